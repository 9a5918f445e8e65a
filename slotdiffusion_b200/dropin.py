"""Run the UNMODIFIED reference (Wuziyi616/SlotDiffusion) with the B200 hot path dropped in.

    python -m slotdiffusion_b200.dropin scripts/train.py --task img_based --params <cfg>.py ...

or, from Python, before the reference builds its model:

    import slotdiffusion_b200.dropin as dropin
    dropin.install()
    model = slotdiffusion.img_based.build_model(params)     # unchanged reference code from here on

install() rebinds the module-level names the reference looks up when it constructs / runs the hot path; no
reference file is edited and every other class of the reference (encoders, VQ-VAE, losses, datasets, nerv
training loop) keeps running as shipped:

  reference name (module global)                                   replaced by
  ---------------------------------------------------------------  ------------------------------------------
  img_based/models/slot_attention.py:15   SlotAttention            slotdiffusion_b200.slot_attention.SlotAttention
  img_based/models/sa_diffusion.py:5,9    SlotAttention(WMask)     ...SlotAttention / SlotAttentionWMask
  video_based/models/savi.py:17           SlotAttention            ...SlotAttention
  video_based/models/savi_diffusion.py:5,10 SlotAttention(WMask)   ...SlotAttention / SlotAttentionWMask
  {img,video}_based/models/ddpm/ddpm.py:24  UNetModel              slotdiffusion_b200.unet.UNetModel
  video_based/models/savi.py:10            TransformerPredictor     slotdiffusion_b200.predictor.TransformerPredictor (fwd + bwd)
  video_based/models/vqvae/VQVAE.py:9      Encoder, Decoder         slotdiffusion_b200.vqvae.Encoder / Decoder (frozen, no-grad)
  img_based/models/slot_attention.py:8, video_based/models/savi.py:8  resnet18, resnet34   slotdiffusion_b200.resnet (fwd + bwd)
  {img,video}_based/models/ddpm/cond_ddpm.py:15  NoiseScheduleVP,  thin adapters (below) that route the one sampler
                                          model_wrapper, DPM_Solver  configuration the repo uses (cond_ddpm.py:155-189)
                                                                    to slotdiffusion_b200.dpm_solver.DPMSolverSampler

Constructor arguments, parameter names / shapes (reference checkpoints load with strict=True) and call signatures
are those of the reference classes; tests/test_dropin_cpu.py checks this against the reference itself.
A sampler request outside the supported configuration (other order / method / guidance, intermediates requested)
is handed to the reference's own DPM_Solver -- reference host code driving the B200 UNet -- with a warning; the
UNet and Slot Attention themselves never fall back: they raise on CPU tensors.
"""
import importlib
import runpy
import sys
import warnings

import torch

_TASKS = ('img_based', 'video_based')
_saved = []          # (module, name, original) for uninstall()
_installed = False


def _rebind(mod, name, new):
    _saved.append((mod, name, getattr(mod, name)))
    setattr(mod, name, new)


class _ModelFn:
    """What the patched model_wrapper returns: the reference closure plus the arguments it was built from."""

    def __init__(self, fn, model, noise_schedule, model_type, guidance_type, condition, guidance_scale,
                 unconditional_condition, model_kwargs):
        self.fn = fn
        self.model = model
        self.noise_schedule = noise_schedule
        self.model_type = model_type
        self.guidance_type = guidance_type
        self.condition = condition
        self.guidance_scale = guidance_scale
        self.unconditional_condition = unconditional_condition
        self.model_kwargs = model_kwargs

    def __call__(self, *args, **kwargs):
        # the reference closure is model_fn(x, t_continuous=None, quantize=False) (dpm_solver.py:386-392): the reference
        # loop calls it as model(x0, None, quantize=True) when vq_denoised (dpm_solver.py:532-533)
        return self.fn(*args, **kwargs)


def _make_adapters(ref_dpm):
    """Adapters around the reference's dpm_solver module (video_based/models/ddpm/dpm_solver.py)."""
    from . import dpm_solver as _ds      # looked up at call time (tests substitute a recording sampler)
    from .unet import UNetModel

    class NoiseScheduleVP(ref_dpm.NoiseScheduleVP):
        """dpm_solver.py:66-235, remembering the betas it was built from (the B200 sampler precomputes its plan from them)."""

        def __init__(self, schedule='discrete', betas=None, alphas_cumprod=None, **kw):
            super().__init__(schedule, betas=betas, alphas_cumprod=alphas_cumprod, **kw)
            self.sdb_betas = betas if schedule == 'discrete' else None

    def model_wrapper(model, noise_schedule, model_type='noise', model_kwargs={}, guidance_type='uncond',
                      condition=None, unconditional_condition=None, guidance_scale=1., **kw):
        """dpm_solver.py:238-416: same closure, with its construction arguments kept alongside."""
        fn = ref_dpm.model_wrapper(model, noise_schedule, model_type=model_type, model_kwargs=model_kwargs,
                                   guidance_type=guidance_type, condition=condition,
                                   unconditional_condition=unconditional_condition, guidance_scale=guidance_scale, **kw)
        return _ModelFn(fn, model, noise_schedule, model_type, guidance_type, condition, guidance_scale,
                        unconditional_condition, model_kwargs)

    class DPM_Solver(ref_dpm.DPM_Solver):
        """dpm_solver.py:419-1328.  sample() runs the CUDA-graph B200 sampler when the request is the repo's own
        configuration (DPM-Solver++ singlestep, orders 1-3, any step count, time_uniform, noise model, guidance scale 1, no x0/xt
        correctors other than vq_denoised); anything else goes through the inherited reference loop."""

        def __init__(self, model_fn, noise_schedule, algorithm_type='dpmsolver++', correcting_x0_fn=None, **kw):
            super().__init__(model_fn, noise_schedule, algorithm_type=algorithm_type,
                             correcting_x0_fn=correcting_x0_fn, **kw)
            self.sdb_model_fn = model_fn

        def _sdb_unsupported(self, steps, order, method, skip_type, t_start, t_end, denoise_to_zero,
                             return_intermediate):
            mf = self.sdb_model_fn
            if not isinstance(mf, _ModelFn):
                return 'model_fn was not built by the patched model_wrapper'
            unet = getattr(mf.model, 'diffusion_model', None)
            if not isinstance(unet, UNetModel):
                return 'the denoiser is not a slotdiffusion_b200.UNetModel'
            if getattr(mf.model, 'conditioning_key', 'crossattn') != 'crossattn':
                return 'conditioning_key != crossattn'
            if getattr(mf.noise_schedule, 'sdb_betas', None) is None:
                return 'noise schedule was not built from discrete betas'
            if mf.model_type != 'noise' or mf.model_kwargs:
                return f'model_type={mf.model_type!r} / model_kwargs'
            if mf.guidance_type != 'classifier-free' or mf.condition is None:     # dpm_solver.py:393-396: cond only
                return f'guidance_type={mf.guidance_type!r}'
            if not (mf.guidance_scale == 1. or mf.unconditional_condition is None):
                return 'classifier-free guidance with scale != 1'
            if self.algorithm_type != 'dpmsolver++' or self.correcting_x0_fn is not None or \
                    self.correcting_xt_fn is not None:
                return 'algorithm_type / correcting functions'
            if method != 'singlestep' or order not in (1, 2, 3) or skip_type != 'time_uniform' or denoise_to_zero:
                return f'method={method!r} order={order} skip_type={skip_type!r}'
            if t_start is not None or t_end is not None or return_intermediate:
                return 't_start / t_end / return_intermediate'
            if self.vq_denoised and getattr(mf.model, 'vae', None) is None:
                return 'vq_denoised without a VQ-VAE on the wrapper'
            return None

        def sample(self, x, steps=20, t_start=None, t_end=None, order=2, skip_type='time_uniform',
                   method='multistep', lower_order_final=True, denoise_to_zero=False, solver_type='dpmsolver',
                   atol=0.0078, rtol=0.05, return_intermediate=False, verbose=False):
            why = self._sdb_unsupported(steps, order, method, skip_type, t_start, t_end, denoise_to_zero,
                                        return_intermediate)
            if why is not None:
                warnings.warn(f'slotdiffusion_b200.dropin: sampler request outside the B200 plan ({why}); '
                              'running the reference DPM_Solver loop around the B200 UNet')
                return super().sample(x, steps=steps, t_start=t_start, t_end=t_end, order=order, skip_type=skip_type,
                                      method=method, lower_order_final=lower_order_final,
                                      denoise_to_zero=denoise_to_zero, solver_type=solver_type, atol=atol, rtol=rtol,
                                      return_intermediate=return_intermediate, verbose=verbose)
            mf = self.sdb_model_fn
            unet = mf.model.diffusion_model
            codebook = None
            if self.vq_denoised:
                # VQVAEWrapper.quantize (VQVAE.py:192-194): Q(h * sf) / sf == nearest row of (codebook / sf)
                vae = mf.model.vae
                emb = vae.vqvae.quantize.embedding.weight.detach()
                codebook = (emb / float(getattr(vae, 'scale_factor', 1.))).float().contiguous()
            # samplers (and their CUDA graphs) live on the UNet they drive, so they die with it
            cache = unet.__dict__.setdefault('_sdb_samplers', {})
            key = (steps, order, bool(self.vq_denoised), str(x.device))
            smp = cache.get(key)
            if smp is None:
                smp = cache[key] = _ds.DPMSolverSampler(unet, mf.noise_schedule.sdb_betas, codebook=codebook,
                                                        steps=steps, order=order)
            if codebook is not None:
                # the VQ-VAE may have been reloaded since the last call; captured graphs hold the buffer's address,
                # so refresh it in place (or drop the graphs if it has to be replaced)
                cur = smp.codebook
                if cur is not None and cur.shape == codebook.shape and cur.device == x.device:
                    cur.copy_(codebook)
                else:
                    smp.codebook = codebook.to(x.device)
                    getattr(smp, '_graphs', {}).clear()
            return smp.sample(x, mf.condition)

    return NoiseScheduleVP, model_wrapper, DPM_Solver


_GRAPH = False


def _graphed_class(cls):
    """Subclass whose instances run their training forward / backward from CUDA graphs (graphed.enable at construction)."""
    from . import graphed

    class Graphed(cls):
        def __init__(self, *a, **kw):
            super().__init__(*a, **kw)
            graphed.enable(self)
    Graphed.__name__, Graphed.__qualname__ = cls.__name__, cls.__qualname__
    return Graphed


def install(tasks=_TASKS, sampler=True, boundary=True, vqvae=True, encoder=True, graph=False, predictor=True):
    """Rebind the reference's hot-path names to the B200 implementations (idempotent).  The reference package
    `slotdiffusion` must be importable (on sys.path / installed).
    boundary: also route q_sample (DDPM._sample_xt_from_x0, ddpm.py:161-165), F.mse_loss of LDM / CondDDPM.loss_function
    (ldm.py:76-77, cond_ddpm.py:205-206) and the eval-time bilinear mask resize of SADiffusion / SAViDiffusion.encode
    (sa_diffusion.py:172-180, savi_diffusion.py:205-213) to csrc/boundary.cu."""
    global _installed
    if _installed:
        return
    from .slot_attention import SlotAttention, SlotAttentionWMask
    from .unet import UNetModel
    from . import boundary as _bd
    SlotAttentionEager, SlotAttentionWMaskEager = SlotAttention, SlotAttentionWMask
    if graph:
        # graph=True: the modules the reference constructs replay their training forward / backward schedules from CUDA
        # graphs (graphed.py) -- for the reference's eager nerv loop, where ~36 ms of Python per un-captured UNet step
        # would otherwise be on the critical path.  isinstance() checks of the sampler adapters still hold (subclasses).
        SlotAttention, SlotAttentionWMask = _graphed_class(SlotAttention), _graphed_class(SlotAttentionWMask)
        UNetModel = _graphed_class(UNetModel)
    if vqvae:
        # the frozen first stage of the LDM (VQVAEWrapper, VQVAE.py:155-194): Encoder / Decoder are looked up as module
        # globals of vqvae/VQVAE.py when VQVAE.__init__ runs (VQVAE.py:9,66-67); img_based re-exports the same module.
        # Inference only -- do not install with vqvae=True to TRAIN a VQ-VAE.
        from . import vqvae as _vq
        m = importlib.import_module('slotdiffusion.video_based.models.vqvae.VQVAE')
        _rebind(m, 'Encoder', _vq.Encoder)
        _rebind(m, 'Decoder', _vq.Decoder)
    ref_dpm = importlib.import_module('slotdiffusion.video_based.models.ddpm.dpm_solver')
    adapters = _make_adapters(ref_dpm) if sampler else None
    for task in tasks:
        base = f'slotdiffusion.{task}.models.'
        sa_mod, wmask_mod = ('slot_attention', 'sa_diffusion') if task == 'img_based' else ('savi', 'savi_diffusion')
        # video models call Slot Attention (and the predictor) once per FRAME before a single backward
        # (savi_diffusion.py:183-196); a graphed callable cannot be replayed twice before its backward, so the recurrent
        # modules of video_based stay launch-by-launch under graph=True (UNet / encoder, called once per step, are graphed)
        sa_cls, sa_wm_cls = (SlotAttention, SlotAttentionWMask) if task == 'img_based' else \
            (SlotAttentionEager, SlotAttentionWMaskEager)
        m = importlib.import_module(base + sa_mod)
        _rebind(m, 'SlotAttention', sa_cls)
        if predictor and hasattr(m, 'TransformerPredictor'):
            from .predictor import TransformerPredictor
            _rebind(m, 'TransformerPredictor', TransformerPredictor)      # savi.py:10, built at savi.py:331-336
        m = importlib.import_module(base + wmask_mod)
        _rebind(m, 'SlotAttention', sa_cls)
        _rebind(m, 'SlotAttentionWMask', sa_wm_cls)
        m = importlib.import_module(base + 'ddpm.ddpm')
        _rebind(m, 'UNetModel', UNetModel)
        if encoder:
            # the image encoder is built by eval(enc_dict['resnet'])(...) in the namespace of slot_attention.py / savi.py
            # (slot_attention.py:8,185-188; savi.py:8,200-206)
            from . import resnet as _rn
            from . import graphed as _gr
            m = importlib.import_module(base + sa_mod)
            if graph:
                _rebind(m, 'resnet18', lambda *a, **kw: _gr.enable(_rn.resnet18(*a, **kw)))
                _rebind(m, 'resnet34', lambda *a, **kw: _gr.enable(_rn.resnet34(*a, **kw)))
            else:
                _rebind(m, 'resnet18', _rn.resnet18)
                _rebind(m, 'resnet34', _rn.resnet34)
        if adapters is not None:
            m = importlib.import_module(base + 'ddpm.cond_ddpm')
            for name, obj in zip(('NoiseScheduleVP', 'model_wrapper', 'DPM_Solver'), adapters):
                _rebind(m, name, obj)
        if boundary:
            proxy = _bd.FunctionalProxy()
            for mod_name in ('ddpm.ldm', 'ddpm.cond_ddpm', wmask_mod):
                m = importlib.import_module(base + mod_name)
                if hasattr(m, 'F'):
                    _rebind(m, 'F', proxy)
            ddpm_cls = importlib.import_module(base + 'ddpm.ddpm').DDPM
            orig_q = ddpm_cls._sample_xt_from_x0

            def _sample_xt_from_x0(self, x0, t, noise=None, _orig=orig_q):
                if noise is None:
                    noise = torch.randn_like(x0)          # same RNG call as the reference's default(...)
                if x0.is_cuda and x0.dtype == torch.float32 and not x0.requires_grad and not noise.requires_grad \
                        and x0[0].numel() % 4 == 0:
                    return _bd.q_sample(x0, t, noise, self.sqrt_alphas_bar, self.sqrt_one_minus_alphas_bar)
                return _orig(self, x0, t, noise)
            _rebind(ddpm_cls, '_sample_xt_from_x0', _sample_xt_from_x0)
    _installed = True


def uninstall():
    """Restore the reference's own classes."""
    global _installed
    while _saved:
        mod, name, orig = _saved.pop()
        setattr(mod, name, orig)
    _installed = False


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    install()
    sys.argv = argv
    runpy.run_path(argv[0], run_name='__main__')
    return 0


if __name__ == '__main__':
    sys.exit(main())
