#!/usr/bin/env python
"""Benchmark of the SlotDiffusion hot path on B200 (see DESIGN.md "measurement").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (BASELINE.json configs[1]): SlotDiffusion img_based LDM, CLEVRTex 128x128, 11 slots --
one STEP = Slot Attention (3 iterations over 32x32x192 features) on a batch, then the 20-NFE
DPM-Solver++ sampling loop of the slot-conditioned UNet (vq_denoised) on that batch.
metric = denoise sample-steps/s = (batch x 20 UNet evaluations) / time, whole job over all ranks.
One process per GPU (torchrun), batch sharded across ranks, no data-path collective (weak scaling).

--impl reference: the reference's own CPU path -- the UNMODIFIED reference from baseline/_ref (kind "reference"; falls
back to the oracle port, kind "port", only if that copy is absent) on the host cores, all threads, on a bounded sample
of the same workload; rank 0 only.
gpu_baseline (N=1): the same unmodified reference, eager PyTorch ON THE SAME B200, same step and batch.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = 'dpm_solver_denoise_steps_per_sec'
UNIT = 'sample-steps/s'
NFE = 20
S, D, N_TOK, SA_ITERS = 11, 192, 1024, 3
# algorithmic work (SURVEY.md 8d / BASELINE.md 2)
UNET_FLOP_PER_SAMPLE = 21.13e9
SA_BYTES_PER_SAMPLE = 848384
SA_WEIGHT_BYTES = 1928448


def load_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return d, 'measured'
    return {'hbm_gbs': 6650.0, 'bf16_tflops': 1590.0, 'bf16_tflops_sustained': 1400.0}, 'fallback'


def gemm_traffic(batch):
    """DRAM bytes per gemm_kernel launch from the committed ncu capture of the same workload (profiles/); None when the
    capture does not match the benchmarked batch."""
    p = os.path.join(ROOT, 'profiles', 'r2c_dram_traffic_per_kernel.json')
    if batch != 256 or not os.path.exists(p):
        return None
    return json.load(open(p))['gemm_kernel_inference_b256']['dram_bytes_per_launch']


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.index), f'--query-gpu={self.Q}',
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in self.rows:
            f = [c.strip() for c in r.split(',')]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        sm.sort()
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx, 'reasons': sorted(reasons),
                'samples': len(sm)}


def build_models(device, seed=0):
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    from slotdiffusion_b200.dpm_solver import DPMSolverSampler
    torch.manual_seed(seed)
    sa = SlotAttentionWMask(D, SA_ITERS, S, D, 2 * D).to(device).eval()
    unet = UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                     attention_resolutions=(8, 4, 2), dropout=0.1, channel_mult=(1, 2, 3, 4), dims=2,
                     use_checkpoint=False, num_head_channels=32, resblock_updown=False, conv_resample=True,
                     transformer_depth=1, context_dim=D, n_embed=None).to(device).eval()
    with torch.no_grad():   # the reference zero-initialises 187 tensors; re-draw them so the work is non-trivial
        for p in unet.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.02)
    betas = (torch.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=torch.float64) ** 2).float()
    codebook = torch.randn(4096, 3, device=device)
    sampler = DPMSolverSampler(unet, betas, codebook=codebook, steps=NFE, use_cuda_graph=True)
    init_slots = torch.randn(1, S, D, device=device)
    return sa, unet, sampler, init_slots


def run_ours(args):
    from slotdiffusion_b200 import _lib, ops
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
    B = args.batch
    sa, unet, sampler, init_slots = build_models(dev, seed=rank)
    g = torch.Generator().manual_seed(1234 + rank)
    feats_h = torch.randn(B, N_TOK, D, generator=g).pin_memory()      # synthetic encoder features (MOVi/CLEVRTex shape)
    noise_h = torch.randn(B, 3, 32, 32, generator=g).pin_memory()
    feats_d, noise_d = feats_h.to(dev), noise_h.to(dev)
    out_h = torch.empty(B, 3, 32, 32).pin_memory()
    slots0 = init_slots.expand(B, -1, -1).contiguous()

    def step_device():
        with torch.no_grad():
            slots, _mask = sa(feats_d, slots0)
            return sampler.sample(noise_d, slots)

    def step_e2e():
        with torch.no_grad():
            f = feats_h.to(dev, non_blocking=True)
            n = noise_h.to(dev, non_blocking=True)
            slots, _mask = sa(f, slots0)
            lat = sampler.sample(n, slots)
            out_h.copy_(lat, non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        return ms

    # launches per step, counted on an un-captured pass (graph replays do not go through the C ABI again)
    sampler.use_cuda_graph = False
    step_device()
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    step_device()
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - n0
    sampler.use_cuda_graph = True

    for _ in range(max(args.warmup, 3)):
        step_device()
    clk = ClockSampler(local)
    clk.start()
    ms = timed(step_device, args.steps)
    clocks = clk.stop()
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, args.steps)

    # the same step at the other batch sizes SURVEY 8d asks for (B = 4: weight-traffic / latency bound; B = 64: the reference's
    # evaluation batch); fewer steps, device-resident inputs
    other = {}
    if world == 1 and not args.no_other_batches:
        for ob in (4, 64):
            if ob == B:
                continue
            f_o, n_o = feats_d[:ob].contiguous(), noise_d[:ob].contiguous()
            s_o = init_slots.expand(ob, -1, -1).contiguous()

            def step_o():
                with torch.no_grad():
                    sl, _ = sa(f_o, s_o)
                    return sampler.sample(n_o, sl)
            for _ in range(3):
                step_o()
            ms_o = timed(step_o, 3)
            other['B%d' % ob] = {'value': ob * NFE * 3 / (ms_o / 1e3), 'unit': UNIT, 'ms_per_step': ms_o / 3,
                                 'images_per_sec': ob * 3 / (ms_o / 1e3)}

    # roofline of the dominant kernel (sdb200 gemm_kernel), measured live with CUDA events around every launch
    roof = gemm_roofline(unet, sampler, B, dev)
    sa_stat = time_sa(sa, feats_d, slots0, load_peaks()[0])

    gpu_base = None
    if world == 1 and not args.no_gpu_baseline:
        try:
            gpu_base = gpu_baseline(dev, B, 0 if args.no_train else args.train_batch)
            gpu_base['speedup_vs_tf32_stock'] = (B * NFE * args.steps / (ms / 1e3)) / gpu_base['tf32_stock']['value'] \
                if 'tf32_stock' in gpu_base else None
            gpu_base['speedup_vs_tf32_all'] = (B * NFE * args.steps / (ms / 1e3)) / gpu_base['tf32_all']['value'] \
                if 'tf32_all' in gpu_base else None
        except Exception as e:                                         # noqa: BLE001
            gpu_base = {'unavailable': repr(e)[:300]}
            torch.cuda.empty_cache()

    train = train_hot = train_video = None
    if not args.no_train:
        del feats_d, noise_d
        torch.cuda.empty_cache()
        train = train_bench(args, dev, world, rank, full=True)
        torch.cuda.empty_cache()
        train_hot = train_bench(args, dev, world, rank, full=False)
        torch.cuda.empty_cache()
        if not args.no_video:
            train_video = train_bench(args, dev, world, rank, full=True, video=6)
            torch.cuda.empty_cache()

    if gpu_base and train and 'train_tf32_stock' in gpu_base:
        gpu_base['train_speedup_vs_tf32_stock'] = train['value'] / gpu_base['train_tf32_stock']['value']
        gpu_base['train_speedup_vs_tf32_all'] = train['value'] / gpu_base['train_tf32_all']['value']
    if rank == 0:
        peaks, peak_src = load_peaks()
        units = B * NFE * world * args.steps
        value = units / (ms / 1e3)
        peak_tf = peaks.get('bf16_tflops_sustained', peaks['bf16_tflops'])
        ach = roof['flops'] / (roof['ms'] / 1e3) / 1e12
        line = {
            'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
            'warmup': max(args.warmup, 3), 'ms_per_step': ms / args.steps, 'higher_is_better': True,
            'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32 (split-fp16 x3 tensor-core products, fp32 accumulate)',
            'data': 'synthetic',
            'config': {
                'workload': 'SlotDiffusion img LDM CLEVRTex 128x128: SlotAttention(3 it, 11 slots, 1024x192 features) '
                            '+ DPM-Solver++ 20 NFE UNet(134M) vq_denoised, per-GPU batch %d' % B,
                'per_gpu_batch': B, 'global_batch': B * world, 'nfe': NFE, 'num_slots': S,
                'parallelism': 'batch-sharded replicas, no data-path collective',
                'l2_policy': 'no flush: per-step working set (1.07 GB packed weights + activations) >> 126 MB L2',
                'nfe_per_sec': NFE * args.steps * world / (ms / 1e3),
                'images_per_sec': B * args.steps * world / (ms / 1e3),
                'slot_attention_ms_B%d' % B: sa_stat['module_ms'],
                'slot_attention_hbm_frac': sa_stat['module_hbm_frac'],
                'slot_attention_persistent_kernel': sa_stat['resident'],
            },
            'clocks': clocks,
            'e2e': {'value': units / (ms_e2e / 1e3), 'unit': UNIT,
                    'h2d_bytes_per_step': feats_h.numel() * 4 + noise_h.numel() * 4,
                    'd2h_bytes_per_step': out_h.numel() * 4},
            'gpu_launches': int(launches_per_step * args.steps),
            'roofline': {'bound': 'tensor', 'achieved': ach, 'peak': peak_tf, 'unit': 'TFLOP/s',
                         'frac': ach / peak_tf, 'traffic': gemm_traffic(B),
                         'traffic_source': 'profiles/r2c_dram_traffic_per_kernel.json: dram__bytes_read.sum + dram__bytes_write.sum '
                                           'per gemm_kernel launch (average over the launches of one UNet evaluation at B=256, ncu)',
                         'kernel': 'sdb::gemm_kernel (tcgen05 kind::f16, 3 MMA passes per algorithmic product)',
                         'peak_source': peak_src + ' bf16 sustained', 'tensor_pipe_frac': 3 * ach / peak_tf,
                         'gemm_share_of_unet_time': roof['ms'] / (ms / args.steps / NFE), 'launches': roof['launches'],
                         'how': 'all %d GEMM launches of one UNet evaluation replayed from a CUDA graph, CUDA events' % roof['launches']},
            'roofline_slot_attention': sa_stat['attend_kernel'],
            'cpu_baseline': cpu_baseline(sample_nfe=NFE, batch=4) if world == 1 else None,   # rank 0, N=1 only
            'gpu_baseline': gpu_base,
            'other_batches': other,
            'train': train,
            'train_hot_modules': train_hot,
            'train_video': train_video,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


class FullImageModel(torch.nn.Module):
    """SADiffusion + LDM of the reference (CLEVRTex config, sa_ldm_clevrtex_params-res128.py) assembled from the B200 modules:
    ResNet18-GN encoder, Slot Attention, frozen VQ-VAE encoder, slot-conditioned UNet, q_sample / eps-MSE kernels.  What
    stays PyTorch is what the reference also runs as a handful of small eager ops between them: SoftPositionEmbed
    (utils.py: x + Linear(4 -> 256)(grid)), the per-token LayerNorm-MLP (slot_attention.py:228-233) and the 1x1 quant_conv
    (VQVAE.py:97-98).  Mirrors SADiffusion.forward + calc_train_loss -> LDM.loss_function (sa_diffusion.py:185-213,
    ldm.py:59-83)."""

    def __init__(self, dev):
        super().__init__()
        from slotdiffusion_b200 import resnet, vqvae
        from slotdiffusion_b200.slot_attention import SlotAttentionWMask
        from slotdiffusion_b200.unet import UNetModel
        nn = torch.nn
        self.encoder = resnet.resnet18(small_inputs=True, use_layer4=False)
        self.pos_dense = nn.Linear(4, 256)
        self.encoder_out_layer = nn.Sequential(nn.LayerNorm(256), nn.Linear(256, D), nn.ReLU(), nn.Linear(D, D))
        self.init_latents = nn.Parameter(torch.randn(1, S, D))
        self.slot_attention = SlotAttentionWMask(D, SA_ITERS, S, D, 2 * D)
        self.vq_encoder = vqvae.Encoder(ch=64, out_ch=3, ch_mult=(1, 2, 4), num_res_blocks=2, attn_resolutions=[], dropout=0.0,
                                        in_channels=3, resolution=128, z_channels=3)
        self.quant_conv = nn.Conv2d(3, 3, 1)
        self.unet = UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                              attention_resolutions=(8, 4, 2), dropout=0.1, channel_mult=(1, 2, 3, 4), dims=2,
                              use_checkpoint=False, num_head_channels=32, resblock_updown=False, conv_resample=True,
                              transformer_depth=1, context_dim=D, n_embed=None)
        for p in list(self.vq_encoder.parameters()) + list(self.quant_conv.parameters()):
            p.requires_grad_(False)                                    # frozen first stage (VQVAE.py:172-176)
        with torch.no_grad():
            for p in self.unet.parameters():
                if p.abs().max() == 0:
                    p.normal_(0, 0.02)
        ys, xs = torch.meshgrid(torch.linspace(0., 1., 32), torch.linspace(0., 1., 32), indexing='ij')
        grid = torch.stack([ys, xs], -1).reshape(1024, 2)
        self.register_buffer('grid', torch.cat([grid, 1. - grid], -1))          # build_grid, utils.py
        betas = (torch.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=torch.float64) ** 2)
        acp = torch.cumprod(1 - betas, 0)
        self.register_buffer('sqrt_abar', acp.sqrt().float())
        self.register_buffer('sqrt_1m_abar', (1 - acp).sqrt().float())
        self.to(dev)

    def trainable(self):
        return [p for p in self.parameters() if p.requires_grad]

    def eager_params(self):
        """parameters whose gradients come from torch autograd (not from a module's own all-reduced flat buffer)"""
        return list(self.pos_dense.parameters()) + list(self.encoder_out_layer.parameters()) + [self.init_latents]

    def loss(self, img):
        from slotdiffusion_b200 import boundary
        B = img.shape[0]
        with torch.no_grad():
            x0 = self.quant_conv(self.vq_encoder(img))                 # ldm.py:62-64
        t = torch.randint(0, 1000, (B,), device=img.device)
        eps = torch.randn_like(x0)
        xt = boundary.q_sample(x0, t, eps, self.sqrt_abar, self.sqrt_1m_abar)   # ddpm.py:161-165
        f = self.encoder(img)                                          # [B, 256, 32, 32]
        f = f + self.pos_dense(self.grid).t().reshape(1, 256, 32, 32)
        feats = self.encoder_out_layer(f.flatten(2).permute(0, 2, 1).contiguous())
        slots, _ = self.slot_attention(feats, self.init_latents.expand(B, -1, -1))
        return boundary.mse_loss(self.unet(xt, t, context=slots), eps)        # ldm.py:76-77


class FullVideoModel(FullImageModel):
    """SAViDiffusion of BASELINE configs[2] (MOVi-D 128x128, T = 6 frames per clip, 11 slots, 2 Slot-Attention iterations per
    frame): the image model applied per frame with the slots carried from frame to frame through the TransformerPredictor
    (savi_diffusion.py:169-216, predictor.py:20-44: 2 layers, 4 heads, ffn 4 D, norm_first -- slotdiffusion_b200.predictor,
    forward + backward on the library's kernels, dropout 0.1 active) and the LDM loss over all B*T frames
    (savi_diffusion.py: flatten(0, 1))."""

    def __init__(self, dev, frames=6, iters=2):
        super().__init__(dev)
        from slotdiffusion_b200.slot_attention import SlotAttentionWMask
        self.frames = frames
        self.slot_attention = SlotAttentionWMask(D, iters, S, D, 2 * D).to(dev)
        from slotdiffusion_b200.predictor import TransformerPredictor
        self.predictor = TransformerPredictor(d_model=D, num_layers=2, num_heads=4, ffn_dim=4 * D, norm_first=True).to(dev)

    def eager_params(self):
        return super().eager_params() + list(self.predictor.parameters())

    def loss(self, video):
        from slotdiffusion_b200 import boundary
        B, T = video.shape[:2]
        img = video.flatten(0, 1)
        with torch.no_grad():
            x0 = self.quant_conv(self.vq_encoder(img))
        t = torch.randint(0, 1000, (B * T,), device=img.device)
        eps = torch.randn_like(x0)
        xt = boundary.q_sample(x0, t, eps, self.sqrt_abar, self.sqrt_1m_abar)
        f = self.encoder(img)
        f = f + self.pos_dense(self.grid).t().reshape(1, 256, 32, 32)
        feats = self.encoder_out_layer(f.flatten(2).permute(0, 2, 1).contiguous()).unflatten(0, (B, T))
        prev, per_frame = None, []
        for k in range(T):                                             # savi_diffusion.py:183-196
            latents = self.init_latents.expand(B, -1, -1) if prev is None else self.predictor(prev)
            prev, _ = self.slot_attention(feats[:, k].contiguous(), latents)
            per_frame.append(prev)
        slots = torch.stack(per_frame, 1).flatten(0, 1)               # [B*T, S, D], clip-major like the frames
        return boundary.mse_loss(self.unet(xt, t, context=slots), eps)


# algorithmic work of the full training step per sample: fwd + bwd (3x) of the trainable modules, fwd of the frozen VQ-VAE
RESNET_FLOP_PER_SAMPLE = 13.45e9
VQENC_FLOP_PER_SAMPLE = 19.5e9
FULL_TRAIN_FLOP_PER_SAMPLE = 3 * (UNET_FLOP_PER_SAMPLE + 203.7e6 + RESNET_FLOP_PER_SAMPLE) + VQENC_FLOP_PER_SAMPLE


def train_bench(args, dev, world, rank, full=True, video=0):
    """One TRAINING step.  full=True: the whole image model of BASELINE configs[1] (FullImageModel: images in, loss out,
    backward through UNet, Slot Attention and the ResNet encoder, data-parallel gradient all-reduce, fused Adam) -- the
    'train-step samples/s' of BASELINE.json.  full=False: the two hot modules alone on synthetic encoder features / latents
    (round-1 number, kept for continuity)."""
    import torch.distributed as dist
    from slotdiffusion_b200 import parallel
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    from slotdiffusion_b200.unet import UNetModel
    B = args.train_batch
    torch.manual_seed(rank)
    g = torch.Generator().manual_seed(4321 + rank)
    if full:
        if video:
            B = args.video_clips
            model = FullVideoModel(dev, frames=video).train()
            img_h = torch.randn(B, video, 3, 128, 128, generator=g).clamp_(-1, 1).pin_memory()
        else:
            model = FullImageModel(dev).train()
            img_h = torch.randn(B, 3, 128, 128, generator=g).clamp_(-1, 1).pin_memory()
        model.vq_encoder.eval()
        params = model.trainable()
        inputs_h = (img_h,)
        wcaches = [model.slot_attention._wcache, model.unet._exec.wc, model.encoder._wc]
    else:
        sa = SlotAttentionWMask(D, SA_ITERS, S, D, 2 * D).to(dev).train()
        unet = UNetModel(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                         attention_resolutions=(8, 4, 2), dropout=0.1, channel_mult=(1, 2, 3, 4), dims=2,
                         use_checkpoint=False, num_head_channels=32, resblock_updown=False, conv_resample=True,
                         transformer_depth=1, context_dim=D, n_embed=None).to(dev).train()
        with torch.no_grad():
            for p in unet.parameters():
                if p.abs().max() == 0:
                    p.normal_(0, 0.02)
        init_slots = torch.nn.Parameter(torch.randn(1, S, D, device=dev))
        params = list(sa.parameters()) + list(unet.parameters()) + [init_slots]
        betas = (torch.linspace(0.0015 ** 0.5, 0.0195 ** 0.5, 1000, dtype=torch.float64) ** 2)
        acp = torch.cumprod(1 - betas, 0).float().to(dev)
        inputs_h = (torch.randn(B, N_TOK, D, generator=g).pin_memory(), torch.randn(B, 3, 32, 32, generator=g).pin_memory())
        wcaches = [sa._wcache, unet._exec.wc]
    if world > 1:
        for p in params:
            dist.broadcast(p.data, 0)
        parallel.enable_grad_allreduce()
    use_graph = not args.no_train_graph
    opt = torch.optim.Adam(params, lr=1e-4, fused=True, capturable=use_graph)
    inputs_d = tuple(t.to(dev) for t in inputs_h)
    from slotdiffusion_b200 import _lib, ops

    def step(*inp):
        ops.dropout_step_counter(dev).add_(1)                        # new dropout masks every step (also under replay)
        if full:
            loss = model.loss(inp[0])
            eager = model.eager_params()
        else:
            feats, x0 = inp
            t = torch.randint(0, 1000, (B,), device=dev)
            eps = torch.randn_like(x0)
            a = acp[t].view(B, 1, 1, 1)
            xt = a.sqrt() * x0 + (1 - a).sqrt() * eps                # q_sample, ddpm.py:161-165 (caller side)
            slots, _ = sa(feats, init_slots.expand(B, -1, -1))
            loss = torch.nn.functional.mse_loss(unet(xt, t, context=slots), eps)
            eager = [init_slots]
        loss.backward()
        if world > 1:                                                # the few torch-autograd parameters (the modules reduce their own)
            for p in eager:
                if p.grad is not None:
                    dist.all_reduce(p.grad, op=dist.ReduceOp.AVG)
        opt.step()
        opt.zero_grad(set_to_none=True)
        return loss

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = tt.item()
        return ms

    # launches per step: counted on eager steps (graph replays do not pass through the C ABI again)
    for _ in range(2):
        step(*inputs_d)
    torch.cuda.synchronize()
    n0 = _lib.launch_count()
    step(*inputs_d)
    torch.cuda.synchronize()
    launches = _lib.launch_count() - n0
    mode = 'eager'
    run = step
    if use_graph:
        # the caller captures the WHOLE step (forward, backward, all-reduce, Adam) in one CUDA graph: every kernel of
        # the library is enqueue-only and allocation-free, weight re-packing is part of the captured step
        try:
            static_in = tuple(t.clone() for t in inputs_d)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(2):
                    step(*static_in)
            torch.cuda.current_stream().wait_stream(side)
            for wc in wcaches:
                wc.clear()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                static_loss = step(*static_in)

            def run(*inp):
                if inp[0] is not static_in[0]:
                    for dst, src in zip(static_in, inp):
                        dst.copy_(src, non_blocking=True)
                graph.replay()
                return static_loss
            run(*static_in)
            torch.cuda.synchronize()
            mode = 'cuda-graph (whole step captured by the caller)'
        except Exception as e:                                         # noqa: BLE001
            print('train: CUDA-graph capture failed, timing the eager step:', repr(e)[:300], file=sys.stderr)
            torch.cuda.synchronize()
            run, mode = step, 'eager'
    for _ in range(max(args.warmup, 3)):
        run(*inputs_d)
    ms = timed(lambda: run(*inputs_d), args.train_steps)
    out = {}

    def e2e_step():
        loss = run(*[t.to(dev, non_blocking=True) for t in inputs_h]) if mode == 'eager' else run(*inputs_h)
        out['loss'] = loss.item()                                    # device -> host read of the step result
    ms_e2e = timed(e2e_step, args.train_steps)
    if world > 1:
        parallel.disable_grad_allreduce()
    units = B * max(video, 1)          # frames for the video model
    fl = units * (FULL_TRAIN_FLOP_PER_SAMPLE if full else (3 * UNET_FLOP_PER_SAMPLE + 3 * 203.7e6))
    what = ('FULL image model (CLEVRTex config): ResNet18-GN encoder fwd+bwd, SlotAttention(3 it), frozen VQ-VAE encoder, '
            'q_sample, UNet(134M) fwd+bwd (dropout 0.1), eps-MSE, gradient all-reduce (bucketed, overlapped), fused Adam; '
            'synthetic images; pos-embed / token MLP / quant_conv are the small PyTorch ops the reference also uses'
            if full else 'SlotAttention(3 it) + UNet(134M) forward+backward (dropout 0.1) + gradient all-reduce + fused Adam; '
            'encoder features / VQ latents synthetic (hot modules only)')
    if video:
        what = ('SAViDiffusion, BASELINE configs[2] (MOVi-D shape): %d clips x %d frames 128x128 per GPU, 11 slots, 2 Slot-Attention '
                'iterations per frame with the slots carried through the TransformerPredictor (B200 kernels, dropout 0.1), ResNet18-GN encoder '
                'fwd+bwd, frozen VQ-VAE encoder, UNet fwd+bwd over all frames, gradient all-reduce, fused Adam; value = frames/s'
                % (B, video))
    return {'metric': 'train_step_frames_per_sec' if video else 'train_step_samples_per_sec',
            'value': units * world * args.train_steps / (ms / 1e3), 'unit': 'frames/s' if video else 'samples/s',
            'ms_per_step': ms / args.train_steps, 'per_gpu_batch': B, 'global_batch': B * world,
            'frames_per_clip': video or None,
            'e2e_value': units * world * args.train_steps / (ms_e2e / 1e3),
            'h2d_bytes_per_step': sum(t.numel() for t in inputs_h) * 4, 'd2h_bytes_per_step': 4,
            'gpu_launches_per_step': launches, 'loss': out.get('loss'), 'mode': mode,
            'algorithmic_tflops': fl * args.train_steps / (ms / 1e3) / 1e12,
            'what': what}


def profile_once(args):
    """ncu helper (use with --profile-from-start off): warm-up passes outside the profiled range, then ONE
    SlotAttention forward + ONE un-captured UNet evaluation between cudaProfilerStart/Stop."""
    dev = torch.device('cuda', 0)
    B = args.batch
    sa, unet, sampler, init_slots = build_models(dev)
    feats = torch.randn(B, N_TOK, D, device=dev)
    x = torch.randn(B, 3, 32, 32, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    s0 = init_slots.expand(B, -1, -1).contiguous()
    with torch.no_grad():
        for _ in range(2):
            slots, _ = sa(feats, s0)
            unet(x, t, context=slots)
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        slots, _ = sa(feats, s0)
        unet(x, t, context=slots)
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()


def profile_train_once(args):
    """ncu helper (use with --profile-from-start off): ONE eager training step between cudaProfilerStart/Stop."""
    dev = torch.device('cuda', 0)
    args.no_train_graph = True
    args.train_steps = 1
    args.warmup = 3
    import slotdiffusion_b200.backward as bw
    real_run = bw.Tape.run
    state = {'n': 0}
    orig_train = train_bench

    # profile the LAST timed step only: start the profiler when the timed region begins
    real_event_record = torch.cuda.Event.record
    counter = {'rec': 0}

    def rec(self, *a, **k):
        counter['rec'] += 1
        if counter['rec'] == 1:
            torch.cuda.synchronize()
            torch.cuda.profiler.start()
        r = real_event_record(self, *a, **k)
        if counter['rec'] == 2:
            torch.cuda.synchronize()
            torch.cuda.profiler.stop()
        return r
    torch.cuda.Event.record = rec
    try:
        orig_train(args, dev, 1, 0)
    finally:
        torch.cuda.Event.record = real_event_record


def gemm_roofline(unet, sampler, B, dev):
    """Device time of the dominant kernel: every sdb_gemm call of ONE UNet evaluation is recorded (operands kept
    alive), then exactly those launches are captured in a CUDA graph and replayed back to back, timed with CUDA
    events on the launching stream -- no host launch latency, no other kernels in the interval."""
    from slotdiffusion_b200 import ops
    real = ops.gemm
    calls = []

    def recording_gemm(a, w, *pa, **kw):
        out = real(a, w, *pa, **kw)
        conv = kw.get('conv')
        M = a.rows if conv is None else conv[1] * conv[2] * conv[3]
        kw2 = dict(kw)
        if kw2.get('gsum') is not None:
            kw2['gsum'] = torch.zeros_like(kw2['gsum'])     # private copy: replays must not touch the live arena
        calls.append((a, w, pa, kw2, 2.0 * M * w.rows * w.K,
                      (M, w.rows, w.K, 'conv' if conv else 'lin', 'res' if kw.get('residual') is not None else '',
                       'geglu' if kw.get('geglu') else '')))
        return out
    x = torch.randn(B, 3, 32, 32, device=dev)
    t = torch.randint(0, 1000, (B,), device=dev)
    ctx = torch.randn(B, S, D, device=dev)
    with torch.no_grad():
        unet(x, t, context=ctx)
        torch.cuda.synchronize()
        ops.gemm = recording_gemm
        try:
            unet(x, t, context=ctx)
        finally:
            ops.gemm = real
        torch.cuda.synchronize()

        def replay(sel):
            for a, w, pa, kw, _, _ in sel:
                real(a, w, *pa, **kw)

        def graph_ms(sel, reps=5):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                replay(sel)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                replay(sel)
            g.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                g.replay()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        ms = graph_ms(calls)
        if os.environ.get('SDB_GEMM_TABLE'):
            import collections
            agg = collections.OrderedDict()
            for c in calls:
                agg.setdefault(c[5], []).append(c)
            rows = []
            for k, sel in agg.items():
                rows.append((k, len(sel), graph_ms(sel, 3) * 1e3, sum(c[4] for c in sel)))
            print('%-44s %4s %9s %8s %9s' % ('M,N,K,kind', 'n', 'total us', 'avg us', 'alg TF/s'), file=sys.stderr)
            for k, n, us, fl in sorted(rows, key=lambda r: -r[2]):
                print('%-44s %4d %9.1f %8.1f %9.1f' % (str(k), n, us, us / n, fl / us / 1e6), file=sys.stderr)
    return {'ms': ms, 'flops': sum(c[4] for c in calls), 'launches': len(calls)}


def _graph_us(fn, reps=10, flush=None):
    """Median device time of fn() replayed from a CUDA graph (optionally after evicting L2)."""
    fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def time_sa(sa, feats, slots0, peaks):
    """Slot Attention: whole-module forward (CUDA-graph replay, features evicted from L2 before every replay) and the
    fused attend kernel alone (one iteration) against the HBM roofline."""
    from slotdiffusion_b200 import ops
    B = feats.shape[0]
    flush = torch.empty(64 * 1024 * 1024, device=feats.device)       # 256 MB > 126 MB L2
    with torch.no_grad():
        mod_us = _graph_us(lambda: sa(feats, slots0), flush=flush)
        qa = torch.randn(B * S, D + 4, device=feats.device) * D ** -0.5
        att_us = _graph_us(lambda: ops.slot_attend_fused(feats, qa, B, N_TOK, S, D, 1e-5, 1e-6, True), flush=flush)
    it_bytes = B * 4 * (N_TOK * D + S * N_TOK + 2 * S * D)            # features in, seg mask + partial updates out
    mod_bytes = B * SA_BYTES_PER_SAMPLE + SA_WEIGHT_BYTES
    # the persistent cluster kernel (whole forward in ONE launch, features read from HBM once): timed at one full wave
    # of resident clusters, the largest batch the module routes to it (autograd.RESIDENT_WAVES)
    resident = None
    wave = ops.slot_attention_resident_wave(N_TOK, S, D, D, 2 * D)
    if wave > 0:
        bw = min(wave, B)
        with torch.no_grad():
            res_us = _graph_us(lambda: sa(feats[:bw], slots0[:bw]), flush=flush)
        res_bytes = bw * SA_BYTES_PER_SAMPLE + SA_WEIGHT_BYTES
        resident = {'batch': bw, 'samples_per_wave': wave, 'us': res_us, 'launches': 1,
                    'hbm_frac': res_bytes / (res_us * 1e-6) / 1e9 / peaks['hbm_gbs'],
                    'kernel': 'sdb::sr::slot_attention_resident_kernel (3 iterations + GRU/MLP update, one launch, L2 flushed)'}
    return {'module_ms': mod_us / 1e3,
            'module_hbm_frac': mod_bytes / (mod_us * 1e-6) / 1e9 / peaks['hbm_gbs'],
            'resident': resident,
            'attend_kernel': {'bound': 'hbm', 'us': att_us, 'achieved': it_bytes / (att_us * 1e-6) / 1e9,
                              'peak': peaks['hbm_gbs'], 'unit': 'GB/s',
                              'frac': it_bytes / (att_us * 1e-6) / 1e9 / peaks['hbm_gbs'],
                              'kernel': 'sdb::slot_attend_fused_kernel (+finalize), one iteration, L2 flushed',
                              'algorithmic_bytes': it_bytes}}


def cpu_step(batch, nfe, seed=0, state={}):
    """One bounded step of the same workload through the CPU oracle (port of the reference)."""
    from oracle import dpm_ref, unet_ref
    from oracle import slot_attention_ref as sa_ref
    if 'sd' not in state:
        state['sd'] = unet_ref.random_state_dict(seed=seed)
        state['p'] = sa_ref.random_params(D, D, 2 * D, seed=seed)
        state['betas'] = dpm_ref.ddpm_buffers(dpm_ref.linear_betas())['betas']
        state['cb'] = torch.randn(4096, 3)
    g = torch.Generator().manual_seed(seed)
    feats = torch.randn(batch, N_TOK, D, generator=g)
    slots0 = torch.randn(1, S, D, generator=g).expand(batch, -1, -1)
    xT = torch.randn(batch, 3, 32, 32, generator=g)
    with torch.no_grad():
        slots, _ = sa_ref.slot_attention_forward(state['p'], feats, slots0, SA_ITERS)
        return dpm_ref.dpm_sample(lambda x, t, c: unet_ref.unet_forward(state['sd'], x, t, c), state['betas'], xT,
                                  slots, state['cb'], steps=nfe)


def cpu_baseline(sample_nfe, batch):
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = load_reference_model(torch.device('cpu')) if sample_nfe == NFE else None
    if model is not None:              # the reference's own code on the host cores
        g = torch.Generator().manual_seed(0)
        feats = torch.randn(batch, N_TOK, D, generator=g)
        slots0 = torch.randn(1, S, D, generator=g).expand(batch, -1, -1).contiguous()
        with torch.no_grad():
            model.slot_attention(feats, slots0)          # warm-up (thread pool)
        fn, kind, what = (lambda: reference_step(model, feats, slots0, batch)), 'reference', \
            'unmodified reference (baseline/_ref), torch CPU fp32'
    else:
        cpu_step(batch, 2)             # warm-up (weight generation, thread pool)
        fn, kind, what = (lambda: cpu_step(batch, sample_nfe)), 'port', \
            'oracle (torch CPU fp32 restatement of the reference)'
    runs, dt = 0, 0.0
    t0 = time.perf_counter()
    while dt < 10.0 and runs < 8:      # bounded sample: ~10-30 s of host work
        fn()
        runs += 1
        dt = time.perf_counter() - t0
    return {'value': runs * batch * sample_nfe / dt, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': f'{what}: SlotAttention + {sample_nfe}-NFE DPM-Solver++ at batch {batch}, {runs} run(s), {dt:.1f} s'}


def load_reference_model(device):
    """The UNMODIFIED reference SADiffusion (CLEVRTex config) from baseline/_ref (baseline/install_ref.sh; travels to the
    GPU box) with tools/ref_stubs for its un-vendored imports, or None when that copy is absent."""
    sys.path.insert(0, os.path.join(ROOT, 'tools'))
    import ref_import
    if not ref_import.box_copy_available():
        return None
    import contextlib
    import io
    import warnings
    ref_import.use_box_copy()
    torch.manual_seed(0)
    with warnings.catch_warnings(), contextlib.redirect_stdout(io.StringIO()):
        warnings.simplefilter('ignore')
        model = ref_import.img_models().build_model(
            ref_import.fresh_params('img_based', 'sa_ldm/sa_ldm_clevrtex_params-res128.py'))
    with torch.no_grad():
        for p in model.parameters():
            if p.abs().max() == 0:
                p.normal_(0, 0.02)
    return model.to(device).eval()


def reference_step(model, feats, slots0, batch):
    """The same step through the reference's own code: SlotAttentionWMask.forward (sa_diffusion.py:15-70) and
    dm_decoder.generate_imgs(use_dpm=True) (cond_ddpm.py:155-189: DPM-Solver++ singlestep order 3, 20 NFE, vq_denoised)."""
    with torch.no_grad():
        slots, _ = model.slot_attention(feats, slots0)
        return model.dm_decoder.generate_imgs(cond=slots, batch_size=batch, use_dpm=True, verbose=False)


def gpu_baseline(dev, batch, train_batch=0):
    """The honest bar: the reference's eager PyTorch path on THIS GPU, same step, same batch; PyTorch's stock
    TF32 policy (cuDNN convolutions TF32, matmul fp32) and TF32 everywhere (the fastest setting a user can pick)."""
    model = load_reference_model(dev)
    if model is None:
        return {'unavailable': 'baseline/_ref missing (run baseline/install_ref.sh in the build container)'}
    g = torch.Generator().manual_seed(99)
    feats = torch.randn(batch, N_TOK, D, generator=g).to(dev)
    slots0 = torch.randn(1, S, D, generator=g).to(dev).expand(batch, -1, -1).contiguous()
    out = {'what': 'unmodified reference (baseline/_ref), eager PyTorch %s on the same GPU: SlotAttentionWMask.forward + '
                   'generate_imgs(use_dpm=True), batch %d, inputs resident, 1 warm-up + 2 timed steps' % (torch.__version__, batch),
           'unit': UNIT}
    saved = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    for name, mm, cd in (('tf32_stock', False, True), ('tf32_all', True, True)):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = mm, cd
        reference_step(model, feats, slots0, batch)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(2):
            reference_step(model, feats, slots0, batch)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 2
        out[name] = {'value': batch * NFE / (ms / 1e3), 'ms_per_step': ms}
    if train_batch > 0:
        # the reference's own full training step (SADiffusion.forward -> calc_train_loss -> backward, Adam over every
        # trainable parameter; img_based/method.py, nerv trainer: eager, no graph capture), same batch as our `train` block
        model.train()
        img = torch.randn(train_batch, 3, 128, 128, generator=g).clamp_(-1, 1).to(dev)
        opt = torch.optim.Adam([p for p in model.parameters() if p.requires_grad], lr=1e-4)

        def ref_train_step():
            data = {'img': img}
            loss = model.calc_train_loss(data, model(data))['denoise_loss']
            opt.zero_grad(set_to_none=True)
            loss.backward()
            opt.step()
        for name, mm, cd in (('train_tf32_stock', False, True), ('train_tf32_all', True, True)):
            torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = mm, cd
            for _ in range(2):
                ref_train_step()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ref_train_step()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            out[name] = {'value': train_batch / (ms / 1e3), 'unit': 'samples/s', 'ms_per_step': ms, 'per_gpu_batch': train_batch}
        del opt
    torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = saved
    del model
    torch.cuda.empty_cache()
    return out


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    batch = 4   # bounded sample of the workload (BASELINE configs[0]: the reference's CPU-runnable batch)
    steps = max(1, min(args.steps, 3))
    warm = 1 if args.warmup > 0 else 0
    model = load_reference_model(torch.device('cpu'))
    kind = 'reference' if model is not None else 'port'
    if model is not None:
        g = torch.Generator().manual_seed(0)
        feats = torch.randn(batch, N_TOK, D, generator=g)
        slots0 = torch.randn(1, S, D, generator=g).expand(batch, -1, -1).contiguous()
        dec = model.dm_decoder
        if warm:      # one 2-evaluation pass: thread pool, allocator
            with torch.no_grad():
                model.slot_attention(feats, slots0)
                dec.model.diffusion_model(torch.randn(batch, 3, 32, 32), torch.full((batch,), 500.), context=slots0)
        t0 = time.perf_counter()
        for _ in range(steps):
            reference_step(model, feats, slots0, batch)
        dt = time.perf_counter() - t0
    else:
        for _ in range(warm):
            cpu_step(batch, 2)
        t0 = time.perf_counter()
        for _ in range(steps):
            cpu_step(batch, NFE)
        dt = time.perf_counter() - t0
    value = batch * NFE * steps / dt
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': args.gpus, 'steps': steps,
        'warmup': warm, 'ms_per_step': dt / steps * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'SlotDiffusion img LDM CLEVRTex 128x128: SlotAttention(3 it, 11 slots, 1024x192 features) + '
                               'DPM-Solver++ 20 NFE UNet(134M) vq_denoised through %s on the host cores, bounded sample: '
                               'batch %d per step, %d step(s) (the B200 arm runs the same step at per-GPU batch 256)'
                               % ('the UNMODIFIED reference (baseline/_ref: SlotAttentionWMask.forward + '
                                  'generate_imgs(use_dpm=True))' if kind == 'reference'
                                  else 'the CPU port of the reference (oracle/)', batch, steps),
                   'per_gpu_batch': batch, 'nfe': NFE, 'num_slots': S, 'same_workload_smaller_batch': True},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': kind,
                         'sample': f'{steps} step(s) of batch {batch} x {NFE} NFE'},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='per-GPU batch of the sampling workload')
    ap.add_argument('--train-batch', type=int, default=64, help='per-GPU batch of the training-step measurement')
    ap.add_argument('--train-steps', type=int, default=5)
    ap.add_argument('--no-train', action='store_true', help='skip the training-step measurement')
    ap.add_argument('--no-video', action='store_true', help='skip the video-model (BASELINE configs[2]) training-step line')
    ap.add_argument('--video-clips', type=int, default=8, help='per-GPU clips of the video training step (x 6 frames)')
    ap.add_argument('--no-train-graph', action='store_true', help='time the eager training step (no CUDA graph)')
    ap.add_argument('--no-other-batches', action='store_true', help='skip the B = 4 / B = 64 sampling lines')
    ap.add_argument('--no-gpu-baseline', action='store_true', help='skip timing the unmodified reference on the same GPU')
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--profile-train-once', action='store_true', help='ncu helper: one eager training step')
    ap.add_argument('--profile-once', action='store_true',
                    help='ncu helper: SlotAttention + ONE un-captured UNet evaluation (after one warm-up pass), no timing')
    args = ap.parse_args()
    if args.profile_train_once:
        profile_train_once(args)
    elif args.profile_once:
        profile_once(args)
    elif args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
