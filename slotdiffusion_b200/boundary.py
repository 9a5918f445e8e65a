"""Fusions at the boundary of the hot path, behind the reference's own call sites:

  q_sample(x0, t, noise, sqrt_abar, sqrt_1m_abar)   DDPM._sample_xt_from_x0          ddpm.py:161-165
  mse_loss(pred, target)                            F.mse_loss in LDM.loss_function  ldm.py:76-77 (autograd: d pred only)
  mask_upsample(masks, size) / mask_argmax(...)     F.interpolate(bilinear) in SADiffusion.encode (sa_diffusion.py:172-180,
                                                    savi_diffusion.py:205-213) and masks.argmax(-3) of test_seg.py:27

CUDA only (csrc/boundary.cu); CPU tensors raise.  dropin.install() routes the reference's calls here.
"""
import ctypes

import torch

from ._lib import check, lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _cuda_f32(*ts):
    for t in ts:
        if not (t.is_cuda and t.dtype == torch.float32):
            raise RuntimeError('slotdiffusion_b200.boundary: CUDA float32 tensors required (no CPU fallback), got '
                               f'{t.dtype} on {t.device}')


def q_sample(x0, t, noise, sqrt_abar, sqrt_1m_abar):
    """x_t = sqrt(abar_t) x0 + sqrt(1 - abar_t) noise; x0 / noise [B, ...], t [B] int64; bit-identical to the eager form.
    No gradient flows to x0 (the VQ-VAE latents are detached, ldm.py:62-64)."""
    _cuda_f32(x0, noise, sqrt_abar, sqrt_1m_abar)
    B = x0.shape[0]
    n = x0[0].numel()
    x0c, nc = x0.detach().contiguous(), noise.detach().contiguous()
    out = torch.empty_like(x0c)
    check(lib().sdb_q_sample(_p(x0c), _p(nc), _p(t.to(torch.int64).contiguous()), _p(sqrt_abar.contiguous()),
                             _p(sqrt_1m_abar.contiguous()), _p(out), B, n, _stream()), 'sdb_q_sample')
    return out


class _MSELoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, target):
        p, t = pred.detach().contiguous(), target.detach().contiguous()
        n = p.numel()
        diff = torch.empty_like(p) if pred.requires_grad else None
        loss = torch.empty(1, dtype=torch.float32, device=p.device)
        check(lib().sdb_mse_loss_fwd(_p(p), _p(t), _p(diff), _p(loss), n, _stream()), 'sdb_mse_loss_fwd')
        ctx.diff, ctx.shape = diff, pred.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, gout):
        if ctx.diff is None:
            return None, None
        dp = torch.empty_like(ctx.diff)
        g = gout.detach().reshape(1).float().contiguous()
        check(lib().sdb_mse_loss_bwd(_p(ctx.diff), _p(g), _p(dp), dp.numel(), _stream()), 'sdb_mse_loss_bwd')
        return dp.view(ctx.shape), None


def mse_loss(pred, target):
    """mean((pred - target)^2) with the gradient w.r.t. pred (the target is the sampled noise / detached latents)."""
    _cuda_f32(pred, target)
    if target.requires_grad:
        raise RuntimeError('slotdiffusion_b200.boundary.mse_loss: the target must not require grad')
    return _MSELoss.apply(pred, target)


def mask_upsample(masks, size, want_up=True, want_argmax=False):
    """masks [B, S, h, w] -> (bilinear resize to `size` (align_corners=False) [B, S, H, W] | None, argmax over S [B, H, W]
    int64 | None) in one pass."""
    _cuda_f32(masks)
    B, S, h, w = masks.shape
    H, W = size
    m = masks.detach().contiguous()
    up = torch.empty(B, S, H, W, dtype=torch.float32, device=m.device) if want_up else None
    idx = torch.empty(B, H, W, dtype=torch.int64, device=m.device) if want_argmax else None
    check(lib().sdb_mask_upsample_argmax(_p(m), _p(up), _p(idx), B, S, h, w, H, W, _stream()), 'sdb_mask_upsample_argmax')
    return up, idx


class FunctionalProxy:
    """Stand-in for `torch.nn.functional` inside the reference modules that call F.mse_loss / F.interpolate on the hot path
    boundary (rebound by dropin.install()): routes exactly those two call shapes to the kernels above and everything else to
    torch.nn.functional."""

    def __init__(self):
        import torch.nn.functional as F
        object.__setattr__(self, '_F', F)

    def __getattr__(self, name):
        return getattr(self._F, name)

    def mse_loss(self, input, target, *a, **kw):
        if (not a and not kw and input.is_cuda and input.dtype == torch.float32 and target.dtype == torch.float32
                and input.shape == target.shape and not target.requires_grad and input.numel() % 4 == 0):
            return mse_loss(input, target)
        return self._F.mse_loss(input, target, *a, **kw)

    def interpolate(self, input, size=None, scale_factor=None, mode='nearest', align_corners=None, **kw):
        if (mode == 'bilinear' and align_corners is False and size is not None and scale_factor is None and not kw
                and input.dim() == 4 and input.is_cuda and input.dtype == torch.float32 and not input.requires_grad):
            size = (size, size) if isinstance(size, int) else tuple(size)
            # the reference flattens [B, S, h, w] to [B*S, 1, h, w] first (sa_diffusion.py:173): any (N, C) split works
            N, C = input.shape[:2]
            return mask_upsample(input.reshape(1, N * C, *input.shape[2:]), size)[0].view(N, C, *size)
        return self._F.interpolate(input, size=size, scale_factor=scale_factor, mode=mode, align_corners=align_corners, **kw)
