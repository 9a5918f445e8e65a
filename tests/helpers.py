"""Shared test helpers: seeded inputs identical to tools/make_golden.py."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def seeded(shape, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g, dtype=dtype)


def checksum(t):
    t = t.detach().double().cpu().flatten()
    return np.array([t.sum().item(), (t * t).sum().item(), t[0].item(), t[-1].item()])


def golden(name):
    return np.load(os.path.join(GOLDEN, name + '.npz'))


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def argmax_mismatch(mask, argmax_ref, margin_ref, tie_margin):
    """Count argmax disagreements; split into those on near-ties (reference top-2
    margin < tie_margin, where fp32 round-off legitimately decides) and real ones."""
    am = torch.as_tensor(mask).argmax(1).cpu().numpy().astype(np.int64)
    bad = am != np.asarray(argmax_ref).astype(np.int64)
    near = np.asarray(margin_ref) < tie_margin
    return int((bad & ~near).sum()), int((bad & near).sum())


SA_CASES = {
    'sa_img_clevrtex': (2, 1024, 192, 11, 192, 384, 3),
    'sa_vid_movid': (2, 1024, 192, 15, 192, 384, 2),
    'sa_movie_24slots': (1, 1024, 192, 24, 192, 384, 2),
    'sa_coco_vitb16': (2, 196, 256, 7, 256, 512, 3),
    'sa_ragged_small': (3, 77, 192, 5, 192, 384, 1),
}


def sa_case(name):
    from oracle import slot_attention_ref as sa_ref
    B, N, Din, S, D, M, I = SA_CASES[name]
    p = sa_ref.random_params(Din, D, M, seed=11)
    x = seeded((B, N, Din), 21)
    s0 = seeded((B, S, D), 22)
    gw = seeded((B, S, D), 23)
    return p, x, s0, gw, I


def vq_mismatch(idx, z, codebook, dz):
    """Nearest-code parity with near-tie accounting (same idea as argmax_mismatch).  idx: indices under test [P];
    z: the reference pre-quantisation latents [P, C]; codebook [K, C]; dz: per-pixel (or scalar) bound on how far z may
    legitimately move (the fp32 round-off of whatever produced z).  In fp64: a pixel whose distance to the bisector
    plane between its best and second-best code, (d2^2 - d1^2) / (2 |e2 - e1|), is below dz is a near-tie -- round-off
    decides it.  Returns (real, near, ref_idx): disagreements outside / inside near-ties."""
    z = torch.as_tensor(z).double().cpu()
    e = torch.as_tensor(codebook).double().cpu()
    d = torch.cdist(z, e) ** 2
    two = d.topk(2, dim=1, largest=False)
    ref = two.indices[:, 0]
    sep = (e[two.indices[:, 0]] - e[two.indices[:, 1]]).norm(dim=1).clamp_min(1e-300)
    bis = (two.values[:, 1] - two.values[:, 0]) / (2 * sep)
    bad = torch.as_tensor(idx).long().cpu().flatten() != ref
    near = bis < torch.as_tensor(dz).double()
    return int((bad & ~near).sum()), int((bad & near).sum()), ref
