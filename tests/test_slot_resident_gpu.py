"""GPU parity of the persistent Slot-Attention kernel (csrc/slot_attention_resident.cu: whole forward in one launch, features
resident in shared memory) against the reference goldens, the fp64 oracle and the per-iteration kernel path."""
import pytest
import torch

from slotdiffusion_b200 import autograd, ops

from helpers import SA_CASES, argmax_mismatch, golden, rel_l2, sa_case, seeded
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TIGHT = 5e-5


def make_sa(name):
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, gw, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M).cuda()
    mod.load_state_dict(p)
    return mod, p, x, s0


def run(mod, x, s0, resident):
    saved = autograd.RESIDENT, autograd.RESIDENT_WAVES
    autograd.RESIDENT, autograd.RESIDENT_WAVES = resident, 1 << 20        # whatever the batch size
    try:
        with torch.no_grad():
            return mod(x.cuda(), s0.cuda())
    finally:
        autograd.RESIDENT, autograd.RESIDENT_WAVES = saved


@pytest.mark.parametrize('name', list(SA_CASES))
def test_resident_matches_reference_golden(name):
    B, N, Din, S, D, M, I = SA_CASES[name]
    assert ops.slot_attention_resident_supported(N, S, Din, D, M)
    g = golden(name)
    mod, p, x, s0 = make_sa(name)
    slots, mask = run(mod, x, s0, True)
    assert rel_l2(slots, g['slots']) < TIGHT
    assert rel_l2(mask, g['mask']) < TIGHT
    assert rel_l2(slots, g['slots64']) < TIGHT
    real, near = argmax_mismatch(mask, g['argmax64'], g['margin64'], 1e-5)
    assert real == 0, (real, near)
    s_old, m_old = run(mod, x, s0, False)
    assert rel_l2(slots, s_old) < 1e-5 and rel_l2(mask, m_old) < 1e-5


@pytest.mark.parametrize('B', [1, 3, 37, 64])
def test_resident_batch_sweep_matches_per_iteration_path(B):
    """persistent loop: more samples than resident clusters, odd batch sizes; every sample equals the per-iteration path"""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    N, S, D = 1024, 11, 192
    p = sa_ref.random_params(D, D, 2 * D, seed=5)
    mod = SlotAttentionWMask(D, 3, S, D, 2 * D).cuda()
    mod.load_state_dict(p)
    x, s0 = seeded((B, N, D), 61), seeded((B, S, D), 62)
    slots, mask = run(mod, x, s0, True)
    assert torch.isfinite(slots).all()
    assert (mask.sum(1) - 1).abs().max().item() < 1e-5
    s_old, m_old = run(mod, x, s0, False)
    assert rel_l2(slots, s_old) < 1e-5 and rel_l2(mask, m_old) < 1e-5
    assert (mask.argmax(1) != m_old.argmax(1)).sum().item() <= B
    # batch independence: bit-identical whatever the batch around a sample (a cluster never mixes samples)
    s1, m1 = run(mod, x[B - 1:], s0[B - 1:], True)
    assert torch.equal(s1, slots[B - 1:]) and torch.equal(m1, mask[B - 1:])
    nb = min(B, 2)
    ref_s, ref_m = sa_ref.slot_attention_forward(p, x[:nb].double(), s0[:nb].double(), 3)
    assert rel_l2(slots[:nb], ref_s) < TIGHT
    margin = ref_m.topk(2, dim=1).values
    real, near = argmax_mismatch(mask[:nb], ref_m.argmax(1).numpy(), (margin[:, 0] - margin[:, 1]).numpy(), 1e-5)
    assert real == 0


def test_resident_without_mask_and_rerun_is_deterministic():
    from slotdiffusion_b200.slot_attention import SlotAttention
    N, S, D = 1024, 11, 192
    p = sa_ref.random_params(D, D, 2 * D, seed=7)
    mod = SlotAttention(D, 3, S, D, 2 * D).cuda()
    mod.load_state_dict(p)
    x, s0 = seeded((5, N, D), 71).cuda(), seeded((5, S, D), 72).cuda()
    assert ops.slot_attention_resident_wave(N, S, D, D, 2 * D) >= 5        # B = 5 takes the persistent kernel by default
    with torch.no_grad():
        a = mod(x, s0)
        b = mod(x, s0)
    assert torch.equal(a, b)
    ref_s, _ = sa_ref.slot_attention_forward(p, x[:2].double().cpu(), s0[:2].double().cpu(), 3)
    assert rel_l2(a[:2], ref_s) < TIGHT
