"""GPU parity of the one-launch slot update (sdb_slot_update + sdb_slot_attend_fused_partials, opt-in path
SDB_SA_FUSED_TAIL=1).  Marker gpu_next: written after the round's GPU minutes were spent; its arithmetic is already
checked through the host emulation (tests/test_slot_update_emulation_cpu.py); this file adds the launch itself."""
import pytest
import torch

from helpers import SA_CASES, argmax_mismatch, golden, rel_l2, sa_case, seeded
from oracle import slot_attention_ref as sa_ref

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TIGHT = 5e-5


@pytest.fixture
def fused_tail(monkeypatch):
    from slotdiffusion_b200 import autograd
    monkeypatch.setattr(autograd, 'FUSED_TAIL', True)
    return autograd


def _module(name):
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, gw, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M).cuda()
    mod.load_state_dict(p)
    return mod, p, x, s0, iters


@pytest.mark.parametrize('name', list(SA_CASES))
def test_fused_tail_module_matches_reference_golden(fused_tail, name):
    g = golden(name)
    mod, p, x, s0, iters = _module(name)
    with torch.no_grad():
        slots, mask = mod(x.cuda(), s0.cuda())
    assert rel_l2(slots, g['slots']) < TIGHT and rel_l2(mask, g['mask']) < TIGHT
    assert rel_l2(slots, g['slots64']) < TIGHT
    real, near = argmax_mismatch(mask, g['argmax64'], g['margin64'], 1e-5)
    assert real == 0, (real, near)


def test_fused_tail_agrees_with_the_gemm_tail(fused_tail):
    """same module, both tails, B = 64 (rows = 704: 8-row tiles) and B = 4 (4-row tiles)"""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    N, S, D = 1024, 11, 192
    p = sa_ref.random_params(D, D, 2 * D, seed=5)
    mod = SlotAttentionWMask(D, 3, S, D, 2 * D).cuda()
    mod.load_state_dict(p)
    for B in (64, 4):
        x, s0 = seeded((B, N, D), 61).cuda(), seeded((B, S, D), 62).cuda()
        with torch.no_grad():
            a_s, a_m = mod(x, s0)
            fused_tail.FUSED_TAIL = False
            b_s, b_m = mod(x, s0)
            fused_tail.FUSED_TAIL = True
        assert rel_l2(a_s, b_s) < 1e-5 and rel_l2(a_m, b_m) < 1e-5
        ref_s, _ = sa_ref.slot_attention_forward(p, x[:2].cpu().double(), s0[:2].cpu().double(), 3)
        assert rel_l2(a_s[:2], ref_s) < TIGHT


def test_slot_update_op_matches_the_host_emulation_math():
    """sdb_slot_update against fp64 torch math of the same folded formulas, q-only and update+q calls"""
    from slotdiffusion_b200 import ops
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, S, Din, D, M, chunks = 5, 7, 256, 256, 512, 3
    mod = SlotAttentionWMask(Din, 3, S, D, M).cuda()
    mod.load_state_dict(sa_ref.random_params(Din, D, M, seed=3))
    w = mod._wcache.slot_update_weights(mod)
    slots = seeded((B * S, D), 71).cuda()
    pu = seeded((B, chunks, S, Din), 72).cuda() * 4096
    pc = (seeded((B, chunks, S), 73).abs() + 0.5).cuda()
    new, qa = ops.slot_update(w, (pu.flatten(), pc.flatten(), chunks, 4096.0), slots, S, Din, D, M, True)
    wd = {k: (v.double() if torch.is_tensor(v) else v) for k, v in w.items()}
    U = (pu.double().sum(1) / (4096.0 * pc.double().sum(1))[..., None]).reshape(B * S, Din)
    gi = U @ wd['w_ivT'] + wd['b_iv']
    gh = slots.double() @ wd['w_hhT'] + wd['b_hh']
    r = torch.sigmoid(gi[:, :D] + gh[:, :D])
    z = torch.sigmoid(gi[:, D:2 * D] + gh[:, D:2 * D])
    n = torch.tanh(gi[:, 2 * D:] + r * gh[:, 2 * D:])
    h = (1 - z) * n + z * slots.double()
    y = torch.relu(sa_ref.layer_norm(h, wd['ln_m_g'], wd['ln_m_b']) @ wd['w1T'] + wd['b1'])
    ref = h + y @ wd['w2T'] + wd['b2']
    ref_qa = sa_ref.layer_norm(ref, wd['ln_q_g'], wd['ln_q_b']) @ wd['w_qaT']
    assert rel_l2(new, ref) < 1e-5 and rel_l2(qa[:, :Din + 1], ref_qa[:, :Din + 1]) < 1e-5
    assert (qa[:, Din + 1:] == 0).all()
    same, qa0 = ops.slot_update(w, None, slots, S, Din, D, M, True)
    assert same is slots
    assert rel_l2(qa0[:, :Din + 1], (sa_ref.layer_norm(slots.double(), wd['ln_q_g'], wd['ln_q_b']) @ wd['w_qaT'])[:, :Din + 1]) < 1e-5


def test_fused_tail_forward_is_graph_capturable(fused_tail):
    """enqueue-only, allocation-free inside the library: the 7-launch forward replays from a CUDA graph"""
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, S, D = 8, 1024, 11, 192
    mod = SlotAttentionWMask(D, 3, S, D, 2 * D).cuda()
    mod.load_state_dict(sa_ref.random_params(D, D, 2 * D, seed=5))
    x, s0 = seeded((B, N, D), 61).cuda(), seeded((B, S, D), 62).cuda()
    with torch.no_grad():
        ref_s, ref_m = mod(x, s0)                              # warm-up: weight fold, cudaFuncSetAttribute
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            mod(x, s0)
        torch.cuda.current_stream().wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            out_s, out_m = mod(x, s0)
        x.copy_(seeded((B, N, D), 63).cuda())
        g.replay()
        torch.cuda.synchronize()
        chk_s, chk_m = mod(x, s0)
    assert torch.equal(out_s, chk_s) and torch.equal(out_m, chk_m)
    assert not torch.equal(out_s, ref_s)
