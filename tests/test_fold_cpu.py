"""CPU check of the Slot-Attention weight fold used by the tensor-core inference path: the folded formulation
(no k / v tensors, LayerNorm affine folded into the slot-side projection and the GRU input weight) must reproduce
the oracle (restatement of slot_attention.py:67-104) in fp64 to round-off."""
import pytest
import torch

from helpers import SA_CASES, rel_l2, sa_case
from oracle import slot_attention_ref as sa_ref


def folded_forward(p, x, slots, iters, scale, eps=1e-6, ln_eps=1e-5):
    from slotdiffusion_b200.ops import slot_attention_fold_math
    P = {k: v.double() for k, v in p.items()}
    wqa, wiv, biv = slot_attention_fold_math(P['project_q.1.weight'], P['project_k.weight'], P['project_v.weight'],
                                             P['norm_inputs.weight'], P['norm_inputs.bias'], P['gru.weight_ih'],
                                             P['gru.bias_ih'], scale)
    x = x.double()
    Din = x.shape[-1]
    n = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + ln_eps)
    s = slots.double()
    D = s.shape[-1]
    ln = torch.nn.functional.layer_norm
    mask = None
    for _ in range(iters):
        sn = ln(s, (D,), P['project_q.0.weight'], P['project_q.0.bias'], 1e-5)
        qa = sn @ wqa.t()                                             # [B, S, Din + 4]
        logits = n @ qa[..., :Din].transpose(1, 2) + qa[..., Din][:, None, :]
        attn = torch.softmax(logits, -1)
        mask = attn.transpose(1, 2)
        a = attn + eps
        U = torch.einsum('bns,bnd->bsd', a / a.sum(1, keepdim=True), n)
        gi = U @ wiv.t() + biv
        gh = s @ P['gru.weight_hh'].t() + P['gru.bias_hh']
        r = torch.sigmoid(gi[..., :D] + gh[..., :D])
        z = torch.sigmoid(gi[..., D:2 * D] + gh[..., D:2 * D])
        nn_ = torch.tanh(gi[..., 2 * D:] + r * gh[..., 2 * D:])
        h = (1 - z) * nn_ + z * s
        hn = ln(h, (D,), P['mlp.0.weight'], P['mlp.0.bias'], 1e-5)
        s = h + torch.relu(hn @ P['mlp.1.weight'].t() + P['mlp.1.bias']) @ P['mlp.3.weight'].t() + P['mlp.3.bias']
    return s, mask


@pytest.mark.parametrize('name', list(SA_CASES))
def test_fold_matches_oracle(name):
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, _, iters = sa_case(name)
    ref_s, ref_m = sa_ref.slot_attention_forward(p, x.double(), s0.double(), iters)
    got_s, got_m = folded_forward(p, x, s0, iters, D ** -0.5)
    assert rel_l2(got_s, ref_s) < 1e-10
    assert rel_l2(got_m, ref_m) < 1e-10
    assert torch.equal(got_m.argmax(1), ref_m.argmax(1))
