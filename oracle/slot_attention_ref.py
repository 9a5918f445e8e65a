"""Oracle: iterative Slot Attention update (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates
  * SlotAttention.forward        /root/reference/slotdiffusion/img_based/models/slot_attention.py:55-104
  * SlotAttentionWMask.forward   /root/reference/slotdiffusion/img_based/models/sa_diffusion.py:15-70
    (video copies: video_based/models/savi.py:57-106, savi_diffusion.py:16-71 -- same math)
from the equations in SURVEY.md Appendix A.1.  Works in the dtype of its inputs
(fp32 for parity, fp64 as the tie-breaker for argmax masks).

`p` is a dict with the reference state_dict keys of the module:
  norm_inputs.{weight,bias}, project_q.0.{weight,bias}, project_q.1.weight,
  project_k.weight, project_v.weight, gru.{weight_ih,weight_hh,bias_ih,bias_hh},
  mlp.0.{weight,bias}, mlp.1.{weight,bias}, mlp.3.{weight,bias}
"""
import torch


def layer_norm(x, w, b, eps=1e-5):
    # nn.LayerNorm over the last dim, biased variance (slot_attention.py:36,40,49)
    mu = x.mean(dim=-1, keepdim=True)
    var = ((x - mu) ** 2).mean(dim=-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gru_cell(x, h, w_ih, w_hh, b_ih, b_hh):
    # nn.GRUCell, gate order (r, z, n) (slot_attention.py:47,97-100)
    D = h.shape[-1]
    gi = x @ w_ih.t() + b_ih
    gh = h @ w_hh.t() + b_hh
    r = torch.sigmoid(gi[..., :D] + gh[..., :D])
    z = torch.sigmoid(gi[..., D:2 * D] + gh[..., D:2 * D])
    n = torch.tanh(gi[..., 2 * D:] + r * gh[..., 2 * D:])
    return (1 - z) * n + z * h


def slot_attention_forward(p, inputs, slots, num_iterations, eps=1e-6, return_trace=False):
    """inputs [B,N,Din], slots [B,S,D] -> (slots [B,S,D], seg_mask [B,S,N]).

    seg_mask is the softmax-over-slots map of the LAST iteration, taken before
    the +eps / spatial renormalisation (sa_diffusion.py:49-51).
    """
    dt = inputs.dtype
    p = {k: v.to(dt) for k, v in p.items()}
    B, N, _ = inputs.shape
    S, D = slots.shape[1], slots.shape[2]
    scale = float(D) ** -0.5                                     # slot_attention.py:34
    x = layer_norm(inputs, p['norm_inputs.weight'], p['norm_inputs.bias'])   # :68
    k = x @ p['project_k.weight'].t()                            # :70
    v = x @ p['project_v.weight'].t()                            # :72
    seg_mask = None
    trace = []
    for it in range(num_iterations):                             # :78
        prev = slots
        q = layer_norm(slots, p['project_q.0.weight'], p['project_q.0.bias']) @ p['project_q.1.weight'].t()  # :82
        logits = scale * torch.einsum('bnd,bsd->bns', k, q)      # :84
        attn = torch.softmax(logits, dim=-1)                     # :85  (over slots)
        if it == num_iterations - 1:
            seg_mask = attn.permute(0, 2, 1).clone()             # sa_diffusion.py:50-51
        a = attn + eps                                           # :89
        a = a / a.sum(dim=1, keepdim=True)                       # :90  (over tokens)
        upd = torch.einsum('bns,bnd->bsd', a, v)                 # :91
        s = gru_cell(upd.reshape(B * S, D), prev.reshape(B * S, D),
                     p['gru.weight_ih'], p['gru.weight_hh'], p['gru.bias_ih'], p['gru.bias_hh'])  # :97-100
        s = s.reshape(B, S, D)
        h = layer_norm(s, p['mlp.0.weight'], p['mlp.0.bias'])
        h = torch.relu(h @ p['mlp.1.weight'].t() + p['mlp.1.bias'])
        slots = s + h @ p['mlp.3.weight'].t() + p['mlp.3.bias']  # :102
        if return_trace:
            trace.append(dict(q=q, logits=logits, attn=attn, updates=upd, gru=s, slots=slots))
    if return_trace:
        return slots, seg_mask, trace
    return slots, seg_mask


def slot_attention_video(p, frames, init_slots, num_iterations, predictor=None, eps=1e-6):
    """Per-frame driver (savi_diffusion.py:183-196): frames [B,T,N,Din].

    slots of frame t start from init_slots (t=0) or predictor(slots of t-1).
    """
    B, T = frames.shape[:2]
    out_s, out_m = [], []
    prev = None
    for t in range(T):
        lat = init_slots if prev is None else (predictor(prev) if predictor is not None else prev)
        s, m = slot_attention_forward(p, frames[:, t], lat, num_iterations, eps)
        out_s.append(s)
        out_m.append(m)
        prev = s
    return torch.stack(out_s, 1), torch.stack(out_m, 1)


def random_params(in_features, slot_size, mlp_hidden, seed=0, dtype=torch.float32):
    """Parameter set with nn-default-like init (for property tests at sizes without a fixture)."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan_in):
        bound = fan_in ** -0.5
        return ((torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * bound).to(dtype)

    def ln(n):
        return ((1 + 0.1 * torch.randn(n, generator=g, dtype=torch.float64)).to(dtype),
                (0.1 * torch.randn(n, generator=g, dtype=torch.float64)).to(dtype))
    D, Din, M = slot_size, in_features, mlp_hidden
    p = {}
    p['norm_inputs.weight'], p['norm_inputs.bias'] = ln(Din)
    p['project_q.0.weight'], p['project_q.0.bias'] = ln(D)
    p['project_q.1.weight'] = u((D, D), D)
    p['project_k.weight'] = u((D, Din), Din)
    p['project_v.weight'] = u((D, Din), Din)
    p['gru.weight_ih'] = u((3 * D, D), D)
    p['gru.weight_hh'] = u((3 * D, D), D)
    p['gru.bias_ih'] = u((3 * D,), D)
    p['gru.bias_hh'] = u((3 * D,), D)
    p['mlp.0.weight'], p['mlp.0.bias'] = ln(D)
    p['mlp.1.weight'] = u((M, D), D)
    p['mlp.1.bias'] = u((M,), D)
    p['mlp.3.weight'] = u((D, M), M)
    p['mlp.3.bias'] = u((D,), M)
    return p
