"""Oracle: slot transition function (TEST INFRASTRUCTURE, see oracle/__init__.py).

Restates TransformerPredictor.forward, /root/reference/slotdiffusion/video_based/models/predictor.py:20-44 -- an
nn.TransformerEncoder (batch_first, ReLU feed-forward, no final norm) over the slots [B, S, D].  The layer math is that of
torch.nn.TransformerEncoderLayer (pinned torch 2.x: `_sa_block` = MultiheadAttention(x, x, x) + dropout1, `_ff_block` =
linear2(dropout(relu(linear1(x)))) + dropout2; pre-LN `x + sa(norm1(x))`, `x + ff(norm2(x))`; post-LN
`norm1(x + sa(x))`, `norm2(x + ff(x))`) and torch.nn.MultiheadAttention (fused in_proj rows q | k | v, head h = columns
h*dh..(h+1)*dh of each third, scores scaled by dh^-1/2, softmax over keys, out_proj).  Dropout is the identity here (eval
mode); the training-mode masks of the product are checked statistically and through mask-consistent gradients in
tests/test_predictor_gpu.py.  Pinned by tests/golden/predictor.npz (outputs + gradients of the unmodified reference
module, tools/make_golden.py gen_predictor).

`p` is a dict with the reference state_dict keys:
  transformer_encoder.layers.{i}.self_attn.{in_proj_weight,in_proj_bias,out_proj.weight,out_proj.bias}
  transformer_encoder.layers.{i}.{linear1,linear2,norm1,norm2}.{weight,bias}
Works in the dtype of its inputs (fp32 for parity, fp64 as the reference point of the GPU tests).
"""
import torch

from .slot_attention_ref import layer_norm


def random_state_dict(d_model, num_layers, ffn_dim, seed, dtype=torch.float32):
    g = torch.Generator().manual_seed(seed)

    def rn(*shape, scale=1.0):
        return (torch.randn(*shape, generator=g, dtype=torch.float64) * scale).to(dtype)
    sd = {}
    for i in range(num_layers):
        k = f'transformer_encoder.layers.{i}.'
        sd[k + 'self_attn.in_proj_weight'] = rn(3 * d_model, d_model, scale=d_model ** -0.5)
        sd[k + 'self_attn.in_proj_bias'] = rn(3 * d_model, scale=0.1)
        sd[k + 'self_attn.out_proj.weight'] = rn(d_model, d_model, scale=d_model ** -0.5)
        sd[k + 'self_attn.out_proj.bias'] = rn(d_model, scale=0.1)
        sd[k + 'linear1.weight'] = rn(ffn_dim, d_model, scale=d_model ** -0.5)
        sd[k + 'linear1.bias'] = rn(ffn_dim, scale=0.1)
        sd[k + 'linear2.weight'] = rn(d_model, ffn_dim, scale=ffn_dim ** -0.5)
        sd[k + 'linear2.bias'] = rn(d_model, scale=0.1)
        for n in ('norm1', 'norm2'):
            sd[k + n + '.weight'] = 1 + rn(d_model, scale=0.1)
            sd[k + n + '.bias'] = rn(d_model, scale=0.1)
    return sd


def self_attention(p, k, x, num_heads):
    """nn.MultiheadAttention(x, x, x, need_weights=False), batch_first; x [B, S, D]."""
    B, S, D = x.shape
    dh = D // num_heads
    qkv = x @ p[k + 'self_attn.in_proj_weight'].t() + p[k + 'self_attn.in_proj_bias']
    q, kk, v = (t.reshape(B, S, num_heads, dh).transpose(1, 2) for t in qkv.split(D, dim=-1))
    a = torch.softmax((q @ kk.transpose(-1, -2)) * dh ** -0.5, dim=-1)
    o = (a @ v).transpose(1, 2).reshape(B, S, D)
    return o @ p[k + 'self_attn.out_proj.weight'].t() + p[k + 'self_attn.out_proj.bias']


def feed_forward(p, k, x):
    h = torch.relu(x @ p[k + 'linear1.weight'].t() + p[k + 'linear1.bias'])
    return h @ p[k + 'linear2.weight'].t() + p[k + 'linear2.bias']


def predictor_forward(p, x, num_layers, num_heads, norm_first=True, eps=1e-5):
    """x [B, S, D] -> [B, S, D] (predictor.py:42-44)."""
    for i in range(num_layers):
        k = f'transformer_encoder.layers.{i}.'
        n1 = lambda t: layer_norm(t, p[k + 'norm1.weight'], p[k + 'norm1.bias'], eps)
        n2 = lambda t: layer_norm(t, p[k + 'norm2.weight'], p[k + 'norm2.bias'], eps)
        if norm_first:
            x = x + self_attention(p, k, n1(x), num_heads)
            x = x + feed_forward(p, k, n2(x))
        else:
            x = n1(x + self_attention(p, k, x, num_heads))
            x = n2(x + feed_forward(p, k, x))
    return x
