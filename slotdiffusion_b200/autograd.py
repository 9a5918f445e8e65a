"""Forward/backward drivers that sequence the C-ABI kernels for the hot modules.

`slot_attention_apply` runs the whole iterative Slot Attention update
(reference: img_based/models/slot_attention.py:67-104, sa_diffusion.py:28-70) as a chain of
hand-written kernels:

  LayerNorm+pack -> [Wk|Wv] GEMM (tcgen05)                               once
  per iteration: LayerNorm+pack -> Wq GEMM -> fused attend (TMA-staged, one pass over k|v)
                 -> W_ih / W_hh GEMMs -> GRU gates -> LayerNorm+pack -> MLP GEMMs (+ReLU, +residual)
"""
import os

import torch

from . import ops


FUSED_ATTEND = True   # inference: tensor-core attend over the raw features (no k / v tensors)
# One-launch slot update (csrc/slot_update.cuh) instead of the ten-launch GEMM tail, for B*S <= FUSED_TAIL_MAX_ROWS.
# Opt-in (SDB_SA_FUSED_TAIL=1) until it has been timed on the B200: written when the round's GPU budget was spent;
# its arithmetic is checked on the CPU through the host emulation (tests/test_slot_update_emulation_cpu.py).
FUSED_TAIL = os.environ.get('SDB_SA_FUSED_TAIL', '0') == '1'
FUSED_TAIL_MAX_ROWS = 1024
# Whole forward as ONE persistent cluster kernel (csrc/slot_attention_resident.cu): features read from HBM once and kept
# in shared memory over all iterations.  SDB_SA_RESIDENT=0 falls back to the per-iteration kernels above.
RESIDENT = os.environ.get('SDB_SA_RESIDENT', '1') == '1'
# Measured (profiles/README.md, round 2): one wave (= one sample per resident cluster, 33 on a B200 for N = 1024, D = 192)
# takes ~180 us whatever its fill, the per-iteration path 240 us at B = 33 and 270 us at B = 64 -- so the persistent kernel
# is used up to RESIDENT_WAVES waves; SDB_SA_RESIDENT_WAVES=1000000 forces it for every batch size.
RESIDENT_WAVES = int(os.environ.get('SDB_SA_RESIDENT_WAVES', '1'))


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def slot_attention_forward(mod, inputs, slots, want_mask, save=None):
    """Returns (slots [B,S,D], seg_mask [B,S,N] or None).  `save`: optional list that receives the
    per-iteration tensors needed by the backward pass."""
    B, N, Din = inputs.shape
    S, D = slots.shape[1], slots.shape[2]
    wc = mod._wcache
    inputs = inputs.contiguous().float()
    slots = slots.contiguous().float().reshape(B * S, D)
    if (save is None and RESIDENT and ops.slot_attention_resident_supported(N, S, Din, D, mod.mlp_hidden_size)
            and B <= RESIDENT_WAVES * ops.slot_attention_resident_wave(N, S, Din, D, mod.mlp_hidden_size)):
        w = wc.slot_resident_weights(mod)                                     # :67-104 in one launch
        out, mask = ops.slot_attention_resident(w, inputs, slots.view(B, S, D), mod.num_iterations, mod.norm_inputs.eps,
                                                mod.eps, mod.mlp_hidden_size, want_mask)
        return out, mask
    if save is None and FUSED_ATTEND and ops.slot_attend_fused_supported(S, Din):
        if FUSED_TAIL and B * S <= FUSED_TAIL_MAX_ROWS and ops.slot_update_supported(S, Din, D, mod.mlp_hidden_size):
            return slot_attention_forward_fused_tail(mod, inputs, slots, want_mask, B, N, Din, S, D)
        return _slot_attention_forward_fused(mod, inputs, slots, want_mask, B, N, Din, S, D)

    # k | v projection of LayerNorm(inputs)                                  (slot_attention.py:68-72)
    xn = ops.layernorm_pack(inputs.reshape(B * N, Din), mod.norm_inputs.weight, mod.norm_inputs.bias,
                            mod.norm_inputs.eps)
    w_kv = wc.linear('kv', mod.project_k.weight, mod.project_v.weight)
    kv = ops.gemm(xn, w_kv)                                                   # [B*N, 2D]
    w_q = wc.linear('q', mod.project_q[1].weight)
    w_ih = wc.linear('ih', mod.gru.weight_ih)
    w_hh = wc.linear('hh', mod.gru.weight_hh)
    w_1 = wc.linear('m1', mod.mlp[1].weight)
    w_2 = wc.linear('m2', mod.mlp[3].weight)
    if save is not None:
        save.append(dict(xn=xn, kv=kv))
    mask = None
    hp = None
    for it in range(mod.num_iterations):
        last = it == mod.num_iterations - 1
        prev = slots
        sn = ops.layernorm_pack(prev, mod.project_q[0].weight, mod.project_q[0].bias, mod.project_q[0].eps)
        q = ops.gemm(sn, w_q)                                                 # :82
        upd, m, upd32 = ops.slot_attend(kv, q, B, N, S, D, mod.attn_scale, mod.eps, want_mask and last,
                                        want_fp32=save is not None)           # :84-91
        if want_mask and last:
            mask = m
        gi = ops.gemm(upd, w_ih, bias=mod.gru.bias_ih)                        # :97-100
        if hp is None:
            hp = ops.pack_rows(prev)
        gh = ops.gemm(hp, w_hh, bias=mod.gru.bias_hh)
        h = ops.gru_gates(gi, gh, prev)
        hn = ops.layernorm_pack(h, mod.mlp[0].weight, mod.mlp[0].bias, mod.mlp[0].eps)
        y1, y1p = ops.gemm(hn, w_1, bias=mod.mlp[1].bias, relu=True, pack_out='none',
                           keep_c=save is not None)                           # :102 (ReLU + operand packing fused)
        slots, sp = ops.gemm(y1p, w_2, bias=mod.mlp[3].bias, residual=h, pack_out='none')
        if save is not None:
            save.append(dict(prev=prev, sn=sn, q=q, upd=upd, upd32=upd32, gi=gi, gh=gh, hp=hp, h=h, hn=hn, y1=y1,
                             y1p=y1p))
        hp = sp                                                              # next iteration's W_hh operand
    return slots.view(B, S, D), mask


def _slot_attention_forward_fused(mod, inputs, slots, want_mask, B, N, Din, S, D):
    """Inference path: no k / v tensors.  Per iteration: LayerNorm+pack -> folded slot-side projection (GEMM) ->
    tensor-core attend over the RAW features (csrc/slot_attention_fused.cu) -> GRU / MLP tail."""
    wc = mod._wcache
    fold = wc.slot_attention_fold(mod)
    w_hh = wc.linear('hh', mod.gru.weight_hh)
    w_1 = wc.linear('m1', mod.mlp[1].weight)
    w_2 = wc.linear('m2', mod.mlp[3].weight)
    mask = None
    hp = None
    for it in range(mod.num_iterations):
        last = it == mod.num_iterations - 1
        prev = slots
        sn = ops.layernorm_pack(prev, mod.project_q[0].weight, mod.project_q[0].bias, mod.project_q[0].eps)
        qa = ops.gemm(sn, fold['w_qa'])                                       # [B*S, Din+4]  (:82 and k-side of :84)
        upd, m, _ = ops.slot_attend_fused(inputs, qa, B, N, S, Din, mod.norm_inputs.eps, mod.eps,
                                          want_mask and last)                 # :68-72, :84-91
        if want_mask and last:
            mask = m
        gi = ops.gemm(upd, fold['w_iv'], bias=fold['b_iv'])                   # v-side of :91 folded into :97-100
        if hp is None:
            hp = ops.pack_rows(prev)
        gh = ops.gemm(hp, w_hh, bias=mod.gru.bias_hh)
        h = ops.gru_gates(gi, gh, prev)
        hn = ops.layernorm_pack(h, mod.mlp[0].weight, mod.mlp[0].bias, mod.mlp[0].eps)
        _, y1p = ops.gemm(hn, w_1, bias=mod.mlp[1].bias, relu=True, pack_out='none', keep_c=False)
        slots, hp = ops.gemm(y1p, w_2, bias=mod.mlp[3].bias, residual=h, pack_out='none')
    return slots.view(B, S, D), mask


def slot_attention_forward_fused_tail(mod, inputs, slots, want_mask, B, N, Din, S, D, attend=None, update=None):
    """Inference path with the one-launch slot update: 1 + 2 * iterations launches per forward
    (projection of the initial slots; then attend, update+projection per iteration).
    attend / update default to the CUDA entry points; the CPU emulation test passes host stand-ins with the same
    signatures, so this sequencing is what both run."""
    attend = attend or ops.slot_attend_fused_partials
    update = update or ops.slot_update
    w = mod._wcache.slot_update_weights(mod)
    M = mod.mlp_hidden_size
    mask = None
    _, qa = update(w, None, slots, S, Din, D, M, True)                       # :82 of iteration 0 (+ k-side fold of :84)
    for it in range(mod.num_iterations):
        last = it == mod.num_iterations - 1
        parts, m = attend(inputs, qa, B, N, S, Din, mod.norm_inputs.eps, mod.eps, want_mask and last)   # :68-72, :84-91
        if want_mask and last:
            mask = m
        slots, qa = update(w, parts, slots, S, Din, D, M, not last)          # :91 (v side), :97-102, next :82
    return slots.view(B, S, D), mask


def slot_attention_apply(mod, inputs, slots, want_mask):
    params = [p for p in mod.parameters()]
    if _needs_grad(inputs, slots, *params):
        from .backward import SlotAttentionFn
        out = SlotAttentionFn.apply(mod, want_mask, inputs, slots, *params)
        return out[0], (out[1] if want_mask else None)
    # always the fp32-faithful three-pass operand format (bit-exact argmax masks), whatever scope the caller is in
    with torch.no_grad(), ops.pack_format(ops.SDB_FMT_F16X2):
        return slot_attention_forward(mod, inputs, slots, want_mask)
