"""ctypes binding of libsdb200.so (the C ABI declared in include/sdb200.h).

There is NO fallback: if the shared library is missing or a call fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p

HERE = os.path.dirname(os.path.abspath(__file__))
# SDB_LIB selects a build variant of the same sources (slotdiffusion_b200/build.py: experimental / diagnostic macros);
# unset = the product build
LIB_PATH = os.environ.get('SDB_LIB') or os.path.join(HERE, 'libsdb200.so')

SDB_A_PLAIN, SDB_A_CONV3, SDB_A_CONV3S2, SDB_A_WGRAD, SDB_A_WGRAD_S2, SDB_A_CONV3S2A = 0, 1, 2, 3, 4, 5
SDB_PACK_PLAIN, SDB_PACK_UP2, SDB_PACK_PHASE2 = 0, 1, 2


class SdbGemm(Structure):
    _fields_ = [
        ('a', c_void_p), ('w', c_void_p), ('c', c_void_p), ('bias', c_void_p), ('rowvec', c_void_p),
        ('residual', c_void_p), ('a_plane_stride', c_int64), ('ldc', c_int64), ('ldv', c_int64), ('ldr', c_int64),
        ('M', c_int32), ('N', c_int32), ('K', c_int32), ('mode', c_int32), ('B', c_int32), ('H', c_int32),
        ('W', c_int32), ('C', c_int32), ('rows_per_group', c_int32), ('passes', c_int32), ('relu', c_int32),
        ('out_packed', c_void_p), ('gsum', c_void_p), ('out_plane_stride', c_int64), ('out_act', c_int32),
        ('geglu', c_int32), ('a_bf16', c_int32), ('w_bf16', c_int32), ('corr_scale', c_float), ('gsum_cb', c_int32), ('w_plane_stride', c_int64),
        ('batch_rows', c_int32), ('w_row_step', c_int32), ('w_k_step', c_int32), ('alpha', c_float),
        ('w_rows', c_int64), ('w_cols', c_int64),
    ]


class SdbSlotUpdate(Structure):
    _fields_ = [(n, c_void_p) for n in (
        'part_upd', 'part_cs', 'slots_in', 'w_ivT', 'b_iv', 'w_hhT', 'b_hh', 'ln_m_g', 'ln_m_b', 'w1T', 'b1', 'w2T', 'b2',
        'ln_q_g', 'ln_q_b', 'w_qaT', 'slots_out', 'qa_out')] + [
        ('rows', c_int64), ('S', c_int32), ('Din', c_int32), ('D', c_int32), ('M', c_int32), ('ldq', c_int32),
        ('chunks', c_int32), ('ascale', c_float), ('ln_m_eps', c_float), ('ln_q_eps', c_float), ('do_update', c_int32)]


class SdbSlotAttentionResident(Structure):
    _fields_ = [(n, c_void_p) for n in (
        'x', 'slots_in', 'slots_out', 'seg_mask', 'w_iv4', 'b_iv', 'w_hh4', 'b_hh', 'ln_m_g', 'ln_m_b', 'w1_4', 'b1', 'w2_4',
        'b2', 'ln_q_g', 'ln_q_b', 'w_qa4')] + [
        ('B', c_int64), ('N', c_int32), ('S', c_int32), ('Din', c_int32), ('D', c_int32), ('M', c_int32), ('ldq', c_int32),
        ('iterations', c_int32), ('ln_in_eps', c_float), ('attn_eps', c_float), ('ln_m_eps', c_float),
        ('ln_q_eps', c_float)]


# name -> (restype, argtypes); must list every symbol declared in include/sdb200.h
SIGNATURES = {
    'sdb_version': (c_int, []),
    'sdb_last_error': (c_char_p, []),
    'sdb_launch_count': (c_int64, []),
    'sdb_gemm': (c_int, [POINTER(SdbGemm), c_void_p]),
    'sdb_gemm_timing': (c_int, [c_void_p, c_int]),
    'sdb_pack_weight': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_pack_weight_fmt': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    'sdb_pack_weight_conv3_fmt': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    'sdb_set_pack_mode': (c_int, [c_int, c_void_p]),
    'sdb_groupnorm_add_relu': (c_int, [c_void_p] * 10 + [c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_softmax_pack': (c_int, [c_void_p, c_int64, c_float, c_float, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_q_sample': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_mse_loss_fwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'sdb_mse_loss_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'sdb_mask_upsample_argmax': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int, c_int,
                                         c_void_p]),
    'sdb_pack_weight_conv3': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_pack_rows': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    'sdb_layernorm_pack': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int64, c_int64,
                                   c_void_p]),
    'sdb_groupnorm_stats': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int,
                                    c_float, c_void_p]),
    'sdb_groupnorm_apply_pack': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                         c_void_p, c_int64, c_int64, c_int, c_int, c_void_p]),
    'sdb_groupnorm_apply_pack_fused': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p,
                                               c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_float, c_int,
                                               c_void_p]),
    'sdb_channel_block_sums': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_groupnorm_finalize': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int,
                                       c_float, c_void_p]),
    'sdb_groupnorm_finalize_cb': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_float, c_int, c_void_p]),
    'sdb_pack_weight_geglu': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_pack_nhwc': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int64,
                              c_int, c_void_p]),
    'sdb_geglu_pack': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_timestep_embedding_pack': (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p]),
    'sdb_attention_pack': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                   c_int64, c_int64, c_int, c_int, c_float, c_void_p]),
    'sdb_attention_tc_supported': (c_int, [c_int64, c_int64, c_int64, c_int64, c_int64]),
    'sdb_attention_tc': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                 c_int64, c_int64, c_int, c_int, c_float, c_void_p]),
    'sdb_attention_fewkeys_supported': (c_int, [c_int64, c_int64, c_int64, c_int64, c_int64, c_int64]),
    'sdb_attention_fewkeys': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64,
                                 c_int64, c_int64, c_int, c_int, c_float, c_void_p]),
    'sdb_conv3_in': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_int64,
                             c_void_p]),
    'sdb_conv3_out': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                              c_int64, c_int64, c_int64, c_int, c_int64, c_void_p]),
    'sdb_slot_attend_workspace': (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    'sdb_slot_attend': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64,
                                c_int64, c_int64, c_float, c_float, c_void_p]),
    'sdb_slot_attend_train': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int64, c_int64, c_int64, c_float, c_float, c_void_p]),
    'sdb_slot_attend_fused_supported': (c_int, [c_int64, c_int64]),
    'sdb_slot_attend_fused_debug': (c_int, [c_void_p]),
    'sdb_slot_attend_fused_workspace': (c_int64, [c_int64, c_int64, c_int64, c_int64]),
    'sdb_slot_attend_fused': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                      c_int64, c_int64, c_int64, c_float, c_float, c_void_p]),
    'sdb_slot_attend_fused_chunks': (c_int64, [c_int64, c_int64]),
    'sdb_slot_attend_fused_ascale': (c_float, []),
    'sdb_slot_attend_fused_partials': (c_int, [c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                                               c_int64, c_int64, c_float, c_float, c_void_p]),
    'sdb_slot_update_supported': (c_int, [c_int64, c_int64, c_int64, c_int64]),
    'sdb_slot_update': (c_int, [POINTER(SdbSlotUpdate), c_void_p]),
    'sdb_token_attention_supported': (c_int, [c_int64, c_int64]),
    'sdb_token_attention': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_float, c_uint64,
                                    c_void_p, c_void_p]),
    'sdb_token_attention_bwd': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_float,
                                        c_float, c_uint64, c_void_p, c_void_p]),
    'sdb_dropout_add': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_float, c_uint64, c_void_p, c_void_p]),
    'sdb_slot_attention_resident_supported': (c_int, [c_int64, c_int64, c_int64, c_int64, c_int64]),
    'sdb_slot_attention_resident': (c_int, [POINTER(SdbSlotAttentionResident), c_void_p]),
    'sdb_slot_attention_resident_debug': (c_int, [c_void_p]),
    'sdb_slot_attention_resident_wave': (c_int64, [c_int64, c_int64, c_int64, c_int64, c_int64]),
    'sdb_groupnorm_apply_pack_dropout': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p,
                                                 c_void_p, c_int64, c_int64, c_int, c_int, c_float, c_uint64, c_void_p, c_void_p]),
    'sdb_grad_pack': (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int64,
                              c_int64, c_int, c_void_p]),
    'sdb_transpose_packed': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_repack_bf16': (c_int, [c_void_p, c_void_p, c_int64, c_void_p]),
    'sdb_pack_weight_conv3_dgrad': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int, c_void_p]),
    'sdb_wgrad_conv3_scatter': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p]),
    'sdb_add3': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
    'sdb_act_bwd': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_int,
                            c_void_p]),
    'sdb_groupnorm_bwd': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int,
                                  c_int, c_float, c_uint64, c_void_p, c_void_p]),
    'sdb_layernorm_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int64, c_int64, c_void_p]),
    'sdb_attention_bwd': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p,
                                  c_int64, c_void_p, c_int64, c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64,
                                  c_int, c_int, c_float, c_void_p]),
    'sdb_geglu_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_up2_adjoint': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_pack_zero_up2': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_pack_nchw_pad': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_nhwc_to_nchw': (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_nchw_to_nhwc_pad': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_im2col_t': (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int64, c_void_p]),
    'sdb_gru_gates_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                  c_int64, c_void_p]),
    'sdb_slot_attend_bwd': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64,
                                    c_int64, c_int64, c_int64, c_float, c_float, c_int, c_void_p]),
    'sdb_gru_gates': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p]),
    'sdb_dpm_x0': (c_int, [c_void_p, c_void_p, c_float, c_float, c_void_p, c_int64, c_void_p, c_void_p, c_int64,
                           c_int64, c_int64, c_void_p]),
    'sdb_lincomb': (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_float, c_float, c_float, c_int64, c_void_p]),
}

_lib = None


def lib():
    """Load libsdb200.so (once).  Raises RuntimeError if it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f'{LIB_PATH} not found: build it with `python -m slotdiffusion_b200.build` '
                '(there is no CPU / PyTorch fallback for the hot path)')
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(l, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = l
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().sdb_last_error().decode(errors='replace')
        raise RuntimeError(f'{what} failed (code {rc}): {msg}')


def launch_count():
    return int(lib().sdb_launch_count())
