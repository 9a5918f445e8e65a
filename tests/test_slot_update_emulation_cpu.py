"""The one-launch slot update (csrc/slot_update.cuh) checked on the CPU.

The kernel is written as barrier-separated phases `phase(ph, thread, ...)`; tests/host_emu/slot_update_emu.cpp compiles
the SAME header with g++ and runs every phase for every thread index.  Driven through the product's own sequencing
(autograd.slot_attention_forward_fused_tail, ops.slot_update_args, WeightCache.slot_update_weights) with a torch
stand-in for the tensor-core attend kernel, the result must match the oracle and the reference golden outputs.
What this does not cover: the CUDA launch itself (grid/block/shared-memory size) -- tests/test_slot_update_gpu.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from helpers import SA_CASES, golden, rel_l2, sa_case
from oracle import slot_attention_ref as sa_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ASCALE = 4096.0


def _build_emu(tmp_path_factory, name, extra=()):
    so = str(tmp_path_factory.mktemp(name) / 'slot_update_emu.so')
    src = os.path.join(ROOT, 'tests', 'host_emu', 'slot_update_emu.cpp')
    flags = ['-O2', '-std=c++17', '-shared', '-fPIC'] + list(extra)
    if 'fma' in open('/proc/cpuinfo').read():
        flags.append('-mfma')          # fmaf -> one instruction, same rounding as the device FFMA
    subprocess.check_call(['g++'] + flags + ['-o', so, src])
    lib = ctypes.CDLL(so)
    lib.su_emulate.restype = ctypes.c_int
    lib.su_canary_violations.restype = ctypes.c_long
    return lib


@pytest.fixture(scope='module')
def emu(tmp_path_factory):
    return _build_emu(tmp_path_factory, 'su_emu')


@pytest.fixture(scope='module')
def emu_padded(tmp_path_factory):
    """same phase code, scratch arrays separated by 64 NaN-filled floats each (SU_LAYOUT_PAD)"""
    return _build_emu(tmp_path_factory, 'su_emu_pad', ['-DSU_LAYOUT_PAD=64'])


def attend_cpu(chunks):
    """torch stand-in for sdb_slot_attend_fused_partials: same contract (raw features in, per-chunk partial sums out)."""
    def attend(x, qa, B, N, S, Din, ln_eps, eps, want_mask):
        x = x.double()
        n = (x - x.mean(-1, keepdim=True)) / torch.sqrt(x.var(-1, unbiased=False, keepdim=True) + ln_eps)
        q = qa.double().view(B, S, -1)
        logits = torch.einsum('bnc,bsc->bns', n, q[..., :Din]) + q[..., Din][:, None, :]
        attn = torch.softmax(logits, dim=-1)
        a = attn + eps
        bounds = np.linspace(0, N, chunks + 1).astype(int)
        pu = torch.stack([torch.einsum('bns,bnc->bsc', a[:, lo:hi], n[:, lo:hi]) for lo, hi in zip(bounds[:-1], bounds[1:])], 1)
        pc = torch.stack([a[:, lo:hi].sum(1) for lo, hi in zip(bounds[:-1], bounds[1:])], 1)
        parts = ((pu * ASCALE).float().contiguous().flatten(), pc.float().contiguous().flatten(), chunks, ASCALE)
        return parts, (attn.permute(0, 2, 1).float().contiguous() if want_mask else None)
    return attend


def update_cpu(emu, RT, order=0, log=None):
    from slotdiffusion_b200 import ops

    def update(w, parts, slots_in, S, Din, D, M, want_q):
        rows = slots_in.shape[0]
        slots_out = torch.full_like(slots_in, float('nan')) if parts is not None else None
        qa = torch.full((rows, w['ldq']), float('nan')) if want_q else None
        a = ops.slot_update_args(w, parts, slots_in.contiguous(), slots_out, qa, S, Din, D, M)
        assert emu.su_emulate(ctypes.byref(a), RT, D, order) == 0
        if log is not None:
            log.append((slots_out, qa))
        return (slots_out if parts is not None else slots_in), qa
    return update


def run_module(emu, name, RT, chunks=2, order=0, log=None):
    from slotdiffusion_b200 import autograd
    from slotdiffusion_b200.slot_attention import SlotAttentionWMask
    B, N, Din, S, D, M, I = SA_CASES[name]
    p, x, s0, _, iters = sa_case(name)
    mod = SlotAttentionWMask(Din, I, S, D, M)
    mod.load_state_dict(p)
    with torch.no_grad():
        slots, mask = autograd.slot_attention_forward_fused_tail(
            mod, x, s0.reshape(B * S, D).contiguous(), True, B, N, Din, S, D,
            attend=attend_cpu(chunks), update=update_cpu(emu, RT, order, log))
    return slots, mask, (p, x, s0, iters)


@pytest.mark.parametrize('RT', [4, 8])
@pytest.mark.parametrize('name', list(SA_CASES))
def test_emulated_kernel_matches_oracle_and_reference_golden(emu, name, RT):
    slots, mask, (p, x, s0, iters) = run_module(emu, name, RT)
    assert torch.isfinite(slots).all() and torch.isfinite(mask).all()      # NaN-filled scratch: no read-before-write
    ref_s, ref_m = sa_ref.slot_attention_forward(p, x.double(), s0.double(), iters)
    assert rel_l2(slots, ref_s) < 2e-6
    assert rel_l2(mask, ref_m) < 2e-6
    g = golden(name)                                                        # outputs of the reference module itself
    assert rel_l2(slots, g['slots']) < 5e-6
    assert rel_l2(slots, g['slots64']) < 5e-6


def test_result_does_not_depend_on_thread_order_within_a_phase(emu):
    """forward vs reversed thread order inside every phase: bitwise equal, or a phase has a read/write race"""
    for name in ('sa_img_clevrtex', 'sa_ragged_small', 'sa_coco_vitb16'):
        for RT in (4, 8):
            a, b = [], []
            run_module(emu, name, RT, order=0, log=a)
            run_module(emu, name, RT, order=1, log=b)
            assert len(a) == len(b) > 0
            for (s0, q0), (s1, q1) in zip(a, b):
                for u, v in ((s0, s1), (q0, q1)):
                    if u is not None:
                        assert torch.equal(u, v)


def test_chunk_count_and_row_tile_do_not_change_the_result(emu):
    name = 'sa_ragged_small'            # 15 rows: partial last tile for RT = 4 and RT = 8
    ref = run_module(emu, name, 8, chunks=1)[0]
    for RT, chunks in ((4, 1), (8, 3), (4, 5)):
        assert rel_l2(run_module(emu, name, RT, chunks=chunks)[0], ref) < 1e-6


def test_no_phase_writes_outside_its_scratch_array(emu_padded):
    """every scratch array followed by 64 canary floats: they stay untouched, and the results stay finite and equal"""
    assert emu_padded.su_layout_pad() == 64
    for name in ('sa_img_clevrtex', 'sa_ragged_small', 'sa_coco_vitb16', 'sa_movie_24slots'):
        for RT in (4, 8):
            slots, mask, (p, x, s0, iters) = run_module(emu_padded, name, RT)
            ref_s, _ = sa_ref.slot_attention_forward(p, x.double(), s0.double(), iters)
            assert rel_l2(slots, ref_s) < 2e-6
    assert emu_padded.su_canary_violations() == 0


def test_pad_columns_of_qa_are_written(emu):
    log = []
    run_module(emu, 'sa_ragged_small', 4, log=log)
    qa = log[0][1]
    Din = SA_CASES['sa_ragged_small'][2]
    assert qa.shape[1] == Din + 4 and (qa[:, Din + 1:] == 0).all() and torch.isfinite(qa).all()


def test_struct_layout_matches_header():
    """SdbSlotUpdate field order in include/sdb200.h == ctypes Structure order"""
    import re
    from slotdiffusion_b200._lib import SdbSlotUpdate
    hdr = open(os.path.join(ROOT, 'include', 'sdb200.h')).read()
    body = re.search(r'typedef struct SdbSlotUpdate \{(.*?)\} SdbSlotUpdate;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for stmt in body.split(';'):
        stmt = stmt.strip()
        if not stmt:
            continue
        decl = stmt.split(None, 2)[2] if stmt.startswith('const') else stmt.split(None, 1)[1]
        names += [n.replace('*', '').strip() for n in decl.split(',')]
    assert names == [f[0] for f in SdbSlotUpdate._fields_], names


def test_entry_point_validates_arguments():
    from slotdiffusion_b200 import _lib
    l = _lib.lib()
    assert l.sdb_slot_update_supported(11, 192, 192, 384) == 1 and l.sdb_slot_update_supported(7, 256, 256, 512) == 1
    assert l.sdb_slot_update_supported(11, 192, 200, 384) == 0 and l.sdb_slot_update_supported(11, 190, 192, 384) == 0
    assert l.sdb_slot_update(None, None) == 1
    a = _lib.SdbSlotUpdate()
    a.rows, a.S, a.Din, a.D, a.M, a.ldq = 22, 11, 192, 192, 384, 196
    assert l.sdb_slot_update(ctypes.byref(a), None) == 1 and b'slots_in' in l.sdb_last_error()
    assert l.sdb_slot_attend_fused_partials(None, None, 196, None, None, 1, 16, 4, 192, 1e-5, 1e-6, None) == 1
    assert abs(l.sdb_slot_attend_fused_ascale() - ASCALE) == 0
