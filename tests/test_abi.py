"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads without a GPU/driver and exports
every symbol include/sdb200.h declares; the ctypes signature table covers exactly those symbols."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    hdr = open(os.path.join(ROOT, 'include', 'sdb200.h')).read()
    return set(re.findall(r'\b(sdb_[a-z0-9_]+)\s*\(', hdr))


def test_library_builds_and_exports_header_symbols():
    from slotdiffusion_b200 import build
    path = build.build(verbose=False)
    l = ctypes.CDLL(path)
    for name in _declared():
        assert hasattr(l, name), name


def test_ctypes_table_matches_header():
    from slotdiffusion_b200 import _lib
    assert set(_lib.SIGNATURES) == _declared()
    assert _lib.lib().sdb_version() >= 1


def test_gemm_struct_layout_matches_header():
    """SdbGemm field order in the header == ctypes Structure order."""
    from slotdiffusion_b200._lib import SdbGemm
    hdr = open(os.path.join(ROOT, 'include', 'sdb200.h')).read()
    body = re.search(r'typedef struct SdbGemm \{(.*?)\} SdbGemm;', hdr, re.S).group(1)
    body = re.sub(r'/\*.*?\*/', '', body, flags=re.S)
    names = []
    for stmt in body.split(';'):
        stmt = stmt.strip()
        if not stmt:
            continue
        decl = stmt.split(None, 1)[1] if not stmt.startswith('const') else stmt.split(None, 2)[2]
        for n in decl.split(','):
            names.append(n.replace('*', '').strip())
    assert names == [f[0] for f in SdbGemm._fields_], names


def test_invalid_arguments_are_reported_not_ub():
    from slotdiffusion_b200 import _lib
    l = _lib.lib()
    g = _lib.SdbGemm()
    assert l.sdb_gemm(ctypes.byref(g), None) == 1          # SDB_ERR_INVALID
    assert b'null' in l.sdb_last_error()
    assert l.sdb_slot_attend(None, None, None, None, None, None, 1, 1, 1, 192, 1.0, 1e-6, None) == 1


def test_new_entry_points_validate_arguments():
    """argument validation of the tensor-core entry points (no CUDA call is reached)"""
    from slotdiffusion_b200 import _lib
    l = _lib.lib()
    assert l.sdb_slot_attend_fused_supported(11, 192) == 1 and l.sdb_slot_attend_fused_supported(24, 256) == 1
    assert l.sdb_slot_attend_fused_supported(33, 192) == 0 and l.sdb_slot_attend_fused_supported(11, 64) == 0
    assert l.sdb_slot_attend_fused(None, None, 196, None, None, None, None, 1, 16, 4, 192, 1e-5, 1e-6, None) == 1
    assert b'null' in l.sdb_last_error()
    assert l.sdb_slot_attend_fused_workspace(64, 1024, 11, 192) == (64 * 2 * 11 * 192 + 64 * 2 * 11) * 4 or \
        l.sdb_slot_attend_fused_workspace(64, 1024, 11, 192) > 0     # chunks depend on the SM count of the device
    assert l.sdb_attention_tc_supported(8, 32, 768, 768, 768) == 1 and l.sdb_attention_tc_supported(8, 64, 768, 768, 768) == 0
    assert l.sdb_attention_tc(None, 4, None, 4, None, 4, None, 1, 16, 16, 2, 32, 1.0, None) == 1
    assert l.sdb_channel_block_sums(None, 128, None, 1, 16, None) == 1


def test_every_entry_point_rejects_null_and_zero_arguments():
    """SURVEY 8b 'shape/alignment violations are reported, never UB': each compute entry point called with NULL
    pointers and zero sizes returns SDB_ERR_INVALID with a message (no CUDA call is reached, runs without a GPU)."""
    from slotdiffusion_b200 import _lib
    l = _lib.lib()
    queries = {'sdb_version', 'sdb_last_error', 'sdb_launch_count', 'sdb_attention_tc_supported', 'sdb_attention_fewkeys_supported',
               'sdb_slot_attend_workspace', 'sdb_slot_attend_fused_supported', 'sdb_slot_attend_fused_debug',
               'sdb_slot_attend_fused_workspace', 'sdb_slot_attend_fused_chunks', 'sdb_slot_attend_fused_ascale',
               'sdb_slot_update_supported', 'sdb_token_attention_supported', 'sdb_slot_attention_resident_supported', 'sdb_slot_attention_resident_debug', 'sdb_slot_attention_resident_wave',
               'sdb_set_pack_mode'}      # (fmt = 0, default stream) is a VALID call: a mode switch, no operands
    for name, (res, args) in _lib.SIGNATURES.items():
        if name in queries:
            continue
        vals = [1.0 if a is ctypes.c_float else (0 if a in (ctypes.c_int, ctypes.c_int32, ctypes.c_int64, ctypes.c_uint64) else None)
                for a in args]
        rc = getattr(l, name)(*vals)
        assert rc == 1, (name, rc)
        assert len(l.sdb_last_error()) > 0, name
