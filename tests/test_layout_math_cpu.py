"""Host-side model of the shared-memory operand layouts the hand-written kernels produce (csrc/slot_attention_fused.cu,
csrc/attention_tc.cu): the thread-level store formulas are restated in Python and checked, for every element, against
the canonical 128B-swizzled UMMA layouts the tcgen05 descriptors declare (K-major: 8-row groups SBO apart, 16-byte chunk
index XOR row%8; MN-major: the same bytes read with rows = K).  Guards the address arithmetic against regressions."""
import pytest


def canon_kmajor(row, k, sbo):
    """byte offset of element (row, k) of one 64-wide fp16 K block, SWIZZLE_128B, 8-row groups `sbo` bytes apart"""
    b = k * 2
    return (row // 8) * sbo + (row % 8) * 128 + (((b // 16) ^ (row % 8)) * 16) + b % 16


def canon_mnmajor(mn, krow, lbo, sbo):
    """byte offset of element (mn, krow): 64-element MN blocks `lbo` apart, groups of 8 K rows `sbo` apart"""
    b = (mn % 64) * 2
    return (mn // 64) * lbo + (krow // 8) * sbo + (krow % 8) * 128 + (((b // 16) ^ (krow % 8)) * 16) + b % 16


@pytest.mark.parametrize('din', [128, 192, 256])
def test_slot_attention_feature_tile(din):
    """converter stores (slot_attention_fused.cu): lane (sub, j), round q, float4 index k -> 4 channels of one token"""
    nkb = din // 64
    gb = 2 * nkb * 1024
    seen = {}
    for g in range(16):                       # 8-token groups of a 128-token tile
        for sub in range(4):
            for j in range(8):
                for q in range(2):
                    tg = 4 * (sub & 1) + (sub >> 1) + 2 * q
                    o0 = tg * 128 + ((((j >> 1) ^ (tg & 3)) | ((tg >> 2) << 2)) << 4) + (j & 1) * 8
                    for k in range(din // 32):
                        dst = (o0 ^ 64 if k & 1 else o0) + (k >> 1) * 1024
                        for plane in range(2):
                            for e in range(4):
                                c = 4 * (j + 8 * k) + e
                                off = g * gb + plane * nkb * 1024 + dst + e * 2
                                r = g * 8 + tg
                                kb = c // 64
                                base = plane * nkb * 1024 + kb * 1024
                                # logits product: A operand K-major, rows = tokens, SBO = group bytes
                                assert off == base + canon_kmajor(r, c % 64, gb)
                                # weighted-sum product: same bytes as MN-major (MN = channels, K rows = tokens), LBO = 1 KB
                                assert off == plane * nkb * 1024 + canon_mnmajor(c, r, 1024, gb)
                                assert off not in seen
                                seen[off] = (r, c, plane)
    assert len(seen) == 128 * din * 2       # every (token, channel, plane) exactly once
    # half-warp bank check of the 8-byte stores: the two tokens of a half-warp use disjoint 64-byte halves of the banks
    for sub_pair in ((0, 1), (2, 3)):
        halves = set()
        for sub in sub_pair:
            tg = 4 * (sub & 1) + (sub >> 1)
            chunk = (((0 >> 1) ^ (tg & 3)) | ((tg >> 2) << 2))
            halves.add(chunk >> 2)
        assert halves == {0, 1}


@pytest.mark.parametrize('sp', [16, 32])
def test_slot_attention_a_operand_paired_stores(sp):
    """softmax threads (token r of the tile) write a = softmax + eps for slot pairs (s, s + 4) as 32-bit words"""
    seen = set()
    for r in range(128):
        col, odd = r & 63, r & 1
        a_lane = (r >> 6) * (2 * sp * 128) + (4 * 128 if odd else 0) + (((col >> 3) ^ (4 if odd else 0)) << 4) + ((col & 7) >> 1) * 4
        for pq in range(sp // 2):
            s = (pq & 3) + 8 * (pq >> 2)
            dst = (a_lane ^ ((s & 3) << 4)) + s * 128
            row = s + 4 if odd else s                   # odd lanes own slot rows s + 4
            for plane in range(2):
                for half, tok in enumerate((r & ~1, r | 1)):
                    off = dst + plane * sp * 128 + half * 2
                    want = (tok >> 6) * (2 * sp * 128) + canon_kmajor(row + plane * sp, tok & 63, 1024)
                    assert off == want
                    seen.add((off, row + plane * sp, tok))
    assert len(seen) == 128 * sp * 2


def test_attention_row_chunks():
    """attention_tc.cu sw_chunk(tile, r, c): 16-byte chunk c of 128-byte row r"""
    for r in range(128):
        for c in range(8):
            off = r * 128 + ((c ^ (r & 7)) << 4)
            assert off == canon_kmajor(r, c * 8, 1024)
            # V' tile read as the MN-major B operand of the second product: MN = [v_hi | v_lo] column, K row = key
            assert off == canon_mnmajor(c * 8, r, 1024, 1024)
