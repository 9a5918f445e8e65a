"""GPU parity of every C-ABI kernel against a plain torch reference of the same op (fp64 on device).
Tolerances: 3-pass split-fp16 products are fp32-faithful (<= 2e-6 relative L2); 1-pass is fp16-class."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(180)]


@pytest.fixture(scope='module')
def ops():
    from slotdiffusion_b200 import ops as o
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(*shape, generator=g, device='cuda') * scale


def test_pack_roundtrip(ops):
    x = rnd(300, 192, seed=1, scale=3.0)
    p = ops.pack_rows(x)
    assert rel_l2(p.unpack(), x) < 1e-6
    w = rnd(96, 64, seed=2, scale=0.02)
    assert rel_l2(ops.pack_weight(w).unpack(), w) < 2e-6
    p = ops.pack_rows(x, act=1)
    assert rel_l2(p.unpack(), F.silu(x.double())) < 2e-6
    p = ops.pack_rows(x, act=2)
    assert rel_l2(p.unpack(), F.relu(x.double())) < 1e-6


@pytest.mark.parametrize('M,N,K', [(128, 128, 64), (256, 128, 128), (704, 576, 192), (64, 512, 128), (1000, 200, 72),
                                   (4096, 384, 1024), (33, 48, 512), (65536, 128, 128)])
def test_gemm_plain(ops, M, N, K):
    a = rnd(M, K, seed=3)
    w = rnd(N, K, seed=4, scale=K ** -0.5)
    bias = rnd(N, seed=5)
    res = rnd(M, N, seed=6)
    ref = a.double() @ w.double().t() + bias.double() + res.double()
    c = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, residual=res)
    assert rel_l2(c, ref) < 5e-6
    c1 = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, residual=res, passes=1)
    assert rel_l2(c1, ref) < 2e-3
    c = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), relu=True)
    assert rel_l2(c, F.relu(a.double() @ w.double().t())) < 5e-6


@pytest.fixture
def gemm_env():
    """Force a GEMM variant through the library's env knobs (read per call); always restored."""
    import os
    saved = {k: os.environ.get(k) for k in ('SDB_GEMM_CG', 'SDB_GEMM_SPLITK')}

    def setenv(cg=None, splitk=None):
        for k, v in (('SDB_GEMM_CG', cg), ('SDB_GEMM_SPLITK', splitk)):
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
    yield setenv
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('splitk', [1, 3])
@pytest.mark.parametrize('M,N,K', [(1000, 200, 192), (4096, 384, 1024), (130, 512, 512), (2048, 1152, 256)])
def test_gemm_variants(ops, gemm_env, cg, splitk, M, N, K):
    """Single-CTA and CTA-pair (cta_group::2) kernels, with and without split-K, ragged M / N edges."""
    gemm_env(cg, splitk)
    a = rnd(M, K, seed=31)
    w = rnd(N, K, seed=32, scale=K ** -0.5)
    bias = rnd(N, seed=33)
    res = rnd(M, N, seed=34)
    big = rnd(M // 8 + 1, 2 * N, seed=35)
    rv = big[:, N:]
    ref = a.double() @ w.double().t() + bias.double() + res.double() + rv.double().repeat_interleave(8, 0)[:M]
    c = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, residual=res, rowvec=rv, rows_per_group=8)
    assert rel_l2(c, ref) < 5e-6


@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_scalar_epilogue(ops, gemm_env, cg):
    """N % 4 != 0 / unaligned output view -> scalar epilogue path."""
    gemm_env(cg, None)
    M, N, K = 300, 30, 64
    a, w, bias = rnd(M, K, seed=36), rnd(N, K, seed=37, scale=K ** -0.5), rnd(N, seed=38)
    out = torch.zeros(M, N + 3, device='cuda')
    c = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, out=out[:, 1:N + 1])
    ref = a.double() @ w.double().t() + bias.double()
    assert rel_l2(c, ref) < 5e-6
    assert out[:, 0].abs().max() == 0 and out[:, N + 1:].abs().max() == 0


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('splitk', [1, 4])
def test_gemm_conv3_variants(ops, gemm_env, cg, splitk):
    gemm_env(cg, splitk)
    B, H, W, C, Cout = 5, 8, 8, 384, 512
    x = rnd(B, C, H, W, seed=10)
    w = rnd(Cout, C, 3, 3, seed=11, scale=(9 * C) ** -0.5)
    bias = rnd(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    xh = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    c = ops.gemm(ops.pack_rows(xh), ops.pack_weight_conv3(w), bias=bias, conv=(ops.SDB_A_CONV3, B, H, W, C))
    assert rel_l2(c, ref) < 2e-5


@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_packed_output_and_geglu(ops, gemm_env, cg):
    gemm_env(cg, None)
    M, K, Fh = 520, 256, 512
    a = rnd(M, K, seed=41)
    w = rnd(2 * Fh, K, seed=42, scale=K ** -0.5)
    bias = rnd(2 * Fh, seed=43)
    u = a.double() @ w.double().t() + bias.double()
    ref = u[:, :Fh] * F.gelu(u[:, Fh:])
    wg, bg = ops.pack_weight_geglu(w, bias)
    out = ops.gemm(ops.pack_rows(a), wg, bias=bg, geglu=True)
    assert out.rows == M and out.K == Fh
    assert rel_l2(out.unpack(), ref) < 5e-6
    # packed copy with activation, with and without the fp32 result
    w2 = rnd(192, K, seed=44, scale=K ** -0.5)
    b2 = rnd(192, seed=45)
    r2 = a.double() @ w2.double().t() + b2.double()
    c, pk = ops.gemm(ops.pack_rows(a), ops.pack_weight(w2), bias=b2, pack_out='silu')
    assert rel_l2(c, r2) < 5e-6 and rel_l2(pk.unpack(), F.silu(r2)) < 5e-6
    c, pk = ops.gemm(ops.pack_rows(a), ops.pack_weight(w2), bias=b2, relu=True, pack_out='none', keep_c=False)
    assert c is None and rel_l2(pk.unpack(), F.relu(r2)) < 5e-6


@pytest.mark.parametrize('cg', [1, 2])
@pytest.mark.parametrize('B,HW,C1,C2', [(3, 256, 128, 0), (5, 64, 512, 384), (9, 16, 384, 256), (2, 1024, 128, 128)])
def test_gemm_groupnorm_partial_sums(ops, gemm_env, cg, B, HW, C1, C2):
    """GN statistics accumulated by the producing GEMM epilogues == statistics of the stats kernel / torch."""
    gemm_env(cg, None)
    K = 128
    xs, gss = [], []
    for i, C in enumerate([C1, C2]):
        if C == 0:
            xs.append(None)
            gss.append(None)
            continue
        a = rnd(B * HW, K, seed=51 + i)
        w = rnd(C, K, seed=53 + i, scale=K ** -0.5)
        bias = rnd(C, seed=55 + i) * 2
        gs = torch.zeros(B, C // 4, 2, device='cuda')
        xs.append(ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, gsum=gs, rows_per_group=HW))
        gss.append(gs)
    C = C1 + C2
    gamma, beta = rnd(C, seed=57), rnd(C, seed=58)
    out = ops.groupnorm_pack_fused(xs[0], xs[1], gamma, beta, B, HW, 32, 1e-5, True, gsum1=gss[0], gsum2=gss[1])
    xcat = xs[0] if C2 == 0 else torch.cat([xs[0], xs[1]], 1)
    xr = xcat.view(B, HW, C).permute(0, 2, 1).double()
    ref = F.silu(F.group_norm(xr, 32, gamma.double(), beta.double(), 1e-5)).permute(0, 2, 1).reshape(B * HW, C)
    assert rel_l2(out.unpack(), ref) < 5e-6
    st = ops.groupnorm_finalize(gss[0], gss[1], C1, C2, B, HW, 32, 1e-5)
    st_ref = ops.groupnorm_stats(xs[0], xs[1], B, HW, 32, 1e-5)
    assert rel_l2(st, st_ref) < 1e-5
    out2 = ops.groupnorm_pack_fused(xs[0], xs[1], gamma, beta, B, HW, 32, 1e-5, True, stats=st_ref)
    assert rel_l2(out2.unpack(), ref) < 5e-6


def test_gemm_rowvec(ops):
    B, HW, K, N = 3, 64, 128, 256
    a = rnd(B * HW, K, seed=7)
    w = rnd(N, K, seed=8, scale=K ** -0.5)
    big = rnd(B, 3 * N, seed=9)
    rv = big[:, N:2 * N]                      # strided view like the fused emb GEMM output
    ref = (a.double() @ w.double().t()).view(B, HW, N) + rv.double()[:, None]
    c = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), rowvec=rv, rows_per_group=HW)
    assert rel_l2(c.view(B, HW, N), ref) < 5e-6


@pytest.mark.parametrize('B,H,W,C,Cout', [(2, 32, 32, 128, 128), (3, 16, 16, 256, 384), (5, 8, 8, 384, 128),
                                          (11, 4, 4, 512, 512), (2, 32, 32, 1024 // 4, 128), (1, 56, 56, 64, 64)])
def test_gemm_conv3(ops, B, H, W, C, Cout):
    x = rnd(B, C, H, W, seed=10)
    w = rnd(Cout, C, 3, 3, seed=11, scale=(9 * C) ** -0.5)
    bias = rnd(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, Cout)
    xh = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()       # NHWC rows
    a = ops.pack_rows(xh)
    c = ops.gemm(a, ops.pack_weight_conv3(w), bias=bias, conv=(ops.SDB_A_CONV3, B, H, W, C))
    assert rel_l2(c, ref) < 2e-5      # fp32 TMEM accumulation over K = 9*C up to 4608


@pytest.mark.parametrize('B,H,W,C', [(2, 32, 32, 128), (3, 16, 16, 256), (5, 8, 8, 384)])
def test_gemm_conv3_stride2(ops, B, H, W, C):
    x = rnd(B, C, H, W, seed=13)
    w = rnd(C, C, 3, 3, seed=14, scale=(9 * C) ** -0.5)
    bias = rnd(C, seed=15)
    ref = F.conv2d(x.double(), w.double(), bias.double(), stride=2, padding=1).permute(0, 2, 3, 1).reshape(-1, C)
    xh = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    a = ops.pack_nhwc(xh, None, B, H, W, mode=ops.SDB_PACK_PHASE2)
    c = ops.gemm(a, ops.pack_weight_conv3(w), bias=bias, conv=(ops.SDB_A_CONV3S2, B, H // 2, W // 2, C))
    assert rel_l2(c, ref) < 2e-5


def test_pack_nhwc_up2_and_concat(ops):
    B, H, W, C1, C2 = 2, 8, 8, 64, 128
    x1, x2 = rnd(B * H * W, C1, seed=16), rnd(B * H * W, C2, seed=17)
    cat = torch.cat([x1, x2], 1)
    p, ycat = ops.pack_nhwc(x1, x2, B, H, W, want_cat=True)
    assert torch.equal(ycat, cat)
    assert rel_l2(p.unpack(), cat) < 1e-6
    up = ops.pack_nhwc(x1, None, B, H, W, mode=ops.SDB_PACK_UP2)
    ref = F.interpolate(x1.view(B, H, W, C1).permute(0, 3, 1, 2), scale_factor=2, mode='nearest')
    assert rel_l2(up.unpack(), ref.permute(0, 2, 3, 1).reshape(-1, C1)) < 1e-6


@pytest.mark.parametrize('C', [192, 256, 384, 512])
def test_layernorm_pack(ops, C):
    x = rnd(777, C, seed=18, scale=2.0) + 0.5
    g, b = rnd(C, seed=19) * 0.1 + 1, rnd(C, seed=20) * 0.1
    p, y = ops.layernorm_pack(x, g, b, 1e-5, want_fp32=True)
    ref = F.layer_norm(x.double(), (C,), g.double(), b.double(), 1e-5)
    assert rel_l2(y, ref) < 2e-6
    assert rel_l2(p.unpack(), ref) < 2e-6


@pytest.mark.parametrize('C1,C2,HW,silu,eps', [(128, 0, 1024, True, 1e-5), (512, 384, 64, True, 1e-5),
                                               (256, 0, 256, False, 1e-6), (384, 256, 16, True, 1e-5)])
def test_groupnorm_pack(ops, C1, C2, HW, silu, eps):
    B = 3
    x1 = rnd(B * HW, C1, seed=21, scale=1.5) + 0.3
    x2 = rnd(B * HW, C2, seed=22) if C2 else None
    C = C1 + C2
    g, b = rnd(C, seed=23) * 0.1 + 1, rnd(C, seed=24) * 0.1
    p = ops.groupnorm_pack(x1, x2, g, b, B, HW, 32, eps, silu)
    x = torch.cat([x1, x2], 1) if C2 else x1
    xn = x.view(B, HW, C).permute(0, 2, 1).double()
    ref = F.group_norm(xn, 32, g.double(), b.double(), eps)
    if silu:
        ref = F.silu(ref)
    assert rel_l2(p.unpack(), ref.permute(0, 2, 1).reshape(-1, C)) < 3e-6


def test_geglu_and_timestep(ops):
    u = rnd(100, 2048, seed=25, scale=2.0)
    ref = u[:, :1024].double() * F.gelu(u[:, 1024:].double())
    assert rel_l2(ops.geglu_pack(u).unpack(), ref) < 2e-6
    t = torch.tensor([0.0, 1.0, 333.25, 998.999, 999.0], device='cuda')
    from oracle.unet_ref import timestep_embedding
    ref = timestep_embedding(t.cpu(), 128)
    got = ops.timestep_embedding_pack(t, 128).unpack().cpu()
    assert (got - ref).abs().max().item() < 2e-4     # fp32 sin/cos of arguments up to 1e3: ~1e-4 abs agreement
    ti = torch.tensor([3, 500], device='cuda')
    got = ops.timestep_embedding_pack(ti, 128).unpack().cpu()
    assert (got - timestep_embedding(ti.cpu(), 128)).abs().max().item() < 2e-4


@pytest.mark.parametrize('tc', [True, False])
@pytest.mark.parametrize('Lq,Lk,heads', [(256, 256, 8), (64, 64, 12), (16, 16, 16), (256, 11, 8), (64, 11, 12),
                                          (16, 11, 16), (100, 37, 4), (1024, 1024, 4), (300, 130, 2), (256, 24, 8),
                                          (3136, 7, 4)])
def test_attention(ops, Lq, Lk, heads, tc):
    """attention core through the tensor-core kernel (tc) and the CUDA-core kernel, vs fp64 torch math"""
    B, d = 2, 32
    C = heads * d
    qkv = rnd(B * Lq, 3 * C, seed=26)
    kv = rnd(B * Lk, 2 * C, seed=27)
    q = qkv[:, :C]
    k, v = kv[:, :C], kv[:, C:]
    out = ops.attention_pack(q, k, v, B, Lq, Lk, heads, d, d ** -0.5, tc=tc).unpack()
    qd = q.double().view(B, Lq, heads, d).transpose(1, 2)
    kd = k.double().view(B, Lk, heads, d).transpose(1, 2)
    vd = v.double().view(B, Lk, heads, d).transpose(1, 2)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * d ** -0.5, -1) @ vd).transpose(1, 2).reshape(B * Lq, C)
    assert rel_l2(out, ref) < 3e-6


@pytest.mark.parametrize('Lq,Lk,heads,scale', [(256, 11, 8, 1.0), (64, 11, 12, 1.0), (16, 11, 16, 1.0), (256, 24, 8, 1.0),
                                                (3136, 7, 4, 1.0), (100, 32, 1, 1.0), (45, 1, 16, 1.0), (256, 11, 8, 6.0)])
def test_attention_fewkeys(ops, Lq, Lk, heads, scale):
    """slot cross-attention (Lk = num_slots keys; csrc/attention_fewkeys.cu, the default route of attention_pack for Lk <= 32):
    vs fp64 torch math, ragged row tiles, q as a column slice of a wider buffer, peaked logits; the other two kernels agree"""
    B, d = 3, 32
    C = heads * d
    qkv = rnd(B * Lq, 3 * C, seed=26, scale=scale)
    kv = rnd(B * Lk, 2 * C, seed=27, scale=scale)
    q = qkv[:, C:2 * C]
    k, v = kv[:, :C], kv[:, C:]
    from slotdiffusion_b200._lib import lib
    assert lib().sdb_attention_fewkeys_supported(heads, d, Lk, q.stride(0), k.stride(0), v.stride(0))
    out = ops.attention_pack(q, k, v, B, Lq, Lk, heads, d, d ** -0.5).unpack()
    qd = q.double().view(B, Lq, heads, d).transpose(1, 2)
    kd = k.double().view(B, Lk, heads, d).transpose(1, 2)
    vd = v.double().view(B, Lk, heads, d).transpose(1, 2)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * d ** -0.5, -1) @ vd).transpose(1, 2).reshape(B * Lq, C)
    assert rel_l2(out, ref) < 2e-6
    assert rel_l2(out, ops.attention_pack(q, k, v, B, Lq, Lk, heads, d, d ** -0.5, tc=True).unpack()) < 3e-6
    assert not lib().sdb_attention_fewkeys_supported(heads, d, 33, q.stride(0), k.stride(0), v.stride(0))
    assert not lib().sdb_attention_fewkeys_supported(heads, 64, Lk, q.stride(0), k.stride(0), v.stride(0))


def test_attention_tc_peaked(ops):
    """large logits (peaked softmax) and a self-attention view of a fused q|k|v projection"""
    B, L, heads, d = 3, 256, 8, 32
    C = heads * d
    qkv = rnd(B * L, 3 * C, seed=50, scale=4.0)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    out = ops.attention_pack(q, k, v, B, L, L, heads, d, d ** -0.5, tc=True).unpack()
    qd = q.double().view(B, L, heads, d).transpose(1, 2)
    kd = k.double().view(B, L, heads, d).transpose(1, 2)
    vd = v.double().view(B, L, heads, d).transpose(1, 2)
    ref = (torch.softmax(qd @ kd.transpose(-1, -2) * d ** -0.5, -1) @ vd).transpose(1, 2).reshape(B * L, C)
    assert rel_l2(out, ref) < 3e-6


@pytest.mark.parametrize('B,H,W', [(3, 32, 32), (2, 7, 8), (1, 56, 56)])
def test_conv_in_out(ops, B, H, W):
    x = rnd(B, 3, H, W, seed=28)
    w, b = rnd(128, 3, 3, 3, seed=29, scale=0.2), rnd(128, seed=30)
    y = ops.conv3_in(x, w, b)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1).permute(0, 2, 3, 1).reshape(-1, 128)
    assert rel_l2(y, ref) < 2e-6
    h = rnd(B * H * W, 128, seed=31, scale=2.0)
    g, be = rnd(128, seed=32) * 0.1 + 1, rnd(128, seed=33) * 0.1
    w2, b2 = rnd(3, 128, 3, 3, seed=34, scale=0.05), rnd(3, seed=35)
    stats = ops.groupnorm_stats(h, None, B, H * W, 32, 1e-5)
    y = ops.conv3_out(h, stats, g, be, w2, b2, B, H, W)
    hn = h.view(B, H, W, 128).permute(0, 3, 1, 2).double()
    ref = F.conv2d(F.silu(F.group_norm(hn, 32, g.double(), be.double(), 1e-5)), w2.double(), b2.double(), padding=1)
    assert rel_l2(y, ref) < 3e-6


@pytest.mark.parametrize('B,N,S,D', [(2, 1024, 11, 192), (1, 1024, 24, 192), (3, 196, 7, 256), (5, 77, 5, 192),
                                     (64, 1024, 11, 192), (2, 100, 15, 64)])
def test_slot_attend(ops, B, N, S, D):
    kv = rnd(B * N, 2 * D, seed=36)
    q = rnd(B * S, D, seed=37)
    scale, eps = D ** -0.5, 1e-6
    upd, mask, upd32 = ops.slot_attend(kv, q, B, N, S, D, scale, eps, want_mask=True, want_fp32=True)
    k = kv[:, :D].double().view(B, N, D)
    v = kv[:, D:].double().view(B, N, D)
    attn = torch.softmax(scale * k @ q.double().view(B, S, D).transpose(1, 2), -1)
    a = attn + eps
    a = a / a.sum(1, keepdim=True)
    ref = torch.einsum('bns,bnd->bsd', a, v).reshape(B * S, D)
    assert rel_l2(mask, attn.transpose(1, 2)) < 3e-6
    assert rel_l2(upd32, ref) < 3e-6
    assert rel_l2(upd.unpack(), ref) < 3e-6


@pytest.mark.parametrize('B,N,S,Din', [(2, 1024, 11, 192), (1, 1024, 24, 192), (3, 196, 7, 256), (5, 77, 5, 192),
                                       (64, 1024, 11, 192), (2, 300, 15, 128), (1, 784, 7, 256), (160, 1024, 11, 192),
                                       (2, 130, 20, 256)])
def test_slot_attend_fused(ops, B, N, S, Din):
    """Tensor-core attend over raw features: in-kernel LayerNorm (no affine), logits = n . qa[:, :Din] + qa[:, Din],
    softmax over slots, U = (a^T n) / sum_n a  -- vs fp64 torch math of the same definition."""
    x = rnd(B, N, Din, seed=46, scale=2.0) + 0.3
    qa = torch.zeros(B * S, Din + 4, device='cuda')
    qa[:, :Din] = rnd(B * S, Din, seed=47, scale=Din ** -0.5)
    qa[:, Din] = rnd(B * S, seed=48)
    ln_eps, eps = 1e-5, 1e-6
    upd, mask, upd32 = ops.slot_attend_fused(x, qa, B, N, S, Din, ln_eps, eps, want_mask=True, want_fp32=True)
    xd = x.double()
    n = (xd - xd.mean(-1, keepdim=True)) / torch.sqrt(xd.var(-1, unbiased=False, keepdim=True) + ln_eps)
    qd = qa.double().view(B, S, Din + 4)
    logits = n @ qd[:, :, :Din].transpose(1, 2) + qd[:, :, Din][:, None, :]
    attn = torch.softmax(logits, -1)
    a = attn + eps
    ref = torch.einsum('bns,bnd->bsd', a / a.sum(1, keepdim=True), n).reshape(B * S, Din)
    assert rel_l2(mask, attn.transpose(1, 2)) < 3e-6
    assert rel_l2(upd32, ref) < 6e-6          # fp32 accumulation over up to 1024 tokens per CTA
    assert rel_l2(upd.unpack(), ref) < 6e-6
    # determinism (no atomics): bit-identical on a second launch
    upd_b, mask_b, upd32_b = ops.slot_attend_fused(x, qa, B, N, S, Din, ln_eps, eps, want_mask=True, want_fp32=True)
    assert torch.equal(upd32, upd32_b) and torch.equal(mask, mask_b)


def test_gru_gates(ops):
    R, D = 77, 192
    gi, gh, h = rnd(R, 3 * D, seed=38), rnd(R, 3 * D, seed=39), rnd(R, D, seed=40)
    out = ops.gru_gates(gi, gh, h)
    a, b = gi.double(), gh.double()
    r = torch.sigmoid(a[:, :D] + b[:, :D])
    z = torch.sigmoid(a[:, D:2 * D] + b[:, D:2 * D])
    n = torch.tanh(a[:, 2 * D:] + r * b[:, 2 * D:])
    assert rel_l2(out, (1 - z) * n + z * h.double()) < 2e-6


def test_dpm_glue(ops):
    from oracle import dpm_ref
    B = 4
    x, eps = rnd(B, 3, 32, 32, seed=41), rnd(B, 3, 32, 32, seed=42)
    cb = rnd(4096, 3, seed=43)
    x0, idx = ops.dpm_x0(x, eps, 0.8, 0.6, cb, want_idx=True)
    z = (x.cpu() - 0.6 * eps.cpu()) / 0.8
    zq, ridx = dpm_ref.vq_quantize(z, cb.cpu())
    # index parity is exact outside fp64 near-ties (bisector distance below the fp32 round-off of z itself, 4 ulp)
    from helpers import vq_mismatch
    zf = z.permute(0, 2, 3, 1).reshape(-1, 3)
    real, near, ref64 = vq_mismatch(idx, zf, cb, 4 * 1.2e-7 * zf.abs().amax(1).clamp_min(1.0))
    assert real == 0, (real, near)
    assert near <= 2
    same = (idx.cpu().view(B, 32, 32) == ridx)[:, None].expand_as(zq)
    assert same.float().mean().item() > 0.999
    assert torch.equal(x0.cpu()[same], zq[same])
    # the same bar applied to the reference's own fp32 argmin: it also only deviates from fp64 on near-ties
    real_r, _, _ = vq_mismatch(ridx.flatten(), zf, cb, 4 * 1.2e-7 * zf.abs().amax(1).clamp_min(1.0))
    assert real_r == 0
    x0n = ops.dpm_x0(x, eps, 0.8, 0.6, None)
    assert rel_l2(x0n, z) < 1e-6
    m0, m1 = rnd(B, 3, 32, 32, seed=44), rnd(B, 3, 32, 32, seed=45)
    y = ops.lincomb(x, m0, m1, 0.5, -0.25, 2.0)
    assert rel_l2(y, 0.5 * x - 0.25 * m0 + 2.0 * (m1 - m0)) < 1e-6


@pytest.mark.parametrize('ncodes', [1, 7, 8, 13, 100, 512])
def test_vq_grouped_search_ragged_codebooks_and_first_minimum(ops, ncodes):
    """the nearest-code search scores the codebook in groups of 8 (csrc/elementwise.cu dpm_x0_vq3_kernel): sizes that are
    not a multiple of the group, and DUPLICATED codes -- argmin's rule is the FIRST minimum (quantize.py:84-94), inside a group,
    across groups and in the tail"""
    B = 2
    x, eps = rnd(B, 3, 16, 16, seed=71), rnd(B, 3, 16, 16, seed=72)
    cb = rnd(ncodes, 3, seed=73)
    cb = torch.cat([cb, cb, cb[: max(1, ncodes // 2)]], 0).contiguous()          # every code appears again later
    x0, idx = ops.dpm_x0(x, eps, 0.8, 0.6, cb, want_idx=True)
    z = ((x - 0.6 * eps) / 0.8).permute(0, 2, 3, 1).reshape(-1, 3)
    # the kernel's own distance expression, evaluated in fp32 on the GPU in the same order (fmaf chain), first minimum
    zz = (z[:, 0] * z[:, 0] + z[:, 1] * z[:, 1]) + z[:, 2] * z[:, 2]
    assert (idx.flatten() < ncodes).all()                        # a duplicate further on never wins
    d64 = torch.cdist(z.double(), cb.double())
    pick = d64.gather(1, idx.flatten().long()[:, None])[:, 0]
    assert (pick - d64.min(1).values <= 1e-5 * (1 + zz.double().sqrt())).all()   # and it is a nearest code
    assert torch.equal(x0.permute(0, 2, 3, 1).reshape(-1, 3), cb[idx.flatten().long()])


@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_groupnorm_sums_two_channel_blocks(ops, gemm_env, cg):
    """gsum_cb = 2: per-(sample, channel PAIR) sums for the 64-channel / 32-group layers (ResNet stem, VQ-VAE level 0)"""
    gemm_env(cg, None)
    B, HW, K, C = 3, 4096, 64, 64
    a, w, bias = rnd(B * HW, K, seed=61), rnd(C, K, seed=62, scale=K ** -0.5), rnd(C, seed=63) * 2
    gs = torch.zeros(B, C // 2, 2, device='cuda')
    x = ops.gemm(ops.pack_rows(a), ops.pack_weight(w), bias=bias, gsum=gs, gsum_cb=2, rows_per_group=HW)
    st = ops.groupnorm_finalize_cb(gs, C, B, HW, 32, 1e-5, 2)
    assert rel_l2(st, ops.groupnorm_stats(x, None, B, HW, 32, 1e-5)) < 1e-5
    xr = x.view(B, HW, C // 2, 2).double()
    assert rel_l2(gs[..., 0], xr.sum((1, 3))) < 1e-5 and rel_l2(gs[..., 1], (xr * xr).sum((1, 3))) < 1e-5


@pytest.mark.parametrize('cg', [1, 2])
def test_gemm_block_diagonal_batches(ops, gemm_env, cg):
    """SdbGemm.batch_rows: S_b = q_b k_b^T (W advances by rows) and O_b = P_b v_b with v^T stored [C, B*L] (W advances by
    columns) in one launch each -- the single-head attention of the VQ-VAE AttnBlock"""
    gemm_env(cg, None)
    B, L, C = 3, 512, 256
    q, k, v = rnd(B * L, C, seed=64), rnd(B * L, C, seed=65), rnd(B * L, C, seed=66)
    s = ops.gemm(ops.pack_rows(q), ops.pack_rows(k), batch=(L, L, L, 0))
    ref_s = torch.einsum('blc,bmc->blm', q.view(B, L, C).double(), k.view(B, L, C).double()).reshape(B * L, L)
    assert s.shape == (B * L, L) and rel_l2(s, ref_s) < 5e-6
    p = ops.softmax_pack(s, C ** -0.5)                       # 2^12 P: keeps the lo plane of ~1e-3 probabilities normal
    ref_p = torch.softmax(ref_s * C ** -0.5, dim=-1)
    assert rel_l2(p.unpack() / ops.SOFTMAX_PACK_SCALE, ref_p) < 5e-6
    vt = ops.transpose_packed(ops.pack_rows(v))
    o = ops.gemm(p, vt, batch=(L, C, 0, L), alpha=1.0 / ops.SOFTMAX_PACK_SCALE)
    ref_o = torch.einsum('blm,bmc->blc', ref_p.view(B, L, L), v.view(B, L, C).double()).reshape(B * L, C)
    assert o.shape == (B * L, C) and rel_l2(o, ref_o) < 5e-6
