"""GPU parity of the backward (training) kernels against torch autograd in fp64 on the same inputs.
Every check goes through the C ABI (slotdiffusion_b200.ops wrappers).  Tolerance: fp32-faithful (<= 2e-5 rel-L2)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import rel_l2

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]
TOLB = 3e-5


@pytest.fixture(scope='module')
def ops():
    from slotdiffusion_b200 import ops as o
    return o


def rnd(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.randn(*shape, generator=g, device='cuda') * scale


@pytest.mark.parametrize('M,N,rpg', [(256, 128, 64), (1000, 200, 8), (130, 48, 16), (4096, 384, 1024)])
def test_grad_pack(ops, M, N, rpg):
    big = rnd(M, N + 8, seed=1)
    dy = big[:, 4:N + 4] if N % 4 == 0 else big[:, :N]
    bias = torch.ones(N, device='cuda')
    grp = torch.zeros((M + rpg - 1) // rpg, N, device='cuda')
    rows, tr = ops.grad_pack(dy, bias_grad=bias, group_grad=grp, rows_per_group=rpg)
    assert rel_l2(rows.unpack(), dy) < 2e-5 and rel_l2(tr.unpack()[:, :M], dy.t()) < 2e-5      # bf16 x 2 split
    tiny = ops.grad_pack(dy * 1e-9)[0].unpack()                                                  # no underflow
    assert rel_l2(tiny, dy.double() * 1e-9) < 2e-5
    assert rel_l2(bias, 1 + dy.double().sum(0)) < 1e-6
    pad = (-M) % rpg
    ref = F.pad(dy.double(), (0, 0, 0, pad)).view(-1, rpg, N).sum(1)
    assert rel_l2(grp, ref) < 1e-6


def test_transpose_packed(ops):
    x = rnd(300, 200, seed=2)
    p = ops.pack_rows(x)
    t = ops.transpose_packed(p)
    assert t.rows == 200 and t.K == 304            # contraction length padded to a multiple of 8, zero filled
    assert torch.equal(t.t.view(2, 200, 304)[:, :, :300], p.t.view(2, 300, 200).transpose(1, 2).contiguous())
    assert t.t.view(2, 200, 304)[:, :, 300:].abs().max() == 0


@pytest.mark.parametrize('M,N,K', [(704, 576, 192), (4096, 384, 1024), (1000, 200, 72), (15, 192, 64), (22, 64, 192)])
def test_linear_backward(ops, M, N, K):
    a, w, dy = rnd(M, K, seed=3), rnd(N, K, seed=4, scale=K ** -0.5), rnd(M, N, seed=5)
    dyp, dyT = ops.grad_pack(dy)
    da = ops.gemm(dyp, ops.pack_weight_T(w))
    assert rel_l2(da, dy.double() @ w.double()) < TOLB
    dw = ops.gemm(dyT, ops.transpose_packed(ops.pack_rows(a), to_bf16=True))
    assert rel_l2(dw, dy.double().t() @ a.double()) < TOLB


@pytest.mark.parametrize('B,H,W,C,Cout', [(2, 32, 32, 128, 128), (3, 16, 16, 256, 384), (5, 8, 8, 384, 128),
                                          (4, 8, 8, 64, 256), (7, 4, 4, 512, 512), (3, 16, 16, 192, 64),
                                          (2, 64, 64, 64, 64), (1, 128, 128, 64, 64), (2, 16, 128, 64, 128)])   # ResNet stem / layer1
def test_conv3_backward(ops, B, H, W, C, Cout):
    x = rnd(B, C, H, W, seed=6).double().requires_grad_()
    w = rnd(Cout, C, 3, 3, seed=7, scale=(9 * C) ** -0.5).double().requires_grad_()
    dy = rnd(B, Cout, H, W, seed=8)
    F.conv2d(x, w, padding=1).backward(dy.double())
    dy_rows = dy.permute(0, 2, 3, 1).reshape(-1, Cout).contiguous()
    dyp, dyT = ops.grad_pack(dy_rows)
    dx = ops.gemm(dyp, ops.pack_weight_conv3_dgrad(w.detach().float()), conv=(ops.SDB_A_CONV3, B, H, W, Cout))
    assert rel_l2(dx, x.grad.permute(0, 2, 3, 1).reshape(-1, C)) < TOLB
    xp = ops.pack_rows(x.detach().float().permute(0, 2, 3, 1).reshape(-1, C).contiguous())
    c9 = ops.gemm_wgrad_conv(xp, dyp, B, H, W, C)
    dw = torch.empty(Cout, C, 3, 3, device='cuda')
    ops.wgrad_conv3_scatter(c9, dw, C)
    assert rel_l2(dw, w.grad) < TOLB


@pytest.mark.parametrize('B,H,W,C', [(2, 32, 32, 128), (3, 16, 16, 256), (2, 128, 128, 64)])
def test_conv3_stride2_backward(ops, B, H, W, C):
    """H, W = input size; output H/2 x W/2."""
    x = rnd(B, C, H, W, seed=9).double().requires_grad_()
    w = rnd(C, C, 3, 3, seed=10, scale=(9 * C) ** -0.5).double().requires_grad_()
    Ho, Wo = H // 2, W // 2
    dy = rnd(B, C, Ho, Wo, seed=11)
    F.conv2d(x, w, stride=2, padding=1).backward(dy.double())
    dy_rows = dy.permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    z = ops.pack_zero_up2(dy_rows, B, Ho, Wo, C)
    dx = ops.gemm(z, ops.pack_weight_conv3_dgrad(w.detach().float()), conv=(ops.SDB_A_CONV3, B, H, W, C))
    assert rel_l2(dx, x.grad.permute(0, 2, 3, 1).reshape(-1, C)) < TOLB
    xh = x.detach().float().permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    xph = ops.pack_nhwc(xh, None, B, H, W, mode=ops.SDB_PACK_PHASE2)
    dyp, _ = ops.grad_pack(dy_rows, want_T=False)
    c9 = ops.gemm_wgrad_conv(xph, dyp, B, Ho, Wo, C, stride2=True)
    dw = torch.empty(C, C, 3, 3, device='cuda')
    ops.wgrad_conv3_scatter(c9, dw, C)
    assert rel_l2(dw, w.grad) < TOLB


@pytest.mark.parametrize('C1,C2,HW,silu,p', [(128, 0, 256, True, 0.0), (512, 384, 64, True, 0.0), (256, 0, 16, False, 0.0),
                                            (128, 128, 64, True, 0.1)])
def test_groupnorm_backward(ops, C1, C2, HW, silu, p):
    B, G, eps = 3, 32, 1e-5
    C = C1 + C2
    x1, x2 = rnd(B * HW, C1, seed=12), (rnd(B * HW, C2, seed=13) if C2 else None)
    gamma, beta = rnd(C, seed=14), rnd(C, seed=15)
    da = rnd(B * HW, C, seed=16)
    stats = ops.groupnorm_stats(x1, x2, B, HW, G, eps)
    seed = 1234
    if p > 0:
        out = ops.groupnorm_pack_dropout(x1, x2, gamma, beta, stats, B, HW, G, silu, p, seed).unpack()
        nodrop = ops.groupnorm_pack_fused(x1, x2, gamma, beta, B, HW, G, eps, silu, stats=stats).unpack()
        keep = out != 0
        assert abs(keep.float().mean().item() - (1 - p)) < 0.01
        assert rel_l2(out[keep], nodrop[keep] / (1 - p)) < 1e-6
        mask = keep.double() / (1 - p)
    else:
        mask = None
    xc = (x1 if x2 is None else torch.cat([x1, x2], 1)).double().requires_grad_()
    g64, b64 = gamma.double().requires_grad_(), beta.double().requires_grad_()
    y = F.group_norm(xc.view(B, HW, C).permute(0, 2, 1), G, g64, b64, eps).permute(0, 2, 1).reshape(B * HW, C)
    if silu:
        y = F.silu(y)
    if mask is not None:
        y = y * mask
    y.backward(da.double())
    dgamma, dbeta = torch.ones(C, device='cuda'), torch.ones(C, device='cuda')
    add1 = rnd(B * HW, C1, seed=17)
    dx1, dx2 = ops.groupnorm_bwd(x1, x2, da, stats, gamma, beta, dgamma, dbeta, B, HW, G, silu, add1=add1, drop_p=p,
                                 seed=seed)
    assert rel_l2(dx1, xc.grad[:, :C1] + add1.double()) < TOLB
    if C2:
        assert rel_l2(dx2, xc.grad[:, C1:]) < TOLB
    assert rel_l2(dgamma - 1, g64.grad) < TOLB and rel_l2(dbeta - 1, b64.grad) < TOLB


@pytest.mark.parametrize('C', [192, 256, 384, 512])
def test_layernorm_backward(ops, C):
    M = 777
    x, dn = rnd(M, C, seed=18), rnd(M, C, seed=19)
    gamma, beta = rnd(C, seed=20), rnd(C, seed=21)
    x64, g64, b64 = x.double().requires_grad_(), gamma.double().requires_grad_(), beta.double().requires_grad_()
    F.layer_norm(x64, (C,), g64, b64, 1e-5).backward(dn.double())
    dg, db = torch.zeros(C, device='cuda'), torch.zeros(C, device='cuda')
    add = rnd(M, C, seed=22)
    dx = ops.layernorm_bwd(x, dn, gamma, 1e-5, dg, db, add=add)
    assert rel_l2(dx, x64.grad + add.double()) < TOLB
    assert rel_l2(dg, g64.grad) < TOLB and rel_l2(db, b64.grad) < TOLB


@pytest.mark.parametrize('Lq,Lk,heads', [(256, 256, 8), (64, 11, 12), (16, 16, 16), (100, 7, 4)])
def test_attention_backward(ops, Lq, Lk, heads):
    B, d = 3, 32
    C = heads * d
    scale = d ** -0.5
    qkv = rnd(B * Lq, 3 * C, seed=23)
    kvc = rnd(B * Lk, 2 * C + 64, seed=24)
    q = qkv[:, :C]
    k, v = (qkv[:, C:2 * C], qkv[:, 2 * C:]) if Lk == Lq else (kvc[:, 32:32 + C], kvc[:, 32 + C:32 + 2 * C])
    do = rnd(B * Lq, C, seed=25)
    q64, k64, v64 = [t.double().clone().requires_grad_() for t in (q, k, v)]

    def heads_view(t, L):
        return t.view(B, L, heads, d).permute(0, 2, 1, 3)
    att = torch.softmax(heads_view(q64, Lq) @ heads_view(k64, Lk).transpose(-1, -2) * scale, -1)
    o = (att @ heads_view(v64, Lk)).permute(0, 2, 1, 3).reshape(B * Lq, C)
    o.backward(do.double())
    dq, dk, dv = torch.empty(B * Lq, C, device='cuda'), torch.empty(B * Lk, C, device='cuda'), torch.empty(B * Lk, C, device='cuda')
    ops.attention_bwd(q, k, v, do, dq, dk, dv, B, Lq, Lk, heads, d, scale)
    assert rel_l2(dq, q64.grad) < TOLB and rel_l2(dk, k64.grad) < TOLB and rel_l2(dv, v64.grad) < TOLB


def test_pointwise_backward(ops):
    M, Fh = 300, 512
    u, dg = rnd(M, 2 * Fh, seed=26), rnd(M, Fh, seed=27)
    u64 = u.double().requires_grad_()
    (u64[:, :Fh] * F.gelu(u64[:, Fh:])).backward(dg.double())
    assert rel_l2(ops.geglu_bwd(u, dg), u64.grad) < TOLB
    pre, dy = rnd(100, 64, seed=28), rnd(100, 64, seed=29)
    p64 = pre.double().requires_grad_()
    F.silu(p64).backward(dy.double())
    assert rel_l2(ops.act_bwd(dy, pre, 'silu'), p64.grad) < TOLB
    assert rel_l2(ops.act_bwd(dy, pre, 'relu'), dy.double() * (pre > 0)) < 1e-7
    a, b, c = rnd(64, 36, seed=30), rnd(64, 36, seed=31), rnd(64, 36, seed=32)
    assert torch.equal(ops.add3(a, b, c), (a + b) + c)
    B, H, W, C = 2, 4, 4, 64
    dup = rnd(B * 4 * H * W, C, seed=33)
    ref = F.avg_pool2d(dup.view(B, 2 * H, 2 * W, C).permute(0, 3, 1, 2).double(), 2) * 4
    assert rel_l2(ops.up2_adjoint(dup, B, H, W, C), ref.permute(0, 2, 3, 1).reshape(-1, C)) < 1e-6
    x = rnd(B, 3, 8, 8, seed=34)
    pk = ops.pack_nchw_pad(x, 64).unpack().view(B, 64, 64)
    assert rel_l2(pk[:, :, :3], x.view(B, 3, 64).transpose(1, 2)) < 1e-6 and pk[:, :, 3:].abs().max() == 0
    rows = rnd(B * 64, 16, seed=35)
    assert torch.equal(ops.nhwc_to_nchw(rows, B, 3, 8, 8), rows[:, :3].view(B, 64, 3).permute(0, 2, 1).reshape(B, 3, 8, 8))
    assert torch.equal(ops.nchw_to_nhwc_pad(x, 8)[:, :3], x.view(B, 3, 64).transpose(1, 2).reshape(-1, 3))


def test_gru_backward(ops):
    R, D = 77, 192
    gi, gh, h, dhn = rnd(R, 3 * D, seed=36), rnd(R, 3 * D, seed=37), rnd(R, D, seed=38), rnd(R, D, seed=39)
    gi64, gh64, h64 = [t.double().requires_grad_() for t in (gi, gh, h)]
    r = torch.sigmoid(gi64[:, :D] + gh64[:, :D])
    z = torch.sigmoid(gi64[:, D:2 * D] + gh64[:, D:2 * D])
    n = torch.tanh(gi64[:, 2 * D:] + r * gh64[:, 2 * D:])
    ((1 - z) * n + z * h64).backward(dhn.double())
    dgi, dgh, dh = ops.gru_gates_bwd(gi, gh, h, dhn)
    assert rel_l2(dgi, gi64.grad) < TOLB and rel_l2(dgh, gh64.grad) < TOLB and rel_l2(dh, h64.grad) < TOLB


@pytest.mark.parametrize('B,N,S,D', [(2, 1024, 11, 192), (3, 196, 7, 256), (2, 77, 24, 64)])
def test_slot_attend_backward(ops, B, N, S, D):
    scale, eps = D ** -0.5, 1e-6
    kv, q, dU = rnd(B * N, 2 * D, seed=40), rnd(B * S, D, seed=41), rnd(B * S, D, seed=42)
    kv64, q64 = kv.double().requires_grad_(), q.double().requires_grad_()
    k64, v64 = kv64.view(B, N, 2 * D)[..., :D], kv64.view(B, N, 2 * D)[..., D:]
    att = torch.softmax(scale * k64 @ q64.view(B, S, D).transpose(1, 2), -1) + eps
    att = att / att.sum(1, keepdim=True)
    (att.transpose(1, 2) @ v64).reshape(B * S, D).backward(dU.double())
    upd, _, upd32, cs = ops.slot_attend_train(kv, q, B, N, S, D, scale, eps, False)
    dkv = torch.empty_like(kv)
    dq = ops.slot_attend_bwd(kv, q, upd32, cs, dU, dkv, B, N, S, D, scale, eps, False)
    assert rel_l2(dkv, kv64.grad) < TOLB and rel_l2(dq, q64.grad) < TOLB
    dq2 = ops.slot_attend_bwd(kv, q, upd32, cs, dU, dkv, B, N, S, D, scale, eps, True)
    assert rel_l2(dkv, 2 * kv64.grad) < TOLB and rel_l2(dq2, q64.grad) < TOLB
