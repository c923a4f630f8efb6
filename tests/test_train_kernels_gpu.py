"""Backward / training kernels on the GPU (through the C ABI) against PyTorch fp32 autograd of the same op on the
same bf16-rounded inputs (the autograd of the diffusers blocks is what the reference's accelerator.backward runs,
train_models/train_svd_lora.py:1683)."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def rnd(*shape, dev, scale=1.0, dtype=bf16, seed=0):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed + (hash(shape) & 0xFFFF))
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev)


# ------------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("NS,R,C1,C2,silu", [(4, 100, 64, 0, True), (2, 257, 320, 0, False), (3, 64, 64, 32, True),
                                             (1, 1000, 640, 320, True)])
def test_groupnorm_bwd(cuda, NS, R, C1, C2, silu):
    from lkgd_b200 import ops
    M, C = NS * R, C1 + C2
    x1 = rnd(M, C1, dev=cuda, dtype=torch.float32, seed=1) * 1.5 + 0.3
    x2 = rnd(M, C2, dev=cuda, dtype=torch.float32, seed=2) if C2 else None
    gamma = rnd(C, dev=cuda, dtype=torch.float32, seed=3) * 0.3 + 1
    beta = rnd(C, dev=cuda, dtype=torch.float32, seed=4) * 0.2
    dy = rnd(M, C, dev=cuda, seed=5)
    add = rnd(M, C, dev=cuda, seed=6)
    eps = 1e-5
    y, stats = ops.groupnorm(x1, gamma, beta, eps, NS=NS, R=R, x2=x2, silu=silu, return_stats=True)
    # torch reference: channels-last rows -> [NS, C, R]
    xc = (torch.cat([x1, x2], 1) if C2 else x1).clone().requires_grad_(True)
    xr = xc.reshape(NS, R, C).permute(0, 2, 1)
    yr = F.group_norm(xr, 32, gamma, beta, eps)
    if silu:
        yr = F.silu(yr)
    yr = yr.permute(0, 2, 1).reshape(M, C)
    assert rel_l2(y.float(), yr) < 6e-3
    yr.backward(dy.float())
    ref = xc.grad + add.float()
    out1 = torch.full((M, C1), 0.25, device=cuda)
    out2 = torch.empty((M, C2), device=cuda) if C2 else None
    ob = torch.empty((M, C), device=cuda, dtype=bf16)
    ops.groupnorm_bwd(x1, dy, stats, gamma, beta, eps, NS=NS, R=R, x2=x2, silu=silu, add=add, out1=out1, acc1=True,
                      out2=out2, acc2=False, out_bf16=ob)
    got = torch.cat([out1 - 0.25, out2], 1) if C2 else out1 - 0.25
    assert rel_l2(got, ref) < 2e-4
    full = torch.cat([out1, out2], 1) if C2 else out1
    assert rel_l2(ob.float(), full) < 4e-3


@pytest.mark.parametrize("M,C,dy_f32", [(1000, 320, False), (77, 1280, True), (513, 64, False)])
def test_layernorm_bwd(cuda, M, C, dy_f32):
    from lkgd_b200 import ops
    x = (rnd(M, C, dev=cuda, dtype=torch.float32, seed=1) * 2 + 0.5).requires_grad_(True)
    gamma = rnd(C, dev=cuda, dtype=torch.float32, seed=2) * 0.3 + 1
    beta = rnd(C, dev=cuda, dtype=torch.float32, seed=3)
    dy = rnd(M, C, dev=cuda, dtype=torch.float32 if dy_f32 else bf16, seed=4)
    G0 = rnd(M, C, dev=cuda, dtype=torch.float32, seed=5)
    F.layer_norm(x, (C,), gamma, beta, 1e-5).backward(dy.float())
    G = G0.clone()
    gb = torch.empty((M, C), device=cuda, dtype=bf16)
    ops.layernorm_bwd(x.detach(), dy, gamma, 1e-5, G, accumulate=True, g_bf16=gb)
    assert rel_l2(G - G0, x.grad) < 1e-4
    assert rel_l2(gb.float(), G) < 4e-3
    G2 = torch.full_like(G0, 7.0)
    ops.layernorm_bwd(x.detach(), dy, gamma, 1e-5, G2, accumulate=False)
    assert rel_l2(G2, x.grad) < 1e-4


def test_geglu_fwd_bwd(cuda):
    from lkgd_b200 import ops
    M, C, H = 300, 64, 256                              # H = 4C
    x = rnd(M, C, dev=cuda, seed=1)
    W = rnd(2 * H, C, dev=cuda, scale=C ** -0.5, seed=2)
    b = rnd(2 * H, dev=cuda, dtype=torch.float32, seed=3)
    Wp, bp = ops.pack_geglu(W, b)
    pre = ops.gemm(x, Wp, bias=bp)                      # tile-interleaved [M, 2H]
    out = ops.geglu_fwd(pre)
    fused = ops.gemm(x, Wp, bias=bp, act=ops.ACT_GEGLU)
    assert rel_l2(out.float(), fused.float()) < 6e-3
    # torch on the natural order
    idx = torch.arange(H, device=cuda).view(-1, 128)
    order = torch.cat([idx, idx + H], 1).reshape(-1)
    nat = torch.empty_like(pre)
    nat[:, order] = pre                                  # undo the interleave: nat = [value | gate]
    natf = nat.float().requires_grad_(True)
    ref = natf[:, :H] * F.gelu(natf[:, H:])
    assert rel_l2(out.float(), ref) < 4e-3
    dout = rnd(M, H, dev=cuda, seed=4)
    ref.backward(dout.float())
    dpre = ops.geglu_bwd(pre, dout)
    assert rel_l2(dpre.float(), natf.grad[:, order]) < 4e-3


# ------------------------------------------------------------------------------------------------- reductions
def test_colsum_grouped(cuda):
    from lkgd_b200 import ops
    B_, Fr, HW, C = 2, 3, 50, 96
    M = B_ * Fr * HW
    G = rnd(M, C, dev=cuda, dtype=torch.float32, seed=1)
    m = torch.arange(M, device=cuda)
    for mode, n, idx in [(ops.RV_BATCH, B_, m // (HW * Fr)), (ops.RV_TCTX_0272, B_, ((m // (HW * Fr)) * HW + m % HW) % B_),
                         (ops.RV_FRAMEPOS, Fr, (m // HW) % Fr)]:
        out = ops.colsum_grouped(G, n, (mode, HW, Fr, B_))
        ref = torch.zeros(n, C, device=cuda).index_add_(0, idx, G)
        assert rel_l2(out, ref) < 1e-5, mode
    # one group (batch 1 per GPU): the 16-byte-load kernel; ragged row count, column count not a multiple of 128,
    # accumulation into a slice of a wider destination
    for M1, C1 in [(1234, 320), (77, 100), (35840, 1280)]:
        G1 = rnd(M1, C1, dev=cuda, dtype=torch.float32, seed=2)
        wide = torch.full((1, C1 + 64), 0.5, device=cuda)
        ops.colsum_grouped(G1, 1, (ops.RV_BATCH, M1, 1, 1), out=wide[:, 32:32 + C1])
        assert rel_l2(wide[:, 32:32 + C1] - 0.5, G1.double().sum(0, keepdim=True).float()) < 1e-5
        assert float((wide[:, :32] - 0.5).abs().max()) == 0 and float((wide[:, 32 + C1:] - 0.5).abs().max()) == 0


def test_downsum_and_zero_stuff(cuda):
    from lkgd_b200 import ops
    N, H, W, C = 3, 5, 7, 64
    x = rnd(N * 4 * H * W, C, dev=cuda, seed=1)
    out = ops.downsum2x(x, N, H, W)
    ref = x.float().reshape(N, H, 2, W, 2, C).sum(dim=(2, 4)).reshape(-1, C)
    assert rel_l2(out, ref) < 1e-6
    for Hin, Win in [(10, 14), (9, 13)]:
        Ho, Wo = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
        y = rnd(N * Ho * Wo, C, dev=cuda, dtype=torch.float32, seed=2)
        z = ops.zero_stuff2x(y, N, Hin, Win).float().reshape(N, Hin, Win, C)
        ref = torch.zeros(N, Hin, Win, C, device=cuda)
        ref[:, ::2, ::2] = y.to(bf16).float().reshape(N, Ho, Wo, C)
        assert torch.equal(z, ref)


@pytest.mark.parametrize("M,I,J", [(1000, 64, 64), (5000, 192, 320), (333, 320, 72), (70000, 128, 1280)])
def test_gemm_tn(cuda, M, I, J):
    from lkgd_b200 import ops
    wide = rnd(M, I + 16, dev=cuda, seed=1)
    X = wide[:, 8:8 + I]
    Y = rnd(M, J, dev=cuda, seed=2)
    out = torch.full((I, J), 1.0, device=cuda)
    ops.gemm_tn(X, Y, out, alpha=0.5)
    ref = 1.0 + 0.5 * (X.float().t() @ Y.float())
    assert rel_l2(out, ref) < 2e-5


# ------------------------------------------------------------------------------------------------- conv data gradients
def _conv_dgrad_weight(w):
    """Conv2d weight [Co, Ci, 3, 3] -> data-gradient GEMM weight [Ci, 9*Co] (taps flipped, channels swapped)."""
    co, ci = w.shape[:2]
    return w.flip(2, 3).permute(1, 2, 3, 0).reshape(ci, 9 * co).contiguous()


@pytest.mark.parametrize("stride", [1, 2])
def test_conv3x3_dgrad_via_gemm(cuda, stride):
    from lkgd_b200 import ops
    N, H, W, Ci, Co = 2, 10, 12, 64, 128
    x = rnd(N, Ci, H, W, dev=cuda, dtype=torch.float32, seed=1).requires_grad_(True)
    w = rnd(Co, Ci, 3, 3, dev=cuda, scale=(9 * Ci) ** -0.5, seed=2)
    y = F.conv2d(x, w.float(), stride=stride, padding=1)
    Ho, Wo = y.shape[-2:]
    dy = rnd(N, Co, Ho, Wo, dev=cuda, seed=3)
    y.backward(dy.float())
    dy_rows = dy.permute(0, 2, 3, 1).reshape(-1, Co).contiguous()
    if stride == 2:
        dy_rows = ops.zero_stuff2x(dy_rows, N, H, W)
    got = ops.gemm(dy_rows, _conv_dgrad_weight(w), mode=ops.A_CONV3X3, conv=(N, H, W, 1), out_f32=True)
    ref = x.grad.permute(0, 2, 3, 1).reshape(-1, Ci)
    assert rel_l2(got, ref) < 1e-5


def test_tconv3_dgrad_via_gemm(cuda):
    from lkgd_b200 import ops
    B_, Fr, HW, C = 2, 5, 70, 64
    x = rnd(B_, C, Fr, HW, 1, dev=cuda, dtype=torch.float32, seed=1).requires_grad_(True)
    w = rnd(C, C, 3, 1, 1, dev=cuda, scale=(3 * C) ** -0.5, seed=2)
    y = F.conv3d(x, w.float(), padding=(1, 0, 0))
    dy = rnd(B_, C, Fr, HW, 1, dev=cuda, seed=3)
    y.backward(dy.float())
    dy_rows = dy[..., 0].permute(0, 2, 3, 1).reshape(-1, C).contiguous()
    wd = w[..., 0, 0].flip(2).permute(1, 2, 0).reshape(C, 3 * C).contiguous()     # [Ci, kt', Co]
    got = ops.gemm(dy_rows, wd, mode=ops.A_TCONV3, tconv=(B_, Fr, HW), out_f32=True)
    ref = x.grad[..., 0].permute(0, 2, 3, 1).reshape(-1, C)
    assert rel_l2(got, ref) < 1e-5


# ------------------------------------------------------------------------------------------------- attention backward
@pytest.mark.parametrize("n_img,heads,d,N", [(2, 2, 64, 200), (1, 3, 16, 64), (3, 1, 32, 130), (1, 5, 64, 640),
                                             (2, 2, 128, 200), (1, 3, 128, 576),      # d = 128: reference-default heads
                                             # d = 64 runs on the tcgen05 / TMEM kernels: ragged tails of the 128-row
                                             # resident tile and of the 64-row streamed tile, one tile, the C5 level-0 size
                                             (3, 1, 64, 130), (1, 2, 64, 64), (2, 1, 64, 65), (1, 5, 64, 2560)])
def test_attention_bwd(cuda, n_img, heads, d, N):
    from lkgd_b200 import ops
    C = heads * d
    qkv = rnd(n_img * N, 3 * C, dev=cuda, seed=1)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o, lse = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N, return_lse=True)
    dO = rnd(n_img * N, C, dev=cuda, seed=2)

    def split(t):
        return t.float().reshape(n_img, N, heads, d).permute(0, 2, 1, 3)
    qf, kf, vf = (split(t).clone().requires_grad_(True) for t in (q, k, v))
    s = (qf @ kf.transpose(-1, -2)) * d ** -0.5
    ref_lse = torch.logsumexp(s, -1) / math.log(2.0)
    assert rel_l2(lse, ref_lse) < 1e-4
    of = torch.softmax(s, -1) @ vf
    assert rel_l2(split(o), of) < 6e-3
    of.backward(split(dO))
    dqkv = torch.zeros_like(qkv)
    ops.attention_bwd(q, k, v, o, dO, lse, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], n_img=n_img, heads=heads,
                      d=d, N=N)
    for name, got, ref in (("dq", dqkv[:, :C], qf.grad), ("dk", dqkv[:, C:2 * C], kf.grad), ("dv", dqkv[:, 2 * C:], vf.grad)):
        assert rel_l2(split(got), ref) < 1.2e-2, name


def test_attention_bwd_tcgen05_equals_the_mma_sync_kernels(cuda, monkeypatch):
    """64-wide heads: the tcgen05 / TMEM backward (attention_bwd_tc.cu) against the mma.sync pair it replaces
    (LKGD_ATTN_BWD_MMA=1) on the same inputs - two independent implementations of the same formulas."""
    from lkgd_b200 import ops
    n_img, heads, d, N = 2, 3, 64, 333
    C = heads * d
    qkv = rnd(n_img * N, 3 * C, dev=cuda, seed=5)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o, lse = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N, return_lse=True)
    dO = rnd(n_img * N, C, dev=cuda, seed=6)
    outs = []
    for legacy in (False, True):
        if legacy:
            monkeypatch.setenv("LKGD_ATTN_BWD_MMA", "1")
        else:
            monkeypatch.delenv("LKGD_ATTN_BWD_MMA", raising=False)
        g = torch.zeros_like(qkv)
        ops.attention_bwd(q, k, v, o, dO, lse, g[:, :C], g[:, C:2 * C], g[:, 2 * C:], n_img=n_img, heads=heads, d=d, N=N)
        outs.append(g)
    for i, name in enumerate(("dq", "dk", "dv")):
        assert rel_l2(outs[0][:, i * C:(i + 1) * C].float(), outs[1][:, i * C:(i + 1) * C].float()) < 6e-3, name


@pytest.mark.parametrize("B_,Fr,HW,heads,d", [(1, 14, 40, 2, 64), (2, 8, 33, 4, 16), (1, 25, 17, 1, 32), (1, 32, 8, 2, 64),
                                                  (1, 14, 33, 2, 128), (2, 25, 9, 1, 128)])
def test_attention_temporal_bwd(cuda, B_, Fr, HW, heads, d):
    from lkgd_b200 import ops
    C = heads * d
    qkv = rnd(B_ * Fr * HW, 3 * C, dev=cuda, seed=1)
    dO = rnd(B_ * Fr * HW, C, dev=cuda, seed=2)

    def split(t):          # [B, F, HW, heads, d] -> [B, HW, heads, F, d]
        return t.float().reshape(B_, Fr, HW, heads, d).permute(0, 2, 3, 1, 4)
    qf, kf, vf = (split(qkv[:, i * C:(i + 1) * C]).clone().requires_grad_(True) for i in range(3))
    of = torch.softmax((qf @ kf.transpose(-1, -2)) * d ** -0.5, -1) @ vf
    o = ops.attention_temporal(qkv, B=B_, F=Fr, HW=HW, heads=heads, d=d)
    assert rel_l2(split(o), of) < 6e-3
    of.backward(split(dO))
    dqkv = ops.attention_temporal_bwd(qkv, dO, B=B_, F=Fr, HW=HW, heads=heads, d=d)
    for i, (name, ref) in enumerate((("dq", qf.grad), ("dk", kf.grad), ("dv", vf.grad))):
        assert rel_l2(split(dqkv[:, i * C:(i + 1) * C]), ref) < 1.2e-2, name


# ------------------------------------------------------------------------------------------------- loss / optimizer
def test_edm_precondition_and_loss(cuda):
    from lkgd_b200 import ops
    B_, Fr, C, H, W = 2, 3, 4, 6, 10
    lat = rnd(B_, Fr, C, H, W, dev=cuda, dtype=torch.float32, seed=1)
    noise = rnd(B_, Fr, C, H, W, dev=cuda, dtype=torch.float32, seed=2)
    cond = rnd(B_, C, H, W, dev=cuda, dtype=torch.float32, seed=3)
    sigma = torch.tensor([0.7, 4.2], device=cuda)
    noisy, x_in = ops.edm_precondition(lat, noise, sigma, cond, 64)
    sg = sigma.view(B_, 1, 1, 1, 1)
    ref_noisy = lat + noise * sg
    assert rel_l2(noisy, ref_noisy) < 1e-6
    ref_in = torch.cat([ref_noisy / (sg ** 2 + 1) ** 0.5, cond.unsqueeze(1).repeat(1, Fr, 1, 1, 1)], 2)
    got_in = x_in.float().reshape(B_, Fr, H, W, 64)
    assert rel_l2(got_in[..., :8].permute(0, 1, 4, 2, 3), ref_in) < 4e-3
    assert float(got_in[..., 8:].abs().max()) == 0.0
    # loss + gradient
    pred_rows = rnd(B_ * Fr * H * W, 4, dev=cuda, dtype=torch.float32, seed=4).requires_grad_(True)
    pred = pred_rows.reshape(B_, Fr, H, W, C).permute(0, 1, 4, 2, 3)
    c_out, c_skip = -sg / (sg ** 2 + 1) ** 0.5, 1 / (sg ** 2 + 1)
    den = pred * c_out + c_skip * ref_noisy
    wgt = (1 + sg ** 2) * sg ** -2.0
    ref_loss = torch.mean((wgt * (den - lat) ** 2).reshape(B_, -1), dim=1).mean()
    ref_loss.backward()
    loss, dpred = ops.edm_loss(pred_rows.detach(), noisy, lat, sigma, 64)
    assert abs(float(loss) - float(ref_loss)) < 1e-5 * abs(float(ref_loss))
    assert rel_l2(dpred[:, :4].float(), pred_rows.grad) < 4e-3
    assert float(dpred[:, 4:].float().abs().max()) == 0.0


def test_adamw_and_clip(cuda):
    from lkgd_b200 import ops
    n = 10007
    p0 = rnd(n, dev=cuda, dtype=torch.float32, seed=1)
    g = rnd(n, dev=cuda, dtype=torch.float32, seed=2) * 3
    p_ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.AdamW([p_ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2)
    p, m, v = p0.clone(), torch.zeros(n, device=cuda), torch.zeros(n, device=cuda)
    ss = torch.empty((), device=cuda, dtype=torch.float64)
    for step in range(1, 4):
        p_ref.grad = (g * step * 0.5).clone()          # "all-reduced" gradient before averaging over 2 ranks
        p_ref.grad.mul_(0.5)
        torch.nn.utils.clip_grad_norm_([p_ref], 1.0)
        opt.step()
        gg = (g * step * 0.5).contiguous()
        ops.sumsq(gg, ss)
        assert abs(float(ss) - float((gg.double() ** 2).sum())) < 1e-6 * float(ss)
        ops.adamw(p, gg, m, v, lr=1e-3, step=step, grad_scale=0.5, sumsq_buf=ss, max_norm=1.0)
    assert rel_l2(p - p0, p_ref.detach() - p0) < 1e-4


def test_cast2d(cuda):
    from lkgd_b200 import ops
    src = rnd(40, 100, dev=cuda, dtype=torch.float32, seed=1)
    dst = torch.zeros(40, 64, device=cuda, dtype=bf16)
    ops.cast2d_bf16(src[:, 10:42], dst[:, 16:48], alpha=-2.0)
    assert torch.equal(dst[:, 16:48], (src[:, 10:42] * -2.0).to(bf16))
    assert float(dst[:, :16].float().abs().max()) == 0.0


@pytest.mark.parametrize("M,N,K,act", [(2, 96, 200, 3), (1, 1500, 1024, 0), (3, 777, 64, 3), (2, 512, 1000, 0)])
def test_small_linear_bwd(cuda, M, N, K, act):
    """dx / dW / db of the fp32 linear (latent-knowledge block, cross-attention vectors) against autograd; N >= 512 with
    K % 8 == 0 takes the wide kernel (one CTA per 8 output columns), everything else the column-per-thread kernel."""
    from lkgd_b200 import ops
    x = rnd(M, K, dev=cuda, dtype=torch.float32, seed=1).requires_grad_(True)
    W = (rnd(N, K, dev=cuda, dtype=torch.float32, seed=2) * 0.1).requires_grad_(True)
    b = rnd(N, dev=cuda, dtype=torch.float32, seed=3).requires_grad_(True)
    dy = rnd(M, N, dev=cuda, dtype=torch.float32, seed=4)
    y = F.linear(x, W, b)
    if act == 3:
        y = F.leaky_relu(y, 0.1)
    y.backward(dy)
    dW = torch.full((N, K), 0.5, device=cuda)
    db = torch.full((N,), -1.0, device=cuda)
    dx = ops.small_linear_bwd(dy, W.detach(), x=x.detach(), y=y.detach() if act else None, act_out=act, dW=dW, db=db)
    assert rel_l2(dx, x.grad) < 1e-5
    assert rel_l2(dW - 0.5, W.grad) < 1e-5
    assert rel_l2(db + 1.0, b.grad) < 1e-5
    acc = torch.full((M, K + 8), 2.0, device=cuda)[:, :K]        # accumulate into a strided destination
    ops.small_linear_bwd(dy, W.detach(), y=y.detach() if act else None, act_out=act, dx=acc)
    assert rel_l2(acc - 2.0, x.grad) < 1e-5
    again = ops.small_linear_bwd(dy, W.detach(), y=y.detach() if act else None, act_out=act)
    assert torch.equal(again, dx)                                 # fixed reduction order


def test_cast2d_batch_equals_the_single_casts(cuda):
    """One launch over a table of jobs (the LoRA repack) writes what the per-job casts write, bit for bit: plain,
    transposed-view and scaled sources, destinations that are slices of wider operands."""
    from lkgd_b200 import ops
    pA = rnd(64, 320, dev=cuda, dtype=torch.float32, seed=1)
    pB = rnd(960, 64, dev=cuda, dtype=torch.float32, seed=2)
    small = rnd(3, 5, dev=cuda, dtype=torch.float32, seed=3)

    def dsts():
        return (torch.zeros(128, 320, device=cuda, dtype=bf16), torch.zeros(320, 128, device=cuda, dtype=bf16),
                torch.zeros(1920, 128, device=cuda, dtype=bf16), torch.zeros(128, 1920, device=cuda, dtype=bf16),
                torch.zeros(3, 8, device=cuda, dtype=bf16))

    def jobs(d):
        return [(pA, d[0][64:128], 1.0), (pA.t(), d[1][:, 64:128], 1.0), (pB, d[2][960:, :64], 0.25),
                (pB.t(), d[3][:64, 960:], 0.25), (small, d[4][:, :5], -2.0)]
    ref, got = dsts(), dsts()
    for src, dst, alpha in jobs(ref):
        ops.cast2d_bf16(src, dst, alpha=alpha)
    batch = ops.Cast2dBatch(jobs(got))
    batch.run()
    for a, b in zip(ref, got):
        assert torch.equal(a, b)
    assert float(got[0][64:128].float().abs().sum()) > 0 and float(got[0][:64].float().abs().sum()) == 0
