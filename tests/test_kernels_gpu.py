"""Kernel parity on the GPU (through the C ABI): each CUDA kernel against a plain PyTorch fp32 evaluation of
the same op on bf16-rounded inputs, and against the SIMT checker kernels at sizes the CPU cannot do."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


def rnd(*shape, dev, scale=1.0, dtype=bf16, seed=None):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed if seed is not None else (hash(shape) & 0xFFFF))
    return (torch.randn(*shape, generator=g) * scale).to(dtype).to(dev)


# ------------------------------------------------------------------------------------------------- GEMM
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 320, 320), (128, 32, 32), (4096, 960, 320),
                                   (300, 1280, 1280), (130, 48, 72)])
def test_gemm_linear(cuda, M, N, K):
    from lkgd_b200 import ops
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    b = rnd(N, dev=cuda, dtype=torch.float32)
    out = ops.gemm(A, W, bias=b)
    ref = A.float() @ W.float().t() + b
    assert rel_l2(out.float(), ref) < 4e-3
    chk = ops.gemm(A, W, bias=b, checker=True)
    assert rel_l2(out.float(), chk.float()) < 3e-3


@pytest.mark.parametrize("M,N,K", [(560, 1280, 1280), (2240, 1280, 640), (560, 640, 2560), (256, 1280, 320)])
def test_gemm_narrow_tiles_of_few_row_problems(cuda, monkeypatch, M, N, K):
    """Few row tiles (training clip at the lower UNet levels): the GEMM picks a narrower N tile that still fits one
    wave (refine_bn: BN = 160 / 128 / 64 instead of 256).  Every epilogue family on those tiles equals the widest-tile
    result bit for bit (the contraction order per output element does not depend on the tile width) and the SIMT
    checker within bf16 rounding; the fused GroupNorm statistics equal the sums of the stored tensor."""
    from lkgd_b200 import ops
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    b = rnd(N, dev=cuda, dtype=torch.float32)
    res = rnd(M, N, dev=cuda, dtype=torch.float32, seed=3)
    res_b = rnd(M, N, dev=cuda, seed=4)
    rows = 128 if M % 128 == 0 else 0

    def run():
        outs = [ops.gemm(A, W, bias=b), ops.gemm(A, W, bias=b, act=ops.ACT_SILU, res1=res_b, s1=0.5),
                ops.gemm(A, W, bias=b, res1=res, s0=0.7, out_f32=True, gn_rows=rows),
                ops.gemm(A, W, bias=b, out_f32=True, want_bf16=True)[1]]
        st = ops.gn_stats_of(outs[2])
        return outs, (st[0].clone() if st is not None else None)
    narrow, st_n = run()
    monkeypatch.setenv("LKGD_GEMM_NO_REFINE", "1")
    wide, st_w = run()
    monkeypatch.delenv("LKGD_GEMM_NO_REFINE")
    for a, w in zip(narrow, wide):
        assert torch.equal(a, w)
    chk = ops.gemm(A, W, bias=b, res1=res, s0=0.7, out_f32=True, checker=True)
    assert rel_l2(narrow[2], chk) < 1e-5
    ref = A.float() @ W.float().t() + b
    assert rel_l2(narrow[0].float(), ref) < 4e-3
    if rows:
        x = narrow[2].double().view(M // rows, rows, N)
        assert rel_l2(st_n[..., 0], x.sum(1)) < 1e-5 and rel_l2(st_n[..., 1], (x * x).sum(1)) < 1e-5
        assert rel_l2(st_n, st_w) < 1e-9


def test_gemm_column_slice_and_f32_out(cuda):
    from lkgd_b200 import ops
    M, K, N = 512, 64, 192
    wide = rnd(M, 3 * K, dev=cuda)
    A = wide[:, K:2 * K]
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    out = ops.gemm(A, W, out_f32=True)
    ref = A.float() @ W.float().t()
    assert out.dtype == torch.float32
    assert rel_l2(out, ref) < 1e-5


def test_gemm_epilogue_terms(cuda):
    from lkgd_b200 import ops
    B_, Fr, HW, N, K = 2, 3, 64, 320, 128
    M = B_ * Fr * HW
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    b = rnd(N, dev=cuda, dtype=torch.float32)
    r1, r2 = rnd(M, N, dev=cuda, seed=1), rnd(M, N, dev=cuda, seed=2)
    base = A.float() @ W.float().t() + b
    m = torch.arange(M, device=cuda)
    for mode, G, idx in [(ops.RV_FRAME, B_ * Fr, m // HW), (ops.RV_FRAMEPOS, Fr, (m // HW) % Fr),
                         (ops.RV_BATCH, B_, m // (HW * Fr)),
                         (ops.RV_TCTX_0272, B_, ((m // (HW * Fr)) * HW + m % HW) % B_)]:
        rv = rnd(G, N, dev=cuda, dtype=torch.float32, seed=mode)
        out = ops.gemm(A, W, bias=b, rowvec=rv, rv=(mode, HW, Fr, B_), act=ops.ACT_SILU, s0=0.3, res1=r1, s1=0.7,
                       res2=r2, s2=-1.5)
        ref = 0.3 * F.silu(base + rv[idx]) + 0.7 * r1.float() - 1.5 * r2.float()
        assert rel_l2(out.float(), ref) < 4e-3, mode


def test_gemm_n_store_conv_out_shape(cuda):
    from lkgd_b200 import ops
    M, K = 700, 64
    A = rnd(M, K, dev=cuda)
    W = torch.zeros(32, K, device=cuda, dtype=bf16)
    W[:4] = rnd(4, K, dev=cuda, scale=K ** -0.5)
    b = torch.zeros(32, device=cuda)
    b[:4] = 0.5
    out = ops.gemm(A, W, bias=b, out_f32=True, n_store=4)
    assert out.shape == (M, 4)
    assert rel_l2(out, A.float() @ W[:4].float().t() + 0.5) < 1e-5


@pytest.mark.parametrize("C", [32, 320])
def test_gemm_geglu(cuda, C):
    from lkgd_b200 import ops
    M = 384
    A = rnd(M, C, dev=cuda)
    W = rnd(8 * C, C, dev=cuda, scale=C ** -0.5)
    b = rnd(8 * C, dev=cuda, dtype=torch.float32)
    Wp, bp = ops.pack_geglu(W, b)
    out = ops.gemm(A, Wp, bias=bp, act=ops.ACT_GEGLU)
    proj = A.float() @ W.float().t() + b
    h, g = proj.chunk(2, dim=-1)
    ref = h * F.gelu(g)
    assert out.shape == (M, 4 * C)
    assert rel_l2(out.float(), ref) < 4e-3
    chk = ops.gemm(A, Wp, bias=bp, act=ops.ACT_GEGLU, checker=True)
    assert rel_l2(out.float(), chk.float()) < 3e-3


@pytest.mark.parametrize("r", [8, 64])
def test_gemm_lora_second_segment(cuda, r):
    from lkgd_b200 import ops
    M, K, N = 640, 320, 320
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    T = rnd(M, r, dev=cuda)
    Bl = rnd(N, r, dev=cuda, scale=0.1)
    out = ops.gemm(A, W, A1=T, Bw1=Bl)
    ref = A.float() @ W.float().t() + T.float() @ Bl.float().t()
    assert rel_l2(out.float(), ref) < 4e-3


@pytest.mark.parametrize("NIMG,H,W,Cin,Cout,stride", [(3, 18, 32, 64, 128, 1), (2, 16, 16, 32, 64, 1),
                                                      (2, 9, 16, 128, 64, 1), (2, 36, 64, 64, 64, 2),
                                                      (3, 10, 16, 32, 32, 2), (1, 72, 128, 64, 320, 1),
                                                      (2, 5, 8, 64, 64, 1)])
def test_gemm_conv3x3(cuda, NIMG, H, W, Cin, Cout, stride):
    from lkgd_b200 import ops
    x = rnd(NIMG, Cin, H, W, dev=cuda)                      # NCHW reference layout
    w = rnd(Cout, Cin, 3, 3, dev=cuda, scale=(9 * Cin) ** -0.5)
    b = rnd(Cout, dev=cuda, dtype=torch.float32)
    ref = F.conv2d(x.float(), w.float(), b, stride=stride, padding=1)        # [N, Cout, Ho, Wo]
    x_nhwc = x.permute(0, 2, 3, 1).contiguous()
    w_k = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
    out = ops.gemm(x_nhwc, w_k, mode=ops.A_CONV3X3, conv=(NIMG, H, W, stride), bias=b)
    Ho, Wo = ref.shape[2:]
    got = out.float().view(NIMG, Ho, Wo, Cout).permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < 4e-3


@pytest.mark.parametrize("NIMG,H,W,Cin,C1,Cout", [(3, 16, 24, 64, 128, 96), (7, 9, 16, 128, 192, 64), (2, 18, 32, 64, 72, 160)])
def test_gemm_conv3x3_with_fused_1x1_shortcut(cuda, NIMG, H, W, Cin, C1, Cout):
    """conv2 of a resblock with its 1x1 shortcut conv as the second K segment (centre tap over the raw block input):
    conv3x3(h) + conv1x1(x) in one launch (diffusers ResnetBlock2D.conv2 + conv_shortcut)."""
    from lkgd_b200 import ops
    h = rnd(NIMG, Cin, H, W, dev=cuda)
    x = rnd(NIMG, C1, H, W, dev=cuda, seed=11)
    w = rnd(Cout, Cin, 3, 3, dev=cuda, scale=(9 * Cin) ** -0.5)
    ws = rnd(Cout, C1, 1, 1, dev=cuda, scale=C1 ** -0.5, seed=12)
    b = rnd(Cout, dev=cuda, dtype=torch.float32)
    ref = F.conv2d(h.float(), w.float(), b, padding=1) + F.conv2d(x.float(), ws.float())
    out = ops.gemm(h.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous(),
                   mode=ops.A_CONV3X3, conv=(NIMG, H, W, 1), bias=b, out_f32=True,
                   A1=x.permute(0, 2, 3, 1).reshape(NIMG * H * W, C1).contiguous(), Bw1=ws.reshape(Cout, C1).contiguous())
    got = out.view(NIMG, H, W, Cout).permute(0, 3, 1, 2)
    assert rel_l2(got, ref) < 4e-3


@pytest.mark.parametrize("B_,Fr,HW,C,N", [(2, 5, 200, 64, 64), (1, 8, 1024, 32, 32), (2, 3, 144, 128, 128),
                                          (1, 1, 64, 64, 64)])
def test_gemm_tconv3(cuda, B_, Fr, HW, C, N):
    from lkgd_b200 import ops
    x = rnd(B_, C, Fr, HW, 1, dev=cuda)                     # [B,C,F,H,W] with H*W flattened
    w = rnd(N, C, 3, 1, 1, dev=cuda, scale=(3 * C) ** -0.5)
    b = rnd(N, dev=cuda, dtype=torch.float32)
    ref = F.conv3d(x.float(), w.float(), b, padding=(1, 0, 0))              # [B,N,F,HW,1]
    x_cl = x[..., 0].permute(0, 2, 3, 1).contiguous()                        # [B,F,HW,C]
    w_k = w[..., 0, 0].permute(0, 2, 1).reshape(N, 3 * C).contiguous()       # k = kt*C + c
    out = ops.gemm(x_cl, w_k, mode=ops.A_TCONV3, tconv=(B_, Fr, HW), bias=b)
    got = out.float().view(B_, Fr, HW, N).permute(0, 3, 1, 2)
    assert rel_l2(got, ref[..., 0]) < 4e-3


@pytest.mark.parametrize("B_,Fr", [(1, 4), (2, 3)])
def test_gemm_tconv3_big_frames_use_the_frame_fastest_tile_order(cuda, B_, Fr):
    """One frame of A above 16 MB (the VAE decoder's upper levels): tiles are ordered (pixel tile, frame); same result as the
    SIMT checker, fused GroupNorm statistics included."""
    from lkgd_b200 import ops
    HW, C, N = 131072 + 200, 64, 32                       # 16.8 MB per frame, ragged last pixel tile
    x = rnd(B_ * Fr * HW, C, dev=cuda)
    w = rnd(N, 3 * C, dev=cuda, scale=(3 * C) ** -0.5)
    b = rnd(N, dev=cuda, dtype=torch.float32)
    r = rnd(B_ * Fr * HW, N, dev=cuda, dtype=torch.float32)
    out = ops.gemm(x, w, mode=ops.A_TCONV3, tconv=(B_, Fr, HW), bias=b, res1=r, out_f32=True, gn_rows=HW)
    chk = ops.gemm(x, w, mode=ops.A_TCONV3, tconv=(B_, Fr, HW), bias=b, res1=r, out_f32=True, checker=True)
    assert rel_l2(out, chk) < 1e-5
    st, rows = ops.gn_stats_of(out)
    assert rows == HW and tuple(st.shape) == (B_ * Fr, N, 2)
    ref = torch.stack([chk.view(B_ * Fr, HW, N).double().sum(1), (chk.view(B_ * Fr, HW, N).double() ** 2).sum(1)], -1)
    assert torch.allclose(st, ref, rtol=1e-6, atol=1e-3)


def test_gemm_large_vs_checker(cuda):
    """SVD level-0 projection shape (M = 2*25*72*128 would be 460800; one CFG half of 14 frames here)."""
    from lkgd_b200 import ops
    M, N, K = 14 * 9216, 320, 320
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    out = ops.gemm(A, W)
    chk = ops.gemm(A, W, checker=True)
    assert rel_l2(out.float(), chk.float()) < 3e-3
    assert torch.isfinite(out.float()).all()


# ------------------------------------------------------------------------------------------------- norms
@pytest.mark.parametrize("NS,R,C1,C2,silu", [(4, 1024, 32, 0, True), (3, 576, 320, 0, True), (2, 2304, 640, 320, True),
                                             (2, 8 * 144, 1280, 0, True), (5, 200, 64, 0, False),
                                             (2, 144, 1280, 1280, True)])
def test_groupnorm(cuda, NS, R, C1, C2, silu):
    from lkgd_b200 import ops
    x1 = rnd(NS * R, C1, dev=cuda) + 0.5
    x2 = rnd(NS * R, C2, dev=cuda, scale=2.0) if C2 else None
    Ct = C1 + C2
    g = rnd(Ct, dev=cuda, dtype=torch.float32) * 0.2 + 1.0
    b = rnd(Ct, dev=cuda, dtype=torch.float32, seed=3) * 0.2
    out = ops.groupnorm(x1, g, b, 1e-5, NS=NS, R=R, x2=x2, silu=silu)
    x = x1 if x2 is None else torch.cat([x1, x2], dim=1)
    xr = x.float().view(NS, R, Ct).permute(0, 2, 1)                       # [NS, C, R]
    ref = F.group_norm(xr, 32, g, b, 1e-5)
    if silu:
        ref = F.silu(ref)
    ref = ref.permute(0, 2, 1).reshape(NS * R, Ct)
    assert rel_l2(out.float(), ref) < 4e-3


@pytest.mark.parametrize("M,C", [(1000, 32), (777, 320), (512, 640), (300, 1280)])
def test_layernorm(cuda, M, C):
    from lkgd_b200 import ops
    x = rnd(M, C, dev=cuda) * 2 + 0.3
    g = rnd(C, dev=cuda, dtype=torch.float32) * 0.2 + 1.0
    b = rnd(C, dev=cuda, dtype=torch.float32, seed=5) * 0.2
    out = ops.layernorm(x, g, b, 1e-5)
    ref = F.layer_norm(x.float(), (C,), g, b, 1e-5)
    assert rel_l2(out.float(), ref) < 4e-3


def test_layernorm_fused_add(cuda):
    from lkgd_b200 import ops
    B_, Fr, HW, C = 2, 4, 50, 320
    M = B_ * Fr * HW
    x = rnd(M, C, dev=cuda)
    g = torch.ones(C, device=cuda)
    b = torch.zeros(C, device=cuda)
    emb = rnd(Fr, C, dev=cuda, dtype=torch.float32)
    s = torch.empty_like(x)
    out = ops.layernorm(x, g, b, 1e-5, addvec=emb, rv=(ops.RV_FRAMEPOS, HW, Fr, B_), sum_out=s)
    f_idx = (torch.arange(M, device=cuda) // HW) % Fr
    s_ref = (x.float() + emb[f_idx]).to(bf16)
    assert torch.equal(s, s_ref)
    assert rel_l2(out.float(), F.layer_norm(s_ref.float(), (C,), g, b, 1e-5)) < 4e-3


# ------------------------------------------------------------------------------------------------- attention
def _attn_ref(q, k, v, n_img, heads, d, Nq, Nk):
    qf = q.float().view(n_img, Nq, heads, d).transpose(1, 2)
    kf = k.float().view(n_img, Nk, heads, d).transpose(1, 2)
    vf = v.float().view(n_img, Nk, heads, d).transpose(1, 2)
    w = torch.softmax(qf @ kf.transpose(-1, -2) * d ** -0.5, dim=-1)
    return (w @ vf).transpose(1, 2).reshape(n_img * Nq, heads * d)


@pytest.mark.parametrize("n_img,heads,d,N", [(2, 2, 64, 128), (2, 5, 64, 576), (3, 2, 16, 1024), (1, 4, 32, 144),
                                             (2, 10, 64, 2304), (1, 1, 64, 100),
                                             # d = 128: the reference-default heads (5,10,10,20) at level 2
                                             (2, 10, 128, 576), (1, 2, 128, 100), (3, 1, 128, 1090)])
def test_attention_self(cuda, n_img, heads, d, N):
    from lkgd_b200 import ops
    Cn = heads * d
    qkv = rnd(n_img * N, 3 * Cn, dev=cuda)
    q, k, v = qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:]
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N)
    ref = _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), n_img, heads, d, N, N)
    assert rel_l2(out.float(), ref) < 6e-3


def test_attention_cross_kv_len(cuda):
    from lkgd_b200 import ops
    n_img, heads, d, Nq, Nk = 2, 3, 64, 300, 77
    q = rnd(n_img * Nq, heads * d, dev=cuda)
    k = rnd(n_img * Nk, heads * d, dev=cuda, seed=1)
    v = rnd(n_img * Nk, heads * d, dev=cuda, seed=2)
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nk)
    assert rel_l2(out.float(), _attn_ref(q, k, v, n_img, heads, d, Nq, Nk)) < 6e-3
    chk = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nk, checker=True)
    assert rel_l2(out.float(), chk.float()) < 6e-3


@pytest.mark.parametrize("n_img,heads,d,Nq,Nk", [(2, 5, 64, 576, 576), (1, 2, 64, 200, 77), (2, 3, 32, 300, 129),
                                                 (1, 2, 16, 130, 64), (1, 5, 64, 2304, 2304)])
def test_attention_q_in_tmem_variant(cuda, monkeypatch, n_img, heads, d, Nq, Nk):
    """LKGD_ATTN_QT=1: the variant that keeps Q in tensor memory (TS-form Q K^T, two CTAs per SM) - same results as the
    default kernel (same arithmetic, same order) and as the fp32 reference."""
    from lkgd_b200 import ops
    q = rnd(n_img * Nq, heads * d, dev=cuda)
    k = rnd(n_img * Nk, heads * d, dev=cuda, seed=1)
    v = rnd(n_img * Nk, heads * d, dev=cuda, seed=2)
    base = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nk)
    monkeypatch.setenv("LKGD_ATTN_QT", "1")
    qt = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nk)
    monkeypatch.delenv("LKGD_ATTN_QT")
    assert rel_l2(qt.float(), _attn_ref(q, k, v, n_img, heads, d, Nq, Nk)) < 6e-3
    assert rel_l2(qt.float(), base.float()) < 1e-3


@pytest.mark.parametrize("d", [64, 128])
@pytest.mark.parametrize("Nk", [1, 5, 63, 64, 65, 129])
def test_attention_short_and_ragged_kv(cuda, Nk, d):
    """KV lengths around the 64-key step of the softmax pipeline (single partial step, exactly one step, one key over)."""
    from lkgd_b200 import ops
    n_img, heads, Nq = 2, 2, 200
    q = rnd(n_img * Nq, heads * d, dev=cuda)
    k = rnd(n_img * Nk, heads * d, dev=cuda, seed=1)
    v = rnd(n_img * Nk, heads * d, dev=cuda, seed=2)
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nk)
    assert rel_l2(out.float(), _attn_ref(q, k, v, n_img, heads, d, Nq, Nk)) < 6e-3


def test_attention_growing_row_maximum_rescales(cuda):
    """Keys whose scores grow block after block (row maxima rise by far more than the lazy-rescale threshold of 2^8
    several times, for some rows only): exercises the deferred O rescale in TMEM and the running sum correction."""
    from lkgd_b200 import ops
    n_img, heads, d, N = 1, 2, 64, 640
    g = torch.Generator().manual_seed(3)
    q = torch.randn(n_img * N, heads * d, generator=g)
    k = torch.randn(n_img * N, heads * d, generator=g)
    ramp = (torch.arange(N) // 64).float()                     # 10 steps of 64 keys
    k = k + q.mean(0, keepdim=True) * 0.0                        # keep k random ...
    k = k * (1.0 + 0.9 * ramp)[:, None]                          # ... but ever larger: later blocks dominate
    q[: N // 2] *= 3.0                                           # half of the rows see steeper growth than the others
    v = torch.randn(n_img * N, heads * d, generator=g)
    q, k, v = (t.to(torch.bfloat16).to(cuda) for t in (q, k, v))
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N)
    ref = _attn_ref(q, k, v, n_img, heads, d, N, N)
    assert torch.isfinite(out.float()).all()
    assert rel_l2(out.float(), ref) < 8e-3


def test_attention_svd_l0_vs_checker(cuda):
    from lkgd_b200 import ops
    n_img, heads, d, N = 1, 5, 64, 9216
    Cn = heads * d
    qkv = rnd(n_img * N, 3 * Cn, dev=cuda)
    q, k, v = qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:]
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N)
    chk = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N, checker=True)
    assert rel_l2(out.float(), chk.float()) < 6e-3


@pytest.mark.parametrize("B_,Fr,HW,heads,d", [(2, 8, 64, 2, 16), (2, 25, 144, 5, 64), (1, 14, 100, 4, 32),
                                              (1, 1, 40, 2, 64), (2, 25, 72, 10, 128), (1, 14, 33, 1, 128)])
def test_attention_temporal(cuda, B_, Fr, HW, heads, d):
    from lkgd_b200 import ops
    Cn = heads * d
    qkv = rnd(B_ * Fr * HW, 3 * Cn, dev=cuda)
    out = ops.attention_temporal(qkv, B=B_, F=Fr, HW=HW, heads=heads, d=d)
    t = qkv.float().view(B_, Fr, HW, 3, heads, d).permute(3, 0, 2, 4, 1, 5)     # [3,B,HW,h,F,d]
    w = torch.softmax(t[0] @ t[1].transpose(-1, -2) * d ** -0.5, dim=-1)
    ref = (w @ t[2]).permute(0, 3, 1, 2, 4).reshape(B_ * Fr * HW, Cn)            # [B,F,HW,h,d]
    assert rel_l2(out.float(), ref) < 4e-3


# ------------------------------------------------------------------------------------------------- glue
def test_small_linear_and_timestep_embedding(cuda):
    from lkgd_b200 import ops
    x = rnd(3, 1280, dev=cuda, dtype=torch.float32)
    W = rnd(320, 1280, dev=cuda, dtype=torch.float32, scale=0.03)
    b = rnd(320, dev=cuda, dtype=torch.float32)
    y = ops.small_linear(x, W, b, act_in=1, act_out=3)
    ref = F.leaky_relu(F.silu(x) @ W.t() + b, 0.1)
    assert rel_l2(y, ref) < 1e-5
    t = torch.tensor([1.6377, -0.92, 6.0, 127.0, 0.02], device=cuda)
    e = ops.timestep_embedding(t, 320)
    k = torch.arange(160, device=cuda, dtype=torch.float32)
    arg = t[:, None] * torch.exp(-math.log(10000.0) * k / 160)[None]
    assert rel_l2(e, torch.cat([arg.cos(), arg.sin()], -1)) < 1e-5


@pytest.mark.parametrize("mode", ["conv", "tconv", "linear", "conv_s2", "conv_9x16"])
def test_gemm_fused_groupnorm_statistics(cuda, mode):
    """lkgd_gemm(gn_stats) + lkgd_groupnorm_from_stats == lkgd_gemm + lkgd_groupnorm (spatial: statistics per frame;
    temporal: across frames; two-source concatenation), and the raw sums match torch."""
    from lkgd_b200 import ops
    from lkgd_b200.ops import A_CONV3X3, A_LINEAR, A_TCONV3
    B, Fr, H, W, Ci, Co = 2, 3, 16, 24, 64, 96
    if mode == "conv_9x16":      # small feature map: four images share a 128-row tile, one frame per epilogue warp
        B, Fr, H, W, mode = 1, 7, 9, 16, "conv"
    HW = H * W
    kw, taps, gn_rows = {}, 1, HW
    if mode == "conv":
        kw, taps = dict(mode=A_CONV3X3, conv=(B * Fr, H, W, 1)), 9
    elif mode == "conv_s2":
        kw, taps, gn_rows = dict(mode=A_CONV3X3, conv=(B * Fr, H, W, 2)), 9, (H // 2) * (W // 2)
    elif mode == "tconv":
        kw, taps = dict(mode=A_TCONV3, tconv=(B, Fr, HW)), 3
    M_in = B * Fr * HW
    M = B * Fr * gn_rows
    A = rnd(M_in, Ci, dev=cuda)
    Wt = rnd(Co, taps * Ci, dev=cuda, scale=(taps * Ci) ** -0.5)
    bias = rnd(Co, dev=cuda, dtype=torch.float32)
    res = rnd(M, Co, dev=cuda, dtype=torch.float32, seed=3)
    plain = ops.gemm(A, Wt, bias=bias, res1=res, s0=0.7, out_f32=True, **kw)
    fused = ops.gemm(A, Wt, bias=bias, res1=res, s0=0.7, out_f32=True, gn_rows=gn_rows, **kw)
    assert torch.equal(plain, fused)
    st, rows = ops.gn_stats_of(fused)
    assert rows == gn_rows and ops.gn_stats_of(plain) is None
    x = fused.double().view(B * Fr, gn_rows, Co)
    assert rel_l2(st[..., 0], x.sum(1)) < 1e-5 and rel_l2(st[..., 1], (x * x).sum(1)) < 1e-5
    g, b = rnd(Co, dev=cuda, dtype=torch.float32, seed=5), rnd(Co, dev=cuda, dtype=torch.float32, seed=6)
    for NS, R in ((B * Fr, gn_rows), (B, Fr * gn_rows)):
        a = ops.groupnorm(plain, g, b, 1e-5, NS=NS, R=R, silu=True)
        f = ops.groupnorm(fused, g, b, 1e-5, NS=NS, R=R, silu=True)
        assert rel_l2(f, a) < 2e-3       # bf16 outputs; statistics agree to ~1e-6
    g2, b2 = rnd(2 * Co, dev=cuda, dtype=torch.float32, seed=7), rnd(2 * Co, dev=cuda, dtype=torch.float32, seed=8)
    other = ops.gemm(A, Wt, bias=bias, out_f32=True, gn_rows=gn_rows, **kw)
    a = ops.groupnorm(plain, g2, b2, 1e-5, NS=B * Fr, R=gn_rows, x2=other.clone(), silu=False)
    f, raw = ops.groupnorm(fused, g2, b2, 1e-5, NS=B * Fr, R=gn_rows, x2=other, silu=False, want_raw=True)
    assert rel_l2(f, a) < 2e-3
    assert torch.equal(raw, torch.cat([fused, other], 1).to(bf16))       # the shortcut conv's operand, same pass
    assert ops.groupnorm(plain, g2, b2, 1e-5, NS=B * Fr, R=gn_rows, x2=other.clone(), silu=False, want_raw=True)[1] is None
    ops.axpby(res, 0.5, fused, 1.0)      # an in-place update drops the (now stale) statistics
    assert ops.gn_stats_of(fused) is None


def test_small_linear_result_does_not_depend_on_alignment(cuda):
    """Parameters trained in the flat fp32 buffer are views at 4-byte granularity; the same numbers must give the same
    bits wherever they live (a 1-ulp difference in the latent-knowledge context flips bf16 roundings downstream)."""
    from lkgd_b200 import ops
    x = rnd(2, 512, dev=cuda, dtype=torch.float32)
    W = rnd(129, 512, dev=cuda, dtype=torch.float32, scale=0.05)
    b = rnd(129, dev=cuda, dtype=torch.float32)
    ref = ops.small_linear(x, W, b)
    for off in (1, 2, 3):
        buf = torch.empty(W.numel() + 4, device=cuda, dtype=torch.float32)
        Wm = buf[off:off + W.numel()].view_as(W)
        Wm.copy_(W)
        xb = torch.empty(x.numel() + 4, device=cuda, dtype=torch.float32)
        xm = xb[off:off + x.numel()].view_as(x)
        xm.copy_(x)
        assert torch.equal(ops.small_linear(x, Wm, b), ref)
        assert torch.equal(ops.small_linear(xm, W, b), ref)


def test_pack_unpack_layouts(cuda):
    from lkgd_b200 import ops
    S, Fr, H, W = 2, 3, 8, 16
    lat = rnd(S, Fr, 4, H, W, dev=cuda, dtype=torch.float32)
    img = rnd(2 * S, Fr, 4, H, W, dev=cuda, dtype=torch.float32, seed=9)
    out = ops.pack_input(lat, 0.25, img, N=2 * S, Cpad=64)
    ref = torch.zeros(2 * S, Fr, H, W, 64, device=cuda)
    ref[..., :4] = (torch.cat([lat, lat]) * 0.25).permute(0, 1, 3, 4, 2)
    ref[..., 4:8] = img.permute(0, 1, 3, 4, 2)
    assert torch.equal(out.view(2 * S, Fr, H, W, 64), ref.to(bf16))
    src = rnd(S * Fr * H * W, 8, dev=cuda, dtype=torch.float32)
    un = ops.unpack_output(src[:, :4], S, Fr, 4, H, W)
    assert torch.equal(un, src[:, :4].reshape(S, Fr, H, W, 4).permute(0, 1, 4, 2, 3))
    x = rnd(5, 48, 9, 7, dev=cuda, dtype=torch.float32)
    cl = ops.nchw_to_nhwc(x)
    assert torch.equal(cl.view(5, 9, 7, 48), x.permute(0, 2, 3, 1).to(bf16))
    assert torch.equal(ops.nhwc_to_nchw(cl, 5, 9, 7), x.to(bf16).float())


def test_upsample_concat_axpby(cuda):
    from lkgd_b200 import ops
    N, H, W, Cn = 3, 5, 6, 64
    x = rnd(N * H * W, Cn, dev=cuda)
    up = ops.upsample2x(x, N, H, W)
    ref = F.interpolate(x.float().view(N, H, W, Cn).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.view(N, 2 * H, 2 * W, Cn), ref.permute(0, 2, 3, 1).to(bf16))
    y = rnd(N * H * W, 32, dev=cuda, seed=4)
    assert torch.equal(ops.concat_channels(x, y), torch.cat([x, y], 1))
    acc = x.clone()
    ops.axpby(rnd(N * H * W, Cn, dev=cuda, seed=6), 3.0, acc, 1.0)
    assert rel_l2(acc.float(), x.float() + 3.0 * rnd(N * H * W, Cn, dev=cuda, seed=6).float()) < 4e-3


def test_cfg_euler_step_matches_formula(cuda):
    """KAT from SURVEY Appendix C (i=3: sigma 322.4537 -> 244.0231, x=1.5, v=-0.25 -> 1.1959649)."""
    from lkgd_b200 import ops
    S, Fr, Cn, H, W = 1, 2, 4, 4, 4
    x = torch.full((S, Fr, Cn, H, W), 1.5, device=cuda)
    pred = torch.zeros(2 * S * Fr * H * W, 4, device=cuda)
    pred[: S * Fr * H * W] = -0.25
    pred[S * Fr * H * W:] = -0.25
    g = torch.tensor([1.0, 3.0], device=cuda)
    xn, v = ops.cfg_euler_step(pred, g, x, 322.45367431640625, 244.0230712890625, cfg=True, want_v=True)
    assert torch.allclose(v, torch.full_like(v, -0.25))
    assert abs(float(xn.flatten()[0]) - 1.1959649324417114) < 2e-6
    # general case vs the reference formula in fp32
    x = rnd(2, 3, 4, 8, 8, dev=cuda, dtype=torch.float32) * 700
    u = rnd(2, 3, 8, 8, 4, dev=cuda, dtype=torch.float32, seed=1)
    c = rnd(2, 3, 8, 8, 4, dev=cuda, dtype=torch.float32, seed=2)
    g = torch.linspace(1, 3, 3, device=cuda)
    pred = torch.cat([u, c]).reshape(-1, 4)
    sig, sig_n = 700.0, 545.729248046875
    xn, v = ops.cfg_euler_step(pred, g, x, sig, sig_n, cfg=True, want_v=True)
    vv = (u + g[None, :, None, None, None] * (c - u)).permute(0, 1, 4, 2, 3)
    x0 = vv * (-sig / (sig ** 2 + 1) ** 0.5) + x / (sig ** 2 + 1)
    ref = x + (x - x0) / sig * (sig_n - sig)
    assert rel_l2(v, vv) < 1e-6
    assert rel_l2(xn, ref) < 1e-6
    # the two halves at separate bases (CFG-pair split: one of them is the partner GPU's peer-mapped buffer; here both are
    # local, wider-pitched and offset slices): the same bits, also with device sigmas and the in-place latent update
    n = 2 * 3 * 8 * 8
    wide_u = torch.zeros(n, 8, device=cuda); wide_c = torch.full((n + 5, 8), 7.0, device=cuda)
    wide_u[:, :4] = u.reshape(-1, 4); wide_c[5:, :4] = c.reshape(-1, 4)
    xn2, v2 = ops.cfg_euler_step(wide_u, g, x, sig, sig_n, cfg=True, want_v=True, pred_cond=wide_c[5:])
    assert torch.equal(xn2, xn) and torch.equal(v2, v)
    xs = x.clone()
    xn3, _ = ops.cfg_euler_step(wide_u, g, xs, 1.0, 1.0, cfg=True, pred_cond=wide_c[5:], in_place=True,
                                sigmas_dev=torch.tensor([sig, sig_n], device=cuda))
    assert xn3.data_ptr() == xs.data_ptr() and torch.equal(xs, xn)
    with pytest.raises(ValueError):
        ops.cfg_euler_step(wide_u, g, x, sig, sig_n, cfg=True, pred_cond=c.reshape(-1, 4))     # different row pitch
    with pytest.raises(ValueError):
        ops.cfg_euler_step(wide_u, None, x, sig, sig_n, cfg=False, pred_cond=wide_c[5:])       # needs guidance


# ------------------------------------------------------------------------------------------------- fp32 stream
def test_fp32_residual_stream_variants(cuda):
    """The residual stream is fp32: norms read it, GEMM epilogues add / write it, glue narrows it to bf16."""
    from lkgd_b200 import ops
    M, K, N = 640, 128, 320
    A = rnd(M, K, dev=cuda)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5)
    r1 = rnd(M, N, dev=cuda, dtype=torch.float32, seed=1)
    r2 = rnd(M, N, dev=cuda, dtype=torch.float32, seed=2)
    out = ops.gemm(A, W, res1=r1, s1=1.0, res2=r2, s2=0.5, s0=0.25, out_f32=True)
    ref = 0.25 * (A.float() @ W.float().t()) + r1 + 0.5 * r2.float()
    assert out.dtype == torch.float32 and rel_l2(out, ref) < 1e-5
    chk = ops.gemm(A, W, res1=r1, s1=1.0, res2=r2, s2=0.5, s0=0.25, out_f32=True, checker=True)
    assert rel_l2(out, chk) < 1e-5
    # the AlphaBlender mix: bf16 output (proj_out's GEMM operand) from two fp32 residual-stream tensors
    mix = ops.gemm(A, W, res1=r1, s1=0.4, res2=r2, s2=0.6, s0=0.4)
    assert mix.dtype == bf16 and rel_l2(mix.float(), 0.4 * (A.float() @ W.float().t()) + 0.4 * r1 + 0.6 * r2) < 4e-3
    # bf16 residual into an fp32 output (ControlNet condition embedding added to conv_in)
    rb = rnd(M, N, dev=cuda, seed=3)
    o2 = ops.gemm(A, W, res1=rb, out_f32=True)
    assert rel_l2(o2, A.float() @ W.float().t() + rb.float()) < 1e-5
    with pytest.raises(ValueError):      # two residuals must share a dtype
        ops.gemm(A, W, res1=r1, res2=rb, out_f32=True)
    # GroupNorm / LayerNorm on fp32 input
    NS, R, C1, C2 = 2, 300, 64, 32
    x1 = rnd(NS * R, C1, dev=cuda, dtype=torch.float32) + 0.3
    x2 = rnd(NS * R, C2, dev=cuda, dtype=torch.float32, seed=4) * 2
    g = rnd(C1 + C2, dev=cuda, dtype=torch.float32) * 0.2 + 1
    b = rnd(C1 + C2, dev=cuda, dtype=torch.float32, seed=6) * 0.2
    got = ops.groupnorm(x1, g, b, 1e-6, NS=NS, R=R, x2=x2, silu=True)
    xr = torch.cat([x1, x2], 1).view(NS, R, -1).permute(0, 2, 1)
    ref = F.silu(F.group_norm(xr, 32, g, b, 1e-6)).permute(0, 2, 1).reshape(NS * R, -1)
    assert got.dtype == bf16 and rel_l2(got.float(), ref) < 4e-3
    x = rnd(500, 320, dev=cuda, dtype=torch.float32) * 3
    gg, bb = torch.ones(320, device=cuda), torch.zeros(320, device=cuda)
    add = rnd(5, 320, dev=cuda, dtype=torch.float32, seed=8)
    s = torch.empty_like(x)
    got = ops.layernorm(x, gg, bb, 1e-5, addvec=add, rv=(ops.RV_FRAME, 100, 1, 1), sum_out=s)
    s_ref = x + add[torch.arange(500, device=cuda) // 100]
    assert torch.equal(s, s_ref)
    assert rel_l2(got.float(), F.layer_norm(s_ref, (320,), gg, bb, 1e-5)) < 4e-3
    # glue: cast, upsample (fp32 -> bf16), concat (fp32 -> bf16), axpby (bf16 into fp32)
    assert torch.equal(ops.cast_bf16(x), x.to(bf16))
    xs = rnd(2 * 3 * 4, 64, dev=cuda, dtype=torch.float32)
    up = ops.upsample2x(xs, 2, 3, 4)
    ref = F.interpolate(xs.view(2, 3, 4, 64).permute(0, 3, 1, 2), scale_factor=2.0, mode="nearest")
    assert torch.equal(up.view(2, 6, 8, 64), ref.permute(0, 2, 3, 1).to(bf16))
    assert torch.equal(ops.concat_channels(x1, x2), torch.cat([x1, x2], 1).to(bf16))
    y = x.clone()
    xb = rnd(500, 320, dev=cuda, seed=12)
    ops.axpby(xb, 4.0, y, 1.0)
    assert rel_l2(y, x + 4.0 * xb.float()) < 1e-6


# ------------------------------------------------------------------------------------------------- CTA pairs
def _both_gemm_variants(fn):
    """Runs ``fn`` with the one-CTA kernel and with the cta_group::2 pair kernel forced (LKGD_GEMM_1CTA / _2CTA are read
    per call)."""
    import os
    os.environ["LKGD_GEMM_1CTA"] = "1"
    try:
        one = fn().clone()
    finally:
        del os.environ["LKGD_GEMM_1CTA"]
    os.environ["LKGD_GEMM_2CTA"] = "1"
    try:
        two = fn().clone()
    finally:
        del os.environ["LKGD_GEMM_2CTA"]
    return one, two


@pytest.mark.parametrize("M,N,K", [(1536, 320, 2880), (1536, 160, 512), (1536, 640, 512), (1536, 256, 512),
                                   (1100, 1280, 1280), (130, 48, 72)])
def test_gemm_cta_pairs_match_single_cta_linear(cuda, M, N, K):
    """Same accumulation order in both kernels -> bit-identical results, for every epilogue flavour (this caught a
    staging-buffer toggle that assumed a 4 KB aligned base: only BN = 160 pair tiles broke it)."""
    from lkgd_b200 import ops
    A = rnd(M, K, dev=cuda, seed=1)
    W = rnd(N, K, dev=cuda, scale=K ** -0.5, seed=2)
    b = rnd(N, dev=cuda, dtype=torch.float32, seed=3)
    r = rnd(M, N, dev=cuda, dtype=torch.float32, seed=4)
    rb = r.to(bf16)
    cases = {"plain": dict(), "f32": dict(out_f32=True), "res32->f32": dict(res1=r, out_f32=True),
             "res32->bf16": dict(res1=r), "resbf->bf16": dict(res1=rb), "resbf->f32": dict(res1=rb, out_f32=True),
             "2res32": dict(res1=r, res2=r, s2=0.5, out_f32=True), "2resbf": dict(res1=rb, res2=rb, s1=0.3, s2=-1.0)}
    for name, kw in cases.items():
        one, two = _both_gemm_variants(lambda: ops.gemm(A, W, bias=b, **kw))
        assert torch.equal(one, two), name
    if N % 256 == 0:
        Wp, bp = ops.pack_geglu(W, b)
        one, two = _both_gemm_variants(lambda: ops.gemm(A, Wp, bias=bp, act=ops.ACT_GEGLU))
        assert torch.equal(one, two)
    ref = A.float() @ W.float().t() + b + r
    assert rel_l2(ops.gemm(A, W, bias=b, res1=r, out_f32=True), ref) < 1e-5


@pytest.mark.parametrize("n,h,w,ci,co,stride", [(6, 16, 16, 320, 320, 1), (5, 12, 20, 192, 640, 1), (3, 16, 16, 64, 32, 1),
                                                (6, 16, 16, 128, 256, 2)])
def test_gemm_cta_pairs_match_single_cta_conv(cuda, n, h, w, ci, co, stride):
    from lkgd_b200 import ops
    A = rnd(n * h * w, ci, dev=cuda, seed=1)
    W = rnd(co, 9 * ci, dev=cuda, scale=(9 * ci) ** -0.5, seed=2)
    b = rnd(co, dev=cuda, dtype=torch.float32, seed=3)
    ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
    r = rnd(n * ho * wo, co, dev=cuda, dtype=torch.float32, seed=4)
    for kw in (dict(out_f32=True), dict(res1=r, out_f32=True), dict()):
        one, two = _both_gemm_variants(lambda: ops.gemm(A, W, mode=ops.A_CONV3X3, conv=(n, h, w, stride), bias=b, **kw))
        assert torch.equal(one, two)
    Bt, Ft, HWt = 2, n, h * w // 2
    At = rnd(Bt * Ft * HWt, ci, dev=cuda, seed=5)
    Wt = rnd(co, 3 * ci, dev=cuda, scale=(3 * ci) ** -0.5, seed=6)
    one, two = _both_gemm_variants(lambda: ops.gemm(At, Wt, mode=ops.A_TCONV3, tconv=(Bt, Ft, HWt), out_f32=True))
    assert torch.equal(one, two)


@pytest.mark.parametrize("N,Cc,H,W", [(2, 2, 64, 96), (1, 3, 50, 70), (3, 1, 17, 33), (1, 4, 8, 256)])
def test_cond_conv_in(cuda, N, Cc, H, W):
    """ControlNet condition stem (models/controlnet_sdv.py:104-105): fp32 planar frames -> SiLU(conv3x3) in bf16
    channels-last, against torch conv2d."""
    from lkgd_b200 import ops
    g = torch.Generator().manual_seed(0)
    x = torch.rand(N, Cc, H, W, generator=g) * 2 - 1
    w = torch.randn(16, Cc, 3, 3, generator=g) * (Cc * 9) ** -0.5
    b = torch.randn(16, generator=g) * 0.1
    ref = torch.nn.functional.silu(torch.nn.functional.conv2d(x, w, b, padding=1)).permute(0, 2, 3, 1).reshape(-1, 16)
    out = ops.cond_conv_in(x.to(cuda), w.to(cuda), b.to(cuda))
    assert out.dtype == torch.bfloat16 and tuple(out.shape) == (N * H * W, 16)
    assert rel_l2(out.float(), ref) < 4e-3


@pytest.mark.parametrize("N,H,W,Cin,Cout,silu", [(2, 64, 96, 16, 16, True), (1, 50, 70, 32, 32, True), (3, 9, 33, 16, 32, False),
                                                 (1, 4, 32, 16, 16, True), (2, 130, 5, 32, 32, False)])
def test_thin_conv3x3(cuda, N, H, W, Cin, Cout, silu):
    """Stride-1 thin convs of the condition encoder (controlnet_sdv.py:107-109) on mma.sync tiles with a cp.async halo:
    image borders, partial 4 x 32 tiles, both channel widths - against torch conv2d on the same bf16 operands."""
    from lkgd_b200 import ops
    g = torch.Generator().manual_seed(1)
    x = (torch.randn(N, H, W, Cin, generator=g)).to(torch.bfloat16)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) * (Cin * 9) ** -0.5).to(torch.bfloat16)
    b = torch.randn(Cout, generator=g) * 0.1
    ref = torch.nn.functional.conv2d(x.float().permute(0, 3, 1, 2), w.float(), b, padding=1)
    if silu:
        ref = torch.nn.functional.silu(ref)
    ref = ref.permute(0, 2, 3, 1).reshape(-1, Cout)
    w9 = w.permute(2, 3, 0, 1).reshape(9, Cout, Cin).contiguous()
    out = ops.thin_conv3x3(x.reshape(-1, Cin).to(cuda), w9.to(cuda), b.to(cuda), N, H, W, silu=silu)
    assert rel_l2(out.float(), ref) < 4e-3
    # same numbers as the tcgen05 implicit-GEMM path it replaces
    wk = w.float().permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).to(torch.bfloat16).to(cuda)
    via_gemm = ops.gemm(x.reshape(-1, Cin).to(cuda), wk, mode=ops.A_CONV3X3, conv=(N, H, W, 1), bias=b.to(cuda),
                        act=ops.ACT_SILU if silu else ops.ACT_NONE)
    assert rel_l2(out.float(), via_gemm.float()) < 4e-3


# ------------------------------------------------------------------------------------------------- CLIP image encoder pieces
@pytest.mark.parametrize("N,Cn,S,P,Kpad", [(2, 3, 56, 14, 592), (1, 3, 224, 14, 592), (3, 4, 32, 8, 256)])
def test_patchify(cuda, N, Cn, S, P, Kpad):
    """Patch rows of the non-overlapping Conv2d (CLIP patch embedding): F.unfold's column order, zero-padded to Kpad."""
    from lkgd_b200 import ops
    x = rnd(N, Cn, S, S, dev=cuda, dtype=torch.float32)
    got = ops.patchify(x, P, Kpad)
    ref = F.unfold(x, P, stride=P).transpose(1, 2).reshape(-1, Cn * P * P).to(bf16)
    assert tuple(got.shape) == (N * (S // P) ** 2, Kpad)
    assert torch.equal(got[:, :Cn * P * P], ref) and float(got[:, Cn * P * P:].abs().max() if Kpad > Cn * P * P else 0) == 0


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_gemm_gelu_epilogues(cuda, act):
    from lkgd_b200 import ops
    A, W, b = rnd(500, 192, dev=cuda), rnd(320, 192, dev=cuda, scale=0.1), rnd(320, dev=cuda, dtype=torch.float32)
    code = ops.ACT_GELU if act == "gelu" else ops.ACT_QUICK_GELU
    out = ops.gemm(A, W, bias=b, act=code)
    y = A.float() @ W.float().t() + b
    ref = F.gelu(y) if act == "gelu" else y * torch.sigmoid(1.702 * y)
    assert rel_l2(out.float(), ref) < 4e-3
    assert rel_l2(ops.gemm(A, W, bias=b, act=code, checker=True).float(), ref) < 4e-3


@pytest.mark.parametrize("n_img,heads,d,N", [(2, 16, 80, 257), (1, 2, 80, 17), (2, 3, 96, 300), (1, 2, 72, 64)])
def test_attention_head_widths_between_64_and_128(cuda, n_img, heads, d, N):
    """ViT-H heads are 80 wide: the two-sub-tile kernel zero-fills the channels above d."""
    from lkgd_b200 import ops
    Cn = heads * d
    qkv = rnd(n_img * N, 3 * Cn, dev=cuda)
    q, k, v = qkv[:, :Cn], qkv[:, Cn:2 * Cn], qkv[:, 2 * Cn:]
    out = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N)
    ref = _attn_ref(q.contiguous(), k.contiguous(), v.contiguous(), n_img, heads, d, N, N)
    assert rel_l2(out.float(), ref) < 6e-3
