"""Torch-tensor front end of the C ABI: validates, allocates outputs with the caching allocator, passes raw
device pointers + the current CUDA stream.  No computation happens in PyTorch here."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import (A_CONV3X3, A_LINEAR, A_TCONV3, ACT_GEGLU, ACT_NONE, ACT_SILU, RV_BATCH, RV_FRAME, RV_FRAMEPOS,
                   RV_NONE, RV_TCTX_0272, GemmArgs)

bf16 = torch.bfloat16


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise ValueError("lkgd_b200 kernels take CUDA tensors only (there is no CPU path)")


def launch_count() -> int:
    return int(L.load().lkgd_launch_count())


def device_check(dev: int = 0):
    L.check(L.load().lkgd_device_check(dev), "device_check")


# ----------------------------------------------------------------------------------------------- GEMM
def gemm(A: torch.Tensor, Bw: torch.Tensor, *, mode: int = A_LINEAR, M: Optional[int] = None,
         bias: Optional[torch.Tensor] = None, A1: Optional[torch.Tensor] = None, Bw1: Optional[torch.Tensor] = None,
         conv: Optional[Tuple[int, int, int, int]] = None, tconv: Optional[Tuple[int, int, int]] = None,
         rowvec: Optional[torch.Tensor] = None, rv: Tuple[int, int, int, int] = (RV_NONE, 1, 1, 1),
         act: int = ACT_NONE, s0: float = 1.0, res1: Optional[torch.Tensor] = None, s1: float = 1.0,
         res2: Optional[torch.Tensor] = None, s2: float = 1.0, out: Optional[torch.Tensor] = None,
         out_f32: bool = False, n_store: int = 0, checker: bool = False) -> torch.Tensor:
    """out = s0*act(A (*) Bw^T [+ A1 Bw1^T] + bias + rowvec[g(m)]) + s1*res1 + s2*res2   (see lkgd_gemm).

    LINEAR: ``A`` is [M, K] (may be a column slice of a wider row-major matrix).  CONV3X3: ``A`` is contiguous
    [NIMG, Hin, Win, C], ``conv=(NIMG, Hin, Win, stride)``, ``Bw`` [N, 9*C].  TCONV3: ``A`` contiguous
    [B, F, HW, C], ``tconv=(B, F, HW)``, ``Bw`` [N, 3*C]."""
    _need_cuda(A, Bw, bias, A1, Bw1, rowvec, res1, res2, out)
    if A.dtype != bf16 or Bw.dtype != bf16:
        raise ValueError("gemm operands must be bfloat16")
    a = GemmArgs()
    a.a_mode = mode
    N = Bw.shape[0]
    if mode == A_LINEAR:
        if A.dim() != 2 or A.stride(1) != 1:
            raise ValueError("LINEAR A must be a 2-D row-major matrix (column slices allowed)")
        M = A.shape[0]
        a.K0, a.lda = A.shape[1], A.stride(0)
    elif mode == A_CONV3X3:
        if conv is None or not A.is_contiguous():
            raise ValueError("CONV3X3 needs conv=(NIMG,Hin,Win,stride) and a contiguous NHWC tensor")
        a.NIMG, a.Hin, a.Win, a.stride = conv
        a.K0 = A.shape[-1]
        s = conv[3]
        M = conv[0] * ((conv[1] - 1) // s + 1) * ((conv[2] - 1) // s + 1)
    elif mode == A_TCONV3:
        if tconv is None or not A.is_contiguous():
            raise ValueError("TCONV3 needs tconv=(B,F,HW) and a contiguous [B,F,HW,C] tensor")
        a.NIMG, a.F, a.HW = tconv
        a.K0 = A.shape[-1]
        M = tconv[0] * tconv[1] * tconv[2]
    else:
        raise ValueError(f"unknown a_mode {mode}")
    taps = {A_LINEAR: 1, A_CONV3X3: 9, A_TCONV3: 3}[mode]
    if Bw.dim() != 2 or Bw.stride(1) != 1 or Bw.shape[1] != taps * a.K0:
        raise ValueError(f"Bw must be [N, {taps}*K0={taps * a.K0}], got {tuple(Bw.shape)}")
    a.M, a.N = M, N
    a.A, a.Bw, a.ldb = A.data_ptr(), Bw.data_ptr(), Bw.stride(0)
    if A1 is not None:
        if Bw1 is None or Bw1.shape[0] != N or A1.dtype != bf16 or Bw1.dtype != bf16:
            raise ValueError("segment 1 needs A1 [M,K1] and Bw1 [N,K1] in bfloat16")
        a.K1 = A1.shape[-1]
        a.A1, a.lda1 = A1.data_ptr(), (A1.stride(0) if A1.dim() == 2 else a.K1)
        a.Bw1, a.ldb1 = Bw1.data_ptr(), Bw1.stride(0)
    n_cols = N // 2 if act == ACT_GEGLU else N
    n_out = n_store if n_store > 0 else n_cols
    if out is None:
        out = torch.empty((M, n_out), device=A.device, dtype=torch.float32 if out_f32 else bf16)
    elif out.dim() != 2 or out.shape[0] != M or out.stride(1) != 1:
        raise ValueError("out must be a [M, >=n] row-major matrix")
    if (out.dtype == torch.float32) != bool(out_f32):
        raise ValueError("out dtype does not match out_f32")
    if bias is not None and (bias.dtype != torch.float32 or bias.numel() != N):
        raise ValueError("bias must be fp32 [N]")
    for r in (res1, res2):
        if r is not None and (r.dtype not in (bf16, torch.float32) or r.dim() != 2 or r.shape[0] != M
                              or r.stride(1) != 1):
            raise ValueError("residuals must be bf16 or fp32 [M, >=n] row-major")
    if rowvec is not None and (rowvec.dtype != torch.float32 or rowvec.dim() != 2 or rowvec.shape[-1] != n_cols
                               or rowvec.stride(1) != 1):
        raise ValueError("rowvec must be fp32 [G, n] with unit column stride (column slices of a wider matrix allowed)")
    a.bias, a.rowvec = _ptr(bias), _ptr(rowvec)
    a.rv_mode, a.rv_HW, a.rv_F, a.rv_B = rv
    a.rv_ld = rowvec.stride(0) if rowvec is not None else 0
    a.act, a.s0 = act, s0
    a.res1, a.ldr1, a.s1 = _ptr(res1), (res1.stride(0) if res1 is not None else 0), s1
    a.res2, a.ldr2, a.s2 = _ptr(res2), (res2.stride(0) if res2 is not None else 0), s2
    a.out, a.ldo, a.out_f32, a.n_store = out.data_ptr(), out.stride(0), int(out_f32), n_store
    a.res1_f32 = int(res1 is not None and res1.dtype == torch.float32)
    a.res2_f32 = int(res2 is not None and res2.dtype == torch.float32)
    lib = L.load()
    fn = lib.lkgd_gemm_simt_check if checker else lib.lkgd_gemm
    if L.PROF.enabled:
        L.PROF.meta = {"flops": 2.0 * M * N * (taps * a.K0 + a.K1), "mode": mode, "M": M, "N": N,
                       "K": taps * a.K0 + a.K1}
    L.check(fn(C.byref(a), _stream()), "lkgd_gemm")
    return out


def pack_geglu(weight: torch.Tensor, bias: Optional[torch.Tensor]):
    """Reorders a GEGLU projection ([8C, K]: value rows then gate rows) into 256-row tiles of 128 value rows
    followed by their 128 gate rows, the layout ``lkgd_gemm(act=GEGLU)`` expects.  Pure data movement."""
    n2 = weight.shape[0]
    half = n2 // 2
    if half % 128:
        raise ValueError("GEGLU inner width must be a multiple of 128")
    idx = torch.arange(half, device=weight.device).view(-1, 128)
    order = torch.cat([idx, idx + half], dim=1).reshape(-1)
    return weight[order].contiguous(), (bias[order].contiguous() if bias is not None else None)


# ----------------------------------------------------------------------------------------------- norms
def groupnorm(x1: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, *, NS: int, R: int,
              x2: Optional[torch.Tensor] = None, groups: int = 32, silu: bool = True,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """x1 [NS*R, C1] (+ x2 [NS*R, C2]) bf16 or fp32 channels-last -> [NS*R, C1+C2] bf16."""
    _need_cuda(x1, x2, gamma, beta)
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    Ct = C1 + C2
    if x1.dtype not in (bf16, torch.float32) or not x1.is_contiguous() or x1.numel() != NS * R * C1:
        raise ValueError("groupnorm: x1 must be contiguous bf16/fp32 [NS*R, C1]")
    if x2 is not None and (x2.dtype != x1.dtype or not x2.is_contiguous() or x2.numel() != NS * R * C2):
        raise ValueError("groupnorm: x2 must be contiguous [NS*R, C2] of x1's dtype")
    if gamma.dtype != torch.float32 or gamma.numel() != Ct or beta.numel() != Ct:
        raise ValueError("groupnorm: gamma/beta must be fp32 [C]")
    lib = L.load()
    if out is None:
        out = torch.empty((NS * R, Ct), device=x1.device, dtype=bf16)
    if L.PROF.enabled:   # algorithmic bytes: input read twice (statistics, then normalise) + bf16 output
        L.PROF.meta = {"bytes": NS * R * Ct * (2 * x1.element_size() + 2)}
    ws_bytes = lib.lkgd_groupnorm_workspace(NS, Ct)
    ws = torch.empty(ws_bytes, device=x1.device, dtype=torch.uint8)
    L.check(lib.lkgd_groupnorm(x1.data_ptr(), C1, _ptr(x2), C2, NS, R, groups, gamma.data_ptr(), beta.data_ptr(),
                               eps, int(silu), int(x1.dtype == torch.float32), out.data_ptr(), ws.data_ptr(), ws_bytes,
                               _stream()), "lkgd_groupnorm")
    return out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, *,
              addvec: Optional[torch.Tensor] = None, rv: Tuple[int, int, int, int] = (RV_NONE, 1, 1, 1),
              sum_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, gamma, beta, addvec)
    if x.dtype not in (bf16, torch.float32) or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("layernorm: x must be contiguous bf16/fp32 [M, C]")
    if sum_out is not None and (sum_out.dtype != x.dtype or not sum_out.is_contiguous()):
        raise ValueError("layernorm: sum_out must be contiguous and of x's dtype")
    M, Cn = x.shape
    if addvec is not None and (addvec.dtype != torch.float32 or addvec.dim() != 2 or addvec.shape[-1] != Cn
                               or addvec.stride(1) != 1):
        raise ValueError("layernorm: addvec must be fp32 [G, C] with unit column stride")
    if out is None:
        out = torch.empty((M, Cn), device=x.device, dtype=bf16)
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": M * Cn * (x.element_size() + 2 + (x.element_size() if sum_out is not None else 0))}
    L.check(L.load().lkgd_layernorm(x.data_ptr(), M, Cn, gamma.data_ptr(), beta.data_ptr(), eps, _ptr(addvec),
                                    addvec.stride(0) if addvec is not None else 0, rv[0], rv[1], rv[2], rv[3], int(x.dtype == torch.float32), _ptr(sum_out),
                                    out.data_ptr(), _stream()),
            "lkgd_layernorm")
    return out


# ----------------------------------------------------------------------------------------------- attention
def attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, *, n_img: int, heads: int, d: int, Nq: int,
              Nk: int, scale: Optional[float] = None, out: Optional[torch.Tensor] = None,
              checker: bool = False) -> torch.Tensor:
    """q [n_img*Nq, >=heads*d], k/v [n_img*Nk, >=heads*d] (column slices of a fused projection are fine)."""
    _need_cuda(q, k, v)
    for t in (q, k, v):
        if t.dtype != bf16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("attention operands must be bf16 2-D row-major (column slices allowed)")
    if out is None:
        out = torch.empty((n_img * Nq, heads * d), device=q.device, dtype=bf16)
    scale = d ** -0.5 if scale is None else scale
    lib = L.load()
    fn = lib.lkgd_attention_simt_check if checker else lib.lkgd_attention
    if L.PROF.enabled:
        L.PROF.meta = {"flops": 4.0 * n_img * heads * Nq * Nk * d}
    L.check(fn(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0), out.data_ptr(),
               out.stride(0), n_img, heads, d, Nq, Nk, scale, _stream()), "lkgd_attention")
    return out


def attention_temporal(qkv: torch.Tensor, *, B: int, F: int, HW: int, heads: int, d: int,
                       scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(qkv)
    Cn = heads * d
    if qkv.dtype != bf16 or not qkv.is_contiguous() or qkv.numel() != B * F * HW * 3 * Cn:
        raise ValueError("attention_temporal: qkv must be contiguous bf16 [B*F*HW, 3*heads*d]")
    if out is None:
        out = torch.empty((B * F * HW, Cn), device=qkv.device, dtype=bf16)
    scale = d ** -0.5 if scale is None else scale
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": B * F * HW * Cn * 2 * 4}
    L.check(L.load().lkgd_attention_temporal(qkv.data_ptr(), out.data_ptr(), B, F, HW, heads, d, scale, _stream()),
            "lkgd_attention_temporal")
    return out


# ----------------------------------------------------------------------------------------------- small fp32
def small_linear(x: torch.Tensor, W: torch.Tensor, b: Optional[torch.Tensor] = None, act_in: int = 0,
                 act_out: int = 0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, W, b)
    if x.dtype != torch.float32 or W.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 \
            or not W.is_contiguous() or W.shape[1] != x.shape[1]:
        raise ValueError("small_linear: x fp32 [M,K] row-major, W contiguous fp32 [N,K]")
    M, K = x.shape
    N = W.shape[0]
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=torch.float32)
    L.check(L.load().lkgd_small_linear(x.data_ptr(), x.stride(0), W.data_ptr(), _ptr(b), out.data_ptr(),
                                       out.stride(0), M, N, K, act_in, act_out, _stream()), "lkgd_small_linear")
    return out


def timestep_embedding(t: torch.Tensor, dim: int) -> torch.Tensor:
    _need_cuda(t)
    t = t.reshape(-1).to(torch.float32).contiguous()
    out = torch.empty((t.numel(), dim), device=t.device, dtype=torch.float32)
    L.check(L.load().lkgd_timestep_embedding(t.data_ptr(), t.numel(), dim, out.data_ptr(), _stream()),
            "lkgd_timestep_embedding")
    return out


def axpy_f32(x: torch.Tensor, y: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """y += alpha * x in place (fp32)."""
    _need_cuda(x, y)
    if x.dtype != torch.float32 or y.dtype != torch.float32 or not x.is_contiguous() or not y.is_contiguous() \
            or x.numel() != y.numel():
        raise ValueError("axpy_f32: contiguous fp32 tensors of equal size")
    L.check(L.load().lkgd_axpy_f32(x.data_ptr(), alpha, y.data_ptr(), x.numel(), _stream()), "lkgd_axpy_f32")
    return y


def scale_f32(x: torch.Tensor, alpha: float) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype != torch.float32:
        raise ValueError("scale_f32: fp32 tensor expected")
    x = x.contiguous()
    y = torch.empty_like(x)
    L.check(L.load().lkgd_scale_f32(x.data_ptr(), alpha, y.data_ptr(), x.numel(), _stream()), "lkgd_scale_f32")
    return y


def polar(a: torch.Tensor, b: torch.Tensor, mode: int):
    """mode 0: (re, im) -> (mag, pha); mode 1: (mag, pha) -> (re, im).  fp32 contiguous."""
    _need_cuda(a, b)
    a, b = a.contiguous(), b.contiguous()
    o0, o1 = torch.empty_like(a), torch.empty_like(a)
    L.check(L.load().lkgd_polar(a.data_ptr(), b.data_ptr(), o0.data_ptr(), o1.data_ptr(), a.numel(), mode,
                                _stream()), "lkgd_polar")
    return o0, o1


# ----------------------------------------------------------------------------------------------- glue
def pack_input(src0: torch.Tensor, scale0: float, src1: Optional[torch.Tensor], N: int, Cpad: int) -> torch.Tensor:
    """src [N?,F,C?,H,W] fp32 -> bf16 [N*F*H*W, Cpad] channels-last (see lkgd_pack_input)."""
    _need_cuda(src0, src1)
    src0 = src0.to(torch.float32).contiguous()
    N0, F, C0, H, W = src0.shape
    N1 = C1 = 0
    if src1 is not None:
        src1 = src1.to(torch.float32).contiguous()
        N1, _, C1 = src1.shape[:3]
    out = torch.empty((N * F * H * W, Cpad), device=src0.device, dtype=bf16)
    L.check(L.load().lkgd_pack_input(src0.data_ptr(), N0, C0, scale0, _ptr(src1), max(N1, 1), C1, out.data_ptr(), N,
                                     F, H, W, Cpad, _stream()), "lkgd_pack_input")
    return out


def unpack_output(src: torch.Tensor, N: int, F: int, Cn: int, H: int, W: int) -> torch.Tensor:
    _need_cuda(src)
    if src.dtype != torch.float32 or src.dim() != 2 or src.stride(1) != 1:
        raise ValueError("unpack_output: src must be fp32 [N*F*H*W, ld]")
    out = torch.empty((N, F, Cn, H, W), device=src.device, dtype=torch.float32)
    L.check(L.load().lkgd_unpack_output(src.data_ptr(), src.stride(0), out.data_ptr(), N * F, Cn, H, W, _stream()),
            "lkgd_unpack_output")
    return out


def nchw_to_nhwc(src: torch.Tensor) -> torch.Tensor:
    _need_cuda(src)
    src = src.to(torch.float32).contiguous()
    N, Cn, H, W = src.shape
    out = torch.empty((N * H * W, Cn), device=src.device, dtype=bf16)
    L.check(L.load().lkgd_nchw_to_nhwc(src.data_ptr(), out.data_ptr(), N, Cn, H, W, _stream()), "lkgd_nchw_to_nhwc")
    return out


def nhwc_to_nchw(src: torch.Tensor, N: int, H: int, W: int) -> torch.Tensor:
    _need_cuda(src)
    Cn = src.shape[-1]
    if src.dtype != bf16 or not src.is_contiguous():
        raise ValueError("nhwc_to_nchw: src must be contiguous bf16")
    out = torch.empty((N, Cn, H, W), device=src.device, dtype=torch.float32)
    L.check(L.load().lkgd_nhwc_to_nchw(src.data_ptr(), out.data_ptr(), N, Cn, H, W, _stream()), "lkgd_nhwc_to_nchw")
    return out


def upsample2x(src: torch.Tensor, N: int, H: int, W: int) -> torch.Tensor:
    _need_cuda(src)
    Cn = src.shape[-1]
    out = torch.empty((N * 4 * H * W, Cn), device=src.device, dtype=bf16)
    L.check(L.load().lkgd_upsample2x(src.data_ptr(), int(src.dtype == torch.float32), out.data_ptr(), N, H, W, Cn,
                                     _stream()), "lkgd_upsample2x")
    return out


def concat_channels(a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    _need_cuda(a, b)
    M = a.shape[0]
    out = torch.empty((M, a.shape[1] + b.shape[1]), device=a.device, dtype=bf16)
    if a.dtype != b.dtype or not a.is_contiguous() or not b.is_contiguous():
        raise ValueError("concat_channels: contiguous sources of one dtype")
    L.check(L.load().lkgd_concat_channels(a.data_ptr(), a.shape[1], b.data_ptr(), b.shape[1],
                                          int(a.dtype == torch.float32), out.data_ptr(), M, _stream()),
            "lkgd_concat_channels")
    return out


def axpby(x: torch.Tensor, alpha: float, y: torch.Tensor, beta: float) -> torch.Tensor:
    """y = alpha*x + beta*y in place; x and y each bf16 or fp32."""
    _need_cuda(x, y)
    ok = (bf16, torch.float32)
    if x.dtype not in ok or y.dtype not in ok or x.numel() != y.numel() or not x.is_contiguous() \
            or not y.is_contiguous():
        raise ValueError("axpby: contiguous bf16/fp32 tensors of equal size")
    L.check(L.load().lkgd_axpby(x.data_ptr(), int(x.dtype == torch.float32), alpha, y.data_ptr(),
                                int(y.dtype == torch.float32), beta, x.numel(), _stream()), "lkgd_axpby")
    return y


def cast_bf16(x: torch.Tensor) -> torch.Tensor:
    """fp32 -> bf16 copy (a residual-stream tensor that a GEMM reads raw)."""
    _need_cuda(x)
    if x.dtype == bf16:
        return x
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("cast_bf16: contiguous fp32 tensor expected")
    out = torch.empty(x.shape, device=x.device, dtype=bf16)
    L.check(L.load().lkgd_cast_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "lkgd_cast_bf16")
    return out


def cfg_euler_step(pred: torch.Tensor, guidance: Optional[torch.Tensor], x: torch.Tensor, sigma: float,
                   sigma_next: float, *, cfg: bool, want_v: bool = False):
    """pred fp32: channels-last rows [(2)S*F*H*W, ld] (2-D) or the latent's own layout [(2)S,F,C,H,W] (5-D);
    x fp32 [S,F,C,H,W] -> (x_next, v or None)."""
    _need_cuda(pred, guidance, x)
    S, F, Cn, H, W = x.shape
    x = x.contiguous()
    if pred.dtype != torch.float32 or x.dtype != torch.float32:
        raise ValueError("cfg_euler_step works in fp32")
    if pred.dim() == 5:
        pred = pred.contiguous()
        ld = 0
        if pred.numel() != (2 if cfg else 1) * x.numel():
            raise ValueError("prediction shape does not match the latent")
    else:
        ld = pred.stride(0)
        if pred.shape[0] != (2 if cfg else 1) * S * F * H * W:
            raise ValueError("prediction rows do not match the latent")
    x_next = torch.empty_like(x)
    v = torch.empty_like(x) if want_v else None
    L.check(L.load().lkgd_cfg_euler_step(pred.data_ptr(), ld, int(cfg), _ptr(guidance), x.data_ptr(),
                                         x_next.data_ptr(), _ptr(v), S, F, Cn, H, W, sigma, sigma_next, _stream()),
            "lkgd_cfg_euler_step")
    return x_next, v
