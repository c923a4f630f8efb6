"""Torch-tensor front end of the C ABI: validates, allocates outputs with the caching allocator, passes raw
device pointers + the current CUDA stream.  No computation happens in PyTorch here."""
from __future__ import annotations

import os
import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib as L
from ._lib import (A_CONV3X3, A_LINEAR, A_TCONV3, ACT_GEGLU, ACT_GELU, ACT_NONE, ACT_QUICK_GELU, ACT_SILU, RV_BATCH, RV_FRAME, RV_FRAMEPOS,
                   RV_BATCH_TCTX, RV_NONE, RV_TCTX_0272, GemmArgs)

bf16 = torch.bfloat16


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    """Every operand on ONE CUDA device, and that device current: the C ABI launches on the calling thread's current
    device with the stream handle it is given, so a tensor of another device would be an illegal address (or a silent
    peer access).  The module / pipeline / trainer entry points switch devices themselves (``on_own_device``)."""
    dev = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise ValueError("lkgd_b200 kernels take CUDA tensors only (there is no CPU path)")
        if dev is None:
            dev = t.device.index
        elif t.device.index != dev:
            raise ValueError(f"lkgd_b200 op with operands on cuda:{dev} and cuda:{t.device.index}")
    if dev is not None and dev != torch.cuda.current_device():
        raise ValueError(f"lkgd_b200 op on cuda:{dev} tensors while the current device is "
                         f"cuda:{torch.cuda.current_device()}: call it under `with torch.cuda.device({dev})`")


def on_own_device(fn):
    """Method decorator for the public entry points (module forward, pipeline step, trainer step): runs the call with
    the object's own CUDA device current, like a PyTorch module of the reference would on any device."""
    import functools

    @functools.wraps(fn)
    def wrap(self, *a, **k):
        dev = getattr(self, "device", None)
        if not isinstance(dev, torch.device) or dev.type != "cuda" or dev.index in (None, torch.cuda.current_device()):
            return fn(self, *a, **k)
        with torch.cuda.device(dev):
            return fn(self, *a, **k)
    return wrap


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
         out_f32: bool = False, n_store: int = 0, checker: bool = False, gn_rows: int = 0,
         out2: Optional[torch.Tensor] = None, want_bf16: bool = False, pad_br: bool = False):
    """out = s0*act(A (*) Bw^T [+ A1 Bw1^T] + bias + rowvec[g(m)]) + s1*res1 + s2*res2   (see lkgd_gemm).

    ``want_bf16`` (fp32 outputs): also write a bf16 copy of the output in the same epilogue and return ``(out, out2)``
    (``out2`` may be given); the copy is the A operand of the GEMM that reads this tensor next.

    ``gn_rows`` > 0 (fp32 outputs): also accumulate the GroupNorm statistics of the output per (frame image of
    ``gn_rows`` rows, channel) in the epilogue; they travel with the returned tensor (``gn_stats_of``) and let the
    ``groupnorm`` that consumes it skip its statistics pass.

    LINEAR: ``A`` is [M, K] (may be a column slice of a wider row-major matrix).  CONV3X3: ``A`` is contiguous
    [NIMG, Hin, Win, C], ``conv=(NIMG, Hin, Win, stride)``, ``Bw`` [N, 9*C] (``pad_br``: stride 2 with the input padded on
    the bottom / right only - diffusers ``Downsample2D(padding=0)`` of the VAE encoder).  TCONV3: ``A`` contiguous
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
        if pad_br:
            if s != 2:
                raise ValueError("pad_br (bottom / right padding only) is the stride-2 convolution of Downsample2D(padding=0)")
            a.pad_br = 1
            M = conv[0] * ((conv[1] - 2) // 2 + 1) * ((conv[2] - 2) // 2 + 1)
        else:
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
    if want_bf16 or out2 is not None:
        if not out_f32:
            raise ValueError("want_bf16 / out2: the primary output must be fp32")
        if out2 is None:
            out2 = torch.empty((M, n_out), device=A.device, dtype=bf16)
        elif out2.dtype != bf16 or out2.dim() != 2 or out2.shape[0] != M or out2.stride(1) != 1 or not out2.is_cuda:
            raise ValueError("out2 must be a bf16 [M, >=n] row-major CUDA matrix")
        a.out2, a.ldo2 = out2.data_ptr(), out2.stride(0)
    lib = L.load()
    fn = lib.lkgd_gemm_simt_check if checker else lib.lkgd_gemm
    stats = None
    if gn_rows > 0 and not checker and not os.environ.get("LKGD_NO_FUSED_GN"):   # switch: A/B measurements only
        stats = STATS_ARENA.take(M // gn_rows, N, A.device)
        a.gn_stats, a.gn_rows = stats.data_ptr(), gn_rows
    if L.PROF.enabled:
        L.PROF.meta = {"flops": 2.0 * M * N * (taps * a.K0 + a.K1), "mode": mode, "M": M, "N": N,
                       "K": taps * a.K0 + a.K1}
    L.check(fn(C.byref(a), _stream()), "lkgd_gemm")
    _set_gn_stats(out, stats, gn_rows)
    return (out, out2) if out2 is not None else out


class _StatsArena:
    """One zeroed fp64 buffer per forward for all fused GroupNorm statistics (a C3 step has ~105 producers: one memset
    instead of 105 fill launches).  ``begin`` is called at the start of a forward; slices are valid until the next
    ``begin`` on the same (device, stream) - every (device, stream) pair has its own buffer, so two models, two devices
    or two streams in one process never share statistics.  Falls back to ``torch.zeros`` when it is not active or too
    small (it then grows at the next ``begin``)."""

    class _Slot:
        __slots__ = ("buf", "off", "want")

        def __init__(self):
            self.buf, self.off, self.want = None, 0, 8 << 20      # doubles

    def __init__(self):
        self.slots = {}

    @staticmethod
    def _key(device):
        return (device.index, torch.cuda.current_stream(device).cuda_stream)

    def begin(self, device) -> None:
        s = self.slots.setdefault(self._key(device), self._Slot())
        if s.buf is None or s.buf.numel() < s.want:
            s.buf = torch.empty(s.want, device=device, dtype=torch.float64)
        s.buf.zero_()
        s.off = 0

    def take(self, frames: int, n: int, device) -> torch.Tensor:
        need = frames * n * 2
        s = self.slots.get(self._key(device))
        if s is None or s.buf is None:
            return torch.zeros((frames, n, 2), device=device, dtype=torch.float64)
        if s.off + need > s.buf.numel():
            s.want = max(s.want, 2 * (s.off + need))
            s.off += need                                      # keep counting so that `want` covers the whole forward
            return torch.zeros((frames, n, 2), device=device, dtype=torch.float64)
        out = s.buf[s.off:s.off + need].view(frames, n, 2)
        s.off += need
        return out


STATS_ARENA = _StatsArena()


def _set_gn_stats(t: torch.Tensor, stats: Optional[torch.Tensor], gn_rows: int = 0) -> None:
    """Attach (or, with None, drop) fused GroupNorm statistics.  Every op that writes a tensor in place drops them."""
    if stats is None:
        if hasattr(t, "_gn_stats"):
            del t._gn_stats
    else:
        t._gn_stats = (stats, gn_rows)


def gn_stats_of(t: Optional[torch.Tensor]):
    """(stats [frames, C, 2] float64, rows per frame) accumulated by the producing ``gemm``, or None."""
    return getattr(t, "_gn_stats", None) if t is not None else None


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
              out: Optional[torch.Tensor] = None, return_stats: bool = False, want_raw: bool = False):
    """x1 [NS*R, C1] (+ x2 [NS*R, C2]) bf16 or fp32 channels-last -> [NS*R, C1+C2] bf16.
    ``return_stats``: also return the per-(sample, channel) sums the backward reuses.
    ``want_raw``: return ``(out, raw)`` where ``raw`` is the un-normalised (concatenated) input narrowed to bf16 - written
    by the same pass when the statistics come fused from the producers, else ``None`` (the caller narrows it itself)."""
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
    ws_bytes = lib.lkgd_groupnorm_workspace(NS, Ct)
    ws = torch.empty(ws_bytes, device=x1.device, dtype=torch.uint8)
    st1, st2 = gn_stats_of(x1), gn_stats_of(x2)
    fused = (st1 is not None and (x2 is None or (st2 is not None and st2[1] == st1[1]))
             and R % st1[1] == 0 and st1[0].shape[0] * st1[1] == NS * R)
    if fused:
        raw = torch.empty((NS * R, Ct), device=x1.device, dtype=bf16) if want_raw else None
        if L.PROF.enabled:   # algorithmic bytes: ONE read of the input + bf16 output (+ the raw bf16 copy)
            L.PROF.meta = {"bytes": NS * R * Ct * (x1.element_size() + 2 + (2 if want_raw else 0))}
        L.check(lib.lkgd_groupnorm_from_stats(x1.data_ptr(), C1, st1[0].data_ptr(), _ptr(x2), C2,
                                              st2[0].data_ptr() if x2 is not None else None, NS, R, R // st1[1],
                                              groups, gamma.data_ptr(), beta.data_ptr(), eps, int(silu),
                                              int(x1.dtype == torch.float32), out.data_ptr(), _ptr(raw), ws.data_ptr(),
                                              ws_bytes, _stream()), "lkgd_groupnorm_from_stats")
        if return_stats:                     # the workspace now holds the per-(sample, channel) sums, as after lkgd_groupnorm
            return out, ws
        return (out, raw) if want_raw else out
    if L.PROF.enabled:   # algorithmic bytes: input read twice (statistics, then normalise) + bf16 output
        L.PROF.meta = {"bytes": NS * R * Ct * (2 * x1.element_size() + 2)}
    L.check(lib.lkgd_groupnorm(x1.data_ptr(), C1, _ptr(x2), C2, NS, R, groups, gamma.data_ptr(), beta.data_ptr(),
                               eps, int(silu), int(x1.dtype == torch.float32), out.data_ptr(), ws.data_ptr(), ws_bytes,
                               _stream()), "lkgd_groupnorm")
    if want_raw:
        return out, None
    return (out, ws) if return_stats else out


def layernorm(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float = 1e-5, *,
              addvec: Optional[torch.Tensor] = None, rv: Tuple[int, int, int, int] = (RV_NONE, 1, 1, 1),
              sum_out: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, gamma, beta, addvec)
    if x.dtype not in (bf16, torch.float32) or x.dim() != 2 or not x.is_contiguous():
        raise ValueError("layernorm: x must be contiguous bf16/fp32 [M, C]")
    if sum_out is not None and (sum_out.dtype != x.dtype or not sum_out.is_contiguous()):
        raise ValueError("layernorm: sum_out must be contiguous and of x's dtype")
    if sum_out is not None:
        _set_gn_stats(sum_out, None)
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
              checker: bool = False, return_lse: bool = False):
    """q [n_img*Nq, >=heads*d], k/v [n_img*Nk, >=heads*d] (column slices of a fused projection are fine).
    ``return_lse``: also return the fp32 [n_img, heads, Nq] log2-sum-exp the backward needs."""
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
    if return_lse:
        lse = torch.empty((n_img, heads, Nq), device=q.device, dtype=torch.float32)
        L.check(lib.lkgd_attention_lse(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                       out.data_ptr(), out.stride(0), n_img, heads, d, Nq, Nk, scale, lse.data_ptr(),
                                       _stream()), "lkgd_attention_lse")
        return out, lse
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
    _set_gn_stats(y, None)
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
def pack_input(src0: torch.Tensor, scale0: float, src1: Optional[torch.Tensor], N: int, Cpad: int,
               scale_dev: Optional[torch.Tensor] = None) -> torch.Tensor:
    """src [N?,F,C?,H,W] fp32 -> bf16 [N*F*H*W, Cpad] channels-last (see lkgd_pack_input).  ``scale_dev``: fp32 device
    scalar that replaces ``scale0`` (a step captured in a CUDA graph reads its per-step scale from device memory)."""
    _need_cuda(src0, src1, scale_dev)
    if scale_dev is not None and (scale_dev.dtype != torch.float32 or scale_dev.numel() < 1):
        raise ValueError("pack_input: scale_dev must be an fp32 device scalar")
    src0 = src0.to(torch.float32).contiguous()
    N0, F, C0, H, W = src0.shape
    N1 = C1 = 0
    if src1 is not None:
        src1 = src1.to(torch.float32).contiguous()
        N1, _, C1 = src1.shape[:3]
    out = torch.empty((N * F * H * W, Cpad), device=src0.device, dtype=bf16)
    L.check(L.load().lkgd_pack_input(src0.data_ptr(), N0, C0, scale0, _ptr(scale_dev), _ptr(src1), max(N1, 1), C1,
                                     out.data_ptr(), N, F, H, W, Cpad, _stream()), "lkgd_pack_input")
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


def cond_conv_in(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor) -> torch.Tensor:
    """x fp32 [N, Cc, H, W] -> bf16 channels-last rows [N*H*W, 16] = SiLU(conv3x3(x) + bias) (lkgd_cond_conv_in)."""
    _need_cuda(x, weight, bias)
    x = x.to(torch.float32).contiguous()
    N, Cc, H, W = x.shape
    if weight.dtype != torch.float32 or not weight.is_contiguous() or tuple(weight.shape) != (16, Cc, 3, 3) \
            or bias.dtype != torch.float32 or bias.numel() != 16:
        raise ValueError("cond_conv_in: weight fp32 [16, Cc, 3, 3], bias fp32 [16]")
    out = torch.empty((N * H * W, 16), device=x.device, dtype=bf16)
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": N * H * W * (Cc * 4 + 32), "flops": 2.0 * N * H * W * 9 * Cc * 16}
    L.check(L.load().lkgd_cond_conv_in(x.data_ptr(), N, Cc, H, W, weight.data_ptr(), bias.data_ptr(), out.data_ptr(),
                                       _stream()), "lkgd_cond_conv_in")
    return out


def thin_conv3x3(x: torch.Tensor, weight9: torch.Tensor, bias: torch.Tensor, N: int, H: int, W: int,
                 silu: bool = True) -> torch.Tensor:
    """x bf16 rows [N*H*W, Cin], weight9 bf16 [9, Cout, Cin] -> bf16 rows [N*H*W, Cout] (lkgd_thin_conv3x3)."""
    _need_cuda(x, weight9, bias)
    Cin = x.shape[-1]
    Cout = weight9.shape[1]
    if x.dtype != bf16 or not x.is_contiguous() or x.numel() != N * H * W * Cin or weight9.dtype != bf16 \
            or not weight9.is_contiguous() or tuple(weight9.shape) != (9, Cout, Cin) or bias.dtype != torch.float32 \
            or bias.numel() != Cout:
        raise ValueError("thin_conv3x3: x bf16 [N*H*W, Cin], weight bf16 [9, Cout, Cin], bias fp32 [Cout]")
    out = torch.empty((N * H * W, Cout), device=x.device, dtype=bf16)
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": N * H * W * (Cin + Cout) * 2, "flops": 2.0 * N * H * W * 9 * Cin * Cout}
    L.check(L.load().lkgd_thin_conv3x3(x.data_ptr(), N, H, W, Cin, Cout, weight9.data_ptr(), bias.data_ptr(), int(silu),
                                       out.data_ptr(), _stream()), "lkgd_thin_conv3x3")
    return out


def patchify(x: torch.Tensor, P: int, Kpad: int) -> torch.Tensor:
    """x fp32 [N, C, H, W] -> bf16 [N * (H/P) * (W/P), Kpad] patch rows in the Conv2d weight's column order (lkgd_patchify)."""
    _need_cuda(x)
    x = x.to(torch.float32).contiguous()
    N, Cn, H, W = x.shape
    out = torch.empty((N * (H // P) * (W // P), Kpad), device=x.device, dtype=bf16)
    L.check(L.load().lkgd_patchify(x.data_ptr(), N, Cn, H, W, P, out.data_ptr(), Kpad, _stream()), "lkgd_patchify")
    return out


def softmax_rows(x: torch.Tensor, scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bf16 softmax(scale * x) over the last axis of an fp32 [M, N] matrix (row slices / pitches allowed; lkgd_softmax_rows)."""
    _need_cuda(x, out)
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1:
        raise ValueError("softmax_rows: x must be fp32 [M, N] with unit column stride")
    M, N = x.shape
    if out is None:
        out = torch.empty((M, N), device=x.device, dtype=bf16)
    elif out.dtype != bf16 or out.dim() != 2 or tuple(out.shape) != (M, N) or out.stride(1) != 1:
        raise ValueError("softmax_rows: out must be bf16 [M, N] with unit column stride")
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": M * N * 6}
    L.check(L.load().lkgd_softmax_rows(x.data_ptr(), x.stride(0), M, N, scale, out.data_ptr(), out.stride(0), _stream()),
            "lkgd_softmax_rows")
    return out


def time_conv_out(x: torch.Tensor, weight: torch.Tensor, bias: torch.Tensor, NB: int, F: int, H: int, W: int) -> torch.Tensor:
    """x fp32 rows [NB*F*H*W, ld] (first C columns) -> fp32 [NB*F, C, H, W] = Conv3d(C, C, (3,1,1), padding (1,0,0)) over the
    frame axis (lkgd_time_conv_out); weight fp32 [C, C, 3]."""
    _need_cuda(x, weight, bias)
    Cn = weight.shape[0]
    if x.dtype != torch.float32 or x.dim() != 2 or x.stride(1) != 1 or x.shape[0] != NB * F * H * W or x.shape[1] < Cn:
        raise ValueError("time_conv_out: x must be fp32 [NB*F*H*W, >= C] rows")
    if weight.dtype != torch.float32 or not weight.is_contiguous() or tuple(weight.shape) != (Cn, Cn, 3) \
            or bias.dtype != torch.float32 or bias.numel() != Cn:
        raise ValueError("time_conv_out: weight fp32 [C, C, 3], bias fp32 [C]")
    out = torch.empty((NB * F, Cn, H, W), device=x.device, dtype=torch.float32)
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": NB * F * H * W * Cn * 8}
    L.check(L.load().lkgd_time_conv_out(x.data_ptr(), x.stride(0), weight.data_ptr(), bias.data_ptr(), out.data_ptr(), NB, F,
                                        H * W, Cn, _stream()), "lkgd_time_conv_out")
    return out


def select_rows(srcs, rv: Tuple[int, int, int, int]) -> torch.Tensor:
    """out[m] = srcs[g(m)][m] over equally shaped contiguous bf16 [M, C] matrices (see lkgd_select_rows)."""
    _need_cuda(*srcs)
    M, Cn = srcs[0].shape
    for t in srcs:
        if t.dtype != bf16 or not t.is_contiguous() or t.shape != (M, Cn):
            raise ValueError("select_rows: contiguous bf16 [M, C] matrices of one shape expected")
    out = torch.empty_like(srcs[0])
    arr = (C.c_void_p * len(srcs))(*[t.data_ptr() for t in srcs])
    L.check(L.load().lkgd_select_rows(arr, len(srcs), out.data_ptr(), M, Cn, rv[0], rv[1], rv[2], rv[3], _stream()),
            "lkgd_select_rows")
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
    _set_gn_stats(y, None)
    return y


def fusion_euler_step(v: torch.Tensor, x: torch.Tensor, weights: torch.Tensor, sigma: float, sigma_next: float):
    """Bidirectional Euler step (lkgd_fusion_euler_step): v, x fp32 [2S, F, C, H, W], weights fp32 [F]."""
    _need_cuda(v, x, weights)
    if v.dtype != torch.float32 or x.dtype != torch.float32 or v.shape != x.shape or v.dim() != 5 or v.shape[0] % 2 \
            or not v.is_contiguous() or not x.is_contiguous():
        raise ValueError("fusion_euler_step: contiguous fp32 [2S, F, C, H, W] tensors of equal shape expected")
    S2, F, Cn, H, W = v.shape
    if weights.dtype != torch.float32 or weights.numel() != F or not weights.is_contiguous():
        raise ValueError("fusion_euler_step: weights must be contiguous fp32 [F]")
    out = torch.empty_like(x)
    L.check(L.load().lkgd_fusion_euler_step(v.data_ptr(), x.data_ptr(), weights.data_ptr(), out.data_ptr(), S2 // 2, F,
                                            Cn, H, W, sigma, sigma_next, _stream()), "lkgd_fusion_euler_step")
    return out


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
                   sigma_next: float, *, cfg: bool, want_v: bool = False, want_x0: bool = False,
                   sigmas_dev: Optional[torch.Tensor] = None, in_place: bool = False,
                   pred_cond: Optional[torch.Tensor] = None):
    """pred fp32: channels-last rows [(2)S*F*H*W, ld] (2-D) or the latent's own layout [(2)S,F,C,H,W] (5-D);
    x fp32 [S,F,C,H,W] -> (x_next, v or None) (-> (x_next, v, x0) with ``want_x0``).  ``sigmas_dev``: fp32 device
    [sigma, sigma_next] replacing the host scalars; ``in_place``: write x_next over x (CUDA-graph replay).
    ``pred_cond`` (CFG-pair split): ``pred`` holds the unconditional rows [S*F*H*W, ld] only and ``pred_cond`` the
    conditional ones at their own base - either may be the partner GPU's peer-mapped buffer."""
    _need_cuda(pred, guidance, x, sigmas_dev, pred_cond)
    S, F, Cn, H, W = x.shape
    if pred_cond is not None:
        if in_place and not x.is_contiguous():
            raise ValueError("cfg_euler_step: in_place needs a contiguous latent")
        x = x.contiguous()
        n = S * F * H * W
        for t in (pred, pred_cond):
            if t.dtype != torch.float32 or t.dim() != 2 or t.shape[0] != n or t.stride(1) != 1:
                raise ValueError("cfg_euler_step(pred_cond=): both halves are fp32 rows [S*F*H*W, ld]")
        if pred.stride(0) != pred_cond.stride(0) or not cfg or guidance is None or x.dtype != torch.float32:
            raise ValueError("cfg_euler_step(pred_cond=): equal row pitch, cfg=True, guidance and an fp32 latent expected")
        x_next = x if in_place else torch.empty_like(x)
        v = torch.empty_like(x) if want_v else None
        x0 = torch.empty_like(x) if want_x0 else None
        L.check(L.load().lkgd_cfg_euler_step_pair(pred.data_ptr(), pred_cond.data_ptr(), pred.stride(0),
                                                  guidance.data_ptr(), x.data_ptr(), x_next.data_ptr(), _ptr(v), _ptr(x0),
                                                  S, F, Cn, H, W, sigma, sigma_next, _ptr(sigmas_dev), _stream()),
                "lkgd_cfg_euler_step_pair")
        return (x_next, v, x0) if want_x0 else (x_next, v)
    if in_place and not x.is_contiguous():
        raise ValueError("cfg_euler_step: in_place needs a contiguous latent")
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
    x_next = x if in_place else torch.empty_like(x)
    v = torch.empty_like(x) if want_v else None
    x0 = torch.empty_like(x) if want_x0 else None
    L.check(L.load().lkgd_cfg_euler_step(pred.data_ptr(), ld, int(cfg), _ptr(guidance), x.data_ptr(),
                                         x_next.data_ptr(), _ptr(v), _ptr(x0), S, F, Cn, H, W, sigma, sigma_next,
                                         _ptr(sigmas_dev), _stream()), "lkgd_cfg_euler_step")
    return (x_next, v, x0) if want_x0 else (x_next, v)


# ----------------------------------------------------------------------------------------------- training step
def attention_bwd(q, k, v, o, dO, lse, dq, dk, dv, *, n_img: int, heads: int, d: int, N: int,
                  scale: Optional[float] = None):
    """Backward of ``attention`` (self-attention): fills dq / dk / dv (bf16, column slices allowed)."""
    _need_cuda(q, k, v, o, dO, lse, dq, dk, dv)
    for t in (q, k, v, o, dO, dq, dk, dv):
        if t.dtype != bf16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("attention_bwd operands must be bf16 2-D row-major (column slices allowed)")
    if o.stride(0) != dO.stride(0):
        raise ValueError("attention_bwd: o and dO must share a row pitch")
    if lse.dtype != torch.float32 or not lse.is_contiguous() or lse.numel() != n_img * heads * N:
        raise ValueError("attention_bwd: lse must be contiguous fp32 [n_img, heads, N]")
    scale = d ** -0.5 if scale is None else scale
    lib = L.load()
    ws_bytes = lib.lkgd_attention_bwd_workspace(n_img, heads, N)
    ws = torch.empty(ws_bytes, device=q.device, dtype=torch.uint8)
    if L.PROF.enabled:
        L.PROF.meta = {"flops": 10.0 * n_img * heads * N * N * d}
    L.check(lib.lkgd_attention_bwd(q.data_ptr(), q.stride(0), k.data_ptr(), k.stride(0), v.data_ptr(), v.stride(0),
                                   o.data_ptr(), dO.data_ptr(), o.stride(0), lse.data_ptr(), dq.data_ptr(),
                                   dq.stride(0), dk.data_ptr(), dk.stride(0), dv.data_ptr(), dv.stride(0), n_img, heads,
                                   d, N, scale, ws.data_ptr(), ws_bytes, _stream()), "lkgd_attention_bwd")


def attention_temporal_bwd(qkv: torch.Tensor, dO: torch.Tensor, *, B: int, F: int, HW: int, heads: int, d: int,
                           scale: Optional[float] = None) -> torch.Tensor:
    _need_cuda(qkv, dO)
    Cn = heads * d
    if qkv.dtype != bf16 or not qkv.is_contiguous() or qkv.numel() != B * F * HW * 3 * Cn:
        raise ValueError("attention_temporal_bwd: qkv must be contiguous bf16 [B*F*HW, 3*heads*d]")
    if dO.dtype != bf16 or not dO.is_contiguous() or dO.numel() != B * F * HW * Cn:
        raise ValueError("attention_temporal_bwd: dO must be contiguous bf16 [B*F*HW, heads*d]")
    out = torch.empty_like(qkv)
    scale = d ** -0.5 if scale is None else scale
    L.check(L.load().lkgd_attention_temporal_bwd(qkv.data_ptr(), dO.data_ptr(), out.data_ptr(), B, F, HW, heads, d,
                                                 scale, _stream()), "lkgd_attention_temporal_bwd")
    return out


def groupnorm_bwd(x1: torch.Tensor, dy: torch.Tensor, stats: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor,
                  eps: float, *, NS: int, R: int, x2: Optional[torch.Tensor] = None, groups: int = 32,
                  silu: bool = True, add: Optional[torch.Tensor] = None, out1: Optional[torch.Tensor] = None,
                  acc1: bool = False, out2: Optional[torch.Tensor] = None, acc2: bool = False,
                  out_bf16: Optional[torch.Tensor] = None):
    """Backward of ``groupnorm`` (see lkgd_groupnorm_bwd): dy bf16 [NS*R, C]; results go to out1 / out2 (fp32, the
    two concatenated sources' gradients, optionally accumulated) and / or out_bf16."""
    _need_cuda(x1, x2, dy, stats, gamma, beta, add, out1, out2, out_bf16)
    C1 = x1.shape[-1]
    C2 = x2.shape[-1] if x2 is not None else 0
    Ct = C1 + C2
    if dy.dtype != bf16 or not dy.is_contiguous() or dy.numel() != NS * R * Ct:
        raise ValueError("groupnorm_bwd: dy must be contiguous bf16 [NS*R, C]")
    if x1.dtype not in (bf16, torch.float32) or not x1.is_contiguous() or x1.numel() != NS * R * C1:
        raise ValueError("groupnorm_bwd: x1 must be contiguous [NS*R, C1]")
    if x2 is not None and (x2.dtype != x1.dtype or not x2.is_contiguous() or x2.numel() != NS * R * C2):
        raise ValueError("groupnorm_bwd: x2 must be contiguous [NS*R, C2] of x1's dtype")
    if add is not None and (add.dtype not in (bf16, torch.float32) or not add.is_contiguous()
                            or add.numel() != NS * R * Ct):
        raise ValueError("groupnorm_bwd: add must be contiguous bf16/fp32 [NS*R, C]")
    for o, cn in ((out1, C1), (out2, C2)):
        if o is not None and (o.dtype != torch.float32 or not o.is_contiguous() or o.numel() != NS * R * cn):
            raise ValueError("groupnorm_bwd: out1 / out2 must be contiguous fp32 [NS*R, C1 / C2]")
    if out_bf16 is not None and (out_bf16.dtype != bf16 or not out_bf16.is_contiguous()
                                 or out_bf16.numel() != NS * R * Ct):
        raise ValueError("groupnorm_bwd: out_bf16 must be contiguous bf16 [NS*R, C]")
    lib = L.load()
    ws_bytes = lib.lkgd_groupnorm_bwd_workspace(NS, Ct)
    ws = torch.empty(ws_bytes, device=x1.device, dtype=torch.uint8)
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": NS * R * Ct * (2 * x1.element_size() + 2 * 2 + 4)}
    L.check(lib.lkgd_groupnorm_bwd(x1.data_ptr(), C1, _ptr(x2), C2, NS, R, groups, gamma.data_ptr(), beta.data_ptr(),
                                   eps, int(silu), int(x1.dtype == torch.float32), dy.data_ptr(), stats.data_ptr(),
                                   _ptr(add), int(add is not None and add.dtype == torch.float32), _ptr(out1),
                                   int(acc1), _ptr(out2), int(acc2), _ptr(out_bf16), ws.data_ptr(), ws_bytes,
                                   _stream()), "lkgd_groupnorm_bwd")


def layernorm_bwd(x: torch.Tensor, dy: torch.Tensor, gamma: torch.Tensor, eps: float, G: torch.Tensor, *,
                  accumulate: bool = True, g_bf16: Optional[torch.Tensor] = None):
    """G (fp32 [M, C]) (+)= LayerNorm'(dy) for the fp32 input ``x``; optional bf16 copy of the new G."""
    _need_cuda(x, dy, gamma, G, g_bf16)
    M, Cn = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous() or G.dtype != torch.float32 or not G.is_contiguous() \
            or G.shape != x.shape:
        raise ValueError("layernorm_bwd: x and G must be contiguous fp32 [M, C]")
    if dy.dtype not in (bf16, torch.float32) or not dy.is_contiguous() or dy.shape != x.shape:
        raise ValueError("layernorm_bwd: dy must be contiguous bf16/fp32 [M, C]")
    if g_bf16 is not None and (g_bf16.dtype != bf16 or not g_bf16.is_contiguous() or g_bf16.shape != x.shape):
        raise ValueError("layernorm_bwd: g_bf16 must be contiguous bf16 [M, C]")
    if L.PROF.enabled:
        L.PROF.meta = {"bytes": M * Cn * (4 + dy.element_size() + 8)}
    L.check(L.load().lkgd_layernorm_bwd(x.data_ptr(), dy.data_ptr(), int(dy.dtype == torch.float32), M, Cn,
                                        gamma.data_ptr(), eps, G.data_ptr(), int(accumulate), _ptr(g_bf16), _stream()),
            "lkgd_layernorm_bwd")
    return G


def geglu_fwd(pre: torch.Tensor) -> torch.Tensor:
    _need_cuda(pre)
    if pre.dtype != bf16 or pre.dim() != 2 or not pre.is_contiguous() or pre.shape[1] % 256:
        raise ValueError("geglu_fwd: pre must be contiguous bf16 [M, 2H], H % 128 == 0")
    M, H2 = pre.shape
    out = torch.empty((M, H2 // 2), device=pre.device, dtype=bf16)
    L.check(L.load().lkgd_geglu_fwd(pre.data_ptr(), out.data_ptr(), M, H2 // 2, _stream()), "lkgd_geglu_fwd")
    return out


def geglu_bwd(pre: torch.Tensor, dout: torch.Tensor) -> torch.Tensor:
    _need_cuda(pre, dout)
    M, H2 = pre.shape
    if pre.dtype != bf16 or not pre.is_contiguous() or dout.dtype != bf16 or not dout.is_contiguous() \
            or dout.shape != (M, H2 // 2):
        raise ValueError("geglu_bwd: pre bf16 [M, 2H], dout bf16 [M, H], contiguous")
    dpre = torch.empty_like(pre)
    L.check(L.load().lkgd_geglu_bwd(pre.data_ptr(), dout.data_ptr(), dpre.data_ptr(), M, H2 // 2, _stream()),
            "lkgd_geglu_bwd")
    return dpre


def colsum_grouped(G: torch.Tensor, n_groups: int, rv: Tuple[int, int, int, int], out: Optional[torch.Tensor] = None):
    """out[g(m)] += G[m] (fp32); ``out`` is created zeroed when not given."""
    _need_cuda(G, out)
    if G.dtype != torch.float32 or G.dim() != 2 or not G.is_contiguous():
        raise ValueError("colsum_grouped: G must be contiguous fp32 [M, C]")
    M, Cn = G.shape
    if out is None:
        out = torch.zeros((n_groups, Cn), device=G.device, dtype=torch.float32)
    elif out.dtype != torch.float32 or out.stride(1) != 1 or out.shape != (n_groups, Cn):
        raise ValueError("colsum_grouped: out must be fp32 [n_groups, C] with unit column stride")
    L.check(L.load().lkgd_colsum_grouped(G.data_ptr(), M, Cn, rv[0], rv[1], rv[2], rv[3], n_groups, out.data_ptr(),
                                         out.stride(0), _stream()), "lkgd_colsum_grouped")
    return out


def downsum2x(x: torch.Tensor, N: int, H: int, W: int) -> torch.Tensor:
    """[N*2H*2W, C] (bf16 / fp32) -> fp32 [N*H*W, C]: backward of the nearest 2x upsample."""
    _need_cuda(x)
    Cn = x.shape[-1]
    if x.dtype not in (bf16, torch.float32) or not x.is_contiguous() or x.numel() != N * 4 * H * W * Cn:
        raise ValueError("downsum2x: contiguous [N*2H*2W, C] expected")
    out = torch.empty((N * H * W, Cn), device=x.device, dtype=torch.float32)
    L.check(L.load().lkgd_downsum2x(x.data_ptr(), int(x.dtype == torch.float32), out.data_ptr(), N, H, W, Cn,
                                    _stream()), "lkgd_downsum2x")
    return out


def zero_stuff2x(x: torch.Tensor, N: int, Hin: int, Win: int) -> torch.Tensor:
    """[N*Ho*Wo, C] -> bf16 [N*Hin*Win, C] with the values at even (h, w) and zeros elsewhere."""
    _need_cuda(x)
    Cn = x.shape[-1]
    Ho, Wo = (Hin - 1) // 2 + 1, (Win - 1) // 2 + 1
    if x.dtype not in (bf16, torch.float32) or not x.is_contiguous() or x.numel() != N * Ho * Wo * Cn:
        raise ValueError("zero_stuff2x: contiguous [N*Ho*Wo, C] expected")
    out = torch.empty((N * Hin * Win, Cn), device=x.device, dtype=bf16)
    L.check(L.load().lkgd_zero_stuff2x(x.data_ptr(), int(x.dtype == torch.float32), out.data_ptr(), N, Hin, Win, Cn,
                                       _stream()), "lkgd_zero_stuff2x")
    return out


def gemm_tn(X: torch.Tensor, Y: torch.Tensor, out: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """out[i, j] += alpha * sum_m X[m, i] Y[m, j];  X / Y bf16 [M, *] (column slices allowed), out fp32."""
    _need_cuda(X, Y, out)
    for t in (X, Y):
        if t.dtype != bf16 or t.dim() != 2 or t.stride(1) != 1:
            raise ValueError("gemm_tn operands must be bf16 2-D row-major (column slices allowed)")
    if X.shape[0] != Y.shape[0] or out.dtype != torch.float32 or out.dim() != 2 or out.stride(1) != 1 \
            or out.shape != (X.shape[1], Y.shape[1]):
        raise ValueError("gemm_tn: out must be fp32 [X.cols, Y.cols]")
    if L.PROF.enabled:
        L.PROF.meta = {"flops": 2.0 * X.shape[0] * X.shape[1] * Y.shape[1]}
    L.check(L.load().lkgd_gemm_tn(X.data_ptr(), X.stride(0), X.shape[1], Y.data_ptr(), Y.stride(0), Y.shape[1],
                                  X.shape[0], alpha, out.data_ptr(), out.stride(0), _stream()), "lkgd_gemm_tn")
    return out


def edm_precondition(latents: torch.Tensor, noise: torch.Tensor, sigma: torch.Tensor, cond: torch.Tensor, Cpad: int):
    """-> (noisy fp32 [B,F,C,H,W], x_in bf16 rows [B*F*H*W, Cpad])."""
    _need_cuda(latents, noise, sigma, cond)
    B, F, Cn, H, W = latents.shape
    for t in (latents, noise, sigma, cond):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("edm_precondition: contiguous fp32 tensors expected")
    if noise.shape != latents.shape or sigma.numel() != B or cond.shape != (B, Cn, H, W):
        raise ValueError("edm_precondition: shape mismatch")
    noisy = torch.empty_like(latents)
    x_in = torch.empty((B * F * H * W, Cpad), device=latents.device, dtype=bf16)
    L.check(L.load().lkgd_edm_precondition(latents.data_ptr(), noise.data_ptr(), sigma.data_ptr(), cond.data_ptr(),
                                           noisy.data_ptr(), x_in.data_ptr(), B, F, Cn, H, W, Cpad, _stream()),
            "lkgd_edm_precondition")
    return noisy, x_in


def edm_loss(pred: torch.Tensor, noisy: torch.Tensor, target: torch.Tensor, sigma: torch.Tensor, Cpad: int,
             grad_scale: float = 1.0, want_grad: bool = True):
    """pred fp32 rows [B*F*H*W, ld] -> (loss: device double scalar tensor, dpred bf16 rows [B*F*H*W, Cpad] | None)."""
    _need_cuda(pred, noisy, target, sigma)
    B, F, Cn, H, W = noisy.shape
    if pred.dtype != torch.float32 or pred.dim() != 2 or pred.stride(1) != 1 or pred.shape[0] != B * F * H * W:
        raise ValueError("edm_loss: pred must be fp32 rows [B*F*H*W, >=C]")
    for t in (noisy, target, sigma):
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise ValueError("edm_loss: contiguous fp32 tensors expected")
    loss = torch.empty((), device=pred.device, dtype=torch.float64)
    dpred = torch.empty((B * F * H * W, Cpad), device=pred.device, dtype=bf16) if want_grad else None
    L.check(L.load().lkgd_edm_loss(pred.data_ptr(), pred.stride(0), noisy.data_ptr(), target.data_ptr(),
                                   sigma.data_ptr(), loss.data_ptr(), _ptr(dpred), B, F, Cn, H, W, Cpad, grad_scale,
                                   _stream()), "lkgd_edm_loss")
    return loss, dpred


def sumsq(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise ValueError("sumsq: contiguous fp32 tensor expected")
    if out is None:
        out = torch.empty((), device=x.device, dtype=torch.float64)
    L.check(L.load().lkgd_sumsq(x.data_ptr(), x.numel(), out.data_ptr(), _stream()), "lkgd_sumsq")
    return out


def adamw(p: torch.Tensor, g: torch.Tensor, m: torch.Tensor, v: torch.Tensor, *, lr: float, beta1: float = 0.9,
          beta2: float = 0.999, eps: float = 1e-8, weight_decay: float = 1e-2, step: int, grad_scale: float = 1.0,
          sumsq_buf: Optional[torch.Tensor] = None, max_norm: float = 0.0):
    _need_cuda(p, g, m, v, sumsq_buf)
    for t in (p, g, m, v):
        if t.dtype != torch.float32 or not t.is_contiguous() or t.numel() != p.numel():
            raise ValueError("adamw: flat contiguous fp32 buffers of equal size expected")
    L.check(L.load().lkgd_adamw(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel(), lr, beta1, beta2,
                                eps, weight_decay, step, grad_scale, _ptr(sumsq_buf), max_norm, _stream()),
            "lkgd_adamw")


def cast2d_bf16(src: torch.Tensor, dst: torch.Tensor, alpha: float = 1.0) -> torch.Tensor:
    """dst (bf16, 2-D, row pitch free) = alpha * src (fp32, 2-D, any strides: transposed views allowed)."""
    _need_cuda(src, dst)
    if src.dtype != torch.float32 or dst.dtype != bf16 or src.dim() != 2 or dst.shape != src.shape \
            or dst.stride(1) != 1:
        raise ValueError("cast2d_bf16: fp32 -> bf16 2-D tensors of equal shape, dst with unit column stride")
    L.check(L.load().lkgd_cast2d_bf16(src.data_ptr(), src.stride(0), src.stride(1), dst.data_ptr(), dst.stride(0),
                                      src.shape[0], src.shape[1], alpha, _stream()), "lkgd_cast2d_bf16")
    return dst


class Cast2dBatch:
    """A fixed table of ``cast2d_bf16`` jobs run as ONE launch (``lkgd_cast2d_bf16_batch``): ``jobs`` is a list of
    ``(src fp32 2-D, dst bf16 2-D, alpha)``; the tensors must stay where they are (the table holds raw pointers)."""

    def __init__(self, jobs):
        if not jobs:
            raise ValueError("Cast2dBatch: no jobs")
        arr = (L.Cast2dJob * len(jobs))()
        self.keep = []
        self.max_elems = 0
        dev = jobs[0][0].device
        for k, (src, dst, alpha) in enumerate(jobs):
            _need_cuda(src, dst)
            if (src.dtype != torch.float32 or dst.dtype != bf16 or src.dim() != 2 or dst.dim() != 2
                    or src.shape != dst.shape or dst.stride(1) != 1 or src.device != dev or dst.device != dev):
                raise ValueError("cast2d_bf16: fp32 -> bf16 2-D tensors of equal shape, dst with unit column stride")
            j = arr[k]
            j.src, j.lds, j.src_cs = src.data_ptr(), src.stride(0), src.stride(1)
            j.dst, j.ldd = dst.data_ptr(), dst.stride(0)
            j.rows, j.cols, j.alpha = src.shape[0], src.shape[1], float(alpha)
            self.max_elems = max(self.max_elems, src.shape[0] * src.shape[1])
            self.keep.append((src, dst))
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        self.table = raw.to(dev)
        self.n = len(jobs)

    def run(self):
        L.check(L.load().lkgd_cast2d_bf16_batch(self.table.data_ptr(), self.n, self.max_elems, _stream()),
                "lkgd_cast2d_bf16_batch")


def small_linear_bwd(dy: torch.Tensor, W: Optional[torch.Tensor] = None, *, x: Optional[torch.Tensor] = None,
                     y: Optional[torch.Tensor] = None, act_out: int = 0, want_dx: bool = True,
                     dx: Optional[torch.Tensor] = None, dW: Optional[torch.Tensor] = None,
                     db: Optional[torch.Tensor] = None) -> Optional[torch.Tensor]:
    """Backward of ``small_linear`` (fp32): returns dx [M, K] (or accumulates into the given ``dx``); ``dW`` [N, K] and
    ``db`` [N] are accumulated in place when given (they need ``x``).  ``act_out`` / ``y``: the forward's output
    activation and output (LeakyReLU only)."""
    _need_cuda(dy, W, x, y, dx, dW, db)
    M, N = dy.shape
    ref = W if W is not None else dW
    K = ref.shape[1]
    for t in (dy, x, y, dx):
        if t is not None and (t.dtype != torch.float32 or t.dim() != 2 or t.stride(1) != 1):
            raise ValueError("small_linear_bwd: fp32 2-D tensors with unit column stride expected")
    for t in (W, dW):
        if t is not None and (t.dtype != torch.float32 or not t.is_contiguous() or t.shape != (N, K)):
            raise ValueError("small_linear_bwd: W / dW must be contiguous fp32 [N, K]")
    if db is not None and (db.dtype != torch.float32 or not db.is_contiguous() or db.numel() != N):
        raise ValueError("small_linear_bwd: db must be contiguous fp32 [N]")
    acc = dx is not None
    if want_dx and dx is None:
        dx = torch.empty((M, K), device=dy.device, dtype=torch.float32)
    if not want_dx:
        dx = None
    L.check(L.load().lkgd_small_linear_bwd(dy.data_ptr(), dy.stride(0), _ptr(y), y.stride(0) if y is not None else 0,
                                           act_out, _ptr(x), x.stride(0) if x is not None else 0, _ptr(W), _ptr(dx),
                                           dx.stride(0) if dx is not None else 0, int(acc), _ptr(dW), _ptr(db), M, N, K,
                                           _stream()), "lkgd_small_linear_bwd")
    return dx


def polar_bwd(a: torch.Tensor, b: torch.Tensor, d0: torch.Tensor, d1: torch.Tensor, mode: int):
    _need_cuda(a, b, d0, d1)
    a, b, d0, d1 = (t.contiguous() for t in (a, b, d0, d1))
    o0, o1 = torch.empty_like(a), torch.empty_like(a)
    L.check(L.load().lkgd_polar_bwd(a.data_ptr(), b.data_ptr(), d0.data_ptr(), d1.data_ptr(), o0.data_ptr(),
                                    o1.data_ptr(), a.numel(), mode, _stream()), "lkgd_polar_bwd")
    return o0, o1


def grouped1x1_bwd_w(dy: torch.Tensor, x: torch.Tensor, dw: torch.Tensor):
    """dw [G, 4] (contiguous fp32, +=) from dy [B, G] and x [B, 4G] (row pitches free)."""
    _need_cuda(dy, x, dw)
    B, G = dy.shape
    if x.shape != (B, 4 * G) or dw.numel() != 4 * G or not dw.is_contiguous() or dy.stride(1) != 1 or x.stride(1) != 1:
        raise ValueError("grouped1x1_bwd_w: dy [B, G], x [B, 4G], dw [G, 4]")
    L.check(L.load().lkgd_grouped1x1_bwd_w(dy.data_ptr(), dy.stride(0), x.data_ptr(), x.stride(0), dw.data_ptr(), B, G,
                                           _stream()), "lkgd_grouped1x1_bwd_w")


def hamilton_bwd(dWt: torch.Tensor, dr: torch.Tensor, di: torch.Tensor, dj: torch.Tensor, dk: torch.Tensor):
    _need_cuda(dWt, dr, di, dj, dk)
    out_f, in_f = dWt.shape
    for t in (dr, di, dj, dk):
        if not t.is_contiguous() or t.numel() != (in_f // 4) * (out_f // 4) or t.dtype != torch.float32:
            raise ValueError("hamilton_bwd: component gradients must be contiguous fp32 [in/4, out/4]")
    if not dWt.is_contiguous():
        raise ValueError("hamilton_bwd: dWt must be contiguous [out, in]")
    L.check(L.load().lkgd_hamilton_bwd(dWt.data_ptr(), in_f, out_f, dr.data_ptr(), di.data_ptr(), dj.data_ptr(),
                                       dk.data_ptr(), _stream()), "lkgd_hamilton_bwd")
