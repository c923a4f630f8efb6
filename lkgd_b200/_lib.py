"""ctypes binding of liblkgd_b200.so (the C ABI declared in include/lkgd_b200.h).

There is NO fallback: if the shared library is missing or a call fails, this raises.  The library is built
in-tree by ``lkgd_b200/build.py`` (``__graft_entry__.build()``)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

ABI_VERSION = 7
LIB_PATH = Path(__file__).resolve().parent / "lib" / "liblkgd_b200.so"

A_LINEAR, A_CONV3X3, A_TCONV3 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_GELU, ACT_QUICK_GELU = 0, 1, 2, 3, 4
RV_NONE, RV_FRAME, RV_FRAMEPOS, RV_BATCH, RV_TCTX_0272, RV_BATCH_TCTX = 0, 1, 2, 3, 4, 5
SL_NONE, SL_SILU, SL_LEAKY = 0, 1, 3

i32, f32, vp, i64, sz = C.c_int32, C.c_float, C.c_void_p, C.c_int64, C.c_size_t


class Cast2dJob(C.Structure):
    """Mirror of ``lkgd_cast2d_job``."""
    _fields_ = [("src", vp), ("lds", i64), ("src_cs", i64), ("dst", vp), ("ldd", i64), ("rows", i32), ("cols", i32),
                ("alpha", f32), ("reserved", i32)]


class GemmArgs(C.Structure):
    """Mirror of ``lkgd_gemm_args``."""
    _fields_ = [
        ("a_mode", i32), ("M", i32), ("N", i32), ("K0", i32), ("K1", i32),
        ("A", vp), ("lda", i32), ("A1", vp), ("lda1", i32),
        ("Bw", vp), ("ldb", i32), ("Bw1", vp), ("ldb1", i32),
        ("NIMG", i32), ("Hin", i32), ("Win", i32), ("stride", i32), ("F", i32), ("HW", i32),
        ("bias", vp), ("rowvec", vp), ("rv_mode", i32), ("rv_HW", i32), ("rv_F", i32), ("rv_B", i32),
        ("act", i32), ("s0", f32), ("res1", vp), ("ldr1", i32), ("s1", f32),
        ("res2", vp), ("ldr2", i32), ("s2", f32),
        ("out", vp), ("ldo", i32), ("out_f32", i32), ("n_store", i32), ("res1_f32", i32), ("res2_f32", i32),
        ("rv_ld", i32), ("gn_stats", vp), ("gn_rows", i32), ("out2", vp), ("ldo2", i32), ("pad_br", i32),
    ]


# name -> (restype, argtypes); must list every symbol include/lkgd_b200.h declares
SIGNATURES = {
    "lkgd_abi_version": (i32, []),
    "lkgd_strerror": (C.c_char_p, [i32]),
    "lkgd_last_cuda_error": (C.c_char_p, []),
    "lkgd_device_check": (i32, [i32]),
    "lkgd_launch_count": (C.c_uint64, []),
    "lkgd_gemm": (i32, [C.POINTER(GemmArgs), vp]),
    "lkgd_gemm_simt_check": (i32, [C.POINTER(GemmArgs), vp]),
    "lkgd_groupnorm_workspace": (sz, [i32, i32]),
    "lkgd_groupnorm": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, f32, i32, i32, vp, vp, sz, vp]),
    "lkgd_groupnorm_from_stats": (i32, [vp, i32, vp, vp, i32, vp, i32, i32, i32, i32, vp, vp, f32, i32, i32, vp, vp, vp,
                                        sz, vp]),
    "lkgd_layernorm": (i32, [vp, i32, i32, vp, vp, f32, vp, i32, i32, i32, i32, i32, i32, vp, vp, vp]),
    "lkgd_attention": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "lkgd_attention_simt_check": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "lkgd_attention_temporal": (i32, [vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "lkgd_small_linear": (i32, [vp, i32, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "lkgd_timestep_embedding": (i32, [vp, i32, i32, vp, vp]),
    "lkgd_axpy_f32": (i32, [vp, f32, vp, i64, vp]),
    "lkgd_scale_f32": (i32, [vp, f32, vp, i64, vp]),
    "lkgd_polar": (i32, [vp, vp, vp, vp, i32, i32, vp]),
    "lkgd_pack_input": (i32, [vp, i32, i32, f32, vp, vp, i32, i32, vp, i32, i32, i32, i32, i32, vp]),
    "lkgd_unpack_output": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "lkgd_nchw_to_nhwc": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "lkgd_nhwc_to_nchw": (i32, [vp, vp, i32, i32, i32, i32, vp]),
    "lkgd_upsample2x": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "lkgd_cast_bf16": (i32, [vp, vp, i64, vp]),
    "lkgd_concat_channels": (i32, [vp, i32, vp, i32, i32, vp, i64, vp]),
    "lkgd_axpby": (i32, [vp, i32, f32, vp, i32, f32, i64, vp]),
    "lkgd_cond_conv_in": (i32, [vp, i32, i32, i32, i32, vp, vp, vp, vp]),
    "lkgd_thin_conv3x3": (i32, [vp, i32, i32, i32, i32, i32, vp, vp, i32, vp, vp]),
    "lkgd_patchify": (i32, [vp, i32, i32, i32, i32, i32, vp, i32, vp]),
    "lkgd_softmax_rows": (i32, [vp, i64, i64, i32, f32, vp, i64, vp]),
    "lkgd_time_conv_out": (i32, [vp, i32, vp, vp, vp, i32, i32, i64, i32, vp]),
    "lkgd_select_rows": (i32, [vp, i32, vp, i64, i32, i32, i32, i32, i32, vp]),
    "lkgd_cfg_euler_step": (i32, [vp, i32, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp, vp]),
    "lkgd_cfg_euler_step_pair": (i32, [vp, vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp, vp]),
    "lkgd_fusion_euler_step": (i32, [vp, vp, vp, vp, i32, i32, i32, i32, i32, f32, f32, vp]),
    # ---- training step
    "lkgd_attention_lse": (i32, [vp, i32, vp, i32, vp, i32, vp, i32, i32, i32, i32, i32, i32, f32, vp, vp]),
    "lkgd_attention_bwd_workspace": (sz, [i32, i32, i32]),
    "lkgd_attention_bwd": (i32, [vp, i32, vp, i32, vp, i32, vp, vp, i32, vp, vp, i32, vp, i32, vp, i32, i32, i32, i32,
                                 i32, f32, vp, sz, vp]),
    "lkgd_attention_temporal_bwd": (i32, [vp, vp, vp, i32, i32, i32, i32, i32, f32, vp]),
    "lkgd_groupnorm_bwd_workspace": (sz, [i32, i32]),
    "lkgd_groupnorm_bwd": (i32, [vp, i32, vp, i32, i32, i32, i32, vp, vp, f32, i32, i32, vp, vp, vp, i32, vp, i32, vp,
                                 i32, vp, vp, sz, vp]),
    "lkgd_layernorm_bwd": (i32, [vp, vp, i32, i32, i32, vp, f32, vp, i32, vp, vp]),
    "lkgd_geglu_fwd": (i32, [vp, vp, i64, i32, vp]),
    "lkgd_geglu_bwd": (i32, [vp, vp, vp, i64, i32, vp]),
    "lkgd_colsum_grouped": (i32, [vp, i64, i32, i32, i32, i32, i32, i32, vp, i64, vp]),
    "lkgd_downsum2x": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "lkgd_zero_stuff2x": (i32, [vp, i32, vp, i32, i32, i32, i32, vp]),
    "lkgd_gemm_tn": (i32, [vp, i64, i32, vp, i64, i32, i64, f32, vp, i64, vp]),
    "lkgd_edm_precondition": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, vp]),
    "lkgd_edm_loss": (i32, [vp, i32, vp, vp, vp, vp, vp, i32, i32, i32, i32, i32, i32, f32, vp]),
    "lkgd_sumsq": (i32, [vp, i64, vp, vp]),
    "lkgd_adamw": (i32, [vp, vp, vp, vp, i64, f32, f32, f32, f32, f32, i32, f32, vp, f32, vp]),
    "lkgd_cast2d_bf16": (i32, [vp, i64, i64, vp, i64, i32, i32, f32, vp]),
    "lkgd_cast2d_bf16_batch": (i32, [vp, i32, i32, vp]),
    "lkgd_small_linear_bwd": (i32, [vp, i32, vp, i32, i32, vp, i32, vp, vp, i32, i32, vp, vp, i32, i32, i32, vp]),
    "lkgd_polar_bwd": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, vp]),
    "lkgd_grouped1x1_bwd_w": (i32, [vp, i32, vp, i32, vp, i32, i32, vp]),
    "lkgd_hamilton_bwd": (i32, [vp, i32, i32, vp, vp, vp, vp, vp]),
}

_lib = None


class Profiler:
    """Optional per-call CUDA-event timing of the C-ABI launches (bench.py's roofline section).  Off by default;
    when off the wrapper costs one attribute test per call."""
    enabled = False
    records = []      # (symbol, start_event, end_event, meta dict)
    meta = None       # set by ops.* just before a call (e.g. {"flops": ..})


PROF = Profiler()
_TIMED = {"lkgd_gemm", "lkgd_groupnorm", "lkgd_groupnorm_from_stats", "lkgd_layernorm", "lkgd_attention", "lkgd_attention_temporal",
          "lkgd_small_linear", "lkgd_timestep_embedding", "lkgd_pack_input", "lkgd_unpack_output",
          "lkgd_upsample2x", "lkgd_concat_channels", "lkgd_cast_bf16", "lkgd_axpby", "lkgd_cfg_euler_step", "lkgd_cfg_euler_step_pair", "lkgd_fusion_euler_step", "lkgd_axpy_f32",
          "lkgd_nchw_to_nhwc", "lkgd_nhwc_to_nchw", "lkgd_polar", "lkgd_scale_f32",
          "lkgd_cond_conv_in", "lkgd_thin_conv3x3", "lkgd_select_rows", "lkgd_attention_lse", "lkgd_attention_bwd", "lkgd_attention_temporal_bwd", "lkgd_groupnorm_bwd",
          "lkgd_layernorm_bwd", "lkgd_geglu_fwd", "lkgd_geglu_bwd", "lkgd_colsum_grouped", "lkgd_downsum2x",
          "lkgd_zero_stuff2x", "lkgd_gemm_tn", "lkgd_edm_precondition", "lkgd_edm_loss", "lkgd_sumsq", "lkgd_adamw",
          "lkgd_cast2d_bf16", "lkgd_cast2d_bf16_batch"}


def _timed(name, fn):
    def call(*args):
        if not PROF.enabled:
            return fn(*args)
        import torch
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        rc = fn(*args)
        e1.record()
        PROF.records.append((name, e0, e1, PROF.meta))
        PROF.meta = None
        return rc
    return call


class LkgdError(RuntimeError):
    pass


def load() -> C.CDLL:
    """Loads the shared library and binds every symbol.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise LkgdError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(lkgd_b200 has no CPU or PyTorch fallback)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the symbol is not exported
        fn.restype, fn.argtypes = res, args
    if lib.lkgd_abi_version() != ABI_VERSION:
        raise LkgdError(f"ABI version mismatch: library reports {lib.lkgd_abi_version()}, binding expects {ABI_VERSION}")
    for name in _TIMED:
        setattr(lib, name, _timed(name, getattr(lib, name)))
    _lib = lib
    return lib


def check(rc: int, what: str):
    """Maps C error codes to exceptions; shape / alignment problems are ValueError like the reference's
    configuration errors (models/unet_spatio_temporal_condition_controlnet.py:101-124)."""
    if rc == 0:
        return
    lib = load()
    msg = lib.lkgd_strerror(rc).decode()
    if rc in (-1, -2):
        raise ValueError(f"{what}: {msg}")
    if rc == -5:
        msg += ": " + lib.lkgd_last_cuda_error().decode()
    raise LkgdError(f"{what}: {msg}")
