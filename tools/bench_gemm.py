#!/usr/bin/env python
"""Micro-benchmark of lkgd_gemm on the shapes that dominate the C3 denoise step (from profiles/*prof*.json).
Usage: python tools/bench_gemm.py [--only i,j] [--iters N]   (CUDA events, L2 flushed between calls)"""
import argparse, json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops
from lkgd_b200.ops import A_CONV3X3, A_LINEAR, A_TCONV3, ACT_GEGLU

bf16 = torch.bfloat16
# (name, mode, M or geometry, N, K0, epilogue)
SHAPES = [
    ("ff1_L0_geglu", "lin", 460800, 2560, 320, "geglu"),
    ("ff1_L1_geglu", "lin", 115200, 5120, 640, "geglu"),
    ("ff1_L2_geglu", "lin", 28800, 10240, 1280, "geglu"),
    ("proj_L0_res32", "lin", 460800, 320, 320, "res32"),
    ("qkv_L0_bf16", "lin", 460800, 960, 320, "bf16"),
    ("ff2_L0_res32", "lin", 460800, 320, 1280, "res32"),
    ("ff2_L1_res32", "lin", 115200, 640, 2560, "res32"),
    ("proj_L1_res32", "lin", 115200, 640, 640, "res32"),
    ("proj_L2_res32", "lin", 28800, 1280, 1280, "res32"),
    ("ff2_L2_res32", "lin", 28800, 1280, 5120, "res32"),
    ("conv_L0_f32", "conv", (50, 72, 128), 320, 320, "f32"),
    ("conv_L0_res32", "conv", (50, 72, 128), 320, 320, "res32"),
    ("conv_L1_f32", "conv", (50, 36, 64), 640, 640, "f32"),
    ("conv_L2_f32", "conv", (50, 18, 32), 1280, 1280, "f32"),
    ("conv_L3_f32", "conv", (50, 9, 16), 1280, 1280, "f32"),
    ("tconv_L0_f32", "tconv", (2, 25, 9216), 320, 320, "f32"),
    ("tconv_L1_res32", "tconv", (2, 25, 2304), 640, 640, "res32"),
    ("tconv_L3_f32", "tconv", (2, 25, 144), 1280, 1280, "f32"),
    ("tconv_L0_res32", "tconv", (2, 25, 9216), 320, 320, "res32"),
    ("proj_L3_res32", "lin", 7200, 1280, 1280, "res32"),
    # ControlNet condition encoder at pixel resolution (C4, 28 frames of 576x1024): thin layers
    ("cond1_k8_silu", "conv", (28, 576, 1024), 16, 8, "silu"),
    ("cond1_k64_silu", "conv", (28, 576, 1024), 16, 64, "silu"),
    ("cond2_k16_silu", "conv", (28, 576, 1024), 16, 16, "silu"),
    # VAE temporal decoder, 8-frame chunk of 576x1024 frames (lkgd_b200/vae.py): top level, 128 channels   [23..]
    ("vae_conv128_res32gn", "conv", (8, 576, 1024), 128, 128, "res32gn"),
    ("vae_conv128_res32", "conv", (8, 576, 1024), 128, 128, "res32"),
    ("vae_conv128_f32gn", "conv", (8, 576, 1024), 128, 128, "f32gn"),
    ("vae_conv128_f32", "conv", (8, 576, 1024), 128, 128, "f32"),
    ("vae_conv128_bf16", "conv", (8, 576, 1024), 128, 128, "bf16"),
    ("vae_tconv128_res32gn", "tconv", (1, 8, 589824), 128, 128, "res32gn"),
    ("vae_tconv128_f32gn", "tconv", (1, 8, 589824), 128, 128, "f32gn"),
    ("vae_tconv128_f32", "tconv", (1, 8, 589824), 128, 128, "f32"),
    ("vae_tconv128_bf16", "tconv", (1, 8, 589824), 128, 128, "bf16"),
    ("vae_tconv256_res32gn", "tconv", (1, 8, 147456), 256, 256, "res32gn"),
    ("vae_conv256_res32gn", "conv", (8, 288, 512), 256, 256, "res32gn"),
    # training shapes (C5: 14 frames, 40x64 latents): few 128-row tiles at the lower levels   [34..]
    ("train_conv_L3", "conv", (14, 5, 8), 1280, 1280, "f32"),
    ("train_conv_L2", "conv", (14, 10, 16), 1280, 1280, "f32"),
    ("train_conv_L1", "conv", (14, 20, 32), 640, 640, "f32"),
    ("train_tconv_L3", "tconv", (1, 14, 40), 1280, 1280, "f32"),
    ("train_tconv_L2", "tconv", (1, 14, 160), 1280, 1280, "f32"),
    ("train_lin_L3", "lin", 560, 1280, 1280, "res32"),
    ("train_lin_L2", "lin", 2240, 1280, 1280, "res32"),
    ("train_lin_L2_ff2", "lin", 2240, 1280, 5120, "res32"),
]


def run(shape, iters, flush, ab=None):
    name, mode, geo, N, K0, epi = shape
    dev = "cuda"
    g = torch.Generator(device=dev).manual_seed(0)
    if mode == "lin":
        M = geo
        A = torch.randn(M, K0, device=dev, dtype=bf16, generator=g)
        taps, kw = 1, dict(mode=A_LINEAR)
    elif mode == "conv":
        n, h, w = geo
        M = n * h * w
        A = torch.randn(M, K0, device=dev, dtype=bf16, generator=g)
        taps, kw = 9, dict(mode=A_CONV3X3, conv=(n, h, w, 1))
    else:
        b, f, hw = geo
        M = b * f * hw
        A = torch.randn(M, K0, device=dev, dtype=bf16, generator=g)
        taps, kw = 3, dict(mode=A_TCONV3, tconv=(b, f, hw))
    W = (torch.randn(N, taps * K0, device=dev, dtype=bf16, generator=g) * (taps * K0) ** -0.5)
    bias = torch.randn(N, device=dev, generator=g)
    n_out = N // 2 if epi == "geglu" else N
    gn = epi.endswith("gn")
    if gn:
        epi = epi[:-2]
        kw.update(gn_rows=geo[1] * geo[2] if mode == "conv" else geo[2])
    out_b = 4 if epi in ("f32", "res32") else 2
    bytes_ = M * K0 * 2 + N * taps * K0 * 2 + M * n_out * out_b
    if epi == "geglu":
        W, bias = ops.pack_geglu(W, bias)
        kw.update(act=ACT_GEGLU)
    elif epi == "res32":
        kw.update(res1=torch.randn(M, N, device=dev, generator=g), out_f32=True)
        bytes_ += M * N * 4
    elif epi == "f32":
        kw.update(out_f32=True)
    elif epi == "silu":
        kw.update(act=1)
    out = torch.empty(M, n_out, device=dev, dtype=torch.float32 if out_b == 4 else bf16)
    for _ in range(2):
        ops.gemm(A, W, bias=bias, out=out, **kw)
    ts, ts_b = [], []
    for it in range(iters * (2 if ab else 1)):
        # --ab NAME: interleave calls with the environment switch NAME unset / set (the library reads it per call), so
        # both variants see the same clocks and power state
        alt = ab is not None and (it & 1)
        if ab:
            if alt:
                os.environ[ab] = "1"
            else:
                os.environ.pop(ab, None)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.gemm(A, W, bias=bias, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        (ts_b if alt else ts).append(e0.elapsed_time(e1))
    if ab:
        os.environ.pop(ab, None)
    ms = sorted(ts)[len(ts) // 2]
    flops = 2.0 * M * N * taps * K0
    extra = {}
    if ab:
        mb = sorted(ts_b)[len(ts_b) // 2]
        extra = {"ms_" + ab: round(mb, 4), "ratio": round(mb / ms, 3)}
    return dict(**extra, name=name, M=M, N=N, K=taps * K0, epi=epi + ("+gn" if gn else ""), ms=round(ms, 4), tflops=round(flops / ms / 1e9, 1),
                gbs=round(bytes_ / ms / 1e6, 1))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--ab", default=None, help="environment switch to A/B in-process (interleaved calls)")
    a = ap.parse_args()
    sel = [int(i) for i in a.only.split(",")] if a.only else range(len(SHAPES))
    flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
    res = []
    for i in sel:
        r = run(SHAPES[i], a.iters, flush, a.ab)
        res.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
