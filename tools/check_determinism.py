#!/usr/bin/env python
"""Run-to-run determinism of the kernels: every op is launched N times on the same inputs and compared bit for bit."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops
from lkgd_b200.ops import A_CONV3X3, A_LINEAR, A_TCONV3, ACT_GEGLU
bf16 = torch.bfloat16
dev = "cuda"
N = 20


def check(name, f):
    ref = f().clone()
    bad = 0
    for _ in range(N):
        out = f()
        torch.cuda.synchronize()
        if not torch.equal(out, ref):
            bad += 1
    print(f"{name:60s} {'DETERMINISTIC' if bad == 0 else f'{bad}/{N} runs differ'}", flush=True)


g = torch.Generator(device=dev).manual_seed(0)
for n_img, heads, d, Nq in [(16, 2, 16, 256), (16, 4, 16, 64), (10, 5, 64, 2304), (4, 20, 64, 144), (2, 5, 64, 9216)]:
    C = heads * d
    qkv = torch.randn(n_img * Nq, 3 * C, device=dev, dtype=bf16, generator=g)
    check(f"attention n_img={n_img} heads={heads} d={d} N={Nq}",
          lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_img=n_img, heads=heads, d=d, Nq=Nq, Nk=Nq))
for M, Nn, K, kw in [(4096, 32, 32, {}), (4096, 64, 64, {}), (115200, 640, 640, {}), (28800, 1280, 1280, {}),
                     (460800, 320, 320, {})]:
    A = torch.randn(M, K, device=dev, dtype=bf16, generator=g)
    W = torch.randn(Nn, K, device=dev, dtype=bf16, generator=g) * K ** -0.5
    b = torch.randn(Nn, device=dev, generator=g)
    r1 = torch.randn(M, Nn, device=dev, generator=g)
    r2 = torch.randn(M, Nn, device=dev, generator=g)
    check(f"gemm {M}x{Nn}x{K} f32", lambda: ops.gemm(A, W, bias=b, out_f32=True))
    check(f"gemm {M}x{Nn}x{K} res32", lambda: ops.gemm(A, W, bias=b, res1=r1, out_f32=True))
    check(f"gemm {M}x{Nn}x{K} 2 x res32 -> bf16", lambda: ops.gemm(A, W, bias=b, res1=r1, res2=r2, s1=0.3, s2=0.7, s0=0.3))
    check(f"gemm {M}x{Nn}x{K} bf16", lambda: ops.gemm(A, W, bias=b))
    if Nn % 256 == 0:
        Wg, bg = ops.pack_geglu(W, b)
        check(f"gemm {M}x{Nn}x{K} geglu", lambda: ops.gemm(A, Wg, bias=bg, act=ACT_GEGLU))
for NS, R, C in [(16, 256, 32), (2, 2048, 32), (50, 2304, 640)]:
    x = torch.randn(NS * R, C, device=dev, generator=g)
    gm, bt = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    check(f"groupnorm NS={NS} R={R} C={C}", lambda: ops.groupnorm(x, gm, bt, 1e-5, NS=NS, R=R, silu=True))
    check(f"layernorm M={NS * R} C={C}", lambda: ops.layernorm(x, gm, bt, 1e-5))
