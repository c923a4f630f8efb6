"""Isolated timing of lkgd_attention_bwd (prep + dQ + dK,dV kernels) on the C5 training shapes, L2 flushed between calls.
usage: python tools/bench_attn_bwd.py        (LKGD_ATTN_BWD_MMA=1: the mma.sync kernels)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops  # noqa: E402

dev = "cuda"
flush = torch.empty(256 << 20, device=dev, dtype=torch.uint8)
for name, n_img, heads, d, N in (("C5 L0", 14, 5, 64, 2560), ("C5 L1", 14, 10, 64, 640), ("C5 L2", 14, 20, 64, 160),
                                 ("C3-size L0", 25, 5, 64, 9216)):
    C = heads * d
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(n_img * N, 3 * C, device=dev, dtype=torch.bfloat16, generator=g)
    q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    o, lse = ops.attention(q, k, v, n_img=n_img, heads=heads, d=d, Nq=N, Nk=N, return_lse=True)
    dO = torch.randn(n_img * N, C, device=dev, dtype=torch.bfloat16, generator=g)
    dqkv = torch.zeros_like(qkv)
    res = {}
    for legacy in (False, True):
        if legacy:
            os.environ["LKGD_ATTN_BWD_MMA"] = "1"
        else:
            os.environ.pop("LKGD_ATTN_BWD_MMA", None)
        ts = []
        for it in range(7):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.attention_bwd(q, k, v, o, dO, lse, dqkv[:, :C], dqkv[:, C:2 * C], dqkv[:, 2 * C:], n_img=n_img, heads=heads, d=d, N=N)
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        res[legacy] = sorted(ts[2:])[len(ts[2:]) // 2]
    fl = 14.0 * n_img * heads * N * N * d
    print(f"{name:12s} n_img {n_img} heads {heads} N {N}: tcgen05 {res[False]:.3f} ms = {fl / res[False] / 1e9:.0f} TF/s | "
          f"mma.sync {res[True]:.3f} ms = {fl / res[True] / 1e9:.0f} TF/s")
os.environ.pop("LKGD_ATTN_BWD_MMA", None)
