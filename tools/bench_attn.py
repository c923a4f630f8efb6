#!/usr/bin/env python
"""Micro-benchmark of lkgd_attention on the spatial self-attention shapes of the C3 step (fused-qkv operands)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops
SHAPES = [("L0", 50, 5, 9216), ("L1", 50, 10, 2304), ("L2", 50, 20, 576), ("L3", 50, 20, 144)]
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
only = [int(i) for i in sys.argv[1].split(",")] if len(sys.argv) > 1 else range(len(SHAPES))
for i in only:
    name, n_img, heads, N = SHAPES[i]
    C = heads * 64
    qkv = torch.randn(n_img * N, 3 * C, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(n_img * N, C, device="cuda", dtype=torch.bfloat16)
    v = qkv[:, 2 * C:]
    f = lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], v, n_img=n_img, heads=heads, d=64, Nq=N, Nk=N, out=out)
    f(); f()
    # LKGD_ATTN_POLY_AB=0,2,3,4 / LKGD_ATTN_QT_AB=0,1: time the variants interleaved in one process (the library reads the
    # switches per call)
    var_name = "LKGD_ATTN_QT" if os.environ.get("LKGD_ATTN_QT_AB") else "LKGD_ATTN_POLY"
    ab = os.environ.get("LKGD_ATTN_QT_AB") or os.environ.get("LKGD_ATTN_POLY_AB")
    variants = ab.split(",") if ab else [None]
    ts = {pv: [] for pv in variants}
    for _ in range(5):
        for pv in variants:
            if pv is not None:
                os.environ[var_name] = pv
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record(); torch.cuda.synchronize()
            ts[pv].append(e0.elapsed_time(e1))
    for pv in variants:
        ms = sorted(ts[pv])[2]
        print(json.dumps(dict(name=name, poly=pv, n_img=n_img, heads=heads, N=N, ms=round(ms, 4),
                              tflops=round(4.0 * n_img * heads * N * N * 64 / ms / 1e9, 1))), flush=True)
    os.environ.pop(var_name, None)
