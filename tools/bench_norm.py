#!/usr/bin/env python
"""Micro-benchmark of the GroupNorm / LayerNorm kernels on the C3 step's shapes (algorithmic GB/s, L2 flushed)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)


AB = os.environ.get("LKGD_BENCH_AB")     # name of a library switch to A/B in-process (interleaved calls)


def timeit(f, n=5):
    f(); f()
    ts, tb = [], []
    for i in range(n * (2 if AB else 1)):
        alt = AB is not None and (i & 1)
        if AB:
            if alt:
                os.environ[AB] = "1"
            else:
                os.environ.pop(AB, None)
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        (tb if alt else ts).append(e0.elapsed_time(e1))
    if AB:
        os.environ.pop(AB, None)
        print(f"    A/B {AB}: unset {sorted(ts)[n // 2]:.4f} ms, set {sorted(tb)[n // 2]:.4f} ms", flush=True)
    return sorted(ts)[n // 2]


for name, M, C in [("L0", 460800, 320), ("L1", 115200, 640), ("L2", 28800, 1280)]:
    x = torch.randn(M, C, device="cuda")
    g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    out = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-5, out=out))
    print(json.dumps(dict(op="layernorm", level=name, ms=round(ms, 4), gbs=round(M * C * 6 / ms / 1e6))), flush=True)
    av = torch.randn(2, C, device="cuda")
    so = torch.empty_like(x)
    ms = timeit(lambda: ops.layernorm(x, g, b, 1e-5, addvec=av, rv=(ops.RV_BATCH, M // 50, 25, 2), sum_out=so, out=out))
    print(json.dumps(dict(op="layernorm+add+sum", level=name, ms=round(ms, 4), gbs=round(M * C * 10 / ms / 1e6))), flush=True)
    for NS in (50, 2):
        ms = timeit(lambda: ops.groupnorm(x, g, b, 1e-5, NS=NS, R=M // NS, silu=True, out=out))
        print(json.dumps(dict(op=f"groupnorm NS={NS}", level=name, ms=round(ms, 4), gbs=round(M * C * 10 / ms / 1e6))), flush=True)
