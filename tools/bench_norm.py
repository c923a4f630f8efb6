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


for name, M, C in ([] if "train" in sys.argv else [("L0", 460800, 320), ("L1", 115200, 640), ("L2", 28800, 1280)]):
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

if "train" in sys.argv:   # GroupNorm backward on the C5 training clip's shapes (14 frames, 40x64 latents)
    for name, F_, HW, C in [("L0", 14, 2560, 320), ("L1", 14, 640, 640), ("L2", 14, 160, 1280), ("L3", 14, 40, 1280)]:
        for NS, R in ((F_, HW), (1, F_ * HW)):
            M = NS * R
            x = torch.randn(M, C, device="cuda")
            g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
            dy = torch.randn(M, C, device="cuda", dtype=torch.bfloat16)
            G = torch.zeros(M, C, device="cuda")
            ob = torch.empty(M, C, device="cuda", dtype=torch.bfloat16)
            _, st = ops.groupnorm(x, g, b, 1e-5, NS=NS, R=R, silu=True, return_stats=True)
            ms = timeit(lambda: ops.groupnorm_bwd(x, dy, st, g, b, 1e-5, NS=NS, R=R, silu=True, out1=G, acc1=True))
            print(json.dumps(dict(op=f"groupnorm_bwd NS={NS} acc", level=name, ms=round(ms, 4),
                                  gbs=round(M * C * 20 / ms / 1e6))), flush=True)
            ms = timeit(lambda: ops.groupnorm_bwd(x, dy, st, g, b, 1e-5, NS=NS, R=R, silu=True, out_bf16=ob))
            print(json.dumps(dict(op=f"groupnorm_bwd NS={NS} bf16", level=name, ms=round(ms, 4),
                                  gbs=round(M * C * 14 / ms / 1e6))), flush=True)
    sys.exit(0)
