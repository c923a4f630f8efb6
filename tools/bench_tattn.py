#!/usr/bin/env python
"""Micro-benchmark of lkgd_attention_temporal on the C3 shapes (HBM-bound: bytes = qkv read + out write)."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops
flush = torch.empty(256 << 20, device="cuda", dtype=torch.uint8)
for name, HW, heads in [("L0", 9216, 5), ("L1", 2304, 10), ("L2", 576, 20), ("L3", 144, 20)]:
    B, F, C = 2, 25, heads * 64
    qkv = torch.randn(B * F * HW, 3 * C, device="cuda", dtype=torch.bfloat16)
    out = torch.empty(B * F * HW, C, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attention_temporal(qkv, B=B, F=F, HW=HW, heads=heads, d=64, out=out)
    f(); f()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); f(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[2]
    print(json.dumps(dict(name=name, ms=round(ms, 4), gbs=round(B * F * HW * C * 2 * 4 / ms / 1e6, 1))), flush=True)
