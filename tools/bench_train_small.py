#!/usr/bin/env python
"""One launch each of the training step's bandwidth-bound kernels at the C5 level-0 shapes (for ncu):
GroupNorm backward (both passes), cross-vector backward (small_linear_bwd over the 56 MB concatenated matrix),
column sums of the cross-vector gradient, GEGLU forward / backward, LayerNorm backward."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lkgd_b200 import ops

dev = "cuda"
M, C = 14 * 2560, 320
x = torch.randn(M, C, device=dev)
g, b = torch.ones(C, device=dev), torch.zeros(C, device=dev)
dy = torch.randn(M, C, device=dev, dtype=torch.bfloat16)
G = torch.zeros(M, C, device=dev)
for _ in range(2):
    _, st = ops.groupnorm(x, g, b, 1e-5, NS=14, R=2560, silu=True, return_stats=True)
    ops.groupnorm_bwd(x, dy, st, g, b, 1e-5, NS=14, R=2560, silu=True, out1=G, acc1=True)
    W = torch.randn(13760, 1024, device=dev) * 0.02
    d = torch.randn(1, 13760, device=dev)
    ops.small_linear_bwd(d, W)
    ops.colsum_grouped(G, 1, (ops.RV_BATCH, M, 1, 1))
    pre = torch.randn(M, 2560, device=dev, dtype=torch.bfloat16)
    act = ops.geglu_fwd(pre)
    ops.geglu_bwd(pre, torch.randn_like(act))
    ops.layernorm_bwd(x, dy, g, 1e-5, G, accumulate=True)
    torch.cuda.synchronize()
print("ok")
