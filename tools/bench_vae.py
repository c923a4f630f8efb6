"""Per-kernel breakdown of the VAE's temporal decode at the headline size (8-frame chunk of 576x1024 frames, SVD widths,
random weights): CUDA events around every C-ABI call (lkgd_b200._lib.PROF), GEMM launches listed by shape.
usage: python tools/bench_vae.py [frames=8] [out.json]
       ncu --profile-from-start off ... python tools/bench_vae.py 8 --profiler     (ONE decode between cudaProfilerStart / Stop)"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from lkgd_b200 import _lib  # noqa: E402
from lkgd_b200.flops import vae_flops  # noqa: E402
from lkgd_b200.vae import SVD_VAE_CONFIG, AutoencoderKLTemporalDecoder  # noqa: E402

PROFILER = "--profiler" in sys.argv
if PROFILER:
    sys.argv.remove("--profiler")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
vae = AutoencoderKLTemporalDecoder(**SVD_VAE_CONFIG).to(dev)
z = torch.randn(n, 4, 72, 128, device=dev)
for _ in range(2):
    vae.decode(z, num_frames=n)
torch.cuda.synchronize()
if PROFILER:
    torch.cuda.cudart().cudaProfilerStart()
    vae.decode(z, num_frames=n)
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    sys.exit(0)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(3):
    vae.decode(z, num_frames=n)
b.record()
torch.cuda.synchronize()
total = a.elapsed_time(b) / 3
_lib.PROF.records, _lib.PROF.enabled = [], True
vae.decode(z, num_frames=n)
torch.cuda.synchronize()
_lib.PROF.enabled = False
by, gemms = {}, {}
for name, ea, eb, meta in _lib.PROF.records:
    ms = ea.elapsed_time(eb)
    d = by.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
    d["ms"] += ms
    d["calls"] += 1
    if meta:
        d["flops"] += meta.get("flops", 0.0)
        d["bytes"] += meta.get("bytes", 0.0)
    if name == "lkgd_gemm" and meta:
        key = f"mode{meta['mode']} M={meta['M']} N={meta['N']} K={meta['K']}"
        g = gemms.setdefault(key, dict(ms=0.0, calls=0, flops=0.0))
        g["ms"] += ms
        g["calls"] += 1
        g["flops"] += meta["flops"]
fl = vae_flops(SVD_VAE_CONFIG, n, 72, 128)["decode"]
print(f"decode of {n} frames 576x1024: {total:.2f} ms, {fl / 1e12:.1f} TFLOP, {fl / 1e12 / (total * 1e-3):.0f} TF/s")
for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"]):
    extra = f" {v['flops'] / 1e12 / (v['ms'] * 1e-3):.0f} TF/s" if v["flops"] else (
        f" {v['bytes'] / 1e9 / (v['ms'] * 1e-3):.0f} GB/s" if v["bytes"] else "")
    print(f"  {k:32s} {v['ms']:8.3f} ms  {v['calls']:4d} calls{extra}")
for k, v in sorted(gemms.items(), key=lambda kv: -kv[1]["ms"]):
    print(f"    {k:44s} {v['ms']:8.3f} ms {v['calls']:3d} calls {v['flops'] / 1e12 / (v['ms'] * 1e-3):7.0f} TF/s")
if len(sys.argv) > 2:
    json.dump({"frames": n, "ms": total, "tflop": fl / 1e12, "by_kernel": by, "gemm_shapes": gemms}, open(sys.argv[2], "w"), indent=1)
