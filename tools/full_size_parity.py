#!/usr/bin/env python
"""FULL-SIZE parity of one UNet forward: BASELINE.json's headline configuration (SVD-XT, 25 frames, 72x128 latents, CFG
batch 2, LoRA r=64 on the temporal attn1 q/k/v, latent-knowledge conditioning) on the CUDA path against the fp32 CPU
oracle on identical random-init weights and inputs.  The oracle forward is ~161 TFLOP of eager fp32 PyTorch on the host
cores (minutes), so this is a tool, not a test:

    python tools/full_size_parity.py [--frames 25] [--out gpurun_out/full_size_parity.json]

The oracle's attention materialises softmax(QK^T); at 9216 tokens that is 1.7 GB per image, so the spatial attention is
evaluated image by image here (same arithmetic, bounded memory)."""
import argparse, json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle as O                      # test infrastructure: the checker, never the thing measured
from oracle import blocks as OB
from lkgd_b200.unet import SVD_XT_CONFIG, UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionModel

ap = argparse.ArgumentParser()
ap.add_argument("--frames", type=int, default=25)
ap.add_argument("--h", type=int, default=72)
ap.add_argument("--w", type=int, default=128)
ap.add_argument("--rank", type=int, default=64)
ap.add_argument("--out", default="gpurun_out/full_size_parity.json")
ap.add_argument("--plain", action="store_true", help="the plain (ControlNet-accepting) UNet instead of the LKGD UNet")
ap.add_argument("--zero-uncond", type=int, default=1, help="1: zero CLIP embedding for the unconditional half (pipeline)")
ap.add_argument("--attn-checker", action="store_true", help="spatial attention through the fp32 SIMT checker kernel")
ap.add_argument("--gemm-checker", action="store_true", help="every GEMM / conv through the fp32 SIMT checker kernel")
ap.add_argument("--ref-cache", default=None, help="file that keeps the oracle output between runs of the same configuration")
a = ap.parse_args()
torch.set_num_threads(os.cpu_count() or 1)

_orig = OB.Attention.forward


def chunked(self, x, encoder_hidden_states=None):
    if encoder_hidden_states is not None or x.shape[1] < 4096:
        return _orig(self, x, encoder_hidden_states)
    return torch.cat([_orig(self, x[i:i + 1]) for i in range(x.shape[0])], 0)


OB.Attention.forward = chunked

cfg = dict(SVD_XT_CONFIG, num_frames=a.frames, cross_attention_dim=1024)
t0 = time.time()
torch.manual_seed(0)
OCLS = O.UNetSpatioTemporalConditionControlNetModel if a.plain else O.UNetSpatioTemporalConditionModel
PCLS = UNetSpatioTemporalConditionControlNetModel if a.plain else UNetSpatioTemporalConditionModel
with torch.device("meta"):
    o = OCLS(**cfg)
    if a.rank:
        O.add_lora(o, a.rank)
o = o.to_empty(device="cpu").eval()
# the unconditional half's zero embedding: take the GPU reference's FFT zeros (+0), not the CPU library's (oracle/unet.py)
o.canonical_zero_phase = True
g = torch.Generator().manual_seed(0)
with torch.no_grad():
    for n, p in o.named_parameters():
        if n.endswith("mix_factor"):
            p.copy_(torch.rand(p.shape, generator=g) * 2 - 1)
        elif "lora_B" in n or n.startswith("quaternion_lora_texts"):
            p.copy_((torch.randn(p.shape, generator=g) * 0.02).to(torch.bfloat16).float())
        elif p.ndim >= 2:
            fan_in = p[0].numel()
            p.copy_(((torch.rand(p.shape, generator=g) * 2 - 1) * fan_in ** -0.5).to(torch.bfloat16).float())
        elif "norm" in n and n.endswith("weight"):
            p.copy_(1 + 0.1 * torch.randn(p.shape, generator=g))
        else:
            p.copy_(0.02 * torch.randn(p.shape, generator=g))
p = PCLS(**cfg)
if a.rank:
    p.add_lora(a.rank)
p.load_state_dict(o.state_dict(), strict=True)
p = p.to("cuda")
B = 2
x = torch.randn(B, a.frames, 8, a.h, a.w, generator=g)
ctx = torch.randn(B, 1, 1024, generator=g)
if a.zero_uncond:
    ctx[0] = 0
ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
dom, flo = torch.randn(1, 1, 1000, generator=g), torch.randn(1, 1, 1000, generator=g)
t_build = time.time() - t0
if a.attn_checker or a.gemm_checker:      # error attribution: same bf16 operands, reference-quality SIMT kernels
    import functools
    from lkgd_b200 import ops
    if a.attn_checker:
        ops.attention = functools.partial(ops.attention, checker=True)
    if a.gemm_checker:
        ops.gemm = functools.partial(ops.gemm, checker=True)
ts = 1.2
extra_o = () if a.plain else (dom, flo)
extra_p = () if a.plain else (dom.cuda(), flo.cuda())
got = p(x.cuda(), ts, ctx.cuda(), *extra_p, added_time_ids=ids.cuda(), return_dict=False)[0]
torch.cuda.synchronize()
t1 = time.time()
if a.ref_cache and os.path.exists(a.ref_cache):
    ref = torch.load(a.ref_cache)
else:
    with torch.no_grad():
        ref = o(x, ts, ctx, *extra_o, added_time_ids=ids, return_dict=False)[0]
    if a.ref_cache:
        torch.save(ref, a.ref_cache)
t_ref = time.time() - t1
d = (got.double().cpu() - ref.double())
rel = float(d.norm() / ref.double().norm())
res = dict(config=f"SVD-XT {'plain' if a.plain else 'LKGD'} UNet, LoRA r={a.rank}, {a.frames} frames {a.h}x{a.w} latents, CFG batch 2, fp32 CPU oracle vs CUDA path",
           rel_l2=rel, max_abs=float(d.abs().max()), ref_rms=float(ref.double().pow(2).mean().sqrt()),
           finite=bool(torch.isfinite(got).all()), tolerance=1e-2, oracle_seconds=round(t_ref, 1),
           build_seconds=round(t_build, 1), host_threads=torch.get_num_threads())
print(json.dumps(res))
os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
json.dump(res, open(a.out, "w"), indent=1)
