"""Generates tests/golden/full_size_<case>.npz: the fp32 CPU oracle's output for ONE full-size UNet forward of the
BASELINE.json configurations, so that `pytest -m gpu` compares the CUDA path at 72x128 latents without paying minutes
of CPU time on the GPU box.

  c3   configs[2]: SVD-XT, 25 frames, 72x128 latents, CFG batch 2 (zero CLIP embedding in the unconditional half),
       LoRA r=64 on the temporal attn1 q/k/v, latent-knowledge conditioning (LKGD UNet)
  c4   configs[3]: SVD, 14 frames, 72x128 latents, CFG batch 2, ControlNetSDVModel (flow, 2 conditioning channels,
       576x1024 condition frames) feeding the plain UNet's F6 residual injection

Weights and inputs are name-seeded (tests/golden/weights.py), so the GPU test rebuilds bit-identical tensors directly
in the product modules; only the oracle's OUTPUT is committed (fp16, <= 3.7 MB per case).

    python tests/golden/make_full_size_golden.py c3 c4        # ~10 min of CPU per case on 8 cores
"""
import json
import os
import pathlib
import sys
import time

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(HERE)]

import oracle as O                      # noqa: E402  (test infrastructure)
from oracle import blocks as OB         # noqa: E402
from weights import fill_seeded_, seeded_tensor  # noqa: E402

CASES = {
    "c3": dict(frames=25, h=72, w=128, rank=64, lkgd=True, controlnet=0, t=1.2),
    "c4": dict(frames=14, h=72, w=128, rank=0, lkgd=False, controlnet=2, t=1.2),
    # small variants of the same recipes (seconds): used by the CPU test of this generator's determinism
    "c3_tiny": dict(frames=3, h=8, w=8, rank=4, lkgd=True, controlnet=0, t=1.2),
    "c4_tiny": dict(frames=2, h=8, w=8, rank=0, lkgd=False, controlnet=2, t=1.2),
}

_orig_attn = OB.Attention.forward


def _chunked(self, x, encoder_hidden_states=None):
    """softmax(QK^T) of 9216 tokens is 1.7 GB per image: evaluate image by image (same arithmetic)."""
    if encoder_hidden_states is not None or x.shape[1] < 4096:
        return _orig_attn(self, x, encoder_hidden_states)
    return torch.cat([_orig_attn(self, x[i:i + 1]) for i in range(x.shape[0])], 0)


def case_inputs(name):
    """Name-seeded inputs of a case (shared with tests/test_full_size_gpu.py)."""
    c = CASES[name]
    B, F, h, w = 2, c["frames"], c["h"], c["w"]
    x = seeded_tensor(f"{name}/sample", (B, F, 8, h, w))
    ctx = seeded_tensor(f"{name}/ctx", (B, 1, 1024))
    ctx[0] = 0                                     # the pipeline's unconditional half
    ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
    extra = ()
    if c["lkgd"]:
        extra = (seeded_tensor(f"{name}/domain", (1, 1, 1000)), seeded_tensor(f"{name}/flow", (1, 1, 1000)))
    cond = None
    if c["controlnet"]:
        cond = seeded_tensor(f"{name}/cond", (B, F, c["controlnet"], 8 * h, 8 * w)).clamp(-1, 1)
    return x, ctx, ids, extra, cond


def unet_config(name):
    from oracle.unet import SVD_XT_CONFIG
    c = CASES[name]
    return dict(SVD_XT_CONFIG, num_frames=c["frames"], cross_attention_dim=1024)


def build_oracle(name):
    c = CASES[name]
    cfg = unet_config(name)
    cls = O.UNetSpatioTemporalConditionModel if c["lkgd"] else O.UNetSpatioTemporalConditionControlNetModel
    with torch.device("meta"):
        u = cls(**cfg)
        if c["rank"]:
            O.add_lora(u, c["rank"])
    u = fill_seeded_(u.to_empty(device="cpu")).eval()
    if c["lkgd"]:
        u.canonical_zero_phase = True      # zero embedding: the GPU reference's FFT zeros (+0), see oracle/unet.py
    cn = None
    if c["controlnet"]:
        with torch.device("meta"):
            cn = O.ControlNetSDVModel(**{k: v for k, v in cfg.items() if k != "up_block_types"},
                                      conditioning_channels=c["controlnet"])
        cn = fill_seeded_(cn.to_empty(device="cpu"), seed=1).eval()
    return u, cn


def run(name):
    torch.set_num_threads(os.cpu_count() or 1)
    OB.Attention.forward = _chunked
    c = CASES[name]
    t0 = time.time()
    u, cn = build_oracle(name)
    x, ctx, ids, extra, cond = case_inputs(name)
    t_build = time.time() - t0
    t1 = time.time()
    out = {}
    with torch.no_grad():
        kw = {}
        if cn is not None:
            down, mid = cn(x, c["t"], ctx, ids, controlnet_cond=cond, conditioning_scale=1.0, return_dict=False)
            kw = dict(down_block_additional_residuals=down, mid_block_additional_residual=mid)
            out["cn_mid"] = mid[:, :128].half().numpy()          # first 128 of 1280 channels: keeps the fixture small
            out["cn_down_norms"] = np.asarray([float(d.double().norm()) for d in down])
        y = u(x, c["t"], ctx, *extra, added_time_ids=ids, return_dict=False, **kw)[0]
    t_fwd = time.time() - t1
    out["out"] = y.half().numpy()
    out["out_rms"] = np.asarray(float(y.double().pow(2).mean().sqrt()))
    out["meta"] = np.asarray(json.dumps(dict(case=name, **c, oracle_seconds=round(t_fwd, 1), build_seconds=round(t_build, 1),
                                             threads=torch.get_num_threads(), torch=torch.__version__)))
    path = HERE / f"full_size_{name}.npz"
    np.savez(path, **out)
    print("wrote", path, {k: getattr(v, "shape", None) for k, v in out.items()}, f"oracle {t_fwd:.0f} s", flush=True)


if __name__ == "__main__":
    for n in sys.argv[1:] or ["c3", "c4"]:
        run(n)
