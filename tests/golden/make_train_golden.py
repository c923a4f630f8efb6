"""Generates tests/golden/full_size_c5.npz: loss and LoRA / quaternion gradients of ONE full-size training step
(BASELINE.json configs[4], "C5": SVD-XT-width LKGD UNet, LoRA r = 64 on the temporal attn1 q / k / v, 14 frames of 320x512 =
40x64 latents, batch 1) by PyTorch autograd through the fp32 CPU oracle - the computation the reference's
`accelerator.backward(loss)` performs (train_models/train_svd_lora.py:1503-1530,1634-1642,1651-1683).

Committed: the loss, the norm of EVERY trainable tensor's gradient (125 names) and a handful of complete gradient tensors
(fp32) from the first down block, the mid block, the last up block and the latent-knowledge block.  Weights / inputs are
name-seeded (tests/golden/weights.py), so `tests/test_full_size_gpu.py` rebuilds identical tensors in the product modules.

    python tests/golden/make_train_golden.py        # ~3 min of CPU on 8 cores, ~25 GB of RAM
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
from oracle.pipeline import train_loss, train_precondition  # noqa: E402
from oracle.unet import SVD_XT_CONFIG   # noqa: E402
from weights import fill_seeded_, seeded_tensor  # noqa: E402

FRAMES, H, W, RANK = 14, 40, 64, 64
FULL = ("down_blocks.0.attentions.0.temporal_transformer_blocks.0.attn1.to_q.lora_A.default.weight",
        "down_blocks.0.attentions.0.temporal_transformer_blocks.0.attn1.to_v.lora_B.default.weight",
        "mid_block.attentions.0.temporal_transformer_blocks.0.attn1.to_k.lora_B.default.weight",
        "up_blocks.3.attentions.2.temporal_transformer_blocks.0.attn1.to_q.lora_B.default.weight",
        "up_blocks.3.attentions.2.temporal_transformer_blocks.0.attn1.to_v.lora_A.default.weight",
        "quaternion_lora_texts", "quaternion_lora_fuse.r_weight", "quaternion_lora_dconv.weight")


def config():
    return dict(SVD_XT_CONFIG, num_frames=FRAMES, cross_attention_dim=1024)


def inputs():
    lat = seeded_tensor("c5/latents", (1, FRAMES, 4, H, W))
    noise = seeded_tensor("c5/noise", (1, FRAMES, 4, H, W))
    cond = seeded_tensor("c5/cond", (1, 4, H, W))
    ctx = seeded_tensor("c5/ctx", (1, 1, 1024))
    sig = torch.tensor([1.3])
    ids = O.add_time_ids_training(5, 127, 0.02, 1)
    extra = (seeded_tensor("c5/domain", (1, 1, 1000)), seeded_tensor("c5/flow", (1, 1, 1000)))
    return lat, noise, cond, ctx, sig, ids, extra


def trainable(name):
    return "lora_" in name and ("lora_A" in name or "lora_B" in name) or "quaternion" in name


if __name__ == "__main__":
    torch.set_num_threads(os.cpu_count() or 1)
    t0 = time.time()
    with torch.device("meta"):
        u = O.UNetSpatioTemporalConditionModel(**config())
        O.add_lora(u, RANK)
    u = fill_seeded_(u.to_empty(device="cpu")).eval()
    for n, p in u.named_parameters():
        p.requires_grad_(trainable(n))
    lat, noise, cond, ctx, sig, ids, extra = inputs()
    noisy, timesteps, inp = train_precondition(lat, noise, sig)
    x = torch.cat([inp, cond.unsqueeze(1).repeat(1, FRAMES, 1, 1, 1)], dim=2)
    t1 = time.time()
    pred = u(x, timesteps, ctx, *extra, added_time_ids=ids, return_dict=False)[0]
    loss = train_loss(pred, noisy, lat, sig)
    loss.backward()
    secs = time.time() - t1
    grads = {n: p.grad for n, p in u.named_parameters() if p.requires_grad}
    out = {"loss": np.asarray(float(loss)), "names": np.array(list(grads)),
           "norms": np.asarray([float(g.double().norm()) for g in grads.values()])}
    for n in FULL:
        out["grad/" + n] = grads[n].float().numpy()         # fp32: some gradients are far below the fp16 range
    out["meta"] = np.asarray(json.dumps(dict(frames=FRAMES, h=H, w=W, rank=RANK, oracle_seconds=round(secs, 1),
                                             build_seconds=round(t1 - t0, 1), threads=torch.get_num_threads(),
                                             torch=torch.__version__)))
    np.savez_compressed(HERE / "full_size_c5.npz", **out)
    print("loss", float(loss), "tensors", len(grads), f"fwd+bwd {secs:.0f} s", {k: v.shape for k, v in out.items() if k.startswith("grad/")})
