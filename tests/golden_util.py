"""Shared helpers for the golden-vector tests (CPU and GPU)."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from weights import fill_seeded_, seeded_tensor  # noqa: E402,F401

REDUCED4 = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
SCHED = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
             prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
             sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1)
B, F, H, W = 2, 4, 16, 16
T_STEP = 1.4439898729


def golden():
    """tests/golden/reference_golden.npz - produced by running the reference's own files
    (tests/golden/make_reference_golden.py)."""
    return np.load(os.path.join(HERE, "golden", "reference_golden.npz"))


def t(a):
    return torch.from_numpy(np.asarray(a))


def unet_inputs():
    sample = seeded_tensor("unet/sample", (B, F, 8, H, W))
    ctx = seeded_tensor("unet/ctx", (B, 1, 32))
    ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
    return sample, ctx, ids


def unet_residuals():
    shapes = [(B * F, 32, 16, 16)] * 3 + [(B * F, 32, 8, 8)] + [(B * F, 64, 8, 8)] * 2
    res = [seeded_tensor(f"unet/res{i}", s, scale=0.5) for i, s in enumerate(shapes)]
    mid = seeded_tensor("unet/resmid", (B * F, 64, 8, 8), scale=0.5)
    return res, mid


def rel(a, b):
    a = torch.as_tensor(a).double().flatten().cpu()
    b = torch.as_tensor(b).double().flatten().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))
