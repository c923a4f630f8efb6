"""Generates tests/golden/joint_attention_golden.npz by RUNNING the reference's joint-attention patch:

  patch/patch.py:94-688   ToMeBlock with enable_joint_attention = True: a second attention `attn1n` whose keys / values
                          come from the PARTNER sample (mask-selected swap :452-456, optional frame flip :458-464), the
                          zero-initialised post layer (`conv` / `scale`, :143-172) and `joint_scale` (:491); the temporal
                          variant :617-658
  patch/patch.py:719-806  apply_patch(flip, with_spatial_block, with_temporal_block); :938-1004 set_joint_* helpers

on the reference's own UNet file (models/unet_spatio_temporal_condition_controlnet.py) whose transformer blocks are the
torch-primitive stand-ins of tests/golden/ref_shim (no oracle arithmetic in them).  Dev container only.

    python tests/golden/make_joint_attention_golden.py
"""
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import fill_seeded_, seeded_tensor  # noqa: E402
from diffusers.models.attention import BasicTransformerBlock, TemporalBasicTransformerBlock  # noqa: E402  (shim)
from patch import patch as ref_patch  # noqa: E402
from models.unet_spatio_temporal_condition_controlnet import UNetSpatioTemporalConditionControlNetModel  # noqa: E402

torch.set_num_threads(8)
REDUCED = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
B, F, H, W = 4, 4, 16, 16
out = {}


def build(post, flip, temporal):
    unet = UNetSpatioTemporalConditionControlNetModel(**REDUCED)
    for mod in list(unet.modules()):
        for lst_name, cls in (("transformer_blocks", BasicTransformerBlock),
                              ("temporal_transformer_blocks", TemporalBasicTransformerBlock)):
            lst = getattr(mod, lst_name, None)
            if lst is None:
                continue
            for i, old in enumerate(lst):
                a = old.attn1
                dim, xdim = a.to_q.in_features, old.attn2.to_k.in_features
                lst[i] = cls(dim, a.heads, a.dim_head, xdim) if cls is BasicTransformerBlock else \
                    cls(dim, dim, a.heads, a.dim_head, xdim)
    ref_patch.apply_patch(unet, flip=flip, with_spatial_block=True, with_temporal_block=temporal)
    ref_patch.initialize_joint_layers(unet, post=post)
    unet = fill_seeded_(unet).eval()            # attn1n.* / conv1n.weight / scale1n get name-seeded, NON-zero values
    ref_patch.set_joint_attention(unet, True)
    if not temporal:                            # un-patched temporal blocks are stand-ins without a forward: patch them
        for m in unet.modules():                # with joint attention OFF (the stock arithmetic, patch.py:659-661)
            if m.__class__.__name__ == "TemporalBasicTransformerBlock":
                m.__class__ = ref_patch.make_diffusers_tome_block(m.__class__)
                m.forward = m.forward_temporal
                m.enable_joint_attention = False
    return unet


sample = seeded_tensor("ja/sample", (B, F, 8, H, W))
ctx = seeded_tensor("ja/ctx", (B, 1, 32))
ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
t = torch.tensor(1.4439898729)
cases = {"conv": ("conv", False, True, [0, 1, 0, 1], 1.0), "conv_flip": ("conv", True, False, [0, 1, 0, 1], 0.7),
         "scale_pair": ("scale", False, True, [0, 1], 1.0),
         # post = "conv_fuse" (:154-157, :488-494; most of the reference's gradio configurations): one [2C, 2C] layer over
         # [masked sample | partner]; the temporal forward has no branch for it and adds the raw attention output
         "conv_fuse": ("conv_fuse", False, True, [0, 1, 0, 1], 0.8), "conv_fuse_flip": ("conv_fuse", True, False, [1, 0], 1.0)}
for tag, (post, flip, temporal, mask, jscale) in cases.items():
    unet = build(post, flip, temporal)
    ref_patch.set_joint_attention_mask(unet, mask)
    ref_patch.set_joint_scale(unet, jscale)
    with torch.no_grad():
        out[f"ja/out_{tag}"] = unet(sample, t, ctx, added_time_ids=ids, return_dict=False)[0].numpy()
    if tag == "conv":
        out["ja/param_names"] = np.array(sorted(n for n, _ in unet.named_parameters() if "1n" in n))
        ref_patch.set_joint_attention(unet, False)
        with torch.no_grad():
            out["ja/out_off"] = unet(sample, t, ctx, added_time_ids=ids, return_dict=False)[0].numpy()
np.savez_compressed(HERE / "joint_attention_golden.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
print("joint on vs off:", np.linalg.norm(out["ja/out_conv"] - out["ja/out_off"]) / np.linalg.norm(out["ja/out_off"]))
