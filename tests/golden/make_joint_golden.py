"""Generates tests/golden/joint_golden.npz by RUNNING the reference's dual-input-head UNet
(/root/reference/models/unet_spatio_temporal_condition_joint.py, UNetSpatioTemporalConditionJointModel, unmodified)
through tests/golden/ref_shim, with the batch masks installed by the reference's own
patch.set_patch_lora_mask (patch/patch.py:872-896) - SURVEY 8f N3.  Dev container only; the .npz is committed.

    python tests/golden/make_joint_golden.py
"""
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import fill_seeded_, seeded_tensor  # noqa: E402
from models.unet_spatio_temporal_condition_joint import UNetSpatioTemporalConditionJointModel  # noqa: E402
from patch import patch as ref_patch  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)
REDUCED = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
B, F, H, W = 4, 4, 16, 16
out = {}
unet = UNetSpatioTemporalConditionJointModel(**REDUCED)
unet.add_y_input_head()
unet = fill_seeded_(unet).eval()               # conv_in_y / time_embedding_y / add_embedding_y get their own tensors
sample = seeded_tensor("joint/sample", (B, F, 8, H, W))
ctx = seeded_tensor("joint/ctx", (B, 1, 32))
ids = torch.tensor([[6.0, 127.0, 0.02], [6.0, 60.0, 0.02], [6.0, 127.0, 0.1], [12.0, 127.0, 0.02]])
t = torch.tensor([1.4439898729, 0.3, 1.4439898729, -0.7])
# (a batch with an EMPTY branch crashes the reference in input_layers' reshape((0, -1)), :415 - not a supported call)
for tag, (xy, yx) in {"alt": ([1, 0, 1, 0], [0, 1, 0, 1]), "pair": ([1, 0], [0, 1]), "yxxx": ([0, 1, 1, 1], [1, 0, 0, 0])}.items():
    ref_patch.set_patch_lora_mask(unet, "xy_lora", xy)        # utils/util.py:432-438 installs them like this
    ref_patch.set_patch_lora_mask(unet, "yx_lora", yx)
    with torch.no_grad():
        out[f"joint/out_{tag}"] = unet(sample, t, ctx, added_time_ids=ids, return_dict=False)[0].numpy()
out["joint/param_names"] = np.array(sorted(n for n, _ in unet.named_parameters() if "_y." in n))
np.savez_compressed(HERE / "joint_golden.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
