"""Generates tests/golden/flow_golden.npz by RUNNING the reference's flow-stem UNet
(/root/reference/models/unet_spatio_temporal_condition_flow.py, UNetSpatioTemporalConditionModelFlow, unmodified)
through tests/golden/ref_shim - SURVEY 8f N3.  Dev container only; the .npz is committed.

    python tests/golden/make_flow_golden.py
"""
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import fill_seeded_, seeded_tensor  # noqa: E402
from models.unet_spatio_temporal_condition_flow import UNetSpatioTemporalConditionModelFlow  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)
REDUCED = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
B, F, H, W = 2, 4, 16, 16
out = {}
unet = UNetSpatioTemporalConditionModelFlow(**REDUCED)
unet.initialize_conv_in()
unet = fill_seeded_(unet).eval()                     # every parameter (incl. conv_in2 / conv_in2_alpha) from its name
sample = seeded_tensor("flow/sample", (B, F, 12, H, W))
ctx = seeded_tensor("unet/ctx", (B, 1, 32))
ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
t = torch.tensor(1.4439898729)
with torch.no_grad():
    out["flow/out"] = unet(sample, t, ctx, added_time_ids=ids, return_dict=False)[0].numpy()
    # a fresh stem (alpha = 0, conv_in2 = copy of conv_in) must not change the model: same as the 8-channel forward
    unet.conv_in2_alpha.zero_()
    out["flow/out_alpha0"] = unet(sample, t, ctx, added_time_ids=ids).sample.numpy()
out["flow/param_names"] = np.array(sorted(n for n, _ in unet.named_parameters() if n.startswith("conv_in")))
np.savez_compressed(HERE / "flow_golden.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
