"""Generates tests/golden/patch_golden.npz by RUNNING THE REFERENCE'S OWN TRANSFORMER-BLOCK FORWARDS:

  patch/patch.py:390-580   ToMeBlock.forward           (restates diffusers BasicTransformerBlock.forward)
  patch/patch.py:582-686   ToMeBlock.forward_temporal  (restates TemporalBasicTransformerBlock.forward)
  patch/patch.py:719-806   apply_patch                 (class swap by class NAME, as the reference's trans* scripts do:
                                                        run_models/run_inference_joint_depth.py:37)

with joint attention off (`enable_joint_attention = False`, the non-joint branches :502-508 and :659-661), i.e. the
arithmetic of the stock blocks.  The blocks those forwards drive are the constructor-only stand-ins of
tests/golden/ref_shim/diffusers/models/attention.py - LayerNorm / Linear / F.scaled_dot_product_attention / F.gelu
straight from torch, NOT from oracle/ - so these vectors pin the oracle's two transformer blocks (and the CUDA path)
against code that does not share the oracle's block arithmetic:

  patch/spatial_*, patch/temporal_*   single blocks, d = 16 and d = 64 heads, batch*frames = 8 / batch = 2, frames = 4
  patch/unet_out                      the whole reduced UNet (models/unet_spatio_temporal_condition_controlnet.py) with
                                      every (Temporal)BasicTransformerBlock replaced by a stand-in + apply_patch

Dev container only (/root/reference); the .npz is committed.      python tests/golden/make_patch_golden.py
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
from patch import patch as ref_patch  # noqa: E402   the reference's file

torch.set_num_threads(8)
out = {}


class ModelMixin(torch.nn.Module):        # apply_patch accepts a model whose class is NAMED ModelMixin (:772-773)
    pass


class Holder(ModelMixin):
    def __init__(self, blk):
        super().__init__()
        self.blk = blk

    def forward(self, x, **kw):
        return self.blk(x, **kw)


def patched(model):
    ref_patch.apply_patch(model, with_spatial_block=True, with_temporal_block=True)
    n = 0
    for m in model.modules():
        if m.__class__.__name__ == "ToMeBlock":
            m.enable_joint_attention = False      # what ToMeBlock.set_joint_attention(False) does (:174-175)
            n += 1
    return n


# ------------------------------------------------------------------------------------------ single blocks
for tag, (dim, heads, dh, xdim) in {"d16": (32, 2, 16, 32), "d64": (128, 2, 64, 48)}.items():
    BF, N, Fr = 8, 24, 4
    sb = Holder(fill_seeded_(BasicTransformerBlock(dim, heads, dh, xdim), seed=11).eval())
    assert patched(sb) == 1
    x = seeded_tensor(f"patch/{tag}/x", (BF, N, dim))
    ctx = seeded_tensor(f"patch/{tag}/ctx", (BF, 1, xdim))
    ctx3 = seeded_tensor(f"patch/{tag}/ctx3", (BF, 3, xdim))
    with torch.no_grad():
        out[f"patch/spatial_{tag}"] = sb(x, encoder_hidden_states=ctx).numpy()
        out[f"patch/spatial_{tag}_kv3"] = sb(x, encoder_hidden_states=ctx3).numpy()
    tb = Holder(fill_seeded_(TemporalBasicTransformerBlock(dim, dim, heads, dh, xdim), seed=12).eval())
    assert patched(tb) == 1
    tctx = seeded_tensor(f"patch/{tag}/tctx", ((BF // Fr) * N, 1, xdim))
    with torch.no_grad():
        out[f"patch/temporal_{tag}"] = tb(x, num_frames=Fr, encoder_hidden_states=tctx).numpy()

# ------------------------------------------------------------------------------------------ whole UNet
from models.unet_spatio_temporal_condition_controlnet import UNetSpatioTemporalConditionControlNetModel  # noqa: E402

REDUCED = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
B, F, H, W = 2, 4, 16, 16
unet = UNetSpatioTemporalConditionControlNetModel(**REDUCED)
swapped = 0
for mod in list(unet.modules()):
    for lst_name, cls in (("transformer_blocks", BasicTransformerBlock),
                          ("temporal_transformer_blocks", TemporalBasicTransformerBlock)):
        lst = getattr(mod, lst_name, None)
        if lst is None:
            continue
        for i, old in enumerate(lst):
            a = old.attn1
            dim = a.to_q.in_features
            xdim = old.attn2.to_k.in_features
            new = cls(dim, a.heads, a.dim_head, xdim) if cls is BasicTransformerBlock else \
                cls(dim, dim, a.heads, a.dim_head, xdim)
            assert [n for n, _ in new.named_parameters()] == [n for n, _ in old.named_parameters()]
            lst[i] = new
            swapped += 1
unet = fill_seeded_(unet).eval()          # name-seeded: the same tensors the un-patched golden ("unet/out") used
n_p = patched(unet)
print('swapped', swapped, 'patched', n_p)
assert n_p == swapped and swapped > 0
sample = seeded_tensor("unet/sample", (B, F, 8, H, W))
ctx = seeded_tensor("unet/ctx", (B, 1, 32))
ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
with torch.no_grad():
    out["patch/unet_out"] = unet(sample, torch.tensor(1.4439898729), ctx, added_time_ids=ids, return_dict=False)[0].numpy()
out["patch/n_blocks"] = np.asarray(swapped)
np.savez_compressed(HERE / "patch_golden.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
ref = np.load(HERE / "reference_golden.npz")["unet/out"]
d = np.linalg.norm(out["patch/unet_out"] - ref) / np.linalg.norm(ref)
print("patched-reference UNet vs the shim/oracle-block golden: rel-L2", d)
