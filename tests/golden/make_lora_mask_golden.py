"""Generates tests/golden/lora_mask_golden.npz by RUNNING the reference's per-sample masked multi-adapter LoRA forward:
`patch.lora_forward_hack` (patch/patch.py:57-92) installed on the reference's own LoRA layer (models/lora_layer.py Linear,
two adapters through its `update_layer`), with masks in the form `patch.set_patch_lora_mask` stores them.  Dev container
only; the .npz is committed.        python tests/golden/make_lora_mask_golden.py
"""
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import fill_seeded_, seeded_tensor  # noqa: E402
from models.lora_layer import Linear as RefLoraLinear  # noqa: E402
from patch import patch as ref_patch  # noqa: E402

out = {}
base = torch.nn.Linear(32, 48)
lora = RefLoraLinear(base, "xy_lora", r=4, lora_alpha=4, init_lora_weights="gaussian")
lora.update_layer("yx_lora", r=8, lora_alpha=4, lora_dropout=0.0, init_lora_weights="gaussian", use_rslora=False)
lora.set_adapter(["xy_lora", "yx_lora"])
lora = fill_seeded_(lora, seed=3)
x = seeded_tensor("loramask/x", (8, 5, 32))               # leading dim 8 = 4 samples x 2 frames
with torch.no_grad():
    out["loramask/y_unmasked"] = lora(x).numpy()           # stock forward: both adapters on every sample
    lora.forward = ref_patch.lora_forward_hack(lora)       # what hack_lora_forward does per module (:911-922)
    lora.lora_mask = {"xy_lora": torch.tensor([1, 0, 1, 0], dtype=torch.bool),
                      "yx_lora": torch.tensor([0, 1, 0, 1], dtype=torch.bool)}
    out["loramask/y_masked"] = lora(x).numpy()
    lora.lora_mask = {"xy_lora": torch.tensor([1, 1], dtype=torch.bool), "yx_lora": torch.tensor([0, 1], dtype=torch.bool)}
    out["loramask/y_masked2"] = lora(x).numpy()
out["loramask/names"] = np.array(sorted(n for n, _ in lora.named_parameters()))
np.savez_compressed(HERE / "lora_mask_golden.npz", **out)
print({k: getattr(v, "shape", v) for k, v in out.items()})
