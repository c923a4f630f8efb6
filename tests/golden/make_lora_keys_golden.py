"""Generates tests/golden/lora_keys.json by running the REFERENCE's own adapter (de)serialisation helpers,
``utils/peft_utils.py::get_peft_model_state_dict`` / ``set_peft_model_state_dict`` (/root/reference, unmodified; peft /
huggingface_hub imports satisfied by tests/golden/ref_shim), on the oracle LKGD UNet with the training script's adapter
config (train_models/train_svd_lora.py:1081-1102).  Dev container only.

    python tests/golden/make_lora_keys_golden.py
"""
import json
import pathlib
import sys
from types import SimpleNamespace

import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

import oracle as O  # noqa: E402
from peft.utils.peft_types import PeftType  # noqa: E402
from utils.peft_utils import get_peft_model_state_dict, set_peft_model_state_dict  # noqa: E402

REDUCED = dict(O.REDUCED_CONFIG, cross_attention_dim=1024)
out = {}
for adapter in ("default", "y_lora"):
    torch.manual_seed(0)
    unet = O.UNetSpatioTemporalConditionModel(**REDUCED)
    O.add_lora(unet, r=4, adapter_name=adapter)
    unet.peft_config = {adapter: SimpleNamespace(peft_type=PeftType.LORA, bias="none", is_prompt_learning=False,
                                                 target_modules=["attn1.to_k", "attn1.to_q", "attn1.to_v"],
                                                 base_model_name_or_path=None)}
    sd = get_peft_model_state_dict(unet, adapter_name=adapter)
    # round trip through the reference's loader: perturb, load back, compare
    with torch.no_grad():
        new = {k: torch.full_like(v, 0.25) for k, v in sd.items()}
    res = set_peft_model_state_dict(unet, new, adapter_name=adapter)
    own = unet.state_dict()
    restored = sorted(k for k, v in own.items() if ("lora_" in k) and bool((v == 0.25).all()))
    out[adapter] = dict(saved_keys=sorted(sd), shapes={k: list(v.shape) for k, v in sd.items()},
                        module_keys_restored=restored, unexpected=list(res.unexpected_keys))
    print(adapter, len(sd), "saved keys;", len(restored), "module tensors restored;", len(res.unexpected_keys), "unexpected")
json.dump(out, open(HERE / "lora_keys.json", "w"), indent=0, sort_keys=True)
