"""Adapter wire format (lkgd_b200/lora_io.py) against the key lists the REFERENCE's own helpers produce
(tests/golden/lora_keys.json <- utils/peft_utils.py::get/set_peft_model_state_dict run by
tests/golden/make_lora_keys_golden.py) and through a save -> load round trip of the safetensors file the reference's
training loop writes (train_models/train_svd_lora.py:1735-1747)."""
import json
import os

import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "lora_keys.json")))


def _model(adapter="default", r=4, lora=True):
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionModel
    torch.manual_seed(0)
    m = UNetSpatioTemporalConditionModel(**dict(REDUCED_CONFIG, cross_attention_dim=1024))
    if lora:
        m.add_lora(r, adapter_name=adapter)
        with torch.no_grad():
            for n, p in m.named_parameters():
                if "lora_" in n:
                    p.copy_(torch.randn(p.shape) * 0.1)
    return m


@pytest.mark.parametrize("adapter", ["default", "y_lora"])
def test_saved_keys_match_the_reference_helper(adapter):
    from lkgd_b200 import lora_io
    sd = lora_io.get_peft_model_state_dict(_model(adapter), adapter)
    assert sorted(sd) == GOLD[adapter]["saved_keys"]                      # 36 LoRA tensors + 29 quaternion tensors
    assert {k: list(v.shape) for k, v in sd.items()} == GOLD[adapter]["shapes"]
    # re-inserting the adapter name reproduces the module keys the reference's loader restores
    lora_keys = [k for k in sd if ".lora_A." in k or ".lora_B." in k]
    assert sorted(lora_io._with_adapter_name(k, adapter) for k in lora_keys) == GOLD[adapter]["module_keys_restored"]
    # reference quirk kept on record: its loader mangles the quaternion keys and reports them as unexpected
    assert len(GOLD[adapter]["unexpected"]) == 29


def test_file_format_and_round_trip(tmp_path):
    from safetensors import safe_open
    from lkgd_b200 import lora_io
    from lkgd_b200.modules import LoraLinear
    src = _model()
    path = lora_io.save_lora_weights(src, str(tmp_path))
    assert os.path.basename(path) == "pytorch_lora_weights.safetensors"
    with safe_open(path, framework="pt") as f:
        keys = list(f.keys())
        assert f.metadata() == {"format": "pt"}
    assert all(k.startswith("unet.") for k in keys) and len(keys) == 65
    assert any(k.endswith("attn1.to_q.lora.down.weight") for k in keys)       # diffusers naming of lora_A / lora_B
    assert any(k.endswith("attn1.to_v.lora.up.weight") for k in keys)
    assert "unet.quaternion_lora_fuse.r_weight" in keys and not any("lora_A" in k for k in keys)
    # load into a model that has NO adapters yet: wrappers are created with the rank found in the file
    dst = _model(lora=False)
    res = lora_io.load_lora_weights(dst, str(tmp_path))
    assert len(res["loaded"]) == 65 and not res["unexpected"]
    a, b = src.state_dict(), dst.state_dict()
    for k in a:
        if "lora_" in k:
            assert torch.equal(a[k], b[k]), k
    wrapped = [m for m in dst.modules() if isinstance(m, LoraLinear)]
    assert len(wrapped) == 18 and all(m.r == 4 and m.scaling == 1.0 for m in wrapped)
    # PEFT-format keys without the unet. prefix load too
    from safetensors.torch import save_file
    alt = tmp_path / "peft.safetensors"
    save_file({k: v.contiguous() for k, v in lora_io.get_peft_model_state_dict(src).items()}, str(alt))
    dst2 = _model(lora=False)
    assert len(lora_io.load_lora_weights(dst2, str(alt))["loaded"]) == 65


def test_rank_mismatch_and_unexpected_keys_raise(tmp_path):
    from safetensors.torch import save_file
    from lkgd_b200 import lora_io
    src = _model(r=4)
    lora_io.save_lora_weights(src, str(tmp_path))
    with pytest.raises(ValueError):
        lora_io.load_lora_weights(_model(r=8), str(tmp_path))
    bad = tmp_path / "bad.safetensors"
    save_file({"unet.not_a_module.lora_A.weight": torch.zeros(4, 8)}, str(bad))
    with pytest.raises((KeyError, AttributeError)):
        lora_io.load_lora_weights(_model(), str(bad))
