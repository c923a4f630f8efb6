"""CPU tests of the drop-in constructors: `add_adapter(LoraConfig)`, `from_reference(module)`,
`from_pretrained(dir, subfolder=...)` / `save_pretrained` - the calls the reference's scripts make
(run_models/run_inference.py:279-281, train_models/train_svd_lora.py:1022-1027,1081-1102,
run_models/run_inference_flow_lora.py:326-331)."""
import json
import os
from types import SimpleNamespace

import pytest
import torch

from golden_util import REDUCED4

HERE = os.path.dirname(os.path.abspath(__file__))


def _svd_meta(cls_name="UNetSpatioTemporalConditionModel"):
    import lkgd_b200.unet as U
    with torch.device("meta"):
        return getattr(U, cls_name)(**U.SVD_XT_CONFIG)


def test_add_adapter_with_the_references_lora_config():
    """train_svd_lora.py:1081-1088: LoraConfig(r, lora_alpha=r, init_lora_weights="gaussian", layers_to_transform=[0],
    layers_pattern="temporal_transformer_blocks.*", target_modules=["attn1.to_k", "attn1.to_q", "attn1.to_v"])."""
    cfg = SimpleNamespace(r=64, lora_alpha=64, init_lora_weights="gaussian", layers_to_transform=[0],
                          layers_pattern="temporal_transformer_blocks.*",
                          target_modules=["attn1.to_k", "attn1.to_q", "attn1.to_v"], lora_dropout=0.0)
    a = _svd_meta()
    hit = a.add_adapter(cfg, adapter_name="default")
    b = _svd_meta()
    assert sorted(hit) == sorted(b.add_lora(64))                   # the regex shorthand selects the same 48 modules
    assert len(hit) == 48 and all(".temporal_transformer_blocks.0.attn1.to_" in n for n in hit)
    names = {n for n, _ in a.named_parameters()}
    d = json.load(open(os.path.join(HERE, "golden", "param_names.json")))
    assert {n for n in names if "lora_" in n} == set(d["trainable"])      # the reference's train_svd_lora_train.txt
    assert a.get_submodule(hit[0]).scaling == 1.0
    # dict-shaped config, custom adapter name (train_svd_lora.py:1099-1102 `adapter_name=lora_name`)
    c = _svd_meta()
    c.add_adapter(dict(r=8, lora_alpha=16, target_modules=["attn1.to_q"], layers_to_transform=0,
                       layers_pattern="temporal_transformer_blocks.*"), adapter_name="y_lora")
    keys = [n for n, _ in c.named_parameters() if "lora_A" in n]
    assert len(keys) == 16 and all(k.endswith("attn1.to_q.lora_A.y_lora.weight") for k in keys)


def test_add_adapter_all_attention_projections_and_errors():
    """run_inference_flow_lora.py:326-331: target_modules=["to_k", "to_q", "to_v", "to_out.0"], r=128."""
    import lkgd_b200.unet as U
    m = _svd_meta("UNetSpatioTemporalConditionControlNetModel")
    hit = m.add_adapter(SimpleNamespace(r=128, lora_alpha=128, init_lora_weights="gaussian",
                                        target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
    assert len(hit) == 16 * 2 * 2 * 4                      # 16 transformers x (spatial, temporal) x (attn1, attn2) x 4
    import re
    assert all(re.match(U.ALL_ATTN_PROJ, n) for n in hit)
    m2 = _svd_meta()
    assert len(m2.add_adapter(dict(r=4, target_modules=r".*attn2\.to_v"))) == 32          # string = full-match regex
    with pytest.raises(ValueError, match="positive integer"):
        _svd_meta().add_adapter(dict(r=0, target_modules=["to_q"]))
    with pytest.raises(ValueError, match="not found"):
        _svd_meta().add_adapter(dict(r=4, target_modules=["nope"]))
    with pytest.raises(ValueError, match="target_modules"):
        _svd_meta().add_adapter(dict(r=4))
    with pytest.raises(ValueError, match="already"):
        m.add_adapter(dict(r=4, target_modules=["to_q"]))


@pytest.mark.parametrize("kind", ["plain", "lkgd_lora", "controlnet"])
def test_from_reference_round_trips_config_and_state_dict(kind):
    """A live reference-shaped module (the oracle's classes carry the reference's config attributes and parameter names -
    pinned by tests/golden/param_names.json) -> lkgd_b200 module: same config, bit-identical state_dict."""
    import oracle as O
    import lkgd_b200.unet as U
    torch.manual_seed(0)
    if kind == "plain":
        ref, cls = O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4), U.UNetSpatioTemporalConditionControlNetModel
    elif kind == "lkgd_lora":
        ref = O.UNetSpatioTemporalConditionModel(**dict(REDUCED4, cross_attention_dim=1024))
        O.add_lora(ref, 4)
        with torch.no_grad():
            for n, p in ref.named_parameters():
                if "lora_B" in n:
                    p.normal_(0, 0.1)
        cls = U.UNetSpatioTemporalConditionModel
    else:
        ref = O.ControlNetSDVModel(**{k: v for k, v in REDUCED4.items() if k != "up_block_types"}, conditioning_channels=2)
        cls = U.ControlNetSDVModel
    got = cls.from_reference(ref)
    a, b = ref.state_dict(), got.state_dict()
    assert sorted(a) == sorted(b)
    assert all(torch.equal(a[k], b[k]) for k in a)
    for k in ("in_channels", "block_out_channels", "num_attention_heads", "cross_attention_dim", "num_frames",
              "addition_time_embed_dim", "layers_per_block"):
        assert getattr(got.config, k) == getattr(ref.config, k), k
    if kind == "lkgd_lora":
        from lkgd_b200.modules import LoraLinear
        assert sum(isinstance(m, LoraLinear) for m in got.modules()) == 6 * 3      # 6 transformers x q,k,v


def test_from_pretrained_and_save_pretrained(tmp_path):
    """run_inference.py:279-280 `Model.from_pretrained(path, subfolder="unet")` on a local diffusers directory."""
    import lkgd_b200.unet as U
    torch.manual_seed(1)
    src = U.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    root = tmp_path / "svd"
    src.save_pretrained(str(root / "unet"))
    cfg = json.load(open(root / "unet" / "config.json"))
    assert cfg["_class_name"] == "UNetSpatioTemporalConditionControlNetModel" and cfg["block_out_channels"] == [32, 64]
    # a real diffusers config.json carries keys the constructor does not take: ignored like ConfigMixin does
    cfg.update({"_name_or_path": "stabilityai/stable-video-diffusion-img2vid", "some_future_key": 1})
    json.dump(cfg, open(root / "unet" / "config.json", "w"))
    dst = U.UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(root), subfolder="unet")
    a, b = src.state_dict(), dst.state_dict()
    assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
    assert dst.config.block_out_channels == (32, 64) and dst.config.num_frames == 4
    # ControlNet from its own directory, the UNet -> ControlNet initialisation of train_svd_controlnet
    cn = U.ControlNetSDVModel.from_unet(dst, conditioning_channels=2)
    cn.save_pretrained(str(root / "controlnet"), safe_serialization=False)
    cn2 = U.ControlNetSDVModel.from_pretrained(str(root), subfolder="controlnet")
    assert cn2.config.conditioning_channels == 2
    assert all(torch.equal(v, cn2.state_dict()[k]) for k, v in cn.state_dict().items())
    with pytest.raises(EnvironmentError, match="local"):
        U.UNetSpatioTemporalConditionControlNetModel.from_pretrained("stabilityai/stable-video-diffusion-img2vid",
                                                                     subfolder="unet")


def test_joint_attention_patch_api_on_cpu():
    """lkgd_b200.patch mirrors the reference's patch/patch.py entry points: parameter names of the joint layers equal the
    reference's (golden from the reference's own initialize_joint_layers), switches land on the blocks, errors are loud."""
    import numpy as np
    from lkgd_b200 import modules as M, patch
    from lkgd_b200.engine import _partners
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    JG = np.load(os.path.join(HERE, "golden", "joint_attention_golden.npz"))
    u = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    patch.apply_patch(u, flip=True, with_temporal_block=True)
    patch.initialize_joint_layers(u, post="conv")
    assert sorted(n for n, _ in u.named_parameters() if "1n" in n) == list(JG["ja/param_names"])
    blocks = [m for m in u.modules() if isinstance(m, (M.BasicTransformerBlock, M.TemporalBasicTransformerBlock))]
    assert len(blocks) == 12 and all(b.patched and b.enable_joint_attention for b in blocks)
    assert all(b.flip == isinstance(b, M.BasicTransformerBlock) for b in blocks)        # flip acts in the spatial block
    assert all(float(b.conv1n.weight.abs().max()) == 0.0 for b in blocks)                # zero-initialised (:152-153)
    patch.set_joint_attention_mask(u, [0, 1, 0, 1])
    patch.set_joint_scale(u, 0.5)
    patch.set_joint_attention(u, False, name_filter="up_blocks")
    assert all(b.joint_scale == 0.5 and b.joint_attn_mask.tolist() == [False, True, False, True] for b in blocks)
    on = [n for n, b in u.named_modules() if isinstance(b, M.BasicTransformerBlock) and b.enable_joint_attention]
    assert on and all("up_blocks" not in n for n in on)
    assert _partners(torch.tensor([0, 1, 0, 1], dtype=torch.bool), 4) == [1, 0, 3, 2]
    assert _partners(torch.tensor([0, 1], dtype=torch.bool), 4) == [2, 3, 0, 1]           # repeat_interleave: x x y y
    with pytest.raises(ValueError, match="half"):
        _partners(torch.tensor([1, 1, 1, 0], dtype=torch.bool), 4)
    patch.update_patch(u, joint_scale=0.25)                                               # :841-853
    assert set(patch.collect_from_patch(u, "joint_scale").values()) == {0.25}             # :856-870
    for prm in u.parameters():
        prm.requires_grad_(False)
    patch.set_joint_layer_requires_grad(u, ["xy_lora"], True)                             # no adapter yet: the post layers
    assert sorted(n for n, prm in u.named_parameters() if prm.requires_grad) == \
        sorted(n for n, _ in u.named_parameters() if n.endswith("conv1n.weight"))
    for prm in u.parameters():
        prm.requires_grad_(True)
    patch.remove_patch(u)
    assert not any(b.patched or b.enable_joint_attention for b in blocks)
    with pytest.raises(NotImplementedError):
        patch.apply_patch(u, single_dir=True)
    with pytest.raises(ValueError, match="Unkown post processing type"):
        blocks[0].initialize_joint_layers(post="mlp")
    # several adapters per layer + per-sample masks (utils/util.py:566-603: y_lora / xy_lora / yx_lora, set_adapters,
    # hack_lora_forward, set_patch_lora_mask)
    cfg_ = dict(r=4, lora_alpha=4, init_lora_weights="gaussian", target_modules=["attn1.to_q", "attn1.to_k"])
    first = u.add_adapter(cfg_, adapter_name="xy_lora")
    second = u.add_adapter(dict(cfg_, r=8), adapter_name="yx_lora")
    assert sorted(first) == sorted(second) and len(first) == 24
    lin = u.get_submodule(first[0])
    assert sorted(lin.lora_A) == ["xy_lora", "yx_lora"] and lin.ranks == {"xy_lora": 4, "yx_lora": 8}
    assert [a[0] for a in lin.adapters()] == ["xy_lora", "yx_lora"] and all(a[4] is None for a in lin.adapters())
    u.set_adapters(["yx_lora"])
    assert [a[0] for a in lin.adapters()] == ["yx_lora"]
    u.set_adapters(["xy_lora", "yx_lora"])
    patch.set_patch_lora_mask(u, "xy_lora", [1, 0])
    patch.set_patch_lora_mask(u, "yx_lora", [1, 1])
    patch.hack_lora_forward(u)
    ads = {a[0]: a[4] for a in lin.adapters()}
    assert ads["xy_lora"].tolist() == [True, False] and ads["yx_lora"] is None          # all-true mask = unmasked
    assert u.lora_mask["xy_lora"].tolist() == [True, False]
    with pytest.raises(ValueError, match="already exists"):
        u.add_adapter(cfg_, adapter_name="xy_lora")
    with pytest.raises(ValueError, match="masked adapter cannot be merged"):
        u.merge_lora()
