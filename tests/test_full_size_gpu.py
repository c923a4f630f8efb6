"""FULL-SIZE parity inside `pytest -m gpu`: the CUDA path at BASELINE.json's own sizes (72x128 latents, 25 / 14 frames,
CFG batch 2) against the fp32 CPU oracle's output of the same name-seeded weights and inputs, committed as fp16
fixtures by tests/golden/make_full_size_golden.py (the oracle forward costs minutes of CPU; the GPU box only loads its
result).  Tolerance: BASELINE.json's per-step prediction rel-L2 <= 1e-2."""
import json
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_full_size_golden as FS  # noqa: E402  (case table + name-seeded inputs; imports the oracle package only)
from weights import fill_seeded_  # noqa: E402

pytestmark = pytest.mark.gpu


def _fixture(name):
    path = os.path.join(HERE, "golden", f"full_size_{name}.npz")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run tests/golden/make_full_size_golden.py {name}")
    return np.load(path)


def _product(name, cuda):
    from lkgd_b200.unet import (ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel,
                                UNetSpatioTemporalConditionModel)
    c = FS.CASES[name]
    cfg = FS.unet_config(name)
    cls = UNetSpatioTemporalConditionModel if c["lkgd"] else UNetSpatioTemporalConditionControlNetModel
    u = cls(**cfg)
    if c["rank"]:
        u.add_lora(c["rank"])
    u = fill_seeded_(u).to(cuda)
    cn = None
    if c["controlnet"]:
        cn = ControlNetSDVModel(**{k: v for k, v in cfg.items() if k != "up_block_types"},
                                conditioning_channels=c["controlnet"])
        cn = fill_seeded_(cn, seed=1).to(cuda)
    return u, cn


def _run(name, cuda):
    c = FS.CASES[name]
    u, cn = _product(name, cuda)
    x, ctx, ids, extra, cond = FS.case_inputs(name)
    xc, cc, ic = x.to(cuda), ctx.to(cuda), ids.to(cuda)
    kw = {}
    mid = None
    if cn is not None:
        down, mid = cn(xc, c["t"], cc, ic, controlnet_cond=cond.to(cuda), conditioning_scale=1.0, return_dict=False,
                       output_layout="nhwc")
        kw = dict(down_block_additional_residuals=down, mid_block_additional_residual=mid)
    out = u(xc, c["t"], cc, *(e.to(cuda) for e in extra), added_time_ids=ic, return_dict=False, **kw)[0]
    torch.cuda.synchronize()
    return out, mid


@pytest.mark.parametrize("name", ["c3_tiny", "c4_tiny"])
def test_fixture_recipe_small(cuda, name):
    """The same recipe at 8x8 latents (seconds on the CPU): guards the fixture plumbing itself."""
    G = _fixture(name)
    out, _ = _run(name, cuda)
    err = rel_l2(out, torch.from_numpy(G["out"].astype(np.float32)))
    print(name, "rel-L2", err)
    assert err < 1e-2


def test_full_size_c3_lkgd_lora_25_frames(cuda):
    """BASELINE.json configs[2] = the headline configuration: SVD-XT LKGD UNet, LoRA r=64, 25 frames, 72x128 latents."""
    G = _fixture("c3")
    out, _ = _run("c3", cuda)
    assert tuple(out.shape) == (2, 25, 4, 72, 128) and bool(torch.isfinite(out).all())
    ref = torch.from_numpy(G["out"].astype(np.float32))
    err = rel_l2(out, ref)
    print("C3 full-size rel-L2", err, "oracle:", json.loads(str(G["meta"])))
    assert err < 1e-2
    for b in range(2):       # both CFG halves individually (the unconditional half carries the zero embedding)
        assert rel_l2(out[b], ref[b]) < 1e-2, b


def test_full_size_c4_controlnet_14_frames(cuda):
    """BASELINE.json configs[3]: SVD 14 frames, ControlNetSDVModel (2-channel flow condition at 576x1024) -> 12 + 1
    residuals -> F6 residual injection in the UNet."""
    G = _fixture("c4")
    out, mid = _run("c4", cuda)
    assert tuple(out.shape) == (2, 14, 4, 72, 128) and bool(torch.isfinite(out).all())
    err_mid = rel_l2(mid.to_nchw()[:, :128], torch.from_numpy(G["cn_mid"].astype(np.float32)))
    err = rel_l2(out, torch.from_numpy(G["out"].astype(np.float32)))
    print("C4 full-size rel-L2: unet", err, "controlnet mid residual", err_mid, json.loads(str(G["meta"])))
    assert err_mid < 1.5e-2
    assert err < 1e-2


def test_full_size_c5_training_step(cuda):
    """BASELINE.json configs[4]: ONE full-size LoRA fine-tuning step (SVD-XT-width LKGD UNet, r = 64, 14 frames of 40x64 latents,
    batch 1): loss and gradients of the hand-scheduled CUDA backward against autograd through the fp32 oracle
    (tests/golden/make_train_golden.py: loss, the gradient norm of all 125 trainable tensors, eight complete gradient tensors).
    Tolerances as in tests/test_training_gpu.py: loss 1e-2 relative, every tensor 5e-2, all norms together 3e-2."""
    import make_train_golden as TG
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import UNetSpatioTemporalConditionModel
    G = _fixture("c5")
    u = UNetSpatioTemporalConditionModel(**TG.config())
    u.add_lora(TG.RANK)
    u = fill_seeded_(u).to(cuda)
    lat, noise, cond, ctx, sig, ids, extra = TG.inputs()
    tr = LoraTrainer(u)
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda),
                               *(e.to(cuda) for e in extra))
    got = dict(tr.named_grads())
    names = [str(n) for n in G["names"]]
    assert sorted(got) == sorted(names) and len(names) == 125
    ref_loss = float(G["loss"])
    assert abs(float(loss) - ref_loss) < 1e-2 * abs(ref_loss), (float(loss), ref_loss)
    norms = {n: float(v) for n, v in zip(names, G["norms"])}
    worst = max((abs(float(got[n].double().norm()) / norms[n] - 1.0), n) for n in names if norms[n] > 1e-12)
    num = den = 0.0
    worst_full = (0.0, "")
    for key in G.files:
        if not key.startswith("grad/"):
            continue
        n = key[5:]
        ref = torch.from_numpy(G[key].astype(np.float32))
        e = rel_l2(got[n], ref)
        num += float((got[n].double().cpu() - ref.double()).pow(2).sum())
        den += float(ref.double().pow(2).sum())
        worst_full = max(worst_full, (e, n))
    print(f"C5 full-size: loss {float(loss):.6f} vs {ref_loss:.6f}; worst gradient-norm deviation {worst[0]:.3e} ({worst[1]}); "
          f"complete tensors: all {((num / den) ** 0.5):.3e}, worst {worst_full[0]:.3e} ({worst_full[1]})",
          json.loads(str(G["meta"])))
    assert worst[0] < 5e-2, worst
    assert worst_full[0] < 5e-2 and (num / den) ** 0.5 < 3e-2
