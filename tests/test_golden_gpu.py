"""GPU: the CUDA path (through the C ABI) against golden vectors produced by running the reference's own files
(tests/golden/make_reference_golden.py).  Tolerances are BASELINE.json's: bf16 prediction rel-L2 <= 1e-2,
fp32 scheduler <= 1e-4, trajectory <= 2e-2."""
import numpy as np
import pytest
import torch

from golden_util import (B, F, H, REDUCED4, SCHED, T_STEP, W, fill_seeded_, golden, rel, seeded_tensor, t,
                         unet_inputs, unet_residuals)

pytestmark = pytest.mark.gpu
G = golden()


@pytest.mark.parametrize("n", [25, 10])
def test_scheduler_kernels_vs_reference_trajectory(cuda, n):
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    s = EulerDiscreteScheduler(**SCHED)
    s.set_timesteps(n, device=cuda)
    x = (seeded_tensor("sched/x0", (1, 3, 4, 8, 8)) * s.init_noise_sigma).to(cuda)
    worst = 0.0
    for i, ts in enumerate(s.timesteps):
        xin = s.scale_model_input(x, ts)
        assert rel(xin, G[f"sched{n}/scaled"][i]) < 1e-6
        o = s.step(seeded_tensor(f"sched/v{i}", x.shape).to(cuda), ts, x)
        x = o.prev_sample
        worst = max(worst, rel(x, G[f"sched{n}/traj"][i]))
        worst = max(worst, rel(o.pred_original_sample, G[f"sched{n}/x0"][i]))     # reference :506,:527
    print("scheduler worst rel-L2", worst)
    assert worst < 1e-4
    assert s.step_index == n


def test_scheduler_step_accepts_any_shape(cuda):
    """The reference's step is elementwise on any shape (4-D image latents, 2-D ...): same numbers as the 5-D call."""
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    outs = []
    for shape in ((1, 3, 4, 8, 8), (3, 4, 8, 8), (12, 64), (768,)):
        s = EulerDiscreteScheduler(**SCHED)
        s.set_timesteps(10, device=cuda)
        x = (seeded_tensor("sched/x0", (1, 3, 4, 8, 8)) * s.init_noise_sigma).to(cuda).reshape(shape)
        v = seeded_tensor("sched/v0", (1, 3, 4, 8, 8)).to(cuda).reshape(shape)
        s.scale_model_input(x, s.timesteps[0])
        o = s.step(v, s.timesteps[0], x)
        assert o.prev_sample.shape == x.shape and o.pred_original_sample.shape == x.shape
        outs.append(o.prev_sample.flatten())
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    assert rel(outs[0], G["sched10/traj"][0]) < 1e-4


def test_direct_fusion_step_vs_reference_lines(cuda):
    """lkgd_fusion_euler_step (one kernel) against the trajectory of the reference's own direct_fusion
    lines (tests/golden/make_fusion_golden.py); fp32 tolerance of the scheduler: 1e-4 (observed ~1e-7)."""
    import os
    from golden_util import HERE
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    FG = np.load(os.path.join(HERE, "golden", "fusion_golden.npz"))["traj"]
    s = EulerDiscreteScheduler(**SCHED)
    s.set_timesteps(10, device=cuda)
    x = (seeded_tensor("fusion/x0", (2, 5, 4, 8, 8)) * s.init_noise_sigma).to(cuda)
    worst = 0.0
    for i, ts in enumerate(s.timesteps):
        x = s.step_direct_fusion(seeded_tensor(f"fusion/v{i}", x.shape).to(cuda), ts, x)
        worst = max(worst, rel(x, FG[i]))
    print("direct-fusion worst rel-L2", worst)
    assert worst < 1e-4 and s.step_index == 10


def _product_unet(cuda, cls=None, cfg=REDUCED4, seed=0):
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    cls = cls or UNetSpatioTemporalConditionControlNetModel
    return fill_seeded_(cls(**cfg), seed=seed).to(cuda)


def test_unet_vs_reference_with_patched_blocks(cuda):
    """CUDA path vs the reference UNet whose transformer blocks ran the reference's own forwards (patch/patch.py:390-686
    via apply_patch, joint attention off) on torch-primitive blocks - a golden with no oracle block arithmetic in the
    transformer blocks (tests/golden/make_patch_golden.py)."""
    import os
    from golden_util import HERE
    PG = np.load(os.path.join(HERE, "golden", "patch_golden.npz"))
    p = _product_unet(cuda)
    sample, ctx, ids = (v.to(cuda) for v in unet_inputs())
    a = p(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0]
    err = rel(a, PG["patch/unet_out"])
    print("rel-L2 vs patched-reference UNet", err)
    assert err < 1e-2


def test_unet_vs_reference(cuda):
    p = _product_unet(cuda)
    sample, ctx, ids = (v.to(cuda) for v in unet_inputs())
    a = p(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0]
    b = p(sample, 0.75, ctx, added_time_ids=ids).sample
    ea, eb = rel(a, G["unet/out"]), rel(b, G["unet/out_float_t"])
    print("unet rel-L2", ea, eb)
    assert ea < 1e-2 and eb < 1e-2


def test_flow_stem_unet_vs_reference(cuda):
    """the flow-stem UNet (two gated stems merged into one 12-channel implicit-GEMM conv at pack time)
    against the reference's own models/unet_spatio_temporal_condition_flow.py run through the shim."""
    import os
    from golden_util import HERE
    from lkgd_b200.unet import UNetSpatioTemporalConditionModelFlow
    FG = np.load(os.path.join(HERE, "golden", "flow_golden.npz"))
    p = UNetSpatioTemporalConditionModelFlow(**REDUCED4)
    p.initialize_conv_in()
    p = fill_seeded_(p).to(cuda)
    assert sorted(n for n, _ in p.named_parameters() if n.startswith("conv_in")) == list(FG["flow/param_names"])
    sample = seeded_tensor("flow/sample", (B, F, 12, H, W)).to(cuda)
    _, ctx, ids = (v.to(cuda) for v in unet_inputs())
    a = p(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0]
    with torch.no_grad():
        p.conv_in2_alpha.zero_()
    p.invalidate()
    b = p(sample, T_STEP, ctx, added_time_ids=ids).sample
    ea, eb = rel(a, FG["flow/out"]), rel(b, FG["flow/out_alpha0"])
    print("flow-stem unet rel-L2", ea, eb)
    assert ea < 1e-2 and eb < 1e-2
    with pytest.raises(ValueError):
        p(sample[:, :, :8], T_STEP, ctx, added_time_ids=ids)


def test_unet_residual_injection_vs_reference(cuda):
    p = _product_unet(cuda)
    sample, ctx, ids = (v.to(cuda) for v in unet_inputs())
    res, mid = unet_residuals()
    a = p(sample, T_STEP, ctx, down_block_additional_residuals=[r.to(cuda) for r in res],
          mid_block_additional_residual=mid.to(cuda), added_time_ids=ids).sample
    e = rel(a, G["unet/out_residuals"])
    print("unet+residuals rel-L2", e)
    assert e < 1e-2


def test_lkgd_vs_reference(cuda):
    from lkgd_b200.unet import UNetSpatioTemporalConditionModel
    p = _product_unet(cuda, UNetSpatioTemporalConditionModel, dict(REDUCED4, cross_attention_dim=1024))
    sample, _, ids = (v.to(cuda) for v in unet_inputs())
    ctx = seeded_tensor("lkgd/ctx", (B, 1, 1024)).to(cuda)
    dom, flo = seeded_tensor("lkgd/domain", (1, 1, 1000)).to(cuda), seeded_tensor("lkgd/flow", (1, 1, 1000)).to(cuda)
    c = p._context(ctx, dom, flo)
    assert rel(c.reshape(B, 1024), G["lkgd/context"]) < 1e-4            # fp32 conditioning block
    a = p(sample, T_STEP, ctx, dom, flo, added_time_ids=ids).sample     # D8: batch-1 features duplicated
    dom2 = seeded_tensor("lkgd/domain2", (B, 1, 1000)).to(cuda)
    flo2 = seeded_tensor("lkgd/flow2", (B, 1, 1000)).to(cuda)
    assert rel(p._context(ctx, dom2, flo2).reshape(B, 1024), G["lkgd/context_b2"]) < 1e-4
    a2 = p(sample, T_STEP, ctx, dom2, flo2, added_time_ids=ids).sample
    e, e2 = rel(a, G["lkgd/out"]), rel(a2, G["lkgd/out_b2"])
    print("lkgd rel-L2", e, e2)
    assert e < 1e-2 and e2 < 1e-2


def _product_controlnet(cuda):
    from lkgd_b200.unet import ControlNetSDVModel
    cfg = {k: v for k, v in REDUCED4.items() if k != "up_block_types"}
    return fill_seeded_(ControlNetSDVModel(**cfg, conditioning_channels=2), seed=1).to(cuda)


def test_controlnet_vs_reference(cuda):
    cn = _product_controlnet(cuda)
    sample, ctx, ids = (v.to(cuda) for v in unet_inputs())
    cond = seeded_tensor("cn/cond", (B, F, 2, 8 * H, 8 * W)).clamp(-1, 1).to(cuda)
    down, mid = cn(sample, T_STEP, ctx, ids, controlnet_cond=cond, conditioning_scale=0.7, return_dict=False)
    assert len(down) == 6
    errs = [rel(d, G[f"cn/down{i}"]) for i, d in enumerate(down)] + [rel(mid, G["cn/mid"])]
    print("controlnet rel-L2", errs)
    assert max(errs) < 1e-2


def test_lora_folded_and_merged_vs_reference(cuda):
    """models/lora_layer.py Linear.forward / merge, executed as the GEMM's second K segment and as W += s*B*A."""
    from lkgd_b200 import modules as M
    from lkgd_b200 import engine, ops
    base = M.Linear(32, 48)
    lo = fill_seeded_(M.LoraLinear(base, 4, 4, "gaussian", "default"), seed=2).to(cuda)
    x = seeded_tensor("lora/x", (5, 7, 32)).reshape(35, 32).to(cuda).to(torch.bfloat16)
    d = engine._dense(lo, fold_lora=True)
    y = engine.dense(x, d, out_f32=True)
    assert rel(y.reshape(5, 7, 48), G["lora/y"]) < 1e-2
    dm = engine._dense(lo, fold_lora=False)
    assert rel(dm.w.float(), G["lora/merged_weight"]) < 5e-3            # bf16 storage of the merged weight
    assert rel(engine.dense(x, dm, out_f32=True).reshape(5, 7, 48), G["lora/y_merged"]) < 1e-2


def test_pipeline_loop_vs_reference(cuda):
    """6 CFG Euler-Karras steps with ControlNet residual injection vs the latents the reference pipeline produced."""
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    pipe = StableVideoDiffusionPipeline(_product_unet(cuda), EulerDiscreteScheduler(**SCHED), _product_controlnet(cuda))
    cond = seeded_tensor("pipe/cond", (F, 2, 8 * H, 8 * W)).clamp(-1, 1)
    lat0 = seeded_tensor("pipe/latents", (1, F, 4, H, W))
    final, preds, traj = pipe(t(G["pipe/image_embeddings"]), t(G["pipe/image_latents"]), num_frames=F,
                              num_inference_steps=6, fps=7, motion_bucket_id=127, noise_aug_strength=0.02,
                              latents=lat0, controlnet_condition=cond, controlnet_cond_scale=0.7,
                              return_trajectory=True)
    assert np.allclose(pipe.guidance_scale.cpu().numpy(), G["pipe/guidance"])
    errs = [rel(x, G["pipe/steps"][i]) for i, x in enumerate(traj)]
    print("pipeline per-step latents rel-L2", errs)
    assert max(errs) < 2e-2 and rel(final, G["pipe/final"]) < 2e-2


def test_joint_input_head_unet_vs_reference(cuda):
    """UNetSpatioTemporalConditionJointModel (x / y input heads chosen per sample) against the reference's
    own file run through the shim (tests/golden/make_joint_golden.py); per-sample timesteps / added-time ids."""
    import os
    from golden_util import HERE
    from test_oracle_golden import JOINT_MASKS, joint_inputs
    from lkgd_b200.unet import UNetSpatioTemporalConditionJointModel
    JG = np.load(os.path.join(HERE, "golden", "joint_golden.npz"))
    u = UNetSpatioTemporalConditionJointModel(**REDUCED4)
    u.add_y_input_head()
    u = fill_seeded_(u).to(cuda)
    assert sorted(n for n, _ in u.named_parameters() if "_y." in n) == list(JG["joint/param_names"])
    sample, ctx, ids, ts = (v.to(cuda) for v in joint_inputs())
    for tag, (xy, yx) in JOINT_MASKS.items():
        u.set_lora_mask("xy_lora", xy)
        u.set_lora_mask("yx_lora", yx)
        out = u(sample, ts, ctx, added_time_ids=ids, return_dict=False)[0]
        err = rel(out, JG[f"joint/out_{tag}"])
        print("joint", tag, err)
        assert err < 1e-2, tag
    u.set_lora_mask("xy_lora", [1, 1, 1, 1])
    u.set_lora_mask("yx_lora", [0, 0, 0, 0])
    with pytest.raises(ValueError, match="at least one sample"):
        u(sample, ts, ctx, added_time_ids=ids)


@pytest.mark.parametrize("tag", ["conv", "conv_flip", "scale_pair", "conv_fuse", "conv_fuse_flip"])
def test_joint_attention_patch_vs_reference(cuda, tag):
    """lkgd_b200.patch (apply_patch / initialize_joint_layers / set_joint_attention_mask / set_joint_scale)
    + the engine's joint-attention branch against the reference's own patch/patch.py forwards with joint attention ON
    (tests/golden/make_joint_attention_golden.py)."""
    import os
    from golden_util import HERE
    from test_oracle_golden import JA_CASES, ja_inputs
    from lkgd_b200 import patch
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    JG = np.load(os.path.join(HERE, "golden", "joint_attention_golden.npz"))
    post, flip, temporal, mask, jscale = JA_CASES[tag]
    u = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    patch.apply_patch(u, flip=flip, with_spatial_block=True, with_temporal_block=temporal)
    patch.initialize_joint_layers(u, post=post)
    u = fill_seeded_(u).to(cuda)
    if tag == "conv":
        assert sorted(n for n, _ in u.named_parameters() if "1n" in n) == list(JG["ja/param_names"])
    patch.set_joint_attention_mask(u, mask)
    patch.set_joint_scale(u, jscale)
    sample, ctx, ids = (v.to(cuda) for v in ja_inputs())
    out = u(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0]
    err = rel(out, JG[f"ja/out_{tag}"])
    print("joint attention", tag, err)
    assert err < 1e-2
    if tag == "conv":
        patch.set_joint_attention(u, False)
        off = u(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0]
        assert rel(off, JG["ja/out_off"]) < 1e-2 and rel(out, off) > 5e-2
        patch.remove_patch(u)
        assert rel(u(sample, torch.tensor(T_STEP, device=cuda), ctx, added_time_ids=ids, return_dict=False)[0],
                   JG["ja/out_off"]) < 1e-2
