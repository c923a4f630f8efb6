"""CPU: the oracle (oracle/) against golden vectors produced by RUNNING THE REFERENCE'S OWN FILES
(tests/golden/make_reference_golden.py: scheduler, UNet wiring, LKGD conditioning, ControlNet, residual injection,
LoRA layer, pipeline CFG loop - with only the un-vendored diffusers/peft/core_qnn pieces shimmed), plus the
reference's parameter-name dumps and the known-answer values of SURVEY.md Appendix C."""
import json
import os

import numpy as np
import pytest
import torch

import oracle as O
from golden_util import (B, F, H, REDUCED4, SCHED, T_STEP, W, fill_seeded_, golden, rel, seeded_tensor, t,
                         unet_inputs, unet_residuals)

HERE = os.path.dirname(os.path.abspath(__file__))
G = golden()


# ------------------------------------------------------------------------------------------------ scheduler
@pytest.mark.parametrize("n", [25, 10])
def test_scheduler_schedule_bit_exact(n):
    s = O.EulerDiscreteScheduler(**SCHED)
    assert np.array_equal(s.sigmas[:8].numpy(), G["sched/init_sigmas_head"])      # the "fix": Karras at __init__
    assert np.array_equal(s.timesteps[:8].numpy(), G["sched/init_timesteps_head"])
    s.set_timesteps(n)
    assert np.array_equal(s.sigmas.numpy(), G[f"sched{n}/sigmas"])
    assert np.array_equal(s.timesteps.numpy(), G[f"sched{n}/timesteps"])
    assert float(s.init_noise_sigma) == float(G[f"sched{n}/init_noise_sigma"])


@pytest.mark.parametrize("n", [25, 10])
def test_scheduler_trajectory_bit_exact(n):
    s = O.EulerDiscreteScheduler(**SCHED)
    s.set_timesteps(n)
    x = seeded_tensor("sched/x0", (1, 3, 4, 8, 8)) * s.init_noise_sigma
    for i, ts in enumerate(s.timesteps):
        xin = s.scale_model_input(x, ts)
        assert np.array_equal(xin.numpy(), G[f"sched{n}/scaled"][i])
        o = s.step(seeded_tensor(f"sched/v{i}", x.shape), ts, x)
        assert np.array_equal(o.prev_sample.numpy(), G[f"sched{n}/traj"][i]), i
        assert np.array_equal(o.pred_original_sample.numpy(), G[f"sched{n}/x0"][i]), i
        x = o.prev_sample


def test_direct_fusion_step_matches_reference_lines():
    """the bidirectional ``direct_fusion`` Euler step against a trajectory produced by executing the
    reference's own lines (tests/golden/make_fusion_golden.py) - bit-exact in fp32."""
    FG = np.load(os.path.join(HERE, "golden", "fusion_golden.npz"))["traj"]
    s = O.EulerDiscreteScheduler(**SCHED)
    s.set_timesteps(10)
    x = seeded_tensor("fusion/x0", (2, 5, 4, 8, 8)) * s.init_noise_sigma
    for i, ts in enumerate(s.timesteps):
        x = s.step_direct_fusion(seeded_tensor(f"fusion/v{i}", x.shape), ts, x)
        assert np.array_equal(x.numpy(), FG[i]), i
    assert s.step_index == 10


def test_scheduler_add_noise_and_errors():
    s = O.EulerDiscreteScheduler(**SCHED)
    orig, noise = seeded_tensor("sched/orig", (4, 2, 4, 4, 4)), seeded_tensor("sched/noise", (4, 2, 4, 4, 4))
    got = s.add_noise(orig, noise, t(G["sched/add_noise_t"]))
    assert np.array_equal(got.numpy(), G["sched/add_noise"])
    s.set_timesteps(25)
    with pytest.raises(ValueError):                 # integer timesteps are rejected (reference :458-469, D6)
        s.step(torch.zeros(1), 3, torch.zeros(1))


def test_scheduler_kat_appendix_c():
    s = O.EulerDiscreteScheduler(**SCHED)
    s.set_timesteps(25)
    sig = s.sigmas.numpy()
    np.testing.assert_allclose(sig[:5], [700.0, 545.729248046875, 421.5691223144531, 322.45367431640625,
                                         244.0230712890625], rtol=1e-6)
    np.testing.assert_allclose(sig[20:], [0.15740464627742767, 0.06639907509088516, 0.02480258047580719,
                                          0.007882495410740376, 0.0020000000949949026, 0.0], rtol=1e-6)
    np.testing.assert_allclose(s.timesteps.numpy()[:5], [1.6377700567245483, 1.575530767440796, 1.5109959840774536,
                                                         1.443989872932434, 1.3743157386779785], rtol=1e-6)
    np.testing.assert_allclose(float(s.init_noise_sigma), 700.000732421875, rtol=1e-7)
    s._step_index = 3
    x, v = torch.full((1,), 1.5), torch.full((1,), -0.25)
    np.testing.assert_allclose(float(s.scale_model_input(x, s.timesteps[3])), 0.004651808645576239, rtol=1e-6)
    o = s.step(v, s.timesteps[3], x)
    np.testing.assert_allclose(float(o.pred_original_sample), 0.25001323223114014, rtol=1e-5)
    np.testing.assert_allclose(float(o.prev_sample), 1.1959649324417114, rtol=1e-6)
    g = O.guidance_ramp(1.0, 3.0, 14).flatten().numpy()
    np.testing.assert_allclose(g[:4], [1.0, 1.1538461446762085, 1.307692289352417, 1.4615384340286255], rtol=1e-6)
    assert g[-1] == 3.0


# ------------------------------------------------------------------------------------------------ structure
def test_param_names_match_reference_dumps():
    d = json.load(open(os.path.join(HERE, "golden", "param_names.json")))
    with torch.device("meta"):
        m = O.UNetSpatioTemporalConditionModel(**dict(O.SVD_XT_CONFIG, num_attention_heads=(5, 10, 20, 20)))
        O.add_lora(m, 4)
    names = {n for n, _ in m.named_parameters()}
    want = set(d["frozen"]) | set(d["trainable"])
    assert names == want, (sorted(names - want)[:5], sorted(want - names)[:5])
    trainable = {n for n in names if "lora_" in n}       # incl. quaternion_lora_* (F10)
    assert trainable == set(d["trainable"])


# ------------------------------------------------------------------------------------------------ UNet wiring
def _oracle_unet():
    return fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)).eval()


def test_unet_forward_matches_reference():
    o = _oracle_unet()
    sample, ctx, ids = unet_inputs()
    with torch.no_grad():
        a = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids, return_dict=False)[0]
        b = o(sample, 0.75, ctx, added_time_ids=ids).sample
    assert rel(a, G["unet/out"]) < 1e-6
    assert rel(b, G["unet/out_float_t"]) < 1e-6
    assert int(G["unet/n_params"]) == sum(p.numel() for p in o.parameters())


PG = np.load(os.path.join(HERE, "golden", "patch_golden.npz"))


@pytest.mark.parametrize("tag,dim,heads,dh,xdim", [("d16", 32, 2, 16, 32), ("d64", 128, 2, 64, 48)])
def test_transformer_blocks_match_the_references_own_forwards(tag, dim, heads, dh, xdim):
    """The oracle's BasicTransformerBlock / TemporalBasicTransformerBlock against vectors produced by executing the
    reference's in-tree restatement of those forwards (patch/patch.py:390-580 and :582-686 through apply_patch, joint
    attention off) on blocks built from plain torch primitives + F.scaled_dot_product_attention - no oracle arithmetic
    on the golden side (tests/golden/make_patch_golden.py)."""
    BF, N, Fr = 8, 24, 4
    x = seeded_tensor(f"patch/{tag}/x", (BF, N, dim))
    sb = fill_seeded_(O.BasicTransformerBlock(dim, heads, dh, xdim), seed=11).eval()
    tb = fill_seeded_(O.TemporalBasicTransformerBlock(dim, dim, heads, dh, xdim), seed=12).eval()
    with torch.no_grad():
        a = sb(x, encoder_hidden_states=seeded_tensor(f"patch/{tag}/ctx", (BF, 1, xdim)))
        a3 = sb(x, encoder_hidden_states=seeded_tensor(f"patch/{tag}/ctx3", (BF, 3, xdim)))
        b = tb(x, num_frames=Fr, encoder_hidden_states=seeded_tensor(f"patch/{tag}/tctx", ((BF // Fr) * N, 1, xdim)))
    assert rel(a, PG[f"patch/spatial_{tag}"]) < 1e-6
    assert rel(a3, PG[f"patch/spatial_{tag}_kv3"]) < 1e-6
    assert rel(b, PG[f"patch/temporal_{tag}"]) < 1e-6


def test_unet_matches_the_reference_unet_with_patched_blocks():
    """Whole reduced UNet: the reference's UNet file with all 12 transformer blocks driven by the reference's own
    patch/patch.py forwards (independent block internals) == the oracle, and == the golden made with oracle blocks."""
    o = _oracle_unet()
    sample, ctx, ids = unet_inputs()
    with torch.no_grad():
        a = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids, return_dict=False)[0]
    assert int(PG["patch/n_blocks"]) == 12
    assert rel(a, PG["patch/unet_out"]) < 2e-6
    assert rel(G["unet/out"], PG["patch/unet_out"]) < 2e-6


JOINT_MASKS = {"alt": ([1, 0, 1, 0], [0, 1, 0, 1]), "pair": ([1, 0], [0, 1]), "yxxx": ([0, 1, 1, 1], [1, 0, 0, 0])}


def joint_inputs():
    sample = seeded_tensor("joint/sample", (4, F, 8, H, W))
    ctx = seeded_tensor("joint/ctx", (4, 1, 32))
    ids = torch.tensor([[6.0, 127.0, 0.02], [6.0, 60.0, 0.02], [6.0, 127.0, 0.1], [12.0, 127.0, 0.02]])
    ts = torch.tensor([1.4439898729, 0.3, 1.4439898729, -0.7])
    return sample, ctx, ids, ts


def test_joint_input_head_unet_matches_reference():
    """the x / y input-head UNet against the reference's own models/unet_spatio_temporal_condition_joint.py
    run through the shim with masks installed by the reference's patch.set_patch_lora_mask
    (tests/golden/make_joint_golden.py): per-sample timesteps and added-time ids, three mask layouts."""
    JG = np.load(os.path.join(HERE, "golden", "joint_golden.npz"))
    o = O.UNetSpatioTemporalConditionJointModel(**REDUCED4)
    o.add_y_input_head()
    o = fill_seeded_(o).eval()
    assert sorted(n for n, _ in o.named_parameters() if "_y." in n) == list(JG["joint/param_names"])
    sample, ctx, ids, ts = joint_inputs()
    for tag, (xy, yx) in JOINT_MASKS.items():
        o.lora_mask = {"xy_lora": torch.tensor(xy, dtype=torch.bool), "yx_lora": torch.tensor(yx, dtype=torch.bool)}
        with torch.no_grad():
            a = o(sample, ts, ctx, added_time_ids=ids, return_dict=False)[0]
        assert rel(a, JG[f"joint/out_{tag}"]) < 1e-6, tag


JA_CASES = {"conv": ("conv", False, True, [0, 1, 0, 1], 1.0), "conv_flip": ("conv", True, False, [0, 1, 0, 1], 0.7),
            "scale_pair": ("scale", False, True, [0, 1], 1.0),      # post, flip, temporal blocks too, mask, joint_scale
            "conv_fuse": ("conv_fuse", False, True, [0, 1, 0, 1], 0.8), "conv_fuse_flip": ("conv_fuse", True, False, [1, 0], 1.0)}


def ja_inputs():
    sample = seeded_tensor("ja/sample", (4, F, 8, H, W))
    ctx = seeded_tensor("ja/ctx", (4, 1, 32))
    return sample, ctx, torch.tensor([[6.0, 127.0, 0.02]] * 4)


@pytest.mark.parametrize("tag", list(JA_CASES))
def test_joint_attention_matches_the_references_patch(tag):
    """the oracle's joint-attention branch (second attention over the partner sample, post layer,
    joint_scale, frame flip; spatial and temporal blocks) against the reference's own patch/patch.py ToMeBlock forwards
    with enable_joint_attention = True, run on torch-primitive blocks (tests/golden/make_joint_attention_golden.py)."""
    JG = np.load(os.path.join(HERE, "golden", "joint_attention_golden.npz"))
    post, flip, temporal, mask, jscale = JA_CASES[tag]
    o = O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    for m in o.modules():
        if isinstance(m, O.BasicTransformerBlock) or (temporal and isinstance(m, O.TemporalBasicTransformerBlock)):
            m.initialize_joint_layers(post)
    o = fill_seeded_(o).eval()
    if tag == "conv":
        assert sorted(n for n, _ in o.named_parameters() if "1n" in n) == list(JG["ja/param_names"])
    for m in o.modules():
        if hasattr(m, "attn1n"):
            m.enable_joint_attention, m.joint_scale, m.flip, m.num_frames = True, jscale, flip, F
            m.joint_attn_mask = torch.tensor(mask, dtype=torch.bool)
    sample, ctx, ids = ja_inputs()
    with torch.no_grad():
        a = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids, return_dict=False)[0]
    assert rel(a, JG[f"ja/out_{tag}"]) < 2e-6
    if tag == "conv":
        for m in o.modules():
            if hasattr(m, "attn1n"):
                m.enable_joint_attention = False
        with torch.no_grad():
            b = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids, return_dict=False)[0]
        assert rel(b, JG["ja/out_off"]) < 2e-6 and rel(a, b) > 5e-2          # the branch matters


def test_masked_multi_adapter_lora_matches_the_references_hacked_forward():
    """patch/patch.py:57-92 (`lora_forward_hack`) on the reference's own LoRA layer with two adapters and per-sample masks
    (tests/golden/make_lora_mask_golden.py): the oracle's LoraLinear reproduces the stock and the masked forward exactly."""
    LG = np.load(os.path.join(HERE, "golden", "lora_mask_golden.npz"))
    l = O.LoraLinear(torch.nn.Linear(32, 48), 4, 4, "gaussian", "xy_lora")
    l.update_layer("yx_lora", 8, 4)
    assert sorted(n for n, _ in l.named_parameters()) == list(LG["loramask/names"])
    l = fill_seeded_(l, seed=3)
    x = seeded_tensor("loramask/x", (8, 5, 32))
    with torch.no_grad():
        assert rel(l(x), LG["loramask/y_unmasked"]) < 1e-6
        l.masked_forward = True
        l.lora_mask = {"xy_lora": torch.tensor([1, 0, 1, 0], dtype=torch.bool),
                       "yx_lora": torch.tensor([0, 1, 0, 1], dtype=torch.bool)}
        assert rel(l(x), LG["loramask/y_masked"]) < 1e-6
        l.lora_mask = {"xy_lora": torch.tensor([1, 1], dtype=torch.bool), "yx_lora": torch.tensor([0, 1], dtype=torch.bool)}
        assert rel(l(x), LG["loramask/y_masked2"]) < 1e-6


@pytest.mark.parametrize("act,heads,hid", [("gelu", 2, 160), ("quick_gelu", 4, 64)])
def test_clip_oracle_matches_transformers(act, heads, hid):
    """SURVEY 8f N1 (CLIP half): the oracle's CLIPVisionModelWithProjection - a restatement of the un-vendored
    transformers class the reference pipelines call (pipeline...controlnet.py:174-214) - against the `transformers` package
    installed in this image, identical state_dict (80- and 16-wide heads, both activations)."""
    tf = pytest.importorskip("transformers")
    cfg = dict(hidden_size=hid, intermediate_size=2 * hid, num_hidden_layers=3, num_attention_heads=heads, image_size=56,
               patch_size=14, num_channels=3, projection_dim=48, hidden_act=act, layer_norm_eps=1e-5)
    torch.manual_seed(0)
    hf = tf.CLIPVisionModelWithProjection(tf.CLIPVisionConfig(**cfg)).eval()
    o = O.CLIPVisionModelWithProjection(**cfg).eval()
    o.load_state_dict({k: v for k, v in hf.state_dict().items() if not k.endswith("position_ids")}, strict=True)
    x = seeded_tensor("clip/x", (2, 3, 56, 56))
    with torch.no_grad():
        assert rel(o(x).image_embeds, hf(pixel_values=x).image_embeds) < 1e-5


def test_image_preprocessing_matches_the_references_function():
    """lkgd_b200/preprocess.py (anti-aliased resize of `_encode_image`) against the reference's own
    `_resize_with_antialiasing` run here (tests/golden/make_preprocess_golden.py; fp16 fixture)."""
    from lkgd_b200.preprocess import clip_pixel_values, resize_with_antialiasing
    PG_ = np.load(os.path.join(HERE, "golden", "preprocess_golden.npz"))
    for tag, shape in {"576x1024": (1, 3, 576, 1024), "224x300": (1, 3, 224, 300)}.items():
        img = seeded_tensor(f"pre/{tag}", shape).sigmoid()
        got = resize_with_antialiasing(img * 2.0 - 1.0, (224, 224))
        assert float((got - t(PG_[f"pre/{tag}"]).float()).abs().max()) < 2e-3
    pv = clip_pixel_values(seeded_tensor("pre/576x1024", (1, 3, 576, 1024)).sigmoid())
    assert tuple(pv.shape) == (1, 3, 224, 224) and abs(float(pv.mean())) < 1.0


def test_flow_stem_unet_matches_reference():
    """the reference's flow-stem UNet (models/unet_spatio_temporal_condition_flow.py, run through the shim
    by tests/golden/make_flow_golden.py) against the oracle restatement; conv_in2 / conv_in2_alpha keep their names."""
    import numpy as np
    import oracle as O
    FG = np.load(os.path.join(HERE, "golden", "flow_golden.npz"))
    o = O.UNetSpatioTemporalConditionModelFlow(**REDUCED4)
    o.initialize_conv_in()
    o = fill_seeded_(o).eval()
    assert sorted(n for n, _ in o.named_parameters() if n.startswith("conv_in")) == list(FG["flow/param_names"])
    sample = seeded_tensor("flow/sample", (B, F, 12, H, W))
    _, ctx, ids = unet_inputs()
    with torch.no_grad():
        a = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids, return_dict=False)[0]
        o.conv_in2_alpha.zero_()
        b = o(sample, torch.tensor(T_STEP), ctx, added_time_ids=ids).sample
    assert rel(a, FG["flow/out"]) < 1e-6
    assert rel(b, FG["flow/out_alpha0"]) < 1e-6
    # a fresh stem (alpha = 0) is the plain 8-channel UNet on (noise | condition)
    base = O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    base.load_state_dict({k: v for k, v in o.state_dict().items() if not k.startswith("conv_in2")})
    with torch.no_grad():
        c = base.eval()(sample[:, :, :8], torch.tensor(T_STEP), ctx, added_time_ids=ids).sample
    assert rel(c, FG["flow/out_alpha0"]) < 1e-6


def test_unet_residual_injection_matches_reference():
    """F6: the residual add sits inside the down-block loop (multipliers (2,2,2,2,1,1) for the 2-level config)."""
    o = _oracle_unet()
    sample, ctx, ids = unet_inputs()
    res, mid = unet_residuals()
    with torch.no_grad():
        a = o(sample, torch.tensor(T_STEP), ctx, down_block_additional_residuals=res,
              mid_block_additional_residual=mid, added_time_ids=ids).sample
    assert rel(a, G["unet/out_residuals"]) < 1e-6
    assert rel(G["unet/out"], G["unet/out_residuals"]) > 1e-2     # the residuals matter


def test_lkgd_conditioning_matches_reference():
    o = fill_seeded_(O.UNetSpatioTemporalConditionModel(**dict(REDUCED4, cross_attention_dim=1024))).eval()
    sample, _, ids = unet_inputs()
    ctx = seeded_tensor("lkgd/ctx", (B, 1, 1024))
    dom, flo = seeded_tensor("lkgd/domain", (1, 1, 1000)), seeded_tensor("lkgd/flow", (1, 1, 1000))
    with torch.no_grad():
        c = o._condition(ctx, dom, flo)
        a = o(sample, torch.tensor(T_STEP), ctx, dom, flo, added_time_ids=ids).sample
        dom2, flo2 = seeded_tensor("lkgd/domain2", (B, 1, 1000)), seeded_tensor("lkgd/flow2", (B, 1, 1000))
        c2 = o._condition(ctx, dom2, flo2)
        a2 = o(sample, torch.tensor(T_STEP), ctx, dom2, flo2, added_time_ids=ids).sample
    assert rel(c.reshape(B, 1024), G["lkgd/context"]) < 1e-5
    assert rel(a, G["lkgd/out"]) < 1e-5
    assert rel(c2.reshape(B, 1024), G["lkgd/context_b2"]) < 1e-5
    assert rel(a2, G["lkgd/out_b2"]) < 1e-5


def test_lkgd_zero_embedding_phase_switch_only_touches_exact_zeros():
    """oracle.UNetSpatioTemporalConditionModel.canonical_zero_phase: for a non-zero embedding the conditioning is
    unchanged bit for bit; for the all-zero embedding of the unconditional CFG half it replaces the CPU FFT's -0 real
    parts (phase pi in 63 bins) by +0 (phase 0, what cuFFT returns) - see the GPU test of the same name family."""
    import oracle as O
    o = fill_seeded_(O.UNetSpatioTemporalConditionModel(**dict(REDUCED4, cross_attention_dim=1024))).eval()
    ctx = torch.cat([torch.zeros(1, 1, 1024), seeded_tensor("lkgd/ctx", (1, 1, 1024))])
    dom, flo = seeded_tensor("lkgd/domain", (1, 1, 1000)), seeded_tensor("lkgd/flow", (1, 1, 1000))
    with torch.no_grad():
        a = o._condition(ctx, dom, flo)
        o.canonical_zero_phase = True
        b = o._condition(ctx, dom, flo)
    assert torch.equal(a[1], b[1])                      # generic spectrum: nothing to canonicalise
    z = torch.fft.rfft(torch.zeros(256))
    if torch.signbit(z.real).any():                     # this torch build's CPU FFT has the artefact
        assert not torch.equal(a[0], b[0])


def test_controlnet_matches_reference():
    cfg = {k: v for k, v in REDUCED4.items() if k != "up_block_types"}
    o = fill_seeded_(O.ControlNetSDVModel(**cfg, conditioning_channels=2), seed=1).eval()
    sample, ctx, ids = unet_inputs()
    cond = seeded_tensor("cn/cond", (B, F, 2, 8 * H, 8 * W)).clamp(-1, 1)
    with torch.no_grad():
        down, mid = o(sample, torch.tensor(T_STEP), ctx, ids, controlnet_cond=cond, conditioning_scale=0.7,
                      return_dict=False)
    assert len(down) == int(G["cn/n_down"]) == 6
    for i, d in enumerate(down):
        assert rel(d, G[f"cn/down{i}"]) < 1e-6, i
    assert rel(mid, G["cn/mid"]) < 1e-6


def test_lora_matches_reference_layer():
    base = torch.nn.Linear(32, 48)
    lo = O.LoraLinear(base, 4, 4, "gaussian")
    fill_seeded_(lo, seed=2)
    x = seeded_tensor("lora/x", (5, 7, 32))
    with torch.no_grad():
        assert rel(lo(x), G["lora/y"]) < 1e-6
        assert rel(lo.get_delta_weight(), G["lora/delta"]) < 1e-6
        lo.merge()
        assert rel(lo.base_layer.weight, G["lora/merged_weight"]) < 1e-6
        assert rel(lo(x), G["lora/y_merged"]) < 1e-6
    lo8 = O.LoraLinear(torch.nn.Linear(32, 48), 8, 4, True)
    assert lo8.scaling == float(G["lora/scaling_r8_a4"]) == 0.5
    assert float(lo8.lora_B["default"].weight.detach().abs().max()) == float(G["lora/B_default_is_zero"]) == 0.0


def test_pipeline_loop_matches_reference():
    """The reference pipeline's own __call__ (CFG dup, scale, concat, ControlNet, UNet with residual injection,
    frame-wise guidance, Euler-Karras step) for 6 steps vs the oracle's denoise_loop."""
    unet = _oracle_unet()
    cfg = {k: v for k, v in REDUCED4.items() if k != "up_block_types"}
    cn = fill_seeded_(O.ControlNetSDVModel(**cfg, conditioning_channels=2), seed=1).eval()
    sched = O.EulerDiscreteScheduler(**SCHED)
    sched.set_timesteps(6)
    lat0 = seeded_tensor("pipe/latents", (1, F, 4, H, W)) * sched.init_noise_sigma
    cond = seeded_tensor("pipe/cond", (F, 2, 8 * H, 8 * W)).clamp(-1, 1).unsqueeze(0)
    cond = torch.cat([cond] * 2)
    ids = O.add_time_ids_inference(6, 127, 0.02, 1)
    final, preds, traj = O.denoise_loop(unet, sched, lat0, t(G["pipe/image_latents"]), t(G["pipe/image_embeddings"]),
                                        ids, 6, 1.0, 3.0, controlnet=cn, controlnet_cond=cond,
                                        controlnet_cond_scale=0.7, return_trajectory=True)
    assert np.array_equal(O.guidance_ramp(1.0, 3.0, F).numpy(), G["pipe/guidance"])
    for i, x in enumerate(traj):
        assert rel(x, G["pipe/steps"][i]) < 1e-5, i
    assert rel(final, G["pipe/final"]) < 1e-5


# ------------------------------------------------------------------------------------------------ algebraic KATs
def test_lora_default_init_is_identity_and_merge_equivalence():
    torch.manual_seed(0)
    lin = torch.nn.Linear(16, 24)
    lo = O.LoraLinear(lin, 4, 4, "gaussian")
    x = torch.randn(3, 16)
    with torch.no_grad():
        assert torch.equal(lo(x), lin(x))                 # B = 0 (reference lora_layer.py:145)
        lo.lora_B["default"].weight.normal_(0, 0.1)
        y = lo(x)
        lo.merge()
        assert rel(lo(x), y) < 1e-6


def test_controlnet_zero_convs_are_identity_for_the_unet():
    o = _oracle_unet()
    cfg = {k: v for k, v in REDUCED4.items() if k != "up_block_types"}
    torch.manual_seed(0)
    cn = O.ControlNetSDVModel(**cfg, conditioning_channels=3).eval()      # default zero-init (controlnet_sdv.py:804-807)
    sample, ctx, ids = unet_inputs()
    with torch.no_grad():
        down, mid = cn(sample, 0.5, ctx, ids, controlnet_cond=torch.rand(B, F, 3, 8 * H, 8 * W), return_dict=False)
        assert all(float(d.abs().max()) == 0.0 for d in down) and float(mid.abs().max()) == 0.0
        a = o(sample, 0.5, ctx, added_time_ids=ids).sample
        b = o(sample, 0.5, ctx, down_block_additional_residuals=down, mid_block_additional_residual=mid,
              added_time_ids=ids).sample
    assert torch.equal(a, b)


def test_residual_multipliers_f6():
    """Constant residuals r_i = i+1 on a UNet whose skips are otherwise unchanged show m = (2,2,2,2,1,1)."""
    from lkgd_b200.engine import residual_multipliers
    assert residual_multipliers(4, [4, 3, 3, 2]) == [4, 4, 4, 4, 3, 3, 3, 2, 2, 2, 1, 1]
    assert residual_multipliers(2, [4, 2]) == [2, 2, 2, 2, 1, 1]
    # the same multipliers out of the oracle's loop: feed unit residuals and look at what reaches the up blocks
    o = _oracle_unet()
    sample, ctx, ids = unet_inputs()
    res, mid = unet_residuals()
    seen = []
    hooks = [r.register_forward_pre_hook(lambda m, a: seen.append(a[0].detach().clone()))
             for blk in o.up_blocks for r in blk.resnets]
    with torch.no_grad():
        o(sample, 0.5, ctx, added_time_ids=ids)
        base = list(seen)
        seen.clear()
        o(sample, 0.5, ctx, down_block_additional_residuals=res, added_time_ids=ids)
    for h in hooks:
        h.remove()
    mult = [2, 2, 2, 2, 1, 1]
    # up-block resnet k consumes skip 5-k as the trailing channels of its concatenated input
    for k, (a, b) in enumerate(zip(base, seen)):
        i = 5 - k
        c = res[i].shape[1]
        assert rel(b[:, -c:] - a[:, -c:], mult[i] * res[i]) < 1e-5, k


def test_kv1_cross_attention_identity():
    torch.manual_seed(0)
    at = O.Attention(32, 2, 16, cross_attention_dim=24).eval()
    x, ctx = torch.randn(3, 10, 32), torch.randn(3, 1, 24)
    with torch.no_grad():
        want = at.to_out[0](at.to_v(ctx)).expand(3, 10, 32)
        assert rel(at(x, ctx), want) < 1e-6           # softmax over one key == 1


def test_cfg_identities():
    p = torch.randn(2, 5, 4, 3, 3)
    assert torch.allclose(O.cfg_combine(p, O.guidance_ramp(1.0, 1.0, 5)), p[1:], atol=1e-6)
    q = torch.cat([p[:1], p[:1]])
    assert torch.allclose(O.cfg_combine(q, O.guidance_ramp(1.0, 3.0, 5)), p[:1])


def test_alphablender_limits_and_single_frame():
    ab = O.AlphaBlender()
    xs, xt = torch.randn(2, 3, 4, 2, 2), torch.randn(2, 3, 4, 2, 2)
    iof = torch.zeros(2, 4)
    with torch.no_grad():
        ab.mix_factor.fill_(50.0)
        assert torch.allclose(ab(xs, xt, iof), xs)
        ab.mix_factor.fill_(-50.0)
        assert torch.allclose(ab(xs, xt, iof), xt)


def test_add_time_ids_orderings_f12():
    assert O.add_time_ids_inference(6, 127, 0.02, 1, do_cfg=False).tolist() == [[6.0, 127.0, pytest.approx(0.02)]]
    assert O.add_time_ids_training(6, 127, 0.02, 1).tolist() == [[6.0, pytest.approx(0.02), 127.0]]


def test_oracle_fp64_self_consistency():
    o = _oracle_unet()
    sample, ctx, ids = unet_inputs()
    with torch.no_grad():
        a = o(sample, 0.5, ctx, added_time_ids=ids).sample
        b = o.double()(sample.double(), 0.5, ctx.double(), added_time_ids=ids.double()).sample
    assert rel(a, b) < 1e-5


def test_vae_state_dict_names():
    """SURVEY 8f N1 (VAE half; oracle/vae.py is a restatement of the un-vendored diffusers class, parity unpinned): the oracle
    and the product module expose the SVD checkpoint's `vae/` parameter names and shapes - 374 tensors, 97.74 M parameters -
    and the structural facts the diffusers class is known by (no post_quant_conv, Conv3d time_conv_out, learned blenders)."""
    from lkgd_b200.vae import AutoencoderKLTemporalDecoder as P
    o = O.AutoencoderKLTemporalDecoder(**O.SVD_VAE_CONFIG)
    p = P(**O.SVD_VAE_CONFIG)
    so, sp = o.state_dict(), p.state_dict()
    assert list(so.keys()) == list(sp.keys()) and all(so[k].shape == sp[k].shape for k in so)
    assert len(so) == 374 and sum(v.numel() for v in so.values()) == 97_742_847
    for k, shape in {"encoder.conv_in.weight": (128, 3, 3, 3), "encoder.down_blocks.1.resnets.0.conv_shortcut.weight": (256, 128, 1, 1),
                     "encoder.down_blocks.2.downsamplers.0.conv.weight": (512, 512, 3, 3),
                     "encoder.mid_block.attentions.0.to_q.weight": (512, 512), "encoder.mid_block.attentions.0.to_out.0.bias": (512,),
                     "encoder.mid_block.attentions.0.group_norm.weight": (512,), "encoder.conv_out.weight": (8, 512, 3, 3),
                     "quant_conv.weight": (8, 8, 1, 1), "decoder.conv_in.weight": (512, 4, 3, 3),
                     "decoder.mid_block.resnets.0.spatial_res_block.conv1.weight": (512, 512, 3, 3),
                     "decoder.mid_block.resnets.1.temporal_res_block.conv2.weight": (512, 512, 3, 1, 1),
                     "decoder.mid_block.resnets.0.time_mixer.mix_factor": (1,),
                     "decoder.mid_block.attentions.0.to_v.bias": (512,),
                     "decoder.up_blocks.2.resnets.0.spatial_res_block.conv_shortcut.weight": (256, 512, 1, 1),
                     "decoder.up_blocks.3.resnets.2.temporal_res_block.norm1.weight": (128,),
                     "decoder.up_blocks.2.upsamplers.0.conv.weight": (256, 256, 3, 3),
                     "decoder.conv_out.weight": (3, 128, 3, 3), "decoder.time_conv_out.weight": (3, 3, 3, 1, 1)}.items():
        assert tuple(so[k].shape) == shape, k
    assert not any(k.startswith("post_quant_conv") or "time_emb_proj" in k or "upsamplers" in k and k.startswith("decoder.up_blocks.3")
                   for k in so)


def test_vae_oracle_semantics():
    """decode mixes frames only inside one `num_frames` clip; the learned blender is switched (alpha -> 1 - alpha); the encoder's
    downsample pads bottom / right; `decode_latents` returns [B, 3, F, H, W]."""
    torch.manual_seed(0)
    v = O.AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 64, 64)).eval()
    z = seeded_tensor("vae/z", (4, 4, 4, 6))
    with torch.no_grad():
        a = v.decode(z, num_frames=2).sample
        b = v.decode(z[:2], num_frames=2).sample
        c = v.decode(z, num_frames=4).sample
        assert torch.allclose(a[:2], b, atol=1e-5) and not torch.allclose(a, c, atol=1e-4)
        for r in v.decoder.mid_block.resnets:
            r.time_mixer.mix_factor.fill_(-30.0)          # sigmoid -> 0, switched alpha -> 1: spatial branch only
        for blk in v.decoder.up_blocks:
            for r in blk.resnets:
                r.time_mixer.mix_factor.fill_(-30.0)
        v.decoder.time_conv_out.weight.zero_()
        v.decoder.time_conv_out.weight[:, :, 1].copy_(torch.eye(3)[..., None, None])
        d2, d4 = v.decode(z, num_frames=2).sample, v.decode(z, num_frames=4).sample
        assert torch.allclose(d2, d4, atol=1e-5)          # no temporal path left
        x = seeded_tensor("vae/x", (1, 32, 6, 6))
        ds = v.encoder.down_blocks[0].downsamplers[0]
        assert torch.allclose(ds(x), torch.nn.functional.conv2d(torch.nn.functional.pad(x, (0, 1, 0, 1)), ds.conv.weight,
                                                                ds.conv.bias, stride=2))
        out = O.decode_latents(v, seeded_tensor("vae/l", (1, 3, 4, 4, 4)), num_frames=3, decode_chunk_size=2)
    assert tuple(out.shape) == (1, 3, 3, 32, 32)


LDM_AUTOENCODER = "torchtitan/experiments/flux/model/autoencoder.py"     # an on-disk CompVis / LDM KL-autoencoder


def _ldm_autoencoder():
    """The image's site-packages hold an INDEPENDENT implementation of the KL-VAE the SVD VAE descends from (torchtitan's Flux
    autoencoder = the CompVis latent-diffusion `Encoder` / `Decoder`: GroupNorm(32, eps 1e-6) + swish ResnetBlocks,
    (0,1,0,1)-padded stride-2 downsampling, single-head mid attention, nearest upsampling).  Loaded by path (no torchtitan
    package import)."""
    import importlib.util
    import sysconfig
    path = os.path.join(sysconfig.get_paths()["purelib"], LDM_AUTOENCODER)
    if not os.path.isfile(path):
        pytest.skip(f"{LDM_AUTOENCODER} is not installed")
    spec = importlib.util.spec_from_file_location("ldm_autoencoder_pin", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _ldm_to_diffusers(name, n_levels, decoder):
    """The published LDM -> diffusers VAE key mapping (diffusers `convert_ldm_vae_checkpoint`)."""
    import re
    name = name.replace("nin_shortcut", "conv_shortcut").replace("norm_out", "conv_norm_out")
    name = re.sub(r"^mid\.block_(\d)\.", lambda m: f"mid_block.resnets.{int(m.group(1)) - 1}.", name)
    name = re.sub(r"^mid\.attn_1\.", "mid_block.attentions.0.", name)
    for a, b in (("attentions.0.norm.", "attentions.0.group_norm."), ("attentions.0.q.", "attentions.0.to_q."),
                 ("attentions.0.k.", "attentions.0.to_k."), ("attentions.0.v.", "attentions.0.to_v."),
                 ("attentions.0.proj_out.", "attentions.0.to_out.0.")):
        name = name.replace(a, b)
    name = re.sub(r"^down\.(\d)\.block\.(\d)\.", r"down_blocks.\1.resnets.\2.", name)
    name = re.sub(r"^down\.(\d)\.downsample\.", r"down_blocks.\1.downsamplers.0.", name)
    name = re.sub(r"^up\.(\d)\.block\.(\d)\.", lambda m: f"up_blocks.{n_levels - 1 - int(m.group(1))}.resnets.{m.group(2)}.", name)
    name = re.sub(r"^up\.(\d)\.upsample\.", lambda m: f"up_blocks.{n_levels - 1 - int(m.group(1))}.upsamplers.0.", name)
    return name


def test_vae_oracle_spatial_skeleton_matches_an_independent_ldm_autoencoder():
    """diffusers is not installed, but the KL-autoencoder the SVD VAE was initialised from is on disk in an
    independent implementation.  With the published LDM -> diffusers key mapping the oracle's ENCODER must reproduce it
    exactly (ResnetBlock eps / activation / shortcut, bottom-right padded downsampling, the single-head attention block and
    its scale, conv_norm_out / conv_out), and the oracle's TEMPORAL DECODER with its temporal branch switched off (mix_factor
    -> -inf, i.e. switched alpha = 1; identity time_conv_out) must reproduce the LDM decoder frame by frame (block order,
    channel wiring, upsampler placement).  What stays restated-only for the VAE: the temporal branch itself
    (TemporalResnetBlock / AlphaBlender with switch_spatial_to_temporal_mix / Conv3d time_conv_out) and quant_conv."""
    L = _ldm_autoencoder()
    boc, z = (32, 64, 128, 128), 4
    kw = dict(resolution=64, in_channels=3, ch=boc[0], ch_mult=[c // boc[0] for c in boc], num_res_blocks=2, z_channels=z)
    torch.manual_seed(0)
    enc = fill_seeded_(L.Encoder(**kw)).eval()
    dec = fill_seeded_(L.Decoder(out_ch=3, **kw), seed=1).eval()
    o = fill_seeded_(O.AutoencoderKLTemporalDecoder(block_out_channels=boc, latent_channels=z), seed=2).eval()
    sd = o.state_dict()
    n_mapped = 0
    for part, ref in (("encoder", enc), ("decoder", dec)):
        for k, v in ref.state_dict().items():
            name = _ldm_to_diffusers(k, len(boc), part == "decoder")
            if part == "decoder" and (".resnets." in name) and "mid_block.attentions" not in name:
                name = name.replace(".resnets.", ".resnets.").replace(".norm1", ".spatial_res_block.norm1") \
                    .replace(".conv1", ".spatial_res_block.conv1").replace(".norm2", ".spatial_res_block.norm2") \
                    .replace(".conv2", ".spatial_res_block.conv2").replace(".conv_shortcut", ".spatial_res_block.conv_shortcut")
            key = f"{part}.{name}"
            assert key in sd, (k, key)
            if v.ndim == 4 and sd[key].ndim == 2:          # 1x1 convs of the LDM attention block are Linears in diffusers
                v = v[:, :, 0, 0]
            assert sd[key].shape == v.shape, (key, sd[key].shape, v.shape)
            sd[key] = v.clone()
            n_mapped += 1
    assert n_mapped == len(enc.state_dict()) + len(dec.state_dict())
    # every encoder tensor of the oracle is now the LDM's; of the decoder everything except the temporal branch
    assert sum(k.startswith("encoder.") for k in sd) == len(enc.state_dict())
    assert sum(k.startswith("decoder.") and "temporal_res_block" not in k and "time_mixer" not in k and "time_conv_out" not in k
               for k in sd) == len(dec.state_dict())
    for k in sd:
        if k.endswith("time_mixer.mix_factor"):
            sd[k] = torch.full_like(sd[k], -40.0)          # sigmoid -> 0; switched blender: alpha = 1 -> spatial branch only
    sd["decoder.time_conv_out.weight"] = torch.zeros_like(sd["decoder.time_conv_out.weight"])
    sd["decoder.time_conv_out.weight"][:, :, 1, 0, 0] = torch.eye(3)
    sd["decoder.time_conv_out.bias"] = torch.zeros(3)
    sd["quant_conv.weight"] = torch.eye(2 * z)[:, :, None, None].clone()
    sd["quant_conv.bias"] = torch.zeros(2 * z)
    o.load_state_dict(sd, strict=True)
    x = torch.tanh(seeded_tensor("ldm/x", (2, 3, 64, 48)))
    zz = seeded_tensor("ldm/z", (6, z, 8, 6))
    with torch.no_grad():
        ref_m, ref_d = enc(x), dec(zz)
        got_m = o.encode(x).latent_dist
        got_d = o.decode(zz, num_frames=3).sample
    assert rel(torch.cat([got_m.mean, got_m.logvar], 1), ref_m) < 2e-6
    assert rel(got_d, ref_d) < 1e-5          # 14 blocks, each blended as 1.0 * spatial + 4e-18 * temporal in fp32


def test_timestep_embedding_matches_an_independent_implementation():
    """`Timesteps(C, flip_sin_to_cos=True, downscale_freq_shift=0)` (restated from diffusers) against the
    sinusoidal embedding of torchtitan's Flux layers (same image, independent code): cos | sin halves over exp(-ln(1e4) k / half)."""
    import importlib.util
    import sysconfig
    path = os.path.join(sysconfig.get_paths()["purelib"], "torchtitan/experiments/flux/model/layers.py")
    if not os.path.isfile(path):
        pytest.skip("torchtitan's flux layers are not installed")
    src = open(path).read()
    start = src.index("def timestep_embedding(")
    ns = {"math": __import__("math"), "torch": torch, "Tensor": torch.Tensor}
    exec(src[start:src.index("\nclass ", start)], ns)                      # the function only (the module imports more)
    ts = torch.tensor([0.0, 0.25 * float(np.log(0.002)), 1.6377, 999.0])
    for dim in (256, 320):
        ref = ns["timestep_embedding"](ts, dim, time_factor=1.0)
        assert rel(O.timestep_embedding(ts, dim), ref) < 1e-6


def test_quaternion_linear_is_the_hamilton_product():
    """SURVEY A.7 / U5 (core_qnn is un-vendored; the block signs were recalled): with 1x1 blocks the layer must be the Hamilton
    product  weight (x) input  of quaternion algebra.  Checked against scipy's independent quaternion composition
    (`Rotation`: unit quaternions, scalar-last, q and -q identified) and against the defining identities i^2 = j^2 = k^2 =
    ijk = -1.  A single wrong sign in the 4x4 block pattern breaks both."""
    from scipy.spatial.transform import Rotation
    from oracle.unet import QuaternionLinear
    q = QuaternionLinear(4, 4)
    g = torch.Generator().manual_seed(0)
    for _ in range(8):
        w = torch.randn(4, generator=g, dtype=torch.float64)
        x = torch.randn(4, generator=g, dtype=torch.float64)
        w, x = w / w.norm(), x / x.norm()
        with torch.no_grad():
            for n_, v in zip(("r_weight", "i_weight", "j_weight", "k_weight"), w):
                getattr(q, n_).copy_(v.reshape(1, 1).float())
            q.bias.zero_()
            y = q(x.float()[None])[0].double()                          # (r, i, j, k)
        as_xyzw = lambda t: np.array([t[1], t[2], t[3], t[0]], dtype=np.float64)
        ref = (Rotation.from_quat(as_xyzw(w)) * Rotation.from_quat(as_xyzw(x))).as_quat()      # w (x) x, scalar last
        got = as_xyzw(y.numpy())
        assert min(np.abs(got - ref).max(), np.abs(got + ref).max()) < 1e-6
    unit = {"1": [1, 0, 0, 0], "i": [0, 1, 0, 0], "j": [0, 0, 1, 0], "k": [0, 0, 0, 1]}

    def mul(a, b):
        with torch.no_grad():
            for n_, v in zip(("r_weight", "i_weight", "j_weight", "k_weight"), a):
                getattr(q, n_).fill_(float(v))
            return q(torch.tensor([b], dtype=torch.float32))[0].tolist()
    assert mul(unit["i"], unit["i"]) == [-1, 0, 0, 0] and mul(unit["j"], unit["j"]) == [-1, 0, 0, 0]
    assert mul(unit["k"], unit["k"]) == [-1, 0, 0, 0]
    assert mul(unit["i"], unit["j"]) == unit["k"] and mul(unit["j"], unit["k"]) == unit["i"] and mul(unit["k"], unit["i"]) == unit["j"]
    assert mul(unit["j"], unit["i"]) == [0, 0, 0, -1]
