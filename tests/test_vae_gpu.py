"""SURVEY 8f N1 (VAE half): lkgd_b200.vae.AutoencoderKLTemporalDecoder against the CPU oracle (oracle/vae.py, a restatement
of diffusers 0.27.2's class - un-vendored, parity unpinned for the VAE-specific assembly, see its header) on small
configurations, against the committed full-size fixture (576x1024 frames), and the three kernels added for it.

Tolerance: the VAE's two chains are deep and have no long skip connections (encoder 11 resnets + attention, decoder 14
spatio-temporal resblocks = 56 convs), so the bf16 rounding of every conv operand accumulates block by block
(`tools/vae_trace.py`, profiles/r02k_vae_stage_errors.txt: +0.5e-3 ... +1.7e-3 rel-L2 per block, smoothly, no stage stands
out; narrow 32-channel blocks are the noisiest).  Small configurations: <= 2.5e-2 (measured 0.9e-2 ... 1.8e-2); the SVD VAE's own
widths at 576x1024: <= 1.2e-2 (measured 0.86e-2 encode, 0.91e-2 decode)."""
TOL_SMALL, TOL_FULL = 2.5e-2, 1.2e-2
import os
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_vae_golden as VG  # noqa: E402  (name-seeded inputs; imports the oracle package only)
from weights import seeded_tensor  # noqa: E402

pytestmark = pytest.mark.gpu
bf16 = torch.bfloat16


# ------------------------------------------------------------------------------------------------- kernels
@pytest.mark.parametrize("M,N,ld", [(64, 96, 96), (300, 9216, 9216), (17, 1000, 1024), (5, 16384, 16384)])
def test_softmax_rows(cuda, M, N, ld):
    from lkgd_b200 import ops
    x = (seeded_tensor(f"sm/{M}/{N}", (M, ld)) * 3).to(cuda)
    got = ops.softmax_rows(x[:, :N], 0.37)
    ref = torch.softmax(x[:, :N].double() * 0.37, -1)
    assert got.dtype == bf16 and tuple(got.shape) == (M, N)
    assert rel_l2(got.float(), ref) < 3e-3
    assert float((got.float().sum(-1) - 1).abs().max()) < 2e-2
    with pytest.raises(Exception):
        ops.softmax_rows(x[:, :N - 1].contiguous(), 1.0)


@pytest.mark.parametrize("NB,Fr,H,W,Cn,ld", [(1, 5, 8, 12, 3, 4), (2, 3, 16, 8, 3, 8), (1, 1, 8, 8, 4, 4), (1, 8, 24, 40, 3, 3)])
def test_time_conv_out(cuda, NB, Fr, H, W, Cn, ld):
    from lkgd_b200 import ops
    x = seeded_tensor("tco/x", (NB * Fr * H * W, ld)).to(cuda)
    w = seeded_tensor("tco/w", (Cn, Cn, 3)).to(cuda)
    b = seeded_tensor("tco/b", (Cn,)).to(cuda)
    got = ops.time_conv_out(x, w, b, NB, Fr, H, W)
    x5 = x[:, :Cn].reshape(NB, Fr, H, W, Cn).permute(0, 4, 1, 2, 3).double().cpu()       # fp64 on the CPU: no TF32 convs
    ref = F.conv3d(x5, w.double().cpu()[..., None, None], b.double().cpu(), padding=(1, 0, 0))
    ref = ref.permute(0, 2, 1, 3, 4).reshape(NB * Fr, Cn, H, W)
    assert torch.allclose(got.double().cpu(), ref, atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("N,H,W,Ci,Co", [(2, 16, 24, 64, 64), (1, 64, 96, 128, 128), (3, 9, 11, 64, 32), (1, 2, 2, 64, 64)])
def test_gemm_conv_stride2_bottom_right_padding(cuda, N, H, W, Ci, Co):
    """diffusers Downsample2D(padding=0): F.pad(x, (0, 1, 0, 1)) + Conv2d(stride=2, padding=0)."""
    from lkgd_b200 import ops
    x = seeded_tensor("ds/x", (N, Ci, H, W)).to(bf16)
    w = (seeded_tensor("ds/w", (Co, Ci, 3, 3)) * (9 * Ci) ** -0.5).to(bf16)
    b = seeded_tensor("ds/b", (Co,))
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), b, stride=2)
    Ho, Wo = ref.shape[2:]
    rows = x.permute(0, 2, 3, 1).reshape(N * H * W, Ci).contiguous().to(cuda)
    wk = w.permute(0, 2, 3, 1).reshape(Co, 9 * Ci).contiguous().to(cuda)
    got = ops.gemm(rows, wk, mode=ops.A_CONV3X3, conv=(N, H, W, 2), pad_br=True, bias=b.to(cuda), out_f32=True)
    assert tuple(got.shape) == (N * Ho * Wo, Co)
    ref_rows = ref.permute(0, 2, 3, 1).reshape(N * Ho * Wo, Co)
    assert rel_l2(got, ref_rows) < 2e-3
    chk = ops.gemm(rows, wk, mode=ops.A_CONV3X3, conv=(N, H, W, 2), pad_br=True, bias=b.to(cuda), out_f32=True, checker=True)
    assert rel_l2(chk, ref_rows) < 2e-3


# ------------------------------------------------------------------------------------------------- the module
def _pair(cuda, **cfg):
    import oracle as O
    from lkgd_b200.vae import AutoencoderKLTemporalDecoder
    o = VG.build(O.AutoencoderKLTemporalDecoder, cfg).eval()
    p = AutoencoderKLTemporalDecoder(**cfg)
    p.load_state_dict(o.state_dict(), strict=True)
    return o, p.to(cuda)


@pytest.mark.parametrize("name,boc", [("flash_d64", (32, 64, 64, 64)), ("gemm_d256", (32, 64, 128, 256))])
def test_vae_small_configs_vs_oracle(cuda, name, boc):
    """Encoder + temporal decoder vs the oracle: 64-wide head (flash kernel) and one 256-wide head (GEMM - softmax - GEMM, the
    path the real VAE's 512-wide head takes); two clips of three frames, so temporal layers see B = 2."""
    o, p = _pair(cuda, block_out_channels=boc)
    img = torch.tanh(seeded_tensor("vae/img", (2, 3, 64, 96)))
    z = seeded_tensor("vae/z", (6, 4, 8, 12))
    with torch.no_grad():
        ref_m = o.quant_conv(o.encoder(img))
        ref_d = o.decode(z, num_frames=3).sample
    dist = p.encode(img.to(cuda)).latent_dist
    e_m = rel_l2(dist.parameters, ref_m)
    got = p.decode(z.to(cuda), num_frames=3).sample
    e_d = rel_l2(got, ref_d)
    print(name, "moments rel-L2", e_m, "decode rel-L2", e_d)
    assert tuple(dist.mode().shape) == (2, 4, 8, 12) and tuple(got.shape) == (6, 3, 64, 96)
    assert e_m < TOL_SMALL and e_d < TOL_SMALL
    # a different chunking changes the temporal mixing exactly as in the oracle
    with torch.no_grad():
        ref_1 = o.decode(z[:2], num_frames=2).sample
    assert rel_l2(p.decode(z[:2].to(cuda), num_frames=2).sample, ref_1) < TOL_SMALL
    g = torch.Generator(device="cpu").manual_seed(3)
    s = dist.sample(g)
    assert tuple(s.shape) == (2, 4, 8, 12) and float((s - dist.mode()).abs().max()) > 0


def test_vae_pipeline_helpers(cuda):
    """`decode_latents` (chunks of decode_chunk_size frames, [B, 3, F, H, W] fp32) and `_encode_vae_image` (zero
    unconditional half, unscaled posterior mode) as the reference pipeline calls them (pipeline...controlnet.py:216-237,268-295)."""
    import oracle as O
    from lkgd_b200.vae import decode_latents, encode_vae_image
    o, p = _pair(cuda, block_out_channels=(32, 32, 64, 64))
    lat = seeded_tensor("vae/lat", (1, 5, 4, 8, 8))
    with torch.no_grad():
        ref = O.decode_latents(o, lat, num_frames=5, decode_chunk_size=2)
    got = decode_latents(p, lat.to(cuda), num_frames=5, decode_chunk_size=2)
    assert tuple(got.shape) == (1, 3, 5, 64, 64) and got.dtype == torch.float32
    assert rel_l2(got, ref) < TOL_SMALL
    img = torch.tanh(seeded_tensor("vae/img2", (1, 3, 64, 64)))
    il = encode_vae_image(p, img.to(cuda), num_videos_per_prompt=2, do_classifier_free_guidance=True)
    with torch.no_grad():
        mode = o.encode(img).latent_dist.mode()
    assert tuple(il.shape) == (4, 4, 8, 8) and float(il[0].abs().max()) == 0.0 and float(il[2].abs().max()) == 0.0
    assert rel_l2(il[1], mode[0]) < TOL_SMALL and torch.equal(il[1], il[3])
    with pytest.raises(ValueError):
        p.decode(lat[0].to(cuda), num_frames=2)          # 5 frames are not a multiple of 2


def test_vae_full_size_against_fixture(cuda):
    """576x1024 frames at the SVD VAE's widths (97.7 M parameters, one 512-wide attention head over 9216 tokens): decode of two
    frames and encode of one image against tests/golden/vae_full_size.npz (fp32 oracle, ~1 min of CPU, generated here)."""
    from lkgd_b200.vae import SVD_VAE_CONFIG, AutoencoderKLTemporalDecoder
    path = os.path.join(HERE, "golden", "vae_full_size.npz")
    if not os.path.exists(path):
        pytest.fail(f"{path} is missing: run tests/golden/make_vae_golden.py")
    G = np.load(path)
    p = VG.build(AutoencoderKLTemporalDecoder, SVD_VAE_CONFIG).to(cuda)
    z, img = VG.inputs()
    mom = p.encode(img.to(cuda)).latent_dist.parameters
    e_m = rel_l2(mom, torch.from_numpy(G["encode/moments"]).float())
    y = p.decode(z.to(cuda), num_frames=2).sample
    e_d = rel_l2(y[:, :, ::VG.STEP, ::VG.STEP], torch.from_numpy(G["decode/sub"]).float())
    n = float(torch.linalg.norm(y.double()))
    print("full-size VAE: moments rel-L2", e_m, "decode rel-L2 (every 4th pixel)", e_d, "norm", n, float(G["decode/norm"][0]))
    assert tuple(y.shape) == (2, 3, 576, 1024)
    assert e_m < TOL_FULL and e_d < TOL_FULL and abs(n / float(G["decode/norm"][0]) - 1) < 5e-3


def test_pipeline_image_to_frames(cuda):
    """The whole reference `__call__` (pipeline...controlnet.py:470-646) with the three lkgd_b200 models registered: image ->
    CLIP embedding + noise-augmented VAE latents -> CFG Euler-Karras loop -> chunked temporal VAE decode -> frames in [0, 1],
    against the same chain of oracle modules (reduced UNet, small VAE / CLIP, 4 frames, 3 steps)."""
    import oracle as O
    from oracle.scheduler import SVD_SCHEDULER_CONFIG
    from lkgd_b200.clip import CLIPVisionModelWithProjection
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.preprocess import clip_pixel_values
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionControlNetModel
    from lkgd_b200.vae import AutoencoderKLTemporalDecoder
    from test_unet_gpu import _pair as unet_pair
    ucfg = dict(REDUCED_CONFIG)
    xdim = ucfg["cross_attention_dim"]
    ou, pu = unet_pair(O.UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionControlNetModel, ucfg, cuda)
    ov, pv = _pair(cuda, block_out_channels=(32, 32, 64, 64))
    ccfg = dict(hidden_size=64, intermediate_size=128, num_hidden_layers=2, num_attention_heads=4, image_size=224,
                patch_size=14, projection_dim=xdim, hidden_act="gelu")
    torch.manual_seed(0)
    oc = O.CLIPVisionModelWithProjection(**ccfg).eval()
    pc = CLIPVisionModelWithProjection(**ccfg)
    pc.load_state_dict(oc.state_dict(), strict=True)
    pc = pc.to(cuda)
    Fr, Hh, Ww, n, nas = 4, 128, 192, 3, 0.02
    image = torch.rand(1, 3, Hh, Ww, generator=torch.Generator().manual_seed(5))
    noise = torch.randn(1, Fr, 4, Hh // 8, Ww // 8, generator=torch.Generator().manual_seed(6))
    # ---- the oracle chain, written out as the reference pipeline does it
    with torch.no_grad():
        emb = oc(clip_pixel_values(image)).image_embeds.unsqueeze(1)
        emb = torch.cat([torch.zeros_like(emb), emb])
        aug = torch.randn(image.shape, generator=torch.Generator().manual_seed(7))
        lat = ov.encode(2.0 * image - 1.0 + nas * aug).latent_dist.mode()
        img_lat = torch.cat([torch.zeros_like(lat), lat]).unsqueeze(1).repeat(1, Fr, 1, 1, 1)
        osched = O.EulerDiscreteScheduler(**SVD_SCHEDULER_CONFIG)
        osched.set_timesteps(n)
        ids = O.add_time_ids_inference(6, 127, nas, 1)
        final = O.denoise_loop(ou, osched, noise * osched.init_noise_sigma, img_lat, emb, ids, n, 1.0, 3.0)
        ref = O.decode_latents(ov, final, Fr, decode_chunk_size=3)
        ref = (ref.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
    pipe = StableVideoDiffusionPipeline(pu, EulerDiscreteScheduler(**SVD_SCHEDULER_CONFIG), vae=pv, image_encoder=pc)
    out = pipe(image=image, height=Hh, width=Ww, num_frames=Fr, num_inference_steps=n, fps=7, noise_aug_strength=nas,
               latents=noise, generator=torch.Generator().manual_seed(7), decode_chunk_size=3, output_type="pt")
    e = rel_l2(out.frames, ref)
    print("image -> frames rel-L2", e)
    assert tuple(out.frames.shape) == (1, Fr, 3, Hh, Ww) and float(out.frames.min()) >= 0 and float(out.frames.max()) <= 1
    assert e < 2e-2
    lat_only = pipe(image=image, height=Hh, width=Ww, num_frames=Fr, num_inference_steps=n, fps=7, noise_aug_strength=nas,
                    latents=noise, generator=torch.Generator().manual_seed(7)).frames
    assert rel_l2(lat_only, final) < 2e-2
    npv = pipe(image=image, height=Hh, width=Ww, num_frames=Fr, num_inference_steps=1, latents=noise, output_type="np").frames
    assert npv.shape == (1, Fr, Hh, Ww, 3)
    with pytest.raises(ValueError):
        StableVideoDiffusionPipeline(pu, EulerDiscreteScheduler(**SVD_SCHEDULER_CONFIG))(image=image, num_frames=Fr)


def test_vae_full_size_chunk_invariance(cuda):
    """Size-independent property at the headline size (25 frames of 576x1024, SVD widths, 14.7 M output rows in one launch, ~54 GB):
    with the temporal branch switched off (mix_factor -> -inf: the switched blender's alpha is exactly 1; identity time_conv_out)
    every frame is decoded independently, so ONE 25-frame chunk must equal the reference scripts' chunks of 8 (8, 8, 8, 1) bit for
    bit - row indexing, tile order (pixel-tile-major temporal convs above 16 MB per frame), fused GroupNorm statistics and the
    NCHW unpack at full size."""
    from lkgd_b200.vae import SVD_VAE_CONFIG, AutoencoderKLTemporalDecoder
    torch.manual_seed(0)
    vae = AutoencoderKLTemporalDecoder(**SVD_VAE_CONFIG)
    with torch.no_grad():
        for n, prm in vae.named_parameters():
            if n.endswith("mix_factor"):
                prm.fill_(-40.0)
        w = vae.decoder.time_conv_out.weight
        w.zero_()
        w[:, :, 1, 0, 0] = torch.eye(3)
        vae.decoder.time_conv_out.bias.zero_()
    vae = vae.to(cuda)
    z = torch.randn(25, 4, 72, 128, device=cuda)
    full = vae.decode(z, num_frames=25).sample
    assert tuple(full.shape) == (25, 3, 576, 1024) and bool(torch.isfinite(full).all())
    parts = torch.cat([vae.decode(z[i:i + 8], num_frames=z[i:i + 8].shape[0]).sample for i in range(0, 25, 8)])
    assert torch.equal(full, parts) and float(parts.abs().mean()) > 1e-2
    del full, parts
    torch.cuda.empty_cache()
