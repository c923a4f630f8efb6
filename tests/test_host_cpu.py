"""CPU tests of the host side: the C-ABI library loads and exports every symbol the header declares (no compute
calls - there is no GPU here), the ctypes struct mirrors the C struct, the reference-facing modules keep the
reference's names / errors, the host-side schedule is bit-exact against the reference-generated golden, and the
product never imports the oracle."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from golden_util import REDUCED4, SCHED, golden

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G = golden()


@pytest.fixture(scope="module")
def lib():
    from lkgd_b200 import _lib, build
    build.build()                      # no-op when lkgd_b200/lib/liblkgd_b200.so is newer than its sources
    return _lib.load()


def _header_symbols():
    src = open(os.path.join(ROOT, "include", "lkgd_b200.h")).read()
    return re.findall(r"LKGD_API\s+[\w\s\*]+?\b(lkgd_\w+)\s*\(", src)


def test_library_exports_every_declared_symbol(lib):
    from lkgd_b200 import _lib
    syms = _header_symbols()
    assert len(syms) >= 25 and len(set(syms)) == len(syms)
    assert set(syms) == set(_lib.SIGNATURES), set(syms) ^ set(_lib.SIGNATURES)
    exported = subprocess.run(["nm", "-D", "--defined-only", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in exported.splitlines() if " T " in ln}
    assert set(syms) <= exported
    assert {s for s in exported if not s.startswith("lkgd_")} <= {"_init", "_fini"}     # nothing else leaks
    assert lib.lkgd_abi_version() == _lib.ABI_VERSION
    assert lib.lkgd_strerror(0) is not None and b"shape" in lib.lkgd_strerror(-1).lower()


@pytest.mark.parametrize("cname,pyname", [("lkgd_gemm_args", "GemmArgs"), ("lkgd_cast2d_job", "Cast2dJob")])
def test_gemm_args_struct_layout_matches_c(lib, tmp_path, cname, pyname):
    """sizeof / offsetof of the ABI's descriptor structs as gcc sees the header == the ctypes mirrors."""
    from lkgd_b200 import _lib
    import ctypes as C
    S = getattr(_lib, pyname)
    fields = [f[0] for f in S._fields_]
    prog = '#include <stdio.h>\n#include <stddef.h>\n#include "lkgd_b200.h"\nint main(){printf("%zu", sizeof(' + cname + '));' \
        + "".join(f'printf(" %zu", offsetof({cname}, {f}));' for f in fields) + "return 0;}\n"
    src = tmp_path / "layout.c"
    src.write_text(prog)
    exe = tmp_path / "layout"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    vals = [int(v) for v in subprocess.check_output([str(exe)], text=True).split()]
    assert vals[0] == C.sizeof(S)
    assert vals[1:] == [getattr(S, f).offset for f in fields]


def test_error_codes_without_a_gpu(lib):
    """Argument validation happens before any CUDA call, so it can be exercised here."""
    from lkgd_b200._lib import GemmArgs
    import ctypes as C
    a = GemmArgs()
    assert lib.lkgd_gemm(C.byref(a), None) == -1                     # LKGD_ESHAPE: null operands
    assert lib.lkgd_attention(None, 8, None, 8, None, 8, None, 8, 0, 1, 64, 1, 1, 1.0, None) == -1
    assert lib.lkgd_attention_temporal(None, None, 1, 33, 1, 1, 64, 1.0, None) == -1   # F > 32


def test_product_does_not_import_the_oracle():
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lkgd_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                s = open(os.path.join(dirpath, f)).read()
                if re.search(r"^\s*(from|import)\s+oracle\b", s, re.M) or "golden" in s:
                    bad.append(f)
    assert not bad, bad


def test_no_cpu_fallback():
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    m = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    x = torch.zeros(1, 4, 8, 16, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(x, 0.5, torch.zeros(1, 1, 32), added_time_ids=torch.zeros(1, 3))
    from lkgd_b200 import ops
    with pytest.raises(ValueError, match="CUDA"):
        ops.layernorm(torch.zeros(4, 8), torch.ones(8), torch.zeros(8))
    # the entry points added with ABI v7 refuse host tensors the same way
    with pytest.raises(ValueError, match="CUDA"):
        ops.Cast2dBatch([(torch.zeros(4, 8), torch.zeros(4, 8, dtype=torch.bfloat16), 1.0)])
    with pytest.raises(ValueError, match="no jobs"):
        ops.Cast2dBatch([])
    with pytest.raises(ValueError, match="CUDA"):
        ops.cfg_euler_step(torch.zeros(16, 4), torch.ones(1), torch.zeros(1, 1, 4, 4, 4), 2.0, 1.0, cfg=True,
                           pred_cond=torch.zeros(16, 4))


def test_module_parameter_names_match_reference_dumps():
    from lkgd_b200.unet import SVD_XT_CONFIG, UNetSpatioTemporalConditionModel
    d = json.load(open(os.path.join(ROOT, "tests", "golden", "param_names.json")))
    with torch.device("meta"):
        m = UNetSpatioTemporalConditionModel(**SVD_XT_CONFIG)
        hits = m.add_lora(4)
    assert len(hits) == 48                                            # 16 temporal blocks x q,k,v (F9)
    names = {n for n, _ in m.named_parameters()}
    assert names == set(d["frozen"]) | set(d["trainable"])
    # the pipeline reads these (reference pipeline ...controlnet.py:252-253,465-468,534)
    assert m.config.in_channels == 8 and m.config.num_frames == 25 and m.config.addition_time_embed_dim == 256
    assert m.add_embedding.linear_1.in_features == 768


def test_constructor_errors_mirror_the_reference():
    from lkgd_b200.unet import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel as U
    with torch.device("meta"):
        with pytest.raises(ValueError, match="same number of `down_block_types` as `up_block_types`"):
            U(**dict(REDUCED4, up_block_types=("UpBlockSpatioTemporal",)))
        with pytest.raises(ValueError, match="`block_out_channels` as `down_block_types`"):
            U(**dict(REDUCED4, block_out_channels=(32,)))
        with pytest.raises(ValueError, match="`num_attention_heads` as `down_block_types`"):
            U(**dict(REDUCED4, num_attention_heads=(2,)))
        with pytest.raises(ValueError, match="`layers_per_block` as `down_block_types`"):
            U(**dict(REDUCED4, layers_per_block=(2,)))
        with pytest.raises(ValueError, match="does not exist"):
            U(**dict(REDUCED4, down_block_types=("CrossAttnDownBlockSpatioTemporal", "Nope")))
        with pytest.raises(ValueError, match="time_context_order"):
            U(**dict(REDUCED4, time_context_order="x"))
        u = U(**REDUCED4)
        with pytest.raises(ValueError, match="either 0 or 1"):
            u.enable_forward_chunking(dim=2)
        cfg = {k: v for k, v in REDUCED4.items() if k != "up_block_types"}
        cn = ControlNetSDVModel(**cfg, conditioning_channels=2)
        assert len(cn.controlnet_down_blocks) == 6


@pytest.mark.parametrize("n", [25, 10])
def test_host_schedule_bit_exact_vs_reference(n):
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    s = EulerDiscreteScheduler(**SCHED)
    assert np.array_equal(s.sigmas[:8].numpy(), G["sched/init_sigmas_head"])
    assert np.array_equal(s.timesteps[:8].numpy(), G["sched/init_timesteps_head"])
    s.set_timesteps(n)
    assert s.order == 1 and s.num_inference_steps == n
    assert np.array_equal(s.sigmas.numpy(), G[f"sched{n}/sigmas"])
    assert np.array_equal(s.timesteps.numpy(), G[f"sched{n}/timesteps"])
    assert float(s.init_noise_sigma) == float(G[f"sched{n}/init_noise_sigma"])
    s._init_step_index(s.timesteps[3])
    assert s.step_index == 3
    with pytest.raises(ValueError, match="integer indices"):
        s.step(torch.zeros(1), 3, torch.zeros(1))
    assert np.array_equal(s.add_noise(torch.zeros(2, 1), torch.ones(2, 1), s.timesteps[[0, 5]]).flatten().numpy(),
                          G[f"sched{n}/sigmas"][[0, 5]])


def test_pipeline_argument_checks():
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionModel
    u = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    p = StableVideoDiffusionPipeline(u, EulerDiscreteScheduler(**SCHED))
    ids = p._get_add_time_ids(6, 127, 0.02, torch.float32, 1, 1, True)
    assert ids.tolist() == [[6.0, 127.0, pytest.approx(0.02)]] * 2      # inference order (F12)
    emb, lat = torch.zeros(2, 1, 32), torch.zeros(2, 4, 4, 16, 16)
    with pytest.raises(ValueError, match="output_type"):
        p(emb, lat, output_type="pil")
    with pytest.raises(ValueError, match="num_frames"):
        p.prepare(emb, lat, num_frames=5)
    with pytest.raises(ValueError, match="controlnet"):
        p.prepare(emb, lat, controlnet_condition=torch.zeros(4, 2, 128, 128))
    with pytest.raises(ValueError, match="generators"):
        p.prepare_latents(2, 4, 8, 16, 16, torch.float32, "cpu", [torch.Generator()])
    # num_videos_per_prompt: the conditioning arrives already repeated (reference `_encode_image` :204 /
    # `_encode_vae_image` :234), so 2 prompts x 2 videos = 4 samples, 8 CFG rows - counted once, not twice
    emb4, lat4 = torch.zeros(8, 1, 32), torch.zeros(8, 4, 4, 16, 16)
    st = p.prepare(emb4, lat4, num_videos_per_prompt=2)
    assert st["S"] == 4 and st["n_batch"] == 8 and tuple(st["added_time_ids"].shape) == (8, 3)
    assert tuple(p.guidance_scale.shape) == (4, 4, 1, 1, 1)
    assert tuple(p.prepare_latents(st["S"], 4, 8, 16, 16, torch.float32, "cpu", None).shape) == (4, 4, 4, 16, 16)
    with pytest.raises(ValueError, match="num_videos_per_prompt"):
        p.prepare(emb, lat, num_videos_per_prompt=2)
    with torch.device("meta"):
        lk = UNetSpatioTemporalConditionModel(**dict(REDUCED4, cross_attention_dim=1024))
    with pytest.raises(ValueError, match="domain_features"):
        StableVideoDiffusionPipeline(lk, EulerDiscreteScheduler(**SCHED)).prepare(emb, lat)


def test_flop_model_matches_baseline_md():
    from lkgd_b200.flops import unet_flops
    from lkgd_b200.unet import SVD_XT_CONFIG
    c2 = unet_flops(dict(SVD_XT_CONFIG, num_frames=14), 2, 14, 72, 128)["total"] / 1e12
    c3 = unet_flops(SVD_XT_CONFIG, 2, 25, 72, 128, lora_rank=64)["total"] / 1e12
    assert abs(c2 - 89.6) < 0.6 and abs(c3 - 160.9) < 0.6


def test_vae_and_clip_argument_checks():
    """SURVEY 8f N1 modules: constructor / argument errors are raised before anything touches the device, and a CPU call
    fails loudly (no fallback)."""
    from lkgd_b200.clip import CLIPVisionModelWithProjection
    from lkgd_b200.flops import vae_flops
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    from lkgd_b200.vae import SVD_VAE_CONFIG, AutoencoderKLTemporalDecoder
    with pytest.raises(ValueError, match="multiples of 32"):
        AutoencoderKLTemporalDecoder(block_out_channels=(24, 32, 64, 64))
    with pytest.raises(ValueError, match="channel counts"):
        AutoencoderKLTemporalDecoder(block_out_channels=(32, 32), out_channels=2)
    v = AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 32, 32))
    with pytest.raises(ValueError, match="encode expects"):
        v.encode(torch.zeros(1, 4, 64, 64))
    with pytest.raises(ValueError, match="multiples of 8"):
        v.encode(torch.zeros(1, 3, 60, 64))
    with pytest.raises(ValueError, match="decode expects"):
        v.decode(torch.zeros(5, 4, 8, 8), num_frames=2)
    with pytest.raises(RuntimeError, match="CUDA"):
        v.decode(torch.zeros(4, 4, 8, 8), num_frames=2)
    with pytest.raises(ValueError, match="hidden_act"):
        CLIPVisionModelWithProjection(hidden_size=64, num_attention_heads=4, hidden_act="relu")
    with pytest.raises(ValueError, match="head width"):
        CLIPVisionModelWithProjection(hidden_size=60, num_attention_heads=4)
    f = vae_flops(SVD_VAE_CONFIG, 25, 72, 128)
    assert 170e12 < f["decode"] < 177e12 and 2.5e12 < f["encode"] < 2.7e12
    with torch.device("meta"):
        u = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
    pipe = StableVideoDiffusionPipeline(u, EulerDiscreteScheduler(**SCHED))
    with pytest.raises(ValueError, match="vae and image_encoder"):
        pipe.encode_inputs(torch.zeros(1, 3, 64, 64), 64, 64, 4)
    with pytest.raises(ValueError, match="needs the pipeline's vae"):
        pipe.decode_latents(torch.zeros(1, 4, 4, 8, 8), 4)
    with pytest.raises(ValueError, match="pass image="):
        pipe(num_frames=4)
    with pytest.raises(ValueError, match="decoded frames need"):
        pipe(torch.zeros(2, 1, 32), torch.zeros(2, 4, 4, 16, 16), output_type="pt")
    vid = pipe.tensor2vid(torch.full((1, 3, 2, 4, 4), 3.0), "np")
    assert vid.shape == (1, 2, 4, 4, 3) and float(vid.max()) == 1.0
