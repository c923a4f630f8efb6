"""Generates tests/golden/preprocess_golden.npz by RUNNING the reference's own `_resize_with_antialiasing`
(pipeline/pipeline_stable_video_diffusion_controlnet.py:672-784) on seeded images - pins lkgd_b200/preprocess.py.
Dev container only.        python tests/golden/make_preprocess_golden.py"""
import pathlib
import sys

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]
from weights import seeded_tensor  # noqa: E402
from pipeline.pipeline_stable_video_diffusion_controlnet import _resize_with_antialiasing  # noqa: E402

out = {}
for tag, shape in {"576x1024": (1, 3, 576, 1024), "224x300": (1, 3, 224, 300)}.items():
    img = seeded_tensor(f"pre/{tag}", shape).sigmoid()          # values in (0, 1)
    out[f"pre/{tag}"] = _resize_with_antialiasing(img * 2.0 - 1.0, (224, 224)).half().numpy()    # fp16: keeps the file small
np.savez_compressed(HERE / "preprocess_golden.npz", **out)
print({k: v.shape for k, v in out.items()})
