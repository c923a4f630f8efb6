"""CPU oracle for the LKGD / Stable-Video-Diffusion denoise hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``lkgd_b200/`` may import this package: only
``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / reported CPU baseline - never as the product path.

PARITY UNPINNED: the reference (/root/reference) ships no tests, golden vectors or
known-answer files for this path, and its block arithmetic lives in un-vendored,
un-installable dependencies (``diffusers==0.27.2``, ``peft==0.10.0``, un-pinned ``core_qnn``;
reference ``requirements.txt:15,44``).  The restatement here follows

* the reference's own wiring / forward code where it exists in-tree (cited per function), and
* the published algorithm of the pinned dependency otherwise (every
  recollection that cannot be checked in-tree is a constructor switch, U1-U6),

and is pinned by (i) the reference's two parameter-name dumps ``train_svd_lora.txt`` /
``train_svd_lora_train.txt`` (1553 names; ``tests/golden/param_names.json``), and (ii) the
known-answer values derivable from in-tree formulas (``tests/test_oracle_kat.py``).

Everything is plain PyTorch on CPU in fp32 (or fp64 for self-consistency checks).
"""

from .blocks import *  # noqa: F401,F403
from .unet import (  # noqa: F401
    UNetSpatioTemporalConditionControlNetModel,
    UNetSpatioTemporalConditionModelFlow,
    UNetSpatioTemporalConditionJointModel,
    UNetSpatioTemporalConditionModel,
    SVD_XT_CONFIG,
    REDUCED_CONFIG,
)
from .controlnet import ControlNetSDVModel  # noqa: F401
from .lora import LoraLinear, add_lora, merge_lora  # noqa: F401
from .scheduler import EulerDiscreteScheduler  # noqa: F401
from .pipeline import (guidance_ramp, cfg_combine, denoise_loop, add_time_ids_inference, add_time_ids_training,  # noqa: F401
                       smooth_chunks, smooth_loop)
from .clip import CLIPVisionModelWithProjection, CLIP_VIT_H_14  # noqa: F401,E402
from .vae import AutoencoderKLTemporalDecoder, SVD_VAE_CONFIG, decode_latents  # noqa: F401,E402
