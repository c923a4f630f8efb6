"""Generates tests/golden/fusion_golden.npz by EXECUTING THE REFERENCE'S OWN LINES of the ``direct_fusion`` branch
(/root/reference/pipeline/pipeline_stable_video_diffusion_trans_controlnet.py, the ``else:`` body at :640-667 - the code
is inline in ``__call__``, so the lines are lifted textually, dedented and run against the reference's unmodified
scheduler).  Dev container only; the .npz is committed.

    python tests/golden/make_fusion_golden.py
"""
import pathlib
import sys
import textwrap
from types import SimpleNamespace

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import seeded_tensor  # noqa: E402
from utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler  # noqa: E402

SRC = pathlib.Path("/root/reference/pipeline/pipeline_stable_video_diffusion_trans_controlnet.py").read_text().splitlines()
start = next(i for i, l in enumerate(SRC) if l.strip() == "model_output = noise_pred" and SRC[i - 1].strip() == "else:")
end = next(i for i in range(start, len(SRC)) if SRC[i].strip() == "self.scheduler._step_index += 1")
body = textwrap.dedent("\n".join(SRC[start:end + 1]))
print(f"executing reference lines {start + 1}-{end + 1}")


def _append_dims(x, target_dims):          # the reference's helper (:70-75), same file
    return x[(...,) + (None,) * (target_dims - x.ndim)]


def randn_tensor(shape, dtype=None, device=None, generator=None):
    return torch.randn(shape, dtype=dtype, generator=generator)


SCHED = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
             prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
             sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1)
sched = EulerDiscreteScheduler(**SCHED)
sched.set_timesteps(10)
self = SimpleNamespace(scheduler=sched)
latents = seeded_tensor("fusion/x0", (2, 5, 4, 8, 8)) * sched.init_noise_sigma
traj = []
for i, t in enumerate(sched.timesteps):
    noise_pred = seeded_tensor(f"fusion/v{i}", latents.shape)
    ns = dict(self=self, noise_pred=noise_pred, latents=latents, t=t, generator=None, torch=torch,
              randn_tensor=randn_tensor, _append_dims=_append_dims)
    exec(body, ns)
    latents = ns["latents"]
    traj.append(latents.numpy())
np.savez_compressed(HERE / "fusion_golden.npz", traj=np.stack(traj))
print(np.stack(traj).shape)
