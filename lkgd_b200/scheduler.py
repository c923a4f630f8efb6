"""Euler-discrete scheduler with Karras sigmas (the reference's "fix" fork), B200 edition.

Same surface and state machine as reference ``utils/scheduling_euler_discrete_karras_fix.py``
(``set_timesteps`` :290-350, ``scale_model_input`` :264-288, ``step`` :418-528, ``add_noise`` :530-553,
``init_noise_sigma`` :248-255, ``_step_index`` / ``_init_step_index`` :243-246,405-416); the schedule itself is
host-side numpy exactly like the reference, the per-step tensor math runs in the fused fp32 CUDA kernels
(``lkgd_cfg_euler_step`` / ``lkgd_scale_f32``).  A host copy of sigmas / timesteps is kept so that stepping
with python floats or the scheduler's own timestep tensors never synchronises with the device.

Deliberate difference: ``step`` does not draw the unused ``randn`` the reference draws when gamma == 0
(quirk D5) unless ``consume_rng=True`` - it only matters for RNG-stream parity with the global generator."""
from __future__ import annotations

from dataclasses import dataclass
from types import SimpleNamespace
from typing import Optional, Union

import numpy as np
import torch

from . import ops

SVD_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True,
    sigma_min=0.002, sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1,
)


@dataclass
class EulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.0001, beta_end: float = 0.02,
                 beta_schedule: str = "linear", trained_betas=None, prediction_type: str = "epsilon",
                 interpolation_type: str = "linear", use_karras_sigmas: Optional[bool] = False,
                 sigma_min: Optional[float] = None, sigma_max: Optional[float] = None,
                 timestep_spacing: str = "linspace", timestep_type: str = "discrete", steps_offset: int = 0,
                 rescale_betas_zero_snr: bool = False):
        if rescale_betas_zero_snr:
            raise NotImplementedError("rescale_betas_zero_snr is not used by any SVD configuration")
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, trained_betas=trained_betas, prediction_type=prediction_type,
            interpolation_type=interpolation_type, use_karras_sigmas=use_karras_sigmas, sigma_min=sigma_min,
            sigma_max=sigma_max, timestep_spacing=timestep_spacing, timestep_type=timestep_type,
            steps_offset=steps_offset, rescale_betas_zero_snr=rescale_betas_zero_snr)
        # host-side torch (CPU) for bit-equality with the reference's float32 schedule arithmetic (:203-221)
        if trained_betas is not None:
            betas = torch.tensor(trained_betas, dtype=torch.float32)
        elif beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self._alphas_cumprod = torch.cumprod(1.0 - betas, dim=0)
        self.use_karras_sigmas = use_karras_sigmas
        train_sigmas = self._train_sigmas()[::-1].copy()
        timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        if use_karras_sigmas:      # the "fix": Karras sigmas also at construction (reference :225-228)
            log_sigmas = np.log(train_sigmas)
            train_sigmas = self._convert_to_karras(train_sigmas, num_train_timesteps)
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in train_sigmas])
        self.num_inference_steps = None
        self._install(train_sigmas, timesteps, None)
        self.is_scale_input_called = False

    # ------------------------------------------------------------------ schedule (host)
    def _train_sigmas(self) -> np.ndarray:
        a = self._alphas_cumprod
        return (((1 - a) / a) ** 0.5).numpy()      # float32 tensor -> float32 ndarray (np.array(tensor) in the reference)

    def _install(self, sigmas: np.ndarray, timesteps: np.ndarray, device):
        sig32 = sigmas.astype(np.float32)
        if self.config.timestep_type == "continuous" and self.config.prediction_type == "v_prediction":
            # 0.25 * sigma.log() element by element on fp32 0-d tensors, as the reference does (:344-345)
            ts32 = torch.Tensor([0.25 * s.log() for s in torch.from_numpy(sig32)]).numpy()
        else:
            ts32 = timesteps.astype(np.float32)
        self._sigmas_host = np.concatenate([sig32, np.zeros(1, np.float32)])
        self._timesteps_host = ts32
        self.sigmas = torch.from_numpy(self._sigmas_host.copy()).to(device=device)
        self.timesteps = torch.from_numpy(ts32.copy()).to(device=device)
        self._step_index = None

    @property
    def init_noise_sigma(self):
        m = torch.tensor(float(self._sigmas_host.max()), dtype=torch.float32)
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return m
        return (m ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _convert_to_karras(self, in_sigmas, num_inference_steps) -> np.ndarray:
        smin = self.config.sigma_min if self.config.sigma_min is not None else float(in_sigmas[-1])
        smax = self.config.sigma_max if self.config.sigma_max is not None else float(in_sigmas[0])
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        return (smax ** (1 / rho) + ramp * (smin ** (1 / rho) - smax ** (1 / rho))) ** rho

    @staticmethod
    def _sigma_to_t(sigma, log_sigmas):
        log_sigma = np.log(np.maximum(sigma, 1e-10))
        dists = log_sigma - log_sigmas[:, np.newaxis]
        low = np.cumsum((dists >= 0), axis=0).argmax(axis=0).clip(max=log_sigmas.shape[0] - 2)
        high = low + 1
        w = np.clip((log_sigmas[low] - log_sigma) / (log_sigmas[low] - log_sigmas[high]), 0, 1)
        return ((1 - w) * low + w * high).reshape(np.shape(sigma))

    def set_timesteps(self, num_inference_steps: int, device: Union[str, torch.device] = None):
        c = self.config
        self.num_inference_steps = num_inference_steps
        if c.timestep_spacing == "linspace":
            timesteps = np.linspace(0, c.num_train_timesteps - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c.timestep_spacing == "leading":
            ratio = c.num_train_timesteps // num_inference_steps
            timesteps = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32)
            timesteps += c.steps_offset
        elif c.timestep_spacing == "trailing":
            ratio = c.num_train_timesteps / num_inference_steps
            timesteps = (np.arange(c.num_train_timesteps, 0, -ratio)).round().copy().astype(np.float32) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported. Please make sure to choose one of "
                             "'linspace', 'leading' or 'trailing'.")
        sigmas = self._train_sigmas()
        log_sigmas = np.log(sigmas)
        if c.interpolation_type == "linear":
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        elif c.interpolation_type == "log_linear":
            sigmas = torch.linspace(np.log(sigmas[-1]), np.log(sigmas[0]), num_inference_steps + 1).exp().numpy()
        else:
            raise ValueError(f"{c.interpolation_type} is not implemented. Please specify interpolation_type to "
                             "either 'linear' or 'log_linear'")
        if self.use_karras_sigmas:
            sigmas = self._convert_to_karras(sigmas, num_inference_steps)
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas])
        self._install(np.asarray(sigmas), np.asarray(timesteps), device)

    # ------------------------------------------------------------------ step index (host)
    def _init_step_index(self, timestep):
        t = float(timestep.item()) if isinstance(timestep, torch.Tensor) else float(timestep)
        cand = np.nonzero(self._timesteps_host == np.float32(t))[0]
        if len(cand) == 0:
            raise IndexError("timestep is not one of scheduler.timesteps")
        self._step_index = int(cand[1] if len(cand) > 1 else cand[0])

    def index_for(self, i: int):
        """Fast path for loops that already know the step number: no device read-back."""
        self._step_index = int(i)

    # ------------------------------------------------------------------ per-step tensor math (device)
    def scale_model_input(self, sample: torch.Tensor, timestep) -> torch.Tensor:
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = float(self._sigmas_host[self._step_index])
        self.is_scale_input_called = True
        out = ops.scale_f32(sample.to(torch.float32), float(1.0 / np.sqrt(np.float32(sigma) ** 2 + 1)))
        return out if sample.dtype == torch.float32 else out.to(sample.dtype)

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, s_churn: float = 0.0,
             s_tmin: float = 0.0, s_tmax: float = float("inf"), s_noise: float = 1.0,
             generator: Optional[torch.Generator] = None, return_dict: bool = True, consume_rng: bool = False):
        if isinstance(timestep, int) or isinstance(timestep, (torch.IntTensor, torch.LongTensor)) or (
                isinstance(timestep, torch.Tensor) and not timestep.is_floating_point()):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        if self.config.prediction_type != "v_prediction":
            raise NotImplementedError("the fused CUDA step implements v_prediction (the SVD configuration)")
        if s_churn != 0.0:
            raise NotImplementedError("s_churn > 0 (stochastic sampling) is never used by the reference pipelines")
        if self._step_index is None:
            self._init_step_index(timestep)
        if consume_rng:   # reference quirk D5 (:485-489)
            torch.randn(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        if not self.is_scale_input_called:      # reference :470-474 (logger.warning)
            import logging
            logging.getLogger(__name__).warning(
                "The `scale_model_input` function should be called before `step` to ensure correct denoising. "
                "See `StableDiffusionPipeline` for a usage example.")
        sigma = float(self._sigmas_host[self._step_index])
        sigma_next = float(self._sigmas_host[self._step_index + 1])
        if model_output.shape != sample.shape:
            raise ValueError("model_output and sample must have the same shape")
        # the update is elementwise: any shape the reference accepts (4-D image latents, 5-D video latents, ...) is one
        # flat fp32 vector for the fused kernel; upcast like the reference (:481), cast back (:520)
        shape = sample.shape
        pred = model_output.to(torch.float32).reshape(1, 1, 1, 1, -1)
        with torch.cuda.device(sample.device):
            prev, _, x0 = ops.cfg_euler_step(pred, None, sample.to(torch.float32).reshape(1, 1, 1, 1, -1), sigma,
                                             sigma_next, cfg=False, want_x0=True)
        prev = prev.reshape(shape).to(model_output.dtype)
        x0 = x0.reshape(shape).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return EulerDiscreteSchedulerOutput(prev_sample=prev, pred_original_sample=x0)

    def step_direct_fusion(self, model_output: torch.Tensor, timestep, sample: torch.Tensor,
                           generator: Optional[torch.Generator] = None, consume_rng: bool = False) -> torch.Tensor:
        """The ``direct_fusion`` branch of the reference's trans pipelines
        (pipeline/pipeline_stable_video_diffusion_trans_controlnet.py:639-667): ``model_output`` / ``sample`` hold the
        forward samples followed by their time-reversed partners; returns the next latents (one fused kernel).
        ``consume_rng``: draw the noise tensor the reference draws and never uses (:646-648), for RNG-stream parity."""
        if self.config.prediction_type != "v_prediction":
            raise NotImplementedError("the fused CUDA step implements v_prediction (the SVD configuration)")
        if self._step_index is None:
            self._init_step_index(timestep)
        if consume_rng:
            torch.randn(model_output.shape, dtype=model_output.dtype, device=model_output.device, generator=generator)
        sigma = float(self._sigmas_host[self._step_index])
        sigma_next = float(self._sigmas_host[self._step_index + 1])
        F = model_output.shape[1]
        w = torch.linspace(1, 0, F).to(model_output.device)
        out = ops.fusion_euler_step(model_output.to(torch.float32).contiguous(), sample.to(torch.float32).contiguous(),
                                    w, sigma, sigma_next)
        self._step_index += 1
        return out.to(model_output.dtype)

    def step_cfg_rows(self, pred_rows: torch.Tensor, guidance: Optional[torch.Tensor], sample: torch.Tensor,
                      cfg: bool, want_v: bool = False, sigmas_dev: Optional[torch.Tensor] = None,
                      in_place: bool = False, pred_cond: Optional[torch.Tensor] = None):
        """Fused CFG combine + Euler update straight from the UNet's channels-last fp32 prediction
        (pipeline/pipeline_stable_video_diffusion_controlnet.py:614-619 in one kernel).  ``sigmas_dev`` / ``in_place``:
        the CUDA-graph form (per-step sigmas read from device memory, latents updated in their static buffer)."""
        sigma = float(self._sigmas_host[self._step_index])
        sigma_next = float(self._sigmas_host[self._step_index + 1])
        out = ops.cfg_euler_step(pred_rows, guidance, sample, sigma, sigma_next, cfg=cfg, want_v=want_v,
                                 sigmas_dev=sigmas_dev, in_place=in_place, pred_cond=pred_cond)
        self._step_index += 1
        return out

    def add_noise(self, original_samples: torch.Tensor, noise: torch.Tensor, timesteps: torch.Tensor):
        ts = timesteps.detach().to("cpu", torch.float32).reshape(-1).numpy()
        idx = [int(np.nonzero(self._timesteps_host == t)[0].item()) for t in ts]
        sigma = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)[idx].flatten()
        while sigma.ndim < original_samples.ndim:
            sigma = sigma.unsqueeze(-1)
        return original_samples + noise * sigma      # training-side helper, off the sampling hot path

    def __len__(self):
        return self.config.num_train_timesteps
