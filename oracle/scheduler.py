"""Oracle Euler-discrete / Karras scheduler (TEST INFRA ONLY).

Restates reference ``utils/scheduling_euler_discrete_karras_fix.py``: ``__init__`` :179-246 (incl. the
"fix": Karras sigmas also at construction, :225-228), ``init_noise_sigma`` :248-255,
``scale_model_input`` :264-288, ``set_timesteps`` :290-350, ``_sigma_to_t`` :352-373,
``_convert_to_karras`` :376-399, ``_init_step_index`` :405-416, ``step`` :418-528, ``add_noise`` :530-553.
No diffusers mixins; config values are read from ``self.config`` (quirk D4)."""
from __future__ import annotations

from types import SimpleNamespace

import numpy as np
import torch

# SVD scheduler_config.json as recalled (U-sched); corroborated in-tree by
# train_models/train_svd_lora.py:309-310 (0.002/700), :1527-1528 (t = 0.25 ln sigma), :1653-1654 (v-pred).
SVD_SCHEDULER_CONFIG = dict(
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True,
    sigma_min=0.002, sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1,
)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 prediction_type="epsilon", interpolation_type="linear", use_karras_sigmas=False,
                 sigma_min=None, sigma_max=None, timestep_spacing="linspace", timestep_type="discrete",
                 steps_offset=0):
        self.config = SimpleNamespace(
            num_train_timesteps=num_train_timesteps, beta_start=beta_start, beta_end=beta_end,
            beta_schedule=beta_schedule, prediction_type=prediction_type, interpolation_type=interpolation_type,
            use_karras_sigmas=use_karras_sigmas, sigma_min=sigma_min, sigma_max=sigma_max,
            timestep_spacing=timestep_spacing, timestep_type=timestep_type, steps_offset=steps_offset)
        if beta_schedule == "linear":
            self.betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            self.betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps,
                                        dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(f"{beta_schedule} does is not implemented for {self.__class__}")
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()[::-1].copy()
        timesteps = np.linspace(0, num_train_timesteps - 1, num_train_timesteps, dtype=float)[::-1].copy()
        if use_karras_sigmas:
            log_sigmas = np.log(sigmas)
            sigmas = self._convert_to_karras(sigmas, num_train_timesteps)
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas])
        sigmas = torch.from_numpy(sigmas).to(torch.float32)
        self.num_inference_steps = None
        if timestep_type == "continuous" and prediction_type == "v_prediction":
            self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas])
        else:
            self.timesteps = torch.from_numpy(timesteps.astype(np.float32))
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.is_scale_input_called = False
        self._step_index = None

    @property
    def init_noise_sigma(self):
        m = self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return m
        return (m ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _convert_to_karras(self, in_sigmas, num_inference_steps):
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

    def set_timesteps(self, num_inference_steps: int, device=None):
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
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        log_sigmas = np.log(sigmas)
        if c.interpolation_type == "linear":
            sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        elif c.interpolation_type == "log_linear":
            sigmas = torch.linspace(np.log(sigmas[-1]), np.log(sigmas[0]), num_inference_steps + 1).exp().numpy()
        else:
            raise ValueError(f"{c.interpolation_type} is not implemented.")
        if c.use_karras_sigmas:
            sigmas = self._convert_to_karras(sigmas, num_inference_steps)
            timesteps = np.array([self._sigma_to_t(s, log_sigmas) for s in sigmas])
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        if c.timestep_type == "continuous" and c.prediction_type == "v_prediction":
            self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas]).to(device=device)
        else:
            self.timesteps = torch.from_numpy(timesteps.astype(np.float32)).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def _init_step_index(self, timestep):
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        cand = (self.timesteps == timestep).nonzero()
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        self.is_scale_input_called = True
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample, s_churn=0.0, s_tmin=0.0, s_tmax=float("inf"), s_noise=1.0,
             generator=None, return_dict=True):
        if isinstance(timestep, int) or isinstance(timestep, (torch.IntTensor, torch.LongTensor)):
            raise ValueError("Passing integer indices (e.g. from `enumerate(timesteps)`) as timesteps to"
                             " `EulerDiscreteScheduler.step()` is not supported. Make sure to pass"
                             " one of the `scheduler.timesteps` as a timestep.")
        if self._step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index]
        gamma = min(s_churn / (len(self.sigmas) - 1), 2 ** 0.5 - 1) if s_tmin <= sigma <= s_tmax else 0.0
        # quirk D5: the reference always draws this noise (consumes the RNG) even when gamma == 0
        noise = torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator).to(model_output.device)
        sigma_hat = sigma * (gamma + 1)
        if gamma > 0:
            sample = sample + noise * s_noise * (sigma_hat ** 2 - sigma ** 2) ** 0.5
        pt = self.config.prediction_type
        if pt in ("original_sample", "sample"):
            x0 = model_output
        elif pt == "epsilon":
            x0 = sample - sigma_hat * model_output
        elif pt == "v_prediction":
            x0 = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        else:
            raise ValueError(f"prediction_type given as {pt} must be one of `epsilon`, or `v_prediction`")
        derivative = (sample - x0) / sigma_hat
        dt = self.sigmas[self._step_index + 1] - sigma_hat
        prev = (sample + derivative * dt).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return SimpleNamespace(prev_sample=prev, pred_original_sample=x0)

    def step_direct_fusion(self, model_output, timestep, sample, generator=None):
        """The ``direct_fusion`` branch of the reference's trans pipelines, restated
        (``pipeline/pipeline_stable_video_diffusion_trans_controlnet.py:639-667``): forward and time-reversed halves of
        the batch share one blended denoised prediction."""
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        torch.randn(model_output.shape, dtype=model_output.dtype, generator=generator)     # drawn, never used (:646-648)
        x0 = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))          # :655
        fwd, bwd = x0.chunk(2)
        w = torch.linspace(1, 0, fwd.shape[1]).to(fwd.device).unsqueeze(0)
        w = w[(...,) + (None,) * (fwd.ndim - w.ndim)]
        x0 = fwd * w + bwd.flip(dims=[1]) * (1 - w)                                                  # :661
        x0 = torch.cat([x0, x0.flip(dims=[1])], dim=0)
        derivative = (sample - x0) / sigma
        dt = self.sigmas[self._step_index + 1] - sigma
        self._step_index += 1
        return sample + derivative * dt

    def add_noise(self, original_samples, noise, timesteps):
        sigmas = self.sigmas.to(device=original_samples.device, dtype=original_samples.dtype)
        sched_t = self.timesteps.to(original_samples.device)
        idx = [(sched_t == t).nonzero().item() for t in timesteps.to(original_samples.device)]
        sigma = sigmas[idx].flatten()
        while sigma.ndim < original_samples.ndim:
            sigma = sigma.unsqueeze(-1)
        return original_samples + noise * sigma

    def __len__(self):
        return self.config.num_train_timesteps
