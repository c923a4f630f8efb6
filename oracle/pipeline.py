"""Oracle sampling loop / training wrapper on latents (TEST INFRA ONLY; no VAE / CLIP).

Restates reference ``pipeline/pipeline_stable_video_diffusion_controlnet.py``: guidance ramp :553-558,
denoise loop :577-630 (CFG dup :579, scale :580, concat :584, ControlNet :585-594, UNet :599-607, CFG
combine :614-616, scheduler step :619), ``_get_add_time_ids`` :239-266; and the training
preconditioning / loss of ``train_models/train_svd_lora.py:1503-1530,1651-1672`` with
``utils/util.py:250-308`` (``_get_add_time_ids``, training order, quirk F12)."""
from __future__ import annotations

from typing import Optional

import torch


def guidance_ramp(min_guidance_scale: float, max_guidance_scale: float, num_frames: int, batch: int = 1,
                  dtype=torch.float32) -> torch.Tensor:
    g = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(dtype)
    return g.repeat(batch, 1)[:, :, None, None, None]            # [B,F,1,1,1]


def cfg_combine(noise_pred: torch.Tensor, guidance: torch.Tensor) -> torch.Tensor:
    uncond, cond = noise_pred.chunk(2)
    return uncond + guidance * (cond - uncond)


def add_time_ids_inference(fps: float, motion_bucket_id: float, noise_aug_strength: float, batch: int,
                           do_cfg: bool = True, dtype=torch.float32) -> torch.Tensor:
    """Inference order [fps, motion_bucket, noise_aug] (pipeline :250); callers pass ``fps - 1``."""
    ids = torch.tensor([[fps, motion_bucket_id, noise_aug_strength]], dtype=dtype).repeat(batch, 1)
    return torch.cat([ids, ids]) if do_cfg else ids


def add_time_ids_training(fps: float, motion_bucket_id: float, noise_aug_strength: float, batch: int,
                          dtype=torch.float32) -> torch.Tensor:
    """Training order [fps, noise_aug, motion_bucket] (``utils/util.py:295``) - kept, not "fixed" (F12)."""
    return torch.tensor([[fps, noise_aug_strength, motion_bucket_id]], dtype=dtype).repeat(batch, 1)


@torch.no_grad()
def denoise_loop(unet, scheduler, latents, image_latents, image_embeddings, added_time_ids,
                 num_inference_steps: int = 25, min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                 controlnet=None, controlnet_cond=None, controlnet_cond_scale: float = 1.0,
                 unet_extra_args: tuple = (), return_trajectory: bool = False, max_steps: Optional[int] = None,
                 direct_fusion: bool = False):
    """``latents`` [S,F,4,h,w] (already scaled by init_noise_sigma); ``image_latents`` / ``image_embeddings`` /
    ``added_time_ids`` already CFG-duplicated ([2S,...], uncond first).  ``unet_extra_args`` are the LKGD
    positional (domain_features, flow_features)."""
    do_cfg = max_guidance_scale > 1.0
    scheduler.set_timesteps(num_inference_steps)
    g = guidance_ramp(min_guidance_scale, max_guidance_scale, latents.shape[1], latents.shape[0], latents.dtype)
    traj, preds = [], []
    for i, t in enumerate(scheduler.timesteps):
        if max_steps is not None and i >= max_steps:
            break
        x = torch.cat([latents] * 2) if do_cfg else latents
        x = scheduler.scale_model_input(x, t)
        x = torch.cat([x, image_latents], dim=2)
        kw = {}
        if controlnet is not None:
            down, mid = controlnet(x, t, encoder_hidden_states=image_embeddings, controlnet_cond=controlnet_cond,
                                   added_time_ids=added_time_ids, conditioning_scale=controlnet_cond_scale,
                                   guess_mode=False, return_dict=False)
            kw = dict(down_block_additional_residuals=down, mid_block_additional_residual=mid)
        noise_pred = unet(x, t, image_embeddings, *unet_extra_args, added_time_ids=added_time_ids,
                          return_dict=False, **kw)[0]
        if do_cfg:
            noise_pred = cfg_combine(noise_pred, g)
        if direct_fusion:      # pipeline_stable_video_diffusion_trans_controlnet.py:637-667 (forward half | time-reversed half)
            latents = scheduler.step_direct_fusion(noise_pred, t, latents)
        else:
            latents = scheduler.step(noise_pred, t, latents).prev_sample
        if return_trajectory:
            preds.append(noise_pred.clone())
            traj.append(latents.clone())
    return (latents, preds, traj) if return_trajectory else latents


def train_precondition(latents, noise, sigmas):
    """``train_svd_lora.py:1525-1530``: noisy = x + n*sigma; t = 0.25 ln sigma; input = noisy / sqrt(sigma^2+1)."""
    s = sigmas.reshape(-1, *([1] * (latents.ndim - 1)))
    noisy = latents + noise * s
    timesteps = torch.Tensor([0.25 * v.log() for v in sigmas])
    return noisy, timesteps, noisy / ((s ** 2 + 1) ** 0.5)


def train_loss(model_pred, noisy_latents, target, sigmas):
    """``train_svd_lora.py:1651-1672``: EDM v-pred wrapper and the (1+s^2)/s^2 weighted MSE in fp32."""
    s = sigmas.reshape(-1, *([1] * (target.ndim - 1)))
    c_out = -s / ((s ** 2 + 1) ** 0.5)
    c_skip = 1 / (s ** 2 + 1)
    denoised = model_pred * c_out + c_skip * noisy_latents
    w = (1 + s ** 2) * (s ** -2.0)
    loss = torch.mean((w.float() * (denoised.float() - target.float()) ** 2).reshape(target.shape[0], -1), dim=1)
    return loss.mean()


def smooth_chunks(total_frames: int, num_frames: int, rng=None):
    """``get_chunks`` of the smooth pipeline (pipeline/pipeline_stable_video_diffusion_smooth.py:526-534): a first chunk
    of random length 1..num_frames, then chunks of ``num_frames``.  ``rng``: object with ``randint`` (default
    ``numpy.random``, the reference's global stream)."""
    import numpy as np
    rng = np.random if rng is None else rng
    x_index = torch.arange(total_frames)
    rand_first = rng.randint(0, num_frames) + 1
    chunks = x_index[rand_first:].split(num_frames, dim=0)
    chunks = [x_index[:rand_first]] + list(chunks) if len(chunks[0]) > 0 else [x_index[:rand_first]]
    return [[int(i) for i in chunk] for chunk in chunks]


@torch.no_grad()
def smooth_loop(unet, scheduler, original_image_latents, noise, image_latents, image_embeddings, added_time_ids,
                num_frames: int, start_step: int, num_inference_steps: int = 25, min_guidance_scale: float = 1.0,
                max_guidance_scale: float = 3.0, rng=None, return_chunks: bool = False):
    """Denoising part of the reference's ``smooth`` pipeline (pipeline/pipeline_stable_video_diffusion_smooth.py:
    :520 add_noise at ``start_step``, :526-534 random chunking per step, :546-590 per chunk a [chunk, flipped chunk] pair
    with first / last frame conditioning, CFG, the forward half's prediction kept, :593 one Euler step over all frames).
    ``original_image_latents`` [1,T,4,h,w] (scaled VAE latents), ``image_latents`` / ``image_embeddings`` per FRAME
    [2T,...] (uncond first), ``added_time_ids`` [2,3] (CFG-duplicated once; the loop duplicates it again, :541)."""
    T = original_image_latents.shape[1]
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    latents = scheduler.add_noise(original_image_latents, noise, timesteps[[start_step]])
    ids4 = torch.cat([added_time_ids] * 2, dim=0)
    used = []
    for i in range(start_step, len(timesteps)):
        t = timesteps[i]
        chunks = smooth_chunks(T, num_frames, rng)
        used.append(chunks)
        noise_pred = torch.empty_like(latents)
        for chunk in chunks:
            lc = latents[:, chunk]
            lc = torch.cat([lc, lc.flip(dims=[1])], dim=0)
            first = [chunk[0], chunk[-1], chunk[0] + T, chunk[-1] + T]
            cur_lat = image_latents[first].unsqueeze(1).repeat(1, len(chunk), 1, 1, 1)
            cur_emb = image_embeddings[first]
            x = scheduler.scale_model_input(torch.cat([lc] * 2), t)
            x = torch.cat([x, cur_lat], dim=2)
            pred = unet(x, t, cur_emb, added_time_ids=ids4, return_dict=False)[0]
            g = torch.linspace(min_guidance_scale, max_guidance_scale, len(chunk)).unsqueeze(0)[..., None, None, None]
            u, c = pred.chunk(2)
            pred = u + g * (c - u)
            noise_pred[:, chunk] = pred[:len(pred) // 2]
        latents = scheduler.step(noise_pred, t, latents).prev_sample
    return (latents, used) if return_chunks else latents
