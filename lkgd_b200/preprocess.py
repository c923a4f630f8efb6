"""Image-side glue of the reference pipelines' ``_encode_image`` (pipeline/pipeline_stable_video_diffusion_controlnet.py:174-214):
anti-aliased resize to the CLIP resolution (:672-784: Gaussian blur with sigma = (factor - 1) / 2 and a 2-sigma window,
reflect padding, then bicubic ``interpolate(align_corners=True)``) and the CLIP normalisation of the feature extractor
(``do_normalize`` only, :190-197).  One conditioning image per video, once per call - OFF the denoise hot path: plain torch
tensor ops on whatever device the image is on (no kernels of ours, nothing to accelerate)."""
from __future__ import annotations

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)          # transformers CLIPImageProcessor defaults
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


def _gaussian(window: int, sigma: float, dtype, device):
    x = torch.arange(window, device=device, dtype=dtype) - window // 2
    if window % 2 == 0:
        x = x + 0.5
    g = torch.exp(-x.pow(2.0) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x, kernel, horizontal: bool):
    b, c, h, w = x.shape
    k = kernel.numel()
    front = (k - 1) // 2
    rear = (k - 1) - front
    pad = (front, rear, 0, 0) if horizontal else (0, 0, front, rear)
    x = F.pad(x, pad, mode="reflect")
    wgt = kernel.reshape(1, 1, 1, k) if horizontal else kernel.reshape(1, 1, k, 1)
    return F.conv2d(x, wgt.expand(c, 1, -1, -1), groups=c)


def resize_with_antialiasing(image: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """``_resize_with_antialiasing(input, size, "bicubic", align_corners=True)`` of the reference (:672-703)."""
    if image.ndim == 3:
        image = image.unsqueeze(0)
    h, w = image.shape[-2:]
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = [int(max(2.0 * 2 * sigmas[0], 3)), int(max(2.0 * 2 * sigmas[1], 3))]
    ks = [k + 1 if k % 2 == 0 else k for k in ks]
    x = _filter(image, _gaussian(ks[1], sigmas[1], image.dtype, image.device), horizontal=True)
    x = _filter(x, _gaussian(ks[0], sigmas[0], image.dtype, image.device), horizontal=False)
    return F.interpolate(x, size=size, mode="bicubic", align_corners=True)


def clip_pixel_values(image: torch.Tensor, size=(224, 224)) -> torch.Tensor:
    """image in [0, 1], [B, 3, H, W] (or [3, H, W]) -> what the reference feeds its image encoder (:181-197)."""
    x = resize_with_antialiasing(image * 2.0 - 1.0, size)
    x = (x + 1.0) / 2.0
    mean = torch.tensor(CLIP_MEAN, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=x.dtype, device=x.device).view(1, 3, 1, 1)
    return (x - mean) / std


@torch.no_grad()
def encode_image(image_encoder, image: torch.Tensor, num_videos_per_prompt: int = 1,
                 do_classifier_free_guidance: bool = True) -> torch.Tensor:
    """``_encode_image`` (:174-214): -> ``image_embeddings`` [(2) * B * nvpp, 1, D], zero unconditional half first."""
    pv = clip_pixel_values(image.to(torch.float32))
    emb = image_encoder(pv.to(image_encoder.device)).image_embeds.unsqueeze(1)
    bs, seq, _ = emb.shape
    emb = emb.repeat(1, num_videos_per_prompt, 1).view(bs * num_videos_per_prompt, seq, -1)
    if do_classifier_free_guidance:
        emb = torch.cat([torch.zeros_like(emb), emb])
    return emb
