import torch


class VaeImageProcessor:
    """Tensor inputs only: ``preprocess`` of a [N,C,H,W] tensor already in [-1, 1] at the target size is the
    identity (the golden generator feeds tensors, never PIL images)."""

    def __init__(self, vae_scale_factor=8, **kw):
        self.vae_scale_factor = vae_scale_factor

    def preprocess(self, image, height=None, width=None):
        if not isinstance(image, torch.Tensor):
            raise TypeError("shim VaeImageProcessor takes tensors only")
        if height is not None and tuple(image.shape[-2:]) != (height, width):
            raise ValueError("shim VaeImageProcessor does not resize")
        return image
