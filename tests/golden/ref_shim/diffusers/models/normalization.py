"""Names patch/patch.py imports (`from diffusers.models.normalization import ...`); none is used on the
``norm_type == "layer_norm"`` path the SVD blocks take (patch/patch.py:415-416)."""
import torch.nn as nn


class AdaLayerNorm(nn.Module):
    pass


class AdaLayerNormContinuous(nn.Module):
    pass


class AdaLayerNormZero(nn.Module):
    pass


class RMSNorm(nn.Module):
    pass
