"""TEST INFRASTRUCTURE (CPU oracle) - ``AutoencoderKLTemporalDecoder`` of the SVD pipelines (VAE half).

The class lives in the reference's UN-VENDORED dependency ``diffusers==0.27.2``
(``models/autoencoders/autoencoder_kl_temporal_decoder.py``, ``models/autoencoders/vae.py`` ``Encoder``,
``models/unets/unet_3d_blocks.py`` ``MidBlockTemporalDecoder`` / ``UpBlockTemporalDecoder``, ``models/unets/unet_2d_blocks.py``
``DownEncoderBlock2D`` / ``UNetMidBlock2D``); diffusers is not installed in this image, so its published algorithm is restated
here on the oracle's block classes.  **Parity partly pinned**: there is neither a diffusers installation nor a
reference-held fixture, but the image holds an INDEPENDENT implementation of the KL-autoencoder this VAE descends from
(torchtitan's Flux autoencoder = the CompVis latent-diffusion Encoder / Decoder).  Under the published LDM -> diffusers key
mapping the ENCODER here equals it to 2e-6 and the TEMPORAL DECODER with its temporal branch switched off equals the LDM
decoder frame by frame (``tests/test_oracle_golden.py::test_vae_oracle_spatial_skeleton_matches_an_independent_ldm_autoencoder``):
resnet arithmetic, bottom / right padded downsampling, the single-head attention block, block order, channel wiring and
upsampler placement are pinned.  **Unpinned (restated only)**: the temporal branch (TemporalResnetBlock, the switched learned
AlphaBlender, Conv3d ``time_conv_out``) and ``quant_conv``.  Further anchors are the reference's call sites
(``pipeline/pipeline_stable_video_diffusion_controlnet.py:216-237`` ``_encode_vae_image`` = ``vae.encode(x).latent_dist.mode()``;
``:268-295`` ``decode_latents`` = ``vae.decode(z / scaling_factor, num_frames=chunk).sample`` in chunks of
``decode_chunk_size`` frames; training ``utils/util.py:234-248`` ``tensor_to_vae_latent`` = ``encode(x).latent_dist.sample() *
scaling_factor``) and the parameter names / shapes of the SVD checkpoint's ``vae/`` folder
(``tests/test_oracle_golden.py::test_vae_state_dict_names``).  The building blocks themselves (ResnetBlock2D,
TemporalResnetBlock, AlphaBlender, Upsample2D) are the ones the UNet goldens exercise.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU legs may import this package."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Sequence

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import Downsample2D, ResnetBlock2D, SpatioTemporalResBlock, Upsample2D

SVD_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                      latent_channels=4, sample_size=768, scaling_factor=0.18215, force_upcast=True)


class VaeAttention(nn.Module):
    """diffusers ``Attention(query_dim=C, heads=C // dim_head, dim_head, eps, norm_num_groups=32, bias=True,
    residual_connection=True)`` on a 4-D input: GroupNorm over the tokens' channels, biased q / k / v, softmax attention over
    the H*W tokens of one image, biased out-projection, residual."""

    def __init__(self, channels: int, dim_head: int, eps: float = 1e-6, groups: int = 32):
        super().__init__()
        self.heads, self.dim_head = channels // dim_head, dim_head
        self.group_norm = nn.GroupNorm(groups, channels, eps=eps, affine=True)
        self.to_q, self.to_k, self.to_v = (nn.Linear(channels, channels, bias=True) for _ in range(3))
        self.to_out = nn.ModuleList([nn.Linear(channels, channels, bias=True), nn.Dropout(0.0)])

    def forward(self, x):
        b, c, h, w = x.shape
        t = self.group_norm(x.view(b, c, h * w)).transpose(1, 2)                # [B, HW, C]
        q, k, v = (f(t).view(b, h * w, self.heads, self.dim_head).transpose(1, 2) for f in (self.to_q, self.to_k, self.to_v))
        o = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(b, h * w, c)
        o = self.to_out[0](o)
        return o.transpose(-1, -2).reshape(b, c, h, w) + x


class DownEncoderBlock2D(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(in_channels if i == 0 else out_channels, out_channels, None, 1e-6)
                                      for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels, padding=0)]) if add_downsample else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x, None)
        if self.downsamplers is not None:
            x = self.downsamplers[0](x)
        return x


class UNetMidBlock2D(nn.Module):
    def __init__(self, channels, attention_head_dim):
        super().__init__()
        self.resnets = nn.ModuleList([ResnetBlock2D(channels, channels, None, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([VaeAttention(channels, attention_head_dim)])

    def forward(self, x):
        x = self.resnets[0](x, None)
        x = self.attentions[0](x)
        return self.resnets[1](x, None)


class Encoder(nn.Module):
    """``Encoder(in, latent, 4 x DownEncoderBlock2D, block_out_channels, layers_per_block, double_z=True)``."""

    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        self.conv_in = nn.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        blocks, c = [], block_out_channels[0]
        for i, co in enumerate(block_out_channels):
            blocks.append(DownEncoderBlock2D(c, co, layers_per_block, add_downsample=i != len(block_out_channels) - 1))
            c = co
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(c, attention_head_dim=c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * out_channels, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


def _temporal_resblock(cin, cout):
    return SpatioTemporalResBlock(cin, cout, None, 1e-6, temporal_eps=1e-5, merge_factor=0.0, merge_strategy="learned",
                                  switch_spatial_to_temporal_mix=True)


class MidBlockTemporalDecoder(nn.Module):
    def __init__(self, channels, attention_head_dim, num_layers):
        super().__init__()
        self.resnets = nn.ModuleList([_temporal_resblock(channels, channels) for _ in range(num_layers)])
        self.attentions = nn.ModuleList([VaeAttention(channels, attention_head_dim)])

    def forward(self, x, image_only_indicator):
        x = self.resnets[0](x, None, image_only_indicator)
        for r, a in zip(self.resnets[1:], self.attentions):
            x = r(a(x), None, image_only_indicator)
        return x


class UpBlockTemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, num_layers, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_temporal_resblock(in_channels if i == 0 else out_channels, out_channels)
                                      for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, x, image_only_indicator):
        for r in self.resnets:
            x = r(x, None, image_only_indicator)
        if self.upsamplers is not None:
            x = self.upsamplers[0](x)
        return x


class TemporalDecoder(nn.Module):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        top = block_out_channels[-1]
        self.conv_in = nn.Conv2d(in_channels, top, 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(top, attention_head_dim=top, num_layers=layers_per_block)
        rev, blocks, c = list(reversed(block_out_channels)), [], top
        for i, co in enumerate(rev):
            blocks.append(UpBlockTemporalDecoder(c, co, layers_per_block + 1, add_upsample=i != len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = nn.GroupNorm(32, block_out_channels[0], eps=1e-6)
        self.conv_out = nn.Conv2d(block_out_channels[0], out_channels, 3, padding=1)
        self.time_conv_out = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, image_only_indicator, num_frames):
        x = self.mid_block(self.conv_in(z), image_only_indicator)
        for b in self.up_blocks:
            x = b(x, image_only_indicator)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        x = x[None, :].reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class DiagonalGaussianDistribution:
    def __init__(self, parameters: torch.Tensor):
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self):
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None, noise: Optional[torch.Tensor] = None):
        if noise is None:
            noise = torch.randn(self.mean.shape, generator=generator, dtype=self.mean.dtype)
        return self.mean + self.std * noise


class AutoencoderKLTemporalDecoder(nn.Module):
    def __init__(self, in_channels=3, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block=2, latent_channels=4, sample_size=768, scaling_factor=0.18215, force_upcast=True, **_):
        super().__init__()
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels,
                                      block_out_channels=tuple(block_out_channels), layers_per_block=layers_per_block,
                                      latent_channels=latent_channels, sample_size=sample_size, scaling_factor=scaling_factor,
                                      force_upcast=force_upcast)
        self.encoder = Encoder(in_channels, latent_channels, block_out_channels, layers_per_block)
        self.decoder = TemporalDecoder(latent_channels, out_channels, block_out_channels, layers_per_block)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)

    def encode(self, x):
        return SimpleNamespace(latent_dist=DiagonalGaussianDistribution(self.quant_conv(self.encoder(x))))

    def decode(self, z, num_frames: int):
        ioi = torch.zeros(z.shape[0] // num_frames, num_frames, dtype=z.dtype)
        return SimpleNamespace(sample=self.decoder(z, ioi, num_frames))


def decode_latents(vae, latents, num_frames: int, decode_chunk_size: int = 14):
    """Reference ``decode_latents`` (pipeline...controlnet.py:268-295): [B, F, C, h, w] -> fp32 [B, 3, F, 8h, 8w]."""
    latents = latents.flatten(0, 1) / vae.config.scaling_factor
    frames = []
    for i in range(0, latents.shape[0], decode_chunk_size):
        chunk = latents[i:i + decode_chunk_size]
        frames.append(vae.decode(chunk, num_frames=chunk.shape[0]).sample)
    frames = torch.cat(frames, dim=0)
    return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
