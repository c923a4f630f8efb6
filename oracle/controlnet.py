"""Oracle ControlNetSDV (TEST INFRA ONLY).  Follows reference ``models/controlnet_sdv.py``:
``ControlNetConditioningEmbeddingSVD`` :64-119, ``ControlNetSDVModel.__init__`` :160-316,
``.forward`` :441-578, ``.from_unet`` :581-638, ``zero_module`` :804-807."""
from __future__ import annotations

from types import SimpleNamespace
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import (CrossAttnDownBlockSpatioTemporal, DownBlockSpatioTemporal, TimestepEmbedding,
                     UNetMidBlockSpatioTemporal)
from .unet import UNetSpatioTemporalConditionControlNetModel, _tuple


def zero_module(m: nn.Module) -> nn.Module:
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class ControlNetConditioningEmbeddingSVD(nn.Module):
    """Pixel-resolution condition encoder: conv3x3 + SiLU, 3x {conv3x3, SiLU, conv3x3 stride 2, SiLU},
    zero-init conv3x3 to ``conditioning_embedding_channels`` at H/8 (reference :64-119)."""

    def __init__(self, conditioning_embedding_channels: int, conditioning_channels: int = 3,
                 block_out_channels: Tuple[int, ...] = (16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = zero_module(nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, 3, padding=1))

    def forward(self, conditioning):
        b, f, c, h, w = conditioning.shape
        x = F.silu(self.conv_in(conditioning.reshape(b * f, c, h, w)))
        for blk in self.blocks:
            x = F.silu(blk(x))
        return self.conv_out(x)


class ControlNetSDVModel(UNetSpatioTemporalConditionControlNetModel):
    """Copy of the UNet encoder + mid block, a condition encoder added after ``conv_in``,
    12 zero-init 1x1 convs over the skip list and a mid 1x1 (reference :219-316, :441-578)."""

    def __init__(self, sample_size=None, in_channels=8, out_channels=4,
                 down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
                 block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
                 transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20), num_frames=25,
                 conditioning_channels=3, conditioning_embedding_out_channels=(16, 32, 96, 256),
                 time_context_order="hw_major_0272"):
        nn.Module.__init__(self)
        n = len(down_block_types)
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), block_out_channels=tuple(block_out_channels),
            addition_time_embed_dim=addition_time_embed_dim,
            projection_class_embeddings_input_dim=projection_class_embeddings_input_dim,
            layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
            transformer_layers_per_block=transformer_layers_per_block, num_attention_heads=num_attention_heads,
            num_frames=num_frames, conditioning_channels=conditioning_channels,
            conditioning_embedding_out_channels=tuple(conditioning_embedding_out_channels),
            time_context_order=time_context_order)
        heads, xdim = _tuple(num_attention_heads, n), _tuple(cross_attention_dim, n)
        lpb, tlpb = _tuple(layers_per_block, n), _tuple(transformer_layers_per_block, n)
        c0 = block_out_channels[0]
        temb = c0 * 4
        self.conv_in = nn.Conv2d(in_channels, c0, 3, padding=1)
        self.time_embedding = TimestepEmbedding(c0, temb)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList()
        self.controlnet_cond_embedding = ControlNetConditioningEmbeddingSVD(
            c0, conditioning_channels, tuple(conditioning_embedding_out_channels))
        out_c = c0
        self.controlnet_down_blocks.append(zero_module(nn.Conv2d(out_c, out_c, 1)))
        for i, t in enumerate(down_block_types):
            in_c, out_c = out_c, block_out_channels[i]
            last = i == n - 1
            if t == "CrossAttnDownBlockSpatioTemporal":
                blk = CrossAttnDownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], tlpb[i], heads[i], xdim[i],
                                                       add_downsample=not last,
                                                       time_context_order=time_context_order)
            else:
                blk = DownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], add_downsample=not last)
            self.down_blocks.append(blk)
            for _ in range(lpb[i] + (0 if last else 1)):
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(out_c, out_c, 1)))
        self.controlnet_mid_block = zero_module(nn.Conv2d(block_out_channels[-1], block_out_channels[-1], 1))
        self.mid_block = UNetMidBlockSpatioTemporal(block_out_channels[-1], temb, 1, tlpb[-1], heads[-1], xdim[-1],
                                                    time_context_order=time_context_order)

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond=None,
                image_only_indicator=None, return_dict=True, guess_mode=False, conditioning_scale=1.0):
        # image_only_indicator / guess_mode are accepted and ignored (reference quirk D7, :448,450,530)
        b, f = sample.shape[:2]
        emb = self._time_embedding(sample, timestep, added_time_ids)
        ctx = encoder_hidden_states.repeat_interleave(f, dim=0)
        x = self.conv_in(sample.flatten(0, 1))
        if controlnet_cond is not None:
            x = x + self.controlnet_cond_embedding(controlnet_cond)
        iof = torch.zeros(b, f, dtype=x.dtype, device=x.device)
        skips = (x,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                x, res = blk(x, emb, ctx, iof)
            else:
                x, res = blk(x, emb, iof)
            skips += res
        x = self.mid_block(x, emb, ctx, iof)
        down = [conv(s) * conditioning_scale for s, conv in zip(skips, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(x) * conditioning_scale
        if not return_dict:
            return (down, mid)
        return SimpleNamespace(down_block_res_samples=down, mid_block_res_sample=mid)

    @classmethod
    def from_unet(cls, unet, conditioning_embedding_out_channels=(16, 32, 96, 256), load_weights_from_unet=True,
                  conditioning_channels=3):
        c = unet.config
        net = cls(in_channels=c.in_channels, down_block_types=c.down_block_types,
                  block_out_channels=c.block_out_channels, addition_time_embed_dim=c.addition_time_embed_dim,
                  transformer_layers_per_block=c.transformer_layers_per_block,
                  cross_attention_dim=c.cross_attention_dim, num_attention_heads=c.num_attention_heads,
                  num_frames=c.num_frames, sample_size=c.sample_size, layers_per_block=c.layers_per_block,
                  projection_class_embeddings_input_dim=c.projection_class_embeddings_input_dim,
                  conditioning_channels=conditioning_channels,
                  conditioning_embedding_out_channels=conditioning_embedding_out_channels,
                  time_context_order=getattr(c, "time_context_order", "hw_major_0272"))
        if load_weights_from_unet:
            for name in ("conv_in", "time_embedding", "add_embedding", "down_blocks", "mid_block"):
                getattr(net, name).load_state_dict(getattr(unet, name).state_dict())
        return net
