"""Oracle restatement of the diffusers==0.27.2 spatio-temporal building blocks (TEST INFRA ONLY).

The reference imports these from the un-vendored dependency
(reference ``models/unet_spatio_temporal_condition_controlnet.py:11-13``); nothing here can be
cited into /root/reference except where noted.  Parameter/module names follow the reference's
parameter dumps ``train_svd_lora.txt`` / ``train_svd_lora_train.txt`` exactly.

Un-checkable recollections are switches:
  U1  GroupNorm eps per block family       -> ``eps`` arguments of the block factories
  U2  temporal cross-attn context ordering -> ``time_context_order``
  U3  attention = softmax(q k^T / sqrt(d)) v, no mask
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

__all__ = [
    "timestep_embedding", "TimestepEmbedding", "Attention", "GEGLU", "FeedForward",
    "BasicTransformerBlock", "TemporalBasicTransformerBlock", "AlphaBlender",
    "TransformerSpatioTemporalModel", "ResnetBlock2D", "TemporalResnetBlock",
    "SpatioTemporalResBlock", "Downsample2D", "Upsample2D", "DownBlockSpatioTemporal",
    "CrossAttnDownBlockSpatioTemporal", "UNetMidBlockSpatioTemporal", "UpBlockSpatioTemporal",
    "CrossAttnUpBlockSpatioTemporal",
]


# --------------------------------------------------------------------------- embeddings (A.1)
def timestep_embedding(t: torch.Tensor, dim: int, flip_sin_to_cos: bool = True,
                       downscale_freq_shift: float = 0.0, max_period: float = 10000.0) -> torch.Tensor:
    """``Timesteps(dim, flip_sin_to_cos, downscale_freq_shift)``; ctor args at reference
    ``models/unet_spatio_temporal_condition_controlnet.py:137,142``.  Always fp32."""
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32, device=t.device)
    freqs = torch.exp(-math.log(max_period) * k / (half - downscale_freq_shift))
    arg = t.reshape(-1).float()[:, None] * freqs[None, :]
    emb = torch.cat([torch.sin(arg), torch.cos(arg)], dim=-1)
    if flip_sin_to_cos:
        emb = torch.cat([emb[:, half:], emb[:, :half]], dim=-1)
    if dim % 2 == 1:
        emb = F.pad(emb, (0, 1))
    return emb


class TimestepEmbedding(nn.Module):
    def __init__(self, in_channels: int, time_embed_dim: int, out_dim: Optional[int] = None):
        super().__init__()
        self.linear_1 = nn.Linear(in_channels, time_embed_dim)
        self.linear_2 = nn.Linear(time_embed_dim, out_dim or time_embed_dim)

    def forward(self, x):
        return self.linear_2(F.silu(self.linear_1(x)))


# --------------------------------------------------------------------------- attention (A.6)
class Attention(nn.Module):
    """``Attention(query_dim, heads, dim_head, cross_attention_dim)`` with the default
    ``AttnProcessor2_0`` (U3): scale d^-0.5, no mask, non-causal, no residual inside."""

    def __init__(self, query_dim: int, heads: int, dim_head: int, cross_attention_dim: Optional[int] = None):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        kv_dim = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(kv_dim, inner, bias=False)
        self.to_v = nn.Linear(kv_dim, inner, bias=False)
        self.to_out = nn.ModuleList([nn.Linear(inner, query_dim), nn.Dropout(0.0)])

    # softmax(QK^T) of a 72x128 latent frame (9216 tokens) is 1.7 GB per image and head group: self-attention over
    # that many tokens is evaluated image by image (same arithmetic, bounded memory) - full-size parity runs only
    CHUNK_TOKENS = 4096

    def forward(self, x, encoder_hidden_states=None):
        if encoder_hidden_states is None and x.shape[1] >= self.CHUNK_TOKENS and x.shape[0] > 1:
            return torch.cat([self.forward(x[i:i + 1]) for i in range(x.shape[0])], 0)
        ctx = x if encoder_hidden_states is None else encoder_hidden_states
        b, n, _ = x.shape
        q = self.to_q(x).view(b, n, self.heads, self.dim_head).transpose(1, 2)
        k = self.to_k(ctx).view(b, -1, self.heads, self.dim_head).transpose(1, 2)
        v = self.to_v(ctx).view(b, -1, self.heads, self.dim_head).transpose(1, 2)
        w = torch.softmax((q @ k.transpose(-1, -2)) * (self.dim_head ** -0.5), dim=-1)
        o = (w @ v).transpose(1, 2).reshape(b, n, self.heads * self.dim_head)
        return self.to_out[0](o)


class GEGLU(nn.Module):
    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        h, gate = self.proj(x).chunk(2, dim=-1)
        return h * F.gelu(gate)  # exact (erf) gelu


class FeedForward(nn.Module):
    def __init__(self, dim: int, dim_out: Optional[int] = None, mult: int = 4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class BasicTransformerBlock(nn.Module):
    """Spatial block; forward restated in-tree at reference ``patch/patch.py:390-580``
    (``norm_type == "layer_norm"`` branches, ``enable_joint_attention=False``)."""

    def __init__(self, dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)

    # ---- joint attention between paired samples (reference ``patch/patch.py`` ToMeBlock): off by default
    enable_joint_attention = False
    joint_scale = 1.0
    flip = False                 # ``_tome_info["args"]["flip"]`` (:460-464)
    num_frames = None            # ``_tome_info["size"][1]``

    def initialize_joint_layers(self, post="conv"):
        """``ToMeBlock.initialize_joint_layers`` (:143-172): a copy of attn1 + a zero-initialised post layer."""
        import copy
        self.attn1n = copy.deepcopy(self.attn1)
        c = self.attn1.to_out[0].out_features
        if post == "scale":
            self.scale1n = nn.Parameter(torch.zeros(1, 1, c))
        elif post == "conv":
            self.conv1n = nn.Linear(c, c, bias=False)
            nn.init.zeros_(self.conv1n.weight)
        elif post == "conv_fuse":                      # :154-157: one [2C, 2C] layer over [masked sample | its partner]
            self.conv1n = nn.Linear(2 * c, 2 * c, bias=False)
            nn.init.zeros_(self.conv1n.weight)
        else:
            raise ValueError(post)
        self.post = post

    def _partner(self, n, mask):
        """:452-456 - joint_enc[~m] = n[m]; joint_enc[m] = n[~m] with the mask repeat-interleaved over the batch."""
        m = mask.repeat_interleave(n.shape[0] // len(mask), dim=0)
        out = torch.empty_like(n)
        out[~m] = n[m]
        out[m] = n[~m]
        return out

    def _post(self, y):
        return self.conv1n(y) if self.post == "conv" else self.scale1n * y

    def _post_spatial(self, y):
        if self.post != "conv_fuse":
            return self._post(y)
        # :488-494 - the masked rows and the unmasked rows (k-th with k-th) go through ONE layer side by side
        m = self.joint_attn_mask.repeat_interleave(y.shape[0] // len(self.joint_attn_mask), dim=0)
        fx, fy = self.conv1n(torch.cat([y[m], y[~m]], dim=-1)).chunk(2, dim=-1)
        out = y.clone()
        out[m] = fx
        out[~m] = fy
        return out

    def forward(self, x, encoder_hidden_states):
        n = self.norm1(x)
        a = self.attn1(n)
        if self.enable_joint_attention:                                   # :434-492
            enc = self._partner(n, self.joint_attn_mask)
            if self.flip:
                f = self.num_frames
                enc = enc.reshape(-1, f, *enc.shape[1:]).flip(dims=[1]).reshape(enc.shape)
            a = a + self._post_spatial(self.attn1n(n, enc)) * self.joint_scale
        x = x + a
        x = x + self.attn2(self.norm2(x), encoder_hidden_states)
        x = x + self.ff(self.norm3(x))
        return x


class TemporalBasicTransformerBlock(nn.Module):
    """Temporal block; forward restated in-tree at reference ``patch/patch.py:582-686``
    (non-joint branch ``:659-661``).  ``is_res = dim == time_mix_inner_dim``."""

    def __init__(self, dim: int, time_mix_inner_dim: int, heads: int, dim_head: int, cross_attention_dim: int):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim, eps=1e-5)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim)
        self.norm1 = nn.LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.attn1 = Attention(time_mix_inner_dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.attn2 = Attention(time_mix_inner_dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.ff = FeedForward(time_mix_inner_dim)

    enable_joint_attention = False
    initialize_joint_layers = BasicTransformerBlock.initialize_joint_layers

    def forward(self, x, num_frames: int, encoder_hidden_states):
        bf, n, c = x.shape
        b = bf // num_frames
        x = x[None, :].reshape(b, num_frames, n, c).permute(0, 2, 1, 3).reshape(b * n, num_frames, c)
        res = x
        x = self.ff_in(self.norm_in(x))
        if self.is_res:
            x = x + res
        nh = self.norm1(x)
        a = self.attn1(nh)
        if self.enable_joint_attention:                                   # :617-658 (no joint_scale, no flip here)
            an = self.attn1n(nh, BasicTransformerBlock._partner(self, nh, self.joint_attn_mask))
            # :647-650 - the temporal forward knows "conv" and "scale" only: under "conv_fuse" there is NO post layer here
            a = a + (an if self.post == "conv_fuse" else BasicTransformerBlock._post(self, an))
        x = a + x
        x = self.attn2(self.norm2(x), encoder_hidden_states) + x
        y = self.ff(self.norm3(x))
        x = y + x if self.is_res else y
        x = x[None, :].reshape(b, n, num_frames, c).permute(0, 2, 1, 3).reshape(b * num_frames, n, c)
        return x


class AlphaBlender(nn.Module):
    """``AlphaBlender(alpha, "learned_with_images")``: alpha = where(iof, 1, sigmoid(mix_factor)) (the UNet);
    ``"learned"``: alpha = sigmoid(mix_factor) for every frame (the VAE's temporal decoder, which also sets
    ``switch_spatial_to_temporal_mix``: alpha -> 1 - alpha)."""

    def __init__(self, alpha: float = 0.5, merge_strategy: str = "learned_with_images",
                 switch_spatial_to_temporal_mix: bool = False):
        super().__init__()
        if merge_strategy not in ("learned", "learned_with_images"):
            raise ValueError(f"merge_strategy {merge_strategy!r}")
        self.merge_strategy, self.switch = merge_strategy, switch_spatial_to_temporal_mix
        self.mix_factor = nn.Parameter(torch.tensor([alpha], dtype=torch.float32))

    def forward(self, x_spatial, x_temporal, image_only_indicator):
        if self.merge_strategy == "learned":
            a = torch.sigmoid(self.mix_factor)
        else:
            a = torch.where(image_only_indicator.bool(),
                            torch.ones(1, 1, device=x_spatial.device, dtype=self.mix_factor.dtype),
                            torch.sigmoid(self.mix_factor)[..., None])
            if x_spatial.ndim == 5:
                a = a[:, None, :, None, None]
            elif x_spatial.ndim == 3:
                a = a.reshape(-1)[:, None, None]
        a = a.to(x_spatial.dtype)
        if self.switch:
            a = 1.0 - a
        return a * x_spatial + (1.0 - a) * x_temporal


class TransformerSpatioTemporalModel(nn.Module):
    """A.5.  ``time_context_order``: "hw_major_0272" reproduces the pinned diffusers 0.27.2 layout
    (row r of the temporal block sees ``ctx_first[r % B]``, U2/F8); "b_major" is the later fix."""

    def __init__(self, heads: int, dim_head: int, in_channels: int, num_layers: int = 1,
                 cross_attention_dim: int = 1024, time_context_order: str = "hw_major_0272"):
        super().__init__()
        inner = heads * dim_head
        self.in_channels = in_channels
        self.time_context_order = time_context_order
        # test hook for the CFG-pair split (tests/test_distributed_cpu.py): (first-frame contexts of the WHOLE batch
        # [B_total, L, D], index of this half's first batch element) - rows then index the contexts exactly as the
        # unsplit batch would under the 0.27.2 order
        self.cfg_split = None
        self.norm = nn.GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = nn.Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.temporal_transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(inner, inner, heads, dim_head, cross_attention_dim) for _ in range(num_layers)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_mixer = AlphaBlender(0.5)
        self.proj_out = nn.Linear(inner, in_channels)

    def forward(self, x, encoder_hidden_states, image_only_indicator):
        bf, c, h, w = x.shape
        f = image_only_indicator.shape[-1]
        b = bf // f
        ctx = encoder_hidden_states
        ctx_first = ctx[None, :].reshape(b, f, -1, ctx.shape[-1])[:, 0]              # [B, L, D]
        if self.time_context_order == "hw_major_0272":
            time_context = ctx_first[None, :].broadcast_to(h * w, b, ctx_first.shape[1], ctx.shape[-1])
        elif self.time_context_order == "b_major":
            time_context = ctx_first[:, None].broadcast_to(b, h * w, ctx_first.shape[1], ctx.shape[-1])
        else:
            raise ValueError(self.time_context_order)
        time_context = time_context.reshape(h * w * b, ctx_first.shape[1], ctx.shape[-1])
        if self.cfg_split is not None and self.time_context_order == "hw_major_0272":
            full, b0 = self.cfg_split
            rows = torch.arange(b * h * w, device=x.device) + b0 * h * w          # row index inside the unsplit batch
            time_context = full[rows % full.shape[0]]

        res = x
        x = self.norm(x)
        x = x.permute(0, 2, 3, 1).reshape(bf, h * w, c)
        x = self.proj_in(x)
        frame_idx = torch.arange(f, device=x.device).repeat(b, 1).reshape(-1)
        emb = self.time_pos_embed(timestep_embedding(frame_idx, self.in_channels).to(x.dtype))[:, None, :]
        for blk, tblk in zip(self.transformer_blocks, self.temporal_transformer_blocks):
            # keyword calls, as diffusers does: the reference's patched forwards (patch/patch.py:390-399, :582-587)
            # take (hidden_states, attention_mask, encoder_hidden_states, ...) positionally
            x = blk(x, encoder_hidden_states=encoder_hidden_states)
            xm = tblk(x + emb, num_frames=f, encoder_hidden_states=time_context)
            x = self.time_mixer(x, xm, image_only_indicator)
        x = self.proj_out(x)
        x = x.reshape(bf, h, w, c).permute(0, 3, 1, 2).contiguous()
        return x + res


# --------------------------------------------------------------------------- resnets (A.3, A.4)
class ResnetBlock2D(nn.Module):
    """``temb_channels=None`` (the VAE): no time-embedding projection."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int], eps: float, groups: int = 32):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, in_channels, eps=eps)
        self.conv1 = nn.Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(groups, out_channels, eps=eps)
        self.conv2 = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, None, None]
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class TemporalResnetBlock(nn.Module):
    """GroupNorm on the 5-D tensor: statistics over (C/32)*F*H*W, i.e. ACROSS frames."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int], eps: float):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, in_channels, eps=eps)
        self.conv1 = nn.Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = nn.Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = nn.GroupNorm(32, out_channels, eps=eps)
        self.conv2 = nn.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.conv_shortcut = nn.Conv3d(in_channels, out_channels, 1) if in_channels != out_channels else None

    def forward(self, x, temb=None):  # x [B,C,F,H,W], temb [B,F,T]
        h = self.conv1(F.silu(self.norm1(x)))
        if self.time_emb_proj is not None:
            h = h + self.time_emb_proj(F.silu(temb))[:, :, :, None, None].permute(0, 2, 1, 3, 4)
        h = self.conv2(F.silu(self.norm2(h)))
        if self.conv_shortcut is not None:
            x = self.conv_shortcut(x)
        return x + h


class SpatioTemporalResBlock(nn.Module):
    """UNet: ``(in, out, temb_channels, eps)``.  VAE temporal decoder: ``temb_channels=None, eps=1e-6, temporal_eps=1e-5,
    merge_factor=0.0, merge_strategy="learned", switch_spatial_to_temporal_mix=True``."""

    def __init__(self, in_channels: int, out_channels: int, temb_channels: Optional[int], eps: float,
                 temporal_eps: Optional[float] = None, merge_factor: float = 0.5,
                 merge_strategy: str = "learned_with_images", switch_spatial_to_temporal_mix: bool = False):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, out_channels, temb_channels,
                                                      temporal_eps if temporal_eps is not None else eps)
        self.time_mixer = AlphaBlender(merge_factor, merge_strategy, switch_spatial_to_temporal_mix)

    def forward(self, x, temb, image_only_indicator):
        f = image_only_indicator.shape[-1]
        x = self.spatial_res_block(x, temb)
        bf, c, h, w = x.shape
        b = bf // f
        xs = x[None, :].reshape(b, f, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(xs, temb.reshape(b, f, -1) if temb is not None else None)
        x = self.time_mixer(xs, xt, image_only_indicator)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


# --------------------------------------------------------------------------- sampling (A.10)
class Downsample2D(nn.Module):
    """``padding=0`` (the VAE encoder): the input is padded (0, 1, 0, 1) - right / bottom only - before the conv."""

    def __init__(self, channels: int, padding: int = 1):
        super().__init__()
        self.padding = padding
        self.conv = nn.Conv2d(channels, channels, 3, stride=2, padding=padding)

    def forward(self, x):
        if self.padding == 0:
            x = F.pad(x, (0, 1, 0, 1), mode="constant", value=0)
        return self.conv(x)


class Upsample2D(nn.Module):
    def __init__(self, channels: int):
        super().__init__()
        self.conv = nn.Conv2d(channels, channels, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


# --------------------------------------------------------------------------- block families (A.2)
class DownBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb_channels, num_layers=2, add_downsample=True, eps=1e-5):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps)
            for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, image_only_indicator):
        outs: Tuple[torch.Tensor, ...] = ()
        for r in self.resnets:
            hidden_states = r(hidden_states, temb, image_only_indicator)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class CrossAttnDownBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb_channels, num_layers=2, transformer_layers_per_block=1,
                 num_attention_heads=1, cross_attention_dim=1024, add_downsample=True, eps=1e-6,
                 time_context_order="hw_major_0272"):
        super().__init__()
        self.resnets = nn.ModuleList([
            SpatioTemporalResBlock(in_channels if i == 0 else out_channels, out_channels, temb_channels, eps)
            for i in range(num_layers)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                                           transformer_layers_per_block, cross_attention_dim, time_context_order)
            for _ in range(num_layers)])
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None

    def forward(self, hidden_states, temb, encoder_hidden_states, image_only_indicator):
        outs: Tuple[torch.Tensor, ...] = ()
        for r, a in zip(self.resnets, self.attentions):
            hidden_states = r(hidden_states, temb, image_only_indicator)
            hidden_states = a(hidden_states, encoder_hidden_states, image_only_indicator)
            outs += (hidden_states,)
        if self.downsamplers is not None:
            for d in self.downsamplers:
                hidden_states = d(hidden_states)
            outs += (hidden_states,)
        return hidden_states, outs


class UNetMidBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, temb_channels, num_layers=1, transformer_layers_per_block=1,
                 num_attention_heads=1, cross_attention_dim=1024, eps=1e-5, time_context_order="hw_major_0272"):
        super().__init__()
        self.resnets = nn.ModuleList(
            [SpatioTemporalResBlock(in_channels, in_channels, temb_channels, eps) for _ in range(num_layers + 1)])
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, in_channels // num_attention_heads, in_channels,
                                           transformer_layers_per_block, cross_attention_dim, time_context_order)
            for _ in range(num_layers)])

    def forward(self, hidden_states, temb, encoder_hidden_states, image_only_indicator):
        hidden_states = self.resnets[0](hidden_states, temb, image_only_indicator)
        for a, r in zip(self.attentions, self.resnets[1:]):
            hidden_states = a(hidden_states, encoder_hidden_states, image_only_indicator)
            hidden_states = r(hidden_states, temb, image_only_indicator)
        return hidden_states


def _up_resnets(in_channels, prev_output_channel, out_channels, temb_channels, num_layers, eps):
    blocks = []
    for i in range(num_layers):
        skip = in_channels if i == num_layers - 1 else out_channels
        rin = prev_output_channel if i == 0 else out_channels
        blocks.append(SpatioTemporalResBlock(rin + skip, out_channels, temb_channels, eps))
    return nn.ModuleList(blocks)


class UpBlockSpatioTemporal(nn.Module):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers=3,
                 add_upsample=True, eps=1e-6):
        super().__init__()
        self.resnets = _up_resnets(in_channels, prev_output_channel, out_channels, temb_channels, num_layers, eps)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, image_only_indicator):
        for r in self.resnets:
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, skip], dim=1)   # hidden first
            hidden_states = r(hidden_states, temb, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states


class CrossAttnUpBlockSpatioTemporal(nn.Module):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb_channels, num_layers=3,
                 transformer_layers_per_block=1, num_attention_heads=1, cross_attention_dim=1024,
                 add_upsample=True, eps=1e-6, time_context_order="hw_major_0272"):
        super().__init__()
        self.resnets = _up_resnets(in_channels, prev_output_channel, out_channels, temb_channels, num_layers, eps)
        self.attentions = nn.ModuleList([
            TransformerSpatioTemporalModel(num_attention_heads, out_channels // num_attention_heads, out_channels,
                                           transformer_layers_per_block, cross_attention_dim, time_context_order)
            for _ in range(num_layers)])
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None

    def forward(self, hidden_states, res_hidden_states_tuple, temb, encoder_hidden_states, image_only_indicator):
        for r, a in zip(self.resnets, self.attentions):
            skip = res_hidden_states_tuple[-1]
            res_hidden_states_tuple = res_hidden_states_tuple[:-1]
            hidden_states = torch.cat([hidden_states, skip], dim=1)
            hidden_states = r(hidden_states, temb, image_only_indicator)
            hidden_states = a(hidden_states, encoder_hidden_states, image_only_indicator)
        if self.upsamplers is not None:
            for u in self.upsamplers:
                hidden_states = u(hidden_states)
        return hidden_states
