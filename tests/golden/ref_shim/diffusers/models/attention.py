"""INDEPENDENT of oracle/: constructor-only stand-ins for diffusers' `BasicTransformerBlock` /
`TemporalBasicTransformerBlock` (attributes as patch/patch.py reads them: norm1/attn1/norm2/attn2/norm3/ff,
norm_in/ff_in/is_res, norm_type, pos_embed, only_cross_attention, _chunk_size, _chunk_dim).  They have NO forward:
tests/golden/make_patch_golden.py class-swaps them with the reference's `patch.apply_patch`, so the forward that
runs is the reference's own restatement (patch/patch.py:390-580 and :582-686)."""
import torch.nn as nn
import torch.nn.functional as F

from .attention_processor import Attention


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim_out or dim)])

    def forward(self, x):
        for m in self.net:
            x = m(x)
        return x


class _Attrs:
    norm_type = "layer_norm"
    pos_embed = None
    only_cross_attention = False
    _chunk_size = None
    _chunk_dim = 0

    def forward(self, *a, **k):
        raise RuntimeError("constructor-only stand-in: apply the reference's patch.apply_patch first")


class BasicTransformerBlock(_Attrs, nn.Module):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(dim)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(dim)
        self.ff = FeedForward(dim)


class TemporalBasicTransformerBlock(_Attrs, nn.Module):
    def __init__(self, dim, time_mix_inner_dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self.norm_in = nn.LayerNorm(dim)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim)
        self.norm1 = nn.LayerNorm(time_mix_inner_dim)
        self.attn1 = Attention(time_mix_inner_dim, heads, dim_head)
        self.norm2 = nn.LayerNorm(time_mix_inner_dim)
        self.attn2 = Attention(time_mix_inner_dim, heads, dim_head, cross_attention_dim)
        self.norm3 = nn.LayerNorm(time_mix_inner_dim)
        self.ff = FeedForward(time_mix_inner_dim)
