"""Oracle CLIP vision tower (TEST INFRA ONLY): the ``image_encoder`` of the reference pipelines
(``CLIPVisionModelWithProjection``; call site ``pipeline/pipeline_stable_video_diffusion_controlnet.py:174-214``:
``image_embeddings = self.image_encoder(image).image_embeds``).  The model class lives in the un-vendored
``transformers==4.40.1`` (requirements.txt:71); this is a restatement of its published algorithm under the same parameter
names, PINNED against the ``transformers`` package that happens to be installed here (5.x, same arithmetic):
``tests/test_oracle_golden.py::test_clip_oracle_matches_transformers``.

ViT: 14x14 patch embedding (conv, no bias) + class token + learned position embedding -> pre-LayerNorm -> L x
[LN, multi-head self-attention with biased q/k/v/out projections, residual, LN, fc1-act-fc2, residual] -> post-LayerNorm
of the class token -> bias-free projection."""
from __future__ import annotations

from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


def _act(name: str, x):
    if name == "gelu":
        return F.gelu(x)
    if name == "quick_gelu":
        return x * torch.sigmoid(1.702 * x)
    raise ValueError(name)


class _Attn(nn.Module):
    def __init__(self, c, heads):
        super().__init__()
        self.heads = heads
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (nn.Linear(c, c) for _ in range(4))

    def forward(self, x):
        b, t, c = x.shape
        d = c // self.heads
        q, k, v = (p(x).view(b, t, self.heads, d).transpose(1, 2) for p in (self.q_proj, self.k_proj, self.v_proj))
        w = torch.softmax((q @ k.transpose(-1, -2)) * d ** -0.5, dim=-1)
        return self.out_proj((w @ v).transpose(1, 2).reshape(b, t, c))


class _MLP(nn.Module):
    def __init__(self, c, inner, act):
        super().__init__()
        self.act = act
        self.fc1, self.fc2 = nn.Linear(c, inner), nn.Linear(inner, c)

    def forward(self, x):
        return self.fc2(_act(self.act, self.fc1(x)))


class _Layer(nn.Module):
    def __init__(self, c, heads, inner, act, eps):
        super().__init__()
        self.self_attn = _Attn(c, heads)
        self.layer_norm1 = nn.LayerNorm(c, eps=eps)
        self.mlp = _MLP(c, inner, act)
        self.layer_norm2 = nn.LayerNorm(c, eps=eps)

    def forward(self, x):
        x = x + self.self_attn(self.layer_norm1(x))
        return x + self.mlp(self.layer_norm2(x))


class _Embeddings(nn.Module):
    def __init__(self, c, image_size, patch_size, channels):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(c))
        self.patch_embedding = nn.Conv2d(channels, c, patch_size, stride=patch_size, bias=False)
        self.position_embedding = nn.Embedding((image_size // patch_size) ** 2 + 1, c)

    def forward(self, pixel_values):
        p = self.patch_embedding(pixel_values).flatten(2).transpose(1, 2)
        x = torch.cat([self.class_embedding.expand(p.shape[0], 1, -1), p], dim=1)
        return x + self.position_embedding.weight[None]


class _Encoder(nn.Module):
    def __init__(self, n, *a):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(*a) for _ in range(n)])


class _VisionModel(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        c = cfg.hidden_size
        self.embeddings = _Embeddings(c, cfg.image_size, cfg.patch_size, cfg.num_channels)
        self.pre_layrnorm = nn.LayerNorm(c, eps=cfg.layer_norm_eps)           # (sic) the upstream attribute name
        self.encoder = _Encoder(cfg.num_hidden_layers, c, cfg.num_attention_heads, cfg.intermediate_size, cfg.hidden_act,
                                cfg.layer_norm_eps)
        self.post_layernorm = nn.LayerNorm(c, eps=cfg.layer_norm_eps)


CLIP_VIT_H_14 = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
                     patch_size=14, num_channels=3, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)


class CLIPVisionModelWithProjection(nn.Module):
    """SVD's image encoder is CLIP ViT-H/14 (``CLIP_VIT_H_14``)."""

    def __init__(self, **config):
        super().__init__()
        self.config = SimpleNamespace(**dict(CLIP_VIT_H_14, **config))
        self.vision_model = _VisionModel(self.config)
        self.visual_projection = nn.Linear(self.config.hidden_size, self.config.projection_dim, bias=False)

    def forward(self, pixel_values):
        vm = self.vision_model
        x = vm.pre_layrnorm(vm.embeddings(pixel_values))
        for layer in vm.encoder.layers:
            x = layer(x)
        pooled = vm.post_layernorm(x[:, 0])
        return SimpleNamespace(image_embeds=self.visual_projection(pooled), last_hidden_state=x)
