"""CLIP image encoder of the SVD pipelines (first half) on the lkgd_b200 kernels.

Reference call site: ``pipeline/pipeline_stable_video_diffusion_controlnet.py:174-214`` (``_encode_image``:
``self.image_encoder(image).image_embeds`` -> ``[B, 1, 1024]`` -> zero unconditional half prepended); the module is
``transformers.CLIPVisionModelWithProjection`` (ViT-H/14 for SVD), loaded by ``from_pretrained(..., subfolder="image_encoder")``.
Same parameter names as the transformers class, so its checkpoints load with ``load_state_dict``.

Execution: patch unfold (1 kernel) -> patch-embedding GEMM with the position embedding as a per-row vector -> fp32 token
stream [N*T, C]; per layer LayerNorm -> fused biased q|k|v GEMM -> flash attention (d = 80 for ViT-H: the two-sub-tile
kernel, TMA zero-fills channels 80..127) -> out-projection GEMM with the residual in its epilogue -> LayerNorm -> fc1 GEMM with
GELU / quick-GELU in the epilogue -> fc2 GEMM with the residual; post-LayerNorm of the class token, projection GEMM."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import Optional

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from .ops import ACT_GELU, ACT_QUICK_GELU, RV_FRAMEPOS, bf16

CLIP_VIT_H_14 = dict(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16, image_size=224,
                     patch_size=14, num_channels=3, projection_dim=1024, hidden_act="gelu", layer_norm_eps=1e-5)


class _Attn(M.Container):
    def __init__(self, c):
        super().__init__()
        self.k_proj, self.v_proj, self.q_proj, self.out_proj = (M.Linear(c, c) for _ in range(4))


class _MLP(M.Container):
    def __init__(self, c, inner):
        super().__init__()
        self.fc1, self.fc2 = M.Linear(c, inner), M.Linear(inner, c)


class _Layer(M.Container):
    def __init__(self, c, inner, eps):
        super().__init__()
        self.self_attn = _Attn(c)
        self.layer_norm1 = M.LayerNorm(c, eps=eps)
        self.mlp = _MLP(c, inner)
        self.layer_norm2 = M.LayerNorm(c, eps=eps)


class _Embeddings(M.Container):
    def __init__(self, c, image_size, patch_size, channels):
        super().__init__()
        self.class_embedding = nn.Parameter(torch.randn(c))
        self.patch_embedding = M.Conv2d(channels, c, patch_size, stride=patch_size, bias=False)
        self.position_embedding = nn.Embedding((image_size // patch_size) ** 2 + 1, c)


class _Encoder(M.Container):
    def __init__(self, n, c, inner, eps):
        super().__init__()
        self.layers = nn.ModuleList([_Layer(c, inner, eps) for _ in range(n)])


class _VisionModel(M.Container):
    def __init__(self, cfg):
        super().__init__()
        c = cfg.hidden_size
        self.embeddings = _Embeddings(c, cfg.image_size, cfg.patch_size, cfg.num_channels)
        self.pre_layrnorm = M.LayerNorm(c, eps=cfg.layer_norm_eps)             # (sic) upstream attribute name
        self.encoder = _Encoder(cfg.num_hidden_layers, c, cfg.intermediate_size, cfg.layer_norm_eps)
        self.post_layernorm = M.LayerNorm(c, eps=cfg.layer_norm_eps)


def _f32(t):
    return t.detach().to(torch.float32).contiguous()


class CLIPVisionModelWithProjection(nn.Module):
    """Drop-in for ``transformers.CLIPVisionModelWithProjection`` as the SVD pipelines use it (``.image_embeds``)."""

    def __init__(self, **config):
        super().__init__()
        cfg = dict(CLIP_VIT_H_14, **{k: v for k, v in config.items() if k in CLIP_VIT_H_14})
        self.config = SimpleNamespace(**cfg)
        c = self.config
        if c.hidden_act not in ("gelu", "quick_gelu"):
            raise ValueError(f"hidden_act {c.hidden_act!r} is not supported (gelu, quick_gelu)")
        if c.hidden_size % c.num_attention_heads or (c.hidden_size // c.num_attention_heads) % 8 \
                or c.hidden_size // c.num_attention_heads > 128:
            raise ValueError("head width must be a multiple of 8 and at most 128")
        if c.image_size % c.patch_size:
            raise ValueError("image_size must be a multiple of patch_size")
        self.vision_model = _VisionModel(c)
        self.visual_projection = M.Linear(c.hidden_size, c.projection_dim, bias=False)
        self._pk = None

    # ---- plumbing shared with the UNet modules
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def invalidate(self):
        self._pk = None

    def load_state_dict(self, sd, *a, **k):
        sd = {key: v for key, v in sd.items() if not key.endswith("position_ids")}     # a buffer of older checkpoints
        out = super().load_state_dict(sd, *a, **k)
        self.invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pk = None
        return out

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, device=None):
        """Local directory with ``config.json`` + ``model.safetensors`` / ``pytorch_model.bin`` (the ``image_encoder``
        sub-folder of an SVD checkpoint)."""
        root = os.path.join(path, subfolder) if subfolder else path
        cfg_path = os.path.join(root, "config.json")
        if not os.path.isfile(cfg_path):
            raise EnvironmentError(f"{cfg_path} not found: lkgd_b200 loads local model directories only")
        cfg = json.load(open(cfg_path))
        cfg = cfg.get("vision_config", cfg)
        net = cls(**cfg)
        for name in ("model.safetensors", "model.fp16.safetensors", "pytorch_model.bin"):
            f = os.path.join(root, name)
            if os.path.isfile(f):
                if f.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(f)
                else:
                    sd = torch.load(f, map_location="cpu", weights_only=True)
                net.load_state_dict(sd, strict=True)
                return net.to(device) if device is not None else net
        raise EnvironmentError(f"no model.safetensors / pytorch_model.bin under {root}")

    # ---- kernel-ready weights
    def _pack(self):
        if self._pk is None:
            if self.device.type != "cuda":
                raise RuntimeError("lkgd_b200 runs on CUDA devices only (there is no CPU or PyTorch fallback)")
            c = self.config
            vm = self.vision_model
            K = c.num_channels * c.patch_size ** 2
            Kpad = (K + 7) // 8 * 8
            wp = torch.zeros(c.hidden_size, Kpad, device=self.device)
            wp[:, :K] = vm.embeddings.patch_embedding.weight.detach().float().reshape(c.hidden_size, K)
            pos = _f32(vm.embeddings.position_embedding.weight)
            layers = []
            for L in vm.encoder.layers:
                a = L.self_attn
                layers.append(SimpleNamespace(
                    ln1=(_f32(L.layer_norm1.weight), _f32(L.layer_norm1.bias)),
                    ln2=(_f32(L.layer_norm2.weight), _f32(L.layer_norm2.bias)),
                    wqkv=torch.cat([a.q_proj.weight, a.k_proj.weight, a.v_proj.weight], 0).detach().to(bf16).contiguous(),
                    bqkv=torch.cat([a.q_proj.bias, a.k_proj.bias, a.v_proj.bias], 0).detach().float().contiguous(),
                    wo=a.out_proj.weight.detach().to(bf16).contiguous(), bo=_f32(a.out_proj.bias),
                    w1=L.mlp.fc1.weight.detach().to(bf16).contiguous(), b1=_f32(L.mlp.fc1.bias),
                    w2=L.mlp.fc2.weight.detach().to(bf16).contiguous(), b2=_f32(L.mlp.fc2.bias)))
            self._pk = SimpleNamespace(
                Kpad=Kpad, wp=wp.to(bf16).contiguous(), pos_patches=pos[1:].contiguous(),
                cls_row=(_f32(vm.embeddings.class_embedding) + pos[0]).contiguous(),
                pre=(_f32(vm.pre_layrnorm.weight), _f32(vm.pre_layrnorm.bias)),
                post=(_f32(vm.post_layernorm.weight), _f32(vm.post_layernorm.bias)),
                proj=self.visual_projection.weight.detach().to(bf16).contiguous(), layers=layers)
        return self._pk

    @ops.on_own_device
    @torch.no_grad()
    def forward(self, pixel_values: torch.Tensor, **unused):
        c = self.config
        pk = self._pack()
        if pixel_values.ndim != 4 or pixel_values.shape[1] != c.num_channels or pixel_values.shape[2] != c.image_size \
                or pixel_values.shape[3] != c.image_size:
            raise ValueError(f"Input image size ({pixel_values.shape[2]}*{pixel_values.shape[3]}) doesn't match model "
                             f"({c.image_size}*{c.image_size}).")
        dev = self.device
        N = pixel_values.shape[0]
        C, heads = c.hidden_size, c.num_attention_heads
        d = C // heads
        P = (c.image_size // c.patch_size) ** 2
        T = P + 1
        eps = c.layer_norm_eps
        act = ACT_GELU if c.hidden_act == "gelu" else ACT_QUICK_GELU
        x = ops.patchify(pixel_values.to(dev), c.patch_size, pk.Kpad)
        emb = torch.empty((N * T, C), device=dev, dtype=torch.float32)
        for n in range(N):      # patches + their position embeddings land behind each sample's class row
            ops.gemm(x[n * P:(n + 1) * P], pk.wp, rowvec=pk.pos_patches, rv=(RV_FRAMEPOS, 1, P, 1),
                     out=emb[n * T + 1:(n + 1) * T], out_f32=True)
        emb.view(N, T, C)[:, 0].copy_(pk.cls_row)                       # class token + position 0 (a constant row)
        # pre-LayerNorm starts the residual stream: bf16 result widened once into the fp32 stream
        h = torch.zeros((N * T, C), device=dev, dtype=torch.float32)
        ops.axpby(ops.layernorm(emb, pk.pre[0], pk.pre[1], eps), 1.0, h, 0.0)
        for L in pk.layers:
            n1 = ops.layernorm(h, L.ln1[0], L.ln1[1], eps)
            qkv = ops.gemm(n1, L.wqkv, bias=L.bqkv)
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_img=N, heads=heads, d=d, Nq=T, Nk=T)
            h = ops.gemm(a, L.wo, bias=L.bo, res1=h, out_f32=True)
            n2 = ops.layernorm(h, L.ln2[0], L.ln2[1], eps)
            f = ops.gemm(n2, L.w1, bias=L.b1, act=act)
            h = ops.gemm(f, L.w2, bias=L.b2, res1=h, out_f32=True)
        pooled = h.view(N, T, C)[:, 0].contiguous()
        pn = ops.layernorm(pooled, pk.post[0], pk.post[1], eps)
        embeds = ops.gemm(pn, pk.proj, out_f32=True)
        return SimpleNamespace(image_embeds=embeds.to(pixel_values.dtype if pixel_values.is_floating_point() else torch.float32),
                               last_hidden_state=h.view(N, T, C))
