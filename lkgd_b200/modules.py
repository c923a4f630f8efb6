"""Parameter containers that reproduce the reference's module tree (names == the reference's
``named_parameters()`` dumps, train_svd_lora.txt / train_svd_lora_train.txt) WITHOUT any arithmetic: every
``forward`` here raises.  The compute lives in ``engine.py`` (CUDA kernels through the C ABI).

Structure follows reference ``models/unet_spatio_temporal_condition_controlnet.py:126-245`` and the
diffusers==0.27.2 block factories it calls; class names
``BasicTransformerBlock`` / ``TemporalBasicTransformerBlock`` are kept because the reference's patching code
matches blocks by class name (patch/patch.py:791-806)."""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn

_MSG = "lkgd_b200 modules are parameter containers; the forward pass runs in the CUDA engine (no PyTorch fallback)"


class _NoForward:
    def forward(self, *a, **k):  # pragma: no cover - guard
        raise RuntimeError(_MSG)


class Linear(_NoForward, nn.Linear):
    pass


class Conv2d(_NoForward, nn.Conv2d):
    pass


class Conv3d(_NoForward, nn.Conv3d):
    pass


class Conv1d(_NoForward, nn.Conv1d):
    pass


class GroupNorm(_NoForward, nn.GroupNorm):
    pass


class LayerNorm(_NoForward, nn.LayerNorm):
    pass


class Container(_NoForward, nn.Module):
    pass


class TimestepEmbedding(Container):
    def __init__(self, in_channels, time_embed_dim, out_dim=None):
        super().__init__()
        self.linear_1 = Linear(in_channels, time_embed_dim)
        self.linear_2 = Linear(time_embed_dim, out_dim or time_embed_dim)


class Attention(Container):
    def __init__(self, query_dim, heads, dim_head, cross_attention_dim=None):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head = heads, dim_head
        kv = cross_attention_dim if cross_attention_dim is not None else query_dim
        self.to_q = Linear(query_dim, inner, bias=False)
        self.to_k = Linear(kv, inner, bias=False)
        self.to_v = Linear(kv, inner, bias=False)
        self.to_out = nn.ModuleList([Linear(inner, query_dim), nn.Dropout(0.0)])


class GEGLU(Container):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = Linear(dim_in, dim_out * 2)


class FeedForward(Container):
    def __init__(self, dim, dim_out=None, mult=4):
        super().__init__()
        self.net = nn.ModuleList([GEGLU(dim, dim * mult), nn.Dropout(0.0), Linear(dim * mult, dim_out or dim)])


class _JointAttention:
    """State of the reference's joint-attention patch (``patch/patch.py`` ToMeBlock) on a transformer
    block: a second attention ``attn1n`` whose keys / values come from the PARTNER sample of the batch, a zero-initialised
    post layer (``conv1n`` = bias-free Linear, or ``scale1n``) and the switches the reference keeps on the patched block.
    Parameter names equal the reference's (``...transformer_blocks.0.attn1n.to_q.weight``, ``...conv1n.weight``)."""
    enable_joint_attention = False          # patch.py:104 is True by default once PATCHED; un-patched blocks have none
    joint_scale = 1.0
    patched = False                         # set by lkgd_b200.patch.apply_patch
    flip = False                            # _tome_info["args"]["flip"] (spatial block only, :458-464)
    joint_attn_mask = None
    post = None
    add_norm = False

    def initialize_joint_layers(self, post: str = "conv", add_norm: bool = False):
        """``ToMeBlock.initialize_joint_layers`` (patch/patch.py:143-172)."""
        import copy
        if add_norm:
            raise NotImplementedError("add_norm=True (AdaLayerNormContinuous on the joint branch) is not built")
        if post not in ("conv", "scale", "conv_fuse"):
            raise ValueError(f"Unkown post processing type {post}")            # (sic) the reference's assert message
        self.attn1n = copy.deepcopy(self.attn1)
        c = self.attn1.to_out[0].out_features
        dev, dt = self.attn1.to_out[0].weight.device, self.attn1.to_out[0].weight.dtype
        if post == "scale":
            self.scale1n = nn.Parameter(torch.zeros(1, 1, c, device=dev, dtype=dt))
        else:
            k = 2 * c if post == "conv_fuse" else c         # conv_fuse: ONE layer over [masked sample | partner] (:154-157)
            self.conv1n = Linear(k, k, bias=False, device=dev, dtype=dt)
            if dev.type != "meta":
                nn.init.zeros_(self.conv1n.weight)
        self.post, self.add_norm, self.joint_scale = post, add_norm, 1.0

    @property
    def post_joint(self):
        return self.scale1n if self.post == "scale" else self.conv1n

    def set_joint_attention(self, enable: bool = True):
        self.enable_joint_attention = enable

    def set_joint_scale(self, scale: float = 1.0):
        self.joint_scale = scale

    def joint_active(self) -> bool:
        return bool(self.patched and self.enable_joint_attention)


class BasicTransformerBlock(_JointAttention, Container):
    def __init__(self, dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.norm_type, self.only_cross_attention, self.pos_embed = "layer_norm", False, None
        self._chunk_size, self._chunk_dim = None, 0
        self.norm1 = LayerNorm(dim, eps=1e-5)
        self.attn1 = Attention(dim, heads, dim_head)
        self.norm2 = LayerNorm(dim, eps=1e-5)
        self.attn2 = Attention(dim, heads, dim_head, cross_attention_dim)
        self.norm3 = LayerNorm(dim, eps=1e-5)
        self.ff = FeedForward(dim)


class TemporalBasicTransformerBlock(_JointAttention, Container):
    def __init__(self, dim, time_mix_inner_dim, heads, dim_head, cross_attention_dim):
        super().__init__()
        self.is_res = dim == time_mix_inner_dim
        self._chunk_size, self._chunk_dim = None, 0
        self.norm_in = LayerNorm(dim, eps=1e-5)
        self.ff_in = FeedForward(dim, dim_out=time_mix_inner_dim)
        self.norm1 = LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.attn1 = Attention(time_mix_inner_dim, heads, dim_head)
        self.norm2 = LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.attn2 = Attention(time_mix_inner_dim, heads, dim_head, cross_attention_dim)
        self.norm3 = LayerNorm(time_mix_inner_dim, eps=1e-5)
        self.ff = FeedForward(time_mix_inner_dim)


class AlphaBlender(Container):
    """``switch_spatial_to_temporal_mix`` (the VAE's temporal decoder): alpha -> 1 - alpha."""

    def __init__(self, alpha=0.5, merge_strategy="learned_with_images", switch_spatial_to_temporal_mix=False):
        super().__init__()
        self.merge_strategy, self.switch_spatial_to_temporal_mix = merge_strategy, switch_spatial_to_temporal_mix
        self.mix_factor = nn.Parameter(torch.tensor([alpha], dtype=torch.float32))


class TransformerSpatioTemporalModel(Container):
    def __init__(self, heads, dim_head, in_channels, num_layers=1, cross_attention_dim=1024):
        super().__init__()
        if num_layers != 1:
            raise ValueError("lkgd_b200 supports transformer_layers_per_block == 1 (the SVD configuration)")
        inner = heads * dim_head
        self.heads, self.dim_head, self.in_channels = heads, dim_head, in_channels
        self.norm = GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(inner, heads, dim_head, cross_attention_dim)])
        self.temporal_transformer_blocks = nn.ModuleList(
            [TemporalBasicTransformerBlock(inner, inner, heads, dim_head, cross_attention_dim)])
        self.time_pos_embed = TimestepEmbedding(in_channels, in_channels * 4, out_dim=in_channels)
        self.time_mixer = AlphaBlender(0.5)
        self.proj_out = Linear(inner, in_channels)


class ResnetBlock2D(Container):
    def __init__(self, in_channels, out_channels, temb_channels, eps):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = GroupNorm(32, in_channels, eps=eps)
        self.conv1 = Conv2d(in_channels, out_channels, 3, padding=1)
        self.time_emb_proj = Linear(temb_channels, out_channels) if temb_channels is not None else None    # None: the VAE
        self.norm2 = GroupNorm(32, out_channels, eps=eps)
        self.conv2 = Conv2d(out_channels, out_channels, 3, padding=1)
        self.conv_shortcut = Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else None


class TemporalResnetBlock(Container):
    def __init__(self, in_channels, out_channels, temb_channels, eps):
        super().__init__()
        if in_channels != out_channels:
            raise ValueError("temporal resblocks of the SVD UNet have in_channels == out_channels")
        self.norm1 = GroupNorm(32, in_channels, eps=eps)
        self.conv1 = Conv3d(in_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.time_emb_proj = Linear(temb_channels, out_channels) if temb_channels is not None else None
        self.norm2 = GroupNorm(32, out_channels, eps=eps)
        self.conv2 = Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))
        self.conv_shortcut = None


class SpatioTemporalResBlock(Container):
    """UNet: ``(in, out, temb_channels, eps)``; VAE temporal decoder: ``temb_channels=None, eps=1e-6, temporal_eps=1e-5,
    merge_factor=0.0, merge_strategy="learned", switch_spatial_to_temporal_mix=True``."""

    def __init__(self, in_channels, out_channels, temb_channels, eps, temporal_eps=None, merge_factor=0.5,
                 merge_strategy="learned_with_images", switch_spatial_to_temporal_mix=False):
        super().__init__()
        self.spatial_res_block = ResnetBlock2D(in_channels, out_channels, temb_channels, eps)
        self.temporal_res_block = TemporalResnetBlock(out_channels, out_channels, temb_channels,
                                                      eps if temporal_eps is None else temporal_eps)
        self.time_mixer = AlphaBlender(merge_factor, merge_strategy, switch_spatial_to_temporal_mix)


class Downsample2D(Container):
    """``padding=0`` (the VAE encoder): the input is padded on the bottom / right only, F.pad(x, (0, 1, 0, 1))."""

    def __init__(self, channels, padding=1):
        super().__init__()
        self.padding = padding
        self.conv = Conv2d(channels, channels, 3, stride=2, padding=padding)


class Upsample2D(Container):
    def __init__(self, channels):
        super().__init__()
        self.conv = Conv2d(channels, channels, 3, padding=1)


def _resnets(chans, temb, eps):
    return nn.ModuleList([SpatioTemporalResBlock(i, o, temb, eps) for i, o in chans])


def _attns(n, heads, channels, layers, xdim):
    return nn.ModuleList([TransformerSpatioTemporalModel(heads, channels // heads, channels, layers, xdim)
                          for _ in range(n)])


class DownBlockSpatioTemporal(Container):
    has_cross_attention = False

    def __init__(self, in_channels, out_channels, temb, num_layers=2, add_downsample=True, eps=1e-5):
        super().__init__()
        self.resnets = _resnets([(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)],
                                temb, eps)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class CrossAttnDownBlockSpatioTemporal(Container):
    has_cross_attention = True

    def __init__(self, in_channels, out_channels, temb, num_layers=2, tlayers=1, heads=1, xdim=1024,
                 add_downsample=True, eps=1e-6):
        super().__init__()
        self.attentions = _attns(num_layers, heads, out_channels, tlayers, xdim)   # diffusers registers these first
        self.resnets = _resnets([(in_channels if i == 0 else out_channels, out_channels) for i in range(num_layers)],
                                temb, eps)
        self.downsamplers = nn.ModuleList([Downsample2D(out_channels)]) if add_downsample else None


class UNetMidBlockSpatioTemporal(Container):
    has_cross_attention = True

    def __init__(self, in_channels, temb, num_layers=1, tlayers=1, heads=1, xdim=1024, eps=1e-5):
        super().__init__()
        self.attentions = _attns(num_layers, heads, in_channels, tlayers, xdim)
        self.resnets = _resnets([(in_channels, in_channels)] * (num_layers + 1), temb, eps)


def _up_chans(in_channels, prev_output_channel, out_channels, num_layers):
    return [((prev_output_channel if i == 0 else out_channels) +
             (in_channels if i == num_layers - 1 else out_channels), out_channels) for i in range(num_layers)]


class UpBlockSpatioTemporal(Container):
    has_cross_attention = False

    def __init__(self, in_channels, prev_output_channel, out_channels, temb, num_layers=3, add_upsample=True,
                 eps=1e-6):
        super().__init__()
        self.resnets = _resnets(_up_chans(in_channels, prev_output_channel, out_channels, num_layers), temb, eps)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class CrossAttnUpBlockSpatioTemporal(Container):
    has_cross_attention = True

    def __init__(self, in_channels, prev_output_channel, out_channels, temb, num_layers=3, tlayers=1, heads=1,
                 xdim=1024, add_upsample=True, eps=1e-6):
        super().__init__()
        self.attentions = _attns(num_layers, heads, out_channels, tlayers, xdim)
        self.resnets = _resnets(_up_chans(in_channels, prev_output_channel, out_channels, num_layers), temb, eps)
        self.upsamplers = nn.ModuleList([Upsample2D(out_channels)]) if add_upsample else None


class QuaternionLinear(Container):
    """Parameter layout of core_qnn ``QuaternionLinearAutograd`` (r/i/j/k_weight [in/4, out/4], bias [out])."""

    def __init__(self, in_features, out_features):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        i4, o4 = in_features // 4, out_features // 4
        bound = (1.0 / (2.0 * (i4 + o4))) ** 0.5
        for n in ("r_weight", "i_weight", "j_weight", "k_weight"):
            setattr(self, n, nn.Parameter(torch.empty(i4, o4).uniform_(-bound, bound)))
        self.bias = nn.Parameter(torch.zeros(out_features))


class LoraLinear(Container):
    """peft-0.10-shaped LoRA wrapper (names ``base_layer``, ``lora_A.<adapter>``, ``lora_B.<adapter>``);
    init as reference ``models/lora_layer.py:132-150``: A ~ N(0, 1/r) ("gaussian") or kaiming-uniform, B = 0."""

    def __init__(self, base_layer: Linear, r: int, lora_alpha: float, init_lora_weights="gaussian",
                 adapter_name="default", use_rslora=False):
        super().__init__()
        if r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        self.base_layer = base_layer
        self.adapter_name, self.r = adapter_name, r
        self.scaling = lora_alpha / math.sqrt(r) if use_rslora else lora_alpha / r
        dev, dt = base_layer.weight.device, base_layer.weight.dtype
        self.lora_A = nn.ModuleDict({adapter_name: Linear(base_layer.in_features, r, bias=False, device=dev, dtype=dt)})
        self.lora_B = nn.ModuleDict({adapter_name: Linear(r, base_layer.out_features, bias=False, device=dev, dtype=dt)})
        self.merged = False
        a, b = self.lora_A[adapter_name].weight, self.lora_B[adapter_name].weight
        if a.device.type != "meta":
            if init_lora_weights is True:
                nn.init.kaiming_uniform_(a, a=math.sqrt(5))
            elif isinstance(init_lora_weights, str) and init_lora_weights.lower() == "gaussian":
                nn.init.normal_(a, std=1 / r)
            elif init_lora_weights is not False:
                raise ValueError(f"Unknown initialization {init_lora_weights=}")
            if init_lora_weights is not False:
                nn.init.zeros_(b)
        self.scalings = {adapter_name: self.scaling}
        self.ranks = {adapter_name: r}
        self.active_adapters = [adapter_name]      # peft `set_adapters`: every listed adapter adds up
        self.lora_mask = {}                        # adapter -> bool mask over the batch (patch.set_patch_lora_mask)
        self.masked_forward = False                # patch.hack_lora_forward: per-sample masked adapters (patch.py:57-92)

    def update_layer(self, adapter_name: str, r: int, lora_alpha: float, init_lora_weights="gaussian"):
        """A further adapter on the same layer (reference models/lora_layer.py:85-130)."""
        if adapter_name in self.lora_A:
            raise ValueError(f"adapter {adapter_name!r} already exists on this layer")
        if r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        base = self.base_layer
        dev, dt = base.weight.device, base.weight.dtype
        self.lora_A[adapter_name] = Linear(base.in_features, r, bias=False, device=dev, dtype=dt)
        self.lora_B[adapter_name] = Linear(r, base.out_features, bias=False, device=dev, dtype=dt)
        a, b = self.lora_A[adapter_name].weight, self.lora_B[adapter_name].weight
        if a.device.type != "meta":
            if init_lora_weights is True:
                nn.init.kaiming_uniform_(a, a=math.sqrt(5))
            elif isinstance(init_lora_weights, str) and init_lora_weights.lower() == "gaussian":
                nn.init.normal_(a, std=1 / r)
            if init_lora_weights is not False:
                nn.init.zeros_(b)
        self.scalings[adapter_name] = lora_alpha / r
        self.ranks[adapter_name] = r
        self.active_adapters.append(adapter_name)

    def adapters(self):
        """[(name, A weight, B weight, scaling, mask | None)] of the active adapters; the mask is None unless the masked
        forward is switched on and the adapter's mask leaves some sample out."""
        out = []
        for name in self.active_adapters:
            if name not in self.lora_A:
                continue
            mask = None
            if self.masked_forward:
                if name not in self.lora_mask:
                    raise KeyError(f"masked LoRA forward without a mask for adapter {name!r} (patch.set_patch_lora_mask)")
                mask = self.lora_mask[name]
                if bool(mask.all()):
                    mask = None
            out.append((name, self.lora_A[name].weight, self.lora_B[name].weight, self.scalings[name], mask))
        return out

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    @property
    def weight(self):
        return self.base_layer.weight

    @property
    def bias(self):
        return self.base_layer.bias
