"""Oracle UNets (TEST INFRA ONLY): the ControlNet-accepting SVD UNet and the LKGD UNet.

Follows the reference wiring at
  ``models/unet_spatio_temporal_condition_controlnet.py:126-245`` (constructor) and ``:358-508`` (forward),
  ``models/unet_spatio_temporal_condition.py:197-225`` (latent-knowledge modules) and ``:448-693`` (forward).
Block arithmetic comes from ``oracle/blocks.py`` (diffusers 0.27.2 restatement).
"""
from __future__ import annotations

from types import SimpleNamespace
from typing import Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn
import torch.nn.functional as F

from .blocks import (CrossAttnDownBlockSpatioTemporal, CrossAttnUpBlockSpatioTemporal, DownBlockSpatioTemporal,
                     TimestepEmbedding, UNetMidBlockSpatioTemporal, UpBlockSpatioTemporal, timestep_embedding)

# released SVD / SVD-XT unet/config.json as recalled (U4: heads [5,10,20,20]); num_frames 25 (XT) / 14.
SVD_XT_CONFIG = dict(
    sample_size=96, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
    up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3,
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20), num_frames=25,
)

# BASELINE.json configs[0]: reduced random-init config, CPU-runnable.
REDUCED_CONFIG = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32,
    projection_class_embeddings_input_dim=96, layers_per_block=2, cross_attention_dim=32,
    transformer_layers_per_block=1, num_attention_heads=(2, 4), num_frames=8,
)


def _tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class UNetSpatioTemporalConditionControlNetModel(nn.Module):
    """Reference ``UNetSpatioTemporalConditionControlNetModel``
    (``models/unet_spatio_temporal_condition_controlnet.py:69-245,358-508``)."""

    def __init__(self, sample_size=None, in_channels=8, out_channels=4,
                 down_block_types=SVD_XT_CONFIG["down_block_types"], up_block_types=SVD_XT_CONFIG["up_block_types"],
                 block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
                 projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
                 transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20), num_frames=25,
                 time_context_order="hw_major_0272"):
        super().__init__()
        self._init_root()    # LKGD: learnable texts* are direct root parameters (first in the reference's dump)
        n = len(down_block_types)
        if len(up_block_types) != n:       # reference :101-104
            raise ValueError("Must provide the same number of `down_block_types` as `up_block_types`.")
        if len(block_out_channels) != n:   # reference :106-109
            raise ValueError("Must provide the same number of `block_out_channels` as `down_block_types`.")
        if not isinstance(num_attention_heads, int) and len(num_attention_heads) != n:  # :111-114
            raise ValueError("Must provide the same number of `num_attention_heads` as `down_block_types`.")
        if isinstance(cross_attention_dim, (list, tuple)) and len(cross_attention_dim) != n:  # :116-119
            raise ValueError("Must provide the same number of `cross_attention_dim` as `down_block_types`.")
        if not isinstance(layers_per_block, int) and len(layers_per_block) != n:  # :121-124
            raise ValueError("Must provide the same number of `layers_per_block` as `down_block_types`.")
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
            block_out_channels=tuple(block_out_channels), addition_time_embed_dim=addition_time_embed_dim,
            projection_class_embeddings_input_dim=projection_class_embeddings_input_dim,
            layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
            transformer_layers_per_block=transformer_layers_per_block, num_attention_heads=num_attention_heads,
            num_frames=num_frames, time_context_order=time_context_order)
        heads = _tuple(num_attention_heads, n)
        xdim = _tuple(cross_attention_dim, n)
        lpb = _tuple(layers_per_block, n)
        tlpb = _tuple(transformer_layers_per_block, n)
        c0 = block_out_channels[0]
        temb = c0 * 4

        self.conv_in = nn.Conv2d(in_channels, c0, 3, padding=1)
        self.time_embedding = TimestepEmbedding(c0, temb)
        self.add_embedding = TimestepEmbedding(projection_class_embeddings_input_dim, temb)

        self.down_blocks = nn.ModuleList()
        self.up_blocks = nn.ModuleList()
        out_c = c0
        for i, t in enumerate(down_block_types):
            in_c, out_c = out_c, block_out_channels[i]
            last = i == n - 1
            if t == "CrossAttnDownBlockSpatioTemporal":
                blk = CrossAttnDownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], tlpb[i], heads[i], xdim[i],
                                                       add_downsample=not last,
                                                       time_context_order=time_context_order)
            elif t == "DownBlockSpatioTemporal":
                blk = DownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], add_downsample=not last)
            else:
                raise ValueError(f"{t} does not exist.")
            self.down_blocks.append(blk)
        self._init_extra()   # LKGD registers its latent-knowledge modules here (dump order: after up_blocks)
        self.mid_block = UNetMidBlockSpatioTemporal(block_out_channels[-1], temb, 1, tlpb[-1], heads[-1], xdim[-1],
                                                    time_context_order=time_context_order)
        rc, rh, rl, rx, rt = (list(reversed(v)) for v in (block_out_channels, heads, lpb, xdim, tlpb))
        out_c = rc[0]
        for i, t in enumerate(up_block_types):
            last = i == n - 1
            prev_c, out_c = out_c, rc[i]
            in_c = rc[min(i + 1, n - 1)]
            if t == "CrossAttnUpBlockSpatioTemporal":
                blk = CrossAttnUpBlockSpatioTemporal(in_c, prev_c, out_c, temb, rl[i] + 1, rt[i], rh[i], rx[i],
                                                     add_upsample=not last, time_context_order=time_context_order)
            elif t == "UpBlockSpatioTemporal":
                blk = UpBlockSpatioTemporal(in_c, prev_c, out_c, temb, rl[i] + 1, add_upsample=not last)
            else:
                raise ValueError(f"{t} does not exist.")
            self.up_blocks.append(blk)
        self.conv_norm_out = nn.GroupNorm(32, c0, eps=1e-5)
        self.conv_out = nn.Conv2d(c0, out_channels, 3, padding=1)

    def _init_root(self):
        pass

    def _init_extra(self):
        pass

    # ------------------------------------------------------------------ forward pieces
    def _time_embedding(self, sample, timestep, added_time_ids):
        """reference ``...controlnet.py:389-426``."""
        t = timestep
        if not torch.is_tensor(t):
            t = torch.tensor([t], dtype=torch.float64 if isinstance(timestep, float) else torch.int64,
                             device=sample.device)
        elif t.ndim == 0:
            t = t[None].to(sample.device)
        b, f = sample.shape[:2]
        t = t.expand(b)
        c0 = self.config.block_out_channels[0]
        t_emb = timestep_embedding(t, c0).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        te = timestep_embedding(added_time_ids.flatten(), self.config.addition_time_embed_dim)
        te = te.reshape(b, -1).to(emb.dtype)
        emb = emb + self.add_embedding(te)
        return emb.repeat_interleave(f, dim=0)

    def _condition(self, encoder_hidden_states, *extra):
        return encoder_hidden_states

    def _stem(self, sample):
        return self.conv_in(sample)

    def _body(self, sample, emb, ctx, b, f, down_block_additional_residuals, mid_block_additional_residual):
        sample = self._stem(sample)
        iof = torch.zeros(b, f, dtype=sample.dtype, device=sample.device)
        skips: Tuple[torch.Tensor, ...] = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(sample, emb, ctx, iof)
            else:
                sample, res = blk(sample, emb, iof)
            skips += res
            # reference quirk F6: the add sits INSIDE the block loop and zip truncates
            # (``...controlnet.py:453-462``) -> multipliers (4,4,4,4,3,3,3,2,2,2,1,1).
            if down_block_additional_residuals is not None:
                skips = tuple(s + r for s, r in zip(skips, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, ctx, iof)
        if mid_block_additional_residual is not None:
            sample = sample + mid_block_additional_residual
        for blk in self.up_blocks:
            k = len(blk.resnets)
            res, skips = skips[-k:], skips[:-k]
            if blk.has_cross_attention:
                sample = blk(sample, res, emb, ctx, iof)
            else:
                sample = blk(sample, res, emb, iof)
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        return sample.reshape(b, f, *sample.shape[1:])

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, added_time_ids=None):
        b, f = sample.shape[:2]
        emb = self._time_embedding(sample, timestep, added_time_ids)
        ctx = encoder_hidden_states.repeat_interleave(f, dim=0)
        out = self._body(sample.flatten(0, 1), emb, ctx, b, f, down_block_additional_residuals,
                         mid_block_additional_residual)
        return SimpleNamespace(sample=out) if return_dict else (out,)


# --------------------------------------------------------------------------- flow stem
class UNetSpatioTemporalConditionModelFlow(UNetSpatioTemporalConditionControlNetModel):
    """The ControlNet-accepting UNet with the second, gated input stem of the reference's flow pipelines
    (``models/unet_spatio_temporal_condition_flow.py:260-273,494-502``): the sample carries THREE 4-channel groups
    (noise | condition | second condition); ``conv_in`` sees (noise, condition), ``conv_in2`` - created by
    ``initialize_conv_in()`` as a copy of ``conv_in`` - sees (noise, second condition) and is scaled per output channel
    by ``conv_in2_alpha`` (zero-initialised, so a fresh stem leaves the model unchanged)."""

    def initialize_conv_in(self):
        c = self.conv_in
        self.conv_in2 = nn.Conv2d(c.in_channels, c.out_channels, 3, padding=1).to(c.weight.device, c.weight.dtype)
        self.conv_in2_alpha = nn.Parameter(torch.zeros(1, c.out_channels, 1, 1, device=c.weight.device,
                                                       dtype=c.weight.dtype))
        self.conv_in2.load_state_dict(c.state_dict())

    def _stem(self, sample):
        noise, cond, cond2 = sample.chunk(3, dim=-3)                                    # reference :494
        return self.conv_in(torch.cat([noise, cond], -3)) + \
            self.conv_in2(torch.cat([noise, cond2], -3)) * self.conv_in2_alpha          # :499-502


# --------------------------------------------------------------------------- x / y input heads
class UNetSpatioTemporalConditionJointModel(UNetSpatioTemporalConditionControlNetModel):
    """``models/unet_spatio_temporal_condition_joint.py``: a second set of INPUT heads (``conv_in_y``,
    ``time_embedding_y``, ``add_embedding_y``; ``add_y_input_head`` :251-280 deep-copies the x heads) and a forward that
    routes every sample of the batch through the x or the y heads according to the boolean masks
    ``lora_mask["xy_lora"]`` / ``lora_mask["yx_lora"]`` (:483-500; set by ``patch.set_patch_lora_mask``,
    patch/patch.py:872-896).  ``add_time_proj`` is shared (:414), the body is the plain UNet's.  ``timestep`` must be a
    [batch] tensor (the reference indexes it with the masks)."""

    def add_y_input_head(self):
        import copy
        self.conv_in_y = copy.deepcopy(self.conv_in)
        self.time_embedding_y = copy.deepcopy(self.time_embedding)
        self.add_embedding_y = copy.deepcopy(self.add_embedding)

    def _masks(self, batch):
        x_mask, y_mask = self.lora_mask["xy_lora"], self.lora_mask["yx_lora"]
        return (x_mask.repeat_interleave(batch // len(x_mask)), y_mask.repeat_interleave(batch // len(y_mask)))

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, added_time_ids=None):
        b, f = sample.shape[:2]
        t = timestep.expand(b) if timestep.ndim else timestep[None].expand(b)
        x_mask, y_mask = self._masks(b)
        c0 = self.config.block_out_channels[0]
        stem = torch.empty(b * f, c0, *sample.shape[-2:], dtype=sample.dtype)
        emb = torch.empty(b * f, c0 * 4, dtype=sample.dtype)
        for mask, conv, te, ae in ((x_mask, self.conv_in, self.time_embedding, self.add_embedding),
                                   (y_mask, self.conv_in_y, self.time_embedding_y, self.add_embedding_y)):
            if not mask.any():
                continue
            e = te(timestep_embedding(t[mask], c0).to(sample.dtype))                                   # :404-410
            ids = timestep_embedding(added_time_ids[mask].flatten(), self.config.addition_time_embed_dim)
            e = e + ae(ids.reshape(int(mask.sum()), -1).to(e.dtype))                                      # :414-418
            fm = mask.repeat_interleave(f)
            emb[fm] = e.repeat_interleave(f, dim=0)                                                      # :425, :499-500
            stem[fm] = conv(sample[mask].flatten(0, 1))                                                  # :430, :497-498
        ctx = encoder_hidden_states.repeat_interleave(f, dim=0)
        self._joint_stem = stem
        out = self._body(sample.flatten(0, 1), emb, ctx, b, f, down_block_additional_residuals,
                         mid_block_additional_residual)
        self._joint_stem = None
        return SimpleNamespace(sample=out) if return_dict else (out,)

    def _stem(self, sample):
        return self._joint_stem


# --------------------------------------------------------------------------- LKGD (A.7, A.8)
class QuaternionLinear(nn.Module):
    """``core_qnn.quaternion_layers.QuaternionLinearAutograd`` (un-vendored, un-pinned; F5/U5).
    y = x @ W + bias with the Hamilton-product block matrix W (rows = input quarters)."""

    def __init__(self, in_features: int, out_features: int):
        super().__init__()
        i4, o4 = in_features // 4, out_features // 4
        bound = (1.0 / (2.0 * (i4 + o4))) ** 0.5
        for n in ("r_weight", "i_weight", "j_weight", "k_weight"):
            setattr(self, n, nn.Parameter(torch.empty(i4, o4).uniform_(-bound, bound)))
        self.bias = nn.Parameter(torch.zeros(out_features))

    def hamilton(self):
        r, i, j, k = self.r_weight, self.i_weight, self.j_weight, self.k_weight
        return torch.cat([torch.cat([r, -i, -j, -k], 0), torch.cat([i, r, -k, j], 0),
                          torch.cat([j, k, r, -i], 0), torch.cat([k, -j, i, r], 0)], 1)

    def forward(self, x):
        return x @ self.hamilton() + self.bias


class UNetSpatioTemporalConditionModel(UNetSpatioTemporalConditionControlNetModel):
    """LKGD UNet: a1 + latent-knowledge conditioning
    (reference ``models/unet_spatio_temporal_condition.py:197-225,536-613``).  The latent-knowledge
    modules hard-code 1024/256 widths, so ``cross_attention_dim`` must be 1024."""

    def _init_root(self):
        self.quaternion_lora_texts = nn.Parameter(torch.zeros(256))
        self.quaternion_lora_texts_fft_mag = nn.Parameter(torch.zeros(129))
        self.quaternion_lora_texts_fft_pha = nn.Parameter(torch.zeros(129))

    def _init_extra(self):
        def dw():
            return nn.Conv1d(1024, 256, kernel_size=1, groups=256, bias=False)
        self.quaternion_lora_dconv, self.quaternion_lora_lconv, self.quaternion_lora_fconv = dw(), dw(), dw()
        self.quaternion_lora_fuse = QuaternionLinear(1024, 512)
        self.quaternion_lora_fuse_fft_mag = QuaternionLinear(512, 256)
        self.quaternion_lora_fuse_fft_pha = QuaternionLinear(512, 256)
        self.quaternion_lora_fuse_fft_mag0 = nn.Linear(4, 1)
        self.quaternion_lora_fuse_fft_pha0 = nn.Linear(4, 1)
        self.quaternion_lora_fuse_sf = nn.Sequential(nn.Linear(1024, 256), nn.LeakyReLU(0.1), nn.Linear(256, 1024))

    def _condition(self, encoder_hidden_states, domain_features, flow_features):
        """reference ``unet_spatio_temporal_condition.py:536-595``."""
        def lower(conv, x):
            return conv(x.permute(0, 2, 1)).permute(0, 2, 1)
        lh = lower(self.quaternion_lora_lconv, encoder_hidden_states)                               # [B,1,256]
        ld = lower(self.quaternion_lora_dconv, F.interpolate(domain_features, size=1024, mode="linear"))
        lf = lower(self.quaternion_lora_fconv, F.interpolate(flow_features, size=1024, mode="linear"))
        if ld.shape[0] != lh.shape[0] and ld.shape[0] == 1:     # D8: keyed on the domain batch only
            ld = torch.cat([ld, ld], 0)
            lf = torch.cat([lf, lf], 0)
        texts = self.quaternion_lora_texts.expand_as(lh)
        spatial = self.quaternion_lora_fuse(torch.cat([lh, ld, lf, texts], -1))                     # [B,1,512]

        ffts = [torch.fft.rfft(v, dim=-1) for v in (lh, ld, lf)]                                    # [B,1,129]
        if getattr(self, "canonical_zero_phase", False):
            # The unconditional CFG half feeds an all-zero CLIP embedding (pipeline...controlnet.py:206-212), so ``lh`` is
            # exactly zero there and the phase fed to ``fuse_fft_pha`` is the angle of 0 + 0j - decided by the SIGN of the
            # zeros the FFT library returns: torch's CPU backend (pocketfft) returns -0 real parts in 63 of the 129 bins
            # (angle = pi), cuFFT - what the reference runs on the GPU - returns +0 everywhere (angle = 0; checked on a
            # B200, tests/test_unet_gpu.py::test_lkgd_zero_embedding_follows_the_gpu_fft).  Default: this module's CPU
            # behaviour, bit for bit.  With the switch set the zeros are canonicalised to +0 = the GPU reference.
            ffts = [torch.complex(v.real + 0.0, v.imag + 0.0) for v in ffts]
        mags = [torch.abs(v) for v in ffts] + [self.quaternion_lora_texts_fft_mag.expand(ffts[0].shape)]
        phas = [torch.angle(v) for v in ffts] + [self.quaternion_lora_texts_fft_pha.expand(ffts[0].shape)]
        mag = self.quaternion_lora_fuse_fft_mag(torch.cat([m[..., :-1] for m in mags], -1))          # [B,1,256]
        pha = self.quaternion_lora_fuse_fft_pha(torch.cat([p[..., :-1] for p in phas], -1))
        spec = torch.complex(mag * torch.cos(pha), mag * torch.sin(pha))
        mag0 = self.quaternion_lora_fuse_fft_mag0(torch.cat([m[..., -1] for m in mags], -1))         # [B,1]
        pha0 = self.quaternion_lora_fuse_fft_pha0(torch.cat([p[..., -1] for p in phas], -1))
        spec0 = torch.complex(mag0 * torch.cos(pha0), mag0 * torch.sin(pha0))
        spec = torch.cat([spec, spec0.unsqueeze(-1)], -1)                                           # [B,1,257]
        freq = torch.fft.irfft(spec, dim=-1)                                                        # [B,1,512]
        return self.quaternion_lora_fuse_sf(torch.cat([spatial, freq], -1))                         # [B,1,1024]

    def forward(self, sample, timestep, encoder_hidden_states, domain_features, flow_features,
                down_block_additional_residuals=None, mid_block_additional_residual=None, return_dict=True,
                added_time_ids=None):
        b, f = sample.shape[:2]
        emb = self._time_embedding(sample, timestep, added_time_ids)
        ctx = self._condition(encoder_hidden_states, domain_features, flow_features).repeat_interleave(f, dim=0)
        out = self._body(sample.flatten(0, 1), emb, ctx, b, f, down_block_additional_residuals,
                         mid_block_additional_residual)
        return SimpleNamespace(sample=out) if return_dict else (out,)
