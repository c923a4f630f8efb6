"""``AutoencoderKLTemporalDecoder`` of the SVD pipelines (VAE half) on the lkgd_b200 kernels.

Reference call sites: ``pipeline/pipeline_stable_video_diffusion_controlnet.py:216-237`` (``_encode_vae_image``:
``vae.encode(image).latent_dist.mode()``), ``:268-295`` (``decode_latents``: ``vae.decode(z / scaling_factor,
num_frames=chunk).sample`` over chunks of ``decode_chunk_size`` frames, ``:636``) and, in training, ``utils/util.py:234-248``
(``tensor_to_vae_latent``: ``encode(x).latent_dist.sample() * scaling_factor``).  The class itself is diffusers 0.27.2's
(un-vendored); parameter names are the checkpoint's, so the ``vae/`` folder of an SVD checkpoint loads with
``load_state_dict`` / ``from_pretrained``.

Execution (channels-last rows, fp32 residual stream, bf16 GEMM operands - the UNet's conventions):
  * every 3x3 conv is an implicit-GEMM ``lkgd_gemm`` launch whose epilogue adds bias / residual / the 1x1 shortcut segment and
    accumulates the GroupNorm statistics of what it stores; the encoder's ``Downsample2D(padding=0)`` is the stride-2 conv with
    bottom / right padding (``pad_br``); ``quant_conv`` (1x1, 8 -> 8) is folded into ``encoder.conv_out`` at pack time;
  * the decoder's ``SpatioTemporalResBlock`` is the UNet's (``engine.run_resblock``) without a time embedding and with the
    ``switch_spatial_to_temporal_mix`` blender; the temporal GroupNorms / (3,1,1) convs see the frames of ONE decode chunk;
  * the mid-block attention has one 512-wide head: Q K^T (GEMM, fp32) -> ``lkgd_softmax_rows`` -> P V (GEMM against V^T,
    which a GEMM with swapped operands produces directly; V's bias is added after P V since softmax rows sum to one);
    head widths <= 128 (small test configurations) take the flash kernel;
  * ``conv_out`` stores fp32 rows and ``lkgd_time_conv_out`` applies the (3,1,1) frame convolution while unpacking to NCHW."""
from __future__ import annotations

import json
import os
from types import SimpleNamespace
from typing import List, Optional, Sequence

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from .engine import Geom, Norm, PackedResBlock, _b16, _conv3x3_weight, _f32, run_resblock
from .ops import A_CONV3X3, bf16

SVD_VAE_CONFIG = dict(in_channels=3, out_channels=3, block_out_channels=(128, 256, 512, 512), layers_per_block=2,
                      latent_channels=4, sample_size=768, scaling_factor=0.18215, force_upcast=True)
IN_CPAD = 64      # input channels of the two conv_in layers padded to one 64-element K block


# ------------------------------------------------------------------------------------------------- parameter containers
class VaeAttention(M.Container):
    def __init__(self, channels, dim_head, eps=1e-6, groups=32):
        super().__init__()
        self.heads, self.dim_head = channels // dim_head, dim_head
        self.group_norm = M.GroupNorm(groups, channels, eps=eps)
        self.to_q, self.to_k, self.to_v = (M.Linear(channels, channels) for _ in range(3))
        self.to_out = nn.ModuleList([M.Linear(channels, channels), nn.Dropout(0.0)])


class DownEncoderBlock2D(M.Container):
    def __init__(self, cin, cout, num_layers, add_downsample):
        super().__init__()
        self.resnets = nn.ModuleList([M.ResnetBlock2D(cin if i == 0 else cout, cout, None, 1e-6) for i in range(num_layers)])
        self.downsamplers = nn.ModuleList([M.Downsample2D(cout, padding=0)]) if add_downsample else None


class UNetMidBlock2D(M.Container):
    def __init__(self, channels, attention_head_dim):
        super().__init__()
        self.resnets = nn.ModuleList([M.ResnetBlock2D(channels, channels, None, 1e-6) for _ in range(2)])
        self.attentions = nn.ModuleList([VaeAttention(channels, attention_head_dim)])


class Encoder(M.Container):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        self.conv_in = M.Conv2d(in_channels, block_out_channels[0], 3, padding=1)
        blocks, c = [], block_out_channels[0]
        for i, co in enumerate(block_out_channels):
            blocks.append(DownEncoderBlock2D(c, co, layers_per_block, i != len(block_out_channels) - 1))
            c = co
        self.down_blocks = nn.ModuleList(blocks)
        self.mid_block = UNetMidBlock2D(c, c)
        self.conv_norm_out = M.GroupNorm(32, c, eps=1e-6)
        self.conv_out = M.Conv2d(c, 2 * out_channels, 3, padding=1)


def _temporal_resblock(cin, cout):
    return M.SpatioTemporalResBlock(cin, cout, None, 1e-6, temporal_eps=1e-5, merge_factor=0.0, merge_strategy="learned",
                                    switch_spatial_to_temporal_mix=True)


class MidBlockTemporalDecoder(M.Container):
    def __init__(self, channels, attention_head_dim, num_layers):
        super().__init__()
        self.resnets = nn.ModuleList([_temporal_resblock(channels, channels) for _ in range(num_layers)])
        self.attentions = nn.ModuleList([VaeAttention(channels, attention_head_dim)])


class UpBlockTemporalDecoder(M.Container):
    def __init__(self, cin, cout, num_layers, add_upsample):
        super().__init__()
        self.resnets = nn.ModuleList([_temporal_resblock(cin if i == 0 else cout, cout) for i in range(num_layers)])
        self.upsamplers = nn.ModuleList([M.Upsample2D(cout)]) if add_upsample else None


class TemporalDecoder(M.Container):
    def __init__(self, in_channels, out_channels, block_out_channels, layers_per_block):
        super().__init__()
        top = block_out_channels[-1]
        self.conv_in = M.Conv2d(in_channels, top, 3, padding=1)
        self.mid_block = MidBlockTemporalDecoder(top, top, layers_per_block)
        rev, blocks, c = list(reversed(block_out_channels)), [], top
        for i, co in enumerate(rev):
            blocks.append(UpBlockTemporalDecoder(c, co, layers_per_block + 1, i != len(rev) - 1))
            c = co
        self.up_blocks = nn.ModuleList(blocks)
        self.conv_norm_out = M.GroupNorm(32, block_out_channels[0], eps=1e-6)
        self.conv_out = M.Conv2d(block_out_channels[0], out_channels, 3, padding=1)
        self.time_conv_out = M.Conv3d(out_channels, out_channels, (3, 1, 1), padding=(1, 0, 0))


class DiagonalGaussianDistribution:
    """diffusers ``DiagonalGaussianDistribution`` over the encoder's moments (a [B, 2*latent, h, w] tensor: elementwise host-side
    arithmetic on a latent-sized tensor, no kernel)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)
        self.var = torch.exp(self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        noise = torch.randn(self.mean.shape, generator=generator, device=self.mean.device if generator is None
                            else generator.device, dtype=self.mean.dtype).to(self.mean.device)
        return self.mean + self.std * noise


# ------------------------------------------------------------------------------------------------- kernel-ready weights
class _PackedResnet2D:
    def __init__(self, r: M.ResnetBlock2D):
        self.cin, self.cout = r.in_channels, r.out_channels
        self.n1, self.n2 = Norm.of(r.norm1), Norm.of(r.norm2)
        self.w1, self.b1 = _conv3x3_weight(r.conv1)
        self.w2, self.b2 = _conv3x3_weight(r.conv2)
        self.wsc = None
        if r.conv_shortcut is not None:
            self.wsc = _b16(r.conv_shortcut.weight.reshape(self.cout, self.cin))
            self.b2 = (self.b2 + _f32(r.conv_shortcut.bias)).contiguous()


class _PackedAttention:
    def __init__(self, a: VaeAttention):
        self.heads, self.d = a.heads, a.dim_head
        self.c = self.heads * self.d
        self.norm = Norm.of(a.group_norm)
        if self.d % 8:
            raise ValueError("VAE attention head width must be a multiple of 8")
        self.flash = self.d <= 128
        if self.flash:
            self.wqkv = _b16(torch.cat([a.to_q.weight, a.to_k.weight, a.to_v.weight], 0))
            self.bqkv = _f32(torch.cat([a.to_q.bias, a.to_k.bias, a.to_v.bias], 0))
        else:
            self.wqk = _b16(torch.cat([a.to_q.weight, a.to_k.weight], 0))
            self.bqk = _f32(torch.cat([a.to_q.bias, a.to_k.bias], 0))
            self.wv, self.bv = _b16(a.to_v.weight), _f32(a.to_v.bias)
        self.wo, self.bo = _b16(a.to_out[0].weight), _f32(a.to_out[0].bias)


def _run_resnet2d(p: _PackedResnet2D, x: torch.Tensor, n: int, H: int, W: int) -> torch.Tensor:
    HW = H * W
    xa = None
    if p.wsc is not None:
        h, xa = ops.groupnorm(x, p.n1.g, p.n1.b, p.n1.eps, NS=n, R=HW, silu=True, want_raw=True)
        if xa is None:
            xa = ops.cast_bf16(x)
    else:
        h = ops.groupnorm(x, p.n1.g, p.n1.b, p.n1.eps, NS=n, R=HW, silu=True)
    h = ops.gemm(h, p.w1, mode=A_CONV3X3, conv=(n, H, W, 1), bias=p.b1, out_f32=True, gn_rows=HW)
    h = ops.groupnorm(h, p.n2.g, p.n2.b, p.n2.eps, NS=n, R=HW, silu=True)
    if p.wsc is not None:       # the 1x1 shortcut conv rides as the second K segment of conv2
        return ops.gemm(h, p.w2, mode=A_CONV3X3, conv=(n, H, W, 1), bias=p.b2, A1=xa, Bw1=p.wsc, out_f32=True, gn_rows=HW)
    return ops.gemm(h, p.w2, mode=A_CONV3X3, conv=(n, H, W, 1), bias=p.b2, res1=x, out_f32=True, gn_rows=HW)


def _run_attention(p: _PackedAttention, x: torch.Tensor, n: int, HW: int) -> torch.Tensor:
    """x fp32 [n*HW, C] -> x + to_out(attention(GroupNorm(x))) (fp32, with fused GroupNorm statistics when HW allows)."""
    C = p.c
    t = ops.groupnorm(x, p.norm.g, p.norm.b, p.norm.eps, NS=n, R=HW, silu=False)
    scale = p.d ** -0.5
    if p.flash:
        qkv = ops.gemm(t, p.wqkv, bias=p.bqkv)
        o = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_img=n, heads=p.heads, d=p.d, Nq=HW, Nk=HW)
    else:
        if HW % 8:
            raise ValueError(f"VAE attention over {HW} tokens: H*W of the latent must be a multiple of 8")
        qk = ops.gemm(t, p.wqk, bias=p.bqk)                                   # [n*HW, 2C]: q | k
        o = torch.empty((n * HW, C), device=x.device, dtype=bf16)
        S = torch.empty((HW, HW), device=x.device, dtype=torch.float32)       # reused by every (image, head): one stream
        P = torch.empty((HW, HW), device=x.device, dtype=bf16)
        for i in range(n):
            rows = slice(i * HW, (i + 1) * HW)
            vT = ops.gemm(p.wv, t[rows])                                      # [C, HW] = W_v t^T (V^T without its bias)
            for hd in range(p.heads):
                cs = slice(hd * p.d, (hd + 1) * p.d)
                ops.gemm(qk[rows, cs], qk[rows, C + hd * p.d:C + (hd + 1) * p.d], out=S, out_f32=True)
                ops.softmax_rows(S, scale, out=P)
                ops.gemm(P, vT[cs], bias=p.bv[cs], out=o[rows, cs])           # rows of P sum to 1: + b_v afterwards
    return ops.gemm(o, p.wo, bias=p.bo, res1=x, out_f32=True, gn_rows=HW if HW % 128 == 0 else 0)


class _PackedVae:
    def __init__(self, vae: "AutoencoderKLTemporalDecoder"):
        e, d = vae.encoder, vae.decoder
        lat = vae.config.latent_channels
        # ---- encoder
        self.e_in_w, self.e_in_b = _conv3x3_weight(e.conv_in, cin_pad=IN_CPAD)
        self.e_down = []
        for b in e.down_blocks:
            ds = None
            if b.downsamplers is not None:
                ds = _conv3x3_weight(b.downsamplers[0].conv)
            self.e_down.append(([_PackedResnet2D(r) for r in b.resnets], ds))
        self.e_mid = ([_PackedResnet2D(r) for r in e.mid_block.resnets], _PackedAttention(e.mid_block.attentions[0]))
        self.e_norm = Norm.of(e.conv_norm_out)
        # quant_conv (1x1) folded into conv_out: moments = Wq (Wc * x + bc) + bq
        wq = vae.quant_conv.weight.detach().float().reshape(2 * lat, 2 * lat)
        wc = torch.einsum("om,mikl->oikl", wq, e.conv_out.weight.detach().float())
        bc = wq @ e.conv_out.bias.detach().float() + vae.quant_conv.bias.detach().float()
        folded = SimpleNamespace(weight=wc, bias=bc)
        self.e_out_w, self.e_out_b = _conv3x3_weight(folded, cout_pad=max(32, (2 * lat + 15) // 16 * 16))
        self.moments = 2 * lat
        # ---- decoder
        self.d_in_w, self.d_in_b = _conv3x3_weight(d.conv_in, cin_pad=IN_CPAD)
        self.d_mid = ([PackedResBlock(r) for r in d.mid_block.resnets], _PackedAttention(d.mid_block.attentions[0]))
        self.d_up = []
        for b in d.up_blocks:
            us = _conv3x3_weight(b.upsamplers[0].conv) if b.upsamplers is not None else None
            self.d_up.append(([PackedResBlock(r) for r in b.resnets], us))
        self.d_norm = Norm.of(d.conv_norm_out)
        self.cout = vae.config.out_channels
        self.d_out_w, self.d_out_b = _conv3x3_weight(d.conv_out, cout_pad=32)
        self.t_out_w = _f32(d.time_conv_out.weight[..., 0, 0])                 # [C, C, 3]
        self.t_out_b = _f32(d.time_conv_out.bias)


# ------------------------------------------------------------------------------------------------- the module
class AutoencoderKLTemporalDecoder(nn.Module):
    """Drop-in for ``diffusers.AutoencoderKLTemporalDecoder`` as the reference uses it: ``.config.scaling_factor``,
    ``encode(x).latent_dist.mode() / .sample()``, ``decode(z, num_frames=n).sample``, ``forward`` accepting ``num_frames``
    (the pipeline inspects its signature, pipeline...controlnet.py:274-275)."""

    def __init__(self, in_channels=3, out_channels=3, block_out_channels: Sequence[int] = (128, 256, 512, 512),
                 layers_per_block=2, latent_channels=4, sample_size=768, scaling_factor=0.18215, force_upcast=True, **_):
        super().__init__()
        boc = tuple(block_out_channels)
        if any(c % 32 for c in boc):
            raise ValueError("block_out_channels must be multiples of 32 (GroupNorm groups)")
        if out_channels not in (1, 3, 4) or in_channels > IN_CPAD or latent_channels > 8:
            raise ValueError("unsupported VAE channel counts (image channels 1 / 3 / 4, at most 8 latent channels)")
        self.config = SimpleNamespace(in_channels=in_channels, out_channels=out_channels, block_out_channels=boc,
                                      layers_per_block=layers_per_block, latent_channels=latent_channels,
                                      sample_size=sample_size, scaling_factor=scaling_factor, force_upcast=force_upcast)
        self.encoder = Encoder(in_channels, latent_channels, boc, layers_per_block)
        self.decoder = TemporalDecoder(latent_channels, out_channels, boc, layers_per_block)
        self.quant_conv = M.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self._pk: Optional[_PackedVae] = None

    # ---- plumbing shared with the other modules
    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def invalidate(self):
        self._pk = None

    def load_state_dict(self, sd, *a, **k):
        out = super().load_state_dict(sd, *a, **k)
        self.invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._pk = None
        return out

    @classmethod
    def from_pretrained(cls, path: str, subfolder: Optional[str] = None, device=None):
        """Local directory with ``config.json`` + ``diffusion_pytorch_model.safetensors`` (the ``vae`` sub-folder of an SVD
        checkpoint)."""
        root = os.path.join(path, subfolder) if subfolder else path
        cfg_path = os.path.join(root, "config.json")
        if not os.path.isfile(cfg_path):
            raise EnvironmentError(f"{cfg_path} not found: lkgd_b200 loads local model directories only")
        cfg = {k: v for k, v in json.load(open(cfg_path)).items() if not k.startswith("_")}
        net = cls(**cfg)
        for name in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
                     "diffusion_pytorch_model.bin"):
            f = os.path.join(root, name)
            if os.path.isfile(f):
                if f.endswith(".safetensors"):
                    from safetensors.torch import load_file
                    sd = load_file(f)
                else:
                    sd = torch.load(f, map_location="cpu", weights_only=True)
                net.load_state_dict(sd, strict=True)
                return net.to(device) if device is not None else net
        raise EnvironmentError(f"no diffusion_pytorch_model.safetensors / .bin under {root}")

    def _pack(self) -> _PackedVae:
        if self._pk is None:
            if self.device.type != "cuda":
                raise RuntimeError("lkgd_b200 runs on CUDA devices only (there is no CPU or PyTorch fallback)")
            self._pk = _PackedVae(self)
        return self._pk

    # ---- encoder
    @ops.on_own_device
    @torch.no_grad()
    def encode(self, x: torch.Tensor, return_dict: bool = True):
        """x [B, 3, H, W] (H, W multiples of 8) -> ``.latent_dist`` over fp32 moments [B, 2*latent, H/8, W/8]."""
        c = self.config
        if x.ndim != 4 or x.shape[1] != c.in_channels:
            raise ValueError(f"encode expects [B, {c.in_channels}, H, W], got {tuple(x.shape)}")
        n, _, H, W = x.shape
        down = len(c.block_out_channels) - 1
        if H % (1 << down) or W % (1 << down):
            raise ValueError(f"image height and width must be multiples of {1 << down}")
        pk = self._pack()
        dev = self.device
        ops.STATS_ARENA.begin(dev)
        rows = ops.pack_input(x.to(dev)[:, None], 1.0, None, n, IN_CPAD)                  # [n*H*W, 64] bf16
        h = ops.gemm(rows, pk.e_in_w, mode=A_CONV3X3, conv=(n, H, W, 1), bias=pk.e_in_b, out_f32=True, gn_rows=H * W)
        for res, ds in pk.e_down:
            for r in res:
                h = _run_resnet2d(r, h, n, H, W)
            if ds is not None:
                hb = ops.cast_bf16(h)
                H, W = H // 2, W // 2
                h = ops.gemm(hb, ds[0], mode=A_CONV3X3, conv=(n, 2 * H, 2 * W, 2), pad_br=True, bias=ds[1], out_f32=True,
                             gn_rows=H * W)
        res, att = pk.e_mid
        h = _run_resnet2d(res[0], h, n, H, W)
        h = _run_attention(att, h, n, H * W)
        h = _run_resnet2d(res[1], h, n, H, W)
        hn = ops.groupnorm(h, pk.e_norm.g, pk.e_norm.b, pk.e_norm.eps, NS=n, R=H * W, silu=True)
        mom = ops.gemm(hn, pk.e_out_w, mode=A_CONV3X3, conv=(n, H, W, 1), bias=pk.e_out_b, out_f32=True, n_store=pk.moments)
        moments = ops.unpack_output(mom, n, 1, pk.moments, H, W)[:, 0]
        dist = DiagonalGaussianDistribution(moments.to(x.dtype if x.is_floating_point() else torch.float32))
        return SimpleNamespace(latent_dist=dist) if return_dict else (dist,)

    # ---- temporal decoder
    @ops.on_own_device
    @torch.no_grad()
    def decode(self, z: torch.Tensor, num_frames: int, return_dict: bool = True):
        """z [B*num_frames, latent, h, w] -> ``.sample`` [B*num_frames, 3, 8h, 8w]; the temporal layers mix the ``num_frames``
        consecutive frames of each batch element (the reference passes one ``decode_chunk_size`` chunk at a time)."""
        c = self.config
        if z.ndim != 4 or z.shape[1] != c.latent_channels or num_frames <= 0 or z.shape[0] % num_frames:
            raise ValueError(f"decode expects [B*num_frames, {c.latent_channels}, h, w], got {tuple(z.shape)} with "
                             f"num_frames={num_frames}")
        n, _, H, W = z.shape
        up = len(c.block_out_channels) - 1
        if n * H * W * (1 << (2 * up)) >= 2 ** 31:
            raise ValueError("decode chunk too large for one launch (2^31 output pixels): lower decode_chunk_size")
        pk = self._pack()
        dev = self.device
        g = Geom(n // num_frames, num_frames, H, W)
        ops.STATS_ARENA.begin(dev)
        rows = ops.pack_input(z.to(dev)[:, None], 1.0, None, n, IN_CPAD)
        h = ops.gemm(rows, pk.d_in_w, mode=A_CONV3X3, conv=(n, H, W, 1), bias=pk.d_in_b, out_f32=True, gn_rows=g.HW)
        res, att = pk.d_mid
        h = run_resblock(res[0], h, None, g, None)
        for r in res[1:]:
            h = _run_attention(att, h, n, g.HW)
            h = run_resblock(r, h, None, g, None)
        for res, us in pk.d_up:
            for r in res:
                h = run_resblock(r, h, None, g, None)
            if us is not None:
                hu = ops.upsample2x(h, n, g.H, g.W)
                g = g.up()
                h = ops.gemm(hu, us[0], mode=A_CONV3X3, conv=(n, g.H, g.W, 1), bias=us[1], out_f32=True, gn_rows=g.HW)
        hn = ops.groupnorm(h, pk.d_norm.g, pk.d_norm.b, pk.d_norm.eps, NS=n, R=g.HW, silu=True)
        y = ops.gemm(hn, pk.d_out_w, mode=A_CONV3X3, conv=(n, g.H, g.W, 1), bias=pk.d_out_b, out_f32=True, n_store=4)
        out = ops.time_conv_out(y, pk.t_out_w, pk.t_out_b, g.B, g.F, g.H, g.W)
        out = out.to(z.dtype if z.is_floating_point() else torch.float32)
        return SimpleNamespace(sample=out) if return_dict else (out,)

    def forward(self, sample: torch.Tensor, sample_posterior: bool = False, return_dict: bool = True,
                generator: Optional[torch.Generator] = None, num_frames: int = 1):
        post = self.encode(sample).latent_dist
        z = post.sample(generator) if sample_posterior else post.mode()
        return self.decode(z, num_frames=num_frames, return_dict=return_dict)


def decode_latents(vae: AutoencoderKLTemporalDecoder, latents: torch.Tensor, num_frames: int, decode_chunk_size: int = 14
                   ) -> torch.Tensor:
    """The reference pipelines' ``decode_latents`` (pipeline...controlnet.py:268-295): [B, F, C, h, w] -> fp32
    [B, 3, F, 8h, 8w], ``decode_chunk_size`` frames at a time."""
    latents = latents.flatten(0, 1) / vae.config.scaling_factor
    frames: List[torch.Tensor] = []
    for i in range(0, latents.shape[0], decode_chunk_size):
        chunk = latents[i:i + decode_chunk_size]
        frames.append(vae.decode(chunk, num_frames=chunk.shape[0]).sample)
    frames = torch.cat(frames, dim=0)
    return frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()


def encode_vae_image(vae: AutoencoderKLTemporalDecoder, image: torch.Tensor, num_videos_per_prompt: int = 1,
                     do_classifier_free_guidance: bool = True) -> torch.Tensor:
    """``_encode_vae_image`` (pipeline...controlnet.py:216-237): posterior mode (UNSCALED, as the reference), zero unconditional
    half first, repeated per video."""
    lat = vae.encode(image.to(vae.device)).latent_dist.mode()
    if do_classifier_free_guidance:
        lat = torch.cat([torch.zeros_like(lat), lat])
    return lat.repeat(num_videos_per_prompt, 1, 1, 1)
