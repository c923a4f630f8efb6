"""Block-graph executor of the SVD / LKGD spatio-temporal UNet on the lkgd_b200 CUDA kernels.

Resident activation layout: channels-last, one row per (batch, frame, pixel): ``[B*F*H*W, C]``.  The RESIDUAL
STREAM (block inputs / outputs, skip tensors, the running hidden state inside a transformer) is kept in fp32 - with
~190 sequential residual updates per forward, rounding it to bf16 at every add costs ~1.5e-2 rel-L2 by itself, above
the 1e-2 parity bar; everything a tensor-core GEMM reads (norm outputs, attention outputs, GEGLU products) is bf16.  Spatial
ops see it as [B*F, HW, C]; temporal ops address the same buffer as [B, F, HW, C] (frame stride HW*C) - there
is no physical transpose between the spatial and temporal halves of a block (the reference alternates NCHW,
[BF,HW,C] and [B*HW,F,C]: patch/patch.py:592-597,682-684).

``Packed*`` objects hold kernel-ready weights (bf16 [N,K] K-major, conv kernels as [Cout, tap, Cin], fused QKV,
tile-interleaved GEGLU, fp32 norms / embeddings MLPs).  The arithmetic order follows SURVEY.md Appendix A
(diffusers 0.27.2) and the reference forward ``models/unet_spatio_temporal_condition_controlnet.py:358-508``.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import modules as M
from . import ops
from .ops import (A_CONV3X3, A_LINEAR, A_TCONV3, ACT_GEGLU, ACT_NONE, ACT_SILU, RV_BATCH, RV_FRAMEPOS, RV_NONE,
                  RV_BATCH_TCTX, RV_TCTX_0272, bf16)

SL_SILU, SL_LEAKY = 1, 3


def _f32(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(torch.float32).contiguous()


def _b16(t: torch.Tensor) -> torch.Tensor:
    return t.detach().to(bf16).contiguous()


def _lin_parts(mod):
    """(weight fp32 [N,K], bias fp32 | None, lora) of a Linear / LoraLinear.  ``lora`` is None or
    (A [r,K], B*scaling [N,r], masks): all ACTIVE adapters stacked along r (their updates add up, reference
    models/lora_layer.py:425-442); ``masks`` = None, or a list of (row_lo, row_hi, bool mask over the batch) for the
    layers on which the masked forward (patch/patch.py:57-92) restricts at least one adapter to some samples (mask None =
    that adapter acts on every sample)."""
    if isinstance(mod, M.LoraLinear):
        base = mod.base_layer
        lora = None
        ads = [] if mod.merged else mod.adapters()
        if ads:
            a = torch.cat([_f32(x[1]) for x in ads], 0)
            b = torch.cat([_f32(x[2]) * x[3] for x in ads], 1)
            masks, off = [], 0
            for _, aw, _, _, m in ads:            # once one adapter is masked every adapter gets its own column range
                masks.append((off, off + aw.shape[0], m))
                off += aw.shape[0]
            lora = (a, b, masks if any(m[2] is not None for m in masks) else None)
        return _f32(base.weight), (None if base.bias is None else _f32(base.bias)), lora
    return _f32(mod.weight), (None if mod.bias is None else _f32(mod.bias)), None


def _merged_weight(mod) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    w, b, lora = _lin_parts(mod)
    if lora is not None:
        if lora[2] is not None:
            raise NotImplementedError("a per-sample masked LoRA adapter (patch.hack_lora_forward) sits on a projection "
                                      "whose adapters are merged at pack time (cross-attention attn2, GEGLU proj): "
                                      "masked adapters are built for attn1 / attn1n / proj_in / proj_out / ff.net.2")
        w = w + lora[1] @ lora[0]          # W + s * B A   (reference models/lora_layer.py:406)
    return w, b


def _conv3x3_weight(conv, cin_pad: Optional[int] = None, cout_pad: Optional[int] = None):
    w = conv.weight.detach().to(torch.float32)               # [Cout, Cin, 3, 3]
    co, ci = w.shape[:2]
    cin_pad, cout_pad = cin_pad or ci, cout_pad or co
    wk = torch.zeros(cout_pad, 3, 3, cin_pad, device=w.device, dtype=torch.float32)
    wk[:co, :, :, :ci] = w.permute(0, 2, 3, 1)
    b = torch.zeros(cout_pad, device=w.device, dtype=torch.float32)
    b[:co] = conv.bias.detach().to(torch.float32)
    return wk.reshape(cout_pad, 9 * cin_pad).to(bf16).contiguous(), b


@dataclass
class Norm:
    g: torch.Tensor
    b: torch.Tensor
    eps: float

    @staticmethod
    def of(mod) -> "Norm":
        return Norm(_f32(mod.weight), _f32(mod.bias), float(mod.eps))


@dataclass
class Dense:
    """bf16 GEMM weight [+ LoRA pair] + fp32 bias."""
    w: torch.Tensor
    b: Optional[torch.Tensor]
    lora_a: Optional[torch.Tensor] = None    # [r_pad, K] bf16   (t = x A^T)
    lora_b: Optional[torch.Tensor] = None    # [N, r_pad] bf16   (scaled B)
    lora_masks: Optional[list] = None        # [(row_lo, row_hi of lora_a, bool mask over the batch)]: masked adapters


class PackedResBlock:
    def __init__(self, blk: M.SpatioTemporalResBlock):
        s, t = blk.spatial_res_block, blk.temporal_res_block
        self.cin, self.cout = s.in_channels, s.out_channels
        self.n1, self.n2 = Norm.of(s.norm1), Norm.of(s.norm2)
        self.w1, self.b1 = _conv3x3_weight(s.conv1)
        self.w2, self.b2 = _conv3x3_weight(s.conv2)
        self.has_temb = s.time_emb_proj is not None          # False: the VAE's temporal decoder
        if self.has_temb:
            self.temb_w, self.temb_b = _f32(s.time_emb_proj.weight), _f32(s.time_emb_proj.bias)
        if s.conv_shortcut is not None:
            self.wsc = _b16(s.conv_shortcut.weight.reshape(self.cout, self.cin))
            self.bsc = _f32(s.conv_shortcut.bias)
            self.b2sc = (self.b2 + self.bsc).contiguous()      # conv2 + fused shortcut segment share one bias
        else:
            self.wsc = self.bsc = self.b2sc = None
        self.tn1, self.tn2 = Norm.of(t.norm1), Norm.of(t.norm2)
        c = self.cout
        # Conv3d weight [C, C, 3, 1, 1] -> [C, kt, Cin]
        self.tw1 = _b16(t.conv1.weight[..., 0, 0].permute(0, 2, 1).reshape(c, 3 * c))
        self.tw2 = _b16(t.conv2.weight[..., 0, 0].permute(0, 2, 1).reshape(c, 3 * c))
        self.tb1, self.tb2 = _f32(t.conv1.bias), _f32(t.conv2.bias)
        if self.has_temb:
            self.ttemb_w, self.ttemb_b = _f32(t.time_emb_proj.weight), _f32(t.time_emb_proj.bias)
        self.alpha = float(torch.sigmoid(blk.time_mixer.mix_factor.detach().float()).item())
        if getattr(blk.time_mixer, "switch_spatial_to_temporal_mix", False):
            self.alpha = 1.0 - self.alpha


def _dense(mod, fold_lora: bool) -> Dense:
    w, b, lora = _lin_parts(mod)
    if lora is None:
        return Dense(w.to(bf16).contiguous(), b)
    a, bs, masks = lora
    if not fold_lora and masks is None:
        return Dense((w + bs @ a).to(bf16).contiguous(), b)
    r = a.shape[0]
    r_pad = (r + 7) // 8 * 8
    ap = torch.zeros(r_pad, a.shape[1], device=a.device)
    ap[:r] = a
    bp = torch.zeros(bs.shape[0], r_pad, device=a.device)
    bp[:, :r] = bs
    return Dense(w.to(bf16).contiguous(), b, ap.to(bf16).contiguous(), bp.to(bf16).contiguous(), masks)


def _fused(mods, fold_lora: bool, mask_map=None) -> Dense:
    """Several projections of ONE input fused into one [sum N, K] GEMM (q|k|v, or k|v of the joint branch); LoRA adapters
    become one stacked [n_proj * r, K] down-projection and a block-diagonal [sum N, n_proj * r] up-projection consumed as the
    GEMM's second K segment.  ``mask_map(i, mask)``: optional re-indexing of projection i's per-sample adapter masks."""
    parts = [_lin_parts(m) for m in mods]
    has_lora = any(p[2] is not None for p in parts)
    masked = any(p[2] is not None and p[2][2] is not None for p in parts)
    if not has_lora or (not fold_lora and not masked):
        ws = [(p[0] + p[2][1] @ p[2][0]) if p[2] is not None else p[0] for p in parts]
        return Dense(torch.cat(ws, 0).to(bf16).contiguous(), None)
    np_ = len(parts)
    n, k = parts[0][0].shape
    r = max(p[2][0].shape[0] for p in parts if p[2] is not None)
    r_pad = (r + 7) // 8 * 8
    a_cat = torch.zeros(np_ * r_pad, k, device=parts[0][0].device)
    b_blk = torch.zeros(np_ * n, np_ * r_pad, device=parts[0][0].device)
    masks = []
    for i, p in enumerate(parts):
        if p[2] is None:
            continue
        a, bs, ms = p[2]
        a_cat[i * r_pad:i * r_pad + a.shape[0]] = a
        b_blk[i * n:(i + 1) * n, i * r_pad:i * r_pad + a.shape[0]] = bs
        for lo, hi, m in (ms or []):
            masks.append((i * r_pad + lo, i * r_pad + hi, mask_map(i, m) if mask_map is not None else m))
    w = torch.cat([p[0] for p in parts], 0)
    return Dense(w.to(bf16).contiguous(), None, a_cat.to(bf16).contiguous(), b_blk.to(bf16).contiguous(), masks or None)


def _qkv(attn: M.Attention, fold_lora: bool) -> Dense:
    return _fused((attn.to_q, attn.to_k, attn.to_v), fold_lora)


def _geglu(ff: M.FeedForward, fold_lora: bool):
    w, b = _merged_weight(ff.net[0].proj)
    wp, bp = ops.pack_geglu(w.to(bf16), b)
    return Dense(wp, bp), _dense(ff.net[2], fold_lora)


class PackedCross:
    """attn2: fp32 to_v / to_out for the KV-length-1 collapse + bf16 q/k/v/out for KV > 1.

    Per-sample MASKED adapters (patch.hack_lora_forward, patch/patch.py:57-92) on to_v / to_out.0 make the collapsed matrix a
    function of the sample: ``variant(key)`` builds (and caches) ``(Wo + sum_a s B A)(Wv + sum_a s B A)`` for the adapters
    that are active on it, ``patterns(B)`` says which samples share a matrix (to_q / to_k never matter with one key)."""

    def __init__(self, attn: M.Attention):
        self.heads, self.d = attn.heads, attn.dim_head
        self._attn = attn
        self._general = None
        self._variants = {}
        self.masked = any(isinstance(m, M.LoraLinear) and not m.merged and any(a[4] is not None for a in m.adapters())
                          for m in (attn.to_v, attn.to_out[0]))
        if self.masked:
            self._ads = [[] if not isinstance(m, M.LoraLinear) or m.merged else m.adapters() for m in (attn.to_v, attn.to_out[0])]
            self.bo = _lin_parts(attn.to_out[0])[1]
            self.wov = self.variant(tuple(tuple(True for _ in ads) for ads in self._ads))
        else:
            wv, _ = _merged_weight(attn.to_v)
            wo, self.bo = _merged_weight(attn.to_out[0])
            # softmax over ONE key is exactly 1, so attn2(x, ctx) = to_out(to_v(ctx)) = (Wo Wv) ctx + bo for every query:
            # one [C, D] matrix per cross-attention, product taken in fp64
            self.wov = (wo.double() @ wv.double()).float().contiguous()
        self.off = 0                # column offset inside the batched cross-vector matrix (set by PackedUNet)

    def variant(self, key) -> torch.Tensor:
        if key not in self._variants:
            ws = []
            for mod, ads, on in zip((self._attn.to_v, self._attn.to_out[0]), self._ads, key):
                w = _lin_parts(mod)[0].double()
                for (_, a, b, sc, _), active in zip(ads, on):
                    if active:
                        w = w + sc * (b.detach().double() @ a.detach().double())
                ws.append(w)
            self._variants[key] = (ws[1] @ ws[0]).float().contiguous()
        return self._variants[key]

    def patterns(self, batch: int):
        """{variant key: [samples]} under the reference's ``mask.repeat_interleave(batch // len(mask))``."""
        out = {}
        for b in range(batch):
            key = tuple(tuple(True if m is None else bool(m[b // max(batch // len(m), 1)]) for (_, _, _, _, m) in ads)
                        for ads in self._ads)
            out.setdefault(key, []).append(b)
        return out

    def general(self):
        if self._general is None:
            a = self._attn
            wq, _ = _merged_weight(a.to_q)
            wk, _ = _merged_weight(a.to_k)
            wv, _ = _merged_weight(a.to_v)
            wo, _ = _merged_weight(a.to_out[0])
            self._general = (wq.to(bf16).contiguous(), torch.cat([wk, wv], 0).to(bf16).contiguous(),
                             wo.to(bf16).contiguous())
        return self._general


def _partners(mask: torch.Tensor, batch: int) -> List[int]:
    """Partner sample of every sample under the reference's swap ``enc[~m] = n[m]; enc[m] = n[~m]`` with ``m`` the mask
    repeat-interleaved over the batch (patch/patch.py:444-456): the k-th unmasked sample pairs with the k-th masked one."""
    if mask is None:
        raise ValueError("joint attention needs a joint attention mask (patch.set_joint_attention_mask)")
    if batch % len(mask):
        raise ValueError(f"batch {batch} is not a multiple of the joint attention mask length {len(mask)}")
    m = mask.repeat_interleave(batch // len(mask)).tolist()
    t = [i for i, v in enumerate(m) if v]
    f = [i for i, v in enumerate(m) if not v]
    if len(t) != len(f):
        raise ValueError("the joint attention mask must select half of the batch")
    out = [0] * batch
    for a, b in zip(f, t):
        out[a], out[b] = b, a
    return out


class PackedJoint:
    """Joint-attention branch of a patched block (patch/patch.py ToMeBlock): ``attn1n`` projections (their LoRA adapters
    folded like everywhere else, per-sample masks included) + the post layer and ``joint_scale`` folded into its output
    projection: post(to_out(a)) * s = a (s P Wo)^T + (a A^T) (s P B)^T + s P bo."""

    def __init__(self, blk, spatial: bool, fold_lora: bool = True):
        if not hasattr(blk, "attn1n") or blk.post is None:
            raise ValueError("joint attention is enabled on a block without joint layers: call "
                             "patch.initialize_joint_layers(unet) after patch.apply_patch(unet)")
        a = blk.attn1n
        self.mask = blk.joint_attn_mask
        self.flip = bool(blk.flip) and spatial
        self.q = _dense(a.to_q, fold_lora)
        # k / v are projected from every sample's OWN rows and read by the partner; the reference projects the swapped
        # tensor, with the adapter masks of to_k / to_v stored inverted (patch.py:889-892): position i of the swapped
        # tensor holds sample partner(i), so sample j's rows carry the adapter iff stored_mask[partner(j)]
        def remap(_, m):
            if self.mask is None:
                raise ValueError("masked adapters on attn1n.to_k / to_v need the joint attention mask")
            batch = max(len(m), len(self.mask))
            part = _partners(self.mask, batch)
            mm = m.repeat_interleave(batch // len(m)).tolist()
            return torch.tensor([mm[part[j]] for j in range(batch)], dtype=torch.bool)
        self.kv = _fused((a.to_k, a.to_v), fold_lora, mask_map=remap)
        wo, bo, lora = _lin_parts(a.to_out[0])
        s = float(blk.joint_scale) if spatial else 1.0        # the temporal branch has no joint_scale (patch.py:655)
        self.fuse = None
        if blk.post == "conv":
            P = blk.conv1n.weight.detach().double()
        elif blk.post == "scale":
            P = torch.diag(blk.scale1n.detach().double().reshape(-1))
        else:
            # conv_fuse (patch.py:488-494): [out_masked | out_partner] -> one [2C, 2C] layer -> (masked, partner) halves.  The
            # output projection stays plain; a second two-segment GEMM per sample mixes a sample's rows with its partner's:
            #   masked sample s:   W[:C, :C] o_s + W[:C, C:] o_p        unmasked sample s:   W[C:, C:] o_s + W[C:, :C] o_p
            # The temporal forward has no conv_fuse branch at all (:647-650): the raw attention output is added.
            P = torch.eye(wo.shape[0], dtype=torch.float64, device=wo.device)
            if spatial:
                c = wo.shape[0]
                W = blk.conv1n.weight.detach().float() * s
                s = 1.0
                self.fuse = {True: (W[:c, :c].to(bf16).contiguous(), W[:c, c:].to(bf16).contiguous()),
                             False: (W[c:, c:].to(bf16).contiguous(), W[c:, :c].to(bf16).contiguous())}
        P = P * s
        w_post = (P @ wo.double()).float()
        b_post = (P @ bo.double()).float().contiguous() if bo is not None else None
        if lora is None:
            self.out = Dense(w_post.to(bf16).contiguous(), b_post)
        else:
            la, lb, masks = lora
            r = la.shape[0]
            r_pad = (r + 7) // 8 * 8
            ap = torch.zeros(r_pad, la.shape[1], device=la.device)
            ap[:r] = la
            bp = torch.zeros(lb.shape[0], r_pad, device=la.device)
            bp[:, :r] = (P @ lb.double()).float()
            self.out = Dense(w_post.to(bf16).contiguous(), b_post, ap.to(bf16).contiguous(), bp.to(bf16).contiguous(), masks)


class PackedTransformer:
    def __init__(self, t: M.TransformerSpatioTemporalModel, fold_lora: bool):
        self.heads, self.d, self.c = t.heads, t.dim_head, t.in_channels
        self.src = t                # the module (training looks up the LoRA parameters it owns)
        self.norm = Norm.of(t.norm)
        self.proj_in, self.proj_out = _dense(t.proj_in, fold_lora), _dense(t.proj_out, fold_lora)
        sb, tb = t.transformer_blocks[0], t.temporal_transformer_blocks[0]
        self.s_ln1, self.s_ln2, self.s_ln3 = Norm.of(sb.norm1), Norm.of(sb.norm2), Norm.of(sb.norm3)
        self.s_qkv, self.s_out = _qkv(sb.attn1, fold_lora), _dense(sb.attn1.to_out[0], fold_lora)
        self.s_cross = PackedCross(sb.attn2)
        self.s_ff1, self.s_ff2 = _geglu(sb.ff, fold_lora)
        self.t_lnin, self.t_ln1, self.t_ln2, self.t_ln3 = (Norm.of(tb.norm_in), Norm.of(tb.norm1), Norm.of(tb.norm2),
                                                           Norm.of(tb.norm3))
        self.t_ffin1, self.t_ffin2 = _geglu(tb.ff_in, fold_lora)
        self.t_qkv, self.t_out = _qkv(tb.attn1, fold_lora), _dense(tb.attn1.to_out[0], fold_lora)
        self.t_cross = PackedCross(tb.attn2)
        self.t_ff1, self.t_ff2 = _geglu(tb.ff, fold_lora)
        self.s_joint = PackedJoint(sb, True, fold_lora) if sb.joint_active() else None
        self.t_joint = PackedJoint(tb, False, fold_lora) if tb.joint_active() else None
        self.alpha = float(torch.sigmoid(t.time_mixer.mix_factor.detach().float()).item())
        pe = t.time_pos_embed
        self.pe = (_f32(pe.linear_1.weight), _f32(pe.linear_1.bias), _f32(pe.linear_2.weight), _f32(pe.linear_2.bias))
        self._pos_cache: Dict[int, torch.Tensor] = {}

    def pos_emb(self, F: int) -> torch.Tensor:
        """time_pos_embed(Timesteps(C)(arange(F))) -> fp32 [F, C]; input-independent, cached per F."""
        if F not in self._pos_cache:
            dev = self.pe[0].device
            e = ops.timestep_embedding(torch.arange(F, device=dev, dtype=torch.float32), self.c)
            h = ops.small_linear(e, self.pe[0], self.pe[1], act_out=SL_SILU)
            self._pos_cache[F] = ops.small_linear(h, self.pe[2], self.pe[3])
        return self._pos_cache[F]


@dataclass
class Geom:
    B: int
    F: int
    H: int
    W: int

    @property
    def BF(self):
        return self.B * self.F

    @property
    def HW(self):
        return self.H * self.W

    @property
    def M(self):
        return self.B * self.F * self.H * self.W

    def rv(self, mode):
        return (mode, self.HW, self.F, self.B)

    def down(self):
        return Geom(self.B, self.F, (self.H - 1) // 2 + 1, (self.W - 1) // 2 + 1)

    def up(self):
        return Geom(self.B, self.F, self.H * 2, self.W * 2)


def _mask_ranges(mask: torch.Tensor, batch: int, rows: int):
    """Row ranges [lo, hi) of the samples an adapter mask selects (mask repeat-interleaved over the batch, as the masked
    forward does with the Linear's leading dimension, patch/patch.py:75-76); adjacent samples are merged."""
    if batch % len(mask):
        raise ValueError(f"batch {batch} is not a multiple of the LoRA mask length {len(mask)}")
    m = mask.repeat_interleave(batch // len(mask)).tolist()
    out = []
    for i, v in enumerate(m):
        if v:
            if out and out[-1][1] == i * rows:
                out[-1][1] = (i + 1) * rows
            else:
                out.append([i * rows, (i + 1) * rows])
    return out


def lora_down(x: torch.Tensor, d: Dense, batch: Optional[int]) -> torch.Tensor:
    """t = x A^T for the stacked adapters.  Adapters masked to some samples (patch.hack_lora_forward) contribute zero rows
    elsewhere: their columns of t are computed sample range by sample range into a zeroed buffer."""
    if not d.lora_masks:
        return ops.gemm(x, d.lora_a)
    if batch is None:
        raise ValueError("masked LoRA adapters need the batch size of the activation")
    M_ = x.shape[0]
    rows = M_ // batch
    t = torch.zeros((M_, d.lora_a.shape[0]), device=x.device, dtype=bf16)
    for lo, hi, mask in d.lora_masks:
        for r0, r1 in ([[0, M_]] if mask is None else _mask_ranges(mask, batch, rows)):
            ops.gemm(x[r0:r1], d.lora_a[lo:hi], out=t[r0:r1, lo:hi])
    return t


def dense(x: torch.Tensor, d: Dense, batch: Optional[int] = None, **kw) -> torch.Tensor:
    """x W^T (+ LoRA second segment) through lkgd_gemm.  ``batch``: samples in x (needed only for masked adapters)."""
    if d.lora_a is not None:
        return ops.gemm(x, d.w, bias=d.b, A1=lora_down(x, d, batch), Bw1=d.lora_b, **kw)
    return ops.gemm(x, d.w, bias=d.b, **kw)


class Conditioning:
    """Everything that depends only on (timestep, added_time_ids, context), computed ONCE per forward by three
    batched fp32 mat-vec launches over weights concatenated at pack time, and consumed as GEMM row-vectors /
    LayerNorm add-vectors (column slices of the batched results):
      * every resblock's ``time_emb_proj(SiLU(emb))`` (44 projections of the same input);
      * every KV-length-1 cross-attention term ``to_out(to_v(ctx))``, spatial and temporal."""

    def __init__(self, pk: "PackedUNet", emb: torch.Tensor, ctx: torch.Tensor, ctx_t: Optional[torch.Tensor] = None):
        self.emb = emb          # fp32 [B, 4*C0]  (time + added-time embedding, reference ...controlnet.py:406-419)
        self.ctx = ctx          # fp32 [B, L, D]
        # contexts the TEMPORAL cross-attention indexes (diffusers 0.27.2 picks ctx[row % B_total]): the
        # local batch normally; every CFG half's context when the pair is split across two GPUs
        self.ctx_t = ctx if ctx_t is None else ctx_t
        self.temb_all = ops.small_linear(emb, pk.temb_w, pk.temb_b, act_in=SL_SILU)
        self.xs_all = self.xt_all = None
        self.t_table = False      # xt_all is a [B * n_ctx, C] table (RV_BATCH_TCTX): masked adapters on temporal attn2
        if ctx.shape[1] == 1 and pk.xs_w is not None:
            self.xs_all = ops.small_linear(ctx[:, 0].contiguous(), pk.xs_w, pk.xs_b)
            self.xt_all = ops.small_linear(self.ctx_t[:, 0].contiguous(), pk.xt_w, pk.xt_b)
            if pk.masked_cross:
                self._masked_cross(pk)

    def _masked_cross(self, pk: "PackedUNet"):
        """Per-sample masked LoRA adapters on attn2.to_v / to_out.0: the cross vector of sample b uses the matrix of b's
        adapter pattern.  Spatial (and b-major temporal) rows are per sample anyway; under the 0.27.2 temporal context
        order the row's own sample picks the MATRIX while (row % B) picks the CONTEXT, so the temporal vectors become a
        [B * B, C] table addressed with RV_BATCH_TCTX."""
        B, n_ctx = self.ctx.shape[0], self.ctx_t.shape[0]
        if n_ctx != B:
            raise NotImplementedError("masked cross-attention adapters under a CFG pair split")
        if pk.tctx_mode == RV_TCTX_0272 and any(not sp for sp, _ in pk.masked_cross):
            self.xt_all = self.xt_all.repeat(B, 1)             # row b * B + g = vector of context g (unmasked layers)
            self.t_table = True
        ctx0, ctxt0 = self.ctx[:, 0].contiguous(), self.ctx_t[:, 0].contiguous()
        for spatial, pc in pk.masked_cross:
            cols = slice(pc.off, pc.off + pc.wov.shape[0])
            for key, rows in pc.patterns(B).items():
                w = pc.variant(key)
                if spatial or not self.t_table:
                    idx = torch.tensor(rows, device=ctx0.device)
                    src = (ctx0 if spatial else ctxt0).index_select(0, idx)
                    (self.xs_all if spatial else self.xt_all)[idx, cols] = ops.small_linear(src, w, pc.bo)
                else:
                    v = ops.small_linear(ctxt0, w, pc.bo)                               # [n_ctx, C]: every context
                    for b in rows:
                        self.xt_all[b * B:(b + 1) * B, cols] = v

    def t_rv(self, tctx_mode: int) -> int:
        return RV_BATCH_TCTX if self.t_table and tctx_mode == RV_TCTX_0272 else tctx_mode

    def temb(self, off: int, n: int):       # time_emb_proj(SiLU(emb)) -> [B, Cout] (view)
        return self.temb_all[:, off:off + n]

    def cross_vec(self, pc: PackedCross):   # spatial attn2 term -> [B, C] (view)
        return self.xs_all[:, pc.off:pc.off + pc.wov.shape[0]]

    def cross_vec_t(self, pc: PackedCross):
        return self.xt_all[:, pc.off:pc.off + pc.wov.shape[0]]


def run_resblock(p: PackedResBlock, x: torch.Tensor, skip: Optional[torch.Tensor], g: Geom, cond: Conditioning,
                 want_bf16: bool = False):
    """x (and skip) are fp32 stream tensors; returns the fp32 block output (``want_bf16``: and its bf16 copy, written
    by the same epilogue)."""
    temb_s = cond.temb(p.off_s, p.cout) if p.has_temb else None
    temb_t = cond.temb(p.off_t, p.cout) if p.has_temb else None
    # blocks with a 1x1 shortcut conv need their raw (concatenated) input as a bf16 GEMM operand: the GroupNorm pass
    # that reads it anyway writes that copy too
    xa = None
    if p.wsc is not None:
        h, xa = ops.groupnorm(x, p.n1.g, p.n1.b, p.n1.eps, NS=g.BF, R=g.HW, x2=skip, silu=True, want_raw=True)
    else:
        h = ops.groupnorm(x, p.n1.g, p.n1.b, p.n1.eps, NS=g.BF, R=g.HW, x2=skip, silu=True)
    # gn_rows: the epilogue also accumulates the GroupNorm statistics of what it stores (per frame image, channel), so
    # the GroupNorm that consumes the tensor reads it once instead of twice
    h = ops.gemm(h, p.w1, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=p.b1, rowvec=temb_s, rv=g.rv(RV_BATCH),
                 out_f32=True, gn_rows=g.HW)        # only GroupNorm reads it: keep fp32 instead of rounding twice
    h = ops.groupnorm(h, p.n2.g, p.n2.b, p.n2.eps, NS=g.BF, R=g.HW, silu=True)
    if p.wsc is not None:
        # the 1x1 shortcut conv reads the raw input: narrow (and concatenate) it to a bf16 operand and run it as the
        # SECOND K segment of conv2 (centre tap) - one launch, no fp32 shortcut tensor written and read back
        if xa is None:           # no fused statistics for this input (e.g. ControlNet-injected skips): narrow it here
            xa = ops.cast_bf16(x) if skip is None else ops.concat_channels(x, skip)
        s = ops.gemm(h, p.w2, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=p.b2sc, A1=xa, Bw1=p.wsc, out_f32=True,
                     gn_rows=g.HW)
    else:
        if skip is not None:
            raise ValueError("resblock with concatenated input must have a shortcut conv")
        s = ops.gemm(h, p.w2, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=p.b2, res1=x, out_f32=True, gn_rows=g.HW)
    # temporal half: GroupNorm statistics across frames, (3,1,1) conv over the frame axis, AlphaBlender
    t = ops.groupnorm(s, p.tn1.g, p.tn1.b, p.tn1.eps, NS=g.B, R=g.F * g.HW, silu=True)
    t = ops.gemm(t, p.tw1, mode=A_TCONV3, tconv=(g.B, g.F, g.HW), bias=p.tb1, rowvec=temb_t, rv=g.rv(RV_BATCH),
                 out_f32=True, gn_rows=g.HW)
    t = ops.groupnorm(t, p.tn2.g, p.tn2.b, p.tn2.eps, NS=g.B, R=g.F * g.HW, silu=True)
    # alpha*s + (1-alpha)*(s + conv2(t)) == s + (1-alpha)*conv2(t)
    return ops.gemm(t, p.tw2, mode=A_TCONV3, tconv=(g.B, g.F, g.HW), bias=p.tb2, s0=1.0 - p.alpha, res1=s, s1=1.0,
                    out_f32=True, gn_rows=g.HW, want_bf16=want_bf16)


def _cross_general(pc: PackedCross, n: torch.Tensor, ctx: torch.Tensor, g: Geom, h: torch.Tensor,
                   tctx_0272: bool = False):
    """KV > 1: h + to_out(attention(to_q(n), to_k(ctx), to_v(ctx))); rows of batch b see ctx[b].
    ``tctx_0272``: the temporal block under the diffusers 0.27.2 context order - row m of the temporal batch sees
    ctx[((m / (HW F)) HW + m % HW) % B]: the attention is evaluated against every context (B is the CFG batch,
    the keys are few) and ``select_rows`` keeps each row's own."""
    wq, wkv, wo = pc.general()
    B, L, D = ctx.shape
    c = pc.heads * pc.d
    q = ops.gemm(n, wq)
    kv = ops.gemm(ctx.reshape(B * L, D).to(bf16).contiguous(), wkv)
    if tctx_0272 and B > 1:
        per_ctx = [ops.attention(q, kv[j * L:(j + 1) * L, :c], kv[j * L:(j + 1) * L, c:], n_img=1, heads=pc.heads,
                                 d=pc.d, Nq=g.M, Nk=L) for j in range(B)]
        a = ops.select_rows(per_ctx, (RV_TCTX_0272, g.HW, g.F, B))
    else:
        a = ops.attention(q, kv[:, :c], kv[:, c:], n_img=B, heads=pc.heads, d=pc.d, Nq=g.F * g.HW, Nk=L)
    return ops.gemm(a, wo, bias=pc.bo, res1=h, out_f32=True)


def run_transformer(p: PackedTransformer, x: torch.Tensor, g: Geom, cond: Conditioning, tctx_mode: int,
                    want_bf16: bool = False):
    C = p.c
    kv1 = cond.ctx.shape[1] == 1
    h = ops.groupnorm(x, p.norm.g, p.norm.b, p.norm.eps, NS=g.BF, R=g.HW, silu=False)
    h = dense(h, p.proj_in, batch=g.B, out_f32=True)                                  # fp32 hidden stream
    # ---- spatial block (patch/patch.py:390-580)
    n = ops.layernorm(h, p.s_ln1.g, p.s_ln1.b, p.s_ln1.eps)
    qkv = dense(n, p.s_qkv, batch=g.B)
    a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], n_img=g.BF, heads=p.heads, d=p.d, Nq=g.HW,
                      Nk=g.HW)
    h = dense(a, p.s_out, batch=g.B, res1=h, out_f32=True)
    if p.s_joint is not None:
        # joint attention (patch/patch.py:434-492): queries of sample i against the keys / values of its PARTNER sample,
        # addressed in place (row offsets into the fused projection; with `flip` frame f meets the partner's frame
        # F-1-f); post layer and joint_scale are folded into the output projection, which accumulates onto h in place
        J = p.s_joint
        rows = g.F * g.HW
        qn, kvn = dense(n, J.q, batch=g.B), dense(n, J.kv, batch=g.B)
        an = torch.empty((g.M, C), device=n.device, dtype=bf16)
        for i, pi in enumerate(_partners(J.mask, g.B)):
            if not J.flip:
                ops.attention(qn[i * rows:(i + 1) * rows], kvn[pi * rows:(pi + 1) * rows, :C],
                              kvn[pi * rows:(pi + 1) * rows, C:], n_img=g.F, heads=p.heads, d=p.d, Nq=g.HW, Nk=g.HW,
                              out=an[i * rows:(i + 1) * rows])
            else:
                for f in range(g.F):
                    q0, k0 = i * rows + f * g.HW, pi * rows + (g.F - 1 - f) * g.HW
                    ops.attention(qn[q0:q0 + g.HW], kvn[k0:k0 + g.HW, :C], kvn[k0:k0 + g.HW, C:], n_img=1,
                                  heads=p.heads, d=p.d, Nq=g.HW, Nk=g.HW, out=an[q0:q0 + g.HW])
        if J.fuse is None:
            h = dense(an, J.out, batch=g.B, res1=h, out=h, out_f32=True)
        else:
            o = dense(an, J.out, batch=g.B)                     # plain to_out; the fuse layer sees it next to the partner's
            mm = J.mask.repeat_interleave(g.B // len(J.mask)).tolist()
            for i, pi in enumerate(_partners(J.mask, g.B)):
                w_self, w_cross = J.fuse[bool(mm[i])]
                ops.gemm(o[i * rows:(i + 1) * rows], w_self, A1=o[pi * rows:(pi + 1) * rows], Bw1=w_cross,
                         res1=h[i * rows:(i + 1) * rows], out=h[i * rows:(i + 1) * rows], out_f32=True)
    if kv1:
        n = ops.layernorm(h, p.s_ln3.g, p.s_ln3.b, p.s_ln3.eps, addvec=cond.cross_vec(p.s_cross),
                          rv=g.rv(RV_BATCH), sum_out=h)
    else:
        n = ops.layernorm(h, p.s_ln2.g, p.s_ln2.b, p.s_ln2.eps)
        h = _cross_general(p.s_cross, n, cond.ctx, g, h)
        n = ops.layernorm(h, p.s_ln3.g, p.s_ln3.b, p.s_ln3.eps)
    ff = dense(n, p.s_ff1, act=ACT_GEGLU)
    xs = dense(ff, p.s_ff2, batch=g.B, res1=h, out_f32=True)                                     # x_spatial
    # ---- temporal block (patch/patch.py:582-686) on the same rows, frame stride HW*C
    t0 = torch.empty_like(xs)
    n = ops.layernorm(xs, p.t_lnin.g, p.t_lnin.b, p.t_lnin.eps, addvec=p.pos_emb(g.F), rv=g.rv(RV_FRAMEPOS),
                      sum_out=t0)                                                      # t0 = x_spatial + emb[f]
    ff = dense(n, p.t_ffin1, act=ACT_GEGLU)
    t = dense(ff, p.t_ffin2, batch=g.B, res1=t0, out_f32=True)
    n = ops.layernorm(t, p.t_ln1.g, p.t_ln1.b, p.t_ln1.eps)
    qkv = dense(n, p.t_qkv, batch=g.B)
    a = ops.attention_temporal(qkv, B=g.B, F=g.F, HW=g.HW, heads=p.heads, d=p.d)
    t = dense(a, p.t_out, batch=g.B, res1=t, out_f32=True)
    if p.t_joint is not None:
        # temporal joint attention (patch/patch.py:617-658): pixel (b, p)'s frames attend to the partner sample's frames
        # at the same pixel.  The fused [q | k | v] buffer of the temporal kernel is written directly in partner order:
        # q from every sample's own rows, k / v from the partner's rows (one projection launch per sample, no copy)
        J = p.t_joint
        rows = g.F * g.HW
        buf = torch.empty((g.M, 3 * C), device=n.device, dtype=bf16)
        dense(n, J.q, batch=g.B, out=buf[:, :C])
        if J.kv.lora_a is None:
            for i, pi in enumerate(_partners(J.mask, g.B)):
                ops.gemm(n[pi * rows:(pi + 1) * rows], J.kv.w, out=buf[i * rows:(i + 1) * rows, C:])
        else:       # adapters on to_k / to_v: project once in natural order, then place each sample's rows at its partner
            kvn = dense(n, J.kv, batch=g.B)
            for i, pi in enumerate(_partners(J.mask, g.B)):
                buf[i * rows:(i + 1) * rows, C:].copy_(kvn[pi * rows:(pi + 1) * rows])
        an = ops.attention_temporal(buf, B=g.B, F=g.F, HW=g.HW, heads=p.heads, d=p.d)
        t = dense(an, J.out, batch=g.B, res1=t, out=t, out_f32=True)
    if kv1:
        # under a CFG pair split ctx_t holds every half's context: index it exactly as the unsplit batch would
        n = ops.layernorm(t, p.t_ln3.g, p.t_ln3.b, p.t_ln3.eps, addvec=cond.cross_vec_t(p.t_cross),
                          rv=(cond.t_rv(tctx_mode), g.HW, g.F, cond.ctx_t.shape[0]), sum_out=t)
    else:
        if tctx_mode != RV_BATCH and cond.ctx_t.shape[0] != cond.ctx.shape[0]:
            raise NotImplementedError("temporal cross-attention with KV length > 1 under a CFG pair split")
        n = ops.layernorm(t, p.t_ln2.g, p.t_ln2.b, p.t_ln2.eps)
        t = _cross_general(p.t_cross, n, cond.ctx, g, t, tctx_0272=tctx_mode != RV_BATCH)
        n = ops.layernorm(t, p.t_ln3.g, p.t_ln3.b, p.t_ln3.eps)
    ff = dense(n, p.t_ff1, act=ACT_GEGLU)
    # AlphaBlender: alpha*x_spatial + (1-alpha)*(ff_out + t); bf16 because proj_out reads it as its GEMM operand
    mix = dense(ff, p.t_ff2, batch=g.B, s0=1.0 - p.alpha, res1=t, s1=1.0 - p.alpha, res2=xs, s2=p.alpha)
    # the next resblock's GroupNorm consumes this tensor: fused statistics where a 128-row tile stays inside a frame
    return dense(mix, p.proj_out, batch=g.B, res1=x, out_f32=True, gn_rows=g.HW if g.HW % 128 == 0 else 0,
                 want_bf16=want_bf16)


class PackedUNet:
    """Kernel-ready copy of a UNet's weights + the forward schedule (down / mid / up, skip handling)."""

    def __init__(self, unet, fold_lora: bool = True):
        cfg = unet.config
        self.cfg = cfg
        self.c0 = cfg.block_out_channels[0]
        self.cin = cfg.in_channels
        self.cin_pad = 64
        # a model may present a merged stem (the flow variant folds its second, gated conv_in into one 12-channel conv)
        stem = unet._stem_conv() if hasattr(unet, "_stem_conv") else unet.conv_in
        self.conv_in_w, self.conv_in_b = _conv3x3_weight(stem, cin_pad=self.cin_pad)
        te, ae = unet.time_embedding, unet.add_embedding
        self.te = tuple(_f32(t) for t in (te.linear_1.weight, te.linear_1.bias, te.linear_2.weight, te.linear_2.bias))
        self.ae = tuple(_f32(t) for t in (ae.linear_1.weight, ae.linear_1.bias, ae.linear_2.weight, ae.linear_2.bias))
        self.down = []
        for blk in unet.down_blocks:
            res = [PackedResBlock(r) for r in blk.resnets]
            att = [PackedTransformer(a, fold_lora) for a in blk.attentions] if blk.has_cross_attention else None
            ds = _conv3x3_weight(blk.downsamplers[0].conv) if blk.downsamplers is not None else None
            self.down.append((res, att, ds))
        mb = unet.mid_block
        self.mid = ([PackedResBlock(r) for r in mb.resnets], [PackedTransformer(a, fold_lora) for a in mb.attentions])
        self.up = []
        for blk in getattr(unet, "up_blocks", []):
            res = [PackedResBlock(r) for r in blk.resnets]
            att = [PackedTransformer(a, fold_lora) for a in blk.attentions] if blk.has_cross_attention else None
            us = _conv3x3_weight(blk.upsamplers[0].conv) if blk.upsamplers is not None else None
            self.up.append((res, att, us))
        if hasattr(unet, "conv_out"):
            self.norm_out = Norm.of(unet.conv_norm_out)
            self.cout = cfg.out_channels
            self.conv_out_w, self.conv_out_b = _conv3x3_weight(unet.conv_out, cout_pad=max(32, (self.cout + 15) // 16 * 16))
        self._batch_conditioning()
        order = getattr(cfg, "time_context_order", "hw_major_0272")
        if order not in ("hw_major_0272", "b_major"):
            raise ValueError(f"time_context_order must be 'hw_major_0272' or 'b_major', got {order}")
        self.tctx_mode = RV_TCTX_0272 if order == "hw_major_0272" else RV_BATCH

    def _batch_conditioning(self):
        """Concatenates every time_emb_proj and every fused KV=1 cross-attention matrix so that one launch each
        serves the whole forward; the per-block tensors become views of the concatenation."""
        res = [r for blk in self.down for r in blk[0]] + list(self.mid[0]) + [r for blk in self.up for r in blk[0]]
        tr = [a for blk in self.down if blk[1] for a in blk[1]] + list(self.mid[1]) + \
             [a for blk in self.up if blk[1] for a in blk[1]]
        ws, bs, off = [], [], 0
        for r in res:
            r.off_s, r.off_t = off, off + r.cout
            ws += [r.temb_w, r.ttemb_w]
            bs += [r.temb_b, r.ttemb_b]
            off += 2 * r.cout
            r.temb_w = r.temb_b = r.ttemb_w = r.ttemb_b = None
        self.temb_w, self.temb_b = torch.cat(ws, 0).contiguous(), torch.cat(bs, 0).contiguous()
        self.xs_w = self.xs_b = self.xt_w = self.xt_b = None
        self.masked_cross = [(key == "s_cross", getattr(t, key)) for t in tr for key in ("s_cross", "t_cross")
                             if getattr(t, key).masked]
        if tr:
            for name, key in (("xs", "s_cross"), ("xt", "t_cross")):
                off, w, b = 0, [], []
                for t in tr:
                    pc = getattr(t, key)
                    pc.off = off
                    w.append(pc.wov)
                    b.append(pc.bo)
                    off += pc.wov.shape[0]
                setattr(self, name + "_w", torch.cat(w, 0).contiguous())
                setattr(self, name + "_b", torch.cat(b, 0).contiguous())
                for t in tr:                     # keep only the shape; the matrix lives in the concatenation
                    pc = getattr(t, key)
                    pc.wov = getattr(self, name + "_w")[pc.off:pc.off + pc.wov.shape[0]]

    # ------------------------------------------------------------------------------------------
    def time_embedding(self, timestep: torch.Tensor, added_time_ids: torch.Tensor, te=None, ae=None) -> torch.Tensor:
        """emb = time_embedding(Timesteps(t)) + add_embedding(Timesteps(added_time_ids).reshape(B,-1)) in fp32.
        ``te`` / ``ae``: another head's MLP weights (the joint UNet's y heads)."""
        te, ae = te or self.te, ae or self.ae
        B = added_time_ids.shape[0]
        t = timestep.to(torch.float32).reshape(-1).expand(B).contiguous()
        e = ops.timestep_embedding(t, self.c0)
        e = ops.small_linear(ops.small_linear(e, te[0], te[1], act_out=SL_SILU), te[2], te[3])
        a = ops.timestep_embedding(added_time_ids.to(torch.float32).reshape(-1), self.cfg.addition_time_embed_dim)
        a = a.reshape(B, -1)
        a = ops.small_linear(ops.small_linear(a, ae[0], ae[1], act_out=SL_SILU), ae[2], ae[3])
        ops.axpy_f32(a, e)
        return e

    def encoder(self, x: torch.Tensor, g: Geom, cond: Conditioning, stem_add: Optional[torch.Tensor] = None,
                bf16_skips: bool = False, stem: Optional[torch.Tensor] = None):
        """conv_in (+ ControlNet condition embedding) -> down blocks -> mid.  Returns (sample, skips, geoms, geometry of
        the mid block).  ``bf16_skips``: every skip tensor and the mid output are ``(fp32, bf16 copy)`` pairs - the copy
        is written by the epilogue that produces the tensor and feeds the ControlNet's zero convs without a narrowing pass
        (the downsampler also reads it instead of narrowing its input itself)."""
        wb = bf16_skips

        def both(t):                 # (fp32, bf16) -> fp32 stream tensor, bf16 copy (None when not requested)
            return t if wb else (t, None)

        if stem is not None:         # conv_in already applied by the caller (per-sample input heads of the joint UNet)
            if wb or stem_add is not None:
                raise ValueError("a precomputed stem excludes stem_add / bf16_skips")
            x, xb = stem, None
        else:
            x, xb = both(ops.gemm(x, self.conv_in_w, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=self.conv_in_b,
                                  res1=stem_add, out_f32=True, gn_rows=g.HW, want_bf16=wb))
        skips, geoms = [(x, xb) if wb else x], [g]
        for res, att, ds in self.down:
            for i, r in enumerate(res):
                if att is not None:
                    x = run_resblock(r, x, None, g, cond)
                    x, xb = both(run_transformer(att[i], x, g, cond, self.tctx_mode, want_bf16=wb))
                else:
                    x, xb = both(run_resblock(r, x, None, g, cond, want_bf16=wb))
                skips.append((x, xb) if wb else x)
                geoms.append(g)
            if ds is not None:
                x, xb = both(ops.gemm(xb if xb is not None else ops.cast_bf16(x), ds[0], mode=A_CONV3X3,
                                      conv=(g.BF, g.H, g.W, 2), bias=ds[1], out_f32=True, gn_rows=g.down().HW,
                                      want_bf16=wb))
                g = g.down()
                skips.append((x, xb) if wb else x)
                geoms.append(g)
        res, att = self.mid
        x = run_resblock(res[0], x, None, g, cond)
        for j, (a, r) in enumerate(zip(att, res[1:])):
            x = run_transformer(a, x, g, cond, self.tctx_mode)
            x = run_resblock(r, x, None, g, cond, want_bf16=wb and j == len(att) - 1)
        return x, skips, geoms, g

    def decoder(self, x: torch.Tensor, skips: List[torch.Tensor], g: Geom, cond: Conditioning) -> torch.Tensor:
        for res, att, us in self.up:
            for i, r in enumerate(res):
                x = run_resblock(r, x, skips.pop(), g, cond)
                if att is not None:
                    x = run_transformer(att[i], x, g, cond, self.tctx_mode)
            if us is not None:
                x = ops.upsample2x(x, g.BF, g.H, g.W)
                g = g.up()
                x = ops.gemm(x, us[0], mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=us[1], out_f32=True, gn_rows=g.HW)
        h = ops.groupnorm(x, self.norm_out.g, self.norm_out.b, self.norm_out.eps, NS=g.BF, R=g.HW, silu=True)
        # conv_out: N padded to 32 rows of zeros, only the first `cout` columns are stored (fp32)
        return ops.gemm(h, self.conv_out_w, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=self.conv_out_b,
                        out_f32=True, n_store=self.cout)


def residual_multipliers(n_down_blocks: int, skips_per_block: Sequence[int]) -> List[int]:
    """Reference quirk F6: the ControlNet residual add sits inside the down-block loop and ``zip`` truncates
    (models/unet_spatio_temporal_condition_controlnet.py:453-462), so a skip produced by down block j (conv_in
    counts with block 0) ends up with ``n_down_blocks - j`` times its residual: (4,4,4,4,3,3,3,2,2,2,1,1)."""
    mult = []
    for j, n in enumerate(skips_per_block):
        mult += [n_down_blocks - j] * n
    return mult
