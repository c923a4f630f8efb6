"""LoRA fine-tuning step of the SVD / LKGD UNet on the lkgd_b200 kernels: forward with saved activations, a
hand-scheduled backward through the block graph, LoRA weight gradients, fused clip + AdamW, one flat all-reduce.

Mirrors one iteration of the reference training loop ``train_models/train_svd_lora.py:1445-1689``:
EDM noising + input preconditioning (:1503-1530), ``unet(inp_noisy_latents, timesteps, encoder_hidden_states,
[domain_features, flow_features,] added_time_ids=...)`` (:1634-1642), the v-prediction wrapper and weighted MSE
(:1651-1672), ``accelerator.backward`` (:1683), ``clip_grad_norm_`` (:1684-1686), ``optimizer.step`` (:1687), and the
DDP gradient all-reduce that ``accelerator.prepare`` installs (:1300-1302).  Trainable parameters are the LoRA pairs the
reference's adapter configurations create: ``temporal_transformer_blocks.*.attn1.to_{q,k,v}`` (:1081-1088, the default) or
every attention projection ``to_q / to_k / to_v / to_out.0`` of the spatial and temporal blocks, attn1 and attn2 (:1091-1096,
run_models/run_inference_flow_lora.py:326-331); every other weight is frozen, so the backward only propagates data gradients
(GEMMs / convolutions re-run with transposed, tap-flipped weights on the same tcgen05 kernel) plus the two skinny
weight-gradient GEMMs per LoRA pair; the KV-length-1 cross-attention adapters get theirs from the gradient of the
per-sample cross vectors.

There is no autograd here and no PyTorch compute: the backward schedule is written out per block, the gradient of the
residual stream is kept in fp32 like the forward stream, GEMM operands are bf16.
"""
from __future__ import annotations

from types import SimpleNamespace as NS
from typing import Dict, List, Optional, Tuple

import torch

from . import modules as M
from . import ops
from .engine import (Conditioning, Geom, PackedResBlock, PackedTransformer, PackedUNet, dense)
from .ops import A_CONV3X3, A_TCONV3, RV_BATCH, RV_FRAMEPOS, bf16


# ------------------------------------------------------------------------------------------------- data-gradient weights
def _t(w: torch.Tensor) -> torch.Tensor:
    return w.t().contiguous()


def _conv_dgrad(w: torch.Tensor, taps: int) -> torch.Tensor:
    """[Co, taps*Ci] ([Co, tap, Ci]) -> [Ci, taps*Co] with the taps reversed: the data gradient of a stride-1 'same'
    convolution is the convolution of the output gradient with the flipped, channel-swapped kernel."""
    co = w.shape[0]
    ci = w.shape[1] // taps
    return w.reshape(co, taps, ci).flip(1).permute(2, 1, 0).reshape(ci, taps * co).contiguous()


def res_train_weights(p: PackedResBlock) -> NS:
    if getattr(p, "_tw", None) is None:
        p._tw = NS(w1=_conv_dgrad(p.w1, 9), w2=_conv_dgrad(p.w2, 9), tw1=_conv_dgrad(p.tw1, 3),
                   tw2=_conv_dgrad(p.tw2, 3), wsc=None if p.wsc is None else _t(p.wsc))
    return p._tw


SITES = ("s_qkv", "s_out", "t_qkv", "t_out")     # projections whose LoRA adapters can be trained (self-attention q|k|v, out)


def site_modules(p: PackedTransformer, name: str):
    """The nn modules behind a packed projection, in the row order of its weight."""
    blk = (p.src.transformer_blocks if name[0] == "s" else p.src.temporal_transformer_blocks)[0]
    a = blk.attn1
    return (a.to_q, a.to_k, a.to_v) if name.endswith("qkv") else (a.to_out[0],)


def tr_train_weights(p: PackedTransformer) -> NS:
    if getattr(p, "_tw", None) is None:
        if p.s_joint is not None or p.t_joint is not None:
            raise NotImplementedError("training through the joint-attention branch is not built (inference only)")
        for d in (p.proj_in, p.proj_out, p.s_ff2, p.t_ffin2, p.t_ff2, p.s_ff1, p.t_ffin1, p.t_ff1):
            if d.lora_a is not None:
                raise NotImplementedError("training supports LoRA on the attention projections to_q / to_k / to_v / "
                                          "to_out.0 (the reference's adapter configs, train_svd_lora.py:1081-1096), not on "
                                          "proj_in / proj_out / the feed-forward layers")
        p._tw = NS(proj_in=_t(p.proj_in.w), proj_out=_t(p.proj_out.w), s_qkv=_t(p.s_qkv.w), s_out=_t(p.s_out.w),
                   s_ff1=_t(p.s_ff1.w), s_ff2=_t(p.s_ff2.w), t_ffin1=_t(p.t_ffin1.w), t_ffin2=_t(p.t_ffin2.w),
                   t_qkv=_t(p.t_qkv.w), t_out=_t(p.t_out.w), t_ff1=_t(p.t_ff1.w), t_ff2=_t(p.t_ff2.w),
                   lora={n: (_t(getattr(p, n).lora_a), _t(getattr(p, n).lora_b)) for n in SITES
                         if getattr(p, n).lora_a is not None})
    return p._tw


def _lora_fwd(S, name: str, x, d, **kw):
    """y = x W^T + b [+ (x A^T)(s B)^T]; keeps the projection's input and the down-projected rows for the weight gradients."""
    if d.lora_a is None:
        return ops.gemm(x, d.w, bias=d.b, **kw)
    t = ops.gemm(x, d.lora_a)
    setattr(S, name + "_x", x)
    setattr(S, name + "_t", t)
    return ops.gemm(x, d.w, bias=d.b, A1=t, Bw1=d.lora_b, **kw)


def _lora_bwd(S, name: str, dy, d, wT, tw, grads):
    """bf16 gradient of a projection's output -> bf16 gradient of its input; accumulates the adapters' fp32 weight
    gradients: dB_i += s dy_i^T t_i, dA_i += (dy_i s B_i)^T x."""
    if d.lora_a is None:
        return ops.gemm(dy, wT)
    aT, bT = tw.lora[name]
    dt = ops.gemm(dy, bT)                                         # [M, n_adapters * r_pad]
    if grads is not None:
        x, t = getattr(S, name + "_x"), getattr(S, name + "_t")
        for ad in grads.get(name, ()):
            rs = slice(ad["r_lo"], ad["r_lo"] + ad["r"])
            ops.gemm_tn(dy[:, ad["n_lo"]:ad["n_hi"]], t[:, rs], ad["B"], alpha=ad["scaling"])
            ops.gemm_tn(dt[:, rs], x, ad["A"])
    return ops.gemm(dy, wT, A1=dt, Bw1=aT)


# ------------------------------------------------------------------------------------------------- resblock
def resblock_fwd(p: PackedResBlock, x, skip, g: Geom, cond: Conditioning):
    """engine.run_resblock with the tensors the backward needs kept (GroupNorm inputs + statistics)."""
    S = NS(x=x, skip=skip)
    h, S.st1 = ops.groupnorm(x, p.n1.g, p.n1.b, p.n1.eps, NS=g.BF, R=g.HW, x2=skip, silu=True, return_stats=True)
    # gn_rows: the producing GEMM's epilogue accumulates the GroupNorm statistics of what it stores (as in inference); the
    # fused GroupNorm leaves the same per-(sample, channel) sums behind that the backward re-reads
    S.c1 = ops.gemm(h, p.w1, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=p.b1, rowvec=cond.temb(p.off_s, p.cout),
                    rv=g.rv(RV_BATCH), out_f32=True, gn_rows=g.HW)
    h, S.st2 = ops.groupnorm(S.c1, p.n2.g, p.n2.b, p.n2.eps, NS=g.BF, R=g.HW, silu=True, return_stats=True)
    if p.wsc is not None:
        xa = ops.cast_bf16(x) if skip is None else ops.concat_channels(x, skip)
        sc = ops.gemm(xa, p.wsc, bias=p.bsc, out_f32=True)
    else:
        if skip is not None:
            raise ValueError("resblock with concatenated input must have a shortcut conv")
        sc = x
    S.s = ops.gemm(h, p.w2, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=p.b2, res1=sc, out_f32=True, gn_rows=g.HW)
    t, S.st3 = ops.groupnorm(S.s, p.tn1.g, p.tn1.b, p.tn1.eps, NS=g.B, R=g.F * g.HW, silu=True, return_stats=True)
    S.c3 = ops.gemm(t, p.tw1, mode=A_TCONV3, tconv=(g.B, g.F, g.HW), bias=p.tb1, rowvec=cond.temb(p.off_t, p.cout),
                    rv=g.rv(RV_BATCH), out_f32=True, gn_rows=g.HW)
    t, S.st4 = ops.groupnorm(S.c3, p.tn2.g, p.tn2.b, p.tn2.eps, NS=g.B, R=g.F * g.HW, silu=True, return_stats=True)
    out = ops.gemm(t, p.tw2, mode=A_TCONV3, tconv=(g.B, g.F, g.HW), bias=p.tb2, s0=1.0 - p.alpha, res1=S.s, s1=1.0,
                   out_f32=True, gn_rows=g.HW)
    return out, S


def resblock_bwd(p: PackedResBlock, S, G: torch.Tensor, g: Geom):
    """G: fp32 gradient of the block output (consumed: overwritten with the gradient of ``s``).  Returns
    (dx fp32 [M, C1], dskip fp32 [M, C2] | None)."""
    tw = res_train_weights(p)
    M_, C = G.shape
    tc = dict(mode=A_TCONV3, tconv=(g.B, g.F, g.HW))
    cv = dict(mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1))
    # out = s + (1 - alpha) * (tconv2(t2) + b)
    gb = torch.empty((M_, C), device=G.device, dtype=bf16)
    ops.cast2d_bf16(G, gb, alpha=1.0 - p.alpha)
    d = ops.gemm(gb, tw.tw2, **tc)                                           # d t2
    ops.groupnorm_bwd(S.c3, d, S.st4, p.tn2.g, p.tn2.b, p.tn2.eps, NS=g.B, R=g.F * g.HW, silu=True, out_bf16=gb)
    d = ops.gemm(gb, tw.tw1, **tc)                                           # d t1
    ops.groupnorm_bwd(S.s, d, S.st3, p.tn1.g, p.tn1.b, p.tn1.eps, NS=g.B, R=g.F * g.HW, silu=True, out1=G, acc1=True,
                      out_bf16=gb)                                           # G = ds (fp32), gb = bf16(ds)
    d = ops.gemm(gb, tw.w2, **cv)                                            # d h2
    dc1 = torch.empty_like(d)
    ops.groupnorm_bwd(S.c1, d, S.st2, p.n2.g, p.n2.b, p.n2.eps, NS=g.BF, R=g.HW, silu=True, out_bf16=dc1)
    d = ops.gemm(dc1, tw.w1, **cv)                                           # d h1 [M, C1 + C2]
    add = ops.gemm(gb, tw.wsc) if tw.wsc is not None else G                  # shortcut path
    c1 = S.x.shape[1]
    dx = torch.empty((M_, c1), device=G.device, dtype=torch.float32)
    dskip = None if S.skip is None else torch.empty((M_, S.skip.shape[1]), device=G.device, dtype=torch.float32)
    ops.groupnorm_bwd(S.x, d, S.st1, p.n1.g, p.n1.b, p.n1.eps, NS=g.BF, R=g.HW, x2=S.skip, silu=True, add=add,
                      out1=dx, out2=dskip)
    return dx, dskip


# ------------------------------------------------------------------------------------------------- transformer
def _ff_fwd(n, d1, d2, **kw):
    pre = ops.gemm(n, d1.w, bias=d1.b)             # tile-interleaved [M, 8C]; GEGLU applied by its own kernel so the
    return pre, dense(ops.geglu_fwd(pre), d2, **kw)  # backward can re-read the pre-activation


def transformer_fwd(p: PackedTransformer, x, g: Geom, cond: Conditioning, tctx_mode: int):
    if cond.ctx.shape[1] != 1:
        raise NotImplementedError("training covers the KV-length-1 context the reference always uses (SURVEY F7)")
    C = p.c
    S = NS(x=x)
    h, S.st0 = ops.groupnorm(x, p.norm.g, p.norm.b, p.norm.eps, NS=g.BF, R=g.HW, silu=False, return_stats=True)
    S.h_a = dense(h, p.proj_in, out_f32=True)
    n = ops.layernorm(S.h_a, p.s_ln1.g, p.s_ln1.b, p.s_ln1.eps)
    S.s_qkv = _lora_fwd(S, "s_qkv", n, p.s_qkv)
    S.s_o, S.s_lse = ops.attention(S.s_qkv[:, :C], S.s_qkv[:, C:2 * C], S.s_qkv[:, 2 * C:], n_img=g.BF, heads=p.heads,
                                   d=p.d, Nq=g.HW, Nk=g.HW, return_lse=True)
    S.h_b = _lora_fwd(S, "s_out", S.s_o, p.s_out, res1=S.h_a, out_f32=True)
    n = ops.layernorm(S.h_b, p.s_ln3.g, p.s_ln3.b, p.s_ln3.eps, addvec=cond.cross_vec(p.s_cross), rv=g.rv(RV_BATCH),
                      sum_out=S.h_b)
    S.s_pre, S.xs = _ff_fwd(n, p.s_ff1, p.s_ff2, res1=S.h_b, out_f32=True)
    S.t0 = torch.empty_like(S.xs)
    n = ops.layernorm(S.xs, p.t_lnin.g, p.t_lnin.b, p.t_lnin.eps, addvec=p.pos_emb(g.F), rv=g.rv(RV_FRAMEPOS),
                      sum_out=S.t0)
    S.t_pre_in, S.t_a = _ff_fwd(n, p.t_ffin1, p.t_ffin2, res1=S.t0, out_f32=True)
    n = ops.layernorm(S.t_a, p.t_ln1.g, p.t_ln1.b, p.t_ln1.eps)
    S.t_qkv = _lora_fwd(S, "t_qkv", n, p.t_qkv)
    a = ops.attention_temporal(S.t_qkv, B=g.B, F=g.F, HW=g.HW, heads=p.heads, d=p.d)
    S.t_b = _lora_fwd(S, "t_out", a, p.t_out, res1=S.t_a, out_f32=True)
    S.n_tctx = cond.ctx_t.shape[0]
    n = ops.layernorm(S.t_b, p.t_ln3.g, p.t_ln3.b, p.t_ln3.eps, addvec=cond.cross_vec_t(p.t_cross),
                      rv=(tctx_mode, g.HW, g.F, S.n_tctx), sum_out=S.t_b)
    S.t_pre, mix = _ff_fwd(n, p.t_ff1, p.t_ff2, s0=1.0 - p.alpha, res1=S.t_b, s1=1.0 - p.alpha, res2=S.xs, s2=p.alpha)
    out = dense(mix, p.proj_out, res1=x, out_f32=True, gn_rows=g.HW if g.HW % 128 == 0 else 0)
    return out, S


def _ff_bwd(gb, pre, w2T, w1T):
    """bf16 gradient of the feed-forward output -> bf16 gradient of its (LayerNorm-ed) input."""
    return ops.gemm(ops.geglu_bwd(pre, ops.gemm(gb, w2T)), w1T)


def transformer_bwd(p: PackedTransformer, S, G: torch.Tensor, g: Geom, tctx_mode: int, lora_grads=None,
                    cross_grads=None):
    """G: fp32 gradient of the transformer output; overwritten with the gradient of its input ``x`` and returned.
    ``lora_grads``: {site name: [adapter dicts with the fp32 ``A`` [r, K] / ``B`` [N_i, r] accumulators, the adapter's row
    range in the projection and its column range in the stacked down-projection, ``scaling``]} of this layer.  ``cross_grads``: (d_xs [B, C] view, d_xt [n_ctx, C] view) accumulators of the two KV-length-1
    cross-attention vectors (LKGD conditioning gradient)."""
    tw = tr_train_weights(p)
    M_, C = G.shape
    dev = G.device
    gb = ops.cast_bf16(G)
    dmix = ops.gemm(gb, tw.proj_out, out_f32=True)                # mix = (1-a)(ff + t_b) + a xs
    Gt = ops.scale_f32(dmix, 1.0 - p.alpha)
    ops.cast2d_bf16(dmix, gb, alpha=1.0 - p.alpha)
    # ---- temporal block, last to first
    ops.layernorm_bwd(S.t_b, _ff_bwd(gb, S.t_pre, tw.t_ff2, tw.t_ff1), p.t_ln3.g, p.t_ln3.eps, Gt, g_bf16=gb)
    if cross_grads is not None:
        ops.colsum_grouped(Gt, S.n_tctx, (tctx_mode, g.HW, g.F, S.n_tctx), out=cross_grads[1])
    da = _lora_bwd(S, "t_out", gb, p.t_out, tw.t_out, tw, lora_grads)
    dqkv = ops.attention_temporal_bwd(S.t_qkv, da, B=g.B, F=g.F, HW=g.HW, heads=p.heads, d=p.d)
    dn = _lora_bwd(S, "t_qkv", dqkv, p.t_qkv, tw.t_qkv, tw, lora_grads)
    ops.layernorm_bwd(S.t_a, dn, p.t_ln1.g, p.t_ln1.eps, Gt, g_bf16=gb)
    ops.layernorm_bwd(S.t0, _ff_bwd(gb, S.t_pre_in, tw.t_ffin2, tw.t_ffin1), p.t_lnin.g, p.t_lnin.eps, Gt)
    # ---- spatial block: d xs = d t0 + alpha * d mix
    ops.axpby(dmix, p.alpha, Gt, 1.0)
    Gs = Gt
    ops.cast2d_bf16(Gs, gb)
    ops.layernorm_bwd(S.h_b, _ff_bwd(gb, S.s_pre, tw.s_ff2, tw.s_ff1), p.s_ln3.g, p.s_ln3.eps, Gs, g_bf16=gb)
    if cross_grads is not None:
        ops.colsum_grouped(Gs, g.B, g.rv(RV_BATCH), out=cross_grads[0])
    da = _lora_bwd(S, "s_out", gb, p.s_out, tw.s_out, tw, lora_grads)
    dqkv = torch.empty((M_, 3 * C), device=dev, dtype=bf16)
    ops.attention_bwd(S.s_qkv[:, :C], S.s_qkv[:, C:2 * C], S.s_qkv[:, 2 * C:], S.s_o, da, S.s_lse, dqkv[:, :C],
                      dqkv[:, C:2 * C], dqkv[:, 2 * C:], n_img=g.BF, heads=p.heads, d=p.d, N=g.HW)
    ops.layernorm_bwd(S.h_a, _lora_bwd(S, "s_qkv", dqkv, p.s_qkv, tw.s_qkv, tw, lora_grads), p.s_ln1.g, p.s_ln1.eps, Gs,
                      g_bf16=gb)
    dh0 = ops.gemm(gb, tw.proj_in)
    ops.groupnorm_bwd(S.x, dh0, S.st0, p.norm.g, p.norm.b, p.norm.eps, NS=g.BF, R=g.HW, silu=False, out1=G, acc1=True)
    return G


# ------------------------------------------------------------------------------------------------- whole UNet
class TrainGraph:
    """Forward-with-save / backward of a PackedUNet.  One instance per step (holds the saved activations)."""

    def __init__(self, pk: PackedUNet):
        self.pk = pk
        self.tape: List[Tuple] = []

    # ---- forward: PackedUNet.encoder + decoder with a tape
    def forward(self, x: torch.Tensor, g: Geom, cond: Conditioning) -> torch.Tensor:
        pk, tape = self.pk, self.tape
        tm = pk.tctx_mode
        x = ops.gemm(x, pk.conv_in_w, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=pk.conv_in_b, out_f32=True,
                     gn_rows=g.HW)
        skips = [x]
        tape.append(("skip", 0))
        for res, att, ds in pk.down:
            for i, r in enumerate(res):
                x, S = resblock_fwd(r, x, None, g, cond)
                tape.append(("res", r, S, g, None))
                if att is not None:
                    x, S = transformer_fwd(att[i], x, g, cond, tm)
                    tape.append(("tr", att[i], S, g))
                tape.append(("skip", len(skips)))
                skips.append(x)
            if ds is not None:
                gin = g
                x = ops.gemm(ops.cast_bf16(x), ds[0], mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 2), bias=ds[1], out_f32=True,
                             gn_rows=g.down().HW)
                g = g.down()
                tape.append(("down", ds, gin))
                tape.append(("skip", len(skips)))
                skips.append(x)
        res, att = pk.mid
        x, S = resblock_fwd(res[0], x, None, g, cond)
        tape.append(("res", res[0], S, g, None))
        for a, r in zip(att, res[1:]):
            x, S = transformer_fwd(a, x, g, cond, tm)
            tape.append(("tr", a, S, g))
            x, S = resblock_fwd(r, x, None, g, cond)
            tape.append(("res", r, S, g, None))
        for res, att, us in pk.up:
            for i, r in enumerate(res):
                k = len(skips) - 1
                x, S = resblock_fwd(r, x, skips.pop(), g, cond)
                tape.append(("res", r, S, g, k))
                if att is not None:
                    x, S = transformer_fwd(att[i], x, g, cond, tm)
                    tape.append(("tr", att[i], S, g))
            if us is not None:
                gin = g
                x = ops.upsample2x(x, g.BF, g.H, g.W)
                g = g.up()
                x = ops.gemm(x, us[0], mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=us[1], out_f32=True, gn_rows=g.HW)
                tape.append(("up", us, gin, g))
        h, st = ops.groupnorm(x, pk.norm_out.g, pk.norm_out.b, pk.norm_out.eps, NS=g.BF, R=g.HW, silu=True,
                              return_stats=True)
        tape.append(("out", x, st, g))
        return ops.gemm(h, pk.conv_out_w, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1), bias=pk.conv_out_b, out_f32=True,
                        n_store=pk.cout)

    # ---- backward
    def backward(self, dpred: torch.Tensor, lora_grads: Optional[Dict] = None, cross_grads=None):
        """dpred: bf16 rows [M, conv_out_w.shape[0]] (gradient of the prediction, zero padded).  Walks the tape in
        reverse; stops below the first (in forward order) block that owns a trainable parameter."""
        pk = self.pk
        tm = pk.tctx_mode
        tw = getattr(pk, "_tw", None)
        if tw is None:
            tw = pk._tw = NS(conv_out=_conv_dgrad(pk.conv_out_w, 9),
                             ups={id(b[2]): _conv_dgrad(b[2][0], 9) for b in pk.up if b[2] is not None},
                             downs={id(b[2]): _conv_dgrad(b[2][0], 9) for b in pk.down if b[2] is not None})
        first_tr = next(i for i, e in enumerate(self.tape) if e[0] == "tr")
        dskips: Dict[int, torch.Tensor] = {}
        G = None
        for idx in range(len(self.tape) - 1, first_tr - 1, -1):
            e = self.tape[idx]
            kind = e[0]
            if kind == "out":
                _, x, st, g = e
                dh = ops.gemm(dpred, tw.conv_out, mode=A_CONV3X3, conv=(g.BF, g.H, g.W, 1))
                G = torch.empty_like(x)
                ops.groupnorm_bwd(x, dh, st, pk.norm_out.g, pk.norm_out.b, pk.norm_out.eps, NS=g.BF, R=g.HW, silu=True,
                                  out1=G)
            elif kind == "up":
                _, us, gin, gout = e
                d = ops.gemm(ops.cast_bf16(G), tw.ups[id(us)], mode=A_CONV3X3, conv=(gout.BF, gout.H, gout.W, 1))
                G = ops.downsum2x(d, gin.BF, gin.H, gin.W)
            elif kind == "down":
                _, ds, gin = e
                z = ops.zero_stuff2x(G, gin.BF, gin.H, gin.W)
                G = ops.gemm(z, tw.downs[id(ds)], mode=A_CONV3X3, conv=(gin.BF, gin.H, gin.W, 1), out_f32=True)
            elif kind == "skip":
                k = e[1]
                if k in dskips:
                    ops.axpby(dskips.pop(k), 1.0, G, 1.0)
            elif kind == "tr":
                _, p, S, g = e
                lg = None if lora_grads is None else lora_grads.get(id(p))
                cg = None if cross_grads is None else cross_grads(p)
                G = transformer_bwd(p, S, G, g, tm, lg, cg)
            elif kind == "res":
                _, p, S, g, k = e
                G, dskip = resblock_bwd(p, S, G, g)
                if dskip is not None:
                    dskips[k] = dskip
        self.tape = []
        return G


# ------------------------------------------------------------------------------------------------- trainer
class LoraTrainer:
    """Owns the trainable state of a LoRA fine-tuning run: ONE flat fp32 parameter buffer (the module's LoRA
    parameters become views of it), flat gradient / Adam moment buffers, the bf16 GEMM operands derived from the
    parameters, and the step: forward -> loss -> backward -> [all-reduce] -> clip -> AdamW -> repack.

    ``world_size`` > 1 (+ optional ``group``, default group when None): the flat gradient is summed across ranks with one
    ``all_reduce`` per step and averaged inside the optimizer kernel - the data-parallel gradient sync the reference
    gets from DDP (train_svd_lora.py:1300-1302)."""

    def __init__(self, unet, lr: float = 1e-4, betas=(0.9, 0.999), weight_decay: float = 1e-2, eps: float = 1e-8,
                 max_grad_norm: float = 1.0, group=None, world_size: int = 1):
        self.unet, self.group, self.world = unet, group, world_size
        self.lr, self.betas, self.wd, self.eps, self.max_norm = lr, betas, weight_decay, eps, max_grad_norm
        self.step_count = 0
        self._repack_batch = None
        if not getattr(unet, "fold_lora", True):
            raise ValueError("training needs the LoRA pairs folded into the GEMM (fold_lora=True)")
        pk = unet.packed()
        self._pk = pk            # the trainer owns this pack: gradient slots are keyed by its layers (see forward_backward)
        self.layers = [a for blk in pk.down if blk[1] for a in blk[1]] + list(pk.mid[1]) + \
                      [a for blk in pk.up if blk[1] for a in blk[1]]
        dev = unet.device

        def trainable(m):
            if not isinstance(m, M.LoraLinear) or m.merged:
                return False
            if len(m.lora_A) != 1 or m.masked_forward:
                raise NotImplementedError("training supports ONE unmasked adapter per layer")
            return True

        # ---- what is trainable: (layer, site, adapter geometry) for the self-attention projections and the KV-length-1
        # cross-attention projections (to_q / to_k of a one-key attention receive exactly zero gradient, but they are
        # optimizer state - weight decay - in the reference, so they are kept)
        plan, n = [], 0
        for p in self.layers:
            tr_train_weights(p)                # raises for adapters on projections the backward does not cover
            for name in SITES:
                d = getattr(p, name)
                mods = site_modules(p, name)
                if d.lora_a is None:
                    continue
                r_pad = d.lora_a.shape[0] // len(mods)
                n_rows = d.w.shape[0] // len(mods)
                for i, m in enumerate(mods):
                    if trainable(m):
                        plan.append(("site", p, name, m, i * n_rows, (i + 1) * n_rows, i * r_pad))
                        n += m.r * (m.in_features + m.out_features)
            for key in ("s_cross", "t_cross"):
                attn = getattr(p, key)._attn
                for role, m in (("q", attn.to_q), ("k", attn.to_k), ("v", attn.to_v), ("o", attn.to_out[0])):
                    if trainable(m):
                        plan.append(("cross", p, key, m, role))
                        n += m.r * (m.in_features + m.out_features)
        if not plan:
            raise ValueError("no trainable LoRA adapter found: call unet.add_lora(r) / add_adapter(config) first "
                             "(reference train_svd_lora.py:1081-1102)")
        # the latent-knowledge block's 'quaternion' parameters are trained with the adapters (train_svd_lora.py:1068-1073)
        self.lk_params = [(n_, p_) for n_, p_ in unet.named_parameters() if "quaternion" in n_] \
            if hasattr(unet, "_context_train") else []
        n += sum(p_.numel() for _, p_ in self.lk_params)
        self.flat_p = torch.empty(n, device=dev, dtype=torch.float32)
        self.flat_g = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_m = torch.zeros(n, device=dev, dtype=torch.float32)
        self.flat_v = torch.zeros(n, device=dev, dtype=torch.float32)
        self.sumsq = torch.zeros((), device=dev, dtype=torch.float64)
        self.slots: Dict[int, Dict] = {}           # id(layer) -> {site: [adapter dicts]}
        self.cross: List[Dict] = []                # cross-attention adapters, one dict per (layer, s/t cross)
        self.adapters: List[Dict] = []             # every adapter, flat-buffer order
        off = 0
        with torch.no_grad():
            cross_by = {}
            for e in plan:
                m = e[3]
                r, K, N = m.r, m.in_features, m.out_features
                ad = dict(module=m, r=r, scaling=float(m.scaling))
                for nm, shape in (("A", (r, K)), ("B", (N, r))):
                    k = r * K if nm == "A" else N * r
                    ad["p" + nm] = self.flat_p[off:off + k].view(shape)
                    ad[nm] = self.flat_g[off:off + k].view(shape)
                    off += k
                a_, b_ = m.lora_A[m.adapter_name].weight, m.lora_B[m.adapter_name].weight
                ad["pA"].copy_(a_)
                ad["pB"].copy_(b_)
                a_.data, b_.data = ad["pA"], ad["pB"]          # the module's parameters now alias the flat buffer
                self.adapters.append(ad)
                if e[0] == "site":
                    _, p, name, _, n_lo, n_hi, r_lo = e
                    ad.update(n_lo=n_lo, n_hi=n_hi, r_lo=r_lo, dense=getattr(p, name), tw=tr_train_weights(p).lora[name])
                    self.slots.setdefault(id(p), {}).setdefault(name, []).append(ad)
                else:
                    _, p, key, _, role = e
                    c = cross_by.get((id(p), key))
                    if c is None:
                        c = cross_by[(id(p), key)] = dict(layer=p, key=key, pc=getattr(p, key), ads={})
                        self.cross.append(c)
                    c["ads"][role] = ad
            self.lk_grads: Dict[str, torch.Tensor] = {}
            for name, prm in self.lk_params:
                k = prm.numel()
                view = self.flat_p[off:off + k].view(prm.shape)
                view.copy_(prm)
                prm.data = view
                self.lk_grads[name] = self.flat_g[off:off + k].view(prm.shape)
                off += k
        assert off == n
        self.repack()

    # ---- fp32 master parameters -> the bf16 operands the GEMMs read (forward: A_cat, B_blk; backward: transposes)
    def repack(self):
        if self._repack_batch is None:       # the master parameters and the operands never move: one table, one launch
            jobs = []
            for ad in self.adapters:
                if "dense" not in ad:
                    continue
                d, (aT, bT), r = ad["dense"], ad["tw"], ad["r"]
                rs = slice(ad["r_lo"], ad["r_lo"] + r)
                ns = slice(ad["n_lo"], ad["n_hi"])
                jobs += [(ad["pA"], d.lora_a[rs], 1.0), (ad["pA"].t(), aT[:, rs], 1.0),
                         (ad["pB"], d.lora_b[ns, rs], ad["scaling"]), (ad["pB"].t(), bT[rs, ns], ad["scaling"])]
            self._repack_batch = ops.Cast2dBatch(jobs) if jobs else False
        if self._repack_batch:
            self._repack_batch.run()
        for c in self.cross:
            wv, wo = self._cross_weights(c)
            c["pc"].wov.copy_(wo @ wv)                     # a view of the batched [sum C, D] cross-vector matrix
        if self.lk_params and not torch.cuda.is_current_stream_capturing():
            self.unet._lk = None          # dense matrices of the latent-knowledge block are rebuilt from the parameters

    @staticmethod
    def _cross_weights(c):
        """Effective fp32 to_v [C, D] and to_out [C, C] of a KV-length-1 cross-attention: W + s B A (host-side weight
        preparation on [C, D]-sized matrices, as at pack time)."""
        attn = c["pc"]._attn
        out = []
        for role, m in (("v", attn.to_v), ("o", attn.to_out[0])):
            w = (m.base_layer if isinstance(m, M.LoraLinear) else m).weight.detach().float()
            ad = c["ads"].get(role)
            if ad is not None:
                w = w + ad["scaling"] * (ad["pB"] @ ad["pA"])
            elif isinstance(m, M.LoraLinear) and not m.merged:
                raise NotImplementedError("frozen, unmerged adapter next to trained ones on a cross-attention")
            out.append(w)
        return out

    def _cross_backward(self, cond: Conditioning, d_xs: torch.Tensor, d_xt: torch.Tensor):
        """Weight gradients of the cross-attention adapters from the gradients of the per-sample cross vectors
        xs = Wo (Wv ctx) + bo: rank-B outer products on [C, r]-sized matrices (B = batch; negligible next to the step, host-
        side tensor algebra like the latent-knowledge block's)."""
        for c in self.cross:
            pc = c["pc"]
            spatial = c["key"] == "s_cross"
            ctx = (cond.ctx if spatial else cond.ctx_t)[:, 0].float()                       # [B, D]
            dx = (d_xs if spatial else d_xt)[:, pc.off:pc.off + pc.wov.shape[0]]            # [B, C]
            wv, wo = self._cross_weights(c)
            u = ctx @ wv.t()                                                                 # [B, C] = to_v(ctx)
            du = dx @ wo                                                                     # [B, C]
            for role, lhs, rhs in (("o", dx, u), ("v", du, ctx)):                            # dW = lhs^T rhs
                ad = c["ads"].get(role)
                if ad is not None:
                    ad["B"].add_(lhs.t() @ (rhs @ ad["pA"].t()), alpha=ad["scaling"])
                    ad["A"].add_((lhs @ ad["pB"]).t() @ rhs, alpha=ad["scaling"])

    def named_grads(self):
        """(qualified parameter name, fp32 gradient view) in the reference's naming
        (``...attn1.to_q.lora_A.<adapter>.weight``; train_svd_lora_train.txt)."""
        names = {id(m): n for n, m in self.unet.named_modules()}
        out = []
        for ad in self.adapters:
            m = ad["module"]
            base = names[id(m)]
            out.append((f"{base}.lora_A.{m.adapter_name}.weight", ad["A"]))
            out.append((f"{base}.lora_B.{m.adapter_name}.weight", ad["B"]))
        return out + list(self.lk_grads.items())

    def state_tensors(self):
        """(qualified name, fp32 parameter view, exp_avg view, exp_avg_sq view) of every trainable tensor, in the
        module's ``named_parameters()`` order - the order the reference hands to its optimizer
        (``filter(requires_grad, unet.parameters())``, train_svd_lora.py:1179; == train_svd_lora_train.txt)."""
        base = self.flat_p.data_ptr()
        by_name = {}
        for name, g in self.named_grads():
            off = (g.data_ptr() - self.flat_g.data_ptr()) // 4
            n = g.numel()
            by_name[name] = (self.flat_p[off:off + n].view(g.shape), self.flat_m[off:off + n].view(g.shape),
                             self.flat_v[off:off + n].view(g.shape))
        order = [n for n, _ in self.unet.named_parameters() if n in by_name]
        if len(order) != len(by_name):
            raise RuntimeError("trainable tensors are missing from the module's parameter list")
        assert base == self.flat_p.data_ptr()
        return [(n,) + by_name[n] for n in order]

    def save_state(self, output_dir: str, global_step: int, lora_name: Optional[str] = None, **kw) -> str:
        """``checkpoint-<global_step>`` directory as the reference's loop writes it (lkgd_b200/checkpoint.py)."""
        from . import checkpoint
        return checkpoint.save_state(self, output_dir, global_step, lora_name or self._adapter_name(), **kw)

    def load_state(self, path: str, lora_name: Optional[str] = None, **kw) -> int:
        from . import checkpoint
        self._graph_stale()
        return checkpoint.load_state(self, path, lora_name or self._adapter_name(), **kw)

    def _adapter_name(self) -> str:
        return self.adapters[0]["module"].adapter_name

    def _graph_stale(self):
        """Parameters restored from a checkpoint are copied INTO the flat buffer the captured graphs read, so the graphs
        stay valid; the latent-knowledge block's derived matrices are rebuilt inside the graph on every replay."""
        return None

    # ---- one training step
    def forward_backward(self, latents: torch.Tensor, noise: torch.Tensor, sigmas: torch.Tensor,
                         cond_latents: torch.Tensor, encoder_hidden_states: torch.Tensor, added_time_ids: torch.Tensor,
                         *extra, zero_grad: bool = True) -> torch.Tensor:
        """latents / noise fp32 [B,F,4,h,w], sigmas fp32 [B], cond_latents fp32 [B,4,h,w] (first-frame latent, already
        masked by the conditioning dropout), encoder_hidden_states [B,1,D], added_time_ids [B,3], extra =
        (domain_features, flow_features) for the LKGD UNet.  Returns the loss (device double scalar); the LoRA
        gradients are accumulated into the flat gradient buffer."""
        unet = self.unet
        pk = unet.packed()
        if pk is not self._pk:
            raise RuntimeError("the UNet was re-packed (invalidate() / load_state_dict / .to) after this LoraTrainer was "
                               "built: its gradient slots and bf16 operand views belong to the previous pack - build a new "
                               "trainer, or restore adapters through LoraTrainer.load_state")
        dev = unet.device
        B, F, Cl, h, w = latents.shape
        f32 = torch.float32
        latents, noise = latents.to(dev, f32).contiguous(), noise.to(dev, f32).contiguous()
        sigmas = sigmas.to(dev, f32).contiguous()
        if zero_grad:
            self.flat_g.zero_()
        noisy, x_in = ops.edm_precondition(latents, noise, sigmas, cond_latents.to(dev, f32).contiguous(), pk.cin_pad)
        timesteps = 0.25 * torch.log(sigmas)                      # train_svd_lora.py:1508-1509 (host-side scalar math)
        lk_saved = cross = None
        if self.lk_params:
            ctx, lk_saved = unet._context_train(encoder_hidden_states.to(dev), *[e.to(dev) for e in extra])
        else:
            ctx = unet._context(encoder_hidden_states.to(dev), *[e.to(dev) for e in extra])
        if self.lk_params or self.cross:
            d_xs = torch.zeros((B, pk.xs_w.shape[0]), device=dev, dtype=f32)    # gradients of every KV=1 cross-attention
            d_xt = torch.zeros((B, pk.xt_w.shape[0]), device=dev, dtype=f32)    # vector, all layers side by side

            def cross(p):
                return (d_xs[:, p.s_cross.off:p.s_cross.off + p.c], d_xt[:, p.t_cross.off:p.t_cross.off + p.c])
        emb = pk.time_embedding(timesteps, added_time_ids.to(dev))
        cond = Conditioning(pk, emb, ctx)
        g = Geom(B, F, h, w)
        ops.STATS_ARENA.begin(dev)            # one zeroed buffer for the fused GroupNorm statistics of this forward
        graph = TrainGraph(pk)
        pred = graph.forward(x_in, g, cond)
        loss, dpred = ops.edm_loss(pred, noisy, latents, sigmas, pk.conv_out_w.shape[0])
        graph.backward(dpred, lora_grads=self.slots, cross_grads=cross)
        if self.cross:
            self._cross_backward(cond, d_xs, d_xt)
        if lk_saved is not None:
            # cross vectors = ctx @ W^T + b for the concatenated [sum C, 1024] matrices: fold back onto the context
            dctx = ops.small_linear_bwd(d_xs, pk.xs_w)
            ops.small_linear_bwd(d_xt, pk.xt_w, dx=dctx)
            unet._context_backward(lk_saved, dctx, self.lk_grads)
        return loss

    def optimizer_step(self, repack: bool = True):
        """[all-reduce(sum)] -> grad-norm -> clip + AdamW (gradient averaged over ranks in-kernel) -> repack."""
        if self.world > 1:
            from .distributed import allreduce_flat_
            allreduce_flat_(self.flat_g, self.group)
        self.step_count += 1
        scale = 1.0 / self.world
        ops.sumsq(self.flat_g, self.sumsq)
        ops.adamw(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr=self.lr, beta1=self.betas[0],
                  beta2=self.betas[1], eps=self.eps, weight_decay=self.wd, step=self.step_count, grad_scale=scale,
                  sumsq_buf=self.sumsq, max_norm=self.max_norm)
        if repack:
            self.repack()

    @property
    def device(self):
        return self.unet.device

    @ops.on_own_device
    def train_step(self, *args, **kw) -> torch.Tensor:
        if self._graph is not None:
            return self._graphed_step(*args)
        loss = self.forward_backward(*args, **kw)
        self.optimizer_step()
        return loss

    # ---- CUDA graphs: the step is ~1300 small launches at training sizes (14 frames, 40x64 latents) and would be
    # bound by the Python / driver launch path; every shape is static, so forward+backward and the repack are captured
    # once and replayed.  The optimizer (host-side step count) and the NCCL all-reduce stay outside the graphs.
    _graph = None

    @ops.on_own_device
    def capture(self, *batch):
        """Captures forward_backward (and the repack) for batches shaped like ``batch`` (device tensors: latents, noise,
        sigmas, cond_latents, encoder_hidden_states, added_time_ids[, domain_features, flow_features]).  Call after at
        least one eager step (lazy weight preprocessing and one-time kernel attributes must not happen under capture)."""
        dev = self.unet.device
        self._static = [b.to(dev).clone() for b in batch]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            self.forward_backward(*self._static)            # warm the allocator on the capture stream
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g1, g2 = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
        if self.lk_params:
            self.unet._lk = None       # rebuilt INSIDE the capture: every replay re-derives the dense matrices of the
            #                            latent-knowledge block from the parameters the optimizer just updated
        with torch.cuda.graph(g1):
            self._static_loss = self.forward_backward(*self._static)
        with torch.cuda.graph(g2, pool=g1.pool()):
            self.repack()
        self._graph = (g1, g2)
        return self

    def _graphed_step(self, *batch) -> torch.Tensor:
        for s, b in zip(self._static, batch):
            if s.data_ptr() != b.data_ptr():
                s.copy_(b, non_blocking=True)
        self._graph[0].replay()
        self.optimizer_step(repack=False)
        self._graph[1].replay()
        return self._static_loss
