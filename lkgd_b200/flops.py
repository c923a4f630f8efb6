"""Algorithmic work of one UNet forward (SURVEY.md Appendix B formulas): conv 2*k*Cin*Cout*px, linear 2*M*N*K,
attention 4*N^2*C per sequence batch.  Used for the roofline numbers in bench.py / DESIGN.md."""
from __future__ import annotations

from typing import Dict


def unet_flops(cfg: dict, B: int, F: int, H: int, W: int, lora_rank: int = 0, count_dead_cross_attn: bool = True,
               encoder_only: bool = False) -> Dict[str, float]:
    """``count_dead_cross_attn``: True counts the KV-length-1 cross-attention as the REFERENCE executes it (to_q, to_k
    and a per-token to_out whose result is independent of the query); False counts what lkgd_b200 executes
    (one [C, D] mat-vec per batch element).  ``encoder_only``: conv_in + down blocks + mid block (the part a
    ControlNetSDVModel copies, models/controlnet_sdv.py:219-316)."""
    chans = cfg["block_out_channels"]
    n = len(chans)
    heads = cfg["num_attention_heads"]
    heads = tuple(heads) if isinstance(heads, (tuple, list)) else (heads,) * n
    lpb = cfg["layers_per_block"]
    lpb = tuple(lpb) if isinstance(lpb, (tuple, list)) else (lpb,) * n
    xdim = cfg["cross_attention_dim"]
    xdim = tuple(xdim) if isinstance(xdim, (tuple, list)) else (xdim,) * n
    out: Dict[str, float] = dict(conv3x3=0.0, tconv=0.0, shortcut=0.0, proj=0.0, geglu_ff=0.0, spatial_attn=0.0,
                                 temporal_attn=0.0, cross_attn=0.0, lora=0.0)
    BF = B * F

    def res(cin, cout, h, w):
        px = BF * h * w
        out["conv3x3"] += 2 * 9 * cin * cout * px + 2 * 9 * cout * cout * px
        out["tconv"] += 2 * (2 * 3 * cout * cout * px)
        if cin != cout:
            out["shortcut"] += 2 * cin * cout * px

    def tr(c, h, w, xd):
        m = BF * h * w
        hw = h * w
        out["proj"] += 2 * (2 * m * c * c)                       # proj_in / proj_out
        out["proj"] += 2 * (4 * 2 * m * c * c)                   # spatial + temporal q,k,v,out
        out["geglu_ff"] += 3 * (2 * m * c * 8 * c + 2 * m * 4 * c * c)
        out["spatial_attn"] += BF * 4 * hw * hw * c
        out["temporal_attn"] += B * hw * 4 * F * F * c
        if count_dead_cross_attn:                                # as the reference executes it
            out["cross_attn"] += 2 * (2 * 2 * m * c * c)
        out["cross_attn"] += 2 * (2 * 2 * B * xd * c)
        if lora_rank:
            out["lora"] += 3 * (2 * m * c * lora_rank + 2 * m * lora_rank * c)

    h, w = H, W
    px = BF * h * w
    out["conv3x3"] += 2 * 9 * cfg["in_channels"] * chans[0] * px
    c = chans[0]
    skip_c = [c]
    for i, t in enumerate(cfg["down_block_types"]):
        for j in range(lpb[i]):
            res(c, chans[i], h, w)
            c = chans[i]
            if t.startswith("CrossAttn"):
                tr(c, h, w, xdim[i])
            skip_c.append(c)
        if i != n - 1:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            out["conv3x3"] += 2 * 9 * c * c * BF * h * w
            skip_c.append(c)
    res(c, c, h, w)
    tr(c, h, w, xdim[-1])
    res(c, c, h, w)
    if encoder_only:
        out["total"] = sum(out.values())
        return out
    rc = list(reversed(chans))
    rh = list(reversed(lpb))
    rx = list(reversed(xdim))
    for i, t in enumerate(cfg["up_block_types"]):
        for j in range(rh[i] + 1):
            res(c + skip_c.pop(), rc[i], h, w)
            c = rc[i]
            if t.startswith("CrossAttn"):
                tr(c, h, w, rx[i])
        if i != n - 1:
            h, w = h * 2, w * 2
            out["conv3x3"] += 2 * 9 * c * c * BF * h * w
    out["conv3x3"] += 2 * 9 * c * cfg["out_channels"] * BF * h * w
    out["total"] = sum(out.values())
    return out


def controlnet_flops(cfg: dict, B: int, F: int, H: int, W: int, cond_channels: int = 3,
                     cond_embed_channels=(16, 32, 96, 256), count_dead_cross_attn: bool = True) -> Dict[str, float]:
    """ControlNetSDVModel.forward (models/controlnet_sdv.py:441-578): the UNet's encoder + mid block, the pixel-resolution
    condition encoder (:64-119: conv3x3 Cc->16, then per level conv3x3 c->c and conv3x3 stride 2 c->c', zero conv3x3
    256->C0 at latent resolution) and the 12 + 1 zero 1x1 convs (:290-316)."""
    out = unet_flops(cfg, B, F, H, W, 0, count_dead_cross_attn, encoder_only=True)
    out.pop("total")
    BF = B * F
    h, w = 8 * H, 8 * W
    ce = list(cond_embed_channels)
    f = 2 * 9 * cond_channels * ce[0] * BF * h * w
    for i in range(len(ce) - 1):
        f += 2 * 9 * ce[i] * ce[i] * BF * h * w
        h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
        f += 2 * 9 * ce[i] * ce[i + 1] * BF * h * w
    f += 2 * 9 * ce[-1] * cfg["block_out_channels"][0] * BF * h * w
    out["cond_embedding"] = float(f)
    chans = cfg["block_out_channels"]
    n = len(chans)
    lpb = cfg["layers_per_block"]
    lpb = tuple(lpb) if isinstance(lpb, (tuple, list)) else (lpb,) * n
    z, h, w = 2 * chans[0] * chans[0] * BF * H * W, H, W
    for i in range(n):
        z += lpb[i] * 2 * chans[i] * chans[i] * BF * h * w
        if i != n - 1:
            h, w = (h - 1) // 2 + 1, (w - 1) // 2 + 1
            z += 2 * chans[i] * chans[i] * BF * h * w
    z += 2 * chans[-1] * chans[-1] * BF * h * w
    out["zero_convs"] = float(z)
    out["total"] = sum(out.values())
    return out


def vae_flops(cfg: dict, n_frames: int, h: int, w: int) -> Dict[str, float]:
    """AutoencoderKLTemporalDecoder (lkgd_b200/vae.py): ``decode`` of ``n_frames`` latent frames of h x w and ``encode`` of
    ONE 8h x 8w image, with the conv / linear / attention formulas above."""
    boc = list(cfg["block_out_channels"])
    lpb, lat, cin, cout = cfg["layers_per_block"], cfg["latent_channels"], cfg["in_channels"], cfg["out_channels"]
    top = boc[-1]

    def conv(ci, co, px, k=9):
        return 2.0 * k * ci * co * px

    def attn(c, px_img, n):
        return n * (4 * 2.0 * px_img * c * c + 4.0 * px_img * px_img * c)

    # ---- temporal decoder
    px = n_frames * h * w

    def tres(ci, co, px):
        f = conv(ci, co, px) + conv(co, co, px) + 2 * conv(co, co, px, 3)
        return f + (conv(ci, co, px, 1) if ci != co else 0.0)

    d = conv(lat, top, px) + lpb * tres(top, top, px) + attn(top, h * w, n_frames)
    c, hh, ww = top, h, w
    rev = list(reversed(boc))
    for i, co in enumerate(rev):
        p_ = n_frames * hh * ww
        d += tres(c, co, p_) + lpb * tres(co, co, p_)
        c = co
        if i != len(rev) - 1:
            hh, ww = 2 * hh, 2 * ww
            d += conv(c, c, n_frames * hh * ww)
    d += conv(c, cout, n_frames * hh * ww) + conv(cout, cout, n_frames * hh * ww, 3)
    # ---- encoder (one image)
    def res(ci, co, px):
        return conv(ci, co, px) + conv(co, co, px) + (conv(ci, co, px, 1) if ci != co else 0.0)

    H, W = 8 * h, 8 * w
    e = conv(cin, boc[0], H * W)
    c = boc[0]
    for i, co in enumerate(boc):
        e += res(c, co, H * W) + (lpb - 1) * res(co, co, H * W)
        c = co
        if i != len(boc) - 1:
            H, W = H // 2, W // 2
            e += conv(c, c, H * W)
    e += 2 * res(c, c, H * W) + attn(c, H * W, 1) + conv(c, 2 * lat, H * W) + conv(2 * lat, 2 * lat, H * W, 1)
    return {"decode": float(d), "encode": float(e)}
