"""Per-stage rel-L2 of the VAE's CUDA path against the CPU oracle (debugging aid: where does the bf16 operand noise of the
deep, skip-free VAE chains accumulate?).  usage: python tools/vae_trace.py [c0,c1,c2,c3]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "tests", "golden")]
import oracle as O  # noqa: E402
import make_vae_golden as VG  # noqa: E402
from weights import seeded_tensor  # noqa: E402
from lkgd_b200 import ops, vae as V  # noqa: E402
from lkgd_b200.engine import Geom, run_resblock  # noqa: E402
from lkgd_b200.ops import A_CONV3X3  # noqa: E402

boc = tuple(int(v) for v in sys.argv[1].split(",")) if len(sys.argv) > 1 else (32, 64, 128, 256)
dev = torch.device("cuda:0")
o = VG.build(O.AutoencoderKLTemporalDecoder, dict(block_out_channels=boc)).eval()
p = V.AutoencoderKLTemporalDecoder(block_out_channels=boc)
p.load_state_dict(o.state_dict())
p = p.to(dev)
pk = p._pack()


def rel(rows, ref, n, H, W):
    a = rows.float().reshape(n, H, W, -1).permute(0, 3, 1, 2).cpu().double()
    b = ref.double()
    return float(torch.linalg.norm(a - b) / torch.linalg.norm(b))


with torch.no_grad():
    img = torch.tanh(seeded_tensor("vae/img", (2, 3, 64, 96)))
    n, H, W = 2, 64, 96
    ops.STATS_ARENA.begin(dev)
    rows = ops.pack_input(img.to(dev)[:, None], 1.0, None, n, V.IN_CPAD)
    h = ops.gemm(rows, pk.e_in_w, mode=A_CONV3X3, conv=(n, H, W, 1), bias=pk.e_in_b, out_f32=True, gn_rows=H * W)
    r = o.encoder.conv_in(img)
    print("enc conv_in", rel(h, r, n, H, W))
    for bi, ((res, ds), ob) in enumerate(zip(pk.e_down, o.encoder.down_blocks)):
        for ri, (rr, orr) in enumerate(zip(res, ob.resnets)):
            h = V._run_resnet2d(rr, h, n, H, W)
            r = orr(r, None)
            print(f"enc down{bi}.res{ri}", rel(h, r, n, H, W))
        if ds is not None:
            hb = ops.cast_bf16(h)
            H, W = H // 2, W // 2
            h = ops.gemm(hb, ds[0], mode=A_CONV3X3, conv=(n, 2 * H, 2 * W, 2), pad_br=True, bias=ds[1], out_f32=True, gn_rows=H * W)
            r = ob.downsamplers[0](r)
            print(f"enc down{bi}.ds", rel(h, r, n, H, W))
    res, att = pk.e_mid
    h = V._run_resnet2d(res[0], h, n, H, W); r = o.encoder.mid_block.resnets[0](r, None)
    print("enc mid.res0", rel(h, r, n, H, W))
    h = V._run_attention(att, h, n, H * W); r = o.encoder.mid_block.attentions[0](r)
    print("enc mid.attn", rel(h, r, n, H, W))
    h = V._run_resnet2d(res[1], h, n, H, W); r = o.encoder.mid_block.resnets[1](r, None)
    print("enc mid.res1", rel(h, r, n, H, W))

    z = seeded_tensor("vae/z", (6, 4, 8, 12))
    n, H, W = 6, 8, 12
    g = Geom(2, 3, H, W)
    ioi = torch.zeros(2, 3)
    ops.STATS_ARENA.begin(dev)
    rows = ops.pack_input(z.to(dev)[:, None], 1.0, None, n, V.IN_CPAD)
    h = ops.gemm(rows, pk.d_in_w, mode=A_CONV3X3, conv=(n, H, W, 1), bias=pk.d_in_b, out_f32=True, gn_rows=g.HW)
    r = o.decoder.conv_in(z)
    print("dec conv_in", rel(h, r, n, g.H, g.W))
    res, att = pk.d_mid
    h = run_resblock(res[0], h, None, g, None); r = o.decoder.mid_block.resnets[0](r, None, ioi)
    print("dec mid.res0", rel(h, r, n, g.H, g.W))
    h = V._run_attention(att, h, n, g.HW); r = o.decoder.mid_block.attentions[0](r)
    print("dec mid.attn", rel(h, r, n, g.H, g.W))
    h = run_resblock(res[1], h, None, g, None); r = o.decoder.mid_block.resnets[1](r, None, ioi)
    print("dec mid.res1", rel(h, r, n, g.H, g.W))
    for bi, ((res, us), ob) in enumerate(zip(pk.d_up, o.decoder.up_blocks)):
        for ri, (rr, orr) in enumerate(zip(res, ob.resnets)):
            h = run_resblock(rr, h, None, g, None); r = orr(r, None, ioi)
            print(f"dec up{bi}.res{ri}", rel(h, r, n, g.H, g.W))
        if us is not None:
            hu = ops.upsample2x(h, n, g.H, g.W)
            g = g.up()
            h = ops.gemm(hu, us[0], mode=A_CONV3X3, conv=(n, g.H, g.W, 1), bias=us[1], out_f32=True, gn_rows=g.HW)
            r = ob.upsamplers[0](r)
            print(f"dec up{bi}.us", rel(h, r, n, g.H, g.W))
