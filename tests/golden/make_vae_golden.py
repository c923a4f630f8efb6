"""Generates tests/golden/vae_full_size.npz: the fp32 CPU oracle's VAE results at the headline resolution (576x1024 frames,
72x128 latents), so that `pytest -m gpu` checks the CUDA path at full size without paying minutes of CPU.

  decode   2 frames of name-seeded latents through `AutoencoderKLTemporalDecoder.decode(z, num_frames=2)` at the SVD VAE's
           own widths (128, 256, 512, 512): the committed fixture is every 4th pixel of the 2 x 3 x 576 x 1024 result (fp16)
  encode   one 576x1024 name-seeded image through `encode(x).latent_dist` : the full 8 x 72 x 128 moments (fp16)

Weights are name-seeded (tests/golden/weights.py; q / k projections of the attention blocks scaled x2 so that the softmax is
not near-uniform), so the GPU test rebuilds bit-identical tensors in the product module.

    python tests/golden/make_vae_golden.py          # ~3 min of CPU
"""
import pathlib
import sys
import time

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(ROOT), str(HERE)]

import oracle as O                      # noqa: E402  (test infrastructure)
from weights import fill_seeded_, seeded_tensor  # noqa: E402

STEP = 4


def sharpen_attention_(vae):
    with torch.no_grad():
        for name, p in vae.named_parameters():
            if ".attentions." in name and (".to_q." in name or ".to_k." in name) and name.endswith("weight"):
                p.mul_(2.0)
    return vae


def inputs():
    z = seeded_tensor("vae/latents", (2, 4, 72, 128))
    img = torch.tanh(seeded_tensor("vae/image", (1, 3, 576, 1024)))
    return z, img


def build(cls, config):
    return sharpen_attention_(fill_seeded_(cls(**config)))


if __name__ == "__main__":
    torch.manual_seed(0)
    vae = build(O.AutoencoderKLTemporalDecoder, O.SVD_VAE_CONFIG).eval()
    z, img = inputs()
    out = {}
    with torch.no_grad():
        t0 = time.time()
        mom = vae.quant_conv(vae.encoder(img))
        print(f"encode {time.time() - t0:.1f} s", tuple(mom.shape), float(mom.abs().mean()))
        out["encode/moments"] = mom.half().numpy()
        t0 = time.time()
        y = vae.decode(z, num_frames=2).sample
        print(f"decode {time.time() - t0:.1f} s", tuple(y.shape), float(y.abs().mean()))
        out["decode/sub"] = y[:, :, ::STEP, ::STEP].half().numpy()
        out["decode/norm"] = np.array([float(torch.linalg.norm(y.double()))])
    np.savez_compressed(HERE / "vae_full_size.npz", **out)
    print({k: v.shape for k, v in out.items()})
