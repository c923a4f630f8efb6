"""Deterministic, construction-order-independent synthetic weights shared by the golden generator and the tests:
every parameter is drawn from a generator seeded by crc32(parameter name), so the reference model (built here from
/root/reference), the oracle and the lkgd_b200 modules get bit-identical tensors without shipping checkpoints.
GEMM / conv weights are rounded to bf16-representable values (the CUDA path stores them in bf16).  Zero-inits that
would hide bugs are overridden: LoRA B, ControlNet zero convs, quaternion texts, mix_factor."""
import zlib

import torch


def fill_seeded_(module: torch.nn.Module, seed: int = 0) -> torch.nn.Module:
    with torch.no_grad():
        for name, p in module.named_parameters():
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
            if name.endswith("mix_factor"):
                v = torch.rand(p.shape, generator=g) * 2 - 1
            elif p.ndim >= 2:
                fan_in = p[0].numel()
                v = (torch.rand(p.shape, generator=g) * 2 - 1) * fan_in ** -0.5
                if "lora_B" in name:
                    v = v * 0.3
                v = v.to(torch.bfloat16).float()
            elif "norm" in name and name.endswith("weight"):
                v = 1.0 + 0.1 * torch.randn(p.shape, generator=g)
            else:
                v = 0.05 * torch.randn(p.shape, generator=g)
            p.copy_(v.to(p.dtype))
    return module


def seeded_tensor(tag: str, shape, seed: int = 0, scale: float = 1.0) -> torch.Tensor:
    g = torch.Generator().manual_seed((zlib.crc32(tag.encode()) + 7919 * seed) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g) * scale
