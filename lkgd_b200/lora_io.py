"""LoRA / latent-knowledge adapter wire format: what the reference writes after training and reads before sampling.

Reference write path (``train_models/train_svd_lora.py:1735-1747``):
``StableDiffusionPipeline.save_lora_weights(save_directory, unet_lora_layers=convert_state_dict_to_diffusers(
get_peft_model_state_dict(unet, adapter_name)), safe_serialization=True)`` ->
``<dir>/pytorch_lora_weights.safetensors``.

* ``get_peft_model_state_dict`` (in-tree, ``utils/peft_utils.py:189-326``): every state-dict entry whose key contains
  ``"lora_"`` - the LoRA pairs AND the ``quaternion_lora_*`` tensors of the latent-knowledge block - with
  the ``.<adapter_name>`` component removed (``...to_q.lora_A.weight``).
* ``convert_state_dict_to_diffusers`` / ``save_lora_weights`` (diffusers 0.27.2, un-vendored; restated from its
  ``utils/state_dict_utils.py::PEFT_TO_DIFFUSERS`` table): ``to_{q,k,v}.lora_A`` -> ``.lora.down``, ``lora_B`` ->
  ``.lora.up``, ``to_out.0`` likewise; every key prefixed with ``unet.``; safetensors metadata ``{"format": "pt"}``.
Reference read paths: ``set_peft_model_state_dict`` (``utils/peft_utils.py:329-394``: re-inserts the adapter name) and
the hand-rolled loop of ``run_models/run_inference_svd.py:183-207`` (strip ``unet.``, copy where the name matches).

Pure host-side tensor bookkeeping - no kernels."""
from __future__ import annotations

import os
import re
from typing import Dict, Optional

import torch

from . import modules as M

LORA_WEIGHT_NAME_SAFE = "pytorch_lora_weights.safetensors"       # diffusers loaders/lora.py
_PEFT_TO_DIFFUSERS = {          # diffusers utils/state_dict_utils.py (entries that occur in a UNet)
    "to_out.0.lora_A": "to_out.0.lora.down", "to_out.0.lora_B": "to_out.0.lora.up",
    "to_k.lora_A": "to_k.lora.down", "to_k.lora_B": "to_k.lora.up",
    "to_q.lora_A": "to_q.lora.down", "to_q.lora_B": "to_q.lora.up",
    "to_v.lora_A": "to_v.lora.down", "to_v.lora_B": "to_v.lora.up",
}
_DIFFUSERS_TO_PEFT = {v: k for k, v in _PEFT_TO_DIFFUSERS.items()}


def get_peft_model_state_dict(unet, adapter_name: str = "default") -> Dict[str, torch.Tensor]:
    """Adapter tensors under the keys the reference's ``get_peft_model_state_dict`` returns (bias="none")."""
    sd = unet.state_dict()
    keep = {k: v for k, v in sd.items() if "lora_" in k or f".{adapter_name}." in k}
    return {k.replace(f".{adapter_name}", ""): v for k, v in keep.items()}


def _with_adapter_name(key: str, adapter_name: str) -> str:
    """Inverse of the name stripping, as ``set_peft_model_state_dict`` does it (utils/peft_utils.py:361-371)."""
    if "lora_" not in key:
        return key
    suffix = key.split("lora_")[1]
    if "." in suffix:
        tail = ".".join(suffix.split(".")[1:])
        return key.replace(tail, f"{adapter_name}.{tail}")
    return f"{key}.{adapter_name}"


def convert_state_dict_to_diffusers(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for a, b in _PEFT_TO_DIFFUSERS.items():
            if a in k:
                k = k.replace(a, b)
                break
        out[k] = v
    return out


def convert_state_dict_to_peft(sd: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    out = {}
    for k, v in sd.items():
        for a, b in _DIFFUSERS_TO_PEFT.items():
            if a in k:
                k = k.replace(a, b)
                break
        out[k] = v
    return out


def save_lora_weights(unet, save_directory: str, adapter_name: str = "default",
                      weight_name: str = LORA_WEIGHT_NAME_SAFE) -> str:
    """Writes the file the reference's training loop writes; returns its path."""
    from safetensors.torch import save_file
    sd = convert_state_dict_to_diffusers(get_peft_model_state_dict(unet, adapter_name))
    sd = {f"unet.{k}": v.detach().to("cpu").contiguous() for k, v in sd.items()}
    os.makedirs(save_directory, exist_ok=True)
    path = os.path.join(save_directory, weight_name)
    save_file(sd, path, metadata={"format": "pt"})
    return path


def load_lora_weights(unet, path: str, adapter_name: str = "default", lora_alpha: Optional[float] = None,
                      strict: bool = True) -> Dict[str, list]:
    """Loads a reference-trained adapter file (or directory holding ``pytorch_lora_weights.safetensors``).

    Accepts diffusers-format (``unet.`` prefix, ``lora.down/up``) and PEFT-format keys.  LoRA wrappers are created where
    the file has a pair and the module has none yet (rank read from the tensor shapes, ``lora_alpha`` defaults to the
    rank as in train_svd_lora.py:1081-1083).  Returns {"loaded": [...], "unexpected": [...]}; ``strict`` raises on
    unexpected keys (the reference's hand-rolled loader silently skips them, run_inference_svd.py:196-203)."""
    from safetensors.torch import load_file
    if os.path.isdir(path):
        path = os.path.join(path, LORA_WEIGHT_NAME_SAFE)
    raw = load_file(path)
    sd = convert_state_dict_to_peft({(k[len("unet."):] if k.startswith("unet.") else k): v for k, v in raw.items()})
    # create missing wrappers
    pat = re.compile(r"^(.*)\.lora_A\.weight$")
    for k, v in sd.items():
        m = pat.match(k)
        if not m:
            continue
        name = m.group(1)
        parent_name, _, leaf = name.rpartition(".")
        parent = unet.get_submodule(parent_name) if parent_name else unet
        mod = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
        if isinstance(mod, M.LoraLinear):
            if mod.r != v.shape[0]:
                raise ValueError(f"{name}: adapter rank {v.shape[0]} in the file, {mod.r} in the model")
            continue
        r = v.shape[0]
        new = M.LoraLinear(mod, r, r if lora_alpha is None else lora_alpha, "gaussian", adapter_name)
        new = new.to(device=mod.weight.device)
        if leaf.isdigit():
            parent[int(leaf)] = new
        else:
            setattr(parent, leaf, new)
    own = unet.state_dict()
    loaded, unexpected = [], []
    with torch.no_grad():
        for k, v in sd.items():
            # LoRA pairs get the adapter name back (set_peft_model_state_dict); the quaternion_lora_* tensors keep their
            # name - the reference's set_peft_model_state_dict mangles those ("quaternion_lora_fuse.default.r_weight",
            # reported as unexpected keys), which is why run_inference_svd.py:196-203 copies them by plain name instead
            for full in (k, _with_adapter_name(k, adapter_name)):
                if full in own and own[full].shape == v.shape:
                    own[full].copy_(v.to(own[full].dtype))
                    loaded.append(full)
                    break
            else:
                unexpected.append(k)
    if strict and unexpected:
        raise KeyError(f"adapter file has {len(unexpected)} tensors the model cannot take, e.g. {unexpected[:3]}")
    unet.invalidate()
    return {"loaded": loaded, "unexpected": unexpected}
