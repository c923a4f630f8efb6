"""Oracle LoRA (TEST INFRA ONLY).  Follows the reference's in-tree peft==0.10.0 copy
``models/lora_layer.py``: ``update_layer`` :85-130, ``reset_lora_parameters`` :132-150,
``Linear.forward`` :417-443, ``get_delta_weight`` :383-415, ``merge`` :300-361; and the adapter
configuration used by training, ``train_models/train_svd_lora.py:1081-1088``."""
from __future__ import annotations

import math
import re
from typing import Iterable, List

import torch
import torch.nn as nn


class LoraLinear(nn.Module):
    """``result = base(x) + lora_B(lora_A(x.to(A.dtype))) * scaling`` cast back to the base dtype
    (``lora_layer.py:425-442``); scaling = alpha / r (``:106``) or alpha / sqrt(r) (rslora ``:104``)."""

    def __init__(self, base_layer: nn.Linear, r: int, lora_alpha: float, init_lora_weights="gaussian",
                 adapter_name: str = "default", use_rslora: bool = False):
        super().__init__()
        if r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        self.base_layer = base_layer
        self.adapter_name = adapter_name
        self.r = r
        self.scaling = lora_alpha / math.sqrt(r) if use_rslora else lora_alpha / r
        self.lora_A = nn.ModuleDict({adapter_name: nn.Linear(base_layer.in_features, r, bias=False)})
        self.lora_B = nn.ModuleDict({adapter_name: nn.Linear(r, base_layer.out_features, bias=False)})
        self.merged = False
        if init_lora_weights is True:
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        elif isinstance(init_lora_weights, str) and init_lora_weights.lower() == "gaussian":
            nn.init.normal_(self.lora_A[adapter_name].weight, std=1 / r)
        elif init_lora_weights is not False:
            raise ValueError(f"Unknown initialization {init_lora_weights=}")
        if init_lora_weights is not False:
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.scalings = {adapter_name: self.scaling}
        self.active_adapters = [adapter_name]
        self.lora_mask = {}              # adapter -> bool mask over the batch (patch.set_patch_lora_mask)
        self.masked_forward = False      # patch.hack_lora_forward replaces forward by the masked one (patch.py:57-92)

    def update_layer(self, adapter_name: str, r: int, lora_alpha: float, init_lora_weights="gaussian"):
        """A further adapter on the same layer (``lora_layer.py:85-130``); all adapters in ``active_adapters`` add up."""
        if adapter_name in self.lora_A:
            raise ValueError(f"adapter {adapter_name!r} exists")
        self.lora_A[adapter_name] = nn.Linear(self.base_layer.in_features, r, bias=False)
        self.lora_B[adapter_name] = nn.Linear(r, self.base_layer.out_features, bias=False)
        if isinstance(init_lora_weights, str) and init_lora_weights.lower() == "gaussian":
            nn.init.normal_(self.lora_A[adapter_name].weight, std=1 / r)
        elif init_lora_weights is True:
            nn.init.kaiming_uniform_(self.lora_A[adapter_name].weight, a=math.sqrt(5))
        if init_lora_weights is not False:
            nn.init.zeros_(self.lora_B[adapter_name].weight)
        self.scalings[adapter_name] = lora_alpha / r
        self.active_adapters.append(adapter_name)

    @property
    def in_features(self):
        return self.base_layer.in_features

    @property
    def out_features(self):
        return self.base_layer.out_features

    @property
    def weight(self):
        return self.base_layer.weight

    def get_delta_weight(self) -> torch.Tensor:
        a, b = self.lora_A[self.adapter_name].weight, self.lora_B[self.adapter_name].weight
        return (b @ a) * self.scaling

    def merge(self):
        if not self.merged:
            self.base_layer.weight.data += self.get_delta_weight().to(self.base_layer.weight.dtype)
            self.merged = True

    def unmerge(self):
        if self.merged:
            self.base_layer.weight.data -= self.get_delta_weight().to(self.base_layer.weight.dtype)
            self.merged = False

    def forward(self, x):
        result = self.base_layer(x)
        if self.merged:
            return result
        dt = result.dtype
        for name in self.active_adapters:
            if name not in self.lora_A:
                continue
            a, b, s = self.lora_A[name], self.lora_B[name], self.scalings[name]
            x = x.to(a.weight.dtype)
            if self.masked_forward:                                   # patch/patch.py:74-88
                m = self.lora_mask[name]
                m = m.repeat_interleave(x.shape[0] // len(m), dim=0)
                result[m] += b(a(x[m])) * s
            else:                                                     # models/lora_layer.py:425-442
                result = result + b(a(x)) * s
        return result.to(dt)


TEMPORAL_QKV = r".*temporal_transformer_blocks\.0\.attn1\.to_[qkv]$"     # train_svd_lora.py:1081-1088
ALL_ATTN_PROJ = r".*\.(to_q|to_k|to_v|to_out\.0)$"                        # run_inference_flow_lora.py:326-331


def add_lora(model: nn.Module, r: int, lora_alpha: float = None, target: str = TEMPORAL_QKV,
             init_lora_weights="gaussian", adapter_name: str = "default") -> List[str]:
    """Wrap every ``nn.Linear`` whose qualified name matches ``target`` (peft ``add_adapter``)."""
    lora_alpha = r if lora_alpha is None else lora_alpha
    pat = re.compile(target)
    hits = [n for n, m in model.named_modules() if isinstance(m, (nn.Linear, LoraLinear)) and pat.match(n)
            and ".lora_" not in n and not n.endswith("base_layer")]
    for name in hits:
        parent_name, _, leaf = name.rpartition(".")
        parent = model.get_submodule(parent_name) if parent_name else model
        old = getattr(parent, leaf) if not leaf.isdigit() else parent[int(leaf)]
        if isinstance(old, LoraLinear):                       # a further adapter on an already wrapped layer
            old.update_layer(adapter_name, r, lora_alpha, init_lora_weights)
            continue
        wrapped = LoraLinear(old, r, lora_alpha,
                             init_lora_weights, adapter_name)
        if leaf.isdigit():
            parent[int(leaf)] = wrapped
        else:
            setattr(parent, leaf, wrapped)
    return hits


def merge_lora(model: nn.Module):
    for m in model.modules():
        if isinstance(m, LoraLinear):
            m.merge()
