"""The reference's joint-attention patch API (``patch/patch.py``) for lkgd_b200 models.

The reference class-swaps every ``BasicTransformerBlock`` / ``TemporalBasicTransformerBlock`` with a ``ToMeBlock`` whose
forward adds a SECOND attention, ``attn1n``, over the partner sample of the batch (x <-> y of a jointly generated pair):
``attn_output = attn1(n) + post(attn1n(n, encoder_hidden_states = partner's n [frames flipped])) * joint_scale``
(patch/patch.py:434-492; temporal variant :617-658).  Here the blocks are parameter containers and the arithmetic runs in
the CUDA engine, so "patching" marks the blocks and creates the extra parameters under the reference's names; the engine
(``engine.run_transformer``) executes the joint branch: the partner's keys / values are addressed in place (per-sample
pointer offsets, no gathered copy), the post layer and ``joint_scale`` are folded into ``attn1n.to_out`` at pack time and
the sum lands in the residual stream through one GEMM epilogue.

Function names, arguments and defaults follow the reference (``apply_patch`` :719-806, ``remove_patch`` :820-838,
``initialize_joint_layers`` :966-977, ``set_joint_attention`` :938-950, ``set_joint_scale`` :952-964,
``set_joint_attention_mask`` :985-1001, ``set_patch_lora_mask`` :872-896).  Not built: ``add_norm=True`` (the
reference's own forward hands ``timestep=None`` to that AdaLayerNormContinuous, :396,:448 - it cannot run with the stock
``TransformerSpatioTemporalModel``), ``single_dir`` (commented out in the reference, "bug in lora mask", :461-467); the per-sample masked LoRA forward (``hack_lora_forward`` :911-922) is built for the
GEMM-path projections and the KV-length-1 cross-attentions (one collapsed matrix per adapter pattern) and refuses partial
masks on the GEGLU projection / attn2 with KV length > 1 (merged at pack time)."""
from __future__ import annotations

import torch

from . import modules as M


def _models(model):
    model0 = model.unet if hasattr(model, "unet") else model
    out = [model0]
    if hasattr(model, "controlnet") and model.controlnet is not None:
        out.append(model.controlnet)
    return out


def _blocks(net):
    for name, m in net.named_modules():
        if isinstance(m, (M.BasicTransformerBlock, M.TemporalBasicTransformerBlock)):
            yield name, m


def apply_patch(model, seed: int = 123, flip: bool = False, with_spatial_block: bool = True,
                with_temporal_block: bool = False, single_dir: bool = False, name_skip=None):
    """Marks the transformer blocks as patched (reference :719-806).  As in the reference a freshly patched block has
    ``enable_joint_attention = True`` (class attribute of ToMeBlock, :104) - the forward then needs
    ``initialize_joint_layers`` and a joint attention mask."""
    if single_dir:
        raise NotImplementedError("single_dir is marked broken in the reference itself (patch.py:445-449)")
    remove_patch(model)
    net = _models(model)[0]        # the reference patches the UNet only (`include_control` is undefined there, :778)
    net._tome_info = {"size": None, "hooks": [], "args": {"generator": None, "seed": seed, "flip": flip,
                                                          "single_dir": single_dir}}
    for name, m in _blocks(net):
        if name_skip is not None and name_skip in name:
            continue
        spatial = isinstance(m, M.BasicTransformerBlock)
        if (spatial and with_spatial_block) or (not spatial and with_temporal_block):
            m.patched, m.enable_joint_attention, m.flip = True, True, bool(flip) and spatial
            m._tome_info = net._tome_info
    net.invalidate()
    return model


def remove_patch(model):
    for net in _models(model):
        for _, m in _blocks(net):
            m.patched, m.enable_joint_attention, m.flip = False, False, False
        if hasattr(net, "invalidate"):
            net.invalidate()
    return model


def _patched(model):
    for net in _models(model):
        for name, m in _blocks(net):
            if m.patched:
                yield net, name, m


def initialize_joint_layers(model, post: str = "conv", add_norm: bool = False):
    nets = set()
    for net, _, m in _patched(model):
        m.initialize_joint_layers(post=post, add_norm=add_norm)
        nets.add(net)
    for net in nets:
        net.invalidate()
    return model


def set_joint_attention(model, enable: bool = True, name_filter=None):
    for net, name, m in _patched(model):
        if name_filter is None or name_filter in name:
            m.set_joint_attention(enable)
            net.invalidate()
    return model


def set_joint_scale(model, scale: float = 1.0):
    for net, _, m in _patched(model):
        m.set_joint_scale(scale)
        net.invalidate()
    return model


def set_joint_attention_mask(model, joint_attn_mask):
    mask = torch.tensor(joint_attn_mask, dtype=torch.bool)
    for net, _, m in _patched(model):
        m.joint_attn_mask = mask
        net.invalidate()
    return model


def set_patch_lora_mask(model, lora_name, lora_mask):
    """Reference :872-896: the mask is stored on the model (the joint UNet's input heads read it) and on every LoRA
    Linear (inverted for ``attn1n.to_k`` / ``attn1n.to_v``, whose input is the partner's hidden state)."""
    mask = torch.tensor(lora_mask, dtype=torch.bool)
    for net in _models(model):
        if not hasattr(net, "lora_mask"):
            net.lora_mask = dict()
        net.lora_mask[lora_name] = mask
        for name, m in net.named_modules():
            if isinstance(m, M.LoraLinear):
                if not hasattr(m, "lora_mask"):
                    m.lora_mask = dict()
                m.lora_mask[lora_name] = ~mask if ("attn1n.to_k" in name or "attn1n.to_v" in name) else mask
        if hasattr(net, "invalidate"):
            net.invalidate()
    return model


def hack_lora_forward(model):
    """Reference :911-922 switches every LoRA Linear to the per-sample masked forward (:57-92): an adapter acts only on the
    samples its ``lora_mask`` selects.  Here the switch is a flag on the LoRA modules; the engine computes a masked
    adapter's down-projection sample range by sample range (``engine.lora_down``).  Masked adapters are built for the
    GEMM-path projections (attn1 / attn1n q, k, v, out; proj_in / proj_out; ff.net.2) and the KV-length-1 cross-attentions
    (``engine.PackedCross``: one collapsed matrix per adapter pattern); on the GEGLU projection and attn2 with more than one
    key, whose adapters are merged at pack time, a partial mask raises when the model is packed."""
    for net in _models(model):
        for name, m in net.named_modules():
            if isinstance(m, M.LoraLinear):
                m.masked_forward = True
        if hasattr(net, "invalidate"):
            net.invalidate()
    return model


def update_patch(model, **kwargs):
    """Reference :841-853: sets attributes on every patched module (the UNet itself carries ``_tome_info`` too)."""
    for net in _models(model):
        for _, m in net.named_modules():
            if hasattr(m, "_tome_info"):
                for k, v in kwargs.items():
                    setattr(m, k, v)
        if hasattr(net, "invalidate"):
            net.invalidate()
    return model


def collect_from_patch(model, attr: str = "tome"):
    """Reference :856-870: {module name: attribute} over the modules that have ``attr``."""
    out = {}
    for net in _models(model):
        for name, m in net.named_modules():
            if hasattr(m, attr):
                out[name] = getattr(m, attr)
    return out


def set_joint_layer_requires_grad(model, adapter_names, requires_grad: bool):
    """Reference :898-909 / :110-133: requires_grad of the named adapters inside ``attn1n`` and of the post layer."""
    if isinstance(adapter_names, str):
        adapter_names = [adapter_names]
    for _, _, m in _patched(model):
        if not hasattr(m, "attn1n"):
            continue
        for sub in m.attn1n.modules():
            if isinstance(sub, M.LoraLinear):
                for d in (sub.lora_A, sub.lora_B):
                    for key, layer in d.items():
                        if key in adapter_names:
                            layer.requires_grad_(requires_grad)
        m.post_joint.requires_grad_(requires_grad)
    return model
