"""Reference-facing UNet / ControlNet modules: same constructor arguments, parameter names, ``forward``
signatures and error behaviour as the reference classes, with the forward pass executed by the CUDA engine.

  UNetSpatioTemporalConditionControlNetModel  <- models/unet_spatio_temporal_condition_controlnet.py:69-245,358-508
  UNetSpatioTemporalConditionModel (LKGD)     <- models/unet_spatio_temporal_condition.py:72-298,448-693
  ControlNetSDVModel                          <- models/controlnet_sdv.py:160-316,441-578,581-638
  UNetSpatioTemporalConditionModelFlow        <- models/unet_spatio_temporal_condition_flow.py (second, gated input stem)
  UNetSpatioTemporalConditionJointModel       <- models/unet_spatio_temporal_condition_joint.py (x / y input heads)

``forward`` is the inference path (runs under no_grad).  The LoRA fine-tuning step (SURVEY.md section 8 rows a14 / a15:
forward with saved activations, hand-scheduled backward, clip + AdamW, flat gradient all-reduce) lives in
``lkgd_b200/training.py`` and works on the same modules; construction from reference artefacts (``from_reference``,
``from_pretrained``, ``add_adapter``) is on ``_Base`` below."""
from __future__ import annotations

import math
import os
import re
from dataclasses import dataclass
from types import SimpleNamespace
from typing import List, Optional, Sequence, Tuple, Union

import torch
import torch.nn as nn

from . import modules as M
from . import ops
from .engine import Conditioning, Geom, PackedUNet, SL_LEAKY, SL_SILU, _conv3x3_weight, _f32, residual_multipliers
from .ops import A_CONV3X3, ACT_SILU, bf16

SVD_XT_CONFIG = dict(
    sample_size=96, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 3 + ("DownBlockSpatioTemporal",),
    up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 3,
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20), num_frames=25,
)
REDUCED_CONFIG = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32,
    projection_class_embeddings_input_dim=96, layers_per_block=2, cross_attention_dim=32,
    transformer_layers_per_block=1, num_attention_heads=(2, 4), num_frames=8,
)

TEMPORAL_QKV = r".*temporal_transformer_blocks\.0\.attn1\.to_[qkv]$"     # train_models/train_svd_lora.py:1081-1088
ALL_ATTN_PROJ = r".*\.(to_q|to_k|to_v|to_out\.0)$"                        # run_models/run_inference_flow_lora.py:326-331


@dataclass
class UNetSpatioTemporalConditionOutput:
    sample: torch.Tensor = None


@dataclass
class ControlNetOutput:
    down_block_res_samples: Tuple[torch.Tensor]
    mid_block_res_sample: torch.Tensor


class ChannelsLast:
    """A [N, H, W, C] bf16 activation kept in the engine's layout (rows = pixels).  Passing these between
    ``ControlNetSDVModel`` and the UNet skips the NCHW round trip of the reference's residual hand-off."""

    def __init__(self, rows: torch.Tensor, N: int, H: int, W: int):
        self.rows, self.N, self.H, self.W = rows, N, H, W

    def to_nchw(self) -> torch.Tensor:
        return ops.nhwc_to_nchw(self.rows, self.N, self.H, self.W)


def _tuple(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _Base(nn.Module):
    _supports_gradient_checkpointing = True

    # -------------------------------------------------------------------- construction helpers
    def _build_encoder(self, in_channels, down_block_types, block_out_channels, layers_per_block,
                       transformer_layers_per_block, num_attention_heads, cross_attention_dim, addition_time_embed_dim,
                       projection_class_embeddings_input_dim):
        n = len(down_block_types)
        heads, xdim = _tuple(num_attention_heads, n), _tuple(cross_attention_dim, n)
        lpb, tlpb = _tuple(layers_per_block, n), _tuple(transformer_layers_per_block, n)
        c0 = block_out_channels[0]
        temb = c0 * 4
        self.conv_in = M.Conv2d(in_channels, c0, 3, padding=1)
        self.time_embedding = M.TimestepEmbedding(c0, temb)
        self.add_embedding = M.TimestepEmbedding(projection_class_embeddings_input_dim, temb)
        self.down_blocks = nn.ModuleList()
        out_c = c0
        for i, t in enumerate(down_block_types):
            in_c, out_c = out_c, block_out_channels[i]
            last = i == n - 1
            if t == "CrossAttnDownBlockSpatioTemporal":
                blk = M.CrossAttnDownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], tlpb[i], heads[i], xdim[i],
                                                         add_downsample=not last)
            elif t == "DownBlockSpatioTemporal":
                blk = M.DownBlockSpatioTemporal(in_c, out_c, temb, lpb[i], add_downsample=not last)
            else:
                raise ValueError(f"{t} does not exist.")
            self.down_blocks.append(blk)
        return heads, xdim, lpb, tlpb, temb

    # -------------------------------------------------------------------- diffusers-style surface
    @property
    def dtype(self):
        return next(self.parameters()).dtype

    @property
    def device(self):
        return next(self.parameters()).device

    @property
    def attn_processors(self):
        """Attention runs in fixed CUDA kernels; the processor API is kept as a read-only shim
        (reference ...controlnet.py:249-271)."""
        return {n + ".processor": "lkgd_b200" for n, m in self.named_modules() if isinstance(m, M.Attention)}

    def set_attn_processor(self, processor):
        return None

    def set_default_attn_processor(self):
        return None

    def enable_forward_chunking(self, chunk_size=None, dim=0):
        if dim not in (0, 1):      # reference ...controlnet.py:341-342
            raise ValueError(f"Make sure to set `dim` to either 0 or 1, not {dim}")
        return None                # feed-forward chunking is a memory knob of the reference; not needed here

    def _set_gradient_checkpointing(self, module, value=False):
        return None

    # -------------------------------------------------------------------- LoRA (peft add_adapter equivalent)
    def add_lora(self, r: int, lora_alpha: Optional[float] = None, target: str = TEMPORAL_QKV,
                 init_lora_weights="gaussian", adapter_name: str = "default") -> List[str]:
        """Wraps every Linear whose qualified name matches ``target`` (reference train_svd_lora.py:1081-1102)."""
        lora_alpha = r if lora_alpha is None else lora_alpha
        pat = re.compile(target)
        hits = [n for n, m in self.named_modules() if isinstance(m, M.Linear) and pat.match(n)
                and ".lora_" not in n and not n.endswith("base_layer")]
        for name in hits:
            parent_name, _, leaf = name.rpartition(".")
            parent = self.get_submodule(parent_name) if parent_name else self
            old = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
            new = M.LoraLinear(old, r, lora_alpha, init_lora_weights, adapter_name)
            if leaf.isdigit():
                parent[int(leaf)] = new
            else:
                setattr(parent, leaf, new)
        self.invalidate()
        return hits

    def add_adapter(self, adapter_config, adapter_name: str = "default") -> List[str]:
        """``unet.add_adapter(LoraConfig(...), adapter_name)`` as the reference calls it
        (train_models/train_svd_lora.py:1081-1102; run_models/run_inference_flow_lora.py:326-331).  ``adapter_config`` is
        anything shaped like a peft ``LoraConfig`` (attributes or dict keys ``r``, ``lora_alpha``, ``init_lora_weights``,
        ``target_modules``, ``layers_to_transform``, ``layers_pattern``) - peft itself is not needed.  Module selection
        follows peft 0.10's ``check_target_module_exists``: a list of ``target_modules`` matches a module whose name equals
        or ends with ``"." + target``, a string is a full-match regex; with ``layers_to_transform`` the module must also
        sit under ``<layers_pattern>.<index>.`` with the index listed."""
        def get(k, default=None):
            if isinstance(adapter_config, dict):
                return adapter_config.get(k, default)
            return getattr(adapter_config, k, default)

        r = get("r")
        if not isinstance(r, int) or r <= 0:
            raise ValueError(f"`r` should be a positive integer value but the value passed is {r}")
        if get("lora_dropout", 0.0):
            raise NotImplementedError("lora_dropout > 0 is not supported (the reference trains with 0.0)")
        targets = get("target_modules")
        if targets is None:
            raise ValueError("Please specify `target_modules` in `peft_config`")
        layers = get("layers_to_transform")
        layers = [layers] if isinstance(layers, int) else layers
        patterns = get("layers_pattern")
        patterns = [patterns] if isinstance(patterns, str) else (list(patterns) if patterns else [])

        def wanted(key: str) -> bool:
            if isinstance(targets, str):
                found = re.fullmatch(targets, key) is not None
            else:
                found = any(key == t or key.endswith("." + t) for t in targets)
            if found and layers is not None:
                m = None
                if not patterns:
                    m = re.match(r".*\.[^.]*\.(\d+)\.", key)
                for pat in patterns:
                    m = re.match(rf".*\.{pat}\.(\d+)\.", key)
                    if m is not None:
                        break
                found = m is not None and int(m.group(1)) in layers
            return found

        names = [n for n, m in self.named_modules() if isinstance(m, M.Linear) and ".lora_" not in n
                 and not n.endswith("base_layer") and wanted(n)]
        taken = [(n, m) for n, m in self.named_modules() if isinstance(m, M.LoraLinear) and wanted(n)]
        if any(adapter_name in m.lora_A for _, m in taken):
            raise ValueError(f"adapter {adapter_name!r} already exists on {taken[0][0]}")
        alpha0 = get("lora_alpha", r)
        for _, m in taken:                 # a further adapter on layers that already carry one (peft update_layer)
            m.update_layer(adapter_name, r, r if alpha0 is None else alpha0, get("init_lora_weights", True))
        if taken:
            self.invalidate()
        if not names and taken:
            return [n for n, _ in taken]
        if not names:
            raise ValueError(f"Target modules {targets} not found in the base model. Please check the target modules "
                             "and try again.")
        alpha = get("lora_alpha", r)
        return [n for n, _ in taken] + self._wrap_lora(names, r, r if alpha is None else alpha,
                                                      get("init_lora_weights", True), adapter_name)

    def set_adapters(self, adapter_names, weights=None):
        """peft ``set_adapters`` (utils/util.py:596-597): the listed adapters are active on every LoRA layer that has
        them; their updates add up."""
        if weights is not None:
            raise NotImplementedError("per-adapter weights are not supported")
        names = [adapter_names] if isinstance(adapter_names, str) else list(adapter_names)
        for m in self.modules():
            if isinstance(m, M.LoraLinear):
                m.active_adapters = [n for n in names if n in m.lora_A]
        self.invalidate()

    def _wrap_lora(self, names, r, lora_alpha, init_lora_weights, adapter_name):
        for name in names:
            parent_name, _, leaf = name.rpartition(".")
            parent = self.get_submodule(parent_name) if parent_name else self
            old = parent[int(leaf)] if leaf.isdigit() else getattr(parent, leaf)
            if isinstance(old, M.LoraLinear):
                raise ValueError(f"{name} already carries an adapter")
            new = M.LoraLinear(old, r, lora_alpha, init_lora_weights, adapter_name)
            if leaf.isdigit():
                parent[int(leaf)] = new
            else:
                setattr(parent, leaf, new)
        self.invalidate()
        return list(names)

    # -------------------------------------------------------------------- construction from reference artefacts
    @classmethod
    def from_config(cls, config, **overrides):
        """Builds the module from a diffusers-style config (dict / namespace / ``FrozenDict``): keys the constructor does
        not know (``_class_name``, ``_diffusers_version``, ``_name_or_path`` ...) are ignored like ``ConfigMixin`` does."""
        import inspect
        cfg = dict(config) if isinstance(config, dict) else dict(vars(config))
        cfg.update(overrides)
        accepted = set(inspect.signature(cls.__init__).parameters) - {"self"}
        kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in cfg.items() if k in accepted}
        return cls(**kw)

    @classmethod
    def from_reference(cls, module, device=None):
        """Drop-in conversion of a live reference module (``models/unet_spatio_temporal_condition*.py`` /
        ``models/controlnet_sdv.py`` instance, with or without peft LoRA wrappers): same config, same ``state_dict``
        (run_models/run_inference.py:279-281 hands such instances to the pipeline)."""
        net = cls.from_config(module.config)
        sd = {k: v.detach() for k, v in module.state_dict().items()}
        net._adopt_state_dict(sd)
        dev = device if device is not None else next(iter(sd.values())).device
        return net.to(dev)

    def _adopt_state_dict(self, sd, strict: bool = True):
        """load_state_dict that first creates the LoRA wrappers the checkpoint implies (``<m>.base_layer.weight`` +
        ``<m>.lora_A.<adapter>.weight``; rank from the tensor shape, lora_alpha = rank as the reference configures it)."""
        pat = re.compile(r"^(.*)\.lora_A\.([^.]+)\.weight$")
        groups = {}
        for k, v in sd.items():
            m = pat.match(k)
            if m:
                groups.setdefault((m.group(2), v.shape[0]), []).append(m.group(1))
        for (adapter, r), names in groups.items():
            self._wrap_lora(names, r, r, "gaussian", adapter)
        return self.load_state_dict(sd, strict=strict)

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path: str, subfolder: Optional[str] = None, device=None,
                        torch_dtype=None, **config_overrides):
        """``Model.from_pretrained(dir, subfolder="unet")`` for a LOCAL diffusers model directory (``config.json`` +
        ``diffusion_pytorch_model.safetensors`` / ``.fp16.safetensors`` / ``.bin``), as run_models/run_inference.py:279-280
        and train_svd_lora.py:1022-1027 call it.  There is no hub download here (no network): a repo id raises."""
        import json
        import os
        root = pretrained_model_name_or_path
        if subfolder:
            root = os.path.join(root, subfolder)
        cfg_path = os.path.join(root, "config.json")
        if not os.path.isfile(cfg_path):
            raise EnvironmentError(f"{cfg_path} not found: lkgd_b200 loads local diffusers model directories only "
                                   "(download the checkpoint with the reference's tooling first)")
        with open(cfg_path) as f:
            config = json.load(f)
        net = cls.from_config(config, **config_overrides)
        for name in ("diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors",
                     "diffusion_pytorch_model.bin"):
            path = os.path.join(root, name)
            if os.path.isfile(path):
                break
        else:
            raise EnvironmentError(f"no diffusion_pytorch_model.safetensors / .bin under {root}")
        if path.endswith(".safetensors"):
            from safetensors.torch import load_file
            sd = load_file(path)
        else:
            sd = torch.load(path, map_location="cpu", weights_only=True)
        net._adopt_state_dict(sd)
        if torch_dtype is not None and torch_dtype != torch.float32:
            # the engine keeps fp32 masters and packs bf16 operands itself; a half-precision checkpoint is upcast
            pass
        return net.to(device) if device is not None else net

    def save_pretrained(self, save_directory: str, safe_serialization: bool = True):
        """Writes ``config.json`` + ``diffusion_pytorch_model.safetensors`` in the diffusers layout (``from_pretrained``
        here and ``ModelMixin.from_pretrained`` of the reference read it)."""
        import json
        import os
        os.makedirs(save_directory, exist_ok=True)
        cfg = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(self.config).items()}
        cfg["_class_name"] = type(self).__name__
        cfg["_diffusers_version"] = "0.27.2"
        with open(os.path.join(save_directory, "config.json"), "w") as f:
            json.dump(cfg, f, indent=2, sort_keys=True)
        sd = {k: v.detach().to("cpu").contiguous() for k, v in self.state_dict().items()}
        if safe_serialization:
            from safetensors.torch import save_file
            save_file(sd, os.path.join(save_directory, "diffusion_pytorch_model.safetensors"), metadata={"format": "pt"})
        else:
            torch.save(sd, os.path.join(save_directory, "diffusion_pytorch_model.bin"))
        return save_directory

    def merge_lora(self):
        """W += scaling * B A (reference models/lora_layer.py:300-361)."""
        for m in self.modules():
            if isinstance(m, M.LoraLinear) and not m.merged:
                for _, a, b, sc, mask in m.adapters():
                    if mask is not None:
                        raise ValueError("a per-sample masked adapter cannot be merged into the base weight")
                    m.base_layer.weight.data += ((b.float() @ a.float()) * sc).to(m.base_layer.weight.dtype)
                m.merged = True
        self.invalidate()

    # -------------------------------------------------------------------- engine plumbing
    def invalidate(self):
        """Call after changing parameters: the kernel-ready weight copies are rebuilt on the next forward."""
        self._packed = None

    def load_state_dict(self, *a, **k):
        out = super().load_state_dict(*a, **k)
        self.invalidate()
        return out

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self._packed = None
        return out

    def packed(self) -> PackedUNet:
        if getattr(self, "_packed", None) is None:
            if self.device.type != "cuda":
                raise RuntimeError("lkgd_b200 runs on CUDA devices only: move the module with .to('cuda') "
                                   "(there is no CPU or PyTorch fallback)")
            ops.device_check(self.device.index or 0)
            with torch.no_grad():
                self._packed = PackedUNet(self, fold_lora=getattr(self, "fold_lora", True))
        return self._packed

    def _timestep_tensor(self, timestep, sample) -> torch.Tensor:
        """python scalar, 0-d or [B] tensor (reference ...controlnet.py:389-404)."""
        if not torch.is_tensor(timestep):
            return torch.tensor([float(timestep)], dtype=torch.float32, device=sample.device)
        t = timestep.to(device=sample.device, dtype=torch.float32)
        return t[None] if t.ndim == 0 else t

    def _pack_sample(self, sample: torch.Tensor, pk: PackedUNet) -> Tuple[torch.Tensor, Geom]:
        if sample.ndim != 5:
            raise ValueError(f"sample must be [batch, frames, channels, height, width], got {tuple(sample.shape)}")
        B, F, C, H, W = sample.shape
        if C != self._sample_channels():
            raise ValueError(f"sample has {C} channels, the model expects {self._sample_channels()}")
        if F > 32:
            # the temporal attention kernels keep one frame per lane; the reference's pipelines run 14 / 25 frames per call
            # (long videos go through the `smooth` sampler's windows of `num_frames`)
            raise ValueError(f"at most 32 frames per forward, got {F}")
        n_down = sum(1 for d in pk.down if d[2] is not None)
        if H % (1 << n_down) or W % (1 << n_down):
            # the reference has no upsample_size forwarding, skip shapes would not match
            raise ValueError(f"height and width must be divisible by {1 << n_down}, got {H}x{W}")
        x = ops.pack_input(sample, 1.0, None, N=B, Cpad=pk.cin_pad)
        return x, Geom(B, F, H, W)

    def _context(self, encoder_hidden_states: torch.Tensor, *extra) -> torch.Tensor:
        return encoder_hidden_states.to(torch.float32).contiguous()

    def _sample_channels(self) -> int:
        return self.config.in_channels


class UNetSpatioTemporalConditionControlNetModel(_Base):
    """SVD UNet that additionally accepts ControlNet residuals."""

    def __init__(self, sample_size: Optional[int] = None, in_channels: int = 8, out_channels: int = 4,
                 down_block_types: Tuple[str] = SVD_XT_CONFIG["down_block_types"],
                 up_block_types: Tuple[str] = SVD_XT_CONFIG["up_block_types"],
                 block_out_channels: Tuple[int] = (320, 640, 1280, 1280), addition_time_embed_dim: int = 256,
                 projection_class_embeddings_input_dim: int = 768, layers_per_block: Union[int, Tuple[int]] = 2,
                 cross_attention_dim: Union[int, Tuple[int]] = 1024,
                 transformer_layers_per_block: Union[int, Tuple[int]] = 1,
                 num_attention_heads: Union[int, Tuple[int]] = (5, 10, 10, 20), num_frames: int = 25,
                 time_context_order: str = "hw_major_0272"):
        super().__init__()
        self._init_root()
        n = len(down_block_types)
        if len(up_block_types) != n:
            raise ValueError(f"Must provide the same number of `down_block_types` as `up_block_types`. "
                             f"`down_block_types`: {down_block_types}. `up_block_types`: {up_block_types}.")
        if len(block_out_channels) != n:
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        if not isinstance(num_attention_heads, int) and len(num_attention_heads) != n:
            raise ValueError(f"Must provide the same number of `num_attention_heads` as `down_block_types`. "
                             f"`num_attention_heads`: {num_attention_heads}. `down_block_types`: {down_block_types}.")
        if isinstance(cross_attention_dim, (list, tuple)) and len(cross_attention_dim) != n:
            raise ValueError(f"Must provide the same number of `cross_attention_dim` as `down_block_types`. "
                             f"`cross_attention_dim`: {cross_attention_dim}. `down_block_types`: {down_block_types}.")
        if not isinstance(layers_per_block, int) and len(layers_per_block) != n:
            raise ValueError(f"Must provide the same number of `layers_per_block` as `down_block_types`. "
                             f"`layers_per_block`: {layers_per_block}. `down_block_types`: {down_block_types}.")
        if time_context_order not in ("hw_major_0272", "b_major"):
            raise ValueError(f"time_context_order must be 'hw_major_0272' or 'b_major', got {time_context_order}")
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), up_block_types=tuple(up_block_types),
            block_out_channels=tuple(block_out_channels), addition_time_embed_dim=addition_time_embed_dim,
            projection_class_embeddings_input_dim=projection_class_embeddings_input_dim,
            layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
            transformer_layers_per_block=transformer_layers_per_block, num_attention_heads=num_attention_heads,
            num_frames=num_frames, time_context_order=time_context_order)
        self.sample_size = sample_size
        heads, xdim, lpb, tlpb, temb = self._build_encoder(
            in_channels, down_block_types, block_out_channels, layers_per_block, transformer_layers_per_block,
            num_attention_heads, cross_attention_dim, addition_time_embed_dim, projection_class_embeddings_input_dim)
        self.up_blocks = nn.ModuleList()   # registered before the LKGD modules and mid_block (reference order)
        self._init_extra()
        self.mid_block = M.UNetMidBlockSpatioTemporal(block_out_channels[-1], temb, 1, tlpb[-1], heads[-1], xdim[-1])
        rc, rh, rl, rx, rt = (list(reversed(v)) for v in (block_out_channels, heads, lpb, xdim, tlpb))
        out_c = rc[0]
        self.num_upsamplers = 0
        for i, t in enumerate(up_block_types):
            last = i == n - 1
            prev_c, out_c = out_c, rc[i]
            in_c = rc[min(i + 1, n - 1)]
            if not last:
                self.num_upsamplers += 1
            if t == "CrossAttnUpBlockSpatioTemporal":
                blk = M.CrossAttnUpBlockSpatioTemporal(in_c, prev_c, out_c, temb, rl[i] + 1, rt[i], rh[i], rx[i],
                                                       add_upsample=not last)
            elif t == "UpBlockSpatioTemporal":
                blk = M.UpBlockSpatioTemporal(in_c, prev_c, out_c, temb, rl[i] + 1, add_upsample=not last)
            else:
                raise ValueError(f"{t} does not exist.")
            self.up_blocks.append(blk)
        self.conv_norm_out = M.GroupNorm(32, block_out_channels[0], eps=1e-5)
        self.conv_out = M.Conv2d(block_out_channels[0], out_channels, 3, padding=1)
        self._packed = None

    def _init_root(self):
        pass

    def _init_extra(self):
        pass

    # -------------------------------------------------------------------- forward
    def _residual_rows(self, r, g: Geom) -> torch.Tensor:
        if isinstance(r, ChannelsLast):
            return r.rows
        if r.ndim != 4:
            raise ValueError("ControlNet residuals must be [batch*frames, C, H, W] tensors or ChannelsLast")
        return ops.nchw_to_nhwc(r)

    @ops.on_own_device
    @torch.no_grad()
    def forward_packed(self, x: torch.Tensor, g: Geom, timestep, encoder_hidden_states: torch.Tensor, *extra,
                       down_block_additional_residuals=None, mid_block_additional_residual=None,
                       added_time_ids: torch.Tensor = None, batch_slice: Optional[Tuple[int, int]] = None,
                       fused_controlnet=None) -> torch.Tensor:
        """Engine-layout forward.  ``x``: bf16 channels-last rows [B*F*H*W, 64] (input channels zero-padded to 64,
        see ``ops.pack_input``).  Returns the fp32 channels-last prediction [B*F*H*W, out_channels].

        ``batch_slice=(lo, hi)``: ``x`` holds only rows ``lo:hi`` of the batch that ``encoder_hidden_states`` /
        ``added_time_ids`` (and the LKGD features) describe - the CFG pair split across two GPUs.  The
        conditioning is computed for the whole batch (it is microscopic) so that the temporal cross-attention can
        index every half's context exactly like the unsplit reference batch does (diffusers 0.27.2 quirk F8).

        ``fused_controlnet=(controlnet, controlnet_cond, conditioning_scale)``: ControlNet residual injection FUSED into
        this forward (BASELINE.json configs[3]).  The UNet's encoder runs first; the ControlNet then runs on the same
        packed input and its 12 + 1 zero 1x1 convs (models/controlnet_sdv.py:558-571) add ``m_i * scale * conv(skip_cn)``
        straight onto this UNet's skip tensors in their GEMM epilogue (residual read + fused GroupNorm statistics of the
        sum, F6 multipliers folded into the epilogue scale) - no residual tensors, no axpby passes, no narrowing passes,
        and the decoder's GroupNorms keep their one-read path
        (models/unet_spatio_temporal_condition_controlnet.py:453-462,472-473)."""
        if added_time_ids is None:
            raise ValueError("added_time_ids is required")
        pk = self.packed()
        ops.STATS_ARENA.begin(x.device)          # one zeroed buffer for this forward's fused GroupNorm statistics
        ctx_all = self._context(encoder_hidden_states, *extra)
        ids = added_time_ids.to(x.device)
        ctx_t = None
        if batch_slice is not None:
            lo, hi = batch_slice
            if hi - lo != g.B or hi > encoder_hidden_states.shape[0]:
                raise ValueError("batch_slice does not match the packed sample")
            if pk.tctx_mode != 3 and g.HW % 2 and ctx_all.shape[0] > 1:     # RV_BATCH == 3
                raise ValueError("the split batch needs an even number of latent pixels (0.27.2 context order)")
            ctx_t = ctx_all if pk.tctx_mode != 3 else None
            ctx_all, ids = ctx_all[lo:hi].contiguous(), ids[lo:hi].contiguous()
        if ctx_all.shape[0] != g.B or ids.shape[0] != g.B:
            raise ValueError("encoder_hidden_states / added_time_ids batch does not match sample")
        emb = self._embedding(pk, self._timestep_tensor(timestep, x), ids)
        cond = Conditioning(pk, emb, ctx_all, ctx_t)
        x_in = x
        x, skips, geoms, gm = pk.encoder(x, g, cond, stem=self._stem_rows(pk, x, g))
        if fused_controlnet is not None:
            if down_block_additional_residuals is not None or mid_block_additional_residual is not None:
                raise ValueError("fused_controlnet excludes explicit residuals")
            cn, cn_cond, cn_scale = fused_controlnet[:3]
            cn_repeat = fused_controlnet[3] if len(fused_controlnet) > 3 else 1
            per_block = [len(d[0]) + (1 if d[2] is not None else 0) for d in pk.down]
            per_block[0] += 1
            mult = residual_multipliers(len(pk.down), per_block)
            cn_ctx, cn_ctx_t = encoder_hidden_states, None
            if batch_slice is not None:      # one half of the guidance batch (CFG pair split): the ControlNet sees the raw
                #                              image embeddings of this half, its temporal cross-attention those of both
                cn_ctx = encoder_hidden_states[lo:hi]
                cn_ctx_t = encoder_hidden_states if cn.packed().tctx_mode != 3 else None        # RV_BATCH == 3
            x = cn.inject_packed(x_in, g, timestep, cn_ctx, ids, cn_cond, cn_scale, skips, mult, x,
                                 cond_repeat=cn_repeat, ctx_t=cn_ctx_t)
        if mid_block_additional_residual is not None:
            ops.axpby(self._residual_rows(mid_block_additional_residual, gm), 1.0, x, 1.0)
        if down_block_additional_residuals is not None:
            per_block = [len(d[0]) + (1 if d[2] is not None else 0) for d in pk.down]
            per_block[0] += 1   # conv_in output travels with block 0
            mult = residual_multipliers(len(pk.down), per_block)
            # zip truncation of the reference: extra residuals / skips are ignored
            for s, r, m, gs in zip(skips, down_block_additional_residuals, mult, geoms):
                ops.axpby(self._residual_rows(r, gs), float(m), s, 1.0)
        return pk.decoder(x, skips, gm, cond)

    def _embedding(self, pk: PackedUNet, t: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
        return pk.time_embedding(t, ids)

    def _stem_rows(self, pk: PackedUNet, x: torch.Tensor, g: Geom) -> Optional[torch.Tensor]:
        return None                      # conv_in runs inside the encoder

    @ops.on_own_device
    @torch.no_grad()
    def forward_rows(self, sample: torch.Tensor, timestep, encoder_hidden_states: torch.Tensor, *extra, **kw):
        x, g = self._pack_sample(sample, self.packed())
        return self.forward_packed(x, g, timestep, encoder_hidden_states, *extra, **kw), g

    @ops.on_own_device
    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True,
                added_time_ids: torch.Tensor = None):
        rows, g = self.forward_rows(sample, timestep, encoder_hidden_states,
                                    down_block_additional_residuals=down_block_additional_residuals,
                                    mid_block_additional_residual=mid_block_additional_residual,
                                    added_time_ids=added_time_ids)
        out = ops.unpack_output(rows, g.B, g.F, self.config.out_channels, g.H, g.W)
        if out.dtype != sample.dtype:
            out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNetSpatioTemporalConditionOutput(sample=out)


# --------------------------------------------------------------------------------------------------- x / y input heads
class UNetSpatioTemporalConditionJointModel(UNetSpatioTemporalConditionControlNetModel):
    """``models/unet_spatio_temporal_condition_joint.py``: a second set of input heads - ``conv_in_y``,
    ``time_embedding_y``, ``add_embedding_y`` (``add_y_input_head`` :251-280; the weight-free ``time_proj`` /
    ``add_time_proj`` need no copy) - and a forward that sends every sample of the batch through the x or the y heads
    according to ``lora_mask["xy_lora"]`` / ``lora_mask["yx_lora"]`` (:483-500; installed by the reference's
    ``patch.set_patch_lora_mask``, patch/patch.py:872-896, or ``set_lora_mask`` here).  The body is the plain UNet's:
    only the stem conv (one implicit-GEMM launch per sample, written into one row matrix) and the fp32 embedding MLPs
    differ per sample."""

    def add_y_input_head(self):
        import copy
        self.conv_in_y = copy.deepcopy(self.conv_in)
        self.time_embedding_y = copy.deepcopy(self.time_embedding)
        self.add_embedding_y = copy.deepcopy(self.add_embedding)
        self.invalidate()

    def load_y_input_head(self, path):
        """Reference :281-284: a ``torch.save``d dict of the four head state_dicts (``time_proj`` has no tensors)."""
        sd = torch.load(path, map_location="cpu", weights_only=True)
        for name in ("conv_in", "time_embedding", "add_embedding"):
            getattr(self, name + "_y").load_state_dict(sd[name])
        self.invalidate()

    def set_lora_mask(self, lora_name: str, lora_mask):
        """What ``patch.set_patch_lora_mask(unet, lora_name, mask)`` stores on the model (patch/patch.py:880-884)."""
        if not hasattr(self, "lora_mask"):
            self.lora_mask = dict()
        self.lora_mask[lora_name] = torch.as_tensor(lora_mask, dtype=torch.bool)

    def invalidate(self):
        super().invalidate()
        self._y = None

    def _y_pack(self):
        if getattr(self, "_y", None) is None:
            if not hasattr(self, "conv_in_y"):
                raise AttributeError("call add_y_input_head() first")      # the reference fails on conv_in_dict likewise
            pk = self.packed()
            w, b = _conv3x3_weight(self.conv_in_y, cin_pad=pk.cin_pad)
            te, ae = self.time_embedding_y, self.add_embedding_y
            self._y = (w, b,
                       tuple(_f32(t) for t in (te.linear_1.weight, te.linear_1.bias, te.linear_2.weight, te.linear_2.bias)),
                       tuple(_f32(t) for t in (ae.linear_1.weight, ae.linear_1.bias, ae.linear_2.weight, ae.linear_2.bias)))
        return self._y

    def _branches(self, batch: int) -> List[bool]:
        """True = y heads, per sample.  The masks are repeat-interleaved to the batch like the reference (:485-486)."""
        xm, ym = self.lora_mask["xy_lora"], self.lora_mask["yx_lora"]
        xm = xm.repeat_interleave(batch // len(xm)).tolist()
        ym = ym.repeat_interleave(batch // len(ym)).tolist()
        if len(xm) != batch or len(ym) != batch or any(a == b for a, b in zip(xm, ym)):
            raise ValueError("xy_lora / yx_lora masks must partition the batch")
        if not any(xm) or not any(ym):
            raise ValueError("both input heads need at least one sample (the reference's input_layers fails on an empty "
                             "branch, :415)")
        return ym

    def _embedding(self, pk, t, ids):
        B = ids.shape[0]
        if t.numel() not in (1, B):
            raise ValueError("timestep must be a scalar or a [batch] tensor")
        _, _, te_y, ae_y = self._y_pack()
        ex, ey = pk.time_embedding(t, ids), pk.time_embedding(t, ids, te=te_y, ae=ae_y)
        for b, is_y in enumerate(self._branches(B)):
            if is_y:
                ex[b].copy_(ey[b])
        return ex

    def _stem_rows(self, pk, x, g):
        wy, by, _, _ = self._y_pack()
        rows = g.F * g.HW
        out = torch.empty((g.M, pk.c0), device=x.device, dtype=torch.float32)
        for b, is_y in enumerate(self._branches(g.B)):
            ops.gemm(x[b * rows:(b + 1) * rows], wy if is_y else pk.conv_in_w, mode=A_CONV3X3, conv=(g.F, g.H, g.W, 1),
                     bias=by if is_y else pk.conv_in_b, out=out[b * rows:(b + 1) * rows], out_f32=True)
        return out

    def forward_packed(self, x, g, timestep, encoder_hidden_states, *extra, **kw):
        if kw.get("batch_slice") is not None or kw.get("fused_controlnet") is not None:
            raise ValueError("the joint UNet supports neither the CFG pair split nor the fused ControlNet path")
        return super().forward_packed(x, g, timestep, encoder_hidden_states, *extra, **kw)


# --------------------------------------------------------------------------------------------------- LKGD
class UNetSpatioTemporalConditionModelFlow(UNetSpatioTemporalConditionControlNetModel):
    """The reference's flow-stem UNet (``models/unet_spatio_temporal_condition_flow.py``): the
    ControlNet-accepting UNet plus ``initialize_conv_in()`` - a second stem ``conv_in2`` (a copy of ``conv_in``) gated
    per output channel by ``conv_in2_alpha`` (:260-273).  ``forward`` takes a 12-channel sample (noise | condition |
    second condition, :494-502):  conv_in(cat(noise, cond)) + conv_in2(cat(noise, cond2)) * alpha.

    Both stems are linear in the sample, so the engine runs them as ONE 3x3 implicit-GEMM conv over the 12 channels with
    weights merged at pack time (noise taps: W1 + alpha W2; cond: W1; cond2: alpha W2; bias b1 + alpha b2) - no second
    launch, no extra pass over the [B*F, 320, H, W] stem output."""

    def initialize_conv_in(self):
        c = self.conv_in
        self.conv_in2 = M.Conv2d(c.in_channels, c.out_channels, 3, padding=1).to(c.weight.device, c.weight.dtype)
        self.conv_in2_alpha = nn.Parameter(torch.zeros(1, c.out_channels, 1, 1, device=c.weight.device,
                                                       dtype=c.weight.dtype))
        self.conv_in2.load_state_dict(c.state_dict())
        self.invalidate()

    def _sample_channels(self) -> int:
        if not hasattr(self, "conv_in2"):
            return self.config.in_channels
        n = self.config.in_channels // 2
        return 3 * n

    def _stem_conv(self):
        if not hasattr(self, "conv_in2"):
            return self.conv_in
        n = self.config.in_channels // 2
        w1, w2 = self.conv_in.weight.detach().float(), self.conv_in2.weight.detach().float()
        a = self.conv_in2_alpha.detach().float().reshape(-1, 1, 1, 1)
        w = torch.cat([w1[:, :n] + a * w2[:, :n], w1[:, n:], a * w2[:, n:]], dim=1)
        b = self.conv_in.bias.detach().float() + a.reshape(-1) * self.conv_in2.bias.detach().float()
        return SimpleNamespace(weight=w, bias=b)


class UNetSpatioTemporalConditionModel(UNetSpatioTemporalConditionControlNetModel):
    """LKGD UNet: the latent-knowledge block (reference unet_spatio_temporal_condition.py:197-225,536-613) fuses the
    CLIP embedding with domain / flow ViT features (grouped 1x1 conv, quaternion linear, rFFT magnitude / phase
    fuse, iFFT, MLP) into the one-token cross-attention context."""

    def _init_root(self):
        self.quaternion_lora_texts = nn.Parameter(torch.zeros(256))
        self.quaternion_lora_texts_fft_mag = nn.Parameter(torch.zeros(129))
        self.quaternion_lora_texts_fft_pha = nn.Parameter(torch.zeros(129))

    def _init_extra(self):
        def dw():
            return M.Conv1d(1024, 256, kernel_size=1, groups=256, bias=False)
        self.quaternion_lora_dconv, self.quaternion_lora_lconv, self.quaternion_lora_fconv = dw(), dw(), dw()
        self.quaternion_lora_fuse = M.QuaternionLinear(1024, 512)
        self.quaternion_lora_fuse_fft_mag = M.QuaternionLinear(512, 256)
        self.quaternion_lora_fuse_fft_pha = M.QuaternionLinear(512, 256)
        self.quaternion_lora_fuse_fft_mag0 = M.Linear(4, 1)
        self.quaternion_lora_fuse_fft_pha0 = M.Linear(4, 1)
        self.quaternion_lora_fuse_sf = nn.Sequential(M.Linear(1024, 256), nn.LeakyReLU(0.1), M.Linear(256, 1024))

    def invalidate(self):
        super().invalidate()
        self._lk = None

    def _lk_const(self):
        """Parameter-independent matrices of the block, built once per device: the 1000->1024 linear interpolation,
        the rDFT(256) and irDFT(512) bases."""
        dev = self.device
        c = getattr(self, "_lkc", None)
        if c is not None and c["interp64"].device == dev:
            return c
        f64 = torch.float64
        # F.interpolate(size=1024, mode="linear", align_corners=False) as a [1024, 1000] matrix
        o = torch.arange(1024, dtype=f64, device=dev)
        src = ((o + 0.5) * (1000 / 1024) - 0.5).clamp(min=0)
        i0 = src.floor().long().clamp(max=999)
        i1 = (i0 + 1).clamp(max=999)
        lam = src - i0
        interp = torch.zeros(1024, 1000, dtype=f64, device=dev)
        interp[torch.arange(1024, device=dev), i0] += 1 - lam
        interp[torch.arange(1024, device=dev), i1] += lam
        n = torch.arange(256, dtype=f64, device=dev)
        k = torch.arange(129, dtype=f64, device=dev)
        ang = 2 * math.pi * k[:, None] * n[None, :] / 256
        dft_im = -torch.sin(ang)
        dft_im[0] = 0
        dft_im[128] = 0           # rfft returns an exactly-zero imaginary part for the DC and Nyquist bins; the phase
        #                           (torch.angle) is discontinuous there, so the zeros must be exact, not ~1e-13
        t = torch.arange(512, dtype=f64, device=dev)
        k2 = torch.arange(257, dtype=f64, device=dev)
        ang2 = 2 * math.pi * t[:, None] * k2[None, :] / 512
        wgt = torch.full((257,), 2.0, dtype=f64, device=dev)
        wgt[0] = wgt[256] = 1.0
        ir = (torch.cos(ang2) * wgt / 512)
        ii = (-torch.sin(ang2) * wgt / 512)
        ii[:, 0] = 0
        ii[:, 256] = 0              # irfft ignores the imaginary part of the DC and Nyquist bins
        self._lkc = dict(interp64=interp, interp=interp.float().contiguous(), j256=torch.arange(256, device=dev),
                         dft_re=torch.cos(ang).float().contiguous(), dft_im=dft_im.float().contiguous(),
                         idft=torch.cat([ir, ii], 1).float().contiguous())
        return self._lkc

    def _lk_pack(self):
        """Weight preprocessing (once per parameter change): every linear stage of the block as a dense fp32 matrix for
        the small-linear kernel - grouped conv (+ the 1000->1024 linear interpolation folded in), Hamilton matrices.
        Only device-side tensor ops on the parameters (safe under CUDA-graph capture: training re-derives these inside
        the captured step); the parameter-independent bases come from ``_lk_const``."""
        if getattr(self, "_lk", None) is not None:
            return self._lk
        dev = self.device
        f64 = torch.float64
        c = self._lk_const()
        j = c["j256"]

        def grouped(conv):          # Conv1d(1024->256, k=1, groups=256): out_j = sum_m w[j,m] x[4j+m]
            w = conv.weight.detach().to(f64).reshape(256, 4)
            full = torch.zeros(256, 1024, dtype=f64, device=dev)
            for m in range(4):
                full[j, 4 * j + m] = w[:, m]
            return full

        def ham(q):                 # y = x @ W  ->  small_linear weight is W^T [out, in]
            r, i, j_, k = (getattr(q, n).detach().to(f64) for n in ("r_weight", "i_weight", "j_weight", "k_weight"))
            W = torch.cat([torch.cat([r, -i, -j_, -k], 0), torch.cat([i, r, -k, j_], 0),
                           torch.cat([j_, k, r, -i], 0), torch.cat([k, -j_, i, r], 0)], 1)
            return W.t().contiguous().float(), _f32(q.bias)

        sf = self.quaternion_lora_fuse_sf
        gd, gf = grouped(self.quaternion_lora_dconv), grouped(self.quaternion_lora_fconv)
        self._lk = dict(
            interp=c["interp"], dconv_g=gd.float().contiguous(), fconv_g=gf.float().contiguous(),
            lconv=grouped(self.quaternion_lora_lconv).float().contiguous(),
            dconv=(gd @ c["interp64"]).float().contiguous(), fconv=(gf @ c["interp64"]).float().contiguous(),
            fuse=ham(self.quaternion_lora_fuse), mag=ham(self.quaternion_lora_fuse_fft_mag),
            pha=ham(self.quaternion_lora_fuse_fft_pha),
            dft_re=c["dft_re"], dft_im=c["dft_im"], idft=c["idft"],
            mag0=(_f32(self.quaternion_lora_fuse_fft_mag0.weight), _f32(self.quaternion_lora_fuse_fft_mag0.bias)),
            pha0=(_f32(self.quaternion_lora_fuse_fft_pha0.weight), _f32(self.quaternion_lora_fuse_fft_pha0.bias)),
            sf0=(_f32(sf[0].weight), _f32(sf[0].bias)), sf2=(_f32(sf[2].weight), _f32(sf[2].bias)),
            texts=_f32(self.quaternion_lora_texts), tmag=_f32(self.quaternion_lora_texts_fft_mag),
            tpha=_f32(self.quaternion_lora_texts_fft_pha))
        return self._lk

    def _context(self, encoder_hidden_states, domain_features, flow_features) -> torch.Tensor:
        if encoder_hidden_states.shape[1] != 1 or encoder_hidden_states.shape[2] != 1024:
            raise ValueError("the latent-knowledge block expects encoder_hidden_states of shape [B, 1, 1024]")
        lk = self._lk_pack()
        dev = encoder_hidden_states.device
        B = encoder_hidden_states.shape[0]
        f32 = torch.float32
        ctx = encoder_hidden_states.to(f32).reshape(B, 1024).contiguous()
        dom = domain_features.to(f32).reshape(-1, 1000).contiguous()
        flo = flow_features.to(f32).reshape(-1, 1000).contiguous()
        Bd = dom.shape[0]
        cat = torch.empty(B, 1024, device=dev, dtype=f32)          # [lh | ld | lf | texts]
        ops.small_linear(ctx, lk["lconv"], out=cat[:, 0:256])
        if Bd == B:
            ops.small_linear(dom, lk["dconv"], out=cat[:, 256:512])
            ops.small_linear(flo, lk["fconv"], out=cat[:, 512:768])
        elif Bd == 1 and B == 2:                                   # quirk D8: duplicated only for a CFG pair
            for b in range(B):
                ops.small_linear(dom, lk["dconv"], out=cat[b:b + 1, 256:512])
                ops.small_linear(flo[:1], lk["fconv"], out=cat[b:b + 1, 512:768])
        else:
            raise ValueError(f"domain_features batch {Bd} is incompatible with encoder_hidden_states batch {B}")
        cat[:, 768:] = lk["texts"]
        spatial = torch.empty(B, 1024, device=dev, dtype=f32)      # [spatial 512 | freq 512]
        ops.small_linear(cat, *lk["fuse"], out=spatial[:, :512])
        # rFFT(256) of the three lowered vectors as real DFT mat-vecs, then |.| and angle
        mags = torch.empty(B, 4, 129, device=dev, dtype=f32)
        phas = torch.empty(B, 4, 129, device=dev, dtype=f32)
        for i in range(3):
            v = cat[:, 256 * i:256 * (i + 1)]
            re = ops.small_linear(v, lk["dft_re"])
            im = ops.small_linear(v, lk["dft_im"])
            m_, p_ = ops.polar(re, im, 0)
            mags[:, i], phas[:, i] = m_, p_
        mags[:, 3], phas[:, 3] = lk["tmag"], lk["tpha"]
        mag = ops.small_linear(mags[:, :, :128].reshape(B, 512), *lk["mag"])
        pha = ops.small_linear(phas[:, :, :128].reshape(B, 512), *lk["pha"])
        mag0 = ops.small_linear(mags[:, :, 128].contiguous(), *lk["mag0"])
        pha0 = ops.small_linear(phas[:, :, 128].contiguous(), *lk["pha0"])
        re, im = ops.polar(torch.cat([mag, mag0], 1), torch.cat([pha, pha0], 1), 1)        # [B, 257] each
        ops.small_linear(torch.cat([re, im], 1), lk["idft"], out=spatial[:, 512:])          # irFFT -> 512 samples
        h = ops.small_linear(spatial, *lk["sf0"], act_out=SL_LEAKY)
        return ops.small_linear(h, *lk["sf2"]).reshape(B, 1, 1024)

    # -------------------------------------------------------------------- training: forward with save / backward
    def _context_train(self, encoder_hidden_states, domain_features, flow_features):
        """``_context`` with the intermediates the backward needs.  Differs only in evaluating the linear
        interpolation and the grouped conv as two mat-vecs (their weights are trained separately)."""
        from types import SimpleNamespace as NS
        lk = self._lk_pack()
        dev = encoder_hidden_states.device
        B = encoder_hidden_states.shape[0]
        f32 = torch.float32
        S = NS(B=B)
        S.ctx = encoder_hidden_states.to(f32).reshape(B, 1024).contiguous()
        dom = domain_features.to(f32).reshape(-1, 1000).contiguous()
        flo = flow_features.to(f32).reshape(-1, 1000).contiguous()
        if dom.shape[0] not in (B, 1) or (dom.shape[0] == 1 and B > 2):
            raise ValueError(f"domain_features batch {dom.shape[0]} is incompatible with encoder_hidden_states batch {B}")
        S.dom_i = ops.small_linear(dom, lk["interp"]).expand(B, 1024).contiguous()
        S.flo_i = ops.small_linear(flo, lk["interp"]).expand(B, 1024).contiguous()
        S.cat = cat = torch.empty(B, 1024, device=dev, dtype=f32)
        ops.small_linear(S.ctx, lk["lconv"], out=cat[:, 0:256])
        ops.small_linear(S.dom_i, lk["dconv_g"], out=cat[:, 256:512])
        ops.small_linear(S.flo_i, lk["fconv_g"], out=cat[:, 512:768])
        cat[:, 768:] = lk["texts"]
        S.spatial = spatial = torch.empty(B, 1024, device=dev, dtype=f32)
        ops.small_linear(cat, *lk["fuse"], out=spatial[:, :512])
        mags = torch.empty(B, 4, 129, device=dev, dtype=f32)
        phas = torch.empty(B, 4, 129, device=dev, dtype=f32)
        S.re, S.im = [], []
        for i in range(3):
            v = cat[:, 256 * i:256 * (i + 1)]
            re, im = ops.small_linear(v, lk["dft_re"]), ops.small_linear(v, lk["dft_im"])
            S.re.append(re)
            S.im.append(im)
            mags[:, i], phas[:, i] = ops.polar(re, im, 0)
        mags[:, 3], phas[:, 3] = lk["tmag"], lk["tpha"]
        S.magin, S.phain = mags[:, :, :128].reshape(B, 512), phas[:, :, :128].reshape(B, 512)
        S.mag128, S.pha128 = mags[:, :, 128].contiguous(), phas[:, :, 128].contiguous()
        mag, pha = ops.small_linear(S.magin, *lk["mag"]), ops.small_linear(S.phain, *lk["pha"])
        mag0, pha0 = ops.small_linear(S.mag128, *lk["mag0"]), ops.small_linear(S.pha128, *lk["pha0"])
        S.magc, S.phac = torch.cat([mag, mag0], 1), torch.cat([pha, pha0], 1)
        re, im = ops.polar(S.magc, S.phac, 1)
        S.reim = torch.cat([re, im], 1)
        ops.small_linear(S.reim, lk["idft"], out=spatial[:, 512:])
        S.h = ops.small_linear(spatial, *lk["sf0"], act_out=SL_LEAKY)
        return ops.small_linear(S.h, *lk["sf2"]).reshape(B, 1, 1024), S

    def _context_backward(self, S, dctx: torch.Tensor, grads: dict):
        """dctx: fp32 [B, 1024] gradient of the block output.  ``grads``: parameter name -> fp32 accumulator of that
        parameter's shape (the 29 ``quaternion_lora_*`` tensors the reference trains, train_svd_lora.py:1068-1073)."""
        lk = self._lk_pack()
        B, dev, f32 = S.B, dctx.device, torch.float32
        q = "quaternion_lora_"

        def gw(name):
            return grads[q + name]

        def colsum(t, name):
            ops.colsum_grouped(t.contiguous(), 1, (ops.RV_NONE, 1, 1, 1), out=gw(name).reshape(1, -1))

        def qlinear(dy, x, name, key):
            wt = lk[key][0]
            dwt = torch.zeros_like(wt)
            dx = ops.small_linear_bwd(dy, wt, x=x, dW=dwt, db=gw(name + ".bias"))
            ops.hamilton_bwd(dwt, *(gw(f"{name}.{c}_weight") for c in "rijk"))
            return dx

        dh = ops.small_linear_bwd(dctx, lk["sf2"][0], x=S.h, dW=gw("fuse_sf.2.weight"), db=gw("fuse_sf.2.bias"))
        dsp = ops.small_linear_bwd(dh, lk["sf0"][0], x=S.spatial, y=S.h, act_out=SL_LEAKY, dW=gw("fuse_sf.0.weight"),
                                   db=gw("fuse_sf.0.bias"))
        dreim = ops.small_linear_bwd(dsp[:, 512:], lk["idft"])
        dmagc, dphac = ops.polar_bwd(S.magc, S.phac, dreim[:, :257], dreim[:, 257:], 1)
        dmags = torch.empty(B, 4, 129, device=dev, dtype=f32)
        dphas = torch.empty(B, 4, 129, device=dev, dtype=f32)
        dmags[:, :, :128] = qlinear(dmagc[:, :256], S.magin, "fuse_fft_mag", "mag").view(B, 4, 128)
        dphas[:, :, :128] = qlinear(dphac[:, :256], S.phain, "fuse_fft_pha", "pha").view(B, 4, 128)
        dmags[:, :, 128] = ops.small_linear_bwd(dmagc[:, 256:257], lk["mag0"][0], x=S.mag128,
                                                dW=gw("fuse_fft_mag0.weight"), db=gw("fuse_fft_mag0.bias"))
        dphas[:, :, 128] = ops.small_linear_bwd(dphac[:, 256:257], lk["pha0"][0], x=S.pha128,
                                                dW=gw("fuse_fft_pha0.weight"), db=gw("fuse_fft_pha0.bias"))
        colsum(dmags[:, 3], "texts_fft_mag")
        colsum(dphas[:, 3], "texts_fft_pha")
        dcat = qlinear(dsp[:, :512], S.cat, "fuse", "fuse")
        for i in range(3):
            dre, dim = ops.polar_bwd(S.re[i], S.im[i], dmags[:, i], dphas[:, i], 0)
            sl = dcat[:, 256 * i:256 * (i + 1)]
            ops.small_linear_bwd(dre, lk["dft_re"], dx=sl)
            ops.small_linear_bwd(dim, lk["dft_im"], dx=sl)
        colsum(dcat[:, 768:], "texts")
        ops.grouped1x1_bwd_w(dcat[:, 0:256], S.ctx, gw("lconv.weight"))
        ops.grouped1x1_bwd_w(dcat[:, 256:512], S.dom_i, gw("dconv.weight"))
        ops.grouped1x1_bwd_w(dcat[:, 512:768], S.flo_i, gw("fconv.weight"))

    @ops.on_own_device
    @torch.no_grad()
    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int], encoder_hidden_states,
                domain_features, flow_features,
                down_block_additional_residuals: Optional[Tuple[torch.Tensor]] = None,
                mid_block_additional_residual: Optional[torch.Tensor] = None, return_dict: bool = True,
                added_time_ids: torch.Tensor = None):
        rows, g = self.forward_rows(sample, timestep, encoder_hidden_states, domain_features, flow_features,
                                    down_block_additional_residuals=down_block_additional_residuals,
                                    mid_block_additional_residual=mid_block_additional_residual,
                                    added_time_ids=added_time_ids)
        out = ops.unpack_output(rows, g.B, g.F, self.config.out_channels, g.H, g.W)
        if out.dtype != sample.dtype:
            out = out.to(sample.dtype)
        if not return_dict:
            return (out,)
        return UNetSpatioTemporalConditionOutput(sample=out)


# --------------------------------------------------------------------------------------------------- ControlNet
class ControlNetConditioningEmbeddingSVD(M.Container):
    def __init__(self, conditioning_embedding_channels, conditioning_channels=3,
                 block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = M.Conv2d(conditioning_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(M.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(M.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = M.Conv2d(block_out_channels[-1], conditioning_embedding_channels, 3, padding=1)
        nn.init.zeros_(self.conv_out.weight)
        nn.init.zeros_(self.conv_out.bias)


def _zero_conv(c):
    m = M.Conv2d(c, c, 1)
    nn.init.zeros_(m.weight)
    nn.init.zeros_(m.bias)
    return m


class ControlNetSDVModel(_Base):
    def __init__(self, sample_size=None, in_channels=8, out_channels=4,
                 down_block_types=SVD_XT_CONFIG["down_block_types"], block_out_channels=(320, 640, 1280, 1280),
                 addition_time_embed_dim=256, projection_class_embeddings_input_dim=768, layers_per_block=2,
                 cross_attention_dim=1024, transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20),
                 num_frames=25, conditioning_channels=3, conditioning_embedding_out_channels=(16, 32, 96, 256),
                 time_context_order="hw_major_0272"):
        super().__init__()
        n = len(down_block_types)
        if len(block_out_channels) != n:
            raise ValueError(f"Must provide the same number of `block_out_channels` as `down_block_types`. "
                             f"`block_out_channels`: {block_out_channels}. `down_block_types`: {down_block_types}.")
        self.config = SimpleNamespace(
            sample_size=sample_size, in_channels=in_channels, out_channels=out_channels,
            down_block_types=tuple(down_block_types), block_out_channels=tuple(block_out_channels),
            addition_time_embed_dim=addition_time_embed_dim,
            projection_class_embeddings_input_dim=projection_class_embeddings_input_dim,
            layers_per_block=layers_per_block, cross_attention_dim=cross_attention_dim,
            transformer_layers_per_block=transformer_layers_per_block, num_attention_heads=num_attention_heads,
            num_frames=num_frames, conditioning_channels=conditioning_channels,
            conditioning_embedding_out_channels=tuple(conditioning_embedding_out_channels),
            time_context_order=time_context_order)
        heads, xdim, lpb, tlpb, temb = self._build_encoder(
            in_channels, down_block_types, block_out_channels, layers_per_block, transformer_layers_per_block,
            num_attention_heads, cross_attention_dim, addition_time_embed_dim, projection_class_embeddings_input_dim)
        self.controlnet_cond_embedding = ControlNetConditioningEmbeddingSVD(
            block_out_channels[0], conditioning_channels, tuple(conditioning_embedding_out_channels))
        self.controlnet_down_blocks = nn.ModuleList([_zero_conv(block_out_channels[0])])
        for i in range(n):
            last = i == n - 1
            for _ in range(lpb[i] + (0 if last else 1)):
                self.controlnet_down_blocks.append(_zero_conv(block_out_channels[i]))
        self.controlnet_mid_block = _zero_conv(block_out_channels[-1])
        self.mid_block = M.UNetMidBlockSpatioTemporal(block_out_channels[-1], temb, 1, tlpb[-1], heads[-1], xdim[-1])
        self._packed = None
        self._cn = None

    def invalidate(self):
        super().invalidate()
        self._cn = None

    @classmethod
    def from_unet(cls, unet, controlnet_conditioning_channel_order: str = "rgb",
                  conditioning_embedding_out_channels=(16, 32, 96, 256), load_weights_from_unet: bool = True,
                  conditioning_channels: int = 3):
        c = unet.config
        net = cls(in_channels=c.in_channels, down_block_types=c.down_block_types,
                  block_out_channels=c.block_out_channels, addition_time_embed_dim=c.addition_time_embed_dim,
                  transformer_layers_per_block=c.transformer_layers_per_block,
                  cross_attention_dim=c.cross_attention_dim, num_attention_heads=c.num_attention_heads,
                  num_frames=c.num_frames, sample_size=c.sample_size, layers_per_block=c.layers_per_block,
                  projection_class_embeddings_input_dim=c.projection_class_embeddings_input_dim,
                  conditioning_channels=conditioning_channels,
                  conditioning_embedding_out_channels=conditioning_embedding_out_channels,
                  time_context_order=getattr(c, "time_context_order", "hw_major_0272"))
        if load_weights_from_unet:
            for name in ("conv_in", "time_embedding", "add_embedding", "down_blocks", "mid_block"):
                getattr(net, name).load_state_dict(getattr(unet, name).state_dict())
        return net.to(unet.device)

    # channels of the packed condition frames (2 = flow, 3 = depth / rgb).  Padding to one full 64-channel k-block is the
    # FASTER choice on this path: with 8 channels (16 bytes per pixel) the TMA box is 7/8 out-of-bounds fill and the conv
    # takes 6.2 ms instead of 2.3 ms at 28 x 576 x 1024 (profiles/r02d_thin_conv.txt; ncu: every warp waits on the load
    # barriers, DRAM 1 %), which outweighs the 1.2 ms saved in the pack kernel
    COND_CPAD = 64

    def _cn_pack(self):
        if getattr(self, "_cn", None) is None:
            ce = self.controlnet_cond_embedding
            convs = [ce.conv_in] + list(ce.blocks) + [ce.conv_out]
            packed = []
            cin_pad = self.COND_CPAD
            if ce.conv_in.in_channels > cin_pad:
                raise ValueError(f"conditioning_channels > {cin_pad} is not supported")
            for i, cv in enumerate(convs):
                cout = cv.out_channels
                cout_pad = (cout + 15) // 16 * 16
                # the bandwidth-bound thin layers at pixel resolution have their own kernels (csrc/thin_conv.cu): the stem
                # reads the fp32 condition frames directly, the stride-1 16 -> 16 / 32 -> 32 convs run on mma.sync tiles
                if i == 0 and cout == 16 and cv.in_channels <= 4 and not os.environ.get("LKGD_NO_THIN_CONV"):
                    packed.append(("stem", _f32(cv.weight), _f32(cv.bias), 1, cout_pad))
                elif i > 0 and cv.stride[0] == 1 and (cv.in_channels, cout) in ((16, 16), (32, 32), (16, 32)) \
                        and cin_pad == cv.in_channels and not os.environ.get("LKGD_NO_THIN_CONV"):
                    w9 = cv.weight.detach().float().permute(2, 3, 0, 1).reshape(9, cout, cv.in_channels)
                    packed.append(("thin", w9.to(bf16).contiguous(), _f32(cv.bias), 1, cout_pad))
                else:
                    w, b = _conv3x3_weight(cv, cin_pad=cin_pad, cout_pad=cout_pad)
                    packed.append((w, b, cv.stride[0], cout_pad))
                cin_pad = cout_pad
            zero = [(cv.weight.detach().reshape(cv.out_channels, cv.in_channels).to(bf16).contiguous(), _f32(cv.bias))
                    for cv in list(self.controlnet_down_blocks) + [self.controlnet_mid_block]]
            self._cn = (packed, zero)
        return self._cn

    def _encode(self, x, g, timestep, encoder_hidden_states, added_time_ids, controlnet_cond, bf16_skips, cond_repeat=1,
                ctx_t=None):
        """Condition encoder (pixel resolution, models/controlnet_sdv.py:98-119) + the copied UNet encoder + mid block.
        ``cond_repeat``: ``controlnet_cond`` holds batch / cond_repeat samples and every one conditions ``cond_repeat``
        batch entries (the pipeline feeds the SAME condition frames to both classifier-free-guidance halves,
        pipeline...controlnet.py:547-550): the seven pixel-resolution convs then run once instead of twice, only the last
        (latent-resolution) conv is issued per copy, writing straight into its slice of the stem tensor."""
        pk = self.packed()
        convs, zero = self._cn_pack()
        emb = pk.time_embedding(self._timestep_tensor(timestep, x), added_time_ids.to(x.device))
        # ctx_t: the contexts of the WHOLE guidance batch when this forward holds one half of it (CFG pair split): the
        # temporal cross-attention of diffusers 0.27.2 indexes them by row, exactly as in the UNet
        cond = Conditioning(pk, emb, encoder_hidden_states.to(torch.float32).contiguous(),
                            None if ctx_t is None else ctx_t.to(torch.float32).contiguous())
        stem_add = None
        if controlnet_cond is not None:
            if controlnet_cond.ndim != 5:
                raise ValueError("controlnet_cond must be [batch, frames, channels, height, width]")
            b_, f_, cc, hc, wc = controlnet_cond.shape
            if b_ * cond_repeat * f_ != g.BF:
                raise ValueError("controlnet_cond batch x frames does not match the sample")
            hh, ww = hc, wc
            e = None
            for entry in convs[:-1]:
                if entry[0] == "stem":           # (kind, fp32 weight, bias, ...): reads the fp32 frames directly
                    e = ops.cond_conv_in(controlnet_cond.reshape(b_ * f_, cc, hc, wc), entry[1], entry[2])
                    continue
                if entry[0] == "thin":           # (kind, bf16 tap-major weight, bias, ...)
                    e = ops.thin_conv3x3(e, entry[1], entry[2], b_ * f_, hh, ww, silu=True)
                    continue
                w, b, stride, _ = entry
                if e is None:
                    e = ops.pack_input(controlnet_cond, 1.0, None, N=b_, Cpad=self.COND_CPAD)
                e = ops.gemm(e, w, mode=A_CONV3X3, conv=(b_ * f_, hh, ww, stride), bias=b, act=ACT_SILU)
                if stride == 2:
                    hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
            if (hh, ww) != (g.H, g.W):
                raise ValueError("controlnet_cond resolution must be 8x the latent resolution")
            w, b, _, _ = convs[-1]
            rows = b_ * f_ * hh * ww
            stem_add = torch.empty((cond_repeat * rows, w.shape[0]), device=e.device, dtype=bf16)
            for r in range(cond_repeat):
                ops.gemm(e, w, mode=A_CONV3X3, conv=(b_ * f_, hh, ww, 1), bias=b, out=stem_add[r * rows:(r + 1) * rows])
        return pk.encoder(x, g, cond, stem_add=stem_add, bf16_skips=bf16_skips), zero

    @ops.on_own_device
    @torch.no_grad()
    def forward_packed(self, x: torch.Tensor, g: Geom, timestep, encoder_hidden_states: torch.Tensor,
                       added_time_ids: torch.Tensor, controlnet_cond: Optional[torch.Tensor] = None,
                       conditioning_scale: float = 1.0, cond_repeat: int = 1):
        """Engine-layout forward: returns (list of 12 ``ChannelsLast`` residuals, ``ChannelsLast`` mid residual)."""
        ops.STATS_ARENA.begin(x.device)          # one zeroed buffer for this forward's fused GroupNorm statistics
        (xm, skips, geoms, gm), zero = self._encode(x, g, timestep, encoder_hidden_states, added_time_ids,
                                                    controlnet_cond, bf16_skips=True, cond_repeat=cond_repeat)
        s = float(conditioning_scale)
        down = [ChannelsLast(ops.gemm(skb, w, bias=b, s0=s), gs.BF, gs.H, gs.W)
                for (_, skb), (w, b), gs in zip(skips, zero[:-1], geoms)]
        mid = ChannelsLast(ops.gemm(xm[1], zero[-1][0], bias=zero[-1][1], s0=s), gm.BF, gm.H, gm.W)
        return down, mid

    @torch.no_grad()
    def inject_packed(self, x: torch.Tensor, g: Geom, timestep, encoder_hidden_states: torch.Tensor,
                      added_time_ids: torch.Tensor, controlnet_cond: Optional[torch.Tensor], conditioning_scale: float,
                      unet_skips: List[torch.Tensor], multipliers: Sequence[int], unet_mid: torch.Tensor,
                      cond_repeat: int = 1, ctx_t: Optional[torch.Tensor] = None) -> torch.Tensor:
        """The fused form of ``forward_packed`` + the UNet's residual adds: every zero conv writes
        ``unet_skip += m_i * scale * (W skip_cn + b)`` in place in its epilogue (fp32 residual read, fused GroupNorm
        statistics of the sum where a 128-row tile stays inside one frame), the mid zero conv does the same on the UNet's
        mid sample, which is returned.  Runs inside the UNet's forward: shares its statistics arena (no ``begin``).
        zip truncation as in the reference (:453-462): extra skips / residuals are ignored."""
        (xm, skips, geoms, gm), zero = self._encode(x, g, timestep, encoder_hidden_states, added_time_ids,
                                                    controlnet_cond, bf16_skips=True, cond_repeat=cond_repeat,
                                                    ctx_t=ctx_t)
        s = float(conditioning_scale)
        for i, ((_, skb), (w, b), gs, m, us) in enumerate(zip(skips, zero[:-1], geoms, multipliers, unet_skips)):
            if us.shape != (gs.M, w.shape[0]):
                raise ValueError("ControlNet and UNet skip shapes differ")
            ops.gemm(skb, w, bias=b, s0=s * float(m), res1=us, s1=1.0, out=us, out_f32=True,
                     gn_rows=gs.HW if gs.HW % 128 == 0 else 0)
        ops.gemm(xm[1], zero[-1][0], bias=zero[-1][1], s0=s, res1=unet_mid, s1=1.0, out=unet_mid, out_f32=True,
                 gn_rows=gm.HW if gm.HW % 128 == 0 else 0)
        return unet_mid

    @ops.on_own_device
    @torch.no_grad()
    def forward(self, sample: torch.FloatTensor, timestep: Union[torch.Tensor, float, int],
                encoder_hidden_states: torch.Tensor, added_time_ids: torch.Tensor,
                controlnet_cond: torch.FloatTensor = None, image_only_indicator: Optional[torch.Tensor] = None,
                return_dict: bool = True, guess_mode: bool = False, conditioning_scale: float = 1.0,
                output_layout: str = "nchw"):
        """``image_only_indicator`` and ``guess_mode`` are accepted and ignored like the reference (quirk D7).
        ``output_layout="nhwc"`` returns ``ChannelsLast`` residuals for the UNet's fast path."""
        x, g = self._pack_sample(sample, self.packed())
        down, mid = self.forward_packed(x, g, timestep, encoder_hidden_states, added_time_ids, controlnet_cond,
                                        conditioning_scale)
        if output_layout == "nchw":
            down = [d.to_nchw().to(sample.dtype) for d in down]
            mid = mid.to_nchw().to(sample.dtype)
        elif output_layout != "nhwc":
            raise ValueError("output_layout must be 'nchw' or 'nhwc'")
        if not return_dict:
            return (down, mid)
        return ControlNetOutput(down_block_res_samples=down, mid_block_res_sample=mid)
