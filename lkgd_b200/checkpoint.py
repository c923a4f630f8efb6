"""``checkpoint-<step>`` directories of a LoRA fine-tuning run, laid out as the reference's loop writes and resumes them
(SURVEY.md 8f, N4 second half).

Reference write path (``train_models/train_svd_lora.py:1702-1748``), every 40 optimizer steps on the main process:
rotate old ``checkpoint-*`` directories down to ``checkpoints_total_limit - 1`` (:1706-1730), then
``accelerator.save_state(<output_dir>/checkpoint-<global_step>)`` (:1733-1734) and
``save_lora_weights(<checkpoint>/<lora_name>/pytorch_lora_weights.safetensors)`` (:1735-1747).
Reference resume path (:1364-1387): ``--resume_from_checkpoint latest`` picks the directory with the largest step,
``accelerator.load_state`` restores optimizer / scheduler / RNG and - through the registered ``load_model_hook``
(:1157-1177) - the adapter file; ``global_step`` / ``first_epoch`` / ``resume_step`` are derived from the directory name.

``accelerate`` itself is un-vendored; its ``save_state`` file set is restated (accelerate ``checkpointing.py``:
``optimizer.bin`` = ``torch.save(optimizer.state_dict())``, ``scheduler.bin`` = LR-scheduler ``state_dict()``,
``random_states_<rank>.pkl`` = python / numpy / torch / cuda RNG states; the full ``model.safetensors`` of the frozen
UNet is optional here - the trainable state is the adapter file).  ``optimizer.bin`` has the exact shape of
``torch.optim.AdamW.state_dict()`` over the reference's parameter order (``filter(requires_grad, unet.parameters())``,
:1179, == ``train_svd_lora_train.txt``), so a reference run can resume from a lkgd_b200 checkpoint and vice versa.

Pure host-side bookkeeping - no kernels."""
from __future__ import annotations

import os
import pickle
import random
import shutil
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from . import lora_io

OPTIMIZER_NAME, SCHEDULER_NAME, RNG_NAME = "optimizer.bin", "scheduler.bin", "random_states_{}.pkl"


def list_checkpoints(output_dir: str) -> List[str]:
    """``checkpoint-*`` directory names sorted by step (train_svd_lora.py:1708-1712, :1370-1373)."""
    if not os.path.isdir(output_dir):
        return []
    dirs = [d for d in os.listdir(output_dir) if d.startswith("checkpoint")]
    return sorted(dirs, key=lambda x: int(x.split("-")[1]))


def latest_checkpoint(output_dir: str) -> Optional[str]:
    dirs = list_checkpoints(output_dir)
    return dirs[-1] if dirs else None


def rotate_checkpoints(output_dir: str, checkpoints_total_limit: Optional[int]) -> List[str]:
    """Before a new checkpoint is written at most ``checkpoints_total_limit - 1`` may remain (:1706-1730).  Returns the
    removed directory names."""
    if checkpoints_total_limit is None:
        return []
    ckpts = list_checkpoints(output_dir)
    removed = []
    if len(ckpts) >= checkpoints_total_limit:
        for name in ckpts[0:len(ckpts) - checkpoints_total_limit + 1]:
            try:
                shutil.rmtree(os.path.join(output_dir, name))
                removed.append(name)
            except OSError:
                pass
    return removed


def resume_position(checkpoint_name: str, gradient_accumulation_steps: int, num_update_steps_per_epoch: int
                    ) -> Tuple[int, int, int]:
    """(global_step, first_epoch, resume_step) exactly as the reference derives them (:1381-1387)."""
    global_step = int(os.path.basename(checkpoint_name.rstrip("/")).split("-")[1])
    resume_global_step = global_step * gradient_accumulation_steps
    first_epoch = global_step // num_update_steps_per_epoch
    resume_step = resume_global_step % (num_update_steps_per_epoch * gradient_accumulation_steps)
    return global_step, first_epoch, resume_step


# ------------------------------------------------------------------------------------------------- optimizer state
def adamw_state_dict(trainer) -> Dict:
    """``torch.optim.AdamW(lora_layers, lr, betas, weight_decay, eps).state_dict()`` rebuilt from the trainer's flat fp32
    moment buffers: parameter i is the i-th trainable tensor in the module's ``named_parameters()`` order."""
    entries = trainer.state_tensors()
    state = {}
    if trainer.step_count > 0:
        for i, (_, _, m, v) in enumerate(entries):
            state[i] = {"step": torch.tensor(float(trainer.step_count)), "exp_avg": m.detach().cpu().clone(),
                        "exp_avg_sq": v.detach().cpu().clone()}
    group = {"lr": trainer.lr, "betas": tuple(trainer.betas), "eps": trainer.eps, "weight_decay": trainer.wd,
             "amsgrad": False, "foreach": None, "maximize": False, "capturable": False, "differentiable": False,
             "fused": None, "params": list(range(len(entries)))}
    return {"state": state, "param_groups": [group], "param_names": [e[0] for e in entries]}


def load_adamw_state_dict(trainer, sd: Dict) -> None:
    entries = trainer.state_tensors()
    if len(sd["param_groups"]) != 1 or len(sd["param_groups"][0]["params"]) != len(entries):
        raise ValueError(f"optimizer state covers {len(sd['param_groups'][0]['params'])} parameters, the trainer has "
                         f"{len(entries)}")
    names = sd.get("param_names")
    if names is not None and list(names) != [e[0] for e in entries]:
        raise ValueError("optimizer state was written for a different set / order of trainable parameters")
    g = sd["param_groups"][0]
    trainer.lr, trainer.betas, trainer.eps, trainer.wd = g["lr"], tuple(g["betas"]), g["eps"], g["weight_decay"]
    step = 0
    with torch.no_grad():
        for i, (name, _, m, v) in enumerate(entries):
            st = sd["state"].get(i)
            if st is None:
                m.zero_()
                v.zero_()
                continue
            if tuple(st["exp_avg"].shape) != tuple(m.shape):
                raise ValueError(f"{name}: moment shape {tuple(st['exp_avg'].shape)} != parameter shape {tuple(m.shape)}")
            m.copy_(st["exp_avg"].to(m.device, m.dtype))
            v.copy_(st["exp_avg_sq"].to(v.device, v.dtype))
            step = max(step, int(float(st["step"])))
    trainer.step_count = step


# ------------------------------------------------------------------------------------------------- save / load
def save_state(trainer, output_dir: str, global_step: int, lora_name: str = "default",
               checkpoints_total_limit: Optional[int] = None, process_index: int = 0, lr_scheduler_state: Optional[Dict] = None,
               save_full_model: bool = False) -> str:
    """Writes ``<output_dir>/checkpoint-<global_step>`` (rotation first) and returns its path."""
    os.makedirs(output_dir, exist_ok=True)
    rotate_checkpoints(output_dir, checkpoints_total_limit)
    path = os.path.join(output_dir, f"checkpoint-{global_step}")
    os.makedirs(path, exist_ok=True)
    torch.save(adamw_state_dict(trainer), os.path.join(path, OPTIMIZER_NAME))
    sched = lr_scheduler_state if lr_scheduler_state is not None else \
        {"base_lrs": [trainer.lr], "last_epoch": global_step, "_step_count": global_step + 1, "_last_lr": [trainer.lr]}
    torch.save(sched, os.path.join(path, SCHEDULER_NAME))
    states = {"step": global_step, "random_state": random.getstate(), "numpy_random_seed": np.random.get_state(),
              "torch_manual_seed": torch.get_rng_state()}
    if torch.cuda.is_available():
        states["torch_cuda_manual_seed"] = torch.cuda.get_rng_state_all()
    with open(os.path.join(path, RNG_NAME.format(process_index)), "wb") as f:
        pickle.dump(states, f)
    lora_io.save_lora_weights(trainer.unet, os.path.join(path, lora_name), adapter_name=lora_name)   # :1735-1747
    if save_full_model:
        from safetensors.torch import save_file
        save_file({k: v.detach().cpu().contiguous() for k, v in trainer.unet.state_dict().items()},
                  os.path.join(path, "model.safetensors"), metadata={"format": "pt"})
    return path


def load_state(trainer, path: str, lora_name: str = "default", process_index: int = 0, restore_rng: bool = True) -> int:
    """``accelerator.load_state(path)`` + the reference's ``load_model_hook`` (:1157-1172): adapter tensors, Adam moments,
    step count and RNG states.  Returns the global step encoded in the directory name."""
    # the trainer OWNS the kernel-ready pack of its UNet (its gradient slots are keyed by the packed layers and `repack`
    # refreshes their bf16 operands in place from the flat fp32 parameters): loading adapter tensors must not drop it
    pack = getattr(trainer.unet, "_packed", None)
    res = lora_io.load_lora_weights(trainer.unet, os.path.join(path, lora_name), adapter_name=lora_name, strict=True)
    if pack is not None:
        trainer.unet._packed = pack
    if not res["loaded"]:
        raise ValueError(f"no adapter tensors found under {path}/{lora_name}")
    opt = os.path.join(path, OPTIMIZER_NAME)
    if os.path.isfile(opt):
        load_adamw_state_dict(trainer, torch.load(opt, map_location="cpu", weights_only=False))
    rng = os.path.join(path, RNG_NAME.format(process_index))
    if restore_rng and os.path.isfile(rng):
        with open(rng, "rb") as f:
            states = pickle.load(f)
        random.setstate(states["random_state"])
        np.random.set_state(states["numpy_random_seed"])
        torch.set_rng_state(states["torch_manual_seed"])
        if torch.cuda.is_available() and "torch_cuda_manual_seed" in states and \
                len(states["torch_cuda_manual_seed"]) == torch.cuda.device_count():
            torch.cuda.set_rng_state_all(states["torch_cuda_manual_seed"])
    if hasattr(trainer, "repack"):
        trainer.repack()        # the bf16 GEMM operands follow the restored fp32 master parameters
    return int(os.path.basename(path.rstrip("/")).split("-")[1])


def resume_from_checkpoint(trainer, output_dir: str, resume: str = "latest", lora_name: str = "default",
                           gradient_accumulation_steps: int = 1, num_update_steps_per_epoch: int = 1):
    """The reference's ``--resume_from_checkpoint`` block (:1364-1387).  Returns (global_step, first_epoch, resume_step)
    or None when there is nothing to resume from."""
    name = latest_checkpoint(output_dir) if resume == "latest" else os.path.basename(resume)
    if name is None or not os.path.isdir(os.path.join(output_dir, name)):
        return None
    load_state(trainer, os.path.join(output_dir, name), lora_name=lora_name)
    return resume_position(name, gradient_accumulation_steps, num_update_steps_per_epoch)
