#!/usr/bin/env python
"""bench.py - headline benchmark of the LKGD / SVD-XT denoise hot path (BASELINE.json).

Workload (config.workload = "C3"): one classifier-free-guided Euler-Karras denoise step of the SVD-XT
spatio-temporal UNet, 25 frames at 576x1024 (72x128 latents), CFG batch 2, LoRA r=64 on the temporal attn1
q/k/v projections and the latent-knowledge (LKGD) cross-attention conditioning; random-init weights, synthetic
inputs.  A "step" = fused pack kernel -> UNet -> fused CFG + Euler kernel.

    python bench.py --gpus N --steps K --warmup W           # ours  (N>1: launched under torchrun, 1 rank/GPU)
    python bench.py --impl reference ...                    # the reference's CPU path (oracle port) on host cores

Multi-GPU: independent initial-frame samples are sharded one per rank (no collective on the data path, weak
scaling); value = steps all ranks completed / max-over-ranks device time.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # name: (unet config key, frames, h, w, lora rank, lkgd)           BASELINE.json configs[i]
    "C3": ("svd_xt", 25, 72, 128, 64, True),        # [2] headline: SVD-XT 25f + LoRA r=64 + LKGD conditioning
    "C2": ("svd_xt", 14, 72, 128, 0, False),        # [1] SVD 14f
    "C4": ("svd_xt", 14, 72, 128, 0, False),        # [3] C2 + ControlNetSDVModel (flow condition, 2 channels)
    "C4d": ("svd_xt", 14, 72, 128, 0, False),       # [3] ... depth condition, 3 channels
    "C1": ("reduced", 8, 32, 32, 0, False),         # [0] reduced config
    # LoRA fine-tuning step (BASELINE.json configs[4]): forward + backward + clip + AdamW, 14 frames 320x512, b=1 / GPU
    "C5": ("svd_xt", 14, 40, 64, 64, True),
}
CONTROLNET_CHANNELS = {"C4": 2, "C4d": 3}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], tf_burst=p["bf16_tflops"], tf_sustained=p["bf16_tflops_sustained"],
                    source="measured (MEASURED_PEAKS.json)")
    return dict(hbm_gbs=6650.0, tf_burst=1590.0, tf_sustained=1400.0, source="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, pw = [], [], set(), []
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_inputs(S, F, h, w, lkgd, device="cpu", seed=0, xdim=1024):
    """latents ~ N(0,1) (scaled by init_noise_sigma later), image_latents ~ N(0,1) with a zero uncond
    half, CLIP embedding ~ N(0,1) with a zero uncond half, domain / flow features ~ N(0,1)."""
    g = torch.Generator().manual_seed(seed)
    noise = torch.randn(S, F, 4, h, w, generator=g)
    cond_lat = torch.randn(S, 1, 4, h, w, generator=g).repeat(1, F, 1, 1, 1)
    image_latents = torch.cat([torch.zeros_like(cond_lat), cond_lat])
    emb = torch.randn(S, 1, xdim, generator=g)
    image_embeddings = torch.cat([torch.zeros_like(emb), emb])
    extra = (torch.randn(1, 1, 1000, generator=g), torch.randn(1, 1, 1000, generator=g)) if lkgd else ()
    return noise, image_latents, image_embeddings, extra


def init_weights_(model, seed=0, zero_conv_std=None):
    """Default-initialiser-equivalent random weights generated directly on the GPU (no checkpoints exist here);
    zero-inits that would hide work are overridden as in SURVEY 8(d)."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if zero_conv_std and ("controlnet_down_blocks" in n or "controlnet_mid_block" in n
                                  or "controlnet_cond_embedding.conv_out" in n):
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * zero_conv_std)   # zero convs
            elif n.endswith("mix_factor"):
                p.copy_(torch.rand(p.shape, generator=g, device="cuda") * 2 - 1)
            elif "lora_B" in n or n.startswith("quaternion_lora_texts"):
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") * 0.02)
            elif "lora_A" in n:
                p.copy_(torch.randn(p.shape, generator=g, device="cuda") / p.shape[0])
            elif p.ndim >= 2:
                fan_in = p[0].numel()
                p.copy_((torch.rand(p.shape, generator=g, device="cuda") * 2 - 1) * fan_in ** -0.5)
            elif ".norm" in n and n.endswith("weight") or "conv_norm_out.weight" in n:
                p.fill_(1.0)
            elif n.endswith("bias") and (".norm" in n or "conv_norm_out" in n):
                p.zero_()
            else:
                p.copy_((torch.rand(p.shape, generator=g, device="cuda") * 2 - 1) * 0.02)


def build_ours(workload, device):
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.scheduler import SVD_SCHEDULER_CONFIG, EulerDiscreteScheduler
    from lkgd_b200.unet import (REDUCED_CONFIG, SVD_XT_CONFIG, UNetSpatioTemporalConditionControlNetModel,
                                UNetSpatioTemporalConditionModel)
    key, F, h, w, rank, lkgd = CONFIGS[workload]
    cfg = dict(SVD_XT_CONFIG if key == "svd_xt" else REDUCED_CONFIG, num_frames=F)
    if lkgd and cfg["cross_attention_dim"] != 1024:
        cfg["cross_attention_dim"] = 1024
    cls = UNetSpatioTemporalConditionModel if lkgd else UNetSpatioTemporalConditionControlNetModel
    with torch.device("meta"):
        unet = cls(**cfg)
        if rank:
            unet.add_lora(rank)
    unet = unet.to_empty(device=device)
    init_weights_(unet)
    unet.invalidate()
    cn = None
    if workload in CONTROLNET_CHANNELS:
        from lkgd_b200.unet import ControlNetSDVModel
        with torch.device("meta"):
            cn = ControlNetSDVModel(**{k: v for k, v in cfg.items() if k != "up_block_types"},
                                    conditioning_channels=CONTROLNET_CHANNELS[workload])
        cn = cn.to_empty(device=device)
        init_weights_(cn, seed=1, zero_conv_std=0.02)
        cn.invalidate()
    pipe = StableVideoDiffusionPipeline(unet, EulerDiscreteScheduler(**SVD_SCHEDULER_CONFIG), controlnet=cn)
    return pipe, cfg, (F, h, w, rank, lkgd)


def workload_desc(workload):
    key, F, h, w, lrank, lkgd = CONFIGS[workload]
    cn = CONTROLNET_CHANNELS.get(workload)
    return (f"{workload}: SVD-XT CFG denoise step, {F} frames 576x1024 ({h}x{w} latents), CFG batch 2, LoRA r={lrank} "
            f"folded in the temporal qkv GEMMs, LKGD cond={lkgd}"
            + (f", ControlNetSDVModel ({cn}-channel condition at 576x1024) + residual injection" if cn else "")
            + "; 1 sample per GPU")


def step_flops(workload, cfg, F, h, w, lrank, as_reference=False):
    """Algorithmic FLOPs of one denoise step.  ``as_reference``: count the KV-length-1 cross-attention as the reference
    executes it (to_q / to_k / per-token to_out); default: what the CUDA path executes."""
    from lkgd_b200.flops import controlnet_flops, unet_flops
    f = unet_flops(cfg, 2, F, h, w, lora_rank=lrank, count_dead_cross_attn=as_reference)["total"]
    if workload in CONTROLNET_CHANNELS:
        f += controlnet_flops(cfg, 2, F, h, w, CONTROLNET_CHANNELS[workload], count_dead_cross_attn=as_reference)["total"]
    return f


def synth_controlnet_cond(workload, F, h, w, seed=0):
    """controlnet_cond ~ U(-1, 1) [F, Cc, 8h, 8w] (the pipeline duplicates it for CFG)."""
    g = torch.Generator().manual_seed(1000 + seed)
    return torch.rand(F, CONTROLNET_CHANNELS[workload], 8 * h, 8 * w, generator=g) * 2 - 1


def gemm_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per gemm_tcgen05_kernel launch, averaged over the launches of one
    step, from the committed ncu capture (profiles/*gemm_traffic*.json; a profiler number is never timed here)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*gemm_traffic*.json")))
    if not files:
        return None
    try:
        return json.load(open(files[-1])).get("avg_dram_bytes_per_launch")
    except Exception:
        return None


def _build_oracle(workload, Fs):
    """fp32 CPU oracle of the workload's modules (UNet [+ LoRA] [+ ControlNet]) with random weights."""
    import oracle as O
    from lkgd_b200.unet import REDUCED_CONFIG, SVD_XT_CONFIG
    key, F, h, w, rank, lkgd = CONFIGS[workload]
    cfg = dict(SVD_XT_CONFIG if key == "svd_xt" else REDUCED_CONFIG, num_frames=Fs)
    if lkgd:
        cfg["cross_attention_dim"] = 1024
    cls = O.UNetSpatioTemporalConditionModel if lkgd else O.UNetSpatioTemporalConditionControlNetModel
    torch.manual_seed(0)
    with torch.device("meta"):
        model = cls(**cfg)
        if rank:
            O.add_lora(model, rank)
        cn = None
        if workload in CONTROLNET_CHANNELS:
            cn = O.ControlNetSDVModel(**{k: v for k, v in cfg.items() if k != "up_block_types"},
                                      conditioning_channels=CONTROLNET_CHANNELS[workload])
    mods = [model.to_empty(device="cpu")] + ([cn.to_empty(device="cpu")] if cn is not None else [])
    with torch.no_grad():
        for m in mods:
            for p in m.parameters():
                p.uniform_(-0.02, 0.02)
            m.eval()
    return cfg, mods[0], (mods[1] if cn is not None else None)


def cpu_oracle_rate(workload, max_seconds=40.0, iters=4, full_size=False):
    """Times the oracle (fp32 PyTorch, eager, all host threads) on the host cores.

    ``full_size=False`` (the `cpu_baseline` leg of our own line: ~10-30 s): a BOUNDED sample - the same full-width
    modules at 2 frames of 16x16 latents - scaled to the workload by algorithmic FLOPs; labelled ``extrapolated``.
    ``full_size=True`` (the `--impl reference` arm): ONE real step of the workload itself (25 / 14 frames, 72x128
    latents, CFG batch 2: ~160 / 90 / 123 TFLOP of eager fp32, minutes), nothing scaled."""
    import oracle as O
    key, F, h, w, rank, lkgd = CONFIGS[workload]
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    if full_size or key != "svd_xt":
        Fs, hs, ws = F, h, w
    else:
        Fs, hs, ws = 2, 16, 16          # sample: full-width SVD-XT UNet, CFG batch 2, 2 frames, 16x16 latents
    t0 = time.time()
    cfg, model, cn = _build_oracle(workload, Fs)
    build_s = time.time() - t0
    noise, img_lat, emb, extra = synth_inputs(1, Fs, hs, ws, lkgd, xdim=cfg["cross_attention_dim"])
    sched = O.EulerDiscreteScheduler(**O.scheduler.SVD_SCHEDULER_CONFIG)
    sched.set_timesteps(25)
    ids = O.add_time_ids_inference(6, 127, 0.02, 1)
    lat = noise * sched.init_noise_sigma
    kw = {}
    if cn is not None:
        cc = synth_controlnet_cond(workload, Fs, hs, ws).unsqueeze(0)
        kw = dict(controlnet=cn, controlnet_cond=torch.cat([cc, cc]))       # the pipeline duplicates it for CFG (D3)
    times = []
    t_start = time.time()
    for it in range(1 if full_size else max(2, iters)):
        sched._step_index = None
        t1 = time.time()
        O.denoise_loop(model, sched, lat, img_lat, emb, ids, 25, 1.0, 3.0, unet_extra_args=extra, max_steps=1, **kw)
        times.append(time.time() - t1)
        if time.time() - t_start > max_seconds:
            break
    t_step = min(times[1:]) if len(times) > 1 else times[0]
    f_sample = step_flops(workload, cfg, Fs, hs, ws, rank, as_reference=True)
    f_full = step_flops(workload, dict(cfg, num_frames=F), F, h, w, rank, as_reference=True)
    value = (1.0 / t_step) * (f_sample / f_full)
    what = (f"oracle (fp32 PyTorch eager, {cores} threads), full-width modules, one CFG denoise step at {Fs} frames "
            f"{hs}x{ws} latents: {t_step:.2f} s/step = {f_sample / t_step / 1e12:.3f} TFLOP/s")
    if full_size or key != "svd_xt":
        what += "; this IS the workload's size - measured, not scaled"
    else:
        what += (f"; EXTRAPOLATED to the {F}f {h}x{w} step by algorithmic FLOPs ({f_sample / 1e12:.2f} vs "
                 f"{f_full / 1e12:.1f} TFLOP as the reference executes them)")
    what += f"; model build {build_s:.0f} s not timed"
    return dict(value=value, unit="denoise_steps/s", cores=cores, kind="port", sample=what,
                extrapolated=not (full_size or key != "svd_xt"), timed_steps=len(times),
                seconds_per_timed_step=t_step), t_step


def train_flops(cfg, B, F, h, w, rank):
    """Algorithmic work of one LoRA training step: forward + data gradients of every GEMM / conv above the first
    adapter (~ the forward's dense FLOPs again) + attention backward (2.5x its forward) + LoRA weight gradients."""
    from lkgd_b200.flops import unet_flops
    f = unet_flops(cfg, B, F, h, w, lora_rank=rank, count_dead_cross_attn=False)
    dense = f["conv3x3"] + f["tconv"] + f["shortcut"] + f["proj"] + f["geglu_ff"] + f["lora"]
    attn = f["spatial_attn"] + f["temporal_attn"]
    return f["total"] + dense + 2.5 * attn + f["lora"]


def run_train(args):
    """--workload C5: one LoRA fine-tuning step per "step" (reference train_models/train_svd_lora.py:1445-1689)."""
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from lkgd_b200 import _lib, ops
    from lkgd_b200.training import LoraTrainer
    ops.device_check(local)
    pipe, cfg, (F, h, w, lrank, lkgd) = build_ours(args.workload, device)
    unet = pipe.unet
    tr = LoraTrainer(unet, lr=1e-4, world_size=world)
    B = 1
    g = torch.Generator().manual_seed(100 + rank)
    host = dict(lat=torch.randn(B, F, 4, h, w, generator=g).pin_memory(),
                noise=torch.randn(B, F, 4, h, w, generator=g).pin_memory(),
                cond=torch.randn(B, 4, h, w, generator=g).pin_memory(),
                ctx=torch.randn(B, 1, 1024, generator=g).pin_memory(),
                sig=torch.tensor([1.3] * B).pin_memory())
    if lkgd:   # domain / flow ViT features of the clip (train_svd_lora.py:1458-1470)
        host.update(dom=torch.randn(B, 1, 1000, generator=g).pin_memory(),
                    flo=torch.randn(B, 1, 1000, generator=g).pin_memory())
    ids = torch.tensor([[5.0, 0.02, 127.0]] * B, device=device)      # training order (utils/util.py:295, quirk F12)
    dev = {k: v.to(device) for k, v in host.items()}

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(d):
        extra = (d["dom"], d["flo"]) if lkgd else ()
        return tr.train_step(d["lat"], d["noise"], d["sig"], d["cond"], d["ctx"], ids, *extra)

    step(dev)                                   # eager: lazy weight preprocessing, one-time kernel attributes
    l0 = ops.launch_count()
    step(dev)
    launches_per_step = ops.launch_count() - l0  # a graph replay re-issues exactly the launches captured from this path
    if not args.no_graph:
        extra = (dev["dom"], dev["flo"]) if lkgd else ()
        tr.capture(dev["lat"], dev["noise"], dev["sig"], dev["cond"], dev["ctx"], ids, *extra)
    for _ in range(args.warmup):
        step(dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        loss = step(dev)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    # end to end: the batch comes from pinned host memory and the loss is read back every step
    h2d = sum(v.numel() * 4 for v in host.values())
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    for _ in range(args.steps):
        d = {k: v.to(device, non_blocking=True) for k, v in host.items()}
        loss_host = float(step(d))
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])
    if rank == 0:
        tr._graph = None                         # per-kernel event breakdown on the eager path
        if args.profiler_step:                   # ncu launch list of ONE eager training step (--profile-from-start off)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            step(dev)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        _lib.PROF.records, _lib.PROF.enabled = [], True
        step(dev)
        torch.cuda.synchronize()
        _lib.PROF.enabled = False
        by = {}
        for name, a, b, meta in _lib.PROF.records:
            d = by.setdefault(name, dict(ms=0.0, calls=0, flops=0.0))
            d["ms"] += a.elapsed_time(b)
            d["calls"] += 1
            if meta and "flops" in meta:
                d["flops"] += meta["flops"]
        total_ms = sum(d["ms"] for d in by.values())
        if args.profile_out:
            json.dump({"by_kernel": by, "gemm_calls": [dict(meta, ms=a.elapsed_time(b)) for n, a, b, meta in
                                                        _lib.PROF.records if n == "lkgd_gemm" and meta]},
                      open(args.profile_out, "w"), indent=1)
        pk = peaks()
        gm = by.get("lkgd_gemm", dict(ms=1e-9, calls=0, flops=0.0))
        achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
        flops = train_flops(cfg, B, F, h, w, lrank)
        line = {
            "metric": "train_steps_per_s", "value": args.steps * world / (ms * 1e-3), "unit": "train_steps/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"C5: SVD LoRA fine-tuning step (forward + backward + clip + AdamW), {F} frames 320x512 "
                                   f"({h}x{w} latents), batch 1 per GPU, LoRA r={lrank} on temporal attn1 q/k/v + the 29 latent-knowledge "
                                   f"'quaternion' tensors (LKGD UNet={lkgd}), "
                                   f"{'one flat NCCL all-reduce of the LoRA gradients per step' if world > 1 else 'single GPU'}",
                       "algorithmic_tflop_per_step": flops / 1e12,
                       "model_tflops_per_gpu": flops / 1e12 / (ms / args.steps * 1e-3),
                       "trainable_parameters": int(tr.flat_p.numel()), "loss": loss_host,
                       "cuda_graph": not args.no_graph,
                       "gpu_launches_note": "launches of our kernels issued per step on the eager path x steps; under "
                                            "CUDA-graph replay the same launches are re-issued by the graph",
                       "l2": "3 GB of bf16 weights (+ their transposed copies) stream through every step; >> 126 MB L2"},
            "e2e": {"value": args.steps * world / (ms_e2e * 1e-3), "unit": "train_steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 8},
            "gpu_launches": int(launches), "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (forward + data-gradient GEMMs / convs)",
                         "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                         "frac": achieved / pk["tf_sustained"], "traffic": None,
                         "peak_source": pk["source"] + ", sustained", "launches_per_step": gm["calls"],
                         "kernel_ms_per_step": gm["ms"], "share_of_step": gm["ms"] / total_ms,
                         "breakdown_ms": {k: round(v["ms"], 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}},
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    """The reference's CPU path on the host cores: ONE real, full-size denoise step of the oracle port (the reference
    itself cannot be installed here: diffusers 0.27.2 / peft / core_qnn are absent).  ``steps`` / ``warmup`` of the
    printed line say what was actually timed (1 / 0): a full-size step is minutes of CPU work, so K steps are not run;
    ``--ref-sample`` switches to the bounded, FLOP-scaled sample (seconds) and labels the line ``extrapolated``."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.workload == "C5":
        args.workload = "C3"
    full = not args.ref_sample
    cb, t_step = cpu_oracle_rate(args.workload, max_seconds=150.0, iters=3, full_size=full)
    line = {
        "impl": "reference", "metric": "denoise_steps_per_s", "value": cb["value"], "unit": "denoise_steps/s",
        "n_gpus": args.gpus, "steps": cb["timed_steps"] if full else args.steps, "warmup": 0 if full else args.warmup,
        "requested_steps": args.steps, "requested_warmup": args.warmup,
        "ms_per_step": 1000.0 / cb["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "extrapolated": cb["extrapolated"],
        "config": {"workload": workload_desc(args.workload),
                   "reference_arm": "the reference's CPU PyTorch path (fp32 oracle port; the reference itself cannot be "
                                    "installed: no diffusers/peft/core_qnn). " +
                                    ("ONE real full-size step timed (steps=1, warmup=0: minutes of CPU per step)" if full
                                     else "every step a bounded sample, FLOP-scaled (extrapolated)")},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "denoise_steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


_RESULT_FD = None


def guard_stdout():
    """The contract is ONE JSON line on stdout.  Libraries print there too (NCCL's "NCCL version ..." banner when the box
    exports NCCL_DEBUG=VERSION, seen on the 8-GPU boxes): everything written to fd 1 from here on goes to stderr, and
    `emit` writes the result line to the real stdout."""
    global _RESULT_FD
    if _RESULT_FD is None:
        sys.stdout.flush()
        _RESULT_FD = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict) -> None:
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_RESULT_FD, data)


def main():
    guard_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3", choices=list(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-sample", action="store_true",
                    help="--impl reference: time the bounded 2-frame 16x16 sample (seconds, FLOP-scaled) instead of one "
                         "real full-size step (minutes)")
    ap.add_argument("--no-extra-sections", action="store_true",
                    help="N >= 2: skip the CFG-pair-split and data-parallel C5 sections after the sample-sharded timing")
    ap.add_argument("--no-graph", action="store_true", help="C5: run the training step eagerly (no CUDA graph replay)")
    ap.add_argument("--profile-out", default=None, help="write the per-kernel event breakdown to this JSON file")
    ap.add_argument("--profiler-step", action="store_true",
                    help="bracket ONE extra step with cudaProfilerStart/Stop (for `ncu --profile-from-start off`)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "C5":
        return run_train(args)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)
    from lkgd_b200 import _lib, ops
    ops.device_check(local)

    pipe, cfg, (F, h, w, lrank, lkgd) = build_ours(args.workload, device)
    S = 1   # one initial-frame sample per GPU (weak scaling: every rank denoises its own sample)
    noise, img_lat, emb, extra = synth_inputs(S, F, h, w, lkgd, seed=rank, xdim=cfg["cross_attention_dim"])
    kw = dict(domain_features=extra[0], flow_features=extra[1]) if lkgd else {}
    if args.workload in CONTROLNET_CHANNELS:
        kw["controlnet_condition"] = synth_controlnet_cond(args.workload, F, h, w, seed=rank)
    st = pipe.prepare(emb, img_lat, num_frames=F, num_inference_steps=25, **kw)
    lat0 = (noise * pipe.scheduler.init_noise_sigma).to(device)
    n_sig = 25

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------------------------------------------------------- device-resident timing
    lat = lat0
    l0 = ops.launch_count()
    lat, _ = pipe.denoise_step(st, 0, lat)               # eager: lazy weight packing, one-time kernel attributes
    launches_per_step = ops.launch_count() - l0          # a CUDA-graph replay re-issues exactly these launches
    graphed = False
    if not args.no_graph and hasattr(pipe, "capture"):
        pipe.capture(st, lat0)
        graphed = True
    for i in range(args.warmup):
        lat, _ = pipe.denoise_step(st, i % n_sig, lat)
    barrier()
    sampler_all = ClockSampler(local)
    sampler_all.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_host0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        lat, _ = pipe.denoise_step(st, (args.warmup + i) % n_sig, lat)
    e1.record()
    host_ms = (time.perf_counter() - t_host0) * 1e3      # time the host spent ISSUING the K steps (launch path)
    barrier()
    ms = e0.elapsed_time(e1)
    launches = launches_per_step * args.steps
    finite = bool(torch.isfinite(lat).all())

    # ---------------------------------------------------------------- end to end through the public API
    pin_noise = lat0.cpu().pin_memory()
    pin_img, pin_emb = img_lat.pin_memory(), emb.pin_memory()
    pin_out = torch.empty_like(pin_noise).pin_memory()
    h2d = pin_noise.numel() * 4 + pin_img.numel() * 4 + pin_emb.numel() * 4
    d2h = pin_out.numel() * 4
    barrier()
    e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2.record()
    lat_d = torch.empty_like(lat0)
    for i in range(args.steps):
        # this step's inputs come from pinned host memory into the step's static device buffers (the CUDA graph reads
        # them in place); the result goes back to the host
        lat_d.copy_(pin_noise, non_blocking=True)
        st["image_latents"].copy_(pin_img, non_blocking=True)
        st["image_embeddings"].copy_(pin_emb, non_blocking=True)
        out, _ = pipe.denoise_step(st, (args.warmup + i) % n_sig, lat_d)
        pin_out.copy_(out, non_blocking=True)
    e3.record()
    barrier()
    ms_e2e = e2.elapsed_time(e3)

    if args.profiler_step and rank == 0:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
        pipe.denoise_step(st, 5, lat0, eager=True)
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStop()

    # per-rank device time and SM clock (attribution of the N > 1 efficiency: there is no data-path collective, so any
    # loss is the slowest rank's clock / host launch path - value uses the MAX over ranks)
    my_clk = sampler_all.stop()
    mine = {"rank": rank, "ms_per_step": ms / args.steps, "e2e_ms_per_step": ms_e2e / args.steps,
            "host_ms_per_step": host_ms / args.steps, "sm_mhz": my_clk.get("sm_mhz"),
            "power_w_max": my_clk.get("power_w_max"), "reasons": my_clk.get("reasons")}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)
    t = torch.tensor([ms, ms_e2e], device=device, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t[0]), float(t[1])

    # ---------------------------------------------------------------- per-kernel breakdown (one extra step, events)
    roof = None
    if rank == 0:
        _lib.PROF.records, _lib.PROF.enabled = [], True
        pipe.denoise_step(st, 5, lat0, eager=True)       # per-launch CUDA events need the eager path (no graph replay)
        torch.cuda.synchronize()
        _lib.PROF.enabled = False
        by = {}
        for name, a, b, meta in _lib.PROF.records:
            d = by.setdefault(name, dict(ms=0.0, calls=0, flops=0.0, bytes=0.0))
            d["ms"] += a.elapsed_time(b)
            d["calls"] += 1
            if meta and "flops" in meta:
                d["flops"] += meta["flops"]
            if meta and "bytes" in meta:
                d["bytes"] += meta["bytes"]
        total_ms = sum(d["ms"] for d in by.values())
        pk = peaks()
        gm = by.get("lkgd_gemm", dict(ms=1e-9, calls=0, flops=0.0))
        achieved = gm["flops"] / (gm["ms"] * 1e-3) / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tcgen05_kernel (GEMM + implicit-GEMM conv)",
                "achieved": achieved, "peak": pk["tf_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["tf_sustained"], "traffic": gemm_traffic(),
                "traffic_unit": "bytes per launch (dram read+write, ncu, averaged over the step's launches)",
                "peak_source": pk["source"] + ", sustained",
                "achieved_def": "algorithmic 2*M*N*K summed over the step's GEMM/conv launches / their summed CUDA-event durations",
                "launches_per_step": gm["calls"], "algorithmic_tflop_per_step": gm["flops"] / 1e12,
                "kernel_ms_per_step": gm["ms"], "share_of_step": gm["ms"] / total_ms,
                "breakdown_ms": {k: round(v["ms"], 3) for k, v in sorted(by.items(), key=lambda kv: -kv[1]["ms"])}}
        at = by.get("lkgd_attention")
        if at:
            roof["attention_tflops"] = at["flops"] / (at["ms"] * 1e-3) / 1e12
        # bandwidth-bound glue: algorithmic bytes (ops.* record them per call) / summed CUDA-event time, against the
        # measured HBM copy bandwidth
        roof["hbm_glue"] = {k: {"ms": round(v["ms"], 3), "gbs": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9, 1),
                                "frac_of_hbm_peak": round(v["bytes"] / (v["ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], 3)}
                            for k, v in by.items() if v["bytes"] > 0 and v["ms"] > 0}
        if args.profile_out:
            json.dump({"by_kernel": by, "gemm_calls": [dict(meta, ms=a.elapsed_time(b)) for n, a, b, meta in
                                                        _lib.PROF.records if n == "lkgd_gemm" and meta]},
                      open(args.profile_out, "w"), indent=1)

    # ---------------------------------------------------------------- N >= 2: the other multi-GPU paths of north_star
    extra_sections = {}
    if world >= 2 and not args.no_extra_sections and args.workload == "C3":
        extra_sections["cfg_pair_split"] = section_cfg_split(args, pipe, F, h, w, lkgd, ms / args.steps, device, rank, world)
        del pipe, st
        torch.cuda.empty_cache()
        extra_sections["c5_data_parallel_training"] = section_c5_dp(args, device, rank, world)
    if world == 1 and not args.no_extra_sections and args.workload == "C3":
        del pipe, st
        torch.cuda.empty_cache()
        try:
            extra_sections["vae_and_clip"] = section_vae_clip(device)
        except Exception as ex:   # the side sections must never sink the headline number
            extra_sections["vae_and_clip"] = {"failed": repr(ex)}

    if rank == 0:
        steps_total = args.steps * world
        value = steps_total / (ms * 1e-3)
        flops = step_flops(args.workload, cfg, F, h, w, lrank)                       # what the CUDA path executes
        flops_ref = step_flops(args.workload, cfg, F, h, w, lrank, as_reference=True)
        line = {
            "metric": "denoise_steps_per_s", "value": value, "unit": "denoise_steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_desc(args.workload),
                       "frames_per_s_per_gpu": F * args.steps / (ms * 1e-3),
                       "algorithmic_tflop_per_step": flops / 1e12,
                       "reference_executed_tflop_per_step": flops_ref / 1e12,
                       "flops_note": "algorithmic = executed by the CUDA path: the KV-length-1 cross-attention is ONE "
                                     "[C,D] mat-vec per batch element; the reference also runs to_q / to_k / a per-token "
                                     "to_out whose result does not depend on the query (SURVEY F7) - not counted as achieved",
                       "model_tflops_per_gpu": flops / 1e12 / (ms / args.steps * 1e-3),
                       "cuda_graph": graphed, "host_issue_ms_per_step": per_rank[0]["host_ms_per_step"],
                       "l2": "per-step working set (3 GB bf16 weights + multi-GB activations) >> 126 MB L2; "
                             "no explicit flush needed",
                       "output_finite": finite},
            "e2e": {"value": steps_total / (ms_e2e * 1e-3), "unit": "denoise_steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches),
            "gpu_launches_note": "launches of our kernels per step (counted on the eager step) x steps; a CUDA-graph "
                                 "replay re-issues exactly these launches",
            "clocks": my_clk,
            "per_rank": per_rank,
            "roofline": roof,
        }
        line.update(extra_sections)
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"], _ = cpu_oracle_rate(args.workload)
            except Exception as ex:  # the CPU leg must never sink the GPU number
                line["cpu_baseline"] = {"value": None, "unit": "denoise_steps/s", "cores": os.cpu_count(),
                                        "kind": "port", "sample": f"failed: {ex!r}"}
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def section_vae_clip(device):
    """The steps either side of the denoise loop at the headline sizes, random weights: temporal VAE decode of
    25 x 576 x 1024 frames in chunks of 8 (the reference scripts' decode_chunk_size, run_models/run_inference.py:295), VAE
    encode of one 576x1024 image, CLIP ViT-H/14 embedding of one image.  CUDA events, one warm-up pass each."""
    from lkgd_b200 import ops
    from lkgd_b200.clip import CLIP_VIT_H_14, CLIPVisionModelWithProjection
    from lkgd_b200.flops import vae_flops
    from lkgd_b200.vae import SVD_VAE_CONFIG, AutoencoderKLTemporalDecoder, decode_latents
    torch.manual_seed(0)
    vae = AutoencoderKLTemporalDecoder(**SVD_VAE_CONFIG).to(device)
    lat = torch.randn(1, 25, 4, 72, 128, device=device)
    img = torch.rand(1, 3, 576, 1024, device=device) * 2 - 1

    def timed(fn, reps):
        fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            out = fn()
        b.record()
        torch.cuda.synchronize()
        return a.elapsed_time(b) / reps, out

    l0 = ops.launch_count()
    dec_ms, frames = timed(lambda: decode_latents(vae, lat, 25, 8), 2)
    dec_launches = (ops.launch_count() - l0) // 3
    enc_ms, dist_ = timed(lambda: vae.encode(img).latent_dist.mode(), 3)
    fl = vae_flops(SVD_VAE_CONFIG, 25, 72, 128)
    out = {"vae_decode_ms": dec_ms, "frames": 25, "frame_size": [576, 1024], "decode_chunk_size": 8,
           "vae_decode_tflop": fl["decode"] / 1e12, "vae_decode_tflops": fl["decode"] / 1e12 / (dec_ms * 1e-3),
           "vae_decode_launches": int(dec_launches), "vae_decode_finite": bool(torch.isfinite(frames).all().item()),
           "vae_encode_ms": enc_ms, "vae_encode_tflops": fl["encode"] / 1e12 / (enc_ms * 1e-3)}
    del vae, frames
    torch.cuda.empty_cache()
    clip = CLIPVisionModelWithProjection(**CLIP_VIT_H_14).to(device)
    pv = torch.randn(1, 3, 224, 224, device=device)
    clip_ms, _ = timed(lambda: clip(pv).image_embeds, 3)
    out["clip_vit_h_ms"] = clip_ms
    out["note"] = ("one pipeline call = 1 CLIP + 1 VAE encode + 25 denoise steps + 1 chunked VAE decode; random weights, "
                   "inputs resident on the device")
    return out


def _rel(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / (b.norm() + 1e-30))


def section_cfg_split(args, pipe, F, h, w, lkgd, unsplit_ms, device, rank, world):
    """north_star: "the CFG cond/uncond pair ... split across GPUs, with no collective in the UNet".  Ranks (2k, 2k+1)
    denoise sample k together: rank 2k the unconditional half, 2k+1 the conditional half (batch 1 each), ONE exchange
    of the fp32 prediction per step - fused into the CFG + Euler kernel over peer memory where the box offers it (the
    kernel reads the partner's half over NVLink; the NCCL all-gather path is timed beside it), else an all-gather - and
    the combine kernel on both ranks (latents stay replicated).  Latency mode:
    reports ms/step against the unsplit step, the split-vs-unsplit difference and whether the pair's latents are equal."""
    import torch.distributed as dist
    from lkgd_b200.distributed import CFGPair
    if world % 2:
        return {"skipped": "odd number of ranks"}
    pair = CFGPair.from_world()
    k = rank // 2
    noise, img_lat, emb, extra = synth_inputs(1, F, h, w, lkgd, seed=k)
    kw = dict(domain_features=extra[0], flow_features=extra[1]) if lkgd else {}
    st = pipe.prepare(emb, img_lat, num_frames=F, num_inference_steps=25, cfg_pair=pair, **kw)
    lat0 = (noise * pipe.scheduler.init_noise_sigma).to(device)

    def timed(eager=True):
        lat = lat0
        for i in range(max(2, args.warmup)):
            lat, _ = pipe.denoise_step(st, i, lat, eager=eager)
        dist.barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            lat, _ = pipe.denoise_step(st, (3 + i) % 25, lat, eager=eager)
        e1.record()
        dist.barrier(); torch.cuda.synchronize()
        tt = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return tt
    peer = pair.peer is not None          # prepare() mapped the partner's prediction buffer (symmetric memory)
    t_nccl = None
    if peer:                              # the same steps through the NCCL all-gather, for comparison
        keep, pair.peer = pair.peer, None
        t_nccl = float(timed()[0])
        pair.peer = keep
    t_eager = timed()
    t, graphed = t_eager, False
    if peer and not args.no_graph:        # no collective on the path any more: the split step is ONE graph replay
        pipe.capture(st, lat0)
        t, graphed = timed(eager=False), True
    # correctness on the same inputs: one split step vs the unsplit batch-2 step on this rank
    split, _ = pipe.denoise_step(st, 3, lat0, eager=True)
    st_whole = pipe.prepare(emb, img_lat, num_frames=F, num_inference_steps=25, **kw)
    whole, _ = pipe.denoise_step(st_whole, 3, lat0, eager=True)
    other = split.clone()
    dist.broadcast(other, src=2 * k, group=pair.group)
    mine = {"split_vs_unsplit": _rel(split, whole), "replicated": bool(torch.equal(other, split))}
    allr = [None] * world
    dist.all_gather_object(allr, mine)
    nbytes = F * h * w * 4 * 4
    return {"ms_per_step": float(t[0]), "unsplit_ms_per_step": unsplit_ms, "ratio_vs_unsplit": float(t[0]) / unsplit_ms,
            "pairs": world // 2, "samples_per_s_equiv": (world // 2) / (float(t[0]) * 1e-3),
            "exchange": "peer memory: the fused CFG + Euler kernel reads the partner's half over NVLink, one device-side "
                        "barrier per step, no collective" if peer else "NCCL all-gather (symmetric memory unavailable: "
                        + str(pair.peer_error) + ")",
            "ms_per_step_nccl_allgather": t_nccl,
            "ms_per_step_eager_launches": float(t_eager[0]),
            "exchanged_bytes_per_step_per_rank": nbytes, "cuda_graph": graphed,
            "split_vs_unsplit_rel_l2_max": max(r["split_vs_unsplit"] for r in allr),
            "replicated": all(r["replicated"] for r in allr),
            "what": "ranks (2k,2k+1) share sample k: uncond half on 2k, cond half on 2k+1 (batch 1 each), one exchange "
                    "of the fp32 prediction per step, fused CFG+Euler on both ranks; max-over-ranks device time"}


def section_c5_dp(args, device, rank, world):
    """BASELINE.json configs[4]: the LoRA fine-tuning step, data parallel: every rank its own sample, ONE flat NCCL
    all-reduce of the LoRA (+ quaternion) gradients per optimizer step (train_models/train_svd_lora.py:1300-1302,1683)."""
    import torch.distributed as dist
    from lkgd_b200.training import LoraTrainer
    pipe, cfg, (F, h, w, lrank, lkgd) = build_ours("C5", device)
    tr = LoraTrainer(pipe.unet, lr=1e-4, world_size=world)
    g = torch.Generator().manual_seed(100 + rank)
    B = 1
    b = [torch.randn(B, F, 4, h, w, generator=g), torch.randn(B, F, 4, h, w, generator=g), torch.tensor([1.3] * B),
         torch.randn(B, 4, h, w, generator=g), torch.randn(B, 1, 1024, generator=g),
         torch.tensor([[5.0, 0.02, 127.0]] * B)]
    if lkgd:
        b += [torch.randn(B, 1, 1000, generator=g), torch.randn(B, 1, 1000, generator=g)]
    b = [t.to(device) for t in b]
    # correctness of the collective on real gradients: all-reduce(SUM) == sum of the gathered per-rank gradients
    tr.forward_backward(*b)
    mine = tr.flat_g.clone()
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    expect = torch.stack(parts).double().sum(0)
    tr.optimizer_step()
    reduced_is_sum = _rel(tr.flat_g, expect)
    other = tr.flat_p.clone()
    dist.broadcast(other, src=0)
    replicated = bool(torch.equal(other, tr.flat_p))
    distinct = _rel(parts[0], parts[-1]) > 1e-3          # ranks really worked on different samples
    if not args.no_graph:
        tr.capture(*b)
    for _ in range(max(2, args.warmup)):
        tr.train_step(*b)
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = tr.train_step(*b)
    e1.record()
    dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ok = torch.tensor([float(replicated), float(torch.isfinite(loss).all())], device=device)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    flops = train_flops(cfg, B, F, h, w, lrank)
    return {"metric": "train_steps_per_s", "value": world / (float(t[0]) * 1e-3), "unit": "train_steps/s",
            "ms_per_step": float(t[0]), "n_gpus": world, "scaling": "weak",
            "allreduce_bytes_per_step": int(tr.flat_g.numel() * 4), "trainable_parameters": int(tr.flat_p.numel()),
            "reduced_is_sum_rel_l2": reduced_is_sum, "params_replicated": bool(ok[0] > 0), "loss_finite": bool(ok[1] > 0),
            "ranks_had_distinct_gradients": bool(distinct), "cuda_graph": not args.no_graph,
            "model_tflops_per_gpu": flops / 1e12 / (float(t[0]) * 1e-3),
            "what": f"C5: {F} frames 320x512 ({h}x{w} latents), batch 1 per GPU, LoRA r={lrank} + LKGD quaternion tensors; "
                    "forward+backward graph replay -> one flat NCCL all-reduce(SUM) -> clip + AdamW (1/world in-kernel)"}


if __name__ == "__main__":
    main()
