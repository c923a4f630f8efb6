"""Latent-space Stable-Video-Diffusion sampling loop on the lkgd_b200 engine.

Mirrors the denoising part of the reference pipelines' ``__call__``
(pipeline/pipeline_stable_video_diffusion_controlnet.py:364-646: added-time ids :517-526, timesteps :529,
latents :535-545, guidance ramp :553-558, loop :577-630).  The loop itself works on what ``_encode_image`` /
``_encode_vae_image`` return (``image_embeddings`` [2S,1,1024] uncond-first, ``image_latents`` [2S,F,4,h,w]) and yields
latents (``output_type="latent"``); with ``vae=`` and ``image_encoder=`` registered (SURVEY.md section 8f, N1:
``lkgd_b200.vae`` / ``lkgd_b200.clip``) ``__call__(image=...)`` also runs the steps either side of it - CLIP embedding,
noise-augmented VAE encode (:486-513), chunked temporal VAE decode (:268-295, :636) - and returns frames.

Per step the loop launches: one fused pack kernel (CFG duplication + scale_model_input + image-latent concat +
NCHW->channels-last), [ControlNet], the UNet, and one fused CFG-combine + Euler-Karras kernel reading the UNet's
channels-last fp32 prediction directly."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, List, Optional, Union

import os

import numpy as np
import torch

from . import ops
from .engine import Conditioning, Geom
from .preprocess import encode_image
from .scheduler import EulerDiscreteScheduler
from .vae import decode_latents, encode_vae_image
from .unet import ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionModel


@dataclass
class StableVideoDiffusionPipelineOutput:
    frames: torch.Tensor


def _append_dims(x, target_dims):
    dims_to_append = target_dims - x.ndim
    if dims_to_append < 0:
        raise ValueError(f"input has {x.ndim} dims but target_dims is {target_dims}, which is less")
    return x[(...,) + (None,) * dims_to_append]


class StableVideoDiffusionPipeline:
    """``unet`` may be the plain/ControlNet-accepting UNet or the LKGD UNet (then ``domain_features`` and
    ``flow_features`` are required, the signature the reference's missing LKGD pipeline would have - SURVEY F11)."""

    def __init__(self, unet, scheduler: EulerDiscreteScheduler, controlnet: Optional[ControlNetSDVModel] = None,
                 vae=None, image_encoder=None):
        self.unet, self.scheduler, self.controlnet = unet, scheduler, controlnet
        self.vae, self.image_encoder = vae, image_encoder      # the reference's register_modules(vae=, image_encoder=), :122-143
        self._guidance_scale = None

    # ---- the steps either side of the loop
    @torch.no_grad()
    def encode_inputs(self, image: torch.Tensor, height: int, width: int, num_frames: int, noise_aug_strength: float = 0.02,
                      num_videos_per_prompt: int = 1, do_classifier_free_guidance: bool = True, generator=None):
        """Reference steps 3-4 (:486-513): ``image`` [B, 3, H, W] in [0, 1] -> (``image_embeddings`` [2S, 1, D],
        ``image_latents`` [2S, F, 4, h, w]); the VAE sees the [-1, 1] image plus ``noise_aug_strength`` * noise."""
        if self.vae is None or self.image_encoder is None:
            raise ValueError("encode_inputs needs the pipeline's vae and image_encoder")
        if image.ndim == 3:
            image = image.unsqueeze(0)
        if image.ndim != 4 or image.shape[1] != 3:
            raise ValueError(f"image must be a [B, 3, H, W] tensor in [0, 1], got {tuple(image.shape)}")
        if height % 8 != 0 or width % 8 != 0:                       # check_inputs, :297-311
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")
        image = image.to(torch.float32)
        emb = encode_image(self.image_encoder, image, num_videos_per_prompt, do_classifier_free_guidance)
        # VaeImageProcessor.preprocess on a tensor: F.interpolate to (height, width) [default mode], then [0,1] -> [-1,1]
        if tuple(image.shape[-2:]) != (height, width):
            image = torch.nn.functional.interpolate(image, size=(height, width))
        x = 2.0 * image - 1.0
        gdev = generator.device if generator is not None else x.device
        noise = torch.randn(x.shape, generator=generator, device=gdev, dtype=x.dtype).to(x.device)
        x = x + noise_aug_strength * noise
        lat = encode_vae_image(self.vae, x, num_videos_per_prompt, do_classifier_free_guidance).to(emb.dtype)
        return emb, lat.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)

    @torch.no_grad()
    def decode_latents(self, latents: torch.Tensor, num_frames: int, decode_chunk_size: int = 14) -> torch.Tensor:
        """:268-295: [B, F, 4, h, w] -> fp32 [B, 3, F, 8h, 8w]."""
        if self.vae is None:
            raise ValueError("decode_latents needs the pipeline's vae")
        return decode_latents(self.vae, latents, num_frames, decode_chunk_size)

    @staticmethod
    def tensor2vid(frames: torch.Tensor, output_type: str):
        """:67-83 for the tensor output types: [B, 3, F, H, W] in [-1, 1] -> ``"pt"`` [B, F, 3, H, W] in [0, 1], ``"np"`` the
        same as a [B, F, H, W, 3] numpy array."""
        vid = (frames.permute(0, 2, 1, 3, 4) / 2 + 0.5).clamp(0, 1)
        if output_type == "pt":
            return vid
        if output_type == "np":
            return vid.permute(0, 1, 3, 4, 2).float().cpu().numpy()
        raise ValueError(f"output_type {output_type!r}: 'latent', 'pt' and 'np' are supported")

    @property
    def guidance_scale(self):
        return self._guidance_scale

    def _get_add_time_ids(self, fps, motion_bucket_id, noise_aug_strength, dtype, batch_size, num_videos_per_prompt,
                          do_classifier_free_guidance):
        """Inference order [fps, motion_bucket_id, noise_aug_strength] (reference pipeline :239-266)."""
        add_time_ids = [fps, motion_bucket_id, noise_aug_strength]
        passed = self.unet.config.addition_time_embed_dim * len(add_time_ids)
        expected = self.unet.add_embedding.linear_1.in_features
        if expected != passed:
            raise ValueError(f"Model expects an added time embedding vector of length {expected}, but a vector of "
                             f"{passed} was created. The model has an incorrect config.")
        ids = torch.tensor([add_time_ids], dtype=dtype).repeat(batch_size * num_videos_per_prompt, 1)
        return torch.cat([ids, ids]) if do_classifier_free_guidance else ids

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height, width)   # latent-space sizes
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}.")
        if latents is None:
            if isinstance(generator, list):
                latents = torch.cat([torch.randn((1,) + shape[1:], generator=g, device=g.device, dtype=dtype).to(device)
                                     for g in generator])
            else:
                gdev = generator.device if generator is not None else device
                latents = torch.randn(shape, generator=generator, device=gdev, dtype=dtype).to(device)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma.to(latents.device)

    @ops.on_own_device
    @torch.no_grad()
    def prepare(self, image_embeddings: torch.Tensor, image_latents: torch.Tensor, num_frames: Optional[int] = None,
                num_inference_steps: int = 25, min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                fps: int = 7, motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                num_videos_per_prompt: int = 1, controlnet_condition: Optional[torch.Tensor] = None,
                controlnet_cond_scale: float = 1.0, domain_features: Optional[torch.Tensor] = None,
                flow_features: Optional[torch.Tensor] = None, cfg_pair=None) -> dict:
        """Everything ``__call__`` does before the loop (reference :475-558): validates, builds added-time ids,
        timesteps, the frame-wise guidance ramp, and moves the conditioning to the device.  Returns the loop state
        consumed by ``denoise_step``."""
        unet, sched = self.unet, self.scheduler
        device = unet.device
        do_cfg = max_guidance_scale > 1.0          # reference :485 (quirk D1: local, not the property)
        n_lat = image_latents.shape[0]
        S = n_lat // 2 if do_cfg else n_lat
        num_frames = num_frames if num_frames is not None else unet.config.num_frames
        if image_latents.shape[1] != num_frames:
            raise ValueError("image_latents must be repeated over num_frames ([2S, F, 4, h, w])")
        if image_embeddings.shape[0] != n_lat:
            raise ValueError("image_embeddings and image_latents must have the same (CFG-duplicated) batch")
        lkgd = isinstance(unet, UNetSpatioTemporalConditionModel)
        if lkgd and (domain_features is None or flow_features is None):
            raise ValueError("the LKGD UNet needs domain_features and flow_features")
        # the incoming conditioning already carries the reference's num_videos_per_prompt repetition (`_encode_image` /
        # `_encode_vae_image` repeat it, :204,:234): S = batch_size * num_videos_per_prompt samples
        if num_videos_per_prompt < 1 or S % num_videos_per_prompt:
            raise ValueError(f"the conditioning batch ({S} samples) is not a multiple of num_videos_per_prompt="
                             f"{num_videos_per_prompt}")
        fps = fps - 1                              # reference :482: the model was conditioned on fps - 1
        added_time_ids = self._get_add_time_ids(fps, motion_bucket_id, noise_aug_strength, torch.float32,
                                                S // num_videos_per_prompt, num_videos_per_prompt, do_cfg).to(device)
        sched.set_timesteps(num_inference_steps, device=device)
        guidance = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(device)
        self._guidance_scale = _append_dims(guidance.repeat(S, 1), 5)
        if controlnet_condition is not None and self.controlnet is None:
            raise ValueError("controlnet_condition given but the pipeline has no controlnet")
        if controlnet_condition is not None:
            cc = controlnet_condition
            if cc.ndim == 4:
                cc = cc.unsqueeze(0)
            # reference :547-550 duplicates the condition for the two CFG halves (D3): kept as ONE copy + a repeat count, the
            # ControlNet's pixel-resolution condition encoder then runs once per step instead of twice
            cond_repeat = 2 if do_cfg else 1
            if cc.shape[0] * cond_repeat != n_lat:
                raise ValueError("controlnet_condition batch does not match the conditioning batch")
            controlnet_condition = cc.to(device=device, dtype=torch.float32)
        if cfg_pair is not None and not do_cfg:
            raise ValueError("cfg_pair needs classifier-free guidance")
        if cfg_pair is not None and hasattr(cfg_pair, "enable_peer") and not os.environ.get("LKGD_CFG_PAIR_NCCL"):
            # exchange through peer memory where the box offers it (collective over the pair: every rank prepares)
            n_rows = S * num_frames * image_latents.shape[-2] * image_latents.shape[-1]
            ld = int(getattr(self.unet.config, "out_channels", 4))
            if getattr(cfg_pair, "peer", None) is None or cfg_pair.peer["n"] != n_rows or cfg_pair.peer["ld"] != ld:
                cfg_pair.enable_peer(n_rows, ld, device)
        return dict(cfg_pair=cfg_pair, S=S, n_batch=n_lat, F=num_frames, h=image_latents.shape[-2],
                    w=image_latents.shape[-1], do_cfg=do_cfg, added_time_ids=added_time_ids,
                    guidance=guidance.reshape(-1).to(torch.float32).contiguous(),
                    image_latents=image_latents.to(device=device, dtype=torch.float32).contiguous(),
                    image_embeddings=image_embeddings.to(device), controlnet_condition=controlnet_condition,
                    controlnet_cond_scale=controlnet_cond_scale,
                    controlnet_cond_repeat=(2 if do_cfg else 1) if controlnet_condition is not None else 1,
                    extra=(domain_features.to(device), flow_features.to(device)) if lkgd else ())

    @property
    def device(self):
        return self.unet.device

    def _step_body(self, st: dict, latents: torch.Tensor, scale, t, sigmas_dev=None, in_place=False, want_v=False):
        """pack -> [ControlNet] -> UNet -> CFG + Euler for the unsplit batch.  ``scale`` / ``t`` are host floats on the
        eager path and fp32 device scalars inside a captured graph."""
        sched, unet = self.scheduler, self.unet
        pk = unet.packed()
        dev_scalars = torch.is_tensor(scale)
        # CFG duplication + scale_model_input + concat(image_latents) + layout, one kernel (:579-584)
        x = ops.pack_input(latents, 0.0 if dev_scalars else scale, st["image_latents"], N=st["n_batch"], Cpad=pk.cin_pad,
                           scale_dev=scale if dev_scalars else None)
        g = Geom(st["n_batch"], st["F"], st["h"], st["w"])
        kw = {}
        if st["controlnet_condition"] is not None:
            if st.get("fuse_controlnet", True):
                # residual injection fused into the UNet forward: the ControlNet's zero convs add onto the UNet's skips
                kw = dict(fused_controlnet=(self.controlnet, st["controlnet_condition"], st["controlnet_cond_scale"],
                                            st["controlnet_cond_repeat"]))
            else:   # the reference's hand-off: 12 + 1 residual tensors, added by the UNet (:585-607)
                down, mid = self.controlnet.forward_packed(x, g, t, st["image_embeddings"], st["added_time_ids"],
                                                           st["controlnet_condition"], st["controlnet_cond_scale"],
                                                           cond_repeat=st["controlnet_cond_repeat"])
                kw = dict(down_block_additional_residuals=down, mid_block_additional_residual=mid)
        rows = unet.forward_packed(x, g, t, st["image_embeddings"], *st["extra"],
                                   added_time_ids=st["added_time_ids"], **kw)
        if st.get("direct_fusion"):
            # trans pipelines (pipeline_stable_video_diffusion_trans_controlnet.py:637-667): the batch is [forward samples |
            # their time-reversed partners]; the guided prediction takes the bidirectional x0 blend instead of the plain step
            if sigmas_dev is not None:
                raise ValueError("direct_fusion is not captured in the CUDA graph")
            i = sched._step_index
            _, v = ops.cfg_euler_step(rows, st["guidance"] if st["do_cfg"] else None, latents,
                                      float(sched._sigmas_host[i]), float(sched._sigmas_host[i + 1]), cfg=st["do_cfg"],
                                      want_v=True)
            return sched.step_direct_fusion(v, sched.timesteps[i], latents), (v if want_v else None)
        return sched.step_cfg_rows(rows, st["guidance"] if st["do_cfg"] else None, latents, cfg=st["do_cfg"],
                                   want_v=want_v, sigmas_dev=sigmas_dev, in_place=in_place)

    def _pair_body(self, st: dict, latents: torch.Tensor, scale, t, sigmas_dev=None, in_place=False, want_v=False):
        """The CFG-pair split of one step (lkgd_b200/distributed.py): this rank runs ONE half of [uncond | cond] with
        batch S; the halves' predictions meet in the fused CFG + Euler kernel, which both ranks run (replicated latents).
        With peer memory the kernel reads the partner's half over NVLink (one device-side barrier, capturable in a CUDA
        graph); else the halves are all-gathered first."""
        sched, unet = self.scheduler, self.unet
        pk = unet.packed()
        pair = st["cfg_pair"]
        lo, hi = pair.batch_slice(st["S"])
        # this rank's half of the conditioning latents: a fixed buffer refreshed every step (also inside a captured graph),
        # so that `st["image_latents"]` stays the one input a caller updates with copy_
        if st.get("image_latents_half") is None:
            st["image_latents_half"] = torch.empty_like(st["image_latents"][lo:hi]).contiguous()
        st["image_latents_half"].copy_(st["image_latents"][lo:hi])
        dev_scalars = torch.is_tensor(scale)
        x = ops.pack_input(latents, 0.0 if dev_scalars else scale, st["image_latents_half"], N=st["S"], Cpad=pk.cin_pad,
                           scale_dev=scale if dev_scalars else None)
        g = Geom(st["S"], st["F"], st["h"], st["w"])
        kw = {}
        if st["controlnet_condition"] is not None:
            # both halves are conditioned on the SAME frames (reference :547-550 duplicates them, D3): this half takes
            # the one copy as it is; injection fused into the UNet forward as in the unsplit step
            if not st.get("fuse_controlnet", True):
                raise ValueError("the CFG pair split runs the ControlNet fused into the UNet forward")
            kw = dict(fused_controlnet=(self.controlnet, st["controlnet_condition"], st["controlnet_cond_scale"], 1))
        rows = unet.forward_packed(x, g, t, st["image_embeddings"], *st["extra"],
                                   added_time_ids=st["added_time_ids"], batch_slice=(lo, hi), **kw)
        if pair.peer is not None:          # exchange inside the combine kernel: the partner's half is read over NVLink
            u, c = pair.publish(rows)
            return sched.step_cfg_rows(u, st["guidance"], latents, cfg=True, want_v=want_v, pred_cond=c,
                                       sigmas_dev=sigmas_dev, in_place=in_place)
        if sigmas_dev is not None:
            raise ValueError("the CFG pair split over an NCCL all-gather is not captured in a CUDA graph")
        rows = pair.exchange(rows)
        return sched.step_cfg_rows(rows, st["guidance"], latents, cfg=True, want_v=want_v)

    @ops.on_own_device
    @torch.no_grad()
    def capture(self, st: dict, latents: torch.Tensor) -> dict:
        """Captures ONE denoise step for the loop state ``st`` in a CUDA graph: every shape is static
        across the 25 steps, only four scalars change (1/sqrt(sigma^2+1), sigma, sigma_next, t = 0.25 ln sigma) - they
        live in a 16-byte device buffer refreshed from a device-resident table before each replay, the latents are
        updated in place in a static buffer.  ~800 kernel launches, their ctypes calls, TMA-descriptor encodes and
        allocations become one ``cudaGraphLaunch``.  Call after at least one eager ``denoise_step`` (lazy weight packing
        and one-time kernel attributes must not happen under capture).  The conditioning tensors in ``st`` are the
        graph's static inputs: refresh them with ``copy_`` (not by rebinding the dict entries)."""
        pair = st.get("cfg_pair")
        if pair is not None and pair.peer is None:
            raise ValueError("the CFG pair split exchanges predictions with NCCL every step: not captured "
                             "(peer memory, CFGPair.enable_peer, makes the split step capturable)")
        sched = self.scheduler
        dev = self.unet.device
        sig = np.asarray(sched._sigmas_host, dtype=np.float32)
        ts = np.asarray(sched._timesteps_host, dtype=np.float32)
        tab = np.stack([np.asarray([float(1.0 / np.sqrt(np.float32(s) ** 2 + 1)) for s in sig[:-1]], dtype=np.float32),
                        sig[:-1], sig[1:], ts], axis=1)
        G = dict(table=torch.from_numpy(np.ascontiguousarray(tab)).to(dev), lat=latents.to(dev).clone().contiguous(),
                 params=torch.zeros(4, device=dev, dtype=torch.float32), n=len(ts))
        G["params"].copy_(G["table"][0])
        keep = G["lat"].clone()

        def body():
            sched.index_for(0)
            step = self._step_body if pair is None else self._pair_body
            step(st, G["lat"], G["params"][0:1], G["params"][3:4], sigmas_dev=G["params"][1:3], in_place=True)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()                                   # warm the allocator / statistics arena off the default stream
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        if pair is not None:
            # the peer buffer has two slots used by alternate steps: one graph per slot, replayed alternately (both ranks
            # capture and replay in lockstep - the barrier inside is the only coupling)
            if pair.peer["tick"] & 1:
                body()                               # start the pair of captures on slot 0 (the warm-up above took one)
            with torch.cuda.graph(graph):
                body()
            odd = torch.cuda.CUDAGraph()
            with torch.cuda.graph(odd, pool=graph.pool()):
                body()
            G["graphs"] = (graph, odd)
        else:
            with torch.cuda.graph(graph):
                body()
        G["lat"].copy_(keep)
        G["graph"] = graph
        st["graph"] = G
        return st

    @ops.on_own_device
    @torch.no_grad()
    def denoise_step(self, st: dict, i: int, latents: torch.Tensor, want_v: bool = False, eager: bool = False):
        """One iteration of the reference's denoising loop (:577-619) for step index ``i``: fused
        dup/scale/concat/layout kernel -> [ControlNet] -> UNet -> fused CFG + Euler-Karras kernel.
        ``latents``: fp32 [S,F,4,h,w] on the device.  Returns (next latents, guided prediction or None).
        After ``capture(st, ...)`` the step is a CUDA-graph replay and the returned latents ARE the graph's static
        buffer (updated in place by the next call); ``eager=True`` / ``want_v`` keep the kernel-by-kernel path."""
        sched, unet = self.scheduler, self.unet
        G = st.get("graph")
        if G is not None and not eager and not want_v:
            if not 0 <= i < G["n"]:
                raise IndexError("step index outside the captured schedule")
            if latents.data_ptr() != G["lat"].data_ptr():
                G["lat"].copy_(latents, non_blocking=True)
            G["params"].copy_(G["table"][i], non_blocking=True)
            if "graphs" in G:                        # CFG pair over peer memory: the slot alternates call by call
                P = st["cfg_pair"].peer
                G["graphs"][P["tick"] & 1].replay()
                P["tick"] += 1
            else:
                G["graph"].replay()
            sched.index_for(i + 1)
            return G["lat"], None
        pk = unet.packed()
        sched.index_for(i)
        sigma = float(sched._sigmas_host[i])
        t = float(sched._timesteps_host[i])
        scale = float(1.0 / np.sqrt(np.float32(sigma) ** 2 + 1))
        if st.get("cfg_pair") is not None:
            return self._pair_body(st, latents, scale, t, want_v=want_v)
        return self._step_body(st, latents, scale, t, want_v=want_v)

    @torch.no_grad()
    def __call__(self, image_embeddings: Optional[torch.Tensor] = None, image_latents: Optional[torch.Tensor] = None,
                 num_frames: Optional[int] = None,
                 num_inference_steps: int = 25, min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                 fps: int = 7, motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                 num_videos_per_prompt: int = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 controlnet_condition: Optional[torch.Tensor] = None, controlnet_cond_scale: float = 1.0,
                 domain_features: Optional[torch.Tensor] = None, flow_features: Optional[torch.Tensor] = None,
                 cfg_pair=None, output_type: str = "latent", callback_on_step_end: Optional[Callable] = None,
                 return_dict: bool = True, max_steps: Optional[int] = None, return_trajectory: bool = False,
                 use_cuda_graph: bool = False, fuse_controlnet: bool = True, direct_fusion: bool = False,
                 image: Optional[torch.Tensor] = None, height: int = 576, width: int = 1024,
                 decode_chunk_size: Optional[int] = None):
        if output_type not in ("latent", "pt", "np"):
            raise ValueError(f"output_type {output_type!r}: 'latent', 'pt' and 'np' are supported")
        if output_type != "latent" and (self.vae is None or return_trajectory):
            raise ValueError("decoded frames need the pipeline's vae (and no return_trajectory): pass vae= to the "
                             "pipeline or use output_type='latent'")
        if num_frames is None:
            num_frames = self.unet.config.num_frames
        if image is not None:
            if image_embeddings is not None or image_latents is not None:
                raise ValueError("pass either image= or (image_embeddings, image_latents)")
            image_embeddings, image_latents = self.encode_inputs(
                image, height, width, num_frames, noise_aug_strength, num_videos_per_prompt, max_guidance_scale > 1.0,
                generator)
        elif image_embeddings is None or image_latents is None:
            raise ValueError("pass image= (with vae and image_encoder registered) or image_embeddings and image_latents")
        st = self.prepare(image_embeddings, image_latents, num_frames, num_inference_steps, min_guidance_scale,
                          max_guidance_scale, fps, motion_bucket_id, noise_aug_strength, num_videos_per_prompt,
                          controlnet_condition, controlnet_cond_scale, domain_features, flow_features, cfg_pair)
        st["fuse_controlnet"] = fuse_controlnet
        st["direct_fusion"] = direct_fusion
        if direct_fusion and (use_cuda_graph or cfg_pair is not None or st["S"] % 2):
            raise ValueError("direct_fusion needs an even number of samples (forward | reversed) and the eager loop")
        device = self.unet.device
        latents = self.prepare_latents(st["S"], st["F"], self.unet.config.in_channels, st["h"], st["w"],
                                       torch.float32, device, generator, latents).to(torch.float32).contiguous()
        timesteps = self.scheduler.timesteps
        traj: List[torch.Tensor] = []
        preds: List[torch.Tensor] = []
        for i in range(len(timesteps)):
            if max_steps is not None and i >= max_steps:
                break
            latents, v = self.denoise_step(st, i, latents, want_v=return_trajectory)
            if i == 0 and use_cuda_graph and not return_trajectory and \
                    (cfg_pair is None or getattr(cfg_pair, "peer", None) is not None):
                self.capture(st, latents)              # steps 1.. replay the captured graph (same kernels, same order)
            if return_trajectory:
                preds.append(v)
                traj.append(latents)
            if callback_on_step_end is not None:
                out = callback_on_step_end(self, i, timesteps[i], {"latents": latents})
                latents = out.pop("latents", latents)
        if return_trajectory:
            return latents, preds, traj
        frames = latents
        if output_type != "latent":                                # :632-636
            chunk = decode_chunk_size if decode_chunk_size is not None else num_frames
            frames = self.tensor2vid(self.decode_latents(latents, num_frames, chunk), output_type)
        if not return_dict:
            return frames
        return StableVideoDiffusionPipelineOutput(frames=frames)


class StableVideoDiffusionSmoothPipeline(StableVideoDiffusionPipeline):
    """Latent-space part of the reference's ``smooth`` pipeline (pipeline/pipeline_stable_video_diffusion_smooth.py,
    SURVEY 8f N3): a long clip of ``T`` frames is re-noised to step ``start_step`` (:520) and denoised in windows of at
    most ``num_frames`` frames whose boundaries are re-drawn at random every step (:526-534); every window runs the UNet
    on [window, time-flipped window] x CFG with the window's first / last frame as the conditioning image (:546-575),
    keeps the forward half's guided prediction (:589-590), and ONE Euler step updates all frames (:593).

    Per window: the fused pack kernel (CFG duplication + scale + concat + layout), the UNet, the fused CFG kernel."""

    @staticmethod
    def get_chunks(flen: int, num_frames: int, rng=None):
        rng = np.random if rng is None else rng          # the reference draws from numpy's global stream
        x_index = torch.arange(flen)
        rand_first = rng.randint(0, num_frames) + 1
        chunks = x_index[rand_first:].split(num_frames, dim=0)
        chunks = [x_index[:rand_first]] + list(chunks) if len(chunks[0]) > 0 else [x_index[:rand_first]]
        return [[int(i) for i in chunk] for chunk in chunks]

    @ops.on_own_device
    @torch.no_grad()
    def __call__(self, image_embeddings: torch.Tensor, image_latents: torch.Tensor, original_image_latents: torch.Tensor,
                 num_frames: Optional[int] = None, start_step: int = 0, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0, fps: int = 7,
                 motion_bucket_id: int = 127, noise_aug_strength: float = 0.02, generator=None,
                 noise: Optional[torch.Tensor] = None, chunk_rng=None, output_type: str = "latent",
                 return_dict: bool = True):
        """``image_embeddings`` [2T,1,D] / ``image_latents`` [2T,4,h,w]: per FRAME, unconditional (zeros) half first
        (:441, :462-469); ``original_image_latents`` [1,T,4,h,w]: the clip's scaled VAE latents (:464)."""
        if output_type != "latent":
            raise ValueError("lkgd_b200 covers the denoise loop only: use output_type='latent'")
        unet, sched = self.unet, self.scheduler
        device = unet.device
        do_cfg = max_guidance_scale > 1.0
        if not do_cfg:
            raise ValueError("the smooth pipeline is defined for classifier-free guidance (max_guidance_scale > 1)")
        T = original_image_latents.shape[1]
        if image_latents.shape[0] != 2 * T or image_embeddings.shape[0] != 2 * T:
            raise ValueError("image_latents / image_embeddings must hold 2*T per-frame entries (uncond half first)")
        num_frames = num_frames if num_frames is not None else unet.config.num_frames
        ids = self._get_add_time_ids(fps - 1, motion_bucket_id, noise_aug_strength, torch.float32, 1, 1, True)
        ids4 = torch.cat([ids] * 2, dim=0).to(device)                      # :541
        sched.set_timesteps(num_inference_steps, device=device)
        if not 0 <= start_step < num_inference_steps:
            raise ValueError("start_step outside the schedule")
        x0 = original_image_latents.to(device=device, dtype=torch.float32).contiguous()
        if noise is None:
            gdev = generator.device if generator is not None else device
            noise = torch.randn(x0.shape, generator=generator, device=gdev, dtype=torch.float32)
        latents = sched.add_noise(x0, noise.to(device), sched.timesteps[[start_step]]).contiguous()      # :520
        img_lat = image_latents.to(device=device, dtype=torch.float32)
        img_emb = image_embeddings.to(device)
        h, w = x0.shape[-2:]
        pk = unet.packed()
        for i in range(start_step, num_inference_steps):
            sched.index_for(i)
            sigma, sigma_next = float(sched._sigmas_host[i]), float(sched._sigmas_host[i + 1])
            t = float(sched._timesteps_host[i])
            scale = float(1.0 / np.sqrt(np.float32(sigma) ** 2 + 1))
            noise_pred = torch.empty_like(latents)
            for chunk in self.get_chunks(T, num_frames, chunk_rng):
                n = len(chunk)
                lc = latents[:, chunk]
                pair = torch.cat([lc, lc.flip(dims=[1])], dim=0).contiguous()                      # :549-551
                first = [chunk[0], chunk[-1], chunk[0] + T, chunk[-1] + T]                         # :553-556
                cur_lat = img_lat[first].unsqueeze(1).repeat(1, n, 1, 1, 1).contiguous()
                x = ops.pack_input(pair, scale, cur_lat, N=4, Cpad=pk.cin_pad)                     # :565-569, one kernel
                rows = unet.forward_packed(x, Geom(4, n, h, w), t, img_emb[first].contiguous(), added_time_ids=ids4)
                g = torch.linspace(min_guidance_scale, max_guidance_scale, n, device=device, dtype=torch.float32)
                _, v = ops.cfg_euler_step(rows, g.contiguous(), pair, sigma, sigma_next, cfg=True, want_v=True)   # :580-587
                noise_pred[:, chunk] = v[:1]                                                       # :589-590
            latents, _ = ops.cfg_euler_step(noise_pred.contiguous(), None, latents, sigma, sigma_next, cfg=False)   # :593
            sched._step_index = i + 1
        if not return_dict:
            return latents
        return StableVideoDiffusionPipelineOutput(frames=latents)
