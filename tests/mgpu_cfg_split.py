"""torchrun worker (2 ranks, NCCL): the CFG pair split across two GPUs must reproduce the single-GPU unsplit step.
Launched by tests/test_multigpu_gpu.py; prints one JSON line on rank 0."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from golden_util import REDUCED4, SCHED, fill_seeded_, rel, seeded_tensor  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from lkgd_b200.distributed import CFGPair
    from lkgd_b200.pipeline import StableVideoDiffusionPipeline
    from lkgd_b200.scheduler import EulerDiscreteScheduler
    from lkgd_b200.unet import (ControlNetSDVModel, UNetSpatioTemporalConditionControlNetModel,
                                UNetSpatioTemporalConditionModel)
    pair = CFGPair.from_world()
    res = {}
    for name, cls, cfg in (("plain_0272", UNetSpatioTemporalConditionControlNetModel, REDUCED4),
                           ("plain_b_major", UNetSpatioTemporalConditionControlNetModel,
                            dict(REDUCED4, time_context_order="b_major")),
                           ("lkgd", UNetSpatioTemporalConditionModel, dict(REDUCED4, cross_attention_dim=1024)),
                           ("controlnet_0272", UNetSpatioTemporalConditionControlNetModel, REDUCED4)):
        F_, H_, W_ = 4, 16, 16
        xd = cfg["cross_attention_dim"]
        unet = fill_seeded_(cls(**cfg)).to(dev)
        img = torch.cat([torch.zeros(1, F_, 4, H_, W_), seeded_tensor("d/img", (1, 1, 4, H_, W_)).repeat(1, F_, 1, 1, 1)])
        emb = torch.cat([torch.zeros(1, 1, xd), seeded_tensor("d/emb", (1, 1, xd))])
        kw = {}
        if name == "lkgd":
            kw = dict(domain_features=seeded_tensor("d/dom", (1, 1, 1000)), flow_features=seeded_tensor("d/flow", (1, 1, 1000)))
        lat = seeded_tensor("d/lat", (1, F_, 4, H_, W_))
        cn = None
        if name.startswith("controlnet"):      # BASELINE configs[3]: ControlNet injection fused into each half's UNet forward
            cn = fill_seeded_(ControlNetSDVModel(**{k: v for k, v in cfg.items() if k != "up_block_types"},
                                                 conditioning_channels=2), seed=1).to(dev)
            kw = dict(controlnet_condition=seeded_tensor("d/cc", (F_, 2, 8 * H_, 8 * W_)))
        pipe = StableVideoDiffusionPipeline(unet, EulerDiscreteScheduler(**SCHED), controlnet=cn)
        split = pipe(emb, img, num_frames=F_, num_inference_steps=25, latents=lat, max_steps=3, cfg_pair=pair,
                     return_dict=False, **kw)
        pipe2 = StableVideoDiffusionPipeline(unet, EulerDiscreteScheduler(**SCHED), controlnet=cn)
        whole = pipe2(emb, img, num_frames=F_, num_inference_steps=25, latents=lat, max_steps=3, return_dict=False, **kw)
        other = split.clone()
        dist.broadcast(other, src=0)
        res[name] = dict(split_vs_unsplit=rel(split, whole), replicated=bool(torch.equal(other, split)),
                         exchange="peer" if pair.peer is not None else "nccl")
        if pair.peer is not None:      # the same steps through the NCCL all-gather must give the same latents
            keep, pair.peer = pair.peer, None
            os.environ["LKGD_CFG_PAIR_NCCL"] = "1"
            again = pipe(emb, img, num_frames=F_, num_inference_steps=25, latents=lat, max_steps=3, cfg_pair=pair,
                         return_dict=False, **kw)
            del os.environ["LKGD_CFG_PAIR_NCCL"]
            pair.peer = keep
            res[name]["peer_equals_nccl"] = bool(torch.equal(again, split))
            # the split step as a CUDA-graph replay (two graphs, one per slot of the peer buffer): three steps again
            st = pipe.prepare(emb, img, num_frames=F_, num_inference_steps=25, cfg_pair=pair, **kw)
            x = (lat * pipe.scheduler.init_noise_sigma).to(dev)
            pipe.capture(st, x)
            for i in range(3):
                x, _ = pipe.denoise_step(st, i, x)
            res[name]["graph_equals_eager"] = bool(torch.equal(x, split.to(dev)))
            viacall = pipe(emb, img, num_frames=F_, num_inference_steps=25, latents=lat, max_steps=3, cfg_pair=pair,
                           return_dict=False, use_cuda_graph=True, **kw)
            res[name]["graph_equals_eager"] &= bool(torch.equal(viacall, split))
    if rank == 0:
        print("RESULT " + json.dumps(res), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
