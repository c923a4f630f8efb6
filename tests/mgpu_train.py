"""torchrun worker (2 ranks, NCCL): data-parallel LoRA training step.  Each rank takes its own sample, the flat
gradient buffer is all-reduced once, and the result must equal the sum of the two per-sample gradients computed on one
GPU; after the optimizer step the parameters must be identical on both ranks.  Launched by tests/test_multigpu_gpu.py."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
from golden_util import REDUCED4, fill_seeded_, rel, seeded_tensor  # noqa: E402


def main():
    rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel

    def make(world):
        unet = UNetSpatioTemporalConditionControlNetModel(**REDUCED4)
        unet.add_lora(8)
        fill_seeded_(unet)
        return LoraTrainer(unet.to(dev), lr=1e-3, world_size=world)

    F_, H_, W_ = 4, 16, 16
    lat, noise = seeded_tensor("dp/lat", (2, F_, 4, H_, W_)), seeded_tensor("dp/noise", (2, F_, 4, H_, W_))
    cond, ctx = seeded_tensor("dp/cond", (2, 4, H_, W_)), seeded_tensor("dp/ctx", (2, 1, 32))
    sig = torch.tensor([0.7, 2.5])
    ids = torch.tensor([[5.0, 0.02, 127.0]] * 2)

    def batch(rows):
        return [t[rows].to(dev) for t in (lat, noise, sig, cond, ctx, ids)]

    tr = make(2)
    tr.forward_backward(*batch(slice(rank, rank + 1)))
    mine = tr.flat_g.clone()
    tr.optimizer_step()                      # all-reduce -> clip -> AdamW -> repack
    reduced = tr.flat_g.clone()
    params = tr.flat_p.clone()
    other = params.clone()
    dist.broadcast(other, src=0)
    res = dict(params_replicated=bool(torch.equal(other, params)))
    # single-process reference: both samples' gradients, summed
    solo = make(1)
    solo.forward_backward(*batch(slice(0, 1)))
    g0 = solo.flat_g.clone()
    solo.forward_backward(*batch(slice(1, 2)))
    g1 = solo.flat_g.clone()
    res["local_matches_solo"] = rel(mine, g0 if rank == 0 else g1)
    res["reduced_is_sum"] = rel(reduced, g0 + g1)
    gathered = [None, None]
    dist.all_gather_object(gathered, res)
    if rank == 0:
        print("RESULT " + json.dumps(gathered), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
