"""Multi-GPU plumbing of the sampling path.  One process per GPU, ``torch.distributed``.

The denoise step shards only where it splits naturally:
  * independent samples -> ``shard_range``: every rank denoises its own samples, NO data-path collective;
  * the classifier-free-guidance pair -> ``CFGPair``: rank 2k runs the unconditional half, rank 2k+1 the conditional
    half (batch 1 each), and ONE exchange per step feeds the fused CFG + Euler kernel, which both ranks run so the latents
    stay replicated.  The exchange is peer memory where the box offers it (``CFGPair.enable_peer``: each rank's [S*F*H*W, 4]
    fp32 prediction, 3.7 MB at 25 frames 72x128, sits in a symmetric buffer and ``lkgd_cfg_euler_step_pair`` reads the
    partner's half over NVLink inside the combine kernel: one device-side barrier, no collective, and the whole split step
    can be captured in a CUDA graph), else an NCCL / gloo all-gather (``CFGPair.exchange``).
  * LoRA training -> ``allreduce_flat_``: replicas; the ONE flat fp32 gradient buffer of all trainable tensors is summed
    with a single all-reduce per optimizer step (the averaging 1/world is folded into the optimizer kernel) - what DDP's
    bucketed all-reduce does for the reference (train_models/train_svd_lora.py:1300-1302,1683).
Frames are never a shard axis: temporal conv / attention and the 5-D GroupNorm couple all of them.
The reference has none of this (single process, ``pipeline...controlnet.py:577-619``); backend-agnostic on purpose so
the host logic is covered by world-size-2 ``gloo`` tests on CPU."""
from __future__ import annotations

from typing import List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced slice [lo, hi) of ``n`` independent samples for ``rank`` (first ``n % world`` ranks get one
    more)."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


class CFGPair:
    """Two ranks that share one sample's classifier-free-guidance batch.  ``role`` 0 = unconditional half (first in the
    reference's ``torch.cat([uncond, cond])`` order, pipeline :206-232), 1 = conditional half."""

    def __init__(self, group, role: int):
        if role not in (0, 1):
            raise ValueError("role must be 0 (uncond) or 1 (cond)")
        self.group, self.role = group, role
        self.peer, self.peer_error = None, None

    @staticmethod
    def from_world() -> "CFGPair":
        """Ranks (2k, 2k+1) of the default group form pair k.  Collective: every rank must call it."""
        world, rank = dist.get_world_size(), dist.get_rank()
        if world % 2:
            raise ValueError("the CFG pair split needs an even number of ranks")
        mine = None
        for k in range(world // 2):
            g = dist.new_group([2 * k, 2 * k + 1])
            if rank // 2 == k:
                mine = g
        return CFGPair(mine, rank % 2)

    def batch_slice(self, S: int) -> Tuple[int, int]:
        """Rows of the CFG-duplicated batch [uncond(S) | cond(S)] this rank computes."""
        return self.role * S, (self.role + 1) * S

    # ---- exchange over peer memory (NVLink / NVSwitch): no collective, the combine kernel reads the partner's half
    def enable_peer(self, n_rows: int, ld: int, device) -> bool:
        """Allocates this rank's prediction buffer in symmetric (peer-mapped) memory and maps the partner's
        (``torch.distributed._symmetric_memory``: CUDA VMM allocations exchanged inside the pair's group).  Two slots,
        used alternately by consecutive steps, so ONE device-side barrier per step orders both the read-after-write of
        this step and the write-after-read of the step after next.  Collective over the pair; returns False (and leaves
        the NCCL ``exchange`` as the path) where the rendezvous is not available: CPU / gloo groups, no P2P access."""
        self.peer = None
        try:
            import torch.distributed._symmetric_memory as symm
            buf = symm.empty((2, n_rows, ld), dtype=torch.float32, device=device)
            hdl = symm.rendezvous(buf, self.group)
            other = 1 - hdl.rank
            theirs = hdl.get_buffer(other, (2, n_rows, ld), torch.float32)
            ok = torch.ones(1, device=device)
        except Exception as e:                     # noqa: BLE001 - any failure means "use the collective"
            self.peer_error = repr(e)
            ok = torch.zeros(1, device=device) if torch.device(device).type == "cuda" else torch.zeros(1)
            buf = hdl = theirs = None
        try:                                       # both ranks must agree, or one would wait in a barrier alone
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.group)
        except Exception:                          # noqa: BLE001
            return False
        if float(ok) < 1.0:
            return False
        self.peer = dict(mine=buf, theirs=theirs, hdl=hdl, n=n_rows, ld=ld, tick=0)
        return True

    def publish(self, pred_rows: torch.Tensor):
        """This half's prediction -> the next slot of the symmetric buffer (slots alternate call by call, on both ranks
        alike), one barrier inside the pair, then the (unconditional rows, conditional rows) the fused CFG + Euler kernel
        reads: one of the two is the partner's memory."""
        P = self.peer
        k = P["tick"] & 1
        P["tick"] += 1
        mine, theirs = P["mine"][k], P["theirs"][k]
        if pred_rows.shape != mine.shape:
            raise ValueError(f"prediction rows {tuple(pred_rows.shape)} do not match the peer buffer {tuple(mine.shape)}")
        mine.copy_(pred_rows)
        P["hdl"].barrier(channel=0, timeout_ms=20000)     # a lost partner traps the kernel instead of hanging the GPU
        return (mine, theirs) if self.role == 0 else (theirs, mine)

    def exchange(self, pred_rows: torch.Tensor) -> torch.Tensor:
        """[n, c] prediction of this half -> [2n, c] (uncond rows first) on both ranks: the one collective per step."""
        pred_rows = pred_rows.contiguous()
        parts = [torch.empty_like(pred_rows) for _ in range(2)]
        dist.all_gather(parts, pred_rows, group=self.group)
        return torch.cat(parts, 0)


def gather_samples(latents: torch.Tensor, dst: int = 0) -> Optional[List[torch.Tensor]]:
    """Collects every rank's finished latents on ``dst`` (outside the timed / data path)."""
    world = dist.get_world_size()
    out = [torch.empty_like(latents) for _ in range(world)] if dist.get_rank() == dst else None
    dist.gather(latents.contiguous(), out, dst=dst)
    return out


def allreduce_flat_(flat: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM of a flat gradient buffer across ``group`` (default group when None).  One collective per optimizer
    step; no-op without an initialised process group or with a single rank."""
    if not dist.is_available() or not dist.is_initialized():
        return flat
    if dist.get_world_size(group) == 1:
        return flat
    if flat.dim() != 1 or not flat.is_contiguous():
        raise ValueError("allreduce_flat_ expects one contiguous 1-D buffer (all trainable gradients back to back)")
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    return flat
