"""World-size-2 ``gloo`` tests (CPU) of the N>1 host logic: sample sharding and the CFG-pair split
(lkgd_b200/distributed.py).  The denoiser in these tests is the CPU oracle - the point is the plumbing: which rank
computes which half, the order of the exchanged predictions, and that indexing the temporal cross-attention
contexts over the WHOLE batch makes a split step equal to the unsplit reference step."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from golden_util import REDUCED4, SCHED, fill_seeded_, rel, seeded_tensor


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _run(fn, world, *args):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    procs = [ctx.Process(target=_entry, args=(fn, r, world, port, q, args)) for r in range(world)]
    for p in procs:
        p.start()
    out = {}
    import time
    t0 = time.time()
    while len(out) < world and time.time() - t0 < 300 and any(p.is_alive() for p in procs) or not q.empty():
        if q.empty():
            time.sleep(0.05)
            continue
        r, v = q.get()
        out[r] = _unplain(v)
    for p in procs:
        p.join(60)
    assert all(p.exitcode == 0 for p in procs), [p.exitcode for p in procs]
    assert len(out) == world
    return out


def _plain(v):
    """tensors -> numpy so results cross the process boundary by value (no shared-memory handles)."""
    if torch.is_tensor(v):
        return ("t", v.detach().cpu().numpy())
    if isinstance(v, (list, tuple)):
        return ("l", [_plain(x) for x in v])
    return ("v", v)


def _unplain(v):
    kind, x = v
    if kind == "t":
        return torch.from_numpy(x)
    if kind == "l":
        return [_unplain(y) for y in x]
    return x


def _entry(fn, rank, world, port, q, args):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        q.put((rank, _plain(fn(rank, world, *args))))
    finally:
        dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    from lkgd_b200.distributed import shard_range
    for n in (0, 1, 7, 8, 25):
        for world in (1, 2, 3, 8):
            parts = [shard_range(n, world, r) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(parts, parts[1:]))
            sizes = [hi - lo for lo, hi in parts]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _exchange_worker(rank, world):
    from lkgd_b200.distributed import CFGPair, gather_samples
    pair = CFGPair.from_world()
    assert pair.role == rank % 2 and pair.batch_slice(3) == ((0, 3) if rank % 2 == 0 else (3, 6))
    mine = torch.full((5, 4), float(rank))
    # no peer memory on a gloo / CPU group: both ranks agree on the collective (a rank alone in a barrier would hang)
    assert pair.enable_peer(5, 4, "cpu") is False and pair.peer is None and pair.peer_error is not None
    both = pair.exchange(mine)
    got = gather_samples(torch.full((2,), float(rank)))
    return both, got


def test_cfg_pair_exchange_order_and_gather():
    out = _run(_exchange_worker, 2)
    want = torch.cat([torch.zeros(5, 4), torch.ones(5, 4)])           # uncond (role 0) rows first on BOTH ranks
    assert torch.equal(out[0][0], want) and torch.equal(out[1][0], want)
    assert [float(t[0]) for t in out[0][1]] == [0.0, 1.0] and out[1][1] is None


def _cfg_split_worker(rank, world, order):
    """One CFG Euler-Karras step with the pair split across two ranks, the oracle as the denoiser."""
    import oracle as O
    from lkgd_b200.distributed import CFGPair
    pair = CFGPair.from_world()
    F_, H_, W_ = 4, 8, 8
    unet = fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**dict(REDUCED4, time_context_order=order))).eval()
    sched = O.EulerDiscreteScheduler(**SCHED)
    sched.set_timesteps(25)
    lat = seeded_tensor("dist/lat", (1, F_, 4, H_, W_)) * sched.init_noise_sigma
    img = torch.cat([torch.zeros(1, F_, 4, H_, W_), seeded_tensor("dist/img", (1, 1, 4, H_, W_)).repeat(1, F_, 1, 1, 1)])
    emb = torch.cat([torch.zeros(1, 1, 32), seeded_tensor("dist/emb", (1, 1, 32))])
    ids = O.add_time_ids_inference(6, 127, 0.02, 1)
    t = sched.timesteps[0]
    g = O.guidance_ramp(1.0, 3.0, F_)
    x = sched.scale_model_input(lat, t)
    lo, hi = pair.batch_slice(1)
    if order == "hw_major_0272":      # every transformer indexes the contexts of the whole CFG batch
        for m in unet.modules():
            if isinstance(m, O.TransformerSpatioTemporalModel):
                m.cfg_split = (emb, lo)
    with torch.no_grad():
        half = unet(torch.cat([x, img[lo:hi]], dim=2), t, emb[lo:hi], added_time_ids=ids[lo:hi]).sample
    both = pair.exchange(half.reshape(-1, 4)).reshape(2, *half.shape[1:])
    pred = O.cfg_combine(both, g)
    sched._step_index = 0
    return sched.step(pred, t, lat).prev_sample


def _cfg_split_controlnet_worker(rank, world, give_cn_both_contexts):
    """The pair split with a ControlNet (BASELINE configs[3]), the oracle as the denoiser: every half runs the ControlNet
    and the UNet on its own batch entry; under the 0.27.2 context order BOTH models' temporal cross-attentions must index
    the embeddings of the whole guidance batch."""
    import oracle as O
    from lkgd_b200.distributed import CFGPair
    pair = CFGPair.from_world()
    F_, H_, W_ = 4, 8, 8
    unet = fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)).eval()
    cn = fill_seeded_(O.ControlNetSDVModel(**{k: v for k, v in REDUCED4.items() if k != "up_block_types"},
                                           conditioning_channels=2), seed=1).eval()
    sched = O.EulerDiscreteScheduler(**SCHED)
    sched.set_timesteps(25)
    lat = seeded_tensor("dist/lat", (1, F_, 4, H_, W_)) * sched.init_noise_sigma
    img = torch.cat([torch.zeros(1, F_, 4, H_, W_), seeded_tensor("dist/img", (1, 1, 4, H_, W_)).repeat(1, F_, 1, 1, 1)])
    emb = torch.cat([torch.zeros(1, 1, 32), seeded_tensor("dist/emb", (1, 1, 32))])
    cc = seeded_tensor("dist/cc", (1, F_, 2, 8 * H_, 8 * W_))
    ids = O.add_time_ids_inference(6, 127, 0.02, 1)
    t = sched.timesteps[0]
    x = sched.scale_model_input(lat, t)
    lo, hi = pair.batch_slice(1)
    for model in ((unet, cn) if give_cn_both_contexts else (unet,)):
        for m in model.modules():
            if isinstance(m, O.TransformerSpatioTemporalModel):
                m.cfg_split = (emb, lo)
    with torch.no_grad():
        xin = torch.cat([x, img[lo:hi]], dim=2)
        down, mid = cn(xin, t, encoder_hidden_states=emb[lo:hi], controlnet_cond=cc, added_time_ids=ids[lo:hi],
                       conditioning_scale=1.0, guess_mode=False, return_dict=False)
        half = unet(xin, t, emb[lo:hi], added_time_ids=ids[lo:hi], down_block_additional_residuals=down,
                    mid_block_additional_residual=mid).sample
    both = pair.exchange(half.reshape(-1, 4)).reshape(2, *half.shape[1:])
    sched._step_index = 0
    return sched.step(O.cfg_combine(both, O.guidance_ramp(1.0, 3.0, F_)), t, lat).prev_sample


def test_cfg_pair_split_with_controlnet_equals_unsplit_step():
    import oracle as O
    out = _run(_cfg_split_controlnet_worker, 2, True)
    assert torch.equal(out[0], out[1])
    F_, H_, W_ = 4, 8, 8
    unet = fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**REDUCED4)).eval()
    cn = fill_seeded_(O.ControlNetSDVModel(**{k: v for k, v in REDUCED4.items() if k != "up_block_types"},
                                           conditioning_channels=2), seed=1).eval()
    sched = O.EulerDiscreteScheduler(**SCHED)
    sched.set_timesteps(25)
    lat = seeded_tensor("dist/lat", (1, F_, 4, H_, W_)) * sched.init_noise_sigma
    img = torch.cat([torch.zeros(1, F_, 4, H_, W_), seeded_tensor("dist/img", (1, 1, 4, H_, W_)).repeat(1, F_, 1, 1, 1)])
    emb = torch.cat([torch.zeros(1, 1, 32), seeded_tensor("dist/emb", (1, 1, 32))])
    cc = seeded_tensor("dist/cc", (1, F_, 2, 8 * H_, 8 * W_))
    ref = O.denoise_loop(unet, sched, lat, img, emb, O.add_time_ids_inference(6, 127, 0.02, 1), 25, 1.0, 3.0,
                         max_steps=1, controlnet=cn, controlnet_cond=torch.cat([cc, cc]))
    assert rel(out[0], ref) < 1e-5
    # the design point the CUDA path follows: a ControlNet that sees only its own half's embedding is NOT the reference
    wrong = _run(_cfg_split_controlnet_worker, 2, False)
    assert rel(wrong[0], ref) > 1e-5


@pytest.mark.parametrize("order", ["hw_major_0272", "b_major"])
def test_cfg_pair_split_equals_unsplit_step(order):
    import oracle as O
    out = _run(_cfg_split_worker, 2, order)
    assert torch.equal(out[0], out[1])                 # latents stay replicated
    F_, H_, W_ = 4, 8, 8
    unet = fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**dict(REDUCED4, time_context_order=order))).eval()
    sched = O.EulerDiscreteScheduler(**SCHED)
    lat = seeded_tensor("dist/lat", (1, F_, 4, H_, W_))
    sched.set_timesteps(25)
    lat = lat * sched.init_noise_sigma
    img = torch.cat([torch.zeros(1, F_, 4, H_, W_), seeded_tensor("dist/img", (1, 1, 4, H_, W_)).repeat(1, F_, 1, 1, 1)])
    emb = torch.cat([torch.zeros(1, 1, 32), seeded_tensor("dist/emb", (1, 1, 32))])
    ref = O.denoise_loop(unet, sched, lat, img, emb, O.add_time_ids_inference(6, 127, 0.02, 1), 25, 1.0, 3.0,
                         max_steps=1)
    assert rel(out[0], ref) < 1e-5


# ------------------------------------------------------------------------------------------------- training all-reduce
def _dp_oracle(order="b_major"):
    import oracle as O
    unet = fill_seeded_(O.UNetSpatioTemporalConditionControlNetModel(**dict(REDUCED4, time_context_order=order)))
    O.add_lora(unet, r=4)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for n, p in unet.named_parameters():
            if "lora_" in n:
                p.copy_(torch.randn(p.shape, generator=g) * 0.05)
    for n, p in unet.named_parameters():
        p.requires_grad_("lora_" in n)
    return unet.eval()


def _dp_batch():
    F_, H_, W_ = 4, 8, 8
    return dict(lat=seeded_tensor("dp/lat", (2, F_, 4, H_, W_)), noise=seeded_tensor("dp/noise", (2, F_, 4, H_, W_)),
                cond=seeded_tensor("dp/cond", (2, 4, H_, W_)), ctx=seeded_tensor("dp/ctx", (2, 1, 32)),
                sig=torch.tensor([0.7, 2.5]))


def _dp_flat_grad(unet, b, rows):
    """Loss + flat LoRA gradient (all trainable tensors back to back, registration order) of the samples ``rows``."""
    import oracle as O
    from oracle.pipeline import train_loss, train_precondition
    lat, noise, cond, ctx, sig = (b[k][rows] for k in ("lat", "noise", "cond", "ctx", "sig"))
    noisy, ts, inp = train_precondition(lat, noise, sig)
    x = torch.cat([inp, cond.unsqueeze(1).repeat(1, lat.shape[1], 1, 1, 1)], dim=2)
    ids = O.add_time_ids_training(5, 127, 0.02, lat.shape[0])
    unet.zero_grad()
    loss = train_loss(unet(x, ts, ctx, added_time_ids=ids, return_dict=False)[0], noisy, lat, sig)
    loss.backward()
    return torch.cat([p.grad.reshape(-1) for p in unet.parameters() if p.requires_grad])


def _dp_worker(rank, world):
    """Each rank: gradient of ITS sample into one flat buffer, then the one collective of the training step."""
    from lkgd_b200.distributed import allreduce_flat_
    flat = _dp_flat_grad(_dp_oracle(), _dp_batch(), slice(rank, rank + 1)).contiguous()
    local = flat.clone()
    allreduce_flat_(flat)
    return local, flat


def test_flat_gradient_allreduce_equals_big_batch_gradient():
    """Two replicas with one sample each + one flat all-reduce (sum) + the 1/world the optimizer kernel applies ==
    the gradient of the two-sample batch on one process: the data-parallel contract of the reference's DDP wrap
    (train_models/train_svd_lora.py:1300-1302)."""
    out = _run(_dp_worker, 2)
    assert torch.equal(out[0][1], out[1][1])                                  # both ranks hold the same reduced buffer
    assert torch.allclose(out[0][1], out[0][0] + out[1][0], rtol=0, atol=0)    # it is the plain sum
    whole = _dp_flat_grad(_dp_oracle(), _dp_batch(), slice(0, 2))
    assert rel(out[0][1] / 2, whole) < 1e-5


def test_allreduce_flat_is_a_noop_without_a_group_and_rejects_views():
    from lkgd_b200.distributed import allreduce_flat_
    x = torch.arange(6.0)
    assert allreduce_flat_(x) is x and torch.equal(x, torch.arange(6.0))
