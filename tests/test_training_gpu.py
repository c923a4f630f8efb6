"""Training-step parity on the GPU: loss and LoRA gradients of the hand-scheduled CUDA backward
(lkgd_b200/training.py) against PyTorch autograd through the fp32 CPU oracle on identical weights and inputs -
the computation the reference's ``accelerator.backward(loss)`` performs (train_models/train_svd_lora.py:1503-1530,
1634-1642,1651-1683).  Tolerance: bf16 operands in ~400 chained GEMMs; per-tensor gradient rel-L2 <= 5e-2, all
gradients together <= 3e-2, loss <= 1e-2 relative."""
import pytest
import torch

from conftest import rel_l2
from test_unet_gpu import _pair

pytestmark = pytest.mark.gpu


def _train_inputs(B, F, h, w, xdim, seed=5):
    g = torch.Generator().manual_seed(seed)
    lat = torch.randn(B, F, 4, h, w, generator=g)
    noise = torch.randn(B, F, 4, h, w, generator=g)
    cond = torch.randn(B, 4, h, w, generator=g)
    ctx = torch.randn(B, 1, xdim, generator=g)
    sig = torch.tensor([0.8, 3.0, 0.3, 11.0][:B])
    return lat, noise, cond, ctx, sig


def _oracle_step(o, lat, noise, cond, ctx, sig, ids, extra=()):
    from oracle.pipeline import train_loss, train_precondition
    noisy, timesteps, inp = train_precondition(lat, noise, sig)
    x = torch.cat([inp, cond.unsqueeze(1).repeat(1, lat.shape[1], 1, 1, 1)], dim=2)
    for p in o.parameters():
        p.requires_grad_(False)
    for n, p in o.named_parameters():
        if "lora_" in n:
            p.requires_grad_(True)
            p.grad = None
    pred = o(x, timesteps, ctx, *extra, added_time_ids=ids, return_dict=False)[0]
    loss = train_loss(pred, noisy, lat, sig)
    loss.backward()
    return float(loss), {n: p.grad.clone() for n, p in o.named_parameters() if p.requires_grad}


def _compare(trainer, ref_grads, ref_loss, loss, tol_each=5e-2, tol_all=3e-2):
    assert abs(float(loss) - ref_loss) < 1e-2 * abs(ref_loss), (float(loss), ref_loss)
    got = dict(trainer.named_grads())
    assert set(got) == set(ref_grads), set(got) ^ set(ref_grads)
    worst, num, den = ("", 0.0), 0.0, 0.0
    for n, gref in ref_grads.items():
        e = rel_l2(got[n], gref)
        num += float((got[n].double().cpu() - gref.double()).pow(2).sum())
        den += float(gref.double().pow(2).sum())
        if gref.numel() <= 4:
            # Linear(4, 1) heads of the Nyquist bin: 1-4 numbers of magnitude 1e-7 that are sums of cancelling terms -
            # the ~1e-2 noise of the incoming gradient is amplified; a wrong formula would be off by O(1)
            assert e < 0.3, (n, e)
            continue
        if e > worst[1]:
            worst = (n, e)
    total = (num / den) ** 0.5
    print(f"loss {float(loss):.6f} vs {ref_loss:.6f}; grads: all {total:.3e}, worst {worst[1]:.3e} ({worst[0]})")
    assert worst[1] < tol_each, worst
    assert total < tol_all, total


@pytest.mark.parametrize("B,r", [(2, 8), (1, 4)])
def test_lora_gradients_reduced(cuda, B, r):
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionControlNetModel
    o, p = _pair(O.UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionControlNetModel,
                 dict(REDUCED_CONFIG), cuda, lora=dict(r=r))
    lat, noise, cond, ctx, sig = _train_inputs(B, 8, 16, 16, 32)
    ids = O.add_time_ids_training(5, 127, 0.02, B)
    ref_loss, ref_grads = _oracle_step(o, lat, noise, cond, ctx, sig, ids)
    tr = LoraTrainer(p)
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda))
    _compare(tr, ref_grads, ref_loss, loss)


def test_lora_gradients_three_levels_d64(cuda):
    """64-wide heads (the SVD head size), a stride-2 chain of two downsamplers, 5 frames, non-square latents."""
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    cfg = dict(sample_size=32, in_channels=8, out_channels=4,
               down_block_types=("CrossAttnDownBlockSpatioTemporal",) * 2 + ("DownBlockSpatioTemporal",),
               up_block_types=("UpBlockSpatioTemporal",) + ("CrossAttnUpBlockSpatioTemporal",) * 2,
               block_out_channels=(64, 128, 128), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
               layers_per_block=2, cross_attention_dim=64, transformer_layers_per_block=1,
               num_attention_heads=(1, 2, 2), num_frames=5)
    o, p = _pair(O.UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionControlNetModel, cfg, cuda,
                 lora=dict(r=16))
    lat, noise, cond, ctx, sig = _train_inputs(2, 5, 24, 40, 64)
    ids = O.add_time_ids_training(5, 127, 0.02, 2)
    ref_loss, ref_grads = _oracle_step(o, lat, noise, cond, ctx, sig, ids)
    tr = LoraTrainer(p)
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda))
    _compare(tr, ref_grads, ref_loss, loss)


@pytest.mark.parametrize("target", ["all_attention", "self_attention_out_only"])
def test_lora_gradients_on_every_attention_projection(cuda, target):
    """The reference's other adapter configuration (train_svd_lora.py:1091-1096 commented alternative,
    run_models/run_inference_flow_lora.py:326-331): LoraConfig(target_modules=["to_k","to_q","to_v","to_out.0"]) puts adapters
    on EVERY attention projection - spatial and temporal, attn1 and attn2.  Gradients of all of them against autograd through
    the oracle: fused q|k|v and out-projection sites of both self-attentions, to_v / to_out of the KV-length-1
    cross-attentions (through the per-sample cross vectors), exactly zero for the cross-attentions' to_q / to_k (one key:
    the softmax is 1 whatever the query)."""
    import oracle as O
    from oracle.lora import ALL_ATTN_PROJ
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionControlNetModel
    cfg = dict(REDUCED_CONFIG)
    torch.manual_seed(0)
    o = O.UNetSpatioTemporalConditionControlNetModel(**cfg).eval()
    p = UNetSpatioTemporalConditionControlNetModel(**cfg)
    if target == "all_attention":
        O.add_lora(o, 8, target=ALL_ATTN_PROJ)
        hit = p.add_adapter(dict(r=8, lora_alpha=8, init_lora_weights="gaussian",
                                 target_modules=["to_k", "to_q", "to_v", "to_out.0"]))
        assert len(hit) == 6 * 2 * 2 * 4
    else:
        O.add_lora(o, 4, target=r".*attn1\.to_out\.0$")
        hit = p.add_adapter(dict(r=4, lora_alpha=4, init_lora_weights="gaussian", target_modules=["attn1.to_out.0"]))
        assert len(hit) == 6 * 2
    from test_unet_gpu import _randomise_zero_inits
    _randomise_zero_inits(o)
    with torch.no_grad():
        for n, prm in o.named_parameters():
            if "lora_B" in n:
                prm.copy_((torch.randn(prm.shape, generator=torch.Generator().manual_seed(len(n))) * 0.05)
                          .to(torch.bfloat16).float())
    p.load_state_dict(o.state_dict(), strict=True)
    p = p.to(cuda)
    B = 2
    lat, noise, cond, ctx, sig = _train_inputs(B, 8, 16, 16, 32)
    ids = O.add_time_ids_training(5, 127, 0.02, B)
    ref_loss, ref_grads = _oracle_step(o, lat, noise, cond, ctx, sig, ids)
    tr = LoraTrainer(p)
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda))
    got = dict(tr.named_grads())
    dead = [n for n in ref_grads if ".attn2.to_q." in n or ".attn2.to_k." in n]
    if target == "all_attention":
        assert len(dead) == 6 * 2 * 2 * 2
    for n in dead:                                    # zero in the oracle (up to fp32 noise of a constant softmax), zero here
        assert float(ref_grads[n].abs().max()) < 1e-6 and float(got[n].abs().max()) == 0.0
        del ref_grads[n], got[n]
    tr.named_grads = lambda: list(got.items())
    _compare(tr, ref_grads, ref_loss, loss)
    # optimizer steps move every live adapter and reduce the loss
    tr2 = LoraTrainer(p, lr=2e-3, weight_decay=0.0)
    args = [t.to(cuda) for t in (lat, noise, sig, cond, ctx, ids)]
    losses = [float(tr2.train_step(*args)) for _ in range(5)]
    print(target, "losses", losses)
    assert losses[-1] < losses[0]


def test_lora_gradients_with_128_wide_heads(cuda):
    """The reference-default head layout (5,10,10,20) gives 128-wide heads at level 2
    (models/unet_spatio_temporal_condition_controlnet.py:93): forward with log-sum-exp and both backward kernels at d = 128."""
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import UNetSpatioTemporalConditionControlNetModel
    cfg = dict(sample_size=16, in_channels=8, out_channels=4,
               down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
               up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
               block_out_channels=(128, 128), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
               layers_per_block=1, cross_attention_dim=64, transformer_layers_per_block=1,
               num_attention_heads=(1, 1), num_frames=5)
    o, p = _pair(O.UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionControlNetModel, cfg, cuda,
                 lora=dict(r=8))
    lat, noise, cond, ctx, sig = _train_inputs(2, 5, 16, 24, 64)
    ids = O.add_time_ids_training(5, 127, 0.02, 2)
    ref_loss, ref_grads = _oracle_step(o, lat, noise, cond, ctx, sig, ids)
    tr = LoraTrainer(p)
    assert tr.layers[0].d == 128
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda))
    _compare(tr, ref_grads, ref_loss, loss)


def test_train_steps_reduce_loss_and_alias_parameters(cuda):
    """A few optimizer steps on one fixed batch: the loss must fall, and the module's LoRA parameters (state_dict)
    must be the tensors the optimizer kernel updates."""
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionControlNetModel
    o, p = _pair(O.UNetSpatioTemporalConditionControlNetModel, UNetSpatioTemporalConditionControlNetModel,
                 dict(REDUCED_CONFIG), cuda, lora=dict(r=8))
    lat, noise, cond, ctx, sig = (t.to(cuda) for t in _train_inputs(2, 8, 16, 16, 32))
    ids = O.add_time_ids_training(5, 127, 0.02, 2).to(cuda)
    tr = LoraTrainer(p, lr=2e-3, weight_decay=0.0, max_grad_norm=1.0)
    before = {n: v.detach().clone() for n, v in p.state_dict().items() if "lora_" in n}
    losses = [float(tr.train_step(lat, noise, sig, cond, ctx, ids)) for _ in range(6)]
    print("losses", losses)
    assert losses[-1] < losses[0]
    after = {n: v for n, v in p.state_dict().items() if "lora_" in n}
    moved = sum(float((after[n] - before[n]).abs().sum()) > 0 for n in before)
    assert moved == len(before)
    assert all(torch.isfinite(v).all() for v in after.values())
    # the forward the sampler runs sees the trained adapters: packed operands were refreshed from the parameters
    pk = p.packed()
    lay = tr.layers[0]
    ad = tr.slots[id(lay)]["t_qkv"][0]
    assert torch.equal(lay.t_qkv.lora_a[:ad["r"]], ad["pA"].to(torch.bfloat16))


def test_lkgd_quaternion_and_lora_gradients(cuda):
    """LKGD UNet: the 29 'quaternion' tensors of the latent-knowledge block are trained with the adapters
    (train_svd_lora.py:1068-1073); their gradients flow through every KV-length-1 cross-attention vector, the
    Hamilton-product layers, the rFFT magnitude / phase fuse and the iFFT."""
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionModel
    cfg = dict(REDUCED_CONFIG, cross_attention_dim=1024)
    o, p = _pair(O.UNetSpatioTemporalConditionModel, UNetSpatioTemporalConditionModel, cfg, cuda, lora=dict(r=8))
    B = 2
    lat, noise, cond, ctx, sig = _train_inputs(B, 8, 16, 16, 1024)
    g = torch.Generator().manual_seed(9)
    dom, flo = torch.randn(B, 1, 1000, generator=g), torch.randn(B, 1, 1000, generator=g)
    ids = O.add_time_ids_training(5, 127, 0.02, B)
    ref_loss, ref_grads = _oracle_step(o, lat, noise, cond, ctx, sig, ids, extra=(dom, flo))
    assert sum("quaternion" in n for n in ref_grads) == 29
    tr = LoraTrainer(p)
    loss = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda),
                               dom.to(cuda), flo.to(cuda))
    got = dict(tr.named_grads())
    for n in sorted(ref_grads):
        if "quaternion" in n:
            print(f"{n:50s} {rel_l2(got[n], ref_grads[n]):.3e}  |g| {float(ref_grads[n].norm()):.3e}")
    _compare(tr, ref_grads, ref_loss, loss, tol_each=6e-2)
    # an optimizer step moves them and the next forward sees the new values
    before = tr.flat_p.clone()
    tr.optimizer_step()
    assert float((tr.flat_p - before).abs().max()) > 0
    loss2 = tr.forward_backward(lat.to(cuda), noise.to(cuda), sig.to(cuda), cond.to(cuda), ctx.to(cuda), ids.to(cuda),
                                dom.to(cuda), flo.to(cuda))
    assert torch.isfinite(loss2)


def test_cuda_graph_step_matches_eager(cuda):
    """forward+backward captured in a CUDA graph (static shapes) must reproduce the eager step: same loss, same
    gradients up to the summation order of the atomics, and the same parameters after three optimizer steps."""
    import oracle as O
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionModel
    cfg = dict(REDUCED_CONFIG, cross_attention_dim=1024)
    B = 1
    batch = list(_train_inputs(B, 8, 16, 16, 1024))
    lat, noise, cond, ctx, sig = batch
    g = torch.Generator().manual_seed(9)
    dom, flo = torch.randn(B, 1, 1000, generator=g), torch.randn(B, 1, 1000, generator=g)
    ids = O.add_time_ids_training(5, 127, 0.02, B)
    args = [t.to(cuda) for t in (lat, noise, sig, cond, ctx, ids, dom, flo)]
    results = []
    for graphed in (False, True):
        o, p = _pair(O.UNetSpatioTemporalConditionModel, UNetSpatioTemporalConditionModel, cfg, cuda, lora=dict(r=8))
        tr = LoraTrainer(p, lr=1e-3)
        losses = [float(tr.train_step(*args))]                  # eager step (also the warm-up before capture)
        if graphed:
            tr.capture(*args)
        for _ in range(3):
            losses.append(float(tr.train_step(*args)))
        results.append((losses, tr.flat_g.clone(), tr.flat_p.clone()))
    (l0, g0, p0), (l1, g1, p1) = results
    print("eager", l0, "graph", l1)
    assert max(abs(a - b) for a, b in zip(l0, l1)) < 2e-4
    assert rel_l2(g1, g0) < 8e-2          # Adam's g / sqrt(v) amplifies atomics-order noise of near-zero gradients
    assert rel_l2(p1, p0) < 5e-3          # first Adam steps move every weight by ~lr * sign(g): noise flips a few signs


def test_trained_adapters_survive_the_wire_format(cuda, tmp_path):
    """train a few steps -> pytorch_lora_weights.safetensors (the file the reference's loop writes) -> fresh model
    without adapters -> identical forward."""
    import oracle as O
    from lkgd_b200 import lora_io
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionModel
    cfg = dict(REDUCED_CONFIG, cross_attention_dim=1024)
    o, p = _pair(O.UNetSpatioTemporalConditionModel, UNetSpatioTemporalConditionModel, cfg, cuda, lora=dict(r=8))
    B = 2
    lat, noise, cond, ctx, sig = (t.to(cuda) for t in _train_inputs(B, 8, 16, 16, 1024))
    g = torch.Generator().manual_seed(9)
    dom, flo = torch.randn(B, 1, 1000, generator=g).to(cuda), torch.randn(B, 1, 1000, generator=g).to(cuda)
    ids = O.add_time_ids_training(5, 127, 0.02, B).to(cuda)
    tr = LoraTrainer(p, lr=1e-3)
    for _ in range(3):
        tr.train_step(lat, noise, sig, cond, ctx, ids, dom, flo)
    path = lora_io.save_lora_weights(p, str(tmp_path))
    base = {k: v for k, v in o.state_dict().items() if ".lora_" not in k}       # frozen weights only
    fresh = UNetSpatioTemporalConditionModel(**cfg)
    fresh.load_state_dict({k.replace(".base_layer", ""): v for k, v in base.items()}, strict=True)
    fresh = fresh.to(cuda)
    res = lora_io.load_lora_weights(fresh, path)
    assert not res["unexpected"] and len(res["loaded"]) == 36 + 29
    x = torch.randn(B, 8, 8, 16, 16, generator=g).to(cuda)
    p.invalidate()                      # re-derive the packed operands from the trained parameters on both sides
    a = p(x, 1.3, ctx, dom, flo, added_time_ids=ids).sample
    b = fresh(x, 1.3, ctx, dom, flo, added_time_ids=ids).sample
    assert torch.equal(a, b)


def test_checkpoint_directory_round_trip(cuda, tmp_path):
    """`checkpoint-<step>` directories (reference train_svd_lora.py:1702-1748 save, :1364-1387 resume) with
    the real trainer: two steps -> save -> a fresh model + trainer resumes -> its third step equals the original's third
    step (same kernels, same Adam moments, same step count); the graph-captured trainer resumes too."""
    import oracle as O
    from lkgd_b200 import checkpoint as C
    from lkgd_b200.training import LoraTrainer
    from lkgd_b200.unet import REDUCED_CONFIG, UNetSpatioTemporalConditionModel
    cfg = dict(REDUCED_CONFIG, cross_attention_dim=1024)
    lat, noise, cond, ctx, sig = _train_inputs(2, 8, 16, 16, 1024)
    ids = O.add_time_ids_training(5, 127, 0.02, 2)
    g = torch.Generator().manual_seed(2)
    extra = (torch.randn(2, 1, 1000, generator=g), torch.randn(2, 1, 1000, generator=g))
    batch = [t.to(cuda) for t in (lat, noise, sig, cond, ctx, ids) + extra]

    def make(seed):
        _, p = _pair(O.UNetSpatioTemporalConditionModel, UNetSpatioTemporalConditionModel, cfg, cuda, lora=dict(r=4),
                     seed=seed)
        return LoraTrainer(p, lr=1e-3)

    a = make(0)
    for _ in range(2):
        a.train_step(*batch)
    path = a.save_state(str(tmp_path), global_step=80, checkpoints_total_limit=3)
    assert sorted(__import__("os").listdir(path)) == ["default", "optimizer.bin", "random_states_0.pkl", "scheduler.bin"]
    loss_a = float(a.train_step(*batch))
    b = make(0)                      # same frozen weights (same seed), fresh adapters / optimizer
    with torch.no_grad():
        b.flat_p.add_(0.05)          # make sure the restore really overwrites the trainable state
    b.repack()
    assert C.resume_from_checkpoint(b, str(tmp_path), "latest", num_update_steps_per_epoch=100) == (80, 0, 80)
    assert b.step_count == 2
    loss_b = float(b.train_step(*batch))
    # identical state in, same kernels: only the summation order of the fp32 atomics in the weight-gradient GEMM differs
    assert abs(loss_a - loss_b) <= 1e-5 * abs(loss_a)
    # (Adam normalises by sqrt(v): gradient elements near zero turn that noise into a visible parameter difference)
    assert rel_l2(b.flat_p, a.flat_p) < 5e-3 and rel_l2(b.flat_m, a.flat_m) < 1e-2 and rel_l2(b.flat_v, a.flat_v) < 1e-2
    # names / order of optimizer.bin == the module's trainable parameters
    sd = torch.load(__import__("os").path.join(path, "optimizer.bin"), weights_only=False)
    assert sd["param_names"] == [n for n, _ in b.unet.named_parameters() if "lora_" in n]
