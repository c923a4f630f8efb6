"""CPU tests of the ``checkpoint-<step>`` directory format (lkgd_b200/checkpoint.py; reference
train_models/train_svd_lora.py:1364-1387 resume, :1702-1748 save): directory naming / rotation / resume arithmetic,
``optimizer.bin`` in the exact shape of ``torch.optim.AdamW.state_dict()`` - a real torch AdamW over the same parameters
loads it and continues identically - and the adapter file inside the checkpoint."""
import os
import pickle
from types import SimpleNamespace

import pytest
import torch

from golden_util import REDUCED4


class _FakeTrainer:
    """The state a LoraTrainer exposes (flat fp32 parameter / moment buffers, views per tensor), on the CPU."""

    def __init__(self, unet, lr=1e-3):
        self.unet, self.lr, self.betas, self.eps, self.wd, self.step_count = unet, lr, (0.9, 0.999), 1e-8, 1e-2, 0
        self.params = [(n, p) for n, p in unet.named_parameters() if "lora_" in n]
        n = sum(p.numel() for _, p in self.params)
        self.flat_m, self.flat_v = torch.zeros(n), torch.zeros(n)
        self.views, off = [], 0
        for name, p in self.params:
            k = p.numel()
            self.views.append((name, p.data, self.flat_m[off:off + k].view(p.shape), self.flat_v[off:off + k].view(p.shape)))
            off += k

    def state_tensors(self):
        return self.views

    def adamw_step(self, grads):
        """torch.optim.AdamW's update rule on the flat state (what lkgd_adamw computes on the GPU)."""
        self.step_count += 1
        b1, b2 = self.betas
        for (_, p, m, v), g in zip(self.views, grads):
            p.mul_(1 - self.lr * self.wd)
            m.mul_(b1).add_(g, alpha=1 - b1)
            v.mul_(b2).addcmul_(g, g, value=1 - b2)
            mh, vh = m / (1 - b1 ** self.step_count), v / (1 - b2 ** self.step_count)
            p.addcdiv_(mh, vh.sqrt() + self.eps, value=-self.lr)


def _unet(seed=0):
    from lkgd_b200.unet import UNetSpatioTemporalConditionModel
    torch.manual_seed(seed)
    u = UNetSpatioTemporalConditionModel(**dict(REDUCED4, cross_attention_dim=1024))
    u.add_lora(4)
    with torch.no_grad():
        for n, p in u.named_parameters():
            if "lora_" in n:
                p.copy_(torch.randn(p.shape) * 0.1)
    return u


def test_directory_naming_rotation_and_resume_arithmetic(tmp_path):
    from lkgd_b200 import checkpoint as C
    out = str(tmp_path)
    for s in (40, 120, 80, 1000):
        os.makedirs(os.path.join(out, f"checkpoint-{s}"))
    os.makedirs(os.path.join(out, "logs"))
    assert C.list_checkpoints(out) == ["checkpoint-40", "checkpoint-80", "checkpoint-120", "checkpoint-1000"]   # numeric
    assert C.latest_checkpoint(out) == "checkpoint-1000"
    assert C.rotate_checkpoints(out, 3) == ["checkpoint-40", "checkpoint-80"]        # at most limit - 1 remain (:1714-1717)
    assert C.list_checkpoints(out) == ["checkpoint-120", "checkpoint-1000"]
    assert C.rotate_checkpoints(out, None) == []
    # :1381-1387 with gradient_accumulation_steps = 4, 250 optimizer steps per epoch
    assert C.resume_position("checkpoint-1000", 4, 250) == (1000, 4, 0)
    assert C.resume_position(os.path.join(out, "checkpoint-120"), 4, 250) == (120, 0, 480)
    assert C.latest_checkpoint(str(tmp_path / "nothing")) is None


def test_checkpoint_round_trip_and_torch_adamw_compatibility(tmp_path):
    from safetensors import safe_open
    from lkgd_b200 import checkpoint as C
    u = _unet()
    tr = _FakeTrainer(u)
    g = torch.Generator().manual_seed(1)
    grads = [[torch.randn(p.shape, generator=g) for _, p in tr.params] for _ in range(3)]
    # a REAL torch AdamW over the same parameters, same order, same gradients
    twin = [torch.nn.Parameter(p.detach().clone()) for _, p in tr.params]
    opt = torch.optim.AdamW(twin, lr=tr.lr, betas=tr.betas, eps=tr.eps, weight_decay=tr.wd)
    for k in range(2):
        tr.adamw_step(grads[k])
        for t, gg in zip(twin, grads[k]):
            t.grad = gg.clone()
        opt.step()
    path = C.save_state(tr, str(tmp_path), global_step=80, lora_name="default", checkpoints_total_limit=2)
    assert os.path.basename(path) == "checkpoint-80"
    assert sorted(os.listdir(path)) == ["default", "optimizer.bin", "random_states_0.pkl", "scheduler.bin"]
    with safe_open(os.path.join(path, "default", "pytorch_lora_weights.safetensors"), framework="pt") as f:
        assert len(list(f.keys())) == 65 and f.metadata() == {"format": "pt"}
    rng = pickle.load(open(os.path.join(path, "random_states_0.pkl"), "rb"))
    assert {"step", "random_state", "numpy_random_seed", "torch_manual_seed"} <= set(rng)
    # (1) the file is a torch AdamW state_dict: a fresh torch optimizer loads it and equals the live one
    sd = torch.load(os.path.join(path, "optimizer.bin"), weights_only=False)
    ref_sd = opt.state_dict()
    assert set(sd["state"]) == set(ref_sd["state"]) and sd["param_groups"][0]["params"] == ref_sd["param_groups"][0]["params"]
    for i in ref_sd["state"]:
        assert float(sd["state"][i]["step"]) == float(ref_sd["state"][i]["step"]) == 2.0
        assert torch.allclose(sd["state"][i]["exp_avg"], ref_sd["state"][i]["exp_avg"], rtol=1e-5, atol=1e-7)
        assert torch.allclose(sd["state"][i]["exp_avg_sq"], ref_sd["state"][i]["exp_avg_sq"], rtol=1e-5, atol=1e-9)
    opt2 = torch.optim.AdamW([torch.nn.Parameter(t.detach().clone()) for t in twin], lr=1.0)
    opt2.load_state_dict({k: v for k, v in sd.items() if k != "param_names"})
    assert opt2.param_groups[0]["lr"] == tr.lr and opt2.param_groups[0]["weight_decay"] == tr.wd
    # (2) resume into a fresh model + trainer: parameters, moments and step come back; the next step is identical
    u2 = _unet(seed=5)
    tr2 = _FakeTrainer(u2, lr=123.0)
    pos = C.resume_from_checkpoint(tr2, str(tmp_path), "latest", gradient_accumulation_steps=1, num_update_steps_per_epoch=50)
    assert pos == (80, 1, 30)
    assert tr2.step_count == 2 and tr2.lr == tr.lr
    for (n1, p1, m1, v1), (n2, p2, m2, v2) in zip(tr.views, tr2.views):
        assert n1 == n2 and torch.equal(p1, p2) and torch.equal(m1, m2) and torch.equal(v1, v2)
    tr.adamw_step(grads[2])
    tr2.adamw_step(grads[2])
    assert all(torch.equal(a[1], b[1]) for a, b in zip(tr.views, tr2.views))
    # a third checkpoint rotates the oldest away (limit 2)
    C.save_state(tr, str(tmp_path), global_step=120, checkpoints_total_limit=2)
    C.save_state(tr, str(tmp_path), global_step=160, checkpoints_total_limit=2)
    assert C.list_checkpoints(str(tmp_path)) == ["checkpoint-120", "checkpoint-160"]
    with pytest.raises(ValueError, match="different set"):
        bad = dict(sd, param_names=list(reversed(sd["param_names"])))
        C.load_adamw_state_dict(tr2, bad)
    assert C.resume_from_checkpoint(tr2, str(tmp_path / "empty")) is None
