"""Generates tests/golden/reference_golden.npz by RUNNING THE REFERENCE'S OWN FILES from /root/reference:

  utils/scheduling_euler_discrete_karras_fix.py            (EulerDiscreteScheduler, unmodified)
  models/unet_spatio_temporal_condition_controlnet.py      (UNetSpatioTemporalConditionControlNetModel)
  models/unet_spatio_temporal_condition.py                 (LKGD UNetSpatioTemporalConditionModel)
  models/controlnet_sdv.py                                 (ControlNetSDVModel)
  models/lora_layer.py                                     (peft-copy LoRA Linear)
  pipeline/pipeline_stable_video_diffusion_controlnet.py   (StableVideoDiffusionPipelineControlNet.__call__)

Their un-vendored dependencies (diffusers 0.27.2 / peft 0.10 / core_qnn) are replaced by tests/golden/ref_shim
(plumbing re-stated; block arithmetic delegated to oracle/ - see ref_shim/README.md).  Run in the dev container
only: /root/reference does not exist on the GPU box, the committed .npz does.

    python tests/golden/make_reference_golden.py
"""
import hashlib
import pathlib
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parent.parent
sys.path[:0] = [str(HERE / "ref_shim"), "/root/reference", str(ROOT), str(HERE)]

from weights import fill_seeded_, seeded_tensor  # noqa: E402

torch.set_num_threads(8)
torch.manual_seed(0)
out = {}

REDUCED = dict(
    sample_size=32, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(32, 64), addition_time_embed_dim=32, projection_class_embeddings_input_dim=96,
    layers_per_block=2, cross_attention_dim=32, transformer_layers_per_block=1, num_attention_heads=(2, 4),
    num_frames=4)
SCHED = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
             prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
             sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1)

# ------------------------------------------------------------------------------------------ G1 scheduler
from utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler  # noqa: E402

for n in (25, 10):
    s = EulerDiscreteScheduler(**SCHED)
    out[f"sched/init_sigmas_head"] = s.sigmas[:8].numpy()
    out[f"sched/init_timesteps_head"] = s.timesteps[:8].numpy()
    s.set_timesteps(n)
    out[f"sched{n}/sigmas"] = s.sigmas.numpy()
    out[f"sched{n}/timesteps"] = s.timesteps.numpy()
    out[f"sched{n}/init_noise_sigma"] = np.asarray(float(s.init_noise_sigma), dtype=np.float32)
    x = seeded_tensor("sched/x0", (1, 3, 4, 8, 8)) * s.init_noise_sigma
    xs, scaled, x0s = [], [], []
    for i, t in enumerate(s.timesteps):
        xin = s.scale_model_input(x, t)
        v = seeded_tensor(f"sched/v{i}", x.shape)          # stand-in model output (v-prediction)
        o = s.step(v, t, x)
        scaled.append(xin.numpy()), xs.append(o.prev_sample.numpy()), x0s.append(o.pred_original_sample.numpy())
        x = o.prev_sample
    out[f"sched{n}/scaled"], out[f"sched{n}/traj"], out[f"sched{n}/x0"] = map(np.stack, (scaled, xs, x0s))
s = EulerDiscreteScheduler(**SCHED)     # add_noise on the construction-time ("fix") schedule, continuous timesteps
idx = torch.tensor([0, 17, 500, 999])
orig, noise = seeded_tensor("sched/orig", (4, 2, 4, 4, 4)), seeded_tensor("sched/noise", (4, 2, 4, 4, 4))
out["sched/add_noise_t"] = s.timesteps[idx].numpy()
out["sched/add_noise"] = s.add_noise(orig, noise, s.timesteps[idx]).numpy()

# ------------------------------------------------------------------------------------------ G2/G3 plain UNet
from models.unet_spatio_temporal_condition_controlnet import UNetSpatioTemporalConditionControlNetModel  # noqa: E402

B, F, H, W = 2, 4, 16, 16
unet = fill_seeded_(UNetSpatioTemporalConditionControlNetModel(**REDUCED)).eval()
sample = seeded_tensor("unet/sample", (B, F, 8, H, W))
ctx = seeded_tensor("unet/ctx", (B, 1, 32))
ids = torch.tensor([[6.0, 127.0, 0.02]] * B)
t = torch.tensor(1.4439898729)
with torch.no_grad():
    out["unet/out"] = unet(sample, t, ctx, added_time_ids=ids, return_dict=False)[0].numpy()
    out["unet/out_float_t"] = unet(sample, 0.75, ctx, added_time_ids=ids).sample.numpy()
    # F6: residuals for the 6 skips of the 2-level config + mid
    shapes = [(B * F, 32, 16, 16)] * 3 + [(B * F, 32, 8, 8)] + [(B * F, 64, 8, 8)] * 2
    res = [seeded_tensor(f"unet/res{i}", s, scale=0.5) for i, s in enumerate(shapes)]
    mid = seeded_tensor("unet/resmid", (B * F, 64, 8, 8), scale=0.5)
    out["unet/out_residuals"] = unet(sample, t, ctx, down_block_additional_residuals=res,
                                     mid_block_additional_residual=mid, added_time_ids=ids).sample.numpy()
out["unet/n_params"] = np.asarray(sum(p.numel() for p in unet.parameters()))
names = sorted(n for n, _ in unet.named_parameters())
out["unet/param_names_sha256"] = np.frombuffer(hashlib.sha256("\n".join(names).encode()).digest(), dtype=np.uint8)

# ------------------------------------------------------------------------------------------ G4 LKGD UNet
from models.unet_spatio_temporal_condition import UNetSpatioTemporalConditionModel as LKGDUNet  # noqa: E402

lk = fill_seeded_(LKGDUNet(**dict(REDUCED, cross_attention_dim=1024))).eval()
ctx_lk = seeded_tensor("lkgd/ctx", (B, 1, 1024))
dom, flo = seeded_tensor("lkgd/domain", (1, 1, 1000)), seeded_tensor("lkgd/flow", (1, 1, 1000))
seen = {}
h = lk.mid_block.register_forward_pre_hook(
    lambda m, a, k: seen.__setitem__("ctx", k["encoder_hidden_states"].detach().clone()), with_kwargs=True)
with torch.no_grad():
    out["lkgd/out"] = lk(sample, t, ctx_lk, dom, flo, added_time_ids=ids).sample.numpy()           # D8: batch-1 dup
    out["lkgd/context"] = seen["ctx"].reshape(B, F, 1024)[:, 0].numpy()                            # [B,1024]
    dom2, flo2 = seeded_tensor("lkgd/domain2", (B, 1, 1000)), seeded_tensor("lkgd/flow2", (B, 1, 1000))
    out["lkgd/out_b2"] = lk(sample, t, ctx_lk, dom2, flo2, added_time_ids=ids).sample.numpy()
    out["lkgd/context_b2"] = seen["ctx"].reshape(B, F, 1024)[:, 0].numpy()
h.remove()

# ------------------------------------------------------------------------------------------ G5 ControlNet
from models.controlnet_sdv import ControlNetSDVModel  # noqa: E402

cn_cfg = dict(REDUCED)   # the reference ctor validates up_block_types although it builds no up blocks
cn = fill_seeded_(ControlNetSDVModel(**cn_cfg, conditioning_channels=2), seed=1).eval()
cond = seeded_tensor("cn/cond", (B, F, 2, 8 * H, 8 * W)).clamp(-1, 1)
with torch.no_grad():
    down, midr = cn(sample, t, ctx, ids, controlnet_cond=cond, conditioning_scale=0.7, return_dict=False)
for i, d in enumerate(down):
    out[f"cn/down{i}"] = d.numpy()
out["cn/mid"] = midr.numpy()
out["cn/n_down"] = np.asarray(len(down))

# ------------------------------------------------------------------------------------------ G6 LoRA
from models.lora_layer import Linear as RefLoraLinear  # noqa: E402

base = torch.nn.Linear(32, 48)
lora = fill_seeded_(RefLoraLinear(base, "default", r=4, lora_alpha=4, init_lora_weights="gaussian"), seed=2)
xl = seeded_tensor("lora/x", (5, 7, 32))
with torch.no_grad():
    out["lora/y"] = lora(xl).numpy()
    out["lora/delta"] = lora.get_delta_weight("default").numpy()
    lora.merge()
    out["lora/merged_weight"] = lora.base_layer.weight.detach().numpy().copy()
    out["lora/y_merged"] = lora(xl).numpy()
lora8 = RefLoraLinear(torch.nn.Linear(32, 48), "default", r=8, lora_alpha=4)
out["lora/scaling_r8_a4"] = np.asarray(lora8.scaling["default"], dtype=np.float32)
out["lora/B_default_is_zero"] = np.asarray(float(lora8.lora_B["default"].weight.abs().max()))

# ------------------------------------------------------------------------------------------ G7 pipeline loop
from pipeline.pipeline_stable_video_diffusion_controlnet import StableVideoDiffusionPipelineControlNet  # noqa: E402


class _Enc(torch.nn.Module):
    def __init__(self, emb):
        super().__init__()
        self.p = torch.nn.Parameter(torch.zeros(1))
        self.emb = emb

    def forward(self, x):
        return SimpleNamespace(image_embeds=self.emb.expand(x.shape[0], -1))


clip_emb = seeded_tensor("pipe/clip", (1, 32))
vae_lat = seeded_tensor("pipe/vae", (1, 4, H, W))
vae = SimpleNamespace(config=SimpleNamespace(block_out_channels=(1, 2, 3, 4), force_upcast=False, scaling_factor=0.18215),
                      dtype=torch.float32,
                      encode=lambda im: SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: vae_lat.clone())))
fe = lambda images, **kw: SimpleNamespace(pixel_values=images)  # noqa: E731
pipe = StableVideoDiffusionPipelineControlNet(vae=vae, image_encoder=_Enc(clip_emb), unet=unet, controlnet=cn,
                                              scheduler=EulerDiscreteScheduler(**SCHED), feature_extractor=fe)
image = seeded_tensor("pipe/image", (1, 3, 8 * H, 8 * W)).clamp(-1, 1)
cond_p = seeded_tensor("pipe/cond", (F, 2, 8 * H, 8 * W)).clamp(-1, 1)
lat0 = seeded_tensor("pipe/latents", (1, F, 4, H, W))
steps = []
res = pipe(image, controlnet_condition=cond_p, height=8 * H, width=8 * W, num_frames=F, num_inference_steps=6,
           min_guidance_scale=1.0, max_guidance_scale=3.0, fps=7, motion_bucket_id=127, noise_aug_strength=0.02,
           latents=lat0, output_type="latent", controlnet_cond_scale=0.7,
           callback_on_step_end=lambda p, i, t_, kw: (steps.append(kw["latents"].numpy().copy()), kw)[1])
out["pipe/final"] = res.frames.numpy()
out["pipe/steps"] = np.stack(steps)
out["pipe/guidance"] = pipe.guidance_scale.numpy()
out["pipe/image_embeddings"] = torch.cat([torch.zeros_like(clip_emb), clip_emb]).unsqueeze(1).numpy()
out["pipe/image_latents"] = torch.cat([torch.zeros_like(vae_lat), vae_lat]).unsqueeze(1).repeat(1, F, 1, 1, 1).numpy()

np.savez_compressed(HERE / "reference_golden.npz", **out)
print("wrote", HERE / "reference_golden.npz", {k: v.shape for k, v in out.items()})
