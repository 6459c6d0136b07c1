"""Generate tests/golden/sampler_*.npz from the UNMODIFIED reference `diffusion` package (/root/reference/diffusion).
Run in the build container only (the reference does not travel to the GPU box):   python oracle/make_golden_sampler.py

Stored per respacing ("250", "8", "ddim20"): every float64 table of the SpacedDiffusion object, its timestep_map, and for
the short chains the full trajectory (x after every step and pred_xstart) of ddim_sample_loop (eta 0 and 0.5) and
p_sample_loop driven by oracle.sampler_oracle.toy_model with torch.manual_seed(seed) fixing the per-step noise draws."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
ref = importlib.import_module("diffusion")           # the reference package, untouched
from oracle.sampler_oracle import toy_model          # noqa: E402

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
          "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2"]


def trajectory(d, method, shape, seed, eta, clip):
    torch.manual_seed(seed)
    z = torch.randn(*shape)
    fn = d.ddim_sample_loop_progressive if method == "ddim" else d.p_sample_loop_progressive
    kw = dict(noise=z, clip_denoised=clip, model_kwargs={}, device="cpu")
    if method == "ddim":
        kw["eta"] = eta
    xs, x0s = [], []
    for out in fn(toy_model, shape, **kw):
        xs.append(out["sample"].numpy().copy())
        x0s.append(out["pred_xstart"].numpy().copy())
    return z.numpy(), np.stack(xs), np.stack(x0s)


def main():
    out_dir = os.path.join(ROOT, "tests", "golden")
    for spacing in ["250", "8", "ddim20"]:
        d = ref.create_diffusion(spacing)
        blob = {k: np.asarray(getattr(d, k), dtype=np.float64) for k in TABLES}
        blob["timestep_map"] = np.asarray(d.timestep_map, dtype=np.int64)
        blob["log_betas"] = np.log(d.betas)
        if spacing != "250":
            shape = (2, 3, 4, 8, 8)
            for name, method, eta, clip in [("ddim_eta0", "ddim", 0.0, False), ("ddim_eta05_clip", "ddim", 0.5, True),
                                            ("ddpm", "ddpm", 0.0, False), ("ddpm_clip", "ddpm", 0.0, True)]:
                z, xs, x0s = trajectory(d, method, shape, 1234, eta, clip)
                blob[f"{name}_z"], blob[f"{name}_x"], blob[f"{name}_x0"] = z, xs, x0s
        if spacing == "250":
            # training_losses on the UNSPACED chain (train.py:131 create_diffusion(timestep_respacing="")), t = 0 included
            dt = ref.create_diffusion("")
            torch.manual_seed(77)
            x0 = torch.randn(4, 3, 4, 8, 8).clamp(-1, 1)
            noise = torch.randn_like(x0)
            t = torch.tensor([0, 1, 500, 999])
            terms = dt.training_losses(toy_model, x0, t, model_kwargs={}, noise=noise)
            blob.update(train_x0=x0.numpy(), train_noise=noise.numpy(), train_t=t.numpy(),
                        **{f"train_{k}": v.detach().numpy() for k, v in terms.items()})
        path = os.path.join(out_dir, f"sampler_{spacing}.npz")
        np.savez_compressed(path, **blob)
        print("wrote", path, {k: v.shape for k, v in blob.items() if k.endswith("_x")})


if __name__ == "__main__":
    main()
