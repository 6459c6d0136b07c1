"""CPU: the sampler oracle (oracle/sampler_oracle.py) is pinned to goldens produced by the UNMODIFIED reference
`diffusion` package (oracle/make_golden_sampler.py), and the host-side mirror (latte_b200/diffusion) builds the same
float64 schedule tables.  No GPU, no compute through the C ABI."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as S

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
          "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2", "log_betas"]
TRAJ = [("ddim_eta0", "ddim", 0.0, False), ("ddim_eta05_clip", "ddim", 0.5, True), ("ddpm", "ddpm", 0.0, False),
        ("ddpm_clip", "ddpm", 0.0, True)]


def _golden(golden_dir, spacing):
    return np.load(os.path.join(golden_dir, f"sampler_{spacing}.npz"))


@pytest.mark.parametrize("spacing", ["250", "8", "ddim20"])
def test_schedule_tables_equal_reference(golden_dir, spacing):
    """create_diffusion(spacing): every float64 table and the timestep map, bit for bit (respace.py:73-88, gaussian_diffusion.py:171-208)."""
    g, s = _golden(golden_dir, spacing), S.make_schedule(spacing)
    for k in TABLES:
        assert np.array_equal(getattr(s, k), g[k]), k
    assert np.array_equal(s.timestep_map, g["timestep_map"])
    if spacing == "250":   # fractional stride 999/249
        assert s.timestep_map[:3].tolist() == [0, 4, 8] and s.timestep_map[-2:].tolist() == [995, 999]


@pytest.mark.parametrize("spacing", ["8", "ddim20"])
@pytest.mark.parametrize("case", TRAJ, ids=[c[0] for c in TRAJ])
def test_trajectories_equal_reference(golden_dir, spacing, case):
    """Whole sampling loops (DDIM eta 0 / 0.5, DDPM, with and without clipping) driven by the toy model: x after every step
    and pred_xstart are bit-identical to the reference's ddim_sample_loop_progressive / p_sample_loop_progressive."""
    name, method, eta, clip = case
    g, s = _golden(golden_dir, spacing), S.make_schedule(spacing)
    torch.manual_seed(1234)
    z = torch.randn(2, 3, 4, 8, 8)
    assert np.array_equal(z.numpy(), g[name + "_z"])
    rec = []
    S.sample_loop(s, S.toy_model, z.shape, z, method=method, clip_denoised=clip, eta=eta, record=rec)
    assert np.array_equal(np.stack([r[0].numpy() for r in rec]), g[name + "_x"])
    assert np.array_equal(np.stack([r[1].numpy() for r in rec]), g[name + "_x0"])


@pytest.mark.parametrize("spacing", ["250", "8", "ddim20", "1000"])
def test_mirror_builds_the_same_tables(golden_dir, spacing):
    """latte_b200.diffusion.create_diffusion (the product's host side) == oracle == reference tables."""
    from latte_b200.diffusion import create_diffusion
    d, s = create_diffusion(spacing), S.make_schedule(spacing)
    assert d.timestep_map == s.timestep_map.tolist() and d.num_timesteps == s.num_timesteps
    for k in TABLES[:-1]:
        assert np.array_equal(getattr(d, k), getattr(s, k)), k


def test_mirror_rejects_what_is_not_built():
    from latte_b200.diffusion import create_diffusion
    with pytest.raises(NotImplementedError):
        create_diffusion("250", learn_sigma=False)
    with pytest.raises(NotImplementedError):
        create_diffusion("250", predict_xstart=True)
    d = create_diffusion("8")
    with pytest.raises(RuntimeError):   # no CPU path
        d.ddim_sample(lambda x, t: torch.cat([x, x], 2), torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, dtype=torch.long))


def test_training_losses_equal_reference(golden_dir):
    """Groundwork for the training row (BASELINE config 5): training_losses (MSE + learned-range VB, gaussian_diffusion.py:
    719-795) on the unspaced chain of train.py:131, t = 0 (decoder NLL branch) included — bit-identical loss / mse / vb."""
    g = _golden(golden_dir, "250")
    s = S.make_schedule("")
    assert s.num_timesteps == 1000 and s.timestep_map.tolist() == list(range(1000))
    out = S.training_losses(s, S.toy_model, torch.from_numpy(g["train_x0"]), torch.from_numpy(g["train_t"]),
                            torch.from_numpy(g["train_noise"]))
    for k in ("loss", "mse", "vb"):
        assert np.array_equal(out[k].numpy(), g["train_" + k]), k
    # the product's own training_losses (latte_b200/diffusion: torch elementwise math around the native denoiser) on the same
    # toy model: same numbers up to the association of a few fp32 products
    from latte_b200.diffusion import create_diffusion
    got = create_diffusion(timestep_respacing="").training_losses(
        S.toy_model, torch.from_numpy(g["train_x0"]), torch.from_numpy(g["train_t"]), noise=torch.from_numpy(g["train_noise"]))
    for k in ("loss", "mse", "vb"):
        np.testing.assert_allclose(got[k].numpy(), g["train_" + k], rtol=2e-5, atol=1e-6, err_msg=k)
