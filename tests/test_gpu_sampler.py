"""GPU parity of the fused sampler step (b200_sampler_step through latte_b200.diffusion) against the CPU oracle, which is
itself pinned bit-for-bit to the reference (tests/test_oracle_sampler.py).
Tolerances: DDIM path is the same IEEE fp32 operations in the same order -> bit-identical for a given model output.
DDPM path goes through exp(): sample = mean + exp(0.5 logvar) * noise with expf within 2 ulp of the CPU libm and
|noise| <= ~5, |terms| <= ~4 -> 1e-6 relative + 2e-6 absolute (the sum can cancel)."""
import os

import numpy as np
import pytest
import torch

from oracle import sampler_oracle as S

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _rand_case(seed, B=2, F=16, C=4, H=32):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, F, C, H, H, generator=g)
    mo = torch.randn(B, F, 2 * C, H, H, generator=g)
    noise = torch.randn(B, F, C, H, H, generator=g)
    return x, mo, noise


@pytest.mark.parametrize("spacing,idx", [("250", [249, 120]), ("250", [0, 1]), ("ddim20", [7, 7]), ("1000", [999, 0])])
@pytest.mark.parametrize("clip", [False, True])
def test_step_matches_oracle(dev, spacing, idx, clip):
    """One step at the BASELINE latent shape (2,16,4,32,32), including t = 0 (no noise) and per-sample different t."""
    from latte_b200.diffusion import create_diffusion
    d, s = create_diffusion(spacing), S.make_schedule(spacing)
    x, mo, noise = _rand_case(len(spacing) + idx[0])
    t = torch.tensor(idx)
    model = lambda xx, tt: mo.to(dev)   # noqa: E731
    for eta in (0.0, 0.7):
        want = S.ddim_sample(s, mo, x, t, noise, clip, eta)
        got = d.ddim_sample(model, x.to(dev), t.to(dev), clip_denoised=clip, eta=eta, noise=noise.to(dev))
        assert torch.equal(got["sample"].cpu(), want["sample"])
        assert torch.equal(got["pred_xstart"].cpu(), want["pred_xstart"])
    want = S.p_sample(s, mo, x, t, noise, clip)
    got = d.p_sample(model, x.to(dev), t.to(dev), clip_denoised=clip, noise=noise.to(dev))
    assert torch.equal(got["pred_xstart"].cpu(), want["pred_xstart"])
    torch.testing.assert_close(got["sample"].cpu(), want["sample"], rtol=1e-6, atol=2e-6)
    pmv_w = S.p_mean_variance(s, mo, x, t, clip)
    pmv = d.p_mean_variance(model, x.to(dev), t.to(dev), clip_denoised=clip)
    assert torch.equal(pmv["mean"].cpu(), pmv_w["mean"]) and torch.equal(pmv["log_variance"].cpu(), pmv_w["log_variance"])
    torch.testing.assert_close(pmv["variance"].cpu(), pmv_w["variance"], rtol=1e-6, atol=0)


@pytest.mark.parametrize("dt", [torch.float16, torch.bfloat16])
def test_step_takes_16bit_model_output(dev, dt):
    """The fp16 sampling path hands the sampler a 16-bit model output (sample.py:72-75): promoted to fp32 like torch does."""
    from latte_b200.diffusion import create_diffusion
    d, s = create_diffusion("250"), S.make_schedule("250")
    x, mo, noise = _rand_case(5)
    mo16 = mo.to(dt)
    t = torch.tensor([200, 3])
    want = S.ddim_sample(s, mo16.float(), x, t, noise, False, 0.0)
    got = d.ddim_sample(lambda xx, tt: mo16.to(dev), x.to(dev), t.to(dev), clip_denoised=False, noise=noise.to(dev))
    assert torch.equal(got["sample"].cpu(), want["sample"])


@pytest.mark.parametrize("spacing", ["8", "ddim20"])
def test_ddim_loop_equals_reference_golden(dev, golden_dir, spacing):
    """The whole loop on the GPU, model = the toy denoiser evaluated on the GPU: eta = 0 does not depend on the RNG, so the
    final sample must agree with the reference trajectory up to the toy model's tanh/sin (GPU libm vs CPU, ~1e-6/step)."""
    from latte_b200.diffusion import create_diffusion
    g = np.load(os.path.join(golden_dir, f"sampler_{spacing}.npz"))
    d = create_diffusion(spacing)
    z = torch.from_numpy(g["ddim_eta0_z"]).to(dev)
    seen_t = []

    def model(x, t):
        seen_t.append(t.clone())
        return S.toy_model(x.cpu(), t.cpu()).to(dev)    # the toy model itself stays on the CPU: only the sampler is under test

    xs = [o["sample"].cpu().numpy() for o in d.ddim_sample_loop_progressive(model, z.shape, noise=z, clip_denoised=False, device=dev)]
    assert np.array_equal(np.stack(xs), g["ddim_eta0_x"])          # bit-identical trajectory
    assert torch.stack(seen_t)[:, 0].cpu().tolist() == g["timestep_map"][::-1].tolist()   # original timesteps, descending
    out = d.ddim_sample_loop(model, z.shape, noise=z, clip_denoised=False, device=dev)
    assert np.array_equal(out.cpu().numpy(), g["ddim_eta0_x"][-1])


def test_ddpm_loop_statistics(dev):
    """p_sample_loop with the device RNG: cannot be compared draw by draw with the CPU generator, so check the update
    against the oracle step by step (same noise fed to both) inside a real loop."""
    from latte_b200.diffusion import create_diffusion
    d, s = create_diffusion("8"), S.make_schedule("8")
    torch.manual_seed(3)
    x = torch.randn(2, 4, 4, 8, 8)
    xd = x.to(dev)
    for i in reversed(range(8)):
        t = torch.tensor([i, i])
        mo = S.toy_model(x, torch.from_numpy(s.timestep_map)[t])
        noise = torch.randn_like(x)
        want = S.p_sample(s, mo, x, t, noise, True)
        got = d.p_sample(lambda xx, tt: mo.to(dev), xd, t.to(dev), clip_denoised=True, noise=noise.to(dev))
        torch.testing.assert_close(got["sample"].cpu(), want["sample"], rtol=1e-6, atol=2e-6)
        x, xd = want["sample"], want["sample"].to(dev)


def test_sampler_drives_latte_module(dev, golden_dir):
    """sample.py:100-107 on this repo's two halves: create_diffusion('8').ddim_sample_loop(model.forward_with_cfg, ...)."""
    from latte_b200 import Latte
    from latte_b200.diffusion import create_diffusion
    from oracle import latte_oracle as O
    cfg = O.make_config("Latte-tiny72/2", input_size=16, num_frames=16)
    sd = O.make_weights(cfg, 5)
    net = Latte(input_size=16, hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, num_frames=16,
                num_classes=cfg.num_classes, extras=2)
    net.load_state_dict(sd)
    net = net.to(dev).eval()
    torch.manual_seed(0)
    z = torch.randn(1, 16, 4, 16, 16)
    zz = torch.cat([z, z], 0)
    y = torch.tensor([3, cfg.num_classes])
    d = create_diffusion("8")
    out = d.ddim_sample_loop(net.forward_with_cfg, zz.shape, zz.to(dev), clip_denoised=False,
                             model_kwargs=dict(y=y.to(dev), cfg_scale=4.0), device=dev)
    # CPU truth: oracle model + oracle sampler, same noise-free DDIM trajectory
    s = S.make_schedule("8")
    ref = S.sample_loop(s, lambda x, t, **kw: O.latte_forward_with_cfg(sd, cfg, x, t, y, 4.0), zz.shape, zz, method="ddim",
                        clip_denoised=False)
    assert out.shape == zz.shape and torch.isfinite(out).all()
    # 8 model calls with fp16 operands (per-call error ~2e-3 of the output scale); with random weights and no clipping the
    # trajectory grows to |x| ~ 4e2, so compare relative to the sample's scale
    assert (out.cpu() - ref).abs().max().item() < 5e-3 * ref.abs().max().item()


def test_trajectory_conditioning_is_bit_identical(dev, monkeypatch):
    """conditioning rows for all steps evaluated before the loop (b200_latte_conditioning +
    b200_latte_forward_conditioned) give bit-identical model outputs and bit-identical trajectories."""
    from latte_b200 import Latte
    from latte_b200.diffusion import create_diffusion
    from oracle import latte_oracle as O
    cfg = O.make_config("Latte-tiny72/2", input_size=16, num_frames=16)
    net = Latte(input_size=16, hidden_size=cfg.hidden_size, depth=cfg.depth, num_heads=cfg.num_heads, num_frames=16,
                num_classes=cfg.num_classes, extras=2)
    net.load_state_dict(O.make_weights(cfg, 9))
    net = net.to(dev).eval()
    x, t, y = (v.to(dev) for v in O.make_inputs(cfg, 2, 10))
    steps = torch.tensor([[999, 999], [500, 500], [int(t[0]), int(t[1])]], device=dev)
    with torch.no_grad():
        plain = net(x, t, y=y)
        plain_cfg = net.forward_with_cfg(x, t, y=y, cfg_scale=3.0)
        traj = net.precompute_conditioning(steps, y)
        assert traj.shape[:2] == (3, 2)
        assert torch.equal(net(x, t, y=y, trajectory_step=2), plain)
        assert torch.equal(net.forward_with_cfg(x, t, y=y, cfg_scale=3.0, trajectory_step=2), plain_cfg)
        assert not torch.equal(net(x, t, y=y, trajectory_step=0), plain)      # a different timestep's rows
        net.clear_conditioning()
        assert torch.equal(net(x, t, y=y, trajectory_step=2), plain)          # cache dropped -> (t, y) path
        d = create_diffusion("8")
        zz = torch.cat([x[:1], x[:1]], 0)
        kw = dict(y=y, cfg_scale=4.0)
        a = d.ddim_sample_loop(net.forward_with_cfg, zz.shape, zz, clip_denoised=False, model_kwargs=kw, device=dev)
        assert net._trajectory is None                                          # dropped when the loop ends
        monkeypatch.setenv("B200_NO_TRAJECTORY_CONDITIONING", "1")
        b = d.ddim_sample_loop(net.forward_with_cfg, zz.shape, zz, clip_denoised=False, model_kwargs=kw, device=dev)
    assert torch.equal(a, b)
