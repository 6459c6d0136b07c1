"""CPU ORACLE (test infrastructure, never the product path) for the sampler-step math that surrounds the denoiser call
(SURVEY.md §8 row a20 / §8f rank 1).  Restates, in plain numpy (tables, float64) + torch CPU fp32 (per-step arithmetic),
what the reference's `diffusion/` package does on the sampling path:

  create_diffusion(str(n))            diffusion/__init__.py:10-47     linear betas 1e-4..2e-2 over 1000 steps, eps-prediction,
                                                                      LEARNED_RANGE variance
  space_timesteps                     diffusion/respace.py:12-62      fractional-stride selection of the kept timesteps
  SpacedDiffusion.__init__            diffusion/respace.py:73-88      betas of the shortened chain from alphas_cumprod
  GaussianDiffusion.__init__ tables   diffusion/gaussian_diffusion.py:171-208
  p_mean_variance (LEARNED_RANGE)     diffusion/gaussian_diffusion.py:254-336
  p_sample / p_sample_loop            diffusion/gaussian_diffusion.py:380-516
  ddim_sample / ddim_sample_loop      diffusion/gaussian_diffusion.py:517-564, 604-689
  _WrappedModel timestep mapping      diffusion/respace.py:118-130
  q_sample, training_losses (MSE+VB)  diffusion/gaussian_diffusion.py:223-236, 686-795     (train.py:220-222; groundwork for the
  normal_kl, discretized log-lik.     diffusion/diffusion_utils.py:10-88                    training row, BASELINE config 5)
  _extract_into_tensor                diffusion/gaussian_diffusion.py:869-881   (float64 table -> fp32 value)

Pinned: tests/test_oracle_sampler.py checks every table and whole short trajectories against tests/golden/sampler_*.npz,
which oracle/make_golden_sampler.py produced by importing the UNMODIFIED reference package from /root/reference.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch


def linear_betas(num_steps: int = 1000) -> np.ndarray:
    """get_named_beta_schedule('linear', n)  (gaussian_diffusion.py:104-121): Ho et al. range scaled by 1000/n, float64."""
    scale = 1000 / num_steps
    return np.linspace(scale * 0.0001, scale * 0.02, num_steps, dtype=np.float64)


def space_timesteps(num_steps: int, section_counts) -> list:
    """respace.py:12-62.  `"ddimN"` = integer stride giving exactly N steps; otherwise comma-separated counts per equal
    section, each section sampled at a fractional stride with round-half-even (python round) positions."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[4:])
            for stride in range(1, num_steps):
                if len(range(0, num_steps, stride)) == want:
                    return sorted(set(range(0, num_steps, stride)))
            raise ValueError(f"cannot create exactly {num_steps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    per, extra = divmod(num_steps, len(section_counts))
    start, kept = 0, []
    for i, count in enumerate(section_counts):
        size = per + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        pos = 0.0
        for _ in range(count):
            kept.append(start + round(pos))
            pos += stride
        start += size
    return sorted(set(kept))


@dataclass
class Schedule:
    """The float64 tables of the SHORTENED chain (what GaussianDiffusion.__init__ holds after SpacedDiffusion re-derives
    its betas) + the map from chain index to original timestep fed to the model."""
    timestep_map: np.ndarray            # int64 [n]
    betas: np.ndarray
    alphas_cumprod: np.ndarray
    alphas_cumprod_prev: np.ndarray
    sqrt_recip_alphas_cumprod: np.ndarray
    sqrt_recipm1_alphas_cumprod: np.ndarray
    posterior_variance: np.ndarray
    posterior_log_variance_clipped: np.ndarray
    posterior_mean_coef1: np.ndarray
    posterior_mean_coef2: np.ndarray
    log_betas: np.ndarray

    @property
    def num_timesteps(self) -> int:
        return int(self.betas.shape[0])


def make_schedule(timestep_respacing, diffusion_steps: int = 1000) -> Schedule:
    base = linear_betas(diffusion_steps)
    if timestep_respacing is None or timestep_respacing == "":
        timestep_respacing = [diffusion_steps]
    use = set(space_timesteps(diffusion_steps, timestep_respacing if not isinstance(timestep_respacing, int) else str(timestep_respacing)))
    base_ac = np.cumprod(1.0 - base, axis=0)
    # respace.py:77-86: beta_i' = 1 - abar_i / abar_(previous kept)
    last, new_betas, tmap = 1.0, [], []
    for i, ac in enumerate(base_ac):
        if i in use:
            new_betas.append(1 - ac / last)
            last = ac
            tmap.append(i)
    betas = np.array(new_betas, dtype=np.float64)
    alphas = 1.0 - betas
    ac = np.cumprod(alphas, axis=0)
    ac_prev = np.append(1.0, ac[:-1])
    post_var = betas * (1.0 - ac_prev) / (1.0 - ac)
    post_logvar = np.log(np.append(post_var[1], post_var[1:])) if len(post_var) > 1 else np.array([])
    return Schedule(
        timestep_map=np.array(tmap, dtype=np.int64), betas=betas, alphas_cumprod=ac, alphas_cumprod_prev=ac_prev,
        sqrt_recip_alphas_cumprod=np.sqrt(1.0 / ac), sqrt_recipm1_alphas_cumprod=np.sqrt(1.0 / ac - 1),
        posterior_variance=post_var, posterior_log_variance_clipped=post_logvar,
        posterior_mean_coef1=betas * np.sqrt(ac_prev) / (1.0 - ac),
        posterior_mean_coef2=(1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac), log_betas=np.log(betas))


def _extract(arr: np.ndarray, t: torch.Tensor, shape) -> torch.Tensor:
    """_extract_into_tensor (gaussian_diffusion.py:869-881): float64 table -> gather -> .float() -> broadcast."""
    res = torch.from_numpy(arr)[t].float()
    while res.dim() < len(shape):
        res = res[..., None]
    return res + torch.zeros(shape)


def p_mean_variance(s: Schedule, model_output: torch.Tensor, x: torch.Tensor, t: torch.Tensor, clip_denoised: bool):
    """gaussian_diffusion.py:288-336 for EPSILON / LEARNED_RANGE given the model output (B,F,2C,H,W)."""
    C = x.shape[2]
    assert model_output.shape == (x.shape[0], x.shape[1], 2 * C, *x.shape[3:])
    eps, var_values = torch.split(model_output, C, dim=2)
    min_log = _extract(s.posterior_log_variance_clipped, t, x.shape)
    max_log = _extract(s.log_betas, t, x.shape)
    frac = (var_values + 1) / 2
    log_variance = frac * max_log + (1 - frac) * min_log
    pred_xstart = _extract(s.sqrt_recip_alphas_cumprod, t, x.shape) * x - _extract(s.sqrt_recipm1_alphas_cumprod, t, x.shape) * eps
    if clip_denoised:
        pred_xstart = pred_xstart.clamp(-1, 1)
    mean = _extract(s.posterior_mean_coef1, t, x.shape) * pred_xstart + _extract(s.posterior_mean_coef2, t, x.shape) * x
    return {"mean": mean, "variance": torch.exp(log_variance), "log_variance": log_variance, "pred_xstart": pred_xstart}


def p_sample(s: Schedule, model_output, x, t, noise, clip_denoised=True):
    """gaussian_diffusion.py:405-419."""
    out = p_mean_variance(s, model_output, x, t, clip_denoised)
    nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    return {"sample": out["mean"] + nonzero * torch.exp(0.5 * out["log_variance"]) * noise, "pred_xstart": out["pred_xstart"]}


def ddim_sample(s: Schedule, model_output, x, t, noise, clip_denoised=True, eta=0.0):
    """gaussian_diffusion.py:531-564 (eps re-derived from pred_xstart exactly as there)."""
    out = p_mean_variance(s, model_output, x, t, clip_denoised)
    eps = (_extract(s.sqrt_recip_alphas_cumprod, t, x.shape) * x - out["pred_xstart"]) / _extract(s.sqrt_recipm1_alphas_cumprod, t, x.shape)
    abar = _extract(s.alphas_cumprod, t, x.shape)
    abar_prev = _extract(s.alphas_cumprod_prev, t, x.shape)
    sigma = eta * torch.sqrt((1 - abar_prev) / (1 - abar)) * torch.sqrt(1 - abar / abar_prev)
    mean_pred = out["pred_xstart"] * torch.sqrt(abar_prev) + torch.sqrt(1 - abar_prev - sigma ** 2) * eps
    nonzero = (t != 0).float().view(-1, *([1] * (x.dim() - 1)))
    return {"sample": mean_pred + nonzero * sigma * noise, "pred_xstart": out["pred_xstart"]}


def sample_loop(s: Schedule, model, shape, noise, method="ddim", clip_denoised=True, eta=0.0, model_kwargs=None, record=None):
    """p_sample_loop / ddim_sample_loop (gaussian_diffusion.py:423-516, 604-689) with the _WrappedModel timestep mapping
    (respace.py:125-130).  `model(x, mapped_t, **kw)`; fresh `torch.randn_like` every step in BOTH methods, as there."""
    kw = model_kwargs or {}
    img = noise
    tmap = torch.from_numpy(s.timestep_map)
    for i in reversed(range(s.num_timesteps)):
        t = torch.tensor([i] * shape[0])
        with torch.no_grad():
            mo = model(img, tmap[t], **kw)
            step_noise = torch.randn_like(img)
            if method == "ddim":
                out = ddim_sample(s, mo, img, t, step_noise, clip_denoised, eta)
            else:
                out = p_sample(s, mo, img, t, step_noise, clip_denoised)
        img = out["sample"]
        if record is not None:
            record.append((img.clone(), out["pred_xstart"].clone()))
    return img


# ------------------------------------------------------------------------------------------------ training losses
def q_sample(s: Schedule, x_start, t, noise):
    """gaussian_diffusion.py:223-236: x_t = sqrt(abar_t) x_0 + sqrt(1 - abar_t) noise."""
    return (_extract(np.sqrt(s.alphas_cumprod), t, x_start.shape) * x_start
            + _extract(np.sqrt(1.0 - s.alphas_cumprod), t, x_start.shape) * noise)


def normal_kl(mean1, logvar1, mean2, logvar2):
    """diffusion_utils.py:10-37."""
    return 0.5 * (-1.0 + logvar2 - logvar1 + torch.exp(logvar1 - logvar2) + ((mean1 - mean2) ** 2) * torch.exp(-logvar2))


def _approx_std_normal_cdf(x):
    """diffusion_utils.py:40-45 (tanh approximation)."""
    return 0.5 * (1.0 + torch.tanh(np.sqrt(2.0 / np.pi) * (x + 0.044715 * torch.pow(x, 3))))


def discretized_gaussian_log_likelihood(x, means, log_scales):
    """diffusion_utils.py:63-88: log-probability of the 1/255-wide bin around x in [-1, 1]."""
    centered = x - means
    inv_stdv = torch.exp(-log_scales)
    cdf_plus = _approx_std_normal_cdf(inv_stdv * (centered + 1.0 / 255.0))
    cdf_min = _approx_std_normal_cdf(inv_stdv * (centered - 1.0 / 255.0))
    log_cdf_plus = torch.log(cdf_plus.clamp(min=1e-12))
    log_one_minus_cdf_min = torch.log((1.0 - cdf_min).clamp(min=1e-12))
    cdf_delta = cdf_plus - cdf_min
    return torch.where(x < -0.999, log_cdf_plus,
                       torch.where(x > 0.999, log_one_minus_cdf_min, torch.log(cdf_delta.clamp(min=1e-12))))


def _mean_flat(t):
    return t.mean(dim=list(range(1, t.dim())))


def vb_terms_bpd(s: Schedule, model_output, x_start, x_t, t):
    """gaussian_diffusion.py:686-716 (clip_denoised=False, as training_losses calls it): KL to the true posterior in bits,
    decoder NLL at t == 0."""
    true_mean = _extract(s.posterior_mean_coef1, t, x_t.shape) * x_start + _extract(s.posterior_mean_coef2, t, x_t.shape) * x_t
    true_logvar = _extract(s.posterior_log_variance_clipped, t, x_t.shape)
    out = p_mean_variance(s, model_output, x_t, t, clip_denoised=False)
    kl = _mean_flat(normal_kl(true_mean, true_logvar, out["mean"], out["log_variance"])) / np.log(2.0)
    nll = _mean_flat(-discretized_gaussian_log_likelihood(x_start, out["mean"], 0.5 * out["log_variance"])) / np.log(2.0)
    return torch.where(t == 0, nll, kl)


def training_losses(s: Schedule, model, x_start, t, noise, model_kwargs=None):
    """gaussian_diffusion.py:719-795 for LossType.MSE + LEARNED_RANGE (what create_diffusion builds, train.py:131):
    loss = mean((noise - eps)^2) + vb(eps.detach(), var_values); the model sees the ORIGINAL timestep (respace.py:125-130)."""
    x_t = q_sample(s, x_start, t, noise)
    mo = model(x_t, torch.from_numpy(s.timestep_map)[t], **(model_kwargs or {}))
    C = x_t.shape[2]
    eps, var_values = torch.split(mo, C, dim=2)
    frozen = torch.cat([eps.detach(), var_values], dim=2)
    vb = vb_terms_bpd(s, frozen, x_start, x_t, t)
    mse = _mean_flat((noise - eps) ** 2)
    return {"loss": mse + vb, "mse": mse, "vb": vb}


def toy_model(x: torch.Tensor, t: torch.Tensor, gain: float = 1.0) -> torch.Tensor:
    """A deterministic stand-in denoiser with the right signature / output shape (B,F,2C,H,W) for sampler-level goldens."""
    tt = (t.float() / 1000.0).view(-1, 1, 1, 1, 1)
    eps = torch.tanh(0.5 * x + tt) * gain
    var = torch.sin(3.0 * x - tt)
    return torch.cat([eps, var], dim=2)
