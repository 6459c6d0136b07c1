// One fused kernel for the sampler-step arithmetic that surrounds the denoiser call:
//   GaussianDiffusion.p_mean_variance   diffusion/gaussian_diffusion.py:288-336   (EPSILON mean, LEARNED_RANGE variance)
//   GaussianDiffusion.p_sample          :405-419        x_{t-1} = mean + [t != 0] exp(0.5 logvar) noise
//   GaussianDiffusion.ddim_sample       :531-564        eps re-derived from pred_xstart, sigma(eta), Eq. 12
//   _extract_into_tensor                :869-881        float64 tables -> fp32 values  (done once, tables live on the device)
// The reference spends ~40 tiny launches and ~12 pageable H2D table copies per step here; this is one launch that reads
// x_t (fp32), the model output (fp32 / fp16 / bf16, (B,F,2C,H,W)) and optionally the step noise, and writes x_{t-1} and
// pred_xstart.  HBM-bound: 4 (x) + 2*{4|2} (eps, v) + 4 (noise) + 8 (outputs) bytes per latent element.
//
// Arithmetic is fp32 in the reference's operation order with FMA contraction disabled (__fmul_rn / __fadd_rn ...).  Every
// per-timestep scalar -- including the DDIM coefficients sqrt(abar_prev), sigma(eta) and sqrt(1 - abar_prev - sigma^2) --
// comes from a table the caller evaluated once with the reference's own fp32 expressions, so the kernel contains only
// IEEE mul/add/sub/div (+ one expf for DDPM): the DDIM path is bit-identical to the torch fp32 restatement, the DDPM
// path differs only by expf's last ulps.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.h"

namespace b200 {
namespace {

struct SamplerDev {
  B200SamplerTables tab;
  const long long* t;       // [B] chain indices
  const float* x;           // [B, F, C, HW]
  const void* model_out;    // [B, F, 2C, HW]
  const float* noise;       // [B, F, C, HW] or NULL
  float* x_prev;            // outputs, each [B, F, C, HW] or NULL
  float* pred_xstart;
  float* mean;
  float* log_variance;
  long long per_sample;     // F * C * HW
  long long total4;         // B * per_sample / 4
  int C, HW;
  int method;               // B200_SAMPLER_DDPM / B200_SAMPLER_DDIM
  int clip;
};

template <int DT>
__device__ __forceinline__ float4 load4(const void* base, long long idx) {
  if constexpr (DT == 0) {
    return *reinterpret_cast<const float4*>(static_cast<const float*>(base) + idx);
  } else {
    const uint2 raw = *reinterpret_cast<const uint2*>(static_cast<const uint16_t*>(base) + idx);
    float2 a, b;
    if constexpr (DT == 1) {
      a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
      b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
    } else {
      a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.x));
      b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&raw.y));
    }
    return make_float4(a.x, a.y, b.x, b.y);
  }
}

struct StepCoef {
  float A, Bc, c1, c2, min_log, max_log, sqrt_abar_prev, sigma, dir, nz_sigma, nonzero;
};

template <int DT>
__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerDev p) {
  const long long i4 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i4 >= p.total4) return;
  const long long e = i4 * 4;
  const int b = static_cast<int>(e / p.per_sample);
  const long long rem = e - static_cast<long long>(b) * p.per_sample;
  const long long chw = static_cast<long long>(p.C) * p.HW;
  const long long f = rem / chw;
  const long long cp = rem - f * chw;                                     // c * HW + pixel
  const long long mo = (static_cast<long long>(b) * (p.per_sample / chw) + f) * 2 * chw + cp;   // eps; var values at + chw

  const long long t = p.t[b];
  StepCoef k;
  k.A = p.tab.sqrt_recip_alphas_cumprod[t];
  k.Bc = p.tab.sqrt_recipm1_alphas_cumprod[t];
  k.c1 = p.tab.posterior_mean_coef1[t];
  k.c2 = p.tab.posterior_mean_coef2[t];
  k.min_log = p.tab.posterior_log_variance_clipped[t];
  k.max_log = p.tab.log_betas[t];
  k.nonzero = t != 0 ? 1.0f : 0.0f;
  if (p.method == B200_SAMPLER_DDIM) {
    k.sigma = p.tab.ddim_sigma[t];
    k.sqrt_abar_prev = p.tab.ddim_sqrt_alpha_prev[t];
    k.dir = p.tab.ddim_dir[t];
    k.nz_sigma = __fmul_rn(k.nonzero, k.sigma);
  }

  const float4 x4 = *reinterpret_cast<const float4*>(p.x + e);
  const float4 e4 = load4<DT>(p.model_out, mo);
  float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f), n4 = make_float4(0.f, 0.f, 0.f, 0.f);
  const bool need_var = p.method == B200_SAMPLER_DDPM || p.log_variance != nullptr;
  if (need_var) v4 = load4<DT>(p.model_out, mo + chw);
  if (p.noise) n4 = *reinterpret_cast<const float4*>(p.noise + e);

  const float xs[4] = {x4.x, x4.y, x4.z, x4.w}, es[4] = {e4.x, e4.y, e4.z, e4.w}, vs[4] = {v4.x, v4.y, v4.z, v4.w},
              ns[4] = {n4.x, n4.y, n4.z, n4.w};
  float out[4], x0[4], mu[4], lv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float ax = __fmul_rn(k.A, xs[j]);
    float pred = __fsub_rn(ax, __fmul_rn(k.Bc, es[j]));                   // _predict_xstart_from_eps (:338-343)
    if (p.clip) pred = fminf(fmaxf(pred, -1.0f), 1.0f);
    x0[j] = pred;
    mu[j] = __fadd_rn(__fmul_rn(k.c1, pred), __fmul_rn(k.c2, xs[j]));     // q_posterior_mean_variance (:236-239)
    if (need_var) {
      const float frac = __fdiv_rn(__fadd_rn(vs[j], 1.0f), 2.0f);         // (:300-302)
      lv[j] = __fadd_rn(__fmul_rn(frac, k.max_log), __fmul_rn(__fsub_rn(1.0f, frac), k.min_log));
    } else {
      lv[j] = 0.f;
    }
    if (p.method == B200_SAMPLER_DDIM) {
      const float eps = __fdiv_rn(__fsub_rn(ax, pred), k.Bc);             // _predict_eps_from_xstart (:345-348)
      const float mean_pred = __fadd_rn(__fmul_rn(pred, k.sqrt_abar_prev), __fmul_rn(k.dir, eps));
      out[j] = __fadd_rn(mean_pred, __fmul_rn(k.nz_sigma, ns[j]));
    } else {
      const float sd = expf(__fmul_rn(0.5f, lv[j]));
      out[j] = __fadd_rn(mu[j], __fmul_rn(__fmul_rn(k.nonzero, sd), ns[j]));
    }
  }
  if (p.x_prev) *reinterpret_cast<float4*>(p.x_prev + e) = make_float4(out[0], out[1], out[2], out[3]);
  if (p.pred_xstart) *reinterpret_cast<float4*>(p.pred_xstart + e) = make_float4(x0[0], x0[1], x0[2], x0[3]);
  if (p.mean) *reinterpret_cast<float4*>(p.mean + e) = make_float4(mu[0], mu[1], mu[2], mu[3]);
  if (p.log_variance) *reinterpret_cast<float4*>(p.log_variance + e) = make_float4(lv[0], lv[1], lv[2], lv[3]);
}

}  // namespace

// ------------------------------------------------------------------------------------------------ training objective
// GaussianDiffusion.training_losses for LossType.MSE + LEARNED_RANGE (gaussian_diffusion.py:719-795; vb term :686-716,
// diffusion_utils.py:10-88) as ONE pass over (x_0, x_t, noise, model output): per-sample sums of the squared error and of the
// variational-bound term (KL to the true posterior in bits, discretised decoder NLL at t == 0), AND the gradient of both with
// respect to the model output -- eps channels get d mse, the variance channels get d vb (the mean is frozen there, :757-765).
// The reference issues ~80 elementwise launches for this (forward + autograd) on 65 k elements per video.
struct LossDev {
  const float *x0, *xt, *noise, *mo;
  const long long* t;
  const float *recip, *recipm1, *coef1, *coef2, *min_log, *max_log;   // schedule tables (fp32 copies of the float64 arrays)
  float* dmo;                                                          // same layout as mo, or nullptr
  float* sums;                                                         // [2][B]: sum of squared error, sum of vb (nats)
  long long per_sample;                                                // F * C * HW
  int C, HW, batch;
};

__device__ __forceinline__ float approx_cdf(float z, float* dz) {      // diffusion_utils.py:40-45 and its derivative
  const float k = 0.7978845608028654f;
  const float th = tanhf(k * (z + 0.044715f * z * z * z));
  *dz = 0.5f * (1.0f - th * th) * k * (1.0f + 3.0f * 0.044715f * z * z);
  return 0.5f * (1.0f + th);
}

__global__ void __launch_bounds__(256) training_loss_kernel(const LossDev p) {
  const int b = blockIdx.y;
  const long long chw = static_cast<long long>(p.C) * p.HW;
  const long long t = p.t[b];
  const float A = p.recip[t], Bc = p.recipm1[t], c1 = p.coef1[t], c2 = p.coef2[t], mn = p.min_log[t], mx = p.max_log[t];
  const float inv_n = 1.0f / static_cast<float>(p.per_sample);
  float s_mse = 0.f, s_vb = 0.f;
  for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < p.per_sample;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long f = i / chw, cp = i - f * chw;
    const long long xi = static_cast<long long>(b) * p.per_sample + i;
    const long long mi = (static_cast<long long>(b) * (p.per_sample / chw) + f) * 2 * chw + cp;   // eps; var values at + chw
    const float x0 = p.x0[xi], xt = p.xt[xi], nz = p.noise[xi], eps = p.mo[mi], v = p.mo[mi + chw];
    const float d = nz - eps;
    s_mse += d * d;
    const float frac = 0.5f * (v + 1.0f);
    const float lv = frac * mx + (1.0f - frac) * mn;
    const float mean = c1 * (A * xt - Bc * eps) + c2 * xt;
    float term, dterm_dlv;
    if (t != 0) {
      const float tm = c1 * x0 + c2 * xt;
      const float e1 = __expf(mn - lv), dm = tm - mean, e2 = dm * dm * __expf(-lv);
      term = 0.5f * (-1.0f + lv - mn + e1 + e2);
      dterm_dlv = 0.5f * (1.0f - e1 - e2);
    } else {
      const float cen = x0 - mean, inv_std = __expf(-0.5f * lv);
      const float za = inv_std * (cen + 1.0f / 255.0f), zb = inv_std * (cen - 1.0f / 255.0f);
      float da, db;
      const float ca = approx_cdf(za, &da), cb = approx_cdf(zb, &db);
      float arg, darg;                       // log_probs = log(max(arg, 1e-12)); z scales with inv_std: dz/dlv = -z/2
      if (x0 < -0.999f) { arg = ca; darg = da * (-0.5f * za); }
      else if (x0 > 0.999f) { arg = 1.0f - cb; darg = -db * (-0.5f * zb); }
      else { arg = ca - cb; darg = da * (-0.5f * za) - db * (-0.5f * zb); }
      const bool live = arg > 1e-12f;
      term = -__logf(live ? arg : 1e-12f);
      dterm_dlv = live ? -darg / arg : 0.f;
    }
    s_vb += term;
    if (p.dmo != nullptr) {
      p.dmo[mi] = -2.0f * d * inv_n;
      p.dmo[mi + chw] = dterm_dlv * 0.5f * (mx - mn) * inv_n * 1.4426950408889634f;   // / ln 2: bits
    }
  }
  __shared__ float red[2][8];
  for (int o = 16; o > 0; o >>= 1) {
    s_mse += __shfl_xor_sync(0xffffffffu, s_mse, o);
    s_vb += __shfl_xor_sync(0xffffffffu, s_vb, o);
  }
  if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s_mse; red[1][threadIdx.x >> 5] = s_vb; }
  __syncthreads();
  if (threadIdx.x < 2) {
    float tsum = 0.f;
    for (int w = 0; w < 8; ++w) tsum += red[threadIdx.x][w];
    atomicAdd(p.sums + threadIdx.x * p.batch + b, tsum);
  }
}

int launch_sampler_step(const B200SamplerTables* tab, int method, int clip_denoised, const long long* t,
                        const float* x, const void* model_out, int model_out_dtype, const float* noise, int batch,
                        int frames, int channels, int hw, float* x_prev, float* pred_xstart, float* mean,
                        float* log_variance, cudaStream_t stream) {
  B200_REQUIRE(tab && t && x && model_out, B200_ERR_SHAPE, "sampler_step: NULL argument");
  B200_REQUIRE(batch > 0 && frames > 0 && channels > 0 && hw > 0 && hw % 4 == 0, B200_ERR_SHAPE,
               "sampler_step: bad shape (B %d, F %d, C %d, H*W %d; H*W must be a multiple of 4)", batch, frames, channels, hw);
  B200_REQUIRE(method == B200_SAMPLER_DDPM || method == B200_SAMPLER_DDIM, B200_ERR_UNSUPPORTED, "sampler_step: unknown method %d", method);
  B200_REQUIRE(model_out_dtype >= 0 && model_out_dtype <= 2, B200_ERR_DTYPE, "sampler_step: model_out dtype %d (0 fp32, 1 fp16, 2 bf16)", model_out_dtype);
  B200_REQUIRE(noise || method == B200_SAMPLER_DDIM || !x_prev, B200_ERR_SHAPE, "sampler_step: DDPM sampling needs noise");
  B200_REQUIRE(tab->sqrt_recip_alphas_cumprod && tab->sqrt_recipm1_alphas_cumprod && tab->posterior_mean_coef1 &&
                   tab->posterior_mean_coef2 && tab->posterior_log_variance_clipped && tab->log_betas,
               B200_ERR_SHAPE, "sampler_step: a schedule table is NULL");
  B200_REQUIRE(method != B200_SAMPLER_DDIM || (tab->ddim_sqrt_alpha_prev && tab->ddim_sigma && tab->ddim_dir), B200_ERR_SHAPE,
               "sampler_step: DDIM needs ddim_sqrt_alpha_prev, ddim_sigma, ddim_dir");
  const uintptr_t al = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(noise) | reinterpret_cast<uintptr_t>(x_prev) |
                       reinterpret_cast<uintptr_t>(pred_xstart) | reinterpret_cast<uintptr_t>(mean) | reinterpret_cast<uintptr_t>(log_variance);
  B200_REQUIRE((al & 15) == 0 && (reinterpret_cast<uintptr_t>(model_out) & (model_out_dtype == 0 ? 15 : 7)) == 0, B200_ERR_ALIGN,
               "sampler_step: tensors must be 16-byte aligned");
  B200_TRY(check_arch());
  SamplerDev p;
  p.tab = *tab;
  p.t = t; p.x = x; p.model_out = model_out; p.noise = noise;
  p.x_prev = x_prev; p.pred_xstart = pred_xstart; p.mean = mean; p.log_variance = log_variance;
  p.per_sample = static_cast<long long>(frames) * channels * hw;
  p.total4 = static_cast<long long>(batch) * p.per_sample / 4;
  p.C = channels; p.HW = hw; p.method = method; p.clip = clip_denoised;
  const unsigned blocks = static_cast<unsigned>((p.total4 + 255) / 256);
  switch (model_out_dtype) {
    case 0: sampler_step_kernel<0><<<blocks, 256, 0, stream>>>(p); break;
    case 1: sampler_step_kernel<1><<<blocks, 256, 0, stream>>>(p); break;
    default: sampler_step_kernel<2><<<blocks, 256, 0, stream>>>(p); break;
  }
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

int launch_training_loss(const B200SamplerTables* tab, const long long* t, const float* x0, const float* xt, const float* noise,
                         const float* model_out, int batch, int frames, int channels, int hw, float* sums, float* dmo, cudaStream_t stream) {
  B200_REQUIRE(tab && t && x0 && xt && noise && model_out && sums, B200_ERR_SHAPE, "training_loss: NULL argument");
  B200_REQUIRE(batch > 0 && frames > 0 && channels > 0 && hw > 0, B200_ERR_SHAPE, "training_loss: bad shape");
  B200_REQUIRE(tab->sqrt_recip_alphas_cumprod && tab->sqrt_recipm1_alphas_cumprod && tab->posterior_mean_coef1 &&
                   tab->posterior_mean_coef2 && tab->posterior_log_variance_clipped && tab->log_betas,
               B200_ERR_SHAPE, "training_loss: a schedule table is NULL");
  LossDev p;
  p.x0 = x0; p.xt = xt; p.noise = noise; p.mo = model_out; p.t = t;
  p.recip = tab->sqrt_recip_alphas_cumprod; p.recipm1 = tab->sqrt_recipm1_alphas_cumprod; p.coef1 = tab->posterior_mean_coef1;
  p.coef2 = tab->posterior_mean_coef2; p.min_log = tab->posterior_log_variance_clipped; p.max_log = tab->log_betas;
  p.dmo = dmo; p.sums = sums;
  p.per_sample = static_cast<long long>(frames) * channels * hw;
  p.C = channels; p.HW = hw; p.batch = batch;
  B200_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * batch, stream));
  long long bx = (p.per_sample + 255) / 256;
  if (bx > 64) bx = 64;
  training_loss_kernel<<<dim3(static_cast<unsigned>(bx), batch), 256, 0, stream>>>(p);
  B200_CHECK_CUDA(cudaGetLastError());
  return B200_OK;
}

}  // namespace b200
