// Loop glue of one denoising step as ONE launch: classifier-free-guidance combine + DDIM update.
//
//   eps      = eps_u + g * (eps_c - eps_u)                     (animatediff/pipelines/controlanimation_pipeline.py:845-846)
//   x0       = (x_t - sqrt(1 - a_t) * eps) / sqrt(a_t)         (diffusers 0.23.0 DDIMScheduler.step, eta = 0, epsilon
//   x_{t-1}  = sqrt(a_prev) * x0 + sqrt(1 - a_prev) * eps       prediction, clip_sample False; pipeline :849)
//
// The reference runs these as ~9 elementwise torch kernels on the [1, 4, f, h, w] latents (the `.to(latents_dtype)` of the
// UNet output, chunk, sub, mul, add, and the scheduler's four).  Here the UNet output (model dtype, rows [uncond | cond] when
// CFG is on) and the latents (their own dtype, typically fp32) are read once, everything is fp32 in registers, and the new
// latents are rounded once.  HBM-bound on (b * s_model + 2 * s_latent) bytes per element — 1-3 MB per step, a launch-latency
// sized kernel: what it buys is eight launches less per step, not bandwidth.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kThreads = 256;

template <typename TM, typename TL>
__global__ void __launch_bounds__(kThreads) cfg_ddim_kernel(const TM* __restrict__ model_out, const TL* latents, TL* latents_out,
                                                            TL* __restrict__ noise_out, long long n,
                                                            int cfg, float guidance, float sqrt_a, float sqrt_1ma,
                                                            float sqrt_ap, float sqrt_1map) {
  const long long i = (long long)blockIdx.x * kThreads + threadIdx.x;
  if (i >= n) return;
  // the reference converts the UNet output to the latents' dtype BEFORE the guidance arithmetic (:841): same here, so a
  // 16-bit latent path sees the same operands (the arithmetic itself stays fp32)
  float eps = Traits<TL>::to_f(Traits<TL>::from_f(Traits<TM>::to_f(model_out[i])));
  if (cfg) {
    const float c = Traits<TL>::to_f(Traits<TL>::from_f(Traits<TM>::to_f(model_out[n + i])));
    eps = eps + guidance * (c - eps);
  }
  const float x = Traits<TL>::to_f(latents[i]);  // latents_out may alias latents (no __restrict__, no .nc load)
  const float x0 = (x - sqrt_1ma * eps) / sqrt_a;
  latents_out[i] = Traits<TL>::from_f(sqrt_ap * x0 + sqrt_1map * eps);
  if (noise_out) noise_out[i] = Traits<TL>::from_f(eps);
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_cfg_ddim_step(const void* model_out, const void* latents,
                                                                       void* latents_out, void* noise_out, long long n, int cfg,
                                                                       float guidance, float sqrt_alpha_t,
                                                                       float sqrt_one_minus_alpha_t, float sqrt_alpha_prev,
                                                                       float sqrt_one_minus_alpha_prev, int model_dtype,
                                                                       int latent_dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(model_out && latents && latents_out, "cfg_ddim_step: null pointer");
  CA_CHECK_ARG(n >= 0 && (cfg == 0 || cfg == 1), "cfg_ddim_step: bad n / cfg");
  CA_CHECK_ARG(sqrt_alpha_t > 0.f, "cfg_ddim_step: sqrt(alpha_t) must be positive");
  // only latents_out may alias an input (latents: same element, read before written); model_out is read through the
  // read-only path and noise_out is a second output
  CA_CHECK_ARG(model_out != latents_out && model_out != noise_out && noise_out != latents && noise_out != latents_out,
               "cfg_ddim_step: model_out / noise_out must not alias the other buffers (latents_out may alias latents)");
  if (n == 0) return CA_OK;
  const long long blocks = (n + kThreads - 1) / kThreads;
  CA_CHECK_ARG(blocks < (1ll << 31), "cfg_ddim_step: tensor too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch_dtype(model_dtype, [&](auto mtag) -> int {
    using TM = decltype(mtag);
    return dispatch_dtype(latent_dtype, [&](auto ltag) -> int {
      using TL = decltype(ltag);
      cfg_ddim_kernel<TM, TL><<<(unsigned)blocks, kThreads, 0, st>>>(
          reinterpret_cast<const TM*>(model_out), reinterpret_cast<const TL*>(latents), reinterpret_cast<TL*>(latents_out),
          reinterpret_cast<TL*>(noise_out), n, cfg, guidance, sqrt_alpha_t, sqrt_one_minus_alpha_t, sqrt_alpha_prev,
          sqrt_one_minus_alpha_prev);
      CA_CUDA(cudaGetLastError());
      return CA_OK;
    });
  });
}
