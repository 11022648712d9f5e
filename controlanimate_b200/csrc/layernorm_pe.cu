// LayerNorm (+ temporal positional encoding) on token-major rows: one warp per token, the row
// lives in registers between the statistics and the normalise step -> one HBM read, one write.
//
// Replaces nn.LayerNorm at reference animatediff/models/motion_module.py:214,221 and the
// PositionalEncoding add of VersatileAttention.forward (:285-288).  The reference first permutes
// '(b f) d c -> (b d) f c' so that pe[:, :f] broadcasts along dim 1; here the frame index of token
// t = (b*f + frame)*d + site is recovered arithmetically and pe[frame] is added in place, so the
// permuted copy never exists.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr int kMaxElemsPerLane = 64;  // row slice kept in registers: supports c <= 2048

template <typename T>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    layernorm_pe_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ pe, long long rows, int c, int f,
                        int d, float eps) {
  constexpr int VEC = Traits<T>::kVec;
  constexpr int kMaxVecPerLane = kMaxElemsPerLane / VEC;
  const int lane = threadIdx.x & 31;
  const long long row = (long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int nvec = c / VEC;
  const T* xr = x + row * c;
  float v[kMaxVecPerLane][VEC];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      Vec16<T> r;
      r.raw = ldg_stream(xr + vi * VEC);
      r.unpack(v[i]);
#pragma unroll
      for (int j = 0; j < VEC; ++j) sum += v[i][j];
    }
  }
  const float mean = warp_sum(sum) / (float)c;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < kMaxVecPerLane; ++i) {
    if (lane + i * 32 < nvec) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float dlt = v[i][j] - mean;
        sq += dlt * dlt;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(sq) / (float)c + eps);
  const float* per = pe ? pe + (long long)((row / d) % f) * c : nullptr;
  T* yr = y + row * c;
#pragma unroll
  for (int i = 0; i < kMaxVecPerLane; ++i) {
    const int vi = lane + i * 32;
    if (vi < nvec) {
      float o[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const int ch = vi * VEC + j;
        o[j] = (v[i][j] - mean) * rstd * __ldg(gamma + ch) + __ldg(beta + ch);
        if (per) o[j] += __ldg(per + ch);
      }
      Vec16<T> r;
      r.pack(o);
      stg_stream(yr + vi * VEC, r.raw);
    }
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_layernorm_pe(const void* x, void* y, const float* gamma, const float* beta, const float* pe,
                               long long rows, int c, int f, int d, float eps, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && gamma && beta, "layernorm_pe: null pointer");
  CA_CHECK_ARG(rows >= 0 && c > 0 && f > 0 && d > 0, "layernorm_pe: bad sizes");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "layernorm_pe: bad dtype");
  const int vec = dtype == CA_F32 ? 4 : 8;
  CA_CHECK_ARG(c % vec == 0 && c <= 32 * kMaxElemsPerLane, "layernorm_pe: c=%d unsupported (multiple of %d, <= %d)", c,
               vec, 32 * kMaxElemsPerLane);
  CA_CHECK_ARG(aligned16(x) && aligned16(y), "layernorm_pe: x/y must be 16-byte aligned");
  if (rows == 0) return CA_OK;
  const long long grid = (rows + kWarpsPerCta - 1) / kWarpsPerCta;
  CA_CHECK_ARG(grid < (1ll << 31), "layernorm_pe: too many rows");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rc = dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    layernorm_pe_kernel<T><<<(unsigned)grid, kWarpsPerCta * 32, 0, st>>>(
        reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), gamma, beta, pe, rows, c, f, d, eps);
    return CA_OK;
  });
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}
