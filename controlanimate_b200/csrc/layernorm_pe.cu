// LayerNorm (+ temporal positional encoding) on token-major rows: one warp per token, the row
// lives in registers between the statistics and the normalise step -> one HBM read, one write.
//
// Replaces nn.LayerNorm at reference animatediff/models/motion_module.py:214,221 and the
// PositionalEncoding add of VersatileAttention.forward (:285-288).  The reference first permutes
// '(b f) d c -> (b d) f c' so that pe[:, :f] broadcasts along dim 1; here the frame index of token
// t = (b*f + frame)*d + site is recovered arithmetically and pe[frame] is added in place, so the
// permuted copy never exists.
//
// Design: persistent warps (grid = SMs x resident CTAs) walk the rows with a grid stride; the
// loads of the NEXT row are issued before the current row is reduced (register double buffer), so
// every warp keeps two rows of 16-byte loads in flight; gamma/beta sit in shared memory; NV (16-byte
// vectors per lane) is a template parameter so small channel counts keep the register file free for
// occupancy (c = 320 -> NV 2, 640 -> 3, 1280 -> 5).
#include "common.cuh"

namespace ca {
namespace {

constexpr int kWarpsPerCta = 8;

// LPR lanes cooperate on one row (32/LPR rows per warp per iteration); lane sl of a row owns 16-byte vectors
// sl, sl+LPR, ... (NV of them).  c = 320 -> LPR 8, NV 5; 640 -> 16, 5; 1280 -> 32, 5: every lane is busy.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    layernorm_pe_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ pe, long long rows, int c, int f,
                        int d, float eps) {
  constexpr int VEC = Traits<T>::kVec;
  constexpr int R = 32 / LPR;  // rows per warp per iteration
  extern __shared__ __align__(16) float s_gb[];  // gamma[c], beta[c]
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    s_gb[i] = gamma[i];
    s_gb[c + i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nvec = c / VEC;
  const long long row0 = ((long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * R + sub;
  const long long stride = (long long)gridDim.x * kWarpsPerCta * R;
  const float inv_c = 1.0f / (float)c;

  uint4 nxt[NV];
  auto load_row = [&](long long row) {
    const T* xr = x + row * c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = sl + i * LPR;
      if (vi < nvec) nxt[i] = ldg_stream(xr + vi * VEC);
    }
  };
  if (row0 < rows) load_row(row0);
  // all lanes of a warp run the same number of iterations (shuffles need the full warp)
  const long long warp_first = row0 - sub;
  for (long long base = warp_first; base < rows; base += stride) {
    const long long row = base + sub;
    const bool live = row < rows;
    float v[NV][VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (live && sl + i * LPR < nvec) {
        Vec16<T> r;
        r.raw = nxt[i];
        r.unpack(v[i]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) sum += v[i][j];
      }
    }
    if (row + stride < rows) load_row(row + stride);  // prefetch the next row while this one is reduced
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (live && sl + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float dlt = v[i][j] - mean;
          sq += dlt * dlt;
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    const float nmr = -mean * rstd;
    if (!live) continue;
    const float* per = pe ? pe + (long long)((row / d) % f) * c : nullptr;
    T* yr = y + row * c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = sl + i * LPR;
      if (vi < nvec) {
        float o[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(s_gb + vi * VEC + j);
          const float4 b4 = *reinterpret_cast<const float4*>(s_gb + c + vi * VEC + j);
          o[j + 0] = fmaf(fmaf(v[i][j + 0], rstd, nmr), g4.x, b4.x);
          o[j + 1] = fmaf(fmaf(v[i][j + 1], rstd, nmr), g4.y, b4.y);
          o[j + 2] = fmaf(fmaf(v[i][j + 2], rstd, nmr), g4.z, b4.z);
          o[j + 3] = fmaf(fmaf(v[i][j + 3], rstd, nmr), g4.w, b4.w);
          if (per) {
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(per + vi * VEC + j));
            o[j + 0] += p4.x;
            o[j + 1] += p4.y;
            o[j + 2] += p4.z;
            o[j + 3] += p4.w;
          }
        }
        Vec16<T> r;
        r.pack(o);
        stg_stream(yr + vi * VEC, r.raw);
      }
    }
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_layernorm_pe(const void* x, void* y, const float* gamma,
                                                                      const float* beta, const float* pe, long long rows,
                                                                      int c, int f, int d, float eps, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && gamma && beta, "layernorm_pe: null pointer");
  CA_CHECK_ARG(rows >= 0 && c > 0 && f > 0 && d > 0, "layernorm_pe: bad sizes");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "layernorm_pe: bad dtype");
  const int vec = dtype == CA_F32 ? 4 : 8;
  const int nvec = c / vec;
  CA_CHECK_ARG(c % vec == 0 && nvec <= 32 * 16, "layernorm_pe: c=%d unsupported (multiple of %d, <= %d)", c, vec, 512 * vec);
  CA_CHECK_ARG(aligned16(x) && aligned16(y) && (!pe || aligned16(pe)), "layernorm_pe: x/y/pe must be 16-byte aligned");
  if (rows == 0) return CA_OK;
  // lanes per row: smallest of 8/16/32 that needs at most 5 vectors per lane (8 for very wide rows)
  int lpr = 8;
  while (lpr < 32 && (nvec + lpr - 1) / lpr > 5) lpr <<= 1;
  const int nv = (nvec + lpr - 1) / lpr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const size_t smem = 2 * (size_t)c * sizeof(float);
  const int rc = dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    auto run = [&](auto kernel) -> int {
      const void* fn = reinterpret_cast<const void*>(kernel);
      int per_sm = 1;
      CA_CUDA(cached_occupancy(&per_sm, fn, kWarpsPerCta * 32, smem));
      const int rows_per_cta = kWarpsPerCta * (32 / lpr);
      long long grid = (long long)sm_count() * (per_sm < 1 ? 1 : per_sm);
      const long long need = (rows + rows_per_cta - 1) / rows_per_cta;
      if (grid > need) grid = need;
      kernel<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), gamma,
                                                              beta, pe, rows, c, f, d, eps);
      return CA_OK;
    };
#define CA_LN_CASE(L_, N_) if (lpr == L_ && nv <= N_) return run(layernorm_pe_kernel<T, L_, N_>)
    CA_LN_CASE(8, 1); CA_LN_CASE(8, 2); CA_LN_CASE(8, 3); CA_LN_CASE(8, 5);
    CA_LN_CASE(16, 3); CA_LN_CASE(16, 5);
    CA_LN_CASE(32, 3); CA_LN_CASE(32, 5); CA_LN_CASE(32, 8); CA_LN_CASE(32, 16);
#undef CA_LN_CASE
    set_error("layernorm_pe: no kernel for c=%d", c);
    return CA_ERR_UNSUPPORTED;
  });
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}
