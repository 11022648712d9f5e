// Convolution epilogue for the 3D ResNet blocks on channels-last rows:
//
//   y[r, :] = (act(x[r, :] + bias[:]) + residual[r, :]) * scale
//
// Replaces the separate passes the reference (through torch/cuDNN) spends after every convolution of the hot path:
// the broadcast bias add behind each InflatedConv3d (animatediff/models/resnet.py:12-20 -> nn.Conv2d), the
// shortcut add + output scale of ResnetBlock3D.forward (resnet.py:213-216), and conv + SiLU in the ControlNet
// conditioning embedding (diffusers ControlNetConditioningEmbedding, called from
// modules/controlresiduals_pipeline.py:294-302).  Pure HBM streaming: (2 or 3) * N * s bytes, 16-byte accesses,
// four independent vectors per thread in flight, the bias vector stays in L1.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

template <typename T, int ACT, bool RES>
__global__ void __launch_bounds__(kThreads) bias_act_residual_kernel(const T* x, const float* __restrict__ bias, const T* res, T* y,
                                                                     long long nvec, int cvec, float scale) {
  // x / res / y carry no __restrict__ and are read without .nc: the epilogue runs in place (y == x or y == res); every
  // thread reads its own vectors before it writes them, which is all the aliasing this kernel allows
  constexpr int VEC = Traits<T>::kVec;
  const long long v0 = (long long)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  uint4 xv[kUnroll], rv[kUnroll];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const long long i = v0 + (long long)u * kThreads;
    ok[u] = i < nvec;
    if (ok[u]) {
      xv[u] = ldg_stream_rw(x + i * VEC);
      if constexpr (RES) rv[u] = ldg_stream_rw(res + i * VEC);
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    if (!ok[u]) continue;
    const long long i = v0 + (long long)u * kThreads;
    const int cv = (int)(i % cvec);
    float f[VEC], r[VEC];
    Vec16<T> v;
    v.raw = xv[u];
    v.unpack(f);
    if (bias) {
      const float4* b4 = reinterpret_cast<const float4*>(bias + (long long)cv * VEC);
#pragma unroll
      for (int q = 0; q < VEC / 4; ++q) {
        const float4 b = __ldg(b4 + q);
        f[4 * q] += b.x;
        f[4 * q + 1] += b.y;
        f[4 * q + 2] += b.z;
        f[4 * q + 3] += b.w;
      }
    }
    if constexpr (ACT == 1) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) f[e] = silu_f(f[e]);
    }
    if constexpr (RES) {
      Vec16<T> w;
      w.raw = rv[u];
      w.unpack(r);
#pragma unroll
      for (int e = 0; e < VEC; ++e) f[e] += r[e];
    }
#pragma unroll
    for (int e = 0; e < VEC; ++e) f[e] *= scale;
    v.pack(f);
    stg_stream(y + i * VEC, v.raw);
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_bias_act_residual(const void* x, const float* bias,
                                                                           const void* residual, void* y, long long rows,
                                                                           int c, float scale, int act, int dtype,
                                                                           void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y, "bias_act_residual: null pointer");
  CA_CHECK_ARG(rows >= 0 && c > 0, "bias_act_residual: bad shape");
  CA_CHECK_ARG(act == 0 || act == 1, "bias_act_residual: act must be 0 (none) or 1 (SiLU)");
  if (rows == 0) return CA_OK;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  return dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    constexpr int VEC = Traits<T>::kVec;
    CA_CHECK_ARG(c % VEC == 0, "bias_act_residual: c=%d must be a multiple of %d", c, VEC);
    CA_CHECK_ARG(aligned16(x) && aligned16(y) && (!residual || aligned16(residual)) && (!bias || aligned16(bias)),
                 "bias_act_residual: pointers must be 16-byte aligned");
    const long long nvec = rows * (c / VEC);
    const long long blocks = (nvec + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    CA_CHECK_ARG(blocks < (1ll << 31), "bias_act_residual: tensor too large");
    const T* xp = reinterpret_cast<const T*>(x);
    const T* rp = reinterpret_cast<const T*>(residual);
    T* yp = reinterpret_cast<T*>(y);
    auto go = [&](auto kern) {
      kern<<<(unsigned)blocks, kThreads, 0, st>>>(xp, bias, rp, yp, nvec, c / VEC, scale);
    };
    if (residual) {
      if (act) go(bias_act_residual_kernel<T, 1, true>);
      else go(bias_act_residual_kernel<T, 0, true>);
    } else {
      if (act) go(bias_act_residual_kernel<T, 1, false>);
      else go(bias_act_residual_kernel<T, 0, false>);
    }
    CA_CUDA(cudaGetLastError());
    return CA_OK;
  });
}
