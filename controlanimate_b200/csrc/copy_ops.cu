// The two pure data-movement steps of the UNet's up path on channels-last rows, as 16-byte streaming kernels:
//
//   ca_upsample_nearest : F.interpolate(mode="nearest") of Upsample3D.forward (animatediff/models/resnet.py:63-69), by a
//                         factor of 2 or to an explicit output size (forward_upsample_size, unet.py:491-499, 596-597)
//   ca_concat_channels  : torch.cat([hidden_states, res_hidden_states], dim=1) in front of every up-block resnet
//                         (animatediff/models/unet_blocks.py:636, :742)
//
// torch's channels-last nearest-upsample kernel moves these tensors at ~0.8 TB/s (263 us per launch in the config-2
// step, profiles/r02_step_launches.md) and the strided concat at ~60 % of the copy roofline; both are HBM-bound copies:
// (in + out) * s bytes, four independent vectors per thread in flight.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kThreads = 256;
constexpr int kUnroll = 4;

// y[n, oh, ow, :] = x[n, src(oh), src(ow), :], src(d) = min(floor(d * scale), in - 1): torch's
// nearest_neighbor_compute_source_index with scale = 1 / scale_factor (0.5 here) or in / out for an explicit size
__global__ void __launch_bounds__(kThreads) upsample_nearest_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, long long nvec,
                                                                    int cvec, int in_h, int in_w, int out_h, int out_w,
                                                                    float scale_h, float scale_w) {
  const long long v0 = (long long)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  uint4 val[kUnroll];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const long long i = v0 + (long long)u * kThreads;
    ok[u] = i < nvec;
    if (ok[u]) {
      const int cv = (int)(i % cvec);
      long long r = i / cvec;
      const int ow = (int)(r % out_w);
      r /= out_w;
      const int oh = (int)(r % out_h);
      const long long n = r / out_h;
      int ih = (int)floorf((float)oh * scale_h), iw = (int)floorf((float)ow * scale_w);
      ih = ih < in_h - 1 ? ih : in_h - 1;
      iw = iw < in_w - 1 ? iw : in_w - 1;
      val[u] = ldg_keep(x + ((n * in_h + ih) * in_w + iw) * cvec + cv);  // every input vector is read four times: keep it cached
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
    if (ok[u]) y[v0 + (long long)u * kThreads] = val[u];
}

// y[r, :] = [a[r, :], b[r, :]]
__global__ void __launch_bounds__(kThreads) concat_channels_kernel(const uint4* __restrict__ a, const uint4* __restrict__ b,
                                                                   uint4* __restrict__ y, long long nvec, int avec, int bvec) {
  const long long v0 = (long long)blockIdx.x * (kThreads * kUnroll) + threadIdx.x;
  const int cvec = avec + bvec;
  uint4 val[kUnroll];
  bool ok[kUnroll];
#pragma unroll
  for (int u = 0; u < kUnroll; ++u) {
    const long long i = v0 + (long long)u * kThreads;
    ok[u] = i < nvec;
    if (ok[u]) {
      const long long r = i / cvec;
      const int j = (int)(i - r * cvec);
      val[u] = j < avec ? ldg_stream(a + r * avec + j) : ldg_stream(b + r * bvec + (j - avec));
    }
  }
#pragma unroll
  for (int u = 0; u < kUnroll; ++u)
    if (ok[u]) y[v0 + (long long)u * kThreads] = val[u];
}

inline unsigned blocks_for(long long nvec) { return (unsigned)((nvec + kThreads * kUnroll - 1) / (kThreads * kUnroll)); }

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_upsample_nearest(const void* x, void* y, long long n, int c, int in_h,
                                                                          int in_w, int out_h, int out_w, int exact_2x,
                                                                          int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y, "upsample_nearest: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "upsample_nearest: bad dtype");
  const int vec = dtype == CA_F32 ? 4 : 8;
  CA_CHECK_ARG(n >= 0 && c > 0 && c % vec == 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "upsample_nearest: bad sizes (c %% %d)", vec);
  CA_CHECK_ARG(!exact_2x || (out_h == 2 * in_h && out_w == 2 * in_w), "upsample_nearest: exact_2x needs out = 2 * in");
  CA_CHECK_ARG(aligned16(x) && aligned16(y), "upsample_nearest: pointers must be 16-byte aligned");
  const long long nvec = n * out_h * out_w * (c / vec);
  if (nvec == 0) return CA_OK;
  CA_CHECK_ARG(blocks_for(nvec) < (1u << 31), "upsample_nearest: tensor too large");
  // scale_factor = 2 -> torch uses 1 / scale_factor; explicit size -> in / out (float, as torch computes it)
  const float sh = exact_2x ? 0.5f : (float)in_h / (float)out_h, sw = exact_2x ? 0.5f : (float)in_w / (float)out_w;
  upsample_nearest_kernel<<<blocks_for(nvec), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(x), reinterpret_cast<uint4*>(y), nvec, c / vec, in_h, in_w, out_h, out_w, sh, sw);
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}

extern "C" __attribute__((visibility("default"))) int ca_concat_channels(const void* a, const void* b, void* y, long long rows,
                                                                         int ca_, int cb_, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(a && b && y, "concat_channels: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "concat_channels: bad dtype");
  const int vec = dtype == CA_F32 ? 4 : 8;
  CA_CHECK_ARG(rows >= 0 && ca_ > 0 && cb_ > 0 && ca_ % vec == 0 && cb_ % vec == 0, "concat_channels: channel counts must be multiples of %d", vec);
  CA_CHECK_ARG(aligned16(a) && aligned16(b) && aligned16(y), "concat_channels: pointers must be 16-byte aligned");
  const long long nvec = rows * ((ca_ + cb_) / vec);
  if (nvec == 0) return CA_OK;
  CA_CHECK_ARG(blocks_for(nvec) < (1u << 31), "concat_channels: tensor too large");
  concat_channels_kernel<<<blocks_for(nvec), kThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const uint4*>(a), reinterpret_cast<const uint4*>(b), reinterpret_cast<uint4*>(y), nvec, ca_ / vec, cb_ / vec);
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}
