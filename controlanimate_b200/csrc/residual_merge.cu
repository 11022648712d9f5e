// Kernel (3): single-pass Multi-ControlNet residual merge.
//
//   dst_i (=|+=) sum_k scale[k][i] * res[k][i]     for the 13 residual tensors, ONE launch.
//
// Replaces: per-net `sample * conditioning_scale` and the `prev + curr` sum over nets (diffusers
// 0.23.0, called from reference modules/controlresiduals_pipeline.py:294-302), the 13 einops
// '(b f) c h w -> b c f h w' transposes (:304-312) and, when add_into_dst, the 13 skip additions
// of animatediff/models/unet.py:567-576, 584-585.  The reference moves 12*E*s bytes for two nets
// (SURVEY.md §8 A9); this kernel moves (N+2)*E*s in place or (N+1)*E*s as a producer.
//
// Pure HBM streaming: every thread issues all of its 16-byte loads (N nets x UNROLL vectors)
// before the first FMA, accumulates in fp32, writes once.  The c<->f transpose of the reference
// layout is free because it only permutes whole rows of h*w contiguous elements.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kThreads = 256;
constexpr int kVecPerThread = 4;  // 64 B of each operand in flight per thread

struct MergeTensor {
  const void* src[CA_MAX_NETS];
  float scale[CA_MAX_NETS];
  void* dst;
  long long nvec_dst;     // 16-byte vectors (or elements when VEC == 1) in dst
  long long src_batch_vecs;  // vectors of one batch item of src; src index = idx % (src_batch_vecs*b_res)
  int c, hwv;             // channels, (h*w)/VEC  (NCFHW transpose only)
  unsigned int block_begin;  // first CTA of this tensor
};

struct MergeParams {
  MergeTensor t[CA_MAX_RESIDUALS];
  int n_res, n_nets, f, b_res, add, transpose;
};

template <typename T, int VEC, int NETS>
__device__ __forceinline__ void merge_body(const MergeParams& p, const MergeTensor& t, long long v0) {
  // NETS > 0: compile-time net count (fully unrolled loads); NETS == 0: runtime p.n_nets.
  const int nets = NETS > 0 ? NETS : p.n_nets;
  using V = Vec16<T>;
  long long idx[kVecPerThread];
  long long sidx[kVecPerThread];
  bool ok[kVecPerThread];
#pragma unroll
  for (int u = 0; u < kVecPerThread; ++u) {
    idx[u] = v0 + (long long)u * kThreads;
    ok[u] = idx[u] < t.nvec_dst;
    long long i = ok[u] ? idx[u] : 0;
    if (p.transpose) {
      // dst vector i lives at [b, c, f, hwv]; source row is [(b f), c, hwv]
      const long long col = i % t.hwv;
      long long r = i / t.hwv;
      const int fi = (int)(r % p.f);
      r /= p.f;
      const int ci = (int)(r % t.c);
      const long long bi = (r / t.c) % p.b_res;  // batch broadcast
      sidx[u] = ((bi * p.f + fi) * t.c + ci) * t.hwv + col;
    } else {
      sidx[u] = i % (t.src_batch_vecs * p.b_res);
    }
  }
  if constexpr (VEC > 1) {
    uint4 raw[(NETS > 0 ? NETS : CA_MAX_NETS)][kVecPerThread];
    uint4 draw[kVecPerThread];
#pragma unroll
    for (int k = 0; k < (NETS > 0 ? NETS : CA_MAX_NETS); ++k)
      if (k < nets) {
#pragma unroll
        for (int u = 0; u < kVecPerThread; ++u)
          if (ok[u]) raw[k][u] = ldg_stream(reinterpret_cast<const uint4*>(t.src[k]) + sidx[u]);
      }
    if (p.add) {
#pragma unroll
      for (int u = 0; u < kVecPerThread; ++u)
        if (ok[u]) draw[u] = ldg_stream_rw(reinterpret_cast<const uint4*>(t.dst) + idx[u]);  // dst is written below: no .nc
    }
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u) {
      if (!ok[u]) continue;
      float acc[VEC];
      V v;
      if (p.add) {
        v.raw = draw[u];
        v.unpack(acc);
      } else {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
      }
#pragma unroll
      for (int k = 0; k < (NETS > 0 ? NETS : CA_MAX_NETS); ++k)
        if (k < nets) {
          float fv[VEC];
          v.raw = raw[k][u];
          v.unpack(fv);
#pragma unroll
          for (int j = 0; j < VEC; ++j) acc[j] = fmaf(fv[j], t.scale[k], acc[j]);
        }
      v.pack(acc);
      stg_stream(reinterpret_cast<uint4*>(t.dst) + idx[u], v.raw);
    }
  } else {
#pragma unroll
    for (int u = 0; u < kVecPerThread; ++u) {
      if (!ok[u]) continue;
      float acc = p.add ? Traits<T>::to_f(reinterpret_cast<const T*>(t.dst)[idx[u]]) : 0.f;
      for (int k = 0; k < nets; ++k)
        acc = fmaf(Traits<T>::to_f(reinterpret_cast<const T*>(t.src[k])[sidx[u]]), t.scale[k], acc);
      reinterpret_cast<T*>(t.dst)[idx[u]] = Traits<T>::from_f(acc);
    }
  }
}

template <typename T, int VEC, int NETS>
__global__ void __launch_bounds__(kThreads) residual_merge_kernel(const __grid_constant__ MergeParams p) {
  // locate the tensor this CTA works on (<= 16 entries, block_begin ascending)
  int ti = 0;
#pragma unroll 1
  for (int i = 1; i < p.n_res; ++i)
    if (blockIdx.x >= p.t[i].block_begin) ti = i;
  const MergeTensor& t = p.t[ti];
  const long long v0 = (long long)(blockIdx.x - t.block_begin) * (kThreads * kVecPerThread) + threadIdx.x;
  merge_body<T, VEC, NETS>(p, t, v0);
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_residual_merge(const void* const* res, const float* scales, void* const* dst, const int* chw,
                                 int n_nets, int n_res, int b_res, int b_dst, int f, int add_into_dst, int layout,
                                 int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(res && scales && dst && chw, "residual_merge: null pointer");
  CA_CHECK_ARG(n_nets >= 1 && n_nets <= CA_MAX_NETS, "residual_merge: n_nets=%d out of range [1,%d]", n_nets, CA_MAX_NETS);
  CA_CHECK_ARG(n_res >= 1 && n_res <= CA_MAX_RESIDUALS, "residual_merge: n_res=%d out of range", n_res);
  CA_CHECK_ARG(b_dst >= 1 && (b_res == b_dst || b_res == 1), "residual_merge: b_res=%d must be 1 or b_dst=%d", b_res, b_dst);
  CA_CHECK_ARG(f >= 1, "residual_merge: f=%d", f);
  CA_CHECK_ARG(layout == CA_LAYOUT_NCFHW || layout == CA_LAYOUT_BFHWC, "residual_merge: bad layout %d", layout);
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "residual_merge: bad dtype %d", dtype);
  const int esz = dtype == CA_F32 ? 4 : 2;
  const int vec16 = 16 / esz;
  // vector width: 16 bytes when every tensor allows it, else scalar
  int vec = vec16;
  for (int i = 0; i < n_res; ++i) {
    const long long c = chw[3 * i], hw = (long long)chw[3 * i + 1] * chw[3 * i + 2];
    CA_CHECK_ARG(c > 0 && hw > 0, "residual_merge: bad shape for residual %d", i);
    const long long inner = layout == CA_LAYOUT_NCFHW ? hw : c * hw;
    if (inner % vec16 != 0 || !aligned16(dst[i])) vec = 1;
    for (int k = 0; k < n_nets; ++k) {
      CA_CHECK_ARG(res[k * n_res + i] != nullptr && dst[i] != nullptr, "residual_merge: null tensor pointer");
      if (!aligned16(res[k * n_res + i])) vec = 1;
    }
  }
  MergeParams p{};
  p.n_res = n_res; p.n_nets = n_nets; p.f = f; p.b_res = b_res; p.add = add_into_dst ? 1 : 0;
  p.transpose = layout == CA_LAYOUT_NCFHW ? 1 : 0;
  unsigned long long blocks = 0;
  for (int i = 0; i < n_res; ++i) {
    const long long c = chw[3 * i], hw = (long long)chw[3 * i + 1] * chw[3 * i + 2];
    MergeTensor& t = p.t[i];
    for (int k = 0; k < n_nets; ++k) {
      t.src[k] = res[k * n_res + i];
      t.scale[k] = scales[k * n_res + i];
    }
    t.dst = dst[i];
    t.c = (int)c;
    t.hwv = (int)(hw / vec);
    t.src_batch_vecs = (long long)f * c * hw / vec;
    t.nvec_dst = t.src_batch_vecs * b_dst;
    t.block_begin = (unsigned int)blocks;
    blocks += (t.nvec_dst + kThreads * kVecPerThread - 1) / (kThreads * kVecPerThread);
  }
  CA_CHECK_ARG(blocks < (1ull << 31), "residual_merge: grid too large");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const unsigned int grid = (unsigned int)blocks;
  const int rc = dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    constexpr int V = Traits<T>::kVec;
#define CA_MERGE_LAUNCH(VEC_, NETS_) residual_merge_kernel<T, VEC_, NETS_><<<grid, kThreads, 0, st>>>(p)
    if (vec > 1) {
      switch (n_nets) {
        case 1: CA_MERGE_LAUNCH(V, 1); break;
        case 2: CA_MERGE_LAUNCH(V, 2); break;
        case 3: CA_MERGE_LAUNCH(V, 3); break;
        case 4: CA_MERGE_LAUNCH(V, 4); break;
        default: CA_MERGE_LAUNCH(V, 0); break;
      }
    } else {
      CA_MERGE_LAUNCH(1, 0);
    }
#undef CA_MERGE_LAUNCH
    return CA_OK;
  });
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}
