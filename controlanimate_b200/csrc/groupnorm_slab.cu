// Kernel (2), native-layout slab path: GroupNorm + SiLU (+ time-embedding add) on a BFHWC video activation when the rows of
// a statistics domain for a slab of whole groups fit the shared memory of ONE CTA (16x16 / 8x8 latent levels) or of one
// thread-block CLUSTER of up to 8 CTAs (32x32 ... 96x96).  Groups are independent, so the CTAs that own every row of
// (domain, S groups) exchange nothing with the rest of the grid: cp.async the [rows][S*cpg] slab into smem, statistics,
// [per-group (mean, M2) partials pushed into every peer's smem over DSMEM + one barrier.cluster], normalise, store -- one
// plain launch, no workspace, no spin-waits, one HBM read + one HBM write.
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31, 191-192, 199-208) and the
// per-frame transformer-entry GroupNorms (motion_module.py:144, attention.py:131) at those levels.
//
// Why next to groupnorm_ring.cu: the ring's cooperative launch, table preset and folder round trip cost ~15-20 us no
// matter how small the tensor is (r01d: 21 us for the 10 MB c1280 8x8 activation, 7 % of the copy roofline), and two
// thirds of the 136 GroupNorm launches of a denoising step are at the small levels.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_paths.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kSlabThreads = 256;
constexpr int kVecE = 8;
constexpr size_t kSlabTileTarget = 40 * 1024;  // per-CTA slab the plan aims for (4-5 CTAs per SM overlap load / compute / store)
constexpr size_t kSlabTileCap = 96 * 1024;     // hard limit (2 CTAs per SM)
constexpr int kMaxCluster = 8;

struct SlabParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;
  long long temb_ld;
  int c, groups, cpg;
  int S, seg_c, nvs, k, gl;  // groups per slab, channels per slab, 16-byte vectors per slab row, row lanes, lanes per group
  int slabs_per_dom;
  int per_frame, f;
  float eps;
  int dom_rows;
  int R, rows_per;  // cluster size (CTAs sharing one (domain, slab)) and rows per CTA
};

__device__ __forceinline__ float tanh_fast_b(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void cp_async16(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}

template <typename T, bool kSilu>
__global__ void __launch_bounds__(kSlabThreads) gn_slab_kernel(const SlabParams p) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int Cs = p.seg_c, nvs = p.nvs, k = p.k, R = p.R;
  uint4* tile = reinterpret_cast<uint4*>(s_raw);                                          // [rows_per][nvs]
  float* s_part = reinterpret_cast<float*>(s_raw + (size_t)p.rows_per * Cs * sizeof(T));  // [k][2][Cs]
  float* s_ch = s_part + (size_t)k * 2 * Cs;                                              // [2][Cs]
  float2* s_fin = reinterpret_cast<float2*>(s_ch + 2 * Cs);                               // [S] (mean, rstd)
  float2* s_x = s_fin + p.S;                                                              // [R][S] partials of the cluster

  const int tid = threadIdx.x;
  const int rank = R > 1 ? (int)cluster_ctarank() : 0;
  const int unit = blockIdx.x / R;  // (domain, slab)
  const int dom = unit / p.slabs_per_dom, slab = unit - dom * p.slabs_per_dom;
  const int bi = p.per_frame ? dom / p.f : dom;
  const bool on = tid < nvs * k;
  const int cv = tid % nvs, rl = tid / nvs;
  const long long co = (long long)slab * Cs;
  const int r_begin = rank * p.rows_per;
  const int rows = min(p.rows_per, p.dom_rows - r_begin);  // >= 1 by construction of the plan
  const T* xs = reinterpret_cast<const T*>(p.x) + ((long long)dom * p.dom_rows + r_begin) * p.c + co + cv * kVecE;
  T* ys = reinterpret_cast<T*>(p.y) + ((long long)dom * p.dom_rows + r_begin) * p.c + co + cv * kVecE;

  if (on)
    for (int r = rl; r < rows; r += k) cp_async16(&tile[r * nvs + cv], xs + (long long)r * p.c);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // "every CTA of the cluster has started" must hold before anyone writes into a peer's shared memory: arrive now, wait only
  // right before the DSMEM pushes, so the barrier latency hides behind the load and the statistics
  if (R > 1) asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncthreads();

  // ---- statistics: per-channel sums shifted by row 0 (packed f32x2), row lanes added in fixed order, channels -> groups ----
  if (on) {
    float2 nx0[4], s1[4], s2[4];
    {
      const uint4 v0 = tile[cv];
      unpack2(v0.x, nx0[0].x, nx0[0].y, T());
      unpack2(v0.y, nx0[1].x, nx0[1].y, T());
      unpack2(v0.z, nx0[2].x, nx0[2].y, T());
      unpack2(v0.w, nx0[3].x, nx0[3].y, T());
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      nx0[e] = make_float2(-nx0[e].x, -nx0[e].y);
      s1[e] = s2[e] = make_float2(0.f, 0.f);
    }
#pragma unroll 4
    for (int r = rl; r < rows; r += k) {
      const uint4 raw = tile[r * nvs + cv];
      float2 v[4];
      unpack2(raw.x, v[0].x, v[0].y, T());
      unpack2(raw.y, v[1].x, v[1].y, T());
      unpack2(raw.z, v[2].x, v[2].y, T());
      unpack2(raw.w, v[3].x, v[3].y, T());
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 d = __fadd2_rn(v[e], nx0[e]);
        s1[e] = __fadd2_rn(s1[e], d);
        s2[e] = __ffma2_rn(d, d, s2[e]);
      }
    }
    float4* d1 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 2 + 0) * Cs + cv * kVecE);
    float4* d2 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 2 + 1) * Cs + cv * kVecE);
    d1[0] = make_float4(s1[0].x, s1[0].y, s1[1].x, s1[1].y);
    d1[1] = make_float4(s1[2].x, s1[2].y, s1[3].x, s1[3].y);
    d2[0] = make_float4(s2[0].x, s2[0].y, s2[1].x, s2[1].y);
    d2[1] = make_float4(s2[2].x, s2[2].y, s2[3].x, s2[3].y);
  }
  __syncthreads();
  {
    const T* row0 = reinterpret_cast<const T*>(tile);
    const float inv_n = 1.0f / (float)rows;
    const float* tp = p.temb ? p.temb + (long long)bi * p.temb_ld + co : nullptr;
    for (int c0 = tid; c0 < Cs; c0 += kSlabThreads) {
      float a1 = 0.f, a2 = 0.f;
      for (int q = 0; q < k; ++q) {
        a1 += s_part[((size_t)q * 2 + 0) * Cs + c0];
        a2 += s_part[((size_t)q * 2 + 1) * Cs + c0];
      }
      const float t = tp ? __ldg(tp + c0) : 0.f;
      const float dm = a1 * inv_n;
      s_ch[c0] = Traits<T>::to_f(row0[c0]) + t + dm;
      s_ch[Cs + c0] = fmaxf(a2 - a1 * dm, 0.f);
    }
  }
  __syncthreads();
  if (R > 1) asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  {
    const int L = p.gl;
    const float inv_cpg = 1.0f / (float)p.cpg;
    for (int g0 = 0; g0 < p.S; g0 += kSlabThreads / L) {
      const int g = g0 + tid / L, l = tid % L;
      const float* mc = s_ch + g * p.cpg;
      float sm = 0.f, sq = 0.f;
      if (g < p.S)
        for (int e = l; e < p.cpg; e += L) {
          sm += mc[e];
          sq += mc[Cs + e];
        }
      for (int o = L >> 1; o > 0; o >>= 1) {
        sm += __shfl_xor_sync(0xffffffffu, sm, o);
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
      }
      const float gmean = sm * inv_cpg;
      float dv = 0.f;
      if (g < p.S)
        for (int e = l; e < p.cpg; e += L) {
          const float d = mc[e] - gmean;
          dv = fmaf(d, d, dv);
        }
      for (int o = L >> 1; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
      if (l == 0 && g < p.S) {
        const float m2 = fmaf((float)rows, dv, sq);
        if (R == 1) {
          s_fin[g] = make_float2(gmean, rsqrtf(m2 / ((float)rows * (float)p.cpg) + p.eps));
        } else {  // push this CTA's (mean, M2) into slot [rank][g] of every CTA of the cluster (DSMEM)
          for (int pr = 0; pr < R; ++pr) {
            const uint32_t dst = mapa_u32(&s_x[rank * p.S + g], (uint32_t)pr);
            asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(dst), "f"(gmean), "f"(m2) : "memory");
          }
        }
      }
    }
  }
  if (R > 1) {
    cluster_sync_all();  // release / acquire: all partials have landed; no remote access happens after this point
    for (int g = tid; g < p.S; g += kSlabThreads) {  // rank order, double: deterministic and identical in every CTA
      double tn = 0, tm = 0, tmm = 0, tq = 0;
      for (int pr = 0; pr < R; ++pr) {
        const double nk = (double)min(p.rows_per, p.dom_rows - pr * p.rows_per) * p.cpg;
        const float2 v = s_x[pr * p.S + g];
        tn += nk;
        tm += nk * (double)v.x;
        tmm += nk * (double)v.x * (double)v.x;
        tq += (double)v.y;
      }
      const double mean = tm / tn;
      double var = (tq + tmm - tn * mean * mean) / tn;
      if (var < 0) var = 0;
      s_fin[g] = make_float2((float)mean, rsqrtf((float)var + p.eps));
    }
  }
  __syncthreads();
  if (!on) return;

  // ---- normalise + affine + SiLU from smem, 16-byte streaming stores ----
  float2 av[4], bv[4];
  {
    const float* tp = p.temb ? p.temb + (long long)bi * p.temb_ld + co + cv * kVecE : nullptr;
#pragma unroll
    for (int e = 0; e < kVecE; e += 2) {
      const float2 m0 = s_fin[(cv * kVecE + e) / p.cpg], m1 = s_fin[(cv * kVecE + e + 1) / p.cpg];
      float a0 = __ldg(p.gamma + co + cv * kVecE + e) * m0.y, a1 = __ldg(p.gamma + co + cv * kVecE + e + 1) * m1.y;
      const float t0 = tp ? __ldg(tp + e) : 0.f, t1 = tp ? __ldg(tp + e + 1) : 0.f;
      float b0 = fmaf(t0 - m0.x, a0, __ldg(p.beta + co + cv * kVecE + e));
      float b1 = fmaf(t1 - m1.x, a1, __ldg(p.beta + co + cv * kVecE + e + 1));
      if constexpr (kSilu) {
        a0 *= 0.5f;
        a1 *= 0.5f;
        b0 *= 0.5f;
        b1 *= 0.5f;
      }
      av[e / 2] = make_float2(a0, a1);
      bv[e / 2] = make_float2(b0, b1);
    }
  }
#pragma unroll 4
  for (int r = rl; r < rows; r += k) {
    const uint4 raw = tile[r * nvs + cv];
    float2 v[4];
    unpack2(raw.x, v[0].x, v[0].y, T());
    unpack2(raw.y, v[1].x, v[1].y, T());
    unpack2(raw.z, v[2].x, v[2].y, T());
    unpack2(raw.w, v[3].x, v[3].y, T());
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float2 hh = __ffma2_rn(v[e], av[e], bv[e]);
      if constexpr (kSilu) v[e] = __ffma2_rn(hh, make_float2(tanh_fast_b(hh.x), tanh_fast_b(hh.y)), hh);
      else v[e] = hh;
    }
    uint4 out;
    out.x = pack2(v[0].x, v[0].y, T());
    out.y = pack2(v[1].x, v[1].y, T());
    out.z = pack2(v[2].x, v[2].y, T());
    out.w = pack2(v[3].x, v[3].y, T());
    stg_stream(ys + (long long)r * p.c, out);
  }
}

struct SlabPlan {
  int domains, rows, S, seg_c, nvs, k, gl, slabs_per_dom, R, rows_per;
  size_t smem;
};

bool make_slab_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype, SlabPlan* pl) {
  static const int on = [] { const char* e = getenv("CA_GN_SLAB"); return (e && e[0] == '0') ? 0 : 1; }();
  // r01d sweep: one CTA per (domain, slab) with slabs up to 96 KB beats both the slice ring and the clustered split wherever it
  // fits (32x32: 32.7 us vs 40 / 37.6); where it does not (64x64: 327 KB) clusters of 4 / 8 measured 62 / 70 us against the
  // ring's 60, so the clustered split is opt-in (CA_GN_SLAB_CLUSTER=2/4/8) and the ring keeps those shapes
  static const int max_r = [] { const char* e = getenv("CA_GN_SLAB_CLUSTER"); const int v = (e && e[0]) ? atoi(e) : 1; return v < 1 ? 1 : (v > kMaxCluster ? kMaxCluster : v); }();
  if (!on) return false;
  if (dtype != CA_BF16 && dtype != CA_F16) return false;
  if (c % kVecE != 0 || groups <= 0 || c % groups != 0) return false;
  const int cpg = c / groups;
  const long long rows = per_frame ? (long long)h * w : (long long)f * h * w;
  const long long domains = per_frame ? (long long)b * f : b;
  if (rows <= 0 || rows >= (1 << 20) || domains <= 0 || domains * groups * kMaxCluster >= (1ll << 31)) return false;
  // groups per slab: S | groups, slab rows 16-byte granular and >= 64 B.  Start from the smallest such S, split the rows over
  // a cluster of R CTAs until the per-CTA slab is ~40 KB; with R == 1 widen S while the slab stays small and the grid large
  int S = 0;
  for (int t = 1; t <= groups; ++t) {
    if (groups % t) continue;
    const long long seg = (long long)t * cpg;
    if (seg % kVecE || seg * 2 < 64) continue;
    if (seg / kVecE > kSlabThreads) return false;
    S = t;
    break;
  }
  if (S == 0) return false;
  int R = 1;
  while (R < max_r && (size_t)((rows + R - 1) / R) * S * cpg * 2 > kSlabTileTarget) R *= 2;
  if (R > 1 && rows < 8 * R) return false;
  long long rows_per = (rows + R - 1) / R;
  if ((rows + rows_per - 1) / rows_per != R) return false;  // every CTA of the cluster must own at least one row
  if ((size_t)rows_per * S * cpg * 2 > kSlabTileCap) return false;
  if (R == 1) {
    const long long want = 2ll * sm_count();
    for (int t = S + 1; t <= groups; ++t) {
      if (groups % t) continue;
      const long long seg = (long long)t * cpg;
      if (seg % kVecE || seg / kVecE > kSlabThreads) continue;
      if ((size_t)(rows * seg * 2) > kSlabTileTarget || domains * (groups / t) < want) break;
      S = t;
    }
  }
  pl->domains = (int)domains;
  pl->rows = (int)rows;
  pl->S = S;
  pl->seg_c = S * cpg;
  pl->nvs = pl->seg_c / kVecE;
  pl->k = kSlabThreads / pl->nvs;
  pl->gl = 1;
  while (pl->gl < 32 && pl->gl * 2 <= cpg && pl->gl * 2 * S <= kSlabThreads) pl->gl *= 2;
  pl->slabs_per_dom = groups / S;
  pl->R = R;
  pl->rows_per = (int)rows_per;
  pl->smem = (size_t)rows_per * pl->seg_c * 2 + sizeof(float) * ((size_t)pl->k * 2 * pl->seg_c + 2 * (size_t)pl->seg_c) +
             sizeof(float2) * (size_t)S * (1 + R);
  return pl->smem <= 110 * 1024;
}

}  // namespace

int gn_slab_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c,
                   int f, int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, cudaStream_t st,
                   bool* handled) {
  *handled = false;
  SlabPlan pl;
  if (!make_slab_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return CA_OK;
  if (!aligned16(x) || !aligned16(y)) return CA_OK;
  SlabParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld;
  p.c = c; p.groups = groups; p.cpg = c / groups;
  p.S = pl.S; p.seg_c = pl.seg_c; p.nvs = pl.nvs; p.k = pl.k; p.gl = pl.gl; p.slabs_per_dom = pl.slabs_per_dom;
  p.per_frame = per_frame ? 1 : 0; p.f = f; p.eps = eps; p.dom_rows = pl.rows; p.R = pl.R; p.rows_per = pl.rows_per;
  const void* fn = nullptr;
  if (dtype == CA_BF16) fn = apply_silu ? (const void*)gn_slab_kernel<__nv_bfloat16, true> : (const void*)gn_slab_kernel<__nv_bfloat16, false>;
  else fn = apply_silu ? (const void*)gn_slab_kernel<__half, true> : (const void*)gn_slab_kernel<__half, false>;
  CA_CUDA(ensure_dynamic_smem(fn, pl.smem));
  void* args[] = {(void*)&p};
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((long long)pl.domains * pl.slabs_per_dom * pl.R));
  cfg.blockDim = dim3(kSlabThreads);
  cfg.dynamicSmemBytes = pl.smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)pl.R;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pl.R > 1 ? 1 : 0;
  CA_CUDA(cudaLaunchKernelExC(&cfg, fn, args));
  *handled = true;
  return CA_OK;
}

}  // namespace ca
