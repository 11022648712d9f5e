// Kernel (2), native-layout streaming path: GroupNorm + SiLU (+ time-embedding add) on a BFHWC video activation as two
// plain streaming kernels whose second read is served by the 126 MB L2:
//
//     statistics : flat grid, one short-lived CTA per slice (k*8 rows of one statistics domain): every thread has all 8 of
//                  its 16-byte loads in flight at once, accumulates per-channel sums shifted by its OWN first row (packed
//                  f32x2; no load depends on another), the row lanes are merged with Chan's formula, channels folded to
//                  per-group (mean, M2): one 8-byte partial per (domain, slice, group)
//     finalize   : one CTA per domain folds the partials (slice order, double: deterministic) -> (mean, rstd) per group
//     normalise  : flat grid again, domains walked in REVERSE order of the statistics launch, so the most recently read
//                  lines are still L2-resident; loads first, then the per-channel scale / shift, then streaming
//                  (evict-first) stores that do not displace x
//
// HBM traffic stays at the algorithmic 2*N*s bytes (one read + one write) as long as the chunk of domains handed to the
// pair of launches fits the L2; larger tensors are cut into chunks of whole domains on the host (domains are independent).
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31, 191-192, 199-208;
// unet.py:614-615) and the transformer-entry GroupNorms (motion_module.py:144, attention.py:131).
//
// Status: an ALTERNATIVE to groupnorm_ring.cu / groupnorm_slab.cu, off by default (CA_GN_STREAM=1 selects it), kept as the
// measured record of the "exchange-free flat launches" idea (profiles/r01d_notes.md §1, §7): 86.7 us at c320 64x64 in this
// form (72.7 us in a first form with 16 rows per thread and the fold inside the normalise CTAs) against 60 us for the slice
// ring.  The per-CTA statistics epilogue and the per-thread scale/shift prologue cost as many issue slots as the 8 rows a
// thread streams, so the two passes run at about half the speed of the plain elementwise epilogue kernel they imitate.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_team.cuh"

namespace ca {
namespace {

constexpr int kSThreads = 256;
constexpr int kVecE = 8;
constexpr int kBatch = 8;  // independent 16-byte loads in flight per thread

struct StreamParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;
  long long temb_ld;
  int c, groups, cpg;       // full row
  int slabs, cs, gs;        // channel slabs of whole groups (blockIdx.y), channels / groups per slab
  int nvec, k, gl;          // per slab: 16-byte vectors per row, row lanes, lanes per group in the group fold
  int per_frame, f;
  float eps;
  int dom_rows, slice_rows, spd;
  int dom0, ndom;           // this launch covers domains [dom0, dom0 + ndom)
  float2* partials;         // [domains][spd][groups] (mean, M2)
  float2* finals;           // [domains][groups] (mean, rstd)
};

__device__ __forceinline__ float tanh_fast_s(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ void stg_evict_first(void* p, const uint4& v) {
  asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ uint4 ldg_l2_keep(const void* p) {  // no L1 allocation, normal L2 residency
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}

template <typename T>
__global__ void __launch_bounds__(kSThreads) gn_stream_stats_kernel(const StreamParams p) {
  extern __shared__ __align__(16) float s_dyn[];
  const int Cs = p.cs, nvec = p.nvec, k = p.k;
  float* s_part = s_dyn;                        // [k][3][Cs]: x0 (the lane's own shift), sum(x - x0), sum((x - x0)^2)
  float* s_ch = s_part + (size_t)k * 3 * Cs;    // [2][Cs]
  const int tid = threadIdx.x;
  const int slab = blockIdx.y;
  const int dl = blockIdx.x / p.spd, sl = blockIdx.x - dl * p.spd;
  const int dom = p.dom0 + dl;
  const int r0 = sl * p.slice_rows;
  const int rows = min(p.slice_rows, p.dom_rows - r0);
  const int bi = p.per_frame ? dom / p.f : dom;
  const bool on = tid < nvec * k;
  const int cv = tid % nvec, rl = tid / nvec;
  const T* xt = reinterpret_cast<const T*>(p.x) + ((long long)dom * p.dom_rows + r0) * p.c + (long long)slab * Cs + cv * kVecE;

  if (on) {
    // slice_rows = k * kBatch: every load of this thread is in flight at once; the shift is the thread's OWN first row, so no
    // load depends on another (row lanes are merged with Chan's formula below)
    uint4 raw[kBatch];
#pragma unroll
    for (int u = 0; u < kBatch; ++u) {
      const int r = rl + u * k;
      raw[u] = r < rows ? ldg_l2_keep(xt + (long long)r * p.c) : make_uint4(0u, 0u, 0u, 0u);
    }
    float2 x0[4], s1[4], s2[4];
    unpack2(raw[0].x, x0[0].x, x0[0].y, T());
    unpack2(raw[0].y, x0[1].x, x0[1].y, T());
    unpack2(raw[0].z, x0[2].x, x0[2].y, T());
    unpack2(raw[0].w, x0[3].x, x0[3].y, T());
    float2 nx0[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      nx0[e] = make_float2(-x0[e].x, -x0[e].y);
      s1[e] = s2[e] = make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int u = 1; u < kBatch; ++u) {
      if (rl + u * k < rows) {
        float2 v[4];
        unpack2(raw[u].x, v[0].x, v[0].y, T());
        unpack2(raw[u].y, v[1].x, v[1].y, T());
        unpack2(raw[u].z, v[2].x, v[2].y, T());
        unpack2(raw[u].w, v[3].x, v[3].y, T());
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 d = __fadd2_rn(v[e], nx0[e]);
          s1[e] = __fadd2_rn(s1[e], d);
          s2[e] = __ffma2_rn(d, d, s2[e]);
        }
      }
    }
    float4* d0 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 3 + 0) * Cs + cv * kVecE);
    float4* d1 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 3 + 1) * Cs + cv * kVecE);
    float4* d2 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 3 + 2) * Cs + cv * kVecE);
    d0[0] = make_float4(x0[0].x, x0[0].y, x0[1].x, x0[1].y);
    d0[1] = make_float4(x0[2].x, x0[2].y, x0[3].x, x0[3].y);
    d1[0] = make_float4(s1[0].x, s1[0].y, s1[1].x, s1[1].y);
    d1[1] = make_float4(s1[2].x, s1[2].y, s1[3].x, s1[3].y);
    d2[0] = make_float4(s2[0].x, s2[0].y, s2[1].x, s2[1].y);
    d2[1] = make_float4(s2[2].x, s2[2].y, s2[3].x, s2[3].y);
  }
  __syncthreads();
  {  // per-channel (mean, M2) of the slice: Chan merge of the row lanes in lane order (fixed order: deterministic)
    const float* tp = p.temb ? p.temb + (long long)bi * p.temb_ld + (long long)slab * Cs : nullptr;
    for (int c0 = tid; c0 < Cs; c0 += kSThreads) {
      float n = 0.f, mean = 0.f, m2 = 0.f;
      for (int q = 0; q < k; ++q) {
        const int nq_i = q < rows ? (rows - q + k - 1) / k : 0;
        if (nq_i == 0) break;
        const float nq = (float)nq_i;
        const float a1 = s_part[((size_t)q * 3 + 1) * Cs + c0], a2 = s_part[((size_t)q * 3 + 2) * Cs + c0];
        const float dm = a1 / nq;
        const float mq = s_part[((size_t)q * 3 + 0) * Cs + c0] + dm;
        const float m2q = fmaxf(a2 - a1 * dm, 0.f);
        const float nt = n + nq, delta = mq - mean;
        mean += delta * (nq / nt);
        m2 += m2q + delta * delta * (n * nq / nt);
        n = nt;
      }
      s_ch[c0] = mean + (tp ? __ldg(tp + c0) : 0.f);
      s_ch[Cs + c0] = m2;
    }
  }
  __syncthreads();
  {  // per-group (mean, M2) from the cpg channels (equal counts: Chan's formula with n_c = rows)
    const int L = p.gl;
    const float inv_cpg = 1.0f / (float)p.cpg;
    for (int g0 = 0; g0 < p.gs; g0 += kSThreads / L) {
      const int g = g0 + tid / L, l = tid % L;
      const float* mc = s_ch + g * p.cpg;
      float sm = 0.f, sq = 0.f;
      if (g < p.gs)
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
      if (g < p.gs)
        for (int e = l; e < p.cpg; e += L) {
          const float d = mc[e] - gmean;
          dv = fmaf(d, d, dv);
        }
      for (int o = L >> 1; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
      if (l == 0 && g < p.gs)
        p.partials[((long long)dom * p.spd + sl) * p.groups + slab * p.gs + g] = make_float2(gmean, fmaf((float)rows, dv, sq));
    }
  }
}

// one CTA per domain: fold the slice partials in slice order, double (deterministic) -> (mean, rstd) per group
__global__ void __launch_bounds__(kSThreads) gn_stream_finalize_kernel(const StreamParams p) {
  __shared__ double s_fold[4][8][33];
  const int dom = p.dom0 + blockIdx.x;
  const int tid = threadIdx.x, lane_q = tid >> 5, lane_g = tid & 31;
  const float2* part = p.partials + (long long)dom * p.spd * p.groups;
  for (int g0 = 0; g0 < p.groups; g0 += 32) {
    const int g = g0 + lane_g;
    double a_n = 0, a_m = 0, a_mm = 0, a_q = 0;
    if (g < p.groups) {
#pragma unroll 4
      for (int q = lane_q; q < p.spd; q += 8) {
        const float2 v = __ldcg(part + (long long)q * p.groups + g);
        const double nk = (double)(min(p.slice_rows, p.dom_rows - q * p.slice_rows)) * p.cpg;
        const double m = (double)v.x;
        a_n += nk;
        a_m += nk * m;
        a_mm += nk * m * m;
        a_q += (double)v.y;
      }
    }
    s_fold[0][lane_q][lane_g] = a_n;
    s_fold[1][lane_q][lane_g] = a_m;
    s_fold[2][lane_q][lane_g] = a_mm;
    s_fold[3][lane_q][lane_g] = a_q;
    __syncthreads();
    if (lane_q == 0 && g < p.groups) {
      double tn = 0, tm = 0, tmm = 0, tq = 0;
      for (int l = 0; l < 8; ++l) {
        tn += s_fold[0][l][lane_g];
        tm += s_fold[1][l][lane_g];
        tmm += s_fold[2][l][lane_g];
        tq += s_fold[3][l][lane_g];
      }
      const double mean = tm / tn;
      double var = (tq + tmm - tn * mean * mean) / tn;
      if (var < 0) var = 0;
      p.finals[(long long)dom * p.groups + g] = make_float2((float)mean, rsqrtf((float)var + p.eps));
    }
    __syncthreads();
  }
}

template <typename T, bool kSilu>
__global__ void __launch_bounds__(kSThreads) gn_stream_apply_kernel(const StreamParams p) {
  const int Cs = p.cs, nvec = p.nvec, k = p.k;
  const int tid = threadIdx.x;
  const int slab = blockIdx.y;
  // reverse domain order: the statistics launch read domain ndom-1 last, so it is the most likely to still be in L2
  const int dl = p.ndom - 1 - (int)(blockIdx.x / p.spd), sl = blockIdx.x % p.spd;
  const int dom = p.dom0 + dl;
  const int r0 = sl * p.slice_rows;
  const int rows = min(p.slice_rows, p.dom_rows - r0);
  const int bi = p.per_frame ? dom / p.f : dom;
  if (tid >= nvec * k) return;
  const int cv = tid % nvec, rl = tid / nvec;
  const long long co = (long long)slab * Cs + cv * kVecE;
  const T* xt = reinterpret_cast<const T*>(p.x) + ((long long)dom * p.dom_rows + r0) * p.c + co;
  T* yt = reinterpret_cast<T*>(p.y) + ((long long)dom * p.dom_rows + r0) * p.c + co;

  // all rows of this thread are requested before the per-channel scale / shift is assembled (its L2 round trip hides behind them)
  uint4 raw[kBatch];
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    const int r = rl + u * k;
    raw[u] = r < rows ? ldg_stream(xt + (long long)r * p.c) : make_uint4(0u, 0u, 0u, 0u);
  }
  float2 av[4], bv[4];
  {
    const float2* fin = p.finals + (long long)dom * p.groups;
    const float* tp = p.temb ? p.temb + (long long)bi * p.temb_ld + co : nullptr;
#pragma unroll
    for (int e = 0; e < kVecE; e += 2) {
      const float2 m0 = __ldg(fin + (int)((co + e) / p.cpg)), m1 = __ldg(fin + (int)((co + e + 1) / p.cpg));
      float a0 = __ldg(p.gamma + co + e) * m0.y, a1 = __ldg(p.gamma + co + e + 1) * m1.y;
      const float t0 = tp ? __ldg(tp + e) : 0.f, t1 = tp ? __ldg(tp + e + 1) : 0.f;
      float b0 = fmaf(t0 - m0.x, a0, __ldg(p.beta + co + e)), b1 = fmaf(t1 - m1.x, a1, __ldg(p.beta + co + e + 1));
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
#pragma unroll
  for (int u = 0; u < kBatch; ++u) {
    const int r = rl + u * k;
    if (r < rows) {
      float2 v[4];
      unpack2(raw[u].x, v[0].x, v[0].y, T());
      unpack2(raw[u].y, v[1].x, v[1].y, T());
      unpack2(raw[u].z, v[2].x, v[2].y, T());
      unpack2(raw[u].w, v[3].x, v[3].y, T());
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 hh = __ffma2_rn(v[e], av[e], bv[e]);
        if constexpr (kSilu) v[e] = __ffma2_rn(hh, make_float2(tanh_fast_s(hh.x), tanh_fast_s(hh.y)), hh);
        else v[e] = hh;
      }
      uint4 out;
      out.x = pack2(v[0].x, v[0].y, T());
      out.y = pack2(v[1].x, v[1].y, T());
      out.z = pack2(v[2].x, v[2].y, T());
      out.w = pack2(v[3].x, v[3].y, T());
      stg_evict_first(yt + (long long)r * p.c, out);
    }
  }
}

struct StreamPlan {
  int domains, dom_rows, slabs, cs, gs, nvec, k, gl, slice_rows, spd, chunk_domains;
  size_t smem_stats, smem_apply, partial_bytes;
};

int s_env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

bool make_stream_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype, StreamPlan* pl) {
  static const int on = s_env_int("CA_GN_STREAM", 0);  // opt-in: r01d measured it slower than the slice ring (see the header)
  static const int rows_per_thread = s_env_int("CA_GN_STREAM_ROWS", 16);
  static const int chunk_mb = s_env_int("CA_GN_STREAM_MB", 96);
  if (!on) return false;
  if (dtype != CA_BF16 && dtype != CA_F16) return false;
  if (c % kVecE != 0 || groups <= 0 || c % groups != 0) return false;
  int slabs = 1;
  while (slabs <= groups && (groups % slabs != 0 || (c / slabs) % kVecE != 0 || c / slabs / kVecE > kSThreads)) ++slabs;
  if (slabs > groups || slabs > 65535) return false;
  const int cs = c / slabs, gs = groups / slabs, nvec = cs / kVecE;
  const long long rows = per_frame ? (long long)h * w : (long long)f * h * w;
  const long long domains = per_frame ? (long long)b * f : b;
  if (rows <= 0 || rows >= (1ll << 30) || domains <= 0 || domains >= (1ll << 24)) return false;
  const int k = kSThreads / nvec;
  (void)rows_per_thread;
  long long slice_rows = (long long)k * kBatch;  // every thread holds all of its rows in registers at once
  if (slice_rows > rows) slice_rows = rows;
  const long long spd = (rows + slice_rows - 1) / slice_rows;
  if (spd >= (1ll << 20) || domains * spd >= (1ll << 31)) return false;
  pl->domains = (int)domains;
  pl->dom_rows = (int)rows;
  pl->slabs = slabs; pl->cs = cs; pl->gs = gs; pl->nvec = nvec; pl->k = k;
  pl->gl = 1;
  while (pl->gl < 32 && pl->gl * 2 <= c / groups && pl->gl * 2 * gs <= kSThreads) pl->gl *= 2;
  pl->slice_rows = (int)slice_rows;
  pl->spd = (int)spd;
  const long long dom_bytes = rows * (long long)c * 2;
  long long cd = ((long long)chunk_mb << 20) / dom_bytes;
  if (cd < 1) cd = 1;
  if (cd > domains) cd = domains;
  pl->chunk_domains = (int)cd;
  pl->smem_stats = sizeof(float) * ((size_t)k * 3 * cs + 2 * (size_t)cs);
  pl->smem_apply = 0;
  pl->partial_bytes = sizeof(float2) * ((size_t)domains * spd * groups + (size_t)domains * groups);
  return pl->smem_stats <= 160 * 1024;
}

}  // namespace

size_t gn_stream_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype) {
  StreamPlan pl;
  if (!make_stream_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return 0;
  return pl.partial_bytes;
}

int gn_stream_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c,
                     int f, int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                     size_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  StreamPlan pl;
  if (!make_stream_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return CA_OK;
  if (!aligned16(x) || !aligned16(y)) return CA_OK;
  if (!workspace || workspace_bytes < pl.partial_bytes || (reinterpret_cast<uintptr_t>(workspace) & 7u)) return CA_OK;

  StreamParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld;
  p.c = c; p.groups = groups; p.cpg = c / groups;
  p.slabs = pl.slabs; p.cs = pl.cs; p.gs = pl.gs; p.nvec = pl.nvec; p.k = pl.k; p.gl = pl.gl;
  p.per_frame = per_frame ? 1 : 0; p.f = f; p.eps = eps;
  p.dom_rows = pl.dom_rows; p.slice_rows = pl.slice_rows; p.spd = pl.spd;
  p.partials = reinterpret_cast<float2*>(workspace);
  p.finals = p.partials + (size_t)pl.domains * pl.spd * groups;

  return dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if constexpr (sizeof(T) != 2) {
      return CA_OK;
    } else {
      auto stats = gn_stream_stats_kernel<T>;
      const void* apply = apply_silu ? (const void*)gn_stream_apply_kernel<T, true> : (const void*)gn_stream_apply_kernel<T, false>;
      CA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(stats), pl.smem_stats));
      for (int d0 = 0; d0 < pl.domains; d0 += pl.chunk_domains) {
        p.dom0 = d0;
        p.ndom = pl.domains - d0 < pl.chunk_domains ? pl.domains - d0 : pl.chunk_domains;
        const dim3 grid((unsigned)((long long)p.ndom * pl.spd), (unsigned)pl.slabs);
        stats<<<grid, kSThreads, pl.smem_stats, st>>>(p);
        gn_stream_finalize_kernel<<<(unsigned)p.ndom, kSThreads, 0, st>>>(p);
        void* args[] = {(void*)&p};
        CA_CUDA(cudaLaunchKernel(apply, grid, dim3(kSThreads), args, 0, st));
      }
      CA_CUDA(cudaGetLastError());
      *handled = true;
      return CA_OK;
    }
  });
}

}  // namespace ca
