// Kernel (2), native-layout fast path: GroupNorm + SiLU (+ time-embedding add) on a BFHWC video
// activation with ONE HBM read and ONE HBM write per element (algorithmic bytes 2*N*s).
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31, 191-192,
// 199-208; unet.py:614-615) and the transformer-entry GroupNorms (motion_module.py:144, attention.py:131).
//
// Design (B200): in BFHWC a statistics domain (b[, frame]) is one contiguous slab of rows*C elements.
// The grid is persistent and co-resident (cooperative launch, 2 CTAs per SM) and organised in TEAMS of
// `team` CTAs.  A team takes one domain at a time; member m owns a contiguous band of rows which it pulls
// into shared memory with 1-D bulk async copies (cp.async.bulk + mbarrier: no registers, no per-thread
// address math, full memory-level parallelism from a single issuing thread).  The band stays resident in
// smem while the team agrees on the statistics:
//     band -> exact two-pass (sum, centred M2) per group from smem
//          -> (mean, M2) partial to a tiny global table, one arrival on the domain's counter
//          -> every member combines the team's partials with Chan's formula in double, fixed order
//             (deterministic, no floating-point atomics)
//          -> normalise + affine + SiLU from smem, 16-byte streaming stores to y.
// While one CTA of an SM waits at its team barrier the other CTA of that SM is loading or storing, so
// HBM stays busy.  SiLU is evaluated as h*(1+tanh(h)), h = o/2: ONE MUFU op per element (ex2+rcp would
// need two and cap the kernel at 16 MUFU/clk/SM = 70% of the HBM roofline).
//
// Shapes that do not fit (domain larger than the team's combined smem, fp32 storage, c % 8 != 0) fall
// back to the split statistics/finalize/apply kernels in groupnorm_silu.cu.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_team.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kTeamThreads = 512;
constexpr int kCtasPerSm = 2;
constexpr size_t kSmemCap = 104 * 1024;  // dynamic smem per CTA so that two CTAs fit one SM
constexpr int kVecE = 8;                 // 16-bit elements per 16-byte vector

struct TeamParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;  // [b, c] (row stride temb_ld) or null
  long long temb_ld;
  int c, groups, cpg, nvec, k;
  int gl;  // lanes per group in the in-CTA group reduction
  int per_frame, f;
  float eps;
  int dom_rows, domains, team, n_teams, slice_rows;
  float2* partials;        // [domains][team][groups] (mean, M2)
  unsigned int* counters;  // [domains]
};

__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ float tanh_fast(float v) {
  float r;
  asm("tanh.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// Sum of the cpg per-channel values of group (threadIdx.x / L) with L = p.gl lanes per group (power of two <= 32):
// strided partial sums then a butterfly inside the L-lane segment -> fixed order, deterministic.
__device__ __forceinline__ float group_sum(const float* s_a, int groups, int cpg, int L) {
  const int g = threadIdx.x / L, l = threadIdx.x % L;
  float s = 0.f;
  if (g < groups)
    for (int e = l; e < cpg; e += L) s += s_a[g * cpg + e];
  for (int o = L >> 1; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  return s;
}

template <typename T, bool kSilu>
__global__ void __launch_bounds__(kTeamThreads, kCtasPerSm) gn_team_kernel(const TeamParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ double s_n[8][33], s_m[8][33], s_q[8][33];
  __shared__ __align__(8) uint64_t s_bar;

  const int C = p.c, nvec = p.nvec, k = p.k;
  const size_t buf_bytes = (size_t)p.slice_rows * C * sizeof(T);
  const uint4* bufv = reinterpret_cast<const uint4*>(smem_raw);
  float* s_a = reinterpret_cast<float*>(smem_raw + buf_bytes);  // [C]
  float* s_gmean = s_a + C;                                     // [groups]
  float* s_grstd = s_gmean + p.groups;                          // [groups]
  float* s_part = s_grstd + p.groups;                           // [k][C]

  const int team_id = blockIdx.x / p.team, member = blockIdx.x - team_id * p.team;
  const int r0 = member * p.slice_rows;
  const int rows = min(p.dom_rows, r0 + p.slice_rows) - r0;  // >= 1 by construction of the plan
  const bool on = threadIdx.x < nvec * k;
  const int cv = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const double cnt = (double)rows * p.cpg;

  if (threadIdx.x == 0) {
    mbar_init(&s_bar, 1);
    fence_mbar_init();
  }
  __syncthreads();

  uint32_t parity = 0;
  for (int dom = team_id; dom < p.domains; dom += p.n_teams, parity ^= 1u) {
    const long long base = ((long long)dom * p.dom_rows + r0) * C;
    const T* __restrict__ xg = reinterpret_cast<const T*>(p.x) + base;
    T* __restrict__ yg = reinterpret_cast<T*>(p.y) + base + cv * kVecE;
    const int bi = p.per_frame ? dom / p.f : dom;

    // ---- band -> smem (previous iteration's generic reads are ordered before these async writes) ----
    if (threadIdx.x == 0) {
      fence_proxy_async();
      const uint32_t total = (uint32_t)((size_t)rows * C * sizeof(T));
      mbar_arrive_expect_tx(&s_bar, total);
      constexpr uint32_t kPiece = 16 * 1024;
      for (uint32_t off = 0; off < total; off += kPiece)
        bulk_load_1d(smem_raw + off, reinterpret_cast<const unsigned char*>(xg) + off, min(kPiece, total - off), &s_bar);
    }
    float tv[kVecE];
#pragma unroll
    for (int e = 0; e < kVecE; ++e) tv[e] = (p.temb && on) ? __ldg(p.temb + (long long)bi * p.temb_ld + cv * kVecE + e) : 0.f;
    mbar_wait(&s_bar, parity);

    // ---- pass 1: per-channel sums -> group means ----
    float acc[kVecE];
#pragma unroll
    for (int e = 0; e < kVecE; ++e) acc[e] = 0.f;
    int my_rows = 0;
    if (on) {
#pragma unroll 4
      for (int r = rl; r < rows; r += k) {
        float fv[kVecE];
        Vec16<T> vv;
        vv.raw = bufv[r * nvec + cv];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < kVecE; ++e) acc[e] += fv[e];
        ++my_rows;
      }
#pragma unroll
      for (int e = 0; e < kVecE; ++e) s_part[rl * C + cv * kVecE + e] = fmaf((float)my_rows, tv[e], acc[e]);
    }
    __syncthreads();
    for (int c0 = threadIdx.x; c0 < C; c0 += kTeamThreads) {  // fixed-order reduction over row lanes
      float t = 0.f;
      for (int q = 0; q < k; ++q) t += s_part[q * C + c0];
      s_a[c0] = t;
    }
    __syncthreads();
    {
      const float s = group_sum(s_a, p.groups, p.cpg, p.gl);
      if (threadIdx.x % p.gl == 0 && threadIdx.x / p.gl < p.groups) s_gmean[threadIdx.x / p.gl] = (float)((double)s / cnt);
    }
    __syncthreads();

    // ---- pass 2: centred second moment around the band's group means ----
    if (on) {
      float gm[kVecE];
#pragma unroll
      for (int e = 0; e < kVecE; ++e) {
        gm[e] = tv[e] - s_gmean[(cv * kVecE + e) / p.cpg];
        acc[e] = 0.f;
      }
#pragma unroll 4
      for (int r = rl; r < rows; r += k) {
        float fv[kVecE];
        Vec16<T> vv;
        vv.raw = bufv[r * nvec + cv];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < kVecE; ++e) {
          const float dlt = fv[e] + gm[e];
          acc[e] = fmaf(dlt, dlt, acc[e]);
        }
      }
#pragma unroll
      for (int e = 0; e < kVecE; ++e) s_part[rl * C + cv * kVecE + e] = acc[e];
    }
    __syncthreads();
    for (int c0 = threadIdx.x; c0 < C; c0 += kTeamThreads) {
      float t = 0.f;
      for (int q = 0; q < k; ++q) t += s_part[q * C + c0];
      s_a[c0] = t;
    }
    __syncthreads();

    const float m2_band = group_sum(s_a, p.groups, p.cpg, p.gl);
    const bool g_owner = threadIdx.x % p.gl == 0 && threadIdx.x / p.gl < p.groups;
    if (p.team == 1) {
      if (g_owner) s_grstd[threadIdx.x / p.gl] = rsqrtf((float)((double)m2_band / cnt) + p.eps);
    } else {
      // ---- publish the band's partials, arrive on the domain counter, wait for the team ----
      float2* part = p.partials + (long long)dom * p.team * p.groups;
      if (g_owner) part[member * p.groups + threadIdx.x / p.gl] = make_float2(s_gmean[threadIdx.x / p.gl], m2_band);
      __syncthreads();
      if (threadIdx.x == 0) {
        unsigned int* counter = p.counters + dom;
        __threadfence();
        atomicAdd(counter, 1u);
        unsigned int seen, spins = 0;
        do {
          asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
          if (++spins == (1u << 26)) {  // a protocol bug must trap, never hang the GPU
            printf("controlanimate_b200: groupnorm team barrier timed out (block %d domain %d seen %u/%d)\n",
                   (int)blockIdx.x, dom, seen, p.team);
            __trap();
          }
        } while (seen < (unsigned)p.team);
      }
      __syncthreads();
      // ---- Chan combine in double, fixed order: 8 member lanes x 32 group lanes ----
      const int lane_q = threadIdx.x >> 5, lane_g = threadIdx.x & 31;
      for (int g0 = 0; g0 < p.groups; g0 += 32) {
        const int g = g0 + lane_g;
        const bool act = lane_q < 8 && g < p.groups;
        double ntot = 0, msum = 0;
        if (act) {
          for (int q = lane_q; q < p.team; q += 8) {
            const double nk = (double)(min(p.dom_rows, (q + 1) * p.slice_rows) - q * p.slice_rows) * p.cpg;
            ntot += nk;
            msum += nk * (double)__ldcg(part + q * p.groups + g).x;
          }
        }
        if (lane_q < 8) {
          s_n[lane_q][lane_g] = ntot;
          s_m[lane_q][lane_g] = msum;
        }
        __syncthreads();
        double nt = 0, ms = 0;
        for (int l = 0; l < 8; ++l) {
          nt += s_n[l][lane_g];
          ms += s_m[l][lane_g];
        }
        const double gmean = nt > 0 ? ms / nt : 0.0;
        double m2 = 0;
        if (act) {
          for (int q = lane_q; q < p.team; q += 8) {
            const double nk = (double)(min(p.dom_rows, (q + 1) * p.slice_rows) - q * p.slice_rows) * p.cpg;
            const float2 v = __ldcg(part + q * p.groups + g);
            const double dm = (double)v.x - gmean;
            m2 += (double)v.y + nk * dm * dm;
          }
        }
        if (lane_q < 8) s_q[lane_q][lane_g] = m2;
        __syncthreads();
        if (lane_q == 0 && g < p.groups) {
          double t = 0;
          for (int l = 0; l < 8; ++l) t += s_q[l][lane_g];
          s_gmean[g] = (float)gmean;
          s_grstd[g] = rsqrtf((float)(t / nt) + p.eps);
        }
        __syncthreads();
      }
    }
    __syncthreads();

    // ---- normalise + affine + SiLU from smem, streaming 16-byte stores ----
    if (on) {
      float av[kVecE], bv[kVecE];
#pragma unroll
      for (int e = 0; e < kVecE; ++e) {
        const int ch = cv * kVecE + e;
        const int g = ch / p.cpg;
        av[e] = __ldg(p.gamma + ch) * s_grstd[g];
        bv[e] = __ldg(p.beta + ch) + (tv[e] - s_gmean[g]) * av[e];
        if constexpr (kSilu) {
          av[e] *= 0.5f;
          bv[e] *= 0.5f;
        }
      }
#pragma unroll 4
      for (int r = rl; r < rows; r += k) {
        float fv[kVecE];
        Vec16<T> vv;
        vv.raw = bufv[r * nvec + cv];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < kVecE; ++e) {
          const float hh = fmaf(fv[e], av[e], bv[e]);
          fv[e] = kSilu ? fmaf(hh, tanh_fast(hh), hh) : hh;
        }
        vv.pack(fv);
        stg_stream(yg + (long long)r * C, vv.raw);
      }
    }
    __syncthreads();  // the band buffer and the statistics scratch are free for the next domain
  }
}

struct TeamPlan {
  int domains, dom_rows, nvec, k, team, n_teams, slice_rows;
  size_t smem, partial_bytes, counter_bytes;
};

size_t scratch_bytes(int c, int groups, int k) { return sizeof(float) * ((size_t)c + 2 * (size_t)groups + (size_t)k * c); }

// Returns false when the shape is outside the fast path.
bool make_team_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype, TeamPlan* pl) {
  if (dtype != CA_BF16 && dtype != CA_F16) return false;
  if (c % kVecE != 0 || groups <= 0 || c % groups != 0 || groups > 256) return false;
  const int nvec = c / kVecE;
  if (nvec > kTeamThreads) return false;
  const long long rows = per_frame ? (long long)h * w : (long long)f * h * w;
  const long long domains = per_frame ? (long long)b * f : b;
  if (rows >= (1ll << 30) || domains >= (1ll << 30)) return false;
  const int k = kTeamThreads / nvec;
  const size_t scratch = scratch_bytes(c, groups, k);
  if (scratch + (size_t)c * 2 > kSmemCap) return false;
  const long long row_bytes = (long long)c * 2;
  const long long max_slice = (long long)(kSmemCap - scratch) / row_bytes;
  const long long min_team = (rows + max_slice - 1) / max_slice;
  const int G = kCtasPerSm * sm_count();
  if (min_team > G) return false;
  // pick the number of teams that minimises rounds * (band bytes + fixed per-round latency)
  const double fixed = 48.0 * 1024;
  double best = 1e300;
  int best_nt = 0;
  const long long nt_max = domains < G / min_team ? domains : G / min_team;
  for (long long nt = 1; nt <= nt_max; ++nt) {
    long long team = G / nt;
    if (team > rows) team = rows;
    const long long slice = (rows + team - 1) / team;
    if (slice > max_slice) continue;
    const long long rounds = (domains + nt - 1) / nt;
    const double cost = (double)rounds * ((double)slice * row_bytes + fixed);
    if (cost < best) {
      best = cost;
      best_nt = (int)nt;
    }
  }
  if (best_nt == 0) return false;
  long long team = G / best_nt;
  if (team > rows) team = rows;
  const long long slice = (rows + team - 1) / team;
  team = (rows + slice - 1) / slice;  // every member owns at least one row
  pl->domains = (int)domains;
  pl->dom_rows = (int)rows;
  pl->nvec = nvec;
  pl->k = k;
  pl->team = (int)team;
  pl->n_teams = best_nt;
  pl->slice_rows = (int)slice;
  pl->smem = (size_t)slice * row_bytes + scratch;
  pl->counter_bytes = team > 1 ? ((size_t)domains * sizeof(unsigned int) + 15) / 16 * 16 : 0;
  pl->partial_bytes = team > 1 ? sizeof(float2) * (size_t)domains * team * groups : 0;
  return true;
}

bool team_enabled() {
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CA_GN_TEAM");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on == 1;
}

}  // namespace

size_t gn_team_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype) {
  TeamPlan pl;
  if (!team_enabled() || !make_team_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return 0;
  return pl.counter_bytes + pl.partial_bytes;
}

int gn_team_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  TeamPlan pl;
  if (!team_enabled() || !make_team_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return CA_OK;
  if (!aligned16(x) || !aligned16(y)) return CA_OK;
  const size_t need = pl.counter_bytes + pl.partial_bytes;
  if (need > 0 && (!workspace || workspace_bytes < need)) return CA_OK;

  TeamParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld;
  p.c = c; p.groups = groups; p.cpg = c / groups; p.nvec = pl.nvec; p.k = pl.k;
  p.gl = 1;
  while (p.gl < 32 && p.gl * 2 * groups <= kTeamThreads) p.gl *= 2;
  p.per_frame = per_frame ? 1 : 0; p.f = f; p.eps = eps;
  p.dom_rows = pl.dom_rows; p.domains = pl.domains; p.team = pl.team; p.n_teams = pl.n_teams; p.slice_rows = pl.slice_rows;
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + pl.counter_bytes);

  const void* fn = nullptr;
  if (dtype == CA_BF16) fn = apply_silu ? (const void*)gn_team_kernel<__nv_bfloat16, true> : (const void*)gn_team_kernel<__nv_bfloat16, false>;
  else fn = apply_silu ? (const void*)gn_team_kernel<__half, true> : (const void*)gn_team_kernel<__half, false>;
  CA_CUDA(ensure_dynamic_smem(fn, pl.smem));
  const int grid = pl.team * pl.n_teams;
  if (pl.team > 1) {
    int per_sm = 0;
    CA_CUDA(cached_occupancy(&per_sm, fn, kTeamThreads, pl.smem));
    if ((long long)per_sm * sm_count() < grid) return CA_OK;  // cannot be co-resident: use the split kernels
    CA_CUDA(cudaMemsetAsync(p.counters, 0, pl.counter_bytes, st));
  }
  void* args[] = {(void*)&p};
  if (pl.team > 1) {
    CA_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kTeamThreads), args, pl.smem, st));
  } else {
    CA_CUDA(cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(kTeamThreads), args, pl.smem, st));
  }
  *handled = true;
  return CA_OK;
}

}  // namespace ca
