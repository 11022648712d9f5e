// Kernel (2): fused GroupNorm + SiLU (+ time-embedding add) over a video activation [b,c,f,h,w].
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31,
// 191-192, 199-208; unet.py:614-615), which costs 4 full read+write passes there (rearrange copy,
// GroupNorm, rearrange copy, SiLU) plus a separate temb-add pass.  Here: ONE HBM read + ONE HBM
// write (algorithmic bytes 2*N*s).
//
// Design (B200): a "domain" is one set of elements that share statistics.  Each domain is cut
// into chunks of <= 32-64 KB; a CTA loads its whole chunk into REGISTERS with all 16-byte loads
// issued up front (maximum memory-level parallelism, nothing staged twice), computes the chunk's
// mean and centred second moment from the registers (exact two-pass), and then normalises
// and stores from the same registers -> HBM is touched exactly once per element.  When a domain
// spans several chunks they exchange (mean, M2) partials through a tiny global workspace and a
// per-domain arrival counter and combine them with Chan's parallel-variance formula in double
// (deterministic: fixed summation order, no floating-point atomics).  CTAs of one domain have
// consecutive block indices and are therefore co-resident; the host checks chunks-per-domain
// against the resident-CTA capacity and otherwise falls back to two launches (statistics, apply).
//   NCFHW : domain = (b, group[, frame]) : cpg rows of (h*w | f*h*w) contiguous elements
//   BFHWC : domain = (b[, frame])        : (h*w | f*h*w) token rows of c contiguous channels; every
//           CTA covers ALL groups of a band of rows so each global access is a dense 16-byte
//           vector of a contiguous row; thread t owns channel vector t % (c/8) for all its rows.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_paths.cuh"

namespace ca {
namespace {

constexpr int kPhaseStats = 1, kPhaseApply = 2, kPhaseFused = 3;
constexpr int kNcfhwThreads = 256;

struct GnParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;  // [b, c] (row stride temb_ld) or null
  long long temb_ld;
  int b, c, f, hw, groups, cpg;
  int ld;               // BFHWC: elements between consecutive token rows (== c unless a channel slab is processed)
  int per_frame, apply_silu, phase;
  float eps;
  int chunks;           // chunks per domain
  int chunk_units;      // vectors (NCFHW) or rows (BFHWC) per chunk
  int k;                // BFHWC: row lanes per CTA
  float2* partials;     // [domains][chunks][groups_per_domain] (mean, M2)
  float2* finals;       // [domains][groups_per_domain] (mean, rstd)
  unsigned int* counters;  // [domains]
};

// Cross-chunk statistics exchange.  Every CTA publishes its (mean, M2) partials, takes a ticket; the CTA that
// draws the last ticket of its domain loads all partials (cooperatively, all loads independent), combines them
// with Chan's formula in double in a FIXED order (deterministic) and publishes (mean, rstd) per group plus a
// ready bit; the others wait for the bit and read the finals.  L2 traffic per CTA stays O(groups), not O(chunks).
constexpr unsigned int kReadyBit = 0x80000000u;

template <int kMaxGroups>
__device__ __forceinline__ void exchange_statistics(const GnParams& p, int domain, int groups_here, long long total_units,
                                                    double per_unit, float* s_mean, float* s_rstd, float2* s_scratch,
                                                    int* s_flag) {
  unsigned int* counter = p.counters + domain;
  float2* part = p.partials + (long long)domain * p.chunks * groups_here;
  float2* fin = p.finals + (long long)domain * groups_here;
  __syncthreads();  // all partial writes of this CTA are issued
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int ticket = atomicAdd(counter, 1u);
    *s_flag = (ticket == (unsigned)p.chunks - 1) ? 1 : 0;
  }
  __syncthreads();
  if (*s_flag) {
    __threadfence();
    const int nvals = p.chunks * groups_here;
    for (int i = threadIdx.x; i < nvals; i += blockDim.x) s_scratch[i] = __ldcg(part + i);
    __syncthreads();
    if (threadIdx.x < groups_here) {
      double ntot = 0, msum = 0;
      for (int q = 0; q < p.chunks; ++q) {
        const long long a0 = (long long)q * p.chunk_units;
        const double nk = (double)(min(total_units, a0 + p.chunk_units) - a0) * per_unit;
        ntot += nk;
        msum += nk * (double)s_scratch[q * groups_here + threadIdx.x].x;
      }
      const double gm = msum / ntot;
      double m2 = 0;
      for (int q = 0; q < p.chunks; ++q) {
        const long long a0 = (long long)q * p.chunk_units;
        const double nk = (double)(min(total_units, a0 + p.chunk_units) - a0) * per_unit;
        const float2 pk = s_scratch[q * groups_here + threadIdx.x];
        const double dm = (double)pk.x - gm;
        m2 += (double)pk.y + nk * dm * dm;
      }
      const float2 out = make_float2((float)gm, rsqrtf((float)(m2 / ntot) + p.eps));
      fin[threadIdx.x] = out;
      s_mean[threadIdx.x] = out.x;
      s_rstd[threadIdx.x] = out.y;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicOr(counter, kReadyBit);
    }
  } else {
    if (threadIdx.x == 0) {
      unsigned int seen;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      } while (!(seen & kReadyBit));
    }
    __syncthreads();
    if (threadIdx.x < groups_here) {
      const float2 v = __ldcg(fin + threadIdx.x);
      s_mean[threadIdx.x] = v.x;
      s_rstd[threadIdx.x] = v.y;
    }
  }
  __syncthreads();
}

// Split mode: statistics launch -> finalize launch (one small CTA per domain) -> apply launch.
__device__ __forceinline__ void read_finals(const GnParams& p, int domain, int groups_here, float* s_mean, float* s_rstd) {
  if (threadIdx.x < groups_here) {
    const float2 v = __ldg(p.finals + (long long)domain * groups_here + threadIdx.x);
    s_mean[threadIdx.x] = v.x;
    s_rstd[threadIdx.x] = v.y;
  }
  __syncthreads();
}

// grid = domains, block = 256 = 8 chunk-lanes x 32 group-lanes; Chan combine in double, fixed order.
__global__ void __launch_bounds__(256) gn_finalize_kernel(const GnParams p, int groups_here, long long total_units, double per_unit) {
  __shared__ double s_n[8][33], s_m[8][33], s_q[8][33];
  const int domain = blockIdx.x;
  const int lane_q = threadIdx.x >> 5, lane_g = threadIdx.x & 31;
  const float2* part = p.partials + (long long)domain * p.chunks * groups_here;
  for (int g0 = 0; g0 < groups_here; g0 += 32) {
    const int g = g0 + lane_g;
    // pass 1: weighted mean
    double ntot = 0, msum = 0;
    if (g < groups_here)
      for (int q = lane_q; q < p.chunks; q += 8) {
        const long long a0 = (long long)q * p.chunk_units;
        const double nk = (double)(min(total_units, a0 + p.chunk_units) - a0) * per_unit;
        ntot += nk;
        msum += nk * (double)__ldcg(part + (long long)q * groups_here + g).x;
      }
    s_n[lane_q][lane_g] = ntot;
    s_m[lane_q][lane_g] = msum;
    __syncthreads();
    double nt = 0, ms = 0;
    for (int l = 0; l < 8; ++l) {
      nt += s_n[l][lane_g];
      ms += s_m[l][lane_g];
    }
    const double gm = nt > 0 ? ms / nt : 0.0;
    // pass 2: M2 around the global mean
    double m2 = 0;
    if (g < groups_here)
      for (int q = lane_q; q < p.chunks; q += 8) {
        const long long a0 = (long long)q * p.chunk_units;
        const double nk = (double)(min(total_units, a0 + p.chunk_units) - a0) * per_unit;
        const float2 pk = __ldcg(part + (long long)q * groups_here + g);
        const double dm = (double)pk.x - gm;
        m2 += (double)pk.y + nk * dm * dm;
      }
    s_q[lane_q][lane_g] = m2;
    __syncthreads();
    if (lane_q == 0 && g < groups_here) {
      double t = 0;
      for (int l = 0; l < 8; ++l) t += s_q[l][lane_g];
      p.finals[(long long)domain * groups_here + g] = make_float2((float)gm, rsqrtf((float)(t / nt) + p.eps));
    }
    __syncthreads();
  }
}

template <typename T, int VEC>
struct Elem {
  using type = uint4;
};
template <typename T>
struct Elem<T, 1> {
  using type = T;
};

// ------------------------------------------------------------------------------------------
// NCFHW.  VEC = elements per access (16-byte vectors, or 1 for ragged h*w); NV accesses per thread.
// ------------------------------------------------------------------------------------------
template <typename T, int VEC, int NV>
__global__ void __launch_bounds__(kNcfhwThreads, 3) gn_ncfhw_kernel(const GnParams p) {
  using E = typename Elem<T, VEC>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float2* s_scratch = reinterpret_cast<float2*>(smem_raw);  // [chunks] partials (finalizing CTA only)
  __shared__ double red_d[32];
  __shared__ float s_mean, s_rstd;
  __shared__ int s_flag;

  const int domain = blockIdx.x / p.chunks, chunk = blockIdx.x - domain * p.chunks;
  int bi, g, fi = 0;
  if (p.per_frame) {
    fi = domain % p.f;
    g = (domain / p.f) % p.groups;
    bi = domain / (p.f * p.groups);
  } else {
    g = domain % p.groups;
    bi = domain / p.groups;
  }
  const int cols = p.per_frame ? p.hw : p.f * p.hw;  // elements per channel row (host guarantees < 2^31)
  const long long row_stride = (long long)p.f * p.hw;
  const long long base = ((long long)bi * p.c + (long long)g * p.cpg) * row_stride + (p.per_frame ? (long long)fi * p.hw : 0);
  const int colv = cols / VEC;
  const int total = colv * p.cpg;  // vectors in the domain
  const int v0 = chunk * p.chunk_units;
  const int n = min(total, v0 + p.chunk_units) - v0;
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x) + base;
  T* __restrict__ y = reinterpret_cast<T*>(p.y) + base;
  const float* temb = p.temb ? p.temb + (long long)bi * p.temb_ld + g * p.cpg : nullptr;

  // ---- all loads up front ----
  E raw[NV];
  int row[NV];  // channel row of access j inside the group, -1 = beyond the chunk
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int i = threadIdx.x + j * kNcfhwThreads;
    row[j] = -1;
    if (i < n) {
      const int v = v0 + i;
      row[j] = v / colv;
      const long long off = row[j] * row_stride + (long long)(v - row[j] * colv) * VEC;
      if constexpr (VEC > 1) raw[j] = ldg_stream(x + off);
      else raw[j] = x[off];
    }
  }
  float tv[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) tv[j] = (temb && row[j] >= 0) ? __ldg(temb + row[j]) : 0.f;

  float mean, rstd;
  if (p.phase & kPhaseStats) {
    float lsum = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (row[j] >= 0) {
        if constexpr (VEC > 1) {
          float fv[VEC];
          Vec16<T> vv;
          vv.raw = raw[j];
          vv.unpack(fv);
#pragma unroll
          for (int e = 0; e < VEC; ++e) lsum += fv[e] + tv[j];
        } else {
          lsum += Traits<T>::to_f(raw[j]) + tv[j];
        }
      }
    }
    const double cnt = (double)n * VEC;
    const float lmean = (float)(block_sum<double>((double)lsum, red_d) / cnt);
    float lsq = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (row[j] >= 0) {
        const float t = tv[j] - lmean;
        if constexpr (VEC > 1) {
          float fv[VEC];
          Vec16<T> vv;
          vv.raw = raw[j];
          vv.unpack(fv);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float dlt = fv[e] + t;
            lsq += dlt * dlt;
          }
        } else {
          const float dlt = Traits<T>::to_f(raw[j]) + t;
          lsq += dlt * dlt;
        }
      }
    }
    const double m2 = block_sum<double>((double)lsq, red_d);
    if (p.chunks == 1 && p.phase == kPhaseFused) {
      mean = lmean;
      rstd = rsqrtf((float)(m2 / cnt) + p.eps);
    } else {
      if (threadIdx.x == 0) p.partials[(long long)domain * p.chunks + chunk] = make_float2(lmean, (float)m2);
      if (p.phase == kPhaseStats) return;
      exchange_statistics<1>(p, domain, 1, total, (double)VEC, &s_mean, &s_rstd, s_scratch, &s_flag);
      mean = s_mean;
      rstd = s_rstd;
    }
  } else {
    read_finals(p, domain, 1, &s_mean, &s_rstd);
    mean = s_mean;
    rstd = s_rstd;
  }

  // ---- normalise + affine + SiLU from registers, single global write ----
  const float* gamma = p.gamma + g * p.cpg;
  const float* beta = p.beta + g * p.cpg;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    if (row[j] >= 0) {
      const float a = __ldg(gamma + row[j]) * rstd;
      const float bb = __ldg(beta + row[j]) + (tv[j] - mean) * a;
      const int v = v0 + threadIdx.x + j * kNcfhwThreads;
      const long long off = row[j] * row_stride + (long long)(v - row[j] * colv) * VEC;
      if constexpr (VEC > 1) {
        float fv[VEC];
        Vec16<T> vv;
        vv.raw = raw[j];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float o = fmaf(fv[e], a, bb);
          fv[e] = p.apply_silu ? silu_f(o) : o;
        }
        vv.pack(fv);
        stg_stream(y + off, vv.raw);
      } else {
        const float o = fmaf(Traits<T>::to_f(raw[j]), a, bb);
        y[off] = Traits<T>::from_f(p.apply_silu ? silu_f(o) : o);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------
// BFHWC (token-major).  One CTA = k*NV consecutive token rows x all c channels.  Thread t < nvec*k owns
// channel vector cv = t % nvec of rows rl + j*k (rl = t / nvec, j < NV): its 8 channels never change, so
// per-channel scale/shift live in registers and per-channel partial sums need one smem exchange.
// ------------------------------------------------------------------------------------------
template <typename T, int NV>
__global__ void __launch_bounds__(320, 2) gn_bfhwc_kernel(const GnParams p) {
  constexpr int VEC = Traits<T>::kVec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = p.c, k = p.k;
  float* s_a = reinterpret_cast<float*>(smem_raw);  // [C] per-channel sum / centred sq
  float* s_gmean = s_a + C;                         // [groups]
  float* s_grstd = s_gmean + p.groups;              // [groups]
  float* s_part = s_grstd + p.groups;               // [k][C] row-lane partials; reused as [chunks][groups] float2 scratch
  __shared__ int s_flag;

  const int domain = blockIdx.x / p.chunks, chunk = blockIdx.x - domain * p.chunks;
  const int bi = p.per_frame ? domain / p.f : domain;
  const int dom_rows = p.per_frame ? p.hw : p.f * p.hw;
  const int r0 = chunk * p.chunk_units;
  const int rows = min(dom_rows, r0 + p.chunk_units) - r0;
  const int nvec = C / VEC;
  const bool on = threadIdx.x < nvec * k;
  const int cv = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const long long ld = p.ld;
  const long long base = ((long long)domain * dom_rows + r0) * ld + cv * VEC;  // domains are contiguous slabs
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x) + base;
  T* __restrict__ y = reinterpret_cast<T*>(p.y) + base;
  const float* temb = p.temb ? p.temb + (long long)bi * p.temb_ld : nullptr;

  uint4 raw[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int r = rl + j * k;
    raw[j] = make_uint4(0u, 0u, 0u, 0u);
    if (on && r < rows) raw[j] = ldg_stream(x + (long long)r * ld);
  }
  float tv[VEC];
#pragma unroll
  for (int e = 0; e < VEC; ++e) tv[e] = (temb && on) ? __ldg(temb + cv * VEC + e) : 0.f;

  const double cnt = (double)rows * p.cpg;
  if (p.phase & kPhaseStats) {
    float acc[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) acc[e] = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      if (on && rl + j * k < rows) {
        float fv[VEC];
        Vec16<T> vv;
        vv.raw = raw[j];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) acc[e] += fv[e] + tv[e];
      }
    }
    if (on) {
#pragma unroll
      for (int e = 0; e < VEC; ++e) s_part[rl * C + cv * VEC + e] = acc[e];
    }
    __syncthreads();
    for (int c0 = threadIdx.x; c0 < C; c0 += blockDim.x) {  // fixed-order (deterministic) reduction over row lanes
      float t = 0.f;
      for (int q = 0; q < k; ++q) t += s_part[q * C + c0];
      s_a[c0] = t;
    }
    __syncthreads();
    if (threadIdx.x < p.groups) {
      float s = 0.f;
      for (int e = 0; e < p.cpg; ++e) s += s_a[threadIdx.x * p.cpg + e];
      s_gmean[threadIdx.x] = (float)((double)s / cnt);
    }
    __syncthreads();
    if (on) {
      float gm[VEC];
#pragma unroll
      for (int e = 0; e < VEC; ++e) {
        gm[e] = tv[e] - s_gmean[(cv * VEC + e) / p.cpg];
        acc[e] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        if (rl + j * k < rows) {
          float fv[VEC];
          Vec16<T> vv;
          vv.raw = raw[j];
          vv.unpack(fv);
#pragma unroll
          for (int e = 0; e < VEC; ++e) {
            const float dlt = fv[e] + gm[e];
            acc[e] += dlt * dlt;
          }
        }
      }
#pragma unroll
      for (int e = 0; e < VEC; ++e) s_part[rl * C + cv * VEC + e] = acc[e];
    }
    __syncthreads();
    for (int c0 = threadIdx.x; c0 < C; c0 += blockDim.x) {
      float t = 0.f;
      for (int q = 0; q < k; ++q) t += s_part[q * C + c0];
      s_a[c0] = t;
    }
    __syncthreads();
    if (threadIdx.x < p.groups) {
      float s = 0.f;
      for (int e = 0; e < p.cpg; ++e) s += s_a[threadIdx.x * p.cpg + e];
      if (p.chunks == 1 && p.phase == kPhaseFused) {
        s_grstd[threadIdx.x] = rsqrtf((float)((double)s / cnt) + p.eps);
      } else {
        p.partials[((long long)domain * p.chunks + chunk) * p.groups + threadIdx.x] = make_float2(s_gmean[threadIdx.x], s);
      }
    }
    if (!(p.chunks == 1 && p.phase == kPhaseFused)) {
      if (p.phase == kPhaseStats) return;
      exchange_statistics<256>(p, domain, p.groups, dom_rows, (double)p.cpg, s_gmean, s_grstd,
                               reinterpret_cast<float2*>(s_part), &s_flag);
    }
  } else {
    read_finals(p, domain, p.groups, s_gmean, s_grstd);
  }
  __syncthreads();
  if (on) {
    float av[VEC], bv[VEC];
#pragma unroll
    for (int e = 0; e < VEC; ++e) {
      const int ch = cv * VEC + e;
      const int g = ch / p.cpg;
      av[e] = __ldg(p.gamma + ch) * s_grstd[g];
      bv[e] = __ldg(p.beta + ch) + (tv[e] - s_gmean[g]) * av[e];
    }
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const int r = rl + j * k;
      if (r < rows) {
        float fv[VEC];
        Vec16<T> vv;
        vv.raw = raw[j];
        vv.unpack(fv);
#pragma unroll
        for (int e = 0; e < VEC; ++e) {
          const float o = fmaf(fv[e], av[e], bv[e]);
          fv[e] = p.apply_silu ? silu_f(o) : o;
        }
        vv.pack(fv);
        stg_stream(y + (long long)r * ld, vv.raw);
      }
    }
  }
}

struct GnPlan {
  long long domains;
  int chunks;
  int chunk_units;  // vectors (NCFHW) or rows (BFHWC)
  int vec;          // elements per access
  int nv;           // accesses per thread (template parameter)
  int k;            // BFHWC row lanes
  int slabs;        // BFHWC: channel slabs (whole groups each) processed by separate launches when a row is wider than a CTA
  size_t smem;
  int threads;
  size_t partial_bytes, counter_bytes, final_bytes;
};

int pow2_at_least(long long v, int cap) {
  int p = 1;
  while (p < v && p < cap) p <<= 1;
  return p;
}

int gn_mode();

int make_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int layout, int dtype, GnPlan* pl) {
  CA_CHECK_ARG(b > 0 && c > 0 && f > 0 && h > 0 && w > 0 && groups > 0, "groupnorm: non-positive dimension");
  CA_CHECK_ARG(c % groups == 0, "groupnorm: c=%d not divisible by groups=%d", c, groups);
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "groupnorm: bad dtype %d", dtype);
  const int esz = dtype == CA_F32 ? 4 : 2;
  const int vec16 = 16 / esz;
  const long long hw = (long long)h * w;
  const int cpg = c / groups;
  constexpr int kMaxNv = 8;
  pl->k = 0;
  pl->smem = 0;
  pl->slabs = 1;
  if (layout == CA_LAYOUT_NCFHW) {
    const long long cols = per_frame ? hw : (long long)f * hw;
    CA_CHECK_ARG(cols * cpg < (1ll << 31), "groupnorm: group too large");
    pl->vec = (hw % vec16 == 0) ? vec16 : 1;
    pl->domains = per_frame ? (long long)b * groups * f : (long long)b * groups;
    const long long total = cols / pl->vec * cpg;
    const long long cap = (long long)kNcfhwThreads * kMaxNv;
    pl->chunks = (int)((total + cap - 1) / cap);
    pl->chunk_units = (int)((total + pl->chunks - 1) / pl->chunks);
    pl->nv = pow2_at_least((pl->chunk_units + kNcfhwThreads - 1) / kNcfhwThreads, kMaxNv);
    pl->threads = kNcfhwThreads;
    pl->smem = pl->chunks > 1 ? sizeof(float2) * (size_t)pl->chunks : 0;
    pl->partial_bytes = sizeof(float2) * pl->domains * pl->chunks;
    pl->final_bytes = sizeof(float2) * pl->domains;
  } else if (layout == CA_LAYOUT_BFHWC) {
    CA_CHECK_ARG(c % vec16 == 0, "groupnorm BFHWC: c=%d must be a multiple of %d", c, vec16);
    // a CTA holds at most 320 channel vectors of a row; wider rows are cut into slabs of whole groups (groups are
    // independent, so each slab is the same problem with fewer channels and the full row stride)
    int slabs = 1;
    while (slabs <= groups && (groups % slabs != 0 || (c / slabs) % vec16 != 0 || c / slabs / vec16 > 320)) ++slabs;
    CA_CHECK_ARG(slabs <= groups && groups / slabs <= 256, "groupnorm BFHWC: c=%d / groups=%d too large", c, groups);
    pl->slabs = slabs;
    c /= slabs;
    groups /= slabs;
    const int nvec = c / vec16;
    const long long rows = per_frame ? hw : (long long)f * hw;
    CA_CHECK_ARG(rows < (1ll << 31), "groupnorm: domain too large");
    pl->vec = vec16;
    pl->domains = per_frame ? (long long)b * f : b;
    const int k = 256 / nvec > 1 ? 256 / nvec : 1;
    pl->k = k;
    const long long cap = (long long)k * kMaxNv;
    pl->chunks = (int)((rows + cap - 1) / cap);
    pl->chunk_units = (int)((rows + pl->chunks - 1) / pl->chunks);
    pl->nv = pow2_at_least((pl->chunk_units + k - 1) / k, kMaxNv);
    pl->threads = ((nvec * k + 31) / 32) * 32;
    if (pl->threads < groups) pl->threads = ((groups + 31) / 32) * 32;
    size_t scratch = sizeof(float) * (size_t)k * c;
    // only the single-launch (fused) mode stages the domain's chunk partials in smem; the split launches never do
    if (pl->chunks > 1 && gn_mode() == 1 && sizeof(float2) * (size_t)pl->chunks * groups > scratch)
      scratch = sizeof(float2) * (size_t)pl->chunks * groups;
    pl->smem = sizeof(float) * ((size_t)c + 2 * (size_t)groups) + scratch;
    pl->partial_bytes = sizeof(float2) * pl->domains * pl->chunks * groups;
    pl->final_bytes = sizeof(float2) * pl->domains * groups;
  } else {
    set_error("groupnorm: unknown layout %d", layout);
    return CA_ERR_INVALID;
  }
  pl->counter_bytes = ((size_t)pl->domains * sizeof(unsigned int) + 15) / 16 * 16;
  if (pl->chunks == 1) pl->partial_bytes = pl->counter_bytes = pl->final_bytes = 0;
  return CA_OK;
}

enum { kModeAuto = 0, kModeFused = 1, kModeSplit = 2 };
int gn_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("CA_GN_MODE");
    mode = !e ? kModeAuto : (e[0] == 'f' ? kModeFused : (e[0] == 's' ? kModeSplit : kModeAuto));
  }
  return mode;
}

template <typename K>
int launch(K kernel, const GnParams& prm, const GnPlan& pl, cudaStream_t st, int groups_here, long long total_units,
           double per_unit) {
  const void* fn = reinterpret_cast<const void*>(kernel);
  if (pl.smem > 48 * 1024) CA_CUDA(ensure_dynamic_smem(fn, pl.smem));
  GnParams p = prm;
  const long long grid = pl.domains * pl.chunks;
  CA_CHECK_ARG(grid < (1ll << 31), "groupnorm: grid too large");
  bool fused = pl.chunks == 1;
  if (pl.chunks > 1 && gn_mode() == kModeFused) {
    int per_sm = 0;
    CA_CUDA(cached_occupancy(&per_sm, fn, pl.threads, pl.smem));
    // every chunk of a domain must be resident at once (with margin for a domain straddling two "waves")
    fused = (long long)pl.chunks * 2 <= (long long)per_sm * sm_count() && pl.smem <= 160 * 1024;
  }
  if (fused) {
    if (pl.chunks > 1) CA_CUDA(cudaMemsetAsync(p.counters, 0, pl.counter_bytes, st));
    p.phase = kPhaseFused;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
  } else {
    p.phase = kPhaseStats;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
    gn_finalize_kernel<<<(unsigned)pl.domains, 256, 0, st>>>(p, groups_here, total_units, per_unit);
    p.phase = kPhaseApply;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
  }
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}

template <typename T>
int launch_ncfhw(const GnParams& p, const GnPlan& pl, cudaStream_t st) {
  constexpr int V = Traits<T>::kVec;
  const long long cols = p.per_frame ? p.hw : (long long)p.f * p.hw;
  const long long tu = cols / pl.vec * p.cpg;
  const double pu = (double)pl.vec;
  auto go = [&](auto kern) { return launch(kern, p, pl, st, 1, tu, pu); };
  if (pl.vec > 1) {
    switch (pl.nv) {
      case 1: return go(gn_ncfhw_kernel<T, V, 1>);
      case 2: return go(gn_ncfhw_kernel<T, V, 2>);
      case 4: return go(gn_ncfhw_kernel<T, V, 4>);
      default: return go(gn_ncfhw_kernel<T, V, 8>);
    }
  }
  switch (pl.nv) {
    case 1: return go(gn_ncfhw_kernel<T, 1, 1>);
    case 2: return go(gn_ncfhw_kernel<T, 1, 2>);
    case 4: return go(gn_ncfhw_kernel<T, 1, 4>);
    default: return go(gn_ncfhw_kernel<T, 1, 8>);
  }
}

template <typename T>
int launch_bfhwc(const GnParams& p, const GnPlan& pl, cudaStream_t st) {
  const long long tu = p.per_frame ? p.hw : (long long)p.f * p.hw;
  const double pu = (double)p.cpg;
  auto go = [&](auto kern) { return launch(kern, p, pl, st, p.groups, tu, pu); };
  switch (pl.nv) {
    case 1: return go(gn_bfhwc_kernel<T, 1>);
    case 2: return go(gn_bfhwc_kernel<T, 2>);
    case 4: return go(gn_bfhwc_kernel<T, 4>);
    default: return go(gn_bfhwc_kernel<T, 8>);
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) size_t ca_groupnorm_workspace_bytes(int b, int c, int f, int h, int w,
                                                                                      int groups, int per_frame, int layout,
                                                                                      int dtype) {
  ca::GnPlan pl;
  if (ca::make_plan(b, c, f, h, w, groups, per_frame, layout, dtype, &pl) != CA_OK) return 0;
  const size_t split = pl.partial_bytes + pl.counter_bytes + pl.final_bytes;
  const size_t ring = layout == CA_LAYOUT_BFHWC ? ca::gn_ring_workspace_bytes(b, c, f, h, w, groups, per_frame, dtype) : 0;
  return split > ring ? split : ring;
}

extern "C" __attribute__((visibility("default"))) int ca_groupnorm_silu(const void* x, void* y, const float* gamma,
                                                                        const float* beta, const float* temb, long long temb_ld,
                                                                        int b, int c, int f, int h, int w, int groups, float eps,
                                                                        int per_frame, int apply_silu, int layout, int dtype,
                                                                        void* workspace, size_t workspace_bytes, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && gamma && beta, "groupnorm: null pointer");
  GnPlan pl;
  int rc = make_plan(b, c, f, h, w, groups, per_frame, layout, dtype, &pl);
  if (rc != CA_OK) return rc;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (layout == CA_LAYOUT_BFHWC) {  // native layout: small domains -> one CTA per (domain, group slab); else the pipelined slice ring
    bool handled = false;
    rc = gn_slab_launch(x, y, gamma, beta, temb, temb_ld > 0 ? temb_ld : c, b, c, f, h, w, groups, eps, per_frame, apply_silu, dtype, st,
                        &handled);
    if (rc != CA_OK || handled) return rc;
    rc = gn_ring_launch(x, y, gamma, beta, temb, temb_ld > 0 ? temb_ld : c, b, c, f, h, w, groups, eps, per_frame, apply_silu, dtype, workspace,
                        workspace_bytes, st, &handled);
    if (rc != CA_OK || handled) return rc;
  }
  CA_CHECK_ARG(pl.smem <= 200 * 1024, "groupnorm: per-CTA scratch does not fit shared memory (%zu B)", pl.smem);
  CA_CHECK_ARG(pl.vec == 1 || (aligned16(x) && aligned16(y)), "groupnorm: x/y must be 16-byte aligned");
  const size_t need = pl.partial_bytes + pl.counter_bytes + pl.final_bytes;
  CA_CHECK_ARG(need == 0 || (workspace && workspace_bytes >= need), "groupnorm: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  GnParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld > 0 ? temb_ld : c;
  p.b = b; p.c = c / pl.slabs; p.f = f; p.hw = h * w; p.groups = groups / pl.slabs; p.cpg = c / groups; p.ld = c;
  p.per_frame = per_frame ? 1 : 0; p.apply_silu = apply_silu ? 1 : 0; p.eps = eps;
  p.chunks = pl.chunks; p.chunk_units = pl.chunk_units; p.k = pl.k;
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + pl.counter_bytes);
  p.finals = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + pl.counter_bytes + pl.partial_bytes);
  return dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if (layout == CA_LAYOUT_NCFHW) return launch_ncfhw<T>(p, pl, st);
    // channel slabs run back to back on the stream, so they can share the statistics workspace
    for (int s = 0; s < pl.slabs; ++s) {
      GnParams q = p;
      const long long off = (long long)s * p.c;
      q.x = reinterpret_cast<const T*>(x) + off;
      q.y = reinterpret_cast<T*>(y) + off;
      q.gamma = gamma + off;
      q.beta = beta + off;
      q.temb = temb ? temb + off : nullptr;
      const int rc2 = launch_bfhwc<T>(q, pl, st);
      if (rc2 != CA_OK) return rc2;
    }
    return CA_OK;
  });
}
