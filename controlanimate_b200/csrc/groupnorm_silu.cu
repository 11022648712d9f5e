// Kernel (2): fused GroupNorm + SiLU (+ time-embedding add) over a video activation [b,c,f,h,w].
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31,
// 191-192, 199-208; unet.py:614-615), which costs 4 full read+write passes there (rearrange copy,
// GroupNorm, rearrange copy, SiLU) plus a separate temb-add pass.  Here: ONE HBM read + ONE HBM
// write (algorithmic bytes 2*N*s).
//
// Design (B200): a "domain" is one set of elements that share statistics.  Each domain is cut
// into chunks of <= ~48 KB that stay RESIDENT IN SHARED MEMORY between the statistics pass and
// the normalise pass, so HBM is touched once.  Chunks of one domain exchange (mean, M2) partials
// through a tiny global workspace and a per-domain arrival counter (Chan's parallel-variance
// combine, in double); CTAs of one domain have consecutive block indices and are therefore
// co-resident (the host checks chunks_per_domain against the resident-CTA capacity and otherwise
// falls back to two launches: statistics, then apply with an L2-assisted re-read).
//   NCFHW : domain = (b, group[, frame]) : cpg rows of (h*w | f*h*w) contiguous elements
//   BFHWC : domain = (b[, frame])        : (h*w | f*h*w) token rows of c contiguous channels, all
//           groups at once so every global access is a full 16-byte vector of a dense row.
#include "common.cuh"

namespace ca {
namespace {

constexpr int kPhaseStats = 1, kPhaseApply = 2, kPhaseFused = 3;

struct GnParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;  // [b, c] or null
  int b, c, f, hw, groups, cpg;
  int per_frame, apply_silu, phase;
  float eps;
  int chunks;          // chunks per domain
  long long chunk_vecs;  // vectors (NCFHW) or rows (BFHWC) per chunk
  double2* partials;   // [domains][chunks][groups_per_domain] (mean, M2)
  unsigned int* counters;  // [domains]
};

__device__ __forceinline__ void domain_barrier(unsigned int* counter, int chunks) {
  // One thread publishes this chunk's partials (written before the call) and waits for its peers.
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
      if (seen < (unsigned)chunks) __nanosleep(64);
    } while (seen < (unsigned)chunks);
  }
  __syncthreads();
}

// ------------------------------------------------------------------------------------------
// NCFHW.  VEC = elements per access (16-byte vectors, or 1 for ragged h*w).
// ------------------------------------------------------------------------------------------
template <typename T, int VEC>
__global__ void __launch_bounds__(512) gn_ncfhw_kernel(const GnParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sdata = reinterpret_cast<T*>(smem_raw);
  __shared__ double red_d[32];
  __shared__ float s_mean, s_rstd;

  const int domain = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
  // domain -> (b, g[, frame])
  int bi, g, fi = 0;
  if (p.per_frame) {
    fi = domain % p.f;
    g = (domain / p.f) % p.groups;
    bi = domain / (p.f * p.groups);
  } else {
    g = domain % p.groups;
    bi = domain / p.groups;
  }
  const long long cols = p.per_frame ? p.hw : (long long)p.f * p.hw;  // elements per channel row
  const long long row_stride = (long long)p.f * p.hw;
  const long long base = ((long long)bi * p.c + (long long)g * p.cpg) * row_stride + (p.per_frame ? (long long)fi * p.hw : 0);
  const long long colv = cols / VEC;
  const long long total = colv * p.cpg;
  const long long v0 = (long long)chunk * p.chunk_vecs;
  const long long v1 = min(total, v0 + p.chunk_vecs);
  const int n = (int)(v1 - v0);
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x);
  T* __restrict__ y = reinterpret_cast<T*>(p.y);
  const float* temb = p.temb ? p.temb + (long long)bi * p.c + g * p.cpg : nullptr;

  // ---- load chunk into smem (+ local sum) ----
  float lsum = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long v = v0 + i;
    const int row = (int)(v / colv);
    const long long col = (v - (long long)row * colv) * VEC;
    const T* src = x + base + row * row_stride + col;
    const float t = temb ? temb[row] : 0.f;
    if constexpr (VEC > 1) {
      Vec16<T> vv;
      vv.raw = ldg_stream(src);
      reinterpret_cast<uint4*>(sdata)[i] = vv.raw;
      float fv[VEC];
      vv.unpack(fv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) lsum += fv[j] + t;
    } else {
      const T e = *src;
      sdata[i] = e;
      lsum += Traits<T>::to_f(e) + t;
    }
  }
  float mean, rstd;
  if (p.phase & kPhaseStats) {
    const double cnt = (double)n * VEC;
    const double bsum = block_sum<double>((double)lsum, red_d);
    const float lmean = (float)(bsum / cnt);
    // centred second moment from smem (exact two-pass inside the chunk)
    float lsq = 0.f;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
      const int row = (int)((v0 + i) / colv);
      const float t = (temb ? temb[row] : 0.f) - lmean;
      if constexpr (VEC > 1) {
        Vec16<T> vv;
        vv.raw = reinterpret_cast<const uint4*>(sdata)[i];
        float fv[VEC];
        vv.unpack(fv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float dlt = fv[j] + t;
          lsq += dlt * dlt;
        }
      } else {
        const float dlt = Traits<T>::to_f(sdata[i]) + t;
        lsq += dlt * dlt;
      }
    }
    const double m2 = block_sum<double>((double)lsq, red_d);
    if (p.chunks == 1 && p.phase == kPhaseFused) {
      mean = lmean;
      rstd = rsqrtf((float)(m2 / cnt) + p.eps);
    } else {
      if (threadIdx.x == 0) p.partials[(long long)domain * p.chunks + chunk] = make_double2((double)lmean, m2);
      if (p.phase == kPhaseStats) return;
      domain_barrier(p.counters + domain, p.chunks);
    }
  }
  if (!(p.chunks == 1 && p.phase == kPhaseFused)) {
    if (threadIdx.x == 0) {
      // Chan combine over the chunks of this domain.
      double ntot = 0, msum = 0;
      for (int k = 0; k < p.chunks; ++k) {
        const long long a0 = (long long)k * p.chunk_vecs;
        const double nk = (double)(min(total, a0 + p.chunk_vecs) - a0) * VEC;
        const double2 pk = __ldcg(p.partials + (long long)domain * p.chunks + k);
        ntot += nk;
        msum += nk * pk.x;
      }
      const double gm = msum / ntot;
      double m2 = 0;
      for (int k = 0; k < p.chunks; ++k) {
        const long long a0 = (long long)k * p.chunk_vecs;
        const double nk = (double)(min(total, a0 + p.chunk_vecs) - a0) * VEC;
        const double2 pk = __ldcg(p.partials + (long long)domain * p.chunks + k);
        m2 += pk.y + nk * (pk.x - gm) * (pk.x - gm);
      }
      s_mean = (float)gm;
      s_rstd = rsqrtf((float)(m2 / ntot) + p.eps);
    }
    __syncthreads();
    mean = s_mean;
    rstd = s_rstd;
  }

  // ---- normalise + affine + SiLU from smem, single global write ----
  const float* gamma = p.gamma + g * p.cpg;
  const float* beta = p.beta + g * p.cpg;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long v = v0 + i;
    const int row = (int)(v / colv);
    const long long col = (v - (long long)row * colv) * VEC;
    const float a = gamma[row] * rstd;
    const float bb = beta[row] + ((temb ? temb[row] : 0.f) - mean) * a;
    T* dst = y + base + row * row_stride + col;
    if constexpr (VEC > 1) {
      Vec16<T> vv;
      vv.raw = reinterpret_cast<const uint4*>(sdata)[i];
      float fv[VEC];
      vv.unpack(fv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float o = fmaf(fv[j], a, bb);
        fv[j] = p.apply_silu ? silu_f(o) : o;
      }
      vv.pack(fv);
      stg_stream(dst, vv.raw);
    } else {
      const float o = fmaf(Traits<T>::to_f(sdata[i]), a, bb);
      *dst = Traits<T>::from_f(p.apply_silu ? silu_f(o) : o);
    }
  }
}

// ------------------------------------------------------------------------------------------
// BFHWC (token-major).  One CTA = `rows` consecutive tokens x all c channels.
// blockDim = roundup32(nvec * k) where nvec = c / VEC; thread t < nvec*k owns channel vector
// t % nvec for rows t / nvec, t / nvec + k, ...
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(1024) gn_bfhwc_kernel(const GnParams p) {
  constexpr int VEC = Traits<T>::kVec;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int C = p.c;
  float* s_a = reinterpret_cast<float*>(smem_raw);  // [C] per-channel sum -> later scale
  float* s_b = s_a + C;                             // [C] per-channel sq  -> later shift
  float* s_gmean = s_b + C;                         // [groups]
  float* s_grstd = s_gmean + p.groups;              // [groups]
  uint4* sdata = reinterpret_cast<uint4*>(s_grstd + p.groups + ((2 * p.groups) % 4 ? 4 - (2 * p.groups) % 4 : 0));

  const int domain = blockIdx.x / p.chunks, chunk = blockIdx.x % p.chunks;
  const int bi = p.per_frame ? domain / p.f : domain;
  const long long dom_rows = p.per_frame ? p.hw : (long long)p.f * p.hw;
  const long long r0 = (long long)chunk * p.chunk_vecs;
  const int rows = (int)(min(dom_rows, r0 + p.chunk_vecs) - r0);
  const int nvec = C / VEC;
  const int k = max(1, 512 / nvec);
  const int active = nvec * k;
  const bool on = threadIdx.x < active;
  const int cv = threadIdx.x % nvec, rl = threadIdx.x / nvec;
  const long long base = ((long long)domain * dom_rows + r0) * C;  // domains are contiguous slabs
  const T* __restrict__ x = reinterpret_cast<const T*>(p.x) + base;
  T* __restrict__ y = reinterpret_cast<T*>(p.y) + base;
  const float* temb = p.temb ? p.temb + (long long)bi * C : nullptr;

  float tv[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) tv[j] = (temb && on) ? temb[cv * VEC + j] : 0.f;

  for (int c0 = threadIdx.x; c0 < 2 * C; c0 += blockDim.x) s_a[c0] = 0.f;  // zero s_a and s_b
  __syncthreads();

  // ---- load + per-channel sums ----
  float acc[VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) acc[j] = 0.f;
  if (on) {
    for (int r = rl; r < rows; r += k) {
      Vec16<T> vv;
      vv.raw = ldg_stream(x + (long long)r * C + cv * VEC);
      sdata[r * nvec + cv] = vv.raw;
      float fv[VEC];
      vv.unpack(fv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) acc[j] += fv[j] + tv[j];
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j) atomicAdd(&s_a[cv * VEC + j], acc[j]);
  }
  __syncthreads();
  const double cnt = (double)rows * p.cpg;
  if (p.phase & kPhaseStats) {
    if (threadIdx.x < p.groups) {
      float s = 0.f;
      for (int j = 0; j < p.cpg; ++j) s += s_a[threadIdx.x * p.cpg + j];
      s_gmean[threadIdx.x] = (float)((double)s / cnt);
    }
    __syncthreads();
    if (on) {
      float gm[VEC];
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        gm[j] = tv[j] - s_gmean[(cv * VEC + j) / p.cpg];
        acc[j] = 0.f;
      }
      for (int r = rl; r < rows; r += k) {
        Vec16<T> vv;
        vv.raw = sdata[r * nvec + cv];
        float fv[VEC];
        vv.unpack(fv);
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float dlt = fv[j] + gm[j];
          acc[j] += dlt * dlt;
        }
      }
#pragma unroll
      for (int j = 0; j < VEC; ++j) atomicAdd(&s_b[cv * VEC + j], acc[j]);
    }
    __syncthreads();
    if (threadIdx.x < p.groups) {
      float s = 0.f;
      for (int j = 0; j < p.cpg; ++j) s += s_b[threadIdx.x * p.cpg + j];
      if (p.chunks == 1 && p.phase == kPhaseFused) {
        s_grstd[threadIdx.x] = rsqrtf((float)((double)s / cnt) + p.eps);
      } else {
        p.partials[((long long)domain * p.chunks + chunk) * p.groups + threadIdx.x] =
            make_double2((double)s_gmean[threadIdx.x], (double)s);
      }
    }
    if (!(p.chunks == 1 && p.phase == kPhaseFused)) {
      if (p.phase == kPhaseStats) return;
      __syncthreads();  // all partial writes of this CTA issued before thread 0 fences
      domain_barrier(p.counters + domain, p.chunks);
    }
  }
  if (!(p.chunks == 1 && p.phase == kPhaseFused)) {
    if (threadIdx.x < p.groups) {
      double ntot = 0, msum = 0;
      for (int q = 0; q < p.chunks; ++q) {
        const long long a0 = (long long)q * p.chunk_vecs;
        const double nk = (double)(min(dom_rows, a0 + p.chunk_vecs) - a0) * p.cpg;
        const double2 pk = __ldcg(p.partials + ((long long)domain * p.chunks + q) * p.groups + threadIdx.x);
        ntot += nk;
        msum += nk * pk.x;
      }
      const double gmn = msum / ntot;
      double m2 = 0;
      for (int q = 0; q < p.chunks; ++q) {
        const long long a0 = (long long)q * p.chunk_vecs;
        const double nk = (double)(min(dom_rows, a0 + p.chunk_vecs) - a0) * p.cpg;
        const double2 pk = __ldcg(p.partials + ((long long)domain * p.chunks + q) * p.groups + threadIdx.x);
        m2 += pk.y + nk * (pk.x - gmn) * (pk.x - gmn);
      }
      s_gmean[threadIdx.x] = (float)gmn;
      s_grstd[threadIdx.x] = rsqrtf((float)(m2 / ntot) + p.eps);
    }
  }
  __syncthreads();
  // per-channel scale/shift
  for (int c0 = threadIdx.x; c0 < C; c0 += blockDim.x) {
    const int g = c0 / p.cpg;
    const float a = p.gamma[c0] * s_grstd[g];
    s_a[c0] = a;
    s_b[c0] = p.beta[c0] + ((temb ? temb[c0] : 0.f) - s_gmean[g]) * a;
  }
  __syncthreads();
  if (on) {
    float av[VEC], bv[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      av[j] = s_a[cv * VEC + j];
      bv[j] = s_b[cv * VEC + j];
    }
    for (int r = rl; r < rows; r += k) {
      Vec16<T> vv;
      vv.raw = sdata[r * nvec + cv];
      float fv[VEC];
      vv.unpack(fv);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        const float o = fmaf(fv[j], av[j], bv[j]);
        fv[j] = p.apply_silu ? silu_f(o) : o;
      }
      vv.pack(fv);
      stg_stream(y + (long long)r * C + cv * VEC, vv.raw);
    }
  }
}

constexpr long long kChunkBytes = 48 * 1024;  // smem-resident chunk target (4 CTAs / SM)

struct GnPlan {
  long long domains;
  int chunks;
  long long chunk_units;  // vectors (NCFHW) or rows (BFHWC)
  int vec;                // NCFHW only
  size_t smem;
  int threads;
  size_t partial_bytes, counter_bytes;
};

int make_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int layout, int dtype, GnPlan* pl) {
  CA_CHECK_ARG(b > 0 && c > 0 && f > 0 && h > 0 && w > 0 && groups > 0, "groupnorm: non-positive dimension");
  CA_CHECK_ARG(c % groups == 0, "groupnorm: c=%d not divisible by groups=%d", c, groups);
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "groupnorm: bad dtype %d", dtype);
  const int esz = dtype == CA_F32 ? 4 : 2;
  const int vec16 = 16 / esz;
  const long long hw = (long long)h * w;
  const int cpg = c / groups;
  if (layout == CA_LAYOUT_NCFHW) {
    const long long cols = per_frame ? hw : (long long)f * hw;
    pl->vec = (cols % vec16 == 0 && (hw % vec16 == 0)) ? vec16 : 1;
    pl->domains = per_frame ? (long long)b * groups * f : (long long)b * groups;
    const long long total = cols / pl->vec * cpg;
    const long long cap = kChunkBytes / (pl->vec * esz);
    pl->chunks = (int)((total + cap - 1) / cap);
    pl->chunk_units = (total + pl->chunks - 1) / pl->chunks;
    pl->smem = (size_t)pl->chunk_units * pl->vec * esz;
    pl->threads = 512;
    pl->partial_bytes = sizeof(double2) * pl->domains * pl->chunks;
  } else if (layout == CA_LAYOUT_BFHWC) {
    CA_CHECK_ARG(c % vec16 == 0, "groupnorm BFHWC: c=%d must be a multiple of %d", c, vec16);
    const int nvec = c / vec16;
    CA_CHECK_ARG(nvec <= 1024 && groups <= 1024, "groupnorm BFHWC: c too large");
    const long long rows = per_frame ? hw : (long long)f * hw;
    pl->vec = vec16;
    pl->domains = per_frame ? (long long)b * f : b;
    long long rcap = kChunkBytes / ((long long)c * esz);
    if (rcap < 1) rcap = 1;
    pl->chunks = (int)((rows + rcap - 1) / rcap);
    pl->chunk_units = (rows + pl->chunks - 1) / pl->chunks;
    const int k = 512 / nvec > 1 ? 512 / nvec : 1;
    pl->threads = ((nvec * k + 31) / 32) * 32;
    if (pl->threads < groups) pl->threads = ((groups + 31) / 32) * 32;
    size_t head = sizeof(float) * (2 * (size_t)c + 2 * (size_t)groups);
    head = (head + 15) / 16 * 16;
    pl->smem = head + (size_t)pl->chunk_units * c * esz;
    pl->partial_bytes = sizeof(double2) * pl->domains * pl->chunks * groups;
  } else {
    set_error("groupnorm: unknown layout %d", layout);
    return CA_ERR_INVALID;
  }
  pl->counter_bytes = ((size_t)pl->domains * sizeof(unsigned int) + 15) / 16 * 16;
  if (pl->chunks == 1) pl->partial_bytes = pl->counter_bytes = 0;
  return CA_OK;
}

template <typename K>
int launch(K kernel, const GnParams& prm, const GnPlan& pl, cudaStream_t st, int capacity_hint) {
  CA_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
  GnParams p = prm;
  const long long grid = pl.domains * pl.chunks;
  CA_CHECK_ARG(grid < (1ll << 31), "groupnorm: grid too large");
  bool fused = true;
  if (pl.chunks > 1) {
    int per_sm = 0;
    CA_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, pl.threads, pl.smem));
    const long long capacity = (long long)per_sm * capacity_hint;
    fused = (long long)pl.chunks * 2 <= capacity;  // all chunks of a domain must be co-resident
  }
  if (fused) {
    if (pl.chunks > 1) CA_CUDA(cudaMemsetAsync(p.counters, 0, pl.counter_bytes, st));
    p.phase = kPhaseFused;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
  } else {
    p.phase = kPhaseStats;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
    p.phase = kPhaseApply;
    kernel<<<(unsigned)grid, pl.threads, pl.smem, st>>>(p);
  }
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) size_t ca_groupnorm_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame,
                                               int layout, int dtype) {
  ca::GnPlan pl;
  if (ca::make_plan(b, c, f, h, w, groups, per_frame, layout, dtype, &pl) != CA_OK) return 0;
  return pl.partial_bytes + pl.counter_bytes;
}

extern "C" __attribute__((visibility("default"))) int ca_groupnorm_silu(const void* x, void* y, const float* gamma, const float* beta, const float* temb,
                                 int b, int c, int f, int h, int w, int groups, float eps, int per_frame,
                                 int apply_silu, int layout, int dtype, void* workspace, size_t workspace_bytes,
                                 void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && gamma && beta, "groupnorm: null pointer");
  GnPlan pl;
  int rc = make_plan(b, c, f, h, w, groups, per_frame, layout, dtype, &pl);
  if (rc != CA_OK) return rc;
  CA_CHECK_ARG(pl.smem <= 200 * 1024, "groupnorm: chunk does not fit shared memory (%zu B)", pl.smem);
  CA_CHECK_ARG(pl.vec == 1 || (aligned16(x) && aligned16(y)), "groupnorm: x/y must be 16-byte aligned");
  const size_t need = pl.partial_bytes + pl.counter_bytes;
  CA_CHECK_ARG(need == 0 || (workspace && workspace_bytes >= need), "groupnorm: workspace too small (%zu < %zu)",
               workspace_bytes, need);
  GnParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb;
  p.b = b; p.c = c; p.f = f; p.hw = h * w; p.groups = groups; p.cpg = c / groups;
  p.per_frame = per_frame ? 1 : 0; p.apply_silu = apply_silu ? 1 : 0; p.eps = eps;
  p.chunks = pl.chunks; p.chunk_vecs = pl.chunk_units;
  p.counters = reinterpret_cast<unsigned int*>(workspace);
  p.partials = reinterpret_cast<double2*>(reinterpret_cast<char*>(workspace) + pl.counter_bytes);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int sms = sm_count();
  return dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if (layout == CA_LAYOUT_NCFHW) {
      if (pl.vec > 1) return launch(gn_ncfhw_kernel<T, Traits<T>::kVec>, p, pl, st, sms);
      return launch(gn_ncfhw_kernel<T, 1>, p, pl, st, sms);
    }
    return launch(gn_bfhwc_kernel<T>, p, pl, st, sms);
  });
}
