// Kernel (2), native-layout pipelined path: GroupNorm + SiLU (+ time-embedding add) on a BFHWC video
// activation with ONE HBM read and ONE HBM write per element (algorithmic bytes 2*N*s), and with the
// statistics exchange taken off the critical path.
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31, 191-192,
// 199-208; unet.py:614-615) and the per-frame transformer-entry GroupNorm (attention.py:131).
//
// Why a second design next to groupnorm_team.cu: the team kernel runs every CTA through
// load -> statistics -> team barrier -> normalise -> store for one band at a time, so the whole GPU moves
// in lock-step and HBM idles during the statistics and the barrier (measured 28 % of the copy roofline at
// c320 64x64, profiles/r01c_microbench_quick.json).  Here the tensor is cut into small SLICES (k*j rows of
// one statistics domain, 8-16 KB) that are dealt round-robin to a persistent co-resident grid of WORKER
// CTAs.  Every worker keeps a RING of slices in shared memory and is warp-specialised:
//     producer warp     : cp.async.bulk + mbarriers, keeps `stages` slices in flight
//     statistics group  : one pass over the slice in smem (sums shifted by the slice's first row, so the
//                         cancellation of E[x^2]-E[x]^2 never sees the mean), per-group (mean, M2) to a
//                         global table, red.release arrival on the domain's counter; never waits on anyone
//     normalise group   : waits for the domain's flag, normalise + affine + SiLU from smem, 16-byte
//                         streaming stores, releases the stage
// and a few FOLDER CTAs do nothing but wait for complete domains and fold their partials in slice order in
// double (deterministic), publishing (mean, rstd) + flag.  r01d measured why the fold must not sit on a
// worker: with "last arriver folds" the late CTA becomes the last arriver of every following domain and
// the 32 folds serialise (200 us).  Loads, statistics, the cross-CTA exchange and stores of different items
// overlap inside each SM and reads and writes are both in flight all the time.  Every worker owns at most
// one slice per domain (slices_per_domain <= workers) and statistics never wait, so the oldest incomplete
// domain can always complete: no deadlock as long as the grid is co-resident (cooperative launch).
//
// Shapes outside this path (domains larger than the ring can hold: the v1 GroupNorm over (f,h,w); fp32
// storage; c % 8 != 0) fall through to groupnorm_team.cu / the split kernels in groupnorm_silu.cu.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_team.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kGroupThreads = 256;                        // threads of the statistics group and of the normalise group
constexpr int kRingThreads = 2 * kGroupThreads + 32;      // + one producer warp
constexpr int kRingCtasPerSm = 2;
constexpr size_t kRingSmemCap = 110 * 1024;  // dynamic smem per CTA so that two CTAs (+ static + 1 KB reserved) fit one SM
constexpr int kVecE = 8;                     // 16-bit elements per 16-byte vector
constexpr int kMaxStages = 12;
constexpr int kFoldLanes = 16;
constexpr size_t kFoldBytes = sizeof(double) * 4 * kFoldLanes * 33;
constexpr int kMaxFolders = 8;

struct RingParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;  // [b, c] (row stride temb_ld) or null
  long long temb_ld;
  int c, groups, cpg, nvec, k, gl;
  int per_frame, f;
  float eps;
  int dom_rows, domains, slice_rows, spd;  // spd = slices per domain
  int n_items, stages, workers, folders;
  unsigned int stage_bytes, part_floats;
  unsigned long long* partials;  // [domains][spd][groups] packed (mean, M2) of one slice; all-ones = not written yet
  unsigned long long* finals;    // [domains][groups] packed (mean, rstd); all-ones = not written yet
};

// ---- cross-CTA exchange without fences: every datum is one naturally aligned 64-bit word that carries its own
// validity (the host fills both tables with 0xFF bytes before the launch; a real (x, y) pair never has y = 0xFFFFFFFF
// because writers canonicalise that NaN payload).  Relaxed gpu-scope accesses go to L2; single-copy atomicity of the
// 64-bit word is all the protocol needs, so no MEMBAR ever waits for the normalise group's streaming stores.
constexpr unsigned int kNotReady = 0xFFFFFFFFu;

__device__ __forceinline__ unsigned long long ld_relaxed64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long pack_pair(float x, float y) {
  unsigned int yb = __float_as_uint(y);
  if (yb == kNotReady) yb = 0x7FC00000u;  // keep the NaN, drop the reserved payload
  return ((unsigned long long)yb << 32) | (unsigned long long)__float_as_uint(x);
}
__device__ __forceinline__ bool pair_ready(unsigned long long v) { return (unsigned int)(v >> 32) != kNotReady; }
__device__ __forceinline__ float2 unpack_pair(unsigned long long v) {
  return make_float2(__uint_as_float((unsigned int)v), __uint_as_float((unsigned int)(v >> 32)));
}
// Re-poll until the word is valid.  A protocol bug must trap, never hang the GPU.
__device__ __forceinline__ unsigned long long wait_pair(const unsigned long long* p, unsigned long long v, int what) {
  unsigned int spins = 0;
  while (!pair_ready(v)) {
    __nanosleep(32);
    v = ld_relaxed64(p);
    if (++spins == (1u << 24)) {
      printf("controlanimate_b200: groupnorm ring wait %d timed out (block %d thread %d)\n", what, (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
  return v;
}

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

__device__ __forceinline__ void group_bar(int id) { asm volatile("bar.sync %0, %1;" ::"r"(id), "n"(kGroupThreads) : "memory"); }

template <typename T, bool kSilu>
__global__ void __launch_bounds__(kRingThreads, kRingCtasPerSm) gn_ring_kernel(const RingParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t s_full[kMaxStages], s_empty[kMaxStages];

  const int C = p.c, nvec = p.nvec, k = p.k;
  const int tid = threadIdx.x;

  // ===================== folder CTAs: fold the slice partials of finished domains =====================
  if ((int)blockIdx.x >= p.workers) {
    double(*s_fold)[kFoldLanes][33] = reinterpret_cast<double(*)[kFoldLanes][33]>(smem_raw);
    const int lane_q = tid >> 5, lane_g = tid & 31;
    constexpr int kBatch = 4;
    for (int dom = (int)blockIdx.x - p.workers; dom < p.domains; dom += p.folders) {
      const unsigned long long* part = p.partials + (long long)dom * p.spd * p.groups;
      for (int g0 = 0; g0 < p.groups; g0 += 32) {  // slice order, double: N, sum n*m, sum n*m^2, sum M2 (deterministic)
        const int g = g0 + lane_g;
        double a_n = 0, a_m = 0, a_mm = 0, a_q = 0;
        if (lane_q < kFoldLanes && g < p.groups) {
          for (int q0 = lane_q; q0 < p.spd; q0 += kFoldLanes * kBatch) {
            unsigned long long w[kBatch];
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {  // all loads of the batch in flight before the first validity test
              const int qq = q0 + u * kFoldLanes;
              w[u] = qq < p.spd ? ld_relaxed64(part + (long long)qq * p.groups + g) : 0ull;
            }
#pragma unroll
            for (int u = 0; u < kBatch; ++u) {
              const int qq = q0 + u * kFoldLanes;
              if (qq < p.spd) {
                const float2 v = unpack_pair(wait_pair(part + (long long)qq * p.groups + g, w[u], 0));
                const double nk = (double)(min(p.slice_rows, p.dom_rows - qq * p.slice_rows)) * p.cpg;
                const double m = (double)v.x;
                a_n += nk;
                a_m += nk * m;
                a_mm += nk * m * m;
                a_q += (double)v.y;
              }
            }
          }
        }
        if (lane_q < kFoldLanes) {
          s_fold[0][lane_q][lane_g] = a_n;
          s_fold[1][lane_q][lane_g] = a_m;
          s_fold[2][lane_q][lane_g] = a_mm;
          s_fold[3][lane_q][lane_g] = a_q;
        }
        __syncthreads();
        if (lane_q == 0 && g < p.groups) {
          double tn = 0, tm = 0, tmm = 0, tq = 0;
          for (int l = 0; l < kFoldLanes; ++l) {
            tn += s_fold[0][l][lane_g];
            tm += s_fold[1][l][lane_g];
            tmm += s_fold[2][l][lane_g];
            tq += s_fold[3][l][lane_g];
          }
          const double mean = tm / tn;
          double var = (tq + tmm - tn * mean * mean) / tn;
          if (var < 0) var = 0;
          st_relaxed64(p.finals + (long long)dom * p.groups + g, pack_pair((float)mean, rsqrtf((float)var + p.eps)));
        }
        __syncthreads();
      }
    }
    return;
  }

  // ===================== worker CTAs =====================
  const int G = p.workers;
  unsigned char* ring = smem_raw;
  float* s_part = reinterpret_cast<float*>(smem_raw + (size_t)p.stages * p.stage_bytes);  // [k][2][C]
  float* s_ch = s_part + p.part_floats;                                                   // [2][C] (mean_c, M2_c)
  float2* s_fin = reinterpret_cast<float2*>(s_ch + 2 * C);                                // [2][groups] (mean, rstd), double buffer
  const int n_my = (int)blockIdx.x < p.n_items ? (p.n_items - (int)blockIdx.x + G - 1) / G : 0;

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&s_full[s], 1);
      mbar_init(&s_empty[s], 1);
    }
    fence_mbar_init();
  }
  __syncthreads();

  const int role = tid / kGroupThreads;  // 0 statistics, 1 normalise, 2 producer
  const int gt = tid - role * kGroupThreads;
  const bool on = gt < nvec * k;
  const int cv = gt % nvec, rl = gt / nvec;

  if (role == 2) {
    // ---------- producer: keeps the ring full ----------
    if (gt == 0) {
      for (int i = 0; i < n_my; ++i) {
        const int q = blockIdx.x + i * G;
        const int dom = q / p.spd, sl = q - dom * p.spd;
        const int r0 = sl * p.slice_rows;
        const int rows = min(p.slice_rows, p.dom_rows - r0);
        const int st = i % p.stages;
        mbar_wait(&s_empty[st], ((uint32_t)(i / p.stages) & 1u) ^ 1u);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.x) + ((long long)dom * p.dom_rows + r0) * C * (long long)sizeof(T);
        const uint32_t total = (uint32_t)((size_t)rows * C * sizeof(T));
        mbar_arrive_expect_tx(&s_full[st], total);
        constexpr uint32_t kPiece = 8 * 1024;
        for (uint32_t off = 0; off < total; off += kPiece)
          bulk_load_1d(ring + (size_t)st * p.stage_bytes + off, src + off, min(kPiece, total - off), &s_full[st]);
      }
    }
    return;
  }

  if (role == 0) {
    // ---------- statistics group: never waits on another CTA ----------
    for (int i = 0; i < n_my; ++i) {
      const int q = blockIdx.x + i * G;
      const int dom = q / p.spd, sl = q - dom * p.spd;
      const int r0 = sl * p.slice_rows;
      const int rows = min(p.slice_rows, p.dom_rows - r0);
      const int st = i % p.stages;
      const int bi = p.per_frame ? dom / p.f : dom;
      const uint4* bufv = reinterpret_cast<const uint4*>(ring + (size_t)st * p.stage_bytes);
      mbar_wait(&s_full[st], (uint32_t)(i / p.stages) & 1u);

      if (on) {
        float x0[kVecE], s1[kVecE], s2[kVecE];
        {
          Vec16<T> v0;
          v0.raw = bufv[cv];
          v0.unpack(x0);
        }
#pragma unroll
        for (int e = 0; e < kVecE; ++e) s1[e] = s2[e] = 0.f;
#pragma unroll 2
        for (int r = rl; r < rows; r += k) {
          float fv[kVecE];
          Vec16<T> vv;
          vv.raw = bufv[r * nvec + cv];
          vv.unpack(fv);
#pragma unroll
          for (int e = 0; e < kVecE; ++e) {
            const float d = fv[e] - x0[e];
            s1[e] += d;
            s2[e] = fmaf(d, d, s2[e]);
          }
        }
        float4* d1 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 2 + 0) * C + cv * kVecE);
        float4* d2 = reinterpret_cast<float4*>(s_part + ((size_t)rl * 2 + 1) * C + cv * kVecE);
        d1[0] = make_float4(s1[0], s1[1], s1[2], s1[3]);
        d1[1] = make_float4(s1[4], s1[5], s1[6], s1[7]);
        d2[0] = make_float4(s2[0], s2[1], s2[2], s2[3]);
        d2[1] = make_float4(s2[4], s2[5], s2[6], s2[7]);
      }
      group_bar(1);
      // per-channel (mean, M2) of the slice: the k row lanes share the shift, so their sums just add (fixed order)
      {
        const T* row0 = reinterpret_cast<const T*>(bufv);
        const float inv_n = 1.0f / (float)rows;
        for (int c0 = gt; c0 < C; c0 += kGroupThreads) {
          float a1 = 0.f, a2 = 0.f;
#pragma unroll 4
          for (int qq = 0; qq < k; ++qq) {
            a1 += s_part[((size_t)qq * 2 + 0) * C + c0];
            a2 += s_part[((size_t)qq * 2 + 1) * C + c0];
          }
          const float t = p.temb ? __ldg(p.temb + (long long)bi * p.temb_ld + c0) : 0.f;
          const float dm = a1 * inv_n;
          s_ch[c0] = Traits<T>::to_f(row0[c0]) + t + dm;
          s_ch[C + c0] = fmaxf(a2 - a1 * dm, 0.f);
        }
      }
      group_bar(1);
      // per-group (mean, M2) of the slice from its cpg channels (equal counts: Chan's formula with n_c = rows); the owner
      // lane publishes the pair as one 64-bit word
      {
        const int L = p.gl;
        for (int g0 = 0; g0 < p.groups; g0 += kGroupThreads / L) {
          const int g = g0 + gt / L, l = gt % L;
          float sm = 0.f, sq = 0.f;
          if (g < p.groups)
            for (int e = l; e < p.cpg; e += L) {
              sm += s_ch[g * p.cpg + e];
              sq += s_ch[C + g * p.cpg + e];
            }
          for (int o = L >> 1; o > 0; o >>= 1) {
            sm += __shfl_xor_sync(0xffffffffu, sm, o);
            sq += __shfl_xor_sync(0xffffffffu, sq, o);
          }
          const float gmean = sm / (float)p.cpg;
          float dv = 0.f;
          if (g < p.groups)
            for (int e = l; e < p.cpg; e += L) {
              const float d = s_ch[g * p.cpg + e] - gmean;
              dv = fmaf(d, d, dv);
            }
          for (int o = L >> 1; o > 0; o >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, o);
          if (l == 0 && g < p.groups)
            st_relaxed64(p.partials + ((long long)dom * p.spd + sl) * p.groups + g, pack_pair(gmean, fmaf((float)rows, dv, sq)));
        }
      }
      // no trailing barrier: s_part is rewritten only after every thread passed the second barrier above, s_ch only after
      // the first barrier of the next item, which every thread reaches after its reads here
    }
    return;
  }

  // ---------- normalise group ----------
  // channel e of this thread's vector belongs to group g_first + nibble e of gmap (8 channels span at most 8 groups)
  const int g_first = (cv * kVecE) / p.cpg;
  unsigned int gmap = 0;
#pragma unroll
  for (int e = 0; e < kVecE; ++e) gmap |= (unsigned int)((cv * kVecE + e) / p.cpg - g_first) << (4 * e);

  auto dom_of = [&](int i) { return (int)((blockIdx.x + (long long)i * G) / p.spd); };
  // (mean, rstd) of item 0's domain -> s_fin[0]
  if (n_my > 0) {
    for (int g = gt; g < p.groups; g += kGroupThreads) {
      const unsigned long long* src = p.finals + (long long)dom_of(0) * p.groups + g;
      s_fin[g] = unpack_pair(wait_pair(src, ld_relaxed64(src), 1));
    }
    group_bar(2);
  }
  for (int i = 0; i < n_my; ++i) {
    const int q = blockIdx.x + i * G;
    const int dom = q / p.spd, sl = q - dom * p.spd;
    const int r0 = sl * p.slice_rows;
    const int rows = min(p.slice_rows, p.dom_rows - r0);
    const int st = i % p.stages;
    const int bi = p.per_frame ? dom / p.f : dom;
    const uint4* bufv = reinterpret_cast<const uint4*>(ring + (size_t)st * p.stage_bytes);
    const float2* fin = s_fin + (i & 1) * p.groups;

    // the next item's statistics are requested now and tested after this item's stores (the L2 round trip is hidden)
    const bool fetch_next = i + 1 < n_my && gt < p.groups;  // groups <= kGroupThreads is checked on the host
    const unsigned long long* nsrc = nullptr;
    unsigned long long nval = 0;
    if (fetch_next) {
      nsrc = p.finals + (long long)dom_of(i + 1) * p.groups + gt;
      nval = ld_relaxed64(nsrc);
    }

    mbar_wait(&s_full[st], (uint32_t)(i / p.stages) & 1u);  // completed long ago (the statistics group consumed it): visibility only
    if (on) {
      float av[kVecE], bv[kVecE];
      {
        const float4* g4 = reinterpret_cast<const float4*>(p.gamma + cv * kVecE);
        const float4* b4 = reinterpret_cast<const float4*>(p.beta + cv * kVecE);
        const float4 ga = __ldg(g4), gb = __ldg(g4 + 1), ba = __ldg(b4), bb = __ldg(b4 + 1);
        const float gam[kVecE] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
        const float bet[kVecE] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
        float tv[kVecE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        if (p.temb) {
          const float* tp = p.temb + (long long)bi * p.temb_ld + cv * kVecE;  // row stride may be unaligned: scalar loads
#pragma unroll
          for (int e = 0; e < kVecE; ++e) tv[e] = __ldg(tp + e);
        }
#pragma unroll
        for (int e = 0; e < kVecE; ++e) {
          const float2 mr = fin[g_first + (int)((gmap >> (4 * e)) & 7u)];
          av[e] = gam[e] * mr.y;
          bv[e] = fmaf(tv[e] - mr.x, av[e], bet[e]);
          if constexpr (kSilu) {
            av[e] *= 0.5f;
            bv[e] *= 0.5f;
          }
        }
      }
      T* yg = reinterpret_cast<T*>(p.y) + ((long long)dom * p.dom_rows + r0) * C + cv * kVecE;
#pragma unroll 2
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
    if (fetch_next) s_fin[((i + 1) & 1) * p.groups + gt] = unpack_pair(wait_pair(nsrc, nval, 1));
    group_bar(2);  // the stage is free again and the next item's (mean, rstd) are in place
    if (gt == 0) mbar_arrive(&s_empty[st]);
  }
}

struct RingPlan {
  int domains, dom_rows, nvec, k, slice_rows, spd, n_items, stages, workers, folders;
  unsigned int stage_bytes, part_floats;
  size_t smem, counter_bytes, partial_bytes, final_bytes;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

// Returns false when the shape is outside the pipelined path.
bool make_ring_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype, RingPlan* pl) {
  static const int on = env_int("CA_GN_RING", 1);
  static const int target_kb = env_int("CA_GN_RING_KB", 16);
  static const int want_stages = env_int("CA_GN_RING_STAGES", 6);
  static const int want_folders = env_int("CA_GN_RING_FOLDERS", kMaxFolders);
  if (!on) return false;
  if (dtype != CA_BF16 && dtype != CA_F16) return false;
  if (c % kVecE != 0 || groups <= 0 || c % groups != 0 || groups > kGroupThreads) return false;
  const int nvec = c / kVecE;
  if (nvec > kGroupThreads) return false;
  const long long rows = per_frame ? (long long)h * w : (long long)f * h * w;
  const long long domains = per_frame ? (long long)b * f : b;
  if (rows <= 0 || rows >= (1ll << 30) || domains <= 0 || domains >= (1ll << 24)) return false;
  const int k = kGroupThreads / nvec;
  const size_t part_floats = (size_t)k * 2 * c;  // row-lane partial sums
  const size_t scratch = sizeof(float) * (part_floats + 2 * (size_t)c) + sizeof(float2) * 2 * (size_t)groups;
  if (scratch >= kRingSmemCap) return false;
  const long long row_bytes = (long long)c * 2;
  int folders = want_folders < 1 ? 1 : (want_folders > kMaxFolders ? kMaxFolders : want_folders);
  if (folders > domains) folders = (int)domains;
  const int G = kRingCtasPerSm * sm_count() - folders;  // worker CTAs

  long long j = ((long long)target_kb * 1024) / ((long long)k * row_bytes);
  if (j < 1) j = 1;
  if ((long long)k * j > rows) j = (rows + k - 1) / k;
  while ((rows + k * j - 1) / (k * j) > G) ++j;  // every worker owns at most one slice per domain
  long long stages;
  for (;; --j) {
    stages = (long long)(kRingSmemCap - scratch) / ((long long)k * j * row_bytes);
    if (stages >= 3 || j == 1) break;
  }
  if (stages < 3) return false;
  const long long slice_rows = (long long)k * j;
  const long long spd = (rows + slice_rows - 1) / slice_rows;
  if (spd > G) return false;
  if (stages > want_stages) stages = want_stages;
  if (stages > kMaxStages) stages = kMaxStages;
  const long long n_items = domains * spd;
  if (n_items >= (1ll << 30)) return false;

  pl->domains = (int)domains;
  pl->dom_rows = (int)rows;
  pl->nvec = nvec;
  pl->k = k;
  pl->slice_rows = (int)slice_rows;
  pl->spd = (int)spd;
  pl->n_items = (int)n_items;
  pl->stages = (int)stages;
  pl->workers = (int)(n_items < G ? n_items : G);
  pl->folders = folders;
  pl->stage_bytes = (unsigned int)(slice_rows * row_bytes);
  pl->part_floats = (unsigned int)part_floats;
  pl->smem = (size_t)stages * pl->stage_bytes + scratch;
  if (pl->smem < kFoldBytes) pl->smem = kFoldBytes;  // folder CTAs use the start of the dynamic smem as fold scratch
  pl->counter_bytes = 0;
  pl->partial_bytes = sizeof(float2) * (size_t)domains * spd * groups;
  pl->final_bytes = sizeof(float2) * (size_t)domains * groups;
  return true;
}

}  // namespace

size_t gn_ring_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype) {
  RingPlan pl;
  if (!make_ring_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return 0;
  return pl.counter_bytes + pl.partial_bytes + pl.final_bytes;
}

int gn_ring_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  RingPlan pl;
  if (!make_ring_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return CA_OK;
  if (!aligned16(x) || !aligned16(y)) return CA_OK;
  const size_t need = pl.counter_bytes + pl.partial_bytes + pl.final_bytes;
  if (!workspace || workspace_bytes < need || !aligned16(workspace)) return CA_OK;

  RingParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld;
  p.c = c; p.groups = groups; p.cpg = c / groups; p.nvec = pl.nvec; p.k = pl.k;
  p.gl = 1;
  while (p.gl < 32 && p.gl * 2 <= p.cpg && p.gl * 2 * groups <= kGroupThreads) p.gl *= 2;
  p.per_frame = per_frame ? 1 : 0; p.f = f; p.eps = eps;
  p.dom_rows = pl.dom_rows; p.domains = pl.domains; p.slice_rows = pl.slice_rows; p.spd = pl.spd;
  p.n_items = pl.n_items; p.stages = pl.stages; p.workers = pl.workers; p.folders = pl.folders;
  p.stage_bytes = pl.stage_bytes; p.part_floats = pl.part_floats;
  char* ws = reinterpret_cast<char*>(workspace);
  p.partials = reinterpret_cast<unsigned long long*>(ws);
  p.finals = reinterpret_cast<unsigned long long*>(ws + pl.partial_bytes);

  const void* fn = nullptr;
  if (dtype == CA_BF16) fn = apply_silu ? (const void*)gn_ring_kernel<__nv_bfloat16, true> : (const void*)gn_ring_kernel<__nv_bfloat16, false>;
  else fn = apply_silu ? (const void*)gn_ring_kernel<__half, true> : (const void*)gn_ring_kernel<__half, false>;
  CA_CUDA(ensure_dynamic_smem(fn, pl.smem));
  int per_sm = 0;
  CA_CUDA(cached_occupancy(&per_sm, fn, kRingThreads, pl.smem));
  const int grid = pl.workers + pl.folders;
  if ((long long)per_sm * sm_count() < grid) return CA_OK;  // cannot be co-resident: use the other kernels
  CA_CUDA(cudaMemsetAsync(workspace, 0xFF, need, st));  // every exchange word starts as 'not written yet'
  void* args[] = {(void*)&p};
  CA_CUDA(cudaLaunchCooperativeKernel(fn, dim3((unsigned)grid), dim3(kRingThreads), args, pl.smem, st));
  *handled = true;
  return CA_OK;
}

}  // namespace ca
