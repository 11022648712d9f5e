// Kernel (2), native-layout pipelined path: GroupNorm + SiLU (+ time-embedding add) on a BFHWC video
// activation with ONE HBM read and ONE HBM write per element (algorithmic bytes 2*N*s), and with the
// statistics exchange taken off the critical path.
//
// Replaces InflatedGroupNorm.forward + F.silu (reference animatediff/models/resnet.py:23-31, 191-192,
// 199-208; unet.py:614-615) and the per-frame transformer-entry GroupNorm (attention.py:131).
//
// Why this design (round 1's persistent "team" kernel, since removed, ran every CTA through
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
// storage; c % 8 != 0) fall through to the split kernels in groupnorm_silu.cu.
#include <stdlib.h>

#include "common.cuh"
#include "groupnorm_paths.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kGroupThreads = 128;                        // threads of the statistics group and of the normalise group
constexpr int kRingThreads = 2 * kGroupThreads + 32;      // + one producer warp
constexpr int kRingCtasPerSm = 2;            // r01d sweep: 2 CTAs/SM with 32 KB slices beat 3 CTAs/SM with 16 KB slices
constexpr size_t kRingSmemTotal = 222 * 1024;  // per SM, shared by the resident CTAs (+ static + 1 KB reserved each)
constexpr int kVecE = 8;                     // 16-bit elements per 16-byte vector
constexpr int kMaxStages = 12;
constexpr int kFoldLanes = 8;
constexpr size_t kFoldBytes = sizeof(double) * 4 * kFoldLanes * 33;
constexpr int kMaxFolders = 8;

struct RingParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* temb;  // [b, c] (row stride temb_ld) or null
  long long temb_ld;
  int c, groups, cpg, nvec, k, gl;         // c = full row width (row pitch); nvec, k, gl describe one channel slab
  int slabs, cs, gs;                       // channel slabs per row (whole groups each), channels / groups per slab
  int per_frame, f;
  float eps;
  int dom_rows, domains, slice_rows, spd;  // spd = row slices per domain
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

// mbarrier wait that backs off between probes: the three roles of a CTA share the SM's issue slots, and r01d's capture
// showed a fifth of all issued instructions were try_wait spins
__device__ __forceinline__ void mbar_wait_backoff(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    __nanosleep(64);
    if (++spins == (1u << 24)) {
      printf("controlanimate_b200: groupnorm ring mbarrier wait timed out (block %d thread %d)\n", (int)blockIdx.x, (int)threadIdx.x);
      __trap();
    }
  }
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
    constexpr int kBatch = 8;
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
  const int Cs = p.cs;  // channels of one slab: the unit a worker handles; rows keep the full pitch C
  unsigned char* ring = smem_raw;
  float* s_part = reinterpret_cast<float*>(smem_raw + (size_t)p.stages * p.stage_bytes);  // [k][2][Cs]
  float* s_ch = s_part + p.part_floats;                                                   // [2][Cs] (mean_c, M2_c)
  float2* s_fin = reinterpret_cast<float2*>(s_ch + 2 * Cs);                               // [2][gs] (mean, rstd), double buffer
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

  // item i of this CTA is q = blockIdx.x + i*G = (dom*spd + sl)*slabs + slab; (dom, sl, slab), the ring stage and its
  // phase are advanced incrementally: r01d showed the per-item integer divisions alone cost more issue slots than the
  // arithmetic
  int slab = (int)blockIdx.x % p.slabs, sl = ((int)blockIdx.x / p.slabs) % p.spd, dom = ((int)blockIdx.x / p.slabs) / p.spd;
  const int d_slab = G % p.slabs, d_sl = (G / p.slabs) % p.spd, d_dom = (G / p.slabs) / p.spd;
  int st = 0;
  uint32_t phase = 0;
  auto advance = [&]() {
    slab += d_slab;
    sl += d_sl;
    dom += d_dom;
    if (slab >= p.slabs) {
      slab -= p.slabs;
      ++sl;
    }
    if (sl >= p.spd) {
      sl -= p.spd;
      ++dom;
    }
    if (++st == p.stages) {
      st = 0;
      phase ^= 1u;
    }
  };

  if (role == 2) {
    // ---------- producer: keeps the ring full ----------
    if (gt == 0) {
      for (int i = 0; i < n_my; ++i, advance()) {
        const int r0 = sl * p.slice_rows;
        const int rows = min(p.slice_rows, p.dom_rows - r0);
        mbar_wait_backoff(&s_empty[st], phase ^ 1u);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.x) +
                                   (((long long)dom * p.dom_rows + r0) * C + (long long)slab * Cs) * (long long)sizeof(T);
        const uint32_t row_b = (uint32_t)(Cs * sizeof(T));
        const uint32_t total = (uint32_t)rows * row_b;
        unsigned char* dst = ring + (size_t)st * p.stage_bytes;
        mbar_arrive_expect_tx(&s_full[st], total);
        if (p.slabs == 1) {  // whole rows: the slice is one contiguous span
          constexpr uint32_t kPiece = 8 * 1024;
          for (uint32_t off = 0; off < total; off += kPiece) bulk_load_1d(dst + off, src + off, min(kPiece, total - off), &s_full[st]);
        } else {             // a channel slab of every row
          for (int r = 0; r < rows; ++r) bulk_load_1d(dst + (size_t)r * row_b, src + (size_t)r * C * sizeof(T), row_b, &s_full[st]);
        }
      }
    }
    return;
  }

  if (role == 0) {
    // ---------- statistics group: never waits on another CTA ----------
    for (int i = 0; i < n_my; ++i, advance()) {
      const int r0 = sl * p.slice_rows;
      const int rows = min(p.slice_rows, p.dom_rows - r0);
      const int bi = p.per_frame ? dom / p.f : dom;
      const uint4* bufv = reinterpret_cast<const uint4*>(ring + (size_t)st * p.stage_bytes);
      mbar_wait_backoff(&s_full[st], phase);

      if (on) {
        // sums of (x - x0) and (x - x0)^2 per channel, x0 = the slice's first row; packed f32x2 arithmetic
        float2 nx0[kVecE / 2], s1[kVecE / 2], s2[kVecE / 2];
        {
          const uint4 r0v = bufv[cv];
          unpack2(r0v.x, nx0[0].x, nx0[0].y, T());
          unpack2(r0v.y, nx0[1].x, nx0[1].y, T());
          unpack2(r0v.z, nx0[2].x, nx0[2].y, T());
          unpack2(r0v.w, nx0[3].x, nx0[3].y, T());
        }
#pragma unroll
        for (int e = 0; e < kVecE / 2; ++e) {
          nx0[e] = make_float2(-nx0[e].x, -nx0[e].y);
          s1[e] = s2[e] = make_float2(0.f, 0.f);
        }
#pragma unroll 4
        for (int r = rl; r < rows; r += k) {
          const uint4 raw = bufv[r * nvec + cv];
          float2 v[kVecE / 2];
          unpack2(raw.x, v[0].x, v[0].y, T());
          unpack2(raw.y, v[1].x, v[1].y, T());
          unpack2(raw.z, v[2].x, v[2].y, T());
          unpack2(raw.w, v[3].x, v[3].y, T());
#pragma unroll
          for (int e = 0; e < kVecE / 2; ++e) {
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
      group_bar(1);
      // per-channel (mean, M2) of the slice: the k row lanes share the shift, so their sums just add (fixed order)
      {
        const T* row0 = reinterpret_cast<const T*>(bufv);
        const float inv_n = 1.0f / (float)rows;
        const float* tp = p.temb ? p.temb + (long long)bi * p.temb_ld + (long long)slab * Cs : nullptr;
        for (int c0 = gt; c0 < Cs; c0 += kGroupThreads) {
          float a1 = 0.f, a2 = 0.f;
          for (int qq = 0; qq < k; ++qq) {
            a1 += s_part[((size_t)qq * 2 + 0) * Cs + c0];
            a2 += s_part[((size_t)qq * 2 + 1) * Cs + c0];
          }
          const float t = tp ? __ldg(tp + c0) : 0.f;
          const float dm = a1 * inv_n;
          s_ch[c0] = Traits<T>::to_f(row0[c0]) + t + dm;
          s_ch[Cs + c0] = fmaxf(a2 - a1 * dm, 0.f);
        }
      }
      group_bar(1);
      // per-group (mean, M2) of the slice from its cpg channels (equal counts: Chan's formula with n_c = rows); the owner
      // lane publishes the pair as one 64-bit word
      {
        const int L = p.gl;
        const float inv_cpg = 1.0f / (float)p.cpg;
        for (int g0 = 0; g0 < p.gs; g0 += kGroupThreads / L) {
          const int g = g0 + gt / L, l = gt % L;
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
            st_relaxed64(p.partials + ((long long)dom * p.spd + sl) * p.groups + slab * p.gs + g, pack_pair(gmean, fmaf((float)rows, dv, sq)));
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

  // (mean, rstd) of item 0's domain -> s_fin[0]
  if (n_my > 0) {
    for (int g = gt; g < p.gs; g += kGroupThreads) {
      const unsigned long long* src = p.finals + (long long)dom * p.groups + slab * p.gs + g;
      s_fin[g] = unpack_pair(wait_pair(src, ld_relaxed64(src), 1));
    }
    group_bar(2);
  }
  for (int i = 0; i < n_my; ++i) {
    const int r0 = sl * p.slice_rows;
    const int rows = min(p.slice_rows, p.dom_rows - r0);
    const int bi = p.per_frame ? dom / p.f : dom;
    const uint4* bufv = reinterpret_cast<const uint4*>(ring + (size_t)st * p.stage_bytes);
    const float2* fin = s_fin + (i & 1) * p.gs;
    T* yg = reinterpret_cast<T*>(p.y) + ((long long)dom * p.dom_rows + r0) * C + (long long)slab * Cs + cv * kVecE;
    const int my_slab = slab;
    uint64_t* full = &s_full[st];
    uint64_t* empty = &s_empty[st];
    const uint32_t full_phase = phase;
    advance();  // (dom, sl, st, phase) now describe item i + 1

    // the next item's statistics are requested now and tested after this item's stores (the L2 round trip is hidden)
    const unsigned long long* nsrc[2] = {nullptr, nullptr};
    unsigned long long nval[2] = {0ull, 0ull};
    if (i + 1 < n_my) {
#pragma unroll
      for (int u = 0; u < 2; ++u) {  // groups per slab <= 2 * kGroupThreads is checked on the host
        const int g = gt + u * kGroupThreads;
        if (g < p.gs) {
          nsrc[u] = p.finals + (long long)dom * p.groups + slab * p.gs + g;
          nval[u] = ld_relaxed64(nsrc[u]);
        }
      }
    }

    mbar_wait(full, full_phase);  // completed long ago (the statistics group consumed it): visibility only
    if (on) {
      // gamma / beta / temb of this thread's 8 channels: L1-resident after the first item (registers are too scarce at
      // three CTAs per SM to keep them across items)
      const long long co = (long long)my_slab * Cs + cv * kVecE;
      const float4 ga = __ldg(reinterpret_cast<const float4*>(p.gamma + co)), gb = __ldg(reinterpret_cast<const float4*>(p.gamma + co) + 1);
      const float4 ba = __ldg(reinterpret_cast<const float4*>(p.beta + co)), bb = __ldg(reinterpret_cast<const float4*>(p.beta + co) + 1);
      const float gam[kVecE] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w};
      const float bet[kVecE] = {ba.x, ba.y, ba.z, ba.w, bb.x, bb.y, bb.z, bb.w};
      float tv[kVecE] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (p.temb) {
        const float* tp = p.temb + (long long)bi * p.temb_ld + co;  // row stride may be unaligned: scalar loads
#pragma unroll
        for (int e = 0; e < kVecE; ++e) tv[e] = __ldg(tp + e);
      }
      float2 av[kVecE / 2], bv[kVecE / 2];
#pragma unroll
      for (int e = 0; e < kVecE; e += 2) {
        const float2 m0 = fin[g_first + (int)((gmap >> (4 * e)) & 7u)];
        const float2 m1 = fin[g_first + (int)((gmap >> (4 * e + 4)) & 7u)];
        float a0 = gam[e] * m0.y, a1 = gam[e + 1] * m1.y;
        float b0 = fmaf(tv[e] - m0.x, a0, bet[e]), b1 = fmaf(tv[e + 1] - m1.x, a1, bet[e + 1]);
        if constexpr (kSilu) {
          a0 *= 0.5f;
          a1 *= 0.5f;
          b0 *= 0.5f;
          b1 *= 0.5f;
        }
        av[e / 2] = make_float2(a0, a1);
        bv[e / 2] = make_float2(b0, b1);
      }
#pragma unroll 4
      for (int r = rl; r < rows; r += k) {
        const uint4 raw = bufv[r * nvec + cv];
        float2 v[kVecE / 2];
        unpack2(raw.x, v[0].x, v[0].y, T());
        unpack2(raw.y, v[1].x, v[1].y, T());
        unpack2(raw.z, v[2].x, v[2].y, T());
        unpack2(raw.w, v[3].x, v[3].y, T());
#pragma unroll
        for (int e = 0; e < kVecE / 2; ++e) {
          const float2 hh = __ffma2_rn(v[e], av[e], bv[e]);
          if constexpr (kSilu) v[e] = __ffma2_rn(hh, make_float2(tanh_fast(hh.x), tanh_fast(hh.y)), hh);
          else v[e] = hh;
        }
        uint4 out;
        out.x = pack2(v[0].x, v[0].y, T());
        out.y = pack2(v[1].x, v[1].y, T());
        out.z = pack2(v[2].x, v[2].y, T());
        out.w = pack2(v[3].x, v[3].y, T());
        stg_stream(yg + (long long)r * C, out);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (nsrc[u]) s_fin[((i + 1) & 1) * p.gs + gt + u * kGroupThreads] = unpack_pair(wait_pair(nsrc[u], nval[u], 1));
    group_bar(2);  // the stage is free again and the next item's (mean, rstd) are in place
    if (gt == 0) mbar_arrive(empty);
  }
}

struct RingPlan {
  int domains, dom_rows, slabs, cs, gs, nvec, k, slice_rows, spd, n_items, stages, workers, folders;
  unsigned int stage_bytes, part_floats;
  size_t smem, partial_bytes, final_bytes;
};

int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return (e && e[0]) ? atoi(e) : dflt;
}

// Returns false when the shape is outside the pipelined path.
bool make_ring_plan(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype, RingPlan* pl) {
  static const int on = env_int("CA_GN_RING", 1);
  static const int target_kb = env_int("CA_GN_RING_KB", 32);
  static const int want_stages = env_int("CA_GN_RING_STAGES", 6);
  static const int want_folders = env_int("CA_GN_RING_FOLDERS", kMaxFolders);
  static const int want_ctas = env_int("CA_GN_RING_CTAS", kRingCtasPerSm);
  if (!on) return false;
  if (dtype != CA_BF16 && dtype != CA_F16) return false;
  if (c % kVecE != 0 || groups <= 0 || c % groups != 0) return false;
  // a worker handles at most kGroupThreads channel vectors of a row; wider rows are cut into slabs of whole groups
  int slabs = 1;
  while (slabs <= groups && (groups % slabs != 0 || (c / slabs) % kVecE != 0 || c / slabs / kVecE > kGroupThreads)) ++slabs;
  if (slabs > groups) return false;
  const int cs = c / slabs, gs = groups / slabs;
  if (gs > 2 * kGroupThreads) return false;
  const int nvec = cs / kVecE;
  const long long rows = per_frame ? (long long)h * w : (long long)f * h * w;
  const long long domains = per_frame ? (long long)b * f : b;
  if (rows <= 0 || rows >= (1ll << 30) || domains <= 0 || domains >= (1ll << 24)) return false;
  const int k = kGroupThreads / nvec;
  const size_t part_floats = (size_t)k * 2 * cs;  // row-lane partial sums
  const size_t scratch = sizeof(float) * (part_floats + 2 * (size_t)cs) + sizeof(float2) * 2 * (size_t)gs;
  const int ctas = want_ctas < 1 ? 1 : (want_ctas > kRingCtasPerSm ? kRingCtasPerSm : want_ctas);
  const size_t cap = kRingSmemTotal / ctas - 1536;  // static smem + the 1 KB the driver reserves per CTA
  if (scratch >= cap) return false;
  const long long row_bytes = (long long)cs * 2;
  int folders = want_folders < 1 ? 1 : (want_folders > kMaxFolders ? kMaxFolders : want_folders);
  if (folders > domains) folders = (int)domains;
  const int G = ctas * sm_count() - folders;  // worker CTAs

  long long j = ((long long)target_kb * 1024) / ((long long)k * row_bytes);
  if (j < 1) j = 1;
  if ((long long)k * j > rows) j = (rows + k - 1) / k;
  while (((rows + k * j - 1) / (k * j)) * slabs > G) ++j;  // every worker owns at most one item per domain
  long long stages;
  for (;; --j) {
    stages = (long long)(cap - scratch) / ((long long)k * j * row_bytes);
    if (stages >= 3 || j == 1) break;
  }
  if (stages < 3) return false;
  const long long slice_rows = (long long)k * j;
  const long long spd = (rows + slice_rows - 1) / slice_rows;
  if (spd * slabs > G) return false;
  if (stages > want_stages) stages = want_stages;
  if (stages > kMaxStages) stages = kMaxStages;
  const long long n_items = domains * spd * slabs;
  if (n_items >= (1ll << 30)) return false;

  pl->domains = (int)domains;
  pl->dom_rows = (int)rows;
  pl->slabs = slabs;
  pl->cs = cs;
  pl->gs = gs;
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
  pl->partial_bytes = sizeof(unsigned long long) * (size_t)domains * spd * groups;
  pl->final_bytes = sizeof(unsigned long long) * (size_t)domains * groups;
  return true;
}

}  // namespace

size_t gn_ring_workspace_bytes(int b, int c, int f, int h, int w, int groups, int per_frame, int dtype) {
  RingPlan pl;
  if (!make_ring_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return 0;
  return pl.partial_bytes + pl.final_bytes;
}

int gn_ring_launch(const void* x, void* y, const float* gamma, const float* beta, const float* temb, long long temb_ld, int b, int c, int f,
                   int h, int w, int groups, float eps, int per_frame, int apply_silu, int dtype, void* workspace,
                   size_t workspace_bytes, cudaStream_t st, bool* handled) {
  *handled = false;
  RingPlan pl;
  if (!make_ring_plan(b, c, f, h, w, groups, per_frame, dtype, &pl)) return CA_OK;
  if (!aligned16(x) || !aligned16(y) || !aligned16(gamma) || !aligned16(beta)) return CA_OK;
  const size_t need = pl.partial_bytes + pl.final_bytes;
  if (!workspace || workspace_bytes < need || !aligned16(workspace)) return CA_OK;

  RingParams p{};
  p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.temb = temb; p.temb_ld = temb_ld;
  p.c = c; p.groups = groups; p.cpg = c / groups; p.nvec = pl.nvec; p.k = pl.k;
  p.slabs = pl.slabs; p.cs = pl.cs; p.gs = pl.gs;
  p.gl = 1;
  while (p.gl < 32 && p.gl * 2 <= p.cpg && p.gl * 2 * pl.gs <= kGroupThreads) p.gl *= 2;
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
