// Spatial self-attention core (row N2): softmax(Q K^T * scale) V over the h*w sites of one frame, per head — the
// arithmetic of BasicTransformerBlock.attn1 (animatediff/models/attention.py:268-271) reached through the reference's
// AttentionProcessor (modules/attention_processor.py:56-62 / :247-256: xformers memory_efficient_attention or SDPA).
//
// One CTA = one (frame, head) and 256 query rows (two 128-row tiles), walking the keys in tiles of 128:
//   S_t = Q_t K_j^T            tcgen05.mma (cta_group::1, 128 x 128 x 16, K-major operands by TMA) into TMEM
//   P_t = exp2(S_t c - m_t)    one thread per query row (TMEM lane = row), packed FFMA2 / MUFU.EX2, row sums in fp32;
//                              the running maximum m_t is only raised when the tile maximum exceeds it by more than 8
//                              (log2 units), so the accumulator is rescaled a handful of times per row, not per tile
//   O_t += P_t V_j             tcgen05.mma with P_t from shared memory (K-major) and V_j as it lies in HBM ([key][dim]:
//                              the MN-major B operand — no transpose anywhere)
// The two query tiles share every K / V tile and alternate on the tensor pipe: S of tile j + 1 is issued as soon as a
// softmax warp group has read S of tile j, so the MUFU pipe — the bound of this kernel at head_dim 40: 128 x 128 exponentials
// per 2 x 128 x 128 x 88 MACs (profiles/r02_notes.md §4) — always has a warp group feeding it.
// Q / K / V are column slices of the packed [T, 3C] projection output: 4-D TMA maps (dim, site, head, frame), pad columns
// of the 64-wide boxes zero-filled; head_dim 40 and 80 (one or two 64-column chunks).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kTile = 128;                 // query rows per tile = keys per tile
constexpr int kSoftmaxWarps = 8;           // two warp groups: warps 0-3 -> query tile 0, warps 4-7 -> query tile 1
// one MMA-issuing warp per query tile (kMmaWarp, kMmaWarp + 1): each walks S_j, P V_{j-1} for ITS warp group in order, so a wait
// for one group's barrier never delays the other group's MMAs
constexpr int kProducerWarp = kSoftmaxWarps, kMmaWarp = kSoftmaxWarps + 1, kAllocWarp = kSoftmaxWarps + 3;
constexpr int kThreads = (kSoftmaxWarps + 4) * 32;
constexpr int kChunkBytes = kTile * 128;   // one 64-column (128-byte) SWIZZLE_128B chunk of a 128-row tile: 16 KB
#ifndef CA_FMHA_POLY_EVERY
#define CA_FMHA_POLY_EVERY 3
#endif
constexpr int kPolyEvery = CA_FMHA_POLY_EVERY;   // every n-th pair of exponentials of the fast path on the FMA pipes (0: none)
constexpr float kRaise = 8.0f;             // raise the running maximum only past this margin (log2 units)

struct FmhaParams {
  int sites, heads, frames, head_dim, chunks;  // chunks = ceil(head_dim / 64)
  int kv_tiles, stages;
  float scale_log2;                            // softmax scale * log2(e)
  void* o;
  long long ldo;                               // output row stride (elements)
  uint32_t idesc_qk, idesc_pv;
  long long* timing;   // CA_FMHA_TIMING=1: [16] counters (development aid), else null
  int dbg;             // CA_FMHA_DBG (development aid, wrong results): 1 = no P V MMAs, 2 = no S MMAs, 3 = P V with a K-major V descriptor
};

// ---- tcgen05 wrappers (cta_group::1) -------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
      "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]),
      "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]),
      "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// SWIZZLE_128B shared-memory matrix descriptors (sm_100 "version 1"): [0,14) start >> 4 | [16,30) LBO >> 4 | [32,46) SBO >> 4 |
// [46,48) version 1 | [61,64) layout 2.  K-major (rows of 128 B = 64 K-elements): SBO = 1024 B between 8-row groups, LBO
// unused.  MN-major (rows of 128 B = 64 MN-elements, one row per K index): SBO = 1024 B between groups of 8 K rows, LBO =
// distance between 64-element MN blocks.
__device__ __forceinline__ uint64_t desc_k_major(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ uint64_t desc_mn_major(uint32_t smem_addr, uint32_t lbo_bytes) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(1024 >> 4) << 32) |
         ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
}

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// 2^x for a pair of x <= 8 on the FMA / ALU pipes instead of the MUFU: x = n + f with n = round(x), f in [-0.5, 0.5];
// 2^f by a cubic (max relative error 7.5e-5, far below the rounding of P to 16 bits) and n added into the exponent field.
// The softmax pass is bound by the 16 exponentials per clock the MUFU pipe delivers; a third of them go this way
// (measured, 4096^2 x 40: none 1683 us, every 4th pair 1600, every 3rd 1568, every 2nd 1794).
__device__ __forceinline__ float2 ex2_poly2(float2 x) {
  x.x = fmaxf(x.x, -126.0f);
  x.y = fmaxf(x.y, -126.0f);
  const float2 magic = make_float2(12582912.0f, 12582912.0f);   // 1.5 * 2^23: the sum's low mantissa bits are round(x)
  const float2 t = __fadd2_rn(x, magic);
  const float2 f = __fadd2_rn(x, __fadd2_rn(magic, make_float2(-t.x, -t.y)));   // x - (t - magic)
  float2 q = __ffma2_rn(make_float2(5.517202416e-02f, 5.517202416e-02f), f, make_float2(2.426111656e-01f, 2.426111656e-01f));
  q = __ffma2_rn(q, f, make_float2(6.932609214e-01f, 6.932609214e-01f));
  q = __ffma2_rn(q, f, make_float2(9.999280683e-01f, 9.999280683e-01f));
  return make_float2(__int_as_float(__float_as_int(q.x) + (__float_as_int(t.x) << 23)),
                     __int_as_float(__float_as_int(q.y) + (__float_as_int(t.y) << 23)));
}

template <typename T>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
template <>
__device__ __forceinline__ uint32_t pack_pair<__half>(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

// development aid (CA_FMHA_TIMING=1): cycles spent in each kind of wait, per role, accumulated into p.timing
__device__ __forceinline__ void twait(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
  if (!on) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += clock64() - t0;
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
    fmha_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                const __grid_constant__ CUtensorMap map_v, const FmhaParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  constexpr int kMaxStages = 4;
  __shared__ uint64_t q_full, k_full[kMaxStages], k_empty[kMaxStages], v_full[kMaxStages], v_empty[kMaxStages];
  __shared__ uint64_t s_full[2], s_free[2], p_ready[2], pv_done[2][2];   // pv_done[t][b]: P V of the tiles with j & 1 == b
  __shared__ uint32_t tmem_base_slot;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int chunks = p.chunks;
  const uint32_t tile_bytes = (uint32_t)chunks * kChunkBytes;     // one Q / K / V tile: 128 rows x chunks x 128 B
  unsigned char* q_s = smem;                                       // [2 query tiles][chunks][128][128 B]
  unsigned char* k_s = q_s + 2 * tile_bytes;                       // [stages][chunks][128][128 B]
  unsigned char* v_s = k_s + (size_t)p.stages * tile_bytes;        // [stages][chunks][128 keys][128 B]
  unsigned char* p_s = v_s + (size_t)p.stages * tile_bytes;        // [2 query tiles][2 buffers][2 key chunks][128][128 B]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q_block = blockIdx.x, head = blockIdx.y, frame = blockIdx.z;
  const int q_row0 = q_block * 2 * kTile;

  if (threadIdx.x == 0) {
    mbar_init(&q_full, 1);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 2);   // one commit per MMA warp
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 2);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1);
      mbar_init(&s_free[t], 4);    // one arrival per softmax warp of the tile
      mbar_init(&p_ready[t], 4);
      mbar_init(&pv_done[t][0], 1);
      mbar_init(&pv_done[t][1], 1);
    }
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  // TMEM columns: S of query tile t at t * 128 (128 fp32 columns), O of query tile t at 256 + t * 128 (chunks * 64 used)
  const int stages = p.stages, kv_tiles = p.kv_tiles;
  const bool timing = p.timing != nullptr;
  long long tw0 = 0, tw1 = 0, tw2 = 0, tw3 = 0;
  const long long t_begin = timing ? clock64() : 0;

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tensormap(&map_q);
      prefetch_tensormap(&map_k);
      prefetch_tensormap(&map_v);
    }
    if (elect_one()) {
      mbar_arrive_expect_tx(&q_full, 2 * tile_bytes);
      for (int t = 0; t < 2; ++t)
        for (int c = 0; c < chunks; ++c)
          tma_load_4d(q_s + t * tile_bytes + c * kChunkBytes, &map_q, &q_full, c * 64, q_row0 + t * kTile, head, frame);
    }
    __syncwarp();
    int stage = 0;
    uint32_t phase = 0;
    for (int j = 0; j < kv_tiles; ++j) {
      mbar_wait(&k_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[stage], tile_bytes);
        for (int c = 0; c < chunks; ++c)
          tma_load_4d(k_s + (size_t)stage * tile_bytes + c * kChunkBytes, &map_k, &k_full[stage], c * 64, j * kTile, head, frame);
      }
      __syncwarp();
      mbar_wait(&v_empty[stage], phase ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[stage], tile_bytes);
        for (int c = 0; c < chunks; ++c)
          tma_load_4d(v_s + (size_t)stage * tile_bytes + c * kChunkBytes, &map_v, &v_full[stage], c * 64, j * kTile, head, frame);
      }
      __syncwarp();
      if (++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == kMmaWarp || warp == kMmaWarp + 1) {
    // ===================== MMA issuer of query tile t =====================
    // per key tile j:  S_j (as soon as the warp group has read S_{j-1}), then O += P_{j-1} V_{j-1} (as soon as P_{j-1} is written)
    const int t = warp - kMmaWarp;
    mbar_wait(&q_full, 0);
    tc_fence_after();
    const int ksteps_qk = (p.head_dim + 15) / 16;   // K = head_dim rounded up to 16 (pad columns are zero)
    const uint32_t qa = smem_u32(q_s + t * tile_bytes), pa0 = smem_u32(p_s + t * 4 * kChunkBytes);
    const uint32_t s_tmem = tmem_base + (uint32_t)t * 128u, o_tmem = tmem_base + 256u + (uint32_t)t * 128u;
    int stage = 0, prev_stage = 0;
    uint32_t phase = 0, prev_phase = 0;
    for (int j = 0; j <= kv_tiles; ++j) {
      const uint32_t ppar = (uint32_t)((j - 1) & 1);
      if (j < kv_tiles) {
        twait(&k_full[stage], phase, timing, tw0);
        if (j > 0) twait(&s_free[t], ppar, timing, tw1);
        tc_fence_after();
        const uint32_t ka = smem_u32(k_s + (size_t)stage * tile_bytes);
        if (elect_one()) {
          for (int ks = 0; ks < (p.dbg == 2 ? 0 : ksteps_qk); ++ks) {
            const uint32_t off = (uint32_t)(ks >> 2) * kChunkBytes + (uint32_t)(ks & 3) * 32;
            umma_f16(s_tmem, desc_k_major(qa + off), desc_k_major(ka + off), p.idesc_qk, ks ? 1u : 0u);
          }
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[stage]);
        }
        __syncwarp();
      }
      if (j > 0) {
        twait(&v_full[prev_stage], prev_phase, timing, tw2);
        twait(&p_ready[t], ppar, timing, tw3);
        tc_fence_after();
        const uint32_t va = smem_u32(v_s + (size_t)prev_stage * tile_bytes);
        const uint32_t pa = pa0 + (uint32_t)((j - 1) & 1) * 2 * kChunkBytes;
        if (elect_one()) {
          for (int ks = 0; ks < (p.dbg == 1 ? 0 : kTile / 16); ++ks) {   // 16 keys per step: P advances 32 B in its row, V by two 8-key groups
            const uint32_t poff = (uint32_t)(ks >> 2) * kChunkBytes + (uint32_t)(ks & 3) * 32;
            umma_f16(o_tmem, desc_k_major(pa + poff), desc_mn_major(va + (uint32_t)ks * 2048u, kChunkBytes), p.idesc_pv,
                     (j == 1 && ks == 0) ? 0u : 1u);
          }
          umma_commit(&pv_done[t][(j - 1) & 1]);
          umma_commit(&v_empty[prev_stage]);
        }
        __syncwarp();
      }
      prev_stage = stage;
      prev_phase = phase;
      if (j < kv_tiles && ++stage == stages) {
        stage = 0;
        phase ^= 1;
      }
    }
    if (timing && lane == 0 && t == 0) {
      atomicAdd((unsigned long long*)&p.timing[0], (unsigned long long)(clock64() - t_begin));
      atomicAdd((unsigned long long*)&p.timing[1], (unsigned long long)tw0);
      atomicAdd((unsigned long long*)&p.timing[2], (unsigned long long)tw1);
      atomicAdd((unsigned long long*)&p.timing[3], (unsigned long long)tw2);
      atomicAdd((unsigned long long*)&p.timing[4], (unsigned long long)tw3);
    }
  } else if (warp < kSoftmaxWarps) {
    // ===================== softmax: one thread per query row =====================
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;                       // row inside the query tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_addr + (uint32_t)t * 128u;
    const uint32_t o_addr = tmem_base + lane_addr + 256u + (uint32_t)t * 128u;
    unsigned char* p_row0 = p_s + t * 4 * kChunkBytes + row * 128;
    const int sw = row & 7;
    const float c = p.scale_log2;
    float m_ref = -INFINITY;     // running maximum (log2 units, scaled), raised lazily
    float l_sum = 0.f;
    const int kv_len = p.sites;
    for (int j = 0; j < kv_tiles; ++j) {
      const uint32_t par = (uint32_t)(j & 1);
      const int valid = kv_len - j * kTile;                    // keys of this tile that exist (>= 128 except in the last tile)
      twait(&s_full[t], par, timing, tw0);
      tc_fence_after();
      const long long t_a = timing ? clock64() : 0;
      uint32_t sa[32], sb[32];
      unsigned char* p_row = p_row0 + (j & 1) * 2 * kChunkBytes;
      // The exponentials of a tile are taken against the running reference m_ref, which is known BEFORE the tile (one pass
      // over S, interleaved with the search for the tile's own maximum).  If that maximum exceeds the reference by more than
      // the margin, the reference is raised, O and l are rescaled and the pass is repeated (always for the first tile of a
      // row, a handful of times afterwards).  P therefore never exceeds 2^8.
      float mx;
      auto max2 = [](float a, float b) {   // two-input maximum: the compiler's fused 3-input FMNMX3 issues at a quarter of the rate
        float d;
        asm("max.ftz.f32 %0, %1, %2;" : "=f"(d) : "f"(a), "f"(b));
        return d;
      };
      auto max_chunk = [&](const uint32_t (&sv)[32], int cb) {
        if (valid >= (cb + 1) * 32) {
          float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            m0 = max2(m0, __uint_as_float(sv[i]));
            m1 = max2(m1, __uint_as_float(sv[i + 1]));
            m2 = max2(m2, __uint_as_float(sv[i + 2]));
            m3 = max2(m3, __uint_as_float(sv[i + 3]));
          }
          mx = max2(mx, max2(max2(m0, m1), max2(m2, m3)));
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cb * 32 + i < valid) mx = max2(mx, __uint_as_float(sv[i]));
        }
      };
      float sum0, sum1, sum2, sum3, nm;
      auto exp_chunk = [&](const uint32_t (&sv)[32], int cb) {
        float pv[32];
#pragma unroll
        for (int i = 0; i < 32; ++i) pv[i] = ex2f(fmaf(__uint_as_float(sv[i]), c, nm));
        if (valid < (cb + 1) * 32) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (cb * 32 + i >= valid) pv[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          sum0 += pv[i];
          sum1 += pv[i + 1];
          sum2 += pv[i + 2];
          sum3 += pv[i + 3];
        }
        unsigned char* dst = p_row + (cb >> 1) * kChunkBytes;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 w;
          w.x = pack_pair<T>(pv[8 * u + 0], pv[8 * u + 1]);
          w.y = pack_pair<T>(pv[8 * u + 2], pv[8 * u + 3]);
          w.z = pack_pair<T>(pv[8 * u + 4], pv[8 * u + 5]);
          w.w = pack_pair<T>(pv[8 * u + 6], pv[8 * u + 7]);
          const int unit = (cb & 1) * 4 + u;
          *reinterpret_cast<uint4*>(dst + ((unit ^ sw) * 16)) = w;
        }
      };
      // P is double buffered: this tile's buffer was last read by the P V of tile j - 2
      if (j >= 2) twait(&pv_done[t][j & 1], (uint32_t)(((j >> 1) - 1) & 1), timing, tw1);
      // ---- fast path: a full tile with a known reference.  Straight-line and software pipelined by one chunk: the block that
      // issues the 32 exponentials of chunk k also holds the row sums, conversions and stores of chunk k - 1 and the maximum
      // search of chunk k, so the MUFU pipe is fed while the other pipes work (with two softmax warps per scheduler nothing
      // else hides those latencies).
      bool done = false, tripped = false;
      if (valid >= kTile && __all_sync(0xffffffffu, m_ref > -INFINITY)) {
        tripped = true;   // (only read when the fast path falls through)
        float pa[32], pb[32];
        float2 acc0 = make_float2(0.f, 0.f), acc1 = make_float2(0.f, 0.f);
        const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
        auto E = [&](const uint32_t (&sv)[32], float (&pv)[32]) {   // exponentials of one chunk: 3 of 4 pairs on the MUFU
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])), c2, nm2);
            if (kPolyEvery > 0 && (i / 2) % (kPolyEvery > 0 ? kPolyEvery : 1) == 0) {
              const float2 e = ex2_poly2(x);
              pv[i] = e.x;
              pv[i + 1] = e.y;
            } else {
              pv[i] = ex2f(x.x);
              pv[i + 1] = ex2f(x.y);
            }
          }
        };
        auto F = [&](const float (&pv)[32], int cb) {               // row sums, conversion, stores of one chunk
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            acc0 = __fadd2_rn(acc0, make_float2(pv[i], pv[i + 1]));
            acc1 = __fadd2_rn(acc1, make_float2(pv[i + 2], pv[i + 3]));
          }
          unsigned char* dst = p_row + (cb >> 1) * kChunkBytes;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            uint4 w;
            w.x = pack_pair<T>(pv[8 * u + 0], pv[8 * u + 1]);
            w.y = pack_pair<T>(pv[8 * u + 2], pv[8 * u + 3]);
            w.z = pack_pair<T>(pv[8 * u + 4], pv[8 * u + 5]);
            w.w = pack_pair<T>(pv[8 * u + 6], pv[8 * u + 7]);
            const int unit = (cb & 1) * 4 + u;
            *reinterpret_cast<uint4*>(dst + ((unit ^ sw) * 16)) = w;
          }
        };
        tmem_ld32(s_addr, sa);
        tmem_ld_wait();
        tmem_ld32(s_addr + 32u, sb);
        E(sa, pa);
        tmem_ld_wait();
        tmem_ld32(s_addr + 64u, sa);
        E(sb, pb);
        F(pa, 0);
        tmem_ld_wait();
        tmem_ld32(s_addr + 96u, sb);
        E(sa, pa);
        F(pb, 1);
        tmem_ld_wait();
        // Does the reference still hold?  No maximum search in this path: a score above m_ref + 8 shows as an exponential
        // above 2^8, hence as a row sum above 2^8 (every term is >= 0; an overflow to inf trips it as well).  The test is
        // conservative — a sum of many small terms can trip it too; the general path then sets the reference to the exact
        // maximum, after which the sum of a 128-key tile cannot exceed 128.  The last chunk is tested on its raw scores so
        // that S can be released before its exponentials are taken.
        float m3 = -INFINITY, m3b = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 2) {
          m3 = max2(m3, __uint_as_float(sb[i]));
          m3b = max2(m3b, __uint_as_float(sb[i + 1]));
        }
        F(pa, 2);
        const float part = (acc0.x + acc0.y) + (acc1.x + acc1.y);
        const bool trip = !(part <= 256.0f) || max2(m3, m3b) * c > m_ref + kRaise;
        if (!__any_sync(0xffffffffu, trip)) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);     // S is in registers for good: the next tile's scores may overwrite it
          E(sb, pb);
          F(pb, 3);
          sum0 = acc0.x + acc0.y;
          sum1 = acc1.x + acc1.y;
          sum2 = sum3 = 0.f;
          done = true;
        }
        // else: fall through to the general path below, which raises the reference and repeats the tile (S is still intact)
      }
      for (; !done;) {
        const bool have_ref = __any_sync(0xffffffffu, m_ref > -INFINITY);   // false only before a row's first tile (warp-uniform)
        sum0 = sum1 = sum2 = sum3 = 0.f;
        nm = -m_ref;
        mx = -INFINITY;
        tmem_ld32(s_addr, sa);
        tmem_ld_wait();
        tmem_ld32(s_addr + 32u, sb);
        max_chunk(sa, 0);
        if (have_ref) exp_chunk(sa, 0);
        tmem_ld_wait();
        tmem_ld32(s_addr + 64u, sa);
        max_chunk(sb, 1);
        if (have_ref) exp_chunk(sb, 1);
        tmem_ld_wait();
        tmem_ld32(s_addr + 96u, sb);
        max_chunk(sa, 2);
        if (have_ref) exp_chunk(sa, 2);
        tmem_ld_wait();
        max_chunk(sb, 3);
        mx *= c;
        // after a trip of the fast path's sum test the reference moves to the exact maximum whenever that is higher at all, so
        // that the following tiles' sums stay below the trip level
        const bool raise = mx > m_ref + kRaise || (tripped && mx > m_ref);
        if (!__any_sync(0xffffffffu, raise)) {
          // the common case: S has been read for good (its last chunk is in registers) — the tensor pipe may overwrite it
          // with the next tile's scores while the last quarter of the exponentials is still being computed
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[t]);
          exp_chunk(sb, 3);
          break;
        }
        // raise the reference (warp-uniform control flow: the tcgen05 loads / stores below are .sync.aligned)
        if (have_ref && j > 0) {                   // O and l were accumulated against the old reference
          twait(&pv_done[t][(j - 1) & 1], (uint32_t)(((j - 1) >> 1) & 1), timing, tw2);
          tc_fence_after();
          const float f = raise ? ex2f(m_ref - mx) : 1.0f;
          l_sum *= f;
          for (int oc = 0; oc < p.chunks * 64; oc += 32) {
            uint32_t o[32];
            tmem_ld32(o_addr + (uint32_t)oc, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * f);
            tmem_st32(o_addr + (uint32_t)oc, o);
          }
          tmem_st_wait();
        }
        if (raise) m_ref = mx;
      }
      if (timing) tw3 += clock64() - t_a;
      tc_fence_before();
      l_sum += (sum0 + sum1) + (sum2 + sum3);
      fence_proxy_async();       // generic-proxy writes of P -> visible to the tensor core's async-proxy reads
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_ready[t]);
    }
    if (timing && lane == 0 && quarter == 0) {
      atomicAdd((unsigned long long*)&p.timing[5 + 5 * t], (unsigned long long)(clock64() - t_begin));
      atomicAdd((unsigned long long*)&p.timing[6 + 5 * t], (unsigned long long)tw0);
      atomicAdd((unsigned long long*)&p.timing[7 + 5 * t], (unsigned long long)tw1);
      atomicAdd((unsigned long long*)&p.timing[8 + 5 * t], (unsigned long long)tw2);
      atomicAdd((unsigned long long*)&p.timing[9 + 5 * t], (unsigned long long)tw3);
    }
    // epilogue: O / l -> global ([frame * sites + site][head * head_dim + e])
    mbar_wait(&pv_done[t][(kv_tiles - 1) & 1], (uint32_t)(((kv_tiles - 1) >> 1) & 1));
    tc_fence_after();
    const int site = q_row0 + t * kTile + row;
    const float inv = 1.0f / l_sum;
    T* out = reinterpret_cast<T*>(p.o) + ((long long)frame * p.sites + site) * p.ldo + (long long)head * p.head_dim;
    for (int oc = 0; oc < p.head_dim; oc += 32) {
      uint32_t o[32];
      tmem_ld32(o_addr + (uint32_t)oc, o);
      tmem_ld_wait();
      if (site < p.sites) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (oc + 8 * u < p.head_dim) {   // head_dim % 8 == 0
            uint4 w;
            w.x = pack_pair<T>(__uint_as_float(o[8 * u + 0]) * inv, __uint_as_float(o[8 * u + 1]) * inv);
            w.y = pack_pair<T>(__uint_as_float(o[8 * u + 2]) * inv, __uint_as_float(o[8 * u + 3]) * inv);
            w.z = pack_pair<T>(__uint_as_float(o[8 * u + 4]) * inv, __uint_as_float(o[8 * u + 5]) * inv);
            w.w = pack_pair<T>(__uint_as_float(o[8 * u + 6]) * inv, __uint_as_float(o[8 * u + 7]) * inv);
            *reinterpret_cast<uint4*>(out + oc + 8 * u) = w;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_spatial_attn_core(const void* q, const void* k, const void* v, void* o,
                                                                           int frames, int sites, int heads, int head_dim,
                                                                           long long ldq, long long ldk, long long ldv,
                                                                           long long ldo, float scale, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(q && k && v && o, "spatial_attn_core: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "spatial_attn_core: dtype must be bf16 or f16");
  CA_CHECK_ARG(frames > 0 && sites > 0 && heads > 0, "spatial_attn_core: bad sizes");
  CA_CHECK_ARG(head_dim % 8 == 0 && head_dim >= 16 && head_dim <= 64, "spatial_attn_core: head_dim %d unsupported (16..64, multiple of 8)", head_dim);
  const long long c = (long long)heads * head_dim;
  CA_CHECK_ARG(ldq >= c && ldk >= c && ldv >= c && ldo >= c && ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "spatial_attn_core: row strides must cover heads * head_dim and be multiples of 8");
  CA_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "spatial_attn_core: pointers must be 16-byte aligned");
  FmhaParams p{};
  p.sites = sites; p.heads = heads; p.frames = frames; p.head_dim = head_dim;
  p.chunks = (head_dim + 63) / 64;
  p.kv_tiles = (sites + kTile - 1) / kTile;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = o; p.ldo = ldo;
  const size_t tile_bytes = (size_t)p.chunks * kChunkBytes;
  const size_t fixed = 2 * tile_bytes + 2 * 2 * 2 * (size_t)kChunkBytes + 1024;   // Q (2 tiles) + P (2 tiles x 2 buffers x 2 chunks) + alignment
  const size_t budget = 227 * 1024 - 2048;
  CA_CHECK_ARG(fixed + 2 * 2 * tile_bytes <= budget, "spatial_attn_core: head_dim %d does not fit shared memory", head_dim);
  p.stages = (int)((budget - fixed) / (2 * tile_bytes));
  if (p.stages > 4) p.stages = 4;
  const size_t smem = fixed + (size_t)p.stages * 2 * tile_bytes;
  // instruction descriptors (kind::f16): D = f32 [4,6) = 1; A / B format [7,10) / [10,13): 1 = bf16, 0 = f16; bit 15 / 16 = A / B
  // MN-major; N >> 3 at [17,23); M >> 4 at [24,29)
  const uint32_t fmt = dtype == CA_BF16 ? 1u : 0u;
  p.idesc_qk = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);
  p.idesc_pv = (1u << 4) | (fmt << 7) | (fmt << 10) | (1u << 16) | ((uint32_t)((p.chunks * 64) >> 3) << 17) | ((uint32_t)(kTile >> 4) << 24);

  CUtensorMap mq, mk, mv;
  const CUtensorMapDataType dt = dtype == CA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  auto make = [&](CUtensorMap* m, const void* base, long long ld) {
    // (dim, site, head, frame); rows of the 64-wide boxes beyond head_dim and sites beyond `sites` are zero-filled
    const uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)sites, (uint64_t)heads, (uint64_t)frames};
    const uint64_t strides[3] = {(uint64_t)ld * 2, (uint64_t)head_dim * 2, (uint64_t)sites * (uint64_t)ld * 2};
    const uint32_t box[4] = {64, (uint32_t)kTile, 1, 1};
    return encode_tensor_map(m, dt, 4, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B);
  };
  if (!make(&mq, q, ldq) || !make(&mk, k, ldk) || !make(&mv, v, ldv)) return CA_ERR_CUDA;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  dim3 grid((unsigned)((sites + 2 * kTile - 1) / (2 * kTile)), (unsigned)heads, (unsigned)frames);
  static long long* timing_buf = nullptr;
  static const bool timing_on = getenv("CA_FMHA_TIMING") != nullptr;
  static const int dbg_env = getenv("CA_FMHA_DBG") ? atoi(getenv("CA_FMHA_DBG")) : 0;
  p.dbg = dbg_env;
  if (timing_on) {
    if (!timing_buf) CA_CUDA(cudaMalloc(&timing_buf, 16 * sizeof(long long)));
    CA_CUDA(cudaMemsetAsync(timing_buf, 0, 16 * sizeof(long long), st));
    p.timing = timing_buf;
  }
  auto run = [&](auto kernel) -> int {
    CA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), 227 * 1024 - 1024));
    kernel<<<grid, kThreads, smem, st>>>(mq, mk, mv, p);
    CA_CUDA(cudaGetLastError());
    if (timing_on) {   // development aid: synchronous; mean cycles per CTA and key tile
      long long h[16];
      CA_CUDA(cudaStreamSynchronize(st));
      CA_CUDA(cudaMemcpy(h, timing_buf, sizeof(h), cudaMemcpyDeviceToHost));
      const double n = (double)grid.x * grid.y * grid.z * p.kv_tiles;
      fprintf(stderr, "[ca_spatial_attn timing] sites=%d hd=%d per key tile: mma loop %.0f clk (wait k %.0f, s_free %.0f, v %.0f, p_ready %.0f) | "
              "softmax t0 %.0f clk (wait s_full %.0f, pv_done %.0f, pass A %.0f, pass B %.0f) | t1 %.0f (%.0f, %.0f, %.0f, %.0f)\n",
              sites, head_dim, h[0] / n, h[1] / n, h[2] / n, h[3] / n, h[4] / n, h[5] / n, h[6] / n, h[7] / n, h[8] / n, h[9] / n,
              h[10] / n, h[11] / n, h[12] / n, h[13] / n, h[14] / n);
    }
    return CA_OK;
  };
  if (dtype == CA_BF16) return run(fmha_kernel<__nv_bfloat16>);
  return run(fmha_kernel<__half>);
}
