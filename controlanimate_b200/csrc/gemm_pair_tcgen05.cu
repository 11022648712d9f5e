// Dense projections on the 5th-generation tensor cores, CTA-pair edition (tcgen05.mma.cta_group::2).
//
//   y[m, n_out] = epilogue( x[m, k] @ w[n, k]^T + bias[n] ) (+ residual[m, n_out])
//
// Replaces every nn.Linear of the motion module (reference animatediff/models/motion_module.py:147 proj_in,
// :155 proj_out, :215-219 to_q/to_k/to_v (fused [3C, C]) and to_out + residual, :221 GEGLU feed-forward) and of
// the spatial transformer (attention.py:133,157-162,268-289), which the reference runs as separate cuBLAS
// GEMMs + elementwise bias / residual / GEGLU passes.
//
// Why a CTA pair (profiles/r01a_gemm_ncu.md): an SM ingests at most ~43 B/clk from L2 (6300 B/clk chip-wide; the
// 1-CTA kernel sat exactly on that line with 45 % tensor-pipe utilisation).  A 128 x BN tile per SM needs
// (128 + BN) * 2 B per 256 * BN flop; two SMs sharing one 256 x BN tile each load their 128 rows of A and only
// HALF of B (the tensor core reads the other half from the peer's shared memory), and when the whole K extent of
// the B block fits in shared memory it is loaded ONCE per CTA ("B-stationary", K = 320: the 64x64-latent level)
// so that only A streams: 128 * 2 B per 256 * BN flop.
//
// Roles (per CTA; cluster = 2 CTAs = one TPC, rank 0 is the leader):
//   warp 0   TMA producer : x tile [128 x 64] + this CTA's half of the w tile [BN/2 x 64] (bf16, K-major,
//                           SWIZZLE_128B) into a smem ring; completion bytes of BOTH CTAs are credited to the
//                           leader's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 1   MMA issuer   : leader only; one thread issues tcgen05.mma.cta_group::2.kind::f16 (M=256, N=BN, K=16);
//                           tcgen05.commit ... .multicast::cluster releases the smem stage in both CTAs and publishes
//                           the accumulator (each CTA's TMEM holds its own 128 rows x BN fp32 columns, 2 stages)
//   warp 2   TMEM allocator (cta_group::2, 512 columns)
//   warps 4-11 epilogue   : 4 TMEM lane quarters x 2 column halves.  Per 32x32 box: tcgen05.ld (32 columns) ->
//                           +bias -> [GEGLU, packed f32x2 math] -> [+residual: TMA-prefetched two boxes ahead into the
//                           same SWIZZLE_64B staging slot] -> bf16 -> slot -> TMA tensor store.  Overlaps the next
//                           tile's MMAs through the second TMEM stage.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace ca {
int linear_1cta(const void* x, const void* w, const float* bias, const void* residual, void* y, long long m, int n, int k,
                long long ldx, long long ldr, long long ldy, int epilogue, int dtype, void* stream);
namespace {

constexpr int BM = 128;       // rows per CTA (UMMA_M = 256 per pair)
constexpr int BK = 64;        // one 128-byte swizzle atom of 16-bit elements
constexpr int UMMA_K = 16;
constexpr int kAccCols = 256; // TMEM columns per accumulator stage
constexpr int kEpiParts = 4;  // epilogue warps per TMEM lane quarter (column parts of the tile)
constexpr int kEpiWarps = 4 * kEpiParts;
constexpr int kMaxSlots = 3;  // staging slots per epilogue warp: 3 with a residual (TMA-prefetched two boxes ahead), else 1
constexpr int kSlotBytes = 2048;  // one 32 x 32 box of 16-bit elements
constexpr int kResAhead = 2;  // residual boxes in flight per warp
constexpr int kThreads = 128 + 32 * kEpiWarps;
constexpr int kMaxStages = 8;
constexpr uint32_t kABytes = BM * BK * 2;

struct PairParams {
  long long m;
  int n, k, bn, n_out;  // bn = accumulator columns per tile; n_out = output columns (n, or n/2 for GEGLU)
  int geglu, has_res;
  int num_m_blocks, num_n_blocks, num_k_blocks, stages;  // m blocks of 256 rows
  int stationary;       // the B block of this pair stays in smem for the whole kernel
  int slots;            // staging slots per epilogue warp
  long long tiles, tile_step;
  const float* bias;
  void* y;
  long long ldy;
  int dbg;            // development aid (CA_GEMM_DBG): 1 = epilogue only releases TMEM, 2 = tcgen05.ld but no stores
  long long* timing;  // CA_GEMM_TIMING=1: per-CTA cycle counters [gridDim.x][8] (development aid), else null
  uint32_t idesc, b_bytes, b_stride, stage_bytes, bres_bytes;  // b_bytes: this CTA's half tile; b_stride: 1024-aligned
};

// ---- tcgen05 wrappers (cta_group::2) --------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// arrive (once the MMAs issued so far have completed) on the barrier at this smem offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                              uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major, SWIZZLE_128B shared-memory matrix descriptor (sm_100 "version 1" format):
//   [0,14) start address >> 4 | [16,30) LBO >> 4 (unused for swizzled K-major, 1) | [32,46) SBO >> 4 (1024 B: one
//   8-row swizzle atom) | [46,48) version = 1 | [61,64) layout type = 2 (SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: two lanes of fp32 per issue slot) ---------------------------
struct f2 {
  uint64_t v;
};
__device__ __forceinline__ f2 pk(float lo, float hi) {
  f2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r.v) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void upk(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a.v)); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) {
  f2 d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d.v) : "l"(a.v), "l"(b.v), "l"(c.v));
  return d;
}
__device__ __forceinline__ f2 mul2(f2 a, f2 b) {
  f2 d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 add2(f2 a, f2 b) {
  f2 d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d.v) : "l"(a.v), "l"(b.v));
  return d;
}
__device__ __forceinline__ f2 splat(float c) { return pk(c, c); }

__device__ __forceinline__ float rcp_approx(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

// a * GELU(g) for two columns, exact-erf definition (diffusers GEGLU uses F.gelu default).  erfc by Abramowitz-Stegun
// 7.1.26 (abs err <= 1.5e-7): Phi(g) = 1/2 + copysign(1/2 - erfc(|g|/sqrt2)/2, g).  With z' = |g| * sqrt(log2(e)/2) the
// Gaussian factor is ex2(-z'^2); the polynomial runs on packed fp32x2 (11.5 issue slots per output instead of ~32).
__device__ __forceinline__ f2 geglu2(f2 a, f2 g) {
  float g0, g1;
  upk(g, g0, g1);
  const f2 zp = mul2(pk(fabsf(g0), fabsf(g1)), splat(0.84932180028801904f));        // |g| * sqrt(log2e / 2)
  const f2 d = fma2(zp, splat(0.3275911f * 0.83255461115769776f), splat(1.0f));     // 1 + p |g| / sqrt2
  float d0, d1;
  upk(d, d0, d1);
  const f2 t = pk(rcp_approx(d0), rcp_approx(d1));
  f2 poly = fma2(t, splat(-0.5f * 1.061405429f), splat(-0.5f * -1.453152027f));     // coefficients pre-halved, negated
  poly = fma2(poly, t, splat(-0.5f * 1.421413741f));
  poly = fma2(poly, t, splat(-0.5f * -0.284496736f));
  poly = fma2(poly, t, splat(-0.5f * 0.254829592f));
  poly = mul2(poly, t);
  const f2 w = mul2(zp, zp);
  float w0, w1;
  upk(w, w0, w1);
  const f2 e = pk(exp2f(-w0), exp2f(-w1));                                          // --use_fast_math: MUFU.EX2
  const f2 half_erf = fma2(poly, e, splat(0.5f));                                   // 1/2 - erfc/2  (>= 0)
  float h0, h1;
  upk(half_erf, h0, h1);
  const f2 phi = add2(pk(copysignf(h0, g0), copysignf(h1, g1)), splat(0.5f));
  return mul2(mul2(a, g), phi);
}

template <typename T>
__device__ __forceinline__ uint32_t cvt_pack(f2 v) {
  float lo, hi;
  upk(v, lo, hi);
  return pack2(lo, hi, T());
}
template <typename T>
__device__ __forceinline__ f2 unpack_f2(uint32_t w) {
  float lo, hi;
  unpack2(w, lo, hi, T());
  return pk(lo, hi);
}

// mbarrier wait that adds the cycles spent to `acc` when timing is on
__device__ __forceinline__ void timed_wait(uint64_t* bar, uint32_t parity, bool on, long long& acc) {
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
    gemm_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_r,
                     const PairParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full[2], tmem_empty[2], bres_full;
  __shared__ uint64_t res_full[kEpiWarps][kMaxSlots];
  __shared__ uint32_t tmem_base_slot;

  // SWIZZLE_128B tiles must start on 1024-byte boundaries (same offsets in both CTAs of the pair)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* bres = smem;                        // [num_k_blocks][b_stride] when stationary
  unsigned char* ring = smem + p.bres_bytes;         // [stages][stage_bytes]: A tile (+ B half tile)
  unsigned char* slots = ring + (size_t)p.stages * p.stage_bytes;  // [kEpiWarps][p.slots][kSlotBytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const long long pair = blockIdx.x >> 1;
  const int stages = p.stages;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);   // the leader's arrive.expect_tx covers the bytes of both CTAs (see producer)
      mbar_init(&empty_bar[s], 1);  // multicast tcgen05.commit
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
        mbar_init(&tmem_empty[s], 2 * kEpiWarps);  // epilogue warps of both CTAs
    }
    mbar_init(&bres_full, 1);
    for (int w = 0; w < kEpiWarps; ++w)
      for (int s = 0; s < kMaxSlots; ++s) mbar_init(&res_full[w][s], 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc_pair(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barrier inits of the peer are visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const bool timing = p.timing != nullptr;
  long long t_wait_a = 0, t_wait_b = 0;
  const long long t_begin = timing ? clock64() : 0;
  const bool active = !(p.stationary && pair >= p.tile_step);  // stationary: pairs beyond G * n_blocks have no tiles
  const int half = p.bn / 2;

  if (warp == 0) {
    // ===================== TMA producer (one thread per CTA) =====================
    if (active) {  // the whole warp walks the loop (convergent, warp-uniform state); one elected lane issues
      if (lane == 0) {
        prefetch_tensormap(&map_x);
        prefetch_tensormap(&map_w);
      }
      // rows of w this CTA supplies for n-block nb: plain = its half of [n0, n0+BN); GEGLU = values (rank 0) / gates (rank 1)
      auto b_row = [&](int nb) { return p.geglu ? (int)rank * (p.n / 2) + nb * half : nb * p.bn + (int)rank * half; };
      if (p.stationary) {
        const int nb = (int)(pair % p.num_n_blocks);
        const uint32_t bar = mapa_u32(&bres_full, 0);
        if (elect_one()) {
          if (leader) mbar_arrive_expect_tx(&bres_full, 2u * p.b_bytes * (uint32_t)p.num_k_blocks);
          for (int kb = 0; kb < p.num_k_blocks; ++kb)
            tma_load_2d_pair(bres + (size_t)kb * p.b_stride, &map_w, bar, kb * BK, b_row(nb));
        }
        __syncwarp();
      }
      const uint32_t stage_tx = kABytes + (p.stationary ? 0u : p.b_bytes);
      int stage = 0;
      uint32_t phase = 0;
      for (long long tile = pair; tile < p.tiles; tile += p.tile_step) {
        const int nb = (int)(tile % p.num_n_blocks);
        const int mb = (int)(tile / p.num_n_blocks);
        const int a_row = mb * 2 * BM + (int)rank * BM;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          timed_wait(&empty_bar[stage], phase ^ 1, timing, t_wait_a);
          unsigned char* sa = ring + (size_t)stage * p.stage_bytes;
          const uint32_t bar = mapa_u32(&full_bar[stage], 0);
          // Only the leader arrives (expecting the bytes of both CTAs); the peer just issues its loads.  A peer load
          // of phase n+1 cannot land before the leader's barrier finished phase n: the peer waits on empty_bar, which
          // the leader's MMAs signal only after consuming phase n.  (A remote release-arrive here costs a MEMBAR per
          // k-block on the producer's critical path: 3x slower main loop, profiles/r01b.)
          if (elect_one()) {
            if (leader) mbar_arrive_expect_tx(&full_bar[stage], 2u * stage_tx);
            tma_load_2d_pair(sa, &map_x, bar, kb * BK, a_row);
            if (!p.stationary) tma_load_2d_pair(sa + kABytes, &map_w, bar, kb * BK, b_row(nb));
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (timing && lane == 0) p.timing[blockIdx.x * 8 + 3] = t_wait_a;
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (leader && active) {  // convergent warp; tcgen05.mma / commit by one elected lane
      if (p.stationary) {
        mbar_wait(&bres_full, 0);
        tc_fence_after();
      }
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = pair; tile < p.tiles; tile += p.tile_step, ++it) {
        const int acc = (int)(it & 1);
        timed_wait(&tmem_empty[acc], (uint32_t)((it >> 1) & 1) ^ 1, timing, t_wait_a);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * kAccCols;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          timed_wait(&full_bar[stage], phase, timing, t_wait_b);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)stage * p.stage_bytes);
          const uint32_t sb = p.stationary ? smem_u32(bres + (size_t)kb * p.b_stride) : sa + kABytes;
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sb);
          if (elect_one()) {
#pragma unroll
            for (int ks = 0; ks < BK / UMMA_K; ++ks) {
              // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
              umma_f16_pair(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), p.idesc, (kb | ks) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty_bar[stage]);  // frees this smem stage in both CTAs once the MMAs above have read it
            if (kb == p.num_k_blocks - 1) umma_commit_pair(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
      if (timing && lane == 0) {
        p.timing[blockIdx.x * 8 + 0] = clock64() - t_begin;
        p.timing[blockIdx.x * 8 + 1] = t_wait_a;
        p.timing[blockIdx.x * 8 + 2] = t_wait_b;
      }
    }
  } else if (warp >= 4 && active) {
    // ===================== epilogue: TMEM -> registers -> swizzled smem slot -> TMA store =====================
    // warp w drains TMEM lane quarter q = w % 4 (hardware restriction) and column part (w - 4) / 4 of the tile, one
    // 32 x 32 box at a time (two tcgen05.ld of 16 columns).  Each warp owns `nslots` 2 KB staging slots
    // (SWIZZLE_64B: conflict-free 16-byte accesses); with a residual the slot is first filled by a TMA load issued two
    // boxes earlier, the sum is written back in place and the slot leaves through a TMA tensor store.
    const int ew = warp - 4, q = warp & 3, part = ew >> 2;
    const int out_cols = p.geglu ? half : p.bn;
    const int boxes = out_cols / 32;
    const int b_begin = (boxes * part) / kEpiParts, b_end = (boxes * (part + 1)) / kEpiParts;
    const uint32_t nslots = (uint32_t)p.slots;
    unsigned char* my_slots = slots + (size_t)ew * nslots * kSlotBytes;
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte chunk c of row r sits at chunk c ^ ((r >> 1) & 3)
    const uint32_t empty_remote0 = mapa_u32(&tmem_empty[0], 0), empty_remote1 = mapa_u32(&tmem_empty[1], 0);
    auto release_acc = [&](int acc) {  // TMEM stage fully read by this warp (tcgen05.wait::ld done): hand it back
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster_relaxed(acc ? empty_remote1 : empty_remote0);
      }
    };
    if (lane == 0 && p.has_res) prefetch_tensormap(&map_r);
    T* __restrict__ y = reinterpret_cast<T*>(p.y);
    auto tile_row0 = [&](long long tile) { return (int)(tile / p.num_n_blocks) * 2 * BM + (int)rank * BM + q * 32; };
    auto tile_col0 = [&](long long tile) { return (int)(tile % p.num_n_blocks) * out_cols; };

    // residual prefetch cursor (lane 0 only): runs kResAhead boxes ahead of the compute cursor
    long long pf_tile = pair;
    int pf_bx = b_begin;
    uint32_t pf_slot = 0;
    auto prefetch_one = [&]() {
      if (pf_tile >= p.tiles) return;
      mbar_arrive_expect_tx(&res_full[ew][pf_slot], kSlotBytes);
      tma_load_2d(my_slots + pf_slot * kSlotBytes, &map_r, &res_full[ew][pf_slot], tile_col0(pf_tile) + pf_bx * 32,
                  tile_row0(pf_tile));
      if (++pf_slot == nslots) pf_slot = 0;
      if (++pf_bx == b_end) {
        pf_bx = b_begin;
        pf_tile += p.tile_step;
      }
    };
    if (b_begin < b_end && p.has_res && lane == 0) {
#pragma unroll 1
      for (int i = 0; i < kResAhead; ++i) prefetch_one();
    }

    uint32_t slot = 0, slot_phase = 0;  // staging slot of the current box; parity of its residual barrier
    long long it = 0;
    for (long long tile = pair; tile < p.tiles; tile += p.tile_step, ++it) {
      const int acc = (int)(it & 1);
      const int row0 = tile_row0(tile), n0 = tile_col0(tile);
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccCols;
      timed_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1), timing, t_wait_a);
      tc_fence_after();
      if (b_begin == b_end || p.dbg == 1) {  // nothing to drain for this warp (tile narrower than 128 columns)
        release_acc(acc);
        continue;
      }
#pragma unroll 1
      for (int bx = b_begin; bx < b_end; ++bx) {
        const int col = bx * 32;
        unsigned char* buf = my_slots + slot * kSlotBytes;
        // the slot refilled below (residual of box + kResAhead) was drained kSlots - kResAhead boxes ago by this warp's
        // own ld.shared / st.global (program order), so lane 0 may hand it to the TMA engine right away
        if (p.has_res && lane == 0) prefetch_one();
        bool res_ready = false;
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
          uint32_t r[16];
          f2 v[8];
          const int c16 = col + 16 * hh;
          const int gc = n0 + c16;  // global output column
          tmem_ld16(taddr + c16, r);
          if (p.geglu) {
            uint32_t gt[16];
            tmem_ld16(taddr + out_cols + c16, gt);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
              if (p.bias) {
                ba = __ldg(reinterpret_cast<const float4*>(p.bias + gc + j));
                bg = __ldg(reinterpret_cast<const float4*>(p.bias + p.n / 2 + gc + j));
              }
              const f2 a0 = add2(pk(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pk(ba.x, ba.y));
              const f2 a1 = add2(pk(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pk(ba.z, ba.w));
              const f2 g0 = add2(pk(__uint_as_float(gt[j]), __uint_as_float(gt[j + 1])), pk(bg.x, bg.y));
              const f2 g1 = add2(pk(__uint_as_float(gt[j + 2]), __uint_as_float(gt[j + 3])), pk(bg.z, bg.w));
              v[j / 2] = geglu2(a0, g0);
              v[j / 2 + 1] = geglu2(a1, g1);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) ba = __ldg(reinterpret_cast<const float4*>(p.bias + gc + j));
              v[j / 2] = add2(pk(__uint_as_float(r[j]), __uint_as_float(r[j + 1])), pk(ba.x, ba.y));
              v[j / 2 + 1] = add2(pk(__uint_as_float(r[j + 2]), __uint_as_float(r[j + 3])), pk(ba.z, ba.w));
            }
          }
          if (hh == 1 && bx == b_end - 1) release_acc(acc);
          if (p.dbg == 2) {
            if (__uint_as_float(r[0]) == 123.456f) y[0] = T(v[0].v != 0);
            continue;
          }
          if (p.has_res) {
            if (!res_ready) {
              timed_wait(&res_full[ew][slot], slot_phase, timing, t_wait_b);
              res_ready = true;
            }
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const uint4 rr = *reinterpret_cast<const uint4*>(buf + lane * 64 + (((2 * hh + c) ^ sw) * 16));
              v[4 * c + 0] = add2(v[4 * c + 0], unpack_f2<T>(rr.x));
              v[4 * c + 1] = add2(v[4 * c + 1], unpack_f2<T>(rr.y));
              v[4 * c + 2] = add2(v[4 * c + 2], unpack_f2<T>(rr.z));
              v[4 * c + 3] = add2(v[4 * c + 3], unpack_f2<T>(rr.w));
            }
          }
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            uint4 o;
            o.x = cvt_pack<T>(v[4 * c + 0]);
            o.y = cvt_pack<T>(v[4 * c + 1]);
            o.z = cvt_pack<T>(v[4 * c + 2]);
            o.w = cvt_pack<T>(v[4 * c + 3]);
            *reinterpret_cast<uint4*>(buf + lane * 64 + (((2 * hh + c) ^ sw) * 16)) = o;
          }
        }
        // transposed write-out: the box sits in the slot row-major (64 B rows); every st.global.v4 of the warp now covers
        // 8 rows x 64 contiguous bytes (full 32-byte sectors) instead of 32 rows x 16 bytes
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int rr = i * 8 + (lane >> 2), cc = lane & 3;
          const uint4 o = *reinterpret_cast<const uint4*>(buf + rr * 64 + ((cc ^ ((rr >> 1) & 3)) * 16));
          const long long grow = (long long)row0 + rr;
          if (grow < p.m) stg_stream(y + grow * p.ldy + n0 + col + cc * 8, o);
        }
        __syncwarp();  // all lanes have read the slot before it is refilled / rewritten
        if (++slot == nslots) {
          slot = 0;
          slot_phase ^= 1;
        }
      }
    }
    if (timing && lane == 0 && q == 0 && part < 2) {  // part 0 and part 1 may own different numbers of boxes
      p.timing[blockIdx.x * 8 + 4 + 2 * part] = t_wait_a;
      p.timing[blockIdx.x * 8 + 5 + 2 * part] = t_wait_b;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast-commit into / read from this CTA's shared memory until here
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

int pick_bn(int n_cols, int cap) {  // largest multiple of 32 <= cap dividing n_cols (epilogue boxes are 32 columns wide)
  for (int bn = cap / 32 * 32; bn >= 32; bn -= 32)
    if (n_cols % bn == 0) return bn;
  return 0;
}

int gemm_impl() {  // CA_GEMM_IMPL=1cta selects the single-CTA yardstick kernel
  static int impl = -1;
  if (impl < 0) {
    const char* e = getenv("CA_GEMM_IMPL");
    impl = (e && e[0] == '1') ? 1 : 2;
  }
  return impl;
}
int gemm_stationary_allowed() {  // CA_GEMM_STATIONARY=0 disables the B-stationary schedule (A/B measurements)
  static int on = -1;
  if (on < 0) {
    const char* e = getenv("CA_GEMM_STATIONARY");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on;
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_linear(const void* x, const void* w, const float* bias,
                                                                const void* residual, void* y, long long m, int n, int k,
                                                                long long ldx, long long ldr, long long ldy, int epilogue,
                                                                int dtype, void* stream) {
  using namespace ca;
  if (gemm_impl() == 1) return linear_1cta(x, w, bias, residual, y, m, n, k, ldx, ldr, ldy, epilogue, dtype, stream);
  CA_CHECK_ARG(x && w && y, "linear: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "linear: dtype must be bf16 or f16 (tcgen05 kind::f16)");
  CA_CHECK_ARG(m >= 0 && n > 0 && k > 0, "linear: bad sizes m=%lld n=%d k=%d", m, n, k);
  CA_CHECK_ARG(epilogue == CA_EPI_NONE || epilogue == CA_EPI_GEGLU, "linear: unknown epilogue %d", epilogue);
  const bool geglu = epilogue == CA_EPI_GEGLU;
  CA_CHECK_ARG(k % 8 == 0 && ldx % 8 == 0 && ldx >= k, "linear: k and ldx must be multiples of 8 (16-byte TMA rows)");
  CA_CHECK_ARG(n % (geglu ? 64 : 32) == 0, "linear: n=%d must be a multiple of %d", n, geglu ? 64 : 32);
  const int n_out = geglu ? n / 2 : n;
  CA_CHECK_ARG(ldy >= n_out && ldy % 8 == 0 && (!residual || (ldr >= n_out && ldr % 8 == 0)), "linear: bad ldy/ldr");
  CA_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)), "linear: pointers must be 16-byte aligned");
  CA_CHECK_ARG(m < (1ll << 31), "linear: m too large");
  if (m == 0) return CA_OK;

  PairParams p{};
  p.m = m; p.n = n; p.k = k; p.geglu = geglu ? 1 : 0; p.n_out = n_out; p.has_res = residual ? 1 : 0;
  int sms = sm_count();
  const int pairs = sms / 2;
  p.num_m_blocks = (int)((m + 2 * BM - 1) / (2 * BM));
  p.num_k_blocks = (k + BK - 1) / BK;
  // accumulator columns per tile: as wide as divides n (<= 256), narrower when that leaves most pairs without a tile
  {
    static const int cap_env = getenv("CA_GEMM_BN") ? atoi(getenv("CA_GEMM_BN")) : 256;  // development aid
    int cap = cap_env;
    for (;;) {
      const int bn = geglu ? 2 * pick_bn(n / 2, cap / 2) : pick_bn(n, cap);
      CA_CHECK_ARG(bn >= 32, "linear: cannot tile n=%d", n);
      p.bn = bn;
      p.num_n_blocks = geglu ? (n / 2) / (bn / 2) : n / bn;
      const long long tiles = (long long)p.num_m_blocks * p.num_n_blocks;
      if (tiles * 2 > pairs || bn <= 64 || cap <= 64) break;  // enough tiles for more than half of the pairs
      cap = bn - (geglu ? 64 : 32) > 64 ? bn - (geglu ? 64 : 32) : 64;
    }
  }
  p.tiles = (long long)p.num_m_blocks * p.num_n_blocks;
  p.bias = bias; p.y = y; p.ldy = ldy;
  p.b_bytes = (uint32_t)(p.bn / 2) * BK * 2;
  p.b_stride = (p.b_bytes + 1023) & ~1023u;
  // instruction descriptor (kind::f16): D=f32 [4,6)=1; A/B format [7,10)/[10,13): 1=bf16, 0=f16; A,B K-major (bits
  // 15,16 = 0); N>>3 at [17,23); M>>4 at [24,29)  (M = 256: the pair's tile)
  const uint32_t fmt = dtype == CA_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);

  p.slots = residual ? kMaxSlots : 1;
  const size_t slot_bytes = (size_t)kEpiWarps * p.slots * kSlotBytes;
  const size_t budget = 227 * 1024 - 1024 /*alignment*/ - 1024 /*static*/ - slot_bytes;
  // B-stationary: the pair keeps its [BN x K] block resident when that still leaves >= 4 A stages and every n-block
  // gets at least one pair
  const size_t bres = (size_t)p.num_k_blocks * p.b_stride;
  p.stationary = gemm_stationary_allowed() && bres + 4 * kABytes <= budget && p.num_n_blocks <= pairs &&
                 p.num_m_blocks >= 2 * (pairs / p.num_n_blocks);
  long long grid_pairs;
  if (p.stationary) {
    p.bres_bytes = (uint32_t)bres;
    p.stage_bytes = kABytes;
    const int g = pairs / p.num_n_blocks;  // pairs per n-block
    p.tile_step = (long long)g * p.num_n_blocks;
    grid_pairs = pairs;  // pairs >= tile_step idle (fewer than num_n_blocks of them)
    if (grid_pairs > p.tile_step) grid_pairs = p.tile_step;
  } else {
    p.bres_bytes = 0;
    p.stage_bytes = kABytes + p.b_stride;
    grid_pairs = pairs < p.tiles ? pairs : p.tiles;
    p.tile_step = grid_pairs;
  }
  int stages = (int)((budget - p.bres_bytes) / p.stage_bytes);
  if (stages > kMaxStages) stages = kMaxStages;
  CA_CHECK_ARG(stages >= 2, "linear: tile does not fit shared memory");
  p.stages = stages;
  const size_t smem = (size_t)p.bres_bytes + (size_t)stages * p.stage_bytes + slot_bytes + 1024;

  CUtensorMap mx, mw, mr;
  const CUtensorMapDataType dt = dtype == CA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)m};
    const uint64_t strides[1] = {(uint64_t)ldx * 2};
    const uint32_t box[2] = {BK, BM};
    if (!encode_tensor_map(&mx, dt, 2, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CA_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)k, (uint64_t)n};
    const uint64_t strides[1] = {(uint64_t)k * 2};
    const uint32_t box[2] = {BK, (uint32_t)(p.bn / 2)};
    if (!encode_tensor_map(&mw, dt, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return CA_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)m};
    const uint32_t box[2] = {32, 32};
    const uint64_t sr[1] = {(uint64_t)(residual ? ldr : ldy) * 2};
    if (!encode_tensor_map(&mr, dt, 2, residual ? residual : y, dims, sr, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CA_ERR_CUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static long long* timing_buf = nullptr;
  static const bool timing_on = getenv("CA_GEMM_TIMING") != nullptr;
  static const int dbg_env = getenv("CA_GEMM_DBG") ? atoi(getenv("CA_GEMM_DBG")) : 0;
  p.dbg = dbg_env;
  if (timing_on) {
    if (!timing_buf) CA_CUDA(cudaMalloc(&timing_buf, 8 * 1024 * sizeof(long long)));
    CA_CUDA(cudaMemsetAsync(timing_buf, 0, 8 * 1024 * sizeof(long long), st));
    p.timing = timing_buf;
  }
  auto run = [&](auto kernel) -> int {
    CA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), 227 * 1024 - 1024));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * grid_pairs));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CA_CUDA(cudaLaunchKernelEx(&cfg, kernel, mx, mw, mr, p));
    if (timing_on) {  // development aid: synchronous, prints the mean per-CTA cycle attribution of this launch
      static long long host[8 * 1024];
      CA_CUDA(cudaStreamSynchronize(st));
      CA_CUDA(cudaMemcpy(host, timing_buf, sizeof(host), cudaMemcpyDeviceToHost));
      double s8[8] = {0}, lead = 0;
      const int ctas = (int)(2 * grid_pairs);
      for (int c = 0; c < ctas; ++c)
        for (int j = 0; j < 8; ++j) s8[j] += (double)host[c * 8 + j];
      lead = ctas / 2.0;
      fprintf(stderr, "[ca_linear timing] m=%lld n=%d k=%d bn=%d stat=%d stages=%d tiles/pair=%.1f | mma loop %.0f clk, wait tmem_empty %.0f, wait full %.0f | "
              "producer wait empty %.0f | epi part0 wait acc %.0f res %.0f | epi part1 wait acc %.0f res %.0f\n", m, n, k, p.bn, p.stationary, p.stages,
              (double)p.tiles / (double)p.tile_step, s8[0] / lead, s8[1] / lead, s8[2] / lead, s8[3] / ctas, s8[4] / ctas, s8[5] / ctas, s8[6] / ctas, s8[7] / ctas);
    }
    return CA_OK;
  };
  if (dtype == CA_BF16) return run(gemm_pair_kernel<__nv_bfloat16>);
  return run(gemm_pair_kernel<__half>);
}
