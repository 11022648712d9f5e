// Dense projections on the 5th-generation tensor cores, CTA-pair edition (tcgen05.mma.cta_group::2), round 2.
//
//   y[m, n_out] = epilogue( x[m, k] @ w[n, k]^T + bias[n] ) (+ residual[m, n_out])
//
// Replaces every nn.Linear of the motion module (reference animatediff/models/motion_module.py:147 proj_in,
// :155 proj_out, :215-219 to_q/to_k/to_v (fused [3C, C]) and to_out + residual, :221 GEGLU feed-forward) and of
// the spatial transformer (attention.py:133,157-162,268-289), which the reference runs as separate cuBLAS
// GEMMs + elementwise bias / residual / GEGLU passes.
//
// What bounds it (profiles/r01_gemm_notes.md, profiles/r02_gemm_notes.md):
//   * an SM ingests ~36-40 B/clk from L2, so operand bytes per flop per SM decide the tensor-pipe ceiling: a CTA PAIR shares
//     one 256 x BN tile (each SM loads its 128 rows of A and HALF of B), and for long K with narrow N the tile covers up to
//     512 accumulator columns (two UMMA sub-tiles sharing every A stage, single TMEM stage);
//   * round 1's epilogue (46 KB of SASS, fully unrolled) took ~2000 cycles per 32 x 32 box: every hot loop now fits the
//     instruction cache (one box per iteration, one kernel instance per epilogue kind);
//   * a lane owns one accumulator ROW, so storing (or reading the residual) straight from registers touches 32 cache lines
//     per instruction (~66 L1TEX cycles each): boxes go through a SWIZZLE_64B smem slot and the TMA engine in both
//     directions (residual: TMA load issued two boxes ahead; output: TMA store, slot reuse gated by bulk-group waits);
//   * exact-erf GELU cost two MUFU ops per output (rcp + ex2: 2048 MUFU cycles per 128 x 128 tile, close to the 2560
//     cycles the K = 320 MMAs take); Phi(g) is now an odd degree-19 polynomial on the packed fp32x2 FMA pipe;
//   * the issuing thread's instruction stream is part of the tensor pipe's schedule: a predicated-off UTCHMMA still costs
//     an issue slot (10 % at K = 1280), descriptor arithmetic inside the unrolled k-loop cost 2x.
//
// Roles (per CTA; cluster = 2 CTAs = one TPC, rank 0 is the leader):
//   warp 16  TMA producer : x tile [128 x 64] + this CTA's half of each w sub-tile [BN/2 x 64] (bf16, K-major,
//                           SWIZZLE_128B) into a smem ring; completion bytes of BOTH CTAs are credited to the
//                           leader's full barrier (cp.async.bulk.tensor ... .cta_group::2)
//   warp 17  MMA issuer   : leader only; one thread issues tcgen05.mma.cta_group::2.kind::f16 (M=256, N=BN, K=16);
//                           tcgen05.commit ... .multicast::cluster releases the smem stage in both CTAs and publishes
//                           the accumulator (each CTA's TMEM holds its own 128 rows x BN fp32 columns)
//   warp 18  TMEM allocator (cta_group::2, 512 columns)
//   warps 0-15 epilogue   : 4 TMEM lane quarters x 4 column parts; 16-column chunks dealt round-robin to the parts.
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
#ifndef CA_EPI_PARTS
#define CA_EPI_PARTS 2
#endif
constexpr int kEpiParts = CA_EPI_PARTS;  // epilogue warps per TMEM lane quarter
constexpr int kEpiWarps = 4 * kEpiParts;
constexpr int kThreads = 128 + 32 * kEpiWarps;
// Warp roles.  The SMSP arbiter favours the HIGHEST warp id among eligible warps (B300_MICROARCH.md), and the single
// MMA-issuing thread must never queue behind sixteen busy epilogue warps (measured: 257 instead of 128 cycles per UMMA
// with the issuer in warp 1), so the epilogue owns warps 0-15 and the producer / issuer / allocator sit above them.
constexpr int kProducerWarp = kEpiWarps, kMmaWarp = kEpiWarps + 1, kAllocWarp = kEpiWarps + 2;
constexpr int kMaxStages = 12;
constexpr int kMaxSlots = 4;    // staging slots per epilogue warp: 4 with a residual (TMA-prefetched two boxes ahead), else 2
constexpr int kSlotBytes = 2048;  // one 32 x 32 box of 16-bit elements (SWIZZLE_64B)
constexpr uint32_t kABytes = BM * BK * 2;
// EPI_LN / EPI_LN_GEGLU: x is the RAW input of a LayerNorm whose gain is folded into w; the epilogue applies the row's
// (mean, rstd):  y = rstd * (acc - mean * colsum[n]) + shift[frame(row)][n]   (see ca_linear_ln)
enum { EPI_BIAS = 0, EPI_RES = 1, EPI_GEGLU = 2, EPI_LN = 3, EPI_LN_GEGLU = 4 };

struct PairParams {
  long long m;
  int n, k;
  int bn;        // accumulator columns of one UMMA sub-tile (UMMA N)
  int nsub;      // sub-tiles per tile: they share every A stage (nsub * bn accumulator columns per tile)
  int out_cols;  // output columns per tile: nsub * bn, or bn / 2 (GEGLU)
  int num_m_blocks, num_n_blocks, num_k_blocks, stages;  // m blocks of 256 rows
  int acc_stages;  // TMEM accumulator stages: 2 when nsub * bn <= 256, else 1
  long long tiles;
  const float* bias;  // [n]; with a folded LayerNorm: the shift table [ln_shift_rows][n]
  const float2* ln_stats;   // (mean, rstd) per row of x (folded LayerNorm), else null
  const float* ln_colsum;   // sum_k w[n][k] per output column (of the gain-folded weights)
  int ln_sites, ln_frames, ln_shift_rows;  // shift row of token row r: (r / ln_sites) % ln_frames (row 0 if the table has one row)
  void* y;
  const void* res;
  long long ldy, ldr;
  int slots;          // staging slots per epilogue warp
  int dbg;            // development aid (CA_GEMM_DBG): 1 = epilogue only releases TMEM
  long long* timing;  // CA_GEMM_TIMING=1: per-CTA cycle counters [gridDim.x][8] (development aid), else null
  uint32_t idesc, b_bytes, stage_bytes;  // b_bytes: this CTA's half of ONE sub-tile k-block
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
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
__device__ __forceinline__ float fma_sat(float a, float b, float c) {
  float d;
  asm("fma.rn.sat.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

// a * GELU(g) for two columns, exact-erf definition (diffusers GEGLU uses F.gelu default): GELU(g) = g * Phi(g),
// Phi(g) = 1/2 + erf(g / sqrt 2) / 2.  With s = g / 4.5 and u = min(s^2, 1):  Phi(g) = sat(1/2 + s * P(u)), P of degree 9
// fitted on |g| <= 4.5 (max |error of Phi| 9e-6 including the clamp: 1 - Phi(4.5) = 3.4e-6; beyond the fit interval
// s * P(1) keeps growing in magnitude and the saturation returns exactly 0 or 1).  No MUFU: ten packed FMAs per PAIR of
// outputs instead of rcp + ex2 per output.
__device__ __forceinline__ f2 geglu2(f2 a, f2 g) {
  const f2 s = mul2(g, splat(1.0f / 4.5f));
  float s0, s1;
  upk(mul2(s, s), s0, s1);
  const f2 u = pk(fminf(s0, 1.0f), fminf(s1, 1.0f));
  f2 p = fma2(splat(-4.261638838e+00f), u, splat(2.484848156e+01f));
  p = fma2(p, u, splat(-6.462894735e+01f));
  p = fma2(p, u, splat(9.980938078e+01f));
  p = fma2(p, u, splat(-1.031514738e+02f));
  p = fma2(p, u, splat(7.648830514e+01f));
  p = fma2(p, u, splat(-4.258644516e+01f));
  p = fma2(p, u, splat(1.823896685e+01f));
  p = fma2(p, u, splat(-6.051783177e+00f));
  p = fma2(p, u, splat(1.795147712e+00f));
  float p0, p1;
  upk(p, p0, p1);
  upk(s, s0, s1);
  const f2 phi = pk(fma_sat(s0, p0, 0.5f), fma_sat(s1, p1, 0.5f));
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
  const long long t0 = on ? clock64() : 0;
  mbar_wait(bar, parity);
  if (on) acc += clock64() - t0;
}

// The tiles of one pair, in the order all three roles walk them (32-bit, division-free stepping): round robin over the
// pairs, n-block-minor (t = mb * num_n_blocks + nb), so that concurrently running pairs share their A rows and the small B
// matrix in L2.  (A "B-stationary" schedule — contiguous tile ranges per pair with the B block resident in smem — was
// measured and lost at every config-2 shape: the groups drift apart, A is re-read from HBM once per n-block, and every
// n-block change drains the pipeline; profiles/r02_gemm_notes.md.)
struct TileWalk {
  int left, mb_, nb_, step_mb, step_nb, num_n_blocks;
  __device__ TileWalk(const PairParams& p, int pair, int pairs) : num_n_blocks(p.num_n_blocks) {
    const int tiles = (int)p.tiles;
    left = pair < tiles ? (tiles - pair + pairs - 1) / pairs : 0;
    mb_ = pair / num_n_blocks;
    nb_ = pair - mb_ * num_n_blocks;
    step_mb = pairs / num_n_blocks;
    step_nb = pairs - step_mb * num_n_blocks;
  }
  __device__ bool valid() const { return left > 0; }
  __device__ void next() {
    --left;
    mb_ += step_mb;
    nb_ += step_nb;
    if (nb_ >= num_n_blocks) {
      nb_ -= num_n_blocks;
      ++mb_;
    }
  }
  __device__ int nb() const { return nb_; }
  __device__ int mb() const { return mb_; }
};

template <typename T, int EPI>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_pair_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                     const __grid_constant__ CUtensorMap map_y, const __grid_constant__ CUtensorMap map_r, const PairParams p) {
  constexpr bool kGeglu = EPI == EPI_GEGLU || EPI == EPI_LN_GEGLU;
  constexpr bool kLn = EPI == EPI_LN || EPI == EPI_LN_GEGLU;
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages], tmem_full[2], tmem_empty[2];
  __shared__ uint64_t res_full[kEpiWarps][kMaxSlots];
  __shared__ uint32_t tmem_base_slot;

  // SWIZZLE_128B tiles must start on 1024-byte boundaries (same offsets in both CTAs of the pair)
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* ring = smem;  // [stages][stage_bytes]: A tile + nsub B half tiles
  unsigned char* slots = ring + (size_t)p.stages * p.stage_bytes;  // [kEpiWarps][p.slots][kSlotBytes]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;
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
    for (int w = 0; w < kEpiWarps; ++w)
      for (int sl = 0; sl < kMaxSlots; ++sl) mbar_init(&res_full[w][sl], 1);
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc_pair(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barrier inits of the peer are visible before any remote arrive / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const bool timing = p.timing != nullptr;
  long long t_wait_a = 0, t_wait_b = 0;
  const long long t_begin = timing ? clock64() : 0;
  const int half = p.bn / 2;

  if (warp == kProducerWarp) {
    // ===================== TMA producer (the whole warp walks the loop; one elected lane issues) =====================
    if (lane == 0) {
      prefetch_tensormap(&map_x);
      prefetch_tensormap(&map_w);
    }
    // rows of w this CTA supplies for sub-tile j of n-block nb: its half of the sub-tile's columns, or (GEGLU) the value
    // rows (rank 0) / gate rows (rank 1) of the n-block
    auto b_row = [&](int nb, int j) {
      return kGeglu ? (int)rank * (p.n / 2) + nb * half : (nb * p.nsub + j) * p.bn + (int)rank * half;
    };
    const uint32_t stage_tx = kABytes + (uint32_t)p.nsub * p.b_bytes;
    int stage = 0;
    uint32_t phase = 0;
    for (TileWalk w(p, pair, pairs); w.valid(); w.next()) {
      const int nb = w.nb();
      const int a_row = w.mb() * 2 * BM + (int)rank * BM;
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
          tma_load_2d_pair(sa + kABytes, &map_w, bar, kb * BK, b_row(nb, 0));
          if (p.nsub == 2) tma_load_2d_pair(sa + kABytes + p.b_bytes, &map_w, bar, kb * BK, b_row(nb, 1));
        }
        __syncwarp();
        if (++stage == stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    if (timing && lane == 0) {
      p.timing[blockIdx.x * 8 + 3] = t_wait_a;
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (one thread of the leader CTA) =====================
    if (leader) {  // convergent warp; tcgen05.mma / commit by one elected lane
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (TileWalk w(p, pair, pairs); w.valid(); w.next(), ++it) {
        const int acc = p.acc_stages == 2 ? (int)(it & 1) : 0;
        const uint32_t acc_par = p.acc_stages == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
        timed_wait(&tmem_empty[acc], acc_par ^ 1, timing, t_wait_a);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * 256u;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          timed_wait(&full_bar[stage], phase, timing, t_wait_b);
          tc_fence_after();
          const uint32_t sa = smem_u32(ring + (size_t)stage * p.stage_bytes);
          const uint32_t sb = sa + kABytes;
          const uint64_t adesc = make_sw128_desc(sa), bdesc0 = make_sw128_desc(sb);
          const uint32_t accum0 = kb != 0 ? 1u : 0u;
          if (elect_one()) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field.  Two separate
            // instruction streams: a predicated-off UTCHMMA still costs a tensor-pipe issue slot (measured: 10 % on the
            // K = 1280 shapes when the second sub-tile's MMAs were merely predicated away).
            if (p.nsub == 1) {
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks)
                umma_f16_pair(tmem_d, adesc + (uint64_t)(ks * 2), bdesc0 + (uint64_t)(ks * 2), p.idesc, ks ? 1u : accum0);
            } else {
              const uint64_t bdesc1 = make_sw128_desc(sb + p.b_bytes);
#pragma unroll
              for (int ks = 0; ks < BK / UMMA_K; ++ks) {
                umma_f16_pair(tmem_d, adesc + (uint64_t)(ks * 2), bdesc0 + (uint64_t)(ks * 2), p.idesc, ks ? 1u : accum0);
                umma_f16_pair(tmem_d + (uint32_t)p.bn, adesc + (uint64_t)(ks * 2), bdesc1 + (uint64_t)(ks * 2), p.idesc,
                              ks ? 1u : accum0);
              }
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
  } else if (warp < kEpiWarps) {
    // ===================== epilogue: TMEM -> registers -> swizzled smem box -> TMA store =====================
    // warp w drains TMEM lane quarter q = w % 4 (hardware restriction); the tile's 32-column boxes are dealt round robin to
    // the parts.  A lane owns one output row, so a direct st.global would touch 32 different cache lines per instruction
    // (measured: ~66 L1TEX cycles per store instruction, 4400 cycles per 128 x 256 tile — longer than the K = 320 MMAs);
    // instead the box is written to a SWIZZLE_64B slot (conflict-free 16-byte st.shared) and leaves through the TMA
    // engine, and the residual box arrives the same way (TMA load issued two boxes ahead into the slot ring).
    const int q = warp & 3, part = warp >> 2;
    const int n_boxes = p.out_cols >> 5;
    const int my_n = (n_boxes - part + kEpiParts - 1) / kEpiParts;  // boxes part, part + kEpiParts, ...
    const uint32_t nslots = (uint32_t)p.slots;
    unsigned char* my_slots = slots + (size_t)warp * nslots * kSlotBytes;
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B: 16-byte unit u of row r sits at unit u ^ ((r >> 1) & 3)
    const uint32_t empty_remote0 = mapa_u32(&tmem_empty[0], 0), empty_remote1 = mapa_u32(&tmem_empty[1], 0);
    auto release_acc = [&](int acc) {  // this warp's share of the stage is in registers (tcgen05.wait::ld done)
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) mbar_arrive(&tmem_empty[acc]);
        else mbar_arrive_cluster_relaxed(acc ? empty_remote1 : empty_remote0);
      }
    };
    if (lane == 0) {
      prefetch_tensormap(&map_y);
      if (EPI == EPI_RES) prefetch_tensormap(&map_r);
    }
    // residual prefetch cursor (meaningful in lane 0 only): runs two boxes ahead of the compute cursor
    TileWalk pf(p, pair, pairs);
    int pf_i = 0;
    uint32_t pf_slot = 0;
    auto prefetch_one = [&]() {
      if (!pf.valid() || my_n == 0) return;
      mbar_arrive_expect_tx(&res_full[warp][pf_slot], kSlotBytes);
      tma_load_2d(my_slots + pf_slot * kSlotBytes, &map_r, &res_full[warp][pf_slot],
                  pf.nb() * p.out_cols + (part + pf_i * kEpiParts) * 32, pf.mb() * 2 * BM + (int)rank * BM + q * 32);
      if (++pf_slot == nslots) pf_slot = 0;
      if (++pf_i == my_n) {
        pf_i = 0;
        pf.next();
      }
    };
    if (EPI == EPI_RES && lane == 0) {
      prefetch_one();
      prefetch_one();
    }
    uint32_t slot = 0, slot_phase = 0;
    long long it = 0;
    // folded LayerNorm: (mean, rstd) of this lane's row, fetched one tile ahead (a cold 8-byte read per row and tile)
    auto load_stats = [&](const TileWalk& tw) {
      const long long r = (long long)tw.mb() * 2 * BM + (int)rank * BM + q * 32 + lane;
      return (tw.valid() && r < p.m) ? __ldg(p.ln_stats + r) : make_float2(0.f, 0.f);
    };
    float2 st_next = make_float2(0.f, 0.f);
    // ... and the per-column parameters of the NEXT box: lane j fetches column j's (colsum, shift) [GEGLU: value and gate
    // column] one box ahead, so the L2 round trip of the table rows (they do not stay in the small L1 next to 200 KB of
    // shared memory) is off the box's critical path; at the box they are broadcast through the staging slot.
    auto shift_row = [&](const TileWalk& tw) {
      const int r0 = tw.mb() * 2 * BM + (int)rank * BM + q * 32;
      return p.ln_shift_rows > 1 ? p.bias + (size_t)((r0 / p.ln_sites) % p.ln_frames) * (size_t)p.n : p.bias;
    };
    auto load_prm = [&](const TileWalk& tw, int box_i) {
      float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
      if (!tw.valid() || my_n == 0) return r;
      const int col = tw.nb() * p.out_cols + (part + box_i * kEpiParts) * 32 + lane;
      const float* sh_row = shift_row(tw);
      r.x = __ldg(p.ln_colsum + col);
      r.y = __ldg(sh_row + col);
      if constexpr (kGeglu) {
        r.z = __ldg(p.ln_colsum + p.n / 2 + col);
        r.w = __ldg(sh_row + p.n / 2 + col);
      }
      return r;
    };
    float4 prm_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if constexpr (kLn) {
      st_next = load_stats(TileWalk(p, pair, pairs));
      prm_next = load_prm(TileWalk(p, pair, pairs), 0);
    }
    for (TileWalk w(p, pair, pairs); w.valid(); w.next(), ++it) {
      const int acc = p.acc_stages == 2 ? (int)(it & 1) : 0;
      const uint32_t acc_par = p.acc_stages == 2 ? (uint32_t)((it >> 1) & 1) : (uint32_t)(it & 1);
      const int row0 = w.mb() * 2 * BM + (int)rank * BM + q * 32;
      const int col0 = w.nb() * p.out_cols;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * 256u;
      timed_wait(&tmem_full[acc], acc_par, timing, t_wait_a);
      tc_fence_after();
      if (my_n == 0 || p.dbg == 1) {
        release_acc(acc);
        continue;
      }
      // folded LayerNorm: this lane's row statistics and the warp's shift row (32 | ln_sites: one frame per warp)
      f2 ln_rs = splat(1.f), ln_nm = splat(0.f);
      const float* bias = p.bias;
      TileWalk ahead = w;
      if constexpr (kLn) {
        const float2 st = st_next;
        ahead.next();
        st_next = load_stats(ahead);
        ln_rs = splat(st.y);
        ln_nm = splat(-st.x * st.y);
      }
      // one box per iteration, NOT unrolled: the hot loops have to stay inside the instruction cache (round 1's fully
      // unrolled epilogue was 46 KB of SASS)
#pragma unroll 1
      for (int i = 0; i < my_n; ++i) {
        const int bc = (part + i * kEpiParts) * 32;  // first column of the box inside the tile
        unsigned char* buf = my_slots + slot * kSlotBytes;
        // slot hygiene (bulk groups are per thread: lane 0 issues every TMA store of this warp).  With a residual the
        // slot refilled now (box + 2) was last read by the store issued two boxes ago; without, the slot written below was
        // last read by the store issued nslots boxes ago.
        if (lane == 0) {
          if (EPI == EPI_RES) {
            bulk_wait_read<1>();
            prefetch_one();
          } else {
            bulk_wait_read<1>();
          }
        }
        __syncwarp();
        f2 v[16];
        if constexpr (kGeglu) {
#pragma unroll
          for (int h2 = 0; h2 < 2; ++h2) {
            uint32_t a[16], g[16];
            tmem_ld16(taddr + (uint32_t)(bc + 16 * h2), a);
            tmem_ld16(taddr + (uint32_t)(half + bc + 16 * h2), g);
            // per-column parameters: fetched while the TMEM reads are in flight
            const int gc = col0 + bc + 16 * h2;
            f2 sa[8], sg[8];
            if constexpr (kLn) {
              if (h2 == 0) {  // broadcast the box's prefetched (colsum, shift) pairs through the (still unused) staging slot
                const float4 prm = prm_next;
                prm_next = i + 1 < my_n ? load_prm(w, i + 1) : load_prm(ahead, 0);
                float* sc = reinterpret_cast<float*>(buf);
                sc[(lane >> 1) * 4 + (lane & 1)] = prm.x;
                sc[(lane >> 1) * 4 + 2 + (lane & 1)] = prm.y;
                sc[64 + (lane >> 1) * 4 + (lane & 1)] = prm.z;
                sc[64 + (lane >> 1) * 4 + 2 + (lane & 1)] = prm.w;
                __syncwarp();
              }
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                const float4 ca = *reinterpret_cast<const float4*>(buf + (8 * h2 + j) * 16);
                const float4 cg = *reinterpret_cast<const float4*>(buf + 256 + (8 * h2 + j) * 16);
                sa[j] = fma2(ln_nm, pk(ca.x, ca.y), pk(ca.z, ca.w));
                sg[j] = fma2(ln_nm, pk(cg.x, cg.y), pk(cg.z, cg.w));
              }
              if (h2 == 1) __syncwarp();  // every lane has read the parameters before the output box overwrites them
            } else {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
                if (bias) {
                  ba = __ldg(reinterpret_cast<const float4*>(bias + gc + j));
                  bg = __ldg(reinterpret_cast<const float4*>(bias + p.n / 2 + gc + j));
                }
                sa[j / 2] = pk(ba.x, ba.y);
                sa[j / 2 + 1] = pk(ba.z, ba.w);
                sg[j / 2] = pk(bg.x, bg.y);
                sg[j / 2 + 1] = pk(bg.z, bg.w);
              }
            }
            tmem_ld_wait();
            if (h2 == 1 && i == my_n - 1) release_acc(acc);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              const f2 av = pk(__uint_as_float(a[j]), __uint_as_float(a[j + 1])), gv = pk(__uint_as_float(g[j]), __uint_as_float(g[j + 1]));
              v[8 * h2 + j / 2] = kLn ? geglu2(fma2(ln_rs, av, sa[j / 2]), fma2(ln_rs, gv, sg[j / 2])) : geglu2(add2(av, sa[j / 2]), add2(gv, sg[j / 2]));
            }
          }
        } else {
          uint32_t a[32];
          tmem_ld32(taddr + (uint32_t)bc, a);
          // the per-column parameters do not depend on the accumulator: fetch them while the TMEM read is in flight
          f2 sh[16];
          if constexpr (kLn) {
            const float4 prm = prm_next;
            prm_next = i + 1 < my_n ? load_prm(w, i + 1) : load_prm(ahead, 0);
            float* sc = reinterpret_cast<float*>(buf);
            sc[(lane >> 1) * 4 + (lane & 1)] = prm.x;
            sc[(lane >> 1) * 4 + 2 + (lane & 1)] = prm.y;
            __syncwarp();
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float4 cs = *reinterpret_cast<const float4*>(buf + j * 16);
              sh[j] = fma2(ln_nm, pk(cs.x, cs.y), pk(cs.z, cs.w));
            }
            __syncwarp();  // every lane has read the parameters before the output box overwrites them
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f);
              if (bias) ba = __ldg(reinterpret_cast<const float4*>(bias + col0 + bc + j));
              sh[j / 2] = pk(ba.x, ba.y);
              sh[j / 2 + 1] = pk(ba.z, ba.w);
            }
          }
          tmem_ld_wait();
          if (i == my_n - 1) release_acc(acc);
#pragma unroll
          for (int j = 0; j < 32; j += 2) {
            const f2 av = pk(__uint_as_float(a[j]), __uint_as_float(a[j + 1]));
            v[j / 2] = kLn ? fma2(ln_rs, av, sh[j / 2]) : add2(av, sh[j / 2]);
          }
        }
        if constexpr (EPI == EPI_RES) {
          timed_wait(&res_full[warp][slot], slot_phase, timing, t_wait_b);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint4 rr = *reinterpret_cast<const uint4*>(buf + lane * 64 + ((u ^ sw) * 16));
            v[4 * u + 0] = add2(v[4 * u + 0], unpack_f2<T>(rr.x));
            v[4 * u + 1] = add2(v[4 * u + 1], unpack_f2<T>(rr.y));
            v[4 * u + 2] = add2(v[4 * u + 2], unpack_f2<T>(rr.z));
            v[4 * u + 3] = add2(v[4 * u + 3], unpack_f2<T>(rr.w));
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          uint4 o;
          o.x = cvt_pack<T>(v[4 * u + 0]);
          o.y = cvt_pack<T>(v[4 * u + 1]);
          o.z = cvt_pack<T>(v[4 * u + 2]);
          o.w = cvt_pack<T>(v[4 * u + 3]);
          *reinterpret_cast<uint4*>(buf + lane * 64 + ((u ^ sw) * 16)) = o;
        }
        fence_proxy_async();  // generic-proxy writes of the box -> visible to the TMA engine
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_y, buf, col0 + bc, row0);  // rows >= m are clipped by the tensor map
          bulk_commit();
        }
        if (++slot == nslots) {
          slot = 0;
          slot_phase ^= 1;
        }
      }
    }
    if (lane == 0) bulk_wait<0>();  // the slots must outlive the last stores' reads (and the writes must land before exit)
    if (timing && lane == 0 && q == 0 && part == 0) {
      p.timing[blockIdx.x * 8 + 6] = t_wait_a;
      p.timing[blockIdx.x * 8 + 7] = t_wait_b;
    }
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still multicast-commit into / read from this CTA's shared memory until here
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

// One way to tile the problem, and what the busiest pair's main loop is expected to cost with it.
struct Plan {
  int bn = 0, nsub = 1, stages = 0, acc_stages = 2, num_n_blocks = 0;
  size_t stage_bytes = 0;
  double cost = 1e300;
};

constexpr size_t kSmemBudget = 227 * 1024 - 1024 /*alignment*/ - 1024 /*static*/;

// Cycles of the busiest pair under the limits this kernel was measured against (profiles/r02_gemm_notes.md): the tensor
// pipe (BN / 2 cycles per UMMA), the ~36 B/clk an SM ingests from L2 (A tile + its half of B per k-step), whole tiles per
// pair (wave quantisation), and the accumulator drain, which is exposed when the tile needs all of TMEM (one stage).
// The constants reproduce the measured order of the candidate tilings on all twenty config-2 shapes.
bool make_plan(Plan& pl, int bn, int nsub, int n, int k, long long m, bool geglu, bool has_res, int pairs) {
  const int tile_cols = bn * nsub;
  const int num_k_blocks = (k + BK - 1) / BK;
  const int num_m_blocks = (int)((m + 2 * BM - 1) / (2 * BM));
  pl.bn = bn;
  pl.nsub = nsub;
  pl.acc_stages = tile_cols <= 256 ? 2 : 1;
  pl.num_n_blocks = n / tile_cols;
  pl.stage_bytes = kABytes + (size_t)nsub * (bn / 2) * BK * 2;  // bn % 16 == 0 -> every B half tile is a multiple of 1024 B
  const size_t budget = kSmemBudget - (size_t)kEpiWarps * (has_res ? kMaxSlots : 2) * kSlotBytes;
  if (3 * pl.stage_bytes > budget) return false;
  pl.stages = (int)(budget / pl.stage_bytes);
  if (pl.stages > kMaxStages) pl.stages = kMaxStages;
  const double steps = num_k_blocks * 4.0;
  const double mma = tile_cols / 2.0;
  const double ingest = (4096.0 + tile_cols * 16.0) / 36.0;
  const double step = mma > ingest ? mma : ingest;
  const int out_cols = geglu ? bn / 2 : tile_cols;
  const double drain = 500.0 + (geglu ? 1400.0 : 700.0) * ((out_cols / 32 + kEpiParts - 1) / kEpiParts);
  double tile = steps * step;
  if (pl.acc_stages == 1) tile += drain;     // nothing overlaps the drain
  else if (drain > tile) tile = drain;       // two stages: the drain hides behind the next tile's MMAs unless it is longer
  tile += 300.0;
  const long long tiles = (long long)num_m_blocks * pl.num_n_blocks;
  pl.cost = (double)((tiles + pairs - 1) / pairs) * tile;
  return true;
}

// a LayerNorm folded into the projection that consumes it (ca_linear_ln)
struct LnFold {
  const float2* stats;
  const float* colsum;
  int sites, frames, shift_rows;
};

int linear_impl(const void* x, const void* w, const float* bias, const void* residual, void* y, long long m, int n, int k,
                long long ldx, long long ldr, long long ldy, int epilogue, int dtype, void* stream, const LnFold* ln) {
  CA_CHECK_ARG(x && w && y, "linear: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "linear: dtype must be bf16 or f16 (tcgen05 kind::f16)");
  CA_CHECK_ARG(m >= 0 && n > 0 && k > 0, "linear: bad sizes m=%lld n=%d k=%d", m, n, k);
  CA_CHECK_ARG(epilogue == CA_EPI_NONE || epilogue == CA_EPI_GEGLU, "linear: unknown epilogue %d", epilogue);
  const bool geglu = epilogue == CA_EPI_GEGLU;
  CA_CHECK_ARG(!(geglu && residual), "linear: GEGLU with a residual is not supported");
  CA_CHECK_ARG(k % 8 == 0 && ldx % 8 == 0 && ldx >= k, "linear: k and ldx must be multiples of 8 (16-byte TMA rows)");
  CA_CHECK_ARG(n % (geglu ? 64 : 32) == 0, "linear: n=%d must be a multiple of %d", n, geglu ? 64 : 32);
  const int n_out = geglu ? n / 2 : n;
  CA_CHECK_ARG(ldy >= n_out && ldy % 8 == 0 && (!residual || (ldr >= n_out && ldr % 8 == 0)), "linear: bad ldy/ldr");
  CA_CHECK_ARG(aligned16(x) && aligned16(w) && aligned16(y) && (!residual || aligned16(residual)), "linear: pointers must be 16-byte aligned");
  CA_CHECK_ARG(m < (1ll << 31), "linear: m too large");
  if (m == 0) return CA_OK;

  const int pairs = sm_count() / 2;
  // ---- choose the tiling: every (bn, nsub) that divides n and fits shared memory / TMEM, cheapest first ----
  Plan best;
  {
    static const char* cfg_env = getenv("CA_GEMM_CFG");  // development aid: "bn,nsub"
    int f_bn = 0, f_nsub = 0;
    if (cfg_env) sscanf(cfg_env, "%d,%d", &f_bn, &f_nsub);
    const int unit = geglu ? 64 : 16;
    for (int bn = 256; bn >= unit; bn -= unit) {
      for (int nsub = 1; nsub <= (geglu ? 1 : 2); ++nsub) {
        if (n % (bn * nsub) != 0 || bn * nsub > 512 || (bn * nsub) % 32 != 0) continue;  // epilogue boxes are 32 columns wide
        if (nsub == 2 && bn * 2 <= 256) continue;  // one wider sub-tile does the same with fewer instructions
        if (f_bn && (bn != f_bn || nsub != f_nsub)) continue;
        Plan pl;
        if (!make_plan(pl, bn, nsub, n, k, m, geglu, residual != nullptr, pairs)) continue;
        if (pl.cost < best.cost) best = pl;
      }
    }
    CA_CHECK_ARG(best.bn > 0, "linear: cannot tile n=%d k=%d", n, k);
  }

  PairParams p{};
  p.m = m; p.n = n; p.k = k;
  p.bn = best.bn; p.nsub = best.nsub; p.acc_stages = best.acc_stages; p.stages = best.stages;
  p.out_cols = geglu ? best.bn / 2 : best.bn * best.nsub;
  p.num_m_blocks = (int)((m + 2 * BM - 1) / (2 * BM));
  p.num_k_blocks = (k + BK - 1) / BK;
  p.num_n_blocks = best.num_n_blocks;
  p.tiles = (long long)p.num_m_blocks * p.num_n_blocks;
  p.bias = bias; p.y = y; p.ldy = ldy; p.res = residual; p.ldr = ldr;
  if (ln) {
    p.ln_stats = ln->stats; p.ln_colsum = ln->colsum;
    p.ln_sites = ln->sites; p.ln_frames = ln->frames; p.ln_shift_rows = ln->shift_rows;
  }
  p.slots = residual ? kMaxSlots : 2;
  p.b_bytes = (uint32_t)(p.bn / 2) * BK * 2;
  p.stage_bytes = (uint32_t)best.stage_bytes;
  // instruction descriptor (kind::f16): D=f32 [4,6)=1; A/B format [7,10)/[10,13): 1=bf16, 0=f16; A,B K-major (bits
  // 15,16 = 0); N>>3 at [17,23); M>>4 at [24,29)  (M = 256: the pair's tile)
  const uint32_t fmt = dtype == CA_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
  const long long grid_pairs = pairs < p.tiles ? pairs : p.tiles;
  const size_t smem = (size_t)p.stages * p.stage_bytes + (size_t)kEpiWarps * p.slots * kSlotBytes + 1024;

  CUtensorMap mx, mw, my, mr;
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
    const uint64_t sy[1] = {(uint64_t)ldy * 2}, sr[1] = {(uint64_t)(residual ? ldr : ldy) * 2};
    if (!encode_tensor_map(&my, dt, 2, y, dims, sy, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CA_ERR_CUDA;
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
    CA_CUDA(cudaLaunchKernelEx(&cfg, kernel, mx, mw, my, mr, p));
    if (timing_on) {  // development aid: synchronous, prints the mean per-CTA cycle attribution of this launch
      static long long host[8 * 1024];
      CA_CUDA(cudaStreamSynchronize(st));
      CA_CUDA(cudaMemcpy(host, timing_buf, sizeof(host), cudaMemcpyDeviceToHost));
      double s8[8] = {0};
      const int ctas = (int)(2 * grid_pairs);
      for (int c = 0; c < ctas; ++c)
        for (int j = 0; j < 8; ++j) s8[j] += (double)host[c * 8 + j];
      const double lead = ctas / 2.0;
      fprintf(stderr,
              "[ca_linear timing] m=%lld n=%d k=%d bn=%dx%d stages=%d acc=%d tiles/pair=%.1f | mma loop %.0f clk: wait tmem_empty "
              "%.0f, full %.0f | producer wait empty %.0f | epi wait acc %.0f res %.0f\n",
              m, n, k, p.bn, p.nsub, p.stages, p.acc_stages, (double)p.tiles / (double)grid_pairs, s8[0] / lead, s8[1] / lead,
              s8[2] / lead, s8[3] / ctas, s8[6] / ctas, s8[7] / ctas);
    }
    return CA_OK;
  };
  const int epi = ln ? (geglu ? EPI_LN_GEGLU : EPI_LN) : geglu ? EPI_GEGLU : (residual ? EPI_RES : EPI_BIAS);
  if (dtype == CA_BF16) {
    if (epi == EPI_LN_GEGLU) return run(gemm_pair_kernel<__nv_bfloat16, EPI_LN_GEGLU>);
    if (epi == EPI_LN) return run(gemm_pair_kernel<__nv_bfloat16, EPI_LN>);
    if (epi == EPI_GEGLU) return run(gemm_pair_kernel<__nv_bfloat16, EPI_GEGLU>);
    if (epi == EPI_RES) return run(gemm_pair_kernel<__nv_bfloat16, EPI_RES>);
    return run(gemm_pair_kernel<__nv_bfloat16, EPI_BIAS>);
  }
  if (epi == EPI_LN_GEGLU) return run(gemm_pair_kernel<__half, EPI_LN_GEGLU>);
  if (epi == EPI_LN) return run(gemm_pair_kernel<__half, EPI_LN>);
  if (epi == EPI_GEGLU) return run(gemm_pair_kernel<__half, EPI_GEGLU>);
  if (epi == EPI_RES) return run(gemm_pair_kernel<__half, EPI_RES>);
  return run(gemm_pair_kernel<__half, EPI_BIAS>);
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_linear(const void* x, const void* w, const float* bias,
                                                                const void* residual, void* y, long long m, int n, int k,
                                                                long long ldx, long long ldr, long long ldy, int epilogue,
                                                                int dtype, void* stream) {
  return ca::linear_impl(x, w, bias, residual, y, m, n, k, ldx, ldr, ldy, epilogue, dtype, stream, nullptr);
}

extern "C" __attribute__((visibility("default"))) int ca_linear_ln(const void* x, const void* w_gain, const float* colsum,
                                                                   const float* shift, int shift_rows, int frames, int sites,
                                                                   const float* stats, void* y, long long m, int n, int k,
                                                                   long long ldx, long long ldy, int epilogue, int dtype,
                                                                   void* stream) {
  using namespace ca;
  CA_CHECK_ARG(colsum && shift && stats, "linear_ln: null pointer");
  CA_CHECK_ARG(shift_rows == 1 || (frames >= 1 && shift_rows >= frames && sites >= 32 && sites % 32 == 0),
               "linear_ln: a per-frame shift table needs frames <= shift_rows and sites %% 32 == 0 (got frames=%d rows=%d sites=%d)",
               frames, shift_rows, sites);
  CA_CHECK_ARG((reinterpret_cast<uintptr_t>(stats) & 7) == 0 && aligned16(colsum) && aligned16(shift), "linear_ln: misaligned tables");
  LnFold ln{reinterpret_cast<const float2*>(stats), colsum, sites < 1 ? 1 : sites, frames < 1 ? 1 : frames, shift_rows};
  return linear_impl(x, w_gain, shift, nullptr, y, m, n, k, ldx, ldy, ldy, epilogue, dtype, stream, &ln);
}
