// Kernel (1), fused: one temporal-attention block of the motion module in ONE launch,
//
//   y = x + to_out( softmax( (LN(x) + pe) Wq  ((LN(x) + pe) Wk)^T / sqrt(hd) ) (LN(x) + pe) Wv ) + b_out
//
// for token-major x [(b f d), C] (row t = (b*f + frame)*d + site).  Replaces, per attention block of
// TemporalTransformerBlock.forward (reference animatediff/models/motion_module.py:212-219): nn.LayerNorm (:214),
// VersatileAttention.forward (:272-329: the '(b f) d c -> (b d) f c' rearrange, the positional-encoding add :287-288, the
// processor call :321, the rearrange back :327), the AttentionProcessor arithmetic (modules/attention_processor.py:
// 186-272: to_q / to_k / to_v, head split, softmax(q k^T scale) v, head merge, to_out[0]) and the residual add (:219).
// Round 1 ran this as four launches (LayerNorm+PE, QKV GEMM, attention core, out-proj GEMM) moving ~13 T*C*s bytes; here
// x is read once and y written once (2 T*C*s + weights from L2), SURVEY.md §8(d) row "(1) fused".
//
// A CTA PAIR (cluster of 2, tcgen05 cta_group::2) owns 2 x 128 rows = 2 x (S sites x f frames), rows ordered frame-major
// inside a CTA (r = frame * S + site: what one 3-D TMA box (64 channels, S sites, f frames) delivers, SWIZZLE_128B = the
// UMMA K-major operand layout).  Shared memory per CTA (C = 320): R0 80 KB (x -> LayerNorm+PE in place = A operand of the
// QKV projection; later the residual x again and the output staging), R1 80 KB (attention output = A operand of the
// out-projection), a 3 x 10 KB weight ring (each CTA loads HALF of every weight tile, the tensor core reads the other half
// from the peer), 30 KB of per-site Q / K / V matrices.  TMEM: two 128-column stages for one head's q|k|v block, reused as
// the 320 accumulator columns of the out-projection.
//
// Roles: warps 0-15 compute (LayerNorm, TMEM drain, attention on mma.sync with warp-shuffle softmax exactly as
// temporal_attn.cu — 16 / S warps share a site and split the P V column groups —, epilogue), warp 16 TMA producer,
// warp 17 MMA issuer (leader CTA), warp 18 TMEM allocator.
// Per tile:  x --TMA--> R0 --LN+PE--> A;  for each head: UMMA(A, Wqkv_h) -> TMEM -> bf16 Q,K,V per site -> f x f attention
// -> O_h into R1 (while the tensor core already works on head h+1);  x --TMA--> R0 again (residual);  UMMA(R1, Wo) -> TMEM
// -> + b_out + x -> R0 --TMA--> y.
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int BM = 128, BK = 64, UMMA_K = 16;
constexpr int kComputeWarps = 16, kComputeThreads = 32 * kComputeWarps;
constexpr int kProducerWarp = 16, kMmaWarp = 17, kAllocWarp = 18;
constexpr int kThreads = 19 * 32;
constexpr int kWSlots = 3;
constexpr uint32_t kChunkBytes = BM * BK * 2;  // [128 rows][64 channels] of 16-bit elements, 128-byte rows

struct FusedParams {
  int b, f, d, heads, hd, C;
  int S;           // sites per CTA (S * f <= 128); a pair covers 2 S sites
  int site_tiles;  // ceil(d / (2 S))
  int tiles;       // b * site_tiles
  int nq;          // columns of one head's q|k|v block: 3 hd rounded up to a multiple of 16
  int bn_o, nsub_o;  // out-projection: nsub_o passes of bn_o accumulator columns
  int fpad, pitch, site_bytes;  // per-site staging: 3 matrices [fpad][pitch bytes] (+16 bytes: bank spread)
  uint32_t x_bytes;             // bytes one x box delivers (64 * S * f * 2)
  uint32_t wq_bytes, wo_bytes;  // this CTA's half of one weight k-chunk
  uint32_t slot_bytes;
  uint32_t idesc_q, idesc_o;
  float eps, scale_log2;
  const float *gamma, *beta, *pe, *bo;  // pe: [>= f, C] fp32 or null
  long long* timing;                    // CA_FUSED_TIMING=1: cycle counters of pair 0's leader CTA (development aid)
};

// ---- tcgen05 (cta_group::2) -------------------------------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit_pair(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void umma_f16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
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
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// K-major SWIZZLE_128B smem matrix descriptor (see gemm_pair_tcgen05.cu)
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// explicit shared-space accesses: the regions are carved out of an integer-aligned base, so plain pointer accesses compile
// to generic LD / ST (measured: the generic ST.E.128 of the drain were among the hottest instructions)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 r;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(a));
  return r;
}
__device__ __forceinline__ void sts128(uint32_t a, const uint4& v) {
  asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(a), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void sts32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ float2 lds64f(uint32_t a) {
  float2 r;
  asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(r.x), "=f"(r.y) : "r"(a));
  return r;
}
__device__ __forceinline__ void compute_bar() { asm volatile("bar.sync 1, %0;" ::"n"(kComputeThreads) : "memory"); }

// byte offset of 16-byte unit u (8 channels) of row r inside a region of 64-channel chunks (SWIZZLE_128B)
__device__ __forceinline__ uint32_t unit_off(int r, int u) {
  return (uint32_t)(u >> 3) * kChunkBytes + (uint32_t)r * 128u + (uint32_t)(((u & 7) ^ (r & 7)) << 4);
}

// ---- mma.sync helpers for the f x f attention (same arithmetic as temporal_attn.cu) -------------------------------------
template <typename T>
struct MmaT;
template <>
struct MmaT<__nv_bfloat16> {
  __device__ static void k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
};
template <>
struct MmaT<__half> {
  __device__ static void k16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  }
  __device__ static void k8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t b0) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(b0));
  }
};
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1(uint32_t& r0, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// One (site, head): O = softmax(Q K^T * scale) V for the f frames of the site; Q, K, V are [fpad][pitch] matrices in smem and
// O is written over Q by the caller's copy loop (returned through `obuf`, this warp's column groups only).  MT = ceil(f / 16)
// query tiles; HD > 0 fixes head_dim at compile time (the generic code is issue-bound on loop and index arithmetic,
// profiles/r01d_notes.md §2).  `wps` warps work on the same site: each computes the scores and the softmax (cheap) and
// the P V product of the 8-column groups g with g % wps == part.  Warp-collective.
template <typename T, int MT, int HD>
__device__ __forceinline__ void attend_site(unsigned char* qbase, uint32_t mat_bytes, int pitch, int f, int hd_rt, float scale_log2, int lane,
                                            int wps, int part, uint32_t R1s, int S, int site, int unit0) {
  using M = MmaT<T>;
  constexpr int NT = 2 * MT;
  const int hd = HD > 0 ? HD : hd_rt;
  const uint32_t q_s = smem_u32(qbase), k_s = q_s + mat_bytes, v_s = q_s + 2 * mat_bytes;
  const int r0 = lane >> 2, cq = (lane & 3) * 2;
  float sacc[MT][NT][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[mt][nt][j] = 0.f;
  const int k16 = hd >> 4;
#pragma unroll
  for (int ks = 0; ks < (HD > 0 ? HD / 16 : 16); ++ks) {
    if (HD == 0 && ks >= k16) break;
    uint32_t a[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) ldsm_x4(a[mt], q_s + (mt * 16 + (lane & 15)) * pitch + (ks * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < f) {
        uint32_t b0, b1;
        ldsm_x2(b0, b1, k_s + (nt * 8 + (lane & 7)) * pitch + (ks * 16 + ((lane >> 3) & 1) * 8) * 2);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) M::k16(sacc[mt][nt], a[mt], b0, b1);
      }
    }
  }
  if (hd & 8) {
    const int c0 = k16 * 16;
    uint32_t a[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) ldsm_x2(a[mt][0], a[mt][1], q_s + (mt * 16 + (lane & 15)) * pitch + c0 * 2);
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 < f) {
        uint32_t b0;
        ldsm_x1(b0, k_s + (nt * 8 + (lane & 7)) * pitch + c0 * 2);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) M::k8(sacc[mt][nt], a[mt][0], a[mt][1], b0);
      }
    }
  }
  // softmax over keys (fp32, exp2 with folded scale); P -> hi + lo 16-bit fragments (P V then carries ~16 mantissa bits)
  uint32_t p_hi[MT][NT][2], p_lo[MT][NT][2];
  float inv_sum[MT][2];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int key = nt * 8 + cq + (j & 1);
        const float sv = key < f ? sacc[mt][nt][j] * scale_log2 : -INFINITY;
        sacc[mt][nt][j] = sv;
        mx[j >> 1] = fmaxf(mx[j >> 1], sv);
      }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      float e[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        e[j] = exp2f(sacc[mt][nt][j] - mx[j >> 1]);
        sum[j >> 1] += e[j];
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const uint32_t hi = pack2(e[2 * h], e[2 * h + 1], T());
        float h0, h1;
        unpack2(hi, h0, h1, T());
        p_hi[mt][nt][h] = hi;
        p_lo[mt][nt][h] = pack2(e[2 * h] - h0, e[2 * h + 1] - h1, T());
      }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
      inv_sum[mt][h] = 1.0f / sum[h];
    }
  }
  // O = P V for this warp's 8-column groups, straight into the rows of R1 (A operand of the out-projection): the fragment
  // holds two adjacent columns of rows r0 / r0 + 8 -> 4-byte stores
#pragma unroll
  for (int g = 0; g < (HD > 0 ? HD / 8 : 32); ++g) {
    if (HD == 0 && g * 8 >= hd) break;
    if (g % wps != part) continue;
    const int c0 = g * 8;
    float oacc[MT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int j = 0; j < 4; ++j) oacc[mt][j] = 0.f;
#pragma unroll
    for (int kt = 0; kt < MT; ++kt) {
      if (kt * 16 < f) {
        uint32_t b0, b1;
        ldsm_x2_trans(b0, b1, v_s + (kt * 16 + (lane & 15)) * pitch + c0 * 2);
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
          const uint32_t ah[4] = {p_hi[mt][2 * kt][0], p_hi[mt][2 * kt][1], p_hi[mt][2 * kt + 1][0], p_hi[mt][2 * kt + 1][1]};
          const uint32_t al[4] = {p_lo[mt][2 * kt][0], p_lo[mt][2 * kt][1], p_lo[mt][2 * kt + 1][0], p_lo[mt][2 * kt + 1][1]};
          M::k16(oacc[mt], ah, b0, b1);
          M::k16(oacc[mt], al, b0, b1);
        }
      }
    }
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int fr = mt * 16 + r0 + h * 8;
        if (fr < f)
          sts32(R1s + unit_off(fr * S + site, unit0 + g) + cq * 2, pack2(oacc[mt][2 * h] * inv_sum[mt][h], oacc[mt][2 * h + 1] * inv_sum[mt][h], T()));
      }
  }
}

template <typename T, int CHUNKS>
__global__ void __launch_bounds__(kThreads, 1)
    temporal_block_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_y,
                          const __grid_constant__ CUtensorMap map_wq, const __grid_constant__ CUtensorMap map_wo, const FusedParams p) {
  constexpr int UQ = 2 * CHUNKS;  // 16-byte units per quarter row (LayerNorm statistics: four threads per row)
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t x_full, res_full, r0_free, a_ready, a_done, o_ready, y_full, y_free;
  __shared__ uint64_t w_full[kWSlots], w_empty[kWSlots], qkv_full[2], qkv_empty[2];
  __shared__ uint32_t tmem_base_slot;

  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  unsigned char* R0 = smem;
  unsigned char* R1 = R0 + CHUNKS * kChunkBytes;
  unsigned char* wring = R1 + CHUNKS * kChunkBytes;
  unsigned char* stg = wring + kWSlots * p.slot_bytes;                           // per-site Q | K | V matrices (+ tail pad)
  float* bo_s = reinterpret_cast<float*>(stg + (size_t)p.S * p.site_bytes + 32 * (size_t)p.pitch);  // [C] out-projection bias

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, pairs = gridDim.x >> 1;

  // rows the TMA boxes never touch (S * f < 128) and the staging pads must hold finite data
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (int)((2 * CHUNKS * kChunkBytes + kWSlots * p.slot_bytes + (size_t)p.S * p.site_bytes + 32 * (size_t)p.pitch) / 16);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    for (int i = threadIdx.x; i < p.C; i += blockDim.x) bo_s[i] = p.bo[i];
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    mbar_init(&x_full, 1);
    mbar_init(&res_full, 1);
    mbar_init(&r0_free, 1);
    mbar_init(&a_ready, 2);   // one arrival per CTA of the pair (on the leader)
    mbar_init(&a_done, 1);    // multicast tcgen05.commit
    mbar_init(&o_ready, 2);
    mbar_init(&y_full, 1);
    mbar_init(&y_free, 2 * kComputeWarps);
    for (int s = 0; s < kWSlots; ++s) {
      mbar_init(&w_full[s], 1);
      mbar_init(&w_empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&qkv_full[s], 1);
      mbar_init(&qkv_empty[s], 2 * kComputeWarps);
    }
    fence_mbar_init();
  }
  if (warp == kAllocWarp) tmem_alloc_pair(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;
  const int S = p.S, f = p.f;

  auto tile_coords = [&](int tile, int& site0, int& frame0) {
    const int bi = tile / p.site_tiles, st = tile - bi * p.site_tiles;
    site0 = st * 2 * S + (int)rank * S;
    frame0 = bi * f;
  };

  if (warp == kProducerWarp) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tensormap(&map_x);
      prefetch_tensormap(&map_wq);
      prefetch_tensormap(&map_wo);
    }
    int wslot = 0;
    uint32_t wphase = 0;
    auto load_w = [&](const CUtensorMap* map, int k0, int row0, uint32_t bytes) {
      mbar_wait(&w_empty[wslot], wphase ^ 1);
      const uint32_t bar = mapa_u32(&w_full[wslot], 0);
      if (elect_one()) {
        if (leader) mbar_arrive_expect_tx(&w_full[wslot], 2u * bytes);  // bytes of both CTAs are credited to the leader
        tma_load_2d_pair(wring + (size_t)wslot * p.slot_bytes, map, bar, k0, row0);
      }
      __syncwarp();
      if (++wslot == kWSlots) {
        wslot = 0;
        wphase ^= 1;
      }
    };
    int it = 0;
    for (int tile = pair; tile < p.tiles; tile += pairs, ++it) {
      int site0, frame0;
      tile_coords(tile, site0, frame0);
      mbar_wait(&r0_free, (uint32_t)(it & 1) ^ 1);  // the previous tile's output has left R0
      if (elect_one()) {
        mbar_arrive_expect_tx(&x_full, (uint32_t)CHUNKS * p.x_bytes);
        for (int c = 0; c < CHUNKS; ++c) tma_load_3d(R0 + c * kChunkBytes, &map_x, &x_full, c * BK, site0, frame0);
      }
      __syncwarp();
      for (int h = 0; h < p.heads; ++h)
        for (int c = 0; c < CHUNKS; ++c) load_w(&map_wq, c * BK, h * p.nq + (int)rank * (p.nq / 2), p.wq_bytes);
      mbar_wait(&a_done, (uint32_t)(it & 1));  // the QKV MMAs have read the LayerNorm'd tile: R0 takes the residual x
      if (elect_one()) {
        mbar_arrive_expect_tx(&res_full, (uint32_t)CHUNKS * p.x_bytes);
        for (int c = 0; c < CHUNKS; ++c) tma_load_3d(R0 + c * kChunkBytes, &map_x, &res_full, c * BK, site0, frame0);
      }
      __syncwarp();
      for (int ps = 0; ps < p.nsub_o; ++ps)
        for (int c = 0; c < CHUNKS; ++c) load_w(&map_wo, c * BK, ps * p.bn_o + (int)rank * (p.bn_o / 2), p.wo_bytes);
    }
  } else if (warp == kMmaWarp) {
    // ===================== MMA issuer (leader CTA) =====================
    if (leader) {
      int wslot = 0, it = 0;
      uint32_t wphase = 0, nq_cnt = 0;
      const bool tm = p.timing != nullptr && blockIdx.x == 0;
      long long tw_w = 0, tw_a = 0, tw_q = 0, tw_o = 0, t0m = 0;
      const long long t_begin = tm ? clock64() : 0;
#define TWAIT(acc, stmt) do { if (tm) t0m = clock64(); stmt; if (tm) acc += clock64() - t0m; } while (0)
      // one k-chunk: 4 UMMAs of K = 16 on A chunk `a_addr` and the weight slot; D accumulates from the first k-step
      auto chunk_mmas = [&](uint32_t a_addr, uint32_t tmem_d, uint32_t idesc, int c) {
        TWAIT(tw_w, mbar_wait(&w_full[wslot], wphase));
        tc_fence_after();
        const uint64_t adesc = make_sw128_desc(a_addr), bdesc = make_sw128_desc(smem_u32(wring + (size_t)wslot * p.slot_bytes));
        const uint32_t acc0 = c != 0 ? 1u : 0u;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks)
            umma_f16_pair(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), idesc, ks ? 1u : acc0);
          umma_commit_pair(&w_empty[wslot]);
        }
        __syncwarp();
        if (++wslot == kWSlots) {
          wslot = 0;
          wphase ^= 1;
        }
      };
      for (int tile = pair; tile < p.tiles; tile += pairs, ++it) {
        TWAIT(tw_a, mbar_wait(&a_ready, (uint32_t)(it & 1)));
        mbar_wait(&y_free, (uint32_t)(it & 1) ^ 1);  // the previous tile's out-projection accumulators have been drained
        tc_fence_after();
        for (int h = 0; h < p.heads; ++h, ++nq_cnt) {
          const uint32_t st = nq_cnt & 1;
          TWAIT(tw_q, mbar_wait(&qkv_empty[st], ((nq_cnt >> 1) & 1) ^ 1));
          tc_fence_after();
          for (int c = 0; c < CHUNKS; ++c) chunk_mmas(smem_u32(R0 + c * kChunkBytes), tmem_base + st * (uint32_t)p.nq, p.idesc_q, c);
          if (elect_one()) {
            umma_commit_pair(&qkv_full[st]);
            if (h == p.heads - 1) umma_commit_pair(&a_done);
          }
          __syncwarp();
        }
        TWAIT(tw_o, mbar_wait(&o_ready, (uint32_t)(it & 1)));
        tc_fence_after();
        for (int ps = 0; ps < p.nsub_o; ++ps)
          for (int c = 0; c < CHUNKS; ++c) chunk_mmas(smem_u32(R1 + c * kChunkBytes), tmem_base + (uint32_t)(ps * p.bn_o), p.idesc_o, c);
        if (elect_one()) umma_commit_pair(&y_full);
        __syncwarp();
      }
      if (tm && lane == 0) {
        p.timing[0] = clock64() - t_begin; p.timing[1] = tw_a; p.timing[2] = tw_w; p.timing[3] = tw_q; p.timing[4] = tw_o;
      }
#undef TWAIT
    }
  } else if (warp < kComputeWarps) {
    // ===================== compute warps =====================
    const int ctid = threadIdx.x;
    const int q = warp & 3, sub = warp >> 2;       // TMEM lane quarter / column quarter
    const int trow = q * 32 + lane;                // this lane's tile row in TMEM
    const int tframe = trow / S, tsite = trow - tframe * S;
    const bool trow_ok = trow < S * f;
    const uint32_t lane_off = (uint32_t)(q * 32) << 16;
    const uint32_t a_ready_l = mapa_u32(&a_ready, 0), o_ready_l = mapa_u32(&o_ready, 0), y_free_l = mapa_u32(&y_free, 0);
    const uint32_t qe_l[2] = {mapa_u32(&qkv_empty[0], 0), mapa_u32(&qkv_empty[1], 0)};
    const uint32_t mat_bytes = (uint32_t)(p.fpad * p.pitch);
    const uint32_t R0s = smem_u32(R0), R1s = smem_u32(R1), stgs = smem_u32(stg);
    // drain: the 8-column groups sub, sub + 4, ... of a head's q|k|v block; their staging offsets are fixed for the kernel
    const int n_groups = 3 * p.hd / 8;
    int dr_off[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int cg = (sub + 4 * j) * 8, which = cg / p.hd, cc = cg - which * p.hd;
      dr_off[j] = tsite * p.site_bytes + which * (int)mat_bytes + tframe * p.pitch + cc * 2;
    }
    // attention: 16 / S warps per site (they split the P V column groups)
    const int wps = kComputeWarps / S > 0 ? kComputeWarps / S : 1;
    const int att_site = warp / wps, att_part = warp - att_site * wps;
    uint32_t nq_cnt = 0;
    int it = 0;
    if (ctid == 0) prefetch_tensormap(&map_y);
#ifdef CA_FUSED_PHASE_TIMING  // development build only: the counters cost ~20 registers in the hot warps
    const bool tmc = p.timing != nullptr && blockIdx.x == 0 && ctid == 0;
    long long tc_x = 0, tc_ln = 0, tc_qw = 0, tc_drain = 0, tc_att = 0, tc_ew = 0, tc_epi = 0, tc_st = 0, tc0 = 0;
#define TC(acc) do { if (tmc) { const long long n_ = clock64(); acc += n_ - tc0; tc0 = n_; } } while (0)
#else
#define TC(acc) do { } while (0)
#endif
    // LayerNorm pass 2: this thread's channel unit and frame group; gamma / beta of the unit stay in registers
    constexpr int UR = 8 * CHUNKS, NFG = kComputeThreads / UR;
    const int ln_u = ctid % UR, ln_fg = ctid / UR;
    const float4 lg0 = __ldg(reinterpret_cast<const float4*>(p.gamma + ln_u * 8)), lg1 = __ldg(reinterpret_cast<const float4*>(p.gamma + ln_u * 8 + 4));
    const float4 lb0 = __ldg(reinterpret_cast<const float4*>(p.beta + ln_u * 8)), lb1 = __ldg(reinterpret_cast<const float4*>(p.beta + ln_u * 8 + 4));
    for (int tile = pair; tile < p.tiles; tile += pairs, ++it) {
      int site0, frame0;
      tile_coords(tile, site0, frame0);
      // ---------- LayerNorm + positional encoding, in place on R0 ----------
      // pass 1 (two threads per row): mean / rstd of every row -> a 1 KB table.  Sums are shifted by the half
      // row's first element and the halves combined with Chan's formula: no cancellation, one read.
      // pass 2 (one 8-channel unit per thread, fixed for the whole kernel: gamma / beta live in registers; the thread walks
      // the frames of its group, so the positional-encoding row is fetched once per frame): normalise in place.
#ifdef CA_FUSED_PHASE_TIMING
      if (tmc) tc0 = clock64();
#endif
      mbar_wait(&x_full, (uint32_t)(it & 1));
      TC(tc_x);
      float2* stats = reinterpret_cast<float2*>(bo_s + p.C);  // [128] (mean, rstd); its own 1 KB: the staging pads must stay zero
      const uint32_t stats_s = smem_u32(stats);
      {
        const int row = ctid >> 2, qd = ctid & 3;
        float pivot;
        {
          Vec16<T> t;
          t.raw = lds128(R0s + unit_off(row, qd * UQ));
          float v[8];
          t.unpack(v);
          pivot = v[0];
        }
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int j = 0; j < UQ; ++j) {
          float v[8];
          Vec16<T> t;
          t.raw = lds128(R0s + unit_off(row, qd * UQ + j));
          t.unpack(v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const float dlt = v[e] - pivot;
            s1 += dlt;
            s2 = fmaf(dlt, dlt, s2);
          }
        }
        float n = (float)(8 * UQ);
        float mean = pivot + s1 / n, m2 = s2 - s1 * s1 / n;
#pragma unroll
        for (int o = 1; o <= 2; o <<= 1) {  // Chan's pairwise combine of (n, mean, M2) over the four quarter rows
          const float mean_o = __shfl_xor_sync(0xffffffffu, mean, o), m2_o = __shfl_xor_sync(0xffffffffu, m2, o);
          const float dm = mean_o - mean;
          m2 = m2 + m2_o + dm * dm * (0.5f * n);
          mean = 0.5f * (mean + mean_o);
          n *= 2.0f;
        }
        if (qd == 0) stats[row] = make_float2(mean, rsqrtf(fmaxf(m2 / n, 0.f) + p.eps));
      }
      compute_bar();
      if (ctid < UR * NFG) {
        for (int fr = ln_fg; fr < f; fr += NFG) {
          float4 e0 = make_float4(0.f, 0.f, 0.f, 0.f), e1 = e0;
          if (p.pe) {
            e0 = __ldg(reinterpret_cast<const float4*>(p.pe + (size_t)fr * p.C + ln_u * 8));
            e1 = __ldg(reinterpret_cast<const float4*>(p.pe + (size_t)fr * p.C + ln_u * 8 + 4));
          }
          const float be[8] = {lb0.x + e0.x, lb0.y + e0.y, lb0.z + e0.z, lb0.w + e0.w, lb1.x + e1.x, lb1.y + e1.y, lb1.z + e1.z, lb1.w + e1.w};
#pragma unroll 4
          for (int sI = 0; sI < S; ++sI) {
            const int row = fr * S + sI;
            const uint32_t up = R0s + unit_off(row, ln_u);
            const float2 ms = lds64f(stats_s + row * 8);
            Vec16<T> t;
            t.raw = lds128(up);
            float v[8];
            t.unpack(v);
            v[0] = fmaf((v[0] - ms.x) * ms.y, lg0.x, be[0]);
            v[1] = fmaf((v[1] - ms.x) * ms.y, lg0.y, be[1]);
            v[2] = fmaf((v[2] - ms.x) * ms.y, lg0.z, be[2]);
            v[3] = fmaf((v[3] - ms.x) * ms.y, lg0.w, be[3]);
            v[4] = fmaf((v[4] - ms.x) * ms.y, lg1.x, be[4]);
            v[5] = fmaf((v[5] - ms.x) * ms.y, lg1.y, be[5]);
            v[6] = fmaf((v[6] - ms.x) * ms.y, lg1.z, be[6]);
            v[7] = fmaf((v[7] - ms.x) * ms.y, lg1.w, be[7]);
            t.pack(v);
            sts128(up, t.raw);
          }
        }
      }
      fence_proxy_async();  // generic-proxy writes of the A operand -> visible to the tensor core
      compute_bar();
      if (ctid == 0) {
        if (leader) mbar_arrive(&a_ready);
        else mbar_arrive_cluster(a_ready_l);
      }
      TC(tc_ln);
      // ---------- heads: drain q|k|v of head h, f x f attention per site, O_h into R1 ----------
      for (int h = 0; h < p.heads; ++h, ++nq_cnt) {
        const uint32_t st = nq_cnt & 1;
        mbar_wait(&qkv_full[st], (nq_cnt >> 1) & 1);
        TC(tc_qw);
        tc_fence_after();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const int g = sub + 4 * j;
          if (g < n_groups) {
            uint32_t a[8];
            tmem_ld8(tmem_base + lane_off + st * (uint32_t)p.nq + (uint32_t)(g * 8), a);
            tmem_ld_wait();
            if (trow_ok) {
              uint4 o;
              o.x = pack2(__uint_as_float(a[0]), __uint_as_float(a[1]), T());
              o.y = pack2(__uint_as_float(a[2]), __uint_as_float(a[3]), T());
              o.z = pack2(__uint_as_float(a[4]), __uint_as_float(a[5]), T());
              o.w = pack2(__uint_as_float(a[6]), __uint_as_float(a[7]), T());
              sts128(stgs + dr_off[j], o);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {  // this warp's share of the stage is in shared memory: hand the TMEM stage back
          if (leader) mbar_arrive(&qkv_empty[st]);
          else mbar_arrive_cluster_relaxed(qe_l[st]);
        }
        compute_bar();  // Q, K, V of every site are complete
        TC(tc_drain);
        for (int sI = att_site; sI < S; sI += kComputeWarps / wps) {
          unsigned char* qb = stg + (size_t)sI * p.site_bytes;
          const int unit0 = h * p.hd >> 3;
          if (f > 16) {
            if (p.hd == 40) attend_site<T, 2, 40>(qb, mat_bytes, p.pitch, f, 40, p.scale_log2, lane, wps, att_part, R1s, S, sI, unit0);
            else attend_site<T, 2, 0>(qb, mat_bytes, p.pitch, f, p.hd, p.scale_log2, lane, wps, att_part, R1s, S, sI, unit0);
          } else {
            if (p.hd == 40) attend_site<T, 1, 40>(qb, mat_bytes, p.pitch, f, 40, p.scale_log2, lane, wps, att_part, R1s, S, sI, unit0);
            else attend_site<T, 1, 0>(qb, mat_bytes, p.pitch, f, p.hd, p.scale_log2, lane, wps, att_part, R1s, S, sI, unit0);
          }
        }
        compute_bar();  // the staging matrices are free for the next head
        TC(tc_att);
      }
      fence_proxy_async();
      compute_bar();
      if (ctid == 0) {
        if (leader) mbar_arrive(&o_ready);
        else mbar_arrive_cluster(o_ready_l);
      }
      // ---------- epilogue: y = acc + b_out + x, staged in R0 (which holds the residual x again), TMA store ----------
      mbar_wait(&res_full, (uint32_t)(it & 1));
      mbar_wait(&y_full, (uint32_t)(it & 1));
      TC(tc_ew);
      tc_fence_after();
      {
        const int half_c = p.C >> 2;  // this warp's quarter of the output columns
        for (int cb = 0; cb < half_c; cb += 16) {
          const int col0 = sub * half_c + cb;
          uint32_t a[16];
          tmem_ld16(tmem_base + lane_off + (uint32_t)col0, a);
          tmem_ld_wait();
          if (cb + 16 >= half_c) {  // last read of this warp: release the accumulator columns
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (leader) mbar_arrive(&y_free);
              else mbar_arrive_cluster_relaxed(y_free_l);
            }
          }
#pragma unroll
          for (int g = 0; g < 2; ++g) {
            const int c0 = col0 + 8 * g;
            const uint32_t up = R0s + unit_off(trow, c0 >> 3);
            Vec16<T> t;
            t.raw = lds128(up);
            float r[8];
            t.unpack(r);
            const float4 b0 = *reinterpret_cast<const float4*>(bo_s + c0), b1 = *reinterpret_cast<const float4*>(bo_s + c0 + 4);
            r[0] += __uint_as_float(a[8 * g + 0]) + b0.x;
            r[1] += __uint_as_float(a[8 * g + 1]) + b0.y;
            r[2] += __uint_as_float(a[8 * g + 2]) + b0.z;
            r[3] += __uint_as_float(a[8 * g + 3]) + b0.w;
            r[4] += __uint_as_float(a[8 * g + 4]) + b1.x;
            r[5] += __uint_as_float(a[8 * g + 5]) + b1.y;
            r[6] += __uint_as_float(a[8 * g + 6]) + b1.z;
            r[7] += __uint_as_float(a[8 * g + 7]) + b1.w;
            t.pack(r);
            sts128(up, t.raw);
          }
        }
      }
      fence_proxy_async();
      compute_bar();
      TC(tc_epi);
      if (ctid == 0) {
        for (int c = 0; c < CHUNKS; ++c) tma_store_3d(&map_y, R0 + c * kChunkBytes, c * BK, site0, frame0);
        bulk_commit();
        bulk_wait_read<0>();  // R0 may be refilled (the global writes complete asynchronously)
        mbar_arrive(&r0_free);
      }
      TC(tc_st);
    }
    if (ctid == 0) bulk_wait<0>();
#ifdef CA_FUSED_PHASE_TIMING
    if (tmc) {
      p.timing[8] = tc_x; p.timing[9] = tc_ln; p.timing[10] = tc_qw; p.timing[11] = tc_drain; p.timing[12] = tc_att;
      p.timing[13] = tc_ew; p.timing[14] = tc_epi; p.timing[15] = tc_st; p.timing[16] = it;
    }
#endif
#undef TC
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == kAllocWarp) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_temporal_attn_fused(const void* x, void* y, const float* ln_gamma,
                                                                             const float* ln_beta, const float* pe,
                                                                             const void* wqkv_perm, const void* wo,
                                                                             const float* bo, int b, int f, int d, int C,
                                                                             int heads, float eps, float scale, int dtype,
                                                                             void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && ln_gamma && ln_beta && wqkv_perm && wo && bo, "temporal_attn_fused: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "temporal_attn_fused: dtype must be bf16 or f16");
  CA_CHECK_ARG(b > 0 && d > 0 && heads > 0 && f >= 1 && f <= 32, "temporal_attn_fused: bad sizes (f <= 32: PE max_len)");
  if (!(C == 64 || C == 128 || C == 320) || C % heads != 0) {
    set_error("temporal_attn_fused: C=%d is not a built variant (64, 128, 320: the LayerNorm'd tile and the attention output of 128 "
              "rows must both fit shared memory)", C);
    return CA_ERR_UNSUPPORTED;
  }
  const int hd = C / heads;
  CA_CHECK_ARG(hd % 8 == 0 && 3 * hd <= 256, "temporal_attn_fused: head_dim=%d must be a multiple of 8 and <= 80", hd);
  CA_CHECK_ARG(aligned16(x) && aligned16(y) && aligned16(wqkv_perm) && aligned16(wo) && aligned16(ln_gamma) && aligned16(ln_beta) &&
                   aligned16(bo) && (!pe || aligned16(pe)),
               "temporal_attn_fused: pointers must be 16-byte aligned");
  CA_CHECK_ARG((long long)b * f * d < (1ll << 31), "temporal_attn_fused: too many rows");

  FusedParams p{};
  p.b = b; p.f = f; p.d = d; p.heads = heads; p.hd = hd; p.C = C;
  p.S = BM / f;
  p.site_tiles = (d + 2 * p.S - 1) / (2 * p.S);
  p.tiles = b * p.site_tiles;
  p.nq = (3 * hd + 15) / 16 * 16;
  p.nsub_o = C > 256 ? 2 : 1;
  p.bn_o = C / p.nsub_o;
  CA_CHECK_ARG(p.bn_o % 16 == 0 && 2 * p.nq <= 512, "temporal_attn_fused: cannot tile C=%d", C);
  p.fpad = (f + 7) / 8 * 8;  // ldmatrix may read up to 16 / 32 rows: what lies behind is another (finite) matrix or the tail pad,
                             // and those rows only ever meet zero probabilities or discarded query rows
  const int hdp = ((hd / 8) & 1) ? hd : hd + 8;          // odd number of 16-byte units per row: conflict-free ldmatrix
  p.pitch = hdp * 2;
  p.site_bytes = 3 * p.fpad * p.pitch + 16;
  p.x_bytes = (uint32_t)(BK * p.S * f * 2);
  p.wq_bytes = (uint32_t)(p.nq / 2) * BK * 2;
  p.wo_bytes = (uint32_t)(p.bn_o / 2) * BK * 2;
  p.slot_bytes = ((p.wq_bytes > p.wo_bytes ? p.wq_bytes : p.wo_bytes) + 1023) & ~1023u;
  const uint32_t fmt = dtype == CA_BF16 ? 1u : 0u;
  p.idesc_q = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.nq >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
  p.idesc_o = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.bn_o >> 3) << 17) | ((uint32_t)((2 * BM) >> 4) << 24);
  p.eps = eps;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.gamma = ln_gamma; p.beta = ln_beta; p.pe = pe; p.bo = bo;
  const int chunks = C / BK;
  const size_t smem = (size_t)2 * chunks * kChunkBytes + (size_t)kWSlots * p.slot_bytes + (size_t)p.S * p.site_bytes + 32 * (size_t)p.pitch + (size_t)C * 4 + BM * 8 + 1024;
  CA_CHECK_ARG(smem <= 227 * 1024 - 1024, "temporal_attn_fused: tile does not fit shared memory (%zu bytes)", smem);

  const CUtensorMapDataType dt = dtype == CA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap mx, my, mq, mo;
  {
    // (channel, site, frame-of-batch): site stride C, frame stride d * C elements
    const uint64_t dims[3] = {(uint64_t)C, (uint64_t)d, (uint64_t)b * f};
    const uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)d * C * 2};
    const uint32_t box[3] = {BK, (uint32_t)p.S, (uint32_t)f};
    if (!encode_tensor_map(&mx, dt, 3, x, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CA_ERR_CUDA;
    if (!encode_tensor_map(&my, dt, 3, y, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return CA_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)heads * p.nq};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {BK, (uint32_t)(p.nq / 2)};
    if (!encode_tensor_map(&mq, dt, 2, wqkv_perm, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return CA_ERR_CUDA;
  }
  {
    const uint64_t dims[2] = {(uint64_t)C, (uint64_t)C};
    const uint64_t strides[1] = {(uint64_t)C * 2};
    const uint32_t box[2] = {BK, (uint32_t)(p.bn_o / 2)};
    if (!encode_tensor_map(&mo, dt, 2, wo, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return CA_ERR_CUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static long long* timing_buf = nullptr;
  static const bool timing_on = getenv("CA_FUSED_TIMING") != nullptr;
  if (timing_on) {
    if (!timing_buf) CA_CUDA(cudaMalloc(&timing_buf, 32 * sizeof(long long)));
    CA_CUDA(cudaMemsetAsync(timing_buf, 0, 32 * sizeof(long long), st));
    p.timing = timing_buf;
  }
  const int pairs_hw = sm_count() / 2;
  const int grid_pairs = pairs_hw < p.tiles ? pairs_hw : p.tiles;
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
    CA_CUDA(cudaLaunchKernelEx(&cfg, kernel, mx, my, mq, mo, p));
    if (timing_on) {
      long long h[32];
      CA_CUDA(cudaStreamSynchronize(st));
      CA_CUDA(cudaMemcpy(h, timing_buf, sizeof(h), cudaMemcpyDeviceToHost));
      fprintf(stderr, "[fused timing] b%d f%d d%d C%d tiles/pair %.1f (%lld on pair 0) | mma loop %lld clk: wait a_ready %lld, weights %lld, qkv_empty %lld, "
              "o_ready %lld | compute: wait x %lld, LN %lld, wait qkv %lld, drain %lld, attention %lld, wait res+y %lld, epilogue %lld, store %lld\n",
              b, f, d, C, (double)p.tiles / grid_pairs, h[16], h[0], h[1], h[2], h[3], h[4], h[8], h[9], h[10], h[11], h[12], h[13], h[14], h[15]);
    }
    return CA_OK;
  };
  if (dtype == CA_BF16) {
    if (chunks == 5) return run(temporal_block_kernel<__nv_bfloat16, 5>);
    if (chunks == 2) return run(temporal_block_kernel<__nv_bfloat16, 2>);
    return run(temporal_block_kernel<__nv_bfloat16, 1>);
  }
  if (chunks == 5) return run(temporal_block_kernel<__half, 5>);
  if (chunks == 2) return run(temporal_block_kernel<__half, 2>);
  return run(temporal_block_kernel<__half, 1>);
}
