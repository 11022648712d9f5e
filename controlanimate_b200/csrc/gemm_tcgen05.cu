// Dense projections of the motion module on the 5th-generation tensor cores.
//
//   y[m, n_out] = epilogue( x[m, k] @ w[n, k]^T + bias[n] ) (+ residual[m, n_out])
//
// Replaces every nn.Linear of TemporalTransformer3DModel / TemporalTransformerBlock /
// VersatileAttention (reference animatediff/models/motion_module.py:147 proj_in, :155 proj_out,
// :215-219 to_q/to_k/to_v (fused [3C, C]) and to_out + residual, :221 GEGLU feed-forward), which the
// reference runs as separate cuBLAS GEMMs + elementwise bias/residual/GEGLU passes.
//
// Design (B200, tcgen05 / TMEM / TMA; persistent, warp-specialised, one CTA per SM):
//   warp 0   TMA producer : x tile [128 x 64] and w tile [BN x 64] (bf16, K-major, SWIZZLE_128B) into a
//                           multi-stage smem ring; mbarrier expect_tx / complete_tx
//   warp 1   MMA issuer   : one elected thread issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=BN<=256,
//                           K=16) x4 per stage; tcgen05.commit releases the smem stage / publishes the accumulator
//   warp 2   TMEM allocator (512 columns = two accumulator stages of up to 256 fp32 columns)
//   warps 4-19 epilogue   : tcgen05.ld -> +bias -> [GEGLU] -> +residual (register-prefetched) -> bf16 ->
//                           SWIZZLE_64B smem staging box -> TMA tensor store;
//                           overlaps the MMA of the next tile through the second TMEM stage.
// Both operands are K-major ("TN"): x rows and nn.Linear weight rows are contiguous along k, so the TMA
// box lands directly in the canonical UMMA SWIZZLE_128B layout (8-row x 128-byte atoms, SBO = 1024 B).
// GEGLU: the B tile is loaded as two half tiles, rows [n0, n0+BN/2) and [N/2+n0, N/2+n0+BN/2), so value and
// gate of the same output column sit in one accumulator and a*gelu(g) never round-trips through HBM.
#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int BM = 128;       // UMMA_M
constexpr int BK = 64;        // one 128-byte swizzle atom of 16-bit elements
constexpr int UMMA_K = 16;
constexpr int kAccCols = 256; // TMEM columns per accumulator stage
constexpr int kEpiParts = 4;            // epilogue warps per TMEM lane quarter
constexpr int kEpiWarps = 4 * kEpiParts;
constexpr int kEpiBytesPerWarp = 4096;  // 2 x 2 KB output staging boxes
constexpr int kThreads = 128 + 32 * kEpiWarps;  // 4 control warps (TMA, MMA, TMEM alloc, spare) + 16 epilogue warps

struct GemmParams {
  long long m;
  int n, k, bn, n_out;  // bn = accumulator columns per tile; n_out = output columns (n, or n/2 for GEGLU)
  int geglu;
  int num_m_blocks, num_n_blocks, num_k_blocks, stages;
  const float* bias;
  const void* residual;
  void* y;
  long long ldr, ldy;
  uint32_t idesc;
};

// ---- tcgen05 wrappers ----------------------------------------------------------------------
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
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

// GELU with the exact-erf definition (diffusers GEGLU uses F.gelu default); erf by Abramowitz-Stegun 7.1.26
// (|abs err| <= 1.5e-7, far below bf16 resolution): one MUFU.RCP + one MUFU.EX2 + 8 FMA instead of libdevice erff.
__device__ __forceinline__ float gelu_erf(float v) {
  const float x = fabsf(v) * 0.70710678118654752f;
  const float t = __frcp_rn(fmaf(0.3275911f, x, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erfc_x = poly * t * __expf(-x * x);          // 1 - erf(|v|/sqrt2)
  const float erf_v = copysignf(1.0f - erfc_x, v);
  return 0.5f * v * (1.0f + erf_v);
}

template <typename T>
__device__ __forceinline__ void store16(T* dst, const float (&v)[16]) {
  Vec16<T> a, b;
  float lo[8], hi[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    lo[j] = v[j];
    hi[j] = v[8 + j];
  }
  a.pack(lo);
  b.pack(hi);
  reinterpret_cast<uint4*>(dst)[0] = a.raw;
  reinterpret_cast<uint4*>(dst)[1] = b.raw;
}
template <typename T>
__device__ __forceinline__ void load16_add(const T* src, float (&v)[16]) {
  Vec16<T> a, b;
  a.raw = reinterpret_cast<const uint4*>(src)[0];
  b.raw = reinterpret_cast<const uint4*>(src)[1];
  float lo[8], hi[8];
  a.unpack(lo);
  b.unpack(hi);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    v[j] += lo[j];
    v[8 + j] += hi[j];
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 1)
    gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w,
                        const __grid_constant__ CUtensorMap map_y, const GemmParams p) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t full_bar[8], empty_bar[8], tmem_full[2], tmem_empty[2];
  __shared__ uint32_t tmem_base_slot;

  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t a_bytes = BM * BK * 2;
  const uint32_t b_bytes = (uint32_t)p.bn * BK * 2;
  const uint32_t stage_bytes = a_bytes + ((b_bytes + 1023) & ~1023u);
  const int stages = p.stages;
  const long long num_tiles = (long long)p.num_m_blocks * p.num_n_blocks;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(&tmem_base_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      prefetch_tensormap(&map_x);
      prefetch_tensormap(&map_w);
      int stage = 0;
      uint32_t phase = 0;
      const int half = p.bn / 2;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int nb = (int)(tile % p.num_n_blocks);
        const int mb = (int)(tile / p.num_n_blocks);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          unsigned char* sa = smem + (size_t)stage * stage_bytes;
          unsigned char* sb = sa + a_bytes;
          mbar_arrive_expect_tx(&full_bar[stage], a_bytes + b_bytes);
          tma_load_2d(sa, &map_x, &full_bar[stage], kb * BK, mb * BM);
          if (p.geglu) {
            tma_load_2d(sb, &map_w, &full_bar[stage], kb * BK, nb * half);
            tma_load_2d(sb + (size_t)half * BK * 2, &map_w, &full_bar[stage], kb * BK, p.n / 2 + nb * half);
          } else {
            tma_load_2d(sb, &map_w, &full_bar[stage], kb * BK, nb * p.bn);
          }
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (single thread) =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      long long it = 0;
      for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = (int)(it & 1);
        mbar_wait(&tmem_empty[acc], (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)acc * kAccCols;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = make_sw128_desc(sa);
          const uint64_t bdesc = make_sw128_desc(sa + a_bytes);
#pragma unroll
          for (int ks = 0; ks < BK / UMMA_K; ++ks) {
            // advance 16 elements (32 B) along K inside the swizzle atom: +2 in the (addr >> 4) field
            umma_f16(tmem_d, adesc + (uint64_t)(ks * 2), bdesc + (uint64_t)(ks * 2), p.idesc, (kb | ks) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);  // frees this smem stage once the MMAs above have read it
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
          if (++stage == stages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue (TMEM -> registers -> swizzled smem -> TMA store), 16 warps =====================
    // warp w: TMEM lane quarter q = w % 4 (hardware restriction), column part (w - 4) / 4 of the tile.  Work unit: a
    // box of 32 rows x 32 output columns.  The residual row segment (64 B = two full sectors per thread) is prefetched
    // into registers one box ahead (the first one before the accumulator is even ready); results are written as bf16
    // into a SWIZZLE_64B staging box (conflict-free 16-byte stores: chunk ^= (row / 2) & 3) and leave through a TMA
    // tensor store, so every global write of the epilogue is a coalesced bulk transfer clipped at the tensor edge.
    const int q = warp & 3, part = (warp - 4) >> 2, ew = warp - 4;
    const int out_cols = p.geglu ? p.bn / 2 : p.bn;
    const int boxes = out_cols / 32;
    const int b_begin = (boxes * part) / kEpiParts, b_end = (boxes * (part + 1)) / kEpiParts;
    unsigned char* stage_base = smem + (size_t)stages * stage_bytes + (size_t)ew * kEpiBytesPerWarp;
    unsigned char* obuf[2] = {stage_base, stage_base + 2048};
    const T* __restrict__ res = reinterpret_cast<const T*>(p.residual);
    const int sw = (lane >> 1) & 3;  // SWIZZLE_64B xor for this thread's row
    int ob = 0;
    if (lane == 0) prefetch_tensormap(&map_y);
    long long it = 0;
    for (long long tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int nb = (int)(tile % p.num_n_blocks);
      const int mb = (int)(tile / p.num_n_blocks);
      const int acc = (int)(it & 1);
      const int row0 = mb * BM + q * 32;   // first row of this warp's boxes
      const int n0 = nb * out_cols;        // first output column of this tile
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)acc * kAccCols;
      const bool live = res && (row0 + lane) < p.m;
      const T* res_row = live ? res + (long long)(row0 + lane) * p.ldr + n0 : nullptr;
      uint4 rres[4];
      if (live && b_begin < b_end) {
#pragma unroll
        for (int j = 0; j < 4; ++j) rres[j] = ldg_stream(res_row + b_begin * 32 + j * 8);
      }
      mbar_wait(&tmem_full[acc], (uint32_t)((it >> 1) & 1));
      tc_fence_after();
#pragma unroll 1
      for (int bx = b_begin; bx < b_end; ++bx) {
        const int col = bx * 32;  // column of this box inside the tile
        // staging buffer `ob` was handed to a TMA store two boxes ago: make sure that store has read it
        if (lane == 0) bulk_wait_read<1>();
        __syncwarp();
#pragma unroll
        for (int half = 0; half < 2; ++half) {
          uint32_t r[16];
          float v[16];
          tmem_ld16(taddr + col + 16 * half, r);
          const int gc = n0 + col + 16 * half;  // global output column
          if (p.geglu) {
            uint32_t g[16];
            tmem_ld16(taddr + out_cols + col + 16 * half, g);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f), bg = ba;
              if (p.bias) {
                ba = __ldg(reinterpret_cast<const float4*>(p.bias + gc + j));
                bg = __ldg(reinterpret_cast<const float4*>(p.bias + p.n / 2 + gc + j));
              }
              v[j + 0] = (__uint_as_float(r[j + 0]) + ba.x) * gelu_erf(__uint_as_float(g[j + 0]) + bg.x);
              v[j + 1] = (__uint_as_float(r[j + 1]) + ba.y) * gelu_erf(__uint_as_float(g[j + 1]) + bg.y);
              v[j + 2] = (__uint_as_float(r[j + 2]) + ba.z) * gelu_erf(__uint_as_float(g[j + 2]) + bg.z);
              v[j + 3] = (__uint_as_float(r[j + 3]) + ba.w) * gelu_erf(__uint_as_float(g[j + 3]) + bg.w);
            }
          } else {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; j += 4) {
              float4 ba = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) ba = __ldg(reinterpret_cast<const float4*>(p.bias + gc + j));
              v[j + 0] = __uint_as_float(r[j + 0]) + ba.x;
              v[j + 1] = __uint_as_float(r[j + 1]) + ba.y;
              v[j + 2] = __uint_as_float(r[j + 2]) + ba.z;
              v[j + 3] = __uint_as_float(r[j + 3]) + ba.w;
            }
          }
          if (live) {
            Vec16<T> a, b;
            a.raw = rres[2 * half];
            b.raw = rres[2 * half + 1];
            float lo[8], hi[8];
            a.unpack(lo);
            b.unpack(hi);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              v[j] += lo[j];
              v[8 + j] += hi[j];
            }
          }
          // this thread's row inside the box: 64 bytes = four 16-byte chunks, chunks 2*half and 2*half+1 here
          Vec16<T> a, b;
          float lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            lo[j] = v[j];
            hi[j] = v[8 + j];
          }
          a.pack(lo);
          b.pack(hi);
          *reinterpret_cast<uint4*>(obuf[ob] + lane * 64 + ((2 * half) ^ sw) * 16) = a.raw;
          *reinterpret_cast<uint4*>(obuf[ob] + lane * 64 + ((2 * half + 1) ^ sw) * 16) = b.raw;
        }
        if (live && bx + 1 < b_end) {  // residual of the next box
#pragma unroll
          for (int j = 0; j < 4; ++j) rres[j] = ldg_stream(res_row + col + 32 + j * 8);
        }
        fence_proxy_async();  // generic-proxy smem writes -> visible to the TMA (async proxy) store
        __syncwarp();
        if (lane == 0) {
          tma_store_2d(&map_y, obuf[ob], n0 + col, row0);
          bulk_commit();
        }
        ob ^= 1;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

int pick_bn(int n_cols, int cap) {  // largest multiple of 32 <= cap dividing n_cols (epilogue boxes are 32 columns wide)
  for (int bn = cap; bn >= 32; bn -= 32)
    if (n_cols % bn == 0) return bn;
  return 0;
}

}  // namespace
}  // namespace ca

namespace ca {
// Single-CTA (cta_group::1) implementation: kept as the A/B yardstick of gemm_pair_tcgen05.cu (CA_GEMM_IMPL=1cta).
int linear_1cta(const void* x, const void* w, const float* bias, const void* residual, void* y, long long m, int n, int k,
                long long ldx, long long ldr, long long ldy, int epilogue, int dtype, void* stream) {
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

  GemmParams p{};
  p.m = m; p.n = n; p.k = k; p.geglu = geglu ? 1 : 0; p.n_out = n_out;
  if (geglu) {
    const int half = pick_bn(n / 2, 128);
    CA_CHECK_ARG(half >= 32, "linear: cannot tile n=%d for GEGLU", n);
    p.bn = 2 * half;
    p.num_n_blocks = (n / 2) / half;
  } else {
    p.bn = pick_bn(n, 256);
    p.num_n_blocks = n / p.bn;
  }
  p.num_m_blocks = (int)((m + BM - 1) / BM);
  p.num_k_blocks = (k + BK - 1) / BK;
  p.bias = bias; p.residual = residual; p.y = y; p.ldr = ldr; p.ldy = ldy;
  // instruction descriptor (kind::f16): D=f32 [4,6)=1; A/B format [7,10)/[10,13): 1=bf16, 0=f16; A,B K-major (bits
  // 15,16 = 0); N>>3 at [17,23); M>>4 at [24,29)
  const uint32_t fmt = dtype == CA_BF16 ? 1u : 0u;
  p.idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(p.bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  const uint32_t stage_bytes = BM * BK * 2 + (((uint32_t)p.bn * BK * 2 + 1023) & ~1023u);
  const size_t epi_bytes = (size_t)kEpiWarps * kEpiBytesPerWarp;
  int stages = (int)((225 * 1024 - 1024 - epi_bytes) / stage_bytes);
  if (stages > 8) stages = 8;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem = (size_t)stages * stage_bytes + epi_bytes + 1024;

  CUtensorMap mx, mw;
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
    const uint32_t box[2] = {BK, (uint32_t)(geglu ? p.bn / 2 : p.bn)};
    if (!encode_tensor_map(&mw, dt, 2, w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B))
      return CA_ERR_CUDA;
  }
  CUtensorMap my;
  {
    const uint64_t dims[2] = {(uint64_t)n_out, (uint64_t)m};
    const uint32_t box[2] = {32, 32};
    const uint64_t sy[1] = {(uint64_t)ldy * 2};
    if (!encode_tensor_map(&my, dt, 2, y, dims, sy, box, CU_TENSOR_MAP_SWIZZLE_64B)) return CA_ERR_CUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const long long tiles = (long long)p.num_m_blocks * p.num_n_blocks;
  long long grid = sm_count();
  if (grid > tiles) grid = tiles;
  auto run = [&](auto kernel) -> int {
    CA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), 226 * 1024));
    kernel<<<(unsigned)grid, kThreads, smem, st>>>(mx, mw, my, p);
    CA_CUDA(cudaGetLastError());
    return CA_OK;
  };
  if (dtype == CA_BF16) return run(gemm_tcgen05_kernel<__nv_bfloat16>);
  return run(gemm_tcgen05_kernel<__half>);
}
}  // namespace ca
