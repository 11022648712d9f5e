// Kernel (1a): temporal self-attention core.  For every (batch, spatial site, head) the f frames
// (f <= 32) attend to each other: O = softmax(Q K^T * scale) V.
//
// Replaces the attention arithmetic the reference reaches through VersatileAttention.forward
// (animatediff/models/motion_module.py:285 rearrange, :321 processor call, :327 rearrange back) ->
// AttentionProcessor (modules/attention_processor.py:56-62 baddbmm/softmax/bmm, :247-256 SDPA;
// xformers memory_efficient_attention on its default GPU path).  The reference materialises two
// permuted copies ('(b f) d c -> (b d) f c' and back) plus head-split copies of Q, K, V; here the
// activations stay token-major [(b f), d, C] and the frame gather is done by the TMA engine.
//
// HBM-bound (AI = f/2 FLOP/B): algorithmic bytes = 4*T*C*s (read Q,K,V, write O).
//
// Design (B200):
//  * 5-D TMA tensor maps over the token-major buffers with dims (head_dim, frame, site, head, batch)
//    and box (hdp, fpad, S, 1, 1): one bulk-tensor copy gathers the f frames of S neighbouring sites of one
//    head into smem as S contiguous [f][hdp] matrices.  hdp = head_dim (+8) is chosen so the row
//    pitch is an odd multiple of 16 B -> ldmatrix is bank-conflict free without swizzle; the
//    columns beyond head_dim are out of bounds of dim 0 and are zero-filled by the TMA unit.
//  * mbarrier full/empty ring (3 stages), one producer warp, S consumer warps (one site each).
//  * per warp: S = Q K^T with mma.sync m16n8k16 (+ one m16n8k8 step when head_dim % 16 == 8),
//    fp32 scores, warp-shuffle row max / row sum (4 lanes per row), P split into bf16 hi + lo so
//    that P V carries ~16 mantissa bits, O scaled by 1/rowsum in fp32, staged in the consumed Q tile and
//    written with 16-byte streaming stores (whole head rows; no TMA store on the consumers' critical path).
#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kStages = 3;

struct AttnParams {
  int b, f, fpad, d, heads, hd, hdp;  // fpad = f rounded up to 8: rows per site in smem (128-B aligned sites)
  int S;             // sites per tile = consumer warps
  int site_tiles;    // ceil(d / S)
  long long units;   // b * site_tiles * heads
  float scale_log2;  // scale * log2(e)
  uint32_t tile_bytes;  // bytes of one operand tile in smem (S*f*hdp*2, padded to 128)
  // output addressing for the vectorised write-out: row(b, frame, site) = b*batch_rows + frame*frame_rows + site*site_rows
  void* o;
  long long ldo, batch_rows, frame_rows, site_rows;
};

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
  __device__ static uint32_t pack(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static float2 unpack(uint32_t u) { return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u)); }
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
  __device__ static uint32_t pack(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
  }
  __device__ static float2 unpack(uint32_t u) { return __half22float2(*reinterpret_cast<__half2*>(&u)); }
};

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
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

// MT = ceil(f / 16) query m-tiles (1 or 2); keys are handled as 2*MT n-tiles of 8.  HD > 0 fixes head_dim (and the smem row
// pitch) at compile time and F > 0 the frame count, so the loops unroll without per-tile bounds tests; r01d's
// capture showed the generic kernel issue-bound on index arithmetic (662 warp instructions per (site, head) unit, only 22
// of them HMMA, 64-bit div/mod per unit), not on HBM.
template <typename T, int MT, int HD, int F>
__global__ void __launch_bounds__(288, (MT == 1 || F > 0) ? 2 : 1)
    temporal_attn_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                         const __grid_constant__ CUtensorMap map_v,
                         const AttnParams p) {
  using M = MmaT<T>;
  constexpr int NT = 2 * MT;
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ uint64_t full_bar[kStages], empty_bar[kStages];

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const uint32_t tile_bytes = p.tile_bytes;
  const uint32_t stage_bytes = 3 * tile_bytes;
  const int hdp = HD > 0 ? (((HD / 8) & 1) ? HD : HD + 8) : p.hdp;
  const int pitch = hdp * 2;  // bytes per smem row

  // zero all of smem once: rows beyond f are read as (masked) padding and must hold finite data
  {
    uint4* z = reinterpret_cast<uint4*>(smem);
    const int n16 = (kStages * stage_bytes + 32 * pitch) / 16;
    for (int i = threadIdx.x; i < n16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
    fence_proxy_async();
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], S);
    }
    fence_mbar_init();
  }
  __syncthreads();

  // contiguous range of work units for this CTA
  const int per = (int)((p.units + gridDim.x - 1) / gridDim.x);
  const int u0 = (int)min(p.units, (long long)blockIdx.x * per);
  const int u1 = (int)min(p.units, (long long)u0 + per);
  // unit u = (bi * site_tiles + st) * heads + head, walked incrementally (no per-unit div/mod)
  int head = u0 % p.heads, st = (u0 / p.heads) % p.site_tiles, bi = (u0 / p.heads) / p.site_tiles;
  auto next_unit = [&]() {
    if (++head == p.heads) {
      head = 0;
      if (++st == p.site_tiles) {
        st = 0;
        ++bi;
      }
    }
  };

  if (warp == S) {
    // ===== TMA producer =====
    if (lane == 0) {
      prefetch_tensormap(&map_q);
      prefetch_tensormap(&map_k);
      prefetch_tensormap(&map_v);
      int stage = 0;
      uint32_t phase = 0;
      for (int u = u0; u < u1; ++u, next_unit()) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        unsigned char* dst = smem + stage * stage_bytes;
        mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
        tma_load_5d(dst, &map_q, &full_bar[stage], 0, 0, st * S, head, bi);
        tma_load_5d(dst + tile_bytes, &map_k, &full_bar[stage], 0, 0, st * S, head, bi);
        tma_load_5d(dst + 2 * tile_bytes, &map_v, &full_bar[stage], 0, 0, st * S, head, bi);
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }
  if (warp > S) return;

  // ===== consumers: warp w owns site (tile_site0 + w) =====
  const int f = F > 0 ? F : p.f, hd = HD > 0 ? HD : p.hd;
  const int site_bytes = (F > 0 ? (F + 7) / 8 * 8 : p.fpad) * pitch;
  const int r0 = lane >> 2, cq = (lane & 3) * 2;  // fragment row / column-pair
  int stage = 0;
  uint32_t phase = 0;
  for (int u = u0; u < u1; ++u, next_unit()) {
    mbar_wait(&full_bar[stage], phase);

    unsigned char* base = smem + stage * stage_bytes + warp * site_bytes;
    const uint32_t q_s = smem_u32(base), k_s = q_s + tile_bytes, v_s = q_s + 2 * tile_bytes;

    // ---- scores = Q K^T ----
    float sacc[MT][NT][4];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt)
#pragma unroll
      for (int nt = 0; nt < NT; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) sacc[mt][nt][j] = 0.f;

    const int k16 = hd >> 4;
#pragma unroll
    for (int ks = 0; ks < k16; ++ks) {
      uint32_t a[MT][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
        ldsm_x4(a[mt], q_s + (mt * 16 + (lane & 15)) * pitch + (ks * 16 + (lane >> 4) * 8) * 2);
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

    // ---- softmax over keys (fp32, exp2 with folded scale), P -> hi/lo 16-bit fragments ----
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
          const float s = key < f ? sacc[mt][nt][j] * p.scale_log2 : -INFINITY;
          sacc[mt][nt][j] = s;
          mx[j >> 1] = fmaxf(mx[j >> 1], s);
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
          e[j] = exp2f(sacc[mt][nt][j] - mx[j >> 1]);  // exp2f(-inf) = 0 for masked keys
          sum[j >> 1] += e[j];
        }
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint32_t hi = M::pack(e[2 * h], e[2 * h + 1]);
          const float2 hf = M::unpack(hi);
          p_hi[mt][nt][h] = hi;
          p_lo[mt][nt][h] = M::pack(e[2 * h] - hf.x, e[2 * h + 1] - hf.y);
        }
      }
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
        inv_sum[mt][h] = 1.0f / sum[h];
      }
    }

    // ---- O = P V, OCH output columns at a time (32 with two query m-tiles: keeps the accumulators at 32 registers so two
    // CTAs fit one SM); staged into this warp's (consumed) Q tile ----
    constexpr int OCH = MT == 2 ? 32 : 64;
#pragma unroll
    for (int c0 = 0; c0 < hd; c0 += OCH) {
      float oacc[MT][OCH / 8][4];
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < OCH / 8; ++nt)
#pragma unroll
          for (int j = 0; j < 4; ++j) oacc[mt][nt][j] = 0.f;
#pragma unroll
      for (int kt = 0; kt < MT; ++kt) {  // 16 keys per step
        if (kt * 16 < f) {
#pragma unroll
          for (int nt = 0; nt < OCH / 8; ++nt) {
            if (c0 + nt * 8 < hd) {
              uint32_t b0, b1;
              ldsm_x2_trans(b0, b1, v_s + (kt * 16 + (lane & 15)) * pitch + (c0 + nt * 8) * 2);
#pragma unroll
              for (int mt = 0; mt < MT; ++mt) {
                const uint32_t ah[4] = {p_hi[mt][2 * kt][0], p_hi[mt][2 * kt][1], p_hi[mt][2 * kt + 1][0],
                                        p_hi[mt][2 * kt + 1][1]};
                const uint32_t al[4] = {p_lo[mt][2 * kt][0], p_lo[mt][2 * kt][1], p_lo[mt][2 * kt + 1][0],
                                        p_lo[mt][2 * kt + 1][1]};
                M::k16(oacc[mt][nt], ah, b0, b1);
                M::k16(oacc[mt][nt], al, b0, b1);
              }
            }
          }
        }
      }
#pragma unroll
      for (int mt = 0; mt < MT; ++mt)
#pragma unroll
        for (int nt = 0; nt < OCH / 8; ++nt) {
          const int col = c0 + nt * 8 + cq;
          if (col < hd) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              const int row = mt * 16 + r0 + h * 8;
              if (row < f) {
                const uint32_t v = M::pack(oacc[mt][nt][2 * h] * inv_sum[mt][h], oacc[mt][nt][2 * h + 1] * inv_sum[mt][h]);
                *reinterpret_cast<uint32_t*>(base + row * pitch + col * 2) = v;
              }
            }
          }
        }
    }
    // ---- write this site's [f][hd] block with 16-byte stores straight from the staging tile, then release the stage.
    // (r01d: a TMA tensor store here cost 23 % of the consumers' time in cp.async.bulk.wait_group.read -- the store
    // queues behind the producer's loads in the TMA unit and the stage cannot be handed back before it has been read.)
    __syncwarp();
    {
      const int site = st * S + warp;
      if (site < p.d) {
        const int nvr = hd >> 3;  // 16-byte vectors per row
        T* og = reinterpret_cast<T*>(p.o) + ((long long)bi * p.batch_rows + (long long)site * p.site_rows) * p.ldo + head * hd;
        for (int v = lane; v < f * nvr; v += 32) {
          const int row = v / nvr, ch = v - row * nvr;
          const uint4 val = *reinterpret_cast<const uint4*>(base + row * pitch + ch * 16);
          stg_stream(og + (long long)row * p.frame_rows * p.ldo + ch * 8, val);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == kStages) {
      stage = 0;
      phase ^= 1;
    }
  }
}

// 5-D map (head_dim, frame, site, head, batch) over a token matrix with row stride `ld` elements where
// row(b, frame, site) = b*batch_rows + frame*frame_rows + site*site_rows.
bool make_attn_map(CUtensorMap* m, const void* base, int dtype, int b, int f, int d, int heads, int hd, long long ld,
                   long long batch_rows, long long frame_rows, long long site_rows, int box_hd, int box_f, int box_sites) {
  const uint64_t dims[5] = {(uint64_t)hd, (uint64_t)f, (uint64_t)d, (uint64_t)heads, (uint64_t)b};
  const uint64_t strides[4] = {(uint64_t)frame_rows * ld * 2, (uint64_t)site_rows * ld * 2, (uint64_t)hd * 2,
                               (uint64_t)batch_rows * ld * 2};
  const uint32_t box[5] = {(uint32_t)box_hd, (uint32_t)box_f, (uint32_t)box_sites, 1u, 1u};
  return encode_tensor_map(m, dtype == CA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5,
                           base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE);
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_temporal_attn_core(const void* q, const void* k, const void* v, void* o, int b, int f, int d,
                                     int heads, int head_dim, long long ldq, long long ldk, long long ldv,
                                     long long ldo, int seq_major, float scale, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(q && k && v && o, "temporal_attn_core: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "temporal_attn_core: dtype must be bf16 or f16");
  CA_CHECK_ARG(b > 0 && d > 0 && heads > 0, "temporal_attn_core: bad sizes");
  CA_CHECK_ARG(f >= 1 && f <= 32, "temporal_attn_core: f=%d outside [1, 32] (reference PE max_len, motion_module.py:236)", f);
  CA_CHECK_ARG(head_dim % 8 == 0 && head_dim >= 8 && head_dim <= 248, "temporal_attn_core: head_dim=%d must be a multiple of 8 in [8,248]", head_dim);
  const long long width = (long long)heads * head_dim;
  CA_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "temporal_attn_core: row stride < heads*head_dim");
  CA_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0, "temporal_attn_core: row strides must be multiples of 8 elements");
  CA_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "temporal_attn_core: pointers must be 16-byte aligned");
  CA_CHECK_ARG((long long)b * f < (1ll << 31) && d < (1 << 30), "temporal_attn_core: too many rows");
  CA_CHECK_ARG((long long)b * d * heads < (1ll << 31), "temporal_attn_core: too many (site, head) units");

  AttnParams p{};
  p.b = b; p.f = f; p.d = d; p.heads = heads; p.hd = head_dim;
  p.hdp = ((head_dim / 8) & 1) ? head_dim : head_dim + 8;  // odd number of 16-byte chunks per row
  p.fpad = (f + 7) / 8 * 8;
  const int site_bytes = p.fpad * p.hdp * 2;  // multiple of 128 B: every site tile is a legal TMA store source
  int S = 8;
  while (S > 1 && 3 * S * site_bytes > 34 * 1024) S >>= 1;
  if (S > d) {
    S = 1;
    while (S * 2 <= d && S < 8) S <<= 1;
  }
  p.S = S;
  p.site_tiles = (d + S - 1) / S;
  p.units = (long long)b * p.site_tiles * heads;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.tile_bytes = (uint32_t)(S * site_bytes);
  p.o = o; p.ldo = ldo;
  const size_t smem = (size_t)kStages * 3 * p.tile_bytes + 32 * p.hdp * 2 + 128;
  CA_CHECK_ARG(smem <= 220 * 1024, "temporal_attn_core: tile does not fit shared memory");

  // token-major rows t = (b*f + frame)*d + site, or the reference's "(b d) f c" rows t = (b*d + site)*f + frame
  const long long batch_rows = (long long)f * d;
  const long long frame_rows = seq_major ? 1 : d;
  const long long site_rows = seq_major ? f : 1;
  p.batch_rows = batch_rows; p.frame_rows = frame_rows; p.site_rows = site_rows;
  CUtensorMap mq, mk, mv;
  if (!make_attn_map(&mq, q, dtype, b, f, d, heads, head_dim, ldq, batch_rows, frame_rows, site_rows, p.hdp, p.fpad, S) ||
      !make_attn_map(&mk, k, dtype, b, f, d, heads, head_dim, ldk, batch_rows, frame_rows, site_rows, p.hdp, p.fpad, S) ||
      !make_attn_map(&mv, v, dtype, b, f, d, heads, head_dim, ldv, batch_rows, frame_rows, site_rows, p.hdp, p.fpad, S))
    return CA_ERR_CUDA;

  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int threads = (S + 1) * 32;
  auto run = [&](auto kernel) -> int {
    CA_CUDA(ensure_dynamic_smem(reinterpret_cast<const void*>(kernel), 220 * 1024));
    int per_sm = 1;
    CA_CUDA(cached_occupancy(&per_sm, reinterpret_cast<const void*>(kernel), threads, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > p.units) grid = p.units;
    kernel<<<(unsigned)grid, threads, smem, st>>>(mq, mk, mv, p);
    CA_CUDA(cudaGetLastError());
    return CA_OK;
  };
  const bool two = f > 16;
  if (dtype == CA_BF16) {
    // the AnimateDiff window lengths (8 / 16 / 24 / 32 frames) x the SD1.5 motion-module head widths are compile-time
    // specialisations; everything else takes the generic kernel
#define CA_TA_CASE(F_, HD_) if (f == F_ && head_dim == HD_) return run(temporal_attn_kernel<__nv_bfloat16, (F_ > 16 ? 2 : 1), HD_, F_>)
    CA_TA_CASE(16, 40); CA_TA_CASE(16, 80); CA_TA_CASE(16, 160);
    CA_TA_CASE(8, 40); CA_TA_CASE(8, 80); CA_TA_CASE(8, 160);
    CA_TA_CASE(24, 40); CA_TA_CASE(24, 80); CA_TA_CASE(24, 160);
    CA_TA_CASE(32, 40); CA_TA_CASE(32, 80); CA_TA_CASE(32, 160);
#undef CA_TA_CASE
    return two ? run(temporal_attn_kernel<__nv_bfloat16, 2, 0, 0>) : run(temporal_attn_kernel<__nv_bfloat16, 1, 0, 0>);
  }
  return two ? run(temporal_attn_kernel<__half, 2, 0, 0>) : run(temporal_attn_kernel<__half, 1, 0, 0>);
}
