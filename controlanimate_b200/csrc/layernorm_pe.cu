// LayerNorm (+ temporal positional encoding) on token-major rows: one warp per token, the row
// lives in registers between the statistics and the normalise step -> one HBM read, one write.
//
// Replaces nn.LayerNorm at reference animatediff/models/motion_module.py:214,221 and the
// PositionalEncoding add of VersatileAttention.forward (:285-288).  The reference first permutes
// '(b f) d c -> (b d) f c' so that pe[:, :f] broadcasts along dim 1; here the frame index of token
// t = (b*f + frame)*d + site is recovered arithmetically and pe[frame] is added in place, so the
// permuted copy never exists.
//
// Design: persistent warps (grid = SMs x resident CTAs) walk the rows with a grid stride; the
// loads of the NEXT row are issued before the current row is reduced (register double buffer), so
// every warp keeps two rows of 16-byte loads in flight; gamma/beta sit in shared memory; NV (16-byte
// vectors per lane) is a template parameter so small channel counts keep the register file free for
// occupancy (c = 320 -> NV 2, 640 -> 3, 1280 -> 5).
#include <stdlib.h>

#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kWarpsPerCta = 8;

// LPR lanes cooperate on one row (32/LPR rows per warp per iteration); lane sl of a row owns 16-byte vectors
// sl, sl+LPR, ... (NV of them).  c = 320 -> LPR 8, NV 5; 640 -> 16, 5; 1280 -> 32, 5: every lane is busy.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
    layernorm_pe_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma,
                        const float* __restrict__ beta, const float* __restrict__ pe, long long rows, int c, int f,
                        int d, float eps) {
  constexpr int VEC = Traits<T>::kVec;
  constexpr int R = 32 / LPR;  // rows per warp per iteration
  extern __shared__ __align__(16) float s_gb[];  // gamma[c], beta[c]
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    s_gb[i] = gamma[i];
    s_gb[c + i] = beta[i];
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nvec = c / VEC;
  const long long row0 = ((long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * R + sub;
  const long long stride = (long long)gridDim.x * kWarpsPerCta * R;
  const float inv_c = 1.0f / (float)c;

  uint4 nxt[NV];
  auto load_row = [&](long long row) {
    const T* xr = x + row * c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = sl + i * LPR;
      if (vi < nvec) nxt[i] = ldg_stream(xr + vi * VEC);
    }
  };
  if (row0 < rows) load_row(row0);
  // all lanes of a warp run the same number of iterations (shuffles need the full warp)
  const long long warp_first = row0 - sub;
  for (long long base = warp_first; base < rows; base += stride) {
    const long long row = base + sub;
    const bool live = row < rows;
    float v[NV][VEC];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (live && sl + i * LPR < nvec) {
        Vec16<T> r;
        r.raw = nxt[i];
        r.unpack(v[i]);
#pragma unroll
        for (int j = 0; j < VEC; ++j) sum += v[i][j];
      }
    }
    if (row + stride < rows) load_row(row + stride);  // prefetch the next row while this one is reduced
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum * inv_c;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      if (live && sl + i * LPR < nvec) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
          const float dlt = v[i][j] - mean;
          sq += dlt * dlt;
        }
      }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq * inv_c + eps);
    const float nmr = -mean * rstd;
    if (!live) continue;
    const float* per = pe ? pe + (long long)((row / d) % f) * c : nullptr;
    T* yr = y + row * c;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int vi = sl + i * LPR;
      if (vi < nvec) {
        float o[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j += 4) {
          const float4 g4 = *reinterpret_cast<const float4*>(s_gb + vi * VEC + j);
          const float4 b4 = *reinterpret_cast<const float4*>(s_gb + c + vi * VEC + j);
          o[j + 0] = fmaf(fmaf(v[i][j + 0], rstd, nmr), g4.x, b4.x);
          o[j + 1] = fmaf(fmaf(v[i][j + 1], rstd, nmr), g4.y, b4.y);
          o[j + 2] = fmaf(fmaf(v[i][j + 2], rstd, nmr), g4.z, b4.z);
          o[j + 3] = fmaf(fmaf(v[i][j + 3], rstd, nmr), g4.w, b4.w);
          if (per) {
            const float4 p4 = __ldg(reinterpret_cast<const float4*>(per + vi * VEC + j));
            o[j + 0] += p4.x;
            o[j + 1] += p4.y;
            o[j + 2] += p4.z;
            o[j + 3] += p4.w;
          }
        }
        Vec16<T> r;
        r.pack(o);
        stg_stream(yr + vi * VEC, r.raw);
      }
    }
  }
}


// ---- flat variant (16-bit storage): no persistence, no ring.  One row group per sub-warp, one pass, exit: the block
// scheduler keeps many short-lived CTAs at different phases on every SM, which is what makes the plain elementwise
// epilogue kernel reach 78 % of the copy roofline in the same harness (profiles/r01d_notes.md §7).  The row is kept as raw
// 16-byte vectors (NV registers x 4) and unpacked per pass so that 4 CTAs x 8 warps fit the register file.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
    layernorm_flat_kernel(const T* __restrict__ x, T* __restrict__ y, const float* __restrict__ gamma, const float* __restrict__ beta,
                          const float* __restrict__ pe, long long rows, int c, int f, int d, float eps) {
  constexpr int VEC = 8;
  constexpr int R = 32 / LPR;
  // gamma | beta | pe[frame of the CTA's first row] | pe[next frame]: staged once per CTA.  r01e's capture showed the first
  // form of this kernel (parameters by __ldg per row) at 88 % L1TEX throughput: 96 B of fp32 parameters per 16 B of data.
  extern __shared__ __align__(16) float s_prm[];
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nvec = c / VEC;
  const long long cta_row0 = (long long)blockIdx.x * kWarpsPerCta * R;
  const long long row = cta_row0 + (long long)(threadIdx.x >> 5) * R + sub;
  const bool live = row < rows;
  const float inv_c = 1.0f / (float)c;
  uint4 raw[NV];
  const T* xr = x + row * c;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sl + i * LPR;
    raw[i] = (live && vi < nvec) ? ldg_stream(xr + vi * VEC) : make_uint4(0u, 0u, 0u, 0u);
  }
  // rows of one CTA are consecutive tokens: they touch at most two frames when rows-per-CTA <= d (checked on the host)
  const long long fr0 = cta_row0 / d;
  const int f0 = (int)(fr0 % f), f1 = f0 + 1 == f ? 0 : f0 + 1;
  {
    const int c4 = c / 4;
    const int parts = pe ? 4 : 2;
    for (int i = threadIdx.x; i < parts * c4; i += blockDim.x) {
      const int which = i / c4, j = i - which * c4;
      const float* src = which == 0 ? gamma : which == 1 ? beta : which == 2 ? pe + (long long)f0 * c : pe + (long long)f1 * c;
      reinterpret_cast<float4*>(s_prm)[i] = __ldg(reinterpret_cast<const float4*>(src) + j);
    }
  }
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float2 v[4];
    unpack2(raw[i].x, v[0].x, v[0].y, T());
    unpack2(raw[i].y, v[1].x, v[1].y, T());
    unpack2(raw[i].z, v[2].x, v[2].y, T());
    unpack2(raw[i].w, v[3].x, v[3].y, T());
    sum2 = __fadd2_rn(sum2, __fadd2_rn(__fadd2_rn(v[0], v[1]), __fadd2_rn(v[2], v[3])));
  }
  float sum = sum2.x + sum2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * inv_c;
  const float2 nmean = make_float2(-mean, -mean);
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (sl + i * LPR < nvec) {
      float2 v[4];
      unpack2(raw[i].x, v[0].x, v[0].y, T());
      unpack2(raw[i].y, v[1].x, v[1].y, T());
      unpack2(raw[i].z, v[2].x, v[2].y, T());
      unpack2(raw[i].w, v[3].x, v[3].y, T());
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 dl = __fadd2_rn(v[j], nmean);
        sq2 = __ffma2_rn(dl, dl, sq2);
      }
    }
  }
  float sq = sq2.x + sq2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  const float rstd = rsqrtf(sq * inv_c + eps);
  __syncthreads();  // parameters staged (every thread reaches this: no early exit above)
  if (!live) return;
  const float2 rstd2 = make_float2(rstd, rstd), nmr2 = make_float2(-mean * rstd, -mean * rstd);
  const float* s_g = s_prm;
  const float* s_b = s_prm + c;
  const float* per = pe ? s_prm + 2 * c + ((row / d) != fr0 ? c : 0) : nullptr;
  T* yr = y + row * c;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sl + i * LPR;
    if (vi < nvec) {
      float2 v[4], o[4];
      unpack2(raw[i].x, v[0].x, v[0].y, T());
      unpack2(raw[i].y, v[1].x, v[1].y, T());
      unpack2(raw[i].z, v[2].x, v[2].y, T());
      unpack2(raw[i].w, v[3].x, v[3].y, T());
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float4 g4 = *reinterpret_cast<const float4*>(s_g + vi * VEC + 4 * j);
        const float4 b4 = *reinterpret_cast<const float4*>(s_b + vi * VEC + 4 * j);
        float2 b01 = make_float2(b4.x, b4.y), b23 = make_float2(b4.z, b4.w);
        if (per) {
          const float4 p4 = *reinterpret_cast<const float4*>(per + vi * VEC + 4 * j);
          b01 = __fadd2_rn(b01, make_float2(p4.x, p4.y));
          b23 = __fadd2_rn(b23, make_float2(p4.z, p4.w));
        }
        o[2 * j] = __ffma2_rn(__ffma2_rn(v[2 * j], rstd2, nmr2), make_float2(g4.x, g4.y), b01);
        o[2 * j + 1] = __ffma2_rn(__ffma2_rn(v[2 * j + 1], rstd2, nmr2), make_float2(g4.z, g4.w), b23);
      }
      uint4 out;
      out.x = pack2(o[0].x, o[0].y, T());
      out.y = pack2(o[1].x, o[1].y, T());
      out.z = pack2(o[2].x, o[2].y, T());
      out.w = pack2(o[3].x, o[3].y, T());
      stg_stream(yr + vi * VEC, out);
    }
  }
}

// ---- pipelined variant (16-bit storage): rows travel HBM -> smem by bulk async copies, registers hold one row only ----
// The register-resident kernel above keeps at most two rows of loads in flight per warp and, at ~100 registers per
// thread, 16 warps per SM: ~40 KB outstanding per SM, measured 46 % of the copy roofline at c = 320
// (profiles/r01c_microbench_quick.json).  Here a producer warp keeps a ring of `stages` tiles (rows_per_tile consecutive
// rows = one contiguous span, ~20 KB) in flight per CTA with cp.async.bulk + mbarriers; consumer warps read a row from
// smem as raw 16-byte vectors, do the exact two-pass statistics in registers and stream the result out.  Bytes in flight
// no longer depend on registers or occupancy, and there is no CTA-wide barrier in the loop.
constexpr int kLnRingWarps = 8;    // consumer warps
constexpr int kLnRingStagesMax = 8;

struct LnRingParams {
  const void* x;
  void* y;
  const float* gamma;
  const float* beta;
  const float* pe;
  long long rows;
  int c, f, d;
  float eps;
  int rows_per_tile, stages;
  long long n_tiles;
  unsigned int tile_bytes;  // rows_per_tile rows (+ two fp32 PE rows when pe_staged)
  int pe_staged;            // the producer copies pe[frame(r0)] and pe[frame(r0) + 1] behind the rows of every tile
};

__device__ __forceinline__ void ln_bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

template <typename T, int LPR, int NV>
__global__ void __launch_bounds__((kLnRingWarps + 1) * 32, 2) layernorm_ring_kernel(const LnRingParams p) {
  constexpr int VEC = 8;
  constexpr int R = 32 / LPR;
  extern __shared__ __align__(128) unsigned char ln_smem[];
  __shared__ __align__(8) uint64_t full_bar[kLnRingStagesMax], empty_bar[kLnRingStagesMax];
  const int c = p.c, nvec = c / VEC;
  unsigned char* ring = ln_smem;
  float* s_g = reinterpret_cast<float*>(ln_smem + (size_t)p.stages * p.tile_bytes);
  float* s_b = s_g + c;
  for (int i = threadIdx.x; i < c; i += blockDim.x) {
    s_g[i] = p.gamma[i];
    s_b[i] = p.beta[i];
  }
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kLnRingWarps);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long row_bytes = (long long)c * (long long)sizeof(T);

  if (warp == kLnRingWarps) {
    // ===== producer =====
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
        const long long r0 = t * p.rows_per_tile;
        const long long nr = min((long long)p.rows_per_tile, p.rows - r0);
        const uint32_t total = (uint32_t)(nr * row_bytes);
        const uint32_t pe_b = p.pe_staged ? (uint32_t)c * 4u : 0u;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], total + 2 * pe_b);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(p.x) + r0 * row_bytes;
        unsigned char* dst = ring + (size_t)stage * p.tile_bytes;
        constexpr uint32_t kPiece = 8 * 1024;
        for (uint32_t off = 0; off < total; off += kPiece) ln_bulk_load(dst + off, src + off, min(kPiece, total - off), &full_bar[stage]);
        if (pe_b) {  // a tile spans at most two frames (rows_per_tile <= d): their PE rows ride along (L2-resident table)
          const int f0 = (int)((r0 / p.d) % p.f), f1 = f0 + 1 == p.f ? 0 : f0 + 1;
          unsigned char* pd = dst + (size_t)p.rows_per_tile * row_bytes;
          ln_bulk_load(pd, p.pe + (long long)f0 * c, pe_b, &full_bar[stage]);
          ln_bulk_load(pd + pe_b, p.pe + (long long)f1 * c, pe_b, &full_bar[stage]);
        }
        if (++stage == p.stages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
    return;
  }

  // ===== consumers: sub-warp `sub` of warp `warp` owns rows warp*R + sub (+ kLnRingWarps*R ...) of every tile =====
  const int sub = lane / LPR, sl = lane % LPR;
  const float inv_c = 1.0f / (float)c;
  int stage = 0;
  uint32_t phase = 0;
  for (long long t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
    const long long r0 = t * p.rows_per_tile;
    const int nr = (int)min((long long)p.rows_per_tile, p.rows - r0);
    mbar_wait(&full_bar[stage], phase);
    const unsigned char* tile = ring + (size_t)stage * p.tile_bytes;
    const float* pe_s = reinterpret_cast<const float*>(tile + (size_t)p.rows_per_tile * row_bytes);
    const int to_next_frame = p.pe_staged ? (int)min((long long)p.rows_per_tile, (r0 / p.d + 1) * p.d - r0) : 0;
    for (int rb = warp * R; rb < nr; rb += kLnRingWarps * R) {  // warp-uniform trip count (shuffles need the full warp)
      const int r = rb + sub;
      const bool live = r < nr;
      const uint4* rowv = reinterpret_cast<const uint4*>(tile + (size_t)r * row_bytes);
      // the row lives in registers as fp32 pairs; all arithmetic is packed f32x2 (FADD2/FFMA2: half the issue slots)
      float2 v[NV][VEC / 2];
      float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int vi = sl + i * LPR;
        uint4 raw = make_uint4(0u, 0u, 0u, 0u);
        if (live && vi < nvec) raw = rowv[vi];
        unpack2(raw.x, v[i][0].x, v[i][0].y, T());
        unpack2(raw.y, v[i][1].x, v[i][1].y, T());
        unpack2(raw.z, v[i][2].x, v[i][2].y, T());
        unpack2(raw.w, v[i][3].x, v[i][3].y, T());
#pragma unroll
        for (int j = 0; j < VEC / 2; ++j) sum2 = __fadd2_rn(sum2, v[i][j]);  // padding vectors are zero
      }
      float sum = sum2.x + sum2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      const float mean = sum * inv_c;
      const float2 nmean = make_float2(-mean, -mean);
      float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        if (sl + i * LPR < nvec) {  // padding vectors must not contribute (0 - mean)^2
#pragma unroll
          for (int j = 0; j < VEC / 2; ++j) {
            v[i][j] = __fadd2_rn(v[i][j], nmean);  // keep the centred value: the normalise step reuses it
            sq2 = __ffma2_rn(v[i][j], v[i][j], sq2);
          }
        }
      }
      float sq = sq2.x + sq2.y;
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
      const float rstd = rsqrtf(sq * inv_c + p.eps);
      const float2 rstd2 = make_float2(rstd, rstd);
      if (live) {
        const long long row = r0 + r;
        const float* per = !p.pe ? nullptr
                           : p.pe_staged ? pe_s + (r >= to_next_frame ? c : 0)
                                         : p.pe + (long long)((row / p.d) % p.f) * c;
        T* yr = reinterpret_cast<T*>(p.y) + row * c;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          const int vi = sl + i * LPR;
          if (vi < nvec) {
            float2 o[VEC / 2];
#pragma unroll
            for (int j = 0; j < VEC / 4; ++j) {
              const float4 g4 = *reinterpret_cast<const float4*>(s_g + vi * VEC + 4 * j);
              const float4 b4 = *reinterpret_cast<const float4*>(s_b + vi * VEC + 4 * j);
              float2 b01 = make_float2(b4.x, b4.y), b23 = make_float2(b4.z, b4.w);
              if (per) {
                const float4 p4 = *reinterpret_cast<const float4*>(per + vi * VEC + 4 * j);  // smem (staged) or global
                b01 = __fadd2_rn(b01, make_float2(p4.x, p4.y));
                b23 = __fadd2_rn(b23, make_float2(p4.z, p4.w));
              }
              o[2 * j] = __ffma2_rn(__fmul2_rn(v[i][2 * j], rstd2), make_float2(g4.x, g4.y), b01);
              o[2 * j + 1] = __ffma2_rn(__fmul2_rn(v[i][2 * j + 1], rstd2), make_float2(g4.z, g4.w), b23);
            }
            uint4 out;
            out.x = pack2(o[0].x, o[0].y, T());
            out.y = pack2(o[1].x, o[1].y, T());
            out.z = pack2(o[2].x, o[2].y, T());
            out.w = pack2(o[3].x, o[3].y, T());
            stg_stream(yr + vi * VEC, out);
          }
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(&empty_bar[stage]);
    if (++stage == p.stages) {
      stage = 0;
      phase ^= 1;
    }
  }
}


// Row statistics only: (mean, rstd) of every row, the arithmetic of layernorm_flat_kernel's first two passes (exact two-pass
// variance from the register-resident row).  Feeds ca_linear_ln, which applies the normalisation in the epilogue of the
// projection that consumes the LayerNorm — the normalised tensor is never written.
template <typename T, int LPR, int NV>
__global__ void __launch_bounds__(kWarpsPerCta * 32, 4)
    row_stats_kernel(const T* __restrict__ x, float2* __restrict__ stats, long long rows, int c, long long ldx, float eps) {
  constexpr int VEC = 8;
  constexpr int R = 32 / LPR;
  const int lane = threadIdx.x & 31;
  const int sub = lane / LPR, sl = lane % LPR;
  const int nvec = c / VEC;
  const long long row = ((long long)blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * R + sub;
  const bool live = row < rows;
  const float inv_c = 1.0f / (float)c;
  uint4 raw[NV];
  const T* xr = x + row * ldx;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int vi = sl + i * LPR;
    raw[i] = (live && vi < nvec) ? ldg_keep(xr + vi * VEC) : make_uint4(0u, 0u, 0u, 0u);  // the projection re-reads x: keep it in L2
  }
  float2 sum2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float2 v[4];
    unpack2(raw[i].x, v[0].x, v[0].y, T());
    unpack2(raw[i].y, v[1].x, v[1].y, T());
    unpack2(raw[i].z, v[2].x, v[2].y, T());
    unpack2(raw[i].w, v[3].x, v[3].y, T());
    sum2 = __fadd2_rn(sum2, __fadd2_rn(__fadd2_rn(v[0], v[1]), __fadd2_rn(v[2], v[3])));
  }
  float sum = sum2.x + sum2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float mean = sum * inv_c;
  const float2 nmean = make_float2(-mean, -mean);
  float2 sq2 = make_float2(0.f, 0.f);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (sl + i * LPR < nvec) {
      float2 v[4];
      unpack2(raw[i].x, v[0].x, v[0].y, T());
      unpack2(raw[i].y, v[1].x, v[1].y, T());
      unpack2(raw[i].z, v[2].x, v[2].y, T());
      unpack2(raw[i].w, v[3].x, v[3].y, T());
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 dl = __fadd2_rn(v[j], nmean);
        sq2 = __ffma2_rn(dl, dl, sq2);
      }
    }
  }
  float sq = sq2.x + sq2.y;
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (live && sl == 0) stats[row] = make_float2(mean, rsqrtf(sq * inv_c + eps));
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_row_stats(const void* x, float* stats, long long rows, int c,
                                                                   long long ldx, float eps, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && stats, "row_stats: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "row_stats: dtype must be bf16 or f16");
  CA_CHECK_ARG(rows >= 0 && c > 0 && c % 8 == 0 && ldx >= c && ldx % 8 == 0, "row_stats: bad sizes rows=%lld c=%d ldx=%lld", rows, c, ldx);
  CA_CHECK_ARG(aligned16(x) && (reinterpret_cast<uintptr_t>(stats) & 7) == 0, "row_stats: misaligned pointer");
  if (rows == 0) return CA_OK;
  const int nvec = c / 8;
  int lpr = 8;
  while (lpr < 32 && (nvec + lpr - 1) / lpr > 5) lpr <<= 1;
  const int nv = (nvec + lpr - 1) / lpr;
  CA_CHECK_ARG(nv <= 5, "row_stats: c=%d too wide (<= 1280)", c);
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int rows_per_cta = kWarpsPerCta * (32 / lpr);
  const long long grid = (rows + rows_per_cta - 1) / rows_per_cta;
  CA_CHECK_ARG(grid < (1ll << 31), "row_stats: too many rows");
  const int rc = dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    if constexpr (sizeof(T) == 2) {
      auto run = [&](auto kernel) -> int {
        kernel<<<(unsigned)grid, kWarpsPerCta * 32, 0, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<float2*>(stats), rows, c, ldx, eps);
        return CA_OK;
      };
#define CA_RS_CASE(L_, N_) if (lpr == L_ && nv <= N_) return run(row_stats_kernel<T, L_, N_>)
      CA_RS_CASE(8, 1); CA_RS_CASE(8, 2); CA_RS_CASE(8, 3); CA_RS_CASE(8, 5);
      CA_RS_CASE(16, 3); CA_RS_CASE(16, 5);
      CA_RS_CASE(32, 3); CA_RS_CASE(32, 5);
#undef CA_RS_CASE
    }
    return CA_ERR_UNSUPPORTED;
  });
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}

namespace ca {
namespace {
}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_layernorm_pe(const void* x, void* y, const float* gamma,
                                                                      const float* beta, const float* pe, long long rows,
                                                                      int c, int f, int d, float eps, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(x && y && gamma && beta, "layernorm_pe: null pointer");
  CA_CHECK_ARG(rows >= 0 && c > 0 && f > 0 && d > 0, "layernorm_pe: bad sizes");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16 || dtype == CA_F32, "layernorm_pe: bad dtype");
  const int vec = dtype == CA_F32 ? 4 : 8;
  const int nvec = c / vec;
  CA_CHECK_ARG(c % vec == 0 && nvec <= 32 * 16, "layernorm_pe: c=%d unsupported (multiple of %d, <= %d)", c, vec, 512 * vec);
  CA_CHECK_ARG(aligned16(x) && aligned16(y) && (!pe || aligned16(pe)), "layernorm_pe: x/y/pe must be 16-byte aligned");
  if (rows == 0) return CA_OK;
  // lanes per row: smallest of 8/16/32 that needs at most 5 vectors per lane (8 for very wide rows)
  int lpr = 8;
  while (lpr < 32 && (nvec + lpr - 1) / lpr > 5) lpr <<= 1;
  const int nv = (nvec + lpr - 1) / lpr;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  static const int ln_mode = [] { const char* e = getenv("CA_LN_MODE"); return !e ? 0 : (e[0] == 'f' ? 1 : (e[0] == 'r' ? 2 : (e[0] == 'p' ? 3 : 0))); }();
  // 0 auto, 1 flat (non-persistent), 2 ring, 3 persistent register-resident
  if ((ln_mode == 1 || ln_mode == 0) && dtype != CA_F32 && nv <= 5 && aligned16(gamma) && aligned16(beta) &&
      (!pe || kWarpsPerCta * (32 / lpr) <= d)) {  // a CTA's rows touch at most two frames
    const int rc2 = dispatch_dtype(dtype, [&](auto tag) -> int {
      using T = decltype(tag);
      if constexpr (sizeof(T) == 2) {
        auto run = [&](auto kernel) -> int {
          const int rows_per_cta = kWarpsPerCta * (32 / lpr);
          const long long grid = (rows + rows_per_cta - 1) / rows_per_cta;
          if (grid >= (1ll << 31)) return CA_ERR_UNSUPPORTED;
          const size_t prm = (size_t)(pe ? 4 : 2) * c * sizeof(float);
          if (prm > 48 * 1024) return CA_ERR_UNSUPPORTED;
          kernel<<<(unsigned)grid, kWarpsPerCta * 32, prm, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), gamma, beta, pe,
                                                                 rows, c, f, d, eps);
          return CA_OK;
        };
#define CA_LNF_CASE(L_, N_) if (lpr == L_ && nv <= N_) return run(layernorm_flat_kernel<T, L_, N_>)
        CA_LNF_CASE(8, 1); CA_LNF_CASE(8, 2); CA_LNF_CASE(8, 3); CA_LNF_CASE(8, 5);
        CA_LNF_CASE(16, 3); CA_LNF_CASE(16, 5);
        CA_LNF_CASE(32, 3); CA_LNF_CASE(32, 5);
#undef CA_LNF_CASE
      }
      return CA_ERR_UNSUPPORTED;
    });
    if (rc2 == CA_OK) {
      CA_CUDA(cudaGetLastError());
      return CA_OK;
    }
    if (rc2 != CA_ERR_UNSUPPORTED) return rc2;
  }
  static const int ring_on = [] { const char* e = getenv("CA_LN_RING"); return (e && e[0] == '0') ? 0 : 1; }();
  static const int ring_kb = [] { const char* e = getenv("CA_LN_RING_KB"); return (e && e[0]) ? atoi(e) : 20; }();
  static const int ring_stages = [] { const char* e = getenv("CA_LN_RING_STAGES"); return (e && e[0]) ? atoi(e) : 4; }();
  if (ring_on && ln_mode != 3 && dtype != CA_F32 && nv <= 5) {
    // pipelined path: tiles of whole row groups (kLnRingWarps * 32/lpr rows), about ring_kb KB each
    const int group = kLnRingWarps * (32 / lpr);
    const long long row_bytes = (long long)c * 2;
    long long rpt = ((long long)ring_kb * 1024 / row_bytes) / group * group;
    if (rpt < group) rpt = group;
    LnRingParams p{};
    p.x = x; p.y = y; p.gamma = gamma; p.beta = beta; p.pe = pe; p.rows = rows; p.c = c; p.f = f; p.d = d; p.eps = eps;
    p.rows_per_tile = (int)rpt;
    p.stages = ring_stages < 2 ? 2 : (ring_stages > kLnRingStagesMax ? kLnRingStagesMax : ring_stages);
    p.n_tiles = (rows + rpt - 1) / rpt;
    p.pe_staged = (pe && rpt <= d) ? 1 : 0;
    p.tile_bytes = (unsigned int)(rpt * row_bytes + (p.pe_staged ? 2 * (size_t)c * sizeof(float) : 0));
    const size_t smem_ring = (size_t)p.stages * p.tile_bytes + 2 * (size_t)c * sizeof(float);
    if (smem_ring <= 100 * 1024) {
      const int threads = (kLnRingWarps + 1) * 32;
      const int rc2 = dispatch_dtype(dtype, [&](auto tag) -> int {
        using T = decltype(tag);
        if constexpr (sizeof(T) == 2) {
          auto run = [&](auto kernel) -> int {
            const void* fn = reinterpret_cast<const void*>(kernel);
            CA_CUDA(ensure_dynamic_smem(fn, smem_ring));
            int per_sm = 1;
            CA_CUDA(cached_occupancy(&per_sm, fn, threads, smem_ring));
            long long grid = (long long)sm_count() * (per_sm < 1 ? 1 : per_sm);
            if (grid > p.n_tiles) grid = p.n_tiles;
            kernel<<<(unsigned)grid, threads, smem_ring, st>>>(p);
            return CA_OK;
          };
#define CA_LNR_CASE(L_, N_) if (lpr == L_ && nv <= N_) return run(layernorm_ring_kernel<T, L_, N_>)
          CA_LNR_CASE(8, 1); CA_LNR_CASE(8, 2); CA_LNR_CASE(8, 3); CA_LNR_CASE(8, 5);
          CA_LNR_CASE(16, 3); CA_LNR_CASE(16, 5);
          CA_LNR_CASE(32, 3); CA_LNR_CASE(32, 5);
#undef CA_LNR_CASE
        }
        return CA_ERR_UNSUPPORTED;
      });
      if (rc2 == CA_OK) {
        CA_CUDA(cudaGetLastError());
        return CA_OK;
      }
      if (rc2 != CA_ERR_UNSUPPORTED) return rc2;
    }
  }
  const size_t smem = 2 * (size_t)c * sizeof(float);
  const int rc = dispatch_dtype(dtype, [&](auto tag) -> int {
    using T = decltype(tag);
    auto run = [&](auto kernel) -> int {
      const void* fn = reinterpret_cast<const void*>(kernel);
      int per_sm = 1;
      CA_CUDA(cached_occupancy(&per_sm, fn, kWarpsPerCta * 32, smem));
      const int rows_per_cta = kWarpsPerCta * (32 / lpr);
      long long grid = (long long)sm_count() * (per_sm < 1 ? 1 : per_sm);
      const long long need = (rows + rows_per_cta - 1) / rows_per_cta;
      if (grid > need) grid = need;
      kernel<<<(unsigned)grid, kWarpsPerCta * 32, smem, st>>>(reinterpret_cast<const T*>(x), reinterpret_cast<T*>(y), gamma,
                                                              beta, pe, rows, c, f, d, eps);
      return CA_OK;
    };
#define CA_LN_CASE(L_, N_) if (lpr == L_ && nv <= N_) return run(layernorm_pe_kernel<T, L_, N_>)
    CA_LN_CASE(8, 1); CA_LN_CASE(8, 2); CA_LN_CASE(8, 3); CA_LN_CASE(8, 5);
    CA_LN_CASE(16, 3); CA_LN_CASE(16, 5);
    CA_LN_CASE(32, 3); CA_LN_CASE(32, 5); CA_LN_CASE(32, 8); CA_LN_CASE(32, 16);
#undef CA_LN_CASE
    set_error("layernorm_pe: no kernel for c=%d", c);
    return CA_ERR_UNSUPPORTED;
  });
  if (rc != CA_OK) return rc;
  CA_CUDA(cudaGetLastError());
  return CA_OK;
}
