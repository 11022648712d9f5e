// Row N2 (SURVEY.md §8 "next"): cross-attention core of the spatial transformer.  Every latent site of every frame
// attends to the L <= 96 prompt tokens (CLIP: 77; IP-Adapter adds a few image tokens):  O = softmax(Q K^T * scale) V.
//
// Replaces the attention arithmetic of attn2 in BasicTransformerBlock (reference animatediff/models/attention.py:
// 283-289 -> modules/attention_processor.py:56-62 baddbmm/softmax/bmm, :247-256 SDPA / xformers).  The library path
// (cuDNN SDPA picks an sm80 wmma flash kernel for this shape) took 259 us per call at the 64x64 level
// (profiles/r01c_torch_profile.txt) although the op only has to stream Q in and O out: K and V of one (prompt, head)
// are 2 x 6 KB and stay in shared memory for hundreds of query tiles.
//
// HBM-bound: algorithmic bytes = 2*T*C*s (read Q, write O); 4*T*L*C flop.
//
// Design (B200):
//  * persistent CTAs walk a contiguous range of units ordered (prompt, head, frame, query tile) so that K/V of a
//    (prompt, head) pair are loaded to smem once per ~hundreds of units;
//  * a 4-D TMA map (head_dim, site, frame, head) gathers the [128 sites][hdp] Q tile of one head (hdp = head_dim (+8):
//    odd number of 16-byte chunks per row -> conflict-free ldmatrix, the pad columns are zero-filled by the TMA unit),
//    3-stage mbarrier ring issued two units ahead by one thread;
//  * 8 warps x 16 query rows: S = Q K^T on mma.sync m16n8k16 (+ k8 tail for head_dim 40), fp32 softmax over the
//    <= 96 keys in registers (quad shuffles), P (bf16) V on mma.sync, O scaled by 1/rowsum, staged in the warp's own
//    (consumed) Q rows and written with 16-byte streaming stores.
#include "common.cuh"
#include "tma.cuh"

namespace ca {
namespace {

constexpr int kXStages = 3;
constexpr int kXWarps = 8;
constexpr int kXRows = 16 * kXWarps;  // query rows (sites) per tile

struct XParams {
  int n_frames, d, heads, hd, hdp, L;
  int frames_per_ctx;       // frame n uses prompt n / frames_per_ctx unless ctx_of_frame is given
  const int* ctx_of_frame;  // [n_frames] or null
  int n_ctx;
  int q_tiles;              // ceil(d / 128)
  long long units;          // n_frames * heads * q_tiles
  const void* k;
  const void* v;
  long long ldk, ldv, ctx_stride_k, ctx_stride_v;  // elements
  void* o;
  long long ldo;
  float scale_log2;
  uint32_t tile_bytes;      // one Q tile in smem
};

template <typename T>
struct XMma;
template <>
struct XMma<__nv_bfloat16> {
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
struct XMma<__half> {
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

__device__ __forceinline__ void x_ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void x_ldsm_x2(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}
__device__ __forceinline__ void x_ldsm_x1(uint32_t& r0, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x1.shared.b16 {%0}, [%1];" : "=r"(r0) : "r"(addr));
}
__device__ __forceinline__ void x_ldsm_x2_trans(uint32_t& r0, uint32_t& r1, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(addr));
}

// HD = head_dim (40 / 80 / 160), KT = key k16-steps (keys padded to 16*KT: 5 -> 80, 6 -> 96)
template <typename T, int HD, int KT>
__global__ void __launch_bounds__(kXWarps * 32) cross_attn_kernel(const __grid_constant__ CUtensorMap map_q, const XParams p) {
  using M = XMma<T>;
  constexpr int NT = 2 * KT;                             // key n-tiles of 8
  constexpr int HDP = ((HD / 8) & 1) ? HD : HD + 8;      // smem row pitch in elements: odd number of 16-byte chunks
  constexpr int PITCH = HDP * 2;
  constexpr int LP = 16 * KT;
  constexpr int OC = HD > 80 ? 80 : HD;                  // output columns per P V pass (bounds the accumulator registers)
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full_bar[kXStages];

  unsigned char* q_ring = smem;                                            // kXStages tiles [128][HDP]
  unsigned char* k_s = smem + (size_t)kXStages * p.tile_bytes;             // [LP][HDP]
  unsigned char* v_s = k_s + (size_t)LP * PITCH;                           // [LP][HDP]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kXStages; ++s) mbar_init(&full_bar[s], 1);
    fence_mbar_init();
    prefetch_tensormap(&map_q);
  }
  {  // K / V pad rows and pad columns must hold finite data (zeros); the Q ring is fully written by TMA (zero fill)
    uint4* z = reinterpret_cast<uint4*>(k_s);
    for (int i = threadIdx.x; i < 2 * LP * PITCH / 16; i += blockDim.x) z[i] = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();

  // contiguous unit range; unit u = ((ctx_slot * heads + head) * frames_in_slot + frame) * q_tiles + tile is walked as
  // (head-major inside a prompt) so K/V change rarely.  Frames are grouped by prompt through frames_per_ctx; with an
  // explicit ctx_of_frame map the frame order is kept and K/V are reloaded whenever the prompt index changes.
  const int per = (int)((p.units + gridDim.x - 1) / gridDim.x);
  const int u0 = (int)min(p.units, (long long)blockIdx.x * per);
  const int u1 = (int)min(p.units, (long long)u0 + per);
  // decomposition: u = ((blk * heads + head) * fpb + fr) * q_tiles + qt, frame = blk * fpb + fr, fpb = frames per block
  const int fpb = p.ctx_of_frame ? 1 : p.frames_per_ctx;
  auto decode = [&](int u, int& frame, int& head, int& qt) {
    qt = u % p.q_tiles;
    int r = u / p.q_tiles;
    const int fr = r % fpb;
    r /= fpb;
    head = r % p.heads;
    frame = (r / p.heads) * fpb + fr;
  };
  auto issue = [&](int u) {  // thread 0
    int frame, head, qt;
    decode(u, frame, head, qt);
    const int st = (u - u0) % kXStages;
    mbar_arrive_expect_tx(&full_bar[st], p.tile_bytes);
    tma_load_4d(q_ring + (size_t)st * p.tile_bytes, &map_q, &full_bar[st], 0, qt * kXRows, frame, head);
  };
  if (threadIdx.x == 0)
    for (int u = u0; u < min(u1, u0 + kXStages - 1); ++u) issue(u);

  const int r0 = lane >> 2, cq = (lane & 3) * 2;
  int cur_ctx = -1, cur_head = -1;
  // (qt, fr, head, blk) of the current unit are advanced incrementally; only thread 0 decodes (the look-ahead unit)
  int qt = u0 % p.q_tiles, fr = (u0 / p.q_tiles) % fpb, head = ((u0 / p.q_tiles) / fpb) % p.heads, blk = ((u0 / p.q_tiles) / fpb) / p.heads;
  int st = 0;
  uint32_t phase = 0;
  for (int u = u0; u < u1; ++u) {
    const int frame = blk * fpb + fr;
    const int ctx = p.ctx_of_frame ? __ldg(p.ctx_of_frame + frame) : frame / p.frames_per_ctx;

    if (ctx != cur_ctx || head != cur_head) {  // CTA-uniform
      __syncthreads();                          // everyone is done with the previous K/V
      const int nv = HD / 8;                    // 16-byte vectors per row
      const T* kg = reinterpret_cast<const T*>(p.k) + (long long)ctx * p.ctx_stride_k + head * HD;
      const T* vg = reinterpret_cast<const T*>(p.v) + (long long)ctx * p.ctx_stride_v + head * HD;
      for (int i = threadIdx.x; i < p.L * nv; i += blockDim.x) {
        const int row = i / nv, ch = i - row * nv;
        *reinterpret_cast<uint4*>(k_s + row * PITCH + ch * 16) = ldg_keep(kg + (long long)row * p.ldk + ch * 8);
        *reinterpret_cast<uint4*>(v_s + row * PITCH + ch * 16) = ldg_keep(vg + (long long)row * p.ldv + ch * 8);
      }
      cur_ctx = ctx;
      cur_head = head;
      __syncthreads();
    }

    mbar_wait(&full_bar[st], phase);
    unsigned char* base = q_ring + (size_t)st * p.tile_bytes + warp * 16 * PITCH;
    const uint32_t q_a = smem_u32(base), k_a = smem_u32(k_s), v_a = smem_u32(v_s);

    // ---- scores = Q K^T ----
    float sacc[NT][4];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt)
#pragma unroll
      for (int j = 0; j < 4; ++j) sacc[nt][j] = 0.f;
#pragma unroll
    for (int ks = 0; ks < HD / 16; ++ks) {
      uint32_t a[4];
      x_ldsm_x4(a, q_a + (lane & 15) * PITCH + (ks * 16 + (lane >> 4) * 8) * 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t b0, b1;
        x_ldsm_x2(b0, b1, k_a + (nt * 8 + (lane & 7)) * PITCH + (ks * 16 + ((lane >> 3) & 1) * 8) * 2);
        M::k16(sacc[nt], a, b0, b1);
      }
    }
    if constexpr (HD & 8) {
      constexpr int c0 = (HD / 16) * 16;
      uint32_t a0, a1;
      x_ldsm_x2(a0, a1, q_a + (lane & 15) * PITCH + c0 * 2);
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        uint32_t b0;
        x_ldsm_x1(b0, k_a + (nt * 8 + (lane & 7)) * PITCH + c0 * 2);
        M::k8(sacc[nt], a0, a1, b0);
      }
    }

    // ---- softmax over the keys (fp32) -> P as 16-bit A fragments.  The row maximum is taken on the raw scores (scale > 0),
    // scale*log2(e) and the maximum are folded into one packed FFMA2 per score pair, only the n-tiles that contain padded
    // keys are masked: r01d's capture showed this kernel issue-bound on the per-score softmax arithmetic. ----
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      if (nt * 8 + 8 > p.L) {  // warp-uniform: at most the last two tiles
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (nt * 8 + cq + (j & 1) >= p.L) sacc[nt][j] = -INFINITY;
      }
      mx[0] = fmaxf(mx[0], fmaxf(sacc[nt][0], sacc[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(sacc[nt][2], sacc[nt][3]));
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
      mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    const float2 sl2 = make_float2(p.scale_log2, p.scale_log2);
    const float2 nm0 = make_float2(-mx[0] * p.scale_log2, -mx[0] * p.scale_log2);
    const float2 nm1 = make_float2(-mx[1] * p.scale_log2, -mx[1] * p.scale_log2);
    float2 sum0 = make_float2(0.f, 0.f), sum1 = make_float2(0.f, 0.f);
    uint32_t pf[NT][2];
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const float2 t0 = __ffma2_rn(make_float2(sacc[nt][0], sacc[nt][1]), sl2, nm0);
      const float2 t1 = __ffma2_rn(make_float2(sacc[nt][2], sacc[nt][3]), sl2, nm1);
      const float2 e0 = make_float2(exp2f(t0.x), exp2f(t0.y));  // exp2f(-inf) = 0 for masked keys
      const float2 e1 = make_float2(exp2f(t1.x), exp2f(t1.y));
      sum0 = __fadd2_rn(sum0, e0);
      sum1 = __fadd2_rn(sum1, e1);
      pf[nt][0] = pack2(e0.x, e0.y, T());
      pf[nt][1] = pack2(e1.x, e1.y, T());
    }
    float sum[2] = {sum0.x + sum0.y, sum1.x + sum1.y};
    float inv_sum[2];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
      sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
      inv_sum[h] = 1.0f / sum[h];
    }

    // ---- O = P V, OC output columns per pass; staged into this warp's (consumed) Q rows ----
    __syncwarp();  // all lanes have read their Q fragments before the rows are overwritten
#pragma unroll
    for (int c0 = 0; c0 < HD; c0 += OC) {
      float oacc[OC / 8][4];
#pragma unroll
      for (int nt = 0; nt < OC / 8; ++nt)
#pragma unroll
        for (int j = 0; j < 4; ++j) oacc[nt][j] = 0.f;
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        const uint32_t a[4] = {pf[2 * kt][0], pf[2 * kt][1], pf[2 * kt + 1][0], pf[2 * kt + 1][1]};
#pragma unroll
        for (int nt = 0; nt < OC / 8; ++nt) {
          uint32_t b0, b1;
          x_ldsm_x2_trans(b0, b1, v_a + (kt * 16 + (lane & 15)) * PITCH + (c0 + nt * 8) * 2);
          M::k16(oacc[nt], a, b0, b1);
        }
      }
#pragma unroll
      for (int nt = 0; nt < OC / 8; ++nt) {
        const int col = c0 + nt * 8 + cq;
#pragma unroll
        for (int h = 0; h < 2; ++h)
          *reinterpret_cast<uint32_t*>(base + (r0 + h * 8) * PITCH + col * 2) =
              pack2(oacc[nt][2 * h] * inv_sum[h], oacc[nt][2 * h + 1] * inv_sum[h], T());
      }
    }
    __syncwarp();
    {  // 16-byte streaming stores of whole head rows
      constexpr int NVR = HD / 8;
      const int site0 = qt * kXRows + warp * 16;
      T* og = reinterpret_cast<T*>(p.o) + ((long long)frame * p.d + site0) * p.ldo + head * HD;
      for (int vI = lane; vI < 16 * NVR; vI += 32) {
        const int row = vI / NVR, ch = vI - row * NVR;
        if (site0 + row < p.d) {
          const uint4 val = *reinterpret_cast<const uint4*>(base + row * PITCH + ch * 16);
          stg_stream(og + (long long)row * p.ldo + ch * 8, val);
        }
      }
    }
    __syncthreads();  // the stage is free: every warp has read its rows back
    if (threadIdx.x == 0 && u + kXStages - 1 < u1) {
      fence_proxy_async();
      issue(u + kXStages - 1);
    }
    if (++qt == p.q_tiles) {
      qt = 0;
      if (++fr == fpb) {
        fr = 0;
        if (++head == p.heads) {
          head = 0;
          ++blk;
        }
      }
    }
    if (++st == kXStages) {
      st = 0;
      phase ^= 1u;
    }
  }
}

}  // namespace
}  // namespace ca

extern "C" __attribute__((visibility("default"))) int ca_cross_attn_core(const void* q, const void* k, const void* v, void* o,
                                                                         int n_frames, int d, int heads, int head_dim, int n_ctx,
                                                                         int kv_len, long long ldq, long long ldk, long long ldv,
                                                                         long long ldo, long long ctx_stride_k,
                                                                         long long ctx_stride_v, const int* ctx_of_frame,
                                                                         float scale, int dtype, void* stream) {
  using namespace ca;
  CA_CHECK_ARG(q && k && v && o, "cross_attn_core: null pointer");
  CA_CHECK_ARG(dtype == CA_BF16 || dtype == CA_F16, "cross_attn_core: dtype must be bf16 or f16");
  CA_CHECK_ARG(n_frames > 0 && d > 0 && heads > 0 && n_ctx > 0 && kv_len > 0, "cross_attn_core: bad sizes");
  CA_CHECK_ARG(ctx_of_frame || n_frames % n_ctx == 0, "cross_attn_core: n_frames=%d not a multiple of n_ctx=%d", n_frames, n_ctx);
  if (!(head_dim == 40 || head_dim == 80 || head_dim == 160) || kv_len > 96) {
    set_error("cross_attn_core: head_dim=%d / kv_len=%d outside the built variants (40/80/160, <= 96)", head_dim, kv_len);
    return CA_ERR_UNSUPPORTED;
  }
  const long long width = (long long)heads * head_dim;
  CA_CHECK_ARG(ldq >= width && ldk >= width && ldv >= width && ldo >= width, "cross_attn_core: row stride < heads*head_dim");
  CA_CHECK_ARG(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0 && ctx_stride_k % 8 == 0 && ctx_stride_v % 8 == 0,
               "cross_attn_core: strides must be multiples of 8 elements");
  CA_CHECK_ARG(aligned16(q) && aligned16(k) && aligned16(v) && aligned16(o), "cross_attn_core: pointers must be 16-byte aligned");

  XParams p{};
  p.n_frames = n_frames; p.d = d; p.heads = heads; p.hd = head_dim; p.L = kv_len;
  p.hdp = ((head_dim / 8) & 1) ? head_dim : head_dim + 8;
  p.frames_per_ctx = ctx_of_frame ? 1 : n_frames / n_ctx;
  p.ctx_of_frame = ctx_of_frame; p.n_ctx = n_ctx;
  p.q_tiles = (d + kXRows - 1) / kXRows;
  p.units = (long long)n_frames * heads * p.q_tiles;
  CA_CHECK_ARG(p.units < (1ll << 31), "cross_attn_core: too many units");
  p.k = k; p.v = v; p.ldk = ldk; p.ldv = ldv; p.ctx_stride_k = ctx_stride_k; p.ctx_stride_v = ctx_stride_v;
  p.o = o; p.ldo = ldo;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.tile_bytes = (uint32_t)(kXRows * p.hdp * 2);
  const int kt = kv_len <= 80 ? 5 : 6;
  const size_t smem = (size_t)kXStages * p.tile_bytes + 2 * (size_t)(16 * kt) * p.hdp * 2 + 1024;

  // 4-D map (head_dim, site, frame, head) over the token-major q [n_frames*d, ldq]
  CUtensorMap mq;
  {
    const uint64_t dims[4] = {(uint64_t)head_dim, (uint64_t)d, (uint64_t)n_frames, (uint64_t)heads};
    const uint64_t strides[3] = {(uint64_t)ldq * 2, (uint64_t)d * ldq * 2, (uint64_t)head_dim * 2};
    const uint32_t box[4] = {(uint32_t)p.hdp, (uint32_t)kXRows, 1u, 1u};
    if (!encode_tensor_map(&mq, dtype == CA_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, q, dims,
                           strides, box, CU_TENSOR_MAP_SWIZZLE_NONE))
      return CA_ERR_CUDA;
  }
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  auto run = [&](auto kernel) -> int {
    const void* fn = reinterpret_cast<const void*>(kernel);
    CA_CUDA(ensure_dynamic_smem(fn, smem));
    int per_sm = 1;
    CA_CUDA(cached_occupancy(&per_sm, fn, kXWarps * 32, smem));
    if (per_sm < 1) per_sm = 1;
    long long grid = (long long)sm_count() * per_sm;
    if (grid > p.units) grid = p.units;
    kernel<<<(unsigned)grid, kXWarps * 32, smem, st>>>(mq, p);
    CA_CUDA(cudaGetLastError());
    return CA_OK;
  };
#define CA_X_CASE(T_, HD_, KT_) if (head_dim == HD_ && kt == KT_) return run(cross_attn_kernel<T_, HD_, KT_>)
  if (dtype == CA_BF16) {
    CA_X_CASE(__nv_bfloat16, 40, 5); CA_X_CASE(__nv_bfloat16, 80, 5); CA_X_CASE(__nv_bfloat16, 160, 5);
    CA_X_CASE(__nv_bfloat16, 40, 6); CA_X_CASE(__nv_bfloat16, 80, 6); CA_X_CASE(__nv_bfloat16, 160, 6);
  } else {
    CA_X_CASE(__half, 40, 5); CA_X_CASE(__half, 80, 5); CA_X_CASE(__half, 160, 5);
    CA_X_CASE(__half, 40, 6); CA_X_CASE(__half, 80, 6); CA_X_CASE(__half, 160, 6);
  }
#undef CA_X_CASE
  set_error("cross_attn_core: no kernel variant");
  return CA_ERR_UNSUPPORTED;
}
