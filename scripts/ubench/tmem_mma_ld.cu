// Does tcgen05.ld (TMEM -> registers) interfere with tcgen05.mma running on the same SM?
// One CTA per SM: warp 0 issues back-to-back UMMAs (M=128, N=256, K=16, cta_group::1) into TMEM columns [0,256);
// warps 4.. loop on tcgen05.ld of columns [256,512).  Prints cycles for MMA alone, LD alone and both together.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t make_sw128_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
template <int COLS> __device__ __forceinline__ uint32_t tld(uint32_t taddr);
template <> __device__ __forceinline__ uint32_t tld<16>(uint32_t taddr) {
  uint32_t r[16];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return r[0] ^ r[sizeof(r) / 4 - 1];
}
template <> __device__ __forceinline__ uint32_t tld<32>(uint32_t taddr) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return r[0] ^ r[sizeof(r) / 4 - 1];
}
template <> __device__ __forceinline__ uint32_t tld<64>(uint32_t taddr) {
  uint32_t r[64];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,"
      "%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]), "=r"(r[32]),
        "=r"(r[33]), "=r"(r[34]), "=r"(r[35]), "=r"(r[36]), "=r"(r[37]), "=r"(r[38]), "=r"(r[39]), "=r"(r[40]),
        "=r"(r[41]), "=r"(r[42]), "=r"(r[43]), "=r"(r[44]), "=r"(r[45]), "=r"(r[46]), "=r"(r[47]), "=r"(r[48]),
        "=r"(r[49]), "=r"(r[50]), "=r"(r[51]), "=r"(r[52]), "=r"(r[53]), "=r"(r[54]), "=r"(r[55]), "=r"(r[56]),
        "=r"(r[57]), "=r"(r[58]), "=r"(r[59]), "=r"(r[60]), "=r"(r[61]), "=r"(r[62]), "=r"(r[63]) : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  return r[0] ^ r[sizeof(r) / 4 - 1];
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (++spins == (1u << 24)) __trap();
  }
}
template <int COLS>
__global__ void __launch_bounds__(640, 1) k(long long* out, int mma_iters, int ld_warps, int ld_iters, int n_cols) {
  extern __shared__ __align__(1024) unsigned char smem_dyn[];
  __shared__ uint64_t done_bar;
  __shared__ uint32_t base_slot;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (16384 + 32768) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&done_bar)), "r"(1) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = base_slot;
  if (warp == 0) {
    if (mma_iters > 0) {
      // idesc: D=f32 [4,6)=1; A,B bf16 (1 at [7,10),[10,13)); N>>3 at [17,23); M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n_cols >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t adesc = make_sw128_desc(smem_u32(smem)), bdesc = make_sw128_desc(smem_u32(smem + 16384));
      const long long t0 = clock64();
      uint32_t pred;
      asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
      if (pred) {
        for (int it = 0; it < mma_iters; ++it) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks)
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                         ::"r"(tmem_base), "l"(adesc + (uint64_t)(ks * 2)), "l"(bdesc + (uint64_t)(ks * 2)), "r"(idesc), "r"((uint32_t)((it | ks) != 0)) : "memory");
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&done_bar)) : "memory");
      }
      __syncwarp();
      mbar_wait(&done_bar, 0);
      if (lane == 0) out[blockIdx.x * 4 + 0] = clock64() - t0;
    }
  } else if (warp >= 4 && warp < 4 + ld_warps) {
    const uint32_t base = tmem_base + 256 + ((uint32_t)((warp & 3) * 32) << 16);
    const long long t0 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < ld_iters; ++i) {
      acc ^= tld<COLS>(base + (uint32_t)((i * COLS) & 255 & ~(COLS - 1)));
    }
    if (warp == 4 && lane == 0) out[blockIdx.x * 4 + 1] = clock64() - t0;
    if (acc == 0x12345u) out[blockIdx.x * 4 + 2] = acc;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 2) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
}
template <int COLS>
void run(long long* d, int mma_iters, int ld_warps, int ld_iters, int n_cols) {
  cudaFuncSetAttribute(k<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
  cudaMemset(d, 0, 148 * 4 * sizeof(long long));
  k<COLS><<<148, 640, 64 * 1024>>>(d, mma_iters, ld_warps, ld_iters, n_cols);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[4];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  printf("N=%3d x%-2d mma_iters %5d ld_warps %2d ld_iters %5d : mma %8lld clk (%.1f clk/MMA)  ld %8lld clk (%.1f clk/ld, %.0f B/clk/SM)  %s\n",
         n_cols, COLS, mma_iters, ld_warps, ld_iters, h[0], mma_iters ? h[0] / (4.0 * mma_iters) : 0.0, h[1], ld_iters ? (double)h[1] / ld_iters : 0.0,
         h[1] ? (double)ld_warps * ld_iters * COLS * 128 / (double)h[1] : 0.0, cudaGetErrorString(e));
}
int main() {
  long long* d;
  cudaMalloc(&d, 148 * 4 * sizeof(long long));
  for (int n : {256, 160}) {
    run<16>(d, 2000, 0, 0, n);
    run<16>(d, 2000, 0, 0, n);
  }
  run<16>(d, 0, 16, 4000, 256);
  run<32>(d, 0, 16, 2000, 256);
  run<64>(d, 0, 16, 1000, 256);
  for (int n : {256, 160}) {
    // matched amount of work: each MMA iteration (4 UMMAs, one 128 x N x 64 block) ~ 4*N/2 cycles
    run<16>(d, 2000, 16, 4000, n);
    run<32>(d, 2000, 16, 2000, n);
    run<64>(d, 2000, 16, 1000, n);
    run<16>(d, 2000, 4, 4000, n);
    run<32>(d, 2000, 4, 2000, n);
    run<64>(d, 2000, 4, 1000, n);
    run<32>(d, 2000, 16, 200, n);   // light drain: one 128 x 256 tile per 10 blocks of MMAs... scaled
  }
  return 0;
}
