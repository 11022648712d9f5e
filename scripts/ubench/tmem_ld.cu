// TMEM -> register read throughput (tcgen05.ld) as a function of resident warps and vector width.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_ld tmem_ld.cu && ./tmem_ld
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
template <int COLS>
__device__ __forceinline__ void tld(uint32_t taddr, uint32_t* r);
template <>
__device__ __forceinline__ void tld<16>(uint32_t taddr, uint32_t* r) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                 "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
               : "r"(taddr) : "memory");
}
template <>
__device__ __forceinline__ void tld<32>(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,"
      "%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
}
template <int COLS, int INFLIGHT>
__global__ void __launch_bounds__(1024, 1) k(long long* out, int iters) {
  __shared__ uint32_t base_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&base_slot)), "r"(512) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = base_slot + ((uint32_t)((warp & 3) * 32) << 16);
  uint32_t acc = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    uint32_t r[INFLIGHT][COLS];
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j) tld<COLS>(base + (uint32_t)(((i * INFLIGHT + j) & 7) * COLS), r[j]);
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < INFLIGHT; ++j)
      acc ^= r[j][0] ^ r[j][COLS - 1];
  }
  const long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0) out[blockIdx.x * 2] = t1 - t0;
  if (acc == 0x12345678) out[blockIdx.x * 2 + 1] = acc;
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base_slot), "r"(512) : "memory");
}
template <int COLS, int INFLIGHT>
void run(int warps, long long* d) {
  const int iters = 2000;
  k<COLS, INFLIGHT><<<148, warps * 32>>>(d, iters);
  cudaDeviceSynchronize();
  k<COLS, INFLIGHT><<<148, warps * 32>>>(d, iters);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
  const double bytes = (double)warps * iters * INFLIGHT * COLS * 32 * 4;
  printf("cols x%-2d inflight %d warps %2d: %8lld clk, %.1f B/clk/SM (%s)\n", COLS, INFLIGHT, warps, h[0], bytes / (double)h[0],
         cudaGetErrorString(e));
}
int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  for (int warps : {1, 4, 8, 16, 32}) {
    run<16, 1>(warps, d);
    run<16, 2>(warps, d);
    run<32, 1>(warps, d);
    run<32, 2>(warps, d);
  }
  return 0;
}
