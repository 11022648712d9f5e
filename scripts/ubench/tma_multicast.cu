// Microbenchmark: delivered L2->SM bandwidth of TMA bulk loads, unicast vs cluster multicast.
// Question it answers (GEMM section): is the ~6300 B/clk chip-wide cap a limit on L2 (LTS) reads, which multicast
// relieves, or on bytes delivered to the SMs, which it does not?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tma_mc scripts/ubench/tma_multicast.cu && /tmp/tma_mc
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0,1,0,p;\n}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void remote_arrive(uint64_t* b, uint32_t cta) {
  uint32_t addr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(addr) : "r"(smem_u32(b)), "r"(cta));
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(addr) : "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync() { asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory"); }

constexpr int kStages = 6;
constexpr int kBox = 32768;  // bytes per stage (one "tile")

// mode 0: every CTA loads its own full box (unicast), address = f(cluster id)  -> CTAs of a cluster read the SAME data
// mode 1: every CTA loads 1/csz of the box and multicasts it to all CTAs of the cluster
// mode 2: unicast, every CTA reads DIFFERENT data
__global__ void __launch_bounds__(128) bw_kernel(const uint8_t* __restrict__ src, size_t span, int iters, int mode, unsigned long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t full[kStages], empty[kStages];
  const uint32_t rank = cluster_rank(), csz = cluster_size(), cid = cluster_id();
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], mode == 1 ? csz : 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync();
  if (threadIdx.x == 0) {
    const unsigned long long t0 = clock64();
    const size_t stream = (mode == 2 ? (size_t)blockIdx.x : (size_t)cid);
    const uint32_t slice = kBox / csz;
    for (int i = 0; i < iters + kStages; ++i) {
      if (i < iters) {
        const int s = i % kStages;
        if (i >= kStages) mbar_wait(&empty[s], ((i / kStages) - 1) & 1);
        mbar_expect(&full[s], kBox);
        const uint8_t* g = src + ((stream * 7919u + (size_t)i) * kBox) % span;
        if (mode == 1) {
          const uint16_t mask = (uint16_t)((1u << csz) - 1);
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;"
                       ::"r"(smem_u32(smem + (size_t)s * kBox + rank * slice)), "l"(g + rank * slice), "r"(slice), "r"(smem_u32(&full[s])), "h"(mask) : "memory");
        } else {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(smem + (size_t)s * kBox)), "l"(g), "r"((uint32_t)kBox), "r"(smem_u32(&full[s])) : "memory");
        }
      }
      const int j = i - kStages + 1;  // consume the oldest outstanding stage
      if (j >= 0 && j < iters) {
        const int s = j % kStages;
        mbar_wait(&full[s], (j / kStages) & 1);
        if (mode == 1) { for (uint32_t c = 0; c < csz; ++c) remote_arrive(&empty[s], c); }
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&empty[s])) : "memory");
      }
    }
    cycles[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  cluster_sync();
}

int main() {
  const size_t span = 64ull << 20;  // 64 MB source: L2 resident (126 MB L2)
  uint8_t* src;
  CK(cudaMalloc(&src, span + kBox));
  CK(cudaMemset(src, 1, span + kBox));
  unsigned long long* cyc;
  CK(cudaMalloc(&cyc, 1024 * sizeof(unsigned long long)));
  int sms = 0;
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kStages * kBox + 1024));
  CK(cudaFuncSetAttribute(bw_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  const int iters = 4000;
  printf("SMs %d, box %d B, stages %d, iters %d\n", sms, kBox, kStages, iters);
  for (int csz : {1, 2, 4, 8}) {
    for (int mode : {0, 1, 2}) {
      if (csz == 1 && mode == 1) continue;
      int grid = sms / csz * csz;
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(grid); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = kStages * kBox + 1024;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = csz; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
      cfg.attrs = at; cfg.numAttrs = 1;
      int maxc = 0;
      cudaOccupancyMaxActiveClusters(&maxc, bw_kernel, &cfg);
      if (maxc * csz < grid) { grid = maxc * csz; cfg.gridDim = dim3(grid); }
      float best = 1e30f;
      for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        CK(cudaLaunchKernelEx(&cfg, bw_kernel, (const uint8_t*)src, span, iters, mode, cyc));
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      unsigned long long h[1024], mx = 0;
      CK(cudaMemcpy(h, cyc, grid * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
      for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
      const double delivered = (double)grid * iters * kBox;
      printf("cluster %d mode %d (%s) grid %d: %.3f ms, delivered %.2f TB/s, %.0f B/clk chip (%.1f B/clk/SM), max cycles %llu\n", csz, mode,
             mode == 0 ? "unicast, cluster-shared data" : mode == 1 ? "multicast" : "unicast, distinct data", grid, best,
             delivered / best / 1e9, delivered / (double)mx, delivered / (double)mx / grid, mx);
    }
  }
  return 0;
}
