// Issue rates of the instructions an attention softmax is made of, per SM (all four sub-partitions busy):
// ex2.approx f32 / f16x2 / bf16x2 (MUFU), 3-input max, packed fma.rn.f32x2, cvt.rn.bf16x2.f32 / f16x2.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o softmax_pipes softmax_pipes.cu && ./softmax_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

constexpr int kChains = 8;      // independent dependency chains per thread
constexpr int kIters = 4096;

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t y) {
  uint32_t r;
  if constexpr (OP == 0) {
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(r) : "r"(x));
  } else if constexpr (OP == 1) {
    asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(r) : "r"(x));
  } else if constexpr (OP == 2) {
    asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(r) : "r"(x));
  } else if constexpr (OP == 3) {
    asm volatile("max.f32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(y), "r"(x ^ 0x3f));
  } else if constexpr (OP == 4) {
    asm volatile("max.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
  } else if constexpr (OP == 5) {
    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
  } else if constexpr (OP == 6) {
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(y));
  } else if constexpr (OP == 7) {
    asm volatile("fma.rn.ftz.f32 %0, %1, %2, %1;" : "=r"(r) : "r"(x), "r"(y));
  } else if constexpr (OP == 9) {
    asm volatile("tanh.approx.f32 %0, %1;" : "=r"(r) : "r"(x));
  } else {
    r = x;
  }
  return r;
}

template <int OP>
__global__ void __launch_bounds__(1024, 1) k(long long* out, uint32_t seed) {
  uint32_t v[kChains];
  unsigned long long w[kChains / 2];
#pragma unroll
  for (int j = 0; j < kChains; ++j) v[j] = seed + threadIdx.x * 7 + j;
#pragma unroll
  for (int j = 0; j < kChains / 2; ++j) w[j] = (unsigned long long)(seed + j) << 20;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < kIters; ++i) {
    if constexpr (OP == 8) {
#pragma unroll
      for (int j = 0; j < kChains / 2; ++j) asm volatile("fma.rn.f32x2 %0, %0, %1, %0;" : "+l"(w[j]) : "l"(w[(j + 1) % (kChains / 2)]));
    } else {
#pragma unroll
      for (int j = 0; j < kChains; ++j) v[j] = op<OP>(v[j], seed);
    }
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < kChains; ++j) acc ^= v[j];
#pragma unroll
  for (int j = 0; j < kChains / 2; ++j) acc ^= (uint32_t)w[j];
  if (threadIdx.x == 0) out[blockIdx.x * 2] = t1 - t0;
  if (acc == 0x12345679) out[blockIdx.x * 2 + 1] = acc;
}

template <int OP>
void run(const char* name, int elems_per_op, long long* d) {
  for (int threads : {128, 256, 512, 1024}) {
    k<OP><<<148, threads>>>(d, 0x3c003c00u);
    cudaDeviceSynchronize();
    k<OP><<<148, threads>>>(d, 0x3c003c00u);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[2];
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const int per_thread = (OP == 8) ? kChains / 2 : kChains;
    const double ops = (double)kIters * per_thread * threads;      // thread-level instructions per SM
    printf("%-28s threads %4d: %7.1f thread-instr/clk/SM, %7.1f elements/clk/SM  (%s)\n", name, threads, ops / h[0],
           ops * elems_per_op / h[0], cudaGetErrorString(e));
  }
}

int main() {
  long long* d;
  cudaMalloc(&d, 148 * 2 * sizeof(long long));
  run<0>("ex2.approx.ftz.f32", 1, d);
  run<1>("ex2.approx.f16x2", 2, d);
  run<2>("ex2.approx.ftz.bf16x2", 2, d);
  run<9>("tanh.approx.f32", 1, d);
  run<3>("max.f32 (3 inputs)", 2, d);
  run<4>("max.f32 (2 inputs)", 1, d);
  run<5>("cvt.rn.bf16x2.f32", 2, d);
  run<6>("cvt.rn.f16x2.f32", 2, d);
  run<7>("fma.rn.ftz.f32", 1, d);
  run<8>("fma.rn.f32x2", 2, d);
  return 0;
}
