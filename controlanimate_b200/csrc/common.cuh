// Shared device/host helpers for the controlanimate_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/controlanimate_b200.h"

namespace ca {

// ---- error plumbing (host) ----------------------------------------------------------------
void set_error(const char* fmt, ...);
int sm_count();
// Raise a kernel's dynamic shared-memory limit once per (device, function); cached occupancy query.
cudaError_t ensure_dynamic_smem(const void* func, size_t bytes);
cudaError_t cached_occupancy(int* per_sm, const void* func, int threads, size_t smem);

#define CA_CHECK_ARG(cond, ...)           \
  do {                                    \
    if (!(cond)) {                        \
      ca::set_error(__VA_ARGS__);         \
      return CA_ERR_INVALID;              \
    }                                     \
  } while (0)

#define CA_CUDA(call)                                                                       \
  do {                                                                                      \
    cudaError_t _e = (call);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ca::set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e));   \
      return CA_ERR_CUDA;                                                                   \
    }                                                                                       \
  } while (0)

// ---- dtype traits ---------------------------------------------------------------------------
template <typename T>
struct Traits;
template <>
struct Traits<__nv_bfloat16> {
  static constexpr int kVec = 8;  // elements per 16-byte vector
  __device__ static float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
  __device__ static __nv_bfloat16 from_f(float v) { return __float2bfloat16_rn(v); }
};
template <>
struct Traits<__half> {
  static constexpr int kVec = 8;
  __device__ static float to_f(__half v) { return __half2float(v); }
  __device__ static __half from_f(float v) { return __float2half_rn(v); }
};
template <>
struct Traits<float> {
  static constexpr int kVec = 4;
  __device__ static float to_f(float v) { return v; }
  __device__ static float from_f(float v) { return v; }
};

// 16-byte vector of T, unpacked to / packed from fp32 with register-only bit manipulation (no address of the
// vector is ever taken, so arrays of Vec16 stay in registers).
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi, __nv_bfloat16) {
  lo = __uint_as_float(w << 16);
  hi = __uint_as_float(w & 0xffff0000u);
}
__device__ __forceinline__ void unpack2(uint32_t w, float& lo, float& hi, __half) {
  const __half2 h = *reinterpret_cast<const __half2*>(&w);
  const float2 f = __half22float2(h);
  lo = f.x;
  hi = f.y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi, __nv_bfloat16) {
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi, __half) {
  uint32_t r;
  asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}

template <typename T>
struct Vec16 {
  static constexpr int N = Traits<T>::kVec;
  uint4 raw;
  __device__ __forceinline__ void unpack(float (&f)[N]) const {
    if constexpr (N == 4) {
      f[0] = __uint_as_float(raw.x);
      f[1] = __uint_as_float(raw.y);
      f[2] = __uint_as_float(raw.z);
      f[3] = __uint_as_float(raw.w);
    } else {
      unpack2(raw.x, f[0], f[1], T());
      unpack2(raw.y, f[2], f[3], T());
      unpack2(raw.z, f[4], f[5], T());
      unpack2(raw.w, f[6], f[7], T());
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[N]) {
    if constexpr (N == 4) {
      raw = make_uint4(__float_as_uint(f[0]), __float_as_uint(f[1]), __float_as_uint(f[2]), __float_as_uint(f[3]));
    } else {
      raw = make_uint4(pack2(f[0], f[1], T()), pack2(f[2], f[3], T()), pack2(f[4], f[5], T()), pack2(f[6], f[7], T()));
    }
  }
};

// Streaming 16-byte global accesses: activations on this path are read once and written once.
__device__ __forceinline__ uint4 ldg_stream(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// Same streaming load WITHOUT the read-only (.nc) path: for operands the kernel may also write (in-place epilogues,
// `dst += ...`).  PTX requires .nc data to stay unmodified for the whole kernel, aliasing included.
__device__ __forceinline__ uint4 ldg_stream_rw(const void* p) {
  uint4 r;
  asm volatile("ld.global.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p)
               : "memory");
  return r;
}
__device__ __forceinline__ uint4 ldg_keep(const void* p) {  // plain (L1/L2-allocating) load
  return *reinterpret_cast<const uint4*>(p);
}
__device__ __forceinline__ void stg_stream(void* p, const uint4& v) {
  asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z),
               "r"(v.w)
               : "memory");
}

__device__ __forceinline__ float silu_f(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Block-wide sum; `scratch` holds >= 32 elements. All threads get the result.
template <typename A>
__device__ __forceinline__ A block_sum(A v, A* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect scratch reuse
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  A r = (lane < nw) ? scratch[lane] : A(0);
  r = warp_sum(r);
  return r;
}

template <typename F>
int dispatch_dtype(int dtype, F&& f) {
  switch (dtype) {
    case CA_BF16: return f(__nv_bfloat16());
    case CA_F16: return f(__half());
    case CA_F32: return f(float());
    default: set_error("unknown dtype %d", dtype); return CA_ERR_INVALID;
  }
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace ca
