// Library-wide pieces of the C ABI: error text, device query, tensor-map encoder.
#include <stdarg.h>

#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "tma.cuh"

namespace ca {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

static std::mutex g_cache_mu;

cudaError_t ensure_dynamic_smem(const void* func, size_t bytes) {
  static std::map<std::pair<int, const void*>, size_t> limit;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cache_mu);
  size_t& cur = limit[{dev, func}];
  if (bytes <= cur) return cudaSuccess;
  const cudaError_t e = cudaFuncSetAttribute(func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) cur = bytes;
  return e;
}

cudaError_t cached_occupancy(int* per_sm, const void* func, int threads, size_t smem) {
  static std::map<std::tuple<int, const void*, int, size_t>, int> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cache_mu);
  auto key = std::make_tuple(dev, func, threads, smem);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *per_sm = it->second;
    return cudaSuccess;
  }
  const cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, func, threads, smem);
  if (e == cudaSuccess) cache[key] = *per_sm;
  return e;
}

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  });
  return fn;
}

bool encode_tensor_map(CUtensorMap* map, CUtensorMapDataType dt, int rank, const void* base, const uint64_t* dims,
                       const uint64_t* strides_bytes, const uint32_t* box, CUtensorMapSwizzle swizzle,
                       CUtensorMapL2promotion promo) {
  EncodeTiledFn fn = get_encode_tiled();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled unavailable (driver too old?)");
    return false;
  }
  cuuint64_t gdims[5], gstr[4];
  cuuint32_t gbox[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
    if (i + 1 < rank) gstr[i] = strides_bytes[i];
  }
  const CUresult r = fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstr, gbox, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
              rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
              rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
    return false;
  }
  return true;
}

}  // namespace ca

extern "C" __attribute__((visibility("default"))) const char* ca_version(void) { return "controlanimate_b200 0.1 (sm_100a)"; }
extern "C" __attribute__((visibility("default"))) const char* ca_last_error(void) { return ca::g_err; }
extern "C" __attribute__((visibility("default"))) int ca_device_sm(void) {
  int dev = 0, major = 0, minor = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
  if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
  return major * 10 + minor;
}
