// Shared device/host helpers for libmedseg_b200 (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/medseg_b200.h"

namespace msb {

void set_error(const char* fmt, ...);

#define MSB_REQUIRE(cond, ...)            \
  do {                                    \
    if (!(cond)) {                        \
      ::msb::set_error(__VA_ARGS__);      \
      return MSB_ERR_INVALID;             \
    }                                     \
  } while (0)

#define MSB_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::msb::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MSB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define MSB_LAUNCH_OK()                                                                     \
  do {                                                                                      \
    cudaError_t _e = cudaGetLastError();                                                    \
    if (_e != cudaSuccess) {                                                                \
      ::msb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MSB_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

constexpr int kNumSMs = 148;

// ---- programmatic dependent launch (PDL) ---------------------------------------------------------------
// Every kernel of the train step is launched with the programmatic-stream-serialization attribute: kernel N+1 may be
// scheduled while kernel N drains, runs its prologue (barrier init, TMEM allocation, descriptor prefetch, parameter
// loads from the kernel arguments) and blocks in pdl_wait() until kernel N has completed and its writes are visible.
// RULES: (1) a kernel launched through MSB_LAUNCH_PDL calls pdl_wait() before its first global-memory access;
// (2) pdl_trigger() is issued only AFTER the kernel's own TMEM allocation (a dependent CTA that became resident first
// and grabbed the TMEM columns would otherwise deadlock its predecessor).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

extern int g_pdl_enabled;  // msb_debug_set(7, 1) disables PDL (A/B measurements)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// same, as thread-block clusters of `csize` CTAs along x (grid.x must be a multiple of csize)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                      int csize, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = g_pdl_enabled;
  attr[1].id = cudaLaunchAttributeClusterDimension;
  attr[1].val.clusterDim.x = (unsigned)csize;
  attr[1].val.clusterDim.y = 1;
  attr[1].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

#define MSB_LAUNCH_PDL(kernel, grid, block, smem, st, ...)                                           \
  do {                                                                                               \
    cudaError_t _e = ::msb::launch_pdl(kernel, grid, block, smem, st, __VA_ARGS__);                  \
    if (_e != cudaSuccess) {                                                                         \
      ::msb::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return MSB_ERR_CUDA;                                                                           \
    }                                                                                                \
  } while (0)

// ---- 8-channel vector load/store in the B8 layout -------------------------------------------------
template <typename T>
struct Vec8;

template <>
struct Vec8<float> {
  static __device__ __forceinline__ void load(const float* p, float (&v)[8]) {
    float4 a = *reinterpret_cast<const float4*>(p);
    float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
  static __device__ __forceinline__ void store(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
  static __device__ __forceinline__ float round(float x) { return x; }
};

template <>
struct Vec8<__nv_bfloat16> {
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&v)[8]) {
    uint4 raw = *reinterpret_cast<const uint4*>(p);
    const uint32_t r[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      v[2 * i] = __uint_as_float(r[i] << 16);
      v[2 * i + 1] = __uint_as_float(r[i] & 0xffff0000u);
    }
  }
  static __device__ __forceinline__ void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t r[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      r[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(r[0], r[1], r[2], r[3]);
  }
  static __device__ __forceinline__ float round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }
};

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

__device__ __forceinline__ float prelu(float x, float a) { return x > 0.f ? x : a * x; }

// typed pointer of a view for (n, c8 plane, voxel v)
template <typename T>
__device__ __forceinline__ T* view_ptr(const msb_tensor& t, int n, int c8, int64_t s, int64_t v) {
  return reinterpret_cast<T*>(t.ptr) + (int64_t)n * t.n_stride + ((int64_t)c8 * s + v) * 8;
}

inline bool view_ok(const msb_tensor& t) {
  return t.ptr != nullptr && t.c > 0 && (t.c % 8) == 0 && (t.dtype == MSB_F32 || t.dtype == MSB_BF16) &&
         (reinterpret_cast<uintptr_t>(t.ptr) % 16 == 0);
}

}  // namespace msb

// dtype dispatch: binds T to float or __nv_bfloat16
#define MSB_DISPATCH_DTYPE(dt, ...)                         \
  do {                                                      \
    if ((dt) == MSB_F32) {                                  \
      using T = float;                                      \
      __VA_ARGS__                                           \
    } else {                                                \
      using T = __nv_bfloat16;                              \
      __VA_ARGS__                                           \
    }                                                       \
  } while (0)
