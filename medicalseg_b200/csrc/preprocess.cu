// GPU rewrite of tools/preprocess_utils/values.py:37-87 (label_remap, normalize, HUnorm) and
// geometry.py:31-69 (resample = scipy.ndimage.zoom, mode='nearest', order 0/1, align-corners mapping).
// All HBM-bound: vectorised streaming for the pointwise ops, one fused gather for (HUnorm|normalize)+resample.
#include <float.h>

#include "common.cuh"

namespace msb {

__device__ __forceinline__ float hunorm_point(float x, float hu_min, float div, float hu_nan) {
  x = isnan(x) ? hu_nan : x;                  // np.nan_to_num(nan=HU_nan)  (values.py:81)
  x = __fdiv_rn(x - hu_min, div);             // (image - HU_min) / ((HU_max - HU_min) / 255)  (:84)
  return fminf(fmaxf(x, 0.f), 255.f);         // np.clip(0, 255) (:85)
}

__device__ __forceinline__ float normalize_point(float x, float lo, float range) {
  x = __fdiv_rn(x - lo, range);               // values.py:58-61
  return fminf(fmaxf(x, 0.f), 1.f);           // np.clip(0, 1) (:62)
}

__global__ void __launch_bounds__(256) hunorm_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                     int64_t count, float hu_min, float div, float hu_nan) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t n4 = count >> 2;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
    v.x = hunorm_point(v.x, hu_min, div, hu_nan);
    v.y = hunorm_point(v.y, hu_min, div, hu_nan);
    v.z = hunorm_point(v.z, hu_min, div, hu_nan);
    v.w = hunorm_point(v.w, hu_min, div, hu_nan);
    reinterpret_cast<float4*>(dst)[i] = v;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = hunorm_point(src[i], hu_min, div, hu_nan);
}

// order-preserving float <-> uint mapping for atomic min/max
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

__global__ void minmax_init_kernel(uint32_t* ord) {
  ord[0] = 0xffffffffu;  // min
  ord[1] = 0u;           // max
}

__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ src, int64_t count,
                                                     uint32_t* __restrict__ ord) {
  float lo = FLT_MAX, hi = -FLT_MAX;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float v = __ldg(src + i);
    if (!isnan(v)) { lo = fminf(lo, v); hi = fmaxf(hi, v); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
    hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(ord, f2ord(lo));
    atomicMax(ord + 1, f2ord(hi));
  }
}

__global__ void minmax_finish_kernel(uint32_t* ord) {
  float* f = reinterpret_cast<float*>(ord);
  const float lo = ord2f(ord[0]), hi = ord2f(ord[1]);
  f[0] = lo;
  f[1] = hi;
}

__global__ void __launch_bounds__(256) normalize_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                        int64_t count, float lo, float hi,
                                                        const float* __restrict__ minmax) {
  if (minmax) { lo = minmax[0]; hi = minmax[1]; }
  const float range = hi - lo;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
    dst[i] = normalize_point(__ldg(src + i), lo, range);
}

// ---- resample ------------------------------------------------------------------------------------
// Per-axis source coordinate in float64 exactly as scipy's zoom: src = dst * (in-1)/(out-1).
__device__ __forceinline__ double axis_coord(int o, int n_in, int n_out) {
  return n_out > 1 ? (double)o * ((double)(n_in - 1) / (double)(n_out - 1)) : 0.0;
}

template <int PRE>
__device__ __forceinline__ float fetch(const float* __restrict__ src, int64_t idx, float p0, float p1, float p2) {
  const float v = __ldg(src + idx);
  if (PRE == 1) return hunorm_point(v, p0, p1, p2);
  if (PRE == 2) return normalize_point(v, p0, p1);
  return v;
}

template <int ORDER, int PRE>
__global__ void __launch_bounds__(256)
    resample_f32_kernel(const float* __restrict__ src, msb_dim3 in, float* __restrict__ dst, msb_dim3 out, float p0,
                        float p1, float p2) {
  // block = (bx <= 256 threads along w, 256 / bx output rows): a 128-wide output row fills half a 256-thread block, so
  // the block takes two rows - no idle half (round 1 launched 256 x 1 threads with 128 active)
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;  // last axis fastest -> coalesced stores
  const int oy = blockIdx.y * blockDim.y + threadIdx.y, oz = blockIdx.z;
  if (ox >= out.w || oy >= out.h) return;
  const double cz = axis_coord(oz, in.d, out.d), cy = axis_coord(oy, in.h, out.h), cx = axis_coord(ox, in.w, out.w);
  const int64_t sy = in.w, sz = (int64_t)in.h * in.w;
  float r;
  if (ORDER == 0) {
    const int iz = min(max((int)floor(cz + 0.5), 0), in.d - 1);
    const int iy = min(max((int)floor(cy + 0.5), 0), in.h - 1);
    const int ix = min(max((int)floor(cx + 0.5), 0), in.w - 1);
    r = fetch<PRE>(src, iz * sz + iy * sy + ix, p0, p1, p2);
  } else {
    const int z0 = (int)floor(cz), y0 = (int)floor(cy), x0 = (int)floor(cx);
    const float tz = (float)(cz - z0), ty = (float)(cy - y0), tx = (float)(cx - x0);
    const int z1 = min(z0 + 1, in.d - 1), y1 = min(y0 + 1, in.h - 1), x1 = min(x0 + 1, in.w - 1);
    const float v000 = fetch<PRE>(src, z0 * sz + y0 * sy + x0, p0, p1, p2);
    const float v001 = fetch<PRE>(src, z0 * sz + y0 * sy + x1, p0, p1, p2);
    const float v010 = fetch<PRE>(src, z0 * sz + y1 * sy + x0, p0, p1, p2);
    const float v011 = fetch<PRE>(src, z0 * sz + y1 * sy + x1, p0, p1, p2);
    const float v100 = fetch<PRE>(src, z1 * sz + y0 * sy + x0, p0, p1, p2);
    const float v101 = fetch<PRE>(src, z1 * sz + y0 * sy + x1, p0, p1, p2);
    const float v110 = fetch<PRE>(src, z1 * sz + y1 * sy + x0, p0, p1, p2);
    const float v111 = fetch<PRE>(src, z1 * sz + y1 * sy + x1, p0, p1, p2);
    const float a00 = fmaf(tx, v001 - v000, v000), a01 = fmaf(tx, v011 - v010, v010);
    const float a10 = fmaf(tx, v101 - v100, v100), a11 = fmaf(tx, v111 - v110, v110);
    const float b0 = fmaf(ty, a01 - a00, a00), b1 = fmaf(ty, a11 - a10, a10);
    r = fmaf(tz, b1 - b0, b0);
  }
  dst[((int64_t)oz * out.h + oy) * out.w + ox] = r;
}

__global__ void __launch_bounds__(256) resample_i32_kernel(const int32_t* __restrict__ src, msb_dim3 in,
                                                           int32_t* __restrict__ dst, msb_dim3 out) {
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y * blockDim.y + threadIdx.y, oz = blockIdx.z;
  if (ox >= out.w || oy >= out.h) return;
  const int iz = min(max((int)floor(axis_coord(oz, in.d, out.d) + 0.5), 0), in.d - 1);
  const int iy = min(max((int)floor(axis_coord(oy, in.h, out.h) + 0.5), 0), in.h - 1);
  const int ix = min(max((int)floor(axis_coord(ox, in.w, out.w) + 0.5), 0), in.w - 1);
  dst[((int64_t)oz * out.h + oy) * out.w + ox] = __ldg(src + ((int64_t)iz * in.h + iy) * in.w + ix);
}

constexpr int kMaxRemap = 64;
struct RemapTable {
  int32_t keys[kMaxRemap], vals[kMaxRemap];
  int n;
};

__global__ void __launch_bounds__(256) label_remap_kernel(int32_t* __restrict__ labels, int64_t count, RemapTable t) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    int32_t x = labels[i];
    for (int k = 0; k < t.n; ++k)  // sequential in-place semantics of values.py:48-49
      if (x == t.keys[k]) x = t.vals[k];
    labels[i] = x;
  }
}

static inline int stream_blocks(int64_t items) {
  int64_t b = (items + 255) / 256;
  const int64_t cap = (int64_t)kNumSMs * 16;
  return (int)(b < cap ? (b > 0 ? b : 1) : cap);
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_hunorm(const float* src, float* dst, int64_t count, float hu_min, float hu_max, float hu_nan, void* stream) {
  MSB_REQUIRE(src && dst && count > 0 && hu_max > hu_min, "msb_hunorm: bad arguments");
  MSB_REQUIRE((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) % 16 == 0,
              "msb_hunorm: buffers must be 16-byte aligned");
  // divisor computed as the reference does: a Python float (double) demoted to the array's float32
  const float div = (float)(((double)hu_max - (double)hu_min) / 255.0);
  hunorm_kernel<<<stream_blocks(count / 4), 256, 0, as_stream(stream)>>>(src, dst, count, hu_min, div, hu_nan);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_minmax(const float* src, int64_t count, float* minmax, void* stream) {
  MSB_REQUIRE(src && minmax && count > 0, "msb_minmax: bad arguments");
  cudaStream_t st = as_stream(stream);
  uint32_t* ord = reinterpret_cast<uint32_t*>(minmax);
  minmax_init_kernel<<<1, 1, 0, st>>>(ord);
  minmax_kernel<<<stream_blocks(count), 256, 0, st>>>(src, count, ord);
  minmax_finish_kernel<<<1, 1, 0, st>>>(ord);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_normalize(const float* src, float* dst, int64_t count, float lo, float hi, const float* minmax, void* stream) {
  MSB_REQUIRE(src && dst && count > 0, "msb_normalize: bad arguments");
  normalize_kernel<<<stream_blocks(count), 256, 0, as_stream(stream)>>>(src, dst, count, lo, hi, minmax);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

static inline dim3 resample_block(int out_w) {
  int bx = (out_w + 31) / 32 * 32;
  if (bx > 256) bx = 256;
  return dim3((unsigned)bx, (unsigned)(256 / bx), 1);
}

int msb_resample_f32(const float* src, msb_dim3 in_dims, float* dst, msb_dim3 out_dims, int order, int pre_op,
                     float p0, float p1, float p2, void* stream) {
  MSB_REQUIRE(src && dst && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 && out_dims.h > 0 &&
                  out_dims.w > 0 && out_dims.h <= 65535 && out_dims.d <= 65535,
              "msb_resample_f32: bad dims");
  MSB_REQUIRE((order == 0 || order == 1) && pre_op >= 0 && pre_op <= 2, "msb_resample_f32: order in {0,1}, pre_op in {0,1,2}");
  float q0 = p0, q1 = p1, q2 = p2;
  if (pre_op == 1) q1 = (float)(((double)p1 - (double)p0) / 255.0);  // (HU_min, HU_max, HU_nan) -> (min, div, nan)
  if (pre_op == 2) q1 = p1 - p0;                                     // (lo, hi) -> (lo, range)
  const dim3 block = resample_block(out_dims.w);
  const dim3 grid((out_dims.w + block.x - 1) / block.x, (out_dims.h + block.y - 1) / block.y, out_dims.d);
  cudaStream_t st = as_stream(stream);
#define MSB_RS(O, P) resample_f32_kernel<O, P><<<grid, block, 0, st>>>(src, in_dims, dst, out_dims, q0, q1, q2)
  if (order == 0) { if (pre_op == 0) MSB_RS(0, 0); else if (pre_op == 1) MSB_RS(0, 1); else MSB_RS(0, 2); }
  else            { if (pre_op == 0) MSB_RS(1, 0); else if (pre_op == 1) MSB_RS(1, 1); else MSB_RS(1, 2); }
#undef MSB_RS
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_resample_i32(const int32_t* src, msb_dim3 in_dims, int32_t* dst, msb_dim3 out_dims, void* stream) {
  MSB_REQUIRE(src && dst && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 && out_dims.h > 0 &&
                  out_dims.w > 0 && out_dims.h <= 65535 && out_dims.d <= 65535,
              "msb_resample_i32: bad dims");
  const dim3 block = resample_block(out_dims.w);
  const dim3 grid((out_dims.w + block.x - 1) / block.x, (out_dims.h + block.y - 1) / block.y, out_dims.d);
  resample_i32_kernel<<<grid, block, 0, as_stream(stream)>>>(src, in_dims, dst, out_dims);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_label_remap(int32_t* labels, int64_t count, const int32_t* keys, const int32_t* vals, int nmap, void* stream) {
  MSB_REQUIRE(labels && count > 0 && nmap >= 0 && nmap <= kMaxRemap && (nmap == 0 || (keys && vals)),
              "msb_label_remap: at most 64 (key,val) pairs (host pointers)");
  RemapTable t;
  t.n = nmap;
  for (int i = 0; i < nmap; ++i) { t.keys[i] = keys[i]; t.vals[i] = vals[i]; }
  label_remap_kernel<<<stream_blocks(count), 256, 0, as_stream(stream)>>>(labels, count, t);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
