// On-device training augmentations (reference: medicalseg/transforms/functional.py:77-110 and the callers in
// transforms/transform.py:46-72,154-167,185-203): plane rotation with scipy.ndimage.rotate(reshape=False,
// mode='constant') semantics, axis flips, and Compose's "divide by the volume maximum".  The reference runs these
// through NumPy / SciPy on the host (a 128^3 rotation takes ~1 s of one core, ~80 train steps of this engine);
// here each is one HBM-bound gather over the volume.  Coordinates are computed in f64 with the operation order of
// SciPy's NI_GeometricTransform (no FMA contraction), so the strict inside test `0 <= c <= n-1` picks the same
// voxels as SciPy does.
#include "common.cuh"

namespace msb {

struct RotParams {
  int d, h, w;
  int axis_a, axis_b;            // rotation plane (axis_a < axis_b), 0 = depth, 1 = height, 2 = width
  double m00, m01, m10, m11;     // [[c, s], [-s, c]]  (scipy: cosdg / sindg of the angle)
  double off0, off1;             // in_center - M @ out_center
  int order;                     // 0 nearest, 1 linear
};

template <typename T>
__device__ __forceinline__ T rot_cast(double v);
template <>
__device__ __forceinline__ float rot_cast<float>(double v) { return (float)v; }
template <>
__device__ __forceinline__ int32_t rot_cast<int32_t>(double v) { return (int32_t)floor(v + 0.5); }  // SciPy rounds integer outputs

template <typename T>
__global__ void __launch_bounds__(256) rotate3d_kernel(const T* __restrict__ src, T* __restrict__ dst, RotParams p,
                                                       T cval) {
  const int64_t total = (int64_t)p.d * p.h * p.w;
  auto sel = [](int ax, int64_t v0, int64_t v1, int64_t v2) { return ax == 0 ? v0 : (ax == 1 ? v1 : v2); };
  const int na = (int)sel(p.axis_a, p.d, p.h, p.w), nb = (int)sel(p.axis_b, p.d, p.h, p.w);
  const int64_t sa = sel(p.axis_a, (int64_t)p.h * p.w, p.w, 1), sb = sel(p.axis_b, (int64_t)p.h * p.w, p.w, 1);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.w);
    const int64_t r = i / p.w;
    const int y = (int)(r % p.h), z = (int)(r / p.h);
    const int ia_out = (int)sel(p.axis_a, z, y, x), ib_out = (int)sel(p.axis_b, z, y, x);
    const double oa = (double)ia_out, ob = (double)ib_out;
    // icoor = 0; icoor += o_a * m[.][0]; icoor += o_b * m[.][1]; icoor += shift
    const double ca = __dadd_rn(__dadd_rn(__dmul_rn(oa, p.m00), __dmul_rn(ob, p.m01)), p.off0);
    const double cb = __dadd_rn(__dadd_rn(__dmul_rn(oa, p.m10), __dmul_rn(ob, p.m11)), p.off1);
    T out = cval;
    if (ca >= 0.0 && ca <= (double)(na - 1) && cb >= 0.0 && cb <= (double)(nb - 1)) {
      const int64_t base = i - (int64_t)ia_out * sa - (int64_t)ib_out * sb;  // the voxel's plane origin
      if (p.order == 0) {
        const int ia = min(max((int)floor(ca + 0.5), 0), na - 1);
        const int ib = min(max((int)floor(cb + 0.5), 0), nb - 1);
        out = __ldg(src + base + ia * sa + ib * sb);
      } else {
        const double fa = floor(ca), fb = floor(cb);
        const double ta = ca - fa, tb = cb - fb;
        const int a0 = (int)fa, b0 = (int)fb;
        const int a1 = min(a0 + 1, na - 1), b1 = min(b0 + 1, nb - 1);  // weight 0 when clamped (c == n-1)
        const double g00 = (double)__ldg(src + base + a0 * sa + b0 * sb);
        const double g01 = (double)__ldg(src + base + a0 * sa + b1 * sb);
        const double g10 = (double)__ldg(src + base + a1 * sa + b0 * sb);
        const double g11 = (double)__ldg(src + base + a1 * sa + b1 * sb);
        const double ua = 1.0 - ta, ub = 1.0 - tb;
        double v = __dmul_rn(__dmul_rn(ua, ub), g00);
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(ua, tb), g01));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(ta, ub), g10));
        v = __dadd_rn(v, __dmul_rn(__dmul_rn(ta, tb), g11));
        out = rot_cast<T>(v);
      }
    }
    dst[i] = out;
  }
}

// np.flip along one axis for any 4-byte element type
__global__ void __launch_bounds__(256) flip3d_kernel(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst,
                                                     int d, int h, int w, int axis) {
  const int64_t total = (int64_t)d * h * w;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int x = (int)(i % w);
    const int64_t r = i / w;
    int y = (int)(r % h), z = (int)(r / h);
    if (axis == 0) z = d - 1 - z;
    else if (axis == 1) y = h - 1 - y;
    else x = w - 1 - x;
    dst[i] = __ldg(src + ((int64_t)z * h + y) * w + x);
  }
}

// Compose (transform.py:67-69): im = im / im.max() if im.max() > 0.  minmax[1] is the device-resident maximum.
__global__ void __launch_bounds__(256) scale_by_max_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                           int64_t count, const float* __restrict__ minmax) {
  const float mx = __ldg(minmax + 1);
  const bool scale = mx > 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = __ldg(src + i);
    dst[i] = scale ? __fdiv_rn(v, mx) : v;
  }
}

static inline int grid_for(int64_t count) {
  const int64_t want = (count + 255) / 256;
  return (int)(want < (int64_t)kNumSMs * 16 ? (want > 0 ? want : 1) : (int64_t)kNumSMs * 16);
}

template <typename T>
static int rotate3d_impl(const char* who, const T* src, T* dst, msb_dim3 dims, int axis_a, int axis_b, double m00,
                         double m01, double m10, double m11, double off0, double off1, int order, T cval,
                         void* stream) {
  MSB_REQUIRE(src && dst && src != dst && dims.d > 0 && dims.h > 0 && dims.w > 0, "%s: bad buffers or dims", who);
  MSB_REQUIRE(axis_a >= 0 && axis_a < axis_b && axis_b <= 2, "%s: the plane must be two sorted axes out of (0, 1, 2)", who);
  MSB_REQUIRE(order == 0 || order == 1, "%s: interpolation order must be 0 or 1", who);
  RotParams p;
  p.d = dims.d; p.h = dims.h; p.w = dims.w; p.axis_a = axis_a; p.axis_b = axis_b;
  p.m00 = m00; p.m01 = m01; p.m10 = m10; p.m11 = m11; p.off0 = off0; p.off1 = off1; p.order = order;
  const int64_t total = (int64_t)dims.d * dims.h * dims.w;
  rotate3d_kernel<T><<<grid_for(total), 256, 0, as_stream(stream)>>>(src, dst, p, cval);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_rotate3d_f32(const float* src, float* dst, msb_dim3 dims, int axis_a, int axis_b, double m00, double m01,
                     double m10, double m11, double off0, double off1, int order, float cval, void* stream) {
  return rotate3d_impl<float>("msb_rotate3d_f32", src, dst, dims, axis_a, axis_b, m00, m01, m10, m11, off0, off1,
                              order, cval, stream);
}

int msb_rotate3d_i32(const int32_t* src, int32_t* dst, msb_dim3 dims, int axis_a, int axis_b, double m00, double m01,
                     double m10, double m11, double off0, double off1, int order, int32_t cval, void* stream) {
  return rotate3d_impl<int32_t>("msb_rotate3d_i32", src, dst, dims, axis_a, axis_b, m00, m01, m10, m11, off0, off1,
                                order, cval, stream);
}

int msb_flip3d(const void* src, void* dst, msb_dim3 dims, int axis, void* stream) {
  MSB_REQUIRE(src && dst && src != dst && dims.d > 0 && dims.h > 0 && dims.w > 0 && axis >= 0 && axis <= 2,
              "msb_flip3d: bad buffers, dims or axis (4-byte elements, out of place)");
  const int64_t total = (int64_t)dims.d * dims.h * dims.w;
  flip3d_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(reinterpret_cast<const uint32_t*>(src),
                                                               reinterpret_cast<uint32_t*>(dst), dims.d, dims.h,
                                                               dims.w, axis);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_scale_by_max(const float* src, float* dst, int64_t count, const float* minmax, void* stream) {
  MSB_REQUIRE(src && dst && minmax && count > 0, "msb_scale_by_max: bad arguments");
  scale_by_max_kernel<<<grid_for(count), 256, 0, as_stream(stream)>>>(src, dst, count, minmax);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
