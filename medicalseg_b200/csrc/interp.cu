// Trilinear resize of NCDHW f32 logits for the deep-supervision heads of VNetDeepSup
// (reference: medicalseg/models/vnet_deepsup.py:257-272, `F.interpolate(d, size=x.shape[2:], mode='trilinear')` with
// Paddle's defaults align_corners=False, align_mode=0) and its adjoint for the backward pass.
// Per axis: src = (in/out) * (dst + 0.5) - 0.5, clamped at 0; i0 = int(src), i1 = min(i0 + 1, in - 1), t = src - i0.
// Forward: one thread per output voxel (write-bound).  Backward: one thread per SOURCE voxel gathers every output
// voxel whose footprint touches it, weights recomputed with the forward's own formula - deterministic, no atomics.
#include "common.cuh"

namespace msb {

struct AxisMap {
  int i0, i1;
  float t;
};

__device__ __forceinline__ AxisMap axis_map(int j, float scale, int n_in) {
  float src = scale * ((float)j + 0.5f) - 0.5f;
  src = src < 0.f ? 0.f : src;
  AxisMap m;
  m.i0 = min((int)src, n_in - 1);
  m.i1 = min(m.i0 + 1, n_in - 1);
  m.t = src - (float)m.i0;
  return m;
}

// grid: x = (plane c, block of kTriZ output depths), y = 256-element chunks of the (y, x) plane.  The in-plane maps are
// computed once per thread and re-used for kTriZ depths (one block per depth is bound by the block launch rate: 491 k
// tiny blocks per MRI head), the depth map is block-uniform, and W = 12 rows leave no thread idle.
constexpr int kTriZ = 8;
__global__ void __launch_bounds__(256) trilinear_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int id, int ih, int iw, int od, int oh, int ow, float sd,
                                                            float sh, float sw) {
  const int zblocks = (od + kTriZ - 1) / kTriZ;
  const int64_t c = blockIdx.x / zblocks;
  const int z0 = (int)(blockIdx.x - c * zblocks) * kTriZ;
  const int p = blockIdx.y * 256 + threadIdx.x;
  if (p >= oh * ow) return;
  const int y = p / ow, x = p - y * ow;
  const AxisMap my = axis_map(y, sh, ih), mx = axis_map(x, sw, iw);
  const float* q = src + c * ((int64_t)id * ih * iw);
  const int o00 = my.i0 * iw + mx.i0, o01 = my.i0 * iw + mx.i1, o10 = my.i1 * iw + mx.i0, o11 = my.i1 * iw + mx.i1;
  const float ux = 1.f - mx.t, uy = 1.f - my.t;
  float* out = dst + (c * od + z0) * ((int64_t)oh * ow) + p;
  const int zend = min(kTriZ, od - z0);
  for (int dz = 0; dz < zend; ++dz) {
    const AxisMap mz = axis_map(z0 + dz, sd, id);
    const float* p0 = q + (int64_t)mz.i0 * ih * iw;
    const float* p1 = q + (int64_t)mz.i1 * ih * iw;
    const float x00 = __ldg(p0 + o00) * ux + __ldg(p0 + o01) * mx.t;
    const float x01 = __ldg(p0 + o10) * ux + __ldg(p0 + o11) * mx.t;
    const float x10 = __ldg(p1 + o00) * ux + __ldg(p1 + o01) * mx.t;
    const float x11 = __ldg(p1 + o10) * ux + __ldg(p1 + o11) * mx.t;
    const float y0 = x00 * uy + x01 * my.t;
    const float y1 = x10 * uy + x11 * my.t;
    out[(int64_t)dz * oh * ow] = y0 * (1.f - mz.t) + y1 * mz.t;
  }
}

// weight with which output index j feeds source index i along one axis
__device__ __forceinline__ float axis_weight(int j, int i, float scale, int n_in) {
  const AxisMap m = axis_map(j, scale, n_in);
  return (m.i0 == i ? 1.f - m.t : 0.f) + (m.i1 == i ? m.t : 0.f);
}

// exact range [lo, hi] of output indices that feed source index i with a non-zero weight: a conservative estimate
// from the inverse map, then shrunk by evaluating the forward's own weight formula at its ends
__device__ __forceinline__ void axis_range(int i, float scale, int n_in, int n_out, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1;
  lo = max(lo, 0);
  hi = min(hi, n_out - 1);
  if (i == 0) lo = 0;  // everything that clamps to src = 0
  while (lo <= hi && axis_weight(lo, i, scale, n_in) == 0.f) ++lo;
  while (hi >= lo && axis_weight(hi, i, scale, n_in) == 0.f) --hi;
}

// grid: x = (plane c, source depth z), y = 128-element chunks of the source (y, x) plane
__global__ void __launch_bounds__(128) trilinear_bwd_kernel(const float* __restrict__ ddst, float* __restrict__ dsrc,
                                                            int id, int ih, int iw, int od, int oh, int ow, float sd,
                                                            float sh, float sw) {
  const int64_t c = blockIdx.x / id;
  const int z = (int)(blockIdx.x - c * id);
  const int p = blockIdx.y * 128 + threadIdx.x;
  if (p >= ih * iw) return;
  const int y = p / iw, x = p - y * iw;
  int zlo, zhi, ylo, yhi, xlo, xhi;
  axis_range(z, sd, id, od, zlo, zhi);
  axis_range(y, sh, ih, oh, ylo, yhi);
  axis_range(x, sw, iw, ow, xlo, xhi);
  const float* g = ddst + c * ((int64_t)od * oh * ow);
  float acc = 0.f;
  for (int jz = zlo; jz <= zhi; ++jz) {
    const float wz = axis_weight(jz, z, sd, id);
    float accy = 0.f;
    for (int jy = ylo; jy <= yhi; ++jy) {
      const float wy = axis_weight(jy, y, sh, ih);
      const float* row = g + ((int64_t)jz * oh + jy) * ow;
      float accx = 0.f;
      for (int jx = xlo; jx <= xhi; ++jx) accx = fmaf(axis_weight(jx, x, sw, iw), __ldg(row + jx), accx);
      accy = fmaf(wy, accx, accy);
    }
    acc = fmaf(wz, accy, acc);
  }
  dsrc[(c * id + z) * ((int64_t)ih * iw) + p] = acc;
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_trilinear_fwd(const float* src, int64_t nc, msb_dim3 in_dims, float* dst, msb_dim3 out_dims, void* stream) {
  MSB_REQUIRE(src && dst && nc > 0 && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 &&
                  out_dims.h > 0 && out_dims.w > 0,
              "msb_trilinear_fwd: bad arguments");
  MSB_REQUIRE(nc * out_dims.d < (int64_t)1 << 31 && (int64_t)out_dims.h * out_dims.w <= (int64_t)65535 * 256 &&
                  (int64_t)in_dims.h * in_dims.w < (int64_t)1 << 31,
              "msb_trilinear_fwd: volume too large for the launch grid");
  const dim3 grid((unsigned)(nc * ((out_dims.d + kTriZ - 1) / kTriZ)),
                  (unsigned)(((int64_t)out_dims.h * out_dims.w + 255) / 256));
  trilinear_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(
      src, dst, in_dims.d, in_dims.h, in_dims.w, out_dims.d, out_dims.h, out_dims.w,
      (float)in_dims.d / out_dims.d, (float)in_dims.h / out_dims.h, (float)in_dims.w / out_dims.w);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_trilinear_bwd(const float* ddst, int64_t nc, msb_dim3 out_dims, float* dsrc, msb_dim3 in_dims, void* stream) {
  MSB_REQUIRE(ddst && dsrc && nc > 0 && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 &&
                  out_dims.h > 0 && out_dims.w > 0,
              "msb_trilinear_bwd: bad arguments");
  MSB_REQUIRE(nc * in_dims.d < (int64_t)1 << 31 && (int64_t)in_dims.h * in_dims.w <= (int64_t)65535 * 128 &&
                  (int64_t)out_dims.h * out_dims.w < (int64_t)1 << 31,
              "msb_trilinear_bwd: volume too large for the launch grid");
  const dim3 grid((unsigned)(nc * in_dims.d), (unsigned)(((int64_t)in_dims.h * in_dims.w + 127) / 128));
  trilinear_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(
      ddst, dsrc, in_dims.d, in_dims.h, in_dims.w, out_dims.d, out_dims.h, out_dims.w,
      (float)in_dims.d / out_dims.d, (float)in_dims.h / out_dims.h, (float)in_dims.w / out_dims.w);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
