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

__global__ void __launch_bounds__(256) trilinear_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                            int64_t nc, int id, int ih, int iw, int od, int oh,
                                                            int ow, float sd, float sh, float sw) {
  const int64_t ovox = (int64_t)od * oh * ow, ivox = (int64_t)id * ih * iw;
  const int64_t total = nc * ovox;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ovox;
    int64_t r = i - c * ovox;
    const int x = (int)(r % ow); r /= ow;
    const int y = (int)(r % oh);
    const int z = (int)(r / oh);
    const AxisMap mz = axis_map(z, sd, id), my = axis_map(y, sh, ih), mx = axis_map(x, sw, iw);
    const float* p = src + c * ivox;
    auto at = [&](int a, int b, int e) { return __ldg(p + ((int64_t)a * ih + b) * iw + e); };
    const float x00 = at(mz.i0, my.i0, mx.i0) * (1.f - mx.t) + at(mz.i0, my.i0, mx.i1) * mx.t;
    const float x01 = at(mz.i0, my.i1, mx.i0) * (1.f - mx.t) + at(mz.i0, my.i1, mx.i1) * mx.t;
    const float x10 = at(mz.i1, my.i0, mx.i0) * (1.f - mx.t) + at(mz.i1, my.i0, mx.i1) * mx.t;
    const float x11 = at(mz.i1, my.i1, mx.i0) * (1.f - mx.t) + at(mz.i1, my.i1, mx.i1) * mx.t;
    const float y0 = x00 * (1.f - my.t) + x01 * my.t;
    const float y1 = x10 * (1.f - my.t) + x11 * my.t;
    dst[i] = y0 * (1.f - mz.t) + y1 * mz.t;
  }
}

// weight with which output index j feeds source index i along one axis
__device__ __forceinline__ float axis_weight(int j, int i, float scale, int n_in) {
  const AxisMap m = axis_map(j, scale, n_in);
  return (m.i0 == i ? 1.f - m.t : 0.f) + (m.i1 == i ? m.t : 0.f);
}

// conservative range of output indices whose footprint can touch source index i (weights outside are exactly 0)
__device__ __forceinline__ void axis_range(int i, float scale, int n_out, int& lo, int& hi) {
  const float inv = 1.f / scale;
  lo = (int)floorf(((float)i - 0.5f) * inv - 0.5f) - 1;
  hi = (int)ceilf(((float)i + 1.5f) * inv - 0.5f) + 1;
  lo = max(lo, 0);
  hi = min(hi, n_out - 1);
  if (i == 0) lo = 0;  // everything that clamps to src = 0
}

__global__ void __launch_bounds__(128) trilinear_bwd_kernel(const float* __restrict__ ddst, float* __restrict__ dsrc,
                                                            int64_t nc, int id, int ih, int iw, int od, int oh,
                                                            int ow, float sd, float sh, float sw) {
  const int64_t ovox = (int64_t)od * oh * ow, ivox = (int64_t)id * ih * iw;
  const int64_t total = nc * ivox;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t c = i / ivox;
    int64_t r = i - c * ivox;
    const int x = (int)(r % iw); r /= iw;
    const int y = (int)(r % ih);
    const int z = (int)(r / ih);
    int zlo, zhi, ylo, yhi, xlo, xhi;
    axis_range(z, sd, od, zlo, zhi);
    axis_range(y, sh, oh, ylo, yhi);
    axis_range(x, sw, ow, xlo, xhi);
    const float* g = ddst + c * ovox;
    float acc = 0.f;
    for (int jz = zlo; jz <= zhi; ++jz) {
      const float wz = axis_weight(jz, z, sd, id);
      if (wz == 0.f) continue;
      float accy = 0.f;
      for (int jy = ylo; jy <= yhi; ++jy) {
        const float wy = axis_weight(jy, y, sh, ih);
        if (wy == 0.f) continue;
        const float* row = g + ((int64_t)jz * oh + jy) * ow;
        float accx = 0.f;
        for (int jx = xlo; jx <= xhi; ++jx) accx = fmaf(axis_weight(jx, x, sw, iw), __ldg(row + jx), accx);
        accy = fmaf(wy, accx, accy);
      }
      acc = fmaf(wz, accy, acc);
    }
    dsrc[i] = acc;
  }
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_trilinear_fwd(const float* src, int64_t nc, msb_dim3 in_dims, float* dst, msb_dim3 out_dims, void* stream) {
  MSB_REQUIRE(src && dst && nc > 0 && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 &&
                  out_dims.h > 0 && out_dims.w > 0,
              "msb_trilinear_fwd: bad arguments");
  const int64_t total = nc * out_dims.d * out_dims.h * out_dims.w;
  const int64_t want = (total + 255) / 256;
  const int blocks = (int)(want < (int64_t)kNumSMs * 16 ? want : (int64_t)kNumSMs * 16);
  trilinear_fwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(
      src, dst, nc, in_dims.d, in_dims.h, in_dims.w, out_dims.d, out_dims.h, out_dims.w,
      (float)in_dims.d / out_dims.d, (float)in_dims.h / out_dims.h, (float)in_dims.w / out_dims.w);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_trilinear_bwd(const float* ddst, int64_t nc, msb_dim3 out_dims, float* dsrc, msb_dim3 in_dims, void* stream) {
  MSB_REQUIRE(ddst && dsrc && nc > 0 && in_dims.d > 0 && in_dims.h > 0 && in_dims.w > 0 && out_dims.d > 0 &&
                  out_dims.h > 0 && out_dims.w > 0,
              "msb_trilinear_bwd: bad arguments");
  const int64_t total = nc * in_dims.d * in_dims.h * in_dims.w;
  const int64_t want = (total + 127) / 128;
  const int blocks = (int)(want < (int64_t)kNumSMs * 32 ? want : (int64_t)kNumSMs * 32);
  trilinear_bwd_kernel<<<blocks, 128, 0, as_stream(stream)>>>(
      ddst, dsrc, nc, in_dims.d, in_dims.h, in_dims.w, out_dims.d, out_dims.h, out_dims.w,
      (float)in_dims.d / out_dims.d, (float)in_dims.h / out_dims.h, (float)in_dims.w / out_dims.w);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
