// w-fold / w-unfold helpers of the folded 5x5x1 convolutions (see conv_k5_umma.cu, "w-folded 5x5x1 convolutions").
//   fold   : F_s(t)[v, (j, c)] = t[v + s*(j-2) e_w, c]            j = 0..4, c < c_real, 5*c_real <= 16 channels (B8 bf16)
//            s = +1 on the single-channel image feeds in_tr.conv1 (vnet.py:67-68); s = -1 on dY of out_tr.conv1
//            (vnet.py:165-166) is the adjoint of `unfold` and feeds its input / weight gradients.
//   unfold : y[v, c] = bias[c] + sum_j P[v + (j-2) e_w, (j, c)]   -> B8 bf16 (+ per-channel BN sums of the rounded y)
// All of it is HBM-bound streaming: one thread per voxel, 16-byte vector accesses.
#include "common.cuh"

namespace msb {

constexpr int kFoldC = 16;  // folded channel count (two 8-channel planes)

// ---- fold from an NCDHW f32 tensor with c_real channels -------------------------------------------------------
__global__ void __launch_bounds__(256) fold_w_f32_kernel(const float* __restrict__ x, int c_real, msb_tensor out, int64_t s,
                                                         int w_ext, int sign) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  const float* xn = x + (int64_t)n * c_real * s;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < s; v += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(v % w_ext);
    float o[kFoldC];
#pragma unroll
    for (int i = 0; i < kFoldC; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ws = w + sign * (j - 2);
      if (ws >= 0 && ws < w_ext) {
        for (int c = 0; c < c_real; ++c) {
          const float val = __ldg(xn + (int64_t)c * s + v + sign * (j - 2));
#pragma unroll
          for (int i = 0; i < kFoldC; ++i)
            if (i == j * c_real + c) o[i] = val;
        }
      }
    }
    float lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { lo[i] = o[i]; hi[i] = o[8 + i]; }
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, 0, s, v), lo);
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, 1, s, v), hi);
  }
}

// ---- fold from a B8 bf16 view whose first c_real (<= 3) channels are real -----------------------------------------
__global__ void __launch_bounds__(256) fold_w_b8_kernel(msb_tensor x, int c_real, msb_tensor out, int64_t s, int w_ext,
                                                        int sign) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < s; v += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(v % w_ext);
    float o[kFoldC];
#pragma unroll
    for (int i = 0; i < kFoldC; ++i) o[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ws = w + sign * (j - 2);
      if (ws >= 0 && ws < w_ext) {
        float a[8];
        Vec8<__nv_bfloat16>::load(view_ptr<__nv_bfloat16>(x, n, 0, s, v + sign * (j - 2)), a);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < c_real) {
#pragma unroll
            for (int i = 0; i < kFoldC; ++i)
              if (i == j * c_real + c) o[i] = a[c];
          }
      }
    }
    float lo[8], hi[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { lo[i] = o[i]; hi[i] = o[8 + i]; }
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, 0, s, v), lo);
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, 1, s, v), hi);
  }
}

// ---- unfold: P (B8, 16 channels, f32 or bf16) -> y (B8 bf16, out.c channels, only c_real non-zero) -----------------
template <typename TP>
__global__ void __launch_bounds__(256) unfold_w_kernel(msb_tensor p, const float* __restrict__ bias, int c_real,
                                                       msb_tensor out, int64_t s, int w_ext, int groups,
                                                       double* __restrict__ sums) {
  pdl_wait();
  pdl_trigger();
  const int n = blockIdx.y;
  float b[3] = {0.f, 0.f, 0.f};
  for (int c = 0; c < c_real; ++c) b[c] = bias != nullptr ? __ldg(bias + c) : 0.f;
  float s1[3] = {0.f, 0.f, 0.f}, s2[3] = {0.f, 0.f, 0.f};
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < s; v += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(v % w_ext);
    float y[3] = {b[0], b[1], b[2]};
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const int ws = w + (j - 2);
      if (ws >= 0 && ws < w_ext) {
        float lo[8], hi[8];
        Vec8<TP>::load(view_ptr<TP>(p, n, 0, s, v + (j - 2)), lo);
#pragma unroll
        for (int i = 0; i < 8; ++i) hi[i] = 0.f;
        if ((j + 1) * c_real > 8) Vec8<TP>::load(view_ptr<TP>(p, n, 1, s, v + (j - 2)), hi);
#pragma unroll
        for (int c = 0; c < 3; ++c)
          if (c < c_real) {
            const int ch = j * c_real + c;
            float val = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (i == ch) val = lo[i];
              if (i + 8 == ch) val = hi[i];
            }
            y[c] += val;
          }
      }
    }
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      if (c < c_real) {
        o[c] = Vec8<__nv_bfloat16>::round(y[c]);
        s1[c] += o[c];
        s2[c] += o[c] * o[c];
      }
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, 0, s, v), o);
    if (out.c > 8) {
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = 0.f;
      for (int c8 = 1; c8 < out.c / 8; ++c8) Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(out, n, c8, s, v), o);
    }
  }
  if (sums != nullptr) {
    __shared__ float red[8][6];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float a = warp_sum(s1[c]), q = warp_sum(s2[c]);
      if (lane == 0) { red[warp][c] = a; red[warp][3 + c] = q; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
      const int stat = threadIdx.x / 3, c = threadIdx.x % 3;
      if (c < c_real) {
        float t = 0.f;
        for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
        const int g = groups > 1 ? n : 0;
        atomicAdd(&sums[((int64_t)stat * groups + g) * out.c + c], (double)t);
      }
    }
  }
}

static inline int fold_blocks(int64_t s) {
  int64_t b = (s + 255) / 256;
  return (int)(b < 148 * 16 ? b : 148 * 16);
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_fold_w_f32(const float* x, int c_real, msb_tensor out, int n, msb_dim3 dims, int sign, void* stream) {
  MSB_REQUIRE(x && view_ok(out) && out.dtype == MSB_BF16 && out.c == kFoldC && n > 0, "msb_fold_w_f32: bf16 B8 16-channel output required");
  MSB_REQUIRE(c_real > 0 && 5 * c_real <= kFoldC, "msb_fold_w_f32: at most 3 channels can be folded");
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0 && (sign == 1 || sign == -1), "msb_fold_w_f32: bad dims / sign");
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  MSB_LAUNCH_PDL(fold_w_f32_kernel, dim3(fold_blocks(s), n), dim3(256), 0, as_stream(stream), x, c_real, out, s, dims.w, sign);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_fold_w(msb_tensor x, int c_real, msb_tensor out, int n, msb_dim3 dims, int sign, void* stream) {
  MSB_REQUIRE(view_ok(x) && x.dtype == MSB_BF16 && view_ok(out) && out.dtype == MSB_BF16 && out.c == kFoldC && n > 0,
              "msb_fold_w: bf16 B8 views (16-channel output) required");
  MSB_REQUIRE(c_real > 0 && c_real <= 3 && c_real <= x.c, "msb_fold_w: at most 3 channels can be folded");
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0 && (sign == 1 || sign == -1), "msb_fold_w: bad dims / sign");
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  MSB_LAUNCH_PDL(fold_w_b8_kernel, dim3(fold_blocks(s), n), dim3(256), 0, as_stream(stream), x, c_real, out, s, dims.w, sign);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_unfold_w(msb_tensor p, const float* bias, int c_real, msb_tensor out, int n, msb_dim3 dims, int groups,
                 double* sums, void* stream) {
  MSB_REQUIRE(view_ok(p) && p.c == kFoldC && view_ok(out) && out.dtype == MSB_BF16 && n > 0,
              "msb_unfold_w: 16-channel B8 input and bf16 B8 output required");
  MSB_REQUIRE(c_real > 0 && c_real <= 3, "msb_unfold_w: at most 3 channels can be unfolded");
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0 && (groups == 1 || groups == n), "msb_unfold_w: bad dims / groups");
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  const dim3 grid(fold_blocks(s), n);
  if (p.dtype == MSB_F32)
    MSB_LAUNCH_PDL(unfold_w_kernel<float>, grid, dim3(256), 0, as_stream(stream), p, bias, c_real, out, s, dims.w, groups, sums);
  else
    MSB_LAUNCH_PDL(unfold_w_kernel<__nv_bfloat16>, grid, dim3(256), 0, as_stream(stream), p, bias, c_real, out, s, dims.w, groups, sums);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
