// Fused Dice + cross-entropy loss over NCDHW f32 logits and int32 labels (one pass forward, one pass backward).
// Replaces the ~8 separate Paddle kernels behind models/losses/dice_loss.py:76-102 and
// cross_entropy_loss.py:47-87 (+ loss_utils.py:31-40 class_weights).
#include "common.cuh"

namespace msb {

constexpr int kLossThreads = 256;
constexpr int kLossVoxPerBlock = 4096;

template <int CMAX>
__device__ __forceinline__ void load_logits(const float* __restrict__ logits, int n, int c, int64_t s, int64_t v,
                                            float (&z)[CMAX]) {
#pragma unroll
  for (int k = 0; k < CMAX; ++k) z[k] = k < c ? __ldg(logits + ((int64_t)n * c + k) * s + v) : -INFINITY;
}

template <int CMAX>
__device__ __forceinline__ float softmax_inplace(float (&z)[CMAX], int c, float& logsum) {
  float m = z[0];
#pragma unroll
  for (int k = 1; k < CMAX; ++k) m = fmaxf(m, z[k]);
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    z[k] = k < c ? expf(z[k] - m) : 0.f;
    sum += z[k];
  }
  logsum = logf(sum);
  const float inv = 1.f / sum;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) z[k] *= inv;
  return m;
}

// block reduce K floats and add them to double accumulators
template <int K>
__device__ __forceinline__ void block_accumulate(float (&acc)[K], int kvalid, double* __restrict__ out) {
  __shared__ float red[kLossThreads / 32][K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float r = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kvalid; i += kLossThreads) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kLossThreads / 32; ++w) t += (double)red[w][i];
    atomicAdd(out + i, t);
  }
}

template <int CMAX>
__global__ void __launch_bounds__(kLossThreads) class_weight_sums_kernel(const float* __restrict__ logits, int c,
                                                                         int64_t s, double* __restrict__ psum) {
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  float acc[CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) acc[k] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    float z[CMAX], ls;
    load_logits<CMAX>(logits, n, c, s, v, z);
    softmax_inplace<CMAX>(z, c, ls);
#pragma unroll
    for (int k = 0; k < CMAX; ++k) acc[k] += z[k];
  }
  block_accumulate<CMAX>(acc, c, psum);
}

__global__ void class_weight_finalize_kernel(const double* __restrict__ psum, double count, int c,
                                             float* __restrict__ w) {
  const int k = threadIdx.x;
  if (k < c) w[k] = (float)((count - psum[k]) / psum[k]);
}

// acc layout: [0,C) I_c ; [C,2C) sum p^2 ; [2C,3C) sum t ; 3C ce_num ; 3C+1 ce_den
template <int CMAX>
__global__ void __launch_bounds__(kLossThreads)
    dice_ce_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                       const float* __restrict__ class_w, int c, int64_t s, int ignore_index,
                       double* __restrict__ out) {
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  float acc[3 * CMAX + 2];
#pragma unroll
  for (int k = 0; k < 3 * CMAX + 2; ++k) acc[k] = 0.f;
  float w[CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) w[k] = k < c ? __ldg(class_w + k) : 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    float z[CMAX];
    load_logits<CMAX>(logits, n, c, s, v, z);
    const int y = __ldg(labels + (int64_t)n * s + v);
    float zy = 0.f, wy = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (k < c) {
        const float p = 1.f / (1.f + expf(-z[k]));
        const bool hit = (y == k);
        acc[k] += hit ? p : 0.f;
        acc[CMAX + k] += p * p;
        acc[2 * CMAX + k] += hit ? 1.f : 0.f;
        if (hit) { zy = z[k]; wy = w[k]; }
      }
    }
    if (y != ignore_index && y >= 0 && y < c) {
      float ls;
      const float m = softmax_inplace<CMAX>(z, c, ls);
      acc[3 * CMAX] += wy * (m + ls - zy);
      acc[3 * CMAX + 1] += wy;
    }
  }
  // compact to the [3C+2] layout while reducing
  __shared__ float red[kLossThreads / 32][3 * CMAX + 2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 3 * CMAX + 2; ++i) {
    float r = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 3 * CMAX + 2; i += kLossThreads) {
    const int grp = i / CMAX, k = i % CMAX;
    int dst;
    if (i >= 3 * CMAX) dst = 3 * c + (i - 3 * CMAX);
    else if (k < c) dst = grp * c + k;
    else continue;
    double t = 0;
#pragma unroll
    for (int wv = 0; wv < kLossThreads / 32; ++wv) t += (double)red[wv][i];
    atomicAdd(out + dst, t);
  }
}

__global__ void dice_ce_finalize_kernel(const double* __restrict__ acc, int c, float* __restrict__ result) {
  if (threadIdx.x != 0) return;
  double mean = 0;
  for (int k = 0; k < c; ++k) {
    double den = acc[c + k] + acc[2 * c + k];
    if (den < 1e-6) den = 1e-6;
    const double d = 2.0 * acc[k] / den;
    result[2 + k] = (float)d;
    mean += d;
  }
  result[0] = (float)(acc[3 * c] / acc[3 * c + 1]);
  result[1] = (float)(1.0 - mean / c);
}

template <int CMAX>
__global__ void __launch_bounds__(kLossThreads)
    dice_ce_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                       const float* __restrict__ class_w, const double* __restrict__ acc, int c, int64_t s,
                       int ignore_index, float coef_ce, float coef_dice, const float* __restrict__ coef_dev,
                       float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  if (coef_dev != nullptr) {
    coef_ce *= __ldg(coef_dev);
    coef_dice *= __ldg(coef_dev + 1);
  }
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  float w[CMAX], ka[CMAX], kb[CMAX];  // d dice_c / dp = ka*t + kb*p
  const float ce_scale = coef_ce / (float)acc[3 * c + 1];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    w[k] = ka[k] = kb[k] = 0.f;
    if (k < c) {
      w[k] = __ldg(class_w + k);
      const double inter = acc[k];
      double den = acc[c + k] + acc[2 * c + k];
      if (den < 1e-6) {
        ka[k] = (float)(2.0 / 1e-6);
      } else {
        ka[k] = (float)(2.0 / den);
        kb[k] = (float)(-4.0 * inter / (den * den));
      }
    }
  }
  const float dscale = -coef_dice / (float)c;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    float z[CMAX], g[CMAX];
    load_logits<CMAX>(logits, n, c, s, v, z);
    const int y = __ldg(labels + (int64_t)n * s + v);
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (k < c) {
        const float p = 1.f / (1.f + expf(-z[k]));
        const float t = (y == k) ? 1.f : 0.f;
        g[k] = dscale * (ka[k] * t + kb[k] * p) * p * (1.f - p);
      }
    }
    if (y != ignore_index && y >= 0 && y < c) {
      float ls, wy = 0.f;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) if (k == y) wy = w[k];
      softmax_inplace<CMAX>(z, c, ls);
      const float f = ce_scale * wy;
#pragma unroll
      for (int k = 0; k < CMAX; ++k)
        if (k < c) g[k] += f * (z[k] - ((y == k) ? 1.f : 0.f));
    }
#pragma unroll
    for (int k = 0; k < CMAX; ++k)
      if (k < c) dlogits[((int64_t)n * c + k) * s + v] = g[k];
  }
}

}  // namespace msb

using namespace msb;

#define MSB_DISPATCH_CMAX(c, ...)                         \
  do {                                                    \
    if ((c) <= 4) { constexpr int CMAX = 4; __VA_ARGS__ } \
    else if ((c) <= 8) { constexpr int CMAX = 8; __VA_ARGS__ } \
    else if ((c) <= 20) { constexpr int CMAX = 20; __VA_ARGS__ } \
    else { constexpr int CMAX = 32; __VA_ARGS__ }         \
  } while (0)

extern "C" {

int msb_class_weight_sums(const float* logits, int n, int c, int64_t s, double* psum, void* stream) {
  MSB_REQUIRE(logits && psum && n > 0 && c > 0 && c <= 32 && s > 0, "msb_class_weight_sums: needs 1 <= C <= 32");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  MSB_DISPATCH_CMAX(c, class_weight_sums_kernel<CMAX><<<grid, kLossThreads, 0, as_stream(stream)>>>(logits, c, s, psum););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_class_weight_finalize(const double* psum, double count, int c, float* weights, void* stream) {
  MSB_REQUIRE(psum && weights && c > 0 && c <= 32 && count > 0, "msb_class_weight_finalize: bad arguments");
  class_weight_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(psum, count, c, weights);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_fwd(const float* logits, const int32_t* labels, const float* class_w, int n, int c, int64_t s,
                    int ignore_index, double* acc, void* stream) {
  MSB_REQUIRE(logits && labels && class_w && acc && n > 0 && c > 0 && c <= 32 && s > 0,
              "msb_dice_ce_fwd: needs 1 <= C <= 32");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  MSB_DISPATCH_CMAX(c, dice_ce_fwd_kernel<CMAX><<<grid, kLossThreads, 0, as_stream(stream)>>>(
                           logits, labels, class_w, c, s, ignore_index, acc););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_finalize(const double* acc, int c, float* result, void* stream) {
  MSB_REQUIRE(acc && result && c > 0 && c <= 32, "msb_dice_ce_finalize: bad arguments");
  dice_ce_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(acc, c, result);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_bwd(const float* logits, const int32_t* labels, const float* class_w, const double* acc, int n, int c,
                    int64_t s, int ignore_index, float coef_ce, float coef_dice, const float* coef_dev, float* dlogits,
                    void* stream) {
  MSB_REQUIRE(logits && labels && class_w && acc && dlogits && n > 0 && c > 0 && c <= 32 && s > 0,
              "msb_dice_ce_bwd: needs 1 <= C <= 32");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  MSB_DISPATCH_CMAX(c, dice_ce_bwd_kernel<CMAX><<<grid, kLossThreads, 0, as_stream(stream)>>>(
                           logits, labels, class_w, acc, c, s, ignore_index, coef_ce, coef_dice, coef_dev, dlogits););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
