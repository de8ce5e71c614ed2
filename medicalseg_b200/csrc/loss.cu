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

// max and log-sum-exp of the first c logits, z left untouched (the cross-entropy needs m + logsum - z_y only)
template <int CMAX>
__device__ __forceinline__ float logsumexp(const float (&z)[CMAX], int c, float& sum) {
  float m = z[0];
#pragma unroll
  for (int k = 1; k < CMAX; ++k) m = k < c ? fmaxf(m, z[k]) : m;
  sum = 0.f;
#pragma unroll
  for (int k = 0; k < CMAX; ++k) sum += k < c ? expf(z[k] - m) : 0.f;
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
__global__ void __launch_bounds__(kLossThreads, CMAX > 8 ? 2 : 1)
    dice_ce_fwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                       const float* __restrict__ class_w, int c, int64_t s, int ignore_index, int dice_softmax,
                       double* __restrict__ out) {
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  // Register diet (round 2; ncu: 164 registers -> 12.5 % occupancy at C = 20): only the p^2 sums need one register per
  // class.  The label-dependent sums (I_c, count_c) touch ONE class per voxel, so they live in a per-thread column of
  // shared memory indexed by the label ([class][thread]: conflict free), and the class weights are read from shared
  // memory (broadcast).
  // (static shared memory ends at 48 KB: beyond 20 classes the label sums go back to registers)
  constexpr bool kSmemHit = CMAX <= 20;
  __shared__ float s_hit[2][kSmemHit ? CMAX : 1][kSmemHit ? kLossThreads : 1];
  __shared__ float s_w[CMAX];
  float r_hit[2][kSmemHit ? 1 : CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    if (kSmemHit) { s_hit[0][k][threadIdx.x] = 0.f; s_hit[1][k][threadIdx.x] = 0.f; }
    else { r_hit[0][kSmemHit ? 0 : k] = 0.f; r_hit[1][kSmemHit ? 0 : k] = 0.f; }
  }
  if (threadIdx.x < CMAX) s_w[threadIdx.x] = threadIdx.x < c ? __ldg(class_w + threadIdx.x) : 0.f;
  __syncthreads();
  float psq[CMAX];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) psq[k] = 0.f;
  float ce_num = 0.f, ce_den = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    float z[CMAX];
    load_logits<CMAX>(logits, n, c, s, v, z);
    const int y = __ldg(labels + (int64_t)n * s + v);
    const bool ce_on = y != ignore_index && y >= 0 && y < c;
    // softmax statistics: needed by the cross-entropy term and by DiceLoss(sigmoid_norm=False) (dice_loss.py:41-43)
    float m = 0.f, sum = 1.f, inv = 0.f;
    if (ce_on || dice_softmax) {
      m = logsumexp<CMAX>(z, c, sum);
      inv = 1.f / sum;
    }
    float zy = 0.f;
#pragma unroll
    for (int k = 0; k < CMAX; ++k) {
      if (k < c) {
        const float p = dice_softmax ? expf(z[k] - m) * inv : 1.f / (1.f + expf(-z[k]));
        psq[k] += p * p;
        if (kSmemHit) {
          if (y == k) { zy = z[k]; s_hit[0][k][threadIdx.x] += p; s_hit[1][k][threadIdx.x] += 1.f; }
        } else {
          const bool hit = (y == k);
          if (hit) zy = z[k];
          r_hit[0][kSmemHit ? 0 : k] += hit ? p : 0.f;
          r_hit[1][kSmemHit ? 0 : k] += hit ? 1.f : 0.f;
        }
      }
    }
    if (ce_on) {
      const float wy = s_w[y];
      ce_num += wy * (m + logf(sum) - zy);
      ce_den += wy;
    }
  }
  float acc[3 * CMAX + 2];
#pragma unroll
  for (int k = 0; k < CMAX; ++k) {
    acc[k] = kSmemHit ? s_hit[0][kSmemHit ? k : 0][kSmemHit ? threadIdx.x : 0] : r_hit[0][kSmemHit ? 0 : k];
    acc[CMAX + k] = psq[k];
    acc[2 * CMAX + k] = kSmemHit ? s_hit[1][kSmemHit ? k : 0][kSmemHit ? threadIdx.x : 0] : r_hit[1][kSmemHit ? 0 : k];
  }
  acc[3 * CMAX] = ce_num;
  acc[3 * CMAX + 1] = ce_den;
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

__global__ void dice_ce_finalize_kernel(const double* __restrict__ acc, int c, const float* __restrict__ dice_w,
                                        float* __restrict__ result) {
  if (threadIdx.x != 0) return;
  double mean = 0;
  for (int k = 0; k < c; ++k) {
    double den = acc[c + k] + acc[2 * c + k];
    if (den < 1e-6) den = 1e-6;
    const double wk = dice_w ? (double)dice_w[k] : 1.0;  // dice_loss.py:64-65: intersect = weight * intersect
    const double d = 2.0 * wk * acc[k] / den;
    result[2 + k] = (float)d;
    mean += d;
  }
  result[0] = (float)(acc[3 * c] / acc[3 * c + 1]);
  result[1] = (float)(1.0 - mean / c);
}

template <int CMAX>
__global__ void __launch_bounds__(kLossThreads, CMAX > 8 ? 2 : 1)
    dice_ce_bwd_kernel(const float* __restrict__ logits, const int32_t* __restrict__ labels,
                       const float* __restrict__ class_w, const double* __restrict__ acc, int c, int64_t s,
                       int ignore_index, float coef_ce, float coef_dice, const float* __restrict__ coef_dev,
                       const float* __restrict__ dice_w, int dice_softmax, float* __restrict__ dlogits) {
  const int n = blockIdx.y;
  if (coef_dev != nullptr) {
    coef_ce *= __ldg(coef_dev);
    coef_dice *= __ldg(coef_dev + 1);
  }
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  // per-class coefficients in shared memory (broadcast reads) instead of 3 x CMAX registers: d dice_c / dp = ka*t + kb*p
  __shared__ float w[CMAX], ka[CMAX], kb[CMAX];
  const float ce_scale = coef_ce / (float)acc[3 * c + 1];
  if (threadIdx.x < CMAX) {
    const int k = threadIdx.x;
    float wk = 0.f, a = 0.f, b = 0.f;
    if (k < c) {
      wk = __ldg(class_w + k);
      const double dw = dice_w ? (double)__ldg(dice_w + k) : 1.0;
      const double inter = acc[k];
      double den = acc[c + k] + acc[2 * c + k];
      if (den < 1e-6) {
        a = (float)(2.0 * dw / 1e-6);
      } else {
        a = (float)(2.0 * dw / den);
        b = (float)(-4.0 * dw * inter / (den * den));
      }
    }
    w[k] = wk; ka[k] = a; kb[k] = b;
  }
  __syncthreads();
  const float dscale = -coef_dice / (float)c;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    // one live array (z): softmax statistics first, then every class's gradient is formed and stored at once
    float z[CMAX];
    load_logits<CMAX>(logits, n, c, s, v, z);
    const int y = __ldg(labels + (int64_t)n * s + v);
    const bool ce_on = y != ignore_index && y >= 0 && y < c;
    float m = 0.f, inv = 0.f, f = 0.f;
    if (ce_on || dice_softmax) {
      float sum;
      m = logsumexp<CMAX>(z, c, sum);
      inv = 1.f / sum;
    }
    if (ce_on) f = ce_scale * w[y];
    if (dice_softmax) {
      // p = softmax(z): dL/dz_k = p_k (dL/dp_k - sum_j dL/dp_j p_j)
      float dot = 0.f;
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (k < c) {
          const float p = expf(z[k] - m) * inv;
          dot += dscale * (ka[k] * ((y == k) ? 1.f : 0.f) + kb[k] * p) * p;
        }
      }
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (k < c) {
          const float p = expf(z[k] - m) * inv;
          const float t = (y == k) ? 1.f : 0.f;
          float g = p * (dscale * (ka[k] * t + kb[k] * p) - dot);
          if (ce_on) g += f * (p - t);
          dlogits[((int64_t)n * c + k) * s + v] = g;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < CMAX; ++k) {
        if (k < c) {
          const float p = 1.f / (1.f + expf(-z[k]));
          const float t = (y == k) ? 1.f : 0.f;
          float g = dscale * (ka[k] * t + kb[k] * p) * p * (1.f - p);
          if (ce_on) g += f * (expf(z[k] - m) * inv - t);
          dlogits[((int64_t)n * c + k) * s + v] = g;
        }
      }
    }
  }
}

// ---- fused evaluation head ---------------------------------------------------------------------------
// core/infer.py:79-92 (argmax of the logits) + core/val.py:101-118 (loss of the same logits) behind out_tr's 1x1x1
// conv (vnet.py:173-174): the logits of a voxel live only in registers - they are never written to HBM (252 MB per
// MRI volume in f32).  Same fmaf order as conv1x1_fwd_kernel, so the logits are bit-identical to the unfused path.
// Optional outputs: pred (int32 argmax, first maximum wins like paddle/torch argmax), acc ([3C+2] as
// dice_ce_fwd_kernel), psum ([C] softmax sums for the CE class weights, loss_utils.py:31-40).
// MODE 0: prediction only; 1: prediction + loss sums; 2: class-weight softmax sums only (separate instantiations
// keep the 3*CMAX+2 loss accumulators and the CMAX softmax sums out of each other's register budget)
template <typename T, int CI8, int CMAX, int MODE>
__global__ void __launch_bounds__(kLossThreads)
    eval_head_kernel(msb_tensor a, const float* __restrict__ w2, const float* __restrict__ b2,
                     const int32_t* __restrict__ labels, const float* __restrict__ class_w, int c, int64_t s,
                     int ignore_index, int32_t* __restrict__ pred, double* __restrict__ out,
                     double* __restrict__ psum) {
  constexpr int kAcc = MODE == 1 ? 3 * CMAX + 2 : (MODE == 2 ? CMAX : 1);
  __shared__ __align__(16) float wsm[CMAX * CI8 * 8 + 2 * CMAX];
  __shared__ float red[kLossThreads / 32][kAcc];
  float* bsm = wsm + CMAX * CI8 * 8;
  float* cwsm = bsm + CMAX;
  const int ci = c;  // the 1x1x1 conv is C -> C
  for (int i = threadIdx.x; i < CMAX * CI8 * 8; i += kLossThreads) {  // transposed: wsm[input j][output o]
    const int j = i / CMAX, o = i % CMAX;
    wsm[i] = (o < c && j < ci) ? w2[o * ci + j] : 0.f;
  }
  for (int i = threadIdx.x; i < CMAX; i += kLossThreads) {
    bsm[i] = (b2 != nullptr && i < c) ? b2[i] : 0.f;
    cwsm[i] = (MODE == 1 && i < c) ? class_w[i] : 0.f;
  }
  __syncthreads();
  const int n = blockIdx.y;
  const int64_t v0 = (int64_t)blockIdx.x * kLossVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kLossVoxPerBlock, s);
  float acc[kAcc];
#pragma unroll
  for (int k = 0; k < kAcc; ++k) acc[k] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kLossThreads) {
    // z[o] = b[o] + sum_i w[o][i] * x[i], i ascending per output (the order of conv1x1_fwd_kernel).  One 8-channel
    // group per (rolled) iteration: a fully unrolled C x C product makes ptxas give up on register allocation.
    float z[CMAX];
#pragma unroll
    for (int o = 0; o < CMAX; ++o) z[o] = bsm[o];
    const int groups = (ci + 7) >> 3;
#pragma unroll 1
    for (int k = 0; k < groups; ++k) {
      float t[8];
      Vec8<T>::load(view_ptr<T>(a, n, k, s, v), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wrow = reinterpret_cast<const float4*>(wsm + (k * 8 + j) * CMAX);
#pragma unroll
        for (int o4 = 0; o4 < CMAX / 4; ++o4) {
          const float4 w4 = wrow[o4];  // rows j >= ci are zero: fmaf(0, finite, z) == z
          z[o4 * 4 + 0] = fmaf(w4.x, t[j], z[o4 * 4 + 0]);
          z[o4 * 4 + 1] = fmaf(w4.y, t[j], z[o4 * 4 + 1]);
          z[o4 * 4 + 2] = fmaf(w4.z, t[j], z[o4 * 4 + 2]);
          z[o4 * 4 + 3] = fmaf(w4.w, t[j], z[o4 * 4 + 3]);
        }
      }
    }
#pragma unroll
    for (int o = 0; o < CMAX; ++o)
      if (o >= c) z[o] = -INFINITY;
    if constexpr (MODE != 2) {
      int best = 0;
      float zbest = z[0];
#pragma unroll
      for (int o = 1; o < CMAX; ++o)
        if (z[o] > zbest) { zbest = z[o]; best = o; }
      if (pred != nullptr) pred[(int64_t)n * s + v] = best;
    }
    if constexpr (MODE == 1) {
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
          if (hit) { zy = z[k]; wy = cwsm[k]; }
        }
      }
      if (y != ignore_index && y >= 0 && y < c) {
        float ls;
        const float m = softmax_inplace<CMAX>(z, c, ls);
        acc[3 * CMAX] += wy * (m + ls - zy);
        acc[3 * CMAX + 1] += wy;
      }
    }
    if constexpr (MODE == 2) {
      float ls;
      softmax_inplace<CMAX>(z, c, ls);
#pragma unroll
      for (int k = 0; k < CMAX; ++k) acc[k] += z[k];
    }
  }
  if constexpr (MODE == 0) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) {
    const float r = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = r;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kAcc; i += kLossThreads) {
    int dst;
    if constexpr (MODE == 1) {
      const int grp = i / CMAX, k = i % CMAX;
      if (i >= 3 * CMAX) dst = 3 * c + (i - 3 * CMAX);
      else if (k < c) dst = grp * c + k;
      else continue;
    } else {
      if (i >= c) continue;
      dst = i;
    }
    double t = 0;
#pragma unroll
    for (int wv = 0; wv < kLossThreads / 32; ++wv) t += (double)red[wv][i];
    atomicAdd((MODE == 1 ? out : psum) + dst, t);
  }
}

// paddle.argmax(logit, axis=1, keepdim=True, dtype='int32') (core/infer.py:92): first maximum wins.  One thread per
// voxel; consecutive threads read consecutive voxels of every class plane (coalesced), writes int32.
__global__ void argmax_channels_kernel(const float* __restrict__ logits, int c, int64_t s, int32_t* __restrict__ pred) {
  const int n = blockIdx.y;
  const float* src = logits + (int64_t)n * c * s;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < s; v += (int64_t)gridDim.x * blockDim.x) {
    int best = 0;
    float zbest = __ldg(src + v);
    for (int o = 1; o < c; ++o) {
      const float z = __ldg(src + (int64_t)o * s + v);
      if (z > zbest) { zbest = z; best = o; }
    }
    pred[(int64_t)n * s + v] = best;
  }
}

// nn.Dropout3D(p) channel masks (vnet.py:103,144-145): out[i] = 0 or 1/(1-p) per (sample, channel), all dropout sites
// of one forward in ONE launch.  Counter-based generator (splitmix64 finaliser of (seed, step, i)); the step counter
// lives in device memory and is advanced by the kernel itself, so a CUDA-graph replay draws fresh masks every step.
__global__ void dropout_masks_kernel(uint64_t seed, unsigned long long* __restrict__ step_counter, float* __restrict__ out,
                                     int total, float p, float keep_scale) {
  const unsigned long long step = *step_counter;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (uint64_t)(step * 0x100000001B3ull + (uint64_t)i + 1ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    const float u = (float)(z >> 40) * (1.0f / 16777216.0f);  // 24 random bits -> [0, 1)
    out[i] = u >= p ? keep_scale : 0.f;
  }
  // every thread has read the counter before any block can finish: a single-block launch is required by the wrapper
  __syncthreads();
  if (blockIdx.x == 0 && threadIdx.x == 0) *step_counter = step + 1ull;
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

int msb_dice_ce_fwd_ex(const float* logits, const int32_t* labels, const float* class_w, int n, int c, int64_t s,
                       int ignore_index, int dice_softmax, double* acc, void* stream) {
  MSB_REQUIRE(logits && labels && class_w && acc && n > 0 && c > 0 && c <= 32 && s > 0,
              "msb_dice_ce_fwd: needs 1 <= C <= 32");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  MSB_DISPATCH_CMAX(c, dice_ce_fwd_kernel<CMAX><<<grid, kLossThreads, 0, as_stream(stream)>>>(
                           logits, labels, class_w, c, s, ignore_index, dice_softmax, acc););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_fwd(const float* logits, const int32_t* labels, const float* class_w, int n, int c, int64_t s,
                    int ignore_index, double* acc, void* stream) {
  return msb_dice_ce_fwd_ex(logits, labels, class_w, n, c, s, ignore_index, 0, acc, stream);
}

int msb_dice_ce_finalize_ex(const double* acc, int c, const float* dice_w, float* result, void* stream) {
  MSB_REQUIRE(acc && result && c > 0 && c <= 32, "msb_dice_ce_finalize: bad arguments");
  dice_ce_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(acc, c, dice_w, result);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_finalize(const double* acc, int c, float* result, void* stream) {
  return msb_dice_ce_finalize_ex(acc, c, nullptr, result, stream);
}

int msb_dice_ce_bwd_ex(const float* logits, const int32_t* labels, const float* class_w, const double* acc, int n,
                       int c, int64_t s, int ignore_index, float coef_ce, float coef_dice, const float* coef_dev,
                       const float* dice_w, int dice_softmax, float* dlogits, void* stream) {
  MSB_REQUIRE(logits && labels && class_w && acc && dlogits && n > 0 && c > 0 && c <= 32 && s > 0,
              "msb_dice_ce_bwd: needs 1 <= C <= 32");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  MSB_DISPATCH_CMAX(c, dice_ce_bwd_kernel<CMAX><<<grid, kLossThreads, 0, as_stream(stream)>>>(
                           logits, labels, class_w, acc, c, s, ignore_index, coef_ce, coef_dice, coef_dev, dice_w,
                           dice_softmax, dlogits););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dice_ce_bwd(const float* logits, const int32_t* labels, const float* class_w, const double* acc, int n, int c,
                    int64_t s, int ignore_index, float coef_ce, float coef_dice, const float* coef_dev, float* dlogits,
                    void* stream) {
  return msb_dice_ce_bwd_ex(logits, labels, class_w, acc, n, c, s, ignore_index, coef_ce, coef_dice, coef_dev, nullptr,
                            0, dlogits, stream);
}

int msb_eval_head(msb_tensor a, const float* w, const float* b, const int32_t* labels, const float* class_w, int n,
                  int c, int64_t s, int ignore_index, int32_t* pred, double* acc, double* psum, void* stream) {
  MSB_REQUIRE(view_ok(a) && w && n > 0 && c > 0 && c <= 32 && c <= a.c && s > 0 &&
                  (a.c == 8 || a.c == 16 || a.c == 32),
              "msb_eval_head: needs 1 <= C <= 32 and an 8-, 16- or 32-channel B8 view with C <= channels");
  MSB_REQUIRE((pred != nullptr || acc != nullptr) != (psum != nullptr),
              "msb_eval_head: request either pred / acc, or psum (the class-weight pass), not both or neither");
  MSB_REQUIRE(acc == nullptr || (labels != nullptr && class_w != nullptr),
              "msb_eval_head: the loss sums need labels and class weights");
  const dim3 grid((unsigned)((s + kLossVoxPerBlock - 1) / kLossVoxPerBlock), (unsigned)n);
  cudaStream_t st = as_stream(stream);
  const int mode = psum != nullptr ? 2 : (acc != nullptr ? 1 : 0);
#define MSB_EVAL_HEAD_M(CI8_, CMAX_, MODE_)                                                                        \
  eval_head_kernel<T, CI8_, CMAX_, MODE_><<<grid, kLossThreads, 0, st>>>(a, w, b, labels, class_w, c, s,           \
                                                                          ignore_index, pred, acc, psum)
#define MSB_EVAL_HEAD(CI8_, CMAX_)                      \
  do {                                                  \
    if (mode == 0) MSB_EVAL_HEAD_M(CI8_, CMAX_, 0);     \
    else if (mode == 1) MSB_EVAL_HEAD_M(CI8_, CMAX_, 1); \
    else MSB_EVAL_HEAD_M(CI8_, CMAX_, 2);               \
  } while (0)
  MSB_DISPATCH_DTYPE(a.dtype, {
    if (a.c == 8) { if (c <= 4) MSB_EVAL_HEAD(1, 4); else MSB_EVAL_HEAD(1, 8); }
    else if (a.c == 16) { if (c <= 4) MSB_EVAL_HEAD(2, 4); else if (c <= 8) MSB_EVAL_HEAD(2, 8); else MSB_EVAL_HEAD(2, 16); }
    else { if (c <= 20) MSB_EVAL_HEAD(4, 20); else MSB_EVAL_HEAD(4, 32); }
  });
#undef MSB_EVAL_HEAD
#undef MSB_EVAL_HEAD_M
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_argmax_channels(const float* logits, int n, int c, int64_t s, int32_t* pred, void* stream) {
  MSB_REQUIRE(logits && pred && n > 0 && c > 0 && s > 0, "msb_argmax_channels: bad arguments");
  int64_t blocks = (s + 255) / 256;
  if (blocks > 8 * kNumSMs) blocks = 8 * kNumSMs;
  argmax_channels_kernel<<<dim3((unsigned)blocks, (unsigned)n), 256, 0, as_stream(stream)>>>(logits, c, s, pred);
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_dropout_masks(uint64_t seed, uint64_t* step_counter, float* out, int total, float p, void* stream) {
  MSB_REQUIRE(step_counter && out && total > 0 && p >= 0.f && p < 1.f, "msb_dropout_masks: bad arguments");
  dropout_masks_kernel<<<1, 1024, 0, as_stream(stream)>>>(seed, reinterpret_cast<unsigned long long*>(step_counter), out,
                                                          total, p, 1.f / (1.f - p));
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
