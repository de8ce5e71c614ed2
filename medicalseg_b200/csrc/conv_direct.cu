// Direct (CUDA-core, f32 accumulate) convolutions for the HBM-leaning layers of VNet:
//   * InputTransition.conv1  1->16, 5x5x5, pad 2            (vnet.py:67-68)   fwd + wgrad (no dgrad: image)
//   * DownTransition.down_conv  k=kernel, s=stride, pad 0     (vnet.py:98-99)   fwd / dgrad / wgrad
//   * UpTransition.up_conv  Conv3DTranspose k, s             (vnet.py:133-137) fwd / dgrad / wgrad
// These are < 1.5 % of the step FLOPs; the 5x5x5 C->C layers run on tcgen05
// (conv_k5_umma.cu).  Weights and weight-gradients stay f32 in the reference (Paddle) layouts.
#include "common.cuh"

namespace msb {

// =====================================================================================================
// in_tr: 1 -> 16 channels, 5x5x5, pad 2.  Block tile 32(w) x 8(h) x 4(d); thread = (w,h) column of 4 d.
// =====================================================================================================
constexpr int kInTW = 32, kInTH = 8, kInTD = 4;
constexpr int kInHW = kInTW + 4, kInHH = kInTH + 4, kInHD = kInTD + 4;

__device__ __forceinline__ void load_in_halo(const float* __restrict__ x, int n, msb_dim3 dims, int d0, int h0,
                                             int w0, float* halo) {
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  for (int i = threadIdx.x; i < kInHD * kInHH * kInHW; i += blockDim.x) {
    const int hw = i % kInHW, hh = (i / kInHW) % kInHH, hd = i / (kInHW * kInHH);
    const int d = d0 + hd - 2, h = h0 + hh - 2, w = w0 + hw - 2;
    float v = 0.f;
    if (d >= 0 && d < dims.d && h >= 0 && h < dims.h && w >= 0 && w < dims.w)
      v = __ldg(x + (int64_t)n * s + ((int64_t)d * dims.h + h) * dims.w + w);
    halo[i] = v;
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    conv_in_fwd_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                       msb_tensor out, msb_dim3 dims, int tiles_w, int tiles_h, int groups,
                       double* __restrict__ sums) {
  __shared__ float halo[kInHD * kInHH * kInHW];
  __shared__ __align__(16) float ws[125 * 16];
  __shared__ float red[8][32];
  const int n = blockIdx.z;
  const int tw = blockIdx.x % tiles_w, th = blockIdx.x / tiles_w;
  const int w0 = tw * kInTW, h0 = th * kInTH, d0 = blockIdx.y * kInTD;
  for (int i = threadIdx.x; i < 125 * 16; i += 256) {
    const int co = i & 15, tap = i >> 4;
    ws[i] = w[co * 125 + tap];  // [16][1][5][5][5] -> [tap][co]
  }
  load_in_halo(x, n, dims, d0, h0, w0, halo);
  __syncthreads();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  float acc[kInTD][16];
#pragma unroll
  for (int i = 0; i < kInTD; ++i)
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[i][c] = bias ? __ldg(bias + c) : 0.f;
  for (int kh = 0; kh < 5; ++kh) {
#pragma unroll
    for (int kw = 0; kw < 5; ++kw) {
      float col[kInHD];
#pragma unroll
      for (int z = 0; z < kInHD; ++z) col[z] = halo[(z * kInHH + ty + kh) * kInHW + tx + kw];
#pragma unroll
      for (int kd = 0; kd < 5; ++kd) {
        const float4* wp = reinterpret_cast<const float4*>(ws + ((kd * 5 + kh) * 5 + kw) * 16);
        const float4 w0v = wp[0], w1v = wp[1], w2v = wp[2], w3v = wp[3];
        const float wv[16] = {w0v.x, w0v.y, w0v.z, w0v.w, w1v.x, w1v.y, w1v.z, w1v.w,
                              w2v.x, w2v.y, w2v.z, w2v.w, w3v.x, w3v.y, w3v.z, w3v.w};
#pragma unroll
        for (int i = 0; i < kInTD; ++i)
#pragma unroll
          for (int c = 0; c < 16; ++c) acc[i][c] = fmaf(col[i + kd], wv[c], acc[i][c]);
      }
    }
  }
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  const int h = h0 + ty, wq = w0 + tx;
  float st[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) st[i] = 0.f;
#pragma unroll
  for (int i = 0; i < kInTD; ++i) {
    const int d = d0 + i;
    if (d < dims.d && h < dims.h && wq < dims.w) {
      const int64_t v = ((int64_t)d * dims.h + h) * dims.w + wq;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float o[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          o[j] = Vec8<T>::round(acc[i][k * 8 + j]);
          st[k * 8 + j] += o[j];
          st[16 + k * 8 + j] += o[j] * o[j];
        }
        Vec8<T>::store(view_ptr<T>(out, n, k, s, v), o);
      }
    }
  }
  if (sums != nullptr) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const float r = warp_sum(st[i]);
      if (lane == 0) red[warp][i] = r;
    }
    __syncthreads();
    if (threadIdx.x < 32) {
      double t = 0;
#pragma unroll
      for (int wv = 0; wv < 8; ++wv) t += (double)red[wv][threadIdx.x];
      const int g = groups == 1 ? 0 : n;
      const int stat = threadIdx.x >> 4, c = threadIdx.x & 15;
      atomicAdd(&sums[((int64_t)stat * groups + g) * 16 + c], t);
    }
  }
}

// wgrad: thread t < 250 owns (tap = t % 125, co half = t / 125); threads 250,251 own the bias halves.
template <typename T>
__global__ void __launch_bounds__(256)
    conv_in_wgrad_kernel(const float* __restrict__ x, msb_tensor dy, float* __restrict__ dw,
                         float* __restrict__ dbias, int nbatch, msb_dim3 dims, int tiles_w, int tiles_h, int tiles_d) {
  __shared__ float halo[kInHD * kInHH * kInHW];
  __shared__ __align__(16) float dys[kInTH * kInTW][16];  // one d-slice (256 voxels) of dy at a time, 16 KB
  const int t = threadIdx.x;
  const int tap = t % 125, half = (t / 125) & 1;
  const bool is_w = t < 250, is_b = (t >= 250 && t < 252);
  const int bhalf = t - 250;
  const int kd = tap / 25, kh = (tap / 5) % 5, kw = tap % 5;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  const int64_t s = (int64_t)dims.d * dims.h * dims.w;
  const int64_t ntiles = (int64_t)nbatch * tiles_d * tiles_h * tiles_w;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int r = (int)(tile % ((int64_t)tiles_d * tiles_h * tiles_w));
    const int n = (int)(tile / ((int64_t)tiles_d * tiles_h * tiles_w));
    const int tw = r % tiles_w; r /= tiles_w;
    const int th = r % tiles_h; const int td = r / tiles_h;
    const int w0 = tw * kInTW, h0 = th * kInTH, d0 = td * kInTD;
    __syncthreads();
    load_in_halo(x, n, dims, d0, h0, w0, halo);
    for (int dz = 0; dz < kInTD; ++dz) {
      __syncthreads();
      {  // stage dy slice: thread = voxel (tx, ty)
        const int tx = t & 31, ty = t >> 5;
        const int d = d0 + dz, h = h0 + ty, wq = w0 + tx;
        float a[8], b[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) a[j] = b[j] = 0.f;
        if (d < dims.d && h < dims.h && wq < dims.w) {
          const int64_t v = ((int64_t)d * dims.h + h) * dims.w + wq;
          Vec8<T>::load(view_ptr<T>(dy, n, 0, s, v), a);
          Vec8<T>::load(view_ptr<T>(dy, n, 1, s, v), b);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) { dys[t][j] = a[j]; dys[t][8 + j] = b[j]; }
      }
      __syncthreads();
      if (is_w) {
        const float* hp = halo + ((dz + kd) * kInHH + kh) * kInHW + kw;
        for (int vy = 0; vy < kInTH; ++vy) {
#pragma unroll 8
          for (int vx = 0; vx < kInTW; ++vx) {
            const float xv = hp[vy * kInHW + vx];
            const float4 g0 = *reinterpret_cast<const float4*>(&dys[vy * kInTW + vx][half * 8]);
            const float4 g1 = *reinterpret_cast<const float4*>(&dys[vy * kInTW + vx][half * 8 + 4]);
            acc[0] = fmaf(xv, g0.x, acc[0]); acc[1] = fmaf(xv, g0.y, acc[1]);
            acc[2] = fmaf(xv, g0.z, acc[2]); acc[3] = fmaf(xv, g0.w, acc[3]);
            acc[4] = fmaf(xv, g1.x, acc[4]); acc[5] = fmaf(xv, g1.y, acc[5]);
            acc[6] = fmaf(xv, g1.z, acc[6]); acc[7] = fmaf(xv, g1.w, acc[7]);
          }
        }
      } else if (is_b) {
        for (int vi = 0; vi < kInTH * kInTW; ++vi)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] += dys[vi][bhalf * 8 + j];
      }
    }
  }
  if (is_w) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dw + (half * 8 + j) * 125 + tap, acc[j]);
  } else if (is_b && dbias) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dbias + bhalf * 8 + j, acc[j]);
  }
}

// =====================================================================================================
// generic strided conv (no padding), B8 activations.  Block = (voxel chunk, 8-out-channel plane, n);
// thread = one output voxel x 8 output channels; the weight slice for this plane lives in smem as
// [tap][rc][8 oc] f32.
// =====================================================================================================
constexpr int kSThreads = 128;

struct ConvGeom {
  msb_dim3 in, out, k, s, p;  // "in" = the larger grid (gather source / scatter target), "out" = the smaller one
  int cred, cout;          // reduction / produced channels of the views (multiples of 8)
  int cred_real, cout_real;  // channel counts of the weight tensor (<= the view counts; the rest is zero padding)
  int taps;
};

// weight element for (produced channel oc, reduction channel rc, tap): W_OUT_FIRST -> w[oc][rc][tap]
// (gather form) else w[rc][oc][tap] (scatter form)
// stages the slice [tap][rc in chunk][8 oc] of the weight tensor (reduction channels rc0 .. rc0+rchunk)
template <bool W_OUT_FIRST>
__device__ __forceinline__ void stage_weights(const float* __restrict__ w, const ConvGeom& g, int oc8, int rc0,
                                              int rchunk, float* ws) {
  const int total = g.taps * rchunk * 8;
  for (int i = threadIdx.x; i < total; i += blockDim.x) {
    const int j = i & 7, rl = (i >> 3) % rchunk, tap = i / (8 * rchunk);
    const int oc = oc8 * 8 + j, rc = rc0 + rl;
    const int64_t idx = W_OUT_FIRST ? ((int64_t)oc * g.cred_real + rc) * g.taps + tap
                                    : ((int64_t)rc * g.cout_real + oc) * g.taps + tap;
    ws[i] = (oc < g.cout_real && rc < g.cred_real) ? __ldg(w + idx) : 0.f;
  }
}

// adds pre-accumulated per-thread sums (st) and sums of squares (sq) of 8 channels to the BN accumulators
__device__ __forceinline__ void block_bn_sums2(const float (&st)[8], const float (&sq)[8], int groups, int n,
                                               int c_total, int oc8, double* __restrict__ sums,
                                               float* red /*[warps][16]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float r0 = warp_sum(st[i]), r1 = warp_sum(sq[i]);
    if (lane == 0) { red[warp * 16 + i] = r0; red[warp * 16 + 8 + i] = r1; }
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0;
    for (int wv = 0; wv < (int)(blockDim.x >> 5); ++wv) t += (double)red[wv * 16 + threadIdx.x];
    const int g = groups == 1 ? 0 : n;
    const int stat = threadIdx.x >> 3, j = threadIdx.x & 7;
    atomicAdd(&sums[((int64_t)stat * groups + g) * c_total + oc8 * 8 + j], t);
  }
}

template <typename T>
__device__ __forceinline__ void block_bn_sums(const float (&o)[8], bool valid, int groups, int n, int c_total,
                                              int oc8, double* __restrict__ sums, float* red /*[4][16]*/) {
  float st[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    st[j] = valid ? o[j] : 0.f;
    st[8 + j] = valid ? o[j] * o[j] : 0.f;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float r = warp_sum(st[i]);
    if (lane == 0) red[warp * 16 + i] = r;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0;
    for (int wv = 0; wv < kSThreads / 32; ++wv) t += (double)red[wv * 16 + threadIdx.x];
    const int g = groups == 1 ? 0 : n;
    const int stat = threadIdx.x >> 3, j = threadIdx.x & 7;
    atomicAdd(&sums[((int64_t)stat * groups + g) * c_total + oc8 * 8 + j], t);
  }
}

constexpr int kGVox = 4;  // output voxels per thread: the 8x8 weight block in registers is reused 4x

__device__ __forceinline__ void fma_8x8(const float (&a)[8], const float (&wr)[64], float (&acc)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = fmaf(a[i], wr[i * 8 + j], acc[j]);
}

__device__ __forceinline__ void load_w64(const float* wt, float (&wr)[64]) {
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float4 t = *reinterpret_cast<const float4*>(wt + i * 4);
    wr[i * 4] = t.x; wr[i * 4 + 1] = t.y; wr[i * 4 + 2] = t.z; wr[i * 4 + 3] = t.w;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSThreads)
    conv_gather_kernel(msb_tensor x, const float* __restrict__ w, const float* __restrict__ bias, msb_tensor out,
                       ConvGeom g, int rchunk, int groups, double* __restrict__ sums) {
  extern __shared__ __align__(16) float ws[];
  __shared__ float red[(kSThreads / 32) * 16];
  const int oc8 = blockIdx.y, n = blockIdx.z;
  const int64_t so = (int64_t)g.out.d * g.out.h * g.out.w, si = (int64_t)g.in.d * g.in.h * g.in.w;
  const int64_t vbase = (int64_t)blockIdx.x * (kSThreads * kGVox) + threadIdx.x;
  bool valid[kGVox];
  int od[kGVox], oh[kGVox], ow[kGVox];
  float acc[kGVox][8];
#pragma unroll
  for (int i = 0; i < kGVox; ++i) {
    const int64_t v = vbase + (int64_t)i * kSThreads;
    valid[i] = v < so;
    const int64_t vv = valid[i] ? v : 0;
    ow[i] = (int)(vv % g.out.w);
    oh[i] = (int)((vv / g.out.w) % g.out.h);
    od[i] = (int)(vv / ((int64_t)g.out.w * g.out.h));
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = (bias && oc8 * 8 + j < g.cout_real) ? __ldg(bias + oc8 * 8 + j) : 0.f;
  }
  for (int rc0 = 0; rc0 < g.cred; rc0 += rchunk) {
    __syncthreads();
    stage_weights<true>(w, g, oc8, rc0, rchunk, ws);
    __syncthreads();
    int tap = 0;
    for (int kd = 0; kd < g.k.d; ++kd)
      for (int kh = 0; kh < g.k.h; ++kh)
        for (int kw = 0; kw < g.k.w; ++kw, ++tap) {
          int64_t vi[kGVox];
          bool ok[kGVox];
#pragma unroll
          for (int i = 0; i < kGVox; ++i) {
            const int id = od[i] * g.s.d + kd - g.p.d, ih = oh[i] * g.s.h + kh - g.p.h, iw = ow[i] * g.s.w + kw - g.p.w;
            ok[i] = valid[i] && id >= 0 && id < g.in.d && ih >= 0 && ih < g.in.h && iw >= 0 && iw < g.in.w;
            vi[i] = ok[i] ? ((int64_t)id * g.in.h + ih) * g.in.w + iw : 0;
          }
          const float* wt = ws + (int64_t)tap * rchunk * 8;
          for (int r8 = 0; r8 < rchunk / 8; ++r8) {
            float a[kGVox][8];
#pragma unroll
            for (int i = 0; i < kGVox; ++i) {
              if (ok[i]) Vec8<T>::load(view_ptr<T>(x, n, rc0 / 8 + r8, si, vi[i]), a[i]);
              else {
#pragma unroll
                for (int j = 0; j < 8; ++j) a[i][j] = 0.f;
              }
            }
            float wr[64];
            load_w64(wt + r8 * 64, wr);
#pragma unroll
            for (int i = 0; i < kGVox; ++i) fma_8x8(a[i], wr, acc[i]);
          }
        }
  }
  float st[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) st[j] = 0.f;
  float sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sq[j] = 0.f;
#pragma unroll
  for (int i = 0; i < kGVox; ++i) {
    if (valid[i]) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[i][j] = Vec8<T>::round(acc[i][j]);
        st[j] += acc[i][j];
        sq[j] += acc[i][j] * acc[i][j];
      }
      Vec8<T>::store(view_ptr<T>(out, n, oc8, so, vbase + (int64_t)i * kSThreads), acc[i]);
    }
  }
  if (sums != nullptr) {
    __syncthreads();
    block_bn_sums2(st, sq, groups, n, g.cout, oc8, sums, red);
  }
}

template <typename T>
__global__ void __launch_bounds__(kSThreads)
    conv_scatter_kernel(msb_tensor x, const float* __restrict__ w, const float* __restrict__ bias, msb_tensor out,
                        ConvGeom g, int rchunk, int accumulate, int groups, double* __restrict__ sums) {
  extern __shared__ __align__(16) float ws[];
  __shared__ float red[(kSThreads / 32) * 16];
  const int oc8 = blockIdx.y, n = blockIdx.z;
  // here g.in is the LARGE grid we produce, g.out the small grid we read
  const int64_t sl = (int64_t)g.in.d * g.in.h * g.in.w, ss = (int64_t)g.out.d * g.out.h * g.out.w;
  const int64_t v = (int64_t)blockIdx.x * kSThreads + threadIdx.x;
  const bool valid = v < sl;
  const int iw = (int)(v % g.in.w), ih = (int)((v / g.in.w) % g.in.h), id = (int)(v / ((int64_t)g.in.w * g.in.h));
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = (bias && oc8 * 8 + j < g.cout_real) ? __ldg(bias + oc8 * 8 + j) : 0.f;
  if (valid && accumulate) {
    float prev[8];
    Vec8<T>::load(view_ptr<T>(out, n, oc8, sl, v), prev);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += prev[j];
  }
  for (int rc0 = 0; rc0 < g.cred; rc0 += rchunk) {
    __syncthreads();
    stage_weights<false>(w, g, oc8, rc0, rchunk, ws);
    __syncthreads();
    if (!valid) continue;
    int tap = 0;
    for (int kd = 0; kd < g.k.d; ++kd) {
      const int rd = id - kd + g.p.d;
      for (int kh = 0; kh < g.k.h; ++kh) {
        const int rh = ih - kh + g.p.h;
        for (int kw = 0; kw < g.k.w; ++kw, ++tap) {
          const int rw = iw - kw + g.p.w;
          if (rd < 0 || rh < 0 || rw < 0 || rd % g.s.d || rh % g.s.h || rw % g.s.w) continue;
          const int od = rd / g.s.d, oh = rh / g.s.h, ow = rw / g.s.w;
          if (od >= g.out.d || oh >= g.out.h || ow >= g.out.w) continue;
          const int64_t vs = ((int64_t)od * g.out.h + oh) * g.out.w + ow;
          const float* wt = ws + (int64_t)tap * rchunk * 8;
          for (int r8 = 0; r8 < rchunk / 8; ++r8) {
            float a[8];
            Vec8<T>::load(view_ptr<T>(x, n, rc0 / 8 + r8, ss, vs), a);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 w0 = *reinterpret_cast<const float4*>(wt + (r8 * 8 + i) * 8);
              const float4 w1 = *reinterpret_cast<const float4*>(wt + (r8 * 8 + i) * 8 + 4);
              acc[0] = fmaf(a[i], w0.x, acc[0]); acc[1] = fmaf(a[i], w0.y, acc[1]);
              acc[2] = fmaf(a[i], w0.z, acc[2]); acc[3] = fmaf(a[i], w0.w, acc[3]);
              acc[4] = fmaf(a[i], w1.x, acc[4]); acc[5] = fmaf(a[i], w1.y, acc[5]);
              acc[6] = fmaf(a[i], w1.z, acc[6]); acc[7] = fmaf(a[i], w1.w, acc[7]);
            }
          }
        }
      }
    }
  }
  if (valid) {
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = Vec8<T>::round(acc[j]);
    Vec8<T>::store(view_ptr<T>(out, n, oc8, sl, v), acc);
  }
  if (sums != nullptr) {
    __syncthreads();
    block_bn_sums<T>(acc, valid, groups, n, g.cout, oc8, sums, red);
  }
}

// Non-overlapping fast path of the scatter form (kernel == stride, kernel.w == 2): every output voxel receives
// exactly one tap.  Block = (chunk of small-grid voxels, (8-out-channel plane, kd, kh), n); a thread owns 2 input
// voxels and both kw taps, so it stores 32 contiguous bytes per (voxel, plane) and reuses each 8x8 weight block.
constexpr int kXVox = 2;

template <typename T>
__global__ void __launch_bounds__(kSThreads)
    conv_expand_k2_kernel(msb_tensor x, const float* __restrict__ w, const float* __restrict__ bias, msb_tensor out,
                          ConvGeom g, int accumulate, int groups, double* __restrict__ sums) {
  extern __shared__ __align__(16) float ws[];  // [kw 2][rc][8 oc]
  __shared__ float red[(kSThreads / 32) * 16];
  const int dh = g.k.d * g.k.h;
  const int oc8 = blockIdx.y / dh, kd = (blockIdx.y % dh) / g.k.h, kh = blockIdx.y % g.k.h, n = blockIdx.z;
  for (int i = threadIdx.x; i < 2 * g.cred * 8; i += blockDim.x) {
    const int j = i & 7, rc = (i >> 3) % g.cred, kw = i / (8 * g.cred);
    const int oc = oc8 * 8 + j;
    const int tap = (kd * g.k.h + kh) * 2 + kw;
    ws[i] = (oc < g.cout_real && rc < g.cred_real) ? __ldg(w + ((int64_t)rc * g.cout_real + oc) * g.taps + tap) : 0.f;
  }
  __syncthreads();
  const int64_t sl = (int64_t)g.in.d * g.in.h * g.in.w, ss = (int64_t)g.out.d * g.out.h * g.out.w;
  const int64_t vbase = (int64_t)blockIdx.x * (kSThreads * kXVox) + threadIdx.x;
  bool valid[kXVox];
  int64_t vo[kXVox];
  float acc[kXVox][2][8];
#pragma unroll
  for (int i = 0; i < kXVox; ++i) {
    const int64_t v = vbase + (int64_t)i * kSThreads;
    valid[i] = v < ss;
    const int64_t vv = valid[i] ? v : 0;
    const int ow = (int)(vv % g.out.w), oh = (int)((vv / g.out.w) % g.out.h), od = (int)(vv / ((int64_t)g.out.w * g.out.h));
    vo[i] = ((int64_t)(od * g.s.d + kd) * g.in.h + (oh * g.s.h + kh)) * g.in.w + ow * 2;
#pragma unroll
    for (int k = 0; k < 2; ++k)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][k][j] = (bias && oc8 * 8 + j < g.cout_real) ? __ldg(bias + oc8 * 8 + j) : 0.f;
    if (valid[i] && accumulate) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        float prev[8];
        Vec8<T>::load(view_ptr<T>(out, n, oc8, sl, vo[i] + k), prev);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][k][j] += prev[j];
      }
    }
  }
  for (int r8 = 0; r8 < g.cred / 8; ++r8) {
    float a[kXVox][8];
#pragma unroll
    for (int i = 0; i < kXVox; ++i) {
      if (valid[i]) Vec8<T>::load(view_ptr<T>(x, n, r8, ss, vbase + (int64_t)i * kSThreads), a[i]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) a[i][j] = 0.f;
      }
    }
#pragma unroll
    for (int k = 0; k < 2; ++k) {
      float wr[64];
      load_w64(ws + ((int64_t)k * g.cred + r8 * 8) * 8, wr);
#pragma unroll
      for (int i = 0; i < kXVox; ++i) fma_8x8(a[i], wr, acc[i][k]);
    }
  }
  float st[8], sq[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) st[j] = sq[j] = 0.f;
#pragma unroll
  for (int i = 0; i < kXVox; ++i) {
    if (valid[i]) {
#pragma unroll
      for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[i][k][j] = Vec8<T>::round(acc[i][k][j]);
          st[j] += acc[i][k][j];
          sq[j] += acc[i][k][j] * acc[i][k][j];
        }
        Vec8<T>::store(view_ptr<T>(out, n, oc8, sl, vo[i] + k), acc[i][k]);
      }
    }
  }
  if (sums != nullptr) {
    __syncthreads();
    block_bn_sums2(st, sq, groups, n, g.cout, oc8, sums, red);
  }
}

// wgrad: block = (small-grid voxel chunk, (bc8, sc8) pair, n); loops taps; 8x8 outer products per voxel.
constexpr int kWThreads = 256;
constexpr int kWVoxPerBlock = 8192;

template <typename T>
__global__ void __launch_bounds__(kWThreads)
    conv_strided_wgrad_kernel(msb_tensor big, msb_tensor small, float* __restrict__ dw, ConvGeom g) {
  __shared__ float red[kWThreads / 32][64];
  const int bc8n = g.cred / 8;
  const int bc8 = blockIdx.y % bc8n, sc8 = blockIdx.y / bc8n, n = blockIdx.z;
  const int64_t sb = (int64_t)g.in.d * g.in.h * g.in.w, ss = (int64_t)g.out.d * g.out.h * g.out.w;
  const int64_t v0 = (int64_t)blockIdx.x * kWVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kWVoxPerBlock, ss);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int tap = 0;
  for (int kd = 0; kd < g.k.d; ++kd)
    for (int kh = 0; kh < g.k.h; ++kh)
      for (int kw = 0; kw < g.k.w; ++kw, ++tap) {
        float acc[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) acc[i] = 0.f;
        for (int64_t vb0 = v0 + threadIdx.x; vb0 < v1; vb0 += 4 * kWThreads) {
          float a[4][8], b[4][8];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int64_t v = vb0 + (int64_t)u * kWThreads;
            bool ok = v < v1;
            int64_t vb = 0;
            if (ok) {
              const int ow = (int)(v % g.out.w), oh = (int)((v / g.out.w) % g.out.h),
                        od = (int)(v / ((int64_t)g.out.w * g.out.h));
              const int bd = od * g.s.d + kd - g.p.d, bh = oh * g.s.h + kh - g.p.h, bw = ow * g.s.w + kw - g.p.w;
              ok = bd >= 0 && bd < g.in.d && bh >= 0 && bh < g.in.h && bw >= 0 && bw < g.in.w;
              vb = ((int64_t)bd * g.in.h + bh) * g.in.w + bw;
            }
            if (ok) {
              Vec8<T>::load(view_ptr<T>(big, n, bc8, sb, vb), a[u]);
              Vec8<T>::load(view_ptr<T>(small, n, sc8, ss, v), b[u]);
            } else {
#pragma unroll
              for (int j = 0; j < 8; ++j) a[u][j] = b[u][j] = 0.f;
            }
          }
#pragma unroll
          for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
              for (int j = 0; j < 8; ++j) acc[i * 8 + j] = fmaf(b[u][i], a[u][j], acc[i * 8 + j]);  // [sc][bc]
        }
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          const float r = warp_sum(acc[i]);
          if (lane == 0) red[warp][i] = r;
        }
        __syncthreads();
        if (threadIdx.x < 64) {
          float t = 0.f;
#pragma unroll
          for (int wv = 0; wv < kWThreads / 32; ++wv) t += red[wv][threadIdx.x];
          const int sc = sc8 * 8 + (threadIdx.x >> 3), bc = bc8 * 8 + (threadIdx.x & 7);
          if (sc < g.cout_real && bc < g.cred_real) atomicAdd(dw + ((int64_t)sc * g.cred_real + bc) * g.taps + tap, t);
        }
        __syncthreads();
      }
}

// per-channel sum of a B8 tensor, added to out[C] (f32): bias gradients
template <typename T>
__global__ void __launch_bounds__(256) channel_sum_kernel(msb_tensor x, int64_t s, int c_real,
                                                          float* __restrict__ out) {
  __shared__ float red[8][8];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kWVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kWVoxPerBlock, s);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {
    float a[8];
    Vec8<T>::load(view_ptr<T>(x, n, c8, s, v), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] += a[j];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float r = warp_sum(acc[j]);
    if (lane == 0) red[warp][j] = r;
  }
  __syncthreads();
  if (threadIdx.x < 8 && c8 * 8 + threadIdx.x < c_real) {
    float t = 0.f;
#pragma unroll
    for (int wv = 0; wv < 8; ++wv) t += red[wv][threadIdx.x];
    atomicAdd(out + c8 * 8 + threadIdx.x, t);
  }
}

static inline bool dims_ok(msb_dim3 d) { return d.d > 0 && d.h > 0 && d.w > 0; }

}  // namespace msb

using namespace msb;

extern "C" {

int msb_conv_in_fwd(const float* x, const float* w, const float* bias, msb_tensor out, int n, msb_dim3 dims,
                    int groups, double* sums, void* stream) {
  MSB_REQUIRE(x && w && view_ok(out) && out.c == 16 && n > 0 && dims_ok(dims), "msb_conv_in_fwd: out must have 16 channels");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_conv_in_fwd: groups must be 1 or n");
  const int tw = (dims.w + kInTW - 1) / kInTW, th = (dims.h + kInTH - 1) / kInTH, td = (dims.d + kInTD - 1) / kInTD;
  MSB_REQUIRE(td <= 65535 && n <= 65535, "msb_conv_in_fwd: volume too deep");
  const dim3 grid(tw * th, td, n);
  MSB_DISPATCH_DTYPE(out.dtype, conv_in_fwd_kernel<T><<<grid, 256, 0, as_stream(stream)>>>(x, w, bias, out, dims, tw,
                                                                                             th, groups, sums););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_conv_in_wgrad(const float* x, msb_tensor dy, float* dw, float* dbias, int n, msb_dim3 dims, void* stream) {
  MSB_REQUIRE(x && dw && view_ok(dy) && dy.c == 16 && n > 0 && dims_ok(dims), "msb_conv_in_wgrad: dy must have 16 channels");
  const int tw = (dims.w + kInTW - 1) / kInTW, th = (dims.h + kInTH - 1) / kInTH, td = (dims.d + kInTD - 1) / kInTD;
  const int64_t ntiles = (int64_t)n * tw * th * td;
  const int blocks = (int)(ntiles < 2 * kNumSMs ? ntiles : 2 * kNumSMs);
  MSB_DISPATCH_DTYPE(dy.dtype, conv_in_wgrad_kernel<T><<<blocks, 256, 0, as_stream(stream)>>>(x, dy, dw, dbias, n,
                                                                                                 dims, tw, th, td););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

static int make_geom(msb_dim3 big, msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int cred, int cout, int cred_real,
                     int cout_real, ConvGeom* g) {
  MSB_REQUIRE(cred_real >= 0 && cred_real <= cred && cout_real >= 0 && cout_real <= cout,
              "strided conv: real channel counts exceed the views");
  g->cred_real = cred_real > 0 ? cred_real : cred;
  g->cout_real = cout_real > 0 ? cout_real : cout;
  MSB_REQUIRE(dims_ok(big) && dims_ok(kernel) && dims_ok(stride) && pad.d >= 0 && pad.h >= 0 && pad.w >= 0,
              "strided conv: bad dims");
  MSB_REQUIRE(big.d + 2 * pad.d >= kernel.d && big.h + 2 * pad.h >= kernel.h && big.w + 2 * pad.w >= kernel.w,
              "strided conv: kernel larger than padded input");
  g->in = big;
  g->k = kernel;
  g->s = stride;
  g->p = pad;
  g->out.d = (big.d + 2 * pad.d - kernel.d) / stride.d + 1;
  g->out.h = (big.h + 2 * pad.h - kernel.h) / stride.h + 1;
  g->out.w = (big.w + 2 * pad.w - kernel.w) / stride.w + 1;
  g->cred = cred;
  g->cout = cout;
  g->taps = kernel.d * kernel.h * kernel.w;
  return MSB_OK;
}

// largest reduction-channel chunk (multiple of 8, divides cred) whose [taps][chunk][8] f32 slice fits 64 KB
static int pick_rchunk(int taps, int cred) {
  int best = 8;
  for (int c = 8; c <= cred; c += 8)
    if (cred % c == 0 && (size_t)taps * c * 8 * sizeof(float) <= 64 * 1024) best = c;
  return best;
}

int msb_conv_strided_fwd(msb_tensor x, const float* w, const float* bias, msb_tensor out, int n, msb_dim3 in_dims,
                         msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_red_real, int c_out_real, int groups,
                         double* sums, void* stream) {
  MSB_REQUIRE(view_ok(x) && view_ok(out) && x.dtype == out.dtype && w && n > 0, "msb_conv_strided_fwd: bad arguments");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_conv_strided_fwd: groups must be 1 or n");
  ConvGeom g;
  int rc = make_geom(in_dims, kernel, stride, pad, x.c, out.c, c_red_real, c_out_real, &g);
  if (rc) return rc;
  const int rchunk = pick_rchunk(g.taps, g.cred);
  const size_t smem = (size_t)g.taps * rchunk * 8 * sizeof(float);
  MSB_REQUIRE(smem <= 200 * 1024, "msb_conv_strided_fwd: kernel volume too large for shared memory");
  const int64_t so = (int64_t)g.out.d * g.out.h * g.out.w;
  const dim3 grid((unsigned)((so + kSThreads * kGVox - 1) / (kSThreads * kGVox)), out.c / 8, n);
  cudaStream_t st = as_stream(stream);
  MSB_DISPATCH_DTYPE(x.dtype, {
    if (smem > 48 * 1024)
      MSB_CUDA_OK(cudaFuncSetAttribute(conv_gather_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_gather_kernel<T><<<grid, kSThreads, smem, st>>>(x, w, bias, out, g, rchunk, groups, sums);
  });
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_conv_strided_bwd_data(msb_tensor x, const float* w, const float* bias, msb_tensor out, int n,
                              msb_dim3 out_dims, msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_red_real,
                              int c_out_real, int accumulate, int groups, double* sums, void* stream) {
  MSB_REQUIRE(view_ok(x) && view_ok(out) && x.dtype == out.dtype && w && n > 0,
              "msb_conv_strided_bwd_data: bad arguments");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_conv_strided_bwd_data: groups must be 1 or n");
  ConvGeom g;
  int rc = make_geom(out_dims, kernel, stride, pad, x.c, out.c, c_red_real, c_out_real, &g);
  if (rc) return rc;
  cudaStream_t st = as_stream(stream);
  const bool fast = kernel.d == stride.d && kernel.h == stride.h && kernel.w == stride.w && kernel.w == 2 &&
                    pad.d == 0 && pad.h == 0 && pad.w == 0 && out_dims.d == g.out.d * stride.d &&
                    out_dims.h == g.out.h * stride.h && out_dims.w == g.out.w * 2 &&
                    (size_t)2 * g.cred * 8 * sizeof(float) <= 48 * 1024;
  if (fast) {
    const int64_t ss = (int64_t)g.out.d * g.out.h * g.out.w;
    const dim3 fgrid((unsigned)((ss + kSThreads * kXVox - 1) / (kSThreads * kXVox)),
                     (out.c / 8) * kernel.d * kernel.h, n);
    const size_t fsmem = (size_t)2 * g.cred * 8 * sizeof(float);
    MSB_DISPATCH_DTYPE(x.dtype, conv_expand_k2_kernel<T><<<fgrid, kSThreads, fsmem, st>>>(x, w, bias, out, g,
                                                                                          accumulate, groups, sums););
    MSB_LAUNCH_OK();
    return MSB_OK;
  }
  const int rchunk = pick_rchunk(g.taps, g.cred);
  const size_t smem = (size_t)g.taps * rchunk * 8 * sizeof(float);
  MSB_REQUIRE(smem <= 200 * 1024, "msb_conv_strided_bwd_data: kernel volume too large for shared memory");
  const int64_t sl = (int64_t)g.in.d * g.in.h * g.in.w;
  const dim3 grid((unsigned)((sl + kSThreads - 1) / kSThreads), out.c / 8, n);
  MSB_DISPATCH_DTYPE(x.dtype, {
    if (smem > 48 * 1024)
      MSB_CUDA_OK(cudaFuncSetAttribute(conv_scatter_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    conv_scatter_kernel<T><<<grid, kSThreads, smem, st>>>(x, w, bias, out, g, rchunk, accumulate, groups, sums);
  });
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_conv_strided_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n, msb_dim3 big_dims,
                           msb_dim3 kernel, msb_dim3 stride, msb_dim3 pad, int c_big_real, int c_small_real,
                           int bias_from_big, void* stream) {
  MSB_REQUIRE(view_ok(big) && view_ok(small) && big.dtype == small.dtype && dw && n > 0,
              "msb_conv_strided_wgrad: bad arguments");
  ConvGeom g;
  int rc = make_geom(big_dims, kernel, stride, pad, big.c, small.c, c_big_real, c_small_real, &g);
  if (rc) return rc;
  const int64_t ss = (int64_t)g.out.d * g.out.h * g.out.w;
  const int pairs = (big.c / 8) * (small.c / 8);
  MSB_REQUIRE(pairs <= 65535, "msb_conv_strided_wgrad: too many channel-plane pairs");
  const dim3 grid((unsigned)((ss + kWVoxPerBlock - 1) / kWVoxPerBlock), pairs, n);
  cudaStream_t st = as_stream(stream);
  MSB_DISPATCH_DTYPE(big.dtype, {
    conv_strided_wgrad_kernel<T><<<grid, kWThreads, 0, st>>>(big, small, dw, g);
    if (dbias != nullptr) {
      const msb_tensor& bt = bias_from_big ? big : small;
      const int64_t sbt = bias_from_big ? (int64_t)g.in.d * g.in.h * g.in.w : ss;
      const dim3 bgrid((unsigned)((sbt + kWVoxPerBlock - 1) / kWVoxPerBlock), bt.c / 8, n);
      channel_sum_kernel<T><<<bgrid, 256, 0, st>>>(bt, sbt, bias_from_big ? g.cred_real : g.cout_real, dbias);
    }
  });
  MSB_LAUNCH_OK();
  return MSB_OK;
}

}  // extern "C"
