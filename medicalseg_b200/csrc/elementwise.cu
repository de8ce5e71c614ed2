// HBM-bound elementwise / reduction kernels of the VNet path (B8 layout), sm_100a.
// One block owns a fixed (n, 8-channel plane) and a chunk of voxels, so per-channel parameters sit in
// registers and per-channel reductions finish with one warp-shuffle tree + one double atomic per block.
#include <stdarg.h>

#include "common.cuh"
#include "umma.cuh"

namespace msb {

int g_pdl_enabled = 1;
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

constexpr int kThreads = 256;
constexpr int kVoxPerBlock = 2048;

static inline dim3 plane_grid(int n, int c, int64_t s) {
  return dim3((unsigned)((s + kVoxPerBlock - 1) / kVoxPerBlock), (unsigned)(c / 8), (unsigned)n);
}
// block-level reduction of K per-thread floats; result valid in threads [0, K) of warp 0.. returned via smem
template <int K>
__device__ __forceinline__ void block_reduce_to_smem(float (&v)[K], float* smem /*[kThreads/32][K]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < K; ++i) {
    float r = warp_sum(v[i]);
    if (lane == 0) smem[warp * K + i] = r;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) to_blocked_kernel(const float* __restrict__ src, int c, int64_t s,
                                                               msb_tensor dst) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ch = c8 * 8 + j;
      x[j] = ch < c ? __ldg(src + ((int64_t)n * c + ch) * s + v) : 0.f;
    }
    Vec8<T>::store(view_ptr<T>(dst, n, c8, s, v), x);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads) from_blocked_kernel(msb_tensor src, float* __restrict__ dst, int c,
                                                                 int64_t s) {
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float x[8];
    Vec8<T>::load(view_ptr<T>(src, n, c8, s, v), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      int ch = c8 * 8 + j;
      if (ch < c) dst[((int64_t)n * c + ch) * s + v] = x[j];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) bn_stats_kernel(msb_tensor x, int64_t s, int groups, int c_total,
                                                            double* __restrict__ sums) {
  __shared__ float red[kThreads / 32][16];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float a[8];
    Vec8<T>::load(view_ptr<T>(x, n, c8, s, v), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += a[j];
      acc[8 + j] += a[j] * a[j];
    }
  }
  block_reduce_to_smem<16>(acc, &red[0][0]);
  if (threadIdx.x < 16) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kThreads / 32; ++w) t += (double)red[w][threadIdx.x];
    const int g = groups == 1 ? 0 : n;
    const int stat = threadIdx.x >> 3, j = threadIdx.x & 7;
    atomicAdd(&sums[((int64_t)stat * groups + g) * c_total + c8 * 8 + j], t);
  }
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ rmean,
                                   float* __restrict__ rvar, float momentum, float eps, int training, int c,
                                   int groups, float* __restrict__ bnbuf) {
  pdl_wait();
  pdl_trigger();
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  const int64_t gc = (int64_t)groups * c;
  double mean_acc = 0, var_acc = 0;
  for (int g = 0; g < groups; ++g) {
    double mean, var;
    if (training) {
      mean = sums[(int64_t)g * c + ch] / count;
      var = sums[gc + (int64_t)g * c + ch] / count - mean * mean;
      if (var < 0) var = 0;
      mean_acc += mean;
      var_acc += var;
    } else {
      mean = rmean[ch];
      var = rvar[ch];
    }
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float scale = gamma[ch] * invstd;
    bnbuf[0 * gc + (int64_t)g * c + ch] = scale;
    bnbuf[1 * gc + (int64_t)g * c + ch] = beta[ch] - (float)mean * scale;
    bnbuf[2 * gc + (int64_t)g * c + ch] = (float)mean;
    bnbuf[3 * gc + (int64_t)g * c + ch] = invstd;
  }
  if (training && rmean != nullptr) {
    rmean[ch] = momentum * rmean[ch] + (1.f - momentum) * (float)(mean_acc / groups);
    rvar[ch] = momentum * rvar[ch] + (1.f - momentum) * (float)(var_acc / groups);
  }
}

// ---------------------------------------------------------------------------------------------------
// TMA-staged plane streaming for the HBM-bound BatchNorm kernels.  A (n, 8-channel plane) of a B8 tensor is one
// contiguous run of S * 8 elements, so a chunk of voxels is a single 1-D bulk copy (cp.async.bulk + mbarrier).  One
// block per SM keeps kStreamStages chunks of every input tensor in flight in shared memory (up to 192 KB), which
// decouples the bytes in flight from the register file: the register-staged version needed 128 registers per thread
// for two voxels in flight and ran at 25 % occupancy / 49-54 % of the HBM peak (ncu, profiles/r1n_*).
constexpr int kStreamStages = 3;
constexpr int kStreamChunkBytes = 16384;  // per tensor and stage: 1024 bf16 voxels (x 8 channels) or 512 f32 voxels
constexpr int kStreamThreads = 512;

template <typename T, int NIN>
struct PlaneStream {
  static constexpr int kVox = kStreamChunkBytes / (8 * (int)sizeof(T));
  static constexpr int kSmemBytes = kStreamStages * NIN * kStreamChunkBytes + 64;
  uint8_t* data;      // [stage][NIN][kStreamChunkBytes]
  uint32_t bar0;      // full barriers [stage]
  const T* src[NIN];  // plane base pointers (element 0 of voxel 0)
  int64_t s, nchunks;
  bool on[NIN];       // inputs that are really streamed (others are skipped)

  __device__ __forceinline__ void init(uint8_t* smem_raw) {
    uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 15) & ~uintptr_t(15));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base);
    data = base + 64;
    bar0 = ptx::smem_u32(bars);
    if (threadIdx.x == 0) {
      for (int i = 0; i < kStreamStages; ++i) ptx::mbar_init(bar0 + 8u * i, 1);
      ptx::fence_mbar_init();
    }
    __syncthreads();
  }
  __device__ __forceinline__ int chunk_voxels(int64_t chunk) const {
    const int64_t v0 = chunk * kVox;
    return (int)min((int64_t)kVox, s - v0);
  }
  // thread 0 only
  __device__ __forceinline__ void issue(int64_t chunk, int slot) {
    const int nv = chunk_voxels(chunk);
    const uint32_t bytes = (uint32_t)nv * 8u * (uint32_t)sizeof(T);
    int active = 0;
#pragma unroll
    for (int i = 0; i < NIN; ++i) active += on[i] ? 1 : 0;
    ptx::mbar_expect_tx(bar0 + 8u * slot, bytes * active);
#pragma unroll
    for (int i = 0; i < NIN; ++i)
      if (on[i])
        ptx::bulk_load(ptx::smem_u32(data + ((size_t)slot * NIN + i) * kStreamChunkBytes), src[i] + chunk * kVox * 8, bytes,
                       bar0 + 8u * slot);
  }
  __device__ __forceinline__ void prologue() {
    if (threadIdx.x == 0) {
      int slot = 0;
      for (int64_t c = blockIdx.x; c < nchunks && slot < kStreamStages; c += gridDim.x, ++slot) issue(c, slot);
    }
  }
  __device__ __forceinline__ void wait(int64_t k) {  // k = index of the chunk in this block's sequence
    ptx::mbar_wait(bar0 + 8u * (uint32_t)(k % kStreamStages), (uint32_t)((k / kStreamStages) & 1));
  }
  __device__ __forceinline__ const T* stage(int64_t k, int i) const {
    return reinterpret_cast<const T*>(data + ((size_t)(k % kStreamStages) * NIN + i) * kStreamChunkBytes);
  }
  // all threads finished reading chunk k of this block's sequence: refill its slot
  __device__ __forceinline__ void release(int64_t k, int64_t chunk) {
    __syncthreads();
    if (threadIdx.x == 0) {
      const int64_t next = chunk + (int64_t)kStreamStages * gridDim.x;
      if (next < nchunks) issue(next, (int)(k % kStreamStages));
    }
  }
};

template <typename T>
__device__ __forceinline__ const T* plane_base(const msb_tensor& t, int n, int c8, int64_t s) {
  return reinterpret_cast<const T*>(t.ptr) + (int64_t)n * t.n_stride + (int64_t)c8 * s * 8;
}

// grid for the streaming kernels: about one block per SM, at most one block per chunk
template <typename T>
static inline dim3 stream_grid(int n, int c, int64_t s) {
  const int kvox = kStreamChunkBytes / (8 * (int)sizeof(T));
  const int64_t chunks = (s + kvox - 1) / kvox;
  const int64_t planes = (int64_t)(c / 8) * n;
  int64_t gx = kNumSMs / planes;
  if (gx < 1) gx = 1;
  if (gx > chunks) gx = chunks;
  return dim3((unsigned)gx, (unsigned)(c / 8), (unsigned)n);
}

// ---------------------------------------------------------------------------------------------------
struct BnActParams {
  float scale[8], shift[8], a1[8], a2[8];
};

__device__ __forceinline__ void load_params(BnActParams& p, const float* bnbuf, const float* alpha1,
                                            const float* alpha2, int c_total, int groups, int g, int c8) {
  const int64_t gc = (int64_t)groups * c_total;
  const int64_t off = (int64_t)g * c_total + c8 * 8;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    p.scale[j] = __ldg(bnbuf + off + j);
    p.shift[j] = __ldg(bnbuf + gc + off + j);
    p.a1[j] = __ldg(alpha1 + c8 * 8 + j);
    p.a2[j] = alpha2 ? __ldg(alpha2 + c8 * 8 + j) : 0.f;
  }
}

template <typename T, bool HAS_RES, bool HAS_TILE>
__global__ void __launch_bounds__(kStreamThreads)
    bn_act_fwd_kernel(msb_tensor y, msb_tensor out, msb_tensor res, const float* __restrict__ tile_src, int tile_c,
                      const float* __restrict__ bnbuf, const float* __restrict__ alpha1,
                      const float* __restrict__ alpha2, int64_t s, int groups) {
  pdl_wait();
  pdl_trigger();
  const int c8 = blockIdx.y, n = blockIdx.z;
  BnActParams p;
  load_params(p, bnbuf, alpha1, alpha2, y.c, groups, groups == 1 ? 0 : n, c8);
  extern __shared__ uint8_t stream_smem[];
  using PS = PlaneStream<T, 2>;
  PS ps;
  ps.s = s; ps.nchunks = (s + PS::kVox - 1) / PS::kVox;
  ps.src[0] = plane_base<T>(y, n, c8, s); ps.on[0] = true;
  ps.src[1] = HAS_RES ? plane_base<T>(res, n, c8, s) : nullptr; ps.on[1] = HAS_RES;
  ps.init(stream_smem);
  ps.prologue();
  T* outp = const_cast<T*>(plane_base<T>(out, n, c8, s));
  int64_t k = 0;
  for (int64_t chunk = blockIdx.x; chunk < ps.nchunks; chunk += gridDim.x, ++k) {
    ps.wait(k);
    const int nv = ps.chunk_voxels(chunk);
    const T* ys = ps.stage(k, 0);
    const T* rs = ps.stage(k, 1);
    const int64_t v0 = chunk * PS::kVox;
    for (int lv = threadIdx.x; lv < nv; lv += kStreamThreads) {
      float a[8], r[8];
      Vec8<T>::load(ys + lv * 8, a);
      if (HAS_RES) Vec8<T>::load(rs + lv * 8, r);
      float tile1 = 0.f;  // in_channels == 1 (every reference config): one load per voxel instead of eight
      if (HAS_TILE && tile_c == 1) tile1 = __ldg(tile_src + (int64_t)n * s + v0 + lv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(a[j], p.scale[j], p.shift[j]);
        if (HAS_TILE) t += tile_c == 1 ? tile1 : __ldg(tile_src + ((int64_t)n * tile_c + ((c8 * 8 + j) % tile_c)) * s + v0 + lv);
        t = prelu(t, p.a1[j]);
        if (HAS_RES) t = prelu(t + r[j], p.a2[j]);
        a[j] = t;
      }
      Vec8<T>::store(outp + (v0 + lv) * 8, a);
    }
    ps.release(k, chunk);
  }
}

// bn_finalize + bn_act_fwd in ONE launch: every block derives the scale/shift of its 8 channels from the f64 sums
// (training) or the running statistics (eval); block x == 0 of each (n, plane) also publishes them in bnbuf for the
// backward kernels, and block (x == 0, n == 0) applies the running-statistics update.
template <typename T, bool HAS_RES, bool HAS_TILE>
__global__ void __launch_bounds__(kStreamThreads)
    bn_fwd_fused_kernel(msb_tensor y, msb_tensor out, msb_tensor res, const float* __restrict__ tile_src, int tile_c,
                        const double* __restrict__ sums, double count, const float* __restrict__ gamma,
                        const float* __restrict__ beta, float* __restrict__ rmean, float* __restrict__ rvar,
                        float momentum, float eps, int training, float* __restrict__ bnbuf,
                        const float* __restrict__ alpha1, const float* __restrict__ alpha2, int64_t s, int groups) {
  pdl_wait();
  pdl_trigger();
  const int c8 = blockIdx.y, n = blockIdx.z, c = y.c;
  const int g = groups == 1 ? 0 : n;
  const int64_t gc = (int64_t)groups * c;
  __shared__ float sh_scale[8], sh_shift[8];
  if (threadIdx.x < 8) {  // f64 statistics -> f32 scale/shift: one thread per channel, broadcast through smem
    const int ch = c8 * 8 + threadIdx.x;
    double mean, var;
    if (training) {
      mean = sums[(int64_t)g * c + ch] / count;
      var = sums[gc + (int64_t)g * c + ch] / count - mean * mean;
      if (var < 0) var = 0;
    } else {
      mean = rmean[ch];
      var = rvar[ch];
    }
    const float invstd = (float)(1.0 / sqrt(var + (double)eps));
    const float scale = __ldg(gamma + ch) * invstd;
    const float shift = __ldg(beta + ch) - (float)mean * scale;
    sh_scale[threadIdx.x] = scale;
    sh_shift[threadIdx.x] = shift;
    if (blockIdx.x == 0) {
      bnbuf[0 * gc + (int64_t)g * c + ch] = scale;
      bnbuf[1 * gc + (int64_t)g * c + ch] = shift;
      bnbuf[2 * gc + (int64_t)g * c + ch] = (float)mean;
      bnbuf[3 * gc + (int64_t)g * c + ch] = invstd;
    }
  }
  __syncthreads();
  BnActParams p;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int ch = c8 * 8 + j;
    p.scale[j] = sh_scale[j];
    p.shift[j] = sh_shift[j];
    p.a1[j] = __ldg(alpha1 + ch);
    p.a2[j] = alpha2 ? __ldg(alpha2 + ch) : 0.f;
  }
  if (training && rmean != nullptr && blockIdx.x == 0 && n == 0 && threadIdx.x < 8) {
    const int ch = c8 * 8 + threadIdx.x;
    double mean_acc = 0, var_acc = 0;
    for (int gg = 0; gg < groups; ++gg) {
      const double mean = sums[(int64_t)gg * c + ch] / count;
      double var = sums[gc + (int64_t)gg * c + ch] / count - mean * mean;
      if (var < 0) var = 0;
      mean_acc += mean;
      var_acc += var;
    }
    rmean[ch] = momentum * rmean[ch] + (1.f - momentum) * (float)(mean_acc / groups);
    rvar[ch] = momentum * rvar[ch] + (1.f - momentum) * (float)(var_acc / groups);
  }
  extern __shared__ uint8_t stream_smem[];
  using PS = PlaneStream<T, 2>;
  PS ps;
  ps.s = s; ps.nchunks = (s + PS::kVox - 1) / PS::kVox;
  ps.src[0] = plane_base<T>(y, n, c8, s); ps.on[0] = true;
  ps.src[1] = HAS_RES ? plane_base<T>(res, n, c8, s) : nullptr; ps.on[1] = HAS_RES;
  ps.init(stream_smem);
  ps.prologue();
  T* outp = const_cast<T*>(plane_base<T>(out, n, c8, s));
  int64_t k = 0;
  for (int64_t chunk = blockIdx.x; chunk < ps.nchunks; chunk += gridDim.x, ++k) {
    ps.wait(k);
    const int nv = ps.chunk_voxels(chunk);
    const T* ys = ps.stage(k, 0);
    const T* rs = ps.stage(k, 1);
    const int64_t v0 = chunk * PS::kVox;
    for (int lv = threadIdx.x; lv < nv; lv += kStreamThreads) {
      float a[8], r[8];
      Vec8<T>::load(ys + lv * 8, a);
      if (HAS_RES) Vec8<T>::load(rs + lv * 8, r);
      float tile1 = 0.f;  // in_channels == 1 (every reference config): one load per voxel instead of eight
      if (HAS_TILE && tile_c == 1) tile1 = __ldg(tile_src + (int64_t)n * s + v0 + lv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(a[j], p.scale[j], p.shift[j]);
        if (HAS_TILE) t += tile_c == 1 ? tile1 : __ldg(tile_src + ((int64_t)n * tile_c + ((c8 * 8 + j) % tile_c)) * s + v0 + lv);
        t = prelu(t, p.a1[j]);
        if (HAS_RES) t = prelu(t + r[j], p.a2[j]);
        a[j] = t;
      }
      Vec8<T>::store(outp + (v0 + lv) * 8, a);
    }
    ps.release(k, chunk);
  }
}

// recompute of the forward chain + gradient at the BN output (g1) and after the residual add (g2)
template <bool HAS_RES>
__device__ __forceinline__ void bwd_point(float yv, float rv, float tile, float go, float scale, float shift,
                                          float a1, float a2, float& g1, float& g2, float& da1, float& da2) {
  const float t = fmaf(yv, scale, shift) + tile;
  const float act1 = prelu(t, a1);
  if (HAS_RES) {
    const float t2 = act1 + rv;
    g2 = t2 > 0.f ? go : a2 * go;
    da2 = t2 > 0.f ? 0.f : go * t2;
  } else {
    g2 = go;
    da2 = 0.f;
  }
  g1 = t > 0.f ? g2 : a1 * g2;
  da1 = t > 0.f ? 0.f : g2 * t;
}

template <typename T, bool HAS_RES, bool HAS_TILE>
__global__ void __launch_bounds__(kStreamThreads)
    bn_act_bwd_reduce_kernel(msb_tensor y, msb_tensor res, const float* __restrict__ tile_src, int tile_c,
                             msb_tensor gout, const float* __restrict__ bnbuf, const float* __restrict__ alpha1,
                             const float* __restrict__ alpha2, int64_t s, int groups, double* __restrict__ red) {
  pdl_wait();
  pdl_trigger();
  __shared__ float sred[kStreamThreads / 32][32];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int g = groups == 1 ? 0 : n;
  BnActParams p;
  load_params(p, bnbuf, alpha1, alpha2, y.c, groups, g, c8);
  float mean[8], invstd[8];
  {
    const int64_t gc = (int64_t)groups * y.c, off = (int64_t)g * y.c + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = __ldg(bnbuf + 2 * gc + off + j);
      invstd[j] = __ldg(bnbuf + 3 * gc + off + j);
    }
  }
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  // streamed over this block's chunks of the plane (about one block per SM: the per-channel f64 atomics at the end hit
  // the same 32 addresses for every block of a plane, so few blocks also keep that tail short)
  extern __shared__ uint8_t stream_smem[];
  using PS = PlaneStream<T, 3>;
  PS ps;
  ps.s = s; ps.nchunks = (s + PS::kVox - 1) / PS::kVox;
  ps.src[0] = plane_base<T>(y, n, c8, s); ps.on[0] = true;
  ps.src[1] = plane_base<T>(gout, n, c8, s); ps.on[1] = true;
  ps.src[2] = HAS_RES ? plane_base<T>(res, n, c8, s) : nullptr; ps.on[2] = HAS_RES;
  ps.init(stream_smem);
  ps.prologue();
  int64_t k = 0;
  for (int64_t chunk = blockIdx.x; chunk < ps.nchunks; chunk += gridDim.x, ++k) {
    ps.wait(k);
    const int nv = ps.chunk_voxels(chunk);
    const T* ys = ps.stage(k, 0);
    const T* gs = ps.stage(k, 1);
    const T* rs = ps.stage(k, 2);
    const int64_t v0 = chunk * PS::kVox;
    for (int lv = threadIdx.x; lv < nv; lv += kStreamThreads) {
      float a[8], r[8], go[8];
      Vec8<T>::load(ys + lv * 8, a);
      Vec8<T>::load(gs + lv * 8, go);
      if (HAS_RES) Vec8<T>::load(rs + lv * 8, r);
      float tile1 = 0.f;  // in_channels == 1 (every reference config): one load per voxel instead of eight
      if (HAS_TILE && tile_c == 1) tile1 = __ldg(tile_src + (int64_t)n * s + v0 + lv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float tile = 0.f;
        if (HAS_TILE) tile = tile_c == 1 ? tile1 : __ldg(tile_src + ((int64_t)n * tile_c + ((c8 * 8 + j) % tile_c)) * s + v0 + lv);
        float g1, g2, da1, da2;
        bwd_point<HAS_RES>(a[j], HAS_RES ? r[j] : 0.f, tile, go[j], p.scale[j], p.shift[j], p.a1[j], p.a2[j], g1, g2,
                           da1, da2);
        const float xhat = (a[j] - mean[j]) * invstd[j];
        acc[j] += g1;
        acc[8 + j] += g1 * xhat;
        acc[16 + j] += da1;
        acc[24 + j] += da2;
      }
    }
    ps.release(k, chunk);
  }
  block_reduce_to_smem<32>(acc, &sred[0][0]);
  if (threadIdx.x < 32) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kStreamThreads / 32; ++w) t += (double)sred[w][threadIdx.x];
    const int stat = threadIdx.x >> 3, j = threadIdx.x & 7;
    atomicAdd(&red[((int64_t)stat * groups + g) * y.c + c8 * 8 + j], t);
  }
}

template <typename T, bool HAS_RES, bool HAS_TILE, bool HAS_DRES>
__global__ void __launch_bounds__(kStreamThreads)
    bn_act_bwd_apply_kernel(msb_tensor y, msb_tensor res, const float* __restrict__ tile_src, int tile_c,
                            msb_tensor gout, const float* __restrict__ bnbuf, const float* __restrict__ alpha1,
                            const float* __restrict__ alpha2, const double* __restrict__ red, double count,
                            int training, msb_tensor dy, msb_tensor dres, int dres_acc, int64_t s, int groups,
                            float* __restrict__ dgamma, float* __restrict__ dbeta, float* __restrict__ dalpha1,
                            float* __restrict__ dalpha2) {
  pdl_wait();
  pdl_trigger();
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int g = groups == 1 ? 0 : n;
  if (blockIdx.x == 0 && n == 0 && threadIdx.x < 8) {  // parameter gradients (one block per plane owns them)
    const int ch = c8 * 8 + threadIdx.x;
    const int64_t gcc = (int64_t)groups * y.c;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    for (int gg = 0; gg < groups; ++gg) {
      s0 += red[0 * gcc + (int64_t)gg * y.c + ch];
      s1 += red[1 * gcc + (int64_t)gg * y.c + ch];
      s2 += red[2 * gcc + (int64_t)gg * y.c + ch];
      s3 += red[3 * gcc + (int64_t)gg * y.c + ch];
    }
    if (dbeta) dbeta[ch] += (float)s0;
    if (dgamma) dgamma[ch] += (float)s1;
    if (dalpha1) dalpha1[ch] += (float)s2;
    if (dalpha2) dalpha2[ch] += (float)s3;
  }
  BnActParams p;
  load_params(p, bnbuf, alpha1, alpha2, y.c, groups, g, c8);
  float mean[8], invstd[8], m_g1[8], m_g1x[8];
  {
    const int64_t gc = (int64_t)groups * y.c, off = (int64_t)g * y.c + c8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      mean[j] = __ldg(bnbuf + 2 * gc + off + j);
      invstd[j] = __ldg(bnbuf + 3 * gc + off + j);
      m_g1[j] = training ? (float)(red[off + j] / count) : 0.f;
      m_g1x[j] = training ? (float)(red[gc + off + j] / count) : 0.f;
    }
  }
  extern __shared__ uint8_t stream_smem[];
  using PS = PlaneStream<T, 4>;
  PS ps;
  ps.s = s; ps.nchunks = (s + PS::kVox - 1) / PS::kVox;
  ps.src[0] = plane_base<T>(y, n, c8, s); ps.on[0] = true;
  ps.src[1] = plane_base<T>(gout, n, c8, s); ps.on[1] = true;
  ps.src[2] = HAS_RES ? plane_base<T>(res, n, c8, s) : nullptr; ps.on[2] = HAS_RES;
  ps.src[3] = (HAS_DRES && dres_acc) ? plane_base<T>(dres, n, c8, s) : nullptr; ps.on[3] = HAS_DRES && dres_acc;
  ps.init(stream_smem);
  ps.prologue();
  T* dyp = const_cast<T*>(plane_base<T>(dy, n, c8, s));
  T* drp = HAS_DRES ? const_cast<T*>(plane_base<T>(dres, n, c8, s)) : nullptr;
  int64_t k = 0;
  for (int64_t chunk = blockIdx.x; chunk < ps.nchunks; chunk += gridDim.x, ++k) {
    ps.wait(k);
    const int nv = ps.chunk_voxels(chunk);
    const T* ys = ps.stage(k, 0);
    const T* gs = ps.stage(k, 1);
    const T* rs = ps.stage(k, 2);
    const T* ds_ = ps.stage(k, 3);
    const int64_t v0 = chunk * PS::kVox;
    for (int lv = threadIdx.x; lv < nv; lv += kStreamThreads) {
      float a[8], r[8], go[8], dr[8];
      Vec8<T>::load(ys + lv * 8, a);
      Vec8<T>::load(gs + lv * 8, go);
      if (HAS_RES) Vec8<T>::load(rs + lv * 8, r);
      if (HAS_DRES && dres_acc) Vec8<T>::load(ds_ + lv * 8, dr);
      float tile1 = 0.f;  // in_channels == 1 (every reference config): one load per voxel instead of eight
      if (HAS_TILE && tile_c == 1) tile1 = __ldg(tile_src + (int64_t)n * s + v0 + lv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float tile = 0.f;
        if (HAS_TILE) tile = tile_c == 1 ? tile1 : __ldg(tile_src + ((int64_t)n * tile_c + ((c8 * 8 + j) % tile_c)) * s + v0 + lv);
        float g1, g2, da1, da2;
        bwd_point<HAS_RES>(a[j], HAS_RES ? r[j] : 0.f, tile, go[j], p.scale[j], p.shift[j], p.a1[j], p.a2[j], g1, g2,
                           da1, da2);
        const float xhat = (a[j] - mean[j]) * invstd[j];
        a[j] = p.scale[j] * (g1 - m_g1[j] - xhat * m_g1x[j]);
        if (HAS_DRES) dr[j] = (dres_acc ? dr[j] : 0.f) + g2;
      }
      Vec8<T>::store(dyp + (v0 + lv) * 8, a);
      if (HAS_DRES) Vec8<T>::store(drp + (v0 + lv) * 8, dr);
    }
    ps.release(k, chunk);
  }
}

// ---------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads) channel_scale_kernel(msb_tensor src, msb_tensor dst,
                                                                 const float* __restrict__ scale, int64_t s,
                                                                 int accumulate) {
  pdl_wait();
  pdl_trigger();
  const int c8 = blockIdx.y, n = blockIdx.z;
  float sc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) sc[j] = scale ? __ldg(scale + (int64_t)n * src.c + c8 * 8 + j) : 1.f;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float a[8], d[8];
    Vec8<T>::load(view_ptr<T>(src, n, c8, s, v), a);
    if (accumulate) Vec8<T>::load(view_ptr<T>(dst, n, c8, s, v), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = (accumulate ? d[j] : 0.f) + a[j] * sc[j];
    Vec8<T>::store(view_ptr<T>(dst, n, c8, s, v), a);
  }
}

// ---------------------------------------------------------------------------------------------------
// OutputTransition.conv2: 1x1x1 conv C->C producing NCDHW f32 logits (vnet.py:169,174)
constexpr int kHeadMaxC = 32;

template <typename T, int CI8>
__global__ void __launch_bounds__(kThreads) conv1x1_fwd_kernel(msb_tensor a, const float* __restrict__ w,
                                                               const float* __restrict__ b,
                                                               float* __restrict__ logits, int ci, int co, int64_t s) {
  pdl_wait();
  pdl_trigger();
  __shared__ float ws[kHeadMaxC * kHeadMaxC + kHeadMaxC];
  for (int i = threadIdx.x; i < co * ci; i += kThreads) ws[i] = w[i];
  for (int i = threadIdx.x; i < co; i += kThreads) ws[kHeadMaxC * kHeadMaxC + i] = b ? b[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float x[CI8 * 8];
#pragma unroll
    for (int k = 0; k < CI8; ++k) {
      float t[8];
      Vec8<T>::load(view_ptr<T>(a, n, k, s, v), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) x[k * 8 + j] = t[j];
    }
    for (int o = 0; o < co; ++o) {
      float acc = ws[kHeadMaxC * kHeadMaxC + o];
#pragma unroll
      for (int i = 0; i < CI8 * 8; ++i)
        if (i < ci) acc = fmaf(ws[o * ci + i], x[i], acc);
      logits[((int64_t)n * co + o) * s + v] = acc;
    }
  }
}

// Fast path of the head for the channel widths the model produces (8 / 16 / 32): the logits of a voxel accumulate in
// registers, the weights are read as float4 rows of the TRANSPOSED matrix (one broadcast shared-memory read feeds four
// FMAs; the scalar loop above is bound by shared-memory bandwidth at C = 20).  Same fmaf order per output: bit-identical.
template <typename T, int CI8, int CMAX>
__global__ void __launch_bounds__(kThreads) conv1x1_fwd_reg_kernel(msb_tensor a, const float* __restrict__ w,
                                                                   const float* __restrict__ b,
                                                                   float* __restrict__ logits, int ci, int co,
                                                                   int64_t s) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float wt[CI8 * 8 * CMAX + CMAX];  // wt[input j][output o], then the bias
  for (int i = threadIdx.x; i < CI8 * 8 * CMAX; i += kThreads) {
    const int j = i / CMAX, o = i % CMAX;
    wt[i] = (o < co && j < ci) ? w[o * ci + j] : 0.f;
  }
  for (int i = threadIdx.x; i < CMAX; i += kThreads) wt[CI8 * 8 * CMAX + i] = (b != nullptr && i < co) ? b[i] : 0.f;
  __syncthreads();
  const int n = blockIdx.z;
  const int64_t v0 = (int64_t)blockIdx.x * kVoxPerBlock;
  const int64_t v1 = min(v0 + (int64_t)kVoxPerBlock, s);
  const int groups = (ci + 7) >> 3;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += kThreads) {
    float z[CMAX];
#pragma unroll
    for (int o = 0; o < CMAX; ++o) z[o] = wt[CI8 * 8 * CMAX + o];
#pragma unroll 1
    for (int k = 0; k < groups; ++k) {
      float t[8];
      Vec8<T>::load(view_ptr<T>(a, n, k, s, v), t);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float4* wrow = reinterpret_cast<const float4*>(wt + (k * 8 + j) * CMAX);
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
      if (o < co) logits[((int64_t)n * co + o) * s + v] = z[o];
  }
}

constexpr int kHeadThreads = 512;  // phase 1: one voxel per thread (da); phase 2: one (o, i) weight pair per thread
constexpr int kHeadTile = 512;     // voxels staged per inner step of the backward kernel
constexpr int kHeadPitch = kHeadTile + 1;
template <int CI8>
constexpr int head_bwd_smem() { return (CI8 * 8 + 1 + kHeadMaxC) * kHeadPitch * (int)sizeof(float); }

// dlogits -> da (B8) and dW / db.  Phase 1: every thread loads one voxel, writes da and stages (a, dlogits) in shared
// memory (plus a row of ones, so the bias gradient is one more column of the pair matrix).  Phase 2 (many pairs, e.g.
// the 420 of the C = 20 MRI head): the [co] x [ci + 1] pair matrix is cut into 4 x 4 REGISTER tiles, one tile per lane,
// and each of the 16 warps runs over its own 32 of the 512 staged voxels: 8 broadcast shared-memory reads feed 16 FMAs
// (a thread-per-pair loop needs 2 reads per FMA and is bound by shared-memory bandwidth).  Few pairs (2-3 classes):
// one warp per pair.  Block-level reduction in shared memory, then one atomic per pair and block.
template <typename T, int CI8>
__global__ void __launch_bounds__(kHeadThreads)
    conv1x1_bwd_kernel(msb_tensor a, const float* __restrict__ w, const float* __restrict__ dlogits, msb_tensor da,
                       float* __restrict__ dw, float* __restrict__ db, int ci, int co, int64_t s) {
  pdl_wait();
  pdl_trigger();
  __shared__ __align__(16) float ws[kHeadMaxC][CI8 * 8];  // W[o][i], rows zero-padded to the plane width
  __shared__ float accw[kHeadThreads / 32 * 4];           // per-pair sums of the warp-per-pair path
  if (threadIdx.x < kHeadThreads / 32 * 4) accw[threadIdx.x] = 0.f;
  extern __shared__ float head_smem[];
  constexpr int kOnes = CI8 * 8;  // row of ones (valid voxels) behind the activation rows
  float (*as)[kHeadPitch] = reinterpret_cast<float (*)[kHeadPitch]>(head_smem);                                // [CI8*8 + 1]
  float (*ds)[kHeadPitch] = reinterpret_cast<float (*)[kHeadPitch]>(head_smem + (CI8 * 8 + 1) * kHeadPitch);   // [kHeadMaxC]
  for (int i = threadIdx.x; i < kHeadMaxC * CI8 * 8; i += kHeadThreads) {
    const int o = i / (CI8 * 8), c = i % (CI8 * 8);
    ws[o][c] = (o < co && c < ci) ? w[o * ci + c] : 0.f;
  }
  const int n = blockIdx.z;
  const int npairs = co * ci + co;  // last `co` pseudo-pairs accumulate the bias gradient
  const bool few_pairs = npairs <= kHeadThreads / 32 * 4;
  // register tiles of the many-pairs path: tile = (4 classes o) x (4 columns c of [a | ones]); <= 8 x 9 tiles
  const int tiles_c = (ci + 1 + 3) / 4, ntiles = ((co + 3) / 4) * tiles_c;
  constexpr int kTileSlots = (kHeadMaxC / 4 * (kHeadMaxC / 4 + 1) + 31) / 32;  // 72 tiles over 32 lanes
  float acct[kTileSlots][16];
#pragma unroll
  for (int k = 0; k < kTileSlots; ++k)
#pragma unroll
    for (int e = 0; e < 16; ++e) acct[k][e] = 0.f;
  for (int i = threadIdx.x; i < (kHeadMaxC - co) * kHeadPitch; i += kHeadThreads)
    ds[co][i] = 0.f;  // class rows beyond co are read by the padded tiles: keep them finite
  __syncthreads();
  // grid-stride over 128-voxel tiles: the launch uses a few blocks per SM, so the final atomics (all blocks hit the
  // same <= 1056 addresses - a few cache lines) stay in the low hundreds per address
  for (int64_t base = (int64_t)blockIdx.x * kHeadTile; base < s; base += (int64_t)gridDim.x * kHeadTile) {
    {
      const int64_t v = base + threadIdx.x;
      const bool valid = v < s;
#pragma unroll
      for (int k = 0; k < CI8; ++k) {
        float t[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) t[j] = 0.f;
        if (valid) Vec8<T>::load(view_ptr<T>(a, n, k, s, v), t);
#pragma unroll
        for (int j = 0; j < 8; ++j) as[k * 8 + j][threadIdx.x] = t[j];
      }
      as[kOnes][threadIdx.x] = valid ? 1.f : 0.f;
      // da[i] = sum_o W[o][i] * g[o]: runtime loop over the classes, the da accumulators stay in registers
      float dacc[CI8 * 8];
#pragma unroll
      for (int i = 0; i < CI8 * 8; ++i) dacc[i] = 0.f;
      for (int o = 0; o < co; ++o) {
        const float go = valid ? __ldg(dlogits + ((int64_t)n * co + o) * s + v) : 0.f;
        ds[o][threadIdx.x] = go;
#pragma unroll
        for (int i4 = 0; i4 < CI8 * 2; ++i4) {
          const float4 w4 = *reinterpret_cast<const float4*>(&ws[o][i4 * 4]);  // broadcast read
          dacc[i4 * 4 + 0] = fmaf(w4.x, go, dacc[i4 * 4 + 0]);
          dacc[i4 * 4 + 1] = fmaf(w4.y, go, dacc[i4 * 4 + 1]);
          dacc[i4 * 4 + 2] = fmaf(w4.z, go, dacc[i4 * 4 + 2]);
          dacc[i4 * 4 + 3] = fmaf(w4.w, go, dacc[i4 * 4 + 3]);
        }
      }
      if (valid) {
#pragma unroll
        for (int k = 0; k < CI8; ++k) {
          float t[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) t[j] = dacc[k * 8 + j];
          Vec8<T>::store(view_ptr<T>(da, n, k, s, v), t);
        }
      }
    }
    __syncthreads();
    if (few_pairs) {
      // few pairs (2-3 classes): one WARP per pair, lanes stride over the staged voxels (a thread-per-pair loop would
      // be one 512-long dependent FMA chain on a handful of threads)
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      for (int pidx = warp; pidx < npairs; pidx += kHeadThreads / 32) {
        float acc = 0.f;
        if (pidx < co * ci) {
          const float* ar = as[pidx % ci];
          const float* dr = ds[pidx / ci];
          for (int t = lane; t < kHeadTile; t += 32) acc = fmaf(ar[t], dr[t], acc);
        } else {
          const float* dr = ds[pidx - co * ci];
          for (int t = lane; t < kHeadTile; t += 32) acc += dr[t];
        }
        acc = warp_sum(acc);
        if (lane == 0) accw[pidx] += acc;
      }
    } else {
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int t0 = warp * (kHeadTile / (kHeadThreads / 32));
#pragma unroll
      for (int slot = 0; slot < kTileSlots; ++slot) {
        const int tile = lane + slot * 32;
        if (tile >= ntiles) continue;
        const int o0 = (tile / tiles_c) * 4, c0 = (tile % tiles_c) * 4;
        const float* ar[4];
        const float* dr[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          ar[q] = as[c0 + q < ci ? c0 + q : kOnes];  // column ci (and the padding behind it) reads the ones row
          dr[q] = ds[o0 + q];                        // rows >= co are zero
        }
#pragma unroll 4
        for (int t = t0; t < t0 + kHeadTile / (kHeadThreads / 32); ++t) {
          float av[4], dv[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) { av[q] = ar[q][t]; dv[q] = dr[q][t]; }
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int q = 0; q < 4; ++q) acct[slot][r * 4 + q] = fmaf(dv[r], av[q], acct[slot][r * 4 + q]);
        }
      }
    }
    __syncthreads();
  }
  if (few_pairs) {
    for (int pidx = threadIdx.x; pidx < npairs; pidx += kHeadThreads) {
      if (pidx < co * ci) atomicAdd(dw + pidx, accw[pidx]);
      else if (db) atomicAdd(db + (pidx - co * ci), accw[pidx]);
    }
    return;
  }
  // block reduction of the 16 warps' register tiles through shared memory (the staging buffers are free now)
  float* red = head_smem;  // [warps][ntiles * 16]
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int slot = 0; slot < kTileSlots; ++slot) {
      const int tile = lane + slot * 32;
      if (tile >= ntiles) continue;
#pragma unroll
      for (int e = 0; e < 16; ++e) red[(warp * ntiles + tile) * 16 + e] = acct[slot][e];
    }
  }
  __syncthreads();
  for (int pidx = threadIdx.x; pidx < npairs; pidx += kHeadThreads) {
    const int o = pidx < co * ci ? pidx / ci : pidx - co * ci;
    const int c = pidx < co * ci ? pidx % ci : ci;
    const int e = ((o >> 2) * tiles_c + (c >> 2)) * 16 + (o & 3) * 4 + (c & 3);
    float sum = 0.f;
#pragma unroll
    for (int wv = 0; wv < kHeadThreads / 32; ++wv) sum += red[wv * ntiles * 16 + e];
    if (pidx < co * ci) atomicAdd(dw + pidx, sum);
    else if (db) atomicAdd(db + (pidx - co * ci), sum);
  }
}

// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) momentum_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                       float* __restrict__ v, int64_t count, float lr,
                                                       const float* __restrict__ lr_dev, float mu, float wd, float gs) {
  pdl_wait();
  pdl_trigger();
  if (lr_dev != nullptr) lr = __ldg(lr_dev);  // learning rate read from device memory (CUDA-graph replays)
  const int64_t n4 = count >> 2;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    const float4 gg = reinterpret_cast<const float4*>(g)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    vv.x = fmaf(mu, vv.x, fmaf(gs, gg.x, wd * pp.x));
    vv.y = fmaf(mu, vv.y, fmaf(gs, gg.y, wd * pp.y));
    vv.z = fmaf(mu, vv.z, fmaf(gs, gg.z, wd * pp.z));
    vv.w = fmaf(mu, vv.w, fmaf(gs, gg.w, wd * pp.w));
    pp.x -= lr * vv.x; pp.y -= lr * vv.y; pp.z -= lr * vv.z; pp.w -= lr * vv.w;
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(v)[i] = vv;
  }
  for (int64_t i = (n4 << 2) + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
    const float vv = fmaf(mu, v[i], fmaf(gs, g[i], wd * p[i]));
    v[i] = vv;
    p[i] -= lr * vv;
  }
}

}  // namespace msb

using namespace msb;

extern "C" {

int msb_version(void) { return MSB_VERSION; }
const char* msb_last_error_string(void) { return msb::last_error(); }

int msb_zero(void* ptr, size_t bytes, void* stream) {
  MSB_REQUIRE(ptr != nullptr || bytes == 0, "msb_zero: null pointer");
  if (bytes) MSB_CUDA_OK(cudaMemsetAsync(ptr, 0, bytes, as_stream(stream)));
  return MSB_OK;
}

int msb_to_blocked(const float* src, int n, int c, int64_t s, msb_tensor dst, void* stream) {
  MSB_REQUIRE(src && view_ok(dst) && n > 0 && c > 0 && c <= dst.c && s > 0, "msb_to_blocked: bad arguments");
  MSB_DISPATCH_DTYPE(dst.dtype, to_blocked_kernel<T><<<plane_grid(n, dst.c, s), kThreads, 0, as_stream(stream)>>>(
                                    src, c, s, dst););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_from_blocked(msb_tensor src, float* dst, int n, int c, int64_t s, void* stream) {
  MSB_REQUIRE(dst && view_ok(src) && n > 0 && c > 0 && c <= src.c && s > 0, "msb_from_blocked: bad arguments");
  MSB_DISPATCH_DTYPE(src.dtype,
                     from_blocked_kernel<T><<<plane_grid(n, ((c + 7) / 8) * 8, s), kThreads, 0, as_stream(stream)>>>(
                         src, dst, c, s););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_bn_stats(msb_tensor x, int n, int64_t s, int groups, double* sums, void* stream) {
  MSB_REQUIRE(view_ok(x) && sums && n > 0 && s > 0 && (groups == 1 || groups == n), "msb_bn_stats: bad arguments");
  MSB_DISPATCH_DTYPE(x.dtype, bn_stats_kernel<T><<<plane_grid(n, x.c, s), kThreads, 0, as_stream(stream)>>>(
                                  x, s, groups, x.c, sums););
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_bn_finalize(const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int training, int c, int groups, float* bnbuf,
                    void* stream) {
  MSB_REQUIRE(gamma && beta && bnbuf && c > 0 && groups > 0, "msb_bn_finalize: bad arguments");
  MSB_REQUIRE(training ? (sums != nullptr && count > 0) : (running_mean && running_var),
              "msb_bn_finalize: training needs sums, eval needs running stats");
  MSB_LAUNCH_PDL(bn_finalize_kernel, dim3((c + 127) / 128), dim3(128), 0, as_stream(stream), sums, count, gamma, beta,
                 running_mean, running_var, momentum, eps, training, c, groups, bnbuf);
  return MSB_OK;
}

// launches a PlaneStream kernel: ~1 block per SM, kStreamThreads threads, NIN * 48 KB of dynamic shared memory
#define MSB_LAUNCH_STREAM(kernel, T_, NIN_, n_, c_, s_, st_, ...)                                              \
  do {                                                                                                         \
    constexpr int _smem = PlaneStream<T_, NIN_>::kSmemBytes;                                                   \
    static bool _attr = false;                                                                                 \
    if (!_attr) {                                                                                              \
      MSB_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, _smem));           \
      _attr = true;                                                                                            \
    }                                                                                                          \
    MSB_LAUNCH_PDL(kernel, stream_grid<T_>(n_, c_, s_), dim3(kStreamThreads), _smem, st_, __VA_ARGS__);        \
  } while (0)

#define MSB_BOOL_DISPATCH2(b0, b1, ...)                          \
  do {                                                           \
    if (b0) {                                                    \
      constexpr bool B0 = true;                                  \
      if (b1) { constexpr bool B1 = true; __VA_ARGS__ }          \
      else { constexpr bool B1 = false; __VA_ARGS__ }            \
    } else {                                                     \
      constexpr bool B0 = false;                                 \
      if (b1) { constexpr bool B1 = true; __VA_ARGS__ }          \
      else { constexpr bool B1 = false; __VA_ARGS__ }            \
    }                                                            \
  } while (0)

int msb_bn_act_fwd(msb_tensor y, msb_tensor out, msb_tensor residual, const float* tile_src, int tile_c,
                   const float* bnbuf, const float* alpha1, const float* alpha2, int n, int64_t s, int groups,
                   void* stream) {
  const bool has_res = residual.ptr != nullptr;
  MSB_REQUIRE(view_ok(y) && view_ok(out) && out.c == y.c && out.dtype == y.dtype && bnbuf && alpha1,
              "msb_bn_act_fwd: bad y/out/bnbuf/alpha1");
  MSB_REQUIRE(!has_res || (view_ok(residual) && residual.c == y.c && residual.dtype == y.dtype && alpha2),
              "msb_bn_act_fwd: residual needs matching view and alpha2");
  MSB_REQUIRE(!tile_src || tile_c > 0, "msb_bn_act_fwd: tile_c must be > 0");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_bn_act_fwd: groups must be 1 or n");
  MSB_DISPATCH_DTYPE(y.dtype, MSB_BOOL_DISPATCH2(has_res, tile_src != nullptr,
                                                 MSB_LAUNCH_STREAM((bn_act_fwd_kernel<T, B0, B1>), T, 2, n, y.c, s,
                                                                   as_stream(stream), y, out, residual, tile_src, tile_c,
                                                                   bnbuf, alpha1, alpha2, s, groups);););
  return MSB_OK;
}

int msb_bn_fwd_fused(msb_tensor y, msb_tensor out, msb_tensor residual, const float* tile_src, int tile_c,
                     const double* sums, double count, const float* gamma, const float* beta, float* running_mean,
                     float* running_var, float momentum, float eps, int training, float* bnbuf, const float* alpha1,
                     const float* alpha2, int n, int64_t s, int groups, void* stream) {
  const bool has_res = residual.ptr != nullptr;
  MSB_REQUIRE(view_ok(y) && view_ok(out) && out.c == y.c && out.dtype == y.dtype && bnbuf && alpha1 && gamma && beta,
              "msb_bn_fwd_fused: bad y/out/bnbuf/alpha1/gamma/beta");
  MSB_REQUIRE(!has_res || (view_ok(residual) && residual.c == y.c && residual.dtype == y.dtype && alpha2),
              "msb_bn_fwd_fused: residual needs matching view and alpha2");
  MSB_REQUIRE(!tile_src || tile_c > 0, "msb_bn_fwd_fused: tile_c must be > 0");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_bn_fwd_fused: groups must be 1 or n");
  MSB_REQUIRE(training ? (sums != nullptr && count > 0) : (running_mean && running_var),
              "msb_bn_fwd_fused: training needs sums, eval needs running stats");
  MSB_DISPATCH_DTYPE(y.dtype, MSB_BOOL_DISPATCH2(has_res, tile_src != nullptr,
                                                 MSB_LAUNCH_STREAM((bn_fwd_fused_kernel<T, B0, B1>), T, 2, n, y.c, s,
                                                                   as_stream(stream), y, out, residual, tile_src, tile_c,
                                                                   sums, count, gamma, beta, running_mean, running_var,
                                                                   momentum, eps, training, bnbuf, alpha1, alpha2, s,
                                                                   groups);););
  return MSB_OK;
}

int msb_bn_act_bwd_reduce(msb_tensor y, msb_tensor residual, const float* tile_src, int tile_c, msb_tensor gout,
                          const float* bnbuf, const float* alpha1, const float* alpha2, int n, int64_t s, int groups,
                          double* red, void* stream) {
  const bool has_res = residual.ptr != nullptr;
  MSB_REQUIRE(view_ok(y) && view_ok(gout) && gout.c == y.c && gout.dtype == y.dtype && bnbuf && alpha1 && red,
              "msb_bn_act_bwd_reduce: bad arguments");
  MSB_REQUIRE(!has_res || (view_ok(residual) && residual.c == y.c && residual.dtype == y.dtype && alpha2),
              "msb_bn_act_bwd_reduce: residual needs matching view and alpha2");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_bn_act_bwd_reduce: groups must be 1 or n");
  MSB_DISPATCH_DTYPE(y.dtype, MSB_BOOL_DISPATCH2(has_res, tile_src != nullptr,
                                                 MSB_LAUNCH_STREAM((bn_act_bwd_reduce_kernel<T, B0, B1>), T, 3, n, y.c, s,
                                                                   as_stream(stream), y, residual, tile_src, tile_c, gout,
                                                                   bnbuf, alpha1, alpha2, s, groups, red);););
  return MSB_OK;
}

int msb_bn_act_bwd_apply(msb_tensor y, msb_tensor residual, const float* tile_src, int tile_c, msb_tensor gout,
                         const float* bnbuf, const float* alpha1, const float* alpha2, const double* red, double count,
                         int training, msb_tensor dy, msb_tensor dres, int dres_accumulate, float* dgamma,
                         float* dbeta, float* dalpha1, float* dalpha2, int n, int64_t s, int groups, void* stream) {
  const bool has_res = residual.ptr != nullptr;
  const bool has_dres = dres.ptr != nullptr;
  MSB_REQUIRE(view_ok(y) && view_ok(gout) && view_ok(dy) && gout.c == y.c && dy.c == y.c && gout.dtype == y.dtype &&
                  dy.dtype == y.dtype && bnbuf && alpha1 && red && count > 0,
              "msb_bn_act_bwd_apply: bad arguments");
  MSB_REQUIRE(!has_res || (view_ok(residual) && residual.c == y.c && residual.dtype == y.dtype && alpha2),
              "msb_bn_act_bwd_apply: residual needs matching view and alpha2");
  MSB_REQUIRE(!has_dres || (has_res && view_ok(dres) && dres.c == y.c && dres.dtype == y.dtype),
              "msb_bn_act_bwd_apply: dres needs a residual and a matching view");
  MSB_REQUIRE(groups == 1 || groups == n, "msb_bn_act_bwd_apply: groups must be 1 or n");
  cudaStream_t st = as_stream(stream);
  MSB_DISPATCH_DTYPE(
      y.dtype, MSB_BOOL_DISPATCH2(has_res, tile_src != nullptr, {
        if (has_dres)
          MSB_LAUNCH_STREAM((bn_act_bwd_apply_kernel<T, B0, B1, true>), T, 4, n, y.c, s, st, y, residual, tile_src,
                            tile_c, gout, bnbuf, alpha1, alpha2, red, count, training, dy, dres, dres_accumulate, s,
                            groups, dgamma, dbeta, dalpha1, has_res ? dalpha2 : nullptr);
        else
          MSB_LAUNCH_STREAM((bn_act_bwd_apply_kernel<T, B0, B1, false>), T, 4, n, y.c, s, st, y, residual, tile_src,
                            tile_c, gout, bnbuf, alpha1, alpha2, red, count, training, dy, dres, dres_accumulate, s,
                            groups, dgamma, dbeta, dalpha1, has_res ? dalpha2 : nullptr);
      }););
  return MSB_OK;
}

int msb_channel_scale(msb_tensor src, msb_tensor dst, const float* scale, int n, int64_t s, int accumulate,
                      void* stream) {
  MSB_REQUIRE(view_ok(src) && view_ok(dst) && src.c == dst.c && src.dtype == dst.dtype && n > 0 && s > 0,
              "msb_channel_scale: bad arguments");
  MSB_DISPATCH_DTYPE(src.dtype, MSB_LAUNCH_PDL(channel_scale_kernel<T>, plane_grid(n, src.c, s), dim3(kThreads), 0,
                                               as_stream(stream), src, dst, scale, s, accumulate););
  return MSB_OK;
}

int msb_conv1x1_fwd(msb_tensor a, const float* w, const float* b, float* logits, int n, int ci, int co, int64_t s,
                    void* stream) {
  MSB_REQUIRE(view_ok(a) && w && logits && ci > 0 && ci <= a.c && a.c <= kHeadMaxC && co > 0 && co <= a.c,
              "msb_conv1x1_fwd: needs ci <= a.c <= 32 and co <= a.c");
  const dim3 grid((unsigned)((s + kVoxPerBlock - 1) / kVoxPerBlock), 1, (unsigned)n);
  cudaStream_t st = as_stream(stream);
#define MSB_HEAD_FWD(CI8_, CMAX_) \
  MSB_LAUNCH_PDL((conv1x1_fwd_reg_kernel<T, CI8_, CMAX_>), grid, dim3(kThreads), 0, st, a, w, b, logits, ci, co, s)
  if (a.c == 8 || a.c == 16 || a.c == 32) {
    MSB_DISPATCH_DTYPE(a.dtype, {
      if (a.c == 8) { if (co <= 4) MSB_HEAD_FWD(1, 4); else MSB_HEAD_FWD(1, 8); }
      else if (a.c == 16) { if (co <= 4) MSB_HEAD_FWD(2, 4); else if (co <= 8) MSB_HEAD_FWD(2, 8); else MSB_HEAD_FWD(2, 16); }
      else { if (co <= 20) MSB_HEAD_FWD(4, 20); else MSB_HEAD_FWD(4, 32); }
    });
    MSB_LAUNCH_OK();
    return MSB_OK;
  }
#undef MSB_HEAD_FWD
  MSB_DISPATCH_DTYPE(a.dtype, {
    if (a.c == 8) MSB_LAUNCH_PDL((conv1x1_fwd_kernel<T, 1>), grid, dim3(kThreads), 0, st, a, w, b, logits, ci, co, s);
    else if (a.c == 16) MSB_LAUNCH_PDL((conv1x1_fwd_kernel<T, 2>), grid, dim3(kThreads), 0, st, a, w, b, logits, ci, co, s);
    else if (a.c == 24) MSB_LAUNCH_PDL((conv1x1_fwd_kernel<T, 3>), grid, dim3(kThreads), 0, st, a, w, b, logits, ci, co, s);
    else MSB_LAUNCH_PDL((conv1x1_fwd_kernel<T, 4>), grid, dim3(kThreads), 0, st, a, w, b, logits, ci, co, s);
  });
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_conv1x1_bwd(msb_tensor a, const float* w, const float* dlogits, msb_tensor da, float* dw, float* db, int n,
                    int ci, int co, int64_t s, void* stream) {
  MSB_REQUIRE(view_ok(a) && view_ok(da) && da.c == a.c && da.dtype == a.dtype && w && dlogits && dw && ci > 0 &&
                  ci <= a.c && a.c <= kHeadMaxC && co > 0 && co <= a.c,
              "msb_conv1x1_bwd: needs ci <= a.c <= 32 and co <= a.c");
  int64_t gx = (s + kHeadTile - 1) / kHeadTile;
  const int64_t cap = (kNumSMs * 3 + n - 1) / n;
  if (gx > cap) gx = cap;
  const dim3 grid((unsigned)gx, 1, (unsigned)n);
  cudaStream_t st = as_stream(stream);
#define MSB_HEAD_BWD(CI8_)                                                                                          \
  do {                                                                                                              \
    static bool _attr = false;                                                                                      \
    if (!_attr) {                                                                                                   \
      MSB_CUDA_OK(cudaFuncSetAttribute(conv1x1_bwd_kernel<T, CI8_>, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                       head_bwd_smem<CI8_>()));                                                    \
      _attr = true;                                                                                                 \
    }                                                                                                               \
    MSB_LAUNCH_PDL((conv1x1_bwd_kernel<T, CI8_>), grid, dim3(kHeadThreads), head_bwd_smem<CI8_>(), st, a, w, dlogits, \
                   da, dw, db, ci, co, s);                                                                          \
  } while (0)
  MSB_DISPATCH_DTYPE(a.dtype, {
    if (a.c == 8) MSB_HEAD_BWD(1);
    else if (a.c == 16) MSB_HEAD_BWD(2);
    else if (a.c == 24) MSB_HEAD_BWD(3);
    else MSB_HEAD_BWD(4);
  });
#undef MSB_HEAD_BWD
  MSB_LAUNCH_OK();
  return MSB_OK;
}

int msb_momentum_step(float* p, const float* g, float* v, int64_t count, float lr, float mu, float wd,
                      float grad_scale, void* stream) {
  MSB_REQUIRE(p && g && v && count > 0, "msb_momentum_step: bad arguments");
  MSB_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
              "msb_momentum_step: buffers must be 16-byte aligned");
  int64_t want = (count / 4 + 255) / 256 + 1;
  const int blocks = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  MSB_LAUNCH_PDL(momentum_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, v, count, lr,
                 (const float*)nullptr, mu, wd, grad_scale);
  return MSB_OK;
}

int msb_momentum_step_lrdev(float* p, const float* g, float* v, int64_t count, const float* lr_dev, float mu, float wd,
                            float grad_scale, void* stream) {
  MSB_REQUIRE(p && g && v && lr_dev && count > 0, "msb_momentum_step_lrdev: bad arguments");
  MSB_REQUIRE((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(v)) % 16 == 0,
              "msb_momentum_step_lrdev: buffers must be 16-byte aligned");
  int64_t want = (count / 4 + 255) / 256 + 1;
  const int blocks = (int)(want < (int64_t)kNumSMs * 8 ? want : (int64_t)kNumSMs * 8);
  MSB_LAUNCH_PDL(momentum_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), p, g, v, count, 0.f, lr_dev, mu, wd,
                 grad_scale);
  return MSB_OK;
}

}  // extern "C"
