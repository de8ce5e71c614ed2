// Shared host helpers of the tcgen05 5x5x5 convolution kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace msb {

constexpr int kNumTaps = 125;

extern int g_debug_flags[8];
extern int g_dynamic_tiles;  // msb_set_tile_scheduler(): 1 = persistent kernels fetch tiles from an atomic counter

// 4-D TMA map over a B8 bf16 view: dims (W*8, H, D, planes), box (box_w*8, box_h, box_d, box_p)
int make_b8_tmap(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_h, int box_d,
                 int box_p);
// same tensor, dims ordered (W*8, planes, H, D): shared-memory image [d][h][plane][w][8]
int make_b8_tmap_hmajor(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_p, int box_h,
                        int box_d);

// strided sub-lattice map (5-D, elementStrides sw / sh along w / h): shared-memory image [plane][box_h][box_w][8]
int make_b8_tmap_s2(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_h, int box_p,
                    int sw = 2, int sh = 2);

// per-translation-unit halves of msb_set_tile_scheduler (each .cu owns the counters of its kernels)
int msb_set_tile_scheduler_wgrad2(int dynamic);
int msb_set_tile_scheduler_k2s2(int dynamic);

// kh-stacked weight-gradient kernel (conv_k5_wgrad2.cu); returns MSB_ERR_UNSUPPORTED when not applicable
// kw_taps: 5 = 5x5x5 kernel (ws [125][cout][cin]); 1 = 5x5x1 kernel of the w-folded convs (ws [25][cout][cin])
int launch_wgrad_v2(const msb_tensor& x, const msb_tensor& dy, int cout, int cin, int n, msb_dim3 dims, float* ws,
                    cudaStream_t st, int kw_taps = 5);

// 16 per-lane values -> per-channel totals over the warp; lane L ends up with the total of channel L>>1.
__device__ __forceinline__ float warp_reduce16(const float (&v)[16], int lane) {
  float a[8], b[4], c[2];
  bool hi = lane & 16;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float send = hi ? v[i] : v[i + 8], keep = hi ? v[i + 8] : v[i];
    a[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
  hi = lane & 8;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float send = hi ? a[i] : a[i + 4], keep = hi ? a[i + 4] : a[i];
    b[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  hi = lane & 4;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const float send = hi ? b[i] : b[i + 2], keep = hi ? b[i + 2] : b[i];
    c[i] = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  hi = lane & 2;
  const float send = hi ? c[0] : c[1], keep = hi ? c[1] : c[0];
  float r = keep + __shfl_xor_sync(0xffffffffu, send, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}

}  // namespace msb
