// Shared host helpers of the tcgen05 5x5x5 convolution kernels.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace msb {

constexpr int kNumTaps = 125;

extern int g_debug_flags[8];

// 4-D TMA map over a B8 bf16 view: dims (W*8, H, D, planes), box (box_w*8, box_h, box_d, box_p)
int make_b8_tmap(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_h, int box_d,
                 int box_p);
// same tensor, dims ordered (W*8, planes, H, D): shared-memory image [d][h][plane][w][8]
int make_b8_tmap_hmajor(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_p, int box_h,
                        int box_d);

// kh-stacked weight-gradient kernel (conv_k5_wgrad2.cu); returns MSB_ERR_UNSUPPORTED when not applicable
// kw_taps: 5 = 5x5x5 kernel (ws [125][cout][cin]); 1 = 5x5x1 kernel of the w-folded convs (ws [25][cout][cin])
int launch_wgrad_v2(const msb_tensor& x, const msb_tensor& dy, int cout, int cin, int n, msb_dim3 dims, float* ws,
                    cudaStream_t st, int kw_taps = 5);

}  // namespace msb
