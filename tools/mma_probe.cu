// Micro-benchmark: issue cost of tcgen05.mma (kind::f16, bf16 in / f32 acc) as a function of the operand
// layout (SWIZZLE_NONE strides as used by the conv kernels vs. dense vs. 128-byte swizzle), N, and cta_group.
// Measurement tool only (not part of libmedseg_b200.so):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu && ./tools/mma_probe
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace cg = cooperative_groups;

struct Params {
  int reps;
  int n;                 // MMA N
  uint32_t a_off, a_lbo, a_sbo, a_layout;
  uint32_t b_off, b_lbo, b_sbo, b_layout;
  int a_major, b_major;  // 0 = K-major, 1 = MN-major
  int a_shift_mode;      // 0 = fixed start address, 1 = walk the 25 (kh,kw) taps of a 12-wide halo row pitch
  int n_acc;             // accumulators cycled through (columns n_acc * n <= 512)
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo, uint32_t sbo, uint32_t layout) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(layout & 7) << 61;
  return d;
}

__device__ __forceinline__ bool mbar_wait_bounded(uint32_t bar, uint32_t parity) {
  for (long long spin = 0; spin < (1ll << 26); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

template <int CTAS>
__global__ void __launch_bounds__(128, 1) probe_kernel(Params p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t cta_rank = 0;
  if (CTAS == 2) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(cta_rank));
  // fill operands with small bf16 values
  {
    uint32_t* w = reinterpret_cast<uint32_t*>(smem);
    uint32_t s = 1234567u + threadIdx.x * 7919u + blockIdx.x * 104729u;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) {
      s = s * 1664525u + 1013904223u;
      // two bf16 in [-1, 1): exponent 0x3f (0.5..1) / sign random
      w[i] = (0x3f003f00u | (s & 0x807f807fu));
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  const uint32_t bar_addr = smem_u32(&bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_addr) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    if (CTAS == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_slot)) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cg::this_cluster().sync(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_slot;
  const uint32_t m = CTAS == 2 ? 256 : 128;
  const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)p.a_major << 15) | ((uint32_t)p.b_major << 16) |
                         ((uint32_t)(p.n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
  long long best = 1ll << 62;
  bool ok = true;
  for (int rep = 0; rep < 3 && ok; ++rep) {
    if (warp == 1 && lane == 0) {
      if (cta_rank == 0) {
        const uint32_t a_base = smem_u32(smem) + p.a_off, b_base = smem_u32(smem) + p.b_off;
        // descriptors precomputed (the issue loop must cost ~1 instruction per MMA, as in the real kernels)
        uint64_t a_desc[8];
        const uint32_t tap_off[8] = {0, 1, 2, 3, 4, 12, 13, 14};  // 16-byte units: (kh,kw) taps of a 12-wide halo
#pragma unroll
        for (int u = 0; u < 8; ++u)
          a_desc[u] = make_desc(a_base + (p.a_shift_mode == 1 ? tap_off[u] * 16u : 0u), p.a_lbo, p.a_sbo, p.a_layout);
        const uint64_t b_desc = make_desc(b_base, p.b_lbo, p.b_sbo, p.b_layout);
        const uint32_t d0 = tmem_base, d1 = tmem_base + (uint32_t)((p.n_acc - 1) * p.n);
        const long long t0 = clock64();
        for (int i = 0; i < p.reps; i += 8) {
          const uint32_t acc = i > 0 ? 1u : 0u;
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t d = (u & 1) ? d1 : d0;
            if (CTAS == 1)
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a_desc[u]),
                           "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
            else
              asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                           "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d), "l"(a_desc[u]),
                           "l"(b_desc), "r"(idesc), "r"(acc) : "memory");
          }
        }
        if (CTAS == 1)
          asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_addr) : "memory");
        else
          asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                       ::"r"(bar_addr), "h"((uint16_t)3) : "memory");
        ok = mbar_wait_bounded(bar_addr, rep & 1);
        const long long t1 = clock64();
        if (rep > 0 && t1 - t0 < best) best = t1 - t0;
      } else {
        ok = mbar_wait_bounded(bar_addr, rep & 1);
      }
    }
    __syncwarp();
  }
  if (warp == 1 && lane == 0 && cta_rank == 0) out[blockIdx.x / CTAS] = ok ? best : -1;
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CTAS == 2) cg::this_cluster().sync(); else __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CTAS == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

static void run(const char* name, int ctas, Params p, int grid) {
  long long* out;
  cudaMalloc(&out, sizeof(long long) * 256);
  cudaMemset(out, 0, sizeof(long long) * 256);
  const int smem = 201 * 1024 + 1024;
  cudaError_t e;
  if (ctas == 1) {
    cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe_kernel<1><<<grid, 128, smem>>>(p, out);
  } else {
    cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, probe_kernel<2>, p, out);
  }
  e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%-44s CUDA error: %s\n", name, cudaGetErrorString(e)); exit(1); }
  long long h[256];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  const int nres = grid / ctas;
  long long mn = 1ll << 62, mx = 0;
  for (int i = 0; i < nres; ++i) { if (h[i] < mn) mn = h[i]; if (h[i] > mx) mx = h[i]; }
  const double flop_per_clk = 2.0 * (ctas == 2 ? 256 : 128) * p.n * 16 / ((double)mx / p.reps) / ctas;
  printf("%-44s cta%d N=%3d grid=%3d  clk/MMA min %7.1f max %7.1f  -> %6.0f flop/clk/SM (%.0f%% of 8192)\n", name, ctas, p.n,
         grid, (double)mn / p.reps, (double)mx / p.reps, flop_per_clk, 100.0 * flop_per_clk / 8192.0);
  fflush(stdout);
  cudaFree(out);
}

int main(int argc, char** argv) {
  const bool do_cta2 = argc > 1 && strcmp(argv[1], "cta2") == 0;
  const int reps = 2000;
  const int ns[5] = {16, 32, 64, 128, 256};
  if (!do_cta2) {
    for (int grid : {1, 148}) {
      for (int ni = 0; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        // conv layout: A = haloed voxel tile (row pitch 12*16 B, c8 plane 8*20*12*16 B), B = weight image (11 blocks of n/4.. rows)
        p.a_off = 0; p.a_lbo = 8 * 20 * 12 * 16; p.a_sbo = 12 * 16; p.a_layout = 0; p.a_shift_mode = 1;
        p.b_off = 128 * 1024; p.b_lbo = (uint32_t)(n * 16 * 2 + 1024); p.b_sbo = 128; p.b_layout = 0;
        run("conv layout (A halo taps, B weights)", 1, p, grid);
      }
      for (int ni = 0; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        p.a_off = 0; p.a_lbo = 2048; p.a_sbo = 128; p.a_layout = 0; p.a_shift_mode = 0;
        p.b_off = 128 * 1024; p.b_lbo = (uint32_t)(n * 16); p.b_sbo = 128; p.b_layout = 0;
        run("dense SWIZZLE_NONE K-major", 1, p, grid);
      }
      for (int ni = 2; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        // swapped roles: A = weights (dense 128 rows), B = voxels (halo pitch), N voxels = 8 w x (n/8) h
        p.a_off = 128 * 1024; p.a_lbo = 128 * 16 * 2 + 1024; p.a_sbo = 128; p.a_layout = 0; p.a_shift_mode = 0;
        p.b_off = 0; p.b_lbo = 36 * 12 * 16; p.b_sbo = 12 * 16; p.b_layout = 0;
        run("swapped (A weights, B halo voxels)", 1, p, grid);
      }
      for (int ni = 2; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        // 128-byte swizzle, K-major, 64-element (128 B) rows: SBO = 1024 (8 rows), LBO ignored
        p.a_off = 0; p.a_lbo = 16; p.a_sbo = 1024; p.a_layout = 2; p.a_shift_mode = 0;
        p.b_off = 128 * 1024; p.b_lbo = 16; p.b_sbo = 1024; p.b_layout = 2;
        run("SWIZZLE_128B K-major", 1, p, grid);
      }
      for (int ni = 1; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        // wgrad layout: both MN-major (8 channels contiguous, 8 k-rows per core matrix)
        p.a_off = 0; p.a_lbo = 128; p.a_sbo = 12 * 20 * 16; p.a_layout = 0; p.a_shift_mode = 0; p.a_major = 1;
        p.b_off = 128 * 1024; p.b_lbo = 128; p.b_sbo = 8 * 16 * 16; p.b_layout = 0; p.b_major = 1;
        run("wgrad layout (both MN-major)", 1, p, grid);
      }
    }
  } else {
    for (int grid : {2, 148}) {
      for (int ni = 2; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        p.a_off = 0; p.a_lbo = 8 * 20 * 12 * 16; p.a_sbo = 12 * 16; p.a_layout = 0; p.a_shift_mode = 1;
        p.b_off = 128 * 1024; p.b_lbo = (uint32_t)(n / 2 * 16 * 2 + 1024); p.b_sbo = 128; p.b_layout = 0;
        run("conv layout, 2-CTA pair (M=256)", 2, p, grid);
      }
      for (int ni = 2; ni < 5; ++ni) {
        const int n = ns[ni];
        Params p = {};
        p.reps = reps; p.n = n; p.n_acc = 2;
        p.a_off = 0; p.a_lbo = 16; p.a_sbo = 1024; p.a_layout = 2; p.a_shift_mode = 0;
        p.b_off = 128 * 1024; p.b_lbo = 16; p.b_sbo = 1024; p.b_layout = 2;
        run("SWIZZLE_128B K-major, 2-CTA pair (M=256)", 2, p, grid);
      }
    }
  }
  return 0;
}
