// Strided Conv3D and Conv3DTranspose without padding (DownTransition.down_conv vnet.py:98-99, UpTransition.up_conv
// vnet.py:133-137, and each other's input gradients) as pointwise GEMMs on tcgen05 tensor cores: the default
// 2x2x2 / stride 2 and the anisotropic MRISpineSeg kernels (2,2,4)/(2,2,1), (2,2,2)/(2,2,1) whose windows overlap along
// the last axis (vnet_mri_spine_seg_512_512_12_15k.yml:9-10).  General form (taps = kd*kh*kw, any strides):
//   gather : out[v, co] = bias[co] + sum_{tap, cr} x[s*v + tap, cr] * w[co][cr][tap]; the strided 5-D TMA box uses
//            elementStrides (s_w, s_h) and the start coordinate (s*v0 + tap).
//   scatter: the d and h axes must have stride == kernel (disjoint windows: their taps go to N and the epilogue writes
//            them depth-to-space); the w axis is either disjoint as well ("scatter-w", taps on N) or has stride 1
//            ("gather-w": out[.., w'] = sum_kw x[.., w' - kw] W[kw] - the kw taps join the K loop as shifted boxes,
//            zero-filled out of bounds, so overlapping windows need no atomics).
// The text below describes the default 2x2x2 / stride-2 case.
//
// The windows do not overlap, so both directions are plain GEMMs over 128-voxel tiles (8 w x 16 h of one d-plane of
// the SMALL grid) with bf16 operands and f32 accumulation in TMEM:
//   gather  (mode 0): out[v, co]       = bias[co] + sum_{tap, cr} x[2v + tap, cr] * w[co][cr][tap]     K = 8 * Cred
//       A = the tap-(kd,kh,kw) sub-lattice of the big grid, fetched by ONE strided 5-D TMA box per (tap, 16 channels)
//       (elementStrides 2 along w and h: the space-to-depth never touches memory), B = packed weights.
//   scatter (mode 1): out[2v + tap, co] (+)= bias[co] + sum_cr x[v, cr] * w[cr][co][tap]                K = Cred
//       N = (taps of a group, co): one MMA produces up to 8 output voxels per input voxel; the epilogue writes each
//       tap's 16-byte channel vectors to its depth-to-space position.
// Both are HBM-bound (AI 26-190 FLOP/B): the kernel's job is to stream x once and write out once.
// Warp roles as in conv_k5_umma.cu: w0 TMA producer, w1 MMA issue (uniform control flow, one elected lane),
// w2 TMEM alloc, w4-11 epilogue: two warpgroups take alternate 16-column blocks (bias, optional accumulate, bf16
// round, BN partial sums, 128-bit stores).  Round 2: the accumulate form serialised (tcgen05.ld -> load old -> add ->
// store) per 16 columns with ~4 KB in flight per SM (ncu: DRAM at 20 % of peak); the `old` vectors of a whole item are
// now requested BEFORE the accumulator barrier, so their HBM latency overlaps the MMAs: 190 -> 100 us for the
// down_tr32 input gradient.  MEASURED negative result: 16 epilogue warps instead of 8 made the store-only forms slower
// (scatter 55 -> 63 us, gather 73 -> 94 us: the extra warps spin on the accumulator barrier and take issue slots).
#include <cuda.h>

#include "common.cuh"
#include "conv_k5.cuh"
#include "umma.cuh"

namespace msb {

constexpr int kK2TileW = 8, kK2TileH = 16;
constexpr int kK2ABytes = 2 * kK2TileH * kK2TileW * 16;  // [2 c8][16 h][8 w][8 ch] bf16 = 4 KB
// shared memory of a kernel variant: stages x (A tile + [2 k8][N <= ACCN][8 ch] weights) + barriers + BN statistics
constexpr int k2_smem_bytes(int ew, int st, int accn) { return st * (kK2ABytes + accn * 32) + 1024 + ew * 2 * 256 * 4 + 128; }

struct K2Params {
  int mode;              // 0 gather, 1 scatter
  int n, chunks;         // chunks = x.c / 16
  int nmma;              // MMA N: gather = cpad, scatter = tg * cpad
  int cpad;              // output channels padded to a multiple of 16
  int tg, tap_groups;    // scatter: taps per MMA and number of tap groups (tg * tap_groups = 8); gather: 1, 1
  int cout_real, out_c8;
  int sd, sh, sw;        // small-grid extents
  int bd, bh, bw;        // big-grid extents
  int kdn, khn, kwn;     // kernel extents
  int std, sth, stw;     // strides
  int wmode;             // scatter only: 0 = w taps on N (disjoint windows), 1 = w taps in the K loop (stride 1)
  int ntap_n;            // scatter: taps carried on N (kd*kh[*kw])
  int tiles_w, tiles_h;
  int x_c8_total;        // planes per n in the TMA coordinate space of x
  const void* packed;
  const float* bias;
  msb_tensor out;
  int accumulate;
  int groups;
  double* sums;
  int sums_c;
  unsigned int* sched;   // dynamic tile scheduler counters {next, done} (nullptr = static round-robin)
};

static __device__ unsigned int g_sched_k2[2];
static unsigned int* g_sched_k2_ptr = nullptr;
static int g_dynamic_tiles_k2 = 0;

int msb_set_tile_scheduler_k2s2(int dynamic) {
  if (dynamic && g_sched_k2_ptr == nullptr) {
    MSB_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&g_sched_k2_ptr), g_sched_k2));
    MSB_CUDA_OK(cudaMemset(g_sched_k2_ptr, 0, 2 * sizeof(unsigned int)));
  }
  g_dynamic_tiles_k2 = dynamic ? 1 : 0;
  return MSB_OK;
}

// EW epilogue warps, ST TMA stages, ACCN TMEM columns per accumulator (two accumulators are allocated), ACC = the
// accumulate form.  <8, 8, 256, *>: one CTA per SM, any N <= 256.  <4, 4, 128, false>: the store-only forms with
// N <= 128 run TWO CTAs per SM (256 TMEM columns, 41 KB of shared memory, 256 threads each): the kernel is latency
// bound (ncu: DRAM 21-25 % of peak, SMs 18-35 % busy with one CTA per SM), a second independent TMA -> MMA -> epilogue
// chain per SM doubles what is in flight.
template <int EW, int ST, int ACCN, bool ACC>
__global__ void __launch_bounds__(128 + 32 * EW, ACCN == 128 ? 2 : 1)
    conv_k2s2_kernel(const __grid_constant__ CUtensorMap tmap_x, const K2Params p) {
  constexpr int kK2Stages = ST, kK2EpiWarps = EW, kK2EpiGroups = EW / 4, kK2MaxIt = (ACCN / 16) / (EW / 4);
  constexpr int kK2StageBytes = kK2ABytes + ACCN * 32, kK2Threads = 128 + 32 * EW;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* stage_smem = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stage_smem + kK2Stages * kK2StageBytes);
  // barrier map: [0,S) full  [S,2S) empty  [2S,2S+2) acc_full  [2S+2,2S+4) acc_empty
  constexpr int kFull = 0, kEmpty = kK2Stages, kAccFull = 2 * kK2Stages, kAccEmpty = 2 * kK2Stages + 2;
  constexpr int kSF = 40, kSE = 40 + sched::kDepth;  // tile-scheduler ring; item slots at byte 512 of the barrier KB
  static_assert(2 * kK2Stages + 4 + 1 <= 40, "barrier map overlap");
  volatile int* sched_slots = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kK2Stages + 4);
  float* stat_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [epilogue warps][2][256]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  if (threadIdx.x == 0) {
    for (int i = 0; i < kK2Stages; ++i) { ptx::mbar_init(BAR(kFull + i), 1); ptx::mbar_init(BAR(kEmpty + i), 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(kAccFull + i), 1); ptx::mbar_init(BAR(kAccEmpty + i), kK2EpiWarps); }
    if (sched::kEnabled) sched::init(BAR(kSF), BAR(kSE), 2 + kK2EpiWarps);  // consumers: producer and MMA warps + the epilogue warps
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < kK2EpiWarps * 2 * 256; i += kK2Threads) stat_smem[i] = 0.f;
  __shared__ int tap_off[64];  // scatter: big-grid voxel offset of every N-side tap (kd, kh[, kw])
  if (threadIdx.x < 64) {
    const int tap = threadIdx.x;
    const int kwl = p.wmode == 1 ? 0 : tap % p.kwn;
    const int r2 = p.wmode == 1 ? tap : tap / p.kwn;
    const int kh = r2 % p.khn, kd = r2 / p.khn;
    tap_off[tap] = (kd * p.bh + kh) * p.bw + kwl;
  }
  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmap_x);
  if (warp == 2) ptx::tmem_alloc<2 * ACCN>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after our own TMEM allocation (common.cuh, PDL rules)
  pdl_wait();

  const int tiles_per_n = p.sd * p.tiles_h * p.tiles_w;
  const int num_items = p.n * tiles_per_n * p.tap_groups;
  const int ktaps = p.mode == 0 ? p.kdn * p.khn * p.kwn : (p.wmode == 1 ? p.kwn : 1);  // taps carried by the K loop
  const int kiters = ktaps * p.chunks;
  const uint32_t b_bytes = (uint32_t)p.nmma * 32u;
  // k-th item of this CTA: static round-robin, or from the scheduler warp (umma.cuh, namespace sched)
  const bool dyn = sched::kEnabled && p.sched != nullptr;
  auto get_item = [&](uint32_t k) -> int {
    if (dyn) return sched::next(BAR(kSF), BAR(kSE), sched_slots, k, lane);
    const int it = (int)blockIdx.x + (int)k * (int)gridDim.x;
    return it < num_items ? it : -1;
  };

  if (warp == 3) {
    if (dyn && lane == 0) sched::run(BAR(kSF), BAR(kSE), sched_slots, p.sched, num_items);
  } else if (warp == 0) {
    // ================= TMA producer: one (tap, 16-channel chunk) A tile + its weight block per stage ==========
    const bool leader = ptx::elect_one();
    uint32_t use = 0;
    for (uint32_t ik = 0;; ++ik) {
      const int item = get_item(ik);
      if (item < 0) break;
      const int tile = item / p.tap_groups, tgp = item % p.tap_groups;
      const int n = tile / tiles_per_n;
      int r = tile % tiles_per_n;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; const int d = r / p.tiles_h;
      for (int ki = 0; ki < kiters; ++ki, ++use) {
        const int tap = ki / p.chunks, ck = ki % p.chunks;
        const uint32_t s = use % kK2Stages, ph = (use / kK2Stages) & 1;
        ptx::mbar_wait(BAR(kEmpty + s), ph ^ 1);
        if (leader) {
          uint8_t* st = stage_smem + s * kK2StageBytes;
          ptx::mbar_expect_tx(BAR(kFull + s), kK2ABytes + b_bytes);
          if (p.mode == 0) {
            const int kw = tap % p.kwn, kh = (tap / p.kwn) % p.khn, kd = tap / (p.kwn * p.khn);
            ptx::tma_load_5d(ptx::smem_u32(st), &tmap_x, BAR(kFull + s), 0, p.stw * tw * kK2TileW + kw,
                             p.sth * th * kK2TileH + kh, p.std * d + kd, n * p.x_c8_total + ck * 2);
          } else {  // gather-w: the tile runs over OUTPUT w', tap kw reads x[.., w' - kw] (zero outside the volume)
            ptx::tma_load_4d(ptx::smem_u32(st), &tmap_x, BAR(kFull + s), (tw * kK2TileW - (p.wmode == 1 ? tap : 0)) * 8,
                             th * kK2TileH, d, n * p.x_c8_total + ck * 2);
          }
          const int blk = (p.mode == 0 ? tap : tgp * ktaps + tap) * p.chunks + ck;
          ptx::bulk_load(ptx::smem_u32(st + kK2ABytes), reinterpret_cast<const uint8_t*>(p.packed) + (size_t)blk * b_bytes,
                         b_bytes, BAR(kFull + s));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
    const uint32_t idesc = ptx::make_idesc_bf16(128, p.nmma, 0, 0);
    constexpr uint32_t a_hi = ptx::desc_hi(128u), b_hi = ptx::desc_hi(128u);
    const uint32_t a_lbo16 = (uint32_t)(kK2TileH * kK2TileW * 16) >> 4, b_lbo16 = (uint32_t)p.nmma;  // nmma*16 B >> 4
    uint32_t use = 0, iuse = 0;
    for (;; ++iuse) {
      if (sched::uniform(get_item(iuse)) < 0) break;
      const uint32_t as = iuse & 1, aph = (iuse >> 1) & 1;
      ptx::mbar_wait(BAR(kAccEmpty + as), aph ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_u + as * (uint32_t)ACCN;
      for (int ki = 0; ki < kiters; ++ki, ++use) {
        const uint32_t s = use % kK2Stages, ph = (use / kK2Stages) & 1;
        ptx::mbar_wait(BAR(kFull + s), ph);
        ptx::tc_fence_after();
        const uint32_t st = ptx::smem_u32(stage_smem + s * kK2StageBytes);
        const uint32_t a_lo = ptx::desc_lo(st, a_lbo16), b_lo = ptx::desc_lo(st + kK2ABytes, b_lbo16);
        if (leader) {
          ptx::mma_bf16_split(d_tmem, a_lo, a_hi, b_lo, b_hi, idesc, ki != 0 ? 1u : 0u);
          ptx::mma_commit(BAR(kEmpty + s));
        }
        __syncwarp();
      }
      if (leader) ptx::mma_commit(BAR(kAccFull + as));
      __syncwarp();
    }
  } else if (warp >= 4) {
    // ================= epilogue: the warpgroups take alternate 16-column blocks ================================
    // (a warp reads the TMEM lane quadrant warp % 4).  Everything that does not depend on the item is hoisted and the
    // bf16 conversion is done once.
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    const int64_t Ss = (int64_t)p.sd * p.sh * p.sw;
    const int64_t So = p.mode == 0 ? Ss : (int64_t)p.bd * p.bh * p.bw;  // voxels of the output grid
    float* my_stats = stat_smem + (warp - 4) * 2 * 256;
    const int nblk16 = p.nmma / 16;
    const bool has_bias = p.bias != nullptr;
    const bool bias_vec = has_bias && (reinterpret_cast<uintptr_t>(p.bias) % 16 == 0);
    const bool want_stats = p.sums != nullptr;
    uint32_t iuse = 0;
    for (;; ++iuse) {
      const int item = get_item(iuse);
      if (item < 0) break;
      const int tile = item / p.tap_groups, tgp = item % p.tap_groups;
      const int n = tile / tiles_per_n;
      int r = tile % tiles_per_n;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; const int d = r / p.tiles_h;
      const int h = th * kK2TileH + hh, w = tw * kK2TileW + ww;
      // gather: (d, h, w) is an output voxel of the small grid; scatter: an input voxel (scatter-w) or, in gather-w
      // mode, w already is the OUTPUT coordinate along w
      const bool ok = h < p.sh && w < ((p.mode == 1 && p.wmode == 1) ? p.bw : p.sw);
      const uint32_t as = iuse & 1, aph = (iuse >> 1) & 1;
      const int64_t v_small = ((int64_t)d * p.sh + h) * p.sw + w;
      const int64_t v_big0 = ((int64_t)(p.std * d) * p.bh + p.sth * h) * p.bw + (p.wmode == 1 ? w : p.stw * w);
      __nv_bfloat16* out_n = reinterpret_cast<__nv_bfloat16*>(p.out.ptr) + (int64_t)n * p.out.n_stride;
      // destination of (column block cb, channel plane k): nullptr when the lane / tap / plane is dead
      auto dest = [&](int cb, int k, int& co0) -> __nv_bfloat16* {
        const int col = cb * 16;
        const int t = col / p.cpad;
        co0 = col - t * p.cpad;
        int64_t v = v_small;
        bool tap_ok = true;
        if (p.mode != 0) {
          const int tap = tgp * p.tg + t;  // N-side tap: (kd, kh[, kw]) row-major; voxel offsets precomputed in smem
          tap_ok = tap < p.ntap_n;
          v = v_big0 + tap_off[tap_ok ? tap : 0];
        }
        const int c8 = (co0 >> 3) + k;
        return (ok && tap_ok && c8 < p.out_c8) ? out_n + ((int64_t)c8 * So + v) * 8 : nullptr;
      };
      // accumulate: request every `old` vector of this item now - the loads fly while the MMAs of the item run
      // kPF column blocks are in flight per thread: all of an item's for the one-CTA variant, a rotating window of
      // kPF for the two-CTAs-per-SM variant (128 registers per thread)
      constexpr int kPF = ACC ? (ACCN == 128 ? 3 : kK2MaxIt) : 1;
      uint4 oldv[kPF][2];
      auto fetch_old = [&](int it, uint4 (&slot)[2]) {
        const int cb = wg + kK2EpiGroups * it;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          slot[k] = make_uint4(0u, 0u, 0u, 0u);
          if (cb < nblk16) {
            int co0;
            const __nv_bfloat16* src = dest(cb, k, co0);
            if (src != nullptr) slot[k] = *reinterpret_cast<const uint4*>(src);
          }
        }
      };
      if (ACC) {
#pragma unroll
        for (int it = 0; it < kPF; ++it) fetch_old(it, oldv[it]);
      }
      ptx::mbar_wait(BAR(kAccFull + as), aph);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + as * (uint32_t)ACCN + ((uint32_t)(q * 32) << 16);
#pragma unroll
      for (int it = 0; it < kK2MaxIt; ++it) {
        const int cb = wg + kK2EpiGroups * it;
        if (cb >= nblk16) break;
        float acc[16];
        ptx::tmem_ld16(t_base + cb * 16, acc);
        int co0 = 0;
        __nv_bfloat16* dst0 = dest(cb, 0, co0);
        int co0b;
        __nv_bfloat16* dst1 = dest(cb, 1, co0b);
        if (has_bias) {
          if (bias_vec && co0 + 16 <= p.cout_real) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + co0) + i);
              acc[4 * i] += b4.x; acc[4 * i + 1] += b4.y; acc[4 * i + 2] += b4.z; acc[4 * i + 3] += b4.w;
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (co0 + i < p.cout_real) acc[i] += __ldg(p.bias + co0 + i);
          }
        }
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          __nv_bfloat16* dst = k == 0 ? dst0 : dst1;
          uint32_t pk[4] = {0u, 0u, 0u, 0u};
          if (dst != nullptr) {
            if (ACC) {
              const uint4 ov = oldv[ACC ? it % kPF : 0][k];
              const uint32_t o[4] = {ov.x, ov.y, ov.z, ov.w};
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[k * 8 + 2 * i] += __uint_as_float(o[i] << 16);
                acc[k * 8 + 2 * i + 1] += __uint_as_float(o[i] & 0xffff0000u);
              }
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const __nv_bfloat162 hpair = __floats2bfloat162_rn(acc[k * 8 + 2 * i], acc[k * 8 + 2 * i + 1]);
              pk[i] = *reinterpret_cast<const uint32_t*>(&hpair);
            }
            *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
          }
          if (want_stats) {  // statistics of the ROUNDED values (what the next kernel reads); dead lanes contribute 0
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              acc[k * 8 + 2 * i] = __uint_as_float(pk[i] << 16);
              acc[k * 8 + 2 * i + 1] = __uint_as_float(pk[i] & 0xffff0000u);
            }
          }
        }
        if (ACC && kPF < kK2MaxIt && it + kPF < kK2MaxIt) fetch_old(it + kPF, oldv[it % kPF]);  // refill the window
        if (want_stats) {
          float sq[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) sq[i] = acc[i] * acc[i];
          const float s1 = warp_reduce16(acc, lane);
          const float s2 = warp_reduce16(sq, lane);
          if ((lane & 1) == 0) {
            my_stats[co0 + (lane >> 1)] += s1;
            my_stats[256 + co0 + (lane >> 1)] += s2;
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(BAR(kAccEmpty + as));
      if (want_stats && p.groups > 1) {
        __syncwarp();
        for (int i = lane; i < 2 * 256; i += 32) {
          const int stat = i / 256, c = i % 256;
          if (c < p.sums_c && my_stats[i] != 0.f)
            atomicAdd(&p.sums[((int64_t)stat * p.groups + n) * p.sums_c + c], (double)my_stats[i]);
          my_stats[i] = 0.f;
        }
        __syncwarp();
      }
    }
    if (want_stats && p.groups == 1) {
      __syncwarp();
      for (int i = lane; i < 2 * 256; i += 32) {
        const int stat = i / 256, c = i % 256;
        if (c < p.sums_c && my_stats[i] != 0.f) atomicAdd(&p.sums[(int64_t)stat * p.sums_c + c], (double)my_stats[i]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<2 * ACCN>(tmem_base);
  }
}

// gather : packed[tap][chunk][k8][co < cpad][j]                 = w[co][cr][tap]   (w = [c_out][c_red][taps])
// scatter: packed[tgp][ktap][chunk][k8][t*cpad + co][j]         = w[cr][co][tap]   (w = [c_red][c_out][taps])
//          tap = (N-side tap tgp*tg + t, K-side tap ktap) -> (kd, kh, kw); cr = chunk*16 + k8*8 + j
struct K2Geom {
  int kdn, khn, kwn;   // kernel
  int wmode;           // scatter: 1 = kw taps in the K loop
  int tg, tap_groups;  // scatter: taps per MMA / number of tap groups
  int ntap_n, ktaps;   // scatter: taps on N in total / taps in the K loop
};

__global__ void __launch_bounds__(256) pack_k2s2_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                        int c_red, int c_out, int mode, int c_red_pad, int cpad,
                                                        K2Geom g) {
  pdl_wait();
  pdl_trigger();
  const int chunks = c_red_pad / 16;
  const int taps = g.kdn * g.khn * g.kwn;
  const int64_t total = mode == 0 ? (int64_t)taps * c_red_pad * cpad
                                  : (int64_t)g.tap_groups * g.ktaps * c_red_pad * g.tg * cpad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    int64_t r = i >> 3;
    int tap, co, k8, ck;
    bool tap_ok = true;
    if (mode == 0) {
      co = (int)(r % cpad); r /= cpad;
      k8 = (int)(r & 1); r >>= 1;
      ck = (int)(r % chunks); tap = (int)(r / chunks);
    } else {
      const int nm = g.tg * cpad;
      const int col = (int)(r % nm); r /= nm;
      k8 = (int)(r & 1); r >>= 1;
      ck = (int)(r % chunks); r /= chunks;
      const int ktap = (int)(r % g.ktaps);
      const int tgp = (int)(r / g.ktaps);
      const int ntap = tgp * g.tg + col / cpad;  // N-side tap: (kd, kh[, kw])
      co = col % cpad;
      tap_ok = ntap < g.ntap_n;
      tap = g.wmode == 1 ? ntap * g.kwn + ktap : ntap;  // full tap index (kd*khn + kh)*kwn + kw
    }
    const int cr = ck * 16 + k8 * 8 + j;
    float v = 0.f;
    if (tap_ok && cr < c_red && co < c_out)
      v = mode == 0 ? __ldg(w + ((int64_t)co * c_red + cr) * taps + tap) : __ldg(w + ((int64_t)cr * c_out + co) * taps + tap);
    packed[i] = __float2bfloat16_rn(v);
  }
}

static inline int k2_pad16(int c) { return (c + 15) / 16 * 16; }

// scatter geometry for an output of cpad channels; returns false when the shape is not supported
static bool k2_geom(int mode, msb_dim3 kernel, msb_dim3 stride, int cpad, K2Geom* g) {
  g->kdn = kernel.d; g->khn = kernel.h; g->kwn = kernel.w;
  g->wmode = 0; g->tg = 1; g->tap_groups = 1; g->ntap_n = 1; g->ktaps = kernel.d * kernel.h * kernel.w;
  if (kernel.d < 1 || kernel.h < 1 || kernel.w < 1 || stride.d < 1 || stride.h < 1 || stride.w < 1) return false;
  if (kernel.d * kernel.h * kernel.w > 64 || stride.w > 4 || stride.h > 4) return false;
  if (mode == 0) return true;
  if (kernel.d != stride.d || kernel.h != stride.h) return false;  // d / h windows must be disjoint
  if (kernel.w == stride.w) g->wmode = 0;
  else if (stride.w == 1) g->wmode = 1;
  else return false;
  g->ntap_n = kernel.d * kernel.h * (g->wmode == 1 ? 1 : kernel.w);
  g->ktaps = g->wmode == 1 ? kernel.w : 1;
  int tg = 1;
  // N = tg * cpad <= 128: every scatter then fits the 128-column accumulators of the two-CTAs-per-SM variant (wider
  // outputs fall back to one tap per MMA)
  while (tg * 2 <= g->ntap_n && tg * 2 * cpad <= 128) tg *= 2;
  g->tg = tg;
  g->tap_groups = (g->ntap_n + tg - 1) / tg;
  return true;
}

static int launch_k2s2(int mode, const msb_tensor& x, const void* packed, const float* bias, int cout,
                       const msb_tensor& out, int n, msb_dim3 big, msb_dim3 kernel, msb_dim3 stride, int accumulate,
                       int groups, double* sums, cudaStream_t st) {
  K2Params p;
  K2Geom g;
  p.cpad = k2_pad16(out.c);
  if (!k2_geom(mode, kernel, stride, p.cpad, &g)) {
    set_error("strided tensor-core conv: unsupported kernel / stride combination");
    return MSB_ERR_UNSUPPORTED;
  }
  const msb_dim3 sd = {(big.d - kernel.d) / stride.d + 1, (big.h - kernel.h) / stride.h + 1,
                       (big.w - kernel.w) / stride.w + 1};
  p.mode = mode; p.n = n; p.chunks = x.c / 16;
  p.tg = g.tg; p.tap_groups = g.tap_groups; p.ntap_n = g.ntap_n; p.wmode = g.wmode;
  p.nmma = mode == 0 ? p.cpad : p.tg * p.cpad;
  p.cout_real = cout; p.out_c8 = out.c / 8;
  p.sd = sd.d; p.sh = sd.h; p.sw = sd.w;
  p.bd = big.d; p.bh = big.h; p.bw = big.w;
  p.kdn = kernel.d; p.khn = kernel.h; p.kwn = kernel.w;
  p.std = stride.d; p.sth = stride.h; p.stw = stride.w;
  const int tile_w_extent = (mode == 1 && p.wmode == 1) ? big.w : sd.w;
  p.tiles_w = (tile_w_extent + kK2TileW - 1) / kK2TileW;
  p.tiles_h = (sd.h + kK2TileH - 1) / kK2TileH;
  p.packed = packed; p.bias = bias; p.out = out; p.accumulate = accumulate;
  p.groups = groups; p.sums = sums; p.sums_c = out.c;
  p.sched = g_dynamic_tiles_k2 ? g_sched_k2_ptr : nullptr;
  CUtensorMap tmap;
  int rc;
  if (mode == 0) {
    const int64_t Sb = (int64_t)big.d * big.h * big.w;
    p.x_c8_total = (int)(x.n_stride / (Sb * 8));
    if ((rc = make_b8_tmap_s2(&tmap, x, n, big, kK2TileW, kK2TileH, 2, stride.w, stride.h))) return rc;
  } else {
    const int64_t Ss = (int64_t)sd.d * sd.h * sd.w;
    p.x_c8_total = (int)(x.n_stride / (Ss * 8));
    if ((rc = make_b8_tmap(&tmap, x, n, sd, kK2TileW, kK2TileH, 1, 2))) return rc;
  }
  const int items = n * sd.d * p.tiles_h * p.tiles_w * p.tap_groups;
  static bool attr_set = false;
  if (!attr_set) {
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k2s2_kernel<8, 8, 256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     k2_smem_bytes(8, 8, 256)));
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k2s2_kernel<8, 8, 256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     k2_smem_bytes(8, 8, 256)));
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k2s2_kernel<4, 4, 128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     k2_smem_bytes(4, 4, 128)));
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k2s2_kernel<4, 4, 128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     k2_smem_bytes(4, 4, 128)));
    attr_set = true;
  }
  if (p.nmma <= 128 && !(g_debug_flags[6] & 64) && !(accumulate && (g_debug_flags[6] & 128))) {
    // narrow N: two CTAs per SM (the accumulate form prefetches a rotating window of 3 column blocks instead of all 8)
    const int grid = items < 2 * kNumSMs ? items : 2 * kNumSMs;
    if (accumulate)
      MSB_LAUNCH_PDL((conv_k2s2_kernel<4, 4, 128, true>), dim3(grid), dim3(128 + 32 * 4), k2_smem_bytes(4, 4, 128), st, tmap, p);
    else
      MSB_LAUNCH_PDL((conv_k2s2_kernel<4, 4, 128, false>), dim3(grid), dim3(128 + 32 * 4), k2_smem_bytes(4, 4, 128), st, tmap, p);
  } else {
    const int grid = items < kNumSMs ? items : kNumSMs;
    if (accumulate)
      MSB_LAUNCH_PDL((conv_k2s2_kernel<8, 8, 256, true>), dim3(grid), dim3(128 + 32 * 8), k2_smem_bytes(8, 8, 256), st, tmap, p);
    else
      MSB_LAUNCH_PDL((conv_k2s2_kernel<8, 8, 256, false>), dim3(grid), dim3(128 + 32 * 8), k2_smem_bytes(8, 8, 256), st, tmap, p);
  }
  return MSB_OK;
}

}  // namespace msb

using namespace msb;

extern "C" {

static const msb_dim3 kTwo = {2, 2, 2};

size_t msb_conv_tc_packed_bytes(int c_red_pad, int c_out_pad, msb_dim3 kernel, msb_dim3 stride, int mode) {
  K2Geom g;
  if (!k2_geom(mode, kernel, stride, c_out_pad, &g)) return 0;
  const size_t taps = (size_t)kernel.d * kernel.h * kernel.w;
  if (mode == 0) return taps * c_red_pad * c_out_pad * sizeof(__nv_bfloat16);
  return (size_t)g.tap_groups * g.ktaps * c_red_pad * g.tg * c_out_pad * sizeof(__nv_bfloat16);
}

int msb_conv_tc_pack(const float* w, void* packed, int c_red, int c_out, int mode, int c_red_pad, int c_out_pad,
                     msb_dim3 kernel, msb_dim3 stride, void* stream) {
  MSB_REQUIRE(w && packed && c_red > 0 && c_out > 0 && (mode == 0 || mode == 1), "msb_conv_tc_pack: bad arguments");
  MSB_REQUIRE(c_red_pad % 16 == 0 && c_out_pad % 16 == 0 && c_red_pad >= c_red && c_out_pad >= c_out && c_out_pad <= 256,
              "msb_conv_tc_pack: padded channel counts must be multiples of 16 covering the real ones (out <= 256)");
  K2Geom g;
  MSB_REQUIRE(k2_geom(mode, kernel, stride, c_out_pad, &g), "msb_conv_tc_pack: unsupported kernel / stride combination");
  const int64_t total = (int64_t)(msb_conv_tc_packed_bytes(c_red_pad, c_out_pad, kernel, stride, mode) / 2);
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  MSB_LAUNCH_PDL(pack_k2s2_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), w,
                 reinterpret_cast<__nv_bfloat16*>(packed), c_red, c_out, mode, c_red_pad, c_out_pad, g);
  return MSB_OK;
}

static int k2s2_check(const char* who, const msb_tensor& x, const msb_tensor& out, const void* packed, int cout, int n,
                      msb_dim3 big, msb_dim3 kernel, msb_dim3 stride, int groups) {
  MSB_REQUIRE(view_ok(x) && view_ok(out) && x.dtype == MSB_BF16 && out.dtype == MSB_BF16 && packed && n > 0,
              "%s: bf16 B8 views required", who);
  MSB_REQUIRE(big.d >= kernel.d && big.h >= kernel.h && big.w >= kernel.w && kernel.d > 0 && stride.d > 0,
              "%s: the large grid must be at least one window in every direction", who);
  MSB_REQUIRE((big.d - kernel.d) % stride.d == 0 && (big.h - kernel.h) % stride.h == 0 && (big.w - kernel.w) % stride.w == 0,
              "%s: (big - kernel) must be a multiple of the stride in every direction (the transposed conv's output "
              "extent (small-1)*stride+kernel)", who);
  MSB_REQUIRE(x.c % 16 == 0 && out.c <= 256 && cout > 0 && cout <= out.c, "%s: x.c must be a multiple of 16, out.c <= 256", who);
  MSB_REQUIRE(groups == 1 || groups == n, "%s: groups must be 1 or n", who);
  return MSB_OK;
}

int msb_conv_tc_gather(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                       msb_dim3 big_dims, msb_dim3 kernel, msb_dim3 stride, int groups, double* sums, void* stream) {
  int rc = k2s2_check("msb_conv_tc_gather", x, out, packed, cout, n, big_dims, kernel, stride, groups);
  if (rc) return rc;
  return launch_k2s2(0, x, packed, bias, cout, out, n, big_dims, kernel, stride, 0, groups, sums, as_stream(stream));
}

int msb_conv_tc_scatter(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                        msb_dim3 big_dims, msb_dim3 kernel, msb_dim3 stride, int accumulate, int groups, double* sums,
                        void* stream) {
  int rc = k2s2_check("msb_conv_tc_scatter", x, out, packed, cout, n, big_dims, kernel, stride, groups);
  if (rc) return rc;
  return launch_k2s2(1, x, packed, bias, cout, out, n, big_dims, kernel, stride, accumulate, groups, sums,
                     as_stream(stream));
}

// ---- the 2x2x2 / stride-2 entry points (kept: thin wrappers) ------------------------------------------------
size_t msb_conv_k2s2_packed_bytes(int c_red_pad, int c_out_pad) {
  return (size_t)8 * c_red_pad * c_out_pad * sizeof(__nv_bfloat16);
}

int msb_conv_k2s2_pack(const float* w, void* packed, int c_red, int c_out, int mode, int c_red_pad, int c_out_pad,
                       void* stream) {
  return msb_conv_tc_pack(w, packed, c_red, c_out, mode, c_red_pad, c_out_pad, kTwo, kTwo, stream);
}

int msb_conv_k2s2_gather(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                         msb_dim3 big_dims, int groups, double* sums, void* stream) {
  return msb_conv_tc_gather(x, packed, bias, cout, out, n, big_dims, kTwo, kTwo, groups, sums, stream);
}

int msb_conv_k2s2_scatter(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                          msb_dim3 big_dims, int accumulate, int groups, double* sums, void* stream) {
  return msb_conv_tc_scatter(x, packed, bias, cout, out, n, big_dims, kTwo, kTwo, accumulate, groups, sums, stream);
}

}  // extern "C"
