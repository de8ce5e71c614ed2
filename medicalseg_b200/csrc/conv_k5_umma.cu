// 5x5x5 Conv3D (pad 2, stride 1) as an implicit GEMM on tcgen05 tensor cores — LUConv.conv1 (vnet.py:36),
// OutputTransition.conv1 (vnet.py:165-166) and, with flipped/transposed weights, their input gradients.
//
// Data layout / mapping:
//   * activations: B8 bf16 [N][C/8][D][H][W][8]; ONE 4-D TMA box per 16 input channels fetches the haloed tile
//     [2 c8][TD+4][20][12][8ch] (zero-filled outside the volume = the conv padding) into shared memory.  That
//     image IS the canonical SWIZZLE_NONE K-major UMMA operand: 8 consecutive w voxels x 16 B form a core matrix,
//     so every one of the 125 taps is the same buffer viewed through a shifted 16-byte-aligned start address.
//     The tile is fetched once and reused by 125 x TD MMAs (no im2col, no per-tap reload).
//   * weights: packed bf16 [Cin/16][125][2][Cout][8] streamed by cp.async.bulk in 5-tap stages.
//   * M = 128 output voxels (8 w x 16 h) of one d-plane, N = Cout, K = 16 channels per MMA; TD planes share a
//     weight stage; accumulators (TD x N f32 columns) are double-buffered in TMEM so the epilogue of item i
//     overlaps the MMAs of item i+1.
//   * warp roles: w0 halo TMA, w1 MMA issue (one elected lane), w2 TMEM alloc + weight TMA, w4-7 epilogue.
#include <cuda.h>

#include "common.cuh"
#include "conv_k5.cuh"
#include "umma.cuh"

namespace msb {

constexpr int kTileW = 8, kTileH = 16;
constexpr int kHaloW = kTileW + 4, kHaloH = kTileH + 4;

// NPAD: padded output channels; TD: output d-planes per work item; J: d-planes stacked along N in one MMA
// (N_mma = J * NPAD, TD % J == 0); ACC_SETS: accumulator sets in TMEM (2 = epilogue overlaps the next item).
template <int NPAD, int TD, int J = 1, int ACC_SETS = 2>
struct FwdCfg {
  static constexpr int kHaloPlaneBytes = (TD + 4) * kHaloH * kHaloW * 16;  // one c8 plane of the haloed tile
  static constexpr int kHaloBytes = 2 * kHaloPlaneBytes;                   // 16 input channels
  // one weight stage = one (kh, kw) column: [k8 (2)][block (5)][co (NPAD)][8 ci]; the blocks hold the 5 kd taps in
  // REVERSED order (W4..W0), so the operand for input plane kd' and the stacked output planes j = j_lo..j_hi (tap
  // kd = kd' - j) is the same image viewed from block 4 - kd' + j_lo, nblk = j_hi - j_lo + 1 blocks long.
  static constexpr int kNB = 5;
  static constexpr int kBlockBytes = NPAD * 16;
  static constexpr int kWK8Bytes = kNB * kBlockBytes;
  static constexpr int kWStageBytes = 2 * kWK8Bytes;
  static constexpr int kWLoadBytes = 5 * kBlockBytes;                      // bytes TMA writes per (stage, k8)
  // as many weight stages as fit (<= 8): the stage ring has to cover the L2 latency of cp.async.bulk
  static constexpr int kFixedBytes = 2 * kHaloBytes + 1024 /*barriers*/ + 8 * 2 * NPAD * 4 /*stats*/ + 128 /*alignment slack*/;
  static constexpr int kWStagesFit = (227 * 1024 - kFixedBytes) / kWStageBytes;
  static constexpr int kWStages = kWStagesFit > 8 ? 8 : kWStagesFit;
  static constexpr int kAccCols = TD * NPAD;
  static constexpr int kNMma = J * NPAD;
  static constexpr int kColsNeeded = ACC_SETS * kAccCols;
  static constexpr int kTmemCols = (kColsNeeded <= 32) ? 32 : (kColsNeeded <= 64) ? 64 : (kColsNeeded <= 128) ? 128
                                 : (kColsNeeded <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kFixedBytes + kWStages * kWStageBytes;
  static_assert(kWStages >= 2, "weight stages do not fit");
  static_assert(TD % J == 0 && (J < 5 ? J : 5) * NPAD <= 256, "bad plane stacking (an MMA covers <= 5 stacked planes)");
  static_assert(kColsNeeded <= 512, "TMEM overflow");
  static_assert(kSmemBytes <= 227 * 1024, "shared memory overflow");
};

// issue order of the J+4 input planes of a stacked group: the overwriting ("init") planes first (see the MMA loop)
template <int J>
__host__ __device__ constexpr int kdp_order(int i) {
  static_assert(J <= 10, "two init MMAs cover at most 10 stacked planes");
  constexpr int kA = J - 1 < 4 ? J - 1 : 4;
  constexpr int kB = 9;
  constexpr int n_init = J > 5 ? 2 : 1;
  if (i == 0) return kA;
  if (n_init == 2 && i == 1) return kB;
  int idx = i - n_init;
  for (int k = 0; k < J + 4; ++k) {
    if (k == kA || (n_init == 2 && k == kB)) continue;
    if (idx == 0) return k;
    --idx;
  }
  return -1;
}

struct FwdParams {
  int n, cin_pad, cout_real, out_c8;     // out_c8: 8-channel planes actually stored
  int d, h, w;
  int tiles_w, tiles_h, dblocks;
  int x_c8_total;                        // planes per n in the TMA coordinate space of x
  const void* packed;
  const float* bias;
  msb_tensor out;
  int accumulate;
  const float* ch_scale;
  const float* ep_scale;   // evaluation epilogue (all three per output channel, length >= out.c) or nullptr
  const float* ep_shift;
  const float* ep_alpha;
  const float* ep_alpha2;  // PReLU applied after the residual add (required with ep_res)
  msb_tensor ep_res;       // optional residual (bf16 B8 view of the output's shape)
  int ep_has_res;
  int groups;
  double* sums;
  int sums_c;                            // channel count of the sums array
  int dbg_swap;
  int nsplit;                            // output-channel slices per tile (fills the SMs on small volumes)
  int kw_taps;                           // 5 = full 5x5x5 kernel; 1 = 5x5x1 (kd,kh) kernel of the w-folded convs
  int out_f32;                           // store f32 (B8 f32 view) instead of bf16; no accumulate, no BN sums
  long long* prof;                       // bring-up: per-CTA clocks the MMA warp spent {total, w_full, halo_full, acc_empty}
  float* ws;                             // split-K: f32 partial sums [n][out_c8][D*H*W][8], zero on entry
  int ksplit, chunks_per_split;          // split-K: item = tile * ksplit + slice; slice covers chunks_per_split 16-channel chunks
  unsigned int* sched;                   // dynamic tile scheduler counters {next, done} (nullptr = static round-robin)
};

static __device__ unsigned int g_sched_fwd[2];
static __device__ unsigned int g_sched_wg[2];
int g_dynamic_tiles = 0;

// f32 vector reduction into global memory (sm_90+): 4 consecutive floats per instruction
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

// SPLITK: small volumes (fewer tiles than SMs).  The reduction over the input channels is split across CTAs (item =
// (tile, slice of 16-channel chunks)): every CTA streams only ITS slice of the weight set (the layer is otherwise
// bound by pulling the whole 4-16 MB weight set through every SM) and adds its f32 partial tile into p.ws with vector
// stores into the slice's PRIVATE copy of the output (ws = [ksplit][n][c8][S][8] f32); splitk_finalize_kernel adds the
// copies in slice order (deterministic - round 1 met in one copy through red.global atomics) and applies bias /
// accumulate / rounding / BN sums.
constexpr int kFwdThreads = 384;  // w0 halo TMA, w1 MMA, w2 TMEM alloc + weight TMA, w3 idle, w4-11 epilogue

template <int NPAD, int TD, int J, int ACC_SETS, int NS, bool SPLITK = false>
__global__ void __launch_bounds__(kFwdThreads, 1)
    conv_k5_fwd_kernel(const __grid_constant__ CUtensorMap tmap_x, const FwdParams p) {
  using Cfg = FwdCfg<NPAD, TD, J, ACC_SETS>;
  static_assert(NS == 1 || J == 1, "channel slicing only without plane stacking");
  static_assert(!SPLITK || (NS == 1 && J == 1), "split-K works on whole tiles");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* halo_smem = smem;                                   // [2][kHaloBytes]
  uint8_t* w_smem = halo_smem + 2 * Cfg::kHaloBytes;           // [kWStages][kWStageBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(w_smem + Cfg::kWStages * Cfg::kWStageBytes);
  // barrier map: [0,2) halo_full  [2,4) halo_empty  [4,6) acc_full  [6,8) acc_empty  [8,8+S) w_full  [8+S,8+2S) w_empty
  constexpr int kWF = 8, kWE = 8 + Cfg::kWStages;
  constexpr int kSF = 40, kSE = 40 + sched::kDepth;  // tile-scheduler ring (full / empty), item slots at byte 512
  static_assert(8 + 2 * Cfg::kWStages + 1 <= 40, "barrier map overlap");
  volatile int* sched_slots = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8 + 2 * Cfg::kWStages);
  float* stat_smem = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + 1024);  // [8 warps][2][NPAD]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(0 + i), 1); ptx::mbar_init(BAR(2 + i), 1); }
    for (int i = 0; i < Cfg::kWStages; ++i) { ptx::mbar_init(BAR(kWF + i), 1); ptx::mbar_init(BAR(kWE + i), 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(4 + i), 1); ptx::mbar_init(BAR(6 + i), 8); }
    if (sched::kEnabled) sched::init(BAR(kSF), BAR(kSE), 11);  // consumers: halo, weight and MMA warps + 8 epilogue warps
    ptx::fence_mbar_init();
  }
  for (int i = threadIdx.x; i < 8 * 2 * NPAD; i += kFwdThreads) stat_smem[i] = 0.f;
  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&tmap_x);
  if (warp == 2) ptx::tmem_alloc<Cfg::kTmemCols>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_trigger();  // after our own TMEM allocation (common.cuh, PDL rules)
  pdl_wait();

  const int chunks = p.cin_pad / 16;
  const int items_per_n = p.dblocks * p.tiles_h * p.tiles_w;
  const int kdiv = SPLITK ? p.ksplit : NS;             // items per tile
  const int num_items = p.n * items_per_n * kdiv;      // item = tile * NS + channel slice | tile * ksplit + K slice
  constexpr int n_slice = Cfg::kNMma / NS;             // MMA N per item
  auto chunk_range = [&](int item, int& ck0, int& ck1) {
    if (SPLITK) {
      ck0 = (item % kdiv) * p.chunks_per_split;
      ck1 = min(chunks, ck0 + p.chunks_per_split);
    } else {
      ck0 = 0; ck1 = chunks;
    }
  };

  // k-th item of this CTA: static round-robin, or handed out by the scheduler warp (see umma.cuh, namespace sched)
  const bool dyn = sched::kEnabled && p.sched != nullptr;
  auto get_item = [&](uint32_t k) -> int {
    if (dyn) return sched::next(BAR(kSF), BAR(kSE), sched_slots, k, lane);
    const int it = (int)blockIdx.x + (int)k * (int)gridDim.x;
    return it < num_items ? it : -1;
  };

  if (warp == 3) {
    if (dyn && lane == 0) sched::run(BAR(kSF), BAR(kSE), sched_slots, p.sched, num_items);
  } else if (warp == 0) {
    // ================= halo TMA producer (whole warp runs the uniform loop, one elected lane issues) ==========
    const bool leader = ptx::elect_one();
    uint32_t use = 0;
    for (uint32_t ik = 0;; ++ik) {
      const int item = get_item(ik);
      if (item < 0) break;
      const int tile = item / kdiv;
      const int n = tile / items_per_n;
      int r = tile % items_per_n;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; const int db = r / p.tiles_h;
      int ck0, ck1;
      chunk_range(item, ck0, ck1);
      for (int ck = ck0; ck < ck1; ++ck, ++use) {
        const uint32_t b = use & 1, ph = (use >> 1) & 1;
        ptx::mbar_wait(BAR(2 + b), ph ^ 1);
        if (leader) {
          ptx::mbar_expect_tx(BAR(0 + b), Cfg::kHaloBytes);
          ptx::tma_load_4d(ptx::smem_u32(halo_smem + b * Cfg::kHaloBytes), &tmap_x, BAR(0 + b),
                           (tw * kTileW - 2) * 8, th * kTileH - 2, db * TD - 2, n * p.x_c8_total + ck * 2);
        }
        __syncwarp();
      }
    }
  } else if (warp == 2) {
    // ================= weight producer =================
    const bool leader = ptx::elect_one();
    uint32_t use = 0;
    const int stages_per_chunk = 5 * p.kw_taps;
    for (uint32_t ik = 0;; ++ik) {
      const int item = get_item(ik);
      if (item < 0) break;
      int ck0, ck1;
      chunk_range(item, ck0, ck1);
      for (int ck = ck0; ck < ck1; ++ck) {
        const uint8_t* src =
            reinterpret_cast<const uint8_t*>(p.packed) + (size_t)ck * stages_per_chunk * 2 * Cfg::kWLoadBytes;
        for (int st = 0; st < stages_per_chunk; ++st, ++use) {
          const uint32_t s = use % Cfg::kWStages, ph = (use / Cfg::kWStages) & 1;
          ptx::mbar_wait(BAR(kWE + s), ph ^ 1);
          if (leader) {
            ptx::mbar_expect_tx(BAR(kWF + s), 2 * Cfg::kWLoadBytes);
#pragma unroll
            for (int k8 = 0; k8 < 2; ++k8)
              ptx::bulk_load(ptx::smem_u32(w_smem + s * Cfg::kWStageBytes + k8 * Cfg::kWK8Bytes),
                             src + (size_t)(st * 2 + k8) * Cfg::kWLoadBytes, Cfg::kWLoadBytes, BAR(kWF + s));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================= MMA issuer =================
    // The WHOLE warp runs the (warp-uniform) control flow and descriptor arithmetic so that everything lives in
    // uniform registers; only the tcgen05 instructions themselves are predicated on one lane.  (Running the loop
    // under `if (lane == 0)` makes every operand a per-thread value and costs ~11 SASS instructions per MMA -
    // measured 102 clk per 128x128x16 MMA instead of the 64-clk tensor-core floor, tools/mma_probe.cu.)
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);  // REDUX result = provably warp-uniform
    constexpr uint32_t a_hi = ptx::desc_hi((uint32_t)(kHaloW * 16)), b_hi = ptx::desc_hi(128u);
    constexpr uint32_t a_lbo16 = (uint32_t)Cfg::kHaloPlaneBytes >> 4, b_lbo16 = (uint32_t)Cfg::kWK8Bytes >> 4;
    uint32_t huse = 0, wuse = 0, iuse = 0;
    const int stages_per_chunk = 5 * p.kw_taps;
    const bool prof = p.prof != nullptr;
    long long t_w = 0, t_h = 0, t_a = 0, t_begin = prof ? clock64() : 0, tq = 0;
    for (;; ++iuse) {
      const int item = sched::uniform(get_item(iuse));
      if (item < 0) break;
      const uint32_t as = iuse % ACC_SETS, aph = (iuse / ACC_SETS) & 1;
      if (prof) tq = clock64();
      ptx::mbar_wait(BAR(6 + as), aph ^ 1);
      if (prof) t_a += clock64() - tq;
      ptx::tc_fence_after();
      const uint32_t d_base = tmem_u + as * Cfg::kAccCols;
      const uint32_t b_slice = SPLITK ? 0u : (uint32_t)((item % NS) * n_slice);  // first weight row (16 B each) of this slice
      int ck0, ck1;
      chunk_range(item, ck0, ck1);
      for (int ck = ck0; ck < ck1; ++ck, ++huse) {
        const uint32_t hb = huse & 1, hph = (huse >> 1) & 1;
        if (prof) tq = clock64();
        ptx::mbar_wait(BAR(0 + hb), hph);
        if (prof) t_h += clock64() - tq;
        const uint32_t a_lo0 = ptx::desc_lo(ptx::smem_u32(halo_smem + hb * Cfg::kHaloBytes), a_lbo16);
        for (int st = 0; st < stages_per_chunk; ++st, ++wuse) {  // st = kh*kw_taps + kw
          const uint32_t s = wuse % Cfg::kWStages, wph = (wuse / Cfg::kWStages) & 1;
          if (prof) tq = clock64();
          ptx::mbar_wait(BAR(kWF + s), wph);
          if (prof) t_w += clock64() - tq;
          ptx::tc_fence_after();
          const uint32_t b_lo0 = ptx::desc_lo(ptx::smem_u32(w_smem + s * Cfg::kWStageBytes), b_lbo16) + b_slice;
          const uint32_t hw_off =  // 16-byte units; the single w tap of the 5x5x1 kernel is the centre one
              p.kw_taps == 5 ? (uint32_t)((st / 5) * kHaloW + (st % 5)) : (uint32_t)(st * kHaloW + 2);
          const uint32_t a_lo1 = a_lo0 + hw_off;
          const uint32_t first = ((ck - ck0) | st) != 0 ? 1u : 0u;
#pragma unroll
          for (int g = 0; g < TD / J; ++g) {
#pragma unroll
            for (int i = 0; i < J + 4; ++i) {
              // input plane g*J + kdp feeds the stacked output planes j = j_lo..j_hi through tap kd = kdp - j.  Only
              // the valid taps are issued (N = nblk * NPAD: 1..J blocks), which trims the triangular ends of the
              // Toeplitz band: (J+4) full-width MMAs would spend 2*(J-1)*J/2 of their J*(J+4) blocks on zeros.
              // An accumulation starts with the MMAs that OVERWRITE: kdp = min(4, J-1) initialises planes 0..4 and,
              // for J > 5, kdp = 9 initialises planes 5..J-1 (an MMA spans at most 5 planes); all others accumulate.
              const int kdp = kdp_order<J>(i);
              const int j_lo = kdp > 4 ? kdp - 4 : 0;
              const int j_hi = kdp < J - 1 ? kdp : J - 1;
              const int nblk = j_hi - j_lo + 1;
              const uint32_t idesc_k = ptx::make_idesc_bf16(128, nblk * (NPAD / NS), 0, 0);
              const uint32_t a_lo = a_lo1 + (uint32_t)((g * J + kdp) * kHaloH * kHaloW);
              const uint32_t b_lo = b_lo0 + (uint32_t)(((4 - kdp + j_lo) * Cfg::kBlockBytes) >> 4);
              if (leader)
                ptx::mma_bf16_split(d_base + g * Cfg::kNMma + j_lo * NPAD, a_lo, a_hi, b_lo, b_hi, idesc_k,
                                    i >= (J > 5 ? 2 : 1) ? 1u : first);
            }
          }
          if (leader) ptx::mma_commit(BAR(kWE + s));  // weight stage free once these MMAs retire
        }
        if (leader) ptx::mma_commit(BAR(2 + hb));   // halo buffer free
      }
      if (leader) ptx::mma_commit(BAR(4 + as));    // accumulators complete -> epilogue
      __syncwarp();
    }
    if (prof && lane == 0) {
      p.prof[blockIdx.x * 4 + 0] = clock64() - t_begin;
      p.prof[blockIdx.x * 4 + 1] = t_w;
      p.prof[blockIdx.x * 4 + 2] = t_h;
      p.prof[blockIdx.x * 4 + 3] = t_a;
    }
  } else if (warp >= 4) {
    // ================= epilogue: TMEM -> registers -> (bias, accumulate, round, BN sums) -> global =================
    // Two warpgroups (warps 4-7 / 8-11; a warp reads TMEM lane quadrant warp % 4) take alternate (plane, 16-column)
    // blocks: with one warp per scheduler the loop is instruction-latency bound, and on the narrow layers (in_tr,
    // out_tr, the 2-chunk 32-channel convs) it - not the MMA stream - set the pace of the kernel.
    const int wg = (warp - 4) >> 2;
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int hh = row >> 3, ww = row & 7;
    const int64_t S = (int64_t)p.d * p.h * p.w;
    float* my_stats = stat_smem + (warp - 4) * 2 * NPAD;
    constexpr int kCb = (NPAD / NS) / 16;          // 16-column blocks per plane
    const bool has_bias = p.bias != nullptr;
    const bool bias_vec = has_bias && (reinterpret_cast<uintptr_t>(p.bias) % 16 == 0);
    const bool want_stats = !SPLITK && p.sums != nullptr;
    uint32_t iuse = 0;
    for (;; ++iuse) {
      const int item = get_item(iuse);
      if (item < 0) break;
      const int tile = item / kdiv;
      const int c_first = SPLITK ? 0 : (item % NS) * n_slice;  // first output channel of this slice (nsplit > 1 only with J == 1)
      const int n = tile / items_per_n;
      int r = tile % items_per_n;
      const int tw = r % p.tiles_w; r /= p.tiles_w;
      const int th = r % p.tiles_h; const int db = r / p.tiles_h;
      const int h = th * kTileH + hh, w = tw * kTileW + ww;
      const bool inb = h < p.h && w < p.w;
      const uint32_t as = iuse % ACC_SETS, aph = (iuse / ACC_SETS) & 1;
      ptx::mbar_wait(BAR(4 + as), aph);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + as * Cfg::kAccCols + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int it = wg; it < TD * kCb; it += 2) {
        const int td = it / kCb, cb = it % kCb;
        const int d = db * TD + td;
        const bool ok = inb && d < p.d;
        const int64_t v = ((int64_t)d * p.h + h) * p.w + w;
        float acc[16];
        ptx::tmem_ld16(t_base + td * NPAD + cb * 16, acc);
        if constexpr (SPLITK) {
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int c8 = cb * 2 + k;
            if (c8 < p.out_c8 && ok) {
              // every K slice owns a private copy of the output tile: plain stores, no atomics -> the finalize kernel
              // adds the slices in a FIXED order and the result is bit-reproducible
              float* dst = p.ws + ((((int64_t)(item % kdiv) * p.n + n) * p.out_c8 + c8) * S + v) * 8;
              *reinterpret_cast<float4*>(dst) = make_float4(acc[k * 8 + 0], acc[k * 8 + 1], acc[k * 8 + 2], acc[k * 8 + 3]);
              *reinterpret_cast<float4*>(dst + 4) = make_float4(acc[k * 8 + 4], acc[k * 8 + 5], acc[k * 8 + 6], acc[k * 8 + 7]);
            }
          }
        } else {
          const int c0 = c_first + cb * 16;
          if (has_bias) {
            if (bias_vec && c0 + 16 <= p.cout_real) {
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + c0) + i);
                acc[4 * i] += b4.x; acc[4 * i + 1] += b4.y; acc[4 * i + 2] += b4.z; acc[4 * i + 3] += b4.w;
              }
            } else {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                if (c0 + i < p.cout_real) acc[i] += __ldg(p.bias + c0 + i);
            }
          }
#pragma unroll
          for (int k = 0; k < 2; ++k) {
            const int c8 = (c0 >> 3) + k;
            const bool live = ok && c8 < p.out_c8;
            uint32_t pk[4] = {0u, 0u, 0u, 0u};
            if (p.out_f32) {
              // f32 output (folded 5x5x1 partial sums; the hi/lo passes of the 3 x bf16 fp32 path, which accumulate)
              float o[8];
#pragma unroll
              for (int j = 0; j < 8; ++j) o[j] = 0.f;
              if (live) {
                float* dst = view_ptr<float>(p.out, n, c8, S, v);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = acc[k * 8 + j];
                if (p.accumulate) {
                  float old[8];
                  Vec8<float>::load(dst, old);
                  if (p.ch_scale != nullptr) {
                    const float* scp = p.ch_scale + (int64_t)n * p.out.c + c8 * 8;
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], __ldg(scp + j), old[j]);
                  } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += old[j];
                  }
                }
                Vec8<float>::store(dst, o);
              }
              if (want_stats) {  // statistics of the final f32 values (callers pass sums on the last pass only)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[k * 8 + j] = o[j];
              }
              continue;
            } else if (live) {
              __nv_bfloat16* dst = view_ptr<__nv_bfloat16>(p.out, n, c8, S, v);
              if (p.accumulate) {
                float old[8];
                Vec8<__nv_bfloat16>::load(dst, old);
                if (p.ch_scale != nullptr) {
                  const float* scp = p.ch_scale + (int64_t)n * p.out.c + c8 * 8;
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[k * 8 + j] = fmaf(acc[k * 8 + j], __ldg(scp + j), old[j]);
                } else {
#pragma unroll
                  for (int j = 0; j < 8; ++j) acc[k * 8 + j] += old[j];
                }
              }
              if (p.ep_scale != nullptr) {
                // evaluation epilogue: running-statistics BatchNorm folded to y*scale + shift, PReLU, and for a
                // block's last LUConv the residual add + second PReLU (vnet.py:41, :110, :154 in eval mode) - the
                // conv output never makes a separate BN pass
                float r[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) r[j] = 0.f;
                if (p.ep_has_res) Vec8<__nv_bfloat16>::load(view_ptr<__nv_bfloat16>(p.ep_res, n, c8, S, v), r);
                const float* scp = p.ep_scale + c8 * 8;
                const float* shp = p.ep_shift + c8 * 8;
                const float* alp = p.ep_alpha + c8 * 8;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                  float t = fmaf(acc[k * 8 + j], __ldg(scp + j), __ldg(shp + j));
                  t = t > 0.f ? t : t * __ldg(alp + j);
                  if (p.ep_has_res) {  // block tail: relu2(relu1(bn(conv)) + residual)
                    t += r[j];
                    t = t > 0.f ? t : t * __ldg(p.ep_alpha2 + c8 * 8 + j);
                  }
                  acc[k * 8 + j] = t;
                }
              }
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const __nv_bfloat162 hpair = __floats2bfloat162_rn(acc[k * 8 + 2 * i], acc[k * 8 + 2 * i + 1]);
                pk[i] = *reinterpret_cast<const uint32_t*>(&hpair);
              }
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
            }
            if (want_stats) {  // statistics of the ROUNDED values (what the next kernel reads); dead lanes add 0
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                acc[k * 8 + 2 * i] = __uint_as_float(pk[i] << 16);
                acc[k * 8 + 2 * i + 1] = __uint_as_float(pk[i] & 0xffff0000u);
              }
            }
          }
          if (want_stats) {
            float sq[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) sq[i] = acc[i] * acc[i];
            const float s1 = warp_reduce16(acc, lane);
            const float s2 = warp_reduce16(sq, lane);
            if ((lane & 1) == 0) {
              my_stats[c0 + (lane >> 1)] += s1;
              my_stats[NPAD + c0 + (lane >> 1)] += s2;
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(BAR(6 + as));
      if (want_stats && p.groups > 1) {
        // per-instance statistics: flush after every item (an item never straddles two n)
        __syncwarp();
        for (int i = lane; i < 2 * NPAD; i += 32) {
          const int stat = i / NPAD, c = i % NPAD;
          if (c < p.sums_c && my_stats[i] != 0.f)
            atomicAdd(&p.sums[((int64_t)stat * p.groups + n) * p.sums_c + c], (double)my_stats[i]);
          my_stats[i] = 0.f;
        }
        __syncwarp();
      }
    }
    if (want_stats && p.groups == 1) {
      __syncwarp();
      for (int i = lane; i < 2 * NPAD; i += 32) {
        const int stat = i / NPAD, c = i % NPAD;
        if (c < p.sums_c && my_stats[i] != 0.f) atomicAdd(&p.sums[(int64_t)stat * p.sums_c + c], (double)my_stats[i]);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
}

// split-K tail: out = round_T(sum_slices ws[slice] + bias [* ch_scale + out]), BN partial sums of the rounded values
// (T = bf16, or f32 for the three-pass fp32 engine where each pass accumulates into the f32 output)
template <typename T>
__global__ void __launch_bounds__(256) splitk_finalize_kernel(const float* __restrict__ ws, int ksplit,
                                                              const float* __restrict__ bias,
                                                              int cout_real, msb_tensor out, int64_t s, int accumulate,
                                                              const float* __restrict__ ch_scale, int groups,
                                                              double* __restrict__ sums, int sums_c,
                                                              const float* __restrict__ ep_scale,
                                                              const float* __restrict__ ep_shift,
                                                              const float* __restrict__ ep_alpha,
                                                              const float* __restrict__ ep_alpha2, msb_tensor ep_res,
                                                              int ep_has_res) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][16];
  const int c8 = blockIdx.y, n = blockIdx.z, out_c8 = gridDim.y;
  const int64_t v0 = (int64_t)blockIdx.x * 2048, v1 = min(v0 + (int64_t)2048, s);
  float b[8], sc[8], acc[16];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c8 * 8 + j;
    b[j] = (bias != nullptr && c < cout_real) ? __ldg(bias + c) : 0.f;
    sc[j] = ch_scale ? __ldg(ch_scale + (int64_t)n * out.c + c) : 1.f;
  }
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = 0.f;
  const float* wp = ws + ((int64_t)n * out_c8 + c8) * s * 8;
  const int64_t slice_stride = (int64_t)gridDim.z * out_c8 * s * 8;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {
    float o[8];
    Vec8<float>::load(wp + v * 8, o);
    for (int ks = 1; ks < ksplit; ++ks) {  // fixed order: slice 0 + slice 1 + ...
      float t[8];
      Vec8<float>::load(wp + ks * slice_stride + v * 8, t);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] += t[j];
    }
    T* dst = view_ptr<T>(out, n, c8, s, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] += b[j];
    if (accumulate) {
      float old[8];
      Vec8<T>::load(dst, old);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(o[j], sc[j], old[j]);
    }
    if (ep_scale != nullptr) {  // evaluation epilogue, as in conv_k5_fwd_kernel
      float r[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = 0.f;
      if (ep_has_res) Vec8<T>::load(view_ptr<T>(ep_res, n, c8, s, v), r);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float t = fmaf(o[j], __ldg(ep_scale + c8 * 8 + j), __ldg(ep_shift + c8 * 8 + j));
        t = t > 0.f ? t : t * __ldg(ep_alpha + c8 * 8 + j);
        if (ep_has_res) {
          t += r[j];
          t = t > 0.f ? t : t * __ldg(ep_alpha2 + c8 * 8 + j);
        }
        o[j] = t;
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j] = Vec8<T>::round(o[j]);
      acc[j] += o[j];
      acc[8 + j] += o[j] * o[j];
    }
    Vec8<T>::store(dst, o);
  }
  if (sums == nullptr) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const float r = warp_sum(acc[i]);
    if (lane == 0) red[warp][i] = r;
  }
  __syncthreads();
  if (threadIdx.x < 16) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += (double)red[w][threadIdx.x];
    const int stat = threadIdx.x >> 3, j = threadIdx.x & 7, c = c8 * 8 + j;
    const int g = groups == 1 ? 0 : n;
    if (c < sums_c) atomicAdd(&sums[((int64_t)stat * groups + g) * sums_c + c], t);
  }
}

// =====================================================================================================
// Weight gradient: dW[tap][co][ci] = sum_v X[v + tap][ci] * dY[v][co]  (K = voxels).
//   A = X  tile, MN-major (8 channels contiguous = 16 B, 8 consecutive w voxels = one 128-B core matrix)
//   B = dY tile, MN-major.   M = 128 rows = QM stacked kd-taps x min(Cin,128) channels: the X tile is stored
//   [plane q][c8][h][w][8] so that the M-group stride (SBO) is uniform across the stacked planes.
//   One MMA per (kh,kw) unit per 16-voxel K-step; up to 512/N units keep their accumulators in TMEM.
// =====================================================================================================
constexpr int kWgTileW = 16;

template <int NPAD, int TH>
struct WgCfg {
  static constexpr int kGroupBytes = (TH + 4) * (kWgTileW + 4) * 16;  // one 8-channel M-group of the X halo tile
  static constexpr int kXBytes = 16 * kGroupBytes;
  static constexpr int kDyPlaneBytes = TH * kWgTileW * 16;
  static constexpr int kDyBytes = (NPAD / 8) * kDyPlaneBytes;
  static constexpr int kMaxUnits = (512 / NPAD) < 25 ? (512 / NPAD) : 25;
  static constexpr int kSmemBytes = 2 * (kXBytes + kDyBytes) + 1024 + 128;
  static_assert(kSmemBytes <= 227 * 1024, "shared memory overflow");
};

struct WgParams {
  int n, cin_pad, cin_real, cout_real, dy_c8;
  int d, h, w;
  int tiles_w, tiles_h;
  int x_c8_total, dy_c8_total;
  int qm, kd_groups, mhalves, cin_m;    // stacked planes, ceil(5/qm), Cin/128, min(Cin,128)
  int units_per_pass, passes_per_group; // (kh,kw) units per pass
  int num_passes, chunks, tiles_per_chunk, total_tiles;
  float* ws;                            // [125][cout_real][cin_real] f32
  int dbg_swap;
  int pad, units_total;                 // 2 / 25 for the 5x5x5 conv; 0 / 1 for the pointwise (1x1x1) weight gradient
  int s2_c8n;                           // > 0: x is the BIG grid of a strided conv (s2_c8n = its planes); the M rows
                                        // are (tap, channel) and every tap's sub-lattice is fetched by a strided 5-D
                                        // TMA box (tmap_x built by make_b8_tmap_s2) - no space-to-depth copy
  int s2_khn, s2_kwn;                   // kernel extents along h, w (tap = (kd*khn + kh)*kwn + kw)
  int s2_sd, s2_sh, s2_sw;              // strides
  int csize, rounds;                    // CL kernels: cluster size and ceil(passes_per_group / csize) - the (kh,kw)
                                        // passes of one (channel half, kd plane) share their tiles by TMA multicast
  unsigned int* sched;                  // dynamic tile scheduler counters (nullptr = static; never with clusters)
};

// CL = true: clusters of p.csize CTAs walk the same tiles of one (channel half, kd plane) and own different (kh,kw)
// units (pass = round * csize + cluster rank); every TMA box is issued by one rank and multicast to all, a stage is
// refilled once the MMAs of all ranks retired (multicast tcgen05.commit).  One 256 x 256 (or 128 x 128 x 4) tap set
// fills the TMEM, so without sharing every loaded tile feeds only 2-4 taps and the kernel is L2 -> SM bound.
template <int NPAD, int TH, bool CL = false>
__global__ void __launch_bounds__(256, 1)
    conv_k5_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                         const WgParams p) {
  using Cfg = WgCfg<NPAD, TH>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* x_smem = smem;                           // [2][kXBytes]
  uint8_t* dy_smem = x_smem + 2 * Cfg::kXBytes;     // [2][kDyBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(dy_smem + 2 * Cfg::kDyBytes);
  // [0,2) full  [2,4) empty  [4] acc_full  [5] acc_empty  [16,16+2D) tile-scheduler ring, item slots at byte 512
  constexpr int kSF = 16, kSE = 16 + sched::kDepth;
  volatile int* sched_slots = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t crank = CL ? ptx::cluster_ctarank() : 0u;
  const int csize = CL ? p.csize : 1;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(2 + i), (uint32_t)csize); }
    ptx::mbar_init(BAR(4), 1);
    ptx::mbar_init(BAR(5), 4);
    if (sched::kEnabled) sched::init(BAR(kSF), BAR(kSE), 6);  // consumers: the producer thread, the MMA warp, 4 epilogue warps
    ptx::fence_mbar_init();
  }
  if (warp == 0 && lane == 0) { ptx::prefetch_tmap(&tmap_x); ptx::prefetch_tmap(&tmap_dy); }
  if (warp == 2) ptx::tmem_alloc<512>(ptx::smem_u32(tmem_slot));
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (CL) ptx::cluster_sync();  // every CTA's barriers are initialised before any peer multicasts into them
  pdl_trigger();  // after our own TMEM allocation (common.cuh, PDL rules)
  pdl_wait();

  // CL = false: item = (pass, chunk) per CTA;  CL = true: item = ((channel half, kd group, round), chunk) per CLUSTER
  const int num_items = CL ? p.mhalves * p.kd_groups * p.rounds * p.chunks : p.num_passes * p.chunks;
  const int item0 = CL ? (int)blockIdx.x / csize : (int)blockIdx.x;
  const int item_step = CL ? (int)gridDim.x / csize : (int)gridDim.x;
  const int tiles_per_n = p.d * p.tiles_h * p.tiles_w;
  // k-th item of this CTA (cluster): static round-robin, or from the scheduler warp (umma.cuh, namespace sched)
  const bool dyn = sched::kEnabled && !CL && p.sched != nullptr;
  auto static_item = [&](uint32_t k) -> int {
    const int it = item0 + (int)k * item_step;
    return it < num_items ? it : -1;
  };
  if (warp == 3 && dyn && lane == 0) sched::run(BAR(kSF), BAR(kSE), sched_slots, p.sched, num_items);

  auto decode_pass = [&](int pass, int& mh, int& g, int& u0, int& u1) {
    int pg;
    if (CL) {
      pg = (pass % p.rounds) * csize + (int)crank;  // may exceed the last pass: that rank idles (u0 == u1)
      g = (pass / p.rounds) % p.kd_groups;
      mh = pass / (p.rounds * p.kd_groups);
    } else {
      pg = pass % p.passes_per_group;
      g = (pass / p.passes_per_group) % p.kd_groups;
      mh = pass / (p.passes_per_group * p.kd_groups);
    }
    u0 = min(p.units_total, pg * p.units_per_pass);
    u1 = min(p.units_total, u0 + p.units_per_pass);
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t use = 0;
      for (uint32_t ik = 0;; ++ik) {
        const int item = dyn ? sched::next_lane(BAR(kSF), BAR(kSE), sched_slots, ik) : static_item(ik);
        if (item < 0) break;
        const int pass = item / p.chunks, chunk = item % p.chunks;
        int mh, g, u0, u1;
        decode_pass(pass, mh, g, u0, u1);
        const int t0 = chunk * p.tiles_per_chunk, t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
        const int planes_valid = min(p.qm, 5 - g * p.qm);
        const int x_planes = p.cin_m / 8;  // c8 planes per stacked kd plane
        const uint32_t bytes = (uint32_t)(planes_valid * x_planes * Cfg::kGroupBytes + p.dy_c8 * Cfg::kDyPlaneBytes);
        for (int t = t0; t < t1; ++t, ++use) {
          const int n = t / tiles_per_n;
          int r = t % tiles_per_n;
          const int tw = r % p.tiles_w; r /= p.tiles_w;
          const int th = r % p.tiles_h; const int d = r / p.tiles_h;
          const uint32_t b = use & 1, ph = (use >> 1) & 1;
          ptx::mbar_wait(BAR(2 + b), ph ^ 1);
          ptx::mbar_expect_tx(BAR(b), bytes);
          if (p.s2_c8n > 0) {
            const int per_tap = p.s2_c8n < 16 ? p.s2_c8n : 16;  // planes of one tap inside this 16-plane M tile
            for (int i = 0; i < 16 / per_tap; ++i) {
              const int plane0 = mh * 16 + i * per_tap;
              const int tap = plane0 / p.s2_c8n, c8 = plane0 % p.s2_c8n;
              const int kw = tap % p.s2_kwn, kh = (tap / p.s2_kwn) % p.s2_khn, kd = tap / (p.s2_kwn * p.s2_khn);
              ptx::tma_load_5d(ptx::smem_u32(x_smem + b * Cfg::kXBytes + i * per_tap * Cfg::kGroupBytes), &tmap_x, BAR(b),
                               0, p.s2_sw * tw * kWgTileW + kw, p.s2_sh * th * TH + kh, p.s2_sd * d + kd,
                               n * p.x_c8_total + c8);
            }
          } else if (CL) {  // box i of the tile is issued by rank i % csize for every rank
            for (int q = 0; q < planes_valid; ++q)
              if ((uint32_t)(q % csize) == crank)
                ptx::tma_load_4d_mc(ptx::smem_u32(x_smem + b * Cfg::kXBytes + q * x_planes * Cfg::kGroupBytes), &tmap_x,
                                    BAR(b), (tw * kWgTileW - p.pad) * 8, th * TH - p.pad, d + g * p.qm + q - p.pad,
                                    n * p.x_c8_total + mh * 16, cmask);
          } else {
            for (int q = 0; q < planes_valid; ++q)
              ptx::tma_load_4d(ptx::smem_u32(x_smem + b * Cfg::kXBytes + q * x_planes * Cfg::kGroupBytes), &tmap_x, BAR(b),
                               (tw * kWgTileW - p.pad) * 8, th * TH - p.pad, d + g * p.qm + q - p.pad,
                               n * p.x_c8_total + mh * 16);
          }
          if (CL) {
            if ((uint32_t)(planes_valid % csize) == crank)
              ptx::tma_load_4d_mc(ptx::smem_u32(dy_smem + b * Cfg::kDyBytes), &tmap_dy, BAR(b), tw * kWgTileW * 8, th * TH, d,
                                  n * p.dy_c8_total, cmask);
          } else
          ptx::tma_load_4d(ptx::smem_u32(dy_smem + b * Cfg::kDyBytes), &tmap_dy, BAR(b), tw * kWgTileW * 8, th * TH, d,
                           n * p.dy_c8_total);
        }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the uniform control flow / descriptor arithmetic; one elected lane issues
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, NPAD, 1, 1);
    const uint32_t a_lbo16 = p.dbg_swap ? (uint32_t)Cfg::kGroupBytes >> 4 : 8u;
    const uint32_t a_hi = ptx::desc_hi(p.dbg_swap ? 128u : (uint32_t)Cfg::kGroupBytes);
    const uint32_t b_lbo16 = p.dbg_swap ? (uint32_t)Cfg::kDyPlaneBytes >> 4 : 8u;
    const uint32_t b_hi = ptx::desc_hi(p.dbg_swap ? 128u : (uint32_t)Cfg::kDyPlaneBytes);
    uint32_t use = 0, iuse = 0;
    for (;; ++iuse) {
      const int item = sched::uniform(dyn ? sched::next(BAR(kSF), BAR(kSE), sched_slots, iuse, lane) : static_item(iuse));
      if (item < 0) break;
      const int pass = item / p.chunks, chunk = item % p.chunks;
      int mh, g, u0, u1;
      decode_pass(pass, mh, g, u0, u1);
      const int t0 = chunk * p.tiles_per_chunk, t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
      ptx::mbar_wait(BAR(5), (iuse & 1) ^ 1);
      ptx::tc_fence_after();
      for (int t = t0; t < t1; ++t, ++use) {
        const uint32_t b = use & 1, ph = (use >> 1) & 1;
        ptx::mbar_wait(BAR(b), ph);
        ptx::tc_fence_after();
        const uint32_t a_lo0 = ptx::desc_lo(ptx::smem_u32(x_smem + b * Cfg::kXBytes), a_lbo16);
        const uint32_t b_lo0 = ptx::desc_lo(ptx::smem_u32(dy_smem + b * Cfg::kDyBytes), b_lbo16);
#pragma unroll 1
        for (int hrow = 0; hrow < TH; ++hrow) {
          const uint32_t b_lo = b_lo0 + (uint32_t)(hrow * kWgTileW);
          const uint32_t acc = (t != t0 || hrow != 0) ? 1u : 0u;
          uint32_t kh = (uint32_t)u0 / 5u, kw = (uint32_t)u0 % 5u;
#pragma unroll 1
          for (int u = u0; u < u1; ++u) {
            const uint32_t a_lo = a_lo0 + (uint32_t)(hrow + kh) * (uint32_t)(kWgTileW + 4) + kw;
            if (leader) ptx::mma_bf16_split(tmem_u + (uint32_t)((u - u0) * NPAD), a_lo, a_hi, b_lo, b_hi, idesc, acc);
            if (++kw == 5u) { kw = 0u; ++kh; }
          }
        }
        if (leader) {
          if (CL) ptx::mma_commit_mc(BAR(2 + b), cmask);  // shared stage: free once every rank's MMAs retired
          else ptx::mma_commit(BAR(2 + b));
        }
      }
      if (leader) ptx::mma_commit(BAR(4));
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q4 = warp - 4;
    const int row = q4 * 32 + lane;
    const int qplane = row / p.cin_m, ci_local = row % p.cin_m;
    uint32_t iuse = 0;
    for (;; ++iuse) {
      const int item = dyn ? sched::next(BAR(kSF), BAR(kSE), sched_slots, iuse, lane) : static_item(iuse);
      if (item < 0) break;
      const int pass = item / p.chunks;
      int mh, g, u0, u1;
      decode_pass(pass, mh, g, u0, u1);
      const int kd = g * p.qm + qplane;
      const int ci = mh * 128 + ci_local;
      const bool row_ok = (qplane < p.qm) && kd < 5 && ci < p.cin_real;
      ptx::mbar_wait(BAR(4), iuse & 1);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
#pragma unroll 1
      for (int u = u0; u < u1; ++u) {
        const int tap = kd * 25 + u;
#pragma unroll 1
        for (int cb = 0; cb < NPAD / 16; ++cb) {
          float acc[16];
          ptx::tmem_ld16(t_base + (u - u0) * NPAD + cb * 16, acc);
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int co = cb * 16 + j;
              if (co < p.cout_real) red_add_f32(p.ws + ((int64_t)tap * p.cout_real + co) * p.cin_real + ci, acc[j]);
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(BAR(5));
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (CL) ptx::cluster_sync();  // no CTA may exit while a peer can still multicast into it or signal its barriers
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<512>(tmem_base);
  }
}

// dw[co][ci][tap] += ws[tap][co][ci]
__global__ void __launch_bounds__(256) wgrad_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                           int cout, int cin) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = (int64_t)cout * cin * kNumTaps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kNumTaps);
    const int64_t r = i / kNumTaps;  // co*cin + ci
    dw[i] += ws[(int64_t)tap * cout * cin + r];
  }
}

__global__ void __launch_bounds__(256) channel_sum_bf16_kernel(msb_tensor x, int64_t s, int c_real,
                                                               float* __restrict__ out) {
  pdl_wait();
  pdl_trigger();
  __shared__ float red[8][8];
  const int c8 = blockIdx.y, n = blockIdx.z;
  const int64_t chunk = 8192;
  const int64_t v0 = (int64_t)blockIdx.x * chunk, v1 = min(v0 + chunk, s);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int64_t v = v0 + threadIdx.x; v < v1; v += 256) {
    float a[8];
    Vec8<__nv_bfloat16>::load(view_ptr<__nv_bfloat16>(x, n, c8, s, v), a);
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

// ---- 2x2x2 / stride-2 weight gradients on the tensor cores ---------------------------------------------
// dW[sc][bc][tap] = sum_o big[2o+tap][bc] * small[o][sc] is a pointwise (1x1x1) weight gradient whose M rows are
// (tap, big channel); conv_k5_wgrad_kernel fetches every tap's sub-lattice of the big grid with a stride-2 5-D TMA box
// (WgParams::s2_c8n), so no space-to-depth copy is materialised.
// dw[sc][bc][tap] += ws[sc][tap*cbig + bc]
__global__ void __launch_bounds__(256) k2s2_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                          int csmall, int cbig, int taps, int cbig_real) {
  pdl_wait();
  pdl_trigger();
  // dw is [csmall_real][cbig_real][taps] or [cbig_real][csmall_real][taps] - the caller passes the matching strides
  const int64_t total = (int64_t)csmall * cbig_real * taps;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % taps);
    const int64_t r = i / taps;
    const int bc = (int)(r % cbig_real), sc = (int)(r / cbig_real);
    dw[i] += ws[((int64_t)sc * taps + tap) * cbig + bc];
  }
}

// ---- weight packing ----------------------------------------------------------------------------------
// packed[chunk][hk = kh*5+kw][k8][kdr = 4-kd][oc][j] (bf16), rc = chunk*16 + k8*8 + j.
// One block per (oc, chunk): the 16 x 125 f32 slab is read with coalesced rows into shared memory, then each
// thread emits one 16-byte vector of 8 reduction channels (mode 1 = input-gradient operand: reduction over the
// conv's Cout, produces its Cin, taps mirrored).
__global__ void __launch_bounds__(256) pack_k5_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                      int cout, int cin, int mode, int cin_pad, int cout_pad) {
  pdl_wait();
  pdl_trigger();
  __shared__ float slab[16][kNumTaps + 1];
  const int oc = blockIdx.x, chunk = blockIdx.y;
  for (int i = threadIdx.x; i < 16 * kNumTaps; i += 256) {
    const int r = i / kNumTaps, tap = i % kNumTaps;
    const int rc = chunk * 16 + r;
    float v = 0.f;
    if (mode == 0) {
      if (oc < cout && rc < cin) v = __ldg(w + ((int64_t)oc * cin + rc) * kNumTaps + tap);
    } else {
      if (rc < cout && oc < cin) v = __ldg(w + ((int64_t)rc * cin + oc) * kNumTaps + (kNumTaps - 1 - tap));
    }
    slab[r][tap] = v;
  }
  __syncthreads();
  if (threadIdx.x < 250) {
    const int kdr = threadIdx.x % 5, k8 = (threadIdx.x / 5) & 1, hk = threadIdx.x / 10;
    const int tap = (4 - kdr) * 25 + hk;  // natural tap index kd*25 + kh*5 + kw
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = slab[k8 * 8 + j][tap];
    const int64_t o = ((((int64_t)chunk * 25 + hk) * 2 + k8) * 5 + kdr) * cout_pad + oc;
    Vec8<__nv_bfloat16>::store(packed + o * 8, v);
  }
}

// Same packed image from a TAP-MAJOR f32 master weight w_tm[tap][co][ci] (the layout the weight-gradient kernels
// accumulate in, so master weights / gradients / momentum can live in it and no gradient transposition is needed).
// One thread per 16-byte output vector: mode 0 reads 8 consecutive ci (one 32-byte sector), mode 1 reads 8 values
// strided by cin that are consecutive across the threads of a warp.
__global__ void __launch_bounds__(256) pack_k5_tm_kernel(const float* __restrict__ w_tm, __nv_bfloat16* __restrict__ packed,
                                                         int cout, int cin, int mode_bits, int cin_pad, int cout_pad) {
  pdl_wait();
  pdl_trigger();
  const int mode = mode_bits & 1;
  const bool lo_part = (mode_bits & 2) != 0;  // 3 x bf16 fp32 path: pack w - bf16(w) instead of w
  const int64_t total = (int64_t)(cin_pad / 8) * kNumTaps * cout_pad;  // output vectors
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i;
    const int oc = (int)(r % cout_pad); r /= cout_pad;
    const int kdr = (int)(r % 5); r /= 5;
    const int k8 = (int)(r & 1); r >>= 1;
    const int hk = (int)(r % 25);
    const int chunk = (int)(r / 25);
    const int tap = (4 - kdr) * 25 + hk;  // natural tap index kd*25 + kh*5 + kw
    const int rc0 = chunk * 16 + k8 * 8;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int rc = rc0 + j;
      float x = 0.f;
      if (mode == 0) {
        if (oc < cout && rc < cin) x = __ldg(w_tm + ((int64_t)tap * cout + oc) * cin + rc);
      } else {
        if (rc < cout && oc < cin) x = __ldg(w_tm + ((int64_t)(kNumTaps - 1 - tap) * cout + rc) * cin + oc);
      }
      if (lo_part) x -= __bfloat162float(__float2bfloat16_rn(x));
      v[j] = x;
    }
    Vec8<__nv_bfloat16>::store(packed + i * 8, v);
  }
}

// BOTH operand images of one layer (forward: mode 0, input-gradient: mode 1) from ONE coalesced read of the tap-major
// master weight.  A block owns one source tap and a 16 co x 32 ci tile of it: it stages the tile in shared memory
// (16 rows of 128 contiguous bytes) and then warps 0-1 write the forward image - 256-byte runs of (16 co) x (8 ci) per
// (16-ci chunk, k8) - while warps 2-3 write the input-gradient image, whose K dimension is co and whose tap is mirrored:
// 512-byte runs of (32 ci) x (8 co) per k8.  8 bytes of traffic per weight instead of 12 and no 32-byte strided gathers
// (the single-image kernel above reads one sector per thread in mode 0: ~1.3 TB/s).  The 2.3 KB tile fits beside the
// largest forward CTA (3 968 B of an SM's shared memory stay free next to conv_k5_fwd_kernel<32, 8, 8>), so the
// re-pack on the side stream still overlaps the convolutions of the next forward.
constexpr int kPairCi = 32, kPairRow = kPairCi + 4;  // +4 floats: 8-float reads of consecutive rows hit distinct banks
__global__ void __launch_bounds__(128)
    pack_k5_tm_pair_kernel(const float* __restrict__ w_tm, __nv_bfloat16* __restrict__ packed_f,
                           __nv_bfloat16* __restrict__ packed_b, int cout, int cin, int lo_part, int f_cin_pad,
                           int f_cout_pad, int b_cin_pad, int b_cout_pad) {
  __shared__ __align__(16) float s[16][kPairRow];
  pdl_wait();
  pdl_trigger();
  const int t = blockIdx.x;               // source tap kd*25 + kh*5 + kw
  const int co0 = blockIdx.y * 16;        // first of the tile's 16 output channels
  const int ci0 = blockIdx.z * kPairCi;   // first of the tile's 32 input channels
  const int tid = threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int e = tid + k * 128, r = e >> 5, c = e & 31;
    float x = 0.f;
    if (co0 + r < cout && ci0 + c < cin) x = __ldg(w_tm + ((int64_t)t * cout + co0 + r) * cin + ci0 + c);
    if (lo_part) x -= __bfloat162float(__float2bfloat16_rn(x));
    s[r][c] = x;
  }
  __syncthreads();
  if (tid < 64) {  // forward image: vector (chunk, hk, k8, kdr, oc = co) holds 8 consecutive ci
    const int ocl = tid & 15, k8 = (tid >> 4) & 1, cl = tid >> 5;
    const int chunk = blockIdx.z * 2 + cl;
    if (co0 < f_cout_pad && chunk * 16 < f_cin_pad) {
      const int kdr = 4 - t / 25, hk = t % 25;
      float v[8];
      Vec8<float>::load(&s[ocl][cl * 16 + k8 * 8], v);
      const int64_t i = ((((int64_t)chunk * 25 + hk) * 2 + k8) * 5 + kdr) * f_cout_pad + co0 + ocl;
      Vec8<__nv_bfloat16>::store(packed_f + i * 8, v);
    }
  } else {  // input-gradient image: tap mirrored, vector (chunk = co / 16, hk, k8, kdr, oc = ci) holds 8 consecutive co
    const int e = tid - 64, k8 = e >> 5, ocl = e & 31, oc = ci0 + ocl;
    if (co0 < b_cin_pad && oc < b_cout_pad) {
      const int tb = kNumTaps - 1 - t;
      const int kdr = 4 - tb / 25, hk = tb % 25, chunk = blockIdx.y;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = s[k8 * 8 + j][ocl];
      const int64_t i = ((((int64_t)chunk * 25 + hk) * 2 + k8) * 5 + kdr) * b_cout_pad + oc;
      Vec8<__nv_bfloat16>::store(packed_b + i * 8, v);
    }
  }
}

// f32 B8 tensor -> bf16 hi = bf16(x) and lo = bf16(x - hi): x = hi + lo up to 2^-17 relative (3 x bf16 fp32 path)
__global__ void __launch_bounds__(256) split_hi_lo_kernel(msb_tensor x, msb_tensor hi, msb_tensor lo, int64_t s) {
  pdl_wait();
  pdl_trigger();
  const int c8 = blockIdx.y, n = blockIdx.z;
  for (int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; v < s; v += (int64_t)gridDim.x * blockDim.x) {
    float a[8], h[8], l[8];
    Vec8<float>::load(view_ptr<float>(x, n, c8, s, v), a);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      h[j] = __bfloat162float(__float2bfloat16_rn(a[j]));
      l[j] = a[j] - h[j];
    }
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(hi, n, c8, s, v), h);
    Vec8<__nv_bfloat16>::store(view_ptr<__nv_bfloat16>(lo, n, c8, s, v), l);
  }
}

// ---- w-folded 5x5x1 convolutions (narrow layers: in_tr 1->16, out_tr 32->classes) -------------------------------
// A 5x5x5 conv whose input (fold_side 0) or output (fold_side 1) has <= 3 real channels wastes a 16-wide K chunk / N
// block per kw tap.  Folding the 5 kw taps into the padded channels turns it into ONE 5x5x1 conv (5x fewer MMAs):
//   fold_side 0:  y = conv551(F(x), Wf),        F(x)[v,(j,ci)] = x[v + (j-2) e_w, ci],  Wf[co,(j,ci),kd,kh] = W[co,ci,kd,kh,j]
//   fold_side 1:  P = conv551(x, Wp),           Wp[(j,co),ci,kd,kh] = W[co,ci,kd,kh,j],  y[v,co] = sum_j P[v + (j-2) e_w,(j,co)]
// Logical (folded) weight element; cin / cout are the REAL channel counts of the 5-D weight.
__device__ __forceinline__ float fold_w_elem(const float* __restrict__ w, int cout, int cin, int fold_side, int oc,
                                             int rc, int tap25) {
  int co, ci, jw;
  if (fold_side == 0) {
    if (oc >= cout || rc >= 5 * cin) return 0.f;
    co = oc; jw = rc / cin; ci = rc % cin;
  } else {
    if (oc >= 5 * cout || rc >= cin) return 0.f;
    jw = oc / cout; co = oc % cout; ci = rc;
  }
  return __ldg(w + ((int64_t)co * cin + ci) * kNumTaps + tap25 * 5 + jw);
}

// packed[chunk][kh][k8][kdr = 4-kd][oc][j] (bf16), rc = chunk*16 + k8*8 + j;  mode 1 = input-gradient operand
__global__ void __launch_bounds__(256) pack_k551_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ packed,
                                                        int cout, int cin, int mode, int fold_side, int cin_pad,
                                                        int cout_pad) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = (int64_t)cin_pad * 25 * cout_pad;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i & 7);
    int64_t r = i >> 3;
    const int oc = (int)(r % cout_pad); r /= cout_pad;
    const int kdr = (int)(r % 5); r /= 5;
    const int k8 = (int)(r & 1); r >>= 1;
    const int kh = (int)(r % 5);
    const int chunk = (int)(r / 5);
    const int tap25 = (4 - kdr) * 5 + kh;
    const int rc = chunk * 16 + k8 * 8 + j;
    const float v = mode == 0 ? fold_w_elem(w, cout, cin, fold_side, oc, rc, tap25)
                              : fold_w_elem(w, cout, cin, fold_side, rc, oc, 24 - tap25);
    packed[i] = __float2bfloat16_rn(v);
  }
}

// dw[co][ci][kd][kh][jw] += ws[kd*5+kh][oc][rc]   (ws strides: cout_f x cin_f = folded channel counts)
__global__ void __launch_bounds__(256) wgrad551_unpack_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                              int cout, int cin, int fold_side) {
  pdl_wait();
  pdl_trigger();
  const int64_t total = (int64_t)cout * cin * kNumTaps;
  const int cout_f = fold_side == 0 ? cout : 5 * cout, cin_f = fold_side == 0 ? 5 * cin : cin;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int tap = (int)(i % kNumTaps);
    const int64_t r = i / kNumTaps;
    const int ci = (int)(r % cin), co = (int)(r / cin);
    const int jw = tap % 5, tap25 = tap / 5;
    const int oc = fold_side == 0 ? co : jw * cout + co;
    const int rc = fold_side == 0 ? jw * cin + ci : ci;
    dw[i] += ws[((int64_t)tap25 * cout_f + oc) * cin_f + rc];
  }
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(ptr);
  }
  return fn;
}

// 4-D map over a B8 bf16 view: dims (W*8, H, D, planes), box (box_w*8, box_h, box_d, box_p)
int make_b8_tmap(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_h, int box_d,
                 int box_p) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return MSB_ERR_CUDA;
  }
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  if (t.n_stride % (S * 8) != 0) {
    set_error("tensor view: n_stride must be a multiple of D*H*W*8");
    return MSB_ERR_INVALID;
  }
  const int64_t planes_total = t.n_stride / (S * 8);
  const cuuint64_t gdim[4] = {(cuuint64_t)dims.w * 8, (cuuint64_t)dims.h, (cuuint64_t)dims.d,
                              (cuuint64_t)((n - 1) * planes_total + t.c / 8)};
  const cuuint64_t gstr[3] = {(cuuint64_t)dims.w * 16, (cuuint64_t)dims.h * dims.w * 16, (cuuint64_t)S * 16};
  const cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_h, (cuuint32_t)box_d, (cuuint32_t)box_p};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (dims %d %d %d, planes %lld)", (int)r, dims.d, dims.h,
              dims.w, (long long)gdim[3]);
    return MSB_ERR_CUDA;
  }
  return MSB_OK;
}

int make_b8_tmap_hmajor(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_p, int box_h,
                        int box_d) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return MSB_ERR_CUDA;
  }
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  if (t.n_stride % (S * 8) != 0) {
    set_error("tensor view: n_stride must be a multiple of D*H*W*8");
    return MSB_ERR_INVALID;
  }
  const int64_t planes_total = t.n_stride / (S * 8);
  const cuuint64_t gdim[4] = {(cuuint64_t)dims.w * 8, (cuuint64_t)((n - 1) * planes_total + t.c / 8),
                              (cuuint64_t)dims.h, (cuuint64_t)dims.d};
  const cuuint64_t gstr[3] = {(cuuint64_t)S * 16, (cuuint64_t)dims.w * 16, (cuuint64_t)dims.h * dims.w * 16};
  const cuuint32_t box[4] = {(cuuint32_t)box_w * 8, (cuuint32_t)box_p, (cuuint32_t)box_h, (cuuint32_t)box_d};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, t.ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (h-major) failed with CUresult %d", (int)r);
    return MSB_ERR_CUDA;
  }
  return MSB_OK;
}

// 5-D map over a B8 bf16 view that samples every sw-th / sh-th voxel along w / h (elementStrides): dims (8, W, H, D,
// planes), box (8, sw*box_w, sh*box_h, 1, box_p) -> shared-memory image [plane][box_h][box_w][8].  The start coordinate
// selects the (kd, kh, kw) sub-lattice of a strided conv window (stride 1 = plain shifted box).
int make_b8_tmap_s2(CUtensorMap* map, const msb_tensor& t, int n, msb_dim3 dims, int box_w, int box_h, int box_p,
                    int sw, int sh) {
  EncodeTiledFn enc = get_encode_fn();
  if (enc == nullptr) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return MSB_ERR_CUDA;
  }
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  if (t.n_stride % (S * 8) != 0) {
    set_error("tensor view: n_stride must be a multiple of D*H*W*8");
    return MSB_ERR_INVALID;
  }
  if (sw * box_w > 256 || sh * box_h > 256) {
    set_error("strided tensor map: box exceeds 256 elements per dimension");
    return MSB_ERR_UNSUPPORTED;
  }
  const int64_t planes_total = t.n_stride / (S * 8);
  const cuuint64_t gdim[5] = {8, (cuuint64_t)dims.w, (cuuint64_t)dims.h, (cuuint64_t)dims.d,
                              (cuuint64_t)((n - 1) * planes_total + t.c / 8)};
  const cuuint64_t gstr[4] = {16, (cuuint64_t)dims.w * 16, (cuuint64_t)dims.h * dims.w * 16, (cuuint64_t)S * 16};
  const cuuint32_t box[5] = {8, (cuuint32_t)(sw * box_w), (cuuint32_t)(sh * box_h), 1, (cuuint32_t)box_p};
  const cuuint32_t estr[5] = {1, (cuuint32_t)sw, (cuuint32_t)sh, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, t.ptr, gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (strided) failed with CUresult %d", (int)r);
    return MSB_ERR_CUDA;
  }
  return MSB_OK;
}

int g_debug_flags[8] = {0, 0, 0, 0, 0, 0, 0, 0};
static long long* g_prof_buf = nullptr;  // debug flag 5: MMA-warp stall profile of the last conv_k5_fwd launch

template <int NPAD, int TD, int J, int ACC_SETS, int NS>
static int launch_fwd_ns(const CUtensorMap& tmap, FwdParams& p, int tiles, cudaStream_t st) {
  using Cfg = FwdCfg<NPAD, TD, J, ACC_SETS>;
  const int items = tiles * NS;
  const int grid = items < kNumSMs ? items : kNumSMs;
  static bool attr_set = false;
  if (!attr_set) {
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_fwd_kernel<NPAD, TD, J, ACC_SETS, NS>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  MSB_LAUNCH_PDL((conv_k5_fwd_kernel<NPAD, TD, J, ACC_SETS, NS>), dim3(grid), dim3(kFwdThreads), Cfg::kSmemBytes, st, tmap, p);
  return MSB_OK;
}

template <int NPAD, int TD, int J = 1, int ACC_SETS = 2>
static int launch_fwd(const msb_tensor& x, msb_dim3 dims, FwdParams& p, cudaStream_t st) {
  CUtensorMap tmap;
  int rc = make_b8_tmap(&tmap, x, p.n, dims, kHaloW, kHaloH, TD + 4, 2);
  if (rc) return rc;
  p.dblocks = (p.d + TD - 1) / TD;
  const int tiles = p.n * p.dblocks * p.tiles_h * p.tiles_w;
  int nsplit = 1;
  if (J == 1 && g_debug_flags[4] == 0)  // small volumes: slice the output channels so that every SM gets work
    while (nsplit < 4 && tiles * nsplit * 2 <= kNumSMs && (NPAD / (nsplit * 2)) % 16 == 0 && NPAD / (nsplit * 2) >= 32)
      nsplit *= 2;
  p.nsplit = nsplit;
  if constexpr (J == 1 && NPAD >= 128) {
    if (nsplit == 4) return launch_fwd_ns<NPAD, TD, J, ACC_SETS, 4>(tmap, p, tiles, st);
    if (nsplit == 2) return launch_fwd_ns<NPAD, TD, J, ACC_SETS, 2>(tmap, p, tiles, st);
  }
  return launch_fwd_ns<NPAD, TD, J, ACC_SETS, 1>(tmap, p, tiles, st);
}

// split-K launch (see conv_k5_fwd_kernel): TD planes x NPAD columns fill the 512 TMEM columns with ONE accumulator set
template <int NPAD, int TD>
static int launch_fwd_splitk(const msb_tensor& x, msb_dim3 dims, FwdParams& p, int ksplit, const float* bias,
                             cudaStream_t st) {
  using Cfg = FwdCfg<NPAD, TD, 1, 1>;
  CUtensorMap tmap;
  int rc = make_b8_tmap(&tmap, x, p.n, dims, kHaloW, kHaloH, TD + 4, 2);
  if (rc) return rc;
  p.dblocks = (p.d + TD - 1) / TD;
  p.nsplit = 1;
  p.ksplit = ksplit;
  p.chunks_per_split = (p.cin_pad / 16 + ksplit - 1) / ksplit;
  const int items = p.n * p.dblocks * p.tiles_h * p.tiles_w * ksplit;
  const int grid = items < kNumSMs ? items : kNumSMs;
  static bool attr_set = false;
  if (!attr_set) {
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_fwd_kernel<NPAD, TD, 1, 1, 1, true>,
                                     cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
    attr_set = true;
  }
  MSB_LAUNCH_PDL((conv_k5_fwd_kernel<NPAD, TD, 1, 1, 1, true>), dim3(grid), dim3(kFwdThreads), Cfg::kSmemBytes, st, tmap, p);
  const int64_t S = (int64_t)p.d * p.h * p.w;
  const dim3 fgrid((unsigned)((S + 2047) / 2048), (unsigned)p.out_c8, (unsigned)p.n);
  if (p.out_f32) {
    MSB_LAUNCH_PDL(splitk_finalize_kernel<float>, fgrid, dim3(256), 0, st, p.ws, ksplit, bias, p.cout_real, p.out, S,
                   p.accumulate, p.ch_scale, p.groups, p.sums, p.sums_c, p.ep_scale, p.ep_shift, p.ep_alpha, p.ep_alpha2,
                   p.ep_res, p.ep_has_res);
  } else {
    MSB_LAUNCH_PDL(splitk_finalize_kernel<__nv_bfloat16>, fgrid, dim3(256), 0, st, p.ws, ksplit, bias, p.cout_real, p.out, S,
                   p.accumulate, p.ch_scale, p.groups, p.sums, p.sums_c, p.ep_scale, p.ep_shift, p.ep_alpha, p.ep_alpha2,
                   p.ep_res, p.ep_has_res);
  }
  return MSB_OK;
}

// number of K slices the split-K path would use for this shape (0 = use the regular path)
static int splitk_slices(int npad, int n, msb_dim3 dims, int cin_pad) {
  if (npad != 128 && npad != 256) return 0;
  const int td = npad == 256 ? 2 : 4;
  const int tiles = n * ((dims.d + td - 1) / td) * ((dims.h + kTileH - 1) / kTileH) * ((dims.w + kTileW - 1) / kTileW);
  if (tiles * 2 > kNumSMs) return 0;  // enough tiles to fill the SMs: whole-K items reuse the weights better
  const int chunks = cin_pad / 16;
  int ks = kNumSMs / tiles;
  if (ks > chunks) ks = chunks;
  while (ks > 1 && chunks % ks != 0) --ks;
  return ks >= 2 ? ks : 0;
}

static inline int pad16(int c) { return (c + 15) / 16 * 16; }

template <int NPAD, int TH>
static int launch_wgrad(const msb_tensor& x, const msb_tensor& dy, WgParams& p, msb_dim3 dims, cudaStream_t st,
                        msb_dim3 p_big_dims = msb_dim3{0, 0, 0}) {
  using Cfg = WgCfg<NPAD, TH>;
  p.tiles_w = (dims.w + kWgTileW - 1) / kWgTileW;
  p.tiles_h = (dims.h + TH - 1) / TH;
  p.total_tiles = p.n * dims.d * p.tiles_h * p.tiles_w;
  int amax = Cfg::kMaxUnits;
  if (amax > p.units_total) amax = p.units_total;
  p.passes_per_group = (p.units_total + amax - 1) / amax;
  p.units_per_pass = (p.units_total + p.passes_per_group - 1) / p.passes_per_group;
  p.num_passes = p.mhalves * p.kd_groups * p.passes_per_group;
  int chunks = p.num_passes <= 2 * kNumSMs ? (2 * kNumSMs) / p.num_passes : 1;  // <= 2 items per CTA
  if (chunks > p.total_tiles) chunks = p.total_tiles;
  if (chunks < 1) chunks = 1;
  p.tiles_per_chunk = (p.total_tiles + chunks - 1) / chunks;
  p.chunks = (p.total_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  CUtensorMap tmx, tmdy;
  int rc;
  if (p.s2_c8n > 0) {  // x = big grid, sampled on the tap sub-lattices
    if ((rc = make_b8_tmap_s2(&tmx, x, p.n, p_big_dims, kWgTileW + 4, TH + 4, p.s2_c8n < 16 ? p.s2_c8n : 16, p.s2_sw,
                              p.s2_sh)))
      return rc;
  } else if ((rc = make_b8_tmap(&tmx, x, p.n, dims, kWgTileW + 4, TH + 4, 1, p.cin_m / 8))) return rc;
  if ((rc = make_b8_tmap(&tmdy, dy, p.n, dims, kWgTileW, TH, 1, dy.c / 8))) return rc;
  p.csize = 1; p.rounds = p.passes_per_group;
  // cluster launch: the passes of one (half, kd plane) share their tiles by TMA multicast.  MEASURED NEGATIVE RESULT
  // (B200, batch 2, 4-CTA clusters): 128 -> 128 @32^3 0.324 ms vs 0.210 ms, 256 -> 256 @16^3 0.538 vs 0.183 ms - with a
  // 2-stage ring every refill waits for the commits of all four ranks (cross-SM round trip per 0.5 us of MMA work) and
  // every (group, chunk) item pays a 64 K-atomic epilogue.  Off unless msb_debug_set(6, 8) (kept verified by the tests).
  if (p.s2_c8n == 0 && p.passes_per_group >= 4 && (g_debug_flags[6] & 8)) {
    const int csize = 4;
    const int nclusters = kNumSMs / csize;
    const int rounds = (p.passes_per_group + csize - 1) / csize;
    const int groups = p.mhalves * p.kd_groups * rounds;
    // chunks: makespan of the static round-robin = ceil(groups * c / nclusters) / c pass-times; a few items per cluster
    // smooth the tail, too many multiply the epilogue (every item flushes its accumulators with atomics)
    int best_c = 1;
    double best = 1e30;
    for (int c = 1; c <= 16 && c <= p.total_tiles; ++c) {
      const int waves = (groups * c + nclusters - 1) / nclusters;
      const double cost = (double)waves / c + 0.01 * c;
      if (cost < best) { best = cost; best_c = c; }
    }
    p.tiles_per_chunk = (p.total_tiles + best_c - 1) / best_c;
    p.chunks = (p.total_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.csize = csize; p.rounds = rounds;
    const int items = groups * p.chunks;
    const int grid = (items < nclusters ? items : nclusters) * csize;
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_wgrad_kernel<NPAD, TH, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     Cfg::kSmemBytes));
    cudaError_t e = launch_pdl_cluster(conv_k5_wgrad_kernel<NPAD, TH, true>, dim3(grid), dim3(256), Cfg::kSmemBytes, st,
                                       csize, tmx, tmdy, p);
    if (e != cudaSuccess) {
      set_error("clustered per-tap wgrad launch failed: %s", cudaGetErrorString(e));
      return MSB_ERR_CUDA;
    }
    return MSB_OK;
  }
  const int items = p.num_passes * p.chunks;
  const int grid = items < kNumSMs ? items : kNumSMs;
  MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_wgrad_kernel<NPAD, TH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
  MSB_LAUNCH_PDL((conv_k5_wgrad_kernel<NPAD, TH>), dim3(grid), dim3(256), Cfg::kSmemBytes, st, tmx, tmdy, p);
  return MSB_OK;
}


}  // namespace msb

using namespace msb;

extern "C" {

int msb_debug_read_prof(long long* host_out /* [148][4] */) {
  MSB_REQUIRE(host_out != nullptr && g_prof_buf != nullptr, "msb_debug_read_prof: enable debug flag 5 and run a conv first");
  MSB_CUDA_OK(cudaMemcpy(host_out, g_prof_buf, kNumSMs * 4 * sizeof(long long), cudaMemcpyDeviceToHost));
  return MSB_OK;
}

// device addresses of the scheduler counters, resolved once (outside any stream capture) by msb_set_tile_scheduler
static unsigned int* g_sched_fwd_ptr = nullptr;
static unsigned int* g_sched_wg_ptr = nullptr;

int msb_set_tile_scheduler(int dynamic) {
  if (dynamic && !sched::kEnabled) {
    set_error("msb_set_tile_scheduler: the dynamic scheduler is compiled out (rebuild with -DMSB_DYNAMIC_TILES=1)");
    return MSB_ERR_UNSUPPORTED;
  }
  if (dynamic && g_sched_fwd_ptr == nullptr) {
    MSB_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&g_sched_fwd_ptr), g_sched_fwd));
    MSB_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&g_sched_wg_ptr), g_sched_wg));
    MSB_CUDA_OK(cudaMemset(g_sched_fwd_ptr, 0, 2 * sizeof(unsigned int)));
    MSB_CUDA_OK(cudaMemset(g_sched_wg_ptr, 0, 2 * sizeof(unsigned int)));
  }
  int rc = msb_set_tile_scheduler_wgrad2(dynamic);
  if (rc) return rc;
  rc = msb_set_tile_scheduler_k2s2(dynamic);
  if (rc) return rc;
  g_dynamic_tiles = dynamic ? 1 : 0;
  return MSB_OK;
}

int msb_debug_set(int key, int value) {
  if (key < 0 || key >= 8) return MSB_ERR_INVALID;
  g_debug_flags[key] = value;
  if (key == 7) g_pdl_enabled = value ? 0 : 1;
  return MSB_OK;
}

size_t msb_conv_k5_packed_bytes(int cin_pad, int cout_pad) {
  return (size_t)cin_pad * kNumTaps * cout_pad * sizeof(__nv_bfloat16);
}

int msb_conv_k5_pack(const float* w, void* packed, int cout, int cin, int mode, int cin_pad, int cout_pad,
                     void* stream) {
  MSB_REQUIRE(w && packed && cout > 0 && cin > 0 && (mode == 0 || mode == 1), "msb_conv_k5_pack: bad arguments");
  MSB_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0, "msb_conv_k5_pack: padded channel counts must be multiples of 16");
  MSB_REQUIRE(mode == 0 ? (cin_pad >= cin && cout_pad >= cout) : (cin_pad >= cout && cout_pad >= cin),
              "msb_conv_k5_pack: padded channel counts too small");
  MSB_LAUNCH_PDL(pack_k5_kernel, dim3(cout_pad, cin_pad / 16), dim3(256), 0, as_stream(stream), w,
                 reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, mode, cin_pad, cout_pad);
  return MSB_OK;
}

int msb_split_hi_lo(msb_tensor x, msb_tensor hi, msb_tensor lo, int n, int64_t s, void* stream) {
  MSB_REQUIRE(view_ok(x) && view_ok(hi) && view_ok(lo) && x.dtype == MSB_F32 && hi.dtype == MSB_BF16 &&
                  lo.dtype == MSB_BF16 && hi.c == x.c && lo.c == x.c && n > 0 && s > 0,
              "msb_split_hi_lo: f32 B8 input and two bf16 B8 outputs of the same channel count required");
  int64_t bx = (s + 255) / 256;
  if (bx > 2048) bx = 2048;
  MSB_LAUNCH_PDL(split_hi_lo_kernel, dim3((unsigned)bx, (unsigned)(x.c / 8), (unsigned)n), dim3(256), 0, as_stream(stream),
                 x, hi, lo, s);
  return MSB_OK;
}

int msb_conv_k5_pack_tm(const float* w_tm, void* packed, int cout, int cin, int mode_bits, int cin_pad, int cout_pad,
                        void* stream) {
  const int mode = mode_bits & 1;
  MSB_REQUIRE(w_tm && packed && cout > 0 && cin > 0 && mode_bits >= 0 && mode_bits <= 3, "msb_conv_k5_pack_tm: bad arguments");
  MSB_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0, "msb_conv_k5_pack_tm: padded channel counts must be multiples of 16");
  MSB_REQUIRE(mode == 0 ? (cin_pad >= cin && cout_pad >= cout) : (cin_pad >= cout && cout_pad >= cin),
              "msb_conv_k5_pack_tm: padded channel counts too small");
  const int64_t total = (int64_t)(cin_pad / 8) * kNumTaps * cout_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  MSB_LAUNCH_PDL(pack_k5_tm_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), w_tm,
                 reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, mode_bits, cin_pad, cout_pad);
  return MSB_OK;
}

int msb_conv_k5_pack_tm_pair(const float* w_tm, void* packed_f, void* packed_b, int cout, int cin, int lo_part,
                             int f_cin_pad, int f_cout_pad, int b_cin_pad, int b_cout_pad, void* stream) {
  MSB_REQUIRE(w_tm && packed_f && packed_b && cout > 0 && cin > 0 && (lo_part == 0 || lo_part == 1),
              "msb_conv_k5_pack_tm_pair: bad arguments");
  MSB_REQUIRE(f_cin_pad % 16 == 0 && f_cout_pad % 16 == 0 && b_cin_pad % 16 == 0 && b_cout_pad % 16 == 0,
              "msb_conv_k5_pack_tm_pair: padded channel counts must be multiples of 16");
  MSB_REQUIRE(f_cin_pad >= cin && f_cout_pad >= cout && b_cin_pad >= cout && b_cout_pad >= cin,
              "msb_conv_k5_pack_tm_pair: padded channel counts too small");
  MSB_REQUIRE(f_cin_pad <= 256 && f_cout_pad <= 256 && b_cin_pad <= 256 && b_cout_pad <= 256,
              "msb_conv_k5_pack_tm_pair: at most 256 (padded) channels");
  const int rows = f_cout_pad > b_cin_pad ? f_cout_pad : b_cin_pad;
  const int width = f_cin_pad > b_cout_pad ? f_cin_pad : b_cout_pad;
  MSB_LAUNCH_PDL(pack_k5_tm_pair_kernel, dim3(kNumTaps, (unsigned)(rows / 16), (unsigned)((width + kPairCi - 1) / kPairCi)),
                 dim3(128), 0, as_stream(stream), w_tm,
                 reinterpret_cast<__nv_bfloat16*>(packed_f), reinterpret_cast<__nv_bfloat16*>(packed_b), cout, cin, lo_part,
                 f_cin_pad, f_cout_pad, b_cin_pad, b_cout_pad);
  return MSB_OK;
}

int msb_conv_k5_out_pad(int cout_view);

// J = TD = 8 stacks cost 736 clk per 8 output planes, J = TD = 4 stacks 416 clk per 4 (measured for N_pad = 32, see the
// dispatch below): the deep stack pays when it does not round the depth up by more than that ratio
static inline bool deep_stack_pays(int d) { return ((d + 7) / 8) * 736 <= ((d + 3) / 4) * 416; }

static int conv_k5_fwd_impl(const char* who, msb_tensor x, const void* packed, const float* bias, int cout,
                           msb_tensor out, int n, msb_dim3 dims, int accumulate, const float* ch_scale, int groups,
                           double* sums, int kw_taps, void* stream, void* workspace = nullptr,
                           size_t workspace_bytes = 0, const float* ep_scale = nullptr, const float* ep_shift = nullptr,
                           const float* ep_alpha = nullptr, const float* ep_alpha2 = nullptr,
                           const msb_tensor* ep_res = nullptr) {
  MSB_REQUIRE(view_ok(x) && view_ok(out) && x.dtype == MSB_BF16 && packed && n > 0, "%s: bf16 B8 input view required", who);
  MSB_REQUIRE(out.dtype == MSB_BF16 || out.dtype == MSB_F32, "%s: bf16 or f32 B8 output view required", who);
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0, "%s: bad dims", who);
  MSB_REQUIRE(x.c % 16 == 0, "%s: input channels must be a multiple of 16 (pad the buffer)", who);
  MSB_REQUIRE(cout > 0 && cout <= out.c && out.c <= 256, "%s: cout must fit the output view (<= 256)", who);
  MSB_REQUIRE(groups == 1 || groups == n, "%s: groups must be 1 or n", who);
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  FwdParams p;
  p.n = n; p.cin_pad = x.c; p.cout_real = cout; p.out_c8 = out.c / 8;
  p.d = dims.d; p.h = dims.h; p.w = dims.w;
  p.tiles_w = (dims.w + kTileW - 1) / kTileW;
  p.tiles_h = (dims.h + kTileH - 1) / kTileH;
  p.x_c8_total = (int)(x.n_stride / (S * 8));
  p.packed = packed; p.bias = bias; p.out = out; p.accumulate = accumulate; p.ch_scale = ch_scale;
  p.groups = groups; p.sums = sums; p.sums_c = out.c; p.dbg_swap = 0;
  p.kw_taps = kw_taps; p.out_f32 = out.dtype == MSB_F32;
  p.ep_scale = ep_scale; p.ep_shift = ep_shift; p.ep_alpha = ep_alpha; p.ep_alpha2 = ep_alpha2;
  p.ep_has_res = ep_res != nullptr;
  p.ep_res = ep_res != nullptr ? *ep_res : out;
  if (ep_scale != nullptr) {
    MSB_REQUIRE(ep_shift && ep_alpha && out.dtype == MSB_BF16 && !accumulate && sums == nullptr,
                "%s: the evaluation epilogue needs scale, shift and alpha, a bf16 output, no accumulate and no BN sums", who);
    MSB_REQUIRE(ep_res == nullptr || (view_ok(*ep_res) && ep_res->dtype == MSB_BF16 && ep_res->c >= out.c && ep_alpha2),
                "%s: the residual must be a bf16 B8 view with at least the output's channels and needs alpha2", who);
  }
  p.prof = nullptr;
  if (g_debug_flags[5]) {
    if (g_prof_buf == nullptr) MSB_CUDA_OK(cudaMalloc(&g_prof_buf, kNumSMs * 4 * sizeof(long long)));
    MSB_CUDA_OK(cudaMemsetAsync(g_prof_buf, 0, kNumSMs * 4 * sizeof(long long), as_stream(stream)));
    p.prof = g_prof_buf;
  }
  p.ws = nullptr; p.ksplit = 1; p.chunks_per_split = x.c / 16;
  p.sched = g_dynamic_tiles ? g_sched_fwd_ptr : nullptr;
  cudaStream_t st = as_stream(stream);
  const int npad_sel = msb_conv_k5_out_pad(out.c);
  // NOTE: the packed operand must have been built with cout_pad == npad_sel.
  if (workspace != nullptr && kw_taps == 5 && (g_debug_flags[6] & 1) == 0) {
    const int ks = splitk_slices(npad_sel, n, dims, x.c);
    if (ks > 0) {
      const size_t need = (size_t)ks * n * out.c * S * sizeof(float);
      MSB_REQUIRE(workspace_bytes >= need && reinterpret_cast<uintptr_t>(workspace) % 16 == 0,
                  "%s: split-K workspace too small or misaligned (%zu < %zu)", who, workspace_bytes, need);
      p.ws = reinterpret_cast<float*>(workspace);
      return npad_sel == 256 ? launch_fwd_splitk<256, 2>(x, dims, p, ks, bias, st)
                             : launch_fwd_splitk<128, 4>(x, dims, p, ks, bias, st);
    }
  }
  // plane stacking along N (debug flag 3 = 1 disables it): narrow outputs are bound by the A-operand fetch, so one
  // MMA produces J output planes from one activation window (see FwdCfg)
  const bool stack = g_debug_flags[3] != 1;
  switch (npad_sel) {
    // deep stacks (J = TD = 8) amortise the triangular ends of the tap band best (736 clk per 8 planes vs 416 per 4
    // for N_pad = 32) and re-fetch less halo ((TD+4)/TD); shallow volumes keep TD = 4
    // ... unless the depth wastes planes: d = 12 (MRISpineSeg 512x512x12) is two 8-plane blocks with 4 idle planes
    // (2 x 736 clk per tile column) but three full 4-plane blocks (3 x 416 clk): pick the cheaper block count
    case 16:
      if (stack && dims.d >= 8 && deep_stack_pays(dims.d)) return launch_fwd<16, 8, 8>(x, dims, p, st);
      return stack ? launch_fwd<16, 4, 4>(x, dims, p, st) : launch_fwd<16, 4>(x, dims, p, st);
    case 32:
      if (stack && dims.d >= 8 && g_debug_flags[3] != 2 && deep_stack_pays(dims.d))
        return launch_fwd<32, 8, 8>(x, dims, p, st);
      return stack ? launch_fwd<32, 4, 4>(x, dims, p, st) : launch_fwd<32, 4>(x, dims, p, st);
    case 64: return stack ? launch_fwd<64, 4, 4>(x, dims, p, st) : launch_fwd<64, 4>(x, dims, p, st);
    case 128: return launch_fwd<128, 2>(x, dims, p, st);
    default: {
      // 256 channels: every item streams the whole 16 MB weight set, so two planes per item (all 512 TMEM columns, one
      // accumulator set) halve the L2 -> SM traffic; worth it as soon as the two-plane tiles still fill most SMs
      const int tiles2 = n * ((dims.d + 1) / 2) * p.tiles_h * p.tiles_w;
      if (dims.d >= 2 && tiles2 * 4 >= kNumSMs * 3) return launch_fwd<256, 2, 1, 1>(x, dims, p, st);
      return launch_fwd<256, 1>(x, dims, p, st);
    }
  }
}

int msb_conv_k5_fwd(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                    msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* stream) {
  return conv_k5_fwd_impl("msb_conv_k5_fwd", x, packed, bias, cout, out, n, dims, accumulate, ch_scale, groups, sums, 5,
                          stream);
}

size_t msb_conv_k5_fwd_workspace_bytes(int n, int cout_view, msb_dim3 dims, int cin_view) {
  const int npad = msb_conv_k5_out_pad(cout_view);
  const int ks = splitk_slices(npad, n, dims, cin_view);
  if (ks == 0) return 0;
  return (size_t)ks * n * cout_view * dims.d * dims.h * dims.w * sizeof(float);  // one private copy per K slice
}

int msb_conv_k5_fwd_ws(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                       msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* workspace,
                       size_t workspace_bytes, void* stream) {
  return conv_k5_fwd_impl("msb_conv_k5_fwd_ws", x, packed, bias, cout, out, n, dims, accumulate, ch_scale, groups, sums,
                          5, stream, workspace, workspace_bytes);
}

int msb_conv_k5_fwd_act(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                        msb_dim3 dims, const float* scale, const float* shift, const float* alpha,
                        const msb_tensor* residual, const float* alpha2, void* workspace, size_t workspace_bytes,
                        void* stream) {
  MSB_REQUIRE(scale != nullptr, "msb_conv_k5_fwd_act: scale / shift / alpha are required");
  return conv_k5_fwd_impl("msb_conv_k5_fwd_act", x, packed, bias, cout, out, n, dims, 0, nullptr, 1, nullptr, 5, stream,
                          workspace, workspace_bytes, scale, shift, alpha, alpha2, residual);
}

int msb_conv_k551_fwd(msb_tensor x, const void* packed, const float* bias, int cout, msb_tensor out, int n,
                      msb_dim3 dims, int accumulate, const float* ch_scale, int groups, double* sums, void* stream) {
  return conv_k5_fwd_impl("msb_conv_k551_fwd", x, packed, bias, cout, out, n, dims, accumulate, ch_scale, groups, sums,
                          1, stream);
}

size_t msb_conv_k5_wgrad_workspace_bytes(int cin, int cout) { return (size_t)kNumTaps * cin * cout * sizeof(float); }

static int conv_k5_wgrad_impl(msb_tensor x, msb_tensor dy, float* dw, float* dbias, int cout, int cin, int n, msb_dim3 dims,
                              void* workspace, size_t workspace_bytes, int dw_tap_major, void* stream);

int msb_conv_k5_wgrad(msb_tensor x, msb_tensor dy, float* dw, float* dbias, int cout, int cin, int n, msb_dim3 dims,
                      void* workspace, size_t workspace_bytes, void* stream) {
  return conv_k5_wgrad_impl(x, dy, dw, dbias, cout, cin, n, dims, workspace, workspace_bytes, 0, stream);
}

int msb_conv_k5_wgrad_tm(msb_tensor x, msb_tensor dy, float* dw_tm, float* dbias, int cout, int cin, int n,
                         msb_dim3 dims, void* stream) {
  return conv_k5_wgrad_impl(x, dy, dw_tm, dbias, cout, cin, n, dims, dw_tm, (size_t)kNumTaps * cin * cout * sizeof(float),
                            1, stream);
}

static int conv_k5_wgrad_impl(msb_tensor x, msb_tensor dy, float* dw, float* dbias, int cout, int cin, int n, msb_dim3 dims,
                              void* workspace, size_t workspace_bytes, int dw_tap_major, void* stream) {
  MSB_REQUIRE(view_ok(x) && view_ok(dy) && x.dtype == MSB_BF16 && dy.dtype == MSB_BF16 && dw && n > 0,
              "msb_conv_k5_wgrad: bf16 B8 views required");
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0, "msb_conv_k5_wgrad: bad dims");
  MSB_REQUIRE(x.c == 16 || x.c == 32 || x.c == 64 || x.c == 128 || x.c == 256,
              "msb_conv_k5_wgrad: input view must have 16/32/64/128/256 channels");
  MSB_REQUIRE(cin > 0 && cin <= x.c && cout > 0 && cout <= dy.c && dy.c <= 256, "msb_conv_k5_wgrad: bad channel counts");
  const size_t need = msb_conv_k5_wgrad_workspace_bytes(cin, cout);
  MSB_REQUIRE(workspace && workspace_bytes >= need, "msb_conv_k5_wgrad: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  // tap-major gradient buffer: the kernels' [tap][co][ci] f32 atomics ARE the += into dw - no memset, no unpack
  if (!dw_tap_major) MSB_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  WgParams p;
  p.n = n; p.cin_pad = x.c; p.cin_real = cin; p.cout_real = cout; p.dy_c8 = dy.c / 8;
  p.d = dims.d; p.h = dims.h; p.w = dims.w;
  p.x_c8_total = (int)(x.n_stride / (S * 8));
  p.dy_c8_total = (int)(dy.n_stride / (S * 8));
  p.cin_m = x.c < 128 ? x.c : 128;
  p.mhalves = x.c > 128 ? x.c / 128 : 1;
  p.qm = 128 / p.cin_m;
  const int qeff = p.qm < 5 ? p.qm : 5;
  p.kd_groups = (5 + qeff - 1) / qeff;
  p.ws = reinterpret_cast<float*>(workspace);
  p.sched = g_dynamic_tiles ? g_sched_wg_ptr : nullptr;
  p.dbg_swap = g_debug_flags[1];
  p.pad = 2; p.units_total = 25; p.s2_c8n = 0; p.s2_khn = p.s2_kwn = p.s2_sd = p.s2_sh = p.s2_sw = 1;
  const int npad = msb_conv_k5_out_pad(dy.c);
  int rc = MSB_ERR_UNSUPPORTED;
  if (g_debug_flags[2] == 0) rc = launch_wgrad_v2(x, dy, cout, cin, n, dims, p.ws, st);
  if (rc != MSB_ERR_UNSUPPORTED) {
    if (rc) return rc;
  } else
  switch (npad) {
    case 16: rc = launch_wgrad<16, 8>(x, dy, p, dims, st); break;
    case 32: rc = launch_wgrad<32, 8>(x, dy, p, dims, st); break;
    case 64: rc = launch_wgrad<64, 8>(x, dy, p, dims, st); break;
    case 128: rc = launch_wgrad<128, 8>(x, dy, p, dims, st); break;
    default: rc = launch_wgrad<256, 4>(x, dy, p, dims, st); break;
  }
  if (rc) return rc;
  if (!dw_tap_major) {
    const int64_t total = (int64_t)cout * cin * kNumTaps;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    MSB_LAUNCH_PDL(wgrad_unpack_kernel, dim3(blocks), dim3(256), 0, st, p.ws, dw, cout, cin);
  }
  if (dbias != nullptr) {
    const dim3 grid((unsigned)((S + 8191) / 8192), dy.c / 8, n);
    MSB_LAUNCH_PDL(channel_sum_bf16_kernel, grid, dim3(256), 0, st, dy, S, cout, dbias);
  }
  return MSB_OK;
}

size_t msb_conv_k551_packed_bytes(int cin_pad, int cout_pad) {
  return (size_t)cin_pad * 25 * cout_pad * sizeof(__nv_bfloat16);
}

int msb_conv_k551_pack(const float* w, void* packed, int cout, int cin, int mode, int fold_side, int cin_pad,
                       int cout_pad, void* stream) {
  MSB_REQUIRE(w && packed && cout > 0 && cin > 0 && (mode == 0 || mode == 1) && (fold_side == 0 || fold_side == 1),
              "msb_conv_k551_pack: bad arguments");
  MSB_REQUIRE(cin_pad % 16 == 0 && cout_pad % 16 == 0, "msb_conv_k551_pack: padded channel counts must be multiples of 16");
  const int cin_f = fold_side == 0 ? 5 * cin : cin, cout_f = fold_side == 0 ? cout : 5 * cout;
  MSB_REQUIRE(mode == 0 ? (cin_pad >= cin_f && cout_pad >= cout_f) : (cin_pad >= cout_f && cout_pad >= cin_f),
              "msb_conv_k551_pack: padded channel counts too small for the folded channels");
  const int64_t total = (int64_t)cin_pad * 25 * cout_pad;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  MSB_LAUNCH_PDL(pack_k551_kernel, dim3(blocks), dim3(256), 0, as_stream(stream), w,
                 reinterpret_cast<__nv_bfloat16*>(packed), cout, cin, mode, fold_side, cin_pad, cout_pad);
  return MSB_OK;
}

size_t msb_conv_k551_wgrad_workspace_bytes(int cin, int cout, int fold_side) {
  const int cin_f = fold_side == 0 ? 5 * cin : cin, cout_f = fold_side == 0 ? cout : 5 * cout;
  return (size_t)25 * cin_f * cout_f * sizeof(float);
}

int msb_conv_k551_wgrad(msb_tensor x, msb_tensor dy, float* dw, int cout, int cin, int fold_side, int n, msb_dim3 dims,
                        void* workspace, size_t workspace_bytes, void* stream) {
  MSB_REQUIRE(view_ok(x) && view_ok(dy) && x.dtype == MSB_BF16 && dy.dtype == MSB_BF16 && dw && n > 0,
              "msb_conv_k551_wgrad: bf16 B8 views required");
  MSB_REQUIRE(dims.d > 0 && dims.h > 0 && dims.w > 0, "msb_conv_k551_wgrad: bad dims");
  MSB_REQUIRE(fold_side == 0 || fold_side == 1, "msb_conv_k551_wgrad: fold_side must be 0 or 1");
  MSB_REQUIRE(x.c == 16 || x.c == 32 || x.c == 64 || x.c == 128, "msb_conv_k551_wgrad: input view must have 16..128 channels");
  const int cin_f = fold_side == 0 ? 5 * cin : cin, cout_f = fold_side == 0 ? cout : 5 * cout;
  MSB_REQUIRE(cin > 0 && cout > 0 && cin_f <= x.c && cout_f <= dy.c, "msb_conv_k551_wgrad: folded channels exceed the views");
  const size_t need = msb_conv_k551_wgrad_workspace_bytes(cin, cout, fold_side);
  MSB_REQUIRE(workspace && workspace_bytes >= need, "msb_conv_k551_wgrad: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  MSB_CUDA_OK(cudaMemsetAsync(workspace, 0, need, st));
  float* ws = reinterpret_cast<float*>(workspace);
  int rc = launch_wgrad_v2(x, dy, cout_f, cin_f, n, dims, ws, st, 1);
  if (rc == MSB_ERR_UNSUPPORTED) set_error("msb_conv_k551_wgrad: shape not supported by the kh-stacked kernel");
  if (rc) return rc;
  const int64_t total = (int64_t)cout * cin * kNumTaps;
  const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
  MSB_LAUNCH_PDL(wgrad551_unpack_kernel, dim3(blocks), dim3(256), 0, st, ws, dw, cout, cin, fold_side);
  return MSB_OK;
}

int msb_channel_sum(msb_tensor x, int c_real, int n, int64_t s, float* out, void* stream) {
  MSB_REQUIRE(view_ok(x) && x.dtype == MSB_BF16 && out && n > 0 && s > 0 && c_real > 0 && c_real <= x.c,
              "msb_channel_sum: bf16 B8 view required");
  const dim3 grid((unsigned)((s + 8191) / 8192), (unsigned)((c_real + 7) / 8), n);
  MSB_LAUNCH_PDL(channel_sum_bf16_kernel, grid, dim3(256), 0, as_stream(stream), x, s, c_real, out);
  return MSB_OK;
}

size_t msb_conv_tc_wgrad_workspace_bytes(int c_big, int c_small, msb_dim3 kernel) {
  return (size_t)c_small * kernel.d * kernel.h * kernel.w * c_big * sizeof(float);
}

int msb_conv_tc_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n, msb_dim3 big_dims,
                      msb_dim3 kernel, msb_dim3 stride, int bias_from_big, void* workspace, size_t workspace_bytes,
                      void* stream) {
  MSB_REQUIRE(view_ok(big) && view_ok(small) && big.dtype == MSB_BF16 && small.dtype == MSB_BF16 && dw && n > 0,
              "msb_conv_tc_wgrad: bf16 B8 views required");
  MSB_REQUIRE(kernel.d > 0 && kernel.h > 0 && kernel.w > 0 && stride.d > 0 && stride.h > 0 && stride.w > 0 &&
                  stride.w <= 4 && stride.h <= 4 && big_dims.d >= kernel.d && big_dims.h >= kernel.h &&
                  big_dims.w >= kernel.w,
              "msb_conv_tc_wgrad: bad kernel / stride / dims");
  const int taps = kernel.d * kernel.h * kernel.w;
  MSB_REQUIRE(big.c == 16 || big.c == 32 || big.c == 64 || big.c == 128, "msb_conv_tc_wgrad: big.c in {16,32,64,128}");
  MSB_REQUIRE((taps * big.c) % 128 == 0, "msb_conv_tc_wgrad: taps * big.c must be a multiple of 128");
  MSB_REQUIRE(small.c % 16 == 0 && small.c <= 256, "msb_conv_tc_wgrad: small.c must be a multiple of 16 (<= 256)");
  const size_t need = msb_conv_tc_wgrad_workspace_bytes(big.c, small.c, kernel);
  MSB_REQUIRE(workspace && workspace_bytes >= need, "msb_conv_tc_wgrad: workspace too small (%zu < %zu)",
              workspace_bytes, need);
  cudaStream_t st = as_stream(stream);
  const msb_dim3 sd = {(big_dims.d - kernel.d) / stride.d + 1, (big_dims.h - kernel.h) / stride.h + 1,
                       (big_dims.w - kernel.w) / stride.w + 1};
  const int64_t Ss = (int64_t)sd.d * sd.h * sd.w, Sb = (int64_t)big_dims.d * big_dims.h * big_dims.w;
  const int cxs = taps * big.c;  // M rows = (tap, big channel): a pointwise weight gradient over the tap sub-lattices
  float* ws = reinterpret_cast<float*>(workspace);
  MSB_CUDA_OK(cudaMemsetAsync(ws, 0, need, st));
  WgParams p;
  p.n = n; p.cin_pad = cxs; p.cin_real = cxs; p.cout_real = small.c; p.dy_c8 = small.c / 8;
  p.d = sd.d; p.h = sd.h; p.w = sd.w;
  p.x_c8_total = (int)(big.n_stride / (Sb * 8));
  p.dy_c8_total = (int)(small.n_stride / (Ss * 8));
  p.cin_m = 128; p.mhalves = cxs / 128; p.qm = 1; p.kd_groups = 1;
  p.ws = ws; p.dbg_swap = 0; p.pad = 0; p.units_total = 1; p.s2_c8n = big.c / 8;
  p.sched = g_dynamic_tiles ? g_sched_wg_ptr : nullptr;
  p.s2_khn = kernel.h; p.s2_kwn = kernel.w; p.s2_sd = stride.d; p.s2_sh = stride.h; p.s2_sw = stride.w;
  int rc;
  switch (msb_conv_k5_out_pad(small.c)) {
    case 16: rc = launch_wgrad<16, 8>(big, small, p, sd, st, big_dims); break;
    case 32: rc = launch_wgrad<32, 8>(big, small, p, sd, st, big_dims); break;
    case 64: rc = launch_wgrad<64, 8>(big, small, p, sd, st, big_dims); break;
    case 128: rc = launch_wgrad<128, 8>(big, small, p, sd, st, big_dims); break;
    default: rc = launch_wgrad<256, 4>(big, small, p, sd, st, big_dims); break;
  }
  if (rc) return rc;
  {
    const int64_t total = (int64_t)small.c * big.c * taps;
    const int blocks = (int)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    MSB_LAUNCH_PDL(k2s2_unpack_kernel, dim3(blocks), dim3(256), 0, st, ws, dw, small.c, big.c, taps, big.c);
  }
  if (dbias != nullptr) {
    const msb_tensor& bt = bias_from_big ? big : small;
    const int64_t sbt = bias_from_big ? Sb : Ss;
    const dim3 grid((unsigned)((sbt + 8191) / 8192), bt.c / 8, n);
    MSB_LAUNCH_PDL(channel_sum_bf16_kernel, grid, dim3(256), 0, st, bt, sbt, bt.c, dbias);
  }
  return MSB_OK;
}

size_t msb_conv_k2s2_wgrad_workspace_bytes(int n, int c_big, int c_small, msb_dim3 big_dims) {
  (void)n; (void)big_dims;
  return msb_conv_tc_wgrad_workspace_bytes(c_big, c_small, msb_dim3{2, 2, 2});
}

int msb_conv_k2s2_wgrad(msb_tensor big, msb_tensor small, float* dw, float* dbias, int n, msb_dim3 big_dims,
                        int bias_from_big, void* workspace, size_t workspace_bytes, void* stream) {
  MSB_REQUIRE(big_dims.d % 2 == 0 && big_dims.h % 2 == 0 && big_dims.w % 2 == 0,
              "msb_conv_k2s2_wgrad: the large grid must have even extents");
  return msb_conv_tc_wgrad(big, small, dw, dbias, n, big_dims, msb_dim3{2, 2, 2}, msb_dim3{2, 2, 2}, bias_from_big,
                           workspace, workspace_bytes, stream);
}

int msb_conv_k5_out_pad(int cout_view) {
  const int npad = pad16(cout_view);
  return npad <= 16 ? 16 : npad <= 32 ? 32 : npad <= 64 ? 64 : npad <= 128 ? 128 : 256;
}

}  // extern "C"
