// Weight gradient of the 5x5x5 conv, "kh-stacked" variant for narrow outputs (Cout <= 48):
//   dW[kd,kh,kw][ci][co] = sum_u X[u + (kd-2, 0, kw-2)][ci] * dY[u - (0, kh-2, 0)][co]     (u = input voxel)
// One tcgen05.mma accumulates D[(kd-stack, ci) x (kh-stack, co)]:
//   * M = 128 rows = QM stacked kd planes x min(Cin,128) channels  (X tile stored [plane][c8][h][w][8])
//   * N = 5 stacked kh shifts x 8*ceil(Cout/8) channels (padded to 16) — the dY tile is stored [h][c8][w][8] with a
//     +-2 row halo, so the five row shifts are FIVE CONSECUTIVE 8-channel groups of one MN-major operand with a
//     uniform group stride: a single descriptor, no copies.
//   * K = 16 consecutive w voxels of one input row.
// Versus the per-tap kernel this issues 5x fewer, 5x wider MMAs (operand bytes per MAC drop ~2.8x), which is what
// bounds small-N tcgen05 work (measured: SWIZZLE_NONE operand fetch ~70 B/clk).
#include <cuda.h>

#include "common.cuh"
#include "conv_k5.cuh"
#include "umma.cuh"

namespace msb {

constexpr int kW2TileW = 16, kW2TileH = 8;
constexpr int kW2GroupBytes = kW2TileH * (kW2TileW + 4) * 16;  // one 8-channel M-group of the X tile (w halo only)
constexpr int kW2XBytes = 16 * kW2GroupBytes;                  // 128 M rows
constexpr int kW2RowBytes = kW2TileW * 16;                     // one (h, c8) row of the dY tile
constexpr int kW2MaxStages = 4;  // stages are a launch parameter: 4 when the tiles fit (32 channels), else 3

struct Wg2Params {
  int n, cin_real, cout_real, dyp, npad;
  int jh, jgroups;                       // kh shifts stacked per MMA, number of kh groups
  int d, h, w, tiles_w, tiles_h;
  int x_c8_total, dy_c8_total;
  int qm, kd_groups, mhalves, cin_m;
  int units_per_pass, passes_per_group, num_passes, chunks, tiles_per_chunk, total_tiles;
  int dy_stage_bytes;
  int stages;                            // TMA ring depth (3 or 4)
  unsigned char pass_order[64];          // passes sorted by MMA cost (heavy first): CTA b gets items b, b+grid, ... =
                                         // one heavy + one light pass instead of two heavy ones
  float* ws;
  int kw_taps, kw_base;                  // 5 / 0 for the 5x5x5 kernel, 1 / 2 (centre tap only) for the 5x5x1 kernel
  int csize;                             // cluster size: the (kh-group, kw-subset) passes of one (channel half, kd group)
                                         // load identical tiles -> they run as ONE cluster sharing them by TMA multicast
  int balance;                           // 2-CTA clusters, 5 kw taps: rank 0 owns kw {0,1}, rank 1 kw {3,4}, the middle
                                         // tap alternates with the tile parity (2.5 MMAs per row each instead of 3 / 2)
  int rep;                               // balanced 32-channel clusters: the LAST kd group (plane kd = 4 alone, 3 of its 4
                                         // M slots idle) is loaded as 4 kw-shifted replicas of that plane instead, so ONE
                                         // MMA covers kw = replica (+ rank): 2 MMAs per row for the group instead of 5
  int gchunks[2], gtpc[2];               // rep: chunks / tiles per chunk of kd group 0 and 1 (sized by their MMA cost)
  unsigned int* sched;                   // dynamic tile scheduler counters (nullptr = static; never with clusters)
};

static __device__ unsigned int g_sched_wg2[2];
static unsigned int* g_sched_wg2_ptr = nullptr;
static int g_dynamic_tiles_wg2 = 0;

int msb_set_tile_scheduler_wgrad2(int dynamic) {
  if (dynamic && g_sched_wg2_ptr == nullptr) {
    MSB_CUDA_OK(cudaGetSymbolAddress(reinterpret_cast<void**>(&g_sched_wg2_ptr), g_sched_wg2));
    MSB_CUDA_OK(cudaMemset(g_sched_wg2_ptr, 0, 2 * sizeof(unsigned int)));
  }
  g_dynamic_tiles_wg2 = dynamic ? 1 : 0;
  return MSB_OK;
}

// CL = true: launched in clusters of p.csize CTAs.  All CTAs of a cluster walk the same tiles of the same (channel half,
// kd group) and differ only in the accumulators they own (kh group x kw subset = cluster rank); every TMA box is issued
// by ONE rank and multicast to all, a stage is refilled once the MMAs of ALL ranks have consumed it (multicast
// tcgen05.commit on the empty barriers).  L2 -> SM traffic drops by the cluster size (4.97 GB per launch for the 32 -> 32
// layer at 128^3 before, ncu) - the kernel was bound by re-loading its tiles once per pass, not by the tensor pipe.
template <bool CL>
__global__ void __launch_bounds__(256, 1)
    conv_k5_wgrad2_kernel(const __grid_constant__ CUtensorMap tmap_x, const __grid_constant__ CUtensorMap tmap_dy,
                          const Wg2Params p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
  uint8_t* dy_smem = smem;                                        // [stages][dy_stage_bytes]
  const int kW2Stages = p.stages;
  uint8_t* x_smem = dy_smem + kW2Stages * p.dy_stage_bytes;       // [stages][kW2XBytes]
  uint64_t* bars = reinterpret_cast<uint64_t*>(x_smem + kW2Stages * kW2XBytes);
  // [0,4) full  [4,8) empty  [8] acc_full  [9] acc_empty  [16,16+2D) tile-scheduler ring, item slots at byte 512
  constexpr int kEmpty = kW2MaxStages, kAccFull = 2 * kW2MaxStages, kAccEmpty = 2 * kW2MaxStages + 1;
  constexpr int kSF = 16, kSE = 16 + sched::kDepth;
  volatile int* sched_slots = reinterpret_cast<volatile int*>(reinterpret_cast<uint8_t*>(bars) + 512);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t bar0 = ptx::smem_u32(bars);
  auto BAR = [&](int i) { return bar0 + 8u * (uint32_t)i; };
  const uint32_t crank = CL ? ptx::cluster_ctarank() : 0u;
  const int csize = CL ? p.csize : 1;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1u);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kW2Stages; ++i) { ptx::mbar_init(BAR(i), 1); ptx::mbar_init(BAR(kEmpty + i), (uint32_t)csize); }
    ptx::mbar_init(BAR(kAccFull), 1);
    ptx::mbar_init(BAR(kAccEmpty), 4);
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

  // work items: CL = false: (pass, chunk) per CTA; CL = true: (channel half x kd group, chunk) per CLUSTER
  const int num_items = CL ? p.mhalves * p.kd_groups * p.chunks : p.num_passes * p.chunks;
  const int item0 = CL ? (int)blockIdx.x / csize : (int)blockIdx.x;
  const int item_step = CL ? (int)gridDim.x / csize : (int)gridDim.x;
  const int tiles_per_n = p.d * p.tiles_h * p.tiles_w;
  auto decode_pass = [&](int pass_slot, int& mh, int& g, int& jg, int& kw0, int& kw1) {
    int pg;
    if (CL) {  // pass_slot = channel half x kd group; the cluster rank picks (kh group, kw subset)
      g = pass_slot % p.kd_groups;
      mh = pass_slot / p.kd_groups;
      pg = (int)crank % p.passes_per_group;
      jg = (int)crank / p.passes_per_group;
    } else {
      const int pass = p.pass_order[pass_slot];
      pg = pass % p.passes_per_group;
      int r = pass / p.passes_per_group;
      jg = r % p.jgroups; r /= p.jgroups;
      g = r % p.kd_groups;
      mh = r / p.kd_groups;
    }
    kw0 = pg * p.units_per_pass;
    kw1 = min(p.kw_taps, kw0 + p.units_per_pass);
  };

  // item -> (pass slot, tile range).  rep: the two kd groups own gchunks[0] / gchunks[1] consecutive items
  auto decode_item = [&](int item, int& pass, int& t0, int& t1) {
    if (CL && p.rep) {
      const int grp = item < p.gchunks[0] ? 0 : 1;
      const int chunk = grp ? item - p.gchunks[0] : item;
      pass = grp;
      t0 = chunk * p.gtpc[grp];
      t1 = min(p.total_tiles, t0 + p.gtpc[grp]);
    } else {
      pass = item / p.chunks;
      t0 = (item % p.chunks) * p.tiles_per_chunk;
      t1 = min(p.total_tiles, t0 + p.tiles_per_chunk);
    }
  };
  const int num_items_all = (CL && p.rep) ? p.gchunks[0] + p.gchunks[1] : num_items;
  // k-th item of this CTA (cluster): static round-robin, or from the scheduler warp (umma.cuh, namespace sched)
  const bool dyn = sched::kEnabled && !CL && p.sched != nullptr;
  auto static_item = [&](uint32_t k) -> int {
    const int it = item0 + (int)k * item_step;
    return it < num_items_all ? it : -1;
  };
  if (warp == 3 && dyn && lane == 0) sched::run(BAR(kSF), BAR(kSE), sched_slots, p.sched, num_items_all);

  if (warp == 0) {
    if (lane == 0) {
      uint32_t use = 0;
      const int x_planes = p.cin_m / 8;
      for (uint32_t ik = 0;; ++ik) {
        const int item = dyn ? sched::next_lane(BAR(kSF), BAR(kSE), sched_slots, ik) : static_item(ik);
        if (item < 0) break;
        int pass, t0, t1;
        decode_item(item, pass, t0, t1);
        int mh, g, jg, kw0, kw1;
        decode_pass(pass, mh, g, jg, kw0, kw1);
        const bool repl = CL && p.rep && g == p.kd_groups - 1;
        const int planes_valid = repl ? p.qm : min(p.qm, 5 - g * p.qm);
        const uint32_t bytes =
            (uint32_t)(planes_valid * x_planes * kW2GroupBytes + (kW2TileH + 4) * p.dyp * kW2RowBytes);
        for (int t = t0; t < t1; ++t, ++use) {
          const int n = t / tiles_per_n;
          int r = t % tiles_per_n;
          const int tw = r % p.tiles_w; r /= p.tiles_w;
          const int th = r % p.tiles_h; const int d = r / p.tiles_h;
          const uint32_t s = use % kW2Stages, ph = (use / kW2Stages) & 1;
          ptx::mbar_wait(BAR(kEmpty + s), ph ^ 1);
          ptx::mbar_expect_tx(BAR(s), bytes);
          if (CL) {  // box i of the tile (planes_valid X planes + the dY tile) is issued by rank i % csize for everyone
            for (int q = 0; q < planes_valid; ++q)
              if ((uint32_t)(q % csize) == crank)
                ptx::tma_load_4d_mc(ptx::smem_u32(x_smem + s * kW2XBytes + q * x_planes * kW2GroupBytes), &tmap_x, BAR(s),
                                    (tw * kW2TileW - 2 + (repl ? q : 0)) * 8, th * kW2TileH,
                                    d + g * p.qm + (repl ? 0 : q) - 2, n * p.x_c8_total + mh * 16, cmask);
            if ((uint32_t)(planes_valid % csize) == crank)
              ptx::tma_load_4d_mc(ptx::smem_u32(dy_smem + s * p.dy_stage_bytes), &tmap_dy, BAR(s), tw * kW2TileW * 8,
                                  n * p.dy_c8_total, th * kW2TileH - 2, d, cmask);
          } else {
            for (int q = 0; q < planes_valid; ++q)
              ptx::tma_load_4d(ptx::smem_u32(x_smem + s * kW2XBytes + q * x_planes * kW2GroupBytes), &tmap_x, BAR(s),
                               (tw * kW2TileW - 2) * 8, th * kW2TileH, d + g * p.qm + q - 2,
                               n * p.x_c8_total + mh * 16);
            ptx::tma_load_4d(ptx::smem_u32(dy_smem + s * p.dy_stage_bytes), &tmap_dy, BAR(s), tw * kW2TileW * 8,
                             n * p.dy_c8_total, th * kW2TileH - 2, d);
          }
        }
      }
    }
  } else if (warp == 1) {
    // whole warp runs the uniform control flow / descriptor arithmetic; one elected lane issues (see conv_k5_umma.cu)
    const bool leader = ptx::elect_one();
    const uint32_t tmem_u = __reduce_or_sync(0xffffffffu, tmem_base);
    // N of a pass = the kh shifts its group really holds (the last group of a 3+2 split is narrower: 128 instead of 192
    // columns for 64 output channels - a third fewer tensor-pipe cycles for those passes)
    auto pass_n = [&](int jg) {
      const int jcount = min(p.jh, 5 - jg * p.jh);
      return (uint32_t)((jcount * 8 * p.dyp + 15) / 16 * 16);
    };
    constexpr uint32_t a_hi = ptx::desc_hi(kW2GroupBytes), b_hi = ptx::desc_hi(kW2RowBytes);
    const uint32_t b_row16 = (uint32_t)(p.dyp * kW2RowBytes) >> 4;  // one h row of the dY tile, 16-byte units
    const uint32_t npad = (uint32_t)p.npad;
    uint32_t use = 0, iuse = 0;
    for (;; ++iuse) {
      const int item = sched::uniform(dyn ? sched::next(BAR(kSF), BAR(kSE), sched_slots, iuse, lane) : static_item(iuse));
      if (item < 0) break;
      int pass, t0, t1;
      decode_item(item, pass, t0, t1);
      int mh, g, jg, kw0, kw1;
      decode_pass(pass, mh, g, jg, kw0, kw1);
      const int nkw = kw1 - kw0;
      const uint32_t idesc = ptx::make_idesc_bf16(128, (int)pass_n(jg), 1, 1);
      const bool repl = CL && p.rep && g == p.kd_groups - 1;
      ptx::mbar_wait(BAR(kAccEmpty), (iuse & 1) ^ 1);
      ptx::tc_fence_after();
      bool mid_started = false;
      for (int t = t0; t < t1; ++t, ++use) {
        const uint32_t s = use % kW2Stages, ph = (use / kW2Stages) & 1;
        ptx::mbar_wait(BAR(s), ph);
        ptx::tc_fence_after();
        const uint32_t a_lo0 = ptx::desc_lo(ptx::smem_u32(x_smem + s * kW2XBytes), 8u) + (uint32_t)(kw0 + p.kw_base);
        const uint32_t b_lo0 =
            ptx::desc_lo(ptx::smem_u32(dy_smem + s * p.dy_stage_bytes), 8u) + (uint32_t)(jg * p.jh) * b_row16;
        if (repl) {
          // M slot q holds the kd = 4 plane shifted by q voxels along w: viewed at +rank it is tap kw = q + rank.
          // rank 0: kw 0..3; rank 1: kw 1..4, of which only kw = 4 (slot 3) is kept by the epilogue
          const uint32_t a_base = ptx::desc_lo(ptx::smem_u32(x_smem + s * kW2XBytes), 8u) + crank;
#pragma unroll
          for (int u = 0; u < kW2TileH; ++u) {
            const uint32_t acc = (t != t0 || u != 0) ? 1u : 0u;
            if (leader)
              ptx::mma_bf16_split(tmem_u, a_base + (uint32_t)(u * (kW2TileW + 4)), a_hi, b_lo0 + (uint32_t)u * b_row16,
                                  b_hi, idesc, acc);
          }
        } else if (CL && p.balance) {
          // accumulator slots 0,1 = this rank's fixed taps (kw 0,1 or 3,4), slot 2 = the middle tap on the tiles it owns
          const uint32_t a_base = ptx::desc_lo(ptx::smem_u32(x_smem + s * kW2XBytes), 8u);
          const uint32_t kwf = crank == 0 ? 0u : 3u;
          const bool own_mid = (uint32_t)(t & 1) == crank;
#pragma unroll
          for (int u = 0; u < kW2TileH; ++u) {
            const uint32_t acc = (t != t0 || u != 0) ? 1u : 0u;
            const uint32_t b_lo = b_lo0 + (uint32_t)u * b_row16;
            const uint32_t a_row = a_base + (uint32_t)(u * (kW2TileW + 4));
            if (leader) {
              ptx::mma_bf16_split(tmem_u, a_row + kwf, a_hi, b_lo, b_hi, idesc, acc);
              ptx::mma_bf16_split(tmem_u + npad, a_row + kwf + 1u, a_hi, b_lo, b_hi, idesc, acc);
              if (own_mid)
                ptx::mma_bf16_split(tmem_u + 2u * npad, a_row + 2u, a_hi, b_lo, b_hi, idesc, (mid_started || u != 0) ? 1u : 0u);
            }
          }
          if (own_mid) mid_started = true;
        } else {
#pragma unroll
          for (int u = 0; u < kW2TileH; ++u) {
            const uint32_t acc = (t != t0 || u != 0) ? 1u : 0u;
            const uint32_t b_lo = b_lo0 + (uint32_t)u * b_row16;
#pragma unroll
            for (int k = 0; k < 5; ++k) {
              if (k < nkw && leader)
                ptx::mma_bf16_split(tmem_u + (uint32_t)k * npad, a_lo0 + (uint32_t)(u * (kW2TileW + 4) + k), a_hi, b_lo,
                                    b_hi, idesc, acc);
            }
          }
        }
        if (leader) {
          if (CL) ptx::mma_commit_mc(BAR(kEmpty + s), cmask);  // the stage is shared: free once every rank's MMAs retired
          else ptx::mma_commit(BAR(kEmpty + s));
        }
      }
      if (leader) ptx::mma_commit(BAR(kAccFull));
      __syncwarp();
    }
  } else if (warp >= 4) {
    const int q4 = warp - 4;
    const int row = q4 * 32 + lane;
    const int qplane = row / p.cin_m, ci_local = row % p.cin_m;
    const int cw = 8 * p.dyp;  // channels per stacked kh group
    uint32_t iuse = 0;
    for (;; ++iuse) {
      const int item = dyn ? sched::next(BAR(kSF), BAR(kSE), sched_slots, iuse, lane) : static_item(iuse);
      if (item < 0) break;
      int pass, t0_, t1_;
      decode_item(item, pass, t0_, t1_);
      int mh, g, jg, kw0, kw1;
      decode_pass(pass, mh, g, jg, kw0, kw1);
      const bool repl = CL && p.rep && g == p.kd_groups - 1;
      const int kd = repl ? 4 : g * p.qm + qplane;
      const int ci = mh * 128 + ci_local;
      // repl: row block q = tap kw = q + rank; rank 1 only contributes kw = 4 (kw 1..3 are rank 0's)
      const bool row_ok = (qplane < p.qm) && kd < 5 && ci < p.cin_real && (!repl || crank == 0 || qplane == p.qm - 1);
      ptx::mbar_wait(BAR(kAccFull), iuse & 1);
      ptx::tc_fence_after();
      const uint32_t t_base = tmem_base + ((uint32_t)(q4 * 32) << 16);
      const bool bal = CL && p.balance;
      const int nslots = repl ? 1 : (bal ? 3 : kw1 - kw0);
#pragma unroll 1
      for (int slot = 0; slot < nslots; ++slot) {
        const int kw = repl ? qplane + (int)crank : (bal ? (slot == 2 ? 2 : (crank == 0 ? 0 : 3) + slot) : kw0 + slot);
        const int jcount_e = min(p.jh, 5 - jg * p.jh);
        const int npass_e = (jcount_e * 8 * p.dyp + 15) / 16 * 16;
#pragma unroll 1
        for (int cb = 0; cb < npass_e / 16; ++cb) {
          float acc[16];
          ptx::tmem_ld16(t_base + slot * p.npad + cb * 16, acc);
          if (row_ok) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const int col = cb * 16 + j;
              const int jj = col / cw, co = col % cw;
              const int jh = jg * p.jh + jj;  // dY row shift: v_h = u_h - 2 + jh  <=>  kh = 4 - jh
              if (jj < p.jh && jh < 5 && co < p.cout_real) {
                const int tap = (kd * 5 + (4 - jh)) * p.kw_taps + kw;
                red_add_f32(p.ws + ((int64_t)tap * p.cout_real + co) * p.cin_real + ci, acc[j]);
              }
            }
          }
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(BAR(kAccEmpty));
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

int launch_wgrad_v2(const msb_tensor& x, const msb_tensor& dy, int cout, int cin, int n, msb_dim3 dims, float* ws,
                    cudaStream_t st, int kw_taps) {
  const int dyp = (cout + 7) / 8;
  if (dyp > dy.c / 8 || dyp > 10) return MSB_ERR_UNSUPPORTED;  // wider outputs: per-tap kernel (N = Cout already wide)
  // stack as many kh shifts per MMA as fit N <= 256, preferring the split of 5 with the least waste
  int jh = 256 / (8 * dyp);
  if (jh >= 5) jh = 5;
  else if (jh == 4) jh = 3;  // 3+2(+1 wasted) beats 4+1(+3 wasted)
  const int jgroups = (5 + jh - 1) / jh;
  const int npad = (jh * 8 * dyp + 15) / 16 * 16;
  const int64_t S = (int64_t)dims.d * dims.h * dims.w;
  Wg2Params p;
  p.n = n; p.cin_real = cin; p.cout_real = cout; p.dyp = dyp; p.npad = npad;
  p.jh = jh; p.jgroups = jgroups;
  p.d = dims.d; p.h = dims.h; p.w = dims.w;
  p.tiles_w = (dims.w + kW2TileW - 1) / kW2TileW;
  p.tiles_h = (dims.h + kW2TileH - 1) / kW2TileH;
  p.x_c8_total = (int)(x.n_stride / (S * 8));
  p.dy_c8_total = (int)(dy.n_stride / (S * 8));
  p.cin_m = x.c < 128 ? x.c : 128;
  p.mhalves = x.c > 128 ? x.c / 128 : 1;
  p.qm = 128 / p.cin_m;
  const int qeff = p.qm < 5 ? p.qm : 5;
  p.kd_groups = (5 + qeff - 1) / qeff;
  p.kw_taps = kw_taps; p.kw_base = kw_taps == 5 ? 0 : 2;
  int amax = 512 / npad;
  if (amax > kw_taps) amax = kw_taps;
  p.passes_per_group = (kw_taps + amax - 1) / amax;
  p.units_per_pass = (kw_taps + p.passes_per_group - 1) / p.passes_per_group;
  p.num_passes = p.mhalves * p.kd_groups * p.jgroups * p.passes_per_group;
  p.total_tiles = n * dims.d * p.tiles_h * p.tiles_w;
  if (p.num_passes > 64) return MSB_ERR_UNSUPPORTED;
  // items = passes x chunks must not exceed 2 items per CTA (a third item on a few CTAs would set the makespan)
  int chunks = (2 * kNumSMs) / p.num_passes;
  if (chunks > p.total_tiles) chunks = p.total_tiles;
  if (chunks < 1) chunks = 1;
  p.tiles_per_chunk = (p.total_tiles + chunks - 1) / chunks;
  p.chunks = (p.total_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  {  // heavy passes (more kw taps per tile) first
    int cost[64], n_order = 0;
    for (int pass = 0; pass < p.num_passes; ++pass) {
      const int pg = pass % p.passes_per_group;
      const int kw0 = pg * p.units_per_pass;
      int kw1 = kw0 + p.units_per_pass;
      if (kw1 > kw_taps) kw1 = kw_taps;
      cost[pass] = kw1 - kw0;
    }
    for (int c = kw_taps; c >= 0; --c)
      for (int pass = 0; pass < p.num_passes; ++pass)
        if (cost[pass] == c) p.pass_order[n_order++] = (unsigned char)pass;
  }
  int dy_rows = (kW2TileH + 4) * dyp;
  const int need = (kW2TileH - 1 + (jgroups - 1) * jh) * dyp + npad / 8;  // groups read past the halo rows
  if (need > dy_rows) dy_rows = need;
  p.dy_stage_bytes = (dy_rows * kW2RowBytes + 1023) / 1024 * 1024;
  p.ws = ws;
  // ring depth: 4 stages when they fit (32 channels: 52 KB per stage) - the kw-replicated leftover group spends only
  // 8 MMAs on a tile, so three tiles in flight did not cover the TMA round trip; 3 otherwise (msb_debug_set(6, 32): 3)
  p.stages = kW2MaxStages;
  if ((g_debug_flags[6] & 32) || p.stages * (p.dy_stage_bytes + kW2XBytes) + 1024 + 128 > 227 * 1024) p.stages = 3;
  const int smem_bytes = p.stages * (p.dy_stage_bytes + kW2XBytes) + 1024 + 128;
  if (smem_bytes > 227 * 1024) return MSB_ERR_UNSUPPORTED;
  CUtensorMap tmx, tmdy;
  int rc;
  if ((rc = make_b8_tmap(&tmx, x, n, dims, kW2TileW + 4, kW2TileH, 1, p.cin_m / 8))) return rc;
  if ((rc = make_b8_tmap_hmajor(&tmdy, dy, n, dims, kW2TileW, dyp, kW2TileH + 4, 1))) return rc;
  // cluster path: the jgroups x passes_per_group passes of one (channel half, kd group) share their tiles.
  // MEASURED (B200, batch 2): 2-CTA clusters with the middle kw tap alternating between the ranks (`balance`): 32 -> 32
  // @128^3 1.11 ms vs 1.26 ms unclustered, @64^3 0.167 vs 0.199 ms -> ON by default for that shape family.  Unbalanced
  // clusters are a NEGATIVE result: 32 -> 32 1.24 ms (ranks run in lock-step at the pace of the rank with 3 of the 5 kw
  // taps), 64 -> 64 @64^3 with 6-CTA clusters 1.20 ms vs 0.65 ms -> OFF unless msb_debug_set(6, 4) forces them (the GPU
  // tests keep that path verified); msb_debug_set(6, 2) disables clustering altogether.
  int csize = p.jgroups * p.passes_per_group;
  const bool balanced_shape = csize == 2 && p.passes_per_group == 2 && kw_taps == 5;
  if ((g_debug_flags[6] & 2) || !(balanced_shape || (g_debug_flags[6] & 4))) csize = 1;
  const int groups = p.mhalves * p.kd_groups;
  if (csize >= 2 && csize <= 8 && (kNumSMs / csize) >= groups && p.total_tiles >= 8 * (kNumSMs / csize)) {
    const int nclusters = kNumSMs / csize;
    int chunks = nclusters / groups;  // one (group, chunk) item per cluster
    if (chunks > p.total_tiles) chunks = p.total_tiles;
    p.tiles_per_chunk = (p.total_tiles + chunks - 1) / chunks;
    p.chunks = (p.total_tiles + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.csize = csize;
    p.sched = nullptr;
    p.balance = (csize == 2 && p.passes_per_group == 2 && p.jgroups == 1 && kw_taps == 5 && p.tiles_per_chunk >= 4) ? 1 : 0;
    // kw-replicated leftover plane (32 input channels: 4 planes per M block, kd groups {0..3}, {4}): group 1 then costs
    // 1 MMA per row and rank against 2.5 for group 0, and the clusters are shared out in that proportion
    p.rep = (p.balance && p.qm == 4 && p.kd_groups == 2 && p.mhalves == 1 && !(g_debug_flags[6] & 16) &&
             p.total_tiles >= 16384) ? 1 : 0;
    if (p.rep) {
      // MMA cost 2.5 : 1 per row and rank would give the leftover group 2/7 of the clusters, but it streams the same
      // bytes per tile for 2.5x fewer MMAs and is bound by the L2 -> SM delivery instead.  MEASURED (B200, 32 -> 32
      // @128^3, batch 2): share 0.23 -> 1.74 ms, 2/7 -> 1.44, 0.35 -> 1.19, 0.42 -> 1.01 ms (round-1 form: 1.10 ms);
      // the two groups balance near 0.43.  At 64^3 the replica form is slower (0.198 vs 0.172 ms) -> large volumes only.
      // msb_debug_set(0, permille) overrides the share.
      const double share = g_debug_flags[0] > 0 ? g_debug_flags[0] / 1000.0 : 0.43;
      int c1 = (int)(nclusters * share + 0.5);
      if (c1 < 1) c1 = 1;
      if (c1 > nclusters - 1) c1 = nclusters - 1;
      const int c0 = nclusters - c1;
      const int cc[2] = {c0, c1};
      for (int gi = 0; gi < 2; ++gi) {
        p.gtpc[gi] = (p.total_tiles + cc[gi] - 1) / cc[gi];
        p.gchunks[gi] = (p.total_tiles + p.gtpc[gi] - 1) / p.gtpc[gi];
      }
    } else {
      p.gchunks[0] = p.gchunks[1] = p.gtpc[0] = p.gtpc[1] = 0;
    }
    const int items = p.rep ? p.gchunks[0] + p.gchunks[1] : groups * p.chunks;
    const int grid = (items < nclusters ? items : nclusters) * csize;
    static bool attr_set_cl = false;
    if (!attr_set_cl) {
      MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_wgrad2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      attr_set_cl = true;
    }
    cudaError_t e = launch_pdl_cluster(conv_k5_wgrad2_kernel<true>, dim3(grid), dim3(256), smem_bytes, st, csize, tmx,
                                       tmdy, p);
    if (e != cudaSuccess) {
      set_error("clustered wgrad launch failed: %s", cudaGetErrorString(e));
      return MSB_ERR_CUDA;
    }
    return MSB_OK;
  }
  p.csize = 1;
  p.balance = 0;
  p.rep = 0;
  p.sched = g_dynamic_tiles_wg2 ? g_sched_wg2_ptr : nullptr;
  p.gchunks[0] = p.gchunks[1] = p.gtpc[0] = p.gtpc[1] = 0;
  const int items = p.num_passes * p.chunks;
  const int grid = items < kNumSMs ? items : kNumSMs;
  static bool attr_set = false;
  if (!attr_set) {
    MSB_CUDA_OK(cudaFuncSetAttribute(conv_k5_wgrad2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  MSB_LAUNCH_PDL(conv_k5_wgrad2_kernel<false>, dim3(grid), dim3(256), smem_bytes, st, tmx, tmdy, p);
  return MSB_OK;
}

}  // namespace msb
