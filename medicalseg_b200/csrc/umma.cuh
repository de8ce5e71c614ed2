// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc / mma / commit / ld).
#pragma once
#include <cuda.h>
#include <stdint.h>

namespace msb {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---- mbarrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// ---- TMA --------------------------------------------------------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                            int c3, int c4) {
  asm volatile(
      "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
      : "memory");
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(bar)
               : "memory");
}

// ---- thread-block clusters: multicast TMA + multicast MMA-completion arrives -------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// the box lands at the same shared-memory offset in every CTA of `mask`, and each copy signals the mbarrier at the
// same offset of ITS CTA with its own byte count
__device__ __forceinline__ void tma_load_4d_mc(uint32_t dst, const CUtensorMap* m, uint32_t bar, int c0, int c1, int c2,
                                               int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(dst), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}

// ---- tcgen05 ----------------------------------------------------------------------------------------
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(kCols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 x bf16 -> f32
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// mbarrier arrives once every previously issued tcgen05.mma of this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// same, arriving on the mbarrier at this offset in EVERY CTA of `mask` (a shared stage is free once all consumers are done)
__device__ __forceinline__ void mma_commit_mc(uint32_t bar, uint16_t mask) {
  asm volatile(
      "tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"(mask)
      : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- descriptors --------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, SWIZZLE_NONE ("interleave") canonical layouts (16-byte units):
//   K-major : ((8,m),(8 elems,2)) : ((16B, SBO),(1, LBO))  -> 8 rows x 16 B core matrix, contiguous 128 B
//   MN-major: ((8 elems,m),(8,k)) : ((1, SBO),(16B, LBO))  -> 8 k-rows x 16 B core matrix, contiguous 128 B
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}

// The same descriptor split into 32-bit halves: only the low word (start address, LBO) changes between MMAs, so issue
// loops add plain 32-bit offsets (16-byte units) to `lo` and keep `hi` (SBO, version, SWIZZLE_NONE) constant.
__host__ __device__ constexpr uint32_t desc_hi(uint32_t sbo_bytes) { return ((sbo_bytes >> 4) & 0x3fffu) | (1u << 14); }
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo16) {
  return ((smem_addr >> 4) & 0x3fffu) | ((lbo16 & 0x3fffu) << 16);
}
__device__ __forceinline__ void mma_bf16_split(uint32_t d_tmem, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                               uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %6, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\tmov.b64 db, {%3, %4};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}\n" ::"r"(d_tmem),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate)
      : "memory");
}

// Instruction descriptor for kind::f16 with bf16 inputs and f32 accumulation.
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
         ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

}  // namespace ptx

// fire-and-forget f32 add (RED): the weight-gradient epilogues issue tens of thousands of these per CTA.  Written as
// PTX because `atomicAdd` with an unused result is only turned into RED by a compiler heuristic - the kernels that also
// contain value-returning atomics (the tile scheduler's counter) got ATOMG for all of them and ran 35-45 % slower.
__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

// ---- dynamic tile scheduler for the persistent tensor-core kernels ------------------------------------------------
// Static round-robin (item = blockIdx.x + k * gridDim.x) makes the makespan of a persistent grid the finish time of
// its LAST-STARTED CTA: when NCCL's all-reduce CTAs hold some SMs at launch (data-parallel step),
// the CTAs that wanted those SMs start late and still have their full share of tiles to do.  With the scheduler one
// otherwise idle warp fetches item numbers from a global atomic counter and hands them to the role warps through a
// small shared-memory ring (mbarrier full / empty pairs), so late CTAs simply take fewer tiles.  The first item stays
// static (no atomic latency before the first TMA), the counter pair {next, done} resets itself: the CTA that finishes
// last writes zeros (every CTA's final fetch precedes its `done` increment).
#ifndef MSB_DYNAMIC_TILES
#define MSB_DYNAMIC_TILES 0
#endif
namespace sched {
// Compiled OUT by default (build with -DMSB_DYNAMIC_TILES=1 to get it back): measured on 2 and 8 B200s it gains nothing
// (DESIGN 6), and merely carrying the branch cost the static path 3.5 % on the forward kernel (the ring's barriers, and
// the BatchNorm-statistics REDG turned into ATOMG next to the counter's value-returning atomics).
constexpr bool kEnabled = MSB_DYNAMIC_TILES != 0;
constexpr int kDepth = 4;

// REDUX keeps an item number that came out of the ring provably warp-uniform (everything the MMA issue loop derives
// from it must live in uniform registers: without this the loop costs ~11 extra SASS instructions per MMA, 35-45 % on
// the per-tap weight-gradient kernels).  The static round-robin number is built from blockIdx / gridDim and needs none.
__device__ __forceinline__ int uniform(int v) {
  return kEnabled ? (int)__reduce_or_sync(0xffffffffu, (unsigned int)v) : v;
}

__device__ __forceinline__ void init(uint32_t full0, uint32_t empty0, uint32_t consumer_warps) {  // one thread
  for (int i = 0; i < kDepth; ++i) {
    ptx::mbar_init(full0 + 8u * i, 1);
    ptx::mbar_init(empty0 + 8u * i, consumer_warps);
  }
}
// one lane of the scheduler warp
__device__ __forceinline__ void run(uint32_t full0, uint32_t empty0, volatile int* slots, unsigned int* counters,
                                    int num_items) {
  const int grid = (int)gridDim.x;
  for (uint32_t k = 0;; ++k) {
    const uint32_t s = k % kDepth, ph = (k / kDepth) & 1u;
    ptx::mbar_wait(empty0 + 8u * s, ph ^ 1u);
    int item = k == 0 ? (int)blockIdx.x : grid + (int)atomicAdd(counters, 1u);
    if (item >= num_items) item = -1;
    slots[s] = item;
    ptx::mbar_arrive(full0 + 8u * s);  // release: the slot write is visible to the waiters
    if (item < 0) break;
  }
  __threadfence();
  if (atomicAdd(counters + 1, 1u) == (unsigned int)grid - 1u) {  // last CTA: nobody fetches any more
    atomicExch(counters, 0u);
    atomicExch(counters + 1, 0u);
  }
}
// every lane of a consumer warp; returns -1 when the work is exhausted
__device__ __forceinline__ int next(uint32_t full0, uint32_t empty0, const volatile int* slots, uint32_t k, int lane) {
  const uint32_t s = k % kDepth, ph = (k / kDepth) & 1u;
  ptx::mbar_wait(full0 + 8u * s, ph);
  const int item = slots[s];
  __syncwarp();
  if (lane == 0) ptx::mbar_arrive(empty0 + 8u * s);
  return item;
}
// the same for a role that runs on a single thread (TMA producers written under `if (lane == 0)`)
__device__ __forceinline__ int next_lane(uint32_t full0, uint32_t empty0, const volatile int* slots, uint32_t k) {
  const uint32_t s = k % kDepth, ph = (k / kDepth) & 1u;
  ptx::mbar_wait(full0 + 8u * s, ph);
  const int item = slots[s];
  ptx::mbar_arrive(empty0 + 8u * s);
  return item;
}
}  // namespace sched
}  // namespace msb
