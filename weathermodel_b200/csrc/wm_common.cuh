// wm_common.cuh -- sm_100a device primitives shared by every kernel of the WeatherModel hot path.
//
// Thin inline-PTX wrappers (mbarrier, TMA, tcgen05/TMEM), the UMMA descriptor encoders, a
// counter-based Philox4x32-10 (bit-compatible with curand, which ATen uses for torch.rand on
// CUDA: torch/include/ATen/native/cuda/DistributionTemplates.h:66-89) and small math helpers.
// Nothing here allocates or holds global state except the device-side error word g_wm_dev_error.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace wm {

// ---------------------------------------------------------------------------------------------
// device-side error word: a bounded mbarrier wait that times out records a code here instead of
// hanging the GPU (a hang on the shared box is a strike). Host reads it via wm_device_error().
// ---------------------------------------------------------------------------------------------
// (the library is built as ONE translation unit -- wm_lib.cu -- so this is the single definition)
__device__ unsigned int g_wm_dev_error = 0;

// host-side count of kernel launches issued by this library (bench.py reports it as gpu_launches)
static long long g_wm_launches = 0;
#define WM_COUNT_LAUNCH() (++::wm::g_wm_launches)

#define WM_DEVICE __device__ __forceinline__

// phase timestamps of CTA 0 (debug aid for the warp-specialised kernels; read with wm_debug_ticks)
__device__ long long g_wm_ticks[64];
#define WM_TICK(i)                                                    \
  do {                                                                \
    if (blockIdx.x == 0 && (threadIdx.x & 31) == 0) g_wm_ticks[(i)] = clock64(); \
  } while (0)

WM_DEVICE uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Round a pointer into the dynamic shared-memory array up to `align` bytes WITHOUT leaving the shared address
// space: an integer offset added to the __shared__ array. Rounding through uintptr_t makes the compiler forget the
// address space -- every access through the result became a generic LD.E / ST.E instead of LDS / STS (found in the
// ncu source view of attn_bwd: the top stall was the scoreboard of those generic loads).
WM_DEVICE uint8_t* smem_align_up(uint8_t* smem_raw, uint32_t align) {
  const uint32_t a = smem_u32(smem_raw);
  return smem_raw + (((a + align - 1u) & ~(align - 1u)) - a);
}

WM_DEVICE uint32_t lane_id() { return threadIdx.x & 31u; }

WM_DEVICE bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t"
      ".reg .b32 %%rx;\n\t"
      ".reg .pred %%px;\n\t"
      "elect.sync %%rx|%%px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, %%px;\n\t"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
WM_DEVICE void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
WM_DEVICE void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
// generic-proxy smem writes -> visible to the async proxy (TMA / tcgen05.mma operand reads)
WM_DEVICE void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
WM_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
WM_DEVICE void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
WM_DEVICE bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: ~seconds at most, then flag + fall through (results are garbage, GPU survives).
WM_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity, uint32_t code = 1) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if ((++spins & 255u) == 0) {
      // ~0.25 s at 1.9 GHz for the first waiter; everyone else bails as soon as the flag is up
      if (*reinterpret_cast<volatile unsigned int*>(&g_wm_dev_error) != 0) break;
      if (clock64() - t0 > 500000000LL) {
        atomicCAS(&g_wm_dev_error, 0u, code);
        break;
      }
    }
  }
}

// NOTE on waiting: all 32 lanes of a waiting warp call mbar_wait. Letting one lane poll (behind `if (lane == 0)` +
// __syncwarp, or an elect.sync inside an asm block) was measured 1.5-2.5x SLOWER end to end in the attention kernels:
// a lone try_wait wakes up late. What does pay off is not waiting at all where in-order MMA completion already
// implies the condition.

// ---------------------------------------------------------------------------------------------
// TMA (cp.async.bulk.tensor) -- 2D tiled loads into 128B-swizzled smem, completion on an mbarrier
// ---------------------------------------------------------------------------------------------
WM_DEVICE void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
WM_DEVICE void tma_load_2d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}

WM_DEVICE void tma_load_3d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
WM_DEVICE void tma_load_4d(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// 1D bulk copy global -> shared (16-byte aligned source / destination, size a multiple of 16)
WM_DEVICE void bulk_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05 / TMEM
// ---------------------------------------------------------------------------------------------
template <uint32_t kCols>
WM_DEVICE void tmem_alloc(uint32_t* smem_slot) {  // whole warp, .sync.aligned
  static_assert(kCols >= 32 && kCols <= 512 && (kCols & (kCols - 1)) == 0, "TMEM cols: pow2 in [32,512]");
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(
                   smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
WM_DEVICE void tmem_dealloc(uint32_t taddr) {  // same warp that allocated
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols)
               : "memory");
}
WM_DEVICE void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
WM_DEVICE void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// D[tmem] (+)= A[smem desc] * B[smem desc]; issued by ONE thread for the whole CTA.
WM_DEVICE void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                       uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once every previously issued tcgen05.mma of this thread has completed
// (implies tcgen05.fence::before_thread_sync)
WM_DEVICE void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// Warp-convergent variants: EVERY lane of the warp executes the call (uniform control flow, so the compiler keeps
// descriptor arithmetic in uniform registers and emits no per-active-lane issue loop); one elected lane issues.
WM_DEVICE void umma_ss_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
WM_DEVICE void umma_commit_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
      "}\n" ::"r"(smem_u32(bar))
      : "memory");
}
// A operand from tensor memory (128 lanes x 8 columns of packed bf16x2 per k-step of 16), B from shared memory
WM_DEVICE void umma_ts_warp(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// registers -> TMEM: lane t of the warp writes 16 consecutive 32-bit columns of TMEM lane (lane_base + t)
WM_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};\n" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
WM_DEVICE void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
WM_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;\n" ::: "memory"); }
// warp index that the compiler can prove warp-uniform (role dispatch on it stays convergent)
WM_DEVICE int warp_idx_uniform() { return __shfl_sync(0xffffffffu, static_cast<int>(threadIdx.x >> 5), 0); }

// ---- cta_group::2 (CTA pair) variants: one MMA spans two SMs (M = 256), operands are split between the two
// CTAs' shared memories, the accumulator rows between their TMEMs. Only the leader CTA (cluster rank 0) issues
// MMAs and commits; both CTAs issue TMA loads that signal the LEADER's mbarrier. ----
WM_DEVICE uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
WM_DEVICE void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta pointer of this CTA) as seen in CTA `rank` of the cluster
WM_DEVICE uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t a;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(a) : "r"(smem_u32(p)), "r"(rank));
  return a;
}
// Arrive on a barrier in another CTA of the cluster. Default semantics (.release at CTA scope), as CUTLASS'
// ClusterBarrier::arrive uses: what this hands over is a TMEM accumulator stage, ordered by the tcgen05 fences around
// it -- no global memory. The explicit `.release.cluster` form compiled to MEMBAR.ALL.GPU + ERRBAR in front of every
// arrive: 46 % of the stall samples of the CTA-pair GEMM's epilogue warps (ncu source view, profiles/r02_gemm_membar.txt).
WM_DEVICE void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];\n" ::"r"(cluster_addr) : "memory");
}
template <uint32_t kCols>
WM_DEVICE void tmem_alloc_2cta(uint32_t* smem_slot) {  // one warp in EACH CTA of the pair
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_slot)),
               "n"(kCols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;\n" ::: "memory");
}
template <uint32_t kCols>
WM_DEVICE void tmem_dealloc_2cta(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols) : "memory");
}
WM_DEVICE void umma_ss_2cta(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive (once all earlier MMAs of this thread are done) on the barrier at this smem offset in BOTH CTAs
WM_DEVICE void umma_commit_2cta(uint64_t* bar) {
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n" ::"r"(
          smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}
// 2-CTA TMA load: data lands in THIS CTA's smem, completion bytes are credited to the leader's mbarrier
// (peer bit 24 of the shared::cluster address cleared, as cute::SM100_TMA_2SM_LOAD does)
WM_DEVICE void tma_load_2d_2cta(void* smem_dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}

// warp-convergent cta_group::2 issue (see umma_ss_warp): every lane of the leader CTA's issue warp calls these
WM_DEVICE void umma_ss_2cta_warp(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, e;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|e, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@e tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
WM_DEVICE void umma_commit_2cta_warp(uint64_t* bar) {
  asm volatile(
      "{\n\t"
      ".reg .pred e;\n\t"
      ".reg .b32 r;\n\t"
      "elect.sync r|e, 0xffffffff;\n\t"
      "@e tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;\n\t"
      "}\n" ::"r"(smem_u32(bar)),
      "h"(static_cast<uint16_t>(3))
      : "memory");
}

// TMA store: one box of a 2D tiled tensor map from shared memory to global memory (bulk async-group of the
// issuing thread). Elements outside the tensor's bounds are not written.
WM_DEVICE void tma_store_2d(const CUtensorMap* m, const void* smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_u32(smem_src)), "r"(c0), "r"(c1)
               : "memory");
}
WM_DEVICE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;\n" ::: "memory"); }
// wait until at most kPending of this thread's bulk groups still have to READ their shared-memory source
template <int kPending>
WM_DEVICE void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;\n" ::"n"(kPending) : "memory");
}

// TMEM -> registers: lane t of the warp receives 32 / 16 consecutive fp32 columns of TMEM lane
// (lane_base + t); lane_base must be 32*(warp_id % 4).
WM_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
WM_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
WM_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// UMMA descriptors (bit layout: cute/arch/mma_sm100_desc.hpp SmemDescriptor / InstrDescriptor)
// ---------------------------------------------------------------------------------------------
enum : uint32_t { UMMA_SWZ_NONE = 0, UMMA_SWZ_128B = 2, UMMA_SWZ_64B = 4, UMMA_SWZ_32B = 6 };

// descriptor of the same layout `byte_off` bytes further into shared memory (start-address field only; the
// caller guarantees the sum stays below 256 KB so the 14-bit field cannot carry)
WM_DEVICE uint64_t umma_desc_advance(uint64_t desc, uint32_t byte_off) {
  return desc + static_cast<uint64_t>(byte_off >> 4);
}
WM_DEVICE uint64_t umma_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                  uint32_t layout_type) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3FFFFu) >> 4);          // [0,14)  start address
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;     // [16,30) leading byte offset
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;     // [32,46) stride byte offset
  d |= static_cast<uint64_t>(1) << 46;                              // [46,48) version = 1 (sm_100)
  d |= static_cast<uint64_t>(layout_type & 7u) << 61;               // [61,64) swizzle mode
  return d;
}
// kind::f16, bf16 x bf16 -> fp32; a_mn / b_mn = 1 selects an MN-major (transposed) operand.
__host__ __device__ inline uint32_t umma_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn,
                                                    uint32_t b_mn) {
  uint32_t d = 0;
  d |= 1u << 4;            // c_format = F32
  d |= 1u << 7;            // a_format = BF16
  d |= 1u << 10;           // b_format = BF16
  d |= (a_mn & 1u) << 15;  // a_major
  d |= (b_mn & 1u) << 16;  // b_major
  d |= ((n >> 3) & 0x3Fu) << 17;
  d |= ((m >> 4) & 0x1Fu) << 24;
  return d;
}

// ---------------------------------------------------------------------------------------------
// Philox4x32-10, curand-compatible. counter = {offset_lo, offset_hi, subseq_lo, subseq_hi}
// (curand_init(seed, subsequence, offset) with offset counted in 128-bit blocks here).
// ---------------------------------------------------------------------------------------------
struct Philox4 {
  uint32_t x, y, z, w;
};
template <int kRounds>
__host__ __device__ inline Philox4 philox4x32(uint64_t seed, uint64_t subsequence, uint64_t block_offset) {
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(seed >> 32);
  uint32_t c0 = static_cast<uint32_t>(block_offset), c1 = static_cast<uint32_t>(block_offset >> 32);
  uint32_t c2 = static_cast<uint32_t>(subsequence), c3 = static_cast<uint32_t>(subsequence >> 32);
#pragma unroll
  for (int r = 0; r < kRounds; ++r) {
#ifdef __CUDA_ARCH__
    uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
#else
    uint64_t p0 = static_cast<uint64_t>(0xD2511F53u) * c0, p1 = static_cast<uint64_t>(0xCD9E8D57u) * c2;
    uint32_t hi0 = static_cast<uint32_t>(p0 >> 32), lo0 = static_cast<uint32_t>(p0);
    uint32_t hi1 = static_cast<uint32_t>(p1 >> 32), lo1 = static_cast<uint32_t>(p1);
#endif
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}
__host__ __device__ inline Philox4 philox4x32_10(uint64_t seed, uint64_t subsequence, uint64_t block_offset) {
  return philox4x32<10>(seed, subsequence, block_offset);
}
// curand_uniform: (0,1]
__host__ __device__ inline float curand_uniform_from_u32(uint32_t x) {
  return static_cast<float>(x) * 2.3283064e-10f + (2.3283064e-10f / 2.0f);
}

// Dropout masks. Nothing here has to reproduce torch's stream (only the INPUT masks replay curand Philox, above), so
// the keep decisions come from a much cheaper counter-based hash: rounds of "multiply 32 x 32 -> 64, fold the
// halves" keyed by words derived from (seed, stream). Philox4x32-7 cost ~12 issue slots per element in the
// attention forward kernel -- more than the softmax itself (ncu instruction mix, profiles/r01_attn_ncu_mix.txt).
// One counter x decides FOUR elements with 15-bit resolution each: the first round gives t(x); two second-round
// products of t under different keys give two well-mixed 32-bit words A, B; element 0 / 1 = low / high half of A,
// element 2 / 3 = low / high half of B. keep iff (half & 0x7FFF) >= thresh15, thresh15 = round(32768 p): the drop
// probability is thresh15 / 32768 (0.100006 for p = 0.1, the reference's nn.Dropout(0.1); round 1 compared 7 bits
// and trained at 0.1016) and the keep scale 32768 / (32768 - thresh15). Measured on 2^18 consecutive counters
// (tests/test_oracle.py): every input bit flips every output bit with probability 0.48-0.52, keep decisions of
// neighbouring elements / words / streams correlate below 6e-3.
// Element e of a row of N elements: counter (row * ceil(N/16) + e/16) * 4 + (e%16)/4, element e%4 of it -- the GEMM
// epilogue, LayerNorm backward and the attention kernels all use this numbering, so a mask can be regenerated anywhere.
struct DropKeys {
  uint32_t k0, k1, k2;
};
__host__ __device__ inline DropKeys drop_keys(uint64_t seed, uint64_t stream) {  // splitmix64 finaliser
  uint64_t z = seed + (stream + 1) * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z ^= z >> 31;
  const uint32_t a = static_cast<uint32_t>(z), b = static_cast<uint32_t>(z >> 32);
  return DropKeys{a, b, (b ^ 0x68E31DA4u) + a * 0x2545F491u};
}
// Per-replay key mix for CUDA-graph replays of a training step: the kernels of a captured step carry the dropout keys
// of the step that was captured as launch parameters; a captured 1-thread kernel (wm_step_params_apply) installs three
// fresh words here at the head of every replay and every dropout site folds them into its keys, so replays draw new
// masks (forward and backward of one replay see the same words). All zero -- the state outside graphs -- changes nothing.
__device__ uint32_t g_wm_drop_mix[4] = {0u, 0u, 0u, 0u};
WM_DEVICE DropKeys drop_keys_live(DropKeys k) {
  k.k0 ^= g_wm_drop_mix[0];
  k.k1 ^= g_wm_drop_mix[1];
  k.k2 ^= g_wm_drop_mix[2];
  return k;
}
struct DropWords {
  uint32_t a, b;  // a: elements 0 (bits 0-15) and 1 (bits 16-31) of the counter; b: elements 2 and 3
};
__host__ __device__ inline DropWords drop_hash64(uint32_t x, DropKeys k) {
  const uint64_t p = static_cast<uint64_t>(x ^ k.k0) * 0x9E3779B1u;
  const uint32_t hl = static_cast<uint32_t>(p >> 32) ^ static_cast<uint32_t>(p);
  const uint64_t q = static_cast<uint64_t>(hl ^ k.k1) * 0x85EBCA77u;
  const uint64_t r = static_cast<uint64_t>(hl ^ k.k2) * 0xC2B2AE3Du;
  return DropWords{static_cast<uint32_t>(q >> 32) ^ static_cast<uint32_t>(q),
                   static_cast<uint32_t>(r >> 32) ^ static_cast<uint32_t>(r)};
}
__host__ __device__ inline uint32_t drop_thresh15(uint32_t thresh16) { return (thresh16 + 1u) >> 1; }
__host__ __device__ inline float drop_keep_scale(uint32_t thresh16) {
  const uint32_t t = drop_thresh15(thresh16);
  return t ? 32768.0f / static_cast<float>(32768u - t) : 1.0f;
}
// per-half addend of the keep test: (half & 0x7FFF) + (0x8000 - thresh15) carries into bit 15 exactly when the
// 15-bit value is >= thresh15, and no carry crosses a half
__host__ __device__ inline uint32_t drop_add2(uint32_t thresh16) { return (0x8000u - drop_thresh15(thresh16)) * 0x00010001u; }
// bit 15 / 31 of .a = keep flags of elements 0 / 1, of .b = elements 2 / 3 (the other bits are noise)
WM_DEVICE DropWords drop_flags4(uint32_t x, DropKeys k, uint32_t add2) {
  const DropWords h = drop_hash64(x, k);
  return DropWords{(h.a & 0x7FFF7FFFu) + add2, (h.b & 0x7FFF7FFFu) + add2};
}
// all-ones / all-zeros 32-bit mask from the flag of element E (0..3) of the counter: one PRMT in sign-replicate mode
template <int E>
WM_DEVICE uint32_t drop_mask32(const DropWords& f) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"((E & 2) ? f.b : f.a), "r"(0u), "r"((((E & 1) ? 3u : 1u) | 8u) * 0x1111u));
  return d;
}
// 0xFFFF / 0x0000 in each half of the result: the keep flags of the two elements of one flag word, for masking a
// packed bf16x2 pair
WM_DEVICE uint32_t drop_pair_mask(uint32_t fword) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(fword), "r"(0u), "r"(0xBB99u));
  return d;
}

// ---------------------------------------------------------------------------------------------
// misc
// ---------------------------------------------------------------------------------------------
WM_DEVICE float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
WM_DEVICE float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// 256-bit global accesses (sm_100+): one full 32-byte sector per lane. Thread-per-row epilogues are bound by the
// number of sectors the LSU touches (~1 per cycle), so 16-byte accesses that each cover half a sector cost twice.
WM_DEVICE void ldg256(const void* p, uint4& lo, uint4& hi) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(lo.x), "=r"(lo.y), "=r"(lo.z), "=r"(lo.w), "=r"(hi.x), "=r"(hi.y), "=r"(hi.z), "=r"(hi.w)
               : "l"(p));
}
WM_DEVICE void stg256(void* p, const uint4& lo, const uint4& hi) {
  asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(lo.x), "r"(lo.y), "r"(lo.z),
               "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w)
               : "memory");
}
// Packed fp32 pairs (sm_100: FFMA2 / FADD2 / FMUL2 work on an aligned register pair, one issue slot for two
// elements) and the three-input maximum (FMNMX3). The softmax loops of the attention kernels are issue-bound, so
// halving the slots of their FMAs / adds is a direct gain.
WM_DEVICE uint64_t f2_pack(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
WM_DEVICE uint64_t f2_pack_u(uint32_t lo, uint32_t hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
WM_DEVICE void f2_unpack(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
WM_DEVICE uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
WM_DEVICE uint64_t f2_add(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
WM_DEVICE uint64_t f2_mul(uint64_t a, uint64_t b) {
  uint64_t d;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
WM_DEVICE float fmax3(float a, float b, float c) {
  float d;
  asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}
WM_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
WM_DEVICE float bf16_lo(uint32_t v) { return __uint_as_float(v << 16); }
WM_DEVICE float bf16_hi(uint32_t v) { return __uint_as_float(v & 0xFFFF0000u); }

}  // namespace wm
