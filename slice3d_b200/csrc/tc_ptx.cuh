// Thin inline-PTX wrappers for the Blackwell (sm_100a) primitives the tensor-core decoder uses:
// mbarrier, bulk async copy (TMA engine, UBLKCP), tcgen05 MMA / TMEM load-store / commit / alloc,
// and the UMMA shared-memory + instruction descriptors.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <stdint.h>

namespace s3d {
namespace ptx {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Suspend-time hint of try_wait: the waiting warp stays suspended (no issue slots, no spin power) until the phase
// completes or this many ns pass, instead of the short system default (an ncu capture showed 26 % of all issued
// instructions were try_wait spins).
constexpr uint32_t MBAR_SUSPEND_NS = 0x989680u;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// ---------------------------------------------------------------- cluster (CTA pair) helpers
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {  // every thread of every CTA of the cluster
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cta address -> shared::cluster address of the same offset in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// Arrive on a barrier of another CTA of the cluster.  Default semantics (release at CTA scope), as CUTLASS's
// ClusterBarrier::arrive(cta_id) does: what the arrive publishes here lives in tensor / shared memory and is ordered
// by tcgen05.fence / fence.proxy.async; an explicit .release.cluster was measured at ~1 kcycle per arrive.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {  // (same as mbar_try_wait, see above)
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait_cluster(bar, parity)) {
  }
}

// Register re-budgeting between warpgroups (every warp of the warpgroup executes the same instruction).
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// One lane of a converged warp (warp-uniform control flow around tcgen05.mma / bulk copies).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() {  // generic-proxy smem writes -> async proxy (UMMA/TMA)
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ---------------------------------------------------------------- bulk async copy (global -> shared)
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// ---------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// CTA-pair variants: the same warp of BOTH CTAs of the pair executes them with the same arguments.
__device__ __forceinline__ void tmem_alloc_pair(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// 32 lanes x 32 consecutive 32-bit columns: thread i of the warp reads lane (addr.lane + i).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float* v) {
  const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------- UMMA
// Shared-memory matrix descriptor, K-major operand stored as [rows][64 bf16] tiles with the
// 128-byte swizzle (rows 128 B apart, 8-row groups 1024 B apart; tile base 1024-B aligned).
// Fields (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version=1 [46,48), layout SWIZZLE_128B=2 [61,64).
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>(1) << 16;
  d |= static_cast<uint64_t>(1024 >> 4) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(2) << 61;
  return d;
}

// Instruction descriptor, kind::f16: D = fp32, A = B = bf16, both K-major; M = 128 (one CTA) or 256 (CTA pair).
__host__ __device__ constexpr uint32_t make_idesc_bf16(int n, int m = 128) {
  return (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; one thread issues for the CTA.  kind::f16 covers bf16 AND fp16 operands: the
// instruction descriptor (make_idesc_bf16 / make_idesc_f16) names the formats.
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void umma_f16kind(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  umma_bf16(d_tmem, adesc, bdesc, idesc, accumulate);
}

// Same with the A operand in tensor memory (lane = row, two bf16 per 32-bit column, K = 16 -> 8 columns):
// no shared-memory read for A, which is what bounds N <= 128 SS-mode MMAs (A+B operand bytes per cycle
// exceed the 128 B/clk shared-memory port at N = 64).
__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// CTA-pair MMAs (cta_group::2, M = 256): issued by one thread of the leader CTA; A rows 0-127 / accumulator come
// from the leader's shared/tensor memory and rows 128-255 from the peer's at the same addresses; each CTA holds
// N/2 rows of B.
__device__ __forceinline__ void umma_bf16_pair(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                                  uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit of the pair's MMAs: arrives on the barrier at this offset in BOTH CTAs
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  const uint16_t mask = 3;
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(mask)
               : "memory");
}

// The same MMAs taking only the LOW words of the shared-memory descriptors: the high word of every SW128 K-major
// descriptor used here is the constant DESC_HI (SBO = 1024 B, version 1, SWIZZLE_128B), and advancing along K or to
// the next k-block only changes the start-address field in the low word.  32-bit operands keep an unrolled issue
// sequence cheap in registers.
constexpr uint32_t DESC_HI = 0x40004040u;
__device__ __forceinline__ uint32_t make_desc_lo(uint32_t saddr) { return ((saddr & 0x3FFFFu) >> 4) | (1u << 16); }
__device__ __forceinline__ void umma_pair_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_ts_pair_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}

// kind::f8f6f4 with E4M3 operands (K = 32 per instruction: twice the rate of kind::f16), fp32 accumulate.  The
// instruction descriptor has the same bit layout as kind::f16's and E4M3 is format code 0 like F16, so make_idesc_f16()
// serves both; 8-bit K-major SW128 tiles are [rows][128 elements = 128 B], a k-step advances the start address by 32 B.
__device__ __forceinline__ void umma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_f8_pair_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 da, {%1, %5};\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], da, db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}
__device__ __forceinline__ void umma_f8_ts_pair_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "mov.b64 db, {%2, %5};\n\t"
      "tcgen05.mma.cta_group::2.kind::f8f6f4 [%0], [%1], db, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(DESC_HI)
      : "memory");
}

// Arrive on an mbarrier once every previously issued MMA of this thread has completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------- packed fp32 FMA (FFMA2)
// d.x += a.x * b.x ; d.y += a.y * b.y in one issue slot (each half rounds exactly like fmaf).
__device__ __forceinline__ void ffma2(float2& d, const float2 a, const float2 b) {
  uint64_t dd, aa, bb;
  asm("mov.b64 %0, {%1, %2};" : "=l"(dd) : "f"(d.x), "f"(d.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(aa) : "f"(a.x), "f"(a.y));
  asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b.x), "f"(b.y));
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(dd) : "l"(aa), "l"(bb));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(dd));
}

// ---------------------------------------------------------------- warp-level MMA (the decoder's 13 x 13 attention)
// Four 8x8 b16 matrices from shared memory: lane l supplies the address of row l % 8 of matrix l / 8.
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
// d (16x8 fp32) += a (16x16 f16, row) * b (16x8 f16, col)
__device__ __forceinline__ void mma_f16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// two floats -> packed fp16 pair hi and the packed fp16 pair of the remainders
__device__ __forceinline__ void split2_h(float x, float y, uint32_t& hi, uint32_t& lo) {
  const __half2 h2 = __floats2half2_rn(x, y);
  const __half2 l2 = __floats2half2_rn(x - __low2float(h2), y - __high2float(h2));
  hi = *reinterpret_cast<const uint32_t*>(&h2);
  lo = *reinterpret_cast<const uint32_t*>(&l2);
}

// ---------------------------------------------------------------- bf16 split helpers
// x = hi + lo + O(2^-17 |x|): hi = bf16_rn(x), lo = bf16_rn(x - hi).
__device__ __forceinline__ void split_bf16(float x, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(x);
  lo = __float2bfloat16_rn(x - __bfloat162float(hi));
}
// 8 floats -> 4 words of packed bf16 hi pairs and 4 words of packed bf16 lo pairs (element 2i in the low half).
template <bool LO>
__device__ __forceinline__ void split8(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&h2);
    if (LO) {
      const float h0 = __uint_as_float(h[i] << 16), h1 = __uint_as_float(h[i] & 0xffff0000u);
      const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - h0, v[2 * i + 1] - h1);
      l[i] = *reinterpret_cast<const uint32_t*>(&l2);
    }
  }
}
// fp16 flavour (the encoder's split activations): x = hi + lo + O(2^-22 |x|) for |x| within the fp16 range;
// hi saturates at the largest finite fp16 instead of overflowing to infinity.
__device__ __forceinline__ void split8_h(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float a = fminf(fmaxf(v[2 * i], -65504.f), 65504.f), b = fminf(fmaxf(v[2 * i + 1], -65504.f), 65504.f);
    const __half2 h2 = __floats2half2_rn(a, b);
    h[i] = *reinterpret_cast<const uint32_t*>(&h2);
    const __half2 l2 = __floats2half2_rn(v[2 * i] - __low2float(h2), v[2 * i + 1] - __high2float(h2));
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
}
// The same without the clamp (the decoder's fp16x3 mode: its activations are LayerNorm outputs, attention outputs and
// FFN hidden values, orders of magnitude below 65504; a value beyond the fp16 range would become inf and the result NaN --
// loud, not silent).  lo parts below 2^-14 are fp16 subnormals (the tensor core does not flush them): the split then
// carries an ABSOLUTE error <= 2^-25 instead of a relative one.
__device__ __forceinline__ void split8_hn(const float* v, uint32_t* h, uint32_t* l) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    h[i] = *reinterpret_cast<const uint32_t*>(&h2);
    const __half2 l2 = __floats2half2_rn(v[2 * i] - __low2float(h2), v[2 * i + 1] - __high2float(h2));
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
}
// ---------------------------------------------------------------- fp16 + 2 x fp8 split ("fp16f8")
// x * w = xh * wh + xl * wh + xh * wl (+ xl * wl, dropped) with xh = fp16(x), xl = x - xh.  The leading term runs on
// kind::f16; the two cross terms are 2^-11 of it, so 3-4 bits of each factor are enough for them: they run on
// kind::f8f6f4 (E4M3) at twice the rate.  All three terms accumulate in ONE fp32 accumulator, so they share one power-of
// -two scale 2^15: the leading term's operands carry 2^7 (activations) and 2^8 (weights), the cross terms 2^11 * 2^4
// (xl * wh) and 2^0 * 2^15 (xh * wl), which also puts every fp8 factor near 1.  The epilogue multiplies by 2^-15.
constexpr float F8_XS = 128.f, F8_XLS = 2048.f, F8_ACC_INV = 1.f / 32768.f;
// 8 floats -> fp16(x * 2^7) x 8 (4 words), e4m3((x - fp16(x)) * 2^11) x 8 (2 words), e4m3(fp16(x)) x 8 (2 words)
__device__ __forceinline__ void split8_f8(const float* v, uint32_t* h, uint32_t* l8, uint32_t* h8) {
  uint16_t pl[4], ph[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h2 = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
    const float h0 = __low2float(h2), h1 = __high2float(h2);
    // exact (a power-of-two multiple of an fp16 number) below 511.75; saturates beyond instead of overflowing to infinity
    const __half2 hs = __floats2half2_rn(fminf(fmaxf(h0 * F8_XS, -65504.f), 65504.f), fminf(fmaxf(h1 * F8_XS, -65504.f), 65504.f));
    h[i] = *reinterpret_cast<const uint32_t*>(&hs);
    pl[i] = __nv_cvt_float2_to_fp8x2(make_float2((v[2 * i] - h0) * F8_XLS, (v[2 * i + 1] - h1) * F8_XLS), __NV_SATFINITE, __NV_E4M3);
    ph[i] = __nv_cvt_float2_to_fp8x2(make_float2(h0, h1), __NV_SATFINITE, __NV_E4M3);
  }
  l8[0] = pl[0] | (static_cast<uint32_t>(pl[1]) << 16);
  l8[1] = pl[2] | (static_cast<uint32_t>(pl[3]) << 16);
  h8[0] = ph[0] | (static_cast<uint32_t>(ph[1]) << 16);
  h8[1] = ph[2] | (static_cast<uint32_t>(ph[3]) << 16);
}

// The same split for the H chunks of the FFN, which arrive as D = 128 h >= 0 (the epilogue of linear1 folds the 2^7 into its
// scale and bias): fp16 operand = fp16(D) (saturating: h above 511.75 clamps), residual (D - fp16(D)) * 2^4 = (h - hh) 2^11,
// e4m3(hh) from the fp16 pair scaled back by 2^-7 (exact).  7 instructions per value instead of 11 (the compute warps'
// issue slots bound the FFN loop once the MMA hand-offs are off the critical path).
__device__ __forceinline__ void split8_f8_h(const float* D, uint32_t* h, uint32_t* l8, uint32_t* h8) {
  uint16_t pl[4], ph[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t h2;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(h2) : "f"(D[2 * i + 1]), "f"(D[2 * i]));
    h[i] = h2;
    const __half2 hv = *reinterpret_cast<const __half2*>(&h2);
    const float h0 = __low2float(hv), h1 = __high2float(hv);
    pl[i] = __nv_cvt_float2_to_fp8x2(make_float2((D[2 * i] - h0) * 16.f, (D[2 * i + 1] - h1) * 16.f), __NV_SATFINITE, __NV_E4M3);
    const __half2 hh = __hmul2(hv, __float2half2_rn(0.0078125f));
    const uint32_t hhb = *reinterpret_cast<const uint32_t*>(&hh);
    asm("cvt.rn.satfinite.e4m3x2.f16x2 %0, %1;" : "=h"(ph[i]) : "r"(hhb));
  }
  l8[0] = pl[0] | (static_cast<uint32_t>(pl[1]) << 16);
  l8[1] = pl[2] | (static_cast<uint32_t>(pl[3]) << 16);
  h8[0] = ph[0] | (static_cast<uint32_t>(ph[1]) << 16);
  h8[1] = ph[2] | (static_cast<uint32_t>(ph[3]) << 16);
}

// Instruction descriptor, kind::f16 with fp16 operands (both K-major), fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_f16(int n, int m = 128) {
  return (1u << 4) | (static_cast<uint32_t>(n >> 3) << 17) | (static_cast<uint32_t>(m >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack2(__nv_bfloat16 a, __nv_bfloat16 b) {
  return static_cast<uint32_t>(__bfloat16_as_ushort(a)) | (static_cast<uint32_t>(__bfloat16_as_ushort(b)) << 16);
}

// Byte offset of the 16-byte chunk (8 bf16, k = 8*kc .. 8*kc+7) of row r inside one
// [rows][64] SW128 tile.
__host__ __device__ __forceinline__ uint32_t sw128_chunk_off(int r, int kc) {
  return static_cast<uint32_t>((r >> 3) * 1024 + (r & 7) * 128 + ((kc ^ (r & 7)) << 4));
}

}  // namespace ptx
}  // namespace s3d
