// Thin inline-PTX wrappers for the sm_100a features the kernels use: mbarrier, bulk async copies
// (TMA engine, 1-D form), tcgen05 (alloc / mma / commit / ld / fences) and the shared-memory
// matrix descriptors tcgen05.mma consumes.  No library code is involved; every wrapper is one
// PTX instruction (or a bounded spin around one).
#pragma once
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

namespace nerfq {

// ------------------------------------------------------------------------------------------
// addresses
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Identity the compiler cannot see through: keeps a derived address in a register instead of re-deriving it
// (the aligned dynamic shared-memory base costs ~10 instructions to rebuild and ptxas rematerialises it freely).
__device__ __forceinline__ uint32_t opaque_u32(uint32_t v) {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
    return r;
}

// ------------------------------------------------------------------------------------------
// mbarrier
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait comes back every ~50 cycles while the phase is incomplete: ncu r02 counted 25 polls per accumulator wait and
// 28 % of all issued instructions in the poll loops.  A suspend-time hint (NERFQ_MBAR_SUSPEND_NS > 0) was measured and does
// not help: 1, 4 and 20 us all ran 1 % SLOWER than no hint (profiles/r02_ab_wait_hint.log), so the default is none; waits
// that are not latency-critical back off with nanosleep instead (mbar_wait_relaxed).
#ifndef NERFQ_MBAR_SUSPEND_NS
#define NERFQ_MBAR_SUSPEND_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
#if NERFQ_MBAR_SUSPEND_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity), "r"((uint32_t)NERFQ_MBAR_SUSPEND_NS) : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
#endif
    return ok != 0;
}
// non-blocking probe (try_wait may suspend the thread for a while when the phase is not complete yet)
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
#ifndef NERFQ_MBAR_SPIN_LIMIT
#define NERFQ_MBAR_SPIN_LIMIT (1u << 20)
#endif
[[noreturn]] static __device__ __noinline__ void mbar_timeout(uint32_t bar, uint32_t parity) {        // cold and out of line: the wait sites stay small
    printf("nerfq: mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
    __trap();
    __builtin_unreachable();
}
// NERFQ_MBAR_POLL_UNROLL probes per pass of the bookkeeping (spin counter, limit check).  A probe comes back after ~50 cycles
// whether or not the phase completed, so the loop's own 7 instructions per probe are what a waiting warp feeds into the
// scheduler it shares with working warps -- yet 4 probes per pass measured 7 % SLOWER on the forward kernel than 1
// (0.746 vs 0.694 ms, profiles/r02_ab_poll_and_save_split.log), so the loop stays rolled.
#ifndef NERFQ_MBAR_POLL_UNROLL
#define NERFQ_MBAR_POLL_UNROLL 1
#endif
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    for (;;) {
#pragma unroll
        for (int i = 0; i < NERFQ_MBAR_POLL_UNROLL; ++i)
            if (mbar_try_wait(bar, parity)) return;
        if (++spins > (threadIdx.x < 128 ? NERFQ_MBAR_SPIN_LIMIT / (8 * NERFQ_MBAR_POLL_UNROLL) : NERFQ_MBAR_SPIN_LIMIT / NERFQ_MBAR_POLL_UNROLL))
            mbar_timeout(bar, parity);   // (the thread-dependent limit also keeps ptxas from restructuring the wait loops: with a constant limit the issuer code spilled)
    }
}

// For waits with slack (the weight loaders: a freed ring slot is needed again three chunks later): sleep between polls
// instead of re-issuing the probe at full rate, which takes issue slots from the epilogue warps on the same scheduler.
#ifndef NERFQ_RELAXED_WAIT_NS
#define NERFQ_RELAXED_WAIT_NS 128
#endif
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
#if NERFQ_RELAXED_WAIT_NS > 0
        __nanosleep(NERFQ_RELAXED_WAIT_NS);
#endif
        if (++spins > NERFQ_MBAR_SPIN_LIMIT / 8) mbar_timeout(bar, parity);
    }
}

// The epilogue warps' wait for an accumulator: latency-critical (the job starts when it returns), but 16 warps polling
// at full rate are a quarter of all issued instructions.  NERFQ_ACC_WAIT_NS > 0 sleeps that long between probes.
#ifndef NERFQ_ACC_WAIT_NS
#define NERFQ_ACC_WAIT_NS 0
#endif
__device__ __forceinline__ void mbar_wait_acc(uint32_t bar, uint32_t parity) {
#if NERFQ_ACC_WAIT_NS > 0
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(NERFQ_ACC_WAIT_NS);
        if (++spins > NERFQ_MBAR_SPIN_LIMIT) mbar_timeout(bar, parity);
    }
#else
    mbar_wait(bar, parity);
#endif
}

// One lane of a converged warp (the compiler keeps warp-uniform operands of the elected code in uniform registers).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// warp index as a provably warp-uniform value
__device__ __forceinline__ int uniform_warp_idx() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }

// ------------------------------------------------------------------------------------------
// proxies / fences
// ------------------------------------------------------------------------------------------
// generic-proxy smem writes -> visible to the async proxy (tcgen05.mma operand reads, bulk stores)
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// ------------------------------------------------------------------------------------------
// bulk async copies (1-D; executed by the TMA engine, SASS UBLKCP)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar) : "memory");
}
// the same copy with an L2 eviction policy: NERFQ_WEIGHT_L2_POLICY = 1 keeps the weight image (read by every SM, every group)
// in the L2 ahead of the streamed activations
__device__ __forceinline__ void bulk_g2s_evict_last(uint32_t dst_smem, const void* src_gmem, uint32_t bytes, uint32_t bar) {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst_smem), "l"(src_gmem), "r"(bytes), "r"(bar), "l"(pol) : "memory");
}
// ask the copy engine to bring [p, p + bytes) into the L2 (no destination, no completion); bytes a multiple of 16
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* dst_gmem, uint32_t src_smem, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                 ::"l"(dst_gmem), "r"(src_smem), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_but1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// TMEM allocation (one full warp executes these)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ------------------------------------------------------------------------------------------
// tcgen05.mma  (single-CTA, kind::f16: fp16 or bf16 operands from shared memory, fp32 accumulate in TMEM)
// ------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor, K-major operand, rows of `row_bytes` (= swizzle span).
//   bits [ 0,14) start address >> 4        bits [16,30) leading-dim byte offset >> 4 (unused for
//   bits [32,46) stride-dim byte offset>>4               swizzled K-major; set to 1)
//   bits [46,48) descriptor version (1 on sm_100)        bits [61,64) swizzle mode
enum : uint32_t { SWZ_NONE = 0, SWZ_128B = 2, SWZ_64B = 4, SWZ_32B = 6 };

__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr, uint32_t sbo_bytes, uint32_t swizzle) {
    uint64_t d = 0;
    d |= static_cast<uint64_t>((saddr >> 4) & 0x3FFF);
    d |= static_cast<uint64_t>(1) << 16;
    d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= static_cast<uint64_t>(1) << 46;
    d |= static_cast<uint64_t>(swizzle) << 61;
    return d;
}

// Instruction descriptor (upper 32 bits of the "idesc" operand):
//   [4,6) D format (1 = f32)   [7,10) A format   [10,13) B format (0 = f16, 1 = bf16)
//   [15] A major (0 = K)   [16] B major (0 = K)   [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t umma_idesc(uint32_t m, uint32_t n, bool bf16) {
    return (1u << 4) | ((bf16 ? 1u : 0u) << 7) | ((bf16 ? 1u : 0u) << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// Same instruction with the accumulate flag as a compile-time constant (no runtime predicate setup).
template <int kAccumulate>
__device__ __forceinline__ void umma_ss_c(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "n"(kAccumulate) : "memory");
}
// All previously issued tcgen05.mma of this thread arrive (once) on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// ------------------------------------------------------------------------------------------
// CTA pairs (cluster of 2, tcgen05 cta_group::2): one thread of the leader CTA issues M=256 MMAs that use both SMs'
// tensor cores; each CTA supplies its 128 rows of A and its half of B's N columns from its own shared memory (same
// CTA-relative addresses) and receives its 128 rows x N columns of D in its own TMEM.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `addr` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t caddr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(caddr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void st_cluster_f32(uint32_t caddr, float v) {
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(caddr), "f"(v) : "memory");
}
// arrive on an mbarrier of any CTA of the cluster (shared::cluster address), releasing this thread's prior writes
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t caddr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
// ... without ordering any of this thread's memory accesses (a pure signal)
__device__ __forceinline__ void mbar_arrive_remote(uint32_t caddr) {
    asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(caddr) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint32_t bar, uint32_t parity) {      // acquire at cluster scope
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait_cluster(bar, parity)) {
        if (++spins > NERFQ_MBAR_SPIN_LIMIT / 8) {
            printf("nerfq: cluster mbarrier timeout (block %d thread %d bar 0x%x parity %u)\n", blockIdx.x, threadIdx.x, bar, parity);
            __trap();
        }
    }
}
// generic-proxy writes (to either CTA's shared memory) -> visible to the async proxy
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc2(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss2(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// completion of this thread's MMAs arrives on the mbarrier at CTA-relative address `bar` in every CTA of `cta_mask`
__device__ __forceinline__ void umma_commit2_mc(uint32_t bar, uint16_t cta_mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(cta_mask) : "memory");
}

// ------------------------------------------------------------------------------------------
// TMEM -> registers.  32x32b: lane i of the warp reads TMEM lane (32*(warp%4)+i), N consecutive columns.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// registers -> 32 lanes x 16 columns of TMEM (the inverse of tmem_ld16)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------------------------------
// K-major swizzled tile addressing shared by the writers of operand tiles.
// A tile block is [rows][64 halves] = rows x 128 B, 8-row groups of 1024 B, 16-byte chunk index
// XOR (row & 7)  (the SWIZZLE_128B pattern).  Returns the byte offset of chunk `c` (0..7) of `row`.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
    return row * 128u + ((chunk ^ (row & 7u)) << 4);
}
// 64-byte rows (32 halves), 8-row groups of 512 B, chunk (0..3) XOR ((row >> 1) & 3)  (SWIZZLE_64B).
__host__ __device__ __forceinline__ uint32_t sw64_offset(uint32_t row, uint32_t chunk) {
    return row * 64u + ((chunk ^ ((row >> 1) & 3u)) << 4);
}

// explicit shared-space accesses with 32-bit addresses (a pointer derived from the manually aligned dynamic
// shared-memory base is a generic pointer to the compiler, which would emit generic ST/ATOM instead of STS/ATOMS)
__device__ __forceinline__ void red_shared_add_s32(uint32_t addr, int v) {          // native (ATOMS.ADD), unlike the float form
    asm volatile("red.shared.add.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_shared_s32(uint32_t addr) {
    int v;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ uint4 ld_shared_v4(uint32_t addr) {
    uint4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint32_t x, uint32_t y, uint32_t z, uint32_t w) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(x), "r"(y), "r"(z), "r"(w) : "memory");
}
__device__ __forceinline__ void st_shared_f32(uint32_t addr, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ float4 ld_shared_v4f(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}
__device__ __forceinline__ void red_shared_add_f32(uint32_t addr, float v) {
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
// 32-byte global store (one full sector; a warp of consecutive addresses writes eight full lines)
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1,
                                             uint32_t b2, uint32_t b3) {
#ifndef NERFQ_SAVE_ST_POLICY
#define NERFQ_SAVE_ST_POLICY 0
#endif
#if NERFQ_SAVE_ST_POLICY == 0
    asm volatile("st.global.L2::evict_first.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
#elif NERFQ_SAVE_ST_POLICY == 1
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
#else
    asm volatile("st.global.L1::no_allocate.L2::evict_last.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1), "r"(b2), "r"(b3) : "memory");
#endif
}
// packed fp32 pairs (sm_100: add / mul / fma .f32x2 on 64-bit registers)
__device__ __forceinline__ uint64_t pack_f32x2(float lo, float hi) {
    uint64_t d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
    return d;
}
// {y0, y1} = {a0, a1} * m2 + c2   (a0, a1: raw fp32 bit patterns, e.g. straight from tcgen05.ld)
__device__ __forceinline__ void fma_f32x2(float& y0, float& y1, uint32_t a0, uint32_t a1, uint64_t m2, uint64_t c2) {
    uint64_t a, d;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a) : "r"(a0), "r"(a1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(m2), "l"(c2));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(y0), "=f"(y1) : "l"(d));
}
// {lo, hi} -> packed fp16 pair, round to nearest, saturating at +-65504 instead of producing inf (one F2FP either way);
// the relu form clamps negative inputs to +0 in the same instruction
__device__ __forceinline__ uint32_t cvt_pack_f16(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}
__device__ __forceinline__ uint32_t cvt_pack_f16_relu(float lo, float hi) {
    uint32_t d;
    asm("cvt.rn.relu.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(d) : "f"(hi), "f"(lo));
    return d;
}

__device__ __forceinline__ uint32_t pack_half2(float a, float b) {
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_bf162(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

}  // namespace nerfq
