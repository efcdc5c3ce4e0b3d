// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk.tensor), tcgen05 (alloc / mma /
// commit / ld / st / fences) and UMMA descriptors.  No CUTLASS dependency; bit layouts follow the
// PTX ISA "tcgen05" matrix/instruction descriptor tables.
#pragma once

#include <cuda.h>          // CUtensorMap (types only; the driver entry point is resolved at run time)
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace sm100 {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}\n" : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// device-scope flag hand-off between CTAs (global memory)
__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// GDKVM_TRYWAIT_HINT_NS > 0: pass a suspend-time hint, so a waiting thread sleeps in hardware (up to the hint) instead
// of coming back to the issue slots of its scheduler every few hundred cycles.
#ifndef GDKVM_TRYWAIT_HINT_NS
#define GDKVM_TRYWAIT_HINT_NS 0
#endif
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
#if GDKVM_TRYWAIT_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)GDKVM_TRYWAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
#endif
    return ok != 0;
}
// -DGDKVM_DEBUG_WAIT: a wait that exceeds ~4 ms records (block, warp, barrier byte offset, parity) and falls
// through instead of trapping, so the host can read which hand-off was missed (scripts/bisect_sizes.py).
#ifdef GDKVM_DEBUG_WAIT
__device__ unsigned int g_wait_dbg[64 * 4];
__device__ unsigned int g_wait_dbg_n;
__device__ __noinline__ void wait_dbg_record(uint64_t* bar, uint32_t parity) {
    if ((threadIdx.x & 31) == 0) {
        const unsigned int i = atomicAdd(&g_wait_dbg_n, 1u);
        if (i < 64) {
            g_wait_dbg[4 * i] = blockIdx.x; g_wait_dbg[4 * i + 1] = threadIdx.x >> 5;
            g_wait_dbg[4 * i + 2] = smem_u32(bar); g_wait_dbg[4 * i + 3] = parity;
        }
    }
}
#define GDKVM_WAIT_TIMEOUT(bar, parity) do { wait_dbg_record(bar, parity); return; } while (0)
#define GDKVM_WAIT_POLLS (1u << 16)
#else
#define GDKVM_WAIT_TIMEOUT(bar, parity) __trap()
#define GDKVM_WAIT_POLLS (1u << 26)     // a failed try_wait suspends for ~0.1-1 us: minutes before a protocol bug traps
#endif

// Bounded wait: a protocol bug traps (error at the next sync on the host) instead of hanging the GPU.
// Deliberately not inlined: there are dozens of call sites and the kernel is instruction-fetch sensitive.
#ifndef GDKVM_SLEEP_WAIT
#define GDKVM_SLEEP_WAIT 0
#endif
#ifndef GDKVM_SLEEP_INL
#define GDKVM_SLEEP_INL 0
#endif
// The time-out counts polls, not clocks: reading the clock (CS2R) goes through the XU pipe, which the fp32 -> bf16
// packs of this kernel already saturate (ncu: sm__inst_executed_pipe_xu > 100 % of sustained peak).
static __device__ __noinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (GDKVM_SLEEP_WAIT) __nanosleep(GDKVM_SLEEP_WAIT);
        if (++polls > GDKVM_WAIT_POLLS) GDKVM_WAIT_TIMEOUT(bar, parity);
    }
}

// Inlined wait for the MMA/TMA issuer warps: a CALL in their loops would force every loop-invariant
// descriptor out of the uniform registers (R2UR before each UTCHMMA).
__device__ __forceinline__ void mbar_wait_inl(uint64_t* bar, uint32_t parity) {
#if defined(GDKVM_DEBUG_WAIT) || GDKVM_SLEEP_INL
    if (mbar_try_wait(bar, parity)) return;
    uint32_t polls = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (GDKVM_SLEEP_INL) __nanosleep(GDKVM_SLEEP_INL);
        if (++polls > GDKVM_WAIT_POLLS) GDKVM_WAIT_TIMEOUT(bar, parity);
    }
#else
    // one PTX block (labels are local to the braces): three instructions when the phase has already completed --
    // the kernel is instruction-issue bound and executes ~70 warp-level waits per chunk
#if GDKVM_TRYWAIT_HINT_NS > 0
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %3;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, %2;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "trap;\n"
        "WAIT_DONE:\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)GDKVM_WAIT_POLLS), "r"((uint32_t)GDKVM_TRYWAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .u32 c;\n\t"
        "mov.u32 c, 0;\n"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "add.u32 c, c, 1;\n\t"
        "setp.lt.u32 p, c, %2;\n\t"
        "@p bra WAIT_LOOP;\n\t"
        "trap;\n"
        "WAIT_DONE:\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(parity), "r"((uint32_t)GDKVM_WAIT_POLLS)
        : "memory");
#endif
#endif
}

// Wait with a warp-uniform exit (vote): every lane polls, the loop condition is the VOTE result, so ptxas keeps
// the issuer warps' loop state and descriptors in uniform registers.
__device__ __forceinline__ void mbar_wait_u(uint64_t* bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!__all_sync(0xffffffffu, mbar_try_wait(bar, parity))) {
#ifdef GDKVM_DEBUG_WAIT
        if (++spins > (1u << 16)) GDKVM_WAIT_TIMEOUT(bar, parity);
#else
        if (++spins > (1u << 26)) __trap();
#endif
    }
}

// ---------------------------------------------------------------- proxies / fences
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2,
                                            int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* m, const void* src, int c0, int c1, int c2, int c3, int c4) {
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5, %6}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
                 : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, const void* src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(src)), "r"(c0), "r"(c1)
                 : "memory");
}
// 32 bytes per thread: a whole sector per lane (sm_100 256-bit store, SASS STG.E.ENL2.256)
__device__ __forceinline__ void st_global_v8(void* p, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6,
                                             uint32_t a7) {
    asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"l"(p), "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_read1() { asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---------------------------------------------------------------- TMEM allocation
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {   // whole warp
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {     // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---------------------------------------------------------------- UMMA descriptors
// Shared-memory matrix descriptor (64 bit): start address >>4 [0,14), leading byte offset >>4
// [16,30), stride byte offset >>4 [32,46), version=1 [46,48), layout type [61,64) (2 = 128B swizzle).
__device__ __forceinline__ uint64_t umma_smem_desc_sw128(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
// Instruction descriptor for kind::f16 with bf16 A/B and fp32 accumulate.
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int M, int N, bool a_mn_major, bool b_mn_major,
                                                       bool a_negate = false, bool b_negate = false) {
    return (1u << 4)                                   // D format: f32
           | (1u << 7) | (1u << 10)                    // A, B format: bf16
           | ((a_negate ? 1u : 0u) << 13) | ((b_negate ? 1u : 0u) << 14)
           | ((a_mn_major ? 1u : 0u) << 15) | ((b_mn_major ? 1u : 0u) << 16)
           | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// Warp-uniform variants: the whole (converged) warp executes the call, one elected lane issues.  ptxas keeps
// the operands in uniform registers and emits a single predicated UTCHMMA; issued from a divergent
// `if (lane == 0)` region the same instruction is wrapped in an ELECT / BRA.U.ANY loop and costs ~113
// cycles instead of ~33 (tests/probes/mma_timing.cu).  elect.sync is deterministic for a given
// membermask, so the MMAs and the commit that tracks them come from the same lane.
__device__ __forceinline__ void umma_ss_w(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// Four K-slices of one 64-deep contraction behind ONE election: the operand moves into uniform registers of all
// four MMAs are issued back to back instead of one dependent elect / move / issue sequence per MMA (~100 cycles each).
__device__ __forceinline__ void umma4_ss_w(uint32_t d_tmem, uint64_t a0, uint64_t a1, uint64_t a2, uint64_t a3, uint64_t b0, uint64_t b1,
                                           uint64_t b2, uint64_t b3, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %10, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %5, %9, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %2, %6, %9, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %3, %7, %9, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %4, %8, %9, 1;\n\t}\n"
        ::"r"(d_tmem), "l"(a0), "l"(a1), "l"(a2), "l"(a3), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
__device__ __forceinline__ void umma4_ts_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3,
                                           uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 a1, a2, a3;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %7, 0;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], %3, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], %4, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], %5, %6, 1;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"((uint32_t)accumulate)
        : "memory");
}
// ... followed by commits on one or two mbarriers, still behind the same election
__device__ __forceinline__ void umma4_ts_commit_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3,
                                                  uint32_t idesc, bool accumulate, uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 a1, a2, a3;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %7, 0;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], %3, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], %4, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], %5, %6, 1;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"((uint32_t)accumulate), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void umma4_ts_commit2_w(uint32_t d_tmem, uint32_t a_tmem, uint64_t b0, uint64_t b1, uint64_t b2, uint64_t b3,
                                                   uint32_t idesc, bool accumulate, uint64_t* bar0, uint64_t* bar1) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t.reg .b32 a1, a2, a3;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, %7, 0;\n\t"
        "add.u32 a1, %1, 8;\n\tadd.u32 a2, %1, 16;\n\tadd.u32 a3, %1, 24;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %6, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a1], %3, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a2], %4, %6, 1;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [a3], %5, %6, 1;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%8];\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%9];\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(b0), "l"(b1), "l"(b2), "l"(b3), "r"(idesc), "r"((uint32_t)accumulate), "r"(smem_u32(bar0)),
          "r"(smem_u32(bar1))
        : "memory");
}
__device__ __forceinline__ void umma_commit_w(uint64_t* bar) {
    asm volatile(
        "{\n\t.reg .pred q;\n\telect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n"
        ::"r"(smem_u32(bar)) : "memory");
}
// All previously issued MMAs of this thread arrive (once) on `bar` when they complete.
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// ---------------------------------------------------------------- TMEM <-> registers
// 32x32b shape: lane i of the warp <-> TMEM lane (base_lane + i); register j <-> column (base_col + j).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
        "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),
          "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),
          "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, uint32_t (&r)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t (&r)[4]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3])
                 : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),
          "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// 16x256b shape: the warp reads 16 TMEM lanes (taddr lane .. +15) x 64 columns in the mma.sync C-fragment
// layout: register 4q + 2h + e of lane l  <->  TMEM lane (l / 4 + 8 h), column 8 q + 2 (l % 4) + e.
// (verified on hardware: tests/probes/frag_probe.cu)
__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// four 8x8 b16 matrices, transposed on the way: fragment element (row r, col c) lands in the 16-byte
// memory row c (address supplied by lane 8 m + c for matrix m) at position r
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// pack two fp32 into one bf16x2 word: `lo` lands in bits [0,16), `hi` in bits [16,32).
// (-DGDKVM_PACK_ALU: integer pack on the ALU pipe, 3 instructions instead of one F2FP -- measured 19 % SLOWER on the
// whole kernel, which is how the instruction-issue bound of this kernel was confirmed: time follows instruction count.)
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
#ifdef GDKVM_PACK_ALU
    return __byte_perm(__float_as_uint(lo) + 0x8000u, __float_as_uint(hi) + 0x8000u, 0x7632);
#else
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
#endif
}

// byte offset of (row, 16-byte chunk) inside a 128B-swizzled tile whose rows are 128 bytes
// (TMA SWIZZLE_128B == UMMA layout type 2; the tile base must be 1024-byte aligned)
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk16) {
    return (uint32_t)(row * 128 + ((chunk16 ^ (row & 7)) << 4));
}

}  // namespace sm100
