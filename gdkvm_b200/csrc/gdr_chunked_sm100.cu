// Chunked (WY/UT) GDR/LKVA kernel on the 5th-gen tensor cores of sm_100a -- warp-specialised.
//
// One CTA = one (clip, head) chain, all V value columns, 64-token chunks (a frame is one chunk;
// frames longer than 64 tokens are cut into 64-token sub-chunks, shorter ones are zero-padded by
// TMA out-of-bounds fill: k = 0, beta = 0, g = 0 rows are exact no-ops).
//
// The problem is held TRANSPOSED so the value dimension sits on the 128 TMEM lanes (M = 128):
//     S^T [V x 64] fp32 lives in TMEM for the whole clip (never touches HBM),
//     Vn^T = V^T T'^T - Sb W^T           (Sb = bf16 copy of S^T, TMEM A-operand)
//     O^T  = Sb Q~^T + Vnb P^T           (Vnb = bf16 copy of Vn^T, TMEM A-operand, aliases Vn)
//     S^T  = gamma S^T + Vnb K'          (fp32 accumulate in place)
// Every contraction is a 128 x 64 x 64 tcgen05.mma (bf16 in, fp32 TMEM accumulate).
//
// Roles (20 warps):
//   warps 0-7   K-side group: everything that does not depend on the state, one chunk AHEAD of the
//               state side: [K;Q]K^T accumulators -> masked/gated A (fp32) and P (bf16); the
//               triangular inverse (I + A)^-1 in fp32-grade arithmetic (16x16 forward substitution on
//               CUDA cores, block merges as 3xTF32 mma.sync); T' = T diag(..) ; W^T ; the in-place row
//               scalings K~, Q~ and K' of the TMA tiles (the 128B swizzle keeps rows intact).
//   warps 8-11 / 12-15   state warpgroups, one per 128 value columns: S -> (Sb, gamma S), Vn -> Vnb,
//               readout O -> bf16 -> staging -> TMA store; initial / final state.
//   warp 16     issuer K: TMA loads (q, k, v tiles, one chunk of prefetch) and the K-side MMAs.
//   warps 17-18  issuers S: the five state-side MMA groups of one value half each.
//   warp 19     idle: completes the issuer warpgroup, which gives most of its registers to the state warpgroups.
// The roles meet only through mbarriers (tcgen05.commit / arrive), so the K-side work of chunk n+1,
// the state-side work of chunk n and the TMA traffic of chunk n+2 overlap.
//
// "f-folded" gating (fast path, chunk decay > e^-60): with e_i = exp(Gamma_i), f_j = exp(-Gamma_j)
//     T = diag(e) X diag(f),  X = (I + strict_tril(beta_i k_i.k_j))^-1   (no decay inside the solve)
// and the chunk is carried in the scaled variable V^_j = f_j Vnew_j, which turns every decay factor
// into a row / column / scalar factor that is applied where the data already passes through registers:
//     T' = X diag(f beta)                      (column factor in the fp32 -> bf16 conversion of X)
//     W  = T' (K e)                            (e_j applied to the K fragments of the in-register mma.sync)
//     O  = diag(scale e) (Q S + tril(Q K^T) V^) (row factor applied in the readout epilogue; raw Q, unscaled P)
//     S' = gamma (S + K^T V^)                  (raw K, loaded twice by TMA; gamma applied by the next S pass)
// so the fast path never rescales a tile in shared memory.  Chunks with stronger decay use the
// per-element exp(Gamma_i - Gamma_j) form (slow path, rare): same MMA sequence, tiles rescaled in place.
//
// Time segments.  One CTA per chain leaves a ragged last wave (512 chains on 148 SMs: 3.46 waves, the fourth one 46 %
// full).  A chain is therefore cut at chunk boundaries into `nseg` segments that are separate work units: unit
// u = seg * chains + chain, taken in ticket order (an atomic counter, so a unit's predecessor is always resident or
// done).  A segment hands its fp32 state to the next one through a per-launch scratch buffer in global memory
// (64 KiB per chain, released by a flag); the arithmetic is the one of an uncut chain, bit for bit.  The host picks
// nseg by simulating the unit schedule (pick_segments).
//
// Packed variable-length sequences (kVar, the `cu_seqlens` call): the clips lie back to back in one token stream.  A
// one-block prologue kernel turns the device-resident sequence offsets into a table of work units (sequence, first
// token, valid tokens, chunks, segment), level by level so that every unit's predecessor comes earlier in ticket
// order; the host only knows an upper bound of the unit count and the surplus CTAs leave at once.  Rows past the end
// of a sequence carry the NEXT sequence's tokens: g = beta = 0 makes them exact no-ops of the recurrence, and the
// last chunk's readout is stored row by row instead of by TMA so that nothing is written past the sequence.
//
// Training forward (kSave): a second instantiation whose S pass also stores the bf16 operand copy Sb of every chunk-start state
// (the only thing the backward kernel, gdr_bwd_sm100.cu, needs from the forward).
//
// Layout facts used here were verified on hardware by tests/probes/umma_probe.cu.
// Math: oracle/gdr_ref.py::gdr_chunk_ref.
#include <algorithm>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <tuple>
#include <vector>

#include "gdr_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.h"
#include "tri_solve.cuh"

namespace gdkvm {
namespace {

using namespace sm100;

constexpr int kKThreads = 256;                 // K-side group
// 19 working warps + one idle warp that completes the issuer warpgroup (setmaxnreg works on whole warpgroups).  The CTA
// starts with 96 registers per thread (5 warps per scheduler: 480 of its 512 registers per lane); the issuer warpgroup then
// shrinks to 40 and the two state warpgroups grow to 120.  An increase can only take what a decrease of the SAME CTA has
// released (4 warps x 56 >= 8 warps x 24); asking for more (state 128, or K-side 112 as well) blocks forever.  With 96
// registers the Vnb pass (64 fp32 accumulator values per thread in flight) saved and restored 14 registers around its
// TMEM loads, on the recurrence: 0 spill bytes now, -2.1 % (scripts/ab.sh).
constexpr int kThreads = 20 * 32;

// ---- shared memory map (bytes from a 1024-aligned base) ----
// Two TMA rings with different lifetimes: the K|Q tiles of a chunk live from the K side of the chunk (one
// chunk ahead) to its state update, the V tile only until U = V^T T'^T has been formed.
constexpr uint32_t kKqSlots = 3, kKqSlotBytes = 16384;   // Kt 8K | Qt 8K   (Qt must follow Kt: stacked [K;Q] operand)
constexpr uint32_t kOffKq = 0;
constexpr uint32_t kOffV = kKqSlots * kKqSlotBytes;       // V ring: [2 slots][2 value halves][2 x 64 values][64 tok][64] bf16
constexpr uint32_t kVSlotBytes = 32768;
constexpr uint32_t kOffPp = kOffV + 2 * kVSlotBytes;      // P   [2]  (B of the intra-chunk readout, K-major)
constexpr uint32_t kOffWt = kOffPp + 16384;      // W^T [2]  (B of the state correction, MN-major)
constexpr uint32_t kOffTp = kOffWt + 16384;      // T'  [2]  (B of U / W, K-major)
constexpr uint32_t kOffOst = kOffTp + 16384;     // readout staging, per value half [2][64 tok][64] bf16
constexpr uint32_t kOffH = kOffOst + 32768;      // fp16 solve matrix (128B-swizzled rows, tri_solve.cuh)
constexpr uint32_t kOffF = kOffH + 8192;
//   floats: beta[2][64] Gam[2][64] E[4][64] Cj[2][64] Kd[4][64] Ofac[4][64] post[4] pre[4] fast[4] Eb[4][32] (bf16 pairs)
constexpr uint32_t kNumFloats = 20 * 64 + 12;
constexpr uint32_t kOffBar = kOffF + kNumFloats * 4;
constexpr uint32_t kNumBars = 27;
constexpr uint32_t kSmemBytes = kOffBar + kNumBars * 8 + 48 + 1024;   // + tmem slot + unit info + alignment slack
static_assert(kSmemBytes <= 232448, "exceeds the 227 KB dynamic shared memory limit");

// ---- tensor memory map (columns) ----
constexpr uint32_t kColS = 0;      // S^T   [h]: +64h   fp32
constexpr uint32_t kColVn = 128;   // Vn^T  [h]: +64h   fp32; its first 32 columns are re-used for Vnb (bf16)
constexpr uint32_t kColO = 256;    // O^T   [h]: +64h   fp32
constexpr uint32_t kColSb = 384;   // Sb    [h]: +32h   bf16 x2 per column
constexpr uint32_t kColKQ = 448;   // [K;Q]K^T, later W^T (64 columns)
constexpr uint32_t kTmemCols = 512;

// ---- mbarrier slots ----
enum Bar : int {
    kKqTile = 0,     // [3] K|Q tiles of a chunk landed                  (tx)      -> issuer K, K group
    kKqFull = 10,     //     [K;Q]K^T accumulators complete                (commit)  -> K group
    kKqFree = 3,     //     [K;Q]K^T accumulators drained                 (256)     -> issuer K
    kTpReady = 4,    // [2] K side done: T', P, decay factors published   (1)       -> issuer S (U), state groups (W)
    kKsideFull = 6,  // [2] W^T operand written by the state groups       (128 NH)  -> issuer S
    kKsideEmpty = 8, // [2] every MMA of the chunk completed              (commit)  -> K group (operand buffers free)
    kSbReady = 11,   // [2] per half: Sb + decayed S in TMEM              (128)     -> issuer S
    kVnFull = 13,    // [2] per half: Vn^T complete                       (commit)  -> state group
    kVnbReady = 15,  // [2] per half: Vnb in TMEM                         (128)     -> issuer S
    kSReady = 17,    // [2] per half: state update complete               (commit)  -> state group
    kOFull = 19,     // [2] per half: readout accumulators complete       (commit)  -> state group
    kOFree = 21,     // [2] per half: readout accumulators drained        (128)     -> issuer S
    kVTile = 23,     // [2][2] V tile of a chunk, per value half         (tx)      -> issuer S (U)
};

// ---- optional phase timers (build with -DGDKVM_PHASE_TIMERS: scripts/phase_timers.py) ----
#ifdef GDKVM_PHASE_TIMERS
// g_phase_cycles[slot]: cycles accumulated per phase; g_phase_trace[slot][c]: clock64 at the END of the phase
// for chunks kTraceFirst + c (a steady-state window), which gives a cross-role timeline of CTA 0.
__device__ unsigned long long g_phase_cycles[64];
__device__ long long g_phase_trace[64][8];
__device__ long long g_unit_marks[16];      // clock64 of CTA 0 at fixed points of a work unit (scripts/unit_marks.py)
#define UM(i, cond) do { if (blockIdx.x == 0 && (cond)) g_unit_marks[i] = clock64(); } while (0)
constexpr int kTraceFirst = 40;
#define PT_DECL long long pt_prev = clock64();
#define PT(slot, cond, chunk)                                                          \
    do {                                                                               \
        if (blockIdx.x == 0 && (cond)) {                                               \
            const long long pt_now = clock64();                                        \
            atomicAdd(&g_phase_cycles[slot], (unsigned long long)(pt_now - pt_prev));  \
            if ((chunk) >= kTraceFirst && (chunk) < kTraceFirst + 8)                   \
                g_phase_trace[slot][(chunk) - kTraceFirst] = pt_now;                   \
            pt_prev = pt_now;                                                          \
        }                                                                              \
    } while (0)
#else
#define PT_DECL
#define PT(slot, cond, chunk) do { } while (0)
#define UM(i, cond) do { } while (0)
#endif

// ---- ablation switches (timing experiments only; results are wrong when any bit is set) ----
// scripts/ablate.py builds one library per bit and times configs[1] with each phase's work removed while the
// synchronisation skeleton stays intact: the drop in ms/step is that phase's share of the critical cycle.
#ifndef GDKVM_ABLATE
#define GDKVM_ABLATE 0
#endif
#define ABL(bit) ((GDKVM_ABLATE >> (bit)) & 1)
// bits: 0 gating  1 solve levels 0-1  2 solve level 2  3 T' conversion  4 W^T  5 S pass  6 Vnb pass  7 readout

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void kbar() { named_bar_sync(1, kKThreads); }

__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t w, float s) {
    const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
    return pack_bf16(lo * s, hi * s);
}
__device__ __forceinline__ uint4 scale_row8(uint4 v, float s) {
    return make_uint4(scale_bf16x2(v.x, s), scale_bf16x2(v.y, s), scale_bf16x2(v.z, s), scale_bf16x2(v.w, s));
}

// ---- warp-level MMA on the legacy tensor path: tiny K-side products that live in registers ----
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}

// Gate scan of one chunk (one warp; lane l holds g, beta of tokens 2l, 2l+1): Gamma = cumsum(g), decay
// factors, fast/slow decision.
__device__ __forceinline__ void gate_scan(const float (&gv)[2], const float (&bv)[2], float* sBt, float* sGam, float* sE, uint32_t* sEb,
                                          float* sCj, float* sKd, float* sFast, float* ofac, float* post, float* pre,
                                          float scale, int lane) {
    const float g0 = gv[0], g1 = gv[1];
    float s = g0 + g1;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, s, off);
        if (lane >= off) s += t;
    }
    const float G1 = s, G0 = s - g1;
    const float Gl = __shfl_sync(0xffffffffu, s, 31);
    const bool fast = Gl > -60.f;
    const float gam = __expf(Gl), e0 = __expf(G0), e1 = __expf(G1);
    *reinterpret_cast<float2*>(sBt + 2 * lane) = make_float2(bv[0], bv[1]);
    *reinterpret_cast<float2*>(sGam + 2 * lane) = make_float2(G0, G1);
    *reinterpret_cast<float2*>(sE + 2 * lane) = make_float2(e0, e1);
    sEb[lane] = pack_bf16(e0, e1);
    *reinterpret_cast<float2*>(sCj + 2 * lane) =                               // column factor of T'
        make_float2(bv[0] * (fast ? __expf(-G0) : 1.f), bv[1] * (fast ? __expf(-G1) : 1.f));
    *reinterpret_cast<float2*>(sKd + 2 * lane) = make_float2(__expf(Gl - G0), __expf(Gl - G1));   // slow path: K' factor
    *reinterpret_cast<float2*>(ofac + 2 * lane) = make_float2(fast ? scale * e0 : 1.f, fast ? scale * e1 : 1.f);   // readout row factor
    if (lane == 0) { *post = fast ? gam : 1.f; *pre = fast ? 1.f : gam; *sFast = fast ? 1.f : 0.f; }
}

// One 16 (key dims d) x 32 (tokens i) unit of  W^T[d][i] = sum_j (K[j][d] e_j) T'[i][j]  in registers:
// bf16 mma.sync with ldmatrix from the 128B-swizzled K tile (transposed) and T' tile; e_j is applied to
// the K fragments; the result is written as the bf16 MN-major operand rows of the Vn correction MMA
// (row = key dim d, contiguous over tokens i).  T' is lower triangular: slices with j > i are skipped.
__device__ __forceinline__ void w_unit_mma(uint32_t aK, uint32_t aT, uint8_t* wt, const uint32_t* eB, int mt, int ng, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
        if (ks * 16 <= ng * 32 + 31) {
            uint32_t af[4], bf0[4], bf1[4];
            ldmatrix_x4_trans(af, aK + sw128_offset(ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, mt * 2 + ((lane >> 3) & 1)));
            {   // K~ = K e_j: fragment registers hold tokens j = 16 ks + 2t (+1) and + 8.  One packed bf16 multiply per
                // register (e_j rounded to bf16 first: one more 2^-9 rounding on W than an fp32 multiply, no visible change in
                // the parity numbers, a fifth of the instructions)
                const uint32_t p01 = eB[ks * 8 + t], p89 = eB[ks * 8 + t + 4];
                asm("mul.rn.bf16x2 %0, %0, %1;" : "+r"(af[0]) : "r"(p01));
                asm("mul.rn.bf16x2 %0, %0, %1;" : "+r"(af[1]) : "r"(p01));
                asm("mul.rn.bf16x2 %0, %0, %1;" : "+r"(af[2]) : "r"(p89));
                asm("mul.rn.bf16x2 %0, %0, %1;" : "+r"(af[3]) : "r"(p89));
            }
            ldmatrix_x4(bf0, aT + sw128_offset(ng * 32 + ((lane >> 4) & 1) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)));
            ldmatrix_x4(bf1, aT + sw128_offset(ng * 32 + 16 + ((lane >> 4) & 1) * 8 + (lane & 7), ks * 2 + ((lane >> 3) & 1)));
            if (ks * 16 <= ng * 32 + 7) mma_bf16(acc[0], af, bf0[0], bf0[1]);
            if (ks * 16 <= ng * 32 + 15) mma_bf16(acc[1], af, bf0[2], bf0[3]);
            if (ks * 16 <= ng * 32 + 23) mma_bf16(acc[2], af, bf1[0], bf1[1]);
            mma_bf16(acc[3], af, bf1[2], bf1[3]);
        }
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
        const int i0 = ng * 32 + nt * 8 + 2 * t;
        *reinterpret_cast<uint32_t*>(wt + sw128_offset(mt * 16 + g, i0 >> 3) + (i0 & 7) * 2) = pack_bf16(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(wt + sw128_offset(mt * 16 + g + 8, i0 >> 3) + (i0 & 7) * 2) = pack_bf16(acc[nt][2], acc[nt][3]);
    }
}

// Rows 0 .. valid-1 of one value half of the readout staging tile ([2 value blocks][64 tokens][64] bf16, 128B swizzle)
// -> global memory with 16-byte stores, by the 128 threads of a state warpgroup.  Out of line on purpose: it runs once
// per clip and must not add to the register pressure of the readout it is called from.
__device__ __noinline__ void store_valid_rows(const uint8_t* stage, __nv_bfloat16* og, int64_t row_stride, int valid, int stid, int nblk) {
    for (int idx = stid; idx < valid * 16; idx += 128) {
        const int row = idx >> 4, blk = (idx >> 3) & 1, ck = idx & 7;
        if (blk >= nblk) continue;
        const uint4 val = *reinterpret_cast<const uint4*>(stage + blk * 8192 + sw128_offset(row, ck));
        *reinterpret_cast<uint4*>(og + (int64_t)row * row_stride + blk * 64 + ck * 8) = val;
    }
}

// four 128x64x16 tcgen05 MMAs covering K = 64; descriptors advance by a fixed step per K slice.
// Warp-uniform: called by every lane of the (converged) issuer warp, one elected lane issues.
__device__ __forceinline__ void umma4_ss(uint32_t d, uint64_t a, uint32_t astep, uint64_t bdesc, uint32_t bstep, uint32_t idesc, bool acc0) {
#ifdef GDKVM_UMMA_SINGLE
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss_w(d, a + (uint64_t)(k * astep), bdesc + (uint64_t)(k * bstep), idesc, acc0 || k > 0);
#else
    umma4_ss_w(d, a, a + (uint64_t)astep, a + (uint64_t)(2 * astep), a + (uint64_t)(3 * astep), bdesc, bdesc + (uint64_t)bstep,
               bdesc + (uint64_t)(2 * bstep), bdesc + (uint64_t)(3 * bstep), idesc, acc0);
#endif
}
__device__ __forceinline__ void umma4_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t bstep, uint32_t idesc, bool acc0) {
#ifdef GDKVM_UMMA_SINGLE
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ts_w(d, a_tmem + k * 8, bdesc + (uint64_t)(k * bstep), idesc, acc0 || k > 0);
#else
    umma4_ts_w(d, a_tmem, bdesc, bdesc + (uint64_t)bstep, bdesc + (uint64_t)(2 * bstep), bdesc + (uint64_t)(3 * bstep), idesc, acc0);
#endif
}

// ... with the commit(s) that follow them behind the same election
__device__ __forceinline__ void umma4_ts_commit(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t bstep, uint32_t idesc, bool acc0,
                                                uint64_t* bar0, uint64_t* bar1 = nullptr) {
#ifdef GDKVM_UMMA_SINGLE
    umma4_ts(d, a_tmem, bdesc, bstep, idesc, acc0);
    umma_commit_w(bar0);
    if (bar1 != nullptr) umma_commit_w(bar1);
#else
    if (bar1 == nullptr)
        umma4_ts_commit_w(d, a_tmem, bdesc, bdesc + (uint64_t)bstep, bdesc + (uint64_t)(2 * bstep), bdesc + (uint64_t)(3 * bstep), idesc, acc0, bar0);
    else
        umma4_ts_commit2_w(d, a_tmem, bdesc, bdesc + (uint64_t)bstep, bdesc + (uint64_t)(2 * bstep), bdesc + (uint64_t)(3 * bstep), idesc, acc0,
                           bar0, bar1);
#endif
}

// n / d for 0 <= n < 2^31 and a run-time d >= 1 without the ~40-instruction integer division sequence (the kernel is
// instruction-issue bound and maps chunk -> (frame, chunk in frame) several times per chunk).  Host-side constants.
struct FastDiv {
    unsigned int mul, shr, d;
    __host__ static FastDiv make(unsigned int d) {
        FastDiv f{0u, 0u, d};
        if (d > 1) {
            unsigned int l = 0;
            while ((1ull << l) < d) ++l;                              // ceil(log2 d)
            const unsigned long long p = 31 + l;
            f.mul = (unsigned int)(((1ull << p) + d - 1) / d);        // ceil(2^p / d) < 2^32
            f.shr = (unsigned int)(p - 32);
        }
        return f;
    }
    __device__ __forceinline__ int div(int n) const { return d == 1 ? n : (int)(__umulhi((unsigned int)n, mul) >> shr); }
};

// work-unit table of the packed variable-length call: utab[0] = number of entries, entry e at utab + 8 + 8 e:
// [0] sequence [1] first token (row of the packed stream) [2] valid tokens [3] chunks [4] segment [5] last segment
// [6] chunk-state slot of the first chunk (training forward)
constexpr int kUnitInts = 8;

// kSave: the training forward (also stores the bf16 chunk-start states) -- a separate instantiation, so that the inference
// kernel's S pass, which is on the recurrence, carries neither the branch nor its register pressure (it cost 2.4 % as a run-time test)
template <bool kVar, bool kSave>
__global__ void __launch_bounds__(kThreads, 1)
gdr_chunk_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                 const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mo,
                 const GdkvmGdrParams p, const int C, const int F, const FastDiv div_cpf,
                 const int nseg, const int seg_chunks, float* __restrict__ xstate, int* __restrict__ xsync,
                 const int* __restrict__ utab, __nv_bfloat16* __restrict__ sdump) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-align inside the shared window with pointer arithmetic only (an integer round trip would
    // demote every access below from LDS/STS to generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    float* sBt = reinterpret_cast<float*>(smem + kOffF);  // [2][64] beta
    float* sGam = sBt + 128;                              // [2][64] Gamma_i (inclusive cumsum of g)
    float* sE = sGam + 128;                               // [4][64] exp(Gamma_i) of chunk n in slot n & 3
    float* sCj = sE + 256;                                // [2][64] column factor of T': f_j beta_j (fast) | beta_j (slow)
    float* sKd = sCj + 128;                               // [4][64] slow path: column factor of the second Vnb copy, exp(Gamma_last - Gamma_j)
    float* sOfac = sKd + 256;                             // [4][64] readout row factor of chunk n in slot n & 3: scale e_i | 1
    float* sPost = sOfac + 256;                           // [4] factor applied to S AFTER chunk n's accumulate: gamma | 1
    float* sPre = sPost + 4;                              // [4] factor applied to S BEFORE chunk n's accumulate: 1 | gamma
    float* sFast = sPre + 4;                              // [4] chunk n in slot n & 3
    uint32_t* sEb = reinterpret_cast<uint32_t*>(sFast + 4);   // [4][32] exp(Gamma_i) as bf16 pairs (tokens 2l, 2l+1): W^T fragment factors
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);
    // work unit of this CTA, written once by thread 0: [0] time segment [1] chain [2] clip [3] head [4] first chunk
    // (kVar: first token) [5] kVar: chunks [6] kVar: valid tokens [7] last segment of the chain.  Read through a volatile pointer where it is used (single threads, outside the hot loops) so that
    // none of it occupies registers of the instruction-bound warps for the whole kernel.
    volatile int* s_info = reinterpret_cast<volatile int*>(s_tmem + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    UM(0, tid == 0);                              // kernel entry
    // V = 64: one state warpgroup whose upper 64 TMEM lanes idle on zeros (the second value block of the V ring is never loaded)
    const int V = p.V, NH = V > 128 ? 2 : 1, VB = V >> 6;
    const uint32_t v_tile_bytes = V >= 128 ? 16384u : 8192u;
    const int cpf = (C + 63) >> 6;
    const float scale = p.scale;

    if constexpr (kVar) {
        if (tid == 0) {       // unit = ticket -> (table entry, head); CTAs past the end of the table have nothing to do
            const int unit = atomicAdd(xsync, 1), entry = unit / p.H, head = unit - entry * p.H;
            s_info[5] = 0;
            if (entry < utab[0]) {
                const int* e = utab + kUnitInts * (1 + entry);
                s_info[0] = e[4]; s_info[1] = e[0] * p.H + head; s_info[2] = 0; s_info[3] = head;
                s_info[4] = e[1]; s_info[5] = e[3]; s_info[6] = e[2]; s_info[7] = e[5]; s_info[8] = e[6];
            }
        }
        __syncthreads();
        if (s_info[5] == 0) return;
    }
    if (tid == 0) {
        if constexpr (!kVar) {    // unit -> (time segment, chain), in ticket order when chains are cut in time
            const int unit = nseg > 1 ? atomicAdd(xsync, 1) : (int)blockIdx.x, nchains = p.B * p.H;
            const int seg = unit / nchains, chain = unit - seg * nchains, clip = chain / p.H, nb = seg * seg_chunks;
            s_info[0] = seg; s_info[1] = chain; s_info[2] = clip; s_info[3] = chain - clip * p.H;
            s_info[4] = nb; s_info[7] = seg == nseg - 1;
        }
        for (int i = 0; i < 3; ++i) mbar_init(&bars[kKqTile + i], 1);
        for (int i = 0; i < 4; ++i) mbar_init(&bars[kVTile + i], 1);
        mbar_init(&bars[kKqFull], 1); mbar_init(&bars[kKqFree], kKThreads);
        mbar_init(&bars[kTpReady], 1); mbar_init(&bars[kTpReady + 1], 1);
        mbar_init(&bars[kKsideFull], 128 * NH); mbar_init(&bars[kKsideFull + 1], 128 * NH);
        mbar_init(&bars[kKsideEmpty], NH); mbar_init(&bars[kKsideEmpty + 1], NH);    // one commit per issuer
        for (int hh = 0; hh < 2; ++hh) {
            mbar_init(&bars[kSbReady + hh], 128); mbar_init(&bars[kVnFull + hh], 1);
            mbar_init(&bars[kVnbReady + hh], 128); mbar_init(&bars[kSReady + hh], 1);
            mbar_init(&bars[kOFull + hh], 1); mbar_init(&bars[kOFree + hh], 128);
        }
        fence_mbar_init();
    }
    // H and P are written on and below the diagonal only: the blocks above it stay zero for the whole kernel
    for (int i = tid; i < 1536; i += kThreads) {
        if (i < 1024) reinterpret_cast<uint4*>(smem + kOffPp)[i] = make_uint4(0u, 0u, 0u, 0u);
        else reinterpret_cast<uint4*>(smem + kOffH)[i - 1024] = make_uint4(0u, 0u, 0u, 0u);
    }
    if (V < 128) {      // the value block that is never loaded feeds the idle upper TMEM lanes: keep it zero
        for (int i = tid; i < 1024; i += kThreads)
            reinterpret_cast<uint4*>(smem + kOffV + (i >> 9) * kVSlotBytes + 8192)[i & 511] = make_uint4(0u, 0u, 0u, 0u);
    }
    fence_proxy_async_smem();
    if (warp == 16) {
        tmem_alloc(s_tmem, kTmemCols);
        if (lane == 0) { tma_prefetch_desc(&mq); tma_prefetch_desc(&mk); tma_prefetch_desc(&mv); tma_prefetch_desc(&mo); }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    UM(1, tid == 0);                              // setup done: barriers, zeroed tiles, TMEM allocated
    const uint32_t tmem = *s_tmem;
    // Chunks s_info[4] .. + NC - 1 of the chain; n below counts inside the segment.  NC is the same for every unit (a
    // kernel parameter, so every loop bound stays warp-uniform for the compiler); chunks past the end of the chain in
    // the last segment are exact no-ops: tiles zero-filled by TMA, gates 0, stores clipped.
    // (kVar: per unit, from the table.)
    const int NC = kVar ? (int)s_info[5] : seg_chunks, nc_chain = F * cpf;
    // chunk n of this unit -> tensor-map coordinates (token in frame, frame)
    auto chunk_coord = [&](int n, int& c0, int& f) {
        if constexpr (kVar) { f = 0; c0 = s_info[4] + (n << 6); }
        else { const int m = s_info[4] + n; f = div_cpf.div(m); c0 = (m - f * cpf) << 6; }
    };
#define U_SEG s_info[0]
#define U_CHAIN s_info[1]
#define U_CLIP s_info[2]
#define U_HEAD s_info[3]
#define U_NB s_info[4]

    constexpr uint32_t kIdKK = umma_idesc_bf16(128, 64, false, false);
    constexpr uint32_t kIdMnA = umma_idesc_bf16(128, 64, true, false);          // A = tile^T (MN-major), B K-major
    constexpr uint32_t kIdMnB = umma_idesc_bf16(128, 64, false, true);          // TS, B MN-major
    constexpr uint32_t kIdMnBneg = umma_idesc_bf16(128, 64, false, true, true); // TS, -A, B MN-major

    if (warp < 8) {
        // =========================================================================================
        // K-side group
        // =========================================================================================
        const int wq = warp & 3, wh = warp >> 2;
        const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
        uint8_t* sH = smem + kOffH;
        const uint32_t aH = sbase + kOffH;
        // warp 7 owns the gates: it holds g, beta of the NEXT chunk in registers (loaded one chunk period
        // before they are scanned, so the global-load latency never stalls the group) and scans them while
        // warps 0-3 run the triangular solve.  Lane l: tokens 2l, 2l+1 of the chunk.
        const int64_t g_off = (int64_t)U_CLIP * p.g_stride[0] + (int64_t)U_HEAD * p.g_stride[2];
        const int64_t bt_off = (int64_t)U_CLIP * p.beta_stride[0] + (int64_t)U_HEAD * p.beta_stride[2];
        auto load_gates = [&](int n, float (&gv)[2], float (&bv)[2]) {
            const int m = U_NB + n, f = kVar ? 0 : div_cpf.div(m);
            const int ntok = kVar ? (int)s_info[6] : 0;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int c = kVar ? (n << 6) + 2 * lane + e : ((m - f * cpf) << 6) + 2 * lane + e;
                gv[e] = 0.f; bv[e] = 0.f;                        // pad rows: exact no-ops
                if (kVar ? c < ntok : (c < C && m < nc_chain)) {
                    const int64_t t = kVar ? (int64_t)s_info[4] + c : (int64_t)f * C + c;
                    gv[e] = load_gate(p.g, g_off + t * p.g_stride[1], p.gate_dtype);
                    bv[e] = load_gate(p.beta, bt_off + t * p.beta_stride[1], p.gate_dtype);
                }
            }
        };
        auto scan_chunk = [&](int m, const float (&gv)[2], const float (&bv)[2]) {
            gate_scan(gv, bv, sBt + (m & 1) * 64, sGam + (m & 1) * 64, sE + (m & 3) * 64, sEb + (m & 3) * 32, sCj + (m & 1) * 64, sKd + (m & 3) * 64,
                      sFast + (m & 3), sOfac + (m & 3) * 64, sPost + (m & 3), sPre + (m & 3), scale, lane);
        };
        float g_nx[2] = {0.f, 0.f}, b_nx[2] = {0.f, 0.f};
        if (warp == 7) {
            load_gates(0, g_nx, b_nx);
            scan_chunk(0, g_nx, b_nx);
            if (NC > 1) load_gates(1, g_nx, b_nx);
        }
        kbar();

        PT_DECL
        for (int n = 0; n < NC; ++n) {
            const int st = n & 1;
            uint8_t* sp = smem + kOffKq + (uint32_t)(n % 3) * kKqSlotBytes;   // K | Q tiles of chunk n
            const float* btS = sBt + st * 64;
            // operand buffers of this stage are free once every MMA of chunk n-2 has completed
            if (n >= 2) mbar_wait_inl(&bars[kKsideEmpty + st], (uint32_t)((n >> 1) - 1) & 1u);
            const bool fast = sFast[n & 3] != 0.f;

            // [K;Q]K^T accumulators -> masked A (fp16 solve matrix H) and P (bf16 operand)
            mbar_wait_inl(&bars[kKqTile + n % 3], (uint32_t)(n / 3) & 1u);   // tiles visible to this thread's loads
            mbar_wait_inl(&bars[kKqFull], (uint32_t)n & 1u);
            tc_fence_after_sync();
            PT(0, tid == 0, n);   // waits: buffers free, tiles landed, [K;Q]K^T done
            // Warp (wq, wh) holds accumulator rows 32 (wq & 1) .. + 31 of K K^T (wq < 2) or Q K^T and one 32-column half `ch`:
            // above the diagonal block everything is masked (those tile regions were zeroed once and are never written),
            // below it nothing is, and only the two diagonal blocks pay for the per-element mask.  Warps 0 and 1 take the
            // diagonal blocks of K K^T: exactly what levels 0-1 of the solve on the same warp read, so no barrier in between.
            const int ch = wq == 1 ? (wh ^ 1) : wh;
            if (!ABL(0) && !((wq & 1) == 0 && ch == 1)) {
                uint32_t r[32], pk[16];
                tmem_ld32(lane_addr + kColKQ + ch * 32, r);
                tmem_wait_ld();
                const int j0 = ch * 32;
                const float* Gm = sGam + st * 64;
                const bool diag = (wq & 1) == ch;
                if (!fast) {           // rare: per-element decay exp(Gamma_i - Gamma_j) (and the readout scale) applied first
                    const int i = (wq & 1) * 32 + lane;
                    const float gi = Gm[i], sc = wq < 2 ? 1.f : scale;
#pragma unroll 1
                    for (int jj = 0; jj < 32; jj += 4) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            const float d = sc * __expf(fminf(gi - Gm[j0 + jj + e], 0.f));
                            // r[] must stay in registers: compile-time indices only
                            switch (jj) {
                                case 0: r[e] = __float_as_uint(__uint_as_float(r[e]) * d); break;
                                case 4: r[4 + e] = __float_as_uint(__uint_as_float(r[4 + e]) * d); break;
                                case 8: r[8 + e] = __float_as_uint(__uint_as_float(r[8 + e]) * d); break;
                                case 12: r[12 + e] = __float_as_uint(__uint_as_float(r[12 + e]) * d); break;
                                case 16: r[16 + e] = __float_as_uint(__uint_as_float(r[16 + e]) * d); break;
                                case 20: r[20 + e] = __float_as_uint(__uint_as_float(r[20 + e]) * d); break;
                                case 24: r[24 + e] = __float_as_uint(__uint_as_float(r[24 + e]) * d); break;
                                default: r[28 + e] = __float_as_uint(__uint_as_float(r[28 + e]) * d); break;
                            }
                        }
                    }
                }
                if (wq < 2) {          // rows of K K^T:  A_ij = beta_i (k_i.k_j),  j < i
                    const int i = wq * 32 + lane;
                    const float bi = btS[i];
                    if (diag) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj)
                            pk[jj] = tri::pack_f16(2 * jj < lane ? __uint_as_float(r[2 * jj]) * bi : 0.f, 2 * jj + 1 < lane ? __uint_as_float(r[2 * jj + 1]) * bi : 0.f);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) pk[jj] = tri::pack_f16(__uint_as_float(r[2 * jj]) * bi, __uint_as_float(r[2 * jj + 1]) * bi);
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(sH + sw128_offset(i, ch * 4 + c)) = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                } else {               // rows of Q K^T:  P_ij = (q_i.k_j),  j <= i   (fast: scale e_i is applied in the readout)
                    const int i = (wq - 2) * 32 + lane;
                    if (diag) {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj)
                            pk[jj] = pack_bf16(2 * jj <= lane ? __uint_as_float(r[2 * jj]) : 0.f, 2 * jj + 1 <= lane ? __uint_as_float(r[2 * jj + 1]) : 0.f);
                    } else {
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) pk[jj] = pack_bf16(__uint_as_float(r[2 * jj]), __uint_as_float(r[2 * jj + 1]));
                    }
#pragma unroll
                    for (int c = 0; c < 4; ++c)
                        *reinterpret_cast<uint4*>(smem + kOffPp + st * 8192 + sw128_offset(i, ch * 4 + c)) =
                            make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bars[kKqFree]);
            if (!fast) {   // rare: chunk decay below e^-60.  Q~ = scale e_i Q in place (the swizzle keeps rows intact);
                           // K' = K exp(Gamma_last - Gamma_i) is folded into a second Vnb copy by the state warps
                const float* eS = sE + (n & 3) * 64;
#pragma unroll 1
                for (int idx = tid; idx < 512; idx += kKThreads) {
                    uint4* pq = reinterpret_cast<uint4*>(sp + 8192 + (idx << 4));
                    *pq = scale_row8(*pq, scale * eS[idx >> 3]);
                }
            }
            PT(1, tid == 0, n);   // gating (own work)

            // (I + A)^-1 in place.  Levels 0-1 on warps 0-1 straight after their own gating (they wrote the diagonal 32 x 32
            // blocks themselves); warp 7 scans the gates of chunk n+1 meanwhile.
            if (warp < 2) {
                __syncwarp();
                if (!ABL(1)) tri::solve_levels01(sH, aH, warp, lane);
            } else if (warp == 7 && n + 1 < NC) {
                scan_chunk(n + 1, g_nx, b_nx);
                if (n + 2 < NC) load_gates(n + 2, g_nx, b_nx);
            }
            PT(3, tid == 0, n);   // diagonal blocks + level 1
            kbar();                // A complete (all eight warps), levels 0-1 done
            PT(2, tid == 0, n);   // barrier before level 2
            // T' = X diag(c) -> bf16, K-major swizzled rows: H and T' share one layout (thread: one 16-byte chunk per task)
            auto convert_rows = [&](int task) {
                const float* cjS = sCj + st * 64;
                const int i = task >> 3, c = task & 7;
                const uint32_t off = sw128_offset(i, c);
                const uint4 hx = *reinterpret_cast<const uint4*>(sH + off);
                const float4 c0 = *reinterpret_cast<const float4*>(cjS + c * 8), c1 = *reinterpret_cast<const float4*>(cjS + c * 8 + 4);
                const float2 x0 = tri::unpack_f16(hx.x), x1 = tri::unpack_f16(hx.y), x2 = tri::unpack_f16(hx.z), x3 = tri::unpack_f16(hx.w);
                *reinterpret_cast<uint4*>(smem + kOffTp + st * 8192 + off) =
                    make_uint4(pack_bf16(x0.x * c0.x, x0.y * c0.y), pack_bf16(x1.x * c0.z, x1.y * c0.w),
                               pack_bf16(x2.x * c1.x, x2.y * c1.y), pack_bf16(x3.x * c1.z, x3.y * c1.w));
            };
            // Level 2 (warps 0-3) only writes rows 32-63, columns 0-31: rows 0-31 of X are final, warps 4-7 convert them now
            if (warp < 4) {
                if (!ABL(2)) tri::solve_level2(aH, warp, lane, 5); else named_bar_sync(5, 128);
            } else if (!ABL(3)) {
                convert_rows(tid - 128);            // tasks 0 .. 255: rows 0-31
                convert_rows(tid);
            }
            PT(4, tid == 0, n);   // level 2
            kbar();
            PT(5, tid == 0, n);   // solve barrier
            if (!ABL(3)) convert_rows(256 + tid);   // rows 32-63, one task per thread
            fence_proxy_async_smem();
            kbar();
            PT(6, tid == 0, n);   // T' conversion
            if (n == 0) UM(3, tid == 0);          // K side of the unit's first chunk published
            if (n == NC - 1) UM(4, tid == 0);     // K side of the unit's last chunk published
            if (tid == 0) mbar_arrive(&bars[kTpReady + st]);
        }
    } else if (warp < 16) {
        // =========================================================================================
        // state warpgroups (one per 128 value columns)
        // =========================================================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 120;");
        const int hh = (warp - 8) >> 2, wq = warp & 3, stid = tid - 256 - hh * 128;
        if (hh < NH) {
            const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
            const int vcol = hh * 128 + wq * 32 + lane;
            const int bar_id = 2 + hh;
            {   // initial state -> TMEM: the caller's for the first segment, the previous segment's hand-off otherwise
                uint32_t r[32];
                const int chain = U_CHAIN, seg = U_SEG;
                const bool col_ok = vcol < V;                 // V = 64: lanes 64-127 carry zeros
                const float* s0 = (p.initial_state && col_ok) ? p.initial_state + (int64_t)chain * 64 * V + vcol : nullptr;
                if (seg != 0) {
                    if (stid == 0) {
                        const int* flag = xsync + 1 + chain * 2 + hh;
                        uint32_t polls = 0;
                        while (ld_acquire_gpu(flag) < seg) {
                            __nanosleep(256);
                            if (++polls > (1u << 28)) __trap();      // minutes (a debugger or a time-sliced GPU may stall a predecessor for seconds): it died
                        }
                    }
                    named_bar_sync(bar_id, 128);
                    s0 = col_ok ? xstate + (int64_t)chain * 64 * V + vcol : nullptr;
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = s0 ? __float_as_uint(__ldcg(s0 + (int64_t)(half * 32 + j) * V)) : 0u;
                    tmem_st32(lane_addr + kColS + hh * 64 + half * 32, r);
                }
                tmem_wait_st();
                UM(2, tid == 256);                    // initial state in TMEM
            }
            // drain the readout of chunk m: O^T accumulators in mma-fragment layout (16x256b TMEM loads) -> * scale e_i
            // -> bf16 pairs -> stmatrix.trans into the 128B-swizzled staging tile [2][tok][64] -> TMA store
            // (rows past the frame are clipped by the tensor map)
            PT_DECL
            auto readout = [&](int m) {
                mbar_wait_inl(&bars[kOFull + hh], (uint32_t)m & 1u);
                tc_fence_after_sync();
                PT(23, tid == 256, m);   // wait: readout accumulators (after Vnb -> state update + intra-chunk MMAs)
                if (stid == 0) tma_store_wait_read0();      // previous readout has left the staging buffer
                named_bar_sync(bar_id, 128);
                PT(24, tid == 256, m);   // wait: staging buffer free + group barrier
                const float* of = sOfac + (m & 3) * 64 + 2 * (lane & 3);
                const uint32_t ost_half = sbase + kOffOst + hh * 16384;
                float2 cf[8];       // the same eight column (token) factors for both lane groups: loaded once
#pragma unroll
                for (int q = 0; q < 8; ++q) cf[q] = *reinterpret_cast<const float2*>(of + 8 * q);
#pragma unroll
                for (int grp = 0; grp < (ABL(7) ? 0 : 2); ++grp) {
                    uint32_t r[32];
                    tmem_ld_16x256b_x8(tmem + ((uint32_t)(wq * 32 + grp * 16) << 16) + kColO + hh * 64, r);
                    tmem_wait_ld();
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 c = cf[q];
                        pk[2 * q] = pack_bf16(__uint_as_float(r[4 * q]) * c.x, __uint_as_float(r[4 * q + 1]) * c.y);
                        pk[2 * q + 1] = pack_bf16(__uint_as_float(r[4 * q + 2]) * c.x, __uint_as_float(r[4 * q + 3]) * c.y);
                    }
                    // lane l addresses matrix l/8 (token block +(l/16), value rows +8 ((l/8)&1)), memory row = token l%8
                    const int v0 = wq * 32 + grp * 16 + ((lane >> 3) & 1) * 8;
                    const uint32_t vaddr = ost_half + (v0 >> 6) * 8192;
#pragma unroll
                    for (int q = 0; q < 8; q += 2) {
                        const int tok = 8 * (q + (lane >> 4)) + (lane & 7);
                        stmatrix_x4_trans(vaddr + sw128_offset(tok, (v0 & 63) >> 3), pk[2 * q], pk[2 * q + 1], pk[2 * q + 2], pk[2 * q + 3]);
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(&bars[kOFree + hh]);
                fence_proxy_async_smem();
                named_bar_sync(bar_id, 128);
                if (kVar && (int)s_info[6] - (m << 6) < 64) {
                    // last chunk of a packed sequence: the rows behind it belong to the next sequence, so the valid rows
                    // leave through ordinary 16-byte stores (staging tile: [value block][token][64], 128B swizzle)
                    store_valid_rows(smem + kOffOst + hh * 16384,
                                     reinterpret_cast<__nv_bfloat16*>(p.o) + ((int64_t)s_info[4] + (m << 6)) * p.o_stride[1] +
                                         (int64_t)s_info[3] * p.o_stride[2] + hh * 128,
                                     p.o_stride[1], (int)s_info[6] - (m << 6), stid, V >= 128 ? 2 : 1);
                } else if (stid == 0) {
                    int c0, f;
                    chunk_coord(m, c0, f);
                    tma_store_5d(&mo, smem + kOffOst + hh * 16384, 0, c0, U_HEAD * VB + hh * 2, f, U_CLIP);
                    tma_store_commit();
                }
            };
            // W^T of chunk m: 8 units of 16 key dims x 32 tokens shared by the state warps (needs the K side of chunk m)
            auto wt_operand = [&](int m) {
                const int st = m & 1;
                mbar_wait_inl(&bars[kTpReady + st], (uint32_t)(m >> 1) & 1u);          // K side of chunk m published
                PT(16, tid == 256, m);   // wait: K side of chunk m
                const int sw = (warp - 8);                                          // 0 .. 4 NH - 1
                for (int u = sw; u < (ABL(4) ? 0 : 8); u += 4 * NH)
                    w_unit_mma(sbase + kOffKq + (uint32_t)(m % 3) * kKqSlotBytes, sbase + kOffTp + st * 8192, smem + kOffWt + st * 8192,
                               sEb + (m & 3) * 32, u & 3, u >> 2, lane);
                fence_proxy_async_smem();
                mbar_arrive(&bars[kKsideFull + st]);
                PT(22, tid == 256, m);   // W^T mma.sync
            };
            // Order inside one chunk: only the S pass and the Vnb pass sit between the state-side MMAs of the
            // recurrence; the readout of chunk n-1 runs under the Vn correction MMA and the W^T operand of chunk
            // n+1 under the state update MMA.
            wt_operand(0);
            for (int n = 0; n < NC; ++n) {
                if (n >= 1) mbar_wait_inl(&bars[kSReady + hh], (uint32_t)(n - 1) & 1u);
                tc_fence_after_sync();
                PT(17, tid == 256, n);   // wait: state update of chunk n-1
                {   // S_n = post_{n-1} * accumulator;  Sb = bf16(S_n) (operand copy);  accumulator <- pre_n * S_n
                    const float post = n >= 1 ? sPost[(n - 1) & 3] : 1.f, pre = sPre[n & 3];
#ifdef GDKVM_ABLATE_DECAY      // measurement only (wrong results): what a deferred decay of the state accumulator could save
                    const bool rescale = false;
#else
                    const bool rescale = pre != 1.f || post != 1.f;
#endif
                    if (!ABL(5)) {   // both 32-column halves in flight at once (one TMEM load latency instead of two)
                        uint32_t ra[32], rb[32], pk[32];
                        tmem_ld32(lane_addr + kColS + hh * 64, ra);
                        tmem_ld32(lane_addr + kColS + hh * 64 + 32, rb);
                        tmem_wait_ld();
#ifndef GDKVM_ABLATE_DECAY
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            ra[j] = __float_as_uint(__uint_as_float(ra[j]) * post);
                            rb[j] = __float_as_uint(__uint_as_float(rb[j]) * post);
                        }
#endif
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            pk[j] = pack_bf16(__uint_as_float(ra[2 * j]), __uint_as_float(ra[2 * j + 1]));
                            pk[16 + j] = pack_bf16(__uint_as_float(rb[2 * j]), __uint_as_float(rb[2 * j + 1]));
                        }
                        tmem_st32(lane_addr + kColSb + hh * 32, pk);
                        if constexpr (kSave) {
                            // training forward: the bf16 chunk-start state (exactly the operand copy Sb) is kept for the backward
                            // pass -- batched: [chain][chunk][value column][key dim]; packed clips: [slot][head][value column][key dim]
                            // -- 128 contiguous bytes per thread
                            const int m = kVar ? n : (int)s_info[4] + n;
                            if ((kVar || m < nc_chain) && vcol < V) {
                                const int64_t blk = kVar ? ((int64_t)s_info[8] + n) * p.H + s_info[3] : (int64_t)s_info[1] * nc_chain + m;
                                // 256-bit stores: a whole 32-byte sector per lane and instruction (1.49 -> 1.36 ms for the training
                                // forward at configs[1] against 16-byte stores, profiles/r4y_state_dump_256bit_stores_ab.log)
                                uint8_t* dst = reinterpret_cast<uint8_t*>(sdump + (blk * V + vcol) * 64);
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    st_global_v8(dst + 32 * j, pk[8 * j], pk[8 * j + 1], pk[8 * j + 2], pk[8 * j + 3], pk[8 * j + 4], pk[8 * j + 5],
                                                 pk[8 * j + 6], pk[8 * j + 7]);
                            }
                        }
                        if (pre != 1.f) {      // slow path only
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                ra[j] = __float_as_uint(__uint_as_float(ra[j]) * pre);
                                rb[j] = __float_as_uint(__uint_as_float(rb[j]) * pre);
                            }
                        }
                        if (rescale) {
                            tmem_st32(lane_addr + kColS + hh * 64, ra);
                            tmem_st32(lane_addr + kColS + hh * 64 + 32, rb);
                        }
                    }
                    tmem_wait_st();
                }
                tc_fence_before_sync();
                mbar_arrive(&bars[kSbReady + hh]);
                PT(18, tid == 256, n);   // S pass
                if (n >= 1) readout(n - 1);
                PT(19, tid == 256, n - 1);   // readout of chunk n-1
                mbar_wait_inl(&bars[kVnFull + hh], (uint32_t)n & 1u);
                tc_fence_after_sync();
                PT(20, tid == 256, n);   // wait: Vn
                if (!ABL(6)) {   // Vnb = bf16(Vn^T) written over the first half of Vn (TMEM A-operand); both fp32 halves are
                    // in registers before the bf16 columns overwrite them
                    uint32_t r0[32], r1[32];
                    tmem_ld32(lane_addr + kColVn + hh * 64, r0);
                    tmem_ld32(lane_addr + kColVn + hh * 64 + 32, r1);
                    tmem_wait_ld();
                    uint32_t pk[32];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        pk[j] = pack_bf16(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                        pk[16 + j] = pack_bf16(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                    }
                    tmem_st32(lane_addr + kColVn + hh * 64, pk);
                    if (sFast[n & 3] == 0.f) {   // rare: second copy Vnb diag(exp(Gamma_last - Gamma_j)) for the state update
                        const float* kd = sKd + (n & 3) * 64;
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            pk[j] = pack_bf16(__uint_as_float(r0[2 * j]) * kd[2 * j], __uint_as_float(r0[2 * j + 1]) * kd[2 * j + 1]);
                            pk[16 + j] = pack_bf16(__uint_as_float(r1[2 * j]) * kd[32 + 2 * j], __uint_as_float(r1[2 * j + 1]) * kd[33 + 2 * j]);
                        }
                        tmem_st32(lane_addr + kColVn + hh * 64 + 32, pk);
                    }
                    tmem_wait_st();
                }
                tc_fence_before_sync();
                mbar_arrive(&bars[kVnbReady + hh]);
                PT(21, tid == 256, n);   // Vnb pass
                if (n + 1 < NC) wt_operand(n + 1);
            }
            UM(6, tid == 256);                        // last Vnb pass done
            readout(NC - 1);
            UM(7, tid == 256);                        // last readout drained
            mbar_wait_inl(&bars[kSReady + hh], (uint32_t)(NC - 1) & 1u);
            tc_fence_after_sync();
            UM(8, tid == 256);                        // last state update complete
            const int chain = U_CHAIN, seg = U_SEG;
            const bool last_seg = s_info[7] != 0;
            float* sT = last_seg ? p.final_state : xstate;       // the caller's final state | hand-off to the next segment
            if (sT != nullptr && vcol < V) {
                uint32_t r[32];
                const float post = sPost[(NC - 1) & 3];
                sT += (int64_t)chain * 64 * V + vcol;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tmem_ld32(lane_addr + kColS + hh * 64 + half * 32, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) sT[(int64_t)(half * 32 + j) * V] = __uint_as_float(r[j]) * post;
                }
            }
            if (!last_seg) {                                     // publish: state writes of the 128 threads, then the flag
                __threadfence();
                named_bar_sync(bar_id, 128);
                if (stid == 0) st_release_gpu(xsync + 1 + chain * 2 + hh, seg + 1);
            }
            UM(9, tid == 256);                        // final state stored / handed over
            if (stid == 0) tma_store_wait_all0();
            UM(10, tid == 256);                       // readout stores complete
            tc_fence_before_sync();
        }
    } else {
      asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");      // the whole issuer warpgroup (warps 16-19) at one instruction
      if (warp == 16) {
        // =========================================================================================
        // issuer K: TMA loads (one chunk of prefetch) + the [K;Q]K^T MMA
        // =========================================================================================
        {
            const int b = U_CLIP, h = U_HEAD;
            auto issue_kq = [&](int m) {         // one elected lane arms the barrier and issues the K and Q tile loads
                const int slot = m % 3;
                int c0, f;
                chunk_coord(m, c0, f);
                uint8_t* sp = smem + kOffKq + (uint32_t)slot * kKqSlotBytes;
                if (elect_one()) {
                    mbar_arrive_expect_tx(&bars[kKqTile + slot], 16384u);
                    tma_load_5d(sp, &mk, &bars[kKqTile + slot], 0, c0, f, h, b);
                    tma_load_5d(sp + 8192, &mq, &bars[kKqTile + slot], 0, c0, f, h, b);
                }
                __syncwarp();
            };
            for (int m = 0; m < 2 && m < NC; ++m) {      // prologue: chunks 0 and 1 (their V tiles too)
                issue_kq(m);
                if (elect_one()) {
                    int c0, f;
                    chunk_coord(m, c0, f);
                    for (int hh = 0; hh < NH; ++hh) {
                        uint64_t* vb = &bars[kVTile + m * 2 + hh];
                        mbar_arrive_expect_tx(vb, v_tile_bytes);
                        tma_load_5d(smem + kOffV + m * kVSlotBytes + hh * 16384, &mv, vb, 0, c0, h * VB + hh * 2, f, b);
                    }
                }
                __syncwarp();
            }
#pragma unroll 1
            for (int n = 0; n < NC; ++n) {
                const int slot = n % 3;
                const uint64_t dK = umma_smem_desc_sw128(sbase + kOffKq + (uint32_t)slot * kKqSlotBytes, 16, 1024);
                mbar_wait_inl(&bars[kKqTile + slot], (uint32_t)(n / 3) & 1u);
                if (n >= 1) mbar_wait_inl(&bars[kKqFree], (uint32_t)(n - 1) & 1u);  // accumulators of chunk n-1 drained
                tc_fence_after_sync();
                umma4_ss(tmem + kColKQ, dK, 2, dK, 2, kIdKK, false);                // [K;Q] K^T
                umma_commit_w(&bars[kKqFull]);
                if (n + 2 < NC) {                    // ring slot of chunk n-1 takes chunk n+2 once every MMA of chunk n-1 completed
                    if (n >= 1) mbar_wait_inl(&bars[kKsideEmpty + ((n - 1) & 1)], (uint32_t)((n - 1) >> 1) & 1u);
                    issue_kq(n + 2);
                }
            }
        }
      } else {
        // =========================================================================================
        // issuers S: the state-side MMAs, one warp per value half.  Issuing a tcgen05.mma costs ~100 cycles whenever
        // its descriptors have to be moved into uniform registers first (tests/probes/mma_timing.cu: 33 cycles only
        // for loop-invariant operands), so the 20 MMAs per chunk and half are ~2 000 cycles of one warp.
        // =========================================================================================
        const int hh = warp - 17;
        if (hh < NH) {
            const int b = U_CLIP, hv = U_HEAD * VB + hh * 2;
            PT_DECL
#pragma unroll 1
            for (int n = 0; n < NC; ++n) {
                const int st = n & 1;
                const uint32_t aKt = sbase + kOffKq + (uint32_t)(n % 3) * kKqSlotBytes, aQt = aKt + 8192;
                const uint32_t aVt = sbase + kOffV + st * kVSlotBytes + hh * 16384;
                const uint64_t dTp = umma_smem_desc_sw128(sbase + kOffTp + st * 8192, 16, 1024);
                const uint64_t dWt = umma_smem_desc_sw128(sbase + kOffWt + st * 8192, 8192, 1024);
                const uint64_t dKp = umma_smem_desc_sw128(aKt, 8192, 1024);      // raw K tile as the MN-major B of the state update
                const uint64_t dPp = umma_smem_desc_sw128(sbase + kOffPp + st * 8192, 16, 1024);
                const uint64_t dQt = umma_smem_desc_sw128(aQt, 16, 1024);
                const uint32_t tVn = tmem + kColVn + hh * 64, tSb = tmem + kColSb + hh * 32, tO = tmem + kColO + hh * 64, tS = tmem + kColS + hh * 64;
                const uint32_t par2 = (uint32_t)(n >> 1) & 1u, par1 = (uint32_t)n & 1u;
                mbar_wait_inl(&bars[kTpReady + st], par2);
                PT(32, lane == 0 && hh == 0, n);   // issuer S: wait K side
                const uint32_t vn_off = sFast[n & 3] != 0.f ? 0u : 32u;   // slow path: the state update reads the second Vnb copy
                mbar_wait_inl(&bars[kVTile + st * 2 + hh], par2);
                tc_fence_after_sync();
                umma4_ss(tVn, umma_smem_desc_sw128(aVt, 8192, 1024), 128, dTp, 2, kIdMnA, false);   // Vn^T = V^T T'^T (after the MMAs of chunk n-1)
                PT(38, lane == 0 && hh == 0, n);   // issuer S: issue U
                mbar_wait_inl(&bars[kKsideFull + st], par2);
                PT(33, lane == 0 && hh == 0, n);   // issuer S: wait W^T operand
                mbar_wait_inl(&bars[kSbReady + hh], par1);
                tc_fence_after_sync();
                umma4_ts_commit(tVn, tSb, dWt, 128, kIdMnBneg, true, &bars[kVnFull + hh]);          // Vn^T -= Sb W^T
                PT(34, lane == 0 && hh == 0, n);   // issuer S: wait Sb, issue Vn correction
                if (n >= 1) mbar_wait_inl(&bars[kOFree + hh], par1 ^ 1u);
                tc_fence_after_sync();
                umma4_ts(tO, tSb, dQt, 2, kIdKK, false);                                             // O^T = Sb Q~^T
                PT(35, lane == 0 && hh == 0, n);   // issuer S: wait O free, issue inter-chunk readout
                mbar_wait_inl(&bars[kVnbReady + hh], par1);
                tc_fence_after_sync();
                umma4_ts_commit(tS, tVn + vn_off, dKp, 128, kIdMnB, true, &bars[kSReady + hh]);     // S^T += Vnb K'
                umma4_ts_commit(tO, tVn, dPp, 2, kIdKK, true, &bars[kOFull + hh], &bars[kKsideEmpty + st]);   // O^T += Vnb P^T; one of NH arrivals
                if (n + 2 < NC) {   // U of chunk n completed before Vnb was published: its V half-tile slot takes chunk n + 2
                    if (elect_one()) {
                        int c0, f;
                        chunk_coord(n + 2, c0, f);
                        uint64_t* vb = &bars[kVTile + st * 2 + hh];
                        mbar_arrive_expect_tx(vb, v_tile_bytes);
                        tma_load_5d(smem + kOffV + st * kVSlotBytes + hh * 16384, &mv, vb, 0, c0, hv, f, b);
                    }
                    __syncwarp();
                }
                PT(37, lane == 0 && hh == 0, n);   // issuer S: wait Vnb, issue state update + intra-chunk readout
            }
        }
        __syncwarp();
      }
    }

    // ---- teardown ----
    tc_fence_before_sync();
    __syncthreads();
    UM(11, tid == 0);
    if (warp == 16) tmem_dealloc(tmem, kTmemCols);
    UM(12, tid == 512);                           // warp 16: TMEM released
}

bool mult16(int64_t elems) { return (elems * 2) % 16 == 0; }

// Number of time segments per chain: list-schedule the units (ticket order, one unit per SM at a time, a unit
// occupies its SM while it waits for its predecessor) for every candidate count and keep the shortest makespan.
// Unit cost = its chunks + kUnitOverhead chunk periods of set-up, pipeline fill and drain.
constexpr double kUnitOverhead = 2.5;
constexpr int kMaxSegments = 8;

double simulate_units(int chains, int nc, int sms, int seg_chunks) {
    const int nseg = (nc + seg_chunks - 1) / seg_chunks;
    std::priority_queue<double, std::vector<double>, std::greater<double>> sm_free;
    for (int i = 0; i < sms; ++i) sm_free.push(0.0);
    std::vector<double> done((size_t)chains, 0.0);
    double makespan = 0.0;
    for (int s = 0; s < nseg; ++s) {
        const double len = (double)std::min(seg_chunks, nc - s * seg_chunks) + kUnitOverhead;
        for (int c = 0; c < chains; ++c) {
            const double t0 = sm_free.top();
            sm_free.pop();
            const double end = std::max(t0, done[(size_t)c]) + len;
            done[(size_t)c] = end;
            sm_free.push(end);
            makespan = std::max(makespan, end);
        }
    }
    return makespan;
}

int pick_segments(int chains, int nc, int sms) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int>, int> cache;
    const auto key = std::make_tuple(chains, nc, sms);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) return it->second;
    int best = 1;
    double best_t = simulate_units(chains, nc, sms, nc);
    if ((int64_t)chains * kMaxSegments <= (1 << 16)) {          // keep the one-off simulation cheap
        for (int s = 2; s <= kMaxSegments && s <= nc / 8; ++s) {   // at least 8 chunks per segment
            const int sc = (nc + s - 1) / s;
            const double t = simulate_units(chains, nc, sms, sc);
            if (t < best_t * 0.98) { best_t = t; best = (nc + sc - 1) / sc; }   // cut only for a real gain
        }
    }
    if (cache.size() > 4096) cache.clear();
    cache[key] = best;
    return best;
}

// Mixed plan for a batch of equal-length chains that fills more than one wave: whole waves of UNCUT chains (a unit boundary costs
// ~2.85 chunk periods -- scripts/unit_overhead.py -- so every cut is paid for), and only the clips left over for the ragged last
// wave are cut, into about one unit per SM.  Ticket order: the first segments of the cut clips, the uncut clips, then the later
// segments of the cut clips level by level (a predecessor always holds a smaller ticket and has long finished).  configs[1]:
// 55 clips uncut + 9 clips in two segments = 4 units per SM instead of 7.  Returns the simulated makespan (chunk periods) and the
// plan, or a negative value when there is nothing to gain over the uniform plan `uniform_t`.
struct MixedPlan { int uncut_clips = 0, cut_clips = 0, nseg = 1, seg_chunks = 0; };

double simulate_mixed(int H, int nc, int sms, const MixedPlan& m) {
    std::priority_queue<double, std::vector<double>, std::greater<double>> sm_free;
    for (int i = 0; i < sms; ++i) sm_free.push(0.0);
    std::vector<double> done((size_t)m.cut_clips * H, 0.0);
    double makespan = 0.0;
    auto run = [&](double len, double pred) {
        const double t0 = sm_free.top();
        sm_free.pop();
        const double end = std::max(t0, pred) + len;
        sm_free.push(end);
        makespan = std::max(makespan, end);
        return end;
    };
    const int nsegs = (nc + m.seg_chunks - 1) / m.seg_chunks;
    for (int c = 0; c < m.cut_clips * H; ++c) done[(size_t)c] = run(std::min(m.seg_chunks, nc) + kUnitOverhead, 0.0);
    for (int c = 0; c < m.uncut_clips * H; ++c) run(nc + kUnitOverhead, 0.0);
    for (int l = 1; l < nsegs; ++l)
        for (int c = 0; c < m.cut_clips * H; ++c)
            done[(size_t)c] = run(std::min(m.seg_chunks, nc - l * m.seg_chunks) + kUnitOverhead, done[(size_t)c]);
    return makespan;
}

double pick_mixed(int B, int H, int nc, int sms, double uniform_t, MixedPlan* out) {
    const int chains = B * H, waves = chains / sms;
    if (waves < 1 || nc < 16 || (int64_t)chains * kMaxSegments > (1 << 16)) return -1.0;
    MixedPlan best;
    double best_t = -1.0;
    // candidates: the clips of `waves` (or one fewer) whole waves stay uncut
    for (int w = waves; w >= std::max(1, waves - 1); --w) {
        MixedPlan m;
        m.uncut_clips = std::min(B, (w * sms) / H);
        m.cut_clips = B - m.uncut_clips;
        if (m.cut_clips == 0) continue;
        const int target = std::max(1, (int)((double)sms / (m.cut_clips * H) + 0.5));
        for (int s = std::max(1, target - 1); s <= std::min({target + 1, kMaxSegments, nc / 8}); ++s) {
            m.seg_chunks = (nc + s - 1) / s;
            m.nseg = (nc + m.seg_chunks - 1) / m.seg_chunks;
            const double t = simulate_mixed(H, nc, sms, m);
            if (best_t < 0 || t < best_t) { best_t = t; best = m; }
        }
    }
    if (best_t < 0 || best_t > uniform_t * 0.985) return -1.0;       // only for a real gain
    *out = best;
    return best_t;
}

// Per-device facts and resources, created on first use under one mutex: SM count, memory-pool support, the dynamic
// shared-memory opt-in of both kernel instantiations (cudaFuncSetAttribute is per device/context, so a process that
// drives several GPUs needs it on each), and a library-PRIVATE stream-ordered memory pool for the hand-off scratch of
// cut launches.  The pool keeps its blocks across synchronisations (release threshold = max) so the scratch is re-used
// launch after launch; the process-wide default pool and its policy are never touched.
struct DeviceCtx {
    bool init = false;
    int sms = 148;
    bool mempools = false;
    cudaMemPool_t pool = nullptr;      // nullptr: fall back to the default pool (unchanged policy)
    cudaError_t attr_err = cudaErrorUnknown;
};

DeviceCtx* device_ctx() {
    static std::mutex mu;
    static DeviceCtx ctx[64];
    static DeviceCtx overflow;         // device ordinals >= 64: re-initialised on every call (no caching)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
    std::lock_guard<std::mutex> lk(mu);
    DeviceCtx* c = (dev >= 0 && dev < 64) ? &ctx[dev] : &overflow;
    if (c == &overflow) *c = DeviceCtx();
    if (!c->init) {
        int n = 0, ok = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) c->sms = n;
        if (cudaDeviceGetAttribute(&ok, cudaDevAttrMemoryPoolsSupported, dev) == cudaSuccess) c->mempools = ok != 0;
        if (c->mempools) {
            cudaMemPoolProps props = {};
            props.allocType = cudaMemAllocationTypePinned;
            props.handleTypes = cudaMemHandleTypeNone;
            props.location.type = cudaMemLocationTypeDevice;
            props.location.id = dev;
            cudaMemPool_t pool = nullptr;
            if (cudaMemPoolCreate(&pool, &props) == cudaSuccess) {
                unsigned long long keep = ~0ull;
                cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
                c->pool = pool;
            }
        }
        (void)cudaGetLastError();
        c->init = true;
    }
    if (c->attr_err != cudaSuccess) {      // retried until it has succeeded on THIS device
        cudaError_t e = cudaFuncSetAttribute(gdr_chunk_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gdr_chunk_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gdr_chunk_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gdr_chunk_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (e != cudaSuccess) (void)cudaGetLastError();
        c->attr_err = e;
    }
    return c;
}

// stream-ordered scratch of one launch: from the library's own pool when there is one
cudaError_t scratch_alloc(void** ws, size_t bytes, const DeviceCtx* c, cudaStream_t stream) {
    if (c != nullptr && c->pool != nullptr) return cudaMallocFromPoolAsync(ws, bytes, c->pool, stream);
    return cudaMallocAsync(ws, bytes, stream);
}

}  // namespace

#ifdef GDKVM_PHASE_TIMERS
}  // namespace gdkvm
// debug-only export of the profiling build: accumulated cycles per phase slot of CTA 0 (and reset)
extern "C" int gdkvm_debug_phase_cycles(unsigned long long* out, int n) {
    unsigned long long h[64];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, gdkvm::g_phase_cycles, sizeof h) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 64; ++i) out[i] = h[i];
    unsigned long long z[64] = {0};
    cudaMemcpyToSymbol(gdkvm::g_phase_cycles, z, sizeof z);
    return 0;
}
extern "C" int gdkvm_debug_unit_marks(long long* out, int n) {
    long long h[16];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, gdkvm::g_unit_marks, sizeof h) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 16; ++i) out[i] = h[i];
    return 0;
}
extern "C" int gdkvm_debug_phase_trace(long long* out, int n) {   // out[64][8]
    long long h[64 * 8];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, gdkvm::g_phase_trace, sizeof h) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 64 * 8; ++i) out[i] = h[i];
    return 0;
}
namespace gdkvm {
#endif

// Why the tcgen05 chunk kernel cannot take this problem ("" when it can).  Host arithmetic only.
const char* chunked_unsupported_reason(const GdkvmGdrParams& p) {
    if (p.io_dtype != GDKVM_BF16) return "q/k/v/o are fp32: the tcgen05 chunk kernel takes bf16 I/O (fp32 I/O runs the fp32 CUDA-core kernel)";
    if (p.K != 64) return "d_k != 64: the tcgen05 chunk kernel is built for d_k = 64";
    if (p.V != 64 && p.V != 128 && p.V != 256) return "d_v not in {64, 128, 256}";
    if (p.T <= 0) return "no tokens";
    // TMA: 16-byte aligned bases and strides; the value/readout head stride must equal V so that
    // (head, 64-wide value block) folds into one tensor-map dimension.
    const void* ptrs[4] = {p.q, p.k, p.v, p.o};
    for (const void* x : ptrs) if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return "q/k/v/o base pointers must be 16-byte aligned for TMA";
    for (int i = 0; i < 3; ++i)
        if (!mult16(p.q_stride[i]) || !mult16(p.k_stride[i]) || !mult16(p.v_stride[i]) || !mult16(p.o_stride[i]))
            return "q/k/v/o strides must be multiples of 16 bytes for TMA";
    if (p.v_stride[2] != p.V || p.o_stride[2] != p.V) return "v/o head stride must equal d_v (heads contiguous inside a token)";
    if (p.q_stride[1] <= 0 || p.k_stride[1] <= 0 || p.v_stride[1] <= 0 || p.o_stride[1] <= 0) return "token strides must be positive";
    return "";
}

bool chunked_supports(const GdkvmGdrParams& p) { return chunked_unsupported_reason(p)[0] == '\0'; }

int chunked_segments(const GdkvmGdrParams& p, int sms) {
    const bool flat = p.frame_tokens <= 0 || (p.flags & GDKVM_FLAG_FLAT_CHUNKS) ||
                      (p.frame_tokens % 64 != 0 && !(p.flags & GDKVM_FLAG_FRAME_CHUNKS));
    const int C = flat ? p.T : p.frame_tokens, nc = (p.T / C) * ((C + 63) / 64), chains = p.B * p.H;
    int nseg = (int)((p.flags >> 8) & 0xfu);
    if (nseg == 0) nseg = pick_segments(chains, nc, sms > 0 ? sms : 148);
    nseg = std::max(1, std::min(nseg, nc));
    const int seg_chunks = (nc + nseg - 1) / nseg;
    return (nc + seg_chunks - 1) / seg_chunks;                     // no empty segment
}

namespace {

// Tensor maps of q, k (dk, token-in-frame, frame, head, clip) and v, o (64 values, token-in-frame, head x value block,
// frame, clip): C tokens per frame, F frames; one 128-column value half per v / o box (each state warpgroup loads and
// stores its own half).
int make_maps(const GdkvmGdrParams& p, int C, int F, CUtensorMap* mq, CUtensorMap* mk, CUtensorMap* mv, CUtensorMap* mo) {
    const uint64_t B = p.B, H = p.H, V = p.V;
    {
        const uint64_t dims[5] = {64, (uint64_t)C, (uint64_t)F, H, B};
        const uint32_t box[5] = {64, 64, 1, 1, 1};
        const uint64_t sq[4] = {(uint64_t)p.q_stride[1] * 2, (uint64_t)p.q_stride[1] * 2 * C, (uint64_t)p.q_stride[2] * 2, (uint64_t)p.q_stride[0] * 2};
        const uint64_t sk[4] = {(uint64_t)p.k_stride[1] * 2, (uint64_t)p.k_stride[1] * 2 * C, (uint64_t)p.k_stride[2] * 2, (uint64_t)p.k_stride[0] * 2};
        int rc = make_tmap(mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.q, dims, sq, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.k, dims, sk, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    {
        const uint64_t dims[5] = {64, (uint64_t)C, H * (V / 64), (uint64_t)F, B};
        const uint32_t box[5] = {64, 64, V >= 128 ? 2u : 1u, 1, 1};     // V = 64: one value block per head
        const uint64_t sv[4] = {(uint64_t)p.v_stride[1] * 2, 128, (uint64_t)p.v_stride[1] * 2 * C, (uint64_t)p.v_stride[0] * 2};
        const uint64_t so[4] = {(uint64_t)p.o_stride[1] * 2, 128, (uint64_t)p.o_stride[1] * 2 * C, (uint64_t)p.o_stride[0] * 2};
        int rc = make_tmap(mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.v, dims, sv, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.o, dims, so, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    return 0;
}

// Work-unit table of the packed variable-length call, built on the device from the sequence offsets (one block).
// Level l holds segment l of every sequence that has one, in sequence order; levels follow each other, so a unit's
// predecessor (same sequence, level l - 1) always has a smaller ticket.  Sequences without tokens get no unit: their
// final state is the initial state (copied here).
template <typename IdxT>
__global__ void __launch_bounds__(256) gdr_units_kernel(const IdxT* __restrict__ cu, int nseq, int seg_chunks, int max_entries,
                                                        int* __restrict__ utab, int H, int state_elems,
                                                        const float* __restrict__ s0, float* __restrict__ sT) {
    __shared__ int s_warp[8];
    __shared__ int s_base, s_levels;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) { s_base = 0; s_levels = 0; }
    __syncthreads();
    // a sequence of c chunks is cut into round(c / seg_chunks) (>= 1) segments of equal size (no short remainder unit)
    auto chunks_of = [&](int n) {
        const long long len = (long long)cu[n + 1] - (long long)cu[n];
        return len > 0 ? (int)((len + 63) >> 6) : 0;
    };
    auto seg_size = [&](int chunks) {            // chunks per segment (chunks > 0)
        const int want = max(1, (chunks + seg_chunks / 2) / seg_chunks);
        return (chunks + want - 1) / want;
    };
    auto segs_of = [&](int n) {                  // every segment non-empty: (segs - 1) * size < chunks
        const int chunks = chunks_of(n);
        return chunks == 0 ? 0 : (chunks + seg_size(chunks) - 1) / seg_size(chunks);
    };
    int mx = 0;
    for (int n = tid; n < nseq; n += 256) mx = max(mx, segs_of(n));
    atomicMax(&s_levels, mx);
    __syncthreads();
    const int levels = s_levels;
    for (int lvl = 0; lvl < levels; ++lvl) {
        for (int n0 = 0; n0 < nseq; n0 += 256) {
            const int n = n0 + tid;
            const int nsegs = n < nseq ? segs_of(n) : 0;
            const bool has = nsegs > lvl;
            const unsigned m = __ballot_sync(0xffffffffu, has);
            if (lane == 0) s_warp[warp] = __popc(m);
            __syncthreads();
            int pos = s_base + __popc(m & ((1u << lane) - 1u)), total = 0;
            for (int w = 0; w < 8; ++w) { if (w < warp) pos += s_warp[w]; total += s_warp[w]; }
            if (has && pos < max_entries) {
                const long long t0 = (long long)cu[n], len = (long long)cu[n + 1] - t0;
                const int sc = seg_size(chunks_of(n));                       // chunks per segment of this sequence
                const int first = lvl * sc;                                  // first chunk of this segment
                const long long rest = len - (long long)first * 64;         // tokens from there to the end of the sequence
                const bool last = lvl == nsegs - 1;
                int* e = utab + kUnitInts * (1 + pos);
                e[0] = n; e[1] = (int)(t0 + (long long)first * 64);
                e[2] = last ? (int)rest : sc * 64;
                e[3] = last ? (int)((rest + 63) >> 6) : sc;
                e[4] = lvl; e[5] = last ? 1 : 0; e[7] = 0;
                e[6] = (int)(t0 >> 6) + n + first;      // training forward: chunk-state slot of the unit's first chunk (unique: see gdkvm_gdr.h)
            }
            __syncthreads();
            if (tid == 0) s_base += total;
            __syncthreads();
        }
    }
    if (tid == 0) utab[0] = min(s_base, max_entries);
    if (sT != nullptr) {
        for (int n = 0; n < nseq; ++n) {
            if ((long long)cu[n + 1] - (long long)cu[n] > 0) continue;
            const size_t off = (size_t)n * H * state_elems;
            for (int i = tid; i < H * state_elems; i += 256) sT[off + i] = s0 != nullptr ? s0[off + i] : 0.f;
        }
    }
}

// Work-unit table of the mixed plan (pick_mixed): clips [0, uncut) whole, clips [uncut, B) in `nsegs` segments of `sc` chunks;
// every clip has T tokens at packed offset clip * T.  Entry layout as in gdr_units_kernel.
__global__ void __launch_bounds__(256) gdr_units_mixed_kernel(int B, int T, int uncut, int nsegs, int sc, int* __restrict__ utab,
                                                              int* __restrict__ xsync, int sync_ints) {
    const int cut = B - uncut, entries = cut + uncut + cut * (nsegs - 1);
    for (int i = threadIdx.x; i < sync_ints; i += 256) xsync[i] = 0;          // ticket counter and hand-off flags (no separate memset)
    for (int pos = threadIdx.x; pos < entries; pos += 256) {
        int n, lvl, segs, size;
        if (pos < cut) { n = uncut + pos; lvl = 0; segs = nsegs; size = sc; }
        else if (pos < cut + uncut) { n = pos - cut; lvl = 0; segs = 1; size = (T + 63) >> 6; }
        else { const int v = pos - cut - uncut; lvl = 1 + v / cut; n = uncut + v % cut; segs = nsegs; size = sc; }
        const long long t0 = (long long)n * T;
        const int first = lvl * size;
        const long long rest = (long long)T - (long long)first * 64;
        const bool last = lvl == segs - 1;
        int* e = utab + kUnitInts * (1 + pos);
        e[0] = n; e[1] = (int)(t0 + (long long)first * 64);
        e[2] = last ? (int)rest : size * 64;
        e[3] = last ? (int)((rest + 63) >> 6) : size;
        e[4] = lvl; e[5] = last ? 1 : 0; e[6] = 0; e[7] = 0;
    }
    if (threadIdx.x == 0) utab[0] = entries;
}

}  // namespace

int plan_time_segments(int chains, int chunks, int sms) { return pick_segments(chains, chunks, sms > 0 ? sms : 148); }

int library_scratch_alloc(void** ws, size_t bytes, cudaStream_t stream, int* sms, bool* mempools) {
    const DeviceCtx* dc = device_ctx();
    if (dc == nullptr) return (int)cudaErrorInvalidDevice;
    if (sms != nullptr) *sms = dc->sms;
    if (mempools != nullptr) *mempools = dc->mempools;
    if (ws == nullptr) return 0;
    const cudaError_t e = scratch_alloc(ws, bytes, dc, stream);
    if (e != cudaSuccess) (void)cudaGetLastError();
    return (int)e;
}

namespace {

bool pick_mixed_cached(int B, int H, int nc, int sms, int uniform_seg_chunks, MixedPlan* out) {
    static std::mutex mu;
    static std::map<std::tuple<int, int, int, int>, MixedPlan> cache;       // cut_clips == 0: the uniform plan stays
    const auto key = std::make_tuple(B, H, nc, sms);
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it == cache.end()) {
        MixedPlan m;
        const double tu = simulate_units(B * H, nc, sms, uniform_seg_chunks);
        if (pick_mixed(B, H, nc, sms, tu, &m) < 0) m = MixedPlan();
        if (cache.size() > 4096) cache.clear();
        it = cache.emplace(key, m).first;
    }
    *out = it->second;
    return out->cut_clips > 0;
}

// Whether an (inference) launch of this problem takes the mixed plan, and which.
bool mixed_plan_for(const GdkvmGdrParams& p, int sms, MixedPlan* mp) {
    // flat 64-token tiling -- or frames of whole 64-token chunks, whose chunks are the flat ones
    const bool flat = p.frame_tokens <= 0 || (p.flags & GDKVM_FLAG_FLAT_CHUNKS) || (p.frame_tokens % 64 != 0 && !(p.flags & GDKVM_FLAG_FRAME_CHUNKS)) ||
                      (p.frame_tokens % 64 == 0 && p.T % p.frame_tokens == 0);
    if (!flat || ((p.flags >> 8) & 0xfu) != 0 || p.B <= 1 || (int64_t)p.B * p.T >= ((int64_t)1 << 31)) return false;
    if (p.q_stride[0] != (int64_t)p.T * p.q_stride[1] || p.k_stride[0] != (int64_t)p.T * p.k_stride[1] ||
        p.v_stride[0] != (int64_t)p.T * p.v_stride[1] || p.o_stride[0] != (int64_t)p.T * p.o_stride[1] ||
        p.g_stride[0] != (int64_t)p.T * p.g_stride[1] || p.beta_stride[0] != (int64_t)p.T * p.beta_stride[1])
        return false;
    const int nc = (p.T + 63) / 64, ns = pick_segments(p.B * p.H, nc, sms), sc = (nc + ns - 1) / ns;
    return pick_mixed_cached(p.B, p.H, nc, sms, sc, mp);
}

// The batch as ONE packed token stream (clip n = rows n T .. (n + 1) T - 1) through the unit-table kernel with the mixed plan's table.
int launch_chunked_mixed(const GdkvmGdrParams& p, const MixedPlan& mp, const DeviceCtx* dc, cudaStream_t stream) {
    GdkvmGdrParams pp = p;
    pp.B = 1;
    pp.T = p.B * p.T;
    CUtensorMap mq, mk, mv, mo;
    const int me = make_maps(pp, pp.T, 1, &mq, &mk, &mv, &mo);
    if (me != 0) return me;
    const int H = p.H, V = p.V, chains = p.B * H;
    const int entries = mp.cut_clips * mp.nseg + mp.uncut_clips;
    const size_t state_bytes = (size_t)chains * 64 * V * sizeof(float);
    const size_t sync_bytes = (((size_t)chains * 2 + 1) * sizeof(int) + 15) & ~(size_t)15;
    const size_t tab_bytes = (size_t)kUnitInts * (1 + (size_t)entries) * sizeof(int);
    void* ws = nullptr;
    cudaError_t e = scratch_alloc(&ws, state_bytes + sync_bytes + tab_bytes, dc, stream);
    if (e != cudaSuccess) return (int)e;
    float* xstate = reinterpret_cast<float*>(ws);
    int* xsync = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + state_bytes);
    int* utab = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + state_bytes + sync_bytes);
    gdr_units_mixed_kernel<<<1, 256, 0, stream>>>(p.B, p.T, mp.uncut_clips, mp.nseg, mp.seg_chunks, utab, xsync, (int)(sync_bytes / sizeof(int)));
    count_launch();
    gdr_chunk_kernel<true, false><<<entries * H, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, pp, pp.T, 1, FastDiv::make(1u), 0, 0, xstate, xsync,
                                                                                utab, nullptr);
    count_launch();
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, stream);
    return (int)le;
}

}  // namespace

// Work units of an inference launch of this problem: out = {units, uncut clips, cut clips, segments of a cut clip}; returns 1 for
// the mixed plan, 0 for the uniform one (out[1] = 0, out[2] = B, out[3] = segments per chain).
int chunked_plan_units(const GdkvmGdrParams& p, int sms, int out[4]) {
    sms = sms > 0 ? sms : 148;
    MixedPlan mp;
    if (mixed_plan_for(p, sms, &mp)) {
        out[0] = (mp.uncut_clips + mp.cut_clips * mp.nseg) * p.H; out[1] = mp.uncut_clips; out[2] = mp.cut_clips; out[3] = mp.nseg;
        return 1;
    }
    const int ns = chunked_segments(p, sms);
    out[0] = p.B * p.H * ns; out[1] = 0; out[2] = p.B; out[3] = ns;
    return 0;
}

int launch_chunked(const GdkvmGdrParams& p, cudaStream_t stream, void* chunk_states) {
    const DeviceCtx* dc = device_ctx();
    if (dc == nullptr) return (int)cudaErrorInvalidDevice;
    if (dc->attr_err != cudaSuccess) return (int)dc->attr_err;
    // frame-aligned chunks when frames are whole 64-token chunks (or when asked for); otherwise tile the
    // flat token stream: identical results (token-causal recurrence), no zero-padded rows to process
    // (a training forward that keeps the chunk-start states always tiles the flat stream: the backward pass does too)
    const bool flat = chunk_states != nullptr || p.frame_tokens <= 0 || (p.flags & GDKVM_FLAG_FLAT_CHUNKS) ||
                      (p.frame_tokens % 64 != 0 && !(p.flags & GDKVM_FLAG_FRAME_CHUNKS));
    const int C = flat ? p.T : p.frame_tokens;
    const int F = p.T / C;
    const uint64_t V = p.V;
    CUtensorMap mq, mk, mv, mo;
    const int me = make_maps(p, C, F, &mq, &mk, &mv, &mo);
    if (me != 0) return me;
    // time segments (see the header comment): explicit count in flags bits 8-11, else the simulated optimum; a launch
    // under stream capture stays uncut (no workspace allocation inside a graph)
    const int chains = p.B * p.H, cpf = (C + 63) / 64, nc = F * cpf;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(stream, &cap) != cudaSuccess) { (void)cudaGetLastError(); cap = cudaStreamCaptureStatusActive; }
    // under capture the scratch becomes allocation / free nodes of the graph (stream-ordered allocator); without memory-pool
    // support a captured launch stays uncut
    const bool capturing = cap != cudaStreamCaptureStatusNone;
    GdkvmGdrParams pf = p;
    if (chunk_states != nullptr) { pf.flags |= GDKVM_FLAG_FLAT_CHUNKS; pf.flags &= ~GDKVM_FLAG_FRAME_CHUNKS; }
    int nseg = (capturing && !dc->mempools) ? 1 : chunked_segments(pf, dc->sms);
    int seg_chunks = (nc + nseg - 1) / nseg;
    // more than one wave of equal chains, flat tiling, the library's own choice of segments, a batch that is one packed token
    // stream in memory: cut only the clips of the ragged last wave (pick_mixed) and run through the unit-table kernel
    if (chunk_states == nullptr && !(capturing && !dc->mempools)) {
        MixedPlan mp;
        if (mixed_plan_for(p, dc->sms, &mp)) return launch_chunked_mixed(p, mp, dc, stream);
    }
    float* xstate = nullptr;
    int* xsync = nullptr;
    void* ws = nullptr;
    if (nseg > 1) {
        const size_t state_bytes = (size_t)chains * 64 * V * sizeof(float), sync_bytes = ((size_t)chains * 2 + 1) * sizeof(int);
        if (scratch_alloc(&ws, state_bytes + sync_bytes, dc, stream) != cudaSuccess) {
            (void)cudaGetLastError();
            ws = nullptr; nseg = 1; seg_chunks = nc;                // same kernel, uncut chains
        } else {
            xstate = reinterpret_cast<float*>(ws);
            xsync = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + state_bytes);
            const cudaError_t me2 = cudaMemsetAsync(xsync, 0, sync_bytes, stream);
            if (me2 != cudaSuccess) { cudaFreeAsync(ws, stream); return (int)me2; }
        }
    }
    if (chunk_states != nullptr)
        gdr_chunk_kernel<false, true><<<chains * nseg, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, C, F, FastDiv::make((unsigned)cpf),
                                                                                       nseg, seg_chunks, xstate, xsync, nullptr,
                                                                                       reinterpret_cast<__nv_bfloat16*>(chunk_states));
    else
        gdr_chunk_kernel<false, false><<<chains * nseg, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, C, F, FastDiv::make((unsigned)cpf),
                                                                                        nseg, seg_chunks, xstate, xsync, nullptr, nullptr);
    count_launch();
    const cudaError_t le = cudaGetLastError();
    if (ws != nullptr) cudaFreeAsync(ws, stream);
    return (int)le;
}

int chunked_varlen_seg_chunks(const GdkvmGdrParams& p, int nseq, int sms) {
    // the lengths live on the device: plan for the average sequence (longer ones simply get more units of this size)
    const int avg = std::max(1, (int)(((int64_t)p.T / std::max(1, nseq) + 63) / 64));
    int nseg = (int)((p.flags >> 8) & 0xfu);
    if (nseg == 0) nseg = pick_segments(nseq * p.H, avg, sms > 0 ? sms : 148);
    nseg = std::max(1, std::min(nseg, avg));
    return (avg + nseg - 1) / nseg;
}

// Packed variable-length sequences: q,k,v,o [1, T, H, *], sequence n = rows cu[n] .. cu[n+1]-1 (offsets on the device,
// int32 or int64), states [nseq, H, K, V].
int launch_chunked_varlen(const GdkvmGdrParams& p, const void* cu, int cu_bytes, int nseq, cudaStream_t stream, void* chunk_states) {
    const DeviceCtx* dc = device_ctx();
    if (dc == nullptr) return (int)cudaErrorInvalidDevice;
    if (dc->attr_err != cudaSuccess) return (int)dc->attr_err;
    CUtensorMap mq, mk, mv, mo;
    const int me = make_maps(p, p.T, 1, &mq, &mk, &mv, &mo);
    if (me != 0) return me;
    const int H = p.H, V = p.V, chains = nseq * H;
    const int seg_chunks = chunked_varlen_seg_chunks(p, nseq, dc->sms);
    // sum over sequences of round(chunks_n / seg_chunks)  <=  nseq + (sum of chunks) / seg_chunks,  sum of chunks <= T / 64 + nseq
    const int64_t max_entries64 = (int64_t)nseq + ((int64_t)p.T / 64 + nseq) / seg_chunks + 1;
    if (max_entries64 * H > 0x3fffffff) return (int)cudaErrorInvalidValue;
    const int max_entries = (int)max_entries64;
    const size_t state_bytes = (size_t)chains * 64 * V * sizeof(float);
    const size_t sync_bytes = (((size_t)chains * 2 + 1) * sizeof(int) + 15) & ~(size_t)15;
    const size_t tab_bytes = (size_t)kUnitInts * (1 + (size_t)max_entries) * sizeof(int);
    void* ws = nullptr;
    cudaError_t e = scratch_alloc(&ws, state_bytes + sync_bytes + tab_bytes, dc, stream);
    if (e != cudaSuccess) return (int)e;
    float* xstate = reinterpret_cast<float*>(ws);
    int* xsync = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + state_bytes);
    int* utab = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(ws) + state_bytes + sync_bytes);
    e = cudaMemsetAsync(xsync, 0, sync_bytes, stream);
    if (e != cudaSuccess) { cudaFreeAsync(ws, stream); return (int)e; }
    if (cu_bytes == 8)
        gdr_units_kernel<long long><<<1, 256, 0, stream>>>(reinterpret_cast<const long long*>(cu), nseq, seg_chunks, max_entries, utab, H,
                                                           64 * V, p.initial_state, p.final_state);
    else
        gdr_units_kernel<int><<<1, 256, 0, stream>>>(reinterpret_cast<const int*>(cu), nseq, seg_chunks, max_entries, utab, H, 64 * V,
                                                     p.initial_state, p.final_state);
    count_launch();
    if (chunk_states != nullptr)
        gdr_chunk_kernel<true, true><<<max_entries * H, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, p.T, 1, FastDiv::make(1u), 0, 0, xstate,
                                                                                        xsync, utab, reinterpret_cast<__nv_bfloat16*>(chunk_states));
    else
        gdr_chunk_kernel<true, false><<<max_entries * H, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, p.T, 1, FastDiv::make(1u), 0, 0,
                                                                                         xstate, xsync, utab, nullptr);
    count_launch();
    const cudaError_t le = cudaGetLastError();
    cudaFreeAsync(ws, stream);
    return (int)le;
}

}  // namespace gdkvm
