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
// Roles (18 warps):
//   warps 0-7   K-side group: everything that does not depend on the state, one chunk AHEAD of the
//               state side: [K;Q]K^T accumulators -> masked/gated A (fp32) and P (bf16); the
//               triangular inverse (I + A)^-1 in fp32-grade arithmetic (16x16 forward substitution on
//               CUDA cores, block merges as 3xTF32 mma.sync); T' = T diag(..) ; W^T ; the in-place row
//               scalings K~, Q~ and K' of the TMA tiles (the 128B swizzle keeps rows intact).
//   warps 8-11 / 12-15   state warpgroups, one per 128 value columns: S -> (Sb, gamma S), Vn -> Vnb,
//               readout O -> bf16 -> staging -> TMA store; initial / final state.
//   warp 16     issuer K: TMA loads (q, k, v tiles, one chunk of prefetch) and the K-side MMAs.
//   warp 17     issuer S: the five state-side MMAs per value half.
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
// Layout facts used here were verified on hardware by tests/probes/umma_probe.cu.
// Math: oracle/gdr_ref.py::gdr_chunk_ref.
#include <mutex>

#include "gdr_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace gdkvm {
namespace {

using namespace sm100;

constexpr int kKThreads = 256;                 // K-side group
constexpr int kThreads = 18 * 32;              // whole CTA
constexpr int kPitchA = 68;   // fp32 pitch of the 64x64 solve matrix: conflict-free mma A-fragment loads
constexpr int kPitchY = 40;   // pitch of the merge scratch: conflict-free mma B-fragment loads

// ---- shared memory map (bytes from a 1024-aligned base) ----
constexpr uint32_t kStageBytes = 49152;          // Kt 8K | Qt 8K | Vt 32K   (Qt must follow Kt: stacked [K;Q] operand)
constexpr uint32_t kOffKt = 0, kOffQt = 8192, kOffVt = 16384;
constexpr uint32_t kOffKp = 2 * kStageBytes;     // K'  [2]  (B of the state update, MN-major)
constexpr uint32_t kOffPp = kOffKp + 16384;      // P   [2]  (B of the intra-chunk readout, K-major)
constexpr uint32_t kOffWt = kOffPp + 16384;      // W^T [2]  (B of the state correction, MN-major)
constexpr uint32_t kOffTp = kOffWt + 16384;      // T'  [2]  (B of U / W, K-major)
constexpr uint32_t kOffOst = kOffTp + 16384;     // readout staging, per value half [2][64 tok][64] bf16
constexpr uint32_t kOffA = kOffOst + 32768;      // fp32 solve matrix
constexpr uint32_t kOffY = kOffA + 64 * kPitchA * 4;
constexpr uint32_t kOffF = kOffY + 32 * kPitchY * 4;
//   floats: g[2][64] beta[2][64] Gam[2][64] E[4][64] Cj[2][64] Kd[2][64] Ofac[4][64] post[4] pre[4] fast[2] pad[2]
constexpr uint32_t kNumFloats = 18 * 64 + 12;
constexpr uint32_t kOffBar = kOffF + kNumFloats * 4;
constexpr uint32_t kNumBars = 26;
constexpr uint32_t kSmemBytes = kOffBar + kNumBars * 8 + 16 + 1024;   // + tmem slot + alignment slack
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
    kTmaFull = 0,    // [2] tiles of a chunk landed                      (tx)      -> issuer K, K group
    kKqFull = 2,     //     [K;Q]K^T accumulators complete                (commit)  -> K group
    kKqFree = 3,     //     [K;Q]K^T accumulators drained                 (256)     -> issuer K
    kTpReady = 4,    // [2] K side done: T', P, decay factors published   (1)       -> issuer S (U), state groups (W)
    kKsideFull = 6,  // [2] W^T operand written by the state groups       (128 NH)  -> issuer S
    kKsideEmpty = 8, // [2] every MMA of the chunk completed              (commit)  -> K group (operand buffers free)
    kD1Done = 10,    //     tile stage no longer read by any MMA          (commit)  -> issuer K (TMA refill)
    kSbReady = 11,   // [2] per half: Sb + decayed S in TMEM              (128)     -> issuer S
    kVnFull = 13,    // [2] per half: Vn^T complete                       (commit)  -> state group
    kVnbReady = 15,  // [2] per half: Vnb in TMEM                         (128)     -> issuer S
    kSReady = 17,    // [2] per half: state update complete               (commit)  -> state group
    kOFull = 19,     // [2] per half: readout accumulators complete       (commit)  -> state group
    kOFree = 21,     // [2] per half: readout accumulators drained        (128)     -> issuer S
    kKpFull = 23,    // [2] second copy of the K tile landed               (tx)      -> issuer S (state update), K group (slow path)
};

// ---- optional phase timers (build with -DGDKVM_PHASE_TIMERS: scripts/phase_timers.py) ----
#ifdef GDKVM_PHASE_TIMERS
__device__ unsigned long long g_phase_cycles[64];
#define PT_DECL long long pt_prev = clock64();
#define PT(slot, cond)                                                                 \
    do {                                                                               \
        if (blockIdx.x == 0 && (cond)) {                                               \
            const long long pt_now = clock64();                                        \
            atomicAdd(&g_phase_cycles[slot], (unsigned long long)(pt_now - pt_prev));  \
            pt_prev = pt_now;                                                          \
        }                                                                              \
    } while (0)
#else
#define PT_DECL
#define PT(slot, cond) do { } while (0)
#endif

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void kbar() { named_bar_sync(1, kKThreads); }
// second K-group barrier (id 4) for the middle of the 16x16 merge, where warp 7 only arrives: a warp must
// never arrive twice in one phase of the same barrier, so this cannot share id 1 with the full syncs
__device__ __forceinline__ void kbar_mid() { named_bar_sync(4, kKThreads); }
__device__ __forceinline__ void kbar_mid_arrive() { asm volatile("bar.arrive 4, %0;" ::"r"(kKThreads) : "memory"); }

__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t w, float s) {
    const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
    return pack_bf16(lo * s, hi * s);
}
__device__ __forceinline__ uint4 scale_row8(uint4 v, float s) {
    return make_uint4(scale_bf16x2(v.x, s), scale_bf16x2(v.y, s), scale_bf16x2(v.z, s), scale_bf16x2(v.w, s));
}

// ---- warp-level MMA on the legacy tensor path: tiny K-side products that live in registers ----
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
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

// c(16x8) += A[0..15][k] * B[k][0..7] over the 16-wide k slices that intersect [kbeg, kend); A, B are
// row-major fp32 in shared memory (pitches pa, pb), rounded to fp16 on the fly (11 significant bits;
// every entry is O(1) and the result is rounded to bf16 afterwards -- tests/chunk_numerics_model.py).
// Entries outside the triangular supports are exact zeros in memory, so partial slices are harmless.
template <int KS>
__device__ __forceinline__ void tile_mma_f16(float (&c)[4], const float* A, int pa, const float* B, int pb,
                                             int kbeg, int kend, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
        const int k0 = ks * 16;
        if (k0 + 16 > kbeg && k0 < kend) {
            const float2 a0 = *reinterpret_cast<const float2*>(A + g * pa + k0 + 2 * t);
            const float2 a1 = *reinterpret_cast<const float2*>(A + (g + 8) * pa + k0 + 2 * t);
            const float2 a2 = *reinterpret_cast<const float2*>(A + g * pa + k0 + 2 * t + 8);
            const float2 a3 = *reinterpret_cast<const float2*>(A + (g + 8) * pa + k0 + 2 * t + 8);
            const uint32_t af[4] = {pack_f16(a0.x, a0.y), pack_f16(a1.x, a1.y), pack_f16(a2.x, a2.y), pack_f16(a3.x, a3.y)};
            const float* Bk = B + (k0 + 2 * t) * pb + g;
            mma_f16(c, af, pack_f16(Bk[0], Bk[pb]), pack_f16(Bk[8 * pb], Bk[9 * pb]));
        }
    }
}

// X21 <- -X22 (L21 X11) for NP independent pairs of adjacent N x N diagonal blocks of the unit
// lower-triangular matrix held in sA (in place; sY is scratch).  Called by the whole K group.
// With LAZY7, warp 7 (which has no tile) only arrives at the middle barrier: it runs the gate scan of
// the next chunk across both stages.
template <int N, int NP, bool LAZY7>
__device__ __forceinline__ void tri_merge(float* sA, float* sY, int warp, int lane) {
    constexpr int NT = N / 8, TPP = (N / 16) * NT, TILES = NP * TPP;
    const int g = lane >> 2, t = lane & 3;
    const int pair = warp / TPP, tile = warp % TPP, mt = tile / NT, nt = tile % NT;
    const int o1 = pair * 2 * N, o2 = o1 + N;
    float* Y = sY + pair * (N * kPitchY);
    if (warp < TILES) {      // Y = L21 X11   (X11 lower triangular: rows k < 8 nt contribute nothing)
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        tile_mma_f16<N / 16>(c, sA + (o2 + mt * 16) * kPitchA + o1, kPitchA, sA + o1 * kPitchA + o1 + nt * 8, kPitchA, nt * 8, N, lane);
        *reinterpret_cast<float2*>(Y + (mt * 16 + g) * kPitchY + nt * 8 + 2 * t) = make_float2(c[0], c[1]);
        *reinterpret_cast<float2*>(Y + (mt * 16 + g + 8) * kPitchY + nt * 8 + 2 * t) = make_float2(c[2], c[3]);
    }
    if (LAZY7) { if (warp != 7) kbar_mid(); } else kbar();
    if (warp < TILES) {      // X21 = -X22 Y  (X22 lower triangular: columns k > row contribute nothing)
        float c[4] = {0.f, 0.f, 0.f, 0.f};
        tile_mma_f16<N / 16>(c, sA + (o2 + mt * 16) * kPitchA + o2, kPitchA, Y + nt * 8, kPitchY, 0, mt * 16 + 16, lane);
        float* X21 = sA + (o2 + mt * 16) * kPitchA + o1 + nt * 8 + 2 * t;
        *reinterpret_cast<float2*>(X21 + g * kPitchA) = make_float2(-c[0], -c[1]);
        *reinterpret_cast<float2*>(X21 + (g + 8) * kPitchA) = make_float2(-c[2], -c[3]);
    }
    kbar();
}

// Gate scan of one chunk (one warp): Gamma = cumsum(g), decay factors, fast/slow decision.
__device__ __forceinline__ void gate_scan(const float* gS, const float* btS, float* sGam, float* sE, float* sCj,
                                          float* sKd, float* sFast, float* ofac, float* post, float* pre,
                                          float scale, int lane) {
    const float g0 = gS[2 * lane], g1 = gS[2 * lane + 1];
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
    sGam[2 * lane] = G0; sGam[2 * lane + 1] = G1;
    sE[2 * lane] = e0; sE[2 * lane + 1] = e1;
    sCj[2 * lane] = btS[2 * lane] * (fast ? __expf(-G0) : 1.f);          // column factor of T'
    sCj[2 * lane + 1] = btS[2 * lane + 1] * (fast ? __expf(-G1) : 1.f);
    sKd[2 * lane] = __expf(Gl - G0);                                      // slow path: row factor of K'
    sKd[2 * lane + 1] = __expf(Gl - G1);
    ofac[2 * lane] = fast ? scale * e0 : 1.f;                             // readout row factor
    ofac[2 * lane + 1] = fast ? scale * e1 : 1.f;
    if (lane == 0) { *post = fast ? gam : 1.f; *pre = fast ? 1.f : gam; *sFast = fast ? 1.f : 0.f; }
}

// One 16 (key dims d) x 32 (tokens i) unit of  W^T[d][i] = sum_j (K[j][d] e_j) T'[i][j]  in registers:
// bf16 mma.sync with ldmatrix from the 128B-swizzled K tile (transposed) and T' tile; e_j is applied to
// the K fragments; the result is written as the bf16 MN-major operand rows of the Vn correction MMA
// (row = key dim d, contiguous over tokens i).  T' is lower triangular: slices with j > i are skipped.
__device__ __forceinline__ void w_unit_mma(uint32_t aK, uint32_t aT, uint8_t* wt, const float* eS, int mt, int ng, int lane) {
    const int g = lane >> 2, t = lane & 3;
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) acc[a][c] = 0.f;
#pragma unroll 1
    for (int ks = 0; ks < 4; ++ks) {
        if (ks * 16 <= ng * 32 + 31) {
            uint32_t af[4], bf0[4], bf1[4];
            ldmatrix_x4_trans(af, aK + sw128_offset(ks * 16 + (lane & 7) + ((lane >> 4) & 1) * 8, mt * 2 + ((lane >> 3) & 1)));
            {   // K~ = K e_j: fragment registers hold tokens j = 16 ks + 2t (+1) and + 8
                const float2 e01 = *reinterpret_cast<const float2*>(eS + ks * 16 + 2 * t);
                const float2 e89 = *reinterpret_cast<const float2*>(eS + ks * 16 + 2 * t + 8);
                af[0] = pack_bf16(__uint_as_float(af[0] << 16) * e01.x, __uint_as_float(af[0] & 0xffff0000u) * e01.y);
                af[1] = pack_bf16(__uint_as_float(af[1] << 16) * e01.x, __uint_as_float(af[1] & 0xffff0000u) * e01.y);
                af[2] = pack_bf16(__uint_as_float(af[2] << 16) * e89.x, __uint_as_float(af[2] & 0xffff0000u) * e89.y);
                af[3] = pack_bf16(__uint_as_float(af[3] << 16) * e89.x, __uint_as_float(af[3] & 0xffff0000u) * e89.y);
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

// four 128x64x16 tcgen05 MMAs covering K = 64; descriptors advance by a fixed step per K slice
__device__ __forceinline__ void umma4_ss(uint32_t d, uint64_t a, uint32_t astep, uint64_t bdesc, uint32_t bstep, uint32_t idesc, bool acc0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ss(d, a + (uint64_t)(k * astep), bdesc + (uint64_t)(k * bstep), idesc, acc0 || k > 0);
}
__device__ __forceinline__ void umma4_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t bstep, uint32_t idesc, bool acc0) {
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_ts(d, a_tmem + k * 8, bdesc + (uint64_t)(k * bstep), idesc, acc0 || k > 0);
}

__global__ void __launch_bounds__(kThreads, 1)
gdr_chunk_kernel(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mk,
                 const __grid_constant__ CUtensorMap mv, const __grid_constant__ CUtensorMap mo,
                 const GdkvmGdrParams p, const int C, const int F) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // 1024-align inside the shared window with pointer arithmetic only (an integer round trip would
    // demote every access below from LDS/STS to generic LD/ST)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    float* sA = reinterpret_cast<float*>(smem + kOffA);
    float* sY = reinterpret_cast<float*>(smem + kOffY);
    float* sG = reinterpret_cast<float*>(smem + kOffF);   // [2][64] log gates of the chunk in each stage
    float* sBt = sG + 128;                                // [2][64] beta
    float* sGam = sBt + 128;                              // [2][64] Gamma_i (inclusive cumsum of g)
    float* sE = sGam + 128;                               // [4][64] exp(Gamma_i) of chunk n in slot n & 3
    float* sCj = sE + 256;                                // [2][64] column factor of T': f_j beta_j (fast) | beta_j (slow)
    float* sKd = sCj + 128;                               // [2][64] row factor of K':    gamma (fast) | exp(Gamma_last - Gamma_i) (slow)
    float* sOfac = sKd + 128;                             // [4][64] readout row factor of chunk n in slot n & 3: scale e_i | 1
    float* sPost = sOfac + 256;                           // [4] factor applied to S AFTER chunk n's accumulate: gamma | 1
    float* sPre = sPost + 4;                              // [4] factor applied to S BEFORE chunk n's accumulate: 1 | gamma
    float* sFast = sPre + 4;                              // [2]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + kNumBars);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int chain = blockIdx.x, b = chain / p.H, h = chain % p.H;
    const int V = p.V, NH = V >> 7, VB = V >> 6;
    const int cpf = (C + 63) >> 6, NC = F * cpf;
    const float scale = p.scale;

    if (tid == 0) {
        mbar_init(&bars[kTmaFull], 1); mbar_init(&bars[kTmaFull + 1], 1);
        mbar_init(&bars[kKqFull], 1); mbar_init(&bars[kKqFree], kKThreads);
        mbar_init(&bars[kTpReady], 1); mbar_init(&bars[kTpReady + 1], 1);
        mbar_init(&bars[kKsideFull], 128 * NH); mbar_init(&bars[kKsideFull + 1], 128 * NH);
        mbar_init(&bars[kKsideEmpty], 1); mbar_init(&bars[kKsideEmpty + 1], 1);
        mbar_init(&bars[kD1Done], 1);
        mbar_init(&bars[kKpFull], 1); mbar_init(&bars[kKpFull + 1], 1);
        for (int hh = 0; hh < 2; ++hh) {
            mbar_init(&bars[kSbReady + hh], 128); mbar_init(&bars[kVnFull + hh], 1);
            mbar_init(&bars[kVnbReady + hh], 128); mbar_init(&bars[kSReady + hh], 1);
            mbar_init(&bars[kOFull + hh], 1); mbar_init(&bars[kOFree + hh], 128);
        }
        fence_mbar_init();
    }
    if (warp == 16) {
        tmem_alloc(s_tmem, kTmemCols);
        if (lane == 0) { tma_prefetch_desc(&mq); tma_prefetch_desc(&mk); tma_prefetch_desc(&mv); tma_prefetch_desc(&mo); }
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

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
        const int64_t g_off = (int64_t)b * p.g_stride[0] + (int64_t)h * p.g_stride[2];
        const int64_t bt_off = (int64_t)b * p.beta_stride[0] + (int64_t)h * p.beta_stride[2];
        auto load_gates = [&](int n, float& gv, float& bv) {   // tid < 64: g, beta of row tid of chunk n
            const int f = n / cpf, c = ((n - f * cpf) << 6) + tid;
            gv = 0.f; bv = 0.f;                                  // pad rows: exact no-ops
            if (c < C) {
                const int64_t t = (int64_t)f * C + c;
                gv = load_gate(p.g, g_off + t * p.g_stride[1], p.gate_dtype);
                bv = load_gate(p.beta, bt_off + t * p.beta_stride[1], p.gate_dtype);
            }
        };
        if (tid < 64) { float gv, bv; load_gates(0, gv, bv); sG[tid] = gv; sBt[tid] = bv; }
        kbar();
        if (warp == 7) gate_scan(sG, sBt, sGam, sE, sCj, sKd, sFast, sOfac, sPost, sPre, scale, lane);   // chunk 0 -> slot 0
        kbar();

        PT_DECL
        for (int n = 0; n < NC; ++n) {
            const int st = n & 1;
            uint8_t* sp = smem + st * kStageBytes;
            const float* btS = sBt + st * 64;
            const float* eS = sE + (n & 3) * 64;
            float g_next = 0.f, b_next = 0.f;
            if (n + 1 < NC && tid < 64) load_gates(n + 1, g_next, b_next);
            // operand buffers of this stage are free once every MMA of chunk n-2 has completed
            if (n >= 2) mbar_wait(&bars[kKsideEmpty + st], (uint32_t)((n >> 1) - 1) & 1u);
            const bool fast = sFast[st] != 0.f;

            // [K;Q]K^T accumulators -> masked A (fp32 solve matrix) and P (bf16 operand), row factors only
            mbar_wait(&bars[kTmaFull + st], (uint32_t)(n >> 1) & 1u);   // tiles visible to this thread's loads
            mbar_wait(&bars[kKqFull], (uint32_t)n & 1u);
            tc_fence_after_sync();
            PT(0, tid == 0);   // waits: buffers free, tiles landed, [K;Q]K^T done
#pragma unroll 1
            for (int c8 = 0; c8 < 4; ++c8) {      // 8 accumulator columns per step keeps the loop body in the L0 i-cache
                uint32_t r[8];
                tmem_ld8(lane_addr + kColKQ + wh * 32 + c8 * 8, r);
                tmem_wait_ld();
                const int j0 = wh * 32 + c8 * 8;
                if (wq < 2) {          // rows of K K^T:  A_ij = beta_i (k_i.k_j),  j < i
                    const int i = wq * 32 + lane;
                    const float bi = btS[i];
                    float o[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) o[jj] = (j0 + jj) < i ? __uint_as_float(r[jj]) * bi : 0.f;
                    *reinterpret_cast<float4*>(sA + i * kPitchA + j0) = make_float4(o[0], o[1], o[2], o[3]);
                    *reinterpret_cast<float4*>(sA + i * kPitchA + j0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
                } else {               // rows of Q K^T:  P_ij = (q_i.k_j),  j <= i   (scale e_i is applied in the readout)
                    const int i = (wq - 2) * 32 + lane;
                    const float sce = fast ? 1.f : scale;
                    float o[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) o[jj] = (j0 + jj) <= i ? __uint_as_float(r[jj]) * sce : 0.f;
                    *reinterpret_cast<uint4*>(smem + kOffPp + st * 8192 + sw128_offset(i, wh * 4 + c8)) =
                        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                }
            }
            tc_fence_before_sync();
            mbar_arrive(&bars[kKqFree]);
            PT(1, tid == 0);   // gating (own work)
            kbar();
            PT(2, tid == 0);   // gating (barrier wait)
            if (!fast) {   // rare: chunk decay below e^-60 -> per-element exp(Gamma_i - Gamma_j) instead of folded factors
                const float* Gm = sGam + st * 64;
#pragma unroll 1
                for (int idx = tid; idx < 4096; idx += kKThreads) {
                    const int i = idx >> 6, j = idx & 63;
                    if (j < i) sA[i * kPitchA + j] *= __expf(Gm[i] - Gm[j]);
                }
#pragma unroll 1
                for (int idx = tid; idx < 2048; idx += kKThreads) {
                    const int i = idx >> 5, j = (idx & 31) * 2;
                    uint32_t* w = reinterpret_cast<uint32_t*>(smem + kOffPp + st * 8192 + sw128_offset(i, j >> 3) + (j & 7) * 2);
                    const uint32_t v = *w;
                    *w = pack_bf16(__uint_as_float(v << 16) * __expf(fminf(Gm[i] - Gm[j], 0.f)),
                                   __uint_as_float(v & 0xffff0000u) * __expf(fminf(Gm[i] - Gm[j + 1], 0.f)));
                }
                mbar_wait(&bars[kKpFull + st], (uint32_t)(n >> 1) & 1u);
                const float* kdS = sKd + st * 64;
#pragma unroll 1
                for (int idx = tid; idx < 512; idx += kKThreads) {   // (row, 16B chunk): the swizzle keeps rows intact
                    const int row = idx >> 3, off = idx << 4;
                    uint4* pk = reinterpret_cast<uint4*>(smem + kOffKp + st * 8192 + off);
                    uint4* pq = reinterpret_cast<uint4*>(sp + kOffQt + off);
                    *pk = scale_row8(*pk, kdS[row]);                 // K' = K exp(Gamma_last - Gamma_i)
                    *pq = scale_row8(*pq, scale * eS[row]);          // Q~ = scale e_i Q
                }
                kbar();
            }

            // 16x16 diagonal blocks by forward substitution (warps 0-1)
            if (warp < 2) {
                const int blk = tid >> 4, c = tid & 15;
                float* Ab = sA + (blk * 16) * kPitchA + blk * 16;
                float x[16], acc[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) acc[i] = 0.f;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    x[j] = (j == c) ? 1.f : -acc[j];
#pragma unroll
                    for (int i = j + 1; i < 16; ++i) acc[i] = fmaf(Ab[i * kPitchA + j], x[j], acc[i]);
                }
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 16; ++i) Ab[i * kPitchA + c] = x[i];
                if (n + 1 < NC) { sG[(st ^ 1) * 64 + tid] = g_next; sBt[(st ^ 1) * 64 + tid] = b_next; }
            }
            PT(3, tid == 0);   // diagonal 16x16 inverses (own work)
            kbar();
            // block merges; warp 7 has no 16x16 merge tile and scans the gates of chunk n+1 meanwhile
            if (warp == 7) kbar_mid_arrive();    // middle barrier of the 16x16 merge: warp 7 is not waited for
            if (warp == 7 && n + 1 < NC)
                gate_scan(sG + (st ^ 1) * 64, sBt + (st ^ 1) * 64, sGam + (st ^ 1) * 64, sE + ((n + 1) & 3) * 64, sCj + (st ^ 1) * 64,
                          sKd + (st ^ 1) * 64, sFast + (st ^ 1), sOfac + ((n + 1) & 3) * 64, sPost + ((n + 1) & 3), sPre + ((n + 1) & 3),
                          scale, lane);
            tri_merge<16, 2, true>(sA, sY, warp, lane);
            PT(4, tid == 0);   // 16x16 merges
            tri_merge<32, 1, false>(sA, sY, warp, lane);
            PT(5, tid == 0);   // 32x32 merge
            {   // T' = X diag(c) -> bf16, K-major swizzled rows (thread: 4 consecutive columns of one row, 4 tasks)
                const float* cjS = sCj + st * 64;
#pragma unroll 1
                for (int it = 0; it < 4; ++it) {
                    const int task = it * kKThreads + tid, i = task >> 4, q4 = task & 15;
                    const float4 x = *reinterpret_cast<const float4*>(sA + i * kPitchA + q4 * 4);
                    const float4 cj = *reinterpret_cast<const float4*>(cjS + q4 * 4);
                    *reinterpret_cast<uint2*>(smem + kOffTp + st * 8192 + sw128_offset(i, q4 >> 1) + (q4 & 1) * 8) =
                        make_uint2(pack_bf16(x.x * cj.x, x.y * cj.y), pack_bf16(x.z * cj.z, x.w * cj.w));
                }
            }
            fence_proxy_async_smem();
            kbar();
            PT(6, tid == 0);   // T' conversion
            if (tid == 0) mbar_arrive(&bars[kTpReady + st]);

        }
    } else if (warp < 16) {
        // =========================================================================================
        // state warpgroups (one per 128 value columns)
        // =========================================================================================
        const int hh = (warp - 8) >> 2, wq = warp & 3, stid = tid - 256 - hh * 128;
        if (hh < NH) {
            const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
            const int vcol = hh * 128 + wq * 32 + lane;
            const int bar_id = 2 + hh;
            {   // initial state -> TMEM
                uint32_t r[32];
                const float* s0 = p.initial_state ? p.initial_state + (int64_t)chain * 64 * V + vcol : nullptr;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) r[j] = s0 ? __float_as_uint(__ldg(s0 + (int64_t)(half * 32 + j) * V)) : 0u;
                    tmem_st32(lane_addr + kColS + hh * 64 + half * 32, r);
                }
                tmem_wait_st();
            }
            // drain the readout of chunk m: O^T accumulators in mma-fragment layout (16x256b TMEM loads) -> * scale e_i
            // -> bf16 pairs -> stmatrix.trans into the 128B-swizzled staging tile [2][tok][64] -> TMA store
            // (rows past the frame are clipped by the tensor map)
            PT_DECL
            auto readout = [&](int m) {
                mbar_wait(&bars[kOFull + hh], (uint32_t)m & 1u);
                tc_fence_after_sync();
                PT(23, tid == 256);   // wait: readout accumulators (after Vnb -> state update + intra-chunk MMAs)
                if (stid == 0) tma_store_wait_read0();      // previous readout has left the staging buffer
                named_bar_sync(bar_id, 128);
                PT(24, tid == 256);   // wait: staging buffer free + group barrier
                const float* of = sOfac + (m & 3) * 64 + 2 * (lane & 3);
                const uint32_t ost_half = sbase + kOffOst + hh * 16384;
#pragma unroll
                for (int grp = 0; grp < 2; ++grp) {
                    uint32_t r[32];
                    tmem_ld_16x256b_x8(tmem + ((uint32_t)(wq * 32 + grp * 16) << 16) + kColO + hh * 64, r);
                    tmem_wait_ld();
                    uint32_t pk[16];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const float2 c = *reinterpret_cast<const float2*>(of + 8 * q);
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
                if (stid == 0) {
                    const int f = m / cpf, c0 = (m - f * cpf) << 6;
                    tma_store_5d(&mo, smem + kOffOst + hh * 16384, 0, c0, h * VB + hh * 2, f, b);
                    tma_store_commit();
                }
            };
            for (int n = 0; n < NC; ++n) {
                const int st = n & 1;
                if (n >= 1) readout(n - 1);                                         // while the K side of chunk n finishes
                PT(19, tid == 256);   // readout of chunk n-1
                mbar_wait(&bars[kTpReady + st], (uint32_t)(n >> 1) & 1u);          // K side of chunk n published
                PT(16, tid == 256);   // wait: K side of chunk n
                {   // W^T of chunk n: 8 units of 16 key dims x 32 tokens shared by the state warps
                    const int sw = (warp - 8);                                      // 0 .. 4 NH - 1
                    for (int u = sw; u < 8; u += 4 * NH)
                        w_unit_mma(sbase + st * kStageBytes + kOffKt, sbase + kOffTp + st * 8192, smem + kOffWt + st * 8192,
                                   sE + (n & 3) * 64, u & 3, u >> 2, lane);
                    fence_proxy_async_smem();
                    mbar_arrive(&bars[kKsideFull + st]);
                }
                PT(22, tid == 256);   // W^T mma.sync
                if (n >= 1) mbar_wait(&bars[kSReady + hh], (uint32_t)(n - 1) & 1u);
                tc_fence_after_sync();
                PT(17, tid == 256);   // wait: state update of chunk n-1
                {   // S_n = post_{n-1} * accumulator;  Sb = bf16(S_n) (operand copy);  accumulator <- pre_n * S_n
                    const float post = n >= 1 ? sPost[(n - 1) & 3] : 1.f, pre = sPre[n & 3];
                    const bool rescale = pre != 1.f || post != 1.f;
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        uint32_t r[32], pk[16];
                        tmem_ld32(lane_addr + kColS + hh * 64 + half * 32, r);
                        tmem_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * post);
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                        tmem_st16(lane_addr + kColSb + hh * 32 + half * 16, pk);
                        if (rescale) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * pre);
                            tmem_st32(lane_addr + kColS + hh * 64 + half * 32, r);
                        }
                    }
                    tmem_wait_st();
                }
                tc_fence_before_sync();
                mbar_arrive(&bars[kSbReady + hh]);
                PT(18, tid == 256);   // S pass
                mbar_wait(&bars[kVnFull + hh], (uint32_t)n & 1u);
                tc_fence_after_sync();
                PT(20, tid == 256);   // wait: Vn
                {   // Vnb = bf16(Vn^T) written over the first half of Vn (TMEM A-operand); both fp32 halves are
                    // in registers before the bf16 columns overwrite them
                    uint32_t r0[32], r1[32], pk[32];
                    tmem_ld32(lane_addr + kColVn + hh * 64, r0);
                    tmem_ld32(lane_addr + kColVn + hh * 64 + 32, r1);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        pk[j] = pack_bf16(__uint_as_float(r0[2 * j]), __uint_as_float(r0[2 * j + 1]));
                        pk[16 + j] = pack_bf16(__uint_as_float(r1[2 * j]), __uint_as_float(r1[2 * j + 1]));
                    }
                    tmem_st32(lane_addr + kColVn + hh * 64, pk);
                    tmem_wait_st();
                }
                tc_fence_before_sync();
                mbar_arrive(&bars[kVnbReady + hh]);
                PT(21, tid == 256);   // Vnb pass
            }
            readout(NC - 1);
            mbar_wait(&bars[kSReady + hh], (uint32_t)(NC - 1) & 1u);
            tc_fence_after_sync();
            if (p.final_state != nullptr) {
                uint32_t r[32];
                const float post = sPost[(NC - 1) & 3];
                float* sT = p.final_state + (int64_t)chain * 64 * V + vcol;
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    tmem_ld32(lane_addr + kColS + hh * 64 + half * 32, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int j = 0; j < 32; ++j) sT[(int64_t)(half * 32 + j) * V] = __uint_as_float(r[j]) * post;
                }
            }
            if (stid == 0) tma_store_wait_all0();
            tc_fence_before_sync();
        }
    } else if (warp == 16) {
        // =========================================================================================
        // issuer K: TMA loads (one chunk of prefetch) + the [K;Q]K^T MMA
        // =========================================================================================
        if (lane == 0) {
            const uint32_t stage_tx = 16384u + (uint32_t)VB * 8192u;
            auto issue_loads = [&](int n) {
                const int st = n & 1, f = n / cpf, c0 = (n - f * cpf) << 6;
                uint8_t* sp = smem + st * kStageBytes;
                mbar_arrive_expect_tx(&bars[kTmaFull + st], stage_tx);
                tma_load_5d(sp + kOffKt, &mk, &bars[kTmaFull + st], 0, c0, f, h, b);
                tma_load_5d(sp + kOffQt, &mq, &bars[kTmaFull + st], 0, c0, f, h, b);
                tma_load_5d(sp + kOffVt, &mv, &bars[kTmaFull + st], 0, c0, h * VB, f, b);
            };
            issue_loads(0);
#pragma unroll 1
            for (int n = 0; n < NC; ++n) {
                const int st = n & 1;
                const uint64_t dK = umma_smem_desc_sw128(sbase + st * kStageBytes + kOffKt, 16, 1024);
                mbar_wait(&bars[kTmaFull + st], (uint32_t)(n >> 1) & 1u);
                if (n >= 1) mbar_wait(&bars[kKqFree], (uint32_t)(n - 1) & 1u);      // accumulators of chunk n-1 drained
                tc_fence_after_sync();
                umma4_ss(tmem + kColKQ, dK, 2, dK, 2, kIdKK, false);                // [K;Q] K^T
                umma_commit(&bars[kKqFull]);
                if (n + 1 < NC) {                    // refill the other stage: chunk n-1 no longer reads it
                    if (n >= 1) mbar_wait(&bars[kD1Done], (uint32_t)(n - 1) & 1u);
                    issue_loads(n + 1);
                }
                {   // second copy of the K tile for the state update (raw K; its buffer is free once chunk n-2 completed)
                    if (n >= 2) mbar_wait(&bars[kKsideEmpty + st], (uint32_t)((n >> 1) - 1) & 1u);
                    const int f = n / cpf, c0 = (n - f * cpf) << 6;
                    mbar_arrive_expect_tx(&bars[kKpFull + st], 8192u);
                    tma_load_5d(smem + kOffKp + st * 8192, &mk, &bars[kKpFull + st], 0, c0, f, h, b);
                }
            }
        }
        __syncwarp();
    } else {
        // =========================================================================================
        // issuer S: the state-side MMAs
        // =========================================================================================
        if (lane == 0) {
            PT_DECL
#pragma unroll 1
            for (int n = 0; n < NC; ++n) {
                const int st = n & 1;
                const uint32_t aQt = sbase + st * kStageBytes + kOffQt, aVt = aQt + 8192;
                const uint64_t dTp = umma_smem_desc_sw128(sbase + kOffTp + st * 8192, 16, 1024);
                const uint64_t dWt = umma_smem_desc_sw128(sbase + kOffWt + st * 8192, 8192, 1024);
                const uint64_t dKp = umma_smem_desc_sw128(sbase + kOffKp + st * 8192, 8192, 1024);
                const uint64_t dPp = umma_smem_desc_sw128(sbase + kOffPp + st * 8192, 16, 1024);
                const uint64_t dQt = umma_smem_desc_sw128(aQt, 16, 1024);
                mbar_wait(&bars[kTpReady + st], (uint32_t)(n >> 1) & 1u);
                tc_fence_after_sync();
                PT(32, true);   // issuer S: wait K side
                for (int hh = 0; hh < NH; ++hh)      // Vn^T[h] = V^T[h] T'^T   (in order after the MMAs of chunk n-1)
                    umma4_ss(tmem + kColVn + hh * 64, umma_smem_desc_sw128(aVt + hh * 16384, 8192, 1024), 128, dTp, 2, kIdMnA, false);
                mbar_wait(&bars[kKsideFull + st], (uint32_t)(n >> 1) & 1u);
                PT(33, true);   // issuer S: issue U, wait W^T operand
                for (int hh = 0; hh < NH; ++hh) {    // Vn^T[h] -= Sb[h] W^T
                    mbar_wait(&bars[kSbReady + hh], (uint32_t)n & 1u);
                    tc_fence_after_sync();
                    umma4_ts(tmem + kColVn + hh * 64, tmem + kColSb + hh * 32, dWt, 128, kIdMnBneg, true);
                    umma_commit(&bars[kVnFull + hh]);
                }
                PT(34, true);   // issuer S: wait Sb, issue Vn correction
                for (int hh = 0; hh < NH; ++hh) {    // O^T[h] = Sb[h] Q~^T
                    if (n >= 1) mbar_wait(&bars[kOFree + hh], (uint32_t)(n - 1) & 1u);
                    tc_fence_after_sync();
                    umma4_ts(tmem + kColO + hh * 64, tmem + kColSb + hh * 32, dQt, 2, kIdKK, false);
                }
                umma_commit(&bars[kD1Done]);
                PT(35, true);   // issuer S: wait O free, issue inter-chunk readout
                mbar_wait(&bars[kKpFull + st], (uint32_t)(n >> 1) & 1u);
                PT(36, true);   // issuer S: wait K copy
                for (int hh = 0; hh < NH; ++hh) {    // S^T[h] += Vnb[h] K' ;  O^T[h] += Vnb[h] P^T
                    mbar_wait(&bars[kVnbReady + hh], (uint32_t)n & 1u);
                    tc_fence_after_sync();
                    umma4_ts(tmem + kColS + hh * 64, tmem + kColVn + hh * 64, dKp, 128, kIdMnB, true);
                    umma_commit(&bars[kSReady + hh]);
                    umma4_ts(tmem + kColO + hh * 64, tmem + kColVn + hh * 64, dPp, 2, kIdKK, true);
                    umma_commit(&bars[kOFull + hh]);
                }
                umma_commit(&bars[kKsideEmpty + st]);
                PT(37, true);   // issuer S: wait Vnb, issue state update + intra-chunk readout
            }
        }
        __syncwarp();
    }

    // ---- teardown ----
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 16) tmem_dealloc(tmem, kTmemCols);
}

bool mult16(int64_t elems) { return (elems * 2) % 16 == 0; }

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
namespace gdkvm {
#endif

bool chunked_supports(const GdkvmGdrParams& p) {
    if (p.io_dtype != GDKVM_BF16 || p.K != 64 || (p.V != 128 && p.V != 256) || p.T <= 0) return false;
    // TMA: 16-byte aligned bases and strides; the value/readout head stride must equal V so that
    // (head, 64-wide value block) folds into one tensor-map dimension.
    const void* ptrs[4] = {p.q, p.k, p.v, p.o};
    for (const void* x : ptrs) if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return false;
    for (int i = 0; i < 3; ++i)
        if (!mult16(p.q_stride[i]) || !mult16(p.k_stride[i]) || !mult16(p.v_stride[i]) || !mult16(p.o_stride[i])) return false;
    if (p.v_stride[2] != p.V || p.o_stride[2] != p.V) return false;
    if (p.q_stride[1] <= 0 || p.k_stride[1] <= 0 || p.v_stride[1] <= 0 || p.o_stride[1] <= 0) return false;
    return true;
}

int launch_chunked(const GdkvmGdrParams& p, cudaStream_t stream) {
    static std::once_flag once;
    static cudaError_t attr_err = cudaSuccess;
    std::call_once(once, [] {
        attr_err = cudaFuncSetAttribute(gdr_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
    });
    if (attr_err != cudaSuccess) {   // per-device attribute: retry (another device may be current now)
        attr_err = cudaFuncSetAttribute(gdr_chunk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes);
        if (attr_err != cudaSuccess) return (int)attr_err;
    }
    // frame-aligned chunks when frames are whole 64-token chunks (or when asked for); otherwise tile the
    // flat token stream: identical results (token-causal recurrence), no zero-padded rows to process
    const bool flat = p.frame_tokens <= 0 || (p.flags & GDKVM_FLAG_FLAT_CHUNKS) ||
                      (p.frame_tokens % 64 != 0 && !(p.flags & GDKVM_FLAG_FRAME_CHUNKS));
    const int C = flat ? p.T : p.frame_tokens;
    const int F = p.T / C;
    const uint64_t B = p.B, H = p.H, V = p.V;
    CUtensorMap mq, mk, mv, mo;
    // q,k: (dk, token-in-frame, frame, head, clip)
    {
        const uint64_t dims[5] = {64, (uint64_t)C, (uint64_t)F, H, B};
        const uint32_t box[5] = {64, 64, 1, 1, 1};
        const uint64_t sq[4] = {(uint64_t)p.q_stride[1] * 2, (uint64_t)p.q_stride[1] * 2 * C, (uint64_t)p.q_stride[2] * 2, (uint64_t)p.q_stride[0] * 2};
        const uint64_t sk[4] = {(uint64_t)p.k_stride[1] * 2, (uint64_t)p.k_stride[1] * 2 * C, (uint64_t)p.k_stride[2] * 2, (uint64_t)p.k_stride[0] * 2};
        int rc = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.q, dims, sq, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.k, dims, sk, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    // v: (64 values, token-in-frame, head x value-block, frame, clip), whole V per box;
    // o: same geometry, one 128-column half per box (each state warpgroup stores its own half)
    {
        const uint64_t dims[5] = {64, (uint64_t)C, H * (V / 64), (uint64_t)F, B};
        const uint32_t boxv[5] = {64, 64, (uint32_t)(V / 64), 1, 1};
        const uint32_t boxo[5] = {64, 64, 2, 1, 1};
        const uint64_t sv[4] = {(uint64_t)p.v_stride[1] * 2, 128, (uint64_t)p.v_stride[1] * 2 * C, (uint64_t)p.v_stride[0] * 2};
        const uint64_t so[4] = {(uint64_t)p.o_stride[1] * 2, 128, (uint64_t)p.o_stride[1] * 2 * C, (uint64_t)p.o_stride[0] * 2};
        int rc = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.v, dims, sv, boxv, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.o, dims, so, boxo, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    gdr_chunk_kernel<<<p.B * p.H, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, C, F);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace gdkvm
