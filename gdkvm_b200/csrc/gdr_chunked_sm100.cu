// Chunked (WY/UT) GDR/LKVA kernel on the 5th-gen tensor cores of sm_100a.
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
// with the state-independent ("K-side") operands built per chunk:
//     [K;Q] K^T -> gate/mask -> A (fp32), P (bf16);  T = (I + A)^-1 in fp32 on CUDA cores
//     (16x16 forward substitution + two block-merge levels);  T' = T diag(beta);
//     W^T = K~^T T'^T via one more MMA;  K~ = K e^Gamma, K' = K e^(Gamma_last - Gamma),
//     Q~ = scale Q e^Gamma are in-place row scalings of the TMA tiles (swizzle-agnostic).
// Every contraction is a 128 x 64 x 64 tcgen05.mma (bf16 in, fp32 TMEM accumulate); q/k/v tiles
// arrive by TMA (128B swizzle, next chunk prefetched while the current one is computed) and the
// readout leaves through a TMA store.  Layout facts used here were verified on hardware by
// tests/probes/umma_probe.cu.
//
// Math: oracle/gdr_ref.py::gdr_chunk_ref.
#include <mutex>

#include "gdr_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace gdkvm {
namespace {

using namespace sm100;

constexpr int kThreads = 256;
constexpr int kPitchA = 68;   // fp32 pitch of the 64x64 solve matrix (16B-aligned rows, conflict-free v4 stores)
constexpr int kPitchY = 36;

// ---- shared memory map (bytes from a 1024-aligned base) ----
constexpr uint32_t kStageBytes = 49152;          // Kt 8K | Qt 8K | Vt 32K   (Qt must follow Kt: stacked [K;Q] operand)
constexpr uint32_t kOffKt = 0, kOffQt = 8192, kOffVt = 16384;
constexpr uint32_t kOffKp = 2 * kStageBytes;     // K'  (B of the state update, MN-major)
constexpr uint32_t kOffTp = kOffKp + 8192;       // T'  (B of U / W, K-major)
constexpr uint32_t kOffPp = kOffTp + 8192;       // P   (B of the intra-chunk readout, K-major)
constexpr uint32_t kOffWt = kOffPp + 8192;       // W^T (B of the state correction, MN-major)
constexpr uint32_t kOffOst = kOffWt + 8192;      // readout staging for the TMA store, [V/64][64 tok][64] bf16
constexpr uint32_t kOffA = kOffOst + 32768;      // fp32 solve matrix
constexpr uint32_t kOffY = kOffA + 64 * kPitchA * 4;
constexpr uint32_t kOffF = kOffY + 32 * kPitchY * 4;   // floats: g[2][64] beta[2][64] Gam[64] E[64] Fi[64] Kd[64] scal[4]
constexpr uint32_t kOffBar = kOffF + (8 * 64 + 4) * 4;
constexpr uint32_t kSmemBytes = kOffBar + 64 + 1024;   // + alignment slack

// ---- tensor memory map (columns) ----
constexpr uint32_t kColS = 0;      // S^T   [h]: +64h   fp32
constexpr uint32_t kColVn = 128;   // Vn^T  [h]: +64h   fp32; its first 32 columns are re-used for Vnb (bf16)
constexpr uint32_t kColO = 256;    // O^T   [h]: +64h   fp32
constexpr uint32_t kColSb = 384;   // Sb    [h]: +32h   bf16 x2 per column
constexpr uint32_t kColKQ = 448;   // [K;Q]K^T, later W^T (64 columns)
constexpr uint32_t kTmemCols = 512;

__device__ __forceinline__ uint32_t scale_bf16x2(uint32_t w, float s) {
    const float lo = __uint_as_float(w << 16), hi = __uint_as_float(w & 0xffff0000u);
    return pack_bf16(lo * s, hi * s);
}
__device__ __forceinline__ uint4 scale_row8(uint4 v, float s) {
    return make_uint4(scale_bf16x2(v.x, s), scale_bf16x2(v.y, s), scale_bf16x2(v.z, s), scale_bf16x2(v.w, s));
}

// X21 <- -X22 (L21 X11) for NP independent pairs of adjacent N x N diagonal blocks of the unit
// lower-triangular matrix held in sA (in place; sY is scratch).  All 256 threads participate.
template <int N, int NP>
__device__ __forceinline__ void tri_merge(float* sA, float* sY, int tid) {
    constexpr int JW = N * N * NP / kThreads;   // outputs per thread, contiguous along j
    constexpr int TPP = kThreads / NP;          // threads per pair
    constexpr int TPR = N / JW;                 // threads per output row
    const int pair = tid / TPP, t = tid % TPP;
    const int i = t / TPR, j0 = (t % TPR) * JW;
    const int o1 = pair * 2 * N, o2 = o1 + N;
    const float* L21 = sA + o2 * kPitchA + o1;
    const float* X11 = sA + o1 * kPitchA + o1;
    const float* X22 = sA + o2 * kPitchA + o2;
    float* Y = sY + pair * (N * kPitchY);
    float acc[JW];
#pragma unroll
    for (int jj = 0; jj < JW; ++jj) acc[jj] = 0.f;
#pragma unroll 4
    for (int k = 0; k < N; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(L21 + i * kPitchA + k);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) acc[jj] = fmaf(av[kk], X11[(k + kk) * kPitchA + j0 + jj], acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < JW; ++jj) Y[i * kPitchY + j0 + jj] = acc[jj];
    __syncthreads();
#pragma unroll
    for (int jj = 0; jj < JW; ++jj) acc[jj] = 0.f;
#pragma unroll 4
    for (int k = 0; k < N; k += 4) {
        const float4 a = *reinterpret_cast<const float4*>(X22 + i * kPitchA + k);
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
            for (int jj = 0; jj < JW; ++jj) acc[jj] = fmaf(av[kk], Y[(k + kk) * kPitchY + j0 + jj], acc[jj]);
    }
#pragma unroll
    for (int jj = 0; jj < JW; ++jj) sA[(o2 + i) * kPitchA + o1 + j0 + jj] = -acc[jj];
    __syncthreads();
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
    float* sG = reinterpret_cast<float*>(smem + kOffF);   // [2][64]
    float* sBt = sG + 128;                                // [2][64]
    float* sGam = sBt + 128;                              // Gamma_i (inclusive cumsum of g)
    float* sE = sGam + 64;                                // exp(Gamma_i)
    float* sFi = sE + 64;                                 // exp(-Gamma_i)   (fast path only)
    float* sKd = sFi + 64;                                // exp(Gamma_last - Gamma_i)
    float* sScal = sKd + 64;                              // [0] exp(Gamma_last)  [1] fast-path flag
    uint64_t* bar_tma = reinterpret_cast<uint64_t*>(smem + kOffBar);   // [2]
    uint64_t* bar_mma = bar_tma + 2;
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bar_mma + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int wq = warp & 3, wh = warp >> 2;     // TMEM lane quadrant, value half
    const int chain = blockIdx.x, b = chain / p.H, h = chain % p.H;
    const int V = p.V, NH = V >> 7, VB = V >> 6;
    const int cpf = (C + 63) >> 6, NC = F * cpf;
    const bool state_warp = wh < NH;
    const float scale = p.scale;
    const int64_t g_off = (int64_t)b * p.g_stride[0] + (int64_t)h * p.g_stride[2];
    const int64_t bt_off = (int64_t)b * p.beta_stride[0] + (int64_t)h * p.beta_stride[2];
    const uint32_t stage_tx = 16384u + (uint32_t)VB * 8192u;

    if (tid == 0) {
        mbar_init(&bar_tma[0], 1); mbar_init(&bar_tma[1], 1); mbar_init(bar_mma, 1);
        fence_mbar_init();
        tma_prefetch_desc(&mq); tma_prefetch_desc(&mk); tma_prefetch_desc(&mv); tma_prefetch_desc(&mo);
    }
    if (warp == 0) tmem_alloc(s_tmem, kTmemCols);
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;
    const uint32_t lane_addr = tmem + ((uint32_t)(wq * 32) << 16);
    const int vcol = wh * 128 + wq * 32 + lane;   // value column this thread owns on the state side

    auto issue_loads = [&](int n, int st) {   // tid 0: TMA q,k,v tiles of chunk n into stage st
        const int f = n / cpf, c0 = (n - f * cpf) << 6;
        uint8_t* sp = smem + st * kStageBytes;
        mbar_arrive_expect_tx(&bar_tma[st], stage_tx);
        tma_load_5d(sp + kOffKt, &mk, &bar_tma[st], 0, c0, f, h, b);
        tma_load_5d(sp + kOffQt, &mq, &bar_tma[st], 0, c0, f, h, b);
        tma_load_5d(sp + kOffVt, &mv, &bar_tma[st], 0, c0, h * VB, f, b);
    };
    auto load_gates = [&](int n, float& gv, float& bv) {   // tid < 64: g, beta of row tid of chunk n
        const int f = n / cpf, c = ((n - f * cpf) << 6) + tid;
        gv = 0.f; bv = 0.f;                                  // pad rows: exact no-ops
        if (c < C) {
            const int64_t t = (int64_t)f * C + c;
            gv = load_gate(p.g, g_off + t * p.g_stride[1], p.gate_dtype);
            bv = load_gate(p.beta, bt_off + t * p.beta_stride[1], p.gate_dtype);
        }
    };

    // ---- prologue: first tiles in flight, initial state into TMEM ----
    if (tid == 0) issue_loads(0, 0);
    if (tid < 64) { float gv, bv; load_gates(0, gv, bv); sG[tid] = gv; sBt[tid] = bv; }
    if (state_warp) {
        uint32_t r[32];
        const float* s0 = p.initial_state ? p.initial_state + (int64_t)chain * 64 * V + vcol : nullptr;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = s0 ? __float_as_uint(__ldg(s0 + (int64_t)(half * 32 + j) * V)) : 0u;
            tmem_st32(lane_addr + kColS + wh * 64 + half * 32, r);
        }
        tmem_wait_st();
    }
    tc_fence_before_sync();
    __syncthreads();

    uint32_t mma_phase = 0;
    constexpr uint32_t kIdKK = umma_idesc_bf16(128, 64, false, false);
    constexpr uint32_t kIdMnA = umma_idesc_bf16(128, 64, true, false);          // A = tile^T (MN-major), B K-major
    constexpr uint32_t kIdMnB = umma_idesc_bf16(128, 64, false, true);          // TS, B MN-major
    constexpr uint32_t kIdMnBneg = umma_idesc_bf16(128, 64, false, true, true); // TS, -A, B MN-major

    for (int n = 0; n < NC; ++n) {
        const int st = n & 1;
        uint8_t* sp = smem + st * kStageBytes;
        const uint32_t aKt = sbase + st * kStageBytes + kOffKt, aQt = aKt + 8192, aVt = aKt + 16384;
        const float* gS = sG + st * 64;
        const float* btS = sBt + st * 64;
        const int f = n / cpf, c0 = (n - f * cpf) << 6;

        // (0) prefetch chunk n+1 (its stage was released at the end of chunk n-1)
        float g_next = 0.f, b_next = 0.f;
        if (n + 1 < NC) {
            if (tid == 0) issue_loads(n + 1, st ^ 1);
            if (tid < 64) load_gates(n + 1, g_next, b_next);
        }

        // (1) tiles of chunk n have landed
        mbar_wait(&bar_tma[st], (uint32_t)(n >> 1) & 1u);

        // (2) [K;Q] K^T  ->  TMEM KQ
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t ad = umma_smem_desc_sw128(aKt + k * 32, 16, 1024);
                umma_ss(tmem + kColKQ, ad, ad, kIdKK, k > 0);
            }
            umma_commit(bar_mma);
        }
        // gate scan (warp 0): Gamma = cumsum(g) and the per-row decay factors
        if (warp == 0) {
            const float g0 = gS[2 * lane], g1 = gS[2 * lane + 1];
            float s = g0 + g1;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const float t = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += t;
            }
            const float G1 = s, G0 = s - g1;
            const float Gl = __shfl_sync(0xffffffffu, s, 31);
            sGam[2 * lane] = G0; sGam[2 * lane + 1] = G1;
            sE[2 * lane] = __expf(G0); sE[2 * lane + 1] = __expf(G1);
            sFi[2 * lane] = __expf(-G0); sFi[2 * lane + 1] = __expf(-G1);
            sKd[2 * lane] = __expf(Gl - G0); sKd[2 * lane + 1] = __expf(Gl - G1);
            if (lane == 0) { sScal[0] = __expf(Gl); sScal[1] = Gl > -60.f ? 1.f : 0.f; }
        }
        __syncthreads();
        const bool fast = sScal[1] != 0.f;

        // (3) KQ -> gated A (fp32, solve matrix) and P (bf16 operand); row scalings of the tiles
        mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
        tc_fence_after_sync();
        {
            uint32_t r[32];
            tmem_ld32(lane_addr + kColKQ + wh * 32, r);
            tmem_wait_ld();
            if (wq < 2) {          // rows of K K^T
                const int i = wq * 32 + lane;
                const float Gi = sGam[i], bi = btS[i], bie = bi * sE[i];
#pragma unroll
                for (int j4 = 0; j4 < 8; ++j4) {
                    float o[4];
#pragma unroll
                    for (int jj = 0; jj < 4; ++jj) {
                        const int j = wh * 32 + j4 * 4 + jj;
                        const float w = fast ? bie * sFi[j] : bi * __expf(Gi - sGam[j]);
                        o[jj] = j < i ? __uint_as_float(r[j4 * 4 + jj]) * w : 0.f;
                    }
                    *reinterpret_cast<float4*>(sA + i * kPitchA + wh * 32 + j4 * 4) = make_float4(o[0], o[1], o[2], o[3]);
                }
            } else {               // rows of Q K^T
                const int i = (wq - 2) * 32 + lane;
                const float Gi = sGam[i], sce = scale * sE[i];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float o[8];
#pragma unroll
                    for (int jj = 0; jj < 8; ++jj) {
                        const int j = wh * 32 + c * 8 + jj;
                        const float w = fast ? sce * sFi[j] : scale * __expf(Gi - sGam[j]);
                        o[jj] = j <= i ? __uint_as_float(r[c * 8 + jj]) * w : 0.f;
                    }
                    *reinterpret_cast<uint4*>(smem + kOffPp + sw128_offset(i, wh * 4 + c)) =
                        make_uint4(pack_bf16(o[0], o[1]), pack_bf16(o[2], o[3]), pack_bf16(o[4], o[5]), pack_bf16(o[6], o[7]));
                }
            }
        }
#pragma unroll
        for (int idx = tid; idx < 512; idx += kThreads) {      // (row, 16B chunk): swizzle keeps rows intact
            const int row = idx >> 3, off = idx << 4;
            const float e = sE[row];
            uint4* pk = reinterpret_cast<uint4*>(sp + kOffKt + off);
            uint4* pq = reinterpret_cast<uint4*>(sp + kOffQt + off);
            const uint4 kv = *pk;
            *reinterpret_cast<uint4*>(smem + kOffKp + off) = scale_row8(kv, sKd[row]);   // K'
            *pk = scale_row8(kv, e);                                                      // K~
            *pq = scale_row8(*pq, scale * e);                                             // Q~
        }
        tc_fence_before_sync();
        __syncthreads();

        // (4) T = (I + A)^-1 : 16x16 forward substitution, then two block-merge levels
        if (tid < 64) {
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
        }
        __syncthreads();
        tri_merge<16, 2>(sA, sY, tid);
        tri_merge<32, 1>(sA, sY, tid);
        {   // T' = T diag(beta) -> bf16, K-major swizzled rows
            const int i = tid >> 2, cb = (tid & 3) * 2;
#pragma unroll
            for (int c = cb; c < cb + 2; ++c) {
                const float4 x0 = *reinterpret_cast<const float4*>(sA + i * kPitchA + c * 8);
                const float4 x1 = *reinterpret_cast<const float4*>(sA + i * kPitchA + c * 8 + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(btS + c * 8);
                const float4 b1 = *reinterpret_cast<const float4*>(btS + c * 8 + 4);
                *reinterpret_cast<uint4*>(smem + kOffTp + sw128_offset(i, c)) =
                    make_uint4(pack_bf16(x0.x * b0.x, x0.y * b0.y), pack_bf16(x0.z * b0.z, x0.w * b0.w),
                               pack_bf16(x1.x * b1.x, x1.y * b1.y), pack_bf16(x1.z * b1.z, x1.w * b1.w));
            }
        }
        fence_proxy_async_smem();
        __syncthreads();

        // (5) W^T = K~^T T'^T -> TMEM KQ region ;  Vn^T[h] = V^T[h] T'^T
        if (tid == 0) {
            tc_fence_after_sync();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint64_t bd = umma_smem_desc_sw128(sbase + kOffTp + k * 32, 16, 1024);
                umma_ss(tmem + kColKQ, umma_smem_desc_sw128(aKt + k * 2048, 8192, 1024), bd, kIdMnA, k > 0);
            }
            for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t bd = umma_smem_desc_sw128(sbase + kOffTp + k * 32, 16, 1024);
                    umma_ss(tmem + kColVn + hh * 64, umma_smem_desc_sw128(aVt + hh * 16384 + k * 2048, 8192, 1024), bd, kIdMnA, k > 0);
                }
            }
            umma_commit(bar_mma);
        }
        //     meanwhile: Sb = bf16(S^T) (operand copy), S^T <- gamma S^T (decay before the accumulate)
        if (state_warp) {
            const float gam = sScal[0];
            uint32_t r[32], pk[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(lane_addr + kColS + wh * 64 + half * 32, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[half * 16 + j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * gam);
                tmem_st32(lane_addr + kColS + wh * 64 + half * 32, r);
            }
            tmem_st32(lane_addr + kColSb + wh * 32, pk);
            tmem_wait_st();
        }
        tc_fence_before_sync();

        // (6) W^T accumulators -> bf16 MN-major operand rows (row = key dim d, contiguous over tokens)
        mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
        tc_fence_after_sync();
        if (wq < 2) {
            uint32_t r[32];
            const int d = wq * 32 + lane;
            tmem_ld32(lane_addr + kColKQ + wh * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int c = 0; c < 4; ++c)
                *reinterpret_cast<uint4*>(smem + kOffWt + sw128_offset(d, wh * 4 + c)) =
                    make_uint4(pack_bf16(__uint_as_float(r[c * 8 + 0]), __uint_as_float(r[c * 8 + 1])),
                               pack_bf16(__uint_as_float(r[c * 8 + 2]), __uint_as_float(r[c * 8 + 3])),
                               pack_bf16(__uint_as_float(r[c * 8 + 4]), __uint_as_float(r[c * 8 + 5])),
                               pack_bf16(__uint_as_float(r[c * 8 + 6]), __uint_as_float(r[c * 8 + 7])));
        }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();

        // (7) Vn^T[h] -= Sb[h] W^T ;  O^T[h] = Sb[h] Q~^T
        if (tid == 0) {
            tc_fence_after_sync();
            for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kColVn + hh * 64, tmem + kColSb + hh * 32 + k * 8,
                            umma_smem_desc_sw128(sbase + kOffWt + k * 2048, 8192, 1024), kIdMnBneg, true);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kColO + hh * 64, tmem + kColSb + hh * 32 + k * 8,
                            umma_smem_desc_sw128(aQt + k * 32, 16, 1024), kIdKK, k > 0);
            }
            umma_commit(bar_mma);
            tma_store_wait_read0();     // previous chunk's readout has left the staging buffer
        }

        // (8) Vnb = bf16(Vn^T) written over the first half of Vn (TMEM A-operand)
        mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
        tc_fence_after_sync();
        if (state_warp) {
            uint32_t r[32], pk[32];
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(lane_addr + kColVn + wh * 64 + half * 32, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 16; ++j) pk[half * 16 + j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
            }
            tmem_st32(lane_addr + kColVn + wh * 64, pk);
            tmem_wait_st();
        }
        tc_fence_before_sync();
        __syncthreads();

        // (9) S^T[h] += Vnb[h] K' ;  O^T[h] += Vnb[h] P^T
        if (tid == 0) {
            tc_fence_after_sync();
            for (int hh = 0; hh < NH; ++hh) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kColS + hh * 64, tmem + kColVn + hh * 64 + k * 8,
                            umma_smem_desc_sw128(sbase + kOffKp + k * 2048, 8192, 1024), kIdMnB, true);
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    umma_ts(tmem + kColO + hh * 64, tmem + kColVn + hh * 64 + k * 8,
                            umma_smem_desc_sw128(sbase + kOffPp + k * 32, 16, 1024), kIdKK, true);
            }
            umma_commit(bar_mma);
        }

        // (10) readout: O^T -> bf16 -> staging [V/64][tok][64] -> TMA store (rows past the frame are clipped)
        mbar_wait(bar_mma, mma_phase); mma_phase ^= 1;
        tc_fence_after_sync();
        if (state_warp) {
            uint32_t r[32];
            __nv_bfloat16* ost = reinterpret_cast<__nv_bfloat16*>(smem + kOffOst) + (vcol >> 6) * 4096 + (vcol & 63);
#pragma unroll
            for (int half = 0; half < 2; ++half) {
                tmem_ld32(lane_addr + kColO + wh * 64 + half * 32, r);
                tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) ost[(half * 32 + j) * 64] = __float2bfloat16_rn(__uint_as_float(r[j]));
            }
        }
        if (n + 1 < NC && tid < 64) { sG[(st ^ 1) * 64 + tid] = g_next; sBt[(st ^ 1) * 64 + tid] = b_next; }
        fence_proxy_async_smem();
        tc_fence_before_sync();
        __syncthreads();
        if (tid == 0) {
            tma_store_5d(&mo, smem + kOffOst, 0, c0, h * VB, f, b);
            tma_store_commit();
        }
    }

    // ---- epilogue: final state, drain the last store, release TMEM ----
    tc_fence_after_sync();
    if (state_warp && p.final_state != nullptr) {
        uint32_t r[32];
        float* sT = p.final_state + (int64_t)chain * 64 * V + vcol;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            tmem_ld32(lane_addr + kColS + wh * 64 + half * 32, r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) sT[(int64_t)(half * 32 + j) * V] = __uint_as_float(r[j]);
        }
    }
    if (tid == 0) tma_store_wait_all0();
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

bool mult16(int64_t elems) { return (elems * 2) % 16 == 0; }

}  // namespace

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
    const bool flat = p.frame_tokens <= 0 || (p.flags & GDKVM_FLAG_FLAT_CHUNKS);
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
    // v,o: (64 values, token-in-frame, head x value-block, frame, clip)
    {
        const uint64_t dims[5] = {64, (uint64_t)C, H * (V / 64), (uint64_t)F, B};
        const uint32_t box[5] = {64, 64, (uint32_t)(V / 64), 1, 1};
        const uint64_t sv[4] = {(uint64_t)p.v_stride[1] * 2, 128, (uint64_t)p.v_stride[1] * 2 * C, (uint64_t)p.v_stride[0] * 2};
        const uint64_t so[4] = {(uint64_t)p.o_stride[1] * 2, 128, (uint64_t)p.o_stride[1] * 2 * C, (uint64_t)p.o_stride[0] * 2};
        int rc = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.v, dims, sv, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, p.o, dims, so, box, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    gdr_chunk_kernel<<<p.B * p.H, kThreads, kSmemBytes, stream>>>(mq, mk, mv, mo, p, C, F);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace gdkvm
