// Fused projection prologue of the memory op: the step immediately before gdr_lkva.
//   replaces: "Key-Pixel Feature Fusion fuses the local key feature, the global key feature with the pixel feature"
//   (reference website/src/content/homepage/en.json:20) as far as it PRODUCES the op's operands -- one linear map of the
//   fused per-token feature to q | k | v | gate | beta -- followed by the normalisations the op expects
//   (fla/ops/gated_delta_rule/chunk.py:374 `use_qk_l2norm_in_kernel`; log-space gate <= 0; beta in (0, 1)).
//
// y = x W^T (+ bias) as a tcgen05 GEMM (bf16 operands staged by TMA in 128B-swizzled K-major tiles, fp32 accumulators in
// TMEM), and an epilogue that works on whole rows -- a thread of the 32x32b TMEM load owns one token's 64 columns of a
// head -- so the L2 norm of q and k is a thread-local sum of squares, beta = sigmoid, g = logsigmoid, and the results
// leave in the op's layouts (q, k [R,H,64], v [R,H,V] bf16; g, beta [R,H] fp32).  Against the unfused route (library
// GEMM -> y in HBM -> two normalisation passes -> elementwise gate kernels) this removes one full write + read of y.
//
// Persistent: one CTA per SM takes 128-token row blocks; the block's feature tile (128 x D bf16, up to 128 KB) is loaded ONCE
// and stays in shared memory while the CTA walks all N / 256 column tiles of the block, so each output tile costs one 256 x D
// weight tile from L2 (the 1.6 MB weight stays L2-resident) instead of a feature tile and a weight tile.  10 warps: TMA producer
// (feature block + a ring of 32 KB weight k-blocks, four stages deep at D <= 256), MMA issuer (two 256-column TMEM accumulators, so the MMAs of
// tile i + 1 overlap the epilogue of tile i), eight epilogue warps (two per TMEM lane quarter, two of a tile's four 64-column groups each) that transpose their rows through
// a swizzled shared-memory staging tile so that every global store instruction writes whole 128-byte lines.
// Bound: HBM writes (772 B per token-head against 2 D bytes read per token).
#include <algorithm>
#include <mutex>

#include "gdr_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace gdkvm {
namespace {

using namespace sm100;

constexpr int kMaxWStages = 8;
constexpr uint32_t kTileBytes = 128 * 64 * 2;                                   // 128 rows x 64 bf16 (one swizzle atom wide)
// Shared memory: 12 operand tiles of 16 KB -- the feature block (BM / 128 tiles per 64-wide k-block) followed by the ring of
// weight k-blocks -- then the staging tiles of the eight epilogue warps, then the barriers.
constexpr uint32_t kOperandTiles = 12;
constexpr uint32_t kStgPerWarp = 4096;
constexpr uint32_t kOffStaging = kOperandTiles * kTileBytes;                    // 8 epilogue warps x 32 rows x 128 B
constexpr uint32_t kOffProjBar = kOffStaging + 8 * kStgPerWarp;
constexpr uint32_t kOffBias = kOffProjBar + 256;                                // two buffers of 128 fp32 bias values (256-row tiles)
constexpr uint32_t kProjSmem = kOffBias + 1024 + 1024;                          // + alignment slack
static_assert(kProjSmem <= 232448, "exceeds the 227 KB dynamic shared memory limit");
#ifndef GDKVM_PROJ_ABLATE
#define GDKVM_PROJ_ABLATE 0     // measurement builds only (scripts/ablate_proj.py): 1 no epilogue work, 2 no weight reloads, 4 no global stores, 8 no MMAs
#endif
static_assert(kProjSmem <= 232448, "exceeds the 227 KB dynamic shared memory limit");

// ---- optional phase timers (-DGDKVM_PROJ_TIMERS, scripts/proj_phase_timers.py): cycles of CTA 0's first epilogue warp (slots 0-7) and
// of its MMA issuer (slots 16-19) between consecutive points ----
#ifdef GDKVM_PROJ_TIMERS
__device__ unsigned long long g_proj_cycles[32];
#define PT_DECL long long pt_prev = clock64();
#define PT(slot)                                                                  \
    do {                                                                          \
        if (blockIdx.x == 0 && (warp == 2 || warp == 1) && lane == 0) {           \
            const long long pt_now = clock64();                                   \
            g_proj_cycles[slot] += (unsigned long long)(pt_now - pt_prev);        \
            pt_prev = pt_now;                                                     \
        }                                                                         \
    } while (0)
#else
#define PT_DECL
#define PT(slot) do { } while (0)
#endif

__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(__expf(-fabsf(x))); }

// Two tile shapes.  BM = 128 (any D <= 512): a CTA takes 128-token row blocks, output tiles of 128 x 256, one MMA issuer
// (N = 256 per instruction: issuing an MMA costs ~100 cycles on its warp, so 64-cycle N = 128 MMAs would be issue-bound).
// BM = 256 (D <= 256, the configs[1] geometry): 256-token row blocks, output tiles of 256 x 128 computed as two 128-row halves by
// TWO issuer warps that consume the same 16 KB weight k-block -- per output element half the weight bytes come from L2.  That is
// what bounds the BM = 128 shape: every SM streams 128 KB of weights + 64 KB of output per tile through L2, 148 SMs x 197 KB in
// ~4.4 k cycles = the ~6.3 KB/clk L2 throughput ceiling (B300_MICROARCH.md "LTS throughput cap"), the tensor pipe waiting on
// weight k-blocks 40-50 % of the time (profiles/r4d_proj_timers.log).
template <int BM>
struct ProjShape {
    static constexpr int kHalves = BM / 128;                      // 128-row accumulators per tile = issuer warps
    static constexpr int kBN = BM == 128 ? 256 : 128;             // output columns per tile = N of one tcgen05.mma
    static constexpr int kThreads = 320 + (kHalves - 1) * 32;     // producer, issuer, eight epilogue warps (, second issuer)
    static constexpr uint32_t kWTileBytes = kBN * 64 * 2;         // a weight k-block: kBN rows x 64 bf16
    static constexpr uint32_t kATileBytes = BM * 64 * 2;          // a feature k-block: BM rows x 64 bf16
};

template <int BM>
__global__ void __launch_bounds__(ProjShape<BM>::kThreads, 1)
qkvgb_proj_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mw, const __grid_constant__ CUtensorMap mq,
                  const __grid_constant__ CUtensorMap mk, const __grid_constant__ CUtensorMap mv, const GdkvmProjParams p, const int n_tiles,
                  const int64_t m_tiles) {
    using SH = ProjShape<BM>;
    constexpr int kBN = SH::kBN, kHalves = SH::kHalves;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffProjBar);
    uint64_t* w_full = bars;                          // [stages] weight k-block landed (tx)
    uint64_t* w_empty = bars + kMaxWStages;           // [stages] its MMAs completed (one commit per issuer)
    uint64_t* a_full = bars + 2 * kMaxWStages;        // feature block landed (tx)
    uint64_t* a_empty = a_full + 1;                   // every MMA of the row block completed (one commit per issuer)
    uint64_t* acc_full = a_empty + 1;                 // [2 buffers][2 halves] accumulator of a tile complete (commit)
    uint64_t* acc_empty = acc_full + 4;               // [2][2] accumulator drained by its epilogue warps
    uint64_t* bias_full = acc_empty + 4;              // [2 buffers] the tile's bias values are in shared memory (256-row tiles)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bias_full + 2);
    float* s_bias = reinterpret_cast<float*>(smem + kOffBias);
    // 256-row tiles: the producer warp stages each tile's 128 bias values in shared memory (16 broadcast LDS.128 per group in the
    // epilogue); with 64 scalar or 16 vector global loads per thread and group the bias cost 0.7 / 0.1 ms at configs[1]
    const bool smem_bias = BM == 256 && p.bias != nullptr;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = p.D >> 6;
    // one feature-block buffer: the shared memory goes to the weight ring; the block's reload is exposed once per n_tiles tiles
    const uint32_t abytes = (uint32_t)KB * SH::kATileBytes;
    const uint32_t kOffW = abytes;
    const uint32_t kWStages = min((kOperandTiles * kTileBytes - abytes) / SH::kWTileBytes, BM == 128 ? 4u : (uint32_t)kMaxWStages);

    if (tid == 0) {
        for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], kHalves); }
        mbar_init(a_full, 1);
        mbar_init(a_empty, kHalves);
        for (int i = 0; i < 4; ++i) { mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8 / kHalves); }
        mbar_init(&bias_full[0], 1);
        mbar_init(&bias_full[1], 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, 512);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mx); tma_prefetch_desc(&mw); tma_prefetch_desc(&mq); tma_prefetch_desc(&mk); tma_prefetch_desc(&mv); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ---- TMA producer: per row block the feature block (once), then the weight k-blocks of every column tile ----
        uint32_t it = 0, j = 0, i = 0;
        const int Ntot_p = p.H * (128 + p.V) + 2 * p.H;
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x, ++j) {
            if (j >= 1) mbar_wait_inl(a_empty, (j - 1) & 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(a_full, abytes);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem + (uint32_t)kb * SH::kATileBytes, &mx, a_full, kb * 64, (int)(tm * BM));
            }
            __syncwarp();
            for (int tn = 0; tn < n_tiles; ++tn, ++i) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % kWStages;
                    if (it >= kWStages) mbar_wait_inl(&w_empty[s], (it / kWStages - 1) & 1u);
                    if (elect_one()) {
                        if ((GDKVM_PROJ_ABLATE & 2) && it >= kWStages) mbar_arrive(&w_full[s]);
                        else {
                            mbar_arrive_expect_tx(&w_full[s], SH::kWTileBytes);
                            tma_load_2d(smem + kOffW + s * SH::kWTileBytes, &mw, &w_full[s], kb * 64, tn * kBN);
                        }
                    }
                    __syncwarp();
                }
                if (smem_bias) {
                    // (after the tile's weight loads are in flight, so that this wait never delays them)  bias buffer i & 1 is free once the
                    // epilogue of tile i - 2 has handed its accumulators back: it reads the bias before that
                    const uint32_t buf = i & 1u;
                    if (i >= 2) { mbar_wait_inl(&acc_empty[buf * 2], (i / 2 - 1) & 1u); mbar_wait_inl(&acc_empty[buf * 2 + 1], (i / 2 - 1) & 1u); }
                    float4 bv;
                    const int c = tn * kBN + lane * 4;
                    bv.x = c < Ntot_p ? __ldg(p.bias + c) : 0.f;
                    bv.y = c + 1 < Ntot_p ? __ldg(p.bias + c + 1) : 0.f;
                    bv.z = c + 2 < Ntot_p ? __ldg(p.bias + c + 2) : 0.f;
                    bv.w = c + 3 < Ntot_p ? __ldg(p.bias + c + 3) : 0.f;
                    *reinterpret_cast<float4*>(s_bias + buf * 128 + lane * 4) = bv;
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&bias_full[buf]);
                }
            }
        }
    } else if (warp == 1 || warp == 10) {
        // ---- MMA issuer of one 128-row half: D[128 x kBN] += X_half[:, kb] W_tile[kBN x 64]^T, four K = 16 slices per weight k-block ----
        constexpr uint32_t kIdesc = umma_idesc_bf16(128, kBN, false, false);
        const uint32_t h = warp == 1 ? 0u : 1u;
        uint32_t it = 0, i = 0, j = 0;
        PT_DECL
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x, ++j) {
            mbar_wait_inl(a_full, j & 1u);
            PT(16);
            for (int tn = 0; tn < n_tiles; ++tn, ++i) {
                const uint32_t buf = i & 1u;
                if (i >= 2) mbar_wait_inl(&acc_empty[buf * 2 + h], (i / 2 - 1) & 1u);
                tc_fence_after_sync();
                PT(17);
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % kWStages;
                    mbar_wait_inl(&w_full[s], (it / kWStages) & 1u);
                    tc_fence_after_sync();
                    PT(18);
                    const uint64_t da = umma_smem_desc_sw128(sbase + (uint32_t)kb * SH::kATileBytes + h * kTileBytes, 16, 1024);
                    const uint64_t db = umma_smem_desc_sw128(sbase + kOffW + s * SH::kWTileBytes, 16, 1024);
                    if (!(GDKVM_PROJ_ABLATE & 8)) umma4_ss_w(tmem + buf * 256 + h * 128, da, da + 2, da + 4, da + 6, db, db + 2, db + 4, db + 6, kIdesc, kb > 0);
                    umma_commit_w(&w_empty[s]);
                    PT(19);
                }
                umma_commit_w(&acc_full[buf * 2 + h]);
            }
            umma_commit_w(a_empty);                // the feature block may be overwritten once every MMA issued so far has completed
        }
    } else {
        // ---- epilogue: one thread = one token row, 64 columns (one head of q / k, a quarter head of v, or the gates) at a time.
        // Eight warps, two per TMEM lane quarter: at BM = 128 one for each 128-column half of the tile, at BM = 256 one for each
        // 128-row half -- either way two 64-column groups per warp and tile (with a single epilogue warp per scheduler the
        // load -> normalise -> stage -> store chain of a tile outlasted its MMAs) ----
        const int quarter = warp & 3, gsel = (warp - 2) >> 2;
        const uint32_t h = BM == 128 ? 0u : (uint32_t)gsel;
        const int g_lo = BM == 128 ? 2 * gsel : 0;
        uint8_t* stg = smem + kOffStaging + (warp - 2) * kStgPerWarp;
        const int H = p.H, Nq = H * 64, Nv = H * p.V, Ntot = 2 * Nq + Nv + 2 * H;
        const bool bias_vec = (reinterpret_cast<uintptr_t>(p.bias) & 15u) == 0;
        uint32_t i = 0;
        // The bulk store of a group is ISSUED (by lane 0, ~250 cycles) behind the next group's TMEM load, so that it overlaps the
        // load's latency instead of extending the chain; `pend_*` describe the staged group that has not been sent yet.
        const CUtensorMap* pend_map = nullptr;
        int pend_col = 0, pend_row = 0;
        auto flush_store = [&]() {
            if (pend_map != nullptr) {
                if (lane == 0 && !(GDKVM_PROJ_ABLATE & 4)) {
                    tma_store_2d(pend_map, stg, pend_col, pend_row);
                    tma_store_commit();
                }
                pend_map = nullptr;
            }
        };
        PT_DECL
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x)
        for (int tn = 0; tn < n_tiles; ++tn, ++i) {
            const int64_t row0 = tm * BM + h * 128 + quarter * 32;
            const uint32_t buf = i & 1u;
            const uint32_t taddr = tmem + buf * 256 + h * 128 + ((uint32_t)(quarter * 32) << 16);
            if (!mbar_try_wait(&acc_full[buf * 2 + h], (i / 2) & 1u)) {
                flush_store();                                         // nothing to hide it behind: send the staged group now
                mbar_wait_inl(&acc_full[buf * 2 + h], (i / 2) & 1u);
            }
            tc_fence_after_sync();
            if (smem_bias) mbar_wait_inl(&bias_full[buf], (i / 2) & 1u);
            PT(0);
            if (GDKVM_PROJ_ABLATE & 1) { __syncwarp(); if (lane == 0) mbar_arrive(&acc_empty[buf * 2 + h]); continue; }
#pragma unroll 1
            for (int gi = g_lo; gi < g_lo + 2; ++gi) {
                const int col0 = tn * kBN + gi * 64;
                uint32_t r0[32], r1[32];
                if (col0 < Ntot) {
                    tmem_ld32(taddr + gi * 64, r0);
                    tmem_ld32(taddr + gi * 64 + 32, r1);
                    flush_store();
                    tmem_wait_ld();
                } else {
                    flush_store();
                }
                PT(1);
                if (smem_bias && col0 < Ntot) {
                    const float4* bs4 = reinterpret_cast<const float4*>(s_bias + buf * 128 + (gi - g_lo) * 64);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b0 = bs4[j], b1 = bs4[8 + j];
                        r0[4 * j] = __float_as_uint(__uint_as_float(r0[4 * j]) + b0.x);
                        r0[4 * j + 1] = __float_as_uint(__uint_as_float(r0[4 * j + 1]) + b0.y);
                        r0[4 * j + 2] = __float_as_uint(__uint_as_float(r0[4 * j + 2]) + b0.z);
                        r0[4 * j + 3] = __float_as_uint(__uint_as_float(r0[4 * j + 3]) + b0.w);
                        r1[4 * j] = __float_as_uint(__uint_as_float(r1[4 * j]) + b1.x);
                        r1[4 * j + 1] = __float_as_uint(__uint_as_float(r1[4 * j + 1]) + b1.y);
                        r1[4 * j + 2] = __float_as_uint(__uint_as_float(r1[4 * j + 2]) + b1.z);
                        r1[4 * j + 3] = __float_as_uint(__uint_as_float(r1[4 * j + 3]) + b1.w);
                    }
                }
                if (gi == g_lo + 1) {       // this warp's TMEM loads (and bias reads) of the tile have completed: hand the accumulator back
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf * 2 + h]);
                }
                if (col0 >= Ntot) continue;                            // (warp-uniform)
                if (p.bias != nullptr && !smem_bias) {
                    const float* bs = p.bias + col0;
                    const int nb = min(64, Ntot - col0);
                    if (nb == 64 && bias_vec) {         // every lane reads the same 16 bytes: one broadcast transaction per four columns
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bs) + j), b1 = __ldg(reinterpret_cast<const float4*>(bs) + 8 + j);
                            r0[4 * j] = __float_as_uint(__uint_as_float(r0[4 * j]) + b0.x);
                            r0[4 * j + 1] = __float_as_uint(__uint_as_float(r0[4 * j + 1]) + b0.y);
                            r0[4 * j + 2] = __float_as_uint(__uint_as_float(r0[4 * j + 2]) + b0.z);
                            r0[4 * j + 3] = __float_as_uint(__uint_as_float(r0[4 * j + 3]) + b0.w);
                            r1[4 * j] = __float_as_uint(__uint_as_float(r1[4 * j]) + b1.x);
                            r1[4 * j + 1] = __float_as_uint(__uint_as_float(r1[4 * j + 1]) + b1.y);
                            r1[4 * j + 2] = __float_as_uint(__uint_as_float(r1[4 * j + 2]) + b1.z);
                            r1[4 * j + 3] = __float_as_uint(__uint_as_float(r1[4 * j + 3]) + b1.w);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (j < nb) r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __ldg(bs + j));
                            if (j + 32 < nb) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __ldg(bs + 32 + j));
                        }
                    }
                }
                if (col0 < 2 * Nq + Nv) {
                    float f = 1.f;
                    const CUtensorMap* om;              // destination tensor and the first of the 64 columns in it
                    int ocol;
                    if (col0 < 2 * Nq) {                // a head of q or k: L2 normalisation over its 64 columns
                        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            s0 = fmaf(__uint_as_float(r0[j]), __uint_as_float(r0[j]), s0);
                            s1 = fmaf(__uint_as_float(r1[j]), __uint_as_float(r1[j]), s1);
                            s2 = fmaf(__uint_as_float(r0[j + 1]), __uint_as_float(r0[j + 1]), s2);
                            s3 = fmaf(__uint_as_float(r1[j + 1]), __uint_as_float(r1[j + 1]), s3);
                        }
                        f = rsqrtf((s0 + s1) + (s2 + s3) + p.eps);
                        om = col0 < Nq ? &mq : &mk;
                        ocol = col0 < Nq ? col0 : col0 - Nq;
                    } else {
                        om = &mv;
                        ocol = col0 - 2 * Nq;
                    }
                    PT(2);
                    // the staging tile is free once the previous bulk store has read it
                    if (lane == 0) tma_store_wait_read0();
                    __syncwarp();
                    PT(3);
                    // this thread's 128 bytes -> staging row `lane`, 16-byte chunk c at slot c ^ (lane & 7): the 128B swizzle of the
                    // output tensor maps, and conflict-free
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        *reinterpret_cast<uint4*>(stg + lane * 128 + ((c ^ (lane & 7)) << 4)) =
                            make_uint4(pack_bf16(__uint_as_float(r0[8 * c]) * f, __uint_as_float(r0[8 * c + 1]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 2]) * f, __uint_as_float(r0[8 * c + 3]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 4]) * f, __uint_as_float(r0[8 * c + 5]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 6]) * f, __uint_as_float(r0[8 * c + 7]) * f));
                        *reinterpret_cast<uint4*>(stg + lane * 128 + (((4 + c) ^ (lane & 7)) << 4)) =
                            make_uint4(pack_bf16(__uint_as_float(r1[8 * c]) * f, __uint_as_float(r1[8 * c + 1]) * f),
                                       pack_bf16(__uint_as_float(r1[8 * c + 2]) * f, __uint_as_float(r1[8 * c + 3]) * f),
                                       pack_bf16(__uint_as_float(r1[8 * c + 4]) * f, __uint_as_float(r1[8 * c + 5]) * f),
                                       pack_bf16(__uint_as_float(r1[8 * c + 6]) * f, __uint_as_float(r1[8 * c + 7]) * f));
                    }
                    PT(4);
                    fence_proxy_async_smem();           // generic-proxy writes -> visible to the bulk (async-proxy) store
                    __syncwarp();
                    PT(5);
                    // ... and out as ONE bulk tensor store of 32 rows x 128 bytes (rows past R are clipped by the tensor map),
                    // issued behind the next TMEM load
                    pend_map = om;
                    pend_col = ocol;
                    pend_row = (int)row0;
                    PT(6);
                } else if (row0 + lane < p.R) {                      // gate columns: g (H of them), then beta (H)
                    float* gd = p.g + (row0 + lane) * H;
                    float* bd = p.beta + (row0 + lane) * H;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const float x0 = __uint_as_float(r0[j]), x1 = __uint_as_float(r1[j]);
                        if (j < H) gd[j] = log_sigmoid(x0);
                        else if (j < 2 * H) bd[j - H] = 1.f / (1.f + __expf(-x0));
                        if (j + 32 < H) gd[j + 32] = log_sigmoid(x1);
                        else if (j + 32 < 2 * H) bd[j + 32 - H] = 1.f / (1.f + __expf(-x1));
                    }
                }
            }
        }
        flush_store();
        if (lane == 0) tma_store_wait_read0();     // the staging tile must outlive the last bulk store's read of it
        tc_fence_before_sync();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

#ifdef GDKVM_PROJ_TIMERS
}  // namespace gdkvm
extern "C" int gdkvm_debug_proj_cycles(unsigned long long* out, int n) {
    unsigned long long h[32];
    if (cudaDeviceSynchronize() != cudaSuccess) return -1;
    if (cudaMemcpyFromSymbol(h, gdkvm::g_proj_cycles, sizeof h) != cudaSuccess) return -1;
    for (int i = 0; i < n && i < 32; ++i) out[i] = h[i];
    unsigned long long z[32] = {0};
    cudaMemcpyToSymbol(gdkvm::g_proj_cycles, z, sizeof z);
    return 0;
}
namespace gdkvm {
#endif

const char* proj_unsupported_reason(const GdkvmProjParams& p) {
    if (p.K != 64) return "projection: d_k must be 64";
    if (p.R >= (int64_t)1 << 31) return "projection: at most 2^31 - 1 rows per call (TMA coordinates are 32-bit)";
    if (p.H < 2 || p.H > 32 || (p.H & 1)) return "projection: the number of heads must be even, 2..32";
    if (p.V <= 0 || p.V % 64 != 0) return "projection: d_v must be a multiple of 64";
    if (p.D <= 0 || p.D % 64 != 0 || p.D > 512) return "projection: the feature dimension must be a multiple of 64, at most 512";
    if ((p.flags & ~3u) != 0 || (p.flags & 3u) == 3u) return "projection: unknown flags";
    if ((p.flags & GDKVM_PROJ_FLAG_TILE_ROWS_256) && p.D > 256) return "projection: 256-row tiles need a feature dimension of at most 256";
    if (p.x_row_stride < p.D || (p.x_row_stride * 2) % 16 != 0) return "projection: feature row stride must be >= D and a multiple of 16 bytes";
    const void* ptrs[5] = {p.x, p.w, p.q, p.k, p.v};
    for (const void* x : ptrs) if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return "projection: x, w, q, k, v must be 16-byte aligned";
    return "";
}

template <int BM>
static int launch_proj_shape(const GdkvmProjParams& p, cudaStream_t stream, int dev, int sms) {
    using SH = ProjShape<BM>;
    static std::mutex mu;
    static bool attr_ok[64];
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_ok[dev]) {
            cudaError_t e = cudaFuncSetAttribute(qkvgb_proj_kernel<BM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProjSmem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) attr_ok[dev] = true;
        }
    }
    const int64_t N = (int64_t)p.H * (128 + p.V) + 2 * p.H;
    CUtensorMap mx, mw, mq, mk, mv;
    {
        const uint64_t dx[2] = {(uint64_t)p.D, (uint64_t)p.R}, sx[1] = {(uint64_t)p.x_row_stride * 2};
        const uint64_t dw[2] = {(uint64_t)p.D, (uint64_t)N}, sw[1] = {(uint64_t)p.D * 2};
        const uint32_t box[2] = {64, (uint32_t)BM}, boxw[2] = {64, (uint32_t)SH::kBN};     // 64 columns = one 128-byte swizzle atom per row
        int rc = make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.x, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.w, dw, sw, boxw, CU_TENSOR_MAP_SWIZZLE_128B);
        // outputs: one bulk store = 32 token rows x 64 columns (128 bytes per row, the staging tile of an epilogue warp)
        const uint64_t Nq = (uint64_t)p.H * 64, Nv = (uint64_t)p.H * p.V;
        const uint64_t dq[2] = {Nq, (uint64_t)p.R}, sq[1] = {Nq * 2}, dv[2] = {Nv, (uint64_t)p.R}, sv[1] = {Nv * 2};
        const uint32_t boxo[2] = {64, 32};
        if (rc == 0) rc = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.q, dq, sq, boxo, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.k, dq, sq, boxo, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.v, dv, sv, boxo, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    const int n_tiles = (int)((N + SH::kBN - 1) / SH::kBN);
    const int64_t m_tiles = (p.R + BM - 1) / BM;
    const unsigned grid = (unsigned)std::min<int64_t>(m_tiles, sms);
    qkvgb_proj_kernel<BM><<<grid, SH::kThreads, kProjSmem, stream>>>(mx, mw, mq, mk, mv, p, n_tiles, m_tiles);
    count_launch();
    return (int)cudaGetLastError();
}

int launch_proj(const GdkvmProjParams& p, cudaStream_t stream) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { (void)cudaGetLastError(); sms = 148; }
    // 256-row tiles halve the weight traffic per output element; 128-row tiles spread a small problem over more SMs
    bool big = p.D <= 256 && (p.R + 127) / 128 > sms;
    if (p.flags & GDKVM_PROJ_FLAG_TILE_ROWS_128) big = false;
    if (p.flags & GDKVM_PROJ_FLAG_TILE_ROWS_256) big = true;
    return big ? launch_proj_shape<256>(p, stream, dev, sms) : launch_proj_shape<128>(p, stream, dev, sms);
}

}  // namespace gdkvm
