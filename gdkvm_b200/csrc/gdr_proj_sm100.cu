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
// One CTA = one 128 (tokens) x 128 (output columns) tile; 6 warps: TMA producer, MMA issuer, four epilogue warps (one per
// TMEM lane quarter).  A 3-stage ring of 32 KB keeps two CTAs per SM resident (2 x 128 TMEM columns), so one CTA's
// epilogue overlaps the other's main loop.  Bound: HBM writes (772 B per token-head against 2 D bytes read per token).
#include <mutex>

#include "gdr_common.cuh"
#include "sm100_ptx.cuh"
#include "tma_host.h"

namespace gdkvm {
namespace {

using namespace sm100;

constexpr int kProjThreads = 192;
constexpr int kProjStages = 3;
constexpr uint32_t kTileBytes = 128 * 64 * 2;                                   // 128 rows x 64 bf16 (one swizzle atom wide)
constexpr uint32_t kProjSmem = kProjStages * 2 * kTileBytes + 128 + 1024;      // + barriers + alignment slack

__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(__expf(-fabsf(x))); }

__global__ void __launch_bounds__(kProjThreads, 2)
qkvgb_proj_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mw, const GdkvmProjParams p, const int n_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kProjStages * 2 * kTileBytes);
    uint64_t* full = bars;                       // [stages] tiles landed (tx)
    uint64_t* empty = bars + kProjStages;        // [stages] MMAs of the stage completed (commit)
    uint64_t* acc_full = bars + 2 * kProjStages; // accumulators complete (commit)
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * kProjStages + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int tn = blockIdx.x % n_tiles;
    const int64_t tm = blockIdx.x / n_tiles;
    const int KB = p.D >> 6;

    if (tid == 0) {
        for (int i = 0; i < kProjStages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, 128);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mx); tma_prefetch_desc(&mw); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ---- TMA producer ----
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % kProjStages;
            if (kb >= kProjStages) mbar_wait_inl(&empty[s], (uint32_t)(kb / kProjStages - 1) & 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&full[s], 2 * kTileBytes);
                tma_load_2d(smem + (uint32_t)s * 2 * kTileBytes, &mx, &full[s], kb * 64, (int)(tm * 128));
                tma_load_2d(smem + (uint32_t)s * 2 * kTileBytes + kTileBytes, &mw, &full[s], kb * 64, tn * 128);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D[128 x 128] += X_tile[128 x 64] W_tile[128 x 64]^T, four K = 16 slices per stage ----
        constexpr uint32_t kIdesc = umma_idesc_bf16(128, 128, false, false);
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % kProjStages;
            mbar_wait_inl(&full[s], (uint32_t)(kb / kProjStages) & 1u);
            tc_fence_after_sync();
            const uint64_t da = umma_smem_desc_sw128(sbase + (uint32_t)s * 2 * kTileBytes, 16, 1024);
            const uint64_t db = umma_smem_desc_sw128(sbase + (uint32_t)s * 2 * kTileBytes + kTileBytes, 16, 1024);
            umma4_ss_w(tmem, da, da + 2, da + 4, da + 6, db, db + 2, db + 4, db + 6, kIdesc, kb > 0);
            umma_commit_w(&empty[s]);
        }
        umma_commit_w(acc_full);
    } else {
        // ---- epilogue: one thread = one token row, 64 columns (one head of q / k, a quarter head of v, or the gates) at a time ----
        const int quarter = warp & 3;
        const int64_t row = tm * 128 + quarter * 32 + lane;
        const uint32_t taddr = tmem + ((uint32_t)(quarter * 32) << 16);
        const int H = p.H, Nq = H * 64, Nv = H * p.V;
        mbar_wait_inl(acc_full, 0u);
        tc_fence_after_sync();
#pragma unroll 1
        for (int gi = 0; gi < 2; ++gi) {
            const int col0 = tn * 128 + gi * 64;
            if (col0 >= 2 * Nq + Nv + 2 * H) break;                // (warp-uniform)
            uint32_t r0[32], r1[32];
            tmem_ld32(taddr + gi * 64, r0);
            tmem_ld32(taddr + gi * 64 + 32, r1);
            tmem_wait_ld();
            if (p.bias != nullptr) {
                const float* bs = p.bias + col0;
                const int nb = min(64, 2 * Nq + Nv + 2 * H - col0);
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j < nb) r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __ldg(bs + j));
                    if (j + 32 < nb) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __ldg(bs + 32 + j));
                }
            }
            if (row >= p.R) continue;
            if (col0 < 2 * Nq + Nv) {
                float f = 1.f;
                __nv_bfloat16* dst;
                if (col0 < 2 * Nq) {                               // a head of q or k: L2 normalisation over its 64 columns
                    float ss = 0.f;
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        ss = fmaf(__uint_as_float(r0[j]), __uint_as_float(r0[j]), ss);
                        ss = fmaf(__uint_as_float(r1[j]), __uint_as_float(r1[j]), ss);
                    }
                    f = rsqrtf(ss + p.eps);
                    dst = col0 < Nq ? reinterpret_cast<__nv_bfloat16*>(p.q) + row * Nq + col0
                                    : reinterpret_cast<__nv_bfloat16*>(p.k) + row * Nq + (col0 - Nq);
                } else {
                    dst = reinterpret_cast<__nv_bfloat16*>(p.v) + row * Nv + (col0 - 2 * Nq);
                }
                uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    d4[c] = make_uint4(pack_bf16(__uint_as_float(r0[8 * c]) * f, __uint_as_float(r0[8 * c + 1]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 2]) * f, __uint_as_float(r0[8 * c + 3]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 4]) * f, __uint_as_float(r0[8 * c + 5]) * f),
                                       pack_bf16(__uint_as_float(r0[8 * c + 6]) * f, __uint_as_float(r0[8 * c + 7]) * f));
                    d4[4 + c] = make_uint4(pack_bf16(__uint_as_float(r1[8 * c]) * f, __uint_as_float(r1[8 * c + 1]) * f),
                                           pack_bf16(__uint_as_float(r1[8 * c + 2]) * f, __uint_as_float(r1[8 * c + 3]) * f),
                                           pack_bf16(__uint_as_float(r1[8 * c + 4]) * f, __uint_as_float(r1[8 * c + 5]) * f),
                                           pack_bf16(__uint_as_float(r1[8 * c + 6]) * f, __uint_as_float(r1[8 * c + 7]) * f));
                }
            } else {                                               // gate columns: g (H of them), then beta (H)
                float* gd = p.g + row * H;
                float* bd = p.beta + row * H;
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
        tc_fence_before_sync();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 128);
}

}  // namespace

const char* proj_unsupported_reason(const GdkvmProjParams& p) {
    if (p.K != 64) return "projection: d_k must be 64";
    if (p.H < 2 || p.H > 32 || (p.H & 1)) return "projection: the number of heads must be even, 2..32";
    if (p.V <= 0 || p.V % 64 != 0) return "projection: d_v must be a multiple of 64";
    if (p.D <= 0 || p.D % 64 != 0) return "projection: the feature dimension must be a multiple of 64";
    if (p.x_row_stride < p.D || (p.x_row_stride * 2) % 16 != 0) return "projection: feature row stride must be >= D and a multiple of 16 bytes";
    const void* ptrs[5] = {p.x, p.w, p.q, p.k, p.v};
    for (const void* x : ptrs) if ((reinterpret_cast<uintptr_t>(x) & 15u) != 0) return "projection: x, w, q, k, v must be 16-byte aligned";
    return "";
}

int launch_proj(const GdkvmProjParams& p, cudaStream_t stream) {
    static std::mutex mu;
    static bool attr_ok[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (dev < 0 || dev >= 64 || !attr_ok[dev]) {
            e = cudaFuncSetAttribute(qkvgb_proj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kProjSmem);
            if (e != cudaSuccess) return (int)e;
            if (dev >= 0 && dev < 64) attr_ok[dev] = true;
        }
    }
    const int64_t N = (int64_t)p.H * (128 + p.V) + 2 * p.H;
    CUtensorMap mx, mw;
    {
        const uint64_t dx[2] = {(uint64_t)p.D, (uint64_t)p.R}, sx[1] = {(uint64_t)p.x_row_stride * 2};
        const uint64_t dw[2] = {(uint64_t)p.D, (uint64_t)N}, sw[1] = {(uint64_t)p.D * 2};
        const uint32_t box[2] = {64, 128};
        int rc = make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.x, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.w, dw, sw, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    const int n_tiles = (int)((N + 127) / 128);
    const int64_t m_tiles = (p.R + 127) / 128;
    if (m_tiles * n_tiles > 0x7fffffff) return (int)cudaErrorInvalidValue;
    qkvgb_proj_kernel<<<(unsigned)(m_tiles * n_tiles), kProjThreads, kProjSmem, stream>>>(mx, mw, p, n_tiles);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace gdkvm
