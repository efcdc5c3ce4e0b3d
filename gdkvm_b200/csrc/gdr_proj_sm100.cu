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

constexpr int kProjThreads = 320;                                               // producer, MMA issuer, eight epilogue warps
constexpr int kMaxWStages = 4;
constexpr int kBN = 256;                                                        // output columns per tile = N of one tcgen05.mma
constexpr uint32_t kTileBytes = 128 * 64 * 2;                                   // 128 rows x 64 bf16 (one swizzle atom wide)
constexpr uint32_t kWTileBytes = kBN * 64 * 2;                                  // a weight k-block: 256 rows x 64 bf16
// feature block (D / 64 tiles of 16 KB: 64 KB at D = 256, at most 128 KB) followed by the ring of 32 KB weight k-blocks: 192 KB
// in all, so D <= 256 leaves four weight stages (128 KB in flight -- a k-block comes from L2, ~2 k cycles away) and D = 512 two.
// N = 256 per MMA because ISSUING an MMA costs ~100 cycles on the one issuer warp: at N = 128 the
// 64-cycle MMAs were issue-bound (0.92 ms).
constexpr uint32_t kOffStaging = 12 * kTileBytes;                               // 8 epilogue warps x 32 rows x 128 B
constexpr uint32_t kOffProjBar = kOffStaging + 8 * 4096;
constexpr uint32_t kProjSmem = kOffProjBar + 512 + 1024;                        // + barriers + alignment slack
static_assert(kProjSmem <= 232448, "exceeds the 227 KB dynamic shared memory limit");

__device__ __forceinline__ float log_sigmoid(float x) { return fminf(x, 0.f) - log1pf(__expf(-fabsf(x))); }

__global__ void __launch_bounds__(kProjThreads, 1)
qkvgb_proj_kernel(const __grid_constant__ CUtensorMap mx, const __grid_constant__ CUtensorMap mw, const GdkvmProjParams p, const int n_tiles,
                  const int64_t m_tiles) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const uint32_t sbase = smem_u32(smem);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffProjBar);
    uint64_t* w_full = bars;                          // [stages] weight k-block landed (tx)
    uint64_t* w_empty = bars + kMaxWStages;           // [stages] its MMAs completed (commit)
    uint64_t* a_full = bars + 2 * kMaxWStages;        // [2] feature block landed (tx)
    uint64_t* a_empty = a_full + 2;                   // [2] every MMA of the row block completed (commit)
    uint64_t* acc_full = a_empty + 2;                 // [2] accumulators of a tile complete (commit)
    uint64_t* acc_empty = acc_full + 2;               // [2] accumulators drained by the four epilogue warps
    uint32_t* s_tmem = reinterpret_cast<uint32_t*>(acc_empty + 2);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int KB = p.D >> 6;
    const uint32_t nbuf = 1u;                         // feature-block buffers (one: the smem goes to the weight ring; the block's
                                                      // reload is exposed once per n_tiles tiles)
    const uint32_t abytes = (uint32_t)KB * kTileBytes;
    const uint32_t kOffW = abytes, kWStages = min((12u - (uint32_t)KB) / 2u, (uint32_t)kMaxWStages);

    if (tid == 0) {
        for (int i = 0; i < kMaxWStages; ++i) { mbar_init(&w_full[i], 1); mbar_init(&w_empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&a_full[i], 1); mbar_init(&a_empty[i], 1); mbar_init(&acc_full[i], 1); mbar_init(&acc_empty[i], 8); }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(s_tmem, 512);
    if (warp == 0 && lane == 0) { tma_prefetch_desc(&mx); tma_prefetch_desc(&mw); }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *s_tmem;

    if (warp == 0) {
        // ---- TMA producer: per row block the feature block (once), then the weight k-blocks of every column tile ----
        uint32_t it = 0, j = 0;
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x, ++j) {
            const uint32_t ab = j % nbuf;
            if (j >= nbuf) mbar_wait_inl(&a_empty[ab], (j / nbuf - 1) & 1u);
            if (elect_one()) {
                mbar_arrive_expect_tx(&a_full[ab], abytes);
                for (int kb = 0; kb < KB; ++kb) tma_load_2d(smem + ab * abytes + (uint32_t)kb * kTileBytes, &mx, &a_full[ab], kb * 64, (int)(tm * 128));
            }
            __syncwarp();
            for (int tn = 0; tn < n_tiles; ++tn) {
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % kWStages;
                    if (it >= kWStages) mbar_wait_inl(&w_empty[s], (it / kWStages - 1) & 1u);
                    if (elect_one()) {
                        mbar_arrive_expect_tx(&w_full[s], kWTileBytes);
                        tma_load_2d(smem + kOffW + s * kWTileBytes, &mw, &w_full[s], kb * 64, tn * kBN);
                    }
                    __syncwarp();
                }
            }
        }
    } else if (warp == 1) {
        // ---- MMA issuer: D[128 x 256] += X_block[:, kb] W_tile[256 x 64]^T, four K = 16 slices per weight k-block ----
        constexpr uint32_t kIdesc = umma_idesc_bf16(128, kBN, false, false);
        uint32_t it = 0, i = 0, j = 0;
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x, ++j) {
            const uint32_t ab = j % nbuf;
            mbar_wait_inl(&a_full[ab], (j / nbuf) & 1u);
            for (int tn = 0; tn < n_tiles; ++tn, ++i) {
                const uint32_t buf = i & 1u;
                if (i >= 2) mbar_wait_inl(&acc_empty[buf], (i / 2 - 1) & 1u);
                tc_fence_after_sync();
                for (int kb = 0; kb < KB; ++kb, ++it) {
                    const uint32_t s = it % kWStages;
                    mbar_wait_inl(&w_full[s], (it / kWStages) & 1u);
                    tc_fence_after_sync();
                    const uint64_t da = umma_smem_desc_sw128(sbase + ab * abytes + (uint32_t)kb * kTileBytes, 16, 1024);
                    const uint64_t db = umma_smem_desc_sw128(sbase + kOffW + s * kWTileBytes, 16, 1024);
                    umma4_ss_w(tmem + buf * kBN, da, da + 2, da + 4, da + 6, db, db + 2, db + 4, db + 6, kIdesc, kb > 0);
                    umma_commit_w(&w_empty[s]);
                }
                umma_commit_w(&acc_full[buf]);
            }
            umma_commit_w(&a_empty[ab]);           // the feature block may be overwritten once every MMA issued so far has completed
        }
    } else {
        // ---- epilogue: one thread = one token row, 64 columns (one head of q / k, a quarter head of v, or the gates) of every tile.
        // Eight warps: two per TMEM lane quarter, one for each 64-column half of the tile -- with a single epilogue warp per
        // scheduler the load -> normalise -> stage -> store chain of a tile (~3.5 k cycles) outlasted its MMAs (~1 k) ----
        const int quarter = warp & 3, gsel = (warp - 2) >> 2;
        uint8_t* stg = smem + kOffStaging + (warp - 2) * 4096;
        const int H = p.H, Nq = H * 64, Nv = H * p.V, Ntot = 2 * Nq + Nv + 2 * H;
        uint32_t i = 0;
        for (int64_t tm = blockIdx.x; tm < m_tiles; tm += gridDim.x)
        for (int tn = 0; tn < n_tiles; ++tn, ++i) {
            const int64_t row0 = tm * 128 + quarter * 32;
            const uint32_t buf = i & 1u;
            const uint32_t taddr = tmem + buf * kBN + ((uint32_t)(quarter * 32) << 16);
            mbar_wait_inl(&acc_full[buf], (i / 2) & 1u);
            tc_fence_after_sync();
#pragma unroll 1
            for (int gi = 2 * gsel; gi < 2 * gsel + 2; ++gi) {
                const int col0 = tn * kBN + gi * 64;
                uint32_t r0[32], r1[32];
                if (col0 < Ntot) {
                    tmem_ld32(taddr + gi * 64, r0);
                    tmem_ld32(taddr + gi * 64 + 32, r1);
                    tmem_wait_ld();
                }
                if (gi == 2 * gsel + 1) {   // this warp's TMEM loads of the tile have completed: hand the accumulator back (8 arrivals)
                    tc_fence_before_sync();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&acc_empty[buf]);
                }
                if (col0 >= Ntot) continue;                            // (warp-uniform)
                if (p.bias != nullptr) {
                    const float* bs = p.bias + col0;
                    const int nb = min(64, Ntot - col0);
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        if (j < nb) r0[j] = __float_as_uint(__uint_as_float(r0[j]) + __ldg(bs + j));
                        if (j + 32 < nb) r1[j] = __float_as_uint(__uint_as_float(r1[j]) + __ldg(bs + 32 + j));
                    }
                }
                if (col0 < 2 * Nq + Nv) {
                    float f = 1.f;
                    __nv_bfloat16* dst;                 // row 0 of this warp's 32 rows, first of the 64 columns
                    int64_t rstride;
                    if (col0 < 2 * Nq) {                // a head of q or k: L2 normalisation over its 64 columns
                        float ss = 0.f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            ss = fmaf(__uint_as_float(r0[j]), __uint_as_float(r0[j]), ss);
                            ss = fmaf(__uint_as_float(r1[j]), __uint_as_float(r1[j]), ss);
                        }
                        f = rsqrtf(ss + p.eps);
                        rstride = Nq;
                        dst = col0 < Nq ? reinterpret_cast<__nv_bfloat16*>(p.q) + row0 * Nq + col0
                                        : reinterpret_cast<__nv_bfloat16*>(p.k) + row0 * Nq + (col0 - Nq);
                    } else {
                        rstride = Nv;
                        dst = reinterpret_cast<__nv_bfloat16*>(p.v) + row0 * Nv + (col0 - 2 * Nq);
                    }
                    // this thread's 128 bytes -> staging row `lane` (16-byte chunk c at slot c ^ (lane & 7): conflict-free both ways)
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
                    __syncwarp();
                    // ... and out again four rows per instruction: eight lanes write one whole 128-byte line
                    const int ch = lane & 7;
#pragma unroll
                    for (int rr = 0; rr < 8; ++rr) {
                        const int row = (lane >> 3) + 4 * rr;
                        const uint4 val = *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
                        if (row0 + row < p.R) *reinterpret_cast<uint4*>(dst + (int64_t)row * rstride + ch * 8) = val;
                    }
                    __syncwarp();
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
        tc_fence_before_sync();
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem, 512);
}

}  // namespace

const char* proj_unsupported_reason(const GdkvmProjParams& p) {
    if (p.K != 64) return "projection: d_k must be 64";
    if (p.R >= (int64_t)1 << 31) return "projection: at most 2^31 - 1 rows per call (TMA coordinates are 32-bit)";
    if (p.H < 2 || p.H > 32 || (p.H & 1)) return "projection: the number of heads must be even, 2..32";
    if (p.V <= 0 || p.V % 64 != 0) return "projection: d_v must be a multiple of 64";
    if (p.D <= 0 || p.D % 64 != 0 || p.D > 512) return "projection: the feature dimension must be a multiple of 64, at most 512";
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
        const uint32_t box[2] = {64, 128}, boxw[2] = {64, (uint32_t)kBN};     // 64 columns = one 128-byte swizzle atom per row
        int rc = make_tmap(&mx, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.x, dx, sx, box, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc == 0) rc = make_tmap(&mw, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, p.w, dw, sw, boxw, CU_TENSOR_MAP_SWIZZLE_128B);
        if (rc != 0) return (int)cudaErrorInvalidValue;
    }
    const int n_tiles = (int)((N + kBN - 1) / kBN);
    const int64_t m_tiles = (p.R + 127) / 128;
    int sms = 148;
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) { (void)cudaGetLastError(); sms = 148; }
    const unsigned grid = (unsigned)std::min<int64_t>(m_tiles, sms);
    qkvgb_proj_kernel<<<grid, kProjThreads, kProjSmem, stream>>>(mx, mw, p, n_tiles, m_tiles);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace gdkvm
