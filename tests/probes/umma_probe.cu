// Hardware probes for the tcgen05/TMA layout assumptions of gdr_chunked_sm100.cu.
// Each probe is one tiny single-CTA kernel checked against a host fp32 computation; the program
// prints one PASS/FAIL line per hypothesis.  Build + run: scripts/run_probe.sh (under gpurun).
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>

#include "../../gdkvm_b200/csrc/sm100_ptx.cuh"
#include "../../gdkvm_b200/csrc/tma_host.h"

using namespace sm100;
typedef __nv_bfloat16 bf16;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

static float bf(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// ------------------------------------------------------------------------------------------------
// Probe A: SS MMA, K-major A[128x64] and K-major B[64x64] (both TMA SWIZZLE_128B), D = A B^T.
// Checks: TMA 2D load + mbarrier, smem descriptor (SBO=1024, +32B per K=16 step), idesc, commit,
//         tcgen05.ld 32x32b mapping (thread <-> lane/row, register <-> column).
__global__ void __launch_bounds__(128) probe_a(const __grid_constant__ CUtensorMap ma, const __grid_constant__ CUtensorMap mb,
                                               float* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sA = base;            // 128 rows x 128 B = 16 KB
    uint8_t* sB = base + 16384;    // 64 rows x 128 B = 8 KB
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_tma, 16384 + 8192);
        tma_load_2d(sA, &ma, &bar_tma, 0, 0);
        tma_load_2d(sB, &mb, &bar_tma, 0, 0);
        mbar_wait(&bar_tma, 0);
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_bf16(128, 64, false, false);
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = umma_smem_desc_sw128(smem_u32(sA) + k * 32, 16, 1024);
            uint64_t bd = umma_smem_desc_sw128(smem_u32(sB) + k * 32, 16, 1024);
            umma_ss(tmem, ad, bd, idesc, k > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();
    uint32_t r[32];
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// Probe B: MN-major A = Vt (V tile [64 tok x 128 v] via one 3D TMA box {64 vi, 64 tok, 2 vo}),
//          K-major B = Tm[64 n x 64 k] written by threads with the manual 128B swizzle.
//          D[v][n] = sum_tok V[tok][v] * Tm[n][tok].   variant: 0 -> LBO=8192,SBO=1024 ; 1 -> swapped
__global__ void __launch_bounds__(128) probe_b(const __grid_constant__ CUtensorMap mv, const float* tm, float* out, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sV = base;            // [2][64][128 B] = 16 KB
    uint8_t* sT = base + 16384;    // 64 rows x 128 B
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_tma, 16384);
        tma_load_3d(sV, &mv, &bar_tma, 0, 0, 0);
    }
    // Tm -> bf16, swizzled K-major rows: thread handles (row = tid/2, 4 chunks)
    {
        const int row = tid >> 1, c0 = (tid & 1) * 4;
        for (int c = c0; c < c0 + 4; ++c) {
            uint32_t w[4];
            for (int j = 0; j < 4; ++j) w[j] = pack_bf16(tm[row * 64 + c * 8 + 2 * j], tm[row * 64 + c * 8 + 2 * j + 1]);
            *reinterpret_cast<uint4*>(sT + sw128_offset(row, c)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bar_tma, 0);
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_bf16(128, 64, true, false);
        const uint32_t lbo = variant == 0 ? 8192 : 1024, sbo = variant == 0 ? 1024 : 8192;
        for (int k = 0; k < 4; ++k) {
            uint64_t ad = umma_smem_desc_sw128(smem_u32(sV) + k * 2048, lbo, sbo);   // 16 tokens = 2 atoms of 1024 B
            uint64_t bd = umma_smem_desc_sw128(smem_u32(sT) + k * 32, 16, 1024);
            umma_ss(tmem, ad, bd, idesc, k > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();
    uint32_t r[32];
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + half * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

// ------------------------------------------------------------------------------------------------
// Probe C: TS MMA.  A = X[128 x 64] packed to bf16 in TMEM by tcgen05.st (variant bit0: 0 = even
//          element in the low half-word, 1 = odd element low), D preloaded with fp32 S0 by
//          tcgen05.st and accumulated into.  B = Kp[64 tok(k) x 64 dk(n)] TMA tile used MN-major.
//          variant bit1: a_negate.  D[v][d] = S0[v][d] +/- sum_tok X[v][tok] * Kp[tok][d]
__global__ void __launch_bounds__(128) probe_c(const __grid_constant__ CUtensorMap mk, const float* x, const float* s0,
                                               float* out, int variant) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sK = base;            // 64 rows(tok) x 128 B
    __shared__ uint64_t bar_tma, bar_mma;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(&bar_tma, 1); mbar_init(&bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base_s, 128);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_tma, 8192);
        tma_load_2d(sK, &mk, &bar_tma, 0, 0);
    }
    uint32_t r[32];
    // D (cols 0..63) <- S0 row tid
    for (int half = 0; half < 2; ++half) {
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(s0[tid * 64 + half * 32 + j]);
        tmem_st32(tmem + lane_off + half * 32, r);
    }
    // A (cols 64..95) <- bf16 pack of X row tid
    for (int j = 0; j < 32; ++j) {
        const float e = x[tid * 64 + 2 * j], o = x[tid * 64 + 2 * j + 1];
        r[j] = (variant & 1) ? pack_bf16(o, e) : pack_bf16(e, o);
    }
    tmem_st32(tmem + lane_off + 64, r);
    tmem_wait_st();
    tc_fence_before_sync();
    __syncthreads();
    if (tid == 0) {
        mbar_wait(&bar_tma, 0);
        tc_fence_after_sync();
        const uint32_t idesc = umma_idesc_bf16(128, 64, false, true, (variant & 2) != 0, false);
        for (int k = 0; k < 4; ++k) {
            // B MN-major: atom = 8 tok x 64 dk (1024 B); 16 tokens per MMA = 2 atoms -> +2048 B per step
            uint64_t bd = umma_smem_desc_sw128(smem_u32(sK) + k * 2048, 8192, 1024);
            umma_ts(tmem, tmem + 64 + k * 8, bd, idesc, true);   // 16 bf16 of A = 8 TMEM columns
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0);
    tc_fence_after_sync();
    for (int half = 0; half < 2; ++half) {
        tmem_ld32(tmem + lane_off + half * 32, r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out[tid * 64 + half * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 128);
}

// ------------------------------------------------------------------------------------------------
// Probe D: rank-5 TMA.  Load a q-like tile {64 dk, 64 tok-in-frame, 1, 1, 1} of frame f / head h
//          from [B,T=F*C,H,64] with C=49 (rows >= 49 must arrive as zeros), dump it de-swizzled;
//          then store a [4 vo][64 tok][64 vi] staging tile through an o-like map (no swizzle) and
//          check only the 49 valid rows are written.
__global__ void __launch_bounds__(128) probe_d(const __grid_constant__ CUtensorMap mq, const __grid_constant__ CUtensorMap mo,
                                               float* out_tile, int f, int h, int b) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* base = (uint8_t*)(((uintptr_t)smem + 1023) & ~(uintptr_t)1023);
    uint8_t* sQ = base;             // 8 KB
    bf16* sO = (bf16*)(base + 8192);   // [4][64][64] bf16 = 32 KB
    __shared__ uint64_t bar_tma;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar_tma, 1); fence_mbar_init(); }
    __syncthreads();
    if (tid == 0) {
        mbar_arrive_expect_tx(&bar_tma, 8192);
        tma_load_5d(sQ, &mq, &bar_tma, 0, 0, f, h, b);
    }
    mbar_wait(&bar_tma, 0);
    for (int i = tid; i < 64 * 64; i += 128) {
        const int row = i / 64, col = i % 64;
        const bf16* p = (const bf16*)(sQ + sw128_offset(row, col / 8)) + (col % 8);
        out_tile[i] = __bfloat162float(*p);
    }
    for (int i = tid; i < 4 * 64 * 64; i += 128) {
        const int vo = i / 4096, tok = (i / 64) % 64, vi = i % 64;
        sO[i] = __float2bfloat16_rn((float)(tok + 1) + 0.001f * (vo * 64 + vi));   // tok+1 in [1,64]
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
        tma_store_5d(&mo, sO, 0, 0, h * 4, f, b);
        tma_store_commit();
        tma_store_wait_all0();
    }
}

// ------------------------------------------------------------------------------------------------
static double max_err(const std::vector<float>& a, const std::vector<float>& b, double* maxref) {
    double e = 0, m = 0;
    for (size_t i = 0; i < a.size(); ++i) { e = fmax(e, fabs((double)a[i] - b[i])); m = fmax(m, fabs((double)b[i])); }
    *maxref = m;
    return e;
}
static void report(const char* name, double err, double maxref) {
    printf("%-44s %s  max_abs_err %.3e (max |ref| %.3e)\n", name, err <= 2e-2 * maxref + 1e-6 ? "PASS" : "FAIL", err, maxref);
}

int main() {
    srand(1);
    auto rnd = []() { return (float)rand() / RAND_MAX * 2.f - 1.f; };
    const int smem_bytes = 64 * 1024;
    CK(cudaFuncSetAttribute(probe_a, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CK(cudaFuncSetAttribute(probe_b, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CK(cudaFuncSetAttribute(probe_c, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    CK(cudaFuncSetAttribute(probe_d, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    float* d_out; CK(cudaMalloc(&d_out, 128 * 64 * 4));
    std::vector<float> out(128 * 64), ref(128 * 64);
    double mr;

    {   // ---- A
        std::vector<bf16> A(128 * 64), B(64 * 64);
        std::vector<float> Af(128 * 64), Bf(64 * 64);
        for (int i = 0; i < 128 * 64; ++i) { Af[i] = bf(rnd()); A[i] = __float2bfloat16_rn(Af[i]); }
        for (int i = 0; i < 64 * 64; ++i) { Bf[i] = bf(rnd()); B[i] = __float2bfloat16_rn(Bf[i]); }
        for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) { float s = 0; for (int k = 0; k < 64; ++k) s += Af[m * 64 + k] * Bf[n * 64 + k]; ref[m * 64 + n] = s; }
        bf16 *dA, *dB; CK(cudaMalloc(&dA, A.size() * 2)); CK(cudaMalloc(&dB, B.size() * 2));
        CK(cudaMemcpy(dA, A.data(), A.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, B.data(), B.size() * 2, cudaMemcpyHostToDevice));
        CUtensorMap ma, mb;
        uint64_t da[2] = {64, 128}, sa[1] = {128}; uint32_t ba[2] = {64, 128};
        uint64_t db[2] = {64, 64}; uint32_t bb[2] = {64, 64};
        int r1 = make_tmap(&ma, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, da, sa, ba, CU_TENSOR_MAP_SWIZZLE_128B);
        int r2 = make_tmap(&mb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, db, sa, bb, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r1 || r2) { printf("tensor map encode failed %d %d\n", r1, r2); return 2; }
        CK(cudaMemset(d_out, 0, 128 * 64 * 4));
        probe_a<<<1, 128, smem_bytes>>>(ma, mb, d_out);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
        double e = max_err(out, ref, &mr); report("A: SS K-major A,B (TMA sw128) + ld 32x32b", e, mr);
    }
    {   // ---- B
        std::vector<bf16> V(64 * 128); std::vector<float> Vf(64 * 128), Tm(64 * 64);
        for (int i = 0; i < 64 * 128; ++i) { Vf[i] = bf(rnd()); V[i] = __float2bfloat16_rn(Vf[i]); }
        for (int i = 0; i < 64 * 64; ++i) Tm[i] = bf(rnd());
        for (int v = 0; v < 128; ++v) for (int n = 0; n < 64; ++n) { float s = 0; for (int t = 0; t < 64; ++t) s += Vf[t * 128 + v] * Tm[n * 64 + t]; ref[v * 64 + n] = s; }
        bf16* dV; float* dT; CK(cudaMalloc(&dV, V.size() * 2)); CK(cudaMalloc(&dT, Tm.size() * 4));
        CK(cudaMemcpy(dV, V.data(), V.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dT, Tm.data(), Tm.size() * 4, cudaMemcpyHostToDevice));
        CUtensorMap mv;
        uint64_t dv[3] = {64, 64, 2}, sv[2] = {256, 128}; uint32_t bv[3] = {64, 64, 2};   // (vi, tok, vo)
        int r1 = make_tmap(&mv, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, dV, dv, sv, bv, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r1) { printf("tensor map encode failed %d\n", r1); return 2; }
        for (int variant = 0; variant < 2; ++variant) {
            CK(cudaMemset(d_out, 0, 128 * 64 * 4));
            probe_b<<<1, 128, smem_bytes>>>(mv, dT, d_out, variant);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
            double e = max_err(out, ref, &mr);
            report(variant == 0 ? "B0: MN-major A LBO=8192 SBO=1024 + manual swz B" : "B1: MN-major A LBO=1024 SBO=8192 + manual swz B", e, mr);
        }
    }
    {   // ---- C
        std::vector<bf16> Kp(64 * 64); std::vector<float> Kf(64 * 64), X(128 * 64), S0(128 * 64);
        for (int i = 0; i < 64 * 64; ++i) { Kf[i] = bf(rnd()); Kp[i] = __float2bfloat16_rn(Kf[i]); }
        for (int i = 0; i < 128 * 64; ++i) { X[i] = bf(rnd()); S0[i] = rnd(); }
        bf16* dK; float *dX, *dS; CK(cudaMalloc(&dK, Kp.size() * 2)); CK(cudaMalloc(&dX, X.size() * 4)); CK(cudaMalloc(&dS, S0.size() * 4));
        CK(cudaMemcpy(dK, Kp.data(), Kp.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dX, X.data(), X.size() * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dS, S0.data(), S0.size() * 4, cudaMemcpyHostToDevice));
        CUtensorMap mk; uint64_t dk[2] = {64, 64}, sk[1] = {128}; uint32_t bk[2] = {64, 64};
        int r1 = make_tmap(&mk, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dK, dk, sk, bk, CU_TENSOR_MAP_SWIZZLE_128B);
        if (r1) { printf("tensor map encode failed %d\n", r1); return 2; }
        for (int variant = 0; variant < 4; ++variant) {
            const float sgn = (variant & 2) ? -1.f : 1.f;
            for (int v = 0; v < 128; ++v) for (int d = 0; d < 64; ++d) { float s = 0; for (int t = 0; t < 64; ++t) s += X[v * 64 + t] * Kf[t * 64 + d]; ref[v * 64 + d] = S0[v * 64 + d] + sgn * s; }
            CK(cudaMemset(d_out, 0, 128 * 64 * 4));
            probe_c<<<1, 128, smem_bytes>>>(mk, dX, dS, d_out, variant);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(out.data(), d_out, out.size() * 4, cudaMemcpyDeviceToHost));
            double e = max_err(out, ref, &mr);
            char name[96]; snprintf(name, sizeof name, "C%d: TS A(tmem bf16 %s-low)%s, MN-major B, accum", variant, (variant & 1) ? "odd" : "even", (variant & 2) ? ", a_negate" : "");
            report(name, e, mr);
        }
    }
    {   // ---- D
        const int B = 2, F = 3, C = 49, H = 2, K = 64, V = 256, T = F * C;
        std::vector<bf16> Q((size_t)B * T * H * K); std::vector<float> Qf(Q.size());
        for (size_t i = 0; i < Q.size(); ++i) { Qf[i] = bf(rnd()); Q[i] = __float2bfloat16_rn(Qf[i]); }
        bf16 *dQ, *dO; CK(cudaMalloc(&dQ, Q.size() * 2)); CK(cudaMalloc(&dO, (size_t)B * T * H * V * 2));
        CK(cudaMemcpy(dQ, Q.data(), Q.size() * 2, cudaMemcpyHostToDevice));
        CK(cudaMemset(dO, 0, (size_t)B * T * H * V * 2));
        CUtensorMap mq, mo;
        uint64_t dq[5] = {(uint64_t)K, (uint64_t)C, (uint64_t)F, (uint64_t)H, (uint64_t)B};
        uint64_t sq[4] = {(uint64_t)H * K * 2, (uint64_t)C * H * K * 2, (uint64_t)K * 2, (uint64_t)T * H * K * 2};
        uint32_t bq[5] = {64, 64, 1, 1, 1};
        uint64_t dO_[5] = {64, (uint64_t)C, (uint64_t)H * V / 64, (uint64_t)F, (uint64_t)B};
        uint64_t sO_[4] = {(uint64_t)H * V * 2, 128, (uint64_t)C * H * V * 2, (uint64_t)T * H * V * 2};
        uint32_t bO_[5] = {64, 64, 4, 1, 1};
        int r1 = make_tmap(&mq, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dQ, dq, sq, bq, CU_TENSOR_MAP_SWIZZLE_128B);
        int r2 = make_tmap(&mo, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, dO, dO_, sO_, bO_, CU_TENSOR_MAP_SWIZZLE_NONE);
        if (r1 || r2) { printf("rank-5 tensor map encode failed %d %d\n", r1, r2); return 2; }
        const int f = 1, h = 1, b = 1;
        float* d_tile; CK(cudaMalloc(&d_tile, 64 * 64 * 4));
        probe_d<<<1, 128, smem_bytes>>>(mq, mo, d_tile, f, h, b);
        CK(cudaDeviceSynchronize());
        std::vector<float> tile(64 * 64), rt(64 * 64);
        CK(cudaMemcpy(tile.data(), d_tile, tile.size() * 4, cudaMemcpyDeviceToHost));
        for (int r = 0; r < 64; ++r) for (int c = 0; c < 64; ++c)
            rt[r * 64 + c] = r < C ? Qf[(((size_t)b * T + f * C + r) * H + h) * K + c] : 0.f;
        double e = max_err(tile, rt, &mr); report("D1: rank-5 TMA load, frame OOB rows zero-filled", e, mr);
        std::vector<bf16> O((size_t)B * T * H * V);
        CK(cudaMemcpy(O.data(), dO, O.size() * 2, cudaMemcpyDeviceToHost));
        double emax = 0; long bad_outside = 0;
        for (int bb = 0; bb < B; ++bb) for (int t = 0; t < T; ++t) for (int hh = 0; hh < H; ++hh) for (int v = 0; v < V; ++v) {
            const float got = __bfloat162float(O[(((size_t)bb * T + t) * H + hh) * V + v]);
            const bool inside = bb == b && hh == h && t >= f * C && t < (f + 1) * C;
            if (inside) { const float want = bf((float)(t - f * C + 1) + 0.001f * v); emax = fmax(emax, fabs(got - want)); }
            else if (got != 0.f) ++bad_outside;
        }
        printf("%-44s %s  max_abs_err %.3e, writes outside the frame: %ld\n", "D2: rank-5 TMA store clips OOB rows", (emax < 0.3 && bad_outside == 0) ? "PASS" : "FAIL", emax, bad_outside);
    }
    printf("probe done\n");
    return 0;
}
