// Token-recurrent GDR/LKVA kernel: exact fp32 math on the CUDA cores of sm_100a.
//
// Role: (1) the fp32-I/O path (max-rel <= 1e-3 contract, in practice ~1e-6 against the oracle),
//       (2) the CUDA path for shapes the tcgen05 chunk kernel does not cover (never a CPU fallback).
//
// Mapping: one CTA = one (clip, head) chain x 128 (or, NC = 2, 256) value columns.  Thread x owns column x (and x + 128) of the
// state: S[0..K-1][x] lives in K registers for the whole clip, so the three per-token contractions
//     dot = S^T k,   S = a S + k (beta (v - a dot)),   o = scale S^T q
// are thread-local FMAs with k_t / q_t broadcast from shared memory; the state never touches HBM
// between the initial load and the final store.  Tokens are staged TB at a time: the next tile's
// q/k/v/g/beta are fetched into registers while the current tile is being consumed.
//
// Follows oracle/gdr_ref.py::gdr_recurrent_ref operation for operation.
#include "gdr_common.cuh"

namespace gdkvm {
namespace {

constexpr int kThreads = 128;
#ifndef GDKVM_REC_NC2_MINB
#define GDKVM_REC_NC2_MINB 2      // CTAs per SM of the two-columns-per-thread variant: 2 = 253 registers, no spills
#endif

template <typename TIO, int N>
__device__ __forceinline__ void load_vec(const TIO* __restrict__ ptr, float (&out)[N]) {
    constexpr int kBytes = N * (int)sizeof(TIO);
    if constexpr (kBytes % 16 == 0) {
        uint4 raw[kBytes / 16];
#pragma unroll
        for (int i = 0; i < kBytes / 16; ++i) raw[i] = __ldg(reinterpret_cast<const uint4*>(ptr) + i);
        const TIO* e = reinterpret_cast<const TIO*>(raw);
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = to_f32(e[i]);
    } else {
        static_assert(kBytes % 8 == 0, "row chunk must be at least 8 bytes");
        uint2 raw[kBytes / 8];
#pragma unroll
        for (int i = 0; i < kBytes / 8; ++i) raw[i] = __ldg(reinterpret_cast<const uint2*>(ptr) + i);
        const TIO* e = reinterpret_cast<const TIO*>(raw);
#pragma unroll
        for (int i = 0; i < N; ++i) out[i] = to_f32(e[i]);
    }
}

// NC = value columns per thread (columns x, x + 128, ...).  With NC = 2 every broadcast k / q load from shared memory feeds
// twice the FMAs: the NC = 1 kernel is bound by those loads (48 LDS.128 per 256 FP instructions, 16.5 ms at configs[1]).
template <int K, typename TIO, int NC>
__global__ void __launch_bounds__(kThreads, (K <= 64 ? (NC > 1 ? GDKVM_REC_NC2_MINB : 3) : 1)) gdr_recurrent_kernel(const GdkvmGdrParams p, const void* __restrict__ cu, const int cu_bytes) {
    constexpr int TB = (K >= 128 || NC > 1) ? 8 : 16;   // tokens per staged tile (static smem < 48 KB, staging registers)
    constexpr int EPT = TB * K / kThreads;  // q/k elements each thread stages per tile
    static_assert(EPT >= 4 && K % EPT == 0, "unsupported K");

    __shared__ __align__(16) float s_q[2][TB][K];
    __shared__ __align__(16) float s_k[2][TB][K];
    __shared__ float s_v[2][TB][kThreads * NC];
    __shared__ float s_alpha[2][TB];
    __shared__ float s_beta[2][TB];

    const int tid = threadIdx.x;
    const int chain = blockIdx.x;
    int b = chain / p.H;
    const int h = chain % p.H;
    const int x = blockIdx.y * (kThreads * NC) + tid;          // first of this thread's columns
    bool col_ok[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) col_ok[c] = x + c * kThreads < p.V;
    int T = p.T;
    const int V = p.V;
    int64_t tok0 = 0;
    if (cu != nullptr) {   // packed variable-length sequences: chain = (sequence, head), rows cu[seq] .. cu[seq+1]-1 of clip 0
        const int64_t lo = cu_bytes == 8 ? reinterpret_cast<const long long*>(cu)[b] : reinterpret_cast<const int*>(cu)[b];
        const int64_t hi = cu_bytes == 8 ? reinterpret_cast<const long long*>(cu)[b + 1] : reinterpret_cast<const int*>(cu)[b + 1];
        tok0 = lo; T = (int)(hi - lo); b = 0;
    }

    const TIO* q_base = reinterpret_cast<const TIO*>(p.q) + (int64_t)b * p.q_stride[0] + (int64_t)h * p.q_stride[2] + tok0 * p.q_stride[1];
    const TIO* k_base = reinterpret_cast<const TIO*>(p.k) + (int64_t)b * p.k_stride[0] + (int64_t)h * p.k_stride[2] + tok0 * p.k_stride[1];
    const TIO* v_base = reinterpret_cast<const TIO*>(p.v) + (int64_t)b * p.v_stride[0] + (int64_t)h * p.v_stride[2] + tok0 * p.v_stride[1];
    TIO* o_base = reinterpret_cast<TIO*>(p.o) + (int64_t)b * p.o_stride[0] + (int64_t)h * p.o_stride[2] + tok0 * p.o_stride[1];
    const int64_t g_off = (int64_t)b * p.g_stride[0] + (int64_t)h * p.g_stride[2] + tok0 * p.g_stride[1];
    const int64_t bt_off = (int64_t)b * p.beta_stride[0] + (int64_t)h * p.beta_stride[2] + tok0 * p.beta_stride[1];

    // state column in registers
    float S[NC][K];
#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (p.initial_state != nullptr && col_ok[c]) {
            const float* s0 = p.initial_state + (int64_t)chain * K * V + x + c * kThreads;
#pragma unroll
            for (int d = 0; d < K; ++d) S[c][d] = __ldg(s0 + (int64_t)d * V);
        } else {
#pragma unroll
            for (int d = 0; d < K; ++d) S[c][d] = 0.f;
        }
    }

    const int n_tiles = (T + TB - 1) / TB;
    const int st_row = (tid * EPT) / K;   // which token of the tile this thread stages
    const int st_col = (tid * EPT) % K;

    float rq[EPT], rk[EPT], rv[NC][TB], rgate = 0.f;

    auto fetch = [&](int tile) {
        const int t0 = tile * TB;
        const int t = t0 + st_row;
        if (t < T) {
            load_vec<TIO, EPT>(q_base + (int64_t)t * p.q_stride[1] + st_col, rq);
            load_vec<TIO, EPT>(k_base + (int64_t)t * p.k_stride[1] + st_col, rk);
        } else {
#pragma unroll
            for (int i = 0; i < EPT; ++i) { rq[i] = 0.f; rk[i] = 0.f; }
        }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int j = 0; j < TB; ++j)
                rv[c][j] = (col_ok[c] && t0 + j < T) ? to_f32(v_base[(int64_t)(t0 + j) * p.v_stride[1] + x + c * kThreads]) : 0.f;
        if (tid < TB) {
            const int tt = t0 + tid;   // alpha = exp(g); pad tokens are exact no-ops
            rgate = tt < T ? expf(load_gate(p.g, g_off + (int64_t)tt * p.g_stride[1], p.gate_dtype)) : 1.f;
        } else if (tid < 2 * TB) {
            const int tt = t0 + tid - TB;
            rgate = tt < T ? load_gate(p.beta, bt_off + (int64_t)tt * p.beta_stride[1], p.gate_dtype) : 0.f;
        }
    };
    auto stage = [&](int buf) {
#pragma unroll
        for (int i = 0; i < EPT; ++i) { s_q[buf][st_row][st_col + i] = rq[i]; s_k[buf][st_row][st_col + i] = rk[i]; }
#pragma unroll
        for (int c = 0; c < NC; ++c)
#pragma unroll
            for (int j = 0; j < TB; ++j) s_v[buf][j][tid + c * kThreads] = rv[c][j];
        if (tid < TB) s_alpha[buf][tid] = rgate;
        else if (tid < 2 * TB) s_beta[buf][tid - TB] = rgate;
    };

    if (n_tiles > 0) { fetch(0); stage(0); }
    __syncthreads();

    const float scale = p.scale;
    for (int tile = 0; tile < n_tiles; ++tile) {
        const int buf = tile & 1;
        if (tile + 1 < n_tiles) fetch(tile + 1);
        const int t0 = tile * TB;
        const int n_tok = min(TB, T - t0);
#pragma unroll 1
        for (int j = 0; j < n_tok; ++j) {
            const float a = s_alpha[buf][j];
            const float bt = s_beta[buf][j];
            const float4* k4 = reinterpret_cast<const float4*>(&s_k[buf][j][0]);
            const float4* q4 = reinterpret_cast<const float4*>(&s_q[buf][j][0]);
            float dt[NC][4];
#pragma unroll
            for (int c = 0; c < NC; ++c) dt[c][0] = dt[c][1] = dt[c][2] = dt[c][3] = 0.f;
#pragma unroll
            for (int d = 0; d < K / 4; ++d) {
                const float4 kk = k4[d];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    dt[c][0] = fmaf(S[c][4 * d + 0], kk.x, dt[c][0]);
                    dt[c][1] = fmaf(S[c][4 * d + 1], kk.y, dt[c][1]);
                    dt[c][2] = fmaf(S[c][4 * d + 2], kk.z, dt[c][2]);
                    dt[c][3] = fmaf(S[c][4 * d + 3], kk.w, dt[c][3]);
                }
            }
            // (a S)^T k = a (S^T k);  r = beta (v - that)
            float r[NC], ot[NC][4];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                r[c] = bt * (s_v[buf][j][tid + c * kThreads] - a * ((dt[c][0] + dt[c][1]) + (dt[c][2] + dt[c][3])));
                ot[c][0] = ot[c][1] = ot[c][2] = ot[c][3] = 0.f;
            }
#pragma unroll
            for (int d = 0; d < K / 4; ++d) {
                const float4 kk = k4[d];
                const float4 qq = q4[d];
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    S[c][4 * d + 0] = fmaf(a, S[c][4 * d + 0], kk.x * r[c]);
                    S[c][4 * d + 1] = fmaf(a, S[c][4 * d + 1], kk.y * r[c]);
                    S[c][4 * d + 2] = fmaf(a, S[c][4 * d + 2], kk.z * r[c]);
                    S[c][4 * d + 3] = fmaf(a, S[c][4 * d + 3], kk.w * r[c]);
                    ot[c][0] = fmaf(S[c][4 * d + 0], qq.x, ot[c][0]);
                    ot[c][1] = fmaf(S[c][4 * d + 1], qq.y, ot[c][1]);
                    ot[c][2] = fmaf(S[c][4 * d + 2], qq.z, ot[c][2]);
                    ot[c][3] = fmaf(S[c][4 * d + 3], qq.w, ot[c][3]);
                }
            }
#pragma unroll
            for (int c = 0; c < NC; ++c)
                if (col_ok[c])
                    from_f32(o_base[(int64_t)(t0 + j) * p.o_stride[1] + x + c * kThreads], scale * ((ot[c][0] + ot[c][1]) + (ot[c][2] + ot[c][3])));
        }
        if (tile + 1 < n_tiles) stage(buf ^ 1);
        __syncthreads();
    }

#pragma unroll
    for (int c = 0; c < NC; ++c) {
        if (p.final_state != nullptr && col_ok[c]) {
            float* sT = p.final_state + (int64_t)chain * K * V + x + c * kThreads;
#pragma unroll
            for (int d = 0; d < K; ++d) sT[(int64_t)d * V] = S[c][d];
        }
    }
}

template <int K>
int launch_k(const GdkvmGdrParams& p, const void* cu, int cu_bytes, int nseq, cudaStream_t stream) {
    const int chains = (cu != nullptr ? nseq : p.B) * p.H;
    if constexpr (K == 64) {
        // two value columns per thread when that still leaves every SM several CTAs (wide value dimension, enough chains)
        if (p.V > kThreads && (int64_t)chains * ((p.V + 2 * kThreads - 1) / (2 * kThreads)) >= 3 * 148) {
            dim3 grid2(chains, (p.V + 2 * kThreads - 1) / (2 * kThreads));
            if (p.io_dtype == GDKVM_BF16) gdr_recurrent_kernel<K, __nv_bfloat16, 2><<<grid2, kThreads, 0, stream>>>(p, cu, cu_bytes);
            else gdr_recurrent_kernel<K, float, 2><<<grid2, kThreads, 0, stream>>>(p, cu, cu_bytes);
            count_launch();
            return (int)cudaGetLastError();
        }
    }
    dim3 grid(chains, (p.V + kThreads - 1) / kThreads);
    if (p.io_dtype == GDKVM_BF16) gdr_recurrent_kernel<K, __nv_bfloat16, 1><<<grid, kThreads, 0, stream>>>(p, cu, cu_bytes);
    else gdr_recurrent_kernel<K, float, 1><<<grid, kThreads, 0, stream>>>(p, cu, cu_bytes);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace

int launch_recurrent(const GdkvmGdrParams& p, cudaStream_t stream) { return launch_recurrent_varlen(p, nullptr, 0, 0, stream); }

int launch_recurrent_varlen(const GdkvmGdrParams& p, const void* cu, int cu_bytes, int nseq, cudaStream_t stream) {
    switch (p.K) {
        case 32: return launch_k<32>(p, cu, cu_bytes, nseq, stream);
        case 64: return launch_k<64>(p, cu, cu_bytes, nseq, stream);
        case 128: return launch_k<128>(p, cu, cu_bytes, nseq, stream);
        default: return (int)cudaErrorInvalidValue;
    }
}

}  // namespace gdkvm
