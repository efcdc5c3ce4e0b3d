// Row-wise L2 normalisation of q / k before the memory op:  y = x * rsqrt(sum(x^2) + eps).
//
// SURVEY.md section 8f rank 3 (the step immediately before the op; fla's `use_qk_l2norm_in_kernel`,
// fla/ops/gated_delta_rule/chunk.py:374).  This is the UNFUSED first step: one streaming pass, 16-byte vector loads and
// stores, D/8 (bf16) or D/4 (fp32) lanes per row with a shuffle reduction; HBM-bound (reads and writes every byte once).
// Folding it into the chunk kernel was measured and rejected; when q and k come out of a projection, the
// fused projection kernel (gdr_proj_sm100.cu) normalises them in its epilogue and this pass is not needed at all.
#include "gdr_common.cuh"

namespace gdkvm {
namespace {

template <typename T, int VEC, int NV>   // VEC elements per 16-byte vector, NV vectors per thread (row = lanes x NV vectors)
__global__ void __launch_bounds__(256) l2norm_rows_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int D,
                                                          int64_t xs, int64_t ys, float eps) {
    const int lanes = D / (VEC * NV);                             // lanes per row: power of two, <= 32
    const int rows_per_block = 256 / lanes;
    const int sub = threadIdx.x % lanes, local_row = threadIdx.x / lanes;
    for (int64_t base = (int64_t)blockIdx.x * rows_per_block; base < rows; base += (int64_t)gridDim.x * rows_per_block) {
        const int64_t row = base + local_row;
        const bool valid = row < rows;                            // the whole warp stays in the loop: full-mask shuffles below
        float v[NV][VEC];
        float ss = 0.f;
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            const uint4 raw = valid ? *reinterpret_cast<const uint4*>(x + row * xs + (n * lanes + sub) * VEC) : make_uint4(0u, 0u, 0u, 0u);
            if constexpr (VEC == 8) {
                const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
                for (int i = 0; i < 4; ++i) { v[n][2 * i] = __uint_as_float(w[i] << 16); v[n][2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
            } else {
                v[n][0] = __uint_as_float(raw.x); v[n][1] = __uint_as_float(raw.y); v[n][2] = __uint_as_float(raw.z); v[n][3] = __uint_as_float(raw.w);
            }
#pragma unroll
            for (int i = 0; i < VEC; ++i) ss = fmaf(v[n][i], v[n][i], ss);
        }
        for (int off = lanes >> 1; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);   // lanes of a row are contiguous
        const float r = rsqrtf(ss + eps);
#pragma unroll
        for (int n = 0; n < NV; ++n) {
            uint4 out;
            if constexpr (VEC == 8) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __nv_bfloat162 p = __floats2bfloat162_rn(v[n][2 * i] * r, v[n][2 * i + 1] * r);
                    w[i] = *reinterpret_cast<const uint32_t*>(&p);
                }
                out = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
                out = make_uint4(__float_as_uint(v[n][0] * r), __float_as_uint(v[n][1] * r), __float_as_uint(v[n][2] * r), __float_as_uint(v[n][3] * r));
            }
            if (valid) *reinterpret_cast<uint4*>(y + row * ys + (n * lanes + sub) * VEC) = out;
        }
    }
}

}  // namespace

int launch_l2norm(const void* x, void* y, int64_t rows, int D, int64_t xs, int64_t ys, int dtype, float eps, cudaStream_t stream) {
    if (rows == 0) return 0;
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int vec = dtype == GDKVM_BF16 ? 8 : 4;
    const int nv = D / vec > 32 ? 2 : 1;                          // fp32 rows of 256: two vectors per lane
    const int rows_per_block = 256 / (D / (vec * nv));
    int64_t blocks = (rows + rows_per_block - 1) / rows_per_block;
    const int64_t cap = (int64_t)sms * 8;                        // 8 resident CTAs of 256 threads per SM, grid-stride beyond
    if (blocks > cap) blocks = cap;
    if (dtype == GDKVM_BF16)
        l2norm_rows_kernel<__nv_bfloat16, 8, 1><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), static_cast<__nv_bfloat16*>(y), rows, D, xs, ys, eps);
    else if (nv == 1)
        l2norm_rows_kernel<float, 4, 1><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(y), rows, D, xs, ys, eps);
    else
        l2norm_rows_kernel<float, 4, 2><<<(unsigned)blocks, 256, 0, stream>>>(static_cast<const float*>(x), static_cast<float*>(y), rows, D, xs, ys, eps);
    count_launch();
    return (int)cudaGetLastError();
}

}  // namespace gdkvm
