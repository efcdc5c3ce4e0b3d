// Probe: the in-register block triangular inverse (gdkvm_b200/csrc/tri_solve.cuh) against a double-precision
// inverse on the host, for random and for strongly correlated keys.  Also times it (cycles, one CTA).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../gdkvm_b200/csrc/tri_solve.cuh"
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)
using namespace gdkvm::tri;

__global__ void __launch_bounds__(256) solve_kernel(const float* A, float* X, long long* cycles) {
    __shared__ __align__(1024) uint8_t H[8192];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        *reinterpret_cast<__half*>(H + sw128_offset(i, j >> 3) + (j & 7) * 2) = __float2half_rn(j < i ? A[idx] : 0.f);
    }
    __syncthreads();
    const uint32_t h = sm100::smem_u32(H);
    const long long t0 = clock64();
    if (warp < 2) solve_levels01(H, h, warp, lane);
    if (warp < 4) {
        asm volatile("bar.sync 5, 128;" ::: "memory");
        solve_level2(h, warp, lane, 5);
    }
    __syncthreads();
    const long long t1 = clock64();
    if (tid == 0) cycles[0] = t1 - t0;
    for (int idx = tid; idx < 4096; idx += 256) {
        const int i = idx >> 6, j = idx & 63;
        X[idx] = __half2float(*reinterpret_cast<const __half*>(H + sw128_offset(i, j >> 3) + (j & 7) * 2));
    }
}

int main() {
    float *dA, *dX; long long* dC;
    CK(cudaMalloc(&dA, 4096 * 4)); CK(cudaMalloc(&dX, 4096 * 4)); CK(cudaMalloc(&dC, 8));
    int fails = 0;
    for (int cs = 0; cs < 3; ++cs) {
        std::vector<float> k(64 * 64), beta(64), A(4096, 0.f), X(4096);
        srand(17 + cs);
        auto rnd = [] { return (float)rand() / RAND_MAX * 2.f - 1.f; };
        std::vector<float> base(64);
        for (auto& b : base) b = rnd();
        const float mix = cs == 0 ? 0.f : (cs == 1 ? 0.7f : 0.95f);      // key correlation inside the chunk
        for (int i = 0; i < 64; ++i) {
            double n2 = 0;
            for (int d = 0; d < 64; ++d) { k[i * 64 + d] = mix * base[d] + (1.f - mix) * rnd(); n2 += (double)k[i * 64 + d] * k[i * 64 + d]; }
            for (int d = 0; d < 64; ++d) k[i * 64 + d] /= (float)std::sqrt(n2);
            beta[i] = 0.5f + 0.5f * (float)rand() / RAND_MAX;
        }
        for (int i = 0; i < 64; ++i)
            for (int j = 0; j < i; ++j) {
                double dot = 0;
                for (int d = 0; d < 64; ++d) dot += (double)k[i * 64 + d] * k[j * 64 + d];
                A[i * 64 + j] = beta[i] * (float)dot;
            }
        std::vector<double> R(4096, 0.0);                                  // exact inverse of I + A by forward substitution
        for (int c = 0; c < 64; ++c)
            for (int i = 0; i < 64; ++i) {
                double s = i == c ? 1.0 : 0.0;
                for (int j = 0; j < i; ++j) s -= (double)A[i * 64 + j] * R[j * 64 + c];
                R[i * 64 + c] = s;
            }
        CK(cudaMemcpy(dA, A.data(), 4096 * 4, cudaMemcpyHostToDevice));
        for (int rep = 0; rep < 3; ++rep) solve_kernel<<<1, 256>>>(dA, dX, dC);
        CK(cudaDeviceSynchronize());
        long long cyc; CK(cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(X.data(), dX, 4096 * 4, cudaMemcpyDeviceToHost));
        double maxerr = 0, maxref = 0;
        for (int i = 0; i < 4096; ++i) { maxerr = std::fmax(maxerr, std::fabs(X[i] - R[i])); maxref = std::fmax(maxref, std::fabs(R[i])); }
        const bool ok = maxerr <= 4e-3 * maxref;
        fails += !ok;
        printf("solve64 mix %.2f: max|X - ref| %.3e (max|ref| %.3f)  %s   %lld cycles\n", mix, maxerr, maxref, ok ? "PASS" : "FAIL", cyc);
    }
    return fails;
}
