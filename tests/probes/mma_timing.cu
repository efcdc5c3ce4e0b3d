// Probe: cycles per tcgen05.mma (cta_group::1, kind::f16, bf16) for the shapes/operand sources the
// chunk kernel uses, on an otherwise idle SM: back-to-back issue of `reps` MMAs, one commit, wait.
#include <cstdio>
#include <cstdlib>
#include "../../gdkvm_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// mode: 0 SS (A K-major smem), 1 SS (A MN-major smem), 2 TS (A in TMEM, B K-major), 3 TS (B MN-major)
__device__ __forceinline__ void umma_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc) {
    asm volatile(
        "{\n\t.reg .pred p, q;\n\telect.sync _|q, 0xffffffff;\n\tsetp.ne.b32 p, 1, 0;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc) : "memory");
}
__global__ void __launch_bounds__(128) timing(long long* out, int mode, int N, int reps, int same_acc) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_s;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 65536 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // small bf16 values
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_s, 512);
    fence_proxy_async_smem();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_s;
    if (mode >= 4 && warp == 0) {          // warp-uniform issue loop, elected lane issues
        // submode: 4 TS B K-major, 5 TS B MN-major, 6 SS A,B K-major, 7 SS A MN-major (B K-major)
        const uint32_t a = smem_u32(smem), b = a + 32768;
        const uint32_t idesc = umma_idesc_bf16(128, N, mode == 7, mode == 5);
        const uint64_t bd0 = mode == 5 ? umma_smem_desc_sw128(b, 8192, 1024) : umma_smem_desc_sw128(b, 16, 1024);
        const uint64_t ad0 = mode == 7 ? umma_smem_desc_sw128(a, 8192, 1024) : umma_smem_desc_sw128(a, 16, 1024);
        const uint32_t bstep = mode == 5 ? 128 : 2, astep = mode == 7 ? 128 : 2;
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t d = tmem + (same_acc ? 0 : (uint32_t)((i & 1) * N));
                const int k = i & 3;
                if (mode < 6) umma_ts_elect(d, tmem + 256 + k * 8, bd0 + (uint64_t)(k * bstep), idesc);
                else umma_ss_w(d, ad0 + (uint64_t)(k * astep), bd0 + (uint64_t)(k * bstep), idesc, true);
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(&bar);
            __syncwarp();
            mbar_wait(&bar, phase); phase ^= 1;
            const long long t2 = clock64();
            if (tid == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        }
    }
    if (mode < 4 && tid == 0) {
        const uint32_t a = smem_u32(smem), b = a + 32768;
        const uint32_t idesc = umma_idesc_bf16(128, N, mode == 1, mode == 3);
        uint32_t phase = 0;
        for (int rep = 0; rep < 3; ++rep) {       // warm-up rounds, keep the last
            const long long t0 = clock64();
            for (int i = 0; i < reps; ++i) {
                const uint32_t d = tmem + (same_acc ? 0 : (uint32_t)((i & 1) * N));
                const int k = i & 3;
                if (mode == 0) umma_ss(d, umma_smem_desc_sw128(a + k * 32, 16, 1024), umma_smem_desc_sw128(b + k * 32, 16, 1024), idesc, true);
                else if (mode == 1) umma_ss(d, umma_smem_desc_sw128(a + k * 2048, 8192, 1024), umma_smem_desc_sw128(b + k * 32, 16, 1024), idesc, true);
                else if (mode == 2) umma_ts(d, tmem + 256 + k * 8, umma_smem_desc_sw128(b + k * 32, 16, 1024), idesc, true);
                else umma_ts(d, tmem + 256 + k * 8, umma_smem_desc_sw128(b + k * 2048, 8192, 1024), idesc, true);
            }
            const long long t1 = clock64();
            umma_commit(&bar);
            mbar_wait(&bar, phase); phase ^= 1;
            const long long t2 = clock64();
            out[0] = t1 - t0; out[1] = t2 - t0;
        }
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 16));
    CK(cudaFuncSetAttribute(timing, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    const char* names[8] = {"SS A K-major ", "SS A MN-major", "TS B K-major ", "TS B MN-major", "uniform TS B K-major ", "uniform TS B MN-major",
                            "uniform SS K-major   ", "uniform SS A MN-major"};
    for (int N : {64, 128})
        for (int mode = 4; mode < 8; ++mode)
            for (int same = 0; same < 2; ++same) {
                const int reps = 64;
                timing<<<1, 128, 100 * 1024>>>(d, mode, N, reps, same);
                CK(cudaDeviceSynchronize());
                long long h[2]; CK(cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost));
                printf("M=128 N=%3d %s %s: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA\n", N, names[mode], same ? "same accumulator " : "2 accumulators   ",
                       (double)h[0] / reps, (double)h[1] / reps);
            }
    return 0;
}
