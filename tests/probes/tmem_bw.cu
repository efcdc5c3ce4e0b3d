// Probe: TMEM load/store throughput per SM as a function of the number of warps issuing
// tcgen05.ld / tcgen05.st 32x32b.x32 (4 KB per instruction per warp), plus one isolated round trip.
#include <cstdio>
#include <cstdlib>
#include "../../gdkvm_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__global__ void __launch_bounds__(512) tmem_bw(long long* out, int nwarps, int mode, int reps) {
    __shared__ uint32_t tmem_s;
    __shared__ long long t_first, t_last;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (warp == 0) tmem_alloc(&tmem_s, 512);
    if (tid == 0) { t_first = 0x7fffffffffffffffLL; t_last = 0; }
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_s;
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = tid + j;
    const uint32_t addr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 64);
    tmem_st32(addr, r); tmem_st32(addr + 32, r); tmem_wait_st();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    uint32_t acc = 0;
    if (warp < nwarps) {
        const long long t0 = clock64();
        for (int i = 0; i < reps; ++i) {
            if (mode == 0) {            // load, wait after every instruction (latency-bound per warp)
                tmem_ld32(addr + (i & 1) * 32, r); tmem_wait_ld();
                acc += r[0] ^ r[31];
            } else if (mode == 1) {     // two loads in flight per warp
                uint32_t r2[32];
                tmem_ld32(addr, r); tmem_ld32(addr + 32, r2); tmem_wait_ld();
                acc += r[0] ^ r2[31];
            } else if (mode == 2) {     // store, wait after every instruction
                r[0] = acc + i;
                tmem_st32(addr + (i & 1) * 32, r); tmem_wait_st();
            } else {                    // ld -> modify -> st (the S pass pattern)
                tmem_ld32(addr + (i & 1) * 32, r); tmem_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * 0.999f);
                tmem_st32(addr + (i & 1) * 32, r); tmem_wait_st();
            }
        }
        const long long t1 = clock64();
        if ((tid & 31) == 0) { atomicMin((unsigned long long*)&t_first, (unsigned long long)t0); atomicMax((unsigned long long*)&t_last, (unsigned long long)t1); }
    }
    if (acc == 0x12345678u) out[3] = acc;
    tc_fence_before_sync(); __syncthreads();
    if (tid == 0) { out[0] = t_last - t_first; }
    if (warp == 0) tmem_dealloc(tmem, 512);
}

int main() {
    long long* d; CK(cudaMalloc(&d, 64));
    const char* names[4] = {"ld x32, wait each      ", "2 x ld x32, wait once   ", "st x32, wait each      ", "ld -> fmul -> st        "};
    for (int mode = 0; mode < 4; ++mode)
        for (int nw : {1, 4, 8, 16}) {
            const int reps = 64;
            tmem_bw<<<1, 512>>>(d, nw, mode, reps);
            CK(cudaDeviceSynchronize());
            long long h; CK(cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost));
            const double instr = (double)reps * (mode == 1 ? 2 : 1);
            printf("%s %2d warps: %7.1f cycles per x32 instruction per warp, %6.1f B/cycle/SM\n", names[mode], nw,
                   (double)h / instr, instr * nw * 4096.0 * (mode == 3 ? 2 : 1) / (double)h);
        }
    return 0;
}
