// Probe: register <-> (lane, column) mapping of tcgen05.ld.16x256b and the memory image of
// stmatrix.x4.trans, needed for the transposing readout epilogue.  Prints the decoded mappings.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../gdkvm_b200/csrc/sm100_ptx.cuh"
using namespace sm100;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ void tmem_ld_16x256b_x8(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}

__global__ void __launch_bounds__(128) probe(float* out_ld, uint16_t* out_st) {
    __shared__ uint32_t tmem_base_s;
    __shared__ __align__(128) uint16_t sm[4 * 64];      // four 8x8 b16 matrices
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) tmem_alloc(&tmem_base_s, 64);
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    const uint32_t tmem = tmem_base_s;
    const uint32_t lane_off = (uint32_t)(warp * 32) << 16;
    uint32_t r[32];
    for (int half = 0; half < 2; ++half) {               // value = 1000*lane + column
        for (int j = 0; j < 32; ++j) r[j] = __float_as_uint((float)(1000 * tid + half * 32 + j));
        tmem_st32(tmem + lane_off + half * 32, r);
    }
    tmem_wait_st();
    tc_fence_before_sync(); __syncthreads(); tc_fence_after_sync();
    for (int grp = 0; grp < 2; ++grp) {                  // lanes +0..15 and +16..31 of the warp's quadrant
        tmem_ld_16x256b_x8(tmem + lane_off + ((uint32_t)(grp * 16) << 16), r);
        tmem_wait_ld();
        for (int j = 0; j < 32; ++j) out_ld[((warp * 2 + grp) * 32 + lane) * 32 + j] = __uint_as_float(r[j]);
    }
    tc_fence_before_sync(); __syncthreads();
    if (warp == 0) {
        // stmatrix.x4.trans: thread holds for matrix m the pair (elem 2*(lane%4), +1) of row lane/4 in register m
        uint32_t v[4];
        for (int m = 0; m < 4; ++m) {
            const uint32_t lo = (uint32_t)(m * 1000 + (lane >> 2) * 10 + (lane & 3) * 2);       // row*10 + col
            v[m] = lo | ((lo + 1) << 16);
        }
        const uint32_t addr = smem_u32(sm) + (lane >> 3) * 128 + (lane & 7) * 16;                 // matrix lane/8, row lane%8
        asm volatile("stmatrix.sync.aligned.m8n8.x4.trans.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]) : "memory");
        __syncwarp();
        for (int i = lane; i < 256; i += 32) out_st[i] = sm[i];
    }
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 64);
}

int main() {
    float* d_ld; uint16_t* d_st;
    CK(cudaMalloc(&d_ld, 8 * 32 * 32 * 4)); CK(cudaMalloc(&d_st, 256 * 2));
    probe<<<1, 128>>>(d_ld, d_st);
    CK(cudaDeviceSynchronize());
    std::vector<float> ld(8 * 32 * 32); std::vector<uint16_t> st(256);
    CK(cudaMemcpy(ld.data(), d_ld, ld.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st.data(), d_st, st.size() * 2, cudaMemcpyDeviceToHost));
    // expected: reg j = 4q + 2h + e  ->  lane = grp*16 + (t/4) + 8h, column = 8q + 2(t%4) + e
    int bad = 0;
    for (int w = 0; w < 4; ++w) for (int grp = 0; grp < 2; ++grp) for (int t = 0; t < 32; ++t) for (int j = 0; j < 32; ++j) {
        const int q = j >> 2, hh = (j >> 1) & 1, e = j & 1;
        const int lane_exp = w * 32 + grp * 16 + (t >> 2) + 8 * hh, col_exp = 8 * q + 2 * (t & 3) + e;
        const float got = ld[((w * 2 + grp) * 32 + t) * 32 + j];
        if (got != (float)(1000 * lane_exp + col_exp)) { if (bad < 6) printf("  ld mismatch w%d grp%d t%d reg%d: got %.0f expected lane %d col %d\n", w, grp, t, j, got, lane_exp, col_exp); ++bad; }
    }
    printf("E1: tcgen05.ld.16x256b.x8 = mma C-fragment layout (reg 4q+2h+e -> lane t/4+8h, col 8q+2(t%%4)+e): %s (%d mismatches)\n", bad ? "FAIL" : "PASS", bad);
    if (bad) { printf("  first thread regs (w0 grp0 t0): "); for (int j = 0; j < 8; ++j) printf("%.0f ", ld[j]); printf("\n  t1: "); for (int j = 0; j < 8; ++j) printf("%.0f ", ld[32 + j]); printf("\n"); }
    // stmatrix.trans expected: memory row c of matrix m (16 B) holds elements (row r = 0..7, col c): value m*1000 + r*10 + c
    int bad2 = 0;
    for (int m = 0; m < 4; ++m) for (int c = 0; c < 8; ++c) for (int rr = 0; rr < 8; ++rr)
        if (st[m * 64 + c * 8 + rr] != (uint16_t)(m * 1000 + rr * 10 + c)) { if (bad2 < 6) printf("  st mismatch m%d memrow%d pos%d: got %u\n", m, c, rr, st[m * 64 + c * 8 + rr]); ++bad2; }
    printf("E2: stmatrix.x4.trans writes fragment (row r, col c) to memory row c, position r: %s (%d mismatches)\n", bad2 ? "FAIL" : "PASS", bad2);
    return 0;
}
