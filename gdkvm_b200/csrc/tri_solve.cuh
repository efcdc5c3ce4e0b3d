// (I + A)^-1 of a 64 x 64 strictly-lower-triangular A, held as fp16 in shared memory in the 128-byte
// swizzled row layout of the operand tiles (row i = 128 bytes, 16-byte chunk c stored at c ^ (i & 7)).
//
// Block recursion over 16 x 16 blocks, operands and intermediates in registers (mma.sync m16n8k16 with
// ldmatrix / stmatrix fragments; the accumulator of the first product is re-used as the A operand of the
// second one), so a level needs no shared-memory round trip and no block-wide barrier:
//   level 0    8 x 8 diagonal blocks: (I + A_bb)^-1       8-step forward substitution, fp32, 64 threads
//   level 0.5  8 -> 16: X_ba = -(X_bb L_ba) X_aa          two mma per 16 x 16 block (zero-padded fragments), two blocks per warp
//   level 1  X_ba = -(X_bb L_ba) X_aa                   (a, b) = (0,1), (2,3): one warp each
//   level 2  [X_20 X_21; X_30 X_31] = -Xbr (Lbl Xtl)    four warps, one 16 x 16 output block each
// On entry H holds A (strict lower triangle, zeros elsewhere, fp16); on exit H holds X = (I + A)^-1
// (unit diagonal, zeros above).  fp16 carries 11 significant bits; every entry is O(1) and the result is
// rounded to bf16 afterwards (tests/chunk_numerics_model.py, tests/probes/solve_probe.cu).
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

#include "sm100_ptx.cuh"

namespace gdkvm {
namespace tri {

using sm100::sw128_offset;

__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
__device__ __forceinline__ float2 unpack_f16(uint32_t w) {
    return __half22float2(*reinterpret_cast<const __half2*>(&w));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x1_trans(uint32_t& r, uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x1.trans.shared.b16 {%0}, [%1];" : "=r"(r) : "r"(addr));
}
__device__ __forceinline__ void stsm_x4(uint32_t addr, uint32_t r0, uint32_t r1, uint32_t r2, uint32_t r3) {
    asm volatile("stmatrix.sync.aligned.m8n8.x4.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(r0), "r"(r1), "r"(r2), "r"(r3)
                 : "memory");
}

// A-operand fragment (16 x 16, rows = block rows) of block (br, bc) of the row-major matrix
__device__ __forceinline__ void frag_a(uint32_t (&a)[4], uint32_t h, int br, int bc, int lane) {
    ldsm_x4(a, h + sw128_offset(16 * br + (lane & 15), 2 * bc + (lane >> 4)));
}
// B-operand fragments (k = block rows, n = block columns; two n8 tiles: {b[0], b[1]} and {b[2], b[3]})
__device__ __forceinline__ void frag_b(uint32_t (&b)[4], uint32_t h, int br, int bc, int lane) {
    ldsm_x4_trans(b, h + sw128_offset(16 * br + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * bc + (lane >> 4)));
}
// c[2][4] (16 x 16 as two n8 tiles) += A B
__device__ __forceinline__ void mma_block(float (&c)[2][4], const uint32_t (&a)[4], const uint32_t (&b)[4]) {
    mma_f16(c[0], a, b[0], b[1]);
    mma_f16(c[1], a, b[2], b[3]);
}
// accumulator (16 x 16) -> A-operand fragment of the same matrix
__device__ __forceinline__ void acc_to_a(uint32_t (&a)[4], const float (&c)[2][4]) {
    a[0] = pack_f16(c[0][0], c[0][1]);
    a[1] = pack_f16(c[0][2], c[0][3]);
    a[2] = pack_f16(c[1][0], c[1][1]);
    a[3] = pack_f16(c[1][2], c[1][3]);
}
// block (br, bc) <- -c
__device__ __forceinline__ void store_neg_block(uint32_t h, int br, int bc, const float (&c)[2][4], int lane) {
    stsm_x4(h + sw128_offset(16 * br + (lane & 7) + 8 * ((lane >> 3) & 1), 2 * bc + (lane >> 4)),
            pack_f16(-c[0][0], -c[0][1]), pack_f16(-c[0][2], -c[0][3]), pack_f16(-c[1][0], -c[1][1]), pack_f16(-c[1][2], -c[1][3]));
}

// Level 0 + level 1: warps 0 and 1 (64 threads).  Thread t of the pair: diagonal block t / 16, column t % 16.
// hs = generic pointer to H, h = its shared-window address.
__device__ __forceinline__ void solve_levels01(uint8_t* hs, uint32_t h, int warp, int lane) {
    {   // level 0: the eight 8 x 8 diagonal blocks, one column per thread: (I + L)^-1 e_c by forward substitution in fp32.
        // Row i of block b8 is ONE 16-byte chunk (chunk b8 of row 8 b8 + i).
        const int b8 = 4 * warp + (lane >> 3), c = lane & 7;
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float s = (i == c) ? 1.f : 0.f;
            if (i > 0) {
                const uint4 rw = *reinterpret_cast<const uint4*>(hs + sw128_offset(8 * b8 + i, b8));
                const uint32_t w[4] = {rw.x, rw.y, rw.z, rw.w};
#pragma unroll
                for (int j = 0; j < i; ++j) {
                    const float2 f = unpack_f16(w[j >> 1]);
                    s = fmaf(-((j & 1) ? f.y : f.x), x[j], s);
                }
            }
            x[i] = s;
        }
        __syncwarp();                       // every row of the block has been read
#pragma unroll
        for (int i = 0; i < 8; ++i)         // column c of the block's inverse
            *reinterpret_cast<__half*>(hs + sw128_offset(8 * b8 + i, b8) + c * 2) = __float2half_rn(x[i]);
        __syncwarp();
    }
    {   // level 0.5: 8 -> 16.  The 16 x 16 block B now reads M = [X_aa 0; L_ba X_bb]; one A fragment of M serves both
        // products: M [0; L_ba] = [.; X_bb L_ba] and, with the lower accumulator rows re-used as A rows 8-15,
        // [.; Q 0] [X_aa; 0] = [.; Q X_aa].  X_ba = -(X_bb L_ba) X_aa goes back over L_ba.
        const int g = lane >> 2, t = lane & 3;
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int B = 2 * warp + u;
            uint32_t fm[4], bl, bx;
            frag_a(fm, h, B, B, lane);
            ldsm_x1_trans(bl, h + sw128_offset(16 * B + 8 + (lane & 7), 2 * B));
            ldsm_x1_trans(bx, h + sw128_offset(16 * B + (lane & 7), 2 * B));
            float q[4] = {0.f, 0.f, 0.f, 0.f}, r[4] = {0.f, 0.f, 0.f, 0.f};
            mma_f16(q, fm, 0u, bl);
            const uint32_t fq[4] = {0u, pack_f16(q[2], q[3]), 0u, 0u};
            mma_f16(r, fq, bx, 0u);
            *reinterpret_cast<uint32_t*>(hs + sw128_offset(16 * B + 8 + g, 2 * B) + 4 * t) = pack_f16(-r[2], -r[3]);
        }
        __syncwarp();
    }
    {   // X_ba = -(X_bb L_ba) X_aa
        const int a = 2 * warp, b = a + 1;
        uint32_t fa[4], fb[4], fp[4];
        float p[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}}, r[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        frag_a(fa, h, b, b, lane);
        frag_b(fb, h, b, a, lane);
        mma_block(p, fa, fb);
        frag_b(fb, h, a, a, lane);
        acc_to_a(fp, p);
        mma_block(r, fp, fb);
        store_neg_block(h, b, a, r, lane);
    }
}

// Level 2: warps 0..3 (after a barrier over them).  Warp w: output block (2 + (w & 1), w >> 1).  The output
// blocks overwrite the L blocks the other warps read, hence the 128-thread barrier `bar_id` before the store.
__device__ __forceinline__ void solve_level2(uint32_t h, int warp, int lane, int bar_id) {
    const int mt = warp & 1, nh = warp >> 1;
    float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
    for (int k2 = 0; k2 < 2; ++k2) {
        if (k2 >= nh) {                               // Xtl block (k2, nh) is zero above the diagonal
            float p[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
#pragma unroll
            for (int k1 = 0; k1 < 2; ++k1) {
                if (k1 <= mt) {                       // Xbr block (mt, k1) is zero above the diagonal
                    uint32_t fa[4], fb[4];
                    frag_a(fa, h, 2 + mt, 2 + k1, lane);
                    frag_b(fb, h, 2 + k1, k2, lane);
                    mma_block(p, fa, fb);
                }
            }
            uint32_t fp[4], fb[4];
            acc_to_a(fp, p);
            frag_b(fb, h, k2, nh, lane);
            mma_block(acc, fp, fb);
        }
    }
    asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
    store_neg_block(h, 2 + mt, nh, acc, lane);
}

}  // namespace tri
}  // namespace gdkvm
