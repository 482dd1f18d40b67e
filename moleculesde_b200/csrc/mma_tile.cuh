// 3xTF32 tensor-core tile GEMM shared by the message-passing kernels (sde2d3d.cu, schnet.cu).
#pragma once

#include "common.cuh"

namespace molsde {

// ---------------------------------------------------------------------------------------
// tensor-core tile GEMM: c[nb] (16 rows x 8 cols per n-block) += A[16 x K] . W[K x 8*NB]
//   As: k-major, As[k*LDA_ + row], pointing at the warp's first row;  Ws: Ws[k*LDW_ + col],
//   pointing at the warp's first column.  3xTF32 split, small terms first.
//   fragment layout (PTX ISA, mma.m16n8k8 .tf32): g = lane/4, t = lane%4
//     a0 (g, t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4);  b0 (k=t, n=g) b1 (k=t+4, n=g)
//     c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
// ---------------------------------------------------------------------------------------
// hi part of the 3xTF32 split: the top 19 bits of the fp32 pattern (what the tensor core keeps of a .tf32 operand).
// `cvt.rna.tf32.f32` is not native on sm_100a (ptxas expands it to FSETP+IADD3+SEL+LOP3, profiles/r1_pc_v2_lines.txt),
// so the split truncates with one LOP3; lo = x - hi is exact, and the dropped lo*lo term plus the tensor core's own
// truncation of lo stay below 2^-20 relative.
__device__ __forceinline__ uint32_t tf32_hi(float x) { return __float_as_uint(x) & 0xffffe000u; }
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NB, int LDA_, int LDW_>
__device__ __forceinline__ void mma_gemm(const float* __restrict__ As, const float* __restrict__ Ws, int K, int lane,
                                         float (&c)[NB][4]) {
    static_assert(NB % 4 == 0, "n-blocks are processed four at a time");
    const int g = lane >> 2, t = lane & 3;
    const float* ap = As + t * LDA_ + g;
    const float* wp = Ws + t * LDW_ + g;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 8) {
        const float av[4] = {ap[k0 * LDA_], ap[k0 * LDA_ + 8], ap[(k0 + 4) * LDA_], ap[(k0 + 4) * LDA_ + 8]};
        uint32_t ah[4], al[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            ah[i] = tf32_hi(av[i]);
            al[i] = __float_as_uint(av[i] - __uint_as_float(ah[i]));
        }
#pragma unroll
        for (int nq = 0; nq < NB / 4; ++nq) {
            // four independent accumulators per pass: the three split terms of one accumulator are 4 MMAs apart
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float b0 = wp[k0 * LDW_ + (nq * 4 + q) * 8], b1 = wp[(k0 + 4) * LDW_ + (nq * 4 + q) * 8];
                bh[q][0] = tf32_hi(b0);
                bh[q][1] = tf32_hi(b1);
                bl[q][0] = __float_as_uint(b0 - __uint_as_float(bh[q][0]));
                bl[q][1] = __float_as_uint(b1 - __uint_as_float(bh[q][1]));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(c[nq * 4 + q], al, bh[q][0], bh[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(c[nq * 4 + q], ah, bl[q][0], bl[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_tf32(c[nq * 4 + q], ah, bh[q][0], bh[q][1]);
        }
    }
}

template <int NB>
__device__ __forceinline__ void zero_frag(float (&c)[NB][4]) {
#pragma unroll
    for (int i = 0; i < NB; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
}

}  // namespace molsde
