// 3xTF32 tensor-core tile GEMM shared by the message-passing kernels (sde2d3d.cu, schnet.cu).
#pragma once

#include <cuda_fp16.h>

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

// ---------------------------------------------------------------------------------------
// The same tile GEMM on the f16 tensor-core shape (mma.m16n8k16, fp32 accumulate) with a two-way fp16 split of BOTH operands:
//   x = hi + lo,  hi = fp16(x) (11 significant bits),  lo = fp16(x - hi) (the next 11 bits),  c += a_lo*b_hi + a_hi*b_lo + a_hi*b_hi
// i.e. the same 22-bit operand coverage as the 3xTF32 split, but one instruction covers K = 16 instead of 8, and legacy
// mma.sync issues f16 k16 at the same rate as tf32 k8 (953 vs 476 MAC/clk/SM, profiles/r1_ubench_mma_rate.txt): half the tensor
// instructions for the same math.  fp16 has a narrow exponent: operands must stay below 65504 in magnitude (activations here are
// sin/cos values, LayerNorm outputs and O(1) hidden features; weights are O(1)), and a `lo` part below 2^-14 is kept with an
// ABSOLUTE precision of 2^-25 instead of a relative one -- an absolute error of <= 3e-8 |other operand| per product, far below
// the 1e-4 parity bar of the score network (checked against the reference's outputs by the same tests as before).
//   fragment layout (PTX ISA, mma.m16n8k16 .f16): g = lane/4, t = lane%4; every register holds two consecutive-k halves
//     A: {(g, 2t..2t+1)} {(g+8, 2t..)} {(g, 2t+8..)} {(g+8, 2t+8..)};  B: {(k=2t..2t+1, n=g)} {(k=2t+8.., n=g)};  C as above
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void split_f16x2(float e0, float e1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(e0, e1);            // .x (low 16 bits) = e0
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(e0 - hf.x, e1 - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int NB, int LDA_, int LDW_>
__device__ __forceinline__ void mma_gemm_h(const float* __restrict__ As, const float* __restrict__ Ws, int K, int lane,
                                           float (&c)[NB][4]) {
    static_assert(NB % 4 == 0, "n-blocks are processed four at a time");
    const int g = lane >> 2, t = lane & 3;
    const float* ap = As + (2 * t) * LDA_ + g;
    const float* wp = Ws + (2 * t) * LDW_ + g;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t ah[4], al[4];
        split_f16x2(ap[k0 * LDA_], ap[(k0 + 1) * LDA_], ah[0], al[0]);
        split_f16x2(ap[k0 * LDA_ + 8], ap[(k0 + 1) * LDA_ + 8], ah[1], al[1]);
        split_f16x2(ap[(k0 + 8) * LDA_], ap[(k0 + 9) * LDA_], ah[2], al[2]);
        split_f16x2(ap[(k0 + 8) * LDA_ + 8], ap[(k0 + 9) * LDA_ + 8], ah[3], al[3]);
#pragma unroll
        for (int nq = 0; nq < NB / 4; ++nq) {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float* w = wp + k0 * LDW_ + (nq * 4 + q) * 8;
                split_f16x2(w[0], w[LDW_], bh[q][0], bl[q][0]);
                split_f16x2(w[8 * LDW_], w[9 * LDW_], bh[q][1], bl[q][1]);
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], al, bh[q][0], bh[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], ah, bl[q][0], bl[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], ah, bh[q][0], bh[q][1]);
        }
    }
}

// mma_gemm_h with the B operand PRE-SPLIT on the host (moleculesde_b200/sde_2d_to_3d.py: pack_f16_pairs): the weight block keeps
// its k-major [K][LDW_] word layout, but row 2p holds half2(hi[2p], hi[2p+1]) and row 2p+1 holds half2(lo[2p], lo[2p+1]) of a
// column -- the B fragments are plain 32-bit loads from the same four addresses, with no conversion in the inner loop.
template <int NB, int LDA_, int LDW_>
__device__ __forceinline__ void mma_gemm_hp(const float* __restrict__ As, const float* __restrict__ Wp, int K, int lane,
                                            float (&c)[NB][4]) {
    static_assert(NB % 4 == 0, "n-blocks are processed four at a time");
    const int g = lane >> 2, t = lane & 3;
    const float* ap = As + (2 * t) * LDA_ + g;
    const uint32_t* wp = reinterpret_cast<const uint32_t*>(Wp) + (2 * t) * LDW_ + g;
#pragma unroll 1
    for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t ah[4], al[4];
        split_f16x2(ap[k0 * LDA_], ap[(k0 + 1) * LDA_], ah[0], al[0]);
        split_f16x2(ap[k0 * LDA_ + 8], ap[(k0 + 1) * LDA_ + 8], ah[1], al[1]);
        split_f16x2(ap[(k0 + 8) * LDA_], ap[(k0 + 9) * LDA_], ah[2], al[2]);
        split_f16x2(ap[(k0 + 8) * LDA_ + 8], ap[(k0 + 9) * LDA_ + 8], ah[3], al[3]);
#pragma unroll
        for (int nq = 0; nq < NB / 4; ++nq) {
            uint32_t bh[4][2], bl[4][2];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t* w = wp + k0 * LDW_ + (nq * 4 + q) * 8;
                bh[q][0] = w[0];
                bl[q][0] = w[LDW_];
                bh[q][1] = w[8 * LDW_];
                bl[q][1] = w[9 * LDW_];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], al, bh[q][0], bh[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], ah, bl[q][0], bl[q][1]);
#pragma unroll
            for (int q = 0; q < 4; ++q) mma_f16(c[nq * 4 + q], ah, bh[q][0], bh[q][1]);
        }
    }
}

template <int NB>
__device__ __forceinline__ void zero_frag(float (&c)[NB][4]) {
#pragma unroll
    for (int i = 0; i < NB; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
}

}  // namespace molsde
