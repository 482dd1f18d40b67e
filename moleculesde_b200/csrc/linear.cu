// Node-level dense layer  Y = act(X . W^T + b), fp32 FFMA register-tiled SGEMM.
// Used for the loop-invariant node projections of the 2D->3D model (node_emb, the node-factored
// first layer of edge_2D_emb; SDE_model_2D_to_3D.py:264-265) and SchNet's node linears.
// Tile 128x64x16, 256 threads, 8x4 micro-tile; M, N, K arbitrary (guards on the edges).
#include "common.cuh"

namespace molsde {

constexpr int BM = 128, BN = 64, BK = 16, LDA_S = BM + 4, LDB_S = BN + 4;

__device__ __forceinline__ float apply_act(float v, int act) {
    switch (act) {
        case 1: return fmaxf(v, 0.0f);
        case 2: return silu_f(v);
        case 3: return softplus_f(v) - 0.69314718246459961f;  // float32(log 2), schnet.py:213
        case 4: return tanhf(v);
        case 5: return v > 0.0f ? v : expm1f(v);  // F.elu
        default: return v;
    }
}

// One K block (16 deep) of a row-major [rows x K] operand -> registers: thread -> (row = item / 4, k-quad = item % 4), one float4
// per item when the row is 16-byte aligned, guarded scalars otherwise; zero beyond the edges.
template <int ROWS>
__device__ __forceinline__ void ld_block(float4 (&r)[ROWS * 4 / 256], const float* __restrict__ P, int64_t row0, int64_t nrows, int64_t ld,
                                         int k0, int K, bool vec) {
#pragma unroll
    for (int i = 0; i < ROWS * 4 / 256; ++i) {
        const int item = threadIdx.x + i * 256, rr = item >> 2, kq = item & 3;
        const int64_t g = row0 + rr;
        const int k = k0 + 4 * kq;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (g < nrows) {
            const float* p = P + g * ld + k;
            if (vec && k + 3 < K) {
                v = __ldg(reinterpret_cast<const float4*>(p));
            } else {
                if (k < K) v.x = __ldg(p);
                if (k + 1 < K) v.y = __ldg(p + 1);
                if (k + 2 < K) v.z = __ldg(p + 2);
                if (k + 3 < K) v.w = __ldg(p + 3);
            }
        }
        r[i] = v;
    }
}
template <int ROWS, int LDS_>
__device__ __forceinline__ void st_block(const float4 (&r)[ROWS * 4 / 256], float* __restrict__ S) {   // k-major: S[k][row]
#pragma unroll
    for (int i = 0; i < ROWS * 4 / 256; ++i) {
        const int item = threadIdx.x + i * 256, rr = item >> 2, kq = item & 3;
        S[(4 * kq) * LDS_ + rr] = r[i].x;
        S[(4 * kq + 1) * LDS_ + rr] = r[i].y;
        S[(4 * kq + 2) * LDS_ + rr] = r[i].z;
        S[(4 * kq + 3) * LDS_ + rr] = r[i].w;
    }
}

// Tile 128 x 64 x 16, 256 threads, 8 x 4 micro-tile, register-prefetched double buffering.  NUMERICS CONTRACT (the reason this kernel
// exists beside the tensor-core GEMMs): every output is ONE fp32 accumulator starting at 0 and updated by fmaf over k = 0 .. K-1 in
// ascending order, then + bias -- linears in front of BatchNorm + ReLU use it in training, where sign decisions near 0 feed
// gradients.  The tiling does not change a single bit of the result (the 64 x 64 / 4 x 4 kernel it replaces computed
// the same chains); it only raises the rate: 8 x 4 outputs per thread = 10.7 FMA per shared-memory load instead of 8, coalesced
// float4 global loads, the next block's loads in flight during the current block's FMAs.
__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int64_t M, int K, int64_t ldx, const float* __restrict__ W,
              const float* __restrict__ bias, int N, float* __restrict__ Y, int64_t ldy, int act,
              const float* __restrict__ R, int64_t ldr, const float* __restrict__ rowscale) {
    __shared__ __align__(16) float Xs[2][BK * LDA_S];
    __shared__ __align__(16) float Ws[2][BK * LDB_S];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> 4 output cols; ty -> rows ty*4 .. +3 and 64 + ty*4 .. +3
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
    const int n0 = blockIdx.x * BN;
    const bool vx = (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (ldx & 3) == 0;
    const bool vw = (reinterpret_cast<uintptr_t>(W) & 15) == 0 && (K & 3) == 0;
    float acc[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    float4 ra[BM * 4 / 256], rb[BN * 4 / 256];
    ld_block<BM>(ra, X, m0, M, ldx, 0, K, vx);
    ld_block<BN>(rb, W, n0, N, K, 0, K, vw);
    st_block<BM, LDA_S>(ra, Xs[0]);
    st_block<BN, LDB_S>(rb, Ws[0]);
    __syncthreads();
    const int nkb = (K + BK - 1) / BK;
    for (int kb = 0; kb < nkb; ++kb) {
        const int cur = kb & 1;
        if (kb + 1 < nkb) {
            ld_block<BM>(ra, X, m0, M, ldx, (kb + 1) * BK, K, vx);
            ld_block<BN>(rb, W, n0, N, K, (kb + 1) * BK, K, vw);
        }
        const float* xs = Xs[cur];
        const float* ws = Ws[cur];
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(xs + k * LDA_S + ty * 4);
            const float4 a1 = *reinterpret_cast<const float4*>(xs + k * LDA_S + 64 + ty * 4);
            const float4 b = *reinterpret_cast<const float4*>(ws + k * LDB_S + tx * 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kb + 1 < nkb) {
            st_block<BM, LDA_S>(ra, Xs[cur ^ 1]);
            st_block<BN, LDB_S>(rb, Ws[cur ^ 1]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t gm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + i - 4);
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j] + (bias ? bias[gn] : 0.0f);
            if (rowscale) v *= rowscale[gm];  // mask_x before the activation (edge_network_dense.py:117-118)
            v = apply_act(v, act);
            if (R) v += R[gm * ldr + gn];  // residual (h = h + interaction(h), schnet.py:97)
            Y[gm * ldy + gn] = v;
        }
    }
}

}  // namespace molsde

using namespace molsde;

extern "C" int molsde_linear(const float* X, int64_t M, int32_t K, int64_t ldx, const float* W, const float* b,
                             int32_t N, float* Y, int64_t ldy, int32_t act, const float* R, int64_t ldr, const float* rowscale,
                             void* stream) {
    if (!X || !W || !Y || M < 0 || K <= 0 || N <= 0) return MOLSDE_ERR_INVALID;
    if (M == 0) return MOLSDE_OK;
    dim3 grid((N + BN - 1) / BN, static_cast<unsigned>((M + BM - 1) / BM));
    linear_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, M, K, ldx, W, b, N, Y, ldy, act, R, ldr, rowscale);
    return check_launch("linear");
}
