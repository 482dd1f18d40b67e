// Node-level dense layer  Y = act(X . W^T + b), fp32 FFMA register-tiled SGEMM.
// Used for the loop-invariant node projections of the 2D->3D model (node_emb, the node-factored
// first layer of edge_2D_emb; SDE_model_2D_to_3D.py:264-265) and SchNet's node linears.
// Tile 64x64x16, 256 threads, 4x4 micro-tile; M, N, K arbitrary (guards on the edges).
#include "common.cuh"

namespace molsde {

constexpr int BM = 64, BN = 64, BK = 16;

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

__global__ void __launch_bounds__(256)
linear_kernel(const float* __restrict__ X, int64_t M, int K, int64_t ldx, const float* __restrict__ W,
              const float* __restrict__ bias, int N, float* __restrict__ Y, int64_t ldy, int act,
              const float* __restrict__ R, int64_t ldr, const float* __restrict__ rowscale) {
    __shared__ float Xs[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;  // tx -> 4 output cols, ty -> 4 rows
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * BM;
    const int n0 = blockIdx.x * BN;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
    for (int k0 = 0; k0 < K; k0 += BK) {
        // load X tile [BM][BK] and W tile [BN][BK], transposed into k-major smem
        for (int idx = tid; idx < BM * BK; idx += 256) {
            int r = idx / BK, c = idx % BK;
            int64_t gm = m0 + r;
            int gk = k0 + c;
            Xs[c][r] = (gm < M && gk < K) ? X[gm * ldx + gk] : 0.0f;
        }
        for (int idx = tid; idx < BN * BK; idx += 256) {
            int r = idx / BK, c = idx % BK;
            int gn = n0 + r, gk = k0 + c;
            Ws[c][r] = (gn < N && gk < K) ? W[static_cast<int64_t>(gn) * K + gk] : 0.0f;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            int gn = n0 + tx * 4 + j;
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
