// Fused row-wise 3-layer MLP   y = W2 . act(W1 . act(W0 . x + b0) + b1) + b2   for very tall, very narrow inputs: the per-pair
// MLPs of the dense 3D->2D edge network (edge_network_dense.py:120-123: 2C -> h -> h -> C', elu, B*Nm^2 rows) and its final
// head (invariant_scorenetwork_dense.py:84-86: 30 -> 60 -> 60 -> 1, silu).  Unfused, every layer streams ~60 floats per row
// through HBM; here a warp keeps a 16-row stripe in shared memory through all three layers (weights staged once per CTA) and
// runs the two wide layers on the tensor cores (mma.sync m16n8k8, 3xTF32 split = fp32-class accuracy), the last, <= 8-wide
// layer with FFMA.  Rows are independent: deterministic by construction.
#include "mma_tile.cuh"

namespace molsde {

constexpr int MR_WARPS = 8, MR_LDA = 24;  // stripe: [k][24] floats, rows 0..15 (24 = conflict-free for the A fragment loads)

// Activations on the SFU fast path (the epilogues evaluate 64 of them per lane and stripe, which would otherwise outweigh the
// MMAs): __expf / __fdividef are accurate to ~2 ulp; expm1 switches to its 4-term series for |v| < 0.1 so that small
// arguments keep full relative accuracy.  Parity with the reference stays within the 1e-4 bar (tests/test_gpu_dense.py).
__device__ __forceinline__ float mr_act(float v, int act) {
    switch (act) {
        case 2: return __fdividef(v, 1.0f + __expf(-v));
        case 4: { const float e = __expf(-2.0f * fabsf(v)); return copysignf(__fdividef(1.0f - e, 1.0f + e), v); }
        case 5: {
            const float series = v * fmaf(v, fmaf(v, fmaf(v, 1.0f / 24.0f, 1.0f / 6.0f), 0.5f), 1.0f);
            const float em1 = v > -0.1f ? series : __expf(v) - 1.0f;
            return v > 0.0f ? v : em1;
        }
        default: return v;
    }
}

template <int NB>
__global__ void __launch_bounds__(MR_WARPS * 32)
mlp3_rows_kernel(const float* __restrict__ X, int64_t rows, int64_t ldx, int K0, const float* __restrict__ W0, const float* __restrict__ b0,
                 int H1, const float* __restrict__ W1, const float* __restrict__ b1, int H2, const float* __restrict__ W2,
                 const float* __restrict__ b2, int NO, int act, float* __restrict__ Y, int64_t ldy) {
    extern __shared__ __align__(16) float mr_smem[];
    constexpr int HP = NB * 8, LDW = HP + 8;          // padded hidden width, weight leading dimension (== 8 mod 32)
    const int K0p = (K0 + 7) & ~7;
    float* W0s = mr_smem;                              // [K0p][LDW]  k-major
    float* W1s = W0s + 32 * LDW;                       // [HP][LDW]
    float* W2s = W1s + HP * LDW;                       // [8][HP]     (row o, k contiguous; rows >= NO zero)
    float* bs = W2s + 8 * HP;                          // b0 [HP] | b1 [HP] | b2 [8]
    float* stripes = bs + 2 * HP + 8;                  // per warp: two buffers [HP][MR_LDA]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * LDW; i += blockDim.x) {
        const int k = i / LDW, n = i % LDW;
        W0s[i] = (k < K0 && n < H1) ? W0[static_cast<int64_t>(n) * K0 + k] : 0.0f;
    }
    for (int i = tid; i < HP * LDW; i += blockDim.x) {
        const int k = i / LDW, n = i % LDW;
        W1s[i] = (k < H1 && n < H2) ? W1[static_cast<int64_t>(n) * H1 + k] : 0.0f;
    }
    for (int i = tid; i < 8 * HP; i += blockDim.x) {
        const int o = i / HP, k = i % HP;
        W2s[i] = (o < NO && k < H2) ? W2[static_cast<int64_t>(o) * H2 + k] : 0.0f;
    }
    for (int i = tid; i < 2 * HP + 8; i += blockDim.x)
        bs[i] = i < HP ? (i < H1 ? b0[i] : 0.0f) : i < 2 * HP ? (i - HP < H2 ? b1[i - HP] : 0.0f) : (i - 2 * HP < NO ? b2[i - 2 * HP] : 0.0f);
    __syncthreads();
    float* A = stripes + warp * (2 * HP * MR_LDA);
    float* Bf = A + HP * MR_LDA;
    const int g = lane >> 2, t = lane & 3;
    const int64_t nstripes = (rows + 15) / 16;
    for (int64_t sidx = static_cast<int64_t>(blockIdx.x) * MR_WARPS + warp; sidx < nstripes; sidx += static_cast<int64_t>(gridDim.x) * MR_WARPS) {
        const int64_t r0 = sidx * 16;
        // stage the stripe transposed: A[k][row]  (lanes along k: coalesced row reads; all 16 loads in flight before the stores)
        {
            float v[16];
#pragma unroll
            for (int r = 0; r < 16; ++r) {
                const int64_t gr = r0 + r;
                v[r] = (gr < rows && lane < K0) ? __ldg(X + gr * ldx + lane) : 0.0f;
            }
            if (lane < K0p) {
#pragma unroll
                for (int r = 0; r < 16; r += 4) *reinterpret_cast<float4*>(A + lane * MR_LDA + r) = make_float4(v[r], v[r + 1], v[r + 2], v[r + 3]);
            }
        }
        __syncwarp();
        float acc[NB][4];
        zero_frag(acc);
        mma_gemm<NB, MR_LDA, LDW>(A, W0s, K0p, lane, acc);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const int c0 = nb * 8 + 2 * t;
            Bf[c0 * MR_LDA + g] = mr_act(acc[nb][0] + bs[c0], act);
            Bf[(c0 + 1) * MR_LDA + g] = mr_act(acc[nb][1] + bs[c0 + 1], act);
            Bf[c0 * MR_LDA + g + 8] = mr_act(acc[nb][2] + bs[c0], act);
            Bf[(c0 + 1) * MR_LDA + g + 8] = mr_act(acc[nb][3] + bs[c0 + 1], act);
        }
        __syncwarp();
        zero_frag(acc);
        mma_gemm<NB, MR_LDA, LDW>(Bf, W1s, HP, lane, acc);
#pragma unroll
        for (int nb = 0; nb < NB; ++nb) {
            const int c0 = nb * 8 + 2 * t;
            A[c0 * MR_LDA + g] = mr_act(acc[nb][0] + bs[HP + c0], act);
            A[(c0 + 1) * MR_LDA + g] = mr_act(acc[nb][1] + bs[HP + c0 + 1], act);
            A[c0 * MR_LDA + g + 8] = mr_act(acc[nb][2] + bs[HP + c0], act);
            A[(c0 + 1) * MR_LDA + g + 8] = mr_act(acc[nb][3] + bs[HP + c0 + 1], act);
        }
        __syncwarp();
        // last layer (<= 8 outputs): lane -> (row = lane % 16, output parity lane / 16)
        {
            const int r = lane & 15, half = lane >> 4;
            const int64_t gr = r0 + r;
            for (int o = half; o < NO; o += 2) {
                float s = 0.0f;
                for (int k = 0; k < HP; ++k) s = fmaf(A[k * MR_LDA + r], W2s[o * HP + k], s);
                if (gr < rows) Y[gr * ldy + o] = s + bs[2 * HP + o];
            }
        }
        __syncwarp();
    }
}

}  // namespace molsde

using namespace molsde;

template <int NB>
static int mr_launch(const float* X, int64_t rows, int64_t ldx, int K0, const float* W0, const float* b0, int H1, const float* W1,
                     const float* b1, int H2, const float* W2, const float* b2, int NO, int act, float* Y, int64_t ldy, cudaStream_t s) {
    constexpr int HP = NB * 8, LDW = HP + 8;
    constexpr size_t smem = sizeof(float) * (32 * LDW + HP * LDW + 8 * HP + 2 * HP + 8 + MR_WARPS * 2 * HP * MR_LDA);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(mlp3_rows_kernel<NB>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MOLSDE_ERR_CUDA; }
        configured = true;
    }
    const int64_t nstripes = (rows + 15) / 16;
    int64_t ctas = (nstripes + MR_WARPS - 1) / MR_WARPS;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * (NB == 4 ? 3 : 1) * 2;
    if (ctas > cap) ctas = cap;
    mlp3_rows_kernel<NB><<<static_cast<unsigned>(ctas), MR_WARPS * 32, smem, s>>>(X, rows, ldx, K0, W0, b0, H1, W1, b1, H2, W2, b2, NO, act,
                                                                                Y, ldy);
    return check_launch("mlp3_rows");
}

extern "C" int molsde_mlp3_rows(const float* X, int64_t rows, int64_t ldx, int32_t K0, const float* W0, const float* b0, int32_t H1,
                                const float* W1, const float* b1, int32_t H2, const float* W2, const float* b2, int32_t NO, int32_t act,
                                float* Y, int64_t ldy, void* stream) {
    if (!X || !W0 || !b0 || !W1 || !b1 || !W2 || !b2 || !Y || rows < 0) return MOLSDE_ERR_INVALID;
    if (K0 < 1 || K0 > 32 || H1 < 1 || H1 > 64 || H2 < 1 || H2 > 64 || NO < 1 || NO > 8 || (act != 2 && act != 4 && act != 5))
        return MOLSDE_ERR_UNSUPPORTED;
    if (rows == 0) return MOLSDE_OK;
    cudaStream_t s = as_stream(stream);
    if (H1 <= 32 && H2 <= 32) return mr_launch<4>(X, rows, ldx, K0, W0, b0, H1, W1, b1, H2, W2, b2, NO, act, Y, ldy, s);
    return mr_launch<8>(X, rows, ldx, K0, W0, b0, H1, W1, b1, H2, W2, b2, NO, act, Y, ldy, s);
}
