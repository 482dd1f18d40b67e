// Whole-chain kernels for the NARROW 3-layer MLPs of the training step (EdgeNetwork_dense.mlp: 2C -> 2max(C,C') -> same -> C',
// edge_network_dense.py:120-123, applied to the B*Nm^2 atom pairs of a batch).  The layer-granular form was 3 GEMM + 2 activation launches forward and
// 3 x (dW, dx) GEMMs + 2 activation launches backward per MLP, every one of them latency-bound at 16-60 columns.
//
//   forward : thread = pair row; the three layers run back to back with the hidden vector in REGISTERS, weights are read from
//             shared memory as broadcast float4 (one LDS.128 per four FFMA); the pre-activations of both hidden layers are stored
//             for the backward.  fp32 FFMA with k-ascending fmaf chains (tighter than the 3xTF32 GEMMs it replaces).
//   backward: thread = pair row; the whole input-gradient chain  dy -> d2 = (W3^T dy) act'(p2) -> d1 = (W2^T d2) act'(p1) ->
//             dx = W1^T d1  in registers.  It also writes a2 = act(p2), d2, a1 = act(p1), d1: the operands of the three weight-gradient
//             GEMMs (dW_l = d_l^T a_{l-1}, bias gradient through the all-ones row), which are LEAVES of the dependency graph and run
//             on the tape's side stream as before.
// Dimensions are template parameters (the hidden vectors must live in registers); `molsde_mlp3_train_supported` tells the host
// which (d0, h, d3, act) combinations exist -- everything else keeps the layer-granular path.
#include "common.cuh"

namespace molsde {

__host__ __device__ constexpr int mp4(int n) { return (n + 3) / 4 * 4; }

template <int ACT>
__device__ __forceinline__ float m3_act(float v) {
    if (ACT == 2) return silu_f(v);
    if (ACT == 5) return v > 0.0f ? v : expm1f(v);
    return v;
}
// activation value and derivative from the pre-activation
template <int ACT>
__device__ __forceinline__ void m3_act_d(float p, float& a, float& d) {
    if (ACT == 2) {
        const float e = 1.0f + expf(-p);
        const float s = 1.0f / e;
        a = p / e;
        d = s * (1.0f + p * (1.0f - s));
    } else if (ACT == 5) {
        if (p > 0.0f) { a = p; d = 1.0f; } else { a = expm1f(p); d = expf(p); }
    } else {
        a = p; d = 1.0f;
    }
}

// W [R][C] (row-major, nn.Linear layout) -> shared [R][CP] with zero padding columns
template <int R, int C, int CP>
__device__ __forceinline__ void m3_stage(float* dst, const float* __restrict__ W) {
    for (int i = threadIdx.x; i < R * CP; i += blockDim.x) {
        const int r = i / CP, c = i % CP;
        dst[i] = c < C ? W[r * C + c] : 0.0f;
    }
}

// ---- row tiles through shared memory -------------------------------------------------------------------------------------
// A CTA owns M3_ROWS consecutive rows, thread = row.  A thread walking its own row in global memory touches a different 128-byte
// line per lane (32 wavefronts per warp access, the lines evicted from L1 before their other sectors are used): every [rows, C]
// array therefore moves between global and shared memory COOPERATIVELY (consecutive threads = consecutive 16 bytes of the
// contiguous row block) and the thread reads / writes its row in shared memory, row stride C + 4 floats (conflict-free LDS.128).
constexpr int M3_ROWS = 128;

template <int C>   // global [nrows][C] (contiguous rows, C % 4 == 0, 16-byte aligned base) -> tile [M3_ROWS][C + 4]
__device__ __forceinline__ void m3_tile_load(float* tile, const float* __restrict__ g, int nrows) {
    constexpr int C4 = C / 4, S = C + 4;
    for (int e = threadIdx.x; e < nrows * C4; e += M3_ROWS) {
        const int r = e / C4, c = e % C4;
        *reinterpret_cast<float4*>(tile + r * S + 4 * c) = *reinterpret_cast<const float4*>(g + static_cast<int64_t>(r) * C + 4 * c);
    }
}
template <int C>
__device__ __forceinline__ void m3_tile_store(const float* tile, float* __restrict__ g, int nrows) {
    constexpr int C4 = C / 4, S = C + 4;
    for (int e = threadIdx.x; e < nrows * C4; e += M3_ROWS) {
        const int r = e / C4, c = e % C4;
        *reinterpret_cast<float4*>(g + static_cast<int64_t>(r) * C + 4 * c) = *reinterpret_cast<const float4*>(tile + r * S + 4 * c);
    }
}
// the same for an arbitrary width / row stride (the MLP input x and its gradient): scalar, lanes along the row
template <int C, int CP>
__device__ __forceinline__ void m3_tile_load_any(float* tile, const float* __restrict__ g, int64_t ld, int nrows) {
    constexpr int S = CP + 4;
    for (int e = threadIdx.x; e < nrows * CP; e += M3_ROWS) {
        const int r = e / CP, c = e % CP;
        tile[r * S + c] = c < C ? g[static_cast<int64_t>(r) * ld + c] : 0.0f;
    }
}
template <int C, int CP>
__device__ __forceinline__ void m3_tile_store_any(const float* tile, float* __restrict__ g, int64_t ld, int nrows) {
    constexpr int S = CP + 4;
    for (int e = threadIdx.x; e < nrows * C; e += M3_ROWS) {
        const int r = e / C, c = e % C;
        g[static_cast<int64_t>(r) * ld + c] = tile[r * S + c];
    }
}

// out[i] = bias[i] + sum_k W[i][k] v[k], i < R: four outputs at a time (four independent k-ascending fmaf chains in flight)
template <int R, int KP>
__device__ __forceinline__ void m3_matvec(float (&out)[R], const float* __restrict__ Ws, const float* __restrict__ bs, const float (&v)[KP]) {
    static_assert(R % 4 == 0, "");
#pragma unroll
    for (int i = 0; i < R; i += 4) {
        float a0 = bs[i], a1 = bs[i + 1], a2 = bs[i + 2], a3 = bs[i + 3];
#pragma unroll
        for (int k = 0; k < KP; k += 4) {
            const float4 w0 = *reinterpret_cast<const float4*>(Ws + (i + 0) * KP + k);
            const float4 w1 = *reinterpret_cast<const float4*>(Ws + (i + 1) * KP + k);
            const float4 w2 = *reinterpret_cast<const float4*>(Ws + (i + 2) * KP + k);
            const float4 w3 = *reinterpret_cast<const float4*>(Ws + (i + 3) * KP + k);
            a0 = fmaf(w0.x, v[k], a0); a1 = fmaf(w1.x, v[k], a1); a2 = fmaf(w2.x, v[k], a2); a3 = fmaf(w3.x, v[k], a3);
            a0 = fmaf(w0.y, v[k + 1], a0); a1 = fmaf(w1.y, v[k + 1], a1); a2 = fmaf(w2.y, v[k + 1], a2); a3 = fmaf(w3.y, v[k + 1], a3);
            a0 = fmaf(w0.z, v[k + 2], a0); a1 = fmaf(w1.z, v[k + 2], a1); a2 = fmaf(w2.z, v[k + 2], a2); a3 = fmaf(w3.z, v[k + 2], a3);
            a0 = fmaf(w0.w, v[k + 3], a0); a1 = fmaf(w1.w, v[k + 3], a1); a2 = fmaf(w2.w, v[k + 3], a2); a3 = fmaf(w3.w, v[k + 3], a3);
        }
        out[i] = a0; out[i + 1] = a1; out[i + 2] = a2; out[i + 3] = a3;
    }
}

template <int D0, int H, int D3>
struct M3Smem {   // dynamic shared memory layout (floats)
    static constexpr int D0P = mp4(D0), HP = mp4(H);
    static constexpr int W1 = 0, W2 = W1 + H * D0P, W3 = W2 + H * HP, B1 = W3 + mp4(D3) * HP, B2 = B1 + HP, B3 = B2 + HP,
                         TILE = B3 + mp4(D3), TILE_FLOATS = M3_ROWS * ((HP > D0P ? HP : D0P) + 4), TOTAL = TILE + TILE_FLOATS;
};

template <int D0, int H, int D3, int ACT>
__global__ void __launch_bounds__(M3_ROWS, 3) mlp3_train_fwd_kernel(const float* __restrict__ x, int64_t rows, int64_t ldx,
                                                                   const float* __restrict__ W1, const float* __restrict__ b1,
                                                                   const float* __restrict__ W2, const float* __restrict__ b2,
                                                                   const float* __restrict__ W3, const float* __restrict__ b3,
                                                                   float* __restrict__ p1, float* __restrict__ p2, float* __restrict__ y) {
    using L = M3Smem<D0, H, D3>;
    constexpr int D0P = L::D0P, HP = L::HP;
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4 (float4 row tiles)");
    extern __shared__ __align__(16) float m3_smem[];
    float *W1s = m3_smem + L::W1, *W2s = m3_smem + L::W2, *W3s = m3_smem + L::W3, *b1s = m3_smem + L::B1, *b2s = m3_smem + L::B2,
          *b3s = m3_smem + L::B3, *tile = m3_smem + L::TILE;
    m3_stage<H, D0, D0P>(W1s, W1);
    m3_stage<H, H, HP>(W2s, W2);
    m3_stage<D3, H, HP>(W3s, W3);
    for (int i = threadIdx.x; i < H; i += M3_ROWS) { b1s[i] = b1[i]; b2s[i] = b2[i]; }
    for (int i = threadIdx.x; i < D3; i += M3_ROWS) b3s[i] = b3[i];
    const int t = threadIdx.x;
    for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * M3_ROWS; r0 < rows; r0 += static_cast<int64_t>(gridDim.x) * M3_ROWS) {
        const int nrows = static_cast<int>(min(static_cast<int64_t>(M3_ROWS), rows - r0));
        __syncthreads();                                   // weights staged / previous tile's stores done
        m3_tile_load_any<D0, D0P>(tile, x + r0 * ldx, ldx, nrows);
        __syncthreads();
        float xv[D0P];
#pragma unroll
        for (int k = 0; k < D0P; k += 4) {
            const float4 v = *reinterpret_cast<const float4*>(tile + t * (D0P + 4) + k);
            xv[k] = v.x; xv[k + 1] = v.y; xv[k + 2] = v.z; xv[k + 3] = v.w;
        }
        float h[H];
        m3_matvec<H, D0P>(h, W1s, b1s, xv);
        __syncthreads();                                   // every thread has read its x row
#pragma unroll
        for (int i = 0; i < H; i += 4) *reinterpret_cast<float4*>(tile + t * (H + 4) + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
        __syncthreads();
        m3_tile_store<H>(tile, p1 + r0 * H, nrows);
#pragma unroll
        for (int i = 0; i < H; ++i) h[i] = m3_act<ACT>(h[i]);
        float g[H];
        m3_matvec<H, HP>(g, W2s, b2s, h);
        __syncthreads();                                   // p1 tile written out
#pragma unroll
        for (int i = 0; i < H; i += 4) *reinterpret_cast<float4*>(tile + t * (H + 4) + i) = make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]);
        __syncthreads();
        m3_tile_store<H>(tile, p2 + r0 * H, nrows);
#pragma unroll
        for (int i = 0; i < H; ++i) g[i] = m3_act<ACT>(g[i]);
        float out[D3];
#pragma unroll
        for (int o = 0; o < D3; ++o) {
            float acc = b3s[o];
#pragma unroll
            for (int k = 0; k < HP; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W3s + o * HP + k);
                acc = fmaf(w.x, g[k], acc); acc = fmaf(w.y, g[k + 1], acc); acc = fmaf(w.z, g[k + 2], acc); acc = fmaf(w.w, g[k + 3], acc);
            }
            out[o] = acc;
        }
        if (t < nrows) {
            float* yo = y + (r0 + t) * D3;
            if (D3 % 4 == 0) {   // (y is a fresh [rows, D3] allocation: 16-byte aligned rows)
#pragma unroll
                for (int o = 0; o + 3 < D3; o += 4) *reinterpret_cast<float4*>(yo + o) = make_float4(out[o], out[o + 1], out[o + 2], out[o + 3]);
            } else {
#pragma unroll
                for (int o = 0; o < D3; ++o) yo[o] = out[o];
            }
        }
    }
}

template <int D0, int H, int D3, int ACT>
__global__ void __launch_bounds__(M3_ROWS, 3) mlp3_train_bwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                                                   const float* __restrict__ dy, int64_t rows,
                                                                   const float* __restrict__ W1, const float* __restrict__ W2,
                                                                   const float* __restrict__ W3, float* __restrict__ a1, float* __restrict__ a2,
                                                                   float* __restrict__ d1, float* __restrict__ d2, float* __restrict__ dx,
                                                                   int64_t lddx) {
    using L = M3Smem<D0, H, D3>;
    constexpr int D0P = L::D0P, HP = L::HP, S = H + 4;
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4 (float4 row tiles)");
    extern __shared__ __align__(16) float m3_smem[];
    float *W1s = m3_smem + L::W1, *W2s = m3_smem + L::W2, *W3s = m3_smem + L::W3, *tile = m3_smem + L::TILE;
    m3_stage<H, D0, D0P>(W1s, W1);
    m3_stage<H, H, HP>(W2s, W2);
    m3_stage<D3, H, HP>(W3s, W3);
    const int t = threadIdx.x;
    for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * M3_ROWS; r0 < rows; r0 += static_cast<int64_t>(gridDim.x) * M3_ROWS) {
        const int nrows = static_cast<int>(min(static_cast<int64_t>(M3_ROWS), rows - r0));
        const bool live = t < nrows;
        __syncthreads();                                   // weights staged / previous tile's stores done
        m3_tile_load<H>(tile, p2 + r0 * H, nrows);
        // v = W3^T dy
        float v[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) v[j] = 0.0f;
#pragma unroll
        for (int o = 0; o < D3; ++o) {
            const float g = live ? dy[(r0 + t) * D3 + o] : 0.0f;
#pragma unroll
            for (int j = 0; j < HP; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W3s + o * HP + j);
                v[j] = fmaf(g, w.x, v[j]); v[j + 1] = fmaf(g, w.y, v[j + 1]); v[j + 2] = fmaf(g, w.z, v[j + 2]); v[j + 3] = fmaf(g, w.w, v[j + 3]);
            }
        }
        __syncthreads();
        // a2 = act(p2) (written back into the tile), d2 = v * act'(p2) (kept in registers)
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 p = *reinterpret_cast<const float4*>(tile + t * S + j);
            float4 a, d;
            m3_act_d<ACT>(p.x, a.x, d.x); m3_act_d<ACT>(p.y, a.y, d.y); m3_act_d<ACT>(p.z, a.z, d.z); m3_act_d<ACT>(p.w, a.w, d.w);
            v[j] *= d.x; v[j + 1] *= d.y; v[j + 2] *= d.z; v[j + 3] *= d.w;
            *reinterpret_cast<float4*>(tile + t * S + j) = a;
        }
        __syncthreads();
        m3_tile_store<H>(tile, a2 + r0 * H, nrows);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < H; j += 4) *reinterpret_cast<float4*>(tile + t * S + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        __syncthreads();
        m3_tile_store<H>(tile, d2 + r0 * H, nrows);
        // u = W2^T d2
        float u[HP];
#pragma unroll
        for (int k = 0; k < HP; ++k) u[k] = 0.0f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float g = v[j];
#pragma unroll
            for (int k = 0; k < HP; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W2s + j * HP + k);
                u[k] = fmaf(g, w.x, u[k]); u[k + 1] = fmaf(g, w.y, u[k + 1]); u[k + 2] = fmaf(g, w.z, u[k + 2]); u[k + 3] = fmaf(g, w.w, u[k + 3]);
            }
        }
        __syncthreads();                                   // d2 tile written out
        m3_tile_load<H>(tile, p1 + r0 * H, nrows);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < H; j += 4) {
            const float4 p = *reinterpret_cast<const float4*>(tile + t * S + j);
            float4 a, d;
            m3_act_d<ACT>(p.x, a.x, d.x); m3_act_d<ACT>(p.y, a.y, d.y); m3_act_d<ACT>(p.z, a.z, d.z); m3_act_d<ACT>(p.w, a.w, d.w);
            u[j] *= d.x; u[j + 1] *= d.y; u[j + 2] *= d.z; u[j + 3] *= d.w;
            *reinterpret_cast<float4*>(tile + t * S + j) = a;
        }
        __syncthreads();
        m3_tile_store<H>(tile, a1 + r0 * H, nrows);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < H; j += 4) *reinterpret_cast<float4*>(tile + t * S + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
        __syncthreads();
        m3_tile_store<H>(tile, d1 + r0 * H, nrows);
        if (dx) {   // dx = W1^T d1
            float w_[D0P];
#pragma unroll
            for (int m = 0; m < D0P; ++m) w_[m] = 0.0f;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float g = u[k];
#pragma unroll
                for (int m = 0; m < D0P; m += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(W1s + k * D0P + m);
                    w_[m] = fmaf(g, w.x, w_[m]); w_[m + 1] = fmaf(g, w.y, w_[m + 1]); w_[m + 2] = fmaf(g, w.z, w_[m + 2]); w_[m + 3] = fmaf(g, w.w, w_[m + 3]);
                }
            }
            __syncthreads();                               // d1 tile written out
#pragma unroll
            for (int m = 0; m < D0P; m += 4)
                *reinterpret_cast<float4*>(tile + t * (D0P + 4) + m) = make_float4(w_[m], w_[m + 1], w_[m + 2], w_[m + 3]);
            __syncthreads();
            m3_tile_store_any<D0, D0P>(tile, dx + r0 * lddx, lddx, nrows);
        }
    }
}

static unsigned m3_grid(int64_t rows) {
    const int64_t blocks = (rows + M3_ROWS - 1) / M3_ROWS;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 6;
    return static_cast<unsigned>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}
template <typename K>
static bool m3_configure(K kernel, size_t smem, bool& done) {
    if (done) return true;
    if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)) != cudaSuccess) return false;
    done = true;
    return true;
}

struct M3Fwd {
    const float *x; int64_t rows, ldx; const float *W1, *b1, *W2, *b2, *W3, *b3; float *p1, *p2, *y;
};
struct M3Bwd {
    const float *p1, *p2, *dy; int64_t rows; const float *W1, *W2, *W3; float *a1, *a2, *d1, *d2, *dx; int64_t lddx;
};
template <int D0, int H, int D3, int ACT>
static void m3_fwd(const M3Fwd& a, cudaStream_t s) {
    constexpr size_t smem = sizeof(float) * M3Smem<D0, H, D3>::TOTAL;
    static bool done = false;
    if (!m3_configure(mlp3_train_fwd_kernel<D0, H, D3, ACT>, smem, done)) return;   // (the launch below then reports the error)
    mlp3_train_fwd_kernel<D0, H, D3, ACT><<<m3_grid(a.rows), M3_ROWS, smem, s>>>(a.x, a.rows, a.ldx, a.W1, a.b1, a.W2, a.b2, a.W3, a.b3, a.p1,
                                                                                a.p2, a.y);
}
template <int D0, int H, int D3, int ACT>
static void m3_bwd(const M3Bwd& a, cudaStream_t s) {
    constexpr size_t smem = sizeof(float) * M3Smem<D0, H, D3>::TOTAL;
    static bool done = false;
    if (!m3_configure(mlp3_train_bwd_kernel<D0, H, D3, ACT>, smem, done)) return;
    mlp3_train_bwd_kernel<D0, H, D3, ACT><<<m3_grid(a.rows), M3_ROWS, smem, s>>>(a.p1, a.p2, a.dy, a.rows, a.W1, a.W2, a.W3, a.a1, a.a2, a.d1,
                                                                                a.d2, a.dx, a.lddx);
}

// the instantiated (d0, h, d3, act) combinations: the pair MLPs of the four EdgeNetwork_dense layers (c_init 2, c_hid 8, c_final 4:
// elu) of the BASELINE configuration.  Measured and NOT instantiated: the 30 -> 60 -> 60 -> 1 silu head
// (EdgeScoreNetwork_dense.final).  With 60 hidden units per thread the fully unrolled chains are ~10^4 straight-line instructions
// executed once per row (instruction-fetch bound: 106 us forward / 182 us backward per 102,400 rows under ncu, against ~110 / ~110
// for the three GEMMs + activations), and the step was 0.2 ms SLOWER with it (8.89 vs 8.69 ms) -- the head stays layer-granular.
#define M3_DISPATCH(CALL)                                              \
    if (d0 == 4 && h == 16 && d3 == 8 && act == 5) { CALL(4, 16, 8, 5); }        \
    else if (d0 == 16 && h == 16 && d3 == 8 && act == 5) { CALL(16, 16, 8, 5); } \
    else if (d0 == 16 && h == 16 && d3 == 4 && act == 5) { CALL(16, 16, 4, 5); } \
    else return MOLSDE_ERR_UNSUPPORTED;

}  // namespace molsde

using namespace molsde;

extern "C" {

int molsde_mlp3_train_supported(int32_t d0, int32_t h, int32_t d3, int32_t act) {
    return act == 5 && h == 16 && ((d0 == 4 && d3 == 8) || (d0 == 16 && d3 == 8) || (d0 == 16 && d3 == 4));
}

int molsde_mlp3_train_fwd(const float* x, int64_t rows, int64_t ldx, int32_t d0, int32_t h, int32_t d3, int32_t act, const float* W1,
                          const float* b1, const float* W2, const float* b2, const float* W3, const float* b3, float* p1, float* p2,
                          float* y, void* stream) {
    if (!x || !W1 || !b1 || !W2 || !b2 || !W3 || !b3 || !p1 || !p2 || !y || rows < 0 || ldx < d0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    const M3Fwd a{x, rows, ldx, W1, b1, W2, b2, W3, b3, p1, p2, y};
    cudaStream_t s = as_stream(stream);
#define M3_CALL_FWD(A, B, C, D) m3_fwd<A, B, C, D>(a, s)
    M3_DISPATCH(M3_CALL_FWD)
    return check_launch("mlp3_train_fwd");
}

int molsde_mlp3_train_bwd(const float* p1, const float* p2, const float* dy, int64_t rows, int32_t d0, int32_t h, int32_t d3, int32_t act,
                          const float* W1, const float* W2, const float* W3, float* a1, float* a2, float* d1, float* d2, float* dx,
                          int64_t lddx, void* stream) {
    if (!p1 || !p2 || !dy || !W1 || !W2 || !W3 || !a1 || !a2 || !d1 || !d2 || rows < 0 || (dx && lddx < d0)) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    const M3Bwd a{p1, p2, dy, rows, W1, W2, W3, a1, a2, d1, d2, dx, lddx};
    cudaStream_t s = as_stream(stream);
#define M3_CALL_BWD(A, B, C, D) m3_bwd<A, B, C, D>(a, s)
    M3_DISPATCH(M3_CALL_BWD)
    return check_launch("mlp3_train_bwd");
}

}  // extern "C"
