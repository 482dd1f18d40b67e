// Whole-chain kernels for the NARROW 3-layer MLPs of the training step (EdgeNetwork_dense.mlp: 2C -> 2max(C,C') -> same -> C',
// edge_network_dense.py:120-123, and EdgeScoreNetwork_dense.final: 30 -> 60 -> 60 -> 1, invariant_scorenetwork_dense.py:60-62,
// applied to the B*Nm^2 atom pairs of a batch).  The layer-granular form was 3 GEMM + 2 activation launches forward and
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

template <int D0, int H, int D3, int ACT>
__global__ void __launch_bounds__(128, 3) mlp3_train_fwd_kernel(const float* __restrict__ x, int64_t rows, int64_t ldx,
                                                             const float* __restrict__ W1, const float* __restrict__ b1,
                                                             const float* __restrict__ W2, const float* __restrict__ b2,
                                                             const float* __restrict__ W3, const float* __restrict__ b3,
                                                             float* __restrict__ p1, float* __restrict__ p2, float* __restrict__ y) {
    constexpr int D0P = mp4(D0), HP = mp4(H);
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4 (float4 row stores)");
    __shared__ __align__(16) float W1s[H * D0P];
    __shared__ __align__(16) float W2s[H * HP];
    __shared__ __align__(16) float W3s[D3 * HP];
    __shared__ float b1s[H], b2s[H], b3s[D3];
    m3_stage<H, D0, D0P>(W1s, W1);
    m3_stage<H, H, HP>(W2s, W2);
    m3_stage<D3, H, HP>(W3s, W3);
    for (int i = threadIdx.x; i < H; i += blockDim.x) { b1s[i] = b1[i]; b2s[i] = b2[i]; }
    for (int i = threadIdx.x; i < D3; i += blockDim.x) b3s[i] = b3[i];
    __syncthreads();
    const bool vx = (D0 % 4 == 0) && (ldx % 4 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
    for (int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; row < rows;
         row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        float xv[D0P];
        const float* xr = x + row * ldx;
        if (vx) {
#pragma unroll
            for (int k = 0; k < D0P; k += 4) {
                const float4 t = *reinterpret_cast<const float4*>(xr + k);
                xv[k] = t.x; xv[k + 1] = t.y; xv[k + 2] = t.z; xv[k + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < D0P; ++k) xv[k] = k < D0 ? xr[k] : 0.0f;
        }
        float h[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            float acc = b1s[i];
#pragma unroll
            for (int k = 0; k < D0P; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W1s + i * D0P + k);
                acc = fmaf(w.x, xv[k], acc); acc = fmaf(w.y, xv[k + 1], acc); acc = fmaf(w.z, xv[k + 2], acc); acc = fmaf(w.w, xv[k + 3], acc);
            }
            h[i] = acc;
        }
        float* o1 = p1 + row * H;
#pragma unroll
        for (int i = 0; i < H; i += 4) *reinterpret_cast<float4*>(o1 + i) = make_float4(h[i], h[i + 1], h[i + 2], h[i + 3]);
#pragma unroll
        for (int i = 0; i < H; ++i) h[i] = m3_act<ACT>(h[i]);
        float g[H];
#pragma unroll
        for (int i = 0; i < H; ++i) {
            float acc = b2s[i];
#pragma unroll
            for (int k = 0; k < HP; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W2s + i * HP + k);
                acc = fmaf(w.x, h[k], acc); acc = fmaf(w.y, h[k + 1], acc); acc = fmaf(w.z, h[k + 2], acc); acc = fmaf(w.w, h[k + 3], acc);
            }
            g[i] = acc;
        }
        float* o2 = p2 + row * H;
#pragma unroll
        for (int i = 0; i < H; i += 4) *reinterpret_cast<float4*>(o2 + i) = make_float4(g[i], g[i + 1], g[i + 2], g[i + 3]);
#pragma unroll
        for (int i = 0; i < H; ++i) g[i] = m3_act<ACT>(g[i]);
        float* yo = y + row * D3;
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
        if (D3 % 4 == 0) {   // (y is a fresh [rows, D3] allocation: 16-byte aligned rows)
#pragma unroll
            for (int o = 0; o + 3 < D3; o += 4) *reinterpret_cast<float4*>(yo + o) = make_float4(out[o], out[o + 1], out[o + 2], out[o + 3]);
        } else {
#pragma unroll
            for (int o = 0; o < D3; ++o) yo[o] = out[o];
        }
    }
}

template <int D0, int H, int D3, int ACT>
__global__ void __launch_bounds__(128, 3) mlp3_train_bwd_kernel(const float* __restrict__ p1, const float* __restrict__ p2,
                                                             const float* __restrict__ dy, int64_t rows,
                                                             const float* __restrict__ W1, const float* __restrict__ W2,
                                                             const float* __restrict__ W3, float* __restrict__ a1, float* __restrict__ a2,
                                                             float* __restrict__ d1, float* __restrict__ d2, float* __restrict__ dx,
                                                             int64_t lddx) {
    constexpr int D0P = mp4(D0), HP = mp4(H);
    static_assert(H % 4 == 0, "hidden width must be a multiple of 4 (float4 row accesses)");
    __shared__ __align__(16) float W1s[H * D0P];
    __shared__ __align__(16) float W2s[H * HP];
    __shared__ __align__(16) float W3s[D3 * HP];
    m3_stage<H, D0, D0P>(W1s, W1);
    m3_stage<H, H, HP>(W2s, W2);
    m3_stage<D3, H, HP>(W3s, W3);
    __syncthreads();
    const bool vdx = dx && (D0 % 4 == 0) && (lddx % 4 == 0) && ((reinterpret_cast<uintptr_t>(dx) & 15) == 0);
    for (int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; row < rows;
         row += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        // t = W3^T dy
        float t[HP];
#pragma unroll
        for (int j = 0; j < HP; ++j) t[j] = 0.0f;
#pragma unroll
        for (int o = 0; o < D3; ++o) {
            const float g = dy[row * D3 + o];
#pragma unroll
            for (int j = 0; j < HP; j += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W3s + o * HP + j);
                t[j] = fmaf(g, w.x, t[j]); t[j + 1] = fmaf(g, w.y, t[j + 1]); t[j + 2] = fmaf(g, w.z, t[j + 2]); t[j + 3] = fmaf(g, w.w, t[j + 3]);
            }
        }
        // a2 = act(p2), d2 = t * act'(p2)
        {
            const float* pr = p2 + row * H;
            float* ar = a2 + row * H;
            float* dr = d2 + row * H;
#pragma unroll
            for (int j = 0; j < H; j += 4) {
                const float4 p = *reinterpret_cast<const float4*>(pr + j);
                float4 a, d;
                m3_act_d<ACT>(p.x, a.x, d.x); m3_act_d<ACT>(p.y, a.y, d.y); m3_act_d<ACT>(p.z, a.z, d.z); m3_act_d<ACT>(p.w, a.w, d.w);
                t[j] *= d.x; t[j + 1] *= d.y; t[j + 2] *= d.z; t[j + 3] *= d.w;
                *reinterpret_cast<float4*>(ar + j) = a;
                *reinterpret_cast<float4*>(dr + j) = make_float4(t[j], t[j + 1], t[j + 2], t[j + 3]);
            }
        }
        // u = W2^T d2
        float u[HP];
#pragma unroll
        for (int k = 0; k < HP; ++k) u[k] = 0.0f;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            const float g = t[j];
#pragma unroll
            for (int k = 0; k < HP; k += 4) {
                const float4 w = *reinterpret_cast<const float4*>(W2s + j * HP + k);
                u[k] = fmaf(g, w.x, u[k]); u[k + 1] = fmaf(g, w.y, u[k + 1]); u[k + 2] = fmaf(g, w.z, u[k + 2]); u[k + 3] = fmaf(g, w.w, u[k + 3]);
            }
        }
        {
            const float* pr = p1 + row * H;
            float* ar = a1 + row * H;
            float* dr = d1 + row * H;
#pragma unroll
            for (int j = 0; j < H; j += 4) {
                const float4 p = *reinterpret_cast<const float4*>(pr + j);
                float4 a, d;
                m3_act_d<ACT>(p.x, a.x, d.x); m3_act_d<ACT>(p.y, a.y, d.y); m3_act_d<ACT>(p.z, a.z, d.z); m3_act_d<ACT>(p.w, a.w, d.w);
                u[j] *= d.x; u[j + 1] *= d.y; u[j + 2] *= d.z; u[j + 3] *= d.w;
                *reinterpret_cast<float4*>(ar + j) = a;
                *reinterpret_cast<float4*>(dr + j) = make_float4(u[j], u[j + 1], u[j + 2], u[j + 3]);
            }
        }
        if (dx) {   // dx = W1^T d1
            float v[D0P];
#pragma unroll
            for (int m = 0; m < D0P; ++m) v[m] = 0.0f;
#pragma unroll
            for (int k = 0; k < H; ++k) {
                const float g = u[k];
#pragma unroll
                for (int m = 0; m < D0P; m += 4) {
                    const float4 w = *reinterpret_cast<const float4*>(W1s + k * D0P + m);
                    v[m] = fmaf(g, w.x, v[m]); v[m + 1] = fmaf(g, w.y, v[m + 1]); v[m + 2] = fmaf(g, w.z, v[m + 2]); v[m + 3] = fmaf(g, w.w, v[m + 3]);
                }
            }
            float* xr = dx + row * lddx;
            if (vdx) {
#pragma unroll
                for (int m = 0; m < D0P; m += 4) *reinterpret_cast<float4*>(xr + m) = make_float4(v[m], v[m + 1], v[m + 2], v[m + 3]);
            } else {
#pragma unroll
                for (int m = 0; m < D0P; ++m)
                    if (m < D0) xr[m] = v[m];
            }
        }
    }
}

static unsigned m3_grid(int64_t rows) {
    const int64_t blocks = (rows + 127) / 128;
    const int64_t cap = static_cast<int64_t>(kNumSMs) * 8;
    return static_cast<unsigned>(blocks < cap ? (blocks < 1 ? 1 : blocks) : cap);
}

struct M3Fwd {
    const float *x; int64_t rows, ldx; const float *W1, *b1, *W2, *b2, *W3, *b3; float *p1, *p2, *y;
};
struct M3Bwd {
    const float *p1, *p2, *dy; int64_t rows; const float *W1, *W2, *W3; float *a1, *a2, *d1, *d2, *dx; int64_t lddx;
};
template <int D0, int H, int D3, int ACT>
static void m3_fwd(const M3Fwd& a, cudaStream_t s) {
    mlp3_train_fwd_kernel<D0, H, D3, ACT><<<m3_grid(a.rows), 128, 0, s>>>(a.x, a.rows, a.ldx, a.W1, a.b1, a.W2, a.b2, a.W3, a.b3, a.p1, a.p2, a.y);
}
template <int D0, int H, int D3, int ACT>
static void m3_bwd(const M3Bwd& a, cudaStream_t s) {
    mlp3_train_bwd_kernel<D0, H, D3, ACT><<<m3_grid(a.rows), 128, 0, s>>>(a.p1, a.p2, a.dy, a.rows, a.W1, a.W2, a.W3, a.a1, a.a2, a.d1, a.d2,
                                                                         a.dx, a.lddx);
}

// the instantiated (d0, h, d3, act) combinations: the pair MLPs of the four EdgeNetwork_dense layers (c_init 2, c_hid 8, c_final 4:
// elu) and the 30 -> 60 -> 60 -> 1 head (silu) of the BASELINE configuration
#define M3_DISPATCH(CALL)                                              \
    if (d0 == 4 && h == 16 && d3 == 8 && act == 5) { CALL(4, 16, 8, 5); }        \
    else if (d0 == 16 && h == 16 && d3 == 8 && act == 5) { CALL(16, 16, 8, 5); } \
    else if (d0 == 16 && h == 16 && d3 == 4 && act == 5) { CALL(16, 16, 4, 5); } \
    else if (d0 == 30 && h == 60 && d3 == 1 && act == 2) { CALL(30, 60, 1, 2); } \
    else return MOLSDE_ERR_UNSUPPORTED;

}  // namespace molsde

using namespace molsde;

extern "C" {

int molsde_mlp3_train_supported(int32_t d0, int32_t h, int32_t d3, int32_t act) {
    return (act == 5 && h == 16 && ((d0 == 4 && d3 == 8) || (d0 == 16 && d3 == 8) || (d0 == 16 && d3 == 4))) ||
           (act == 2 && d0 == 30 && h == 60 && d3 == 1);
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
