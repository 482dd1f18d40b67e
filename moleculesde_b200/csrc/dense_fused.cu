// Fused inference kernels of the dense 3D->2D EDGE score network (sampling / get_score_fn path).
//
// Reference: layers/edge_network_dense.py:56-128 (EdgeLayer attention, EdgeNetwork_dense.forward) and
// invariant_scorenetwork_dense.py:68-93 (EdgeScoreNetwork_dense.forward).  The layer-granular kernels of dense.cu +
// mlp_rows.cu materialise, per layer, the channels-last pair tensor [B,Nm,Nm,2C] (written strided, read back by the
// per-pair MLP), the MLP output [B*Nm^2, C'] and the all-channels buffer [B,Nm,Nm,30]: ~330 MB of HBM traffic and three
// launches per layer for ~100 MB of algorithmic bytes.  Here every adjacency stack stays CHANNEL-MAJOR [B,C,Nm,Nm] (so
// the thread of pair (i,j) reads and writes coalesced along j), and per layer there are two kernels:
//   dense_attn_sym      S[b,c,i,j] = (A_ij + A_ji)/2,  A_ij = mean_h tanh(<q_c[i,h], k_c[j,h]>/sqrt(ds))      (:66-80)
//   dense_pair_mlp      adjc'[b,c',i,j] = (m_ij + m_ji) f_i f_j,  m = MLP_elu([S | adjc][b,:,i,j])              (:120-126)
// and one for the head:
//   dense_edge_final_mlp  out[b,i,j] = MLP_silu(all channels of all layers)[i,j] (i != j) f_i f_j scale_b     (:84-93)
// Pairs with f_i f_j = 0 (padding atoms) are skipped: their outputs are 0 whatever the MLP returns (the caller zero-fills the
// output stacks), nothing else reads their intermediates, and the kernels enumerate the n x n valid pairs of a graph densely
// (compact valid-atom list per CTA), so padding occupies no lanes.  m_ji: the inputs of layers >= 1 are symmetric BITWISE by construction (S is symmetrised, adjc' of the
// previous layer is (m_ij + m_ji) f f with commutative fp adds), so there m_ji == m_ij and `symmetric = 1` skips the second
// evaluation; layer 0 sees adj, adj^2 of an arbitrary (in the reference's sampler: non-symmetric) adjacency and evaluates
// the transposed pair too.  All math is fp32 FFMA with the weights broadcast from shared memory (rows are independent:
// deterministic); activations use the SFU exp with a series branch near 0 (relative error ~1e-7).
#include "common.cuh"
#include "tc05.cuh"
#include <stdlib.h>

namespace molsde {

constexpr int DF_NM = 64;       // max padded atoms per graph
constexpr int DF_K0 = 16;       // pair MLP: padded input width (2C <= 16)
constexpr int DF_H = 16;        //           hidden width
constexpr int DF_CO = 8;        //           padded output channels
constexpr int DF_FK = 32;       // final MLP: padded input channels (fdim <= 32)
constexpr int DF_FH = 64;       //            padded hidden width (2*fdim <= 64)
constexpr int DF_MAX_SEGS = 6;

__device__ __forceinline__ float df_tanh(float v) {
    // tanh(v) = 1 - 2 / (exp(2v) + 1): two SFU ops + three FMA-pipe ops, ABSOLUTE error ~1e-7 everywhere (exp overflow -> 1, underflow
    // -> -1).  Near 0 the relative error grows (cancellation), which is irrelevant here: every use feeds an average / a bounded MLP
    // whose parity bar is 1e-4 of the output scale.  (A series branch for |v| < 0.1 cost 9 more instructions per call: the attention
    // kernel evaluates 8 tanh per pair and channel and is issue-bound.)
    const float e = __expf(2.0f * v);
    return 1.0f - __fdividef(2.0f, e + 1.0f);
}
__device__ __forceinline__ float df_elu(float v) {
    const float series = v * fmaf(v, fmaf(v, fmaf(v, 1.0f / 24.0f, 1.0f / 6.0f), 0.5f), 1.0f);
    const float em1 = v > -0.1f ? series : __expf(v) - 1.0f;
    return v > 0.0f ? v : em1;
}
__device__ __forceinline__ float df_silu(float v) { return __fdividef(v, 1.0f + __expf(-v)); }

// Compact list of the valid atoms of one graph (flags != 0) in shared memory: idx[0..n) ascending; returns n.  Threads then enumerate
// the n*n valid ordered pairs densely (t -> (idx[t / n], idx[t % n])), so padding atoms do not occupy lanes.  Call with all threads.
__device__ __forceinline__ int df_valid_nodes(const float* __restrict__ fl, int Nm, int* idx, int* count) {
    if (threadIdx.x < 32) {
        int n = 0;
        for (int base = 0; base < Nm; base += 32) {
            const int node = base + threadIdx.x;
            const bool ok = node < Nm && fl[node] != 0.0f;
            const unsigned m = __ballot_sync(0xffffffffu, ok);
            if (ok) idx[n + __popc(m & ((1u << threadIdx.x) - 1u))] = node;
            n += __popc(m);
        }
        if (threadIdx.x == 0) *count = n;
    }
    __syncthreads();
    return *count;
}

// ---------------------------------------------------------------------------------------
// S[b,c,i,j]: grid (B, C), 256 threads.  Q, K: [B*Nm, ldq], channel c at columns c*W .. c*W+W-1 (W = H*ds <= 32, ds % 4 == 0)
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_attn_sym_kernel(const float* __restrict__ Q, const float* __restrict__ K, int64_t ldq, int W, int ds,
                      const float* __restrict__ flags, int C, int Nm, float* __restrict__ S) {
    __shared__ __align__(16) float sQ[DF_NM][36], sK[DF_NM][36];   // 144 B rows: float4 aligned, conflict-free per quarter warp
    __shared__ float sA[DF_NM][DF_NM + 1];
    const int b = blockIdx.x, c = blockIdx.y, w4 = W >> 2;
    for (int idx = threadIdx.x; idx < Nm * w4; idx += blockDim.x) {
        const int r = idx / w4, k4 = idx % w4;
        const int64_t g = (static_cast<int64_t>(b) * Nm + r) * ldq + c * W + 4 * k4;
        *reinterpret_cast<float4*>(&sQ[r][4 * k4]) = __ldg(reinterpret_cast<const float4*>(Q + g));
        *reinterpret_cast<float4*>(&sK[r][4 * k4]) = __ldg(reinterpret_cast<const float4*>(K + g));
    }
    __shared__ int vidx[DF_NM], vcount;
    const float* fl = flags + static_cast<int64_t>(b) * Nm;
    const int n = df_valid_nodes(fl, Nm, vidx, &vcount);   // (also the barrier after the Q / K staging)
    const int H = W / ds, d4 = ds >> 2;
    const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(ds));
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        const int a = t / n, cc = t % n;
        const float4* q = reinterpret_cast<const float4*>(sQ[vidx[a]]);
        const float4* k = reinterpret_cast<const float4*>(sK[vidx[cc]]);
        float s = 0.0f;
        for (int h = 0; h < H; ++h) {
            float d = 0.0f;
            for (int kk = 0; kk < d4; ++kk) {
                const float4 x = q[h * d4 + kk], y = k[h * d4 + kk];
                d = fmaf(x.x, y.x, d); d = fmaf(x.y, y.y, d); d = fmaf(x.z, y.z, d); d = fmaf(x.w, y.w, d);
            }
            s += df_tanh(d * inv_sqrt);
        }
        sA[a][cc] = s / static_cast<float>(H);
    }
    __syncthreads();
    float* out = S + (static_cast<int64_t>(b) * C + c) * Nm * Nm;   // (entries of padding pairs are never read: left unwritten)
    for (int t = threadIdx.x; t < n * n; t += blockDim.x) {
        const int a = t / n, cc = t % n;
        out[vidx[a] * Nm + vidx[cc]] = (sA[a][cc] + sA[cc][a]) * 0.5f;
    }
}

// ---------------------------------------------------------------------------------------
// per-pair MLP (2C -> H -> H -> C', elu) + symmetrise + mask; thread = ordered pair (i, j) of graph blockIdx.y
// ---------------------------------------------------------------------------------------
struct PairMlpSmem {   // k-major weights: every input k streams one row of 16 (8) output weights as broadcast float4 loads
    float W0[DF_K0][DF_H], W1[DF_H][DF_H], W2[DF_H][DF_CO], b0[DF_H], b1[DF_H], b2[DF_CO];
};
template <int K, int N>
__device__ __forceinline__ void df_layer(const float* __restrict__ Wk, const float* __restrict__ bias, const float (&x)[K], float (&y)[N]) {
#pragma unroll
    for (int o = 0; o < N; ++o) y[o] = bias[o];
#pragma unroll
    for (int k = 0; k < K; ++k) {
        const float4* w = reinterpret_cast<const float4*>(Wk + k * N);
#pragma unroll
        for (int o4 = 0; o4 < N / 4; ++o4) {
            const float4 ww = w[o4];
            y[4 * o4] = fmaf(x[k], ww.x, y[4 * o4]); y[4 * o4 + 1] = fmaf(x[k], ww.y, y[4 * o4 + 1]);
            y[4 * o4 + 2] = fmaf(x[k], ww.z, y[4 * o4 + 2]); y[4 * o4 + 3] = fmaf(x[k], ww.w, y[4 * o4 + 3]);
        }
    }
}
__device__ __forceinline__ void pair_mlp_eval(const PairMlpSmem& P, const float (&x)[DF_K0], float (&y)[DF_CO]) {
    float h1[DF_H], h2[DF_H];
    df_layer<DF_K0, DF_H>(&P.W0[0][0], P.b0, x, h1);
#pragma unroll
    for (int o = 0; o < DF_H; ++o) h1[o] = df_elu(h1[o]);
    df_layer<DF_H, DF_H>(&P.W1[0][0], P.b1, h1, h2);
#pragma unroll
    for (int o = 0; o < DF_H; ++o) h2[o] = df_elu(h2[o]);
    df_layer<DF_H, DF_CO>(&P.W2[0][0], P.b2, h2, y);
}

__global__ void __launch_bounds__(256, 2)
dense_pair_mlp_kernel(const float* __restrict__ S, const float* __restrict__ adjc, const float* __restrict__ flags,
                      const float* __restrict__ W0, const float* __restrict__ b0, const float* __restrict__ W1,
                      const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2, int Cin, int Hd,
                      int Co, int Nm, int symmetric, float* __restrict__ adjc_next) {
    __shared__ __align__(16) PairMlpSmem P;
    const int K0 = 2 * Cin;
    for (int i = threadIdx.x; i < DF_K0 * DF_H; i += blockDim.x) {
        const int k = i / DF_H, o = i % DF_H;
        P.W0[k][o] = (o < Hd && k < K0) ? W0[o * K0 + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < DF_H * DF_H; i += blockDim.x) {
        const int k = i / DF_H, o = i % DF_H;
        P.W1[k][o] = (o < Hd && k < Hd) ? W1[o * Hd + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < DF_H * DF_CO; i += blockDim.x) {
        const int k = i / DF_CO, o = i % DF_CO;
        P.W2[k][o] = (o < Co && k < Hd) ? W2[o * Hd + k] : 0.0f;
    }
    if (threadIdx.x < DF_H) {
        P.b0[threadIdx.x] = threadIdx.x < Hd ? b0[threadIdx.x] : 0.0f;
        P.b1[threadIdx.x] = threadIdx.x < Hd ? b1[threadIdx.x] : 0.0f;
    }
    if (threadIdx.x < DF_CO) P.b2[threadIdx.x] = threadIdx.x < Co ? b2[threadIdx.x] : 0.0f;
    const int NN = Nm * Nm, b = blockIdx.y;
    __shared__ int vidx[DF_NM], vcount;
    const int n = df_valid_nodes(flags + static_cast<int64_t>(b) * Nm, Nm, vidx, &vcount);   // (also the barrier after the weight staging)
    // (adjc_next is zero-filled by the caller: padding pairs stay 0.  One pair per thread: a CTA that walks several pair blocks to
    //  amortise the weight staging was measured 4x SLOWER -- 295 vs 68 us -- the kernel is latency-bound per pair and needs the threads)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n) return;
    const int i = vidx[t / n], j = vidx[t % n], p = i * Nm + j;
    const float fi = flags[b * Nm + i], fj = flags[b * Nm + j];
    float* out = adjc_next + static_cast<int64_t>(b) * Co * NN + p;
    const float* Sb = S + static_cast<int64_t>(b) * Cin * NN;
    const float* Ab = adjc + static_cast<int64_t>(b) * Cin * NN;
    float x[DF_K0], y[DF_CO], yt[DF_CO];
#pragma unroll
    for (int k = 0; k < DF_K0; ++k) x[k] = 0.0f;
#pragma unroll
    for (int c = 0; c < DF_K0 / 2; ++c)
        if (c < Cin) { x[c] = __ldg(Sb + c * NN + p); }
    // (the adjacency half sits at columns Cin .. 2Cin-1 of the reference's cat([A, adj]), :120)
    if (Cin == DF_K0 / 2) {
#pragma unroll
        for (int c = 0; c < DF_K0 / 2; ++c) x[DF_K0 / 2 + c] = __ldg(Ab + c * NN + p);
    } else {   // Cin == 2 (first layer): columns 2, 3
        x[2] = __ldg(Ab + p);
        x[3] = __ldg(Ab + NN + p);
    }
    pair_mlp_eval(P, x, y);
    if (symmetric || i == j) {
#pragma unroll
        for (int c = 0; c < DF_CO; ++c) yt[c] = y[c];
    } else {
        const int pt = j * Nm + i;
        if (Cin == DF_K0 / 2) {
#pragma unroll
            for (int c = 0; c < DF_K0 / 2; ++c) x[DF_K0 / 2 + c] = __ldg(Ab + c * NN + pt);
        } else {
            x[2] = __ldg(Ab + pt);
            x[3] = __ldg(Ab + NN + pt);
        }
        pair_mlp_eval(P, x, yt);   // (S is symmetric by construction: only the adjacency half changes)
    }
#pragma unroll
    for (int c = 0; c < DF_CO; ++c)
        if (c < Co) out[static_cast<int64_t>(c) * NN] = ((y[c] + yt[c]) * fj) * fi;   // pair_post order (:124-126)
}

// ---------------------------------------------------------------------------------------
// final head: MLP_silu(fdim -> H1 -> H2 -> 1) over the channels of all adjacency stacks, zero diagonal, mask, scale
// ---------------------------------------------------------------------------------------
struct DenseSegs {
    const float* ptr[DF_MAX_SEGS];   // [B, ch, Nm, Nm] each
    int ch[DF_MAX_SEGS];
    int n;
};

__global__ void __launch_bounds__(256, 2)
dense_edge_final_mlp_kernel(DenseSegs segs, const float* __restrict__ flags, const float* __restrict__ scale,
                            const float* __restrict__ W0, const float* __restrict__ b0, const float* __restrict__ W1,
                            const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2, int F, int H1,
                            int H2, int Nm, float* __restrict__ out) {
    extern __shared__ __align__(16) float fsm[];
    float* W0k = fsm;                         // [DF_FK k][DF_FH o]   k-major
    float* W1k = W0k + DF_FK * DF_FH;         // [DF_FH k][DF_FH o]
    float* b0s = W1k + DF_FH * DF_FH;         // [DF_FH]
    float* b1s = b0s + DF_FH;
    float* w2s = b1s + DF_FH;
    for (int i = threadIdx.x; i < DF_FK * DF_FH; i += blockDim.x) {
        const int k = i / DF_FH, o = i % DF_FH;
        W0k[i] = (k < F && o < H1) ? W0[o * F + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < DF_FH * DF_FH; i += blockDim.x) {
        const int k = i / DF_FH, o = i % DF_FH;
        W1k[i] = (k < H1 && o < H2) ? W1[o * H1 + k] : 0.0f;
    }
    if (threadIdx.x < DF_FH) {
        b0s[threadIdx.x] = threadIdx.x < H1 ? b0[threadIdx.x] : 0.0f;
        b1s[threadIdx.x] = threadIdx.x < H2 ? b1[threadIdx.x] : 0.0f;
        w2s[threadIdx.x] = threadIdx.x < H2 ? W2[threadIdx.x] : 0.0f;
    }
    const int NN = Nm * Nm, b = blockIdx.y;
    __shared__ int vidx[DF_NM], vcount;
    const int n = df_valid_nodes(flags + static_cast<int64_t>(b) * Nm, Nm, vidx, &vcount);   // (also the barrier after the weight staging)
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n * n) return;                  // `out` is zero-filled by the caller: padding pairs stay 0
    const int i = vidx[t / n], j = vidx[t % n], p = i * Nm + j;
    const float fi = flags[b * Nm + i], fj = flags[b * Nm + j];
    float* o = out + static_cast<int64_t>(b) * NN + p;
    if (i == j) return;                      // zero diagonal
    float h1[DF_FH];
#pragma unroll
    for (int q = 0; q < DF_FH; ++q) h1[q] = b0s[q];
    int k = 0;
    for (int s = 0; s < segs.n; ++s) {
        const float* base = segs.ptr[s] + static_cast<int64_t>(b) * segs.ch[s] * NN + p;
        for (int c = 0; c < segs.ch[s]; ++c, ++k) {
            const float xk = __ldg(base + static_cast<int64_t>(c) * NN);
            const float4* w = reinterpret_cast<const float4*>(W0k + k * DF_FH);
#pragma unroll
            for (int q4 = 0; q4 < DF_FH / 4; ++q4) {
                const float4 ww = w[q4];
                h1[4 * q4] = fmaf(xk, ww.x, h1[4 * q4]); h1[4 * q4 + 1] = fmaf(xk, ww.y, h1[4 * q4 + 1]);
                h1[4 * q4 + 2] = fmaf(xk, ww.z, h1[4 * q4 + 2]); h1[4 * q4 + 3] = fmaf(xk, ww.w, h1[4 * q4 + 3]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < DF_FH; ++q) h1[q] = df_silu(h1[q]);   // (padded units: silu(0) = 0)
    float r = b2[0];
#pragma unroll 1
    for (int oc = 0; oc < DF_FH / 16; ++oc) {
        float acc[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) acc[q] = b1s[16 * oc + q];
#pragma unroll
        for (int kk = 0; kk < DF_FH; ++kk) {
            const float4* w = reinterpret_cast<const float4*>(W1k + kk * DF_FH + 16 * oc);
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 ww = w[q4];
                acc[4 * q4] = fmaf(h1[kk], ww.x, acc[4 * q4]); acc[4 * q4 + 1] = fmaf(h1[kk], ww.y, acc[4 * q4 + 1]);
                acc[4 * q4 + 2] = fmaf(h1[kk], ww.z, acc[4 * q4 + 2]); acc[4 * q4 + 3] = fmaf(h1[kk], ww.w, acc[4 * q4 + 3]);
            }
        }
#pragma unroll
        for (int q = 0; q < 16; ++q) r = fmaf(df_silu(acc[q]), w2s[16 * oc + q], r);
    }
    float v = (r * fi) * fj;                       // edge_final order: (raw * f_i) * f_j, then the per-graph scale
    *o = scale ? v * scale[b] : v;
}

// ---------------------------------------------------------------------------------------
// Node-side chain of an EdgeLayer whose input is narrow (layers >= 1: Fin = nhid <= 16), one launch instead of three GEMMs:
//   h1 = tanh(W1 x + b1)  [2C*W]   (func_q / func_k layer 0 of all C channels, edge_network_dense.py:45-46)
//   qk[g*W:(g+1)*W] = W2_g h1[g*W:(g+1)*W] + b2_g   for the 2C groups (layer 1)
//   xw = Wv x  [C*Fo]            (x @ func_v.weight of all channels, node_network_dense.py:73)
// thread = (row, group): lanes along 32 consecutive rows, so every weight read is a shared-memory broadcast; fp32 FFMA.
// ---------------------------------------------------------------------------------------
constexpr int NS_W = 32, NS_FIN = 16, NS_ROWS = 32, NS_WARPS = 8;

__global__ void __launch_bounds__(NS_WARPS * 32, 2)
dense_node_side_kernel(const float* __restrict__ X, int64_t rows, int64_t ldx, int Fin, const float* __restrict__ W1,
                       const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2,
                       const float* __restrict__ Wv, int G /*2C groups*/, int NV /*C*Fo*/, float* __restrict__ QK, int64_t ldqk,
                       float* __restrict__ XW, int64_t ldxw) {
    extern __shared__ __align__(16) float ns_smem[];
    float* W1s = ns_smem;                          // [G*W][16]  (k padded to 16)
    float* W2s = W1s + G * NS_W * NS_FIN;          // [G][W][W]
    float* b1s = W2s + G * NS_W * NS_W;            // [G*W]
    float* b2s = b1s + G * NS_W;                   // [G*W]
    float* Wvs = b2s + G * NS_W;                   // [NV][16]
    for (int i = threadIdx.x; i < G * NS_W * NS_FIN; i += blockDim.x) {
        const int o = i / NS_FIN, k = i % NS_FIN;
        W1s[i] = k < Fin ? W1[o * Fin + k] : 0.0f;
    }
    for (int i = threadIdx.x; i < G * NS_W * NS_W; i += blockDim.x) W2s[i] = W2[i];
    for (int i = threadIdx.x; i < G * NS_W; i += blockDim.x) { b1s[i] = b1[i]; b2s[i] = b2[i]; }
    for (int i = threadIdx.x; i < NV * NS_FIN; i += blockDim.x) {
        const int o = i / NS_FIN, k = i % NS_FIN;
        Wvs[i] = k < Fin ? Wv[o * Fin + k] : 0.0f;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nvg = (NV + G - 1) / G;               // xw columns per group-thread
    for (int64_t r0 = static_cast<int64_t>(blockIdx.x) * NS_ROWS; r0 < rows; r0 += static_cast<int64_t>(gridDim.x) * NS_ROWS) {
        const int64_t row = r0 + lane;
        const bool live = row < rows;
        float x[NS_FIN];
#pragma unroll
        for (int k = 0; k < NS_FIN; ++k) x[k] = (live && k < Fin) ? __ldg(X + row * ldx + k) : 0.0f;
        for (int g = warp; g < G; g += NS_WARPS) {
            float h[NS_W];
#pragma unroll
            for (int o = 0; o < NS_W; ++o) {
                const float4* w = reinterpret_cast<const float4*>(W1s + (g * NS_W + o) * NS_FIN);
                float a = b1s[g * NS_W + o];
#pragma unroll
                for (int k4 = 0; k4 < NS_FIN / 4; ++k4) {
                    const float4 ww = w[k4];
                    a = fmaf(x[4 * k4], ww.x, a); a = fmaf(x[4 * k4 + 1], ww.y, a); a = fmaf(x[4 * k4 + 2], ww.z, a); a = fmaf(x[4 * k4 + 3], ww.w, a);
                }
                h[o] = df_tanh(a);
            }
            float* qrow = QK + row * ldqk + g * NS_W;
#pragma unroll 1
            for (int o4 = 0; o4 < NS_W / 4; ++o4) {
                float acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float4* w = reinterpret_cast<const float4*>(W2s + (g * NS_W + 4 * o4 + q) * NS_W);
                    float a = b2s[g * NS_W + 4 * o4 + q];
#pragma unroll
                    for (int k4 = 0; k4 < NS_W / 4; ++k4) {
                        const float4 ww = w[k4];
                        a = fmaf(h[4 * k4], ww.x, a); a = fmaf(h[4 * k4 + 1], ww.y, a); a = fmaf(h[4 * k4 + 2], ww.z, a); a = fmaf(h[4 * k4 + 3], ww.w, a);
                    }
                    acc[q] = a;
                }
                if (live) *reinterpret_cast<float4*>(qrow + 4 * o4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            }
            // this thread's share of xw = Wv x
            for (int c = g * nvg; c < min(NV, (g + 1) * nvg); ++c) {
                const float4* w = reinterpret_cast<const float4*>(Wvs + c * NS_FIN);
                float a = 0.0f;
#pragma unroll
                for (int k4 = 0; k4 < NS_FIN / 4; ++k4) {
                    const float4 ww = w[k4];
                    a = fmaf(x[4 * k4], ww.x, a); a = fmaf(x[4 * k4 + 1], ww.y, a); a = fmaf(x[4 * k4 + 2], ww.z, a); a = fmaf(x[4 * k4 + 3], ww.w, a);
                }
                if (live) XW[row * ldxw + c] = a;
            }
        }
    }
}

// multi_channel MLP of EdgeNetwork_dense (edge_network_dense.py:113-118): out = tanh(W1 elu(W0 v + b0) + b1) * flag, v [rows, K] -> H -> NO;
// thread = row, weights broadcast from shared memory.
__global__ void __launch_bounds__(128)
dense_multi_channel_kernel(const float* __restrict__ V, int64_t rows, int K, const float* __restrict__ W0, const float* __restrict__ b0,
                           int H, const float* __restrict__ W1, const float* __restrict__ b1, int NO, const float* __restrict__ rowflag,
                           float* __restrict__ out) {
    extern __shared__ __align__(16) float mc_smem[];
    float* W0s = mc_smem;            // [16][K]
    float* W1s = W0s + 16 * K;       // [16][16]
    float* bs = W1s + 256;           // b0[16] | b1[16]
    for (int i = threadIdx.x; i < 16 * K; i += blockDim.x) W0s[i] = (i / K) < H ? W0[i] : 0.0f;
    for (int i = threadIdx.x; i < 256; i += blockDim.x) { const int o = i >> 4, k = i & 15; W1s[i] = (o < NO && k < H) ? W1[o * H + k] : 0.0f; }
    if (threadIdx.x < 16) { bs[threadIdx.x] = threadIdx.x < H ? b0[threadIdx.x] : 0.0f; bs[16 + threadIdx.x] = threadIdx.x < NO ? b1[threadIdx.x] : 0.0f; }
    __syncthreads();
    const int64_t row = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (row >= rows) return;
    float h[16];
#pragma unroll
    for (int o = 0; o < 16; ++o) h[o] = bs[o];
    const float4* v4 = reinterpret_cast<const float4*>(V + row * K);
    for (int k4 = 0; k4 < K / 4; ++k4) {
        const float4 v = __ldg(v4 + k4);
#pragma unroll
        for (int o = 0; o < 16; ++o) {
            const float4 w = *reinterpret_cast<const float4*>(W0s + o * K + 4 * k4);
            h[o] = fmaf(v.x, w.x, h[o]); h[o] = fmaf(v.y, w.y, h[o]); h[o] = fmaf(v.z, w.z, h[o]); h[o] = fmaf(v.w, w.w, h[o]);
        }
    }
#pragma unroll
    for (int o = 0; o < 16; ++o) h[o] = df_elu(h[o]);
    const float f = rowflag ? rowflag[row] : 1.0f;
    for (int o = 0; o < NO; ++o) {
        float a = bs[16 + o];
#pragma unroll
        for (int k = 0; k < 16; ++k) a = fmaf(h[k], W1s[o * 16 + k], a);
        out[row * NO + o] = df_tanh(a) * f;
    }
}

// ---------------------------------------------------------------------------------------
// The same head on tcgen05 (round 2): persistent CTAs of four QUADS (128 threads = 128 valid pairs, thread = pair = TMEM lane).
//   A1 row = the pair's <= 32 channel values (fp16 hi/lo)  -> MMA 1 [128 x 32].[64 x 32]^T  -> epilogue: silu(acc + b0) -> A2 row
//   MMA 2 [128 x 64].[64 x 64]^T -> epilogue: r = b2 + sum_k silu(acc + b1[k]) w2[k] -> (r f_i) f_j scale_b
// Both GEMMs are kind::f16 MMAs with the two-way fp16 operand split (22-bit operands, fp32 accumulate in TMEM).  ~1,300 instructions
// per pair instead of ~8,500 on the FFMA path above (which stays as the fallback for wider heads).
// ---------------------------------------------------------------------------------------
constexpr int FT_QUADS = 4, FT_THREADS = 128 * FT_QUADS, FT_BLOCKS_PER_GRAPH = (DF_NM * DF_NM) / 128;
constexpr int FT_W1H = 0, FT_W1L = 4096, FT_W2H = 8192, FT_W2L = 16384;      // B tiles: [K/8 chunks][64 n][16 B]
constexpr int FT_U = 24576, FT_U_BYTES = 32768;                                // per quad: A1 hi|lo (8+8 KB) / A2 hi|lo (16+16 KB)
constexpr int FT_SMALL = FT_U + FT_QUADS * FT_U_BYTES;                         // b0[64] b1[64] w2[64] | per quad idx[64] + count
constexpr int FT_IDX = FT_SMALL + 3 * 64 * 4, FT_BARS = FT_IDX + FT_QUADS * 66 * 4, FT_TMEM = FT_BARS + FT_QUADS * 8;
constexpr size_t FT_SMEM = FT_TMEM + 16;

struct DenseChannels {
    const float* ptr[DF_FK];     // channel k of graph 0: [Nm*Nm] floats
    int64_t bstride[DF_FK];      // floats between consecutive graphs for that channel's stack
    int F;
};

__global__ void __launch_bounds__(FT_THREADS, 1)
dense_edge_final_tc_kernel(DenseChannels ch, const float* __restrict__ flags, const float* __restrict__ scale,
                           const float* __restrict__ W0, const float* __restrict__ b0, const float* __restrict__ W1,
                           const float* __restrict__ b1, const float* __restrict__ W2, const float* __restrict__ b2, int H1, int H2,
                           int B, int Nm, float* __restrict__ out, int32_t* __restrict__ status) {
    extern __shared__ __align__(128) uint8_t ft_smem[];
    const int tid = threadIdx.x, q = tid >> 7, e = tid & 127, warp = tid >> 5;
    float* b0s = reinterpret_cast<float*>(ft_smem + FT_SMALL);
    float* b1s = b0s + 64;
    float* w2s = b1s + 64;
    int* vidx = reinterpret_cast<int*>(ft_smem + FT_IDX) + q * 66;
    // ---- one-time: weights -> fp16 hi/lo B tiles (element (n, k) at (k/8) * 1024 + n * 16 + (k%8) * 2), zero padded
    for (int item = tid; item < 4 * 64; item += FT_THREADS) {
        const int kc = item >> 6, n = item & 63;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = 8 * kc + j; v[j] = (n < H1 && k < ch.F) ? W0[n * ch.F + k] : 0.0f; }
        tc05::store_chunk(ft_smem + FT_W1H, ft_smem + FT_W1L, n, kc, 1024, v);
    }
    for (int item = tid; item < 8 * 64; item += FT_THREADS) {
        const int kc = item >> 6, n = item & 63;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = 8 * kc + j; v[j] = (n < H2 && k < H1) ? W1[n * H1 + k] : 0.0f; }
        tc05::store_chunk(ft_smem + FT_W2H, ft_smem + FT_W2L, n, kc, 1024, v);
    }
    if (tid < 64) {
        b0s[tid] = tid < H1 ? b0[tid] : 0.0f;
        b1s[tid] = tid < H2 ? b1[tid] : 0.0f;
        w2s[tid] = tid < H2 ? W2[tid] : 0.0f;
    }
    if (tid < FT_QUADS) tc05::mbar_init(tc05::smem_u32(ft_smem + FT_BARS + 8 * tid), 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc05::smem_u32(ft_smem + FT_TMEM)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before();
    __syncthreads();
    tc05::fence_after();
    const uint32_t tmem_q = *reinterpret_cast<const uint32_t*>(ft_smem + FT_TMEM) + q * 128;   // acc1 [0,64) | acc2 [64,128)
    const uint32_t tlane = tmem_q + (static_cast<uint32_t>(e & ~31) << 16);
    const uint32_t bar = tc05::smem_u32(ft_smem + FT_BARS + 8 * q);
    uint8_t* U = ft_smem + FT_U + q * FT_U_BYTES;
    const uint32_t u_addr = tc05::smem_u32(U);
    const uint32_t w1h = tc05::smem_u32(ft_smem + FT_W1H), w1l = tc05::smem_u32(ft_smem + FT_W1L);
    const uint32_t w2h = tc05::smem_u32(ft_smem + FT_W2H), w2l = tc05::smem_u32(ft_smem + FT_W2L);
    const int NN = Nm * Nm;
    const float bias2 = b2[0];
    uint32_t ph = 0;
    bool ok = true;
    const int items = B * FT_BLOCKS_PER_GRAPH;
    for (int w = blockIdx.x * FT_QUADS + q; w < items; w += gridDim.x * FT_QUADS) {
        const int b = w / FT_BLOCKS_PER_GRAPH, blk = w % FT_BLOCKS_PER_GRAPH;
        // compact valid-atom list of graph b (the quad's first warp), then the quad's 128 pairs t = blk*128 + e of the n x n valid ones
        tc05::group_sync(1 + q, 128);      // the previous item's readers of vidx are done
        if (e < 32) {
            const float* fl = flags + static_cast<int64_t>(b) * Nm;
            int n = 0;
            for (int base = 0; base < Nm; base += 32) {
                const int node = base + e;
                const bool okn = node < Nm && fl[node] != 0.0f;
                const unsigned m = __ballot_sync(0xffffffffu, okn);
                if (okn) vidx[n + __popc(m & ((1u << e) - 1u))] = node;
                n += __popc(m);
            }
            if (e == 0) vidx[64] = n;
        }
        tc05::group_sync(1 + q, 128);
        const int n = vidx[64];
        if (blk * 128 >= n * n) continue;            // (quad-uniform)
        const int t = blk * 128 + e;
        const bool live = t < n * n;
        const int i = live ? vidx[t / n] : 0, j = live ? vidx[t % n] : 0, p = i * Nm + j;
        // ---- A1 row: the pair's channel values
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            float v[8];
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int k = 8 * kc + jj;
                v[jj] = (live && k < ch.F) ? __ldg(ch.ptr[k] + b * ch.bstride[k] + p) : 0.0f;
            }
            tc05::store_chunk(U, U + 8192, e, kc, 2048, v);
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before();
        tc05::group_sync(1 + q, 128);
        if (e == 0) {
            tc05::fence_after();
            tc05::mma_split_f16<64, 2>(tmem_q, u_addr, u_addr + 8192, w1h, w1l, 0u);
            tc05::commit(bar);
        }
        ok &= tc05::mbar_wait(bar, ph);
        ph ^= 1u;
        tc05::fence_after();
        // ---- epilogue 1: h1 = silu(acc + b0) -> A2 row [128 x 64] (overlays A1: its MMAs are complete)
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            float hv[32];
            tc05::tmem_ld32(tlane + 32 * cb, hv);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) hv[jj] = df_silu(hv[jj] + b0s[32 * cb + jj]);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) tc05::store_chunk(U, U + 16384, e, 4 * cb + k4, 2048, hv + 8 * k4);
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before();
        tc05::group_sync(1 + q, 128);
        if (e == 0) {
            tc05::fence_after();
            tc05::mma_split_f16<64, 4>(tmem_q + 64, u_addr, u_addr + 16384, w2h, w2l, 0u);
            tc05::commit(bar);
        }
        ok &= tc05::mbar_wait(bar, ph);
        ph ^= 1u;
        tc05::fence_after();
        // ---- epilogue 2: r = b2 + sum_k silu(acc + b1[k]) w2[k];  out = ((r f_i) f_j) scale_b, zero diagonal (left to the zero fill)
        float r = bias2;
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
            float hv[32];
            tc05::tmem_ld32(tlane + 64 + 32 * cb, hv);
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) r = fmaf(df_silu(hv[jj] + b1s[32 * cb + jj]), w2s[32 * cb + jj], r);
        }
        tc05::fence_before();
        if (live && i != j) {
            const float v = (r * flags[b * Nm + i]) * flags[b * Nm + j];
            out[static_cast<int64_t>(b) * NN + p] = scale ? v * scale[b] : v;
        }
    }
    if (!ok && status) *status = 1;
    tc05::fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*reinterpret_cast<const uint32_t*>(ft_smem + FT_TMEM)), "r"(512u));
}

}  // namespace molsde

using namespace molsde;

extern "C" {

int molsde_dense_attn_sym(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, const float* flags, int32_t B,
                          int32_t C, int32_t Nm, float* S, void* stream) {
    if (!Q || !K || !flags || !S || B <= 0 || C <= 0 || Nm <= 0 || W <= 0 || ds <= 0 || W % ds) return MOLSDE_ERR_INVALID;
    if (Nm > DF_NM || W > 32 || (ds & 3) || (ldq & 3) || (reinterpret_cast<uintptr_t>(Q) & 15) || (reinterpret_cast<uintptr_t>(K) & 15))
        return MOLSDE_ERR_UNSUPPORTED;
    dense_attn_sym_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(Q, K, ldq, W, ds, flags, C, Nm, S);
    return check_launch("dense_attn_sym");
}

int molsde_dense_pair_mlp(const float* S, const float* adjc, const float* flags, const float* W0, const float* b0, const float* W1,
                          const float* b1, const float* W2, const float* b2, int32_t B, int32_t Cin, int32_t Hd, int32_t Co,
                          int32_t Nm, int32_t symmetric, float* adjc_next, void* stream) {
    if (!S || !adjc || !flags || !W0 || !b0 || !W1 || !b1 || !W2 || !b2 || !adjc_next || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    if (Nm > DF_NM || (Cin != 2 && Cin != DF_K0 / 2) || Hd < 1 || Hd > DF_H || Co < 1 || Co > DF_CO) return MOLSDE_ERR_UNSUPPORTED;
    dense_pair_mlp_kernel<<<dim3((Nm * Nm + 255) / 256, B), 256, 0, as_stream(stream)>>>(S, adjc, flags, W0, b0, W1, b1, W2, b2, Cin, Hd,
                                                                                      Co, Nm, symmetric, adjc_next);
    return check_launch("dense_pair_mlp");
}

int molsde_dense_node_side(const float* X, int64_t rows, int64_t ldx, int32_t Fin, const float* W1, const float* b1, const float* W2,
                           const float* b2, const float* Wv, int32_t G, int32_t W, int32_t NV, float* QK, int64_t ldqk, float* XW,
                           int64_t ldxw, void* stream) {
    if (!X || !W1 || !b1 || !W2 || !b2 || !Wv || !QK || !XW || rows < 0 || G < 1) return MOLSDE_ERR_INVALID;
    if (W != NS_W || Fin < 1 || Fin > NS_FIN || G > 16 || NV < 1 || NV > 256 || (ldqk & 3) || (reinterpret_cast<uintptr_t>(QK) & 15))
        return MOLSDE_ERR_UNSUPPORTED;
    if (rows == 0) return MOLSDE_OK;
    const size_t smem = sizeof(float) * (G * NS_W * NS_FIN + G * NS_W * NS_W + 2 * G * NS_W + NV * NS_FIN);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(dense_node_side_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MOLSDE_ERR_CUDA; }
        configured = smem;
    }
    int64_t ctas = (rows + NS_ROWS - 1) / NS_ROWS;
    if (ctas > 2 * kNumSMs) ctas = 2 * kNumSMs;
    dense_node_side_kernel<<<static_cast<unsigned>(ctas), NS_WARPS * 32, smem, as_stream(stream)>>>(X, rows, ldx, Fin, W1, b1, W2, b2, Wv, G, NV,
                                                                                              QK, ldqk, XW, ldxw);
    return check_launch("dense_node_side");
}

int molsde_dense_multi_channel(const float* V, int64_t rows, int32_t K, const float* W0, const float* b0, int32_t H, const float* W1,
                               const float* b1, int32_t NO, const float* rowflag, float* out, void* stream) {
    if (!V || !W0 || !b0 || !W1 || !b1 || !out || rows < 0) return MOLSDE_ERR_INVALID;
    if (K < 4 || (K & 3) || K > 256 || H < 1 || H > 16 || NO < 1 || NO > 16 || (reinterpret_cast<uintptr_t>(V) & 15)) return MOLSDE_ERR_UNSUPPORTED;
    if (rows == 0) return MOLSDE_OK;
    const size_t smem = sizeof(float) * (16 * K + 256 + 32);
    dense_multi_channel_kernel<<<static_cast<unsigned>((rows + 127) / 128), 128, smem, as_stream(stream)>>>(V, rows, K, W0, b0, H, W1, b1, NO,
                                                                                                      rowflag, out);
    return check_launch("dense_multi_channel");
}

int molsde_dense_edge_final_mlp(const float* const* seg_ptrs, const int32_t* seg_channels, int32_t nseg, const float* flags,
                                const float* scale, const float* W0, const float* b0, const float* W1, const float* b1,
                                const float* W2, const float* b2, int32_t F, int32_t H1, int32_t H2, int32_t B, int32_t Nm, float* out,
                                void* stream) {
    if (!seg_ptrs || !seg_channels || !flags || !W0 || !b0 || !W1 || !b1 || !W2 || !b2 || !out || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    if (nseg < 1 || nseg > DF_MAX_SEGS || Nm > DF_NM || F > DF_FK || H1 > DF_FH || H2 > DF_FH) return MOLSDE_ERR_UNSUPPORTED;
    DenseSegs segs;
    int tot = 0;
    for (int s = 0; s < DF_MAX_SEGS; ++s) {
        segs.ptr[s] = s < nseg ? seg_ptrs[s] : nullptr;
        segs.ch[s] = s < nseg ? seg_channels[s] : 0;
        if (s < nseg && (!seg_ptrs[s] || seg_channels[s] < 1)) return MOLSDE_ERR_INVALID;
        tot += segs.ch[s];
    }
    segs.n = nseg;
    if (tot != F) return MOLSDE_ERR_INVALID;
    static const bool no_tc = getenv("MOLSDE_DENSE_FINAL_FFMA") != nullptr;
    if (!no_tc) {   // tcgen05 path: one persistent CTA per SM
        DenseChannels ch;
        int k = 0;
        for (int s = 0; s < nseg; ++s)
            for (int c = 0; c < seg_channels[s]; ++c, ++k) {
                ch.ptr[k] = seg_ptrs[s] + static_cast<int64_t>(c) * Nm * Nm;
                ch.bstride[k] = static_cast<int64_t>(seg_channels[s]) * Nm * Nm;
            }
        for (; k < DF_FK; ++k) { ch.ptr[k] = seg_ptrs[0]; ch.bstride[k] = 0; }
        ch.F = F;
        static bool configured = false;
        if (!configured) {
            cudaError_t e = cudaFuncSetAttribute(dense_edge_final_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(FT_SMEM));
            if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MOLSDE_ERR_CUDA; }
            configured = true;
        }
        const int items = B * FT_BLOCKS_PER_GRAPH;
        const int want = (items + FT_QUADS - 1) / FT_QUADS;
        dense_edge_final_tc_kernel<<<want < kNumSMs ? want : kNumSMs, FT_THREADS, FT_SMEM, as_stream(stream)>>>(
            ch, flags, scale, W0, b0, W1, b1, W2, b2, H1, H2, B, Nm, out, nullptr);
        return check_launch("dense_edge_final_tc");
    }
    const size_t smem = sizeof(float) * (DF_FK * DF_FH + DF_FH * DF_FH + 3 * DF_FH);
    dense_edge_final_mlp_kernel<<<dim3((Nm * Nm + 255) / 256, B), 256, smem, as_stream(stream)>>>(segs, flags, scale, W0, b0, W1, b1, W2, b2, F,
                                                                                               H1, H2, Nm, out);
    return check_launch("dense_edge_final_mlp");
}

}  // extern "C"
