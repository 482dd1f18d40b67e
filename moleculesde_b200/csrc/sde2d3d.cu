// SDEModel2Dto3D_02 score network (K3) and the fused predictor-corrector reverse-SDE loop (K5).
//
// Reference path: Geom3D/models/MoleculeSDE/SDE_model_2D_to_3D.py:393-445 (get_score),
// equivariant_scorenetwork.py:121-169, and the sampler in
// examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-212.
//
// Design (B200-first):
//   * one persistent CTA (256 threads, 1 CTA/SM, ~214 KB smem) owns one "chunk" = a set of whole
//     molecules with <= 224 atoms; node state (hidden features, q/k/v, positions, score) lives in
//     shared memory for the WHOLE score evaluation -- and, in the PC kernel, for all 1000 reverse
//     steps -- so HBM sees only the initial/final positions;
//   * edges are processed in CSR-by-target order in tiles of <= 128 edges aligned to target nodes,
//     so the segment softmax / mean aggregation of a tile is self-contained and runs in a fixed,
//     atomic-free, ascending-source order (deterministic, same order as the reference scatter);
//   * every per-edge MLP is a register-tiled fp32 FFMA GEMM over the tile (A operand k-major in
//     smem, weights streamed from the packed parameter blob into smem once per phase);
//   * the per-edge attribute (32 floats) is the only per-edge state that survives between phases;
//     it goes to an L2-resident per-CTA scratch in the tile layout [tile][32][128], so re-loading
//     it is a straight 16 KB cp.async copy.
#include <math_constants.h>

#include "common.cuh"
#include "sde2d3d_params.h"

namespace molsde {

constexpr int TE = MOLSDE_TILE_EDGES;         // 128 edges per tile
constexpr int NTHREADS = 256;
constexpr int MAXN = MOLSDE_CHUNK_MAX_NODES;  // 224 atoms per chunk
constexpr int MAXT = 64;                      // tiles per chunk
constexpr float EPS = 1e-6f;                  // SDE_model_2D_to_3D.py:10
constexpr float LN_EPS = 1e-5f;

// ---- shared memory carve-up (float offsets) ----
constexpr int S_XT = 0;                      // [32][MAXN]  node hidden, k-major
constexpr int S_Q = S_XT + 32 * MAXN;        // [MAXN][32]  query  (aggregate written in place)
constexpr int S_K = S_Q + 32 * MAXN;         // [MAXN][32]
constexpr int S_V = S_K + 32 * MAXN;         // [MAXN][32]
constexpr int S_WG = S_V + 32 * MAXN;        // [7488]      weights of the current GAT layer
constexpr int S_A = S_WG + MOLSDE_P_GAT_SZ;  // [64][TE]    A operand (k-major); E0 spills 4 rows into S_M
constexpr int S_M = S_A + 64 * TE;           // [TE][32]    weighted messages / basis mix
constexpr int S_L = S_M + TE * 32;           // [TE][8]     logits / geometry scalars / dyn coeffs
constexpr int S_MS = S_L + TE * 8;           // [2][TE][8]  softmax max / sum
constexpr int S_POS = S_MS + 2 * TE * 8;     // [MAXN*3]
constexpr int S_GRAD = S_POS + MAXN * 3;     // [MAXN*3]  network output ("gradient")
constexpr int S_SCORE = S_GRAD + MAXN * 3;   // [MAXN*3]
constexpr int S_NOISE = S_SCORE + MAXN * 3;  // [MAXN*3]
constexpr int S_RED = S_NOISE + MAXN * 3;    // [64]
constexpr int S_FLOATS = S_RED + 64;
// int region (after the floats)
constexpr int SI_ROWL = 0;                 // [MAXN+1] edge offsets local to the chunk
constexpr int SI_TTGT = SI_ROWL + MAXN + 1;  // [MAXT+1] tile target boundaries local to the chunk
constexpr int SI_ESRC = SI_TTGT + MAXT + 1;  // [TE]
constexpr int SI_ETGT = SI_ESRC + TE;        // [TE]
constexpr int SI_MISC = SI_ETGT + TE;        // [4]
constexpr int S_INTS = SI_MISC + 4;
constexpr size_t SMEM_BYTES = sizeof(float) * S_FLOATS + sizeof(int) * S_INTS;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
static_assert(MOLSDE_P_E0_END <= 3 * 32 * MAXN, "E0 weights are staged in the q/k/v region");
static_assert(MOLSDE_P_BASIS_SZ <= 3 * 32 * MAXN, "basis weights are staged in the q/k/v region");

struct Chunk {
    float* sm;
    int* si;
    int n;       // atoms
    int node0;   // first global node
    int edge0;   // first global CSR edge
    int tile0;   // first global tile
    int ntiles;
};

// ---------------------------------------------------------------------------------------
// register-tiled GEMM over one tile:  acc[TM][TN] += A[k][row] * W[k][col]
//   rows  = te*TM .. te*TM+TM-1                      (edges or nodes)
//   cols  = to*4..to*4+3 (TN==4)   or additionally NOUT/2 + to*4..+3 (TN==8)
// ---------------------------------------------------------------------------------------
template <int TM, int TN, int NOUT, int LDA>
__device__ __forceinline__ void gemm_acc(const float* __restrict__ As, const float* __restrict__ Ws, int K, int te,
                                         int to, float (&acc)[TM][TN]) {
    static_assert(TM % 4 == 0 && (TN == 4 || TN == 8), "tile shape");
    const float* ap = As + te * TM;
    const float* wp = Ws + to * 4;
#pragma unroll 4
    for (int k = 0; k < K; ++k) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM / 4; ++i) {
            const float4 v = *reinterpret_cast<const float4*>(ap + k * LDA + 4 * i);
            a[4 * i] = v.x; a[4 * i + 1] = v.y; a[4 * i + 2] = v.z; a[4 * i + 3] = v.w;
        }
        {
            const float4 v = *reinterpret_cast<const float4*>(wp + k * NOUT);
            b[0] = v.x; b[1] = v.y; b[2] = v.z; b[3] = v.w;
        }
        if (TN == 8) {
            const float4 v = *reinterpret_cast<const float4*>(wp + k * NOUT + NOUT / 2);
            b[4] = v.x; b[5] = v.y; b[6] = v.z; b[7] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
}

template <int TM, int TN>
__device__ __forceinline__ void zero_acc(float (&acc)[TM][TN]) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.0f;
}

// cooperative global->shared copy of `nfloat` floats (multiple of 4, 16B aligned both sides)
__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int nfloat) {
    for (int i = threadIdx.x * 4; i < nfloat; i += NTHREADS * 4) cp_async16(dst + i, src + i);
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------
// geometry, SDE_model_2D_to_3D.py:35-47 (coord2basis) with the reference's unfused op order
// ---------------------------------------------------------------------------------------
struct Frame {
    float dx, dy, dz, cx, cy, cz, vx, vy, vz, dist;
};
__device__ __forceinline__ float dot3_rn(float a0, float a1, float a2, float b0, float b1, float b2) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a0, b0), __fmul_rn(a1, b1)), __fmul_rn(a2, b2));
}
__device__ __forceinline__ Frame coord2basis(const float* pr, const float* pc) {
    Frame f;
    float dx = __fsub_rn(pr[0], pc[0]), dy = __fsub_rn(pr[1], pc[1]), dz = __fsub_rn(pr[2], pc[2]);
    const float radial = dot3_rn(dx, dy, dz, dx, dy, dz);
    float cx = __fsub_rn(__fmul_rn(pr[1], pc[2]), __fmul_rn(pr[2], pc[1]));
    float cy = __fsub_rn(__fmul_rn(pr[2], pc[0]), __fmul_rn(pr[0], pc[2]));
    float cz = __fsub_rn(__fmul_rn(pr[0], pc[1]), __fmul_rn(pr[1], pc[0]));
    f.dist = sqrtf(radial);
    const float norm = __fadd_rn(f.dist, EPS);
    dx = __fdiv_rn(dx, norm); dy = __fdiv_rn(dy, norm); dz = __fdiv_rn(dz, norm);
    const float cnorm = __fadd_rn(sqrtf(dot3_rn(cx, cy, cz, cx, cy, cz)), EPS);
    cx = __fdiv_rn(cx, cnorm); cy = __fdiv_rn(cy, cnorm); cz = __fdiv_rn(cz, cnorm);
    f.dx = dx; f.dy = dy; f.dz = dz;
    f.cx = cx; f.cy = cy; f.cz = cz;
    f.vx = __fsub_rn(__fmul_rn(dy, cz), __fmul_rn(dz, cy));
    f.vy = __fsub_rn(__fmul_rn(dz, cx), __fmul_rn(dx, cz));
    f.vz = __fsub_rn(__fmul_rn(dx, cy), __fmul_rn(dy, cx));
    return f;
}

// per-tile edge bookkeeping: local source / target of every slot, returns #edges in the tile
__device__ __forceinline__ int build_tile_edges(const Chunk& c, const int32_t* __restrict__ src_g, int t, int& ta,
                                                int& tb, int& ea) {
    const int* rowl = c.si + SI_ROWL;
    const int* ttgt = c.si + SI_TTGT;
    int* esrc = c.si + SI_ESRC;
    int* etgt = c.si + SI_ETGT;
    ta = ttgt[t];
    tb = ttgt[t + 1];
    ea = rowl[ta];
    const int ne = rowl[tb] - ea;
    for (int i = ta + threadIdx.x; i < tb; i += NTHREADS) {
        for (int e = rowl[i]; e < rowl[i + 1]; ++e) {
            etgt[e - ea] = i;
            esrc[e - ea] = src_g[c.edge0 + e] - c.node0;
        }
    }
    for (int s = ne + threadIdx.x; s < TE; s += NTHREADS) { esrc[s] = 0; etgt[s] = 0; }
    return ne;
}

// sin/cos Fourier features of one scalar per edge into A rows [row0, row0+64):
// GaussianFourierProjection.forward, SDE_model_2D_to_3D.py:64-66  (x * W * 2 * pi, fp32, in that order)
__device__ __forceinline__ void fill_fourier(float* A, int row0, const float* __restrict__ xs, const float* __restrict__ W) {
    const int edge = threadIdx.x & (TE - 1);
    const int w0 = threadIdx.x >> 7;  // 0 or 1
    const float x = xs[edge];
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
        const int w = w0 + 2 * it;
        const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, W[w]), 2.0f), 3.14159274101257324f);
        float s, co;
        sincosf(arg, &s, &co);
        A[(row0 + w) * TE + edge] = s;
        A[(row0 + 32 + w) * TE + edge] = co;
    }
}

// ---------------------------------------------------------------------------------------
// Phase E0: per-edge attribute  edge_attr = input_mlp(gfp(d)) * e2d + project([sin,cos,emb_i,emb_j])
// SDE_model_2D_to_3D.py:402-432
// ---------------------------------------------------------------------------------------
__device__ void phase_edge_features(const Chunk& c, const float* __restrict__ blob, const int32_t* __restrict__ src_g,
                                    const float* __restrict__ e2d_tiles, float* __restrict__ scratch) {
    float* sm = c.sm;
    float* W = sm + S_Q;  // E0 weights staged over the (currently dead) q/k/v region
    float* A = sm + S_A;
    float* geo = sm + S_L;  // [7][TE]: d, ci0, ci2, cj0, cj2, psin, pcos
    const float* pos = sm + S_POS;
    const int tid = threadIdx.x;
    const int to = tid & 7, te = tid >> 3;
    stage_async(W, blob, MOLSDE_P_E0_END);
    cp_async_wait<0>();
    __syncthreads();
    for (int t = 0; t < c.ntiles; ++t) {
        int ta, tb, ea;
        const int ne = build_tile_edges(c, src_g, t, ta, tb, ea);
        __syncthreads();
        if (tid < TE) {
            float g[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (tid < ne) {
                const float* pr = pos + 3 * (c.si + SI_ESRC)[tid];  // row = source j
                const float* pc = pos + 3 * (c.si + SI_ETGT)[tid];  // col = target i
                const Frame f = coord2basis(pr, pc);
                // coff = edge_basis @ r  (:417-418), |.| on component 1 (:419-420)
                const float ci0 = dot3_rn(f.dx, f.dy, f.dz, pr[0], pr[1], pr[2]);
                const float ci1 = fabsf(dot3_rn(f.cx, f.cy, f.cz, pr[0], pr[1], pr[2]));
                const float ci2 = dot3_rn(f.vx, f.vy, f.vz, pr[0], pr[1], pr[2]);
                const float cj0 = dot3_rn(f.dx, f.dy, f.dz, pc[0], pc[1], pc[2]);
                const float cj1 = fabsf(dot3_rn(f.cx, f.cy, f.cz, pc[0], pc[1], pc[2]));
                const float cj2 = dot3_rn(f.vx, f.vy, f.vz, pc[0], pc[1], pc[2]);
                const float ni = sqrtf(dot3_rn(ci0, ci1, ci2, ci0, ci1, ci2));
                const float nj = sqrtf(dot3_rn(cj0, cj1, cj2, cj0, cj1, cj2));
                const float pcos = __fdiv_rn(__fdiv_rn(dot3_rn(ci0, ci1, ci2, cj0, cj1, cj2), __fadd_rn(ni, EPS)),
                                             __fadd_rn(nj, EPS));
                const float psin = sqrtf(__fsub_rn(1.0f, __fmul_rn(pcos, pcos)));  // :425 (NaN if |cos|>1, as the reference)
                g[0] = f.dist; g[1] = ci0; g[2] = ci2; g[3] = cj0; g[4] = cj2; g[5] = psin; g[6] = pcos;
            }
#pragma unroll
            for (int q = 0; q < 7; ++q) geo[q * TE + tid] = g[q];
        }
        __syncthreads();
        // ---- edge_attr_3D_invariant = input_mlp(gfp_dist(d))  (:409-410) ----
        float inv[4][4];
        zero_acc(inv);
        fill_fourier(A, 0, geo, W + MOLSDE_P_GFP_DIST_W);
        __syncthreads();
        gemm_acc<4, 4, 32, TE>(A, W + MOLSDE_P_IN_WT, 64, te, to, inv);
        __syncthreads();
        // ---- embed_i / embed_j = coff_mlp([gfp(c0), gfp(c2)])  (:297-304, 427-428) ----
        float emb[2][4][4];
#pragma unroll
        for (int side = 0; side < 2; ++side) {
            zero_acc(emb[side]);
#pragma unroll
            for (int comp = 0; comp < 2; ++comp) {
                fill_fourier(A, 0, geo + (1 + 2 * side + comp) * TE, W + MOLSDE_P_GFP_COFF_W);
                __syncthreads();
                gemm_acc<4, 4, 32, TE>(A, W + MOLSDE_P_COFF_WT + comp * 64 * 32, 64, te, to, emb[side]);
                __syncthreads();
            }
        }
        // ---- project: Linear(66,32) silu Linear(32,32) on [psin, pcos, emb_i, emb_j]  (:429-430) ----
        if (tid < TE) {
            A[0 * TE + tid] = geo[5 * TE + tid];
            A[1 * TE + tid] = geo[6 * TE + tid];
            A[66 * TE + tid] = 0.0f;
            A[67 * TE + tid] = 0.0f;
        }
#pragma unroll
        for (int side = 0; side < 2; ++side)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float bj = W[MOLSDE_P_COFF_B + to * 4 + j];
                float4 v = make_float4(emb[side][0][j] + bj, emb[side][1][j] + bj, emb[side][2][j] + bj,
                                       emb[side][3][j] + bj);
                *reinterpret_cast<float4*>(&A[(2 + 32 * side + to * 4 + j) * TE + te * 4]) = v;
            }
        __syncthreads();
        float h[4][4];
        zero_acc(h);
        gemm_acc<4, 4, 32, TE>(A, W + MOLSDE_P_PROJ0_WT, 68, te, to, h);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float bj = W[MOLSDE_P_PROJ0_B + to * 4 + j];
            float4 v = make_float4(silu_f(h[0][j] + bj), silu_f(h[1][j] + bj), silu_f(h[2][j] + bj),
                                   silu_f(h[3][j] + bj));
            *reinterpret_cast<float4*>(&A[(to * 4 + j) * TE + te * 4]) = v;
        }
        __syncthreads();
        float fr[4][4];
        zero_acc(fr);
        gemm_acc<4, 4, 32, TE>(A, W + MOLSDE_P_PROJ1_WT, 32, te, to, fr);
        // ---- edge_attr = inv3d * e2d + frame  (:432) -> scratch tile ----
        const float* e2d_t = e2d_tiles + static_cast<size_t>(c.tile0 + t) * (32 * TE);
        float* sc_t = scratch + static_cast<size_t>(t) * (32 * TE);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = to * 4 + j;
            const float bi = W[MOLSDE_P_IN_B + col], bf = W[MOLSDE_P_PROJ1_B + col];
            const float4 e2 = __ldg(reinterpret_cast<const float4*>(e2d_t + col * TE + te * 4));
            float4 o;
            o.x = fmaf(inv[0][j] + bi, e2.x, fr[0][j] + bf);
            o.y = fmaf(inv[1][j] + bi, e2.y, fr[1][j] + bf);
            o.z = fmaf(inv[2][j] + bi, e2.z, fr[2][j] + bf);
            o.w = fmaf(inv[3][j] + bi, e2.w, fr[3][j] + bf);
            *reinterpret_cast<float4*>(sc_t + col * TE + te * 4) = o;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// GAT layer pieces  (equivariant_scorenetwork.py:34-40, TransformerConv heads=8 C=4)
// ---------------------------------------------------------------------------------------
__device__ void node_qkv(const Chunk& c) {
    float* sm = c.sm;
    const float* Wg = sm + S_WG;
    const int tid = threadIdx.x, to = tid & 7, te = tid >> 3;
    for (int nt = 0; nt * TE < c.n; ++nt) {
#pragma unroll
        for (int which = 0; which < 3; ++which) {
            float acc[4][4];
            zero_acc(acc);
            gemm_acc<4, 4, 32, MAXN>(sm + S_XT + nt * TE, Wg + MOLSDE_G_WQ_T + which * 1024, 32, te, to, acc);
            const float4 b = *reinterpret_cast<const float4*>(Wg + MOLSDE_G_BQ + which * 32 + to * 4);
            float* dst = sm + S_Q + which * 32 * MAXN;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int node = nt * TE + te * 4 + i;
                if (node < c.n)
                    *reinterpret_cast<float4*>(dst + node * 32 + to * 4) =
                        make_float4(acc[i][0] + b.x, acc[i][1] + b.y, acc[i][2] + b.z, acc[i][3] + b.w);
            }
        }
    }
}

// attention over the incoming edges of every target: logits, segment softmax (+1e-16), weighted
// messages, deterministic ascending-source sum; the aggregate overwrites q[target].
__device__ void gat_edge_phase(const Chunk& c, const int32_t* __restrict__ src_g, const float* __restrict__ scratch) {
    float* sm = c.sm;
    float* A = sm + S_A;
    float* Mm = sm + S_M;
    float* L = sm + S_L;
    float* smax = sm + S_MS;
    float* ssum = sm + S_MS + TE * 8;
    float* Q = sm + S_Q;
    const float* Kk = sm + S_K;
    const float* V = sm + S_V;
    const float* Wg = sm + S_WG;
    const int* rowl = c.si + SI_ROWL;
    const int* esrc = c.si + SI_ESRC;
    const int* etgt = c.si + SI_ETGT;
    const int tid = threadIdx.x, to = tid & 7, te = tid >> 3;
    for (int t = 0; t < c.ntiles; ++t) {
        int ta, tb, ea;
        stage_async(A, scratch + static_cast<size_t>(t) * (32 * TE), 32 * TE);
        const int ne = build_tile_edges(c, src_g, t, ta, tb, ea);
        cp_async_wait<0>();
        __syncthreads();
        // e = lin_edge(edge_attr): thread holds head `to` of 4 consecutive edges
        float e[4][4];
        zero_acc(e);
        gemm_acc<4, 4, 32, TE>(A, Wg + MOLSDE_G_WE_T, 32, te, to, e);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = te * 4 + i;
            if (s < ne) {
                const float4 k4 = *reinterpret_cast<const float4*>(Kk + esrc[s] * 32 + to * 4);
                const float4 q4 = *reinterpret_cast<const float4*>(Q + etgt[s] * 32 + to * 4);
                const float4 v4 = *reinterpret_cast<const float4*>(V + esrc[s] * 32 + to * 4);
                // alpha = (q_i . (k_j + e)) / sqrt(C)
                float lg = q4.x * (k4.x + e[i][0]);
                lg = fmaf(q4.y, k4.y + e[i][1], lg);
                lg = fmaf(q4.z, k4.z + e[i][2], lg);
                lg = fmaf(q4.w, k4.w + e[i][3], lg);
                L[s * 8 + to] = lg * 0.5f;
                e[i][0] += v4.x; e[i][1] += v4.y; e[i][2] += v4.z; e[i][3] += v4.w;  // v_j + e
            }
        }
        __syncthreads();
        // per (target, head): max and sum(exp) over the target's contiguous edge segment
        const int ntg = tb - ta;
        for (int p = tid; p < ntg * 8; p += NTHREADS) {
            const int i = ta + (p >> 3), hd = p & 7;
            const int s0 = rowl[i] - ea, s1 = rowl[i + 1] - ea;
            float m = -CUDART_INF_F;
            for (int s = s0; s < s1; ++s) m = fmaxf(m, L[s * 8 + hd]);
            float z = 0.0f;
            for (int s = s0; s < s1; ++s) z += expf(L[s * 8 + hd] - m);
            smax[p] = m;
            ssum[p] = z;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = te * 4 + i;
            if (s < ne) {
                const int p = (etgt[s] - ta) * 8 + to;
                const float a = __fdiv_rn(expf(L[s * 8 + to] - smax[p]), ssum[p] + 1e-16f);
                *reinterpret_cast<float4*>(Mm + s * 32 + to * 4) =
                    make_float4(e[i][0] * a, e[i][1] * a, e[i][2] * a, e[i][3] * a);
            }
        }
        __syncthreads();
        for (int p = tid; p < ntg * 32; p += NTHREADS) {
            const int i = ta + (p >> 5), col = p & 31;
            const int s0 = rowl[i] - ea, s1 = rowl[i + 1] - ea;
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc += Mm[s * 32 + col];
            Q[i * 32 + col] = acc;
        }
        __syncthreads();
    }
}

// LayerNorm over the 32 columns of a row held by the 8 `to` lanes (4 columns each)
__device__ __forceinline__ void layer_norm_rows(float (&v)[4][4], const float* __restrict__ w, const float* __restrict__ b,
                                                int to) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + to * 4);
    const float4 b4 = *reinterpret_cast<const float4*>(b + to * 4);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float s = (v[i][0] + v[i][1]) + (v[i][2] + v[i][3]);
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        s += __shfl_xor_sync(0xffffffffu, s, 4);
        const float mean = s * (1.0f / 32.0f);
        const float d0 = v[i][0] - mean, d1 = v[i][1] - mean, d2 = v[i][2] - mean, d3 = v[i][3] - mean;
        float q = (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        q += __shfl_xor_sync(0xffffffffu, q, 4);
        const float rstd = 1.0f / sqrtf(q * (1.0f / 32.0f) + LN_EPS);
        v[i][0] = d0 * rstd * w4.x + b4.x;
        v[i][1] = d1 * rstd * w4.y + b4.y;
        v[i][2] = d2 * rstd * w4.z + b4.z;
        v[i][3] = d3 * rstd * w4.w + b4.w;
    }
}

// x <- x + LN1(agg + skip(x));  x <- x + LN2(FFN(x));  optional SiLU  (equivariant_scorenetwork.py:35-38,140-141)
__device__ void node_update(const Chunk& c, bool silu_after) {
    float* sm = c.sm;
    float* XT = sm + S_XT;
    float* A = sm + S_A;
    const float* Q = sm + S_Q;
    const float* Wg = sm + S_WG;
    const int tid = threadIdx.x, to = tid & 7, te = tid >> 3;
    for (int nt = 0; nt * TE < c.n; ++nt) {
        float acc[4][4], x1[4][4];
        zero_acc(acc);
        gemm_acc<4, 4, 32, MAXN>(XT + nt * TE, Wg + MOLSDE_G_WS_T, 32, te, to, acc);
        const float4 bs = *reinterpret_cast<const float4*>(Wg + MOLSDE_G_BS + to * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int node = nt * TE + te * 4 + i;
            float4 ag = make_float4(0.f, 0.f, 0.f, 0.f);
            if (node < c.n) ag = *reinterpret_cast<const float4*>(Q + node * 32 + to * 4);
            acc[i][0] += bs.x + ag.x; acc[i][1] += bs.y + ag.y; acc[i][2] += bs.z + ag.z; acc[i][3] += bs.w + ag.w;
        }
        layer_norm_rows(acc, Wg + MOLSDE_G_LN1_W, Wg + MOLSDE_G_LN1_B, to);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int node = nt * TE + te * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float xo = (node < c.n) ? XT[(to * 4 + j) * MAXN + node] : 0.0f;
                x1[i][j] = xo + acc[i][j];
            }
        }
        // FFN: Linear silu Linear on x1 (A operand staged k-major)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(&A[(to * 4 + j) * TE + te * 4]) = make_float4(x1[0][j], x1[1][j], x1[2][j], x1[3][j]);
        __syncthreads();
        zero_acc(acc);
        gemm_acc<4, 4, 32, TE>(A, Wg + MOLSDE_G_F0_WT, 32, te, to, acc);
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float bj = Wg[MOLSDE_G_F0_B + to * 4 + j];
            *reinterpret_cast<float4*>(&A[(to * 4 + j) * TE + te * 4]) =
                make_float4(silu_f(acc[0][j] + bj), silu_f(acc[1][j] + bj), silu_f(acc[2][j] + bj), silu_f(acc[3][j] + bj));
        }
        __syncthreads();
        zero_acc(acc);
        gemm_acc<4, 4, 32, TE>(A, Wg + MOLSDE_G_F3_WT, 32, te, to, acc);
        const float4 b3 = *reinterpret_cast<const float4*>(Wg + MOLSDE_G_F3_B + to * 4);
#pragma unroll
        for (int i = 0; i < 4; ++i) { acc[i][0] += b3.x; acc[i][1] += b3.y; acc[i][2] += b3.z; acc[i][3] += b3.w; }
        layer_norm_rows(acc, Wg + MOLSDE_G_LN2_W, Wg + MOLSDE_G_LN2_B, to);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int node = nt * TE + te * 4 + i;
            if (node < c.n) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    float x2 = x1[i][j] + acc[i][j];
                    if (silu_after) x2 = silu_f(x2);
                    XT[(to * 4 + j) * MAXN + node] = x2;
                }
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// basis MLP + equivariant mean aggregation  (equivariant_scorenetwork.py:154-164)
// ---------------------------------------------------------------------------------------
__device__ void phase_basis(const Chunk& c, const float* __restrict__ blob, const int32_t* __restrict__ src_g,
                            const float* __restrict__ scratch, int module) {
    float* sm = c.sm;
    float* Wb = sm + S_Q;  // staged over q/k/v (dead between GAT blocks)
    float* A = sm + S_A;
    float* mix = sm + S_M;   // [TE][4]
    float* dyn = sm + S_L;   // [TE][4]
    const float* XT = sm + S_XT;
    const float* pos = sm + S_POS;
    float* grad = sm + S_GRAD;
    const int* rowl = c.si + SI_ROWL;
    const int* esrc = c.si + SI_ESRC;
    const int* etgt = c.si + SI_ETGT;
    const int tid = threadIdx.x;
    const int to = tid & 15, te = tid >> 4;  // 16 column groups x 16 edge groups, 8x8 micro-tile
    stage_async(Wb, blob + MOLSDE_P_BASIS0 + module * MOLSDE_P_BASIS_SZ, MOLSDE_P_BASIS_SZ);
    cp_async_wait<0>();
    __syncthreads();
    for (int t = 0; t < c.ntiles; ++t) {
        int ta, tb, ea;
        stage_async(A + 32 * TE, scratch + static_cast<size_t>(t) * (32 * TE), 32 * TE);
        const int ne = build_tile_edges(c, src_g, t, ta, tb, ea);
        __syncthreads();
        {   // rows 0..31: h_row + h_col
            const int edge = tid & (TE - 1), k0 = (tid >> 7) * 16;
            const int sj = esrc[edge], ti = etgt[edge];
            const bool live = edge < ne;
#pragma unroll 4
            for (int k = k0; k < k0 + 16; ++k) A[k * TE + edge] = live ? XT[k * MAXN + sj] + XT[k * MAXN + ti] : 0.0f;
        }
        cp_async_wait<0>();
        __syncthreads();
        float acc[8][8];
        zero_acc(acc);
        gemm_acc<8, 8, 128, TE>(A, Wb + MOLSDE_B_W1_T, 64, te, to, acc);
        float part[8][3];
#pragma unroll
        for (int i = 0; i < 8; ++i) part[i][0] = part[i][1] = part[i][2] = 0.0f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = (j < 4) ? (to * 4 + j) : (64 + to * 4 + (j - 4));
            const float b1 = Wb[MOLSDE_B_B1 + col];
            const float w0 = Wb[MOLSDE_B_W2 + col], w1 = Wb[MOLSDE_B_W2 + 128 + col], w2 = Wb[MOLSDE_B_W2 + 256 + col];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const float hv = silu_f(acc[i][j] + b1);
                part[i][0] = fmaf(hv, w0, part[i][0]);
                part[i][1] = fmaf(hv, w1, part[i][1]);
                part[i][2] = fmaf(hv, w2, part[i][2]);
            }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int o = 0; o < 3; ++o) {
                float v = part[i][o];
                v += __shfl_xor_sync(0xffffffffu, v, 1);
                v += __shfl_xor_sync(0xffffffffu, v, 2);
                v += __shfl_xor_sync(0xffffffffu, v, 4);
                v += __shfl_xor_sync(0xffffffffu, v, 8);
                part[i][o] = v;
            }
        if (to == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int o = 0; o < 3; ++o) dyn[(te * 8 + i) * 4 + o] = part[i][o] + Wb[MOLSDE_B_B2 + o];
        }
        __syncthreads();
        if (tid < ne) {
            const Frame f = coord2basis(pos + 3 * esrc[tid], pos + 3 * etgt[tid]);
            const float d0 = dyn[tid * 4], d1 = dyn[tid * 4 + 1], d2 = dyn[tid * 4 + 2];
            mix[tid * 4 + 0] = d0 * f.dx + d1 * f.cx + d2 * f.vx;
            mix[tid * 4 + 1] = d0 * f.dy + d1 * f.cy + d2 * f.vy;
            mix[tid * 4 + 2] = d0 * f.dz + d1 * f.cz + d2 * f.vz;
        }
        __syncthreads();
        const int ntg = tb - ta;
        for (int p = tid; p < ntg * 3; p += NTHREADS) {
            const int i = ta + p / 3, ax = p % 3;
            const int s0 = rowl[i] - ea, s1 = rowl[i + 1] - ea;
            float s = 0.0f;
            for (int q = s0; q < s1; ++q) s += mix[q * 4 + ax];
            s = __fdiv_rn(s, static_cast<float>(max(s1 - s0, 1)));  // aggr='mean'
            grad[i * 3 + ax] = (module == 0) ? s : grad[i * 3 + ax] + s;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------
// one full network evaluation on the chunk: positions in smem -> "gradient" in smem
// ---------------------------------------------------------------------------------------
__device__ void score_eval(const Chunk& c, const float* __restrict__ blob, const int32_t* __restrict__ src_g,
                           const float* __restrict__ nattr, const float* __restrict__ e2d_tiles,
                           float* __restrict__ scratch) {
    float* sm = c.sm;
    phase_edge_features(c, blob, src_g, e2d_tiles, scratch);
    // conv_input = node_attr (loop-invariant node_emb output), k-major
    for (int idx = threadIdx.x; idx < c.n * 32; idx += NTHREADS) {
        const int node = idx >> 5, k = idx & 31;
        sm[S_XT + k * MAXN + node] = __ldg(nattr + static_cast<size_t>(c.node0 + node) * 32 + k);
    }
    __syncthreads();
    for (int module = 0; module < 2; ++module) {
        for (int conv = 0; conv < 2; ++conv) {
            stage_async(sm + S_WG, blob + MOLSDE_P_GAT0 + (2 * module + conv) * MOLSDE_P_GAT_SZ, MOLSDE_P_GAT_SZ);
            cp_async_wait<0>();
            __syncthreads();
            node_qkv(c);
            __syncthreads();
            gat_edge_phase(c, src_g, scratch);
            node_update(c, conv == 0);
        }
        phase_basis(c, blob, src_g, scratch, module);
    }
}

__device__ __forceinline__ bool load_chunk(Chunk& c, const molsde_plan& plan, int chunk, int32_t* status_flag) {
    const int tid = threadIdx.x;
    c.tile0 = plan.chunk_tile_ptr[chunk];
    c.ntiles = plan.chunk_tile_ptr[chunk + 1] - c.tile0;
    c.node0 = plan.tile_tgt_ptr[c.tile0];
    c.n = plan.tile_tgt_ptr[c.tile0 + c.ntiles] - c.node0;
    c.edge0 = plan.rowptr[c.node0];
    bool ok = (c.n <= MAXN) && (c.ntiles <= MAXT) && (c.n >= 0);
    if (ok) {
        int* rowl = c.si + SI_ROWL;
        int* ttgt = c.si + SI_TTGT;
        for (int i = tid; i <= c.n; i += NTHREADS) rowl[i] = plan.rowptr[c.node0 + i] - c.edge0;
        for (int t = tid; t <= c.ntiles; t += NTHREADS) ttgt[t] = plan.tile_tgt_ptr[c.tile0 + t] - c.node0;
        __syncthreads();
        int bad = 0;
        for (int t = tid; t < c.ntiles; t += NTHREADS)
            if (rowl[ttgt[t + 1]] - rowl[ttgt[t]] > TE) bad = 1;
        ok = !__syncthreads_or(bad);
    }
    if (!ok && tid == 0 && status_flag) atomicExch(status_flag, 1 + chunk);
    return ok;
}

// ---------------------------------------------------------------------------------------
// K3: get_score for a whole batch, SDE_model_2D_to_3D.py:393-445
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
sde2d3d_score_kernel(molsde_plan plan, const float* __restrict__ blob, const float* __restrict__ nattr,
                     const float* __restrict__ e2d_tiles, const float* __restrict__ pos,
                     const float* __restrict__ stdv, float* __restrict__ score, float* __restrict__ scratch,
                     int64_t scratch_stride, int32_t* status_flag) {
    extern __shared__ __align__(16) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(smem + S_FLOATS);
    float* my_scratch = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
    for (int chunk = blockIdx.x; chunk < plan.num_chunks; chunk += gridDim.x) {
        __syncthreads();
        if (!load_chunk(c, plan, chunk, status_flag)) continue;
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) smem[S_POS + i] = pos[static_cast<size_t>(c.node0) * 3 + i];
        __syncthreads();
        score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch);
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) {
            // scores = -output / std  (:440-443)
            score[static_cast<size_t>(c.node0) * 3 + i] = __fdiv_rn(-smem[S_GRAD + i], stdv[c.node0 + i / 3]);
        }
    }
}

// ---------------------------------------------------------------------------------------
// Philox4x32-10 + Box-Muller (throughput-mode noise)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t (&ctr)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr[0]), lo0 = 0xD2511F53u * ctr[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr[2]), lo1 = 0xCD9E8D57u * ctr[2];
        const uint32_t n0 = hi1 ^ ctr[1] ^ k0, n1 = lo1, n2 = hi0 ^ ctr[3] ^ k1, n3 = lo0;
        ctr[0] = n0; ctr[1] = n1; ctr[2] = n2; ctr[3] = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
}
__device__ __forceinline__ void normal3(uint64_t seed, uint32_t node, uint32_t step, uint32_t stream, float* out) {
    uint32_t ctr[4] = {node, step, stream, 0x5DEu};
    philox4x32_10(ctr, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    const float u1 = (static_cast<float>(ctr[0] >> 8) + 1.0f) * (1.0f / 16777216.0f);  // (0,1]
    const float u2 = static_cast<float>(ctr[1] >> 8) * (1.0f / 16777216.0f);
    const float u3 = (static_cast<float>(ctr[2] >> 8) + 1.0f) * (1.0f / 16777216.0f);
    const float u4 = static_cast<float>(ctr[3] >> 8) * (1.0f / 16777216.0f);
    const float r1 = sqrtf(-2.0f * logf(u1)), r2 = sqrtf(-2.0f * logf(u3));
    float s1, c1, s2, c2;
    sincospif(2.0f * u2, &s1, &c1);
    sincospif(2.0f * u4, &s2, &c2);
    out[0] = r1 * c1; out[1] = r1 * s1; out[2] = r2 * c2;
    (void)s2;
}

// deterministic block sum of `v` (one value per thread): warp shuffle tree, then a fixed-order pass
__device__ __forceinline__ float block_sum(float v, float* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.0f;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; ++w) s += red[w];
    return s;
}

// ---------------------------------------------------------------------------------------
// K5: position_PC_generation -- all reverse steps of one sampling group inside one CTA
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
sde2d3d_pc_kernel(molsde_plan plan, const float* __restrict__ blob, const float* __restrict__ nattr,
                  const float* __restrict__ e2d_tiles, const float* __restrict__ pos_init,
                  const float* __restrict__ step_table, molsde_pc_config cfg, const float* __restrict__ noise_corr,
                  const float* __restrict__ noise_pred, float* __restrict__ pos_out, float* __restrict__ pos_mean_out,
                  float* __restrict__ scratch, int64_t scratch_stride, int32_t* work_counter, int32_t* status_flag) {
    extern __shared__ __align__(16) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(smem + S_FLOATS);
    int* misc = c.si + SI_MISC;
    float* my_scratch = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
    float* P = smem + S_POS;
    float* G = smem + S_GRAD;
    float* SC = smem + S_SCORE;
    float* NZ = smem + S_NOISE;
    float* red = smem + S_RED;
    const int tid = threadIdx.x;
    const size_t N3 = static_cast<size_t>(plan.N) * 3;
    while (true) {
        __syncthreads();
        if (tid == 0) misc[0] = atomicAdd(work_counter, 1);
        __syncthreads();
        const int chunk = misc[0];
        if (chunk >= plan.num_chunks) break;
        if (!load_chunk(c, plan, chunk, status_flag)) continue;
        const int n3 = c.n * 3;
        const size_t g3 = static_cast<size_t>(c.node0) * 3;
        for (int i = tid; i < n3; i += NTHREADS) P[i] = pos_init[g3 + i];
        __syncthreads();
        for (int step = 0; step < cfg.steps; ++step) {
            const float stdv = step_table[step * 8 + 0], Gd = step_table[step * 8 + 1];
            const float sqrt_alpha = step_table[step * 8 + 2], calpha = step_table[step * 8 + 3];
            // ---------------- corrector (LangevinCorrector.update_fn :191-212) ----------------
            score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch);
            for (int i = tid; i < n3; i += NTHREADS) SC[i] = __fdiv_rn(-G[i], stdv);
            if (noise_corr) {
                for (int i = tid; i < n3; i += NTHREADS) NZ[i] = noise_corr[static_cast<size_t>(step) * N3 + g3 + i];
            } else {
                for (int i = tid; i < c.n; i += NTHREADS) normal3(cfg.seed, c.node0 + i, step, 0u, NZ + 3 * i);
            }
            __syncthreads();
            float gn = 0.0f, nn = 0.0f;
            for (int i = tid; i < c.n; i += NTHREADS) {
                gn += sqrtf(SC[3 * i] * SC[3 * i] + SC[3 * i + 1] * SC[3 * i + 1] + SC[3 * i + 2] * SC[3 * i + 2]);
                nn += sqrtf(NZ[3 * i] * NZ[3 * i] + NZ[3 * i + 1] * NZ[3 * i + 1] + NZ[3 * i + 2] * NZ[3 * i + 2]);
            }
            gn = block_sum(gn, red) / static_cast<float>(c.n);
            nn = block_sum(nn, red) / static_cast<float>(c.n);
            // step_size = (snr * noise_norm / grad_norm)^2 * 2 * alpha   (:209)
            const float ratio = __fdiv_rn(__fmul_rn(cfg.snr, nn), gn);
            const float step_size = __fmul_rn(__fmul_rn(__fmul_rn(ratio, ratio), 2.0f), calpha);
            const float nscale = sqrtf(__fmul_rn(step_size, 2.0f));
            for (int i = tid; i < n3; i += NTHREADS) {
                const float xm = __fadd_rn(P[i], __fmul_rn(step_size, SC[i]));                // :210
                P[i] = __fadd_rn(xm, __fmul_rn(__fmul_rn(nscale, NZ[i]), cfg.scale_eps));     // :211
            }
            __syncthreads();
            // ---------------- predictor (ReverseDiffusionPredictor.update_fn :163-168) ----------------
            score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch);
            const bool last = (step == cfg.steps - 1);
            for (int i = tid; i < c.n; i += NTHREADS) {
                float nz[3];
                if (noise_pred) {
                    nz[0] = noise_pred[static_cast<size_t>(step) * N3 + g3 + 3 * i];
                    nz[1] = noise_pred[static_cast<size_t>(step) * N3 + g3 + 3 * i + 1];
                    nz[2] = noise_pred[static_cast<size_t>(step) * N3 + g3 + 3 * i + 2];
                } else {
                    normal3(cfg.seed, c.node0 + i, step, 1u, nz);
                }
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    const float x = P[3 * i + a];
                    const float sc = __fdiv_rn(-G[3 * i + a], stdv);
                    const float f = __fsub_rn(__fmul_rn(sqrt_alpha, x), x);                    // SDE_sparse.py:160 / 220
                    const float rev_f = __fsub_rn(f, __fmul_rn(__fmul_rn(Gd, Gd), sc));        // SDE_sparse.py:98
                    const float xmean = __fsub_rn(x, rev_f);                                   // :166
                    P[3 * i + a] = __fadd_rn(xmean, __fmul_rn(Gd, nz[a]));                     // :167
                    if (last) pos_mean_out[g3 + 3 * i + a] = xmean;
                }
            }
            __syncthreads();
        }
        for (int i = tid; i < n3; i += NTHREADS) pos_out[g3 + i] = P[i];
    }
}

// ---------------------------------------------------------------------------------------
// edge_2D_emb (eval): e2d tile = W3 . relu(U[src] + V[tgt]) + b3,  SDE_model_2D_to_3D.py:405-407
// uv [N][600]: columns 0..299 = folded first layer applied to h[row], 300..599 to h[col]
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
edge2d_emb_kernel(molsde_plan plan, const float* __restrict__ uv, const float* __restrict__ w3t,
                  const float* __restrict__ b3, float* __restrict__ e2d_tiles) {
    extern __shared__ __align__(16) float smem[];
    float* A = smem;             // [64][TE]
    float* W = smem + 64 * TE;   // [320][32] (rows >= 300 zero)
    __shared__ int s_src[TE], s_tgt[TE];
    const int tid = threadIdx.x, to = tid & 7, te = tid >> 3;
    for (int i = tid; i < 320 * 32; i += NTHREADS) W[i] = (i < 300 * 32) ? w3t[i] : 0.0f;
    __syncthreads();
    for (int tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
        const int ta = plan.tile_tgt_ptr[tile], tb = plan.tile_tgt_ptr[tile + 1];
        const int ea = plan.rowptr[ta], ne = plan.rowptr[tb] - ea;
        __syncthreads();
        for (int i = ta + tid; i < tb; i += NTHREADS)
            for (int e = plan.rowptr[i]; e < plan.rowptr[i + 1]; ++e) { s_tgt[e - ea] = i; s_src[e - ea] = plan.src[e]; }
        __syncthreads();
        float acc[4][4];
        zero_acc(acc);
        for (int k0 = 0; k0 < 320; k0 += 64) {
            const int edge = tid & (TE - 1), kh = (tid >> 7) * 32;
            const bool live = edge < ne;
            const float* up = live ? uv + static_cast<size_t>(s_src[edge]) * 600 : uv;
            const float* vp = live ? uv + static_cast<size_t>(s_tgt[edge]) * 600 + 300 : uv;
#pragma unroll 4
            for (int kk = 0; kk < 32; ++kk) {
                const int k = k0 + kh + kk;
                float v = 0.0f;
                if (live && k < 300) v = fmaxf(__ldg(up + k) + __ldg(vp + k), 0.0f);
                A[(kh + kk) * TE + edge] = v;
            }
            __syncthreads();
            gemm_acc<4, 4, 32, TE>(A, W + k0 * 32, 64, te, to, acc);
            __syncthreads();
        }
        float* out = e2d_tiles + static_cast<size_t>(tile) * (32 * TE);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int col = to * 4 + j;
            const float bj = b3[col];
            float4 o = make_float4(acc[0][j] + bj, acc[1][j] + bj, acc[2][j] + bj, acc[3][j] + bj);
            const int s = te * 4;
            if (s >= ne) o = make_float4(0.f, 0.f, 0.f, 0.f);
            else {
                if (s + 1 >= ne) o.y = 0.f;
                if (s + 2 >= ne) o.z = 0.f;
                if (s + 3 >= ne) o.w = 0.f;
            }
            *reinterpret_cast<float4*>(out + col * TE + s) = o;
        }
    }
}

}  // namespace molsde

using namespace molsde;

static int plan_ok(const molsde_plan* p) {
    return p && p->num_chunks >= 0 && p->num_tiles >= 0 && p->chunk_tile_ptr && p->tile_tgt_ptr && p->rowptr &&
           (p->E == 0 || p->src);
}

extern "C" {

int64_t molsde_sde2d3d_scratch_floats(const molsde_plan* plan, int32_t max_chunk_tiles, int32_t* num_ctas_out) {
    if (!plan || max_chunk_tiles < 0) return MOLSDE_ERR_INVALID;
    int ctas = plan->num_chunks < kNumSMs ? plan->num_chunks : kNumSMs;
    if (ctas < 1) ctas = 1;
    if (num_ctas_out) *num_ctas_out = ctas;
    return static_cast<int64_t>(ctas) * max_chunk_tiles * 32 * TE;
}

int molsde_edge2d_emb_eval(const molsde_plan* plan, const float* uv, const float* w3t, const float* b3,
                           float* e2d_tiles, void* stream) {
    if (!plan_ok(plan) || !uv || !w3t || !b3 || !e2d_tiles) return MOLSDE_ERR_INVALID;
    if (plan->num_tiles == 0) return MOLSDE_OK;
    const size_t smem = sizeof(float) * (64 * TE + 320 * 32);
    cudaError_t err = cudaFuncSetAttribute(edge2d_emb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    int grid = plan->num_tiles < 2 * kNumSMs ? plan->num_tiles : 2 * kNumSMs;
    edge2d_emb_kernel<<<grid, NTHREADS, smem, as_stream(stream)>>>(*plan, uv, w3t, b3, e2d_tiles);
    return check_launch("edge2d_emb");
}

int molsde_sde2d3d_score(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                         const float* e2d_tiles, const float* pos, const float* stdv, float* score, float* scratch,
                         int64_t scratch_floats, int32_t* status_flag, void* stream) {
    if (!plan_ok(plan) || !params || !params->blob || !nattr || !e2d_tiles || !pos || !stdv || !score || !scratch)
        return MOLSDE_ERR_INVALID;
    if (params->blob_floats < MOLSDE_P_TOTAL) return MOLSDE_ERR_INVALID;
    if (plan->num_chunks == 0) return MOLSDE_OK;
    int ctas = plan->num_chunks < kNumSMs ? plan->num_chunks : kNumSMs;
    const int64_t stride = (scratch_floats / ctas) / (32 * TE) * (32 * TE);
    if (stride < 32 * TE) return MOLSDE_ERR_WORKSPACE;
    cudaError_t err = cudaFuncSetAttribute(sde2d3d_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    sde2d3d_score_kernel<<<ctas, NTHREADS, SMEM_BYTES, as_stream(stream)>>>(*plan, params->blob, nattr, e2d_tiles, pos,
                                                                          stdv, score, scratch, stride, status_flag);
    return check_launch("sde2d3d_score");
}

int molsde_sde2d3d_pc_sample(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                             const float* e2d_tiles, const float* pos_init, const float* step_table,
                             const molsde_pc_config* cfg, const float* noise_corr, const float* noise_pred,
                             float* pos_out, float* pos_mean_out, float* scratch, int64_t scratch_floats,
                             int32_t* work_counter, int32_t* status_flag, void* stream) {
    if (!plan_ok(plan) || !params || !params->blob || !nattr || !e2d_tiles || !pos_init || !step_table || !cfg ||
        !pos_out || !pos_mean_out || !scratch || !work_counter)
        return MOLSDE_ERR_INVALID;
    if (params->blob_floats < MOLSDE_P_TOTAL || cfg->steps <= 0) return MOLSDE_ERR_INVALID;
    if ((noise_corr == nullptr) != (noise_pred == nullptr)) return MOLSDE_ERR_INVALID;
    if (plan->num_chunks == 0) return MOLSDE_OK;
    int ctas = plan->num_chunks < kNumSMs ? plan->num_chunks : kNumSMs;
    const int64_t stride = (scratch_floats / ctas) / (32 * TE) * (32 * TE);
    if (stride < 32 * TE) return MOLSDE_ERR_WORKSPACE;
    cudaError_t err = cudaFuncSetAttribute(sde2d3d_pc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    err = cudaMemsetAsync(work_counter, 0, sizeof(int32_t), as_stream(stream));
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    sde2d3d_pc_kernel<<<ctas, NTHREADS, SMEM_BYTES, as_stream(stream)>>>(
        *plan, params->blob, nattr, e2d_tiles, pos_init, step_table, *cfg, noise_corr, noise_pred, pos_out,
        pos_mean_out, scratch, stride, work_counter, status_flag);
    return check_launch("sde2d3d_pc");
}

}  // extern "C"
