// Building blocks of the PRETRAINING step (pretrain_MoleculeSDE.py:105-152): the layer-granular forward ops that keep
// their intermediates, and the backward kernels of every op on the path.  All fp32 FFMA, deterministic (no atomics:
// every reduction has a fixed order), so gradients are bit-reproducible and comparable to the reference's autograd at
// 1e-4.  The sampling path never uses these; it runs the fused kernels of sde2d3d.cu / schnet.cu / dense.cu.
#include "common.cuh"

namespace molsde {

// =====================================================================================================
// GEMM  C[M,N] (+)= op(A)[M,K] . op(B)[K,N]     (64x64x16 tile, 256 threads, 4x4 micro-tile, optional split-K)
//   TA = 0: A stored [M][lda] (k contiguous)      TA = 1: A stored [K][lda] (m contiguous)
//   TB = 0: B stored [K][ldb] (n contiguous)      TB = 1: B stored [N][ldb] (k contiguous)
// Backward of y = x W^T:  dx = dy . W  (TA=0,TB=0),   dW = dy^T . x  (TA=1,TB=0; K = #rows -> split-K).
// =====================================================================================================
constexpr int GM = 64, GN = 64, GK = 16;

template <int TA, int TB>
__global__ void __launch_bounds__(256)
gemm_kernel(int64_t M, int64_t N, int64_t K, const float* __restrict__ A, int64_t lda, const float* __restrict__ B, int64_t ldb,
            float* __restrict__ C, int64_t ldc, int accumulate, int64_t k_per_split, float* __restrict__ ws) {
    __shared__ float As[GK][GM + 4];
    __shared__ float Bs[GK][GN + 4];
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * GM, n0 = static_cast<int64_t>(blockIdx.x) * GN;
    const int64_t kb = static_cast<int64_t>(blockIdx.z) * k_per_split;
    const int64_t ke = min(K, kb + k_per_split);
    float acc[4][4] = {};
    for (int64_t k0 = kb; k0 < ke; k0 += GK) {
        for (int idx = tid; idx < GM * GK; idx += 256) {
            int r, c;
            if (TA) { c = idx / GM; r = idx % GM; } else { r = idx / GK; c = idx % GK; }
            const int64_t gm = m0 + r, gk = k0 + c;
            float v = 0.0f;
            if (gm < M && gk < ke) v = TA ? A[gk * lda + gm] : A[gm * lda + gk];
            As[c][r] = v;
        }
        for (int idx = tid; idx < GN * GK; idx += 256) {
            int r, c;
            if (TB) { r = idx / GK; c = idx % GK; } else { c = idx / GN; r = idx % GN; }
            const int64_t gn = n0 + r, gk = k0 + c;
            float v = 0.0f;
            if (gn < N && gk < ke) v = TB ? B[gn * ldb + gk] : B[gk * ldb + gn];
            Bs[c][r] = v;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < GK; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    float* out = ws ? ws + static_cast<size_t>(blockIdx.z) * M * N : C;
    const int64_t ldo = ws ? N : ldc;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int64_t gm = m0 + ty * 4 + i;
        if (gm >= M) continue;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int64_t gn = n0 + tx * 4 + j;
            if (gn >= N) continue;
            float v = acc[i][j];
            if (!ws && accumulate) v += out[gm * ldo + gn];
            out[gm * ldo + gn] = v;
        }
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ ws, int splits, int64_t M, int64_t N, float* __restrict__ C,
                                     int64_t ldc, int accumulate) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= M * N) return;
    float v = 0.0f;
    for (int s = 0; s < splits; ++s) v += ws[static_cast<size_t>(s) * M * N + idx];
    float* c = C + (idx / N) * ldc + idx % N;
    *c = accumulate ? *c + v : v;
}

// out[n] (+)= sum_m X[m,n]: CTA = 32 columns x 8 row lanes over one row chunk; partials [chunks][N], fixed-order finish
__global__ void __launch_bounds__(256)
colsum_partial_kernel(const float* __restrict__ X, int64_t M, int N, int64_t ldx, int64_t rows_per_chunk, float* __restrict__ part) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n = blockIdx.x * 32 + tx;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
    float acc = 0.0f;
    if (n < N)
        for (int64_t r = r0 + ty; r < r1; r += 8) acc += X[r * ldx + n];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && n < N) {
        float t = 0.0f;
        for (int q = 0; q < 8; ++q) t += red[q][tx];
        part[static_cast<size_t>(blockIdx.y) * N + n] = t;
    }
}
__global__ void colsum_finish_kernel(const float* __restrict__ part, int chunks, int N, float* __restrict__ out, int accumulate) {
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    float t = 0.0f;
    for (int c = 0; c < chunks; ++c) t += part[static_cast<size_t>(c) * N + n];
    out[n] = accumulate ? out[n] + t : t;
}

// ---- activations (act codes of molsde_linear): y = f(x); dx = dy * f'(x) with x the PRE-activation ----
__device__ __forceinline__ float act_f(float v, int act) {
    switch (act) {
        case 1: return fmaxf(v, 0.0f);
        case 2: return silu_f(v);
        case 3: return softplus_f(v) - 0.69314718246459961f;
        case 4: return tanhf(v);
        case 5: return v > 0.0f ? v : expm1f(v);
        default: return v;
    }
}
__device__ __forceinline__ float act_df(float x, int act) {
    switch (act) {
        case 1: return x > 0.0f ? 1.0f : 0.0f;
        case 2: { const float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f + x * (1.0f - s)); }
        case 3: return 1.0f / (1.0f + expf(-x));
        case 4: { const float t = tanhf(x); return 1.0f - t * t; }
        case 5: return x > 0.0f ? 1.0f : expf(x);
        default: return 1.0f;
    }
}
// second derivative of the activations (double backward of SchNet: a force term inside the training loss, finetune_MD17.py:66-77)
__device__ __forceinline__ float act_d2f(float x, int act) {
    switch (act) {
        case 2: { const float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f - s) * (2.0f + x * (1.0f - 2.0f * s)); }
        case 3: { const float s = 1.0f / (1.0f + expf(-x)); return s * (1.0f - s); }
        case 4: { const float t = tanhf(x); return -2.0f * t * (1.0f - t * t); }
        case 5: return x > 0.0f ? 0.0f : expf(x);
        default: return 0.0f;   // identity, relu
    }
}
// out = act''(x) * a * b  (+ out if accumulate)
__global__ void act_bwd2_kernel(const float* __restrict__ x, const float* __restrict__ a, const float* __restrict__ b, int64_t n, int act,
                                int accumulate, float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const float v = act_d2f(x[i], act) * a[i] * b[i];
    out[i] = accumulate ? out[i] + v : v;
}
__global__ void act_fwd_kernel(const float* __restrict__ x, int64_t n, int act, float* __restrict__ y) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) y[i] = act_f(x[i], act);
}
// f'(pre) from the OUTPUT y = f(pre), for the activations where that is a closed form: relu (y > 0), shifted softplus
// (sigmoid(pre) = 1 - exp(-(y + ln 2))), tanh (1 - y^2), elu (y > 0 ? 1 : y + 1).  Lets a linear layer apply the activation in its
// GEMM epilogue and keep ONE tensor instead of pre-activation + output (SiLU has no such form and keeps its pre-activation).
__device__ __forceinline__ float act_df_from_y(float y, int act) {
    switch (act) {
        case 1: return y > 0.0f ? 1.0f : 0.0f;
        case 3: return 1.0f - expf(-(y + 0.69314718246459961f));
        case 4: return 1.0f - y * y;
        case 5: return y > 0.0f ? 1.0f : y + 1.0f;
        default: return 1.0f;
    }
}
__global__ void act_bwd_y_kernel(const float* __restrict__ y, const float* __restrict__ dy, int64_t n, int act, float* __restrict__ dx) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) dx[i] = dy[i] * act_df_from_y(y[i], act);
}
__global__ void act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t n, int act, float* __restrict__ dx) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i < n) dx[i] = dy[i] * act_df(x[i], act);
}

// ---- elementwise: op 0: out = a + alpha*b (b NULL: alpha*a);  op 1: out = a*b (+ c);  op 2: out = a * alpha * b[row];  op 3: out = a + b[col];  op 4: out = a*(1+b[0]) (+c) ----
__global__ void ew_kernel(int op, const float* a, const float* b, const float* c, float alpha, int64_t n, int64_t cols, float* out) {
    // no __restrict__: `out` may alias an input (in-place accumulation)
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float v;
    if (op == 0) v = b ? fmaf(alpha, b[i], a[i]) : alpha * a[i];
    else if (op == 1) v = c ? fmaf(a[i], b[i], c[i]) : a[i] * b[i];
    else if (op == 2) v = a[i] * alpha * b[i / cols];
    else if (op == 3) v = a[i] + b[i % cols];
    else v = c ? fmaf(a[i], 1.0f + b[0], c[i]) : a[i] * (1.0f + b[0]);
    out[i] = v;
}

// out[r,:] = A[ia[r],:] (+ B[ib[r],:])      (x_j / x_i gathers of message passing; int32 indices, NULL index = r)
__global__ void gather_pair_kernel(const float* __restrict__ A, const int32_t* __restrict__ ia, const float* __restrict__ B,
                                   const int32_t* __restrict__ ib, int64_t rows, int cols, float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float v = A[static_cast<int64_t>(ia ? ia[r] : r) * cols + c];
    if (B) v += B[static_cast<int64_t>(ib ? ib[r] : r) * cols + c];
    out[idx] = v;
}

// out[s,:] (+)= scale[s] * sum_{p in [ptr[s],ptr[s+1])} X[perm ? perm[p] : p, :]   in ascending p   (backward of a gather)
__global__ void seg_gather_sum_kernel(const float* __restrict__ X, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                                      int64_t segments, int cols, const float* __restrict__ scale, int accumulate, int row_div,
                                      float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= segments * cols) return;
    const int64_t s = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float acc = 0.0f;
    for (int p = ptr[s]; p < ptr[s + 1]; ++p) acc += X[static_cast<int64_t>((perm ? perm[p] : p) / row_div) * cols + c];
    if (scale) acc *= scale[s];
    out[idx] = accumulate ? out[idx] + acc : acc;
}

// the same for FEW, LONG segments (embedding-table gradients: ~100 rows of the table, thousands of lookups each):
// CTA (32 columns x 32 row lanes) per (segment, column block), fixed-order sum over the 32 lanes
__global__ void __launch_bounds__(1024)
seg_gather_sum_wide_kernel(const float* __restrict__ X, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm, int cols,
                           const float* __restrict__ scale, int accumulate, int row_div, float* __restrict__ out) {
    __shared__ float red[32][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 columns x 32 row lanes (the carbon row of an atom table holds
    const int64_t s = blockIdx.x;                              // most of the batch: long segments need the parallelism)
    const int c = blockIdx.y * 32 + tx;
    float acc = 0.0f;
    if (c < cols)
        for (int p = ptr[s] + ty; p < ptr[s + 1]; p += 32) acc += X[static_cast<int64_t>((perm ? perm[p] : p) / row_div) * cols + c];
    red[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < cols) {
        float t = 0.0f;
        for (int q = 0; q < 32; ++q) t += red[q][tx];
        if (scale) t *= scale[s];
        out[s * cols + c] = accumulate ? out[s * cols + c] + t : t;
    }
}

// stable bucket sort: rowptr[b], perm = element ids ordered by (key, id).  One CTA per bucket, ascending scan.
__global__ void bucket_count_kernel(const int64_t* __restrict__ keys, int64_t n, int32_t* __restrict__ count) {
    __shared__ int red[256];
    const int b = blockIdx.x;
    int c = 0;
    for (int64_t i = threadIdx.x; i < n; i += 256) c += (keys[i] == b);
    red[threadIdx.x] = c;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) count[b] = red[0];
}
__global__ void bucket_fill_kernel(const int64_t* __restrict__ keys, int64_t n, const int32_t* __restrict__ rowptr,
                                   int32_t* __restrict__ perm) {
    // one warp per bucket: ballot-scan in ascending element order
    const int b = blockIdx.x, lane = threadIdx.x;
    int pos = rowptr[b];
    for (int64_t base = 0; base < n; base += 32) {
        const int64_t i = base + lane;
        const bool hit = i < n && keys[i] == b;
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (hit) perm[pos + __popc(m & ((1u << lane) - 1u))] = static_cast<int32_t>(i);
        pos += __popc(m);
    }
}

// out[r,:] = sum_f T[keys[r*F + f], :]     (AtomEncoder / BondEncoder / nn.Embedding: keys carry the per-feature table offset)
__global__ void embed_sum_kernel(const float* __restrict__ T, const int32_t* __restrict__ keys, int64_t rows, int F, int cols,
                                 float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float acc = 0.0f;
    for (int f = 0; f < F; ++f) acc += T[static_cast<int64_t>(keys[r * F + f]) * cols + c];
    out[idx] = acc;
}

// out[s,c] = sum_{p in [ptr[s],ptr[s+1])} A[ia[e],c] * W[e,c],  e = perm ? perm[p] : p
//   CFConv message + aggregation (schnet.py:186-195): s = target, ia = source;   its dx: s = source (CSR by source), ia = target
// `escale` (optional, per edge): the filter is W[e,:] * escale[e] -- the cosine cutoff C(d_e) of CFConv applied on the fly (rounded
// once, exactly as the materialised product was), so the scaled filter stack is never written
__global__ void edge_mul_reduce_kernel(const float* __restrict__ A, const int32_t* __restrict__ ia, const float* __restrict__ W, int64_t ldw,
                                       const float* __restrict__ escale, const int32_t* __restrict__ ptr, const int32_t* __restrict__ perm,
                                       int64_t segments, int cols, float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= segments * cols) return;
    const int64_t s = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float acc = 0.0f;
    for (int p = ptr[s]; p < ptr[s + 1]; ++p) {
        const int64_t e = perm ? perm[p] : p;
        float w = W[e * ldw + c];
        if (escale) w = __fmul_rn(w, escale[e]);
        acc = fmaf(A[static_cast<int64_t>(ia[e]) * cols + c], w, acc);
    }
    out[idx] = acc;
}
// out[e,c] = A[ia[e],c] * B[ib[e],c]      (dW of the CFConv message; `ldo` = row stride of out)
__global__ void edge_mul_gather_kernel(const float* __restrict__ A, const int32_t* __restrict__ ia, const float* __restrict__ B,
                                       const int32_t* __restrict__ ib, const float* __restrict__ escale, int64_t E, int cols,
                                       float* __restrict__ out, int64_t ldo) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= E * cols) return;
    const int64_t e = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float v = __fmul_rn(A[static_cast<int64_t>(ia[e]) * cols + c], B[static_cast<int64_t>(ib[e]) * cols + c]);
    if (escale) v = __fmul_rn(v, escale[e]);     // gradient w.r.t. the UNSCALED filter (two roundings, as product-then-row-scale had)
    out[e * ldo + c] = v;
}

// out[0] (+)= alpha * <a, b>: DOT_CTAS fp64 partial sums into the caller's workspace, then a fixed-order finish
constexpr int DOT_CTAS = 128;
__global__ void __launch_bounds__(256) dot_partial_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n,
                                                          double* __restrict__ part) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t i = blockIdx.x * 256 + threadIdx.x; i < n; i += static_cast<int64_t>(DOT_CTAS) * 256)
        acc += static_cast<double>(a[i]) * b[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void dot_finish_kernel(const double* __restrict__ part, float alpha, int accumulate, float* __restrict__ out) {
    double t = 0.0;
    for (int q = 0; q < DOT_CTAS; ++q) t += part[q];
    out[0] = (accumulate ? out[0] : 0.0f) + alpha * static_cast<float>(t);
}

// ---- GINConv (molecule_gnn_model.py:13-32):  pre[i] = (1+eps) x_i + sum_{e->i} relu(x_src + bond_emb_e),
//      bond_emb_e = sum_f T[ekeys[e*F+f]] (BondEncoder fused; its table is a few rows) ----
__global__ void gin_aggregate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ T, const int32_t* __restrict__ ekeys, int F,
                                         const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const float* __restrict__ eps,
                                         int64_t N, int cols, float* __restrict__ pre) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * cols) return;
    const int64_t i = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float acc = 0.0f;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) {
        float em = 0.0f;
        for (int f = 0; f < F; ++f) em += T[static_cast<int64_t>(ekeys[static_cast<int64_t>(e) * F + f]) * cols + c];
        acc += fmaxf(x[static_cast<int64_t>(src[e]) * cols + c] + em, 0.0f);
    }
    pre[idx] = (1.0f + eps[0]) * x[idx] + acc;
}
// dmsg[e,c] = dpre[tgt_e,c] * [x_src + bond_emb_e > 0]
__global__ void gin_message_bwd_kernel(const float* __restrict__ x, const float* __restrict__ T, const int32_t* __restrict__ ekeys, int F,
                                       const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, const float* __restrict__ dpre,
                                       int64_t E, int cols, float* __restrict__ dmsg) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= E * cols) return;
    const int64_t e = idx / cols;
    const int c = static_cast<int>(idx % cols);
    float em = 0.0f;
    for (int f = 0; f < F; ++f) em += T[static_cast<int64_t>(ekeys[e * F + f]) * cols + c];
    dmsg[idx] = (x[static_cast<int64_t>(src[e]) * cols + c] + em > 0.0f) ? dpre[static_cast<int64_t>(tgt[e]) * cols + c] : 0.0f;
}

// ---- SchNet edge features: GaussianSmearing (schnet.py:205-207) and the cosine cutoff (:186) ----
__global__ void schnet_edge_feat_kernel(const float* __restrict__ pos, const int32_t* __restrict__ src, const int32_t* __restrict__ tgt,
                                        int64_t E, const float* __restrict__ mu, int ng, float coeff, float cutoff,
                                        float* __restrict__ ea, float* __restrict__ C) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= E * ng) return;
    const int64_t e = idx / ng;
    const int k = static_cast<int>(idx % ng);
    const int r = src[e], c = tgt[e];
    const float dx = __fsub_rn(pos[3 * r], pos[3 * c]), dy = __fsub_rn(pos[3 * r + 1], pos[3 * c + 1]),
                dz = __fsub_rn(pos[3 * r + 2], pos[3 * c + 2]);
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float t = d - mu[k];
    ea[idx] = expf(coeff * (t * t));
    if (k == 0) C[e] = 0.5f * (cosf(d * 3.14159274101257324f / cutoff) + 1.0f);
}

// Directional derivative of the edge features along a displacement field v [N,3] of the positions (forward-mode tangent; the
// double backward of SchNet differentiates  <v, dE/dpos> = d/d eps E(pos + eps v)  with respect to the parameters):
//   dd = <pos[src] - pos[tgt], v[src] - v[tgt]> / d;   ea_dot[e,k] = ea[e,k] 2 coeff (d - mu_k) dd;   C_dot[e] = -pi/(2 cutoff) sin(pi d/cutoff) dd
__global__ void schnet_edge_feat_tangent_kernel(const float* __restrict__ pos, const float* __restrict__ v, const int32_t* __restrict__ src,
                                                const int32_t* __restrict__ tgt, int64_t E, const float* __restrict__ mu, int ng,
                                                float coeff, float cutoff, const float* __restrict__ ea, float* __restrict__ ea_dot,
                                                float* __restrict__ C_dot) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= E * ng) return;
    const int64_t e = idx / ng;
    const int k = static_cast<int>(idx % ng);
    const int r = src[e], c = tgt[e];
    const float dx = __fsub_rn(pos[3 * r], pos[3 * c]), dy = __fsub_rn(pos[3 * r + 1], pos[3 * c + 1]),
                dz = __fsub_rn(pos[3 * r + 2], pos[3 * c + 2]);
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    const float pv = dx * (v[3 * r] - v[3 * c]) + dy * (v[3 * r + 1] - v[3 * c + 1]) + dz * (v[3 * r + 2] - v[3 * c + 2]);
    const float dd = d > 0.0f ? pv / d : 0.0f;
    ea_dot[idx] = ea[idx] * (2.0f * coeff * (d - mu[k])) * dd;
    if (k == 0) {
        const float w = 3.14159274101257324f / cutoff;
        C_dot[e] = -0.5f * w * sinf(d * w) * dd;
    }
}

// ---- d/d pos of the SchNet edge features (finetune_MD17.py:66: force = -dE/dpos) ----
// row-wise dot, one warp per row, fixed shuffle order:  out[r] (+)= sum_c a[r,c] * b[r,c]      (d C[e] = <dWf[e,:], f2[e,:]>)
__global__ void __launch_bounds__(256)
rowdot_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t rows, int cols, int accumulate, float* __restrict__ out) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= rows) return;
    const float* ar = a + r * cols;
    const float* br = b + r * cols;
    float acc = 0.0f;
    for (int c = lane; c < cols; c += 32) acc = fmaf(ar[c], br[c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) out[r] = accumulate ? out[r] + acc : acc;
}
// per edge e = (src -> tgt), d = |pos[src] - pos[tgt]|:
//   dd = sum_k dea[e,k] * ea[e,k] * 2 coeff (d - mu_k)  +  dC[e] * (-pi / (2 cutoff)) sin(pi d / cutoff)
//   g[e,:] = dd * (pos[src] - pos[tgt]) / d      (= d loss / d pos[src] of this edge; d pos[tgt] gets -g)
// one warp per edge; the k-sum is a fixed-order shuffle reduction (deterministic).
__global__ void __launch_bounds__(256)
schnet_edge_feat_bwd_kernel(const float* __restrict__ pos, const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t E,
                            const float* __restrict__ mu, int ng, float coeff, float cutoff, const float* __restrict__ ea,
                            const float* __restrict__ dea, const float* __restrict__ dC, float* __restrict__ g) {
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= E) return;
    const int r = src[e], c = tgt[e];
    const float dx = __fsub_rn(pos[3 * r], pos[3 * c]), dy = __fsub_rn(pos[3 * r + 1], pos[3 * c + 1]),
                dz = __fsub_rn(pos[3 * r + 2], pos[3 * c + 2]);
    const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    float acc = 0.0f;
    for (int k = lane; k < ng; k += 32) acc = fmaf(dea[e * ng + k] * ea[e * ng + k], 2.0f * coeff * (d - mu[k]), acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
        const float w = 3.14159274101257324f / cutoff;
        const float dd = acc + (dC ? dC[e] * (-0.5f * w * sinf(d * w)) : 0.0f);
        const float s = d > 0.0f ? dd / d : 0.0f;   // coincident atoms: the norm's subgradient 0, as torch's norm backward
        g[3 * e] = s * dx; g[3 * e + 1] = s * dy; g[3 * e + 2] = s * dz;
    }
}

// ---- EBM_node_dot_prod backward (examples/util.py:52-68): loss = mean softplus(-pp) + mean softplus(pn) ----
//   dX[r] (+)= coef/(N T) * (-sigmoid(-pp_r) Y[r] + sigmoid(pn_r) Y[perm[r]])
//   dY[r] (+)= coef/(N T) * (-sigmoid(-pp_r) X[r] + sigmoid(pn_q) X[q]),  q = invperm[r]
__global__ void ebm_node_dot_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ perm,
                                        const int64_t* __restrict__ invperm, const float* __restrict__ pp, const float* __restrict__ pn,
                                        int64_t N, int D, float scale, int accumulate, float* __restrict__ dX, float* __restrict__ dY) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * D) return;
    const int64_t r = idx / D;
    const int c = static_cast<int>(idx % D);
    const float gp = -1.0f / (1.0f + expf(pp[r]));  // -sigmoid(-pp)
    const float gn = 1.0f / (1.0f + expf(-pn[r]));
    const int64_t q = invperm[r];
    const float gq = 1.0f / (1.0f + expf(-pn[q]));
    const float dx = scale * (gp * Y[idx] + gn * Y[perm[r] * D + c]);
    const float dy = scale * (gp * X[idx] + gq * X[q * D + c]);
    dX[idx] = accumulate ? dX[idx] + dx : dx;
    dY[idx] = accumulate ? dY[idx] + dy : dy;
}

// ---- InfoNCE_dot_prod (examples/util.py:23-32): CrossEntropy(logits, arange) over the rows of logits [B,B] = X Y^T / T ----
// one warp per row: loss_row = logsumexp(row) - row[r], correct_row = (argmax(row) == r);  optionally the gradient
// dlogits = (softmax(row) - onehot(r)) * scale in place of the logits (scale = coef / B).
__global__ void __launch_bounds__(256)
infonce_rows_kernel(float* __restrict__ logits, int64_t B, int64_t ld, float grad_scale, int write_grad, float* __restrict__ loss_row,
                    float* __restrict__ correct_row) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= B) return;
    float* row = logits + r * ld;
    float mx = -INFINITY;
    int64_t arg = 0;
    for (int64_t c = lane; c < B; c += 32) {
        const float v = row[c];
        if (v > mx) { mx = v; arg = c; }   // first maximum per lane (ascending c)
    }
    for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int64_t oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }   // torch.argmax: first index of the maximum
    }
    float s = 0.0f;
    for (int64_t c = lane; c < B; c += 32) s += expf(row[c] - mx);
    s = warp_sum(s);
    const float lse = mx + logf(s);
    if (lane == 0) {
        if (loss_row) loss_row[r] = lse - row[r];
        if (correct_row) correct_row[r] = (arg == r) ? 1.0f : 0.0f;
    }
    if (write_grad) {
        __syncwarp();
        for (int64_t c = lane; c < B; c += 32) {
            const float p = expf(row[c] - lse);
            row[c] = (p - (c == r ? 1.0f : 0.0f)) * grad_scale;
        }
    }
}

// ---- LayerNorm over the last dim (D <= 1024), one warp per row ----
__global__ void layernorm_fwd_kernel(const float* __restrict__ x, int64_t M, int D, const float* __restrict__ g,
                                     const float* __restrict__ b, float eps, float* __restrict__ y, float* __restrict__ mean_out,
                                     float* __restrict__ rstd_out) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const float* xr = x + r * D;
    float s = 0.0f;
    for (int c = lane; c < D; c += 32) s += xr[c];
    const float mu = warp_sum(s) / D;
    float v = 0.0f;
    for (int c = lane; c < D; c += 32) { const float d = xr[c] - mu; v += d * d; }
    const float rstd = rsqrtf(warp_sum(v) / D + eps);
    for (int c = lane; c < D; c += 32) y[r * D + c] = (xr[c] - mu) * rstd * g[c] + b[c];
    if (lane == 0) { mean_out[r] = mu; rstd_out[r] = rstd; }
}
// dx = rstd * (dy*g - mean(dy*g) - xhat * mean(dy*g*xhat));   dyx = dy * xhat (column-summed by the caller for dgamma)
__global__ void layernorm_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t M, int D,
                                     const float* __restrict__ g, const float* __restrict__ mean, const float* __restrict__ rstd,
                                     float* __restrict__ dx, float* __restrict__ dyx) {
    const int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (r >= M) return;
    const float mu = mean[r], rs = rstd[r];
    float s1 = 0.0f, s2 = 0.0f;
    for (int c = lane; c < D; c += 32) {
        const float xh = (x[r * D + c] - mu) * rs, dg = dy[r * D + c] * g[c];
        s1 += dg;
        s2 += dg * xh;
    }
    s1 = warp_sum(s1) / D;
    s2 = warp_sum(s2) / D;
    for (int c = lane; c < D; c += 32) {
        const float xh = (x[r * D + c] - mu) * rs, d = dy[r * D + c];
        dx[r * D + c] = rs * (d * g[c] - s1 - xh * s2);
        dyx[r * D + c] = d * xh;
    }
}

// ---- BatchNorm1d in train mode over rows [M,F] ----
// stats: fp64 sum / sum of squares per column in row chunks (fixed order) -> mean, biased var
__global__ void __launch_bounds__(256)
bn_partial_kernel(const float* __restrict__ X, const float* __restrict__ Y, int64_t M, int F, int64_t rows_per_chunk,
                  const float* __restrict__ mean, const float* __restrict__ rstd, double* __restrict__ part,
                  const float* __restrict__ relu_y = nullptr) {
    // relu_y (mode B only): output of the fused ReLU; dy counts only where relu_y > 0 (the mask of act_bwd, applied in place)
    // mode A (Y == NULL): part0 = sum x, part1 = sum x^2.   mode B: part0 = sum dy, part1 = sum dy * xhat  (X = x, Y = dy)
    __shared__ double red[2][8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int f = blockIdx.x * 32 + tx;
    const int64_t r0 = static_cast<int64_t>(blockIdx.y) * rows_per_chunk, r1 = min(M, r0 + rows_per_chunk);
    double a0 = 0.0, a1 = 0.0;
    if (f < F) {
        const float mu = Y ? mean[f] : 0.0f, rs = Y ? rstd[f] : 0.0f;
        for (int64_t r = r0 + ty; r < r1; r += 8) {
            const float x = X[r * F + f];
            if (Y) {
                float d = Y[r * F + f];
                if (relu_y && !(relu_y[r * F + f] > 0.0f)) d = 0.0f;
                a0 += d; a1 += static_cast<double>(d) * ((x - mu) * rs);
            }
            else { a0 += x; a1 += static_cast<double>(x) * x; }
        }
    }
    red[0][ty][tx] = a0;
    red[1][ty][tx] = a1;
    __syncthreads();
    if (ty == 0 && f < F) {
        double t0 = 0.0, t1 = 0.0;
        for (int q = 0; q < 8; ++q) { t0 += red[0][q][tx]; t1 += red[1][q][tx]; }
        part[(static_cast<size_t>(blockIdx.y) * 2 + 0) * F + f] = t0;
        part[(static_cast<size_t>(blockIdx.y) * 2 + 1) * F + f] = t1;
    }
}
__global__ void bn_stats_finish_kernel(const double* __restrict__ part, int chunks, int F, int64_t M, float eps, float momentum,
                                       float* __restrict__ mean, float* __restrict__ rstd, float* __restrict__ running_mean,
                                       float* __restrict__ running_var) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < chunks; ++c) { s += part[(static_cast<size_t>(c) * 2) * F + f]; q += part[(static_cast<size_t>(c) * 2 + 1) * F + f]; }
    const double mu = s / M;
    double var = q / M - mu * mu;
    if (var < 0.0) var = 0.0;
    mean[f] = static_cast<float>(mu);
    rstd[f] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
    if (running_mean) {
        const float unbiased = static_cast<float>(var * (static_cast<double>(M) / static_cast<double>(M > 1 ? M - 1 : 1)));
        running_mean[f] = (1.0f - momentum) * running_mean[f] + momentum * static_cast<float>(mu);
        running_var[f] = (1.0f - momentum) * running_var[f] + momentum * unbiased;
    }
}
__global__ void bn_bwd_finish_kernel(const double* __restrict__ part, int chunks, int F, float* __restrict__ dbeta,
                                     float* __restrict__ dgamma, int accumulate, float* __restrict__ gbeta = nullptr,
                                     float* __restrict__ ggamma = nullptr) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < chunks; ++c) { s += part[(static_cast<size_t>(c) * 2) * F + f]; q += part[(static_cast<size_t>(c) * 2 + 1) * F + f]; }
    dbeta[f] = (accumulate ? dbeta[f] : 0.0f) + static_cast<float>(s);
    dgamma[f] = (accumulate ? dgamma[f] : 0.0f) + static_cast<float>(q);
    if (gbeta) gbeta[f] += static_cast<float>(s);     // parameter-gradient buffers: accumulate (same add as a separate a += b pass)
    if (ggamma) ggamma[f] += static_cast<float>(q);
}
// y = act((x - mean) * rstd * gamma + beta)   (act: 0 none, 1 relu)
__global__ void bn_apply_kernel(const float* __restrict__ x, int64_t M, int F, const float* __restrict__ mean,
                                const float* __restrict__ rstd, const float* __restrict__ gamma, const float* __restrict__ beta,
                                int act, float* __restrict__ y) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= M * F) return;
    const int f = static_cast<int>(i % F);
    float v = (x[i] - mean[f]) * rstd[f] * gamma[f] + beta[f];
    if (act == 1) v = fmaxf(v, 0.0f);
    y[i] = v;
}
__global__ void rsqrt_eps_kernel(const float* __restrict__ v, int n, float eps, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = 1.0f / sqrtf(v[i] + eps);
}
// dx = gamma * rstd * (dy' - dbeta/M - xhat * dgamma/M), dy' = dy * [y > 0] for act == 1 (caller passes dy' already masked
// in the reductions: the mask is applied here AND in bn_partial via the masked dy buffer the host builds with act_bwd).
__global__ void bn_bwd_dx_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t M, int F,
                                 const float* __restrict__ mean, const float* __restrict__ rstd, const float* __restrict__ gamma,
                                 const float* __restrict__ dbeta, const float* __restrict__ dgamma, float* __restrict__ dx,
                                 const float* __restrict__ relu_y = nullptr) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= M * F) return;
    const int f = static_cast<int>(i % F);
    const float xh = (x[i] - mean[f]) * rstd[f];
    const float inv_m = 1.0f / static_cast<float>(M);
    float d = dy[i];
    if (relu_y && !(relu_y[i] > 0.0f)) d = 0.0f;
    dx[i] = gamma[f] * rstd[f] * (d - dbeta[f] * inv_m - xh * dgamma[f] * inv_m);
}

// ---- flat Adam (torch.optim.Adam, amsgrad=False, weight_decay as L2):  one launch over the whole parameter buffer ----
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v, int64_t n,
                            float lr, float beta1, float beta2, float eps, float weight_decay, float bc1, float bc2_sqrt,
                            float grad_scale) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float gi = g[i] * grad_scale;
    if (weight_decay != 0.0f) gi = fmaf(weight_decay, p[i], gi);
    const float mi = m[i] + (1.0f - beta1) * (gi - m[i]);        // torch: exp_avg.lerp_(grad, 1 - beta1)
    const float vi = beta2 * v[i] + (1.0f - beta2) * gi * gi;    // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / bc2_sqrt + eps;
    p[i] = p[i] - (lr / bc1) * (mi / denom);
}

}  // namespace molsde

using namespace molsde;

static inline unsigned blocks_for(int64_t n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

static int gemm_splits(int64_t M, int64_t N, int64_t K) {
    // weight gradients are tall-skinny reductions (K = #rows >> M, N): split K until ~4 CTAs per SM are in flight, each
    // with at least 2 k-steps; the partial buffer (splits x M x N) stays small because M x N is a weight matrix
    const int64_t tiles = ((M + GM - 1) / GM) * ((N + GN - 1) / GN);
    int64_t s = (4 * kNumSMs + tiles - 1) / tiles;
    const int64_t maxs = (K + 2 * GK - 1) / (2 * GK);
    if (s > maxs) s = maxs;
    if (s > 1024) s = 1024;
    return s < 1 ? 1 : static_cast<int>(s);
}

extern "C" {

int64_t molsde_gemm_ws_floats(int64_t M, int64_t N, int64_t K) {
    const int s = gemm_splits(M, N, K);
    return s > 1 ? s * M * N : 0;
}

int molsde_gemm(int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                int64_t ldb, float* C, int64_t ldc, int32_t accumulate, float* ws, int64_t ws_floats, void* stream) {
    if (!A || !B || !C || M < 0 || N < 0 || K < 0) return MOLSDE_ERR_INVALID;
    if (M == 0 || N == 0) return MOLSDE_OK;
    int splits = gemm_splits(M, N, K);
    if (splits > 1 && (!ws || ws_floats < static_cast<int64_t>(splits) * M * N)) splits = 1;
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + GK - 1) / GK * GK;
    if (kps < GK) kps = GK;
    dim3 grid(static_cast<unsigned>((N + GN - 1) / GN), static_cast<unsigned>((M + GM - 1) / GM), splits);
    float* w = splits > 1 ? ws : nullptr;
    cudaStream_t s = as_stream(stream);
    if (transA && transB) gemm_kernel<1, 1><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kps, w);
    else if (transA) gemm_kernel<1, 0><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kps, w);
    else if (transB) gemm_kernel<0, 1><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kps, w);
    else gemm_kernel<0, 0><<<grid, 256, 0, s>>>(M, N, K, A, lda, B, ldb, C, ldc, accumulate, kps, w);
    int st = check_launch("gemm");
    if (st != MOLSDE_OK || splits == 1) return st;
    splitk_reduce_kernel<<<blocks_for(M * N), 256, 0, s>>>(ws, splits, M, N, C, ldc, accumulate);
    return check_launch("gemm.splitk_reduce");
}

int64_t molsde_colsum_ws_floats(int64_t M, int32_t N) {
    int64_t chunks = (M + 511) / 512;
    if (chunks > 128) chunks = 128;
    if (chunks < 1) chunks = 1;
    return chunks * N;
}

int molsde_colsum(const float* X, int64_t M, int32_t N, int64_t ldx, float* out, int32_t accumulate, float* ws, int64_t ws_floats,
                  void* stream) {
    if (!X || !out || !ws || M < 0 || N <= 0) return MOLSDE_ERR_INVALID;
    int64_t chunks = (M + 511) / 512;
    if (chunks > 128) chunks = 128;
    if (chunks < 1) chunks = 1;
    if (ws_floats < chunks * N) return MOLSDE_ERR_INVALID;
    const int64_t rpc = (M + chunks - 1) / chunks;
    dim3 grid((N + 31) / 32, static_cast<unsigned>(chunks));
    colsum_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(X, M, N, ldx, rpc > 0 ? rpc : 1, ws);
    int st = check_launch("colsum");
    if (st != MOLSDE_OK) return st;
    colsum_finish_kernel<<<(N + 127) / 128, 128, 0, as_stream(stream)>>>(ws, static_cast<int>(chunks), N, out, accumulate);
    return check_launch("colsum.finish");
}

int molsde_act_fwd(const float* x, int64_t n, int32_t act, float* y, void* stream) {
    if (!x || !y || n < 0) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    act_fwd_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, n, act, y);
    return check_launch("act_fwd");
}
int molsde_act_bwd(const float* x, const float* dy, int64_t n, int32_t act, float* dx, void* stream) {
    if (!x || !dy || !dx || n < 0) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    act_bwd_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, dy, n, act, dx);
    return check_launch("act_bwd");
}
int molsde_act_bwd_y(const float* y, const float* dy, int64_t n, int32_t act, float* dx, void* stream) {
    if (!y || !dy || !dx || n < 0) return MOLSDE_ERR_INVALID;
    if (act != 1 && act != 3 && act != 4 && act != 5) return MOLSDE_ERR_UNSUPPORTED;
    if (n == 0) return MOLSDE_OK;
    act_bwd_y_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(y, dy, n, act, dx);
    return check_launch("act_bwd_y");
}
int molsde_act_bwd2(const float* x, const float* a, const float* b, int64_t n, int32_t act, int32_t accumulate, float* out, void* stream) {
    if (!x || !a || !b || !out || n < 0) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    act_bwd2_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, a, b, n, act, accumulate, out);
    return check_launch("act_bwd2");
}
int molsde_schnet_edge_feat_tangent(const float* pos, const float* v, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu,
                                    int32_t ng, float coeff, float cutoff, const float* ea, float* ea_dot, float* C_dot, void* stream) {
    if (!pos || !v || !src || !tgt || !mu || !ea || !ea_dot || !C_dot || E < 0 || ng < 1) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    schnet_edge_feat_tangent_kernel<<<blocks_for(E * ng), 256, 0, as_stream(stream)>>>(pos, v, src, tgt, E, mu, ng, coeff, cutoff, ea, ea_dot,
                                                                                    C_dot);
    return check_launch("schnet_edge_feat_tangent");
}
int molsde_ew(int32_t op, const float* a, const float* b, const float* c, float alpha, int64_t n, int64_t cols, float* out,
              void* stream) {
    if (!a || !out || n < 0 || op < 0 || op > 4 || (op >= 1 && !b) || ((op == 2 || op == 3) && cols <= 0)) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    ew_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(op, a, b, c, alpha, n, cols, out);
    return check_launch("ew");
}
int molsde_gather_pair(const float* A, const int32_t* ia, const float* B, const int32_t* ib, int64_t rows, int32_t cols, float* out,
                       void* stream) {
    if (!A || !out || rows < 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    gather_pair_kernel<<<blocks_for(rows * cols), 256, 0, as_stream(stream)>>>(A, ia, B, ib, rows, cols, out);
    return check_launch("gather_pair");
}
int molsde_seg_gather_sum(const float* X, const int32_t* ptr, const int32_t* perm, int64_t segments, int32_t cols, const float* scale,
                          int32_t accumulate, int32_t row_div, float* out, void* stream) {
    if (!X || !ptr || !out || segments < 0 || cols <= 0 || row_div <= 0) return MOLSDE_ERR_INVALID;
    if (segments == 0) return MOLSDE_OK;
    if (segments <= 1024) {  // few segments: one CTA per (segment, 32 columns) instead of one thread per output
        seg_gather_sum_wide_kernel<<<dim3(static_cast<unsigned>(segments), (cols + 31) / 32), 1024, 0, as_stream(stream)>>>(
            X, ptr, perm, cols, scale, accumulate, row_div, out);
        return check_launch("seg_gather_sum_wide");
    }
    seg_gather_sum_kernel<<<blocks_for(segments * cols), 256, 0, as_stream(stream)>>>(X, ptr, perm, segments, cols, scale, accumulate,
                                                                                   row_div, out);
    return check_launch("seg_gather_sum");
}
int molsde_embed_sum(const float* T, const int32_t* keys, int64_t rows, int32_t F, int32_t cols, float* out, void* stream) {
    if (!T || !keys || !out || rows < 0 || F <= 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    embed_sum_kernel<<<blocks_for(rows * cols), 256, 0, as_stream(stream)>>>(T, keys, rows, F, cols, out);
    return check_launch("embed_sum");
}
int molsde_edge_mul_reduce_ld(const float* A, const int32_t* ia, const float* W, int64_t ldw, const float* escale, const int32_t* ptr,
                              const int32_t* perm, int64_t segments, int32_t cols, float* out, void* stream) {
    if (!A || !ia || !W || !ptr || !out || segments < 0 || cols <= 0 || ldw < cols) return MOLSDE_ERR_INVALID;
    if (segments == 0) return MOLSDE_OK;
    edge_mul_reduce_kernel<<<blocks_for(segments * cols), 256, 0, as_stream(stream)>>>(A, ia, W, ldw, escale, ptr, perm, segments, cols, out);
    return check_launch("edge_mul_reduce");
}
int molsde_edge_mul_reduce(const float* A, const int32_t* ia, const float* W, const int32_t* ptr, const int32_t* perm, int64_t segments,
                           int32_t cols, float* out, void* stream) {
    return molsde_edge_mul_reduce_ld(A, ia, W, cols, nullptr, ptr, perm, segments, cols, out, stream);
}
int molsde_edge_mul_gather_ld(const float* A, const int32_t* ia, const float* B, const int32_t* ib, const float* escale, int64_t E,
                              int32_t cols, float* out, int64_t ldo, void* stream) {
    if (!A || !ia || !B || !ib || !out || E < 0 || cols <= 0 || ldo < cols) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    edge_mul_gather_kernel<<<blocks_for(E * cols), 256, 0, as_stream(stream)>>>(A, ia, B, ib, escale, E, cols, out, ldo);
    return check_launch("edge_mul_gather");
}
int molsde_edge_mul_gather(const float* A, const int32_t* ia, const float* B, const int32_t* ib, int64_t E, int32_t cols, float* out,
                           void* stream) {
    return molsde_edge_mul_gather_ld(A, ia, B, ib, nullptr, E, cols, out, cols, stream);
}
int molsde_dot(const float* a, const float* b, int64_t n, float alpha, int32_t accumulate, float* out, double* ws, void* stream) {
    if (!a || !b || !out || !ws || n < 0) return MOLSDE_ERR_INVALID;  // ws: >= 128 doubles
    dot_partial_kernel<<<DOT_CTAS, 256, 0, as_stream(stream)>>>(a, b, n, ws);
    dot_finish_kernel<<<1, 1, 0, as_stream(stream)>>>(ws, alpha, accumulate, out);
    return check_launch("dot");
}
int molsde_gin_aggregate_fwd(const float* x, const float* T, const int32_t* ekeys, int32_t F, const int32_t* rowptr, const int32_t* src,
                             const float* eps, int64_t N, int32_t cols, float* pre, void* stream) {
    if (!x || !T || !ekeys || !rowptr || !src || !eps || !pre || N < 0 || F <= 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    gin_aggregate_fwd_kernel<<<blocks_for(N * cols), 256, 0, as_stream(stream)>>>(x, T, ekeys, F, rowptr, src, eps, N, cols, pre);
    return check_launch("gin_aggregate_fwd");
}
int molsde_gin_message_bwd(const float* x, const float* T, const int32_t* ekeys, int32_t F, const int32_t* src, const int32_t* tgt,
                           const float* dpre, int64_t E, int32_t cols, float* dmsg, void* stream) {
    if (!x || !T || !ekeys || !src || !tgt || !dpre || !dmsg || E < 0 || F <= 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    gin_message_bwd_kernel<<<blocks_for(E * cols), 256, 0, as_stream(stream)>>>(x, T, ekeys, F, src, tgt, dpre, E, cols, dmsg);
    return check_launch("gin_message_bwd");
}
int molsde_schnet_edge_feat(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu, int32_t ng,
                            float coeff, float cutoff, float* ea, float* C, void* stream) {
    if (!pos || !src || !tgt || !mu || !ea || !C || E < 0 || ng <= 0) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    schnet_edge_feat_kernel<<<blocks_for(E * ng), 256, 0, as_stream(stream)>>>(pos, src, tgt, E, mu, ng, coeff, cutoff, ea, C);
    return check_launch("schnet_edge_feat");
}
int molsde_rowdot(const float* a, const float* b, int64_t rows, int32_t cols, int32_t accumulate, float* out, void* stream) {
    if (!a || !b || !out || rows < 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    rowdot_kernel<<<blocks_for(rows, 8), 256, 0, as_stream(stream)>>>(a, b, rows, cols, accumulate, out);
    return check_launch("rowdot");
}
int molsde_schnet_edge_feat_bwd(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu, int32_t ng,
                                float coeff, float cutoff, const float* ea, const float* dea, const float* dC, float* g, void* stream) {
    if (!pos || !src || !tgt || !mu || !ea || !dea || !g || E < 0 || ng <= 0) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    schnet_edge_feat_bwd_kernel<<<blocks_for(E, 8), 256, 0, as_stream(stream)>>>(pos, src, tgt, E, mu, ng, coeff, cutoff, ea, dea, dC, g);
    return check_launch("schnet_edge_feat_bwd");
}
int molsde_ebm_node_dot_bwd(const float* X, const float* Y, const int64_t* perm, const int64_t* invperm, const float* pred_pos,
                            const float* pred_neg, int64_t N, int32_t D, float T, float coef, int32_t accumulate, float* dX, float* dY,
                            void* stream) {
    if (!X || !Y || !perm || !invperm || !pred_pos || !pred_neg || !dX || !dY || N <= 0 || D <= 0) return MOLSDE_ERR_INVALID;
    ebm_node_dot_bwd_kernel<<<blocks_for(N * D), 256, 0, as_stream(stream)>>>(X, Y, perm, invperm, pred_pos, pred_neg, N, D,
                                                                           coef / (static_cast<float>(N) * T), accumulate, dX, dY);
    return check_launch("ebm_node_dot_bwd");
}
int molsde_infonce_rows(float* logits, int64_t B, int64_t ld, float grad_scale, int32_t write_grad, float* loss_row, float* correct_row,
                        void* stream) {
    if (!logits || B <= 0 || ld < B) return MOLSDE_ERR_INVALID;
    infonce_rows_kernel<<<blocks_for(B, 8), 256, 0, as_stream(stream)>>>(logits, B, ld, grad_scale, write_grad, loss_row, correct_row);
    return check_launch("infonce_rows");
}
int molsde_bucket_count(const int64_t* keys, int64_t n, int32_t buckets, int32_t* count, void* stream) {
    if (!keys || !count || n < 0 || buckets <= 0) return MOLSDE_ERR_INVALID;
    bucket_count_kernel<<<buckets, 256, 0, as_stream(stream)>>>(keys, n, count);
    return check_launch("bucket_count");
}
int molsde_bucket_fill(const int64_t* keys, int64_t n, int32_t buckets, const int32_t* rowptr, int32_t* perm, void* stream) {
    if (!keys || !rowptr || !perm || n < 0 || buckets <= 0) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    bucket_fill_kernel<<<buckets, 32, 0, as_stream(stream)>>>(keys, n, rowptr, perm);
    return check_launch("bucket_fill");
}
int molsde_layernorm_fwd(const float* x, int64_t M, int32_t D, const float* g, const float* b, float eps, float* y, float* mean,
                         float* rstd, void* stream) {
    if (!x || !g || !b || !y || !mean || !rstd || M < 0 || D <= 0) return MOLSDE_ERR_INVALID;
    if (M == 0) return MOLSDE_OK;
    layernorm_fwd_kernel<<<blocks_for(M, 8), 256, 0, as_stream(stream)>>>(x, M, D, g, b, eps, y, mean, rstd);
    return check_launch("layernorm_fwd");
}
int molsde_layernorm_bwd(const float* x, const float* dy, int64_t M, int32_t D, const float* g, const float* mean, const float* rstd,
                         float* dx, float* dyx, void* stream) {
    if (!x || !dy || !g || !mean || !rstd || !dx || !dyx || M < 0 || D <= 0) return MOLSDE_ERR_INVALID;
    if (M == 0) return MOLSDE_OK;
    layernorm_bwd_kernel<<<blocks_for(M, 8), 256, 0, as_stream(stream)>>>(x, dy, M, D, g, mean, rstd, dx, dyx);
    return check_launch("layernorm_bwd");
}

static int bn_chunks(int64_t M) {
    int64_t c = (M + 255) / 256;
    if (c > 64) c = 64;
    return c < 1 ? 1 : static_cast<int>(c);
}
int64_t molsde_bn_ws_doubles(int64_t M, int32_t F) { return static_cast<int64_t>(bn_chunks(M)) * 2 * F; }

int molsde_bn_train_fwd(const float* x, int64_t M, int32_t F, const float* gamma, const float* beta, float eps, float momentum,
                        float* running_mean, float* running_var, int32_t act, float* y, float* mean, float* rstd, double* ws,
                        void* stream) {
    if (!x || !gamma || !beta || !y || !mean || !rstd || !ws || M <= 0 || F <= 0) return MOLSDE_ERR_INVALID;
    const int chunks = bn_chunks(M);
    const int64_t rpc = (M + chunks - 1) / chunks;
    dim3 grid((F + 31) / 32, chunks);
    bn_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, nullptr, M, F, rpc, nullptr, nullptr, ws);
    int st = check_launch("bn_partial");
    if (st != MOLSDE_OK) return st;
    bn_stats_finish_kernel<<<(F + 127) / 128, 128, 0, as_stream(stream)>>>(ws, chunks, F, M, eps, momentum, mean, rstd, running_mean,
                                                                         running_var);
    st = check_launch("bn_stats_finish");
    if (st != MOLSDE_OK) return st;
    bn_apply_kernel<<<blocks_for(M * F), 256, 0, as_stream(stream)>>>(x, M, F, mean, rstd, gamma, beta, act, y);
    return check_launch("bn_apply");
}
/* eval-mode BatchNorm1d (+ReLU): y = act((x - running_mean) / sqrt(running_var + eps) * gamma + beta) */
int molsde_bn_eval(const float* x, int64_t M, int32_t F, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float eps, int32_t act, float* y, float* rstd_tmp, void* stream) {
    if (!x || !gamma || !beta || !running_mean || !running_var || !y || !rstd_tmp || M < 0 || F <= 0) return MOLSDE_ERR_INVALID;
    if (M == 0) return MOLSDE_OK;
    rsqrt_eps_kernel<<<blocks_for(F), 256, 0, as_stream(stream)>>>(running_var, F, eps, rstd_tmp);
    bn_apply_kernel<<<blocks_for(M * F), 256, 0, as_stream(stream)>>>(x, M, F, running_mean, rstd_tmp, gamma, beta, act, y);
    return check_launch("bn_eval");
}
/* relu_y: output of the fused ReLU (NULL: dy is used as is); dgamma / dbeta are fresh outputs (this layer's totals, needed by
 * dx); grad_gamma / grad_beta (optional): parameter-gradient buffers the totals are ADDED to in the same launch. */
int molsde_bn_train_bwd_fused(const float* x, const float* dy, const float* relu_y, int64_t M, int32_t F, const float* gamma,
                              const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta, float* grad_gamma,
                              float* grad_beta, double* ws, void* stream) {
    if (!x || !dy || !gamma || !mean || !rstd || !dx || !dgamma || !dbeta || !ws || M <= 0 || F <= 0) return MOLSDE_ERR_INVALID;
    const int chunks = bn_chunks(M);
    const int64_t rpc = (M + chunks - 1) / chunks;
    dim3 grid((F + 31) / 32, chunks);
    bn_partial_kernel<<<grid, 256, 0, as_stream(stream)>>>(x, dy, M, F, rpc, mean, rstd, ws, relu_y);
    int st = check_launch("bn_bwd_partial");
    if (st != MOLSDE_OK) return st;
    bn_bwd_finish_kernel<<<(F + 127) / 128, 128, 0, as_stream(stream)>>>(ws, chunks, F, dbeta, dgamma, 0, grad_beta, grad_gamma);
    st = check_launch("bn_bwd_finish");
    if (st != MOLSDE_OK) return st;
    bn_bwd_dx_kernel<<<blocks_for(M * F), 256, 0, as_stream(stream)>>>(x, dy, M, F, mean, rstd, gamma, dbeta, dgamma, dx, relu_y);
    return check_launch("bn_bwd_dx");
}
/* dy must already carry the activation mask (relu'); dgamma / dbeta are fresh outputs */
int molsde_bn_train_bwd(const float* x, const float* dy, int64_t M, int32_t F, const float* gamma, const float* mean,
                        const float* rstd, float* dx, float* dgamma, float* dbeta, double* ws, void* stream) {
    return molsde_bn_train_bwd_fused(x, dy, nullptr, M, F, gamma, mean, rstd, dx, dgamma, dbeta, nullptr, nullptr, ws, stream);
}

int molsde_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int32_t step, float grad_scale, void* stream) {
    if (!p || !g || !m || !v || n < 0 || step < 1) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    const float bc1 = 1.0f - powf(beta1, static_cast<float>(step));
    const float bc2_sqrt = sqrtf(1.0f - powf(beta2, static_cast<float>(step)));
    adam_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(p, g, m, v, n, lr, beta1, beta2, eps, weight_decay, bc1, bc2_sqrt,
                                                            grad_scale);
    return check_launch("adam_step");
}

}  // extern "C"
