// Dense 3D->2D score networks (SDEModel3Dto2D_node_adj_dense): building-block kernels.
//
// Reference: Geom3D/models/MoleculeSDE/SDE_model_3D_to_2D_node_adj_dense.py:101-179,523-562,
// invariant_scorenetwork_dense.py:28-131, layers/edge_network_dense.py:33-128,
// layers/node_network_dense.py:25-88.  Tensors are dense per graph: x [B,Nm,F], adj [B,Nm,Nm],
// channels-first adjacency stacks [B,C,Nm,Nm] and channels-last pair features [B,Nm,Nm,C] (the layout
// the per-pair MLPs consume as rows of a plain linear layer).  Round-1 structure: one kernel per
// reference op group, all channels of a layer batched in one launch; the MLPs go through
// molsde_linear / molsde_grouped_linear.  (Fusion + tcgen05 for the 364->728->728->119 node MLP is
// the planned next step, DESIGN.md.)
#include "common.cuh"

namespace molsde {

constexpr int DN_MAX = 64;  // max padded atoms per graph

// ---------------------------------------------------------------------------------------
// to_dense_batch / to_dense_adj / node_flags   (SDE_model_3D_to_2D_node_adj_dense.py:124-134,523-529)
// ---------------------------------------------------------------------------------------
__global__ void to_dense_batch_kernel(const float* __restrict__ x, const int32_t* __restrict__ node_ptr, int B, int Nm,
                                      int F, float* __restrict__ out, int64_t ldo) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<int64_t>(B) * Nm * F) return;
    const int f = static_cast<int>(i % F);
    const int64_t r = i / F;
    const int b = static_cast<int>(r / Nm), a = static_cast<int>(r % Nm);
    const int n0 = node_ptr[b], n = node_ptr[b + 1] - n0;
    out[r * ldo + f] = a < n ? x[static_cast<int64_t>(n0 + a) * F + f] : 0.0f;
}

__global__ void to_dense_adj_kernel(const int64_t* __restrict__ edge_index, int64_t E, const float* __restrict__ val,
                                    const int64_t* __restrict__ val_i64, float val_add, const int32_t* __restrict__ node_ptr,
                                    int B, int Nm, float* __restrict__ adj) {
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (e >= E) return;
    const int r = static_cast<int>(edge_index[e]), c = static_cast<int>(edge_index[E + e]);
    int lo = 0, hi = B;  // graph of the source node
    while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (node_ptr[mid] <= r) lo = mid; else hi = mid;
    }
    const int n0 = node_ptr[lo];
    const float v = (val ? val[e] : static_cast<float>(val_i64[e])) + val_add;
    atomicAdd(&adj[(static_cast<int64_t>(lo) * Nm + (r - n0)) * Nm + (c - n0)], v);  // scatter-ADD (to_dense_adj)
}

__global__ void node_flags_kernel(const float* __restrict__ adj, int B, int Nm, float eps, float* __restrict__ flags) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * Nm) return;
    const float* row = adj + static_cast<int64_t>(i) * Nm;
    float s = 0.0f;
    for (int j = 0; j < Nm; ++j) s += fabsf(row[j]);
    flags[i] = s > eps ? 1.0f : 0.0f;
}

// ---------------------------------------------------------------------------------------
// grouped linear:  Y[:, g*No:(g+1)*No] = act(X[:, g*Ki:(g+1)*Ki] . W[g]^T + b[g])     (per-channel 2nd MLP layers)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dense_act(float v, int act) {
    switch (act) {
        case 1: return fmaxf(v, 0.0f);
        case 2: return v / (1.0f + expf(-v));
        case 4: return tanhf(v);
        case 5: return v > 0.0f ? v : expm1f(v);  // F.elu, alpha = 1
        default: return v;
    }
}

__global__ void grouped_linear_kernel(const float* __restrict__ X, int64_t rows, int64_t ldx, const float* __restrict__ W,
                                      const float* __restrict__ b, int G, int Ki, int No, float* __restrict__ Y,
                                      int64_t ldy, int act) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int per_row = G * No;
    if (i >= rows * per_row) return;
    const int64_t r = i / per_row;
    const int q = static_cast<int>(i % per_row), g = q / No, o = q % No;
    const float* xr = X + r * ldx + g * Ki;
    const float* wr = W + (static_cast<int64_t>(g) * No + o) * Ki;
    float acc = 0.0f;
    for (int k = 0; k < Ki; ++k) acc = fmaf(xr[k], wr[k], acc);
    Y[r * ldy + q] = dense_act(acc + (b ? b[g * No + o] : 0.0f), act);
}

// ---------------------------------------------------------------------------------------
// pow_tensor (invariant_scorenetwork_dense.py:28-37) for c_init = 2:  adjc[b,0] = adj, adjc[b,1] = adj @ adj;
// also writes both channels into the channels-last "all channels" buffer (columns all_off, all_off+1).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
pow2_kernel(const float* __restrict__ adj, int Nm, float* __restrict__ adjc, float* __restrict__ allc, int ld_all,
            int all_off) {
    __shared__ float A[DN_MAX][DN_MAX + 1];
    const int b = blockIdx.x;
    const float* a = adj + static_cast<int64_t>(b) * Nm * Nm;
    for (int i = threadIdx.x; i < Nm * Nm; i += blockDim.x) A[i / Nm][i % Nm] = a[i];
    __syncthreads();
    float* o0 = adjc + static_cast<int64_t>(b) * 2 * Nm * Nm;
    float* o1 = o0 + Nm * Nm;
    for (int i = threadIdx.x; i < Nm * Nm; i += blockDim.x) {
        const int r = i / Nm, c = i % Nm;
        float s = 0.0f;
        for (int k = 0; k < Nm; ++k) s = fmaf(A[r][k], A[k][c], s);
        o0[i] = A[r][c];
        o1[i] = s;
        if (allc) {   // (NULL on the fused inference path: the head reads the channel-major stacks directly)
            float* ac = allc + (static_cast<int64_t>(b) * Nm * Nm + i) * ld_all + all_off;
            ac[0] = A[r][c];
            ac[1] = s;
        }
    }
}

// ---------------------------------------------------------------------------------------
// NodeNetwork_dense / DenseGCNConv clone (node_network_dense.py:46-85), all channels in one launch:
//   A~ = adjc[b,c] with unit diagonal;  dis = clamp(rowsum(A~), 1)^-1/2;
//   out[b,i, c*Fo + f] = act( sum_j ((dis_i A~_ij) dis_j) xw[b,j, c*Fo + f] + bias[c*Fo + f] )
// grid (B, C); adjacency stack strides are given so that a plain [B,Nm,Nm] adj works with C = 1.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_gcn_kernel(const float* __restrict__ adjc, int64_t adj_stride_b, int64_t adj_stride_c, int Nm,
                 const float* __restrict__ xw, int64_t ldxw, const float* __restrict__ bias, int Fo,
                 float* __restrict__ out, int64_t ldo, int out_off, int act) {
    __shared__ float A[DN_MAX][DN_MAX + 1];
    __shared__ float dis[DN_MAX];
    const int b = blockIdx.x, c = blockIdx.y;
    const float* a = adjc + b * adj_stride_b + c * adj_stride_c;
    for (int i = threadIdx.x; i < Nm * Nm; i += blockDim.x) {
        const int r = i / Nm, cc = i % Nm;
        A[r][cc] = (r == cc) ? 1.0f : a[i];
    }
    __syncthreads();
    if (threadIdx.x < Nm) {
        float s = 0.0f;
        for (int j = 0; j < Nm; ++j) s += A[threadIdx.x][j];
        dis[threadIdx.x] = 1.0f / sqrtf(fmaxf(s, 1.0f));  // clamp(min=1).pow(-0.5)
    }
    __syncthreads();
    const float* xb = xw + static_cast<int64_t>(b) * Nm * ldxw + c * Fo;
    float* ob = out + static_cast<int64_t>(b) * Nm * ldo + out_off + c * Fo;
    // the graph's xw tile [Nm][Fo] is read Nm times by every output row: staged once in shared memory (Fo <= 32), pre-scaled by dis_j
    __shared__ float Xs[DN_MAX][33];
    if (Fo <= 32) {
        for (int p = threadIdx.x; p < Nm * Fo; p += blockDim.x) {
            const int j = p / Fo, f = p % Fo;
            Xs[j][f] = xb[j * ldxw + f];
        }
        __syncthreads();
        for (int p = threadIdx.x; p < Nm * Fo; p += blockDim.x) {
            const int i = p / Fo, f = p % Fo;
            float s = 0.0f;
            for (int j = 0; j < Nm; ++j) s = fmaf((dis[i] * A[i][j]) * dis[j], Xs[j][f], s);
            ob[i * ldo + f] = dense_act(s + bias[c * Fo + f], act);
        }
        return;
    }
    for (int p = threadIdx.x; p < Nm * Fo; p += blockDim.x) {
        const int i = p / Fo, f = p % Fo;
        float s = 0.0f;
        for (int j = 0; j < Nm; ++j) s = fmaf((dis[i] * A[i][j]) * dis[j], xb[j * ldxw + f], s);
        ob[i * ldo + f] = dense_act(s + bias[c * Fo + f], act);
    }
}

// ---------------------------------------------------------------------------------------
// EdgeLayer attention (edge_network_dense.py:66-80), all channels in one launch:
//   A_ij = mean_h tanh( <Q_i[h], K_j[h]> / sqrt(ds) ),  S = (A + A^T) / 2
//   pair[b,i,j, c] = S_ij,  pair[b,i,j, C + c] = adjc[b,c,i,j]      (mlp_in of EdgeNetwork_dense, :120)
// Q,K: [B,Nm,ldq] with channel c at columns c*W .. c*W+W-1 (W = H*ds).  grid (B, C).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_attn_kernel(const float* __restrict__ Q, const float* __restrict__ K, int64_t ldq, int W, int ds,
                  const float* __restrict__ adjc, int C, int Nm, float* __restrict__ pair) {
    __shared__ float sQ[DN_MAX][33], sK[DN_MAX][33];
    __shared__ float sA[DN_MAX][DN_MAX + 1];
    const int b = blockIdx.x, c = blockIdx.y;
    for (int i = threadIdx.x; i < Nm * W; i += blockDim.x) {
        const int r = i / W, k = i % W;
        sQ[r][k] = Q[(static_cast<int64_t>(b) * Nm + r) * ldq + c * W + k];
        sK[r][k] = K[(static_cast<int64_t>(b) * Nm + r) * ldq + c * W + k];
    }
    __syncthreads();
    const int H = W / ds;
    const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(ds));
    for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
        const int i = p / Nm, j = p % Nm;
        float s = 0.0f;
        for (int h = 0; h < H; ++h) {
            float d = 0.0f;
            for (int k = 0; k < ds; ++k) d = fmaf(sQ[i][h * ds + k], sK[j][h * ds + k], d);
            s += tanhf(d * inv_sqrt);
        }
        sA[i][j] = s / static_cast<float>(H);
    }
    __syncthreads();
    const float* a = adjc + (static_cast<int64_t>(b) * C + c) * Nm * Nm;
    float* pb = pair + static_cast<int64_t>(b) * Nm * Nm * (2 * C);
    for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
        const int i = p / Nm, j = p % Nm;
        pb[static_cast<int64_t>(p) * (2 * C) + c] = (sA[i][j] + sA[j][i]) * 0.5f;
        pb[static_cast<int64_t>(p) * (2 * C) + C + c] = a[p];
    }
}

// ---------------------------------------------------------------------------------------
// symmetrise + mask the per-pair MLP output (edge_network_dense.py:124-126) and fan it out:
//   v = (m[b,i,j,c] + m[b,j,i,c]) * f_i * f_j  ->  adjc_next[b,c,i,j]  and  allc[b,i,j, all_off + c]
// ---------------------------------------------------------------------------------------
__global__ void pair_post_kernel(const float* __restrict__ m, const float* __restrict__ flags, int B, int Nm, int Co,
                                 float* __restrict__ adjc_next, float* __restrict__ allc, int ld_all, int all_off) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm * Co) return;
    const int c = static_cast<int>(idx % Co);
    const int64_t p = idx / Co;
    const int j = static_cast<int>(p % Nm), i = static_cast<int>((p / Nm) % Nm), b = static_cast<int>(p / (Nm * Nm));
    const int64_t base = static_cast<int64_t>(b) * Nm * Nm;
    const float v = ((m[(base + i * Nm + j) * Co + c] + m[(base + j * Nm + i) * Co + c]) * flags[b * Nm + j]) * flags[b * Nm + i];
    adjc_next[((static_cast<int64_t>(b) * Co + c) * Nm + i) * Nm + j] = v;
    allc[(base + i * Nm + j) * ld_all + all_off + c] = v;
}

// final edge score (invariant_scorenetwork_dense.py:84-91) and the -1/std scaling of get_score_fn (:83,93):
//   out[b,i,j] = raw[b,i,j] * (i != j) * f_i * f_j * scale[b]        (scale == NULL -> 1)
__global__ void edge_final_kernel(const float* __restrict__ raw, const float* __restrict__ flags,
                                  const float* __restrict__ scale, int B, int Nm, float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm) return;
    const int j = static_cast<int>(idx % Nm), i = static_cast<int>((idx / Nm) % Nm), b = static_cast<int>(idx / (Nm * Nm));
    float v = (i == j) ? 0.0f : raw[idx];
    v = (v * flags[b * Nm + i]) * flags[b * Nm + j];
    out[idx] = scale ? v * scale[b] : v;
}

}  // namespace molsde

using namespace molsde;

static inline unsigned nblk(int64_t n, int t) { return static_cast<unsigned>((n + t - 1) / t); }

extern "C" {

int molsde_to_dense_batch(const float* x, const int32_t* node_ptr, int32_t B, int32_t Nm, int32_t F, float* out, int64_t ldo,
                          void* stream) {
    if (!x || !node_ptr || !out || B <= 0 || Nm <= 0 || F <= 0) return MOLSDE_ERR_INVALID;
    to_dense_batch_kernel<<<nblk(static_cast<int64_t>(B) * Nm * F, 256), 256, 0, as_stream(stream)>>>(x, node_ptr, B, Nm, F, out, ldo);
    return check_launch("to_dense_batch");
}

int molsde_to_dense_adj(const int64_t* edge_index, int64_t E, const float* val, const int64_t* val_i64, float val_add,
                        const int32_t* node_ptr, int32_t B, int32_t Nm, float* adj, void* stream) {
    if (!node_ptr || !adj || B <= 0 || Nm <= 0 || (E > 0 && (!edge_index || (!val && !val_i64)))) return MOLSDE_ERR_INVALID;
    cudaError_t err = cudaMemsetAsync(adj, 0, sizeof(float) * static_cast<size_t>(B) * Nm * Nm, as_stream(stream));
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    if (E == 0) return MOLSDE_OK;
    to_dense_adj_kernel<<<nblk(E, 256), 256, 0, as_stream(stream)>>>(edge_index, E, val, val_i64, val_add, node_ptr, B, Nm, adj);
    return check_launch("to_dense_adj");
}

int molsde_node_flags(const float* adj, int32_t B, int32_t Nm, float eps, float* flags, void* stream) {
    if (!adj || !flags || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    node_flags_kernel<<<nblk(static_cast<int64_t>(B) * Nm, 128), 128, 0, as_stream(stream)>>>(adj, B, Nm, eps, flags);
    return check_launch("node_flags");
}

int molsde_grouped_linear(const float* X, int64_t rows, int64_t ldx, const float* W, const float* b, int32_t G, int32_t Ki,
                          int32_t No, float* Y, int64_t ldy, int32_t act, void* stream) {
    if (!X || !W || !Y || rows < 0 || G <= 0 || Ki <= 0 || No <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    grouped_linear_kernel<<<nblk(rows * G * No, 256), 256, 0, as_stream(stream)>>>(X, rows, ldx, W, b, G, Ki, No, Y, ldy, act);
    return check_launch("grouped_linear");
}

int molsde_dense_pow2(const float* adj, int32_t B, int32_t Nm, float* adjc, float* allc, int32_t ld_all, int32_t all_off,
                      void* stream) {
    if (!adj || !adjc || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    if (Nm > DN_MAX) return MOLSDE_ERR_UNSUPPORTED;
    pow2_kernel<<<B, 256, 0, as_stream(stream)>>>(adj, Nm, adjc, allc, ld_all, all_off);
    return check_launch("dense_pow2");
}

int molsde_dense_gcn(const float* adjc, int64_t adj_stride_b, int64_t adj_stride_c, int32_t B, int32_t C, int32_t Nm,
                     const float* xw, int64_t ldxw, const float* bias, int32_t Fo, float* out, int64_t ldo, int32_t out_off,
                     int32_t act, void* stream) {
    if (!adjc || !xw || !bias || !out || B <= 0 || C <= 0 || Nm <= 0 || Fo <= 0) return MOLSDE_ERR_INVALID;
    if (Nm > DN_MAX) return MOLSDE_ERR_UNSUPPORTED;
    dense_gcn_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(adjc, adj_stride_b, adj_stride_c, Nm, xw, ldxw, bias, Fo, out,
                                                              ldo, out_off, act);
    return check_launch("dense_gcn");
}

int molsde_dense_attn(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, const float* adjc, int32_t B,
                      int32_t C, int32_t Nm, float* pair, void* stream) {
    if (!Q || !K || !adjc || !pair || B <= 0 || C <= 0 || Nm <= 0 || W <= 0 || ds <= 0 || W % ds) return MOLSDE_ERR_INVALID;
    if (Nm > DN_MAX || W > 32) return MOLSDE_ERR_UNSUPPORTED;
    dense_attn_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(Q, K, ldq, W, ds, adjc, C, Nm, pair);
    return check_launch("dense_attn");
}

int molsde_dense_pair_post(const float* m, const float* flags, int32_t B, int32_t Nm, int32_t Co, float* adjc_next, float* allc,
                           int32_t ld_all, int32_t all_off, void* stream) {
    if (!m || !flags || !adjc_next || !allc || B <= 0 || Nm <= 0 || Co <= 0) return MOLSDE_ERR_INVALID;
    pair_post_kernel<<<nblk(static_cast<int64_t>(B) * Nm * Nm * Co, 256), 256, 0, as_stream(stream)>>>(m, flags, B, Nm, Co, adjc_next,
                                                                                                 allc, ld_all, all_off);
    return check_launch("dense_pair_post");
}

int molsde_dense_edge_final(const float* raw, const float* flags, const float* scale, int32_t B, int32_t Nm, float* out,
                            void* stream) {
    if (!raw || !flags || !out || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    edge_final_kernel<<<nblk(static_cast<int64_t>(B) * Nm * Nm, 256), 256, 0, as_stream(stream)>>>(raw, flags, scale, B, Nm, out);
    return check_launch("dense_edge_final");
}

}  // extern "C"

// ---------------------------------------------------------------------------------------
// Perturbation prologue and loss epilogue of SDEModel3Dto2D_node_adj_dense.forward (:134-152, :160-179) and the
// elementwise parts of the 3D->2D predictor-corrector sampler
// (examples/pretrain_MoleculeSDE_inference_3D_to_2D_VE_VP.py:167-252).
// ---------------------------------------------------------------------------------------
namespace molsde {

// gen_noise(sym=True) (:532-538): z = triu(raw,1) + triu(raw,1)^T, masked by flags
__global__ void sym_noise_kernel(const float* __restrict__ raw, const float* __restrict__ flags, int B, int Nm,
                                 float* __restrict__ z) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm) return;
    const int j = static_cast<int>(idx % Nm), i = static_cast<int>((idx / Nm) % Nm), b = static_cast<int>(idx / (Nm * Nm));
    const int64_t base = static_cast<int64_t>(b) * Nm * Nm;
    float v = 0.0f;
    if (i < j) v = raw[base + i * Nm + j];
    else if (i > j) v = raw[base + j * Nm + i];
    z[idx] = (v * flags[b * Nm + i]) * flags[b * Nm + j];
}

// out = ((coef[b] * x + std[b] * z) * f_i) * f_j     perturbed adjacency (:136-138)
__global__ void perturb_adj_kernel(const float* __restrict__ x, const float* __restrict__ z, const float* __restrict__ flags,
                                   const float* __restrict__ coef, const float* __restrict__ stdv, int B, int Nm,
                                   float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm) return;
    const int j = static_cast<int>(idx % Nm), i = static_cast<int>((idx / Nm) % Nm), b = static_cast<int>(idx / (Nm * Nm));
    const float v = __fadd_rn(__fmul_rn(coef[b], x[idx]), __fmul_rn(stdv[b], z[idx]));
    out[idx] = (v * flags[b * Nm + i]) * flags[b * Nm + j];
}

// one-hot perturbation (:143-152): zx = raw * f;  px = (coef[b] * onehot(zidx) + std[b] * zx) * f
__global__ void perturb_onehot_kernel(const int64_t* __restrict__ zidx /*[B,Nm]*/, const float* __restrict__ raw,
                                      const float* __restrict__ flags, const float* __restrict__ coef,
                                      const float* __restrict__ stdv, int B, int Nm, int K, float* __restrict__ zx,
                                      float* __restrict__ px) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * K) return;
    const int k = static_cast<int>(idx % K);
    const int64_t r = idx / K;
    const int b = static_cast<int>(r / Nm);
    const float f = flags[r];
    const float zv = raw[idx] * f;
    const float oh = (zidx[r] == k) ? 1.0f : 0.0f;
    zx[idx] = zv;
    px[idx] = __fadd_rn(__fmul_rn(coef[b], oh), __fmul_rn(stdv[b], zv)) * f;
}

// per-graph deterministic reductions over M contiguous elements: mode 0: sqrt(sum a^2) (Frobenius norm),
// mode 1: mean((a + b)^2 * w[g])  (denoising score-matching loss per graph, :163-177)
__global__ void __launch_bounds__(256)
graph_reduce_kernel(const float* __restrict__ a, const float* __restrict__ bb, const float* __restrict__ w, int64_t M,
                    int mode, float* __restrict__ out) {
    __shared__ float red[8];
    const int g = blockIdx.x;
    const float* pa = a + g * M;
    const float* pb = bb ? bb + g * M : nullptr;
    float s = 0.0f;
    for (int64_t i = threadIdx.x; i < M; i += blockDim.x) {
        const float v = mode == 0 ? pa[i] : pa[i] + pb[i];
        s = fmaf(v, v, s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.0f;
        for (int q = 0; q < 8; ++q) t += red[q];
        out[g] = mode == 0 ? sqrtf(t) : (t / static_cast<float>(M)) * (w ? w[g] : 1.0f);
    }
}

// Langevin step size (inference_3D_to_2D:239-241):  step[b] = (snr * mean(noise_norm) / mean(grad_norm))^2 * 2 * alpha[b]
__global__ void langevin_step_kernel(const float* __restrict__ gnorm, const float* __restrict__ nnorm,
                                     const float* __restrict__ alpha, int B, float snr, float* __restrict__ step) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        float g = 0.0f, n = 0.0f;
        for (int b = 0; b < B; ++b) { g += gnorm[b]; n += nnorm[b]; }
        g /= static_cast<float>(B);
        n /= static_cast<float>(B);
        const float r = __fdiv_rn(__fmul_rn(snr, n), g);
        for (int b = 0; b < B; ++b) step[b] = __fmul_rn(__fmul_rn(__fmul_rn(r, r), 2.0f), alpha ? alpha[b] : 1.0f);
    }
}

// corrector: x_mean = x + step[b] * grad;  x = x_mean + sqrt(2 step[b]) * noise * seps      (:242-243, :250-251)
__global__ void langevin_update_kernel(const float* __restrict__ x, const float* __restrict__ grad,
                                       const float* __restrict__ noise, const float* __restrict__ step, int64_t M, int64_t total,
                                       float seps, float* __restrict__ x_new, float* __restrict__ x_mean) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const float st = step[idx / M];
    const float xm = __fadd_rn(x[idx], __fmul_rn(st, grad[idx]));
    x_mean[idx] = xm;
    x_new[idx] = __fadd_rn(xm, __fmul_rn(__fmul_rn(sqrtf(__fmul_rn(st, 2.0f)), noise[idx]), seps));
}

// predictor (:172-190, SDE_dense.py:97-105,158-166,218-225):
//   f = sqrt_alpha[b] x - x;  rev_f = f - G[b]^2 score;  x_mean = x - rev_f;  x = x_mean + G[b] z
__global__ void reverse_update_kernel(const float* __restrict__ x, const float* __restrict__ score,
                                      const float* __restrict__ z, const float* __restrict__ sqrt_alpha,
                                      const float* __restrict__ G, int64_t M, int64_t total, float* __restrict__ x_new,
                                      float* __restrict__ x_mean) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= total) return;
    const int64_t b = idx / M;
    const float xv = x[idx], g = G[b];
    const float f = __fsub_rn(__fmul_rn(sqrt_alpha[b], xv), xv);
    const float rev_f = __fsub_rn(f, __fmul_rn(__fmul_rn(g, g), score[idx]));
    const float xm = __fsub_rn(xv, rev_f);
    x_mean[idx] = xm;
    x_new[idx] = __fadd_rn(xm, __fmul_rn(g, z[idx]));
}

// out[r, :] = x[r, :] * flags[r]      (mask_x, :559-562)
__global__ void mask_rows_kernel(const float* __restrict__ x, const float* __restrict__ flags, int64_t rows, int cols,
                                 float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows * cols) return;
    out[idx] = x[idx] * flags[idx / cols];
}

}  // namespace molsde

extern "C" {

int molsde_dense_sym_noise(const float* raw, const float* flags, int32_t B, int32_t Nm, float* z, void* stream) {
    if (!raw || !flags || !z || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    sym_noise_kernel<<<nblk(static_cast<int64_t>(B) * Nm * Nm, 256), 256, 0, as_stream(stream)>>>(raw, flags, B, Nm, z);
    return check_launch("dense_sym_noise");
}

int molsde_dense_perturb_adj(const float* x, const float* z, const float* flags, const float* coef, const float* stdv, int32_t B,
                             int32_t Nm, float* out, void* stream) {
    if (!x || !z || !flags || !coef || !stdv || !out || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    perturb_adj_kernel<<<nblk(static_cast<int64_t>(B) * Nm * Nm, 256), 256, 0, as_stream(stream)>>>(x, z, flags, coef, stdv, B, Nm, out);
    return check_launch("dense_perturb_adj");
}

int molsde_dense_perturb_onehot(const int64_t* zidx, const float* raw, const float* flags, const float* coef, const float* stdv,
                                int32_t B, int32_t Nm, int32_t K, float* zx, float* px, void* stream) {
    if (!zidx || !raw || !flags || !coef || !stdv || !zx || !px || B <= 0 || Nm <= 0 || K <= 0) return MOLSDE_ERR_INVALID;
    perturb_onehot_kernel<<<nblk(static_cast<int64_t>(B) * Nm * K, 256), 256, 0, as_stream(stream)>>>(zidx, raw, flags, coef, stdv, B,
                                                                                               Nm, K, zx, px);
    return check_launch("dense_perturb_onehot");
}

int molsde_graph_reduce(const float* a, const float* b, const float* w, int32_t B, int64_t M, int32_t mode, float* out,
                        void* stream) {
    if (!a || !out || B <= 0 || M <= 0 || (mode == 1 && !b) || mode < 0 || mode > 1) return MOLSDE_ERR_INVALID;
    graph_reduce_kernel<<<B, 256, 0, as_stream(stream)>>>(a, b, w, M, mode, out);
    return check_launch("graph_reduce");
}

int molsde_langevin_step(const float* gnorm, const float* nnorm, const float* alpha, int32_t B, float snr, float* step,
                         void* stream) {
    if (!gnorm || !nnorm || !step || B <= 0) return MOLSDE_ERR_INVALID;
    langevin_step_kernel<<<1, 32, 0, as_stream(stream)>>>(gnorm, nnorm, alpha, B, snr, step);
    return check_launch("langevin_step");
}

int molsde_langevin_update(const float* x, const float* grad, const float* noise, const float* step, int32_t B, int64_t M,
                           float seps, float* x_new, float* x_mean, void* stream) {
    if (!x || !grad || !noise || !step || !x_new || !x_mean || B <= 0 || M <= 0) return MOLSDE_ERR_INVALID;
    langevin_update_kernel<<<nblk(B * M, 256), 256, 0, as_stream(stream)>>>(x, grad, noise, step, M, B * M, seps, x_new, x_mean);
    return check_launch("langevin_update");
}

int molsde_reverse_update(const float* x, const float* score, const float* z, const float* sqrt_alpha, const float* G, int32_t B,
                          int64_t M, float* x_new, float* x_mean, void* stream) {
    if (!x || !score || !z || !sqrt_alpha || !G || !x_new || !x_mean || B <= 0 || M <= 0) return MOLSDE_ERR_INVALID;
    reverse_update_kernel<<<nblk(B * M, 256), 256, 0, as_stream(stream)>>>(x, score, z, sqrt_alpha, G, M, B * M, x_new, x_mean);
    return check_launch("reverse_update");
}

int molsde_mask_rows(const float* x, const float* flags, int64_t rows, int32_t cols, float* out, void* stream) {
    if (!x || !flags || !out || rows <= 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    mask_rows_kernel<<<nblk(rows * cols, 256), 256, 0, as_stream(stream)>>>(x, flags, rows, cols, out);
    return check_launch("mask_rows");
}

}  // extern "C"
