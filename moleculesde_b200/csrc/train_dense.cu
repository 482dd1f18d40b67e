// Backward kernels of the dense 3D->2D score networks (SDE_model_3D_to_2D_node_adj_dense.py:101-179,
// invariant_scorenetwork_dense.py, layers/edge_network_dense.py, layers/node_network_dense.py); the forward kernels are
// in dense.cu.  One CTA per (graph, channel) with the Nm x Nm (Nm <= 64) operands in shared memory; deterministic.
#include "common.cuh"

namespace molsde {

constexpr int TDN = 64;  // max padded atoms per graph (dense.cu DN_MAX)

// -----------------------------------------------------------------------------------------------------
// NodeNetwork_dense backward (node_network_dense.py:63-85), forward = dense_gcn_kernel:
//   A~ = adj with unit diagonal, s_i = sum_j A~_ij, dis = clamp(s,1)^-1/2, Ahat_ij = dis_i A~_ij dis_j,
//   out = act(Ahat . xw + bias)
// Given dout (and out for tanh'):  dpre = dout * act'(.)  [written out: its column sum is dbias],
//   dxw_j = sum_i Ahat_ij dpre_i,    and, if dadj != NULL, the gradient w.r.t. the off-diagonal adjacency entries.
// -----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
dense_gcn_bwd_kernel(const float* __restrict__ adjc, int64_t adj_stride_b, int64_t adj_stride_c, int Nm, const float* __restrict__ xw,
                     int64_t ldxw, int Fo, const float* __restrict__ out, const float* __restrict__ dout, int64_t ldo, int out_off,
                     int act, float* __restrict__ dpre_g, float* __restrict__ dxw, int64_t lddx, float* __restrict__ dadj,
                     int64_t dadj_stride_b, int dadj_accumulate) {
    __shared__ float A[TDN][TDN + 1];
    __shared__ float G[TDN][TDN + 1];
    __shared__ float dis[TDN], ssum[TDN], ddis[TDN];
    __shared__ float dp[TDN][17];
    __shared__ float sx[TDN][17];
    const int b = blockIdx.x, c = blockIdx.y, C = gridDim.y;
    const float* a = adjc + b * adj_stride_b + c * adj_stride_c;
    for (int i = threadIdx.x; i < Nm * Nm; i += blockDim.x) {
        const int r = i / Nm, cc = i % Nm;
        A[r][cc] = (r == cc) ? 1.0f : a[i];
    }
    for (int p = threadIdx.x; p < Nm * Fo; p += blockDim.x) {
        const int i = p / Fo, f = p % Fo;
        const int64_t row = static_cast<int64_t>(b) * Nm + i;
        float d = dout[row * ldo + out_off + c * Fo + f];
        if (act == 4) { const float o = out[row * ldo + out_off + c * Fo + f]; d *= 1.0f - o * o; }
        dp[i][f] = d;
        dpre_g[row * (static_cast<int64_t>(C) * Fo) + c * Fo + f] = d;
        sx[i][f] = xw[row * ldxw + c * Fo + f];
    }
    __syncthreads();
    if (threadIdx.x < Nm) {
        float s = 0.0f;
        for (int j = 0; j < Nm; ++j) s += A[threadIdx.x][j];
        ssum[threadIdx.x] = s;
        dis[threadIdx.x] = 1.0f / sqrtf(fmaxf(s, 1.0f));
    }
    __syncthreads();
    for (int p = threadIdx.x; p < Nm * Fo; p += blockDim.x) {
        const int j = p / Fo, f = p % Fo;
        float s = 0.0f;
        for (int i = 0; i < Nm; ++i) s = fmaf((dis[i] * A[i][j]) * dis[j], dp[i][f], s);
        dxw[(static_cast<int64_t>(b) * Nm + j) * lddx + c * Fo + f] = s;
    }
    if (!dadj) return;
    for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
        const int i = p / Nm, j = p % Nm;
        float g = 0.0f;
        for (int f = 0; f < Fo; ++f) g = fmaf(dp[i][f], sx[j][f], g);
        G[i][j] = g;  // d loss / d Ahat_ij
    }
    __syncthreads();
    if (threadIdx.x < Nm) {
        const int i = threadIdx.x;
        float d = 0.0f;
        for (int j = 0; j < Nm; ++j) d += G[i][j] * A[i][j] * dis[j] + G[j][i] * A[j][i] * dis[j];
        // dis = clamp(s, 1)^-1/2 : d dis / d s = -1/2 s^-3/2 where s >= 1, else 0
        ddis[i] = (ssum[i] >= 1.0f) ? d * (-0.5f) * dis[i] * dis[i] * dis[i] : 0.0f;
    }
    __syncthreads();
    float* da = dadj + b * dadj_stride_b + static_cast<int64_t>(c) * Nm * Nm;
    for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
        const int i = p / Nm, j = p % Nm;
        const float v = (i == j) ? 0.0f : G[i][j] * dis[i] * dis[j] + ddis[i];
        da[p] = dadj_accumulate ? da[p] + v : v;
    }
}

// -----------------------------------------------------------------------------------------------------
// EdgeLayer attention backward (forward = dense_attn_kernel): S = (A + A^T)/2, A_ij = mean_h tanh(<Q_i[h],K_j[h]>/sqrt(ds)).
//   dpair [B,Nm,Nm,2C]: column c is dS, column C+c is the pass-through gradient of adjc[b,c].
// Outputs dQ/dK in the layout of the forward's qk buffer and dadjc (pass-through part, overwrite).
// -----------------------------------------------------------------------------------------------------
// Per head h the factor  dU_ij = dA_ij / H * (1 - tanh^2(<Q_i[h],K_j[h]>/sqrt(ds))) / sqrt(ds)  is shared by the ds columns of
// that head, so it is tabulated once per head in shared memory (Nm^2 tanh per head instead of Nm^2 * ds per output row pass);
// the output sums run over o in ascending order with the same fmaf sequence as before (bit-identical, deterministic).
// Dynamic shared memory: sQ, sK [Nm][33] | dA, dU [Nm][Nm+1].
__global__ void __launch_bounds__(256)
dense_attn_bwd_kernel(const float* __restrict__ Q, const float* __restrict__ K, int64_t ldq, int W, int ds, int C, int Nm,
                      const float* __restrict__ dpair, float* __restrict__ dQ, float* __restrict__ dK, float* __restrict__ dadjc) {
    extern __shared__ float smem_attn[];
    const int ldp = Nm + 1;
    float* sQ = smem_attn;              // [Nm][33]
    float* sK = sQ + Nm * 33;           // [Nm][33]
    float* dA = sK + Nm * 33;           // [Nm][Nm+1]
    float* dU = dA + Nm * ldp;          // [Nm][Nm+1]
    const int b = blockIdx.x, c = blockIdx.y;
    for (int i = threadIdx.x; i < Nm * W; i += blockDim.x) {
        const int r = i / W, k = i % W;
        sQ[r * 33 + k] = Q[(static_cast<int64_t>(b) * Nm + r) * ldq + c * W + k];
        sK[r * 33 + k] = K[(static_cast<int64_t>(b) * Nm + r) * ldq + c * W + k];
    }
    const float* pb = dpair + static_cast<int64_t>(b) * Nm * Nm * (2 * C);
    float* da = dadjc ? dadjc + (static_cast<int64_t>(b) * C + c) * Nm * Nm : nullptr;
    for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
        const int i = p / Nm, j = p % Nm;
        dA[i * ldp + j] = 0.5f * (pb[static_cast<int64_t>(p) * (2 * C) + c] + pb[static_cast<int64_t>(j * Nm + i) * (2 * C) + c]);
        if (da) da[p] = pb[static_cast<int64_t>(p) * (2 * C) + C + c];
    }
    __syncthreads();
    const int H = W / ds;
    const float inv_sqrt = 1.0f / sqrtf(static_cast<float>(ds)), inv_h = 1.0f / static_cast<float>(H);
    for (int h = 0; h < H; ++h) {
        for (int p = threadIdx.x; p < Nm * Nm; p += blockDim.x) {
            const int i = p / Nm, j = p % Nm;
            float d = 0.0f;
            for (int k = 0; k < ds; ++k) d = fmaf(sQ[i * 33 + h * ds + k], sK[j * 33 + h * ds + k], d);
            const float t = tanhf(d * inv_sqrt);
            dU[i * ldp + j] = dA[i * ldp + j] * inv_h * (1.0f - t * t) * inv_sqrt;
        }
        __syncthreads();
        for (int p = threadIdx.x; p < 2 * Nm * ds; p += blockDim.x) {
            const bool forK = p >= Nm * ds;
            const int q = forK ? p - Nm * ds : p;
            const int r = q / ds, col = h * ds + q % ds;
            float acc = 0.0f;
            if (forK) {
                for (int o = 0; o < Nm; ++o) acc = fmaf(dU[o * ldp + r], sQ[o * 33 + col], acc);
            } else {
                for (int o = 0; o < Nm; ++o) acc = fmaf(dU[r * ldp + o], sK[o * 33 + col], acc);
            }
            float* dst = forK ? dK : dQ;
            dst[(static_cast<int64_t>(b) * Nm + r) * ldq + c * W + col] = acc;
        }
        __syncthreads();
    }
}

// pair_post backward: v_ij = (m_ij + m_ji) f_i f_j  ->  dm_ij = (dv_ij + dv_ji) f_i f_j,
//   dv_ij = dadjc_next[b,c,i,j] (NULL for the last layer) + dallc[b,i,j, all_off + c]
__global__ void pair_post_bwd_kernel(const float* __restrict__ dadjc_next, const float* __restrict__ dallc, int ld_all, int all_off,
                                     const float* __restrict__ flags, int B, int Nm, int Co, float* __restrict__ dm) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm * Co) return;
    const int c = static_cast<int>(idx % Co);
    const int64_t p = idx / Co;
    const int j = static_cast<int>(p % Nm), i = static_cast<int>((p / Nm) % Nm), b = static_cast<int>(p / (Nm * Nm));
    const int64_t base = static_cast<int64_t>(b) * Nm * Nm;
    float dv = dallc[(base + i * Nm + j) * ld_all + all_off + c] + dallc[(base + j * Nm + i) * ld_all + all_off + c];
    if (dadjc_next) {
        const int64_t cb = (static_cast<int64_t>(b) * Co + c) * Nm * Nm;
        dv += dadjc_next[cb + i * Nm + j] + dadjc_next[cb + j * Nm + i];
    }
    dm[idx] = dv * flags[b * Nm + i] * flags[b * Nm + j];
}

// edge_final backward: draw_ij = dout_ij * (i != j) f_i f_j scale_b
__global__ void edge_final_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ flags, const float* __restrict__ scale,
                                      int B, int Nm, float* __restrict__ draw) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<int64_t>(B) * Nm * Nm) return;
    const int j = static_cast<int>(idx % Nm), i = static_cast<int>((idx / Nm) % Nm), b = static_cast<int>(idx / (Nm * Nm));
    float v = (i == j) ? 0.0f : dout[idx];
    v = v * flags[b * Nm + i] * flags[b * Nm + j];
    draw[idx] = scale ? v * scale[b] : v;
}

// DSM loss (graph_reduce mode 1 then mean over graphs) backward:  da = coef * 2 (a + b) w[g] / (M B)
__global__ void graph_mse_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ w, int B,
                                     int64_t M, float coef, float* __restrict__ da) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= B * M) return;
    const int g = static_cast<int>(idx / M);
    da[idx] = coef * 2.0f * (a[idx] + b[idx]) * (w ? w[g] : 1.0f) / (static_cast<float>(M) * static_cast<float>(B));
}

// to_dense_batch backward: dx[n,:] = ddense[b, n - node_ptr[b], :]
__global__ void from_dense_batch_kernel(const float* __restrict__ dense, int64_t ldd, const int32_t* __restrict__ node_ptr,
                                        const int32_t* __restrict__ node2graph, int64_t N, int Nm, int F, float* __restrict__ x) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * F) return;
    const int64_t n = idx / F;
    const int f = static_cast<int>(idx % F);
    const int g = node2graph[n];
    x[idx] = dense[(static_cast<int64_t>(g) * Nm + (n - node_ptr[g])) * ldd + f];
}

// dst[r, 0:cols] = src[r, 0:cols] with independent row strides (concat / slice without torch)
__global__ void copy2d_kernel(const float* __restrict__ src, int64_t lds, float* __restrict__ dst, int64_t ldd, int64_t rows, int cols,
                              int accumulate) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= rows * cols) return;
    const int64_t r = idx / cols;
    const int c = static_cast<int>(idx % cols);
    const float v = src[r * lds + c];
    dst[r * ldd + c] = accumulate ? dst[r * ldd + c] + v : v;
}
__global__ void __launch_bounds__(256) mean1_kernel(const float* __restrict__ v, int64_t n, float* __restrict__ out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int64_t i = threadIdx.x; i < n; i += 256) acc += v[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 256; ++q) t += red[q];
        out[0] = static_cast<float>(t / static_cast<double>(n));
    }
}

// out[i] = sum_s X[s*n + i] in ascending s
__global__ void sum_slices_kernel(const float* __restrict__ X, int slices, int64_t n, float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float v = 0.0f;
    for (int s = 0; s < slices; ++s) v += X[static_cast<size_t>(s) * n + i];
    out[i] = v;
}

}  // namespace molsde

using namespace molsde;

static inline unsigned nb(int64_t n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

extern "C" {

int molsde_dense_gcn_bwd(const float* adjc, int64_t adj_stride_b, int64_t adj_stride_c, int32_t B, int32_t C, int32_t Nm,
                         const float* xw, int64_t ldxw, int32_t Fo, const float* out, const float* dout, int64_t ldo, int32_t out_off,
                         int32_t act, float* dpre, float* dxw, int64_t lddx, float* dadj, int64_t dadj_stride_b, int32_t dadj_accumulate,
                         void* stream) {
    if (!adjc || !xw || !out || !dout || !dpre || !dxw || B <= 0 || C <= 0 || Nm <= 0 || Fo <= 0) return MOLSDE_ERR_INVALID;
    if (Nm > TDN || Fo > 16 || (act != 0 && act != 4)) return MOLSDE_ERR_UNSUPPORTED;
    dense_gcn_bwd_kernel<<<dim3(B, C), 256, 0, as_stream(stream)>>>(adjc, adj_stride_b, adj_stride_c, Nm, xw, ldxw, Fo, out, dout, ldo,
                                                                  out_off, act, dpre, dxw, lddx, dadj, dadj_stride_b, dadj_accumulate);
    return check_launch("dense_gcn_bwd");
}

int molsde_dense_attn_bwd(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, int32_t B, int32_t C, int32_t Nm,
                          const float* dpair, float* dQ, float* dK, float* dadjc, void* stream) {
    if (!Q || !K || !dpair || !dQ || !dK || B <= 0 || C <= 0 || Nm <= 0 || W <= 0 || ds <= 0 || W % ds) return MOLSDE_ERR_INVALID;
    if (Nm > TDN || W > 32) return MOLSDE_ERR_UNSUPPORTED;
    const size_t smem = sizeof(float) * (2 * static_cast<size_t>(Nm) * 33 + 2 * static_cast<size_t>(Nm) * (Nm + 1));
    static bool attr_set = false;   // 50.2 KB at Nm = 64: above the 48 KB default, opt in once
    if (!attr_set) {
        if (cudaFuncSetAttribute(dense_attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 static_cast<int>(sizeof(float) * (2 * TDN * 33 + 2 * TDN * (TDN + 1)))) != cudaSuccess)
            return check_launch("dense_attn_bwd(attr)");
        attr_set = true;
    }
    dense_attn_bwd_kernel<<<dim3(B, C), 256, smem, as_stream(stream)>>>(Q, K, ldq, W, ds, C, Nm, dpair, dQ, dK, dadjc);
    return check_launch("dense_attn_bwd");
}

int molsde_dense_pair_post_bwd(const float* dadjc_next, const float* dallc, int32_t ld_all, int32_t all_off, const float* flags,
                               int32_t B, int32_t Nm, int32_t Co, float* dm, void* stream) {
    if (!dallc || !flags || !dm || B <= 0 || Nm <= 0 || Co <= 0) return MOLSDE_ERR_INVALID;
    pair_post_bwd_kernel<<<nb(static_cast<int64_t>(B) * Nm * Nm * Co), 256, 0, as_stream(stream)>>>(dadjc_next, dallc, ld_all, all_off,
                                                                                              flags, B, Nm, Co, dm);
    return check_launch("dense_pair_post_bwd");
}

int molsde_dense_edge_final_bwd(const float* dout, const float* flags, const float* scale, int32_t B, int32_t Nm, float* draw,
                                void* stream) {
    if (!dout || !flags || !draw || B <= 0 || Nm <= 0) return MOLSDE_ERR_INVALID;
    edge_final_bwd_kernel<<<nb(static_cast<int64_t>(B) * Nm * Nm), 256, 0, as_stream(stream)>>>(dout, flags, scale, B, Nm, draw);
    return check_launch("dense_edge_final_bwd");
}

int molsde_graph_mse_bwd(const float* a, const float* b, const float* w, int32_t B, int64_t M, float coef, float* da, void* stream) {
    if (!a || !b || !da || B <= 0 || M <= 0) return MOLSDE_ERR_INVALID;
    graph_mse_bwd_kernel<<<nb(B * M), 256, 0, as_stream(stream)>>>(a, b, w, B, M, coef, da);
    return check_launch("graph_mse_bwd");
}

int molsde_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t cols, int32_t accumulate, void* stream) {
    if (!src || !dst || rows < 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    copy2d_kernel<<<nb(rows * cols), 256, 0, as_stream(stream)>>>(src, lds, dst, ldd, rows, cols, accumulate);
    return check_launch("copy2d");
}
int molsde_sum_slices(const float* X, int32_t slices, int64_t n, float* out, void* stream) {
    if (!X || !out || slices <= 0 || n < 0) return MOLSDE_ERR_INVALID;
    if (n == 0) return MOLSDE_OK;
    sum_slices_kernel<<<nb(n), 256, 0, as_stream(stream)>>>(X, slices, n, out);
    return check_launch("sum_slices");
}
int molsde_mean(const float* v, int64_t n, float* out, void* stream) {
    if (!v || !out || n <= 0) return MOLSDE_ERR_INVALID;
    mean1_kernel<<<1, 256, 0, as_stream(stream)>>>(v, n, out);
    return check_launch("mean");
}

int molsde_from_dense_batch(const float* dense, int64_t ldd, const int32_t* node_ptr, const int32_t* node2graph, int64_t N, int32_t Nm,
                            int32_t F, float* x, void* stream) {
    if (!dense || !node_ptr || !node2graph || !x || N < 0 || Nm <= 0 || F <= 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    from_dense_batch_kernel<<<nb(N * F), 256, 0, as_stream(stream)>>>(dense, ldd, node_ptr, node2graph, N, Nm, F, x);
    return check_launch("from_dense_batch");
}

}  // extern "C"
