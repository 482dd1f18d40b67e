// Layer-granular training kernels of SDEModel2Dto3D_02 (SDE_model_2D_to_3D.py:306-391) and its EquivariantScoreNetwork
// (equivariant_scorenetwork.py): the ops whose intermediates the backward pass needs, and their backward kernels.
// Edges are in CSR-by-target order: edge e has target tgt[e] (= PyG x_i, `col`) and source src[e] (= x_j, `row`).
// hidden_dim = 32, 8 heads x 4 channels.  Deterministic: per-target loops in CSR order, no atomics.
#include "common.cuh"

namespace molsde {

constexpr float kEps = 1e-6f;  // EPSILON, SDE_model_2D_to_3D.py:10

// -------------------------------------------------------------------------------------------------
// Geometric edge features at the perturbed positions (:342-369).  One thread per (edge, frequency).
//   gfd  [E,64]  = [sin, cos](2 pi d W_dist)                         (dist_gaussian_fourier, :349)
//   gfi/gfj [E,128] = [sin,cos](2 pi c0 W_coff) | [sin,cos](2 pi c2 W_coff)   (get_embedding input, :297-303)
//   emb  [E,66] columns 0,1 = pseudo_sin, pseudo_cos  (:364-366; columns 2.. are filled by coff_mlp afterwards)
//   basis [E,9]  = coord_diff | coord_cross | coord_vertical       (coord2basis, :35-47)
// No gradient flows into positions or the frozen Fourier frequencies, so this stage has no backward.
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
edge_geom_kernel(const float* __restrict__ pos, const int32_t* __restrict__ src, const int32_t* __restrict__ tgt, int64_t E,
                 const float* __restrict__ w_dist, const float* __restrict__ w_coff, float* __restrict__ gfd, float* __restrict__ gfi,
                 float* __restrict__ gfj, float* __restrict__ emb, float* __restrict__ basis) {
    const int64_t e = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (e >= E) return;
    const int r = src[e], c = tgt[e];
    const float rx = pos[3 * r], ry = pos[3 * r + 1], rz = pos[3 * r + 2];
    const float cx = pos[3 * c], cy = pos[3 * c + 1], cz = pos[3 * c + 2];
    float dx = __fsub_rn(rx, cx), dy = __fsub_rn(ry, cy), dz = __fsub_rn(rz, cz);
    const float radial = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float d = sqrtf(radial);
    float kx = __fsub_rn(__fmul_rn(ry, cz), __fmul_rn(rz, cy));
    float ky = __fsub_rn(__fmul_rn(rz, cx), __fmul_rn(rx, cz));
    float kz = __fsub_rn(__fmul_rn(rx, cy), __fmul_rn(ry, cx));
    const float nrm = d + kEps;
    dx = dx / nrm; dy = dy / nrm; dz = dz / nrm;
    const float cn = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(kx, kx), __fmul_rn(ky, ky)), __fmul_rn(kz, kz))) + kEps;
    kx = kx / cn; ky = ky / cn; kz = kz / cn;
    const float vx = __fsub_rn(__fmul_rn(dy, kz), __fmul_rn(dz, ky));
    const float vy = __fsub_rn(__fmul_rn(dz, kx), __fmul_rn(dx, kz));
    const float vz = __fsub_rn(__fmul_rn(dx, ky), __fmul_rn(dy, kx));
    // coff = Basis . r   (:357-360), |.| on the cross component
    const float ci0 = __fadd_rn(__fadd_rn(__fmul_rn(dx, rx), __fmul_rn(dy, ry)), __fmul_rn(dz, rz));
    const float ci1 = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(kx, rx), __fmul_rn(ky, ry)), __fmul_rn(kz, rz)));
    const float ci2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, rx), __fmul_rn(vy, ry)), __fmul_rn(vz, rz));
    const float cj0 = __fadd_rn(__fadd_rn(__fmul_rn(dx, cx), __fmul_rn(dy, cy)), __fmul_rn(dz, cz));
    const float cj1 = fabsf(__fadd_rn(__fadd_rn(__fmul_rn(kx, cx), __fmul_rn(ky, cy)), __fmul_rn(kz, cz)));
    const float cj2 = __fadd_rn(__fadd_rn(__fmul_rn(vx, cx), __fmul_rn(vy, cy)), __fmul_rn(vz, cz));
    if (lane == 0) {
        const float ni = sqrtf(ci0 * ci0 + ci1 * ci1 + ci2 * ci2), nj = sqrtf(cj0 * cj0 + cj1 * cj1 + cj2 * cj2);
        const float pc = (ci0 * cj0 + ci1 * cj1 + ci2 * cj2) / (ni + kEps) / (nj + kEps);
        emb[e * 66 + 0] = sqrtf(fmaxf(1.0f - pc * pc, 0.0f));  // clamp: rounding can push cos^2 above 1 (DESIGN.md deviations)
        emb[e * 66 + 1] = pc;
        float* b = basis + e * 9;
        b[0] = dx; b[1] = dy; b[2] = dz; b[3] = kx; b[4] = ky; b[5] = kz; b[6] = vx; b[7] = vy; b[8] = vz;
    }
    const float pi_f = 3.14159274101257324f;  // x * W * 2 * np.pi evaluated left to right in fp32
    const float wd = w_dist[lane], wc = w_coff[lane];
    float s, co;
    sincosf(d * wd * 2.0f * pi_f, &s, &co);
    gfd[e * 64 + lane] = s; gfd[e * 64 + 32 + lane] = co;
    sincosf(ci0 * wc * 2.0f * pi_f, &s, &co);
    gfi[e * 128 + lane] = s; gfi[e * 128 + 32 + lane] = co;
    sincosf(ci2 * wc * 2.0f * pi_f, &s, &co);
    gfi[e * 128 + 64 + lane] = s; gfi[e * 128 + 96 + lane] = co;
    sincosf(cj0 * wc * 2.0f * pi_f, &s, &co);
    gfj[e * 128 + lane] = s; gfj[e * 128 + 32 + lane] = co;
    sincosf(cj2 * wc * 2.0f * pi_f, &s, &co);
    gfj[e * 128 + 64 + lane] = s; gfj[e * 128 + 96 + lane] = co;
}

// -------------------------------------------------------------------------------------------------
// TransformerConv (heads 8 x 4, edge_dim 32, root weight) message passing, forward.  One warp per target, lane = channel.
//   logit_e[h] = <q_i, k_j + eproj_e>_h / 2;  alpha = softmax over incoming edges (max shift, +1e-16);
//   out_i = sum_e alpha_e * keep_e / (1-p) * (v_j + eproj_e) + skip_i.      alpha [E,8] is kept for the backward.
// qkvs [N,128] = [q | k | v | skip] (one fused linear), eproj [E,32] = lin_edge(edge_attr).
// -------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
tconv_fwd_kernel(const float* __restrict__ qkvs, const float* __restrict__ eproj, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ src, int64_t N, const float* __restrict__ keep, float inv_keep, float* __restrict__ alpha,
                 float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, hd = lane >> 2;
    if (i >= N) return;
    const float q = qkvs[i * 128 + lane];
    const int a = rowptr[i], b = rowptr[i + 1];
    float mx = -INFINITY;
    for (int e = a; e < b; ++e) {
        float t = q * (qkvs[static_cast<int64_t>(src[e]) * 128 + 32 + lane] + eproj[static_cast<int64_t>(e) * 32 + lane]);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        mx = fmaxf(mx, t * 0.5f);
    }
    float sum = 0.0f;
    for (int e = a; e < b; ++e) {
        float t = q * (qkvs[static_cast<int64_t>(src[e]) * 128 + 32 + lane] + eproj[static_cast<int64_t>(e) * 32 + lane]);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        sum += expf(t * 0.5f - mx);
    }
    float acc = 0.0f;
    for (int e = a; e < b; ++e) {
        const float ep = eproj[static_cast<int64_t>(e) * 32 + lane];
        float t = q * (qkvs[static_cast<int64_t>(src[e]) * 128 + 32 + lane] + ep);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        const float al = expf(t * 0.5f - mx) / (sum + 1e-16f);
        if ((lane & 3) == 0) alpha[static_cast<int64_t>(e) * 8 + hd] = al;
        const float ad = keep ? al * keep[static_cast<int64_t>(e) * 8 + hd] * inv_keep : al;
        acc = fmaf(ad, qkvs[static_cast<int64_t>(src[e]) * 128 + 64 + lane] + ep, acc);
    }
    out[i * 32 + lane] = acc + qkvs[i * 128 + 96 + lane];
}

// backward: given dout [N,32] (gradient of `out`), produce
//   dqkvs [N,128]: columns 0..31 = dq, 96..127 = dskip (= dout); columns 32..95 (dk, dv) are written by the caller's
//                  by-source reduction of dkvE;   dkvE [E,64] = [d k_j | d v_j] per edge;   deproj [E,32] = dk_e + dv_e.
__global__ void __launch_bounds__(256)
tconv_bwd_kernel(const float* __restrict__ qkvs, const float* __restrict__ eproj, const int32_t* __restrict__ rowptr,
                 const int32_t* __restrict__ src, int64_t N, const float* __restrict__ keep, float inv_keep,
                 const float* __restrict__ alpha, const float* __restrict__ dout, float* __restrict__ dqkvs, float* __restrict__ dkvE,
                 float* __restrict__ deproj) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31, hd = lane >> 2;
    if (i >= N) return;
    const float q = qkvs[i * 128 + lane], go = dout[i * 32 + lane];
    const int a = rowptr[i], b = rowptr[i + 1];
    // s[h] = sum_e alpha_e * dalpha_e,  dalpha_e = keep/(1-p) * <dout_i, v_j + eproj_e>_h
    float s = 0.0f;
    for (int e = a; e < b; ++e) {
        float t = go * (qkvs[static_cast<int64_t>(src[e]) * 128 + 64 + lane] + eproj[static_cast<int64_t>(e) * 32 + lane]);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        const float kp = keep ? keep[static_cast<int64_t>(e) * 8 + hd] * inv_keep : 1.0f;
        s = fmaf(alpha[static_cast<int64_t>(e) * 8 + hd], t * kp, s);
    }
    float dq = 0.0f;
    for (int e = a; e < b; ++e) {
        const float ep = eproj[static_cast<int64_t>(e) * 32 + lane];
        const float kj = qkvs[static_cast<int64_t>(src[e]) * 128 + 32 + lane] + ep;
        float t = go * (qkvs[static_cast<int64_t>(src[e]) * 128 + 64 + lane] + ep);
        t += __shfl_xor_sync(0xffffffffu, t, 1);
        t += __shfl_xor_sync(0xffffffffu, t, 2);
        const float kp = keep ? keep[static_cast<int64_t>(e) * 8 + hd] * inv_keep : 1.0f;
        const float al = alpha[static_cast<int64_t>(e) * 8 + hd];
        const float dlogit = al * (t * kp - s) * 0.5f;  // through softmax and the 1/sqrt(C) scale
        dq = fmaf(dlogit, kj, dq);
        const float dk = dlogit * q, dv = al * kp * go;
        dkvE[static_cast<int64_t>(e) * 64 + lane] = dk;
        dkvE[static_cast<int64_t>(e) * 64 + 32 + lane] = dv;
        deproj[static_cast<int64_t>(e) * 32 + lane] = dk + dv;
    }
    dqkvs[i * 128 + lane] = dq;
    dqkvs[i * 128 + 96 + lane] = go;
}

// dqkvs[j, 32..95] = sum over edges leaving j (by-source CSR: sptr, sperm) of dkvE[e]
__global__ void tconv_bwd_src_kernel(const float* __restrict__ dkvE, const int32_t* __restrict__ sptr, const int32_t* __restrict__ sperm,
                                     int64_t N, float* __restrict__ dqkvs) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * 64) return;
    const int64_t j = idx >> 6;
    const int c = static_cast<int>(idx & 63);
    float acc = 0.0f;
    for (int p = sptr[j]; p < sptr[j + 1]; ++p) acc += dkvE[static_cast<int64_t>(sperm[p]) * 64 + c];
    dqkvs[j * 128 + 32 + c] = acc;
}

// -------------------------------------------------------------------------------------------------
// EquiLayer (equivariant_scorenetwork.py:43-78, activation off, aggr mean):
//   grad[i] (+)= mean_{e -> i} sum_k dyn[e,k] * basis[e,k,:]          backward:  ddyn[e,k] = <basis[e,k,:], dgrad[tgt_e]> / deg
// -------------------------------------------------------------------------------------------------
__global__ void equi_fwd_kernel(const float* __restrict__ dyn, const float* __restrict__ basis, const int32_t* __restrict__ rowptr,
                                int64_t N, int accumulate, float* __restrict__ grad) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * 3) return;
    const int64_t i = idx / 3;
    const int ax = static_cast<int>(idx % 3);
    const int a = rowptr[i], b = rowptr[i + 1];
    float acc = 0.0f;
    for (int e = a; e < b; ++e) {
        const float* bs = basis + static_cast<int64_t>(e) * 9;
        const float* dn = dyn + static_cast<int64_t>(e) * 3;
        acc += __fadd_rn(__fadd_rn(__fmul_rn(dn[0], bs[ax]), __fmul_rn(dn[1], bs[3 + ax])), __fmul_rn(dn[2], bs[6 + ax]));
    }
    acc /= static_cast<float>(max(b - a, 1));
    grad[idx] = accumulate ? grad[idx] + acc : acc;
}
__global__ void equi_bwd_kernel(const float* __restrict__ dgrad, const float* __restrict__ basis, const int32_t* __restrict__ rowptr,
                                const int32_t* __restrict__ tgt, int64_t E, float* __restrict__ ddyn) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= E * 3) return;
    const int64_t e = idx / 3;
    const int k = static_cast<int>(idx % 3);
    const int i = tgt[e];
    const float inv = 1.0f / static_cast<float>(max(rowptr[i + 1] - rowptr[i], 1));
    const float* bs = basis + e * 9 + 3 * k;
    ddyn[idx] = (bs[0] * dgrad[3 * i] + bs[1] * dgrad[3 * i + 1] + bs[2] * dgrad[3 * i + 2]) * inv;
}

// d loss / d score[i,:] = 2 (score - noise) * w[i] / (n_g * B) * upstream        (:380-390)
__global__ void dsm_pos_loss_bwd_kernel(const float* __restrict__ score, const float* __restrict__ noise, const float* __restrict__ w,
                                        const int32_t* __restrict__ node_ptr, const int32_t* __restrict__ node2graph, int64_t N, int B,
                                        float upstream, float* __restrict__ dscore) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * 3) return;
    const int64_t i = idx / 3;
    const int g = node2graph[i];
    const float cnt = static_cast<float>(max(node_ptr[g + 1] - node_ptr[g], 1));
    dscore[idx] = 2.0f * (score[idx] - noise[idx]) * (w ? w[i] : 1.0f) / cnt / static_cast<float>(B) * upstream;
}

// expand a CSR row pointer into the per-edge row index
__global__ void expand_rowptr_kernel(const int32_t* __restrict__ rowptr, int64_t N, int32_t* __restrict__ row) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= N) return;
    for (int e = rowptr[i]; e < rowptr[i + 1]; ++e) row[e] = static_cast<int32_t>(i);
}

}  // namespace molsde

using namespace molsde;

static inline unsigned nblk(int64_t n, int t = 256) { return static_cast<unsigned>((n + t - 1) / t); }

extern "C" {

int molsde_expand_rowptr(const int32_t* rowptr, int64_t N, int32_t* row, void* stream) {
    if (!rowptr || !row || N < 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    expand_rowptr_kernel<<<nblk(N), 256, 0, as_stream(stream)>>>(rowptr, N, row);
    return check_launch("expand_rowptr");
}

int molsde_sde2d3d_edge_geom(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* w_dist,
                             const float* w_coff, float* gfd, float* gfi, float* gfj, float* emb, float* basis, void* stream) {
    if (!pos || !src || !tgt || !w_dist || !w_coff || !gfd || !gfi || !gfj || !emb || !basis || E < 0) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    edge_geom_kernel<<<nblk(E, 8), 256, 0, as_stream(stream)>>>(pos, src, tgt, E, w_dist, w_coff, gfd, gfi, gfj, emb, basis);
    return check_launch("edge_geom");
}

int molsde_tconv_fwd(const float* qkvs, const float* eproj, const int32_t* rowptr, const int32_t* src, int64_t N, const float* keep,
                     float dropout_p, float* alpha, float* out, void* stream) {
    if (!qkvs || !rowptr || !src || !alpha || !out || N < 0 || dropout_p < 0.0f || dropout_p >= 1.0f) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    tconv_fwd_kernel<<<nblk(N, 8), 256, 0, as_stream(stream)>>>(qkvs, eproj, rowptr, src, N, keep, 1.0f / (1.0f - dropout_p), alpha, out);
    return check_launch("tconv_fwd");
}

int molsde_tconv_bwd(const float* qkvs, const float* eproj, const int32_t* rowptr, const int32_t* src, const int32_t* sptr,
                     const int32_t* sperm, int64_t N, const float* keep, float dropout_p, const float* alpha, const float* dout,
                     float* dqkvs, float* dkvE, float* deproj, void* stream) {
    if (!qkvs || !rowptr || !src || !sptr || !sperm || !alpha || !dout || !dqkvs || !dkvE || !deproj || N < 0)
        return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    tconv_bwd_kernel<<<nblk(N, 8), 256, 0, as_stream(stream)>>>(qkvs, eproj, rowptr, src, N, keep, 1.0f / (1.0f - dropout_p), alpha,
                                                              dout, dqkvs, dkvE, deproj);
    int st = check_launch("tconv_bwd");
    if (st != MOLSDE_OK) return st;
    tconv_bwd_src_kernel<<<nblk(N * 64), 256, 0, as_stream(stream)>>>(dkvE, sptr, sperm, N, dqkvs);
    return check_launch("tconv_bwd_src");
}

int molsde_equi_fwd(const float* dyn, const float* basis, const int32_t* rowptr, int64_t N, int32_t accumulate, float* grad,
                    void* stream) {
    if (!dyn || !basis || !rowptr || !grad || N < 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    equi_fwd_kernel<<<nblk(N * 3), 256, 0, as_stream(stream)>>>(dyn, basis, rowptr, N, accumulate, grad);
    return check_launch("equi_fwd");
}
int molsde_equi_bwd(const float* dgrad, const float* basis, const int32_t* rowptr, const int32_t* tgt, int64_t E, float* ddyn,
                    void* stream) {
    if (!dgrad || !basis || !rowptr || !tgt || !ddyn || E < 0) return MOLSDE_ERR_INVALID;
    if (E == 0) return MOLSDE_OK;
    equi_bwd_kernel<<<nblk(E * 3), 256, 0, as_stream(stream)>>>(dgrad, basis, rowptr, tgt, E, ddyn);
    return check_launch("equi_bwd");
}
int molsde_dsm_pos_loss_bwd(const float* score, const float* noise, const float* w, const int32_t* node_ptr, const int32_t* node2graph,
                            int64_t N, int32_t B, float upstream, float* dscore, void* stream) {
    if (!score || !noise || !node_ptr || !node2graph || !dscore || N < 0 || B <= 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    dsm_pos_loss_bwd_kernel<<<nblk(N * 3), 256, 0, as_stream(stream)>>>(score, noise, w, node_ptr, node2graph, N, B, upstream, dscore);
    return check_launch("dsm_pos_loss_bwd");
}

}  // extern "C"
