// SchNet forward pieces (Geom3D/models/schnet.py:85-125) and the EBM_node_dot_prod contrastive
// logits/loss (examples/util.py:52-68).
//
//   schnet_cfconv_kernel  -- one launch per InteractionBlock: for every radius-graph edge, in CSR-by-target
//       tiles of <=128 edges, compute in-kernel from the distance d (GaussianSmearing is never
//       materialised):   W = (Lin(51,128) -> ssp -> Lin(128,128))(exp(coeff (d - mu_k)^2)) * 0.5 (cos(d pi / rc) + 1)
//       (schnet.py:185-187,205-207), message x_j * W (:194-195) and the deterministic ascending-source
//       sum per target (:190).  Both filter GEMMs run on the tensor cores (3xTF32, mma_tile.cuh) on
//       warp-private 16-edge stripes; x = lin1(h) is gathered from L2.
//   gather_rows / segment_reduce -- embedding lookup (:89) and per-graph readout (:115).
//   ebm_node_dot_kernel   -- row-wise dots of X with Y and with Y[perm] (HBM-bound, not a GEMM),
//       BCE-with-logits partial sums reduced in a fixed order.
#include <math_constants.h>

#include "common.cuh"
#include "mma_tile.cuh"

namespace molsde {

constexpr int SN_THREADS = 256;
constexpr int SN_TE = MOLSDE_TILE_EDGES;  // 128
constexpr int SN_LDA = MOLSDE_TILE_LD;    // 136
constexpr int SN_F = 128;                 // num_filters (config.py:66)
constexpr int SN_G = 56;                  // num_gaussians (51) padded to a multiple of 8
constexpr int SN_LDM = 132;               // slot-major message tile [128][132]

// smem float offsets
constexpr int SS_W1 = 0;                        // [56][136]  mlp.0.weight^T (rows >= num_gaussians zero)
constexpr int SS_W2 = SS_W1 + SN_G * SN_LDA;    // [128][136] mlp.2.weight^T
constexpr int SS_B1 = SS_W2 + SN_F * SN_LDA;    // [128]
constexpr int SS_B2 = SS_B1 + SN_F;             // [128]
constexpr int SS_MU = SS_B2 + SN_F;             // [64]   gaussian offsets
constexpr int SS_A1 = SS_MU + 64;               // [56][136]  gaussian features, k-major
constexpr int SS_A2 = SS_A1 + SN_G * SN_LDA;    // [128][136] hidden, k-major; later the message tile [128][132]
constexpr int SS_D = SS_A2 + SN_F * SN_LDA;     // [128] distance
constexpr int SS_C = SS_D + SN_TE;              // [128] cosine cutoff
constexpr int SS_FLOATS = SS_C + SN_TE;
constexpr int SSI_SRC = 0, SSI_TGT = SN_TE, SS_INTS = 2 * SN_TE;
constexpr size_t SN_SMEM = sizeof(float) * SS_FLOATS + sizeof(int) * SS_INTS;
static_assert(SN_SMEM <= 232448, "smem");
static_assert(SN_TE * SN_LDM <= SN_F * SN_LDA, "message tile aliases the hidden tile");

__device__ __forceinline__ float ssp_fast(float x) {
    // ShiftedSoftplus (schnet.py:210-216): softplus(x) - float32(log 2), softplus threshold 20
    const float sp = x > 20.0f ? x : __logf(1.0f + __expf(x));
    return sp - 0.69314718246459961f;
}

__global__ void __launch_bounds__(SN_THREADS, 1)
schnet_cfconv_kernel(molsde_plan plan, const float* __restrict__ pos, const float* __restrict__ x /*[N][128]*/,
                     const float* __restrict__ w1t /*[56][136]*/, const float* __restrict__ b1,
                     const float* __restrict__ w2t /*[128][136]*/, const float* __restrict__ b2,
                     const float* __restrict__ mu /*[56]*/, int num_gaussians, float coeff, float cutoff,
                     float* __restrict__ agg /*[N][128]*/) {
    extern __shared__ __align__(16) float smem[];
    int* si = reinterpret_cast<int*>(smem + SS_FLOATS);
    const int tid = threadIdx.x, lane = tid & 31, slab = tid >> 5;
    const int g = lane >> 2, t4 = lane & 3;
    for (int i = tid * 4; i < SN_G * SN_LDA; i += SN_THREADS * 4) cp_async16(smem + SS_W1 + i, w1t + i);
    for (int i = tid * 4; i < SN_F * SN_LDA; i += SN_THREADS * 4) cp_async16(smem + SS_W2 + i, w2t + i);
    cp_async_commit();
    if (tid < SN_F) { smem[SS_B1 + tid] = b1[tid]; smem[SS_B2 + tid] = b2[tid]; }
    if (tid < 64) smem[SS_MU + tid] = tid < SN_G ? mu[tid] : 0.0f;
    cp_async_wait<0>();
    __syncthreads();
    float* A1 = smem + SS_A1 + slab * 16;
    float* A2 = smem + SS_A2 + slab * 16;
    float* Mm = smem + SS_A2;
    const float pi_over_rc = 3.14159265358979323846f / cutoff;
    for (int tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
        const int ta = plan.tile_tgt_ptr[tile], tb = plan.tile_tgt_ptr[tile + 1];
        const int ea = plan.rowptr[ta], ne = plan.rowptr[tb] - ea;
        if (lane < 16) {
            const int slot = slab * 16 + lane;
            int sj = 0, tg = ta;
            float d = 0.0f, cc = 0.0f;
            if (slot < ne) {
                const int e = ea + slot;
                int lo = ta, hi = tb;
                while (hi - lo > 1) {
                    const int mid = (lo + hi) >> 1;
                    if (plan.rowptr[mid] <= e) lo = mid; else hi = mid;
                }
                tg = lo;
                sj = plan.src[e];
                // edge_weight = (pos[row] - pos[col]).norm(dim=-1)   (:93)
                const float dx = __fsub_rn(pos[3 * sj], pos[3 * tg]), dy = __fsub_rn(pos[3 * sj + 1], pos[3 * tg + 1]);
                const float dz = __fsub_rn(pos[3 * sj + 2], pos[3 * tg + 2]);
                d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
                cc = 0.5f * (cosf(d * pi_over_rc) + 1.0f);  // :186
            }
            si[SSI_SRC + slot] = sj;
            si[SSI_TGT + slot] = tg;
            smem[SS_D + slot] = d;
            smem[SS_C + slot] = cc;
        }
        __syncwarp();
        {   // GaussianSmearing (:205-207) of the warp's 16 edges into its A1 stripe
            const int fe = lane & 15, kb = (lane >> 4) * (SN_G / 2);
            const float d = smem[SS_D + slab * 16 + fe];
#pragma unroll 4
            for (int i = 0; i < SN_G / 2; ++i) {
                const int k = kb + i;
                const float u = d - smem[SS_MU + k];
                A1[k * SN_LDA + fe] = k < num_gaussians ? __expf(coeff * u * u) : 0.0f;
            }
        }
        __syncwarp();
        float acc[16][4];
        zero_frag(acc);
        mma_gemm<16, SN_LDA, SN_LDA>(A1, smem + SS_W1, SN_G, lane, acc);
#pragma unroll
        for (int nb = 0; nb < 16; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = nb * 8 + 2 * t4 + j;
                const float bj = smem[SS_B1 + col];
                A2[col * SN_LDA + g] = ssp_fast(acc[nb][j] + bj);
                A2[col * SN_LDA + g + 8] = ssp_fast(acc[nb][2 + j] + bj);
            }
        __syncwarp();
        zero_frag(acc);
        mma_gemm<16, SN_LDA, SN_LDA>(A2, smem + SS_W2, SN_F, lane, acc);
        __syncthreads();  // every warp is done reading its hidden stripe: the region becomes the message tile
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int slot = slab * 16 + g + 8 * rr;
            if (slot < ne) {
                const float cc = smem[SS_C + slot];
                const float* xr = x + static_cast<size_t>(si[SSI_SRC + slot]) * SN_F;
#pragma unroll
                for (int nb = 0; nb < 16; ++nb) {
                    const int col = nb * 8 + 2 * t4;
                    const float2 xv = __ldg(reinterpret_cast<const float2*>(xr + col));
                    // message = x_j * (nn(edge_attr) * C)   (:187,194-195)
                    const float w0 = (acc[nb][2 * rr] + smem[SS_B2 + col]) * cc;
                    const float w1 = (acc[nb][2 * rr + 1] + smem[SS_B2 + col + 1]) * cc;
                    *reinterpret_cast<float2*>(Mm + slot * SN_LDM + col) = make_float2(xv.x * w0, xv.y * w1);
                }
            }
        }
        __syncthreads();
        const int ntg = tb - ta;
        for (int p = tid; p < ntg * SN_F; p += SN_THREADS) {
            const int i = ta + (p >> 7), col = p & (SN_F - 1);
            const int s0 = plan.rowptr[i] - ea, s1 = plan.rowptr[i + 1] - ea;
            float sacc = 0.0f;
            for (int s = s0; s < s1; ++s) sacc += Mm[s * SN_LDM + col];  // ascending source order, aggr="add"
            agg[static_cast<size_t>(i) * SN_F + col] = sacc;
        }
        __syncthreads();
    }
}

// out[r, :] = table[idx[r], :]
__global__ void gather_rows_kernel(const float* __restrict__ table, const int64_t* __restrict__ idx, int64_t rows,
                                   int cols, float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols;
    const int c = static_cast<int>(i % cols);
    out[i] = table[idx[r] * cols + c];
}

// out[s, :] = sum or mean of x[ptr[s]:ptr[s+1], :] in ascending row order (torch_scatter.scatter, schnet.py:115)
__global__ void segment_reduce_kernel(const float* __restrict__ x, const int32_t* __restrict__ ptr, int32_t segments,
                                      int cols, int mean, float* __restrict__ out) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= static_cast<int64_t>(segments) * cols) return;
    const int s = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const int a = ptr[s], b = ptr[s + 1];
    float acc = 0.0f;
    for (int r = a; r < b; ++r) acc += x[static_cast<int64_t>(r) * cols + c];
    if (mean) acc = __fdiv_rn(acc, static_cast<float>(max(b - a, 1)));
    out[i] = acc;
}

// ---------------------------------------------------------------------------------------
// EBM_node_dot_prod (examples/util.py:52-68): one warp per row
//   pred_pos = sum(X*Y)/T, pred_neg = sum(X*Y[perm])/T;  partial[block] = {sum softplus(-pos), sum softplus(neg),
//   #(pos>0) + #(neg<0)};  a second single-CTA pass reduces the partials in block order.
// ---------------------------------------------------------------------------------------
constexpr int CL_WARPS = 8;

__global__ void __launch_bounds__(CL_WARPS * 32)
ebm_node_dot_kernel(const float* __restrict__ X, const float* __restrict__ Y, const int64_t* __restrict__ perm,
                    int64_t N, int D, float inv_T, float* __restrict__ pred_pos, float* __restrict__ pred_neg,
                    float* __restrict__ partial /*[grid][4]*/) {
    __shared__ float red[CL_WARPS][3];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float lp = 0.0f, ln = 0.0f, hit = 0.0f;
    for (int64_t r = static_cast<int64_t>(blockIdx.x) * CL_WARPS + warp; r < N; r += static_cast<int64_t>(gridDim.x) * CL_WARPS) {
        const float* xr = X + r * D;
        const float* yr = Y + r * D;
        const float* yn = Y + perm[r] * D;
        float sp = 0.0f, sn = 0.0f;
        for (int c = lane; c < D; c += 32) {
            const float xv = xr[c];
            sp = fmaf(xv, yr[c], sp);
            sn = fmaf(xv, yn[c], sn);
        }
        sp = warp_sum(sp) * inv_T;
        sn = warp_sum(sn) * inv_T;
        if (lane == 0) {
            pred_pos[r] = sp;
            pred_neg[r] = sn;
            // BCEWithLogits(x, 1) = softplus(-x);  BCEWithLogits(x, 0) = softplus(x)  (numerically stable form)
            lp += fmaxf(-sp, 0.0f) + log1pf(expf(-fabsf(sp)));
            ln += fmaxf(sn, 0.0f) + log1pf(expf(-fabsf(sn)));
            hit += (sp > 0.0f ? 1.0f : 0.0f) + (sn < 0.0f ? 1.0f : 0.0f);
        }
    }
    if (lane == 0) { red[warp][0] = lp; red[warp][1] = ln; red[warp][2] = hit; }
    __syncthreads();
    if (threadIdx.x < 3) {
        float s = 0.0f;
        for (int w = 0; w < CL_WARPS; ++w) s += red[w][threadIdx.x];
        partial[blockIdx.x * 4 + threadIdx.x] = s;
    }
}

__global__ void ebm_finalize_kernel(const float* __restrict__ partial, int blocks, int64_t N, float* __restrict__ out /*[2]*/) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        float lp = 0.0f, ln = 0.0f, hit = 0.0f;
        for (int b = 0; b < blocks; ++b) { lp += partial[4 * b]; ln += partial[4 * b + 1]; hit += partial[4 * b + 2]; }
        out[0] = lp / static_cast<float>(N) + ln / static_cast<float>(N);  // CL_loss = loss_pos + loss_neg  (:61-63)
        out[1] = hit / static_cast<float>(2 * N);                         // CL_acc (:65-67)
    }
}

}  // namespace molsde

using namespace molsde;

extern "C" {

int molsde_schnet_cfconv(const molsde_plan* plan, const float* pos, const float* x, const float* w1t, const float* b1,
                         const float* w2t, const float* b2, const float* mu, int32_t num_gaussians, float coeff,
                         float cutoff, float* agg, void* stream) {
    if (!plan || !plan->tile_tgt_ptr || !plan->rowptr || !pos || !x || !w1t || !b1 || !w2t || !b2 || !mu || !agg)
        return MOLSDE_ERR_INVALID;
    if (num_gaussians <= 0 || num_gaussians > SN_G) return MOLSDE_ERR_UNSUPPORTED;
    if (plan->num_tiles == 0) return MOLSDE_OK;
    cudaError_t err = cudaFuncSetAttribute(schnet_cfconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SN_SMEM);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    const int grid = plan->num_tiles < kNumSMs ? plan->num_tiles : kNumSMs;
    schnet_cfconv_kernel<<<grid, SN_THREADS, SN_SMEM, as_stream(stream)>>>(*plan, pos, x, w1t, b1, w2t, b2, mu,
                                                                          num_gaussians, coeff, cutoff, agg);
    return check_launch("schnet_cfconv");
}

int molsde_gather_rows(const float* table, const int64_t* idx, int64_t rows, int32_t cols, float* out, void* stream) {
    if (!table || !idx || !out || rows < 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (rows == 0) return MOLSDE_OK;
    const int64_t total = rows * cols;
    gather_rows_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(table, idx, rows, cols, out);
    return check_launch("gather_rows");
}

int molsde_segment_reduce(const float* x, const int32_t* ptr, int32_t segments, int32_t cols, int32_t mean, float* out,
                          void* stream) {
    if (!x || !ptr || !out || segments < 0 || cols <= 0) return MOLSDE_ERR_INVALID;
    if (segments == 0) return MOLSDE_OK;
    const int64_t total = static_cast<int64_t>(segments) * cols;
    segment_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(x, ptr, segments, cols,
                                                                                                 mean, out);
    return check_launch("segment_reduce");
}

int molsde_ebm_node_dot(const float* X, const float* Y, const int64_t* perm, int64_t N, int32_t D, float T,
                        float* pred_pos, float* pred_neg, float* loss_acc, float* workspace, int64_t workspace_floats,
                        void* stream) {
    if (!X || !Y || !perm || !pred_pos || !pred_neg || !loss_acc || !workspace || N <= 0 || D <= 0 || T == 0.0f)
        return MOLSDE_ERR_INVALID;
    int blocks = static_cast<int>((N + CL_WARPS - 1) / CL_WARPS);
    if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
    if (workspace_floats < 4 * static_cast<int64_t>(blocks)) return MOLSDE_ERR_WORKSPACE;
    ebm_node_dot_kernel<<<blocks, CL_WARPS * 32, 0, as_stream(stream)>>>(X, Y, perm, N, D, 1.0f / T, pred_pos, pred_neg,
                                                                         workspace);
    int st = check_launch("ebm_node_dot");
    if (st != MOLSDE_OK) return st;
    ebm_finalize_kernel<<<1, 32, 0, as_stream(stream)>>>(workspace, blocks, N, loss_acc);
    return check_launch("ebm_finalize");
}

}  // extern "C"
