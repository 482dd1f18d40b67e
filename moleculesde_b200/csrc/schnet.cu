// SchNet forward pieces (Geom3D/models/schnet.py:85-125) and the EBM_node_dot_prod contrastive
// logits/loss (examples/util.py:52-68).
//
//   schnet_cfconv_kernel  -- one launch per InteractionBlock: for every radius-graph edge, in CSR-by-target
//       tiles of <=128 edges, compute in-kernel from the distance d (GaussianSmearing is never
//       materialised):   W = (Lin(51,128) -> ssp -> Lin(128,128))(exp(coeff (d - mu_k)^2)) * 0.5 (cos(d pi / rc) + 1)
//       (schnet.py:185-187,205-207), message x_j * W (:194-195) and the deterministic ascending-source
//       sum per target (:190).  Both filter GEMMs run on tcgen05 (kind::f16, two-way fp16 operand split, TMEM accumulators,
//       thread = edge slot); x = lin1(h) is gathered from L2.
//   gather_rows / segment_reduce -- embedding lookup (:89) and per-graph readout (:115).
//   ebm_node_dot_kernel   -- row-wise dots of X with Y and with Y[perm] (HBM-bound, not a GEMM),
//       BCE-with-logits partial sums reduced in a fixed order.
#include <math_constants.h>

#include "common.cuh"
#include "tc05.cuh"

namespace molsde {

// ---- CFConv on tcgen05 (round 2): two QUADS of 128 threads per CTA, a quad owns one edge tile at a time, thread = edge slot = TMEM
// lane.  Both filter GEMMs are tcgen05.mma kind::f16 (M = 128 edges, N = 128 filters, fp32 accumulator in the quad's TMEM columns)
// with the two-way fp16 operand split (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi: 22-bit operands; Gaussian features lie in [0,1], ssp
// outputs and trained weights are O(1)):
//   thread: (src, tgt) by bisection, d, C(d) -> 51 Gaussians -> fp16 hi/lo A row [128 x 64]      -> MMA 1 (4 K-steps x 3 terms)
//   epilogue 1: accumulator row (tcgen05.ld) + b1 -> ssp -> fp16 hi/lo A row [128 x 128]          -> MMA 2 (8 K-steps x 3 terms)
//   epilogue 2: (row + b2) * C * x[src] -> message tile (fp32, granules XOR-swizzled by the slot) -> per-target ascending-source sum
// Weights are converted ONCE per persistent CTA into the canonical K-major core-matrix B tiles.  Shared memory: B tiles 96 KB +
// one 64 KB union per quad (A1 | A2 | messages) = 225.5 KB.  The legacy kernel this replaces ran both GEMMs as warp-level
// mma.sync 3xTF32 (profiles/r2_cfconv_ncu.txt: tensor pipe 42 % busy, 8 warps, 1.58 ms per 1.1 M edges).
constexpr int SN_THREADS = 256, SN_QUADS = 2, SN_QT = 128;
constexpr int SN_F = 128;                 // num_filters (config.py:66)
constexpr int SN_G = 56;                  // num_gaussians (51) padded (layout of the host-packed w1t)
constexpr int SN_LDA = MOLSDE_TILE_LD;    // 136: row length of the host-packed k-major weights w1t / w2t
constexpr int SN_K1 = 64;                 // K of GEMM 1 (Gaussians padded to 4 K-steps of 16)
// byte offsets
constexpr int SB_W1H = 0, SB_W1L = SB_W1H + SN_F * SN_K1 * 2;          // [8 k-chunks][128 n][16 B] each
constexpr int SB_W2H = SB_W1L + SN_F * SN_K1 * 2, SB_W2L = SB_W2H + SN_F * SN_F * 2;   // [16 k-chunks][128 n][16 B] each
constexpr int SB_U = SB_W2L + SN_F * SN_F * 2;                          // per quad: 64 KB union
constexpr int SN_U_BYTES = 65536;
constexpr int SB_B1 = SB_U + SN_QUADS * SN_U_BYTES, SB_B2 = SB_B1 + SN_F * 4, SB_MU = SB_B2 + SN_F * 4;
constexpr int SB_BARS = SB_MU + 64 * 4, SB_TMEM = SB_BARS + SN_QUADS * 8;
constexpr size_t SN_SMEM = SB_TMEM + 16;
static_assert(SN_SMEM <= 232448, "smem");
static_assert(SB_U % 128 == 0 && SB_W2H % 128 == 0, "operand tile alignment");

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
                     float* __restrict__ agg /*[N][128]*/, int32_t* __restrict__ status) {
    extern __shared__ __align__(128) uint8_t sn_smem[];
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (SN_QT - 1), warp = tid >> 5;
    float* B1 = reinterpret_cast<float*>(sn_smem + SB_B1);
    float* B2 = reinterpret_cast<float*>(sn_smem + SB_B2);
    float* MU = reinterpret_cast<float*>(sn_smem + SB_MU);
    // ---- one-time: weights -> fp16 hi/lo B tiles (element (n, k) at (k/8) * 2048 + n * 16 + (k%8) * 2), biases, offsets, barriers, TMEM
    for (int item = tid; item < (SN_K1 / 8) * SN_F; item += SN_THREADS) {
        const int kc = item / SN_F, n = item % SN_F;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) { const int k = 8 * kc + j; v[j] = k < SN_G ? w1t[k * SN_LDA + n] : 0.0f; }
        tc05::store_chunk(sn_smem + SB_W1H, sn_smem + SB_W1L, n, kc, SN_F * 16, v);
    }
    for (int item = tid; item < (SN_F / 8) * SN_F; item += SN_THREADS) {
        const int kc = item / SN_F, n = item % SN_F;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = w2t[(8 * kc + j) * SN_LDA + n];
        tc05::store_chunk(sn_smem + SB_W2H, sn_smem + SB_W2L, n, kc, SN_F * 16, v);
    }
    if (tid < SN_F) { B1[tid] = b1[tid]; B2[tid] = b2[tid]; }
    if (tid < 64) MU[tid] = tid < SN_G ? mu[tid] : 0.0f;
    if (tid < SN_QUADS) tc05::mbar_init(tc05::smem_u32(sn_smem + SB_BARS + 8 * tid), 1);
    if (tid == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc05::smem_u32(sn_smem + SB_TMEM)), "r"(512u));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc05::fence_proxy_async_smem();
    tc05::fence_before();
    __syncthreads();
    tc05::fence_after();
    const uint32_t tmem_q = *reinterpret_cast<const uint32_t*>(sn_smem + SB_TMEM) + q * 256;   // 2 x 128 accumulator columns per quad
    const uint32_t tlane = tmem_q + (static_cast<uint32_t>(e & ~31) << 16);
    const uint32_t bar = tc05::smem_u32(sn_smem + SB_BARS + 8 * q);
    uint8_t* U = sn_smem + SB_U + q * SN_U_BYTES;
    const uint32_t u_addr = tc05::smem_u32(U), w1h = tc05::smem_u32(sn_smem + SB_W1H), w1l = tc05::smem_u32(sn_smem + SB_W1L);
    const uint32_t w2h = tc05::smem_u32(sn_smem + SB_W2H), w2l = tc05::smem_u32(sn_smem + SB_W2L);
    float* Mm = reinterpret_cast<float*>(U);
    const float pi_over_rc = 3.14159265358979323846f / cutoff;
    uint32_t ph = 0;
    bool ok = true;
    for (int tile = blockIdx.x * SN_QUADS + q; tile < plan.num_tiles; tile += gridDim.x * SN_QUADS) {
        const int ta = plan.tile_tgt_ptr[tile], tb = plan.tile_tgt_ptr[tile + 1];
        const int ea = plan.rowptr[ta], ne = plan.rowptr[tb] - ea;
        const bool live = e < ne;
        int sj = 0;
        float d = 0.0f, cc = 0.0f;
        if (live) {
            const int eg = ea + e;
            int lo = ta, hi = tb;
            while (hi - lo > 1) {
                const int mid = (lo + hi) >> 1;
                if (plan.rowptr[mid] <= eg) lo = mid; else hi = mid;
            }
            sj = plan.src[eg];
            // edge_weight = (pos[row] - pos[col]).norm(dim=-1)   (:93)
            const float dx = __fsub_rn(pos[3 * sj], pos[3 * lo]), dy = __fsub_rn(pos[3 * sj + 1], pos[3 * lo + 1]);
            const float dz = __fsub_rn(pos[3 * sj + 2], pos[3 * lo + 2]);
            d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            cc = 0.5f * (cosf(d * pi_over_rc) + 1.0f);  // :186
        }
        // ---- A1 row: GaussianSmearing (:205-207) of this edge, 64 columns (rows of dead slots are zero)
#pragma unroll
        for (int kc = 0; kc < SN_K1 / 8; ++kc) {
            float v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int k = 8 * kc + j;
                const float u = d - MU[k];
                v[j] = (live && k < num_gaussians) ? __expf(coeff * u * u) : 0.0f;
            }
            tc05::store_chunk(U, U + 16384, e, kc, 2048, v);
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before();
        tc05::group_sync(1 + q, SN_QT);
        if (e == 0) {
            tc05::fence_after();
            tc05::mma_split_f16<SN_F, SN_K1 / 16>(tmem_q, u_addr, u_addr + 16384, w1h, w1l, 0u);
            tc05::commit(bar);
        }
        ok &= tc05::mbar_wait(bar, ph);
        ph ^= 1u;
        tc05::fence_after();
        // ---- epilogue 1: hidden = ssp(acc + b1) -> A2 row [128 x 128] (overlays A1: its MMAs are complete)
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            float hv[32];
            tc05::tmem_ld32(tlane + 32 * cb, hv);
#pragma unroll
            for (int j = 0; j < 32; ++j) hv[j] = ssp_fast(hv[j] + B1[32 * cb + j]);
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) tc05::store_chunk(U, U + 32768, e, 4 * cb + k4, 2048, hv + 8 * k4);
        }
        tc05::fence_proxy_async_smem();
        tc05::fence_before();
        tc05::group_sync(1 + q, SN_QT);
        if (e == 0) {
            tc05::fence_after();
            tc05::mma_split_f16<SN_F, SN_F / 16>(tmem_q + 128, u_addr, u_addr + 32768, w2h, w2l, 0u);
            tc05::commit(bar);
        }
        ok &= tc05::mbar_wait(bar, ph);
        ph ^= 1u;
        tc05::fence_after();
        // ---- epilogue 2: message = x_j * ((acc + b2) * C)   (:187,194-195) -> message tile (overlays A2: its MMAs are complete)
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<size_t>(sj) * SN_F);
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            float wv[32];
            tc05::tmem_ld32(tlane + 128 + 32 * cb, wv);
            if (live) {
#pragma unroll
                for (int g4 = 0; g4 < 8; ++g4) {
                    const int g = 8 * cb + g4;
                    const float4 xv = __ldg(xr + g);
                    float4 m;
                    m.x = xv.x * ((wv[4 * g4] + B2[4 * g]) * cc);
                    m.y = xv.y * ((wv[4 * g4 + 1] + B2[4 * g + 1]) * cc);
                    m.z = xv.z * ((wv[4 * g4 + 2] + B2[4 * g + 2]) * cc);
                    m.w = xv.w * ((wv[4 * g4 + 3] + B2[4 * g + 3]) * cc);
                    *reinterpret_cast<float4*>(Mm + e * SN_F + ((g ^ (e & 31)) << 2)) = m;
                }
            }
        }
        tc05::fence_before();
        tc05::group_sync(1 + q, SN_QT);
        // ---- per-target ascending-source sum, aggr="add" (:190); thread = (target, column)
        const int ntg = tb - ta;
        for (int p = e; p < ntg * SN_F; p += SN_QT) {
            const int i = ta + (p >> 7), col = p & (SN_F - 1);
            const int s0 = plan.rowptr[i] - ea, s1 = plan.rowptr[i + 1] - ea;
            float sacc = 0.0f;
            for (int s = s0; s < s1; ++s) sacc += Mm[s * SN_F + ((((col >> 2) ^ (s & 31)) << 2) | (col & 3))];
            agg[static_cast<size_t>(i) * SN_F + col] = sacc;
        }
        tc05::fence_proxy_async_smem();   // the union is rewritten through the generic proxy, then read by the tensor core
        tc05::group_sync(1 + q, SN_QT);
    }
    if (!ok && status) *status = 1;
    tc05::fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*reinterpret_cast<const uint32_t*>(sn_smem + SB_TMEM)), "r"(512u));
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
    const int want = (plan->num_tiles + SN_QUADS - 1) / SN_QUADS;
    const int grid = want < kNumSMs ? want : kNumSMs;
    schnet_cfconv_kernel<<<grid, SN_THREADS, SN_SMEM, as_stream(stream)>>>(*plan, pos, x, w1t, b1, w2t, b2, mu,
                                                                          num_gaussians, coeff, cutoff, agg, nullptr);
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
