// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05, sm_100a):
//     C[M,N] (+)= A[M,K] . B[N,K]^T  (+ bias[n]) -> rowscale -> activation (+ residual)
// with arbitrary element strides for A and B (one of the two strides of each operand must be 1), so that the same
// kernel serves  y = x W^T  (forward),  dx = dy . W  and  dW = dy^T . x  (backward, split-K over the rows).
//
// Arithmetic: 3xTF32.  Every fp32 operand value v is split into hi = top 19 bits (exactly a tf32 number) and the exact
// remainder lo = v - hi; the accumulator receives lo_a*hi_b + hi_a*lo_b + hi_a*hi_b in fp32 (TMEM).  The dropped
// lo*lo term is < 2^-22 relative, i.e. the result is fp32-class (measured ~1e-6 max-norm relative vs fp64), which keeps
// the 1e-4 parity bar of scores, losses and gradients (BASELINE.json north_star).
//
// Structure (one CTA = one 128 x BN output tile, 256 threads):
//   * all 8 warps stage a 32-deep K block of A and B from global memory: load -> split hi/lo in registers -> store into
//     the canonical K-major no-swizzle core-matrix layout (8 rows x 16 B) that the UMMA shared-memory descriptors address;
//   * one thread issues the 12 tcgen05.mma (3 terms x 4 K-steps, M=128, N=BN, K=8) of the block and commits them to an
//     mbarrier; the two smem stages ping-pong, so staging of block k+1 overlaps the tensor-core work of block k;
//   * the fp32 accumulator lives in TMEM (BN columns x 128 lanes).  Its adds truncate, so every TC_FLUSH K blocks it is
//     drained with tcgen05.ld into fp32 registers (warp w owns lanes 32*(w%4).., warps 0-3 / 4-7 split the columns);
//     the epilogue applies bias / rowscale / activation / residual and stores, or writes a raw split-K partial.
#include "common.cuh"
#include <cuda.h>      // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <stdlib.h>

namespace molsde {

constexpr int TC_BM = 128, TC_BK = 32, TC_THREADS = 256;
constexpr int TC_FLUSH = 8;  // K blocks (of 32) accumulated in TMEM between two drains into registers
// Shared-memory operand layout (K-major, no swizzle): 8 rows x 16 B core matrices; consecutive K chunks of a row group are
// TC_LBO = 144 B apart (128 + 16 B of padding: the eight 16-byte stores of a quarter warp that walks along K then hit 32
// distinct banks), consecutive 8-row groups TC_SBO = 8 * 144 B apart.   byte(r, kc) = (r / 8) * TC_SBO + kc * TC_LBO + (r % 8) * 16
constexpr uint32_t TC_LBO = 144;
constexpr uint32_t TC_SBO = (TC_BK / 4) * TC_LBO;

__device__ __forceinline__ uint32_t tc_smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t tc_desc(uint32_t saddr) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(TC_LBO >> 4) << 16) |
           (static_cast<uint64_t>(TC_SBO >> 4) << 32) | (static_cast<uint64_t>(1) << 46);
}
template <int R>
__host__ __device__ constexpr int tc_tile_floats() { return (R / 8) * static_cast<int>(TC_SBO) / 4; }
__device__ __forceinline__ int tc_idx(int r, int kc) { return (r >> 3) * static_cast<int>(TC_SBO / 4) + kc * static_cast<int>(TC_LBO / 4) + (r & 7) * 4; }
template <int N>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    // instruction descriptor: D = F32, A = B = TF32, both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((static_cast<uint32_t>(N) >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ bool tc_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)  // bounded: a descriptor bug must not hang the GPU
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}

__device__ __forceinline__ float tc_act(float v, int act) {
    switch (act) {
        case 1: return fmaxf(v, 0.0f);
        case 2: return silu_f(v);
        case 3: return softplus_f(v) - 0.69314718246459961f;
        case 4: return tanhf(v);
        case 5: return v > 0.0f ? v : expm1f(v);
        default: return v;
    }
}

__device__ __noinline__ float tc_act_call(float v, int act) { return tc_act(v, act); }   // (keeps unrolled epilogues small)

// Staging of one [R x 32] K block of an operand in two phases so that the global-memory latency overlaps the tensor-core
// work of the previous block:  tc_load (all loads of the block issued back to back into registers)  ->  ...  ->
// tc_store (split hi/lo, store into the canonical core-matrix layout).   element (r, k) = src[r * sr + k * sk]; rows >=
// rvalid / k >= kvalid read as zero.  Work item = (row r, k-chunk kc of 4): lanes run along r, so the 16-byte smem stores of
// a warp are contiguous (conflict free) and, for r-contiguous operands, the global loads are coalesced; k-contiguous
// aligned operands use one float4 load per item.
template <int R>
struct TcRegs { float v[R * (TC_BK / 4) / TC_THREADS][4]; };

// item -> (row, k-chunk): lanes run along K for k-contiguous operands (a warp reads 4 rows x 128 contiguous bytes) and along
// the rows for row-contiguous operands (each of the 4 scalar loads of a warp is one contiguous 128-byte line)
template <int R>
__device__ __forceinline__ void tc_item(int item, bool kfast, int& r, int& kc) {
    if (kfast) { r = item / (TC_BK / 4); kc = item % (TC_BK / 4); } else { r = item % R; kc = item / R; }
}

template <int R>
__device__ __forceinline__ void tc_load(TcRegs<R>& g, const float* __restrict__ src, int64_t sr, int64_t sk, int rvalid, int kvalid,
                                        bool vec4, int ones_row = -1) {
    const bool kfast = sk == 1;
#pragma unroll
    for (int i = 0; i < R * (TC_BK / 4) / TC_THREADS; ++i) {
        int r, kc;
        tc_item<R>(threadIdx.x + i * TC_THREADS, kfast, r, kc);
        g.v[i][0] = g.v[i][1] = g.v[i][2] = g.v[i][3] = 0.0f;
        if (r == ones_row) {  // the extra all-ones B row: output column N = sum_k A(m,k) (bias gradient of a dW GEMM)
#pragma unroll
            for (int j = 0; j < 4; ++j) g.v[i][j] = (kc * 4 + j < kvalid) ? 1.0f : 0.0f;
            continue;
        }
        if (r < rvalid) {
            const float* p = src + r * sr + static_cast<int64_t>(kc) * 4 * sk;
            if (vec4 && kc * 4 + 3 < kvalid) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(p));
                g.v[i][0] = t.x; g.v[i][1] = t.y; g.v[i][2] = t.z; g.v[i][3] = t.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (kc * 4 + j < kvalid) g.v[i][j] = __ldg(p + j * sk);
            }
        }
    }
}
template <int R>
__device__ __forceinline__ void tc_store(const TcRegs<R>& g, float* __restrict__ hi, float* __restrict__ lo, bool kfast) {
#pragma unroll
    for (int i = 0; i < R * (TC_BK / 4) / TC_THREADS; ++i) {
        int r, kc;
        tc_item<R>(threadIdx.x + i * TC_THREADS, kfast, r, kc);
        float4 h4, l4;
        h4.x = __uint_as_float(__float_as_uint(g.v[i][0]) & 0xFFFFE000u); h4.y = __uint_as_float(__float_as_uint(g.v[i][1]) & 0xFFFFE000u);
        h4.z = __uint_as_float(__float_as_uint(g.v[i][2]) & 0xFFFFE000u); h4.w = __uint_as_float(__float_as_uint(g.v[i][3]) & 0xFFFFE000u);
        l4.x = g.v[i][0] - h4.x; l4.y = g.v[i][1] - h4.y; l4.z = g.v[i][2] - h4.z; l4.w = g.v[i][3] - h4.w;
        const int idx = tc_idx(r, kc);
        *reinterpret_cast<float4*>(hi + idx) = h4;
        *reinterpret_cast<float4*>(lo + idx) = l4;
    }
}

struct TcArgs {
    int64_t M, N, K;
    const float* A; int64_t sam, sak;
    const float* B; int64_t sbn, sbk;
    const float* bias; const float* rowscale; const float* R; int64_t ldr;
    float* C; int64_t ldc;
    int act, accumulate;
    int64_t k_per_split;   // multiple of TC_BK
    int splits, batch;     // grid.z = batch * splits
    int64_t bsA, bsB, bsC, bsBias, bsRow;  // element offsets between consecutive batches (A, B, C, bias, rowscale/R unused)
    float* ws;             // split-K partials [batch][splits][M][Nx] or NULL
    float* colsum;         // optional: colsum[m] (+)= sum_k A(m,k) through an all-ones B row at n == N
    int64_t Nx;            // N + (colsum ? 1 : 0): number of output columns incl. the ones column
    int vecA, vecB, vecC;
    int32_t* status;       // set to 1 if an mbarrier wait timed out (never expected)
};

template <int BN>
__global__ void __launch_bounds__(TC_THREADS, (BN <= 64 ? 2 : 1)) tc_gemm_kernel(const TcArgs a) {
    extern __shared__ __align__(128) float tc_smem[];
    constexpr int A_FLOATS = tc_tile_floats<TC_BM>(), B_FLOATS = tc_tile_floats<BN>();
    constexpr int STAGE = 2 * A_FLOATS + 2 * B_FLOATS;  // A hi | A lo | B hi | B lo
    __shared__ uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * TC_BM, n0 = static_cast<int64_t>(blockIdx.x) * BN;
    const int bz = static_cast<int>(blockIdx.z) / a.splits, sz = static_cast<int>(blockIdx.z) % a.splits;
    const int64_t kb0 = static_cast<int64_t>(sz) * a.k_per_split;
    const int64_t kend = min(a.K, kb0 + a.k_per_split);
    const int nkb = static_cast<int>((kend - kb0 + TC_BK - 1) / TC_BK);

    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[0])));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&bars[1])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    bool ok = true;
    constexpr int HALF = BN / 2;                 // warps 0-3 own columns [0, HALF), warps 4-7 own [HALF, BN) of their 32 lanes
    const int cbase = (warp >> 2) * HALF;
    float accr[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) accr[j] = 0.0f;

    const int rvalidA = static_cast<int>(min(static_cast<int64_t>(TC_BM), a.M - m0));
    const int rvalidB = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(BN), a.N - n0)));
    const int ones_row = (a.colsum && a.N >= n0 && a.N < n0 + BN) ? static_cast<int>(a.N - n0) : -1;
    TcRegs<TC_BM> ga;
    TcRegs<BN> gb;
    const float* Abase = a.A + bz * a.bsA + m0 * a.sam;
    const float* Bbase = a.B + bz * a.bsB + n0 * a.sbn;
    if (nkb > 0) {
        const int kv0 = static_cast<int>(min(static_cast<int64_t>(TC_BK), kend - kb0));
        tc_load<TC_BM>(ga, Abase + kb0 * a.sak, a.sam, a.sak, rvalidA, kv0, a.vecA != 0);
        tc_load<BN>(gb, Bbase + kb0 * a.sbk, a.sbn, a.sbk, rvalidB, kv0, a.vecB != 0, ones_row);
    }
    for (int kb = 0; kb < nkb; ++kb) {
        const int s = kb & 1;
        float* st = tc_smem + s * STAGE;
        if (kb >= 2) ok &= tc_wait(tc_smem_u32(&bars[s]), ((kb >> 1) - 1) & 1);  // tensor core finished reading stage s
        tc_store<TC_BM>(ga, st, st + A_FLOATS, a.sak == 1);
        tc_store<BN>(gb, st + 2 * A_FLOATS, st + 2 * A_FLOATS + B_FLOATS, a.sbk == 1);
        if (kb + 1 < nkb) {  // next block's global loads fly while this block's MMAs run
            const int64_t k1 = kb0 + static_cast<int64_t>(kb + 1) * TC_BK;
            const int kv1 = static_cast<int>(min(static_cast<int64_t>(TC_BK), kend - k1));
            tc_load<TC_BM>(ga, Abase + k1 * a.sak, a.sam, a.sak, rvalidA, kv1, a.vecA != 0);
            tc_load<BN>(gb, Bbase + k1 * a.sbk, a.sbn, a.sbk, rvalidB, kv1, a.vecB != 0, ones_row);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t ah = tc_smem_u32(st), al = ah + A_FLOATS * 4, bh = al + A_FLOATS * 4, bl = bh + B_FLOATS * 4;
            uint32_t acc = (kb % TC_FLUSH) > 0 ? 1u : 0u;
#pragma unroll 1
            for (int term = 0; term < 3; ++term) {
                const uint32_t pa = (term == 0) ? al : ah, pb = (term == 1) ? bl : bh;
#pragma unroll
                for (int ks = 0; ks < TC_BK / 8; ++ks) {
                    tc_mma<BN>(tmem, tc_desc(pa + ks * 2 * TC_LBO), tc_desc(pb + ks * 2 * TC_LBO), acc);
                    acc = 1u;
                }
            }
            asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&bars[s]))
                         : "memory");
        }
        // The TMEM accumulator adds with truncation (measured: relative error grows ~7e-9 per unit of K), so every
        // TC_FLUSH K blocks the partial sum is drained into fp32 registers (round-to-nearest adds) and TMEM restarts at 0.
        if ((kb + 1) % TC_FLUSH == 0 || kb == nkb - 1) {
            ok &= tc_wait(tc_smem_u32(&bars[s]), (kb >> 1) & 1);  // this commit covers every MMA issued so far
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
            for (int c0 = 0; c0 < HALF; c0 += 16) {
                uint32_t r[16];
                const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(cbase + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) accr[c0 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");  // ordered before the next block's MMAs by its __syncthreads
        }
    }
    if (!ok && a.status) *a.status = 1;

    // ---- epilogue: the thread of output row r holds HALF columns in registers -> bias / rowscale / activation there -> staged through
    // shared memory (the operand stages are idle: every MMA completed) -> written out by whole rows with the lanes along the columns
    // (coalesced; residual / accumulate / split-K partials / the ones column are handled on the way out).  The row-per-thread stores
    // this replaces cost ~30 % of the kernel on the TMA-fed variant.
    {
        constexpr int LDS = BN + 4;
        float* stage = tc_smem;                      // [128][LDS] floats (<= 66 KB of the staging ring)
        const int r = 32 * (warp & 3) + lane;
        const int64_t gm = m0 + r;
        float* part = a.ws ? a.ws + static_cast<size_t>(blockIdx.z) * a.M * a.Nx : nullptr;   // [batch][split][M][Nx]
        float* Cb = a.C + bz * a.bsC;
        const float* biasb = (a.bias && !part) ? a.bias + bz * a.bsBias : nullptr;
        const float rs = (a.rowscale && !part && gm < a.M) ? a.rowscale[gm] : 1.0f;
        const bool plain = part != nullptr;          // split-K partial: raw accumulator
        const bool slow_act = !plain && a.act >= 2;
        __syncthreads();                             // (all warps are past their last read of the operand stages)
#pragma unroll
        for (int j = 0; j < HALF; j += 4) {
            float4 o;
            float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t gn = n0 + cbase + j + q;
                float v = accr[j + q];
                if (!plain && gn < a.N) {            // (the ones column n == N stays raw)
                    if (biasb) v += biasb[gn];
                    if (a.rowscale) v *= rs;
                    v = slow_act ? tc_act_call(v, a.act) : (a.act == 1 ? fmaxf(v, 0.0f) : v);
                }
                ov[q] = v;
            }
            *reinterpret_cast<float4*>(stage + r * LDS + cbase + j) = o;
        }
        __syncthreads();
        const int ncols = static_cast<int>(max(static_cast<int64_t>(0), min(static_cast<int64_t>(BN), a.Nx - n0)));
        for (int rr = warp; rr < TC_BM; rr += TC_THREADS / 32) {
            const int64_t row = m0 + rr;
            if (row >= a.M) break;
            const float* srow = stage + rr * LDS;
            if (part) {
                float* pr = part + row * a.Nx + n0;
                for (int col = lane; col < ncols; col += 32) pr[col] = srow[col];
                continue;
            }
            float* c = Cb + row * a.ldc + n0;
            const float* res = a.R ? a.R + row * a.ldr + n0 : nullptr;
            const bool vec = a.vecC && !res && !a.accumulate && !a.colsum;
            for (int c4 = lane; 4 * c4 < ncols; c4 += 32) {
                if (vec && 4 * c4 + 3 < ncols) {
                    *reinterpret_cast<float4*>(c + 4 * c4) = *reinterpret_cast<const float4*>(srow + 4 * c4);
                    continue;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 4 * c4 + q;
                    if (col >= ncols) break;
                    float v = srow[col];
                    if (n0 + col == a.N) { a.colsum[row] = a.accumulate ? a.colsum[row] + v : v; continue; }   // the ones column
                    if (res) v += res[col];
                    c[col] = a.accumulate ? c[col] + v : v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN));
}

// C[b] (+)= sum over the split-K partials ws[b][s] in ascending s.  CTA = 32 outputs x 8 split lanes: lane l adds the
// partials s = l, l+8, ... (fixed order), then the 8 lane sums are added in lane order -> deterministic, and the long
// dependent chain of the tall-skinny weight gradients (hundreds of splits, a few hundred outputs) is 8x shorter.
__global__ void __launch_bounds__(256)
tc_splitk_reduce_kernel(const float* __restrict__ ws, int splits, int batch, int64_t M, int64_t N, int64_t Nreal, float* __restrict__ C,
                        int64_t ldc, int64_t bsC, int accumulate, float* __restrict__ colsum) {
    __shared__ float red[8][33];
    const int tx = threadIdx.x & 31, sl = threadIdx.x >> 5;
    const int64_t idx = static_cast<int64_t>(blockIdx.x) * 32 + tx;
    const int64_t MN = M * N;
    float v = 0.0f;
    if (idx < batch * MN) {
        const int64_t b = idx / MN, e = idx % MN;
        const float* p = ws + static_cast<size_t>(b) * splits * MN + e;
        for (int s = sl; s < splits; s += 8) v += p[static_cast<size_t>(s) * MN];
    }
    red[sl][tx] = v;
    __syncthreads();
    if (sl == 0 && idx < batch * MN) {
        float t = 0.0f;
#pragma unroll
        for (int q = 0; q < 8; ++q) t += red[q][tx];
        const int64_t b = idx / MN, e = idx % MN;
        float* c = (e % N < Nreal) ? C + b * bsC + (e / N) * ldc + e % N : colsum + e / N;   // column Nreal = the ones column
        *c = accumulate ? *c + t : t;
    }
}

// few splits: one thread per output
__global__ void tc_splitk_reduce_simple_kernel(const float* __restrict__ ws, int splits, int batch, int64_t M, int64_t N, int64_t Nreal,
                                               float* __restrict__ C, int64_t ldc, int64_t bsC, int accumulate, float* __restrict__ colsum) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    const int64_t MN = M * N;
    if (idx >= batch * MN) return;
    const int64_t b = idx / MN, e = idx % MN;
    const float* p = ws + static_cast<size_t>(b) * splits * MN + e;
    float t = 0.0f;
    for (int s = 0; s < splits; ++s) t += p[static_cast<size_t>(s) * MN];
    float* c = (e % N < Nreal) ? C + b * bsC + (e / N) * ldc + e % N : colsum + e / N;
    *c = accumulate ? *c + t : t;
}

// =======================================================================================
// TMA-fed variant (plain y = x W^T shapes: both operands K-contiguous, 16-byte aligned rows, no split-K / batch / ones row).
//   * raw fp32 operand tiles [128 x 32] / [BN x 32] arrive by TMA (cp.async.bulk.tensor.2d, tensor maps encoded on the host
//     with SWIZZLE_128B, out-of-bounds rows / K tail zero-filled by the hardware) into a 3-stage ring, completion on mbarriers;
//   * 256 converter threads derive the lo tile IN THE SWIZZLED LAYOUT: the split is elementwise, so a thread reads a float4 at
//     byte offset o of the raw tile and writes lo = v - trunc19(v) at offset o of the lo tile -- no address math, no layout
//     transform, conflict-free 16-byte accesses; the raw tile IS the hi operand (kind::tf32 ignores the 13 low mantissa bits);
//   * one thread issues the 12 tcgen05.mma (M = 128, N = BN = 128: the full-rate tf32 shape) per K block against
//     SWIZZLE_128B K-major descriptors (8-row groups 1024 B apart, K step = +32 B on the start address) and commits to the
//     stage's "empty" mbarrier, which gates the TMA refill of that stage; accumulator drains as in the kernel above.
// =======================================================================================
constexpr int TT_STAGES = 3;
constexpr uint32_t TT_A_BYTES = TC_BM * 128;   // [128 rows][32 floats]

__device__ __forceinline__ uint64_t tt_desc(uint32_t saddr) {   // K-major, SWIZZLE_128B: LBO unused (1), SBO = 1024 B
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4) | (static_cast<uint64_t>(1) << 16) | (static_cast<uint64_t>(1024 >> 4) << 32) |
           (static_cast<uint64_t>(1) << 46) | (static_cast<uint64_t>(2) << 61);
}
__device__ __forceinline__ void tt_tma_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar) : "memory");
}
// raw tile -> lo tile (`n16` 16-byte chunks, all 256 converter threads).  The raw tile itself serves as the hi operand: kind::tf32
// reads the top 19 bits of every fp32 container and ignores the 13 low mantissa bits, i.e. the tensor core sees exactly
// hi = v & 0xFFFFE000 (checked against fp64 to 5e-6 in tests/test_gpu_tcgemm.py: a rounding read would show up as ~2e-4).
__device__ __forceinline__ void tt_split(const uint8_t* raw, uint8_t* lo, int n16) {
    for (int i = threadIdx.x; i < n16; i += TC_THREADS) {
        const float4 v = *reinterpret_cast<const float4*>(raw + i * 16);
        float4 l;
        l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
        l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
        *reinterpret_cast<float4*>(lo + i * 16) = l;
    }
}

constexpr int TT_THREADS = TC_THREADS + 32;   // 8 converter / epilogue warps + 1 control warp (TMA producer + MMA issuer)

template <int BN>
__global__ void __launch_bounds__(TT_THREADS, 1)
tc_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcArgs a) {
    extern __shared__ uint8_t tt_smem_raw[];
    constexpr uint32_t B_BYTES = BN * 128, STAGE = 2 * (TT_A_BYTES + B_BYTES);
    __shared__ uint64_t full[TT_STAGES], ready[TT_STAGES], empty[TT_STAGES];   // TMA landed | hi/lo written (256 arrivals) | MMAs done
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t* tiles = tt_smem_raw + ((1024u - (tc_smem_u32(tt_smem_raw) & 1023u)) & 1023u);   // SWIZZLE_128B atoms: 1024-byte aligned
    const int64_t m0 = static_cast<int64_t>(blockIdx.y) * TC_BM, n0 = static_cast<int64_t>(blockIdx.x) * BN;
    const int nkb = static_cast<int>((a.K + TC_BK - 1) / TC_BK);
    if (tid == 0) {
        for (int s = 0; s < TT_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(&ready[s])), "r"(TC_THREADS));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tc_smem_u32(&empty[s])));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tc_smem_u32(&tmem_slot)), "r"(BN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = tmem_slot;
    bool ok = true;
    constexpr int HALF = BN / 2;
    const int cbase = ((warp >> 2) & 1) * HALF;
    float accr[HALF];
#pragma unroll
    for (int j = 0; j < HALF; ++j) accr[j] = 0.0f;

    if (warp == TC_THREADS / 32) {
        // ---------------- control warp: one thread feeds the ring by TMA and issues the tensor-core work ----------------
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmA)) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmB)) : "memory");
            auto issue_tma = [&](int kb) {
                const int s = kb % TT_STAGES;
                const uint32_t st = tc_smem_u32(tiles) + s * STAGE, bar = tc_smem_u32(&full[s]);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(TT_A_BYTES + B_BYTES) : "memory");
                tt_tma_2d(st, &tmA, kb * TC_BK, static_cast<int>(m0), bar);
                tt_tma_2d(st + 2 * TT_A_BYTES, &tmB, kb * TC_BK, static_cast<int>(n0), bar);
            };
            for (int kb = 0; kb < TT_STAGES && kb < nkb; ++kb) issue_tma(kb);
            for (int kb = 0; kb < nkb; ++kb) {
                const int s = kb % TT_STAGES;
                const uint32_t ph = static_cast<uint32_t>(kb / TT_STAGES) & 1u;
                ok &= tc_wait(tc_smem_u32(&ready[s]), ph);   // hi / lo tiles of block kb written by all converter threads
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t ah = tc_smem_u32(tiles) + s * STAGE, al = ah + TT_A_BYTES, bh = al + TT_A_BYTES, bl = bh + B_BYTES;
                uint32_t acc = (kb % TC_FLUSH) > 0 ? 1u : 0u;
#pragma unroll 1
                for (int term = 0; term < 3; ++term) {
                    const uint32_t pa = (term == 0) ? al : ah, pb = (term == 1) ? bl : bh;
#pragma unroll
                    for (int ks = 0; ks < TC_BK / 8; ++ks) {
                        tc_mma<BN>(tmem, tt_desc(pa + ks * 32), tt_desc(pb + ks * 32), acc);
                        acc = 1u;
                    }
                }
                asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(&empty[s]))
                             : "memory");
                // refill the stage of the previous block (its MMAs were issued one block ago) with block kb - 1 + STAGES
                if (kb >= 1 && kb - 1 + TT_STAGES < nkb) {
                    ok &= tc_wait(tc_smem_u32(&empty[(kb - 1) % TT_STAGES]), static_cast<uint32_t>((kb - 1) / TT_STAGES) & 1u);
                    issue_tma(kb - 1 + TT_STAGES);
                }
            }
        }
    } else {
        // ---------------- converter warps: raw tile -> hi (in place) + lo, then arrive; drain TMEM at the flush points ----------------
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % TT_STAGES;
            const uint32_t ph = static_cast<uint32_t>(kb / TT_STAGES) & 1u;
            uint8_t* st = tiles + s * STAGE;
            ok &= tc_wait(tc_smem_u32(&full[s]), ph);
            tt_split(st, st + TT_A_BYTES, TT_A_BYTES / 16);
            tt_split(st + 2 * TT_A_BYTES, st + 2 * TT_A_BYTES + B_BYTES, B_BYTES / 16);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to the tensor core
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");   // (orders an earlier TMEM drain before the next MMAs)
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(&ready[s])) : "memory");
            if ((kb + 1) % TC_FLUSH == 0 || kb == nkb - 1) {
                ok &= tc_wait(tc_smem_u32(&empty[s]), ph);   // this commit covers every MMA issued so far
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
                for (int c0 = 0; c0 < HALF; c0 += 16) {
                    uint32_t r[16];
                    const uint32_t taddr = tmem + (static_cast<uint32_t>(32 * (warp & 3)) << 16) + static_cast<uint32_t>(cbase + c0);
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
                        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                        : "r"(taddr));
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) accr[c0 + j] += __uint_as_float(r[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
        }
    }
    if (!ok && a.status) *a.status = 1;
    // ---- epilogue: the thread of output row r holds HALF columns in registers -> bias / rowscale / activation there -> staged
    // through shared memory (the operand ring is idle: every TMA landed, every MMA completed) -> written out by whole rows, one
    // coalesced 512-byte store per warp and row (residual / accumulate are read the same way).
    if (warp < TC_THREADS / 32) {
        constexpr int LDS = BN + 4;                          // 528-byte rows: the 16-byte row-segment stores of a quarter warp hit 8 bank groups
        float* stage = reinterpret_cast<float*>(tiles);      // [128][LDS] floats = 66 KB of the 192 KB ring
        const int r = 32 * (warp & 3) + lane;
        const int64_t gm = m0 + r;
        const float rs = (a.rowscale && gm < a.M) ? a.rowscale[gm] : 1.0f;
        const bool slow_act = a.act >= 2;
#pragma unroll
        for (int j = 0; j < HALF; j += 4) {
            float4 o;
            float* ov = reinterpret_cast<float*>(&o);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int64_t gn = n0 + cbase + j + q;
                float v = accr[j + q];
                if (a.bias && gn < a.N) v += a.bias[gn];
                if (a.rowscale) v *= rs;
                ov[q] = slow_act ? tc_act_call(v, a.act) : (a.act == 1 ? fmaxf(v, 0.0f) : v);
            }
            *reinterpret_cast<float4*>(stage + r * LDS + cbase + j) = o;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(TC_THREADS) : "memory");
        const int ncols = static_cast<int>(min(static_cast<int64_t>(BN), a.N - n0));
        for (int rr = warp; rr < TC_BM; rr += TC_THREADS / 32) {
            const int64_t row = m0 + rr;
            if (row >= a.M) break;
            float* c = a.C + row * a.ldc + n0;
            const float* res = a.R ? a.R + row * a.ldr + n0 : nullptr;
            if (a.vecC && !a.R && !a.accumulate && 4 * lane + 3 < ncols) {
                *reinterpret_cast<float4*>(c + 4 * lane) = *reinterpret_cast<const float4*>(stage + rr * LDS + 4 * lane);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const int col = 32 * q + lane;           // lanes along the columns: coalesced scalar accesses
                    if (col >= ncols) continue;
                    float v = stage[rr * LDS + col];
                    if (res) v += res[col];
                    c[col] = a.accumulate ? c[col] + v : v;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(BN));
}

typedef CUresult (*tt_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                 CUtensorMapFloatOOBfill);
static tt_encode_fn tt_encoder() {
    static tt_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<tt_encode_fn>(p);
    }
    return fn;
}
// [rows x K] fp32, K contiguous, row stride ld floats -> tensor map with a [box_rows x 32] box, 128-byte swizzle, zero fill
static bool tt_make_map(CUtensorMap* tm, const float* base, int64_t rows, int64_t K, int64_t ld, int box_rows) {
    tt_encode_fn enc = tt_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(K), static_cast<cuuint64_t>(rows)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
    const cuuint32_t box[2] = {TC_BK, static_cast<cuuint32_t>(box_rows)};
    const cuuint32_t estr[2] = {1, 1};
    return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool tt_eligible(const TcArgs& a, int splits) {
    static const bool off = getenv("MOLSDE_TC_NO_TMA") != nullptr;
    return !off && splits == 1 && a.batch == 1 && !a.colsum && a.sak == 1 && a.sbk == 1 && a.sam % 4 == 0 && a.sbn % 4 == 0 &&
           (reinterpret_cast<uintptr_t>(a.A) & 15) == 0 && (reinterpret_cast<uintptr_t>(a.B) & 15) == 0 && a.N > 64 && a.M >= 256 &&
           a.K >= 64 && a.M < (1ll << 31) && a.N < (1ll << 31) && a.K < (1ll << 31);
}
static int tt_launch(const TcArgs& a, cudaStream_t s) {
    constexpr int BN = 128;
    constexpr size_t smem = TT_STAGES * 2 * (TT_A_BYTES + BN * 128) + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_tma_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MOLSDE_ERR_CUDA; }
        configured = true;
    }
    CUtensorMap tmA, tmB;
    if (!tt_make_map(&tmA, a.A, a.M, a.K, a.sam, TC_BM) || !tt_make_map(&tmB, a.B, a.N, a.K, a.sbn, BN)) return -1000;  // caller falls back
    dim3 grid(static_cast<unsigned>((a.N + BN - 1) / BN), static_cast<unsigned>((a.M + TC_BM - 1) / TC_BM), 1);
    tc_gemm_tma_kernel<BN><<<grid, TT_THREADS, smem, s>>>(tmA, tmB, a);
    return check_launch("tc_gemm_tma");
}

template <int BN>
static int tc_launch(const TcArgs& a, int splits, cudaStream_t s) {
    constexpr size_t smem = 2 * (2 * tc_tile_floats<TC_BM>() + 2 * tc_tile_floats<BN>()) * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(tc_gemm_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) { set_last_error(cudaGetErrorString(e)); return MOLSDE_ERR_CUDA; }
        configured = true;
    }
    dim3 grid(static_cast<unsigned>((a.Nx + BN - 1) / BN), static_cast<unsigned>((a.M + TC_BM - 1) / TC_BM), splits * a.batch);
    tc_gemm_kernel<BN><<<grid, TC_THREADS, smem, s>>>(a);
    return check_launch("tc_gemm");
}

}  // namespace molsde

using namespace molsde;

static int tc_bn(int64_t N) {
    // BN = 64 keeps a CTA at 96 KB of staging (2 CTAs per SM: one CTA's epilogue / prologue overlaps the other's main loop);
    // MOLSDE_TC_BN=128 selects the wide tile (1 CTA per SM) for experiments
    static const int forced = getenv("MOLSDE_TC_BN") ? atoi(getenv("MOLSDE_TC_BN")) : 0;
    if (N <= 32) return 32;
    if (forced == 128 && N > 64) return 128;
    return 64;
}

static int tc_splits(int64_t M, int64_t N, int64_t K, int batch) {
    const int bn = tc_bn(N);
    const int64_t tiles = ((M + TC_BM - 1) / TC_BM) * ((N + bn - 1) / bn) * batch;
    // Split-K policy, tuned on the whole pretraining step (graph replay, batch 256), where these GEMMs never run alone -- two or three
    // streams keep the SMs busy, so a GEMM does not have to fill the GPU by itself and every split costs partial-tile writes plus a
    // reduce launch:  no split once the output has >= 100 tiles (2/3 of a wave);  otherwise ~0.8 waves of CTAs (was two);  >= 8 K
    // blocks per split (was 4).  Step 8.81 -> 8.20 ms (one wave) -> 7.94 ms (0.8 waves, with two side streams for the weight-gradient
    // leaves); the plateau is flat: 0.5-0.8 waves and cut-offs of 60-100 tiles all land within 7.90-8.02 ms.
    // MOLSDE_TC_NOSPLIT_TILES / MOLSDE_TC_SPLIT_WAVES / MOLSDE_TC_SPLIT_MIN_KB re-open the A/B.
    static const int64_t no_split_tiles = getenv("MOLSDE_TC_NOSPLIT_TILES") ? atoll(getenv("MOLSDE_TC_NOSPLIT_TILES")) : 100;
    if (tiles >= no_split_tiles) return 1;
    static const double waves = getenv("MOLSDE_TC_SPLIT_WAVES") ? atof(getenv("MOLSDE_TC_SPLIT_WAVES")) : 0.8;
    int64_t s = (static_cast<int64_t>(waves * kNumSMs) + tiles - 1) / tiles;
    static const int64_t min_kb = getenv("MOLSDE_TC_SPLIT_MIN_KB") ? atoll(getenv("MOLSDE_TC_SPLIT_MIN_KB")) : 8;
    const int64_t maxs = (K + min_kb * TC_BK - 1) / (min_kb * TC_BK);  // >= min_kb K blocks per split
    if (s > maxs) s = maxs;
    if (s > 512) s = 512;
    return s < 1 ? 1 : static_cast<int>(s);
}

static int tc_run(int32_t batch, int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, int64_t bsA, const float* B,
                  int64_t sbn, int64_t sbk, int64_t bsB, const float* bias, int64_t bsBias, int32_t act, const float* rowscale,
                  const float* R, int64_t ldr, float* C, int64_t ldc, int64_t bsC, int32_t accumulate, float* ws, int64_t ws_floats,
                  int32_t* status, void* stream, float* colsum = nullptr) {
    if (colsum && (batch != 1 || bias || rowscale || R || act != 0)) return MOLSDE_ERR_UNSUPPORTED;
    if (!A || !B || !C || M < 0 || N < 0 || K < 0 || batch < 1 || (sam != 1 && sak != 1) || (sbn != 1 && sbk != 1)) return MOLSDE_ERR_INVALID;
    if (batch > 1 && (rowscale || R)) return MOLSDE_ERR_UNSUPPORTED;
    if (M == 0 || N == 0) return MOLSDE_OK;
    TcArgs a;
    a.M = M; a.N = N; a.K = K; a.A = A; a.sam = sam; a.sak = sak; a.B = B; a.sbn = sbn; a.sbk = sbk;
    a.bias = bias; a.rowscale = rowscale; a.R = R; a.ldr = ldr; a.C = C; a.ldc = ldc; a.act = act; a.accumulate = accumulate;
    a.colsum = colsum; a.Nx = N + (colsum ? 1 : 0);
    a.status = status; a.batch = batch; a.bsA = bsA; a.bsB = bsB; a.bsC = bsC; a.bsBias = bsBias; a.bsRow = 0;
    const bool plain = !bias && !rowscale && !R && act == 0;
    int splits = plain ? tc_splits(M, a.Nx, K, batch) : 1;
    if (splits > 1 && (!ws || ws_floats < static_cast<int64_t>(splits) * batch * M * a.Nx)) splits = 1;
    int64_t kps = (K + splits - 1) / splits;
    kps = (kps + TC_BK - 1) / TC_BK * TC_BK;
    if (kps < TC_BK) kps = TC_BK;
    splits = static_cast<int>((K + kps - 1) / kps);
    if (splits < 1) splits = 1;
    a.k_per_split = kps;
    a.splits = splits;
    a.ws = splits > 1 ? ws : nullptr;
    // float4 global loads need k-contiguity and 16-byte aligned base, row stride and batch stride
    a.vecA = (sak == 1 && (reinterpret_cast<uintptr_t>(A) & 15) == 0 && sam % 4 == 0 && bsA % 4 == 0) ? 1 : 0;
    a.vecB = (sbk == 1 && (reinterpret_cast<uintptr_t>(B) & 15) == 0 && sbn % 4 == 0 && bsB % 4 == 0) ? 1 : 0;
    a.vecC = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0 && bsC % 4 == 0) ? 1 : 0;
    cudaStream_t s = as_stream(stream);
    if (tt_eligible(a, splits)) {
        const int st_tma = tt_launch(a, s);
        if (st_tma != -1000) return st_tma;
    }
    const int bn = tc_bn(a.Nx);
    int st = bn == 32 ? tc_launch<32>(a, splits, s) : bn == 64 ? tc_launch<64>(a, splits, s) : tc_launch<128>(a, splits, s);
    if (st != MOLSDE_OK || splits == 1) return st;
    if (splits <= 16)
        tc_splitk_reduce_simple_kernel<<<static_cast<unsigned>((batch * M * a.Nx + 255) / 256), 256, 0, s>>>(ws, splits, batch, M, a.Nx, N, C,
                                                                                                         ldc, bsC, accumulate, colsum);
    else
        tc_splitk_reduce_kernel<<<static_cast<unsigned>((batch * M * a.Nx + 31) / 32), 256, 0, s>>>(ws, splits, batch, M, a.Nx, N, C, ldc,
                                                                                                bsC, accumulate, colsum);
    return check_launch("tc_gemm.splitk_reduce");
}

extern "C" {

int64_t molsde_tc_gemm_ws_floats(int64_t M, int64_t N, int64_t K) {
    const int s = tc_splits(M, N, K, 1);
    return s > 1 ? s * M * N : 0;
}
int64_t molsde_tc_gemm_batched_ws_floats(int32_t batch, int64_t M, int64_t N, int64_t K) {
    const int s = tc_splits(M, N, K, batch);
    return s > 1 ? static_cast<int64_t>(s) * batch * M * N : 0;
}

/* element A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]; for each operand one of its two strides must be 1.
 * split-K (only when ws is given, bias/act/rowscale/R are all unset and K is large relative to the tile count). */
int molsde_tc_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                   const float* bias, int32_t act, const float* rowscale, const float* R, int64_t ldr, float* C, int64_t ldc,
                   int32_t accumulate, float* ws, int64_t ws_floats, int32_t* status, void* stream) {
    return tc_run(1, M, N, K, A, sam, sak, 0, B, sbn, sbk, 0, bias, 0, act, rowscale, R, ldr, C, ldc, 0, accumulate, ws, ws_floats, status,
                  stream);
}

/* Weight + bias gradient of nn.Linear in ONE GEMM: dW[M=out, N=in] (+)= dy^T . x and db[m] (+)= sum_rows dy[row, m], the
 * latter through an all-ones extra operand row (workspace: molsde_tc_gemm_ws_floats(M, N + 1, K)). */
int molsde_tc_gemm_dw_db(int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn,
                         int64_t sbk, float* dW, int64_t ldc, float* db, int32_t accumulate, float* ws, int64_t ws_floats,
                         int32_t* status, void* stream) {
    if (!db) return MOLSDE_ERR_INVALID;
    return tc_run(1, M, N, K, A, sam, sak, 0, B, sbn, sbk, 0, nullptr, 0, 0, nullptr, nullptr, 0, dW, ldc, 0, accumulate, ws, ws_floats,
                  status, stream, db);
}

/* `batch` independent GEMMs of the same shape in one launch: operand / output / bias pointers of batch b are offset by
 * b*bsA, b*bsB, b*bsC, b*bsBias elements (0 = shared).  The per-channel layers of EdgeNetwork_dense (edge_network_dense.py:
 * 105-128) and their backward are instances. */
int molsde_tc_gemm_batched(int32_t batch, int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, int64_t bsA,
                           const float* B, int64_t sbn, int64_t sbk, int64_t bsB, const float* bias, int64_t bsBias, int32_t act,
                           float* C, int64_t ldc, int64_t bsC, int32_t accumulate, float* ws, int64_t ws_floats, int32_t* status,
                           void* stream) {
    return tc_run(batch, M, N, K, A, sam, sak, bsA, B, sbn, sbk, bsB, bias, bsBias, act, nullptr, nullptr, 0, C, ldc, bsC, accumulate, ws,
                  ws_floats, status, stream);
}

}  // extern "C"
