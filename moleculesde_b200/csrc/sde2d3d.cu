// SDEModel2Dto3D_02 score network (K3) and the fused predictor-corrector reverse-SDE loop (K5).
//
// Reference path: Geom3D/models/MoleculeSDE/SDE_model_2D_to_3D.py:393-445 (get_score),
// equivariant_scorenetwork.py:121-169, and the sampler in
// examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-212.
//
// Design (B200-first):
//   * one persistent CTA (512 threads = 2 decoupled groups of 8 warps, 1 CTA/SM, ~219 KB smem) owns one "chunk" = a set of whole
//     molecules with <= 224 atoms; node state (hidden features, q/k/v, positions, score) lives in
//     shared memory for the WHOLE score evaluation -- and, in the PC kernel, for all 1000 reverse
//     steps -- so HBM sees only the initial/final positions;
//   * edges are processed in CSR-by-target order in tiles of <= 128 edges aligned to target nodes,
//     so the segment softmax / mean aggregation of a tile is self-contained and runs in a fixed,
//     atomic-free, ascending-source order (deterministic, same order as the reference scatter);
//   * every per-edge / per-node MLP is a tile GEMM on the tensor cores: mma.sync m16n8k16 (f16 inputs, fp32 accumulate)
//     with an error-compensated two-way fp16 split of both operands (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi; weights
//     pre-split on the host), which keeps fp32-grade accuracy (north_star: 1e-4) at ~6x the math rate the FFMA
//     register tile reaches from shared memory (profiles/r1_ubench_mma_rate.txt); the basis MLP runs on tcgen05
//     (3xTF32, TMEM accumulator).  A operands are k-major in smem with a padded leading dimension (== 8 mod 32);
//   * the per-edge attribute (32 floats) is the only per-edge state that survives between phases; it
//     goes to an L2-resident per-CTA scratch in the smem tile layout [32][136], so re-loading it is
//     a straight 17 KB cp.async copy.
#include <math_constants.h>

#include "common.cuh"
#include "mma_tile.cuh"
#include "sde2d3d_params.h"

// tile GEMMs of the score network: mma.m16n8k16 (f16, fp32 accumulate) with the two-way fp16 split of both operands
// (mma_tile.cuh); every K here is 32 and the weight blocks of the parameter blob arrive pre-split from the host (pack_f16_pairs)
#define MOLSDE_MMA_GEMM mma_gemm_hp

namespace molsde {

constexpr int TE = MOLSDE_TILE_EDGES;         // 128 edges per tile
constexpr int NTHREADS = 512;
constexpr int NWARPS = NTHREADS / 32;
constexpr int GROUPS = 2;                     // two 8-warp groups work on alternate tiles, decoupled
constexpr int GTHREADS = NTHREADS / GROUPS;
constexpr int MAXN = MOLSDE_CHUNK_MAX_NODES;  // 224 atoms per chunk
constexpr int MAXT = 64;                      // tiles per chunk
constexpr int LDA = MOLSDE_TILE_LD;           // 136: leading dim of a k-major edge tile
constexpr int LDX = 232;                      // leading dim of the k-major node matrix (>= MAXN, == 8 mod 32)
constexpr int TILE_FLOATS = 32 * LDA;         // one [32][136] per-edge attribute tile
constexpr int FRAME_FLOATS = 9 * TE;          // per-edge SE(3) frame (diff, cross, vertical) cached by E0 for the basis phases
constexpr int SCR_TILE = TILE_FLOATS + FRAME_FLOATS;  // per-tile scratch record: edge_attr [32][136] | frame [9][128]
constexpr int LDM = 33;                       // padded row of the slot-major message tile [TE][33]
constexpr int LD32 = MOLSDE_LD32, LD96 = MOLSDE_LD96;
constexpr float EPS = 1e-6f;                  // SDE_model_2D_to_3D.py:10
constexpr float LN_EPS = 1e-5f;

// ---- shared memory carve-up (float offsets) ----
constexpr int S_XT = 0;                      // [32][LDX]   node hidden, k-major
constexpr int S_Q = S_XT + 32 * LDX;         // [MAXN][32]  query  (aggregate written in place)
constexpr int S_K = S_Q + 32 * MAXN;         // [MAXN][32]
constexpr int S_V = S_K + 32 * MAXN;         // [MAXN][32]
constexpr int S_WG = S_V + 32 * MAXN;        // [P_GAT_SZ]  weights of the current GAT layer
constexpr int S_A = S_WG + MOLSDE_P_GAT_SZ;  // [GROUPS][32][LDA] A operand per group (k-major); also the message tile
                                             //                   [TE][33] of the group and the node staging [32][LDX]
constexpr int S_L = S_A + GROUPS * TILE_FLOATS;  // [GROUPS][TE][8]    logits / per-warp geometry scalars
constexpr int S_MS = S_L + GROUPS * TE * 8;      // [GROUPS][2][TE][8] softmax max, sum / basis mix
constexpr int S_POS = S_MS + GROUPS * 2 * TE * 8;  // [MAXN*3]
constexpr int S_GRAD = S_POS + MAXN * 3;     // [MAXN*3]  network output ("gradient")
constexpr int S_SCORE = S_GRAD + MAXN * 3;   // [MAXN*3]
constexpr int S_NOISE = S_SCORE + MAXN * 3;  // [MAXN*3]
constexpr int S_RED = S_NOISE + MAXN * 3;    // [64]
constexpr int S_FLOATS = S_RED + 64;
// int region (after the floats)
constexpr int SI_ROWL = 0;                   // [MAXN+1] edge offsets local to the chunk
constexpr int SI_TTGT = SI_ROWL + MAXN + 1;  // [MAXT+1] tile target boundaries local to the chunk
constexpr int SI_ESRC = SI_TTGT + MAXT + 1;  // [GROUPS][TE]
constexpr int SI_ETGT = SI_ESRC + GROUPS * TE;  // [GROUPS][TE]
constexpr int SI_MISC = SI_ETGT + GROUPS * TE;  // [4]  (0: work item, 1: TMEM base address)
constexpr int SI_BAR = ((S_FLOATS + SI_MISC + 4 + 1) & ~1) - S_FLOATS;  // [2] mbarrier of the tcgen05 commits (8-byte aligned)
constexpr int SI_TMABAR = SI_BAR + 2;        // [2] mbarrier of the TMA bulk weight copies (SI_MISC + 2 holds its phase)
constexpr int S_INTS = SI_TMABAR + 2;
// (source, target) of every edge slot of the first SLOT_CACHE_TILES tiles of the chunk as chunk-local byte indices (< MAXN <= 255):
// static per batch, resolved once per chunk (load_chunk) instead of by a dependent global load + bisection at each of the
// 7 tile visits of every score evaluation.  [tile][0: src, 1: tgt][TE] uint8, behind the int region.
constexpr int SLOT_CACHE_TILES = 31;
constexpr size_t SMEM_BYTES = sizeof(float) * S_FLOATS + sizeof(int) * S_INTS + static_cast<size_t>(SLOT_CACHE_TILES) * 2 * TE;
static_assert(MAXN <= 255, "slot cache stores chunk-local node indices as bytes");
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
static_assert(MOLSDE_P_E0_END <= 3 * 32 * MAXN, "E0 weights are staged in the q/k/v region");
static_assert(MOLSDE_P_BASIS_SZ <= 3 * 32 * MAXN, "basis weights (tcgen05 B tiles + small vectors) are staged in the q/k/v region");
static_assert(32 * LDX <= GROUPS * TILE_FLOATS, "node staging aliases the A region");
static_assert(TE * LDM <= TILE_FLOATS, "message tile aliases the group's A region");
static_assert(LDX >= MAXN && LDX % 32 == 8 && LDA % 32 == 8, "padded leading dimensions");
static_assert(S_A % 4 == 0 && S_WG % 4 == 0 && S_Q % 4 == 0 && TILE_FLOATS % 4 == 0, "16B alignment for cp.async");
// tcgen05 operand tiles of the basis MLP: A hi/lo [128 x 64] over the (dead) GAT-weight + A regions, B hi/lo in the q/k/v region
constexpr int UMMA_TILE = 128 * 64;          // floats per operand tile
constexpr int S_UA_HI = S_WG, S_UA_LO = S_WG + UMMA_TILE;
static_assert(S_UA_LO + UMMA_TILE <= S_L, "tcgen05 A tiles must fit before the logits region");
static_assert((S_WG * 4) % 128 == 0 && (S_Q * 4) % 128 == 0, "operand tiles are 128B aligned");
static_assert(GROUPS * TE * 8 >= 4 * TE * 4, "partial dyn buffer [4][TE][4]");
constexpr uint32_t UMMA_LBO = 2048, UMMA_SBO = 128;  // bytes: next 16B K-chunk / next 8-row group (K-major, no swizzle)
constexpr uint32_t TMEM_COLS = 256;  // two 128-column accumulators (tile t / t+1 of the basis GEMM pipeline)

// optional per-phase cycle accounting (build with MOLSDE_PROF=1; read back with molsde_debug_read_prof)
#ifdef MOLSDE_PROF
__device__ unsigned long long g_prof[kNumSMs][8];
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(slot) do { __syncthreads(); if (threadIdx.x == 0) { long long prof_t1 = clock64(); g_prof[blockIdx.x][slot] += prof_t1 - prof_t0; prof_t0 = prof_t1; } } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot)
#endif

struct Chunk {
    float* sm;
    int* si;
    int n;       // atoms
    int node0;   // first global node
    int edge0;   // first global CSR edge
    int tile0;   // first global tile
    int ntiles;
    // train mode (SDEModel2Dto3D_02.forward): keep-masks of the attention / FFN dropouts, NULL in eval mode
    const float* attn_keep;  // [4 layers][E][8] in CSR edge order (0 or 1)
    const float* ffn_keep;   // [4 layers][N][32]
    float inv_keep;          // 1 / (1 - p)
    int64_t E_total, N_total;
};

// ---------------------------------------------------------------------------------------
// fast, accuracy-checked elementwise helpers (absolute / relative error ~1e-6, far below the 1e-4 bar)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// sin/cos of an fp32 argument of magnitude up to ~1e5: 2-term Cody-Waite reduction by 2*pi (exact
// products through FMA), then the SFU approximations on |r| <= pi (abs error < 1e-6).
__device__ __forceinline__ void sincos_reduced(float x, float& s, float& c) {
    const float k = rintf(x * 0.15915494309189535f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484555314695172e-7f, r);
    s = __sinf(r);
    c = __cosf(r);
}

// cooperative global->shared copy of `nfloat` floats (multiple of 4, 16B aligned both sides)
__device__ __forceinline__ void stage_async(float* dst, const float* __restrict__ src, int nfloat) {
    for (int i = threadIdx.x * 4; i < nfloat; i += NTHREADS * 4) cp_async16(dst + i, src + i);
    cp_async_commit();
}

// Weight blocks of a phase (35-68 KB, contiguous in the parameter blob): ONE TMA bulk copy (cp.async.bulk, 1-D) issued by
// thread 0 and tracked by an mbarrier (complete_tx) instead of ~8 cp.async per thread.  Call with all threads after a
// __syncthreads() (the destination may still be read by the previous phase before that); returns when the data is
// visible to every thread.  The barrier's phase bit lives in shared memory (si[SI_MISC + 2]) because the number of copies
// per score evaluation is odd.
__device__ __forceinline__ void stage_bulk(int* si, float* dst, const float* __restrict__ src, int nfloat) {
    const uint32_t bar = static_cast<uint32_t>(__cvta_generic_to_shared(si + SI_TMABAR));
    const uint32_t phase = static_cast<uint32_t>(si[SI_MISC + 2]);
    if (threadIdx.x == 0) {
        const uint32_t bytes = static_cast<uint32_t>(nfloat) * 4u;
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of dst before the async-proxy write
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(dst))), "l"(src), "r"(bytes), "r"(bar) : "memory");
    }
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(phase) : "memory");
    __syncthreads();
    if (threadIdx.x == 0) si[SI_MISC + 2] = static_cast<int>(phase ^ 1u);  // read again only after several more barriers
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

// q / k / v rows are [node][32] with the column XOR-swizzled by the node index (4-float granules): the per-edge gathers
// k[src], v[src] of 8 consecutive slots then hit 8 different bank groups instead of one (profiles/r1_pc_v4_conflicts.txt)
__device__ __forceinline__ int qkv_idx(int node, int col) { return node * 32 + (col ^ ((node & 7) << 2)); }

// ---- group / tile helpers -------------------------------------------------------------------
// The 16 warps form two groups of 8; group g walks tiles g, g+2, ... of the chunk.  Inside a group
// warp `slab` owns edge slots [16*slab, 16*slab+16) of the current tile and the matching 16-column
// stripe of the group's A buffer, so producer -> GEMM -> epilogue chains need only __syncwarp();
// the group meets on a named barrier only where a target's edge segment may span stripes.
__device__ __forceinline__ void group_sync(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(GTHREADS) : "memory");
}

struct TileInfo {
    int ta, tb, ea, ne;  // first / end target (chunk-local), first edge (chunk-local), #edges
};
__device__ __forceinline__ TileInfo tile_info(const Chunk& c, int t) {
    const int* rowl = c.si + SI_ROWL;
    const int* ttgt = c.si + SI_TTGT;
    TileInfo ti;
    ti.ta = ttgt[t];
    ti.tb = ttgt[t + 1];
    ti.ea = rowl[ti.ta];
    ti.ne = rowl[ti.tb] - ti.ea;
    return ti;
}

__device__ __forceinline__ uint8_t* slot_cache(const Chunk& c) { return reinterpret_cast<uint8_t*>(c.si + S_INTS); }

// (source, target) of edge slot `slot` of tile t, chunk-local: bisection on the row pointer + one global load
__device__ __forceinline__ void resolve_slot(const Chunk& c, const int32_t* __restrict__ src_g, const TileInfo& ti, int slot,
                                             int& sj, int& tg) {
    const int* rowl = c.si + SI_ROWL;
    sj = 0;
    tg = ti.ta;
    if (slot < ti.ne) {
        const int e = ti.ea + slot;
        int lo = ti.ta, hi = ti.tb;  // largest i in [ta, tb) with rowl[i] <= e
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (rowl[mid] <= e) lo = mid; else hi = mid;
        }
        tg = lo;
        sj = src_g[c.edge0 + e] - c.node0;
    }
}

// warp-private bookkeeping: lanes 0..15 fetch (source, target) of their slot (dead slots: source 0, target = first of the tile)
__device__ __forceinline__ void slot_edges(const Chunk& c, const int32_t* __restrict__ src_g, const TileInfo& ti, int t,
                                           int slab, int lane, int* esrc, int* etgt) {
    if (lane < 16) {
        const int slot = slab * 16 + lane;
        int sj, tg;
        if (t < SLOT_CACHE_TILES) {
            const uint8_t* sc = slot_cache(c) + t * 2 * TE;
            sj = sc[slot];
            tg = sc[TE + slot];
        } else {
            resolve_slot(c, src_g, ti, slot, sj, tg);
        }
        esrc[slot] = sj;
        etgt[slot] = tg;
    }
    __syncwarp();
}

// copy the warp's 16-column stripe of a [32][LDA] tile from global into its smem stripe (4 x 16 B per lane)
__device__ __forceinline__ void load_stripe_async(float* stripe, const float* __restrict__ tile_stripe, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int chunk = lane + 32 * i, k = chunk >> 2, c4 = (chunk & 3) * 4;
        cp_async16(stripe + k * LDA + c4, tile_stripe + k * LDA + c4);
    }
    cp_async_commit();
}

// ---------------------------------------------------------------------------------------
// Phase E0: per-edge attribute  edge_attr = input_mlp(gfp(d)) * e2d + project([sin,cos,emb_i,emb_j])
// SDE_model_2D_to_3D.py:402-432.  coff_mlp (a bare Linear) is folded into project.layers.0 on the
// host (MOLSDE_P_H_W), so the hidden layer accumulates directly over the four Fourier blocks.
// Entirely warp-private: no block- or group-level barrier inside the tile loop.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ void phase_edge_features(const Chunk c, const float* __restrict__ blob,
                                                 const int32_t* __restrict__ src_g, const float* __restrict__ e2d_tiles,
                                                 float* __restrict__ scratch) {
    float* sm = c.sm;
    float* W = sm + S_Q;  // E0 weights staged over the (currently dead) q/k/v region
    const float* pos = sm + S_POS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 3, slab = warp & 7;
    const int g = lane >> 2, t4 = lane & 3;
    float* A = sm + S_A + grp * TILE_FLOATS + slab * 16;   // stripe: element (k, r) at A[k*LDA + r]
    float* geo = sm + S_L + grp * (TE * 8) + slab * 112;    // [7][16]: d, ci0, ci2, cj0, cj2, psin, pcos
    int* esrc = c.si + SI_ESRC + grp * TE;
    int* etgt = c.si + SI_ETGT + grp * TE;
    __syncthreads();  // the q/k/v region may still be read by the tail of the previous evaluation
    stage_bulk(c.si, W, blob, MOLSDE_P_E0_END);
    for (int t = grp; t < c.ntiles; t += GROUPS) {
        const TileInfo ti = tile_info(c, t);
        slot_edges(c, src_g, ti, t, slab, lane, esrc, etgt);
        // this thread's 16 elements of the e2d tile (rows g, g+8; columns nb*8 + 2*t4 + j), fetched early
        const float* e2d_t = e2d_tiles + static_cast<size_t>(c.tile0 + t) * TILE_FLOATS + slab * 16;
        float e2[4][4];
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                    e2[nb][2 * rr + j] = __ldg(e2d_t + (nb * 8 + 2 * t4 + j) * LDA + g + 8 * rr);
        if (lane < 16) {
            const int slot = slab * 16 + lane;
            float gq[7] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (slot < ti.ne) {
                const float* pr = pos + 3 * esrc[slot];  // row = source j
                const float* pc = pos + 3 * etgt[slot];  // col = target i
                const Frame f = coord2basis(pr, pc);
                {   // cache the equivariant basis of this edge for the two basis phases (equivariant_scorenetwork.py:159)
                    float* fr_t = scratch + static_cast<size_t>(t) * SCR_TILE + TILE_FLOATS + slot;
                    fr_t[0 * TE] = f.dx; fr_t[1 * TE] = f.dy; fr_t[2 * TE] = f.dz;
                    fr_t[3 * TE] = f.cx; fr_t[4 * TE] = f.cy; fr_t[5 * TE] = f.cz;
                    fr_t[6 * TE] = f.vx; fr_t[7 * TE] = f.vy; fr_t[8 * TE] = f.vz;
                }
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
                // :425  sqrt(1 - cos^2).  For (anti)parallel r_i, r_j rounding can make the argument a tiny negative
                // number and the reference then returns NaN for the whole batch; the limit value 0 is used instead
                // (only inputs on which the reference output is NaN are affected).
                const float psin = sqrtf(fmaxf(__fsub_rn(1.0f, __fmul_rn(pcos, pcos)), 0.0f));
                gq[0] = f.dist; gq[1] = ci0; gq[2] = ci2; gq[3] = cj0; gq[4] = cj2; gq[5] = psin; gq[6] = pcos;
            }
#pragma unroll
            for (int q = 0; q < 7; ++q) geo[q * 16 + lane] = gq[q];
        }
        __syncwarp();
        // Fourier block 0 (distance) feeds input_mlp (:409-410); blocks 1..4 (ci0, ci2, cj0, cj2) feed the fused
        // hidden layer of `project` (:427-430).  Each block: sin half -> GEMM(K=32), cos half -> GEMM(K=32).
        float inv[4][4], acc[4][4];
        zero_frag(acc);
        // lane -> (edge fe, half hf); frequency of step i is w(i) = 4*(i/2) + 2*hf + (i&1): at every step the two half-warps
        // write A rows 2 apart = 16 banks apart (conflict-free), profiles/r1_pc_v4_conflicts.txt
        const int fe = lane & 15, hf = lane >> 4;
#pragma unroll 1
        for (int blk = 0; blk < 5; ++blk) {
            const float x = geo[blk * 16 + fe];
            const float* Wf = W + (blk == 0 ? MOLSDE_P_GFP_DIST_W : MOLSDE_P_GFP_COFF_W);
            const float* Wm = (blk == 0) ? W + MOLSDE_P_IN_W : W + MOLSDE_P_H_W + (blk - 1) * 64 * LD32;
            float cs[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // GaussianFourierProjection.forward, :64-66  (x * W * 2 * pi, fp32, in that order)
                const int w = 4 * (i >> 1) + 2 * hf + (i & 1);
                const float arg = __fmul_rn(__fmul_rn(__fmul_rn(x, Wf[w]), 2.0f), 3.14159274101257324f);
                float sn;
                sincos_reduced(arg, sn, cs[i]);
                A[w * LDA + fe] = sn;
            }
            __syncwarp();
            MOLSDE_MMA_GEMM<4, LDA, LD32>(A, Wm, 32, lane, acc);
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 16; ++i) A[(4 * (i >> 1) + 2 * hf + (i & 1)) * LDA + fe] = cs[i];
            __syncwarp();
            MOLSDE_MMA_GEMM<4, LDA, LD32>(A, Wm + 32 * LD32, 32, lane, acc);
            __syncwarp();
            if (blk == 0) {
#pragma unroll
                for (int nb = 0; nb < 4; ++nb)
#pragma unroll
                    for (int q = 0; q < 4; ++q) { inv[nb][q] = acc[nb][q]; acc[nb][q] = 0.0f; }
            }
        }
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = nb * 8 + 2 * t4 + j;
                const float bh = W[MOLSDE_P_H_B + col], ws = W[MOLSDE_P_H_WSIN + col], wc = W[MOLSDE_P_H_WCOS + col];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int r = g + 8 * rr;
                    const float v = acc[nb][2 * rr + j] + bh + geo[5 * 16 + r] * ws + geo[6 * 16 + r] * wc;
                    A[col * LDA + r] = silu_fast(v);
                }
            }
        __syncwarp();
        zero_frag(acc);
        MOLSDE_MMA_GEMM<4, LDA, LD32>(A, W + MOLSDE_P_P1_W, 32, lane, acc);
        // ---- edge_attr = inv3d * e2d + frame  (:432) -> scratch tile (own stripe) ----
        float* sc_t = scratch + static_cast<size_t>(t) * SCR_TILE + slab * 16;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = nb * 8 + 2 * t4 + j;
                const float bi = W[MOLSDE_P_IN_B + col], bf = W[MOLSDE_P_P1_B + col];
#pragma unroll
                for (int rr = 0; rr < 2; ++rr)
                    sc_t[col * LDA + g + 8 * rr] = fmaf(inv[nb][2 * rr + j] + bi, e2[nb][2 * rr + j], acc[nb][2 * rr + j] + bf);
            }
        __syncwarp();
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------
// GAT layer pieces  (equivariant_scorenetwork.py:34-40, TransformerConv heads=8 C=4)
// ---------------------------------------------------------------------------------------
// q|k|v = Linear(x): each warp owns 16 nodes and all 96 output columns
__device__ __noinline__ void node_qkv(const Chunk c) {
    float* sm = c.sm;
    const float* Wg = sm + S_WG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = warp * 16, g = lane >> 2, t4 = lane & 3;
    if (m0 >= c.n) return;
    float acc[12][4];
    zero_frag(acc);
    MOLSDE_MMA_GEMM<12, LDX, LD96>(sm + S_XT + m0, Wg + MOLSDE_G_WQKV, 32, lane, acc);
#pragma unroll
    for (int nb = 0; nb < 12; ++nb) {
        const int col = nb * 8 + 2 * t4;  // 0..95: q | k | v
        const float b0 = Wg[MOLSDE_G_BQKV + col], b1 = Wg[MOLSDE_G_BQKV + col + 1];
        float* dst = sm + S_Q + (col >> 5) * (32 * MAXN);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int node = m0 + g + 8 * rr;
            if (node < c.n)
                *reinterpret_cast<float2*>(dst + qkv_idx(node, col & 31)) = make_float2(acc[nb][2 * rr] + b0, acc[nb][2 * rr + 1] + b1);
        }
    }
}

// attention over the incoming edges of every target: logits, segment softmax (+1e-16), weighted
// messages, deterministic ascending-source sum; the aggregate overwrites q[target].
__device__ __noinline__ void gat_edge_phase(const Chunk c, const int32_t* __restrict__ src_g,
                                            const float* __restrict__ scratch, int layer) {
    float* sm = c.sm;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int grp = warp >> 3, slab = warp & 7, gt = tid & (GTHREADS - 1);
    const int g = lane >> 2, t4 = lane & 3;
    float* Ag = sm + S_A + grp * TILE_FLOATS;
    float* stripe = Ag + slab * 16;
    float* Mm = Ag;  // [TE][LDM] slot-major messages, written only after the group's GEMMs are done
    float* L = sm + S_L + grp * (TE * 8);
    float* ssum = sm + S_MS + grp * (2 * TE * 8) + TE * 8;
    float* Q = sm + S_Q;
    const float* Kk = sm + S_K;
    const float* V = sm + S_V;
    const float* Wg = sm + S_WG;
    const int* rowl = c.si + SI_ROWL;
    int* esrc = c.si + SI_ESRC + grp * TE;
    int* etgt = c.si + SI_ETGT + grp * TE;
    for (int t = grp; t < c.ntiles; t += GROUPS) {
        const TileInfo ti = tile_info(c, t);
        load_stripe_async(stripe, scratch + static_cast<size_t>(t) * SCR_TILE + slab * 16, lane);
        slot_edges(c, src_g, ti, t, slab, lane, esrc, etgt);
        cp_async_wait<0>();
        __syncwarp();
        // e = lin_edge(edge_attr); this thread: slots 16*slab + g, +8; columns nb*8 + 2*t4 + {0,1}
        float e[4][4];
        zero_frag(e);
        MOLSDE_MMA_GEMM<4, LDA, LD32>(stripe, Wg + MOLSDE_G_WE, 32, lane, e);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int s = slab * 16 + g + 8 * rr;
            const bool live = s < ti.ne;
            const int sj = esrc[s], tg = etgt[s];
#pragma unroll
            for (int nb = 0; nb < 4; ++nb) {
                const int col = nb * 8 + 2 * t4;
                const float2 k2 = *reinterpret_cast<const float2*>(Kk + qkv_idx(sj, col));
                const float2 q2 = *reinterpret_cast<const float2*>(Q + qkv_idx(tg, col));
                const float2 v2 = *reinterpret_cast<const float2*>(V + qkv_idx(sj, col));
                // alpha = (q_i . (k_j + e)) / sqrt(C): a head (4 columns) is split over the lane pair (t4, t4^1)
                float part = fmaf(q2.y, k2.y + e[nb][2 * rr + 1], q2.x * (k2.x + e[nb][2 * rr]));
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                if (live && (t4 & 1) == 0) L[s * 8 + (col >> 2)] = part * 0.5f;
                e[nb][2 * rr] += v2.x;      // v_j + e
                e[nb][2 * rr + 1] += v2.y;
            }
        }
        group_sync(grp);
        // per (target, head): max and sum(exp) over the target's contiguous edge segment
        const int ntg = ti.tb - ti.ta;
        for (int p = gt; p < ntg * 8; p += GTHREADS) {
            const int i = ti.ta + (p >> 3), hd = p & 7;
            const int s0 = rowl[i] - ti.ea, s1 = rowl[i + 1] - ti.ea;
            float m = -CUDART_INF_F;
            for (int s = s0; s < s1; ++s) m = fmaxf(m, L[s * 8 + hd]);
            float z = 0.0f;
            for (int s = s0; s < s1; ++s) {   // the exponentials are kept (in place of the logits) for the normalisation pass
                const float ex = __expf(L[s * 8 + hd] - m);
                L[s * 8 + hd] = ex;
                z += ex;
            }
            ssum[p] = z;
        }
        group_sync(grp);
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int s = slab * 16 + g + 8 * rr;
            if (s < ti.ne) {
                const int pb = (etgt[s] - ti.ta) * 8;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int col = nb * 8 + 2 * t4, hd = col >> 2;
                    float a = __fdividef(L[s * 8 + hd], ssum[pb + hd] + 1e-16f);
                    if (c.attn_keep)  // F.dropout(alpha, p) in train mode (TransformerConv.message)
                        a *= c.attn_keep[(static_cast<size_t>(layer) * c.E_total + c.edge0 + ti.ea + s) * 8 + hd] * c.inv_keep;
                    Mm[s * LDM + col] = e[nb][2 * rr] * a;
                    Mm[s * LDM + col + 1] = e[nb][2 * rr + 1] * a;
                }
            }
        }
        group_sync(grp);
        for (int p = gt; p < ntg * 32; p += GTHREADS) {
            const int i = ti.ta + (p >> 5), col = p & 31;
            const int s0 = rowl[i] - ti.ea, s1 = rowl[i + 1] - ti.ea;
            float acc = 0.0f;
            for (int s = s0; s < s1; ++s) acc += Mm[s * LDM + col];
            Q[qkv_idx(i, col)] = acc;
        }
        group_sync(grp);
    }
}

// LayerNorm over the 32 columns of two rows held by a lane quad (8 columns per lane and row)
__device__ __forceinline__ void layer_norm_quad(float (&v)[4][4], const float* __restrict__ w, const float* __restrict__ b,
                                                int t4) {
#pragma unroll
    for (int rr = 0; rr < 2; ++rr) {
        float s = 0.0f;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) s += v[nb][2 * rr] + v[nb][2 * rr + 1];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const float mean = s * (1.0f / 32.0f);
        float q = 0.0f;
#pragma unroll
        for (int nb = 0; nb < 4; ++nb) {
            const float d0 = v[nb][2 * rr] - mean, d1 = v[nb][2 * rr + 1] - mean;
            q += d0 * d0 + d1 * d1;
        }
        q += __shfl_xor_sync(0xffffffffu, q, 1);
        q += __shfl_xor_sync(0xffffffffu, q, 2);
        const float rstd = 1.0f / sqrtf(q * (1.0f / 32.0f) + LN_EPS);
#pragma unroll
        for (int nb = 0; nb < 4; ++nb)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const int col = nb * 8 + 2 * t4 + j;
                v[nb][2 * rr + j] = (v[nb][2 * rr + j] - mean) * rstd * w[col] + b[col];
            }
    }
}

// x <- x + LN1(agg + skip(x));  x <- x + LN2(FFN(x));  optional SiLU  (equivariant_scorenetwork.py:35-38,140-141)
// Each warp owns 16 nodes end to end (only __syncwarp between its GEMMs).
__device__ __noinline__ void node_update(const Chunk c, bool silu_after, int layer) {
    float* sm = c.sm;
    float* XT = sm + S_XT;
    float* NT = sm + S_A;  // [32][LDX] staging of the FFN input / hidden, k-major
    const float* Q = sm + S_Q;
    const float* Wg = sm + S_WG;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int m0 = warp * 16, g = lane >> 2, t4 = lane & 3;
    if (m0 >= c.n) return;
    float acc[4][4], x1[4][4];
    zero_frag(acc);
    MOLSDE_MMA_GEMM<4, LDX, LD32>(XT + m0, Wg + MOLSDE_G_WS, 32, lane, acc);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb) {
        const int col = nb * 8 + 2 * t4;
        const float b0 = Wg[MOLSDE_G_BS + col], b1 = Wg[MOLSDE_G_BS + col + 1];
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int node = m0 + g + 8 * rr;
            float2 ag = make_float2(0.f, 0.f);
            if (node < c.n) ag = *reinterpret_cast<const float2*>(Q + qkv_idx(node, col));
            acc[nb][2 * rr] += b0 + ag.x;
            acc[nb][2 * rr + 1] += b1 + ag.y;
        }
    }
    layer_norm_quad(acc, Wg + MOLSDE_G_LN1_W, Wg + MOLSDE_G_LN1_B, t4);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = nb * 8 + 2 * t4 + j;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int node = m0 + g + 8 * rr;
                const float xo = (node < c.n) ? XT[col * LDX + node] : 0.0f;
                x1[nb][2 * rr + j] = xo + acc[nb][2 * rr + j];
                NT[col * LDX + node] = x1[nb][2 * rr + j];
            }
        }
    __syncwarp();
    zero_frag(acc);
    MOLSDE_MMA_GEMM<4, LDX, LD32>(NT + m0, Wg + MOLSDE_G_F0, 32, lane, acc);
    __syncwarp();
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = nb * 8 + 2 * t4 + j;
            const float bj = Wg[MOLSDE_G_F0_B + col];
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int node = m0 + g + 8 * rr;
                float hv = silu_fast(acc[nb][2 * rr + j] + bj);
                if (c.ffn_keep && node < c.n)  // nn.Dropout between FFN.1 (SiLU) and FFN.3 in train mode
                    hv *= c.ffn_keep[(static_cast<size_t>(layer) * c.N_total + c.node0 + node) * 32 + col] * c.inv_keep;
                NT[col * LDX + node] = hv;
            }
        }
    __syncwarp();
    zero_frag(acc);
    MOLSDE_MMA_GEMM<4, LDX, LD32>(NT + m0, Wg + MOLSDE_G_F3, 32, lane, acc);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const float bj = Wg[MOLSDE_G_F3_B + nb * 8 + 2 * t4 + j];
            acc[nb][j] += bj;
            acc[nb][2 + j] += bj;
        }
    layer_norm_quad(acc, Wg + MOLSDE_G_LN2_W, Wg + MOLSDE_G_LN2_B, t4);
#pragma unroll
    for (int nb = 0; nb < 4; ++nb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            const int col = nb * 8 + 2 * t4 + j;
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int node = m0 + g + 8 * rr;
                float x2 = x1[nb][2 * rr + j] + acc[nb][2 * rr + j];
                if (silu_after) x2 = silu_fast(x2);
                if (node < c.n) XT[col * LDX + node] = x2;
            }
        }
}

// ---------------------------------------------------------------------------------------
// tcgen05 helpers (descriptor formats: cute/arch/mma_sm100_desc.hpp; bring-up test tools/ubench/tcgen05_gemm.cu)
// Operand tiles are K-major in the canonical no-swizzle core-matrix layout (8 rows x 16 B):
//   float index(r, k) = (k/4)*(R/8)*32 + (r/8)*32 + (r%8)*4 + (k%4),  LBO = (R/8)*128 B, SBO = 128 B.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4)                   // start address  [0,14)
           | (static_cast<uint64_t>(lbo_bytes >> 4) << 16)                  // leading byte offset [16,30)
           | (static_cast<uint64_t>(UMMA_SBO >> 4) << 32)                   // stride byte offset  [32,46)
           | (static_cast<uint64_t>(1) << 46);                              // version 1 (Blackwell), layout = no swizzle
}
// D[tmem] (+)= A[smem] . B[smem]^T, M = 128, K = 8 (tf32), issued by ONE thread
template <int N>
__device__ __forceinline__ void umma_tf32_m128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    // instruction descriptor: D = F32 (1<<4), A = B = TF32 (2<<7, 2<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((static_cast<uint32_t>(N) >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0;
    for (int it = 0; it < (1 << 24) && !done; ++it)  // bounded: a descriptor bug must not hang the GPU
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    return done != 0;
}
// 3xTF32: three passes (lo*hi, hi*lo, hi*hi) over `ksteps` K=8 steps of an A tile [128 x 8*ksteps] and a B tile [N x ...]
template <int N>
__device__ __forceinline__ void umma_3xtf32(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, int ksteps,
                                            uint32_t lbo_a, uint32_t lbo_b, uint32_t accumulate) {
#pragma unroll 1
    for (int term = 0; term < 3; ++term) {
        const uint32_t pa = (term == 0) ? a_lo : a_hi, pb = (term == 1) ? b_lo : b_hi;
#pragma unroll 1
        for (int kb = 0; kb < ksteps; ++kb) {
            umma_tf32_m128<N>(tmem_d, umma_desc(pa + kb * 2 * lbo_a, lbo_a), umma_desc(pb + kb * 2 * lbo_b, lbo_b), accumulate);
            accumulate = 1;
        }
    }
}
// 8 consecutive accumulator columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
// write 4 consecutive-k values of row `r` (k-chunk kc) of a [128 x K] A tile as tf32 hi + exact lo
__device__ __forceinline__ void store_a_chunk(float* AH, float* AL, int r, int kc, const float (&v)[4]) {
    float4 h4, l4;
    h4.x = __uint_as_float(tf32_hi(v[0])); h4.y = __uint_as_float(tf32_hi(v[1]));
    h4.z = __uint_as_float(tf32_hi(v[2])); h4.w = __uint_as_float(tf32_hi(v[3]));
    l4.x = v[0] - h4.x; l4.y = v[1] - h4.y; l4.z = v[2] - h4.z; l4.w = v[3] - h4.w;
    const int idx = kc * 512 + (r >> 3) * 32 + (r & 7) * 4;
    *reinterpret_cast<float4*>(AH + idx) = h4;
    *reinterpret_cast<float4*>(AL + idx) = l4;
}

// ---------------------------------------------------------------------------------------
// basis MLP + equivariant mean aggregation  (equivariant_scorenetwork.py:154-164)
//   hidden[128 edges x 128] = [h_row + h_col | edge_attr][128 x 64] . W1^T on tcgen05 (M=128, N=128, 8 K-steps x 3 split
//   terms, accumulator in TMEM); epilogue TMEM -> registers: +bias, SiLU, 128 -> 3 projection; frame mix; per-target mean.
// All 16 warps work on one tile at a time; returns the updated mbarrier phase.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t phase_basis(const Chunk c, const float* __restrict__ blob, const int32_t* __restrict__ src_g,
                                             const float* __restrict__ scratch, int module, uint32_t tmem_base, uint32_t phase,
                                             int32_t* status_flag) {
    float* sm = c.sm;
    float* Wb = sm + S_Q;  // staged over q/k/v (dead between GAT blocks)
    float* AH = sm + S_UA_HI;
    float* AL = sm + S_UA_LO;
    float* dynp = sm + S_L;   // [4][TE][4] partial dyn coefficients of the four column blocks
    float* frames = sm + S_MS;  // [2][9][TE] cached edge frames of tile t / t+1
    float* mix = sm + S_MS + 2 * FRAME_FLOATS;  // [3][TE]
    const float* XT = sm + S_XT;
    float* grad = sm + S_GRAD;
    const int* rowl = c.si + SI_ROWL;
    const uint32_t bar = smem_u32(c.si + SI_BAR);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    stage_bulk(c.si, Wb, blob + MOLSDE_P_BASIS0 + module * MOLSDE_P_BASIS_SZ, MOLSDE_P_BASIS_SZ);

    // stage 1 of the software pipeline: slot bookkeeping + A operand of tile t + MMA issue (asynchronous)
    auto produce_and_issue = [&](int t) {
        const TileInfo ti = tile_info(c, t);
        int* esrc = c.si + SI_ESRC + (t & 1) * TE;
        int* etgt = c.si + SI_ETGT + (t & 1) * TE;
        // items (e, kc) of the A operand: kc = (tid >> 7) + 4 i.  i = 2, 3 (kc >= 8) are the edge_attr half, read straight from
        // the L2-resident scratch record (written by E0): issued FIRST, their latency overlaps the slot bookkeeping, the
        // barrier and the shared-memory gathers of i = 0, 1.
        static_assert(NTHREADS == 4 * TE, "item mapping of the basis A operand");
        const float* sc_t = scratch + static_cast<size_t>(t) * SCR_TILE;
        const int e = tid & (TE - 1), kq = tid >> 7;
        float va[2][4];
#pragma unroll
        for (int i = 0; i < 2; ++i)
#pragma unroll
            for (int q = 0; q < 4; ++q)   // the record is rewritten by E0 of every evaluation of this launch: L2-coherent load, never .nc
                va[i][q] = __ldcg(sc_t + ((kq + 4 * i) * 4 + q) * LDA + e);
        if (tid < TE) {  // (source, target) of every slot
            int sj, tg;
            if (t < SLOT_CACHE_TILES) {
                const uint8_t* sc = slot_cache(c) + t * 2 * TE;
                sj = sc[tid];
                tg = sc[TE + tid];
            } else {
                resolve_slot(c, src_g, ti, tid, sj, tg);
            }
            esrc[tid] = sj;
            etgt[tid] = tg;
        }
        __syncthreads();
        // A operand [128 x 64], K-major canonical core-matrix layout, split into tf32 hi + exact lo:
        //   k < 32: h_row + h_col (:154-155),  k >= 32: edge_attr (scratch tile)
        stage_async(frames + (t & 1) * FRAME_FLOATS, sc_t + TILE_FLOATS, FRAME_FLOATS);  // consumed by the epilogue of tile t
        {
            const int sj = esrc[e], tg = etgt[e];
            const bool live = e < ti.ne;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int kc = kq + 4 * i, k0 = kc * 4;
                float v[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) v[q] = live ? XT[(k0 + q) * LDX + sj] + XT[(k0 + q) * LDX + tg] : 0.0f;
                store_a_chunk(AH, AL, e, kc, v);
            }
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) store_a_chunk(AH, AL, e, kq + 4 * i + 8, va[i]);
        cp_async_wait<0>();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the tensor core
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (tid == 0) {
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            umma_3xtf32<128>(tmem_base + (t & 1) * 128, smem_u32(AH), smem_u32(AL), smem_u32(Wb + MOLSDE_B_W1C_HI),
                             smem_u32(Wb + MOLSDE_B_W1C_LO), 8, UMMA_LBO, UMMA_LBO, 0u);
            umma_commit(bar);
        }
    };

    bool ok = true;
    if (c.ntiles > 0) produce_and_issue(0);
    for (int t = 0; t < c.ntiles; ++t) {
        const TileInfo ti = tile_info(c, t);
        ok &= mbar_wait(bar, phase);  // MMAs of tile t complete: accumulator (t&1) ready, A buffer free
        phase ^= 1u;
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (t + 1 < c.ntiles) produce_and_issue(t + 1);  // its MMAs overlap the epilogue below
        {   // epilogue: warp -> TMEM lane quarter (edge slots 32*lq..) x column block cb (hidden units 32*cb..)
            const int lq = warp & 3, cb = warp >> 2;
            uint32_t v[32];
            const uint32_t taddr = tmem_base + (t & 1) * 128 + (static_cast<uint32_t>(lq * 32) << 16) + cb * 32;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
                  "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
                  "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
                  "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const int col = cb * 32 + j;
                const float hv = silu_fast(__uint_as_float(v[j]) + Wb[MOLSDE_B_B1 + col]);
                p0 = fmaf(hv, Wb[MOLSDE_B_W2 + col], p0);
                p1 = fmaf(hv, Wb[MOLSDE_B_W2 + 128 + col], p1);
                p2 = fmaf(hv, Wb[MOLSDE_B_W2 + 256 + col], p2);
            }
            float* dp = dynp + (cb * TE + lq * 32 + lane) * 4;
            dp[0] = p0; dp[1] = p1; dp[2] = p2;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        // basis_mix = dyn0 * coord_diff + dyn1 * coord_cross + dyn2 * coord_vertical  (:159), frames cached by E0
        const float* F = frames + (t & 1) * FRAME_FLOATS;
        if (tid < 3 * TE) {
            const int q = tid & (TE - 1), ax = tid >> 7;
            if (q < ti.ne) {
                float d[3];
#pragma unroll
                for (int o = 0; o < 3; ++o)
                    d[o] = ((dynp[q * 4 + o] + dynp[(TE + q) * 4 + o]) + (dynp[(2 * TE + q) * 4 + o] + dynp[(3 * TE + q) * 4 + o])) +
                           Wb[MOLSDE_B_B2 + o];
                mix[ax * TE + q] = d[0] * F[ax * TE + q] + d[1] * F[(3 + ax) * TE + q] + d[2] * F[(6 + ax) * TE + q];
            }
        }
        __syncthreads();
        // gradient_i (+)= mean over the incoming edges (:162-164), ascending-source order
        const int ntg = ti.tb - ti.ta;
        for (int p = tid; p < ntg * 3; p += NTHREADS) {
            const int i = ti.ta + p / 3, ax = p % 3;
            const int s0 = rowl[i] - ti.ea, s1 = rowl[i + 1] - ti.ea;
            float sacc = 0.0f;
            for (int q = s0; q < s1; ++q) sacc += mix[ax * TE + q];
            sacc = __fdiv_rn(sacc, static_cast<float>(max(s1 - s0, 1)));  // aggr='mean'
            grad[i * 3 + ax] = (module == 0) ? sacc : grad[i * 3 + ax] + sacc;
        }
        __syncthreads();
    }
    if (!ok && tid == 0 && status_flag) atomicExch(status_flag, -7);
    return phase;
}

// ---------------------------------------------------------------------------------------
// one full network evaluation on the chunk: positions in smem -> "gradient" in smem
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t score_eval(const Chunk c, const float* __restrict__ blob, const int32_t* __restrict__ src_g,
                                            const float* __restrict__ nattr, const float* __restrict__ e2d_tiles,
                                            float* __restrict__ scratch, uint32_t tmem_base, uint32_t phase,
                                            int32_t* status_flag) {
    float* sm = c.sm;
    PROF_T0();
    phase_edge_features(c, blob, src_g, e2d_tiles, scratch);
    PROF_ADD(0);
    // conv_input = node_attr (loop-invariant node_emb output), k-major
    for (int idx = threadIdx.x; idx < c.n * 32; idx += NTHREADS) {
        const int node = idx >> 5, k = idx & 31;
        sm[S_XT + k * LDX + node] = __ldg(nattr + static_cast<size_t>(c.node0 + node) * 32 + k);
    }
    __syncthreads();
    for (int module = 0; module < 2; ++module) {
        for (int conv = 0; conv < 2; ++conv) {
            stage_bulk(c.si, sm + S_WG, blob + MOLSDE_P_GAT0 + (2 * module + conv) * MOLSDE_P_GAT_SZ, MOLSDE_P_GAT_SZ);
            PROF_ADD(1);
            node_qkv(c);
            __syncthreads();
            PROF_ADD(2);
            gat_edge_phase(c, src_g, scratch, 2 * module + conv);
            __syncthreads();
            PROF_ADD(3);
            node_update(c, conv == 0, 2 * module + conv);
            __syncthreads();
            PROF_ADD(4);
        }
        phase = phase_basis(c, blob, src_g, scratch, module, tmem_base, phase, status_flag);
        PROF_ADD(5);
    }
    return phase;
}

// TMEM accumulator (128 columns) + mbarrier for the tcgen05 basis GEMM; call with all threads of the CTA
__device__ __forceinline__ uint32_t tmem_setup(int* si) {
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(si + SI_BAR)));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(si + SI_TMABAR)));
        si[SI_MISC + 2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(si + SI_MISC + 1)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    return static_cast<uint32_t>(si[SI_MISC + 1]);
}
__device__ __forceinline__ void tmem_teardown(uint32_t tmem_base) {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

__device__ __forceinline__ bool load_chunk(Chunk& c, const molsde_plan& plan, int chunk, int32_t* status_flag) {
    // (also fills the slot cache of the chunk's first SLOT_CACHE_TILES tiles)
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
        if (ok) {
            uint8_t* sc = slot_cache(c);
            const int nt = c.ntiles < SLOT_CACHE_TILES ? c.ntiles : SLOT_CACHE_TILES;
            for (int item = tid; item < nt * TE; item += NTHREADS) {
                const int t = item / TE, slot = item % TE;
                const TileInfo ti = tile_info(c, t);
                int sj, tg;
                resolve_slot(c, plan.src, ti, slot, sj, tg);
                sc[t * 2 * TE + slot] = static_cast<uint8_t>(sj);
                sc[t * 2 * TE + TE + slot] = static_cast<uint8_t>(tg);
            }
            __syncthreads();
        }
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
                     int64_t scratch_stride, int32_t* status_flag, const float* __restrict__ attn_keep,
                     const float* __restrict__ ffn_keep, float inv_keep) {
    extern __shared__ __align__(128) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(smem + S_FLOATS);
    c.attn_keep = attn_keep; c.ffn_keep = ffn_keep; c.inv_keep = inv_keep;
    c.E_total = plan.E; c.N_total = plan.N;
    float* my_scratch = scratch + static_cast<size_t>(blockIdx.x) * scratch_stride;
    const uint32_t tmem_base = tmem_setup(c.si);
    uint32_t phase = 0;
    for (int chunk = blockIdx.x; chunk < plan.num_chunks; chunk += gridDim.x) {
        __syncthreads();
        if (!load_chunk(c, plan, chunk, status_flag)) continue;
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) smem[S_POS + i] = pos[static_cast<size_t>(c.node0) * 3 + i];
        __syncthreads();
        phase = score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch, tmem_base, phase, status_flag);
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) {
            // get_score: scores = -output / std  (:440-443);  forward (stdv == NULL): the raw network output (:379)
            score[static_cast<size_t>(c.node0) * 3 + i] = stdv ? __fdiv_rn(-smem[S_GRAD + i], stdv[c.node0 + i / 3]) : smem[S_GRAD + i];
        }
    }
    tmem_teardown(tmem_base);
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
__device__ __noinline__ void normal3(uint64_t seed, uint32_t node, uint32_t step, uint32_t stream, float* out) {
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
    for (int w = 0; w < NWARPS; ++w) s += red[w];
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
    extern __shared__ __align__(128) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(smem + S_FLOATS);
    int* misc = c.si + SI_MISC;
    c.attn_keep = nullptr; c.ffn_keep = nullptr; c.inv_keep = 1.0f; c.E_total = plan.E; c.N_total = plan.N;
    const uint32_t tmem_base = tmem_setup(c.si);
    uint32_t phase = 0;
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
        if (misc[0] >= plan.num_chunks) break;
        const int chunk = plan.chunk_order ? plan.chunk_order[misc[0]] : misc[0];  // longest groups first
        if (!load_chunk(c, plan, chunk, status_flag)) continue;
        const int n3 = c.n * 3;
        const size_t g3 = static_cast<size_t>(c.node0) * 3;
        for (int i = tid; i < n3; i += NTHREADS) P[i] = pos_init[g3 + i];
        __syncthreads();
        for (int step = 0; step < cfg.steps; ++step) {
            const float stdv = step_table[step * 8 + 0], Gd = step_table[step * 8 + 1];
            const float sqrt_alpha = step_table[step * 8 + 2], calpha = step_table[step * 8 + 3];
            // ---------------- corrector (LangevinCorrector.update_fn :191-212) ----------------
            phase = score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch, tmem_base, phase, status_flag);
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
            phase = score_eval(c, blob, plan.src, nattr, e2d_tiles, my_scratch, tmem_base, phase, status_flag);
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
    tmem_teardown(tmem_base);
}

// ---------------------------------------------------------------------------------------
// edge_2D_emb (eval): e2d tile = W3 . relu(U[src] + V[tgt]) + b3,  SDE_model_2D_to_3D.py:405-407
// uv [N][600]: columns 0..299 = folded first layer applied to h[row], 300..599 to h[col].
// One-time (loop-invariant) kernel: fp32 FFMA register tile, 256 threads, output in the [32][136] tile layout.
// ---------------------------------------------------------------------------------------
constexpr int E2D_THREADS = 256;

__global__ void __launch_bounds__(E2D_THREADS, 1)
edge2d_emb_kernel(molsde_plan plan, const float* __restrict__ uv, const float* __restrict__ w3t,
                  const float* __restrict__ b3, float* __restrict__ e2d_tiles) {
    extern __shared__ __align__(128) float smem[];
    float* A = smem;             // [64][TE]
    float* W = smem + 64 * TE;   // [320][32] (rows >= 300 zero)
    __shared__ int s_src[TE], s_tgt[TE];
    const int tid = threadIdx.x, to = tid & 7, te = tid >> 3;
    for (int i = tid; i < 320 * 32; i += E2D_THREADS) W[i] = (i < 300 * 32) ? w3t[i] : 0.0f;
    __syncthreads();
    for (int tile = blockIdx.x; tile < plan.num_tiles; tile += gridDim.x) {
        const int ta = plan.tile_tgt_ptr[tile], tb = plan.tile_tgt_ptr[tile + 1];
        const int ea = plan.rowptr[ta], ne = plan.rowptr[tb] - ea;
        __syncthreads();
        for (int i = ta + tid; i < tb; i += E2D_THREADS)
            for (int e = plan.rowptr[i]; e < plan.rowptr[i + 1]; ++e) { s_tgt[e - ea] = i; s_src[e - ea] = plan.src[e]; }
        __syncthreads();
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.0f;
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
            const float* wk = W + k0 * 32;
#pragma unroll 4
            for (int k = 0; k < 64; ++k) {
                const float4 a = *reinterpret_cast<const float4*>(A + k * TE + te * 4);
                const float4 b = *reinterpret_cast<const float4*>(wk + k * 32 + to * 4);
                const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
            }
            __syncthreads();
        }
        float* out = e2d_tiles + static_cast<size_t>(tile) * TILE_FLOATS;
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
            *reinterpret_cast<float4*>(out + col * LDA + s) = o;
        }
        if (tid < 32 * (LDA - TE)) out[(tid / (LDA - TE)) * LDA + TE + tid % (LDA - TE)] = 0.0f;  // pad columns
    }
}

// ---------------------------------------------------------------------------------------
// edge_2D_emb in TRAIN mode (SDE_model_2D_to_3D.py:265,345-347): BatchNorm1d(300) uses the batch statistics of the
// E x 300 pre-activations  pre[e,f] = U[src_e,f] + V[tgt_e,f]  (uv = node-factored first layer incl. bias).
// One CTA per feature; fp64 two-pass mean / biased variance in a fixed order (deterministic); then the affine
// normalisation is folded into uv in place (U' = U*s + shift, V' = V*s) so that the eval-mode tile kernel applies, and
// the running statistics are updated (momentum 0.1, unbiased variance) like nn.BatchNorm1d.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
bn_edge_stats_kernel(const float* __restrict__ uv, const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, int64_t N,
                     int64_t E, int F, float* __restrict__ mean_out, float* __restrict__ var_out) {
    __shared__ double red[256];
    __shared__ double s_mean;
    const int f = blockIdx.x, tid = threadIdx.x;
    // pass 1: sum over edges = sum over targets i of (sum_{j in N(i)} U[j,f]) + deg(i) * V[i,f]
    double acc = 0.0;
    for (int64_t i = tid; i < N; i += 256) {
        const int a = rowptr[i], b = rowptr[i + 1];
        const double v = uv[i * 2 * F + F + f];
        for (int e = a; e < b; ++e) acc += static_cast<double>(uv[static_cast<int64_t>(src[e]) * 2 * F + f]) + v;
    }
    red[tid] = acc;
    __syncthreads();
    if (tid == 0) { double t = 0.0; for (int q = 0; q < 256; ++q) t += red[q]; s_mean = t / static_cast<double>(E); }
    __syncthreads();
    const double mu = s_mean;
    acc = 0.0;
    for (int64_t i = tid; i < N; i += 256) {
        const int a = rowptr[i], b = rowptr[i + 1];
        const double v = uv[i * 2 * F + F + f];
        for (int e = a; e < b; ++e) {
            const double d = static_cast<double>(uv[static_cast<int64_t>(src[e]) * 2 * F + f]) + v - mu;
            acc += d * d;
        }
    }
    red[tid] = acc;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int q = 0; q < 256; ++q) t += red[q];
        mean_out[f] = static_cast<float>(mu);
        var_out[f] = static_cast<float>(t / static_cast<double>(E));  // biased (normalisation)
    }
}

__global__ void bn_fold_uv_kernel(float* __restrict__ uv, int64_t N, int F, const float* __restrict__ mean,
                                  const float* __restrict__ var, const float* __restrict__ gamma, const float* __restrict__ beta,
                                  float eps) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * 2 * F) return;
    const int c = static_cast<int>(idx % (2 * F)), f = c % F;
    const float sc = gamma[f] / sqrtf(var[f] + eps);
    uv[idx] = (c < F) ? fmaf(uv[idx], sc, beta[f] - mean[f] * sc) : uv[idx] * sc;
}

__global__ void bn_running_update_kernel(const float* __restrict__ mean, const float* __restrict__ var, int F, int64_t E,
                                         float momentum, float* __restrict__ running_mean, float* __restrict__ running_var) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const float unbiased = var[f] * (static_cast<float>(E) / static_cast<float>(E > 1 ? E - 1 : 1));
    running_mean[f] = (1.0f - momentum) * running_mean[f] + momentum * mean[f];
    running_var[f] = (1.0f - momentum) * running_var[f] + momentum * unbiased;
}

// pos_perturbed = mean_coeff[i] * pos + std[i] * noise   (SDE_model_2D_to_3D.py:331-332; mean_coeff NULL = 1, VE)
__global__ void perturb_rows_kernel(const float* __restrict__ x, const float* __restrict__ mean_coeff, const float* __restrict__ stdv,
                                    const float* __restrict__ noise, int64_t N, int D, float* __restrict__ out) {
    const int64_t idx = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (idx >= N * D) return;
    const int64_t i = idx / D;
    const float m = mean_coeff ? __fmul_rn(mean_coeff[i], x[idx]) : x[idx];
    out[idx] = __fadd_rn(m, __fmul_rn(stdv[i], noise[idx]));
}

// loss_pos[g] = mean_{i in g} sum_xyz (score - noise)^2 * w[i]   (SDE_model_2D_to_3D.py:380-386), one warp per graph
__global__ void dsm_pos_loss_kernel(const float* __restrict__ score, const float* __restrict__ noise, const float* __restrict__ w,
                                    const int32_t* __restrict__ node_ptr, int B, float* __restrict__ out) {
    const int g = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (g >= B) return;
    const int a = node_ptr[g], b = node_ptr[g + 1];
    float acc = 0.0f;
    for (int i = a + lane; i < b; i += 32) {
        const float dx = score[3 * i] - noise[3 * i], dy = score[3 * i + 1] - noise[3 * i + 1], dz = score[3 * i + 2] - noise[3 * i + 2];
        acc += (dx * dx + dy * dy + dz * dz) * (w ? w[i] : 1.0f);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[g] = acc / static_cast<float>(max(b - a, 1));
}

// fixed-order mean of B floats (one CTA): loss_pos.mean()
__global__ void __launch_bounds__(256) mean_kernel(const float* __restrict__ v, int B, float* __restrict__ out) {
    __shared__ double red[256];
    double acc = 0.0;
    for (int i = threadIdx.x; i < B; i += 256) acc += v[i];
    red[threadIdx.x] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int q = 0; q < 256; ++q) t += red[q];
        out[0] = static_cast<float>(t / B);
    }
}

}  // namespace molsde

using namespace molsde;

static int plan_ok(const molsde_plan* p) {
    return p && p->num_chunks >= 0 && p->num_tiles >= 0 && p->chunk_tile_ptr && p->tile_tgt_ptr && p->rowptr &&
           (p->E == 0 || p->src);
}

extern "C" {

#ifdef MOLSDE_PROF
// debug: copy the per-CTA phase cycle counters (148 x 8 uint64) to the host and reset them
int molsde_debug_read_prof(unsigned long long* host_out) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(host_out, g_prof, sizeof(unsigned long long) * kNumSMs * 8);
    static unsigned long long zeros[kNumSMs * 8];
    cudaMemcpyToSymbol(g_prof, zeros, sizeof(zeros));
    return 0;
}
#endif

int64_t molsde_tile_floats(void) { return TILE_FLOATS; }

int molsde_edge2d_bn_train(const molsde_plan* plan, float* uv, int32_t F, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, float* batch_mean, float* batch_var,
                           void* stream) {
    if (!plan_ok(plan) || !uv || !gamma || !beta || !batch_mean || !batch_var || F <= 0 || plan->E <= 0) return MOLSDE_ERR_INVALID;
    bn_edge_stats_kernel<<<F, 256, 0, as_stream(stream)>>>(uv, plan->rowptr, plan->src, plan->N, plan->E, F, batch_mean, batch_var);
    int st = check_launch("bn_edge_stats");
    if (st != MOLSDE_OK) return st;
    const int64_t total = plan->N * 2 * F;
    bn_fold_uv_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, as_stream(stream)>>>(uv, plan->N, F, batch_mean, batch_var,
                                                                                            gamma, beta, eps);
    st = check_launch("bn_fold_uv");
    if (st != MOLSDE_OK) return st;
    if (running_mean && running_var) {
        bn_running_update_kernel<<<(F + 127) / 128, 128, 0, as_stream(stream)>>>(batch_mean, batch_var, F, plan->E, momentum,
                                                                               running_mean, running_var);
        st = check_launch("bn_running_update");
    }
    return st;
}

int molsde_perturb_rows(const float* x, const float* mean_coeff, const float* stdv, const float* noise, int64_t N, int32_t D,
                        float* out, void* stream) {
    if (!x || !stdv || !noise || !out || N < 0 || D <= 0) return MOLSDE_ERR_INVALID;
    if (N == 0) return MOLSDE_OK;
    perturb_rows_kernel<<<static_cast<unsigned>((N * D + 255) / 256), 256, 0, as_stream(stream)>>>(x, mean_coeff, stdv, noise, N, D, out);
    return check_launch("perturb_rows");
}

int molsde_dsm_pos_loss(const float* score, const float* noise, const float* w, const int32_t* node_ptr, int32_t B, float* out,
                        float* mean_out, void* stream) {
    if (!score || !noise || !node_ptr || !out || B <= 0) return MOLSDE_ERR_INVALID;
    dsm_pos_loss_kernel<<<(B + 3) / 4, 128, 0, as_stream(stream)>>>(score, noise, w, node_ptr, B, out);
    int st = check_launch("dsm_pos_loss");
    if (st != MOLSDE_OK || !mean_out) return st;
    mean_kernel<<<1, 256, 0, as_stream(stream)>>>(out, B, mean_out);
    return check_launch("dsm_pos_loss.mean");
}

int64_t molsde_sde2d3d_scratch_floats(const molsde_plan* plan, int32_t max_chunk_tiles, int32_t* num_ctas_out) {
    if (!plan || max_chunk_tiles < 0) return MOLSDE_ERR_INVALID;
    int ctas = plan->num_chunks < kNumSMs ? plan->num_chunks : kNumSMs;
    if (ctas < 1) ctas = 1;
    if (num_ctas_out) *num_ctas_out = ctas;
    return static_cast<int64_t>(ctas) * max_chunk_tiles * SCR_TILE;
}

int molsde_edge2d_emb_eval(const molsde_plan* plan, const float* uv, const float* w3t, const float* b3,
                           float* e2d_tiles, void* stream) {
    if (!plan_ok(plan) || !uv || !w3t || !b3 || !e2d_tiles) return MOLSDE_ERR_INVALID;
    if (plan->num_tiles == 0) return MOLSDE_OK;
    const size_t smem = sizeof(float) * (64 * TE + 320 * 32);
    cudaError_t err = cudaFuncSetAttribute(edge2d_emb_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    int grid = plan->num_tiles < 2 * kNumSMs ? plan->num_tiles : 2 * kNumSMs;
    edge2d_emb_kernel<<<grid, E2D_THREADS, smem, as_stream(stream)>>>(*plan, uv, w3t, b3, e2d_tiles);
    return check_launch("edge2d_emb");
}

static int launch_score(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr, const float* e2d_tiles,
                        const float* pos, const float* stdv, float* score, float* scratch, int64_t scratch_floats,
                        int32_t* status_flag, const float* attn_keep, const float* ffn_keep, float inv_keep, void* stream) {
    if (!plan_ok(plan) || !params || !params->blob || !nattr || !e2d_tiles || !pos || !score || !scratch)
        return MOLSDE_ERR_INVALID;
    if (params->blob_floats < MOLSDE_P_TOTAL) return MOLSDE_ERR_INVALID;
    if (plan->num_chunks == 0) return MOLSDE_OK;
    int ctas = plan->num_chunks < kNumSMs ? plan->num_chunks : kNumSMs;
    const int64_t stride = (scratch_floats / ctas) / SCR_TILE * SCR_TILE;
    if (stride < SCR_TILE) return MOLSDE_ERR_WORKSPACE;
    cudaError_t err = cudaFuncSetAttribute(sde2d3d_score_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    sde2d3d_score_kernel<<<ctas, NTHREADS, SMEM_BYTES, as_stream(stream)>>>(*plan, params->blob, nattr, e2d_tiles, pos,
                                                                          stdv, score, scratch, stride, status_flag,
                                                                          attn_keep, ffn_keep, inv_keep);
    return check_launch("sde2d3d_score");
}

int molsde_sde2d3d_score(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                         const float* e2d_tiles, const float* pos, const float* stdv, float* score, float* scratch,
                         int64_t scratch_floats, int32_t* status_flag, void* stream) {
    if (!stdv) return MOLSDE_ERR_INVALID;
    return launch_score(plan, params, nattr, e2d_tiles, pos, stdv, score, scratch, scratch_floats, status_flag, nullptr, nullptr,
                        1.0f, stream);
}

int molsde_sde2d3d_forward_net(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                               const float* e2d_tiles, const float* pos, const float* attn_keep, const float* ffn_keep,
                               float dropout_p, float* gradient, float* scratch, int64_t scratch_floats, int32_t* status_flag,
                               void* stream) {
    if ((attn_keep == nullptr) != (ffn_keep == nullptr) || dropout_p < 0.0f || dropout_p >= 1.0f) return MOLSDE_ERR_INVALID;
    return launch_score(plan, params, nattr, e2d_tiles, pos, nullptr, gradient, scratch, scratch_floats, status_flag, attn_keep,
                        ffn_keep, 1.0f / (1.0f - dropout_p), stream);
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
    const int64_t stride = (scratch_floats / ctas) / SCR_TILE * SCR_TILE;
    if (stride < SCR_TILE) return MOLSDE_ERR_WORKSPACE;
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
