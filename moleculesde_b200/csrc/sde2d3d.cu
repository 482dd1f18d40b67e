// SDEModel2Dto3D_02 score network (K3) and the fused predictor-corrector reverse-SDE loop (K5).
//
// Reference path: Geom3D/models/MoleculeSDE/SDE_model_2D_to_3D.py:393-445 (get_score),
// equivariant_scorenetwork.py:121-169, and the sampler in
// examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-212.
//
// Design (B200-first):
//   * one persistent CTA (512 threads, 1 CTA/SM, ~220 KB smem) owns one "chunk" = a set of whole molecules with <= 224 atoms;
//     node state (hidden features, q/k/v, positions, score) lives in shared memory for the WHOLE score evaluation -- and, in
//     the PC kernel, for all 1000 reverse steps -- so HBM sees only the initial/final positions;
//   * edges are processed in CSR-by-target order in tiles of <= 128 edges aligned to target nodes, so the segment softmax /
//     mean aggregation of a tile is self-contained and runs in a fixed, atomic-free, ascending-source order (deterministic,
//     same order as the reference scatter);
//   * the CTA is four decoupled QUADS of 128 threads.  A quad owns one edge tile at a time and THREAD = EDGE SLOT = TMEM LANE:
//     every per-edge GEMM (Fourier-feature layers, project.1, lin_edge, basis MLP layer 0) is a tcgen05.mma kind::f16
//     (M = 128 edges, fp32 accumulator in the quad's 128 TMEM columns) with an error-compensated two-way fp16 split of both
//     operands (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi: 22-bit operands, north_star's 1e-4 holds with 3 orders of margin), issued
//     by the quad's first thread and tracked by the quad's own mbarriers; epilogues read the accumulator row of their edge
//     with tcgen05.ld and keep all 32 (or 128) columns in registers, so bias / SiLU / attention logits / the 128 -> 3
//     projection need no cross-lane traffic.  Quads synchronise on named barriers only; four tiles are in flight per SM and
//     hide each other's MMA / TMA / MUFU latencies;
//   * the per-edge attribute is written ONCE per evaluation by the feature phase, already as the fp16 hi/lo A-operand tile
//     (canonical K-major core-matrix layout) into an L2-resident per-CTA scratch record; the four GAT layers and the two
//     basis modules fetch it with one TMA bulk copy per tile (cp.async.bulk + mbarrier complete_tx) straight into the
//     operand slot -- no thread touches it again;
//   * the per-node GEMMs run the same way with THREAD = NODE (<= 224 nodes = two M = 128 tiles on quads 0 and 1): one N = 128
//     GEMM gives q | k | v | lin_skip (the skip part stays in TMEM through the edge phase), FFN.0 / FFN.3 follow as N = 32
//     GEMMs; both LayerNorms and the residuals are thread-local on the accumulator row.  No legacy mma.sync is left.
#include <math_constants.h>

#include "common.cuh"
#include "mma_tile.cuh"   // split_f16x2
#include "sde2d3d_params.h"

namespace molsde {

constexpr int TE = MOLSDE_TILE_EDGES;         // 128 edges per tile
constexpr int NTHREADS = 512;
constexpr int NWARPS = NTHREADS / 32;
constexpr int QUADS = 4;                      // four 128-thread quads, one edge tile each
constexpr int QT = NTHREADS / QUADS;
constexpr int MAXN = MOLSDE_CHUNK_MAX_NODES;  // 224 atoms per chunk
constexpr int MAXT = 64;                      // tiles per chunk
constexpr int E2D_TILE_FLOATS = 32 * TE;      // edge_2D_emb output tile [8 feature quads][128 slots][4]
constexpr float EPS = 1e-6f;                  // SDE_model_2D_to_3D.py:10
constexpr float LN_EPS = 1e-5f;
static_assert(QT == TE, "thread = edge slot inside a quad");

// ---- per-tile scratch record (global memory, L2 resident), byte offsets ----
constexpr int REC_EA_HI = 0;                  // edge_attr, fp16 hi part: A-operand tile [128 rows][32 k] = [4 k-chunks][128][8 halves]
constexpr int REC_EA_LO = 8192;               // fp16 lo part
constexpr int REC_FRAME = 16384;              // fp32 [9][128]: coord_diff | coord_cross | coord_vertical per slot
constexpr int REC_SLOT = REC_FRAME + 9 * TE * 4;  // uint8 [2][128]: chunk-local source / target of every slot (static per chunk)
constexpr int REC_BYTES = REC_SLOT + 2 * TE;  // 21,248
constexpr int REC_FLOATS = REC_BYTES / 4;
static_assert(REC_BYTES % 128 == 0 && MAXN <= 255, "record alignment / byte-sized node indices");

// ---- shared memory carve-up (byte offsets) ----
constexpr int SB_XT = 0;                                // f32 [MAXN][32] node hidden, node-major rows, 16-byte granules XOR-swizzled
constexpr int SB_BIG = SB_XT + MAXN * 32 * 4;           // phase-dependent union
//   GAT layers
constexpr int SB_Q = SB_BIG;                            // f32 [MAXN][32] query (aggregate written in place), rows XOR-swizzled
constexpr int SB_K = SB_Q + MAXN * 32 * 4;
constexpr int SB_V = SB_K + MAXN * 32 * 4;
constexpr int SB_WP = SB_V + MAXN * 32 * 4;             // resident weights of the current GAT layer [G_WP_SZ floats]
constexpr int SB_R = SB_WP + MOLSDE_G_WP_SZ * 4;        // node phases: A operand of the two node tiles + q|k|v|skip weights; edge phase: per-quad slots
constexpr int NODE_A = 16384;                           // per node tile (quad 0 / 1): [128 nodes][32 k] fp16 hi (8 KB) | lo (8 KB)
constexpr int SB_WQ = SB_R + 2 * NODE_A;                // q|k|v|skip B tile [128 n][32 k] hi | lo (16 KB), dead once the edge phase starts
constexpr int GAT_SLOT = 16384 + 4096;                  // per quad: A operand hi|lo (later the message tile [128][32] f32) + logits [128][8]
constexpr int SB_BIG_END = SB_R + QUADS * GAT_SLOT;
//   feature phase (E0)
constexpr int SB_E0W = SB_BIG;                          // E0 section of the blob
constexpr int SB_E0A = SB_E0W + MOLSDE_P_E0_END * 4;    // per quad: two A slots [hi 8 KB | lo 8 KB]
constexpr int E0_SLOT = 16384;
//   basis phase
constexpr int SB_BW = SB_BIG;                           // basis section of the blob
constexpr int SB_BA = SB_BW + MOLSDE_P_BASIS_STRIDE * 4;  // per quad: A operand [128 x 64] hi (16 KB) | lo (16 KB)
constexpr int B_SLOT = 32768;
constexpr int SB_MIX = SB_BA + QUADS * B_SLOT;          // per quad f32 [3][128]
//   predictor / corrector update (between evaluations)
constexpr int SB_SCORE = SB_BIG;                        // f32 [MAXN*3]
constexpr int SB_NOISE = SB_SCORE + MAXN * 3 * 4;
// persistent tail
constexpr int SB_POS = SB_BIG_END;                      // f32 [MAXN*3]
constexpr int SB_GRAD = SB_POS + MAXN * 3 * 4;          // f32 [MAXN*3] network output ("gradient")
constexpr int SB_RED = SB_GRAD + MAXN * 3 * 4;          // f32 [64]
constexpr int SB_INT = SB_RED + 256;
constexpr int SI_ROWL = 0;                              // [MAXN+1] edge offsets local to the chunk
constexpr int SI_TTGT = SI_ROWL + MAXN + 1;             // [MAXT+1] tile target boundaries local to the chunk
constexpr int SI_MISC = SI_TTGT + MAXT + 1;             // [0] work item, [1] TMEM base, [2] phase of the CTA-wide TMA barrier
constexpr int SB_BAR = SB_INT + ((SI_MISC + 6) * 4 + 127) / 128 * 128;  // mbarriers (8 B): quad MMA [q][2], quad TMA [q], CTA TMA
constexpr int NUM_BARS = 5 * QUADS + 1;                        //   ... quad FULL [q][2] (count = QT arrivals) behind them
constexpr int BAR_CTA_TMA = 3 * QUADS, BAR_FULL0 = 3 * QUADS + 1;
constexpr size_t SMEM_BYTES = SB_BAR + NUM_BARS * 8;
static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB per-CTA shared memory of sm_100");
static_assert(SB_BIG % 128 == 0 && SB_WP % 128 == 0 && SB_R % 128 == 0 && SB_E0A % 128 == 0 && SB_BA % 128 == 0 && SB_BAR % 8 == 0,
              "operand tiles / barriers alignment");
static_assert(SB_E0A + QUADS * 2 * E0_SLOT <= SB_BIG_END && SB_MIX + QUADS * 3 * TE * 4 <= SB_BIG_END, "phase unions fit");
static_assert(QUADS * GAT_SLOT >= 2 * NODE_A + (MOLSDE_P_GAT_SZ - MOLSDE_G_WQKVS) * 4 && MAXN <= 2 * QT, "R region uses / two node tiles");
static_assert((MOLSDE_G_WEC * 4) % 128 == 0 && (MOLSDE_P_E0_BT * 4) % 128 == 0, "B tiles are 128B aligned inside their sections");
constexpr uint32_t TMEM_COLS = 512;  // 128 accumulator columns per quad

// optional per-phase cycle accounting (build with MOLSDE_PROF=1; read back with molsde_debug_read_prof)
#ifdef MOLSDE_PROF
__device__ unsigned long long g_prof[kNumSMs][8];
#define PROF_T0() long long prof_t0 = clock64()
#define PROF_ADD(slot) do { __syncthreads(); if (threadIdx.x == 0) { long long prof_t1 = clock64(); g_prof[blockIdx.x][slot] += prof_t1 - prof_t0; prof_t0 = prof_t1; } } while (0)
#else
#define PROF_T0()
#define PROF_ADD(slot)
#endif

// The quad's MMA / TMA issuing thread sits in warp (4q + q): warps map to the four SM sub-partitions by warp index mod 4, so the
// four leaders (and their single-thread issue sequences and barrier polls) run on four different schedulers.
__device__ __forceinline__ bool quad_leader(int q, int e) { return e == (q << 5); }

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
    int32_t* status_flag;
};
__device__ __forceinline__ uint8_t* smem_bytes(const Chunk& c) { return reinterpret_cast<uint8_t*>(c.sm); }
template <typename T>
__device__ __forceinline__ T* smem_at(const Chunk& c, int byte_off) { return reinterpret_cast<T*>(smem_bytes(c) + byte_off); }

// Per-thread synchronisation state, carried through the phases in one register:
//   bit 0 / 1: phase parity of the quad's two MMA barriers, bit 2: of the quad's TMA barrier, bit 3: a wait timed out (sticky),
//   bit 4 / 5: phase parity of the quad's two FULL barriers (operand rows written by all 128 threads; only the leader waits).
//   One FULL barrier per operand slot: a thread may run one ring use ahead of the slowest warp of its quad, never two (re-using a
//   slot needs the commit of its previous MMAs, which the leader issues only after all 128 arrivals), so arrivals of
//   consecutive uses must not share a barrier.
constexpr uint32_t QS_MMA0 = 1u, QS_MMA1 = 2u, QS_TMA = 4u, QS_DEAD = 8u, QS_FULL0 = 16u, QS_FULL1 = 32u;

// ---------------------------------------------------------------------------------------
// fast, accuracy-checked elementwise helpers (absolute / relative error ~1e-6, far below the 1e-4 bar)
// ---------------------------------------------------------------------------------------
// (Measured and rejected: the reciprocal on the FMA pipe -- integer-trick seed + three Newton steps, ONE SFU instruction per SiLU
// instead of two -- is 6 % SLOWER on the whole PC pass, 20.6 vs 19.35 ms on the 296-group probe: the SiLU bursts are bound by issue
// slots, 11 instructions against 5, not by the 16 SFU lanes.)
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.0f + __expf(-x)); }

// sin/cos of an fp32 argument of magnitude up to ~1e5: 2-term Cody-Waite reduction by 2*pi (exact products through FMA; the
// rounding to the nearest multiple uses the 1.5*2^23 magic constant, i.e. two FADDs instead of a conversion-pipe FRND), then
// the SFU approximations on |r| <= pi (abs error < 1e-6).
__device__ __forceinline__ void sincos_reduced(float x, float& s, float& c) {
    const float k = __fsub_rn(__fadd_rn(x * 0.15915494309189535f, 12582912.0f), 12582912.0f);
    float r = fmaf(-k, 6.2831854820251465f, x);
    r = fmaf(-k, -1.7484555314695172e-7f, r);
    s = __sinf(r);
    c = __cosf(r);
}

// sin / cos of 2*pi*u for a phase u given in TURNS (GaussianFourierProjection: u = x * W): the integer part is removed exactly
// (u - rint(u) is exact in fp32), the remaining |f| <= 0.5 is scaled to radians once and fed to the SFU.  5 FMA-pipe instructions
// instead of 8 for "form x*W*2*pi as the reference rounds it, then Cody-Waite by 2*pi".  Against the reference's argument
// rn(rn(rn(x W) 2) pi_f32) the phase differs by <= 0.5 ulp(arg) + 2.8e-8 arg (pi_f32 vs pi) -- 1e-6 rad at arg = 30, far below the
// 1e-4 bar; at the VE03 preset's args of ~1e4 rad both sit inside the stated 2e-3 (one ulp of the argument there is 1e-3 rad).
__device__ __forceinline__ void sincos_turns(float u, float& s, float& c) {
    const float k = __fsub_rn(__fadd_rn(u, 12582912.0f), 12582912.0f);   // rint(u), |u| < 2^22
    const float r = __fmul_rn(__fsub_rn(u, k), 6.2831854820251465f);
    s = __sinf(r);
    c = __cosf(r);
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

// ---- mbarrier / TMA bulk copy / tcgen05 primitives --------------------------------------------
__device__ __forceinline__ uint32_t bar_addr(const Chunk& c, int idx) { return smem_u32(smem_bytes(c) + SB_BAR + idx * 8); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// Bounded parity wait: a descriptor / protocol bug must not hang the GPU.  After the first timeout the thread's state word
// carries QS_DEAD, every later wait returns at once (the launch finishes quickly with garbage) and the status word is set.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t& qs, uint32_t which, int32_t* status_flag) {
    if (!(qs & QS_DEAD)) {
        const uint32_t parity = (qs & which) ? 1u : 0u;
        uint32_t done = 0;
        for (int it = 0; it < (1 << 16) && !done; ++it)   // each poll sleeps in hardware up to the hinted time (ns) before it returns
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                         : "=r"(done) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (!done) {
            qs |= QS_DEAD;
            if (status_flag) atomicExch(status_flag, -7);
        }
    }
    qs ^= which;
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void quad_sync(int q) { asm volatile("bar.sync %0, %1;" ::"r"(q + 1), "n"(QT) : "memory"); }

// Weight sections of a phase (13-46 KB each, contiguous in the parameter blob): TMA bulk copies (cp.async.bulk, 1-D) issued by
// thread 0 and tracked by the CTA-wide mbarrier (complete_tx).  Call with all threads after a __syncthreads() (the destinations
// may still be read by the previous phase before that); returns when the data is visible to every thread.
__device__ __forceinline__ void stage_bulk2(const Chunk& c, void* dst0, const float* __restrict__ src0, int nfloat0, void* dst1,
                                            const float* __restrict__ src1, int nfloat1) {
    const uint32_t bar = bar_addr(c, BAR_CTA_TMA);
    const uint32_t parity = static_cast<uint32_t>(c.si[SI_MISC + 2]);
    if (threadIdx.x == 0) {
        fence_proxy_async_smem();  // earlier generic accesses of the destinations before the async-proxy writes
        mbar_expect_tx(bar, static_cast<uint32_t>(nfloat0 + nfloat1) * 4u);
        bulk_g2s(smem_u32(dst0), src0, static_cast<uint32_t>(nfloat0) * 4u, bar);
        if (nfloat1 > 0) bulk_g2s(smem_u32(dst1), src1, static_cast<uint32_t>(nfloat1) * 4u, bar);
    }
    uint32_t done = 0;
    for (int it = 0; it < (1 << 22) && !done; ++it)
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}\n"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    __syncthreads();
    if (threadIdx.x == 0) c.si[SI_MISC + 2] = static_cast<int>(parity ^ 1u);  // read again only after several more barriers
}

// tcgen05 shared-memory matrix descriptor, K-major, no swizzle (cute/arch/mma_sm100_desc.hpp): core matrices of 8 rows x 16 B;
// LBO = byte distance between core matrices adjacent in K, SBO = 128 B between 8-row groups.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes) {
    return static_cast<uint64_t>((saddr & 0x3FFFF) >> 4)                   // start address  [0,14)
           | (static_cast<uint64_t>(lbo_bytes >> 4) << 16)                  // leading byte offset [16,30)
           | (static_cast<uint64_t>(128u >> 4) << 32)                       // stride byte offset  [32,46)
           | (static_cast<uint64_t>(1) << 46);                              // version 1 (Blackwell), layout = no swizzle
}
// D[tmem, 128 lanes x N columns] (+)= A[smem, 128 x 16] . B[smem, N x 16]^T, fp16 inputs, fp32 accumulate; issued by ONE thread
template <int N>
__device__ __forceinline__ void umma_f16_m128(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t accumulate) {
    // instruction descriptor: D = F32 (1<<4), A = B = F16 (0<<7, 0<<10), both K-major, N>>3 at [17,23), M>>4 at [24,29)
    constexpr uint32_t idesc = (1u << 4) | ((static_cast<uint32_t>(N) >> 3) << 17) | ((128u >> 4) << 24);
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Split-fp16 GEMM: three passes (lo*hi, hi*lo, hi*hi; small terms first) over `ksteps` K=16 steps of an A tile [128 x 16*ksteps]
// (hi at a_hi, lo at a_lo; k-chunk stride 2048 B) and a B tile [N x 16*ksteps] (k-chunk stride N*16 B).
template <int N, int KSTEPS>
__device__ __forceinline__ void umma_split_f16(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                               uint32_t accumulate) {
    constexpr uint32_t LBO_A = 2048, LBO_B = N * 16;
    const uint64_t dah = umma_desc(a_hi, LBO_A), dal = umma_desc(a_lo, LBO_A), dbh = umma_desc(b_hi, LBO_B), dbl = umma_desc(b_lo, LBO_B);
#pragma unroll
    for (int term = 0; term < 3; ++term) {
        const uint64_t da = (term == 0) ? dal : dah, db = (term == 1) ? dbl : dbh;
#pragma unroll
        for (int kb = 0; kb < KSTEPS; ++kb) {   // the start-address field counts 16-byte units: one K step = two k-chunks
            umma_f16_m128<N>(tmem_d, da + static_cast<uint64_t>((kb * 2 * LBO_A) >> 4), db + static_cast<uint64_t>((kb * 2 * LBO_B) >> 4),
                             accumulate);
            accumulate = 1;
        }
    }
}
// accumulator columns [col, col+32) / [col, col+8) of this thread's TMEM lane (all 32 lanes of the warp must call)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&f)[32]) {
    uint32_t v[32];
    __syncwarp();  // .sync.aligned: the quad leader's lane may still be behind after its single-thread MMA issue block
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&f)[8]) {
    uint32_t r[8];
    __syncwarp();
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];\n\t"
                 "tcgen05.wait::ld.sync.aligned;\n"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(r[i]);
}
// 8 consecutive-k values of row r -> fp16 hi / lo, one 16 B chunk each, into an operand tile (k-chunk stride 2048 B)
__device__ __forceinline__ void split8(const float* v, uint4& h, uint4& l) {
    split_f16x2(v[0], v[1], h.x, l.x);
    split_f16x2(v[2], v[3], h.y, l.y);
    split_f16x2(v[4], v[5], h.z, l.z);
    split_f16x2(v[6], v[7], h.w, l.w);
}
__device__ __forceinline__ void store_a_chunk(uint8_t* hi, uint8_t* lo, int r, int kc, const float* v) {
    uint4 h, l;
    split8(v, h, l);
    *reinterpret_cast<uint4*>(hi + kc * 2048 + r * 16) = h;
    *reinterpret_cast<uint4*>(lo + kc * 2048 + r * 16) = l;
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

// q / k / v rows are [node][32] with the 16-byte granules XOR-swizzled by the node index: the per-edge row gathers
// k[src], v[src] of 8 consecutive slots then hit 8 different bank groups instead of one.
__device__ __forceinline__ int qkv_idx(int node, int col) { return node * 32 + (col ^ ((node & 7) << 2)); }

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

// ---------------------------------------------------------------------------------------
// Phase E0: per-edge attribute  edge_attr = input_mlp(gfp(d)) * e2d + project([sin,cos,emb_i,emb_j])
// SDE_model_2D_to_3D.py:402-432.  coff_mlp (a bare Linear) is folded into project.layers.0 on the host, so the hidden layer
// accumulates directly over the four Fourier blocks.
//   Per tile (one quad): every thread builds the 5 x 64 Fourier features of ITS edge in ten K = 32 sub-blocks
//   [sin f(16h..) | cos f(16h..)] and writes them as fp16 hi/lo rows into a two-slot operand ring; the quad leader issues
//   3 x 2 tcgen05.mma (N = 32) per ring use into the TMEM accumulators "inv" (block 0) and "hidden" (blocks 1..4).
//   Nothing blocks on the way: threads ARRIVE on the quad's FULL barrier and go on computing the next sub-block; only the
//   leader waits for the 128 arrivals, and a slot is re-written only after the commit of its previous MMAs (mbarrier).
//   The epilogue of a tile is software-pipelined INTO the next tile of the quad (accumulators double-buffered by tile parity):
//   after sub-block 1 of tile n+1 the ring protocol guarantees that all Fourier MMAs of tile n are complete -> hidden -> +bias,
//   SiLU -> A operand -> project.1 as one more ring use (its accumulator overwrites "hidden"); after sub-block 3 that MMA is
//   complete -> edge_attr = (inv + b) * e2d + (proj + b) -> scratch record, already as the fp16 hi/lo operand tile of the
//   later phases.  The MMA latency is never exposed except when the quad drains its last tile.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t phase_edge_features(const Chunk c, const float* __restrict__ blob, const float* __restrict__ e2d_tiles,
                                                     uint8_t* __restrict__ scratch, uint32_t tmem_base, uint32_t qs) {
    const float* W = smem_at<const float>(c, SB_E0W);
    const float* pos = smem_at<const float>(c, SB_POS);
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1);
    uint8_t* Aq = smem_bytes(c) + SB_E0A + q * (2 * E0_SLOT);
    const uint32_t a_base = smem_u32(Aq);
    const uint32_t w_bt = smem_u32(W + MOLSDE_P_E0_BT);
    const uint32_t bar0 = bar_addr(c, 2 * q), bar1 = bar_addr(c, 2 * q + 1);
    const uint32_t full0 = bar_addr(c, BAR_FULL0 + 2 * q), full1 = bar_addr(c, BAR_FULL0 + 2 * q + 1);
    const uint32_t tq = tmem_base + q * 128;                                  // the quad's accumulator columns
    const uint32_t lane_off = static_cast<uint32_t>(e & ~31) << 16;           // this warp's TMEM lane quarter
    __syncthreads();  // the union region may still be read by the tail of the previous evaluation / the PC update
    stage_bulk2(c, smem_bytes(c) + SB_E0W, blob, MOLSDE_P_E0_END, nullptr, nullptr, 0);

    uint32_t ring = 0;     // operand-ring uses so far (slot = ring & 1)
    uint32_t pend = 0;     // bit s: the last commit on slot s has not been waited for yet
    // one ring use: 32 operand columns of this thread's row -> slot, arrive; the leader issues the split-fp16 MMA group
    auto ring_use = [&](const float* v, uint32_t tmem_d, uint32_t w_tile, uint32_t accumulate) {
        const uint32_t sl = ring & 1u;
        const uint32_t bar = sl ? bar1 : bar0;
        if (pend & (1u << sl)) mbar_wait(bar, qs, sl ? QS_MMA1 : QS_MMA0, c.status_flag);  // MMAs of use ring-2 done: slot free
        uint8_t* Ah = Aq + sl * E0_SLOT;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) store_a_chunk(Ah, Ah + 8192, e, kc, v + 8 * kc);
        fence_proxy_async_smem();  // generic-proxy writes -> visible to the tensor core
        const uint32_t full = sl ? full1 : full0, fbit = sl ? QS_FULL1 : QS_FULL0;
        mbar_arrive(full);
        if (quad_leader(q, e)) {
            uint32_t qt = qs;
            mbar_wait(full, qt, fbit, c.status_flag);
            qs |= (qt & QS_DEAD);
            tc_fence_after();
            const uint32_t ah = a_base + sl * E0_SLOT;
            umma_split_f16<32, 2>(tmem_d, ah, ah + 8192, w_tile, w_tile + 2048, accumulate);
            umma_commit(bar);
        }
        qs ^= fbit;
        pend |= (1u << sl);
        ++ring;
    };
    auto drain = [&]() {   // wait for every outstanding commit (older slot first)
        const uint32_t older = ring & 1u;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
            const uint32_t sl = older ^ static_cast<uint32_t>(k);
            if (pend & (1u << sl)) mbar_wait(sl ? bar1 : bar0, qs, sl ? QS_MMA1 : QS_MMA0, c.status_flag);
        }
        pend = 0;
    };
    // hidden(tile) -> +bias, SiLU -> operand ring -> project.1 into the columns of "hidden"
    auto epilogue1 = [&](uint32_t tcol, float psin, float pcos) {
        tc_fence_after();
        float hv[32];
        tmem_ld32(tq + lane_off + tcol + 32, hv);
        tc_fence_before();
        const float4* HV = reinterpret_cast<const float4*>(W + MOLSDE_P_E0_HV);
#pragma unroll
        for (int col = 0; col < 32; ++col) {
            const float4 p = HV[col];  // {bias, w_sin, w_cos, 0}
            hv[col] = silu_fast(hv[col] + p.x + psin * p.y + pcos * p.z);
        }
        ring_use(hv, tq + tcol + 32, w_bt + 10 * (MOLSDE_E0_BT_FLOATS * 4), 0u);
    };
    // edge_attr = inv3d * e2d + frame  (:432) -> scratch record, as the fp16 hi/lo operand rows of the later phases
    auto epilogue2 = [&](int t, uint32_t tcol) {
        uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES;
        // edge_2D_emb tile of this edge (loop invariant, L2): [8 feature quads][128 slots][4]
        float4 e2[8];
        {
            const float4* e2d_t = reinterpret_cast<const float4*>(e2d_tiles + static_cast<size_t>(c.tile0 + t) * E2D_TILE_FLOATS) + e;
#pragma unroll
            for (int fq = 0; fq < 8; ++fq) e2[fq] = __ldg(e2d_t + fq * TE);
        }
        tc_fence_after();
        const float2* OB = reinterpret_cast<const float2*>(W + MOLSDE_P_E0_OB);
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            float inv[8], pr[8], ea[8];
            tmem_ld8(tq + lane_off + tcol + 8 * kc, inv);
            tmem_ld8(tq + lane_off + tcol + 32 + 8 * kc, pr);
            const float ev[8] = {e2[2 * kc].x, e2[2 * kc].y, e2[2 * kc].z, e2[2 * kc].w,
                                 e2[2 * kc + 1].x, e2[2 * kc + 1].y, e2[2 * kc + 1].z, e2[2 * kc + 1].w};
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 ob = OB[8 * kc + j];  // {input_mlp bias, project.1 bias}
                ea[j] = fmaf(inv[j] + ob.x, ev[j], pr[j] + ob.y);
            }
            uint4 h, l;
            split8(ea, h, l);
            *reinterpret_cast<uint4*>(rec + REC_EA_HI + kc * 2048 + e * 16) = h;
            *reinterpret_cast<uint4*>(rec + REC_EA_LO + kc * 2048 + e * 16) = l;
        }
        tc_fence_before();  // these TMEM reads are ordered before the next MMAs into the same columns by the FULL barrier
    };

    int tp = -1;                      // previous tile of this quad (its epilogue is pipelined into the current one)
    uint32_t tcol_p = 0;
    float psin_p = 0.f, pcos_p = 0.f;
    uint32_t par = 0;
    for (int t = q; t < c.ntiles; t += QUADS, par ^= 1u) {
        const TileInfo ti = tile_info(c, t);
        uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES;
        const bool live = e < ti.ne;
        const uint32_t tcol = par * 64;   // accumulators of this tile: inv = [tcol, +32), hidden / proj = [tcol + 32, +32)
        float x[5] = {0.f, 0.f, 0.f, 0.f, 0.f}, psin = 0.f, pcos = 0.f;
        if (live) {
            const int sj = rec[REC_SLOT + e], tg = rec[REC_SLOT + TE + e];
            const float* pr = pos + 3 * sj;  // row = source j
            const float* pc = pos + 3 * tg;  // col = target i
            const Frame f = coord2basis(pr, pc);
            {   // cache the equivariant basis of this edge for the two basis phases (equivariant_scorenetwork.py:159)
                float* fr = reinterpret_cast<float*>(rec + REC_FRAME) + e;
                fr[0 * TE] = f.dx; fr[1 * TE] = f.dy; fr[2 * TE] = f.dz;
                fr[3 * TE] = f.cx; fr[4 * TE] = f.cy; fr[5 * TE] = f.cz;
                fr[6 * TE] = f.vx; fr[7 * TE] = f.vy; fr[8 * TE] = f.vz;
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
            pcos = __fdiv_rn(__fdiv_rn(dot3_rn(ci0, ci1, ci2, cj0, cj1, cj2), __fadd_rn(ni, EPS)), __fadd_rn(nj, EPS));
            // :425  sqrt(1 - cos^2).  For (anti)parallel r_i, r_j rounding can make the argument a tiny negative
            // number and the reference then returns NaN for the whole batch; the limit value 0 is used instead
            // (only inputs on which the reference output is NaN are affected).
            psin = sqrtf(fmaxf(__fsub_rn(1.0f, __fmul_rn(pcos, pcos)), 0.0f));
            x[0] = f.dist; x[1] = ci0; x[2] = ci2; x[3] = cj0; x[4] = cj2;
        }
        // Fourier block 0 (distance) feeds input_mlp (:409-410); blocks 1..4 (ci0, ci2, cj0, cj2) feed the fused hidden layer
        // of `project` (:427-430).
        // (one rolled loop over the ten sub-blocks: the body is ~400 instructions and four warps per scheduler sit at different
        //  places of it -- a fully unrolled copy does not fit the instruction caches)
#pragma unroll 1
        for (int b = 0; b < 10; ++b) {
            const int blk = b >> 1, half = b & 1;
            const float xb = blk == 0 ? x[0] : blk == 1 ? x[1] : blk == 2 ? x[2] : blk == 3 ? x[3] : x[4];
            const float4* Wf = reinterpret_cast<const float4*>(W + (blk == 0 ? MOLSDE_P_GFP_DIST_W : MOLSDE_P_GFP_COFF_W) + half * 16);
            float v[32];
#pragma unroll
            for (int i4 = 0; i4 < 4; ++i4) {
                const float4 w4 = Wf[i4];
                const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    // GaussianFourierProjection.forward, :64-66: sin / cos of x * W * 2 * pi, evaluated from the phase in turns
#ifdef MOLSDE_SINCOS_REF_ROUNDING
                    sincos_reduced(__fmul_rn(__fmul_rn(__fmul_rn(xb, wv[j]), 2.0f), 3.14159274101257324f), v[4 * i4 + j], v[16 + 4 * i4 + j]);
#else
                    sincos_turns(__fmul_rn(xb, wv[j]), v[4 * i4 + j], v[16 + 4 * i4 + j]);
#endif
                }
            }
            ring_use(v, tq + tcol + (blk == 0 ? 0 : 32), w_bt + b * (MOLSDE_E0_BT_FLOATS * 4), (b == 0 || b == 2) ? 0u : 1u);
            // pipelined epilogue of the quad's previous tile: the slot wait inside the ring use above has just proven that
            //   after sub-block 1: every Fourier MMA of the previous tile is complete (its last commit was two uses ago)
            //   after sub-block 3: its project.1 MMA (issued after sub-block 1) is complete
            if (b == 1 && tp >= 0) epilogue1(tcol_p, psin_p, pcos_p);
            if (b == 3 && tp >= 0) epilogue2(tp, tcol_p);
        }
        tp = t; tcol_p = tcol; psin_p = psin; pcos_p = pcos;
    }
    if (tp >= 0) {   // the quad's last tile: drain the ring, then both epilogue stages with their MMA latency exposed once
        drain();
        epilogue1(tcol_p, psin_p, pcos_p);
        drain();
        epilogue2(tp, tcol_p);
    }
    asm volatile("fence.proxy.async;" ::: "memory");  // the records are read back by TMA bulk copies (async proxy)
    __syncthreads();
    return qs;
}

// ---------------------------------------------------------------------------------------
// GAT layer pieces  (equivariant_scorenetwork.py:34-40, TransformerConv heads=8 C=4)
// ---------------------------------------------------------------------------------------
// node rows are [node][32] fp32 with the 16-byte granules XOR-swizzled by the node index (same scheme as q / k / v)
__device__ __forceinline__ float4* row_granule(float* base, int node, int g) {
    return reinterpret_cast<float4*>(base + node * 32) + (g ^ (node & 7));
}
// fp16 hi / lo operand row of a node tile from 32 fp32 values (4 k-chunks)
__device__ __forceinline__ void store_node_operand(uint8_t* A, int e, const float* v) {
#pragma unroll
    for (int kc = 0; kc < 4; ++kc) store_a_chunk(A, A + 8192, e, kc, v + 8 * kc);
}

// x of every node (thread = node; node tile m on quad m) -> XT row (from `src_rows` [N][32] in global memory when given) and the
// fp16 hi/lo A operand of the layer's first GEMM.  Called after E0 (from nattr) and after a basis phase (from XT), whose
// operand buffers overlay the node operand tiles.
__device__ __forceinline__ void node_stage(const Chunk c, const float* __restrict__ src_rows) {
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1), node = q * QT + e;
    if (q * QT >= c.n) return;
    float* XT = smem_at<float>(c, SB_XT);
    float x[32];
    if (node < c.n) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            float4 v4;
            if (src_rows) {
                v4 = __ldg(reinterpret_cast<const float4*>(src_rows + static_cast<size_t>(c.node0 + node) * 32) + g);
                *row_granule(XT, node, g) = v4;
            } else {
                v4 = *row_granule(XT, node, g);
            }
            x[4 * g] = v4.x; x[4 * g + 1] = v4.y; x[4 * g + 2] = v4.z; x[4 * g + 3] = v4.w;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) x[i] = 0.0f;
    }
    store_node_operand(smem_bytes(c) + SB_R + q * NODE_A, e, x);
    fence_proxy_async_smem();
}

// q | k | v | skip = Linear(x) as ONE N = 128 GEMM per node tile (A operand staged by node_stage / the previous node_update);
// q, k, v (+bias) go to their swizzled rows, the lin_skip part stays in TMEM columns [96, 128) for node_update.
__device__ __noinline__ uint32_t node_qkvs(const Chunk c, uint32_t tmem_base, uint32_t qs) {
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1), node = q * QT + e;
    if (q * QT >= c.n) return qs;  // (whole quad: no node tile)
    const float* Wp = smem_at<const float>(c, SB_WP);
    float* Qb = smem_at<float>(c, SB_Q);
    const uint32_t a_addr = smem_u32(smem_bytes(c) + SB_R + q * NODE_A), w_addr = smem_u32(smem_bytes(c) + SB_WQ);
    const uint32_t bar_mma = bar_addr(c, 2 * q);
    const uint32_t tq = tmem_base + q * 128;
    const uint32_t tlane = tq + (static_cast<uint32_t>(e & ~31) << 16);
    if (quad_leader(q, e)) {
        tc_fence_after();
        umma_split_f16<128, 2>(tq, a_addr, a_addr + 8192, w_addr, w_addr + 8192, 0u);
        umma_commit(bar_mma);
    }
    mbar_wait(bar_mma, qs, QS_MMA0, c.status_flag);
    tc_fence_after();
#pragma unroll 1
    for (int part = 0; part < 3; ++part) {   // q, k, v
        float v[32];
        tmem_ld32(tlane + 32 * part, v);
        if (node < c.n) {
            const float4* b4 = reinterpret_cast<const float4*>(Wp + MOLSDE_G_BQKVS + 32 * part);
            float* dst = Qb + part * (32 * MAXN);
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 b = b4[g];
                *row_granule(dst, node, g) = make_float4(v[4 * g] + b.x, v[4 * g + 1] + b.y, v[4 * g + 2] + b.z, v[4 * g + 3] + b.w);
            }
        }
    }
    tc_fence_before();
    return qs;
}

// attention over the incoming edges of every target: e = lin_edge(edge_attr) on the tensor core (A operand = the scratch record,
// fetched by one TMA bulk copy), logits, segment softmax (+1e-16), weighted messages, deterministic ascending-source sum; the
// aggregate overwrites q[target].  One quad per tile, thread = edge slot; two quad barriers per tile: the edge threads publish
// their logits and unweighted messages (v_j + e), then one thread per (target, head) runs max / exp / weighted sum over the
// target's contiguous segment and normalises once:  sum_s ex_s m_s / (sum_s ex_s + 1e-16).
__device__ __noinline__ uint32_t gat_edge_phase(const Chunk c, const uint8_t* __restrict__ scratch, int layer, uint32_t tmem_base,
                                                uint32_t qs) {
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1);
    uint8_t* slot = smem_bytes(c) + SB_R + q * GAT_SLOT;
    float* Mm = reinterpret_cast<float*>(slot);          // [TE][32] messages (granules swizzled by the slot), over the consumed operand
    float* L = reinterpret_cast<float*>(slot + 16384);   // [TE][8] logits
    float* Q = smem_at<float>(c, SB_Q);
    const float* Kk = smem_at<const float>(c, SB_K);
    const float* V = smem_at<const float>(c, SB_V);
    const int* rowl = c.si + SI_ROWL;
    const uint32_t a_addr = smem_u32(slot);
    const uint32_t w_addr = smem_u32(smem_at<float>(c, SB_WP) + MOLSDE_G_WEC);
    const uint32_t bar_mma = bar_addr(c, 2 * q), bar_tma = bar_addr(c, 2 * QUADS + q);
    const uint32_t tq = tmem_base + q * 128;
    const uint32_t tlane = tq + (static_cast<uint32_t>(e & ~31) << 16);
    for (int t = q; t < c.ntiles; t += QUADS) {
        const TileInfo ti = tile_info(c, t);
        const uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES;
        if (quad_leader(q, e)) {
            fence_proxy_async_smem();  // the slot was last touched through the generic proxy (message tile of the previous tile)
            mbar_expect_tx(bar_tma, 16384u);
            bulk_g2s(a_addr, rec + REC_EA_HI, 16384u, bar_tma);
        }
        const bool live = e < ti.ne;
        const int sj = live ? rec[REC_SLOT + e] : 0, tg = live ? rec[REC_SLOT + TE + e] : ti.ta;
        // q / k rows of this edge: in flight while the operand lands and the MMA runs
        float4 q4[8], k4[8];
        {
            const float4* Q4 = reinterpret_cast<const float4*>(Q) + tg * 8;
            const float4* K4 = reinterpret_cast<const float4*>(Kk) + sj * 8;
            const int sq = tg & 7, sk = sj & 7;
#pragma unroll
            for (int hd = 0; hd < 8; ++hd) { q4[hd] = Q4[hd ^ sq]; k4[hd] = K4[hd ^ sk]; }
        }
        if (quad_leader(q, e)) {
            uint32_t qt = qs;  // (every thread flips its own copy below)
            mbar_wait(bar_tma, qt, QS_TMA, c.status_flag);
            qs |= (qt & QS_DEAD);
            tc_fence_after();
            umma_split_f16<32, 2>(tq, a_addr, a_addr + 8192, w_addr, w_addr + 2048, 0u);
            umma_commit(bar_mma);
        }
        qs ^= QS_TMA;
        mbar_wait(bar_mma, qs, QS_MMA0, c.status_flag);
        tc_fence_after();
        float ev[32];
        tmem_ld32(tlane, ev);
        tc_fence_before();
        if (live) {
            const float4* V4 = reinterpret_cast<const float4*>(V) + sj * 8;
            const int sk = sj & 7;
            float lg[8];
#pragma unroll
            for (int hd = 0; hd < 8; ++hd) {
                const float4 v4 = V4[hd ^ sk];
                // alpha = (q_i . (k_j + e)) / sqrt(C)
                float part = fmaf(q4[hd].y, k4[hd].y + ev[4 * hd + 1], q4[hd].x * (k4[hd].x + ev[4 * hd]));
                part += fmaf(q4[hd].w, k4[hd].w + ev[4 * hd + 3], q4[hd].z * (k4[hd].z + ev[4 * hd + 2]));
                lg[hd] = part * 0.5f;
                // message before weighting: v_j + e  (the operand slot is free: its MMA completed)
                *reinterpret_cast<float4*>(Mm + e * 32 + ((hd ^ (e & 7)) << 2)) =
                    make_float4(ev[4 * hd] + v4.x, ev[4 * hd + 1] + v4.y, ev[4 * hd + 2] + v4.z, ev[4 * hd + 3] + v4.w);
            }
            *reinterpret_cast<float4*>(L + e * 8) = make_float4(lg[0], lg[1], lg[2], lg[3]);
            *reinterpret_cast<float4*>(L + e * 8 + 4) = make_float4(lg[4], lg[5], lg[6], lg[7]);
        }
        quad_sync(q);
        // per (target, head): softmax over the target's contiguous edge segment fused with the weighted message sum
        const int ntg = ti.tb - ti.ta;
        for (int p = e; p < ntg * 8; p += QT) {
            const int i = ti.ta + (p >> 3), hd = p & 7;
            const int s0 = rowl[i] - ti.ea, s1 = rowl[i + 1] - ti.ea;
            float m = -CUDART_INF_F;
            for (int s = s0; s < s1; ++s) m = fmaxf(m, L[s * 8 + hd]);
            float z = 0.0f;
            float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
            const float* kp = c.attn_keep ? c.attn_keep + (static_cast<size_t>(layer) * c.E_total + c.edge0 + ti.ea) * 8 + hd : nullptr;
            for (int s = s0; s < s1; ++s) {
                float ex = __expf(L[s * 8 + hd] - m);
                z += ex;
                if (kp) ex *= kp[s * 8];  // F.dropout(alpha, p) in train mode (TransformerConv.message): 0/1 keep mask, 1/(1-p) below
                const float4 mv = *reinterpret_cast<const float4*>(Mm + s * 32 + ((hd ^ (s & 7)) << 2));
                acc.x = fmaf(ex, mv.x, acc.x); acc.y = fmaf(ex, mv.y, acc.y);
                acc.z = fmaf(ex, mv.z, acc.z); acc.w = fmaf(ex, mv.w, acc.w);
            }
            const float rz = __fdividef(c.inv_keep, z + 1e-16f);   // (inv_keep = 1 in eval mode)
            *reinterpret_cast<float4*>(Q + qkv_idx(i, 4 * hd)) = make_float4(acc.x * rz, acc.y * rz, acc.z * rz, acc.w * rz);
        }
        quad_sync(q);  // the slot (messages) and the logits are rewritten by the next tile
    }
    return qs;
}

// LayerNorm over the 32 features of one node, all in this thread's registers (biased variance, eps 1e-5)
__device__ __forceinline__ void layer_norm_row(float (&v)[32], const float* __restrict__ w, const float* __restrict__ b) {
    float s4[4] = {0.f, 0.f, 0.f, 0.f};   // four independent partial sums: the node phases are latency-bound (one or two quads)
#pragma unroll
    for (int i = 0; i < 32; ++i) s4[i & 3] += v[i];
    const float mean = ((s4[0] + s4[1]) + (s4[2] + s4[3])) * (1.0f / 32.0f);
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; q4[i & 3] = fmaf(d, d, q4[i & 3]); }
    const float rstd = 1.0f / sqrtf(((q4[0] + q4[1]) + (q4[2] + q4[3])) * (1.0f / 32.0f) + LN_EPS);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 w4 = reinterpret_cast<const float4*>(w)[g], b4 = reinterpret_cast<const float4*>(b)[g];
        v[4 * g] = (v[4 * g] - mean) * rstd * w4.x + b4.x;
        v[4 * g + 1] = (v[4 * g + 1] - mean) * rstd * w4.y + b4.y;
        v[4 * g + 2] = (v[4 * g + 2] - mean) * rstd * w4.z + b4.z;
        v[4 * g + 3] = (v[4 * g + 3] - mean) * rstd * w4.w + b4.w;
    }
}

// x <- x + LN1(agg + skip(x));  x <- x + LN2(FFN(x));  optional SiLU  (equivariant_scorenetwork.py:35-38,140-141)
// Thread = node: lin_skip(x) waits in TMEM (node_qkvs), the aggregate in q[node]; FFN.0 / FFN.3 are two N = 32 tcgen05 GEMMs on
// the node tile (operand rows written by the node's thread, FULL barrier -> leader issue); the new x goes to XT and, as the
// fp16 hi/lo operand, to the node tile for the next layer's q|k|v|skip GEMM.
__device__ __noinline__ uint32_t node_update(const Chunk c, bool silu_after, int layer, uint32_t tmem_base, uint32_t qs) {
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1), node = q * QT + e;
    if (q * QT >= c.n) return qs;
    const bool live = node < c.n;
    float* XT = smem_at<float>(c, SB_XT);
    float* Q = smem_at<float>(c, SB_Q);
    const float* Wp = smem_at<const float>(c, SB_WP);
    uint8_t* A = smem_bytes(c) + SB_R + q * NODE_A;
    const uint32_t a_addr = smem_u32(A), w_addr = smem_u32(Wp);
    const uint32_t bar_mma = bar_addr(c, 2 * q), bar_full = bar_addr(c, BAR_FULL0 + 2 * q);
    const uint32_t tq = tmem_base + q * 128;
    const uint32_t tlane = tq + (static_cast<uint32_t>(e & ~31) << 16);
    auto gemm32 = [&](const float* rows, uint32_t w_off_floats, uint32_t tcol) {   // rows -> operand, D[tcol, +32) = rows . W^T
        store_node_operand(A, e, rows);
        fence_proxy_async_smem();
        mbar_arrive(bar_full);
        if (quad_leader(q, e)) {
            uint32_t qt = qs;
            mbar_wait(bar_full, qt, QS_FULL0, c.status_flag);
            qs |= (qt & QS_DEAD);
            tc_fence_after();
            umma_split_f16<32, 2>(tq + tcol, a_addr, a_addr + 8192, w_addr + w_off_floats * 4, w_addr + w_off_floats * 4 + 2048, 0u);
            umma_commit(bar_mma);
        }
        qs ^= QS_FULL0;
        mbar_wait(bar_mma, qs, QS_MMA0, c.status_flag);
        tc_fence_after();
    };
    float y[32], x1[32];
    tc_fence_after();
    tmem_ld32(tlane + 96, y);   // lin_skip(x) without bias
    {
        const float4* bs = reinterpret_cast<const float4*>(Wp + MOLSDE_G_BQKVS + 96);
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float4 b = bs[g];
            const float4 ag = live ? *row_granule(Q, node, g) : make_float4(0.f, 0.f, 0.f, 0.f);   // attention aggregate
            y[4 * g] += b.x + ag.x; y[4 * g + 1] += b.y + ag.y; y[4 * g + 2] += b.z + ag.z; y[4 * g + 3] += b.w + ag.w;
        }
    }
    layer_norm_row(y, Wp + MOLSDE_G_LN1_W, Wp + MOLSDE_G_LN1_B);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 xo = live ? *row_granule(XT, node, g) : make_float4(0.f, 0.f, 0.f, 0.f);
        x1[4 * g] = xo.x + y[4 * g]; x1[4 * g + 1] = xo.y + y[4 * g + 1]; x1[4 * g + 2] = xo.z + y[4 * g + 2]; x1[4 * g + 3] = xo.w + y[4 * g + 3];
    }
    gemm32(x1, MOLSDE_G_F0C, 0);
    tmem_ld32(tlane, y);
    tc_fence_before();
    {
        const float* kp = (c.ffn_keep && live) ? c.ffn_keep + (static_cast<size_t>(layer) * c.N_total + c.node0 + node) * 32 : nullptr;
#pragma unroll
        for (int g = 0; g < 8; ++g) {
            const float4 b = reinterpret_cast<const float4*>(Wp + MOLSDE_G_F0_B)[g];
            float4 h = make_float4(silu_fast(y[4 * g] + b.x), silu_fast(y[4 * g + 1] + b.y), silu_fast(y[4 * g + 2] + b.z),
                                   silu_fast(y[4 * g + 3] + b.w));
            if (kp) {  // nn.Dropout between FFN.1 (SiLU) and FFN.3 in train mode
                const float4 k4 = __ldg(reinterpret_cast<const float4*>(kp) + g);
                h.x *= k4.x * c.inv_keep; h.y *= k4.y * c.inv_keep; h.z *= k4.z * c.inv_keep; h.w *= k4.w * c.inv_keep;
            }
            y[4 * g] = h.x; y[4 * g + 1] = h.y; y[4 * g + 2] = h.z; y[4 * g + 3] = h.w;
        }
    }
    gemm32(y, MOLSDE_G_F3C, 32);
    tmem_ld32(tlane + 32, y);
    tc_fence_before();
#pragma unroll
    for (int g = 0; g < 8; ++g) {
        const float4 b = reinterpret_cast<const float4*>(Wp + MOLSDE_G_F3_B)[g];
        y[4 * g] += b.x; y[4 * g + 1] += b.y; y[4 * g + 2] += b.z; y[4 * g + 3] += b.w;
    }
    layer_norm_row(y, Wp + MOLSDE_G_LN2_W, Wp + MOLSDE_G_LN2_B);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
        float x2 = x1[i] + y[i];
        if (silu_after) x2 = silu_fast(x2);
        y[i] = live ? x2 : 0.0f;
    }
    if (live) {
#pragma unroll
        for (int g = 0; g < 8; ++g) *row_granule(XT, node, g) = make_float4(y[4 * g], y[4 * g + 1], y[4 * g + 2], y[4 * g + 3]);
    }
    store_node_operand(A, e, y);   // A operand of the next layer's q|k|v|skip GEMM (its FFN.3 reader has completed)
    fence_proxy_async_smem();
    return qs;
}

// ---------------------------------------------------------------------------------------
// basis MLP + equivariant mean aggregation  (equivariant_scorenetwork.py:154-164)
//   hidden[128 edges x 128] = [h_row + h_col | edge_attr][128 x 64] . W1^T on tcgen05 (M = 128, N = 128, 4 K-steps x 3 split
//   terms, accumulator = the quad's 128 TMEM columns); the edge_attr half of the A operand arrives by TMA bulk copy from the
//   scratch record, the node half is gathered and split by the edge's thread.  Epilogue: the thread reads the 128 hidden
//   pre-activations of ITS edge (4 x tcgen05.ld x32): +bias, SiLU, 128 -> 3 projection in registers; frame mix; per-target mean.
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t phase_basis(const Chunk c, const float* __restrict__ blob, const uint8_t* __restrict__ scratch, int module,
                                             uint32_t tmem_base, uint32_t qs) {
    const float* Wb = smem_at<const float>(c, SB_BW);
    float* XTm = smem_at<float>(c, SB_XT);
    float* grad = smem_at<float>(c, SB_GRAD);
    const int* rowl = c.si + SI_ROWL;
    const int tid = threadIdx.x, q = tid >> 7, e = tid & (QT - 1);
    uint8_t* Ah = smem_bytes(c) + SB_BA + q * B_SLOT;
    uint8_t* Al = Ah + 16384;
    float* mix = smem_at<float>(c, SB_MIX) + q * (3 * TE);
    const uint32_t a_addr = smem_u32(Ah);
    const uint32_t w_addr = smem_u32(Wb);
    const uint32_t bar_mma = bar_addr(c, 2 * q), bar_tma = bar_addr(c, 2 * QUADS + q);
    const uint32_t tq = tmem_base + q * 128;
    const uint32_t tlane = tq + (static_cast<uint32_t>(e & ~31) << 16);
    const uint32_t bar_full = bar_addr(c, BAR_FULL0 + 2 * q);
    __syncthreads();
    stage_bulk2(c, smem_bytes(c) + SB_BW, blob + MOLSDE_P_BASIS0 + module * MOLSDE_P_BASIS_STRIDE, MOLSDE_P_BASIS_SZ, nullptr, nullptr, 0);
    const float4* EPI = reinterpret_cast<const float4*>(Wb + MOLSDE_B_EPI);
    // A operand of tile t: the edge_attr half (k-chunks 4..7 of the hi and the lo tile) by TMA from the scratch record, the node
    // half h_row + h_col (:154-155, k-chunks 0..3) gathered and split by the edge's thread; every thread then ARRIVES on the
    // quad's FULL barrier (nobody blocks here -- only the leader waits for it before issuing the MMAs)
    auto produce = [&](int t) {
        const TileInfo ti = tile_info(c, t);
        const uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES;
        if (quad_leader(q, e)) {
            fence_proxy_async_smem();
            mbar_expect_tx(bar_tma, 16384u);
            bulk_g2s(a_addr + 4 * 2048, rec + REC_EA_HI, 8192u, bar_tma);
            bulk_g2s(a_addr + 16384 + 4 * 2048, rec + REC_EA_LO, 8192u, bar_tma);
        }
        const bool live = e < ti.ne;
        const int sj = live ? rec[REC_SLOT + e] : 0, tg = live ? rec[REC_SLOT + TE + e] : 0;
#pragma unroll
        for (int kc = 0; kc < 4; ++kc) {
            const float4 a0 = *row_granule(XTm, sj, 2 * kc), a1 = *row_granule(XTm, sj, 2 * kc + 1);
            const float4 b0 = *row_granule(XTm, tg, 2 * kc), b1 = *row_granule(XTm, tg, 2 * kc + 1);
            const float v[8] = {a0.x + b0.x, a0.y + b0.y, a0.z + b0.z, a0.w + b0.w, a1.x + b1.x, a1.y + b1.y, a1.z + b1.z, a1.w + b1.w};
            store_a_chunk(Ah, Al, e, kc, v);
        }
        fence_proxy_async_smem();
        mbar_arrive(bar_full);
    };
    if (q < c.ntiles) produce(q);
    for (int t = q; t < c.ntiles; t += QUADS) {
        const TileInfo ti = tile_info(c, t);
        const uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES;
        const bool live = e < ti.ne;
        if (quad_leader(q, e)) {
            uint32_t qt = qs;   // (every thread flips its own copy of the two parities below)
            mbar_wait(bar_full, qt, QS_FULL0, c.status_flag);
            mbar_wait(bar_tma, qt, QS_TMA, c.status_flag);
            qs |= (qt & QS_DEAD);
            tc_fence_after();
            umma_split_f16<128, 4>(tq, a_addr, a_addr + 16384, w_addr + MOLSDE_B_W1_HI * 4, w_addr + MOLSDE_B_W1_LO * 4, 0u);
            umma_commit(bar_mma);
        }
        qs ^= (QS_TMA | QS_FULL0);
        float F[9];
        {
            const float* fr = reinterpret_cast<const float*>(rec + REC_FRAME) + e;
#pragma unroll
            for (int i = 0; i < 9; ++i) F[i] = live ? fr[i * TE] : 0.0f;
        }
        mbar_wait(bar_mma, qs, QS_MMA0, c.status_flag);
        tc_fence_after();
        // the operand slot is free again: stage the quad's next tile now, its TMA and gathers overlap the epilogue below
        if (t + QUADS < c.ntiles) produce(t + QUADS);
        float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
#pragma unroll 1
        for (int cb = 0; cb < 4; ++cb) {
            float hv[32];
            tmem_ld32(tlane + 32 * cb, hv);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const float4 ep = EPI[32 * cb + j];  // {bias, w2[0], w2[1], w2[2]} of hidden unit 32*cb + j
                const float h = silu_fast(hv[j] + ep.x);
                p0 = fmaf(h, ep.y, p0);
                p1 = fmaf(h, ep.z, p1);
                p2 = fmaf(h, ep.w, p2);
            }
        }
        tc_fence_before();
        // basis_mix = dyn0 * coord_diff + dyn1 * coord_cross + dyn2 * coord_vertical  (:159), frames cached by E0
        if (live) {
            const float d0 = p0 + Wb[MOLSDE_B_B2], d1 = p1 + Wb[MOLSDE_B_B2 + 1], d2 = p2 + Wb[MOLSDE_B_B2 + 2];
#pragma unroll
            for (int ax = 0; ax < 3; ++ax) mix[ax * TE + e] = d0 * F[ax] + d1 * F[3 + ax] + d2 * F[6 + ax];
        }
        quad_sync(q);
        // gradient_i (+)= mean over the incoming edges (:162-164), ascending-source order
        const int ntg = ti.tb - ti.ta;
        for (int p = e; p < ntg * 3; p += QT) {
            const int i = ti.ta + p / 3, ax = p % 3;
            const int s0 = rowl[i] - ti.ea, s1 = rowl[i + 1] - ti.ea;
            float sacc = 0.0f;
            for (int s = s0; s < s1; ++s) sacc += mix[ax * TE + s];
            sacc = __fdiv_rn(sacc, static_cast<float>(max(s1 - s0, 1)));  // aggr='mean'
            grad[i * 3 + ax] = (module == 0) ? sacc : grad[i * 3 + ax] + sacc;
        }
        quad_sync(q);  // `mix` is rewritten by the next tile; its TMEM loads are ordered before the next MMA issue
    }
    __syncthreads();
    return qs;
}

// ---------------------------------------------------------------------------------------
// one full network evaluation on the chunk: positions in smem -> "gradient" in smem
// ---------------------------------------------------------------------------------------
__device__ __noinline__ uint32_t score_eval(const Chunk c, const float* __restrict__ blob, const float* __restrict__ nattr,
                                            const float* __restrict__ e2d_tiles, uint8_t* __restrict__ scratch, uint32_t tmem_base,
                                            uint32_t qs) {
    PROF_T0();
    qs = phase_edge_features(c, blob, e2d_tiles, scratch, tmem_base, qs);
    PROF_ADD(0);
    // conv_input = node_attr (loop-invariant node_emb output): XT rows + the operand of the first q|k|v|skip GEMM
    node_stage(c, nattr);
    for (int module = 0; module < 2; ++module) {
        for (int conv = 0; conv < 2; ++conv) {
            const float* sec = blob + MOLSDE_P_GAT0 + (2 * module + conv) * MOLSDE_P_GAT_SZ;
            __syncthreads();  // previous users of the WP / SB_WQ regions are done; the node operand tiles are complete
            stage_bulk2(c, smem_bytes(c) + SB_WP, sec, MOLSDE_G_WP_SZ, smem_bytes(c) + SB_WQ, sec + MOLSDE_G_WQKVS,
                        MOLSDE_P_GAT_SZ - MOLSDE_G_WQKVS);
            PROF_ADD(1);
            qs = node_qkvs(c, tmem_base, qs);
            __syncthreads();
            PROF_ADD(2);
            qs = gat_edge_phase(c, scratch, 2 * module + conv, tmem_base, qs);
            __syncthreads();
            PROF_ADD(3);
            qs = node_update(c, conv == 0, 2 * module + conv, tmem_base, qs);
            PROF_ADD(4);
        }
        qs = phase_basis(c, blob, scratch, module, tmem_base, qs);
        PROF_ADD(5);
        if (module == 0) node_stage(c, nullptr);  // the basis operand buffers overlaid the node operand tiles
    }
    return qs;
}

// TMEM accumulators (128 columns per quad) + mbarriers; call with all threads of the CTA
__device__ __forceinline__ uint32_t tmem_setup(const Chunk& c) {
    if (threadIdx.x == 0) {
        for (int i = 0; i < NUM_BARS; ++i)
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_addr(c, i)), "r"(i >= BAR_FULL0 ? QT : 1));
        c.si[SI_MISC + 2] = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(c.si + SI_MISC + 1)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    return static_cast<uint32_t>(c.si[SI_MISC + 1]);
}
__device__ __forceinline__ void tmem_teardown(uint32_t tmem_base) {
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
}

// (source, target) of edge slot `slot` of a tile, chunk-local: bisection on the row pointer + one global load
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

// Chunk bookkeeping into shared memory + the static slot tables (chunk-local source / target bytes of every edge slot) into the
// chunk's scratch records: resolved once per chunk instead of at each of the 7 tile visits of every score evaluation.
__device__ __forceinline__ bool load_chunk(Chunk& c, const molsde_plan& plan, int chunk, uint8_t* __restrict__ scratch,
                                           int64_t scratch_tiles) {
    const int tid = threadIdx.x;
    c.tile0 = plan.chunk_tile_ptr[chunk];
    c.ntiles = plan.chunk_tile_ptr[chunk + 1] - c.tile0;
    c.node0 = plan.tile_tgt_ptr[c.tile0];
    c.n = plan.tile_tgt_ptr[c.tile0 + c.ntiles] - c.node0;
    c.edge0 = plan.rowptr[c.node0];
    bool ok = (c.n <= MAXN) && (c.ntiles <= MAXT) && (c.n >= 0) && (c.ntiles <= scratch_tiles);
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
            for (int item = tid; item < c.ntiles * TE; item += NTHREADS) {
                const int t = item / TE, slot = item % TE;
                const TileInfo ti = tile_info(c, t);
                int sj, tg;
                resolve_slot(c, plan.src, ti, slot, sj, tg);
                uint8_t* rec = scratch + static_cast<size_t>(t) * REC_BYTES + REC_SLOT;
                rec[slot] = static_cast<uint8_t>(sj);
                rec[TE + slot] = static_cast<uint8_t>(tg);
            }
            __syncthreads();
        }
    }
    if (!ok && tid == 0 && c.status_flag) atomicExch(c.status_flag, 1 + chunk);
    return ok;
}

// ---------------------------------------------------------------------------------------
// K3: get_score for a whole batch, SDE_model_2D_to_3D.py:393-445
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS, 1)
sde2d3d_score_kernel(molsde_plan plan, const float* __restrict__ blob, const float* __restrict__ nattr,
                     const float* __restrict__ e2d_tiles, const float* __restrict__ pos,
                     const float* __restrict__ stdv, float* __restrict__ score, float* __restrict__ scratch,
                     int64_t scratch_tiles, int32_t* status_flag, const float* __restrict__ attn_keep,
                     const float* __restrict__ ffn_keep, float inv_keep) {
    extern __shared__ __align__(128) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(smem) + SB_INT);
    c.attn_keep = attn_keep; c.ffn_keep = ffn_keep; c.inv_keep = inv_keep;
    c.E_total = plan.E; c.N_total = plan.N;
    c.status_flag = status_flag;
    uint8_t* my_scratch = reinterpret_cast<uint8_t*>(scratch) + static_cast<size_t>(blockIdx.x) * scratch_tiles * REC_BYTES;
    const uint32_t tmem_base = tmem_setup(c);
    uint32_t qs = 0;
    float* P = smem_at<float>(c, SB_POS);
    const float* G = smem_at<const float>(c, SB_GRAD);
    for (int chunk = blockIdx.x; chunk < plan.num_chunks; chunk += gridDim.x) {
        __syncthreads();
        if (!load_chunk(c, plan, chunk, my_scratch, scratch_tiles)) continue;
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) P[i] = pos[static_cast<size_t>(c.node0) * 3 + i];
        __syncthreads();
        qs = score_eval(c, blob, nattr, e2d_tiles, my_scratch, tmem_base, qs);
        for (int i = threadIdx.x; i < c.n * 3; i += NTHREADS) {
            // get_score: scores = -output / std  (:440-443);  forward (stdv == NULL): the raw network output (:379)
            score[static_cast<size_t>(c.node0) * 3 + i] = stdv ? __fdiv_rn(-G[i], stdv[c.node0 + i / 3]) : G[i];
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
                  float* __restrict__ scratch, int64_t scratch_tiles, int32_t* work_counter, int32_t* status_flag) {
    extern __shared__ __align__(128) float smem[];
    Chunk c;
    c.sm = smem;
    c.si = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(smem) + SB_INT);
    int* misc = c.si + SI_MISC;
    c.attn_keep = nullptr; c.ffn_keep = nullptr; c.inv_keep = 1.0f; c.E_total = plan.E; c.N_total = plan.N;
    c.status_flag = status_flag;
    const uint32_t tmem_base = tmem_setup(c);
    uint32_t qs = 0;
    uint8_t* my_scratch = reinterpret_cast<uint8_t*>(scratch) + static_cast<size_t>(blockIdx.x) * scratch_tiles * REC_BYTES;
    float* P = smem_at<float>(c, SB_POS);
    float* G = smem_at<float>(c, SB_GRAD);
    float* SC = smem_at<float>(c, SB_SCORE);   // (score / noise alias the phase union: only live between two evaluations)
    float* NZ = smem_at<float>(c, SB_NOISE);
    float* red = smem_at<float>(c, SB_RED);
    const int tid = threadIdx.x;
    const size_t N3 = static_cast<size_t>(plan.N) * 3;
    while (true) {
        __syncthreads();
        if (tid == 0) misc[0] = atomicAdd(work_counter, 1);
        __syncthreads();
        if (misc[0] >= plan.num_chunks) break;
        const int chunk = plan.chunk_order ? plan.chunk_order[misc[0]] : misc[0];  // longest groups first
        if (!load_chunk(c, plan, chunk, my_scratch, scratch_tiles)) continue;
        const int n3 = c.n * 3;
        const size_t g3 = static_cast<size_t>(c.node0) * 3;
        for (int i = tid; i < n3; i += NTHREADS) P[i] = pos_init[g3 + i];
        __syncthreads();
        for (int step = 0; step < cfg.steps; ++step) {
            const float stdv = step_table[step * 8 + 0], Gd = step_table[step * 8 + 1];
            const float sqrt_alpha = step_table[step * 8 + 2], calpha = step_table[step * 8 + 3];
            // ---------------- corrector (LangevinCorrector.update_fn :191-212) ----------------
            qs = score_eval(c, blob, nattr, e2d_tiles, my_scratch, tmem_base, qs);
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
            qs = score_eval(c, blob, nattr, e2d_tiles, my_scratch, tmem_base, qs);
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
// Step-wise predictor / corrector updates for sampling groups that do NOT fit one CTA (> 224 atoms or > 64 edge tiles; the
// reference has no limit: `num_repeat` conformers of a 30-atom molecule are already 300 atoms).  The host captures
//   [score network over the batch (sde2d3d_score_kernel, chunks of whole molecules) -> corrector update -> score -> predictor update]
// in a CUDA graph and replays it once per reverse step; the step index lives in device memory (`step_counter`), the per-step
// schedule constants come from the same `step_table` as the fused kernel, the noise from the same Philox streams
// (seed, node, step, 0 | 1), so a group small enough for both paths is sampled identically up to summation order.
//   corrector: one CTA per group -- the Langevin step size is a per-group mean of norms (F9) -- fixed-order block reduction;
//   predictor: elementwise over the atoms.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NTHREADS)
pc_corrector_update_kernel(const float* __restrict__ net_out, float* __restrict__ pos, const int32_t* __restrict__ group_node_ptr,
                           const float* __restrict__ step_table, const int32_t* __restrict__ step_counter, float snr, float scale_eps,
                           uint64_t seed, const float* __restrict__ noise_corr, int64_t N) {
    __shared__ float red[64];
    const int g = blockIdx.x, tid = threadIdx.x;
    const int a = group_node_ptr[g], n = group_node_ptr[g + 1] - a;
    if (n <= 0) return;
    const int step = *step_counter;
    const float stdv = step_table[step * 8 + 0], calpha = step_table[step * 8 + 3];
    const float* nc = noise_corr ? noise_corr + static_cast<size_t>(step) * N * 3 : nullptr;
    float gn = 0.0f, nn = 0.0f;
    for (int i = tid; i < n; i += NTHREADS) {
        const size_t o = static_cast<size_t>(a + i) * 3;
        const float s0 = __fdiv_rn(-net_out[o], stdv), s1 = __fdiv_rn(-net_out[o + 1], stdv), s2 = __fdiv_rn(-net_out[o + 2], stdv);
        float nz[3];
        if (nc) { nz[0] = nc[o]; nz[1] = nc[o + 1]; nz[2] = nc[o + 2]; } else normal3(seed, a + i, step, 0u, nz);
        gn += sqrtf(s0 * s0 + s1 * s1 + s2 * s2);
        nn += sqrtf(nz[0] * nz[0] + nz[1] * nz[1] + nz[2] * nz[2]);
    }
    gn = block_sum(gn, red) / static_cast<float>(n);
    nn = block_sum(nn, red) / static_cast<float>(n);
    const float ratio = __fdiv_rn(__fmul_rn(snr, nn), gn);
    const float step_size = __fmul_rn(__fmul_rn(__fmul_rn(ratio, ratio), 2.0f), calpha);   // :209
    const float nscale = sqrtf(__fmul_rn(step_size, 2.0f));
    for (int i = tid; i < n; i += NTHREADS) {
        const size_t o = static_cast<size_t>(a + i) * 3;
        float nz[3];
        if (nc) { nz[0] = nc[o]; nz[1] = nc[o + 1]; nz[2] = nc[o + 2]; } else normal3(seed, a + i, step, 0u, nz);
#pragma unroll
        for (int ax = 0; ax < 3; ++ax) {
            const float sc = __fdiv_rn(-net_out[o + ax], stdv);
            const float xm = __fadd_rn(pos[o + ax], __fmul_rn(step_size, sc));                       // :210
            pos[o + ax] = __fadd_rn(xm, __fmul_rn(__fmul_rn(nscale, nz[ax]), scale_eps));          // :211
        }
    }
}

__global__ void __launch_bounds__(256)
pc_predictor_update_kernel(const float* __restrict__ net_out, float* __restrict__ pos, float* __restrict__ pos_mean,
                           const float* __restrict__ step_table, const int32_t* __restrict__ step_counter, uint64_t seed,
                           const float* __restrict__ noise_pred, int64_t N) {
    const int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
    if (i >= N) return;
    const int step = *step_counter;
    const float stdv = step_table[step * 8 + 0], Gd = step_table[step * 8 + 1], sqrt_alpha = step_table[step * 8 + 2];
    float nz[3];
    if (noise_pred) {
        const float* np = noise_pred + (static_cast<size_t>(step) * N + i) * 3;
        nz[0] = np[0]; nz[1] = np[1]; nz[2] = np[2];
    } else {
        normal3(seed, static_cast<uint32_t>(i), step, 1u, nz);
    }
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
        const float x = pos[3 * i + ax];
        const float sc = __fdiv_rn(-net_out[3 * i + ax], stdv);
        const float f = __fsub_rn(__fmul_rn(sqrt_alpha, x), x);                    // SDE_sparse.py:160 / 220
        const float rev_f = __fsub_rn(f, __fmul_rn(__fmul_rn(Gd, Gd), sc));        // SDE_sparse.py:98
        const float xmean = __fsub_rn(x, rev_f);                                   // :166
        pos[3 * i + ax] = __fadd_rn(xmean, __fmul_rn(Gd, nz[ax]));                 // :167
        pos_mean[3 * i + ax] = xmean;
    }
}
__global__ void pc_step_advance_kernel(int32_t* step_counter) { *step_counter += 1; }

// ---------------------------------------------------------------------------------------
// edge_2D_emb (eval): e2d tile = W3 . relu(U[src] + V[tgt]) + b3,  SDE_model_2D_to_3D.py:405-407
// uv [N][600]: columns 0..299 = folded first layer applied to h[row], 300..599 to h[col].
// One-time (loop-invariant) kernel: fp32 FFMA register tile, 256 threads, output in the [8][128][4] tile layout.
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
        // output tile [8 feature quads][128 slots][4]: the edge's thread of the score kernels reads its 32 values as 8 coalesced 16 B loads
        float* out = e2d_tiles + static_cast<size_t>(tile) * E2D_TILE_FLOATS;
        const float4 bq = *reinterpret_cast<const float4*>(b3 + to * 4);  // this thread: feature quad `to`, slots 4*te .. 4*te+3
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int s = te * 4 + i;
            float4 o = make_float4(acc[i][0] + bq.x, acc[i][1] + bq.y, acc[i][2] + bq.z, acc[i][3] + bq.w);
            if (s >= ne) o = make_float4(0.f, 0.f, 0.f, 0.f);  // dead slots stay zero
            *reinterpret_cast<float4*>(out + (to * TE + s) * 4) = o;
        }
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

int64_t molsde_tile_floats(void) { return E2D_TILE_FLOATS; }

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
    return static_cast<int64_t>(ctas) * max_chunk_tiles * REC_FLOATS;
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
    const int64_t stride = (scratch_floats / ctas) / REC_FLOATS;  // scratch records (tiles) per CTA
    if (stride < 1) return MOLSDE_ERR_WORKSPACE;
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
    const int64_t stride = (scratch_floats / ctas) / REC_FLOATS;  // scratch records (tiles) per CTA
    if (stride < 1) return MOLSDE_ERR_WORKSPACE;
    cudaError_t err = cudaFuncSetAttribute(sde2d3d_pc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES);
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    err = cudaMemsetAsync(work_counter, 0, sizeof(int32_t), as_stream(stream));
    if (err != cudaSuccess) { set_last_error(cudaGetErrorString(err)); return MOLSDE_ERR_CUDA; }
    sde2d3d_pc_kernel<<<ctas, NTHREADS, SMEM_BYTES, as_stream(stream)>>>(
        *plan, params->blob, nattr, e2d_tiles, pos_init, step_table, *cfg, noise_corr, noise_pred, pos_out,
        pos_mean_out, scratch, stride, work_counter, status_flag);
    return check_launch("sde2d3d_pc");
}

int molsde_sde2d3d_pc_corrector_update(const float* net_out, float* pos, const int32_t* group_node_ptr, int32_t num_groups,
                                       const float* step_table, const int32_t* step_counter, float snr, float scale_eps, uint64_t seed,
                                       const float* noise_corr, int64_t N, void* stream) {
    if (!net_out || !pos || !group_node_ptr || !step_table || !step_counter || num_groups < 0 || N < 0) return MOLSDE_ERR_INVALID;
    if (num_groups == 0) return MOLSDE_OK;
    pc_corrector_update_kernel<<<num_groups, NTHREADS, 0, as_stream(stream)>>>(net_out, pos, group_node_ptr, step_table, step_counter, snr,
                                                                              scale_eps, seed, noise_corr, N);
    return check_launch("pc_corrector_update");
}

int molsde_sde2d3d_pc_predictor_update(const float* net_out, float* pos, float* pos_mean, const float* step_table,
                                       int32_t* step_counter, uint64_t seed, const float* noise_pred, int64_t N, void* stream) {
    if (!net_out || !pos || !pos_mean || !step_table || !step_counter || N < 0) return MOLSDE_ERR_INVALID;
    if (N > 0) {
        pc_predictor_update_kernel<<<static_cast<unsigned>((N + 255) / 256), 256, 0, as_stream(stream)>>>(net_out, pos, pos_mean, step_table,
                                                                                                   step_counter, seed, noise_pred, N);
        int st = check_launch("pc_predictor_update");
        if (st != MOLSDE_OK) return st;
    }
    pc_step_advance_kernel<<<1, 1, 0, as_stream(stream)>>>(step_counter);   // the next replay of the captured step reads step + 1
    return check_launch("pc_step_advance");
}

}  // extern "C"
