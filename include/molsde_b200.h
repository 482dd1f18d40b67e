/*
 * molsde_b200 -- C ABI of the B200-native (sm_100a) MoleculeSDE hot path.
 *
 * The reference (chao1224/MoleculeSDE) is pure Python and has no FFI: its hot path calls
 * third-party wheels (torch_cluster / torch_sparse / torch_scatter / torch_geometric) and
 * torch ops.  Each entry point below replaces one of those call sequences; the reference
 * site it replaces is cited as `file:line` relative to the reference root.  A Python host
 * binds these with ctypes (see INTEGRATION.md and moleculesde_b200/_abi.py).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - no entry point allocates, synchronises or keeps global state; all work is enqueued on
 *     `stream` (a cudaStream_t passed as void*);
 *   - return value: 0 = MOLSDE_OK, negative = error (see molsde_status); launch errors are
 *     reported as MOLSDE_ERR_CUDA, and molsde_last_error_string() describes the last one;
 *   - node / edge indices inside the kernels are int32; the int64 `[2,E]` tensors of the
 *     reference API are accepted/produced where the reference exposes them.
 */
#ifndef MOLSDE_B200_H_
#define MOLSDE_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum molsde_status {
    MOLSDE_OK = 0,
    MOLSDE_ERR_INVALID = -1,     /* bad argument (null pointer, size out of range) */
    MOLSDE_ERR_UNSUPPORTED = -2, /* shape beyond a compiled limit (see limits below) */
    MOLSDE_ERR_CUDA = -3,        /* CUDA runtime / launch failure */
    MOLSDE_ERR_WORKSPACE = -4    /* workspace too small */
} molsde_status;

/* compiled limits */
#define MOLSDE_MAX_MOL_NODES 128   /* atoms per molecule for the graph builders */
#define MOLSDE_CHUNK_MAX_NODES 224 /* atoms per CTA chunk of the 2D->3D score / PC kernels */
#define MOLSDE_TILE_EDGES 128      /* edges per tile (tiles are aligned to target nodes) */
#define MOLSDE_TILE_LD 136         /* padded row length of a per-edge attribute tile [MOLSDE_HID][MOLSDE_TILE_LD] */
#define MOLSDE_HID 32              /* hidden_dim of SDEModel2Dto3D_02 (pretrain_MoleculeSDE.py:226) */
#define MOLSDE_EMB 300             /* emb_dim (config.py:84) */

/* Host-side bookkeeping behind the `molsde_plan` struct, no device work: chunk / tile boundaries from HOST copies of the CSR row pointer and the
 * molecule offsets; `groups` (optional, [G+1] molecule offsets) fixes one chunk per sampling group.  counts_out[4] = {num_chunks,
 * num_tiles, max tiles per chunk, offending index}; see csrc/core.cu. */
int molsde_build_plan_host(const int64_t* rowptr, const int64_t* node_ptr, int32_t B, const int64_t* groups, int32_t G, int32_t tile_edges,
                           int32_t max_nodes, int32_t max_tiles, int32_t* chunk_tile_ptr, int32_t* tile_tgt_ptr, int64_t* counts_out);

/* test hook: echoes its arguments into out[12] and returns 7 (calling-convention check of the Python fast-call path) */
int molsde_debug_echo(int64_t a, float x, int32_t b, const void* p, float y, int64_t c, int32_t d, uint64_t e, int64_t f, float z,
                      int64_t g, int32_t h, double* out);

const char* molsde_version(void);
const char* molsde_last_error_string(void);
/* compute capability check: returns MOLSDE_OK only on an sm_100 device */
int molsde_check_device(int device);

/* ------------------------------------------------------------------------------------
 * Graph construction (bit-exact integer kernels)
 * ---------------------------------------------------------------------------------- */

/* ptr[s] = first position p in [0,M) with key(p) >= s, s = 0..num_segments; key(p) =
 * keys[p] (indirect == NULL) or keys[indirect[p]].  keys must be ascending along p.
 * Replaces the per-graph offsets PyG's Batch keeps (`batch` vector bookkeeping,
 * SDE_model_3D_to_2D_node_adj_dense.py:124-127). */
int molsde_segment_ptr(const int64_t* keys, const int64_t* indirect, int64_t M, int32_t num_segments,
                       int32_t* ptr, void* stream);

/* exclusive scan of int32 counts[n] into out[n+1] (out[n] = total), single launch. */
int molsde_exclusive_scan_i32(const int32_t* counts, int64_t n, int32_t* out, void* stream);

/* extend_graph, Geom3D/datasets/dataset_3D.py:12-35 (torch_sparse.spspmm + coalesce, twice):
 * per molecule E2 = E u (E.E \ diag), E4 = E2 u (E2.E2 \ diag).
 *   edge_index  int64 [2,E_b] (global node ids, grouped by molecule), node_ptr/edge_ptr
 *   int32 [B+1] molecule offsets.  Pass 1 writes deg[N] (row lengths); after an exclusive
 *   scan into rowptr[N+1], pass 2 writes col int32[E_x] (row-major sorted == coalesce order)
 *   and, if non-NULL, the reference-layout int64 [2,E_x] tensor `ext_edge_index`. */
int molsde_extend_graph_count(const int64_t* edge_index, int64_t E_b, const int32_t* node_ptr,
                              const int32_t* edge_ptr, int32_t B, int32_t* deg, void* stream);
int molsde_extend_graph_fill(const int64_t* edge_index, int64_t E_b, const int32_t* node_ptr,
                             const int32_t* edge_ptr, int32_t B, const int32_t* rowptr, int64_t E_x,
                             int32_t* col, int64_t* ext_edge_index, void* stream);

/* radius_graph, Geom3D/models/schnet.py:91 (torch_cluster.radius_graph, CUDA semantics:
 * d^2 < r^2 strict, first max_num_neighbors+1 hits by ascending index incl. self, self dropped).
 * Output CSR by target with ascending sources; `edge_index` int64 [2,E_r] (row0 = source). */
int molsde_radius_graph_count(const float* pos, const int32_t* node_ptr, int32_t B, float r,
                              int32_t max_num_neighbors, int32_t* deg, void* stream);
int molsde_radius_graph_fill(const float* pos, const int32_t* node_ptr, int32_t B, float r,
                             int32_t max_num_neighbors, const int32_t* rowptr, int64_t E_r, int32_t* col,
                             int64_t* edge_index, void* stream);

/* CSR-by-target view of a generic `[2,E]` int64 edge_index (row0 = source j, row1 = target i,
 * edges grouped by molecule): rowptr[N+1], src[E] and perm[E] (position in the input list), stable
 * in input order -- the accumulation order of MessagePassing.propagate
 * (schnet.py:190, equivariant_scorenetwork.py:71). Two passes like the builders above. */
int molsde_csr_by_target_count(const int64_t* edge_index, int64_t E, int64_t N, int32_t* deg, void* stream);
int molsde_csr_by_target_fill(const int64_t* edge_index, int64_t E, const int32_t* node_ptr,
                              const int32_t* edge_ptr, int32_t B, const int32_t* rowptr, int32_t* src,
                              int32_t* perm, void* stream);

/* ------------------------------------------------------------------------------------
 * Dense node-level linear layer  Y[M,N] = act(X[M,K] . W[N,K]^T + b) (+ R)   (fp32 FFMA)
 * Replaces torch.nn.Linear on node-major tensors (SDE_model_2D_to_3D.py:264,375; schnet.py:99-101,
 * 163-167,189-191; layers/common.py:31-40).  act: 0 none, 1 relu, 2 silu, 3 shifted softplus
 * (schnet.py:210-216), 4 tanh, 5 elu.  rowscale (nullable, [M]) multiplies the pre-activation (mask_x);
 * R (nullable, leading dim ldr) is added after the activation (residual, schnet.py:97).
 * ---------------------------------------------------------------------------------- */
int molsde_linear(const float* X, int64_t M, int32_t K, int64_t ldx, const float* W, const float* b,
                  int32_t N, float* Y, int64_t ldy, int32_t act, const float* R, int64_t ldr, const float* rowscale,
                  void* stream);

/* ------------------------------------------------------------------------------------
 * SDEModel2Dto3D_02 (Geom3D/models/MoleculeSDE/SDE_model_2D_to_3D.py:252-445)
 *
 * A "plan" describes how a batch is cut into CTA chunks (whole molecules, <=
 * MOLSDE_CHUNK_MAX_NODES atoms) and target-aligned tiles of <= MOLSDE_TILE_EDGES edges:
 *   chunk_tile_ptr int32 [C+1], tile_tgt_ptr int32 [T+1] (first target node of each tile; chunk
 *   boundaries are tile boundaries), rowptr int32 [N+1] / src int32 [E] the CSR by target.
 * Per-edge tensors exchanged between kernels use the tile layout  [T][MOLSDE_HID][MOLSDE_TILE_LD]
 * (feature-major inside a tile, edge slot = csr_position - rowptr[first target of the tile], slots
 * >= the tile's edge count and the 8 pad columns are zero); molsde_tile_floats() floats per tile.
 * ---------------------------------------------------------------------------------- */

/* float offsets of the packed parameter blob (built by the host from the state_dict) */
typedef struct molsde_sde2d3d_params {
    const float* blob;  /* layout: see csrc/sde2d3d_params.h (MOLSDE_P_* offsets) */
    int64_t blob_floats;
} molsde_sde2d3d_params;

typedef struct molsde_plan {
    int32_t num_chunks, num_tiles;
    int64_t N, E;
    const int32_t* chunk_tile_ptr; /* [C+1] */
    const int32_t* tile_tgt_ptr;   /* [T+1] */
    const int32_t* rowptr;         /* [N+1] */
    const int32_t* src;            /* [E]   */
    const int32_t* chunk_order;    /* [C] or NULL: order in which the persistent PC kernel hands out chunks
                                      (host sorts by descending tile count = longest-processing-time first) */
} molsde_plan;

/* edge_2D_emb in eval mode (BatchNorm running stats), SDE_model_2D_to_3D.py:265,405-407:
 *   uv [N,600] = node-factored first layer with BN folded in (host: molsde_linear on the folded
 *   weights), out e2d in tile layout = W3 . relu(uv[src,:300] + uv[tgt,300:]) + b3. */
int molsde_edge2d_emb_eval(const molsde_plan* plan, const float* uv, const float* w3t /*[300][32]*/,
                           const float* b3 /*[32]*/, float* e2d_tiles, void* stream);

/* get_score, SDE_model_2D_to_3D.py:393-445: score[N,3] = -gradient / std.
 *   nattr [N,32] = node_emb(node_2D_repr) (loop invariant), e2d_tiles from the call above,
 *   pos [N,3], std [N] = marGINal_prob(.., t)[1].  scratch: num_ctas * max_chunk_tiles *
 *   molsde_tile_floats() floats. */
int molsde_sde2d3d_score(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                         const float* e2d_tiles, const float* pos, const float* std, float* score,
                         float* scratch, int64_t scratch_floats, int32_t* status_flag, void* stream);
/* SDEModel2Dto3D_02.forward pieces (train mode, SDE_model_2D_to_3D.py:306-391; forward values):
 *  - molsde_edge2d_bn_train: BatchNorm1d(300) of edge_2D_emb with BATCH statistics over the E pre-activations
 *    uv[src,:F] + uv[tgt,F:]; folds the normalisation into uv in place (then molsde_edge2d_emb_eval applies), returns the
 *    batch mean / biased variance and updates the running statistics (momentum, unbiased variance) when given;
 *  - molsde_sde2d3d_forward_net: raw network output ("gradient", :379) at perturbed positions; attn_keep [4][E][8] (CSR
 *    edge order) / ffn_keep [4][N][32] are the 0/1 keep-masks of the two dropouts of every GATLayer (NULL = eval);
 *  - molsde_dsm_pos_loss: out[g] = mean_{i in g} sum_xyz (score-noise)^2 * w[i], mean_out[0] = mean_g out[g] (:380-390;
 *    w NULL = 1, mean_out optional). */
int molsde_edge2d_bn_train(const molsde_plan* plan, float* uv, int32_t F, const float* gamma, const float* beta, float eps,
                           float momentum, float* running_mean, float* running_var, float* batch_mean, float* batch_var,
                           void* stream);
int molsde_sde2d3d_forward_net(const molsde_plan* plan, const molsde_sde2d3d_params* params, const float* nattr,
                               const float* e2d_tiles, const float* pos, const float* attn_keep, const float* ffn_keep,
                               float dropout_p, float* gradient, float* scratch, int64_t scratch_floats, int32_t* status_flag,
                               void* stream);
/* out[i,:] = mean_coeff[i] * x[i,:] + stdv[i] * noise[i,:]  (forward perturbation, :331-332; mean_coeff NULL = 1) */
int molsde_perturb_rows(const float* x, const float* mean_coeff, const float* stdv, const float* noise, int64_t N, int32_t D,
                        float* out, void* stream);
int molsde_dsm_pos_loss(const float* score, const float* noise, const float* w, const int32_t* node_ptr, int32_t B, float* out,
                        float* mean_out, void* stream);
int64_t molsde_sde2d3d_scratch_floats(const molsde_plan* plan, int32_t max_chunk_tiles, int32_t* num_ctas_out);
int64_t molsde_tile_floats(void);

/* position_PC_generation, examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-138:
 * the whole reverse-SDE loop (LangevinCorrector :191-212 + ReverseDiffusionPredictor :163-168 over
 * RSDE.discretize, SDE_sparse.py:94-100) for independent sampling groups; one chunk == one group
 * (the corrector step size is a mean over the group's atoms, F9), one persistent CTA per group.
 *   step_table float [steps][8]: {std, G, sqrt_alpha (VE: 1), corr_alpha, 0,0,0,0} per reverse step,
 *   computed by the host from the SDE object exactly as the reference does.
 *   noise: if noise_corr/noise_pred are non-NULL they are [steps][N][3] injected draws (parity
 *   mode); otherwise Philox4x32-10 + Box-Muller in-kernel with (seed, node, step) counters. */
typedef struct molsde_pc_config {
    int32_t steps;
    float snr, scale_eps;
    uint64_t seed;
} molsde_pc_config;
int molsde_sde2d3d_pc_sample(const molsde_plan* plan, const molsde_sde2d3d_params* params,
                             const float* nattr, const float* e2d_tiles, const float* pos_init,
                             const float* step_table, const molsde_pc_config* cfg, const float* noise_corr,
                             const float* noise_pred, float* pos_out, float* pos_mean_out, float* scratch,
                             int64_t scratch_floats, int32_t* work_counter, int32_t* status_flag,
                             void* stream);

/* Step-wise form of the same loop for sampling groups beyond the fused kernel's shared-memory limits (> 224 atoms / 64 tiles; the
 * reference's position_PC_generation has no size limit, ..._inference_2D_to_3D_VE_VP.py:92-138).  One reverse step =
 *   molsde_sde2d3d_forward_net (raw network output) -> pc_corrector_update (LangevinCorrector.update_fn :191-212, per-GROUP mean
 *   norms: group_node_ptr int32 [G+1] node offsets) -> forward_net -> pc_predictor_update (ReverseDiffusionPredictor.update_fn
 *   :163-168; also advances *step_counter).  step_table / seed / noise semantics are those of molsde_sde2d3d_pc_sample; the step
 *   index is read from device memory so the four calls can be captured once in a CUDA graph and replayed per step. */
int molsde_sde2d3d_pc_corrector_update(const float* net_out, float* pos, const int32_t* group_node_ptr, int32_t num_groups,
                                       const float* step_table, const int32_t* step_counter, float snr, float scale_eps, uint64_t seed,
                                       const float* noise_corr, int64_t N, void* stream);
int molsde_sde2d3d_pc_predictor_update(const float* net_out, float* pos, float* pos_mean, const float* step_table,
                                       int32_t* step_counter, uint64_t seed, const float* noise_pred, int64_t N, void* stream);

/* ------------------------------------------------------------------------------------
 * SchNet (Geom3D/models/schnet.py:16-216), forward
 * ---------------------------------------------------------------------------------- */

/* CFConv message passing of one InteractionBlock (schnet.py:185-195) fused with GaussianSmearing
 * (:205-207), the filter MLP and the cosine cutoff:  agg[i,:] = sum_{j->i} x[j,:] * W_ij,
 *   W_ij = (mlp.2 . ssp . mlp.0)(exp(coeff (d_ij - mu_k)^2)) * 0.5 (cos(d_ij pi / cutoff) + 1).
 * plan: tiles over the radius-graph CSR (tile_tgt_ptr / rowptr / src; chunks unused);
 * x [N,128] = conv.lin1(h); w1t [56][136] / w2t [128][136] k-major, zero padded; mu [56] offsets. */
int molsde_schnet_cfconv(const molsde_plan* plan, const float* pos, const float* x, const float* w1t, const float* b1,
                         const float* w2t, const float* b2, const float* mu, int32_t num_gaussians, float coeff,
                         float cutoff, float* agg, void* stream);
/* out[r,:] = table[idx[r],:]  (Embedding lookup, schnet.py:89) */
int molsde_gather_rows(const float* table, const int64_t* idx, int64_t rows, int32_t cols, float* out, void* stream);
/* per-segment sum (mean != 0: mean with count clamped to 1) in ascending row order (readout, schnet.py:115) */
int molsde_segment_reduce(const float* x, const int32_t* ptr, int32_t segments, int32_t cols, int32_t mean, float* out,
                          void* stream);

/* ------------------------------------------------------------------------------------
 * do_CL, metric EBM_node_dot_prod (examples/util.py:52-68), forward:
 *   pred_pos[r] = <X[r],Y[r]>/T, pred_neg[r] = <X[r],Y[perm[r]]>/T,
 *   loss_acc[0] = BCEWithLogits(pred_pos,1) + BCEWithLogits(pred_neg,0) (means), loss_acc[1] = CL_acc.
 * workspace: >= 4 * min(ceil(N/8), 592) floats.
 * ---------------------------------------------------------------------------------- */
int molsde_ebm_node_dot(const float* X, const float* Y, const int64_t* perm, int64_t N, int32_t D, float T,
                        float* pred_pos, float* pred_neg, float* loss_acc, float* workspace, int64_t workspace_floats,
                        void* stream);

/* ------------------------------------------------------------------------------------
 * Dense 3D->2D score networks (SDE_model_3D_to_2D_node_adj_dense.py, invariant_scorenetwork_dense.py,
 * layers/edge_network_dense.py, layers/node_network_dense.py) -- building blocks, Nm <= 64.
 * Layouts: x [B,Nm,F]; adj [B,Nm,Nm]; adjacency stacks [B,C,Nm,Nm]; pair features [B,Nm,Nm,C].
 * ---------------------------------------------------------------------------------- */
/* to_dense_batch (:130-131): out[b,a,:] = x[node_ptr[b]+a,:] or 0 */
int molsde_to_dense_batch(const float* x, const int32_t* node_ptr, int32_t B, int32_t Nm, int32_t F, float* out, int64_t ldo,
                          void* stream);
/* to_dense_adj (:129): adj[b,i,j] += val[e] + val_add  (val float or int64; :121 uses bond_type + 1) */
int molsde_to_dense_adj(const int64_t* edge_index, int64_t E, const float* val, const int64_t* val_i64, float val_add,
                        const int32_t* node_ptr, int32_t B, int32_t Nm, float* adj, void* stream);
/* node_flags (:523-529): flags[b,i] = sum_j |adj[b,i,j]| > eps */
int molsde_node_flags(const float* adj, int32_t B, int32_t Nm, float eps, float* flags, void* stream);
/* Y[:, g*No:(g+1)*No] = act(X[:, g*Ki:(g+1)*Ki] . W[g]^T + b[g]), W [G][No][Ki] */
int molsde_grouped_linear(const float* X, int64_t rows, int64_t ldx, const float* W, const float* b, int32_t G, int32_t Ki,
                          int32_t No, float* Y, int64_t ldy, int32_t act, void* stream);
/* pow_tensor with c_init = 2 (invariant_scorenetwork_dense.py:28-37) */
int molsde_dense_pow2(const float* adj, int32_t B, int32_t Nm, float* adjc, float* allc, int32_t ld_all, int32_t all_off,
                      void* stream);
/* NodeNetwork_dense.forward (node_network_dense.py:46-85) for C channels at once */
int molsde_dense_gcn(const float* adjc, int64_t adj_stride_b, int64_t adj_stride_c, int32_t B, int32_t C, int32_t Nm,
                     const float* xw, int64_t ldxw, const float* bias, int32_t Fo, float* out, int64_t ldo, int32_t out_off,
                     int32_t act, void* stream);
/* EdgeLayer tanh-attention + symmetrisation (edge_network_dense.py:66-80) -> pair[..., c], adjc copy -> pair[..., C+c] */
int molsde_dense_attn(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, const float* adjc, int32_t B,
                      int32_t C, int32_t Nm, float* pair, void* stream);
/* (m + m^T) masked (edge_network_dense.py:124-126) -> next adjacency stack + all-channels buffer */
int molsde_dense_pair_post(const float* m, const float* flags, int32_t B, int32_t Nm, int32_t Co, float* adjc_next, float* allc,
                           int32_t ld_all, int32_t all_off, void* stream);
/* zero diagonal, mask, optional per-graph scale (invariant_scorenetwork_dense.py:86-91; get_score_fn :83,93) */
int molsde_dense_edge_final(const float* raw, const float* flags, const float* scale, int32_t B, int32_t Nm, float* out,
                            void* stream);

/* Fused inference path of the edge score network (csrc/dense_fused.cu): channel-major stacks [B,C,Nm,Nm] end to end.
 *  dense_attn_sym:  S[b,c,i,j] = (A_ij + A_ji)/2 of EdgeLayer (edge_network_dense.py:66-80); pairs with f_i f_j = 0 get 0;
 *  dense_pair_mlp:  adjc_next[b,c',i,j] = ((m_ij + m_ji) f_j) f_i, m = MLP_elu(cat([S, adjc])[b,:,i,j]) (2Cin -> Hd -> Hd -> Co,
 *                   :120-126); `symmetric` != 0 asserts bitwise-symmetric inputs (every layer but the first) and evaluates m once;
 *  dense_edge_final_mlp: out[b,i,j] = MLP_silu(F -> H1 -> H2 -> 1)(all channels of the `nseg` stacks)[i,j] (i != j) f_i f_j scale[b]
 *                   (invariant_scorenetwork_dense.py:84-93; get_score_fn :83,93).  seg_ptrs / seg_channels are HOST arrays. */
int molsde_dense_attn_sym(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, const float* flags, int32_t B,
                          int32_t C, int32_t Nm, float* S, void* stream);
int molsde_dense_pair_mlp(const float* S, const float* adjc, const float* flags, const float* W0, const float* b0, const float* W1,
                          const float* b1, const float* W2, const float* b2, int32_t B, int32_t Cin, int32_t Hd, int32_t Co,
                          int32_t Nm, int32_t symmetric, float* adjc_next, void* stream);
/* Node-side chain of an EdgeLayer with a narrow input (Fin <= 16: every layer but the first) in one launch: h1 = tanh(W1 x + b1)
 * [G*W], qk group g = W2[g] h1[g*W:(g+1)*W] + b2 (G = 2C groups, W = 32), xw = Wv x [NV]  (edge_network_dense.py:45-53,
 * node_network_dense.py:73); and the multi_channel MLP (:113-118): out = tanh(W1 elu(W0 v + b0) + b1) * rowflag. */
int molsde_dense_node_side(const float* X, int64_t rows, int64_t ldx, int32_t Fin, const float* W1, const float* b1, const float* W2,
                           const float* b2, const float* Wv, int32_t G, int32_t W, int32_t NV, float* QK, int64_t ldqk, float* XW,
                           int64_t ldxw, void* stream);
int molsde_dense_multi_channel(const float* V, int64_t rows, int32_t K, const float* W0, const float* b0, int32_t H, const float* W1,
                               const float* b1, int32_t NO, const float* rowflag, float* out, void* stream);
int molsde_dense_edge_final_mlp(const float* const* seg_ptrs, const int32_t* seg_channels, int32_t nseg, const float* flags,
                                const float* scale, const float* W0, const float* b0, const float* W1, const float* b1,
                                const float* W2, const float* b2, int32_t F, int32_t H1, int32_t H2, int32_t B, int32_t Nm, float* out,
                                void* stream);

/* perturbation prologue / loss epilogue of SDEModel3Dto2D_node_adj_dense.forward (:134-152,160-179) and the elementwise
 * steps of the 3D->2D PC sampler (examples/pretrain_MoleculeSDE_inference_3D_to_2D_VE_VP.py:167-252) */
int molsde_dense_sym_noise(const float* raw, const float* flags, int32_t B, int32_t Nm, float* z, void* stream);
int molsde_dense_perturb_adj(const float* x, const float* z, const float* flags, const float* coef, const float* stdv, int32_t B,
                             int32_t Nm, float* out, void* stream);
int molsde_dense_perturb_onehot(const int64_t* zidx, const float* raw, const float* flags, const float* coef, const float* stdv,
                                int32_t B, int32_t Nm, int32_t K, float* zx, float* px, void* stream);
/* per-graph reductions over M contiguous floats: mode 0 Frobenius norm of a; mode 1 mean((a+b)^2) * w[g] */
int molsde_graph_reduce(const float* a, const float* b, const float* w, int32_t B, int64_t M, int32_t mode, float* out,
                        void* stream);
int molsde_langevin_step(const float* gnorm, const float* nnorm, const float* alpha, int32_t B, float snr, float* step,
                         void* stream);
int molsde_langevin_update(const float* x, const float* grad, const float* noise, const float* step, int32_t B, int64_t M,
                           float seps, float* x_new, float* x_mean, void* stream);
int molsde_reverse_update(const float* x, const float* score, const float* z, const float* sqrt_alpha, const float* G, int32_t B,
                          int64_t M, float* x_new, float* x_mean, void* stream);
int molsde_mask_rows(const float* x, const float* flags, int64_t rows, int32_t cols, float* out, void* stream);

/* ------------------------------------------------------------------------------------
 * Pretraining step (examples/pretrain_MoleculeSDE.py:105-152): layer-granular forward ops that keep their
 * intermediates, and the backward kernel of every op on the path (csrc/train*.cu).  fp32, deterministic.
 * The reference gets these from torch.autograd; a maintainer binds them as autograd.Function pairs.
 * ---------------------------------------------------------------------------------- */

/* C[M,N] (+)= op(A)[M,K] . op(B)[K,N];  transA: A stored [K][lda];  transB: B stored [N][ldb].
 * Backward of nn.Linear y = x W^T:  dx = dy . W (0,0);  dW = dy^T . x (1,0; split-K over the rows, workspace
 * molsde_gemm_ws_floats, may be NULL = no split). */
int64_t molsde_gemm_ws_floats(int64_t M, int64_t N, int64_t K);
int molsde_gemm(int32_t transA, int32_t transB, int64_t M, int64_t N, int64_t K, const float* A, int64_t lda, const float* B,
                int64_t ldb, float* C, int64_t ldc, int32_t accumulate, float* ws, int64_t ws_floats, void* stream);
/* out[n] (+)= sum_m X[m,n]  (bias gradients) */
int64_t molsde_colsum_ws_floats(int64_t M, int32_t N);
int molsde_colsum(const float* X, int64_t M, int32_t N, int64_t ldx, float* out, int32_t accumulate, float* ws, int64_t ws_floats,
                  void* stream);
/* y = act(x);  dx = dy * act'(x) with x the pre-activation (act codes of molsde_linear) */
int molsde_act_fwd(const float* x, int64_t n, int32_t act, float* y, void* stream);
int molsde_act_bwd(const float* x, const float* dy, int64_t n, int32_t act, float* dx, void* stream);
/* the same from the OUTPUT y = f(pre) for relu (1), shifted softplus (3), tanh (4), elu (5): dx = dy * f'(pre(y)); a linear layer
 * whose activation is one of these applies it in the GEMM epilogue and keeps one tensor (others: MOLSDE_ERR_UNSUPPORTED) */
int molsde_act_bwd_y(const float* y, const float* dy, int64_t n, int32_t act, float* dx, void* stream);
/* op 0: out = a + alpha*b (b NULL: alpha*a);  op 1: out = a*b (+c);  op 2: out[r,:] = a[r,:] * alpha * b[r];  op 3: out[r,c] = a[r,c] + b[c]  (cols per row);
 * op 4: out = a * (1 + b[0]) (+c).
 * out may alias a. */
int molsde_ew(int32_t op, const float* a, const float* b, const float* c, float alpha, int64_t n, int64_t cols, float* out,
              void* stream);
/* out[r,:] = A[ia[r],:] (+ B[ib[r],:]);  NULL index = identity  (x_j / x_i gathers of MessagePassing) */
int molsde_gather_pair(const float* A, const int32_t* ia, const float* B, const int32_t* ib, int64_t rows, int32_t cols, float* out,
                       void* stream);
/* out[s,:] (+)= scale[s] * sum_{p in [ptr[s],ptr[s+1])} X[perm ? perm[p] : p, :]  (scatter-add / backward of a gather,
 * as a deterministic gather-reduce over a CSR of the index) */
int molsde_seg_gather_sum(const float* X, const int32_t* ptr, const int32_t* perm, int64_t segments, int32_t cols, const float* scale,
                          int32_t accumulate, int32_t row_div, float* out, void* stream);  /* X row = perm[p] / row_div */
/* out[r,:] = sum_f T[keys[r*F+f],:]  (AtomEncoder / BondEncoder / nn.Embedding; keys include the table offsets) */
int molsde_embed_sum(const float* T, const int32_t* keys, int64_t rows, int32_t F, int32_t cols, float* out, void* stream);
/* out[s,:] = sum_{p in ptr[s]..ptr[s+1]} A[ia[e],:] * W[e,:], e = perm ? perm[p] : p   (CFConv message+aggregate, and its dx) */
int molsde_edge_mul_reduce(const float* A, const int32_t* ia, const float* W, const int32_t* ptr, const int32_t* perm, int64_t segments,
                           int32_t cols, float* out, void* stream);
/* out[e,:] = A[ia[e],:] * B[ib[e],:] */
int molsde_edge_mul_gather(const float* A, const int32_t* ia, const float* B, const int32_t* ib, int64_t E, int32_t cols, float* out,
                           void* stream);
/* the same two with a leading dimension: W is a column block of a wider [E, ldw] filter stack (all SchNet interactions' filters
 * side by side, schnet.py:185-195 evaluated once per batch), out a column block of the matching gradient stack [E, ldo];
 * escale (optional, [E]): the cosine cutoff C(d_e) (schnet.py:187-188) applied on the fly -- reduce uses W[e,:] * escale[e], gather
 * returns (A[ia[e],:] * B[ib[e],:]) * escale[e], i.e. the gradient with respect to the UNSCALED filter */
int molsde_edge_mul_reduce_ld(const float* A, const int32_t* ia, const float* W, int64_t ldw, const float* escale, const int32_t* ptr,
                              const int32_t* perm, int64_t segments, int32_t cols, float* out, void* stream);
int molsde_edge_mul_gather_ld(const float* A, const int32_t* ia, const float* B, const int32_t* ib, const float* escale, int64_t E,
                              int32_t cols, float* out, int64_t ldo, void* stream);
/* Whole-chain kernels of the narrow 3-layer MLPs applied to the B*Nm^2 atom pairs in TRAINING (EdgeNetwork_dense.mlp,
 * edge_network_dense.py:120-123): y = W3 act(W2 act(W1 x + b1)
 * + b2) + b3, thread = row, hidden vectors in registers, fp32 FFMA.  fwd keeps the pre-activations p1, p2 [rows, h]; bwd runs the
 * whole input-gradient chain and leaves a1 = act(p1), a2 = act(p2), d1, d2 [rows, h] (the operands of the three weight-gradient
 * GEMMs) and dx [rows, d0] (row stride lddx; NULL = not wanted).  x row stride ldx; y, dy dense [rows, d3].
 * act: 2 silu, 5 elu.  Only the (d0, h, d3, act) combinations reported by molsde_mlp3_train_supported exist (others:
 * MOLSDE_ERR_UNSUPPORTED; the host keeps the layer-granular path). */
int molsde_mlp3_train_supported(int32_t d0, int32_t h, int32_t d3, int32_t act);   /* 1 / 0 */
int molsde_mlp3_train_fwd(const float* x, int64_t rows, int64_t ldx, int32_t d0, int32_t h, int32_t d3, int32_t act, const float* W1,
                          const float* b1, const float* W2, const float* b2, const float* W3, const float* b3, float* p1, float* p2,
                          float* y, void* stream);
int molsde_mlp3_train_bwd(const float* p1, const float* p2, const float* dy, int64_t rows, int32_t d0, int32_t h, int32_t d3, int32_t act,
                          const float* W1, const float* W2, const float* W3, float* a1, float* a2, float* d1, float* d2, float* dx,
                          int64_t lddx, void* stream);
/* out[0] (+)= alpha <a,b>;  ws: >= 128 doubles */
int molsde_dot(const float* a, const float* b, int64_t n, float alpha, int32_t accumulate, float* out, double* ws, void* stream);
/* do_CL, metric InfoNCE_dot_prod (examples/util.py:23-32): rows of logits [B,B] = X Y^T / T (a molsde_tc_gemm):
 * loss_row[r] = logsumexp(logits[r,:]) - logits[r,r], correct_row[r] = (argmax == r); with write_grad the logits are replaced
 * in place by (softmax - onehot) * grad_scale, from which dX = dlogits . Y and dY = dlogits^T . X are two more GEMMs. */
int molsde_infonce_rows(float* logits, int64_t B, int64_t ld, float grad_scale, int32_t write_grad, float* loss_row, float* correct_row,
                        void* stream);
/* GINConv (molecule_gnn_model.py:13-32): pre = (1+eps) x + sum_{e->i} relu(x_src + BondEncoder(e)); dmsg = dpre[tgt] * relu' */
int molsde_gin_aggregate_fwd(const float* x, const float* T, const int32_t* ekeys, int32_t F, const int32_t* rowptr, const int32_t* src,
                             const float* eps, int64_t N, int32_t cols, float* pre, void* stream);
int molsde_gin_message_bwd(const float* x, const float* T, const int32_t* ekeys, int32_t F, const int32_t* src, const int32_t* tgt,
                           const float* dpre, int64_t E, int32_t cols, float* dmsg, void* stream);
/* GaussianSmearing ea [E,ng] and cosine cutoff C [E] of the radius edges (schnet.py:93,186,205-207) */
int molsde_schnet_edge_feat(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu, int32_t ng,
                            float coeff, float cutoff, float* ea, float* C, void* stream);
/* Backward of molsde_schnet_edge_feat towards the positions (the reference obtains forces as -dE/dpos through autograd,
 * examples/finetune_MD17.py:66): g[e,:] = (sum_k dea[e,k] d ea[e,k]/dd + dC[e] dC/dd) * (pos[src]-pos[tgt]) / d, the contribution
 * of edge e to d loss / d pos[src] (and, negated, to d pos[tgt]); dC may be NULL.  molsde_rowdot: out[r] (+)= <a[r,:], b[r,:]>. */
int molsde_schnet_edge_feat_bwd(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu, int32_t ng,
                                float coeff, float cutoff, const float* ea, const float* dea, const float* dC, float* g, void* stream);
int molsde_rowdot(const float* a, const float* b, int64_t rows, int32_t cols, int32_t accumulate, float* out, void* stream);
/* Second-order pieces of SchNet (a force term INSIDE the training loss: examples/finetune_MD17.py:66-77 differentiates
 * pred_force = -grad(E, pos, create_graph=True) with respect to the parameters).  The host propagates a forward-mode tangent
 * along the position displacement v = d loss / d force through the network (the same GEMM / CFConv / activation kernels) and
 * back-propagates through primal + tangent; the two kernels it needs beyond the first-order set:
 *   act_bwd2:                out (+)= act''(x) * a * b          (x = pre-activation)
 *   schnet_edge_feat_tangent: ea_dot[e,k], C_dot[e] = d/d eps of GaussianSmearing / cosine cutoff at pos + eps v (schnet.py:185-188) */
int molsde_act_bwd2(const float* x, const float* a, const float* b, int64_t n, int32_t act, int32_t accumulate, float* out, void* stream);
int molsde_schnet_edge_feat_tangent(const float* pos, const float* v, const int32_t* src, const int32_t* tgt, int64_t E, const float* mu,
                                    int32_t ng, float coeff, float cutoff, const float* ea, float* ea_dot, float* C_dot, void* stream);
/* backward of molsde_ebm_node_dot's loss_acc[0] scaled by coef; invperm = inverse permutation of perm */
int molsde_ebm_node_dot_bwd(const float* X, const float* Y, const int64_t* perm, const int64_t* invperm, const float* pred_pos,
                            const float* pred_neg, int64_t N, int32_t D, float T, float coef, int32_t accumulate, float* dX, float* dY,
                            void* stream);
/* stable counting sort of n int64 keys in [0,buckets): count[b]; then (after an exclusive scan -> rowptr) perm */
int molsde_bucket_count(const int64_t* keys, int64_t n, int32_t buckets, int32_t* count, void* stream);
int molsde_bucket_fill(const int64_t* keys, int64_t n, int32_t buckets, const int32_t* rowptr, int32_t* perm, void* stream);
int molsde_expand_rowptr(const int32_t* rowptr, int64_t N, int32_t* row, void* stream);
/* nn.LayerNorm over the last dim; bwd returns dx and dyx = dy * xhat (dgamma = colsum(dyx), dbeta = colsum(dy)) */
int molsde_layernorm_fwd(const float* x, int64_t M, int32_t D, const float* g, const float* b, float eps, float* y, float* mean,
                         float* rstd, void* stream);
int molsde_layernorm_bwd(const float* x, const float* dy, int64_t M, int32_t D, const float* g, const float* mean, const float* rstd,
                         float* dx, float* dyx, void* stream);
/* nn.BatchNorm1d in train mode over rows [M,F] (+ optional fused ReLU, act = 1); running stats updated when given */
int64_t molsde_bn_ws_doubles(int64_t M, int32_t F);
int molsde_bn_train_fwd(const float* x, int64_t M, int32_t F, const float* gamma, const float* beta, float eps, float momentum,
                        float* running_mean, float* running_var, int32_t act, float* y, float* mean, float* rstd, double* ws,
                        void* stream);
int molsde_bn_eval(const float* x, int64_t M, int32_t F, const float* gamma, const float* beta, const float* running_mean,
                   const float* running_var, float eps, int32_t act, float* y, float* rstd_tmp, void* stream);
int molsde_bn_train_bwd(const float* x, const float* dy, int64_t M, int32_t F, const float* gamma, const float* mean,
                        const float* rstd, float* dx, float* dgamma, float* dbeta, double* ws, void* stream);
/* the same with the ReLU mask of a fused BatchNorm+ReLU applied in place (relu_y = the forward output; NULL = none) and the
 * totals also ADDED to the parameter-gradient buffers grad_gamma / grad_beta (optional) -- one call instead of
 * act_bwd + bn_train_bwd + two accumulation passes. */
int molsde_bn_train_bwd_fused(const float* x, const float* dy, const float* relu_y, int64_t M, int32_t F, const float* gamma,
                              const float* mean, const float* rstd, float* dx, float* dgamma, float* dbeta, float* grad_gamma,
                              float* grad_beta, double* ws, void* stream);
/* torch.optim.Adam step over one flat buffer (pretrain_MoleculeSDE.py:337); g is scaled by grad_scale first (1/world) */
int molsde_adam_step(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
                     float weight_decay, int32_t step, float grad_scale, void* stream);

/* SDEModel2Dto3D_02 training ops; edges in CSR-by-target order (tgt = expanded rowptr, src = plan src):
 *  edge_geom: Fourier features of distance / frame coefficients, pseudo angle (emb[:,0:2], ld 66), frame basis [E,9]
 *             (SDE_model_2D_to_3D.py:342-366);
 *  tconv_fwd/bwd: TransformerConv 8x4 message passing on qkvs [N,128] = [q|k|v|skip], eproj [E,32]; keep [E,8] or NULL;
 *             bwd fills dqkvs [N,128], dkvE [E,64] (scratch) and deproj [E,32]; sptr/sperm = CSR by source;
 *  equi_fwd/bwd: EquiLayer mean aggregation of sum_k dyn_k basis_k (equivariant_scorenetwork.py:43-78);
 *  dsm_pos_loss_bwd: gradient of molsde_dsm_pos_loss's mean_out w.r.t. score. */
int molsde_sde2d3d_edge_geom(const float* pos, const int32_t* src, const int32_t* tgt, int64_t E, const float* w_dist,
                             const float* w_coff, float* gfd, float* gfi, float* gfj, float* emb, float* basis, void* stream);
int molsde_tconv_fwd(const float* qkvs, const float* eproj, const int32_t* rowptr, const int32_t* src, int64_t N, const float* keep,
                     float dropout_p, float* alpha, float* out, void* stream);
int molsde_tconv_bwd(const float* qkvs, const float* eproj, const int32_t* rowptr, const int32_t* src, const int32_t* sptr,
                     const int32_t* sperm, int64_t N, const float* keep, float dropout_p, const float* alpha, const float* dout,
                     float* dqkvs, float* dkvE, float* deproj, void* stream);
int molsde_equi_fwd(const float* dyn, const float* basis, const int32_t* rowptr, int64_t N, int32_t accumulate, float* grad,
                    void* stream);
int molsde_equi_bwd(const float* dgrad, const float* basis, const int32_t* rowptr, const int32_t* tgt, int64_t E, float* ddyn,
                    void* stream);
int molsde_dsm_pos_loss_bwd(const float* score, const float* noise, const float* w, const int32_t* node_ptr, const int32_t* node2graph,
                            int64_t N, int32_t B, float upstream, float* dscore, void* stream);

/* Fused row-wise 3-layer MLP  Y = W2 act(W1 act(W0 x + b0) + b1) + b2  (nn.Linear weights [out,in]; K0 <= 32, H1,H2 <= 64,
 * NO <= 8; act 2 silu / 4 tanh / 5 elu) for the B*Nm^2-row per-pair MLPs of EdgeNetwork_dense (edge_network_dense.py:120-123) and
 * the final head of EdgeScoreNetwork_dense (invariant_scorenetwork_dense.py:84-86): one pass over HBM instead of three. */
int molsde_mlp3_rows(const float* X, int64_t rows, int64_t ldx, int32_t K0, const float* W0, const float* b0, int32_t H1,
                     const float* W1, const float* b1, int32_t H2, const float* W2, const float* b2, int32_t NO, int32_t act,
                     float* Y, int64_t ldy, void* stream);

/* fp32-accurate GEMM on tcgen05 tensor cores (3xTF32 split, TMEM accumulator; csrc/tc_gemm.cu):
 *   C[M,N] (+)= A . B^T (+ bias[n]) -> * rowscale[m] -> act (+ R),   A(m,k) = A[m*sam + k*sak],  B(n,k) = B[n*sbn + k*sbk]
 * (one stride of each operand must be 1).  nn.Linear forward: A = x (sam = ldx, sak = 1), B = W [out,in] (sbn = in, sbk = 1);
 * dx = dy . W: B = W with sbn = 1, sbk = in;  dW = dy^T . x: A = dy (sam = 1, sak = ldy), B = x (sbn = 1, sbk = ldx), split-K
 * over the rows when `ws` (molsde_tc_gemm_ws_floats) is given.  status (optional): set to 1 on an internal wait time-out. */
int64_t molsde_tc_gemm_ws_floats(int64_t M, int64_t N, int64_t K);
int molsde_tc_gemm(int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn, int64_t sbk,
                   const float* bias, int32_t act, const float* rowscale, const float* R, int64_t ldr, float* C, int64_t ldc,
                   int32_t accumulate, float* ws, int64_t ws_floats, int32_t* status, void* stream);
/* nn.Linear backward, weight and bias gradient in one GEMM: dW[M=out,N=in] (+)= dy^T x (A = dy: sam = 1, sak = ldy; B = x:
 * sbn = 1, sbk = ldx; K = rows) and db[m] (+)= sum_rows dy[row,m] via an all-ones extra operand row.
 * Workspace: molsde_tc_gemm_ws_floats(M, N + 1, K). */
int molsde_tc_gemm_dw_db(int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbn,
                         int64_t sbk, float* dW, int64_t ldc, float* db, int32_t accumulate, float* ws, int64_t ws_floats,
                         int32_t* status, void* stream);
/* `batch` independent GEMMs of one shape in ONE launch; pointers of batch b are offset by b*bsA / b*bsB / b*bsC / b*bsBias
 * elements (0 = shared operand).  Instances: the per-channel func_q / func_k / func_v layers of EdgeNetwork_dense
 * (layers/edge_network_dense.py:105-128) and their dx / dW. */
int64_t molsde_tc_gemm_batched_ws_floats(int32_t batch, int64_t M, int64_t N, int64_t K);
int molsde_tc_gemm_batched(int32_t batch, int64_t M, int64_t N, int64_t K, const float* A, int64_t sam, int64_t sak, int64_t bsA,
                           const float* B, int64_t sbn, int64_t sbk, int64_t bsB, const float* bias, int64_t bsBias, int32_t act,
                           float* C, int64_t ldc, int64_t bsC, int32_t accumulate, float* ws, int64_t ws_floats, int32_t* status,
                           void* stream);

/* Backward kernels of the dense 3D->2D score networks (forward: molsde_dense_*):
 *  dense_gcn_bwd: dpre [B*Nm, C*Fo] = dout * act'(out) (its column sum = dbias), dxw [B*Nm, lddx], and (dadj != NULL) the
 *                 gradient w.r.t. the off-diagonal adjacency entries [B,C,Nm,Nm] (Fo <= 16; act none or tanh);
 *  dense_attn_bwd: dQ/dK in the qk layout of the forward, dadjc = pass-through part of dpair (may be NULL);
 *  dense_pair_post_bwd: dm [B*Nm*Nm, Co] from dadjc_next (NULL for the last layer) + the layer's slice of dallc;
 *  dense_edge_final_bwd, graph_mse_bwd (mode-1 molsde_graph_reduce followed by the mean over graphs, times coef),
 *  from_dense_batch (backward of molsde_to_dense_batch). */
int molsde_dense_gcn_bwd(const float* adjc, int64_t adj_stride_b, int64_t adj_stride_c, int32_t B, int32_t C, int32_t Nm,
                         const float* xw, int64_t ldxw, int32_t Fo, const float* out, const float* dout, int64_t ldo, int32_t out_off,
                         int32_t act, float* dpre, float* dxw, int64_t lddx, float* dadj, int64_t dadj_stride_b, int32_t dadj_accumulate,
                         void* stream);
int molsde_dense_attn_bwd(const float* Q, const float* K, int64_t ldq, int32_t W, int32_t ds, int32_t B, int32_t C, int32_t Nm,
                          const float* dpair, float* dQ, float* dK, float* dadjc, void* stream);
int molsde_dense_pair_post_bwd(const float* dadjc_next, const float* dallc, int32_t ld_all, int32_t all_off, const float* flags,
                               int32_t B, int32_t Nm, int32_t Co, float* dm, void* stream);
int molsde_dense_edge_final_bwd(const float* dout, const float* flags, const float* scale, int32_t B, int32_t Nm, float* draw,
                                void* stream);
int molsde_graph_mse_bwd(const float* a, const float* b, const float* w, int32_t B, int64_t M, float coef, float* da, void* stream);
/* dst[r,0:cols] (+)= src[r,0:cols] with independent row strides;  out[0] = mean(v[0:n]) */
int molsde_copy2d(const float* src, int64_t lds, float* dst, int64_t ldd, int64_t rows, int32_t cols, int32_t accumulate, void* stream);
int molsde_mean(const float* v, int64_t n, float* out, void* stream);
int molsde_sum_slices(const float* X, int32_t slices, int64_t n, float* out, void* stream);  /* out[i] = sum_s X[s*n+i] */
int molsde_from_dense_batch(const float* dense, int64_t ldd, const int32_t* node_ptr, const int32_t* node2graph, int64_t N, int32_t Nm,
                            int32_t F, float* x, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MOLSDE_B200_H_ */
