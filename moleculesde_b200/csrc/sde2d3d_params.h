/* Float offsets inside the packed SDEModel2Dto3D_02 parameter blob (molsde_sde2d3d_params.blob).
 * The host (moleculesde_b200/sde_2d_to_3d.py::packed_params) builds it from the reference
 * state_dict keys.
 *
 * Every GEMM of the score network runs on tcgen05; its weight is stored as a B-operand tile:
 *      fp16, two-way split  w = hi + lo  (hi = fp16(w), lo = fp16(w - hi)), each part stored as an [N rows][K] K-major tile in
 *      the canonical no-swizzle core-matrix layout (8 rows x 16 B = 8 halves):
 *        byte offset(n, k) = (k/8)*(N*16) + (n/8)*128 + (n%8)*16 + (k%8)*2        (LBO = N*16 B, SBO = 128 B)
 * All section offsets are multiples of 32 floats (128 B); every section is copied to shared memory by ONE TMA bulk copy. */
#ifndef MOLSDE_SDE2D3D_PARAMS_H_
#define MOLSDE_SDE2D3D_PARAMS_H_

/* ---- per-edge feature stage (SDE_model_2D_to_3D.py:402-432) ---- */
#define MOLSDE_P_GFP_DIST_W 0   /* dist_gaussian_fourier.W [32] */
#define MOLSDE_P_GFP_COFF_W 32  /* coff_gaussian_fourier.W [32] */
#define MOLSDE_P_E0_HV 64       /* [32][4]: {fused hidden bias (project.0.bias + P_i b_c + P_j b_c), project.0.weight[:,0] (pseudo_sin),
                                   project.0.weight[:,1] (pseudo_cos), 0} per hidden unit */
#define MOLSDE_P_E0_OB 192      /* [32][2]: {input_mlp.layers.0.bias, project.layers.1.bias} per output column */
#define MOLSDE_P_E0_BT 256      /* 11 B tiles of [32 n][32 k] fp16, each 1024 floats = hi (2048 B) | lo (2048 B):
                                   tiles 0..9 = Fourier sub-blocks b = 2*blk + half (blk 0: input_mlp.layers.0.weight over gfp(d);
                                   blk 1..4: (project.0.weight[:,2:] (.) coff_mlp.weight) over gfp(ci0, ci2, cj0, cj2)), K order inside a
                                   sub-block = [sin f(16*half .. +15) | cos f(16*half .. +15)];  tile 10 = project.layers.1.weight */
#define MOLSDE_E0_BT_FLOATS 1024
#define MOLSDE_P_E0_END 11520

/* ---- one GATLayer (score_network.gnn_layers.{m}.{c}), base = P_GAT0 + (2m+c)*P_GAT_SZ ----
 * All GEMMs of the layer run on tcgen05: B tiles [N rows][32 k] fp16, hi | lo.
 * [0, G_WP_SZ): resident for the whole layer;  [G_WQKVS, P_GAT_SZ): only needed by the q|k|v|skip GEMM (staged over the edge-phase buffers) */
#define MOLSDE_P_GAT0 11520
#define MOLSDE_G_F0C 0      /* FFN.0.weight [32 n][32 k]: hi (512 floats) | lo (512 floats) */
#define MOLSDE_G_F3C 1024   /* FFN.3.weight */
#define MOLSDE_G_WEC 2048   /* MHA.lin_edge.weight (no bias) */
#define MOLSDE_G_BQKVS 3072 /* [128]: lin_query | lin_key | lin_value | lin_skip biases */
#define MOLSDE_G_LN1_W 3200
#define MOLSDE_G_LN1_B 3232
#define MOLSDE_G_F0_B 3264
#define MOLSDE_G_F3_B 3296
#define MOLSDE_G_LN2_W 3328
#define MOLSDE_G_LN2_B 3360
#define MOLSDE_G_WP_SZ 3392
#define MOLSDE_G_WQKVS 3392 /* [lin_query | lin_key | lin_value | lin_skip].weight as ONE B tile [128 n][32 k]: hi (2048 floats) | lo (2048) */
#define MOLSDE_P_GAT_SZ 7488

/* ---- one basis MLP (score_network.basis_mlp_modules.{m}), base = P_BASIS0 + m*P_BASIS_SZ ----
 * layer 0 (64 -> 128; input k 0..31 = h_row+h_col, 32..63 = edge_attr) as a tcgen05 B tile [128 n][64 k] fp16 hi | lo. */
#define MOLSDE_P_BASIS0 41472
#define MOLSDE_B_W1_HI 0     /* 16384 B = 4096 floats */
#define MOLSDE_B_W1_LO 4096
#define MOLSDE_B_EPI 8192    /* [128][4]: {.0.bias[n], .2.weight[0][n], .2.weight[1][n], .2.weight[2][n]} */
#define MOLSDE_B_B2 8704     /* .2.bias [3] + 1 pad */
#define MOLSDE_P_BASIS_SZ 8708
#define MOLSDE_P_BASIS_STRIDE 8736  /* section stride (multiple of 32 floats) */

#define MOLSDE_P_TOTAL 58944

#endif
