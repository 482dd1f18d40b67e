/* Float offsets inside the packed SDEModel2Dto3D_02 parameter blob (molsde_sde2d3d_params.blob).
 * The host (moleculesde_b200/sde_2d_to_3d.py::pack_params) builds it from the reference
 * state_dict keys; every `*_WT` block is the nn.Linear weight transposed to
 * [in][out] (k-major) so the kernels stream it as the B operand.  All offsets are multiples of 4
 * floats (16 B, cp.async granularity). */
#ifndef MOLSDE_SDE2D3D_PARAMS_H_
#define MOLSDE_SDE2D3D_PARAMS_H_

#define MOLSDE_P_GFP_DIST_W 0    /* dist_gaussian_fourier.W [32] */
#define MOLSDE_P_GFP_COFF_W 32   /* coff_gaussian_fourier.W [32] */
#define MOLSDE_P_IN_WT 64        /* input_mlp.layers.0.weight^T [64][32] */
#define MOLSDE_P_IN_B 2112       /* [32] */
#define MOLSDE_P_COFF_WT 2144    /* coff_mlp.weight^T [128][32] */
#define MOLSDE_P_COFF_B 6240     /* [32] */
#define MOLSDE_P_PROJ0_WT 6272   /* project.layers.0.weight^T [68][32] (rows 66,67 zero) */
#define MOLSDE_P_PROJ0_B 8448    /* [32] */
#define MOLSDE_P_PROJ1_WT 8480   /* project.layers.1.weight^T [32][32] */
#define MOLSDE_P_PROJ1_B 9504    /* [32] */
#define MOLSDE_P_E0_END 9536

/* one GATLayer (score_network.gnn_layers.{m}.{c}), base = MOLSDE_P_GAT0 + (2*m+c)*MOLSDE_P_GAT_SZ */
#define MOLSDE_P_GAT0 9536
#define MOLSDE_G_WQ_T 0      /* MHA.lin_query.weight^T [32][32] */
#define MOLSDE_G_WK_T 1024   /* MHA.lin_key */
#define MOLSDE_G_WV_T 2048   /* MHA.lin_value */
#define MOLSDE_G_WS_T 3072   /* MHA.lin_skip */
#define MOLSDE_G_BQ 4096
#define MOLSDE_G_BK 4128
#define MOLSDE_G_BV 4160
#define MOLSDE_G_BS 4192
#define MOLSDE_G_WE_T 4224   /* MHA.lin_edge.weight^T [32][32] (no bias) */
#define MOLSDE_G_LN1_W 5248
#define MOLSDE_G_LN1_B 5280
#define MOLSDE_G_F0_WT 5312  /* FFN.0 */
#define MOLSDE_G_F0_B 6336
#define MOLSDE_G_F3_WT 6368  /* FFN.3 */
#define MOLSDE_G_F3_B 7392
#define MOLSDE_G_LN2_W 7424
#define MOLSDE_G_LN2_B 7456
#define MOLSDE_P_GAT_SZ 7488

/* one basis MLP (score_network.basis_mlp_modules.{m}), base = MOLSDE_P_BASIS0 + m*MOLSDE_P_BASIS_SZ */
#define MOLSDE_P_BASIS0 39488
#define MOLSDE_B_W1_T 0      /* .0.weight^T [64][128]: rows 0..31 act on h_row+h_col, 32..63 on edge_attr */
#define MOLSDE_B_B1 8192     /* [128] */
#define MOLSDE_B_W2 8320     /* .2.weight [3][128] */
#define MOLSDE_B_B2 8704     /* [3] + 1 pad */
#define MOLSDE_P_BASIS_SZ 8708

#define MOLSDE_P_TOTAL 56904

#endif
