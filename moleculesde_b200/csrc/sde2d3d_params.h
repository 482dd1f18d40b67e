/* Float offsets inside the packed SDEModel2Dto3D_02 parameter blob (molsde_sde2d3d_params.blob).
 * The host (moleculesde_b200/sde_2d_to_3d.py::packed_params) builds it from the reference
 * state_dict keys.  Every `*_W` matrix is stored k-major (`[in][ld]`, i.e. the
 * nn.Linear weight transposed) with a padded leading dimension ld == 8 (mod 32) so that the
 * mma.sync B-fragment loads (lane -> (k = lane%4, n = lane/4)) are shared-memory bank-conflict free.
 * All offsets are multiples of 4 floats (16 B, cp.async granularity). */
#ifndef MOLSDE_SDE2D3D_PARAMS_H_
#define MOLSDE_SDE2D3D_PARAMS_H_

#define MOLSDE_LD32 40   /* leading dimension of a 32-column weight block  */
#define MOLSDE_LD96 104  /* q|k|v block                                     */
#define MOLSDE_LD128 136 /* 128-column weight block                         */

/* ---- per-edge feature stage (SDE_model_2D_to_3D.py:402-432) ---- */
#define MOLSDE_P_GFP_DIST_W 0   /* dist_gaussian_fourier.W [32] */
#define MOLSDE_P_GFP_COFF_W 32  /* coff_gaussian_fourier.W [32] */
#define MOLSDE_P_IN_B 64        /* input_mlp.layers.0.bias [32] */
#define MOLSDE_P_H_B 96         /* fused bias: project.0.bias + P_i b_c + P_j b_c [32] */
#define MOLSDE_P_H_WSIN 128     /* project.layers.0.weight[:,0] (pseudo_sin) [32] */
#define MOLSDE_P_H_WCOS 160     /* project.layers.0.weight[:,1] (pseudo_cos) [32] */
#define MOLSDE_P_P1_B 192       /* project.layers.1.bias [32] */
#define MOLSDE_P_IN_W 224       /* input_mlp.layers.0.weight^T [64][40] */
#define MOLSDE_P_H_W 2784       /* (project.0.weight[:,2:34] @ coff_mlp.weight)^T rows 0..127,
                                   (project.0.weight[:,34:66] @ coff_mlp.weight)^T rows 128..255; [256][40] */
#define MOLSDE_P_P1_W 13024     /* project.layers.1.weight^T [32][40] */
#define MOLSDE_P_E0_END 14304

/* ---- one GATLayer (score_network.gnn_layers.{m}.{c}), base = P_GAT0 + (2m+c)*P_GAT_SZ ---- */
#define MOLSDE_P_GAT0 14304
#define MOLSDE_G_WQKV 0     /* [lin_query | lin_key | lin_value].weight^T [32][104] */
#define MOLSDE_G_WS 3328    /* MHA.lin_skip.weight^T [32][40] */
#define MOLSDE_G_WE 4608    /* MHA.lin_edge.weight^T [32][40] (no bias) */
#define MOLSDE_G_F0 5888    /* FFN.0.weight^T [32][40] */
#define MOLSDE_G_F3 7168    /* FFN.3.weight^T [32][40] */
#define MOLSDE_G_BQKV 8448  /* [96] */
#define MOLSDE_G_BS 8544
#define MOLSDE_G_LN1_W 8576
#define MOLSDE_G_LN1_B 8608
#define MOLSDE_G_F0_B 8640
#define MOLSDE_G_F3_B 8672
#define MOLSDE_G_LN2_W 8704
#define MOLSDE_G_LN2_B 8736
#define MOLSDE_P_GAT_SZ 8768

/* ---- one basis MLP (score_network.basis_mlp_modules.{m}), base = P_BASIS0 + m*P_BASIS_SZ ----
 * The first layer (64 -> 128; input rows 0..31 = h_row+h_col, 32..63 = edge_attr) runs on tcgen05: its weight is
 * stored as the B operand tile [N=128][K=64], K-major, in the canonical no-swizzle core-matrix layout
 *   float index(n, k) = (k/4)*512 + (n/8)*32 + (n%8)*4 + (k%4)        (LBO = 2048 B, SBO = 128 B)
 * twice: the tf32 "hi" part (top 19 bits) and the exact remainder "lo" (3xTF32 split done on the host). */
#define MOLSDE_P_BASIS0 49376
#define MOLSDE_B_W1C_HI 0      /* [8192] */
#define MOLSDE_B_W1C_LO 8192   /* [8192] */
#define MOLSDE_B_B1 16384      /* .0.bias [128] */
#define MOLSDE_B_W2 16512      /* .2.weight [3][128] */
#define MOLSDE_B_B2 16896      /* .2.bias [3] + 1 pad */
#define MOLSDE_P_BASIS_SZ 16900

#define MOLSDE_P_TOTAL 83176

#endif
