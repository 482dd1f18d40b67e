"""ctypes binding of the C ABI in `include/molsde_b200.h` (the only way Python reaches the kernels).

The library is built in-tree by `__graft_entry__.build()` / `moleculesde_b200.build.build()` as
`moleculesde_b200/libmolsde_b200.so`.  There is no fallback: if the library is missing or the
device is not sm_100, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_float, c_int32, c_int64, c_uint64, c_void_p
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MOLSDE_LIB_PATH") or os.path.join(_HERE, "libmolsde_b200.so")   # (override: A/B of two builds)

# keep in sync with include/molsde_b200.h
MAX_MOL_NODES = 128
CHUNK_MAX_NODES = 224
TILE_EDGES = 128
TILE_LD = 136                      # SchNet edge tiles (csrc/schnet.cu)
TILE_FLOATS = 32 * TILE_EDGES      # edge_2D_emb tile [8 feature quads][128 slots][4] (molsde_tile_floats())
HID = 32
EMB = 300
MAX_CHUNK_TILES = 64

EXPORTS = [
    "molsde_version", "molsde_last_error_string", "molsde_check_device",
    "molsde_segment_ptr", "molsde_exclusive_scan_i32",
    "molsde_extend_graph_count", "molsde_extend_graph_fill",
    "molsde_radius_graph_count", "molsde_radius_graph_fill",
    "molsde_csr_by_target_count", "molsde_csr_by_target_fill",
    "molsde_linear",
    "molsde_edge2d_emb_eval", "molsde_sde2d3d_score", "molsde_sde2d3d_scratch_floats", "molsde_tile_floats",
    "molsde_sde2d3d_pc_sample", "molsde_edge2d_bn_train", "molsde_sde2d3d_forward_net", "molsde_dsm_pos_loss", "molsde_perturb_rows",
    "molsde_gemm_ws_floats", "molsde_gemm", "molsde_colsum_ws_floats", "molsde_colsum", "molsde_act_fwd", "molsde_act_bwd", "molsde_act_bwd_y",
    "molsde_ew", "molsde_gather_pair", "molsde_seg_gather_sum", "molsde_bucket_count", "molsde_bucket_fill",
    "molsde_expand_rowptr", "molsde_layernorm_fwd", "molsde_layernorm_bwd", "molsde_bn_ws_doubles", "molsde_bn_train_fwd",
    "molsde_bn_train_bwd", "molsde_adam_step", "molsde_sde2d3d_edge_geom", "molsde_tconv_fwd", "molsde_tconv_bwd",
    "molsde_equi_fwd", "molsde_equi_bwd", "molsde_dsm_pos_loss_bwd", "molsde_embed_sum", "molsde_edge_mul_reduce",
    "molsde_edge_mul_gather", "molsde_edge_mul_reduce_ld", "molsde_edge_mul_gather_ld", "molsde_mlp3_train_supported", "molsde_mlp3_train_fwd", "molsde_mlp3_train_bwd", "molsde_dot", "molsde_gin_aggregate_fwd", "molsde_gin_message_bwd", "molsde_schnet_edge_feat",
    "molsde_schnet_edge_feat_bwd", "molsde_rowdot", "molsde_bn_train_bwd_fused",
    "molsde_ebm_node_dot_bwd", "molsde_infonce_rows", "molsde_bn_eval", "molsde_dense_gcn_bwd", "molsde_dense_attn_bwd",
    "molsde_dense_pair_post_bwd", "molsde_dense_edge_final_bwd", "molsde_graph_mse_bwd", "molsde_from_dense_batch", "molsde_copy2d", "molsde_mean", "molsde_sum_slices", "molsde_tc_gemm_ws_floats", "molsde_tc_gemm", "molsde_tc_gemm_batched_ws_floats", "molsde_tc_gemm_batched", "molsde_tc_gemm_dw_db", "molsde_mlp3_rows",
    "molsde_schnet_cfconv", "molsde_gather_rows", "molsde_segment_reduce", "molsde_ebm_node_dot",
    "molsde_to_dense_batch", "molsde_to_dense_adj", "molsde_node_flags", "molsde_grouped_linear", "molsde_dense_pow2",
    "molsde_dense_gcn", "molsde_dense_attn", "molsde_dense_pair_post", "molsde_dense_edge_final",
    "molsde_dense_sym_noise", "molsde_dense_perturb_adj", "molsde_dense_perturb_onehot", "molsde_graph_reduce",
    "molsde_langevin_step", "molsde_langevin_update", "molsde_reverse_update", "molsde_mask_rows",
    "molsde_dense_attn_sym", "molsde_dense_pair_mlp", "molsde_dense_edge_final_mlp", "molsde_dense_node_side", "molsde_dense_multi_channel",
    "molsde_sde2d3d_pc_corrector_update", "molsde_sde2d3d_pc_predictor_update", "molsde_act_bwd2", "molsde_schnet_edge_feat_tangent", "molsde_build_plan_host", "molsde_debug_echo",
]


class MolsdeError(RuntimeError):
    pass


class Plan(Structure):
    _fields_ = [("num_chunks", c_int32), ("num_tiles", c_int32), ("N", c_int64), ("E", c_int64),
                ("chunk_tile_ptr", c_void_p), ("tile_tgt_ptr", c_void_p), ("rowptr", c_void_p), ("src", c_void_p),
                ("chunk_order", c_void_p)]


class Params(Structure):
    _fields_ = [("blob", c_void_p), ("blob_floats", c_int64)]


class PCConfig(Structure):
    _fields_ = [("steps", c_int32), ("snr", c_float), ("scale_eps", c_float), ("seed", c_uint64)]


_lib: Optional[ctypes.CDLL] = None

# Kernels that update parameters or buffers IN PLACE through raw pointers (the flat Adam step, the train-mode BatchNorm running
# statistics) do not bump torch's per-tensor `_version`.  Every such wrapper calls `touch_params()`; every packed-weight cache
# (SDEModel2Dto3D_02.packed_params, SchNet._filters, EdgeScoreNetwork_dense._pack) carries `param_epoch()` in its key.
_param_epoch = 0


def param_epoch() -> int:
    return _param_epoch


def touch_params() -> None:
    global _param_epoch
    _param_epoch += 1


def lib() -> ctypes.CDLL:
    """Load (once) and return the CUDA library; raises MolsdeError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise MolsdeError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). moleculesde_b200 has no CPU fallback.")
    L = ctypes.CDLL(LIB_PATH)
    L.molsde_version.restype = c_char_p
    L.molsde_last_error_string.restype = c_char_p
    L.molsde_check_device.argtypes = [c_int32]
    L.molsde_segment_ptr.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.molsde_exclusive_scan_i32.argtypes = [c_void_p, c_int64, c_void_p, c_void_p]
    L.molsde_extend_graph_count.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p]
    L.molsde_extend_graph_fill.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_int64,
                                           c_void_p, c_void_p, c_void_p]
    L.molsde_radius_graph_count.argtypes = [c_void_p, c_void_p, c_int32, c_float, c_int32, c_void_p, c_void_p]
    L.molsde_radius_graph_fill.argtypes = [c_void_p, c_void_p, c_int32, c_float, c_int32, c_void_p, c_int64,
                                           c_void_p, c_void_p, c_void_p]
    L.molsde_csr_by_target_count.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p]
    L.molsde_csr_by_target_fill.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_int32, c_void_p, c_void_p,
                                            c_void_p, c_void_p]
    L.molsde_linear.argtypes = [c_void_p, c_int64, c_int32, c_int64, c_void_p, c_void_p, c_int32, c_void_p,
                                c_int64, c_int32, c_void_p, c_int64, c_void_p, c_void_p]
    L.molsde_to_dense_batch.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p]
    L.molsde_to_dense_adj.argtypes = [c_void_p, c_int64, c_void_p, c_void_p, c_float, c_void_p, c_int32, c_int32, c_void_p,
                                      c_void_p]
    L.molsde_node_flags.argtypes = [c_void_p, c_int32, c_int32, c_float, c_void_p, c_void_p]
    L.molsde_grouped_linear.argtypes = [c_void_p, c_int64, c_int64, c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                        c_int64, c_int32, c_void_p]
    L.molsde_dense_pow2.argtypes = [c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p]
    L.molsde_dense_gcn.argtypes = [c_void_p, c_int64, c_int64, c_int32, c_int32, c_int32, c_void_p, c_int64, c_void_p, c_int32,
                                   c_void_p, c_int64, c_int32, c_int32, c_void_p]
    L.molsde_dense_attn.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32,
                                    c_void_p, c_void_p]
    L.molsde_dense_pair_post.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_int32, c_int32,
                                         c_void_p]
    L.molsde_dense_edge_final.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    L.molsde_dense_sym_noise.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    L.molsde_dense_perturb_adj.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]
    L.molsde_dense_perturb_onehot.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int32, c_int32,
                                              c_void_p, c_void_p, c_void_p]
    L.molsde_graph_reduce.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_int32, c_void_p, c_void_p]
    L.molsde_langevin_step.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_float, c_void_p, c_void_p]
    L.molsde_langevin_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_float, c_void_p, c_void_p,
                                         c_void_p]
    L.molsde_reverse_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_int64, c_void_p, c_void_p,
                                        c_void_p]
    L.molsde_mask_rows.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.molsde_dense_attn_sym.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_int32, c_int32, c_int32, c_void_p,
                                        c_void_p]
    L.molsde_dense_node_side.argtypes = [c_void_p, c_int64, c_int64, c_int32, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int32,
                                         c_int32, c_int32, c_void_p, c_int64, c_void_p, c_int64, c_void_p]
    L.molsde_dense_multi_channel.argtypes = [c_void_p, c_int64, c_int32, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p,
                                             c_void_p, c_void_p]
    L.molsde_dense_pair_mlp.argtypes = [c_void_p] * 9 + [c_int32] * 6 + [c_void_p, c_void_p]
    L.molsde_dense_edge_final_mlp.argtypes = [POINTER(c_void_p), POINTER(c_int32), c_int32] + [c_void_p] * 8 + [c_int32] * 5 + \
                                             [c_void_p, c_void_p]
    L.molsde_schnet_cfconv.argtypes = [POINTER(Plan), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_int32, c_float, c_float, c_void_p, c_void_p]
    L.molsde_gather_rows.argtypes = [c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.molsde_segment_reduce.argtypes = [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]
    L.molsde_ebm_node_dot.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_float, c_void_p, c_void_p,
                                      c_void_p, c_void_p, c_int64, c_void_p]
    L.molsde_edge2d_emb_eval.argtypes = [POINTER(Plan), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.molsde_sde2d3d_score.argtypes = [POINTER(Plan), POINTER(Params), c_void_p, c_void_p, c_void_p, c_void_p,
                                       c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.molsde_sde2d3d_scratch_floats.argtypes = [POINTER(Plan), c_int32, POINTER(c_int32)]
    L.molsde_sde2d3d_scratch_floats.restype = c_int64
    L.molsde_tile_floats.restype = c_int64
    P = c_void_p
    L.molsde_gemm_ws_floats.restype = c_int64
    L.molsde_gemm_ws_floats.argtypes = [c_int64, c_int64, c_int64]
    L.molsde_gemm.argtypes = [c_int32, c_int32, c_int64, c_int64, c_int64, P, c_int64, P, c_int64, P, c_int64, c_int32, P, c_int64, P]
    L.molsde_colsum_ws_floats.restype = c_int64
    L.molsde_colsum_ws_floats.argtypes = [c_int64, c_int32]
    L.molsde_colsum.argtypes = [P, c_int64, c_int32, c_int64, P, c_int32, P, c_int64, P]
    L.molsde_act_fwd.argtypes = [P, c_int64, c_int32, P, P]
    L.molsde_act_bwd.argtypes = [P, P, c_int64, c_int32, P, P]
    L.molsde_act_bwd_y.argtypes = [P, P, c_int64, c_int32, P, P]
    L.molsde_ew.argtypes = [c_int32, P, P, P, c_float, c_int64, c_int64, P, P]
    L.molsde_gather_pair.argtypes = [P, P, P, P, c_int64, c_int32, P, P]
    L.molsde_seg_gather_sum.argtypes = [P, P, P, c_int64, c_int32, P, c_int32, c_int32, P, P]
    L.molsde_embed_sum.argtypes = [P, P, c_int64, c_int32, c_int32, P, P]
    L.molsde_edge_mul_reduce.argtypes = [P, P, P, P, P, c_int64, c_int32, P, P]
    L.molsde_edge_mul_gather.argtypes = [P, P, P, P, c_int64, c_int32, P, P]
    L.molsde_edge_mul_reduce_ld.argtypes = [P, P, P, c_int64, P, P, P, c_int64, c_int32, P, P]
    L.molsde_edge_mul_gather_ld.argtypes = [P, P, P, P, P, c_int64, c_int32, P, c_int64, P]
    L.molsde_mlp3_train_supported.argtypes = [c_int32, c_int32, c_int32, c_int32]
    L.molsde_mlp3_train_fwd.argtypes = [P, c_int64, c_int64, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, P, P, P, P, P]
    L.molsde_mlp3_train_bwd.argtypes = [P, P, P, c_int64, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P, P, P, P, c_int64, P]
    L.molsde_dot.argtypes = [P, P, c_int64, c_float, c_int32, P, P, P]
    L.molsde_gin_aggregate_fwd.argtypes = [P, P, P, c_int32, P, P, P, c_int64, c_int32, P, P]
    L.molsde_gin_message_bwd.argtypes = [P, P, P, c_int32, P, P, P, c_int64, c_int32, P, P]
    L.molsde_schnet_edge_feat.argtypes = [P, P, P, c_int64, P, c_int32, c_float, c_float, P, P, P]
    L.molsde_schnet_edge_feat_bwd.argtypes = [P, P, P, c_int64, P, c_int32, c_float, c_float, P, P, P, P, P]
    L.molsde_rowdot.argtypes = [P, P, c_int64, c_int32, c_int32, P, P]
    L.molsde_ebm_node_dot_bwd.argtypes = [P, P, P, P, P, P, c_int64, c_int32, c_float, c_float, c_int32, P, P, P]
    L.molsde_bucket_count.argtypes = [P, c_int64, c_int32, P, P]
    L.molsde_bucket_fill.argtypes = [P, c_int64, c_int32, P, P, P]
    L.molsde_expand_rowptr.argtypes = [P, c_int64, P, P]
    L.molsde_layernorm_fwd.argtypes = [P, c_int64, c_int32, P, P, c_float, P, P, P, P]
    L.molsde_layernorm_bwd.argtypes = [P, P, c_int64, c_int32, P, P, P, P, P, P]
    L.molsde_bn_ws_doubles.restype = c_int64
    L.molsde_bn_ws_doubles.argtypes = [c_int64, c_int32]
    L.molsde_bn_train_fwd.argtypes = [P, c_int64, c_int32, P, P, c_float, c_float, P, P, c_int32, P, P, P, P, P]
    L.molsde_dense_gcn_bwd.argtypes = [P, c_int64, c_int64, c_int32, c_int32, c_int32, P, c_int64, c_int32, P, P, c_int64, c_int32,
                                       c_int32, P, P, c_int64, P, c_int64, c_int32, P]
    L.molsde_tc_gemm_ws_floats.restype = c_int64
    L.molsde_tc_gemm_ws_floats.argtypes = [c_int64, c_int64, c_int64]
    L.molsde_tc_gemm.argtypes = [c_int64, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, P, c_int32, P, P, c_int64, P,
                                 c_int64, c_int32, P, c_int64, P, P]
    L.molsde_tc_gemm_dw_db.argtypes = [c_int64, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, c_int64, P, c_int64, P, c_int32, P, c_int64,
                                       P, P]
    L.molsde_mlp3_rows.argtypes = [P, c_int64, c_int64, c_int32, P, P, c_int32, P, P, c_int32, P, P, c_int32, c_int32, P, c_int64, P]
    L.molsde_tc_gemm_batched_ws_floats.restype = c_int64
    L.molsde_tc_gemm_batched_ws_floats.argtypes = [c_int32, c_int64, c_int64, c_int64]
    L.molsde_tc_gemm_batched.argtypes = [c_int32, c_int64, c_int64, c_int64, P, c_int64, c_int64, c_int64, P, c_int64, c_int64, c_int64,
                                         P, c_int64, c_int32, P, c_int64, c_int64, c_int32, P, c_int64, P, P]
    L.molsde_copy2d.argtypes = [P, c_int64, P, c_int64, c_int64, c_int32, c_int32, P]
    L.molsde_mean.argtypes = [P, c_int64, P, P]
    L.molsde_sum_slices.argtypes = [P, c_int32, c_int64, P, P]
    L.molsde_dense_attn_bwd.argtypes = [P, P, c_int64, c_int32, c_int32, c_int32, c_int32, c_int32, P, P, P, P, P]
    L.molsde_dense_pair_post_bwd.argtypes = [P, P, c_int32, c_int32, P, c_int32, c_int32, c_int32, P, P]
    L.molsde_dense_edge_final_bwd.argtypes = [P, P, P, c_int32, c_int32, P, P]
    L.molsde_graph_mse_bwd.argtypes = [P, P, P, c_int32, c_int64, c_float, P, P]
    L.molsde_from_dense_batch.argtypes = [P, c_int64, P, P, c_int64, c_int32, c_int32, P, P]
    L.molsde_infonce_rows.argtypes = [P, c_int64, c_int64, c_float, c_int32, P, P, P]
    L.molsde_bn_eval.argtypes = [P, c_int64, c_int32, P, P, P, P, c_float, c_int32, P, P, P]
    L.molsde_bn_train_bwd.argtypes = [P, P, c_int64, c_int32, P, P, P, P, P, P, P, P]
    L.molsde_bn_train_bwd_fused.argtypes = [P, P, P, c_int64, c_int32, P, P, P, P, P, P, P, P, P, P]
    L.molsde_adam_step.argtypes = [P, P, P, P, c_int64, c_float, c_float, c_float, c_float, c_float, c_int32, c_float, P]
    L.molsde_sde2d3d_edge_geom.argtypes = [P, P, P, c_int64, P, P, P, P, P, P, P, P]
    L.molsde_tconv_fwd.argtypes = [P, P, P, P, c_int64, P, c_float, P, P, P]
    L.molsde_tconv_bwd.argtypes = [P, P, P, P, P, P, c_int64, P, c_float, P, P, P, P, P, P]
    L.molsde_equi_fwd.argtypes = [P, P, P, c_int64, c_int32, P, P]
    L.molsde_equi_bwd.argtypes = [P, P, P, P, c_int64, P, P]
    L.molsde_dsm_pos_loss_bwd.argtypes = [P, P, P, P, P, c_int64, c_int32, c_float, P, P]
    L.molsde_edge2d_bn_train.argtypes = [POINTER(Plan), c_void_p, c_int32, c_void_p, c_void_p, c_float, c_float, c_void_p, c_void_p,
                                         c_void_p, c_void_p, c_void_p]
    L.molsde_sde2d3d_forward_net.argtypes = [POINTER(Plan), POINTER(Params), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                             c_float, c_void_p, c_void_p, c_int64, c_void_p, c_void_p]
    L.molsde_perturb_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_void_p, c_void_p]
    L.molsde_dsm_pos_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_void_p]
    L.molsde_sde2d3d_pc_sample.argtypes = [POINTER(Plan), POINTER(Params), c_void_p, c_void_p, c_void_p, c_void_p,
                                           POINTER(PCConfig), c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                           c_int64, c_void_p, c_void_p, c_void_p]
    L.molsde_debug_echo.argtypes = [c_int64, c_float, c_int32, c_void_p, c_float, c_int64, c_int32, c_uint64, c_int64, c_float, c_int64,
                                    c_int32, c_void_p]
    L.molsde_build_plan_host.argtypes = [c_void_p, c_void_p, c_int32, c_void_p, c_int32, c_int32, c_int32, c_int32, c_void_p, c_void_p, c_void_p]
    L.molsde_act_bwd2.argtypes = [c_void_p, c_void_p, c_void_p, c_int64, c_int32, c_int32, c_void_p, c_void_p]
    L.molsde_schnet_edge_feat_tangent.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int32, c_float, c_float,
                                                  c_void_p, c_void_p, c_void_p, c_void_p]
    L.molsde_sde2d3d_pc_corrector_update.argtypes = [c_void_p, c_void_p, c_void_p, c_int32, c_void_p, c_void_p, c_float, c_float, c_uint64,
                                                     c_void_p, c_int64, c_void_p]
    L.molsde_sde2d3d_pc_predictor_update.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_uint64, c_void_p, c_int64, c_void_p]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is ctypes.c_int:
            fn.restype = c_int32
    _lib = _fast_lib(L)
    return _lib


class _FastLib:
    """The CDLL with its status-returning entry points rebound through `_molsde_fastcall` (csrc/fastcall.c): ~0.3 us per call
    instead of ~3 us of ctypes marshalling -- the eager training step makes ~750 calls per iteration.  Entry points that take
    ctypes structures / arrays by reference, return something else than the int status, or have no declared argtypes stay on ctypes."""

    def __init__(self, cdll, fc):
        self._cdll = cdll
        plain = {c_void_p: "i", c_int32: "i", c_int64: "i", c_uint64: "i", c_float: "f"}
        for name in EXPORTS:
            fn = getattr(cdll, name)
            at = fn.argtypes
            if at is None or fn.restype is not c_int32 or any(t not in plain for t in at):
                continue
            sig = "".join(plain[t] for t in at)
            if sig.count("i") > 28 or sig.count("f") > 8:
                continue
            setattr(self, name, fc.bind(ctypes.cast(fn, c_void_p).value, sig))

    def __getattr__(self, name):   # anything not rebound: the ctypes function
        return getattr(self._cdll, name)


def _fast_lib(cdll):
    if os.environ.get("MOLSDE_NO_FASTCALL") == "1":
        return cdll
    try:
        from . import _molsde_fastcall as fc
    except ImportError:
        return cdll   # the extension is an optimisation of the HOST path only; ctypes reaches the same kernels
    return _FastLib(cdll, fc)


_device_checked = set()


def require_device(t: torch.Tensor) -> None:
    """Fail loudly unless `t` lives on an sm_100 CUDA device."""
    if not t.is_cuda:
        raise MolsdeError("moleculesde_b200 kernels need CUDA tensors (no CPU fallback)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx not in _device_checked:
        check(lib().molsde_check_device(idx), "check_device")
        _device_checked.add(idx)


def check(status: int, what: str) -> None:
    if status != 0:
        msg = lib().molsde_last_error_string().decode()
        raise MolsdeError(f"{what} failed with status {status}: {msg}")


# Layout checks of every pointer that crosses the C ABI (contiguity / unit column stride).  On in the test suite (conftest sets
# MOLSDE_CHECK_ABI=1); off by default: the eager training step makes ~4,000 pointer conversions, and is bound by host time.
CHECK_ABI = os.environ.get("MOLSDE_CHECK_ABI", "0") == "1"


def ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    if t is None:
        return None
    if CHECK_ABI:
        assert t.is_contiguous(), "C ABI needs contiguous tensors"
    return t.data_ptr()


def stream_ptr(t: torch.Tensor) -> int:
    return torch.cuda.current_stream(t.device).cuda_stream
