"""Checkpoint I/O and model construction with the reference's conventions.

* `save_model` / `load_model`: the `model_complete.pth` layout of `examples/pretrain_MoleculeSDE.py:71-103` — one dict with the
  four state_dicts under `model_2D`, `model_3D`, `SDE_2Dto3D_model`, `SDE_3Dto2D_model`.  Our modules keep the reference's
  parameter names, so a file written here loads into the reference classes and the published checkpoints
  (`README_checkpoints.md:9-49`) load here key for key (`strict=True`).
* `resolve_sde_type`: the noise-schedule presets the pretraining script hard-codes per direction
  (`pretrain_MoleculeSDE.py:226-256` for 2D->3D, `:272-302` for 3D->2D): "VE02" -> ("VE", 0.1, 10) etc.
* `build_models`: the four modules of `pretrain_MoleculeSDE.py:181-315` from the script's argument names.

Host-side bookkeeping only (no kernels); torch.save / torch.load do the serialisation as in the reference.
"""
from __future__ import annotations

import os
from typing import Dict, Optional, Tuple

import torch
from torch import nn

KEYS = ("model_2D", "model_3D", "SDE_2Dto3D_model", "SDE_3Dto2D_model")

# (SDE family, beta/sigma min, beta/sigma max); num_diffusion_timesteps is 1000 for every preset
_PRESETS_2D3D = {"VE": ("VE", 0.2, 1.0), "VP": ("VP", 0.2, 1.0), "VE02": ("VE", 0.1, 10.0), "VP02": ("VP", 0.2, 30.0),
                 "VE03": ("VE", 0.1, 1000.0), "VP03": ("VP", 0.2, 1000.0)}
_PRESETS_3D2D = {"VE": ("VE", 0.1, 1.0), "VP": ("VP", 0.2, 1.0), "VE02": ("VE", 0.1, 10.0), "VP02": ("VP", 0.1, 30.0),
                 "VE03": ("VE", 0.1, 1000.0), "VP03": ("VP", 0.1, 1000.0)}


def resolve_sde_type(name: str, direction: str) -> Tuple[str, float, float, int]:
    """`--SDE_type_2Dto3D` / `--SDE_type_3Dto2D` value -> (SDE_type, beta_min, beta_max, num_diffusion_timesteps)."""
    table = {"2Dto3D": _PRESETS_2D3D, "3Dto2D": _PRESETS_3D2D}[direction]
    if name not in table:
        raise NotImplementedError(f"SDE type {name!r} for {direction} (reference presets: {sorted(table)})")
    kind, lo, hi = table[name]
    return kind, lo, hi, 1000


def build_models(emb_dim: int = 300, SDE_type_2Dto3D: str = "VE", SDE_type_3Dto2D: str = "VE",
                 SDE_2Dto3D_model: str = "SDEModel2Dto3D_02", SDE_3Dto2D_model: str = "SDEModel3Dto2D_node_adj_dense",
                 use_extend_graph: bool = True, noise_on_one_hot: bool = True, num_layer: int = 5, JK: str = "last",
                 dropout_ratio: float = 0.0, gnn_type: str = "GIN", num_filters: int = 128, num_interactions: int = 6,
                 num_gaussians: int = 51, cutoff: float = 10.0, readout: str = "mean") -> Dict[str, nn.Module]:
    """The module set of `pretrain_MoleculeSDE.py:181-315`, keyed like the checkpoint."""
    from .gnn import GNN
    from .schnet import SchNet
    from . import sde_2d_to_3d as M23
    from . import sde_3d_to_2d as M32
    if SDE_2Dto3D_model not in ("SDEModel2Dto3D_01", "SDEModel2Dto3D_02"):
        raise NotImplementedError(f"{SDE_2Dto3D_model}: only the _01 and _02 variants are selectable in the reference script (:258-271)")
    if SDE_3Dto2D_model not in ("SDEModel3Dto2D_node_adj_dense", "SDEModel3Dto2D_node_adj_dense_02"):
        raise NotImplementedError(f"{SDE_3Dto2D_model}: the _dense_03 ablation is out of scope (DESIGN.md section 7)")
    k23, lo23, hi23, n23 = resolve_sde_type(SDE_type_2Dto3D, "2Dto3D")
    k32, lo32, hi32, n32 = resolve_sde_type(SDE_type_3Dto2D, "3Dto2D")
    return {
        "model_2D": GNN(num_layer, emb_dim, JK=JK, drop_ratio=dropout_ratio, gnn_type=gnn_type),
        "model_3D": SchNet(hidden_channels=emb_dim, num_filters=num_filters, num_interactions=num_interactions,
                           num_gaussians=num_gaussians, cutoff=cutoff, readout=readout, node_class=119),
        "SDE_2Dto3D_model": getattr(M23, SDE_2Dto3D_model)(
            emb_dim=emb_dim, hidden_dim=32, beta_schedule=None, beta_min=lo23, beta_max=hi23, num_diffusion_timesteps=n23,
            SDE_type=k23, use_extend_graph=use_extend_graph),
        "SDE_3Dto2D_model": getattr(M32, SDE_3Dto2D_model)(
            dim3D=emb_dim, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=emb_dim,
            num_linears=3, beta_min=lo32, beta_max=hi32, num_diffusion_timesteps=n32, SDE_type=k32, num_class_X=119,
            noise_on_one_hot=noise_on_one_hot),
    }


def save_model(models: Dict[str, nn.Module], output_model_dir: str, save_best: bool = True) -> Optional[str]:
    """`save_model(save_best)` of `pretrain_MoleculeSDE.py:71-103`: `model_complete.pth` (best) / `model_complete_final.pth`."""
    if not output_model_dir:
        return None
    os.makedirs(output_model_dir, exist_ok=True)
    path = os.path.join(output_model_dir, "model_complete.pth" if save_best else "model_complete_final.pth")
    torch.save({k: {n: v.detach().cpu() for n, v in models[k].state_dict().items()} for k in KEYS}, path)
    return path


def variant_of(state_dict: Dict[str, torch.Tensor]) -> str:
    """Which 2D->3D class a `SDE_2Dto3D_model` state_dict belongs to (the published checkpoints mix `_01` and `_02`)."""
    return "SDEModel2Dto3D_02" if "input_mlp.layers.0.weight" in state_dict else "SDEModel2Dto3D_01"


def variant_3d2d_of(state_dict: Dict[str, torch.Tensor]) -> str:
    """Which 3D->2D class a `SDE_3Dto2D_model` state_dict belongs to: `_dense_02` feeds its score networks the 2 x dim3D-wide
    concatenation (`SDE_model_3D_to_2D_node_adj_dense.py:225-236`), visible in the first GCN weight of the node network."""
    w = state_dict["node_score_network.layers.0.weight"]
    f = state_dict["embedding_3D.weight"].shape[0]
    return "SDEModel3Dto2D_node_adj_dense_02" if w.shape[0] == 2 * f else "SDEModel3Dto2D_node_adj_dense"


def load_model(path: str, models: Optional[Dict[str, nn.Module]] = None, strict: bool = True, **build_kwargs) -> Dict[str, nn.Module]:
    """Load a reference-format checkpoint.  Without `models`, the module set is built first (`build_models(**build_kwargs)`),
    picking the 2D->3D variant from the keys present in the file.  Entries missing from the file (the fine-tuning scripts
    store only `model` / `model_3D`, `finetune_QM9.py:97-105`) are left untouched."""
    blob = torch.load(path, map_location="cpu", weights_only=True)
    if models is None:
        if "SDE_2Dto3D_model" in blob:
            build_kwargs.setdefault("SDE_2Dto3D_model", variant_of(blob["SDE_2Dto3D_model"]))
        if "SDE_3Dto2D_model" in blob:
            build_kwargs.setdefault("SDE_3Dto2D_model", variant_3d2d_of(blob["SDE_3Dto2D_model"]))
        models = build_models(**build_kwargs)
    for k in KEYS:
        if k in blob:
            models[k].load_state_dict(blob[k], strict=strict)
    return models
