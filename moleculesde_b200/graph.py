"""Host wrappers of the graph-construction kernels (csrc/graph.cu).

`extend_graph`  replaces `Geom3D/datasets/dataset_3D.py:12-35` (torch_sparse spspmm+coalesce),
`radius_graph`  replaces `torch_cluster.radius_graph` as called at `Geom3D/models/schnet.py:91`,
`csr_by_target` gives the kernels the CSR view (grouped by target, input order kept) in which
`MessagePassing.propagate` accumulates.  All three return int32 CSR for the kernels plus the
reference-layout int64 `[2,E]` tensor.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr


@dataclass
class CSR:
    rowptr: torch.Tensor  # int32 [N+1], grouped by target
    col: torch.Tensor     # int32 [E], source node of every edge (ascending inside a row for the builders)
    edge_index: Optional[torch.Tensor] = None  # int64 [2,E] in the reference layout
    perm: Optional[torch.Tensor] = None        # int32 [E] position in the input edge list (csr_by_target only)

    @property
    def num_edges(self) -> int:
        return int(self.col.numel())


def segment_ptr(keys: torch.Tensor, num_segments: int, indirect: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Offsets of an ascending key vector (`batch`) -> int32 [num_segments+1]."""
    require_device(keys)
    assert keys.dtype == torch.long
    M = int(indirect.numel()) if indirect is not None else int(keys.numel())
    out = torch.empty(num_segments + 1, dtype=torch.int32, device=keys.device)
    check(lib().molsde_segment_ptr(ptr(keys), ptr(indirect), M, num_segments, ptr(out), stream_ptr(keys)), "segment_ptr")
    return out


def _scan(deg: torch.Tensor) -> torch.Tensor:
    out = torch.empty(deg.numel() + 1, dtype=torch.int32, device=deg.device)
    check(lib().molsde_exclusive_scan_i32(ptr(deg), deg.numel(), ptr(out), stream_ptr(deg)), "exclusive_scan")
    return out


def _check_mol_sizes(node_ptr: torch.Tensor) -> None:
    sizes = node_ptr[1:] - node_ptr[:-1]
    if sizes.numel() and int(sizes.max().item()) > _abi.MAX_MOL_NODES:
        raise _abi.MolsdeError(f"molecule with more than {_abi.MAX_MOL_NODES} atoms is not supported by the graph kernels")


def batch_ptrs(batch: torch.Tensor, edge_index: torch.Tensor, num_graphs: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """(node_ptr, edge_ptr) int32 [B+1] of a collated batch (edges grouped by molecule)."""
    node_ptr = segment_ptr(batch, num_graphs)
    edge_ptr = segment_ptr(batch, num_graphs, indirect=edge_index[0].contiguous())
    return node_ptr, edge_ptr


def extend_graph(edge_index: torch.Tensor, batch: torch.Tensor, num_graphs: int, want_edge_index: bool = True) -> CSR:
    """<=4-hop extended graph of every molecule of a batch.  Symmetric, so the
    row-major (coalesce) order the reference stores is also CSR-by-target with ascending sources:
    entry k = (a, b) of the reference tensor is read by the kernels as target a, source b."""
    require_device(edge_index)
    edge_index = edge_index.contiguous()
    N, E_b = int(batch.numel()), int(edge_index.size(1))
    node_ptr, edge_ptr = batch_ptrs(batch, edge_index, num_graphs)
    _check_mol_sizes(node_ptr)
    deg = torch.empty(N, dtype=torch.int32, device=batch.device)
    s = stream_ptr(batch)
    check(lib().molsde_extend_graph_count(ptr(edge_index), E_b, ptr(node_ptr), ptr(edge_ptr), num_graphs, ptr(deg), s),
          "extend_graph_count")
    rowptr = _scan(deg)
    E_x = int(rowptr[-1].item())  # the one host sync of graph construction (output size)
    col = torch.empty(E_x, dtype=torch.int32, device=batch.device)
    ext = torch.empty(2, E_x, dtype=torch.long, device=batch.device) if want_edge_index else None
    check(lib().molsde_extend_graph_fill(ptr(edge_index), E_b, ptr(node_ptr), ptr(edge_ptr), num_graphs, ptr(rowptr),
                                         E_x, ptr(col), ptr(ext), s), "extend_graph_fill")
    return CSR(rowptr, col, ext)


def radius_graph(pos: torch.Tensor, r: float, batch: torch.Tensor, num_graphs: int, max_num_neighbors: int = 32,
                 want_edge_index: bool = True) -> CSR:
    """Neighbour list with torch_cluster's CUDA semantics (see include/molsde_b200.h)."""
    require_device(pos)
    pos = pos.contiguous().float()
    N = int(pos.size(0))
    node_ptr = segment_ptr(batch, num_graphs)
    _check_mol_sizes(node_ptr)
    deg = torch.empty(N, dtype=torch.int32, device=pos.device)
    s = stream_ptr(pos)
    check(lib().molsde_radius_graph_count(ptr(pos), ptr(node_ptr), num_graphs, float(r), max_num_neighbors, ptr(deg), s),
          "radius_graph_count")
    rowptr = _scan(deg)
    E = int(rowptr[-1].item())
    col = torch.empty(E, dtype=torch.int32, device=pos.device)
    ei = torch.empty(2, E, dtype=torch.long, device=pos.device) if want_edge_index else None
    check(lib().molsde_radius_graph_fill(ptr(pos), ptr(node_ptr), num_graphs, float(r), max_num_neighbors, ptr(rowptr), E,
                                         ptr(col), ptr(ei), s), "radius_graph_fill")
    return CSR(rowptr, col, ei)


def edge_segments(edge_index: torch.Tensor, num_nodes: int) -> Tuple[torch.Tensor, torch.Tensor, int]:
    """(node_ptr, edge_ptr, G) of a collated `[2,E]` edge list WITHOUT the `batch` vector (the reference's
    `GNN.forward(x, edge_index, edge_attr)` form, `molecule_gnn_model.py:160-163`): position e of the list is a cut when every
    node referenced before it is smaller than every node referenced from it on (molecules are contiguous node blocks whose edges
    are stored together), so each run between cuts is a self-contained sub-graph and the per-graph CSR kernels apply.  A few
    [E]-sized prefix-max / suffix-min passes and one size read; nodes without edges join the preceding segment."""
    E, dev = int(edge_index.size(1)), edge_index.device
    if E == 0:
        z = torch.tensor([0, num_nodes], dtype=torch.int32, device=dev)
        return z, torch.zeros(2, dtype=torch.int32, device=dev), 1
    lo, hi = torch.minimum(edge_index[0], edge_index[1]), torch.maximum(edge_index[0], edge_index[1])
    pm = torch.cummax(hi, 0).values
    sm = torch.flip(torch.cummin(torch.flip(lo, [0]), 0).values, [0])
    cut = torch.ones(E + 1, dtype=torch.bool, device=dev)
    cut[1:E] = pm[:-1] < sm[1:]
    edge_ptr = torch.nonzero(cut).reshape(-1)                       # [G+1], ascending, first 0, last E
    G = int(edge_ptr.numel()) - 1
    node_ptr = torch.empty(G + 1, dtype=torch.int64, device=dev)
    node_ptr[:G] = sm[edge_ptr[:G]]
    node_ptr[0] = 0
    node_ptr[G] = num_nodes
    return node_ptr.to(torch.int32).contiguous(), edge_ptr.to(torch.int32).contiguous(), G


def csr_by_target(edge_index: torch.Tensor, batch: Optional[torch.Tensor], num_graphs: int, num_nodes: Optional[int] = None) -> CSR:
    """CSR-by-target of an arbitrary collated `[2,E]` edge list, stable in input order.  `batch=None` (with `num_nodes`): the
    sub-graph boundaries are recovered from the edge list itself (`edge_segments`)."""
    require_device(edge_index)
    edge_index = edge_index.contiguous()
    E = int(edge_index.size(1))
    if batch is None:
        assert num_nodes is not None
        N = int(num_nodes)
        node_ptr, edge_ptr, num_graphs = edge_segments(edge_index, N)
        batch = edge_index   # device / stream carrier below
    else:
        N = int(batch.numel())
        node_ptr, edge_ptr = batch_ptrs(batch, edge_index, num_graphs)
    deg = torch.empty(N, dtype=torch.int32, device=batch.device)
    s = stream_ptr(batch)
    check(lib().molsde_csr_by_target_count(ptr(edge_index), E, N, ptr(deg), s), "csr_count")
    rowptr = _scan(deg)
    src = torch.empty(E, dtype=torch.int32, device=batch.device)
    perm = torch.empty(E, dtype=torch.int32, device=batch.device)
    check(lib().molsde_csr_by_target_fill(ptr(edge_index), E, ptr(node_ptr), ptr(edge_ptr), num_graphs, ptr(rowptr),
                                          ptr(src), ptr(perm), s), "csr_fill")
    return CSR(rowptr, src, edge_index, perm)
