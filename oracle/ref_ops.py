"""Restatement (pure torch, CPU, loop-clear) of the third-party ops on the hot path.

TEST INFRASTRUCTURE.  The reference (`/root/reference`) calls these through
un-vendored wheels pinned only in its README (`README.md:35-47`: pytorch 1.9.1,
pyg 2.0.2 -> torch_scatter ~2.0.9, torch_sparse ~0.6.12, torch_cluster ~1.5.9).
None of them is installed here, so their published algorithms are restated; the
reference's own call sites are cited per function.  Parity at this boundary is
UNPINNED by the reference (it has no tests); `tests/test_ref_ops.py` checks every
function against a brute-force dense/loop implementation.
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
from torch import Tensor


# ----------------------------------------------------------------------------
# torch_scatter
# ----------------------------------------------------------------------------
def scatter(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None,
            reduce: str = "sum") -> Tensor:
    """torch_scatter.scatter(src, index, dim, dim_size=, reduce=) for a 1-D index along `dim`.

    Call sites: `Geom3D/models/schnet.py:106,115`,
    `Geom3D/models/MoleculeSDE/SDE_model_3D_to_2D_node_adj_dense.py:126`.
    Accumulation order is ascending position in `src` (torch_scatter CPU loop /
    `index_add_` CPU), 'mean' divides by the count clamped to >= 1.
    """
    if dim < 0:
        dim += src.dim()
    assert index.dim() == 1 and index.numel() == src.size(dim)
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = list(src.shape)
    shape[dim] = dim_size
    if reduce in ("sum", "add"):
        out = torch.zeros(shape, dtype=src.dtype, device=src.device)
        return out.index_add_(dim, index, src)
    if reduce == "mean":
        out = torch.zeros(shape, dtype=src.dtype, device=src.device).index_add_(dim, index, src)
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp_(min=1)
        view = [1] * src.dim()
        view[dim] = dim_size
        if src.is_floating_point():
            return out / cnt.view(view)
        return torch.div(out, cnt.view(view), rounding_mode="floor")
    if reduce == "max":
        out = torch.full(shape, float("-inf"), dtype=src.dtype, device=src.device)
        idx = index.view([-1 if d == dim else 1 for d in range(src.dim())]).expand_as(src)
        out = out.scatter_reduce(dim, idx, src, reduce="amax", include_self=True)
        return torch.where(torch.isinf(out) & (out < 0), torch.zeros_like(out), out)
    raise ValueError(reduce)


def scatter_mean(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None) -> Tensor:
    """torch_scatter.scatter_mean; call site `SDE_model_2D_to_3D.py:386`."""
    return scatter(src, index, dim, dim_size, "mean")


def scatter_add(src: Tensor, index: Tensor, dim: int = 0, dim_size: Optional[int] = None) -> Tensor:
    return scatter(src, index, dim, dim_size, "sum")


# ----------------------------------------------------------------------------
# torch_geometric.utils
# ----------------------------------------------------------------------------
def segment_softmax(src: Tensor, index: Tensor, num_nodes: int) -> Tensor:
    """torch_geometric.utils.softmax (pyg 2.0.2): per-target max-shifted softmax with
    `+1e-16` in the denominator.  Used by TransformerConv.message
    (reference call site `equivariant_scorenetwork.py:18-24,35`)."""
    idx = index.view(-1, *([1] * (src.dim() - 1))).expand_as(src)
    smax = torch.full((num_nodes,) + tuple(src.shape[1:]), float("-inf"), dtype=src.dtype)
    smax = smax.scatter_reduce(0, idx, src, reduce="amax", include_self=True)
    out = (src - smax.index_select(0, index)).exp()
    ssum = torch.zeros((num_nodes,) + tuple(src.shape[1:]), dtype=src.dtype).index_add_(0, index, out)
    return out / (ssum.index_select(0, index) + 1e-16)


def remove_self_loops(edge_index: Tensor, edge_attr: Optional[Tensor] = None):
    """torch_geometric.utils.remove_self_loops; call sites `dataset_3D.py:20,30`."""
    mask = edge_index[0] != edge_index[1]
    edge_index = edge_index[:, mask]
    return edge_index, (None if edge_attr is None else edge_attr[mask])


def to_dense_batch(x: Tensor, batch: Tensor, fill_value: float = 0.0,
                   max_num_nodes: Optional[int] = None, batch_size: Optional[int] = None):
    """torch_geometric.utils.to_dense_batch; call sites
    `SDE_model_3D_to_2D_node_adj_dense.py:130-131`.  `batch` must be ascending."""
    if batch_size is None:
        batch_size = int(batch.max().item()) + 1
    num_nodes = torch.zeros(batch_size, dtype=torch.long).index_add_(0, batch, torch.ones_like(batch))
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max().item())
    local = torch.arange(batch.numel()) - cum[batch]
    idx = batch * max_num_nodes + local
    size = [batch_size * max_num_nodes] + list(x.shape[1:])
    out = x.new_full(size, fill_value)
    out[idx] = x
    out = out.view([batch_size, max_num_nodes] + list(x.shape[1:]))
    mask = torch.zeros(batch_size * max_num_nodes, dtype=torch.bool)
    mask[idx] = True
    return out, mask.view(batch_size, max_num_nodes)


def to_dense_adj(edge_index: Tensor, batch: Tensor, edge_attr: Optional[Tensor] = None,
                 max_num_nodes: Optional[int] = None) -> Tensor:
    """torch_geometric.utils.to_dense_adj (scatter-ADD of edge_attr into [B,Nmax,Nmax]);
    call site `SDE_model_3D_to_2D_node_adj_dense.py:129`."""
    batch_size = int(batch.max().item()) + 1
    num_nodes = torch.zeros(batch_size, dtype=torch.long).index_add_(0, batch, torch.ones_like(batch))
    cum = torch.cat([num_nodes.new_zeros(1), num_nodes.cumsum(0)])
    if max_num_nodes is None:
        max_num_nodes = int(num_nodes.max().item())
    g = batch[edge_index[0]]
    i = edge_index[0] - cum[g]
    j = edge_index[1] - cum[g]
    if edge_attr is None:
        edge_attr = torch.ones(edge_index.size(1))
    size = [batch_size * max_num_nodes * max_num_nodes] + list(edge_attr.shape[1:])
    flat = g * max_num_nodes * max_num_nodes + i * max_num_nodes + j
    adj = torch.zeros(size, dtype=edge_attr.dtype).index_add_(0, flat, edge_attr)
    return adj.view([batch_size, max_num_nodes, max_num_nodes] + list(edge_attr.shape[1:]))


# ----------------------------------------------------------------------------
# torch_sparse
# ----------------------------------------------------------------------------
def coalesce(index: Tensor, value: Optional[Tensor], m: int, n: int):
    """torch_sparse.coalesce: sort by row*n+col, drop duplicates (values summed);
    call sites `dataset_3D.py:24,34`."""
    key = index[0] * n + index[1]
    uniq, inv = torch.unique(key, sorted=True, return_inverse=True)
    out = torch.stack([uniq // n, uniq % n], dim=0)
    if value is None:
        return out, None
    val = torch.zeros((uniq.numel(),) + tuple(value.shape[1:]), dtype=value.dtype).index_add_(0, inv, value)
    return out, val


def spspmm(indexA: Tensor, valueA: Tensor, indexB: Tensor, valueB: Tensor, m: int, k: int, n: int):
    """torch_sparse.spspmm: C = A @ B for COO operands, result coalesced (row-major
    sorted); call sites `dataset_3D.py:18,28`.  Structural zeros are kept out only
    where no (i,k),(k,j) pair exists."""
    A = torch.zeros(m, k, dtype=valueA.dtype).index_put_((indexA[0], indexA[1]), valueA, accumulate=True)
    B = torch.zeros(k, n, dtype=valueB.dtype).index_put_((indexB[0], indexB[1]), valueB, accumulate=True)
    SA = torch.zeros(m, k).index_put_((indexA[0], indexA[1]), torch.ones(indexA.size(1)), accumulate=True)
    SB = torch.zeros(k, n).index_put_((indexB[0], indexB[1]), torch.ones(indexB.size(1)), accumulate=True)
    C = A @ B
    S = (SA @ SB) > 0
    idx = S.nonzero().t().contiguous()
    return idx, C[idx[0], idx[1]]


# ----------------------------------------------------------------------------
# torch_cluster
# ----------------------------------------------------------------------------
def sq_dist_f32(a: Tensor, b: Tensor) -> Tensor:
    """fp32 squared distance with the fixed, FMA-free evaluation order
    ((dx*dx + dy*dy) + dz*dz); the CUDA kernel uses __fmul_rn/__fadd_rn to match
    it bit for bit."""
    d = a - b
    return (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]


def radius_graph(x: Tensor, r: float, batch: Optional[Tensor] = None, loop: bool = False,
                 max_num_neighbors: int = 32, flow: str = "source_to_target") -> Tensor:
    """torch_cluster.radius_graph, CUDA-kernel semantics (the reference trains on GPU):
    for every target i scan the atoms of its own graph in ascending index, keep those with
    d^2 < r^2 (strict) until `max_num_neighbors + 1` are found (self included in the scan
    when `loop=False`), then drop the self pair.  Output `[2,E]` int64, row 0 = source j,
    row 1 = target i, grouped by ascending target with ascending sources.
    Call site `Geom3D/models/schnet.py:91` (default `max_num_neighbors=32`)."""
    assert flow == "source_to_target"
    n = x.size(0)
    if batch is None:
        batch = torch.zeros(n, dtype=torch.long)
    cap = max_num_neighbors if loop else max_num_neighbors + 1
    r2 = torch.tensor(float(r) * float(r), dtype=torch.float32)  # torch_cluster: (scalar_t)(r * r)
    counts = torch.zeros(int(batch.max().item()) + 1 if n else 0, dtype=torch.long)
    counts.index_add_(0, batch, torch.ones_like(batch))
    ptr = torch.cat([counts.new_zeros(1), counts.cumsum(0)]).tolist()
    rows, cols = [], []
    for g in range(len(ptr) - 1):
        s, e = ptr[g], ptr[g + 1]
        if e == s:
            continue
        p = x[s:e].float()
        d2 = sq_dist_f32(p[:, None, :], p[None, :, :])  # [target, source]
        within = d2 < r2
        rank = within.long().cumsum(1)  # 1-based rank among found, ascending source index
        keep = within & (rank <= cap)
        if not loop:
            keep = keep & ~torch.eye(e - s, dtype=torch.bool)
        tgt, src = keep.nonzero(as_tuple=True)  # row-major: ascending target, then source
        rows.append(src + s)
        cols.append(tgt + s)
    if not rows:
        return torch.zeros(2, 0, dtype=torch.long)
    return torch.stack([torch.cat(rows), torch.cat(cols)], dim=0)


# ----------------------------------------------------------------------------
# torch_geometric.nn message passing pieces
# ----------------------------------------------------------------------------
def propagate(edge_index: Tensor, message: Tensor, num_nodes: int, aggr: str) -> Tensor:
    """MessagePassing.propagate aggregation step (flow source_to_target): reduce the
    per-edge `message` at `edge_index[1]` with `dim_size=num_nodes`.
    Call sites `schnet.py:190` (add), `equivariant_scorenetwork.py:71` (mean),
    `molecule_gnn_model.py:24` (add)."""
    return scatter(message, edge_index[1], 0, num_nodes, "mean" if aggr == "mean" else "sum")


def transformer_conv(x: Tensor, edge_index: Tensor, edge_attr: Tensor, w: dict, heads: int,
                     out_channels: int, dropout_mask: Optional[Tensor] = None,
                     dropout_p: float = 0.0) -> Tensor:
    """torch_geometric.nn.TransformerConv (pyg 2.0.2; concat=True, root_weight=True,
    beta=False, edge_dim set, lin_edge without bias).  Reference instantiation
    `equivariant_scorenetwork.py:18-24`, call `:35`.

    `w` holds lin_query/lin_key/lin_value/lin_skip `.weight/.bias` and `lin_edge.weight`.
    `dropout_mask` ([E,heads], entries 0 or 1) reproduces F.dropout(alpha, p) in train mode
    as `alpha * mask / (1 - p)`."""
    H, C = heads, out_channels
    n = x.size(0)
    src, dst = edge_index[0], edge_index[1]
    q = torch.nn.functional.linear(x, w["lin_query.weight"], w["lin_query.bias"]).view(-1, H, C)
    k = torch.nn.functional.linear(x, w["lin_key.weight"], w["lin_key.bias"]).view(-1, H, C)
    v = torch.nn.functional.linear(x, w["lin_value.weight"], w["lin_value.bias"]).view(-1, H, C)
    e = torch.nn.functional.linear(edge_attr, w["lin_edge.weight"]).view(-1, H, C)
    key_j = k.index_select(0, src) + e
    alpha = (q.index_select(0, dst) * key_j).sum(dim=-1) / math.sqrt(C)
    alpha = segment_softmax(alpha, dst, n)
    if dropout_mask is not None:
        alpha = alpha * dropout_mask / (1.0 - dropout_p)
    msg = (v.index_select(0, src) + e) * alpha.view(-1, H, 1)
    out = torch.zeros(n, H, C, dtype=x.dtype).index_add_(0, dst, msg).view(n, H * C)
    return out + torch.nn.functional.linear(x, w["lin_skip.weight"], w["lin_skip.bias"])


def extend_graph_index(edge_index: Tensor, num_nodes: int) -> Tensor:
    """`extend_graph` of `Geom3D/datasets/dataset_3D.py:12-35` on one molecule:
    E2 = coalesce(E u (E.E minus diag)); E4 = coalesce(E2 u (E2.E2 minus diag)).
    Returns the `extended_edge_index` int64 `[2,E_x]`, sorted by row*N+col, unique."""
    ei = edge_index
    for _ in range(2):
        ones = torch.ones(ei.size(1))
        idx, val = spspmm(ei, ones, ei, ones, num_nodes, num_nodes, num_nodes)
        idx, _ = remove_self_loops(idx, val)
        ei = torch.cat([ei, idx], dim=1)
        ei, _ = coalesce(ei, None, num_nodes, num_nodes)
    return ei
