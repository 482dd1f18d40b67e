"""Chunk / tile plan of a batch for the 2D->3D score and PC kernels (host bookkeeping only).

A chunk is a run of whole molecules that one CTA keeps in shared memory (<= CHUNK_MAX_NODES atoms,
<= MAX_CHUNK_TILES tiles); a tile is a run of whole target nodes with <= TILE_EDGES incoming edges.
Built once per batch from the CSR row pointer (static graph structure); the arrays are described in
`include/molsde_b200.h` (`molsde_plan`).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _abi
from .graph import CSR


@dataclass
class TilePlan:
    csr: CSR
    chunk_tile_ptr: torch.Tensor  # int32 [C+1]
    tile_tgt_ptr: torch.Tensor    # int32 [T+1]
    chunk_order: torch.Tensor     # int32 [C], chunks by descending tile count (work-queue order of the PC kernel)
    num_chunks: int
    num_tiles: int
    max_chunk_tiles: int
    N: int
    E: int

    def as_struct(self) -> _abi.Plan:
        return _abi.Plan(self.num_chunks, self.num_tiles, self.N, self.E, self.chunk_tile_ptr.data_ptr(),
                         self.tile_tgt_ptr.data_ptr(), self.csr.rowptr.data_ptr(),
                         self.csr.col.data_ptr() if self.E else None, self.chunk_order.data_ptr())


def _tiles_of_range(rowptr: np.ndarray, a: int, b: int, te: int) -> List[int]:
    """Greedy target-aligned tiling of nodes [a,b): returns the first node of every tile."""
    starts = []
    i = a
    while i < b:
        starts.append(i)
        # last j with rowptr[j] - rowptr[i] <= te
        j = int(np.searchsorted(rowptr, rowptr[i] + te, side="right")) - 1
        if j <= i:
            raise _abi.MolsdeError(f"node {i} has more than {te} incoming edges (unsupported)")
        i = min(j, b, i + te)  # whole targets only; also bounds the targets per tile
    return starts


def build_plan(csr: CSR, node_ptr: torch.Tensor, groups: Optional[Sequence[int]] = None) -> TilePlan:
    """`node_ptr`: molecule offsets [B+1] (any device).  `groups`: optional molecule offsets
    [G+1] of fixed chunks (sampling groups: one chunk per group, required by the PC kernel);
    otherwise molecules are packed greedily into chunks."""
    rowptr = np.ascontiguousarray(csr.rowptr.detach().cpu().numpy().astype(np.int64))
    nptr = np.ascontiguousarray(node_ptr.detach().cpu().numpy().astype(np.int64))
    N, E = int(nptr[-1]), int(rowptr[-1])
    B = len(nptr) - 1
    te, maxn, maxt = _abi.TILE_EDGES, _abi.CHUNK_MAX_NODES, _abi.MAX_CHUNK_TILES
    grp = None if groups is None else np.ascontiguousarray(np.asarray(groups, dtype=np.int64))
    ctp = np.empty((B if grp is None else len(grp) - 1) + 2, dtype=np.int32)
    ttp = np.empty(N + 2, dtype=np.int32)
    counts = np.zeros(4, dtype=np.int64)
    st = _abi.lib().molsde_build_plan_host(rowptr.ctypes.data, nptr.ctypes.data, B, None if grp is None else grp.ctypes.data,
                                           0 if grp is None else len(grp) - 1, te, maxn, maxt, ctp.ctypes.data, ttp.ctypes.data,
                                           counts.ctypes.data)
    if st != 0:
        code, where = int(counts[2]), int(counts[3])
        if code == -2:
            raise _abi.MolsdeError(f"node {where} has more than {te} incoming edges (unsupported)")
        if code == -3:
            raise _abi.MolsdeError(f"sampling group {where} has more than {maxn} atoms; the fused PC kernel holds at most {maxn}")
        if code == -4:
            raise _abi.MolsdeError(f"chunk {where} needs {int(counts[1])} tiles (> {maxt})")
        raise _abi.MolsdeError(f"build_plan_host failed with status {st}")
    nchunks, ntiles, max_tiles = int(counts[0]), int(counts[1]), int(counts[2])
    chunk_tile_ptr = ctp[:nchunks + 1]
    tile_starts = ttp[:ntiles + 1]
    dev = csr.rowptr.device
    tiles_per_chunk = np.diff(chunk_tile_ptr.astype(np.int64))
    order = np.argsort(-tiles_per_chunk, kind="stable").astype(np.int32)
    return TilePlan(csr, torch.from_numpy(chunk_tile_ptr.copy()).to(dev), torch.from_numpy(tile_starts.copy()).to(dev),
                    torch.from_numpy(order).to(dev), nchunks, ntiles, max_tiles, N, E)


def build_plan_py(csr: CSR, node_ptr: torch.Tensor, groups: Optional[Sequence[int]] = None) -> TilePlan:
    """Pure-Python statement of the same plan (the specification `molsde_build_plan_host` is tested against)."""
    rowptr = csr.rowptr.detach().cpu().numpy().astype(np.int64)
    nptr = node_ptr.detach().cpu().numpy().astype(np.int64)
    N, E = int(nptr[-1]), int(rowptr[-1])
    te, maxn, maxt = _abi.TILE_EDGES, _abi.CHUNK_MAX_NODES, _abi.MAX_CHUNK_TILES
    if groups is not None:
        chunk_bounds = [int(nptr[m]) for m in np.asarray(groups, dtype=np.int64)]
    else:
        chunk_bounds = [0]
        B = len(nptr) - 1
        edge_budget = int(maxt * te * 0.6)  # tiles are ~60-95% full; verified exactly below
        m = 0
        while m < B:
            end = m
            while end < B and nptr[end + 1] - nptr[m] <= maxn and \
                    rowptr[nptr[end + 1]] - rowptr[nptr[m]] <= edge_budget:
                end += 1
            if end == m:
                end = m + 1  # a single big molecule: let the exact checks below decide
            chunk_bounds.append(int(nptr[end]))
            m = end
    chunk_tile_ptr = [0]
    tile_starts: List[int] = []
    max_tiles = 0
    for c in range(len(chunk_bounds) - 1):
        a, b = chunk_bounds[c], chunk_bounds[c + 1]
        if b - a > maxn:
            raise _abi.MolsdeError(f"sampling group {c} has {b - a} atoms; the fused PC kernel holds at most {maxn}")
        ts = _tiles_of_range(rowptr, a, b, te)
        if len(ts) > maxt:
            raise _abi.MolsdeError(f"chunk {c} needs {len(ts)} tiles (> {maxt})")
        tile_starts.extend(ts)
        chunk_tile_ptr.append(len(tile_starts))
        max_tiles = max(max_tiles, len(ts))
    tile_starts.append(N)
    dev = csr.rowptr.device
    tiles_per_chunk = np.diff(np.asarray(chunk_tile_ptr, dtype=np.int64))
    order = np.argsort(-tiles_per_chunk, kind="stable").astype(np.int32)
    return TilePlan(csr, torch.tensor(chunk_tile_ptr, dtype=torch.int32, device=dev),
                    torch.tensor(tile_starts, dtype=torch.int32, device=dev),
                    torch.from_numpy(order).to(dev),
                    len(chunk_bounds) - 1, len(tile_starts) - 1, max_tiles, N, E)
