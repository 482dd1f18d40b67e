"""Batch container and the synthetic PCQM4Mv2-shaped molecule generator.

`Data` / `Batch` carry exactly the fields the reference's hot path touches
(`Geom3D/datasets/dataset_utils.py:115-169` defines the layout:
`x i64[n,9]` (col 0 = atomic number - 1), `edge_index i64[2,E_b]` stored as adjacent
(i,j),(j,i) pairs, `edge_attr i64[E_b,3]`, `positions f32[n,3]`), collated the way PyG's
`Batch.from_data_list` does: every key containing "index" is shifted by the node offset and
concatenated along dim 1, everything else along dim 0; `batch` is ascending.

The generator follows SURVEY.md section 8(d).  It is numpy/PCG64-seeded so that the build
container and the GPU box produce identical molecules.
"""
from __future__ import annotations

from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np
import torch

# OGB 1.2.1 vocabulary sizes (ogb/utils/features.py): atom features and bond features.
ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]
BOND_FEATURE_DIMS = [5, 6, 2]


class Data:
    """Minimal stand-in for `torch_geometric.data.Data` (attribute bag of tensors)."""

    def __init__(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)

    @classmethod
    def from_dict(cls, mapping: Dict[str, object]) -> "Data":
        return cls(**dict(mapping))

    @property
    def keys(self) -> List[str]:
        return [k for k, v in self.__dict__.items() if not k.startswith("_") and v is not None]

    def __getitem__(self, key: str):
        return getattr(self, key)

    def __setitem__(self, key: str, value) -> None:
        setattr(self, key, value)

    def __contains__(self, key: str) -> bool:
        return key in self.keys

    @property
    def num_nodes(self) -> int:
        for k in ("x", "positions"):
            v = self.__dict__.get(k)
            if v is not None:
                return int(v.size(0))
        raise AttributeError("cannot infer num_nodes")

    def _apply(self, fn):
        out = self.__class__.__new__(self.__class__)
        for k, v in self.__dict__.items():
            out.__dict__[k] = fn(v) if isinstance(v, torch.Tensor) else v
        return out

    def to(self, device, non_blocking: bool = False):
        return self._apply(lambda t: t.to(device, non_blocking=non_blocking))

    def clone(self):
        return self._apply(lambda t: t.clone())

    def detach(self):
        return self._apply(lambda t: t.detach())

    def __repr__(self) -> str:
        body = ", ".join(f"{k}={list(getattr(self, k).shape) if isinstance(getattr(self, k), torch.Tensor) else getattr(self, k)}" for k in self.keys)
        return f"{self.__class__.__name__}({body})"


class Batch(Data):
    """Collated molecules; `batch` ascending, `ptr` the node offsets, `num_graphs` an int."""

    @classmethod
    def from_data_list(cls, data_list: Sequence[Data]) -> "Batch":
        keys = data_list[0].keys
        cols: Dict[str, list] = {k: [] for k in keys}
        batch_vec, ptr, off = [], [0], 0
        for g, d in enumerate(data_list):
            n = d.num_nodes
            for k in keys:
                v = d[k]
                if isinstance(v, torch.Tensor) and "index" in k:
                    v = v + off
                cols[k].append(v)
            batch_vec.append(torch.full((n,), g, dtype=torch.long))
            off += n
            ptr.append(off)
        out = cls()
        for k in keys:
            v0 = cols[k][0]
            if isinstance(v0, torch.Tensor):
                out[k] = torch.cat(cols[k], dim=1 if "index" in k else 0)
            else:
                out[k] = cols[k]
        out.batch = torch.cat(batch_vec)
        out.ptr = torch.tensor(ptr, dtype=torch.long)
        out.num_graphs = len(data_list)
        return out

    def to_data_list(self) -> List[Data]:
        ptr = self.ptr.tolist()
        out = []
        for g in range(self.num_graphs):
            s, e = ptr[g], ptr[g + 1]
            d = Data()
            for k in self.keys:
                v = self[k]
                if k in ("batch", "ptr", "num_graphs") or not isinstance(v, torch.Tensor):
                    continue
                if "index" in k:
                    m = (v[0] >= s) & (v[0] < e)
                    d[k] = v[:, m] - s
                elif k == "edge_attr":  # per-edge tensor follows edge_index
                    ei = self.edge_index
                    m = (ei[0] >= s) & (ei[0] < e)
                    d[k] = v[m]
                else:
                    d[k] = v[s:e]
            out.append(d)
        return out


# ----------------------------------------------------------------------------
# synthetic molecules
# ----------------------------------------------------------------------------
_ATOM_TYPES = np.array([5, 6, 7, 8, 15, 16])  # atomic number - 1: C N O F S Cl
_ATOM_PROBS = np.array([0.72, 0.12, 0.13, 0.01, 0.01, 0.01])
_BOND_TYPES = np.array([0, 1, 2, 3])
_BOND_PROBS = np.array([0.70, 0.10, 0.02, 0.18])


def _num_atoms(rng: np.random.Generator, kind: str) -> int:
    if kind == "pcqm":
        return int(np.clip(np.rint(rng.normal(14.1, 3.5)), 3, 20))
    if kind == "drug":
        return int(rng.integers(30, 101))
    if kind == "padded64":
        return int(rng.integers(20, 65))
    raise ValueError(kind)


def synth_molecule(rng: np.random.Generator, kind: str = "pcqm", n: Optional[int] = None) -> Data:
    """One synthetic molecule: random tree with max degree 4 plus a few
    ring closures at tree distance 4-5, BFS-embedded at 1.5 A bond length, centred."""
    if n is None:
        n = _num_atoms(rng, kind)
    deg = np.zeros(n, dtype=np.int64)
    parent = np.full(n, -1, dtype=np.int64)
    adj: List[List[int]] = [[] for _ in range(n)]
    bonds: List[tuple] = []
    for v in range(1, n):
        free = np.flatnonzero(deg[:v] < 4)
        p = int(free[rng.integers(len(free))])
        parent[v] = p
        deg[p] += 1
        deg[v] += 1
        adj[p].append(v)
        adj[v].append(p)
        bonds.append((p, v))
    # ring closures: ~Binomial(n, 0.035) extra bonds between nodes at tree distance 4..5
    n_ring = int(rng.binomial(n, 0.035))
    for _ in range(n_ring):
        for _try in range(20):
            a = int(rng.integers(n))
            if deg[a] >= 4:
                continue
            dist = np.full(n, -1)
            dist[a] = 0
            frontier = [a]
            while frontier:
                nxt = []
                for u in frontier:
                    for w in adj[u]:
                        if dist[w] < 0:
                            dist[w] = dist[u] + 1
                            nxt.append(w)
                frontier = nxt
            cand = np.flatnonzero(((dist == 4) | (dist == 5)) & (deg < 4))
            if len(cand) == 0:
                continue
            b = int(cand[rng.integers(len(cand))])
            deg[a] += 1
            deg[b] += 1
            adj[a].append(b)
            adj[b].append(a)
            bonds.append((a, b))
            break
    # features
    x = np.zeros((n, 9), dtype=np.int64)
    x[:, 0] = rng.choice(_ATOM_TYPES, size=n, p=_ATOM_PROBS)
    for c in range(1, 9):
        x[:, c] = rng.integers(0, ATOM_FEATURE_DIMS[c], size=n)
    nb = len(bonds)
    ei = np.zeros((2, 2 * nb), dtype=np.int64)
    ea = np.zeros((2 * nb, 3), dtype=np.int64)
    for k, (a, b) in enumerate(bonds):
        f = (int(rng.choice(_BOND_TYPES, p=_BOND_PROBS)), int(rng.integers(0, 6)), int(rng.integers(0, 2)))
        ei[:, 2 * k] = (a, b)
        ei[:, 2 * k + 1] = (b, a)
        ea[2 * k] = f
        ea[2 * k + 1] = f
    # geometry: BFS embedding along the tree
    pos = np.zeros((n, 3), dtype=np.float64)
    for v in range(1, n):
        p = parent[v]
        for _try in range(50):
            d = rng.normal(size=3)
            d /= np.linalg.norm(d) + 1e-12
            cand = pos[p] + 1.5 * d
            if v == 1 or np.min(np.linalg.norm(pos[:v] - cand, axis=1)) >= 1.0:
                break
        pos[v] = cand
    pos -= pos.mean(axis=0, keepdims=True)  # remove_center, dataset_3D.py:120-122
    return Data(
        x=torch.from_numpy(x),
        edge_index=torch.from_numpy(ei),
        edge_attr=torch.from_numpy(ea),
        positions=torch.from_numpy(pos.astype(np.float32)),
    )


def synth_molecules(num: int, seed: int = 0, kind: str = "pcqm") -> List[Data]:
    rng = np.random.default_rng(seed)
    return [synth_molecule(rng, kind) for _ in range(num)]


def synth_batch(num: int, seed: int = 0, kind: str = "pcqm") -> Batch:
    return Batch.from_data_list(synth_molecules(num, seed, kind))


def repeat_data(data: Data, num_repeat: int) -> Batch:
    """`repeat_data` of `examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:29-39`:
    a batch of `num_repeat` copies of one molecule (plus a `pos_gen` placeholder)."""
    copies = []
    for _ in range(num_repeat):
        d = Data.from_dict({k: data[k] for k in data.keys})
        d.pos_gen = torch.ones_like(data.positions)
        copies.append(d)
    return Batch.from_data_list(copies)
