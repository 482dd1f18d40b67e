"""GIN 2D encoder with the reference's constructor and state_dict keys (`Geom3D/models/molecule_gnn_model.py:13-32,
132-197`, `GNN(num_layer, emb_dim, JK="last", drop_ratio=0, gnn_type="GIN")`): AtomEncoder / BondEncoder lookups
(ogb 1.2.1 feature dims), GINConv `mlp((1+eps) x + sum relu(x_j + e))`, BatchNorm, ReLU between layers.
The module only OWNS parameters; `forward` runs the CUDA kernels (eval or train statistics, no autograd — the
training step records the same ops on a `Tape`, see `pretrain.tape_gin`).
"""
from __future__ import annotations

import torch
from torch import nn

ATOM_FEATURE_DIMS = [119, 4, 12, 12, 10, 6, 6, 2, 2]   # ogb.utils.features.get_atom_feature_dims()
BOND_FEATURE_DIMS = [5, 6, 2]                           # get_bond_feature_dims()


class _Encoder(nn.Module):
    def __init__(self, dims, emb_dim, list_name):
        super().__init__()
        lst = nn.ModuleList()
        for d in dims:
            e = nn.Embedding(d, emb_dim)
            nn.init.xavier_uniform_(e.weight.data)
            lst.append(e)
        setattr(self, list_name, lst)
        self.dims = list(dims)

    def tables(self):
        return [e.weight for e in next(iter(self.children()))]


class GINConv(nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(emb_dim, 2 * emb_dim), nn.BatchNorm1d(2 * emb_dim), nn.ReLU(),
                                 nn.Linear(2 * emb_dim, emb_dim))
        self.eps = nn.Parameter(torch.Tensor([0]))
        self.bond_encoder = _Encoder(BOND_FEATURE_DIMS, emb_dim, "bond_embedding_list")


class GNN(nn.Module):
    def __init__(self, num_layer, emb_dim, JK="last", drop_ratio=0, gnn_type="GIN"):
        super().__init__()
        if num_layer < 2:
            raise ValueError("Number of GNN layers must be greater than 1.")
        if gnn_type != "GIN" or JK != "last" or drop_ratio != 0:
            raise NotImplementedError("the pretraining demo runs GIN, JK=last, dropout_ratio=0 (README.md:89)")
        self.num_layer, self.emb_dim, self.JK, self.drop_ratio = num_layer, emb_dim, JK, drop_ratio
        self.atom_encoder = _Encoder(ATOM_FEATURE_DIMS, emb_dim, "atom_embedding_list")
        self.gnns = nn.ModuleList(GINConv(emb_dim) for _ in range(num_layer))
        self.batch_norms = nn.ModuleList(nn.BatchNorm1d(emb_dim) for _ in range(num_layer))

    def forward(self, *argv):
        """`forward(x, edge_index, edge_attr)` or `forward(data)` (`molecule_gnn_model.py:160-167`) -> [N, emb_dim].
        With autograd enabled and trainable parameters the call is one autograd node (`autograd.py`): `.backward()` runs our
        backward kernels; otherwise it is a plain inference call."""
        from . import autograd as AG
        from .pretrain import tape_gin
        from .tape import Tape, Var
        batch, num_graphs, cache = None, 1, None
        if len(argv) == 3:
            x, edge_index, edge_attr = argv
        elif len(argv) == 1:
            x, edge_index, edge_attr = argv[0].x, argv[0].edge_index, argv[0].edge_attr
            batch, num_graphs = getattr(argv[0], "batch", None), getattr(argv[0], "num_graphs", 1)
            cache = argv[0].__dict__.setdefault("_molsde_train_cache", {})
        else:
            raise ValueError("unmatched number of arguments.")
        if self.training and AG.grad_mode(self):
            cache = cache if cache is not None else {}

            def build(tp, ins, P):
                h = tape_gin(tp, self, P, x, edge_index, edge_attr, cache, batch, num_graphs)

                def seed(gouts):
                    h.grad = gouts[0]
                return [h.data], seed
            return AG.apply(self, build, [])
        with torch.no_grad():
            tp = Tape(x.device)
            P = {n: Var(p.data, False) for n, p in self.named_parameters()}
            return tape_gin(tp, self, P, x, edge_index, edge_attr, None, batch, num_graphs).data
