"""torch_geometric.nn.conv.MessagePassing (pyg 2.0.2), Tensor edge_index, flow
source_to_target, node_dim=0: gather `*_j` from edge_index[0], `*_i` from edge_index[1],
call `message`, reduce at edge_index[1], call `update`."""
import inspect

import torch
from oracle.ref_ops import propagate as _aggregate


class MessagePassing(torch.nn.Module):
    def __init__(self, aggr="add", flow="source_to_target", node_dim=0, **kwargs):
        super().__init__()
        assert flow == "source_to_target" and node_dim == 0
        self.aggr = aggr
        self.node_dim = node_dim

    def propagate(self, edge_index, size=None, **kwargs):
        params = list(inspect.signature(self.message).parameters)
        num_nodes = None
        for v in kwargs.values():
            if isinstance(v, (tuple, list)):
                v = v[1] if v[1] is not None else v[0]
            if isinstance(v, torch.Tensor) and v.dim() >= 1 and num_nodes is None:
                num_nodes = v.size(0)
        # prefer a node-level tensor (x / query / key / value) for the node count
        for name in ("x", "query", "key", "value"):
            if name in kwargs:
                v = kwargs[name]
                if isinstance(v, (tuple, list)):
                    v = v[1] if v[1] is not None else v[0]
                num_nodes = v.size(0)
                break
        args = {}
        for p in params:
            if p == "size_i":
                args[p] = num_nodes
            elif p.endswith("_j") or p.endswith("_i"):
                src = kwargs[p[:-2]]
                side = 0 if p.endswith("_j") else 1
                if isinstance(src, (tuple, list)):
                    src = src[side] if src[side] is not None else src[1 - side]
                args[p] = src.index_select(0, edge_index[side])
            elif p == "index":
                args[p] = edge_index[1]
            elif p == "ptr":
                args[p] = None
            else:
                args[p] = kwargs[p]
        out = self.message(**args)
        out = _aggregate(edge_index, out, num_nodes, self.aggr)
        return self.update(out)

    def message(self, x_j):
        return x_j

    def update(self, aggr_out):
        return aggr_out
