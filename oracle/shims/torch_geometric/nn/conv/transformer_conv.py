"""torch_geometric.nn.TransformerConv (pyg 2.0.2), restated with the same parameter
names so the reference state_dict layout (`MHA.lin_key.weight`, ...) is reproduced."""
import math

import torch
import torch.nn.functional as F
from torch.nn import Linear

from torch_geometric.utils import softmax
from .message_passing import MessagePassing


class TransformerConv(MessagePassing):
    def __init__(self, in_channels, out_channels, heads=1, concat=True, beta=False, dropout=0.0,
                 edge_dim=None, bias=True, root_weight=True, **kwargs):
        kwargs.setdefault("aggr", "add")
        super().__init__(node_dim=0, **kwargs)
        assert concat and root_weight and not beta
        self.in_channels, self.out_channels, self.heads = in_channels, out_channels, heads
        self.dropout, self.edge_dim = dropout, edge_dim
        self.lin_key = Linear(in_channels, heads * out_channels)
        self.lin_query = Linear(in_channels, heads * out_channels)
        self.lin_value = Linear(in_channels, heads * out_channels)
        if edge_dim is not None:
            self.lin_edge = Linear(edge_dim, heads * out_channels, bias=False)
        else:
            self.lin_edge = self.register_parameter("lin_edge", None)
        self.lin_skip = Linear(in_channels, heads * out_channels, bias=bias)

    def forward(self, x, edge_index, edge_attr=None):
        H, C = self.heads, self.out_channels
        query = self.lin_query(x).view(-1, H, C)
        key = self.lin_key(x).view(-1, H, C)
        value = self.lin_value(x).view(-1, H, C)
        out = self.propagate(edge_index, query=query, key=key, value=value, edge_attr=edge_attr, size=None)
        out = out.view(-1, H * C)
        return out + self.lin_skip(x)

    def message(self, query_i, key_j, value_j, edge_attr, index, ptr, size_i):
        if self.lin_edge is not None:
            edge_attr = self.lin_edge(edge_attr).view(-1, self.heads, self.out_channels)
            key_j = key_j + edge_attr
        alpha = (query_i * key_j).sum(dim=-1) / math.sqrt(self.out_channels)
        alpha = softmax(alpha, index, ptr, size_i)
        alpha = F.dropout(alpha, p=self.dropout, training=self.training)
        out = value_j
        if edge_attr is not None:
            out = out + edge_attr
        return out * alpha.view(-1, self.heads, 1)
