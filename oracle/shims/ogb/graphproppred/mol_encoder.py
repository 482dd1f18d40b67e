"""ogb.graphproppred.mol_encoder (1.2.1): sum of per-column embeddings, xavier init."""
import torch
from ogb.utils.features import get_atom_feature_dims, get_bond_feature_dims


class AtomEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.atom_embedding_list = torch.nn.ModuleList()
        for dim in get_atom_feature_dims():
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.atom_embedding_list.append(emb)

    def forward(self, x):
        out = 0
        for i in range(x.shape[1]):
            out = out + self.atom_embedding_list[i](x[:, i])
        return out


class BondEncoder(torch.nn.Module):
    def __init__(self, emb_dim):
        super().__init__()
        self.bond_embedding_list = torch.nn.ModuleList()
        for dim in get_bond_feature_dims():
            emb = torch.nn.Embedding(dim, emb_dim)
            torch.nn.init.xavier_uniform_(emb.weight.data)
            self.bond_embedding_list.append(emb)

    def forward(self, edge_attr):
        out = 0
        for i in range(edge_attr.shape[1]):
            out = out + self.bond_embedding_list[i](edge_attr[:, i])
        return out
