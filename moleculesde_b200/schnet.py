"""SchNet with the reference's constructor, forward signature and state_dict keys
(`Geom3D/models/schnet.py:16-216`), forward pass through the sm_100a kernels:

  radius graph (K1, `molsde_radius_graph_*`) -> embedding gather -> 6 x [conv.lin1 (linear),
  fused CFConv edge kernel (`molsde_schnet_cfconv`: GaussianSmearing + filter MLP + cutoff + message +
  deterministic per-target sum), conv.lin2 + ShiftedSoftplus (linear epilogue), lin + residual] ->
  lin1 / ssp / lin2 -> per-graph readout.

Inference / feature extraction only this round (no autograd through the kernels).
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
from torch import nn

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr
from .graph import radius_graph, segment_ptr
from .ops import gather_rows, linear, segment_reduce
from .plan import build_plan

_NG_PAD, _LD = 56, 136


class ShiftedSoftplus(nn.Module):
    def __init__(self):
        super().__init__()
        self.shift = torch.log(torch.tensor(2.0)).item()


class GaussianSmearing(nn.Module):
    """`schnet.py:198-207`: buffer `offset`, python-float `coeff` from the fp32 spacing."""

    def __init__(self, start=0.0, stop=5.0, num_gaussians=50):
        super().__init__()
        offset = torch.linspace(start, stop, num_gaussians)
        self.coeff = -0.5 / (offset[1] - offset[0]).item() ** 2
        self.register_buffer("offset", offset)


class CFConv(nn.Module):
    """Parameter container of `schnet.py:170-183` (`lin1` without bias, `lin2`, shared filter MLP `nn`)."""

    def __init__(self, in_channels, out_channels, num_filters, nn_module, cutoff):
        super().__init__()
        self.lin1 = nn.Linear(in_channels, num_filters, bias=False)
        self.lin2 = nn.Linear(num_filters, out_channels)
        self.nn = nn_module
        self.cutoff = cutoff
        nn.init.xavier_uniform_(self.lin1.weight)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)


class InteractionBlock(nn.Module):
    """`schnet.py:138-161`: the filter MLP is registered twice (`mlp` and `conv.nn`), as in the reference."""

    def __init__(self, hidden_channels, num_gaussians, num_filters, cutoff):
        super().__init__()
        self.mlp = nn.Sequential(nn.Linear(num_gaussians, num_filters), ShiftedSoftplus(), nn.Linear(num_filters, num_filters))
        self.conv = CFConv(hidden_channels, hidden_channels, num_filters, self.mlp, cutoff)
        self.act = ShiftedSoftplus()
        self.lin = nn.Linear(hidden_channels, hidden_channels)
        nn.init.xavier_uniform_(self.mlp[0].weight)
        self.mlp[0].bias.data.fill_(0)
        nn.init.xavier_uniform_(self.mlp[2].weight)
        nn.init.xavier_uniform_(self.lin.weight)
        self.lin.bias.data.fill_(0)


# Standard atomic masses (IUPAC 2016, index = atomic number, entry 0 = the dummy element 'X'), the values `ase.data.atomic_masses`
# holds: the reference registers them as the float64 buffer `atomic_mass` (schnet.py:47-48) and reads them for the dipole
# readout's centre of mass (:105-107).  Embedded so that a checkpoint written here carries the real table into the reference class.
_ATOMIC_MASSES = (
    1.0, 1.008, 4.002602, 6.94, 9.0121831, 10.81, 12.011, 14.007, 15.999, 18.998403163, 20.1797,
    22.98976928, 24.305, 26.9815385, 28.085, 30.973761998, 32.06, 35.45, 39.948, 39.0983, 40.078,
    44.955908, 47.867, 50.9415, 51.9961, 54.938044, 55.845, 58.933194, 58.6934, 63.546, 65.38,
    69.723, 72.630, 74.921595, 78.971, 79.904, 83.798, 85.4678, 87.62, 88.90584, 91.224,
    92.90637, 95.95, 97.90721, 101.07, 102.90550, 106.42, 107.8682, 112.414, 114.818, 118.710,
    121.760, 127.60, 126.90447, 131.293, 132.90545196, 137.327, 138.90547, 140.116, 140.90766, 144.242,
    144.91276, 150.36, 151.964, 157.25, 158.92535, 162.500, 164.93033, 167.259, 168.93422, 173.054,
    174.9668, 178.49, 180.94788, 183.84, 186.207, 190.23, 192.217, 195.084, 196.966569, 200.592,
    204.38, 207.2, 208.98040, 208.98243, 209.98715, 222.01758, 223.01974, 226.02541, 227.02775, 232.0377,
    231.03588, 238.02891, 237.04817, 244.06421, 243.06138, 247.07035, 247.07031, 251.07959, 252.0830, 257.09511,
    258.09843, 259.1010, 262.110, 267.122, 268.126, 271.134, 270.133, 269.1338, 278.156, 281.165,
    281.166, 285.177, 286.182, 289.190, 289.194, 293.204, 293.208, 294.214)
assert len(_ATOMIC_MASSES) == 119


class SchNet(nn.Module):
    def __init__(self, hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=50, cutoff=10.0,
                 node_class=None, readout="mean", dipole=False, mean=None, std=None, atomref=None):
        super().__init__()
        assert readout in ["add", "sum", "mean"]
        if dipole or atomref is not None or mean is not None or std is not None:
            raise NotImplementedError("dipole / atomref / mean-std heads are not on the pretraining path")
        if num_filters != 128 or num_gaussians > _NG_PAD:
            raise NotImplementedError("kernels are compiled for num_filters=128, num_gaussians<=56 (config.py:66-70)")
        self.hidden_channels, self.num_filters = hidden_channels, num_filters
        self.num_interactions, self.num_gaussians, self.cutoff = num_interactions, num_gaussians, cutoff
        self.readout, self.dipole, self.mean, self.std, self.scale = readout, dipole, mean, std, None
        # `ase.data.atomic_masses` in the reference (schnet.py:47-48); only read when dipole=True (not built here); kept with the real values for checkpoint parity.
        self.register_buffer("atomic_mass", torch.tensor(_ATOMIC_MASSES, dtype=torch.float64))
        self.embedding = nn.Embedding(node_class, hidden_channels)
        self.distance_expansion = GaussianSmearing(0.0, cutoff, num_gaussians)
        self.interactions = nn.ModuleList(InteractionBlock(hidden_channels, num_gaussians, num_filters, cutoff)
                                          for _ in range(num_interactions))
        self.lin1 = nn.Linear(hidden_channels, hidden_channels)
        self.act = ShiftedSoftplus()
        self.lin2 = nn.Linear(hidden_channels, hidden_channels)
        self.register_buffer("initial_atomref", atomref)
        self.atomref = None
        nn.init.xavier_uniform_(self.lin1.weight)
        self.lin1.bias.data.fill_(0)
        nn.init.xavier_uniform_(self.lin2.weight)
        self.lin2.bias.data.fill_(0)
        self._packed = None

    # filter-MLP weights in the kernel layout (k-major, ld 136, zero padded), rebuilt when parameters change
    def _filters(self):
        ver = (_abi.param_epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed[0] == ver:
            return self._packed[1]
        packed = []
        dev = self.lin1.weight.device
        mu = torch.zeros(_NG_PAD, dtype=torch.float32, device=dev)
        mu[:self.num_gaussians] = self.distance_expansion.offset.float()
        for blk in self.interactions:
            w1 = torch.zeros(_NG_PAD, _LD, dtype=torch.float32, device=dev)
            w1[:self.num_gaussians, :128] = blk.mlp[0].weight.detach().float().t()
            w2 = torch.zeros(128, _LD, dtype=torch.float32, device=dev)
            w2[:, :128] = blk.mlp[2].weight.detach().float().t()
            packed.append((w1.contiguous(), blk.mlp[0].bias.detach().float().contiguous(), w2.contiguous(),
                           blk.mlp[2].bias.detach().float().contiguous()))
        self._packed = (ver, (mu, packed))
        return self._packed[1]

    def forward(self, z, pos, batch=None, return_latent=False):
        """Reference signature (`schnet.py:85`).  Inference: the fused edge-kernel path below.  With autograd enabled and
        trainable parameters: one autograd node over the layer-granular training kernels (`autograd.py`, `pretrain.tape_schnet`);
        gradients flow to the parameters (positions are inputs of the pretraining step, not differentiated)."""
        assert z.dim() == 1 and z.dtype == torch.long
        require_device(pos)
        batch = torch.zeros_like(z) if batch is None else batch
        num_graphs = int(batch[-1].item()) + 1 if batch.numel() else 0
        from . import autograd as AG
        if (self.training and AG.grad_mode(self)) or (torch.is_grad_enabled() and pos.requires_grad):
            from .pretrain import tape_schnet
            node_ptr = segment_ptr(batch, num_graphs)
            row2seg = batch.to(torch.int32).contiguous()
            mean = self.readout == "mean"

            def build(tp, ins, P):
                pv = ins[0] if ins else None
                h = tape_schnet(tp, self, P, z.contiguous(), pos, batch, num_graphs, {}, pos_var=pv)
                out = tp.segment_readout(h, node_ptr, row2seg, mean)                   # :115, differentiable
                if self.scale is not None:
                    raise NotImplementedError("SchNet.scale with gradients (unused by the reference scripts)")

                def seed(gouts):   # torch's grad_outputs for (out, h); either may be absent
                    out.grad = gouts[0]
                    h.grad = None if gouts[1] is None else gouts[1].clone()   # the readout backward accumulates into it in place
                return [out.data, h.data], seed
            def dual(tp, ins, tans, P):
                """primal + forward-mode tangent along a position displacement (the double backward of the force term)"""
                from .pretrain import tape_schnet_dual
                h, hd = tape_schnet_dual(tp, self, P, z.contiguous(), pos, tans[0], batch, num_graphs, {})
                out = tp.segment_readout(h, node_ptr, row2seg, mean)
                outd = tp.segment_readout(hd, node_ptr, row2seg, mean)
                return [out, h], [outd, hd]
            # positions take part in autograd only when the caller asked for it (`positions.requires_grad_()`, finetune_MD17.py:49):
            # forces -dE/dpos; with `create_graph=True` (a force term inside the training loss, :66-77) the force is itself
            # differentiable in the parameters through the tangent builder above
            out, h = AG.apply(self, build, [pos] if pos.requires_grad else [], dual=dual)
            return (out, h) if return_latent else out
        with torch.no_grad():
            return self._forward_inference(z, pos, batch, num_graphs, return_latent)

    def _forward_inference(self, z, pos, batch, num_graphs, return_latent):
        pos = pos.detach().float().contiguous()
        csr = radius_graph(pos, self.cutoff, batch, num_graphs, want_edge_index=False)   # schnet.py:91
        node_ptr = segment_ptr(batch, num_graphs)
        plan = build_plan(csr, node_ptr)
        st = plan.as_struct()
        mu, filters = self._filters()
        h = gather_rows(self.embedding.weight, z)                                      # :89
        N = h.size(0)
        agg = torch.empty(N, 128, dtype=torch.float32, device=h.device)
        s = stream_ptr(h)
        for blk, (w1, b1, w2, b2) in zip(self.interactions, filters):
            x = linear(h, blk.conv.lin1.weight)                                         # :189
            check(lib().molsde_schnet_cfconv(ctypes.byref(st), ptr(pos), ptr(x), ptr(w1), ptr(b1), ptr(w2), ptr(b2), ptr(mu),
                                             self.num_gaussians, float(self.distance_expansion.coeff), float(self.cutoff),
                                             ptr(agg), s), "schnet_cfconv")                 # :186-190
            t = linear(agg, blk.conv.lin2.weight, blk.conv.lin2.bias, act="ssp")        # :191, :165
            h = linear(t, blk.lin.weight, blk.lin.bias, residual=h)                     # :166, :97
        h = linear(h, self.lin1.weight, self.lin1.bias, act="ssp")                      # :99-100
        h = linear(h, self.lin2.weight, self.lin2.bias)                                 # :101
        out = segment_reduce(h, node_ptr, mean=(self.readout == "mean"))               # :115
        if self.scale is not None:
            out = self.scale * out
        return (out, h) if return_latent else out
