"""SDEModel3Dto2D_node_adj_dense with the reference's constructor, `forward` / `get_score_fn` signatures
and state_dict keys (`Geom3D/models/MoleculeSDE/SDE_model_3D_to_2D_node_adj_dense.py:13-179`,
`invariant_scorenetwork_dense.py:40-131`, `layers/edge_network_dense.py`, `layers/node_network_dense.py`),
computing through the dense building-block kernels of `csrc/dense.cu` + `molsde_linear`.

Forward values only this round (score networks, DSM losses, sampler); no autograd through the kernels.
"""
from __future__ import annotations

import math
import os
from typing import Dict, List, Optional, Tuple

import torch
from torch import nn

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr
from .graph import segment_ptr
from .ops import gather_rows, linear
from .sde_2d_to_3d import MultiLayerPerceptron
from .sde_dense import VESDE, VPSDE, subVPSDE

EPSILON = 1e-6
ACT = {"none": 0, "relu": 1, "silu": 2, "tanh": 4, "elu": 5}


# ----------------------------------------------------------------------------------------------
# thin wrappers of the dense kernels
# ----------------------------------------------------------------------------------------------
def _grouped_linear(x2: torch.Tensor, W: torch.Tensor, b: torch.Tensor, G: int, Ki: int, No: int, act: str) -> torch.Tensor:
    """y[:, g*No:(g+1)*No] = act(x[:, g*Ki:(g+1)*Ki] W_g^T + b_g): G small GEMMs in one batched tensor-core launch."""
    rows = x2.size(0)
    y = torch.empty(rows, G * No, dtype=torch.float32, device=x2.device)
    if rows >= 256:
        check(lib().molsde_tc_gemm_batched(G, rows, No, Ki, x2.data_ptr(), x2.stride(0), 1, Ki, ptr(W), Ki, 1, No * Ki, ptr(b), No,
                                           ACT[act], ptr(y), G * No, No, 0, None, 0, None, stream_ptr(x2)), "tc_gemm_batched")
    else:
        check(lib().molsde_grouped_linear(x2.data_ptr(), rows, x2.stride(0), ptr(W), ptr(b), G, Ki, No, ptr(y), G * No, ACT[act],
                                          stream_ptr(x2)), "grouped_linear")
    return y


def _dense_gcn(adjc: torch.Tensor, C: int, xw: torch.Tensor, bias: torch.Tensor, Fo: int, out: torch.Tensor, out_off: int,
               act: str) -> None:
    """adjc [B,C,Nm,Nm] (or [B,Nm,Nm] with C == 1), xw [B*Nm, C*Fo], out 2-D view with unit column stride."""
    B, Nm = adjc.size(0), adjc.size(-1)
    sb = adjc.stride(0)
    sc = adjc.stride(1) if adjc.dim() == 4 else 0
    check(lib().molsde_dense_gcn(adjc.data_ptr(), sb, sc, B, C, Nm, xw.data_ptr(), xw.stride(0), ptr(bias), Fo, out.data_ptr(),
                                 out.stride(0), out_off, ACT[act], stream_ptr(xw)), "dense_gcn")


def _mlp_rows(x2: torch.Tensor, mlp: "MultiLayerPerceptron", act: str) -> torch.Tensor:
    """MultiLayerPerceptron over very many rows.  3 layers with narrow widths run as ONE fused kernel (`molsde_mlp3_rows`);
    anything else layer by layer."""
    layers = list(mlp.layers)
    if len(layers) == 3 and layers[0].in_features <= 32 and layers[0].out_features <= 64 and layers[1].out_features <= 64 \
            and layers[2].out_features <= 8 and x2.stride(1) == 1:
        rows = x2.size(0)
        w = [l.weight.detach().float().contiguous() for l in layers]
        b = [l.bias.detach().float().contiguous() for l in layers]
        y = torch.empty(rows, w[2].size(0), dtype=torch.float32, device=x2.device)
        check(lib().molsde_mlp3_rows(x2.data_ptr(), rows, x2.stride(0), w[0].size(1), ptr(w[0]), ptr(b[0]), w[0].size(0), ptr(w[1]),
                                     ptr(b[1]), w[1].size(0), ptr(w[2]), ptr(b[2]), w[2].size(0), ACT[act], ptr(y), y.stride(0),
                                     stream_ptr(x2)), "mlp3_rows")
        return y
    m = x2
    for i, lyr in enumerate(layers):
        m = linear(m, lyr.weight, lyr.bias, act=act if i < len(layers) - 1 else None)
    return m


def node_flags(adj: torch.Tensor, eps: float = 1e-5) -> torch.Tensor:
    """`SDE_model_3D_to_2D_node_adj_dense.py:523-529` for a [B,Nm,Nm] adjacency."""
    require_device(adj)
    adj = adj.contiguous()
    B, Nm = adj.size(0), adj.size(1)
    flags = torch.empty(B, Nm, dtype=torch.float32, device=adj.device)
    check(lib().molsde_node_flags(ptr(adj), B, Nm, eps, ptr(flags), stream_ptr(adj)), "node_flags")
    return flags


def mask_x(x: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    x = x.contiguous()
    out = torch.empty_like(x)
    check(lib().molsde_mask_rows(ptr(x), ptr(flags.contiguous()), x.numel() // x.size(-1), x.size(-1), ptr(out), stream_ptr(x)),
          "mask_rows")
    return out


def mask_adjs(adj: torch.Tensor, flags: torch.Tensor) -> torch.Tensor:
    """`:543-556` for [B,Nm,Nm]: adj * f_i * f_j (perturb kernel with coef 1, std 0)."""
    adj = adj.contiguous()
    B, Nm = adj.size(0), adj.size(1)
    one = torch.ones(B, dtype=torch.float32, device=adj.device)
    zero = torch.zeros(B, dtype=torch.float32, device=adj.device)
    out = torch.empty_like(adj)
    check(lib().molsde_dense_perturb_adj(ptr(adj), ptr(adj), ptr(flags.contiguous()), ptr(one), ptr(zero), B, Nm, ptr(out),
                                         stream_ptr(adj)), "mask_adjs")
    return out


def gen_noise(x: torch.Tensor, flags: torch.Tensor, sym: bool = True, raw: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`:532-540`; `raw` injects the `randn_like(x)` draw."""
    raw = torch.randn_like(x) if raw is None else raw.to(x.device).float()
    raw = raw.contiguous()
    if not sym:
        return mask_x(raw, flags)
    B, Nm = x.size(0), x.size(1)
    z = torch.empty_like(raw)
    check(lib().molsde_dense_sym_noise(ptr(raw), ptr(flags.contiguous()), B, Nm, ptr(z), stream_ptr(raw)), "sym_noise")
    return z


def graph_reduce(a: torch.Tensor, b: Optional[torch.Tensor], w: Optional[torch.Tensor], mode: int) -> torch.Tensor:
    a = a.contiguous()
    B = a.size(0)
    M = a.numel() // B
    out = torch.empty(B, dtype=torch.float32, device=a.device)
    check(lib().molsde_graph_reduce(ptr(a), None if b is None else ptr(b.contiguous()), None if w is None else ptr(w.contiguous()),
                                    B, M, mode, ptr(out), stream_ptr(a)), "graph_reduce")
    return out


# ----------------------------------------------------------------------------------------------
# parameter containers (reference key layout) + kernel-driven forward
# ----------------------------------------------------------------------------------------------
class NodeNetwork_dense(nn.Module):
    """`layers/node_network_dense.py:25-44`: weight stored [in, out], glorot init."""

    def __init__(self, in_channels, out_channels, improved=False, bias=True):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.weight = nn.Parameter(torch.empty(in_channels, out_channels))
        self.bias = nn.Parameter(torch.zeros(out_channels))
        stdv = math.sqrt(6.0 / (in_channels + out_channels))
        self.weight.data.uniform_(-stdv, stdv)


class EdgeLayer(nn.Module):
    """`layers/edge_network_dense.py:33-53` with conv='MLP'."""

    def __init__(self, in_dim, attn_dim, out_dim, num_heads, conv):
        super().__init__()
        if conv != "MLP":
            raise NotImplementedError("conv='GCN' is not used by the reference scripts")
        self.num_heads, self.attn_dim, self.out_dim = num_heads, attn_dim, out_dim
        self.func_q = MultiLayerPerceptron(in_dim, [2 * attn_dim, 2 * attn_dim], activation="tanh")
        self.func_k = MultiLayerPerceptron(in_dim, [2 * attn_dim, 2 * attn_dim], activation="tanh")
        self.func_v = NodeNetwork_dense(in_dim, out_dim)


class EdgeNetwork_dense(nn.Module):
    """`layers/edge_network_dense.py:85-128`."""

    def __init__(self, num_linears, conv_input_dim, attn_dim, conv_output_dim, input_dim, output_dim, num_heads, conv):
        super().__init__()
        self.attn_dim, self.num_heads = attn_dim, num_heads
        self.in_ch, self.out_ch, self.conv_out = input_dim, output_dim, conv_output_dim
        self.attn = nn.ModuleList(EdgeLayer(conv_input_dim, attn_dim, conv_output_dim, num_heads, conv) for _ in range(input_dim))
        self.hidden_dim = 2 * max(input_dim, output_dim)
        self.mlp = MultiLayerPerceptron(2 * input_dim, [self.hidden_dim] * (num_linears - 1) + [output_dim], activation="elu")
        self.multi_channel = MultiLayerPerceptron(input_dim * conv_output_dim, [self.hidden_dim, conv_output_dim], activation="elu")
        self._packed = None

    def _pack(self):
        ver = (_abi.param_epoch(),) + tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self._packed is not None and self._packed[0] == ver:
            return self._packed[1]
        q1 = [a.func_q.layers[0] for a in self.attn] + [a.func_k.layers[0] for a in self.attn]
        q2 = [a.func_q.layers[1] for a in self.attn] + [a.func_k.layers[1] for a in self.attn]
        pk = {
            "w1": torch.cat([l.weight.detach().float() for l in q1], 0).contiguous(),   # [2C*W, Fin]: q_0..q_C-1, k_0..k_C-1
            "b1": torch.cat([l.bias.detach().float() for l in q1], 0).contiguous(),
            "w2": torch.stack([l.weight.detach().float() for l in q2], 0).contiguous(),  # [2C][W][W]
            "b2": torch.cat([l.bias.detach().float() for l in q2], 0).contiguous(),
            "wv": torch.cat([a.func_v.weight.detach().float().t() for a in self.attn], 0).contiguous(),  # [C*Fo, Fin]
            "bv": torch.cat([a.func_v.bias.detach().float() for a in self.attn], 0).contiguous(),
        }
        self._packed = (ver, pk)
        return pk

    def _fusable(self) -> bool:
        """`csrc/dense_fused.cu` covers the reference's pretraining / sampling configuration (3-layer pair MLP, <= 8 channels)."""
        ls = self.mlp.layers
        return (len(ls) == 3 and self.in_ch in (2, 8) and ls[0].out_features <= 16 and ls[1].out_features == ls[0].out_features
                and self.out_ch <= 8 and 2 * self.attn_dim <= 32 and (self.attn_dim // self.num_heads) % 4 == 0)

    @torch.no_grad()
    def forward_fused(self, x, adjc, flags, symmetric: bool):
        """Same values as `forward`, channel-major end to end: node-level GEMMs, `dense_attn_sym`, `dense_pair_mlp`.
        Returns (x_out [B,Nm,conv_out], adj_out [B,C',Nm,Nm]); no channels-last pair / all-channels tensors are written."""
        pk = self._pack()
        B, C, Nm = adjc.size(0), adjc.size(1), adjc.size(-1)
        W = 2 * self.attn_dim
        ds = self.attn_dim // self.num_heads
        s = stream_ptr(adjc)
        x2 = x.flatten(0, 1) if x.dim() == 3 else x
        rows, Fin = x2.size(0), x2.size(1)
        if Fin <= 16 and W == 32 and 2 * C <= 16 and x2.stride(1) == 1:   # narrow input (every layer but the first): one row kernel
            qk = torch.empty(rows, 2 * C * W, dtype=torch.float32, device=adjc.device)
            xw = torch.empty(rows, C * self.conv_out, dtype=torch.float32, device=adjc.device)
            check(lib().molsde_dense_node_side(x2.data_ptr(), rows, x2.stride(0), Fin, ptr(pk["w1"]), ptr(pk["b1"]), ptr(pk["w2"]),
                                               ptr(pk["b2"]), ptr(pk["wv"]), 2 * C, W, C * self.conv_out, ptr(qk), qk.stride(0), ptr(xw),
                                               xw.stride(0), s), "dense_node_side")
        else:
            h1 = linear(x2, pk["w1"], pk["b1"], act="tanh")
            qk = _grouped_linear(h1, pk["w2"], pk["b2"], 2 * C, W, W, "none")
            xw = linear(x2, pk["wv"])
        V = torch.empty(B * Nm, C * self.conv_out, dtype=torch.float32, device=adjc.device)
        _dense_gcn(adjc, C, xw, pk["bv"], self.conv_out, V, 0, "none")
        S = torch.empty(B, C, Nm, Nm, dtype=torch.float32, device=adjc.device)
        check(lib().molsde_dense_attn_sym(qk.data_ptr(), qk.data_ptr() + 4 * C * W, qk.stride(0), W, ds, ptr(flags), B, C, Nm, ptr(S), s),
              "dense_attn_sym")
        mc = self.multi_channel.layers
        Kv = V.size(1)
        if len(mc) == 2 and mc[0].out_features <= 16 and mc[1].out_features <= 16 and Kv % 4 == 0 and Kv <= 256:
            xo = torch.empty(rows, mc[1].out_features, dtype=torch.float32, device=adjc.device)
            check(lib().molsde_dense_multi_channel(ptr(V), rows, Kv, ptr(mc[0].weight.detach().float().contiguous()),
                                                   ptr(mc[0].bias.detach().float().contiguous()), mc[0].out_features,
                                                   ptr(mc[1].weight.detach().float().contiguous()),
                                                   ptr(mc[1].bias.detach().float().contiguous()), mc[1].out_features,
                                                   ptr(flags.reshape(-1).contiguous()), ptr(xo), s), "dense_multi_channel")
        else:
            xo = linear(linear(V, mc[0].weight, mc[0].bias, act="elu"), mc[1].weight, mc[1].bias, act="tanh", rowscale=flags)
        ml = self.mlp.layers
        w = [l.weight.detach().float().contiguous() for l in ml]
        bb = [l.bias.detach().float().contiguous() for l in ml]
        adj_out = torch.zeros(B, self.out_ch, Nm, Nm, dtype=torch.float32, device=adjc.device)   # padding pairs stay 0
        check(lib().molsde_dense_pair_mlp(ptr(S), ptr(adjc), ptr(flags), ptr(w[0]), ptr(bb[0]), ptr(w[1]), ptr(bb[1]), ptr(w[2]),
                                          ptr(bb[2]), B, C, w[0].size(0), self.out_ch, Nm, int(symmetric), ptr(adj_out), s),
              "dense_pair_mlp")
        return xo.view(B, Nm, -1), adj_out

    @torch.no_grad()
    def forward(self, x, adjc, flags, allc: Optional[torch.Tensor] = None, all_off: int = 0):
        """x [B,Nm,Fin] (row view allowed), adjc [B,C,Nm,Nm] -> (x_out [B,Nm,conv_out], adj_out [B,C',Nm,Nm])."""
        pk = self._pack()
        B, C, Nm = adjc.size(0), adjc.size(1), adjc.size(-1)
        W = 2 * self.attn_dim
        ds = self.attn_dim // self.num_heads
        s = stream_ptr(adjc)
        x2 = x.flatten(0, 1) if x.dim() == 3 else x
        h1 = linear(x2, pk["w1"], pk["b1"], act="tanh")                                   # func_q/k layer 0 (+tanh), all channels
        qk = _grouped_linear(h1, pk["w2"], pk["b2"], 2 * C, W, W, "none")                 # func_q/k layer 1
        xw = linear(x2, pk["wv"])                                                          # x @ func_v.weight
        V = torch.empty(B * Nm, C * self.conv_out, dtype=torch.float32, device=adjc.device)
        _dense_gcn(adjc, C, xw, pk["bv"], self.conv_out, V, 0, "none")                    # func_v (dense GCN)
        pair = torch.empty(B, Nm, Nm, 2 * C, dtype=torch.float32, device=adjc.device)
        check(lib().molsde_dense_attn(qk.data_ptr(), qk.data_ptr() + 4 * C * W, qk.stride(0), W, ds, ptr(adjc), B, C, Nm, ptr(pair), s),
              "dense_attn")                                                                # tanh attention + [A, adj] concat
        mc = self.multi_channel.layers
        xo = linear(linear(V, mc[0].weight, mc[0].bias, act="elu"), mc[1].weight, mc[1].bias, act="tanh", rowscale=flags)
        m = _mlp_rows(pair.view(-1, 2 * C), self.mlp, "elu")
        adj_out = torch.empty(B, self.out_ch, Nm, Nm, dtype=torch.float32, device=adjc.device)
        if allc is None:
            allc, all_off = torch.empty(B, Nm, Nm, self.out_ch, dtype=torch.float32, device=adjc.device), 0
        check(lib().molsde_dense_pair_post(ptr(m), ptr(flags), B, Nm, self.out_ch, ptr(adj_out), ptr(allc), allc.size(-1), all_off, s),
              "pair_post")
        return xo.view(B, Nm, -1), adj_out


class EdgeScoreNetwork_dense(nn.Module):
    """`invariant_scorenetwork_dense.py:40-93`."""

    def __init__(self, dim3D, nhid, num_layers, num_linears, c_init, c_hid, c_final, adim, num_heads, conv):
        super().__init__()
        if c_init != 2:
            raise NotImplementedError("pow_tensor kernel is built for c_init=2 (pretrain_MoleculeSDE.py:311)")
        self.c_init, self.c_hid, self.c_final, self.num_layers = c_init, c_hid, c_final, num_layers
        self.layers = nn.ModuleList()
        for i in range(num_layers):
            if i == 0:
                self.layers.append(EdgeNetwork_dense(num_linears, dim3D, nhid, nhid, c_init, c_hid, num_heads, conv))
            elif i == num_layers - 1:
                self.layers.append(EdgeNetwork_dense(num_linears, nhid, adim, nhid, c_hid, c_final, num_heads, conv))
            else:
                self.layers.append(EdgeNetwork_dense(num_linears, nhid, adim, nhid, c_hid, c_hid, num_heads, conv))
        self.fdim = c_hid * (num_layers - 1) + c_final + c_init
        self.final = MultiLayerPerceptron(self.fdim, [2 * self.fdim, 2 * self.fdim, 1], activation="silu")

    @torch.no_grad()
    def forward(self, x, adj, flags, scale: Optional[torch.Tensor] = None):
        """score [B,Nm,Nm] = final-MLP over all adjacency channels, zero diagonal, masked (x scale[b] if given)."""
        require_device(adj)
        adj = adj.contiguous().float()
        flags = flags.contiguous().float()
        B, Nm = adj.size(0), adj.size(1)
        s = stream_ptr(adj)
        fl = self.final.layers
        if os.environ.get("MOLSDE_DENSE_UNFUSED") != "1" and len(fl) == 3 and self.fdim <= 32 and fl[0].out_features <= 64 and fl[1].out_features <= 64 and fl[2].out_features == 1 \
                and all(l._fusable() for l in self.layers):
            return self._forward_fused(x, adj, flags, scale)
        allc = torch.empty(B, Nm, Nm, self.fdim, dtype=torch.float32, device=adj.device)
        adjc = torch.empty(B, 2, Nm, Nm, dtype=torch.float32, device=adj.device)
        check(lib().molsde_dense_pow2(ptr(adj), B, Nm, ptr(adjc), ptr(allc), self.fdim, 0, s), "pow2")
        off = self.c_init
        for lyr in self.layers:
            x, adjc = lyr(x, adjc, flags, allc, off)
            off += lyr.out_ch
        m = _mlp_rows(allc.view(-1, self.fdim), self.final, "silu")
        out = torch.empty(B, Nm, Nm, dtype=torch.float32, device=adj.device)
        check(lib().molsde_dense_edge_final(ptr(m), ptr(flags), None if scale is None else ptr(scale.contiguous()), B, Nm, ptr(out), s),
              "edge_final")
        return out


    def _forward_fused(self, x, adj, flags, scale):
        """Inference path on `csrc/dense_fused.cu`: channel-major adjacency stacks end to end; layers >= 1 see bitwise-symmetric
        inputs by construction (outputs of `dense_pair_mlp`), layer 0 and the head make no symmetry assumption on `adj`."""
        import ctypes
        B, Nm = adj.size(0), adj.size(1)
        s = stream_ptr(adj)
        adjc = torch.empty(B, 2, Nm, Nm, dtype=torch.float32, device=adj.device)
        check(lib().molsde_dense_pow2(ptr(adj), B, Nm, ptr(adjc), None, 0, 0, s), "pow2")
        stacks = [adjc]
        for li, lyr in enumerate(self.layers):
            x, adjc = lyr.forward_fused(x, adjc, flags, symmetric=li > 0)
            stacks.append(adjc)
        fl = self.final.layers
        w = [l.weight.detach().float().contiguous() for l in fl]
        bb = [l.bias.detach().float().contiguous() for l in fl]
        out = torch.zeros(B, Nm, Nm, dtype=torch.float32, device=adj.device)   # padding pairs and the diagonal stay 0
        n = len(stacks)
        ptrs = (ctypes.c_void_p * n)(*[t.data_ptr() for t in stacks])
        chs = (ctypes.c_int32 * n)(*[t.size(1) for t in stacks])
        check(lib().molsde_dense_edge_final_mlp(ptrs, chs, n, ptr(flags), None if scale is None else ptr(scale.contiguous()), ptr(w[0]),
                                                ptr(bb[0]), ptr(w[1]), ptr(bb[1]), ptr(w[2]), ptr(bb[2]), self.fdim, w[0].size(0),
                                                w[1].size(0), B, Nm, ptr(out), s), "dense_edge_final_mlp")
        return out


class NodeScoreNetwork_dense(nn.Module):
    """`invariant_scorenetwork_dense.py:96-131`."""

    def __init__(self, nfeat, depth, nhid, nout):
        super().__init__()
        self.nfeat, self.depth, self.nhid, self.nout = nfeat, depth, nhid, nout
        self.layers = nn.ModuleList(NodeNetwork_dense(nfeat if i == 0 else nhid, nhid) for i in range(depth))
        self.fdim = nfeat + depth * nhid
        self.final = MultiLayerPerceptron(self.fdim, [2 * self.fdim, 2 * self.fdim, nout], activation="silu")

    @torch.no_grad()
    def forward(self, x, adj, flags, scale: Optional[torch.Tensor] = None):
        require_device(adj)
        adj = adj.contiguous().float()
        flags_arg = flags
        flags = flags.contiguous().float()
        B, Nm = adj.size(0), adj.size(1)
        xs = torch.empty(B * Nm, self.fdim, dtype=torch.float32, device=adj.device)  # cat(x_list), built in place
        xs[:, :self.nfeat].copy_(x.reshape(B * Nm, self.nfeat))
        cur, off = xs[:, :self.nfeat], self.nfeat
        for lyr in self.layers:
            xw = linear(cur, lyr.weight.detach().t().contiguous())
            _dense_gcn(adj, 1, xw, lyr.bias.detach().float().contiguous(), self.nhid, xs, off, "tanh")  # tanh(GCN), :121-122
            cur, off = xs[:, off:off + self.nhid], off + self.nhid
        fl = self.final.layers
        comp = self._valid_rows(flags_arg, flags)
        if comp is not None:
            # The final MLP (fdim -> 2 fdim -> 2 fdim -> nout: the largest GEMMs of a score evaluation) is row-wise and its output
            # is masked: run it on the VALID atom rows only (a padded batch of 256 x 64 holds ~10.8k atoms in 16.4k rows) and put
            # the rows back with the padding rows reading a zero row.  Same kernels, same per-row arithmetic: bit-identical.
            idx, inv, gidx, N = comp
            h = linear(gather_rows(xs, idx), fl[0].weight, fl[0].bias, act="silu")
            h = linear(h, fl[1].weight, fl[1].bias, act="silu")
            outc = torch.empty(N + 1, self.nout, dtype=torch.float32, device=adj.device)
            outc[N:].zero_()
            linear(h, fl[2].weight, fl[2].bias, rowscale=None if scale is None else scale.index_select(0, gidx), out=outc[:N])
            return gather_rows(outc, inv).view(B, Nm, self.nout)
        h = linear(xs, fl[0].weight, fl[0].bias, act="silu")
        h = linear(h, fl[1].weight, fl[1].bias, act="silu")
        rs = flags.reshape(-1) if scale is None else (flags * scale[:, None]).reshape(-1)
        out = linear(h, fl[2].weight, fl[2].bias, rowscale=rs)                          # mask_x (and -1/std)
        return out.view(B, Nm, self.nout)

    def _valid_rows(self, key_tensor, flags):
        """Row indices of the valid atoms of a padded batch, cached on the identity (+ version) of the caller's `flags` tensor
        (static along a sampling trajectory; the cache holds the tensor, so its address cannot be reused).  None when the padding
        is small, the batch is small, or the index would have to be built (host sync) while a CUDA graph is being captured."""
        if os.environ.get("MOLSDE_DENSE_NO_COMPACT") == "1":
            return None
        rows = flags.numel()
        if rows < 4096:
            return None
        c = self.__dict__.get("_rows_cache")
        if c is None or c[0] is not key_tensor or c[1] != key_tensor._version:
            if torch.cuda.is_current_stream_capturing():
                return None
            valid = flags.reshape(-1) > 0
            if not bool(((flags == 0) | (flags == 1)).all()):     # the compact path drops the mask multiply: exact only for 0/1 flags
                c = (key_tensor, key_tensor._version, None)
            else:
                idx = valid.nonzero().reshape(-1)
                N = int(idx.numel())
                inv = torch.full((rows,), N, dtype=torch.int64, device=flags.device)
                inv[idx] = torch.arange(N, device=flags.device)
                c = (key_tensor, key_tensor._version, (idx, inv, idx // flags.size(1), N) if 0 < N <= 0.8 * rows else None)
            self.__dict__["_rows_cache"] = c
        return c[2]


class SDEModel3Dto2D_node_adj_dense(nn.Module):
    concat_embedding = False   # True in SDEModel3Dto2D_node_adj_dense_02: cat([embedding_3D, embedding_X]) instead of their sum

    def __init__(self, dim3D, nhid, num_layers, num_linears, c_hid, c_final, adim, emb_dim, beta_min, beta_max,
                 num_diffusion_timesteps, c_init=1, num_heads=4, conv="MLP", noise_mode="discrete", SDE_type="VE",
                 num_class_X=119, noise_on_one_hot=True):
        super().__init__()
        if not noise_on_one_hot:
            raise NotImplementedError("noise_on_one_hot=False is not the pretraining default (config.py:121-123)")
        self.emb_dim, self.beta_min, self.beta_max = emb_dim, beta_min, beta_max
        self.num_diffusion_timesteps, self.nfeat, self.nhid = num_diffusion_timesteps, dim3D, nhid
        self.num_layers, self.num_linears, self.c_init, self.c_hid, self.c_final = num_layers, num_linears, c_init, c_hid, c_final
        self.adim, self.num_heads, self.conv, self.noise_mode, self.SDE_type = adim, num_heads, conv, noise_mode, SDE_type
        if SDE_type == "VE":
            self.sde_x = VESDE(sigma_min=beta_min, sigma_max=beta_max, N=num_diffusion_timesteps)
            self.sde_adj = VESDE(sigma_min=beta_min, sigma_max=beta_max, N=num_diffusion_timesteps)
        elif SDE_type == "VP":
            self.sde_x = VPSDE(beta_min=beta_min, beta_max=beta_max, N=num_diffusion_timesteps)
            self.sde_adj = VPSDE(beta_min=beta_min, beta_max=beta_max, N=num_diffusion_timesteps)
        else:
            raise NotImplementedError(SDE_type)
        self.num_class_X, self.noise_on_one_hot = num_class_X, noise_on_one_hot
        self.embedding_X = nn.Linear(num_class_X, dim3D)
        self.embedding_3D = nn.Linear(dim3D, dim3D)
        net_in = 2 * dim3D if self.concat_embedding else dim3D
        self.edge_score_network = EdgeScoreNetwork_dense(dim3D=net_in, nhid=nhid, num_layers=num_layers, num_linears=num_linears,
                                                         c_init=c_init, c_hid=c_hid, c_final=c_final, adim=adim, num_heads=4,
                                                         conv=conv)
        self.node_score_network = NodeScoreNetwork_dense(nfeat=net_in, depth=num_layers, nhid=nhid, nout=num_class_X)

    # ---- `embedding_3D(representation) + embedding_X(x)` (:156; inference_3D_to_2D:228,240; SDE_dense.py:88,99) ----
    @torch.no_grad()
    def embed_3d(self, representation):
        return linear(representation.contiguous().float(), self.embedding_3D.weight, self.embedding_3D.bias)

    @torch.no_grad()
    def embed(self, representation, x, rep3d=None):
        """`rep3d`: a cached `embed_3d(representation)` (the 3D part does not change along a sampling trajectory)."""
        if rep3d is None:
            rep3d = self.embed_3d(representation)
        if self.concat_embedding:   # `_dense_02` (:333): [B,Nm,2F] = cat([embedding_3D(rep), embedding_X(x)], -1)
            F = self.nfeat
            out = torch.empty(*x.shape[:-1], 2 * F, dtype=torch.float32, device=x.device)
            out[..., :F].copy_(rep3d.view(*x.shape[:-1], F))
            linear(x.contiguous().float(), self.embedding_X.weight, self.embedding_X.bias, out=out[..., F:])
            return out
        return linear(x.contiguous().float(), self.embedding_X.weight, self.embedding_X.bias, residual=rep3d)

    def get_score_fn(self, sde, model, train=True, continuous=True):
        """`:68-99`: returns score_fn(x, adj, flags, t) = -model(x, adj, flags) / std(t)[:,None,None]."""
        if not train:
            model.eval()
        if not isinstance(sde, (VPSDE, subVPSDE, VESDE)):
            raise NotImplementedError(f"SDE class {sde.__class__.__name__} not supported.")

        def score_fn(x, adj, flags, t):
            if not continuous:
                raise NotImplementedError("Discrete not supported")
            std = sde.marGINal_prob(torch.zeros(t.numel(), 1, 1, device=t.device), t)[1]
            return model(x, adj, flags, scale=(-1.0 / std).float())

        return score_fn

    @torch.no_grad()
    def dense_inputs(self, node_3D_repr, data):
        """to_dense_adj / to_dense_batch prologue (:118-134): (adj, rep_dense, z_dense, flags, Nm)."""
        require_device(node_3D_repr)
        batch = data.batch
        cached = getattr(data, "_molsde_dense_dims", None)                 # static per batch: keeps replays sync-free
        if cached is None:
            B = int(batch.max().item()) + 1                                # :124 (host sync, as the reference)
            node_ptr = segment_ptr(batch, B)
            Nm = int((node_ptr[1:] - node_ptr[:-1]).max().item())          # :127
            try:
                data._molsde_dense_dims = (B, Nm, node_ptr)
            except AttributeError:
                pass
        else:
            B, Nm, node_ptr = cached
        if Nm > 64:
            raise _abi.MolsdeError("dense 3D->2D kernels support at most 64 atoms per graph")
        dev, s = node_3D_repr.device, stream_ptr(node_3D_repr)
        ei = data.edge_index.contiguous()
        bond = data.edge_attr[:, 0].contiguous()
        adj = torch.empty(B, Nm, Nm, dtype=torch.float32, device=dev)
        check(lib().molsde_to_dense_adj(ptr(ei), ei.size(1), None, ptr(bond), 1.0, ptr(node_ptr), B, Nm, ptr(adj), s), "to_dense_adj")
        h = node_3D_repr.detach().float().contiguous()
        rep = torch.empty(B, Nm, h.size(1), dtype=torch.float32, device=dev)
        check(lib().molsde_to_dense_batch(ptr(h), ptr(node_ptr), B, Nm, h.size(1), ptr(rep), h.size(1), s), "to_dense_batch")
        zf = data.x[:, 0].float().contiguous().view(-1, 1)
        zd = torch.empty(B, Nm, 1, dtype=torch.float32, device=dev)
        check(lib().molsde_to_dense_batch(ptr(zf), ptr(node_ptr), B, Nm, 1, ptr(zd), 1, s), "to_dense_batch(z)")
        return adj, rep, zd.view(B, Nm).long(), node_flags(adj), Nm

    def forward(self, node_3D_repr, data, continuous, train, reduce_mean, anneal_power, *, draws=None):
        """DSM losses (loss_x, loss_adj), `:101-179`; `draws` = (randint, randn adj, randn one-hot) injects the three random
        draws in reference order.  With autograd enabled (training loop) the call is one autograd node (`autograd.py`,
        `pretrain.tape_3d2d`); under no_grad the forward kernels compute the values."""
        if not reduce_mean:
            raise NotImplementedError("reduce_mean=False only pairs with noise_on_one_hot=False")
        from . import autograd as AG
        if train and AG.grad_mode(self, node_3D_repr):
            if not continuous:
                raise NotImplementedError("Discrete not supported")
            from .pretrain import tape_3d2d

            def build(tp, ins, P):
                c = [1.0, 1.0]
                lx, la = tape_3d2d(tp, self, P, ins[0], data, anneal_power, draws, coef=c)

                def seed(gouts):
                    for i in range(2):
                        c[i] = float(gouts[i].reshape(-1)[0].item()) if gouts[i] is not None else 0.0
                return [lx.reshape(()), la.reshape(())], seed
            return AG.apply(self, build, [node_3D_repr])
        with torch.no_grad():
            return self._forward_value(node_3D_repr, data, continuous, train, reduce_mean, anneal_power, draws=draws)

    def _forward_value(self, node_3D_repr, data, continuous, train, reduce_mean, anneal_power, *, draws=None):
        adj, rep, zd, flags, Nm = self.dense_inputs(node_3D_repr, data)
        dev, B, N, s = adj.device, adj.size(0), self.num_diffusion_timesteps, stream_ptr(adj)
        if self.noise_mode == "discrete":
            th = torch.randint(0, N, size=(B // 2 + 1,), device=dev) if draws is None else draws[0].to(dev)
            t = torch.cat([th, N - th - 1], dim=0)[:B]
            t = t / N * (1 - EPSILON) + EPSILON
        else:
            t = torch.rand(B, device=dev) * (1 - EPSILON) + EPSILON
        z_adj = gen_noise(adj, flags, sym=True, raw=None if draws is None else draws[1])
        std_adj = self.sde_adj.marGINal_prob(torch.zeros(B, 1, 1, device=dev), t)[1].float().contiguous()
        coef_adj = self.sde_adj.mean_coeff(t).float().contiguous()
        p_adj = torch.empty_like(adj)
        check(lib().molsde_dense_perturb_adj(ptr(adj), ptr(z_adj), ptr(flags), ptr(coef_adj), ptr(std_adj), B, Nm, ptr(p_adj), s),
              "perturb_adj")
        K = self.num_class_X
        raw_x = torch.randn(B, Nm, K, device=dev) if draws is None else draws[2].to(dev).float().contiguous()
        std_x = self.sde_x.marGINal_prob(torch.zeros(B, 1, 1, device=dev), t)[1].float().contiguous()
        coef_x = self.sde_x.mean_coeff(t).float().contiguous()
        z_x, p_x = torch.empty(B, Nm, K, device=dev), torch.empty(B, Nm, K, device=dev)
        check(lib().molsde_dense_perturb_onehot(ptr(zd.contiguous()), ptr(raw_x), ptr(flags), ptr(coef_x), ptr(std_x), B, Nm, K,
                                                ptr(z_x), ptr(p_x), s), "perturb_onehot")
        emb = self.embed(rep, p_x)
        score_adj = self.get_score_fn(self.sde_adj, self.edge_score_network, train=train, continuous=continuous)(emb, p_adj, flags, t)
        score_x = self.get_score_fn(self.sde_x, self.node_score_network, train=train, continuous=continuous)(emb, p_adj, flags, t)
        wx = None if anneal_power == 0 else (std_x ** anneal_power).contiguous()
        wa = None if anneal_power == 0 else (std_adj ** anneal_power).contiguous()
        losses_x = graph_reduce(score_x, z_x, wx, 1)
        losses_adj = graph_reduce(score_adj, z_adj, wa, 1)
        return torch.mean(losses_x), torch.mean(losses_adj)


class SDEModel3Dto2D_node_adj_dense_02(SDEModel3Dto2D_node_adj_dense):
    """`SDE_model_3D_to_2D_node_adj_dense.py:182-350`: the variant whose score networks see the CONCATENATION of
    `embedding_3D(representation)` and `embedding_X(perturbed x)` (600 features) instead of their sum (`:333`).  Same constructor,
    `forward` / `get_score_fn` signatures and state_dict keys as the reference class; the same kernels (the first-layer GEMMs and the
    node network's final MLP are simply wider: 664 -> 1328 -> 1328 -> 119)."""
    concat_embedding = True
