"""One MoleculeSDE pretraining iteration (`examples/pretrain_MoleculeSDE.py:105-152,331-337`) on our kernels:
GIN -> SchNet -> dual_CL -> SDEModel2Dto3D_02 -> SDEModel3Dto2D_node_adj_dense -> backward -> (all-reduce) -> Adam.

`ParamStore` flattens the trainable parameters of the four modules into ONE fp32 buffer (the modules' `.data` become
views, so `state_dict()` keeps the reference keys) with a matching flat gradient buffer: the unit of the data-parallel
NCCL all-reduce and of the single-launch Adam step.  The per-model `tape_*` functions build the forward on a `Tape`.
"""
from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch
from torch import nn

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr
from .graph import csr_by_target
from .tape import Index, Tape, Var, _p


class ParamStore:
    def __init__(self, modules: Dict[str, nn.Module], device: torch.device):
        self.modules, self.dev = modules, device
        self.index: Dict[str, Dict[str, tuple]] = {}
        total = 0
        for mname, m in modules.items():
            self.index[mname] = {}
            for pname, p in m.named_parameters():
                if not p.requires_grad:
                    continue
                n = p.numel()
                self.index[mname][pname] = (total, n, tuple(p.shape))
                total += (n + 3) // 4 * 4  # keep every tensor 16-byte aligned
        self.numel = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=device)
        self.grad = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=device)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=device)
        self.step_count = 0
        self.ranges = {}
        for mname, m in modules.items():
            params = dict(m.named_parameters())
            offs = [v[0] for v in self.index[mname].values()]
            self.ranges[mname] = (min(offs), max(o + (n + 3) // 4 * 4 for o, n, _ in self.index[mname].values())) if offs else (0, 0)
            for pname, (off, n, shape) in self.index[mname].items():
                view = self.flat[off:off + n].view(shape)
                view.copy_(params[pname].data.to(device))
                params[pname].data = view
            m.to(device)  # buffers (running stats, frozen Fourier frequencies); parameters already live in `flat`
            for pname, (off, n, shape) in self.index[mname].items():
                assert params[pname].data.data_ptr() == self.flat[off:off + n].data_ptr(), "module.to() re-allocated a parameter"

    def vars(self, mname: str) -> Dict[str, Var]:
        out = {}
        for pname, (off, n, shape) in self.index[mname].items():
            out[pname] = Var(self.flat[off:off + n].view(shape), True, self.grad[off:off + n].view(shape))
        return out

    def grad_view(self, mname: str, pname: str) -> torch.Tensor:
        off, n, shape = self.index[mname][pname]
        return self.grad[off:off + n].view(shape)

    def zero_grad(self) -> None:
        self.grad.zero_()  # cudaMemsetAsync

    def all_reduce(self) -> float:
        """Data-parallel gradient exchange: one NCCL all-reduce (sum) over the flat buffer; returns 1/world for Adam."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grad, op=dist.ReduceOp.SUM)
            return 1.0 / dist.get_world_size()
        return 1.0

    def adam_step(self, lr: float = 1e-4, lr_scale: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                  weight_decay: float = 0.0, grad_scale: float = 1.0) -> None:
        """torch.optim.Adam over the four parameter groups (`pretrain_MoleculeSDE.py:331-337`); groups with the same lr
        share one launch."""
        self.step_count += 1
        s = torch.cuda.current_stream(self.dev).cuda_stream
        scales = {m: (lr_scale or {}).get(m, 1.0) for m in self.modules}
        if len(set(scales.values())) == 1:
            spans = [(0, self.numel, lr * next(iter(scales.values())))]
        else:
            spans = [(a, b, lr * scales[m]) for m, (a, b) in self.ranges.items() if b > a]
        for a, b, l in spans:
            check(lib().molsde_adam_step(self.flat[a:b].data_ptr(), self.grad[a:b].data_ptr(), self.exp_avg[a:b].data_ptr(),
                                         self.exp_avg_sq[a:b].data_ptr(), b - a, l, betas[0], betas[1], eps, weight_decay,
                                         self.step_count, grad_scale, s), "adam_step")


# ======================================================================================================
# graph bookkeeping shared by the message-passing tapes
# ======================================================================================================
class EdgeSet:
    """Edges in CSR-by-target order with both gather indices: `tgt` (grouped, perm = identity) and `src` (CSR by
    source from the same kernels)."""

    def __init__(self, rowptr: torch.Tensor, src: torch.Tensor, batch: torch.Tensor, num_graphs: int):
        dev, N, E = rowptr.device, rowptr.numel() - 1, src.numel()
        s = stream_ptr(rowptr)
        tgt = torch.empty(max(E, 1), dtype=torch.int32, device=dev)
        check(lib().molsde_expand_rowptr(ptr(rowptr), N, ptr(tgt), s), "expand_rowptr")
        tgt = tgt[:E]
        self.N, self.E, self.rowptr = N, E, rowptr
        self.tgt = Index(tgt, rowptr, None, N)
        flipped = torch.stack([tgt.long(), src.long()])  # [2,E]: "target" row of csr_by_target = our source
        by_src = csr_by_target(flipped, batch, num_graphs)
        self.src = Index(src, by_src.rowptr, by_src.perm, N)


# ======================================================================================================
# SDEModel2Dto3D_02.forward (SDE_model_2D_to_3D.py:306-391)
# ======================================================================================================
def _slice_cols(v: Var, a: int, b: int) -> Var:
    return Var(v.data[:, a:b], v.needs, v.grad[:, a:b] if v.grad is not None else None)


def tape_2d3d(tp: Tape, model, P: Dict[str, Var], h2d: Var, data, anneal_power: float = 0.0, draws: Optional[dict] = None,
              coef: float = 1.0) -> torch.Tensor:
    """Records the 2D->3D denoising score-matching loss on `tp` (train mode: BatchNorm batch statistics, attention and
    FFN dropout) and returns the loss tensor [1]; its backward seeds d loss = `coef`."""
    L, dev = tp.L, tp.dev
    draws = draws or {}
    pos = data.positions.detach().float().contiguous()
    require_device(pos)
    prep = model.prepared(data)
    es = getattr(prep, "_edge_set", None)
    if es is None:
        es = prep._edge_set = EdgeSet(prep.csr.rowptr, prep.csr.col, data.batch, data.num_graphs)
    N, E, B, s = es.N, es.E, data.num_graphs, tp.s
    T = model.num_diffusion_timesteps
    noise = draws.get("noise")
    noise = torch.randn_like(pos) if noise is None else noise.to(dev).float().contiguous()
    ts = draws.get("time_step")
    ts = torch.randint(0, T, size=(B // 2 + 1,), device=dev) if ts is None else ts.to(dev)
    ts = torch.cat([ts, T - ts - 1], dim=0)[:B]
    if model.SDE_type in ("VE", "VP"):
        ts = ts / T * (1 - 1e-6) + 1e-6
    t_pos = ts.index_select(0, data.batch)
    coeff, std = model.sde_pos.marGINal_prob(torch.ones(N, 1, device=dev), t_pos)
    from .sde import VPSDE
    coeff = coeff.reshape(-1).float().contiguous() if isinstance(model.sde_pos, VPSDE) else None
    std = std.float().contiguous()
    pos_p = tp.empty(N, 3)
    tp._call(L.molsde_perturb_rows, ptr(pos), ptr(coeff), ptr(std), ptr(noise), N, 3, ptr(pos_p), s, what="perturb_rows")

    p_drop = model.score_network.dropout
    masks = draws.get("dropout")
    if not model.training:
        keep = [(None, None)] * 4
    elif masks is None:
        keep = [((torch.rand(E, 8, device=dev) >= p_drop).float(), (torch.rand(N, 32, device=dev) >= p_drop).float())
                for _ in range(4)]
    else:
        assert prep.csr.perm is not None, "injected attention masks need the edge permutation (build the CSR from edge_index)"
        perm = prep.csr.perm.long()
        keep = [(m[0].to(dev).float()[perm].contiguous(), m[1].to(dev).float().contiguous()) for m in masks]

    # ---- edge_2D_emb: Linear(600,300) on cat[h[row], h[col]] factored per node, BatchNorm (batch stats) + ReLU, Linear(300,32)
    F = model.emb_dim
    W0, b0 = P["edge_2D_emb.0.weight"], P["edge_2D_emb.0.bias"]
    U = tp.linear(h2d, _slice_cols(W0, 0, F), b0)
    V = tp.linear(h2d, _slice_cols(W0, F, 2 * F), None)
    pre = tp.gather_pair(U, es.src, V, es.tgt)
    bn = model.edge_2D_emb[1]
    if model.training:
        act = tp.batchnorm(pre, P["edge_2D_emb.1.weight"], P["edge_2D_emb.1.bias"], bn.running_mean, bn.running_var, bn.eps,
                           bn.momentum, relu=True)
        bn.num_batches_tracked += 1
    else:
        raise NotImplementedError("tape_2d3d is the training step; eval-mode losses go through SDEModel2Dto3D_02.forward")
    e2d = tp.linear(act, P["edge_2D_emb.3.weight"], P["edge_2D_emb.3.bias"])

    # ---- geometric features at the perturbed positions (no gradient into positions / frozen frequencies)
    gfd, gfi, gfj = tp.empty(E, 64), tp.empty(E, 128), tp.empty(E, 128)
    emb = Var(tp.empty(E, 66), False)
    basis = tp.empty(E, 9)
    tp._call(L.molsde_sde2d3d_edge_geom, ptr(pos_p), ptr(es.src.idx), ptr(es.tgt.idx), E, ptr(model.dist_gaussian_fourier.W.data),
             ptr(model.coff_gaussian_fourier.W.data), ptr(gfd), ptr(gfi), ptr(gfj), ptr(emb.data), ptr(basis), s, what="edge_geom")
    inv3d = tp.linear(Var(gfd), P["input_mlp.layers.0.weight"], P["input_mlp.layers.0.bias"])
    tp.linear(Var(gfi), P["coff_mlp.weight"], P["coff_mlp.bias"], into=emb, col0=2)
    tp.linear(Var(gfj), P["coff_mlp.weight"], P["coff_mlp.bias"], into=emb, col0=2 + 32)
    hid = tp.linear(emb, P["project.layers.0.weight"], P["project.layers.0.bias"], act="silu")
    frame = tp.linear(hid, P["project.layers.1.weight"], P["project.layers.1.bias"])
    edge_attr = tp.mul(inv3d, e2d, frame)
    x = tp.linear(h2d, P["node_emb.layers.0.weight"], P["node_emb.layers.0.bias"])

    # ---- EquivariantScoreNetwork (equivariant_scorenetwork.py:121-169)
    grad = tp.empty(N, 3)
    dyn_list = []
    li = 0
    for m in range(2):
        for c in range(2):
            pf = f"score_network.gnn_layers.{m}.{c}."
            x = _gat_layer(tp, P, pf, x, edge_attr, es, keep[li], p_drop if model.training else 0.0)
            li += 1
            if c == 0:
                x = tp.act(x, "silu")
        hsum = tp.gather_pair(x, es.src, x, es.tgt)
        bp = f"score_network.basis_mlp_modules.{m}."
        W1 = P[bp + "0.weight"]
        y1 = tp.linear(hsum, _slice_cols(W1, 0, 32), P[bp + "0.bias"])
        y2 = tp.linear(edge_attr, _slice_cols(W1, 32, 64), None)
        hcoef = tp.act(tp.add(y1, y2), "silu")
        dyn = tp.linear(hcoef, P[bp + "2.weight"], P[bp + "2.bias"])
        tp._call(L.molsde_equi_fwd, ptr(dyn.data), ptr(basis), ptr(es.rowptr), N, int(m > 0), ptr(grad), s, what="equi_fwd")
        dyn_list.append(dyn)

    w = None if anneal_power == 0 else (std ** anneal_power).contiguous()
    per_graph, loss = tp.empty(B), tp.empty(1)
    tp._call(L.molsde_dsm_pos_loss, ptr(grad), ptr(noise), ptr(w), ptr(prep.node_ptr), B, ptr(per_graph), ptr(loss), s,
             what="dsm_pos_loss")
    node2graph = data.batch.to(torch.int32)

    def bwd():
        dgrad = tp.empty(N, 3)
        tp._call(L.molsde_dsm_pos_loss_bwd, ptr(grad), ptr(noise), ptr(w), ptr(prep.node_ptr), ptr(node2graph), N, B, float(coef),
                 ptr(dgrad), s, what="dsm_pos_loss_bwd")
        for dyn in dyn_list:
            dd = tp.empty(E, 3)
            tp._call(L.molsde_equi_bwd, ptr(dgrad), ptr(basis), ptr(es.rowptr), ptr(es.tgt.idx), E, ptr(dd), s, what="equi_bwd")
            tp.accum(dyn, dd)
    # recorded last => runs first in the reverse sweep, seeding dyn.grad before the basis-MLP backward
    tp.ops.append(bwd)
    return loss


def _gat_layer(tp: Tape, P: Dict[str, Var], pf: str, x: Var, edge_attr: Var, es: EdgeSet, keep, p_drop: float) -> Var:
    """GATLayer.forward (`equivariant_scorenetwork.py:34-40`) over TransformerConv heads 8 x 4."""
    L, N, E, s = tp.L, es.N, es.E, tp.s
    attn_keep, ffn_keep = keep
    qkvs = Var(tp.empty(N, 128), False)
    tp.linear(x, P[pf + "MHA.lin_query.weight"], P[pf + "MHA.lin_query.bias"], into=qkvs, col0=0)
    tp.linear(x, P[pf + "MHA.lin_key.weight"], P[pf + "MHA.lin_key.bias"], into=qkvs, col0=32)
    tp.linear(x, P[pf + "MHA.lin_value.weight"], P[pf + "MHA.lin_value.bias"], into=qkvs, col0=64)
    tp.linear(x, P[pf + "MHA.lin_skip.weight"], P[pf + "MHA.lin_skip.bias"], into=qkvs, col0=96)
    eproj = tp.linear(edge_attr, P[pf + "MHA.lin_edge.weight"], None)
    alpha, out = tp.empty(E, 8), Var(tp.empty(N, 32), True)
    tp._call(L.molsde_tconv_fwd, ptr(qkvs.data), ptr(eproj.data), ptr(es.rowptr), ptr(es.src.idx), N, _p(attn_keep), p_drop,
             ptr(alpha), ptr(out.data), s, what="tconv_fwd")

    def bwd():
        if out.grad is None:
            return
        dq, dkv, de = tp.empty(N, 128), tp.empty(E, 64), tp.empty(E, 32)
        tp._call(L.molsde_tconv_bwd, ptr(qkvs.data), ptr(eproj.data), ptr(es.rowptr), ptr(es.src.idx), ptr(es.src.ptr),
                 ptr(es.src.perm), N, _p(attn_keep), p_drop, ptr(alpha), ptr(out.grad), ptr(dq), ptr(dkv), ptr(de), s,
                 what="tconv_bwd")
        tp.accum(qkvs, dq)
        tp.accum(eproj, de)
    tp.ops.append(bwd)
    x1 = tp.add(x, tp.layernorm(out, P[pf + "norm1.weight"], P[pf + "norm1.bias"]))
    h = tp.linear(x1, P[pf + "FFN.0.weight"], P[pf + "FFN.0.bias"], act="silu")
    if ffn_keep is not None:
        h = tp.scale_mask(h, ffn_keep, 1.0 / (1.0 - p_drop))
    h2 = tp.linear(h, P[pf + "FFN.3.weight"], P[pf + "FFN.3.bias"])
    return tp.add(x1, tp.layernorm(h2, P[pf + "norm2.weight"], P[pf + "norm2.bias"]))


# ======================================================================================================
# GIN 2D encoder (Geom3D/models/molecule_gnn_model.py:132-197)
# ======================================================================================================
def _concat_tables(tp: Tape, P: Dict[str, Var], names) -> Var:
    """The embedding tables of one encoder as ONE [rows, emb] table.  Inside a ParamStore they are adjacent in the flat
    buffer (a zero-copy view, gradient included); otherwise (inference) they are concatenated once."""
    vs = [P[n] for n in names]
    cols = vs[0].data.shape[1]
    adjacent = all(vs[i].data.data_ptr() + vs[i].data.numel() * 4 == vs[i + 1].data.data_ptr() for i in range(len(vs) - 1))
    rows = sum(v.data.shape[0] for v in vs)
    if adjacent and vs[0].grad is not None:
        base = vs[0].data
        data = torch.as_strided(base, (rows, cols), (cols, 1))
        grad = torch.as_strided(vs[0].grad, (rows, cols), (cols, 1))
        return Var(data, True, grad)
    assert not any(v.needs for v in vs), "trainable embedding tables must live in a ParamStore"
    return Var(torch.cat([v.data for v in vs], dim=0).contiguous(), False)


def _keys(idx: torch.Tensor, dims) -> torch.Tensor:
    """int64 [R,F] categorical features -> int32 keys into the concatenated table."""
    off = torch.tensor([0] + list(torch.tensor(dims).cumsum(0)[:-1]), dtype=torch.int64, device=idx.device)
    return (idx + off[None, :]).to(torch.int32).contiguous()


def tape_gin(tp: Tape, model, P: Dict[str, Var], x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor,
             cache: Optional[dict], batch: Optional[torch.Tensor] = None, num_graphs: int = 1) -> Var:
    """GNN.forward, GIN / JK=last / dropout 0.  `cache` (a dict living on the batch) keeps the index structures;
    `batch` (node -> graph, ascending) only speeds up the CSR build."""
    from .gnn import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS
    from .tape import bucket_index
    dev = x.device
    require_device(x)
    train = any(v.needs for v in P.values())
    cache = cache if cache is not None else {}
    if "gin" not in cache:
        N = x.size(0)
        if batch is None:
            batch, num_graphs = torch.zeros(N, dtype=torch.long, device=dev), 1
        csr = csr_by_target(edge_index, batch, num_graphs)
        es = EdgeSet(csr.rowptr, csr.col, batch, num_graphs) if train else None
        akeys = _keys(x, ATOM_FEATURE_DIMS)
        ekeys = _keys(edge_attr[csr.perm.long()], BOND_FEATURE_DIMS)   # CSR edge order
        aidx = bucket_index(akeys.reshape(-1).long(), sum(ATOM_FEATURE_DIMS)) if train else None
        eidx = bucket_index(ekeys.reshape(-1).long(), sum(BOND_FEATURE_DIMS)) if train else None
        cache["gin"] = (csr, es, akeys, ekeys, aidx, eidx)
    csr, es, akeys, ekeys, aidx, eidx = cache["gin"]
    src = es.src if es is not None else Index(csr.col, None, None, x.size(0))
    tgt = es.tgt if es is not None else None
    T_atom = _concat_tables(tp, P, [f"atom_encoder.atom_embedding_list.{i}.weight" for i in range(len(ATOM_FEATURE_DIMS))])
    h = tp.embed_sum(T_atom, akeys, aidx)
    for l in range(model.num_layer):
        pf = f"gnns.{l}."
        T_bond = _concat_tables(tp, P, [pf + f"bond_encoder.bond_embedding_list.{i}.weight" for i in range(len(BOND_FEATURE_DIMS))])
        pre = tp.gin_aggregate(h, T_bond, ekeys, eidx, csr.rowptr, src, tgt, P[pf + "eps"])
        bn1, bn2 = model.gnns[l].mlp[1], model.batch_norms[l]
        z = tp.linear(pre, P[pf + "mlp.0.weight"], P[pf + "mlp.0.bias"])
        last = l == model.num_layer - 1
        if model.training:
            z = tp.batchnorm(z, P[pf + "mlp.1.weight"], P[pf + "mlp.1.bias"], bn1.running_mean, bn1.running_var, bn1.eps,
                             bn1.momentum, relu=True)
            bn1.num_batches_tracked += 1
        else:
            z = tp.batchnorm_eval(z, P[pf + "mlp.1.weight"], P[pf + "mlp.1.bias"], bn1.running_mean, bn1.running_var, bn1.eps, relu=True)
        z = tp.linear(z, P[pf + "mlp.3.weight"], P[pf + "mlp.3.bias"])
        if model.training:
            h = tp.batchnorm(z, P[f"batch_norms.{l}.weight"], P[f"batch_norms.{l}.bias"], bn2.running_mean, bn2.running_var,
                             bn2.eps, bn2.momentum, relu=not last)
            bn2.num_batches_tracked += 1
        else:
            h = tp.batchnorm_eval(z, P[f"batch_norms.{l}.weight"], P[f"batch_norms.{l}.bias"], bn2.running_mean, bn2.running_var,
                                  bn2.eps, relu=not last)
    return h


# ======================================================================================================
# SchNet (Geom3D/models/schnet.py:85-125), return_latent representation
# ======================================================================================================
def tape_schnet(tp: Tape, model, P: Dict[str, Var], z: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor, num_graphs: int,
                cache: Optional[dict]) -> Var:
    """Node representation h [N, hidden] of SchNet.forward(return_latent=True) with every intermediate kept."""
    from .graph import radius_graph
    from .tape import bucket_index
    L, dev = tp.L, pos.device
    require_device(pos)
    pos = pos.detach().float().contiguous()
    cache = cache if cache is not None else {}
    if "schnet" not in cache:  # positions are static during pretraining: the radius graph is built once per batch
        csr = radius_graph(pos, model.cutoff, batch, num_graphs, want_edge_index=False)
        es = EdgeSet(csr.rowptr, csr.col, batch, num_graphs)
        zkeys = z.to(torch.int32).reshape(-1, 1).contiguous()
        zidx = bucket_index(z, model.embedding.weight.shape[0])
        E = es.E
        ng = model.num_gaussians
        ea, C = torch.empty(E, ng, dtype=torch.float32, device=dev), torch.empty(max(E, 1), dtype=torch.float32, device=dev)
        check(L.molsde_schnet_edge_feat(ptr(pos), ptr(es.src.idx), ptr(es.tgt.idx), E, ptr(model.distance_expansion.offset), ng,
                                        float(model.distance_expansion.coeff), float(model.cutoff), ptr(ea), ptr(C), tp.s),
              "schnet_edge_feat")
        cache["schnet"] = (es, zkeys, zidx, ea, C)
    es, zkeys, zidx, ea, C = cache["schnet"]
    h = tp.embed_sum(P["embedding.weight"], zkeys, zidx)
    ea_v = Var(ea)
    for i in range(model.num_interactions):
        pf = f"interactions.{i}."
        f1 = tp.linear(ea_v, P[pf + "mlp.0.weight"], P[pf + "mlp.0.bias"], act="ssp")
        f2 = tp.linear(f1, P[pf + "mlp.2.weight"], P[pf + "mlp.2.bias"])
        Wf = tp.rowscale(f2, C)
        x = tp.linear(h, P[pf + "conv.lin1.weight"], None)
        agg = tp.edge_mul_reduce(x, Wf, es.rowptr, es.src, es.tgt)
        t = tp.linear(agg, P[pf + "conv.lin2.weight"], P[pf + "conv.lin2.bias"], act="ssp")
        u = tp.linear(t, P[pf + "lin.weight"], P[pf + "lin.bias"])
        h = tp.add(h, u)
    h = tp.linear(h, P["lin1.weight"], P["lin1.bias"], act="ssp")
    return tp.linear(h, P["lin2.weight"], P["lin2.bias"])


# ======================================================================================================
# dual_CL, EBM_node_dot_prod (examples/util.py:52-79)
# ======================================================================================================
def tape_dual_cl(tp: Tape, X: Var, Y: Var, T: float, neg_index_1: Optional[torch.Tensor] = None,
                 neg_index_2: Optional[torch.Tensor] = None, coef: float = 1.0):
    """(loss [1], loss_acc pairs) of dual_CL(X, Y); backward seeds coef * d loss."""
    L, dev, s = tp.L, tp.dev, tp.s
    N, D = X.data.shape
    outs = []
    dX, dY = tp.empty(N, D), tp.empty(N, D)
    saved = []
    for k, (A, Bv, neg) in enumerate(((X, Y, neg_index_1), (Y, X, neg_index_2))):
        neg = torch.randperm(N) if neg is None else neg   # CPU generator in the reference (util.py:55)
        perm = neg.to(dev).long().contiguous()
        inv = torch.empty_like(perm)
        inv[perm] = torch.arange(N, device=dev)
        pp, pn, out = tp.empty(N), tp.empty(N), tp.empty(2)
        ws = tp.empty(4 * 592)
        tp._call(L.molsde_ebm_node_dot, ptr(A.data), ptr(Bv.data), ptr(perm), N, D, float(T), ptr(pp), ptr(pn), ptr(out), ptr(ws),
                 ws.numel(), s, what="ebm_node_dot")
        outs.append(out)
        saved.append((A, Bv, perm, inv, pp, pn))
    loss = tp.empty(1)
    tp.ew(0, outs[0][:1], outs[1][:1], None, 1.0, loss)
    tp.ew(0, loss, None, None, 0.5, loss)

    def bwd():
        (A, Bv, perm, inv, pp, pn) = saved[0]
        tp._call(L.molsde_ebm_node_dot_bwd, ptr(A.data), ptr(Bv.data), ptr(perm), ptr(inv), ptr(pp), ptr(pn), N, D, float(T),
                 0.5 * coef, 0, ptr(dX), ptr(dY), s, what="ebm_node_dot_bwd")
        (A, Bv, perm, inv, pp, pn) = saved[1]   # roles swapped: A = Y, B = X
        tp._call(L.molsde_ebm_node_dot_bwd, ptr(A.data), ptr(Bv.data), ptr(perm), ptr(inv), ptr(pp), ptr(pn), N, D, float(T),
                 0.5 * coef, 1, ptr(dY), ptr(dX), s, what="ebm_node_dot_bwd")
        tp.accum(X, dX)
        tp.accum(Y, dY)
    tp.ops.append(bwd)
    return loss, outs
