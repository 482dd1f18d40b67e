"""One MoleculeSDE pretraining iteration (`examples/pretrain_MoleculeSDE.py:105-152,331-337`) on our kernels:
GIN -> SchNet -> dual_CL -> SDEModel2Dto3D_02 -> SDEModel3Dto2D_node_adj_dense -> backward -> (all-reduce) -> Adam.

`ParamStore` flattens the trainable parameters of the four modules into ONE fp32 buffer (the modules' `.data` become
views, so `state_dict()` keeps the reference keys) with a matching flat gradient buffer: the unit of the data-parallel
NCCL all-reduce and of the single-launch Adam step.  The per-model `tape_*` functions build the forward on a `Tape`.
"""
from __future__ import annotations

import ctypes
import re
from typing import Dict, Optional, Sequence

import torch
from torch import nn

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr
from .graph import csr_by_target
from .tape import Index, Tape, Var, _p

import os as _os
_GIN_EXACT = _os.environ.get("MOLSDE_GIN_TC") != "1"   # experiment switch: GIN linears on the tensor cores in training too
_NO_MLP3 = _os.environ.get("MOLSDE_NO_MLP3") == "1"   # A/B switch: layer-granular pair MLPs instead of the whole-chain kernels
_SINGLE_STREAM = _os.environ.get("MOLSDE_SINGLE_STREAM") == "1"   # A/B switch: issue the whole iteration on one stream
# parameter gradients on side streams (Tape.wgrad): "capture" = only while a CUDA graph is being captured (the eager step is
# bound by host launch time, where the extra event calls cost more than the overlap returns), "1" always, "0" never
_WGRAD_STREAMS = _os.environ.get("MOLSDE_WGRAD_STREAMS", "capture")
# side streams per tape for the parameter-gradient leaves (consecutive leaves rotate over them): 8.21 -> 8.04 ms at 2 once the GEMMs
# stopped filling the GPU by themselves (tc_gemm split-K policy); 3, 4, 6 give the same
_WGRAD_FANOUT = max(1, int(_os.environ.get("MOLSDE_WGRAD_FANOUT", "2")))
_STREAM_PRIORITY = _os.environ.get("MOLSDE_STREAM_PRIORITY", "0") == "1"   # A/B switch: high-priority branch streams (round 2, after the chain fusions: 9.10 ms with, 8.79 without)


_CH = re.compile(r"^edge_score_network\.layers\.(\d+)\.attn\.(\d+)\.(func_q|func_k)\.layers\.(\d)\.(weight|bias)$")
_CV = re.compile(r"^edge_score_network\.layers\.(\d+)\.attn\.(\d+)\.func_v\.(weight|bias)$")
_QKVS = ("lin_query", "lin_key", "lin_value", "lin_skip")
_CG = re.compile(r"^score_network\.gnn_layers\.(\d+)\.(\d+)\.MHA\.(lin_query|lin_key|lin_value|lin_skip)\.(weight|bias)$")
_CF = re.compile(r"^interactions\.(\d+)\.mlp\.(0|2)\.(weight|bias)$")


def _layout_key(pname: str):
    """Placement of a parameter in the flat buffer.  The per-channel MLPs of EdgeNetwork_dense (`attn.{c}.func_q/func_k/func_v`,
    edge_network_dense.py:33-53) are laid out channel after channel per (layer, tensor kind), so that the stack of all
    channels is ONE contiguous [G, ...] array: the batched GEMM kernels then run all channels of a layer in one launch and
    their weight gradients land directly in the flat gradient buffer.  Everything else keeps registration order."""
    m = _CH.match(pname)
    if m:
        l, c, qk, li, kind = int(m.group(1)), int(m.group(2)), m.group(3), int(m.group(4)), m.group(5)
        return (0, l, li * 2 + (kind == "bias"), qk == "func_k", c)
    m = _CV.match(pname)
    if m:
        return (0, int(m.group(1)), 4 + (m.group(3) == "bias"), 0, int(m.group(2)))
    m = _CG.match(pname)   # TransformerConv projections of one GATLayer: q | k | v | skip = ONE [128, 32] weight (one GEMM per layer)
    if m:
        return (0, 2 * int(m.group(1)) + int(m.group(2)), m.group(4) == "bias", _QKVS.index(m.group(3)), 0)
    m = _CF.match(pname)   # SchNet filter networks: the same layer of all interactions = one [G*128, 51] / [G, 128, 128] stack
    if m:
        return (0, 0, 2 * (m.group(2) == "2") + (m.group(3) == "bias"), 0, int(m.group(1)))
    return (1,)


def _stacked(P: Dict[str, "Var"], names, shape) -> Optional["Var"]:
    """Zero-copy [G, ...] view over parameters that are adjacent in a ParamStore (data and gradient), else None.
    The views of a ParamStore never move, so the result is cached in `P` itself."""
    cache = P.get("__stacked__")
    if cache is None:
        cache = P["__stacked__"] = {}
    key = (names[0], shape)
    if key not in cache:
        cache[key] = _stacked_uncached(P, names, shape)
    return cache[key]


def _stacked_uncached(P: Dict[str, "Var"], names, shape) -> Optional["Var"]:
    vs = [P[n] for n in names]
    if any(v.grad is None for v in vs):
        return None
    for a, b in zip(vs[:-1], vs[1:]):
        if a.data.data_ptr() + a.data.numel() * 4 != b.data.data_ptr() or a.grad.data_ptr() + a.grad.numel() * 4 != b.grad.data_ptr():
            return None
    strides = []
    acc = 1
    for d in reversed(shape):
        strides.insert(0, acc)
        acc *= d
    return Var(torch.as_strided(vs[0].data, shape, strides), True, torch.as_strided(vs[0].grad, shape, strides))


class ParamStore:
    def __init__(self, modules: Dict[str, nn.Module], device: torch.device):
        self.modules, self.dev = modules, device
        self.index: Dict[str, Dict[str, tuple]] = {}
        total = 0
        for mname, m in modules.items():
            self.index[mname] = {}
            named = [(pn, p) for pn, p in m.named_parameters() if p.requires_grad]
            named.sort(key=lambda kv: _layout_key(kv[0]))  # stable: only the per-channel dense layers move (see _layout_key)
            for pname, p in named:
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
        """name -> Var(parameter view, gradient view); the views never change, so the dict is built once."""
        cache = self.__dict__.setdefault("_vars", {})
        if mname not in cache:
            cache[mname] = {pname: Var(self.flat[off:off + n].view(shape), True, self.grad[off:off + n].view(shape))
                            for pname, (off, n, shape) in self.index[mname].items()}
        return cache[mname]

    def grad_view(self, mname: str, pname: str) -> torch.Tensor:
        off, n, shape = self.index[mname][pname]
        return self.grad[off:off + n].view(shape)

    def zero_grad(self) -> None:
        self.grad.zero_()  # cudaMemsetAsync

    def all_reduce(self, modules: Optional[Sequence[str]] = None, async_op: bool = False):
        """Data-parallel gradient exchange: NCCL all-reduce (sum) over the flat buffer, or over the contiguous slice that holds the
        gradients of `modules` (a bucket).  Returns 1/world for Adam -- with `async_op`, (1/world, work handle): the collective is
        ordered after the current stream's work and runs on NCCL's stream, `work.wait()` makes the current stream wait for it."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            return (1.0, None) if async_op else 1.0
        buf = self.grad
        if modules is not None:
            spans = sorted(self.ranges[m] for m in modules if self.ranges[m][1] > self.ranges[m][0])
            assert all(spans[i][1] == spans[i + 1][0] for i in range(len(spans) - 1)), "bucket modules must be adjacent in the flat buffer"
            buf = self.grad[spans[0][0]:spans[-1][1]]
        work = dist.all_reduce(buf, op=dist.ReduceOp.SUM, async_op=async_op)
        scale = 1.0 / dist.get_world_size()
        return (scale, work) if async_op else scale

    def adam_step(self, lr: float = 1e-4, lr_scale: Optional[Dict[str, float]] = None, betas=(0.9, 0.999), eps: float = 1e-8,
                  weight_decay: float = 0.0, grad_scale: float = 1.0, skip: Sequence[str] = ()) -> None:
        """torch.optim.Adam over the four parameter groups (`pretrain_MoleculeSDE.py:331-337`); groups with the same lr
        share one launch.  `skip`: modules that received no gradient this run (a loss term with coefficient 0 is never evaluated,
        `pretrain_MoleculeSDE.py:136-153`, so their `.grad` stays None and torch's Adam leaves them untouched -- no weight
        decay, no moment update)."""
        self.step_count += 1
        s = torch.cuda.current_stream(self.dev).cuda_stream
        scales = {m: (lr_scale or {}).get(m, 1.0) for m in self.modules if m not in skip}
        if not scales:
            return
        if len(set(scales.values())) == 1 and not skip:
            spans = [(0, self.numel, lr * next(iter(scales.values())))]
        else:
            spans = [(a, b, lr * scales[m]) for m, (a, b) in self.ranges.items() if b > a and m in scales]
        for a, b, l in spans:
            check(lib().molsde_adam_step(self.flat[a:b].data_ptr(), self.grad[a:b].data_ptr(), self.exp_avg[a:b].data_ptr(),
                                         self.exp_avg_sq[a:b].data_ptr(), b - a, l, betas[0], betas[1], eps, weight_decay,
                                         self.step_count, grad_scale, s), "adam_step")
        _abi.touch_params()   # raw-pointer update: torch's _version does not move, the packed-weight caches key on this epoch


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
        by_src = csr_by_target(flipped, batch, num_graphs, num_nodes=N)
        self.src = Index(src, by_src.rowptr, by_src.perm, N)


# ======================================================================================================
# SDEModel2Dto3D_02.forward (SDE_model_2D_to_3D.py:306-391)
# ======================================================================================================
def _coef(c, i: int = 0) -> float:
    """Loss weight seeding a backward sweep: a float fixed at record time, or a mutable list filled in right before
    `Tape.backward()` (the autograd bridge passes torch's grad_output this way; two entries = (loss_x, loss_adj))."""
    if isinstance(c, (list, tuple)):
        return float(c[min(i, len(c) - 1)])
    return float(c)


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
    U = tp.linear(h2d, _slice_cols(W0, 0, F), b0, exact=True)      # feeds BatchNorm + ReLU
    V = tp.linear(h2d, _slice_cols(W0, F, 2 * F), None, exact=True)
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
    has_dist = getattr(model, "has_distance_branch", True)
    w_dist = model.dist_gaussian_fourier.W.data if has_dist else torch.zeros(32, dtype=torch.float32, device=dev)
    tp._call(L.molsde_sde2d3d_edge_geom, ptr(pos_p), ptr(es.src.idx), ptr(es.tgt.idx), E, ptr(w_dist),
             ptr(model.coff_gaussian_fourier.W.data), ptr(gfd), ptr(gfi), ptr(gfj), ptr(emb.data), ptr(basis), s, what="edge_geom")
    inv3d = tp.linear(Var(gfd), P["input_mlp.layers.0.weight"], P["input_mlp.layers.0.bias"]) if has_dist else None
    tp.linear(Var(gfi), P["coff_mlp.weight"], P["coff_mlp.bias"], into=emb, col0=2)
    tp.linear(Var(gfj), P["coff_mlp.weight"], P["coff_mlp.bias"], into=emb, col0=2 + 32)
    hid = tp.linear(emb, P["project.layers.0.weight"], P["project.layers.0.bias"], act="silu")
    frame = tp.linear(hid, P["project.layers.1.weight"], P["project.layers.1.bias"])
    edge_attr = tp.mul(inv3d, e2d, frame) if has_dist else tp.add(e2d, frame)   # :372 (_02) / :181 (SDEModel2Dto3D_01)
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
        tp._call(L.molsde_dsm_pos_loss_bwd, ptr(grad), ptr(noise), ptr(w), ptr(prep.node_ptr), ptr(node2graph), N, B, _coef(coef),
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
    Wq = _stacked(P, [pf + f"MHA.{n}.weight" for n in _QKVS], (128, 32))
    bq = _stacked(P, [pf + f"MHA.{n}.bias" for n in _QKVS], (128,))
    if Wq is not None and bq is not None:   # q | k | v | skip adjacent in the flat buffer (`_layout_key`): ONE [N,32] x [128,32]^T GEMM
        qkvs = tp.linear(x, Wq, bq)
    else:
        qkvs = Var(tp.empty(N, 128), False)
        for j, n in enumerate(_QKVS):
            tp.linear(x, P[pf + f"MHA.{n}.weight"], P[pf + f"MHA.{n}.bias"], into=qkvs, col0=32 * j)
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


def prepare_gin(cache: dict, x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor, batch: Optional[torch.Tensor],
                num_graphs: int, train: bool) -> None:
    """Per-batch index structures of the GIN tape (bond CSR by target, its by-source inverse, feature keys and their bucket
    indices), on the current stream; idempotent.  Static for a batch, so an input pipeline can build them ahead of the step."""
    if "gin" in cache:
        return
    from .gnn import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS
    from .tape import bucket_index
    csr = csr_by_target(edge_index, batch, num_graphs, num_nodes=x.size(0))   # batch None: segments recovered from the edge list
    es = EdgeSet(csr.rowptr, csr.col, batch, num_graphs) if train else None
    akeys = _keys(x, ATOM_FEATURE_DIMS)
    ekeys = _keys(edge_attr[csr.perm.long()], BOND_FEATURE_DIMS)   # CSR edge order
    aidx = bucket_index(akeys.reshape(-1).long(), sum(ATOM_FEATURE_DIMS)) if train else None
    eidx = bucket_index(ekeys.reshape(-1).long(), sum(BOND_FEATURE_DIMS)) if train else None
    cache["gin"] = (csr, es, akeys, ekeys, aidx, eidx)


def tape_gin(tp: Tape, model, P: Dict[str, Var], x: torch.Tensor, edge_index: torch.Tensor, edge_attr: torch.Tensor,
             cache: Optional[dict], batch: Optional[torch.Tensor] = None, num_graphs: int = 1) -> Var:
    """GNN.forward, GIN / JK=last / dropout 0.  `cache` (a dict living on the batch) keeps the index structures;
    `batch` (node -> graph, ascending) only speeds up the CSR build."""
    from .gnn import ATOM_FEATURE_DIMS, BOND_FEATURE_DIMS
    from .tape import bucket_index
    dev = x.device
    require_device(x)
    train = any(v.needs for v in P.values() if isinstance(v, Var))
    cache = cache if cache is not None else {}
    prepare_gin(cache, x, edge_index, edge_attr, batch, num_graphs, train)
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
        z = tp.linear(pre, P[pf + "mlp.0.weight"], P[pf + "mlp.0.bias"], exact=model.training and _GIN_EXACT)   # feeds BatchNorm + ReLU (train: sign decisions feed gradients)
        last = l == model.num_layer - 1
        if model.training:
            z = tp.batchnorm(z, P[pf + "mlp.1.weight"], P[pf + "mlp.1.bias"], bn1.running_mean, bn1.running_var, bn1.eps,
                             bn1.momentum, relu=True)
            bn1.num_batches_tracked += 1
        else:
            z = tp.batchnorm_eval(z, P[pf + "mlp.1.weight"], P[pf + "mlp.1.bias"], bn1.running_mean, bn1.running_var, bn1.eps, relu=True)
        z = tp.linear(z, P[pf + "mlp.3.weight"], P[pf + "mlp.3.bias"], exact=model.training and _GIN_EXACT)
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
def prepare_schnet(cache: dict, model, z: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor, num_graphs: int) -> None:
    """Radius graph (CSR + by-source inverse), atomic-number bucket index, GaussianSmearing features and cosine cutoff of the
    edges, on the current stream; idempotent.  Depends on the positions only, i.e. static for a pretraining batch."""
    if "schnet" in cache:
        return
    from .graph import radius_graph
    from .tape import bucket_index
    L, dev = lib(), pos.device
    pos = pos.detach().float().contiguous()
    csr = radius_graph(pos, model.cutoff, batch, num_graphs, want_edge_index=False)
    es = EdgeSet(csr.rowptr, csr.col, batch, num_graphs)
    zkeys = z.to(torch.int32).reshape(-1, 1).contiguous()
    zidx = bucket_index(z, model.embedding.weight.shape[0])
    E = es.E
    ng = model.num_gaussians
    ea, C = torch.empty(E, ng, dtype=torch.float32, device=dev), torch.empty(max(E, 1), dtype=torch.float32, device=dev)
    check(L.molsde_schnet_edge_feat(ptr(pos), ptr(es.src.idx), ptr(es.tgt.idx), E, ptr(model.distance_expansion.offset), ng,
                                    float(model.distance_expansion.coeff), float(model.cutoff), ptr(ea), ptr(C), stream_ptr(pos)),
          "schnet_edge_feat")
    cache["schnet"] = (es, zkeys, zidx, ea, C)


def tape_schnet(tp: Tape, model, P: Dict[str, Var], z: torch.Tensor, pos: torch.Tensor, batch: torch.Tensor, num_graphs: int,
                cache: Optional[dict], pos_var: Optional[Var] = None) -> Var:
    """Node representation h [N, hidden] of SchNet.forward(return_latent=True) with every intermediate kept.
    `pos_var` (a Var over `pos` with needs=True): also differentiate with respect to the positions -- GaussianSmearing and the
    cosine cutoff are functions of the edge length -- which is how the reference obtains forces (`finetune_MD17.py:66`)."""
    from .graph import radius_graph
    from .tape import bucket_index
    L, dev = tp.L, pos.device
    require_device(pos)
    pos = pos.detach().float().contiguous()
    need_pos = pos_var is not None and pos_var.needs
    cache = cache if cache is not None else {}
    prepare_schnet(cache, model, z, pos, batch, num_graphs)   # positions are static during pretraining: once per batch
    es, zkeys, zidx, ea, C = cache["schnet"]
    h = tp.embed_sum(P["embedding.weight"], zkeys, zidx)
    ea_v = Var(ea, need_pos)
    C_v = Var(C[:es.E], need_pos)
    if need_pos:
        N, E = es.N, es.E

        def pos_bwd():   # recorded first => runs last, after every interaction added its share to d ea / d C
            if ea_v.grad is None or E == 0:
                return
            g = tp.empty(E, 3)
            tp._call(L.molsde_schnet_edge_feat_bwd, ptr(pos), ptr(es.src.idx), ptr(es.tgt.idx), E, ptr(model.distance_expansion.offset),
                     model.num_gaussians, float(model.distance_expansion.coeff), float(model.cutoff), ptr(ea), ptr(ea_v.grad),
                     _p(C_v.grad), ptr(g), tp.s, what="schnet_edge_feat_bwd")
            dpos, dneg = tp.empty(N, 3), tp.empty(N, 3)
            tp.seg_sum(g, es.src, 3, dpos)                    # edges leaving the atom: + g
            tp.seg_sum(g, es.tgt, 3, dneg)                    # edges arriving at the atom: - g
            tp.ew(0, dpos, dneg, None, -1.0, dpos)
            tp.accum(pos_var, dpos)
        tp.ops.append(pos_bwd)
    # The filter networks W_i(e) = C(d_e) * mlp_i(GaussianSmearing(d_e)) depend on the edges only, not on h: all interactions'
    # filters are evaluated up front as ONE stack [E, G*F] (first layers = one GEMM against the [G*F, 51] weight stack, second
    # layers = one grouped GEMM) when the parameters are adjacent in the flat buffer (`_layout_key`); their backward runs once,
    # after every interaction deposited d W_i in its column block.  Same dot products as the per-interaction form.
    G = model.num_interactions
    Wf_all = None
    if not need_pos and G > 0 and es.E > 0:
        F_, ng = P["interactions.0.mlp.2.weight"].data.shape[0], ea.shape[1]
        names = lambda k: [f"interactions.{i}.mlp.{k}" for i in range(G)]   # noqa: E731
        W0s, b0s = _stacked(P, names("0.weight"), (G * F_, ng)), _stacked(P, names("0.bias"), (G * F_,))
        W2s, b2s = _stacked(P, names("2.weight"), (G, F_, F_)), _stacked(P, names("2.bias"), (G * F_,))
        if None not in (W0s, b0s, W2s, b2s):
            f1 = tp.linear(ea_v, W0s, b0s, act="ssp")
            Wf_all = tp.grouped_linear(f1, W2s, b2s, G)      # UNSCALED: the cosine cutoff C(d_e) is applied inside the CFConv kernels
    for i in range(G):
        pf = f"interactions.{i}."
        x = tp.linear(h, P[pf + "conv.lin1.weight"], None)
        if Wf_all is not None:
            agg = tp.edge_mul_reduce(x, Wf_all, es.rowptr, es.src, es.tgt, wcol0=i * F_, escale=C)
        else:
            f1 = tp.linear(ea_v, P[pf + "mlp.0.weight"], P[pf + "mlp.0.bias"], act="ssp")
            f2 = tp.linear(f1, P[pf + "mlp.2.weight"], P[pf + "mlp.2.bias"])
            Wf = tp.rowscale_var(f2, C_v) if need_pos else tp.rowscale(f2, C)
            agg = tp.edge_mul_reduce(x, Wf, es.rowptr, es.src, es.tgt)
        t = tp.linear(agg, P[pf + "conv.lin2.weight"], P[pf + "conv.lin2.bias"], act="ssp")
        u = tp.linear(t, P[pf + "lin.weight"], P[pf + "lin.bias"])
        h = tp.add(h, u)
    h = tp.linear(h, P["lin1.weight"], P["lin1.bias"], act="ssp")
    return tp.linear(h, P["lin2.weight"], P["lin2.bias"])


def tape_schnet_dual(tp: Tape, model, P: Dict[str, Var], z: torch.Tensor, pos: torch.Tensor, vdir: torch.Tensor, batch: torch.Tensor,
                     num_graphs: int, cache: Optional[dict]):
    """SchNet node representation h [N, hidden] TOGETHER with its directional derivative h_dot = d/d eps h(pos + eps vdir)
    (forward-mode tangent), every op recorded on the tape, so that a reverse sweep seeded on h_dot yields
    d/d theta <vdir, d(.)/d pos> -- the parameter gradient of a force term (`finetune_MD17.py:66-77`).  The tangent flows through
    the same kernels as the primal: Linear without bias, `act_tangent` (act', with act'' in its backward), CFConv applied to
    (x_dot, W) and (x, W_dot); the edge features' tangent (GaussianSmearing, cosine cutoff) does not depend on the parameters and
    is one forward kernel.  Returns (h, h_dot)."""
    L, dev = tp.L, pos.device
    require_device(pos)
    pos = pos.detach().float().contiguous()
    vdir = vdir.detach().float().contiguous()
    cache = cache if cache is not None else {}
    prepare_schnet(cache, model, z, pos, batch, num_graphs)
    es, zkeys, zidx, ea, C = cache["schnet"]
    E, ng = es.E, model.num_gaussians
    ea_dot, C_dot = tp.empty(E, ng), tp.empty(max(E, 1))
    tp._call(L.molsde_schnet_edge_feat_tangent, ptr(pos), ptr(vdir), ptr(es.src.idx), ptr(es.tgt.idx), E,
             ptr(model.distance_expansion.offset), ng, float(model.distance_expansion.coeff), float(model.cutoff), ptr(ea), ptr(ea_dot),
             ptr(C_dot), tp.s, what="schnet_edge_feat_tangent")
    ea_v, ead_v = Var(ea, False), Var(ea_dot, False)
    Cc, Cd = C[:E], C_dot[:E]
    h = tp.embed_sum(P["embedding.weight"], zkeys, zidx)
    hd = None                                        # the embedding does not depend on the positions
    for i in range(model.num_interactions):
        pf = f"interactions.{i}."
        W0, b0, W2, b2 = P[pf + "mlp.0.weight"], P[pf + "mlp.0.bias"], P[pf + "mlp.2.weight"], P[pf + "mlp.2.bias"]
        p1 = tp.linear(ea_v, W0, b0)
        f2 = tp.linear(tp.act(p1, "ssp"), W2, b2)
        Wf = tp.rowscale(f2, Cc)
        f2d = tp.linear(tp.act_tangent(p1, tp.linear(ead_v, W0, None), "ssp"), W2, None)
        Wfd = tp.add(tp.rowscale(f2d, Cc), tp.rowscale(f2, Cd))
        W1 = P[pf + "conv.lin1.weight"]
        x = tp.linear(h, W1, None)
        agg = tp.edge_mul_reduce(x, Wf, es.rowptr, es.src, es.tgt)
        aggd = tp.edge_mul_reduce(x, Wfd, es.rowptr, es.src, es.tgt)
        if hd is not None:
            aggd = tp.add(aggd, tp.edge_mul_reduce(tp.linear(hd, W1, None), Wf, es.rowptr, es.src, es.tgt))
        Wc, bc = P[pf + "conv.lin2.weight"], P[pf + "conv.lin2.bias"]
        p2 = tp.linear(agg, Wc, bc)
        td = tp.act_tangent(p2, tp.linear(aggd, Wc, None), "ssp")
        Wl, bl = P[pf + "lin.weight"], P[pf + "lin.bias"]
        h = tp.add(h, tp.linear(tp.act(p2, "ssp"), Wl, bl))
        ud = tp.linear(td, Wl, None)
        hd = ud if hd is None else tp.add(hd, ud)
    p3 = tp.linear(h, P["lin1.weight"], P["lin1.bias"])
    out = tp.linear(tp.act(p3, "ssp"), P["lin2.weight"], P["lin2.bias"])
    outd = tp.linear(tp.act_tangent(p3, tp.linear(hd, P["lin1.weight"], None), "ssp"), P["lin2.weight"], None)
    return out, outd


# ======================================================================================================
# dual_CL, EBM_node_dot_prod (examples/util.py:52-79)
# ======================================================================================================
def tape_cl(tp: Tape, X: Var, Y: Var, T: float, neg_index: Optional[torch.Tensor] = None, coef=1.0):
    """do_CL(X, Y) with metric EBM_node_dot_prod (`examples/util.py:52-68`): returns the [loss, acc] tensor; the backward seeds
    coef * d loss into X.grad and Y.grad."""
    L, dev, s = tp.L, tp.dev, tp.s
    N, D = X.data.shape
    # the reference draws the permutation on the CPU generator (util.py:55); here on the device, so that the step stays capturable
    neg = torch.randperm(N, device=dev) if neg_index is None else neg_index
    perm = neg.to(dev).long().contiguous()
    inv = torch.empty_like(perm)
    inv[perm] = torch.arange(N, device=dev)
    pp, pn, out = tp.empty(N), tp.empty(N), tp.empty(2)
    ws = tp.empty(4 * 592)
    tp._call(L.molsde_ebm_node_dot, ptr(X.data), ptr(Y.data), ptr(perm), N, D, float(T), ptr(pp), ptr(pn), ptr(out), ptr(ws),
             ws.numel(), s, what="ebm_node_dot")

    def bwd():
        dX, dY = tp.empty(N, D), tp.empty(N, D)
        tp._call(L.molsde_ebm_node_dot_bwd, ptr(X.data), ptr(Y.data), ptr(perm), ptr(inv), ptr(pp), ptr(pn), N, D, float(T),
                 _coef(coef), 0, ptr(dX), ptr(dY), s, what="ebm_node_dot_bwd")
        tp.accum(X, dX)
        tp.accum(Y, dY)
    tp.ops.append(bwd)
    return out


def tape_infonce(tp: Tape, X: Var, Y: Var, T: float, coef=1.0):
    """do_CL(X, Y) with metric InfoNCE_dot_prod (`examples/util.py:23-32`): logits = X Y^T / T on the tensor cores (tcgen05
    3xTF32), CrossEntropy against the diagonal.  Returns the [loss, acc] tensor; backward: d logits = (softmax - I) coef / B, then
    dX = d logits . Y / T and dY = d logits^T . X / T as two more tensor-core GEMMs."""
    L, s = tp.L, tp.s
    B, D = X.data.shape
    inv_t = torch.full((B,), 1.0 / float(T), dtype=torch.float32, device=tp.dev)
    logits = tp.empty(B, B)
    tp._call(L.molsde_tc_gemm, B, B, D, ptr(X.data), D, 1, ptr(Y.data), D, 1, None, 0, ptr(inv_t), None, 0, ptr(logits), B, 0, None, 0,
             None, s, what="tc_gemm")
    loss_row, correct_row, out = tp.empty(B), tp.empty(B), tp.empty(2)
    tp._call(L.molsde_infonce_rows, ptr(logits), B, B, 0.0, 0, ptr(loss_row), ptr(correct_row), s, what="infonce_rows")
    tp._call(L.molsde_mean, ptr(loss_row), B, out.data_ptr(), s, what="mean")
    tp._call(L.molsde_mean, ptr(correct_row), B, out.data_ptr() + 4, s, what="mean")

    def bwd():
        tp._call(L.molsde_infonce_rows, ptr(logits), B, B, _coef(coef) / (B * float(T)), 1, None, None, s, what="infonce_rows")
        if X.needs:
            dX = tp.empty(B, D)
            tp.gemm(0, 0, B, D, B, logits, B, Y.data, D, dX, D)          # d logits . Y
            tp.accum(X, dX)
        if Y.needs:
            dY = tp.empty(B, D)
            tp.gemm(1, 0, B, D, B, logits, B, X.data, D, dY, D)          # d logits^T . X
            tp.accum(Y, dY)
    tp.ops.append(bwd)
    return out


def tape_dual_cl(tp: Tape, X: Var, Y: Var, T: float, neg_index_1: Optional[torch.Tensor] = None,
                 neg_index_2: Optional[torch.Tensor] = None, coef=1.0):
    """(loss [1], [loss, acc] pairs) of dual_CL(X, Y) = (do_CL(X,Y) + do_CL(Y,X)) / 2 (`util.py:76-79`)."""
    half = [0.5 * _coef(coef)] if not isinstance(coef, list) else _Half(coef)
    o1 = tape_cl(tp, X, Y, T, neg_index_1, half)
    o2 = tape_cl(tp, Y, X, T, neg_index_2, half)
    loss = tp.empty(1)
    tp.ew(0, o1[:1], o2[:1], None, 1.0, loss)
    tp.ew(0, loss, None, None, 0.5, loss)
    return loss, [o1, o2]


class _Half(list):
    """view of a mutable coefficient list scaled by 1/2 (read at backward time)"""

    def __init__(self, src):
        super().__init__([0.0])
        self.src = src

    def __getitem__(self, i):
        return 0.5 * float(self.src[0])


# ======================================================================================================
# SDEModel3Dto2D_node_adj_dense.forward (SDE_model_3D_to_2D_node_adj_dense.py:101-179), train=True, reduce_mean=True
# ======================================================================================================
def _dense_gcn(tp: Tape, adjc: Var, c: int, C: int, xw: Var, xw_col0: int, bias: Var, Fo: int, out: Var, out_off: int, act: str,
               B: int, Nm: int) -> None:
    """One channel of NodeNetwork_dense (dense GCN) writing act(Ahat . xw + bias) into out[:, out_off:out_off+Fo]."""
    L, s = tp.L, tp.s
    a = adjc.data
    sb = a.stride(0)
    a_ptr = a.data_ptr() + 4 * (c * Nm * Nm if a.dim() == 4 else 0)
    code = {"none": 0, "tanh": 4}[act]
    tp._call(L.molsde_dense_gcn, a_ptr, sb, 0, B, 1, Nm, xw.data.data_ptr() + 4 * xw_col0, xw.data.stride(0), ptr(bias.data), Fo,
             out.data.data_ptr(), out.data.stride(0), out_off, code, s, what="dense_gcn")
    out.needs = True

    def bwd():
        dout = tp.grad_of(out)
        dpre = tp.empty(B * Nm, Fo)
        dxw = tp.grad_of(xw)
        want_dadj = adjc.needs
        da_ptr = None
        if want_dadj:
            da = tp.grad_of(adjc)
            da_ptr = da.data_ptr() + 4 * c * Nm * Nm
        tp._call(L.molsde_dense_gcn_bwd, a_ptr, sb, 0, B, 1, Nm, xw.data.data_ptr() + 4 * xw_col0, xw.data.stride(0), Fo,
                 out.data.data_ptr(), dout.data_ptr(), out.data.stride(0), out_off, code, ptr(dpre),
                 dxw.data_ptr() + 4 * xw_col0, dxw.stride(0), da_ptr, C * Nm * Nm, 1, s, what="dense_gcn_bwd")
        if bias.needs:
            with tp.wgrad(dpre, key=bias.grad):
                tp.colsum(dpre, B * Nm, Fo, Fo, bias.grad, accumulate=True)
    tp.ops.append(bwd)


def _mlp(tp: Tape, P: Dict[str, Var], prefix: str, x: Var, n_layers: int, act: str, last_rowscale=None, last_act: str = "none") -> Var:
    if n_layers == 3 and last_rowscale is None and last_act == "none" and not _NO_MLP3:
        Ws = [P[f"{prefix}.layers.{i}.weight"] for i in range(3)]
        bs = [P.get(f"{prefix}.layers.{i}.bias") for i in range(3)]
        if tp.mlp3_supported(x, Ws, bs, act):     # narrow pair MLPs: one forward + one input-gradient launch (csrc/train_mlp.cu)
            return tp.mlp3(x, Ws, bs, act)
    for i in range(n_layers):
        last = i == n_layers - 1
        x = tp.linear(x, P[f"{prefix}.layers.{i}.weight"], P[f"{prefix}.layers.{i}.bias"], act=last_act if last else act,
                      rowscale=last_rowscale if last else None)
    return x


def _edge_network(tp: Tape, P: Dict[str, Var], pf: str, lyr, x: Var, adjc: Var, flags: torch.Tensor, allc: Var, all_off: int,
                  B: int, Nm: int, last: bool):
    """EdgeNetwork_dense.forward (edge_network_dense.py:105-128) -> (x_out, adjc_next)."""
    L, s = tp.L, tp.s
    C, W, Fo, Co = lyr.in_ch, 2 * lyr.attn_dim, lyr.conv_out, lyr.out_ch
    ds = lyr.attn_dim // lyr.num_heads
    rows = B * Nm
    Fin = x.data.shape[1]
    qk_names = lambda li, kind: ([f"{pf}attn.{c}.func_q.layers.{li}.{kind}" for c in range(C)] +
                                 [f"{pf}attn.{c}.func_k.layers.{li}.{kind}" for c in range(C)])
    W0p = _stacked(P, qk_names(0, "weight"), (2 * C * W, Fin))
    b0p = _stacked(P, qk_names(0, "bias"), (2 * C * W,))
    W1p = _stacked(P, qk_names(1, "weight"), (2 * C, W, W))
    b1p = _stacked(P, qk_names(1, "bias"), (2 * C * W,))
    Wvp = _stacked(P, [f"{pf}attn.{c}.func_v.weight" for c in range(C)], (C, Fin, Fo))
    bvp = _stacked(P, [f"{pf}attn.{c}.func_v.bias" for c in range(C)], (C * Fo,))
    if all(v is not None for v in (W0p, b0p, W1p, b1p, Wvp, bvp)):
        # all channels of the layer per launch (parameters are stacked in the ParamStore, see _layout_key)
        h1 = tp.linear(x, W0p, b0p, act="tanh")                     # func_q/func_k layer 0 of every channel: one GEMM
        qk = tp.grouped_linear(h1, W1p, b1p, 2 * C)                 # layer 1: 2C independent 32x32 GEMMs, one launch
        xw = tp.grouped_matmul_shared(x, Wvp, C)                    # x @ func_v.weight for every channel
        V = Var(tp.empty(rows, C * Fo), False)
        _dense_gcn_all(tp, adjc, C, xw, bvp, Fo, V, B, Nm)
        return _edge_network_tail(tp, P, pf, lyr, adjc, flags, allc, all_off, B, Nm, last, qk, V, C, W, ds, Fo, Co)
    # fallback: parameters not stacked (not owned by a ParamStore) -> one launch per channel
    # func_q / func_k (MLP in -> 2a -> 2a, tanh between) for every channel, written side by side: q_0..q_{C-1}, k_0..k_{C-1}
    h1pre = Var(tp.empty(rows, 2 * C * W), False)
    for c in range(C):
        tp.linear(x, P[f"{pf}attn.{c}.func_q.layers.0.weight"], P[f"{pf}attn.{c}.func_q.layers.0.bias"], into=h1pre, col0=c * W)
        tp.linear(x, P[f"{pf}attn.{c}.func_k.layers.0.weight"], P[f"{pf}attn.{c}.func_k.layers.0.bias"], into=h1pre, col0=(C + c) * W)
    h1 = tp.act(h1pre, "tanh")
    qk = Var(tp.empty(rows, 2 * C * W), False)
    for c in range(C):
        tp.linear(h1, P[f"{pf}attn.{c}.func_q.layers.1.weight"], P[f"{pf}attn.{c}.func_q.layers.1.bias"], into=qk, col0=c * W,
                  x_cols=(c * W, (c + 1) * W))
        tp.linear(h1, P[f"{pf}attn.{c}.func_k.layers.1.weight"], P[f"{pf}attn.{c}.func_k.layers.1.bias"], into=qk, col0=(C + c) * W,
                  x_cols=((C + c) * W, (C + c + 1) * W))
    # func_v: dense GCN per channel
    xw = Var(tp.empty(rows, C * Fo), False)
    V = Var(tp.empty(rows, C * Fo), False)
    for c in range(C):
        tp.matmul(x, P[f"{pf}attn.{c}.func_v.weight"], into=xw, col0=c * Fo)
    for c in range(C):
        _dense_gcn(tp, adjc, c, C, xw, c * Fo, P[f"{pf}attn.{c}.func_v.bias"], Fo, V, c * Fo, "none", B, Nm)
    return _edge_network_tail(tp, P, pf, lyr, adjc, flags, allc, all_off, B, Nm, last, qk, V, C, W, ds, Fo, Co)


def _dense_gcn_all(tp: Tape, adjc: Var, C: int, xw: Var, bias: Var, Fo: int, out: Var, B: int, Nm: int) -> None:
    """All C channels of func_v (dense GCN, no activation) in one launch; bias [C*Fo] stacked."""
    L, s = tp.L, tp.s
    a = adjc.data
    sb, sc = a.stride(0), a.stride(1)
    tp._call(L.molsde_dense_gcn, ptr(a), sb, sc, B, C, Nm, ptr(xw.data), xw.data.stride(0), ptr(bias.data), Fo, ptr(out.data),
             out.data.stride(0), 0, 0, s, what="dense_gcn")
    out.needs = True

    def bwd():
        dout = tp.grad_of(out)
        dpre = tp.empty(B * Nm, C * Fo)
        dxw = tp.empty(B * Nm, C * Fo)
        da = tp.grad_of(adjc) if adjc.needs else None
        tp._call(L.molsde_dense_gcn_bwd, ptr(a), sb, sc, B, C, Nm, ptr(xw.data), xw.data.stride(0), Fo, ptr(out.data), ptr(dout),
                 out.data.stride(0), 0, 0, ptr(dpre), ptr(dxw), C * Fo, _p(da), C * Nm * Nm, 1, s, what="dense_gcn_bwd")
        if bias.needs:
            with tp.wgrad(dpre, key=bias.grad):
                tp.colsum(dpre, B * Nm, C * Fo, C * Fo, bias.grad, accumulate=True)
        tp.accum(xw, dxw)
    tp.ops.append(bwd)


def _edge_network_tail(tp: Tape, P, pf, lyr, adjc, flags, allc, all_off, B, Nm, last, qk, V, C, W, ds, Fo, Co):
    L, s = tp.L, tp.s
    # attention scores + [A, adj] concat (recorded AFTER the GCNs: its backward overwrites d adjc, theirs accumulate)
    pair = Var(tp.empty(B * Nm * Nm, 2 * C), True)
    tp._call(L.molsde_dense_attn, qk.data.data_ptr(), qk.data.data_ptr() + 4 * C * W, qk.data.stride(0), W, ds, ptr(adjc.data), B, C, Nm,
             ptr(pair.data), s, what="dense_attn")

    def attn_bwd():
        if pair.grad is None:
            return
        dqk = tp.grad_of(qk)
        da = tp.grad_of(adjc) if adjc.needs else None
        tp._call(L.molsde_dense_attn_bwd, qk.data.data_ptr(), qk.data.data_ptr() + 4 * C * W, qk.data.stride(0), W, ds, B, C, Nm,
                 ptr(pair.grad), dqk.data_ptr(), dqk.data_ptr() + 4 * C * W, _p(da), s, what="dense_attn_bwd")
    tp.ops.append(attn_bwd)
    x_out = None
    if not last:  # the last layer's node features are never used (invariant_scorenetwork_dense.py:80-83)
        hmc = tp.linear(V, P[pf + "multi_channel.layers.0.weight"], P[pf + "multi_channel.layers.0.bias"], act="elu")
        x_out = tp.linear(hmc, P[pf + "multi_channel.layers.1.weight"], P[pf + "multi_channel.layers.1.bias"], act="tanh",
                          rowscale=flags.reshape(-1))
    m = _mlp(tp, P, pf + "mlp", pair, len(lyr.mlp.layers), "elu")
    adj_next = Var(tp.empty(B, Co, Nm, Nm), not last)
    tp._call(L.molsde_dense_pair_post, ptr(m.data), ptr(flags), B, Nm, Co, ptr(adj_next.data), ptr(allc.data), allc.data.size(-1),
             all_off, s, what="dense_pair_post")

    def post_bwd():
        dm = tp.empty(B * Nm * Nm, Co)
        tp._call(L.molsde_dense_pair_post_bwd, _p(adj_next.grad), ptr(tp.grad_of(allc)), allc.data.size(-1), all_off, ptr(flags), B, Nm,
                 Co, ptr(dm), s, what="dense_pair_post_bwd")
        tp.accum(m, dm)
    tp.ops.append(post_bwd)
    return x_out, adj_next


def tape_3d2d(tp: Tape, model, P: Dict[str, Var], h3d: Var, data, anneal_power: float = 0.0, draws=None, coef: float = 0.5):
    """Records both DSM losses of the dense 3D->2D model; returns (loss_x [1], loss_adj [1]).  The backward seeds
    d(total) = coef * (loss_x + loss_adj)  (pretrain_MoleculeSDE.py:146: (x + adj) * 0.5)."""
    from .sde_3d_to_2d import EPSILON, gen_noise
    L, dev, s = tp.L, tp.dev, tp.s
    adj, rep, zd, flags, Nm = model.dense_inputs(h3d.data, data)
    B, T, K, F = adj.size(0), model.num_diffusion_timesteps, model.num_class_X, model.nfeat
    rows = B * Nm
    N = h3d.data.size(0)
    node_ptr = getattr(data, "_molsde_node_ptr", None)
    if node_ptr is None:
        from .graph import segment_ptr
        node_ptr = data._molsde_node_ptr = segment_ptr(data.batch, B)
    rep_v = Var(rep.view(rows, F), h3d.needs)
    if h3d.needs:
        node2graph = data.batch.to(torch.int32)

        def rep_bwd():
            if rep_v.grad is None:
                return
            g = tp.empty(N, F)
            tp._call(L.molsde_from_dense_batch, ptr(rep_v.grad), F, ptr(node_ptr), ptr(node2graph), N, Nm, F, ptr(g), s,
                     what="from_dense_batch")
            tp.accum(h3d, g)
        tp.ops.append(rep_bwd)
    # ---- perturbation (no gradient) :111-152
    th = torch.randint(0, T, size=(B // 2 + 1,), device=dev) if draws is None else draws[0].to(dev)
    t = torch.cat([th, T - th - 1], dim=0)[:B]
    t = t / T * (1 - EPSILON) + EPSILON
    z_adj = gen_noise(adj, flags, sym=True, raw=None if draws is None else draws[1])
    std_adj = model.sde_adj.marGINal_prob(torch.zeros(B, 1, 1, device=dev), t)[1].float().contiguous()
    coef_adj = model.sde_adj.mean_coeff(t).float().contiguous()
    p_adj = tp.empty(B, Nm, Nm)
    tp._call(L.molsde_dense_perturb_adj, ptr(adj), ptr(z_adj), ptr(flags), ptr(coef_adj), ptr(std_adj), B, Nm, ptr(p_adj), s,
             what="perturb_adj")
    raw_x = torch.randn(B, Nm, K, device=dev) if draws is None else draws[2].to(dev).float().contiguous()
    std_x = model.sde_x.marGINal_prob(torch.zeros(B, 1, 1, device=dev), t)[1].float().contiguous()
    coef_x = model.sde_x.mean_coeff(t).float().contiguous()
    z_x, p_x = tp.empty(B, Nm, K), tp.empty(B, Nm, K)
    tp._call(L.molsde_dense_perturb_onehot, ptr(zd.contiguous()), ptr(raw_x), ptr(flags), ptr(coef_x), ptr(std_x), B, Nm, K, ptr(z_x),
             ptr(p_x), s, what="perturb_onehot")
    scale_adj = (-1.0 / std_adj).contiguous()
    scale_x = (-1.0 / std_x).contiguous()
    # ---- embedding :156
    e3 = tp.linear(rep_v, P["embedding_3D.weight"], P["embedding_3D.bias"])
    ex = tp.linear(Var(p_x.view(rows, K)), P["embedding_X.weight"], P["embedding_X.bias"])
    if getattr(model, "concat_embedding", False):   # `_dense_02` (:333): cat([embedding_3D, embedding_X], -1)
        emb = Var(tp.empty(rows, 2 * F), False)
        tp.copy_cols(e3, emb, 0)
        tp.copy_cols(ex, emb, F)
    else:
        emb = tp.add(e3, ex)

    # ---- EdgeScoreNetwork_dense (invariant_scorenetwork_dense.py:74-93)
    esn = model.edge_score_network
    allc = Var(tp.empty(B * Nm * Nm, esn.fdim), False)
    adjc0 = tp.empty(B, 2, Nm, Nm)
    tp._call(L.molsde_dense_pow2, ptr(p_adj), B, Nm, ptr(adjc0), ptr(allc.data), esn.fdim, 0, s, what="dense_pow2")
    x, adjc, off = emb, Var(adjc0, False), esn.c_init
    for li, lyr in enumerate(esn.layers):
        x, adjc = _edge_network(tp, P, f"edge_score_network.layers.{li}.", lyr, x, adjc, flags, allc, off, B, Nm,
                                last=li == len(esn.layers) - 1)
        off += lyr.out_ch
    allc.needs = True
    raw = _mlp(tp, P, "edge_score_network.final", allc, len(esn.final.layers), "silu")
    score_adj = tp.empty(B, Nm, Nm)
    tp._call(L.molsde_dense_edge_final, ptr(raw.data), ptr(flags), ptr(scale_adj), B, Nm, ptr(score_adj), s, what="dense_edge_final")

    # ---- NodeScoreNetwork_dense (:118-131)
    nsn = model.node_score_network
    xs = Var(tp.empty(rows, nsn.fdim), False)
    tp.copy_cols(emb, xs, 0)
    adj_in = Var(p_adj, False)
    cur_off, off = 0, nsn.nfeat
    for li in range(nsn.depth):
        width = nsn.nfeat if li == 0 else nsn.nhid
        xw = Var(tp.empty(rows, nsn.nhid), False)
        W = P[f"node_score_network.layers.{li}.weight"]
        # x @ weight on the column slice of the concat buffer
        xin = Var(xs.data[:, cur_off:cur_off + width], False)
        tp.gemm(0, 0, rows, nsn.nhid, width, xin.data, xs.data.stride(0), W.data, nsn.nhid, xw.data, nsn.nhid)
        xw.needs = True

        def mm_bwd(xw=xw, W=W, width=width, cur_off=cur_off):
            if xw.grad is None:
                return
            xv = xs.data[:, cur_off:cur_off + width]
            with tp.wgrad(xw.grad, xs.data, key=W.grad):
                tp.gemm(1, 0, width, nsn.nhid, rows, xv, xs.data.stride(0), xw.grad, nsn.nhid, W.grad, nsn.nhid, accumulate=True)
            g = tp.grad_of(xs)[:, cur_off:cur_off + width]
            tp.gemm(0, 1, rows, width, nsn.nhid, xw.grad, nsn.nhid, W.data, nsn.nhid, g, xs.data.stride(0), accumulate=True)
        tp.ops.append(mm_bwd)
        _dense_gcn(tp, adj_in, 0, 1, xw, 0, P[f"node_score_network.layers.{li}.bias"], nsn.nhid, xs, off, "tanh", B, Nm)
        cur_off, off = off, off + nsn.nhid
    rs = (flags * scale_x[:, None]).reshape(-1).contiguous()
    score_x = _mlp(tp, P, "node_score_network.final", xs, len(nsn.final.layers), "silu", last_rowscale=rs)

    # ---- losses :160-179
    wx = None if anneal_power == 0 else (std_x ** anneal_power).contiguous()
    wa = None if anneal_power == 0 else (std_adj ** anneal_power).contiguous()
    lx_g, la_g, loss_x, loss_adj = tp.empty(B), tp.empty(B), tp.empty(1), tp.empty(1)
    tp._call(L.molsde_graph_reduce, ptr(score_x.data), ptr(z_x), _p(wx), B, Nm * K, 1, ptr(lx_g), s, what="graph_reduce")
    tp._call(L.molsde_graph_reduce, ptr(score_adj), ptr(z_adj), _p(wa), B, Nm * Nm, 1, ptr(la_g), s, what="graph_reduce")
    tp._call(L.molsde_mean, ptr(lx_g), B, ptr(loss_x), s, what="mean")
    tp._call(L.molsde_mean, ptr(la_g), B, ptr(loss_adj), s, what="mean")

    def loss_bwd():
        dsx = tp.empty(rows, K)
        tp._call(L.molsde_graph_mse_bwd, ptr(score_x.data), ptr(z_x), _p(wx), B, Nm * K, _coef(coef, 0), ptr(dsx), s, what="graph_mse_bwd")
        tp.accum(score_x, dsx)
        dsa = tp.empty(B, Nm, Nm)
        tp._call(L.molsde_graph_mse_bwd, ptr(score_adj), ptr(z_adj), _p(wa), B, Nm * Nm, _coef(coef, 1), ptr(dsa), s, what="graph_mse_bwd")
        draw = tp.empty(B * Nm * Nm, 1)
        tp._call(L.molsde_dense_edge_final_bwd, ptr(dsa), ptr(flags), ptr(scale_adj), B, Nm, ptr(draw), s, what="dense_edge_final_bwd")
        tp.accum(raw, draw)
    tp.ops.append(loss_bwd)
    return loss_x, loss_adj


# ======================================================================================================
# the training iteration (pretrain_MoleculeSDE.py:124-152)
# ======================================================================================================
class PretrainStep:
    """`train()` loop body of the reference for one batch: four modules in train mode, loss =
    c_CL * dual_CL + c_2D3D * loss_2Dto3D + c_3D2D * (loss_x + loss_adj)/2, backward, data-parallel all-reduce, Adam."""

    def __init__(self, gnn, schnet, sde_2d3d, sde_3d2d, device, lr: float = 1e-4, T: float = 0.1, coeff_contrastive: float = 1.0,
                 coeff_2Dto3D: float = 1.0, coeff_3Dto2D: float = 1.0, anneal_power: float = 0.0, gnn_2d_lr_scale: float = 1.0,
                 gnn_3d_lr_scale: float = 1.0, weight_decay: float = 0.0):
        self.gnn, self.schnet, self.m23, self.m32 = gnn, schnet, sde_2d3d, sde_3d2d
        self.dev = device
        self.store = ParamStore({"gnn": gnn, "schnet": schnet, "sde2d3d": sde_2d3d, "sde3d2d": sde_3d2d}, device)
        self.lr, self.T, self.anneal_power, self.weight_decay = lr, T, anneal_power, weight_decay
        self.c_cl, self.c_23, self.c_32 = coeff_contrastive, coeff_2Dto3D, coeff_3Dto2D
        # pretrain_MoleculeSDE.py:331-335: 2D GNN and 2D->3D share gnn_2d_lr_scale; SchNet and 3D->2D share gnn_3d_lr_scale
        self.lr_scale = {"gnn": gnn_2d_lr_scale, "sde2d3d": gnn_2d_lr_scale, "schnet": gnn_3d_lr_scale, "sde3d2d": gnn_3d_lr_scale}
        self.launches = 0
        self._streams = None
        for m in (gnn, schnet, sde_2d3d, sde_3d2d):
            m.train()

    def prepare(self, batch, max_nodes: Optional[int] = None):
        """Build, on the CURRENT stream, every per-batch index structure the iteration needs and attach it to `batch`: extended
        graph (`dataset_3D.py:12-35`) with its CSR / tile plan / by-source inverse, bond CSR and feature bucket indices of the GIN,
        radius graph + edge features of SchNet, dense-batch pointers.  The step itself then launches no graph construction and
        makes no host sync.  `max_nodes` (largest molecule, known on the host from `batch.ptr`) avoids the one device read the
        dense 3D->2D prologue would otherwise make.  Used by `loader.DeviceLoader` one batch ahead of the step."""
        from . import graph as G
        from .graph import segment_ptr
        cache = batch.__dict__.setdefault("_molsde_train_cache", {})
        if "z" not in cache:
            cache["z"] = batch.x[:, 0].contiguous()
        prepare_gin(cache, batch.x, batch.edge_index, batch.edge_attr, batch.batch, batch.num_graphs, True)
        prepare_schnet(cache, self.schnet, cache["z"], batch.positions, batch.batch, batch.num_graphs)
        if self.c_23 > 0:
            if self.m23.use_extend_graph and getattr(batch, "extended_edge_index", None) is None:
                csr = G.extend_graph(batch.edge_index, batch.batch, batch.num_graphs)
                batch.extended_edge_index, batch._molsde_ext_csr = csr.edge_index, csr
            prep = self.m23.prepared(batch)
            if getattr(prep, "_edge_set", None) is None:
                prep._edge_set = EdgeSet(prep.csr.rowptr, prep.csr.col, batch.batch, batch.num_graphs)
        if self.c_32 > 0:
            node_ptr = getattr(batch, "_molsde_node_ptr", None)
            if node_ptr is None:
                node_ptr = batch._molsde_node_ptr = segment_ptr(batch.batch, batch.num_graphs)
            if getattr(batch, "_molsde_dense_dims", None) is None and max_nodes is not None:
                batch._molsde_dense_dims = (batch.num_graphs, int(max_nodes), node_ptr)
        return batch

    def forward_backward(self, batch, draws: Optional[dict] = None, split: bool = False):
        """Forward + backward of one batch; gradients are left in `store.grad`.  `split=True` stops after the loss branches
        (forward of everything, backward of dual_CL / 2D->3D / 3D->2D: the gradients of the two SDE models are final) and returns
        `(out, finish)`; `finish()` runs the two encoder backwards.  Data-parallel callers put the all-reduce of the SDE models'
        gradient bucket between the two halves, so that it overlaps the encoders' backward (`step`, bench.py).  `draws` (parity tests):
        {"cl": (perm1, perm2), "sde2d3d": {...}, "sde3d2d": [randint, randn_adj, randn_x]}.

        The iteration is a fork/join graph, not a chain (pretrain_MoleculeSDE.py:131-147: the two encoders are independent, and
        each of the three loss terms depends only on the encoder outputs), so it is issued on three streams:

            main:  GIN forward ........ | 3D->2D forward+backward | d h2d = d_23 + d_CL -> GIN backward    |
            s1:    SchNet forward ..... | 2D->3D forward+backward | d h3d = d_32 + d_CL -> SchNet backward | join
            s2:                         | dual_CL forward+backward|

        Under CUDA-graph replay the branches overlap on the GPU (most kernels of this step are far smaller than 148 SMs).  Every
        branch owns its tape and its own gradient tensor for the shared representations, the two contributions are added after
        the join in a fixed order, and each branch writes a disjoint slice of the flat gradient buffer — results are
        bit-identical to the single-stream order (`MOLSDE_SINGLE_STREAM=1`)."""
        caller = torch.cuda.current_stream(self.dev) if self.dev.type == "cuda" else None
        if caller is None or _SINGLE_STREAM:
            out, fin = self._forward_backward(batch, draws or {}, caller, caller, caller, None, None)
            if split:
                return out, fin
            fin()
            return out
        if self._streams is None:
            # the dependent chains (branch streams) get the high priority, the weight-gradient side streams the low one: a
            # pending CTA of the critical path is scheduled before the leaves that only have to finish by the end of the step
            hi = -1 if _STREAM_PRIORITY else 0
            self._streams = tuple(torch.cuda.Stream(self.dev, priority=p) for p in (hi, hi, hi) + (0,) * (2 * _WGRAD_FANOUT))
        s0, s1, s2 = self._streams[:3]
        w0, w1 = list(self._streams[3:3 + _WGRAD_FANOUT]), list(self._streams[3 + _WGRAD_FANOUT:])
        if not (_WGRAD_STREAMS == "1" or (_WGRAD_STREAMS == "capture" and torch.cuda.is_current_stream_capturing())):
            w0 = w1 = None
        s0.wait_stream(caller)
        with torch.cuda.stream(s0):
            out, fin = self._forward_backward(batch, draws or {}, s0, s1, s2, w0, w1)
            if not split:
                fin()
        caller.wait_stream(s0)
        if not split:
            return out

        def finish():
            cur = torch.cuda.current_stream(self.dev)
            s0.wait_stream(cur)
            with torch.cuda.stream(s0):
                fin()
            cur.wait_stream(s0)
        return out, finish

    def _forward_backward(self, batch, draws, main, s1, s2, w0, w1) -> Dict[str, torch.Tensor]:
        st = self.store
        st.zero_grad()
        cache = batch.__dict__.setdefault("_molsde_train_cache", {})
        z = cache.get("z")
        if z is None:
            z = cache["z"] = batch.x[:, 0].contiguous()

        def fork(dst, *srcs):
            for src in srcs:
                if dst is not src:
                    dst.wait_stream(src)

        # ---- encoders: GIN on main, SchNet on s1
        fork(s1, main)
        tp_g = Tape(self.dev, w0)
        h2d = tape_gin(tp_g, self.gnn, st.vars("gnn"), batch.x, batch.edge_index, batch.edge_attr, cache, batch.batch, batch.num_graphs)
        with torch.cuda.stream(s1):
            tp_s = Tape(self.dev, w1)
            h3d = tape_schnet(tp_s, self.schnet, st.vars("schnet"), z, batch.positions, batch.batch, batch.num_graphs, cache)
        # ---- loss branches, each forward + backward on its own stream with private gradient tensors for h2d / h3d
        out = {}
        keep = []   # cross-stream tensors stay referenced until the final join (the caching allocator reuses per stream)
        g2d, g3d = [], []
        fork(s2, main, s1)
        fork(s1, main)
        fork(main, s1)
        if self.c_cl > 0:
            with torch.cuda.stream(s2):
                tp = Tape(self.dev)
                x2, x3 = Var(h2d.data, h2d.needs), Var(h3d.data, h3d.needs)
                n1, n2 = draws.get("cl", (None, None))
                out["cl_loss"], accs = tape_dual_cl(tp, x2, x3, self.T, n1, n2, coef=self.c_cl)
                out["cl_acc_pair"] = accs
                tp.backward()
                keep += [tp, x2, x3]
                cl2, cl3 = x2, x3
        else:
            cl2 = cl3 = None
        if self.c_23 > 0:
            with torch.cuda.stream(s1):
                tp = Tape(self.dev, w1)
                x2 = Var(h2d.data, h2d.needs)
                out["loss_2d3d"] = tape_2d3d(tp, self.m23, st.vars("sde2d3d"), x2, batch, self.anneal_power, draws.get("sde2d3d"),
                                             coef=self.c_23)
                tp.backward()
                keep += [tp, x2]
                g2d.append(x2)
        if self.c_32 > 0:
            tp = Tape(self.dev, w0)
            x3 = Var(h3d.data, h3d.needs)
            out["loss_x"], out["loss_adj"] = tape_3d2d(tp, self.m32, st.vars("sde3d2d"), x3, batch, self.anneal_power,
                                                       draws.get("sde3d2d"), coef=0.5 * self.c_32)
            tp.backward()
            keep += [tp, x3]
            g3d.append(x3)
        if cl2 is not None:
            g2d.append(cl2)
            g3d.append(cl3)
        launches = sum(t.launches for t in keep if isinstance(t, Tape))
        # ---- join: every loss branch is done (the gradients of sde2d3d / sde3d2d are final here)
        fork(main, s1, s2)
        fork(s1, main, s2)
        out["h2d"], out["h3d"] = h2d, h3d
        out["_keep"] = keep

        def finish():
            """add the branch gradients (fixed order: SDE branch, then CL), encoder backwards on main / s1"""
            fork(s1, main)
            for v in g2d:
                if v.grad is not None:
                    tp_g.accum(h2d, v.grad)
            tp_g.backward()
            with torch.cuda.stream(s1):
                for v in g3d:
                    if v.grad is not None:
                        tp_s.accum(h3d, v.grad)
                tp_s.backward()
            fork(main, s1)
            self.launches = launches + tp_g.launches + tp_s.launches
        return out, finish

    SDE_BUCKET, ENCODER_BUCKET = ("sde2d3d", "sde3d2d"), ("gnn", "schnet")

    def step(self, batch, draws: Optional[dict] = None) -> Dict[str, torch.Tensor]:
        """One iteration.  Data-parallel: the gradient all-reduce runs in two buckets -- the SDE models' gradients are final when
        the loss branches end, their exchange overlaps the encoders' backward; the encoders' bucket follows."""
        import torch.distributed as dist
        skip = [m for m, c in (("sde2d3d", self.c_23), ("sde3d2d", self.c_32)) if not c > 0]
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            out, finish = self.forward_backward(batch, draws, split=True)
            scale, work = self.store.all_reduce(self.SDE_BUCKET, async_op=True)
            finish()
            work.wait()
            self.store.all_reduce(self.ENCODER_BUCKET)
        else:
            out = self.forward_backward(batch, draws)
            scale = 1.0
        self.store.adam_step(self.lr, self.lr_scale, weight_decay=self.weight_decay, grad_scale=scale, skip=skip)
        return out

    @staticmethod
    def total_loss(out, c_cl=1.0, c_23=1.0, c_32=1.0) -> float:
        t = 0.0
        if "cl_loss" in out:
            t += c_cl * float(out["cl_loss"])
        if "loss_2d3d" in out:
            t += c_23 * float(out["loss_2d3d"])
        if "loss_x" in out:
            t += c_32 * 0.5 * (float(out["loss_x"]) + float(out["loss_adj"]))
        return t
