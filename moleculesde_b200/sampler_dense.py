"""3D->2D predictor-corrector sampler over (x, adj) with the reference's interface
(`examples/pretrain_MoleculeSDE_inference_3D_to_2D_VE_VP.py:95-252`).  Every score-network evaluation and every
elementwise update is a kernel of `csrc/dense.cu` / `tc_gemm.cu` / `mlp_rows.cu`; one predictor-corrector step is captured in a
CUDA graph and replayed per step (`node_adj_PC_generation(use_graph=...)`), the eager loop stays for injected draws.
"""
from __future__ import annotations

import os
from typing import Callable, Optional

import torch

from ._abi import check, lib, ptr, require_device, stream_ptr
from .sde_3d_to_2d import gen_noise, graph_reduce, mask_adjs, mask_x, node_flags


def _per_graph_consts(sde, t):
    """(sqrt_alpha[B], G[B]) of `sde.discretize` (SDE_dense.py:158-166 / 218-225) for per-graph t."""
    _, G = sde.discretize(torch.zeros(t.numel(), 1, 1, device=t.device), t)
    if hasattr(sde, "alphas"):
        ts = (t * (sde.N - 1) / sde.T).long()
        sa = torch.sqrt(sde._on("alphas", t.device)[ts])
    else:
        sa = torch.ones_like(t)
    return sa.float().contiguous(), G.float().contiguous()


class ReverseDiffusionPredictor:
    """`:167-190`"""

    def __init__(self, obj, sde, SDE_model, score_fn, probability_flow=False):
        if probability_flow:
            raise NotImplementedError
        self.obj, self.sde, self.SDE_model, self.score_fn = obj, sde, SDE_model, score_fn
        self.rsde = sde.reverse(score_fn, probability_flow)

    @torch.no_grad()
    def update_fn(self, representation, x, adj, flags, t, raw_noise=None, emb=None):
        """`emb` (keyword extension): `SDE_model.embed(representation, x)` when the caller already has it -- the adjacency and
        the node update of one half-step embed the same (representation, x)."""
        if emb is None:
            emb = self.SDE_model.embed(representation, x)
        score = self.score_fn(emb, adj, flags, t)
        cur = adj if self.obj == "adj" else x
        z = gen_noise(cur, flags, sym=(self.obj == "adj"), raw=raw_noise)
        sa, G = _per_graph_consts(self.sde, t)
        cur = cur.contiguous()
        new, mean = torch.empty_like(cur), torch.empty_like(cur)
        B = cur.size(0)
        check(lib().molsde_reverse_update(ptr(cur), ptr(score.contiguous()), ptr(z), ptr(sa), ptr(G), B, cur.numel() // B, ptr(new),
                                          ptr(mean), stream_ptr(cur)), "reverse_update")
        return new, mean


class LangevinCorrector:
    """`:208-252`.  As in the reference, `isinstance(sde, VPSDE)` there tests against the *sparse* VPSDE class
    (`:21`), so for the dense model's SDE objects alpha is always 1."""

    def __init__(self, obj, sde, SDE_model, score_fn, snr, scale_eps, n_steps):
        self.obj, self.sde, self.SDE_model, self.score_fn = obj, sde, SDE_model, score_fn
        self.snr, self.scale_eps, self.n_steps = snr, scale_eps, n_steps

    @torch.no_grad()
    def update_fn(self, representation, x, adj, flags, t, raw_noise=None, emb=None):
        if emb is None:
            emb = self.SDE_model.embed(representation, x)
        cur = (adj if self.obj == "adj" else x).contiguous()
        B = cur.size(0)
        M = cur.numel() // B
        s = stream_ptr(cur)
        mean = cur
        for it in range(self.n_steps):  # the reference re-evaluates on the updated state for the dense sampler (:236-252)
            grad = self.score_fn(emb, adj if self.obj == "x" else cur, flags, t).contiguous()
            noise = gen_noise(cur, flags, sym=(self.obj == "adj"), raw=raw_noise if it == self.n_steps - 1 else None)
            gn, nn = graph_reduce(grad, None, None, 0), graph_reduce(noise, None, None, 0)
            step = torch.empty(B, dtype=torch.float32, device=cur.device)
            check(lib().molsde_langevin_step(ptr(gn), ptr(nn), None, B, float(self.snr), ptr(step), s), "langevin_step")
            new, mean = torch.empty_like(cur), torch.empty_like(cur)
            check(lib().molsde_langevin_update(ptr(cur), ptr(grad), ptr(noise), ptr(step), B, M, float(self.scale_eps), ptr(new),
                                               ptr(mean), s), "langevin_update")
            cur = new
        return cur, mean


@torch.no_grad()
def node_adj_PC_generation(representation, data, SDE_model, B, max_num_nodes, num_class_X, probability_flow=False,
                           eps=1e-4, snr=0.2, scale_eps=0.9, n_steps=1, *, x_init=None, adj_init=None,
                           draws: Optional[Callable[[str, int], torch.Tensor]] = None, diffusion_steps: Optional[int] = None,
                           use_graph: Optional[bool] = None, return_graph: bool = False):
    """Reference signature (`:95-101`).  Keyword-only extensions: `x_init` / `adj_init` inject the prior draws,
    `draws(kind, step)` with kind in {'c_adj','c_x','p_adj','p_x'} injects the raw `randn_like` of each update,
    `diffusion_steps` truncates the `linspace(T, eps, N)` grid, `use_graph` (default: on when no draws are injected) captures
    ONE predictor-corrector step -- four score-network evaluations + the four state updates, ~170 kernels -- in a CUDA graph
    over static state buffers and replays it for every step (the step's time comes from a device-side counter), so a
    trajectory costs one graph launch per step instead of ~230 host launches.  `return_graph=True` returns
    `(GraphedPCStep, x0, adj0)` instead of running the trajectory."""
    require_device(representation)
    dev = representation.device
    sde_x, sde_adj = SDE_model.sde_x, SDE_model.sde_adj
    score_fn_x = SDE_model.get_score_fn(sde_x, SDE_model.node_score_network, train=False, continuous=True)
    score_fn_adj = SDE_model.get_score_fn(sde_adj, SDE_model.edge_score_network, train=False, continuous=True)
    pred_x = ReverseDiffusionPredictor("x", sde_x, SDE_model, score_fn_x, probability_flow)
    corr_x = LangevinCorrector("x", sde_x, SDE_model, score_fn_x, snr, scale_eps, n_steps)
    pred_adj = ReverseDiffusionPredictor("adj", sde_adj, SDE_model, score_fn_adj, probability_flow)
    corr_adj = LangevinCorrector("adj", sde_adj, SDE_model, score_fn_adj, snr, scale_eps, n_steps)
    if x_init is None:
        x_init = sde_x.prior_sampling((B, max_num_nodes, num_class_X))
    if adj_init is None:
        adj_init = sde_adj.prior_sampling((B, max_num_nodes, max_num_nodes))
    # flags from the oracle adjacency (:121-126)
    adj_oracle, _, _, flags, Nm = SDE_model.dense_inputs(torch.zeros(data.batch.numel(), 1, device=dev), data)
    assert Nm == max_num_nodes
    x = mask_x(x_init.to(dev).float(), flags)
    adj = mask_adjs(adj_init.to(dev).float(), flags)
    N = sde_adj.N
    steps = N if diffusion_steps is None else diffusion_steps
    timesteps = torch.linspace(sde_adj.T, eps, N, device=dev)
    get = (lambda k, i: None) if draws is None else draws
    # embedding_3D(representation) does not depend on the state: evaluated once per trajectory (the reference recomputes the
    # same values inside every one of its 4 x N_steps embeds, `:228,240`)
    rep3d = SDE_model.embed_3d(representation)

    side = {"stream": None}   # set by the graph path: the node-network half of every half-step runs on a second stream

    def both(upd_adj, upd_x, x, adj, vec_t, emb, za, zx):
        """adjacency update (edge score network) and node update (node score network) of one half-step: independent given
        (x, adj, emb) -- under graph capture they are forked onto two streams and overlap (the node network is three large GEMMs,
        the edge network many small kernels)."""
        s2 = side["stream"]
        if s2 is None:
            a = upd_adj.update_fn(representation, x, adj, flags, vec_t, za, emb=emb)
            b = upd_x.update_fn(representation, x, adj, flags, vec_t, zx, emb=emb)
            return a, b
        cur = torch.cuda.current_stream(dev)
        s2.wait_stream(cur)
        with torch.cuda.stream(s2):
            b = upd_x.update_fn(representation, x, adj, flags, vec_t, zx, emb=emb)
        a = upd_adj.update_fn(representation, x, adj, flags, vec_t, za, emb=emb)
        cur.wait_stream(s2)
        for t in b:
            t.record_stream(cur)
        return a, b

    def pc_step(x, adj, vec_t, i, get=get):
        """one iteration of the reference loop (`:134-147`)"""
        if get is None:
            get = lambda k, i: None  # noqa: E731
        emb = SDE_model.embed(representation, x, rep3d=rep3d)
        (adj1, _), (x1, _) = both(corr_adj, corr_x, x, adj, vec_t, emb, get("c_adj", i), get("c_x", i))
        emb = SDE_model.embed(representation, x1, rep3d=rep3d)
        (adj2, adj_mean), (x2, x_mean) = both(pred_adj, pred_x, x1, adj1, vec_t, emb, get("p_adj", i), get("p_x", i))
        return x2, adj2, x_mean, adj_mean

    if use_graph is None:
        use_graph = draws is None and steps >= 8
    if not use_graph:
        x_mean, adj_mean = x, adj
        for i in range(steps):
            x, adj, x_mean, adj_mean = pc_step(x, adj, torch.ones(B, device=dev) * timesteps[i], i)
        return x, adj, x_mean, adj_mean

    # ---- graph replay: static state, device-side step counter ----
    if os.environ.get("MOLSDE_DENSE_ONE_STREAM") != "1":    # measured (256 graphs x 64 atoms, device-timed replays): 4.81 ms/step on one
        side["stream"] = torch.cuda.Stream(device=dev)      # stream, 4.44 with the node network forked onto a second one
    pc = GraphedPCStep(pc_step, x, adj, timesteps, [sde_x, sde_adj], draws, steps)
    if return_graph:   # (x, adj) masked prior draws + the captured step: the caller drives `reset` / `run`
        return pc, x, adj
    pc.run(steps)
    return pc.x, pc.adj, pc.x_mean, pc.adj_mean


class GraphedPCStep:
    """One predictor-corrector step of `node_adj_PC_generation` captured in a CUDA graph over static state buffers.  The step's
    time is read from `timesteps[counter]` with a device-side counter that the captured step advances, so `run(n)` is n graph
    launches and nothing else.  `reset(x, adj, step_index)` rewinds / re-seeds the state (used by the benchmark, where an
    UNTRAINED network leaves the basin of finite states after ~100 steps)."""

    def __init__(self, pc_step, x, adj, timesteps, sdes, draws=None, steps=None):
        dev = x.device
        for sde in sdes:
            if hasattr(sde, "to_device"):
                sde.to_device(dev)   # schedule tables resident on the device: no host copy inside the captured step
        self.x, self.adj = x.clone(), adj.clone()
        self.x_mean, self.adj_mean = torch.empty_like(x), torch.empty_like(adj)
        self.idx = torch.zeros(1, dtype=torch.long, device=dev)
        ones = torch.ones(x.size(0), device=dev)
        self._pc_step = pc_step
        if draws is not None:   # injected draws under replay: per-kind tables [steps, ...] indexed by the device-side step counter
            tables = {k: torch.stack([draws(k, i).to(dev).float() for i in range(steps)]) for k in ("c_adj", "c_x", "p_adj", "p_x")}
            self.get = lambda k, i: tables[k].index_select(0, self.idx).squeeze(0)
        else:
            self.get = None

        def body():
            vec_t = ones * timesteps.index_select(0, self.idx)
            x2, adj2, x_mean, adj_mean = pc_step(self.x, self.adj, vec_t, 0, self.get)
            self.x.copy_(x2); self.adj.copy_(adj2); self.x_mean.copy_(x_mean); self.adj_mean.copy_(adj_mean)
            self.idx.add_(1)

        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            body()                                   # warm-up outside capture (weight packs, lazy inits), then restore the state
            side.synchronize()
            self.reset(x, adj)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph, stream=side):
                body()
            self.reset(x, adj)                       # (capture does not execute)
        torch.cuda.current_stream(dev).wait_stream(side)

    def reset(self, x, adj, step_index: int = 0):
        self.x.copy_(x)
        self.adj.copy_(adj)
        self.idx.fill_(step_index)

    def run(self, n: int):
        for _ in range(n):
            self.graph.replay()
