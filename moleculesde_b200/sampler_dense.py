"""3D->2D predictor-corrector sampler over (x, adj) with the reference's interface
(`examples/pretrain_MoleculeSDE_inference_3D_to_2D_VE_VP.py:95-252`).  Host-driven loop this round:
every score-network evaluation and every elementwise update is a kernel of `csrc/dense.cu`.
"""
from __future__ import annotations

from typing import Callable, Optional

import torch

from ._abi import check, lib, ptr, require_device, stream_ptr
from .sde_3d_to_2d import gen_noise, graph_reduce, mask_adjs, mask_x, node_flags


def _per_graph_consts(sde, t):
    """(sqrt_alpha[B], G[B]) of `sde.discretize` (SDE_dense.py:158-166 / 218-225) for per-graph t."""
    _, G = sde.discretize(torch.zeros(t.numel(), 1, 1, device=t.device), t)
    if hasattr(sde, "alphas"):
        ts = (t * (sde.N - 1) / sde.T).long()
        sa = torch.sqrt(sde.alphas.to(t.device)[ts])
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
    def update_fn(self, representation, x, adj, flags, t, raw_noise=None):
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
    def update_fn(self, representation, x, adj, flags, t, raw_noise=None):
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
                           draws: Optional[Callable[[str, int], torch.Tensor]] = None, diffusion_steps: Optional[int] = None):
    """Reference signature (`:95-101`).  Keyword-only extensions: `x_init` / `adj_init` inject the prior draws,
    `draws(kind, step)` with kind in {'c_adj','c_x','p_adj','p_x'} injects the raw `randn_like` of each update,
    `diffusion_steps` truncates the `linspace(T, eps, N)` grid."""
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
    x_mean, adj_mean = x, adj
    get = (lambda k, i: None) if draws is None else draws
    for i in range(steps):
        vec_t = torch.ones(B, device=dev) * timesteps[i]
        _x, _adj = x, adj
        adj, adj_mean = corr_adj.update_fn(representation, _x, _adj, flags, vec_t, get("c_adj", i))
        x, x_mean = corr_x.update_fn(representation, _x, _adj, flags, vec_t, get("c_x", i))
        _x, _adj = x, adj
        adj, adj_mean = pred_adj.update_fn(representation, _x, _adj, flags, vec_t, get("p_adj", i))
        x, x_mean = pred_x.update_fn(representation, _x, _adj, flags, vec_t, get("p_x", i))
    return x, adj, x_mean, adj_mean
