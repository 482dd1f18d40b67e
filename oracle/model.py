"""CPU restatement (plain torch fp32) of the reference's hot-path modules.

TEST INFRASTRUCTURE -- see `oracle/__init__.py`.  Functional style: every function takes
the reference `state_dict` (same key names) plus plain tensors, and cites the reference
file:line it follows (paths relative to `/root/reference`).  Pinned against the reference's
own sources executed over `oracle/shims` by `tests/test_oracle_vs_reference.py` (container)
and `tests/golden/*.pt` (everywhere).  Third-party op semantics come from `ref_ops`
(unpinned by the reference; brute-force tested).
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import Tensor

from . import ref_ops as R

EPSILON = 1e-6  # SDE_model_2D_to_3D.py:10


def _sub(sd: Dict[str, Tensor], prefix: str) -> Dict[str, Tensor]:
    n = len(prefix)
    return {k[n:]: v for k, v in sd.items() if k.startswith(prefix)}


def _lin(sd, name, x):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


def mlp(sd: Dict[str, Tensor], prefix: str, x: Tensor, act) -> Tensor:
    """MultiLayerPerceptron.forward, `layers/common.py:31-40` (activation between layers,
    none after the last; dropout=0 everywhere on the hot path)."""
    n = 0
    while f"{prefix}.layers.{n}.weight" in sd:
        n += 1
    for i in range(n):
        x = _lin(sd, f"{prefix}.layers.{i}", x)
        if i < n - 1:
            x = act(x)
    return x


# ----------------------------------------------------------------------------
# SDE schedules (per-node t): SDE_sparse.py
# ----------------------------------------------------------------------------
class VESDE:
    """`SDE_sparse.py:172-230`."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
        self.sigma_min, self.sigma_max, self.N = sigma_min, sigma_max, N
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))  # :183
        self.T = 1

    def marginal_prob(self, x, t):  # :196-199
        std = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        return x, std

    def discretize(self, x, t):  # :214-222
        timestep = (t * (self.N - 1) / self.T).long()
        sigma = self.discrete_sigmas[timestep]
        adjacent = torch.where(timestep == 0, torch.zeros_like(t), self.discrete_sigmas[timestep - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)

    def corrector_alpha(self, t):  # inference_2D_to_3D:198-202
        return torch.ones_like(t)


class VPSDE:
    """`SDE_sparse.py:105-169`."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        self.beta_0, self.beta_1, self.N = beta_min, beta_max, N
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)  # :117
        self.alphas = 1.0 - self.discrete_betas
        self.T = 1

    def marginal_prob(self, x, t):  # :134-139
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        mean = torch.exp(log_mean_coeff[:, None]) * x
        std = torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))
        return mean, std

    def discretize(self, x, t):  # :154-162
        timestep = (t * (self.N - 1) / self.T).long()
        beta = self.discrete_betas[timestep]
        alpha = self.alphas[timestep]
        f = torch.sqrt(alpha)[:, None] * x - x
        return f, torch.sqrt(beta)

    def corrector_alpha(self, t):  # inference_2D_to_3D:198-200
        timestep = (t * (self.N - 1) / self.T).long()
        return self.alphas[timestep]


def make_sde(sde_type: str, beta_min: float, beta_max: float, N: int):
    """`SDE_model_2D_to_3D.py:276-279`."""
    if sde_type == "VE":
        return VESDE(sigma_min=beta_min, sigma_max=beta_max, N=N)
    if sde_type == "VP":
        return VPSDE(beta_min=beta_min, beta_max=beta_max, N=N)
    raise NotImplementedError(sde_type)


# ----------------------------------------------------------------------------
# 2D -> 3D score model: SDE_model_2D_to_3D.py (SDEModel2Dto3D_02)
# ----------------------------------------------------------------------------
def gaussian_fourier(W: Tensor, x: Tensor) -> Tensor:
    """GaussianFourierProjection.forward, `SDE_model_2D_to_3D.py:64-66`."""
    x_proj = x * W[None, :] * 2 * np.pi
    return torch.cat([torch.sin(x_proj), torch.cos(x_proj)], dim=-1)


def coord2basis(pos: Tensor, row: Tensor, col: Tensor):
    """`SDE_model_2D_to_3D.py:35-47`.  The reference calls `torch.cross` without `dim`
    (first size-3 dim); restated with dim=1, identical whenever E != 3."""
    coord_diff = pos[row] - pos[col]
    radial = torch.sum(coord_diff ** 2, 1).unsqueeze(1)
    coord_cross = torch.linalg.cross(pos[row], pos[col], dim=1)
    norm = torch.sqrt(radial) + EPSILON
    coord_diff = coord_diff / norm
    cross_norm = torch.sqrt(torch.sum(coord_cross ** 2, 1).unsqueeze(1)) + EPSILON
    coord_cross = coord_cross / cross_norm
    coord_vertical = torch.linalg.cross(coord_diff, coord_cross, dim=1)
    return coord_diff, coord_cross, coord_vertical


def edge_2d_emb(sd, h2d: Tensor, ei: Tensor, training: bool = False,
                stats: Optional[dict] = None) -> Tensor:
    """`edge_2D_emb` = Linear(600,300) -> BatchNorm1d(300) -> ReLU -> Linear(300,32) on
    cat[h[row], h[col]]; `SDE_model_2D_to_3D.py:265,345-347` (train) / `:405-407` (eval).
    With `training=True` batch statistics are used and, if `stats` is a dict, the updated
    running stats (momentum 0.1, unbiased variance) are written into it."""
    row, col = ei
    x = torch.cat([h2d[row], h2d[col]], dim=-1)
    x = _lin(sd, "edge_2D_emb.0", x)
    rm, rv = sd["edge_2D_emb.1.running_mean"], sd["edge_2D_emb.1.running_var"]
    if training:
        rm, rv = rm.clone(), rv.clone()
        x = F.batch_norm(x, rm, rv, sd["edge_2D_emb.1.weight"], sd["edge_2D_emb.1.bias"], True, 0.1, 1e-5)
        if stats is not None:
            stats["running_mean"], stats["running_var"] = rm, rv
    else:
        x = F.batch_norm(x, rm, rv, sd["edge_2D_emb.1.weight"], sd["edge_2D_emb.1.bias"], False, 0.1, 1e-5)
    x = F.relu(x)
    return _lin(sd, "edge_2D_emb.3", x)


def edge_features_2d3d(sd, e2d: Tensor, ei: Tensor, pos: Tensor):
    """Per-edge geometric pipeline shared by forward and get_score,
    `SDE_model_2D_to_3D.py:342-372` == `:402-432`.  Returns (edge_attr [E,32], basis 3x[E,3])."""
    row, col = ei
    d = (pos[row] - pos[col]).norm(dim=-1).unsqueeze(-1)  # get_perturb_distance :50-54
    d_emb = gaussian_fourier(sd["dist_gaussian_fourier.W"], d)  # :349
    inv3d = mlp(sd, "input_mlp", d_emb, F.silu)  # :350 (single layer)
    coord_diff, coord_cross, coord_vertical = coord2basis(pos, row, col)  # :353
    edge_basis = torch.cat([coord_diff.unsqueeze(1), coord_cross.unsqueeze(1), coord_vertical.unsqueeze(1)], dim=1)
    r_i, r_j = pos[row], pos[col]
    coff_i = torch.matmul(edge_basis, r_i.unsqueeze(-1)).squeeze(-1)  # :357
    coff_j = torch.matmul(edge_basis, r_j.unsqueeze(-1)).squeeze(-1)
    coff_i = torch.stack([coff_i[:, 0], coff_i[:, 1].abs(), coff_i[:, 2]], dim=1)  # :359
    coff_j = torch.stack([coff_j[:, 0], coff_j[:, 1].abs(), coff_j[:, 2]], dim=1)  # :360
    coff_mul = coff_i * coff_j
    ni = coff_i.norm(dim=-1, keepdim=True)
    nj = coff_j.norm(dim=-1, keepdim=True)
    pseudo_cos = coff_mul.sum(dim=-1, keepdim=True) / (ni + EPSILON) / (nj + EPSILON)  # :364
    pseudo_sin = torch.sqrt(1 - pseudo_cos ** 2)  # :365
    pseudo_angle = torch.cat([pseudo_sin, pseudo_cos], dim=-1)

    def get_embedding(c):  # :297-304
        parts = [gaussian_fourier(sd["coff_gaussian_fourier.W"], c[:, i:i + 1]) for i in (0, 2)]
        return _lin(sd, "coff_mlp", torch.cat(parts, dim=-1))

    edge_embed = torch.cat([pseudo_angle, get_embedding(coff_i), get_embedding(coff_j)], dim=-1)  # :369
    frame_inv = mlp(sd, "project", edge_embed, F.silu)  # :370
    edge_attr = inv3d * e2d + frame_inv  # :372
    return edge_attr, (coord_diff, coord_cross, coord_vertical)


def gat_layer(sd, prefix: str, ei: Tensor, x: Tensor, edge_attr: Tensor,
              attn_mask: Optional[Tensor] = None, ffn_mask: Optional[Tensor] = None, p: float = 0.1) -> Tensor:
    """GATLayer.forward, `equivariant_scorenetwork.py:34-40` (TransformerConv heads=8, C=4;
    LayerNorm eps 1e-5; FFN Linear-SiLU-Dropout-Linear).  Masks reproduce train-mode dropout."""
    w = _sub(sd, prefix + ".MHA.")
    h = R.transformer_conv(x, ei, edge_attr, w, heads=8, out_channels=x.size(1) // 8,
                           dropout_mask=attn_mask, dropout_p=p)
    x = x + F.layer_norm(h, (x.size(1),), sd[prefix + ".norm1.weight"], sd[prefix + ".norm1.bias"], 1e-5)
    h = F.silu(_lin(sd, prefix + ".FFN.0", x))
    if ffn_mask is not None:
        h = h * ffn_mask / (1.0 - p)
    h = _lin(sd, prefix + ".FFN.3", h)
    return x + F.layer_norm(h, (x.size(1),), sd[prefix + ".norm2.weight"], sd[prefix + ".norm2.bias"], 1e-5)


def equivariant_score_network(sd, prefix: str, ei: Tensor, node_attr: Tensor, edge_attr: Tensor, basis,
                              dropout: Optional[List[Tuple[Tensor, Tensor]]] = None) -> Tensor:
    """EquivariantScoreNetwork.forward, `equivariant_scorenetwork.py:121-169` with
    short_cut=False, concat_hidden=False, num_layers=2, num_convs=2; EquiLayer with
    activation=False => message = basis_mix, aggr = mean (`:74-78,111`).  Returns gradient [N,3].
    `dropout`: 4 (attention mask [E,8], FFN mask [N,32]) pairs in layer order, or None (eval)."""
    coord_diff, coord_cross, coord_vertical = basis
    n = node_attr.size(0)
    conv_input = node_attr
    gradient = None
    li = 0
    for module_idx in range(2):
        for conv_idx in range(2):
            am, fm = dropout[li] if dropout is not None else (None, None)
            hidden = gat_layer(sd, f"{prefix}.gnn_layers.{module_idx}.{conv_idx}", ei, conv_input, edge_attr, am, fm)
            li += 1
            if conv_idx < 1:
                hidden = F.silu(hidden)  # :140-141
            conv_input = hidden
        node_feature = conv_input
        h_row, h_col = node_feature[ei[0]], node_feature[ei[1]]
        edge_feature = torch.cat([h_row + h_col, edge_attr], dim=-1)  # :155
        bp = f"{prefix}.basis_mlp_modules.{module_idx}"
        dyn = _lin(sd, bp + ".2", F.silu(_lin(sd, bp + ".0", edge_feature)))  # :158
        basis_mix = dyn[:, :1] * coord_diff + dyn[:, 1:2] * coord_cross + dyn[:, 2:3] * coord_vertical  # :159
        g = R.propagate(ei, basis_mix, n, "mean")  # :162-164
        gradient = g if gradient is None else gradient + g
    return gradient


def node_emb(sd, h2d: Tensor) -> Tensor:
    """`node_emb` = MultiLayerPerceptron(300,[32]) (single Linear), `SDE_model_2D_to_3D.py:264,375`."""
    return mlp(sd, "node_emb", h2d, F.silu)


def get_score_2d3d(sd, sde, h2d: Tensor, ei: Tensor, pos: Tensor, t: Tensor,
                   e2d: Optional[Tensor] = None, nattr: Optional[Tensor] = None) -> Tensor:
    """SDEModel2Dto3D_02.get_score, `SDE_model_2D_to_3D.py:393-445` (eval mode, no_grad).
    `e2d` / `nattr` may be passed in when hoisted out of a sampling loop (loop-invariant)."""
    with torch.no_grad():
        if nattr is None:
            nattr = node_emb(sd, h2d)
        if e2d is None:
            e2d = edge_2d_emb(sd, h2d, ei, training=False)
        edge_attr, basis = edge_features_2d3d(sd, e2d, ei, pos)
        grad = equivariant_score_network(sd, "score_network", ei, nattr, edge_attr, basis)
        _, std = sde.marginal_prob(pos, t)
        return -grad / std[:, None]


def perturb_2d3d(sde, pos: Tensor, batch: Tensor, num_graphs: int, noise: Tensor, time_step_half: Tensor, N: int):
    """Antithetic time sampling + perturbation kernel, `SDE_model_2D_to_3D.py:318-329`.
    `time_step_half` is the `randint(0,N,(B//2+1,))` draw."""
    ts = torch.cat([time_step_half, N - time_step_half - 1], dim=0)[:num_graphs]
    ts = ts / N * (1 - EPSILON) + EPSILON
    t_pos = ts.index_select(0, batch)
    mean, std = sde.marginal_prob(pos, t_pos)
    return mean + std[:, None] * noise, std, t_pos


def loss_2d3d(sd, sde, h2d: Tensor, ei: Tensor, pos: Tensor, batch: Tensor, num_graphs: int,
              noise: Tensor, time_step_half: Tensor, N: int = 1000, anneal_power: float = 0.0,
              dropout=None, training: bool = True, stats: Optional[dict] = None) -> Tensor:
    """SDEModel2Dto3D_02.forward, `SDE_model_2D_to_3D.py:306-391` with injected draws."""
    pos_p, std, _ = perturb_2d3d(sde, pos, batch, num_graphs, noise, time_step_half, N)
    e2d = edge_2d_emb(sd, h2d, ei, training=training, stats=stats)
    edge_attr, basis = edge_features_2d3d(sd, e2d, ei, pos_p)
    nattr = node_emb(sd, h2d)
    scores = equivariant_score_network(sd, "score_network", ei, nattr, edge_attr, basis, dropout)
    if anneal_power == 0:
        loss = torch.sum((scores - noise) ** 2, -1)
    else:
        loss = torch.sum((scores - noise) ** 2 * (std ** anneal_power).unsqueeze(1), -1)
    return R.scatter_mean(loss, batch, 0, num_graphs).mean()


# ----------------------------------------------------------------------------
# predictor-corrector sampler: examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py
# ----------------------------------------------------------------------------
def pc_sample_2d3d(sd, sde, h2d: Tensor, ei: Tensor, batch: Tensor, num_graphs: int, pos_init: Tensor,
                   noise_corr: Tensor, noise_pred: Tensor, n_diff_steps: Optional[int] = None,
                   eps: float = 1e-4, snr: float = 0.2, scale_eps: float = 0.9,
                   record: bool = False):
    """position_PC_generation (`:92-138`) with LangevinCorrector.update_fn (`:191-212`, n_steps=1;
    larger n_steps only changes which noise draw is used, F9) and
    ReverseDiffusionPredictor.update_fn (`:163-168`) over RSDE.discretize (`SDE_sparse.py:94-100`).
    `noise_corr[i]`, `noise_pred[i]` are the injected `randn_like(pos)` draws of step i.
    Returns (pos, pos_mean) and, if `record`, the per-step list of (pos_in, score_c, pos_c, score_p)."""
    N = sde.N
    steps = N if n_diff_steps is None else n_diff_steps
    timesteps = torch.linspace(sde.T, eps, N)  # :118
    e2d = edge_2d_emb(sd, h2d, ei, training=False)
    nattr = node_emb(sd, h2d)
    pos = pos_init
    pos_mean = pos_init
    trace = []
    for i in range(steps):
        t = timesteps[i]
        vec_t = (torch.ones(num_graphs) * t).index_select(0, batch)  # :121-123
        # corrector :191-212
        alpha = sde.corrector_alpha(vec_t)
        grad = get_score_2d3d(sd, sde, h2d, ei, pos, vec_t, e2d, nattr)
        noise = noise_corr[i]
        grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
        noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
        step_size = (snr * noise_norm / grad_norm) ** 2 * 2 * alpha
        x_mean = pos + step_size[:, None] * grad
        pos_c = x_mean + torch.sqrt(step_size * 2)[:, None] * noise * scale_eps
        # predictor :163-168 + SDE_sparse.py:94-100
        f, G = sde.discretize(pos_c, vec_t)
        score = get_score_2d3d(sd, sde, h2d, ei, pos_c, vec_t, e2d, nattr)
        rev_f = f - G[:, None] ** 2 * score
        pos_mean = pos_c - rev_f
        pos_new = pos_mean + G[:, None] * noise_pred[i]
        if record:
            trace.append((pos, grad, pos_c, score, pos_mean, pos_new))
        pos = pos_new
    return (pos, pos_mean, trace) if record else (pos, pos_mean)


# ----------------------------------------------------------------------------
# SchNet: Geom3D/models/schnet.py
# ----------------------------------------------------------------------------
def shifted_softplus(x: Tensor) -> Tensor:
    """`schnet.py:210-216`: softplus(x) - float32(log 2)."""
    return F.softplus(x) - torch.log(torch.tensor(2.0)).item()


def schnet_forward(sd, z: Tensor, pos: Tensor, batch: Tensor, num_graphs: int, cutoff: float = 10.0,
                   readout: str = "mean", edge_index: Optional[Tensor] = None):
    """SchNet.forward(return_latent=True), `schnet.py:85-125` (dipole=False, no mean/std/atomref).
    Returns (out [B,H], h [N,H], edge_index [2,E_r])."""
    h = sd["embedding.weight"][z]  # :89
    ei = R.radius_graph(pos, cutoff, batch) if edge_index is None else edge_index  # :91
    row, col = ei
    edge_weight = (pos[row] - pos[col]).norm(dim=-1)  # :93
    offset = sd["distance_expansion.offset"]
    coeff = -0.5 / (offset[1] - offset[0]).item() ** 2  # :202
    edge_attr = torch.exp(coeff * torch.pow(edge_weight.view(-1, 1) - offset.view(1, -1), 2))  # :205-207
    n = h.size(0)
    n_inter = 0
    while f"interactions.{n_inter}.lin.weight" in sd:
        n_inter += 1
    for i in range(n_inter):
        p = f"interactions.{i}"
        C = 0.5 * (torch.cos(edge_weight * math.pi / cutoff) + 1.0)  # :186
        W = _lin(sd, p + ".mlp.2", shifted_softplus(_lin(sd, p + ".mlp.0", edge_attr))) * C.view(-1, 1)  # :187
        x = F.linear(h, sd[p + ".conv.lin1.weight"])  # :189
        x = R.propagate(ei, x[row] * W, n, "add")  # :190,194-195
        x = _lin(sd, p + ".conv.lin2", x)  # :191
        x = _lin(sd, p + ".lin", shifted_softplus(x))  # :165-166
        h = h + x  # :97
    h = _lin(sd, "lin2", shifted_softplus(_lin(sd, "lin1", h)))  # :99-101
    out = R.scatter(h, batch, 0, num_graphs, readout)  # :115
    return out, h, ei


# ----------------------------------------------------------------------------
# contrastive loss: examples/util.py
# ----------------------------------------------------------------------------
def do_cl_ebm_node_dot_prod(X: Tensor, Y: Tensor, T: float, neg_index: Tensor):
    """do_CL, metric EBM_node_dot_prod, `examples/util.py:52-68` with injected `randperm`."""
    pred_pos = torch.sum(X * Y, dim=1) / T
    pred_neg = torch.sum(X * Y[neg_index], dim=1) / T
    loss_pos = F.binary_cross_entropy_with_logits(pred_pos, torch.ones_like(pred_pos))
    loss_neg = F.binary_cross_entropy_with_logits(pred_neg, torch.zeros_like(pred_neg))
    acc = (torch.sum(pred_pos > 0).float() + torch.sum(pred_neg < 0).float()) / (len(pred_pos) + len(pred_neg))
    return loss_pos + loss_neg, acc.item()


def dual_cl(X: Tensor, Y: Tensor, T: float, neg_index_1: Tensor, neg_index_2: Tensor):
    """dual_CL, `examples/util.py:76-79`."""
    l1, a1 = do_cl_ebm_node_dot_prod(X, Y, T, neg_index_1)
    l2, a2 = do_cl_ebm_node_dot_prod(Y, X, T, neg_index_2)
    return (l1 + l2) / 2, (a1 + a2) / 2
