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
    has_dist = "input_mlp.layers.0.weight" in sd  # SDEModel2Dto3D_01 (:69-250) has no distance branch
    if has_dist:
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
    edge_attr = inv3d * e2d + frame_inv if has_dist else e2d + frame_inv  # :372 (_02) / :181 (_01)
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


# ----------------------------------------------------------------------------
# dense 3D -> 2D score model: SDE_model_3D_to_2D_node_adj_dense.py, invariant_scorenetwork_dense.py,
# layers/edge_network_dense.py, layers/node_network_dense.py, SDE_dense.py
# ----------------------------------------------------------------------------
def mask_x(x: Tensor, flags: Tensor) -> Tensor:
    """`SDE_model_3D_to_2D_node_adj_dense.py:559-562`."""
    return x * flags[:, :, None]


def mask_adjs(adjs: Tensor, flags: Tensor) -> Tensor:
    """`SDE_model_3D_to_2D_node_adj_dense.py:543-556` (B x N x N or B x C x N x N)."""
    if adjs.dim() == 4:
        flags = flags.unsqueeze(1)
    return adjs * flags.unsqueeze(-1) * flags.unsqueeze(-2)


def node_flags(adj: Tensor, eps: float = 1e-5) -> Tensor:
    """`:523-529`."""
    return torch.abs(adj).sum(-1).gt(eps).to(torch.float32)


def dense_gcn(weight: Tensor, bias: Tensor, x: Tensor, adj: Tensor) -> Tensor:
    """NodeNetwork_dense.forward (DenseGCNConv clone, add_loop=True, weight stored [in,out]),
    `layers/node_network_dense.py:46-85`."""
    n = adj.size(-1)
    adj = adj.clone()
    idx = torch.arange(n)
    adj[:, idx, idx] = 1
    out = torch.matmul(x, weight)
    dis = adj.sum(dim=-1).clamp(min=1).pow(-0.5)
    adj = dis.unsqueeze(-1) * adj * dis.unsqueeze(-2)
    return torch.matmul(adj, out) + bias


def edge_layer(sd, prefix: str, x: Tensor, adj: Tensor, num_heads: int):
    """EdgeLayer.forward with conv='MLP', `layers/edge_network_dense.py:55-82`."""
    Q = mlp(sd, prefix + ".func_q", x, torch.tanh)
    K = mlp(sd, prefix + ".func_k", x, torch.tanh)
    V = dense_gcn(sd[prefix + ".func_v.weight"], sd[prefix + ".func_v.bias"], x, adj)
    attn_dim = Q.size(-1) // 2  # hidden_dims = [2*attn_dim, 2*attn_dim] (:46)
    dim_split = attn_dim // num_heads
    Q_ = torch.cat(Q.split(dim_split, 2), 0)
    K_ = torch.cat(K.split(dim_split, 2), 0)
    A = torch.tanh(Q_.bmm(K_.transpose(1, 2)) / math.sqrt(dim_split))
    A = A.view(-1, *adj.shape).mean(dim=0)
    return V, (A + A.transpose(-1, -2)) / 2


def edge_network_dense(sd, prefix: str, x: Tensor, adj: Tensor, flags: Tensor, num_heads: int = 4):
    """EdgeNetwork_dense.forward, `layers/edge_network_dense.py:105-128`."""
    n_in = adj.size(1)
    masks, xs = [], []
    for c in range(n_in):
        V, A = edge_layer(sd, f"{prefix}.attn.{c}", x, adj[:, c], num_heads)
        masks.append(A.unsqueeze(-1))
        xs.append(V)
    x_out = torch.tanh(mask_x(mlp(sd, prefix + ".multi_channel", torch.cat(xs, dim=-1), F.elu), flags))
    mlp_in = torch.cat([torch.cat(masks, dim=-1), adj.permute(0, 2, 3, 1)], dim=-1)
    shape = mlp_in.shape
    out = mlp(sd, prefix + ".mlp", mlp_in.reshape(-1, shape[-1]), F.elu)
    _adj = out.view(shape[0], shape[1], shape[2], -1).permute(0, 3, 1, 2)
    _adj = _adj + _adj.transpose(-1, -2)
    return x_out, mask_adjs(_adj, flags)


def edge_score_network_dense(sd, prefix: str, x: Tensor, adj: Tensor, flags: Tensor, c_init: int = 2, num_layers: int = 4):
    """EdgeScoreNetwork_dense.forward, `invariant_scorenetwork_dense.py:74-93` (pow_tensor :28-37)."""
    chans = [adj.unsqueeze(1)]
    a = adj.clone()
    for _ in range(c_init - 1):
        a = torch.bmm(a, adj)
        chans.append(a.unsqueeze(1))
    adjc = torch.cat(chans, dim=1)
    adj_list = [adjc]
    for li in range(num_layers):
        x, adjc = edge_network_dense(sd, f"{prefix}.layers.{li}", x, adjc, flags)
        adj_list.append(adjc)
    adjs = torch.cat(adj_list, dim=1).permute(0, 2, 3, 1)
    score = mlp(sd, prefix + ".final", adjs, F.silu).view(*adjs.shape[:-1])
    n = adjs.size(1)
    score = score * (torch.ones(n, n) - torch.eye(n)).unsqueeze(0)
    return mask_adjs(score, flags)


def node_score_network_dense(sd, prefix: str, x: Tensor, adj: Tensor, flags: Tensor, depth: int = 4):
    """NodeScoreNetwork_dense.forward, `invariant_scorenetwork_dense.py:118-131`."""
    x_list = [x]
    for li in range(depth):
        x = torch.tanh(dense_gcn(sd[f"{prefix}.layers.{li}.weight"], sd[f"{prefix}.layers.{li}.bias"], x, adj))
        x_list.append(x)
    xs = torch.cat(x_list, dim=-1)
    out = mlp(sd, prefix + ".final", xs, F.silu).view(adj.shape[0], adj.shape[1], -1)
    return mask_x(out, flags)


class DenseVESDE(VESDE):
    """`SDE_dense.py:176-233`: per-graph t [B]."""

    def marginal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t


class DenseVPSDE(VPSDE):
    """`SDE_dense.py:110-173`."""

    def marginal_prob(self, x, t):
        lmc = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(lmc[:, None, None]) * x, torch.sqrt(1.0 - torch.exp(2.0 * lmc))

    def discretize(self, x, t):
        timestep = (t * (self.N - 1) / self.T).long()
        beta, alpha = self.discrete_betas[timestep], self.alphas[timestep]
        return torch.sqrt(alpha)[:, None, None] * x - x, torch.sqrt(beta)


def make_dense_sde(kind: str, beta_min: float, beta_max: float, N: int):
    return DenseVESDE(beta_min, beta_max, N) if kind == "VE" else DenseVPSDE(beta_min, beta_max, N)


def embed_3d2d(sd, rep_dense: Tensor, x: Tensor) -> Tensor:
    """`embedding_3D(h3D) + embedding_X(x)`, `SDE_model_3D_to_2D_node_adj_dense.py:156`."""
    return _lin(sd, "embedding_3D", rep_dense) + _lin(sd, "embedding_X", x)


def score_3d2d(sd, sde, which: str, emb: Tensor, adj: Tensor, flags: Tensor, t: Tensor) -> Tensor:
    """`get_score_fn(...)(x, adj, flags, t)`, `:68-99`: -net(x, adj, flags) / std[:, None, None]."""
    if which == "adj":
        out = edge_score_network_dense(sd, "edge_score_network", emb, adj, flags)
    else:
        out = node_score_network_dense(sd, "node_score_network", emb, adj, flags)
    std = sde.marginal_prob(torch.zeros_like(adj), t)[1]
    return -out / std[:, None, None]


def dense_inputs(h3d: Tensor, z: Tensor, edge_index: Tensor, bond_type: Tensor, batch: Tensor):
    """to_dense_adj / to_dense_batch prologue of forward, `:118-134`: returns (adj, rep_dense, z_dense, flags)."""
    edge_attr = bond_type.float() + 1
    B = int(batch.max().item()) + 1
    num_nodes = R.scatter(torch.ones_like(batch), batch, 0, B, "sum")
    nmax = int(num_nodes.max().item())
    adj = R.to_dense_adj(edge_index, batch, edge_attr=edge_attr, max_num_nodes=nmax)
    rep, _ = R.to_dense_batch(h3d, batch, max_num_nodes=nmax)
    zd, _ = R.to_dense_batch(z, batch, max_num_nodes=nmax)
    return adj, rep, zd, node_flags(adj)


def loss_3d2d(sd, sde_x, sde_adj, h3d: Tensor, z: Tensor, edge_index: Tensor, bond_type: Tensor, batch: Tensor,
              t_half: Tensor, noise_adj: Tensor, noise_x: Tensor, N: int = 1000, num_class: int = 119,
              anneal_power: float = 0.0):
    """SDEModel3Dto2D_node_adj_dense.forward (noise_on_one_hot, reduce_mean=True, continuous), `:101-179`,
    with the three random draws injected (`randint` :112, `randn_like(adj)` :135/533, `randn_like(one_hot)` :144)."""
    adj, rep, zd, flags = dense_inputs(h3d, z, edge_index, bond_type, batch)
    B = adj.size(0)
    t = torch.cat([t_half, N - t_half - 1], dim=0)[:B]
    t = t / N * (1 - EPSILON) + EPSILON
    z_adj = noise_adj.triu(1)
    z_adj = mask_adjs(z_adj + z_adj.transpose(-1, -2), flags)  # gen_noise(sym=True) :532-538
    mean_adj, std_adj = sde_adj.marginal_prob(adj, t)
    p_adj = mask_adjs(mean_adj + std_adj[:, None, None] * z_adj, flags)
    one_hot = F.one_hot(zd, num_class).float()
    z_x = mask_x(noise_x, flags)
    mean_x, std_x = sde_x.marginal_prob(one_hot, t)
    p_x = mask_x(mean_x + std_x[:, None, None] * z_x, flags)
    emb = embed_3d2d(sd, rep, p_x)
    s_adj = score_3d2d(sd, sde_adj, "adj", emb, p_adj, flags, t)
    s_x = score_3d2d(sd, sde_x, "x", emb, p_adj, flags, t)
    lx = torch.square(s_x + z_x)
    la = torch.square(s_adj + z_adj)
    if anneal_power != 0:
        lx = lx * (std_x ** anneal_power)[:, None, None]
        la = la * (std_adj ** anneal_power)[:, None, None]
    lx = torch.mean(lx.reshape(B, -1), dim=-1)
    la = torch.mean(la.reshape(B, -1), dim=-1)
    return torch.mean(lx), torch.mean(la)


# ----------------------------------------------------------------------------
# GIN 2D encoder: Geom3D/models/molecule_gnn_model.py
# ----------------------------------------------------------------------------
def gin_forward(sd, x: Tensor, edge_index: Tensor, edge_attr: Tensor, num_layer: int = 5, training: bool = False,
                stats: Optional[dict] = None) -> Tensor:
    """GNN.forward, gnn_type="GIN", JK="last", drop_ratio=0 (`molecule_gnn_model.py:160-184`) with GINConv
    (`:13-32`: out = mlp((1+eps) x + sum_{j->i} relu(x_j + BondEncoder(e)))) and the ogb Atom/BondEncoder (sum of the
    per-column embeddings).  `training`: BatchNorm batch statistics; updated running stats go into `stats[key]`."""
    def bn(prefix, v):
        rm, rv = sd[prefix + ".running_mean"], sd[prefix + ".running_var"]
        if training:
            rm, rv = rm.clone(), rv.clone()
            out = F.batch_norm(v, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], True, 0.1, 1e-5)
            if stats is not None:
                stats[prefix + ".running_mean"], stats[prefix + ".running_var"] = rm, rv
            return out
        return F.batch_norm(v, rm, rv, sd[prefix + ".weight"], sd[prefix + ".bias"], False, 0.1, 1e-5)

    h = 0
    for i in range(x.shape[1]):
        h = h + sd[f"atom_encoder.atom_embedding_list.{i}.weight"][x[:, i]]
    n = x.size(0)
    row, col = edge_index
    for l in range(num_layer):
        p = f"gnns.{l}"
        e = 0
        for i in range(edge_attr.shape[1]):
            e = e + sd[f"{p}.bond_encoder.bond_embedding_list.{i}.weight"][edge_attr[:, i]]
        agg = R.propagate(edge_index, F.relu(h[row] + e), n, "add")  # message x_j + edge_attr, aggr add (:28-29)
        z = (1 + sd[p + ".eps"]) * h + agg
        z = _lin(sd, p + ".mlp.0", z)
        z = F.relu(bn(p + ".mlp.1", z))
        z = _lin(sd, p + ".mlp.3", z)
        z = bn(f"batch_norms.{l}", z)
        h = z if l == num_layer - 1 else F.relu(z)  # :176-180 (dropout p = 0)
    return h


# ----------------------------------------------------------------------------
# one pretraining iteration: examples/pretrain_MoleculeSDE.py:124-152
# ----------------------------------------------------------------------------
def pretrain_losses(sds, sde_kind: str, batch, draws: dict, T: float = 0.1, anneal_power: float = 0.0):
    """Forward of the reference `train()` body with injected draws.  `sds`: {"gnn","schnet","sde2d3d","sde3d2d"} state
    dicts (tensors may require grad); `draws`: {"cl": (perm1, perm2), "sde2d3d": {noise, time_step, dropout}, "sde3d2d":
    [randint, randn_adj, randn_x]}.  Returns dict(loss, cl_loss, loss_2d3d, loss_x, loss_adj, h2d, h3d)."""
    h2d = gin_forward(sds["gnn"], batch.x, batch.edge_index, batch.edge_attr, training=True)
    _, h3d, _ = schnet_forward(sds["schnet"], batch.x[:, 0], batch.positions, batch.batch, batch.num_graphs)
    cl, _ = dual_cl(h2d, h3d, T, draws["cl"][0], draws["cl"][1])
    d23 = draws["sde2d3d"]
    l23 = loss_2d3d(sds["sde2d3d"], make_sde(sde_kind, 0.2, 1.0, 1000), h2d, batch.extended_edge_index, batch.positions,
                    batch.batch, batch.num_graphs, d23["noise"], d23["time_step"], 1000, anneal_power, d23["dropout"], True)
    bmin = 0.1 if sde_kind == "VE" else 0.2
    d32 = draws["sde3d2d"]
    lx, la = loss_3d2d(sds["sde3d2d"], make_dense_sde(sde_kind, bmin, 1.0, 1000), make_dense_sde(sde_kind, bmin, 1.0, 1000), h3d,
                       batch.x[:, 0], batch.edge_index, batch.edge_attr[:, 0], batch.batch, d32[0], d32[1], d32[2], 1000, 119,
                       anneal_power)
    loss = cl + l23 + (lx + la) * 0.5
    return {"loss": loss, "cl_loss": cl, "loss_2d3d": l23, "loss_x": lx, "loss_adj": la, "h2d": h2d, "h3d": h3d}
