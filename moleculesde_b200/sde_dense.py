"""VE / VP schedules for dense per-graph states (`Geom3D/models/MoleculeSDE/SDE_dense.py`): time `t`
has shape [B] and broadcasts as `[:, None, None]` over `[B,Nm,*]` tensors.  Same interface as the
reference, including `marGINal_prob` and `reverse(score_fn).discretize(x, adj, flags, t,
representation, SDE_model, is_adj)`.  Schedule bookkeeping only (CPU tables, a few [B]-sized torch
ops); the heavy work is in the score-network kernels.
"""
from __future__ import annotations

import numpy as np
import torch

from . import sde as _sparse


class _DenseMixin:
    def reverse(self, score_fn, probability_flow: bool = False):
        """`SDE_dense.py:64-107`."""
        N, T, fwd, disc = self.N, self.T, self.sde, self.discretize

        class RSDE(self.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, adj, flags, t, representation, SDE_model, is_adj):
                perturbed_x = SDE_model.embed(representation, x)
                drift, diffusion = fwd(adj, t) if is_adj else fwd(x, t)
                score = score_fn(perturbed_x, adj, flags, t)
                drift = drift - diffusion[:, None, None] ** 2 * score * (0.5 if self.probability_flow else 1.0)
                return drift, (0.0 if self.probability_flow else diffusion)

            def discretize(self, x, adj, flags, t, representation, SDE_model, is_adj):
                perturbed_x = SDE_model.embed(representation, x)
                f, G = disc(adj, t) if is_adj else disc(x, t)
                score = score_fn(perturbed_x, adj, flags, t)
                rev_f = f - G[:, None, None] ** 2 * score * (0.5 if self.probability_flow else 1.0)
                rev_G = torch.zeros_like(G) if self.probability_flow else G
                return rev_f, rev_G

        return RSDE()


class VPSDE(_DenseMixin, _sparse.VPSDE):
    """`SDE_dense.py:110-173`."""

    def marGINal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(log_mean_coeff[:, None, None]) * x, torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))

    def mean_coeff(self, t):
        return torch.exp(-0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0)

    def discretize(self, x, t):
        ts = self._timestep(t)
        beta = self._on("discrete_betas", x.device)[ts]
        alpha = self._on("alphas", x.device)[ts]
        return torch.sqrt(alpha)[:, None, None] * x - x, torch.sqrt(beta)


class VESDE(_DenseMixin, _sparse.VESDE):
    """`SDE_dense.py:176-233`."""

    def mean_coeff(self, t):
        return torch.ones_like(t)

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)


class subVPSDE(VPSDE):
    """`SDE_dense.py:236-277` (schedule only)."""

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1.0 - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t * discount)

    def discretize(self, x, t):
        return _sparse.SDE.discretize(self, x, t)
