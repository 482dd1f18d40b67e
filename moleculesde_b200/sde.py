"""VE / VP SDE schedules with the reference's interface (`Geom3D/models/MoleculeSDE/SDE_sparse.py`).

Plain Python objects holding CPU tensors, like the reference (not `nn.Module`, not serialised).
Method names follow the reference, including the `marGINal_prob` spelling.  These are
schedule bookkeeping, not hot-path compute: the fused PC kernel receives them as a per-step table
(`step_table`), and `reverse(...).discretize` is kept for callers that drive the loop themselves.
Per-node time `t` has shape [N] ("sparse" flavour).
"""
from __future__ import annotations

import abc

import numpy as np
import torch


class SDE(abc.ABC):
    def __init__(self, N: int):
        self.N = N

    @property
    @abc.abstractmethod
    def T(self):
        ...

    @abc.abstractmethod
    def sde(self, x, t):
        ...

    @abc.abstractmethod
    def marGINal_prob(self, x, t):
        ...

    def prior_sampling(self, shape):
        return torch.randn(*shape)

    def prior_sampling_sym(self, shape):
        x = torch.randn(*shape).triu(1)
        return x + x.transpose(-1, -2)

    @abc.abstractmethod
    def prior_logp(self, z):
        ...

    def discretize(self, x, t):
        # Euler-Maruyama default (SDE_sparse.py:48-62)
        dt = 1 / self.N
        drift, diffusion = self.sde(x, t)
        return drift * dt, diffusion * torch.sqrt(torch.tensor(dt, device=t.device))

    def reverse(self, score_fn, probability_flow: bool = False):
        """Reverse-time SDE (SDE_sparse.py:64-102): `score_fn` is a model exposing `get_score`."""
        N, T, fwd, disc = self.N, self.T, self.sde, self.discretize

        class RSDE(self.__class__):
            def __init__(self):
                self.N = N
                self.probability_flow = probability_flow

            @property
            def T(self):
                return T

            def sde(self, x, representation, data, t):
                drift, diffusion = fwd(x, t)
                score = score_fn.get_score(representation, data, x, None, t)
                drift = drift - diffusion[:, None] ** 2 * score * (0.5 if self.probability_flow else 1.0)
                return drift, (0.0 if self.probability_flow else diffusion)

            def discretize(self, x, representation, data, t):
                f, G = disc(x, t)
                score = score_fn.get_score(representation, data, x, None, t)
                rev_f = f - G[:, None] ** 2 * score * (0.5 if self.probability_flow else 1.0)
                rev_G = torch.zeros_like(G) if self.probability_flow else G
                return rev_f, rev_G

        return RSDE()

    # ---- per-step constants for the fused PC kernel (csrc/sde2d3d.cu) ----
    def step_table(self, timesteps: torch.Tensor) -> torch.Tensor:
        """float32 [steps, 8] = {std, G, sqrt_alpha, corrector_alpha, 0...} evaluated on the CPU with
        the very expressions the reference evaluates per step (`marGINal_prob`, `discretize`,
        LangevinCorrector's alpha)."""
        t = timesteps.detach().cpu().float()
        _, std = self.marGINal_prob(torch.zeros(t.numel(), 1), t)
        _, G = self.discretize(torch.zeros(t.numel(), 1), t)
        table = torch.zeros(t.numel(), 8, dtype=torch.float32)
        table[:, 0], table[:, 1] = std, G
        table[:, 2], table[:, 3] = self._sqrt_alpha(t), self._corrector_alpha(t)
        return table

    def _sqrt_alpha(self, t):
        return torch.ones_like(t)

    def _corrector_alpha(self, t):
        return torch.ones_like(t)


class VPSDE(SDE):
    """SDE_sparse.py:105-169."""

    def __init__(self, beta_min=0.1, beta_max=20, N=1000):
        super().__init__(N)
        self.beta_0, self.beta_1 = beta_min, beta_max
        self.discrete_betas = torch.linspace(beta_min / N, beta_max / N, N)
        self.alphas = 1.0 - self.discrete_betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.sqrt_alphas_cumprod = torch.sqrt(self.alphas_cumprod)
        self.sqrt_1m_alphas_cumprod = torch.sqrt(1.0 - self.alphas_cumprod)

    @property
    def T(self):
        return 1

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t)

    def marGINal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(log_mean_coeff[:, None]) * x, torch.sqrt(1.0 - torch.exp(2.0 * log_mean_coeff))

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi) - torch.sum(z ** 2, dim=(1, 2)) / 2.0

    def _timestep(self, t):
        return (t * (self.N - 1) / self.T).long()

    def to_device(self, device):
        """Keep device copies of the schedule tables (`_on`): needed when a sampling step is captured in a CUDA graph."""
        self._dev_tables = {n: getattr(self, n).to(device) for n in ("discrete_betas", "alphas")}
        self._dev_tables["device"] = torch.device(device)

    def _on(self, name, device):
        tabs = getattr(self, "_dev_tables", None)
        if tabs is not None and tabs["device"] == torch.device(device):
            return tabs[name]
        return getattr(self, name).to(device)

    def discretize(self, x, t):
        ts = self._timestep(t)
        beta = self._on("discrete_betas", x.device)[ts]
        alpha = self._on("alphas", x.device)[ts]
        return torch.sqrt(alpha)[:, None] * x - x, torch.sqrt(beta)

    def _sqrt_alpha(self, t):
        return torch.sqrt(self.alphas[self._timestep(t)])

    def _corrector_alpha(self, t):
        return self.alphas[self._timestep(t)]


class VESDE(SDE):
    """SDE_sparse.py:172-230."""

    def __init__(self, sigma_min=0.01, sigma_max=50, N=1000):
        super().__init__(N)
        self.sigma_min, self.sigma_max = sigma_min, sigma_max
        self.discrete_sigmas = torch.exp(torch.linspace(np.log(sigma_min), np.log(sigma_max), N))

    @property
    def T(self):
        return 1

    def to_device(self, device):
        self._dev_tables = {"discrete_sigmas": self.discrete_sigmas.to(device), "device": torch.device(device)}

    def _on(self, name, device):
        tabs = getattr(self, "_dev_tables", None)
        if tabs is not None and tabs["device"] == torch.device(device):
            return tabs[name]
        return getattr(self, name).to(device)

    def sde(self, x, t):
        sigma = self.sigma_min * (self.sigma_max / self.sigma_min) ** t
        diffusion = sigma * torch.sqrt(
            torch.tensor(2 * (np.log(self.sigma_max) - np.log(self.sigma_min)), device=t.device))
        return torch.zeros_like(x), diffusion

    def marGINal_prob(self, x, t):
        return x, self.sigma_min * (self.sigma_max / self.sigma_min) ** t

    def prior_logp(self, z):
        n = np.prod(z.shape[1:])
        return -n / 2.0 * np.log(2 * np.pi * self.sigma_max ** 2) - torch.sum(z ** 2, dim=(1, 2, 3)) / (2 * self.sigma_max ** 2)

    def discretize(self, x, t):
        # the reference indexes a CPU table with a device index (F10); here the table follows t
        ts = (t * (self.N - 1) / self.T).long()
        sig = self._on("discrete_sigmas", t.device)
        sigma = sig[ts]
        adjacent = torch.where(ts == 0, torch.zeros_like(t), sig[ts - 1])
        return torch.zeros_like(x), torch.sqrt(sigma ** 2 - adjacent ** 2)


class subVPSDE(VPSDE):
    """SDE_sparse.py:233-275 (schedule only; not used by the BASELINE configs)."""

    def sde(self, x, t):
        beta_t = self.beta_0 + t * (self.beta_1 - self.beta_0)
        discount = 1.0 - torch.exp(-2 * self.beta_0 * t - (self.beta_1 - self.beta_0) * t ** 2)
        return -0.5 * beta_t[:, None, None] * x, torch.sqrt(beta_t * discount)

    def marGINal_prob(self, x, t):
        log_mean_coeff = -0.25 * t ** 2 * (self.beta_1 - self.beta_0) - 0.5 * t * self.beta_0
        return torch.exp(log_mean_coeff)[:, None, None] * x, 1 - torch.exp(2.0 * log_mean_coeff)

    def discretize(self, x, t):
        return SDE.discretize(self, x, t)
