"""2D->3D predictor-corrector conformer sampler with the reference's interface
(`examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-212`).

`position_PC_generation` runs the WHOLE reverse-SDE loop in one persistent kernel launch
(`molsde_sde2d3d_pc_sample`): one CTA per sampling group, positions and node features resident in
shared memory for all steps.  `ReverseDiffusionPredictor` / `LangevinCorrector` keep the reference's
step-wise `update_fn` API on top of `get_score` for callers that drive the loop themselves.
"""
from __future__ import annotations

import abc
import ctypes
from typing import Optional, Tuple

import torch

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr
from .sde import VPSDE, subVPSDE


class Predictor(abc.ABC):
    def __init__(self, sde, score_fn, probability_flow=False):
        self.sde = sde
        self.rsde = sde.reverse(score_fn, probability_flow)
        self.score_fn = score_fn

    @abc.abstractmethod
    def update_fn(self, representation, data, pos, t, args=None):
        ...


class ReverseDiffusionPredictor(Predictor):
    """`:156-168`"""

    def update_fn(self, representation, data, pos, t, args=None, noise=None):
        f, G = self.rsde.discretize(pos, representation, data, t)
        noise = torch.randn_like(pos) if noise is None else noise
        x_mean = pos - f
        return x_mean + G[:, None] * noise, x_mean


class Corrector(abc.ABC):
    def __init__(self, sde, score_fn, snr, scale_eps, n_steps):
        self.sde, self.score_fn = sde, score_fn
        self.snr, self.scale_eps, self.n_steps = snr, scale_eps, n_steps

    @abc.abstractmethod
    def update_fn(self, representation, data, pos, t, args=None):
        ...


class LangevinCorrector(Corrector):
    """`:187-212`.  The reference's inner loop never feeds `x` back into `pos` (F9): every
    iteration evaluates the same score and draws a fresh noise, only the last draw survives -- so
    the score is evaluated once and `n_steps - 1` draws are consumed and dropped."""

    def update_fn(self, representation, data, pos, t, args=None, noise=None):
        sde = self.sde
        if isinstance(sde, (VPSDE, subVPSDE)):
            timestep = (t * (sde.N - 1) / sde.T).long()
            alpha = sde.alphas.to(t.device)[timestep]
        else:
            alpha = torch.ones_like(t)
        grad = self.score_fn.get_score(representation, data, pos, None, t)
        if noise is None:
            for _ in range(self.n_steps):
                noise = torch.randn_like(pos)
        grad_norm = torch.norm(grad.reshape(grad.shape[0], -1), dim=-1).mean()
        noise_norm = torch.norm(noise.reshape(noise.shape[0], -1), dim=-1).mean()
        step_size = (self.snr * noise_norm / grad_norm) ** 2 * 2 * alpha
        x_mean = pos + step_size[:, None] * grad
        x = x_mean + torch.sqrt(step_size * 2)[:, None] * noise * self.scale_eps
        return x, x_mean


@torch.no_grad()
def position_PC_generation(representation, data, pos_init, scorenet, sde, probability_flow=False, denoise=True,
                           eps=1e-4, snr=0.2, scale_eps=0.9, n_steps=1, *, group_ptr: Optional[torch.Tensor] = None,
                           noise_corr: Optional[torch.Tensor] = None, noise_pred: Optional[torch.Tensor] = None,
                           seed: int = 0, diffusion_steps: Optional[int] = None):
    """Reference signature (`:92-96`) plus keyword-only extensions:

    group_ptr        int64 CPU tensor [G+1] of molecule offsets: each group is an independent
                     sampling batch with its own Langevin step size (the reference calls this function
                     once per group of `num_repeat` conformers; here many groups share one launch).
                     Default: the whole batch is one group, exactly the reference semantics.
    noise_corr/pred  float32 [steps,N,3] injected draws (parity mode); default in-kernel Philox(seed).
    diffusion_steps  run only the first k reverse steps of the `linspace(T, eps, N)` grid.
    Returns (data, pos_mean) if `denoise` else (data, pos), like the reference (`:135-138`).
    """
    if probability_flow:
        raise NotImplementedError("probability-flow ODE sampling is not used by the reference scripts")
    require_device(pos_init)
    dev = pos_init.device
    N_diff = scorenet.sde_pos.N
    timesteps = torch.linspace(scorenet.sde_pos.T, eps, N_diff)  # :118 (CPU table; same values for every atom)
    steps = N_diff if diffusion_steps is None else int(diffusion_steps)
    table = sde.step_table(timesteps[:steps]).to(dev).contiguous()
    if group_ptr is None:
        group_ptr = torch.tensor([0, data.num_graphs], dtype=torch.long)
    prep = scorenet.prepared(data, group_ptr)
    pk = scorenet.packed_params()
    nattr, e2d = scorenet.invariants(representation, prep)
    pos0 = pos_init.detach().float().contiguous()
    n_atoms = pos0.size(0)
    if (noise_corr is None) != (noise_pred is None):
        raise ValueError("give both noise_corr and noise_pred or neither")
    if noise_corr is not None:
        noise_corr = noise_corr.float().contiguous()
        noise_pred = noise_pred.float().contiguous()
        assert tuple(noise_corr.shape) == (steps, n_atoms, 3) and tuple(noise_pred.shape) == (steps, n_atoms, 3)
    pos_out = torch.full_like(pos0, float("nan"))   # a skipped chunk stays NaN
    pos_mean = torch.full_like(pos0, float("nan"))
    scratch = prep.get_scratch()
    st = prep.plan.as_struct()
    prm = _abi.Params(pk["blob"].data_ptr(), pk["blob"].numel())
    cfg = _abi.PCConfig(steps, float(snr), float(scale_eps), int(seed) & 0xFFFFFFFFFFFFFFFF)
    prep.status.zero_()
    check(lib().molsde_sde2d3d_pc_sample(ctypes.byref(st), ctypes.byref(prm), ptr(nattr), ptr(e2d), ptr(pos0), ptr(table),
                                         ctypes.byref(cfg), ptr(noise_corr), ptr(noise_pred), ptr(pos_out), ptr(pos_mean),
                                         ptr(scratch), scratch.numel(), ptr(prep.counter), ptr(prep.status),
                                         stream_ptr(pos0)), "sde2d3d_pc_sample")
    # device status word (one sync; the reference's driver copies the positions to the host right after this call anyway):
    # > 0 = chunk (1 + index) exceeded a compiled limit and was skipped, -7 = a tensor-core / TMA completion wait timed out
    status = int(prep.status.item())
    if status != 0:
        raise _abi.MolsdeError(f"sde2d3d_pc_sample: device status {status} "
                               f"({'chunk %d beyond the compiled limits' % (status - 1) if status > 0 else 'completion wait timed out'}); "
                               "the returned positions would be undefined")
    return (data, pos_mean) if denoise else (data, pos_out)
