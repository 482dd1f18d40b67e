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
    if not _groups_fit_fused(scorenet, data, group_ptr):
        # a sampling group beyond one CTA's shared memory (> 224 atoms / 64 edge tiles): same algorithm, one CUDA-graph replay per step
        return _position_PC_stepwise(representation, data, pos_init, scorenet, table, group_ptr, snr, scale_eps, noise_corr, noise_pred,
                                     seed, steps, denoise)
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


def _groups_fit_fused(scorenet, data, group_ptr) -> bool:
    """True when every sampling group fits the fused kernel (one CTA per group: <= CHUNK_MAX_NODES atoms, <= MAX_CHUNK_TILES tiles)."""
    from .graph import segment_ptr
    node_ptr = getattr(data, "_molsde_node_ptr_cpu", None)
    if node_ptr is None:
        node_ptr = segment_ptr(data.batch, data.num_graphs).cpu()
        try:
            data._molsde_node_ptr_cpu = node_ptr
        except AttributeError:
            pass
    sizes = node_ptr[group_ptr.long()]
    if int((sizes[1:] - sizes[:-1]).max()) > _abi.CHUNK_MAX_NODES:
        return False
    try:
        scorenet.prepared(data, group_ptr)   # builds (and caches) the one-chunk-per-group plan; raises beyond the tile limit
    except _abi.MolsdeError:
        return False
    return True


def _position_PC_stepwise(representation, data, pos_init, scorenet, table, group_ptr, snr, scale_eps, noise_corr, noise_pred, seed,
                          steps, denoise):
    """`position_PC_generation` for sampling groups of any size: the score network runs over chunks of whole MOLECULES
    (`molsde_sde2d3d_forward_net`), the Langevin step size is reduced per group by `molsde_sde2d3d_pc_corrector_update`, and
    one reverse step [score, corrector, score, predictor] is captured in a CUDA graph and replayed `steps` times (the step index
    is a device-side counter).  Same schedule table, same Philox noise streams as the fused kernel."""
    from .sde_2d_to_3d import prepare_graph
    dev = pos_init.device
    key = "_molsde_prep_ext_free" if scorenet.use_extend_graph else "_molsde_prep_bond_free"
    prep = getattr(data, key, None)
    if prep is None:
        csr = getattr(data, "_molsde_ext_csr", None) if scorenet.use_extend_graph else None
        prep = prepare_graph(scorenet._edge_index(data), data.batch, data.num_graphs, None, csr)   # molecules packed into chunks
        try:
            setattr(data, key, prep)
        except AttributeError:
            pass
    pk = scorenet.packed_params()
    nattr, e2d = scorenet.invariants(representation, prep)
    pos = pos_init.detach().float().contiguous().clone()
    n_atoms = pos.size(0)
    if (noise_corr is None) != (noise_pred is None):
        raise ValueError("give both noise_corr and noise_pred or neither")
    if noise_corr is not None:
        noise_corr, noise_pred = noise_corr.float().contiguous(), noise_pred.float().contiguous()
        assert tuple(noise_corr.shape) == (steps, n_atoms, 3) and tuple(noise_pred.shape) == (steps, n_atoms, 3)
    gnp = prep.node_ptr.cpu()[group_ptr.long()].to(torch.int32).to(dev)      # node offsets of the groups
    G = gnp.numel() - 1
    net = torch.full_like(pos, float("nan"))
    pos_mean = torch.empty_like(pos)
    scratch = prep.get_scratch()
    st = prep.plan.as_struct()
    prm = _abi.Params(pk["blob"].data_ptr(), pk["blob"].numel())
    counter = torch.zeros(1, dtype=torch.int32, device=dev)
    prep.status.zero_()
    L, sd = lib(), int(seed) & 0xFFFFFFFFFFFFFFFF

    def one_step():
        s = stream_ptr(pos)
        check(L.molsde_sde2d3d_forward_net(ctypes.byref(st), ctypes.byref(prm), ptr(nattr), ptr(e2d), ptr(pos), None, None, 0.0, ptr(net),
                                           ptr(scratch), scratch.numel(), ptr(prep.status), s), "sde2d3d_forward_net")
        check(L.molsde_sde2d3d_pc_corrector_update(ptr(net), ptr(pos), ptr(gnp), G, ptr(table), ptr(counter), float(snr), float(scale_eps),
                                                   sd, ptr(noise_corr), n_atoms, s), "pc_corrector_update")
        check(L.molsde_sde2d3d_forward_net(ctypes.byref(st), ctypes.byref(prm), ptr(nattr), ptr(e2d), ptr(pos), None, None, 0.0, ptr(net),
                                           ptr(scratch), scratch.numel(), ptr(prep.status), s), "sde2d3d_forward_net")
        check(L.molsde_sde2d3d_pc_predictor_update(ptr(net), ptr(pos), ptr(pos_mean), ptr(table), ptr(counter), sd, ptr(noise_pred),
                                                   n_atoms, s), "pc_predictor_update")

    if steps < 4:
        for _ in range(steps):
            one_step()
    else:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=side):
                one_step()
        torch.cuda.current_stream(dev).wait_stream(side)
        for _ in range(steps):
            graph.replay()
    status = int(prep.status.item())
    if status != 0:
        raise _abi.MolsdeError(f"sde2d3d_forward_net: device status {status} (a molecule beyond the compiled limits of the score kernel: "
                               f"{_abi.CHUNK_MAX_NODES} atoms / {_abi.MAX_CHUNK_TILES} tiles per chunk)")
    return (data, pos_mean) if denoise else (data, pos)
