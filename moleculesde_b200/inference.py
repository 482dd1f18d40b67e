"""2D->3D conformer generation driver with the reference's interface
(`examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:28-91`: `repeat_data`, `generate_samples_from_testset`).

The reference walks the test set one molecule at a time: `num_repeat` copies of the molecule form a batch, the 2D GNN encodes it,
`position_PC_generation` runs the 1000-step predictor-corrector loop, the conformers are split back per copy.  One such batch is
~140 atoms -- a single CTA of work -- so here `molecules_per_launch` molecules are processed per launch: their repeated copies
are concatenated into ONE batch (one GIN forward, one extended-graph build, one fused PC kernel launch) and passed as independent
sampling groups (`group_ptr`), which keeps the reference's per-group Langevin step size (a mean over the copies of ONE molecule).
Returned records are the reference's: a clone of the input molecule with `pos_gen [num_repeat * n, 3]` and `num_pos_gen`.
Not replicated: the debugging `if i >= 9: break` of `:78-80` (the whole `[start, end)` range is generated).
"""
from __future__ import annotations

import os
import pickle
from typing import List, Optional, Sequence

import torch

from . import graph as G
from .data import Batch, Data
from .sampler import position_PC_generation


@torch.no_grad()
def generate_samples_from_testset(data: Sequence[Data], molecule_model_2D, SDE_2Dto3D_model, args, out_path: Optional[str] = None,
                                  molecules_per_launch: int = 148, seed: int = 0, diffusion_steps: Optional[int] = None,
                                  _injected=None) -> List[Data]:
    """`data`: the test molecules (anything indexable yielding `Data` with x / edge_index / edge_attr / positions).
    `args`: the reference's namespace -- `start`, `end`, `num_repeat_SDE_inference`, `steps_pos`, `device` (and, for the output
    file name, `generator`, `eval_epoch`, `min_sigma`).  `_injected(N) -> (pos_init, noise_corr, noise_pred)` feeds fixed draws in
    the parity tests; by default the prior is drawn like the reference (`sde_pos.prior_sampling`, CPU) and the per-step noise
    comes from the kernel's Philox generator."""
    molecule_model_2D.eval()
    SDE_2Dto3D_model.eval()
    dev = torch.device(args.device)
    n_total = len(data)
    start, end = max(0, int(args.start)), min(n_total, int(args.end))
    R = int(args.num_repeat_SDE_inference)
    all_data_list: List[Data] = []
    for lo in range(start, end, max(1, int(molecules_per_launch))):
        mols = [data[i] for i in range(lo, min(end, lo + max(1, int(molecules_per_launch))))]
        copies = []
        for m in mols:                          # repeat_data (:28-39) for every molecule of the launch
            for _ in range(R):
                d = Data.from_dict({k: m[k] for k in m.keys})
                d.pos_gen = torch.ones_like(m.positions)
                copies.append(d)
        batch = Batch.from_data_list(copies).to(dev)
        group_ptr = torch.arange(0, len(mols) * R + 1, R, dtype=torch.long)
        if SDE_2Dto3D_model.use_extend_graph and getattr(batch, "extended_edge_index", None) is None:
            # dataset_3D.py:114-115 (per sample, CPU) moved after collation, on the GPU
            csr = G.extend_graph(batch.edge_index, batch.batch, batch.num_graphs)
            batch.extended_edge_index, batch._molsde_ext_csr = csr.edge_index, csr
        representation = molecule_model_2D(batch.x, batch.edge_index, batch.edge_attr)          # :60
        kw = {}
        if _injected is None:
            pos_init = SDE_2Dto3D_model.sde_pos.prior_sampling(batch.positions.shape).to(dev)   # :62
        else:
            pos_init, nc, npd = _injected(batch.positions.size(0))
            pos_init, kw = pos_init.to(dev), {"noise_corr": nc.to(dev), "noise_pred": npd.to(dev)}
        _, pos_gen = position_PC_generation(representation=representation, data=batch, pos_init=pos_init,
                                            scorenet=SDE_2Dto3D_model, sde=SDE_2Dto3D_model.sde_pos, n_steps=args.steps_pos,
                                            group_ptr=group_ptr, seed=seed + lo, diffusion_steps=diffusion_steps, **kw)  # :63-66
        pos_gen = pos_gen.cpu()
        ptr = batch.ptr.cpu() if torch.is_tensor(getattr(batch, "ptr", None)) else None
        for k, m in enumerate(mols):
            a, b = int(ptr[k * R]), int(ptr[(k + 1) * R])
            rec = m.clone().detach()
            rec.pos_gen = pos_gen[a:b].clone()                                                   # (num_repeat * num_node, 3), :74
            rec.num_pos_gen = torch.tensor([R], dtype=torch.long)
            all_data_list.append(rec)
    if out_path is not None:
        name = "2D_to_3D_{}_{}_{}_epoch_{}_min_sig_{:.3f}_repeat_{}".format(
            getattr(args, "generator", "MoleculeSDE"), start, end, getattr(args, "eval_epoch", 0), getattr(args, "min_sigma", 0.0), R)
        os.makedirs(out_path, exist_ok=True)
        with open(os.path.join(out_path, name), "wb") as fout:
            pickle.dump(all_data_list, fout)
    return all_data_list
