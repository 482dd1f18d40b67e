"""Long-trajectory fixture for the per-step score agreement (BASELINE.json north_star: "per-step score agreement along a shared
sampling trajectory"): the UNMODIFIED reference `position_PC_generation`
(`examples/pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py:92-138`) runs all 1000 predictor-corrector steps on 4 conformers of the
first seeded synthetic molecule, VE and VP, and every `get_score` call of 51 reverse steps spread over t in [1, 1e-4]
(steps 0, 20, ..., 980 and 999; corrector call + predictor call each) is recorded as (pos, t, score).  The GPU test feeds the
recorded positions to the CUDA `get_score` (teacher forcing) and compares the scores.

The reference loop carries a debugging leftover, `if i >= 10: break` (`:130-132`, "TODO: hacking"), which would stop the run after
11 steps.  The sources stay unmodified: the loop's `trange` is replaced by a generator of int-subclass indices whose `>=`
comparison is always False, so the full `sde.N` steps run.

Build container only (needs /root/reference):  python tests/golden/make_golden_traj.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import DATA_SEED, NOISE_SEED, NUM_MOLS, WEIGHT_SEED, build_reference_models, record_rng  # noqa: E402
from moleculesde_b200.data import repeat_data, synth_molecules  # noqa: E402
from oracle import refload  # noqa: E402

REPEAT = 4


class _Idx(int):
    """loop index that indexes `timesteps[i]` like an int but never satisfies the reference's `i >= 10` early exit"""

    def __ge__(self, other):
        return False
KEEP_STEPS = sorted(set(list(range(0, 1000, 20)) + [999]))


def main():
    R = refload.load()
    torch.set_num_threads(4)
    mols = synth_molecules(NUM_MOLS, DATA_SEED)
    for m in mols:
        R.extend_graph(m)
    out = {"meta": {"num_mols": NUM_MOLS, "data_seed": DATA_SEED, "weight_seed": WEIGHT_SEED, "noise_seed": NOISE_SEED + 40,
                    "repeat": REPEAT, "steps": KEEP_STEPS, "torch": str(torch.__version__)}}
    for sde_type in ("VE", "VP"):
        gnn, _, m23, _ = build_reference_models(R, sde_type)
        gnn.eval(); m23.eval()
        rb = repeat_data(mols[0], REPEAT)
        with torch.no_grad():
            rep = gnn(rb.x, rb.edge_index, rb.edge_attr)
        calls, count = [], {"n": 0}
        orig = m23.get_score

        def traced(representation, data, pos, sigma, tt):
            s = orig(representation, data, pos, sigma, tt)
            step, which = divmod(count["n"], 2)      # call 0 of a step = corrector, call 1 = predictor
            if step in KEEP_STEPS:
                calls.append((step, which, pos.clone(), tt.clone(), s.clone()))
            count["n"] += 1
            return s

        m23.get_score = traced
        log = []
        glb = R.examples.pc2d3d_globals
        o_trange = glb["trange"]
        glb["trange"] = lambda a, b, **kw: (_Idx(i) for i in range(a, b))
        try:
            with record_rng(log, NOISE_SEED + 40), torch.no_grad():
                pos_init = m23.sde_pos.prior_sampling(rb.positions.shape)
                _, pos_mean = R.examples.pc2d3d.position_PC_generation(
                    representation=rep, data=rb, pos_init=pos_init, scorenet=m23, sde=m23.sde_pos, n_steps=1)
        finally:
            glb["trange"] = o_trange
        del m23.get_score
        assert count["n"] == 2000, count
        out["sde2d3d_" + sde_type] = {"representation": rep, "calls": calls, "pos_mean": pos_mean, "pos_init": pos_init}
        print(sde_type, "calls kept", len(calls), "final |pos_mean| max", float(pos_mean.abs().max()))
    path = os.path.join(HERE, "golden_traj.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
