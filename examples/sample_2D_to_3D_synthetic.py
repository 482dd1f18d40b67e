#!/usr/bin/env python
"""Conformer generation with the flow of the reference's `pretrain_MoleculeSDE_inference_2D_to_3D_VE_VP.py` on synthetic
PCQM4Mv2-shaped molecules: load a `model_complete.pth` (or use seeded random weights), encode with the 2D GNN, run the 1000-step
VE/VP predictor-corrector sampler for `num_repeat_SDE_inference` conformers per molecule, report conformers/s.

    python examples/sample_2D_to_3D_synthetic.py --num_molecules 296 --input_model_file out/model_complete.pth
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from moleculesde_b200 import checkpoint  # noqa: E402
from moleculesde_b200.data import synth_molecules  # noqa: E402
from moleculesde_b200.inference import generate_samples_from_testset  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--num_molecules", type=int, default=296)
    p.add_argument("--start", type=int, default=0)
    p.add_argument("--end", type=int, default=10 ** 9)
    p.add_argument("--num_repeat_SDE_inference", type=int, default=10)     # config.py:133
    p.add_argument("--steps_pos", type=int, default=100)                    # config.py:136
    p.add_argument("--SDE_type_2Dto3D", default="VE")
    p.add_argument("--SDE_2Dto3D_model", default=None, help="SDEModel2Dto3D_01 / _02 (default: from the checkpoint, else _02)")
    p.add_argument("--input_model_file", default="")
    p.add_argument("--output_dir", default="")
    p.add_argument("--molecules_per_launch", type=int, default=592, help="sampling groups per kernel launch (148 SMs x 4)")
    p.add_argument("--diffusion_steps", type=int, default=None, help="run only the first k reverse steps (smoke runs)")
    p.add_argument("--generator", default="MoleculeSDE")
    p.add_argument("--eval_epoch", type=int, default=0)
    p.add_argument("--min_sigma", type=float, default=0.0)
    return p.parse_args()


def main():
    args = parse()
    torch.manual_seed(args.seed)
    dev = torch.device(args.device)
    kw = {"SDE_type_2Dto3D": args.SDE_type_2Dto3D}
    if args.SDE_2Dto3D_model:
        kw["SDE_2Dto3D_model"] = args.SDE_2Dto3D_model
    models = checkpoint.load_model(args.input_model_file, **kw) if args.input_model_file else checkpoint.build_models(**kw)
    gnn, m23 = models["model_2D"].to(dev), models["SDE_2Dto3D_model"].to(dev)
    test_set = synth_molecules(args.num_molecules, args.seed, "pcqm")
    # warm-up (CUDA module loading, parameter packing): one reverse step on the first molecules, not timed
    warm = argparse.Namespace(**{**vars(args), "start": 0, "end": min(4, args.num_molecules)})
    generate_samples_from_testset(test_set, gnn, m23, warm, diffusion_steps=1)
    torch.cuda.synchronize()
    t0 = time.time()
    out = generate_samples_from_testset(test_set, gnn, m23, args, out_path=args.output_dir or None,
                                        molecules_per_launch=args.molecules_per_launch, seed=args.seed,
                                        diffusion_steps=args.diffusion_steps)
    torch.cuda.synchronize()
    dt = time.time() - t0
    n_conf = sum(int(d.num_pos_gen) for d in out)
    print("generated {} conformers of {} molecules in {:.2f} s ({:.1f} conformers/s), model {}".format(
        n_conf, len(out), dt, n_conf / dt, type(m23).__name__))


if __name__ == "__main__":
    main()
