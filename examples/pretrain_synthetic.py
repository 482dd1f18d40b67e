#!/usr/bin/env python
"""The reference's pretraining driver (`examples/pretrain_MoleculeSDE.py:105-180,320-350`) on the molsde_b200 kernels, with
synthetic PCQM4Mv2-shaped molecules in place of the dataset (no network / RDKit in this image).

Same structure: build GIN / SchNet / SDEModel2Dto3D_02 / SDEModel3Dto2D_node_adj_dense, loop over epochs and batches, print the
averaged losses per epoch in the reference's format, save `model_complete.pth` with the reference's checkpoint keys
(`model_2D`, `model_3D`, `SDE_2Dto3D_model`, `SDE_3Dto2D_model`), whose state_dicts load into the reference classes.

    python examples/pretrain_synthetic.py --epochs 3 --num_molecules 2048 --batch_size 256
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 examples/pretrain_synthetic.py ...     # data-parallel
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from moleculesde_b200 import checkpoint  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules  # noqa: E402
from moleculesde_b200.loader import DeviceLoader, pin_batch  # noqa: E402
from moleculesde_b200.pretrain import PretrainStep  # noqa: E402


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--epochs", type=int, default=3)
    p.add_argument("--batch_size", type=int, default=256)
    p.add_argument("--num_molecules", type=int, default=2048)
    p.add_argument("--lr", type=float, default=1e-4)
    p.add_argument("--decay", type=float, default=0.0)
    p.add_argument("--emb_dim", type=int, default=300)
    p.add_argument("--num_layer", type=int, default=5)
    p.add_argument("--T", type=float, default=0.1)
    p.add_argument("--SDE_type_2Dto3D", default="VE", choices=["VE", "VP", "VE02", "VP02", "VE03", "VP03"])
    p.add_argument("--SDE_type_3Dto2D", default="VE", choices=["VE", "VP", "VE02", "VP02", "VE03", "VP03"])
    p.add_argument("--SDE_2Dto3D_model", default="SDEModel2Dto3D_02", choices=["SDEModel2Dto3D_01", "SDEModel2Dto3D_02"])
    p.add_argument("--SDE_coeff_contrastive", type=float, default=1.0)
    p.add_argument("--SDE_coeff_generative_2Dto3D", type=float, default=1.0)
    p.add_argument("--SDE_coeff_generative_3Dto2D", type=float, default=1.0)
    p.add_argument("--SDE_anneal_power", type=float, default=0.0)
    p.add_argument("--gnn_2d_lr_scale", type=float, default=1.0)
    p.add_argument("--gnn_3d_lr_scale", type=float, default=1.0)
    p.add_argument("--output_model_dir", default="")
    return p.parse_args()


def main():
    args = parse()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(args.seed)
    # the module set of pretrain_MoleculeSDE.py:181-315 (schedule presets VE/VP/VE02/VP02/VE03/VP03, _01 / _02 variants)
    models = checkpoint.build_models(emb_dim=args.emb_dim, SDE_type_2Dto3D=args.SDE_type_2Dto3D, SDE_type_3Dto2D=args.SDE_type_3Dto2D,
                                     SDE_2Dto3D_model=args.SDE_2Dto3D_model, num_layer=args.num_layer)
    molecule_model_2D, molecule_model_3D = models["model_2D"], models["model_3D"]
    SDE_2Dto3D_model, SDE_3Dto2D_model = models["SDE_2Dto3D_model"], models["SDE_3Dto2D_model"]
    step = PretrainStep(molecule_model_2D, molecule_model_3D, SDE_2Dto3D_model, SDE_3Dto2D_model, dev, lr=args.lr, T=args.T,
                        coeff_contrastive=args.SDE_coeff_contrastive, coeff_2Dto3D=args.SDE_coeff_generative_2Dto3D,
                        coeff_3Dto2D=args.SDE_coeff_generative_3Dto2D, anneal_power=args.SDE_anneal_power,
                        gnn_2d_lr_scale=args.gnn_2d_lr_scale, gnn_3d_lr_scale=args.gnn_3d_lr_scale, weight_decay=args.decay)
    if world > 1:
        dist.broadcast(step.store.flat, src=0)
    # "dataset": every rank its own shard of synthetic molecules, collated once into pinned HOST batches (the reference
    # re-collates every epoch); the DeviceLoader copies each batch and builds its graph structures (extended graph of
    # dataset_3D.py:114-115, CSRs, radius graph, ...) on the GPU one batch ahead of the step, every epoch
    mols = synth_molecules(args.num_molecules // world, args.seed + 1000 * rank, "pcqm")
    host_batches = [pin_batch(Batch.from_data_list(mols[i:i + args.batch_size]))
                    for i in range(0, len(mols) - args.batch_size + 1, args.batch_size)]
    loader = DeviceLoader(host_batches, dev, prepare=step.prepare)
    optimal_loss = 1e10
    for epoch in range(1, args.epochs + 1):
        if rank == 0:
            print("epoch: {}".format(epoch))
        start_time = time.time()
        acc = {"cl_loss": 0.0, "cl_acc": 0.0, "loss_2d3d": 0.0, "loss_3d2d": 0.0}
        for batch in loader:
            out = step.step(batch)
            acc["cl_loss"] += float(out["cl_loss"]) if "cl_loss" in out else 0.0
            if "cl_acc_pair" in out:
                acc["cl_acc"] += 0.5 * (float(out["cl_acc_pair"][0][1]) + float(out["cl_acc_pair"][1][1]))
            acc["loss_2d3d"] += float(out["loss_2d3d"]) if "loss_2d3d" in out else 0.0
            acc["loss_3d2d"] += 0.5 * (float(out["loss_x"]) + float(out["loss_adj"])) if "loss_x" in out else 0.0
        n = max(len(loader), 1)
        temp_loss = (args.SDE_coeff_contrastive * acc["cl_loss"] + args.SDE_coeff_generative_2Dto3D * acc["loss_2d3d"] +
                     args.SDE_coeff_generative_3Dto2D * acc["loss_3d2d"]) / n
        if rank == 0:
            if temp_loss < optimal_loss:
                optimal_loss = temp_loss
                if args.output_model_dir:
                    print("save model with loss: {:.5f}".format(optimal_loss))
                    checkpoint.save_model(models, args.output_model_dir, save_best=True)   # model_complete.pth, reference keys
            print("CL Loss: {:.5f}\tCL Acc: {:.5f}\t\tSDE 2Dto3D Loss: {:.5f}\tSDE 3Dto2D Loss: {:.5f}".format(
                acc["cl_loss"] / n, acc["cl_acc"] / n, acc["loss_2d3d"] / n, acc["loss_3d2d"] / n))
            dt = time.time() - start_time
            print("Time: {:.5f}  ({:.0f} molecules/s)\n".format(dt, world * n * args.batch_size / dt))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
