#!/usr/bin/env python
"""Energy + force fine-tuning of SchNet with the loop body of the reference's `finetune_MD17.py:47-77`, unchanged, on synthetic
molecules: `positions.requires_grad_()`, `pred_force = -grad(pred_energy, positions, create_graph=True, retain_graph=True)`,
`loss = 0.05 L1(E, E*) + 0.95 L1(F, F*)`, `loss.backward()`, `optimizer.step()` -- the force term is differentiated through the kernels
(double backward: forward-mode tangent on the kernel tape).

    python examples/finetune_force_synthetic.py --epochs 3
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
from torch.autograd import grad  # noqa: E402

from moleculesde_b200.data import synth_batch  # noqa: E402
from moleculesde_b200.schnet import SchNet  # noqa: E402


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--device", default="cuda:0")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--batch_size", type=int, default=32)
    p.add_argument("--num_batches", type=int, default=4)
    p.add_argument("--epochs", type=int, default=3)
    p.add_argument("--lr", type=float, default=5e-4)
    p.add_argument("--MD17_energy_coeff", type=float, default=0.05)     # config.py:35
    p.add_argument("--MD17_force_coeff", type=float, default=0.95)      # config.py:36
    args = p.parse_args()
    torch.manual_seed(args.seed)
    device = torch.device(args.device)
    model = SchNet(hidden_channels=128, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="add", node_class=119)
    model = model.to(device)
    graph_pred_linear = torch.nn.Linear(128, 1).to(device)
    optimizer = torch.optim.Adam(list(model.parameters()) + list(graph_pred_linear.parameters()), lr=args.lr)
    criterion = torch.nn.L1Loss()
    g = torch.Generator().manual_seed(args.seed)
    loader = []
    for i in range(args.num_batches):   # synthetic targets: a fixed random "teacher" energy / force per batch
        b = synth_batch(args.batch_size, 100 + i, "pcqm")
        b.y = torch.randn(args.batch_size, generator=g) * 0.1
        b.force = torch.randn(b.positions.size(0), 3, generator=g) * 0.05
        loader.append(b)
    for epoch in range(1, args.epochs + 1):
        model.train()
        loss_acc, t0 = 0.0, time.time()
        for batch_data in loader:
            batch_data = batch_data.to(device)
            positions = batch_data.positions
            positions.requires_grad_()
            molecule_3D_repr = model(batch_data.x[:, 0].contiguous(), positions, batch_data.batch)
            pred_energy = graph_pred_linear(molecule_3D_repr).squeeze(1)
            pred_force = -grad(outputs=pred_energy, inputs=positions, grad_outputs=torch.ones_like(pred_energy), create_graph=True,
                               retain_graph=True)[0]
            loss = args.MD17_energy_coeff * criterion(pred_energy, batch_data.y.to(device)) + \
                args.MD17_force_coeff * criterion(pred_force, batch_data.force.to(device))
            optimizer.zero_grad()
            loss.backward()
            optimizer.step()
            loss_acc += loss.detach().cpu().item()
        print(f"Epoch: {epoch}\tLoss: {loss_acc / len(loader):.5f}\tTime: {time.time() - t0:.3f}")


if __name__ == "__main__":
    main()
