"""Golden fixture for the SchNet energy + FORCE training loss (double backward), from the UNMODIFIED reference
SchNet over `oracle/shims` (build container only):

    python tests/golden/make_golden_force.py   ->  golden_force.pt

The loop body of `examples/finetune_MD17.py:47-77`: positions.requires_grad_(), pred_energy = graph_pred_linear(SchNet(x, pos,
batch)).squeeze(1), pred_force = -grad(pred_energy, pos, ones, create_graph=True, retain_graph=True), loss = 0.05 L1(E, E*) +
0.95 L1(F, F*) (config.py:35-36, :288), loss.backward().  Recorded: energies, forces, loss, every SchNet parameter gradient
(norm / sum / strided sample) and the gradients of the output layer -- for the L1 criterion of the script and for an MSE criterion
(smooth: no sign() in the chain).  Same molecules (seed 0) and name-keyed weights (seed 1) as `make_golden.py`.
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import torch  # noqa: E402

from make_golden import DATA_SEED, NOISE_SEED, NUM_MOLS, WEIGHT_SEED  # noqa: E402
from make_golden_grads import summarize  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules  # noqa: E402
from oracle import refload  # noqa: E402
from oracle.weights import fill_state_dict  # noqa: E402


def main():
    R = refload.load()
    torch.set_num_threads(1)
    batch = Batch.from_data_list(synth_molecules(NUM_MOLS, DATA_SEED))
    N = batch.positions.size(0)
    g = torch.Generator().manual_seed(NOISE_SEED + 70)
    lin_w = torch.randn(1, 300, generator=g) / 17.0
    lin_b = torch.randn(1, generator=g) * 0.1
    e_true = torch.randn(NUM_MOLS, generator=g) * 0.05 + 0.3
    f_true = torch.randn(N, 3, generator=g) * 0.05
    out = {"meta": {"num_mols": NUM_MOLS, "data_seed": DATA_SEED, "weight_seed": WEIGHT_SEED, "torch": str(torch.__version__)},
           "lin_w": lin_w, "lin_b": lin_b, "e_true": e_true, "f_true": f_true}
    for crit_name, crit in (("l1", torch.nn.L1Loss()), ("mse", torch.nn.MSELoss())):
        sch = R.SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean",
                       node_class=119)
        sch.load_state_dict(fill_state_dict(sch.state_dict(), WEIGHT_SEED))
        sch.train()
        lin = torch.nn.Linear(300, 1)
        with torch.no_grad():
            lin.weight.copy_(lin_w); lin.bias.copy_(lin_b)
        pos = batch.positions.clone().requires_grad_(True)
        rep = sch(batch.x[:, 0], pos, batch.batch)
        e = lin(rep).squeeze(1)
        f = -torch.autograd.grad(outputs=e, inputs=pos, grad_outputs=torch.ones_like(e), create_graph=True, retain_graph=True)[0]
        loss = 0.05 * crit(e, e_true) + 0.95 * crit(f, f_true)
        loss.backward()
        out[crit_name] = {"energy": e.detach(), "force": f.detach(), "loss": loss.detach(),
                          "grads": {n: summarize(p.grad) for n, p in sch.named_parameters() if p.grad is not None},
                          "lin_w_grad": lin.weight.grad.clone(), "lin_b_grad": lin.bias.grad.clone()}
        print(crit_name, "loss", float(loss), "|F|", float(f.norm()), "|d lin.w|", float(lin.weight.grad.norm()))
    path = os.path.join(HERE, "golden_force.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
