"""Golden fixture for the PRETRAINING STEP (SURVEY.md section 8, config 1 shape at 8 molecules): one
`pretrain_MoleculeSDE.py:105-152` iteration of the UNMODIFIED reference sources (imported over `oracle/shims`)
in train mode — GIN -> SchNet -> dual_CL -> SDEModel2Dto3D_02 -> SDEModel3Dto2D_node_adj_dense -> backward -> Adam —
with every random draw recorded.  Stored: the draws, the losses, both representations, dL/d(representation), and for
every parameter the gradient's norm, sum and a strided sample (full tensor when <= 512 elements), plus the same
sample of the parameter after one Adam step (lr 1e-4).

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden_grads.py                 # golden_grads.pt   (8 molecules, VE + VP, full representations)
    python tests/golden/make_golden_grads.py --b32           # golden_grads_b32.pt: BASELINE.json configs[0] -- batch 32, VE/VE,
                                                             # data seed 32; representations stored as summaries only
"""
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from make_golden import DATA_SEED, NOISE_SEED, NUM_MOLS, WEIGHT_SEED, build_reference_models, record_rng  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules  # noqa: E402
from oracle import refload  # noqa: E402

SAMPLE = 512


def summarize(t, n=SAMPLE):
    f = t.detach().reshape(-1).float()
    stride = max(1, f.numel() // n)
    return {"norm": f.double().norm().float(), "sum": f.double().sum().float(), "stride": stride,
            "sample": f[::stride][:n].clone(), "numel": f.numel()}


def main(num_mols=NUM_MOLS, data_seed=DATA_SEED, kinds=("VE", "VP"), out_name="golden_grads.pt", full_repr=True):
    R = refload.load()
    torch.set_num_threads(1)
    mols = synth_molecules(num_mols, data_seed)
    for m in mols:
        R.extend_graph(m)
    batch = Batch.from_data_list(mols)
    out = {"meta": {"num_mols": num_mols, "data_seed": data_seed, "weight_seed": WEIGHT_SEED, "noise_seed": NOISE_SEED + 20,
                    "lr": 1e-4, "torch": str(torch.__version__), "num_atoms": int(batch.positions.size(0)),
                    "num_ext_edges": int(batch.extended_edge_index.size(1))}}
    for sde_type in kinds:
        gnn, sch, m23, m32 = build_reference_models(R, sde_type)
        models = {"gnn": gnn, "schnet": sch, "sde2d3d": m23, "sde3d2d": m32}
        for m in models.values():
            m.train()
        opt = torch.optim.Adam([{"params": m.parameters(), "lr": 1e-4} for m in models.values()], lr=1e-4, weight_decay=0)
        args = types.SimpleNamespace(CL_similarity_metric="EBM_node_dot_prod", T=0.1)
        log = []
        with record_rng(log, NOISE_SEED + 20):
            h2d = gnn(batch.x, batch.edge_index, batch.edge_attr)
            _, h3d = sch(batch.x[:, 0], batch.positions, batch.batch, return_latent=True)
            h2d.retain_grad(); h3d.retain_grad()
            cl_loss, cl_acc = R.examples.dual_CL(h2d, h3d, args)
            b2 = batch.clone()
            l23 = m23(h2d, b2, anneal_power=0)["position"]
            lx, la = m32(h3d, batch, reduce_mean=True, continuous=True, train=True, anneal_power=0)
            loss = cl_loss + l23 + (lx + la) * 0.5
        opt.zero_grad()
        loss.backward()
        keep = (lambda t: t.detach().clone()) if full_repr else summarize
        sec = {"draws": [(k, v) for k, v in log], "h2d": keep(h2d), "h3d": keep(h3d),
               "d_h2d": keep(h2d.grad), "d_h3d": keep(h3d.grad),
               "loss": loss.detach(), "cl_loss": cl_loss.detach(), "cl_acc": torch.tensor(cl_acc), "loss_2d3d": l23.detach(),
               "loss_x": lx.detach(), "loss_adj": la.detach(), "grads": {}, "after_step": {}, "buffers": {}}
        for name, m in models.items():
            sec["grads"][name] = {n: (summarize(p.grad) if p.grad is not None else None)
                                  for n, p in m.named_parameters() if p.requires_grad}
        opt.step()
        for name, m in models.items():
            sec["after_step"][name] = {n: summarize(p, 64) for n, p in m.named_parameters() if p.requires_grad}
            sec["buffers"][name] = {n: summarize(b) for n, b in m.named_buffers() if "running" in n}
        out["pretrain_" + sde_type] = sec
        print(sde_type, "loss", float(loss), "cl", float(cl_loss), "2d3d", float(l23), "x", float(lx), "adj", float(la),
              "draws", [k for k, _ in log])
    path = os.path.join(HERE, out_name)
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    if "--b32" in sys.argv:
        main(num_mols=32, data_seed=32, kinds=("VE",), out_name="golden_grads_b32.pt", full_repr=False)
    else:
        main()
