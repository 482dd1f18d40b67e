"""Golden fixture for `SDEModel3Dto2D_node_adj_dense_02` (`SDE_model_3D_to_2D_node_adj_dense.py:182-350`: the
variant that CONCATENATES embedding_3D(rep) and embedding_X(x) into a 600-wide input of both score networks), from the
UNMODIFIED reference over `oracle/shims` (build container only):

    python tests/golden/make_golden_dense02.py   ->  golden_dense02.pt

Per SDE type (VE, VP): state_dict manifest; both scores on the fixed perturbed state of `golden_pcqm8.pt`; one training forward
(`train=True`) with every draw recorded, both losses, every parameter gradient (norm / sum / strided sample) and d loss / d h3d for
loss = (loss_x + loss_adj) / 2 (`pretrain_MoleculeSDE.py:146`).  Same molecules (seed 0) and name-keyed weights (seed 1).
"""
from __future__ import annotations

import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)

import torch  # noqa: E402

from make_golden import DATA_SEED, NOISE_SEED, NUM_MOLS, WEIGHT_SEED, manifest, record_rng  # noqa: E402
from make_golden_grads import summarize  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules  # noqa: E402
from oracle import refload  # noqa: E402
from oracle.weights import fill_state_dict  # noqa: E402


def main():
    R = refload.load()
    from Geom3D.models.MoleculeSDE import SDEModel3Dto2D_node_adj_dense_02
    from torch_geometric.utils import to_dense_batch
    torch.set_num_threads(1)
    mols = synth_molecules(NUM_MOLS, DATA_SEED)
    for m in mols:
        R.extend_graph(m)
    batch = Batch.from_data_list(mols)
    base = torch.load(os.path.join(HERE, "golden_pcqm8.pt"), weights_only=False)
    h3d = base["schnet"]["h"]
    out = {"meta": {"num_mols": NUM_MOLS, "data_seed": DATA_SEED, "weight_seed": WEIGHT_SEED, "torch": str(torch.__version__)}}
    for kind in ("VE", "VP"):
        m = SDEModel3Dto2D_node_adj_dense_02(
            dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300, num_linears=3,
            beta_min=0.1 if kind == "VE" else 0.2, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type=kind, num_class_X=119,
            noise_on_one_hot=True)
        m.load_state_dict(fill_state_dict(m.state_dict(), WEIGHT_SEED))
        sec = {"manifest": manifest(m.state_dict())}
        fix = base["sde3d2d_" + kind]           # the fixed perturbed state (x, adj, t, flags) of the base fixture
        m.eval()
        rep, _ = to_dense_batch(h3d, batch.batch, max_num_nodes=fix["nmax"])
        with torch.no_grad():
            emb = torch.cat([m.embedding_3D(rep), m.embedding_X(fix["x"])], -1)
            sec["score_adj"] = m.get_score_fn(m.sde_adj, m.edge_score_network, train=False)(emb, fix["adj"], fix["flags"], fix["t"])
            sec["score_x"] = m.get_score_fn(m.sde_x, m.node_score_network, train=False)(emb, fix["adj"], fix["flags"], fix["t"])
        m.train()
        hv = h3d.clone().requires_grad_(True)
        log = []
        with record_rng(log, NOISE_SEED + 60):
            lx, la = m(hv, batch, continuous=True, train=True, reduce_mean=True, anneal_power=0)
        ((lx + la) * 0.5).backward()
        sec.update({"loss_x": lx.detach(), "loss_adj": la.detach(), "draws": [v for _, v in log], "d_h3d": hv.grad.clone(),
                    "grads": {n: (summarize(p.grad) if p.grad is not None else None) for n, p in m.named_parameters()}})
        out["dense02_" + kind] = sec
        print(kind, "loss_x", float(lx), "loss_adj", float(la), "|score_adj|", float(sec["score_adj"].norm()),
              "|score_x|", float(sec["score_x"].norm()))
    path = os.path.join(HERE, "golden_dense02.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
