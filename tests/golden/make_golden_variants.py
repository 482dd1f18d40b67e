"""Golden fixture for the sibling variants of the 2D->3D model, from the UNMODIFIED reference over
`oracle/shims` (build container only):

    python tests/golden/make_golden_variants.py   ->  golden_variants.pt

* `SDEModel2Dto3D_01` (`SDE_model_2D_to_3D.py:69-250`, the variant several published checkpoints use,
  `README_checkpoints.md:9-49`): state_dict manifest, `get_score`, one training loss with every draw recorded, all parameter
  gradients (norm / sum / strided sample) and d loss / d node_2D_repr.
* the noise-schedule presets of `pretrain_MoleculeSDE.py:226-256` (VE02, VP02, VE03, VP03) on `SDEModel2Dto3D_02.get_score`.
Same synthetic molecules (seed 0) and name-keyed weights (seed 1) as `make_golden.py`.
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

# pretrain_MoleculeSDE.py:226-256 (2D->3D direction)
PRESETS = {"VE02": ("VE", 0.1, 10.0), "VP02": ("VP", 0.2, 30.0), "VE03": ("VE", 0.1, 1000.0), "VP03": ("VP", 0.2, 1000.0)}


def main():
    R = refload.load()
    torch.set_num_threads(1)
    mols = synth_molecules(NUM_MOLS, DATA_SEED)
    for m in mols:
        R.extend_graph(m)
    batch = Batch.from_data_list(mols)
    N = batch.positions.size(0)
    base = torch.load(os.path.join(HERE, "golden_pcqm8.pt"), weights_only=False)
    h2d = base["gnn"]["h_eval"]
    # training section: the train-mode GIN output of the golden_grads fixture.  edge_2D_emb (Linear -> BatchNorm -> ReLU) is the
    # same in _01 and _02, and for this representation no ReLU input sits within rounding distance of zero, so fp32
    # implementations agree to 1e-4 on every gradient entry (with the eval-mode representation one unit of one edge flips, which
    # moves two rows of d loss / d node_2D_repr by 1e-3 -- the effect DESIGN.md section 4b describes).
    grads_fix = torch.load(os.path.join(HERE, "golden_grads.pt"), weights_only=False)
    out = {"meta": {"num_mols": NUM_MOLS, "data_seed": DATA_SEED, "weight_seed": WEIGHT_SEED, "torch": str(torch.__version__)}}
    g = torch.Generator().manual_seed(NOISE_SEED + 40)
    pos_p = batch.positions + 0.3 * torch.randn(N, 3, generator=g)
    t = (torch.rand(NUM_MOLS, generator=g) * 0.9 + 0.05)[batch.batch]
    out["inputs"] = {"pos_perturbed": pos_p, "t": t}

    for kind in ("VE", "VP"):
        m = R.SDEModel2Dto3D_01(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                                num_diffusion_timesteps=1000, SDE_type=kind, use_extend_graph=True)
        m.load_state_dict(fill_state_dict(m.state_dict(), WEIGHT_SEED))
        sec = {"manifest": manifest(m.state_dict())}
        m.eval()
        sec["score"] = m.get_score(h2d, batch, pos_p, None, t).detach()
        m.train()
        log = []
        hv = grads_fix["pretrain_" + kind]["h2d"].clone().requires_grad_(True)
        with record_rng(log, NOISE_SEED + 41):
            loss = m(hv, batch.clone(), anneal_power=0)["position"]
        loss.backward()
        sec.update({"train_h2d": hv.detach().clone(), "train_loss": loss.detach(), "train_draws": [(k, v) for k, v in log], "d_h2d": hv.grad.clone(),
                    "grads": {n: (summarize(p.grad) if p.grad is not None else None) for n, p in m.named_parameters()
                              if p.requires_grad},
                    "bn_running_mean": m.edge_2D_emb[1].running_mean.clone(),
                    "bn_running_var": m.edge_2D_emb[1].running_var.clone()})
        out["sde2d3d_01_" + kind] = sec
        print("SDEModel2Dto3D_01", kind, "loss", float(loss), "|score|", float(sec["score"].norm()))

    for name, (kind, bmin, bmax) in PRESETS.items():
        m = R.SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=bmin, beta_max=bmax,
                                num_diffusion_timesteps=1000, SDE_type=kind, use_extend_graph=True)
        m.load_state_dict(fill_state_dict(m.state_dict(), WEIGHT_SEED))
        m.eval()
        # positions perturbed at the preset's own noise level (large sigma: the regime the Fourier features must survive)
        _, std = m.sde_pos.marGINal_prob(batch.positions, t)
        gp = torch.Generator().manual_seed(NOISE_SEED + 42)
        pp = batch.positions + std[:, None] * torch.randn(N, 3, generator=gp) if std.dim() == 1 else \
            batch.positions + std * torch.randn(N, 3, generator=gp)
        out["preset_" + name] = {"kind": kind, "beta_min": bmin, "beta_max": bmax, "pos_perturbed": pp,
                                 "score": m.get_score(h2d, batch, pp, None, t).detach()}
        print(name, "|score|", float(out["preset_" + name]["score"].norm()), "max |pos|", float(pp.abs().max()))
    # ---- SchNet energy head + forces (finetune_MD17.py:49-66): E_g = <readout(h)_g, w>, F = -dE/dpos through torch.autograd
    sch = R.SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean",
                   node_class=119)
    sch.load_state_dict(fill_state_dict(sch.state_dict(), WEIGHT_SEED))
    sch.train()
    gw = torch.Generator().manual_seed(NOISE_SEED + 43)
    w = torch.randn(300, generator=gw) / 17.0
    pos = batch.positions.clone().requires_grad_(True)
    out_g = sch(batch.x[:, 0], pos, batch.batch)
    energy = out_g @ w
    force = -torch.autograd.grad(energy, pos, torch.ones_like(energy), retain_graph=True)[0]
    # a fine-tuning style scalar loss on the energies: parameter gradients through the differentiable readout
    target = torch.linspace(-0.05, 0.05, NUM_MOLS)
    loss = ((energy - target) ** 2).mean()
    loss.backward()
    out["schnet_force"] = {"w": w, "out": out_g.detach(), "energy": energy.detach(), "force": force.detach(), "target": target,
                           "loss": loss.detach(), "d_pos": pos.grad.clone(),
                           "grads": {n: summarize(p.grad) for n, p in sch.named_parameters() if p.grad is not None}}
    print("SchNet energy", energy.detach()[:3], "|F|", float(force.norm()), "loss", float(loss))
    path = os.path.join(HERE, "golden_variants.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
