"""Generate the golden fixtures in this directory by executing the UNMODIFIED reference
sources (`/root/reference`) over `oracle/shims`.  Build-container only:

    python tests/golden/make_golden.py

Inputs are the seeded synthetic molecules of `moleculesde_b200.data` (seed 0), weights the
name-keyed fill of `oracle/weights.py` (seed 1); random draws made by the reference
(`randn_like`, `randint`, `randperm`, dropout masks) are recorded in call order so that the
oracle and the CUDA path can be fed the same numbers.  Only outputs, recorded draws and
state_dict key/shape manifests are stored (no weights), so the files stay small.
"""
from __future__ import annotations

import contextlib
import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import refload  # noqa: E402
from oracle.weights import fill_state_dict  # noqa: E402
from moleculesde_b200.data import Batch, synth_molecules, repeat_data  # noqa: E402

NUM_MOLS = 8
DATA_SEED, WEIGHT_SEED, NOISE_SEED = 0, 1, 2


@contextlib.contextmanager
def record_rng(log: list, seed: int):
    """Patch the RNG entry points the reference uses; every draw comes from one seeded
    generator and is appended to `log` as (kind, tensor)."""
    g = torch.Generator().manual_seed(seed)
    o_randn_like, o_randint, o_randperm, o_dropout, o_randn = (
        torch.randn_like, torch.randint, torch.randperm, F.dropout, torch.randn)

    def randn_like(x, **kw):
        v = o_randn(x.shape, generator=g, dtype=x.dtype)
        log.append(("randn", v))
        return v.clone()

    def randn(*shape, **kw):
        if len(shape) == 1 and isinstance(shape[0], (tuple, list, torch.Size)):
            shape = tuple(shape[0])
        v = o_randn(shape, generator=g)
        log.append(("randn", v))
        return v.clone()

    def randint(low, high, size, **kw):
        v = o_randint(low, high, size, generator=g)
        log.append(("randint", v))
        return v.clone()

    def randperm(n, **kw):
        v = o_randperm(n, generator=g)
        log.append(("randperm", v))
        return v.clone()

    def dropout(x, p=0.5, training=True, inplace=False):
        if not training or p == 0:
            return x
        mask = (torch.rand(x.shape, generator=g) >= p).to(x.dtype)
        log.append(("dropout", mask))
        return x * mask / (1.0 - p)

    torch.randn_like, torch.randint, torch.randperm, F.dropout, torch.randn = (
        randn_like, randint, randperm, dropout, randn)
    try:
        yield
    finally:
        torch.randn_like, torch.randint, torch.randperm, F.dropout, torch.randn = (
            o_randn_like, o_randint, o_randperm, o_dropout, o_randn)


def manifest(sd):
    return {k: (tuple(v.shape), str(v.dtype)) for k, v in sd.items()}


def build_reference_models(R, sde_type="VE"):
    gnn = R.GNN(5, 300, JK="last", drop_ratio=0.0, gnn_type="GIN")
    sch = R.SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51,
                   cutoff=10, readout="mean", node_class=119)
    m23 = R.SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                              num_diffusion_timesteps=1000, SDE_type=sde_type, use_extend_graph=True)
    m32 = R.SDEModel3Dto2D_node_adj_dense(
        dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300,
        num_linears=3, beta_min=0.1 if sde_type == "VE" else 0.2, beta_max=1.0, num_diffusion_timesteps=1000,
        SDE_type=sde_type, num_class_X=119, noise_on_one_hot=True)
    for m in (gnn, sch, m23, m32):
        m.load_state_dict(fill_state_dict(m.state_dict(), WEIGHT_SEED))
    return gnn, sch, m23, m32


def main():
    R = refload.load()
    torch.set_num_threads(1)
    mols = synth_molecules(NUM_MOLS, DATA_SEED)
    for m in mols:
        R.extend_graph(m)  # dataset_3D.py:12-35, attaches extended_edge_index
    batch = Batch.from_data_list(mols)
    N = batch.positions.size(0)
    out = {"meta": {"num_mols": NUM_MOLS, "data_seed": DATA_SEED, "weight_seed": WEIGHT_SEED,
                    "noise_seed": NOISE_SEED, "torch": torch.__version__}}
    out["graph"] = {"extended_edge_index": batch.extended_edge_index.clone(),
                    "num_nodes": torch.tensor([m.num_nodes for m in mols])}

    for sde_type in ("VE", "VP"):
        gnn, sch, m23, m32 = build_reference_models(R, sde_type)
        if sde_type == "VE":
            out["manifest"] = {"gnn": manifest(gnn.state_dict()), "schnet": manifest(sch.state_dict()),
                               "sde2d3d": manifest(m23.state_dict()), "sde3d2d": manifest(m32.state_dict())}
            # ---- encoders (eval mode: what the samplers see) ----
            gnn.eval(); sch.eval()
            with torch.no_grad():
                h2d_eval = gnn(batch.x, batch.edge_index, batch.edge_attr)
                g_out, h3d = sch(batch.x[:, 0], batch.positions, batch.batch, return_latent=True)
            from torch_cluster import radius_graph
            out["schnet"] = {"out": g_out, "h": h3d,
                             "radius_edge_index": radius_graph(batch.positions, r=10, batch=batch.batch)}
            out["gnn"] = {"h_eval": h2d_eval}
            # ---- contrastive ----
            log = []
            with record_rng(log, NOISE_SEED):
                args = types.SimpleNamespace(CL_similarity_metric="EBM_node_dot_prod", T=0.1)
                loss, acc = R.examples.dual_CL(h2d_eval, h3d, args)
            out["cl"] = {"loss": loss, "acc": torch.tensor(acc), "neg_index": [v for _, v in log]}
        h2d_eval = out["gnn"]["h_eval"]
        # ---- 2D->3D: get_score (eval) ----
        m23.eval()
        g = torch.Generator().manual_seed(NOISE_SEED)
        pos_p = batch.positions + 0.3 * torch.randn(N, 3, generator=g)
        t = (torch.rand(NUM_MOLS, generator=g) * 0.9 + 0.05)[batch.batch]
        score = m23.get_score(h2d_eval, batch, pos_p, None, t)
        sec = {"pos_perturbed": pos_p, "t": t, "score": score}
        # ---- 2D->3D: training loss with recorded draws (dropout active, BN batch stats) ----
        m23.train()
        log = []
        with record_rng(log, NOISE_SEED + 1):
            b2 = batch.clone()
            loss = m23(h2d_eval.clone().requires_grad_(True), b2, anneal_power=0)["position"]
        sec["train_loss"] = loss.detach()
        sec["train_draws"] = [(k, v) for k, v in log]
        sec["bn_running_mean"] = m23.edge_2D_emb[1].running_mean.clone()
        sec["bn_running_var"] = m23.edge_2D_emb[1].running_var.clone()
        grads = torch.autograd.grad(loss, [p for p in m23.parameters() if p.requires_grad], allow_unused=True)
        names = [n for n, p in m23.named_parameters() if p.requires_grad]
        sec["train_grad_norms"] = {n: (gr.norm() if gr is not None else torch.tensor(0.0)) for n, gr in zip(names, grads)}
        # ---- 2D->3D: PC sampler, R=4 conformers of molecule 0, 6 reverse steps ----
        m23.load_state_dict(fill_state_dict(m23.state_dict(), WEIGHT_SEED))  # undo BN stat update
        m23.eval()
        mol0 = mols[0]
        rb = repeat_data(mol0, 4)
        gnn.eval()
        with torch.no_grad():
            rep = gnn(rb.x, rb.edge_index, rb.edge_attr)
        calls = []
        orig_get_score = m23.get_score

        def traced(representation, data, pos, sigma, tt):
            s = orig_get_score(representation, data, pos, sigma, tt)
            calls.append((pos.clone(), tt.clone(), s.clone()))
            return s

        m23.get_score = traced
        log = []
        n_steps = 6
        glb = R.examples.pc2d3d_globals
        # run the reference loop for n_steps by truncating the linspace it iterates over
        real_N = m23.sde_pos.N
        with record_rng(log, NOISE_SEED + 2):
            pos_init = m23.sde_pos.prior_sampling(rb.positions.shape)
            class _Stop(Exception):
                pass
            count = {"n": 0}
            o_trange = glb["trange"]
            glb["trange"] = lambda a, b, **kw: range(a, min(b, n_steps))
            try:
                _, pos_mean = R.examples.pc2d3d.position_PC_generation(
                    representation=rep, data=rb, pos_init=pos_init, scorenet=m23, sde=m23.sde_pos, n_steps=1)
            finally:
                glb["trange"] = o_trange
        del m23.get_score
        assert m23.sde_pos.N == real_N
        sec["pc"] = {"repeat": 4, "steps": n_steps, "representation": rep, "pos_mean": pos_mean,
                     "draws": [v for _, v in log],
                     "calls": calls}
        out["sde2d3d_" + sde_type] = sec

    # ---------------- dense 3D -> 2D (VE and VP): score networks, DSM losses, 2 PC sampler steps ----------------
    h3d = out["schnet"]["h"]
    for sde_type in ("VE", "VP"):
        _, _, _, m32 = build_reference_models(R, sde_type)
        m32.train()
        log = []
        with record_rng(log, NOISE_SEED + 5):
            lx, la = m32(h3d, batch, continuous=True, train=True, reduce_mean=True, anneal_power=0)
        sec = {"loss_x": lx.detach(), "loss_adj": la.detach(), "draws": [v for _, v in log]}
        # score networks on a fixed perturbed state
        from torch_geometric.utils import to_dense_adj, to_dense_batch
        import Geom3D.models.MoleculeSDE.SDE_model_3D_to_2D_node_adj_dense as M32
        m32.eval()
        edge_attr = batch.edge_attr[:, 0].float() + 1
        nmax = int(torch.bincount(batch.batch).max())
        adj = to_dense_adj(batch.edge_index, batch.batch, edge_attr=edge_attr, max_num_nodes=nmax)
        rep, _ = to_dense_batch(h3d, batch.batch, max_num_nodes=nmax)
        flags = M32.node_flags(adj)
        g = torch.Generator().manual_seed(NOISE_SEED + 6)
        Bg = adj.size(0)
        xs = M32.mask_x(torch.randn(Bg, nmax, 119, generator=g), flags)
        za = torch.randn(Bg, nmax, nmax, generator=g).triu(1)
        pa = M32.mask_adjs(adj + 0.5 * (za + za.transpose(-1, -2)), flags)
        tt = torch.rand(Bg, generator=g) * 0.9 + 0.05
        with torch.no_grad():
            emb = m32.embedding_3D(rep) + m32.embedding_X(xs)
            s_adj = m32.get_score_fn(m32.sde_adj, m32.edge_score_network, train=False)(emb, pa, flags, tt)
            s_x = m32.get_score_fn(m32.sde_x, m32.node_score_network, train=False)(emb, pa, flags, tt)
        sec.update({"x": xs, "adj": pa, "t": tt, "flags": flags, "score_adj": s_adj, "score_x": s_x, "nmax": nmax})
        # sampler: lift node_adj_PC_generation + classes from the inference script (lines 95-252)
        import abc
        glb = {"torch": torch, "abc": abc, "trange": lambda a, b, **kw: range(a, min(b, 2)), "device": "cpu",
               "args": types.SimpleNamespace(device="cpu"), "print": lambda *a, **k: None,
               "to_dense_adj": to_dense_adj, "node_flags": M32.node_flags, "mask_x": M32.mask_x, "mask_adjs": M32.mask_adjs,
               "gen_noise": M32.gen_noise}
        from Geom3D.models.MoleculeSDE.SDE_sparse import VPSDE, VESDE, subVPSDE
        glb.update(VPSDE=VPSDE, VESDE=VESDE, subVPSDE=subVPSDE)
        refload._exec_slice(os.path.join(refload.REFERENCE_ROOT, "examples", "pretrain_MoleculeSDE_inference_3D_to_2D_VE_VP.py"),
                            95, 252, glb)
        log = []
        with record_rng(log, NOISE_SEED + 7):
            with torch.no_grad():
                x, a, xm, am = glb["node_adj_PC_generation"](representation=rep, data=batch, SDE_model=m32, B=Bg, max_num_nodes=nmax,
                                                             num_class_X=119, n_steps=1)
        sec["pc"] = {"steps": 2, "draws": [v for _, v in log], "x": x, "adj": a, "x_mean": xm, "adj_mean": am}
        out["sde3d2d_" + sde_type] = sec

    path = os.path.join(HERE, "golden_pcqm8.pt")
    torch.save(out, path)
    print("wrote", path, os.path.getsize(path) / 1e6, "MB")


if __name__ == "__main__":
    main()
