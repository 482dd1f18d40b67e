"""GPU parity of the pretraining step (forward losses, parameter gradients, Adam update) against the golden fixture
`tests/golden/golden_grads.pt` (the unmodified reference run over the shims with recorded draws).  Tolerance 1e-4
relative (max-norm per tensor), BASELINE.json north_star."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import sd_from_manifest  # noqa: E402

REL_TOL = 1e-4
# Measured on a B200 (gpurun_out/grad_parity_errors.txt, 1044 gradient tensors): every tensor of the 8-molecule fixture is within
# 5.6e-5 (sample, max-norm relative) / 5.4e-5 (norm, sum); at batch 32 all but 7 are within 1e-4.  The 7 are two documented families:
#  (a) CANCELLATION-DOMINATED tensors -- gradient norm below 1e-3 of the largest gradient norm of their module (the q/k biases of one
#      nearly saturated tanh-attention channel of the dense edge network, 1.5e-4 of the module scale; GINConv.eps, a scalar dot
#      product of ~1e5 cancelling terms): measured <= 2.3e-4, allowed 3e-4 (eps: 1e-3 as before);
#  (b) tensors in front of a BatchNorm + ReLU -- the GIN chain (gnn.*, d loss / d h2d) and edge_2D_emb.{0,1}.* of the 2D->3D model:
#      a ReLU unit sitting within rounding distance of 0 takes the other branch under a different fp32 summation order, which
#      moves the few gradient entries that depend on that unit while norm and sum stay within 2e-5.
#      Measured at batch 32: GIN <= 1.6e-4 (allowed 2e-4); edge_2D_emb.0.weight / .1.weight / .1.bias 5.3e-4 / 1.7e-4 / 6.8e-4
#      (allowed 1e-3, the bound round 1 used for every tensor).  The same values repeat bit for bit run after run.
SMALL_GRAD = 1e-3
BN_RELU_FAMILY = ("edge_2D_emb.0.", "edge_2D_emb.1.")
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gg():
    return torch.load(os.path.join(HERE, "golden", "golden_grads.pt"))


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


ERR_LOG = {}   # what -> (sample error, norm error, sum error), max-norm relative; dumped by test_zz_error_report


def check_grad_summary(got: torch.Tensor, want: dict, what: str, scale_floor: float = 0.0, tol: float = REL_TOL,
                       sample_tol: float = None):
    """Compare a gradient tensor with its fixture summary (norm, sum, strided sample).
      sample: max |got - ref| over the strided sample, relative to the tensor's scale max(rms, max |ref|)  (max-norm relative);
      norm:   | ||got|| - ||ref|| | / ||ref||;
      sum:    | sum(got) - sum(ref) | relative to ||ref||_2 * sqrt(numel) (>= the L1 norm: the scale of a sum's rounding error)."""
    sample_tol = tol if sample_tol is None else sample_tol
    f = got.detach().reshape(-1).float().cpu()
    assert f.numel() == want["numel"], what
    assert torch.isfinite(f).all(), f"{what}: non-finite"
    ref = want["sample"]
    smp = f[::want["stride"]][:ref.numel()]
    scale = max(float(want["norm"]) / max(f.numel(), 1) ** 0.5, float(ref.abs().max()), scale_floor, 1e-30)
    err = float((smp - ref).abs().max()) / scale
    nerr = abs(float(f.double().norm()) - float(want["norm"])) / max(float(want["norm"]), scale_floor, 1e-30)
    serr = abs(float(f.double().sum()) - float(want["sum"])) / max(float(want["norm"]) * max(f.numel(), 1) ** 0.5, scale_floor, 1e-30)
    ERR_LOG[what] = (err, nerr, serr)
    assert err <= sample_tol, f"{what}: sample error {err:.3e} (relative to {scale:.3e})"
    assert nerr <= tol, f"{what}: norm {float(f.double().norm()):.6e} vs {float(want['norm']):.6e} ({nerr:.3e})"
    assert serr <= tol, f"{what}: sum {float(f.double().sum()):.6e} vs {float(want['sum']):.6e} ({serr:.3e})"


def _draws_2d3d(sec):
    d = sec["draws"]
    kinds = [k for k, _ in d]
    assert kinds[:12] == ["randperm", "randperm", "randn", "randint"] + ["dropout"] * 8
    masks = [v for _, v in d[4:12]]
    return {"noise": d[2][1], "time_step": d[3][1], "dropout": [(masks[2 * i], masks[2 * i + 1]) for i in range(4)]}


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_2d3d_loss_and_grads(kind, gg, golden, golden_batch):
    from moleculesde_b200.pretrain import ParamStore, tape_2d3d
    from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02
    from moleculesde_b200.tape import Tape, Var
    from test_gpu_sde2d3d import _gpu_batch
    dev = _dev()
    sec = gg["pretrain_" + kind]
    _, batch = golden_batch
    model = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                              num_diffusion_timesteps=1000, SDE_type=kind, use_extend_graph=True)
    model.load_state_dict(sd_from_manifest(golden["manifest"]["sde2d3d"], golden["meta"]["weight_seed"]))
    model.train()
    store = ParamStore({"sde2d3d": model}, dev)
    b = _gpu_batch(batch, dev)
    tp = Tape(dev)
    h2d = Var(sec["h2d"].to(dev).contiguous(), True)
    loss = tape_2d3d(tp, model, store.vars("sde2d3d"), h2d, b, 0.0, _draws_2d3d(sec))
    ref = float(sec["loss_2d3d"])
    assert abs(float(loss) - ref) <= REL_TOL * abs(ref), (float(loss), ref)
    tp.backward()
    torch.cuda.synchronize()
    bad = []
    gmax = max(float(w["norm"]) for w in sec["grads"]["sde2d3d"].values() if w is not None)
    for name, want in sec["grads"]["sde2d3d"].items():
        if want is None:
            continue
        got = store.grad_view("sde2d3d", name)
        if name == "edge_2D_emb.0.bias" or name.endswith("lin_key.bias"):
            # analytically ZERO gradients (a bias in front of BatchNorm; a key bias shifts all logits of a softmax row
            # equally): the reference holds round-off noise here, so only the magnitude is checked
            assert float(want["norm"]) <= 1e-6 * gmax and float(got.norm()) <= 1e-6 * gmax, name
            continue
        try:
            check_grad_summary(got, want, name)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, "\n".join(bad)
    # BatchNorm running statistics were updated by the step
    for n, want in sec["buffers"]["sde2d3d"].items():
        got = dict(model.named_buffers())[n]
        check_grad_summary(got, want, n)


def _check_module_grads(store, mname, sec, skip_zero=(), sample_tol=None, tag=""):
    bad = []
    gmax = max(float(w["norm"]) for w in sec["grads"][mname].values() if w is not None)
    for name, want in sec["grads"][mname].items():
        if want is None:
            continue
        got = store.grad_view(mname, name)
        if any(name.endswith(z) or name == z for z in skip_zero):
            assert float(want["norm"]) <= 1e-5 * gmax and float(got.norm()) <= 1e-5 * gmax, name
            continue
        try:
            # GINConv.eps is a scalar whose gradient <d pre, x> is a heavily cancelling dot product of ~3e4 terms:
            # its relative error is the summands' 1e-6 times the cancellation factor
            t = REL_TOL * (10 if name.endswith(".eps") else 3 if float(want["norm"]) < SMALL_GRAD * gmax else 1)
            st = max(t, sample_tol or 0.0, 10 * REL_TOL if (tag and name.startswith(BN_RELU_FAMILY)) else 0.0)
            check_grad_summary(got, want, f"{tag}{mname}.{name}", tol=t, sample_tol=st)
        except AssertionError as e:
            bad.append(str(e))
    assert not bad, "\n".join(bad)


def _encoders(golden, dev):
    from moleculesde_b200.gnn import GNN
    from moleculesde_b200.schnet import SchNet
    gnn = GNN(5, 300, JK="last", drop_ratio=0.0, gnn_type="GIN")
    gnn.load_state_dict(sd_from_manifest(golden["manifest"]["gnn"], golden["meta"]["weight_seed"]))
    sch = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean",
                 node_class=119)
    sch.load_state_dict(sd_from_manifest(golden["manifest"]["schnet"], golden["meta"]["weight_seed"]))
    return gnn.train(), sch.train()


def test_gin_eval_forward(golden, golden_batch):
    """GNN.forward in eval mode (the representation the samplers consume) vs the reference output."""
    from test_gpu_sde2d3d import assert_parity
    dev = _dev()
    _, batch = golden_batch
    gnn, _ = _encoders(golden, dev)
    gnn = gnn.to(dev).eval()
    b = batch.to(dev)
    h = gnn(b.x, b.edge_index, b.edge_attr)
    assert_parity(h, golden["gnn"]["h_eval"], "GIN eval forward (x, edge_index, edge_attr)")
    assert torch.equal(h, gnn(b))


def test_encoders_forward_and_grads(gg, golden, golden_batch):
    """GIN and SchNet in train mode: representations vs the reference, then backward from the reference's total
    d loss / d representation -> every encoder parameter gradient."""
    from moleculesde_b200.pretrain import ParamStore, tape_gin, tape_schnet
    from moleculesde_b200.tape import Tape
    from test_gpu_sde2d3d import assert_parity
    dev = _dev()
    sec = gg["pretrain_VE"]
    _, batch = golden_batch
    gnn, sch = _encoders(golden, dev)
    store = ParamStore({"gnn": gnn, "schnet": sch}, dev)
    b = batch.to(dev)
    tp = Tape(dev)
    cache = {}
    h2d = tape_gin(tp, gnn, store.vars("gnn"), b.x, b.edge_index, b.edge_attr, cache, b.batch, b.num_graphs)
    h3d = tape_schnet(tp, sch, store.vars("schnet"), b.x[:, 0].contiguous(), b.positions, b.batch, b.num_graphs, cache)
    assert_parity(h2d.data, sec["h2d"], "GIN train-mode representation")
    assert_parity(h3d.data, sec["h3d"], "SchNet representation")
    h2d.grad = sec["d_h2d"].to(dev).contiguous()
    h3d.grad = sec["d_h3d"].to(dev).contiguous()
    tp.backward()
    torch.cuda.synchronize()
    # biases in front of a BatchNorm have analytically zero gradients
    _check_module_grads(store, "gnn", sec, skip_zero=("mlp.0.bias", "mlp.3.bias"))
    _check_module_grads(store, "schnet", sec)
    bufs = dict(gnn.named_buffers())
    for n, want in sec["buffers"]["gnn"].items():
        check_grad_summary(bufs[n], want, "gnn." + n)


def test_dual_cl_loss_and_backward(gg, golden):
    """dual_CL (EBM_node_dot_prod): loss vs the reference, gradients vs autograd over the oracle restatement."""
    from moleculesde_b200.pretrain import tape_dual_cl
    from moleculesde_b200.tape import Tape, Var
    from oracle import model as O
    from test_gpu_sde2d3d import assert_parity
    dev = _dev()
    sec = gg["pretrain_VE"]
    n1, n2 = sec["draws"][0][1], sec["draws"][1][1]
    X, Y = sec["h2d"], sec["h3d"]
    tp = Tape(dev)
    Xv, Yv = Var(X.to(dev).contiguous(), True), Var(Y.to(dev).contiguous(), True)
    loss, _ = tape_dual_cl(tp, Xv, Yv, 0.1, n1, n2, coef=1.0)
    ref = float(sec["cl_loss"])
    assert abs(float(loss) - ref) <= REL_TOL * abs(ref), (float(loss), ref)
    tp.backward()
    Xc, Yc = X.clone().requires_grad_(True), Y.clone().requires_grad_(True)
    l1, _ = O.do_cl_ebm_node_dot_prod(Xc, Yc, 0.1, n1)
    l2, _ = O.do_cl_ebm_node_dot_prod(Yc, Xc, 0.1, n2)
    ((l1 + l2) / 2).backward()
    assert_parity(Xv.grad, Xc.grad, "d CL / d X")
    assert_parity(Yv.grad, Yc.grad, "d CL / d Y")


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_3d2d_losses_and_grads(kind, gg, golden, golden_batch):
    """Dense 3D->2D model: both DSM losses and every parameter gradient of (loss_x + loss_adj) / 2."""
    from moleculesde_b200.pretrain import ParamStore, tape_3d2d
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    from moleculesde_b200.tape import Tape, Var
    dev = _dev()
    sec = gg["pretrain_" + kind]
    _, batch = golden_batch
    m32 = SDEModel3Dto2D_node_adj_dense(
        dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300, num_linears=3,
        beta_min=0.1 if kind == "VE" else 0.2, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type=kind, num_class_X=119,
        noise_on_one_hot=True)
    m32.load_state_dict(sd_from_manifest(golden["manifest"]["sde3d2d"], golden["meta"]["weight_seed"]))
    m32.train()
    store = ParamStore({"sde3d2d": m32}, dev)
    b = batch.to(dev)
    tp = Tape(dev)
    h3d = Var(sec["h3d"].to(dev).contiguous(), True)
    draws = [v for _, v in sec["draws"][12:15]]
    lx, la = tape_3d2d(tp, m32, store.vars("sde3d2d"), h3d, b, 0.0, draws, coef=0.5)
    for got, want, what in ((lx, sec["loss_x"], "loss_x"), (la, sec["loss_adj"], "loss_adj")):
        assert abs(float(got) - float(want)) <= REL_TOL * abs(float(want)), (what, float(got), float(want))
    tp.backward()
    torch.cuda.synchronize()
    _check_module_grads(store, "sde3d2d", sec)
    assert torch.isfinite(h3d.grad).all()


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_full_pretrain_step(kind, gg, golden, golden_batch):
    """The whole iteration: losses, d loss / d representations, every gradient, and the parameters after one Adam step."""
    from moleculesde_b200.pretrain import PretrainStep
    from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    from test_gpu_sde2d3d import _gpu_batch, assert_parity
    dev = _dev()
    sec = gg["pretrain_" + kind]
    _, batch = golden_batch
    gnn, sch = _encoders(golden, dev)
    m23 = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                            num_diffusion_timesteps=1000, SDE_type=kind, use_extend_graph=True)
    m23.load_state_dict(sd_from_manifest(golden["manifest"]["sde2d3d"], golden["meta"]["weight_seed"]))
    m32 = SDEModel3Dto2D_node_adj_dense(
        dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300, num_linears=3,
        beta_min=0.1 if kind == "VE" else 0.2, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type=kind, num_class_X=119,
        noise_on_one_hot=True)
    m32.load_state_dict(sd_from_manifest(golden["manifest"]["sde3d2d"], golden["meta"]["weight_seed"]))
    ps = PretrainStep(gnn, sch, m23, m32, dev, lr=1e-4, T=0.1)
    b = _gpu_batch(batch, dev)
    d = sec["draws"]
    draws = {"cl": (d[0][1], d[1][1]), "sde2d3d": _draws_2d3d(sec), "sde3d2d": [v for _, v in d[12:15]]}
    out = ps.forward_backward(b, draws)
    torch.cuda.synchronize()
    for key, want in (("cl_loss", "cl_loss"), ("loss_2d3d", "loss_2d3d"), ("loss_x", "loss_x"), ("loss_adj", "loss_adj")):
        assert abs(float(out[key]) - float(sec[want])) <= REL_TOL * abs(float(sec[want])), key
    assert abs(PretrainStep.total_loss(out) - float(sec["loss"])) <= REL_TOL * abs(float(sec["loss"]))
    assert_parity(out["h2d"].grad, sec["d_h2d"], "d loss / d node_2D_repr")
    assert_parity(out["h3d"].grad, sec["d_h3d"], "d loss / d node_3D_repr")
    _check_module_grads(ps.store, "gnn", sec, skip_zero=("mlp.0.bias", "mlp.3.bias"))
    _check_module_grads(ps.store, "schnet", sec)
    _check_module_grads(ps.store, "sde3d2d", sec)
    zero = ("edge_2D_emb.0.bias", "lin_key.bias")
    _check_module_grads(ps.store, "sde2d3d", sec, skip_zero=zero)
    assert ps.launches > 0
    # Adam (lr 1e-4, first step => |update| ~ lr): compare the parameter samples after the step
    ps.store.adam_step(ps.lr, ps.lr_scale)
    torch.cuda.synchronize()
    mods = {"gnn": gnn, "schnet": sch, "sde2d3d": m23, "sde3d2d": m32}
    worst, n_bad, n_all = 0.0, 0, 0
    for mname, m in mods.items():
        params = dict(m.named_parameters())
        for n, want in sec["after_step"][mname].items():
            if sec["grads"][mname].get(n) is None:
                continue
            if any(n.endswith(zz) or n == zz for zz in zero + ("mlp.0.bias", "mlp.3.bias")):
                continue  # sign of round-off noise decides the +-lr update of analytically-zero gradients
            gn = float(sec["grads"][mname][n]["norm"]) / max(sec["grads"][mname][n]["numel"], 1) ** 0.5
            if gn < 1e-6:
                continue
            f = params[n].detach().reshape(-1).float().cpu()
            smp = f[::want["stride"]][:want["sample"].numel()]
            diff = (smp - want["sample"]).abs()
            worst = max(worst, float(diff.max()))
            n_bad += int((diff > 2e-6).sum())
            n_all += diff.numel()
    # the first Adam step moves every element by ~lr * sign(g): elements whose gradient is ~0 may flip sign between
    # implementations (error up to 2 lr); all others must agree to a small fraction of lr
    assert worst <= 2.1e-4, worst
    assert n_bad <= 0.01 * n_all, (n_bad, n_all)


def test_full_pretrain_step_b32(golden):
    """BASELINE.json configs[0]: ONE pretrain_MoleculeSDE step at batch 32 (VE / VE, extended graph) against the unmodified
    reference run on the same molecules with recorded draws (`tests/golden/golden_grads_b32.pt`, made by
    `make_golden_grads.py --b32`): the four losses, both representations and d loss / d representation (summaries), every
    parameter gradient (norm, sum, strided sample) and the BatchNorm running statistics."""
    from moleculesde_b200.data import Batch, synth_molecules
    from moleculesde_b200.pretrain import PretrainStep
    from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    from oracle.ref_ops import extend_graph_index
    from test_gpu_sde2d3d import _gpu_batch
    dev = _dev()
    g32 = torch.load(os.path.join(HERE, "golden", "golden_grads_b32.pt"))
    sec, meta = g32["pretrain_VE"], g32["meta"]
    assert meta["num_mols"] == 32
    mols = synth_molecules(meta["num_mols"], meta["data_seed"])
    for m in mols:
        m.extended_edge_index = extend_graph_index(m.edge_index, m.num_nodes)
    batch = Batch.from_data_list(mols)
    assert batch.positions.size(0) == meta["num_atoms"] and batch.extended_edge_index.size(1) == meta["num_ext_edges"]
    gnn, sch = _encoders(golden, dev)
    m23 = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                            num_diffusion_timesteps=1000, SDE_type="VE", use_extend_graph=True)
    m23.load_state_dict(sd_from_manifest(golden["manifest"]["sde2d3d"], meta["weight_seed"]))
    m32 = SDEModel3Dto2D_node_adj_dense(
        dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300, num_linears=3,
        beta_min=0.1, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type="VE", num_class_X=119, noise_on_one_hot=True)
    m32.load_state_dict(sd_from_manifest(golden["manifest"]["sde3d2d"], meta["weight_seed"]))
    ps = PretrainStep(gnn, sch, m23, m32, dev, lr=1e-4, T=0.1)
    b = _gpu_batch(batch, dev)
    d = sec["draws"]
    draws = {"cl": (d[0][1], d[1][1]), "sde2d3d": _draws_2d3d(sec), "sde3d2d": [v for _, v in d[12:15]]}
    out = ps.forward_backward(b, draws)
    torch.cuda.synchronize()
    for key in ("cl_loss", "loss_2d3d", "loss_x", "loss_adj"):
        assert abs(float(out[key]) - float(sec[key])) <= REL_TOL * abs(float(sec[key])), (key, float(out[key]), float(sec[key]))
    assert abs(PretrainStep.total_loss(out) - float(sec["loss"])) <= REL_TOL * abs(float(sec["loss"]))
    check_grad_summary(out["h2d"].data, sec["h2d"], "b32.h2d", sample_tol=REL_TOL)
    check_grad_summary(out["h3d"].data, sec["h3d"], "b32.h3d", sample_tol=REL_TOL)
    check_grad_summary(out["h2d"].grad, sec["d_h2d"], "b32.d_h2d", sample_tol=2 * REL_TOL)   # family (b)
    check_grad_summary(out["h3d"].grad, sec["d_h3d"], "b32.d_h3d")
    _check_module_grads(ps.store, "gnn", sec, skip_zero=("mlp.0.bias", "mlp.3.bias"), sample_tol=2 * REL_TOL, tag="b32.")   # (b)
    _check_module_grads(ps.store, "schnet", sec, tag="b32.")
    _check_module_grads(ps.store, "sde3d2d", sec, tag="b32.")
    _check_module_grads(ps.store, "sde2d3d", sec, skip_zero=("edge_2D_emb.0.bias", "lin_key.bias"), tag="b32.")
    for mname, m in (("gnn", gnn), ("sde2d3d", m23)):
        bufs = dict(m.named_buffers())
        for n, want in sec["buffers"][mname].items():
            check_grad_summary(bufs[n], want, f"b32.{mname}.{n}", sample_tol=REL_TOL)


@pytest.mark.parametrize("num_mols,seed", [(1, 3), (3, 4), (17, 5)])
def test_pretrain_step_small_and_odd_batches(num_mols, seed, golden):
    """Batch sizes the antithetic time sampling treats specially (B = 1, odd B): losses vs the oracle with the same draws,
    gradients finite and non-trivial, one Adam step keeps the parameters finite."""
    import bench
    from moleculesde_b200 import graph as G
    from moleculesde_b200.data import Batch, synth_molecules
    from moleculesde_b200.pretrain import PretrainStep
    from oracle import model as O
    from oracle.ref_ops import extend_graph_index
    from oracle.weights import fill_state_dict
    dev = _dev()
    mods = bench.build_pretrain_models()
    sds = {}
    for k, m in zip(("gnn", "schnet", "sde2d3d", "sde3d2d"), mods):
        m.load_state_dict(fill_state_dict(m.state_dict(), 1))
        sds[k] = {n: v.clone() for n, v in m.state_dict().items()}
    mols = synth_molecules(num_mols, seed, "pcqm")
    for m in mols:
        m.extended_edge_index = extend_graph_index(m.edge_index, m.num_nodes)
    hb = Batch.from_data_list(mols)
    N, E, B = hb.positions.size(0), hb.extended_edge_index.size(1), hb.num_graphs
    nmax = int(torch.bincount(hb.batch).max())
    g = torch.Generator().manual_seed(seed)
    draws = {"cl": (torch.randperm(N, generator=g), torch.randperm(N, generator=g)),
             "sde2d3d": {"noise": torch.randn(N, 3, generator=g), "time_step": torch.randint(0, 1000, (B // 2 + 1,), generator=g),
                         "dropout": [((torch.rand(E, 8, generator=g) >= 0.1).float(), (torch.rand(N, 32, generator=g) >= 0.1).float())
                                     for _ in range(4)]},
             "sde3d2d": [torch.randint(0, 1000, (B // 2 + 1,), generator=g), torch.randn(B, nmax, nmax, generator=g),
                         torch.randn(B, nmax, 119, generator=g)]}
    with torch.no_grad():
        ref = O.pretrain_losses(sds, "VE", hb, draws)
    ps = PretrainStep(*mods, dev)
    b = hb.to(dev)
    b.extended_edge_index = G.extend_graph(b.edge_index, b.batch, b.num_graphs).edge_index
    out = ps.step(b, draws)
    for k in ("cl_loss", "loss_2d3d", "loss_x", "loss_adj"):
        assert abs(float(out[k]) - float(ref[k])) <= REL_TOL * max(abs(float(ref[k])), 1e-3), (k, float(out[k]), float(ref[k]))
    assert torch.isfinite(ps.store.grad).all() and float(ps.store.grad.abs().max()) > 0
    assert torch.isfinite(ps.store.flat).all()


def test_reference_training_loop_through_autograd(gg, golden, golden_batch):
    """The reference's own loop body (`pretrain_MoleculeSDE.py:124-152`): module calls, `loss.backward()`, torch.optim.Adam —
    unchanged, on our modules through the autograd bridge.  Losses and gradients match the reference step fixture."""
    import types
    from moleculesde_b200.contrastive import dual_CL
    from moleculesde_b200.sde_2d_to_3d import SDEModel2Dto3D_02
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    from test_gpu_sde2d3d import _gpu_batch
    dev = _dev()
    sec = gg["pretrain_VE"]
    _, batch = golden_batch
    molecule_model_2D, molecule_model_3D = _encoders(golden, dev)
    SDE_2Dto3D_model = SDEModel2Dto3D_02(emb_dim=300, hidden_dim=32, beta_schedule=None, beta_min=0.2, beta_max=1.0,
                                         num_diffusion_timesteps=1000, SDE_type="VE", use_extend_graph=True)
    SDE_2Dto3D_model.load_state_dict(sd_from_manifest(golden["manifest"]["sde2d3d"], golden["meta"]["weight_seed"]))
    SDE_3Dto2D_model = SDEModel3Dto2D_node_adj_dense(
        dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4, emb_dim=300, num_linears=3,
        beta_min=0.1, beta_max=1.0, num_diffusion_timesteps=1000, SDE_type="VE", num_class_X=119, noise_on_one_hot=True)
    SDE_3Dto2D_model.load_state_dict(sd_from_manifest(golden["manifest"]["sde3d2d"], golden["meta"]["weight_seed"]))
    mods = {"gnn": molecule_model_2D, "schnet": molecule_model_3D, "sde2d3d": SDE_2Dto3D_model, "sde3d2d": SDE_3Dto2D_model}
    for m in mods.values():
        m.to(dev).train()
    optimizer = torch.optim.Adam([{"params": m.parameters(), "lr": 1e-4} for m in mods.values()], lr=1e-4, weight_decay=0)
    args = types.SimpleNamespace(CL_similarity_metric="EBM_node_dot_prod", T=0.1)
    b = _gpu_batch(batch, dev)
    d = sec["draws"]
    # ---- the reference loop body, with the recorded draws injected through the keyword-only extensions ----
    node_2D_repr = molecule_model_2D(b.x, b.edge_index, b.edge_attr)
    _, node_3D_repr = molecule_model_3D(b.x[:, 0].contiguous(), b.positions, b.batch, return_latent=True)
    CL_loss, CL_acc = dual_CL(node_2D_repr, node_3D_repr, args, d[0][1], d[1][1])
    loss = CL_loss * 1.0
    SDE_loss_2Dto3D = SDE_2Dto3D_model(node_2D_repr, b, anneal_power=0, draws=_draws_2d3d(sec))["position"]
    loss = loss + SDE_loss_2Dto3D * 1.0
    lx, la = SDE_3Dto2D_model(node_3D_repr, b, reduce_mean=True, continuous=True, train=True, anneal_power=0,
                              draws=[v for _, v in d[12:15]])
    loss = loss + (lx + la) * 0.5 * 1.0
    optimizer.zero_grad()
    loss.backward()
    # ----
    assert abs(float(loss.detach()) - float(sec["loss"])) <= REL_TOL * abs(float(sec["loss"]))
    assert abs(float(CL_loss.detach()) - float(sec["cl_loss"])) <= REL_TOL * abs(float(sec["cl_loss"]))

    class _S:  # adapter: _check_module_grads reads store.grad_view(module, name)
        def grad_view(self, mname, pname):
            return dict(mods[mname].named_parameters())[pname].grad
    zero = ("edge_2D_emb.0.bias", "lin_key.bias")
    _check_module_grads(_S(), "gnn", sec, skip_zero=("mlp.0.bias", "mlp.3.bias"))
    _check_module_grads(_S(), "schnet", sec)
    _check_module_grads(_S(), "sde2d3d", sec, skip_zero=zero)
    _check_module_grads(_S(), "sde3d2d", sec)
    optimizer.step()
    assert all(torch.isfinite(p).all() for m in mods.values() for p in m.parameters())


def test_zz_error_report():
    """Not a check: writes the achieved max-norm relative errors of every gradient comparison of this module (sample / norm / sum)
    to gpurun_out/ so that the tolerances above can be read against measurements."""
    if not ERR_LOG:
        pytest.skip("no gradient comparison ran")
    out_dir = os.path.join(os.path.dirname(HERE), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    rows = sorted(ERR_LOG.items(), key=lambda kv: -kv[1][0])
    with open(os.path.join(out_dir, "grad_parity_errors.txt"), "w") as f:
        f.write("# what  sample_err  norm_err  sum_err   (max-norm relative, see check_grad_summary)\n")
        for k, (a, b_, c) in rows:
            f.write(f"{k:70s} {a:.3e} {b_:.3e} {c:.3e}\n")
        worst = [max(v[i] for v in ERR_LOG.values()) for i in range(3)]
        f.write(f"# worst: sample {worst[0]:.3e} norm {worst[1]:.3e} sum {worst[2]:.3e} over {len(rows)} tensors\n")
