"""GPU parity of the dense 3D->2D model (score networks, DSM losses, PC sampler) against the golden outputs of the
reference's own sources, through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import sd_from_manifest  # noqa: E402
from test_gpu_sde2d3d import assert_parity, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _model(golden, kind, dev):
    from moleculesde_b200.sde_3d_to_2d import SDEModel3Dto2D_node_adj_dense
    m = SDEModel3Dto2D_node_adj_dense(dim3D=300, c_init=2, c_hid=8, c_final=4, num_heads=4, adim=16, nhid=16, num_layers=4,
                                      emb_dim=300, num_linears=3, beta_min=0.1 if kind == "VE" else 0.2, beta_max=1.0,
                                      num_diffusion_timesteps=1000, SDE_type=kind, num_class_X=119, noise_on_one_hot=True)
    mine = {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
    assert mine == golden["manifest"]["sde3d2d"]
    m.load_state_dict(sd_from_manifest(golden["manifest"]["sde3d2d"], golden["meta"]["weight_seed"]))
    return m.to(dev).eval()


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_dense_prologue_and_scores(kind, golden, golden_batch):
    dev = _dev()
    _, batch = golden_batch
    sec = golden["sde3d2d_" + kind]
    model = _model(golden, kind, dev)
    b = batch.to(dev)
    adj, rep, zd, flags, Nm = model.dense_inputs(golden["schnet"]["h"].to(dev), b)
    assert Nm == sec["nmax"]
    assert torch.equal(flags.cpu(), sec["flags"])
    emb = model.embed(rep, sec["x"].to(dev))
    t = sec["t"].to(dev)
    s_adj = model.get_score_fn(model.sde_adj, model.edge_score_network, train=False)(emb, sec["adj"].to(dev), flags, t)
    s_x = model.get_score_fn(model.sde_x, model.node_score_network, train=False)(emb, sec["adj"].to(dev), flags, t)
    assert_parity(s_adj, sec["score_adj"], f"edge score [{kind}]")
    assert_parity(s_x, sec["score_x"], f"node score [{kind}]")


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_dense_dsm_losses(kind, golden, golden_batch):
    dev = _dev()
    _, batch = golden_batch
    sec = golden["sde3d2d_" + kind]
    model = _model(golden, kind, dev).train()
    lx, la = model(golden["schnet"]["h"].to(dev), batch.to(dev), continuous=True, train=True, reduce_mean=True, anneal_power=0,
                   draws=sec["draws"])
    assert rel_err(lx.cpu().reshape(1), sec["loss_x"].reshape(1)) < 1e-4
    assert rel_err(la.cpu().reshape(1), sec["loss_adj"].reshape(1)) < 1e-4


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_dense_pc_sampler(kind, golden, golden_batch):
    from moleculesde_b200.sampler_dense import node_adj_PC_generation
    dev = _dev()
    _, batch = golden_batch
    sec = golden["sde3d2d_" + kind]
    pc = sec["pc"]
    model = _model(golden, kind, dev)
    b = batch.to(dev)
    _, rep, _, _, Nm = model.dense_inputs(golden["schnet"]["h"].to(dev), b)
    d = pc["draws"]
    order = {"c_adj": 0, "c_x": 1, "p_adj": 2, "p_x": 3}
    x, adj, xm, am = node_adj_PC_generation(rep, b, model, B=rep.size(0), max_num_nodes=Nm, num_class_X=119, n_steps=1,
                                            x_init=d[0], adj_init=d[1], draws=lambda k, i: d[2 + 4 * i + order[k]],
                                            diffusion_steps=pc["steps"])
    for got, ref, name in ((x, pc["x"], "x"), (adj, pc["adj"], "adj"), (xm, pc["x_mean"], "x_mean"), (am, pc["adj_mean"], "adj_mean")):
        assert torch.isfinite(got).all()
        assert rel_err(got.cpu(), ref) < 1e-3, f"{name} [{kind}] {rel_err(got.cpu(), ref):.3e}"


def test_dense_padded64(golden):
    """Config-4 shape: graphs padded to 64 atoms, VP, vs the oracle."""
    from moleculesde_b200.data import synth_batch
    from oracle import model as O
    dev = _dev()
    b = synth_batch(12, 9, "padded64")
    model = _model(golden, "VP", dev)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(3)
    h3d = torch.randn(b.positions.size(0), 300, generator=g)
    adj0, rep, _, flags = O.dense_inputs(h3d, b.x[:, 0], b.edge_index, b.edge_attr[:, 0], b.batch)
    Bg, Nm = adj0.size(0), adj0.size(1)
    assert Nm > 32
    x = O.mask_x(torch.randn(Bg, Nm, 119, generator=g), flags)
    za = torch.randn(Bg, Nm, Nm, generator=g).triu(1)
    pa = O.mask_adjs(adj0 + 0.3 * (za + za.transpose(-1, -2)), flags)
    t = torch.rand(Bg, generator=g) * 0.9 + 0.05
    sde = O.make_dense_sde("VP", 0.2, 1.0, 1000)
    emb = O.embed_3d2d(sd, rep, x)
    ref_a = O.score_3d2d(sd, sde, "adj", emb, pa, flags, t)
    ref_x = O.score_3d2d(sd, sde, "x", emb, pa, flags, t)
    _, rep_d, _, flags_d, Nm_d = model.dense_inputs(h3d.to(dev), b.to(dev))
    assert Nm_d == Nm and torch.equal(flags_d.cpu(), flags)
    emb_d = model.embed(rep_d, x.to(dev))
    s_a = model.get_score_fn(model.sde_adj, model.edge_score_network, train=False)(emb_d, pa.to(dev), flags_d, t.to(dev))
    s_x = model.get_score_fn(model.sde_x, model.node_score_network, train=False)(emb_d, pa.to(dev), flags_d, t.to(dev))
    assert_parity(s_a, ref_a, "edge score padded64")
    assert_parity(s_x, ref_x, "node score padded64")


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_dense_pc_sampler_graph_replay_vs_golden(kind, golden, golden_batch):
    """The CUDA-graph replay of one PC step (device-side step counter, injected draws read from per-kind tables) reproduces the
    reference's recorded 2-step trajectory and is bit-identical to the eager loop."""
    from moleculesde_b200.sampler_dense import node_adj_PC_generation
    dev = _dev()
    _, batch = golden_batch
    pc = golden["sde3d2d_" + kind]["pc"]
    model = _model(golden, kind, dev)
    b = batch.to(dev)
    _, rep, _, _, Nm = model.dense_inputs(golden["schnet"]["h"].to(dev), b)
    d = pc["draws"]
    order = {"c_adj": 0, "c_x": 1, "p_adj": 2, "p_x": 3}
    outs = {}
    for ug in (False, True):
        outs[ug] = node_adj_PC_generation(rep, b, model, B=rep.size(0), max_num_nodes=Nm, num_class_X=119, n_steps=1,
                                          x_init=d[0], adj_init=d[1], draws=lambda k, i: d[2 + 4 * i + order[k]],
                                          diffusion_steps=pc["steps"], use_graph=ug)
    for got, ref, name in zip(outs[True], (pc["x"], pc["adj"], pc["x_mean"], pc["adj_mean"]), ("x", "adj", "x_mean", "adj_mean")):
        assert rel_err(got.cpu(), ref) < 1e-3, f"{name} [{kind}] {rel_err(got.cpu(), ref):.3e}"
    for a, e in zip(outs[True], outs[False]):
        assert torch.equal(a, e), "graph replay differs from the eager loop"


def test_dense_pc_sampler_20_steps_teacher_forced_vs_oracle(golden):
    """20 PC steps (VP, graphs padded to 64 atoms) with injected draws: graph replay == eager loop bit for bit, and along the
    visited states the GPU scores agree with the oracle's (1e-4) -- the per-step agreement that bounds the trajectory drift."""
    from moleculesde_b200.data import synth_batch
    from moleculesde_b200.sampler_dense import node_adj_PC_generation
    from oracle import model as O
    dev = _dev()
    steps, Bg = 20, 6
    b = synth_batch(Bg, 11, "padded64")
    model = _model(golden, "VP", dev)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    g = torch.Generator().manual_seed(5)
    h3d = torch.randn(b.positions.size(0), 300, generator=g)
    bd = b.to(dev)
    _, rep, _, flags, Nm = model.dense_inputs(h3d.to(dev), bd)
    x0 = torch.randn(Bg, Nm, 119, generator=g)
    a0 = torch.randn(Bg, Nm, Nm, generator=g).triu(1)
    a0 = a0 + a0.transpose(-1, -2)
    tabs = {"c_adj": torch.randn(steps, Bg, Nm, Nm, generator=g), "p_adj": torch.randn(steps, Bg, Nm, Nm, generator=g),
            "c_x": torch.randn(steps, Bg, Nm, 119, generator=g), "p_x": torch.randn(steps, Bg, Nm, 119, generator=g)}
    visited = {}

    def run(n, ug):
        return node_adj_PC_generation(rep, bd, model, B=Bg, max_num_nodes=Nm, num_class_X=119, n_steps=1, x_init=x0, adj_init=a0,
                                      draws=lambda k, i: tabs[k][i], diffusion_steps=n, use_graph=ug)
    eager = run(steps, False)
    graph = run(steps, True)
    for a, e in zip(graph, eager):
        assert torch.isfinite(a).all() and torch.equal(a, e)
    # teacher forcing: the states the GPU trajectory visits after 5, 10, 19 steps, scored by both implementations
    sde = O.make_dense_sde("VP", 0.2, 1.0, 1000)
    ts = torch.linspace(1.0, 1e-4, 1000)
    _, rep_o, _, flags_o = O.dense_inputs(h3d, b.x[:, 0], b.edge_index, b.edge_attr[:, 0], b.batch)
    for n in (5, 10, 19):
        xs, adjs, _, _ = run(n, False)
        t = torch.full((Bg,), float(ts[n]))
        emb_o = O.embed_3d2d(sd, rep_o, xs.cpu())
        emb_d = model.embed(rep, xs)
        for which, net, s in (("adj", model.edge_score_network, model.sde_adj), ("x", model.node_score_network, model.sde_x)):
            ref = O.score_3d2d(sd, sde, which, emb_o, adjs.cpu(), flags_o, t)
            got = model.get_score_fn(s, net, train=False)(emb_d, adjs, flags, t.to(dev))
            assert_parity(got, ref, f"{which} score after {n} PC steps")


def test_node_network_valid_row_compaction_is_bit_identical(monkeypatch):
    """NodeScoreNetwork_dense: the final MLP evaluated on the valid atom rows only (sde_3d_to_2d._valid_rows) gives the same bits
    as the padded evaluation (`invariant_scorenetwork_dense.py:118-131`), including an isolated (flag 0) atom in the middle of a graph."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from moleculesde_b200.sde_3d_to_2d import NodeScoreNetwork_dense
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    B, Nm, F = 128, 48, 40
    net = NodeScoreNetwork_dense(nfeat=F, depth=2, nhid=16, nout=23).to(dev).eval()
    n = torch.randint(5, 30, (B,))
    flags = (torch.arange(Nm)[None, :] < n[:, None]).float()
    flags[3, 2] = 0.0                                           # node_flags can leave holes (an atom without bonds)
    flags = flags.to(dev)
    adj = torch.rand(B, Nm, Nm, device=dev)
    adj = (adj + adj.transpose(1, 2)) * flags[:, :, None] * flags[:, None, :]
    x = torch.randn(B, Nm, F, device=dev) * flags[:, :, None]
    scale = -1.0 / (0.5 + torch.rand(B, device=dev))
    got = net(x, adj, flags, scale)
    assert net.__dict__["_rows_cache"][2] is not None, "compaction active"
    got2 = net(x, adj, flags, scale)                            # cache hit
    monkeypatch.setenv("MOLSDE_DENSE_NO_COMPACT", "1")
    want = net(x, adj, flags, scale)
    assert torch.equal(got, want) and torch.equal(got2, want)
    assert float(want.abs().max()) > 0 and float((want * (1 - flags[:, :, None])).abs().max()) == 0
