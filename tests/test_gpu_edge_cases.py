"""Edge cases of the 2D->3D kernels: degenerate molecules, size limits, step-wise API vs the fused loop."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import sd_from_manifest  # noqa: E402
from oracle import model as O  # noqa: E402
from oracle import ref_ops as R  # noqa: E402
from test_gpu_sde2d3d import _gpu_batch, _model, assert_parity, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _tiny_molecules():
    from moleculesde_b200.data import Data
    one = Data(x=torch.zeros(1, 9, dtype=torch.long), edge_index=torch.zeros(2, 0, dtype=torch.long),
               edge_attr=torch.zeros(0, 3, dtype=torch.long), positions=torch.zeros(1, 3))
    two = Data(x=torch.zeros(2, 9, dtype=torch.long), edge_index=torch.tensor([[0, 1], [1, 0]]),
               edge_attr=torch.zeros(2, 3, dtype=torch.long), positions=torch.tensor([[0.75, 0.0, 0.0], [-0.75, 0.0, 0.0]]))
    return one, two


def test_get_score_degenerate_molecules(golden):
    """A single atom (no edges: score 0), a diatomic, and regular molecules in one batch."""
    from moleculesde_b200.data import Batch, synth_molecules
    dev = _dev()
    one, two = _tiny_molecules()
    mols = [one, two] + synth_molecules(3, 44, "pcqm") + [one]
    batch = Batch.from_data_list(mols)
    model, sd = _model(golden, "VE", dev)
    b = _gpu_batch(batch, dev)
    ext = torch.cat([R.extend_graph_index(m.edge_index, m.num_nodes) + int(o) for m, o in zip(mols, batch.ptr[:-1])], dim=1)
    assert torch.equal(b.extended_edge_index.cpu(), ext)
    g = torch.Generator().manual_seed(0)
    N = batch.positions.size(0)
    h2d = torch.randn(N, 300, generator=g)
    pos = batch.positions + 0.3 * torch.randn(N, 3, generator=g)
    t = torch.full((N,), 0.37)
    ref = O.get_score_2d3d(sd, O.make_sde("VE", 0.2, 1.0, 1000), h2d, ext, pos, t)
    got = model.get_score(h2d.to(dev), b, pos.to(dev), None, t.to(dev))
    assert torch.all(got[0] == 0) and torch.all(got[-1] == 0)  # isolated atoms: mean over no edges
    assert torch.isfinite(got).all()
    assert_parity(got, ref, "get_score with degenerate molecules")


def test_pc_group_at_size_limit_and_overflow(golden):
    """A sampling group of 11 x 20 = 220 atoms (the fused kernel's limit is 224) runs in one CTA; 12 x 20 = 240 atoms takes the
    step-wise path (score kernel over molecule chunks + per-group update kernels) and follows the same oracle trajectory."""
    import numpy as np
    from moleculesde_b200 import _abi
    from moleculesde_b200.data import repeat_data, synth_molecule
    from moleculesde_b200.sampler import position_PC_generation
    dev = _dev()
    mol = synth_molecule(np.random.default_rng(3), "pcqm", n=20)
    model, sd = _model(golden, "VE", dev)
    rb = repeat_data(mol, 11)
    b = _gpu_batch(rb, dev)
    g = torch.Generator().manual_seed(1)
    N = rb.positions.size(0)
    rep, pos0 = torch.randn(N, 300, generator=g), torch.randn(N, 3, generator=g)
    steps = 2
    nc, npd = torch.randn(steps, N, 3, generator=g), torch.randn(steps, N, 3, generator=g)
    _, pm = position_PC_generation(rep.to(dev), b, pos0.to(dev), model, model.sde_pos, noise_corr=nc.to(dev), noise_pred=npd.to(dev),
                                   diffusion_steps=steps)
    _, ref = O.pc_sample_2d3d(sd, O.make_sde("VE", 0.2, 1.0, 1000), rep, b.extended_edge_index.cpu(), rb.batch, rb.num_graphs, pos0,
                              nc, npd, n_diff_steps=steps)
    assert rel_err(pm.cpu(), ref) < 1e-3
    rb2 = repeat_data(mol, 12)
    big = _gpu_batch(rb2, dev)
    N2 = rb2.positions.size(0)
    assert N2 == 240 > _abi.CHUNK_MAX_NODES
    rep2, pos2 = torch.randn(N2, 300, generator=g), torch.randn(N2, 3, generator=g)
    nc2, npd2 = torch.randn(steps, N2, 3, generator=g), torch.randn(steps, N2, 3, generator=g)
    _, pm2 = position_PC_generation(rep2.to(dev), big, pos2.to(dev), model, model.sde_pos, noise_corr=nc2.to(dev), noise_pred=npd2.to(dev),
                                    diffusion_steps=steps)
    _, ref2 = O.pc_sample_2d3d(sd, O.make_sde("VE", 0.2, 1.0, 1000), rep2, big.extended_edge_index.cpu(), rb2.batch, rb2.num_graphs, pos2,
                               nc2, npd2, n_diff_steps=steps)
    assert rel_err(pm2.cpu(), ref2) < 1e-3


@pytest.mark.parametrize("kind", ["VE", "VP"])
def test_stepwise_api_matches_fused_loop(kind, golden, golden_batch):
    """ReverseDiffusionPredictor / LangevinCorrector.update_fn (reference step-wise API on get_score) == fused kernel."""
    from moleculesde_b200.data import repeat_data
    from moleculesde_b200.sampler import LangevinCorrector, ReverseDiffusionPredictor, position_PC_generation
    dev = _dev()
    mols, _ = golden_batch
    model, _ = _model(golden, kind, dev)
    rb = _gpu_batch(repeat_data(mols[2], 3), dev)
    g = torch.Generator().manual_seed(7)
    N = rb.positions.size(0)
    rep, pos0 = torch.randn(N, 300, generator=g).to(dev), torch.randn(N, 3, generator=g).to(dev)
    steps = 3
    nc, npd = torch.randn(steps, N, 3, generator=g).to(dev), torch.randn(steps, N, 3, generator=g).to(dev)
    _, fused = position_PC_generation(rep, rb, pos0, model, model.sde_pos, noise_corr=nc, noise_pred=npd, diffusion_steps=steps)
    sde = model.sde_pos
    pred, corr = ReverseDiffusionPredictor(sde, model), LangevinCorrector(sde, model, 0.2, 0.9, 1)
    timesteps = torch.linspace(sde.T, 1e-4, sde.N, device=dev)
    pos = pos0
    for i in range(steps):
        vec_t = (torch.ones(rb.num_graphs, device=dev) * timesteps[i]).index_select(0, rb.batch)
        pos, _ = corr.update_fn(rep, rb, pos, vec_t, noise=nc[i])
        pos, pos_mean = pred.update_fn(rep, rb, pos, vec_t, noise=npd[i])
    assert rel_err(fused.cpu(), pos_mean.cpu()) < 1e-4


def test_radius_graph_rejects_oversized_molecule():
    from moleculesde_b200 import _abi
    from moleculesde_b200.graph import radius_graph
    dev = _dev()
    pos = torch.randn(200, 3, device=dev)
    with pytest.raises(_abi.MolsdeError):
        radius_graph(pos, 10.0, torch.zeros(200, dtype=torch.long, device=dev), 1)


def test_empty_edge_list_and_tiny_batches():
    from moleculesde_b200 import graph as G
    from moleculesde_b200.data import Batch
    dev = _dev()
    one, two = _tiny_molecules()
    b = Batch.from_data_list([one, one, two]).to(dev)
    csr = G.extend_graph(b.edge_index, b.batch, b.num_graphs)
    assert csr.rowptr.cpu().tolist() == [0, 0, 0, 1, 2] and csr.col.cpu().tolist() == [3, 2]
    rg = G.radius_graph(b.positions, 10.0, b.batch, b.num_graphs)
    assert rg.edge_index.cpu().tolist() == [[3, 2], [2, 3]]
    b1 = Batch.from_data_list([one]).to(dev)
    assert G.extend_graph(b1.edge_index, b1.batch, 1).num_edges == 0
