"""Input pipeline (`moleculesde_b200/loader.py`): thread / hand-over logic on the CPU, and on the GPU that a
batch staged ahead by the loader gives the same bits as the batch prepared inline."""
import pytest
import torch

from moleculesde_b200.data import Batch, synth_molecules
from moleculesde_b200.loader import DeviceLoader, InlineLoader, pin_batch


def _host_batches(n, size=4):
    return [Batch.from_data_list(synth_molecules(size, 40 + s, "pcqm")) for s in range(n)]


@pytest.mark.parametrize("Loader", [DeviceLoader, InlineLoader])
def test_loader_order_prepare_and_max_nodes(Loader):
    hbs = _host_batches(6)
    calls = []

    def prepare(b, max_nodes):
        calls.append(max_nodes)
        b.tag = len(calls)

    got = list(Loader(hbs, torch.device("cpu"), prepare, depth=2))
    assert len(got) == 6 and [b.tag for b in got] == [1, 2, 3, 4, 5, 6], "every batch once, in order"
    for b, hb in zip(got, hbs):
        assert torch.equal(b.x, hb.x) and torch.equal(b.edge_index, hb.edge_index) and b.num_graphs == hb.num_graphs
    assert calls == [int((hb.ptr[1:] - hb.ptr[:-1]).max()) for hb in hbs], "largest molecule from the host offsets"
    assert len(Loader(hbs, torch.device("cpu"))) == 6
    assert [b.num_graphs for b in Loader(hbs[:1], torch.device("cpu"), None, switch_interval=1e-3)] == [hbs[0].num_graphs]


def test_inline_loader_propagates_errors():
    hbs = _host_batches(4)
    seen = []

    def bad(b, max_nodes):
        if len(seen) == 1:
            raise ValueError("boom")

    with pytest.raises(ValueError, match="boom"):
        for b in InlineLoader(hbs, torch.device("cpu"), bad):
            seen.append(b)


def test_loader_propagates_errors_and_stops_early():
    hbs = _host_batches(5)

    calls = []

    def bad(b, max_nodes):    # the third batch fails in the loader thread (a count of calls: independent of the consumer's pace)
        calls.append(1)
        if len(calls) == 3:
            raise ValueError("boom")

    seen = []
    with pytest.raises(ValueError, match="boom"):
        for b in DeviceLoader(hbs, torch.device("cpu"), bad, depth=1):
            seen.append(b)
    assert len(seen) == 2, "the batches staged before the failure are delivered, then the error surfaces in the consumer"
    import threading
    before = threading.active_count()
    for i, b in enumerate(DeviceLoader(hbs, torch.device("cpu"), None, depth=1)):
        if i == 1:
            break
    import time
    time.sleep(0.3)
    assert threading.active_count() <= before, "the loader thread ends when the consumer stops early"


@pytest.mark.gpu
@pytest.mark.parametrize("Loader", [DeviceLoader, InlineLoader])
def test_loader_batches_give_identical_gradients(Loader):
    import bench
    from moleculesde_b200 import graph as G
    from moleculesde_b200.pretrain import PretrainStep
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    dev = torch.device("cuda:0")
    hbs = [Batch.from_data_list(synth_molecules(48, 300 + s, "pcqm")) for s in range(4)]
    ps = PretrainStep(*bench.build_pretrain_models(), dev)
    g = torch.Generator().manual_seed(3)

    def draws_for(hb):
        N, B = hb.positions.size(0), hb.num_graphs
        nmax = int((hb.ptr[1:] - hb.ptr[:-1]).max())
        return {"cl": (torch.randperm(N, generator=g), torch.randperm(N, generator=g)),
                "sde2d3d": {"noise": torch.randn(N, 3, generator=g), "time_step": torch.randint(0, 1000, (B // 2 + 1,), generator=g)},
                "sde3d2d": [torch.randint(0, 1000, (B // 2 + 1,), generator=g), torch.randn(B, nmax, nmax, generator=g),
                            torch.randn(B, nmax, 119, generator=g)]}
    draws = [draws_for(hb) for hb in hbs]
    for m in (ps.gnn, ps.schnet, ps.m23, ps.m32):
        m.train()
    ps.m23.score_network.dropout = 0.0   # no dropout masks to inject: they would need the extended edge count up front
    want = []
    for hb, d in zip(hbs, draws):
        b = hb.to(dev)
        b.extended_edge_index = G.extend_graph(b.edge_index, b.batch, b.num_graphs).edge_index
        out = ps.forward_backward(b, d)
        want.append((ps.store.grad.clone(), float(out["loss_2d3d"]), float(out["loss_adj"])))
    torch.cuda.synchronize()
    n = 0
    for b, d, w in zip(Loader([pin_batch(hb) for hb in hbs], dev, prepare=ps.prepare, depth=2), draws, want):
        assert b.x.is_cuda and getattr(b, "extended_edge_index", None) is not None and "schnet" in b._molsde_train_cache
        assert b._molsde_dense_dims[1] == int((hbs[n].ptr[1:] - hbs[n].ptr[:-1]).max())
        out = ps.forward_backward(b, d)
        torch.cuda.synchronize()
        assert torch.equal(ps.store.grad, w[0]), f"batch {n}: staged-ahead batch == inline batch, bit for bit"
        assert float(out["loss_2d3d"]) == w[1] and float(out["loss_adj"]) == w[2]
        n += 1
    assert n == len(hbs)
