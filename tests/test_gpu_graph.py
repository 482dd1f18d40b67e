"""GPU parity (bit-exact) of the graph-construction kernels against the oracle, via the C ABI."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import ref_ops as R  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _batch(kind, num, seed):
    from moleculesde_b200.data import synth_batch
    return synth_batch(num, seed, kind)


def test_segment_ptr_and_scan():
    from moleculesde_b200 import graph as G
    dev = _dev()
    b = _batch("pcqm", 37, 5)
    ptr = G.segment_ptr(b.batch.to(dev), b.num_graphs)
    assert torch.equal(ptr.cpu().long(), b.ptr)
    g = torch.Generator().manual_seed(0)
    for n in (0, 1, 5, 1023, 1024, 1025, 40000):
        deg = torch.randint(0, 50, (n,), generator=g, dtype=torch.int32)
        out = G._scan(deg.to(dev)).cpu()
        exp = torch.cat([torch.zeros(1, dtype=torch.int64), deg.long().cumsum(0)])
        assert torch.equal(out.long(), exp)


@pytest.mark.parametrize("kind,num,seed", [("pcqm", 8, 0), ("pcqm", 300, 3), ("drug", 40, 1), ("padded64", 33, 2)])
def test_extend_graph_bit_exact(kind, num, seed, golden):
    from moleculesde_b200 import graph as G
    from moleculesde_b200.data import Batch, synth_molecules
    dev = _dev()
    mols = synth_molecules(num, seed, kind)
    exp = []
    off = 0
    for m in mols:
        exp.append(R.extend_graph_index(m.edge_index, m.num_nodes) + off)
        off += m.num_nodes
    exp = torch.cat(exp, dim=1)
    b = Batch.from_data_list(mols)
    csr = G.extend_graph(b.edge_index.to(dev), b.batch.to(dev), b.num_graphs)
    assert torch.equal(csr.edge_index.cpu(), exp)
    assert torch.equal(csr.col.cpu().long(), exp[1])
    deg = torch.bincount(exp[0], minlength=b.positions.size(0))
    assert torch.equal(csr.rowptr.cpu().long()[1:], deg.cumsum(0))
    if (kind, num, seed) == ("pcqm", 8, 0):
        assert torch.equal(csr.edge_index.cpu(), golden["graph"]["extended_edge_index"])


def test_extend_graph_edge_cases():
    from moleculesde_b200 import graph as G
    from moleculesde_b200.data import Batch, Data
    dev = _dev()
    # single atom without bonds, a 2-atom molecule, and a 5-node path (the demo of dataset_3D.py:164-168)
    path = torch.tensor([[0, 1, 1, 2, 2, 3, 3, 4], [1, 0, 2, 1, 3, 2, 4, 3]])
    mols = [Data(x=torch.zeros(1, 9, dtype=torch.long), edge_index=torch.zeros(2, 0, dtype=torch.long), positions=torch.zeros(1, 3)),
            Data(x=torch.zeros(2, 9, dtype=torch.long), edge_index=torch.tensor([[0, 1], [1, 0]]), positions=torch.zeros(2, 3)),
            Data(x=torch.zeros(5, 9, dtype=torch.long), edge_index=path, positions=torch.zeros(5, 3))]
    b = Batch.from_data_list(mols)
    csr = G.extend_graph(b.edge_index.to(dev), b.batch.to(dev), b.num_graphs)
    exp = torch.cat([R.extend_graph_index(m.edge_index, m.num_nodes) + o for m, o in zip(mols, (0, 1, 3))], dim=1)
    assert torch.equal(csr.edge_index.cpu(), exp)
    # path graph: everything within 4 hops -> complete graph minus the diagonal
    assert csr.num_edges == 0 + 2 + 20


@pytest.mark.parametrize("kind,num,seed,r,cap", [("pcqm", 8, 0, 10.0, 32), ("pcqm", 200, 4, 3.0, 32),
                                                 ("drug", 48, 1, 10.0, 32), ("drug", 16, 2, 4.0, 8)])
def test_radius_graph_bit_exact(kind, num, seed, r, cap, golden):
    from moleculesde_b200 import graph as G
    dev = _dev()
    b = _batch(kind, num, seed)
    exp = R.radius_graph(b.positions, r, b.batch, max_num_neighbors=cap)
    csr = G.radius_graph(b.positions.to(dev), r, b.batch.to(dev), b.num_graphs, max_num_neighbors=cap)
    assert torch.equal(csr.edge_index.cpu(), exp)
    assert torch.equal(csr.col.cpu().long(), exp[0])
    if kind == "drug" and r == 10.0:
        deg = torch.bincount(exp[1])
        assert int(deg.max()) >= 32  # the neighbour cap binds on drug-sized molecules
    if (kind, num, seed) == ("pcqm", 8, 0):
        assert torch.equal(csr.edge_index.cpu(), golden["schnet"]["radius_edge_index"])


def test_csr_by_target_is_stable():
    from moleculesde_b200 import graph as G
    dev = _dev()
    b = _batch("pcqm", 50, 7)
    ei = b.edge_index
    csr = G.csr_by_target(ei.to(dev), b.batch.to(dev), b.num_graphs)
    order = torch.sort(ei[1], stable=True).indices
    assert torch.equal(csr.perm.cpu().long(), order)
    assert torch.equal(csr.col.cpu().long(), ei[0][order])
    deg = torch.bincount(ei[1], minlength=b.positions.size(0))
    assert torch.equal(csr.rowptr.cpu().long()[1:], deg.cumsum(0))
