"""GPU parity of SchNet forward and the EBM_node_dot_prod contrastive loss against golden / oracle."""
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from conftest import sd_from_manifest  # noqa: E402
from oracle import model as O  # noqa: E402
from test_gpu_sde2d3d import assert_parity, rel_err  # noqa: E402


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _schnet(golden, dev):
    from moleculesde_b200.schnet import SchNet
    m = SchNet(hidden_channels=300, num_filters=128, num_interactions=6, num_gaussians=51, cutoff=10, readout="mean",
               node_class=119)
    mine = {k: (tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()}
    assert mine == golden["manifest"]["schnet"]  # incl. the duplicated mlp / conv.nn keys and atomic_mass f64[119]
    sd = sd_from_manifest(golden["manifest"]["schnet"], golden["meta"]["weight_seed"])
    m.load_state_dict(sd)
    return m.to(dev).eval(), sd


def test_schnet_vs_golden(golden, golden_batch):
    dev = _dev()
    _, batch = golden_batch
    model, _ = _schnet(golden, dev)
    out, h = model(batch.x[:, 0].to(dev), batch.positions.to(dev), batch.batch.to(dev), return_latent=True)
    assert_parity(h, golden["schnet"]["h"], "SchNet node latent vs reference")
    assert_parity(out, golden["schnet"]["out"], "SchNet readout vs reference")
    out2 = model(batch.x[:, 0].to(dev), batch.positions.to(dev), batch.batch.to(dev))
    assert torch.equal(out, out2)  # deterministic


@pytest.mark.parametrize("kind,num,seed", [("pcqm", 96, 5), ("drug", 10, 6), ("drug", 64, 7)])
def test_schnet_vs_oracle(kind, num, seed, golden):
    """Larger batches; on drug-sized molecules the 32-neighbour cap binds."""
    from moleculesde_b200.data import synth_batch
    dev = _dev()
    b = synth_batch(num, seed, kind)
    model, sd = _schnet(golden, dev)
    ref_out, ref_h, ei = O.schnet_forward(sd, b.x[:, 0], b.positions, b.batch, b.num_graphs)
    out, h = model(b.x[:, 0].to(dev), b.positions.to(dev), b.batch.to(dev), return_latent=True)
    assert_parity(h, ref_h, f"SchNet latent {kind}")
    assert_parity(out, ref_out, f"SchNet readout {kind}")


def test_dual_cl_vs_golden(golden):
    from moleculesde_b200.contrastive import dual_CL
    dev = _dev()
    X, Y = golden["gnn"]["h_eval"].to(dev), golden["schnet"]["h"].to(dev)
    n1, n2 = golden["cl"]["neg_index"]
    args = types.SimpleNamespace(CL_similarity_metric="EBM_node_dot_prod", T=0.1)
    loss, acc = dual_CL(X, Y, args, n1, n2)
    assert rel_err(loss.cpu().reshape(1), golden["cl"]["loss"].reshape(1)) < 1e-5
    assert abs(acc - golden["cl"]["acc"].item()) < 1e-6


def test_cl_edge_cases():
    from moleculesde_b200.contrastive import do_CL
    dev = _dev()
    args = types.SimpleNamespace(CL_similarity_metric="EBM_node_dot_prod", T=0.1)
    g = torch.Generator().manual_seed(0)
    for n in (1, 7, 5000):
        X, Y = torch.randn(n, 300, generator=g) * 0.3, torch.randn(n, 300, generator=g) * 0.3
        perm = torch.randperm(n, generator=g)
        ref_loss, ref_acc = O.do_cl_ebm_node_dot_prod(X, Y, 0.1, perm)
        loss, acc = do_CL(X.to(dev), Y.to(dev), args, perm)
        assert rel_err(loss.cpu().reshape(1), ref_loss.reshape(1)) < 2e-5
        assert abs(acc - ref_acc) < 1e-6


@pytest.mark.parametrize("B,D", [(109, 300), (700, 300), (33, 64)])
def test_infonce_dot_prod_loss_acc_and_grads(B, D):
    """do_CL with metric InfoNCE_dot_prod (`examples/util.py:23-32`): logits on the tcgen05 GEMM, CrossEntropy vs the diagonal;
    loss, accuracy and both input gradients against torch (fp64) restating the reference lines."""
    import types
    import torch.nn.functional as F
    from moleculesde_b200.contrastive import do_CL
    dev = _dev()
    g = torch.Generator().manual_seed(B)
    X = (0.2 * torch.randn(B, D, generator=g)).to(dev).requires_grad_(True)
    Y = (0.2 * torch.randn(B, D, generator=g)).to(dev).requires_grad_(True)
    args = types.SimpleNamespace(CL_similarity_metric="InfoNCE_dot_prod", T=0.1)
    loss, acc = do_CL(X, Y, args)
    (loss * 3.0).backward()
    Xr, Yr = X.detach().double().cpu().requires_grad_(True), Y.detach().double().cpu().requires_grad_(True)
    logits = torch.mm(Xr, Yr.t()) / 0.1
    labels = torch.arange(B)
    ref = F.cross_entropy(logits, labels)
    ref_acc = float((logits.argmax(dim=1) == labels).sum()) / B
    (ref * 3.0).backward()
    assert abs(float(loss.detach()) - float(ref.detach())) <= 1e-5 * abs(float(ref.detach()))
    assert abs(acc - ref_acc) < 1e-6
    for got, want in ((X.grad, Xr.grad), (Y.grad, Yr.grad)):
        err = float((got.double().cpu() - want).abs().max() / want.abs().max())
        assert err < 1e-4, err
    with torch.no_grad():
        l2, a2 = do_CL(X.detach(), Y.detach(), args)
    assert float(l2) == float(loss.detach()) and a2 == acc
