"""Whole-chain 3-layer MLP training kernels (`csrc/train_mlp.cu`, `Tape.mlp3`) against a torch fp64 reference of the same MLP:
forward value, input gradient and every parameter gradient; and against the layer-granular tape path they replace
(EdgeNetwork_dense.mlp, edge_network_dense.py:120-123).
Tolerance 1e-5 max-norm relative (fp32 FFMA chains vs fp64)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _rel(a, b):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("d0,h,d3,act,rows", [(4, 16, 8, "elu", 5000), (16, 16, 8, "elu", 102400), (16, 16, 4, "elu", 777),
                                               (16, 16, 4, "elu", 1), (4, 16, 8, "elu", 128)])
def test_mlp3_train_matches_fp64(d0, h, d3, act, rows):
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from moleculesde_b200.tape import Tape, Var
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(rows + d0)
    dims = [(h, d0), (h, h), (d3, h)]
    Wt = [(torch.randn(o, i, generator=g) / i ** 0.5).to(dev) for o, i in dims]
    bt = [(0.3 * torch.randn(o, generator=g)).to(dev) for o, _ in dims]
    x = (1.5 * torch.randn(rows, d0, generator=g)).to(dev)
    dy = torch.randn(rows, d3, generator=g).to(dev)

    def run(fused):
        tp = Tape(dev)
        Ws = [Var(w, True, torch.zeros_like(w)) for w in Wt]
        bs = [Var(b, True, torch.zeros_like(b)) for b in bt]
        xv = Var(x, True)
        if fused:
            assert tp.mlp3_supported(xv, Ws, bs, act)
            y = tp.mlp3(xv, Ws, bs, act)
        else:
            y = tp.linear(tp.linear(tp.linear(xv, Ws[0], bs[0], act=act), Ws[1], bs[1], act=act), Ws[2], bs[2])
        y.grad = dy.clone()
        tp.backward()
        torch.cuda.synchronize()
        return y.data, xv.grad, [w.grad for w in Ws], [b.grad for b in bs], tp.launches

    yf, dxf, dWf, dbf, nf = run(True)
    yl, dxl, dWl, dbl, nl = run(False)
    assert nf < nl
    # fp64 reference
    fn = torch.nn.functional.elu
    W64 = [w.double().requires_grad_() for w in Wt]
    b64 = [b.double().requires_grad_() for b in bt]
    x64 = x.double().requires_grad_()
    y64 = fn(fn(x64 @ W64[0].T + b64[0]) @ W64[1].T + b64[1]) @ W64[2].T + b64[2]
    y64.backward(dy.double())
    assert _rel(yf, y64.detach()) <= TOL
    assert _rel(dxf, x64.grad) <= TOL
    for k in range(3):
        assert _rel(dWf[k], W64[k].grad) <= 2e-5, k   # weight gradients: tcgen05 3xTF32 GEMMs over the row dimension (as before)
        assert _rel(dbf[k], b64[k].grad) <= 2e-5, k
    # and the path it replaces agrees at the same level
    assert _rel(yl, y64.detach()) <= 1e-4 and _rel(dxl, x64.grad) <= 1e-4


def test_mlp3_unsupported_dims_fall_back():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    from moleculesde_b200.tape import Tape, Var
    dev = torch.device("cuda:0")
    tp = Tape(dev)
    Ws = [Var(torch.zeros(o, i, device=dev), True) for o, i in ((24, 8), (24, 24), (8, 24))]
    bs = [Var(torch.zeros(o, device=dev), True) for o in (24, 24, 8)]
    assert not tp.mlp3_supported(Var(torch.zeros(10, 8, device=dev)), Ws, bs, "elu")
    assert not tp.mlp3_supported(Var(torch.zeros(10, 16, device=dev)), Ws, bs, "tanh")
    # the 30 -> 60 -> 60 -> 1 silu head of the edge score network was measured slower than its GEMMs and is not instantiated
    Wh = [Var(torch.zeros(o, i, device=dev), True) for o, i in ((60, 30), (60, 60), (1, 60))]
    bh = [Var(torch.zeros(o, device=dev), True) for o in (60, 60, 1)]
    assert not tp.mlp3_supported(Var(torch.zeros(10, 30, device=dev)), Wh, bh, "silu")
