"""tcgen05 3xTF32 GEMM (`molsde_tc_gemm`) against an fp64 reference: forward / dx / dW access patterns, ragged sizes,
strided and misaligned views, epilogue (bias, rowscale, activation, residual), split-K, accumulate."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    return torch.device("cuda:0")


def _tc(M, N, K, A, sam, sak, B, sbn, sbk, C, ldc, bias=None, act=0, rowscale=None, R=None, accumulate=False, split=True):
    from moleculesde_b200._abi import check, lib
    L = lib()
    n = L.molsde_tc_gemm_ws_floats(M, N, K) if split else 0
    ws = torch.empty(max(n, 1), dtype=torch.float32, device=C.device)
    status = torch.zeros(1, dtype=torch.int32, device=C.device)
    p = lambda t: None if t is None else t.data_ptr()
    check(L.molsde_tc_gemm(M, N, K, p(A), sam, sak, p(B), sbn, sbk, p(bias), act, p(rowscale), p(R), 0 if R is None else R.stride(0),
                           p(C), ldc, int(accumulate), p(ws) if n else None, n, p(status), torch.cuda.current_stream().cuda_stream),
          "tc_gemm")
    torch.cuda.synchronize()
    assert int(status) == 0, "mbarrier wait timed out"


def _rel(a, b):
    return float((a.double() - b).abs().max() / b.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (5120, 728, 364), (3585, 300, 300), (49000, 128, 51), (777, 119, 728),
                                   (100, 32, 16), (38424, 32, 300), (1, 1, 1), (130, 65, 33)])
def test_forward_linear_shapes(M, N, K):
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).to(dev)
    w = torch.randn(N, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    y = torch.empty(M, N, device=dev)
    _tc(M, N, K, x, K, 1, w, K, 1, y, N, bias=b)
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(y, ref) < 5e-6, _rel(y, ref)


def test_epilogue_and_strided_views():
    dev = _dev()
    g = torch.Generator().manual_seed(3)
    M, N, K = 1000, 48, 70
    xbig = torch.randn(M, K + 9, generator=g).to(dev)
    x = xbig[:, 5:5 + K]                       # misaligned row-strided view
    w = torch.randn(N, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    rs = torch.rand(M, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    out_big = torch.zeros(M, N + 7, device=dev)
    y = out_big[:, 3:3 + N]
    _tc(M, N, K, x, x.stride(0), 1, w, K, 1, y, y.stride(0), bias=b, act=4, rowscale=rs, R=res)
    ref = torch.tanh((x.double() @ w.double().t() + b.double()) * rs.double()[:, None]) + res.double()
    assert _rel(y, ref) < 5e-6
    assert float(out_big[:, :3].abs().max()) == 0 and float(out_big[:, 3 + N:].abs().max()) == 0


def test_backward_patterns_split_k_and_accumulate():
    dev = _dev()
    g = torch.Generator().manual_seed(5)
    rows, nin, nout = 40000, 51, 128
    x = torch.randn(rows, nin, generator=g).to(dev)
    dy = torch.randn(rows, nout, generator=g).to(dev)
    w = torch.randn(nout, nin, generator=g).to(dev)
    # dx = dy . W : A = dy [rows, nout], B(n = kin, k = nout) = W[k, n]
    dx = torch.empty(rows, nin, device=dev)
    _tc(rows, nin, nout, dy, nout, 1, w, 1, nin, dx, nin)
    assert _rel(dx, dy.double() @ w.double()) < 5e-6
    # dW = dy^T . x : A(m = nout, k = row) = dy[k, m], B(n = kin, k = row) = x[k, n]; split-K over the rows, accumulate
    dw = torch.ones(nout, nin, device=dev)
    _tc(nout, nin, rows, dy, 1, nout, x, 1, nin, dw, nin, accumulate=True)
    ref = dy.double().t() @ x.double() + 1.0
    assert _rel(dw, ref) < 5e-6
    dw2 = torch.empty(nout, nin, device=dev)
    _tc(nout, nin, rows, dy, 1, nout, x, 1, nin, dw2, nin, split=False)
    assert _rel(dw2, ref - 1.0) < 5e-6


@pytest.mark.parametrize("rows,nin,nout", [(40000, 51, 128), (5120, 32, 32), (3585, 300, 600), (777, 64, 64), (102400, 16, 16)])
def test_fused_weight_and_bias_gradient(rows, nin, nout):
    """dW = dy^T x and db = colsum(dy) from one GEMM (all-ones extra operand row), accumulate semantics, split-K."""
    from moleculesde_b200._abi import check, lib
    dev = _dev()
    L = lib()
    g = torch.Generator().manual_seed(rows + nin)
    x = torch.randn(rows, nin, generator=g).to(dev)
    dy = torch.randn(rows, nout, generator=g).to(dev)
    dw = torch.full((nout, nin), 2.0, device=dev)
    db = torch.full((nout,), -1.0, device=dev)
    n = L.molsde_tc_gemm_ws_floats(nout, nin + 1, rows)
    ws = torch.empty(max(n, 1), device=dev)
    check(L.molsde_tc_gemm_dw_db(nout, nin, rows, dy.data_ptr(), 1, nout, x.data_ptr(), 1, nin, dw.data_ptr(), nin, db.data_ptr(), 1,
                                 ws.data_ptr() if n else None, n, None, torch.cuda.current_stream().cuda_stream), "dw_db")
    torch.cuda.synchronize()
    assert _rel(dw, dy.double().t() @ x.double() + 2.0) < 5e-6
    assert _rel(db, dy.double().sum(0) - 1.0) < 5e-6


@pytest.mark.parametrize("M,N,K", [(16384, 728, 728), (16384, 728, 364), (4096, 300, 300), (1000, 200, 128), (257, 65, 64), (3607, 600, 300)])
def test_tma_fed_path_shapes_and_epilogue(M, N, K):
    """Shapes that take the TMA-fed kernel (K-contiguous operands, 16-byte aligned rows, N > 64, M >= 256, no split-K): ragged
    M / N / K tails are zero-filled by the tensor maps; epilogue with bias, rowscale, activation, residual, strided output."""
    dev = _dev()
    g = torch.Generator(device="cpu").manual_seed(M * 3 + N + K)
    xbig = torch.randn(M, K + 8, generator=g).to(dev)
    x = xbig[:, 4:4 + K]                       # 16-byte aligned row-strided view (ld = K + 8)
    w = torch.randn(N, K, generator=g).to(dev)
    b = torch.randn(N, generator=g).to(dev)
    rs = torch.rand(M, generator=g).to(dev)
    res = torch.randn(M, N, generator=g).to(dev)
    y = torch.empty(M, N, device=dev)
    _tc(M, N, K, x, x.stride(0), 1, w, K, 1, y, N, bias=b, split=False)
    ref = x.double() @ w.double().t() + b.double()
    assert _rel(y, ref) < 5e-6, _rel(y, ref)
    out_big = torch.zeros(M, N + 5, device=dev)
    y2 = out_big[:, 2:2 + N]
    _tc(M, N, K, x, x.stride(0), 1, w, K, 1, y2, y2.stride(0), bias=b, act=2, rowscale=rs, R=res, split=False)
    pre = ref * rs.double()[:, None]
    ref2 = pre * torch.sigmoid(pre) + res.double()
    assert _rel(y2, ref2) < 5e-6, _rel(y2, ref2)
    assert float(out_big[:, :2].abs().max()) == 0 and float(out_big[:, 2 + N:].abs().max()) == 0
    y3 = torch.empty(M, N, device=dev)       # bit-reproducible run to run
    _tc(M, N, K, x, x.stride(0), 1, w, K, 1, y3, N, bias=b, split=False)
    assert torch.equal(y, y3)
