import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_tcgemm import _tc, _rel
dev = torch.device("cuda:0")
def run_t(nout, nin, rows, split, ldy=None, ldx=None):
    g = torch.Generator().manual_seed(5)
    ldy = ldy or nout; ldx = ldx or nin
    xb = torch.randn(rows, ldx, generator=g).to(dev); dyb = torch.randn(rows, ldy, generator=g).to(dev)
    x, dy = xb[:, :nin], dyb[:, :nout]
    dw = torch.empty(nout, nin, device=dev)
    _tc(nout, nin, rows, dy, 1, ldy, x, 1, ldx, dw, nin, split=split)
    ref = dy.double().t() @ x.double()
    print("TN", nout, nin, rows, split, ldy, ldx, "rel %.2e" % _rel(dw, ref))
def run_nn(M, N, K):
    g = torch.Generator().manual_seed(5)
    dy = torch.randn(M, K, generator=g).to(dev); w = torch.randn(K, N, generator=g).to(dev)
    y = torch.empty(M, N, device=dev)
    _tc(M, N, K, dy, K, 1, w, 1, N, y, N, split=True)
    print("NN", M, N, K, "rel %.2e" % _rel(y, dy.double() @ w.double()))
for rows in (109, 113, 1126, 13):
    run_t(600, 300, rows, True); run_t(300, 600, rows, True); run_t(300, 600, rows, False); run_t(32, 32, rows, True, 512, 64)
run_nn(109, 300, 600); run_nn(109, 600, 300); run_nn(1126, 32, 300)
