import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from test_gpu_tcgemm import _tc, _rel
dev = torch.device("cuda:0")
def run_t(nout, nin, rows, split):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(rows, nin, generator=g).to(dev); dy = torch.randn(rows, nout, generator=g).to(dev)
    dw = torch.empty(nout, nin, device=dev)
    _tc(nout, nin, rows, dy, 1, nout, x, 1, nin, dw, nin, split=split)
    ref = dy.double().t() @ x.double()
    print("TN", nout, nin, rows, split, "rel %.2e" % _rel(dw, ref))
def run_n(M, N, K):
    g = torch.Generator().manual_seed(5)
    x = torch.randn(M, K, generator=g).to(dev); w = torch.randn(N, K, generator=g).to(dev)
    y = torch.empty(M, N, device=dev)
    _tc(M, N, K, x, K, 1, w, K, 1, y, N, split=False)
    print("NT", M, N, K, "rel %.2e" % _rel(y, x.double() @ w.double().t()))
for rows in (512, 4096, 40000):
    for nin in (51, 64, 128):
        run_t(128, nin, rows, False)
run_t(128, 51, 40000, True)
run_n(128, 51, 40000); run_n(128, 64, 40000); run_n(128, 128, 40000); run_n(128, 52, 40000)
