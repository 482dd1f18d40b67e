import sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
from moleculesde_b200._abi import check, lib
from test_gpu_tcgemm import _tc
dev = torch.device("cuda:0")
L = lib()
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for (M, N, K) in [(16384, 728, 728), (16384, 728, 364), (16384, 128, 300), (5120, 728, 728), (5120, 728, 364), (49000, 128, 128), (49000, 128, 51), (3585, 300, 300), (3585, 600, 300), (38424, 32, 300),
                  (38424, 32, 64), (38424, 128, 64), (102400, 16, 16), (102400, 60, 60), (5120, 32, 300), (5120, 32, 32), (40960, 300, 300)]:
    x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev); y = torch.empty(M, N, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    st = torch.zeros(1, dtype=torch.int32, device=dev)
    t_tc = timeit(lambda: L.molsde_tc_gemm(M, N, K, x.data_ptr(), K, 1, w.data_ptr(), K, 1, b.data_ptr(), 0, None, None, 0, y.data_ptr(), N, 0, None, 0, st.data_ptr(), s))
    t_ff = timeit(lambda: L.molsde_linear(x.data_ptr(), M, K, K, w.data_ptr(), b.data_ptr(), N, y.data_ptr(), N, 0, None, 0, None, s))
    fl = 2.0 * M * N * K
    # weight gradient pattern dW[N,K] = dy[M,N]^T x[M,K]
    dy = torch.randn(M, N, device=dev); dw = torch.empty(N, K, device=dev)
    n1 = L.molsde_tc_gemm_ws_floats(N, K, M); ws1 = torch.empty(max(n1, 1), device=dev)
    n2 = L.molsde_gemm_ws_floats(N, K, M); ws2 = torch.empty(max(n2, 1), device=dev)
    t_tc_w = timeit(lambda: L.molsde_tc_gemm(N, K, M, dy.data_ptr(), 1, N, x.data_ptr(), 1, K, None, 0, None, None, 0, dw.data_ptr(), K, 0, ws1.data_ptr(), n1, st.data_ptr(), s))
    t_ff_w = timeit(lambda: L.molsde_gemm(1, 0, N, K, M, dy.data_ptr(), N, x.data_ptr(), K, dw.data_ptr(), K, 0, ws2.data_ptr(), n2, s))
    print(f"{M:7d}x{N:4d}x{K:4d}  fwd: tc {t_tc:7.1f} us ({fl / t_tc / 1e6:6.1f} TF/s)  ffma {t_ff:7.1f} us ({fl / t_ff / 1e6:5.1f} TF/s) | dW: tc {t_tc_w:7.1f} us  ffma {t_ff_w:7.1f} us")
