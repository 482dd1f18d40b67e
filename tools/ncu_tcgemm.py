"""One launch of each interesting molsde_tc_gemm shape (for `ncu --set full`): forward 5120x728x728, dW 728x728x5120."""
import sys, torch
sys.path.insert(0, "/root/repo")
from moleculesde_b200._abi import lib
dev = torch.device("cuda:0"); L = lib(); s = torch.cuda.current_stream().cuda_stream
M, N, K = 5120, 728, 728
x = torch.randn(M, K, device=dev); w = torch.randn(N, K, device=dev); b = torch.randn(N, device=dev); y = torch.empty(M, N, device=dev)
dy = torch.randn(M, N, device=dev); dw = torch.empty(N, K, device=dev)
n1 = L.molsde_tc_gemm_ws_floats(N, K, M); ws = torch.empty(max(n1, 1), device=dev)
for _ in range(2):
    L.molsde_tc_gemm(M, N, K, x.data_ptr(), K, 1, w.data_ptr(), K, 1, b.data_ptr(), 2, None, None, 0, y.data_ptr(), N, 0, None, 0, None, s)
    L.molsde_tc_gemm(N, K, M, dy.data_ptr(), 1, N, x.data_ptr(), 1, K, None, 0, None, None, 0, dw.data_ptr(), K, 0, ws.data_ptr(), n1, None, s)
torch.cuda.synchronize()
