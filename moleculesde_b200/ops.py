"""Thin host wrappers of single C-ABI ops shared by the model modules."""
from __future__ import annotations

from typing import Optional

import torch

from ._abi import check, lib, ptr, require_device, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "silu": 2, "ssp": 3, "tanh": 4, "elu": 5}
TC_MIN_ROWS = 256  # below this the 128-row tensor-core tile is mostly padding


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
           residual: Optional[torch.Tensor] = None, rowscale: Optional[torch.Tensor] = None,
           out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = act((x @ weight.T + bias) * rowscale[:,None]) (+ residual) through `molsde_linear` (fp32).
    `x` may be any tensor whose last dim is K with unit stride and uniform row stride (views into wider
    buffers are fine); `out` may be such a view too."""
    require_device(x)
    K = x.size(-1)
    assert x.stride(-1) == 1
    x2 = x if x.dim() == 2 else x.reshape(-1, K) if x.is_contiguous() else x.flatten(0, -2)
    M, ldx = x2.size(0), x2.stride(0)
    w = weight.detach().float().contiguous()
    b = None if bias is None else bias.detach().float().contiguous()
    N = w.size(0)
    assert w.size(1) == K and x2.dtype == torch.float32
    if out is None:
        y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    else:
        y = out if out.dim() == 2 else out.flatten(0, -2)
        assert y.size(0) == M and y.size(1) == N and y.stride(1) == 1
    r2, ldr = None, 0
    if residual is not None:
        r2 = residual if residual.dim() == 2 else residual.flatten(0, -2)
        assert r2.stride(1) == 1
        ldr = r2.stride(0)
    rs = None if rowscale is None else rowscale.reshape(-1).contiguous()
    if M >= TC_MIN_ROWS:   # tensor-core path (tcgen05, 3xTF32 = fp32-class accuracy); tiny problems stay on the FFMA kernel
        check(lib().molsde_tc_gemm(M, N, K, x2.data_ptr(), ldx, 1, ptr(w), K, 1, ptr(b), ACT[act], ptr(rs),
                                   None if r2 is None else r2.data_ptr(), ldr, y.data_ptr(), y.stride(0), 0, None, 0, None,
                                   stream_ptr(x)), "tc_gemm")
    else:
        check(lib().molsde_linear(x2.data_ptr(), M, K, ldx, ptr(w), ptr(b), N, y.data_ptr(), y.stride(0), ACT[act],
                                  None if r2 is None else r2.data_ptr(), ldr, ptr(rs), stream_ptr(x)), "linear")
    if out is not None:
        return out
    return y if x.dim() == 2 else y.view(*x.shape[:-1], N)


def gather_rows(table: torch.Tensor, idx: torch.Tensor) -> torch.Tensor:
    require_device(table)
    t = table.detach().float().contiguous()
    idx = idx.contiguous()
    assert idx.dtype == torch.long
    out = torch.empty(idx.numel(), t.size(1), dtype=torch.float32, device=t.device)
    check(lib().molsde_gather_rows(ptr(t), ptr(idx), idx.numel(), t.size(1), ptr(out), stream_ptr(t)), "gather_rows")
    return out


def segment_reduce(x: torch.Tensor, seg_ptr: torch.Tensor, mean: bool) -> torch.Tensor:
    require_device(x)
    x = x.contiguous()
    segs = seg_ptr.numel() - 1
    out = torch.empty(segs, x.size(1), dtype=torch.float32, device=x.device)
    check(lib().molsde_segment_reduce(ptr(x), ptr(seg_ptr), segs, x.size(1), int(mean), ptr(out), stream_ptr(x)),
          "segment_reduce")
    return out
