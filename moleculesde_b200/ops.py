"""Thin host wrappers of single C-ABI ops shared by the model modules."""
from __future__ import annotations

from typing import Optional

import torch

from ._abi import check, lib, ptr, require_device, stream_ptr

ACT = {None: 0, "none": 0, "relu": 1, "silu": 2, "ssp": 3}


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, act: Optional[str] = None,
           residual: Optional[torch.Tensor] = None) -> torch.Tensor:
    """Y = act(x @ weight.T + bias) (+ residual) through `molsde_linear` (fp32)."""
    require_device(x)
    x = x.detach().float().contiguous()
    w = weight.detach().float().contiguous()
    b = None if bias is None else bias.detach().float().contiguous()
    M, K = x.shape
    N = w.size(0)
    assert w.size(1) == K
    y = torch.empty(M, N, dtype=torch.float32, device=x.device)
    r = None if residual is None else residual.detach().float().contiguous()
    check(lib().molsde_linear(ptr(x), M, K, K, ptr(w), ptr(b), N, ptr(y), N, ACT[act], ptr(r), N if r is not None else 0,
                              stream_ptr(x)), "linear")
    return y


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
