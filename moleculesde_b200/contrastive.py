"""`do_CL` / `dual_CL` of `examples/util.py:22-79`: the pretraining metric `EBM_node_dot_prod` (row-wise dots, HBM-bound,
`molsde_ebm_node_dot`) and `InfoNCE_dot_prod` (a dense B x B contraction on the tcgen05 GEMM + a row-softmax kernel).
Differentiable through the autograd bridge when an input requires grad."""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ._abi import check, lib, ptr, require_device, stream_ptr


def do_CL(X: torch.Tensor, Y: torch.Tensor, args, neg_index: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, float]:
    """Reference signature `do_CL(X, Y, args)`; `neg_index` injects the `torch.randperm(len(Y))` draw
    (`util.py:55`, CPU generator in the reference)."""
    if args.CL_similarity_metric not in ("EBM_node_dot_prod", "InfoNCE_dot_prod"):
        raise NotImplementedError("built: EBM_node_dot_prod (the pretraining metric, README.md:86-94) and InfoNCE_dot_prod")
    require_device(X)
    infonce = args.CL_similarity_metric == "InfoNCE_dot_prod"
    if infonce or (torch.is_grad_enabled() and (X.requires_grad or Y.requires_grad)):   # tape path (differentiable when asked)
        from . import autograd as AG
        from .pretrain import tape_cl, tape_infonce
        T = float(args.T)
        holder = {}

        def build(tp, ins, P):
            c = [1.0]
            out = tape_infonce(tp, ins[0], ins[1], T, c) if infonce else tape_cl(tp, ins[0], ins[1], T, neg_index, c)

            def seed(gouts):
                c[0] = float(gouts[0].reshape(-1)[0].item()) if gouts[0] is not None else 0.0
            holder["acc"] = out
            return [out[:1].reshape(())], seed
        loss = AG.apply(torch.nn.Module(), build, [X, Y])
        return loss, float(holder["acc"][1].item())
    X = X.detach().float().contiguous()
    Y = Y.detach().float().contiguous()
    N, D = X.shape
    if neg_index is None:
        neg_index = torch.randperm(N)
    perm = neg_index.to(X.device).long().contiguous()
    pred_pos = torch.empty(N, dtype=torch.float32, device=X.device)
    pred_neg = torch.empty(N, dtype=torch.float32, device=X.device)
    out = torch.empty(2, dtype=torch.float32, device=X.device)
    ws = torch.empty(4 * 592, dtype=torch.float32, device=X.device)
    check(lib().molsde_ebm_node_dot(ptr(X), ptr(Y), ptr(perm), N, D, float(args.T), ptr(pred_pos), ptr(pred_neg), ptr(out),
                                    ptr(ws), ws.numel(), stream_ptr(X)), "ebm_node_dot")
    return out[0], float(out[1].item())  # the reference also syncs for CL_acc (.cpu().item(), util.py:68)


def dual_CL(X, Y, args, neg_index_1=None, neg_index_2=None):
    l1, a1 = do_CL(X, Y, args, neg_index_1)
    l2, a2 = do_CL(Y, X, args, neg_index_2)
    return (l1 + l2) / 2, (a1 + a2) / 2
