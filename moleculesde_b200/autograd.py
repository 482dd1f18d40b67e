"""torch.autograd bridge: lets the reference's own training loop (`loss.backward(); optimizer.step()`,
`examples/pretrain_MoleculeSDE.py:148-152`) run unchanged on the molsde_b200 kernels.

Each module forward that is called with gradients enabled becomes ONE `torch.autograd.Function`: its forward records a private
`Tape` (our forward kernels keeping their intermediates), its backward seeds the tape with torch's grad_output and replays it
(our backward kernels), returning the gradients of the tensor inputs and of every parameter.  torch.autograd only stitches the
four modules together (and accumulates `p.grad`); no torch compute op is involved in the layers themselves.

For the kernels' channel-stacked parameter views the parameters of the module are first copied into a scratch flat buffer in
`ParamStore` order (one small copy per parameter and step); `PretrainStep` avoids even that by owning the flat buffer.
"""
from __future__ import annotations

from typing import Callable, Dict, List, Sequence

import torch

from .pretrain import _layout_key
from .tape import Tape, Var


class _TapeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, build: Callable, names: Sequence[str], n_in: int, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        dev = params[0].device if params else inputs[0].device
        tp = Tape(dev)
        # scratch flat copies (data + gradient) in the kernels' layout order
        order = sorted(range(len(names)), key=lambda i: _layout_key(names[i]))
        sizes = [(params[i].numel() + 3) // 4 * 4 for i in order]
        flat = torch.empty(sum(sizes), dtype=torch.float32, device=dev)
        gflat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        P: Dict[str, Var] = {}
        views: List[torch.Tensor] = [None] * len(names)
        off = 0
        for i, n in zip(order, sizes):
            p = params[i]
            d = flat[off:off + p.numel()].view(p.shape)
            d.copy_(p.detach())
            g = gflat[off:off + p.numel()].view(p.shape)
            P[names[i]] = Var(d, bool(p.requires_grad), g)
            views[i] = g
            off += n
        in_vars = [Var(t.detach().float().contiguous(), bool(t.requires_grad)) for t in inputs]
        outs, seed = build(tp, in_vars, P)   # outs: list of tensors; seed(grad_outputs) prepares the tape for the reverse sweep
        ctx.tp, ctx.in_vars, ctx.views, ctx.seed = tp, in_vars, views, seed
        ctx.param_needs = [bool(p.requires_grad) for p in params]
        ctx.n_in = n_in
        return tuple(o.detach() for o in outs) if len(outs) > 1 else outs[0].detach()

    @staticmethod
    def backward(ctx, *gouts):
        if getattr(ctx, "consumed", False):
            raise RuntimeError("moleculesde_b200: this forward was already back-propagated once; the kernel tape frees its "
                               "intermediates in the reverse sweep (retain_graph / double backward are not supported -- run the "
                               "forward again)")
        ctx.consumed = True
        ctx.seed([None if g is None else g.detach().float().contiguous() for g in gouts])
        ctx.tp.backward()
        gin = []
        for v in ctx.in_vars:
            gin.append(v.grad if (v.needs and v.grad is not None) else None)
        gpar = [g if need else None for g, need in zip(ctx.views, ctx.param_needs)]
        return (None, None, None, *gin, *gpar)


def apply(module: torch.nn.Module, build: Callable, inputs: Sequence[torch.Tensor]):
    """Run `build(tp, input_vars, P) -> (outputs, seed)` as one autograd node over `inputs` and the module's parameters."""
    named = [(n, p) for n, p in module.named_parameters()]
    names = [n for n, _ in named]
    return _TapeFn.apply(build, names, len(inputs), *inputs, *[p for _, p in named])


def grad_mode(module: torch.nn.Module, *tensors) -> bool:
    """True when the call should be differentiable: autograd is on and a parameter or an input wants a gradient."""
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())
