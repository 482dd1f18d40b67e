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


def _param_vars(names: Sequence[str], params: Sequence[torch.Tensor], dev):
    """Scratch flat copies (data + zeroed gradient) of the parameters in the kernels' layout order -> (P, gradient views)."""
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
    return P, views


def _record(ctx, build, names, inputs, params):
    dev = params[0].device if params else inputs[0].device
    tp = Tape(dev)
    P, views = _param_vars(names, params, dev)
    in_vars = [Var(t.detach().float().contiguous(), bool(t.requires_grad)) for t in inputs]
    outs, seed = build(tp, in_vars, P)   # outs: list of tensors; seed(grad_outputs) prepares the tape for the reverse sweep
    ctx.tp, ctx.in_vars, ctx.views, ctx.seed = tp, in_vars, views, seed
    ctx.consumed = False
    return outs


class _TapeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, build: Callable, names: Sequence[str], n_in: int, dual, *tensors):
        inputs, params = tensors[:n_in], tensors[n_in:]
        outs = _record(ctx, build, names, inputs, params)
        ctx.build, ctx.names, ctx.inputs, ctx.params, ctx.dual = build, names, inputs, params, dual
        ctx.param_needs = [bool(p.requires_grad) for p in params]
        ctx.n_in = n_in
        return tuple(o.detach() for o in outs) if len(outs) > 1 else outs[0].detach()

    @staticmethod
    def backward(ctx, *gouts):
        if ctx.consumed:
            # a second reverse sweep over the same forward (retain_graph=True: e.g. the energy term of an energy + force loss after
            # the forces were taken): the kernel tape freed its intermediates in the first sweep -> record it again
            _record(ctx, ctx.build, ctx.names, ctx.inputs, ctx.params)
        if torch.is_grad_enabled() and ctx.dual is not None and ctx.n_in > 0:
            # create_graph=True: the input gradients (forces) must themselves be differentiable -- in the output cotangents and
            # in the parameters.  They are produced by a second autograd node whose backward runs the forward-mode tangent.
            gin = _TapeGradFn.apply(ctx, len(gouts), *gouts, *ctx.params)
            gin = gin if isinstance(gin, tuple) else (gin,)
            return (None, None, None, None, *gin, *([None] * len(ctx.params)))
        ctx.consumed = True
        ctx.seed([None if g is None else g.detach().float().contiguous() for g in gouts])
        ctx.tp.backward()
        gin = []
        for v in ctx.in_vars:
            gin.append(v.grad if (v.needs and v.grad is not None) else None)
        gpar = [g if need else None for g, need in zip(ctx.views, ctx.param_needs)]
        return (None, None, None, None, *gin, *gpar)


class _TapeGradFn(torch.autograd.Function):
    """The first-order reverse sweep of a `_TapeFn` node as a DIFFERENTIABLE function  (cotangents g, parameters) -> d inputs =
    J^T g.  Its own backward receives gbar = d L / d (d inputs) and returns
        d L / d g     = J gbar                       (the forward-mode tangent of the outputs along gbar)
        d L / d theta = d/d theta <J(theta) gbar, g>  (reverse sweep through primal + tangent, seeded with g on the tangent outputs)
    through the module's `dual` builder -- `loss = ... + criterion(-grad(E, pos, create_graph=True), F*)` of
    `examples/finetune_MD17.py:66-77`.  Second derivatives with respect to the INPUTS (d^2/dpos^2) are not produced."""

    @staticmethod
    def forward(ctx, owner, n_g: int, *tensors):
        gouts = tensors[:n_g]
        owner.consumed = True
        owner.seed([None if g is None else g.detach().float().contiguous() for g in gouts])
        owner.tp.backward()
        ctx.owner, ctx.gouts, ctx.n_g = owner, [None if g is None else g.detach().float().contiguous() for g in gouts], n_g
        gin = [v.grad if (v.needs and v.grad is not None) else torch.zeros_like(v.data) for v in owner.in_vars]
        return tuple(g.detach() for g in gin) if len(gin) > 1 else gin[0].detach()

    @staticmethod
    def backward(ctx, *gbars):
        o = ctx.owner
        dev = o.params[0].device if o.params else o.inputs[0].device
        tp = Tape(dev)
        P, views = _param_vars(o.names, o.params, dev)
        ins = [Var(t.detach().float().contiguous(), False) for t in o.inputs]
        tans = [None if g is None else g.detach().float().contiguous() for g in gbars]
        outs, out_dots = o.dual(tp, ins, tans, P)            # Vars: primal outputs and their tangents along `tans`
        for od, g in zip(out_dots, ctx.gouts):
            od.grad = None if g is None else g.clone()
        tp.backward()
        ggouts = [od.data if g is not None else None for od, g in zip(out_dots, ctx.gouts)]
        gpar = [v if need else None for v, need in zip(views, o.param_needs)]
        return (None, None, *ggouts, *gpar)


def apply(module: torch.nn.Module, build: Callable, inputs: Sequence[torch.Tensor], dual: Callable = None):
    """Run `build(tp, input_vars, P) -> (outputs, seed)` as one autograd node over `inputs` and the module's parameters.
    `dual(tp, input_vars, input_tangents, P) -> (output Vars, output-tangent Vars)`: optional forward-mode builder that makes the
    node's input gradients differentiable (`create_graph=True`, see `_TapeGradFn`)."""
    named = [(n, p) for n, p in module.named_parameters()]
    names = [n for n, _ in named]
    return _TapeFn.apply(build, names, len(inputs), dual, *inputs, *[p for _, p in named])


def grad_mode(module: torch.nn.Module, *tensors) -> bool:
    """True when the call should be differentiable: autograd is on and a parameter or an input wants a gradient."""
    if not torch.is_grad_enabled():
        return False
    return any(t is not None and t.requires_grad for t in tensors) or any(p.requires_grad for p in module.parameters())
