"""Reverse-mode tape over the C-ABI training kernels (`csrc/train*.cu`).

The reference obtains gradients from torch.autograd over PyTorch/PyG ops (`pretrain_MoleculeSDE.py:150`).
Here every forward op is one of our CUDA kernels and records a closure that launches the matching backward
kernel(s); `Tape.backward` runs the closures in reverse.  torch only owns the memory (`torch.empty`) and the
stream — no torch compute op, no torch.autograd.  Gradient accumulation is `molsde_ew` (a += b), parameter
gradients land in views of one flat buffer (the unit of the NCCL all-reduce and of the flat Adam step).
"""
from __future__ import annotations

from contextlib import contextmanager
from typing import Callable, List, Optional

import torch

from . import _abi
from ._abi import check, lib, ptr, require_device, stream_ptr

ACT = {"none": 0, "relu": 1, "silu": 2, "ssp": 3, "tanh": 4, "elu": 5}
import os as _os
# M*N*K below which a GEMM stays on the FFMA kernels (tensor-core tiles would be mostly padding); MOLSDE_NO_TC=1 disables
TC_MIN_WORK = (1 << 62) if _os.environ.get("MOLSDE_NO_TC") == "1" else int(_os.environ.get("MOLSDE_TC_MIN_WORK", 1 << 20))
FUSED_DB = _os.environ.get("MOLSDE_NO_FUSED_DB") != "1"   # bias gradient through the all-ones row of the dW GEMM
_FUSE_ACT = _os.environ.get("MOLSDE_NO_FUSED_ACT") != "1"  # A/B switch: activation in the GEMM epilogue, derivative from the output
_ACT_FROM_Y = (1, 3, 4, 5)                                 # relu, ssp, tanh, elu (molsde_act_bwd_y)


_CHECK = _abi.CHECK_ABI


def _p(t: Optional[torch.Tensor]) -> Optional[int]:
    """data pointer of a contiguous tensor or of a 2-D row-strided view (stride(1) == 1)."""
    if t is None:
        return None
    if _CHECK and not t.is_contiguous():
        assert t.dim() == 2 and t.stride(1) == 1, "only contiguous tensors and row-strided 2-D views cross the C ABI"
    return t.data_ptr()


def _ld(t: torch.Tensor) -> int:
    return t.stride(0) if t.dim() == 2 else t.numel()


class Var:
    """A tensor on the tape: `data`, its gradient `grad` (allocated lazily) and whether anything upstream needs it."""
    __slots__ = ("data", "grad", "needs")

    def __init__(self, data: torch.Tensor, needs: bool = False, grad: Optional[torch.Tensor] = None):
        self.data, self.needs, self.grad = data, needs, grad

    @property
    def shape(self):
        return self.data.shape


class Index:
    """An int32 row-index vector together with the CSR of its inverse (`ptr`, `perm`): the backward of a gather
    `out[r] = x[idx[r]]` is the deterministic gather-reduce `dx[s] = sum_{p in ptr[s]..ptr[s+1]} dout[perm[p]]`.
    `perm=None` means the index is already grouped (ascending), i.e. perm = identity."""

    def __init__(self, idx: torch.Tensor, ptr_: torch.Tensor, perm: Optional[torch.Tensor], size: int):
        self.idx, self.ptr, self.perm, self.size = idx, ptr_, perm, size


def bucket_index(keys: torch.Tensor, size: int) -> Index:
    """Index for arbitrary int64 keys in [0,size) via the stable counting sort kernels."""
    require_device(keys)
    L, s = lib(), stream_ptr(keys)
    n = keys.numel()
    count = torch.empty(size, dtype=torch.int32, device=keys.device)
    check(L.molsde_bucket_count(ptr(keys), n, size, ptr(count), s), "bucket_count")
    rowptr = torch.empty(size + 1, dtype=torch.int32, device=keys.device)
    check(L.molsde_exclusive_scan_i32(ptr(count), size, ptr(rowptr), s), "scan")
    perm = torch.empty(max(n, 1), dtype=torch.int32, device=keys.device)
    check(L.molsde_bucket_fill(ptr(keys), n, size, ptr(rowptr), ptr(perm), s), "bucket_fill")
    return Index(keys.to(torch.int32), rowptr, perm, size)


class Tape:
    def __init__(self, device: torch.device, wstream: Optional["torch.cuda.Stream"] = None):
        """`wstream`: optional side stream for PARAMETER gradients.  In the backward pass only the input-gradient chain is
        sequential; dW / db / table gradients are leaves of the dependency graph, so they are forked onto `wstream` (after the
        main stream produced dy) and joined once at the end of `backward()` — under CUDA-graph replay they fill the SMs the
        short dx kernels leave idle.  Operands the side kernels read are kept referenced until the join."""
        self.dev = device
        self.ops: List[Callable[[], None]] = []
        self.L = lib()
        self.launches = 0
        self._main = torch.cuda.current_stream(device) if device.type == "cuda" else None
        self._s = self._main.cuda_stream if self._main is not None else 0  # one stream per tape (+ the optional wstream)
        # `wstream` may be a list of side streams: consecutive leaves rotate over them (they are independent of each other too)
        ws = list(wstream) if isinstance(wstream, (list, tuple)) else ([wstream] if wstream is not None else [])
        self._ws = [w for w in ws if self._main is not None and w != self._main]
        self._wi = 0
        self._w = self._ws[0] if self._ws else None
        self._hold: list = []
        self._w_used = False

    # ------------------------------------------------------------------ helpers
    @property
    def s(self) -> int:
        return self._s

    def empty(self, *shape, dtype=torch.float32) -> torch.Tensor:
        return torch.empty(*shape, dtype=dtype, device=self.dev)

    def var(self, data: torch.Tensor, needs: bool = False) -> Var:
        return Var(data, needs)

    def param(self, data: torch.Tensor, grad_view: torch.Tensor) -> Var:
        return Var(data, True, grad_view)

    def _call(self, fn, *args, what=""):
        self.launches += 1
        st = fn(*args)
        if st:
            check(st, what or fn.__name__)

    def ew(self, op, a, b, c, alpha, out, cols=1):
        self._call(self.L.molsde_ew, op, _p(a), _p(b), _p(c), float(alpha), out.numel(), cols, _p(out), self.s, what="ew")

    def accum(self, v: Var, g: torch.Tensor) -> None:
        """v.grad += g (g is a fresh contiguous tensor this op owns)."""
        if not v.needs:
            return
        if v.grad is None:
            v.grad = g if g.is_contiguous() and g.shape == v.data.shape else None
            if v.grad is None:
                v.grad = self.empty(v.data.shape)
                self.ew(0, g.contiguous(), None, None, 1.0, v.grad)
        else:
            assert v.grad.is_contiguous()
            self.ew(0, v.grad, g, None, 1.0, v.grad)

    def grad_of(self, v: Var) -> torch.Tensor:
        if v.grad is None:  # nothing flowed into it
            v.grad = torch.zeros(v.data.shape, dtype=torch.float32, device=self.dev)
        return v.grad

    @contextmanager
    def wgrad(self, *operands, key: Optional[torch.Tensor] = None):
        """Run the enclosed launches (a parameter-gradient leaf) on a side stream, ordered after everything issued so far on
        the tape's stream.  `operands`: tensors read by those launches that could otherwise be freed before the join.
        `key`: the gradient tensor the leaf ACCUMULATES into.  With several side streams the stream is a function of the key's
        address, so every leaf that adds to the same parameter gradient (a weight used twice, e.g. `coff_mlp` on both edge ends)
        runs on the same stream in issue order -- no concurrent read-modify-write, and the same summation order as the eager path."""
        if self._w is None:
            yield
            return
        self._hold.extend(t for t in operands if t is not None)
        n = len(self._ws)
        w = self._ws[0 if (key is None or n == 1) else ((key.data_ptr() >> 4) * 2654435761 >> 12) % n]
        self._wi = n   # (join waits for every side stream)
        w.wait_stream(self._main)
        prev, self._s, self._w_used = self._s, w.cuda_stream, True
        try:
            with torch.cuda.stream(w):   # workspaces allocated inside come from the side stream's pool
                yield
        finally:
            self._s = prev

    def join(self) -> None:
        if self._w_used:
            for w in self._ws[:max(1, min(self._wi, len(self._ws)))]:
                self._main.wait_stream(w)
            self._w_used = False
            self._wi = 0
        self._hold.clear()

    def backward(self) -> None:
        for fn in reversed(self.ops):
            fn()
        self.join()
        self.ops.clear()

    # ------------------------------------------------------------------ dense ops
    def gemm(self, ta, tb, M, N, K, A, lda, B, ldb, C, ldc, accumulate=False):
        """C[M,N] (+)= op(A) op(B);  ta: A stored [K][lda];  tb: B stored [N][ldb].  tcgen05 (3xTF32) unless the problem is tiny."""
        if M * N * K >= TC_MIN_WORK:
            n = self.L.molsde_tc_gemm_ws_floats(M, N, K)
            ws = self.empty(n) if n > 0 else None
            sam, sak = (1, lda) if ta else (lda, 1)
            sbn, sbk = (ldb, 1) if tb else (1, ldb)
            self._call(self.L.molsde_tc_gemm, M, N, K, _p(A), sam, sak, _p(B), sbn, sbk, None, 0, None, None, 0, _p(C), ldc,
                       int(accumulate), _p(ws), n, None, self.s, what="tc_gemm")
            return
        n = self.L.molsde_gemm_ws_floats(M, N, K)
        ws = self.empty(n) if n > 0 else None
        self._call(self.L.molsde_gemm, ta, tb, M, N, K, _p(A), lda, _p(B), ldb, _p(C), ldc, int(accumulate), _p(ws), n, self.s,
                   what="gemm")

    def colsum(self, X, M, N, ldx, out, accumulate=False):
        n = self.L.molsde_colsum_ws_floats(M, N)
        ws = self.empty(n)
        self._call(self.L.molsde_colsum, _p(X), M, N, ldx, _p(out), int(accumulate), _p(ws), n, self.s, what="colsum")

    def linear(self, x: Var, W: Var, b: Optional[Var], act: str = "none", into: Optional[Var] = None, col0: int = 0,
               rowscale: Optional[torch.Tensor] = None, x_cols: Optional[tuple] = None, exact: bool = False) -> Var:
        """y = act(rowscale * (x W^T + b)); W [out,in] as nn.Linear.  `into`/`col0`: write y into columns
        [col0, col0+out) of the wider buffer Var `into` (a concat without a copy); its gradient is read from there.
        `x_cols=(a,b)`: the input is the column slice x[:, a:b]; its gradient is accumulated into that slice of x.grad.
        `exact`: forward on the fp32 FFMA kernel (round-to-nearest adds) instead of tcgen05 3xTF32 — used where the output
        feeds BatchNorm + ReLU, whose sign decisions near zero amplify the ~1e-6 tensor-core error into O(1/rows) gradient
        changes (the backward GEMMs are continuous and stay on the tensor cores)."""
        if x_cols is not None:
            full = x
            x = Var(full.data[:, x_cols[0]:x_cols[1]], False)
        M, K = x.data.shape
        Nout = W.data.shape[0]
        a = ACT[act]
        if into is not None:
            assert a == 0 and rowscale is None
            y = into.data[:, col0:col0 + Nout]
        else:
            y = self.empty(M, Nout)
        use_tc = M * Nout * K >= TC_MIN_WORK and not exact
        # relu / ssp / tanh / elu: f'(pre) is a closed form of the OUTPUT, so the activation runs in the GEMM epilogue and only y is
        # kept (one launch and one [M, Nout] tensor less; SiLU keeps its pre-activation)
        act_in_gemm = a in _ACT_FROM_Y and _FUSE_ACT and (use_tc or W.data.is_contiguous())
        pre = y if (a == 0 or act_in_gemm) else self.empty(M, Nout)
        a_gemm = a if act_in_gemm else 0
        if use_tc:  # tensor cores; W may be a column slice of a wider weight (ld = its row stride)
            self._call(self.L.molsde_tc_gemm, M, Nout, K, _p(x.data), _ld(x.data), 1, _p(W.data), _ld(W.data), 1,
                       _p(b.data) if b is not None else None, a_gemm, _p(rowscale), None, 0, _p(pre), _ld(pre), 0, None, 0, None, self.s,
                       what="tc_gemm")
        elif W.data.is_contiguous():
            self._call(self.L.molsde_linear, _p(x.data), M, K, _ld(x.data), _p(W.data), _p(b.data) if b is not None else None, Nout,
                       _p(pre), _ld(pre), a_gemm, None, 0, _p(rowscale), self.s, what="linear")
        else:
            assert rowscale is None and pre.is_contiguous()
            n = self.L.molsde_gemm_ws_floats(M, Nout, K)
            ws = self.empty(n) if n > 0 else None
            self._call(self.L.molsde_gemm, 0, 1, M, Nout, K, _p(x.data), _ld(x.data), _p(W.data), _ld(W.data), _p(pre), Nout, 0,
                       _p(ws), n, self.s, what="gemm")
            if b is not None:
                self.ew(3, pre, b.data, None, 1.0, pre, cols=Nout)
        if a and not act_in_gemm:
            self._call(self.L.molsde_act_fwd, _p(pre), pre.numel(), a, _p(y), self.s, what="act_fwd")
        x_needs = full.needs if x_cols is not None else x.needs
        needs = x_needs or W.needs or (b is not None and b.needs)
        out = into if into is not None else Var(y, needs)
        if into is not None:
            into.needs = into.needs or needs
        if needs:
            def bwd():
                if into is not None:
                    dy = self.grad_of(into)[:, col0:col0 + Nout]
                else:
                    if out.grad is None:
                        return
                    dy = out.grad
                if a:
                    dpre = self.empty(M, Nout)
                    if act_in_gemm:
                        self._call(self.L.molsde_act_bwd_y, _p(y), _p(dy), dpre.numel(), a, _p(dpre), self.s, what="act_bwd_y")
                    else:
                        self._call(self.L.molsde_act_bwd, _p(pre), _p(dy), dpre.numel(), a, _p(dpre), self.s, what="act_bwd")
                else:
                    dpre = dy
                if rowscale is not None:
                    t = self.empty(M, Nout)
                    self.ew(2, dpre, rowscale, None, 1.0, t, cols=Nout)
                    dpre = t
                self.linear_wgrad(dpre, x.data, W, b)
                if x_cols is not None:
                    if full.needs:
                        g = self.grad_of(full)[:, x_cols[0]:x_cols[1]]
                        self.gemm(0, 0, M, K, Nout, dpre, _ld(dpre), W.data, _ld(W.data), g, _ld(g), accumulate=True)
                elif x.needs:
                    if x.grad is not None and x.grad.is_contiguous() and x.grad.shape == x.data.shape:
                        # another consumer already produced a gradient: accumulate in the GEMM epilogue (no extra add pass)
                        self.gemm(0, 0, M, K, Nout, dpre, _ld(dpre), W.data, _ld(W.data), x.grad, K, accumulate=True)
                    else:
                        dx = self.empty(M, K)
                        self.gemm(0, 0, M, K, Nout, dpre, _ld(dpre), W.data, _ld(W.data), dx, K)
                        self.accum(x, dx)
            self.ops.append(bwd)
        return out

    def linear_wgrad(self, dpre: torch.Tensor, xd: torch.Tensor, W: Var, b: Optional[Var]) -> None:
        """Parameter gradients of y = x W^T + b on the side stream: W.grad += dpre^T x, b.grad += column sums of dpre."""
        M, K = xd.shape
        Nout = dpre.shape[1]
        with self.wgrad(dpre, xd, key=W.grad if W.needs else (b.grad if b is not None and b.needs else None)):
            if FUSED_DB and W.needs and b is not None and b.needs and Nout * K * M >= TC_MIN_WORK:
                # dW and db in one tensor-core GEMM (db = the product with an all-ones extra row)
                n = self.L.molsde_tc_gemm_ws_floats(Nout, K + 1, M)
                ws = self.empty(n) if n > 0 else None
                self._call(self.L.molsde_tc_gemm_dw_db, Nout, K, M, _p(dpre), 1, _ld(dpre), _p(xd), 1, _ld(xd),
                           _p(W.grad), _ld(W.grad), _p(b.grad), 1, _p(ws), n, None, self.s, what="tc_gemm_dw_db")
            else:
                if W.needs:
                    self.gemm(1, 0, Nout, K, M, dpre, _ld(dpre), xd, _ld(xd), W.grad, _ld(W.grad), accumulate=True)
                if b is not None and b.needs:
                    self.colsum(dpre, M, Nout, _ld(dpre), b.grad, accumulate=True)

    _mlp3_ok: dict = {}

    def mlp3_supported(self, x: Var, Ws, bs, act: str) -> bool:
        """True when `molsde_mlp3_train_*` has an instantiation for these three layers (narrow, contiguous weights, all biases)."""
        if len(Ws) != 3 or any(b is None for b in bs) or ACT[act] not in (2, 5) or x.data.dim() != 2 or x.data.stride(1) != 1:
            return False
        d0, h, d3 = x.data.shape[1], Ws[0].data.shape[0], Ws[2].data.shape[0]
        if Ws[0].data.shape != (h, d0) or Ws[1].data.shape != (h, h) or Ws[2].data.shape != (d3, h):
            return False
        if not all(w.data.is_contiguous() for w in Ws):
            return False
        key = (d0, h, d3, ACT[act])
        ok = Tape._mlp3_ok.get(key)
        if ok is None:
            ok = Tape._mlp3_ok[key] = bool(self.L.molsde_mlp3_train_supported(*key))
        return ok

    def mlp3(self, x: Var, Ws, bs, act: str) -> Var:
        """y = W3 act(W2 act(W1 x + b1) + b2) + b3 as ONE forward and ONE input-gradient launch (csrc/train_mlp.cu); the three
        weight-gradient GEMMs stay leaves on the side stream.  Call only when `mlp3_supported`."""
        rows, d0 = x.data.shape
        h, d3 = Ws[0].data.shape[0], Ws[2].data.shape[0]
        a = ACT[act]
        p1, p2, y = self.empty(rows, h), self.empty(rows, h), self.empty(rows, d3)
        self._call(self.L.molsde_mlp3_train_fwd, _p(x.data), rows, _ld(x.data), d0, h, d3, a, _p(Ws[0].data), _p(bs[0].data),
                   _p(Ws[1].data), _p(bs[1].data), _p(Ws[2].data), _p(bs[2].data), _p(p1), _p(p2), _p(y), self.s, what="mlp3_train_fwd")
        out = Var(y, True)

        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            a1, a2, d1, d2 = self.empty(rows, h), self.empty(rows, h), self.empty(rows, h), self.empty(rows, h)
            dx = self.empty(rows, d0) if x.needs else None
            self._call(self.L.molsde_mlp3_train_bwd, _p(p1), _p(p2), _p(dy), rows, d0, h, d3, a, _p(Ws[0].data), _p(Ws[1].data),
                       _p(Ws[2].data), _p(a1), _p(a2), _p(d1), _p(d2), _p(dx), d0, self.s, what="mlp3_train_bwd")
            self.linear_wgrad(dy, a2, Ws[2], bs[2])
            self.linear_wgrad(d2, a1, Ws[1], bs[1])
            self.linear_wgrad(d1, x.data, Ws[0], bs[0])
            if dx is not None:
                self.accum(x, dx)
        self.ops.append(bwd)
        return out

    def matmul(self, x: Var, Wio: Var, into: Optional[Var] = None, col0: int = 0) -> Var:
        """y = x @ W with W stored [in, out] (NodeNetwork_dense.weight, node_network_dense.py:33,73)."""
        M, K = x.data.shape
        Nout = Wio.data.shape[1]
        y = into.data[:, col0:col0 + Nout] if into is not None else self.empty(M, Nout)
        self.gemm(0, 0, M, Nout, K, x.data, _ld(x.data), Wio.data, _ld(Wio.data), y, _ld(y))
        needs = x.needs or Wio.needs
        out = into if into is not None else Var(y, needs)
        if into is not None:
            into.needs = into.needs or needs
        if needs:
            def bwd():
                if into is not None:
                    dy = self.grad_of(into)[:, col0:col0 + Nout]
                else:
                    if out.grad is None:
                        return
                    dy = out.grad
                if Wio.needs:
                    with self.wgrad(dy, x.data, key=Wio.grad):
                        self.gemm(1, 0, K, Nout, M, x.data, _ld(x.data), dy, _ld(dy), Wio.grad, _ld(Wio.grad), accumulate=True)
                if x.needs:
                    dx = self.empty(M, K)
                    self.gemm(0, 1, M, K, Nout, dy, _ld(dy), Wio.data, _ld(Wio.data), dx, K)
                    self.accum(x, dx)
            self.ops.append(bwd)
        return out

    def copy_cols(self, src: Var, dst: Var, col0: int) -> None:
        """dst[:, col0:col0+w] = src (a concat slot); backward adds that slice of dst.grad to src.grad."""
        rows, w = src.data.shape
        d = dst.data[:, col0:col0 + w]
        self._call(self.L.molsde_copy2d, _p(src.data), _ld(src.data), _p(d), _ld(d), rows, w, 0, self.s, what="copy2d")
        dst.needs = dst.needs or src.needs
        if src.needs:
            def bwd():
                g = self.empty(rows, w)
                gs = self.grad_of(dst)[:, col0:col0 + w]
                self._call(self.L.molsde_copy2d, _p(gs), _ld(gs), _p(g), w, rows, w, 0, self.s, what="copy2d")
                self.accum(src, g)
            self.ops.append(bwd)

    def act(self, x: Var, act: str) -> Var:
        a = ACT[act]
        y = self.empty(x.data.shape)
        self._call(self.L.molsde_act_fwd, _p(x.data), y.numel(), a, _p(y), self.s, what="act_fwd")
        out = Var(y, x.needs)
        if x.needs:
            def bwd():
                if out.grad is None:
                    return
                dx = self.empty(x.data.shape)
                self._call(self.L.molsde_act_bwd, _p(x.data), _p(out.grad), dx.numel(), a, _p(dx), self.s, what="act_bwd")
                self.accum(x, dx)
            self.ops.append(bwd)
        return out

    def act_tangent(self, pre: Var, xdot: Var, act: str) -> Var:
        """Forward-mode tangent of y = act(pre): ydot = act'(pre) * xdot, differentiable in BOTH arguments
        (d pre += act''(pre) * xdot * dydot, d xdot = act'(pre) * dydot) -- the op the double backward of SchNet is built from."""
        a = ACT[act]
        y = self.empty(pre.data.shape)
        self._call(self.L.molsde_act_bwd, _p(pre.data), _p(xdot.data), y.numel(), a, _p(y), self.s, what="act_bwd")
        out = Var(y, pre.needs or xdot.needs)
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                if xdot.needs:
                    g = self.empty(y.shape)
                    self._call(self.L.molsde_act_bwd, _p(pre.data), _p(out.grad), g.numel(), a, _p(g), self.s, what="act_bwd")
                    self.accum(xdot, g)
                if pre.needs:
                    g = self.empty(y.shape)
                    self._call(self.L.molsde_act_bwd2, _p(pre.data), _p(xdot.data), _p(out.grad), g.numel(), a, 0, _p(g), self.s,
                               what="act_bwd2")
                    self.accum(pre, g)
            self.ops.append(bwd)
        return out

    def add(self, a: Var, b: Var, alpha: float = 1.0) -> Var:
        y = self.empty(a.data.shape)
        self.ew(0, a.data, b.data, None, alpha, y)
        out = Var(y, a.needs or b.needs)
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                if a.needs:
                    g = self.empty(y.shape)
                    self.ew(0, out.grad, None, None, 1.0, g)
                    self.accum(a, g)
                if b.needs:
                    g = self.empty(y.shape)
                    self.ew(0, out.grad, None, None, alpha, g)
                    self.accum(b, g)
            self.ops.append(bwd)
        return out

    def mul(self, a: Var, b: Var, c: Optional[Var] = None) -> Var:
        """a*b (+c)"""
        y = self.empty(a.data.shape)
        self.ew(1, a.data, b.data, c.data if c is not None else None, 1.0, y)
        out = Var(y, a.needs or b.needs or (c is not None and c.needs))
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                if a.needs:
                    g = self.empty(y.shape)
                    self.ew(1, out.grad, b.data, None, 1.0, g)
                    self.accum(a, g)
                if b.needs:
                    g = self.empty(y.shape)
                    self.ew(1, out.grad, a.data, None, 1.0, g)
                    self.accum(b, g)
                if c is not None and c.needs:
                    g = self.empty(y.shape)
                    self.ew(0, out.grad, None, None, 1.0, g)
                    self.accum(c, g)
            self.ops.append(bwd)
        return out

    def scale_mask(self, x: Var, mask: torch.Tensor, alpha: float) -> Var:
        """x * mask * alpha (dropout with a given keep mask)."""
        y = self.empty(x.data.shape)
        self.ew(2, x.data, mask, None, alpha, y, cols=1)   # x * alpha * mask in one launch (mask is 0/1: same bits as (x*mask)*alpha)
        out = Var(y, x.needs)
        if x.needs:
            def bwd():
                if out.grad is None:
                    return
                g = self.empty(y.shape)
                self.ew(2, out.grad, mask, None, alpha, g, cols=1)
                self.accum(x, g)
            self.ops.append(bwd)
        return out

    # ------------------------------------------------------------------ gather / scatter
    def gather_pair(self, A: Var, ia: Optional[Index], B: Optional[Var] = None, ib: Optional[Index] = None,
                    rows: Optional[int] = None) -> Var:
        """out[r] = A[ia[r]] (+ B[ib[r]])"""
        cols = A.data.shape[1]
        rows = rows if rows is not None else (ia.idx.numel() if ia is not None else A.data.shape[0])
        y = self.empty(rows, cols)
        self._call(self.L.molsde_gather_pair, _p(A.data), _p(ia.idx) if ia else None, _p(B.data) if B is not None else None,
                   _p(ib.idx) if ib else None, rows, cols, _p(y), self.s, what="gather_pair")
        out = Var(y, A.needs or (B is not None and B.needs))
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                for V, ix in ((A, ia), (B, ib)):
                    if V is None or not V.needs:
                        continue
                    g = self.empty(V.data.shape)
                    self.seg_sum(out.grad, ix, cols, g)
                    self.accum(V, g)
            self.ops.append(bwd)
        return out

    def seg_sum(self, X: torch.Tensor, ix: Index, cols: int, out: torch.Tensor, scale: Optional[torch.Tensor] = None,
                accumulate: bool = False, row_div: int = 1):
        self._call(self.L.molsde_seg_gather_sum, _p(X), _p(ix.ptr), _p(ix.perm), ix.size, cols, _p(scale), int(accumulate),
                   row_div, _p(out), self.s, what="seg_gather_sum")

    def scatter_sum(self, X: Var, ix: Index) -> Var:
        """out[s] = sum_{r: idx[r] == s} X[r]   (scatter-add as a gather-reduce); backward is a gather."""
        cols = X.data.shape[1]
        y = self.empty(ix.size, cols)
        self.seg_sum(X.data, ix, cols, y)
        out = Var(y, X.needs)
        if X.needs:
            def bwd():
                if out.grad is None:
                    return
                g = self.empty(X.data.shape)
                self._call(self.L.molsde_gather_pair, _p(out.grad), _p(ix.idx), None, None, g.shape[0], cols, _p(g), self.s,
                           what="gather_pair")
                self.accum(X, g)
            self.ops.append(bwd)
        return out

    # ------------------------------------------------------------------ normalisation
    def layernorm(self, x: Var, g: Var, b: Var, eps: float = 1e-5) -> Var:
        M, D = x.data.shape
        y, mean, rstd = self.empty(M, D), self.empty(M), self.empty(M)
        self._call(self.L.molsde_layernorm_fwd, _p(x.data), M, D, _p(g.data), _p(b.data), eps, _p(y), _p(mean), _p(rstd), self.s,
                   what="layernorm_fwd")
        out = Var(y, True)

        def bwd():
            if out.grad is None:
                return
            dx, dyx = self.empty(M, D), self.empty(M, D)
            self._call(self.L.molsde_layernorm_bwd, _p(x.data), _p(out.grad), M, D, _p(g.data), _p(mean), _p(rstd), _p(dx), _p(dyx),
                       self.s, what="layernorm_bwd")
            with self.wgrad(dyx, out.grad, key=g.grad if g.needs else b.grad):
                if g.needs:
                    self.colsum(dyx, M, D, D, g.grad, accumulate=True)
                if b.needs:
                    self.colsum(out.grad, M, D, D, b.grad, accumulate=True)
            self.accum(x, dx)
        self.ops.append(bwd)
        return out

    def batchnorm(self, x: Var, g: Var, b: Var, running_mean: torch.Tensor, running_var: torch.Tensor, eps: float = 1e-5,
                  momentum: float = 0.1, relu: bool = False) -> Var:
        """nn.BatchNorm1d in train mode (+ fused ReLU)."""
        M, F = x.data.shape
        y, mean, rstd = self.empty(M, F), self.empty(F), self.empty(F)
        ws = self.empty(self.L.molsde_bn_ws_doubles(M, F), dtype=torch.float64)
        self._call(self.L.molsde_bn_train_fwd, _p(x.data), M, F, _p(g.data), _p(b.data), eps, momentum, _p(running_mean),
                   _p(running_var), int(relu), _p(y), _p(mean), _p(rstd), _p(ws), self.s, what="bn_train_fwd")
        _abi.touch_params()   # running statistics updated in place through raw pointers (eval-mode BN folds are cached)
        out = Var(y, True)

        def bwd():
            if out.grad is None:
                return
            # one call: ReLU mask applied inside the reductions and the dx pass, gamma / beta totals added to the parameter
            # gradients by the finish kernel (was act_bwd + bn_train_bwd + two accumulation launches)
            dx, dg, db = self.empty(M, F), self.empty(F), self.empty(F)
            ws2 = self.empty(self.L.molsde_bn_ws_doubles(M, F), dtype=torch.float64)
            gg = g.grad if (g.needs and g.grad is not None) else None
            gb = b.grad if (b.needs and b.grad is not None) else None
            self._call(self.L.molsde_bn_train_bwd_fused, _p(x.data), _p(out.grad), _p(y) if relu else None, M, F, _p(g.data), _p(mean),
                       _p(rstd), _p(dx), _p(dg), _p(db), _p(gg), _p(gb), _p(ws2), self.s, what="bn_train_bwd")
            if g.needs and gg is None:
                self.accum(g, dg)
            if b.needs and gb is None:
                self.accum(b, db)
            self.accum(x, dx)
        self.ops.append(bwd)
        return out

    def batchnorm_eval(self, x: Var, g: Var, b: Var, running_mean: torch.Tensor, running_var: torch.Tensor, eps: float = 1e-5,
                       relu: bool = False) -> Var:
        """nn.BatchNorm1d in eval mode (running statistics); forward only."""
        assert not x.needs, "eval-mode BatchNorm is on the inference path only"
        M, F = x.data.shape
        y, tmp = self.empty(M, F), self.empty(F)
        self._call(self.L.molsde_bn_eval, _p(x.data), M, F, _p(g.data), _p(b.data), _p(running_mean), _p(running_var), eps,
                   int(relu), _p(y), _p(tmp), self.s, what="bn_eval")
        return Var(y, False)

    # ------------------------------------------------------------------ message passing pieces
    def embed_sum(self, T: Var, keys: torch.Tensor, index: Optional[Index]) -> Var:
        """out[r] = sum_f T[keys[r,f]] (AtomEncoder / BondEncoder / nn.Embedding over one concatenated table).
        `index`: bucket Index of keys.reshape(-1) for the backward (None on the inference path)."""
        rows, F = keys.shape
        cols = T.data.shape[1]
        y = self.empty(rows, cols)
        self._call(self.L.molsde_embed_sum, _p(T.data), _p(keys), rows, F, cols, _p(y), self.s, what="embed_sum")
        out = Var(y, T.needs and index is not None)
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                with self.wgrad(out.grad, key=T.grad):
                    self.seg_sum(out.grad, index, cols, T.grad, accumulate=True, row_div=F)
            self.ops.append(bwd)
        return out

    def gin_aggregate(self, x: Var, T: Var, ekeys: torch.Tensor, ekey_index: Optional[Index], rowptr: torch.Tensor, src: Index,
                      tgt: Index, eps: Var) -> Var:
        """GINConv pre-MLP: (1+eps) x + sum_{e->i} relu(x_src + BondEncoder(e))  (molecule_gnn_model.py:23-31)."""
        N, cols = x.data.shape
        E, F = ekeys.shape
        pre = self.empty(N, cols)
        self._call(self.L.molsde_gin_aggregate_fwd, _p(x.data), _p(T.data), _p(ekeys), F, _p(rowptr), _p(src.idx), _p(eps.data), N, cols,
                   _p(pre), self.s, what="gin_aggregate_fwd")
        out = Var(pre, x.needs or T.needs or eps.needs)
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                dmsg = self.empty(E, cols)
                self._call(self.L.molsde_gin_message_bwd, _p(x.data), _p(T.data), _p(ekeys), F, _p(src.idx), _p(tgt.idx), _p(out.grad),
                           E, cols, _p(dmsg), self.s, what="gin_message_bwd")
                with self.wgrad(dmsg, out.grad, x.data, key=T.grad if T.needs else eps.grad):
                    if T.needs:
                        self.seg_sum(dmsg, ekey_index, cols, T.grad, accumulate=True, row_div=F)
                    if eps.needs:
                        ws = self.empty(128, dtype=torch.float64)
                        self._call(self.L.molsde_dot, _p(out.grad), _p(x.data), out.grad.numel(), 1.0, 1, _p(eps.grad), _p(ws), self.s,
                                   what="dot")
                if x.needs:
                    dx = self.empty(N, cols)
                    self.seg_sum(dmsg, src, cols, dx)
                    self.ew(4, out.grad, eps.data, dx, 1.0, dx)
                    self.accum(x, dx)
            self.ops.append(bwd)
        return out

    def edge_mul_reduce(self, x: Var, W: Var, rowptr: torch.Tensor, src: Index, tgt: Index, wcol0: Optional[int] = None,
                        escale: Optional[torch.Tensor] = None) -> Var:
        """CFConv message + aggregation: out[i] = sum_{e->i} x[src_e] * W[e]  (schnet.py:186-195).
        `wcol0`: the filter is the column block [wcol0, wcol0+cols) of the wider stack W (all interactions' filters side by
        side); its gradient is WRITTEN into the same block of W.grad (each block has exactly one consumer).
        `escale` [E] (no gradient): the filter is W[e] * escale[e] -- the cosine cutoff applied inside the kernels, so that neither
        the scaled stack nor the gradient of the scaled stack is ever materialised (same roundings as the two-step form)."""
        N, cols = x.data.shape
        E = W.data.shape[0]
        y = self.empty(N, cols)
        Wd = W.data if wcol0 is None else W.data[:, wcol0:wcol0 + cols]
        ldw = _ld(Wd)
        self._call(self.L.molsde_edge_mul_reduce_ld, _p(x.data), _p(src.idx), _p(Wd), ldw, _p(escale), _p(rowptr), None, N, cols, _p(y),
                   self.s, what="edge_mul_reduce")
        out = Var(y, x.needs or W.needs)
        if out.needs:
            def bwd():
                if out.grad is None:
                    if wcol0 is not None and W.needs:   # an unused output: its column block of the stack's gradient is zero, not garbage
                        if W.grad is None:
                            W.grad = self.empty(W.data.shape)
                        W.grad[:, wcol0:wcol0 + cols].zero_()
                    return
                if W.needs:
                    if wcol0 is None:
                        dW = self.empty(E, cols)
                        self._call(self.L.molsde_edge_mul_gather_ld, _p(out.grad), _p(tgt.idx), _p(x.data), _p(src.idx), _p(escale), E, cols,
                                   _p(dW), cols, self.s, what="edge_mul_gather")
                        self.accum(W, dW)
                    else:
                        if W.grad is None:
                            W.grad = self.empty(W.data.shape)   # every column block is written by its interaction's backward
                        dW = W.grad[:, wcol0:wcol0 + cols]
                        self._call(self.L.molsde_edge_mul_gather_ld, _p(out.grad), _p(tgt.idx), _p(x.data), _p(src.idx), _p(escale), E, cols,
                                   _p(dW), _ld(dW), self.s, what="edge_mul_gather")
                if x.needs:
                    dx = self.empty(N, cols)
                    self._call(self.L.molsde_edge_mul_reduce_ld, _p(out.grad), _p(tgt.idx), _p(Wd), ldw, _p(escale), _p(src.ptr),
                               _p(src.perm), N, cols, _p(dx), self.s, what="edge_mul_reduce")
                    self.accum(x, dx)
            self.ops.append(bwd)
        return out

    def rowscale(self, x: Var, r: torch.Tensor) -> Var:
        """y[row,:] = x[row,:] * r[row]"""
        cols = x.data.shape[1]
        y = self.empty(x.data.shape)
        self.ew(2, x.data, r, None, 1.0, y, cols=cols)
        out = Var(y, x.needs)
        if x.needs:
            def bwd():
                if out.grad is None:
                    return
                g = self.empty(y.shape)
                self.ew(2, out.grad, r, None, 1.0, g, cols=cols)
                self.accum(x, g)
            self.ops.append(bwd)
        return out

    def rowscale_var(self, x: Var, r: Var) -> Var:
        """y[row,:] = x[row,:] * r[row] where the per-row factor is differentiable too: d r[row] = <dy[row,:], x[row,:]>
        (the cosine cutoff C(d) of CFConv when forces are wanted, schnet.py:187-188)."""
        rows, cols = x.data.shape
        y = self.empty(rows, cols)
        self.ew(2, x.data, r.data, None, 1.0, y, cols=cols)
        out = Var(y, x.needs or r.needs)
        if out.needs:
            def bwd():
                if out.grad is None:
                    return
                if r.needs:
                    if r.grad is None:
                        r.grad = self.empty(rows)
                        acc = 0
                    else:
                        acc = 1
                    self._call(self.L.molsde_rowdot, _p(out.grad), _p(x.data), rows, cols, acc, _p(r.grad), self.s, what="rowdot")
                if x.needs:
                    g = self.empty(rows, cols)
                    self.ew(2, out.grad, r.data, None, 1.0, g, cols=cols)
                    self.accum(x, g)
            self.ops.append(bwd)
        return out

    def segment_readout(self, x: Var, seg_ptr: torch.Tensor, row2seg: torch.Tensor, mean: bool) -> Var:
        """out[g,:] = sum (or mean) of the rows of segment g (`scatter(h, batch, reduce=readout)`, schnet.py:115); `row2seg` int32 [rows]."""
        rows, cols = x.data.shape
        segs = seg_ptr.numel() - 1
        y = self.empty(segs, cols)
        self._call(self.L.molsde_segment_reduce, _p(x.data), _p(seg_ptr), segs, cols, int(mean), _p(y), self.s, what="segment_reduce")
        out = Var(y, x.needs)
        if x.needs:
            def bwd():
                if out.grad is None:
                    return
                g = self.empty(rows, cols)
                self._call(self.L.molsde_gather_pair, _p(out.grad), _p(row2seg), None, None, rows, cols, _p(g), self.s, what="gather_pair")
                if mean:
                    cnt = (seg_ptr[1:] - seg_ptr[:-1]).clamp_min(1).float()
                    inv = (1.0 / cnt)[row2seg.long()].contiguous()   # [rows] bookkeeping
                    self.ew(2, g, inv, None, 1.0, g, cols=cols)
                self.accum(x, g)
            self.ops.append(bwd)
        return out

    # ------------------------------------------------------------------ batched small GEMMs (per-channel layers)
    def gemm_batched(self, batch, M, N, K, A, sam, sak, bsA, B, sbn, sbk, bsB, C, ldc, bsC, bias=None, bsBias=0, accumulate=False):
        n = self.L.molsde_tc_gemm_batched_ws_floats(batch, M, N, K)
        ws = self.empty(n) if n > 0 else None
        self._call(self.L.molsde_tc_gemm_batched, batch, M, N, K, _p(A), sam, sak, bsA, _p(B), sbn, sbk, bsB, _p(bias), bsBias, 0,
                   _p(C), ldc, bsC, int(accumulate), _p(ws), n, None, self.s, what="tc_gemm_batched")

    def grouped_linear(self, x: Var, Wp: Var, bp: Var, G: int) -> Var:
        """G independent nn.Linear layers in one launch: y[:, g*No:(g+1)*No] = x[:, g*Ki:(g+1)*Ki] W_g^T + b_g with the weights
        stacked as Wp [G, No, Ki] (adjacent parameters of a ParamStore) and bp [G*No]."""
        rows = x.data.shape[0]
        _, No, Ki = Wp.data.shape
        ldx = _ld(x.data)
        y = self.empty(rows, G * No)
        self.gemm_batched(G, rows, No, Ki, x.data, ldx, 1, Ki, Wp.data, Ki, 1, No * Ki, y, G * No, No, bias=bp.data, bsBias=No)
        out = Var(y, True)

        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            with self.wgrad(dy, x.data, key=Wp.grad if Wp.needs else bp.grad):
                if Wp.needs:   # dW_g = dy_g^T x_g
                    self.gemm_batched(G, No, Ki, rows, dy, 1, G * No, No, x.data, 1, ldx, Ki, Wp.grad, Ki, No * Ki, accumulate=True)
                if bp.needs:
                    self.colsum(dy, rows, G * No, G * No, bp.grad, accumulate=True)
            if x.needs:    # dx_g = dy_g W_g
                dx = self.empty(rows, G * Ki)
                self.gemm_batched(G, rows, Ki, No, dy, G * No, 1, No, Wp.data, 1, Ki, No * Ki, dx, G * Ki, Ki)
                self.accum(x, dx)
        self.ops.append(bwd)
        return out

    def grouped_matmul_shared(self, x: Var, Wp: Var, G: int) -> Var:
        """y[:, g*Fo:(g+1)*Fo] = x @ W_g for G weights stacked as Wp [G, Fin, Fo] (NodeNetwork_dense.weight layout), shared x."""
        rows, Fin = x.data.shape
        Fo = Wp.data.shape[2]
        ldx = _ld(x.data)
        y = self.empty(rows, G * Fo)
        self.gemm_batched(G, rows, Fo, Fin, x.data, ldx, 1, 0, Wp.data, 1, Fo, Fin * Fo, y, G * Fo, Fo)
        out = Var(y, True)

        def bwd():
            if out.grad is None:
                return
            dy = out.grad
            if Wp.needs:   # dW_g [Fin,Fo] = x^T dy_g
                with self.wgrad(dy, x.data, key=Wp.grad):
                    self.gemm_batched(G, Fin, Fo, rows, x.data, 1, ldx, 0, dy, 1, G * Fo, Fo, Wp.grad, Fo, Fin * Fo, accumulate=True)
            if x.needs:    # dx = sum_g dy_g W_g^T : per-group products into a scratch, then a fixed-order sum over the groups
                tmp = self.empty(G, rows, Fin)
                self.gemm_batched(G, rows, Fin, Fo, dy, G * Fo, 1, Fo, Wp.data, Fo, 1, Fin * Fo, tmp, Fin, rows * Fin)
                dx = self.empty(rows, Fin)
                self._call(self.L.molsde_sum_slices, _p(tmp), G, rows * Fin, _p(dx), self.s, what="sum_slices")
                self.accum(x, dx)
        self.ops.append(bwd)
        return out
