"""Differentiable / generic solve path (CUDA tensors, autograd-visible).

Used when (a) gradients must flow through the solver -- every training loop of the reference
backpropagates through ``odeint`` with plain autograd (heat_dynamics.py:333, dgnn.py:204;
SURVEY.md section 3.5) -- or (b) ``func`` is an arbitrary callable the fused kernels do not know.
The state stays on the GPU; the RHS of a recognised ``ODEFunc`` still runs the hand-written
SpMM kernel through ``SpmmFn`` (forward Phi x, backward Phi^T g).  The solver algebra here is
issued as PyTorch ops so that autograd records it, with the same operation order as
torchdiffeq/_impl (so training trajectories track the reference's); the fully fused
backward is the "next" row N1 of SURVEY.md section 8(f) and is NOT claimed by this module.

There is no CPU execution here: inputs must be CUDA tensors.
"""
from __future__ import annotations

from typing import Callable, List

import torch

from . import solver as _solver
from .graph import CsrGraph

# Dormand-Prince(5,4)/Shampine tableau (torchdiffeq/_impl/dopri5.py:11-36)
_A = (1 / 5, 3 / 10, 4 / 5, 8 / 9, 1.0, 1.0)
_B = (
    (1 / 5,),
    (3 / 40, 9 / 40),
    (44 / 45, -56 / 15, 32 / 9),
    (19372 / 6561, -25360 / 2187, 64448 / 6561, -212 / 729),
    (9017 / 3168, -355 / 33, 46732 / 5247, 49 / 176, -5103 / 18656),
    (35 / 384, 0, 500 / 1113, 125 / 192, -2187 / 6784, 11 / 84),
)
_E = (35 / 384 - 1951 / 21600, 0, 500 / 1113 - 22642 / 50085, 125 / 192 - 451 / 720,
      -2187 / 6784 - -12231 / 42400, 11 / 84 - 649 / 6300, -1.0 / 60.0)
_M = (6025192743 / 30085553152 / 2, 0, 51252292925 / 65400821598 / 2, -2691868925 / 45128329728 / 2,
      187940372067 / 1594534317056 / 2, -1776094331 / 19743644256 / 2, 11237099 / 235043384 / 2)


class SpmmFn(torch.autograd.Function):
    """Phi @ x on the CUDA SpMM kernel; d/dx = Phi^T @ g on the same kernel (transposed CSR)."""

    @staticmethod
    def forward(ctx, x: torch.Tensor, graph: CsrGraph) -> torch.Tensor:
        ctx.graph = graph
        return _solver.spmm(graph, x.contiguous())

    @staticmethod
    def backward(ctx, g: torch.Tensor):
        return _solver.spmm(ctx.graph.transpose(), g.contiguous()), None


class RhsFn(torch.autograd.Function):
    """One ``ODEFunc`` evaluation k = relu((Phi x) W^T + b) as a single autograd node: forward = the fused RHS kernel(s),
    backward = ``ndcn_rhs_vjp_f32`` (gather, GEMM with ReLU-mask epilogue, GEMM with W, gather with Phi^T) plus
    ``ndcn_weight_grads_f32`` for dW / db.  What plain autograd would record through neural_dynamics.py:27-36 as
    sparse.mm + addmm + relu (3 forward, ~6 backward launches, z and the pre-activation kept alive); here only x
    is saved and z / the mask are recomputed."""

    @staticmethod
    def forward(ctx, x, W, b, graph, graph_t, flags):
        from . import _ffi

        no_control = bool(flags & _ffi.F_NO_CONTROL)
        spec = _solver.RhsSpec(_ffi.RHS_NDCN, int(x.shape[1]), flags, None if no_control else W, None if no_control else b)
        ctx.save_for_backward(x, W, b)
        ctx.graph, ctx.graph_t, ctx.flags = graph, graph_t, flags
        return _solver.rhs_eval(graph, spec, x.contiguous(), cache_weights=True)

    @staticmethod
    def backward(ctx, gk):
        from . import _ffi

        x, W, b = ctx.saved_tensors
        flags = ctx.flags
        no_control = bool(flags & _ffi.F_NO_CONTROL)
        spec = _solver.RhsSpec(_ffi.RHS_NDCN, int(x.shape[1]), flags, None if no_control else W, None if no_control else b)
        x = x.contiguous()
        gx = torch.empty_like(x)
        gp, z = _vjp(ctx.graph, ctx.graph_t, spec, x, gk.contiguous(), 1.0, gx, False)
        dW = db = None
        if not no_control:
            need_w, need_b = ctx.needs_input_grad[1], ctx.needs_input_grad[2]
            if need_w or need_b:
                dW = torch.empty_like(W)
                db = torch.empty_like(b)
                _solver.weight_grads(gp, z, dW, db, accumulate=False)
                if not need_w:
                    dW = None
                if not need_b:
                    db = None
        return (gx if ctx.needs_input_grad[0] else None), dW, db, None, None, None


def _require_cuda_state(y0: torch.Tensor) -> None:
    if not y0.is_cuda:
        raise RuntimeError(
            "ndcn_b200: the differentiable / generic solve path needs CUDA tensors (run the script with "
            "--gpu 0 / without --no-cuda); there is no CPU fallback in this backend.")


def _wsum(dt, coeffs, ks):
    acc = 0
    for c, k in zip(coeffs, ks):
        acc = acc + (dt * c) * k  # misc.py:22-25
    return acc


def _rms(x):
    return x.norm() / (x.numel() ** 0.5)


def _first_step(func, t0, y0, order, rtol, atol, f0):  # misc.py:84-143
    t0 = t0.to(y0)
    scale = atol + torch.abs(y0) * rtol
    d0, d1 = _rms(y0 / scale), _rms(f0 / scale)
    if d0.item() < 1e-5 or d1.item() < 1e-5:
        h0 = torch.tensor(1e-6).to(t0)
    else:
        h0 = 0.01 * (d0 / d1)
    f1 = func(t0 + h0, y0 + h0 * f0)
    d2 = _rms((f1 - f0) / scale) / h0
    if d1.item() <= 1e-15 and d2.item() <= 1e-15:
        h1 = torch.max(torch.tensor(1e-6).to(h0), h0 * 1e-3)
    else:
        h1 = (0.01 / max(d1, d2)) ** (1.0 / float(order + 1))
    return torch.min(100 * h0, h1)


def _as64(v, device):
    return torch.tensor(v).type(torch.float64).to(device)  # misc.py:39-47: via an fp32 tensor


def _resize(dt, msr, dev):  # misc.py:160-170 with dopri5.py:71-73's constants
    safety, ifactor, dfactor = _as64(0.9, dev), _as64(10.0, dev), _as64(0.2, dev)
    if msr == 0:
        return dt * ifactor
    if msr < 1:
        dfactor = _as64(1, dev)
    root = torch.sqrt(msr).to(dt)
    expo = torch.tensor(1 / 5).to(dt)
    return dt / torch.max(1 / ifactor, torch.min(root ** expo / safety, 1 / dfactor))


def _dense_coeffs(y0, y1, ks, dt):  # dopri5.py:39-45, interp.py:5-35
    ymid = y0 + _wsum(dt, _M, ks)
    f0, f1 = ks[0], ks[-1]
    vs = (f0, f1, y0, y1, ymid)

    def dot(ws):
        acc = 0
        for w, v in zip(ws, vs):
            acc = acc + w * v
        return acc

    return [dot((-2 * dt, 2 * dt, -8, -8, 16)), dot((5 * dt, -3 * dt, 18, 14, -32)),
            dot((-4 * dt, dt, -11, -5, 16)), dt * f0, y0]


def _dense_eval(cs, t0, t1, t):  # interp.py:38-65
    dtype, dev = cs[0].dtype, cs[0].device
    t0, t1, t = t0.to(dev, dtype), t1.to(dev, dtype), t.to(dev, dtype)
    assert (t0 <= t) & (t <= t1), "invalid interpolation, fails `t0 <= t <= t1`: {}, {}, {}".format(t0, t, t1)
    x = ((t - t0) / (t1 - t0)).type(dtype)
    xs = [torch.tensor(1).type(dtype).to(dev), x]
    for _ in range(2, len(cs)):
        xs.append(xs[-1] * x)
    acc = 0
    for c, xp in zip(cs, reversed(xs)):
        acc = acc + c * xp
    return acc


def solve(func: Callable, y0: torch.Tensor, t: torch.Tensor, rtol: float, atol: float, method: str,
          max_num_steps: int = 2 ** 31 - 1) -> torch.Tensor:
    """Single-tensor odeint on CUDA tensors with autograd (euler | midpoint | rk4 | dopri5)."""
    _require_cuda_state(y0)
    assert (t[1:] > t[:-1]).all(), "t must be strictly increasing or decrasing"
    outs: List[torch.Tensor] = [y0]
    if method in ("euler", "midpoint", "rk4"):  # solvers.py:79-99, fixed_grid.py, rk_common.py:72-78
        tt = t.type_as(y0).to(y0.device)
        y = y0
        for t0, t1 in zip(tt[:-1], tt[1:]):
            dt = t1 - t0
            if method == "euler":
                dy = dt * func(t0, y)
            elif method == "midpoint":
                dy = dt * func(t0 + dt / 2, y + func(t0, y) * dt / 2)
            else:
                k1 = func(t0, y)
                k2 = func(t0 + dt / 3, y + dt * k1 / 3)
                k3 = func(t0 + dt * 2 / 3, y + dt * (k1 / -3 + k2))
                k4 = func(t0 + dt, y + dt * (k1 - k2 + k3))
                dy = (k1 + 3 * k2 + 3 * k3 + k4) * (dt / 8)
            y = y + dy
            outs.append(y)
        return torch.stack(outs)
    if method != "dopri5":
        raise ValueError("ndcn_b200 covers euler | midpoint | rk4 | dopri5, got %r" % (method,))
    # solvers.py:25-33, dopri5.py:58-122
    dev = y0.device
    t = t.to(dev, torch.float64)
    f = func(t[0].type_as(y0), y0)
    dt = _first_step(func, t[0], y0, 4, rtol, atol, f).to(t)
    y, lo, hi = y0, t[0], t[0]
    cs = [y0] * 5
    for i in range(1, len(t)):
        n = 0
        while t[i] > hi:
            assert n < max_num_steps, "max_num_steps exceeded ({}>={})".format(n, max_num_steps)
            assert hi + dt > hi, "underflow in dt {}".format(dt.item())
            assert bool(torch.isfinite(y).all()), "non-finite values in state `y`: {}".format(y)
            h = dt.type(y.dtype)
            tb = hi.type(y.dtype)
            ks = [f]
            ys = y
            for a, b in zip(_A, _B):
                ys = y + _wsum(h, b, ks)
                ks.append(func(tb + a * h, ys))
            err = _wsum(h, _E, ks)
            ratio = err / (atol + rtol * torch.max(torch.abs(y), torch.abs(ys)))
            msr = torch.mean(ratio * ratio)
            lo = hi
            if bool(msr <= 1):
                cs = _dense_coeffs(y, ys, ks, h)
                y, f, hi = ys, ks[-1], hi + dt
            dt = _resize(dt, msr, dev)
            n += 1
        outs.append(_dense_eval(cs, lo, hi, t[i]))
    return torch.stack(outs)


# ------------------------------------------------------------------------------------------------
# Fused training path for the fixed-grid solvers (first cut).
#
# The dynamics scripts train NDCN with `--method euler` by default (heat_dynamics.py:22,317-334):
# forward = the fused solver (one launch per RHS evaluation), backward = the discrete adjoint of the
# same scheme with every RHS vjp on the library's kernels (`ndcn_rhs_vjp_f32`: gather, GEMM with a
# ReLU-mask epilogue, GEMM with W untransposed, gather with Phi^T fused with the adjoint update).
# For a fixed grid the discrete adjoint IS what autograd computes through torchdiffeq's loop
# (solvers.py:79-99), up to fp32 summation order.  dW = gp^T z and db = sum gp are plain library
# GEMM / reduction calls.  dopri5 keeps the op-by-op autograd path above: its step-size controller is
# part of the reference's autograd graph (misc.py:160-170), which a hand-written adjoint would drop.
# ------------------------------------------------------------------------------------------------
def _vjp(graph: CsrGraph, graph_t: CsrGraph, spec, x, gk, scale: float, gx, accumulate: bool, cache_weights: bool = True):
    """(gp, z): gx (+)= d<gk*scale, f(x)>/dx;  dW = gp^T z, db = gp.sum(0) are left to the caller."""
    import ctypes as C

    from . import _ffi

    gp = torch.empty_like(x)
    z = torch.empty_like(x) if not (spec.flags & _ffi.F_NO_GRAPH) else x
    keep: list = []
    desc = spec.to_c(keep, prepare=cache_weights)
    with torch.cuda.device(x.device):
        rc = _ffi.lib().ndcn_rhs_vjp_f32(graph.handle, graph_t.handle, C.byref(desc), x.data_ptr(), gk.data_ptr(),
                                         float(scale), gx.data_ptr(), 1 if accumulate else 0, gp.data_ptr(),
                                         z.data_ptr() if z is not x else None,
                                         _solver.current_stream_ptr(x.device))
    _ffi.check(rc, "ndcn_rhs_vjp_f32")
    return gp, z


class FusedFixedGridFn(torch.autograd.Function):
    """``odeint`` for euler | midpoint | rk4 on a recognised ODEFunc, differentiable in (y0, W, b)."""

    persistent_backward = True  # H <= 32, small graphs: one cooperative launch for the whole backward pass

    @staticmethod
    def forward(ctx, y0, W, b, t32, graph, graph_t, flags, method):
        from . import _ffi

        H = y0.shape[1]
        no_control = bool(flags & _ffi.F_NO_CONTROL)
        spec = _solver.RhsSpec(_ffi.RHS_NDCN, H, flags, None if no_control else W, None if no_control else b)
        with torch.no_grad():
            slab = _solver.odeint_fused(graph, spec, y0.contiguous(), t32, method=method)
        ctx.save_for_backward(slab, W, b, t32)
        ctx.graph, ctx.graph_t, ctx.flags, ctx.method = graph, graph_t, flags, method
        return slab

    @staticmethod
    def backward(ctx, g_slab):
        from . import _ffi

        slab, W, b, t32 = ctx.saved_tensors
        graph, graph_t, flags, method = ctx.graph, ctx.graph_t, ctx.flags, ctx.method
        H = slab.shape[2]
        no_control = bool(flags & _ffi.F_NO_CONTROL)
        spec = _solver.RhsSpec(_ffi.RHS_NDCN, H, flags, None if no_control else W, None if no_control else b)
        g_slab = g_slab.contiguous()
        dW = torch.zeros_like(W) if not no_control else None
        db = torch.zeros_like(b) if not no_control else None
        tt = t32.detach().to("cpu", torch.float32)
        if H <= 32 and graph.n_rows <= 16384 and FusedFixedGridFn.persistent_backward:
            # the whole backward pass as one cooperative launch (csrc/small_solver.cuh::k_adjoint_small)
            import ctypes as C

            lam = torch.empty_like(g_slab[0])
            keep: list = []
            desc = spec.to_c(keep)
            t64 = tt.to(torch.float64).contiguous()
            with torch.cuda.device(slab.device):
                rc = _ffi.lib().ndcn_fixed_grid_adjoint_small_f32(
                    graph.handle, graph_t.handle, C.byref(desc), _ffi.METHODS[method],
                    C.cast(t64.data_ptr(), _ffi.c_double_p), int(t64.numel()), slab.data_ptr(), g_slab.data_ptr(),
                    lam.data_ptr(), dW.data_ptr() if dW is not None else None, db.data_ptr() if db is not None else None,
                    _solver.current_stream_ptr(slab.device))
            if rc == 0:
                return lam, dW, db, None, None, None, None, None
            if rc != _ffi.E_ARG:
                _ffi.check(rc, "ndcn_fixed_grid_adjoint_small_f32")
        lam = g_slab[-1].clone()

        def f(x):
            return _solver.rhs_eval(graph, spec, x)

        def vjp(x, gk, scale, gx, accumulate=True):
            gp, z = _vjp(graph, graph_t, spec, x, gk, scale, gx, accumulate)
            if not no_control:
                _solver.weight_grads(gp, z, dW, db, accumulate=True)  # dW += gp^T z, db += sum_r gp

        for i in range(slab.shape[0] - 2, -1, -1):
            y = slab[i]
            dt = float(tt[i + 1] - tt[i])  # fp32 difference of the fp32 grid (solvers.py:81,89)
            if method == "euler":  # y' = y + dt f(y)
                vjp(y, lam, dt, lam)
            elif method == "midpoint":  # y' = y + dt f(y + dt/2 f(y))
                ym = y + f(y) * (dt / 2)
                gm = torch.zeros_like(lam)
                vjp(ym, lam, dt, gm, accumulate=False)      # gm = dL/dym
                lam.add_(gm)
                vjp(y, gm, dt / 2, lam)
            else:  # rk4 = 3/8 rule, rk_common.py:72-78
                k1 = f(y)
                y2 = y + k1 * (dt / 3)
                k2 = f(y2)
                y3 = y + (k2 - k1 / 3) * dt
                k3 = f(y3)
                y4 = y + (k1 - k2 + k3) * dt
                g4 = torch.empty_like(lam)
                vjp(y4, lam, dt / 8, g4, accumulate=False)          # g4 = dL/dy4 via k4
                gk3 = lam * (3 * dt / 8) + g4 * dt
                g3 = torch.empty_like(lam)
                vjp(y3, gk3, 1.0, g3, accumulate=False)
                gk2 = lam * (3 * dt / 8) - g4 * dt + g3 * dt
                g2 = torch.empty_like(lam)
                vjp(y2, gk2, 1.0, g2, accumulate=False)
                gk1 = lam * (dt / 8) + g4 * dt - g3 * (dt / 3) + g2 * (dt / 3)
                lam.add_(g4).add_(g3).add_(g2)
                vjp(y, gk1, 1.0, lam)
            lam.add_(g_slab[i])
        return lam, dW, db, None, None, None, None, None


def solve_fixed_grid_fused(odefunc, y0: torch.Tensor, t: torch.Tensor, method: str, graph: CsrGraph, flags: int):
    """Entry used by ``odeint`` when gradients are required, the RHS is a recognised ODEFunc without
    active dropout and the method is euler | midpoint | rk4."""
    _require_cuda_state(y0)
    t32 = t.detach().to(torch.float32)
    no_graph = bool(flags & 1)
    graph_t = graph if no_graph else graph.transpose()
    W, b = odefunc.wt.weight, odefunc.wt.bias
    return FusedFixedGridFn.apply(y0, W, b, t32, graph, graph_t, flags, method)
