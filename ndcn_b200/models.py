"""Host-side mirror of the reference's operator surface for the hot path:
``ODEFunc`` / ``ODEBlock`` / ``ODEBlock2`` / ``NDCN`` (neural_dynamics.py:8-160).

Same constructor signatures, attribute and sub-module names (so ``state_dict`` keys match:
``neural_dynamic_layer.odefunc.wt.{weight,bias}``, ``input_layer.{0,2}.*``, ``output_layer.*``),
same forward contracts.  The arithmetic of ``ODEFunc.forward`` and of the ODE solve lives in
libndcn_b200.so; these classes only hold parameters and route tensors.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from .odeint import odeint as _odeint, odeint_adjoint as _odeint_adjoint
from . import solver as _solver
from .autograd_solver import RhsFn, SpmmFn
from .graph import cached_graph, require_cuda
from .solver import RhsSpec


class ODEFunc(nn.Module):
    """dX/dt = relu(dropout(Linear(Phi X)))   -- neural_dynamics.py:8-39.

    ``A`` stays a plain attribute, not a buffer (``model.to(device)`` does not move it in the
    reference either, neural_dynamics.py:14); it may be dense or sparse COO, on any device:
    it is converted to CSR on the GPU once and cached.
    """

    def __init__(self, hidden_size, A, dropout=0.0, no_graph=False, no_control=False):
        super(ODEFunc, self).__init__()
        self.hidden_size = hidden_size
        self.dropout = dropout
        self.dropout_layer = nn.Dropout(dropout)
        self.A = A  # N_node * N_node
        self.wt = nn.Linear(hidden_size, hidden_size)
        self.no_graph = no_graph
        self.no_control = no_control

    def forward(self, t, x):
        """t is ignored (autonomous system); x is [N, hidden]."""
        dev = require_cuda(x.device if x.is_cuda else None)
        if not x.is_cuda:
            raise RuntimeError("ndcn_b200.ODEFunc.forward needs a CUDA state (got %s); run with --gpu 0 or call "
                               "odeint(), which stages CPU inputs of gradient-free solves itself" % x.device)
        active_dropout = self.training and self.dropout > 0
        params = [p for p in self.wt.parameters()] if not self.no_control else []
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in params))
        fusable = x.dim() == 2 and x.dtype == torch.float32 and not active_dropout
        if fusable and not needs_grad:
            # one fused kernel: gather + W GEMM + bias + ReLU
            graph = None if self.no_graph else cached_graph(self, self.A, dev)
            if graph is None:
                graph = _identity_graph(x.shape[0], dev)
            spec = RhsSpec.ndcn(x.shape[1], None if self.no_control else self.wt.weight,
                                None if self.no_control else self.wt.bias,
                                no_graph=self.no_graph, no_control=self.no_control)
            return _solver.rhs_eval(graph, spec, x)
        if fusable and (self.no_control or self.wt.bias is not None):
            # one autograd node: fused forward, fused vjp + own dW/db reduction in backward
            from . import _ffi
            graph = _identity_graph(x.shape[0], dev) if self.no_graph else cached_graph(self, self.A, dev)
            graph_t = graph if self.no_graph else graph.transpose()
            flags = (_ffi.F_NO_GRAPH if self.no_graph else 0) | (_ffi.F_NO_CONTROL if self.no_control else 0)
            W = self.wt.weight if not self.no_control else x.new_zeros(())
            b = self.wt.bias if not self.no_control else x.new_zeros(())
            return RhsFn.apply(x, W, b, graph, graph_t, flags)
        # everything else (active dropout: RNG-dependent; non-fp32 or non-2-D states, which the reference's
        # dtype-agnostic ATen path accepts, neural_dynamics.py:27-36): differentiable composition of our SpMM kernel
        # (fp32 2-D) or torch's own sparse/dense product with autograd ops
        if not self.no_graph:
            if x.dim() == 2 and x.dtype == torch.float32:
                x = SpmmFn.apply(x, cached_graph(self, self.A, dev))
            else:
                A = self.A.to(device=x.device, dtype=x.dtype)
                x = torch.sparse.mm(A, x) if A.is_sparse else torch.matmul(A, x)
        if not self.no_control:
            x = self.wt(x)
        x = self.dropout_layer(x)
        return F.relu(x)


_IDENTITY = {}


def _identity_graph(n: int, dev: torch.device):
    """no_graph still needs a graph handle for the row count; an empty CSR does."""
    from .graph import CsrGraph

    key = (n, str(dev))
    if key not in _IDENTITY:
        _IDENTITY[key] = CsrGraph(torch.zeros(n + 1, dtype=torch.int32, device=dev),
                                  torch.zeros(0, dtype=torch.int32, device=dev),
                                  torch.zeros(0, dtype=torch.float32, device=dev), n, n)
    return _IDENTITY[key]


class ODEBlock(nn.Module):
    """neural_dynamics.py:42-79."""

    def __init__(self, odefunc, rtol=.01, atol=.001, method='dopri5', adjoint=False, terminal=False):
        super(ODEBlock, self).__init__()
        self.odefunc = odefunc
        self.rtol = rtol
        self.atol = atol
        self.method = method
        self.adjoint = adjoint
        self.terminal = terminal

    def forward(self, vt, x, decoder=None):
        """``decoder=(W, b)``: extension used by ``NDCN.forward`` -- the output Linear is applied to every
        returned state inside the solve (no ``[T, N, H]`` slab)."""
        integration_time_vector = vt.type_as(x)  # rounds the grid to the state's dtype FIRST (:71)
        solve = _odeint_adjoint if self.adjoint else _odeint  # :72-78
        return solve(self.odefunc, x, integration_time_vector, rtol=self.rtol, atol=self.atol,
                     method=self.method, terminal_only=bool(self.terminal), decoder=decoder)


class ODEBlock2(nn.Module):
    """neural_dynamics.py:82-119 (time vector fixed at construction)."""

    def __init__(self, odefunc, vt, rtol=.01, atol=.001, method='dopri5', adjoint=False, terminal=False):
        super(ODEBlock2, self).__init__()
        self.odefunc = odefunc
        self.integration_time_vector = vt
        self.rtol = rtol
        self.atol = atol
        self.method = method
        self.adjoint = adjoint
        self.terminal = terminal

    def forward(self, x):
        integration_time_vector = self.integration_time_vector.type_as(x)
        solve = _odeint_adjoint if self.adjoint else _odeint  # :111-118
        return solve(self.odefunc, x, integration_time_vector, rtol=self.rtol, atol=self.atol,
                     method=self.method, terminal_only=bool(self.terminal))


class NDCN(nn.Module):
    """Encoder -> ODEBlock -> decoder   -- neural_dynamics.py:122-160."""

    def __init__(self, input_size, hidden_size, A, num_classes, dropout=0.0,
                 no_embed=False, no_graph=False, no_control=False,
                 rtol=.01, atol=.001, method='dopri5'):
        super(NDCN, self).__init__()
        self.input_size = input_size
        self.hidden_size = hidden_size
        self.A = A
        self.num_classes = num_classes
        self.dropout = dropout
        self.dropout_layer = nn.Dropout(dropout)
        self.no_embed = no_embed
        self.no_graph = no_graph
        self.no_control = no_control
        self.rtol = rtol
        self.atol = atol
        self.method = method
        self.input_layer = nn.Sequential(nn.Linear(input_size, hidden_size, bias=True), nn.Tanh(),
                                         nn.Linear(hidden_size, hidden_size, bias=True))
        self.neural_dynamic_layer = ODEBlock(
            ODEFunc(hidden_size, A, dropout=dropout, no_graph=no_graph, no_control=no_control),
            rtol=rtol, atol=atol, method=method)
        self.output_layer = nn.Linear(hidden_size, num_classes, bias=True)

    def forward(self, vt, x):
        if not self.no_embed:
            x = self.input_layer(x)
        needs_grad = torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))
        if x.is_cuda and not needs_grad and self.num_classes <= 8:
            # inference: output_layer is fused into the solve's emission kernels (neural_dynamics.py:159 applies it
            # to the whole [T, N, H] slab; here only [T, N, num_classes] is ever written)
            return self.neural_dynamic_layer(vt, x, decoder=(self.output_layer.weight, self.output_layer.bias))
        hvx = self.neural_dynamic_layer(vt, x)
        return self.output_layer(hvx)
