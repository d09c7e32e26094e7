"""Ground-truth network dynamics as importable operators.

In the reference these three classes are defined INSIDE the driver scripts
(heat_dynamics.py:186-204, gene_dynamics.py:186-205, mutualistic_dynamics.py:186-232), so a
script run through the launcher keeps using its own class objects; ``ndcn_b200.odeint``
recognises those by name/attributes.  The classes below have the same constructors and
``forward(t, x)`` contract for users who import them directly; ``forward`` evaluates the RHS
with the CUDA kernels (no autograd: these are data generators, always run under ``no_grad``).
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import solver as _solver
from .graph import cached_graph, require_cuda
from .solver import RhsSpec


class _Dynamics(nn.Module):
    def _spec(self, width: int) -> RhsSpec:  # pragma: no cover - abstract
        raise NotImplementedError

    def _operator(self) -> torch.Tensor:  # pragma: no cover - abstract
        raise NotImplementedError

    def forward(self, t, x):
        dev = require_cuda(x.device if x.is_cuda else None)
        graph = cached_graph(self, self._operator(), dev)
        out = _solver.rhs_eval(graph, self._spec(int(x.shape[1])), x.detach().to(dev, torch.float32))
        return out if x.is_cuda else out.to(x.device)


class HeatDiffusion(_Dynamics):
    """dX/dt = -k L X  (the module stores -L, heat_dynamics.py:190)."""

    def __init__(self, L, k=1):
        super(HeatDiffusion, self).__init__()
        self.L = -L
        self.k = k

    def _operator(self):
        return self.L

    def _spec(self, width):
        return RhsSpec.heat(width, self.k)


class GeneDynamics(_Dynamics):
    """dx_i/dt = -b x_i^f + sum_j A_ij x_j^h / (x_j^h + 1)  (gene_dynamics.py:186-205)."""

    def __init__(self, A, b, f=1, h=2):
        super(GeneDynamics, self).__init__()
        self.A = A
        self.b = b
        self.f = f
        self.h = h

    def _operator(self):
        return self.A

    def _spec(self, width):
        return RhsSpec.gene(width, self.b, self.f, self.h)


class MutualDynamics(_Dynamics):
    """dx_i/dt = b + x_i(1 - x_i/k)(x_i/c - 1) + sum_j A_ij x_i x_j / (d + e.. + h..)
    (mutualistic_dynamics.py:186-232).  The reference's [N,1] branch attaches ``e`` to the
    neighbour and ``h`` to the node itself, its [N,d>1] loop the other way round; each width
    reproduces the branch the reference would take."""

    def __init__(self, A, b=0.1, k=5., c=1., d=5., e=0.9, h=0.1):
        super(MutualDynamics, self).__init__()
        self.A = A
        self.b = b
        self.k = k
        self.c = c
        self.d = d
        self.e = e
        self.h = h

    def _operator(self):
        return self.A

    def _spec(self, width):
        return RhsSpec.mutual(width, self.b, self.k, self.c, self.d, self.e, self.h)
