"""ndcn_b200 -- B200-native (sm_100a) backend for the NDCN ODE-integrated graph-convolution
hot path: ``ODEFunc`` (normalized-Laplacian SpMM -> Linear -> ReLU) integrated by
dopri5 / rk4 / midpoint / euler, plus the Heat / Gene / Mutualistic ground-truth dynamics.

Layout
  csrc/            hand-written CUDA kernels + the C ABI (libndcn_b200.so, include/ndcn_b200.h)
  _ffi.py          ctypes binding (no fallback: raises if the library is missing)
  graph.py         CSR operator on the GPU
  solver.py        tensor plumbing around ndcn_odeint_f32 / ndcn_rhs_eval_f32
  odeint.py        torchdiffeq.odeint-compatible dispatch
  models.py        ODEFunc / ODEBlock / ODEBlock2 / NDCN with the reference's signatures
  dynamics.py      HeatDiffusion / GeneDynamics / MutualDynamics
  partition.py     1-D row partition + halo exchange for multi-GPU solves
  shims/           modules named like the reference's (neural_dynamics, torchdiffeq)
  run.py           launcher: runs an unmodified reference script on this backend
  workloads.py     sparse graph generators, operators and node orderings (host side, start-up only)
  experiment.py    the dynamics scripts' experiment on sparse operators (100k - 4M nodes)
"""
from .graph import CsrGraph, cached_graph
from .solver import RhsSpec, SolveInfo, odeint_fused, rhs_eval, spmm
from .odeint import odeint, odeint_adjoint
from .models import NDCN, ODEBlock, ODEBlock2, ODEFunc
from .dynamics import GeneDynamics, HeatDiffusion, MutualDynamics

__all__ = ["CsrGraph", "cached_graph", "RhsSpec", "SolveInfo", "odeint_fused", "rhs_eval", "spmm", "odeint",
           "odeint_adjoint", "NDCN", "ODEBlock", "ODEBlock2", "ODEFunc", "GeneDynamics", "HeatDiffusion",
           "MutualDynamics"]
__version__ = "0.1.0"
