"""Stand-in for the reference's vendored ``torchdiffeq`` package (torchdiffeq/__init__.py:1-2):
same two public names, backed by ndcn_b200's CUDA solver."""
from ndcn_b200.odeint import odeint, odeint_adjoint  # noqa: F401

__all__ = ["odeint", "odeint_adjoint"]
