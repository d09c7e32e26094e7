"""Import the upstream reference in-process (build container only).

TEST INFRASTRUCTURE ONLY.  ``/root/reference`` does not exist on the GPU box, so
nothing that runs there may depend on this module succeeding: callers check
``reference_available()`` first.  The three environment shims below touch no
hot-path arithmetic: a matplotlib stub, two networkx
aliases removed in networkx>=3, and a scipy ``csr_matrix((data, zip(...)))``
wrapper needed by ``utils.py:193``.
"""
from __future__ import annotations

import contextlib
import io
import os
import runpy
import sys
import types

REFERENCE_ROOT = os.environ.get("NDCN_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "neural_dynamics.py"))


_shimmed = False


def _apply_env_shims():
    global _shimmed
    if _shimmed:
        return
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.cm", "matplotlib.ticker",
                 "mpl_toolkits", "mpl_toolkits.mplot3d"):
        try:
            __import__(name)
        except Exception:
            sys.modules.setdefault(name, types.ModuleType(name))
    tick = sys.modules["matplotlib.ticker"]
    for attr in ("LinearLocator", "FormatStrFormatter"):
        if not hasattr(tick, attr):
            setattr(tick, attr, object)
    if not hasattr(sys.modules["mpl_toolkits.mplot3d"], "Axes3D"):
        sys.modules["mpl_toolkits.mplot3d"].Axes3D = object

    import networkx as nx
    import scipy.sparse as sp

    if not hasattr(nx, "to_scipy_sparse_matrix"):
        nx.to_scipy_sparse_matrix = lambda G, format="coo", **k: sp.coo_matrix(
            nx.to_scipy_sparse_array(G, format=format))
        nx.from_scipy_sparse_matrix = nx.from_scipy_sparse_array

    base = sp.csr_matrix
    if not getattr(base, "_ndcn_zip_ok", False):
        class _CsrAcceptsZip(base):
            _ndcn_zip_ok = True

            def __init__(self, arg1, *a, **k):
                if isinstance(arg1, tuple) and len(arg1) == 2 and isinstance(arg1[1], zip):
                    arg1 = (arg1[0], tuple(arg1[1]))
                super().__init__(arg1, *a, **k)

        sp.csr_matrix = _CsrAcceptsZip
    _shimmed = True


def import_reference():
    """Returns (neural_dynamics, torchdiffeq) modules OF THE REFERENCE."""
    assert reference_available(), "reference tree not present at " + REFERENCE_ROOT
    _apply_env_shims()
    # make sure we do not pick up this repo's same-named shims
    for name in list(sys.modules):
        if name in ("neural_dynamics", "utils", "propagation", "utils_in_learn_dynamics") or \
                name == "torchdiffeq" or name.startswith("torchdiffeq."):
            mod = sys.modules[name]
            if REFERENCE_ROOT not in (getattr(mod, "__file__", "") or ""):
                del sys.modules[name]
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import neural_dynamics  # noqa
    import torchdiffeq  # noqa
    assert REFERENCE_ROOT in neural_dynamics.__file__ and REFERENCE_ROOT in torchdiffeq.__file__
    return neural_dynamics, torchdiffeq


def run_reference_script(script: str, argv):
    """runpy the unmodified script with run_name != '__main__' (its training loop is under
    ``if __name__ == '__main__'``) and harvest module-level objects: RHS classes, A, L, OM,
    x0, t, solution_numerical, model ..."""
    import_reference()
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [script] + list(argv)
    os.chdir(REFERENCE_ROOT)
    try:
        with contextlib.redirect_stdout(io.StringIO()):
            return runpy.run_path(os.path.join(REFERENCE_ROOT, script), run_name="oracle_harvest")
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)
