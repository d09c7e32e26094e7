"""Launcher: run an UNMODIFIED reference script on the B200 backend.

    python -m ndcn_b200.run /path/to/ndcn/heat_dynamics.py --network grid --T 5 --baseline ndcn --gpu 0

``python script.py`` puts the script's own directory first on sys.path, so the reference's
``neural_dynamics`` / ``torchdiffeq`` would shadow any PYTHONPATH entry; here the shim directory
is inserted in front of it and the script is executed with runpy as ``__main__``.
Environment shims that touch no arithmetic are applied when needed:
a matplotlib stub if matplotlib is absent, networkx>=3 / scipy>=1.1x compatibility aliases.
"""
from __future__ import annotations

import os
import runpy
import sys
import types


def _env_shims():
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
    try:
        import networkx as nx
        import scipy.sparse as sp

        if not hasattr(nx, "to_scipy_sparse_matrix"):
            nx.to_scipy_sparse_matrix = lambda G, format="coo", **k: sp.coo_matrix(
                nx.to_scipy_sparse_array(G, format=format))
            nx.from_scipy_sparse_matrix = nx.from_scipy_sparse_array
        base = sp.csr_matrix
        if not getattr(base, "_accepts_zip", False):
            class _Csr(base):
                _accepts_zip = True

                def __init__(self, arg1, *a, **k):
                    if isinstance(arg1, tuple) and len(arg1) == 2 and isinstance(arg1[1], zip):
                        arg1 = (arg1[0], tuple(arg1[1]))
                    super().__init__(arg1, *a, **k)

            sp.csr_matrix = _Csr
    except ImportError:
        pass


def install(script_dir: str) -> None:
    """Make ``import neural_dynamics`` / ``import torchdiffeq`` resolve to the shims while the
    rest of the script's directory (utils, propagation, ...) stays importable."""
    shim_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "shims")
    repo_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (repo_root, script_dir, shim_dir):  # last inserted wins: shim_dir first
        if p in sys.path:
            sys.path.remove(p)
        sys.path.insert(0, p)
    for name in [n for n in sys.modules if n == "neural_dynamics" or n == "torchdiffeq" or n.startswith("torchdiffeq.")]:
        del sys.modules[name]
    # solvers outside the accelerated path (--method tsit5 | adams | ...) are handed to the script's own torchdiffeq
    import importlib

    importlib.import_module("ndcn_b200.odeint").register_out_of_scope_solver(script_dir)  # the package re-exports the function under that name


def run_script(argv, seed=None, run_name="__main__"):
    """Run the unmodified script ``argv[0]`` with arguments ``argv[1:]`` on this backend; returns the
    script's module globals (``solution_numerical``, ``model``, ... for the dynamics scripts).

    ``seed``: the dynamics scripts never seed torch (heat_dynamics.py:82,89 seed only networkx), so two runs
    differ in their weight initialisation; a seed given here (or NDCN_RUN_SEED in the environment) is applied
    to torch and numpy BEFORE the script starts -- the script itself stays untouched."""
    script = os.path.abspath(argv[0])
    _env_shims()
    install(os.path.dirname(script))
    if seed is None and os.environ.get("NDCN_RUN_SEED"):
        seed = int(os.environ["NDCN_RUN_SEED"])
    if seed is not None:
        import numpy as np
        import torch

        np.random.seed(seed)
        torch.manual_seed(seed)
    old_argv, old_cwd = sys.argv, os.getcwd()
    sys.argv = [script] + list(argv[1:])
    os.chdir(os.path.dirname(script))  # dgnn.py reads data/<dataset>/... relative to cwd (utils.py:122)
    try:
        return runpy.run_path(script, run_name=run_name)
    finally:
        sys.argv = old_argv
        os.chdir(old_cwd)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        print(__doc__)
        return 2
    g = run_script(argv)
    save = os.environ.get("NDCN_RUN_SAVE")  # test hook: persist the ground-truth tensor the script integrated
    if save and "solution_numerical" in g:
        import numpy as np

        np.save(save, g["solution_numerical"].detach().cpu().numpy())
    return 0


if __name__ == "__main__":
    sys.exit(main())
