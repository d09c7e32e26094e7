"""Stand-in for the reference's ``neural_dynamics`` module.

The reference module star-re-exports ``utils`` (neural_dynamics.py:5) and the scripts rely on
that for ``torch``, ``nn``, ``np``, ``nx`` ... (dgnn.py has ``import torch`` commented out,
dgnn.py:4).  So: re-export the reference's ``utils`` when it is importable (launcher case: the
script's own directory is on sys.path), then overlay the accelerated classes.  The discrete
baselines (``GraphConvolution``, ``TemporalGCN``, neural_dynamics.py:163-238) are out of scope
and are taken from the reference file itself when it is available.
"""
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
import numpy as np  # noqa: F401

try:  # the reference's helper namespace (only present next to the reference scripts)
    from utils import *  # noqa: F401,F403
except ImportError:
    pass

import torchdiffeq as ode  # noqa: F401,E402  (the shim next to this file)
from ndcn_b200.models import NDCN, ODEBlock, ODEBlock2, ODEFunc  # noqa: F401,E402


def _load_reference_baselines():
    """GraphConvolution / TemporalGCN stay the reference's own plain-PyTorch code."""
    import importlib.util
    import os
    import sys

    for d in sys.path:
        cand = os.path.join(d, "neural_dynamics.py")
        if os.path.isfile(cand) and os.path.abspath(cand) != os.path.abspath(__file__):
            spec = importlib.util.spec_from_file_location("_reference_neural_dynamics", cand)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            return mod
    return None


try:
    _ref = _load_reference_baselines()
except Exception:  # pragma: no cover
    _ref = None
if _ref is not None:
    GraphConvolution = _ref.GraphConvolution
    TemporalGCN = _ref.TemporalGCN
