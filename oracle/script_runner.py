#!/usr/bin/env python
"""Run an UNMODIFIED reference script on the reference's own CPU code path (test infrastructure).

    python oracle/script_runner.py <root> <seed> <save.npy|-> script.py [script args ...]

<root> holds the reference files (/root/reference in the build container, baseline/_ref on the GPU box).
Applies the environment shims of SURVEY.md appendix A (no arithmetic touched), seeds torch/numpy (the dynamics
scripts do not), runs the script as __main__ and optionally saves its ``solution_numerical``.
"""
from __future__ import annotations

import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))


def main():
    root, seed, save, script = os.path.abspath(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4]
    save = None if save == "-" else os.path.abspath(save)
    os.environ["NDCN_REFERENCE_ROOT"] = root
    from oracle import ref_loader

    ref_loader.REFERENCE_ROOT = root
    ref_loader._apply_env_shims()
    import numpy as np
    import torch

    sys.path.insert(0, root)
    np.random.seed(seed)
    torch.manual_seed(seed)
    sys.argv = [os.path.join(root, script)] + sys.argv[5:]
    os.chdir(root)
    g = runpy.run_path(os.path.join(root, script), run_name="__main__")
    if save and "solution_numerical" in g:
        np.save(save, g["solution_numerical"].detach().cpu().numpy())


if __name__ == "__main__":
    main()
