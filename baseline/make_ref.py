#!/usr/bin/env python
"""Place an UNMODIFIED copy of the reference's hot-path files under baseline/_ref/ (git-ignored,
NOT gpurun-ignored: it travels to the GPU box with the snapshot, where /root/reference does not exist).

Used by
  * ``bench.py --impl reference``: times the reference's own ``torchdiffeq.odeint(neural_dynamics.ODEFunc ...)``
    on the box's host cores (``cpu_baseline.kind = "reference"``);
  * ``tests/test_gpu_scripts.py``: runs the unmodified ``heat_dynamics.py`` / ``dgnn.py`` through
    ``python -m ndcn_b200.run`` on the B200 backend.
Nothing under ``ndcn_b200/`` reads this directory.  The reference is a directory of scripts without
packaging, so "installing" it is a file copy: the top-level ``*.py`` modules, the vendored ``torchdiffeq``
package and ``data/cora`` (576 KB).
"""
from __future__ import annotations

import glob
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("NDCN_REFERENCE_ROOT", "/root/reference")
DST = os.path.join(HERE, "_ref")


def available() -> bool:
    return os.path.isfile(os.path.join(DST, "neural_dynamics.py")) and \
        os.path.isfile(os.path.join(DST, "torchdiffeq", "__init__.py"))


def make(force: bool = False) -> str:
    if not os.path.isfile(os.path.join(REF, "neural_dynamics.py")):
        if available():
            return DST
        raise RuntimeError("reference tree not found at %s and baseline/_ref is empty" % REF)
    if available() and not force:
        return DST
    os.makedirs(DST, exist_ok=True)
    for src in glob.glob(os.path.join(REF, "*.py")):
        shutil.copy2(src, os.path.join(DST, os.path.basename(src)))
    shutil.copytree(os.path.join(REF, "torchdiffeq"), os.path.join(DST, "torchdiffeq"), dirs_exist_ok=True)
    shutil.copytree(os.path.join(REF, "data", "cora"), os.path.join(DST, "data", "cora"), dirs_exist_ok=True)
    return DST


if __name__ == "__main__":
    print(make(force="--force" in sys.argv))
