"""GPU: the reference's UNMODIFIED scripts run on this backend through the launcher
(``python -m ndcn_b200.run <script> ...``) and print what the reference prints on its own CPU path.

The script files come from ``baseline/_ref`` (git-ignored copy of the reference made by
``baseline/make_ref.py``; it ships with the gpurun snapshot).  Goldens: ``tests/golden/script_runs.json``,
produced by ``tests/golden/make_script_golden.py`` from /root/reference on the CPU with the same seed.
BASELINE.json configs 1 (heat_dynamics.py, 400-node grid) and 2 (dgnn.py, Cora, README flags).
"""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import GOLDEN, ROOT

sys.path.insert(0, os.path.join(GOLDEN))
from make_script_golden import parse_dgnn, parse_dynamics  # noqa: E402

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "baseline", "_ref")


def _need_ref():
    if not os.path.isfile(os.path.join(REF, "heat_dynamics.py")):
        pytest.skip("baseline/_ref is empty (run `python baseline/make_ref.py` where /root/reference exists)")


def _golden():
    with open(os.path.join(GOLDEN, "script_runs.json")) as f:
        return json.load(f)


def _launch(script, args, save=None, timeout=600):
    env = dict(os.environ, NDCN_RUN_SEED="0", PYTHONPATH=ROOT)
    if save:
        env["NDCN_RUN_SAVE"] = save
    cmd = [sys.executable, "-m", "ndcn_b200.run", os.path.join(REF, script)] + list(args)
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, env=env, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + "\n" + res.stderr[-3000:]
    return res.stdout


def test_heat_dynamics_script_unmodified(tmp_path):
    """heat_dynamics.py:207-209 (ground truth through our odeint), :248 (NDCN), :313-344 (training loop with
    backward through the Euler solver, evaluation every 20 iterations)."""
    _need_ref()
    g = _golden()["heat"]
    save = str(tmp_path / "truth.npy")
    out = _launch("heat_dynamics.py", g["args"] + ["--gpu", "0"], save=save)
    truth = np.load(save)
    ref = np.load(os.path.join(GOLDEN, "truth_heat.npz"))["sol_dense"]
    np.testing.assert_allclose(truth, ref, rtol=1e-4, atol=1e-4)  # dopri5 at rtol 1e-7 in fp32: round-off level
    lines = parse_dynamics(out)
    assert [l["iter"] for l in lines] == [l["iter"] for l in g["lines"]]
    for got, want in zip(lines, g["lines"]):
        # 20 / 40 Adam iterations on fp32 gradients computed in a different summation order
        assert got["train"] == pytest.approx(want["train"], rel=2e-2), (got, want)
        assert got["test"] == pytest.approx(want["test"], rel=2e-2), (got, want)


def test_gene_dynamics_script_rk4_sparse():
    """gene_dynamics.py with --sparse (COO operators) and --method rk4 (3/8 rule, fused discrete adjoint)."""
    _need_ref()
    g = _golden()["gene_rk4_sparse"]
    out = _launch("gene_dynamics.py", g["args"] + ["--gpu", "0"])
    lines = parse_dynamics(out)
    assert [l["iter"] for l in lines] == [l["iter"] for l in g["lines"]]
    for got, want in zip(lines, g["lines"]):
        assert got["train"] == pytest.approx(want["train"], rel=2e-2), (got, want)
        assert got["test"] == pytest.approx(want["test"], rel=2e-2), (got, want)


def test_dgnn_script_unmodified():
    """dgnn.py:159-237 with the README flags (Cora, hidden 256, dopri5 rtol=atol=.1, no_control, alpha 0):
    Sequential(Linear, Tanh, ODEBlock2(ODEFunc), Linear), trained through dopri5."""
    _need_ref()
    g = _golden()["dgnn"]
    out = _launch("dgnn.py", g["args"])
    got = parse_dgnn(out)
    assert len(got["epochs"]) == len(g["epochs"])
    for a, b in zip(got["epochs"], g["epochs"]):
        assert a["loss_train"] == pytest.approx(b["loss_train"], abs=5e-3), (a, b)
        assert a["loss_val"] == pytest.approx(b["loss_val"], abs=5e-3), (a, b)
        assert a["acc_train"] == pytest.approx(b["acc_train"], abs=0.03), (a, b)
    assert got["test_loss"] == pytest.approx(g["test_loss"], abs=5e-3)
    assert got["test_acc"] == pytest.approx(g["test_acc"], abs=0.03)
