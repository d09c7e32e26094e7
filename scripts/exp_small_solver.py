#!/usr/bin/env python
"""Persistent whole-solve kernel vs launch-per-stage path on BASELINE configs 1-2 (one B200).

config 1: NDCN(1, 20, OM, 1) forward over 100 output times on the 400-node grid, Euler (heat_dynamics.py:344)
config 2: the Cora ODE block of dgnn.py (H=256, dopri5 rtol=atol=.1, terminal state), no_control and control
Prints device ms per call (CUDA events, 200 calls) and launches per call."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import ndcn_b200 as nb  # noqa: E402
from conftest import csr_to_coo, csr_to_dense, load_golden  # noqa: E402
from ndcn_b200 import solver  # noqa: E402


def timed(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


out = {}
g = load_golden("ndcn_grid400")
OM = csr_to_dense(g, "OM")
W = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__weight"]).cuda()
b = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__bias"]).cuda()
h0, t = torch.from_numpy(g["h0"]).cuda(), torch.from_numpy(g["t"]).float()
graph = nb.CsrGraph.from_tensor(OM, torch.device("cuda"))
spec = nb.RhsSpec.ndcn(20, W, b)
Wd, bd = torch.randn(1, 20, device="cuda"), torch.randn(1, device="cuda")
for method in ("euler", "rk4", "dopri5"):
    for small in (True, False):
        f = lambda: nb.odeint_fused(graph, spec, h0, t, method=method, rtol=.01, atol=.001, small=small, decoder=(Wd, bd))  # noqa: E731
        ms = timed(f)
        out["grid400_H20_%s_%s" % (method, "persistent" if small else "staged")] = {
            "ms": round(ms, 4), "launches": solver.last_solve_info.n_launches, "nfe": solver.last_solve_info.nfe}

c = load_golden("cora_block")
x = torch.from_numpy(np.tanh(np.random.RandomState(11).standard_normal((2708, 256))).astype(np.float32)).cuda()
tt = torch.linspace(0, 1.2, 16).float()
for key, noctl in (("a00_h256_noctl", True), ("a05_h256_ctl", False)):
    gr = nb.CsrGraph.from_tensor(csr_to_coo(c, "adj_" + key[:3]), torch.device("cuda"))
    sp = nb.RhsSpec.ndcn(256, torch.from_numpy(c["W_" + key]).cuda(), torch.from_numpy(c["b_" + key]).cuda(), no_control=noctl)
    for small in (True, False):
        f = lambda: nb.odeint_fused(gr, sp, x, tt, method="dopri5", rtol=.1, atol=.1, terminal_only=True, small=small)  # noqa: E731
        ms = timed(f)
        out["cora_%s_%s" % (key, "persistent" if small else "staged")] = {
            "ms": round(ms, 4), "launches": solver.last_solve_info.n_launches, "nfe": solver.last_solve_info.nfe}

# ground truth of the dynamics scripts: dopri5 rtol 1e-7 on [400,1]
th = load_golden("truth_heat")
L = csr_to_dense(th, "L")
gh = nb.CsrGraph.from_tensor(-L, torch.device("cuda"))
x0 = torch.from_numpy(th["x0"]).cuda()
for small in (True, False):
    f = lambda: nb.odeint_fused(gh, nb.RhsSpec.heat(1, 1), x0, torch.from_numpy(th["t"]), method="dopri5", small=small)  # noqa: E731
    ms = timed(f, reps=50)
    out["truth_heat_%s" % ("persistent" if small else "staged")] = {
        "ms": round(ms, 4), "launches": solver.last_solve_info.n_launches, "nfe": solver.last_solve_info.nfe}
print(json.dumps(out, indent=1))
