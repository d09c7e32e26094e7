#!/usr/bin/env python
"""One training iteration of the dynamics scripts (heat_dynamics.py:313-334: NDCN forward over the
training times with --method euler, L1 loss, backward) on the fused fixed-grid training path
(fused forward + discrete adjoint on ndcn_rhs_vjp_f32) vs the op-by-op autograd path."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ndcn_b200 as nb  # noqa: E402
from ndcn_b200 import autograd_solver, workloads as wl  # noqa: E402


def run(n_side, H, T, method, reps=10):
    dev = torch.device("cuda")
    phi = wl.graph_operator(wl.grid_adjacency(n_side), "norm_lap")
    n = phi.shape[0]
    torch.manual_seed(0)
    model = nb.NDCN(1, H, wl.to_reference_coo(phi), 1, method=method).to(dev)
    x0 = torch.rand(n, 1, device=dev) * 10
    t = torch.linspace(0, 5.0, T, device=dev)
    truth = torch.rand(T, n, 1, device=dev)
    func = model.neural_dynamic_layer.odefunc

    def step(fused):
        model.zero_grad(set_to_none=True)
        h0 = model.input_layer(x0)
        if fused:
            hv = nb.odeint(func, h0, t, method=method)
        else:
            hv = autograd_solver.solve(func, h0, t.type_as(h0), 0.01, 0.001, method)
        loss = torch.nn.functional.l1_loss(model.output_layer(hv), truth)
        loss.backward()
        return float(loss)

    out = {}
    for fused in (False, True):
        step(fused); step(fused)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            l = step(fused)
        torch.cuda.synchronize()
        out[fused] = ((time.perf_counter() - t0) / reps * 1e3, l)
    print("N=%d H=%d T=%d %s: op-by-op autograd %.1f ms/iter (loss %.6f), fused path %.1f ms/iter (loss %.6f), x%.2f" %
          (n, H, T, method, out[False][0], out[False][1], out[True][0], out[True][1], out[False][0] / out[True][0]))


if __name__ == "__main__":
    run(20, 20, 80, "euler")        # BASELINE config 1
    run(20, 20, 80, "rk4")
    run(100, 64, 80, "euler")       # 10k nodes
    run(316, 256, 20, "rk4", reps=3)  # ~100k nodes, H=256 (config 3's model in training)
