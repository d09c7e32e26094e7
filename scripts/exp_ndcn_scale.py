#!/usr/bin/env python
"""The reference's model-level call at the north-star scale: NDCN(1, 256, Phi, 1).forward(t, x0) with
100 output times on a 1M-node power-law graph (heat_dynamics.py:248,344 at 2500x the script's size).
The reference returns output_layer applied to a [100, 1M, 256] slab (102 GB); here the decoder is fused
into the emission kernels and only [100, 1M, 1] (400 MB) is written."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ndcn_b200 as nb  # noqa: E402
from ndcn_b200 import solver, workloads as wl  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    dev = torch.device("cuda")
    phi = wl.graph_operator(wl.power_law_adjacency(n, 5, seed=0), "norm_lap")
    torch.manual_seed(0)
    model = nb.NDCN(1, 256, wl.to_reference_coo(phi), 1, rtol=.01, atol=.001, method="dopri5")
    model.neural_dynamic_layer.odefunc.wt.weight.data.mul_(0.5)
    model = model.to(dev).eval()
    x0 = torch.rand(n, 1, device=dev) * 10
    t = torch.linspace(0, 5.0, 100, device=dev)
    with torch.no_grad():
        model(t[:3], x0)  # builds the CSR operator once, warms the allocator
        torch.cuda.synchronize()
        torch.cuda.reset_peak_memory_stats()
        t0 = time.perf_counter()
        pred = model(t, x0)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    info = solver.last_solve_info
    print("NDCN forward, N=%d, H=256, T=100 outputs: %.1f ms, output %s (%.0f MB), nfe=%d accepted=%d rejected=%d, "
          "peak GPU memory %.1f GB, finite=%s" %
          (n, dt * 1e3, tuple(pred.shape), pred.numel() * 4 / 1e6, info.nfe, info.n_accepted, info.n_rejected,
           torch.cuda.max_memory_allocated() / 1e9, bool(torch.isfinite(pred).all())))


if __name__ == "__main__":
    main()
