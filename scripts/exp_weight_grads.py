#!/usr/bin/env python
"""Device time of dW = gp^T z, db = sum gp (ndcn_weight_grads_f32) against torch's fp32 matmul (cuBLAS SGEMM)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from ndcn_b200 import solver  # noqa: E402


def timed(fn, reps=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


if __name__ == "__main__":
    for n, H in ((8192, 128), (99856, 256), (1_000_000, 256)):
        gp = torch.randn(n, H, device="cuda")
        z = torch.randn(n, H, device="cuda")
        dW = torch.zeros(H, H, device="cuda")
        db = torch.zeros(H, device="cuda")
        ours = timed(lambda: solver.weight_grads(gp, z, dW, db, accumulate=True))
        ref = timed(lambda: (dW.addmm_(gp.t(), z), db.add_(gp.sum(0))))
        want = gp.double().t() @ z.double()
        solver.weight_grads(gp, z, dW, db, accumulate=False)
        err = float((dW.double() - want).norm() / want.norm())
        err_t = float(((gp.t() @ z).double() - want).norm() / want.norm())
        print("n=%d H=%d: ndcn_weight_grads_f32 %.3f ms (%.1f TFLOP/s, rel err %.1e), torch addmm_ + sum %.3f ms (rel err %.1e), NDCN_WG_MMA=%s"
              % (n, H, ours, 2.0 * n * H * H / ours / 1e9, err, ref, err_t, os.environ.get("NDCN_WG_MMA", "1")))
