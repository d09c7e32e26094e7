#!/usr/bin/env python
"""Kernel-level timing of the RHS building blocks on the bench workload (1M-node power-law graph,
H=256): gather variants (full rows vs chunk-major at several widths), the tcgen05 GEMM kernel
alone (no_graph), and the complete RHS for both kernel families.  CUDA events on the launch
stream, L2 state between repetitions is whatever the previous repetition left (inputs are 8x
larger than L2).  Usage: python scripts/exp_kernels.py [--nodes N] [--graph power_law|er]"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import ndcn_b200 as nb  # noqa: E402
from ndcn_b200 import _ffi, workloads as wl  # noqa: E402


QUICK = False


def timed(fn, reps=5, warm=2):
    if QUICK:
        reps, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--graph", default="power_law")
    ap.add_argument("--quick", action="store_true", help="one untimed pass per variant (for ncu captures)")
    ap.add_argument("--spmm-only", action="store_true")
    args = ap.parse_args()
    global QUICK
    QUICK = args.quick
    dev = torch.device("cuda")
    n, H = args.nodes, args.hidden
    if args.graph == "power_law":
        a = wl.power_law_adjacency(n, 5, seed=0)
    elif args.graph == "grid":
        a = wl.grid_adjacency(int(round(n ** 0.5)))
        n = a.shape[0]
    else:
        a = wl.erdos_renyi_adjacency(n, 10.0, seed=0)
    phi = wl.graph_operator(a, "norm_lap")
    g = nb.CsrGraph.from_scipy(phi, dev)
    nnz = g.nnz
    torch.manual_seed(0)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5).to(dev), lin.bias.detach().to(dev)
    x = torch.randn(n, H, device=dev)
    res = {"nodes": n, "hidden": H, "nnz": nnz, "graph": args.graph}
    gather_bytes = nnz * H * 4

    ref = None
    for cw, ver in ((-1, 1), (32, 1), (32, 2), (64, 2)):
        _ffi.configure(gather_cw=cw, gather_version=ver)
        med, best = timed(lambda: nb.spmm(g, x))
        out = nb.spmm(g, x)
        if ref is None:
            ref = out
        diff = float((out - ref).abs().max())
        res["spmm_cw%d_v%d" % (cw, ver)] = {"ms": med, "best_ms": best, "gathered_GBps": gather_bytes / med / 1e6,
                                 "max_abs_diff_vs_fullrow": diff}
        print("spmm cw=%d v%d: %.3f ms (best %.3f)  gathered %.0f GB/s  diff %.2e" %
              (cw, ver, med, best, gather_bytes / med / 1e6, diff), flush=True)
    del ref
    if args.spmm_only:
        print(json.dumps(res))
        return

    spec = nb.RhsSpec.ndcn(H, W, b)
    spec_ng = nb.RhsSpec.ndcn(H, W, b, no_graph=True)
    _ffi.configure(stage_impl=_ffi.IMPL_SIMT, gather_cw=-1)
    med, best = timed(lambda: nb.rhs_eval(g, spec, x))
    simt = nb.rhs_eval(g, spec, x)
    res["rhs_simt"] = {"ms": med, "best_ms": best}
    print("rhs simt fused: %.3f ms" % med, flush=True)
    med, best = timed(lambda: nb.rhs_eval(g, spec_ng, x))
    res["rhs_simt_no_graph"] = {"ms": med, "best_ms": best}
    print("rhs simt no_graph (GEMM+epilogue only): %.3f ms" % med, flush=True)

    _ffi.configure(stage_impl=_ffi.IMPL_UMMA, gather_cw=0)
    med, best = timed(lambda: nb.rhs_eval(g, spec_ng, x))
    res["rhs_umma_no_graph"] = {"ms": med, "best_ms": best, "GBps_2NH": 2 * n * H * 4 / med / 1e6}
    print("rhs umma no_graph (GEMM+epilogue only): %.3f ms  (%.0f GB/s of read+write)" %
          (med, 2 * n * H * 4 / med / 1e6), flush=True)
    for cw in (-1, 32, 64):
        _ffi.configure(stage_impl=_ffi.IMPL_UMMA, gather_cw=cw, gather_version=2)
        med, best = timed(lambda: nb.rhs_eval(g, spec, x))
        out = nb.rhs_eval(g, spec, x)
        d = float((out - simt).abs().max())
        res["rhs_umma_cw%d" % cw] = {"ms": med, "best_ms": best, "max_abs_diff_vs_simt": d}
        print("rhs umma cw=%d: %.3f ms  max|umma - simt| = %.2e (max|k| %.2f)" %
              (cw, med, d, float(simt.abs().max())), flush=True)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
