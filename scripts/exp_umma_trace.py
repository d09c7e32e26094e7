#!/usr/bin/env python
"""Device-side timeline of CTA 0 of the tcgen05 stage kernel (ndcn_debug_umma_trace): where do the
roles wait?  Prints a merged event list for the first tiles and per-role summaries.
Events: 0 loader: stage free | 1 producer: stage free | 2 producer: A atom published |
        3 MMA: stage full | 4 MMA: atom committed | 5 MMA: accumulator free |
        6 epilogue: accumulator full | 7 epilogue: chunk done"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import ndcn_b200 as nb  # noqa: E402
from ndcn_b200 import _ffi, workloads as wl  # noqa: E402

NAMES = {0: "LD stage-free", 1: "PR stage-free", 2: "PR published", 3: "MMA stage-full", 4: "MMA committed",
         5: "MMA acc-free", 6: "EPI acc-full", 7: "EPI chunk-done"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--mode", default="store", choices=["store", "stage4"])
    args = ap.parse_args()
    dev = torch.device("cuda")
    n, H = args.nodes, 256
    torch.manual_seed(0)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5).to(dev), lin.bias.detach().to(dev)
    x = torch.randn(n, H, device=dev)
    a = wl.grid_adjacency(int(round(n ** 0.5)))
    g = nb.CsrGraph.from_scipy(wl.graph_operator(a, "norm_lap"), dev)
    n = g.n_rows
    x = x[:n].contiguous()
    _ffi.configure(stage_impl=_ffi.IMPL_UMMA, gather_cw=-1)
    spec = nb.RhsSpec.ndcn(H, W, b, no_graph=True)
    buf = torch.zeros(4 * 4096, dtype=torch.int64, device=dev)
    lib = _ffi.lib()

    def run():
        if args.mode == "store":
            nb.rhs_eval(g, spec, x)
        else:  # a dopri5 solve: the traced launch is the last stage kernel (error-estimate stage)
            t = torch.tensor([0.0, 0.05 * 1.5], dtype=torch.float64)
            nb.odeint_fused(g, spec, x, t, method="dopri5", forced_dt=0.05, terminal_only=True)

    run()
    torch.cuda.synchronize()
    _ffi.check(lib.ndcn_debug_umma_trace(buf.data_ptr()))
    run()
    torch.cuda.synchronize()
    _ffi.check(lib.ndcn_debug_umma_trace(None))
    t = buf.cpu().numpy().astype(np.uint64).reshape(4, 4096)
    evs = []
    for role in range(4):
        cnt = int(t[role, 0])
        for e in t[role, 1:1 + cnt]:
            evs.append((int(e & np.uint64(0x00ffffffffffffff)), int(e >> np.uint64(56)), role))
    evs.sort()
    t0 = evs[0][0]
    print("events:", len(evs), " span: %.1f us (at 1.965 GHz)" % ((evs[-1][0] - t0) / 1965.0))
    for clk, code, role in evs[:150]:
        print("%9d cyc  %8.2f us  %s" % (clk - t0, (clk - t0) / 1965.0, NAMES[code]))
    # summaries
    by = {}
    for clk, code, role in evs:
        by.setdefault(code, []).append(clk)
    for code, lst in sorted(by.items()):
        d = np.diff(np.array(lst, dtype=np.int64))
        if len(d):
            print("%-16s n=%5d  mean gap %8.0f cyc  median %8.0f  p90 %8.0f" %
                  (NAMES[code], len(lst), d.mean(), np.median(d), np.percentile(d, 90)))
    # MMA: wait for full (stage-full minus previous commit) ; producer: time from stage-free to published
    if 1 in by and 2 in by:
        m = min(len(by[1]), len(by[2]))
        d = np.array(by[2][:m]) - np.array(by[1][:m])
        print("producer convert+store: mean %.0f cyc" % d.mean())
    if 3 in by and 4 in by:
        m = min(len(by[3]), len(by[4]))
        d = np.array(by[4][:m]) - np.array(by[3][:m])
        print("MMA issue (full -> committed): mean %.0f cyc" % d.mean())
        d2 = np.array(by[3][1:m]) - np.array(by[4][:m - 1])
        print("MMA waiting for next full stage: mean %.0f cyc" % d2.mean())


if __name__ == "__main__":
    main()
