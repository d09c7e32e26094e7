"""How many of the gather's row fetches can the 126 MB L2 absorb?  (host-only experiment, no GPU)

The full-row gather (k_stage_ndcn_row) reads one 1 KB state row per stored entry of Phi.  This script replays
those fetches, in the order the grid walks the rows, through an LRU cache of C rows (scripts/l2sim/lru.c) for the
bench graph (1M-node power-law, generation order) and for several row PROCESSING orders -- only the order in which
warps take rows changes, the data layout and every result stay the same.

Calibration: C = 90 000 rows (92 MB of the 126 MB L2) reproduces the L2 hit rate ncu measures for the kernel
(15.9 %, profiles/r01_rhs_kernels_ncu_summary.json).  Result (profiles/README.md): no order gets past 21 % --
a preferential-attachment graph is an expander; the re-read traffic of the gather is a property of the graph and of
1 GB of state against 126 MB of L2, not of the kernel.

    python scripts/l2sim/run.py [--nodes 1000000]
"""
import argparse
import os
import subprocess
import sys
import tempfile

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndcn_b200 import workloads as wl  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--caps", type=int, nargs="+", default=[30000, 60000, 90000, 120000],
                    help="LRU capacities (rows of 1 KB) tried for the generation order")
    ap.add_argument("--cap", type=int, default=90000, help="capacity used to compare the processing orders")
    args = ap.parse_args()
    n = args.nodes
    tmp = tempfile.mkdtemp(prefix="l2sim_")
    exe = os.path.join(tmp, "lru")
    subprocess.check_call(["gcc", "-O2", "-o", exe, os.path.join(ROOT, "scripts", "l2sim", "lru.c")])
    phi = wl.graph_operator(wl.power_law_adjacency(n, 5, 0), "norm_lap")
    rp, col = phi.indptr.astype(np.int64), phi.indices.astype(np.int32)
    rp.tofile(os.path.join(tmp, "rp.bin"))
    col.tofile(os.path.join(tmp, "col.bin"))
    deg = np.diff(rp)
    rows = np.repeat(np.arange(n), deg)
    offd = col != rows

    def run(name, order, caps):
        order.astype(np.int32).tofile(os.path.join(tmp, "ord.bin"))
        out = subprocess.run([exe, os.path.join(tmp, "rp.bin"), os.path.join(tmp, "col.bin"), os.path.join(tmp, "ord.bin"),
                              str(n)] + [str(c) for c in caps], capture_output=True, text=True).stdout.strip()
        for line in out.splitlines():
            print("%-22s %s" % (name, line), flush=True)

    run("generation order", np.arange(n), args.caps)
    big = np.iinfo(np.int32).max
    c2 = np.where(offd, col, big)
    minnb = np.minimum.reduceat(c2, rp[:-1])
    run("by smallest neighbour", np.argsort(minnb, kind="stable"), [args.cap])
    key = np.where(offd, -deg[col].astype(np.int64) * n + col, 0)
    run("by biggest-hub nbr", np.argsort(np.minimum.reduceat(key, rp[:-1]), kind="stable"), [args.cap])
    from scipy.sparse.csgraph import breadth_first_order, reverse_cuthill_mckee
    A = sp.csr_matrix((np.ones(len(col), np.float32), col, rp), shape=(n, n))
    o, _ = breadth_first_order(A, 0, directed=False, return_predecessors=True)
    run("breadth first", np.concatenate([o, np.setdiff1d(np.arange(n), o)]), [args.cap])
    run("reverse Cuthill-McKee", reverse_cuthill_mckee(A, symmetric_mode=True), [args.cap])
    run("degree descending", np.argsort(-deg, kind="stable"), [args.cap])


if __name__ == "__main__":
    main()
