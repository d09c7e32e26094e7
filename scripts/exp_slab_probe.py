#!/usr/bin/env python
"""Writes the bench graph (1M-node power-law, normalized Laplacian CSR) as a flat binary for
tests/cuda/slab_gather_probe.cu:  int64 n, int64 nnz, int32 rowptr[n+1], int32 col[nnz], f32 val[nnz]."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndcn_b200 import workloads as wl  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "power_law"
out = sys.argv[3] if len(sys.argv) > 3 else "gpurun_out/graph.bin"
a = wl.power_law_adjacency(n, 5, seed=0) if kind == "power_law" else wl.erdos_renyi_adjacency(n, 10.0, seed=0)
phi = wl.graph_operator(a, "norm_lap")
with open(out, "wb") as f:
    np.array([phi.shape[0], phi.nnz], np.int64).tofile(f)
    phi.indptr.astype(np.int32).tofile(f)
    phi.indices.astype(np.int32).tofile(f)
    phi.data.astype(np.float32).tofile(f)
print("wrote", out, phi.shape[0], phi.nnz)
