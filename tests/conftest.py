import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name + ".npz")))


def csr_to_dense(g, prefix):
    n = int(g[prefix + "_n"])
    rp, col, val = g[prefix + "_rowptr"], g[prefix + "_col"], g[prefix + "_val"]
    csr = torch.sparse_csr_tensor(torch.from_numpy(rp.astype(np.int64)), torch.from_numpy(col.astype(np.int64)),
                                  torch.from_numpy(val), size=(n, n))
    return csr.to_dense()


def csr_to_coo(g, prefix):
    """uncoalesced-style COO in row-major entry order (the reference's at-scale format)."""
    n = int(g[prefix + "_n"])
    rp, col, val = g[prefix + "_rowptr"], g[prefix + "_col"], g[prefix + "_val"]
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(rp))
    idx = torch.from_numpy(np.vstack((rows, col.astype(np.int64))))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(val), (n, n))


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def get(name):
        if name not in cache:
            cache[name] = load_golden(name)
        return cache[name]

    return get
