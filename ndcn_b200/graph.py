"""Device-resident CSR form of the graph operator Phi.

The reference hands the operator to ``ODEFunc`` / ``HeatDiffusion`` / ... either as a dense
``[N, N]`` fp32 tensor (default) or as an *uncoalesced* sparse COO tensor with int64 indices
(``utils_in_learn_dynamics.py:193-201``, ``utils.py:12-23``) and dispatches on ``A.is_sparse``
at every RHS evaluation (``neural_dynamics.py:27-31``).  Here the operator is converted ONCE
to int32 CSR on the GPU and cached on the module that owns it.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch

from . import _ffi


def require_cuda(device: Optional[torch.device] = None) -> torch.device:
    """The hot path exists only on the GPU: fail loudly, never fall back to the CPU."""
    if not torch.cuda.is_available():
        raise RuntimeError(
            "ndcn_b200: no CUDA device is visible. The NDCN hot path is implemented as sm_100a CUDA "
            "kernels only; there is no CPU fallback.")
    if device is not None and device.type == "cuda":
        return device
    return torch.device("cuda", torch.cuda.current_device())


class CsrGraph:
    """int32 CSR (rowptr, col) + fp32 values on one CUDA device, plus the C handle.

    ``n_cols >= n_rows``: columns ``>= n_rows`` index halo rows of a 1-D row partition
    (``ndcn_b200.partition``); on a single GPU ``n_cols == n_rows``.
    """

    def __init__(self, rowptr: torch.Tensor, col: torch.Tensor, val: torch.Tensor, n_rows: int, n_cols: int):
        assert rowptr.is_cuda and col.is_cuda and val.is_cuda, "CSR arrays must live on the GPU"
        assert rowptr.dtype == torch.int32 and col.dtype == torch.int32 and val.dtype == torch.float32
        assert rowptr.numel() == n_rows + 1
        self.rowptr = rowptr.contiguous()
        self.col = col.contiguous()
        self.val = val.contiguous()
        self.n_rows = int(n_rows)
        self.n_cols = int(n_cols)
        self.nnz = int(col.numel())
        self.device = rowptr.device
        handle = C.c_void_p()
        with torch.cuda.device(self.device):
            rc = _ffi.lib().ndcn_graph_create(self.n_rows, self.n_cols, self.nnz, self.rowptr.data_ptr(),
                                              self.col.data_ptr() if self.nnz else None,
                                              self.val.data_ptr() if self.nnz else None, C.byref(handle))
        _ffi.check(rc, "ndcn_graph_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, _ffi.lib().ndcn_graph_destroy, handle)
        self._transpose: Optional["CsrGraph"] = None

    # ------------------------------------------------------------------------------
    @classmethod
    def from_tensor(cls, A: torch.Tensor, device: Optional[torch.device] = None) -> "CsrGraph":
        """Dense, sparse-COO (coalesced or not) or sparse-CSR torch tensor -> CsrGraph."""
        dev = require_cuda(device if device is not None else (A.device if A.is_cuda else None))
        if A.dim() != 2:
            raise ValueError("graph operator must be 2-D, got shape %s" % (tuple(A.shape),))
        n_rows, n_cols = int(A.shape[0]), int(A.shape[1])
        if n_rows != n_cols:
            raise ValueError("graph operator must be square, got %s" % (tuple(A.shape),))
        A = A.detach()
        if A.layout == torch.sparse_coo:
            csr = A.to(dev).to(torch.float32).coalesce().to_sparse_csr()
        elif A.layout == torch.sparse_csr:
            csr = A.to(dev).to(torch.float32)
        else:
            csr = A.to(dev).to(torch.float32).to_sparse_csr()
        return cls(csr.crow_indices().to(torch.int32), csr.col_indices().to(torch.int32),
                   csr.values().to(torch.float32), n_rows, n_cols)

    @classmethod
    def from_scipy(cls, m, device: Optional[torch.device] = None, n_cols: Optional[int] = None) -> "CsrGraph":
        import numpy as np

        dev = require_cuda(device)
        m = m.tocsr()
        m.sort_indices()
        return cls(torch.from_numpy(m.indptr.astype(np.int32)).to(dev),
                   torch.from_numpy(m.indices.astype(np.int32)).to(dev),
                   torch.from_numpy(m.data.astype(np.float32)).to(dev),
                   m.shape[0], n_cols if n_cols is not None else m.shape[1])

    def transpose(self) -> "CsrGraph":
        """Phi^T (for the backward of Phi x); cached.  Square single-GPU graphs only."""
        if self._transpose is None:
            assert self.n_rows == self.n_cols
            csr = torch.sparse_csr_tensor(self.rowptr.long(), self.col.long(), self.val,
                                          size=(self.n_rows, self.n_cols))
            t = csr.to_sparse_coo().t().coalesce().to_sparse_csr()
            self._transpose = CsrGraph(t.crow_indices().to(torch.int32), t.col_indices().to(torch.int32),
                                       t.values(), self.n_rows, self.n_cols)
        return self._transpose

    def algorithmic_bytes(self) -> int:
        """8E + 4(N+1): CSR val + col + rowptr."""
        return 8 * self.nnz + 4 * (self.n_rows + 1)


_GRAPH_CACHE_ATTR = "_ndcn_b200_graph_cache"


def cached_graph(owner, A: torch.Tensor, device: Optional[torch.device] = None, negate: bool = False) -> CsrGraph:
    """CsrGraph for operator tensor ``A``, cached on ``owner`` (an nn.Module) and keyed on the
    tensor's identity + in-place version, so the conversion is paid once per operator."""
    dev = require_cuda(device if device is not None else (A.device if A.is_cuda else None))
    key = (id(A), A._version, str(dev), bool(negate))
    cache = getattr(owner, _GRAPH_CACHE_ATTR, None)
    if cache is not None and cache[0] == key and cache[2] is A:
        return cache[1]
    g = CsrGraph.from_tensor(-A if negate else A, dev)
    try:
        object.__setattr__(owner, _GRAPH_CACHE_ATTR, (key, g, A))
    except Exception:
        pass
    return g
