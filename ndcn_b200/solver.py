"""Host side of the fused solve: tensor plumbing around ``ndcn_odeint_f32``.

PyTorch is used for device memory (state, output slab, workspace) and the current stream;
every arithmetic operation of the path happens inside libndcn_b200.so.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import threading
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional, Sequence, Tuple

import torch

from . import _ffi
from .graph import CsrGraph, require_cuda


@dataclass
class RhsSpec:
    """Which right-hand side to evaluate (mirrors ``ndcn_rhs_desc_t``)."""

    kind: int
    H: int
    flags: int = 0
    W: Optional[torch.Tensor] = None  # [H, H] nn.Linear weight (y = x W^T + b)
    b: Optional[torch.Tensor] = None  # [H]
    p: Sequence[float] = field(default_factory=lambda: (0.0,) * 8)
    callback: Optional[Callable] = None  # python callable (y_ptr, k_ptr, t_ptr) -> int

    @staticmethod
    def ndcn(H: int, W: Optional[torch.Tensor], b: Optional[torch.Tensor], no_graph=False, no_control=False,
             relu=True) -> "RhsSpec":
        flags = (_ffi.F_NO_GRAPH if no_graph else 0) | (_ffi.F_NO_CONTROL if no_control else 0) | \
                (0 if relu else _ffi.F_NO_RELU)
        return RhsSpec(_ffi.RHS_NDCN, H, flags, W, b)

    @staticmethod
    def heat(d: int, k: float = 1.0) -> "RhsSpec":
        return RhsSpec(_ffi.RHS_HEAT, d, p=(float(k),) + (0.0,) * 7)

    @staticmethod
    def gene(d: int, b: float = 1.0, f: float = 1.0, h: float = 2.0) -> "RhsSpec":
        return RhsSpec(_ffi.RHS_GENE, d, p=(float(b), float(f), float(h)) + (0.0,) * 5)

    @staticmethod
    def mutual(dim: int, b=0.1, k=5.0, c=1.0, d=5.0, e=0.9, h=0.1) -> "RhsSpec":
        return RhsSpec(_ffi.RHS_MUTUAL, dim, p=(float(b), float(k), float(c), float(d), float(e), float(h), 0.0, 0.0))

    def to_c(self, keep: list, prepare: bool = False) -> _ffi.RhsDesc:
        """``prepare``: also hand over the derived weight forms (W^T, tf32 images), cached per weight version --
        for the per-evaluation entry points (``rhs_eval`` / vjp); the solver derives its own once per solve."""
        d = _ffi.RhsDesc()
        d.kind, d.flags, d.H = self.kind, self.flags, self.H
        if self.W is not None:
            W = self.W.detach().to(torch.float32).contiguous()
            keep.append(W)
            d.W = W.data_ptr()
            if prepare and self.kind == _ffi.RHS_NDCN and W.is_cuda:
                prep = prepared_weights(self.W, W)
                keep.append(prep)
                d.prepared = prep.data_ptr()
        if self.b is not None:
            b = self.b.detach().to(torch.float32).contiguous()
            keep.append(b)
            d.b = b.data_ptr()
        for i, v in enumerate(self.p):
            d.p[i] = v
        if self.callback is not None:
            cb = _ffi.RHS_CALLBACK(self.callback)
            keep.append(cb)
            d.callback = cb
        return d


@dataclass
class SolveInfo:
    """Counters of one solve (the reference keeps ``nfe`` only in comments, neural_dynamics.py:15)."""

    nfe: int = 0
    n_accepted: int = 0
    n_rejected: int = 0
    n_launches: int = 0
    first_step: float = float("nan")
    last_dt: float = float("nan")
    t_final: float = float("nan")
    status: int = 0
    class_ms: Tuple[float, ...] = ()       # with time_kernels=True: device ms per kernel class (_ffi.K_*)
    class_launches: Tuple[int, ...] = ()


last_solve_info: Optional[SolveInfo] = None

_PREPARED: list = []  # [(weight tensor, version, prepared buffer)], most recent last


def clear_prepared() -> None:
    """Forget the derived weight forms (call after mutating a weight through ``.data``, which bumps no version)."""
    _PREPARED.clear()


def prepared_weights(W_param: torch.Tensor, W32: torch.Tensor) -> torch.Tensor:
    """Device buffer with the derived forms of an ``nn.Linear`` weight the kernels consume (``ndcn_prepare_weights_f32``),
    cached per (tensor object, in-place version): an optimiser step bumps the version, so a training iteration
    prepares once and its dozens of RHS evaluations / vjps reuse the buffer.  The cache keeps the weight tensor
    alive, so a recycled ``data_ptr`` can never alias an entry."""
    for i in range(len(_PREPARED) - 1, -1, -1):
        w, ver, buf = _PREPARED[i]
        if w is W_param:
            if ver == W_param._version:
                return buf
            del _PREPARED[i]
            break
    if len(_PREPARED) >= 8:
        del _PREPARED[0]
    lib = _ffi.lib()
    H = int(W32.shape[0])
    nbytes = int(lib.ndcn_prepared_weights_bytes(H))
    raw = torch.empty(nbytes + 1024, dtype=torch.uint8, device=W32.device)
    off = (-raw.data_ptr()) % 1024
    buf = raw[off:off + nbytes]
    with torch.cuda.device(W32.device):
        _ffi.check(lib.ndcn_prepare_weights_f32(W32.data_ptr(), H, buf.data_ptr(), current_stream_ptr(W32.device)),
                   "ndcn_prepare_weights_f32")
    _PREPARED.append((W_param, W_param._version, buf))
    return buf


def weight_grads(gp: torch.Tensor, z: torch.Tensor, dW: torch.Tensor, db: Optional[torch.Tensor],
                 accumulate: bool = True) -> None:
    """dW (+)= gp^T z, db (+)= column sums of gp on the library's own reduction kernels (``ndcn_weight_grads_f32``):
    the parameter gradients of ODEFunc's Linear (neural_dynamics.py:33) from what the RHS vjp leaves behind."""
    assert gp.is_cuda and gp.dtype == torch.float32 and gp.is_contiguous() and z.is_contiguous() and gp.shape == z.shape
    n, H = gp.shape
    assert dW.shape == (H, H) and dW.is_contiguous() and dW.dtype == torch.float32
    with torch.cuda.device(gp.device):
        rc = _ffi.lib().ndcn_weight_grads_f32(gp.data_ptr(), z.data_ptr(), n, H, dW.data_ptr(),
                                              db.data_ptr() if db is not None else None, 1 if accumulate else 0,
                                              current_stream_ptr(gp.device))
    _ffi.check(rc, "ndcn_weight_grads_f32")

_WORKSPACES: Dict[Tuple[str, int], torch.Tensor] = {}
# One solve at a time per process: the workspace, the solver-handle cache and ``last_solve_info`` are module state,
# and a handle is "one handle, one stream at a time" (include/ndcn_b200.h).  Solves from several Python threads
# (or streams) serialise here instead of sharing state buffers; ranks-as-threads tests hold their own workspaces
# (``peers=``) and are exempt.
_SOLVE_LOCK = threading.RLock()


def _workspace(device: torch.device, nbytes: int, stream: int = 0) -> torch.Tensor:
    """One reusable byte buffer per (device, stream), grown on demand (training loops call odeint
    thousands of times with the same shapes): solves enqueued on different streams never share state buffers."""
    key = (str(device), int(stream))
    buf = _WORKSPACES.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = None
        _WORKSPACES.pop(key, None)
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = buf
    return buf


_SOLVERS: Dict[tuple, tuple] = {}


def release_workspaces() -> None:
    _PREPARED.clear()
    for handle, _g in _SOLVERS.values():
        _ffi.lib().ndcn_solver_destroy(handle)
    _SOLVERS.clear()
    _WORKSPACES.clear()


def current_stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def rhs_eval(graph: CsrGraph, spec: RhsSpec, x: torch.Tensor, cache_weights: bool = False) -> torch.Tensor:
    """One evaluation f(x) on the GPU (ODEFunc.forward / *Dynamics.forward).

    ``cache_weights``: reuse the derived weight forms across calls while the weight's autograd version is unchanged
    (the training path, where weights change through optimiser steps only); off by default because a mutation
    through ``.data`` bumps no version."""
    assert x.is_cuda and x.dtype == torch.float32 and x.dim() == 2
    x = x.contiguous()
    if x.shape[0] != graph.n_cols or x.shape[1] != spec.H:
        raise ValueError("state of shape %s does not match graph with %d nodes / width %d" %
                         (tuple(x.shape), graph.n_cols, spec.H))
    out = torch.empty((graph.n_rows, spec.H), dtype=torch.float32, device=x.device)
    keep: list = []
    desc = spec.to_c(keep, prepare=cache_weights)
    with torch.cuda.device(x.device):
        rc = _ffi.lib().ndcn_rhs_eval_f32(graph.handle, C.byref(desc), x.data_ptr(), out.data_ptr(),
                                          current_stream_ptr(x.device))
    _ffi.check(rc, "ndcn_rhs_eval_f32")
    return out


def spmm(graph: CsrGraph, x: torch.Tensor) -> torch.Tensor:
    """Phi @ x (torch.sparse.mm / torch.mm at neural_dynamics.py:27-31)."""
    return rhs_eval(graph, RhsSpec.ndcn(x.shape[1], None, None, no_control=True, relu=False), x)


def odeint_fused(graph: CsrGraph, spec: RhsSpec, y0: torch.Tensor, t: torch.Tensor, *, method: str = "dopri5",
                 rtol: float = 1e-7, atol: float = 1e-9, terminal_only: bool = False,
                 forced_dt: Optional[float] = None, max_num_steps: int = 0,
                 exchange: Optional[Callable] = None, out: Optional[torch.Tensor] = None,
                 time_kernels: bool = False, first_step: Optional[float] = None, safety: float = 0.0,
                 ifactor: float = 0.0, dfactor: float = 0.0, z_block_cols: int = 0,
                 decoder: Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]] = None,
                 peers=None, small: Optional[bool] = None) -> torch.Tensor:
    """``torchdiffeq.odeint`` for a recognised RHS, entirely inside the CUDA library.

    y0: [n_rows, H] fp32 CUDA.  t: 1-D float tensor (any device); the values are used as
    given, promoted to float64 (callers that mirror ``ODEBlock`` round to fp32 first,
    neural_dynamics.py:71).  Returns ``[len(t), n_rows, H]`` (or ``[n_rows, H]`` if
    ``terminal_only``), and leaves counters in ``ndcn_b200.solver.last_solve_info``.
    ``decoder=(W_d, b_d)``: apply ``Linear(H -> C)`` (NDCN.output_layer) to every returned state inside
    the solve; the result is ``[len(t), n_rows, C]`` and the ``[len(t), n_rows, H]`` slab is never written.
    ``exchange`` / ``z_block_cols``: multi-GPU hooks of ``ndcn_b200.partition`` (halo exchange of a
    1-D row partition, or the feature-sharded gather).  ``peers``: a ``partition.PushPartition`` or
    ``partition.FeaturePushPartition`` -- the peer-push schemes (no hook: the solve runs in the partition's IPC-shared workspace and the library's
    kernels store new gather-source rows straight into the other ranks' buffers); ``graph`` must be
    ``peers.graph`` and every rank must make the same calls in the same order.
    ``small``: None = ``ndcn_odeint_f32`` (takes the persistent whole-solve kernel by itself when the problem is small
    enough), True = require that kernel (``ndcn_odeint_small_f32``: ONE cooperative launch for the whole solve),
    False = always one launch per stage (``ndcn_odeint_staged_f32``).
    """
    global last_solve_info
    if method not in _ffi.METHODS:
        raise ValueError("fused path covers %s, got %r" % (sorted(_ffi.METHODS), method))
    dev = require_cuda(y0.device)
    assert y0.is_cuda and y0.dtype == torch.float32 and y0.dim() == 2
    if y0.shape[0] != graph.n_rows or y0.shape[1] != spec.H:
        raise ValueError("y0 of shape %s does not match graph with %d rows / width %d" %
                         (tuple(y0.shape), graph.n_rows, spec.H))
    y0 = y0.contiguous()
    if y0.data_ptr() % 16:
        y0 = y0.clone()
    t64 = t.detach().to("cpu", torch.float64).contiguous()
    n_t = int(t64.numel())
    if n_t < 1:
        raise ValueError("t must hold at least one time")
    if n_t > 1 and not bool((t64[1:] > t64[:-1]).all()):
        # misc.py:59-60
        raise AssertionError("t must be strictly increasing or decrasing")
    lib = _ffi.lib()
    method_id = _ffi.METHODS[method]
    keep: list = []
    desc = spec.to_c(keep)
    last = spec.H
    dec_W = dec_b = None
    if decoder is not None:
        # fused NDCN.output_layer: Linear(H -> C) applied to every returned state, C <= 8
        dec_W = decoder[0].detach().to(dev, torch.float32).contiguous()
        dec_b = decoder[1].detach().to(dev, torch.float32).contiguous() if decoder[1] is not None else None
        if dec_W.dim() != 2 or dec_W.shape[1] != spec.H or not (1 <= dec_W.shape[0] <= 8):
            raise ValueError("decoder weight must be [C, %d] with 1 <= C <= 8, got %s" % (spec.H, tuple(dec_W.shape)))
        last = int(dec_W.shape[0])
        keep.extend([dec_W, dec_b])
    shape = (graph.n_rows, last) if terminal_only else (n_t, graph.n_rows, last)
    if out is None:
        out = torch.empty(shape, dtype=torch.float32, device=dev)
    else:
        assert out.shape == shape and out.is_cuda and out.is_contiguous() and out.dtype == torch.float32
    opts = _ffi.SolveOpts()
    opts.method = method_id
    opts.flags = (_ffi.O_TERMINAL_ONLY if terminal_only else 0) | (_ffi.O_FORCED_DT if forced_dt is not None else 0) \
        | (_ffi.O_TIME_KERNELS if time_kernels else 0)
    opts.rtol, opts.atol = float(rtol), float(atol)
    opts.forced_dt = float(forced_dt) if forced_dt is not None else 0.0
    opts.max_num_steps = int(max_num_steps)
    opts.first_step = float(first_step) if first_step is not None else 0.0
    opts.safety, opts.ifactor, opts.dfactor = float(safety), float(ifactor), float(dfactor)
    if exchange is not None:
        cb = _ffi.EXCHANGE_CALLBACK(exchange)
        keep.append(cb)
        opts.exchange = cb
    if dec_W is not None:
        opts.dec_W, opts.dec_classes = dec_W.data_ptr(), last
        opts.dec_b = dec_b.data_ptr() if dec_b is not None else None
    if z_block_cols:
        # feature-sharded multi-GPU gather (partition.FeaturePartition): the exchange hook produces Phi x
        assert exchange is not None
        opts.gather_mode, opts.z_block_cols = _ffi.GATHER_EXTERNAL, int(z_block_cols)
    stats = _ffi.SolveStats()
    # ranks that run as threads of one process (peers=) meet in device barriers: they must NOT serialise here
    with (_SOLVE_LOCK if peers is None else contextlib.nullcontext()):
        with torch.cuda.device(dev):
            nbytes = int(lib.ndcn_solver_workspace_bytes(graph.n_rows, graph.n_cols, spec.H, method_id))
            if peers is not None:
                if exchange is not None or z_block_cols or spec.callback is not None:
                    raise ValueError("peers= excludes exchange hooks and callback right-hand sides")
                if graph is not peers.graph or peers.H != spec.H or peers.workspace_bytes < nbytes:
                    raise ValueError("peers was built for another graph / width / method")
                ws_ptr, ws_bytes = peers.workspace_ptr, peers.workspace_bytes
            else:
                ws = _workspace(dev, nbytes, current_stream_ptr(dev))
                ws_ptr, ws_bytes = ws.data_ptr(), ws.numel()
            # solver handles own pinned/ctrl scratch (cudaMallocHost is slow): keep a few alive,
            # keyed on everything the handle captured by pointer
            key = (id(graph), spec.kind, spec.flags, spec.H, method_id, desc.W, desc.b, tuple(spec.p),
                   ws_ptr, spec.callback is not None)
            entry = _SOLVERS.get(key)
            if entry is None or spec.callback is not None:
                handle = C.c_void_p()
                _ffi.check(lib.ndcn_solver_create(graph.handle, C.byref(desc), method_id, ws_ptr, ws_bytes,
                                                  C.byref(handle)), "ndcn_solver_create")
                if peers is not None:
                    peers.configure(handle)
                if len(_SOLVERS) >= 16:
                    oldest = next(iter(_SOLVERS))  # FIFO: never the handle another in-flight rank just created
                    old, _g = _SOLVERS.pop(oldest)
                    lib.ndcn_solver_destroy(old)
                if spec.callback is None:
                    _SOLVERS[key] = (handle, graph)
            else:
                handle = entry[0]
            try:
                t_ptr = C.cast(t64.data_ptr(), _ffi.c_double_p)
                entry = lib.ndcn_odeint_f32 if small is None else (lib.ndcn_odeint_small_f32 if small else
                                                                   lib.ndcn_odeint_staged_f32)
                rc = entry(handle, y0.data_ptr(), t_ptr, n_t, out.data_ptr(), C.byref(opts), C.byref(stats),
                           current_stream_ptr(dev))
            finally:
                if spec.callback is not None:
                    lib.ndcn_solver_destroy(handle)
    if peers is not None:
        peers.n_solves += 1
    last_solve_info = SolveInfo(stats.nfe, stats.n_accepted, stats.n_rejected, stats.n_launches, stats.first_step,
                                stats.last_dt, stats.t_final, stats.status, tuple(stats.class_ms),
                                tuple(stats.class_launches))
    _ffi.check(rc, "ndcn_odeint_f32")
    return out
