"""Multi-GPU: 1-D node-row partition of the state and the four ways the neighbour rows travel.

The reference is single-device.  Here rank p owns the contiguous row block
[row0, row1) of the state; the GEMM, the Runge-Kutta algebra and the error norm are row-local, the gather
Phi x is not.  Host logic (index building) is numpy and covered by gloo / numpy tests on the CPU.

NCCL schemes (exchange hook of ``ndcn_solve_opts_t``, ``solver.odeint_fused(..., exchange=part.exchange)``):

``RowPartition`` -- the north star's halo exchange.  The local CSR has its columns remapped to
``[local rows | halo rows]``, the halo being the sorted set of remote rows the block references.  Before every
RHS evaluation (``what=0``) boundary rows are packed by ``ndcn_pack_rows_f32`` and moved with ONE
``all_to_all_single`` into the halo region of the gather source; the dopri5 error norm is one 2-double
all-reduce per step (``what=1``).  Right for graphs WITH locality (a grid block needs one line of nodes).

``FeaturePartition`` -- for graphs WITHOUT locality (power-law, ER) the halo of a row block is nearly every
remote row (0.86 GB per RHS and rank at N=1M, P=8), while transposing the state costs 2 (P-1)/P * 4NH/P bytes
(0.22 GB).  Every rank holds the whole graph and gathers ALL rows on its own H/P-column slice:
    rows x all columns --all_to_all--> all rows x my columns --Phi--> --all_to_all--> rows x all columns
The second all-to-all delivers one [n_local, H/P] block per peer; the tcgen05 stage kernel reads that blocked
layout directly (``ndcn_solve_opts_t::z_block_cols``, ``what=2``).

Peer-memory schemes (``solver.odeint_fused(..., peers=part)``): no hook, no NCCL call, no Python between
kernels and no per-step host synchronisation.  Every rank maps the other ranks' buffers with CUDA IPC
(NVLink), the library's own kernels store into them while they compute, and a one-block barrier kernel over
IPC-shared signal pads (``k_peer_barrier``) orders those stores before the next gather and carries the dopri5
controller's 2-double all-reduce:

``PushPartition`` -- whole rows: the local CSR takes a FULL halo (every remote row, ``n_cols`` = all nodes) and
the kernels that produce a gather source (tcgen05 stage kernels, pre-stage algebra) store each new row into
their own buffer and into the halo region of every peer -- the all-gather, fused into the producer's epilogue.
(P-1) * n_local * H * 4 bytes out per RHS and rank: what the halo exchange moves at P=2, but overlapped with
the stage kernel and without the pack pass.  Row blocks may be cost-balanced (``cost_balanced_blocks``).

``FeaturePushPartition`` -- column slices: the volume of the feature-sharded gather with the push mechanism.
Producers scatter every new row, slice by slice, into the slice buffers of all ranks; the slice gather stores
z = Phi x straight into the blocked Z of the rank that owns each row; two device barriers per RHS replace
the two all-to-alls.

``exchange_volumes`` reports the bytes each scheme moves (``bench.py --exchange auto`` decides with it; measured
on the 1M-node power-law graph: peer push at 2 GPUs, feature-sharded push from 4 GPUs on).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import numpy as np
import torch

from . import _ffi
from .graph import CsrGraph


def row_blocks(n: int, world: int) -> np.ndarray:
    """Block boundaries [world + 1]: contiguous, sizes differ by at most one."""
    base, rem = divmod(n, world)
    sizes = np.full(world, base, np.int64)
    sizes[:rem] += 1
    return np.concatenate([[0], np.cumsum(sizes)])


def cost_balanced_blocks(phi, world: int, edges_per_row: float = 12.5, align: int = 128) -> np.ndarray:
    """Block boundaries [world + 1] that equalise the per-rank cost of one RHS evaluation instead of the
    row count.  The gather costs time per stored entry of Phi, the GEMM / stage algebra / push per row;
    with ``edges_per_row`` = (time per row) / (time per entry) a row weighs ``deg + edges_per_row``
    (measured on B200 at H=256: 0.16 ns per gathered entry, 2.0 ns per row -> 12.5).  On a power-law graph
    in generation order the first rows are the hubs: an even split at P=2 leaves 70 % of the entries on
    rank 0.  Cuts are rounded to ``align`` rows (whole tcgen05 tiles) where that keeps every block non-empty."""
    phi = phi.tocsr()
    n = phi.shape[0]
    if world <= 1:
        return np.array([0, n], np.int64)
    cost = np.diff(phi.indptr).astype(np.float64) + float(edges_per_row)
    cum = np.cumsum(cost)
    cuts = []
    for r in range(1, world):
        c = int(np.searchsorted(cum, cum[-1] * r / world)) + 1
        if align > 1 and n >= 4 * align * world:
            c = int(round(c / align)) * align
        lo = (cuts[-1] if cuts else 0) + 1
        cuts.append(min(max(c, lo), n - (world - r)))
    return np.array([0] + cuts + [n], np.int64)


@dataclass
class LocalBlock:
    """Host-side description of one rank's share (pure numpy; no device, no process group)."""

    rank: int
    world: int
    bounds: np.ndarray            # [world + 1]
    rowptr: np.ndarray            # int32 [n_local + 1]
    col: np.ndarray               # int32, remapped to [local | halo]
    val: np.ndarray               # fp32
    halo_global: np.ndarray       # int64 [n_halo] global ids of the halo rows, sorted (=> grouped by owner)
    recv_counts: np.ndarray       # int64 [world] halo rows received from each owner
    need_from: List[np.ndarray]   # per owner q: LOCAL row ids (in q's block) this rank needs

    @property
    def n_local(self) -> int:
        return int(self.bounds[self.rank + 1] - self.bounds[self.rank])

    @property
    def n_halo(self) -> int:
        return int(len(self.halo_global))


def uniform_blocks(n: int, world: int) -> np.ndarray:
    """Block boundaries [world + 1] with ceil(n / world) rows in every block but the last: the owner of a row is
    ``row // block`` (the feature-sharded push computes it per store instead of walking a table)."""
    nl = -(-n // world)
    b = np.minimum(np.arange(world + 1, dtype=np.int64) * nl, n)
    if not bool((np.diff(b) > 0).all()):
        return row_blocks(n, world)  # tiny graphs: fall back to sizes differing by one
    return b


def halo_row_offset(bounds: np.ndarray, src_rank: int, dst_rank: int) -> int:
    """Full-halo layout: the row of ``dst_rank``'s gather-source buffer at which ``src_rank``'s block
    starts.  A buffer is [own rows | all remote rows in global order], so the blocks of lower ranks
    follow the own block directly and the blocks of higher ranks sit at their global offset."""
    assert src_rank != dst_rank
    n_local_dst = int(bounds[dst_rank + 1] - bounds[dst_rank])
    row0_src = int(bounds[src_rank])
    return row0_src + (n_local_dst if src_rank < dst_rank else 0)


def build_local_block(phi, world: int, rank: int, full_halo: bool = False,
                      bounds: Optional[np.ndarray] = None) -> LocalBlock:
    """Slice rows [row0,row1) of the scipy CSR operator and remap its columns.  ``full_halo``: the halo
    is EVERY remote row, referenced or not (peer-push scheme: remote blocks arrive whole).  ``bounds``:
    block boundaries [world + 1] (default: even row counts, ``row_blocks``)."""
    phi = phi.tocsr()
    n = phi.shape[0]
    bounds = row_blocks(n, world) if bounds is None else np.asarray(bounds, np.int64)
    assert len(bounds) == world + 1 and bounds[0] == 0 and bounds[-1] == n and bool((np.diff(bounds) > 0).all())
    r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    blk = phi[r0:r1].tocsr()
    blk.sort_indices()
    col = blk.indices.astype(np.int64)
    local = (col >= r0) & (col < r1)
    if full_halo:
        halo_global = np.concatenate([np.arange(0, r0, dtype=np.int64), np.arange(r1, n, dtype=np.int64)])
    else:
        halo_global = np.unique(col[~local])
    new_col = np.empty_like(col)
    new_col[local] = col[local] - r0
    new_col[~local] = (r1 - r0) + np.searchsorted(halo_global, col[~local])
    owner = np.searchsorted(bounds, halo_global, side="right") - 1
    recv_counts = np.bincount(owner, minlength=world).astype(np.int64)
    need_from = [halo_global[owner == q] - bounds[q] for q in range(world)]
    return LocalBlock(rank, world, bounds, blk.indptr.astype(np.int32), new_col.astype(np.int32),
                      blk.data.astype(np.float32), halo_global, recv_counts, need_from)


class RowPartition:
    """One rank's device-side share + the exchange hook handed to ``odeint_fused``."""

    def __init__(self, block: LocalBlock, device: torch.device, H: int, group=None):
        import torch.distributed as dist

        self.block = block
        self.device = device
        self.H = int(H)
        self.group = group
        self.rank, self.world = block.rank, block.world
        self.row0, self.row1 = int(block.bounds[self.rank]), int(block.bounds[self.rank + 1])
        self.n_local, self.n_halo = block.n_local, block.n_halo
        on_gpu = device.type == "cuda"
        self.graph: Optional[CsrGraph] = None
        if on_gpu:
            self.graph = CsrGraph(torch.from_numpy(block.rowptr).to(device), torch.from_numpy(block.col).to(device),
                                  torch.from_numpy(block.val).to(device), self.n_local, self.n_local + self.n_halo)
        # who needs which of MY rows: exchange the need lists once (host side, variable length)
        counts = self._comm(torch.from_numpy(block.recv_counts.copy()))
        send_counts = torch.empty_like(counts)
        dist.all_to_all_single(send_counts, counts, group=group)
        self.recv_counts = [int(c) for c in block.recv_counts]
        self.send_counts = [int(c) for c in send_counts.cpu().tolist()]
        need = np.concatenate(block.need_from).astype(np.int64) if self.n_halo else np.zeros(0, np.int64)
        need = self._comm(torch.from_numpy(need))
        send_idx = self._comm(torch.empty(sum(self.send_counts), dtype=torch.int64))
        dist.all_to_all_single(send_idx, need, output_split_sizes=self.send_counts,
                               input_split_sizes=self.recv_counts, group=group)
        send_idx = send_idx.cpu()
        assert send_idx.numel() == 0 or (int(send_idx.min()) >= 0 and int(send_idx.max()) < self.n_local)
        self.send_idx = send_idx.to(torch.int32).to(device)
        self.n_send = int(send_idx.numel())
        self.send_buf = torch.empty((max(self.n_send, 1), self.H), dtype=torch.float32, device=device)
        self.n_exchanges = 0
        self.bytes_sent = 0

    # communication tensors must live where the backend works: CUDA for nccl, CPU for gloo
    def _comm(self, t: torch.Tensor) -> torch.Tensor:
        return t.to(self.device) if self.device.type == "cuda" else t

    @classmethod
    def build(cls, phi, world: int, rank: int, device: torch.device, H: int, group=None,
              bounds: Optional[np.ndarray] = None) -> "RowPartition":
        return cls(build_local_block(phi, world, rank, bounds=bounds), device, H, group)

    # ------------------------------------------------------------------------------------
    def fill_halo(self, buf: torch.Tensor) -> None:
        """buf: [n_local + n_halo, H]; rows >= n_local are (re)filled from their owners."""
        import torch.distributed as dist

        if self.world == 1:
            return
        if buf.is_cuda:
            _ffi.check(_ffi.lib().ndcn_pack_rows_f32(buf.data_ptr(), self.send_idx.data_ptr(), self.n_send, self.H,
                                                     self.send_buf.data_ptr(),
                                                     torch.cuda.current_stream(buf.device).cuda_stream),
                       "ndcn_pack_rows_f32")
            send = self.send_buf[:self.n_send]
        else:  # CPU tensors: gloo tests of the host logic only
            send = buf[self.send_idx.long()].contiguous()
        halo = buf[self.n_local:self.n_local + self.n_halo]
        dist.all_to_all_single(halo, send, output_split_sizes=self.recv_counts, input_split_sizes=self.send_counts,
                               group=self.group)
        self.n_exchanges += 1
        self.bytes_sent += self.n_send * self.H * 4

    def exchange(self, user, what: int, buf_ptr: int) -> int:
        """``ndcn_exchange_callback_t``: called by the solve driver on the launching thread, between
        kernels, with the current stream = the solver's stream."""
        import torch.distributed as dist

        try:
            if what == 0:
                rows = self.n_local + self.n_halo
                buf = _tensor_from_ptr(buf_ptr, (rows, self.H), torch.float32, self.device)
                self.fill_halo(buf)
            else:
                red = _tensor_from_ptr(buf_ptr, (2,), torch.float64, self.device)
                dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.group)
            return 0
        except Exception as exc:  # pragma: no cover - surfaced as a status code through the C ABI
            import traceback
            traceback.print_exc()
            print("ndcn_b200.partition.exchange failed:", exc)
            return _ffi.E_ARG

    def describe(self) -> dict:
        return {"scheme": "halo exchange", "rows_local": self.n_local, "halo_rows": self.n_halo, "send_rows": self.n_send,
                "halo_bytes_per_rhs": self.n_halo * self.H * 4, "exchanges": self.n_exchanges}


class FeaturePartition:
    """Row-sharded state + feature-sharded gather (see the module docstring).  ``graph`` is the local
    handle the solver works on (n_local rows, no entries: the solver never gathers itself),
    ``exchange`` the hook for ``odeint_fused(..., exchange=, z_block_cols=part.Hc)``."""

    def __init__(self, phi, world: int, rank: int, device: torch.device, H: int, group=None):
        n = phi.shape[0]
        if H % world != 0 or (H // world) % 32 != 0:
            raise ValueError("feature-sharded gather needs H / world to be a multiple of 32 (H=%d, world=%d)" % (H, world))
        self.world, self.rank, self.device, self.group = int(world), int(rank), device, group
        self.H, self.Hc, self.n = int(H), int(H // world), int(n)
        self.bounds = row_blocks(n, world)
        self.rows = [int(self.bounds[q + 1] - self.bounds[q]) for q in range(world)]
        self.row0, self.row1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.n_local, self.n_halo = self.rows[rank], 0
        on_gpu = device.type == "cuda"
        self.phi = phi.tocsr()
        self.full_graph: Optional[CsrGraph] = CsrGraph.from_scipy(self.phi, device) if on_gpu else None
        self.graph: Optional[CsrGraph] = None
        if on_gpu:
            self.graph = CsrGraph(torch.zeros(self.n_local + 1, dtype=torch.int32, device=device),
                                  torch.zeros(0, dtype=torch.int32, device=device),
                                  torch.zeros(0, dtype=torch.float32, device=device), self.n_local, self.n_local)
        else:
            import scipy.sparse as sp  # noqa: F401  (CPU branch: gloo tests of the exchange logic only)
            c = self.phi.tocoo()
            self._phi_cpu = torch.sparse_coo_tensor(torch.from_numpy(np.vstack((c.row, c.col)).astype(np.int64)),
                                                    torch.from_numpy(c.data.astype(np.float32)), c.shape).coalesce()
        self.send = torch.empty((world * self.n_local, self.Hc), dtype=torch.float32, device=device)
        self.x_cs = torch.empty((n, self.Hc), dtype=torch.float32, device=device)
        self.z_cs = torch.empty((n, self.Hc), dtype=torch.float32, device=device)
        self.n_exchanges = 0

    def gather(self, src: torch.Tensor, z_blocked: torch.Tensor) -> None:
        """src [n_local, H] (this rank's rows) -> z_blocked [world * n_local, Hc]: block q holds
        columns [q Hc, (q+1) Hc) of (Phi x)[row0:row1]."""
        import torch.distributed as dist

        P, nl, Hc = self.world, self.n_local, self.Hc
        if src.is_cuda:
            st = torch.cuda.current_stream(src.device).cuda_stream
            _ffi.check(_ffi.lib().ndcn_pack_cols_f32(src.data_ptr(), nl, self.H, Hc, self.send.data_ptr(), st),
                       "ndcn_pack_cols_f32")
        else:
            self.send.view(P, nl, Hc).copy_(src.view(nl, P, Hc).permute(1, 0, 2))
        # block q of `send` (my rows, peer q's columns) -> peer q; I receive every rank's rows of MY columns,
        # in rank order = row order: x_cs is [n, Hc] without any unpacking
        dist.all_to_all_single(self.x_cs, self.send, output_split_sizes=self.rows, input_split_sizes=[nl] * P,
                               group=self.group)
        if src.is_cuda:
            _ffi.check(_ffi.lib().ndcn_spmm_f32(self.full_graph.handle, self.x_cs.data_ptr(), self.z_cs.data_ptr(), Hc,
                                                torch.cuda.current_stream(src.device).cuda_stream), "ndcn_spmm_f32")
        else:
            self.z_cs.copy_(torch.sparse.mm(self._phi_cpu, self.x_cs))
        # rows of peer q (a contiguous block of z_cs) -> peer q; I receive one [n_local, Hc] block per peer
        dist.all_to_all_single(z_blocked, self.z_cs, output_split_sizes=[nl] * P, input_split_sizes=self.rows,
                               group=self.group)
        self.n_exchanges += 1

    def exchange(self, user, what: int, buf_ptr: int) -> int:
        """``ndcn_exchange_callback_t`` (what = 1: error-norm all-reduce; what = 2: external gather)."""
        import torch.distributed as dist

        try:
            if what == 2:
                req = C.cast(buf_ptr, C.POINTER(_ffi.GatherRequest)).contents
                src = _tensor_from_ptr(req.src_dev, (self.n_local, self.H), torch.float32, self.device)
                z = _tensor_from_ptr(req.z_dev, (self.world * self.n_local, self.Hc), torch.float32, self.device)
                self.gather(src, z)
            elif what == 1:
                red = _tensor_from_ptr(buf_ptr, (2,), torch.float64, self.device)
                dist.all_reduce(red, op=dist.ReduceOp.SUM, group=self.group)
            else:
                raise RuntimeError("halo exchange requested from a feature-sharded partition")
            return 0
        except Exception as exc:  # pragma: no cover - surfaced as a status code through the C ABI
            import traceback
            traceback.print_exc()
            print("ndcn_b200.partition.exchange failed:", exc)
            return _ffi.E_ARG

    def describe(self) -> dict:
        per_rank = 2 * (self.world - 1) * self.n_local * self.Hc * 4
        return {"scheme": "feature-sharded gather", "rows_local": self.n_local, "columns_local": self.Hc,
                "all_to_all_bytes_per_rhs": per_rank, "exchanges": self.n_exchanges}


class PushPartition:
    """Peer-push scheme (module docstring): full-halo local graph + IPC-mapped workspaces.

    ``graph`` / ``workspace_ptr`` / ``workspace_bytes`` / ``peer_config()`` are what
    ``solver.odeint_fused(..., peers=part)`` needs.  Build with ``PushPartition.build`` (one process per
    GPU, handles exchanged through the process group) or ``build_in_process`` (all ranks in one process,
    e.g. two threads on one GPU in the tests: the device addresses are valid as they are).
    """

    PAD_BYTES = 4096

    def __init__(self, block: LocalBlock, device: torch.device, H: int, method: str = "dopri5"):
        self.block = block
        self.device = device
        self.H = int(H)
        self.method = method
        self.rank, self.world = block.rank, block.world
        if self.world > 8:
            raise ValueError("peer push covers one NVSwitch domain (world <= 8), got %d" % self.world)
        self.bounds = block.bounds
        self.row0, self.row1 = int(block.bounds[self.rank]), int(block.bounds[self.rank + 1])
        self.n_local, self.n_halo = block.n_local, block.n_halo
        self.n = self.n_local + self.n_halo
        self.graph = CsrGraph(torch.from_numpy(block.rowptr).to(device), torch.from_numpy(block.col).to(device),
                              torch.from_numpy(block.val).to(device), self.n_local, self.n)
        lib = _ffi.lib()
        with torch.cuda.device(device):
            self.workspace_bytes = int(lib.ndcn_solver_workspace_bytes(self.n_local, self.n, self.H, _ffi.METHODS[method]))
            ptr = C.c_void_p()
            self._handle = C.create_string_buffer(64)
            _ffi.check(lib.ndcn_peer_alloc(self.PAD_BYTES + self.workspace_bytes, C.byref(ptr), self._handle),
                       "ndcn_peer_alloc")
        self.base_ptr = int(ptr.value)
        self.mapped: List[int] = []   # base address of every rank's allocation in THIS process
        self._opened: List[int] = []
        self.n_solves = 0

    # ---- construction ----------------------------------------------------------------------
    @classmethod
    def build(cls, phi, world: int, rank: int, device: torch.device, H: int, method: str = "dopri5",
              group=None, bounds: Optional[np.ndarray] = None) -> "PushPartition":
        """One process per GPU: exchange the 64-byte IPC handles through the process group and map the
        peers (``cudaIpcOpenMemHandle``; the peers' GPUs must be reachable over NVLink / PCIe P2P).
        ``bounds``: row-block boundaries, e.g. ``cost_balanced_blocks(phi, world)``."""
        import torch.distributed as dist

        self = cls(build_local_block(phi, world, rank, full_halo=True, bounds=bounds), device, H, method)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, bytes(self._handle.raw), group=group)
        lib = _ffi.lib()
        with torch.cuda.device(device):
            for r in range(world):
                if r == rank:
                    self.mapped.append(self.base_ptr)
                    continue
                p = C.c_void_p()
                _ffi.check(lib.ndcn_peer_open(handles[r], C.byref(p)), "ndcn_peer_open")
                self.mapped.append(int(p.value))
                self._opened.append(int(p.value))
        dist.barrier(group=group)  # everybody has mapped everybody before the first push
        return self

    @classmethod
    def build_in_process(cls, phi, world: int, devices, H: int, method: str = "dopri5",
                         bounds: Optional[np.ndarray] = None) -> List["PushPartition"]:
        """All ranks inside one process (tests): rank r lives on ``devices[r]`` (the same GPU is fine)."""
        parts = [cls(build_local_block(phi, world, r, full_halo=True, bounds=bounds), torch.device(devices[r]), H, method)
                 for r in range(world)]
        lib = _ffi.lib()
        for a in parts:
            a.mapped = [b.base_ptr for b in parts]
            with torch.cuda.device(a.device):
                for b in parts:
                    if b.device != a.device:
                        _ffi.check(lib.ndcn_peer_enable_access(b.device.index), "ndcn_peer_enable_access")
        return parts

    # ---- what the solver needs ---------------------------------------------------------------
    @property
    def workspace_ptr(self) -> int:
        return self.base_ptr + self.PAD_BYTES

    def peer_config(self) -> "_ffi.PeerConfig":
        assert len(self.mapped) == self.world, "peers not mapped yet"
        cfg = _ffi.PeerConfig()
        cfg.rank, cfg.world = self.rank, self.world
        own_ws = self.base_ptr + self.PAD_BYTES
        for r in range(self.world):
            cfg.pad[r] = self.mapped[r]
            if r != self.rank:
                rows = halo_row_offset(self.bounds, self.rank, r)
                cfg.delta_bytes[r] = (self.mapped[r] + self.PAD_BYTES - own_ws) + rows * self.H * 4
        return cfg

    def configure(self, handle) -> None:
        """Attach the peers to a freshly created solver handle (called by ``solver.odeint_fused``)."""
        cfg = self.peer_config()
        _ffi.check(_ffi.lib().ndcn_solver_set_peers(handle, C.byref(cfg)), "ndcn_solver_set_peers")

    def describe(self) -> dict:
        return {"scheme": "peer push (stage-kernel stores into IPC-mapped peer buffers + device barrier)",
                "rows_local": self.n_local, "halo_rows": self.n_halo, "nnz_local": int(len(self.block.col)),
                "row_bounds": [int(b) for b in self.bounds],
                "push_bytes_per_rhs": (self.world - 1) * self.n_local * self.H * 4, "solves": self.n_solves}

    def close(self, group=None) -> None:
        """Unmap the peers, then (after everybody has unmapped) free the own allocation."""
        lib = _ffi.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.ndcn_peer_close(p)
            self._opened = []
            if group is not False:
                try:
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized():
                        dist.barrier(group=group)
                except Exception:
                    pass
            if self.base_ptr:
                lib.ndcn_peer_free(self.base_ptr)
                self.base_ptr = 0


def _align(v: int, a: int) -> int:
    return (v + a - 1) // a * a


class FeaturePushPartition:
    """Feature-sharded gather over peer memory (module docstring; ``ndcn_solver_set_feature_peers``).

    Rank p owns rows ``[row0, row1)`` of the state (GEMM, solver algebra) and the column slice
    ``[p Hc, (p+1) Hc)``, ``Hc = H / world``, of EVERY node (gather).  Per rank one IPC allocation holds the
    signal pad, the slice buffer ``[N, Hc]`` and the blocked ``Z [world][n_local][Hc]``; the solver workspace
    is private.  Same interface towards ``solver.odeint_fused(..., peers=part)`` as ``PushPartition``.
    """

    PAD_BYTES = 4096

    def __init__(self, phi, world: int, rank: int, device: torch.device, H: int, method: str = "dopri5",
                 bounds: Optional[np.ndarray] = None):
        n = phi.shape[0]
        if world > 8 or H not in (128, 256) or H % world != 0 or (H // world) < 32 or ((H // world) & (H // world - 1)):
            raise ValueError("feature-sharded push needs H in {128, 256} and H / world a power of two >= 32 "
                             "(H=%d, world=%d)" % (H, world))
        self.world, self.rank, self.device = int(world), int(rank), device
        self.H, self.Hc, self.n, self.method = int(H), int(H // world), int(n), method
        self.bounds = uniform_blocks(n, world) if bounds is None else np.asarray(bounds, np.int64)
        assert len(self.bounds) == world + 1 and self.bounds[0] == 0 and self.bounds[-1] == n
        self.row0, self.row1 = int(self.bounds[rank]), int(self.bounds[rank + 1])
        self.n_local, self.n_halo = self.row1 - self.row0, 0
        self.full_graph = CsrGraph.from_scipy(phi.tocsr(), device)
        # the solver works on a handle with the local rows and no entries: it never gathers itself
        self.graph = CsrGraph(torch.zeros(self.n_local + 1, dtype=torch.int32, device=device),
                              torch.zeros(0, dtype=torch.int32, device=device),
                              torch.zeros(0, dtype=torch.float32, device=device), self.n_local, self.n_local)
        lib = _ffi.lib()
        self.xcs_bytes = _align(self.n * self.Hc * 4, 256)            # same on every rank: Z sits at the same offset
        self.z_bytes = _align(self.n_local * self.H * 4, 256)
        with torch.cuda.device(device):
            self.workspace_bytes = int(lib.ndcn_solver_workspace_bytes(self.n_local, self.n_local, self.H,
                                                                        _ffi.METHODS[method]))
            self._ws = torch.empty(self.workspace_bytes, dtype=torch.uint8, device=device)
            ptr = C.c_void_p()
            self._handle = C.create_string_buffer(64)
            _ffi.check(lib.ndcn_peer_alloc(self.PAD_BYTES + self.xcs_bytes + self.z_bytes, C.byref(ptr), self._handle),
                       "ndcn_peer_alloc")
        self.base_ptr = int(ptr.value)
        self.mapped: List[int] = []
        self._opened: List[int] = []
        self.n_solves = 0

    @classmethod
    def build(cls, phi, world: int, rank: int, device: torch.device, H: int, method: str = "dopri5",
              group=None) -> "FeaturePushPartition":
        import torch.distributed as dist

        self = cls(phi, world, rank, device, H, method)
        handles: List[Optional[bytes]] = [None] * world
        dist.all_gather_object(handles, bytes(self._handle.raw), group=group)
        lib = _ffi.lib()
        with torch.cuda.device(device):
            for r in range(world):
                if r == rank:
                    self.mapped.append(self.base_ptr)
                    continue
                p = C.c_void_p()
                _ffi.check(lib.ndcn_peer_open(handles[r], C.byref(p)), "ndcn_peer_open")
                self.mapped.append(int(p.value))
                self._opened.append(int(p.value))
        dist.barrier(group=group)
        return self

    @classmethod
    def build_in_process(cls, phi, world: int, devices, H: int, method: str = "dopri5",
                         bounds: Optional[np.ndarray] = None) -> List["FeaturePushPartition"]:
        parts = [cls(phi, world, r, torch.device(devices[r]), H, method, bounds) for r in range(world)]
        lib = _ffi.lib()
        for a in parts:
            a.mapped = [b.base_ptr for b in parts]
            with torch.cuda.device(a.device):
                for b in parts:
                    if b.device != a.device:
                        _ffi.check(lib.ndcn_peer_enable_access(b.device.index), "ndcn_peer_enable_access")
        return parts

    @property
    def workspace_ptr(self) -> int:
        return int(self._ws.data_ptr())

    def configure(self, handle) -> None:
        assert len(self.mapped) == self.world, "peers not mapped yet"
        cfg = _ffi.FeaturePeerConfig()
        cfg.rank, cfg.world = self.rank, self.world
        for r in range(self.world):
            cfg.pad[r] = self.mapped[r]
            cfg.xcs[r] = self.mapped[r] + self.PAD_BYTES
            cfg.z[r] = self.mapped[r] + self.PAD_BYTES + self.xcs_bytes
        for r in range(9):
            cfg.row_bounds[r] = int(self.bounds[min(r, self.world)])
        _ffi.check(_ffi.lib().ndcn_solver_set_feature_peers(handle, self.full_graph.handle, C.byref(cfg)),
                   "ndcn_solver_set_feature_peers")

    def describe(self) -> dict:
        per_rank = 2 * (self.world - 1) * self.n_local * self.Hc * 4
        return {"scheme": "feature-sharded peer push (slice scatter in the stage epilogue, z scatter in the slice "
                          "gather, 2 device barriers per RHS)",
                "rows_local": self.n_local, "columns_local": self.Hc, "push_bytes_per_rhs": per_rank,
                "solves": self.n_solves}

    def close(self, group=None) -> None:
        lib = _ffi.lib()
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self._opened:
                lib.ndcn_peer_close(p)
            self._opened = []
            if group is not False:
                try:
                    import torch.distributed as dist
                    if dist.is_available() and dist.is_initialized():
                        dist.barrier(group=group)
                except Exception:
                    pass
            if self.base_ptr:
                lib.ndcn_peer_free(self.base_ptr)
                self.base_ptr = 0


def exchange_volumes(phi, world: int, H: int) -> dict:
    """Bytes one rank receives per RHS evaluation under either scheme (rank 0's block as the sample)."""
    blk = build_local_block(phi, world, 0)
    return {"halo": int(blk.n_halo) * H * 4,
            "push": (world - 1) * blk.n_local * H * 4,
            "feature": 2 * (world - 1) * blk.n_local * (H // world) * 4 if H % (32 * world) == 0 else None}


class _DevArray:
    """``__cuda_array_interface__`` shim so torch can wrap a raw device pointer without a copy."""

    def __init__(self, ptr: int, shape, typestr: str):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def _tensor_from_ptr(ptr: int, shape, dtype: torch.dtype, device: torch.device) -> torch.Tensor:
    typestr = {torch.float32: "<f4", torch.float64: "<f8"}[dtype]
    return torch.as_tensor(_DevArray(ptr, shape, typestr), device=device)
