"""CPU, world_size 2 (and 3) over gloo: the 1-D row partition's host logic and halo exchange.
Each rank rebuilds Phi x for its row block from [local | halo] rows and the result must equal the
global product; the error-norm all-reduce hook is exercised with CPU tensors."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from ndcn_b200 import partition, workloads as wl


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, H, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        phi = wl.graph_operator(wl.power_law_adjacency(n, 3, seed=1), "norm_lap")
        x = torch.from_numpy(np.random.RandomState(0).standard_normal((n, H)).astype(np.float32))
        part = partition.RowPartition.build(phi, world, rank, torch.device("cpu"), H)
        blk = part.block
        buf = torch.zeros(part.n_local + part.n_halo, H)
        buf[:part.n_local] = x[part.row0:part.row1]
        part.fill_halo(buf)
        # halo rows arrived in the order of halo_global
        ok_halo = torch.equal(buf[part.n_local:], x[torch.from_numpy(blk.halo_global)])
        local = sp.csr_matrix((blk.val, blk.col, blk.rowptr), shape=(part.n_local, part.n_local + part.n_halo))
        mine = local @ buf.numpy()
        ref = (phi @ x.numpy())[part.row0:part.row1]
        ok_spmm = np.allclose(mine, ref, rtol=1e-5, atol=1e-6)
        red = torch.tensor([float(rank + 1), float(part.n_local)], dtype=torch.float64)
        dist.all_reduce(red)
        ok_red = red.tolist() == [world * (world + 1) / 2, float(n)]
        out_q.put((rank, ok_halo, ok_spmm, ok_red, part.n_halo, part.n_send, sum(part.send_counts)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    n, H = 257, 8
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok_halo, ok_spmm, ok_red, n_halo, n_send, tot in res:
        assert ok_halo and ok_spmm and ok_red, (rank, ok_halo, ok_spmm, ok_red)
        assert n_halo > 0 and n_send == tot
    # every halo row received somewhere was sent by someone
    assert sum(r[4] for r in res) == sum(r[5] for r in res)


def test_row_blocks_and_remap():
    b = partition.row_blocks(10, 3)
    assert b.tolist() == [0, 4, 7, 10]
    phi = wl.graph_operator(wl.grid_adjacency(6), "norm_lap")
    seen = 0
    for r in range(3):
        blk = partition.build_local_block(phi, 3, r)
        assert blk.rowptr[-1] == len(blk.col) == len(blk.val)
        assert blk.col.max() < blk.n_local + blk.n_halo
        assert np.all(np.diff(blk.halo_global) > 0)
        assert blk.recv_counts[r] == 0 and blk.recv_counts.sum() == blk.n_halo
        seen += len(blk.val)
    assert seen == phi.nnz


def test_single_rank_partition_is_the_whole_graph():
    phi = wl.graph_operator(wl.erdos_renyi_adjacency(100, 6, seed=2), "norm_lap")
    blk = partition.build_local_block(phi, 1, 0)
    assert blk.n_halo == 0 and blk.n_local == 100
    assert np.array_equal(blk.col, phi.indices)


def _feature_worker(rank, world, port, n, H, out_q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        phi = wl.graph_operator(wl.power_law_adjacency(n, 3, seed=1), "norm_lap")
        x = torch.from_numpy(np.random.RandomState(0).standard_normal((n, H)).astype(np.float32))
        part = partition.FeaturePartition(phi, world, rank, torch.device("cpu"), H)
        src = x[part.row0:part.row1].contiguous()
        zb = torch.zeros(world * part.n_local, part.Hc)
        part.gather(src, zb)
        # block q = columns [q Hc, (q+1) Hc) of (Phi x)[row0:row1]
        z = zb.view(world, part.n_local, part.Hc).permute(1, 0, 2).reshape(part.n_local, H)
        ref = torch.from_numpy((phi @ x.numpy())[part.row0:part.row1])
        ok = bool(torch.allclose(z, ref, rtol=1e-5, atol=1e-6))
        # the send buffer is the column-blocked copy of this rank's rows
        ok_send = torch.equal(part.send.view(world, part.n_local, part.Hc)[1], src[:, part.Hc:2 * part.Hc])
        out_q.put((rank, ok, ok_send, part.describe()["all_to_all_bytes_per_rhs"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 257), (2, 64), (4, 130)])
def test_feature_sharded_gather_gloo(world, n):
    """all-to-all into column slices, full-graph gather on the slice, all-to-all back: the blocked
    result must be this rank's rows of Phi x (uneven row blocks included)"""
    H = 32 * world * 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_feature_worker, args=(r, world, port, n, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, ok_send, nbytes in res:
        assert ok and ok_send, (rank, ok, ok_send)
        assert nbytes > 0


def test_exchange_volumes_prefers_feature_sharding_without_locality():
    phi = wl.graph_operator(wl.power_law_adjacency(20000, 5, seed=0), "norm_lap")
    v = partition.exchange_volumes(phi, 8, 256)
    assert v["feature"] < v["halo"]           # nearly every remote row is a halo row
    grid = wl.graph_operator(wl.grid_adjacency(140), "norm_lap")
    v = partition.exchange_volumes(grid, 8, 256)
    assert v["halo"] < v["feature"]           # a grid block needs one line of nodes from each neighbour
    assert partition.exchange_volumes(grid, 3, 256)["feature"] is None


# ---- peer-push scheme: host-side index logic (no device, no process group) -------------------
@pytest.mark.parametrize("world,n", [(2, 257), (3, 100), (4, 64), (8, 1001)])
def test_full_halo_layout_and_push_offsets(world, n):
    """Every rank 'pushes' its block into every peer's buffer at halo_row_offset(); afterwards each
    rank's buffer must be [own rows | remote rows in global order] and its local CSR applied to that
    buffer must give its rows of Phi x -- the arithmetic PushPartition.peer_config() does in bytes."""
    H = 4
    phi = wl.graph_operator(wl.power_law_adjacency(n, 3, seed=2), "norm_lap")
    x = np.random.RandomState(1).standard_normal((n, H)).astype(np.float32)
    blocks = [partition.build_local_block(phi, world, r, full_halo=True) for r in range(world)]
    bounds = blocks[0].bounds
    bufs = [np.full((n, H), np.nan, np.float32) for _ in range(world)]
    for src in range(world):
        r0, r1 = int(bounds[src]), int(bounds[src + 1])
        bufs[src][:r1 - r0] = x[r0:r1]
        for dst in range(world):
            if dst != src:
                off = partition.halo_row_offset(bounds, src, dst)
                assert np.isnan(bufs[dst][off:off + (r1 - r0)]).all()  # blocks never overlap
                bufs[dst][off:off + (r1 - r0)] = x[r0:r1]
    ref = phi @ x
    for r, blk in enumerate(blocks):
        assert blk.n_local + blk.n_halo == n
        assert not np.isnan(bufs[r]).any()
        assert np.array_equal(bufs[r][blk.n_local:], x[blk.halo_global])
        local = sp.csr_matrix((blk.val, blk.col, blk.rowptr), shape=(blk.n_local, n))
        assert np.allclose(local @ bufs[r], ref[int(bounds[r]):int(bounds[r + 1])], rtol=1e-5, atol=1e-6)
    vols = partition.exchange_volumes(phi, world, H)
    assert vols["push"] == (world - 1) * blocks[0].n_local * H * 4


@pytest.mark.parametrize("world", [2, 4, 8])
def test_cost_balanced_blocks(world):
    """Power-law graph in generation order: the hubs come first, so equal row counts leave most of the
    gather work on rank 0; the cost-balanced cuts must cover all rows, stay ordered, lower the maximum
    cost, and the full-halo push layout must hold for ragged blocks."""
    n, H = 40000, 4
    phi = wl.graph_operator(wl.power_law_adjacency(n, 5, seed=3), "norm_lap")
    epr = 12.5
    even = partition.row_blocks(n, world)
    bal = partition.cost_balanced_blocks(phi, world, epr)
    assert bal[0] == 0 and bal[-1] == n and len(bal) == world + 1 and (np.diff(bal) > 0).all()
    assert all(int(b) % 128 == 0 for b in bal[1:-1])

    def max_cost(b):
        return max((phi.indptr[b[i + 1]] - phi.indptr[b[i]]) + epr * (b[i + 1] - b[i]) for i in range(world))

    assert max_cost(bal) < 0.95 * max_cost(even)
    assert max_cost(bal) < 1.02 * (phi.nnz + epr * n) / world
    # tiny graphs: no alignment, still a valid partition
    small = wl.graph_operator(wl.power_law_adjacency(50, 3, seed=1), "norm_lap")
    bs = partition.cost_balanced_blocks(small, world)
    assert bs[0] == 0 and bs[-1] == 50 and (np.diff(bs) > 0).all()
    # ragged full-halo layout
    x = np.random.RandomState(0).standard_normal((n, H)).astype(np.float32)
    ref = phi @ x
    for r in (0, world - 1):
        blk = partition.build_local_block(phi, world, r, full_halo=True, bounds=bal)
        buf = np.empty((n, H), np.float32)
        buf[:blk.n_local] = x[bal[r]:bal[r + 1]]
        for src in range(world):
            if src != r:
                off = partition.halo_row_offset(bal, src, r)
                buf[off:off + (bal[src + 1] - bal[src])] = x[bal[src]:bal[src + 1]]
        local = sp.csr_matrix((blk.val, blk.col, blk.rowptr), shape=(blk.n_local, n))
        assert np.allclose(local @ buf, ref[bal[r]:bal[r + 1]], rtol=1e-5, atol=1e-6)


def test_uniform_blocks():
    for n, world in [(1_000_000, 8), (4099, 8), (9001, 2), (5003, 4), (64, 8)]:
        b = partition.uniform_blocks(n, world)
        assert b[0] == 0 and b[-1] == n and len(b) == world + 1 and (np.diff(b) > 0).all()
        nl = int(b[1])
        rows = np.arange(n)
        assert np.array_equal(np.searchsorted(b, rows, side="right") - 1, np.minimum(rows // nl, world - 1))
    # too few rows for ceil-sized blocks: falls back to sizes differing by one
    assert np.array_equal(partition.uniform_blocks(9, 8), partition.row_blocks(9, 8))
