"""GPU: the peer-push multi-GPU scheme (partition.PushPartition, ndcn_solver_set_peers).

The kernels that produce a gather source also store every new row into the other ranks' buffers and a
device barrier kernel orders those stores before the next gather.  Here two (or three) "ranks" run as
threads of ONE process on ONE GPU -- the device addresses of the other ranks' workspaces are valid as
they are, no IPC needed -- so the 1-GPU box exercises the whole path: full-halo graphs, the pushes from
the tcgen05 stage kernels / the FP32-FMA kernels / the pre-stage algebra, the barrier and its 2-double
all-reduce.  Reference = the same solve on the unpartitioned graph (itself parity-tested against the
oracle in test_gpu_solver.py / test_gpu_umma.py) and, for one case, the CPU oracle.  The real
one-process-per-GPU variant (CUDA IPC, NVLink) is test_push_two_gpus_torchrun, which needs >= 2 GPUs.
"""
import os
import subprocess
import sys
import threading

import numpy as np
import pytest
import torch

from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _operator(n, seed=0):
    from ndcn_b200 import workloads as wl
    return wl.graph_operator(wl.power_law_adjacency(n, 5, seed=seed), "norm_lap")


def _solve_ranks(parts, spec_of, x0, t, **kw):
    """Run one solve per rank concurrently (a thread and a stream each); returns (results, infos)."""
    import ndcn_b200 as nb
    from ndcn_b200 import solver

    world = len(parts)
    dev = parts[0].device
    specs = [spec_of() for _ in range(world)]
    y0s = [x0[p.row0:p.row1].to(dev).contiguous() for p in parts]
    # create + cache every rank's solver handle first (cudaMalloc inside must not meet a spinning barrier)
    for p, s, y in zip(parts, specs, y0s):
        nb.odeint_fused(p.graph, s, y, t[:1], peers=p, **kw)
    torch.cuda.synchronize(dev)
    res, infos, errs = [None] * world, [None] * world, []
    lock = threading.Lock()

    def run(r):
        try:
            with torch.cuda.stream(torch.cuda.Stream(dev)):
                out = nb.odeint_fused(parts[r].graph, specs[r], y0s[r], t, peers=parts[r], **kw)
                with lock:
                    infos[r] = solver.last_solve_info  # module global: all ranks report the same counters
                res[r] = out.cpu()
        except Exception as exc:  # surfaced below
            errs.append((r, exc))

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for th in threads:
        th.start()
    for th in threads:
        th.join(timeout=120)
    assert not errs, errs
    assert all(r is not None for r in res)
    return res, infos


def _close(parts):
    from ndcn_b200 import solver
    solver.release_workspaces()  # cached solver handles point into the partitions' workspaces
    for p in parts:
        p.close(group=False)


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("method", ["dopri5", "rk4"])
def test_push_umma_family_matches_single_gpu(world, method):
    """H=128, >= 8192 rows per rank: full-row gather + tcgen05 stage kernels push the new rows."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition, solver

    n, H = 8192 * world + 37, 128
    phi = _operator(n, seed=world)
    dev = torch.device("cuda", 0)
    torch.manual_seed(3)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5).to(dev), lin.bias.detach().to(dev)
    x0 = torch.randn(n, H)
    if method == "dopri5":
        t = torch.tensor([0.0, 0.3, 0.55, 1.0], dtype=torch.float64)
        kw = dict(method="dopri5", rtol=1e-2, atol=1e-3)
    else:
        t = torch.linspace(0, 1, 6, dtype=torch.float64)
        kw = dict(method="rk4")
    g = nb.CsrGraph.from_scipy(phi, dev)
    ref = nb.odeint_fused(g, nb.RhsSpec.ndcn(H, W, b), x0.to(dev), t, **kw).cpu()
    ref_info = solver.last_solve_info
    parts = partition.PushPartition.build_in_process(phi, world, [dev] * world, H, method)
    try:
        for _ in range(2):  # twice: barrier epochs and payload slots carry over between solves
            res, infos = _solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W, b), x0, t, **kw)
            got = torch.cat(res, dim=1)
            assert got.shape == ref.shape
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
            assert torch.equal(got[0], x0)
    finally:
        _close(parts)
    assert ref_info.n_accepted > 0


def test_push_small_width_matches_oracle():
    """H=20 (the dynamics scripts' default): FP32-FMA stage kernels and k_epi_only push; adaptive dopri5
    with rejected steps must take the reference's step sequence on both ranks (the all-reduced error norm
    feeds both controllers the same bits)."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition
    from ndcn_b200 import workloads as wl

    n, H, world = 3001, 20, 2
    phi = _operator(n, seed=5)
    Phi = wl.to_reference_coo(phi)
    dev = torch.device("cuda", 0)
    torch.manual_seed(11)
    lin = torch.nn.Linear(H, H)
    W, b = lin.weight.detach(), lin.bias.detach()
    x0 = torch.randn(n, H)
    t = torch.tensor([0.0, 0.4, 1.0, 1.7], dtype=torch.float64)
    st = O.SolveStats()
    ref = O.odeint(lambda tt, xx: O.rhs_ndcn(Phi, W, b, xx), x0, t, rtol=1e-2, atol=1e-3, method="dopri5", stats=st)
    parts = partition.PushPartition.build_in_process(phi, world, [dev] * world, H, "dopri5")
    try:
        res, infos = _solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W.to(dev), b.to(dev)), x0, t,
                                  method="dopri5", rtol=1e-2, atol=1e-3)
    finally:
        _close(parts)
    got = torch.cat(res, dim=1)
    g = nb.CsrGraph.from_scipy(phi, dev)
    single = nb.odeint_fused(g, nb.RhsSpec.ndcn(H, W.to(dev), b.to(dev)), x0.to(dev), t, method="dopri5",
                             rtol=1e-2, atol=1e-3).cpu()
    torch.testing.assert_close(got, single, rtol=1e-4, atol=1e-5)
    assert infos[0].nfe == infos[1].nfe and infos[0].n_accepted == infos[1].n_accepted
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
    assert (infos[0].nfe, infos[0].n_accepted, infos[0].n_rejected) == (st.nfe, st.n_accepted, st.n_rejected)


def test_push_heat_dynamics_euler():
    """[N, 1] ground-truth dynamics (k_stage_dyn1) on three ranks, fixed grid."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition
    from ndcn_b200 import workloads as wl

    n, world = 5000, 3
    a = wl.power_law_adjacency(n, 5, seed=9)
    neg_lap = (-wl.graph_operator(a, "lap")).tocsr()
    dev = torch.device("cuda", 0)
    x0 = torch.rand(n, 1) * 25
    t = torch.linspace(0, 0.05, 11, dtype=torch.float64)
    g = nb.CsrGraph.from_scipy(neg_lap, dev)
    single = nb.odeint_fused(g, nb.RhsSpec.heat(1, 1.0), x0.to(dev), t, method="euler").cpu()
    parts = partition.PushPartition.build_in_process(neg_lap, world, [dev] * world, 1, "euler")
    try:
        res, _ = _solve_ranks(parts, lambda: nb.RhsSpec.heat(1, 1.0), x0, t, method="euler")
    finally:
        _close(parts)
    torch.testing.assert_close(torch.cat(res, dim=1), single, rtol=1e-4, atol=1e-5)


def test_push_two_gpus_torchrun():
    """One process per GPU, CUDA IPC over NVLink: tests/push_worker.py compares the 2-rank solve with
    the single-GPU solve on rank 0."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tests", "push_worker.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "PUSH_WORKER_OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
