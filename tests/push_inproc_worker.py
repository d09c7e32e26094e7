"""Worker of tests/test_gpu_push.py::test_push_ranks_in_one_process: several peer-push "ranks" as threads
of ONE process on ONE GPU (the other ranks' workspaces are plain device pointers, no IPC).

Runs with CUDA_MODULE_LOADING=EAGER: the ranks share one CUDA context here, and with lazy loading the
first launch of a kernel loads it under a context-wide lock while another rank's barrier kernel may be
spinning on exactly that launch (the documented lazy-loading hazard for kernels that wait on each other;
measured: 20 s barrier time-outs and stale halo rows).  One process per GPU -- the real configuration,
tests/push_worker.py -- has a context per rank and is not affected.

CUDA_DEVICE_MAX_CONNECTIONS=32 for the same reason: with the default of 8 hardware queues the nine streams of the
8-rank case (one per rank + the default stream) are multiplexed, and a rank's kernel queued behind ANOTHER rank's
spinning barrier kernel in the same hardware queue can never run (observed once in three runs: seven ranks report
the 20 s barrier time-out).  Because co-scheduling of several ranks' kernels on one GPU is not something the
library can guarantee, a case that ends in that time-out is repeated once on fresh partitions.
"""
import os
import sys
import threading

os.environ["CUDA_MODULE_LOADING"] = "EAGER"
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from oracle import ndcn_oracle as O  # noqa: E402  (checker only)


def _operator(n, seed=0):
    from ndcn_b200 import workloads as wl
    return wl.graph_operator(wl.power_law_adjacency(n, 5, seed=seed), "norm_lap")


def solve_ranks(parts, spec_of, x0, t, **kw):
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


def close(parts):
    from ndcn_b200 import solver
    solver.release_workspaces()  # cached solver handles point into the partitions' workspaces
    for p in parts:
        p.close(group=False)


def case_umma(world, method):
    """H=128, >= 8192 rows per rank: full-row gather + tcgen05 stage kernels push the new rows."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition

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
        kw = dict(method=method)
    g = nb.CsrGraph.from_scipy(phi, dev)
    ref = nb.odeint_fused(g, nb.RhsSpec.ndcn(H, W, b), x0.to(dev), t, **kw).cpu()
    parts = partition.PushPartition.build_in_process(phi, world, [dev] * world, H, method)
    try:
        for _ in range(2):  # twice: barrier epochs and payload slots carry over between solves
            res, infos = solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W, b), x0, t, **kw)
            got = torch.cat(res, dim=1)
            assert got.shape == ref.shape
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
            assert torch.equal(got[0], x0)
    finally:
        close(parts)


def case_feature_push(world, H, method, n, ragged=False):
    """Feature-sharded peer push: slice scatter from the stage epilogues / pre-stage algebra / y0 copy, slice
    gather with z scattered to the row owners, blocked-Z tcgen05 stage kernel; ragged last block."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition

    phi = _operator(n, seed=world + H)
    dev = torch.device("cuda", 0)
    torch.manual_seed(5)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5).to(dev), lin.bias.detach().to(dev)
    x0 = torch.randn(n, H)
    if method == "dopri5":
        t = torch.tensor([0.0, 0.3, 0.55, 1.0], dtype=torch.float64)
        kw = dict(method="dopri5", rtol=1e-2, atol=1e-3)
    else:
        t = torch.linspace(0, 1, 5, dtype=torch.float64)
        kw = dict(method=method)
    g = nb.CsrGraph.from_scipy(phi, dev)
    ref = nb.odeint_fused(g, nb.RhsSpec.ndcn(H, W, b), x0.to(dev), t, **kw).cpu()
    # default: uniform blocks (owner = row // block); ragged: sizes differing by one (owner by table walk)
    bounds = partition.row_blocks(n, world) if ragged else None
    parts = partition.FeaturePushPartition.build_in_process(phi, world, [dev] * world, H, method, bounds=bounds)
    try:
        for _ in range(2):
            res, infos = solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W, b), x0, t, **kw)
            got = torch.cat(res, dim=1)
            torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
            assert torch.equal(got[0], x0)
        terminal = solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W, b), x0, t, terminal_only=True, **kw)[0]
        torch.testing.assert_close(torch.cat(terminal, dim=0), ref[-1], rtol=1e-4, atol=1e-5)
    finally:
        close(parts)


def case_small_width():
    """H=20 (the dynamics scripts' default): FP32-FMA stage kernels and k_epi_only push; adaptive dopri5
    must take the reference's step sequence on both ranks (the all-reduced error norm feeds both
    controllers the same bits) and match the CPU oracle."""
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
        res, infos = solve_ranks(parts, lambda: nb.RhsSpec.ndcn(H, W.to(dev), b.to(dev)), x0, t,
                                 method="dopri5", rtol=1e-2, atol=1e-3)
    finally:
        close(parts)
    got = torch.cat(res, dim=1)
    torch.testing.assert_close(got, ref, rtol=1e-4, atol=1e-5)
    assert infos[0].nfe == infos[1].nfe and infos[0].n_accepted == infos[1].n_accepted
    assert (infos[0].nfe, infos[0].n_accepted, infos[0].n_rejected) == (st.nfe, st.n_accepted, st.n_rejected)


def case_heat(world, method, n):
    """[N, 1] ground-truth dynamics (k_stage_dyn1), ragged row blocks (peer offsets not 16-byte aligned)."""
    import ndcn_b200 as nb
    from ndcn_b200 import partition
    from ndcn_b200 import workloads as wl

    a = wl.power_law_adjacency(n, 5, seed=9)
    neg_lap = (-wl.graph_operator(a, "lap")).tocsr()
    dev = torch.device("cuda", 0)
    x0 = torch.rand(n, 1) * 25
    t = torch.linspace(0, 0.05, 11, dtype=torch.float64)
    g = nb.CsrGraph.from_scipy(neg_lap, dev)
    single = nb.odeint_fused(g, nb.RhsSpec.heat(1, 1.0), x0.to(dev), t, method=method).cpu()
    parts = partition.PushPartition.build_in_process(neg_lap, world, [dev] * world, 1, method)
    try:
        res, _ = solve_ranks(parts, lambda: nb.RhsSpec.heat(1, 1.0), x0, t, method=method)
    finally:
        close(parts)
    torch.testing.assert_close(torch.cat(res, dim=1), single, rtol=1e-4, atol=1e-4)


CASES = [
    ("umma dopri5 x2", lambda: case_umma(2, "dopri5")),
    ("umma dopri5 x3", lambda: case_umma(3, "dopri5")),
    ("umma rk4 x2", lambda: case_umma(2, "rk4")),
    ("umma euler x3", lambda: case_umma(3, "euler")),
    ("small width dopri5 x2 vs oracle", case_small_width),
    ("heat euler x3 ragged", lambda: case_heat(3, "euler", 5000)),
    ("heat rk4 x4", lambda: case_heat(4, "rk4", 5001)),
    ("heat midpoint x2", lambda: case_heat(2, "midpoint", 4999)),
    ("feature push dopri5 x2 H=256", lambda: case_feature_push(2, 256, "dopri5", 9001)),
    ("feature push dopri5 x4 H=128", lambda: case_feature_push(4, 128, "dopri5", 5003)),
    ("feature push rk4 x8 H=256", lambda: case_feature_push(8, 256, "rk4", 4099)),
    ("feature push euler x4 H=256", lambda: case_feature_push(4, 256, "euler", 3000)),
    ("feature push midpoint x2 H=128", lambda: case_feature_push(2, 128, "midpoint", 2500)),
    ("feature push dopri5 x4 H=256 ragged blocks", lambda: case_feature_push(4, 256, "dopri5", 4099, ragged=True)),
]


def main():
    failed = 0
    for name, fn in CASES:
        try:
            try:
                fn()
            except AssertionError as exc:
                if "did not reach the barrier" not in str(exc):
                    raise
                print("CASE repeated after a barrier time-out: %s" % name, flush=True)
                torch.cuda.synchronize()
                fn()
            print("CASE ok: %s" % name, flush=True)
        except Exception as exc:
            import traceback
            traceback.print_exc()
            print("CASE FAILED: %s: %s" % (name, str(exc)[:300]), flush=True)
            failed += 1
    print("PUSH_INPROC_%s %d/%d" % ("OK" if failed == 0 else "FAILED", len(CASES) - failed, len(CASES)), flush=True)
    return 1 if failed else 0


if __name__ == "__main__":
    sys.exit(main())
