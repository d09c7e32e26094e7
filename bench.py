#!/usr/bin/env python
"""bench.py -- node-state-updates/s of the NDCN ODE hot path on B200 (BASELINE.json metric).

Workload (``config.workload``): the north-star configuration -- a 1M-node power-law graph
(preferential attachment, m=5), Phi = I - D^-1/2 A D^-1/2 as fp32 CSR, hidden width 256,
``relu((Phi X) W^T + b)`` integrated by dopri5 with a forced step dt = T/100 = 0.05 (error
estimate computed every step, every step accepted: SURVEY.md section 8(d) "S dopri5 steps").
One "step" = one dopri5 step over the whole state = 6 RHS evaluations + the stage algebra +
the error norm.  ``value`` = N*H*steps / seconds with the state resident in HBM; ``e2e`` = the
same solve through the public ``ndcn_b200.odeint(ODEFunc, y0, t)`` call with y0 in pinned HOST
memory and the result returned to the host (H2D + D2H inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our CUDA path
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own CPU path: the unmodified
                                                                   torchdiffeq.odeint(neural_dynamics.ODEFunc ...) from
                                                                   baseline/_ref (oracle port when that copy is absent)
    python bench.py --config {3,4,5}                               the other BASELINE.json configurations (presets)
    python bench.py --rhs {heat,gene,mutual} --hidden 1 ...        the ground-truth right-hand sides at scale

N > 1: launched by torch.distributed.run, one rank per GPU; the state is partitioned 1-D by
node rows ("strong" scaling: the 1M-node problem is fixed) and the neighbour rows every RHS
evaluation needs travel by one of four schemes (``--exchange``, default ``auto`` = pick_exchange():
peer push at 2 GPUs, feature-sharded peer push from 4 GPUs on -- the library's kernels store into
the other ranks' IPC-mapped memory, no NCCL call on the path -- and the NCCL halo exchange for graphs
with locality).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "node-state-updates/sec (nodes x hidden x steps / s)"
UNIT = "updates/s"
T_TOTAL, STEPS_TOTAL = 5.0, 100  # north star: T=5, 100 dopri5 steps  ->  dt = 0.05
DT = T_TOTAL / STEPS_TOTAL


PUSH_MAX_WORLD = 4    # whole-row peer push: beyond this every rank would receive > 0.8 GB per RHS over NVLink
FPUSH_MIN_WORLD = 4   # feature-sharded peer push takes over here (pick_exchange has the measured table)
PUSH_IN_AUTO = True  # `--exchange auto` may pick the peer-push scheme


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--nodes", type=int, default=1_000_000)
    ap.add_argument("--hidden", type=int, default=256)
    ap.add_argument("--graph", choices=["power_law", "er", "grid"], default="power_law")
    ap.add_argument("--layout", choices=["generation", "degree"], default="generation")
    ap.add_argument("--method", choices=["dopri5", "rk4", "euler"], default="dopri5")
    ap.add_argument("--adaptive", action="store_true",
                    help="dopri5 with the NDCN tolerances (rtol .01, atol .001) over T=5 instead of forced steps; "
                         "reports the measured accepted/rejected steps (SURVEY.md section 8(d)); single GPU")
    ap.add_argument("--exchange", choices=["auto", "halo", "feature", "push", "fpush"], default="auto",
                    help="multi-GPU exchange scheme: halo rows of the row partition, feature-sharded gather, "
                         "or whichever moves fewer bytes per RHS (auto)")
    ap.add_argument("--even-rows", action="store_true",
                    help="multi-GPU push: equal row counts per rank instead of cost-balanced row blocks")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="multi-GPU: skip the comparison with the single-GPU solve")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-budget-s", type=float, default=25.0, help="CPU seconds for the cpu_baseline sample")
    ap.add_argument("--ref-budget-s", type=float, default=240.0, help="--impl reference: CPU seconds for the whole run")
    ap.add_argument("--config", type=int, default=0, choices=[0, 3, 4, 5],
                    help="BASELINE.json preset: 0 north star (1M power-law, dopri5; default), 3 power-law 100489 nodes "
                         "RK4 (3/8 rule), 4 ER 1M dopri5, 5 power-law 4M dopri5; explicit flags override the preset")
    ap.add_argument("--dt", type=float, default=None,
                    help="step size (default T/100 = 0.05; the ground-truth dynamics on a power-law graph need ~1e-4: "
                         "hub degrees in the thousands bound the stable step of an explicit method)")
    ap.add_argument("--rhs", choices=["ndcn", "heat", "gene", "mutual"], default="ndcn",
                    help="right-hand side: the NDCN ODEFunc (default) or a ground-truth dynamics (use --hidden 1)")
    args = ap.parse_args()
    preset = {0: {}, 3: dict(nodes=317 * 317, graph="power_law", method="rk4"),
              4: dict(nodes=1_000_000, graph="er", method="dopri5"),
              5: dict(nodes=4_000_000, graph="power_law", method="dopri5")}[args.config]
    given = {a.split("=")[0].lstrip("-").replace("-", "_") for a in sys.argv[1:] if a.startswith("--")}
    for k, v in preset.items():
        if k not in given:
            setattr(args, k, v)
    if args.dt is not None:
        global DT
        DT = float(args.dt)
    return args


# ----------------------------------------------------------------------------------------------
# workload
# ----------------------------------------------------------------------------------------------
def build_operator(args, n):
    """Phi of the workload: the normalized Laplacian for the NDCN ODEFunc (the scripts' default operator,
    heat_dynamics.py:160-161), -L = A - D for HeatDiffusion (heat_dynamics.py:116-117,190), the adjacency for
    GeneDynamics / MutualDynamics (gene_dynamics.py:209)."""
    from ndcn_b200 import workloads as wl

    if args.graph == "power_law":
        a = wl.power_law_adjacency(n, 5, seed=0)
    elif args.graph == "er":
        a = wl.erdos_renyi_adjacency(n, 10.0, seed=0)
    else:
        side = int(round(n ** 0.5))
        a = wl.grid_adjacency(side)
    if args.layout == "degree":
        a, _ = wl.reorder_by_degree(a)
    kind = getattr(args, "rhs", "ndcn")
    if kind == "heat":
        m = (-wl.graph_operator(a, "lap")).tocsr()
        m.sort_indices()
        return m
    if kind in ("gene", "mutual"):
        return a.astype(np.float32).tocsr()
    return wl.graph_operator(a, "norm_lap")


def make_weights(H):
    """nn.Linear(H, H) default init under seed 0, halved so that the state stays finite over T=5."""
    torch.manual_seed(0)
    lin = torch.nn.Linear(H, H)
    return (lin.weight.detach() * 0.5).contiguous(), lin.bias.detach().contiguous()


def make_state(n, H, pin, positive=False):
    g = torch.Generator().manual_seed(0)
    x = torch.empty((n, H), dtype=torch.float32, pin_memory=pin)
    x.normal_(generator=g)
    if positive:  # the ground-truth dynamics live on non-negative states (x0 of the scripts: 0 .. 25)
        x.abs_()
    return x


def bytes_rhs(n, nnz, H, rhs="ndcn"):
    """algorithmic bytes of one RHS evaluation (state in, k out, CSR once [, W, b])."""
    base = 2 * 4 * n * H + 8 * nnz + 4 * (n + 1)
    return base + (4 * H * H + 4 * H if rhs == "ndcn" else 0)


RHS_PER_STEP = {"dopri5": 6, "rk4": 4, "euler": 1}


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for ts, line in self.rows:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                mx = float(parts[1])
                if t0 <= ts <= t1 + 0.2:
                    sm.append(float(parts[0]))
                    for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"),
                                         parts[3:7]):
                        if val.lower().startswith("active"):
                            reasons.add(name)
            except ValueError:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ----------------------------------------------------------------------------------------------
# the reference's own implementation of the path, on the host cores (cpu_baseline / --impl reference) or on
# cuda:0 (gpu_baseline: cuSPARSE + cuBLAS + Python-issued elementwise kernels)
# ----------------------------------------------------------------------------------------------
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _reference_modules():
    """(neural_dynamics, torchdiffeq) of the UNMODIFIED reference copy in baseline/_ref, or None."""
    if not os.path.isfile(os.path.join(REF_DIR, "neural_dynamics.py")):
        return None
    from oracle import ref_loader

    ref_loader.REFERENCE_ROOT = REF_DIR
    return ref_loader.import_reference()


def _sample_operator(args, n):
    from ndcn_b200 import workloads as wl

    a = wl.power_law_adjacency(n, 5, seed=0) if args.graph == "power_law" else \
        (wl.erdos_renyi_adjacency(n, 10.0, seed=0) if args.graph == "er" else wl.grid_adjacency(int(round(n ** .5))))
    return wl.to_reference_coo(wl.graph_operator(a, "norm_lap"))


def reference_run(args, device, budget_s, n_fixed=None):
    """The reference algorithm on `device` for the bench workload (NDCN ODEFunc, H, graph family), on a graph
    whose size is the full one when the estimated run fits `budget_s` seconds, else the largest that does.

    With baseline/_ref present this is the reference's own code: ``torchdiffeq.odeint(neural_dynamics.ODEFunc(H,
    Phi_coo), x0, t, ...)`` under no_grad, Phi an uncoalesced fp32 COO tensor as ``utils.py:12-23`` builds it
    (kind "reference"); otherwise the oracle port (kind "port").  The reference's dopri5 has no fixed-step mode, so
    it runs adaptively with the NDCN tolerances (rtol .01 / atol .001, neural_dynamics.py:123-126) over T=5 and the
    work is counted in step equivalents of 6 RHS evaluations (nfe / 6: the initial-step probe's 2 evaluations
    count as a third of a step); rk4 / euler run exactly `steps` grid steps."""
    from oracle import ndcn_oracle as O

    is_cpu = device.type == "cpu"
    cores = os.cpu_count() or 1
    if is_cpu:
        torch.set_num_threads(cores)
    H = args.hidden
    W, b = make_weights(H)
    per_step = RHS_PER_STEP[args.method]
    mods = _reference_modules()

    def sync():
        if not is_cpu:
            torch.cuda.synchronize(device)

    def make_func(phi):
        if mods is not None:
            nd, _ = mods
            f = nd.ODEFunc(H, phi.to(device))
            with torch.no_grad():
                f.wt.weight.copy_(W)
                f.wt.bias.copy_(b)
            return f.to(device).eval()
        Wd, bd, pd = W.to(device), b.to(device), phi.to(device)
        return lambda tt, xx: O.rhs_ndcn(pd, Wd, bd, xx)

    # calibrate: seconds per (node x RHS evaluation) on a small graph of the same family
    n_cal = min(args.nodes, 32768)
    phi = _sample_operator(args, n_cal)
    f = make_func(phi)
    x = make_state(phi.shape[0], H, False).to(device)
    with torch.no_grad():
        f(None, x)
        sync()
        t0 = time.perf_counter()
        for _ in range(2):
            f(None, x)
        sync()
        per_node_eval = (time.perf_counter() - t0) / 2 / n_cal
    # a step costs ~ per_step RHS + ~as much again in solver algebra; dopri5 adaptive over
    # T=5 takes ~3.3 step equivalents + a 1.3-step warm-up solve, fixed grids 3 steps + 1 warm-up
    steps_est = 4.7 if args.method == "dopri5" else 4.0
    n = n_fixed or int(budget_s / max(per_node_eval * per_step * 2.0 * steps_est, 1e-12))
    n = max(4096, min(args.nodes, n))
    if n != n_cal:
        del f, x, phi
        phi = _sample_operator(args, n)
        f = make_func(phi)
        x = make_state(phi.shape[0], H, False).to(device)
    n = phi.shape[0]

    class Counted(torch.nn.Module):
        def __init__(self, inner):
            super().__init__()
            self.inner, self.nfe = inner, 0

        def forward(self, tt, xx):
            self.nfe += 1
            return self.inner(tt, xx)

    cf = Counted(f)
    odeint = mods[1].odeint if mods is not None else None

    def run(t, method, **kw):
        with torch.no_grad():
            if odeint is not None:
                return odeint(cf, x, t.to(device), method=method, **kw)
            return O.odeint(cf, x, t, method=method, **kw)

    if args.method == "dopri5":
        run(torch.tensor([0.0, 0.05]), "dopri5", rtol=.01, atol=.001)  # warm-up: init probe + first attempts
        sync()
        cf.nfe = 0
        t0 = time.perf_counter()
        run(torch.tensor([0.0, T_TOTAL]), "dopri5", rtol=.01, atol=.001)
        sync()
        dt = time.perf_counter() - t0
        steps_eq = cf.nfe / 6.0
        how = "adaptive dopri5 rtol=.01 atol=.001 over T=%g: nfe=%d = %.2f step equivalents of 6 RHS" % (T_TOTAL, cf.nfe, steps_eq)
    else:
        run(torch.linspace(0, DT, 2), args.method)
        sync()
        k = 3
        t0 = time.perf_counter()
        run(torch.linspace(0, DT * k, k + 1), args.method)
        sync()
        dt = time.perf_counter() - t0
        steps_eq = float(k)
        how = "%d %s grid steps of dt=%g after 1 warm-up step" % (k, args.method, DT)
    value = n * H * steps_eq / dt
    kind = "reference" if mods is not None else "port"
    what = ("unmodified reference torchdiffeq.odeint(neural_dynamics.ODEFunc) from baseline/_ref" if mods is not None
            else "oracle port of the reference algorithm (baseline/_ref absent)")
    where = ("%d host threads" % cores) if is_cpu else ("cuda:%d (cuSPARSE/cuBLAS/ATen elementwise)" % (device.index or 0))
    return {"value": value, "unit": UNIT, "cores": cores if is_cpu else 0, "kind": kind,
            "sample": "%d-node %s graph (same generator as the workload), H=%d, Phi as uncoalesced fp32 COO; %s; %s; "
                      "torch %s on %s" % (n, args.graph, H, how, what, torch.__version__, where),
            "seconds": dt, "ms_per_step": 1e3 * dt / steps_eq, "nodes": n, "steps_eq": steps_eq}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    if args.rhs != "ndcn":
        print(json.dumps({"impl": "reference", "unavailable": "reference arm covers the NDCN ODEFunc workload only"}))
        return 0
    res = reference_run(args, torch.device("cpu"), args.ref_budget_s)
    cfg = workload_config(args, res["nodes"], None)
    cfg["workload"] += "; REFERENCE ARM SAMPLE: " + res["sample"]
    cfg["requested_nodes"] = args.nodes
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "steps_timed": res["steps_eq"], "ms_per_step": res["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, n, nnz):
    how = ("forced dt=%.3g (T=5 / 100 steps), one step = 6 RHS evals + stage algebra + error norm" % DT
           if args.method == "dopri5" else
           "fixed grid dt=%.3g, one step = %d RHS evals + stage algebra" % (DT, RHS_PER_STEP[args.method]))
    rhs = {"ndcn": "relu((Phi X) W^T + b), Phi = norm. Laplacian", "heat": "k (-L) X  (HeatDiffusion)",
           "gene": "-b x^f + A (x^h/(x^h+1))  (GeneDynamics)", "mutual": "MutualDynamics (edge-wise)"}[args.rhs]
    return {
        "workload": "BASELINE config %s: %s graph %d nodes (%s order)%s, RHS %s, width=%d, %s %s"
                    % (args.config or "north star", args.graph, n, args.layout, "" if nnz is None else " nnz=%d" % nnz,
                       rhs, args.hidden, args.method, how),
        "nodes": n, "hidden": args.hidden, "method": args.method, "dt": DT, "rhs": args.rhs,
        "l2_policy": "inputs larger than L2 (state %.0f MB x >=11 buffers vs 126 MB L2), no flush" %
                     (n * args.hidden * 4 / 1e6) if n * args.hidden * 4 * 11 > (126 << 20) else
                     "state %.1f MB x 11 buffers: 256 MB L2 flush buffer written between warm-up and timed region"
                     % (n * args.hidden * 4 / 1e6),
    }


# ----------------------------------------------------------------------------------------------
# our arm
def pick_exchange(vols: dict, world: int, H: int, allow_push: bool = True) -> str:
    """`--exchange auto`, from the measurements on the 1M-node power-law graph (ms per step, profiles/README.md):

        GPUs   halo (NCCL)   feature (NCCL)   push    fpush
          2       18.7           25.0         12.5    13.8
          4       17.8           12.3         11.1     8.9
          8       14.8            7.2           -      6.0

    A graph WITH locality (halo far smaller than the remote rows, e.g. a grid) keeps the NCCL halo exchange: it
    moves kilobytes where every other scheme moves the whole state.  Otherwise the peer-memory schemes win: whole
    rows pushed from the stage kernels at 2 GPUs (same bytes as the halo exchange, but overlapped and without pack
    pass), column slices from 4 GPUs on (2 (P-1)/P^2 of the state per RHS instead of (P-1)/P).  `allow_push` =
    False (CUDA IPC unavailable) falls back to the NCCL schemes by volume."""
    feature_ok = vols.get("feature") is not None and H in (128, 256)
    fpush_ok = feature_ok and (H // world) >= 32 and ((H // world) & (H // world - 1)) == 0 and world <= 8
    halo, push = vols["halo"], vols["push"]
    if halo < 0.5 * push:
        return "feature" if (feature_ok and vols["feature"] < halo) else "halo"
    if allow_push and PUSH_IN_AUTO and world <= 8:
        if world >= FPUSH_MIN_WORLD and fpush_ok:
            return "fpush"
        if world <= PUSH_MAX_WORLD or not feature_ok:
            return "push"
    if feature_ok and vols["feature"] < halo:
        return "feature"
    return "halo"


def solve_on(nb, graph, spec, y0, k, method):
    """K forced steps on one GPU (the parity reference of the multi-GPU lines)."""
    if method == "dopri5":
        t = torch.tensor([0.0, DT * (k - 0.5)], dtype=torch.float64)
        return nb.odeint_fused(graph, spec, y0, t, method="dopri5", forced_dt=DT, terminal_only=True)
    t = torch.linspace(0, DT * k, k + 1, dtype=torch.float64)
    return nb.odeint_fused(graph, spec, y0, t, method=method, terminal_only=True)


# ----------------------------------------------------------------------------------------------
def main_ours(args):
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi, solver

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (impl=ours) needs a CUDA device: the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa
        dist.init_process_group("nccl", device_id=dev)
    _ffi.lib()  # fail loudly if the CUDA library is missing

    n, H = args.nodes, args.hidden
    phi = build_operator(args, n)
    n = phi.shape[0]
    nnz = int(phi.nnz)
    W, b = make_weights(H)
    W, b = W.to(dev), b.to(dev)
    if args.rhs == "ndcn":
        spec = nb.RhsSpec.ndcn(H, W, b)
    elif args.rhs == "heat":
        spec = nb.RhsSpec.heat(H, 1.0)
    elif args.rhs == "gene":
        spec = nb.RhsSpec.gene(H, 1.0, 1.0, 2.0)
    else:
        spec = nb.RhsSpec.mutual(H)
    if args.rhs != "ndcn" and world > 1:
        raise RuntimeError("--rhs %s: the [N,d] ground-truth dynamics run as replicas (DESIGN.md section 6), --gpus 1" % args.rhs)
    x0_host = make_state(n, H, pin=True, positive=args.rhs != "ndcn")

    z_block_cols = 0
    vols = None
    peers = None
    if world == 1:
        graph = nb.CsrGraph.from_scipy(phi, dev)
        exchange = None
        x0 = x0_host.to(dev)
        part = None
    else:
        from ndcn_b200 import partition
        # four exchange schemes (ndcn_b200/partition.py): NCCL halo exchange of a 1-D row partition, the
        # feature-sharded gather (2 NCCL all-to-alls), and the two peer-memory schemes (the library's kernels store
        # whole rows / column slices into IPC-mapped peer buffers); `auto` = pick_exchange()
        vols = partition.exchange_volumes(phi, world, H)
        scheme = args.exchange
        if scheme == "auto":
            scheme = pick_exchange(vols, world, H)
        peers = None
        if scheme in ("push", "fpush"):
            # peer push needs CUDA IPC between the ranks' processes; agree collectively whether it came up
            ok = torch.ones(1, device=dev)
            meth = args.method if not args.adaptive else "dopri5"
            try:
                if scheme == "fpush":
                    part = partition.FeaturePushPartition.build(phi, world, rank, dev, H, meth)
                else:
                    # rows per rank balanced by cost: 0.16 ns per gathered entry; per row the larger of the stage
                    # kernel's own time and the NVLink push of (P-1) KB at ~750 GB/s.  2.9 ns per row = the fit to
                    # the per-rank class times of the 2-GPU run with the first guess of 2.0 (rank 0 busy 11.3 ms with
                    # 401k rows / 6.74M entries, rank 1 12.2 ms with 599k / 4.26M: 0.134 ns per entry, 2.44 ns per row)
                    row_ns = max(2.9, 1.37 * (world - 1) * H / 256.0)
                    bounds = None if args.even_rows else partition.cost_balanced_blocks(phi, world, row_ns / 0.16)
                    part = partition.PushPartition.build(phi, world, rank, dev, H, meth, bounds=bounds)
            except Exception as exc:  # pragma: no cover - depends on the box
                print("rank %d: peer push unavailable (%s); falling back" % (rank, exc), file=sys.stderr)
                ok.zero_()
                part = None
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 1.0:
                if part is not None:
                    part.close(group=False)
                if args.exchange in ("push", "fpush"):
                    raise RuntimeError("--exchange %s: CUDA IPC peer mapping failed on some rank" % args.exchange)
                scheme = pick_exchange(vols, world, H, allow_push=False)
            else:
                peers = part
        if scheme in ("push", "fpush"):
            pass
        elif scheme == "feature":
            part = partition.FeaturePartition(phi, world, rank, dev, H)
            z_block_cols = part.Hc
        else:
            part = partition.RowPartition.build(phi, world, rank, dev, H)
        graph = part.graph
        exchange = part.exchange if peers is None else None
        x0 = x0_host[part.row0:part.row1].to(dev)

    K, Wm = args.steps, args.warmup
    method = args.method
    per_step = RHS_PER_STEP[method]

    def solve(k, y0, time_kernels=False, out=None):
        if args.adaptive:
            t = torch.tensor([0.0, T_TOTAL], dtype=torch.float64)
            return nb.odeint_fused(graph, spec, y0, t, method="dopri5", rtol=.01, atol=.001, terminal_only=True,
                                   exchange=exchange, time_kernels=time_kernels, out=out, z_block_cols=z_block_cols, peers=peers)
        if method == "dopri5":
            t = torch.tensor([0.0, DT * (k - 0.5)], dtype=torch.float64)  # inside the k-th step: exactly k steps
            return nb.odeint_fused(graph, spec, y0, t, method="dopri5", forced_dt=DT, terminal_only=True,
                                   exchange=exchange, time_kernels=time_kernels, out=out, z_block_cols=z_block_cols, peers=peers)
        t = torch.linspace(0, DT * k, k + 1, dtype=torch.float64)
        return nb.odeint_fused(graph, spec, y0, t, method=method, terminal_only=True, exchange=exchange,
                               time_kernels=time_kernels, out=out, z_block_cols=z_block_cols, peers=peers)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    out_buf = torch.empty((graph.n_rows, H), dtype=torch.float32, device=dev)
    # ---- warm-up (>= 3 steps, untimed) ----
    solve(max(Wm, 3), x0, out=out_buf)
    barrier()
    small_state = n * H * 4 * 11 <= (126 << 20)
    if small_state:
        # the solver's buffers would sit in the 126 MB L2 after the warm-up: flush it with a 256 MB write
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        flush.fill_(1)
        del flush
        barrier()

    # ---- timed region: exactly K steps, state resident in HBM ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    yT = solve(K, x0, out=out_buf)   # no per-launch instrumentation inside the region that produces `value`
    ev1.record()
    barrier()
    wall1 = time.time()
    info = solver.last_solve_info
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(wall0, wall1)
    if dist is not None:
        tmax = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        ms = float(tmax.item())
    if args.adaptive:
        K = int(info.n_accepted + info.n_rejected)  # step attempts actually computed
    else:
        assert info.n_accepted == K and info.nfe == per_step * K + (1 if method == "dopri5" else 0), info
    finite = bool(torch.isfinite(yT).all())
    y_keep = yT.clone() if world > 1 else None
    value = n * H * K / (ms * 1e-3)
    launches = int(info.n_launches)

    # ---- the same K steps once more with CUDA events around every launch (NDCN_O_TIME_KERNELS): per-class device
    # times for the roofline; this instrumented solve does not contribute to `value`
    evi0, evi1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    evi0.record()
    solve(K, x0, time_kernels=True, out=out_buf)
    evi1.record()
    barrier()
    info_t = solver.last_solve_info
    ms_instr = evi0.elapsed_time(evi1)
    info = info_t

    # One RHS evaluation = the stage kernel (GEMM + bias/ReLU + RK epilogue; with the FP32-FMA family it
    # also contains the gather) plus, on the tcgen05 path, the chunk-major gather launch that feeds it.
    stage_ms = info.class_ms[_ffi.K_STAGE]
    stage_n = info.class_launches[_ffi.K_STAGE]
    gather_ms = info.class_ms[_ffi.K_GATHER]
    gather_n = info.class_launches[_ffi.K_GATHER]
    rhs_avg_ms = (stage_ms + gather_ms) / max(stage_n, 1)
    n_rows_local = graph.n_rows
    slice_cols = z_block_cols or (part.Hc if (world > 1 and scheme == "fpush") else 0)
    nnz_local = graph.nnz if not slice_cols else nnz // world  # feature-sharded: every rank gathers all rows on 1/world of the columns
    algo_bytes = bytes_rhs(n_rows_local, nnz_local, H, args.rhs)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (rhs_avg_ms * 1e-3) / 1e9 if rhs_avg_ms > 0 else 0.0
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "stage_kernel_traffic.json")))
        if prof.get("nodes") == n and prof.get("hidden") == H and world == 1:
            traffic = prof.get("dram_bytes_per_rhs")
    except Exception:
        pass
    split = gather_n > 0
    # which gather the library's auto rule picks (csrc/ndcn_api.cu::pick_gather_cw)
    cw_cfg = int(_ffi.lib().ndcn_config_get(_ffi.CFG_GATHER_CW))
    Hg = slice_cols or H
    slab_fits = Hg > 32 and graph.n_cols * Hg * 4 > (64 << 20) and graph.n_cols * 128 <= (48 << 20)
    chunked = cw_cfg > 0 or (cw_cfg == 0 and (slab_fits or (Hg <= 64 and graph.n_cols >= 4096)))
    gather_name = ("chunk-major CSR gather (k_stage_gather_chunk)" if chunked
                   else "full-row CSR gather, one warp per row (k_stage_ndcn_row)")
    roofline = {
        "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
        "traffic": traffic,
        "kernel": (f"RHS evaluation = {gather_name} + tcgen05 3xTF32 GEMM with bias/ReLU "
                   "and RK stage epilogue (k_stage_gemm_umma); time = sum of the two launches") if split else
                  ("fused RHS stage kernel (CSR gather + W GEMM + bias/ReLU + RK stage epilogue)" if args.rhs == "ndcn"
                   else "fused ground-truth RHS stage kernel (k_stage_dyn1 / k_stage_dynv: CSR gather + pointwise "
                        "dynamics + RK stage epilogue)"),
        "algorithmic_bytes_per_launch": algo_bytes, "avg_launch_ms": rhs_avg_ms, "launches_timed": int(stage_n),
        "share_of_step": (stage_ms + gather_ms) / ms_instr if ms_instr > 0 else None,
        "instrumented_ms_per_step": ms_instr / max(K, 1),
        "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (of fallback)",
        "class_ms": {"stage": stage_ms, "gather": gather_ms, "algebra": info.class_ms[_ffi.K_ALGEBRA],
                     "control": info.class_ms[_ffi.K_CONTROL], "emit": info.class_ms[_ffi.K_EMIT],
                     "exchange": info.class_ms[_ffi.K_EXCHANGE]},
        "per_kernel": {
            "gemm_epilogue": {"avg_ms": stage_ms / max(stage_n, 1), "launches": int(stage_n)},
            "gather": {"avg_ms": gather_ms / max(gather_n, 1), "launches": int(gather_n),
                       "algorithmic_bytes": 2 * 4 * n_rows_local * H + 8 * nnz_local + 4 * (n_rows_local + 1)},
        } if split else None,
    }

    # ---- e2e: public API, host buffers, H2D + D2H inside the timed region ----
    e2e = None
    if not args.no_e2e and not args.adaptive:
        if world == 1:
            from ndcn_b200 import workloads as wl
            # the reference's own operator format at scale: uncoalesced fp32 sparse COO (utils.py:12-23);
            # ODEFunc converts it to CSR on the GPU at its first use (the warm-up call below)
            coo = wl.to_reference_coo(phi)
            if args.rhs == "ndcn":
                func = nb.ODEFunc(H, coo)
                func.wt.weight.data.copy_(W)
                func.wt.bias.data.copy_(b)
                func = func.to(dev).eval()
            elif args.rhs == "heat":
                func = nb.HeatDiffusion(-coo, 1)  # the module negates its argument (heat_dynamics.py:190)
            elif args.rhs == "gene":
                func = nb.GeneDynamics(coo, 1)
            else:
                func = nb.MutualDynamics(coo)
            t_host = torch.tensor([0.0, DT * (K - 0.5)], dtype=torch.float64)
            opts = {"forced_dt": DT} if method == "dopri5" else None
            t_arg = t_host if method == "dopri5" else torch.linspace(0, DT * K, K + 1, dtype=torch.float64)

            def e2e_call():
                with torch.no_grad():
                    return nb.odeint(func, x0_host, t_arg, method=method, options=opts, terminal_only=True)

            e2e_call()  # warm-up of the host path (pageable result buffer, allocator)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            res = e2e_call()
            torch.cuda.synchronize(dev)
            e_s = time.perf_counter() - t0
            assert res.device.type == "cpu" and res.shape == (n, H)
            e2e = {"value": n * H * K / e_s, "unit": UNIT, "h2d_bytes_per_step": n * H * 4 / K,
                   "d2h_bytes_per_step": n * H * 4 / K, "seconds": e_s,
                   "note": "one ndcn_b200.odeint(func, y0_pinned_host, t) call covering K steps; y0 H2D (%d B) and "
                           "y(T) D2H (%d B) inside the timed region, bytes amortised over K steps" % (n * H * 4, n * H * 4)}
        else:
            # multi-GPU e2e: every rank stages its row block from pinned host memory and returns it
            y_host = x0_host[part.row0:part.row1].contiguous().pin_memory()
            r_host = torch.empty_like(y_host).pin_memory()
            barrier()
            t0 = time.perf_counter()
            yd = y_host.to(dev, non_blocking=True)
            r = solve(K, yd, out=out_buf)
            r_host.copy_(r, non_blocking=True)
            barrier()
            e_s = time.perf_counter() - t0
            tt = torch.tensor([e_s], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            e_s = float(tt.item())
            e2e = {"value": n * H * K / e_s, "unit": UNIT, "h2d_bytes_per_step": n * H * 4 / K,
                   "d2h_bytes_per_step": n * H * 4 / K, "seconds": e_s,
                   "note": "per-rank row blocks staged from pinned host memory and returned to it, whole job bytes"}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(Wm, 1),
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": workload_config(args, n, nnz),
        "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline,
        "solver": {"nfe": info.nfe, "accepted": info.n_accepted, "rejected": info.n_rejected, "finite": finite},
    }
    if world > 1:
        if peers is not None and scheme == "fpush":
            line["config"]["parallelism"] = ("1-D node-row partition x%d for the state / GEMM / solver algebra, feature-sharded "
                                             "gather over peer memory: stage epilogues scatter column slices into the "
                                             "IPC-mapped slice buffers of all ranks, the slice gather stores z into the "
                                             "row owners' Z, 2 device barrier kernels per RHS eval, no NCCL on the path" % world)
        elif peers is not None:
            line["config"]["parallelism"] = ("1-D node-row partition x%d, peer push: stage kernels store new rows into "
                                             "the IPC-mapped gather sources of all peers (NVLink), device barrier "
                                             "kernel per RHS eval, no NCCL on the path" % world)
        elif z_block_cols:
            line["config"]["parallelism"] = ("1-D node-row partition x%d for the state / GEMM / solver algebra, "
                                             "feature-sharded gather: 2 NCCL all-to-alls per RHS eval" % world)
        else:
            line["config"]["parallelism"] = "1-D node-row partition x%d, NCCL halo exchange before every RHS eval" % world
        line["partition"] = part.describe()
        line["partition"]["exchange_bytes_per_rhs_and_rank"] = vols
        # NVLink side of the roofline: bytes THIS rank stores into peer memory per RHS
        # evaluation over the device time of the kernels that issue those stores (stage kernels + slice gather);
        # peak = the measured peer-copy bandwidth per direction and GPU (B200_PROFILING.md: 770 GB/s, nominal 900)
        sent = None
        if peers is not None and scheme == "fpush":
            sent = 2 * graph.n_rows * H * 4 * (world - 1) // world      # y slices out + z blocks out
        elif peers is not None:
            sent = graph.n_rows * H * 4 * (world - 1)                     # whole rows to every peer
        elif vols is not None:
            sent = vols.get("feature") if z_block_cols else vols.get("halo")
        if sent:
            busy_ms = (stage_ms + gather_ms) / max(stage_n, 1)
            line["nvlink"] = {"sent_bytes_per_rhs_and_rank": int(sent), "unit": "GB/s",
                              "achieved": sent / (busy_ms * 1e-3) / 1e9 if busy_ms > 0 else None, "peak": 770.0,
                              "frac": sent / (busy_ms * 1e-3) / 1e9 / 770.0 if busy_ms > 0 else None,
                              "over": "device time of one RHS evaluation on rank 0 (stage kernel + gather), %.3f ms" % busy_ms,
                              "peak_source": "measured peer copy per direction (B200_PROFILING.md)"}
        # per-rank device time by kernel class (ms per step, CUDA events of the timed solve): where the ranks differ
        # -- `exchange` is the time inside the barrier kernels, i.e. mostly waiting for the slowest rank
        try:
            mine = {k: round(float(v) / K, 4) for k, v in roofline["class_ms"].items()}
            mine["rows"] = int(graph.n_rows)
            by_rank = [None] * world
            dist.all_gather_object(by_rank, mine)
            line["partition"]["class_ms_per_step_by_rank"] = by_rank
        except Exception as exc:  # diagnostics only: never cost the bench line
            line["partition"]["class_ms_per_step_by_rank"] = "unavailable: %s" % (exc,)

    # ---- multi-GPU parity: the K-step result against the single-GPU solve of the same steps (rank 0 runs it on the
    # whole graph), on every 64th row of every rank's block.  A line whose result is off is not printed as a result.
    if world > 1 and not args.no_parity and not args.adaptive:
        try:
            idx = torch.arange(0, graph.n_rows, 64, device=dev)
            mine_rows = y_keep.index_select(0, idx).cpu()
            gathered = [None] * world
            dist.gather_object((int(part.row0), idx.cpu(), mine_rows), gathered if rank == 0 else None, dst=0)
            parity = None
            if rank == 0:
                g1 = nb.CsrGraph.from_scipy(phi, dev)
                y_ref = solve_on(nb, g1, spec, x0_host.to(dev), K, method)
                worst, worst_rel, n_cmp = 0.0, 0.0, 0
                for row0, ii, rows in gathered:
                    ref = y_ref.index_select(0, (ii + row0).to(dev)).cpu()
                    d = (rows - ref).abs()
                    worst = max(worst, float(d.max()))
                    worst_rel = max(worst_rel, float((d / (1e-5 + 1e-4 * ref.abs())).max()))
                    n_cmp += rows.numel()
                parity = {"reference": "single-GPU solve of the same %d forced steps on rank 0" % K,
                          "rows_compared": n_cmp // H, "max_abs_diff": worst,
                          "max_violation_of_rtol1e-4_atol1e-5": worst_rel, "ok": bool(worst_rel <= 1.0)}
                del y_ref, g1
            box = [parity]
            dist.broadcast_object_list(box, src=0)
            line["parity"] = box[0]
        except Exception as exc:
            line["parity"] = {"ok": False, "error": repr(exc)}
        if not line["parity"].get("ok", False):
            line["value"] = None
            line["invalid"] = "multi-GPU result differs from the single-GPU solve: see parity"

    if rank == 0 and world == 1 and args.rhs == "ndcn" and not (args.no_cpu_baseline and args.no_gpu_baseline):
        del x0, out_buf, yT
        solver.release_workspaces()
        torch.cuda.empty_cache()
        if not args.no_gpu_baseline:
            # the reference's own GPU path on this B200 (cuSPARSE SpMM on the uncoalesced COO operator, cuBLAS Linear,
            # ~123 Python-issued elementwise kernels per dopri5 step): the existing GPU path the fused kernels replace
            try:
                res = reference_run(args, dev, 1e9, n_fixed=n)
                line["gpu_baseline"] = {k: res[k] for k in ("value", "unit", "kind", "sample", "ms_per_step")}
            except Exception as exc:  # e.g. out of memory in torch.sparse.mm's coalesce at 4M nodes
                line["gpu_baseline"] = {"unavailable": repr(exc)[:300]}
            torch.cuda.empty_cache()
        if not args.no_cpu_baseline:
            res = reference_run(args, torch.device("cpu"), args.cpu_budget_s)
            line["cpu_baseline"] = {k: res[k] for k in ("value", "unit", "cores", "kind", "sample")}
    if rank == 0:
        print(json.dumps(line))
    if peers is not None:
        solver.release_workspaces()  # solver handles point into the IPC-shared workspace
        peers.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse_args()
    if args.impl == "reference":
        return main_reference(args)
    return main_ours(args)


if __name__ == "__main__":
    sys.exit(main())
