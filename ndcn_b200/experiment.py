"""The dynamics experiment of the reference's driver scripts at any graph size.

``heat_dynamics.py`` / ``gene_dynamics.py`` / ``mutualistic_dynamics.py`` build the graph through networkx
into a dense ``[n, n]`` matrix (heat_dynamics.py:83-117), which ends at a few tens of thousands of nodes; run
unmodified through ``python -m ndcn_b200.run`` they keep that limit.  This module is the same experiment --
same flags, same initial value, same time sampling, same ground-truth solve, same model, loss, optimiser and
test schedule -- with every ``[n, n]`` object sparse (``workloads.network`` / ``workloads.graph_operator``), so
it reaches the BASELINE sizes (100k - 4M nodes).  The compute is the library's: the ground truth is one fused
dopri5 solve of the Heat / Gene / Mutualistic RHS, the model is ``ndcn_b200.NDCN`` (fused forward, kernels of
the discrete adjoint in the backward).

    python -m ndcn_b200.experiment --dynamics heat --network power_law --n 1000000 --mean_degree 10 \
        --sampled_time equal --baseline ndcn --method euler --hidden 20 --niters 50

Differences from the scripts, all forced by size: ``--n`` need not be a perfect square (the scripts lay the
initial value out on a ceil(sqrt(n))-sided square and need n = side**2; here the first n entries of that
square are used), ``--network random|community`` accept ``--mean_degree`` (the scripts' edge probabilities 0.1
and 0.25/0.01 mean 1e5 neighbours per node at 1M nodes), ``--layout`` additionally accepts ``rcm``/``bfs``/
``none`` and defaults to ``degree`` (``community`` runs networkx's greedy modularity: small graphs only), and
only the continuous-time baselines (ndcn, no_embed, no_control, no_graph) exist.
"""
from __future__ import annotations

import argparse
import time
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F

from . import workloads
from .dynamics import GeneDynamics, HeatDiffusion, MutualDynamics
from .models import NDCN
from .odeint import odeint


def parser() -> argparse.ArgumentParser:
    """The scripts' arguments (heat_dynamics.py:18-64) minus --viz/--dump/--adjoint, plus --dynamics and
    --mean_degree."""
    p = argparse.ArgumentParser("ndcn_b200.experiment")
    p.add_argument("--dynamics", choices=["heat", "gene", "mutualistic"], default="heat")
    p.add_argument("--method", choices=["dopri5", "euler", "midpoint", "rk4"], default="euler")
    p.add_argument("--rtol", type=float, default=0.01)
    p.add_argument("--atol", type=float, default=0.001)
    p.add_argument("--lr", type=float, default=0.01)
    p.add_argument("--weight_decay", type=float, default=1e-3)
    p.add_argument("--dropout", type=float, default=0)
    p.add_argument("--hidden", type=int, default=20)
    p.add_argument("--time_tick", type=int, default=100)
    p.add_argument("--sampled_time", choices=["irregular", "equal"], default="irregular")
    p.add_argument("--niters", type=int, default=2000)
    p.add_argument("--test_freq", type=int, default=20)
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--n", type=int, default=400)
    p.add_argument("--network", choices=["grid", "random", "power_law", "small_world", "community"], default="grid")
    p.add_argument("--layout", choices=["community", "degree", "rcm", "bfs", "none"], default="degree")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--T", type=float, default=5.0)
    p.add_argument("--operator", choices=["lap", "norm_lap", "kipf", "norm_adj"], default="norm_lap")
    p.add_argument("--baseline", choices=["ndcn", "no_embed", "no_control", "no_graph"], default="ndcn")
    p.add_argument("--mean_degree", type=float, default=None)
    return p


def initial_value(n: int) -> torch.Tensor:
    """The scripts' three constant patches on a side x side square (heat_dynamics.py:178-183), flattened
    row-major; the first n entries when n is not a perfect square."""
    side = int(np.ceil(np.sqrt(n)))
    x0 = torch.zeros(side, side)
    x0[int(0.05 * side):int(0.25 * side), int(0.05 * side):int(0.25 * side)] = 25
    x0[int(0.45 * side):int(0.75 * side), int(0.45 * side):int(0.75 * side)] = 20
    x0[int(0.05 * side):int(0.25 * side), int(0.35 * side):int(0.65 * side)] = 17
    return x0.view(-1, 1).float()[:n].contiguous()


def time_ticks(sampled_time: str, T: float, time_tick: int):
    """``t`` and the train / extrapolation / interpolation index lists (heat_dynamics.py:119-151); draws from
    numpy's global generator in the scripts' order."""
    if sampled_time == "equal":
        t = torch.linspace(0.0, T, time_tick)
        id_train = list(range(int(time_tick * 0.8)))
        id_test = list(range(int(time_tick * 0.8), time_tick))
        return t, id_train, id_test, None
    sparse_scale = 10
    t = torch.linspace(0.0, T, time_tick * sparse_scale)
    t = np.random.permutation(t)[:int(time_tick * 1.2)]
    t = torch.tensor(np.sort(t))
    t[0] = 0
    id_test = list(range(time_tick, int(time_tick * 1.2)))
    id_test2 = np.random.permutation(range(1, time_tick))[:int(time_tick * 0.2)].tolist()
    id_test2.sort()
    id_train = sorted(set(range(time_tick)) - set(id_test2))
    return t, id_train, id_test, id_test2


def build_graph(args):
    """Adjacency (scipy CSR, reordered by ``--layout`` like networkx_reorder_nodes, heat_dynamics.py:90) --
    the grid keeps its row-major ids as in the scripts."""
    a = workloads.network(args.network, args.n, args.seed, args.mean_degree)
    if args.network != "grid" and args.layout != "none":
        a, _ = workloads.reorder(a, args.layout)
    return a


def truth_dynamics(kind: str, a_t: torch.Tensor, lap_t: Optional[torch.Tensor]):
    """The module each script integrates for its ground truth (heat_dynamics.py:208, gene_dynamics.py:209,
    mutualistic_dynamics.py:236)."""
    if kind == "heat":
        return HeatDiffusion(lap_t, 1)
    if kind == "gene":
        return GeneDynamics(a_t, 1)
    return MutualDynamics(a_t)


def run(args, log=print) -> Dict[str, object]:
    """Ground truth + training loop of the scripts (heat_dynamics.py:206-344); returns the losses and timings."""
    device = torch.device("cuda:%d" % args.gpu)
    np.random.seed(args.seed)
    torch.manual_seed(args.seed)
    t0 = time.time()
    a = build_graph(args)
    n = a.shape[0]
    a_t = workloads.to_reference_coo(a).to(device)
    lap_t = workloads.to_reference_coo(workloads.graph_operator(a, "lap")).to(device) if args.dynamics == "heat" else None
    om_t = workloads.to_reference_coo(workloads.graph_operator(a, args.operator)).to(device)
    build_s = time.time() - t0
    log("graph: %s, %d nodes, %d entries, built in %.1f s" % (args.network, n, a.nnz, build_s))

    x0 = initial_value(n).to(device)
    t, id_train, id_test, id_test2 = time_ticks(args.sampled_time, args.T, args.time_tick)
    t = t.to(device)

    torch.cuda.synchronize(device)
    t1 = time.time()
    with torch.no_grad():
        solution_numerical = odeint(truth_dynamics(args.dynamics, a_t, lap_t), x0, t, method="dopri5")
    torch.cuda.synchronize(device)
    truth_s = time.time() - t1
    log("ground truth: %s, %s in %.3f s" % (args.dynamics, tuple(solution_numerical.shape), truth_s))

    true_y = solution_numerical.squeeze(-1).t()  # [n, T]
    true_y_train = true_y[:, id_train]
    true_y_test = true_y[:, id_test]
    true_y_test2 = true_y[:, id_test2] if id_test2 is not None else None
    t_train = t[id_train]

    flags = dict(no_embed=args.baseline == "no_embed", no_graph=args.baseline == "no_graph",
                 no_control=args.baseline == "no_control")
    hidden = 1 if flags["no_embed"] else args.hidden
    model = NDCN(input_size=1, hidden_size=hidden, A=om_t, num_classes=1, dropout=args.dropout,
                 rtol=args.rtol, atol=args.atol, method=args.method, **flags).to(device)
    optimizer = torch.optim.Adam(model.parameters(), lr=args.lr, weight_decay=args.weight_decay)
    criterion = F.l1_loss

    out: Dict[str, object] = {"nodes": n, "entries": int(a.nnz), "build_s": build_s, "truth_s": truth_s,
                              "train_loss": [], "test": [], "solution_numerical": solution_numerical}
    torch.cuda.synchronize(device)
    t2 = time.time()
    for itr in range(1, args.niters + 1):
        optimizer.zero_grad()
        pred_y = model(t_train, x0).squeeze(-1).t()
        loss_train = criterion(pred_y, true_y_train)
        loss_train.backward()
        optimizer.step()
        out["train_loss"].append(loss_train.detach())
        if itr % args.test_freq == 0:
            with torch.no_grad():
                pred_y = model(t, x0).squeeze(-1).t()
                loss = criterion(pred_y[:, id_test], true_y_test)
                rec = {"iter": itr, "train": float(loss_train), "test": float(loss),
                       "test_rel": float(loss / true_y_test.mean())}
                if true_y_test2 is not None:
                    rec["test2"] = float(criterion(pred_y[:, id_test2], true_y_test2))
                out["test"].append(rec)
                log("Iter {:04d}| Train Loss {:.6f} | Test Loss {:.6f}({:.6f} Relative) | Time {:.4f}".format(
                    itr, rec["train"], rec["test"], rec["test_rel"], time.time() - t2))
    torch.cuda.synchronize(device)
    out["train_s"] = time.time() - t2
    out["train_loss"] = [float(v) for v in out["train_loss"]]
    out["model"] = model
    return out


def main(argv=None) -> int:
    args = parser().parse_args(argv)
    res = run(args)
    print("Total Time Used: %.3f s (graph %.1f, ground truth %.3f, %d iterations %.3f)" % (
        res["build_s"] + res["truth_s"] + res["train_s"], res["build_s"], res["truth_s"], args.niters, res["train_s"]))
    return 0


if __name__ == "__main__":
    import sys

    sys.exit(main())
