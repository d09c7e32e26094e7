"""GPU: the scripts' experiment at sizes their dense graph construction cannot reach
(ndcn_b200/experiment.py; heat_dynamics.py:83-117,206-344 and the gene / mutualistic siblings) --
ground truth against the CPU oracle's dopri5 on the same sparse operator, and a short training run.
Tolerance of the ground truth: rtol 1e-4, atol 1e-4 of a state of magnitude 25 (the bar of
test_truth_dynamics_golden)."""
import numpy as np
import pytest
import torch

from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu


def _args(*extra):
    from ndcn_b200 import experiment as ex

    return ex.parser().parse_args(list(extra))


@pytest.mark.parametrize("dynamics,network,n", [("heat", "community", 3000), ("gene", "power_law", 20000),
                                                ("mutualistic", "small_world", 10000)])
def test_ground_truth_matches_oracle_and_training_reduces_loss(dynamics, network, n):
    from ndcn_b200 import experiment as ex, workloads as wl

    args = _args("--dynamics", dynamics, "--network", network, "--n", str(n), "--sampled_time", "equal",
                 "--time_tick", "20", "--T", "1.0", "--niters", "12", "--test_freq", "6", "--method", "euler",
                 "--hidden", "20", "--mean_degree", "20")
    lines = []
    res = ex.run(args, log=lines.append)
    assert res["nodes"] == n and len(res["train_loss"]) == 12 and len(res["test"]) == 2
    assert all(np.isfinite(res["train_loss"]))
    assert min(res["train_loss"][1:]) < res["train_loss"][0]

    # the same ground-truth solve on the CPU oracle (same sparse operators, reference algorithm)
    a = ex.build_graph(args)
    a_t = wl.to_reference_coo(a).coalesce()
    x0 = ex.initial_value(n)
    t, _, _, _ = ex.time_ticks("equal", 1.0, 20)
    if dynamics == "heat":
        lap = wl.to_reference_coo(wl.graph_operator(a, "lap")).coalesce()
        f = lambda tt, xx: O.rhs_heat(lap, xx, 1)  # noqa: E731
    elif dynamics == "gene":
        f = lambda tt, xx: O.rhs_gene(a_t, xx, 1)  # noqa: E731
    else:
        f = lambda tt, xx: O.rhs_mutual_edgewise(a_t, xx)  # noqa: E731
    with torch.no_grad():
        ref = O.odeint(f, x0, t, method="dopri5")
    torch.testing.assert_close(res["solution_numerical"].cpu(), ref, rtol=1e-4, atol=1e-4)


def test_irregular_sampling_and_rk4_no_control_run():
    from ndcn_b200 import experiment as ex

    args = _args("--dynamics", "heat", "--network", "random", "--n", "5000", "--mean_degree", "10", "--time_tick", "30",
                 "--niters", "4", "--test_freq", "2", "--method", "rk4", "--baseline", "no_control", "--hidden", "16")
    res = ex.run(args, log=lambda s: None)
    assert len(res["test"]) == 2 and "test2" in res["test"][0]
    assert all(np.isfinite(res["train_loss"])) and res["solution_numerical"].shape == (36, 5000, 1)
