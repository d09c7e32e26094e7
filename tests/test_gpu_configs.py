"""GPU: the BASELINE.json configurations at their full sizes against the CPU oracle.

  config 3   power-law 100 489 nodes (317^2), H=256, RK4 = 3/8 rule  (mutualistic_dynamics.py NDCN, rk_common.py:72-78)
  config 4   Erdos-Renyi 1M nodes (mean degree 10), H=256, dopri5: one RHS + one forced dopri5 step
             (neural_dynamics.py:20-39, dopri5.py:94-122)
  ground truth  Heat / Gene / Mutualistic RHS on a [1M, 1] state (heat_dynamics.py:186-204,
             gene_dynamics.py:186-205 incl. :202, mutualistic_dynamics.py:206-216) + a short adaptive solve
Tolerances: the north star's rtol=1e-4 with atol as written per test (fp32, sums of ~11 products per entry; the
3xTF32 Linear agrees with fp32 FMA to ~5e-6 abs at |k| <= 2).
"""
import numpy as np
import pytest
import torch

from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _weights(H, scale=0.5):
    torch.manual_seed(0)
    lin = torch.nn.Linear(H, H)
    return (lin.weight.detach() * scale).contiguous(), lin.bias.detach().contiguous()


def _state(n, H, positive=False):
    g = torch.Generator().manual_seed(0)
    x = torch.empty((n, H)).normal_(generator=g)
    return x.abs_() if positive else x


def test_config3_powerlaw_100489_rk4_three_steps_vs_oracle():
    import ndcn_b200 as nb
    from ndcn_b200 import solver, workloads as wl

    n, H = 317 * 317, 256
    a = wl.power_law_adjacency(n, 5, seed=0)
    phi = wl.graph_operator(a, "norm_lap")
    W, b = _weights(H)
    x0 = _state(n, H)
    t = torch.linspace(0, 0.15, 4)
    g = nb.CsrGraph.from_scipy(phi, torch.device("cuda"))
    spec = nb.RhsSpec.ndcn(H, W.cuda(), b.cuda())
    y = nb.odeint_fused(g, spec, x0.cuda(), t, method="rk4").cpu()
    info = solver.last_solve_info
    assert info.nfe == 12 and info.n_accepted == 3
    coo = wl.to_reference_coo(phi)
    with torch.no_grad():
        ref = O.odeint(lambda tt, xx: O.rhs_ndcn(coo, W, b, xx), x0, t, method="rk4")
    torch.testing.assert_close(y, ref, rtol=RTOL, atol=1e-5)
    # the rows of the 20 largest hubs (degree ~ 5 sqrt(N)): the long-row path of the gather
    deg = np.diff(phi.indptr)
    hubs = torch.from_numpy(np.argsort(-deg)[:20].copy())
    torch.testing.assert_close(y[-1][hubs], ref[-1][hubs], rtol=RTOL, atol=1e-5)


def test_config4_er_1m_rhs_and_forced_dopri5_step_vs_oracle():
    import ndcn_b200 as nb
    from ndcn_b200 import solver, workloads as wl

    n, H = 1_000_000, 256
    a = wl.erdos_renyi_adjacency(n, 10.0, seed=0)
    phi = wl.graph_operator(a, "norm_lap")
    W, b = _weights(H)
    x0 = _state(n, H)
    g = nb.CsrGraph.from_scipy(phi, torch.device("cuda"))
    spec = nb.RhsSpec.ndcn(H, W.cuda(), b.cuda())
    coo = wl.to_reference_coo(phi)
    torch.set_num_threads(max(1, torch.get_num_threads()))
    with torch.no_grad():
        f_gpu = nb.rhs_eval(g, spec, x0.cuda()).cpu()
        f_ref = O.rhs_ndcn(coo, W, b, x0)
    torch.testing.assert_close(f_gpu, f_ref, rtol=RTOL, atol=3e-6)
    del f_gpu, f_ref
    # one forced dopri5 step: 7 RHS evaluations, the six stage combinations, y1 through the FSAL shortcut
    t = torch.tensor([0.0, 0.025])
    y_gpu = nb.odeint_fused(g, spec, x0.cuda(), t, method="dopri5", forced_dt=0.05, terminal_only=True).cpu()
    info = solver.last_solve_info
    assert info.n_accepted == 1 and info.nfe == 7
    with torch.no_grad():
        y_ref = O.odeint(lambda tt, xx: O.rhs_ndcn(coo, W, b, xx), x0, t, method="dopri5", forced_dt=0.05)[-1]
    torch.testing.assert_close(y_gpu, y_ref, rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("kind", ["heat", "gene", "mutual"])
def test_ground_truth_rhs_1m_nodes_d1(kind):
    import ndcn_b200 as nb
    from ndcn_b200 import solver, workloads as wl

    n = 1_000_000
    a = wl.power_law_adjacency(n, 5, seed=0)
    x0 = _state(n, 1, positive=True) * 3.0
    dev = torch.device("cuda")
    if kind == "heat":
        lap = wl.graph_operator(a, "lap")
        op = (-lap).tocsr()
        op.sort_indices()
        spec = nb.RhsSpec.heat(1, 1.0)
        L = wl.to_reference_coo(lap)
        ref_f = lambda x: O.rhs_heat(L, x, 1)  # noqa: E731
        scale = 1e-4  # hub degrees reach thousands: dt * lambda_max(L) stays inside DP5's stability region
    elif kind == "gene":
        op = a.astype(np.float32).tocsr()
        spec = nb.RhsSpec.gene(1, 1.0, 1.0, 2.0)
        A = wl.to_reference_coo(op)
        ref_f = lambda x: O.rhs_gene(A, x, 1, 1, 2)  # noqa: E731
        scale = 0.005
    else:
        op = a.astype(np.float32).tocsr()
        spec = nb.RhsSpec.mutual(1)
        A = wl.to_reference_coo(op)
        ref_f = lambda x: O.rhs_mutual_edgewise(A, x)  # noqa: E731
        scale = 0.001
    g = nb.CsrGraph.from_scipy(op, dev)
    with torch.no_grad():
        f_gpu = nb.rhs_eval(g, spec, x0.cuda()).cpu()
        f_ref = ref_f(x0)
    # hub rows sum thousands of terms; torch's CPU kernel and the GPU kernel add them in different orders
    torch.testing.assert_close(f_gpu, f_ref, rtol=RTOL, atol=1e-4 * float(f_ref.abs().max()) * 1e-2 + 1e-5)
    # three forced dopri5 steps (same dt on both sides): the solver algebra on the [N,1] kernels
    t = torch.tensor([0.0, scale * 2.5])
    y_gpu = nb.odeint_fused(g, spec, x0.cuda(), t, method="dopri5", forced_dt=scale, terminal_only=True).cpu()
    assert solver.last_solve_info.n_accepted == 3
    with torch.no_grad():
        y_ref = O.odeint(lambda tt, xx: ref_f(xx), x0, t, method="dopri5", forced_dt=scale)[-1]
    # hub rows (degree up to ~5 sqrt(N)) carry the summation-order difference of thousands of terms through 21 RHS
    # evaluations: 5e-4 relative on those few rows, everything else sits inside 1e-4
    torch.testing.assert_close(y_gpu, y_ref, rtol=5e-4, atol=2e-4 * float(y_ref.abs().max()) * 1e-2 + 1e-5)
    deg = np.diff(op.indptr)
    ordinary = torch.from_numpy(np.flatnonzero(deg <= 64))
    torch.testing.assert_close(y_gpu[ordinary], y_ref[ordinary], rtol=RTOL, atol=1e-4 * float(y_ref.abs().max()) * 1e-2 + 1e-5)
