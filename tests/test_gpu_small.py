"""GPU: the persistent whole-solve kernel (ndcn_odeint_small_f32, csrc/small_solver.cuh) -- one cooperative launch
per odeint() for the small graphs of BASELINE configs 1-2 -- against the reference goldens, the oracle and the
launch-per-stage path."""
import numpy as np
import pytest
import torch

from conftest import csr_to_coo, csr_to_dense
from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _info():
    from ndcn_b200 import solver
    return solver.last_solve_info


def _grid(golden):
    import ndcn_b200 as nb
    g = golden("ndcn_grid400")
    OM = csr_to_dense(g, "OM")
    W = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__weight"]).cuda()
    b = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__bias"]).cuda()
    h0, t = torch.from_numpy(g["h0"]).cuda(), torch.from_numpy(g["t"])
    graph = nb.CsrGraph.from_tensor(OM, torch.device("cuda"))
    return g, graph, nb.RhsSpec.ndcn(20, W, b), h0, t


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4", "dopri5"])
def test_config1_grid400_one_launch_vs_golden_and_staged(golden, method):
    """BASELINE config 1: the 400-node grid, H=20 (heat_dynamics.py:33,313-344): 99 steps in ONE launch"""
    import ndcn_b200 as nb
    g, graph, spec, h0, t = _grid(golden)
    kw = dict(method=method, rtol=.01, atol=.001)
    hv = nb.odeint_fused(graph, spec, h0, t.float(), small=True, **kw)
    i = _info()
    assert i.n_launches <= 3, i
    torch.testing.assert_close(hv[::10].cpu(), torch.from_numpy(g["hv_every10_" + method]), rtol=RTOL, atol=2e-6)
    assert [i.nfe, i.n_accepted, i.n_rejected] == g["stats_" + method].tolist()
    staged = nb.odeint_fused(graph, spec, h0, t.float(), small=False, **kw)
    assert _info().n_launches > 20
    if method != "dopri5":
        assert torch.equal(hv, staged)  # same device code per element: bit-identical
    else:
        torch.testing.assert_close(hv, staged, rtol=1e-6, atol=1e-7)
    # the default entry point takes the persistent kernel by itself on this size
    auto = nb.odeint_fused(graph, spec, h0, t.float(), **kw)
    assert _info().n_launches <= 3 and torch.equal(auto, hv)
    yT = nb.odeint_fused(graph, spec, h0, t.float(), small=True, terminal_only=True, **kw)
    assert torch.equal(yT, hv[-1])


def test_config2_cora_block_one_launch(golden):
    """BASELINE config 2 (dgnn.py:159-182, README flags): Cora, dopri5 rtol=atol=.1, terminal state only"""
    import ndcn_b200 as nb
    g = golden("cora_block")
    t = torch.linspace(0, 1.2, 16).float()
    for H in (32, 256):
        x = torch.from_numpy(np.tanh(np.random.RandomState(11).standard_normal((2708, H))).astype(np.float32)).cuda()
        for tag in ("a05", "a00"):
            graph = nb.CsrGraph.from_tensor(csr_to_coo(g, "adj_" + tag), torch.device("cuda"))
            for ctl in ("ctl", "noctl"):
                key = "%s_h%d_%s" % (tag, H, ctl)
                W, b = torch.from_numpy(g["W_" + key]).cuda(), torch.from_numpy(g["b_" + key]).cuda()
                spec = nb.RhsSpec.ndcn(H, W, b, no_control=(ctl == "noctl"))
                kw = dict(method="dopri5", rtol=.1, atol=.1, terminal_only=True)
                y = nb.odeint_fused(graph, spec, x, t, small=True, **kw)
                i = _info()
                assert i.n_launches <= 3, (key, i)
                assert [i.nfe, i.n_accepted, i.n_rejected] == g["stats_" + key].tolist(), key
                y2 = nb.odeint_fused(graph, spec, x, t, small=False, **kw)
                torch.testing.assert_close(y, y2, rtol=1e-5, atol=2e-6)
                ref = torch.from_numpy(g["yT_" + key])
                if H == 32:
                    torch.testing.assert_close(y.cpu(), ref, rtol=RTOL, atol=2e-6)
                else:  # tolerance: see tests/test_gpu_solver.py::test_cora_block_golden
                    torch.testing.assert_close(y.cpu()[::8], ref, rtol=RTOL, atol=2e-5)


@pytest.mark.parametrize("key", ["heat", "gene", "mutual"])
def test_ground_truth_solves_one_launch(golden, key):
    """heat_dynamics.py:207-209 & co: dopri5 at rtol 1e-7 / atol 1e-9 on the [400,1] state, ~60-140 steps in one launch"""
    import ndcn_b200 as nb
    g = golden("truth_" + key)
    A, L = csr_to_dense(g, "A"), csr_to_dense(g, "L")
    op = -L if key == "heat" else A
    spec = {"heat": nb.RhsSpec.heat(1, 1), "gene": nb.RhsSpec.gene(1, 1, 1, 2), "mutual": nb.RhsSpec.mutual(1)}[key]
    graph = nb.CsrGraph.from_tensor(op, torch.device("cuda"))
    x0, t = torch.from_numpy(g["x0"]).cuda(), torch.from_numpy(g["t"])
    sol = nb.odeint_fused(graph, spec, x0, t, method="dopri5", small=True)
    i = _info()
    assert i.n_launches <= 3 and i.status == 0 and abs(i.nfe - int(g["nfe"])) < 0.2 * int(g["nfe"])
    torch.testing.assert_close(sol.cpu(), torch.from_numpy(g["sol_dense"]), rtol=1e-4, atol=1e-4)
    # [N, d > 1] states take the column-per-lane row code
    x3 = torch.from_numpy(g["x_probe3"]).cuda()
    spec3 = {"heat": nb.RhsSpec.heat(3, 1), "gene": nb.RhsSpec.gene(3, 1, 1, 2), "mutual": nb.RhsSpec.mutual(3)}[key]
    tt = torch.tensor([0.0, 0.01, 0.02])
    a = nb.odeint_fused(graph, spec3, x3, tt, method="rk4", small=True)
    b = nb.odeint_fused(graph, spec3, x3, tt, method="rk4", small=False)
    assert torch.equal(a, b)


def test_decoder_fused_and_irregular_times(golden):
    """NDCN.output_layer inside the persistent kernel (no [T,N,H] slab), irregular output times (heat_dynamics.py:129-147)"""
    import ndcn_b200 as nb
    g, graph, spec, h0, _ = _grid(golden)
    rs = np.random.RandomState(5)
    t = torch.from_numpy(np.concatenate([[0.0], np.sort(rs.uniform(0, 5, 23))]).astype(np.float32))
    Wd = torch.from_numpy(rs.standard_normal((3, 20)).astype(np.float32)).cuda()
    bd = torch.from_numpy(rs.standard_normal(3).astype(np.float32)).cuda()
    for method in ("euler", "dopri5"):
        full = nb.odeint_fused(graph, spec, h0, t, method=method, rtol=.01, atol=.001, small=True)
        dec = nb.odeint_fused(graph, spec, h0, t, method=method, rtol=.01, atol=.001, small=True, decoder=(Wd, bd))
        assert dec.shape == (24, 400, 3) and _info().n_launches <= 3
        torch.testing.assert_close(dec, torch.nn.functional.linear(full, Wd, bd), rtol=1e-5, atol=1e-5)
        ref = O.odeint(lambda tt, x: O.rhs_ndcn(csr_to_dense(g, "OM"), spec.W.cpu(), spec.b.cpu(), x), h0.cpu(), t,
                       rtol=.01, atol=.001, method=method)
        torch.testing.assert_close(full.cpu(), ref, rtol=RTOL, atol=2e-6)


def test_errors_and_eligibility(golden):
    import ndcn_b200 as nb
    g, graph, spec, h0, t = _grid(golden)
    bad = h0.clone()
    bad[3, 2] = float("nan")
    # a NaN in y0 poisons the initial-step norms first: the reference trips over `t0 + dt > t0` (dopri5.py:100) ...
    with pytest.raises(AssertionError, match="underflow in dt"):
        nb.odeint_fused(graph, spec, bad, t.float(), method="dopri5", rtol=.01, atol=.001, small=True)
    # ... and over the finite-state assert (dopri5.py:102) once a step size is given
    with pytest.raises(AssertionError, match="non-finite"):
        nb.odeint_fused(graph, spec, bad, t.float(), method="dopri5", rtol=.01, atol=.001, small=True, first_step=0.01)
    with pytest.raises(AssertionError, match="max_num_steps"):
        nb.odeint_fused(graph, spec, h0, torch.tensor([0.0, 50.0]), method="dopri5", rtol=1e-6, atol=1e-8, small=True,
                        max_num_steps=3)
    # forced-dt steps and a user first_step behave as on the staged path
    a = nb.odeint_fused(graph, spec, h0, torch.tensor([0.0, 0.24]), method="dopri5", forced_dt=0.1, small=True)
    ia = _info()
    b = nb.odeint_fused(graph, spec, h0, torch.tensor([0.0, 0.24]), method="dopri5", forced_dt=0.1, small=False)
    assert ia.n_accepted == 3 and torch.equal(a, b)
    # a graph beyond the kernel's limits is refused by the explicit entry point and routed to the staged path by auto
    n = 20000
    rs = np.random.RandomState(0)
    r, c = rs.randint(0, n, 5 * n), rs.randint(0, n, 5 * n)
    k = r != c
    Phi = O.normalized_laplacian_coo(np.concatenate([r[k], c[k]]), np.concatenate([c[k], r[k]]), n)
    big = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    x = torch.randn(n, 20, device="cuda")
    with pytest.raises(ValueError):
        nb.odeint_fused(big, spec, x, torch.tensor([0.0, 0.1]), method="euler", small=True)
    nb.odeint_fused(big, spec, x, torch.tensor([0.0, 0.1]), method="euler")


@pytest.mark.parametrize("method", ["euler", "midpoint", "rk4"])
@pytest.mark.parametrize("flags", ["full", "no_graph", "no_control"])
def test_persistent_adjoint_matches_step_by_step_backward(method, flags):
    """ndcn_fixed_grid_adjoint_small_f32: the backward pass of a fixed-grid solve at H <= 32 (the dynamics scripts'
    training loop, heat_dynamics.py:317-334) as ONE cooperative launch, against the launch-per-vjp backward and
    CPU autograd through the oracle."""
    import ndcn_b200 as nb
    from ndcn_b200 import autograd_solver
    n, H = 500, 20
    rs = np.random.RandomState(3)
    r, c = rs.randint(0, n, 4 * n), rs.randint(0, n, 4 * n)
    k = r != c
    A = torch.zeros(n, n)
    A[r[k], c[k]] = 1.0
    A = A / A.sum(1, keepdim=True).clamp(min=1.0)  # non-symmetric: Phi^T is a different matrix
    kw = dict(no_graph=flags == "no_graph", no_control=flags == "no_control")
    torch.manual_seed(1)
    func = nb.ODEFunc(H, A.to_sparse(), **kw).cuda()
    x0 = torch.randn(n, H)
    t = torch.tensor([0.0, 0.2, 0.35, 0.7, 1.0])
    wts = torch.randn(5, n, H) / (n * H) ** 0.5

    def grads(persistent):
        autograd_solver.FusedFixedGridFn.persistent_backward = persistent
        try:
            func.zero_grad(set_to_none=True)
            xg = x0.clone().cuda().requires_grad_()
            out = nb.odeint(func, xg, t.cuda(), method=method)
            (out * wts.cuda()).sum().backward()
            gw = func.wt.weight.grad.clone() if flags != "no_control" else None
            gb = func.wt.bias.grad.clone() if flags != "no_control" else None
            return xg.grad.clone(), gw, gb
        finally:
            autograd_solver.FusedFixedGridFn.persistent_backward = True

    a, b = grads(True), grads(False)

    def rel(u, v):
        return float((u - v).norm() / v.norm().clamp(min=1e-30))

    assert rel(a[0], b[0]) < 1e-5
    if flags != "no_control":
        assert rel(a[1], b[1]) < 1e-5 and rel(a[2], b[2]) < 1e-5
    W = func.wt.weight.detach().cpu().clone().requires_grad_()
    bb = func.wt.bias.detach().cpu().clone().requires_grad_()
    xr = x0.clone().requires_grad_()
    (O.odeint(lambda tt, x: O.rhs_ndcn(A, W, bb, x, **kw), xr, t, method=method) * wts).sum().backward()
    assert rel(a[0].cpu(), xr.grad) < 1e-4
    if flags != "no_control":
        assert rel(a[1].cpu(), W.grad) < 1e-4 and rel(a[2].cpu(), bb.grad) < 1e-4
