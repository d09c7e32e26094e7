"""CPU, build container only: the oracle is bit-identical to the reference imported in-process."""
import pytest
import torch

from oracle import ndcn_oracle as O
from oracle import ref_loader

pytestmark = pytest.mark.skipif(not ref_loader.reference_available(), reason="/root/reference not present")

ARGV = ["--network", "grid", "--T", "5", "--sampled_time", "equal", "--baseline", "ndcn", "--gpu", "-1"]


def test_ndcn_bit_exact_all_methods():
    nd, ode = ref_loader.import_reference()
    g = ref_loader.run_reference_script("heat_dynamics.py", ARGV)
    OM, x0, t = g["OM"], g["x0"], g["t"][:30]
    for method in ("euler", "midpoint", "rk4", "dopri5"):
        torch.manual_seed(0)
        m = nd.NDCN(1, 20, OM, 1, rtol=.01, atol=.001, method=method)
        with torch.no_grad():
            h0 = m.input_layer(x0)
            ref = m.neural_dynamic_layer(t, h0)
            W, b = m.neural_dynamic_layer.odefunc.wt.weight, m.neural_dynamic_layer.odefunc.wt.bias
            mine = O.odeint(lambda tt, x: O.rhs_ndcn(OM, W, b, x), h0, t.type_as(h0), rtol=.01, atol=.001, method=method)
        assert torch.equal(ref, mine), method


def test_truth_heat_bit_exact():
    ref_loader.import_reference()
    g = ref_loader.run_reference_script("heat_dynamics.py", ARGV)
    mine = O.odeint(lambda tt, x: O.rhs_heat(g["L"], x, 1), g["x0"], g["t"], method="dopri5")
    assert torch.equal(mine, g["solution_numerical"])


def test_irregular_grid_and_tight_tolerance():
    """irregular sampling (heat_dynamics.py:129-147) and default rtol/atol through odeint itself."""
    nd, ode = ref_loader.import_reference()
    torch.manual_seed(3)
    A = (torch.rand(50, 50) < 0.1).float()
    A = ((A + A.t()) > 0).float()
    A.fill_diagonal_(0)
    L = torch.diag(A.sum(1)) - A
    fn = nd.ODEFunc(12, L * 0.05)
    x = torch.randn(50, 12)
    t = torch.sort(torch.rand(17))[0]
    t[0] = 0
    with torch.no_grad():
        for method, kw in (("rk4", {}), ("euler", {}), ("dopri5", {}), ("dopri5", dict(rtol=1e-3, atol=1e-4))):
            ref = ode.odeint(fn, x, t, method=method, **kw)
            mine = O.odeint(lambda tt, xx: O.rhs_ndcn(fn.A, fn.wt.weight, fn.wt.bias, xx), x, t, method=method, **kw)
            assert torch.equal(ref, mine), (method, kw)


def test_fixed_grid_step_size_option_bit_exact():
    """FixedGridODESolver's `step_size` grid and its output rule (solvers.py:39-99): every requested time reports
    the END state of the first grid step that reaches it; a `grid_constructor` always raises."""
    import pytest

    from ndcn_b200.odeint import _fixed_grid_and_picks

    nd, ode = ref_loader.import_reference()
    torch.manual_seed(5)
    A = (torch.rand(40, 40) < 0.15).float()
    A = ((A + A.t()) > 0).float()
    fn = nd.ODEFunc(8, A * 0.1)
    x = torch.randn(40, 8)
    t = torch.tensor([0.0, 0.13, 0.5, 0.51, 0.9, 1.0])
    rhs = lambda tt, xx: O.rhs_ndcn(fn.A, fn.wt.weight, fn.wt.bias, xx)  # noqa: E731
    with torch.no_grad():
        for method in ("euler", "midpoint", "rk4"):
            for h in (0.1, 0.07, 0.3):
                ref = ode.odeint(fn, x, t, method=method, options={"step_size": h})
                mine = O.odeint(rhs, x, t, method=method, step_size=h)
                assert torch.equal(ref, mine), (method, h)
                # the host logic of the product: same grid, same picks
                grid, pick = _fixed_grid_and_picks(fn, x, t, h, None)
                full = ode.odeint(fn, x, grid, method=method)
                assert torch.equal(full[pick], ref), (method, h)
        for opts in ({"grid_constructor": lambda f, y, tt: tt}, {"grid_constructor": lambda f, y, tt: tt, "step_size": 0.1}):
            with pytest.raises(ValueError):
                ode.odeint(fn, x, t, method="rk4", options=opts)
            with pytest.raises(ValueError):
                O.odeint(rhs, x, t, method="rk4", **opts)
            with pytest.raises(ValueError):
                _fixed_grid_and_picks(fn, x, t, opts.get("step_size"), opts["grid_constructor"])
