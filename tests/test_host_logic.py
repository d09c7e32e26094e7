"""CPU: host-side logic that needs no GPU -- argument validation with the reference's exception
types, module surface, loud failure without CUDA."""
import inspect

import pytest
import torch

import ndcn_b200 as nb
from ndcn_b200 import _ffi


def test_surface_signatures_match_reference():
    assert list(inspect.signature(nb.ODEFunc.__init__).parameters) == \
        ["self", "hidden_size", "A", "dropout", "no_graph", "no_control"]
    assert list(inspect.signature(nb.ODEBlock.__init__).parameters) == \
        ["self", "odefunc", "rtol", "atol", "method", "adjoint", "terminal"]
    assert list(inspect.signature(nb.ODEBlock2.__init__).parameters) == \
        ["self", "odefunc", "vt", "rtol", "atol", "method", "adjoint", "terminal"]
    assert list(inspect.signature(nb.NDCN.__init__).parameters) == \
        ["self", "input_size", "hidden_size", "A", "num_classes", "dropout", "no_embed", "no_graph", "no_control",
         "rtol", "atol", "method"]
    p = inspect.signature(nb.odeint).parameters
    assert list(p)[:7] == ["func", "y0", "t", "rtol", "atol", "method", "options"]
    assert p["rtol"].default == 1e-7 and p["atol"].default == 1e-9 and p["method"].default is None
    pa = inspect.signature(nb.odeint_adjoint).parameters
    assert pa["rtol"].default == 1e-6 and pa["atol"].default == 1e-12


def test_odeint_adjoint_argument_check_matches_reference():
    """adjoint.py:109-110: a plain callable is refused before anything else happens"""
    with pytest.raises(ValueError, match="func is required to be an instance of nn.Module"):
        nb.odeint_adjoint(lambda t, y: y, torch.ones(2, 2), torch.tensor([0.0, 1.0]))


def test_state_dict_keys():
    m = nb.NDCN(1, 20, torch.eye(4), 1)
    assert sorted(m.state_dict()) == sorted([
        "input_layer.0.weight", "input_layer.0.bias", "input_layer.2.weight", "input_layer.2.bias",
        "neural_dynamic_layer.odefunc.wt.weight", "neural_dynamic_layer.odefunc.wt.bias",
        "output_layer.weight", "output_layer.bias"])
    assert not any(k.endswith(".A") for k in m.state_dict())  # A is a plain attribute (neural_dynamics.py:14)


def test_odeint_argument_errors_match_reference():
    f = lambda t, y: y  # noqa: E731
    y0, t = torch.ones(2, 2), torch.tensor([0.0, 1.0])
    with pytest.raises(ValueError):  # odeint.py:65-66
        nb.odeint(f, y0, t, options={"x": 1})
    with pytest.raises(KeyError):
        nb.odeint(f, y0, t, method="nope")
    with pytest.raises(TypeError):  # misc.py:190-193
        nb.odeint(f, torch.ones(2, 2, dtype=torch.int64), t)
    with pytest.raises(TypeError):
        nb.odeint(f, y0, torch.tensor([0, 1]))
    with pytest.raises(AssertionError):  # misc.py:180
        nb.odeint(f, [y0], t)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    """without a CUDA device the hot path raises; it never computes on the CPU"""
    L = torch.eye(4)
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nb.odeint(nb.HeatDiffusion(L, 1), torch.ones(4, 1), torch.tensor([0.0, 1.0]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nb.ODEFunc(4, L)(None, torch.ones(4, 4))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nb.odeint(lambda t, y: y, torch.ones(2, 2), torch.tensor([0.0, 1.0]))
    with pytest.raises(RuntimeError, match="no CUDA device"):
        nb.CsrGraph.from_tensor(L)


def test_status_codes_map_to_reference_exceptions():
    with pytest.raises(AssertionError, match="non-finite"):
        _ffi.check(_ffi.E_NONFINITE)
    with pytest.raises(AssertionError, match="underflow in dt"):
        _ffi.check(_ffi.E_DT_UNDERFLOW)
    with pytest.raises(AssertionError, match="max_num_steps"):
        _ffi.check(_ffi.E_MAX_STEPS)
    with pytest.raises(ValueError):
        _ffi.check(_ffi.E_ARG)
    with pytest.raises(RuntimeError):
        _ffi.check(700)
    _ffi.check(0)


def test_product_never_imports_oracle():
    import os
    root = os.path.dirname(os.path.abspath(nb.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_bench_exchange_rule():
    """bench.py --exchange auto: the measured table in pick_exchange (no GPU needed)."""
    import importlib.util
    import os

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    no_locality = {2: {"halo": 511e6, "push": 512e6, "feature": 512e6},
                   4: {"halo": 760e6, "push": 768e6, "feature": 384e6},
                   8: {"halo": 856e6, "push": 896e6, "feature": 224e6}}
    assert [bench.pick_exchange(no_locality[p], p, 256) for p in (2, 4, 8)] == ["push", "fpush", "fpush"]
    # CUDA IPC unavailable: NCCL schemes by volume
    assert [bench.pick_exchange(no_locality[p], p, 256, allow_push=False) for p in (2, 4, 8)] == ["halo", "feature", "feature"]
    # a grid: the halo is one line of nodes per neighbour
    assert bench.pick_exchange({"halo": 2e6, "push": 768e6, "feature": 384e6}, 4, 256) == "halo"
    # widths without a tcgen05 path have no slice schemes
    assert bench.pick_exchange({"halo": 190e6, "push": 192e6, "feature": None}, 4, 64) == "push"
    # H / world below 32 columns: no feature-sharded push
    assert bench.pick_exchange({"halo": 428e6, "push": 448e6, "feature": None}, 8, 128) in ("push", "halo")


def test_out_of_scope_solvers_are_handed_to_the_scripts_own_torchdiffeq():
    """SURVEY.md section 2 "Other solvers": adams / fixed_adams / explicit_adams / tsit5 and tuple states of several
    tensors are not accelerated.  Without a registered package they raise; under the launcher (run.install) the
    script's own vendored torchdiffeq takes them unchanged (torchdiffeq/_impl/odeint.py:20-76)."""
    import importlib
    import os

    om = importlib.import_module("ndcn_b200.odeint")
    f = lambda t, y: -y  # noqa: E731
    y0, t = torch.ones(3, 2), torch.linspace(0, 1, 5)
    om.register_out_of_scope_solver(None)
    with pytest.raises(NotImplementedError, match="out of scope"):
        nb.odeint(f, y0, t, method="adams")
    with pytest.raises(NotImplementedError, match="tuple state of 2 tensors"):
        nb.odeint(lambda t, y: (-y[0], y[0]), (y0, y0), t)
    with pytest.raises(KeyError):
        nb.odeint(f, y0, t, method="rk45")
    ref_dir = next((d for d in ("/root/reference", os.path.join(os.path.dirname(os.path.dirname(__file__)), "baseline", "_ref"))
                    if os.path.isfile(os.path.join(d, "torchdiffeq", "__init__.py"))), None)
    if ref_dir is None:
        pytest.skip("no copy of the reference's torchdiffeq on this machine")
    assert not om.register_out_of_scope_solver(os.path.dirname(__file__))  # no torchdiffeq there
    assert om.register_out_of_scope_solver(ref_dir)
    try:
        theirs = om._OUT_OF_SCOPE_SOLVER
        for method in ("adams", "explicit_adams"):
            assert torch.equal(nb.odeint(f, y0, t, method=method), theirs.odeint(f, y0, t, method=method))
        g = lambda t, y: (-y[0], y[0])  # noqa: E731
        ours = nb.odeint(g, (y0, torch.zeros(3, 2)), t, rtol=1e-5, atol=1e-7, method="dopri5")
        want = theirs.odeint(g, (y0, torch.zeros(3, 2)), t, rtol=1e-5, atol=1e-7, method="dopri5")
        assert isinstance(ours, tuple) and all(torch.equal(a, b) for a, b in zip(ours, want))
    finally:
        om.register_out_of_scope_solver(None)
