"""GPU: the large-state kernel family -- chunk-major gather (gather_kernels.cuh) and the tcgen05
3xTF32 GEMM + stage-epilogue kernel (umma_kernels.cuh) -- forced on at test sizes through
``ndcn_config_set`` and checked against the CPU oracle / the committed reference outputs at the
same fp32 parity bar as the FP32-FMA kernels (rtol 1e-4, atol as written)."""
import numpy as np
import pytest
import torch

from conftest import csr_to_coo
from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu
RTOL = 1e-4


@pytest.fixture
def knobs():
    from ndcn_b200 import _ffi
    prev = _ffi.configure()

    def set_(**kw):
        _ffi.configure(**kw)

    yield set_
    _ffi.configure(**prev)


def _info():
    from ndcn_b200 import solver
    return solver.last_solve_info


def _graph(n, avg_deg, seed, hub=0):
    rs = np.random.RandomState(seed)
    m = int(n * avg_deg / 2)
    r, c = rs.randint(0, n, m), rs.randint(0, n, m)
    if hub:
        r = np.concatenate([r, np.full(hub, 3, np.int64), np.full(hub // 2, n - 1, np.int64)])
        c = np.concatenate([c, rs.randint(0, n, hub), rs.randint(0, n, hub // 2)])
    k = r != c
    r, c = r[k], c[k]
    return O.normalized_laplacian_coo(np.concatenate([r, c]), np.concatenate([c, r]), n)


@pytest.mark.parametrize("H", [64, 128, 256])
@pytest.mark.parametrize("cw", [16, 32, 64])
@pytest.mark.parametrize("version", [1, 2])
def test_chunked_gather_vs_full_row_and_torch(knobs, H, cw, version):
    """regular rows accumulate in CSR order in both kernels -> bitwise equal; rows above 256
    entries (two hubs here) are reduced by a CTA of their own -> fp32 tolerance"""
    import ndcn_b200 as nb
    n = 4099 if version == 1 else 40961  # v2: more 128-row items than resident CTAs, a ragged last block
    Phi = _graph(n, 11, seed=cw + H, hub=1500)
    g = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    torch.manual_seed(H)
    x = torch.randn(n, H)
    knobs(gather_cw=-1)
    full = nb.spmm(g, x.cuda()).cpu()
    knobs(gather_cw=cw, gather_version=version)
    chunked = nb.spmm(g, x.cuda()).cpu()
    deg = np.diff(g.rowptr.cpu().numpy())
    short = torch.from_numpy(deg <= 256)
    assert int((~short).sum()) >= 2
    assert torch.equal(chunked[short], full[short])
    torch.testing.assert_close(chunked, torch.sparse.mm(Phi, x), rtol=RTOL, atol=2e-6)
    # no_control RHS through the chunked kernel (ReLU + epilogue in the gather itself)
    ref = O.rhs_ndcn(Phi, None, None, x, no_control=True)
    out = nb.rhs_eval(g, nb.RhsSpec.ndcn(H, None, None, no_control=True), x.cuda()).cpu()
    torch.testing.assert_close(out, ref, rtol=RTOL, atol=2e-6)


@pytest.mark.parametrize("H", [128, 256])
@pytest.mark.parametrize("n", [777, 128, 2049])
@pytest.mark.parametrize("flags", ["full", "no_graph"])
def test_umma_rhs_vs_oracle(knobs, H, n, flags):
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi
    Phi = _graph(n, 9, seed=n + H, hub=300)
    torch.manual_seed(H + n)
    lin = torch.nn.Linear(H, H)
    W, b = lin.weight.detach(), lin.bias.detach()
    x = torch.randn(n, H)
    kw = dict(no_graph=flags == "no_graph")
    ref = O.rhs_ndcn(Phi, W, b, x, **kw)
    g = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    spec = nb.RhsSpec.ndcn(H, W.cuda(), b.cuda(), **kw)
    for cw in (-1, 16):
        knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=cw)
        out = nb.rhs_eval(g, spec, x.cuda()).cpu()
        torch.testing.assert_close(out, ref, rtol=RTOL, atol=3e-6)
    # and the two kernel families agree with each other far inside the parity bar
    knobs(stage_impl=_ffi.IMPL_SIMT, gather_cw=-1)
    simt = nb.rhs_eval(g, spec, x.cuda()).cpu()
    assert float((simt - out).abs().max()) < 1e-5


def test_umma_rhs_golden_powerlaw_h256(knobs, golden):
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi
    knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=16)
    g = golden("powerlaw2048_h256")
    graph = nb.CsrGraph.from_tensor(csr_to_coo(g, "Phi"), torch.device("cuda"))
    W, b = torch.from_numpy(g["W"]).cuda(), torch.from_numpy(g["b"]).cuda()
    x = torch.from_numpy(np.random.RandomState(5).standard_normal((2048, 256)).astype(np.float32)).cuda()
    out = nb.rhs_eval(graph, nb.RhsSpec.ndcn(256, W, b), x).cpu()
    torch.testing.assert_close(out[::4], torch.from_numpy(g["f_x"]), rtol=RTOL, atol=3e-6)


@pytest.mark.parametrize("cw", [-1, 16])
def test_umma_solver_golden_powerlaw_h256(knobs, golden, cw):
    """whole solves (rk4 = 3/8 rule, adaptive dopri5) on the reference's outputs, same step counts"""
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi
    knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=cw)
    g = golden("powerlaw2048_h256")
    graph = nb.CsrGraph.from_tensor(csr_to_coo(g, "Phi"), torch.device("cuda"))
    W, b = torch.from_numpy(g["W"]).cuda(), torch.from_numpy(g["b"]).cuda()
    x = torch.from_numpy(np.random.RandomState(5).standard_normal((2048, 256)).astype(np.float32)).cuda()
    for method, kw in (("rk4", {}), ("dopri5", dict(rtol=.01, atol=.001))):
        y = nb.odeint_fused(graph, nb.RhsSpec.ndcn(256, W, b), x, torch.from_numpy(g["t_" + method]), method=method, **kw)
        torch.testing.assert_close(y[-1][::4].cpu(), torch.from_numpy(g["y_" + method]), rtol=RTOL, atol=1e-5)
        i = _info()
        assert [i.nfe, i.n_accepted, i.n_rejected] == g["stats_" + method].tolist()


@pytest.mark.parametrize("method", ["euler", "midpoint", "dopri5"])
def test_umma_solver_vs_oracle_ragged(knobs, method):
    """n not a multiple of the 128-row tile, more tiles than one wave of CTAs would need on a small
    grid, irregular output times, forced and adaptive stepping"""
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi
    knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=32)
    n, H = 1000, 128
    Phi = _graph(n, 10, seed=3)
    torch.manual_seed(2)
    lin = torch.nn.Linear(H, H)
    W, b = lin.weight.detach() * 0.5, lin.bias.detach()
    x = torch.randn(n, H)
    t = torch.tensor([0.0, 0.3, 0.35, 1.1])
    st = O.SolveStats()
    ref = O.odeint(lambda tt, xx: O.rhs_ndcn(Phi, W, b, xx), x, t, rtol=1e-3, atol=1e-4, method=method, stats=st)
    graph = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    out = nb.odeint_fused(graph, nb.RhsSpec.ndcn(H, W.cuda(), b.cuda()), x.cuda(), t, method=method, rtol=1e-3, atol=1e-4)
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=1e-5)
    i = _info()
    assert [i.nfe, i.n_accepted, i.n_rejected] == [st.nfe, st.n_accepted, st.n_rejected]


def test_umma_many_tiles_per_cta(knobs):
    """more 128-row tiles than SMs: every persistent CTA walks several tiles (both TMEM
    accumulators, both SMEM stages wrap many times); linearity in W as a size-independent check"""
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi
    knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=16)
    n, H = 128 * 148 * 3 + 77, 256
    Phi = _graph(n, 6, seed=8)
    g = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    torch.manual_seed(5)
    W = (torch.randn(H, H) / 16).cuda()
    b = torch.randn(H).cuda()
    x = torch.randn(n, H).cuda()
    spec = nb.RhsSpec.ndcn(H, W, b, relu=False)
    out = nb.rhs_eval(g, spec, x)
    z = nb.spmm(g, x)
    ref = (z.double() @ W.double().t() + b.double())
    err = float((out.double() - ref).abs().max())
    assert err < 5e-6 * max(1.0, float(ref.abs().max())), err
    # sampled rows against the CPU oracle as well
    rows = torch.arange(0, n, 997)
    ref_cpu = O.rhs_ndcn(Phi, W.cpu(), b.cpu(), x.cpu())[rows]
    out_relu = nb.rhs_eval(g, nb.RhsSpec.ndcn(H, W, b), x).cpu()[rows]
    torch.testing.assert_close(out_relu, ref_cpu, rtol=RTOL, atol=3e-6)


@pytest.mark.timeout(900)
def test_full_size_bench_workload_rhs_and_affine_solve(knobs):
    """BASELINE.json's full size (1M-node power-law graph, H=256), default kernel selection:
    (a) one RHS evaluation against the CPU oracle on every row (the oracle needs ~2 s for it);
    (b) size-independent property of the whole solve: without the ReLU the system is affine, so three
        forced dopri5 steps map an affine combination of initial states to the same combination of
        the solutions (every kernel of the step -- gather, tcgen05 GEMM, stage epilogues, error
        stage -- takes part)."""
    import ndcn_b200 as nb
    from ndcn_b200 import workloads as wl
    n, H = 1_000_000, 256
    phi = wl.graph_operator(wl.power_law_adjacency(n, 5, seed=0), "norm_lap")
    g = nb.CsrGraph.from_scipy(phi, torch.device("cuda"))
    torch.manual_seed(0)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5), lin.bias.detach()
    gen = torch.Generator().manual_seed(1)
    x = torch.randn(n, H, generator=gen)
    out = nb.rhs_eval(g, nb.RhsSpec.ndcn(H, W.cuda(), b.cuda()), x.cuda()).cpu()
    ref = O.rhs_ndcn(wl.to_reference_coo(phi), W, b, x)
    torch.testing.assert_close(out, ref, rtol=RTOL, atol=3e-6)
    del out, ref
    spec = nb.RhsSpec.ndcn(H, W.cuda(), b.cuda(), relu=False)
    t = torch.tensor([0.0, 0.125], dtype=torch.float64)

    def solve(y):
        return nb.odeint_fused(g, spec, y, t, method="dopri5", forced_dt=0.05, terminal_only=True).clone()

    xa = x.cuda()
    xb = torch.randn(n, H, generator=gen).cuda()
    sa, sb = solve(xa), solve(xb)
    assert _info().n_accepted == 3
    sc = solve(0.25 * xa + 0.75 * xb)
    comb = 0.25 * sa + 0.75 * sb
    scale = float(comb.abs().max())
    assert float((sc - comb).abs().max()) < 2e-5 * max(scale, 1.0)
    assert bool(torch.isfinite(sc).all())


@pytest.mark.parametrize("H,bc", [(256, 32), (256, 128), (128, 64)])
def test_external_gather_blocked_z(knobs, H, bc):
    """NDCN_GATHER_EXTERNAL (the multi-GPU feature-sharded gather's interface) on one GPU: the exchange
    hook produces z = Phi x in column blocks of `bc` (ndcn_pack_cols_f32), the tcgen05 kernel reads the
    blocked layout; results must equal the library-side gather bit for bit"""
    import ctypes as C
    import ndcn_b200 as nb
    from ndcn_b200 import _ffi, partition
    n = 3000
    Phi = _graph(n, 8, seed=H + bc)
    g = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    torch.manual_seed(3)
    lin = torch.nn.Linear(H, H)
    W, b = (lin.weight.detach() * 0.5).cuda(), lin.bias.detach().cuda()
    x = torch.randn(n, H).cuda()
    spec = nb.RhsSpec.ndcn(H, W, b)
    t = torch.tensor([0.0, 0.4, 1.0])
    knobs(stage_impl=_ffi.IMPL_UMMA, gather_cw=-1)
    ref = nb.odeint_fused(g, spec, x, t, method="dopri5", rtol=1e-3, atol=1e-4).clone()
    ref_info = _info()
    calls = {"gather": 0, "reduce": 0}
    dev = torch.device("cuda")

    def hook(user, what, buf_ptr):
        if what == 2:
            req = C.cast(buf_ptr, C.POINTER(_ffi.GatherRequest)).contents
            src = partition._tensor_from_ptr(req.src_dev, (n, H), torch.float32, dev)
            z = nb.spmm(g, src)
            st = torch.cuda.current_stream(dev).cuda_stream
            _ffi.check(_ffi.lib().ndcn_pack_cols_f32(z.data_ptr(), n, H, bc, req.z_dev, st))
            calls["gather"] += 1
        elif what == 1:
            calls["reduce"] += 1  # one rank: the sums are already global
        else:
            return _ffi.E_ARG
        return 0

    out = nb.odeint_fused(g, spec, x, t, method="dopri5", rtol=1e-3, atol=1e-4, exchange=hook, z_block_cols=bc)
    i = _info()
    assert (i.nfe, i.n_accepted, i.n_rejected) == (ref_info.nfe, ref_info.n_accepted, ref_info.n_rejected)
    assert calls["gather"] == i.nfe and calls["reduce"] > 0
    assert torch.equal(out, ref)
    # fixed-grid solver through the same hook
    ref4 = nb.odeint_fused(g, spec, x, t, method="rk4").clone()
    out4 = nb.odeint_fused(g, spec, x, t, method="rk4", exchange=hook, z_block_cols=bc)
    assert torch.equal(out4, ref4)
