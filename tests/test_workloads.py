"""CPU: the scale-capable sparse operator builders and node orderings (ndcn_b200/workloads.py) against the
reference's dense / scipy code on a 2k-node graph:
utils_in_learn_dynamics.py:80-134,212-247, propagation.py:45-103, heat_dynamics.py:116-117."""
import numpy as np
import pytest
import scipy.sparse as sp

from ndcn_b200 import workloads as wl
from oracle import ref_loader

N = 2000


def _graph(kind="power_law"):
    if kind == "power_law":
        return wl.power_law_adjacency(N, 5, seed=3)
    if kind == "er":
        return wl.erdos_renyi_adjacency(N, 8.0, seed=3)
    return wl.grid_adjacency(45)


def _dense_formulas(A):
    """closed forms, float64 (independent of the reference import)"""
    deg = A.sum(1)
    dis = np.where(deg > 0, deg ** -0.5, 0.0)
    eye = np.eye(A.shape[0])
    d1 = (deg + 1.0) ** -0.5
    return {"lap": np.diag(deg) - A, "norm_adj": dis[:, None] * A * dis[None, :],
            "norm_lap": eye - dis[:, None] * A * dis[None, :], "kipf": d1[:, None] * (A + eye) * d1[None, :]}


@pytest.mark.parametrize("graph", ["power_law", "er", "grid"])
def test_operators_match_closed_forms(graph):
    a = _graph(graph)
    A = a.toarray().astype(np.float64)
    assert (A == A.T).all() and A.diagonal().sum() == 0 and set(np.unique(A)) <= {0.0, 1.0}
    want = _dense_formulas(A)
    for kind, ref in want.items():
        got = wl.graph_operator(a, kind)
        assert got.dtype == np.float32 and got.has_sorted_indices
        np.testing.assert_allclose(got.toarray(), ref, rtol=2e-6, atol=2e-7, err_msg=kind)
    for alpha in (0.0, 0.1, 0.5):
        ap = alpha * np.eye(len(A)) + (1 - alpha) * A
        d = ap.sum(1)
        dis = np.where(d > 0, d ** -0.5, 0.0)
        np.testing.assert_allclose(wl.graph_operator(a, "alpha", alpha=alpha).toarray(), dis[:, None] * ap * dis[None, :],
                                   rtol=2e-6, atol=2e-7)


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")
def test_operators_match_the_reference_functions():
    """utils_in_learn_dynamics.{normalized_laplacian, normalized_adj, zipf_smoothing}, L = D - A,
    Propagation.zipf_smoothing_alpha (what dgnn.py's load_data applies, utils.py:204-211)"""
    import torch

    ref_loader.import_reference()
    import propagation
    import utils_in_learn_dynamics as uld

    a = _graph("power_law")
    A = a.toarray().astype(np.float32)
    assert (A.sum(1) > 0).all()  # the reference leaves D^-1/2 uninitialised on isolated nodes: none here
    ref = {"norm_lap": uld.normalized_laplacian(A), "norm_adj": uld.normalized_adj(A), "kipf": uld.zipf_smoothing(A),
           "lap": (torch.diag(torch.from_numpy(A).sum(1)) - torch.from_numpy(A)).numpy()}
    for kind, want in ref.items():
        np.testing.assert_allclose(wl.graph_operator(a, kind).toarray(), np.asarray(want, dtype=np.float32),
                                   rtol=1e-6, atol=1e-7, err_msg=kind)
    import contextlib
    import io

    for alpha in (0.0, 0.5, 0.9):
        with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints the matrix type
            want = propagation.Propagation(a.astype(np.float64)).zipf_smoothing_alpha(alpha)
        np.testing.assert_allclose(wl.graph_operator(a, "alpha", alpha=alpha).toarray(),
                                   np.asarray(want.todense(), dtype=np.float32), rtol=1e-6, atol=1e-7)
    # the tensor format the reference hands to ODEFunc at scale (utils.py:12-23): uncoalesced fp32 COO, int64 indices
    import utils

    phi = wl.graph_operator(a, "norm_lap")
    t_ref = utils.sparse_csr_matrix_to_torch_sparse_tensor(phi)
    t_ours = wl.to_reference_coo(phi)
    assert t_ours.dtype == t_ref.dtype and t_ours._indices().dtype == t_ref._indices().dtype
    assert torch.equal(t_ours._indices(), t_ref._indices()) and torch.equal(t_ours._values(), t_ref._values())


@pytest.mark.skipif(not ref_loader.reference_available(), reason="needs /root/reference")
@pytest.mark.parametrize("layout", ["degree", "community"])
def test_reorder_matches_networkx_reorder_nodes(layout):
    """--layout degree | community (utils_in_learn_dynamics.py:212-247)"""
    import networkx as nx

    ref_loader.import_reference()
    import utils_in_learn_dynamics as uld

    g = nx.barabasi_albert_graph(300, 5, seed=0)
    a = sp.csr_matrix(nx.to_scipy_sparse_array(g, format="csr"), dtype=np.float32)
    want = nx.to_scipy_sparse_array(uld.networkx_reorder_nodes(g, layout), nodelist=range(300), format="csr")
    got, order = wl.reorder(a, layout)
    assert sorted(order.tolist()) == list(range(300))
    assert (got.toarray() == np.asarray(want.todense(), dtype=np.float32)).all()


@pytest.mark.parametrize("kind", ["rcm", "bfs", "degree"])
def test_scalable_orderings_are_permutations_that_preserve_the_operator(kind):
    a = _graph("power_law")
    p, order = wl.reorder(a, kind)
    assert sorted(order.tolist()) == list(range(N))
    assert (p.toarray() == a.toarray()[np.ix_(order, order)]).all()
    # Phi of the permuted graph = permuted Phi: reordering is a similarity transform the solver is indifferent to
    phi, phi_p = wl.graph_operator(a, "norm_lap").toarray(), wl.graph_operator(p, "norm_lap").toarray()
    np.testing.assert_allclose(phi_p, phi[np.ix_(order, order)], rtol=0, atol=0)
    if kind == "rcm":  # the point of RCM: smaller bandwidth than generation order
        bw = lambda m: int(np.abs(np.subtract(*m.nonzero())).max())  # noqa: E731
        assert bw(p) < bw(a)


def test_generators_shapes_and_degree_law():
    a = wl.power_law_adjacency(20000, 5, seed=0)
    deg = np.asarray(a.sum(1)).ravel()
    assert a.shape == (20000, 20000) and deg.min() >= 1 and 9.0 < deg.mean() < 10.1
    assert deg.max() > 150  # hubs ~ m sqrt(N)
    e = wl.erdos_renyi_adjacency(20000, 10.0, seed=0)
    assert 9.5 < np.asarray(e.sum(1)).mean() < 10.5
    g = wl.grid_adjacency(20)
    assert g.nnz == 2964  # grid_8_neighbor_graph(20), SURVEY.md section 3.1


# ------------------------------------------------------------------------------------------
# the scripts' --network choices without dense matrices (heat_dynamics.py:83-110)
# ------------------------------------------------------------------------------------------
def _is_simple_undirected(a):
    d = a.toarray()
    return (d == d.T).all() and d.diagonal().sum() == 0 and set(np.unique(d)) <= {0.0, 1.0}


def test_small_world_ring_plus_shortcuts():
    n, k, p = 3000, 5, 0.5
    a = wl.small_world_adjacency(n, k, p, seed=1)
    assert _is_simple_undirected(a)
    d = a.toarray()
    i = np.arange(n)
    for j in (1, 2):  # the ring lattice is always there
        assert (d[i, (i + j) % n] == 1).all()
    ring = n * (k // 2)
    extra = a.nnz // 2 - ring
    # one Bernoulli(p) shortcut per ring edge, minus the few that hit an existing edge
    assert abs(extra - p * ring) < 4 * np.sqrt(ring * p * (1 - p)) + 0.01 * ring
    # networkx draws from the same law: same mean degree within sampling noise
    nx = pytest.importorskip("networkx")
    g = nx.newman_watts_strogatz_graph(n, k, p, seed=1)
    assert abs(g.number_of_edges() - a.nnz // 2) < 0.03 * g.number_of_edges()


def test_community_graph_block_densities():
    sizes = [700, 700, 500, 200]
    a = wl.community_adjacency(sizes, 0.25, 0.01, seed=2)
    assert _is_simple_undirected(a) and a.shape[0] == sum(sizes)
    d = a.toarray()
    start = np.concatenate([[0], np.cumsum(sizes)])
    for x in range(4):
        for y in range(4):
            blk = d[start[x]:start[x + 1], start[y]:start[y + 1]]
            dens = blk.sum() / (sizes[x] * (sizes[x] - 1) if x == y else blk.size)
            # pairs are sampled with replacement: the density is 1 - exp(-p) (~p - p^2/2)
            want = 1 - np.exp(-(0.25 if x == y else 0.01))
            assert abs(dens - want) < 0.06 * want + 1e-3, (x, y, dens, want)


@pytest.mark.parametrize("kind", ["grid", "random", "power_law", "small_world", "community"])
def test_network_dispatch_sizes_and_mean_degree(kind):
    n = 900
    a = wl.network(kind, n, seed=0)
    assert a.shape == (n, n) and _is_simple_undirected(a)
    if kind in ("random", "community"):
        b = wl.network(kind, 20000, seed=0, mean_degree=12.0)
        assert b.shape == (20000, 20000)
        assert abs(b.nnz / 20000 - 12.0) < 0.5
    with pytest.raises(ValueError):
        wl.network("ring", n)


def test_experiment_host_logic_follows_the_scripts():
    """initial value and time sampling of the scale driver = the scripts' own statements
    (heat_dynamics.py:119-151,178-183), re-evaluated here from the same numpy seed."""
    import torch

    from ndcn_b200 import experiment as ex

    x0 = ex.initial_value(400)
    ref = torch.zeros(20, 20)
    ref[1:5, 1:5] = 25
    ref[9:15, 9:15] = 20
    ref[1:5, 7:13] = 17
    assert torch.equal(x0, ref.view(-1, 1))
    assert ex.initial_value(390).shape == (390, 1) and torch.equal(ex.initial_value(390), x0[:390])

    t, tr, te, te2 = ex.time_ticks("equal", 5.0, 100)
    assert torch.equal(t, torch.linspace(0.0, 5.0, 100)) and tr == list(range(80)) and te == list(range(80, 100))
    assert te2 is None

    np.random.seed(7)
    t, tr, te, te2 = ex.time_ticks("irregular", 5.0, 100)
    np.random.seed(7)
    tt = torch.linspace(0.0, 5.0, 1000)
    tt = torch.tensor(np.sort(np.random.permutation(tt)[:120]))
    tt[0] = 0
    want2 = sorted(np.random.permutation(range(1, 100))[:20].tolist())
    assert torch.equal(t, tt) and te == list(range(100, 120)) and te2 == want2
    assert tr == sorted(set(range(100)) - set(want2)) and len(tr) == 80
    args = ex.parser().parse_args(["--network", "community", "--n", "1200", "--layout", "degree"])
    a = ex.build_graph(args)
    deg = np.diff(a.indptr)
    assert a.shape == (1200, 1200) and (np.diff(deg) <= 0).all()
