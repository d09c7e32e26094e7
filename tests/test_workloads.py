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
