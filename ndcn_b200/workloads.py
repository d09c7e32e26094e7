"""Scale-capable graph generators and sparse operator builders (host side, start-up only).

The reference builds every graph through networkx into a DENSE ``[N, N]`` numpy matrix
(heat_dynamics.py:83-110) and every operator with dense algebra
(utils_in_learn_dynamics.py:80-134), which stops at a few tens of thousands of nodes.  The
BASELINE configurations (100k - 4M nodes) need the same objects built sparsely; this module
does that with numpy/scipy only and never materialises N x N.  Nothing here is on the timed
path: the output is a scipy CSR matrix that ``CsrGraph.from_scipy`` uploads once.

Operator definitions follow the reference:
  norm_lap  I - D^-1/2 A D^-1/2        utils_in_learn_dynamics.py:109-120   (the scripts' default)
  norm_adj  D^-1/2 A D^-1/2            utils_in_learn_dynamics.py:123-134
  kipf      (D+I)^-1/2 (A+I) (D+I)^-1/2   utils_in_learn_dynamics.py:80-92
  lap       D - A                       heat_dynamics.py:116-117
  alpha     (aI+(1-a)D)^-1/2 (aI+(1-a)A) (aI+(1-a)D)^-1/2   propagation.py:91-103 -- what dgnn.py runs on
            (utils.py:204-211, ``--alpha``, default .5; README flags use 0: the plain normalized adjacency)
with D^-1/2 := 0 on isolated nodes (the reference leaves those entries uninitialised,
utils_in_learn_dynamics.py:117-118).

Node orderings (``reorder``): the reference's ``--layout degree | community``
(utils_in_learn_dynamics.py:212-247) and, for graphs beyond networkx's reach, reverse Cuthill-McKee and
breadth-first orders from scipy.sparse.csgraph (gather locality / halo size of a row partition).
"""
from __future__ import annotations

from typing import Tuple

import numpy as np
import scipy.sparse as sp


# ----------------------------------------------------------------------------------------
# edge generators: return (rows, cols) of an undirected simple graph, both directions present
# ----------------------------------------------------------------------------------------
def _symmetrize(src: np.ndarray, dst: np.ndarray, n: int) -> sp.csr_matrix:
    keep = src != dst
    src, dst = src[keep], dst[keep]
    a = sp.coo_matrix((np.ones(2 * len(src), np.float32), (np.concatenate([src, dst]), np.concatenate([dst, src]))),
                      shape=(n, n)).tocsr()
    a.sum_duplicates()
    a.data[:] = 1.0  # multi-edges collapse to one
    a.sort_indices()
    return a


def power_law_adjacency(n: int, m: int = 5, seed: int = 0) -> sp.csr_matrix:
    """Barabasi-Albert-style preferential attachment (``nx.barabasi_albert_graph(n, 5)`` in the
    scripts, heat_dynamics.py:94), generated in O(n m) vectorised numpy.

    Every new node v >= m draws m targets from the list of all earlier edge endpoints
    ("repeated nodes" list): endpoint slot 2e is the source of edge e, slot 2e+1 its target.
    Drawing slot r < 2*m*(v-m) either names a source directly or points at an earlier target,
    which is resolved by pointer jumping.  Multi-edges are collapsed, so a few nodes end up with
    fewer than m new links (networkx re-draws instead); the degree law is the same k^-3.
    """
    assert n > m >= 1
    rs = np.random.RandomState(seed)
    n_new = n - m
    n_edges = n_new * m
    src = np.repeat(np.arange(m, n, dtype=np.int64), m)
    e = np.arange(n_edges, dtype=np.int64)
    v_idx = e // m  # 0-based index of the new node
    avail = 2 * m * v_idx  # endpoint slots that exist before node v arrives
    draw = (rs.random_sample(n_edges) * np.maximum(avail, 1)).astype(np.int64)
    # the first new node has no earlier edges: it links to the m seed nodes
    first = avail == 0
    tgt = np.full(n_edges, -1, np.int64)
    tgt[first] = e[first] % m
    # even slot -> source of edge draw//2 (known); odd slot -> target of edge draw//2 (chase)
    ref = draw // 2
    is_src = (draw % 2 == 0) & ~first
    tgt[is_src] = src[ref[is_src]]
    pending = np.flatnonzero(tgt < 0)
    ptr = ref.copy()
    while len(pending):
        t = tgt[ptr[pending]]
        done = t >= 0
        tgt[pending[done]] = t[done]
        # still unresolved: follow the chain one more hop (ptr of the referenced edge)
        rest = pending[~done]
        ptr[rest] = ptr[ptr[rest]]
        pending = rest
    return _symmetrize(src, tgt, n)


def erdos_renyi_adjacency(n: int, mean_degree: float = 10.0, seed: int = 0) -> sp.csr_matrix:
    """G(n, p = mean_degree / n) by sampling the edge count's expectation worth of pairs.
    (The script's ``nx.erdos_renyi_graph(n, 0.1)``, heat_dynamics.py:89, would be 5e10 edges at
    n = 1M; BASELINE config 4 uses mean degree 10.)"""
    rs = np.random.RandomState(seed)
    k = int(round(n * mean_degree / 2.0))
    src = rs.randint(0, n, k).astype(np.int64)
    dst = rs.randint(0, n, k).astype(np.int64)
    return _symmetrize(src, dst, n)


def grid_adjacency(side: int) -> sp.csr_matrix:
    """8-neighbour grid on side x side nodes, row-major ids (grid_8_neighbor_graph,
    utils_in_learn_dynamics.py:137-157)."""
    idx = np.arange(side * side, dtype=np.int64).reshape(side, side)
    src, dst = [], []
    for di, dj in ((0, 1), (1, 0), (1, 1), (1, -1)):
        i0, i1 = max(0, -di), side - max(0, di)
        j0, j1 = max(0, -dj), side - max(0, dj)
        src.append(idx[i0:i1, j0:j1].ravel())
        dst.append(idx[i0 + di:i1 + di, j0 + dj:j1 + dj].ravel())
    return _symmetrize(np.concatenate(src), np.concatenate(dst), side * side)


def small_world_adjacency(n: int, k: int = 5, p: float = 0.5, seed: int = 0) -> sp.csr_matrix:
    """Newman-Watts-Strogatz graph (``nx.newman_watts_strogatz_graph(400, 5, 0.5)`` in the scripts,
    heat_dynamics.py:99): a ring in which every node is linked to its k // 2 nearest neighbours on
    either side, plus, for every ring edge, one extra shortcut from its first endpoint to a uniformly
    drawn node with probability p.  (networkx re-draws a shortcut that hits an existing edge or the
    node itself; here such a draw is dropped -- a fraction ~k/n of the shortcuts.)"""
    assert n > k >= 2
    rs = np.random.RandomState(seed)
    u = np.arange(n, dtype=np.int64)
    src = [u for _ in range(1, k // 2 + 1)]
    dst = [(u + j) % n for j in range(1, k // 2 + 1)]
    ring_src = np.concatenate(src)
    add = rs.random_sample(len(ring_src)) < p
    w = rs.randint(0, n, int(add.sum())).astype(np.int64)
    src.append(ring_src[add])
    dst.append(w)
    return _symmetrize(np.concatenate(src), np.concatenate(dst), n)


def community_adjacency(sizes, p_in: float = 0.25, p_out: float = 0.01, seed: int = 0) -> sp.csr_matrix:
    """Planted-partition graph (``nx.random_partition_graph([n/3, n/3, n/4, rest], .25, .01)`` in
    the scripts, heat_dynamics.py:104-109) without the N x N Bernoulli matrix: for every pair of
    blocks the number of edges is drawn from its binomial law and that many node pairs are sampled
    uniformly (duplicates collapse, a fraction ~p of them), nodes numbered block after block."""
    rs = np.random.RandomState(seed)
    sizes = [int(s) for s in sizes]
    start = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    src, dst = [], []
    for a in range(len(sizes)):
        for b in range(a, len(sizes)):
            pairs = sizes[a] * (sizes[a] - 1) // 2 if a == b else sizes[a] * sizes[b]
            m = rs.binomial(pairs, p_in if a == b else p_out) if pairs else 0
            if m == 0:
                continue
            src.append(start[a] + rs.randint(0, sizes[a], m).astype(np.int64))
            dst.append(start[b] + rs.randint(0, sizes[b], m).astype(np.int64))
    if not src:
        return sp.csr_matrix((int(start[-1]), int(start[-1])), dtype=np.float32)
    return _symmetrize(np.concatenate(src), np.concatenate(dst), int(start[-1]))


def network(kind: str, n: int, seed: int = 0, mean_degree: float = None) -> sp.csr_matrix:
    """The scripts' ``--network`` choices (heat_dynamics.py:83-110) at any size, sparse.

    ``grid`` uses side = ceil(sqrt(n)) like the scripts (so the graph has side**2 >= n nodes);
    ``random`` and ``community`` keep the scripts' edge probabilities (0.1; 0.25 / 0.01) unless
    ``mean_degree`` is given, in which case the probabilities are scaled to that mean degree --
    at 1M nodes p = 0.1 would be 5e10 edges."""
    if kind == "grid":
        return grid_adjacency(int(np.ceil(np.sqrt(n))))
    if kind == "random":
        return erdos_renyi_adjacency(n, 0.1 * (n - 1) if mean_degree is None else mean_degree, seed)
    if kind == "power_law":
        return power_law_adjacency(n, 5, seed)
    if kind == "small_world":
        return small_world_adjacency(n, 5, 0.5, seed)
    if kind == "community":
        n1, n2, n3 = n // 3, n // 3, n // 4
        sizes = [n1, n2, n3, n - n1 - n2 - n3]
        p_in, p_out = 0.25, 0.01
        if mean_degree is not None:
            # same in/out ratio of expected degrees as the scripts' 0.25 / 0.01 on four blocks
            base = sum(s * (p_in * (s - 1) + p_out * (n - s)) for s in sizes) / n
            p_in, p_out = p_in * mean_degree / base, p_out * mean_degree / base
        return community_adjacency(sizes, p_in, p_out, seed)
    raise ValueError("unknown network %r" % (kind,))


def reorder_by_degree(a: sp.csr_matrix) -> Tuple[sp.csr_matrix, np.ndarray]:
    """``--layout degree`` (utils_in_learn_dynamics.py:212-247): nodes sorted by decreasing degree
    (stable).  Returns the permuted adjacency and ``perm`` with new_id = rank of old id."""
    deg = np.asarray(a.sum(1)).ravel()
    order = np.argsort(-deg, kind="stable")
    p = a[order][:, order].tocsr()
    p.sort_indices()
    return p, order


# ----------------------------------------------------------------------------------------
# operators
# ----------------------------------------------------------------------------------------
def _inv_sqrt(deg: np.ndarray) -> np.ndarray:
    out = np.zeros_like(deg, dtype=np.float32)
    nz = deg != 0
    out[nz] = np.power(deg[nz].astype(np.float32), -0.5)
    return out


def reorder(a: sp.csr_matrix, kind: str = "degree") -> Tuple[sp.csr_matrix, np.ndarray]:
    """Permuted adjacency and ``order`` (order[new_id] = old_id).

    ``degree``     decreasing degree, ties in node order (generate_node_mapping, utils_in_learn_dynamics.py:218-220)
    ``community``  greedy-modularity communities one after the other, largest first, members in networkx's set
                   order (:221-226; networkx, so a few 10k nodes at most)
    ``rcm``        reverse Cuthill-McKee (bandwidth reduction; scales to millions of nodes)
    ``bfs``        breadth-first order from the highest-degree node
    """
    n = a.shape[0]
    if kind == "degree":
        return reorder_by_degree(a)
    if kind == "community":
        import networkx as nx
        from networkx.algorithms import community

        g = nx.from_scipy_sparse_array(a)
        order = np.array([v for c in community.greedy_modularity_communities(g) for v in c], dtype=np.int64)
    elif kind == "rcm":
        from scipy.sparse.csgraph import reverse_cuthill_mckee

        order = np.asarray(reverse_cuthill_mckee(a.tocsr(), symmetric_mode=True), dtype=np.int64)
    elif kind == "bfs":
        from scipy.sparse.csgraph import breadth_first_order

        deg = np.asarray(a.sum(1)).ravel()
        seen = np.zeros(n, bool)
        parts = []
        for start in np.argsort(-deg, kind="stable"):  # one tree per connected component
            if seen[start]:
                continue
            o = breadth_first_order(a, int(start), directed=False, return_predecessors=False)
            seen[o] = True
            parts.append(o.astype(np.int64))
        order = np.concatenate(parts)
    else:
        raise ValueError("unknown ordering %r" % (kind,))
    assert len(order) == n and len(np.unique(order)) == n
    p = a[order][:, order].tocsr()
    p.sort_indices()
    return p, order


def graph_operator(a: sp.csr_matrix, kind: str = "norm_lap", alpha: float = 0.5) -> sp.csr_matrix:
    """fp32 CSR of the chosen operator (see module docstring)."""
    n = a.shape[0]
    a = a.astype(np.float32).tocsr()
    deg = np.asarray(a.sum(1)).ravel().astype(np.float32)
    eye = sp.identity(n, dtype=np.float32, format="csr")
    if kind == "lap":
        m = sp.diags(deg).tocsr() - a
    elif kind == "norm_adj":
        d = sp.diags(_inv_sqrt(deg))
        m = d @ a @ d
    elif kind == "kipf":
        d = sp.diags(_inv_sqrt(deg + 1.0))
        m = d @ (a + eye) @ d
    elif kind == "norm_lap":
        d = sp.diags(_inv_sqrt(deg))
        m = eye - d @ a @ d
    elif kind == "alpha":
        # propagation.py:91-103: degrees of A' = a I + (1-a) A taken in fp32, rows and columns scaled separately
        ap = (alpha * eye.astype(np.float64) + (1.0 - alpha) * a.astype(np.float64)).tocsr()
        out_deg = np.asarray(ap.sum(1), dtype=np.float32).ravel()
        in_deg = np.asarray(ap.sum(0), dtype=np.float32).ravel()
        m = sp.diags(_inv_sqrt(out_deg).astype(np.float64)) @ ap @ sp.diags(_inv_sqrt(in_deg).astype(np.float64))
    else:
        raise ValueError("unknown operator %r" % (kind,))
    m = m.tocsr().astype(np.float32)
    m.sort_indices()
    return m


def to_reference_coo(m: sp.csr_matrix):
    """The tensor format the reference hands to ODEFunc for large graphs: an UNCOALESCED fp32 torch
    sparse COO tensor with int64 indices in row-major entry order
    (utils.py:12-23 sparse_csr_matrix_to_torch_sparse_tensor)."""
    import torch

    c = m.tocoo()
    idx = torch.from_numpy(np.vstack((c.row, c.col)).astype(np.int64))
    return torch.sparse_coo_tensor(idx, torch.from_numpy(c.data.astype(np.float32)), c.shape)
