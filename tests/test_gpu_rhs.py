"""GPU: one RHS evaluation through the C ABI vs the CPU oracle (rtol 1e-4 / atol 1e-6, the
north-star parity bar for fp32) and vs the committed reference outputs."""
import numpy as np
import pytest
import torch

from conftest import csr_to_coo, csr_to_dense
from oracle import ndcn_oracle as O

pytestmark = pytest.mark.gpu
RTOL, ATOL = 1e-4, 1e-6


def _rand_graph(n, avg_deg, seed, hub=0, empty_rows=()):
    rs = np.random.RandomState(seed)
    m = int(n * avg_deg / 2)
    r = rs.randint(0, n, m)
    c = rs.randint(0, n, m)
    if hub:
        r = np.concatenate([r, np.zeros(hub, np.int64)])
        c = np.concatenate([c, rs.randint(0, n, hub)])
    keep = r != c
    r, c = r[keep], c[keep]
    rows = np.concatenate([r, c])
    cols = np.concatenate([c, r])
    if len(empty_rows):
        k = ~np.isin(rows, empty_rows) & ~np.isin(cols, empty_rows)
        rows, cols = rows[k], cols[k]
    return rows, cols


def _ours_rhs(A_cpu, spec_fn, x):
    import ndcn_b200 as nb
    g = nb.CsrGraph.from_tensor(A_cpu, torch.device("cuda"))
    return nb.rhs_eval(g, spec_fn(nb), x.cuda()).cpu()


@pytest.mark.parametrize("H", [1, 16, 20, 32, 64, 128, 256, 257])
@pytest.mark.parametrize("flags", ["full", "no_graph", "no_control"])
def test_ndcn_rhs_vs_oracle(H, flags):
    n = 777  # not a multiple of the 64-row tile
    rows, cols = _rand_graph(n, 9, seed=H, hub=300, empty_rows=(5, 6, 700))
    Phi = O.normalized_laplacian_coo(rows, cols, n)
    torch.manual_seed(H)
    lin = torch.nn.Linear(H, H)
    W, b = lin.weight.detach(), lin.bias.detach()
    x = torch.randn(n, H)
    kw = dict(no_graph=flags == "no_graph", no_control=flags == "no_control")
    ref = O.rhs_ndcn(Phi, W, b, x, **kw)
    out = _ours_rhs(Phi, lambda nb: nb.RhsSpec.ndcn(H, W.cuda(), b.cuda(), **kw), x)
    torch.testing.assert_close(out, ref, rtol=RTOL, atol=2e-6)


def test_ndcn_rhs_golden_grid(golden):
    g = golden("ndcn_grid400")
    OM = csr_to_dense(g, "OM")
    W = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__weight"])
    b = torch.from_numpy(g["sd_neural_dynamic_layer__odefunc__wt__bias"])
    x = torch.from_numpy(g["x_probe"])
    for tag, kw in (("full", {}), ("no_graph", dict(no_graph=True)), ("no_control", dict(no_control=True))):
        out = _ours_rhs(OM, lambda nb: nb.RhsSpec.ndcn(20, W.cuda(), b.cuda(), **kw), x)
        torch.testing.assert_close(out, torch.from_numpy(g["f_probe_" + tag]), rtol=RTOL, atol=ATOL)


def test_ndcn_rhs_golden_powerlaw_h256(golden):
    g = golden("powerlaw2048_h256")
    Phi = csr_to_coo(g, "Phi")
    W, b = torch.from_numpy(g["W"]), torch.from_numpy(g["b"])
    x = torch.from_numpy(np.random.RandomState(5).standard_normal((2048, 256)).astype(np.float32))
    out = _ours_rhs(Phi, lambda nb: nb.RhsSpec.ndcn(256, W.cuda(), b.cuda()), x)
    torch.testing.assert_close(out[::4], torch.from_numpy(g["f_x"]), rtol=RTOL, atol=2e-6)


def test_spmm_matches_torch_sparse_and_is_linear():
    import ndcn_b200 as nb
    n, H = 5000, 128
    rows, cols = _rand_graph(n, 12, seed=1, hub=3000)
    Phi = O.normalized_laplacian_coo(rows, cols, n)
    g = nb.CsrGraph.from_tensor(Phi, torch.device("cuda"))
    torch.manual_seed(0)
    x, y = torch.randn(n, H), torch.randn(n, H)
    ref = torch.sparse.mm(Phi, x)
    out = nb.spmm(g, x.cuda())
    torch.testing.assert_close(out.cpu(), ref, rtol=RTOL, atol=2e-6)
    lin = nb.spmm(g, (2.0 * x + y).cuda()) - (2.0 * out + nb.spmm(g, y.cuda()))
    assert float(lin.abs().max()) < 1e-4


@pytest.mark.parametrize("key", ["heat", "gene", "mutual"])
def test_dynamics_rhs_golden(golden, key):
    g = golden("truth_" + key)
    A, L = csr_to_dense(g, "A"), csr_to_dense(g, "L")
    op = -L if key == "heat" else A
    mk = {"heat": lambda nb, d: nb.RhsSpec.heat(d, 1), "gene": lambda nb, d: nb.RhsSpec.gene(d, 1, 1, 2),
          "mutual": lambda nb, d: nb.RhsSpec.mutual(d)}[key]
    for probe, d in (("1", 1), ("3", 3)):
        x = torch.from_numpy(g["x_probe" + probe])
        out = _ours_rhs(op, lambda nb: mk(nb, d), x)
        torch.testing.assert_close(out, torch.from_numpy(g["f_probe" + probe]), rtol=RTOL, atol=1e-5)


@pytest.mark.parametrize("avg_deg", [3, 10, 20, 40])
def test_dynamics_rhs_degree_bins(avg_deg):
    """every lanes-per-row variant of the [N,1] kernel, incl. a hub row and empty rows"""
    n = 3001
    rows, cols = _rand_graph(n, avg_deg, seed=avg_deg, hub=500, empty_rows=(0, 17))
    A = torch.sparse_coo_tensor(torch.from_numpy(np.vstack((rows, cols))), torch.ones(len(rows)), (n, n)).coalesce()
    x = torch.rand(n, 1) * 5 + 0.1
    for name, ref, mk in (("gene", O.rhs_gene(A, x, 1.0), lambda nb: nb.RhsSpec.gene(1, 1.0, 1, 2)),
                          ("mutual", O.rhs_mutual_edgewise(A, x), lambda nb: nb.RhsSpec.mutual(1)),
                          ("heat", O.rhs_heat(-A, x, 0.5), lambda nb: nb.RhsSpec.heat(1, 0.5))):
        out = _ours_rhs(A, mk, x)
        torch.testing.assert_close(out, ref, rtol=RTOL, atol=1e-4, msg=lambda m: name + ": " + m)


def test_empty_graph_and_tiny_inputs():
    import ndcn_b200 as nb
    A = torch.zeros(3, 3)
    x = torch.randn(3, 32)
    out = _ours_rhs(A, lambda nb_: nb_.RhsSpec.ndcn(32, None, None, no_control=True), x)
    assert torch.equal(out, torch.zeros(3, 32))
    A1 = torch.tensor([[2.0]])
    out = _ours_rhs(A1, lambda nb_: nb_.RhsSpec.heat(1, 1.0), torch.tensor([[3.0]]))
    assert float(out) == 6.0
