"""Generate tests/golden/*.npz by running the UNMODIFIED reference in-process.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
Every fixture stores the exact inputs handed to the reference (graph operator in CSR
form, initial state, time grid, weights) and the reference's outputs (fp32, torch 2.11
CPU).  While generating, the CPU oracle is checked to be BIT-IDENTICAL to the reference
on every case -- that is what pins the oracle.
"""
from __future__ import annotations

import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")

from oracle import ndcn_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402

ARGV = ["--network", "grid", "--T", "5", "--sampled_time", "equal", "--baseline", "ndcn", "--gpu", "-1"]


def csr_of(dense_or_coo: torch.Tensor):
    m = dense_or_coo.to_dense() if dense_or_coo.is_sparse else dense_or_coo
    csr = m.to_sparse_csr()
    return dict(rowptr=csr.crow_indices().numpy().astype(np.int32),
                col=csr.col_indices().numpy().astype(np.int32),
                val=csr.values().numpy().astype(np.float32),
                n=np.int64(m.shape[0]))


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024))


def pref(prefix, d):
    return {prefix + "_" + k: v for k, v in d.items()}


def main():
    nd, ode = ref_loader.import_reference()
    import utils_in_learn_dynamics as uld

    torch.set_num_threads(1)  # reproducible summation order

    # ---------------- ground-truth dynamics on the 400-node grid (BASELINE config 1) --------------
    for script, key, rhs in (("heat_dynamics.py", "heat", None),
                             ("gene_dynamics.py", "gene", None),
                             ("mutualistic_dynamics.py", "mutual", None)):
        g = ref_loader.run_reference_script(script, ARGV)
        A, L, x0, t = g["A"], g["L"], g["x0"], g["t"]
        ref = g["solution_numerical"]  # [100, 400, 1]
        if key == "heat":
            f = lambda tt, x: O.rhs_heat(L, x, 1)
            cls = g["HeatDiffusion"](L, 1)
        elif key == "gene":
            f = lambda tt, x: O.rhs_gene(A, x, 1)
            cls = g["GeneDynamics"](A, 1)
        else:
            f = lambda tt, x: O.rhs_mutual_dense(A, x)
            cls = g["MutualDynamics"](A)
        st = O.SolveStats()
        mine = O.odeint(f, x0, t, method="dopri5", stats=st)
        assert torch.equal(mine, ref), key
        # also the sparse-COO flavour of the same RHS (--sparse)
        As, Ls = uld.torch_sensor_to_torch_sparse_tensor(A), uld.torch_sensor_to_torch_sparse_tensor(L)
        with torch.no_grad():
            if key == "heat":
                ref_sp = ode.odeint(g["HeatDiffusion"](Ls, 1), x0, t, method="dopri5")
            elif key == "gene":
                ref_sp = ode.odeint(g["GeneDynamics"](As, 1), x0, t, method="dopri5")
            else:
                ref_sp = ode.odeint(g["MutualDynamics"](As), x0, t, method="dopri5")
            # one RHS evaluation on a rough state, d = 1 and (mutual: the disagreeing) d = 3
            rs = np.random.RandomState(7)
            xr1 = torch.from_numpy(rs.uniform(0.1, 8.0, (400, 1)).astype(np.float32))
            f1 = cls(torch.tensor(0.0), xr1)
            xr3 = torch.from_numpy(rs.uniform(0.1, 8.0, (400, 3)).astype(np.float32))
            if key == "mutual":
                f3 = g["MutualDynamics"](As)(torch.tensor(0.0), xr3)  # python loop over nnz, slow but exact
                assert torch.allclose(O.rhs_mutual_edgewise(As, xr3), f3, rtol=1e-5, atol=1e-6)
                assert torch.allclose(O.rhs_mutual_edgewise(As, xr1), f1, rtol=1e-5, atol=1e-5)
            else:
                f3 = cls(torch.tensor(0.0), xr3)
        save("truth_" + key,
             **pref("A", csr_of(A)), **pref("L", csr_of(L)), x0=x0.numpy(), t=t.numpy(),
             sol_dense=ref.numpy(), sol_sparse=ref_sp.numpy(),
             x_probe1=xr1.numpy(), f_probe1=f1.numpy(), x_probe3=xr3.numpy(), f_probe3=f3.numpy(),
             nfe=np.int64(st.nfe), n_accepted=np.int64(st.n_accepted), n_rejected=np.int64(st.n_rejected))
        if key == "heat":
            OM = g["OM"]

    # ---------------- NDCN forward on the grid: all four methods, dense and sparse Phi --------------
    g = ref_loader.run_reference_script("heat_dynamics.py", ARGV)
    OM, x0, t = g["OM"], g["x0"], g["t"]
    OMs = uld.torch_sensor_to_torch_sparse_tensor(OM)
    out = dict(**pref("OM", csr_of(OM)), x0=x0.numpy(), t=t.numpy())
    for method in ("euler", "midpoint", "rk4", "dopri5"):
        torch.manual_seed(0)
        model = nd.NDCN(1, 20, OM, 1, rtol=.01, atol=.001, method=method)
        with torch.no_grad():
            y = model(t, x0)                      # [100, 400, 1]
            h0 = model.input_layer(x0)
            hv_ref = model.neural_dynamic_layer(t, h0)  # [100, 400, 20]
            W, b = model.neural_dynamic_layer.odefunc.wt.weight, model.neural_dynamic_layer.odefunc.wt.bias
            st = O.SolveStats()
            hv = O.odeint(lambda tt, x: O.rhs_ndcn(OM, W, b, x), h0, t.type_as(h0),
                          rtol=.01, atol=.001, method=method, stats=st)
            assert torch.equal(hv, hv_ref), method
            model_s = nd.NDCN(1, 20, OMs, 1, rtol=.01, atol=.001, method=method)
            model_s.load_state_dict(model.state_dict())
            y_s = model_s(t, x0)
        if method == "euler":
            for k, v in model.state_dict().items():
                out["sd_" + k.replace(".", "__")] = v.numpy()
            out["h0"] = h0.numpy()
        out["y_" + method] = y.numpy()
        out["y_sparse_" + method] = y_s.numpy()
        out["hv_every10_" + method] = hv_ref[::10].numpy()
        out["stats_" + method] = np.array([st.nfe, st.n_accepted, st.n_rejected], np.int64)
    # ODEFunc single evaluations incl. ablation switches (neural_dynamics.py:20-39)
    rs = np.random.RandomState(3)
    xp = torch.from_numpy(rs.standard_normal((400, 20)).astype(np.float32))
    out["x_probe"] = xp.numpy()
    torch.manual_seed(0)
    for tag, kw in (("full", {}), ("no_graph", dict(no_graph=True)), ("no_control", dict(no_control=True))):
        fn = nd.ODEFunc(20, OM, **kw)
        fn.wt.weight.data.copy_(W); fn.wt.bias.data.copy_(b)
        with torch.no_grad():
            fr = fn(torch.tensor(0.0), xp)
            assert torch.equal(fr, O.rhs_ndcn(OM, W, b, xp, **kw))
        out["f_probe_" + tag] = fr.numpy()
    save("ndcn_grid400", **out)

    # ---------------- Cora differential_gcn block (BASELINE config 2) --------------
    cwd = os.getcwd()
    os.chdir(ref_loader.REFERENCE_ROOT)
    try:
        import utils as ref_utils
        cora = {}
        for alpha in (0.5, 0.0):
            adj = ref_utils.load_data("cora", alpha)[0]
            cora[alpha] = adj
    finally:
        os.chdir(cwd)
    out = {}
    for alpha, adj in cora.items():
        tag = "a%02d" % int(alpha * 10)
        out.update(pref("adj_" + tag, csr_of(adj)))
        for H in (32, 256):
            for no_control in (False, True):
                torch.manual_seed(1)
                tt = torch.linspace(0, 1.2, 16).float()
                blk = nd.ODEBlock2(nd.ODEFunc(H, adj, dropout=0.0, no_control=no_control), tt,
                                   rtol=.1, atol=.1, method="dopri5", terminal=True)
                x = torch.from_numpy(np.tanh(np.random.RandomState(11).standard_normal((2708, H))).astype(np.float32))
                with torch.no_grad():
                    yT = blk(x)
                    W, b = blk.odefunc.wt.weight, blk.odefunc.wt.bias
                    st = O.SolveStats()
                    full = O.odeint(lambda t_, x_: O.rhs_ndcn(adj, W, b, x_, no_control=no_control), x, tt,
                                    rtol=.1, atol=.1, method="dopri5", stats=st)
                    assert torch.equal(full[-1], yT)
                key = "%s_h%d_%s" % (tag, H, "noctl" if no_control else "ctl")
                out["W_" + key] = W.detach().numpy()
                out["b_" + key] = b.detach().numpy()
                # H=256 states are 2.7 MB each: keep every 8th row (x is regenerated from its seed)
                out["yT_" + key] = yT.numpy() if H == 32 else yT[::8].numpy()
                out["stats_" + key] = np.array([st.nfe, st.n_accepted, st.n_rejected], np.int64)
    save("cora_block", **out)

    # ---------------- small power-law graph: operator in the at-scale COO format, H=256 --------------
    import networkx as nx
    n = 2048
    G = nx.barabasi_albert_graph(n, 5, seed=0)
    e = np.array(G.edges(), dtype=np.int64)
    rows = np.concatenate([e[:, 0], e[:, 1]]); cols = np.concatenate([e[:, 1], e[:, 0]])
    Phi = O.normalized_laplacian_coo(rows, cols, n)           # uncoalesced COO, as utils.py:12-23 builds it
    # the same operator through the reference's dense builder (utils_in_learn_dynamics.py:109-120)
    Ad = torch.zeros(n, n); Ad[rows, cols] = 1
    Phi_ref = torch.FloatTensor(uld.normalized_laplacian(Ad.numpy()))
    assert torch.allclose(Phi.to_dense(), Phi_ref, atol=1e-6)
    H = 256
    torch.manual_seed(2)
    fn = nd.ODEFunc(H, Phi)
    with torch.no_grad():
        fn.wt.weight.mul_(0.5)
    x = torch.from_numpy(np.random.RandomState(5).standard_normal((n, H)).astype(np.float32))
    out = dict(edges=e.astype(np.int32), **pref("Phi", csr_of(Phi)), W=fn.wt.weight.detach().numpy(),
               b=fn.wt.bias.detach().numpy())
    with torch.no_grad():
        out["f_x"] = fn(torch.tensor(0.0), x).numpy()[::4]
        for method, tt, kw in (("rk4", torch.linspace(0, 1.0, 6), {}),
                               ("dopri5", torch.tensor([0.0, 0.4, 1.0]), dict(rtol=.01, atol=.001))):
            yr = ode.odeint(fn, x, tt, method=method, **kw)
            st = O.SolveStats()
            W, b = fn.wt.weight, fn.wt.bias
            ym = O.odeint(lambda t_, x_: O.rhs_ndcn(Phi, W, b, x_), x, tt, method=method, stats=st, **kw)
            assert torch.equal(yr, ym), method
            out["y_" + method] = yr[-1].numpy()[::4]
            out["t_" + method] = tt.numpy()
            out["stats_" + method] = np.array([st.nfe, st.n_accepted, st.n_rejected], np.int64)
    save("powerlaw2048_h256", **out)
    print("all oracle == reference checks passed (bit-exact)")


if __name__ == "__main__":
    main()
