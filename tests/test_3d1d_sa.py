"""SA_AMG, the HAZmath .dat input, the .npy/solution.txt formats and the synthetic 3D-1D system
(BASELINE configs[4])."""
import os

import numpy as np
import pytest
import scipy.sparse as sp

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import datfile, haznics_compat as haznics, params, problems, solver_files
from metric_amg_examples_b200.problems.emi3d1d import circle_average, emi3d1d_system, segment_graph
from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DAT = os.path.join(ROOT, "tests", "golden", "input_metric.dat")


def test_dat_reader_matches_reference_file():
    """tests/golden/input_metric.dat restates the settings of src/input_metric.dat:51-100."""
    solver, amg = datfile.read_input(DAT)
    assert solver == {"type": "cg", "maxit": 1000, "tol": 1e-6, "stop_type": 1, "restart": 100,
                      "precond_type": 16, "print_level": 2}
    assert amg["AMG_type"] == haznics.SA_AMG and amg["cycle_type"] == haznics.V_CYCLE
    assert amg["smoother"] == haznics.SMOOTHER_GS and amg["coarse_scaling"] == haznics.OFF
    assert amg["coarse_dof"] == 300 and amg["aggregation_type"] == haznics.VMB
    assert amg["max_aggregation"] == 20 and amg["strong_coupled"] == 0.0
    assert amg["Schwarz_mmsize"] == 200 and amg["Schwarz_maxlvl"] == 2 and amg["Schwarz_type"] == 3
    ref = "/root/reference/src/input_metric.dat"
    if os.path.exists(ref):
        assert datfile.read_input(ref) == (solver, amg)


def test_segment_graph_is_a_tree_on_mesh_edges():
    n = 12
    ids, edges, xyz = segment_graph(n, 200, seed=0)
    assert len(edges) == len(ids) - 1 and len(np.unique(ids)) == len(ids)
    d = np.abs(xyz[edges[:, 0]] - xyz[edges[:, 1]]) * n
    assert np.allclose(d, np.round(d)) and np.all(d < 1 + 1e-9) and np.all(d.sum(axis=1) > 1 - 1e-9)
    ids2, edges2, _ = segment_graph(n, 200, seed=0)
    assert np.array_equal(ids, ids2) and np.array_equal(edges, edges2)


def test_3d1d_system_properties():
    s = emi3d1d_system(8, gamma=1e3)
    n3, n1 = s.W[0].dim(), s.W[1].dim()
    assert s.ndofs == n3 + n1 and np.array_equal(s.interface_dofs, np.arange(n3, n3 + n1))
    assert abs(s.A - s.A.T).max() < 1e-12
    assert np.linalg.eigvalsh(s.A.toarray()).min() > 0
    # the coupling term annihilates (u, p) with p = trace of u: on constants only the mass terms act,
    # 1' A 1 = k3 |Omega| + k1 |Lambda|
    one = np.ones(s.ndofs)
    length = s.params["nsegments"] and sum(
        np.linalg.norm(s.W[1].coords[a] - s.W[1].coords[b]) for a, b in segment_graph(8, s.params["nsegments"], 0)[1])
    assert abs(one @ (s.A @ one) - (3.0 * 1.0 + 7.0 * np.pi * length)) < 1e-9


def test_circle_averaged_coupling():
    """radius > 0 (src/emi_3d1d.py:63-66): Pi averages over a circle around the curve."""
    n = 12
    ids, edges, xyz = segment_graph(n, 48, 3)
    Pi = circle_average(n, xyz, edges, 0.07)
    assert abs(np.asarray(Pi.sum(axis=1)).ravel() - 1.0).max() < 1e-14 and Pi.data.min() > 0
    assert Pi.nnz / Pi.shape[0] > 8                       # a row touches every tetrahedron its circle crosses
    idx = np.indices((n + 1,) * 3)[::-1]
    x3 = np.stack([idx[a].ravel() / n for a in range(3)], axis=1)
    coef = np.array([0.3, -1.7, 0.9])
    inside = np.all((xyz > 0.08) & (xyz < 0.92), axis=1)
    assert inside.sum() > 5
    # the average of a linear function over a circle is its value at the centre; P1 interpolation is exact for it
    assert abs(Pi @ (x3 @ coef + 0.2) - (xyz @ coef + 0.2))[inside].max() < 1e-13
    # radius -> 0 is the trace coupling
    s0, se = emi3d1d_system(8, gamma=1e2, seed=2), emi3d1d_system(8, gamma=1e2, seed=2, radius=1e-13)
    assert abs(s0.A - se.A).max() < 1e-9
    with pytest.raises(ValueError):
        emi3d1d_system(8, radius=-1.0)
    # the averaged system: symmetric positive definite, denser rows near the curve, and the path converges on it
    s = emi3d1d_system(10, gamma=1e3, radius=0.1)
    assert abs(s.A - s.A.T).max() < 1e-12 and np.linalg.eigvalsh(s.A.toarray()).min() > 0
    assert np.diff(s.A.indptr).max() > 2 * np.diff(s0.A.indptr).max()
    _, amg = datfile.read_input(DAT)
    H = mamg.Hierarchy(s.A, amg, s.interface_dofs)
    for order in ("natural", "multicolor"):
        x, info = Oracle(H.export(), order).pcg(s.b, tolerance=1e-6, relative=2, maxiter=1000)
        assert info["niters"] < 60
        assert np.linalg.norm(s.A @ x - s.b) <= 1.05e-6 * np.linalg.norm(s.b)


def test_sa_amg_hierarchy():
    s = problems.bidomain_system(2, 32, gamma=10.0)
    prm = dict(params.parameters_standard, AMG_type=haznics.SA_AMG, cycle_type=haznics.V_CYCLE,
               smoother=haznics.SMOOTHER_GS, coarse_scaling=haznics.OFF, strong_coupled=0.0,
               max_aggregation=8, coarse_dof=50)
    ex = mamg.Hierarchy(s.A, prm).export()
    assert len(ex["levels"]) >= 3
    for l in range(len(ex["levels"]) - 1):
        L, Lc = ex["levels"][l], ex["levels"][l + 1]
        A = sp.csr_matrix((L["data"], L["indices"], L["indptr"]), shape=(L["n"],) * 2)
        Ac = sp.csr_matrix((Lc["data"], Lc["indices"], Lc["indptr"]), shape=(Lc["n"],) * 2)
        P = sp.csr_matrix((L["P_data"], L["P_indices"], L["P_indptr"]), shape=(L["n"], L["n_aggregates"]))
        assert abs(P.T @ A @ P - Ac).max() <= 1e-12 * abs(Ac).max()
        # P = (I - 0.67 D^-1 A) P_tent
        agg = L["agg"]
        rows = np.flatnonzero(agg >= 0)
        Pt = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=P.shape)
        Dinv = sp.diags(1.0 / A.diagonal())
        keep = sp.diags((agg >= 0).astype(float))
        assert abs(keep @ (Pt - 0.67 * Dinv @ A @ Pt) - P).max() < 1e-13


def test_dump_load_roundtrip(tmp_path):
    s = emi3d1d_system(6, gamma=10.0)
    solver_files.dump_system(s.A, s.b, s.W, str(tmp_path) + "/")
    T = np.load(tmp_path / "A.npy")
    assert T.shape == (s.A.nnz, 3) and T.dtype == np.float64       # COO N x 3 float (src/utils.py:313-316)
    A, b, idofs = solver_files.load_system(str(tmp_path) + "/")
    assert abs(A - s.A).max() == 0 and np.array_equal(b, s.b) and np.array_equal(idofs, s.interface_dofs)
    assert np.array_equal(np.load(tmp_path / "idofs3d.npy"), np.arange(s.W[0].dim()))


@pytest.mark.gpu
def test_sa_cycle_matches_oracle_on_device():
    s = problems.bidomain_system(2, 32, gamma=10.0)
    prm = dict(params.parameters_standard, AMG_type=haznics.SA_AMG, cycle_type=haznics.V_CYCLE,
               smoother=haznics.SMOOTHER_GS, coarse_scaling=haznics.ON, strong_coupled=0.0,
               max_aggregation=8, coarse_dof=50)
    H = mamg.Hierarchy(s.A, prm).to_device(0)
    orc = Oracle(H.export(), "multicolor")
    r = np.random.default_rng(0).standard_normal(s.ndofs)
    z, zo = H.apply(r), orc.apply(r)
    assert np.linalg.norm(z - zo) / np.linalg.norm(zo) < 1e-10


@pytest.mark.gpu
def test_3d1d_file_pipeline_on_device(tmp_path):
    """run_emi_3d1d.sh: dump (emi_3d1d.py -dump 1) -> run_solver_3d1d.py -> solution.txt."""
    s = emi3d1d_system(16, gamma=1e4)
    mdir, odir = str(tmp_path / "data") + "/", str(tmp_path / "out") + "/"
    solver_files.dump_system(s.A, s.b, s.W, mdir)
    niters = solver_files.fenics_metric_solver_xd_1d(DAT, mdir, odir)
    sol = np.loadtxt(odir + "solution.txt")
    assert int(sol[0]) == s.ndofs
    x = sol[1:]
    assert np.linalg.norm(s.A @ x - s.b) <= 1.05e-6 * np.linalg.norm(s.b)
    _, amg = datfile.read_input(DAT)
    H = mamg.Hierarchy(s.A, amg, s.interface_dofs)
    assert H.level_info(0)["max_patch_size"] > 32      # exercises the general (CTA-per-patch) Schwarz kernel
    _, ref = Oracle(H.export(), "multicolor").pcg(s.b, tolerance=1e-6, relative=2, maxiter=1000)
    assert abs(niters - ref["niters"]) <= 1
    n2, wh, dt = solver_files.solve_haznics(s.A, s.b, s.W, s.interface_dofs)
    assert len(wh[0]) == s.W[0].dim() and len(wh[1]) == s.W[1].dim() and n2 > 0
