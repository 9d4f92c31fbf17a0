"""CPU tests of the oracle itself (the reference ships no vectors: SURVEY 8c, parity unpinned):
dense linear-algebra cross-checks, the SURVEY section 6 sanity band, golden fixtures."""
import glob
import importlib.util
import os

import numpy as np
import pytest
import scipy.sparse as sp
import scipy.sparse.linalg as spla

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import haznics_compat as haznics, params, problems
from oracle import Oracle

HERE = os.path.dirname(os.path.abspath(__file__))
_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)


def level_matrix(L):
    return sp.csr_matrix((L["data"], L["indices"], L["indptr"]), shape=(L["n"], L["n"]))


def test_natural_gs_matches_triangular_solve():
    """Forward GS in natural order == x + (D+L)^{-1}(b - A x); backward uses (D+U)."""
    s = problems.bidomain_system(2, 12, gamma=50.0)
    prm = dict(params.parameters_metric, smoother=haznics.SMOOTHER_GS)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    ex = H.export()
    orc = Oracle(ex, "natural")
    A = level_matrix(ex["levels"][0])
    rng = np.random.default_rng(0)
    b, x = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    fwd = x + spla.spsolve_triangular(sp.tril(A).tocsr(), b - A @ x, lower=True)
    bwd = x + spla.spsolve_triangular(sp.triu(A).tocsr(), b - A @ x, lower=False)
    assert np.allclose(orc.smooth(b, x, 0, post=False), fwd, rtol=1e-12, atol=1e-13)
    assert np.allclose(orc.smooth(b, x, 0, post=True), bwd, rtol=1e-12, atol=1e-13)


def test_multicolor_gs_is_gs_in_colour_order():
    s = problems.bidomain_system(2, 12, gamma=50.0)
    prm = dict(params.parameters_metric, smoother=haznics.SMOOTHER_GS)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    ex = H.export()
    L0 = ex["levels"][0]
    A = level_matrix(L0)
    order = np.argsort(L0["color"], kind="stable")
    Ap = A[order][:, order].tocsr()
    rng = np.random.default_rng(1)
    b, x = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    ref = x[order] + spla.spsolve_triangular(sp.tril(Ap).tocsr(), (b - A @ x)[order], lower=True)
    out = Oracle(ex, "multicolor").smooth(b, x, 0, post=False)
    assert np.allclose(out[order], ref, rtol=1e-12, atol=1e-13)


def test_schwarz_patch_solve_is_exact_block_solve():
    s = problems.bidomain_system(2, 10, gamma=1e3)
    prm = dict(params.parameters_metric_schwarz, presmooth_iter=0, Schwarz_type=haznics.SCHWARZ_FORWARD)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    ex = H.export()
    L0 = ex["levels"][0]
    A = level_matrix(L0).toarray()
    rng = np.random.default_rng(2)
    b, x = rng.standard_normal(A.shape[0]), rng.standard_normal(A.shape[0])
    ref = x.copy()
    for p in range(len(L0["patch_seed"])):
        d = L0["patch_dofs"][L0["patch_ptr"][p]:L0["patch_ptr"][p + 1]]
        ref[d] += np.linalg.solve(A[np.ix_(d, d)], (b - A @ ref)[d])
    out = Oracle(ex, "natural").smooth(b, x, 0, post=False)
    assert np.allclose(out, ref, rtol=1e-10, atol=1e-12)
    # the multicolour order visits conflict-free patches: same fixed point structure, different order
    out_mc = Oracle(ex, "multicolor").smooth(b, x, 0, post=False)
    ref = x.copy()
    for c in range(L0["n_patch_colors"]):
        for p in np.flatnonzero(L0["patch_color"] == c):
            d = L0["patch_dofs"][L0["patch_ptr"][p]:L0["patch_ptr"][p + 1]]
            ref[d] += np.linalg.solve(A[np.ix_(d, d)], (b - A @ ref)[d])
    assert np.allclose(out_mc, ref, rtol=1e-10, atol=1e-12)


def test_two_level_cycle_matches_dense_formula():
    """V-cycle with two levels, scaling off: z = S'(S(0,r) + P Ac^{-1} P'(r - A S(0,r)))."""
    s = problems.bidomain_system(2, 8, gamma=10.0)
    prm = dict(params.parameters_metric, cycle_type=haznics.V_CYCLE, coarse_scaling=haznics.OFF,
               max_levels=2, smoother=haznics.SMOOTHER_GS)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    ex = H.export()
    assert len(ex["levels"]) == 2
    L0 = ex["levels"][0]
    A = level_matrix(L0)
    agg = L0["agg"]
    rows = np.flatnonzero(agg >= 0)
    P = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(L0["n"], L0["n_aggregates"]))
    Ac = (P.T @ A @ P).toarray()
    r = np.random.default_rng(3).standard_normal(A.shape[0])
    x = spla.spsolve_triangular(sp.tril(A).tocsr(), r, lower=True)
    x = x + P @ np.linalg.solve(Ac, P.T @ (r - A @ x))
    x = x + spla.spsolve_triangular(sp.triu(A).tocsr(), r - A @ x, lower=False)
    assert np.allclose(Oracle(ex, "natural").apply(r), x, rtol=1e-10, atol=1e-12)


@pytest.mark.parametrize("order", ["natural", "multicolor"])
def test_cycle_symmetric_without_scaling(order):
    s = problems.bidomain_system(2, 16, gamma=1e3)
    prm = dict(params.parameters_metric_schwarz, coarse_scaling=haznics.OFF)
    orc = Oracle(mamg.Hierarchy(s.A, prm, s.interface_dofs).export(), order)
    rng = np.random.default_rng(4)
    u, v = rng.standard_normal(s.ndofs), rng.standard_normal(s.ndofs)
    a, b = v @ orc.apply(u), u @ orc.apply(v)
    assert abs(a - b) / abs(a) < 1e-12
    assert u @ orc.apply(u) > 0


def test_w_cycle_visit_counts():
    s = problems.bidomain_system(2, 32, gamma=1e3)
    H = mamg.Hierarchy(s.A, params.parameters_metric, s.interface_dofs)
    orc = Oracle(H.export(), "natural")
    orc.apply(np.ones(s.ndofs))
    L = H.num_levels
    assert orc.visits() == sum(2 ** l for l in range(L - 1))   # level l visited 2^l times
    orc.set_cycle(haznics.V_CYCLE)
    orc.apply(np.ones(s.ndofs))
    assert orc.visits() == L - 1


# SURVEY section 6 sanity band (throw-away scipy sketch of the survey, NOT reference data):
# bidomain 2-D, random rhs, abs tol 1e-8, natural order.  (n, gamma): (V, W, W+Schwarz)
BAND = {(32, 1e3): (15, 12, 9), (64, 1e6): (22, 9, 8), (128, 1e3): (32, 14, 14)}


@pytest.mark.parametrize("n,gamma", sorted(BAND))
def test_iteration_counts_in_survey_band(n, gamma):
    s = problems.bidomain_system(2, n, gamma=gamma)
    b = np.random.default_rng(0).standard_normal(s.ndofs)
    got = []
    for prm in (dict(params.parameters_metric, cycle_type=haznics.V_CYCLE), params.parameters_metric,
                params.parameters_metric_schwarz):
        orc = Oracle(mamg.Hierarchy(s.A, prm, s.interface_dofs).export(), "natural")
        _, info = orc.pcg(b, tolerance=1e-8, maxiter=500)
        got.append(info["niters"])
        assert info["residuals"][-1] <= 1e-8
    for g, want in zip(got, BAND[(n, gamma)]):
        assert abs(g - want) <= 2, (got, BAND[(n, gamma)])


def test_multicolor_vs_natural_iteration_delta():
    """the multicolour order costs +0..3 iterations over HAZmath's natural order."""
    s = problems.bidomain_system(2, 64, gamma=1e3)
    b = np.random.default_rng(0).standard_normal(s.ndofs)
    ex = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, s.interface_dofs).export()
    it = {o: Oracle(ex, o).pcg(b, tolerance=1e-8)[1]["niters"] for o in ("natural", "multicolor")}
    assert 0 <= it["multicolor"] - it["natural"] <= 3


def test_threaded_oracle_equals_serial():
    s = problems.bidomain_system(2, 24, gamma=1e3)
    ex = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, s.interface_dofs).export()
    r = np.random.default_rng(5).standard_normal(s.ndofs)
    orc = Oracle(ex, "multicolor")
    z1 = orc.apply(r)
    orc.set_threads(4)
    z4 = orc.apply(r)
    assert np.allclose(z1, z4, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_oracle_reproduces_golden(path):
    z = np.load(path)
    hier = make_golden.unflatten(z)
    for order in ("multicolor", "natural"):
        orc = Oracle(hier, order)
        out = orc.apply(z["r"])
        assert np.linalg.norm(out - z[f"z_{order}"]) <= 1e-12 * np.linalg.norm(z[f"z_{order}"])
        _, info = orc.pcg(z["b"], tolerance=float(z["tol"]), maxiter=500)
        assert len(info["residuals"]) == len(z[f"residuals_{order}"])
        assert np.allclose(info["residuals"], z[f"residuals_{order}"], rtol=1e-6)


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))))
def test_setup_reproduces_golden_hierarchy(path, monkeypatch):
    """The host setup is deterministic: re-building the case gives the frozen hierarchy.  The fixtures
    were frozen with the caller's full pattern (MAMG_DROP_ZEROS=0); the default setup stores the same
    hierarchy without its explicit zeros (the diagonal always stays)."""
    name = os.path.basename(path)[:-4]
    z = np.load(path)
    system, prm, _ = make_golden.CASES[name]()
    monkeypatch.setenv("MAMG_DROP_ZEROS", "0")
    ex = mamg.Hierarchy(system.A, prm, system.interface_dofs).export()
    assert len(ex["levels"]) == int(z["nlevels"])
    for l, L in enumerate(ex["levels"]):
        for k in ("indptr", "indices", "agg", "color", "patch_ptr", "patch_dofs", "patch_color"):
            assert np.array_equal(L[k], z[f"L{l}_{k}"]), (l, k)
        assert np.allclose(L["data"], z[f"L{l}_data"], rtol=1e-13, atol=0)
    monkeypatch.delenv("MAMG_DROP_ZEROS")
    lean = mamg.Hierarchy(system.A, prm, system.interface_dofs).export()
    for l, L in enumerate(lean["levels"]):
        ip, ix, dv = z[f"L{l}_indptr"], z[f"L{l}_indices"], z[f"L{l}_data"]
        rows = np.repeat(np.arange(len(ip) - 1), np.diff(ip))
        keep = (dv != 0.0) | (ix == rows)
        assert np.array_equal(L["indices"], ix[keep]) and np.allclose(L["data"], dv[keep], rtol=1e-13, atol=0)
        assert np.array_equal(np.diff(L["indptr"]), np.bincount(rows[keep], minlength=len(ip) - 1))
        for k in ("agg", "color", "patch_ptr", "patch_dofs", "patch_color"):
            assert np.array_equal(L[k], z[f"L{l}_{k}"]), (l, k)


def test_krylov_solvers_against_scipy():
    """Independent pin of the Krylov restatements: scipy's CG / MINRES / GMRES with the oracle's cycle as
    the preconditioner M reach the same solution, and CG with HAZmath's stop rule (||r|| <= tol ||b||,
    relative=2, which is scipy's rule for x0 = 0) takes the same number of iterations (+-1)."""
    import scipy.sparse.linalg as spla
    s = problems.bidomain_system(2, 24, gamma=1e3)
    prm = dict(params.parameters_metric_schwarz, coarse_scaling=haznics.OFF)   # a linear, symmetric B
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    orc = Oracle(H.export(), "natural")
    b, xt = s.random_rhs(5)
    M = spla.LinearOperator(s.A.shape, matvec=orc.apply, dtype=np.float64)
    its = [0]

    def count(_):
        its[0] += 1

    x_sp, info = spla.cg(s.A, b, rtol=1e-9, atol=0.0, M=M, maxiter=200, callback=count)
    assert info == 0
    x, mine = orc.pcg(b, tolerance=1e-9, relative=2, maxiter=200)
    assert abs(mine["niters"] - its[0]) <= 1
    assert np.linalg.norm(x - x_sp) <= 1e-7 * np.linalg.norm(x_sp)
    assert np.linalg.norm(x - xt) <= 1e-6 * np.linalg.norm(xt)
    x_mr, info = spla.minres(s.A, b, M=M, rtol=1e-10, maxiter=300)
    assert info == 0
    xm, _ = orc.minres(b, tolerance=1e-10, relative=True, maxiter=300)
    assert np.linalg.norm(xm - x_mr) <= 1e-6 * np.linalg.norm(x_mr)
    x_gm, info = spla.gmres(s.A, b, M=M, rtol=1e-10, atol=0.0, restart=30, maxiter=20)
    assert info == 0
    xg, _ = orc.gmres(b, tolerance=1e-10, relative=True, maxiter=300, restart=30)
    assert np.linalg.norm(xg - x_gm) <= 1e-6 * np.linalg.norm(x_gm)


def test_dropping_explicit_zeros_is_bit_neutral(monkeypatch):
    """Dropping explicit zeros (the default; MAMG_DROP_ZEROS=0 keeps them) removes the exact zeros of the P1
    pattern from every level: same aggregates, colours and patches, about half the stored entries for
    EMI-3D, and bit-identical cycle outputs and residual histories in both smoother orders."""
    s = problems.emi_system(3, 12, gamma=1e6)
    b, _ = s.random_rhs(0)
    r = np.random.default_rng(0).standard_normal(s.ndofs)
    monkeypatch.setenv("MAMG_DROP_ZEROS", "0")
    full = mamg.Hierarchy(s.A, params.default_metric_parameters, s.interface_dofs).export()
    monkeypatch.delenv("MAMG_DROP_ZEROS", raising=False)
    Hl = mamg.Hierarchy(s.A, params.default_metric_parameters, s.interface_dofs)
    lean = Hl.export()
    assert Hl.level_info(0)["nnz_structural"] == len(full["levels"][0]["data"]) > Hl.level_info(0)["nnz"]
    nnz = lambda ex: sum(len(L["data"]) for L in ex["levels"])
    assert nnz(lean) < 0.6 * nnz(full)
    for F, Z in zip(full["levels"], lean["levels"]):
        assert np.all(Z["data"][Z["indices"] != np.repeat(np.arange(Z["n"]), np.diff(Z["indptr"]))] != 0.0)
        for k in ("agg", "color", "patch_ptr", "patch_dofs", "patch_color"):
            if k in F:
                assert np.array_equal(F[k], Z[k]), k
    for order in ("natural", "multicolor"):
        assert np.array_equal(Oracle(full, order).apply(r), Oracle(lean, order).apply(r))
        h0 = Oracle(full, order).pcg(b, tolerance=1e-10)[1]["residuals"]
        h1 = Oracle(lean, order).pcg(b, tolerance=1e-10)[1]["residuals"]
        assert h0 == h1


# ---- AMLI / nonlinear AMLI / additive cycles (the remaining values of cycle_type, src/amg_parameters.py:6) -------

def _smoother_matrices(orc, level, n, post):
    """The smoother is affine, x' = E x + N b: extract E and N from the oracle's own sweeps."""
    Z, I = np.zeros(n), np.eye(n)
    E = np.column_stack([orc.smooth(Z, I[:, k], level, post) for k in range(n)])
    N = np.column_stack([orc.smooth(I[:, k], Z, level, post) for k in range(n)])
    return E, N


def _dense_hierarchy(ex, orc):
    lv = []
    for l, L in enumerate(ex["levels"]):
        A = level_matrix(L).toarray()
        d = {"A": A}
        if l + 1 < len(ex["levels"]):
            agg = L["agg"]
            rows = np.flatnonzero(agg >= 0)
            d["P"] = sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(L["n"], L["n_aggregates"])).toarray()
            d["pre"] = _smoother_matrices(orc, l, L["n"], False)
            d["post"] = _smoother_matrices(orc, l, L["n"], True)
        lv.append(d)
    return lv


def test_amli_coefficients_are_the_best_uniform_approximation_of_the_reciprocal():
    """q_k (three-term recurrence in the degree, lambda in [1/2, 2]) is the polynomial of best uniform approximation
    to 1/t: the error 1/t - q_k(t) equioscillates at k + 2 points (Chebyshev's criterion), with magnitude
    (3/4) 3^-k, i.e. the factor (sqrt(kappa)-1)/(sqrt(kappa)+1) = 1/3 per degree."""
    from oracle.oracle import amli_coefficients
    t = np.linspace(0.5, 2.0, 300001)
    for deg in range(0, 8):
        q = amli_coefficients(deg)
        err = 1.0 / t - np.polyval(q[::-1], t)
        d = np.diff(err)
        idx = [0] + [i + 1 for i in range(len(d) - 1) if d[i] * d[i + 1] < 0] + [len(t) - 1]
        ext = err[idx]
        assert len(ext) == deg + 2
        assert np.all(ext[:-1] * ext[1:] < 0)
        assert np.allclose(np.abs(ext), 0.75 * 3.0 ** (-deg), rtol=1e-6)


@pytest.mark.parametrize("degree", [0, 1, 2, 3])
def test_amli_cycle_matches_dense_polynomial_formula(degree):
    """Without coarse scaling the AMLI cycle is linear: B_l = post(pre + P q(B_c A_c) B_c P'(I - A pre)) with
    q(t) = sum_i q_i t^i, built here as dense matrices level by level."""
    from oracle.oracle import amli_coefficients
    s = problems.bidomain_system(2, 8, gamma=10.0)
    prm = dict(params.parameters_metric, cycle_type=haznics.AMLI_CYCLE, amli_degree=degree,
               coarse_scaling=haznics.OFF, coarse_dof=10)
    ex = mamg.Hierarchy(s.A, prm, s.interface_dofs).export()
    assert len(ex["levels"]) >= 3
    orc = Oracle(ex, "multicolor")
    lv = _dense_hierarchy(ex, orc)
    q = amli_coefficients(degree)
    B = np.linalg.inv(lv[-1]["A"])
    for d in reversed(lv[:-1]):
        A, P = d["A"], d["P"]
        Ac = P.T @ A @ P
        X = B @ Ac
        Q = sum(q[i] * np.linalg.matrix_power(X, i) for i in range(degree + 1)) @ B
        (Epre, Npre), (Epost, Npost) = d["pre"], d["post"]
        x1 = Npre                                          # pre-smoothing from zero
        x2 = x1 + P @ Q @ P.T @ (np.eye(A.shape[0]) - A @ x1)
        B = Epost @ x2 + Npost
    r = np.random.default_rng(5).standard_normal(s.ndofs)
    z = orc.apply(r)
    assert np.allclose(z, B @ r, rtol=1e-9, atol=1e-11)
    # degree 0 with q_0 = 1.25 is a V-cycle whose coarse corrections are over-relaxed by q_0
    assert orc.visits() == sum((degree + 1) ** l for l in range(len(lv) - 1))


def test_nonlinear_amli_two_level_is_a_v_cycle_and_kcycle_formula():
    """Two levels: the coarse problem is solved directly, so NL-AMLI == V-cycle.  Three levels: the coarse
    correction of level 0 is the two-step Krylov combination of the level-1 cycle (dense restatement)."""
    s = problems.bidomain_system(2, 8, gamma=10.0)
    prm = dict(params.parameters_metric, cycle_type=haznics.NL_AMLI_CYCLE, coarse_scaling=haznics.OFF, max_levels=2)
    ex = mamg.Hierarchy(s.A, prm, s.interface_dofs).export()
    orc = Oracle(ex, "multicolor")
    r = np.random.default_rng(6).standard_normal(s.ndofs)
    z = orc.apply(r)
    orc.set_cycle(haznics.V_CYCLE)
    assert np.array_equal(z, orc.apply(r))

    for krylov in (haznics.SOLVER_GCG, haznics.SOLVER_VFGMRES):
        prm = dict(params.parameters_metric, cycle_type=haznics.NL_AMLI_CYCLE, coarse_scaling=haznics.OFF,
                   max_levels=3, coarse_dof=5, nl_amli_krylov_type=krylov)
        ex = mamg.Hierarchy(s.A, prm, s.interface_dofs).export()
        assert len(ex["levels"]) == 3
        orc = Oracle(ex, "multicolor")
        lv = _dense_hierarchy(ex, orc)
        A0, P0, A1, P1 = lv[0]["A"], lv[0]["P"], lv[1]["A"], lv[1]["P"]
        (E1, N1), (E1p, N1p) = lv[1]["pre"], lv[1]["post"]
        x = N1
        x = x + P1 @ np.linalg.inv(lv[2]["A"]) @ P1.T @ (np.eye(A1.shape[0]) - A1 @ x)
        B1 = E1p @ x + N1p                                  # the level-1 cycle (its coarse level is the last)
        (E0, N0), (E0p, N0p) = lv[0]["pre"], lv[0]["post"]
        x0 = N0 @ r
        b1 = P0.T @ (r - A0 @ x0)
        c1 = B1 @ b1
        v1 = A1 @ c1
        w1 = c1 if krylov == haznics.SOLVER_GCG else v1
        rho1, alpha1 = w1 @ v1, w1 @ b1
        rt = b1 - alpha1 / rho1 * v1
        if rt @ rt < 0.04 * (b1 @ b1):
            e = alpha1 / rho1 * c1
        else:
            c2 = B1 @ rt
            v2 = A1 @ c2
            w2 = c2 if krylov == haznics.SOLVER_GCG else v2
            gamma, alpha2, rho2 = w2 @ v1, w2 @ v2, w2 @ rt
            beta2 = alpha2 - gamma * gamma / rho1
            e = (alpha1 - gamma * rho2 / beta2) / rho1 * c1 + rho2 / beta2 * c2
            # the two-step combination minimises the A-norm (GCG) / residual norm (GCR) over span{c1, c2}
            C = np.column_stack([c1, c2])
            if krylov == haznics.SOLVER_GCG:
                best = C @ np.linalg.solve(C.T @ A1 @ C, C.T @ b1)
            else:
                best = C @ np.linalg.lstsq(A1 @ C, b1, rcond=None)[0]
            assert np.allclose(e, best, rtol=1e-8, atol=1e-10)
        x0 = x0 + P0 @ e
        z_ref = E0p @ x0 + N0p @ r
        assert np.allclose(orc.apply(r), z_ref, rtol=1e-9, atol=1e-11)


def test_additive_cycle_is_the_sum_of_level_corrections_and_symmetric():
    s = problems.bidomain_system(2, 8, gamma=10.0)
    prm = dict(params.parameters_metric, cycle_type=haznics.ADD_CYCLE, coarse_dof=10)
    ex = mamg.Hierarchy(s.A, prm, s.interface_dofs).export()
    orc = Oracle(ex, "multicolor")
    lv = _dense_hierarchy(ex, orc)
    n = s.ndofs
    B = np.zeros((n, n))
    T = np.eye(n)                                           # P_0 ... P_{l-1}
    for d in lv[:-1]:
        (Epre, Npre), (Epost, Npost) = d["pre"], d["post"]
        S = Epost @ Npre + Npost                            # pre- then post-smoothing from zero
        B += T @ S @ T.T
        T = T @ d["P"]
    B += T @ np.linalg.inv(lv[-1]["A"]) @ T.T
    r = np.random.default_rng(7).standard_normal(n)
    assert np.allclose(orc.apply(r), B @ r, rtol=1e-9, atol=1e-11)
    assert np.allclose(B, B.T, rtol=1e-9, atol=1e-11)
    assert np.linalg.eigvalsh(0.5 * (B + B.T)).min() > 0


@pytest.mark.parametrize("cycle", ["AMLI_CYCLE", "NL_AMLI_CYCLE", "ADD_CYCLE"])
def test_further_cycles_precondition_cg(cycle):
    """The further cycle types are usable preconditioners of the reference's solve (scaling on, Schwarz patches):
    AMLI / NL-AMLI need no more iterations than the V-cycle, the additive cycle converges."""
    s = problems.bidomain_system(2, 32, gamma=1e3)
    b = np.random.default_rng(8).standard_normal(s.ndofs)
    its = {}
    for ct in ("V_CYCLE", cycle):
        prm = dict(params.parameters_metric_schwarz, cycle_type=getattr(haznics, ct))
        orc = Oracle(mamg.Hierarchy(s.A, prm, s.interface_dofs).export(), "multicolor")
        x, info = orc.pcg(b, tolerance=1e-8, relative=True, maxiter=200)
        assert np.linalg.norm(b - s.A @ x) <= 2e-8 * np.linalg.norm(b) * 10
        its[ct] = info["niters"]
    if cycle != "ADD_CYCLE":
        assert its[cycle] <= its["V_CYCLE"]
    else:
        assert its[cycle] < 200
