"""Property-based host tests (hypothesis): random sparse SPD matrices with random interface sets go
through the same setup as the mesh systems; the structural invariants of the hierarchy and the
convergence of the oracle on it must hold for all of them (general, unstructured input: ragged rows,
isolated rows, duplicate-free random patterns).  CPU only."""
import numpy as np
import scipy.sparse as sp
from hypothesis import HealthCheck, given, settings, strategies as st

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import haznics_compat as haznics, params
from oracle import Oracle


def random_spd(n, density, n_isolated, seed):
    """Weighted graph Laplacian + positive diagonal (an M-matrix like the FEM systems of the path);
    the last `n_isolated` rows are identity rows (Dirichlet rows after symmetric elimination)."""
    rng = np.random.default_rng(seed)
    m = n - n_isolated
    k = max(1, int(density * m))
    i = np.repeat(np.arange(m), k)
    j = rng.integers(0, m, size=i.size)
    keep = i != j
    w = rng.uniform(0.1, 10.0, size=i.size)
    W = sp.coo_matrix((w[keep], (i[keep], j[keep])), shape=(n, n)).tocsr()
    W = W + W.T
    # a ring keeps the active part connected
    ring = sp.coo_matrix((np.ones(m), (np.arange(m), (np.arange(m) + 1) % m)), shape=(n, n)).tocsr()
    W = W + ring + ring.T
    d = np.asarray(W.sum(axis=1)).ravel()
    A = sp.diags(d + rng.uniform(0.01, 1.0, size=n)) - W
    A = A.tolil()
    for r in range(m, n):
        A[r, r] = 1.0
    A = A.tocsr()
    A.sort_indices()
    return A


@settings(max_examples=30, deadline=None, derandomize=True, database=None, suppress_health_check=[HealthCheck.too_slow])
@given(n=st.integers(130, 900), density=st.floats(0.002, 0.02), iso=st.integers(0, 7), frac=st.floats(0.0, 0.5),
       seed=st.integers(0, 2**31 - 1), schwarz=st.booleans(), vmb=st.booleans())
def test_random_spd_hierarchy_invariants_and_convergence(n, density, iso, frac, seed, schwarz, vmb):
    A = random_spd(n, density, iso, seed)
    rng = np.random.default_rng(seed + 1)
    m = n - iso
    idofs = np.sort(rng.choice(m, size=int(frac * m), replace=False)).astype(np.int32) if frac > 0 else None
    prm = dict(params.parameters_metric_schwarz if schwarz else params.parameters_metric)
    prm["cycle_type"] = haznics.V_CYCLE
    if vmb:
        prm.update(aggregation_type=haznics.VMB, strong_coupled=0.05, max_aggregation=20)
    H = mamg.Hierarchy(A, prm, idofs)
    ex = H.export()
    L = ex["levels"]
    # coarsening stops at coarse_dof, at max_levels, or when the aggregation makes no progress (the
    # coarsest level is then still small enough for the dense inverse)
    assert L[0]["n"] == n and L[-1]["n"] <= 8192 and all(L[l + 1]["n"] < L[l]["n"] for l in range(len(L) - 1))
    for l in range(len(L) - 1):
        F, Cc = L[l], L[l + 1]
        Af = sp.csr_matrix((F["data"], F["indices"], F["indptr"]), shape=(F["n"],) * 2)
        Ac = sp.csr_matrix((Cc["data"], Cc["indices"], Cc["indptr"]), shape=(Cc["n"],) * 2)
        agg = F["agg"]
        ok = agg >= 0
        P = sp.csr_matrix((np.ones(ok.sum()), (np.flatnonzero(ok), agg[ok])), shape=(F["n"], Cc["n"]))
        G = (P.T @ Af @ P).tocsr()
        assert abs(G - Ac).max() <= 1e-12 * abs(Af).max()                      # Galerkin product
        assert np.all(np.bincount(agg[ok], minlength=Cc["n"]) >= 1)            # no empty aggregate
        skip = F["gs_skip"].astype(bool) if len(F.get("gs_skip", ())) else np.zeros(F["n"], bool)
        Co = Af.tocoo()
        off = (Co.row != Co.col) & (Co.data != 0) & ~skip[Co.row] & ~skip[Co.col]
        assert np.all(F["color"][Co.row[off]] != F["color"][Co.col[off]])       # valid GS colouring
    L0 = L[0]
    if schwarz and len(L) > 1 and len(L0.get("patch_seed", ())):
        A0 = sp.csr_matrix((L0["data"], L0["indices"], L0["indptr"]), shape=(n, n))
        nz = (A0 != 0).astype(np.int8)
        ptr, dofs, pc = L0["patch_ptr"], L0["patch_dofs"], L0["patch_color"]
        assert np.all(np.diff(ptr) <= prm["Schwarz_mmsize"])
        touched = [set(dofs[ptr[p]:ptr[p + 1]]) for p in range(len(ptr) - 1)]
        for c in range(int(pc.max()) + 1):                                      # conflict-free colours
            members = np.flatnonzero(pc == c)
            seen = np.zeros(n, bool)
            for p in members:
                d = dofs[ptr[p]:ptr[p + 1]]
                reach = np.unique(np.concatenate([d, nz[d].indices]))
                assert not seen[reach].any()
                seen[list(touched[p])] = True
    # the cycle is a convergent preconditioner on every one of these systems
    x_true = rng.standard_normal(n)
    b = A @ x_true
    for order in ("natural", "multicolor"):
        x, info = Oracle(ex, order).pcg(b, tolerance=1e-10, relative=True, maxiter=200)
        assert info["niters"] < 120, (order, info["niters"])
        assert np.linalg.norm(x - x_true) <= 1e-6 * np.linalg.norm(x_true)
