"""GPU parity tests: every device kernel and the whole apply / PCG path against the CPU oracle
on the SAME exported hierarchy (north_star: apply within 1e-10 relative, iteration counts
within +-1, same final tolerance).  All calls go through the C-ABI (ctypes)."""
import numpy as np
import pytest

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import haznics_compat as haznics, params, problems
from oracle import Oracle

pytestmark = pytest.mark.gpu

APPLY_TOL = 1e-10  # north_star: single cycle apply vs the reference apply, fp64 relative


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def make(system, prm, metric=True):
    H = mamg.Hierarchy(system.A, prm, system.interface_dofs if metric else None)
    H.to_device(0)
    return H, Oracle(H.export(), "multicolor")


CASES = {
    "bidomain2d_metric": lambda: (problems.bidomain_system(2, 32, gamma=1e3), params.parameters_metric),
    "bidomain2d_schwarz": lambda: (problems.bidomain_system(2, 32, gamma=1e3), params.parameters_metric_schwarz),
    "bidomain2d_g1e8": lambda: (problems.bidomain_system(2, 48, gamma=1e8), params.parameters_metric_schwarz),
    "emi2d_default": lambda: (problems.emi_system(2, 64, gamma=1e6), params.default_metric_parameters),
    "bidomain3d_schwarz": lambda: (problems.bidomain_system(3, 10, gamma=1e4), params.parameters_metric_schwarz),
    "emi3d_default": lambda: (problems.emi_system(3, 12, gamma=1e6), params.default_metric_parameters),
}


@pytest.fixture(scope="module", params=sorted(CASES))
def case(request):
    system, prm = CASES[request.param]()
    H, orc = make(system, prm)
    return request.param, system, prm, H, orc


def test_spmv_every_level(case):
    _, system, _, H, orc = case
    rng = np.random.default_rng(1)
    for l in range(H.num_levels):
        x = rng.standard_normal(H.level_info(l)["rows"])
        assert rel(H.spmv(x, l), orc.spmv(x, l)) < 1e-14


def test_smoother_every_level(case):
    _, system, _, H, orc = case
    rng = np.random.default_rng(2)
    for l in range(H.num_levels - 1):
        n = H.level_info(l)["rows"]
        b, x = rng.standard_normal(n), rng.standard_normal(n)
        for post in (False, True):
            assert rel(H.smooth(b, x, l, post), orc.smooth(b, x, l, post)) < 1e-12


def test_apply_matches_oracle(case):
    name, system, _, H, orc = case
    rng = np.random.default_rng(3)
    for k in range(3):
        r = rng.standard_normal(system.ndofs)
        z, zo = H.apply(r), orc.apply(r)
        assert rel(z, zo) < APPLY_TOL, name


def test_apply_linear_scaling_and_symmetry(case):
    """B is homogeneous of degree one (coarse scaling keeps alpha invariant under r -> c r) and,
    with its symmetric smoothers, symmetric up to the alpha clipping."""
    _, system, _, H, _ = case
    rng = np.random.default_rng(4)
    u, v = rng.standard_normal(system.ndofs), rng.standard_normal(system.ndofs)
    # power-of-two factor: scaling is exact in fp64, so homogeneity must hold to rounding even for
    # gamma = 1e8 (a generic factor perturbs r at the eps level, which the ill-conditioned patch
    # solves amplify to ~cond*eps)
    assert rel(H.apply(4.0 * u), 4.0 * H.apply(u)) < 1e-13
    assert rel(H.apply(3.5 * u), 3.5 * H.apply(u)) < 1e-7
    Bu, Bv = H.apply(u), H.apply(v)
    assert Bu @ u > 0 and Bv @ v > 0


def test_pcg_matches_oracle(case):
    name, system, prm, H, orc = case
    b, xt = system.random_rhs(0)
    tol = 1e-10 if name.startswith("emi") else 1e-8   # src/emi_2d.py:211 / src/bidomain_2d.py:205
    x, info = H.pcg(b, tolerance=tol, maxiter=500)
    xo, ref = orc.pcg(b, tolerance=tol, maxiter=500)
    assert abs(info["niters"] - ref["niters"]) <= 1
    assert info["residuals"][-1] <= tol and ref["residuals"][-1] <= tol
    k = min(info["niters"], ref["niters"], 5)
    assert np.allclose(info["residuals"][:k + 1], ref["residuals"][:k + 1], rtol=1e-8)
    assert np.allclose(info["alphas"][:k], ref["alphas"][:k], rtol=1e-8)
    assert rel(x, xt) < 1e-6
    # true residual
    assert np.linalg.norm(system.A @ x - b) / np.linalg.norm(b) < 1e-7


def test_pcg_relative_and_initial_guess(case):
    _, system, _, H, orc = case
    b, xt = system.random_rhs(5)
    x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=500)
    assert info["residuals"][-1] <= 1e-8 * info["residuals"][0]
    x2, info2 = H.pcg(b, x0=x, tolerance=1e-8, relative=False, maxiter=500)
    assert info2["niters"] < info["niters"]   # restart from the converged iterate: (almost) nothing left to do
    _, ref = orc.pcg(b, tolerance=1e-8, relative=True, maxiter=500)
    assert abs(info["niters"] - ref["niters"]) <= 1


@pytest.mark.parametrize("cycle", [haznics.V_CYCLE, haznics.W_CYCLE])
@pytest.mark.parametrize("smoother,relax", [(haznics.SMOOTHER_JACOBI, 0.6), (haznics.SMOOTHER_GS, 1.0),
                                            (haznics.SMOOTHER_SGS, 1.0), (haznics.SMOOTHER_SOR, 1.2),
                                            (haznics.SMOOTHER_SSOR, 1.2), (haznics.SMOOTHER_L1DIAG, 1.0)])
def test_option_space(cycle, smoother, relax):
    system = problems.bidomain_system(2, 24, gamma=1e2)
    prm = dict(params.parameters_metric, cycle_type=cycle, smoother=smoother, relaxation=relax,
               presmooth_iter=2, postsmooth_iter=1)
    H, orc = make(system, prm)
    r = np.random.default_rng(6).standard_normal(system.ndofs)
    assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL


@pytest.mark.parametrize("stype", [haznics.SCHWARZ_FORWARD, haznics.SCHWARZ_BACKWARD, haznics.SCHWARZ_SYMMETRIC])
def test_schwarz_types_and_large_patches(stype):
    system = problems.emi_system(2, 32, gamma=1e4)
    prm = dict(params.default_metric_parameters, Schwarz_type=stype, Schwarz_maxlvl=4, Schwarz_mmsize=60)
    H, orc = make(system, prm)
    assert 32 < H.level_info(0)["max_patch_size"] <= 60
    r = np.random.default_rng(7).standard_normal(system.ndofs)
    assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL


def test_standard_amg_vmb():
    system = problems.bidomain_system(2, 32, gamma=1.0)
    for prm in (params.parameters_standard, params.parameters_standard_schwarz):
        H, orc = make(system, prm, metric=False)
        r = np.random.default_rng(8).standard_normal(system.ndofs)
        assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL


def test_coarse_scaling_off_gives_symmetric_operator():
    system = problems.bidomain_system(2, 24, gamma=1e3)
    prm = dict(params.parameters_metric_schwarz, coarse_scaling=haznics.OFF)
    H, _ = make(system, prm)
    rng = np.random.default_rng(9)
    u, v = rng.standard_normal(system.ndofs), rng.standard_normal(system.ndofs)
    a, b = v @ H.apply(u), u @ H.apply(v)
    assert abs(a - b) / abs(a) < 1e-11


@pytest.mark.parametrize("gamma", [1e0, 1e2, 1e4, 1e6, 1e8, 1e10])
def test_gamma_sweep_iterations(gamma):
    """run_bidomain_2d.sh sweeps gamma over 1e0..1e10: iteration counts stay bounded and equal the
    oracle's within one."""
    system = problems.bidomain_system(2, 32, gamma=gamma)
    H, orc = make(system, params.parameters_metric_schwarz)
    b, _ = system.random_rhs(0)
    _, info = H.pcg(b, tolerance=1e-8, maxiter=500)
    _, ref = orc.pcg(b, tolerance=1e-8, maxiter=500)
    assert abs(info["niters"] - ref["niters"]) <= 1
    assert info["niters"] <= 40


def test_torch_device_vectors():
    import torch
    system = problems.bidomain_system(2, 32, gamma=1e3)
    H, orc = make(system, params.parameters_metric_schwarz)
    r = np.random.default_rng(10).standard_normal(system.ndofs)
    rt = torch.from_numpy(r).cuda()
    z = H.apply(rt)
    torch.cuda.synchronize()
    assert z.is_cuda and rel(z.cpu().numpy(), orc.apply(r)) < APPLY_TOL
    b, xt = system.random_rhs(1)
    x, info = H.pcg(torch.from_numpy(b).cuda(), tolerance=1e-8)
    assert rel(x.cpu().numpy(), xt) < 1e-6


def test_larger_properties():
    """Size-independent checks at a size the oracle does not need to finish: PCG reaches the
    tolerance, true residual agrees, B stays positive and homogeneous."""
    system = problems.bidomain_system(3, 40, gamma=1e4)  # 137 842 dofs
    H = mamg.Hierarchy(system.A, params.parameters_metric_schwarz, system.interface_dofs).to_device(0)
    b, xt = system.random_rhs(2)
    x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=200)
    assert info["residuals"][-1] <= 1e-8 * info["residuals"][0]
    assert np.linalg.norm(system.A @ x - b) / np.linalg.norm(b) < 1e-6
    u = np.random.default_rng(11).standard_normal(system.ndofs)
    assert rel(H.apply(-2.0 * u), -2.0 * H.apply(u)) < 1e-12


import glob  # noqa: E402
import importlib.util  # noqa: E402
import os  # noqa: E402

_HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("path", sorted(glob.glob(os.path.join(_HERE, "golden", "*.npz"))))
def test_device_reproduces_golden(path):
    """The committed fixtures (tests/golden/make_golden.py): device apply and PCG history."""
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(_HERE, "golden", "make_golden.py"))
    mg = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mg)
    name = os.path.basename(path)[:-4]
    z = np.load(path)
    system, prm, tol = mg.CASES[name]()
    H = mamg.Hierarchy(system.A, prm, system.interface_dofs).to_device(0)
    assert rel(H.apply(z["r"]), z["z_multicolor"]) < APPLY_TOL
    _, info = H.pcg(z["b"], tolerance=tol, maxiter=500)
    want = z["residuals_multicolor"]
    assert abs(len(info["residuals"]) - len(want)) <= 1
    k = min(len(want), len(info["residuals"]), 6)
    assert np.allclose(info["residuals"][:k], want[:k], rtol=1e-7)


def test_minres_and_gmres_match_oracle():
    system = problems.bidomain_system(2, 32, gamma=1e3)
    H, orc = make(system, params.parameters_metric_schwarz)
    b, xt = system.random_rhs(3)
    x, info = H.minres(b, tolerance=1e-8, relative=True, maxiter=200)
    xo, ref = orc.minres(b, tolerance=1e-8, relative=True, maxiter=200)
    assert abs(info["niters"] - ref["niters"]) <= 1
    assert info["residuals"][-1] <= 1e-8 * info["residuals"][0]
    assert rel(x, xt) < 1e-6
    k = min(info["niters"], ref["niters"], 5)
    assert np.allclose(info["residuals"][:k + 1], ref["residuals"][:k + 1], rtol=1e-7)
    x, info = H.gmres(b, tolerance=1e-8, relative=True, maxiter=200, restart=10)
    xo, ref = orc.gmres(b, tolerance=1e-8, relative=True, maxiter=200, restart=10)
    assert abs(info["niters"] - ref["niters"]) <= 1
    assert np.linalg.norm(system.A @ x - b) <= 1.1e-8 * np.linalg.norm(b)
    assert rel(x, xt) < 1e-6


def test_drop_in_classes_fused_and_callback():
    """The reference's driver lines (src/bidomain_2d.py:192-216, src/emi_2d.py:204-212)."""
    from metric_amg_examples_b200 import utils
    from metric_amg_examples_b200.block import block_vec, split_blocks
    from metric_amg_examples_b200.iterative import ConjGrad, MinRes, GMRES
    s = problems.bidomain_system(2, 32, gamma=1e3)
    b, xt = s.random_rhs(4)
    BB = utils.get_hazmath_metric_precond_mono(s.A, s.W, None, params.parameters_metric_schwarz, s.interface_dofs)
    AAinv = ConjGrad(s.A, precond=BB, tolerance=1e-8, show=0, maxiter=500)
    xx = AAinv * b
    assert AAinv.mode == "fused" and rel(xx, xt) < 1e-6
    niters = len(AAinv.residuals) - 1
    ev = AAinv.eigenvalue_estimates()
    assert len(ev) == niters and ev.min() > 0 and ev.max() / ev.min() < 50
    seen = []
    cb = ConjGrad(s.A, precond=BB, tolerance=1e-8, show=0, maxiter=500,
                  callback=lambda k, x, r: seen.append((k, np.linalg.norm(r))))
    xc = cb * b
    assert cb.mode == "drop-in" and len(seen) == len(cb.residuals) - 1
    assert abs(len(cb.residuals) - len(AAinv.residuals)) <= 1 and rel(xc, xt) < 1e-6
    assert rel((MinRes(s.A, precond=BB, tolerance=1e-9, show=0, maxiter=300) * b), xt) < 1e-6
    assert rel((GMRES(s.A, precond=BB, tolerance=1e-9, show=0, maxiter=300, relativeconv=True) * b), xt) < 1e-6
    # block variant R.T * Minv * R on a 2-block system (src/utils.py:45-53, src/emi_2d.py:207-212)
    e = problems.emi_system(2, 32, gamma=1e4)
    AA = split_blocks(e.A, [w.dim() for w in e.W])
    be, xe = e.random_rhs(5)
    n0 = e.W[0].dim()
    bb = block_vec([be[:n0], be[n0:]])
    Bblk = utils.get_hazmath_metric_precond(AA, e.W, None, interface_dofs=e.interface_dofs)
    inv = ConjGrad(AA, precond=Bblk, tolerance=1e-10, show=0, maxiter=500)
    xb = inv * bb
    assert inv.mode == "fused" and isinstance(xb, block_vec) and len(xb) == 2
    assert rel(np.concatenate(list(xb)), xe) < 1e-6
    zb = Bblk * bb   # one preconditioner application on a block vector
    assert isinstance(zb, block_vec) and len(zb[0]) == n0


# ---- edge cases: degenerate hierarchies and inputs ------------------------------------------------
def test_single_level_hierarchy_is_a_direct_solve():
    """n <= coarse_dof: no coarsening at all, the apply is the dense coarse solve (UMFPACK upstream)."""
    s = problems.bidomain_system(2, 6, gamma=1e3)   # 98 dofs <= coarse_dof 100
    H, orc = make(s, params.parameters_metric_schwarz)
    assert H.num_levels == 1
    b, _ = s.random_rhs(0)
    z = H.apply(b)
    assert rel(z, orc.apply(b)) < APPLY_TOL
    assert rel(z, np.linalg.solve(s.A.toarray(), b)) < 1e-10
    x, info = H.pcg(b, tolerance=1e-10, relative=True, maxiter=20)
    assert info["niters"] == orc.pcg(b, tolerance=1e-10, relative=True, maxiter=20)[1]["niters"] <= 2


def test_zero_right_hand_side():
    """b = 0: zero iterations, x = 0, no NaN from the 0/0 of the relative stopping rule."""
    s = problems.bidomain_system(2, 24, gamma=1e3)
    H, orc = make(s, params.parameters_metric_schwarz)
    b = np.zeros(s.ndofs)
    for relative in (False, True):
        x, info = H.pcg(b, tolerance=1e-8, relative=relative, maxiter=50)
        xo, io = orc.pcg(b, tolerance=1e-8, relative=relative, maxiter=50)
        assert info["niters"] == io["niters"] == 0
        assert np.all(x == 0.0) and np.all(np.isfinite(x))
    assert np.all(H.apply(b) == 0.0)


def test_truncated_hierarchy_and_two_cycles_per_apply():
    """max_levels 2 leaves a 575-row coarsest level (dense inverse far above coarse_dof); maxit 2 runs two
    cycles per apply."""
    s = problems.bidomain_system(2, 24, gamma=1e3)
    prm = dict(params.parameters_metric_schwarz, max_levels=2, maxit=2)
    H, orc = make(s, prm)
    assert H.num_levels == 2 and H.level_info(1)["rows"] > 500
    rng = np.random.default_rng(11)
    r = rng.standard_normal(s.ndofs)
    assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL
    b, _ = s.random_rhs(2)
    x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=100)
    xo, io = orc.pcg(b, tolerance=1e-8, relative=True, maxiter=100)
    assert abs(info["niters"] - io["niters"]) <= 1 and rel(x, xo) < 1e-6


def test_metric_amg_without_interface_dofs():
    """metricAMG(A, W, parameters=...) (src/utils.py:88): no idofs, an empty list means the same; the
    Schwarz seeds then come from an independent set over all dofs."""
    s = problems.bidomain_system(2, 24, gamma=1e3)
    Hn = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, None)
    He = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, np.zeros(0, np.int32))
    assert Hn.level_info(0)["n_patches"] == He.level_info(0)["n_patches"] > 0
    He.to_device(0)
    orc = Oracle(He.export(), "multicolor")
    r = np.random.default_rng(12).standard_normal(s.ndofs)
    assert rel(He.apply(r), orc.apply(r)) < APPLY_TOL


def test_block_diag_precond_matches_exact_block_solves():
    """`-precond diag` (src/emi_2d.py:149,181,209; src/utils.py:9-12): block_diag_mat of exact block solves.
    The device realises each LU(A_ii) as an AMG-preconditioned CG to 1e-12; the outer ConjGrad must then take
    the iterations it takes with a true sparse LU of the blocks (scipy splu, test side only)."""
    import scipy.sparse.linalg as spla
    from metric_amg_examples_b200 import utils
    from metric_amg_examples_b200.block import block_vec, split_blocks
    from metric_amg_examples_b200.iterative import ConjGrad
    e = problems.emi_system(2, 32, gamma=1e2)
    n0 = e.W[0].dim()
    AA = split_blocks(e.A, [w.dim() for w in e.W])
    be, xe = e.random_rhs(7)
    bb = block_vec([be[:n0], be[n0:]])
    BB = utils.get_block_diag_precond(AA, e.W, None)
    inv = ConjGrad(AA, precond=BB, tolerance=1e-10, show=0, maxiter=500)
    xb = inv * bb
    assert inv.mode == "drop-in" and rel(np.concatenate(list(xb)), xe) < 1e-6
    lus = [spla.splu(AA[i, i].tocsc()) for i in range(2)]

    class Exact:
        def __mul__(self, r):
            return block_vec([lus[i].solve(np.asarray(r[i])) for i in range(2)])
    ref = ConjGrad(AA, precond=Exact(), tolerance=1e-10, show=0, maxiter=500)
    xr = ref * bb
    assert abs(len(inv.residuals) - len(ref.residuals)) <= 1
    assert np.allclose(inv.residuals[:5], ref.residuals[:5], rtol=1e-6)
    assert rel(np.concatenate(list(xb)), np.concatenate(list(xr))) < 1e-7
    assert all(k < 60 for op in BB.ops for k in op.iterations)


# ---- the remaining values of cycle_type (src/amg_parameters.py:6,26,49,69): AMLI, nonlinear AMLI, additive ----
FURTHER_CYCLES = {
    "amli3": dict(cycle_type=haznics.AMLI_CYCLE),
    "amli1": dict(cycle_type=haznics.AMLI_CYCLE, amli_degree=1),
    "nlamli_gcr": dict(cycle_type=haznics.NL_AMLI_CYCLE),
    "nlamli_gcg": dict(cycle_type=haznics.NL_AMLI_CYCLE, nl_amli_krylov_type=haznics.SOLVER_GCG),
    "additive": dict(cycle_type=haznics.ADD_CYCLE),
}


def _check_cycle(system, prm, seed, pcg=True):
    H, orc = make(system, prm)
    r = np.random.default_rng(seed).standard_normal(system.ndofs)
    zo = orc.apply(r)
    nl = prm["cycle_type"] == haznics.NL_AMLI_CYCLE
    if nl:   # no K-cycle stopping decision on this input is within rounding of its threshold
        assert orc.kcycle_margin() > 1e-3
    z = H.apply(r)
    assert rel(z, zo) < APPLY_TOL
    assert np.array_equal(H.apply(r), z)
    if pcg:
        b, _ = system.random_rhs(seed)
        x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=300)
        xo, io = orc.pcg(b, tolerance=1e-8, relative=True, maxiter=300)
        assert info["residuals"][-1] <= 1e-8 * info["residuals"][0]
        if not nl or orc.kcycle_margin(since_creation=True) > 1e-6:
            assert abs(info["niters"] - io["niters"]) <= 1 and rel(x, xo) < 1e-6
    return H, orc


@pytest.mark.parametrize("base", ["parameters_metric", "parameters_metric_schwarz"])
@pytest.mark.parametrize("name", sorted(FURTHER_CYCLES))
def test_further_cycle_types_match_oracle(name, base):
    """2-D bidomain at gamma 1e6: the nonlinear AMLI cycle takes its second Krylov step on most levels here."""
    system = problems.bidomain_system(2, 32, gamma=1e6)
    _check_cycle(system, dict(getattr(params, base), **FURTHER_CYCLES[name]), 21)


@pytest.mark.parametrize("name", sorted(FURTHER_CYCLES))
def test_further_cycle_types_3d_emi_and_sa(name):
    """3-D EMI with the reference's default dict (general Schwarz kernel, wide rows) and smoothed aggregation
    (stored prolongators) under the further cycle types."""
    system = problems.emi_system(3, 8, gamma=1e4)
    _check_cycle(system, dict(params.default_metric_parameters, **FURTHER_CYCLES[name]), 22, pcg=False)
    s2 = problems.bidomain_system(2, 24, gamma=1e2)
    prm = dict(params.parameters_metric, AMG_type=haznics.SA_AMG, smoother=haznics.SMOOTHER_GS, **FURTHER_CYCLES[name])
    _check_cycle(s2, prm, 23, pcg=False)


def test_set_cycle_switches_between_all_cycle_types():
    """mamg_set_cycle on an uploaded hierarchy: every cycle type on the same device arrays, ending where it began."""
    system = problems.bidomain_system(2, 24, gamma=1e3)
    H, orc = make(system, dict(params.parameters_metric_schwarz, cycle_type=haznics.V_CYCLE))
    r = np.random.default_rng(24).standard_normal(system.ndofs)
    z_v = H.apply(r)
    for ct in (haznics.AMLI_CYCLE, haznics.W_CYCLE, haznics.NL_AMLI_CYCLE, haznics.ADD_CYCLE, haznics.V_CYCLE):
        H.set_cycle(ct)
        orc.set_cycle(ct)
        assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL, ct
    assert np.array_equal(H.apply(r), z_v)
