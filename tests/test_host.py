"""CPU tests: C-ABI surface, synthetic systems, host setup, parameter translation.
(No compute call needs a GPU here; `-m "not gpu"` runs this file in well under a minute.)"""
import ctypes as C
import os
import re
import sys

import numpy as np
import pytest
import scipy.sparse as sp

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import _capi, haznics_compat as haznics, params, problems
from oracle import fem_ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported():
    """Every function include/mamg.h declares is exported by libmamg.so and bound in _capi."""
    hdr = open(os.path.join(ROOT, "include", "mamg.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = sorted(set(re.findall(r"\b(mamg_[a-z_0-9]+)\s*\(", hdr)))
    assert len(names) >= 25
    lib = C.CDLL(_capi.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} declared in mamg.h but not exported"
    assert set(names) == set(_capi.SYMBOLS)
    assert b"sm_100a" in _capi.lib.mamg_version()


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    s = problems.bidomain_system(2, 8, gamma=10.0)
    H = mamg.Hierarchy(s.A, params.parameters_metric, s.interface_dofs)
    with pytest.raises(_capi.MamgError, match="no CUDA device|no CPU fallback"):
        H.to_device(0)
    with pytest.raises(_capi.MamgError):
        H.apply(np.ones(s.ndofs))


# nnz / dof counts of SURVEY 8a (closed forms of the structural P1 pattern)
@pytest.mark.parametrize("n,dofs,nnz", [(32, 2178, 29444), (64, 8450, 116228), (128, 33282, 461828),
                                        (256, 132098, 1841156)])
def test_bidomain_2d_counts(n, dofs, nnz):
    s = problems.bidomain_system(2, n, gamma=1e3)
    assert s.ndofs == dofs and s.A.nnz == nnz
    assert len(s.interface_dofs) == dofs // 2 and s.interface_dofs[0] == dofs // 2


def test_emi_counts_and_interface_dofs():
    s = problems.emi_system(2, 64, gamma=1e6)
    assert s.ndofs == 4290
    assert len(s.interface_dofs) == 65           # Omega_1 side only in 2-D (src/emi_2d.py:205)
    s3 = problems.emi_system(3, 8, gamma=1e6)
    assert len(s3.interface_dofs) == 2 * 81      # both sides in 3-D (src/emi_3d.py:134-138)
    assert s3.interface_dofs[81] >= s3.W[0].dim()
    with pytest.raises(_capi.MamgError):
        problems.emi_system(2, 7)


@pytest.mark.parametrize("kind", ["bidomain", "emi"])
@pytest.mark.parametrize("dim,n", [(2, 8), (3, 4), (2, 12), (3, 6)])
def test_assembler_matches_cellwise_fem(kind, dim, n):
    s = (problems.bidomain_system if kind == "bidomain" else problems.emi_system)(dim, n, 2.0, 3.0, 7.0)
    R = getattr(fem_ref, kind)(dim, n, 2.0, 3.0, 7.0)
    assert np.array_equal(s.A.indptr, R.indptr) and np.array_equal(s.A.indices, R.indices)
    assert abs(s.A - R).max() < 1e-13
    assert abs(s.A - s.A.T).max() == 0.0


def test_systems_are_spd():
    for s in (problems.bidomain_system(2, 12, gamma=1e4), problems.emi_system(2, 12, gamma=1e4),
              problems.emi_system(3, 4, gamma=1e2)):
        w = np.linalg.eigvalsh(s.A.toarray())
        assert w.min() > 0


def _P(level):
    agg = level["agg"]
    rows = np.flatnonzero(agg >= 0)
    return sp.csr_matrix((np.ones(len(rows)), (rows, agg[rows])), shape=(level["n"], level["n_aggregates"]))


@pytest.mark.parametrize("prm", ["parameters_metric", "parameters_metric_schwarz", "parameters_standard",
                                 "parameters_standard_schwarz", "default_metric_parameters"])
def test_hierarchy_invariants(prm):
    s = problems.bidomain_system(2, 32, gamma=1e3)
    P = getattr(params, prm)
    metric = "metric" in prm
    H = mamg.Hierarchy(s.A, P, s.interface_dofs if metric else None)
    ex = H.export()
    L = ex["levels"]
    assert L[-1]["n"] <= P["coarse_dof"] or len(L) == P["max_levels"]
    for l in range(len(L) - 1):
        A = sp.csr_matrix((L[l]["data"], L[l]["indices"], L[l]["indptr"]), shape=(L[l]["n"],) * 2)
        Ac = sp.csr_matrix((L[l + 1]["data"], L[l + 1]["indices"], L[l + 1]["indptr"]), shape=(L[l + 1]["n"],) * 2)
        Pm = _P(L[l])
        # Galerkin product
        G = (Pm.T @ A @ Pm).tocsr()
        assert abs(G - Ac).max() <= 1e-12 * abs(Ac).max()
        # aggregates partition the non-isolated rows; isolated rows (identity) are left out
        agg = L[l]["agg"]
        offdiag = abs(A - sp.diags(A.diagonal())).sum(axis=1).A1
        assert np.all((agg < 0) == (offdiag == 0)) or "standard" in prm
        assert np.all(np.bincount(agg[agg >= 0], minlength=L[l]["n_aggregates"]) >= 1)
        # colouring: no two coupled rows that the point smoother touches share a colour; the rows
        # it never touches (Schwarz seeds) all sit in colour 0, which then holds nothing else
        col = L[l]["color"]
        skip = L[l]["gs_skip"].astype(bool) if len(L[l].get("gs_skip", ())) else np.zeros(len(col), bool)
        C_ = A.tocoo()
        m = (C_.row != C_.col) & (C_.data != 0) & ~skip[C_.row] & ~skip[C_.col]
        assert np.all(col[C_.row[m]] != col[C_.col[m]])
        assert col.max() + 1 == L[l]["n_colors"]
        if skip.any():
            assert np.all(col[skip] == 0) and np.all(col[~skip] > 0)
    if metric and P["Schwarz_levels"] > 0:
        l0 = L[0]
        A = sp.csr_matrix((l0["data"], l0["indices"], l0["indptr"]), shape=(l0["n"],) * 2)
        npatch = len(l0["patch_seed"])
        assert npatch == len(s.interface_dofs)
        assert np.array_equal(l0["patch_seed"], s.interface_dofs)
        assert np.all(l0["gs_skip"][s.interface_dofs] == 1) and l0["gs_skip"].sum() == npatch
        # patch-conflict colouring: same colour => patches neither overlap nor touch through A
        Anz = A.copy()
        Anz.eliminate_zeros()  # structural zeros (BC-eliminated entries) carry no coupling
        owner = {}
        for c in range(l0["n_patch_colors"]):
            touched = np.zeros(l0["n"], bool)
            for p in np.flatnonzero(l0["patch_color"] == c):
                d = l0["patch_dofs"][l0["patch_ptr"][p]:l0["patch_ptr"][p + 1]]
                assert len(d) <= P["Schwarz_mmsize"] and s.interface_dofs[p] in d
                nb = np.unique(Anz[d].indices)
                assert not touched[d].any(), "patches of one colour conflict"
                touched[nb] = True
                touched[d] = True
        assert owner == {}
    inv = ex["coarse_inv"]
    Ac = sp.csr_matrix((L[-1]["data"], L[-1]["indices"], L[-1]["indptr"]), shape=(L[-1]["n"],) * 2).toarray()
    assert np.allclose(inv @ Ac, np.eye(L[-1]["n"]), atol=1e-9)


def test_hem_pairs_metric_coupling_first():
    """At large gamma HEM matches (u1_i, u2_i): the metric term is collapsed on level 1."""
    s = problems.bidomain_system(2, 16, gamma=1e8)
    H = mamg.Hierarchy(s.A, params.parameters_metric, s.interface_dofs)
    agg = H.export_level(0)["agg"]
    nv = s.W[0].dim()
    free = agg[:nv] >= 0
    assert np.all(agg[:nv][free] == agg[nv:][free])


def test_parameter_translation():
    p = params.to_struct(params.parameters_metric_schwarz)
    assert p.cycle_type == haznics.W_CYCLE and p.aggregation_type == haznics.HEM
    assert p.Schwarz_maxlvl == 1 and p.Schwarz_levels == 1 and p.coarse_dof == 100
    d = params.to_struct(None)  # src/utils.py:60-82
    assert d.Schwarz_maxlvl == 2 and d.smoother == haznics.SMOOTHER_SGS
    with pytest.warns(UserWarning):
        params.to_struct({"no_such_key": 1})
    for ct in (haznics.AMLI_CYCLE, haznics.NL_AMLI_CYCLE, haznics.ADD_CYCLE):   # the whole cycle_type option space
        assert params.to_struct({"cycle_type": ct}).cycle_type == ct
    assert params.to_struct(None).nl_amli_krylov_type == haznics.SOLVER_VFGMRES and params.to_struct(None).amli_degree == 3
    with pytest.raises(NotImplementedError):
        params.to_struct({"cycle_type": 6})
    with pytest.raises(NotImplementedError):
        params.to_struct({"cycle_type": haznics.ADD_CYCLE, "maxit": 2})
    with pytest.raises(NotImplementedError):
        params.to_struct({"cycle_type": haznics.AMLI_CYCLE, "amli_degree": 16})
    with pytest.raises(NotImplementedError):
        params.to_struct({"aggregation_type": 77})
    with pytest.raises(NotImplementedError):
        params.to_struct({"smoother": 4})
    # 0 = "iterative" upstream: served by the same dense inverses (the limit of that iteration)
    assert params.to_struct({"coarse_solver": 0, "Schwarz_blksolver": 0}).coarse_solver == 0
    with pytest.raises(NotImplementedError):
        params.to_struct({"coarse_solver": 7})


def test_reference_parameter_file_runs_unchanged():
    """src/amg_parameters.py imports `haznics`; our constants module must satisfy it."""
    import sys
    import types
    ref = "/root/reference/src/amg_parameters.py"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present on this box")
    sys.modules["haznics"] = haznics
    try:
        mod = types.ModuleType("ref_amg_parameters")
        exec(compile(open(ref).read(), ref, "exec"), mod.__dict__)
    finally:
        del sys.modules["haznics"]
    for name in ("parameters_standard", "parameters_standard_schwarz", "parameters_metric",
                 "parameters_metric_schwarz"):
        assert getattr(mod, name) == getattr(params, name), name


def test_setup_errors():
    A = sp.eye(4, format="csr")
    bad = sp.csr_matrix(np.array([[0.0, 1.0], [1.0, 0.0]]))
    with pytest.raises(_capi.MamgError, match="diagonal"):
        mamg.Hierarchy(bad, params.parameters_metric)
    with pytest.raises(_capi.MamgError, match="out of range"):
        mamg.Hierarchy(A, params.parameters_metric, idofs=[7])
    H = mamg.Hierarchy(A, params.parameters_metric)  # 4 rows <= coarse_dof: one level
    assert H.num_levels == 1


def test_header_is_plain_c_and_usable_from_c(tmp_path):
    """include/mamg.h must compile as C (the boundary is a C-ABI: no C++ or torch types), and a C
    program can drive the host part of the library through it (setup, queries, errors, destroy)."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "mamg.h"
int main(void) {
  /* 1-D Laplacian, 300 rows: coarsened by HEM down to <= 100 rows */
  enum { N = 300 };
  static int32_t ia[N + 1], ja[3 * N];
  static double a[3 * N];
  int k = 0;
  for (int i = 0; i < N; ++i) {
    ia[i] = k;
    if (i > 0) { ja[k] = i - 1; a[k++] = -1.0; }
    ja[k] = i; a[k++] = 2.0;
    if (i < N - 1) { ja[k] = i + 1; a[k++] = -1.0; }
  }
  ia[N] = k;
  mamg_params p;
  if (mamg_params_default(&p) != 0) return 1;
  mamg_handle h = NULL;
  int32_t idofs[3] = {10, 11, 12};
  if (mamg_setup(&p, N, ia, ja, a, 3, idofs, &h) != 0) { printf("setup: %s\n", mamg_last_error()); return 2; }
  int32_t nl = 0;
  if (mamg_num_levels(h, &nl) != 0 || nl < 2) return 3;
  int64_t info[12];
  if (mamg_level_info(h, 0, info) != 0 || info[0] != N || info[1] != k) return 4;
  if (mamg_level_info(h, nl, info) == 0) return 5;             /* out of range must fail ... */
  if (strlen(mamg_last_error()) == 0) return 6;                /* ... with a message */
  double r[N], z[N];
  for (int i = 0; i < N; ++i) r[i] = 1.0;
  if (mamg_apply(h, r, z, 0) == 0) return 7;                   /* not on a device: no CPU fallback */
  printf("levels %d version %s\n", (int)nl, mamg_version());
  return mamg_destroy(h);
}
''')
    exe = tmp_path / "abi"
    libdir = os.path.dirname(_capi.LIB_PATH)
    cmd = [gcc, "-std=c11", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
           "-L", libdir, "-l:libmamg.so", f"-Wl,-rpath,{libdir}"]
    nccl = os.path.dirname(getattr(_capi, "NCCL_PATH", "") or "")
    if nccl:
        cmd += [f"-Wl,-rpath,{nccl}", f"-Wl,-rpath-link,{nccl}"]
    subprocess.run(cmd, check=True, capture_output=True, text=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""))
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert "sm_100a" in out.stdout


def test_setup_does_not_depend_on_storage_order():
    """Unsorted rows and duplicate entries (a COO-style hand-over) give the hierarchy of the canonical CSR."""
    s = problems.bidomain_system(2, 16, gamma=1e3)
    A = s.A
    ref = mamg.Hierarchy(A, params.parameters_metric_schwarz, s.interface_dofs).export()
    rng = np.random.default_rng(0)
    indptr, indices, data = A.indptr.copy(), A.indices.copy(), A.data.copy()
    for i in range(A.shape[0]):
        sl = slice(indptr[i], indptr[i + 1])
        p = rng.permutation(indptr[i + 1] - indptr[i])
        indices[sl], data[sl] = indices[sl][p], data[sl][p]
    # split every entry into two halves stored apart: duplicates that must be summed (0.5 a + 0.5 a is exact)
    ip2 = 2 * indptr
    idx2, dat2 = np.empty(2 * len(indices), np.int32), np.empty(2 * len(indices))
    for i in range(A.shape[0]):
        k = indptr[i + 1] - indptr[i]
        idx2[ip2[i]:ip2[i] + k] = idx2[ip2[i] + k:ip2[i + 1]] = indices[indptr[i]:indptr[i + 1]]
        dat2[ip2[i]:ip2[i] + k] = dat2[ip2[i] + k:ip2[i + 1]] = 0.5 * data[indptr[i]:indptr[i + 1]]
    for trip in ((indptr, indices, data), (ip2.astype(np.int32), idx2, dat2)):
        ex = mamg.Hierarchy(trip, params.parameters_metric_schwarz, s.interface_dofs).export()
        assert len(ex["levels"]) == len(ref["levels"])
        for L, R in zip(ex["levels"], ref["levels"]):
            for k in ("indptr", "indices", "data", "agg", "color"):
                assert np.array_equal(L[k], R[k]), k


def test_vector_lengths_are_checked_before_the_abi():
    """The C-ABI takes bare pointers; the Python face rejects wrong shapes before any device call."""
    s = problems.bidomain_system(2, 8, gamma=10.0)
    H = mamg.Hierarchy(s.A, params.parameters_metric, s.interface_dofs)
    bad = np.ones(s.ndofs + 1)
    for call in (lambda: H.apply(bad), lambda: H.pcg(bad), lambda: H.pcg(np.ones(s.ndofs), x0=bad),
                 lambda: H.minres(bad), lambda: H.gmres(bad), lambda: H.spmv(bad, 0),
                 lambda: H.smooth(bad, bad, 0), lambda: H.apply(np.ones((s.ndofs, 1))),
                 lambda: H.pcg(np.ones(s.ndofs), maxiter=-1)):
        with pytest.raises(ValueError):
            call()


def test_params_struct_layout_matches_ctypes(tmp_path):
    """struct mamg_params: the header, the ctypes mirror in _capi.py and the stub in INTEGRATION.md list
    the same fields in the same order, and gcc lays them out as ctypes does."""
    import shutil
    import subprocess
    hdr = open(os.path.join(ROOT, "include", "mamg.h")).read()
    body = re.search(r"typedef struct mamg_params \{(.*?)\} mamg_params;", hdr, flags=re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"(int32_t|double)\s+(\w+)(\[\d+\])?;", body)
    names = [f[1] for f in fields]
    assert names == [n for n, _ in _capi.PARAM_FIELDS]
    for (ctype, name, arr), (_, pyt) in zip(fields, _capi.PARAM_FIELDS):
        assert (ctype == "double") == (pyt is C.c_double), name
    stub = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub_names = re.findall(r'"(\w+)"', stub[stub.index("class Params(C.Structure)"):stub.index("lib.mamg_last_error.restype")])
    assert stub_names == names
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "mamg.h"\nint main(void){printf("%zu", sizeof(mamg_params));'
                   + "".join(f'printf(" %zu", offsetof(mamg_params, {n}));' for n in names) + "return 0;}\n")
    exe = tmp_path / "layout"
    subprocess.run([gcc, "-std=c11", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got[0] == C.sizeof(_capi.MamgParams)
    assert got[1:] == [getattr(_capi.MamgParams, n).offset for n in names]


def test_import_hierarchy_round_trip():
    """mamg_import_hierarchy: export -> import -> export is the identity, for UA with Schwarz
    patches, SA with stored prolongators and a partitioned hierarchy; and the golden fixtures (an
    'externally produced' hierarchy as far as the library is concerned) import as they are."""
    import glob
    from metric_amg_examples_b200 import haznics_compat as hz
    cases = [
        (problems.bidomain_system(2, 16, gamma=1e3), params.parameters_metric_schwarz, True, None),
        (problems.emi_system(3, 8, gamma=1e6), params.default_metric_parameters, True, 2),
        (problems.bidomain_system(2, 12, gamma=10.0),
         dict(params.parameters_standard, AMG_type=hz.SA_AMG, cycle_type=hz.V_CYCLE, coarse_dof=40, max_aggregation=8), False, None),
    ]
    for s, prm, metric, nparts in cases:
        part = problems.slab_partition(s, nparts) if nparts else None
        H = mamg.Hierarchy(s.A, prm, s.interface_dofs if metric else None, part=part)
        ex = H.export()
        H2 = mamg.Hierarchy.from_export(ex)
        ex2 = H2.export()
        assert len(ex2["levels"]) == len(ex["levels"]) and H2.nparts == H.nparts
        for L, M in zip(ex["levels"], ex2["levels"]):
            for k in L:
                if isinstance(L[k], np.ndarray):
                    assert np.array_equal(L[k], M[k]), k
                else:
                    assert L[k] == M[k], k
        assert np.array_equal(ex["coarse_inv"], ex2["coarse_inv"])
        # the library recolours when no colouring is handed over, with the same greedy rule
        ex3 = mamg.Hierarchy.from_export(ex, recolor=True).export()
        assert all(np.array_equal(L["color"], M["color"]) for L, M in zip(ex["levels"][:-1], ex3["levels"][:-1]))
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    for path in sorted(glob.glob(os.path.join(here, "golden", "*.npz"))):
        hier = make_golden.unflatten(np.load(path))
        Hg = mamg.Hierarchy.from_export(hier)
        assert Hg.num_levels == len(hier["levels"])


def test_import_hierarchy_rejects_bad_input():
    s = problems.bidomain_system(2, 12, gamma=1e3)
    ex = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, s.interface_dofs).export()
    import copy
    bad = copy.deepcopy(ex)
    bad["levels"][1]["color"][:] = 0          # coupled rows in one colour: a data race on the device
    with pytest.raises(_capi.MamgError, match="share a colour"):
        mamg.Hierarchy.from_export(bad)
    bad = copy.deepcopy(ex)
    bad["levels"][0]["patch_color"][:] = 0    # overlapping patches in one colour
    with pytest.raises(_capi.MamgError, match="one colour"):
        mamg.Hierarchy.from_export(bad)
    bad = copy.deepcopy(ex)
    bad["levels"][0]["agg"][0] = 10 ** 6
    with pytest.raises(_capi.MamgError, match="aggregate id"):
        mamg.Hierarchy.from_export(bad)
    bad = copy.deepcopy(ex)
    bad["levels"] = bad["levels"][:1] + bad["levels"][2:]
    with pytest.raises(_capi.MamgError, match="aggregates"):
        mamg.Hierarchy.from_export(bad)
    # the parameter struct is validated on the C side too (a caller that bypasses params.to_struct)
    prm = params.to_struct(params.parameters_metric)
    prm.cycle_type = 9
    recs = (_capi.MamgLevelArrays * 1)()
    h = C.c_void_p()
    assert _capi.lib.mamg_import_hierarchy(C.byref(prm), 1, recs, None, 1, C.byref(h)) != 0
    assert b"cycle_type" in _capi.lib.mamg_last_error()


@pytest.mark.parametrize("agg", ["VMB", "MIS", "MWM", "HEC", "HEM"])
def test_every_aggregation_type_builds_a_valid_hierarchy(agg):
    """src/amg_parameters.py:16 lists VMB, MIS, MWM, HEC beside the HEM of the metric dicts: each gives aggregates
    that cover every non-isolated row, a Galerkin coarse operator P'AP and a cycle under which the oracle's PCG
    converges (the device path does not depend on how the aggregates were formed)."""
    from oracle import Oracle
    s = problems.bidomain_system(2, 24, gamma=1e2)
    prm = dict(params.parameters_metric, aggregation_type=getattr(haznics, agg), strong_coupled=0.05, max_aggregation=6)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
    ex = H.export()
    assert len(ex["levels"]) >= 3
    for l in range(len(ex["levels"]) - 1):
        L, Lc = ex["levels"][l], ex["levels"][l + 1]
        A = sp.csr_matrix((L["data"], L["indices"], L["indptr"]), shape=(L["n"], L["n"]))
        agg_map = L["agg"]
        ok = agg_map >= 0
        assert set(agg_map[ok]) == set(range(L["n_aggregates"])) and Lc["n"] == L["n_aggregates"] < L["n"]
        P = sp.csr_matrix((np.ones(ok.sum()), (np.flatnonzero(ok), agg_map[ok])), shape=(L["n"], Lc["n"]))
        Ac = sp.csr_matrix((Lc["data"], Lc["indices"], Lc["indptr"]), shape=(Lc["n"], Lc["n"]))
        assert abs(P.T @ A @ P - Ac).max() <= 1e-12 * abs(Ac).max()
        if agg in ("HEM", "MWM"):
            assert np.bincount(agg_map[ok]).max() <= 3          # matchings: pairs (+ a leftover)
    b, xt = s.random_rhs(0)
    x, info = Oracle(ex, "multicolor").pcg(b, tolerance=1e-8, relative=True, maxiter=200)
    assert info["niters"] < 60 and np.linalg.norm(x - xt) < 1e-5 * np.linalg.norm(xt)
