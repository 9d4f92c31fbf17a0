"""Oracle parity at the sizes and kernel variants that bench.py actually runs.

The small cases of test_gpu_parity.py never reach the code paths that carry the bytes of the
benchmark: levels of >= 32 k rows (the wide row kernels), a long stack of per-level kernels above
the persistent tail, CUDA-graph replay of a > 1000-node cycle.  Here every BASELINE config is run at
the reference's own refinements (src/bidomain_2d.py:168 n = 32..256, src/emi_2d.py:190 n = 64..512)
and the 3-D systems at sizes whose finest level has > 32 k rows; the CPU oracle finishes each in
seconds.  Tolerances are north_star's: apply 1e-10 relative, iteration counts +-1, first residuals.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import haznics_compat as haznics, params, problems
from oracle import Oracle

pytestmark = pytest.mark.gpu

APPLY_TOL = 1e-10
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def check_against_oracle(system, prm, tol, relative=False, threads=0):
    H = mamg.Hierarchy(system.A, prm, system.interface_dofs)
    H.to_device(0)
    orc = Oracle(H.export(), "multicolor")
    orc.set_threads(threads)
    rng = np.random.default_rng(21)
    r = rng.standard_normal(system.ndofs)
    assert rel(H.apply(r), orc.apply(r)) < APPLY_TOL
    # every kernel class on the biggest levels: SpMV and one smoothing application
    for l in range(min(2, H.num_levels - 1)):
        nl = H.level_info(l)["rows"]
        x, b = rng.standard_normal(nl), rng.standard_normal(nl)
        assert rel(H.spmv(x, l), orc.spmv(x, l)) < 1e-13
        assert rel(H.smooth(b, x, l), orc.smooth(b, x, l)) < 1e-11
        assert rel(H.smooth(b, x, l, post=True), orc.smooth(b, x, l, post=True)) < 1e-11
    b, xt = system.random_rhs(0)
    x, info = H.pcg(b, tolerance=tol, relative=relative, maxiter=500)
    xo, ref = orc.pcg(b, tolerance=tol, relative=relative, maxiter=500)
    assert abs(info["niters"] - ref["niters"]) <= 1, (info["niters"], ref["niters"])
    k = min(info["niters"], ref["niters"], 6)
    assert np.allclose(info["residuals"][:k + 1], ref["residuals"][:k + 1], rtol=1e-8)
    assert rel(x, xo) < 1e-6
    return H, orc, info


@pytest.mark.parametrize("n", [32, 64, 128, 256])
def test_c1_bidomain2d_all_refinements(n):
    """BASELINE configs[0]: bidomain_2d.py -nrefs 4 -gamma 1e3 -precond metric_mono (W-cycle,
    Schwarz on level 0, absolute tolerance 1e-8: src/bidomain_2d.py:168,201-205)."""
    check_against_oracle(problems.bidomain_system(2, n, gamma=1e3), params.parameters_metric_schwarz, 1e-8)


@pytest.mark.parametrize("n", [64, 128, 256, 512])
def test_c2_emi2d_refinements(n):
    """BASELINE configs[1]: emi_2d.py -gamma 1e6 with the inline default parameters
    (src/emi_2d.py:190,207-211; tolerance 1e-10 absolute)."""
    check_against_oracle(problems.emi_system(2, n, gamma=1e6), params.default_metric_parameters, 1e-10)


@pytest.mark.parametrize("cycle", ["V", "W"])
def test_c3_bidomain3d_wide_kernels(cycle):
    """bidomain_3d n=32 (71 874 dofs): level 0 and level 1 have >= 32 k rows, so the wide row kernels
    the benchmark uses, the <24,4> Schwarz fast path and a non-tail stack of levels are compared."""
    prm = dict(params.parameters_metric_schwarz,
               cycle_type=haznics.V_CYCLE if cycle == "V" else haznics.W_CYCLE)
    H, _, _ = check_against_oracle(problems.bidomain_system(3, 32, gamma=1e4), prm, 1e-8, relative=True)
    assert H.level_info(0)["rows"] >= 32768 and H.level_info(1)["rows"] >= 32768


def test_c4_emi3d_general_schwarz_at_scale():
    """emi_3d n=40 (70 602 dofs), both interface sides seeded (src/emi_3d.py:134-138), 2-ring patches
    of up to 100 dofs (general Schwarz kernel)."""
    H, _, _ = check_against_oracle(problems.emi_system(3, 40, gamma=1e6), params.default_metric_parameters, 1e-10)
    assert H.level_info(0)["max_patch_size"] > 32


ENV_CASES = [
    {"MAMG_UNROLL": "1"}, {"MAMG_UNROLL": "2"}, {"MAMG_UNROLL": "4"},
    {"MAMG_LANES": "2"}, {"MAMG_LANES": "4"}, {"MAMG_LANES": "8"}, {"MAMG_LANES": "16"}, {"MAMG_LANES": "32"},
    {"MAMG_SCHWARZ_GENERAL": "1"}, {"MAMG_GRAPH": "0"}, {"MAMG_TAIL_ROWS": "0"}, {"MAMG_DROP_ZEROS": "1"},
    {"MAMG_DROP_ZEROS": "0"}, {"MAMG_SW_DEDUP": "0"}, {"MAMG_ROWS": "csr"}, {"MAMG_ROWS": "sell"}, {"MAMG_NVTX": "1"},
    {"MAMG_STAGE_MIN_KB": "0", "MAMG_STAGE_CHUNK_KB": "4"},   # every upload through the pinned bounce buffers, 4 KB chunks
]


@pytest.mark.parametrize("env", ENV_CASES, ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_env_variants_match_oracle(env, monkeypatch):
    """Every tuning switch selects another kernel variant of the same arithmetic."""
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    check_against_oracle(problems.bidomain_system(3, 20, gamma=1e4), params.parameters_metric_schwarz, 1e-8)
    if "MAMG_SCHWARZ_GENERAL" in env or "MAMG_DROP_ZEROS" in env or "MAMG_ROWS" in env or "MAMG_SW_DEDUP" in env or "MAMG_STAGE_MIN_KB" in env:
        check_against_oracle(problems.emi_system(3, 16, gamma=1e6), params.default_metric_parameters, 1e-10)


def test_properties_at_bench_like_size():
    """bidomain_3d n=64 (549 250 dofs; the oracle still finishes in seconds): apply parity on a hierarchy
    with five levels above 32 k rows, graph replay equals the eager launch sequence bit for bit."""
    system = problems.bidomain_system(3, 64, gamma=1e4)
    prm = dict(params.parameters_metric_schwarz, cycle_type=haznics.V_CYCLE)
    H = mamg.Hierarchy(system.A, prm, system.interface_dofs).to_device(0)
    orc = Oracle(H.export(), "multicolor")
    orc.set_threads(0)
    r = np.random.default_rng(5).standard_normal(system.ndofs)
    z1 = H.apply(r)     # first call captures the graph
    z2 = H.apply(r)     # replay
    assert np.array_equal(z1, z2)
    assert rel(z1, orc.apply(r)) < APPLY_TOL
    b, xt = system.random_rhs(3)
    x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=200)
    _, ref = orc.pcg(b, tolerance=1e-8, relative=True, maxiter=200)
    assert abs(info["niters"] - ref["niters"]) <= 1
    assert np.linalg.norm(system.A @ x - b) <= 1e-6 * np.linalg.norm(b)


@pytest.mark.parametrize("halo", ["1", "0"])
def test_multi_gpu_parity_under_torchrun(halo):
    """tests/dist_check.py (apply/PCG parity with the oracle, all ranks bit-identical) on 2 GPUs, in halo
    mode (partitioned matrices, neighbour-only halo exchange) and in the all-gather mode of round 1."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.join(ROOT, "tests", "dist_check.py")]
    p = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                       env=dict(os.environ, MAMG_HALO=halo))
    assert p.returncode == 0, p.stdout[-4000:]
    assert "FAIL" not in p.stdout


def test_imported_hierarchy_on_device():
    """mamg_import_hierarchy: a hierarchy that comes in through the import door (here: the arrays of a
    golden fixture and of an export) is applied on the device exactly like the one the library built."""
    sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
    import make_golden
    z = np.load(os.path.join(ROOT, "tests", "golden", "bidomain3d_n8_g1e4.npz"))
    hier = make_golden.unflatten(z)
    H = mamg.Hierarchy.from_export(hier).to_device(0)
    assert rel(H.apply(z["r"]), z["z_multicolor"]) < APPLY_TOL
    _, info = H.pcg(z["b"], tolerance=float(z["tol"]), maxiter=500)
    want = z["residuals_multicolor"]
    assert abs(len(info["residuals"]) - len(want)) <= 1
    s = problems.emi_system(3, 16, gamma=1e6)
    H0 = mamg.Hierarchy(s.A, params.default_metric_parameters, s.interface_dofs)
    H1 = mamg.Hierarchy.from_export(H0.export())
    H0.to_device(0)
    H1.to_device(0)
    r = np.random.default_rng(1).standard_normal(s.ndofs)
    assert np.array_equal(H0.apply(r), H1.apply(r))


def test_block_vec_of_device_tensors_is_addressed_by_offsets():
    """R.T * Minv * R and the block ConjGrad on a block_vec of CUDA tensors (mamg_apply_blocks /
    mamg_pcg_blocks): same numbers as the monolithic calls, no concatenated copy on either side."""
    import torch
    from metric_amg_examples_b200 import utils
    from metric_amg_examples_b200.block import block_vec, split_blocks
    from metric_amg_examples_b200.iterative import ConjGrad
    e = problems.emi_system(2, 64, gamma=1e4)
    n0 = e.W[0].dim()
    AA = split_blocks(e.A, [w.dim() for w in e.W])
    Bblk = utils.get_hazmath_metric_precond(AA, e.W, None, interface_dofs=e.interface_dofs)
    Minv = Bblk.chain[1]
    r = np.random.default_rng(2).standard_normal(e.ndofs)
    z_mono = Minv.hierarchy.to_device(0).apply(r)
    rb = block_vec([torch.from_numpy(r[:n0]).cuda(), torch.from_numpy(r[n0:]).cuda()])
    zb = Bblk * rb
    assert isinstance(zb, block_vec) and zb[0].is_cuda and len(zb[0]) == n0
    assert np.array_equal(np.concatenate([v.cpu().numpy() for v in zb]), z_mono)
    zh = Bblk * block_vec([r[:n0], r[n0:]])
    assert np.array_equal(np.concatenate(list(zh)), z_mono)
    b, xt = e.random_rhs(5)
    x_mono, info_mono = Minv.hierarchy.pcg(b, tolerance=1e-10, maxiter=500)
    inv = ConjGrad(AA, precond=Bblk, tolerance=1e-10, show=0, maxiter=500)
    xb = inv * block_vec([torch.from_numpy(b[:n0]).cuda(), torch.from_numpy(b[n0:]).cuda()])
    assert inv.mode == "fused" and xb[0].is_cuda
    assert np.array_equal(np.concatenate([v.cpu().numpy() for v in xb]), x_mono)
    assert inv.residuals == info_mono["residuals"]


def test_static_race_check_of_the_device_layout(monkeypatch):
    """compute-sanitizer is closed on this GPU pool; mamg_race_check verifies the data-race property the
    coloured smoothers rest on, on the arrays the kernels stream: no two rows of one Gauss-Seidel colour
    launch couple, no patch of a colour reads or writes a dof another patch of that colour writes.  A
    deliberately broken colouring (imported without validation) must be detected."""
    for system, prm in ((problems.bidomain_system(3, 24, gamma=1e4), params.parameters_metric_schwarz),
                        (problems.emi_system(3, 24, gamma=1e6), params.default_metric_parameters),
                        (problems.bidomain_system(2, 128, gamma=1e3), params.parameters_metric_schwarz)):
        H = mamg.Hierarchy(system.A, prm, system.interface_dofs).to_device(0)
        gs, pt, checked = H.race_check()
        assert (gs, pt) == (0, 0) and checked > 20
    s = problems.bidomain_system(2, 24, gamma=1e3)
    ex = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, s.interface_dofs).export()
    ex["levels"][1]["color"][:] = np.minimum(ex["levels"][1]["color"], 1)      # coupled rows share colour 1
    ex["levels"][0]["patch_color"][:] = ex["levels"][0]["patch_color"] // 2     # overlapping patches share colours
    ex["levels"][0]["n_patch_colors"] = int(ex["levels"][0]["patch_color"].max()) + 1
    monkeypatch.setenv("MAMG_IMPORT_NOCHECK", "1")
    Hbad = mamg.Hierarchy.from_export(ex).to_device(0)
    gs, pt, _ = Hbad.race_check()
    assert gs > 0 and pt > 0
