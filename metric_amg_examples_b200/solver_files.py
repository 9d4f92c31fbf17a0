"""The file-driven / all-in-C entry points of the reference's 3D-1D pipeline, on the device:

    dump_system(AA, bb, W, folder)                      src/utils.py:304-333
        A.npy = COO triplets (N x 3 float: row, col, value), b.npy, idofs.npy, idofs3d.npy
    fenics_metric_solver_xd_1d(sfile, mdir, odir)       src/run_solver_3d1d.py:38 (HAZmath C solver)
        reads the .dat input (src/input_metric.dat) and the dumped system, solves, writes
        solution.txt = [size, values...] (read back by src/emi_3d1d.py:146-152)
    fenics_metric_amg_solver_dcsr(A, b, x, idofs)       src/utils.py:119
    solve_haznics(A, b, W, interface_dofs)              src/utils.py:95-127
"""
import os
import time

import numpy as np
import scipy.sparse as sp

from . import datfile
from .iterative import ConjGrad, GMRES, MinRes
from .params import default_metric_parameters
from .precond import metricAMG


def dump_system(AA, bb, W, folder=None):
    folder = "./data/" if folder is None else folder
    os.makedirs(folder, exist_ok=True)
    m = sp.csr_matrix(AA).tocoo()
    assert np.all(np.isfinite(m.data)) and np.all(np.isfinite(bb))
    np.save(os.path.join(folder, "A.npy"), np.c_[m.row, m.col, m.data])
    np.save(os.path.join(folder, "b.npy"), np.asarray(bb, dtype=np.float64))
    n0 = W[0].dim()
    np.save(os.path.join(folder, "idofs3d.npy"), np.arange(n0, dtype=np.int32))
    np.save(os.path.join(folder, "idofs.npy"), np.arange(n0, n0 + W[1].dim(), dtype=np.int32))


def load_system(folder):
    T = np.load(os.path.join(folder, "A.npy"))
    b = np.load(os.path.join(folder, "b.npy"))
    n = len(b)
    A = sp.coo_matrix((T[:, 2], (T[:, 0].astype(np.int64), T[:, 1].astype(np.int64))), shape=(n, n)).tocsr()
    A.sort_indices()
    idofs = np.load(os.path.join(folder, "idofs.npy")).astype(np.int32)
    return A, b, idofs


def _solve(A, b, idofs, solver, amg):
    """Krylov per the .dat settings with the metric-AMG preconditioner; HAZmath's stopping rule
    linear_stop_type 1 = ||r||/||r0|| (src/input_metric.dat:54)."""
    B = metricAMG(A, None, idofs=idofs, parameters=amg)
    H = B.hierarchy
    H.to_device(0)
    if solver["type"] == "cg":
        rel = 2 if solver["stop_type"] == 1 else 1
        x, info = H.pcg(b, tolerance=solver["tol"], relative=rel, maxiter=solver["maxit"])
    elif solver["type"] == "minres":
        x, info = H.minres(b, tolerance=solver["tol"], relative=True, maxiter=solver["maxit"])
    elif solver["type"] == "gmres":
        x, info = H.gmres(b, tolerance=solver["tol"], relative=True, maxiter=solver["maxit"], restart=solver["restart"])
    else:
        raise NotImplementedError(f"linear_itsolver_type {solver['type']}")
    return x, info, B


def fenics_metric_solver_xd_1d(sfile, mdir, odir):
    """Drop-in for haznics.fenics_metric_solver_xd_1d: returns the iteration count."""
    solver, amg = datfile.read_input(sfile)
    A, b, idofs = load_system(mdir)
    t0 = time.time()
    x, info, _ = _solve(A, b, idofs, solver, amg)
    dt = time.time() - t0
    os.makedirs(odir, exist_ok=True)
    np.savetxt(os.path.join(odir, "solution.txt"), np.concatenate([[len(x)], x]))
    print(f"metric AMG {solver['type'].upper()}: {info['niters']} iterations, "
          f"relative residual {info['residuals'][-1] / info['residuals'][0]:.3e}, {dt:.3f} s (setup + solve)")
    return info["niters"]


def fenics_metric_amg_solver_dcsr(A, b, x, idofs=None, parameters=None, tol=1e-6, maxit=1000):
    """Drop-in for haznics.fenics_metric_amg_solver_dcsr(Ahaz, bhaz, xhaz, idofs) (src/utils.py:119):
    solves in place into x and returns the iteration count.  The parameters HAZmath hard-codes in
    that C function are not visible in the reference; the inline defaults of src/utils.py:60-82
    are used unless given."""
    solver = {"type": "cg", "tol": tol, "maxit": maxit, "stop_type": 1, "restart": 30}
    sol, info, _ = _solve(A, b, idofs, solver, parameters or default_metric_parameters)
    x[:] = sol
    return info["niters"]


def solve_haznics(A, b, W, interface_dofs=None):
    """src/utils.py:95-127: (niters, [u0, u1], solve seconds)."""
    dimW = sum(V.dim() for V in W)
    x = np.zeros(dimW)
    t0 = time.time()
    niters = fenics_metric_amg_solver_dcsr(A, b, x, interface_dofs)
    dt = time.time() - t0
    return niters, [x[:W[0].dim()], x[W[0].dim():]], dt
