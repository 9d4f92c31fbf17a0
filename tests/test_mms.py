"""Manufactured-solution check of the bidomain discretisation + solve (the reference's own
self-check: H1 errornorm per refinement and its rate, src/bidomain_2d.py:241-255; P1 => rate -> 1).

2-D bidomain with the reference's MMS (src/bidomain_2d.py:21-28): u1 = cos(pi (x + y)),
u2 = sin(pi (x - y)), f_i = -kappa_i Lap u_i + gamma (u_i - u_j); Dirichlet data on x = 0, 1 (tags 1, 2),
Neumann data kappa_i du_i/dn on y = 0, 1 (tags 3, 4).  The linear system solved is the library's
assembled matrix; the load vector is built here with P1 interpolated data (second-order accurate, so
the H1 rate of P1 is not affected); the solve is the metric-AMG PCG of the CPU oracle.
"""
import numpy as np
import scipy.sparse as sp

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems
from oracle import Oracle, fem_ref

K1, K2, GAMMA = 2.0, 3.0, 1e3   # src/bidomain_2d.py:116-117, run_bidomain_2d.sh gamma grid


def exact(x, y):
    pi = np.pi
    u1, u2 = np.cos(pi * (x + y)), np.sin(pi * (x - y))
    g1 = np.stack([-pi * np.sin(pi * (x + y)), -pi * np.sin(pi * (x + y))], axis=-1)
    g2 = np.stack([pi * np.cos(pi * (x - y)), -pi * np.cos(pi * (x - y))], axis=-1)
    f1 = K1 * 2 * pi ** 2 * u1 + GAMMA * (u1 - u2)
    f2 = K2 * 2 * pi ** 2 * u2 + GAMMA * (u2 - u1)
    return u1, u2, g1, g2, f1, f2


def solve(n):
    s = problems.bidomain_system(2, n, K1, K2, GAMMA)
    coords, cells = fem_ref.box_mesh([n, n], [0.0, 0.0], [1.0 / n, 1.0 / n])
    K, M = fem_ref.p1_matrices(coords, cells)
    nv = coords.shape[0]
    x, y = coords[:, 0], coords[:, 1]
    u1, u2, g1, g2, f1, f2 = exact(x, y)
    b = np.concatenate([M @ f1, M @ f2])
    # Neumann data on y = 0 (n = (0,-1)) and y = 1 (n = (0,1)): 1-D P1 mass on the boundary edges
    h = 1.0 / n
    for yb, sign in ((0.0, -1.0), (1.0, 1.0)):
        ids = np.flatnonzero(np.isclose(y, yb))
        ids = ids[np.argsort(x[ids])]
        for k, (gk, kap) in enumerate(((g1, K1), (g2, K2))):
            flux = kap * sign * gk[ids, 1]
            contrib = np.zeros(len(ids))
            contrib[:-1] += h / 6 * (2 * flux[:-1] + flux[1:])
            contrib[1:] += h / 6 * (flux[:-1] + 2 * flux[1:])
            b[k * nv + ids] += contrib
    # Dirichlet lifting with the unconstrained operator, then the boundary values themselves
    Afull = sp.bmat([[K1 * K + GAMMA * M, -GAMMA * M], [-GAMMA * M, K2 * K + GAMMA * M]], format="csr")
    ud = np.zeros(2 * nv)
    dd = s.dirichlet_dofs
    ud[dd] = np.concatenate([u1, u2])[dd]
    b -= Afull @ ud
    b[dd] = ud[dd]
    H = mamg.Hierarchy(s.A, params.parameters_metric_schwarz, s.interface_dofs)
    xh, info = Oracle(H.export(), "multicolor").pcg(b, tolerance=1e-10, maxiter=200)
    assert info["residuals"][-1] <= 1e-10
    # H1 seminorm error by the centroid rule on every triangle
    err2 = 0.0
    p = coords[cells]
    T = np.transpose(p[:, 1:, :] - p[:, :1, :], (0, 2, 1))
    area = np.abs(np.linalg.det(T)) / 2
    Ti = np.linalg.inv(T)
    gl = np.concatenate([-Ti.sum(axis=1, keepdims=True), Ti], axis=1)   # gradients of the barycentrics
    cx, cy = p[:, :, 0].mean(axis=1), p[:, :, 1].mean(axis=1)
    _, _, e1, e2, _, _ = exact(cx, cy)
    for k, ge in enumerate((e1, e2)):
        uh = xh[k * nv:(k + 1) * nv][cells]
        gh = np.einsum("ci,cia->ca", uh, gl)
        err2 += float((area * ((gh - ge) ** 2).sum(axis=1)).sum())
    return np.sqrt(err2), info["niters"]


def test_h1_rate_and_iteration_counts():
    errs, its = [], []
    for n in (8, 16, 32):
        e, k = solve(n)
        errs.append(e)
        its.append(k)
    rates = [np.log(errs[i] / errs[i + 1]) / np.log(2.0) for i in range(2)]
    assert all(0.9 < r < 1.15 for r in rates), (errs, rates)
    assert max(its) <= 20, its     # gamma-robust metric AMG: iteration counts stay small
