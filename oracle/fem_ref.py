"""Cell-by-cell P1 assembly in numpy/scipy: an independent restatement of the systems the
reference assembles with dolfin/FEniCS_ii, used to check the library's stencil assembler on
small meshes.  TEST INFRASTRUCTURE.

Follows src/bidomain_2d.py:64-68,93-97 (blocks, symmetric Dirichlet), src/emi_2d.py:83-94,
104-108 (trace coupling, Dirichlet tags), src/utils.py:149-260 (meshes and tags).
"""
import itertools
import math

import numpy as np
import scipy.sparse as sp


def box_mesh(ncell, origin, h):
    """Vertices (lexicographic, x fastest) and Kuhn simplices of a box with ncell cells per axis."""
    d = len(ncell)
    shape = [n + 1 for n in ncell]
    grids = np.meshgrid(*[np.arange(s) for s in shape], indexing="ij")
    # flatten with x fastest
    idx = np.stack([g.ravel(order="F") for g in grids], axis=1)
    coords = np.array(origin)[None, :] + idx * np.array(h)[None, :]
    strides = np.cumprod([1] + shape[:-1])
    cells = []
    cgrids = np.meshgrid(*[np.arange(n) for n in ncell], indexing="ij")
    cidx = np.stack([g.ravel(order="F") for g in cgrids], axis=1)
    for perm in itertools.permutations(range(d)):
        verts = [np.zeros(d, int)]
        for a in perm:
            v = verts[-1].copy()
            v[a] += 1
            verts.append(v)
        cell = np.stack([((cidx + v[None, :]) * strides[None, :]).sum(axis=1) for v in verts], axis=1)
        cells.append(cell)
    return coords, np.concatenate(cells, axis=0)


def p1_matrices(coords, cells):
    """Stiffness K and mass M (scipy CSR, duplicates summed, structural zeros kept)."""
    d = cells.shape[1] - 1
    nv = coords.shape[0]
    p = coords[cells]  # (nc, d+1, gdim)
    T = np.transpose(p[:, 1:, :] - p[:, :1, :], (0, 2, 1))  # columns = edges
    if T.shape[1] != d:  # embedded simplex (not needed here)
        raise ValueError("gdim must equal the simplex dimension")
    det = np.linalg.det(T)
    vol = np.abs(det) / math.factorial(d)
    Ti = np.linalg.inv(T)  # rows = gradients of lambda_1..d
    g = np.concatenate([-Ti.sum(axis=1, keepdims=True), Ti], axis=1)  # (nc, d+1, d)
    Ke = vol[:, None, None] * np.einsum("cia,cja->cij", g, g)
    Me = vol[:, None, None] / ((d + 1) * (d + 2)) * (np.ones((d + 1, d + 1)) + np.eye(d + 1))[None]
    rows = np.repeat(cells, d + 1, axis=1).ravel()
    cols = np.tile(cells, (1, d + 1)).ravel()
    K = sp.coo_matrix((Ke.ravel(), (rows, cols)), shape=(nv, nv)).tocsr()
    M = sp.coo_matrix((Me.ravel(), (rows, cols)), shape=(nv, nv)).tocsr()
    K.sort_indices()
    M.sort_indices()
    return K, M


def _apply_dirichlet(A, dofs):
    """Symmetric elimination keeping the sparsity: rows and columns zeroed, unit diagonal."""
    A = A.tocsr().copy()
    n = A.shape[0]
    mask = np.zeros(n, bool)
    mask[dofs] = True
    rows = np.repeat(np.arange(n), np.diff(A.indptr))
    kill = mask[rows] | mask[A.indices]
    A.data[kill] = 0.0
    diag = (rows == A.indices) & mask[rows]
    A.data[diag] = 1.0
    return A


def _union_pattern(*mats):
    """Sum that keeps explicit zeros (scipy's + drops nothing but may not keep zero entries
    of a single operand; go through COO to be safe)."""
    rows, cols, vals = [], [], []
    for m in mats:
        c = m.tocoo()
        rows.append(c.row)
        cols.append(c.col)
        vals.append(c.data)
    n = mats[0].shape
    A = sp.coo_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=n).tocsr()
    A.sort_indices()
    return A


def bidomain(dim, n, k1, k2, g):
    coords, cells = box_mesh([n] * dim, [0.0] * dim, [1.0 / n] * dim)
    K, M = p1_matrices(coords, cells)
    A00 = _union_pattern(k1 * K, g * M)
    A11 = _union_pattern(k2 * K, g * M)
    A01 = _union_pattern(0.0 * K, -g * M)  # full P1 pattern
    A = sp.bmat([[A00, A01], [A01, A11]], format="coo")
    # bmat drops nothing; convert keeping explicit zeros
    A = sp.csr_matrix((A.data, (A.row, A.col)), shape=A.shape)
    A.sort_indices()
    axis = 0 if dim == 2 else 2
    on = np.isclose(coords[:, axis], 0.0) | np.isclose(coords[:, axis], 1.0)
    dd = np.flatnonzero(on)
    nv = coords.shape[0]
    return _apply_dirichlet(A, np.concatenate([dd, dd + nv]))


def emi(dim, n, k1, k2, g):
    half = n // 2
    h = [1.0 / n] * dim
    nc = [n] * (dim - 1) + [half]
    o1 = [0.0] * (dim - 1) + [0.5]
    o2 = [0.0] * dim
    c1, cells1 = box_mesh(nc, o1, h)
    c2, cells2 = box_mesh(nc, o2, h)
    K1, _ = p1_matrices(c1, cells1)
    K2, _ = p1_matrices(c2, cells2)
    nv = c1.shape[0]
    plane = (n + 1) ** (dim - 1)
    # interface mesh = (dim-1)-box with the same Kuhn split; its vertices are the first `plane`
    # vertices of Omega_1 (last index 0) and the last `plane` vertices of Omega_2
    cg, cellsg = box_mesh([n] * (dim - 1), [0.0] * (dim - 1), h[:-1])
    _, Mg = p1_matrices(cg, cellsg)
    i1 = np.arange(plane)
    i2 = np.arange(half * plane, (half + 1) * plane)
    T1 = sp.csr_matrix((np.ones(plane), (np.arange(plane), i1)), shape=(plane, nv))
    T2 = sp.csr_matrix((np.ones(plane), (np.arange(plane), i2)), shape=(plane, nv))
    A00 = _union_pattern(k1 * K1, g * (T1.T @ Mg @ T1))
    A11 = _union_pattern(k2 * K2, g * (T2.T @ Mg @ T2))
    A01 = -g * (T1.T @ Mg @ T2)
    A = sp.bmat([[A00, A01], [A01.T, A11]], format="coo")
    A = sp.csr_matrix((A.data, (A.row, A.col)), shape=A.shape)
    A.sort_indices()
    d1 = np.arange(half * plane, (half + 1) * plane)   # top of Omega_1 (tag 3)
    d2 = np.arange(plane) + nv                          # bottom of Omega_2 (tag 6)
    return _apply_dirichlet(A, np.concatenate([d1, d2]))
