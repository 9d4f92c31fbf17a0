"""Synthetic systems of BASELINE.json's configs (no dolfin): what the reference's
get_system() + ii_convert hand to the preconditioner factory.

    bidomain_system(dim, n, ...)  <->  src/bidomain_2d.py:51-99,178-179,192 (2-D) and
                                       src/bidomain_3d.py:59,119,138 (3-D)
    emi_system(dim, n, ...)       <->  src/emi_2d.py:58-128,204-206 and
                                       src/emi_3d.py:67,125,133-139

Each returns a `System` with the monolithic CSR matrix `A` ([W0 dofs; W1 dofs]), the block
sizes `W` (stand-ins for the FunctionSpaces: objects with .dim()), `interface_dofs` exactly as
the reference driver builds them, and helpers for right-hand sides.
"""
import ctypes as C
from dataclasses import dataclass, field

import numpy as np
import scipy.sparse as sp

from .._capi import check, lib, ptr


class Space:
    """Minimal stand-in for dolfin.FunctionSpace: only .dim() is used on the path (src/utils.py:99)."""

    def __init__(self, dim, coords=None):
        self._dim = int(dim)
        self.coords = coords

    def dim(self):
        return self._dim


@dataclass
class System:
    name: str
    A: sp.csr_matrix
    W: list
    interface_dofs: np.ndarray
    dirichlet_dofs: np.ndarray
    gdim: int
    ncell: int
    params: dict = field(default_factory=dict)

    @property
    def ndofs(self):
        return self.A.shape[0]

    def random_rhs(self, seed=0):
        """b = A x_true with x_true ~ N(0,1) (numpy default_rng(seed)); Dirichlet rows consistent."""
        rng = np.random.default_rng(seed)
        x = rng.standard_normal(self.ndofs)
        return self.A @ x, x


def _assemble(fn, dim, n, k1, k2, g):
    # one call: rows are known in closed form and rows * (own stencil + coupling stencil) bounds nnz
    # (np.empty does not touch the pages beyond the actual nnz)
    emi = fn is lib.mamg_assemble_emi
    rows = 2 * (n + 1) ** (dim - 1) * ((n // 2 + 1) if emi else (n + 1))
    per_row = {2: 7, 3: 15}[dim] * (1 if emi else 2) + ({2: 3, 3: 7}[dim] if emi else 0)
    nrows, cap = C.c_int64(), C.c_int64(rows * per_row)
    indptr = np.empty(rows + 1, np.int32)
    indices = np.empty(cap.value, np.int32)
    data = np.empty(cap.value, np.float64)
    check(fn(dim, n, k1, k2, g, C.byref(nrows), C.byref(cap), ptr(indptr), ptr(indices), ptr(data)))
    assert nrows.value == rows
    A = sp.csr_matrix((data[:cap.value], indices[:cap.value], indptr), shape=(rows, rows))
    A.has_sorted_indices = True
    return A


def _grid_coords(shape, h, origin):
    idx = np.indices(shape[::-1])[::-1]  # x fastest
    return np.stack([origin[a] + h[a] * idx[a].ravel() for a in range(len(shape))], axis=1)


def bidomain_system(dim, n, kappa1=2.0, kappa2=3.0, gamma=5.0):
    """Defaults kappa1=2, kappa2=3, gamma=5 as src/bidomain_2d.py:116-118."""
    A = _assemble(lib.mamg_assemble_bidomain, dim, n, kappa1, kappa2, gamma)
    nv = (n + 1) ** dim
    coords = _grid_coords((n + 1,) * dim, (1.0 / n,) * dim, (0.0,) * dim) if nv <= 5_000_000 else None
    W = [Space(nv, coords), Space(nv, coords)]
    # src/bidomain_2d.py:192: every dof of the second field
    idofs = np.arange(nv, 2 * nv, dtype=np.int32)
    shape = (n + 1,) * dim
    idx = np.indices(shape[::-1])[::-1]
    daxis = 0 if dim == 2 else 2
    on_d = ((idx[daxis] == 0) | (idx[daxis] == n)).ravel()
    dd = np.flatnonzero(on_d).astype(np.int32)
    return System(f"bidomain_{dim}d", A, W, idofs, np.concatenate([dd, dd + nv]), dim, n,
                  dict(kappa1=kappa1, kappa2=kappa2, gamma=gamma))


def emi_system(dim, n, kappa1=2.0, kappa2=3.0, gamma=5.0, both_sides=None):
    """EMI on the unit square/cube split at 1/2.  interface_dofs: 2-D = the Omega_1 side only
    (src/emi_2d.py:205-206); 3-D = both sides, the second offset by dim(W0) (src/emi_3d.py:134-138)."""
    A = _assemble(lib.mamg_assemble_emi, dim, n, kappa1, kappa2, gamma)
    half = n // 2
    shape = (n + 1,) * (dim - 1) + (half + 1,)
    nv = int(np.prod(shape))
    h = (1.0 / n,) * dim
    c1 = _grid_coords(shape, h, (0.0,) * (dim - 1) + (0.5,)) if nv <= 5_000_000 else None
    c2 = _grid_coords(shape, h, (0.0,) * dim) if nv <= 5_000_000 else None
    W = [Space(nv, c1), Space(nv, c2)]
    plane = (n + 1) ** (dim - 1)
    i1 = np.arange(0, plane, dtype=np.int32)                      # Omega_1: first plane (y or z = 1/2)
    i2 = np.arange(half * plane, (half + 1) * plane, dtype=np.int32) + nv  # Omega_2: last plane
    if both_sides is None:
        both_sides = dim == 3
    idofs = np.concatenate([i1, i2]) if both_sides else i1
    d1 = np.arange(half * plane, (half + 1) * plane, dtype=np.int32)       # top of Omega_1 (tag 3)
    d2 = np.arange(0, plane, dtype=np.int32) + nv                          # bottom of Omega_2 (tag 6)
    return System(f"emi_{dim}d", A, W, idofs, np.concatenate([d1, d2]), dim, n,
                  dict(kappa1=kappa1, kappa2=kappa2, gamma=gamma))


def slab_partition(system, nparts, axis=None):
    """Owner part of every dof for multi-GPU runs: slabs along one mesh axis, both fields (bidomain) or
    both sides of the interface (EMI) of a vertex column in the same part.

    axis=None: the last axis (z in 3-D, y in 2-D).  For EMI that axis is the interface normal, so the
    slab boundaries keep the interface plane and 3 planes on either side inside one part and no Schwarz
    patch straddles parts -- but then one rank owns every patch.  axis=0 (x-strips) cuts across the
    interface instead: every part owns a strip of it, patches near a cut straddle parts (their updates
    travel through the export lists, as for bidomain)."""
    n, dim = system.ncell, system.gdim
    plane = (n + 1) ** (dim - 1)
    nv = system.W[0].dim()
    last = axis is None or axis == dim - 1
    if not last and not 0 <= axis < dim - 1:
        raise ValueError(f"axis={axis} for a {dim}-d mesh")
    stride = 1 if last else (n + 1) ** axis
    if system.name.startswith("bidomain"):
        k = np.arange(nv) // plane if last else (np.arange(nv) // stride) % (n + 1)
        z = np.tile(k, 2)                                        # index of every dof along the cut axis
        cuts = [round(q * (n + 1) / nparts) for q in range(1, nparts)]
    elif system.name.startswith("emi") and system.name != "emi_3d1d":
        half = n // 2
        if last:
            k = np.arange(nv) // plane
            z = np.concatenate([k + half, k])                    # Omega_1 sits on top of Omega_2
            cuts = []
            for q in range(1, nparts):
                c = round(q * (n + 1) / nparts)
                if abs(c - half) <= 3:
                    c = half + 4
                cuts.append(c)
        else:
            z = np.tile((np.arange(nv) // stride) % (n + 1), 2)   # both halves are (n+1)^(dim-1) x (half+1) boxes
            cuts = [round(q * (n + 1) / nparts) for q in range(1, nparts)]
    else:
        raise NotImplementedError(f"no slab partition for {system.name}")
    return np.searchsorted(np.array(sorted(cuts)), z, side="right").astype(np.int32)


def scalar_p1(dim, ncell, h, cK=1.0, cM=0.0):
    """cK*K + cM*M for P1 on a box mesh (used for mass-matrix right-hand sides and tests)."""
    ncell = np.ascontiguousarray(ncell, np.int32)
    h = np.ascontiguousarray(h, np.float64)
    nrows, nnz = C.c_int64(), C.c_int64()
    check(lib.mamg_assemble_scalar(dim, ptr(ncell), ptr(h), cK, cM, C.byref(nrows), C.byref(nnz), None, None, None))
    indptr = np.empty(nrows.value + 1, np.int32)
    indices = np.empty(nnz.value, np.int32)
    data = np.empty(nnz.value, np.float64)
    cap = C.c_int64(nnz.value)
    check(lib.mamg_assemble_scalar(dim, ptr(ncell), ptr(h), cK, cM, C.byref(nrows), C.byref(cap),
                                   ptr(indptr), ptr(indices), ptr(data)))
    return sp.csr_matrix((data, indices, indptr), shape=(nrows.value, nrows.value))
