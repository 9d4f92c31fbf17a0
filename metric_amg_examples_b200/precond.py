"""Drop-in preconditioner objects: the names and call signatures of
block.algebraic.hazmath.{AMG, metricAMG} as the reference uses them.

    Minv = metricAMG(A, W, idofs=interface_dofs, parameters=parameters)   src/utils.py:86
    Minv = metricAMG(A, W, parameters=parameters)                         src/utils.py:88
    Minv = AMG(A, parameters=parameters)                                  src/utils.py:40
    precond = R.T * Minv * R                                              src/utils.py:53
    AAinv = ConjGrad(AA, precond=Minv, ...)                               src/bidomain_2d.py:205

Upstream, Precond.matvec copies the dolfin vector to numpy, calls
haznics.apply_precond(b_np, x_np, precond) and copies back; here matvec hands the vector to
mamg_apply (numpy arrays through a staged copy, torch CUDA tensors zero-copy).  Setup failure
raises RuntimeError like upstream ("AMG levels failed to set up").
"""
import numpy as np

from ._capi import MamgError
from .block import block_base, block_vec
from .hierarchy import Hierarchy, csr_arrays
from .params import default_amg_parameters, default_metric_parameters


class _Precond(block_base):
    def __init__(self, A, parameters, idofs, device=0, part=None):
        self.A = A
        self.parameters = dict(parameters)
        try:
            self.hierarchy = Hierarchy(A, self.parameters, idofs, part=part)
        except MamgError as e:
            raise RuntimeError(str(e)) from e
        self.n = self.hierarchy.n
        self._device = device

    def _ensure_device(self):
        if not self.hierarchy.on_device:
            self.hierarchy.to_device(self._device)

    def matvec(self, b):
        """x = B b: one multigrid cycle; returns a new vector of the same kind as b."""
        self._ensure_device()
        wrapped = isinstance(b, block_vec)
        if wrapped:
            if len(b) != 1:
                raise ValueError("monolithic preconditioner expects a 1-block vector; wrap with "
                                 "R.T * Minv * R for block systems (src/utils.py:53)")
            b = b[0]
        if hasattr(b, "get_local"):  # dolfin GenericVector
            x = b.copy()
            x.set_local(self.hierarchy.apply(np.asarray(b.get_local())))
            return x
        x = self.hierarchy.apply(b)
        return block_vec([x]) if wrapped else x

    def transpmult(self, b):  # the cycle is symmetric by construction
        return self.matvec(b)

    @property
    def T(self):
        return self

    def create_vec(self, dim=1):
        return np.zeros(self.n)

    def down_cast(self):
        return self

    def __str__(self):
        h = self.hierarchy
        return (f"<{type(self).__name__} prec of {self.n} dofs, {h.num_levels} levels, "
                f"setup {h.setup_seconds:.3f}s>")


class AMG(_Precond):
    """block.algebraic.hazmath.AMG(A, parameters=None) (src/utils.py:40)."""

    def __init__(self, A, parameters=None, device=0):
        super().__init__(A, parameters if parameters is not None else default_amg_parameters, None, device)


class metricAMG(_Precond):
    """block.algebraic.hazmath.metricAMG(A, W, idofs=None, parameters=None) (src/utils.py:86,88).
    W is only used for the block sizes (list of objects with .dim() or ints)."""

    def __init__(self, A, W=None, idofs=None, parameters=None, device=0, part=None):
        self.W = W
        if W is not None:
            dims = [w.dim() if hasattr(w, "dim") else int(w) for w in W]
            n = csr_arrays(A)[3]
            if sum(dims) != n:
                raise ValueError(f"block sizes {dims} do not add up to the matrix size {n}")
        if idofs is not None:
            idofs = np.asarray(idofs)
            if idofs.dtype == bool:
                idofs = np.flatnonzero(idofs)
        super().__init__(A, parameters if parameters is not None else default_metric_parameters,
                         idofs, device, part)


class LU(block_base):
    """block.algebraic.petsc.LU(A) as the reference uses it for the `diag` preconditioner
    (src/utils.py:9-12: `xii.block_diag_mat([LU(A[i, i]) ...])`): the action of A^{-1} on one block.

    There is no sparse direct factorisation on the device; the block is solved by CG on the device,
    preconditioned with a plain UA-AMG V-cycle of the same library (mamg_pcg), to a relative residual of
    `tolerance` (1e-12: exact as far as the outer Krylov iteration can tell).
    A solve that does not reach the tolerance raises, it is never returned silently."""

    def __init__(self, A, tolerance=1e-12, maxiter=500, device=0):
        from . import haznics_compat as haznics
        prm = dict(default_amg_parameters, cycle_type=haznics.V_CYCLE)
        self.A = A
        try:
            self.hierarchy = Hierarchy(A, prm, None)
        except MamgError as e:
            raise RuntimeError(str(e)) from e
        self.n = self.hierarchy.n
        self.tolerance, self.maxiter, self._device = tolerance, maxiter, device
        self.iterations = []

    def matvec(self, b):
        if not self.hierarchy.on_device:
            self.hierarchy.to_device(self._device)
        wrapped = isinstance(b, block_vec)
        v = b[0] if wrapped else b
        x, info = self.hierarchy.pcg(v, tolerance=self.tolerance, relative=True, maxiter=self.maxiter)
        res = info["residuals"]
        if res[0] > 0.0 and not res[-1] <= self.tolerance * res[0]:
            raise RuntimeError(f"LU: the block solve stopped at a relative residual of {res[-1] / res[0]:.2e} "
                               f"after {info['niters']} iterations")
        self.iterations.append(info["niters"])
        return block_vec([x]) if wrapped else x

    def transpmult(self, b):   # the blocks are symmetric
        return self.matvec(b)

    @property
    def T(self):
        return self

    def create_vec(self, dim=1):
        return np.zeros(self.n)
