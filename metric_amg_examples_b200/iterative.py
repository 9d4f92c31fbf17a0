"""Krylov objects with the surface of block.iterative (cbc.block), as the reference calls them:

    AAinv = ConjGrad(AA_, precond=BB, tolerance=1E-8, show=4, maxiter=500, callback=cbk)
    xx = AAinv * bb_                                     src/bidomain_2d.py:205-206
    niters = len(AAinv.residuals) - 1; AAinv.eigenvalue_estimates()      :213-216

Two execution modes with identical results:
  fused    A is the matrix the preconditioner was built from and no callback is given: the whole
           loop runs on the device through mamg_pcg (SpMV + dot fused, scalars on the device).
  drop-in  anything else (operator A, per-iteration callback(k=, x=, r=) as in
           src/bidomain_2d.py:180): the loop of cbc.block's precondconjgrad runs here and calls
           A*d and B*r per iteration; every B*r is one mamg_apply on the device.
Stopping rule (cbc.block): sqrt(r.Br) <= tolerance, or <= tolerance*sqrt(r0.Br0) when
relativeconv=True; residuals[0] is the initial sqrt(r.Br).
"""
import numpy as np

from .block import block_base, block_mat, block_mul, block_transpose, block_vec, ii_convert, ReductionOperator
from .precond import _Precond


def _inner(a, b):
    if isinstance(a, block_vec):
        return a.inner(b)
    return float((a * b).sum())


class _Iterative(block_base):
    def __init__(self, A, precond=1.0, tolerance=1e-5, initial_guess=None, iter=None, maxiter=200,
                 name=None, show=1, rprecond=None, nonconvergence_is_fatal=False, retain_guess=False,
                 relativeconv=False, callback=None, restart=30, **kwargs):
        self.A = A
        self.B = precond
        self.tolerance = tolerance
        self.initial_guess = initial_guess
        self.maxiter = iter if iter is not None else maxiter
        self.name = name or type(self).__name__
        self.show = show
        self.nonconvergence_is_fatal = nonconvergence_is_fatal
        self.retain_guess = retain_guess
        self.relativeconv = relativeconv
        self.callback = callback
        self.restart = restart
        self.residuals, self.alphas, self.betas = [], [], []
        self.converged = False
        self.mode = None

    # ---- which preconditioner object / layout do we have? ----
    def _unwrap(self):
        """Returns (Minv, blocked): Minv the _Precond behind B (possibly inside R.T*Minv*R)."""
        B = self.B
        if isinstance(B, _Precond):
            return B, False
        if isinstance(B, block_mul) and len(B.chain) == 3:
            Rt, M, R = B.chain
            if isinstance(M, _Precond) and isinstance(R, ReductionOperator) and isinstance(Rt, block_transpose):
                return M, True
        return None, False

    def _same_matrix(self, Minv, blocked):
        A = self.A
        if A is Minv.A:
            return True
        if blocked and isinstance(A, block_mat):
            mono = getattr(A, "_mono", None)
            if mono is None:
                mono = A._mono = ii_convert(A)
            ref = Minv.A
            if hasattr(ref, "nnz") and mono.shape == ref.shape and mono.nnz == ref.nnz:
                return bool(np.array_equal(mono.indptr, ref.indptr) and np.array_equal(mono.indices, ref.indices)
                            and np.array_equal(mono.data, ref.data))
        return False

    def matvec(self, b):
        Minv, blocked = self._unwrap()
        fused_ok = Minv is not None and self.callback is None and self._same_matrix(Minv, blocked)
        if fused_ok:
            self.mode = "fused"
            x = self._solve_fused(Minv, blocked, b)
        else:
            self.mode = "drop-in"
            x = self._solve_generic(b)
        if not self.converged:
            msg = f"{self.name} did not converge in {self.maxiter} iterations"
            if self.nonconvergence_is_fatal:
                raise RuntimeError(msg)
            if self.show:
                print(msg)
        elif self.show >= 2:
            print(f"{self.name} converged [iter={len(self.residuals) - 1}, "
                  f"rel={self.residuals[-1] / max(self.residuals[0], 1e-300):.2e}]")
        return x

    def __mul__(self, b):
        return self.matvec(b)

    @property
    def iterations(self):
        return len(self.residuals) - 1

    def _threshold(self):
        return self.tolerance * (self.residuals[0] if self.relativeconv else 1.0)

    def _mono_vec(self, v, blocked):
        return ii_convert(v) if blocked and isinstance(v, block_vec) else v

    def _split(self, Minv, x, like):
        if isinstance(like, block_vec):
            offs = np.concatenate([[0], np.cumsum([len(v) for v in like])])
            return block_vec([x[offs[i]:offs[i + 1]] for i in range(len(like))])
        return x

    def eigenvalue_estimates(self):
        """Lanczos tridiagonal from the CG coefficients (cbc.block ConjGrad.eigenvalue_estimates):
        T[k,k] = 1/alpha_k + beta_{k-1}/alpha_{k-1}, T[k,k-1] = sqrt(beta_{k-1})/alpha_{k-1}."""
        n = len(self.alphas)
        if n == 0:
            return np.zeros(0)
        T = np.zeros((n, n))
        for k in range(n):
            T[k, k] = 1.0 / self.alphas[k]
            if k > 0:
                T[k, k] += self.betas[k - 1] / self.alphas[k - 1]
                T[k, k - 1] = T[k - 1, k] = np.sqrt(self.betas[k - 1]) / self.alphas[k - 1]
        return np.sort(np.linalg.eigvalsh(T))


class ConjGrad(_Iterative):
    """block.iterative.ConjGrad (preconditioned CG)."""

    def _solve_fused(self, Minv, blocked, b):
        Minv._ensure_device()
        if blocked and isinstance(b, block_vec):
            # block system: the blocks cross the ABI as they are (mamg_pcg_blocks), no concatenated copy
            x0 = list(self.initial_guess) if self.initial_guess is not None else None
            xs, info = Minv.hierarchy.pcg_blocks(list(b), x0_blocks=x0, tolerance=self.tolerance,
                                                 relative=self.relativeconv, maxiter=self.maxiter)
            self.residuals, self.alphas, self.betas = info["residuals"], info["alphas"], info["betas"]
            if info["breakdown"] and self.show:
                print("ConjGrad breakdown")
            self.converged = self.residuals[-1] <= self._threshold()
            return block_vec(xs)
        bm = self._mono_vec(b, blocked)
        x0 = self._mono_vec(self.initial_guess, blocked) if self.initial_guess is not None else None
        x, info = Minv.hierarchy.pcg(bm, x0=x0, tolerance=self.tolerance, relative=self.relativeconv,
                                     maxiter=self.maxiter)
        self.residuals, self.alphas, self.betas = info["residuals"], info["alphas"], info["betas"]
        if info["breakdown"] and self.show:
            print("ConjGrad breakdown")
        self.converged = self.residuals[-1] <= self._threshold()
        return self._split(Minv, x, b)

    def _solve_generic(self, b):
        A, B = self.A, self.B
        Bm = (lambda r: B * r) if not np.isscalar(B) else (lambda r: B * r)
        x = self.initial_guess.copy() if self.initial_guess is not None else 0.0 * b
        r = b - A * x if self.initial_guess is not None else 1.0 * b
        z = Bm(r)
        d = 1.0 * z
        rz = _inner(r, z)
        self.residuals, self.alphas, self.betas = [np.sqrt(rz)], [], []
        k = 0
        while self.residuals[-1] > self._threshold() and k < self.maxiter:
            q = A * d
            dq = _inner(d, q)
            if dq == 0.0:
                print("ConjGrad breakdown")
                break
            alpha = rz / dq
            x = x + alpha * d
            r = r - alpha * q
            z = Bm(r)
            rz_new = _inner(r, z)
            beta = rz_new / rz
            d = z + beta * d
            rz = rz_new
            k += 1
            self.alphas.append(alpha)
            self.betas.append(beta)
            if rz < 0:
                print("ConjGrad breakdown")
                self.residuals.append(float("nan"))
                break
            self.residuals.append(np.sqrt(rz))
            if self.callback is not None:
                self.callback(k=k, x=x, r=r)
        self.converged = self.residuals[-1] <= self._threshold()
        return x


class MinRes(_Iterative):
    """block.iterative.MinRes (preconditioned MINRES, residual estimate in the B-norm)."""

    def _solve_fused(self, Minv, blocked, b):
        bm = self._mono_vec(b, blocked)
        Minv._ensure_device()
        x, info = Minv.hierarchy.minres(bm, tolerance=self.tolerance, relative=self.relativeconv,
                                        maxiter=self.maxiter)
        self.residuals = info["residuals"]
        self.converged = self.residuals[-1] <= self._threshold()
        return self._split(Minv, x, b)

    def _solve_generic(self, b):
        raise NotImplementedError("MinRes runs fused on the device only (A must be the matrix the "
                                  "preconditioner was built from, no callback)")


class GMRES(_Iterative):
    """Restarted right-preconditioned GMRES (the reference's north_star names GMRES; cbc.block
    ships LGMRES with the same constructor)."""

    def _solve_fused(self, Minv, blocked, b):
        bm = self._mono_vec(b, blocked)
        Minv._ensure_device()
        x, info = Minv.hierarchy.gmres(bm, tolerance=self.tolerance, relative=self.relativeconv,
                                       maxiter=self.maxiter, restart=self.restart)
        self.residuals = info["residuals"]
        self.converged = self.residuals[-1] <= self._threshold()
        return self._split(Minv, x, b)

    def _solve_generic(self, b):
        raise NotImplementedError("GMRES runs fused on the device only")


LGMRES = GMRES
