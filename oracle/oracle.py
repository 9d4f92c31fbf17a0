"""ctypes face of oracle/mamg_oracle.c plus small numpy restatements used to cross-check it."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libmamg_oracle.so")


def build_oracle(force=False):
    src = os.path.join(_HERE, "mamg_oracle.c")
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True,
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    return _LIB


def _load():
    lib = C.CDLL(build_oracle())
    vp, i32, dbl = C.c_void_p, C.c_int, C.c_double
    lib.orc_create.restype = vp
    lib.orc_create.argtypes = [i32, i32, i32, dbl, i32, i32, i32, i32]
    lib.orc_add_level.restype = i32
    lib.orc_add_level.argtypes = [vp, i32, vp, vp, vp, vp, i32, vp, i32, vp, i32, vp, vp, vp, i32]
    lib.orc_set_coarse.argtypes = [vp, vp]
    lib.orc_set_prolongator.argtypes = [vp, i32, vp, vp, vp]
    lib.orc_set_ordering.argtypes = [vp, i32]
    lib.orc_set_cycle.argtypes = [vp, i32]
    lib.orc_set_amli.argtypes = [vp, i32, i32]
    lib.orc_amli_coef.argtypes = [i32, vp]
    lib.orc_set_threads.restype = i32
    lib.orc_set_threads.argtypes = [vp, i32]
    lib.orc_visits.restype = C.c_long
    lib.orc_visits.argtypes = [vp]
    lib.orc_kcycle_margin.restype = dbl
    lib.orc_kcycle_margin.argtypes = [vp, i32]
    lib.orc_destroy.argtypes = [vp]
    lib.orc_apply.argtypes = [vp, vp, vp]
    lib.orc_spmv.argtypes = [vp, i32, vp, vp]
    lib.orc_smooth.argtypes = [vp, i32, vp, vp, i32]
    lib.orc_pcg.restype = i32
    lib.orc_pcg.argtypes = [vp, vp, vp, dbl, i32, i32, i32, vp, vp, vp]
    lib.orc_minres.restype = i32
    lib.orc_minres.argtypes = [vp, vp, vp, dbl, i32, i32, vp]
    lib.orc_gmres.restype = i32
    lib.orc_gmres.argtypes = [vp, vp, vp, dbl, i32, i32, i32, vp]
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class Oracle:
    """CPU cycle / PCG on an exported hierarchy (`Hierarchy.export()` dict or the same dict
    loaded from a golden fixture).  ordering: 'multicolor' (device order) or 'natural'."""

    def __init__(self, hier, ordering="multicolor"):
        self.lib = _load()
        P = hier["params"]
        self.params = P
        self.h = self.lib.orc_create(P["cycle_type"], P["maxit"], P["smoother"], P["relaxation"],
                                     P["presmooth_iter"], P["postsmooth_iter"], P["coarse_scaling"],
                                     P["Schwarz_type"])
        self.lib.orc_set_amli(self.h, int(P.get("amli_degree", 3)), int(P.get("nl_amli_krylov_type", 4)))
        self._keep = []
        for L in hier["levels"]:
            arrs = [np.ascontiguousarray(L["indptr"], np.int32), np.ascontiguousarray(L["indices"], np.int32),
                    np.ascontiguousarray(L["data"], np.float64), np.ascontiguousarray(L["agg"], np.int32),
                    np.ascontiguousarray(L["color"], np.int32), np.ascontiguousarray(L["gs_skip"], np.uint8),
                    np.ascontiguousarray(L["patch_ptr"], np.int32), np.ascontiguousarray(L["patch_dofs"], np.int32),
                    np.ascontiguousarray(L["patch_color"], np.int32)]
            ia, ja, a, agg, color, skip, pptr, pdofs, pcolor = arrs
            npatch = len(pptr) - 1
            lev = self.lib.orc_add_level(self.h, int(L["n"]), _p(ia), _p(ja), _p(a), _p(agg), int(L["n_aggregates"]),
                                         _p(color), int(L["n_colors"]), _p(skip), npatch, _p(pptr), _p(pdofs),
                                         _p(pcolor), int(L["n_patch_colors"]))
            if "P_indptr" in L:
                pi = np.ascontiguousarray(L["P_indptr"], np.int32)
                pj = np.ascontiguousarray(L["P_indices"], np.int32)
                pv = np.ascontiguousarray(L["P_data"], np.float64)
                self.lib.orc_set_prolongator(self.h, lev, _p(pi), _p(pj), _p(pv))
        inv = np.ascontiguousarray(hier["coarse_inv"], np.float64)
        self.lib.orc_set_coarse(self.h, _p(inv))
        self.n = int(hier["levels"][0]["n"])
        self.sizes = [int(L["n"]) for L in hier["levels"]]
        self.set_ordering(ordering)

    def __del__(self):
        if getattr(self, "h", None):
            self.lib.orc_destroy(self.h)
            self.h = None

    def set_ordering(self, ordering):
        self.lib.orc_set_ordering(self.h, {"natural": 0, "multicolor": 1}[ordering])

    def set_threads(self, threads):
        """>1 only takes effect for the multicolour ordering; returns the thread count in use."""
        return int(self.lib.orc_set_threads(self.h, int(threads)))

    def set_cycle(self, cycle_type):
        self.lib.orc_set_cycle(self.h, int(cycle_type))

    def set_amli(self, degree, krylov_type=4):
        """amli_degree and HAZmath's nl_amli_krylov_type (5 = GCG, anything else GCR)."""
        self.lib.orc_set_amli(self.h, int(degree), int(krylov_type))

    def apply(self, r):
        r = np.ascontiguousarray(r, np.float64)
        z = np.empty(self.n)
        self.lib.orc_apply(self.h, _p(r), _p(z))
        return z

    def visits(self):
        return int(self.lib.orc_visits(self.h))

    def kcycle_margin(self, since_creation=False):
        """Nonlinear AMLI: how close the closest K-cycle stopping decision of the last apply (or of every apply so
        far) was to its threshold, relatively; tests use it to make sure rounding cannot flip a decision on the
        inputs they compare."""
        return float(self.lib.orc_kcycle_margin(self.h, int(since_creation)))

    def spmv(self, x, level=0):
        x = np.ascontiguousarray(x, np.float64)
        y = np.empty(self.sizes[level])
        self.lib.orc_spmv(self.h, level, _p(x), _p(y))
        return y

    def smooth(self, b, x, level=0, post=False):
        b = np.ascontiguousarray(b, np.float64)
        x = np.array(x, np.float64, copy=True)
        self.lib.orc_smooth(self.h, level, _p(b), _p(x), int(post))
        return x

    def pcg(self, b, x0=None, tolerance=1e-8, relative=False, maxiter=500):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(self.n) if x0 is None else np.array(x0, np.float64, copy=True)
        res = np.zeros(maxiter + 1)
        al = np.zeros(max(maxiter, 1))
        be = np.zeros(max(maxiter, 1))
        k = self.lib.orc_pcg(self.h, _p(b), _p(x), tolerance, int(relative), maxiter, int(x0 is not None),
                             _p(res), _p(al), _p(be))
        return x, {"niters": k, "residuals": res[:k + 1].tolist(), "alphas": al[:k].tolist(),
                   "betas": be[:k].tolist()}

    def minres(self, b, tolerance=1e-8, relative=False, maxiter=500):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(self.n)
        res = np.zeros(maxiter + 2)
        k = self.lib.orc_minres(self.h, _p(b), _p(x), tolerance, int(relative), maxiter, _p(res))
        return x, {"niters": k, "residuals": res[:k + 1].tolist()}

    def gmres(self, b, tolerance=1e-8, relative=False, maxiter=500, restart=30):
        b = np.ascontiguousarray(b, np.float64)
        x = np.zeros(self.n)
        res = np.zeros(maxiter + 2)
        k = self.lib.orc_gmres(self.h, _p(b), _p(x), tolerance, int(relative), maxiter, int(restart), _p(res))
        return x, {"niters": k, "residuals": res[:k + 1].tolist()}


def amli_coefficients(degree):
    """Coefficients q_0..q_degree of the AMLI polynomial the oracle uses (lambda_max = 2, lambda_min = 1/2)."""
    lib = _load()
    coef = np.zeros(16)
    lib.orc_amli_coef(int(degree), _p(coef))
    return coef[:degree + 1].copy()
