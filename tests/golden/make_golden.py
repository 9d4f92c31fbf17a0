"""Generates the golden fixtures in this directory (run from the repo root):

    python tests/golden/make_golden.py

The reference ships no vectors for this path (parity unpinned), and its Python modules
cannot be imported here (dolfin / cbc.block / haznics are not installed), so the fixtures freeze the
outputs of the repo's own CPU oracle on small seeded cases: the exported hierarchy, an input
vector, the cycle output in both smoother orders and the PCG residual history.  They guard both
the oracle and the device path against silent changes.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import metric_amg_examples_b200 as mamg  # noqa: E402
from metric_amg_examples_b200 import params, problems  # noqa: E402
from oracle import Oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

CASES = {
    # BASELINE configs[0] at its first refinement: bidomain_2d n=32, gamma=1e3, parameters_metric_schwarz
    "bidomain2d_n32_g1e3": lambda: (problems.bidomain_system(2, 32, gamma=1e3), params.parameters_metric_schwarz, 1e-8),
    # BASELINE configs[1] at its first refinement: emi_2d n=64, gamma=1e6, inline defaults (maxlvl 2)
    "emi2d_n64_g1e6": lambda: (problems.emi_system(2, 64, gamma=1e6), params.default_metric_parameters, 1e-10),
    "bidomain3d_n8_g1e4": lambda: (problems.bidomain_system(3, 8, gamma=1e4), params.parameters_metric_schwarz, 1e-8),
    "emi3d_n8_g1e6": lambda: (problems.emi_system(3, 8, gamma=1e6), params.default_metric_parameters, 1e-10),
}


def flatten(hier):
    out = {"nlevels": np.array(len(hier["levels"])), "coarse_inv": hier["coarse_inv"]}
    for k, v in hier["params"].items():
        out[f"param_{k}"] = np.array(v)
    for l, L in enumerate(hier["levels"]):
        for k, v in L.items():
            out[f"L{l}_{k}"] = np.asarray(v)
    return out


def unflatten(z):
    n = int(z["nlevels"])
    levels = []
    for l in range(n):
        pre = f"L{l}_"
        L = {k[len(pre):]: z[k] for k in z.files if k.startswith(pre)}
        for k in ("n", "n_aggregates", "n_colors", "n_patch_colors"):
            L[k] = int(L[k])
        levels.append(L)
    prm = {k[len("param_"):]: z[k].item() for k in z.files if k.startswith("param_")}
    return {"levels": levels, "coarse_inv": z["coarse_inv"], "params": prm}


def main():
    for name, mk in CASES.items():
        system, prm, tol = mk()
        H = mamg.Hierarchy(system.A, prm, system.interface_dofs)
        hier = H.export()
        rng = np.random.default_rng(12345)
        r = rng.standard_normal(system.ndofs)
        b, _ = system.random_rhs(0)
        data = flatten(hier)
        data["r"] = r
        data["b"] = b
        data["tol"] = np.array(tol)
        data["idofs"] = system.interface_dofs
        for order in ("multicolor", "natural"):
            orc = Oracle(hier, order)
            data[f"z_{order}"] = orc.apply(r)
            _, info = orc.pcg(b, tolerance=tol, maxiter=500)
            data[f"residuals_{order}"] = np.array(info["residuals"])
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **data)
        print(name, system.ndofs, len(hier["levels"]), len(data["residuals_multicolor"]) - 1,
              len(data["residuals_natural"]) - 1)


if __name__ == "__main__":
    main()
