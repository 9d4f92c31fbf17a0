"""One small apply + PCG of every kernel family, for `a sanitizer (compute-sanitizer is closed on the development pool)`:
bidomain 3-D (SELL row kernels, Schwarz fast path, persistent tail) and EMI 3-D (general Schwarz
kernel).  Checks the result against the oracle so that a sanitizer-clean run is also a correct one."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems
from oracle import Oracle

for name, s, prm in (("bidomain3d n=10", problems.bidomain_system(3, 10, gamma=1e4), params.parameters_metric_schwarz),
                     ("emi3d n=12", problems.emi_system(3, 12, gamma=1e6), params.default_metric_parameters)):
    H = mamg.Hierarchy(s.A, dict(prm, cycle_type=1), s.interface_dofs).to_device(0)
    orc = Oracle(H.export(), "multicolor")
    r = np.random.default_rng(0).standard_normal(s.ndofs)
    z = H.apply(r)
    err = np.linalg.norm(z - orc.apply(r)) / np.linalg.norm(z)
    b, _ = s.random_rhs(0)
    x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=50)
    print(f"{name}: apply rel err {err:.2e}, pcg {info['niters']} its, launches {H.launch_count()}", flush=True)
    assert err < 1e-10
