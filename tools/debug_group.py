"""Developer probe of the grouped Schwarz path: statistics and the real launch error (graph capture off)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["MAMG_GRAPH"] = "0"
import numpy as np

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems
from oracle import Oracle

for name, s, prm in (("emi3d n=16", problems.emi_system(3, 16, gamma=1e6), params.default_metric_parameters),
                     ("bidomain3d n=20", problems.bidomain_system(3, 20, gamma=1e4), params.parameters_metric_schwarz)):
    H = mamg.Hierarchy(s.A, dict(prm, cycle_type=1), s.interface_dofs).to_device(0)
    st = H.stats(0)
    print(name, {k: st[k] for k in ("n_patches", "unique_blobs", "schwarz_grouped", "schwarz_groups", "schwarz_grouped_patches",
                                    "schwarz_group_smem", "schwarz_group_nn_max", "schwarz_group_s_max", "max_patch_size")}, flush=True)
    r = np.random.default_rng(0).standard_normal(s.ndofs)
    try:
        z = H.apply(r)
        orc = Oracle(H.export(), "multicolor")
        print("  apply rel err", np.linalg.norm(z - orc.apply(r)) / np.linalg.norm(z), flush=True)
    except Exception as e:
        print("  FAILED:", e, flush=True)
