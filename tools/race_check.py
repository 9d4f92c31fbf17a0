"""Static race check of the device layouts of the BASELINE configurations (mamg_race_check), as a
stand-in for `compute-sanitizer --tool racecheck`, which is closed on this GPU pool:
    python tools/race_check.py > profiles/r02_race_check.log"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems

CASES = [
    ("C1 bidomain_2d n=256", lambda: problems.bidomain_system(2, 256, gamma=1e3), params.parameters_metric_schwarz),
    ("C2 emi_2d n=512", lambda: problems.emi_system(2, 512, gamma=1e6), params.default_metric_parameters),
    ("C3-like bidomain_3d n=64", lambda: problems.bidomain_system(3, 64, gamma=1e4), params.parameters_metric_schwarz),
    ("C4-like emi_3d n=96", lambda: problems.emi_system(3, 96, gamma=1e6), params.default_metric_parameters),
]
for name, mk, prm in CASES:
    s = mk()
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs).to_device(0)
    gs, pt, checked = H.race_check()
    st = H.stats(0)
    print(f"{name}: {s.ndofs} dofs, {H.num_levels} levels, {st['n_patches']} patches in {st['n_patch_colors']} colours: "
          f"{checked} colour launches checked, GS conflicts {gs}, patch conflicts {pt}", flush=True)
    assert gs == 0 and pt == 0
print("race check clean")
