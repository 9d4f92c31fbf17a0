import os, sys, time
sys.path.insert(0, ".")
os.environ["MAMG_SETUP_TIMING"] = "1"
import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems
kind, n = sys.argv[1], int(sys.argv[2])
t = time.time()
s = problems.emi_system(3, n, gamma=1e6) if kind == "emi" else problems.bidomain_system(3, n, gamma=1e4)
prm = params.default_metric_parameters if kind == "emi" else params.parameters_metric_schwarz
print("assemble", round(time.time() - t, 2), s.ndofs, flush=True)
t = time.time(); H = mamg.Hierarchy(s.A, dict(prm, cycle_type=1), s.interface_dofs); print("setup", round(time.time() - t, 2), flush=True)
t = time.time(); H.to_device(0); print("upload", round(time.time() - t, 2), flush=True)
import numpy as np
b, _ = s.random_rhs(0)
t = time.time(); x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=200)
print("pcg", info["niters"], "iterations", round(time.time() - t, 2), "s; true relative residual",
      float(np.linalg.norm(s.A @ x - b) / np.linalg.norm(b)), flush=True)
