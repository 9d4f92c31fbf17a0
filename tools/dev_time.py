"""Developer timing probe (not the bench): apply / PCG / SpMV times on one GPU."""
import argparse
import json
import sys
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import metric_amg_examples_b200 as mamg
from metric_amg_examples_b200 import params, problems, haznics_compat as hz

ap = argparse.ArgumentParser()
ap.add_argument("--kind", default="bidomain")
ap.add_argument("--dim", type=int, default=3)
ap.add_argument("-n", type=int, default=64)
ap.add_argument("--gamma", type=float, default=1e4)
ap.add_argument("--prm", default="parameters_metric_schwarz")
ap.add_argument("--cycle", default="W")
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--pcg", type=int, default=1)
a = ap.parse_args()

t0 = time.time()
s = (problems.bidomain_system if a.kind == "bidomain" else problems.emi_system)(a.dim, a.n, gamma=a.gamma)
t_asm = time.time() - t0
prm = dict(getattr(params, a.prm), cycle_type=hz.W_CYCLE if a.cycle == "W" else hz.V_CYCLE)
t0 = time.time()
H = mamg.Hierarchy(s.A, prm, s.interface_dofs)
t_setup = time.time() - t0
stream = torch.cuda.Stream()
t0 = time.time()
H.to_device(0, stream.cuda_stream)
t_up = time.time() - t0
out = {"ndofs": s.ndofs, "nnz": int(s.A.nnz), "levels": H.num_levels, "t_asm": t_asm, "t_setup": t_setup,
       "t_upload": t_up, "dev_GB": H.device_bytes() / 1e9, "cycle_GB": H.cycle_bytes() / 1e9,
       "lvl": [(H.level_info(l)["rows"], H.level_info(l)["nnz"], H.level_info(l)["n_colors"]) for l in range(min(H.num_levels, 4))],
       "patches": H.level_info(0)["n_patches"], "pcolors": H.level_info(0)["n_patch_colors"]}
r = torch.randn(s.ndofs, dtype=torch.float64, device="cuda")
with torch.cuda.stream(stream):
    def timeit(fn, reps):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    import ctypes as C
    from metric_amg_examples_b200._capi import lib
    z = torch.empty_like(r)
    H.launch_count(reset=True)
    ms = timeit(lambda: lib.mamg_apply(H._h, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr()), 1), a.reps)
    out["apply_ms"] = ms
    out["apply_launches"] = H.launch_count(reset=True) / (a.reps + 1)
    out["apply_GBs_alg"] = H.cycle_bytes() / ms / 1e6
    y = torch.empty_like(r)
    ms = timeit(lambda: lib.mamg_spmv(H._h, 0, C.c_void_p(r.data_ptr()), C.c_void_p(y.data_ptr()), 1), 20)
    n, nnz = s.ndofs, s.A.nnz
    out["spmv+2gather_ms"] = ms
    out["spmv_alg_GBs"] = (12 * nnz + 4 * (n + 1) + 16 * n) / ms / 1e6
    if a.pcg:
        b = torch.from_numpy(s.random_rhs(0)[0]).cuda()
        torch.cuda.synchronize()
        t0 = time.time()
        x, info = H.pcg(b, tolerance=1e-8, relative=True, maxiter=300)
        torch.cuda.synchronize()
        out["pcg_s"] = time.time() - t0
        out["pcg_iters"] = info["niters"]
        out["dof_per_s"] = s.ndofs / out["pcg_s"]
print(json.dumps(out))
