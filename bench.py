#!/usr/bin/env python
"""bench.py -- metric-AMG preconditioned CG solve on B200 (BASELINE.json metric:
"V-cycle ms & solve DOF/s to rtol 1e-8 at 1/2/4/8 B200; SpMV % of HBM peak").

A "step" is one complete solve of the hot path: cbc.block-style PCG to rtol 1e-8
(relativeconv=True) with the metric-AMG cycle as preconditioner on one synthetic system.
  value   DOF/s with b already resident in HBM (CUDA events around K solves)
  e2e     the same solves through the public API (ConjGrad(A, precond=B) * b) with HOST
          vectors: H2D of b and D2H of x inside the timed region
Default workload at every N: BASELINE.json configs[3], emi_3d on UnitCubeMesh(464) (100.8 M DOFs, the
configuration the north-star target is quoted on; it fits one B200) with the reference's default metric
parameters (src/utils.py:60-82); `--workload bidomain_3d` is configs[2] (n=199, 16.0 M DOFs,
parameters_metric_schwarz).  The headline uses cycle_type V_CYCLE (the metric names the V-cycle); one apply of
the W-cycle that src/amg_parameters.py configures is timed beside it on the same hierarchy at N=1
(`wcycle_ms`, --wcycle K applies; 0 disables).
`--impl reference` times the CPU restatement (oracle/) on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# torchrun pins OMP_NUM_THREADS=1; the host setup (aggregation, Galerkin products, patch lists) is
# OpenMP code: give every rank its share of the cores before libgomp starts
# stdout carries exactly one JSON line: NCCL's own banner (NCCL_DEBUG=VERSION/INFO on some boxes) goes to stderr
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if int(os.environ.get("WORLD_SIZE", "1")) > 1 and os.environ.get("OMP_NUM_THREADS", "1") == "1":
    _ref_arm = any(sys.argv[k] == "--impl" and sys.argv[k + 1] == "reference" for k in range(len(sys.argv) - 1)) \
        or "--impl=reference" in sys.argv
    # the reference arm runs on rank 0 alone and gets every host core at every N; the GPU arm's ranks share them
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1) if _ref_arm else \
        str(max(1, (os.cpu_count() or 1) // int(os.environ.get("LOCAL_WORLD_SIZE", os.environ["WORLD_SIZE"]))))


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="mamg", choices=["mamg", "reference"])
    ap.add_argument("--workload", default="emi_3d", choices=["bidomain_2d", "bidomain_3d", "emi_2d", "emi_3d"])
    ap.add_argument("-n", type=int, default=None, help="cells per direction (default: the BASELINE config size)")
    ap.add_argument("--gamma", type=float, default=1e4)
    ap.add_argument("--cycle", default="V", choices=["V", "W"])
    ap.add_argument("--wcycle", type=int, default=1, help="also time this many W-cycle applies on the same hierarchy (0: skip)")
    ap.add_argument("--cpu-sample-n", type=int, default=None, help="mesh size of the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--rtol", type=float, default=1e-8)
    return ap.parse_args()


DEFAULT_N = {"bidomain_2d": 256, "bidomain_3d": 199, "emi_2d": 2048, "emi_3d": 464}
CPU_SAMPLE_N = {"bidomain_2d": 128, "bidomain_3d": 56, "emi_2d": 256, "emi_3d": 128}   # 10-30 s of one-thread work


def make_system(workload, n, gamma):
    from metric_amg_examples_b200 import params, problems
    kind, dim = workload.split("_")
    dim = int(dim[0])
    if kind == "bidomain":
        s = problems.bidomain_system(dim, n, gamma=gamma)
        prm = dict(params.parameters_metric_schwarz)   # src/bidomain_2d.py:201, src/bidomain_3d.py:145
    else:
        s = problems.emi_system(dim, n, gamma=gamma)
        prm = dict(params.default_metric_parameters)   # src/emi_2d.py:207 passes no parameters
    return s, prm


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[k] for r in self.rows if len(r) >= 7 for k in range(4) if r[3 + k].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def class_bytes(H):
    """Algorithmic bytes per kernel class (fp64 values, int32 columns, every vector
    counted once per kernel), counted from the rows each launch really processes.  Returns
    f(niters, cycle_applies) -> {class: bytes} for that many V-cycle applies plus CG iterations (the
    level data is read here, while the host copy of the hierarchy still exists)."""
    import numpy as np
    import ctypes as C
    from metric_amg_examples_b200._capi import lib, ptr
    L = H.num_levels
    infos = [H.level_info(l) for l in range(L)]
    prm = H.params
    out = {k: 0.0 for k in H.KERNEL_CLASSES}
    gs_lvl = []
    sgs = prm.smoother in (3, 6)
    for l in range(L - 1):
        n, nnz, nc = infos[l]["rows"], infos[l]["nnz"], infos[l + 1]["rows"]
        nnzc = infos[l + 1]["nnz"]
        # rows the point smoother really touches (Schwarz seeds are skipped) and the share of the
        # colour that the backward SGS sweep skips
        indptr = np.empty(n + 1, np.int32)
        color = np.empty(n, np.int32)
        skip = np.empty(n, np.uint8)
        lib.mamg_level_export(H._h, l, ptr(indptr), None, None, None, ptr(color), ptr(skip))
        rl = np.diff(indptr).astype(np.int64)
        act = skip == 0
        nnz_act, n_act = int(rl[act].sum()), int(act.sum())
        last = act & (color == infos[l]["n_colors"] - 1)
        nnz_last, n_last = int(rl[last].sum()), int(last.sum())
        sweep = 12 * nnz_act + 4 * (n_act + 1) + 24 * n_act
        sweep_bw = sweep - (12 * nnz_last + 28 * n_last) if prm.smoother == 3 else sweep
        per_smooth = (sweep + sweep_bw) if sgs else sweep
        out["gs"] += per_smooth * (prm.presmooth_iter + prm.postsmooth_iter)
        gs_lvl.append(per_smooth * (prm.presmooth_iter + prm.postsmooth_iter))
        if infos[l]["n_patches"]:
            # "Schwarz level-0 (stored factors)" -- every patch owns its inverse and its row values;
            # the shared-blob figure (what the kernel has to bring in at least once per launch) is kept beside it
            st = H.stats(l)
            nsw = 2 if prm.Schwarz_type == 3 else 1
            out["schwarz"] += 2 * nsw * st["schwarz_sweep_bytes_stored_factors"]
            out["schwarz_shared"] = out.get("schwarz_shared", 0.0) + 2 * nsw * st["schwarz_sweep_bytes"]
        agg = np.empty(n, np.int32)
        lib.mamg_level_export(H._h, l, None, None, None, ptr(agg), None, None)
        inagg = agg >= 0
        out["restrict"] += 12 * int(rl[inagg].sum()) + (4 + 4 + 8 + 8) * int(inagg.sum()) + 16 * nc
        if prm.coarse_scaling:
            out["scale"] += 12 * nnzc + 4 * (nc + 1) + 16 * nc
        out["prolong"] += 4 * n + 16 * int(inagg.sum()) + 8 * nc
    ncst = infos[-1]["rows"]
    out["coarse"] += 8 * ncst * ncst + 16 * ncst
    n0, nnz0 = infos[0]["rows"], infos[0]["nnz"]
    per_apply = out

    def finish(niters, cycle_applies):
        res = {k: v * cycle_applies for k, v in per_apply.items()}
        res["spmv"] += niters * (12 * nnz0 + 4 * (n0 + 1) + 24 * n0)
        res["dot"] += (niters + 1) * 16 * n0
        # pcg update (48 n) + direction (24 n) per iteration, zero-fill of z per apply, gathers/copies at both ends
        res["vector"] += niters * 72 * n0 + cycle_applies * 8 * n0 + 5 * 20 * n0
        res["gs_by_level"] = [v * cycle_applies for v in gs_lvl]
        return res
    return finish


def host_copies(world):
    """Ranks that hold a host copy of the whole hierarchy while it is built."""
    return world


def problems_slab(system, nparts):
    """z-slabs for bidomain; x-strips for EMI, so that every rank owns a strip of the interface and its
    Schwarz patches (z-slabs would hand the whole interface to one rank)."""
    from metric_amg_examples_b200 import problems
    return problems.slab_partition(system, nparts, axis=0 if system.name.startswith("emi") else None)


def run_mamg(a):
    import numpy as np
    import torch
    import torch.distributed as dist
    import metric_amg_examples_b200 as mamg
    from metric_amg_examples_b200 import haznics_compat as hz
    from metric_amg_examples_b200.iterative import ConjGrad
    from metric_amg_examples_b200.precond import metricAMG

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n = a.n or DEFAULT_N[a.workload]
    note = ""
    # Host memory: every rank assembles the system and builds the hierarchy on the host (setup + upload
    # temporaries per DOF below), uploads its share and frees the host copy.  The ranks do this in waves of
    # as many ranks as the node's free memory takes; the mesh is never reduced silently.
    import psutil
    dim = int(a.workload[-2])
    ndof_est = 2 * (n + 1) ** dim if a.workload.startswith("bidomain") else (n + 1) ** (dim - 1) * (n + 2)
    need = (3600.0 if a.workload.startswith("bidomain") else 650.0) * ndof_est   # measured peaks: emi_3d n=464 50.5 GB
    avail = float(psutil.virtual_memory().available)
    if world > 1:
        free_t = torch.tensor([avail], dtype=torch.float64, device="cuda")
        dist.all_reduce(free_t, op=dist.ReduceOp.MIN)
        avail = float(free_t.item())
    if need > 0.9 * avail:
        raise SystemExit(f"bench.py: {a.workload} n={n} needs ~{need / 1e9:.0f} GB of host memory per rank for the setup, "
                         f"{avail / 1e9:.0f} GB available; the mesh is never reduced silently")
    conc = int(max(1, min(world, 0.9 * avail // need)))
    stream = torch.cuda.Stream()
    t_asm = t_setup = t_upload = 0.0
    for w0 in range(0, world, conc):
        if w0 <= rank < w0 + conc:
            t0 = time.time()
            system, prm = make_system(a.workload, n, a.gamma)
            t_asm = time.time() - t0
            prm["cycle_type"] = hz.V_CYCLE if a.cycle == "V" else hz.W_CYCLE
            ndofs, nnz0 = system.ndofs, int(system.A.nnz)
            b_host, x_true = system.random_rhs(0)       # the same system and right-hand side on every rank
            part = problems_slab(system, world) if world > 1 else None
            t0 = time.time()
            B = metricAMG(system.A, system.W, idofs=system.interface_dofs, parameters=prm, device=local, part=part)
            H = B.hierarchy
            t_setup = time.time() - t0
            t0 = time.time()
            if world > 1:
                B.A = None
                system.A = None                      # the library has its own copy; free the scipy one
                H.to_device(local, stream.cuda_stream, rank=rank, world=world)
            else:
                H.to_device(local, stream.cuda_stream)
            t_upload = time.time() - t0
            cb_of = class_bytes(H) if rank == 0 else None
            nnz_levels = [[H.level_info(l)["nnz"], H.level_info(l)["nnz_structural"]] for l in range(min(H.num_levels, 4))]
            if world > 1:
                H.release_host()                     # the next wave needs the memory
        if world > 1:
            dist.barrier()
    if world > 1:
        H.dist_init()
    import resource
    host_peak_gb = resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1e6
    b_pin = torch.from_numpy(b_host).pin_memory()
    b_dev = b_pin.cuda(non_blocking=False)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    iters = []

    def solve_dev():
        x, info = H.pcg(b_dev, tolerance=a.rtol, relative=True, maxiter=500)
        iters.append(info["niters"])
        return x, info

    with torch.cuda.stream(stream):
        # the first solve through the public API with host vectors: what the reference's timer sees after
        # its setup (src/bidomain_2d.py:184-207 times metricAMG(...) + ConjGrad solve); graph capture included
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        H.pcg(b_pin.numpy(), tolerance=a.rtol, relative=True, maxiter=500)
        t_first = time.perf_counter() - t0
        for _ in range(a.warmup):
            x, info = solve_dev()
        rel_err = float(torch.linalg.norm(x - torch.from_numpy(x_true).cuda()) / np.linalg.norm(x_true))
        # ---- device-resident timing: K solves between CUDA events on the launching stream ----
        H.launch_count(reset=True)
        barrier()
        sampler = ClockSampler(local) if rank == 0 else None
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            x, info = solve_dev()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = H.launch_count(reset=True)
        clocks = sampler.stop() if sampler else None
        # ---- end-to-end through the public API with host vectors (H2D b, D2H x inside) ----
        bh = b_pin.numpy()
        if world == 1:
            solver = ConjGrad(system.A, precond=B, tolerance=a.rtol, relativeconv=True, maxiter=500, show=0)
            e2e_solve = lambda: solver * bh
        else:
            e2e_solve = lambda: H.pcg(bh, tolerance=a.rtol, relative=True, maxiter=500)[0]
        xs = e2e_solve()  # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(a.steps):
            xs = e2e_solve()
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        if world == 1:
            assert solver.mode == "fused"
        # ---- one V-cycle apply, and per-kernel-class times of one solve ----
        import ctypes as C
        from metric_amg_examples_b200._capi import lib
        r = torch.randn(ndofs, dtype=torch.float64, device="cuda")
        z = torch.empty_like(r)
        for _ in range(3):
            lib.mamg_apply(H._h, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr()), 1)
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record(stream)
        for _ in range(5):
            lib.mamg_apply(H._h, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr()), 1)
        c1.record(stream)
        torch.cuda.synchronize()
        cycle_ms = c0.elapsed_time(c1) / 5
        # the W-cycle that src/amg_parameters.py:5,25,49,69 configures, on the same hierarchy
        wcycle_ms = None
        # (one rank only: below the distributed levels every rank runs the same coarse sequence, which is where a
        # W apply spends its time; no warm-up apply: the W sequence is far above the graph-capture limit, so every
        # apply launches eagerly, and all of its kernels have already run in the V-cycles above)
        if a.wcycle > 0 and a.cycle == "V" and world == 1:
            H.set_cycle(hz.W_CYCLE)
            torch.cuda.synchronize()
            c0.record(stream)
            for _ in range(a.wcycle):
                lib.mamg_apply(H._h, C.c_void_p(r.data_ptr()), C.c_void_p(z.data_ptr()), 1)
            c1.record(stream)
            torch.cuda.synchronize()
            wcycle_ms = c0.elapsed_time(c1) / a.wcycle
            H.set_cycle(hz.V_CYCLE)
        H.profile_start()
        _, pinfo = solve_dev()
        prof_lv = H.profile_levels()
        prof = H.profile_stop()
        ncoll = H.collective_count(reset=True)
        H.exchange_bytes(reset=True)
        _, _ = solve_dev()
        ncoll = H.collective_count()
        xbytes = H.exchange_bytes()
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    ms, e2e_s = float(t.item()), float(te.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    nit = pinfo["niters"]
    halo_mode = world > 1 and os.environ.get("MAMG_HALO", "1") != "0" and os.environ.get("MAMG_P2P", "1") != "0"
    exch_mode = ("NCCL grouped broadcasts" if os.environ.get("MAMG_P2P", "1") == "0" else
                 "halo index lists stored into the neighbour ranks' vectors over NVLink (CUDA IPC), neighbour-only flags, "
                 "rank-ordered all-reduce of the dots" if halo_mode else
                 "peer-memory stores over NVLink fused with the flag barrier (CUDA IPC), all-gather of the updated ranges")
    cb = cb_of(nit, nit + 1)
    tot_ms = sum(v[0] for v in prof.values())
    dom = max(prof, key=lambda k: prof[k][0])
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    dom_ms, dom_launches = prof[dom]
    # with several ranks the smoothing / transfer / SpMV classes run on 1/world of the rows of the
    # distributed levels (per-rank numbers; small replicated levels make this a slight under-estimate)
    gs_by_level = cb.pop("gs_by_level")
    share = {k: (1.0 / world if k in ("spmv", "gs", "schwarz", "schwarz_shared", "restrict", "scale", "prolong") else 1.0) for k in cb}
    achieved = cb[dom] * share[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    kernels = {k: {"ms": round(prof[k][0], 3), "launches": prof[k][1], "share": round(prof[k][0] / tot_ms, 4),
                   "alg_GBs": round(cb[k] * share[k] / (prof[k][0] * 1e-3) / 1e9, 1) if prof[k][0] > 0 else None}
               for k in prof}
    if prof["schwarz"][0] > 0:
        kernels["schwarz"]["alg_GBs_shared_blob_model"] = round(cb.get("schwarz_shared", 0.0) * share["schwarz"] / (prof["schwarz"][0] * 1e-3) / 1e9, 1)
    # DRAM traffic per launch of the dominant class: (dram bytes / algorithmic bytes) of that kernel class in
    # the committed `ncu --set full` capture of this workload (profiles/traffic.json, smaller mesh, same
    # kernels) x this run's algorithmic bytes per launch; null when no capture of that class is committed
    traffic, traffic_src = None, "no committed ncu --set full capture of this kernel class for this workload"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        ent = tj.get(a.workload, {}).get(dom)
        if ent:
            base = cb["schwarz_shared"] if dom == "schwarz" else cb[dom]   # the Schwarz ratio was taken against the shared-blob bytes
            traffic = ent["dram_over_algorithmic"] * base * share[dom] / max(prof[dom][1], 1)
            traffic_src = ent["source"]
    except Exception:
        pass
    out = {
        "metric": "solve DOF/s to rtol 1e-8 (metric-AMG V-cycle PCG)", "value": ndofs * a.steps / (ms * 1e-3),
        "unit": "DOF/s", "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": (f"{a.workload} UnitCubeMesh({n}) P1 gamma={a.gamma:g} ndofs={ndofs} nnz={nnz0}"
                                if a.workload.endswith("3d") else
                                f"{a.workload} UnitSquareMesh({n}) P1 gamma={a.gamma:g} ndofs={ndofs} nnz={nnz0}") + note,
                   "precond": "metricAMG parameters_metric_schwarz" if a.workload.startswith("bidomain") else "metricAMG default_metric_parameters",
                   "cycle_type": a.cycle, "krylov": f"ConjGrad relativeconv tolerance={a.rtol:g}",
                   "levels": H.num_levels,
                   "multi_gpu": (f"one system row-partitioned over {world} GPUs ({world} {'x-strips' if a.workload.startswith('emi') else 'z-slabs'}), "
                                 f"{'matrices stored per rank, ' if halo_mode else 'hierarchy replicated, '}"
                                 f"exchange: {exch_mode}; {ncoll} exchanges and {xbytes / 1e6:.1f} MB sent per rank and solve; levels < "
                                 f"{os.environ.get('MAMG_DIST_MIN_ROWS', '6000000' if halo_mode else '1000000')} rows replicated")
                   if world > 1 else "single",
                   "l2_note": (f"level-0 matrix {12 * nnz0 / 1e9:.2f} GB and vectors {8 * ndofs / 1e6:.0f} MB each: "
                               + ("larger than" if 12 * nnz0 > 126e6 else "NOT larger than") + " the 126 MB L2")},
        "iterations": nit, "vcycle_ms": cycle_ms, "wcycle_ms": wcycle_ms, "rel_error_vs_x_true": rel_err,
        "first_solve_e2e_s": round(t_setup + t_upload + t_first, 2),
        "first_solve_note": "host setup + upload + first solve with host vectors (assembly excluded): what the "
                            "reference's timer brackets (src/bidomain_2d.py:184-207)",
        "spmv_hbm_frac": {"alg_GBs": kernels["spmv"]["alg_GBs"],
                          "of_nominal_8000": round((kernels["spmv"]["alg_GBs"] or 0) / 8000.0, 3),
                          "of_measured": round((kernels["spmv"]["alg_GBs"] or 0) / peak, 3)},
        "nnz_stored_vs_structural": nnz_levels,
        "e2e": {"value": ndofs * a.steps / e2e_s, "unit": "DOF/s", "h2d_bytes_per_step": 8 * ndofs,
                "d2h_bytes_per_step": 8 * ndofs},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                     "algorithmic_bytes_per_launch": cb[dom] * share[dom] / max(dom_launches, 1),
                     "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy)" if peaks else "fallback 6650 GB/s",
                     "byte_model": ("SURVEY 8(d) stored-factor model (every patch owns its inverse and row values). The kernels share "
                                    "the blobs of look-alike patches, so their DRAM traffic is far below this figure and they are "
                                    "instruction/latency-bound; against the bytes they really have to bring in the fraction is "
                                    "frac_shared_blob_model" if dom == "schwarz" else "SURVEY 8(d) on stored entries"),
                     "frac_shared_blob_model": (round(kernels["schwarz"].get("alg_GBs_shared_blob_model", 0.0) / peak, 4) if dom == "schwarz" else None),
                     "avg_launch_ms": dom_ms / max(dom_launches, 1)},
        "kernels": kernels,
        "schwarz_blobs": {k: H.stats(0)[k] for k in ("n_patches", "unique_blobs", "schwarz_sweep_bytes",
                                                      "schwarz_sweep_bytes_stored_factors", "schwarz_fast_path")},
        "sell": [{"rows": H.stats(l)["rows"], "nnz_stored": H.stats(l)["nnz_stored"], "slots": H.stats(l)["sell_slots"]}
                 for l in range(min(H.num_levels, 4))],
        "gs_ms_by_level": [round(float(v), 2) for v in prof_lv[:, 1]],
        "gs_alg_GBs_by_level": [round(gs_by_level[l] / world / (float(prof_lv[l, 1]) * 1e-3) / 1e9, 1) if l < len(gs_by_level) and prof_lv[l, 1] > 0
                                else None for l in range(min(H.num_levels, 8))],
        "level_rows": [H.level_info(l)["rows"] for l in range(H.num_levels)],
        "host": {"assemble_s": round(t_asm, 2), "setup_s": round(t_setup, 2), "upload_s": round(t_upload, 2),
                 "device_GB": round(H.device_bytes() / 1e9, 2), "host_peak_GB": round(host_peak_gb, 1),
                 "setup_waves": (world + conc - 1) // conc},
    }
    if not a.no_cpu_baseline and world == 1:
        out["cpu_baseline"] = cpu_sample(a, threads=1, ordering="natural", gpu_check=True)
        out["sample_iters"] = {"gpu": out["cpu_baseline"].pop("iters_gpu"), "oracle_multicolor": out["cpu_baseline"].pop("iters_oracle_multicolor"),
                               "oracle_natural": out["cpu_baseline"]["iters"]}
    print(json.dumps(out), file=RESULT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def cpu_sample(a, threads, ordering, steps=1, gpu_check=False):
    """The oracle (CPU port of the path) on a bounded sample of the same workload.  gpu_check: also run
    the device path on the sample's mesh, so that the line carries a GPU-vs-oracle iteration check."""
    import metric_amg_examples_b200 as mamg
    from metric_amg_examples_b200 import haznics_compat as hz
    from oracle import Oracle
    n = a.cpu_sample_n or CPU_SAMPLE_N[a.workload]
    system, prm = make_system(a.workload, n, a.gamma)
    prm["cycle_type"] = hz.V_CYCLE if a.cycle == "V" else hz.W_CYCLE
    H = mamg.Hierarchy(system.A, prm, system.interface_dofs)
    orc = Oracle(H.export(), ordering)
    used = orc.set_threads(threads) if ordering == "multicolor" else 1
    b, _ = system.random_rhs(0)
    t0 = time.perf_counter()
    for _ in range(steps):
        _, info = orc.pcg(b, tolerance=a.rtol, relative=True, maxiter=500)
    dt = time.perf_counter() - t0
    extra = {}
    if gpu_check:
        orc.set_ordering("multicolor")
        orc.set_threads(0)
        _, im = orc.pcg(b, tolerance=a.rtol, relative=True, maxiter=500)
        H.to_device(int(os.environ.get("LOCAL_RANK", "0")))
        _, ig = H.pcg(b, tolerance=a.rtol, relative=True, maxiter=500)
        extra = {"iters_gpu": ig["niters"], "iters_oracle_multicolor": im["niters"]}
    return {**extra, "iters": info["niters"],
            "value": system.ndofs * steps / dt, "unit": "DOF/s", "cores": used, "kind": "port",
            "sample": f"{a.workload} n={n} ({system.ndofs} dofs), same parameters, full PCG solve to rtol {a.rtol:g}, "
                      f"{ordering} smoother order, {info['niters']} iterations, {dt / steps:.2f} s per solve",
            "host_cores_available": os.cpu_count(), "seconds_per_solve": dt / steps}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.n or DEFAULT_N[a.workload]
    steps = max(1, a.steps)
    for _ in range(min(a.warmup, 1)):
        cpu_sample(a, threads=0, ordering="multicolor")
    t0 = time.perf_counter()
    cb = cpu_sample(a, threads=0, ordering="multicolor", steps=steps)
    dt = time.perf_counter() - t0
    out = {
        "impl": "reference", "metric": "solve DOF/s to rtol 1e-8 (metric-AMG V-cycle PCG)", "value": cb["value"],
        "unit": "DOF/s", "n_gpus": a.gpus, "steps": steps, "warmup": min(a.warmup, 1),
        "ms_per_step": cb["seconds_per_solve"] * 1e3, "sample_wall_s_with_setup": round(dt, 2),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{a.workload} n={n} gamma={a.gamma:g} (timed on the bounded sample below)",
                   "cycle_type": a.cycle, "krylov": f"ConjGrad relativeconv tolerance={a.rtol:g}"},
        "cpu_baseline": cb,
        "e2e": {"value": cb["value"], "unit": "DOF/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "HAZmath/cbc.block are not installable here (SURVEY 8c); this arm is the repo's CPU oracle of "
                "the same path (multicolour order, OpenMP over colour classes) on all host threads",
    }
    print(json.dumps(out), file=RESULT, flush=True)


if __name__ == "__main__":
    args = parse()
    # the real stdout carries exactly one JSON line; whatever libraries print to fd 1 (NCCL's version
    # banner under NCCL_DEBUG=VERSION ignores NCCL_DEBUG_FILE) is sent to stderr instead
    sys.stdout.flush()
    RESULT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_mamg(args)
