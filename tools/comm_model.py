"""Communication volume of a row-partitioned hierarchy (planning tool for the multi-GPU path; CPU only).

    python tools/comm_model.py emi_3d 64 8 [464]      # workload, cells per direction, parts [, size to scale to]
    python tools/comm_model.py bidomain_3d 48 4 [199]

For every level that would be row-distributed it prints the rows per part, the Gauss-Seidel colours,
the halo (rows of a part that another part's rows reference) per neighbour, and for the Schwarz level
the patches whose neighbourhood crosses a cut; then a latency/bandwidth estimate of one V-cycle for
(a) the round-1 scheme (every updated range completed on every rank, one all-rank barrier per colour;
MAMG_HALO=0) and (b) the halo mode of round 2 (neighbour-only exchange of halo rows).  With a fourth
argument the counted sizes are scaled to that mesh size (rows, entries and volume patches by f^3, halos
and interface patches by f^2, colour counts unchanged) before the estimate, and the distribution
thresholds are the library's (1 M rows all-gather scheme, 6 M rows halo mode).

The constants are the round-2 measurements: a dependent exchange step costs its
kernel (never less than the latency floor of a launch whose CTAs walk a chain of dependent loads: about
15 us for a Schwarz colour, 4 us for a row kernel) plus a 12 us neighbour handshake.  Measured at
emi_3d n=464 (profiles/r02_scale_emi3d_n464_<N>gpu.json): 39.9 / 32.2 / 26.5 / 21.4 ms per V-cycle on
1 / 2 / 4 / 8 GPUs.
"""
import os
import sys

import numpy as np
import scipy.sparse as sp

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import metric_amg_examples_b200 as mamg  # noqa: E402
from metric_amg_examples_b200 import params, problems  # noqa: E402

T_EXCH = 22e-6        # measured (round 1): one all-rank exchange (push + flag barrier), seconds
T_P2P = 12e-6         # measured (round 2): neighbour push + flag round trip behind a kernel
T_FLOOR_SW = 15e-6    # measured: one Schwarz colour launch, however few patches a rank owns
T_FLOOR_ROW = 4e-6    # a row-kernel launch inside a replayed graph
BW_HBM = 4.3e12       # algorithmic bandwidth of a whole V-cycle on one GPU (39.9 ms at emi_3d n=464), B/s
BW_LINK = 600e9       # usable NVLink bandwidth per direction, B/s


def straddle_scaled(count, f, f2, kind):
    """patches whose neighbourhood crosses a cut: a surface (bidomain) or a line on the interface (EMI)"""
    return int(count * (f2 if kind == "bidomain" else f))


def main():
    workload, n, nparts = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    f = (int(sys.argv[4]) / n) if len(sys.argv) > 4 else 1.0
    kind, dim = workload.split("_")
    dim = int(dim[0])
    if kind == "bidomain":
        # keep gamma h^2 (mass against stiffness coupling) of the target size: it decides how HEM pairs
        gamma = 1e4 * (n / int(sys.argv[4])) ** 2 if len(sys.argv) > 4 else 1e4
        s, prm, axis = problems.bidomain_system(dim, n, gamma=gamma), params.parameters_metric_schwarz, None
    else:
        s, prm, axis = problems.emi_system(dim, n, gamma=1e4), params.default_metric_parameters, 0
    part = problems.slab_partition(s, nparts, axis=axis)
    H = mamg.Hierarchy(s.A, prm, s.interface_dofs, part=part)
    ex = H.export()
    min_rows = int(os.environ.get("MODEL_DIST_MIN_ROWS", 1000000 if f > 1 else max(2000, s.ndofs // 100)))
    min_rows_halo = int(os.environ.get("MODEL_DIST_MIN_ROWS_HALO", 6000000 if f > 1 else min_rows))
    f3, f2 = f ** 3, f ** 2
    fp = f3 if kind == "bidomain" else f2                  # patches fill the volume / the interface
    t_now = t_halo = t_one = 0.0
    print(f"{workload} n={n}: {s.ndofs} dofs, {nparts} parts, levels with >= {min_rows} rows distributed")
    for l, L in enumerate(ex["levels"][:-1]):
        nl0 = L["n"]
        A = sp.csr_matrix((L["data"], L["indices"], L["indptr"]), shape=(nl0, nl0))
        nl, nnz = int(nl0 * f3), int(A.nnz * f3)
        ncol = L["n_colors"]
        npc = L.get("n_patch_colors", 0) if len(L.get("patch_ptr", ())) > 1 else 0
        gs_bytes = 4 * (12 * nnz + 28 * nl)                                    # 4 GS sweeps
        tr_bytes = (12 * nnz + 36 * nl) * 2                                    # residual + restriction, scaling, prolongation
        sw_level_bytes = 0
        if npc:
            ptr = L["patch_ptr"]
            sw_level_bytes = int(fp * 4 * (8 * np.sum(np.diff(ptr).astype(np.int64) ** 2) // 2)) + 4 * 12 * nnz // 4
        visit_bytes = gs_bytes + tr_bytes + sw_level_bytes
        t_level = visit_bytes / BW_HBM
        t_one += t_level
        if nl < min_rows:
            t_now += t_level
        if nl < min_rows_halo:
            t_halo += t_level
        if nl < min_rows:
            continue
        p = L["part"]
        C = A.tocoo()
        nz = (C.data != 0) & (C.row != C.col)
        cross = nz & (p[C.row] != p[C.col])
        halo = {}                               # (owner, reader) -> rows of owner that reader references
        for o, r_, j in zip(p[C.col[cross]], p[C.row[cross]], C.col[cross]):
            halo.setdefault((int(o), int(r_)), set()).add(int(j))
        hmax = int(f2 * max((len(v) for v in halo.values()), default=0))
        pairs = len(halo)
        straddle = 0
        if npc:
            ptr, dofs = L["patch_ptr"], L["patch_dofs"]
            adj = (A != 0).astype(np.int8).tocsr()
            for q in range(len(ptr) - 1):
                d = dofs[ptr[q]:ptr[q + 1]]
                reach = np.unique(np.concatenate([d, adj[d].indices]))
                straddle += len(np.unique(p[reach])) > 1
        e_gs, e_sw, e_tr = 2 * (2 * ncol - 1), 4 * npc, 4   # per level visit: SGS pre+post, Schwarz sweeps, transfers
        exch = e_gs + e_sw + e_tr
        own = 8.0 * nl / nparts                              # bytes of one rank's share of a level vector
        sw_bytes = 8.0 * 60 * straddle_scaled(straddle, f, f2, kind) / max(npc, 1) / nparts if npc else 0.0
        # today: a colour's rows (own / ncol) go to every peer, whole blocks after transfers and sweeps
        t_now += (t_level / nparts + exch * T_EXCH + e_gs * own / max(ncol, 1) * (nparts - 1) / BW_LINK
                  + (e_tr + (4 if npc else 0)) * own * (nparts - 1) / BW_LINK + e_sw * sw_bytes / BW_LINK)
        if nl >= min_rows_halo:
            # every exchange step waits for its kernel (at least the latency floor) and for the neighbour handshake
            step = lambda nbytes, steps, floor: steps * max(nbytes / max(steps, 1) / BW_HBM / nparts, floor)
            t_halo += (step(gs_bytes, e_gs, T_FLOOR_ROW) + step(sw_level_bytes, e_sw, T_FLOOR_SW) + step(tr_bytes, e_tr, T_FLOOR_ROW)
                       + exch * T_P2P + (e_gs / max(ncol, 1) + e_tr) * 8.0 * hmax / BW_LINK + e_sw * sw_bytes / BW_LINK)
        straddle = straddle_scaled(straddle, f, f2, kind)
        print(f"  level {l}: rows {nl} ({nl // nparts}/part) nnz/row {nnz / nl:.1f} GS colours {ncol} patch colours {npc}"
              f" | neighbour pairs {pairs}, halo rows per pair <= {hmax} ({8 * hmax / 1e3:.1f} KB)"
              + (f", patches crossing a cut {straddle} of {int(fp * (len(L['patch_ptr']) - 1))}" if npc else "")
              + f" | exchanges per visit {exch}")
    print(f"one V-cycle, model: 1 GPU {t_one * 1e3:.2f} ms; {nparts} GPUs all-gather scheme of round 1 {t_now * 1e3:.2f} ms "
          f"(speed-up {t_one / t_now:.2f}); halo mode {t_halo * 1e3:.2f} ms (speed-up {t_one / t_halo:.2f})")


if __name__ == "__main__":
    main()
