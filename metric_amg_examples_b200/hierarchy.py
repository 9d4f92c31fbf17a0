"""Thin Python owner of a mamg_handle: setup, export (for the oracle), device calls."""
import ctypes as C

import numpy as np

from . import _capi
from ._capi import as_f64, as_i32, check, lib, ptr
from .params import to_struct


def csr_arrays(A):
    """(indptr i32, indices i32, data f64, n) from whatever the caller holds.

    Accepted, in the reference's own order of preference: dolfin PETScMatrix / petsc4py Mat
    (what PETSc_to_dCSRmat takes, src/utils.py:108) when those modules exist, scipy.sparse
    matrices, torch sparse-CSR tensors, or a plain (indptr, indices, data[, shape]) tuple.
    """
    if hasattr(A, "mat") and callable(A.mat):  # dolfin PETScMatrix
        A = A.mat()
    if hasattr(A, "getValuesCSR"):  # petsc4py.PETSc.Mat
        indptr, indices, data = A.getValuesCSR()
        return as_i32(indptr), as_i32(indices), as_f64(data), len(indptr) - 1
    if isinstance(A, (tuple, list)):
        indptr, indices, data = A[:3]
        return as_i32(indptr), as_i32(indices), as_f64(data), len(indptr) - 1
    if hasattr(A, "crow_indices"):  # torch sparse CSR
        return (as_i32(A.crow_indices().cpu().numpy()), as_i32(A.col_indices().cpu().numpy()),
                as_f64(A.values().cpu().numpy()), A.shape[0])
    if hasattr(A, "tocsr"):
        A = A.tocsr()
        if A.shape[0] != A.shape[1]:
            raise ValueError("matrix must be square")
        if not A.has_sorted_indices:
            A = A.sorted_indices()
        return as_i32(A.indptr), as_i32(A.indices), as_f64(A.data), A.shape[0]
    raise TypeError(f"cannot interpret {type(A)} as a CSR matrix")


def _comm_device(group, device):
    import torch
    import torch.distributed as dist
    return torch.device("cuda", device) if dist.get_backend(group) == "nccl" else torch.device("cpu")


def broadcast_bytes(payload, src=0, group=None, device=0):
    """Broadcast a fixed-size byte string (NCCL unique id) through torch.distributed (nccl or gloo)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(payload), dtype=torch.uint8, device=_comm_device(group, device))
    dist.broadcast(t, src=src, group=group)
    return bytes(t.cpu().tolist())


def allgather_bytes(payload, group=None, device=0):
    """All-gather equal-size byte strings (CUDA IPC handles): list indexed by rank."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(payload), dtype=torch.uint8, device=_comm_device(group, device))
    out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
    dist.all_gather(out, t, group=group)
    return [bytes(o.cpu().tolist()) for o in out]


def owned_blocks(nparts, rank, world):
    """Parts executed by `rank` (contiguous): the host-side mirror of blk_lo/blk_hi in device.cu."""
    if nparts % world:
        raise ValueError(f"{nparts} parts cannot be spread evenly over {world} ranks")
    per = nparts // world
    return range(rank * per, (rank + 1) * per)


def halo_send_rows(level, nparts, rank, world):
    """Host-side mirror of build_halo_lists() in device.cu for one exported level: {neighbour rank: rows
    (natural numbering) of `rank`'s parts that a row of the neighbour couples to}.  The device code relies
    on this relation being symmetric (a rank waits for exactly the ranks it sends to)."""
    per = nparts // world
    owner = np.asarray(level["part"]) // per
    indptr, indices, data = level["indptr"], level["indices"], level["data"]
    rows = np.repeat(np.arange(level["n"]), np.diff(indptr))
    mine = owner[rows] == rank
    other = owner[indices] != rank
    out = {}
    sel = mine & other
    for q in np.unique(owner[indices[sel]]):
        out[int(q)] = np.unique(rows[sel & (owner[indices] == q)])
    return out


class Hierarchy:
    """Owns one metric-AMG hierarchy (host) and, after to_device(), its copy on one B200."""

    def __init__(self, A, parameters=None, idofs=None, part=None, nparts=None):
        indptr, indices, data, n = csr_arrays(A)
        self.n = int(n)
        self.params = to_struct(parameters)
        idofs = as_i32(idofs) if idofs is not None else np.zeros(0, np.int32)
        self.idofs = idofs
        h = C.c_void_p()
        if part is not None:
            part = as_i32(part)
            if len(part) != self.n:
                raise ValueError("part must have one entry per row")
            self.nparts = int(nparts if nparts is not None else part.max() + 1)
        else:
            self.nparts = 1
        check(lib.mamg_setup_partitioned(C.byref(self.params), self.n, ptr(indptr), ptr(indices), ptr(data),
                                         len(idofs), ptr(idofs), ptr(part), self.nparts, C.byref(h)))
        self._h = h
        self.on_device = False
        self.device = None

    @classmethod
    def from_export(cls, hier, parameters=None, recolor=False):
        """Build a handle from an exported hierarchy (the dict `export()` returns, a golden fixture, or the
        arrays of an external CPU setup such as HAZmath's): mamg_import_hierarchy.  `parameters` overrides
        hier["params"]; recolor=True lets the library compute the Gauss-Seidel colouring itself."""
        from ._capi import MamgLevelArrays
        self = cls.__new__(cls)
        prm = parameters if parameters is not None else hier["params"]
        self.params = to_struct(prm)
        levels = hier["levels"]
        recs = (MamgLevelArrays * len(levels))()
        keep = []

        def arr(L, key, dtype):
            if key not in L or L[key] is None:
                return None
            a = np.ascontiguousarray(L[key], dtype=dtype)
            keep.append(a)
            return a.ctypes.data_as(C.c_void_p)

        nparts = 1
        for l, L in enumerate(levels):
            r = recs[l]
            last = l == len(levels) - 1
            r.n = int(L["n"])
            r.n_aggregates = 0 if last else int(L["n_aggregates"])
            npatch = len(L["patch_ptr"]) - 1 if "patch_ptr" in L and L["patch_ptr"] is not None else 0
            r.n_patches = npatch
            r.n_patch_colors = int(L.get("n_patch_colors", 0)) if npatch else 0
            r.indptr, r.indices, r.data = arr(L, "indptr", np.int32), arr(L, "indices", np.int32), arr(L, "data", np.float64)
            r.agg = None if last else arr(L, "agg", np.int32)
            if not recolor and not last:
                r.color, r.n_colors = arr(L, "color", np.int32), int(L["n_colors"])
            r.gs_skip = arr(L, "gs_skip", np.uint8) if npatch else None
            if npatch:
                r.patch_ptr, r.patch_dofs = arr(L, "patch_ptr", np.int32), arr(L, "patch_dofs", np.int32)
                r.patch_seed, r.patch_color = arr(L, "patch_seed", np.int32), arr(L, "patch_color", np.int32)
            if "P_indptr" in L:
                r.P_indptr, r.P_indices, r.P_data = arr(L, "P_indptr", np.int32), arr(L, "P_indices", np.int32), arr(L, "P_data", np.float64)
            if "part" in L and L["part"] is not None and np.asarray(L["part"]).max(initial=0) > 0:
                nparts = max(nparts, int(np.asarray(L["part"]).max()) + 1)
        if nparts > 1:
            for l, L in enumerate(levels):
                recs[l].part = arr(L, "part", np.int32)
        inv = hier.get("coarse_inv")
        inv = np.ascontiguousarray(inv, np.float64) if inv is not None else None
        h = C.c_void_p()
        check(lib.mamg_import_hierarchy(C.byref(self.params), len(levels), recs, ptr(inv), nparts, C.byref(h)))
        self._h = h
        self.n = int(levels[0]["n"])
        self.nparts = nparts
        self.idofs = np.zeros(0, np.int32)
        self.on_device = False
        self.device = None
        return self

    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            lib.mamg_destroy(h)
            self._h = None

    # ---- introspection -------------------------------------------------------------------
    @property
    def num_levels(self):
        v = C.c_int32()
        check(lib.mamg_num_levels(self._h, C.byref(v)))
        return v.value

    @property
    def setup_seconds(self):
        v = C.c_double()
        check(lib.mamg_setup_seconds(self._h, C.byref(v)))
        return v.value

    def level_info(self, level):
        if getattr(self, "_infos", None):
            return self._infos[level]
        info = (C.c_int64 * 12)()
        check(lib.mamg_level_info(self._h, level, info))
        keys = ["rows", "nnz", "n_aggregates", "n_colors", "n_patches", "n_patch_entries",
                "n_patch_colors", "max_patch_size", "patch_row_entries", "patch_inv_entries", "nnz_P",
                "nnz_structural"]
        return dict(zip(keys, [int(x) for x in info]))

    def export_level(self, level):
        """Natural-ordering arrays of one level (what the oracle consumes)."""
        info = self.level_info(level)
        n, nnz = info["rows"], info["nnz"]
        out = {
            "n": n, "indptr": np.empty(n + 1, np.int32), "indices": np.empty(nnz, np.int32),
            "data": np.empty(nnz, np.float64), "agg": np.empty(n, np.int32),
            "color": np.empty(n, np.int32), "gs_skip": np.empty(n, np.uint8),
            "n_aggregates": info["n_aggregates"], "n_colors": info["n_colors"],
            "part": np.empty(n, np.int32),
        }
        check(lib.mamg_part_export(self._h, level, ptr(out["part"])))
        check(lib.mamg_level_export(self._h, level, ptr(out["indptr"]), ptr(out["indices"]),
                                    ptr(out["data"]), ptr(out["agg"]), ptr(out["color"]),
                                    ptr(out["gs_skip"])))
        if info["nnz_P"]:
            out["P_indptr"] = np.empty(n + 1, np.int32)
            out["P_indices"] = np.empty(info["nnz_P"], np.int32)
            out["P_data"] = np.empty(info["nnz_P"], np.float64)
            check(lib.mamg_prolongator_export(self._h, level, ptr(out["P_indptr"]), ptr(out["P_indices"]),
                                              ptr(out["P_data"])))
        npatch = info["n_patches"]
        out["patch_ptr"] = np.zeros(npatch + 1, np.int32)
        out["patch_dofs"] = np.empty(info["n_patch_entries"], np.int32)
        out["patch_seed"] = np.empty(npatch, np.int32)
        out["patch_color"] = np.empty(npatch, np.int32)
        out["n_patch_colors"] = info["n_patch_colors"]
        if npatch:
            check(lib.mamg_schwarz_export(self._h, level, ptr(out["patch_ptr"]), ptr(out["patch_dofs"]),
                                          ptr(out["patch_seed"]), ptr(out["patch_color"])))
        return out

    def export(self):
        levels = [self.export_level(l) for l in range(self.num_levels)]
        nc = levels[-1]["n"]
        inv = np.empty((nc, nc), np.float64)
        check(lib.mamg_coarse_export(self._h, ptr(inv)))
        from .params import struct_to_dict
        return {"levels": levels, "coarse_inv": inv, "params": struct_to_dict(self.params)}

    def cycle_bytes(self):
        v = C.c_int64()
        check(lib.mamg_cycle_bytes(self._h, C.byref(v)))
        return v.value

    # ---- device ------------------------------------------------------------------------------
    def to_device(self, device=0, stream=None, rank=None, world=None):
        """Upload to one GPU.  With rank/world (one process per GPU) the rank keeps only the matrix rows of
        its own parts on the row-distributed levels (mamg_to_device_dist); dist_init() follows."""
        if world is not None and world > 1:
            check(lib.mamg_to_device_dist(self._h, int(device), C.c_void_p(stream) if stream else None, int(rank), int(world)))
        else:
            check(lib.mamg_to_device(self._h, int(device), C.c_void_p(stream) if stream else None))
        self.on_device = True
        self.device = int(device)
        return self

    def dist_init(self, rank=None, world=None, group=None):
        """Join the NCCL communicator of a torchrun job (one process per GPU).  Needs an initialised
        torch.distributed process group: rank 0 creates the NCCL id, it is broadcast as a byte tensor."""
        self._require_device()
        if world is None:
            import torch.distributed as dist
            if dist.is_available() and dist.is_initialized():
                rank, world = dist.get_rank(group), dist.get_world_size(group)
            else:
                rank, world = 0, 1
        buf = (C.c_ubyte * 128)()
        if world > 1:
            if rank == 0:
                check(lib.mamg_nccl_unique_id(buf))
            buf = (C.c_ubyte * 128)(*broadcast_bytes(bytes(buf), 0, group, self.device))
        check(lib.mamg_dist_init(self._h, int(rank), int(world), buf))
        self.rank, self.world = int(rank), int(world)
        if world > 1:
            # peer-memory exchange: all-gather the CUDA IPC handles of the vector arenas
            import torch.distributed as dist
            mine = (C.c_ubyte * 64)()
            check(lib.mamg_ipc_handle(self._h, mine))
            handles = allgather_bytes(bytes(mine), group, self.device)
            flat = (C.c_ubyte * (64 * world))(*b"".join(handles))
            check(lib.mamg_dist_peers(self._h, flat))
            dist.barrier(group=group)
        return self

    def collective_count(self, reset=False):
        v = C.c_int64()
        check(lib.mamg_collective_count(self._h, C.byref(v), int(reset)))
        return v.value

    def exchange_bytes(self, reset=False):
        v = C.c_int64()
        check(lib.mamg_exchange_bytes(self._h, C.byref(v), int(reset)))
        return v.value

    def set_stream(self, stream):
        check(lib.mamg_set_stream(self._h, C.c_void_p(stream) if stream else None))

    def set_cycle(self, cycle_type):
        """Switch V_CYCLE / W_CYCLE on the existing hierarchy (host and device)."""
        check(lib.mamg_set_cycle(self._h, int(cycle_type)))
        self.params.cycle_type = int(cycle_type)

    def release_host(self):
        """Free the host copy of the level matrices (after to_device); export() is no longer possible."""
        self._infos = [self.level_info(l) for l in range(self.num_levels)]
        check(lib.mamg_release_host(self._h))

    def sync(self):
        check(lib.mamg_sync(self._h))

    def device_bytes(self):
        v = C.c_int64()
        check(lib.mamg_device_bytes(self._h, C.byref(v)))
        return v.value

    def launch_count(self, reset=False):
        v = C.c_int64()
        check(lib.mamg_launch_count(self._h, C.byref(v), int(reset)))
        return v.value

    KERNEL_CLASSES = ["spmv", "gs", "schwarz", "restrict", "scale", "prolong", "coarse", "vector", "dot", "exchange"]

    def schwarz_sweep_bytes(self, level=0):
        v = C.c_int64()
        check(lib.mamg_schwarz_sweep_bytes(self._h, level, C.byref(v)))
        return v.value

    STAT_KEYS = ["rows", "nnz_stored", "nnz_structural", "sell_slots", "device_bytes", "n_patches", "unique_blobs",
                 "schwarz_sweep_bytes", "schwarz_sweep_bytes_stored_factors", "n_colors", "n_patch_colors", "in_tail",
                 "sell", "csr_kept", "row_blocks", "schwarz_fast_path", "max_patch_size", "max_patch_nbr"]

    def stats(self, level=0):
        """Device-side statistics of one level (mamg_stats)."""
        self._require_device()
        out = (C.c_int64 * 24)()
        check(lib.mamg_stats(self._h, int(level), out))
        return dict(zip(self.STAT_KEYS, [int(v) for v in out]))

    def race_check(self):
        """(gs_conflicts, patch_conflicts, launches_checked) of the device layout: mamg_race_check."""
        self._require_device()
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.mamg_race_check(self._h, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def profile_start(self):
        check(lib.mamg_profile(self._h, 1, None, None))

    def profile_levels(self):
        """ms[level][class] of everything launched since profile_start() (call before profile_stop)."""
        L = self.num_levels
        ms = np.zeros((L, 16), np.float64)
        check(lib.mamg_profile_levels(self._h, ptr(ms), L))
        return ms[:, :len(self.KERNEL_CLASSES)]

    def profile_stop(self):
        """{class: (ms, launches)} of everything launched since profile_start()."""
        ms = np.zeros(16, np.float64)
        cnt = np.zeros(16, np.int64)
        check(lib.mamg_profile(self._h, 0, ptr(ms), ptr(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.KERNEL_CLASSES)}

    def _require_device(self):
        if not self.on_device:
            self.to_device(0)

    @staticmethod
    def _is_torch_cuda(x):
        return hasattr(x, "is_cuda") and x.is_cuda

    def _vec_call(self, fn, ins, n_out, *scalars):
        """Run fn(handle, in..., out, scalars..., on_device) for numpy or torch-CUDA vectors."""
        self._require_device()
        if self._is_torch_cuda(ins[0]):
            import torch
            ins = [v.contiguous() for v in ins]
            for v in ins:
                if v.dtype != torch.float64:
                    raise TypeError("device vectors must be float64")
            out = torch.empty(n_out, dtype=torch.float64, device=ins[0].device)
            torch.cuda.current_stream(ins[0].device).synchronize()
            check(fn(self._h, *[C.c_void_p(v.data_ptr()) for v in ins], C.c_void_p(out.data_ptr()), *scalars, 1))
            check(lib.mamg_sync(self._h))  # the library runs on its own stream
            return out
        ins = [as_f64(v) for v in ins]
        out = np.empty(n_out, np.float64)
        check(fn(self._h, *[ptr(v) for v in ins], ptr(out), *scalars, 0))
        return out

    @staticmethod
    def _check_len(v, n, what="vector"):
        """The C-ABI takes bare pointers: a wrong length must be caught on this side."""
        shape = tuple(v.shape) if hasattr(v, "shape") else (len(v),)
        if len(shape) != 1 or shape[0] != n:
            raise ValueError(f"{what} has shape {shape}, expected ({n},)")

    def apply(self, r):
        """z = B r: one multigrid cycle (mamg_apply)."""
        self._check_len(r, self.n)
        return self._vec_call(lib.mamg_apply, [r], self.n)

    def _block_args(self, blocks, writable=False):
        """(nblocks, sizes, pointer array, on_device, kept-alive arrays) of a list of per-field vectors."""
        dev = self._is_torch_cuda(blocks[0])
        if dev:
            import torch
            vs = [v.contiguous() for v in blocks]
            for v in vs:
                if v.dtype != torch.float64 or not v.is_cuda:
                    raise TypeError("device blocks must be float64 CUDA tensors")
            ptrs = (C.c_void_p * len(vs))(*[v.data_ptr() for v in vs])
        else:
            vs = [np.array(v, np.float64, copy=True) if writable else as_f64(v) for v in blocks]
            ptrs = (C.c_void_p * len(vs))(*[v.ctypes.data for v in vs])
        sizes = (C.c_int32 * len(vs))(*[int(v.shape[0]) for v in vs])
        if sum(sizes) != self.n:
            raise ValueError(f"block sizes {list(sizes)} do not add up to {self.n}")
        return len(vs), sizes, ptrs, dev, vs

    def apply_blocks(self, r_blocks):
        """z = B r for a block_vec r (the R.T * Minv * R of src/utils.py:53) without concatenating the
        blocks: mamg_apply_blocks addresses them by offsets in its boundary kernels."""
        self._require_device()
        nb, sizes, rp, dev, rs = self._block_args(r_blocks)
        if dev:
            import torch
            zs = [torch.empty_like(v) for v in rs]
            torch.cuda.current_stream(rs[0].device).synchronize()
            zp = (C.c_void_p * nb)(*[v.data_ptr() for v in zs])
        else:
            zs = [np.empty_like(v) for v in rs]
            zp = (C.c_void_p * nb)(*[v.ctypes.data for v in zs])
        check(lib.mamg_apply_blocks(self._h, nb, sizes, rp, zp, int(dev)))
        if dev:
            check(lib.mamg_sync(self._h))
        return zs

    def pcg_blocks(self, b_blocks, x0_blocks=None, tolerance=1e-8, relative=False, maxiter=500):
        """cbc.block ConjGrad on a block system: right-hand side and solution as lists of blocks."""
        self._require_device()
        nb, sizes, bp, dev, bs = self._block_args(b_blocks)
        if dev:
            import torch
            xs = [torch.zeros_like(v) for v in bs] if x0_blocks is None else [v.clone().contiguous() for v in x0_blocks]
            torch.cuda.current_stream(bs[0].device).synchronize()
            xp = (C.c_void_p * nb)(*[v.data_ptr() for v in xs])
        else:
            xs = [np.zeros_like(v) for v in bs] if x0_blocks is None else [np.array(v, np.float64, copy=True) for v in x0_blocks]
            xp = (C.c_void_p * nb)(*[v.ctypes.data for v in xs])
        res = np.zeros(maxiter + 1, np.float64)
        al = np.zeros(max(maxiter, 1), np.float64)
        be = np.zeros(max(maxiter, 1), np.float64)
        nit = C.c_int32()
        rc = lib.mamg_pcg_blocks(self._h, nb, sizes, bp, xp, tolerance, int(relative), maxiter, int(x0_blocks is not None),
                                 int(dev), C.byref(nit), ptr(res), ptr(al), ptr(be))
        check(rc, allow=(1,))
        k = nit.value
        return xs, {"niters": k, "residuals": res[:k + 1].tolist(), "alphas": al[:k].tolist(),
                    "betas": be[:k].tolist(), "breakdown": rc == 1}

    def spmv(self, x, level=0):
        n = self.level_info(level)["rows"]
        self._check_len(x, n)
        self._require_device()
        x = as_f64(x)
        y = np.empty(n, np.float64)
        check(lib.mamg_spmv(self._h, level, ptr(x), ptr(y), 0))
        return y

    def smooth(self, b, x, level=0, post=False):
        n = self.level_info(level)["rows"]
        self._check_len(b, n, "b")
        self._check_len(x, n, "x")
        self._require_device()
        b = as_f64(b)
        x = as_f64(x).copy()
        check(lib.mamg_smooth(self._h, level, ptr(b), ptr(x), int(post), 0))
        return x

    def pcg(self, b, x0=None, tolerance=1e-8, relative=False, maxiter=500):
        """cbc.block ConjGrad on the device; returns (x, info)."""
        self._check_len(b, self.n, "b")
        if x0 is not None:
            self._check_len(x0, self.n, "x0")
        if maxiter < 0:
            raise ValueError("maxiter must be >= 0")
        self._require_device()
        res = np.zeros(maxiter + 1, np.float64)
        al = np.zeros(max(maxiter, 1), np.float64)
        be = np.zeros(max(maxiter, 1), np.float64)
        nit = C.c_int32()
        if self._is_torch_cuda(b):
            import torch
            b = b.contiguous()
            x = torch.zeros_like(b) if x0 is None else x0.clone().contiguous()
            torch.cuda.current_stream(b.device).synchronize()
            rc = lib.mamg_pcg(self._h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), tolerance,
                              int(relative), maxiter, int(x0 is not None), 1, C.byref(nit), ptr(res),
                              ptr(al), ptr(be))
        else:
            b = as_f64(b)
            x = np.zeros(self.n, np.float64) if x0 is None else as_f64(x0).copy()
            rc = lib.mamg_pcg(self._h, ptr(b), ptr(x), tolerance, int(relative), maxiter,
                              int(x0 is not None), 0, C.byref(nit), ptr(res), ptr(al), ptr(be))
        check(rc, allow=(1,))
        k = nit.value
        info = {"niters": k, "residuals": res[:k + 1].tolist(), "alphas": al[:k].tolist(),
                "betas": be[:k].tolist(), "breakdown": rc == 1}
        return x, info

    def _krylov(self, fn, b, tolerance, relative, maxiter, *extra):
        self._check_len(b, self.n, "b")
        if maxiter < 0:
            raise ValueError("maxiter must be >= 0")
        self._require_device()
        res = np.zeros(maxiter + 2, np.float64)
        nit = C.c_int32()
        if self._is_torch_cuda(b):
            import torch
            b = b.contiguous()
            x = torch.zeros_like(b)
            torch.cuda.current_stream(b.device).synchronize()
            rc = fn(self._h, C.c_void_p(b.data_ptr()), C.c_void_p(x.data_ptr()), tolerance, int(relative),
                    maxiter, *extra, 1, C.byref(nit), ptr(res))
        else:
            b = as_f64(b)
            x = np.zeros(self.n, np.float64)
            rc = fn(self._h, ptr(b), ptr(x), tolerance, int(relative), maxiter, *extra, 0, C.byref(nit), ptr(res))
        check(rc, allow=(1,))
        return x, {"niters": nit.value, "residuals": res[:nit.value + 1].tolist(), "breakdown": rc == 1}

    def minres(self, b, tolerance=1e-8, relative=False, maxiter=500):
        """Preconditioned MINRES on the device (mamg_minres); residuals are B-norm estimates."""
        return self._krylov(lib.mamg_minres, b, tolerance, relative, maxiter)

    def gmres(self, b, tolerance=1e-8, relative=False, maxiter=500, restart=30):
        """Right-preconditioned restarted GMRES on the device (mamg_gmres); residuals are ||b - A x||_2."""
        return self._krylov(lib.mamg_gmres, b, tolerance, relative, maxiter, int(restart))
