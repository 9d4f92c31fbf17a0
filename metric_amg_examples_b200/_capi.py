"""ctypes binding of libmamg.so (include/mamg.h).  There is no fallback: if the shared
object is missing the import fails loudly and tells how to build it."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MAMG_LIB") or os.path.join(_HERE, "libmamg.so")

PARAM_FIELDS = [
    ("AMG_type", C.c_int32), ("cycle_type", C.c_int32), ("max_levels", C.c_int32),
    ("maxit", C.c_int32), ("smoother", C.c_int32), ("relaxation", C.c_double),
    ("presmooth_iter", C.c_int32), ("postsmooth_iter", C.c_int32), ("coarse_dof", C.c_int32),
    ("coarse_solver", C.c_int32), ("coarse_scaling", C.c_int32), ("aggregation_type", C.c_int32),
    ("strong_coupled", C.c_double), ("max_aggregation", C.c_int32), ("amli_degree", C.c_int32),
    ("Schwarz_levels", C.c_int32), ("Schwarz_mmsize", C.c_int32), ("Schwarz_maxlvl", C.c_int32),
    ("Schwarz_type", C.c_int32), ("Schwarz_blksolver", C.c_int32), ("print_level", C.c_int32),
    ("nl_amli_krylov_type", C.c_int32), ("reserved", C.c_int32 * 7),
]


class MamgParams(C.Structure):
    _fields_ = PARAM_FIELDS


class MamgLevelArrays(C.Structure):
    """mamg_level_arrays of include/mamg.h (mamg_import_hierarchy)."""
    _fields_ = [(k, C.c_int32) for k in ("n", "n_aggregates", "n_colors", "n_patches", "n_patch_colors", "reserved")] + \
               [(k, C.c_void_p) for k in ("indptr", "indices", "data", "agg", "color", "gs_skip", "patch_ptr", "patch_dofs",
                                          "patch_seed", "patch_color", "P_indptr", "P_indices", "P_data", "part")]


def _load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python metric_amg_examples_b200/build.py` "
            "(nvcc, sm_100a). This package has no CPU or pure-Python fallback.")
    # libmamg.so needs libnccl.so.2.  PyTorch bundles a newer NCCL under the same soname and fails to
    # import if an older one is already mapped, so map PyTorch's copy first when it exists (the
    # system library is the fallback; the few entry points used here are stable across both).
    try:
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                C.CDLL(cand, mode=C.RTLD_GLOBAL)
                break
    except OSError:
        pass
    lib = C.CDLL(LIB_PATH)
    i32, i64, dbl, vp = C.c_int32, C.c_int64, C.c_double, C.c_void_p
    pi32, pi64, pdbl, pu8 = C.POINTER(i32), C.POINTER(i64), C.POINTER(dbl), C.POINTER(C.c_uint8)
    sig = {
        "mamg_last_error": (C.c_char_p, []),
        "mamg_version": (C.c_char_p, []),
        "mamg_params_default": (i32, [C.POINTER(MamgParams)]),
        "mamg_setup": (i32, [C.POINTER(MamgParams), i32, vp, vp, vp, i32, vp, C.POINTER(vp)]),
        "mamg_setup_partitioned": (i32, [C.POINTER(MamgParams), i32, vp, vp, vp, i32, vp, vp, i32, C.POINTER(vp)]),
        "mamg_part_export": (i32, [vp, i32, vp]),
        "mamg_import_hierarchy": (i32, [C.POINTER(MamgParams), i32, C.POINTER(MamgLevelArrays), vp, i32, C.POINTER(vp)]),
        "mamg_destroy": (i32, [vp]),
        "mamg_num_levels": (i32, [vp, pi32]),
        "mamg_level_info": (i32, [vp, i32, pi64]),
        "mamg_level_export": (i32, [vp, i32, vp, vp, vp, vp, vp, vp]),
        "mamg_schwarz_export": (i32, [vp, i32, vp, vp, vp, vp]),
        "mamg_prolongator_export": (i32, [vp, i32, vp, vp, vp]),
        "mamg_coarse_export": (i32, [vp, vp]),
        "mamg_setup_seconds": (i32, [vp, pdbl]),
        "mamg_to_device": (i32, [vp, i32, vp]),
        "mamg_to_device_dist": (i32, [vp, i32, vp, i32, i32]),
        "mamg_set_stream": (i32, [vp, vp]),
        "mamg_nccl_unique_id": (i32, [vp]),
        "mamg_dist_init": (i32, [vp, i32, i32, vp]),
        "mamg_collective_count": (i32, [vp, pi64, i32]),
        "mamg_exchange_bytes": (i32, [vp, pi64, i32]),
        "mamg_ipc_handle": (i32, [vp, vp]),
        "mamg_dist_peers": (i32, [vp, vp]),
        "mamg_device_bytes": (i32, [vp, pi64]),
        "mamg_set_cycle": (i32, [vp, i32]),
        "mamg_sync": (i32, [vp]),
        "mamg_release_host": (i32, [vp]),
        "mamg_apply": (i32, [vp, vp, vp, i32]),
        "mamg_apply_blocks": (i32, [vp, i32, vp, vp, vp, i32]),
        "mamg_pcg_blocks": (i32, [vp, i32, vp, vp, vp, dbl, i32, i32, i32, i32, pi32, vp, vp, vp]),
        "mamg_spmv": (i32, [vp, i32, vp, vp, i32]),
        "mamg_smooth": (i32, [vp, i32, vp, vp, i32, i32]),
        "mamg_pcg": (i32, [vp, vp, vp, dbl, i32, i32, i32, i32, pi32, vp, vp, vp]),
        "mamg_minres": (i32, [vp, vp, vp, dbl, i32, i32, i32, pi32, vp]),
        "mamg_gmres": (i32, [vp, vp, vp, dbl, i32, i32, i32, i32, pi32, vp]),
        "mamg_launch_count": (i32, [vp, pi64, i32]),
        "mamg_profile": (i32, [vp, i32, vp, vp]),
        "mamg_profile_levels": (i32, [vp, vp, i32]),
        "mamg_schwarz_sweep_bytes": (i32, [vp, i32, pi64]),
        "mamg_cycle_bytes": (i32, [vp, pi64]),
        "mamg_race_check": (i32, [vp, pi64, pi64, pi64]),
        "mamg_stats": (i32, [vp, i32, pi64]),
        "mamg_assemble_scalar": (i32, [i32, vp, vp, dbl, dbl, pi64, pi64, vp, vp, vp]),
        "mamg_assemble_bidomain": (i32, [i32, i32, dbl, dbl, dbl, pi64, pi64, vp, vp, vp]),
        "mamg_assemble_emi": (i32, [i32, i32, dbl, dbl, dbl, pi64, pi64, vp, vp, vp]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here == header/library mismatch
        fn.restype = res
        fn.argtypes = args
    lib._mamg_symbols = sorted(sig)
    return lib


lib = _load()
SYMBOLS = lib._mamg_symbols


class MamgError(RuntimeError):
    pass


def check(rc, allow=()):
    if rc != 0 and rc not in allow:
        raise MamgError(lib.mamg_last_error().decode())
    return rc


def ptr(a):
    """void* of a numpy array (None -> NULL) or pass-through of an int device pointer."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    return a.ctypes.data_as(C.c_void_p)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def as_i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
