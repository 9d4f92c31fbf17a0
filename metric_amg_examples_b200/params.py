"""Parameter dicts of the reference and their translation to the C struct.

The four dicts restate src/amg_parameters.py:3-89 key for key (values via the
haznics-compatible constants); `default_metric_parameters` is the inline dict of
src/utils.py:60-82 and `default_amg_parameters` the one of src/utils.py:20-38.
`to_struct` plays the role of haznics.param_amg_set_dict: unknown keys warn,
values outside the implemented option space raise NotImplementedError (never a silent fallback).
"""
import warnings

from . import haznics_compat as haznics
from ._capi import MamgParams, lib

parameters_standard = {
    "prectype": 2,
    "AMG_type": haznics.UA_AMG,
    "cycle_type": haznics.W_CYCLE,
    "max_levels": 20,
    "maxit": 1,
    "smoother": haznics.SMOOTHER_SGS,
    "relaxation": 1.2,
    "presmooth_iter": 1,
    "postsmooth_iter": 1,
    "coarse_dof": 100,
    "coarse_solver": 32,
    "coarse_scaling": haznics.ON,
    "aggregation_type": haznics.VMB,
    "strong_coupled": 0.1,
    "max_aggregation": 100,
    "Schwarz_levels": 0,
    "print_level": 10,
}

parameters_standard_schwarz = dict(
    parameters_standard,
    Schwarz_levels=1,
    Schwarz_mmsize=100,
    Schwarz_maxlvl=1,
    Schwarz_type=haznics.SCHWARZ_SYMMETRIC,
    Schwarz_blksolver=32,
    print_level=5,
)

parameters_metric = {
    "AMG_type": haznics.UA_AMG,
    "cycle_type": haznics.W_CYCLE,
    "max_levels": 20,
    "maxit": 1,
    "smoother": haznics.SMOOTHER_SGS,
    "relaxation": 1.2,
    "presmooth_iter": 1,
    "postsmooth_iter": 1,
    "coarse_dof": 100,
    "coarse_solver": 32,
    "coarse_scaling": haznics.ON,
    "aggregation_type": haznics.HEM,
    "strong_coupled": 0.1,
    "max_aggregation": 100,
    "amli_degree": 3,
    "Schwarz_levels": 0,
    "print_level": 5,
}

parameters_metric_schwarz = dict(
    parameters_metric,
    Schwarz_levels=1,
    Schwarz_mmsize=100,
    Schwarz_maxlvl=1,
    Schwarz_type=haznics.SCHWARZ_SYMMETRIC,
    Schwarz_blksolver=32,
)

# src/utils.py:60-82 (used when get_hazmath_metric_precond* gets parameters=None)
default_metric_parameters = dict(parameters_metric_schwarz, Schwarz_maxlvl=2, print_level=10)
# src/utils.py:20-38
default_amg_parameters = dict(parameters_standard)

_STRUCT_KEYS = [f[0] for f in MamgParams._fields_ if f[0] != "reserved"]
_IGNORED_KEYS = {"prectype"}  # accepted by HAZmath's generic precond struct, irrelevant for AMG


def to_struct(parameters=None):
    p = MamgParams()
    lib.mamg_params_default(p)
    if parameters is None:
        return p
    for key, val in parameters.items():
        if key in _IGNORED_KEYS:
            continue
        if key not in _STRUCT_KEYS:
            warnings.warn(f"AMG parameter '{key}' is not known to this implementation; ignored")
            continue
        setattr(p, key, type(getattr(p, key))(val))
    if p.AMG_type not in (haznics.UA_AMG, haznics.SA_AMG):
        raise NotImplementedError(f"AMG_type={p.AMG_type}")
    if p.cycle_type not in (haznics.V_CYCLE, haznics.W_CYCLE, haznics.AMLI_CYCLE, haznics.NL_AMLI_CYCLE,
                            haznics.ADD_CYCLE):
        raise NotImplementedError(f"cycle_type={p.cycle_type}")
    if p.cycle_type == haznics.ADD_CYCLE and p.maxit > 1:
        raise NotImplementedError("ADD_CYCLE: the additive cycle is applied once per call (maxit 1)")
    if p.cycle_type == haznics.AMLI_CYCLE and not 0 <= p.amli_degree <= 15:
        raise NotImplementedError(f"amli_degree={p.amli_degree}: 0..15")
    if p.aggregation_type not in (haznics.VMB, haznics.MIS, haznics.MWM, haznics.HEC, haznics.HEM):
        raise NotImplementedError(f"aggregation_type={p.aggregation_type}")
    if p.smoother not in (haznics.SMOOTHER_JACOBI, haznics.SMOOTHER_GS, haznics.SMOOTHER_SGS,
                          haznics.SMOOTHER_SOR, haznics.SMOOTHER_SSOR, haznics.SMOOTHER_L1DIAG):
        raise NotImplementedError(f"smoother={p.smoother}")
    # 0 = "iterative" upstream (src/amg_parameters.py:14,43): the coarsest operator / the patch blocks are
    # iterated to tol*1e-4 there; the dense inverses used here are the limit of that iteration
    if p.coarse_solver not in (haznics.SOLVER_UMFPACK, haznics.SOLVER_DEFAULT):
        raise NotImplementedError(f"coarse_solver={p.coarse_solver}: 32 (direct) or 0 (iterative)")
    if p.Schwarz_levels > 0 and p.Schwarz_blksolver not in (haznics.SOLVER_UMFPACK, haznics.SOLVER_DEFAULT):
        raise NotImplementedError(f"Schwarz_blksolver={p.Schwarz_blksolver}: 32 (direct) or 0 (iterative)")
    return p


def struct_to_dict(p):
    return {k: getattr(p, k) for k in _STRUCT_KEYS}
