"""Reader of HAZmath solver input files (the reference's src/input_metric.dat, consumed upstream by
haznics.fenics_metric_solver_xd_1d through HAZmath's input parser; src/run_solver_3d1d.py:17-38).

Format: `key = value % comment`, `%` starts a comment, symbolic values (SA, V, GS, OFF ...) are the
HAZmath macro names.  `read_input` returns (linear-solver settings, AMG parameter dict with the keys
of src/amg_parameters.py).
"""
from . import haznics_compat as haznics

_SYMBOLS = {
    "AMG_type": {"UA": haznics.UA_AMG, "SA": haznics.SA_AMG},
    "AMG_cycle_type": {"V": haznics.V_CYCLE, "W": haznics.W_CYCLE, "A": haznics.AMLI_CYCLE,
                       "NA": haznics.NL_AMLI_CYCLE, "ADD": haznics.ADD_CYCLE},
    "AMG_smoother": {"JACOBI": haznics.SMOOTHER_JACOBI, "GS": haznics.SMOOTHER_GS, "SGS": haznics.SMOOTHER_SGS,
                     "SOR": haznics.SMOOTHER_SOR, "SSOR": haznics.SMOOTHER_SSOR, "L1DIAG": haznics.SMOOTHER_L1DIAG},
    "AMG_coarse_scaling": {"OFF": haznics.OFF, "ON": haznics.ON},
}

# .dat key -> key of the parameter dicts (src/amg_parameters.py)
_AMG_KEYS = {
    "AMG_type": "AMG_type", "AMG_cycle_type": "cycle_type", "AMG_levels": "max_levels", "AMG_maxit": "maxit",
    "AMG_smoother": "smoother", "AMG_relaxation": "relaxation", "AMG_presmooth_iter": "presmooth_iter",
    "AMG_postsmooth_iter": "postsmooth_iter", "AMG_coarse_dof": "coarse_dof", "AMG_coarse_solver": "coarse_solver",
    "AMG_coarse_scaling": "coarse_scaling", "AMG_amli_degree": "amli_degree",
    "AMG_aggregation_type": "aggregation_type", "AMG_strong_coupled": "strong_coupled",
    "AMG_max_aggregation": "max_aggregation", "AMG_Schwarz_levels": "Schwarz_levels",
    "Schwarz_mmsize": "Schwarz_mmsize", "Schwarz_maxlvl": "Schwarz_maxlvl", "Schwarz_type": "Schwarz_type",
    "Schwarz_blksolver": "Schwarz_blksolver",
}


def _value(key, text):
    text = text.strip().rstrip(";").strip()
    if key in _SYMBOLS and text.upper() in _SYMBOLS[key]:
        return _SYMBOLS[key][text.upper()]
    try:
        return int(text)
    except ValueError:
        try:
            return float(text)
        except ValueError:
            return text


def parse(path):
    """All `key = value` pairs of the file as a dict."""
    out = {}
    with open(path) as fh:
        for line in fh:
            line = line.split("%", 1)[0].strip()
            if "=" not in line:
                continue
            key, val = line.split("=", 1)
            key = key.strip()
            if key and val.strip():
                out[key] = _value(key, val)
    return out


def read_input(path):
    """(solver, amg): solver = {'type': 'cg'|'minres'|'gmres', 'maxit', 'tol', 'stop_type', 'restart',
    'precond_type', 'print_level'}; amg = parameter dict for metricAMG."""
    raw = parse(path)
    kinds = {0: "direct", 1: "cg", 2: "minres", 3: "gmres"}
    solver = {
        "type": kinds.get(raw.get("linear_itsolver_type", 1), "cg"),
        "maxit": int(raw.get("linear_itsolver_maxit", 500)),
        "tol": float(raw.get("linear_itsolver_tol", 1e-6)),
        "stop_type": int(raw.get("linear_stop_type", 1)),
        "restart": int(raw.get("linear_restart", 30)),
        "precond_type": int(raw.get("linear_precond_type", 16)),
        "print_level": int(raw.get("print_level", 0)),
    }
    amg = {new: raw[old] for old, new in _AMG_KEYS.items() if old in raw}
    return solver, amg
