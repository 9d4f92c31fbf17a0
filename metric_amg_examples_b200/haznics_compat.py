"""Constants namespace standing in for the `haznics` SWIG module.

The reference's src/amg_parameters.py does `import haznics` and uses only the
integer macros below (src/amg_parameters.py:5-16,42; src/input_metric.dat:68-99).
With `import metric_amg_examples_b200.haznics_compat as haznics` that file runs
unchanged.  Values mirror include/mamg.h; only symbolic use matters.
"""
UA_AMG, SA_AMG = 1, 2
V_CYCLE, W_CYCLE, AMLI_CYCLE, NL_AMLI_CYCLE, ADD_CYCLE = 1, 2, 3, 4, 5
SMOOTHER_JACOBI, SMOOTHER_GS, SMOOTHER_SGS = 1, 2, 3
SMOOTHER_SOR, SMOOTHER_SSOR, SMOOTHER_L1DIAG = 5, 6, 10
VMB, MIS, MWM, HEC, HEM = 1, 2, 3, 4, 5
SCHWARZ_FORWARD, SCHWARZ_BACKWARD, SCHWARZ_SYMMETRIC = 1, 2, 3
SOLVER_DEFAULT, SOLVER_VFGMRES, SOLVER_GCG, SOLVER_GCR, SOLVER_UMFPACK = 0, 4, 5, 6, 32
OFF, ON = 0, 1
