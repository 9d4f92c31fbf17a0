"""Preconditioner factories with the names and signatures of the reference's src/utils.py
(the L2 "drop-in boundary" of SURVEY section 1), so a driver line such as

    BB = utils.get_hazmath_metric_precond_mono(AA_, W, bcs, amgparams, interface_dofs)

(src/bidomain_2d.py:203) works against this package.  `bcs` is accepted and unused, exactly as
in the reference (src/utils.py:9,15,45,56).
"""
from .block import ReductionOperator, block_diag_mat, ii_convert
from .precond import AMG, LU, metricAMG


def get_block_diag_precond(A, W, bcs):
    """src/utils.py:9-12 -- 'Exact blocks LU as preconditioner' (`-precond diag` of src/emi_2d.py:149):
    block_diag_mat of one exact solve per diagonal block (see precond.LU for how the solve is done)."""
    n, = set(A.blocks.shape)
    return block_diag_mat([LU(A[i, i]) for i in range(n)])


def get_hazmath_amg_precond(A, W, bcs, parameters=None, interface_dofs=None):
    """src/utils.py:15-42 -- standard UA-AMG (VMB aggregation) on the monolithic matrix."""
    return AMG(A, parameters=parameters)


def get_hazmath_metric_precond_mono(A, W, bcs, parameters=None, interface_dofs=None):
    """src/utils.py:56-90 -- metric AMG on the monolithic matrix.  With interface_dofs the
    interface dofs get the Schwarz smoother and the rest Gauss-Seidel (src/utils.py:84)."""
    if interface_dofs is not None:
        return metricAMG(A, W, idofs=interface_dofs, parameters=parameters)
    return metricAMG(A, W, parameters=parameters)


def get_hazmath_metric_precond(A, W, bcs, parameters=None, interface_dofs=None):
    """src/utils.py:45-53 -- block variant R.T * Minv * R acting on 2-block vectors."""
    AA = ii_convert(A)
    R = ReductionOperator([len(W)], W)
    Minv = get_hazmath_metric_precond_mono(AA, W, bcs, parameters=parameters, interface_dofs=interface_dofs)
    return R.T * Minv * R
