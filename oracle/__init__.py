"""CPU oracle of the metric-AMG apply path -- TEST INFRASTRUCTURE (see mamg_oracle.c header).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package.  PARITY UNPINNED: the reference ships no golden vectors for this path.
"""
from .oracle import Oracle, build_oracle  # noqa: F401
