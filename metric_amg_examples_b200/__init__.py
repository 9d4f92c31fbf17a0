"""B200-native metric-AMG apply path (drop-in for the HAZniCS/HAZmath preconditioner path of
anabudisa/metric-amg-examples).  See DESIGN.md; the C-ABI is include/mamg.h."""
from . import haznics_compat  # noqa: F401
from .hierarchy import Hierarchy  # noqa: F401

__all__ = ["Hierarchy", "haznics_compat"]
