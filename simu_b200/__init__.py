"""simu_b200 -- B200-native phenotype-synthesis hot path of precimed/simu.

The product is ``lib/libsimu_b200.so`` (hand-written sm_100a kernels behind the
C ABI of ``include/simu_b200.h``) and the C++ command line in ``host/``.  The
Python modules here are the ctypes binding (``capi``), the mirror of the
reference's four hot-path functions (``hotpath``) and PLINK file helpers
(``plink``).  There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401
from .capi import MODE_EXACT, MODE_FAST, Context, SimuB200Error  # noqa: F401

__all__ = ["capi", "Context", "SimuB200Error", "MODE_EXACT", "MODE_FAST"]
