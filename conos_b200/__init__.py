"""conos_b200 -- B200-native joint-graph construction (the Conos$buildGraph hot path).

Python host mirror of the reference's R interface over the C ABI of libconosgpu.so.  Importing the package
does not need a GPU; creating a context (or calling any operator) does, and fails loudly without one.
"""
from ._lib import ConosGpuError, Context, Job, default_context, load  # noqa: F401
from .conos import Conos, ConosGraph, Pagoda2  # noqa: F401
