"""gpu-pruner idle-decision engine for B200 (sm_100a).

The product is ``libgpr.so`` (``csrc/``: hand-written CUDA + the C ABI of ``include/gpr.h``).
This package holds the Python binding used by the tests and ``bench.py``; the C++ host mirror
of the reference controller lives under ``host/``.
"""
from . import ffi
from .engine import Decision, GprError, IdleEngine, from_biased_u8, to_biased_u8
from . import sharding
from .sharding import Shard, shard_pods

__all__ = ["ffi", "Decision", "GprError", "IdleEngine", "from_biased_u8", "to_biased_u8", "Shard", "shard_pods"]
