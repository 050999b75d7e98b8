"""caaa_b200 -- B200-native rollout path for cseelhoff/caaa (Axis & Allies MCTS playouts).

The product is the CUDA library ``libcaaa_gpu.so`` (C-ABI in include/caaa_gpu.h, sources in
caaa_b200/csrc).  This package is a thin ctypes binding over that ABI plus workload generators; it
contains no game logic and no CPU fallback: importing :mod:`caaa_b200.gpu` without the built
library, or initialising it without a CUDA device, raises.
"""
from .layout import (CACHE_DTYPE, STATE_BYTES, STATE_DTYPE, STATS_DTYPE, SUMMARY_DTYPE, TABLES_DTYPE,
                     start_state)

__all__ = ["CACHE_DTYPE", "STATE_BYTES", "STATE_DTYPE", "STATS_DTYPE", "SUMMARY_DTYPE", "TABLES_DTYPE", "start_state"]
