#!/usr/bin/env python3
"""Knob sweep of the rollout kernels (tools; not part of the product).

    python tools/sweep_knobs.py [playouts] [spec ...]     spec = "CAAA_PATIENCE=8,CAAA_BUCKET_MONEY=12"

One process; the library reads its knobs from the environment at every launch.  Kernel time comes from
the library's own CUDA events (CaaaBatchSummary.kernel_ms) around the launch.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout  # noqa: E402

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
specs = sys.argv[2:] or [""]
eng = gpu.GpuEngine(0)
s0 = layout.start_state()
eng.rollout_batch(s0, 1 << 16, 1, 0)
KNOBS = set()
for spec in specs:
    for kv in filter(None, spec.split(",")):
        KNOBS.add(kv.split("=")[0])
for spec in specs:
    for k in KNOBS:
        os.environ.pop(k, None)
    for kv in filter(None, spec.split(",")):
        k, v = kv.split("=")
        os.environ[k] = v
    best = None
    for rep in range(2):
        res, _, summ = eng.rollout_batch(s0, P, 0xCAAA2024, rep * P)
        ms = float(summ["kernel_ms"])
        best = ms if best is None else min(best, ms)
    print(f"{spec or 'default':60s} {P / best * 1e3 / 1e6:8.3f} M playouts/s  kernel {best:8.1f} ms  "
          f"plies/playout {float(summ['plies']) / P:.2f} mean {res.mean():.5f}", flush=True)
