#!/usr/bin/env python3
"""Latency of small rollout batches through the host-buffer C-ABI call, per kernel (tools; not part of the product)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout

eng = gpu.GpuEngine(0)
s0 = layout.start_state()
eng.rollout_batch(s0, 1024, 1, 0)
kernels = sys.argv[1:] or ["lockstep", "resident", "stream"]
for n_states, per in ((1, 1), (1, 256), (1, 4096), (6, 4096), (64, 4096), (1, 262144)):
    st = np.repeat(s0.reshape(1, -1), n_states, axis=0)
    for k in kernels:
        os.environ["CAAA_ROLLOUT_KERNEL"] = k.split(":")[0]
        if ":" in k:
            os.environ["CAAA_WARPS"] = k.split(":")[1]
        best = None
        for rep in range(3):
            t0 = time.perf_counter()
            res, _, summ = eng.rollout_batch(st, per, 7, 0)
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"{k:12s} n_states={n_states} per={per} playouts={n_states*per} wall={best*1e3:.2f} ms kernel={float(summ['kernel_ms']):.2f} ms "
              f"rate={n_states*per/best:.0f}/s", flush=True)
