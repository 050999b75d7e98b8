#!/usr/bin/env python3
"""Latency of small rollout batches through the host-buffer C-ABI call (tools; not part of the product)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout

eng = gpu.GpuEngine(0)
s0 = layout.start_state()
eng.rollout_batch(s0, 1024, 1, 0)
for n_states, per in ((1, 1), (1, 256), (1, 4096), (6, 4096), (64, 4096), (1, 262144)):
    st = np.repeat(s0.reshape(1, -1), n_states, axis=0)
    t0 = time.perf_counter()
    res, _, summ = eng.rollout_batch(st, per, 7, 0)
    dt = time.perf_counter() - t0
    print(f"n_states={n_states} per={per} playouts={n_states*per} wall={dt*1e3:.1f} ms kernel={float(summ['kernel_ms']):.1f} ms "
          f"rate={n_states*per/dt:.0f}/s mean_plies={float(summ['plies'])/(n_states*per):.1f}", flush=True)
