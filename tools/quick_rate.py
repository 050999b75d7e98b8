#!/usr/bin/env python3
"""Quick kernel-rate probe: python tools/quick_rate.py N [N ...] -- playouts/s from the library's own CUDA-event timer."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout

eng = gpu.GpuEngine(0)
s0 = layout.start_state()
eng.rollout_batch(s0, 4096, 1, 0)
for n in [int(x) for x in sys.argv[1:]]:
    t0 = time.perf_counter()
    res, _, summ = eng.rollout_batch(s0, n, 0xCAAA2024, 0)
    print(f"n={n} kernel_ms={summ['kernel_ms']:.1f} playouts/s={n / (summ['kernel_ms'] * 1e-3):.0f} "
          f"plies={int(summ['plies'])} wall={time.perf_counter() - t0:.2f}s mean={res.mean():.4f}", flush=True)
