#!/usr/bin/env python3
"""Consecutive rollout batches on one stream vs alternating streams (tools; not part of the product)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout

P = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
K = int(sys.argv[2]) if len(sys.argv) > 2 else 4
eng = gpu.GpuEngine(0)
dev = torch.device("cuda", 0)
d_state = torch.from_numpy(layout.start_state()).to(dev)
res = [torch.empty(P, dtype=torch.float64, device=dev) for _ in range(3)]
summ = torch.zeros(8, dtype=torch.int64, device=dev)
for n_streams in (1, 2, 3):
    streams = [torch.cuda.Stream() for _ in range(n_streams)]
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for st in streams:
            st.wait_event(e0)
        for k in range(K):
            st = streams[k % n_streams]
            eng.rollout_batch_device(d_state.data_ptr(), 1, P, 0xCAAA2024, (rep * K + k) * P, res[k % 3].data_ptr(), None, summ.data_ptr(), st.cuda_stream)
        for st in streams:
            torch.cuda.current_stream().wait_stream(st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
    eng.synchronize()
    print(f"streams={n_streams} steps={K} P={P}: {ms:.1f} ms total, {K * P / ms * 1e3 / 1e6:.3f} M playouts/s", flush=True)
