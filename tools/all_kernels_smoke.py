#!/usr/bin/env python3
"""Small invocations of every kernel through the C-ABI (tools; not part of the product): a quick all-kernels smoke run.
(Written for `compute-sanitizer --tool memcheck`, which is closed on this GPU pool.)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout, synth

eng = gpu.GpuEngine(0)
s0 = layout.start_state()
for kernel, n in (("stream", 6000), ("lockstep", 500), ("resident", 1500)):
    os.environ["CAAA_ROLLOUT_KERNEL"] = kernel
    res, stats, summ = eng.rollout_batch(s0, n, 0xCAAA2024, 0, want_stats=True)
    assert int(summ["playouts"]) == n
    print(kernel, n, float(res.mean()))
os.environ.pop("CAAA_ROLLOUT_KERNEL")
bs = synth.battle_states(777, seed=3)
out, draws, summ = eng.resolve_battles(bs, 5, 0)
print("battles", int(summ["playouts"]))
na, acts, ch, st = eng.expand(np.repeat(s0.reshape(1, -1), 5, axis=0))
print("expand", na.tolist())
pick = np.arange(5, dtype=np.uint64)
eng.leaf_batch_submit(0, np.repeat(s0.reshape(1, -1), 5, axis=0), 64, 1, 0, pick)
o = eng.leaf_batch_collect(0, 5)
print("leaf batch", o["picked"].tolist(), o["sums"].round(3).tolist())
st2, caches, d = eng.advance(s0, 1, 0, 3)
print("advance ok", int(d[0]))
