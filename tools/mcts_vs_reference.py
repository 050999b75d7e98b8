import sys, time, numpy as np
sys.path.insert(0, '/root/repo')
from caaa_b200 import gpu, host, layout
eng = gpu.GpuEngine(0)
s0 = layout.start_state()
ref = np.array([892, 4664, 5166, 3693, 5054, 531]) / 20000.0
for lpb, seed in ((32, 1), (32, 2), (8, 3), (128, 4)):
    t = time.time()
    a, v, x, best, info = host.mcts_search(eng, s0, 20000, 1, lpb, seed)
    sh = v / v.sum()
    print(lpb, seed, a.tolist(), np.round(sh, 3).tolist(), "ref", np.round(ref, 3).tolist(), "max|d|", round(float(np.abs(sh - ref).max()), 3), "best", best, "means", np.round(x / np.maximum(v, 1), 3).tolist(), f"{time.time()-t:.1f}s", flush=True)
