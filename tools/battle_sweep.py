import os, sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from caaa_b200 import gpu, synth
eng = gpu.GpuEngine(0)
one = synth.battle_states(1 << 20, seed=0xBA77)
d_in = torch.from_numpy(one).cuda(); d_out = torch.empty_like(d_in); d_draws = torch.empty(1 << 20, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for div in (1, 2, 3, 4, 6, 8, 16):
    os.environ['CAAA_BATTLE_YIELD_DIV'] = str(div)
    for _ in range(2):
        eng.resolve_battles_device(d_in.data_ptr(), 1 << 20, 0xBA77, 0, d_out.data_ptr(), d_draws.data_ptr(), None, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(4):
        eng.resolve_battles_device(d_in.data_ptr(), 1 << 20, 0xBA77, r << 20, d_out.data_ptr(), d_draws.data_ptr(), None, st)
    e1.record(); torch.cuda.synchronize()
    print(f"yield_div={div}: {4 * (1 << 20) / e0.elapsed_time(e1) * 1e3 / 1e6:.1f} M battles/s", flush=True)
