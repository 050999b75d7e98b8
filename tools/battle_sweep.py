"""Battle kernels (config 3): the refill kernel (owning lanes per warp) against the pool kernel's configurations x yield
divisor; device-resident buffers, CUDA events; every variant must reproduce the first one's states and draw counts."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, synth
eng = gpu.GpuEngine(0)
N = 1 << 20
one = synth.battle_states(N, seed=0xBA77)
d_in = torch.from_numpy(one).cuda(); d_out = torch.empty_like(d_in); d_draws = torch.empty(N, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
ref = None
variants = [("refill", {"CAAA_BATTLE_LANES": "16"})] + [("pool", {"CAAA_BATTLE_POOL": str(c)}) for c in range(3)]
divs = [int(x) for x in (sys.argv[1].split(",") if len(sys.argv) > 1 else ["4"])]
for kern, env in variants:
    os.environ['CAAA_BATTLE_KERNEL'] = kern
    os.environ.update(env)
    for div in divs:
        os.environ['CAAA_BATTLE_YIELD_DIV'] = str(div)
        d_out.zero_(); d_draws.zero_()
        for _ in range(2):
            eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, 0, d_out.data_ptr(), d_draws.data_ptr(), None, st)
        torch.cuda.synchronize()
        if ref is None:
            ref = (d_out.clone(), d_draws.clone())
        same = bool(torch.equal(d_out, ref[0]) and torch.equal(d_draws, ref[1]))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(4):
            eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, r << 20, d_out.data_ptr(), d_draws.data_ptr(), None, st)
        e1.record(); torch.cuda.synchronize()
        print(f"{kern} {env} yield_div={div}: {4 * N / e0.elapsed_time(e1) * 1e3 / 1e6:.1f} M battles/s  same={same}", flush=True)
