"""Where the battle kernel's time goes: the shipped library against a build with -DCAAA_BATTLE_PROBE (layout conversion +
cache rebuild + export, no combat), with and without the state export.  usage: battle_probe.py [path/to/lib.so]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, synth
if len(sys.argv) > 1:
    gpu.load_library.__defaults__ = (os.path.abspath(sys.argv[1]),)
eng = gpu.GpuEngine(0)
N = 1 << 20
d_in = torch.from_numpy(synth.battle_states(N, seed=0xBA77)).cuda()
d_out = torch.empty_like(d_in); d_draws = torch.empty(N, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
for lanes in (16,):
    os.environ['CAAA_BATTLE_LANES'] = str(lanes); os.environ.setdefault('CAAA_BATTLE_POOL', '0')
    for label, outp in (("export", d_out.data_ptr()), ("no export", None)):
        for _ in range(2):
            eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, 0, outp, d_draws.data_ptr(), None, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for r in range(4):
            eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, r << 20, outp, d_draws.data_ptr(), None, st)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 4
        print(f"{sys.argv[1] if len(sys.argv) > 1 else 'shipped'} lanes={lanes} {label}: {N / ms * 1e3 / 1e6:.1f} M battles/s ({ms * 1e6 / N:.1f} ns/battle)", flush=True)
