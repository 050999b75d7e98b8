"""Per-stage counters of the pool battle kernel (CAAA_BATTLE_DEBUG) for a few configurations."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, synth
eng = gpu.GpuEngine(0)
N = 1 << 20
d_in = torch.from_numpy(synth.battle_states(N, seed=0xBA77)).cuda()
d_out = torch.empty_like(d_in); d_draws = torch.empty(N, dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream().cuda_stream
eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, 0, d_out.data_ptr(), d_draws.data_ptr(), None, st)
torch.cuda.synchronize()
os.environ['CAAA_BATTLE_DEBUG'] = '1'
for cfg, div in ((0, 4), (1, 4)):
    os.environ['CAAA_BATTLE_POOL'] = str(cfg); os.environ['CAAA_BATTLE_YIELD_DIV'] = str(div)
    print(f"--- pool config {cfg} yield_div {div}", file=sys.stderr, flush=True)
    eng.resolve_battles_device(d_in.data_ptr(), N, 0xBA77, 0, d_out.data_ptr(), d_draws.data_ptr(), None, st)
    torch.cuda.synchronize()
