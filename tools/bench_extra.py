#!/usr/bin/env python3
"""Secondary measurements (BASELINE.json configs 3, 4, 5); bench.py is the headline (config 2).

    python tools/bench_extra.py config1 [--iterations 100000]          (CPU only: the reference's own search)
    python tools/bench_extra.py battles [--n 10485760]
    python tools/bench_extra.py leaf    [--leaves 1024 --rollouts 4096]
    [torchrun --nproc-per-node N] python tools/bench_extra.py mcts [--iterations 2048 --rollouts 4096 --leaves-per-batch 64]

Each prints one JSON line.  Timing: CUDA events inside the library (kernel_ms) and wall clock around
the C-ABI call (host buffers, copies included).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from caaa_b200 import gpu, host, layout, synth  # noqa: E402


def battles(args):
    """Config 3 at full size: ONE host-buffer call over args.n states (kernel rate from the library's CUDA events, e2e
    from the wall clock around the call: pinned staging, H2D, kernel, D2H inside), then the kernel alone on
    device-resident states."""
    import torch
    eng = gpu.GpuEngine(0)
    chunk = min(args.n, 1 << 20)
    one = synth.battle_states(chunk, seed=0xBA77)  # the sweep tiles one generated chunk; every copy gets its own game id
    reps = (args.n + chunk - 1) // chunk
    states = np.ascontiguousarray(np.tile(one, (reps, 1))[:args.n])
    eng.resolve_battles(states[:1 << 18], 0xBA77, 0)  # warm-up: allocates the pinned pipeline
    t0 = time.perf_counter()
    out, d, summ = eng.resolve_battles(states, 0xBA77, 0)
    wall = time.perf_counter() - t0
    k_ms, draws = float(summ["kernel_ms"]), int(summ["draws"])
    # the same call with page-locked caller buffers: the copy engines read and write them directly
    p_in = torch.empty(states.shape, dtype=torch.uint8, pin_memory=True)
    p_in.numpy()[:] = states
    p_out = torch.empty(states.shape, dtype=torch.uint8, pin_memory=True)
    p_draws = torch.empty(states.shape[0], dtype=torch.int32, pin_memory=True)
    eng.resolve_battles(p_in.numpy()[:1 << 18], 0xBA77, 0, out=p_out.numpy()[:1 << 18], draws=p_draws.numpy()[:1 << 18])
    t0 = time.perf_counter()
    _, _, summ_p = eng.resolve_battles(p_in.numpy(), 0xBA77, 0, out=p_out.numpy(), draws=p_draws.numpy())
    wall_p = time.perf_counter() - t0
    same_p = bool(np.array_equal(p_out.numpy(), out) and np.array_equal(p_draws.numpy(), d))
    del p_in, p_out, p_draws
    # kernel alone, inputs resident in HBM (one chunk at a time to bound device memory)
    dev = torch.device("cuda", 0)
    d_in = torch.from_numpy(one).to(dev)
    d_out = torch.empty_like(d_in)
    d_draws = torch.empty(chunk, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        eng.resolve_battles_device(d_in.data_ptr(), chunk, 0xBA77, 0, d_out.data_ptr(), d_draws.data_ptr(), None, stream)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for r in range(reps):
        eng.resolve_battles_device(d_in.data_ptr(), chunk, 0xBA77, r * chunk, d_out.data_ptr(), d_draws.data_ptr(), None, stream)
    e1.record()
    torch.cuda.synchronize()
    dev_ms = e0.elapsed_time(e1)
    n_dev = reps * chunk
    print(json.dumps({"workload": "config 3: isolated battle resolution sweep", "battles": int(args.n),
                      "battles_per_sec_kernel_resident": n_dev / (dev_ms * 1e-3),
                      "hbm_gbs_kernel_resident": n_dev * 1340 / (dev_ms * 1e-3) / 1e9,
                      "hbm_frac_of_6465": n_dev * 1340 / (dev_ms * 1e-3) / 1e9 / 6465.2,
                      "battles_per_sec_kernel_in_pipeline": args.n / (k_ms * 1e-3), "battles_per_sec_e2e": args.n / wall,
                      "battles_per_sec_e2e_page_locked_buffers": args.n / wall_p, "page_locked_equals_pageable": same_p,
                      "battles_per_sec_kernel_in_pipeline_page_locked": args.n / (float(summ_p["kernel_ms"]) * 1e-3),
                      "e2e_seconds": wall, "h2d_bytes": int(args.n) * 670, "d2h_bytes": int(args.n) * 674,
                      "mean_draws": draws / args.n}))


def leaf(args):
    eng = gpu.GpuEngine(0)
    s0 = layout.start_state()
    # randomised mid-game leaves made on the device: game g advanced 1 + (g * 7) % 60 turns of random play
    leaves = np.zeros((args.leaves, layout.STATE_BYTES), dtype=np.uint8)
    for t in range(60):
        idx = np.nonzero((1 + (np.arange(args.leaves) * 7) % 60) == t + 1)[0]
        for i in idx:
            leaves[i] = eng.advance(s0, 0x1EAF, int(i), t + 1)[0][0]
    eng.rollout_batch(leaves[:64], 64, 0x1EAF, 0)
    t0 = time.perf_counter()
    res, _, summ = eng.rollout_batch(leaves, args.rollouts, 0x1EAF, 0)
    wall = time.perf_counter() - t0
    n = args.leaves * args.rollouts
    print(json.dumps({"workload": "config 4: leaf-parallel rollouts from randomised mid-game states", "leaves": args.leaves,
                      "rollouts_per_leaf": args.rollouts, "playouts_per_sec_kernel": n / (float(summ["kernel_ms"]) * 1e-3),
                      "playouts_per_sec_e2e": n / wall, "plies_per_playout": float(summ["plies"]) / n,
                      "flagged": int(summ["flagged"]), "mean_result": float(res.mean())}))


def mcts(args):
    import torch
    import torch.distributed as dist
    from caaa_b200 import root_parallel
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = gpu.GpuEngine(local)
    s0 = layout.start_state()

    def search(state, iterations, seed):
        a, v, x, _, info = host.mcts_search(eng, state, iterations, args.rollouts, args.leaves_per_batch, seed)
        search.info = info
        return a, v, x

    search(s0, args.leaves_per_batch, 1)  # warm-up
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a, v, x, best = root_parallel.root_parallel_search(search, s0, args.iterations, 0xCAAA, device=torch.device("cuda", local))
    torch.cuda.synchronize()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=torch.device("cuda", local))
    if world > 1:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if rank == 0:
        total = int(v.sum())
        print(json.dumps({"workload": "config 5: root-parallel MCTS from the start position, root statistics all-reduced",
                          "n_gpus": world, "iterations_per_gpu": args.iterations, "rollouts_per_leaf": args.rollouts,
                          "leaves_per_batch": args.leaves_per_batch, "playouts": total, "seconds": float(dt.item()),
                          "playouts_per_sec": total / float(dt.item()), "root_actions": a.tolist(), "root_visits": v.tolist(),
                          "root_mean_value": (x / np.maximum(v, 1)).round(4).tolist(), "best_action": best}))
    if world > 1:
        dist.destroy_process_group()


def config1(args):
    """BASELINE config 1: the reference's own mcts_search (reference src/mcts.cpp:65-108) from the start position, fixed
    iteration count, ONE CPU thread, the reference's own glibc byte stream and literal behaviour (no per-playout
    valid_moves reset), i.e. the build SURVEY.md Appendix E took its known answers from."""
    import subprocess
    import tempfile
    cli = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        path = f.name
    layout.start_state().tofile(path)
    out = subprocess.run([cli, "mcts", path, str(args.iterations), "literal"], capture_output=True, text=True).stdout
    os.unlink(path)
    secs, children, best = None, [], None
    for line in out.splitlines():
        f = line.split()
        if f[:1] == ["MCTS"]:
            secs = float(f[2])
        elif f[:1] == ["CHILD"]:
            children.append({"action": int(f[1]), "visits": int(f[2]), "value": float(f[3])})
        elif f[:1] == ["BEST"]:
            best = int(f[1])
    print(json.dumps({"workload": "config 1: reference mcts_search from the start position, one CPU thread (oracle/_ref/ref_cli mcts)",
                      "iterations": args.iterations, "seconds": secs, "iterations_per_sec": args.iterations / secs,
                      "root_children": children, "best_action": best, "cores": 1,
                      "appendix_e": "20 000 iterations -> visits 892/4664/5166/3693/5054/531, best 4; 100 000 -> best 2"}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("what", choices=["battles", "leaf", "mcts", "config1"])
    ap.add_argument("--n", type=int, default=10 * (1 << 20))
    ap.add_argument("--leaves", type=int, default=1024)
    ap.add_argument("--rollouts", type=int, default=4096)
    ap.add_argument("--iterations", type=int, default=1024)
    ap.add_argument("--leaves-per-batch", type=int, default=64)
    args = ap.parse_args()
    {"battles": battles, "leaf": leaf, "mcts": mcts, "config1": config1}[args.what](args)


if __name__ == "__main__":
    main()
