#!/usr/bin/env python3
"""bench.py -- headline measurement: random playouts per second from the 1942 start position.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--playouts P]

A "step" is one pass of the hot path over one batch: P seeded playouts (default 2**20, BASELINE
config 2) of the start position per GPU.  N > 1 is launched by torchrun, one rank per GPU; the batch
is sharded by game-id range with no data-path collective (weak scaling).  Rank 0 prints ONE JSON line.

  value     playouts/s, states and results resident in HBM (device-pointer C-ABI entry point)
  e2e       same metric through the host-buffer C-ABI call (H2D of the states, D2H of the results inside)
  roofline  HBM roofline of the rollout kernel from algorithmic bytes (1340 B per ply)
            and the live CUDA-event duration.  The BINDING limit of this kernel is instruction issue /
            per-warp latency, so `roofline.issue` carries the issue-side counters (issue-slot utilisation,
            active threads per issued instruction, thread instructions per ply) and `roofline.traffic` the
            DRAM bytes per launch, both read from the committed ncu export profiles/r2_rollout_counters.json
            (made by tools/ncu_extract.py from an `ncu --set full` capture of this same command)
  cpu_baseline  the reference's own CPU rollout loop (oracle/_ref) on ONE host core, bounded sample, on
            the reference's own glibc byte stream (mode 0) and on the Philox stream the GPU uses (mode 1)

`--impl reference` times the reference's CPU loop on all host cores instead (one process per core: the
reference engine keeps its state in file-scope globals and cannot be threaded), on its own byte stream.
With N > 1 the line also carries `config.root_parallel`: a short root-parallel MCTS (BASELINE config 5) whose
root statistics are all-reduced by the C entry point caaa_root_allreduce and checked against the host-side sum.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED = 0xCAAA2024
BYTES_PER_PLY = 1340  # one 670-byte GameState read + written per player-turn
METRIC = "playouts/sec"
# kernels of one rollout batch: prepare_kernel (import of the start states), group_init_kernel (phase bitmaps),
# rollout_kernel_stream (the playouts)
KERNELS_PER_BATCH = 3
ROLLOUT_KERNEL = "rollout_kernel_stream<false, 16>"
# counters of one rollout_kernel_stream launch of 2**20 playouts, exported from the ncu capture of this command
NCU_COUNTERS = os.path.join(ROOT, "profiles", "r2_rollout_counters.json")


def ncu_counters(playouts):
    """Issue-side counters and DRAM traffic of the dominant kernel (None when absent or for another batch size)."""
    if playouts != (1 << 20) or not os.path.exists(NCU_COUNTERS):
        return None
    with open(NCU_COUNTERS) as f:
        return json.load(f)


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except (ValueError, IndexError):
                continue
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def write_start_state(path):
    from caaa_b200 import layout
    layout.start_state().tofile(path)


def run_ref_cli(n_procs, per_proc, mode, first=0):
    """per_proc playouts in each of n_procs reference processes; returns (playouts, plies, wall seconds)."""
    cli = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
    if not os.path.exists(cli):
        return None
    with tempfile.NamedTemporaryFile(suffix=".bin", delete=False) as f:
        path = f.name
    write_start_state(path)
    t0 = time.perf_counter()
    procs = [subprocess.Popen([cli, path, str(mode), str(SEED), str(first + j * per_proc), str(per_proc)],
                              stdout=subprocess.PIPE, text=True) for j in range(n_procs)]
    outs = [p.communicate()[0] for p in procs]
    wall = time.perf_counter() - t0
    os.unlink(path)
    playouts = plies = 0
    inner = 0.0
    for o in outs:
        f = o.split()
        playouts += int(f[0])
        plies += int(f[1])
        inner = max(inner, float(f[4]))
    return playouts, plies, inner, wall


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = os.cpu_count() or 1
    per_proc = args.ref_playouts_per_core
    mode = 0  # the reference's own byte stream (glibc table); the games differ from the GPU's, the rate is what is compared
    for _ in range(args.warmup):
        run_ref_cli(cores, max(200, per_proc // 10), mode)
    tot_playouts = tot_plies = 0
    tot_t = 0.0
    kind = "reference"
    for k in range(args.steps):
        r = run_ref_cli(cores, per_proc, mode, first=k * cores * per_proc)
        if r is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/ref_cli has not been built"}))
            return 0
        tot_playouts += r[0]
        tot_plies += r[1]
        tot_t += r[2]  # slowest process' own timer (excludes process start-up and table init)
    value = tot_playouts / tot_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "playouts/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * tot_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": "seeded random playouts from the 1942 start position (game_data.json), "
                               "reference CPU loop random_play_until_terminal, one process per host core",
                   "playouts_per_step": cores * per_proc, "plies_per_sec": tot_plies / tot_t},
        "cpu_baseline": {"value": value, "unit": "playouts/s", "cores": cores, "kind": kind,
                         "sample": f"{args.steps} x {cores} processes x {per_proc} playouts, the reference's own byte stream"},
        "e2e": {"value": value, "unit": "playouts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from caaa_b200 import gpu, layout
    from caaa_b200.root_parallel import shard_first_game_id

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; this implementation has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = gpu.GpuEngine(local)
    P = args.playouts
    dev = torch.device("cuda", local)
    s0 = layout.start_state()
    d_state = torch.from_numpy(s0).to(dev)
    d_results = torch.empty(P, dtype=torch.float64, device=dev)
    d_summary = torch.zeros(7, dtype=torch.int64, device=dev)  # CaaaBatchSummary (56 bytes)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step(k):
        eng.rollout_batch_device(d_state.data_ptr(), 1, P, SEED, shard_first_game_id(rank, world, k, P), d_results.data_ptr(), None,
                                 d_summary.data_ptr(), stream)

    # ---- value: inputs resident in HBM ----
    for k in range(args.warmup):
        device_step(k)
    barrier()
    d_summary.zero_()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    launches = 0
    for k in range(args.steps):
        flush.fill_(k & 0xff)  # L2 flush between timed steps (untimed)
        barrier()
        ev[k][0].record()
        device_step(args.warmup + k)
        launches += KERNELS_PER_BATCH
        ev[k][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    clocks = sampler.stop() if rank == 0 else None
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    summ = d_summary.cpu().numpy().view(np.uint64)
    plies = torch.tensor([float(summ[1])], dtype=torch.float64, device=dev)
    flagged = torch.tensor([float(summ[4])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(plies, op=dist.ReduceOp.SUM)
        dist.all_reduce(flagged, op=dist.ReduceOp.SUM)
    total_s = float(total_ms.item()) / 1000.0
    value = world * P * args.steps / total_s
    plies_total = float(plies.item())

    # ---- e2e: host buffers through the C-ABI, H2D + D2H inside the timed region ----
    states_host = s0.reshape(1, -1)
    for k in range(max(1, args.warmup // 2)):
        eng.rollout_batch(states_host, P, SEED, rank * P)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = args.steps
    for k in range(e2e_steps):
        res, _, _ = eng.rollout_batch(states_host, P, SEED, shard_first_game_id(rank, world, 1000 + k, P))
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * P * e2e_steps / float(e2e_s.item())

    # ---- N > 1: BASELINE config 5 beside the headline: root-parallel MCTS, one tree per GPU, root statistics all-reduced by
    # the library's NCCL entry point and checked against the host-side sum of the per-rank statistics ----
    rp = None
    if world > 1 and not args.no_root_parallel:
        from caaa_b200 import host, root_parallel
        comm = root_parallel.NcclRootComm(eng)
        host.mcts_search(eng, s0, 256, 64, 256, 1)  # warm-up
        rp = root_parallel.measured_root_parallel(eng, comm, s0, args.mcts_iterations, 4096, 256, 0xCAAA)
        comm.close()
        rp["note"] = ("one pipelined search per GPU (caaa_mcts_search: 256 leaves x 4096 rollouts per device batch, two batches in "
                      "flight), seed = base + rank; exchange = caaa_root_allreduce_host (ncclAllReduce of <= 20 x {int64, double})")
    if rank == 0:
        peak, peak_src = measured_peaks()
        per_launch_ms = statistics.mean(step_ms)
        plies_per_launch = plies_total / (world * args.steps)
        achieved = plies_per_launch * BYTES_PER_PLY / (per_launch_ms * 1e-3) / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r0 = run_ref_cli(1, args.cpu_sample // 2, 0)  # the reference's own glibc table stream
            r1 = run_ref_cli(1, args.cpu_sample // 2, 1)  # the Philox stream the GPU plays (same games as the GPU's)
            if r0 is not None and r1 is not None:
                cpu = {"value": r0[0] / r0[2], "unit": "playouts/s", "cores": 1, "kind": "reference",
                       "sample": f"{args.cpu_sample // 2} playouts from the start position, one process, the reference's own "
                                 f"byte stream ({r0[2]:.1f} s); oracle/_ref = the patched reference build",
                       "plies_per_sec": r0[1] / r0[2], "plies_per_playout": r0[1] / r0[0],
                       "philox_stream": {"value": r1[0] / r1[2], "plies_per_sec": r1[1] / r1[2], "plies_per_playout": r1[1] / r1[0],
                                         "sample": f"{args.cpu_sample // 2} playouts, Philox4x32-10 stream ({r1[2]:.1f} s)"}}
        nc = ncu_counters(P)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": nc["dram_bytes_per_launch"] if nc else None, "peak_source": peak_src, "kernel": ROLLOUT_KERNEL,
                    "note": "HBM view as the contract asks (algorithmic 1340 B/ply); this integer, branchy path is bound by "
                            "instruction issue and per-warp latency, see `issue`"}
        if nc:
            # issue roofline: 148 SMs x 4 schedulers x 1 warp instruction per cycle; lanes: 32 threads per instruction
            roofline["issue"] = {"warp_issue_frac": nc["warp_issue_frac"], "active_threads_per_inst": nc["active_threads_per_inst"],
                                 "thread_issue_frac": nc["thread_issue_frac"], "thread_inst_per_ply": nc["thread_inst_per_ply"],
                                 "warp_inst_per_ply": nc["warp_inst_per_ply"], "branch_uniformity_pct": nc["branch_uniformity_pct"],
                                 "warps_active_pct": nc["warps_active_pct"], "source": "profiles/r2_rollout_counters.json "
                                 "(ncu --set full of this command; tools/ncu_extract.py)"}
        line = {
            "metric": METRIC, "value": value, "unit": "playouts/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1000.0 * total_s / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "BASELINE config 2: seeded random playouts from the 1942 start position "
                                   "(game_data.json), bit-exact vs the patched reference",
                       "playouts_per_gpu_per_step": P, "seed": SEED, "plies_per_sec": plies_total / total_s,
                       "plies_per_playout": plies_total / (world * P * args.steps), "flagged_playouts": float(flagged.item()),
                       "l2": "256 MiB write between timed steps (untimed); the 62 MB game pool is re-initialised by every step",
                       "sharding": "game-id ranges per rank, no data-path collective"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": "playouts/s", "h2d_bytes_per_step": 670, "d2h_bytes_per_step": 8 * P + 56},
            "gpu_launches": launches, "clocks": clocks,
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        if rp is not None:
            line["config"]["root_parallel"] = rp
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--playouts", type=int, default=1 << 20, help="playouts per GPU per step")
    ap.add_argument("--cpu-sample", type=int, default=40000, help="playouts of the single-core CPU baseline")
    ap.add_argument("--ref-playouts-per-core", type=int, default=4000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-root-parallel", action="store_true")
    ap.add_argument("--mcts-iterations", type=int, default=2048, help="leaves per GPU of the root-parallel block (N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        return reference_arm(args)
    return ours(args)


if __name__ == "__main__":
    sys.exit(main())
