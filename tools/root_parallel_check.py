#!/usr/bin/env python3
"""Root-parallel MCTS on the GPUs of one box (BASELINE config 5): torchrun --nproc-per-node N tools/root_parallel_check.py

Every rank searches its own tree on its own GPU; the root statistics are all-reduced with the library's C entry
point (caaa_root_allreduce over NCCL) and checked against the host-side sum of the per-rank statistics.
Rank 0 prints one JSON line; exit code 1 if the check fails.
"""
import argparse
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from caaa_b200 import gpu, layout, root_parallel  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=512)
    ap.add_argument("--rollouts", type=int, default=4096)
    ap.add_argument("--leaves-per-batch", type=int, default=256)
    args = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    eng = gpu.GpuEngine(local)
    comm = root_parallel.NcclRootComm(eng)
    s0 = layout.start_state()
    from caaa_b200 import host
    host.mcts_search(eng, s0, args.leaves_per_batch, 64, args.leaves_per_batch, 1)  # warm-up
    out = root_parallel.measured_root_parallel(eng, comm, s0, args.iterations, args.rollouts, args.leaves_per_batch, 0xCAAA)
    comm.close()
    if dist.get_rank() == 0:
        out["workload"] = "config 5: root-parallel MCTS from the start position, root statistics all-reduced by caaa_root_allreduce (NCCL)"
        print(json.dumps(out))
    dist.destroy_process_group()
    return 0 if out["allreduced_equals_host_sum"] else 1


if __name__ == "__main__":
    sys.exit(main())
