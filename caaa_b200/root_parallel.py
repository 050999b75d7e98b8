"""Root-parallel MCTS across the GPUs of one box (BASELINE config 5).

Every rank grows its own tree from the same root with its own seed (no state is exchanged while
searching); the only exchange step is one all-reduce(sum) of the root children's visit counts and
values (<= 20 x (int64, double) = 320 bytes), over NCCL/NVLink on the GPU box and over gloo in the
CPU tests.  Root actions are in the same order on every rank because expansion is deterministic.
"""
import numpy as np
import torch
import torch.distributed as dist

MAX_ACTIONS = 20


def combine_root_stats(actions, visits, values, group=None, device=None):
    """All-reduce (sum) of per-rank root statistics; returns (actions, visits, values) of the whole job."""
    n = len(actions)
    dev = torch.device("cpu") if device is None else device
    buf_v = torch.zeros(MAX_ACTIONS, dtype=torch.int64, device=dev)
    buf_x = torch.zeros(MAX_ACTIONS, dtype=torch.float64, device=dev)
    buf_a = torch.zeros(MAX_ACTIONS, dtype=torch.int64, device=dev)
    buf_v[:n] = torch.as_tensor(np.asarray(visits, dtype=np.int64))
    buf_x[:n] = torch.as_tensor(np.asarray(values, dtype=np.float64))
    buf_a[:n] = torch.as_tensor(np.asarray(actions, dtype=np.int64))
    if dist.is_available() and dist.is_initialized():
        amax = buf_a.clone()
        dist.all_reduce(amax, op=dist.ReduceOp.MAX, group=group)
        if not torch.equal(amax, buf_a):
            raise RuntimeError("root action lists differ between ranks")
        dist.all_reduce(buf_v, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(buf_x, op=dist.ReduceOp.SUM, group=group)
    return np.asarray(actions), buf_v[:n].cpu().numpy(), buf_x[:n].cpu().numpy()


class NcclRootComm:
    """The library's own NCCL communicator (caaa_root_comm_init): one rank per GPU, the 128-byte id of rank 0 is
    handed round through the torch.distributed group that launched the job (any backend)."""

    def __init__(self, engine, group=None):
        import ctypes as C
        self.lib, self.C = engine.lib, C
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        ident = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            self._check(self.lib.caaa_root_unique_id(ident.ctypes.data))
        box = [ident.tobytes()]
        dist.broadcast_object_list(box, src=0, group=group)
        ident = np.frombuffer(box[0], dtype=np.uint8).copy()
        self._check(self.lib.caaa_root_comm_init(ident.ctypes.data, rank, world))

    def _check(self, rc):
        if rc != 0:
            raise RuntimeError(f"caaa_root error {rc}: {self.lib.caaa_root_last_error().decode()}")

    def allreduce(self, visits, values):
        """Sum over the ranks; returns (visits, values, microseconds of the call on this rank)."""
        v = np.ascontiguousarray(visits, dtype=np.int64).copy()
        x = np.ascontiguousarray(values, dtype=np.float64).copy()
        us = self.C.c_double(0.0)
        self._check(self.lib.caaa_root_allreduce_host(len(v), v.ctypes.data, x.ctypes.data, self.C.byref(us)))
        return v, x, us.value

    def close(self):
        self.lib.caaa_root_comm_destroy()


def select_best_action(actions, values):
    """select_best_action, reference src/mcts.cpp:247-258: the child with the largest accumulated value."""
    return int(np.asarray(actions)[int(np.argmax(values))])


def root_parallel_search(search_fn, state, iterations_per_rank, base_seed, group=None, device=None):
    """search_fn(state, iterations, seed) -> (actions, visits, values); seed = base_seed + rank."""
    rank = dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0
    actions, visits, values = search_fn(state, iterations_per_rank, base_seed + rank)
    actions, visits, values = combine_root_stats(actions, visits, values, group=group, device=device)
    return actions, visits, values, select_best_action(actions, values)


def shard_first_game_id(rank, world, step, playouts_per_rank_per_step):
    """Game-id range of one rank for one step of a sharded playout job (bench.py): ranks play disjoint,
    contiguous id ranges, so the union over ranks and steps is exactly what one GPU would play alone -- there is no
    data-path collective.  Returns the first id; the rank plays [first, first + playouts_per_rank_per_step)."""
    return (rank + world * step) * playouts_per_rank_per_step


def measured_root_parallel(engine, comm, state, iterations_per_rank, rollouts_per_leaf, leaves_per_batch, base_seed, group=None,
                           device=None):
    """BASELINE config 5 on the ranks of `group` (one per GPU): every rank runs the real pipelined search
    (caaa_mcts_search) with seed base_seed + rank, the root statistics are summed with the C entry point
    (caaa_root_allreduce_host over NCCL), and the result is checked against the host-side sum of the per-rank
    statistics gathered separately.  Returns a dict (same on every rank); timing = max over ranks."""
    import time
    from . import host
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    dist.barrier(group)
    t0 = time.perf_counter()
    a, v, x, _, info = host.mcts_search(engine, state, iterations_per_rank, rollouts_per_leaf, leaves_per_batch, base_seed + rank)
    search_s = time.perf_counter() - t0
    tv, tx, us = comm.allreduce(v, x)
    comm.allreduce(v, x)  # second call: the steady-state latency, without NCCL's first-use set-up
    _, _, us2 = comm.allreduce(v, x)
    per_rank = [None] * world
    dist.all_gather_object(per_rank, (a.tolist(), v.tolist(), x.tolist(), search_s, int(info["playouts"])), group=group)
    same_actions = all(p[0] == per_rank[0][0] for p in per_rank)
    sum_v = np.sum([np.asarray(p[1], dtype=np.int64) for p in per_rank], axis=0)
    sum_x = np.zeros(len(a))
    for p in per_rank:  # NCCL's reduction order over ranks is its own: compare within an ulp-scale tolerance
        sum_x = sum_x + np.asarray(p[2], dtype=np.float64)
    ok = bool(same_actions and np.array_equal(tv, sum_v) and np.allclose(tx, sum_x, rtol=1e-12, atol=1e-9))
    slowest = max(p[3] for p in per_rank)
    playouts = sum(p[4] for p in per_rank)
    return {"n_gpus": world, "iterations_per_gpu": int(iterations_per_rank), "rollouts_per_leaf": int(rollouts_per_leaf),
            "leaves_per_batch": int(leaves_per_batch), "playouts": int(playouts), "search_seconds_max_rank": slowest,
            "playouts_per_sec": playouts / slowest, "allreduce_us_first": us, "allreduce_us": us2,
            "allreduced_equals_host_sum": ok, "root_actions": [int(t) for t in a], "root_visits": [int(t) for t in tv],
            "best_action": select_best_action(a, tx)}
