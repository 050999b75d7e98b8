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
