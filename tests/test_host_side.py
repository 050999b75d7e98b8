"""CPU tier: host-side code that needs no GPU -- game_data.json I/O in the reference's wire format and
the root-statistics exchange of root-parallel MCTS (2 ranks over gloo)."""
import json
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from caaa_b200 import host, layout, root_parallel, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_JSON = "/root/reference/game_data.json"


def test_start_position_json():
    here = os.path.join(ROOT, "tests", "golden", "start_position.json")
    st = host.load_game_data(here)
    assert np.array_equal(st, layout.start_state())


@pytest.mark.skipif(not os.path.exists(REF_JSON), reason="reference tree not mounted")
def test_reference_game_data_json():
    """The reference's own game_data.json (game_data.json:1-42) parses to the start position."""
    assert np.array_equal(host.load_game_data(REF_JSON), layout.start_state())


def test_json_round_trip_random_states():
    states = synth.battle_states(64, seed=5)
    rng = np.random.default_rng(3)
    for s in states:
        s = s.copy()
        s[606:] = 0  # skipped_moves is not part of the wire format
        s[14 + 3:14 + 125:25] = 0  # ... and neither is bombard_max (reference src/serialize_data.cpp:223-256)
        s[14 + 1] = rng.integers(0, 256)  # a negative factory_hp
        text = host.state_to_json(s)
        doc = json.loads(text)  # well-formed JSON with the reference's keys
        assert set(doc) == {"player_index", "money", "builds_left", "land_state", "units_sea", "flagged_for_combat",
                            "other_land_units", "other_sea_units"}
        assert len(doc["other_sea_units"][0][0]) == 15  # reference src/serialize_data.cpp:313
        back = host.state_from_json(text)
        assert np.array_equal(back, s)


def _rank_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    # CPU tier: there is no GPU here, so no search can run (the library has no CPU path).  This test covers the exchange
    # plumbing only -- seed = base + rank, action-list agreement, sum over ranks -- on synthetic per-rank statistics;
    # the REAL search on two GPUs with the NCCL C entry point is tests/test_gpu_mcts.py::test_root_parallel_two_ranks_real_search.
    def synthetic_rank_stats(state, iterations, seed):
        rng = np.random.default_rng(seed)
        visits = rng.integers(1, 1000, 6)
        return np.array([6, 5, 4, 3, 2, 0]), visits, visits * rng.random(6)

    a, v, x, best = root_parallel.root_parallel_search(synthetic_rank_stats, None, 100, 1000)
    if rank == 0:
        torch.save({"a": a, "v": v, "x": x, "best": best}, out)
    dist.destroy_process_group()


def test_root_statistics_allreduce_two_ranks(tmp_path):
    out = str(tmp_path / "r0.pt")
    mp.spawn(_rank_main, args=(2, 29517, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    want_v = np.zeros(6, dtype=np.int64)
    want_x = np.zeros(6)
    for r in range(2):
        rng = np.random.default_rng(1000 + r)
        v = rng.integers(1, 1000, 6)
        want_v += v
        want_x += v * rng.random(6)
    assert np.array_equal(got["v"], want_v)
    assert np.allclose(got["x"], want_x, rtol=0, atol=1e-9)
    assert got["best"] == [6, 5, 4, 3, 2, 0][int(np.argmax(want_x))]


def _shard_main(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per, steps = 1000, 3
    mine = torch.tensor([[root_parallel.shard_first_game_id(rank, world, k, per), per] for k in range(steps)], dtype=torch.int64)
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    # the job-wide playout count is the only reduction the sharded benchmark needs (bench.py sums plies the same way)
    total = torch.tensor([per * steps], dtype=torch.int64)
    dist.all_reduce(total)
    if rank == 0:
        torch.save({"ranges": torch.stack(gathered), "total": int(total.item())}, out)
    dist.destroy_process_group()


def test_sharded_game_id_ranges_two_ranks(tmp_path):
    """bench.py --gpus N shards game-id ranges over ranks with no data-path collective: the ranges of 2 ranks x 3 steps
    are disjoint and tile [0, 2 * 3 * 1000) exactly."""
    out = str(tmp_path / "s.pt")
    mp.spawn(_shard_main, args=(2, 29519, out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    r = got["ranges"].reshape(-1, 2).numpy()
    r = r[np.argsort(r[:, 0])]
    assert r[0, 0] == 0 and (r[1:, 0] == r[:-1, 0] + r[:-1, 1]).all()
    assert r[-1, 0] + r[-1, 1] == got["total"] == 6000
