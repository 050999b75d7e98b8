"""GPU tier (-m gpu): the CUDA path, called through the C-ABI (libcaaa_gpu.so), against

* the golden fixtures generated from the patched reference (tests/golden),
* the oracle port on fresh seeds (sizes the oracle finishes in seconds),
* size-independent properties at BASELINE.json's full size (1 M playouts): determinism,
  partition invariance, counter consistency, and the distribution test against the
  unpatched-stream reference with a stated binomial tolerance.

Everything is bit-exact (integer / byte work and an IEEE double ratio) except the distribution test.
"""
import os

import numpy as np
import pytest

from caaa_b200 import layout, synth
from test_engine_hostsim import assert_cache_equal

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def fnv_rows(a):
    h = np.full(a.shape[0], 0xcbf29ce484222325, dtype=np.uint64)
    prime = np.uint64(0x100000001b3)
    with np.errstate(over="ignore"):
        for j in range(a.shape[1]):
            h = (h ^ a[:, j].astype(np.uint64)) * prime
    return h


@pytest.fixture(scope="module")
def eng():
    from caaa_b200 import gpu
    return gpu.GpuEngine(0)  # raises if the library or the device is missing: no fallback


def test_device_is_blackwell(eng):
    info = eng.device_info()
    assert info["smem_per_block"] >= 227 * 1024
    assert info["sm_count"] >= 100


def test_start_rollouts_golden(eng):
    g = np.load(os.path.join(GOLD, "start_rollouts.npz"))
    s0 = layout.start_state()
    res, stats, summ = eng.rollout_batch(s0, 4096, int(g["seed"]), 0, want_stats=True)
    assert np.array_equal(stats, g["stats0"])
    assert np.array_equal(res, g["stats0"]["result"])
    assert summ["playouts"] == 4096 and summ["plies"] == g["stats0"]["turns"].sum()
    assert summ["draws"] == g["stats0"]["draws"].sum() and summ["decisions"] == g["stats0"]["decisions"].sum()
    assert summ["flagged"] == 0
    _, stats1, _ = eng.rollout_batch(s0, 256, int(g["seed"]), int(g["first1"]), want_stats=True)
    assert np.array_equal(stats1, g["stats1"])


def test_start_rollouts_vs_oracle_fresh_seed(eng, port):
    s0 = layout.start_state()
    for seed, first, n in ((0x5EED5EED1234, 77, 2500), (3, (1 << 62) + 1, 500)):
        port.set_stream(1, seed)
        want = port.rollout_many(s0, first, n)
        _, got, _ = eng.rollout_batch(s0, n, seed, first, want_stats=True)
        assert np.array_equal(got, want)


def test_advance_golden(eng):
    a = np.load(os.path.join(GOLD, "advance.npz"))
    s0 = layout.start_state()
    for i in range(len(a["turns"])):
        st, caches, draws = eng.advance(s0, int(a["seed"]), int(a["first"]) + i, int(a["turns"][i]))
        assert np.array_equal(st[0], a["states"][i]), i
        assert draws[0] == a["draws"][i]
        assert_cache_equal(a["caches"][i], caches[0], i)


def test_leaf_parallel_midgame_golden(eng):
    a = np.load(os.path.join(GOLD, "advance.npz"))
    m = np.load(os.path.join(GOLD, "midgame_rollouts.npz"))
    per_leaf = int(m["per_leaf"])
    _, stats, summ = eng.rollout_batch(a["states"], per_leaf, int(m["seed"]), int(m["first"]), want_stats=True)
    assert np.array_equal(stats.reshape(-1, per_leaf), m["stats"])
    assert summ["flagged"] == 0


def test_battles_golden(eng):
    g = np.load(os.path.join(GOLD, "battles.npz"))
    bs = synth.battle_states(int(g["n"]), seed=int(g["seed"]))
    out, draws, summ = eng.resolve_battles(bs, int(g["seed"]), 0)
    assert np.array_equal(fnv_rows(out), g["out_hash"])
    assert np.array_equal(draws, g["draws"])
    assert summ["draws"] == g["draws"].sum() and summ["flagged"] == 0


def test_battles_vs_oracle_ragged(eng, port):
    """A ragged batch (not a multiple of the block size) with a different seed and id offset."""
    bs = synth.battle_states(1000 + 37, seed=99)
    out, draws, _ = eng.resolve_battles(bs, 4242, 10**12)
    port.set_stream(1, 4242)
    for i in range(0, len(bs), 5):
        port.load_state(bs[i], 10**12 + i)
        port.run_phase(9)
        port.run_phase(11)
        assert np.array_equal(port.get_state(), out[i]), i
        assert port.draws() == draws[i]


def test_battle_sweep_properties(eng):
    """BASELINE config 3 shape at a size the oracle cannot reach in seconds (1 M battles): the result of a battle is a
    pure function of (state, seed, game id) -- identical however the sweep is batched -- battles only ever remove
    units, and a resolved territory is no longer flagged unless the fight was broken off by a retreat decision."""
    n = 1 << 20
    chunk = synth.battle_states(1 << 16, seed=0xBA77)
    bs = np.tile(chunk, (n // len(chunk), 1))
    out, draws, summ = eng.resolve_battles(bs, 0xBA77, 0)
    assert summ["playouts"] == n and summ["draws"] == draws.astype(np.int64).sum() and summ["flagged"] == 0
    cut = 300_001
    oa, da, _ = eng.resolve_battles(bs[:cut], 0xBA77, 0)
    ob, db, _ = eng.resolve_battles(bs[cut:], 0xBA77, cut)
    assert np.array_equal(np.concatenate([oa, ob]), out) and np.array_equal(np.concatenate([da, db]), draws)
    # same states, other game ids -> other dice: the tiled copies of one state do not all end alike
    k = len(chunk)
    assert (fnv_rows(out[:k]) != fnv_rows(out[k:2 * k])).mean() > 0.3
    # unit counts never grow (bytes 14..597 hold every unit count; owner / factory bytes may change on conquest)
    before = bs.view(layout.STATE_DTYPE).reshape(-1)
    after = out.view(layout.STATE_DTYPE).reshape(-1)
    assert (after["other_land_units"] <= before["other_land_units"]).all()
    sea_b, sea_a = before["other_sea_units"].astype(np.int32), after["other_sea_units"].astype(np.int32)
    assert (sea_a[..., :12] <= sea_b[..., :12]).all()  # every defending type but the two battleship states
    assert (sea_a[..., 12:].sum(axis=-1) <= sea_b[..., 12:].sum(axis=-1)).all()  # a hit battleship turns into a damaged one
    # attackers may retreat to a neighbouring land, so only their total over the map is monotone
    assert (after["land_state"]["infantry"].astype(np.int32).sum(axis=(-1, -2)) <=
            before["land_state"]["infantry"].astype(np.int32).sum(axis=(-1, -2))).all()


def test_battle_kernels_and_buffers_agree(eng):
    """The shipped pool kernel in its three warp splits and the refill kernel (16 / 32 battle-owning lanes) resolve a
    ragged multi-chunk batch identically, and page-locked caller buffers (moved by the copy engines directly, no staging
    copy) give what pageable ones give."""
    import torch
    n = 2 * 131072 + 4321  # three pipeline chunks, the last one ragged
    bs = np.tile(synth.battle_states(1 << 14, seed=77), (n // (1 << 14) + 1, 1))[:n]
    want, want_draws, summ = eng.resolve_battles(bs, 555, 9_000_000_000)
    assert summ["playouts"] == n
    saved = {k: os.environ.get(k) for k in ("CAAA_BATTLE_KERNEL", "CAAA_BATTLE_POOL", "CAAA_BATTLE_LANES")}
    try:
        for env in ({"CAAA_BATTLE_POOL": "1"}, {"CAAA_BATTLE_POOL": "2"}, {"CAAA_BATTLE_KERNEL": "refill", "CAAA_BATTLE_LANES": "16"},
                    {"CAAA_BATTLE_KERNEL": "refill", "CAAA_BATTLE_LANES": "32"}):
            for k in saved:
                os.environ.pop(k, None)
            os.environ.update(env)
            out, draws, _ = eng.resolve_battles(bs[:70_001], 555, 9_000_000_000)
            assert np.array_equal(out, want[:70_001]) and np.array_equal(draws, want_draws[:70_001]), env
    finally:
        for k, v in saved.items():
            os.environ.pop(k, None)
            if v is not None:
                os.environ[k] = v
    p_in = torch.empty(bs.shape, dtype=torch.uint8, pin_memory=True)
    p_in.numpy()[:] = bs
    p_out = torch.zeros(bs.shape, dtype=torch.uint8, pin_memory=True)
    p_draws = torch.zeros(n, dtype=torch.int32, pin_memory=True)
    eng.resolve_battles(p_in.numpy(), 555, 9_000_000_000, out=p_out.numpy(), draws=p_draws.numpy())
    assert np.array_equal(p_out.numpy(), want) and np.array_equal(p_draws.numpy(), want_draws)


def test_expansion_golden(eng):
    g = np.load(os.path.join(GOLD, "expansion.npz"))
    n_actions, actions, children, status = eng.expand(g["leaves"])
    assert np.array_equal(n_actions, g["n_actions"])
    for i in range(len(n_actions)):
        n = int(n_actions[i])
        assert np.array_equal(actions[i][:n], g["actions"][i][:n]), i
        ok = g["stuck"][i][:n] == 0
        assert np.array_equal(fnv_rows(children[i][:n])[ok], g["child_hash"][i][:n][ok]), i
        assert (status[i] != 0) == bool(g["stuck"][i][:n].any()), i


def test_regrouped_kernel_small_batches_vs_oracle(port, tmp_path):
    """Small batches default to the lockstep kernel; force the phase-regrouped kernel through the same ragged shapes."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from caaa_b200 import gpu, layout; e = gpu.GpuEngine(0); "
            "out = [e.rollout_batch(layout.start_state(), n, 0xBA7C4, 1000 * n, want_stats=True)[1] for n in (1, 2, 17, 33, 225, 449)]; "
            "np.save(sys.argv[1], np.concatenate(out))" % root)
    out = str(tmp_path / "s.npy")
    subprocess.check_call([sys.executable, "-c", code, out], env={**os.environ, "CAAA_ROLLOUT_KERNEL": "stream"})
    port.set_stream(1, 0xBA7C4)
    s0 = layout.start_state()
    want = np.concatenate([port.rollout_many(s0, 1000 * n, n) for n in (1, 2, 17, 33, 225, 449)])
    assert np.array_equal(np.load(out), want)


def test_ragged_batches_vs_oracle(eng, port):
    """Batch shapes around the kernel's internal sizes (32 lanes, 240 slots per SM, 16 games per warp): one game,
    partial warps, partial CTAs, several start states with one or a few playouts each."""
    s0 = layout.start_state()
    seed = 0xBA7C4
    port.set_stream(1, seed)
    first = 123456
    for n in (1, 2, 15, 17, 31, 33, 223, 225, 239, 241, 449, 481):
        _, got, summ = eng.rollout_batch(s0, n, seed, first, want_stats=True)
        assert np.array_equal(got, port.rollout_many(s0, first, n)), n
        assert summ["playouts"] == n
        first += n
    a = np.load(os.path.join(GOLD, "advance.npz"))
    states = a["states"][:37]
    for per in (1, 3):
        _, got, _ = eng.rollout_batch(states, per, seed, 999, want_stats=True)
        want = np.concatenate([port.rollout_many(states[i], 999 + i * per, per) for i in range(len(states))])
        assert np.array_equal(got, want), per


def test_lockstep_and_regrouped_kernels_agree(tmp_path):
    """The walk-the-pipeline kernel (CAAA_ROLLOUT_KERNEL=lockstep, with and without the per-phase CTA barrier), the
    phase-regrouped kernel with other group sizes and with yielding phases, and the SM-resident experiment
    (CAAA_ROLLOUT_KERNEL=resident) play the same playouts bit for bit."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import sys, numpy as np; sys.path.insert(0, %r); from caaa_b200 import gpu, layout; "
            "e = gpu.GpuEngine(0); r, s, _ = e.rollout_batch(layout.start_state(), 3000, 0xABCD, 5, want_stats=True); "
            "np.save(sys.argv[1], s)" % root)
    outs = []
    for k, env in enumerate(({"CAAA_ROLLOUT_KERNEL": "stream"}, {"CAAA_ROLLOUT_KERNEL": "lockstep"},
                             {"CAAA_ROLLOUT_KERNEL": "lockstep", "CAAA_SYNC": "1"},
                             {"CAAA_ROLLOUT_KERNEL": "stream", "CAAA_GROUP": "1"},
                             {"CAAA_ROLLOUT_KERNEL": "stream", "CAAA_YIELD": "50", "CAAA_YIELD_BATTLES": "25"},
                             {"CAAA_ROLLOUT_KERNEL": "resident", "CAAA_WARPS": "12", "CAAA_YIELD": "50"})):
        out = str(tmp_path / f"k{k}.npy")
        subprocess.check_call([sys.executable, "-c", code, out], env={**os.environ, **env})
        outs.append(np.load(out))
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


def test_single_playout_shim(eng):
    g = np.load(os.path.join(GOLD, "start_rollouts.npz"))
    eng.set_stream(int(g["seed"]), 0)
    s0 = layout.start_state()
    for i in range(5):
        assert eng.random_play_until_terminal(s0) == g["stats0"]["result"][i]


def test_evaluate_matches_first_score(eng, port):
    a = np.load(os.path.join(GOLD, "advance.npz"))
    sc = eng.evaluate(a["states"][:32])
    for i in range(32):
        port.load_state(a["states"][i], 0)
        assert port.evaluate() == sc[i]


def test_full_size_properties_and_distribution(eng):
    """BASELINE config 2 at full size: 1 M seeded playouts from the start position."""
    n = 1 << 20
    seed = 0xCAAA2024
    s0 = layout.start_state()
    res, stats, summ = eng.rollout_batch(s0, n, seed, 0, want_stats=True)
    g = np.load(os.path.join(GOLD, "start_rollouts.npz"))
    assert np.array_equal(stats[:4096], g["stats0"])
    # determinism and partition invariance: same (seed, game id) -> same playout, however it is batched
    res2, _, _ = eng.rollout_batch(s0, n, seed, 0)
    assert np.array_equal(res, res2)
    cut = 333_333
    ra, _, _ = eng.rollout_batch(s0, cut, seed, 0)
    rb, _, _ = eng.rollout_batch(s0, n - cut, seed, cut)
    assert np.array_equal(np.concatenate([ra, rb]), res)
    # counters
    assert summ["playouts"] == n and summ["flagged"] == 0
    assert summ["plies"] == stats["turns"].astype(np.int64).sum()
    assert summ["draws"] == stats["draws"].astype(np.int64).sum()
    assert summ["decisions"] == stats["decisions"].astype(np.int64).sum()
    assert stats["turns"].min() >= 1 and stats["turns"].max() <= 1000
    assert (res >= 0).all() and (res <= 1).all()
    assert (stats["draws"] >= stats["decisions"]).all()
    # distribution vs the reference on its own glibc stream (tests/golden/make_distribution.py)
    d = np.load(os.path.join(GOLD, "distribution.npz"))
    n_ref = float(d["n"])
    p_ref, p_gpu = float(d["win_rate"]), float((res > 0.5).mean())
    p = (p_ref * n_ref + p_gpu * n) / (n_ref + n)
    sigma = np.sqrt(p * (1 - p) * (1 / n_ref + 1 / n))
    assert abs(p_gpu - p_ref) <= 4 * sigma, (p_gpu, p_ref, sigma)            # binomial, 4 sigma
    sd = float(res.std())
    assert abs(res.mean() - float(d["mean_result"])) <= 4 * sd * np.sqrt(1 / n_ref + 1 / n)
    # chi-square homogeneity on the result deciles and the game-length histogram (p > 1e-4)
    from scipy.stats import chi2_contingency
    dec = np.histogram(res, bins=np.linspace(0, 1, 11))[0]
    assert chi2_contingency(np.array([dec, d["deciles"]]))[1] > 1e-4
    th = np.histogram(stats["turns"], bins=d["turn_edges"])[0]
    assert chi2_contingency(np.array([th, d["turn_hist"]]))[1] > 1e-4


def test_reference_named_api_drop_in(eng, tmp_path):
    """The reference's API names (caaa_engine_compat.h) against SURVEY.md Appendix E's root expansion and
    the golden rollouts: what mcts.cpp / main.cpp would see after switching libraries."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "compat_driver")
    subprocess.check_call(["g++", "-std=c++17", "-O1", os.path.join(root, "tests", "compat_driver.cpp"), "-o", exe,
                           "-L" + os.path.join(root, "caaa_b200"), "-lcaaa_gpu", "-Wl,-rpath," + os.path.join(root, "caaa_b200")])
    out = subprocess.check_output([exe, os.path.join(GOLD, "start_position.json")], text=True).splitlines()
    assert out[0] == "actions 6 5 4 3 2 0"
    assert out[1:7] == ["child 6 player 1 money0 20 terminal 0", "child 5 player 0 money0 5 terminal 0",
                        "child 4 player 0 money0 5 terminal 0", "child 3 player 0 money0 6 terminal 0",
                        "child 2 player 0 money0 7 terminal 0", "child 0 player 1 money0 20 terminal 0"]
    g = np.load(os.path.join(GOLD, "start_rollouts.npz"))
    got = [float(l.split()[1]) for l in out[7:11]]
    assert got == g["stats0"]["result"][:4].tolist()


def test_mcts_search_on_gpu(eng):
    from caaa_b200 import host
    s0 = layout.start_state()
    # reference-shaped loop: one leaf, one rollout per iteration
    a, v, x, best, info = host.mcts_search(eng, s0, 3000, rollouts_per_leaf=1, leaves_per_batch=1, seed=11)
    assert a.tolist() == [6, 5, 4, 3, 2, 0]            # root expansion, SURVEY.md Appendix E
    assert v.sum() == 3000 and info["playouts"] == 3000
    assert (x >= 0).all() and (x <= v).all()
    # same seed -> same tree
    a2, v2, x2, best2, _ = host.mcts_search(eng, s0, 3000, 1, 1, seed=11)
    assert np.array_equal(v, v2) and np.array_equal(x, x2) and best == best2
    # leaf-parallel + batched leaves (BASELINE config 4 shape)
    a3, v3, x3, best3, info3 = host.mcts_search(eng, s0, 512, rollouts_per_leaf=256, leaves_per_batch=16, seed=5)
    assert v3.sum() == 512 * 256 and info3["playouts"] == 512 * 256
    assert best3 in a3.tolist()
    # pure purchases of the first mover cannot change the material ratio much: all children's means are close
    means = x3 / np.maximum(v3, 1)
    assert means.max() - means.min() < 0.2


def test_soa_pack_unpack_round_trip(eng):
    """caaa_gpu_pack / caaa_gpu_unpack: chunk-major SoA form and back, ragged sizes; chunk c of state i sits where the
    header says, pad bytes are zero."""
    import torch
    a = np.load(os.path.join(GOLD, "advance.npz"))["states"]
    for n in (1, 31, 32, 33, 96):
        st = np.ascontiguousarray(a[:n])
        d_aos = torch.from_numpy(st).cuda()
        nb = eng.soa_bytes(n)
        n_pad = (n + 31) // 32 * 32
        assert nb == 42 * n_pad * 16
        d_soa = torch.full((nb,), 0xAB, dtype=torch.uint8, device="cuda")
        d_back = torch.zeros_like(d_aos)
        eng.pack(d_aos.data_ptr(), n, d_soa.data_ptr())
        eng.unpack(d_soa.data_ptr(), n, d_back.data_ptr())
        torch.cuda.synchronize()
        assert np.array_equal(d_back.cpu().numpy(), st)
        soa = d_soa.cpu().numpy().reshape(42, n_pad, 16)
        padded = np.zeros((n, 672), dtype=np.uint8)
        padded[:, :670] = st
        assert np.array_equal(soa[:, :n, :].transpose(1, 0, 2).reshape(n, 672), padded)
