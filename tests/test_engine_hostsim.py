"""CPU-tier parity of the product's device engine (caaa_b200/csrc/engine.cuh).

The same source that the CUDA kernels run is compiled for the host (tests/hostsim) and compared
bit-for-bit with the plain-C oracle port, which is itself pinned against the patched reference
(tests/test_oracle_vs_ref.py).  The GPU tier (-m gpu) repeats the playout checks through the
C-ABI on the device.
"""
import numpy as np

from oracle import pyref
from test_oracle_vs_ref import SEED, snapshots


ROTATED = ("income_per_turn", "total_factory_count", "total_player_land_units", "total_player_sea_units")


def assert_cache_equal(ca, cb, where):
    """Everything the engine can observe: row 5 of the rotated caches is rotate_turns scratch, and list
    entries beyond their counts (factory_locations, enemies_0) are never read."""
    for name in pyref.CACHE_DTYPE.names:
        if name in ("enemies_0", "factory_locations", "pad_"):
            continue
        a, b = (ca[name][:5], cb[name][:5]) if name in ROTATED else (ca[name], cb[name])
        assert np.array_equal(a, b), (where, name, ca[name], cb[name])
    n_en = int(ca["enemies_count_0"])
    assert np.array_equal(ca["enemies_0"][:n_en], cb["enemies_0"][:n_en]), where
    for p in range(5):
        k = max(int(ca["total_factory_count"][p]), 0)
        assert np.array_equal(ca["factory_locations"][p][:k], cb["factory_locations"][p][:k]), (where, p)


def test_tables_match_oracle(port, hostsim):
    tp, th = port.get_tables(), hostsim.get_tables()
    for name in pyref.TABLES_DTYPE.names:
        assert np.array_equal(tp[name], th[name]), name


def test_playouts_bit_exact(port, hostsim):
    s = pyref.start_state()
    for seed, first, n in ((SEED, 0, 2000), (99, (1 << 33) + 5, 300)):
        port.set_stream(1, seed)
        hostsim.set_stream(1, seed)
        assert np.array_equal(port.rollout_many(s, first, n), hostsim.rollout_many(s, first, n))


def test_every_phase_and_cache_bit_exact(port, hostsim):
    s0 = pyref.start_state()
    port.set_stream(1, SEED)
    hostsim.set_stream(1, SEED)
    # row 5 of the rotated caches is scratch in both, enemies_0 beyond the count is stale in the port
    for g in range(40):
        port.load_state(s0, g)
        hostsim.load_state(s0, g)
        for turn in range(150):
            for ph in range(pyref.N_PHASES):
                assert port.run_phase(ph) == hostsim.run_phase(ph)
                assert np.array_equal(port.get_state(), hostsim.get_state()), (g, turn, ph)
                assert_cache_equal(port.get_cache(), hostsim.get_cache(), (g, turn, ph))
            assert port.draws() == hostsim.draws()
            sa = port.evaluate()
            assert sa == hostsim.evaluate()
            if not (0.01 < sa < 0.99):
                break


def test_playouts_from_midgame_snapshots(port, hostsim):
    port.set_stream(1, SEED)
    hostsim.set_stream(1, SEED)
    snaps = snapshots(port, 40)
    assert len(snaps) > 150
    for i, st in enumerate(snaps):
        assert np.array_equal(port.rollout_many(st, 1000 * i, 3), hostsim.rollout_many(st, 1000 * i, 3)), i


def test_expansion_mode_bit_exact(port, hostsim):
    rng = np.random.default_rng(11)
    for walk in range(40):
        sp = pyref.start_state()
        sh = sp.copy()
        for step in range(150):
            ap, ah = port.get_possible_actions(sp), hostsim.get_possible_actions(sh)
            assert port.stuck() == hostsim.stuck()
            assert np.array_equal(ap, ah), (walk, step)
            if len(ap) == 0:
                break
            act = ap[rng.integers(len(ap))]
            sp, sh = port.apply_action(sp, act), hostsim.apply_action(sh, act)
            stuck_p, stuck_h = port.stuck(), hostsim.stuck()
            assert stuck_p == stuck_h
            if stuck_p:
                break  # the reference never returns from this action; both implementations flag it
            assert np.array_equal(sp, sh), (walk, step)


def test_phases_resume_from_their_saved_cursor(port, hostsim):
    """The phase-regrouped kernel time-slices long phases: a phase yields, its loop cursor is saved in the game and
    it is resumed later.  Slicing (here: after every 1, 2 or 5 loop iterations) must not change a single byte."""
    s = pyref.start_state()
    port.set_stream(1, SEED)
    hostsim.set_stream(1, SEED)
    want = port.rollout_many(s, 0, 400)
    for budget in (1, 2, 5):
        assert np.array_equal(want, hostsim.rollout_many_sliced(s, 0, 400, budget)), budget
    snaps = snapshots(port, 12)
    for i, st in enumerate(snaps[::3]):
        assert np.array_equal(port.rollout_many(st, 5000 + 10 * i, 2), hostsim.rollout_many_sliced(st, 5000 + 10 * i, 2, 1)), i


def test_merged_land_phases_bit_exact(port, hostsim):
    """Tanks, artillery and infantry moved in one merged flat loop (what the phase-regrouped kernel does) play exactly
    the same game as the three separate phases, with and without yielding."""
    s = pyref.start_state()
    port.set_stream(1, SEED)
    hostsim.set_stream(1, SEED)
    want = port.rollout_many(s, 0, 400)
    for budget in (0, 1, 3):
        assert np.array_equal(want, hostsim.rollout_many_merged(s, 0, 400, budget)), budget
    for i, st in enumerate(snapshots(port, 12)[::3]):
        assert np.array_equal(port.rollout_many(st, 7000 + 10 * i, 2), hostsim.rollout_many_merged(st, 7000 + 10 * i, 2, 2)), i


def test_isolated_battles_bit_exact(port, hostsim):
    """The combat rounds and casualty loops (written without early exits so that the lanes of a warp reconverge) against
    the oracle on randomised battle states, from sparse stacks to per-type counts that wrap the reference's uint8 sums."""
    from caaa_b200 import synth
    for rep, max_count in enumerate((1, 3, 12, 60, 255)):
        bs = synth.battle_states(6000, seed=40 + rep, max_count=max_count)
        port.set_stream(1, 900 + rep)
        hostsim.set_stream(1, 900 + rep)
        for i in range(len(bs)):
            port.load_state(bs[i], 10**9 + i)
            hostsim.load_state(bs[i], 10**9 + i)
            for ph in (9, 11):  # resolve_sea_battles, resolve_land_battles
                assert port.run_phase(ph) == hostsim.run_phase(ph), (max_count, i, ph)
            assert np.array_equal(port.get_state(), hostsim.get_state()), (max_count, i)
            assert port.draws() == hostsim.draws(), (max_count, i)
