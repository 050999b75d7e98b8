"""Pins the plain-C oracle port against the patched reference build (oracle/_ref).

The reference ships no tests or golden vectors, so the port is checked
against the reference's own functions on the same inputs: whole playouts, every phase of the
pipeline with all derived caches, and the stop-at-decision mode used for tree expansion.
"""
import os

import numpy as np
import pytest

from oracle import pyref

SEED = 0xCAAA2024


def snapshots(eng, n_games, every=7, max_turns=120):
    """Mid-game states at turn boundaries and in the middle of turns, taken from `eng` playouts."""
    out = []
    s0 = pyref.start_state()
    for g in range(n_games):
        eng.load_state(s0, g)
        for turn in range(max_turns):
            for ph in range(pyref.N_PHASES):
                eng.run_phase(ph)
                if (turn * 21 + ph + g) % (every * 21 + 5) == 0:
                    out.append(eng.get_state())
            sc = eng.evaluate()
            if not (0.01 < sc < 0.99):
                break
            if (turn + g) % every == 0:
                out.append(eng.get_state())
    return out


def test_tables_match_reference(ref, port):
    tr, tp = ref.get_tables(), port.get_tables()
    for name in pyref.TABLES_DTYPE.names:
        assert np.array_equal(tr[name], tp[name]), name


def test_playouts_from_start_bit_exact(ref, port):
    s = pyref.start_state()
    ref.set_stream(1, SEED)
    port.set_stream(1, SEED)
    a = ref.rollout_many(s, 0, 1500)
    b = port.rollout_many(s, 0, 1500)
    assert np.array_equal(a, b)
    # a different seed and a far-away game id range
    ref.set_stream(1, 12345)
    port.set_stream(1, 12345)
    a = ref.rollout_many(s, (1 << 40) + 17, 300)
    b = port.rollout_many(s, (1 << 40) + 17, 300)
    assert np.array_equal(a, b)


def test_every_phase_and_cache_bit_exact(ref, port):
    s0 = pyref.start_state()
    ref.set_stream(1, SEED)
    port.set_stream(1, SEED)
    for g in range(40):
        ref.load_state(s0, g)
        port.load_state(s0, g)
        for turn in range(150):
            for ph in range(pyref.N_PHASES):
                ra, rb = ref.run_phase(ph), port.run_phase(ph)
                assert ra == rb
                assert np.array_equal(ref.get_state(), port.get_state()), (g, turn, ph)
                ca, cb = ref.get_cache(), port.get_cache()
                for name in pyref.CACHE_DTYPE.names:
                    assert np.array_equal(ca[name], cb[name]), (g, turn, ph, name, ca[name], cb[name])
            assert ref.draws() == port.draws()
            sa, sb = ref.evaluate(), port.evaluate()
            assert sa == sb
            if not (0.01 < sa < 0.99):
                break


def test_playouts_from_midgame_snapshots(ref, port):
    ref.set_stream(1, SEED)
    port.set_stream(1, SEED)
    snaps = snapshots(ref, 25)
    assert len(snaps) > 100
    for i, st in enumerate(snaps):
        a = ref.rollout_many(st, 1000 * i, 3)
        b = port.rollout_many(st, 1000 * i, 3)
        assert np.array_equal(a, b), i


def test_expansion_mode_bit_exact(ref, port):
    """get_possible_actions / apply_action (reference src/engine.cpp:4050-4183) along random action walks.

    With the expansion mode's deterministic dice a sea battle can score 0 hits on both sides for
    ever: the reference then never returns.  The port has a round guard for that, so every action
    is tried on the port first and the reference is only asked about actions that terminate.
    """
    rng = np.random.default_rng(7)
    n_stuck = 0
    for walk in range(30):
        sr = pyref.start_state()
        sp = sr.copy()
        for step in range(120):
            ap = port.get_possible_actions(sp)
            assert not port.stuck()
            ar = ref.get_possible_actions(sr)
            assert np.array_equal(ar, ap), (walk, step)
            if len(ar) == 0:
                break
            nxt = None
            for act in rng.permutation(ar):
                cand = port.apply_action(sp, act)
                if port.stuck():
                    n_stuck += 1
                    continue
                nxt = (act, cand)
                break
            if nxt is None:
                break
            act, sp = nxt
            sr = ref.apply_action(sr, act)
            assert np.array_equal(sr, sp), (walk, step)
            if ref.is_terminal(sr):
                break
        ref.set_stream(1, SEED)
        port.set_stream(1, SEED)
        a = ref.rollout_many(sr, walk * 10, 4)
        b = port.rollout_many(sp, walk * 10, 4)
        assert np.array_equal(a, b), walk
    print("actions skipped because the reference would hang:", n_stuck)


def test_reference_mcts_known_answer_appendix_e(tmp_path):
    """BASELINE config 1 pinned: the reference's own mcts_search (src/mcts.cpp:65-108), 20 000 iterations from the start
    position on its own byte stream, reproduces SURVEY.md Appendix E's root statistics exactly."""
    import subprocess
    from caaa_b200 import layout
    cli = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "ref_cli")
    if not os.path.exists(cli):
        pytest.skip("oracle/_ref/ref_cli not built")
    p = str(tmp_path / "s0.bin")
    layout.start_state().tofile(p)
    out = subprocess.run([cli, "mcts", p, "20000", "literal"], capture_output=True, text=True).stdout
    kids = [l.split() for l in out.splitlines() if l.startswith("CHILD")]
    assert [(int(k[1]), int(k[2])) for k in kids] == [(6, 892), (5, 4664), (4, 5166), (3, 3693), (2, 5054), (0, 531)]
    assert [round(float(k[3]), 2) for k in kids] == [775.97, 4444.87, 4940.16, 3489.80, 4829.48, 438.85]
    assert "BEST 4" in out
