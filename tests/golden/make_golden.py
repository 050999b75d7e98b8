#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ from the *patched reference itself*
(oracle/_ref, built from /root/reference by oracle/build_ref.py).  Run in the build container:

    python oracle/build_ref.py && python tests/golden/make_golden.py

The fixtures are what the GPU tier checks against on the B200 box, where /root/reference does not
exist.  Philox stream (patch set P2) with the seeds recorded in each file.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from caaa_b200 import synth  # noqa: E402
from oracle import pyref  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 0xCAAA2024


def fnv_rows(a):
    return np.array([pyref.fnv1a64(r) for r in a], dtype=np.uint64)


def main():
    ref = pyref.Ref()
    port = pyref.Port()  # only used to find expansion actions on which the reference would hang
    s0 = pyref.start_state()

    # 0. tables
    t = ref.get_tables()
    np.save(os.path.join(OUT, "tables.npy"), np.frombuffer(t.tobytes(), dtype=np.uint8))

    # 1. playouts from the start position (BASELINE config 2, first 4096 game ids + a far id range)
    ref.set_stream(1, SEED)
    a = ref.rollout_many(s0, 0, 4096)
    b = ref.rollout_many(s0, (1 << 40) + 17, 256)
    np.savez_compressed(os.path.join(OUT, "start_rollouts.npz"), seed=np.uint64(SEED), stats0=a,
                        first1=np.uint64((1 << 40) + 17), stats1=b)

    # 2. whole turns from the start position: state + caches + draw count after n turns
    n_adv = 96
    turns = 1 + (np.arange(n_adv) * 7) % 45
    states = np.zeros((n_adv, 670), dtype=np.uint8)
    caches = np.zeros(n_adv, dtype=pyref.CACHE_DTYPE)
    draws = np.zeros(n_adv, dtype=np.int32)
    for i in range(n_adv):
        ref.load_state(s0, 5000 + i)
        for _ in range(int(turns[i])):
            for ph in range(pyref.N_PHASES):
                ref.run_phase(ph)
        states[i] = ref.get_state()
        caches[i] = ref.get_cache()
        draws[i] = ref.draws()
    np.savez_compressed(os.path.join(OUT, "advance.npz"), seed=np.uint64(SEED), first=np.uint64(5000), turns=turns,
                        states=states, caches=caches, draws=draws)

    # 3. playouts from those mid-game states (leaf-parallel shape: several rollouts per leaf)
    per_leaf = 8
    mid = np.zeros((n_adv, per_leaf), dtype=pyref.STATS_DTYPE)
    for i in range(n_adv):
        mid[i] = ref.rollout_many(states[i], 100000 + i * per_leaf, per_leaf)
    np.savez_compressed(os.path.join(OUT, "midgame_rollouts.npz"), seed=np.uint64(SEED), first=np.uint64(100000),
                        per_leaf=np.int32(per_leaf), stats=mid)

    # 4. isolated battles (BASELINE config 3 generator, first 4096 battles)
    nb = 4096
    bs = synth.battle_states(nb, seed=0xBA77)
    ref.set_stream(1, 0xBA77)
    out_hash = np.zeros(nb, dtype=np.uint64)
    out_draws = np.zeros(nb, dtype=np.int32)
    for i in range(nb):
        ref.load_state(bs[i], i)
        ref.run_phase(9)
        ref.run_phase(11)
        out_hash[i] = pyref.fnv1a64(ref.get_state())
        out_draws[i] = ref.draws()
    np.savez_compressed(os.path.join(OUT, "battles.npz"), seed=np.uint64(0xBA77), n=np.int32(nb),
                        in_hash=fnv_rows(bs[:64]), out_hash=out_hash, draws=out_draws)
    ref.set_stream(1, SEED)

    # 5. tree expansion along random action walks (reference get_possible_actions / apply_action)
    rng = np.random.default_rng(2024)
    leaves, n_actions, actions, child_hash, stuck = [], [], [], [], []
    for walk in range(12):
        s = s0.copy()
        for step in range(60):
            acts = ref.get_possible_actions(s)
            if len(acts) == 0:
                break
            row_a = np.zeros(20, dtype=np.uint8)
            row_a[:len(acts)] = acts
            row_h = np.zeros(20, dtype=np.uint64)
            row_s = np.zeros(20, dtype=np.uint8)
            ok = []
            for k, act in enumerate(acts):
                cand = port.apply_action(s, act)
                if port.stuck():  # the reference never returns from this action
                    row_s[k] = 1
                    continue
                child = ref.apply_action(s, act)
                row_h[k] = pyref.fnv1a64(child)
                ok.append(child)
            leaves.append(s.copy())
            n_actions.append(len(acts))
            actions.append(row_a)
            child_hash.append(row_h)
            stuck.append(row_s)
            if not ok:
                break
            s = ok[rng.integers(len(ok))]
    np.savez_compressed(os.path.join(OUT, "expansion.npz"), leaves=np.array(leaves), n_actions=np.array(n_actions, dtype=np.uint8),
                        actions=np.array(actions), child_hash=np.array(child_hash), stuck=np.array(stuck))
    print("golden fixtures written:", sorted(f for f in os.listdir(OUT) if f.endswith((".npz", ".npy"))))


if __name__ == "__main__":
    main()
