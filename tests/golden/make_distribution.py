#!/usr/bin/env python3
"""Distribution fixture for the second correctness test of BASELINE.json's north_star: the
*unpatched-stream* reference (its own glibc rand() table and cursor, patch sets P0 + P1 i-iii
only, no per-playout valid_moves reset) playing N playouts from the start position.

The reference is single-threaded with global state, so the N playouts are split over worker
processes; worker j first discards j*per rand() values so that together the workers replay
exactly the first N playouts of one fresh reference process (each playout consumes one rand()).
"""
import ctypes
import multiprocessing as mp
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))
PER = 50000
WORKERS = 8


def work(j):
    from oracle import pyref
    ref = pyref.Ref()  # initialize_constants(): consumes the first 65 536 rand() values
    libc = ctypes.CDLL("libc.so.6")
    for _ in range(j * PER):
        libc.rand()
    ref.set_stream(0, 0)
    ref.set_reset_valid_moves(False)
    st = ref.rollout_many(pyref.start_state(), -1, PER)
    return st["result"].copy(), st["turns"].copy()


def main():
    with mp.get_context("spawn").Pool(WORKERS) as pool:
        parts = pool.map(work, range(WORKERS))
    res = np.concatenate([p[0] for p in parts])
    turns = np.concatenate([p[1] for p in parts])
    deciles = np.histogram(res, bins=np.linspace(0, 1, 11))[0]
    turn_edges = np.array([0, 25, 50, 75, 100, 150, 200, 300, 500, 999, 1001])
    turn_hist = np.histogram(turns, bins=turn_edges)[0]
    np.savez(os.path.join(OUT, "distribution.npz"), n=np.int64(len(res)), mean_result=res.mean(),
             win_rate=(res > 0.5).mean(), deciles=deciles, turn_edges=turn_edges, turn_hist=turn_hist,
             mean_turns=turns.mean(), first12=res[:12])
    print(len(res), res.mean(), (res > 0.5).mean(), deciles, turn_hist, turns.mean(), res[:12])


if __name__ == "__main__":
    main()
