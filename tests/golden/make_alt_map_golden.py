#!/usr/bin/env python3
"""Golden vectors for the map-as-data path (run where /root/reference exists, after oracle/build_ref.py).

The oracle is the reference engine itself compiled with the data TUs of the second board
(oracle/_ref/libcaaa_ref_alt.so, generated from tests/golden/alt_map.json by oracle/build_ref.py).
Writes tests/golden/alt_map.npz: the board's topology tables, seeded playouts from its start position
(Philox stream), root expansion and a few n-turn advances.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from caaa_b200 import layout  # noqa: E402
from oracle import pyref  # noqa: E402

SEED = 0xA17A17
N = 3000


def alt_start_state():
    """Every land owned by its original owner, a factory on every land (as on the reference's board)."""
    s = np.zeros((), dtype=layout.STATE_DTYPE)
    s["money"] = [12, 18, 8, 15, 10]
    s["land_state"]["owner_idx"] = [0, 1, 2, 3, 4]
    s["land_state"]["factory_hp"] = [6, 4, 12, 3, 9]
    s["land_state"]["factory_max"] = [6, 4, 12, 3, 9]
    s["builds_left"] = [6, 4, 12, 3, 9, 0, 0, 0]
    return np.frombuffer(s.tobytes(), dtype=np.uint8).copy()


def main():
    ra = pyref.RefAlt()
    ra.set_stream(1, SEED)
    s0 = alt_start_state()
    stats = ra.rollout_many(s0, 0, N)
    tables = np.zeros(1, dtype=layout.TABLES_DTYPE)
    tables[0] = ra.get_tables()
    acts = ra.get_possible_actions(s0)
    children = []
    for a in acts:
        children.append(ra.apply_action(s0, int(a)))
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "alt_map.npz"), seed=np.uint64(SEED), start=s0, stats=stats,
                        tables=tables, root_actions=np.asarray(acts, dtype=np.uint8), root_children=np.asarray(children, dtype=np.uint8))
    print("alt_map.npz:", N, "playouts, mean result", stats["result"].mean(), "mean turns", stats["turns"].mean(), "root actions", list(acts))


if __name__ == "__main__":
    main()
