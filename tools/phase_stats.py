#!/usr/bin/env python3
"""Loop-iteration statistics of the phase loops (CPU; uses the host build of the device engine, tests/hostsim).

For every phase call that had work during N playouts from the start position: units waiting in the phase's cells at
entry (money for the purchase phase) and loop iterations of the flat phase loop.  From that: the lane efficiency a
lockstep batch of 16 games can reach (sum of iterations / (16 x max of iterations)) when batches are formed at random
and when they are formed from three work-size buckets.  Output: profiles/r1_phase_iteration_stats.md
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyref  # noqa: E402  (tools only)

NAMES = ["move fighters", "move bombers", "stage transports", "move tanks", "move artillery", "move infantry", "move transports",
         "move subs", "move ships", "sea battles", "unload transports", "land battles", "move AA guns", "land fighters", "land bombers",
         "purchase"]


def main(n_playouts=600):
    hs = pyref.HostSim()
    hs.set_stream(1, 0xCAAA2024)
    s = pyref.start_state()
    cap = 3 * 4_000_000
    out = np.zeros(cap, dtype=np.int32)
    buf = np.ascontiguousarray(s).view(np.uint8).reshape(-1)
    n = hs.lib.caaa_hostsim_phase_stats(buf.ctypes.data_as(C.c_void_p), C.c_int64(0), C.c_int(n_playouts),
                                        out.ctypes.data_as(C.c_void_p), C.c_int(cap))
    a = out[:3 * n].reshape(-1, 3)
    rng = np.random.default_rng(1)

    def eff(it, b=16, trials=4000):
        if len(it) < b:
            return float("nan")
        m = it[rng.integers(0, len(it), (trials, b))]
        return m.sum() / (m.max(axis=1).sum() * b)

    lines = ["# Iterations of the phase loops per game (host build of the device engine, %d playouts from the start position)" % n_playouts, "",
             "| phase | calls with work | mean iterations | p90 | max | lane efficiency, random batches of 16 | ... batches from 3 work-size buckets | corr(units, iterations) |",
             "|---|---|---|---|---|---|---|---|"]
    for ph in range(16):
        x = a[a[:, 0] == ph]
        if len(x) < 100:
            continue
        it, u = x[:, 2], x[:, 1]
        if u.std() > 0:
            b = np.digitize(u, np.quantile(u, [0.5, 0.8]))
            ks = [k for k in range(3) if (b == k).sum() >= 16]
            cost = sum(it[b == k].sum() / eff(it[b == k]) for k in ks)
            eb, corr = "%.2f" % (it[np.isin(b, ks)].sum() / cost), "%.2f" % np.corrcoef(u, it)[0, 1]
        else:
            eb, corr = "--", "--"
        lines.append("| %s | %d | %.2f | %d | %d | %.2f | %s | %s |" % (NAMES[ph], len(x), it.mean(), np.quantile(it, 0.9), it.max(), eff(it), eb, corr))
    text = "\n".join(lines) + "\n"
    open(os.path.join(ROOT, "profiles", "r1_phase_iteration_stats.md"), "w").write(text)
    print(text)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 600)
