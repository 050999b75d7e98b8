"""GPU tier (-m gpu): parity at BASELINE.json's FULL sizes, every unit checked, not a sample.

* config 2 -- every one of the 2**20 seeded playouts from the start position, all six side outputs (result,
  plies, decisions, draws, final player, hash of the final state), against the reference itself
  (`oracle/_ref/ref_cli`, the patched reference build, one process per host core);
* config 3 -- the whole 10 485 760-battle sweep runs on the GPU; a strided 65 536-battle subset spread over the
  whole id range is checked bit for bit (670 state bytes + draw count) against the C oracle;
* config 4 -- 64 mid-game leaves x 4096 playouts: every playout, and therefore every per-leaf mean and sigma,
  against the reference.

The reference binary travels to the GPU box with the snapshot (oracle/_ref is built here by oracle/build_ref.py);
/root/reference itself is never read.
"""
import os
import subprocess
import tempfile

import numpy as np
import pytest

from caaa_b200 import layout, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
REF_CLI = os.path.join(ROOT, "oracle", "_ref", "ref_cli")
SEED = 0xCAAA2024


@pytest.fixture(scope="module")
def eng():
    from caaa_b200 import gpu
    return gpu.GpuEngine(0)


def reference_stats(states, per_state, seed, first, n):
    """CaaaRolloutStats of playouts first .. first+n-1 from the reference's CPU loop, all host cores."""
    if not os.path.exists(REF_CLI):
        pytest.skip("oracle/_ref/ref_cli not shipped")
    cores = os.cpu_count() or 1
    per = (n + cores - 1) // cores
    with tempfile.TemporaryDirectory() as d:
        sp = os.path.join(d, "states.bin")
        np.ascontiguousarray(states, dtype=np.uint8).tofile(sp)
        procs, outs = [], []
        for j in range(cores):
            lo = j * per
            cnt = min(per, n - lo)
            if cnt <= 0:
                break
            out = os.path.join(d, f"o{j}.bin")
            outs.append(out)
            procs.append(subprocess.Popen([REF_CLI, sp, "1", str(seed), str(first + lo), str(cnt), out, str(per_state)],
                                          stdout=subprocess.DEVNULL, env={**os.environ, "CAAA_REF_ID0": str(first)}))
        for p in procs:
            assert p.wait() == 0
        return np.concatenate([np.fromfile(o, dtype=layout.STATS_DTYPE) for o in outs])


def test_config2_every_playout_vs_reference(eng):
    n = 1 << 20
    s0 = layout.start_state()
    res, stats, summ = eng.rollout_batch(s0, n, SEED, 0, want_stats=True)
    want = reference_stats(s0, 0, SEED, 0, n)
    assert len(want) == n
    for f in layout.STATS_DTYPE.names:
        bad = np.nonzero(stats[f] != want[f])[0]
        assert bad.size == 0, (f, bad[:8], stats[f][bad[:8]], want[f][bad[:8]])
    assert np.array_equal(res, want["result"])
    assert summ["playouts"] == n and summ["plies"] == want["turns"].astype(np.int64).sum() and summ["flagged"] == 0


def test_config3_ten_million_battles(eng, port):
    chunk, n_chunks, stride = 1 << 20, 10, 160
    total = checked = 0
    port.set_stream(1, 0xBA77)
    for c in range(n_chunks):
        bs = synth.battle_states(chunk, seed=0xBA77 + c)
        first = c * chunk
        out, draws, summ = eng.resolve_battles(bs, 0xBA77, first)
        assert summ["playouts"] == chunk and summ["flagged"] == 0 and summ["draws"] == draws.astype(np.int64).sum()
        total += chunk
        for i in range(c % stride, chunk, stride):
            port.load_state(bs[i], first + i)
            port.run_phase(9)    # resolve_sea_battles, reference src/engine.cpp:2312
            port.run_phase(11)   # resolve_land_battles, reference src/engine.cpp:3055
            assert np.array_equal(port.get_state(), out[i]), (c, i)
            assert port.draws() == draws[i], (c, i)
            checked += 1
    assert total == 10_485_760 and checked >= 65_536


def test_config4_leaf_parallel_every_playout_vs_reference(eng):
    a = np.load(os.path.join(GOLD, "advance.npz"))
    leaves = a["states"][:64]
    per_leaf, first = 4096, 7_000_000
    res, stats, summ = eng.rollout_batch(leaves, per_leaf, 0x1EAF, first, want_stats=True)
    want = reference_stats(leaves, per_leaf, 0x1EAF, first, len(leaves) * per_leaf)
    assert np.array_equal(stats, want)
    got_l, want_l = res.reshape(64, per_leaf), want["result"].reshape(64, per_leaf)
    assert np.array_equal(got_l.mean(axis=1), want_l.mean(axis=1)) and np.array_equal(got_l.std(axis=1), want_l.std(axis=1))
    assert summ["flagged"] == 0
