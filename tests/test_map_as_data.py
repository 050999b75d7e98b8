"""The board as data (first step): `caaa_gpu_init_map` takes the connectivity, land values, original
owners, capitals, canals and teams that the reference compiles in from src/land.cpp, sea.cpp, canal.cpp, player.cpp
(dimensions stay 5 / 3 / 2 / 5).

The oracle for a second board is the reference engine itself, compiled by oracle/build_ref.py with data TUs generated
from tests/golden/alt_map.json in the reference's own initialiser format (oracle/_ref/libcaaa_ref_alt.so); the golden
file tests/golden/alt_map.npz was produced from it (tests/golden/make_alt_map_golden.py).
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest

from caaa_b200 import gpu, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def alt():
    return layout.map_from_json(os.path.join(GOLD, "alt_map.json")), np.load(os.path.join(GOLD, "alt_map.npz"))


def test_default_map_is_the_reference_board():
    m = gpu.default_map()
    assert m["land_value"].tolist() == [10, 8, 10, 8, 8] and m["land_owner"].tolist() == [4, 2, 1, 0, 3]   # src/land.cpp:4-20
    assert m["capital"].tolist() == [3, 2, 1, 4, 0]                                                         # src/player.cpp:9-13
    assert m["canal_seas"].tolist() == [[0, 1], [0, 2]] and m["canal_lands"].tolist() == [[0, 1], [1, 2]]   # src/canal.cpp:3-6
    assert m["sea_land"][0][:2].tolist() == [0, 4] and m["land_sea"][1][:2].tolist() == [1, 2]
    t = np.zeros(1, dtype=layout.TABLES_DTYPE)
    t[0] = gpu.tables_for_map(m)
    gold = np.load(os.path.join(GOLD, "tables.npy"))  # the reference's own tables, byte for byte (make_golden.py)
    assert np.array_equal(t.view(np.uint8).reshape(-1), gold)


def test_tables_of_the_second_board_match_the_reference_build(alt):
    m, g = alt
    t = gpu.tables_for_map(m)
    for name in layout.TABLES_DTYPE.names:
        assert np.array_equal(np.asarray(t[name]), np.asarray(g["tables"][0][name])), name
    from oracle import pyref
    if os.path.exists(pyref.RefAlt.PATH):  # and against the reference build itself where it exists
        tr = pyref.RefAlt().get_tables()
        for name in layout.TABLES_DTYPE.names:
            assert np.array_equal(np.asarray(t[name]), np.asarray(tr[name])), name


def test_invalid_maps_are_rejected(alt):
    lib = gpu.load_library()
    for field, idx, val in (("land_sea", (0, 0), 3), ("land_land", (1, 0), 1), ("land_land_count", (2,), 6), ("capital", (0,), 5),
                            ("canal_seas", (0, 0), 9), ("land_owner", (4,), 5), ("sea_sea", (0, 0), 0)):
        m = np.array(alt[0])
        m[field][idx] = val
        t = np.zeros(1, dtype=layout.TABLES_DTYPE)
        assert lib.caaa_gpu_tables_for_map(np.ascontiguousarray(m).ctypes.data, t.ctypes.data) == -3, field  # CAAA_E_ARG


def test_engine_on_the_second_board_host_build(alt, hostsim):
    """The device engine (compiled for the host, tests/hostsim) on the second board against the reference build's playouts."""
    m, g = alt
    hostsim.lib.caaa_hostsim_set_map.argtypes = [C.c_void_p]
    mb = np.ascontiguousarray(m)
    try:
        hostsim.lib.caaa_hostsim_set_map(mb.ctypes.data)
        hostsim.set_stream(1, int(g["seed"]))
        got = hostsim.rollout_many(g["start"], 0, 1200)
        assert np.array_equal(got, g["stats"][:1200])
        assert np.array_equal(hostsim.get_possible_actions(g["start"]), g["root_actions"])
        for a, child in zip(g["root_actions"], g["root_children"]):
            assert np.array_equal(hostsim.apply_action(g["start"], int(a)), child), a
    finally:  # the hostsim fixture is shared by the session: back to the reference's board
        d = np.ascontiguousarray(gpu.default_map())
        hostsim.lib.caaa_hostsim_set_map(d.ctypes.data)


_GPU_SCRIPT = """
import sys, numpy as np
sys.path.insert(0, %r)
from caaa_b200 import gpu, layout
g = np.load(%r)
eng = gpu.GpuEngine(0, game_map=layout.map_from_json(%r))
res, stats, summ = eng.rollout_batch(g["start"], 3000, int(g["seed"]), 0, want_stats=True)          # lockstep kernel
res2, stats2, summ2 = eng.rollout_batch(g["start"], 40000, int(g["seed"]), 0, want_stats=True)      # phase-regrouped kernel
na, acts, ch, st = eng.expand(g["start"].reshape(1, -1))
np.savez(sys.argv[1], stats=stats, stats2=stats2[:3000], flagged=int(summ2["flagged"]), n_actions=na, actions=acts, children=ch)
# a second initialisation with another board must be refused, not silently ignored
rc = eng.lib.caaa_gpu_init(0)
assert rc == -3, rc
"""


@pytest.mark.gpu
def test_engine_on_the_second_board_gpu(alt, tmp_path):
    """Both rollout kernels and the expansion kernels on the second board, bit-exact against the reference build.  (Own process:
    the library holds one board per process.)"""
    m, g = alt
    out = str(tmp_path / "alt.npz")
    code = _GPU_SCRIPT % (ROOT, os.path.join(GOLD, "alt_map.npz"), os.path.join(GOLD, "alt_map.json"))
    subprocess.check_call([sys.executable, "-c", code, out])
    r = np.load(out)
    assert np.array_equal(r["stats"], g["stats"]) and np.array_equal(r["stats2"], g["stats"])
    assert int(r["flagged"]) == 0
    n = int(r["n_actions"][0])
    assert np.array_equal(r["actions"][0][:n], g["root_actions"])
    assert np.array_equal(r["children"][0][:n], g["root_children"])
