"""ctypes bindings for the two CPU checkers.  TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py CPU legs).

* ``Ref``  -- the patched reference itself (oracle/_ref/libcaaa_ref.so, built by oracle/build_ref.py
  from /root/reference; prebuilt .so travels to the GPU box).
* ``Port`` -- our plain-C restatement (oracle/libcaaa_oracle.so, built by oracle/Makefile).

Both expose the same calls, so tests can run one against the other.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
STATE_BYTES = 670
LIVE_BYTES = 606
N_PHASES = 21


class RolloutStats(C.Structure):
    _fields_ = [("result", C.c_double), ("turns", C.c_int32), ("decisions", C.c_int32),
                ("draws", C.c_int32), ("final_player", C.c_int32), ("state_hash", C.c_uint64)]


STATS_DTYPE = np.dtype([("result", "<f8"), ("turns", "<i4"), ("decisions", "<i4"),
                        ("draws", "<i4"), ("final_player", "<i4"), ("state_hash", "<u8")])

CACHE_DTYPE = np.dtype([
    ("income_per_turn", "<i4", (6,)), ("total_factory_count", "<i4", (6,)),
    ("factory_locations", "<i4", (6, 5)), ("total_player_land_units", "<i4", (6, 5)),
    ("total_player_sea_units", "<i4", (6, 3)), ("transports_with_large_cargo_space", "<i4", (3,)),
    ("transports_with_small_cargo_space", "<i4", (3,)), ("allied_carriers", "<i4", (3,)),
    ("enemy_units_count", "<i4", (8,)), ("answers_remaining", "<i4"),
    ("current_player_land_unit_types", "u1", (5, 6)), ("current_player_sea_unit_types", "u1", (3, 15)),
    ("enemy_blockade_total", "u1", (3,)), ("enemy_destroyers_total", "u1", (3,)),
    ("is_land_path_blocked", "u1", (25,)), ("is_sea_path_blocked", "u1", (9,)),
    ("is_sub_path_blocked", "u1", (9,)), ("enemies_0", "u1", (4,)), ("is_allied_0", "u1", (5,)),
    ("enemies_count_0", "u1"), ("canal_state", "u1"), ("valid_moves", "u1", (20,)),
    ("valid_moves_count", "u1"), ("pad_", "u1", (1,))])

TABLES_DTYPE = np.dtype([
    ("land_value", "u1", (5,)), ("original_owner", "u1", (5,)), ("capital", "u1", (5,)),
    ("is_allied", "u1", (5, 5)), ("canal_lands", "u1", (2, 2)), ("land_dist", "u1", (5, 8)),
    ("air_dist", "u1", (8, 8)), ("sea_dist", "u1", (4, 3, 3)), ("sea_path", "u1", (4, 3, 3)),
    ("sea_path_alt", "u1", (4, 3, 3)), ("land_path", "u1", (5, 8)), ("land_path_alt", "u1", (5, 8)),
    ("air_conn_count", "u1", (8,)), ("air_conn", "u1", (8, 7)), ("l2l_count", "u1", (5,)),
    ("l2l_conn", "u1", (5, 6)), ("l2s_count", "u1", (5,)), ("l2s_conn", "u1", (5, 4)),
    ("s2s_count", "u1", (3,)), ("s2s_conn", "u1", (3, 7)), ("s2l_count", "u1", (3,)),
    ("s2l_conn", "u1", (3, 6)), ("lands_within_2_count", "u1", (5,)), ("lands_within_2", "u1", (5, 4)),
    ("load_within_2_count", "u1", (5,)), ("load_within_2", "u1", (5, 3)),
    ("seas_within_1_count", "u1", (4, 3)), ("seas_within_1", "u1", (4, 3, 2)),
    ("seas_within_2_count", "u1", (4, 3)), ("seas_within_2", "u1", (4, 3, 2)),
    ("air_within_count", "u1", (6, 8)), ("air_within", "u1", (6, 8, 7)),
    ("air_to_land_within_count", "u1", (6, 8)), ("air_to_land_within", "u1", (6, 8, 5))])

# GameState as a numpy structured dtype (reference include/game_state.h:20-68)
LAND_DTYPE = np.dtype([("owner_idx", "u1"), ("factory_hp", "i1"), ("factory_max", "u1"), ("bombard_max", "u1"),
                       ("fighters", "u1", (5,)), ("bombers", "u1", (7,)), ("infantry", "u1", (2,)),
                       ("artillery", "u1", (2,)), ("tanks", "u1", (3,)), ("aa_guns", "u1", (2,))])
SEA_DTYPE = np.dtype([("fighters", "u1", (5,)), ("trans_empty", "u1", (4,)), ("trans_1i", "u1", (5,)),
                      ("trans_1a", "u1", (5,)), ("trans_1t", "u1", (5,)), ("trans_2i", "u1", (4,)),
                      ("trans_1i_1a", "u1", (4,)), ("trans_1i_1t", "u1", (4,)), ("submarines", "u1", (3,)),
                      ("destroyers", "u1", (2,)), ("carriers", "u1", (2,)), ("cruisers", "u1", (3,)),
                      ("battleships", "u1", (3,)), ("bs_damaged", "u1", (3,)), ("bombers", "u1", (5,))])
STATE_DTYPE = np.dtype([("player_index", "u1"), ("money", "u1", (5,)), ("builds_left", "u1", (8,)),
                        ("land_state", LAND_DTYPE, (5,)), ("units_sea", SEA_DTYPE, (3,)),
                        ("other_land_units", "u1", (4, 5, 6)), ("other_sea_units", "u1", (4, 3, 14)),
                        ("flagged_for_combat", "u1", (8,)), ("skipped_moves", "u1", (8, 8))])
assert STATE_DTYPE.itemsize == STATE_BYTES and LAND_DTYPE.itemsize == 25 and SEA_DTYPE.itemsize == 57


def start_state():
    """The reference's game_data.json start position (game_data.json:1-42) as 670 bytes."""
    s = np.zeros((), dtype=STATE_DTYPE)
    s["player_index"] = 0
    s["money"] = [10, 20, 2, 20, 2]
    s["builds_left"] = [0, 0, 0, 8, 0, 0, 0, 0]
    s["land_state"]["owner_idx"] = [4, 2, 1, 0, 3]
    s["land_state"]["factory_hp"] = [10, 8, 10, 8, 8]
    s["land_state"]["factory_max"] = [10, 8, 10, 8, 8]
    return np.frombuffer(s.tobytes(), dtype=np.uint8).copy()


def fnv1a64(buf):
    h = 0xcbf29ce484222325
    for b in bytes(buf):
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


class _Engine:
    """Common wrapper; `prefix` selects caaa_ref_* or caaa_oracle_* symbols."""

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = C.CDLL(path)
        self.p = prefix
        L = self.lib
        f = self._f
        f("init", None, [])
        f("set_stream", None, [C.c_int, C.c_uint64])
        f("set_reset_valid_moves", None, [C.c_int])
        f("set_game_id", None, [C.c_int64])
        f("rollout", C.c_double, [C.c_void_p, C.c_void_p])
        f("rollout_many", None, [C.c_void_p, C.c_int64, C.c_int, C.c_void_p, C.c_void_p])
        f("get_state", None, [C.c_void_p])
        f("get_possible_actions", C.c_int, [C.c_void_p, C.c_void_p])
        f("apply_action", None, [C.c_void_p, C.c_uint8])
        f("is_terminal", C.c_int, [C.c_void_p])
        f("load_state", None, [C.c_void_p, C.c_int64])
        f("run_phase", C.c_int, [C.c_int])
        f("evaluate", C.c_double, [])
        f("draws", C.c_int, [])
        f("get_cache", None, [C.c_void_p])
        f("get_tables", None, [C.c_void_p])
        self.init()

    def _f(self, name, res, args):
        fn = getattr(self.lib, self.p + name)
        fn.restype = res
        fn.argtypes = args
        setattr(self, "_" + name, fn)

    def init(self):
        self._init()

    def set_stream(self, mode, seed=0):
        self._set_stream(mode, seed)

    def set_reset_valid_moves(self, on):
        self._set_reset_valid_moves(1 if on else 0)

    def set_game_id(self, gid):
        self._set_game_id(gid)

    def rollout(self, state):
        st = np.zeros(1, dtype=STATS_DTYPE)
        state = np.ascontiguousarray(state, dtype=np.uint8)
        self._rollout(state.ctypes.data, st.ctypes.data)
        return st[0]

    def rollout_many(self, state, first_game_id, n):
        st = np.zeros(n, dtype=STATS_DTYPE)
        state = np.ascontiguousarray(state, dtype=np.uint8)
        self._rollout_many(state.ctypes.data, first_game_id, n, st.ctypes.data, None)
        return st

    def get_state(self):
        out = np.zeros(STATE_BYTES, dtype=np.uint8)
        self._get_state(out.ctypes.data)
        return out

    def get_possible_actions(self, state):
        state = np.ascontiguousarray(state, dtype=np.uint8)
        acts = np.zeros(20, dtype=np.uint8)
        n = self._get_possible_actions(state.ctypes.data, acts.ctypes.data)
        return acts[:n].copy()

    def apply_action(self, state, action):
        s = np.ascontiguousarray(state, dtype=np.uint8).copy()
        self._apply_action(s.ctypes.data, int(action))
        return s

    def is_terminal(self, state):
        state = np.ascontiguousarray(state, dtype=np.uint8)
        return bool(self._is_terminal(state.ctypes.data))

    def load_state(self, state, game_id=0):
        state = np.ascontiguousarray(state, dtype=np.uint8)
        self._load_state(state.ctypes.data, game_id)

    def run_phase(self, ph):
        return self._run_phase(ph)

    def evaluate(self):
        return self._evaluate()

    def draws(self):
        return self._draws()

    def get_cache(self):
        c = np.zeros(1, dtype=CACHE_DTYPE)
        self._get_cache(c.ctypes.data)
        return c[0]

    def get_tables(self):
        t = np.zeros(1, dtype=TABLES_DTYPE)
        self._get_tables(t.ctypes.data)
        return t[0]


class Ref(_Engine):
    PATH = os.path.join(HERE, "_ref", "libcaaa_ref.so")

    def __init__(self):
        super().__init__(self.PATH, "caaa_ref_")
        self.lib.caaa_ref_mcts.restype = C.c_int
        self.lib.caaa_ref_mcts.argtypes = [C.c_void_p, C.c_long, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self.lib.caaa_ref_get_random_table.argtypes = [C.c_void_p]
        self.lib.caaa_ref_set_playout_counter.argtypes = [C.c_uint64]

    def mcts(self, state, iterations):
        state = np.ascontiguousarray(state, dtype=np.uint8)
        acts = np.zeros(20, dtype=np.uint8)
        visits = np.zeros(20, dtype=np.int32)
        values = np.zeros(20, dtype=np.float64)
        best = np.zeros(1, dtype=np.uint8)
        n = self.lib.caaa_ref_mcts(state.ctypes.data, iterations, acts.ctypes.data, visits.ctypes.data,
                                   values.ctypes.data, best.ctypes.data)
        return acts[:n], visits[:n], values[:n], int(best[0])

    def random_table(self):
        out = np.zeros(65536, dtype=np.uint8)
        self.lib.caaa_ref_get_random_table(out.ctypes.data)
        return out


class RefAlt(Ref):
    """The reference engine compiled with the data TUs of the second board (tests/golden/alt_map.json)."""
    PATH = os.path.join(HERE, "_ref", "libcaaa_ref_alt.so")


class Port(_Engine):
    PATH = os.path.join(HERE, "libcaaa_oracle.so")

    def __init__(self):
        super().__init__(self.PATH, "caaa_oracle_")
        self.lib.caaa_oracle_stuck.restype = C.c_int

    def stuck(self):
        """True if the last call hit the sea-battle round guard (the reference would never return)."""
        return bool(self.lib.caaa_oracle_stuck())


class HostSim(_Engine):
    """The product's device engine compiled for the host (tests/hostsim) -- CPU test tier only."""
    PATH = os.path.join(os.path.dirname(HERE), "tests", "hostsim", "libcaaa_hostsim.so")

    def __init__(self):
        super().__init__(self.PATH, "caaa_hostsim_")
        self.lib.caaa_hostsim_stuck.restype = C.c_int

    def stuck(self):
        return bool(self.lib.caaa_hostsim_stuck())

    def rollout_many_sliced(self, state, first_game_id, n, budget):
        """Playouts in which every phase yields after `budget` loop iterations and resumes from its saved cursor."""
        st = np.zeros(n, dtype=STATS_DTYPE)
        buf = np.ascontiguousarray(state).view(np.uint8).reshape(-1)
        self.lib.caaa_hostsim_rollout_many_sliced(buf.ctypes.data_as(C.c_void_p), C.c_int64(first_game_id), C.c_int(n), C.c_int(budget),
                                                  st.ctypes.data_as(C.c_void_p))
        return st


    def rollout_many_merged(self, state, first_game_id, n, budget):
        """Playouts with the three consecutive land-unit phases run as one merged visit (budget = yield granularity, 0 = none)."""
        st = np.zeros(n, dtype=STATS_DTYPE)
        buf = np.ascontiguousarray(state).view(np.uint8).reshape(-1)
        self.lib.caaa_hostsim_rollout_many_merged(buf.ctypes.data_as(C.c_void_p), C.c_int64(first_game_id), C.c_int(n), C.c_int(budget),
                                                  st.ctypes.data_as(C.c_void_p))
        return st


def have_ref():
    return os.path.exists(Ref.PATH)


def have_port():
    return os.path.exists(Port.PATH)
