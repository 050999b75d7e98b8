"""ctypes binding of the host-side C-ABI (include/caaa_host.h): game_data.json I/O and the MCTS driver."""
import ctypes as C

import numpy as np

from . import gpu
from .layout import STATE_BYTES

SYMBOLS = ["caaa_state_from_json_file", "caaa_state_from_json_text", "caaa_state_to_json_text",
           "caaa_state_to_json_file", "caaa_mcts_search", "caaa_mcts_search_pipelined"]


class RootStats(C.Structure):
    _fields_ = [("n_children", C.c_int32), ("actions", C.c_uint8 * 20), ("visits", C.c_int64 * 20),
                ("values", C.c_double * 20), ("iterations", C.c_int64), ("playouts", C.c_int64), ("nodes", C.c_int64)]


def _lib():
    lib = gpu.load_library()
    lib.caaa_state_from_json_file.argtypes = [C.c_char_p, C.c_void_p]
    lib.caaa_state_from_json_text.argtypes = [C.c_char_p, C.c_void_p]
    lib.caaa_state_to_json_text.argtypes = [C.c_void_p, C.c_char_p, C.c_int]
    lib.caaa_state_to_json_file.argtypes = [C.c_char_p, C.c_void_p]
    lib.caaa_mcts_search.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
    lib.caaa_mcts_search_pipelined.argtypes = [C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_uint64, C.c_void_p, C.c_void_p]
    return lib


def load_game_data(path, base=None):
    """game_data.json -> 670-byte state (uint8 array); fields absent from the file keep `base` (zeros by default)."""
    st = np.zeros(STATE_BYTES, dtype=np.uint8) if base is None else np.array(base, dtype=np.uint8).copy()
    rc = _lib().caaa_state_from_json_file(str(path).encode(), st.ctypes.data)
    if rc != 0:
        raise OSError(f"caaa_state_from_json_file({path!r}) failed with {rc}")
    return st


def state_from_json(text, base=None):
    st = np.zeros(STATE_BYTES, dtype=np.uint8) if base is None else np.array(base, dtype=np.uint8).copy()
    rc = _lib().caaa_state_from_json_text(text.encode(), st.ctypes.data)
    if rc != 0:
        raise ValueError(f"caaa_state_from_json_text failed with {rc}")
    return st


def state_to_json(state):
    state = np.ascontiguousarray(state, dtype=np.uint8)
    lib = _lib()
    need = lib.caaa_state_to_json_text(state.ctypes.data, None, 0)
    buf = C.create_string_buffer(need)
    lib.caaa_state_to_json_text(state.ctypes.data, buf, need)
    return buf.value.decode()


def mcts_search(engine, state, iterations, rollouts_per_leaf=1, leaves_per_batch=1, seed=0, depth=None):
    """caaa_mcts_search (depth=None) / caaa_mcts_search_pipelined on an initialised GpuEngine;
    returns (actions, visits, values, best_action, info)."""
    state = np.ascontiguousarray(state, dtype=np.uint8)
    out = RootStats()
    best = C.c_uint8(255)
    if depth is None:
        engine._check(_lib().caaa_mcts_search(state.ctypes.data, int(iterations), int(rollouts_per_leaf), int(leaves_per_batch),
                                              int(seed), C.byref(out), C.byref(best)))
    else:
        engine._check(_lib().caaa_mcts_search_pipelined(state.ctypes.data, int(iterations), int(rollouts_per_leaf),
                                                        int(leaves_per_batch), int(depth), int(seed), C.byref(out), C.byref(best)))
    n = out.n_children
    info = {"iterations": out.iterations, "playouts": out.playouts, "nodes": out.nodes}
    return (np.array(out.actions[:n], dtype=np.uint8), np.array(out.visits[:n], dtype=np.int64),
            np.array(out.values[:n], dtype=np.float64), int(best.value), info)
