"""ctypes binding of libcaaa_gpu.so (include/caaa_gpu.h).  No game logic lives here."""
import ctypes as C
import os

import numpy as np

from .layout import CACHE_DTYPE, MAP_DTYPE, STATE_BYTES, STATS_DTYPE, SUMMARY_DTYPE, TABLES_DTYPE

LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libcaaa_gpu.so")

# every symbol include/caaa_gpu.h declares
SYMBOLS = ["caaa_gpu_init", "caaa_gpu_init_map", "caaa_gpu_default_map", "caaa_gpu_tables_for_map", "caaa_gpu_shutdown", "caaa_gpu_last_error", "caaa_gpu_get_tables",
           "caaa_gpu_device_info", "caaa_gpu_rollout_batch", "caaa_gpu_rollout_batch_device",
           "caaa_gpu_synchronize", "caaa_gpu_set_stream", "caaa_gpu_random_play_until_terminal", "caaa_gpu_advance",
           "caaa_gpu_resolve_battles", "caaa_gpu_resolve_battles_device", "caaa_gpu_expand", "caaa_gpu_evaluate",
           "caaa_gpu_soa_bytes", "caaa_gpu_pack", "caaa_gpu_unpack", "caaa_gpu_leaf_batch_submit", "caaa_gpu_leaf_batch_collect", "caaa_gpu_set_rollout_sms"]


# the root-parallel exchange step (NCCL, bound at run time)
ROOT_SYMBOLS = ["caaa_root_allreduce", "caaa_root_unique_id", "caaa_root_comm_init", "caaa_root_allreduce_host",
                "caaa_root_comm_destroy", "caaa_root_last_error"]


class CaaaGpuError(RuntimeError):
    pass


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise CaaaGpuError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(there is no CPU fallback)")
    lib = C.CDLL(path)
    vp, i32, u64 = C.c_void_p, C.c_int, C.c_uint64
    lib.caaa_gpu_init.argtypes = [i32]
    lib.caaa_gpu_init_map.argtypes = [i32, vp]
    lib.caaa_gpu_default_map.argtypes = [vp]
    lib.caaa_gpu_tables_for_map.argtypes = [vp, vp]
    lib.caaa_gpu_last_error.restype = C.c_char_p
    lib.caaa_gpu_get_tables.argtypes = [vp]
    lib.caaa_gpu_device_info.argtypes = [vp, vp, vp]
    lib.caaa_gpu_rollout_batch.argtypes = [vp, i32, i32, u64, u64, vp, vp, vp]
    lib.caaa_gpu_rollout_batch_device.argtypes = [vp, i32, i32, u64, u64, vp, vp, vp, vp]
    lib.caaa_gpu_set_stream.argtypes = [u64, u64]
    lib.caaa_gpu_random_play_until_terminal.argtypes = [vp]
    lib.caaa_gpu_random_play_until_terminal.restype = C.c_double
    lib.caaa_gpu_advance.argtypes = [vp, i32, u64, u64, i32, vp, vp, vp]
    lib.caaa_gpu_resolve_battles.argtypes = [vp, i32, u64, u64, vp, vp, vp]
    lib.caaa_gpu_resolve_battles_device.argtypes = [vp, i32, u64, u64, vp, vp, vp, vp]
    lib.caaa_gpu_expand.argtypes = [vp, i32, vp, vp, vp, vp]
    lib.caaa_gpu_evaluate.argtypes = [vp, i32, vp]
    lib.caaa_gpu_soa_bytes.argtypes = [i32, vp]
    lib.caaa_gpu_pack.argtypes = [vp, i32, vp, vp]
    lib.caaa_gpu_unpack.argtypes = [vp, i32, vp, vp]
    lib.caaa_gpu_leaf_batch_submit.argtypes = [i32, vp, i32, i32, u64, u64, vp]
    lib.caaa_gpu_leaf_batch_collect.argtypes = [i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.caaa_gpu_set_rollout_sms.argtypes = [i32]
    lib.caaa_root_allreduce.argtypes = [vp, i32, vp, vp, vp]
    lib.caaa_root_unique_id.argtypes = [vp]
    lib.caaa_root_comm_init.argtypes = [vp, i32, i32]
    lib.caaa_root_allreduce_host.argtypes = [i32, vp, vp, vp]
    lib.caaa_root_last_error.restype = C.c_char_p
    return lib


def _states(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    if a.ndim == 1:
        a = a.reshape(1, -1)
    if a.shape[-1] != STATE_BYTES:
        raise ValueError("states must be (n, 670) uint8")
    return a


def default_map(lib=None):
    """The reference's compiled-in board as a CaaaMap record (host only; no GPU needed)."""
    lib = lib or load_library()
    m = np.zeros(1, dtype=MAP_DTYPE)
    if lib.caaa_gpu_default_map(m.ctypes.data) != 0:
        raise CaaaGpuError("caaa_gpu_default_map failed")
    return m[0]


def tables_for_map(game_map, lib=None):
    """initialize_constants() for a board (host only); raises on an invalid map."""
    lib = lib or load_library()
    m = np.ascontiguousarray(game_map)
    t = np.zeros(1, dtype=TABLES_DTYPE)
    rc = lib.caaa_gpu_tables_for_map(m.ctypes.data, t.ctypes.data)
    if rc != 0:
        raise CaaaGpuError(f"caaa_gpu_tables_for_map error {rc}: {lib.caaa_gpu_last_error().decode()}")
    return t[0]


class GpuEngine:
    """One context per process (the library keeps one context per device)."""

    def __init__(self, device=0, game_map=None):
        self.lib = load_library()
        if game_map is None:
            self._check(self.lib.caaa_gpu_init(device))
        else:
            m = np.ascontiguousarray(game_map)
            self._check(self.lib.caaa_gpu_init_map(device, m.ctypes.data))

    def shutdown(self):
        self._check(self.lib.caaa_gpu_shutdown())

    def _check(self, rc):
        if rc != 0:
            raise CaaaGpuError(f"libcaaa_gpu error {rc}: {self.lib.caaa_gpu_last_error().decode()}")

    def device_info(self):
        sm, smem = C.c_int(0), C.c_int(0)
        name = C.create_string_buffer(64)
        self._check(self.lib.caaa_gpu_device_info(C.byref(sm), C.byref(smem), name))
        return {"sm_count": sm.value, "smem_per_block": smem.value, "name": name.value.decode()}

    def tables(self):
        t = np.zeros(1, dtype=TABLES_DTYPE)
        self._check(self.lib.caaa_gpu_get_tables(t.ctypes.data))
        return t[0]

    def rollout_batch(self, states, rollouts_per_state, seed, first_game_id=0, want_stats=False):
        states = _states(states)
        n = states.shape[0] * int(rollouts_per_state)
        results = np.empty(n, dtype=np.float64)
        stats = np.zeros(n, dtype=STATS_DTYPE) if want_stats else None
        summary = np.zeros(1, dtype=SUMMARY_DTYPE)
        self._check(self.lib.caaa_gpu_rollout_batch(states.ctypes.data, states.shape[0], int(rollouts_per_state),
                                                    int(seed), int(first_game_id), results.ctypes.data,
                                                    stats.ctypes.data if want_stats else None, summary.ctypes.data))
        return results, stats, summary[0]

    def rollout_batch_device(self, d_states_ptr, n_states, rollouts_per_state, seed, first_game_id, d_results_ptr,
                             d_stats_ptr=None, d_summary_ptr=None, stream_ptr=None):
        """All pointers are raw device addresses (ints), e.g. torch.Tensor.data_ptr(); asynchronous on the stream."""
        self._check(self.lib.caaa_gpu_rollout_batch_device(d_states_ptr, int(n_states), int(rollouts_per_state),
                                                           int(seed), int(first_game_id), d_results_ptr,
                                                           d_stats_ptr, d_summary_ptr, stream_ptr))

    def synchronize(self):
        self._check(self.lib.caaa_gpu_synchronize())

    def set_stream(self, seed, next_game_id=0):
        self._check(self.lib.caaa_gpu_set_stream(int(seed), int(next_game_id)))

    def random_play_until_terminal(self, state):
        state = _states(state)
        return float(self.lib.caaa_gpu_random_play_until_terminal(state.ctypes.data))

    def advance(self, states, seed, first_game_id, n_turns):
        states = _states(states)
        n = states.shape[0]
        out = np.empty_like(states)
        caches = np.zeros(n, dtype=CACHE_DTYPE)
        draws = np.zeros(n, dtype=np.int32)
        self._check(self.lib.caaa_gpu_advance(states.ctypes.data, n, int(seed), int(first_game_id), int(n_turns),
                                              out.ctypes.data, caches.ctypes.data, draws.ctypes.data))
        return out, caches, draws

    def resolve_battles(self, states, seed, first_game_id=0, out=None, draws=None):
        """`out` / `draws`: optional caller-owned result arrays (e.g. page-locked ones, which the library then fills by
        DMA without its staging copy; the same holds for a page-locked `states`)."""
        states = _states(states)
        n = states.shape[0]
        out = np.empty_like(states) if out is None else out
        draws = np.zeros(n, dtype=np.int32) if draws is None else draws
        assert out.shape == states.shape and out.dtype == np.uint8 and out.flags.c_contiguous
        assert draws.shape == (n,) and draws.dtype == np.int32 and draws.flags.c_contiguous
        summary = np.zeros(1, dtype=SUMMARY_DTYPE)
        self._check(self.lib.caaa_gpu_resolve_battles(states.ctypes.data, n, int(seed), int(first_game_id),
                                                      out.ctypes.data, draws.ctypes.data, summary.ctypes.data))
        return out, draws, summary[0]

    def resolve_battles_device(self, d_states_ptr, n, seed, first_game_id, d_out_ptr=None, d_draws_ptr=None, d_summary_ptr=None,
                               stream_ptr=None):
        """Raw device addresses (ints), asynchronous on the stream; d_summary is accumulated into."""
        self._check(self.lib.caaa_gpu_resolve_battles_device(d_states_ptr, int(n), int(seed), int(first_game_id), d_out_ptr,
                                                             d_draws_ptr, d_summary_ptr, stream_ptr))

    def expand(self, leaves, want_children=True):
        leaves = _states(leaves)
        n = leaves.shape[0]
        n_actions = np.zeros(n, dtype=np.uint8)
        actions = np.zeros((n, 20), dtype=np.uint8)
        children = np.zeros((n, 20, STATE_BYTES), dtype=np.uint8) if want_children else None
        status = np.zeros(n, dtype=np.uint32)
        self._check(self.lib.caaa_gpu_expand(leaves.ctypes.data, n, n_actions.ctypes.data, actions.ctypes.data,
                                             children.ctypes.data if want_children else None, status.ctypes.data))
        return n_actions, actions, children, status

    def leaf_batch_submit(self, handle, leaves, rollouts_per_leaf, seed, first_game_id, child_pick):
        leaves = _states(leaves)
        pick = np.ascontiguousarray(child_pick, dtype=np.uint64)
        assert len(pick) == leaves.shape[0]
        self._check(self.lib.caaa_gpu_leaf_batch_submit(int(handle), leaves.ctypes.data, leaves.shape[0], int(rollouts_per_leaf),
                                                        int(seed), int(first_game_id), pick.ctypes.data))
        return leaves.shape[0]

    def leaf_batch_collect(self, handle, n):
        out = {"scores": np.zeros(n), "n_actions": np.zeros(n, np.uint8), "actions": np.zeros((n, 20), np.uint8),
               "children": np.zeros((n, 20, STATE_BYTES), np.uint8), "status": np.zeros(n, np.uint32),
               "picked": np.zeros(n, np.int32), "sums": np.zeros(n), "summary": np.zeros(1, dtype=SUMMARY_DTYPE)}
        self._check(self.lib.caaa_gpu_leaf_batch_collect(int(handle), out["scores"].ctypes.data, out["n_actions"].ctypes.data,
                                                         out["actions"].ctypes.data, out["children"].ctypes.data,
                                                         out["status"].ctypes.data, out["picked"].ctypes.data,
                                                         out["sums"].ctypes.data, out["summary"].ctypes.data))
        out["summary"] = out["summary"][0]
        return out

    def soa_bytes(self, n):
        b = C.c_size_t(0)
        self._check(self.lib.caaa_gpu_soa_bytes(int(n), C.byref(b)))
        return b.value

    def pack(self, d_aos_ptr, n, d_soa_ptr, stream_ptr=None):
        self._check(self.lib.caaa_gpu_pack(d_aos_ptr, int(n), d_soa_ptr, stream_ptr))

    def unpack(self, d_soa_ptr, n, d_aos_ptr, stream_ptr=None):
        self._check(self.lib.caaa_gpu_unpack(d_soa_ptr, int(n), d_aos_ptr, stream_ptr))

    def set_rollout_sms(self, sms):
        self._check(self.lib.caaa_gpu_set_rollout_sms(int(sms)))

    def evaluate(self, states):
        states = _states(states)
        scores = np.zeros(states.shape[0], dtype=np.float64)
        self._check(self.lib.caaa_gpu_evaluate(states.ctypes.data, states.shape[0], scores.ctypes.data))
        return scores
