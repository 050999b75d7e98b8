"""CPU tier: the C-ABI library loads and exports every symbol include/caaa_gpu.h declares, fails
loudly without a device (no CPU fallback), and the GameState layout is the reference's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from caaa_b200 import gpu, layout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _have_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def test_header_symbols_exported():
    if not os.path.exists(gpu.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    lib = gpu.load_library()
    from caaa_b200 import host
    header = open(os.path.join(ROOT, "include", "caaa_gpu.h")).read()
    declared = sorted(set(re.findall(r"\b(caaa_gpu_\w+)\s*\(", header)))
    assert declared == sorted(gpu.SYMBOLS)
    header2 = open(os.path.join(ROOT, "include", "caaa_host.h")).read()
    declared2 = sorted(set(re.findall(r"\b(caaa_(?:state|mcts)_\w+)\s*\(", header2)))
    assert declared2 == sorted(host.SYMBOLS)
    declared3 = sorted(set(re.findall(r"\b(caaa_root_\w+)\s*\(", header)))
    assert declared3 == sorted(gpu.ROOT_SYMBOLS)
    for name in declared + declared2 + declared3:
        assert hasattr(lib, name), name
    # the reference-named C++ entry points of caaa_engine_compat.h (Itanium-mangled)
    for name in ("_Z20initialize_constantsv", "_Z14load_game_dataPKc", "_Z19get_game_state_copyv",
                 "_Z11clone_stateP9GameState", "_Z10free_stateP9GameState",
                 "_Z20get_possible_actionsP9GameStatePhPA20_h", "_Z12apply_actionP9GameStateh",
                 "_Z17is_terminal_stateP9GameState", "_Z14evaluate_stateP9GameState",
                 "_Z26random_play_until_terminalP9GameState"):
        assert hasattr(lib, name), name


@pytest.mark.skipif(_have_cuda(), reason="only meaningful on a box without a GPU")
def test_no_cpu_fallback():
    lib = gpu.load_library()
    assert lib.caaa_gpu_init(0) == -1  # CAAA_E_NO_DEVICE
    assert b"no CPU path" in lib.caaa_gpu_last_error()
    out = np.zeros(1)
    st = layout.start_state()
    assert lib.caaa_gpu_rollout_batch(st.ctypes.data, 1, 1, 0, 0, out.ctypes.data, None, None) == -4  # NOT_INIT
    with pytest.raises(gpu.CaaaGpuError):
        gpu.GpuEngine()


def test_gamestate_layout_matches_reference():
    """SURVEY.md Appendix C (probe offsetof on the reference's include/game_state.h:20-68)."""
    d = layout.STATE_DTYPE
    off = {n: d.fields[n][1] for n in d.names}
    assert off == {"player_index": 0, "money": 1, "builds_left": 6, "land_state": 14, "units_sea": 139,
                   "other_land_units": 310, "other_sea_units": 430, "flagged_for_combat": 598, "skipped_moves": 606}
    assert d.itemsize == 670
    lo = {n: layout.LAND_DTYPE.fields[n][1] for n in layout.LAND_DTYPE.names}
    assert lo == {"owner_idx": 0, "factory_hp": 1, "factory_max": 2, "bombard_max": 3, "fighters": 4, "bombers": 9,
                  "infantry": 16, "artillery": 18, "tanks": 20, "aa_guns": 23}
    so = {n: layout.SEA_DTYPE.fields[n][1] for n in layout.SEA_DTYPE.names}
    assert so == {"fighters": 0, "trans_empty": 5, "trans_1i": 9, "trans_1a": 14, "trans_1t": 19, "trans_2i": 24,
                  "trans_1i_1a": 28, "trans_1i_1t": 32, "submarines": 36, "destroyers": 39, "carriers": 41,
                  "cruisers": 43, "battleships": 46, "bs_damaged": 49, "bombers": 52}


def test_reference_struct_sizes(ref):
    assert ref.lib.caaa_ref_sizeof_state() == 670
    assert ref.lib.caaa_ref_sizeof_node() == 704


def test_host_tables_match_golden():
    """The product's host table builder (caaa_gpu_get_tables needs no device) against the reference's tables."""
    lib = gpu.load_library()
    t = np.zeros(1, dtype=layout.TABLES_DTYPE)
    assert lib.caaa_gpu_get_tables(t.ctypes.data) == 0
    gold = np.load(os.path.join(ROOT, "tests", "golden", "tables.npy"))
    assert np.array_equal(np.frombuffer(t.tobytes(), dtype=np.uint8), gold)
    # spot values from SURVEY.md Appendix D
    assert t[0]["land_dist"][0].tolist() == [0, 243, 251, 250, 249, 1, 1, 0]
    assert t[0]["air_dist"][3].tolist() == [3, 3, 1, 2, 1, 2, 3, 2]
