"""The drop-in boundary, proven with the reference's OWN translation units.

The reference's search (`src/mcts.cpp`) and CLI driver (`src/main.cpp`) bind the engine through plain C++
externs (reference src/mcts.cpp:16-22, include/engine.h:29,75,85,150-159).  These tests compile those two files
and the reference's static-data TUs exactly as they are (from `oracle/_ref/src`, the copy `oracle/build_ref.py`
makes; mcts.cpp and main.cpp are unpatched), link them against `libcaaa_gpu.so` instead of the reference's
`engine.cpp`, and -- on the GPU box -- run the resulting `caaa` binary.

CPU tier: the link has no undefined reference (the mangled names match, e.g.
`_Z26random_play_until_terminalP9GameState`).  GPU tier: `./caaa 2000` searches and prints a best action.
"""
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
TUS = ["main.cpp", "mcts.cpp", "land.cpp", "sea.cpp", "canal.cpp", "player.cpp", "units.cpp"]


def _have_ref_sources():
    return all(os.path.exists(os.path.join(REF, "src", t)) for t in TUS)


def _link(tmp_path):
    lib_dir = os.path.join(ROOT, "caaa_b200")
    if not os.path.exists(os.path.join(lib_dir, "libcaaa_gpu.so")):
        import __graft_entry__
        __graft_entry__.build()
    exe = str(tmp_path / "caaa")
    cmd = ["g++", "-std=c++11", "-O2", "-w", f"-I{os.path.join(REF, 'include')}"] + \
          [os.path.join(REF, "src", t) for t in TUS] + \
          [f"-L{lib_dir}", "-lcaaa_gpu", f"-Wl,-rpath,{lib_dir}", "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, "the reference's mcts.cpp / main.cpp do not link against libcaaa_gpu.so:\n" + r.stderr
    return exe


@pytest.mark.skipif(not _have_ref_sources(), reason="oracle/_ref/src not built (run oracle/build_ref.py where /root/reference exists)")
def test_reference_tus_link_against_the_library(tmp_path):
    exe = _link(tmp_path)
    # every engine symbol the reference's objects import is defined by the library, under the reference's mangling
    undefined = subprocess.run(["nm", "-u", "-C", exe], capture_output=True, text=True).stdout
    for name in ("random_play_until_terminal(GameState*)", "get_possible_actions(GameState*, unsigned char*, unsigned char (*) [20])",
                 "apply_action(GameState*, unsigned char)", "clone_state(GameState*)", "is_terminal_state(GameState*)",
                 "initialize_constants()", "load_game_data(char const*)", "get_game_state_copy()",
                 "refresh_full_cache()", "load_single_game()"):
        assert name in undefined, name  # imported from the shared library ...
    exported = subprocess.run(["nm", "-D", "-C", "--defined-only", os.path.join(ROOT, "caaa_b200", "libcaaa_gpu.so")],
                              capture_output=True, text=True).stdout
    for name in ("random_play_until_terminal(GameState*)", "apply_action(GameState*, unsigned char)", "load_single_game()"):
        assert name in exported, name   # ... which defines it


@pytest.mark.gpu
@pytest.mark.skipif(not _have_ref_sources(), reason="oracle/_ref/src not shipped")
def test_reference_cli_runs_on_the_gpu(tmp_path):
    """`./caaa 2000` (reference src/main.cpp:10-34): 2000 MCTS iterations, every rollout / expansion on the GPU."""
    exe = _link(tmp_path)
    subprocess.check_call(["cp", os.path.join(REF, "game_data.json"), str(tmp_path / "game_data.json")])
    r = subprocess.run([exe, "2000"], cwd=str(tmp_path), capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    m = re.search(r"Best action: (\d+)", r.stdout)
    assert m is not None, r.stdout[-2000:]
    assert int(m.group(1)) in (6, 5, 4, 3, 2, 0)  # the root actions of the start position
