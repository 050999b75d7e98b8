#!/usr/bin/env python3
"""Build the patched reference (`oracle/_ref/`) from the sources under /root/reference.

TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is on the product path.

The reference HEAD does not compile, so this recipe copies the
reference translation units into the git-ignored `oracle/_ref/src/`, applies three small,
named patch sets to the *copies*, and compiles them together with `oracle/ref_shim.cpp`
(our own glue, not reference code) into

    oracle/_ref/libcaaa_ref.so      C-ABI wrapper around the reference engine
    oracle/_ref/ref_cli             small CLI used by bench.py --impl reference

Patch sets (every edit asserts that the expected reference text is present):

  P0  compile/link fixes only (9 sites), no behaviour change.
  P1  canonical meaning for layout-dependent out-of-bounds accesses and per-playout reset
      of the global `valid_moves` buffer (game independence).
  P2  the three random-byte draw sites and the per-playout stream reset call out to
      `caaa_ref_next_byte()` / `caaa_ref_begin_playout()` in ref_shim.cpp, which serve
      either the reference's own glibc table stream (mode 0) or Philox4x32-10 keyed by
      (seed, game_id, draw_index) (mode 1).

No reference source is ever written into the tracked tree; `oracle/_ref/` is git-ignored.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("CAAA_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "_ref")
SRC = os.path.join(OUT, "src")
INC = os.path.join(OUT, "include")

ENGINE_TUS = ["engine.cpp", "mcts.cpp", "land.cpp", "sea.cpp", "canal.cpp", "player.cpp", "units.cpp"]


def sub(text, old, new, count=1, where=""):
    n = text.count(old)
    if n != count:
        raise SystemExit(f"patch {where!r}: expected {count} occurrence(s) of {old!r}, found {n}")
    return text.replace(old, new)


def patch_engine(t):
    # ---- P0: compile fixes (engine.cpp:20,192,251,321,669,696,697,737,3447,4338-4341) ----
    t = sub(t, "#include <pybind11/pybind11.h>\n", '#include "caaa_ref_hooks.h"\n', where="P0 pybind include")
    t = sub(t, "std::fill(std::begin(LAND_DIST), std::end(LAND_DIST), MAX_UINT8_T);",
            "memset(&LAND_DIST, MAX_UINT8_T, sizeof(LAND_DIST));", where="P0 192")
    t = sub(t, "std::fill(std::begin(SEA_DIST), std::end(SEA_DIST), MAX_UINT8_T);",
            "memset(&SEA_DIST, MAX_UINT8_T, sizeof(SEA_DIST));", where="P0 251")
    t = sub(t, "std::fill(std::begin(AIR_DIST), std::end(AIR_DIST), MAX_UINT8_T);",
            "memset(&AIR_DIST, MAX_UINT8_T, sizeof(AIR_DIST));", where="P0 321")
    t = sub(t, "std::fill(std::begin(total_player_land_units), std::end(total_player_land_units), 0);",
            "memset(&total_player_land_units, 0, sizeof(total_player_land_units));", where="P0 669")
    t = sub(t, "std::fill(std::begin(current_player_sea_unit_types), std::end(current_player_sea_unit_types), 0);",
            "memset(&current_player_sea_unit_types, 0, sizeof(current_player_sea_unit_types));", where="P0 696")
    t = sub(t, "std::fill(std::begin(total_player_sea_units), std::end(total_player_sea_units), 0);",
            "memset(&total_player_sea_units, 0, sizeof(total_player_sea_units));", where="P0 697")
    t = sub(t, "std::copy(std::begin(SEA_PATH[0][canal_state]), std::end(SEA_PATH[0][canal_state]),",
            "std::copy(std::begin(SEA_PATH[canal_state]), std::end(SEA_PATH[canal_state]),", where="P0 737")
    t = sub(t, "memset(canFighterLandHere, 0, sizeof(canFighterLandHere));",
            "memset(canFighterLandHere.data(), 0, sizeof(canFighterLandHere));", where="P0 3447")
    t = sub(t, 'PYBIND11_MODULE(engine, handle) {\n  handle.doc() = "engine doc";\n'
               '  handle.def("random_play_until_terminal", &random_play_until_terminal);\n}\n',
            "", where="P0 pybind module")

    # ---- P1: canonical out-of-bounds meaning + game independence ----
    # (i) the byte table is filled and read for 65 536 entries (engine.cpp:98,110,514-519)
    t = sub(t, "std::array<uint8_t, RANDOM_NUMBERS_SIZE> RANDOM_NUMBERS = {0};",
            "std::array<uint8_t, 65536> RANDOM_NUMBERS = {0};", where="P1i")
    # (ii) skipped_moves is a flat 64-flag array; flat index >= 64 reads 0 / drops the write
    t = sub(t, "state.skipped_moves[src_air][valid_action].bit = true;",
            "CAAA_SKIP_SET(src_air, valid_action);", where="P1ii 1316")
    t = sub(t, "if (state.skipped_moves[dst_air][i].bit) {\n      state.skipped_moves[src_air][i].bit = true;",
            "if (CAAA_SKIP_GET(dst_air, i)) {\n      CAAA_SKIP_SET(src_air, i);", where="P1ii 1324")
    t = sub(t, "if (state.skipped_moves[0][unit_type].bit) {\n            last_purchased = unit_type;",
            "if (CAAA_SKIP_GET(0, unit_type)) {\n            last_purchased = unit_type;", where="P1ii 3707")
    t = sub(t, "for (uint8_t unit_type2 = last_purchased; unit_type2 < purchase; unit_type2++) {\n"
               "            state.skipped_moves[0][unit_type2].bit = true;",
            "for (uint8_t unit_type2 = last_purchased; unit_type2 < purchase; unit_type2++) {\n"
            "            CAAA_SKIP_SET(0, unit_type2);", where="P1ii 3756")
    # (iii) sea_dist[src_sea][dst_air] and is_land_path_blocked[src][dst_air] are flat reads
    t = sub(t, "uint8_t sea_distance = sea_dist[src_sea][dst_air];",
            "uint8_t sea_distance = CAAA_FLAT_U8(sea_dist, src_sea * SEAS_COUNT + dst_air, SEAS_COUNT * SEAS_COUNT);",
            where="P1iii 1630")
    t = sub(t, "if (land_dist[dst_air] == 1 && is_land_path_blocked_src_land[dst_air]) {",
            "if (land_dist[dst_air] == 1 && CAAA_FLAT_U8(is_land_path_blocked, src_land * LANDS_COUNT + dst_air, LANDS_COUNT * LANDS_COUNT)) {",
            where="P1iii 1474")
    # (iv) playouts must not see the previous playout's valid_moves leftovers; P2 stream reset
    t = sub(t, "  random_number_index = (uint16_t)(rand() % RANDOM_NUMBERS_SIZE);\n",
            "  random_number_index = caaa_ref_begin_playout((uint16_t)RANDOM_NUMBERS_SIZE);\n"
            "  if (caaa_ref_reset_valid_moves) { memset(valid_moves, 0, sizeof(valid_moves)); valid_moves_count = 0; }\n",
            where="P1iv/P2 4202")

    # ---- P2: draw sites (engine.cpp:1296 (+1288 DEBUG twin), 2562, 2577) ----
    t = sub(t, "return valid_moves[RANDOM_NUMBERS[random_number_index++] % valid_moves_count];",
            "return valid_moves[caaa_ref_next_byte() % valid_moves_count];", count=2, where="P2 1296")
    t = sub(t, "(RANDOM_NUMBERS[random_number_index++] % 6 < attacker_damage % 6 ? 1 : 0);",
            "(caaa_ref_next_byte() % 6 < attacker_damage % 6 ? 1 : 0);", where="P2 2562")
    t = sub(t, "(RANDOM_NUMBERS[random_number_index++] % 6 < defender_damage % 6 ? 1 : 0);",
            "(caaa_ref_next_byte() % 6 < defender_damage % 6 ? 1 : 0);", where="P2 2577")
    return t


def patch_land_h(t):
    # P0: `using Land = struct {...}` has no linkage name, so LANDS does not link across TUs.
    return sub(t, "using Land = struct {", "struct Land {", where="P0 land.h:12")


HOOKS_H = r"""// generated by oracle/build_ref.py -- hooks the patched reference calls (defined in ref_shim.cpp)
#pragma once
#include <math.h>   // fmax (the reference got it transitively through pybind11)
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
uint8_t caaa_ref_next_byte();
uint16_t caaa_ref_begin_playout(uint16_t table_size);
extern bool caaa_ref_reset_valid_moves;
#define CAAA_SKIP_FLAT(a, b) ((unsigned)(a) * AIRS_COUNT_PAREN + (unsigned)(b))
#define AIRS_COUNT_PAREN (LANDS_COUNT + SEAS_COUNT)
#define CAAA_SKIP_GET(a, b) \
  (CAAA_SKIP_FLAT(a, b) < 64u ? (bool)(((BitField*)state.skipped_moves)[CAAA_SKIP_FLAT(a, b)].bit) : false)
#define CAAA_SKIP_SET(a, b) \
  do { if (CAAA_SKIP_FLAT(a, b) < 64u) ((BitField*)state.skipped_moves)[CAAA_SKIP_FLAT(a, b)].bit = true; } while (0)
#define CAAA_FLAT_U8(arr, idx, n) ((unsigned)(idx) < (unsigned)(n) ? ((const uint8_t*)&(arr))[(unsigned)(idx)] : (uint8_t)0)
"""

CJSON_STUB = """// stand-in for the system libcjson header (not installed here); declarations only
#pragma once
typedef struct cJSON cJSON;
"""


def write_alt_map_sources(map_json, dst):
    """The reference's four static-data TUs (src/land.cpp, sea.cpp, canal.cpp, player.cpp) for ANOTHER board of the same
    dimensions, generated from tests/golden/alt_map.json in the reference's own initialiser format.  The engine, the
    search and the unit tables are the reference's; only the data differs."""
    import json
    with open(map_json) as f:
        d = json.load(f)
    pad = lambda xs, n: ", ".join(str(x) for x in (list(xs) + [0] * n)[:n])
    land = ['#include "land.h"', "#include <array>", "const std::array<Land, LANDS_COUNT> LANDS = {{"]
    land.append(",\n".join('  {"%s", %d, %d, %d, %d, {%s}, {%s}}' % (l["name"], l["owner"], l["value"], len(l["seas"]), len(l["lands"]),
                                                                     pad(l["seas"], 4), pad(l["lands"], 6)) for l in d["lands"]))
    land.append("}};")
    sea = ['#include "sea.h"', "const Sea SEAS[SEAS_COUNT] = {"]
    sea.append(",\n".join('  {(char*)"%s", %d, %d, {%s}, {%s}}' % (s["name"], len(s["seas"]), len(s["lands"]), pad(s["seas"], 7), pad(s["lands"], 6))
                          for s in d["seas"]))
    sea.append("};")
    canal = ['#include "canal.h"', "const Canal CANALS[CANALS_COUNT] = {"]
    canal.append(",\n".join('  {"%s", {%s}, {%s}}' % (c["name"], pad(c["seas"], 2), pad(c["lands"], 2)) for c in d["canals"]))
    canal.append("};")
    axis = d["axis"]
    player = ['#include "player.h"', "const Player PLAYERS[PLAYERS_COUNT] = {"]
    rows = []
    for p in range(5):
        allied = ", ".join("true" if axis[p] == axis[q] else "false" for q in range(5))
        rows.append('  {"P%d", "P%d", "", %d, %d, {%s}, false}' % (p, p, d["capitals"][p], 1 if axis[p] else 0, allied))
    player.append(",\n".join(rows))
    player.append("};")
    os.makedirs(dst, exist_ok=True)
    for name, lines in (("land.cpp", land), ("sea.cpp", sea), ("canal.cpp", canal), ("player.cpp", player)):
        with open(os.path.join(dst, name), "w") as f:
            f.write("\n".join(lines) + "\n")


def main():
    if not os.path.isdir(os.path.join(REF, "src")):
        print(f"build_ref: {REF} not present; keeping any prebuilt oracle/_ref as is", file=sys.stderr)
        return 0
    shutil.rmtree(SRC, ignore_errors=True)
    shutil.rmtree(INC, ignore_errors=True)
    os.makedirs(SRC)
    shutil.copytree(os.path.join(REF, "include"), INC)
    os.makedirs(os.path.join(INC, "cjson"), exist_ok=True)
    with open(os.path.join(INC, "cjson", "cJSON.h"), "w") as f:
        f.write(CJSON_STUB)
    with open(os.path.join(INC, "caaa_ref_hooks.h"), "w") as f:
        f.write(HOOKS_H)
    for tu in ENGINE_TUS:
        with open(os.path.join(REF, "src", tu)) as f:
            t = f.read()
        if tu == "engine.cpp":
            t = patch_engine(t)
        with open(os.path.join(SRC, tu), "w") as f:
            f.write(t)
    # the reference's CLI driver, unpatched: tests/test_reference_links.py links it (with mcts.cpp and the data
    # TUs above) against libcaaa_gpu.so to prove the drop-in boundary; it is not part of libcaaa_ref.so / ref_cli
    shutil.copy(os.path.join(REF, "src", "main.cpp"), os.path.join(SRC, "main.cpp"))
    os.chmod(os.path.join(SRC, "main.cpp"), 0o644)
    shutil.copy(os.path.join(REF, "game_data.json"), os.path.join(OUT, "game_data.json"))
    os.chmod(os.path.join(OUT, "game_data.json"), 0o644)
    p = os.path.join(INC, "land.h")
    os.chmod(p, 0o644)
    with open(p) as f:
        t = f.read()
    with open(p, "w") as f:
        f.write(patch_land_h(t))

    cxx = os.environ.get("CXX", "g++")
    flags = ["-std=c++11", "-O3", "-fPIC", "-fno-strict-aliasing", "-w", f"-I{INC}"]
    srcs = [os.path.join(SRC, tu) for tu in ENGINE_TUS] + [os.path.join(HERE, "ref_shim.cpp")]
    so = os.path.join(OUT, "libcaaa_ref.so")
    subprocess.check_call([cxx] + flags + ["-shared", "-o", so] + srcs + ["-lm"])
    cli = os.path.join(OUT, "ref_cli")
    subprocess.check_call([cxx] + flags + ["-o", cli, os.path.join(HERE, "ref_cli.cpp")] + srcs + ["-lm"])
    # the same engine with the data TUs of a second board (tests/golden/alt_map.json): the oracle of the map-as-data path
    alt_json = os.path.join(os.path.dirname(HERE), "tests", "golden", "alt_map.json")
    if os.path.exists(alt_json):
        alt_src = os.path.join(OUT, "alt_src")
        shutil.rmtree(alt_src, ignore_errors=True)
        write_alt_map_sources(alt_json, alt_src)
        alt = [os.path.join(SRC, tu) for tu in ("engine.cpp", "mcts.cpp", "units.cpp")] + \
              [os.path.join(alt_src, tu) for tu in ("land.cpp", "sea.cpp", "canal.cpp", "player.cpp")] + [os.path.join(HERE, "ref_shim.cpp")]
        subprocess.check_call([cxx] + flags + ["-shared", "-o", os.path.join(OUT, "libcaaa_ref_alt.so")] + alt + ["-lm"])
    print(f"build_ref: built {so} and {cli}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
