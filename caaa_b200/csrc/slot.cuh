// slot.cuh -- unit-type tables as immediates and the read-only topology tables, as they live in shared
// memory while a CTA steps its games.  The per-game working set itself is `Game` in game.cuh.
#pragma once
#include <stdint.h>
#include "../../include/caaa_types.h"

#if defined(__CUDACC__)
#define CAAA_HD __host__ __device__ __forceinline__
#define CAAA_HD_NOINLINE __host__ __device__ __noinline__
#else
#define CAAA_HD inline
#define CAAA_HD_NOINLINE
#endif

namespace caaa {

constexpr int NL = CAAA_LANDS, NS = CAAA_SEAS, NA = CAAA_AIRS, NP = CAAA_PLAYERS;
constexpr int NLT = CAAA_LAND_UNIT_TYPES, NST = CAAA_SEA_UNIT_TYPES;

// unit type ids, reference include/units/units.h:19-44
enum : int { FIGHTERS = 0, BOMBERS_LAND_AIR = 1, INFANTRY = 2, ARTILLERY = 3, TANKS = 4, AA_GUNS = 5 };
enum : int { TRANS_EMPTY = 1, TRANS_1I, TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, SUBMARINES,
             DESTROYERS, CARRIERS, CRUISERS, BATTLESHIPS, BS_DAMAGED, BOMBERS_SEA };

// Small per-unit-type tables (reference src/units.cpp) packed one byte per entry into 64-bit
// immediates: a lookup is a shift and a mask, costs no memory access, and constant-folds when
// the unit type is known at compile time.
CAAA_HD constexpr uint32_t pk8(int i, uint64_t lo, uint64_t hi = 0) {
  return (uint32_t)(((i < 8 ? lo : hi) >> (8 * (i & 7))) & 0xffu);
}
#define CAAA_PK(a, b, c, d, e, f, g, h)                                                                \
  ((uint64_t)(a) | ((uint64_t)(b) << 8) | ((uint64_t)(c) << 16) | ((uint64_t)(d) << 24) |              \
   ((uint64_t)(e) << 32) | ((uint64_t)(f) << 40) | ((uint64_t)(g) << 48) | ((uint64_t)(h) << 56))

CAAA_HD constexpr uint32_t off_land(int t) { return pk8(t, CAAA_PK(4, 9, 16, 18, 20, 23, 0, 0)); }
CAAA_HD constexpr uint32_t states_land(int t) { return pk8(t, CAAA_PK(5, 7, 2, 2, 3, 2, 0, 0)); }
CAAA_HD constexpr uint32_t cost_land(int t) { return pk8(t, CAAA_PK(10, 14, 3, 4, 5, 5, 0, 0)); }
CAAA_HD constexpr uint32_t weight_land(int t) { return pk8(t, CAAA_PK(255, 255, 2, 3, 3, 6, 0, 0)); }
CAAA_HD constexpr uint32_t max_move_land(int t) { return pk8(t, CAAA_PK(4, 6, 1, 1, 2, 1, 0, 0)); }
CAAA_HD constexpr uint32_t off_sea(int t) {
  return pk8(t, CAAA_PK(0, 5, 9, 14, 19, 24, 28, 32), CAAA_PK(36, 39, 41, 43, 46, 49, 52, 0));
}
CAAA_HD constexpr uint32_t states_sea(int t) {
  return pk8(t, CAAA_PK(5, 4, 5, 5, 5, 4, 4, 4), CAAA_PK(3, 2, 2, 3, 3, 3, 5, 0));
}
CAAA_HD constexpr uint32_t cost_sea(int t) {
  return pk8(t, CAAA_PK(10, 8, 11, 12, 13, 14, 15, 16), CAAA_PK(8, 8, 14, 10, 22, 22, 14, 0));
}
CAAA_HD constexpr uint32_t attack_sea(int t) {
  return pk8(t, CAAA_PK(3, 0, 0, 0, 0, 0, 0, 0), CAAA_PK(2, 2, 1, 3, 4, 4, 4, 0));
}
CAAA_HD constexpr uint32_t defense_sea(int t) {
  return pk8(t, CAAA_PK(4, 0, 0, 0, 0, 0, 0, 0), CAAA_PK(1, 2, 2, 3, 4, 4, 1, 0));
}
CAAA_HD constexpr uint32_t staging_sea(int t) { return pk8(t, CAAA_PK(0, 1, 1, 1, 1, 0, 0, 0), 0); }
CAAA_HD constexpr uint32_t unloading_sea(int t) { return pk8(t, CAAA_PK(0, 0, 1, 1, 1, 1, 1, 1), 0); }
CAAA_HD constexpr uint32_t unmoved_sea(int t) {
  return pk8(t, CAAA_PK(4, 1, 1, 1, 1, 1, 1, 1), CAAA_PK(2, 1, 1, 2, 2, 2, 255, 0));
}
CAAA_HD constexpr uint32_t done_moving_sea(int t) {
  return pk8(t, CAAA_PK(0, 0, 0, 0, 0, 0, 0, 0), CAAA_PK(0, 0, 0, 1, 1, 1, 0, 0));
}
CAAA_HD constexpr uint32_t unload_cargo1(int t) {
  return pk8(t, CAAA_PK(255, 255, INFANTRY, ARTILLERY, TANKS, INFANTRY, ARTILLERY, TANKS), ~0ull);
}
// LOAD_UNIT_TYPE rows for infantry / artillery / tanks (reference src/units.cpp:99-117), indices 0..4
CAAA_HD constexpr uint32_t load_result(int land_type, int trans_type) {
  return land_type == INFANTRY ? pk8(trans_type, CAAA_PK(255, TRANS_1I, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, 255, 255, 255))
         : land_type == ARTILLERY ? pk8(trans_type, CAAA_PK(255, TRANS_1A, TRANS_1I_1A, 255, 255, 255, 255, 255))
                                  : pk8(trans_type, CAAA_PK(255, TRANS_1T, TRANS_1I_1T, 255, 255, 255, 255, 255));
}
// casualty orders, reference src/engine.cpp:81-97
CAAA_HD constexpr uint32_t order_land_def(int k) { return pk8(k, CAAA_PK(AA_GUNS, BOMBERS_LAND_AIR, INFANTRY, ARTILLERY, TANKS, FIGHTERS, 0, 0)); }
CAAA_HD constexpr uint32_t order_sea_def(int k) {
  return pk8(k, CAAA_PK(SUBMARINES, DESTROYERS, CARRIERS, CRUISERS, FIGHTERS, BS_DAMAGED, TRANS_EMPTY, TRANS_1I),
             CAAA_PK(TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, 0, 0, 0));
}
CAAA_HD constexpr uint32_t order_sea_att3(int k) {
  return pk8(k, CAAA_PK(BS_DAMAGED, TRANS_EMPTY, TRANS_1I, TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T));
}
CAAA_HD constexpr uint32_t buy_sea(int k) { return pk8(k, CAAA_PK(FIGHTERS, TRANS_EMPTY, SUBMARINES, DESTROYERS, CARRIERS, CRUISERS, BATTLESHIPS, 0)); }

// sections of DevTables::scan
enum : int {
  SC_FIGHTERS = 0, SC_FIGHTERS_N = 8,
  SC_BOMBERS = SC_FIGHTERS + SC_FIGHTERS_N, SC_BOMBERS_N = 5,
  SC_STAGE = SC_BOMBERS + SC_BOMBERS_N, SC_STAGE_N = 12,
  SC_TANKS = SC_STAGE + SC_STAGE_N, SC_TANKS_N = 10,
  SC_ARTILLERY = SC_TANKS + SC_TANKS_N, SC_ARTILLERY_N = 5,
  SC_INFANTRY = SC_ARTILLERY + SC_ARTILLERY_N, SC_INFANTRY_N = 5,
  SC_AA_GUNS = SC_INFANTRY + SC_INFANTRY_N, SC_AA_GUNS_N = 5,
  SC_TRANSPORTS = SC_AA_GUNS + SC_AA_GUNS_N, SC_TRANSPORTS_N = 36,
  SC_SHIPS = SC_TRANSPORTS + SC_TRANSPORTS_N, SC_SHIPS_N = 15,
  SC_UNLOAD = SC_SHIPS + SC_SHIPS_N, SC_UNLOAD_N = 18,
  SC_LAND_FIGHTERS = SC_UNLOAD + SC_UNLOAD_N, SC_LAND_FIGHTERS_N = 24,
  SC_LAND_BOMBERS = SC_LAND_FIGHTERS + SC_LAND_FIGHTERS_N, SC_LAND_BOMBERS_N = 40,
  SC_SUBS = SC_LAND_BOMBERS + SC_LAND_BOMBERS_N, SC_SUBS_N = 3,
  SCAN_TOTAL = SC_SUBS + SC_SUBS_N
};

// Read-only tables staged once per CTA in shared memory: the reference's topology tables plus
// per-player-to-move team lists (refresh_allies, reference src/engine.cpp:716-725, depends only
// on player_index, so it is tabulated instead of cached per game).
struct DevTables {
  CaaaTables t;
  uint8_t allied_mask[NP];    // bit p set <=> relative player p is allied with the player to move
  uint8_t n_enemies[NP];
  uint8_t enemies[NP][4];     // relative indices of the enemies, ascending (casualty order depends on it)
  uint8_t pad_[2];
  // the same lists in absolute player ids (engine2: per-player data is stored in absolute order, never rotated)
  uint8_t allied_abs[NP];     // bit a set <=> absolute player a is allied with the player to move
  uint8_t enemies_abs[NP][4]; // absolute ids of the enemies, in the relative order above
  uint8_t pad2_[3];
  uint32_t recip16[32];       // ceil(2^16 / n): byte % n without a division (n = 0 is never used)
  // Cell lists of the move phases, in the reference's loop order: one entry per (unit type, location, movement
  // state) a phase looks at = byte offset of that count inside a Game (bits 0..9) | type << 10 | location << 14 |
  // state << 17.  Finding the next non-empty cell is then one table read + one byte read per cell.
  uint32_t scan[SCAN_TOTAL];
};

// status bits reported per game (0 = clean)
enum : uint32_t {
  CAAA_ERR_SEA_ROUND_CAP = 1u,   // a sea battle exceeded SEA_ROUND_CAP rounds (the reference would never return)
  CAAA_ERR_LOAD_FAILED = 2u,     // load_transport found no transport (the reference prints and loops)
  CAAA_ERR_NO_DECISION = 4u,     // expansion mode: no player reached a decision within EXPAND_TURN_CAP turns (the reference would spin)
};
constexpr int EXPAND_TURN_CAP = 5000;
constexpr int SEA_ROUND_CAP = 10000;

static_assert(sizeof(CaaaGameState) == 670, "GameState layout");

}  // namespace caaa
