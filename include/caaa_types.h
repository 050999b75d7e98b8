/* caaa_types.h -- plain-old-data layouts shared by the C-ABI (caaa_gpu.h), the host code and
 * the test oracles.  C99 / C++11 / CUDA clean; no torch, no STL.
 *
 * Map dimensions are the reference's compile-time macros and are pinned at exactly these
 * values (reference include/land.h:7, sea.h:7, canal.h:6-7, player.h:7).
 */
#ifndef CAAA_TYPES_H
#define CAAA_TYPES_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CAAA_LANDS 5   /* LANDS_COUNT   include/land.h:7   */
#define CAAA_SEAS 3    /* SEAS_COUNT    include/sea.h:7    */
#define CAAA_AIRS 8    /* AIRS_COUNT    include/game_state.h:13 */
#define CAAA_PLAYERS 5 /* PLAYERS_COUNT include/player.h:7 */
#define CAAA_CANALS 2
#define CAAA_CANAL_STATES 4
#define CAAA_ACTIONS 20 /* ACTIONS_COUNT include/game_state.h:16 */
#define CAAA_LAND_UNIT_TYPES 6
#define CAAA_SEA_UNIT_TYPES 15

/* ---- GameState: byte-for-byte the reference's struct (include/game_state.h:20-68).
 * 670 bytes, alignof 1; offsets are SURVEY.md Appendix C and are asserted in tests. ---- */
typedef struct CaaaLandState {
  uint8_t owner_idx;    /* relative to the player to move; rotates each turn */
  int8_t factory_hp;    /* may go negative down to -factory_max */
  uint8_t factory_max;
  uint8_t bombard_max;
  uint8_t fighters[5];  /* index = moves left */
  uint8_t bombers[7];
  uint8_t infantry[2];
  uint8_t artillery[2];
  uint8_t tanks[3];
  uint8_t aa_guns[2];
} CaaaLandState; /* 25 bytes */

typedef struct CaaaUnitsSea {
  uint8_t fighters[5];
  uint8_t trans_empty[4];
  uint8_t trans_1i[5];
  uint8_t trans_1a[5];
  uint8_t trans_1t[5];
  uint8_t trans_2i[4];
  uint8_t trans_1i_1a[4];
  uint8_t trans_1i_1t[4];
  uint8_t submarines[3];
  uint8_t destroyers[2];
  uint8_t carriers[2];
  uint8_t cruisers[3];
  uint8_t battleships[3];
  uint8_t bs_damaged[3];
  uint8_t bombers[5];
} CaaaUnitsSea; /* 57 bytes */

/* the member list, shared with caaa_engine_compat.h (which must declare the same bytes under the
 * reference's own type name so that its C++ symbols mangle like the reference's) */
#define CAAA_GAME_STATE_MEMBERS                                                            \
  uint8_t player_index;                                                                    \
  uint8_t money[CAAA_PLAYERS];                                                             \
  uint8_t builds_left[CAAA_AIRS];                                                          \
  CaaaLandState land_state[CAAA_LANDS];                                                    \
  CaaaUnitsSea units_sea[CAAA_SEAS];                                                       \
  uint8_t other_land_units[CAAA_PLAYERS - 1][CAAA_LANDS][CAAA_LAND_UNIT_TYPES];            \
  uint8_t other_sea_units[CAAA_PLAYERS - 1][CAAA_SEAS][CAAA_SEA_UNIT_TYPES - 1];           \
  uint8_t flagged_for_combat[CAAA_AIRS];                                                   \
  uint8_t skipped_moves[CAAA_AIRS][CAAA_AIRS]; /* one byte per flag, bit 0 only */

typedef struct CaaaGameState {
  CAAA_GAME_STATE_MEMBERS
} CaaaGameState; /* 670 bytes */

#define CAAA_STATE_BYTES 670
#define CAAA_STATE_LIVE_BYTES 606 /* everything before skipped_moves (zero between plies) */

/* ---- Topology tables as the reference's initialize_constants() computes them
 * (src/engine.cpp:165-568), including its uint8 overflow quirks. ---- */
typedef struct CaaaTables {
  uint8_t land_value[5];
  uint8_t original_owner[5];      /* LANDS[].original_owner_index, src/land.cpp:4-20 */
  uint8_t capital[5];             /* PLAYERS[].capital_territory_index, src/player.cpp:9-13 */
  uint8_t is_allied[5][5];        /* PLAYERS[].is_allied */
  uint8_t canal_lands[2][2];      /* CANALS[].lands, src/canal.cpp:3-6 */
  uint8_t land_dist[5][8];
  uint8_t air_dist[8][8];
  uint8_t sea_dist[4][3][3];
  uint8_t sea_path[4][3][3];
  uint8_t sea_path_alt[4][3][3];
  uint8_t land_path[5][8];
  uint8_t land_path_alt[5][8];
  uint8_t air_conn_count[8];
  uint8_t air_conn[8][7];
  uint8_t l2l_count[5];
  uint8_t l2l_conn[5][6];
  uint8_t l2s_count[5];
  uint8_t l2s_conn[5][4];
  uint8_t s2s_count[3];
  uint8_t s2s_conn[3][7];
  uint8_t s2l_count[3];
  uint8_t s2l_conn[3][6];
  uint8_t lands_within_2_count[5];
  uint8_t lands_within_2[5][4];
  uint8_t load_within_2_count[5];
  uint8_t load_within_2[5][3];
  uint8_t seas_within_1_count[4][3];
  uint8_t seas_within_1[4][3][2];
  uint8_t seas_within_2_count[4][3];
  uint8_t seas_within_2[4][3][2];
  uint8_t air_within_count[6][8];
  uint8_t air_within[6][8][7];
  uint8_t air_to_land_within_count[6][8];
  uint8_t air_to_land_within[6][8][5];
} CaaaTables;

/* ---- The board as data: what the reference compiles in from src/land.cpp:4-20, src/sea.cpp:5-8,
 * src/canal.cpp:3-6 and src/player.cpp:9-13.  Dimensions stay the reference's compile-time macros (5 lands, 3 seas,
 * 2 canals, 5 players); the connectivity, values, owners, capitals and teams are loadable (caaa_gpu_init_map).
 * Neighbour lists are directional exactly as in the reference (a land's sea list and a sea's land list are
 * separate data and need not mirror each other: they do not on the reference's own map). ---- */
typedef struct CaaaMap {
  uint8_t land_owner[CAAA_LANDS];        /* Land.original_owner_index */
  uint8_t land_value[CAAA_LANDS];        /* Land.land_value */
  uint8_t land_sea_count[CAAA_LANDS];    /* Land.sea_conn_count  (<= 4) */
  uint8_t land_land_count[CAAA_LANDS];   /* Land.land_conn_count (<= 6) */
  uint8_t land_sea[CAAA_LANDS][4];       /* Land.connected_sea_index */
  uint8_t land_land[CAAA_LANDS][6];      /* Land.connected_land_index */
  uint8_t sea_sea_count[CAAA_SEAS];      /* Sea.sea_conn_count  (<= 7) */
  uint8_t sea_land_count[CAAA_SEAS];     /* Sea.land_conn_count (<= 6) */
  uint8_t sea_sea[CAAA_SEAS][7];         /* Sea.connected_sea_index */
  uint8_t sea_land[CAAA_SEAS][6];        /* Sea.connected_land_index */
  uint8_t canal_seas[CAAA_CANALS][2];    /* Canal.seas */
  uint8_t canal_lands[CAAA_CANALS][2];   /* Canal.lands */
  uint8_t capital[CAAA_PLAYERS];         /* Player.capital_territory_index */
  uint8_t is_allied[CAAA_PLAYERS][CAAA_PLAYERS]; /* Player.is_allied */
} CaaaMap;

/* ---- Derived per-game caches (the reference's file-scope globals, src/engine.cpp:120-159),
 * in a fixed wire layout so tests can compare implementations field by field. ---- */
typedef struct CaaaCacheDump {
  int32_t income_per_turn[6];
  int32_t total_factory_count[6];
  int32_t factory_locations[6][5];
  int32_t total_player_land_units[6][5];
  int32_t total_player_sea_units[6][3];
  int32_t transports_with_large_cargo_space[3];
  int32_t transports_with_small_cargo_space[3];
  int32_t allied_carriers[3];
  int32_t enemy_units_count[8];
  int32_t answers_remaining;
  uint8_t current_player_land_unit_types[5][6];
  uint8_t current_player_sea_unit_types[3][15];
  uint8_t enemy_blockade_total[3];
  uint8_t enemy_destroyers_total[3];
  uint8_t is_land_path_blocked[25];
  uint8_t is_sea_path_blocked[9];
  uint8_t is_sub_path_blocked[9];
  uint8_t enemies_0[4];
  uint8_t is_allied_0[5];
  uint8_t enemies_count_0;
  uint8_t canal_state;
  uint8_t valid_moves[CAAA_ACTIONS];
  uint8_t valid_moves_count;
  uint8_t pad_[1];
} CaaaCacheDump;

/* Per-playout side outputs used for parity checks. */
typedef struct CaaaRolloutStats {
  double result;        /* the reference's return value (Allies' share of material) */
  int32_t turns;        /* player-turns (plies) executed */
  int32_t decisions;    /* 100000 - answers_remaining (src/engine.cpp:1292,4198) */
  int32_t draws;        /* random bytes consumed */
  int32_t final_player; /* state.player_index at the end */
  uint64_t state_hash;  /* FNV-1a-64 over the first 606 bytes of the final GameState */
} CaaaRolloutStats;

/* Phase ids = position in the reference's fixed pipeline (src/engine.cpp:4214-4235). */
enum CaaaPhase {
  CAAA_PH_MOVE_FIGHTERS = 0,
  CAAA_PH_MOVE_BOMBERS = 1,
  CAAA_PH_STAGE_TRANSPORTS = 2,
  CAAA_PH_MOVE_TANKS = 3,
  CAAA_PH_MOVE_ARTILLERY = 4,
  CAAA_PH_MOVE_INFANTRY = 5,
  CAAA_PH_MOVE_TRANSPORTS = 6,
  CAAA_PH_MOVE_SUBS = 7,
  CAAA_PH_MOVE_SHIPS = 8,
  CAAA_PH_SEA_BATTLES = 9,
  CAAA_PH_UNLOAD_TRANSPORTS = 10,
  CAAA_PH_LAND_BATTLES = 11,
  CAAA_PH_MOVE_AA_GUNS = 12,
  CAAA_PH_LAND_FIGHTERS = 13,
  CAAA_PH_LAND_BOMBERS = 14,
  CAAA_PH_BUY_UNITS = 15,
  CAAA_PH_CRASH_AIR = 16,
  CAAA_PH_RESET_UNITS = 17,
  CAAA_PH_BUY_FACTORY = 18,
  CAAA_PH_COLLECT_MONEY = 19,
  CAAA_PH_ROTATE_TURNS = 20,
  CAAA_PH_COUNT = 21
};

#define CAAA_FNV_OFFSET 0xcbf29ce484222325ULL
#define CAAA_FNV_PRIME 0x100000001b3ULL

#ifdef __cplusplus
}
#endif
#endif /* CAAA_TYPES_H */
