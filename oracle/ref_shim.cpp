// ref_shim.cpp -- TEST INFRASTRUCTURE.  Our own glue around the *patched reference* engine
// built by oracle/build_ref.py into oracle/_ref/.  Not reference code and not on the product
// path: only tests/, __graft_entry__.smoke() and bench.py's CPU legs load the resulting .so.
//
// It provides (a) the hooks the P1/P2 patches call, (b) a stand-in for the cJSON loader, and
// (c) an extern "C" API so Python/ctypes can drive the reference's own functions:
// random_play_until_terminal (src/engine.cpp:4184), get_possible_actions (4050), apply_action
// (4112), the 21 phase functions (4214-4235), evaluate_state (4301) and mcts_search
// (src/mcts.cpp:65), and dump its init tables and derived caches.
#include "engine.h"
#include "mcts.h"
#include "caaa_ref_hooks.h"
#include "../include/caaa_types.h"
#include "philox_ref.h"
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>

using std::array;

// ---- engine.cpp globals (external linkage there) ----
extern GameState state;
extern array<uint8_t, 65536> RANDOM_NUMBERS;
extern uint16_t random_number_index;
extern Actions valid_moves;
extern uint8_t valid_moves_count;
extern int answers_remaining;
extern bool use_selected_action;
extern int max_loops;
extern uint8_t current_player_land_unit_types[LANDS_COUNT][LAND_UNIT_TYPES_COUNT];
extern uint8_t current_player_sea_unit_types[SEAS_COUNT][SEA_UNIT_TYPES_COUNT];
extern array<int, PLAYERS_COUNT + 1> income_per_turn;
extern array<int, PLAYERS_COUNT + 1> total_factory_count;
extern array<array<int, LANDS_COUNT>, PLAYERS_COUNT + 1> factory_locations;
extern array<array<int, LANDS_COUNT>, PLAYERS_COUNT + 1> total_player_land_units;
extern array<array<int, SEAS_COUNT>, PLAYERS_COUNT + 1> total_player_sea_units;
extern array<uint8_t, SEAS_COUNT> enemy_blockade_total;
extern array<uint8_t, SEAS_COUNT> enemy_destroyers_total;
extern array<array<bool, LANDS_COUNT>, LANDS_COUNT> is_land_path_blocked;
extern array<array<bool, SEAS_COUNT>, SEAS_COUNT> is_sea_path_blocked;
extern array<array<bool, SEAS_COUNT>, SEAS_COUNT> is_sub_path_blocked;
extern array<int, SEAS_COUNT> transports_with_large_cargo_space;
extern array<int, SEAS_COUNT> transports_with_small_cargo_space;
extern array<uint8_t, PLAYERS_COUNT - 1> enemies_0;
extern array<bool, PLAYERS_COUNT> is_allied_0;
extern uint8_t enemies_count_0;
extern uint8_t canal_state;
extern array<int, SEAS_COUNT> allied_carriers;
extern array<int, AIRS_COUNT> enemy_units_count;
// init tables
extern array<uint8_t, LANDS_COUNT> LAND_VALUE;
extern array<array<uint8_t, AIRS_COUNT>, LANDS_COUNT> LAND_DIST;
extern array<uint8_t, AIRS_COUNT> AIR_CONN_COUNT;
extern array<array<uint8_t, MAX_AIR_TO_AIR_CONNECTIONS>, AIRS_COUNT> AIR_CONNECTIONS;
extern array<array<uint8_t, AIRS_COUNT>, AIRS_COUNT> AIR_DIST;
extern LandPath LAND_PATH;
extern LandPath LAND_PATH_ALT;
extern array<array<uint8_t, LANDS_COUNT - 1>, LANDS_COUNT> LANDS_WITHIN_2_MOVES;
extern array<uint8_t, LANDS_COUNT> LANDS_WITHIN_2_MOVES_COUNT;
extern array<array<uint8_t, SEAS_COUNT>, LANDS_COUNT> LOAD_WITHIN_2_MOVES;
extern array<uint8_t, LANDS_COUNT> LOAD_WITHIN_2_MOVES_COUNT;
extern array<array<array<uint8_t, SEAS_COUNT - 1>, SEAS_COUNT>, CANAL_STATES> SEAS_WITHIN_1_MOVE;
extern array<array<uint8_t, SEAS_COUNT>, CANAL_STATES> SEAS_WITHIN_1_MOVE_COUNT;
extern array<array<array<uint8_t, SEAS_COUNT - 1>, SEAS_COUNT>, CANAL_STATES> SEAS_WITHIN_2_MOVES;
extern array<array<uint8_t, SEAS_COUNT>, CANAL_STATES> SEAS_WITHIN_2_MOVES_COUNT;
extern array<array<array<uint8_t, AIRS_COUNT - 1>, AIRS_COUNT>, 6> AIR_WITHIN_X_MOVES;
extern array<array<uint8_t, AIRS_COUNT>, 6> AIR_WITHIN_X_MOVES_COUNT;
extern array<array<array<uint8_t, LANDS_COUNT>, AIRS_COUNT>, 6> AIR_TO_LAND_WITHIN_X_MOVES;
extern array<array<uint8_t, AIRS_COUNT>, 6> AIR_TO_LAND_WITHIN_X_MOVES_COUNT;
extern array<array<array<uint8_t, SEAS_COUNT>, SEAS_COUNT>, CANAL_STATES> SEA_DIST;
extern array<array<array<uint8_t, SEAS_COUNT>, SEAS_COUNT>, CANAL_STATES> SEA_PATH;
extern array<array<array<uint8_t, SEAS_COUNT>, SEAS_COUNT>, CANAL_STATES> SEA_PATH_ALT;
extern array<uint8_t, LANDS_COUNT> LAND_TO_LAND_COUNT;
extern array<array<uint8_t, MAX_LAND_TO_LAND_CONNECTIONS>, LANDS_COUNT> LAND_TO_LAND_CONN;
extern array<uint8_t, LANDS_COUNT> LAND_TO_SEA_COUNT;
extern array<array<uint8_t, MAX_LAND_TO_SEA_CONNECTIONS>, LANDS_COUNT> LAND_TO_SEA_CONN;
extern array<uint8_t, SEAS_COUNT> SEA_TO_SEA_COUNT;
extern array<array<uint8_t, MAX_SEA_TO_SEA_CONNECTIONS>, SEAS_COUNT> SEA_TO_SEA_CONN;
extern array<uint8_t, SEAS_COUNT> SEA_TO_LAND_COUNT;
extern array<array<uint8_t, MAX_SEA_TO_LAND_CONNECTIONS>, SEAS_COUNT> SEA_TO_LAND_CONN;

static_assert(sizeof(GameState) == CAAA_STATE_BYTES, "reference GameState must be 670 bytes");
static_assert(sizeof(CaaaGameState) == CAAA_STATE_BYTES, "CaaaGameState must be 670 bytes");

// ---- hooks called by the patched engine ----
bool caaa_ref_reset_valid_moves = true;   // P1(iv)
static int g_stream_mode = 0;             // 0 = reference glibc table stream, 1 = Philox (P2)
static uint64_t g_seed = 0;
static uint64_t g_game_id = 0;
static bool g_auto_game_id = true;        // game id = running playout counter
static uint64_t g_playout_counter = 0;
static uint32_t g_draws = 0;

static uint32_t g_block[4];               // current 16-byte Philox block (one block serves 16 draws, as on the GPU)
static uint32_t g_block_index = 0xffffffffu;
static uint64_t g_block_game = 0, g_block_seed = 0;

uint8_t caaa_ref_next_byte() {
  uint32_t d = g_draws++;
  if (g_stream_mode == 0) {
    return RANDOM_NUMBERS[random_number_index++];  // src/engine.cpp:1296,2562,2577 verbatim
  }
  if ((d >> 4) != g_block_index || g_block_game != g_game_id || g_block_seed != g_seed) {
    const uint32_t ctr[4] = {d >> 4, (uint32_t)g_game_id, (uint32_t)(g_game_id >> 32), 0u};
    const uint32_t key[2] = {(uint32_t)g_seed, (uint32_t)(g_seed >> 32)};
    caaa_philox4x32_10_ref(ctr, key, g_block);
    g_block_index = d >> 4;
    g_block_game = g_game_id;
    g_block_seed = g_seed;
  }
  const uint32_t k = d & 15u;
  return (uint8_t)(g_block[k >> 2] >> (8u * (k & 3u)));
}

uint16_t caaa_ref_begin_playout(uint16_t table_size) {
  g_draws = 0;
  if (g_auto_game_id) g_game_id = g_playout_counter;
  g_playout_counter++;
  if (g_stream_mode == 0) {
    return (uint16_t)(rand() % table_size);  // src/engine.cpp:4202 verbatim
  }
  return 0;
}

// ---- stand-in for src/serialize_data.cpp (needs cJSON, which is not installed) ----
void load_game_data_from_json(char const* filename, GameState* data) {
  (void)filename;
  memset(data, 0, sizeof(GameState));
}

static uint64_t fnv1a(const uint8_t* p, size_t n) {
  uint64_t h = CAAA_FNV_OFFSET;
  for (size_t i = 0; i < n; i++) {
    h ^= p[i];
    h *= CAAA_FNV_PRIME;
  }
  return h;
}

static bool g_inited = false;

extern "C" {

void caaa_ref_init(void) {
  if (g_inited) return;
  initialize_constants();  // consumes the first 65 536 rand() values (src/engine.cpp:516-520)
  g_inited = true;
}

void caaa_ref_set_stream(int mode, uint64_t seed) {
  g_stream_mode = mode;
  g_seed = seed;
}
void caaa_ref_set_reset_valid_moves(int on) { caaa_ref_reset_valid_moves = on != 0; }
void caaa_ref_set_game_id(int64_t game_id) {  // < 0: automatic running counter
  if (game_id < 0) {
    g_auto_game_id = true;
  } else {
    g_auto_game_id = false;
    g_game_id = (uint64_t)game_id;
  }
}
void caaa_ref_set_playout_counter(uint64_t v) { g_playout_counter = v; }

static void fill_stats(double result, CaaaRolloutStats* st) {
  if (!st) return;
  st->result = result;
  st->turns = max_loops < 0 ? 1000 : 1000 - max_loops;
  st->decisions = 100000 - answers_remaining;
  st->draws = (int32_t)g_draws;
  st->final_player = state.player_index;
  st->state_hash = fnv1a((const uint8_t*)&state, CAAA_STATE_LIVE_BYTES);
}

// random_play_until_terminal on a copy of *state670 (the reference never modifies its input)
double caaa_ref_rollout(const void* state670, CaaaRolloutStats* st) {
  GameState tmp;
  memcpy(&tmp, state670, sizeof(tmp));
  double r = random_play_until_terminal(&tmp);
  fill_stats(r, st);
  return r;
}

// n playouts of the same state with game ids first_game_id .. first_game_id+n-1
void caaa_ref_rollout_many(const void* state670, int64_t first_game_id, int n, CaaaRolloutStats* st,
                           double* results) {
  GameState tmp;
  for (int i = 0; i < n; i++) {
    memcpy(&tmp, state670, sizeof(tmp));
    if (first_game_id >= 0) {
      g_auto_game_id = false;
      g_game_id = (uint64_t)first_game_id + (uint64_t)i;
    }
    double r = random_play_until_terminal(&tmp);
    if (results) results[i] = r;
    if (st) fill_stats(r, st + i);
  }
}

void caaa_ref_get_state(void* out670) { memcpy(out670, &state, sizeof(state)); }

int caaa_ref_get_possible_actions(const void* state670, uint8_t* actions20) {
  GameState tmp;
  memcpy(&tmp, state670, sizeof(tmp));
  uint8_t n = 0;
  Actions acts = {0};
  get_possible_actions(&tmp, &n, &acts);
  memcpy(actions20, acts, ACTIONS_COUNT);
  return n;
}

void caaa_ref_apply_action(void* state670, uint8_t action) {
  GameState tmp;
  memcpy(&tmp, state670, sizeof(tmp));
  apply_action(&tmp, action);
  memcpy(state670, &tmp, sizeof(tmp));
}

int caaa_ref_is_terminal(const void* state670) {
  GameState tmp;
  memcpy(&tmp, state670, sizeof(tmp));
  return is_terminal_state(&tmp) ? 1 : 0;
}

// Load a state into the engine's globals the way random_play_until_terminal does
// (src/engine.cpp:4185-4199), for phase-by-phase differential tests.
void caaa_ref_load_state(const void* state670, int64_t game_id) {
  memcpy(&state, state670, sizeof(state));
  refresh_full_cache();
  answers_remaining = 100000;
  use_selected_action = false;
  max_loops = 1000;
  g_auto_game_id = false;
  g_game_id = (uint64_t)game_id;
  g_draws = 0;
  random_number_index = 0;
  memset(valid_moves, 0, sizeof(valid_moves));
  valid_moves_count = 0;
}

int caaa_ref_run_phase(int phase) {
  switch (phase) {
  case CAAA_PH_MOVE_FIGHTERS: return move_fighter_units();
  case CAAA_PH_MOVE_BOMBERS: return move_bomber_units();
  case CAAA_PH_STAGE_TRANSPORTS: return stage_transport_units();
  case CAAA_PH_MOVE_TANKS: return move_land_unit_type(TANKS);
  case CAAA_PH_MOVE_ARTILLERY: return move_land_unit_type(ARTILLERY);
  case CAAA_PH_MOVE_INFANTRY: return move_land_unit_type(INFANTRY);
  case CAAA_PH_MOVE_TRANSPORTS: return move_transport_units();
  case CAAA_PH_MOVE_SUBS: return move_subs();
  case CAAA_PH_MOVE_SHIPS: return move_destroyers_battleships();
  case CAAA_PH_SEA_BATTLES: return resolve_sea_battles();
  case CAAA_PH_UNLOAD_TRANSPORTS: return unload_transports();
  case CAAA_PH_LAND_BATTLES: return resolve_land_battles();
  case CAAA_PH_MOVE_AA_GUNS: return move_land_unit_type(AA_GUNS);
  case CAAA_PH_LAND_FIGHTERS: return land_fighter_units();
  case CAAA_PH_LAND_BOMBERS: return land_bomber_units();
  case CAAA_PH_BUY_UNITS: return buy_units();
  case CAAA_PH_CRASH_AIR: crash_air_units(); return 0;
  case CAAA_PH_RESET_UNITS: reset_units_fully(); return 0;
  case CAAA_PH_BUY_FACTORY: buy_factory(); return 0;
  case CAAA_PH_COLLECT_MONEY: collect_money(); return 0;
  case CAAA_PH_ROTATE_TURNS: rotate_turns(); return 0;
  default: return -1;
  }
}

double caaa_ref_evaluate(void) { return evaluate_state(&state); }
int caaa_ref_draws(void) { return (int)g_draws; }

void caaa_ref_get_cache(CaaaCacheDump* c) {
  memset(c, 0, sizeof(*c));
  for (int p = 0; p < 6; p++) {
    c->income_per_turn[p] = income_per_turn[p];
    c->total_factory_count[p] = total_factory_count[p];
    for (int l = 0; l < 5; l++) {
      c->factory_locations[p][l] = factory_locations[p][l];
      c->total_player_land_units[p][l] = total_player_land_units[p][l];
    }
    for (int s = 0; s < 3; s++) c->total_player_sea_units[p][s] = total_player_sea_units[p][s];
  }
  for (int s = 0; s < 3; s++) {
    c->transports_with_large_cargo_space[s] = transports_with_large_cargo_space[s];
    c->transports_with_small_cargo_space[s] = transports_with_small_cargo_space[s];
    c->allied_carriers[s] = allied_carriers[s];
    c->enemy_blockade_total[s] = enemy_blockade_total[s];
    c->enemy_destroyers_total[s] = enemy_destroyers_total[s];
  }
  for (int a = 0; a < 8; a++) c->enemy_units_count[a] = enemy_units_count[a];
  c->answers_remaining = answers_remaining;
  memcpy(c->current_player_land_unit_types, current_player_land_unit_types, 30);
  memcpy(c->current_player_sea_unit_types, current_player_sea_unit_types, 45);
  for (int i = 0; i < 25; i++) c->is_land_path_blocked[i] = is_land_path_blocked[i / 5][i % 5];
  for (int i = 0; i < 9; i++) {
    c->is_sea_path_blocked[i] = is_sea_path_blocked[i / 3][i % 3];
    c->is_sub_path_blocked[i] = is_sub_path_blocked[i / 3][i % 3];
  }
  for (int i = 0; i < 4; i++) c->enemies_0[i] = enemies_0[i];
  for (int i = 0; i < 5; i++) c->is_allied_0[i] = is_allied_0[i];
  c->enemies_count_0 = enemies_count_0;
  c->canal_state = canal_state;
  memcpy(c->valid_moves, valid_moves, CAAA_ACTIONS);
  c->valid_moves_count = valid_moves_count;
}

#define CP(dst, src) do { static_assert(sizeof(dst) == sizeof(src), #dst); memcpy(&(dst), &(src), sizeof(dst)); } while (0)

void caaa_ref_get_tables(CaaaTables* t) {
  memset(t, 0, sizeof(*t));
  for (int l = 0; l < 5; l++) {
    t->land_value[l] = LANDS[l].land_value;
    t->original_owner[l] = LANDS[l].original_owner_index;
  }
  for (int p = 0; p < 5; p++) {
    t->capital[p] = (uint8_t)PLAYERS[p].capital_territory_index;
    for (int q = 0; q < 5; q++) t->is_allied[p][q] = PLAYERS[p].is_allied[q];
  }
  for (int c = 0; c < 2; c++)
    for (int k = 0; k < 2; k++) t->canal_lands[c][k] = CANALS[c].lands[k];
  CP(t->land_dist, LAND_DIST);
  CP(t->air_dist, AIR_DIST);
  CP(t->sea_dist, SEA_DIST);
  CP(t->sea_path, SEA_PATH);
  CP(t->sea_path_alt, SEA_PATH_ALT);
  CP(t->land_path, LAND_PATH);
  CP(t->land_path_alt, LAND_PATH_ALT);
  CP(t->air_conn_count, AIR_CONN_COUNT);
  CP(t->air_conn, AIR_CONNECTIONS);
  CP(t->l2l_count, LAND_TO_LAND_COUNT);
  CP(t->l2l_conn, LAND_TO_LAND_CONN);
  CP(t->l2s_count, LAND_TO_SEA_COUNT);
  CP(t->l2s_conn, LAND_TO_SEA_CONN);
  CP(t->s2s_count, SEA_TO_SEA_COUNT);
  CP(t->s2s_conn, SEA_TO_SEA_CONN);
  CP(t->s2l_count, SEA_TO_LAND_COUNT);
  CP(t->s2l_conn, SEA_TO_LAND_CONN);
  CP(t->lands_within_2_count, LANDS_WITHIN_2_MOVES_COUNT);
  CP(t->lands_within_2, LANDS_WITHIN_2_MOVES);
  CP(t->load_within_2_count, LOAD_WITHIN_2_MOVES_COUNT);
  CP(t->load_within_2, LOAD_WITHIN_2_MOVES);
  CP(t->seas_within_1_count, SEAS_WITHIN_1_MOVE_COUNT);
  CP(t->seas_within_1, SEAS_WITHIN_1_MOVE);
  CP(t->seas_within_2_count, SEAS_WITHIN_2_MOVES_COUNT);
  CP(t->seas_within_2, SEAS_WITHIN_2_MOVES);
  CP(t->air_within_count, AIR_WITHIN_X_MOVES_COUNT);
  CP(t->air_within, AIR_WITHIN_X_MOVES);
  CP(t->air_to_land_within_count, AIR_TO_LAND_WITHIN_X_MOVES_COUNT);
  CP(t->air_to_land_within, AIR_TO_LAND_WITHIN_X_MOVES);
}

void caaa_ref_get_random_table(uint8_t* out65536) { memcpy(out65536, RANDOM_NUMBERS.data(), 65536); }

// mcts_search (src/mcts.cpp:65) on a copy of the state; returns the number of root children
// and their (action, visits, value).  The reference prints its tree every 50 000 iterations.
int caaa_ref_mcts(const void* state670, long iterations, uint8_t* actions, int32_t* visits,
                  double* values, uint8_t* best_action) {
  GameState tmp;
  memcpy(&tmp, state670, sizeof(tmp));
  MCTSNode* root = mcts_search(&tmp, iterations);
  int n = root->num_children;
  for (int i = 0; i < n && i < ACTIONS_COUNT; i++) {
    actions[i] = root->children[i]->action;
    visits[i] = root->children[i]->visits;
    values[i] = root->children[i]->value;
  }
  if (best_action) *best_action = n > 0 ? select_best_action(root) : 255;
  return n;
}

int caaa_ref_sizeof_state(void) { return (int)sizeof(GameState); }
int caaa_ref_sizeof_node(void) { return (int)sizeof(MCTSNode); }

}  // extern "C"
