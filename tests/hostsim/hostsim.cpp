// hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the *device* engine (caaa_b200/csrc/engine.cuh,
// which is __host__ __device__) with g++ and drives it on the CPU, one slot at a time, through
// the same C API as the oracle.  It exists so that the kernel's game logic can be checked
// bit-for-bit against the oracle in the CPU-only test tier (-m "not gpu"), where no GPU is
// available.  It is never part of the product library, and the product never falls back to it.
#include <cstring>

#include "../../caaa_b200/csrc/engine.cuh"
#include "../../caaa_b200/csrc/host_tables.h"

using namespace caaa;

static DevTables g_tables;
static GameSlot g_slot;
static RunCtx g_cx;
static bool g_inited = false;
static uint64_t g_seed = 0;
static int g_max_loops = 0;

static void set_ctx() {
  g_cx.T = &g_tables;
  g_cx.seed_lo = (uint32_t)g_seed;
  g_cx.seed_hi = (uint32_t)(g_seed >> 32);
}

// random_play_until_terminal, reference src/engine.cpp:4184-4242, on one slot
static double play(const void* state670, uint64_t game_id, CaaaRolloutStats* st) {
  GameSlot* g = &g_slot;
  std::memcpy(&g->st, state670, CAAA_STATE_BYTES);
  begin_playout(g, g_cx, game_id);
  double score = evaluate_state(g, g_cx);
  while (score > 0.01 && score < 0.99 && g->turns < 1000) {
    for (int ph = 0; ph < CAAA_PH_COUNT; ph++) run_phase<false>(g, g_cx, ph);
    g->turns++;
    score = evaluate_state(g, g_cx);
  }
  if (g->st.player_index % 2 == 1) score = 1 - score;
  if (st) {
    st->result = score;
    st->turns = g->turns;
    st->decisions = 100000 - g->answers_remaining;
    st->draws = (int32_t)g->draws;
    st->final_player = g->st.player_index;
    st->state_hash = hash_live_state(g);
  }
  return score;
}

static void run_until_decision(GameSlot* g) {
  for (;;) {
    bool stopped = false;
    for (int ph = 0; ph <= CAAA_PH_BUY_UNITS; ph++)
      if (run_phase<true>(g, g_cx, ph)) { stopped = true; break; }
    if (stopped) break;
    for (int ph = CAAA_PH_CRASH_AIR; ph < CAAA_PH_COUNT; ph++) run_phase<true>(g, g_cx, ph);
  }
}

extern "C" {
void caaa_hostsim_init(void) {
  if (g_inited) return;
  build_dev_tables(&g_tables);
  std::memset(&g_slot, 0, sizeof(g_slot));
  set_ctx();
  g_inited = true;
}
void caaa_hostsim_set_stream(int, uint64_t seed) { g_seed = seed; set_ctx(); }
void caaa_hostsim_set_reset_valid_moves(int) {}
static int64_t g_fixed_game_id = -1;
static uint64_t g_counter = 0;
void caaa_hostsim_set_game_id(int64_t gid) { g_fixed_game_id = gid; }
double caaa_hostsim_rollout(const void* s, CaaaRolloutStats* st) {
  uint64_t gid = g_fixed_game_id >= 0 ? (uint64_t)g_fixed_game_id : g_counter;
  g_counter++;
  return play(s, gid, st);
}
void caaa_hostsim_rollout_many(const void* s, int64_t first, int n, CaaaRolloutStats* st, double* results) {
  for (int i = 0; i < n; i++) {
    double r = play(s, (uint64_t)first + (uint64_t)i, st ? st + i : nullptr);
    if (results) results[i] = r;
  }
}
void caaa_hostsim_get_state(void* out) { std::memcpy(out, &g_slot.st, CAAA_STATE_BYTES); }
int caaa_hostsim_stuck(void) {
  int s = (g_slot.status & CAAA_ERR_SEA_ROUND_CAP) != 0;
  g_slot.status = 0;
  return s;
}
int caaa_hostsim_get_possible_actions(const void* s, uint8_t* actions20) {
  GameSlot* g = &g_slot;
  std::memcpy(&g->st, s, CAAA_STATE_BYTES);
  refresh_full_cache(g, g_cx);
  g->unlucky = 0;
  g->answers_remaining = 0;
  g->status = 0;
  run_until_decision(g);
  std::memset(actions20, 0, CAAA_ACTIONS);
  std::memcpy(actions20, g->vm, g->nvm < CAAA_ACTIONS ? g->nvm : CAAA_ACTIONS);
  return g->nvm;
}
void caaa_hostsim_apply_action(void* s, uint8_t action) {
  GameSlot* g = &g_slot;
  std::memcpy(&g->st, s, CAAA_STATE_BYTES);
  refresh_full_cache(g, g_cx);
  g->unlucky = 0;
  g->answers_remaining = 1;
  g->selected_action = action;
  g->use_selected = 1;
  g->status = 0;
  run_until_decision(g);
  std::memcpy(s, &g->st, CAAA_STATE_BYTES);
}
int caaa_hostsim_is_terminal(const void* s) {
  GameSlot tmp = g_slot;
  const CaaaGameState* in = (const CaaaGameState*)s;
  tmp.st.player_index = in->player_index;
  std::memcpy(tmp.st.money, in->money, NP);
  double sc = evaluate_state(&tmp, g_cx);
  return sc > 0.99 || sc < 0.01;
}
void caaa_hostsim_load_state(const void* s, int64_t game_id) {
  std::memcpy(&g_slot.st, s, CAAA_STATE_BYTES);
  begin_playout(&g_slot, g_cx, (uint64_t)game_id);
}
int caaa_hostsim_run_phase(int ph) { return run_phase<false>(&g_slot, g_cx, ph) ? 1 : 0; }
double caaa_hostsim_evaluate(void) { return evaluate_state(&g_slot, g_cx); }
int caaa_hostsim_draws(void) { return (int)g_slot.draws; }
void caaa_hostsim_get_cache(CaaaCacheDump* d) {
  GameSlot* g = &g_slot;
  std::memset(d, 0, sizeof(*d));
  const int pi = pidx(g);
  for (int p = 0; p < 6; p++) {
    d->income_per_turn[p] = g->income[p];
    d->total_factory_count[p] = g->nfact[p];
    for (int l = 0; l < NL; l++) {
      d->factory_locations[p][l] = g->factloc[p][l];
      d->total_player_land_units[p][l] = g->tot_land[p][l];
    }
    for (int s = 0; s < NS; s++) d->total_player_sea_units[p][s] = g->tot_sea[p][s];
  }
  for (int s = 0; s < NS; s++) {
    d->transports_with_large_cargo_space[s] = g->large_cargo[s];
    d->transports_with_small_cargo_space[s] = g->small_cargo[s];
    d->allied_carriers[s] = g->carriers[s];
    d->enemy_blockade_total[s] = g->blockade[s];
    d->enemy_destroyers_total[s] = g->edestroyers[s];
  }
  for (int a = 0; a < NA; a++) d->enemy_units_count[a] = g->enemy_count[a];
  d->answers_remaining = g->answers_remaining;
  std::memcpy(d->current_player_land_unit_types, g->cur_land, 30);
  std::memcpy(d->current_player_sea_unit_types, g->cur_sea, 45);
  std::memcpy(d->is_land_path_blocked, g->land_blocked, 25);
  std::memcpy(d->is_sea_path_blocked, g->sea_blocked, 9);
  std::memcpy(d->is_sub_path_blocked, g->sub_blocked, 9);
  // the team lists are tabulated per player-to-move in this implementation
  for (int e = 0; e < g_tables.n_enemies[pi]; e++) d->enemies_0[e] = g_tables.enemies[pi][e];
  for (int p = 0; p < NP; p++) d->is_allied_0[p] = (g_tables.allied_mask[pi] >> p) & 1;
  d->enemies_count_0 = g_tables.n_enemies[pi];
  d->canal_state = g->canal_state;
  std::memcpy(d->valid_moves, g->vm, CAAA_ACTIONS);
  d->valid_moves_count = g->nvm;
}
void caaa_hostsim_get_tables(CaaaTables* t) { *t = g_tables.t; }
}
extern "C" void caaa_hostsim_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
extern "C" uint8_t caaa_hostsim_stream_byte(uint64_t seed, uint64_t game_id, uint32_t draw_index) {
  GameSlot g;
  std::memset(&g, 0, sizeof(g));
  g.game_lo = (uint32_t)game_id;
  g.game_hi = (uint32_t)(game_id >> 32);
  RunCtx cx{&g_tables, (uint32_t)seed, (uint32_t)(seed >> 32)};
  g.draws = draw_index & ~15u;
  uint8_t b = 0;
  for (uint32_t d = g.draws; d <= draw_index; d++) b = next_byte(&g, cx);
  return b;
}
// playout driven the way the queue scheduler drives it: only phases with work are executed
extern "C" void caaa_hostsim_rollout_many_routed(const void* s, int64_t first, int n, CaaaRolloutStats* st) {
  for (int i = 0; i < n; i++) {
    GameSlot* g = &g_slot;
    std::memcpy(&g->st, s, CAAA_STATE_BYTES);
    begin_playout(g, g_cx, (uint64_t)first + (uint64_t)i);
    double score = evaluate_state(g, g_cx);
    while (score > 0.01 && score < 0.99 && g->turns < 1000) {
      int ph = next_phase_with_work(g, 0);
      while (ph < CAAA_PH_BUY_UNITS) {
        run_phase<false>(g, g_cx, ph);
        ph = next_phase_with_work(g, ph + 1);
      }
      for (ph = CAAA_PH_BUY_UNITS; ph < CAAA_PH_COUNT; ph++) run_phase<false>(g, g_cx, ph);
      g->turns++;
      score = evaluate_state(g, g_cx);
    }
    if (g->st.player_index % 2 == 1) score = 1 - score;
    st[i].result = score;
    st[i].turns = g->turns;
    st[i].decisions = 100000 - g->answers_remaining;
    st[i].draws = (int32_t)g->draws;
    st[i].final_player = g->st.player_index;
    st[i].state_hash = hash_live_state(g);
  }
}
