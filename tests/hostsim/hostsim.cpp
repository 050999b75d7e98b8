// hostsim.cpp -- TEST INFRASTRUCTURE.  Compiles the *device* engine (caaa_b200/csrc/engine2.cuh,
// which is __host__ __device__) with g++ and drives it on the CPU, one game at a time, through
// the same C API as the oracle.  It exists so that the kernel's game logic can be checked
// bit-for-bit against the oracle in the CPU-only test tier (-m "not gpu"), where no GPU is
// available.  It is never part of the product library, and the product never falls back to it.
#include <cstring>

#include "../../caaa_b200/csrc/engine2.cuh"
#include "../../caaa_b200/csrc/host_tables.h"

using namespace caaa;
using namespace caaa::e2;

static DevTables g_tables;
static Game g_game;
static RunCtx g_cx;
static bool g_inited = false;
static uint64_t g_seed = 0;

static void set_ctx() {
  g_cx.T = &g_tables;
  g_cx.seed_lo = (uint32_t)g_seed;
  g_cx.seed_hi = (uint32_t)(g_seed >> 32);
}

static void fill_stats(Game* g, double score, CaaaRolloutStats* st) {
  st->result = score;
  st->turns = g->turns;
  st->decisions = 100000 - g->answers_remaining;
  st->draws = (int32_t)g->draws;
  st->final_player = g->player;
  st->state_hash = hash_live_state(g);
}

// random_play_until_terminal, reference src/engine.cpp:4184-4242, on one game
static double play(const void* state670, uint64_t game_id, CaaaRolloutStats* st) {
  Game* g = &g_game;
  import_state(g, (const uint8_t*)state670);
  begin_playout(g, g_cx, game_id);
  double score = evaluate_state(g, g_cx);
  while (score > 0.01 && score < 0.99 && g->turns < 1000) {
    for (int ph = 0; ph < CAAA_PH_COUNT; ph++) run_phase<false>(g, g_cx, ph);
    g->turns++;
    score = evaluate_state(g, g_cx);
  }
  if (g->player % 2 == 1) score = 1 - score;
  if (st) fill_stats(g, score, st);
  return score;
}

static void run_until_decision(Game* g) {
  for (;;) {
    bool stopped = false;
    for (int ph = 0; ph <= CAAA_PH_BUY_UNITS; ph++)
      if (run_phase<true>(g, g_cx, ph)) { stopped = true; break; }
    if (stopped) break;
    for (int ph = CAAA_PH_CRASH_AIR; ph < CAAA_PH_COUNT; ph++) run_phase<true>(g, g_cx, ph);
  }
}

extern "C" {
void caaa_hostsim_set_map(const CaaaMap* m) {  // other boards (same dimensions): rebuilds the tables
  build_dev_tables(*m, &g_tables);
  std::memset(&g_game, 0, sizeof(g_game));
  set_ctx();
  g_inited = true;
}
void caaa_hostsim_init(void) {
  if (g_inited) return;
  CaaaMap m;
  default_map(&m);
  build_dev_tables(m, &g_tables);
  std::memset(&g_game, 0, sizeof(g_game));
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
void caaa_hostsim_get_state(void* out) { export_state(&g_game, (uint8_t*)out); }
int caaa_hostsim_stuck(void) {
  int s = (g_game.status & CAAA_ERR_SEA_ROUND_CAP) != 0;
  g_game.status = 0;
  return s;
}
int caaa_hostsim_get_possible_actions(const void* s, uint8_t* actions20) {
  Game* g = &g_game;
  import_state(g, (const uint8_t*)s);
  begin_common(g, g_cx);
  g->answers_remaining = 0;
  run_until_decision(g);
  std::memset(actions20, 0, CAAA_ACTIONS);
  std::memcpy(actions20, g->vm, g->nvm < CAAA_ACTIONS ? g->nvm : CAAA_ACTIONS);
  return g->nvm;
}
void caaa_hostsim_apply_action(void* s, uint8_t action) {
  Game* g = &g_game;
  import_state(g, (const uint8_t*)s);
  begin_common(g, g_cx);
  g->answers_remaining = 1;
  g->selected_action = action;
  g->use_selected = 1;
  run_until_decision(g);
  export_state(g, (uint8_t*)s);
}
int caaa_hostsim_is_terminal(const void* s) {
  // the reference scores a node with the unit tables left behind by the previous call (src/engine.cpp:4314)
  // but the node's own player index and money (4305-4311)
  Game tmp = g_game;
  const CaaaGameState* in = (const CaaaGameState*)s;
  const int old_cur = tmp.cur;
  uint8_t rel_money[NP];
  for (int p = 0; p < NP; p++) rel_money[p] = in->money[p];
  tmp.player = in->player_index;
  // keep the unit tables in the frame they were left in: relative player p stays relative player p
  Game t2 = tmp;
  t2.cur = (uint8_t)(in->player_index % NP);
  for (int p = 0; p < NP; p++) {
    std::memcpy(t2.ut[(t2.cur + p) % NP], tmp.ut[(old_cur + p) % NP], UT_BYTES);
    t2.money[(t2.cur + p) % NP] = rel_money[p];
  }
  double sc = evaluate_state(&t2, g_cx);
  return sc > 0.99 || sc < 0.01;
}
void caaa_hostsim_load_state(const void* s, int64_t game_id) {
  import_state(&g_game, (const uint8_t*)s);
  begin_playout(&g_game, g_cx, (uint64_t)game_id);
}
int caaa_hostsim_run_phase(int ph) { return run_phase<false>(&g_game, g_cx, ph) ? 1 : 0; }
double caaa_hostsim_evaluate(void) { return evaluate_state(&g_game, g_cx); }
int caaa_hostsim_draws(void) { return (int)g_game.draws; }
void caaa_hostsim_get_cache(CaaaCacheDump* d) {
  std::memset(d, 0, sizeof(*d));
  export_cache(&g_game, &g_tables, d);
}
void caaa_hostsim_get_tables(CaaaTables* t) { *t = g_tables.t; }
}
extern "C" void caaa_hostsim_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  philox4x32_10(ctr[0], ctr[1], ctr[2], ctr[3], key[0], key[1], out);
}
extern "C" uint8_t caaa_hostsim_stream_byte(uint64_t seed, uint64_t game_id, uint32_t draw_index) {
  Game g;
  std::memset(&g, 0, sizeof(g));
  g.game_lo = (uint32_t)game_id;
  g.game_hi = (uint32_t)(game_id >> 32);
  RunCtx cx{&g_tables, (uint32_t)seed, (uint32_t)(seed >> 32)};
  g.draws = draw_index & ~15u;
  uint8_t b = 0;
  for (uint32_t d = g.draws; d <= draw_index; d++) b = next_byte(&g, cx);
  return b;
}
// playout in which every phase yields after `budget` loop iterations and is resumed from its saved cursor until it
// is done -- the way the phase-regrouped kernel time-slices long phases
extern "C" void caaa_hostsim_rollout_many_sliced(const void* s, int64_t first, int n, int budget, CaaaRolloutStats* st) {
  for (int i = 0; i < n; i++) {
    Game* g = &g_game;
    import_state(g, (const uint8_t*)s);
    begin_playout(g, g_cx, (uint64_t)first + (uint64_t)i);
    double score = evaluate_state(g, g_cx);
    while (score > 0.01 && score < 0.99 && g->turns < 1000) {
      for (int ph = 0; ph < CAAA_PH_COUNT; ph++)
        while (run_phase<false>(g, g_cx, ph, 0xffffffffu, budget) == PH_YIELDED) {
        }
      g->turns++;
      score = evaluate_state(g, g_cx);
    }
    if (g->player % 2 == 1) score = 1 - score;
    fill_stats(g, score, st + i);
  }
}
// like rollout_many_sliced, but tanks / artillery / infantry (phases 3..5, move_land_units) and transports / subs / ships
// (phases 6..8, move_sea_units) each move in one merged visit, the way the phase-regrouped kernel runs them
extern "C" void caaa_hostsim_rollout_many_merged(const void* s, int64_t first, int n, int budget, CaaaRolloutStats* st) {
  for (int i = 0; i < n; i++) {
    Game* g = &g_game;
    import_state(g, (const uint8_t*)s);
    begin_playout(g, g_cx, (uint64_t)first + (uint64_t)i);
    double score = evaluate_state(g, g_cx);
    while (score > 0.01 && score < 0.99 && g->turns < 1000) {
      for (int ph = 0; ph < CAAA_PH_COUNT; ph++) {
        if (ph == CAAA_PH_MOVE_TANKS) {
          while (move_land_units<false>(g, g_cx, 0xffffffffu, CAAA_PH_MOVE_TANKS, CAAA_PH_MOVE_INFANTRY, budget) == PH_YIELDED) {
          }
          ph = CAAA_PH_MOVE_INFANTRY;
          continue;
        }
        if (ph == CAAA_PH_MOVE_TRANSPORTS) {
          while (move_sea_units<false>(g, g_cx, 0xffffffffu, budget) == PH_YIELDED) {
          }
          ph = CAAA_PH_MOVE_SHIPS;
          continue;
        }
        while (run_phase<false>(g, g_cx, ph, 0xffffffffu, budget) == PH_YIELDED) {
        }
      }
      g->turns++;
      score = evaluate_state(g, g_cx);
    }
    if (g->player % 2 == 1) score = 1 - score;
    fill_stats(g, score, st + i);
  }
}
extern "C" void caaa_hostsim_rollout_many_routed(const void* s, int64_t first, int n, CaaaRolloutStats* st) {
  caaa_hostsim_rollout_many_sliced(s, first, n, 1, st);
}
// iteration statistics of the phase loops (tools only): runs `n` playouts and, for every phase call that had work,
// adds one sample (phase, units waiting in the phase's cells at entry, loop iterations) to the caller's arrays
extern "C" int caaa_hostsim_phase_stats(const void* s, int64_t first, int n, int32_t* out, int cap) {
  int k = 0;
  for (int i = 0; i < n; i++) {
    Game* g = &g_game;
    import_state(g, (const uint8_t*)s);
    begin_playout(g, g_cx, (uint64_t)first + (uint64_t)i);
    double score = evaluate_state(g, g_cx);
    while (score > 0.01 && score < 0.99 && g->turns < 1000) {
      for (int ph = 0; ph < CAAA_PH_COUNT; ph++) {
        const bool has = ph <= CAAA_PH_BUY_UNITS && (ph == CAAA_PH_BUY_UNITS || (g->work >> ph) & 1u);
        int units = 0;
        if (has && ph < CAAA_PH_BUY_UNITS) {
          static const int sec[15][2] = {{SC_FIGHTERS, SC_FIGHTERS_N}, {SC_BOMBERS, SC_BOMBERS_N}, {SC_STAGE, SC_STAGE_N}, {SC_TANKS, SC_TANKS_N},
                                         {SC_ARTILLERY, SC_ARTILLERY_N}, {SC_INFANTRY, SC_INFANTRY_N}, {SC_TRANSPORTS, SC_TRANSPORTS_N}, {-1, 0},
                                         {SC_SHIPS, SC_SHIPS_N}, {-1, 0}, {SC_UNLOAD, SC_UNLOAD_N}, {-1, 0}, {SC_AA_GUNS, SC_AA_GUNS_N},
                                         {SC_LAND_FIGHTERS, SC_LAND_FIGHTERS_N}, {SC_LAND_BOMBERS, SC_LAND_BOMBERS_N}};
          if (sec[ph][0] >= 0)
            for (int c = 0; c < sec[ph][1]; c++) units += ((const uint8_t*)g)[g_tables.scan[sec[ph][0] + c] & 1023u];
        }
        if (ph == CAAA_PH_BUY_UNITS) units = g->money[g->cur];
        int iters = 1;
        while (run_phase<false>(g, g_cx, ph, 0xffffffffu, 1) == PH_YIELDED) iters++;
        if (has && k + 3 <= cap) {
          out[k++] = ph;
          out[k++] = units;
          out[k++] = iters;
        }
      }
      g->turns++;
      score = evaluate_state(g, g_cx);
    }
  }
  return k / 3;
}
