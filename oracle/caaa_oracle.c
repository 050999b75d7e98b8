/* caaa_oracle.c -- TEST INFRASTRUCTURE.  Plain-C CPU restatement of the reference's rollout
 * path: the 21-phase player turn, its derived caches, scoring and the stop-at-decision mode
 * (reference src/engine.cpp:642-814, 1291-4337).  Every function cites the reference lines it
 * follows.  Reference quirks are reproduced on purpose; the canonical
 * meaning of the layout-dependent out-of-bounds accesses is patch set P1.
 *
 * Pinned against the patched reference build oracle/_ref (see caaa_oracle.h).  Never linked
 * into the product library.
 */
#include "caaa_oracle.h"
#include "philox_ref.h"
#include <string.h>

/* ---- unit tables: reference src/units.cpp, include/units/*.h ---- */
enum { FIGHTERS = 0, BOMBERS_LAND_AIR = 1, INFANTRY = 2, ARTILLERY = 3, TANKS = 4, AA_GUNS = 5 };
enum { TRANS_EMPTY = 1, TRANS_1I, TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, SUBMARINES,
       DESTROYERS, CARRIERS, CRUISERS, BATTLESHIPS, BS_DAMAGED, BOMBERS_SEA };
#define NL 5
#define NS 3
#define NA 8
#define NP 5
#define NLT 6
#define NST 15

static const uint8_t UNIT_WEIGHTS[NLT] = {255, 255, 2, 3, 3, 6};
static const uint8_t ATTACK_LAND[NLT] = {3, 4, 1, 2, 3, 0};
static const uint8_t MAX_MOVE_LAND[NLT] = {4, 6, 1, 1, 2, 1};
static const uint8_t STATES_LAND[NLT] = {5, 7, 2, 2, 3, 2};
static const uint8_t COST_LAND[NLT] = {10, 14, 3, 4, 5, 5};
static const uint8_t OFF_LAND[NLT] = {4, 9, 16, 18, 20, 23}; /* byte offsets in LandState */
static const uint8_t ATTACK_SEA[NST] = {3, 0, 0, 0, 0, 0, 0, 0, 2, 2, 1, 3, 4, 4, 4};
static const uint8_t DEFENSE_SEA[NST] = {4, 0, 0, 0, 0, 0, 0, 0, 1, 2, 2, 3, 4, 4, 1};
static const uint8_t MAX_MOVE_SEA[NST] = {4, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 2, 6};
static const uint8_t STATES_SEA[NST] = {5, 4, 5, 5, 5, 4, 4, 4, 3, 2, 2, 3, 3, 3, 5};
static const uint8_t COST_SEA[NST] = {10, 8, 11, 12, 13, 14, 15, 16, 8, 8, 14, 10, 22, 22, 14};
static const uint8_t OFF_SEA[NST] = {0, 5, 9, 14, 19, 24, 28, 32, 36, 39, 41, 43, 46, 49, 52};
static const uint8_t BUY_SEA[7] = {FIGHTERS, TRANS_EMPTY, SUBMARINES, DESTROYERS, CARRIERS, CRUISERS, BATTLESHIPS};
static const uint8_t STATES_STAGING[NST] = {0, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t STATES_UNLOADING[NST] = {0, 0, 1, 1, 1, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0};
static const uint8_t LOAD_INF[NST] = {255, TRANS_1I, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255};
static const uint8_t LOAD_ART[NST] = {255, TRANS_1A, TRANS_1I_1A, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255};
static const uint8_t LOAD_TANK[NST] = {255, TRANS_1T, TRANS_1I_1T, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255, 255};
static const uint8_t UNLOAD_CARGO1[NST] = {255, 255, INFANTRY, ARTILLERY, TANKS, INFANTRY, ARTILLERY, TANKS, 255, 255, 255, 255, 255, 255, 255};
static const uint8_t UNLOAD_CARGO2[NST] = {255, 255, 255, 255, 255, INFANTRY, INFANTRY, INFANTRY, 255, 255, 255, 255, 255, 255, 255};
static const uint8_t UNMOVED_SEA[NST] = {4, 1, 1, 1, 1, 1, 1, 1, 2, 1, 1, 2, 2, 2, 255};
static const uint8_t DONE_MOVING_SEA[NST] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 0};
/* casualty orders, engine.cpp:81-97 */
static const uint8_t ORDER_LAND_DEF[6] = {AA_GUNS, BOMBERS_LAND_AIR, INFANTRY, ARTILLERY, TANKS, FIGHTERS};
static const uint8_t ORDER_LAND_ATT1[3] = {INFANTRY, ARTILLERY, TANKS};
static const uint8_t ORDER_LAND_ATT2[2] = {FIGHTERS, BOMBERS_LAND_AIR};
static const uint8_t ORDER_SEA_DEF[13] = {SUBMARINES, DESTROYERS, CARRIERS, CRUISERS, FIGHTERS, BS_DAMAGED, TRANS_EMPTY,
                                          TRANS_1I, TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T};
static const uint8_t ORDER_SEA_ATT1[2] = {SUBMARINES, DESTROYERS};
static const uint8_t ORDER_SEA_ATT2[2] = {FIGHTERS, BOMBERS_SEA};
static const uint8_t ORDER_SEA_ATT3[8] = {BS_DAMAGED, TRANS_EMPTY, TRANS_1I, TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T};
static const uint8_t BLOCKADE_TYPES[5] = {DESTROYERS, CARRIERS, CRUISERS, BATTLESHIPS, BS_DAMAGED};

/* ---- the engine's file-scope globals (engine.cpp:98-163), gathered in one context ---- */
typedef struct {
  CaaaGameState st;
  uint8_t cur_land[NL][NLT];  /* current_player_land_unit_types */
  uint8_t cur_sea[NS][NST];   /* current_player_sea_unit_types */
  int income[6], nfact[6], factloc[6][NL];
  int tot_land[6][NL], tot_sea[6][NS];
  uint8_t blockade[NS], edestroyers[NS];
  uint8_t land_blocked[25], sea_blocked[9], sub_blocked[9];
  int large_cargo[NS], small_cargo[NS];
  uint8_t enemies[4], allied[NP], n_enemies, canal_state;
  int carriers[NS], enemy_count[NA];
  uint8_t can_f_here[NA], can_f_1[NA], can_b_here[NA], can_b_1[NA], can_b_2[NA];
  int answers_remaining, use_selected, max_loops;
  uint8_t selected_action, unlucky;
  uint8_t vm[CAAA_ACTIONS]; /* valid_moves */
  uint8_t nvm;              /* valid_moves_count */
  /* random stream (P2) */
  uint64_t seed, game_id;
  uint32_t draws;
  int auto_game_id;
  uint64_t playout_counter;
  int reset_vm;
  int stuck; /* guard: set when a sea battle exceeds CAAA_SEA_ROUND_CAP rounds (the reference would spin forever) */
} Ctx;
#define CAAA_SEA_ROUND_CAP 10000

static Ctx G;
static CaaaTables T;
static int g_inited = 0;

/* pointer views that mirror the reference's pointer tables (engine.cpp:521-568) */
static inline uint8_t* land_ptr(Ctx* c, int land, int type) { return (uint8_t*)&c->st.land_state[land] + OFF_LAND[type]; }
static inline uint8_t* sea_ptr(Ctx* c, int sea, int type) { return (uint8_t*)&c->st.units_sea[sea] + OFF_SEA[type]; }
static inline uint8_t* air_ptr(Ctx* c, int air, int type) { /* air_units_state: FIGHTERS / BOMBERS_LAND_AIR only */
  if (air < NL) return land_ptr(c, air, type);
  return sea_ptr(c, air - NL, type == FIGHTERS ? FIGHTERS : BOMBERS_SEA);
}
static inline uint8_t* tplut(Ctx* c, int player, int land) { /* total_player_land_unit_types */
  return player == 0 ? c->cur_land[land] : c->st.other_land_units[player - 1][land];
}
static inline uint8_t* tpsut(Ctx* c, int player, int sea) { /* total_player_sea_unit_types */
  return player == 0 ? c->cur_sea[sea] : c->st.other_sea_units[player - 1][sea];
}
#define OWNER(c, land) ((c)->st.land_state[land].owner_idx)
#define FMAX(c, land) ((c)->st.land_state[land].factory_max)
#define FHP(c, land) ((c)->st.land_state[land].factory_hp)
#define BOMBARD(c, land) ((c)->st.land_state[land].bombard_max)
#define FLAG(c, air) ((c)->st.flagged_for_combat[air])

/* skipped_moves as a flat 64-flag array, P1(ii) */
static inline int skip_get(Ctx* c, unsigned a, unsigned b) {
  unsigned i = a * 8u + b;
  return i < 64u ? (((uint8_t*)c->st.skipped_moves)[i] & 1) : 0;
}
static inline void skip_set(Ctx* c, unsigned a, unsigned b) {
  unsigned i = a * 8u + b;
  if (i < 64u) ((uint8_t*)c->st.skipped_moves)[i] |= 1;
}
static void clear_move_history(Ctx* c) { memset(c->st.skipped_moves, 0, 64); } /* engine.cpp:1330 */

static uint8_t next_byte(Ctx* c) { return caaa_philox_byte_ref(c->seed, c->game_id, c->draws++); }

/* ---- cache refresh, engine.cpp:642-814 ---- */
static void refresh_economy(Ctx* c) { /* 661-666 */
  for (int p = 0; p < NP; p++) c->income[p] = c->nfact[p] = 0;
}
static void refresh_land_armies(Ctx* c) { /* 667-694 */
  memset(c->cur_land, 0, sizeof(c->cur_land));
  memset(c->tot_land, 0, sizeof(c->tot_land));
  for (int l = 0; l < NL; l++) {
    int owner = OWNER(c, l);
    if (FMAX(c, l) > 0) c->factloc[owner][c->nfact[owner]++] = l;
    c->income[owner] += T.land_value[l];
    for (int u = 0; u < NLT; u++) {
      const uint8_t* s = land_ptr(c, l, u);
      for (int k = 0; k < STATES_LAND[u]; k++) c->cur_land[l][u] = (uint8_t)(c->cur_land[l][u] + s[k]);
    }
    for (int p = 0; p < NP; p++)
      for (int u = 0; u < NLT; u++) c->tot_land[p][l] += tplut(c, p, l)[u];
  }
}
static void refresh_sea_navies(Ctx* c) { /* 695-715 */
  memset(c->cur_sea, 0, sizeof(c->cur_sea));
  memset(c->tot_sea, 0, sizeof(c->tot_sea));
  for (int s = 0; s < NS; s++) {
    for (int u = 0; u < NST; u++) {
      const uint8_t* p = sea_ptr(c, s, u);
      for (int k = 0; k < STATES_SEA[u]; k++) c->cur_sea[s][u] = (uint8_t)(c->cur_sea[s][u] + p[k]);
    }
    for (int p = 0; p < NP; p++)
      for (int u = 0; u < NST - 1; u++) c->tot_sea[p][s] += tpsut(c, p, s)[u];
  }
}
static void refresh_allies(Ctx* c) { /* 716-725 */
  c->n_enemies = 0;
  int pi = c->st.player_index;
  for (int p = 0; p < NP; p++) {
    c->allied[p] = T.is_allied[pi][(pi + p) % NP];
    if (!c->allied[p]) c->enemies[c->n_enemies++] = (uint8_t)p;
  }
}
static void refresh_canals(Ctx* c) { /* 726-751; the per-canal table copies are read through canal_state here */
  c->canal_state = 0;
  for (int k = 0; k < CAAA_CANALS; k++)
    if (c->allied[OWNER(c, T.canal_lands[k][0])] && c->allied[OWNER(c, T.canal_lands[k][1])])
      c->canal_state = (uint8_t)(c->canal_state + (1u << k));
}
static void refresh_enemy_armies(Ctx* c) { /* 752-760 */
  memset(c->enemy_count, 0, sizeof(c->enemy_count));
  for (int l = 0; l < NL; l++)
    for (int e = 0; e < c->n_enemies; e++) c->enemy_count[l] += c->tot_land[c->enemies[e]][l];
}
static void refresh_fleets(Ctx* c) { /* 761-784 */
  memset(c->edestroyers, 0, NS);
  memset(c->blockade, 0, NS);
  for (int s = 0; s < NS; s++) {
    const uint8_t* mine = c->cur_sea[s];
    c->carriers[s] = mine[CARRIERS];
    for (int p = 1; p < NP; p++) {
      const uint8_t* su = tpsut(c, p, s);
      if (c->allied[p]) {
        c->carriers[s] += su[CARRIERS];
      } else {
        c->enemy_count[s + NL] += c->tot_sea[p][s];
        c->edestroyers[s] = (uint8_t)(c->edestroyers[s] + su[DESTROYERS]);
        c->blockade[s] = (uint8_t)(c->blockade[s] + su[DESTROYERS] + su[CARRIERS] + su[CRUISERS] +
                                   su[BATTLESHIPS] + su[BS_DAMAGED]);
      }
    }
    c->large_cargo[s] = mine[TRANS_EMPTY] + mine[TRANS_1I];
    c->small_cargo[s] = mine[TRANS_EMPTY] + mine[TRANS_1I] + mine[TRANS_1A] + mine[TRANS_1T];
  }
}
static void refresh_land_path_blocked(Ctx* c) { /* 785-798; entries outside lands_within_2 are never written */
  for (int s = 0; s < NL; s++)
    for (int i = 0; i < T.lands_within_2_count[s]; i++) {
      int d = T.lands_within_2[s][i];
      int n1 = T.land_path[s][d], n2 = T.land_path_alt[s][d];
      c->land_blocked[s * NL + d] = (c->enemy_count[n1] > 0 || FMAX(c, n1) > 0) &&
                                    (c->enemy_count[n2] > 0 || FMAX(c, n2) > 0);
    }
}
static void refresh_sea_path_blocked(Ctx* c) { /* 799-814 */
  int cs = c->canal_state;
  for (int s = 0; s < NS; s++)
    for (int i = 0; i < T.seas_within_2_count[cs][s]; i++) {
      int d = T.seas_within_2[cs][s][i];
      int n1 = T.sea_path[cs][s][d], n2 = T.sea_path_alt[cs][s][d];
      c->sea_blocked[s * NS + d] = c->blockade[n1] > 0 && c->blockade[n2] > 0;
      c->sub_blocked[s * NS + d] = c->edestroyers[n1] > 0 && c->edestroyers[n2] > 0;
    }
}
static void refresh_eot_cache(Ctx* c) { /* 653-660 */
  refresh_allies(c);
  refresh_canals(c);
  refresh_enemy_armies(c);
  refresh_fleets(c);
  refresh_land_path_blocked(c);
  refresh_sea_path_blocked(c);
}
static void refresh_full_cache(Ctx* c) { /* 642-652 */
  refresh_economy(c);
  refresh_land_armies(c);
  refresh_sea_navies(c);
  refresh_eot_cache(c);
}

/* ---- decisions and dice ---- */
static uint8_t ai_input(Ctx* c) { /* getAIInput 1291-1297 */
  c->answers_remaining--;
  if (c->use_selected) return c->selected_action;
  return c->vm[next_byte(c) % c->nvm];
}
static int is_unlucky_team(Ctx* c) { return T.is_allied[c->st.player_index][c->unlucky]; }
static uint8_t attacker_hits(Ctx* c, int dmg) { /* 2552-2565 */
  if (c->answers_remaining < 2) {
    if (is_unlucky_team(c)) return (uint8_t)(dmg / 6);
    return (uint8_t)(dmg / 6 + (1 < dmg % 6 ? 1 : 0));
  }
  return (uint8_t)(dmg / 6 + (next_byte(c) % 6 < dmg % 6 ? 1 : 0));
}
static uint8_t defender_hits(Ctx* c, int dmg) { /* 2567-2580 */
  if (c->answers_remaining < 2) {
    if (is_unlucky_team(c)) return (uint8_t)(dmg / 6 + (1 < dmg % 6 ? 1 : 0));
    return (uint8_t)(dmg / 6);
  }
  return (uint8_t)(dmg / 6 + (next_byte(c) % 6 < dmg % 6 ? 1 : 0));
}

static void apply_skip(Ctx* c, unsigned src, unsigned dst) { /* 1322-1328 */
  for (unsigned i = 0; i < NA; i++)
    if (skip_get(c, dst, i)) skip_set(c, src, i);
}
static void update_move_history_4air(Ctx* c, uint8_t src, uint8_t dst) { /* 1300-1320: starts one past the list */
  for (;;) {
    uint8_t va = c->nvm < CAAA_ACTIONS ? c->vm[c->nvm] : 0;
    if (va == dst) break;
    skip_set(c, src, va);
    apply_skip(c, src, va);
    c->nvm--;
  }
}

static void set_cargo_space(Ctx* c, int sea) {
  c->large_cargo[sea] = c->cur_sea[sea][TRANS_EMPTY] + c->cur_sea[sea][TRANS_1I];
  c->small_cargo[sea] = c->large_cargo[sea] + c->cur_sea[sea][TRANS_1A] + c->cur_sea[sea][TRANS_1T];
}
static int load_transport(Ctx* c, int ut, int src_land, int dst_sea, int land_state) { /* 1397-1443 */
  const uint8_t* load = ut == INFANTRY ? LOAD_INF : ut == ARTILLERY ? LOAD_ART : LOAD_TANK;
  for (uint8_t tt = UNIT_WEIGHTS[ut] > 2 ? TRANS_1I : TRANS_1T; tt >= TRANS_EMPTY; tt--) {
    uint8_t* arr = sea_ptr(c, dst_sea, tt);
    uint8_t unl = STATES_UNLOADING[tt];
    for (uint8_t ts = (uint8_t)(STATES_SEA[tt] - STATES_STAGING[tt]); ts >= unl && ts != 255; ts--) {
      if (arr[ts] == 0) continue;
      uint8_t ntt = load[tt];
      arr[ts]--;
      if (tt == TRANS_EMPTY && ts == 0) ts = STATES_UNLOADING[ntt];
      sea_ptr(c, dst_sea, ntt)[ts]++;
      c->cur_sea[dst_sea][tt]--;
      c->cur_sea[dst_sea][ntt]++;
      c->tot_land[0][src_land]--;
      c->cur_land[src_land][ut]--;
      land_ptr(c, src_land, ut)[land_state]--;
      set_cargo_space(c, dst_sea);
      return 1;
    }
  }
  return 0;
}

static void add_valid_land_moves(Ctx* c, int src, int moves, int ut) { /* 1445-1516 */
  if (moves == 2) {
    for (int i = 0; i < T.lands_within_2_count[src]; i++) {
      int d = T.lands_within_2[src][i];
      if (skip_get(c, src, d)) continue;
      if (T.land_dist[src][d] == 2 && c->land_blocked[src * NL + d]) continue;
      c->vm[c->nvm++] = (uint8_t)d;
    }
    for (int i = 0; i < T.load_within_2_count[src]; i++) {
      int ds = T.load_within_2[src][i];
      if (c->large_cargo[ds] == 0) continue;
      int da = ds + NL;
      if (skip_get(c, src, da)) continue;
      unsigned flat = (unsigned)(src * NL + da); /* 1474: air index into a 5-wide row, P1(iii) */
      if (T.land_dist[src][da] == 1 && (flat < 25u ? c->land_blocked[flat] : 0)) continue;
      c->vm[c->nvm++] = (uint8_t)da;
    }
    return;
  }
  int non_combat = ATTACK_LAND[ut] == 0, unloadable = UNIT_WEIGHTS[ut] > 5, heavy = UNIT_WEIGHTS[ut] > 2;
  for (int i = 0; i < T.l2l_count[src]; i++) {
    int d = T.l2l_conn[src][i];
    if (skip_get(c, src, d)) continue;
    if (non_combat && !c->allied[OWNER(c, d)]) continue;
    c->vm[c->nvm++] = (uint8_t)d;
  }
  if (unloadable) return;
  for (int i = 0; i < T.l2s_count[src]; i++) {
    int ds = T.l2s_conn[src][i];
    if (c->small_cargo[ds] == 0) continue;
    if (heavy && c->large_cargo[ds] == 0) continue;
    int da = ds + NL;
    if (skip_get(c, src, da)) continue;
    c->vm[c->nvm++] = (uint8_t)da;
  }
}

static void add_valid_sea_moves_x(Ctx* c, int src_sea, int moves, const uint8_t* blocked) { /* 1518-1586 */
  int cs = c->canal_state, sa = src_sea + NL;
  if (moves == 2) {
    for (int i = 0; i < T.seas_within_2_count[cs][src_sea]; i++) {
      int d = T.seas_within_2[cs][src_sea][i];
      if (blocked[src_sea * NS + d]) continue;
      if (skip_get(c, sa, d + NL)) continue;
      c->vm[c->nvm++] = (uint8_t)(d + NL);
    }
  } else {
    for (int i = 0; i < T.seas_within_1_count[cs][src_sea]; i++) {
      int d = T.seas_within_1[cs][src_sea][i];
      if (skip_get(c, sa, d + NL)) continue;
      c->vm[c->nvm++] = (uint8_t)(d + NL);
    }
  }
}
static void add_valid_sea_moves(Ctx* c, int s, int m) { add_valid_sea_moves_x(c, s, m, c->sea_blocked); }
static void add_valid_sub_moves(Ctx* c, int s, int m) { add_valid_sea_moves_x(c, s, m, c->sub_blocked); }

/* ---- phase 2: stage_transport_units, 1588-1655 ---- */
static int stage_transport_units(Ctx* c) {
  int processed = 0;
  for (int ut = TRANS_EMPTY; ut <= TRANS_1T; ut++) {
    uint8_t staging = (uint8_t)(STATES_SEA[ut] - 1), done = (uint8_t)(staging - 1);
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(c, s, ut);
      if (arr[staging] == 0) continue;
      uint8_t sa = (uint8_t)(s + NL);
      c->vm[0] = sa;
      c->nvm = 1;
      add_valid_sea_moves(c, s, 2);
      while (arr[staging] > 0) {
        processed = 1;
        uint8_t da = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          da = ai_input(c);
        }
        if (sa == da) {
          arr[done] = (uint8_t)(arr[done] + arr[staging]);
          arr[staging] = 0;
          continue;
        }
        int ds = da - NL;
        unsigned flat = (unsigned)(s * NS + da); /* 1630: sea_dist indexed with an air id, P1(iii) */
        uint8_t dist = flat < 9u ? ((const uint8_t*)T.sea_dist[c->canal_state])[flat] : 0;
        if (c->blockade[ds] > 0) {
          FLAG(c, da) = 2;
          continue;
        }
        sea_ptr(c, ds, ut)[(uint8_t)(done - dist)]++;
        c->cur_sea[ds][ut]++;
        c->tot_sea[0][ds]++;
        c->small_cargo[ds]++;
        arr[staging]--;
        c->cur_sea[s][ut]--;
        c->tot_sea[0][s]--;
        c->small_cargo[s]--;
        if (ut <= TRANS_1I) {
          c->large_cargo[s]--;
          c->large_cargo[ds]++;
        }
      }
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 0: fighters, 1657-1803, 3385-3407 ---- */
static void pre_move_fighter_units(Ctx* c) { /* 1657-1714 */
  memset(c->can_f_here, 0, NA);
  for (int l = 0; l < NL; l++) {
    int owner = OWNER(c, l);
    c->can_f_here[l] = c->allied[owner] && FLAG(c, l) == 0;
    if (owner == c->st.player_index && FMAX(c, l) > 0) /* 1676: relative owner vs absolute player */
      for (int i = 0; i < T.l2s_count[l]; i++) c->can_f_here[NL + T.l2s_conn[l][i]] = 1;
  }
  for (int s = 0; s < NS; s++) {
    if (c->carriers[s] <= 0) continue;
    c->can_f_here[s + NL] = 1;
    if (sea_ptr(c, s, CARRIERS)[2] > 0) { /* 1688: carriers[2] aliases cruisers[0] */
      for (int i = 0; i < T.s2s_count[s]; i++) {
        int s1 = T.s2s_conn[s][i];
        c->can_f_here[NL + s1] = 1;
        for (int j = 0; j < T.s2s_count[s1]; j++) c->can_f_here[NL + T.s2s_conn[s1][j]] = 1;
      }
    }
  }
  for (int a = 0; a < NA; a++) {
    c->can_f_1[a] = 0;
    for (int i = 0; i < T.air_conn_count[a]; i++)
      if (c->can_f_here[T.air_conn[a][i]]) {
        c->can_f_1[a] = 1;
        break;
      }
  }
}
static void add_valid_fighter_moves(Ctx* c, int src, int moves) { /* 3385-3407 */
  for (int i = 0; i < T.air_within_count[moves - 1][src]; i++) {
    int d = T.air_within[moves - 1][src][i];
    int dist = T.air_dist[src][d];
    if (dist <= 2 || c->can_f_here[d] || (dist == 3 && c->can_f_1[d])) {
      if (!c->can_f_here[d] && c->enemy_count[d] == 0) continue;
      if (!skip_get(c, src, d)) c->vm[c->nvm++] = (uint8_t)d;
    }
  }
}
static int move_fighter_units(Ctx* c) { /* 1715-1803 */
  int processed = 0;
  for (int src = 0; src < NA; src++) {
    uint8_t* arr = air_ptr(c, src, FIGHTERS);
    if (arr[4] == 0) continue;
    if (!processed) {
      processed = 1;
      pre_move_fighter_units(c);
    }
    c->vm[0] = (uint8_t)src;
    c->nvm = 1;
    add_valid_fighter_moves(c, src, 4);
    while (arr[4] > 0) {
      uint8_t dst = c->vm[0];
      if (c->nvm > 1) {
        if (c->answers_remaining == 0) return 1;
        dst = ai_input(c);
      }
      update_move_history_4air(c, (uint8_t)src, dst);
      if (src == dst) {
        arr[0] = (uint8_t)(arr[0] + arr[4]);
        arr[4] = 0;
        continue;
      }
      uint8_t dist = T.air_dist[src][dst];
      if (dst < NL) {
        if (!c->allied[OWNER(c, dst)]) FLAG(c, dst) = 2;
        else dist = 4;
        land_ptr(c, dst, FIGHTERS)[(uint8_t)(4 - dist)]++;
        c->cur_land[dst][FIGHTERS]++;
        c->tot_land[0][dst]++;
      } else {
        FLAG(c, dst) = c->enemy_count[dst] > 0;
        sea_ptr(c, dst - NL, FIGHTERS)[(uint8_t)(4 - dist)]++;
        c->cur_sea[dst - NL][FIGHTERS]++;
        c->tot_sea[0][dst - NL]++;
      }
      if (src < NL) {
        c->cur_land[src][FIGHTERS]--;
        c->tot_land[0][src]--;
      } else {
        c->cur_sea[src - NL][FIGHTERS]--;
        c->tot_sea[0][src - NL]--;
      }
      arr[4]--;
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 1: bombers, 1805-1916, 3423-3443 ---- */
static void add_valid_bomber_moves(Ctx* c, int src, int moves) { /* 3423-3443 */
  for (int i = 0; i < T.air_within_count[moves - 1][src]; i++) {
    int d = T.air_within[moves - 1][src][i];
    int dist = T.air_dist[src][d];
    if (dist <= 3 || c->can_b_here[d] || (dist == 4 && c->can_b_2[d]) || (dist == 5 && c->can_b_1[d])) {
      if (!c->can_b_here[d] && c->enemy_count[d] == 0) {
        if (d >= NL || FMAX(c, d) == 0 || FHP(c, d) == -(int)FMAX(c, d)) continue;
      }
      if (!skip_get(c, src, d)) c->vm[c->nvm++] = (uint8_t)d;
    }
  }
}
static int move_bomber_units(Ctx* c) {
  int processed = 0;
  for (int l = 0; l < NL; l++) c->can_b_here[l] = c->allied[OWNER(c, l)] && FLAG(c, l) == 0;
  for (int l = 0; l < NL; l++) {
    c->can_b_1[l] = 0;
    for (int i = 0; i < T.l2l_count[l]; i++)
      if (c->can_b_here[T.l2l_conn[l][i]]) { c->can_b_1[l] = 1; break; }
  }
  for (int s = 0; s < NS; s++) {
    c->can_b_1[NL + s] = 0;
    for (int i = 0; i < T.s2l_count[s]; i++)
      if (c->can_b_here[T.s2l_conn[s][i]]) { c->can_b_1[NL + s] = 1; break; }
  }
  for (int a = 0; a < NA; a++) {
    c->can_b_2[a] = 0;
    for (int i = 0; i < T.air_conn_count[a]; i++)
      if (c->can_b_1[T.air_conn[a][i]]) { c->can_b_2[a] = 1; break; }
  }
  for (int src = 0; src < NL; src++) {
    uint8_t* arr = land_ptr(c, src, BOMBERS_LAND_AIR);
    if (arr[6] == 0) continue;
    processed = 1;
    c->vm[0] = (uint8_t)src;
    c->nvm = 1;
    add_valid_bomber_moves(c, src, 6);
    while (arr[6] > 0) {
      uint8_t dst = c->vm[0];
      if (c->nvm == 1) { /* 1860: inverted test */
        if (c->answers_remaining == 0) return 1;
        dst = ai_input(c);
      }
      if (src == dst) {
        arr[0] = (uint8_t)(arr[0] + arr[6]);
        arr[6] = 0;
        continue;
      }
      if (dst < NL) {
        if (!c->allied[OWNER(c, dst)]) FLAG(c, dst) = 2;
      } else {
        FLAG(c, dst) = c->enemy_count[dst] > 0;
      }
      uint8_t dist = T.air_dist[src][dst];
      if (dst < NL) {
        land_ptr(c, dst, BOMBERS_LAND_AIR)[(uint8_t)(6 - dist)]++;
        c->cur_land[dst][BOMBERS_LAND_AIR]++;
        c->tot_land[0][dst]++;
      } else {
        sea_ptr(c, dst - NL, BOMBERS_SEA)[(uint8_t)(5 - dist)]++;
        c->cur_sea[dst - NL][BOMBERS_SEA]++;
        c->tot_sea[0][dst - NL]++;
      }
      c->cur_land[src][BOMBERS_LAND_AIR]--;
      c->tot_land[0][src]--;
      arr[6]--;
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- conquer_land, 1918-1984 ---- */
static void conquer_land(Ctx* c, int dst) {
  int pi = c->st.player_index;
  uint8_t old = OWNER(c, dst);
  if (T.capital[(pi + old) % NP] == dst) {
    c->st.money[0] = (uint8_t)(c->st.money[0] + c->st.money[old]);
    c->st.money[old] = 0;
  }
  c->income[old] -= T.land_value[dst];
  uint8_t nw = 0;
  uint8_t orig = (uint8_t)((T.original_owner[dst] + NP - pi) % NP);
  if (c->allied[orig]) nw = orig;
  OWNER(c, dst) = nw;
  c->income[nw] += T.land_value[dst];
  c->factloc[nw][c->nfact[nw]++] = dst;
  c->nfact[old]--;
  for (int i = 0; i < c->nfact[old]; i++)
    if (c->factloc[old][i] == dst) {
      for (int j = i; j < c->nfact[old]; j++) c->factloc[old][j] = c->factloc[old][j + 1];
      break;
    }
}

/* ---- phases 3,4,5,12: move_land_unit_type, 1986-2060 ---- */
static int move_land_unit_type(Ctx* c, int ut) {
  int processed = 0;
  for (int src = 0; src < NL; src++)
    for (int moves = MAX_MOVE_LAND[ut]; moves > 0; moves--) {
      uint8_t* arr = land_ptr(c, src, ut);
      if (arr[moves] == 0) continue;
      processed = 1;
      c->vm[0] = (uint8_t)src;
      c->nvm = 1;
      add_valid_land_moves(c, src, moves, ut);
      while (arr[moves] > 0) {
        uint8_t dst = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          dst = ai_input(c);
        }
        if (src == dst) {
          arr[0] = (uint8_t)(arr[0] + arr[moves]);
          arr[moves] = 0;
          continue;
        }
        if (dst >= NL) {
          load_transport(c, ut, src, dst - NL, moves);
          c->nvm = 1;
          add_valid_land_moves(c, src, moves, ut);
          continue;
        }
        FLAG(c, dst) = c->enemy_count[dst] > 0;
        uint8_t dist = T.land_dist[src][dst];
        if (c->allied[OWNER(c, dst)] || c->enemy_count[dst] > 0) dist = (uint8_t)moves;
        land_ptr(c, dst, ut)[(uint8_t)(moves - dist)]++;
        c->cur_land[dst][ut]++;
        c->tot_land[0][dst]++;
        c->cur_land[src][ut]--;
        c->tot_land[0][src]--;
        arr[moves]--;
        if (!c->allied[OWNER(c, dst)] && c->enemy_count[dst] == 0) {
          conquer_land(c, dst);
          FLAG(c, dst) = 2;
        }
      }
    }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 6: move_transport_units, 2062-2159 ---- */
static int move_transport_units(Ctx* c) {
  int processed = 0;
  for (int s = 0; s < NS; s++) { /* skip_empty_transports 2062-2075 */
    uint8_t* e = sea_ptr(c, s, TRANS_EMPTY);
    for (int k = 3; k >= 1; k--) {
      e[0] = (uint8_t)(e[0] + e[k]);
      e[k] = 0;
    }
  }
  for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
    uint8_t max_state = (uint8_t)(STATES_SEA[ut] - STATES_STAGING[ut]);
    for (int s = 0; s < NS; s++)
      for (int i = 1; i < 3; i++) {
        uint8_t cur = (uint8_t)(max_state - i);
        uint8_t* arr = sea_ptr(c, s, ut);
        if (arr[cur] == 0) continue;
        processed = 1;
        uint8_t sa = (uint8_t)(s + NL);
        c->vm[0] = sa;
        c->nvm = 1;
        add_valid_sea_moves(c, s, cur - 1);
        while (arr[cur] > 0) {
          uint8_t da = c->vm[0];
          if (c->nvm > 1) {
            if (c->answers_remaining == 0) return 1;
            da = ai_input(c);
          }
          int ds = da - NL;
          if (c->blockade[ds] > 0) FLAG(c, da) = 2;
          if (sa == da) {
            arr[1] = (uint8_t)(arr[1] + arr[cur]);
            arr[cur] = 0;
            continue;
          }
          sea_ptr(c, ds, ut)[1]++;
          c->cur_sea[ds][ut]++;
          c->tot_sea[0][ds]++;
          arr[cur]--;
          c->cur_sea[s][ut]--;
          c->tot_sea[0][s]--;
          if (ut <= TRANS_1T) {
            c->small_cargo[ds]++;
            c->small_cargo[s]--;
            if (ut <= TRANS_1I) {
              c->large_cargo[ds]++;
              c->large_cargo[s]--;
            }
          }
        }
      }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 7: move_subs, 2160-2214 ---- */
static int move_subs(Ctx* c) {
  int processed = 0;
  for (int s = 0; s < NS; s++) {
    uint8_t* arr = sea_ptr(c, s, SUBMARINES);
    if (arr[2] == 0) continue;
    uint8_t sa = (uint8_t)(s + NL);
    c->vm[0] = sa;
    c->nvm = 1;
    add_valid_sub_moves(c, s, 2);
    while (arr[2] > 0) {
      processed = 1;
      uint8_t da = c->vm[0];
      if (c->nvm > 1) {
        if (c->answers_remaining == 0) return 1;
        da = ai_input(c);
      }
      int ds = da - NL;
      if (c->enemy_count[ds] > 0) FLAG(c, ds) = 2; /* 2188,2194: sea index used as air index */
      if (sa == da) {
        arr[0] = (uint8_t)(arr[0] + arr[2]);
        arr[2] = 0;
        continue;
      }
      sea_ptr(c, ds, SUBMARINES)[0]++;
      c->cur_sea[ds][SUBMARINES]++;
      c->tot_sea[0][ds]++;
      arr[2]--;
      c->cur_sea[s][SUBMARINES]--;
      c->tot_sea[0][s]--;
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 8: move_destroyers_battleships, 2216-2310 ---- */
static void carry_allied_fighters(Ctx* c, int src, int dst) { /* 2286-2310 */
  int moved = 0;
  for (int p = 1; p < NP; p++) {
    if (!c->allied[p]) continue;
    while (tpsut(c, p, src)[FIGHTERS] > 0) {
      tpsut(c, p, src)[FIGHTERS]--;
      c->tot_sea[p][src]--;
      tpsut(c, p, dst)[FIGHTERS]++;
      c->tot_sea[p][dst]++;
      if (moved == 1) return;
      moved++;
    }
  }
}
static int move_destroyers_battleships(Ctx* c) {
  int processed = 0;
  for (int ut = DESTROYERS; ut <= BS_DAMAGED; ut++) {
    uint8_t unmoved = UNMOVED_SEA[ut], done = DONE_MOVING_SEA[ut];
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(c, s, ut);
      if (arr[unmoved] == 0) continue;
      processed = 1;
      uint8_t sa = (uint8_t)(s + NL);
      c->vm[0] = sa;
      c->nvm = 1;
      add_valid_sea_moves(c, s, MAX_MOVE_SEA[ut]);
      while (arr[unmoved] > 0) {
        uint8_t da = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          da = ai_input(c);
        }
        if (c->enemy_count[da] > 0) FLAG(c, da) = 2;
        if (sa == da) {
          arr[done] = (uint8_t)(arr[done] + arr[unmoved]);
          arr[unmoved] = 0;
          continue;
        }
        int ds = da - NL;
        sea_ptr(c, ds, ut)[done]++;
        c->cur_sea[ds][ut]++;
        c->tot_sea[0][ds]++;
        arr[unmoved]--;
        c->cur_sea[s][ut]--;
        c->tot_sea[0][s]--;
        if (ut == CARRIERS) carry_allied_fighters(c, s, ds);
      }
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- casualty removal ---- */
static void remove_land_defenders(Ctx* c, int land, uint8_t hits) { /* 2613-2641 */
  for (int k = 0; k < 6; k++)
    for (int e = 0; e < c->n_enemies; e++) {
      int p = c->enemies[e];
      uint8_t* n = &tplut(c, p, land)[ORDER_LAND_DEF[k]];
      if (*n == 0) continue;
      if (*n < hits) {
        hits = (uint8_t)(hits - *n);
        c->tot_land[p][land] -= *n;
        c->enemy_count[land] -= *n;
        *n = 0;
      } else {
        *n = (uint8_t)(*n - hits);
        c->tot_land[p][land] -= hits;
        c->enemy_count[land] -= hits;
        return;
      }
    }
}
static void remove_land_attackers(Ctx* c, int land, uint8_t hits) { /* 2642-2698 */
  for (int k = 0; k < 3; k++) {
    int ut = ORDER_LAND_ATT1[k];
    uint8_t* n = &land_ptr(c, land, ut)[0];
    if (*n == 0) continue;
    if (*n < hits) {
      hits = (uint8_t)(hits - *n);
      c->tot_land[0][land] -= *n;
      *n = 0;
      c->cur_land[land][ut] = 0;
    } else {
      *n = (uint8_t)(*n - hits);
      c->cur_land[land][ut] = (uint8_t)(c->cur_land[land][ut] - hits);
      c->tot_land[0][land] -= hits;
      return;
    }
  }
  for (int k = 0; k < 2; k++) {
    int ut = ORDER_LAND_ATT2[k];
    if (c->cur_land[land][ut] == 0) continue;
    for (int s = 1; s < STATES_LAND[ut] - 1; s++) {
      uint8_t* n = &land_ptr(c, land, ut)[s];
      if (*n == 0) continue;
      if (*n < hits) {
        hits = (uint8_t)(hits - *n);
        c->cur_land[land][ut] = (uint8_t)(c->cur_land[land][ut] - *n);
        c->tot_land[0][land] -= *n;
        *n = 0;
      } else {
        *n = (uint8_t)(*n - hits);
        c->cur_land[land][ut] = (uint8_t)(c->cur_land[land][ut] - hits);
        c->tot_land[0][land] -= hits;
        return;
      }
    }
  }
}
static void remove_sea_defenders(Ctx* c, int sea, uint8_t hits, int submerged) { /* 2700-2806 */
  int air = sea + NL;
  for (int e = 0; e < c->n_enemies; e++) { /* battleships absorb one hit each */
    uint8_t* su = tpsut(c, c->enemies[e], sea);
    if (su[BATTLESHIPS] == 0) continue;
    if (su[BATTLESHIPS] < hits) {
      hits = (uint8_t)(hits - su[BATTLESHIPS]);
      su[BS_DAMAGED] = (uint8_t)(su[BS_DAMAGED] + su[BATTLESHIPS]);
      su[BATTLESHIPS] = 0;
    } else {
      su[BS_DAMAGED] = (uint8_t)(su[BS_DAMAGED] + hits);
      su[BATTLESHIPS] = (uint8_t)(su[BATTLESHIPS] - hits);
      return;
    }
  }
  if (!submerged)
    for (int e = 0; e < c->n_enemies; e++) {
      int p = c->enemies[e];
      uint8_t* n = &tpsut(c, p, sea)[SUBMARINES];
      if (*n == 0) continue;
      if (*n < hits) {
        hits = (uint8_t)(hits - *n);
        c->tot_sea[p][sea] -= *n;
        c->enemy_count[air] -= *n;
        *n = 0;
      } else {
        *n = (uint8_t)(*n - hits);
        c->tot_sea[p][sea] -= hits;
        c->enemy_count[air] -= hits;
        return;
      }
    }
  for (int k = 1; k < 13; k++)
    for (int e = 0; e < c->n_enemies; e++) {
      int p = c->enemies[e];
      uint8_t* n = &tpsut(c, p, sea)[ORDER_SEA_DEF[k]];
      if (*n == 0) continue;
      int in_blockade_range = k >= DESTROYERS && k <= BS_DAMAGED; /* 2783,2791: order index vs type ids */
      if (*n < hits) {
        hits = (uint8_t)(hits - *n);
        c->tot_sea[p][sea] -= *n;
        c->enemy_count[air] -= *n;
        if (in_blockade_range) c->blockade[sea] = (uint8_t)(c->blockade[sea] - *n);
        *n = 0;
      } else {
        *n = (uint8_t)(*n - hits);
        c->tot_sea[p][sea] -= hits;
        c->enemy_count[air] -= hits;
        if (in_blockade_range) c->blockade[sea] = (uint8_t)(c->blockade[sea] - hits);
        return;
      }
    }
}
/* one attacker casualty step on state-0 units of `ut`; returns 1 when all hits are absorbed */
static int hit_sea_state0(Ctx* c, int sea, int ut, uint8_t* hits) {
  uint8_t* n = &sea_ptr(c, sea, ut)[0];
  if (*n == 0) return 0;
  if (*n < *hits) {
    *hits = (uint8_t)(*hits - *n);
    c->tot_sea[0][sea] -= *n;
    if (ut >= TRANS_EMPTY && ut <= TRANS_1T) {
      c->small_cargo[sea] -= *n;
      if (ut <= TRANS_1I) c->large_cargo[sea] -= *n;
    }
    *n = 0;
    c->cur_sea[sea][ut] = 0;
    if (ut == CARRIERS) c->carriers[sea] = 0;
    return 0;
  }
  *n = (uint8_t)(*n - *hits);
  c->cur_sea[sea][ut] = (uint8_t)(c->cur_sea[sea][ut] - *hits);
  c->tot_sea[0][sea] -= *hits;
  if (ut == CARRIERS) c->carriers[sea] -= *hits;
  if (ut >= TRANS_EMPTY && ut <= TRANS_1T) {
    c->small_cargo[sea] -= *hits;
    if (ut <= TRANS_1I) c->large_cargo[sea] -= *hits;
  }
  return 1;
}
static void remove_sea_attackers(Ctx* c, int sea, uint8_t hits) { /* 2808-2982 */
  uint8_t* bb = &sea_ptr(c, sea, BATTLESHIPS)[0];
  uint8_t* bd = &sea_ptr(c, sea, BS_DAMAGED)[0];
  if (*bb > 0) {
    if (*bb < hits) {
      hits = (uint8_t)(hits - *bb);
      *bd = (uint8_t)(*bd + *bb);
      c->cur_sea[sea][BS_DAMAGED] = (uint8_t)(c->cur_sea[sea][BS_DAMAGED] + *bb);
      *bb = 0;
      c->cur_sea[sea][BATTLESHIPS] = 0;
    } else {
      *bd = (uint8_t)(*bd + hits);
      c->cur_sea[sea][BS_DAMAGED] = (uint8_t)(c->cur_sea[sea][BS_DAMAGED] + hits);
      *bb = (uint8_t)(*bb - hits);
      c->cur_sea[sea][BATTLESHIPS] = (uint8_t)(c->cur_sea[sea][BATTLESHIPS] - hits);
      return;
    }
  }
  for (int k = 0; k < 2; k++)
    if (hit_sea_state0(c, sea, ORDER_SEA_ATT1[k], &hits)) return;
  if (hit_sea_state0(c, sea, CARRIERS, &hits)) return;
  if (hit_sea_state0(c, sea, CRUISERS, &hits)) return;
  for (int k = 0; k < 2; k++) { /* air units in any state, 2915-2942 */
    int ut = ORDER_SEA_ATT2[k];
    if (c->cur_sea[sea][ut] == 0) continue;
    for (int s = 0; s < STATES_SEA[ut]; s++) {
      uint8_t* n = &sea_ptr(c, sea, ut)[s];
      if (*n == 0) continue;
      if (*n < hits) {
        hits = (uint8_t)(hits - *n);
        c->tot_sea[0][sea] -= *n;
        c->cur_sea[sea][ut] = (uint8_t)(c->cur_sea[sea][ut] - *n);
        *n = 0;
      } else {
        *n = (uint8_t)(*n - hits);
        c->cur_sea[sea][ut] = (uint8_t)(c->cur_sea[sea][ut] - hits);
        c->tot_sea[0][sea] -= hits;
        return;
      }
    }
  }
  for (int k = 0; k < 8; k++)
    if (hit_sea_state0(c, sea, ORDER_SEA_ATT3[k], &hits)) return;
}

/* ---- phase 9: resolve_sea_battles, 2312-2550, 2582-2603 ---- */
static void sea_retreat(Ctx* c, int src, int dst) { /* 2582-2603 */
  for (int ut = TRANS_EMPTY; ut <= BS_DAMAGED; ut++) {
    uint8_t n = c->cur_sea[src][ut];
    uint8_t* d = sea_ptr(c, dst, ut);
    d[0] = (uint8_t)(d[0] + n);
    c->cur_sea[dst][ut] = (uint8_t)(c->cur_sea[dst][ut] + n);
    c->tot_sea[0][dst] += n;
    c->tot_sea[0][src] -= n;
    sea_ptr(c, src, ut)[0] = 0;
    sea_ptr(c, src, ut)[1] = 0;
    c->cur_sea[src][ut] = 0;
  }
  FLAG(c, src + NL) = 0;
}
static int resolve_sea_battles(Ctx* c) {
  for (int s = 0; s < NS; s++) {
    int air = s + NL;
    uint8_t* mine = c->cur_sea[s];
    if (FLAG(c, air) == 0) continue;
    if (c->tot_sea[0][s] == 0) continue;
    int submerged = mine[DESTROYERS] == 0;
    if (c->tot_sea[0][s] == 2) { /* 2336: literal "== 2" */
      if (submerged) {
        uint8_t subs = 0;
        for (int e = 0; e < c->n_enemies; e++) subs = (uint8_t)(subs + tpsut(c, c->enemies[e], s)[SUBMARINES]);
        if (subs == c->enemy_count[s]) continue; /* 2344: land index */
      }
      for (int ut = CRUISERS; ut <= BS_DAMAGED; ut++) {
        uint8_t* a = sea_ptr(c, s, ut);
        a[0] = (uint8_t)(a[0] + a[1]);
        a[1] = 0;
      }
    }
    for (int rounds = 0;; rounds++) {
      if (rounds >= CAAA_SEA_ROUND_CAP) { /* not in the reference: with the deterministic dice of the
                                             expansion mode both sides can score 0 hits for ever */
        c->stuck = 1;
        return 1;
      }
      if (FLAG(c, air) == 1) { /* retreat option, 2368-2406 */
        if (mine[FIGHTERS] + mine[BOMBERS_SEA] + mine[SUBMARINES] + mine[DESTROYERS] + mine[CARRIERS] +
                    mine[CRUISERS] + mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0 ||
            c->blockade[s] == 0) {
          c->vm[0] = (uint8_t)air;
          c->nvm = 1;
        } else {
          c->nvm = 0;
        }
        for (int i = 0; i < T.s2s_count[s]; i++) {
          int d = T.s2s_conn[s][i];
          if (c->blockade[d] == 0) c->vm[c->nvm++] = (uint8_t)(d + NL);
        }
        if (c->nvm > 0) {
          uint8_t da;
          if (c->nvm == 1) {
            da = c->vm[0];
          } else {
            if (c->answers_remaining == 0) return 1;
            da = ai_input(c);
          }
          uint8_t ds = (uint8_t)(da - NL);
          if (T.sea_dist[c->canal_state][s][ds] == 1 && FLAG(c, da) == 0) {
            sea_retreat(c, s, ds);
            break;
          }
        }
      }
      FLAG(c, air) = 1;
      int targets = 0;
      if (mine[DESTROYERS]) {
        if (c->enemy_count[air] > 0) targets = 1;
      } else if (mine[CARRIERS] + mine[CRUISERS] + mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0) {
        if (c->enemy_count[air] > 0)
          for (int e = 0; e < c->n_enemies; e++) {
            int p = c->enemies[e];
            if (c->tot_sea[p][s] - tpsut(c, p, s)[SUBMARINES] > 0) { targets = 1; break; }
          }
      } else if (mine[SUBMARINES] > 0) {
        for (int e = 0; e < c->n_enemies; e++) {
          int p = c->enemies[e];
          if (c->tot_sea[p][s] - (tpsut(c, p, s)[FIGHTERS] + tpsut(c, p, s)[SUBMARINES]) > 0) { targets = 1; break; }
        }
      } else if (mine[FIGHTERS] + mine[BOMBERS_SEA] > 0) {
        for (int e = 0; e < c->n_enemies; e++) {
          int p = c->enemies[e];
          if (c->tot_sea[p][s] - tpsut(c, p, s)[SUBMARINES] > 0) { targets = 1; break; }
        }
      }
      if (!targets) { /* 2460-2482 */
        if (c->enemy_count[air] > 0) {
          c->tot_sea[0][s] -= mine[TRANS_EMPTY];
          mine[TRANS_EMPTY] = 0;
          sea_ptr(c, s, TRANS_EMPTY)[0] = 0;
          for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
            c->tot_sea[0][s] -= mine[ut];
            mine[ut] = 0;
            sea_ptr(c, s, ut)[1] = 0;
          }
        }
        FLAG(c, air) = 0;
        break;
      }
      /* sub volley: only state-0 attacking subs count (2484) */
      int admg = sea_ptr(c, s, SUBMARINES)[0] * 2;
      uint8_t ahits = attacker_hits(c, admg);
      int ddmg = 0, dhits = 0;
      if (!submerged) {
        for (int e = 0; e < c->n_enemies; e++) ddmg += tpsut(c, c->enemies[e], s)[SUBMARINES];
        dhits = defender_hits(c, ddmg);
        if (dhits > 0) remove_sea_attackers(c, s, (uint8_t)dhits);
      }
      if (ahits > 0) remove_sea_defenders(c, s, ahits, submerged);
      /* main volley, 2513-2542 */
      admg = 0;
      for (int k = 0; k < 5; k++) admg += mine[BLOCKADE_TYPES[k]] * ATTACK_SEA[BLOCKADE_TYPES[k]];
      admg += mine[FIGHTERS] * ATTACK_SEA[FIGHTERS];
      admg += mine[BOMBERS_SEA] * ATTACK_SEA[BOMBERS_SEA];
      ahits = attacker_hits(c, admg);
      ddmg = 0;
      for (int e = 0; e < c->n_enemies; e++) {
        const uint8_t* eu = tpsut(c, c->enemies[e], s);
        for (int ut = 0; ut < 5; ut++) ddmg += eu[ut] * DEFENSE_SEA[ut]; /* 2527: type ids 0..4 */
        ddmg += eu[FIGHTERS] * DEFENSE_SEA[FIGHTERS];
      }
      dhits = defender_hits(c, ddmg);
      if (dhits > 0) remove_sea_attackers(c, s, (uint8_t)dhits);
      if (ahits > 0) remove_sea_defenders(c, s, ahits, submerged);
      if (c->enemy_count[air] == 0 || c->tot_sea[0][s] == 0) {
        FLAG(c, air) = 0;
        break;
      }
    }
  }
  return 0;
}

/* ---- phase 10: unload_transports, 2984-3053, 3373-3383 ---- */
static int unload_transports(Ctx* c) {
  int processed = 0;
  for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
    uint8_t ust = STATES_UNLOADING[ut], c1 = UNLOAD_CARGO1[ut], c2 = UNLOAD_CARGO2[ut];
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(c, s, ut);
      if (arr[ust] == 0) continue;
      processed = 1;
      uint8_t sa = (uint8_t)(s + NL);
      c->vm[0] = sa;
      c->nvm = 1;
      for (int i = 0; i < T.s2l_count[s]; i++) {
        int d = T.s2l_conn[s][i];
        if (!skip_get(c, sa, d)) c->vm[c->nvm++] = (uint8_t)d;
      }
      while (arr[ust] > 0) {
        uint8_t dst = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          dst = ai_input(c);
        }
        if (sa == dst) {
          arr[0] = (uint8_t)(arr[0] + arr[ust]);
          arr[ust] = 0;
          continue;
        }
        BOMBARD(c, dst)++;
        land_ptr(c, dst, c1)[0]++;
        c->cur_land[dst][c1]++;
        c->tot_land[0][dst]++;
        sea_ptr(c, s, TRANS_EMPTY)[0]++;
        c->cur_sea[s][TRANS_EMPTY]++;
        c->cur_sea[s][ut]--;
        arr[ust]--;
        if (ut > TRANS_1T) {
          BOMBARD(c, dst)++;
          land_ptr(c, dst, c2)[0]++;
          c->cur_land[dst][c2]++;
          c->tot_land[0][dst]++;
        }
        if (!c->allied[OWNER(c, dst)]) {
          FLAG(c, dst) = 2;
          if (c->enemy_count[dst] == 0) conquer_land(c, dst);
        }
      }
    }
  }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 11: resolve_land_battles, 3055-3371 ---- */
static void hit_states(Ctx* c, int land, int ut, int lo, int hi, uint8_t* hits, int stop_on_absorb) {
  /* shared shape of the AA-fire / strategic-bombing casualty loops (3103-3117, 3179-3210) */
  for (int s = lo; s < hi; s++) {
    uint8_t* n = &land_ptr(c, land, ut)[s];
    if (*n < *hits) {
      *hits = (uint8_t)(*hits - *n);
      c->cur_land[land][ut] = (uint8_t)(c->cur_land[land][ut] - *n);
      c->tot_land[0][land] -= *n;
      *n = 0;
    } else {
      *n = (uint8_t)(*n - *hits);
      c->cur_land[land][ut] = (uint8_t)(c->cur_land[land][ut] - *hits);
      c->tot_land[0][land] -= *hits;
      *hits = 0;
      if (stop_on_absorb) return;
    }
  }
}
static int resolve_land_battles(Ctx* c) {
  for (int l = 0; l < NL; l++) {
    if (FLAG(c, l) == 0) continue;
    uint8_t* mine = c->cur_land[l];
    int* my_total = &c->tot_land[0][l];
    int admg;
    uint8_t ahits = 0; /* the reference leaves this uninitialised on one path (3075-3123); result is the same */
    if (*my_total == 0) continue;
    if (FLAG(c, l) == 2) {
      if (mine[BOMBERS_LAND_AIR] > 0 && *my_total == mine[BOMBERS_LAND_AIR]) { /* strategic bombing 3088-3125 */
        if (FHP(c, l) > -(int)FMAX(c, l)) {
          uint8_t ddmg = mine[BOMBERS_LAND_AIR];
          uint8_t dh = defender_hits(c, ddmg);
          if (dh > 0) hit_states(c, l, BOMBERS_LAND_AIR, 1, 6, &dh, 0);
          admg = mine[BOMBERS_LAND_AIR] * 21;
          ahits = attacker_hits(c, admg);
        }
        int hp = FHP(c, l) - ahits, floor_hp = -(int)FMAX(c, l);
        FHP(c, l) = (int8_t)(hp > floor_hp ? hp : floor_hp);
        continue;
      }
      if (BOMBARD(c, l) > 0) { /* shore bombardment 3134-3158 */
        admg = 0;
        for (int ut = BS_DAMAGED; ut >= CRUISERS; ut--)
          for (int i = 0; i < T.l2s_count[l]; i++) {
            uint8_t* ships = sea_ptr(c, T.l2s_conn[l][i], ut);
            while (ships[1] > 0 && BOMBARD(c, l) > 0) {
              admg += ATTACK_SEA[ut];
              ships[0]++;
              ships[1]--;
              BOMBARD(c, l)--;
            }
          }
        BOMBARD(c, l) = 0;
        ahits = attacker_hits(c, admg);
        if (ahits > 0) remove_land_defenders(c, l, ahits);
      }
      uint8_t air_units = (uint8_t)(mine[FIGHTERS] + mine[BOMBERS_LAND_AIR]); /* AA fire 3160-3216 */
      if (air_units > 0) {
        int aa = 0;
        for (int e = 0; e < c->n_enemies; e++) aa += tplut(c, c->enemies[e], l)[AA_GUNS];
        if (aa > 0) {
          uint8_t ddmg = (uint8_t)(air_units * 3);
          uint8_t dh = defender_hits(c, ddmg);
          if (dh > 0) hit_states(c, l, FIGHTERS, 0, 5, &dh, 1);
          if (dh > 0) hit_states(c, l, BOMBERS_LAND_AIR, 0, 7, &dh, 1);
        }
      }
      if (*my_total == 0) continue;
      if (c->enemy_count[l] == 0) {
        if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(c, l);
        continue;
      }
    }
    uint8_t rounds = 0;
    for (;;) {
      rounds++;
      if (*my_total == 0) break;
      if (FLAG(c, l) == 1) { /* retreat option 3256-3299 */
        c->vm[0] = (uint8_t)l;
        c->nvm = 1;
        for (int i = 0; i < T.l2l_count[l]; i++) {
          int d = T.l2l_conn[l][i];
          if (c->enemy_count[d] == 0 && FLAG(c, d) == 0 && c->allied[OWNER(c, d)]) c->vm[c->nvm++] = (uint8_t)d;
        }
        uint8_t dst;
        if (c->nvm == 1) {
          dst = c->vm[0];
        } else {
          if (c->answers_remaining == 0) return 1;
          dst = ai_input(c);
        }
        if (l != dst || rounds > 100) {
          for (int ut = INFANTRY; ut <= TANKS; ut++) {
            int n = land_ptr(c, l, ut)[0];
            land_ptr(c, dst, ut)[0] = (uint8_t)(land_ptr(c, dst, ut)[0] + n);
            c->cur_land[dst][ut] = (uint8_t)(c->cur_land[dst][ut] + n);
            c->tot_land[0][dst] += n;
            c->cur_land[l][ut] = (uint8_t)(c->cur_land[l][ut] - n);
            *my_total -= n;
            land_ptr(c, l, ut)[0] = 0;
          }
          FLAG(c, l) = 0;
          break;
        }
      }
      FLAG(c, l) = 1;
      int inf = mine[INFANTRY], art = mine[ARTILLERY];
      admg = mine[FIGHTERS] * 3 + mine[BOMBERS_LAND_AIR] * 4 + inf * 1 + art * 2 + mine[TANKS] * 3;
      admg += inf < art ? inf : art;
      ahits = attacker_hits(c, admg);
      int ddmg = 0;
      uint8_t dh = 0;
      for (int e = 0; e < c->n_enemies; e++) { /* 3319-3327: one draw per enemy, last one wins */
        const uint8_t* eu = tplut(c, c->enemies[e], l);
        ddmg += eu[INFANTRY] * 2 + eu[ARTILLERY] * 2 + eu[TANKS] * 3 + eu[FIGHTERS] * 4 + eu[BOMBERS_LAND_AIR] * 1;
        dh = defender_hits(c, ddmg);
      }
      if (dh > 0) remove_land_attackers(c, l, dh);
      if (ahits > 0) remove_land_defenders(c, l, ahits);
      if (*my_total == 0) break;
      if (c->enemy_count[l] == 0) {
        if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(c, l);
        break;
      }
    }
  }
  return 0;
}

/* ---- phases 13/14: landing, 3409-3421, 3445-3654 ---- */
static void refresh_can_fighter_land_here(Ctx* c) { /* 3445-3466 */
  memset(c->can_f_here, 0, NA);
  for (int s = 0; s < NS; s++) c->can_f_here[s + NL] = c->carriers[s] > 0;
  for (int l = 0; l < NL; l++) {
    int owner = OWNER(c, l);
    c->can_f_here[l] = c->allied[owner] && FLAG(c, l) == 0;
    if (FMAX(c, l) > 0 && owner == c->st.player_index)
      for (int i = 0; i < T.l2s_count[l]; i++) c->can_f_here[NL + T.l2s_conn[l][i]] = 1;
  }
}
static int land_fighter_units(Ctx* c) { /* 3468-3536 */
  int processed = 0;
  for (int cur = 1; cur < 4; cur++)
    for (int src = 0; src < NA; src++) {
      uint8_t* arr = air_ptr(c, src, FIGHTERS);
      if (arr[cur] == 0) continue;
      if (!processed) {
        processed = 1;
        refresh_can_fighter_land_here(c);
      }
      c->nvm = 0;
      for (int i = 0; i < T.air_within_count[cur - 1][src]; i++) { /* add_valid_fighter_landing */
        int d = T.air_within[cur - 1][src][i];
        if (c->can_f_here[d] && !skip_get(c, src, d)) c->vm[c->nvm++] = (uint8_t)d;
      }
      if (c->nvm == 0 || c->can_f_here[src]) c->vm[c->nvm++] = (uint8_t)src;
      while (arr[cur] > 0) {
        uint8_t dst = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          dst = ai_input(c);
        }
        if (src == dst) {
          arr[0]++;
          arr[cur]--;
          continue;
        }
        air_ptr(c, dst, FIGHTERS)[0]++;
        if (dst < NL) {
          c->tot_land[0][dst]++;
          c->cur_land[dst][FIGHTERS]++;
        } else {
          c->tot_sea[0][dst - NL]++;
          c->cur_sea[dst - NL][FIGHTERS]++;
        }
        if (src < NL) {
          c->tot_land[0][src]--;
          c->cur_land[src][FIGHTERS]--;
        } else {
          c->tot_sea[0][src - NL]--;
          c->cur_sea[src - NL][FIGHTERS]--;
        }
        arr[cur]--;
      }
    }
  if (processed) clear_move_history(c);
  return 0;
}
static int land_bomber_units(Ctx* c) { /* 3565-3654 */
  int processed = 0;
  for (int cur1 = 0; cur1 < 5; cur1++)
    for (int src = 0; src < NA; src++) {
      int cur = src < NL ? cur1 + 1 : cur1;
      uint8_t* arr = air_ptr(c, src, BOMBERS_LAND_AIR);
      if (arr[cur] == 0) continue;
      if (!processed) {
        processed = 1;
        for (int l = 0; l < NL; l++) c->can_b_here[l] = c->allied[OWNER(c, l)] && FLAG(c, l) == 0; /* 3557-3563 */
      }
      c->nvm = 0;
      int moves = cur + (src < NL ? 0 : 1);
      for (int i = 0; i < T.air_to_land_within_count[moves - 1][src]; i++) { /* add_valid_bomber_landing */
        int d = T.air_to_land_within[moves - 1][src][i];
        if (c->can_b_here[d]) c->vm[c->nvm++] = (uint8_t)d;
      }
      while (arr[cur] > 0) {
        processed = 1;
        if (c->nvm == 0) c->vm[c->nvm++] = (uint8_t)src;
        uint8_t dst = c->vm[0];
        if (c->nvm > 1) {
          if (c->answers_remaining == 0) return 1;
          dst = ai_input(c);
        }
        if (src == dst) {
          arr[0]++;
          arr[cur]--;
          continue;
        }
        air_ptr(c, dst, BOMBERS_LAND_AIR)[0]++;
        c->tot_land[0][dst]++;
        c->cur_land[dst][BOMBERS_LAND_AIR]++;
        if (src < NL) {
          c->tot_land[0][src]--;
          c->cur_land[src][BOMBERS_LAND_AIR]--;
        } else {
          c->tot_sea[0][src - NL]--;
          c->cur_sea[src - NL][BOMBERS_LAND_AIR]--; /* 3638: type index 1 = empty transports */
        }
        arr[cur]--;
      }
    }
  if (processed) clear_move_history(c);
  return 0;
}

/* ---- phase 15: buy_units (purchase and placement), 3663-3834 ---- */
static int buy_units(Ctx* c) {
  int processed;
  for (uint8_t f = 0; f < c->nfact[0]; f++) {
    int land = c->factloc[0][f];
    if (c->st.builds_left[land] == 0) continue;
    uint8_t repair = 0;
    for (int si = 0; si < T.l2s_count[land]; si++) {
      processed = 0;
      int sea = T.l2s_conn[land][si], air = sea + NL;
      c->vm[0] = NST; /* pass */
      uint8_t last = 0;
      while (c->st.builds_left[air] > 0) {
        if (c->st.money[0] < 8) {
          c->st.builds_left[air] = 0;
          break;
        }
        uint8_t built = (uint8_t)(FMAX(c, land) - c->st.builds_left[land]);
        if (FHP(c, land) <= built) repair = (uint8_t)(1 + built - FHP(c, land));
        c->nvm = 1;
        for (int k = 6; k >= 0; k--) {
          uint8_t ut = BUY_SEA[k];
          if (skip_get(c, 0, ut)) {
            last = ut;
            break;
          }
          if (c->st.money[0] < COST_SEA[ut] + repair) continue;
          if (ut == FIGHTERS) {
            int fighters = 0;
            for (int p = 0; p < NP; p++) fighters += tpsut(c, p, sea)[FIGHTERS];
            if (c->carriers[sea] * 2 <= fighters) continue;
          }
          c->vm[c->nvm++] = ut;
        }
        if (c->nvm == 1) {
          c->st.builds_left[air] = 0;
          break;
        }
        if (c->answers_remaining == 0) return 1;
        processed = 1;
        uint8_t buy = ai_input(c);
        if (buy == NST) {
          c->st.builds_left[air] = 0;
          break;
        }
        for (int s2 = si; s2 < T.l2s_count[land]; s2++) c->st.builds_left[T.l2s_conn[land][s2] + NL]--;
        c->st.builds_left[land]--;
        FHP(c, land) = (int8_t)(FHP(c, land) + repair);
        c->st.money[0] = (uint8_t)(c->st.money[0] - (COST_SEA[buy] + repair));
        sea_ptr(c, sea, buy)[0]++;
        c->tot_sea[0][sea]++;
        c->cur_sea[sea][buy]++;
        if (buy > last) {
          for (uint8_t u = last; u < buy; u++) skip_set(c, 0, u);
          last = buy;
        }
      }
      if (processed) clear_move_history(c);
    }
    c->vm[0] = NLT; /* pass */
    uint8_t last = 0;
    processed = 0;
    while (c->st.builds_left[land] > 0) {
      if (c->st.money[0] < 3) {
        c->st.builds_left[land] = 0;
        break;
      }
      uint8_t built = (uint8_t)(FMAX(c, land) - c->st.builds_left[land]);
      if (FHP(c, land) <= built) repair = (uint8_t)(1 + built - FHP(c, land));
      c->nvm = 1;
      for (int ut = NLT - 1; ut >= 0; ut--) {
        if (skip_get(c, 0, (unsigned)ut)) {
          last = (uint8_t)ut;
          break;
        }
        if (c->st.money[0] < COST_LAND[ut] + repair) continue;
        c->vm[c->nvm++] = (uint8_t)ut;
      }
      if (c->nvm == 1) {
        c->st.builds_left[land] = 0;
        break;
      }
      if (c->answers_remaining == 0) return 1;
      processed = 1;
      uint8_t buy = ai_input(c);
      if (buy == NLT) {
        c->st.builds_left[land] = 0;
        break;
      }
      c->st.builds_left[land]--;
      FHP(c, land) = (int8_t)(FHP(c, land) + repair);
      c->st.money[0] = (uint8_t)(c->st.money[0] - (COST_LAND[buy] + repair));
      land_ptr(c, land, buy)[0]++;
      c->tot_land[0][land]++;
      c->cur_land[land][buy]++;
      if (buy > last)
        for (uint8_t u = last; u < buy; u++) skip_set(c, 0, u);
      last = buy;
    }
    if (processed) clear_move_history(c);
  }
  return 0;
}

/* ---- phases 16-19: end-of-turn bookkeeping, 3836-3894 ---- */
static void crash_air_units(Ctx* c) {
  pre_move_fighter_units(c);
  for (int l = 0; l < NL; l++) {
    if (c->can_f_here[l]) continue;
    if (c->cur_land[l][FIGHTERS] == 0) continue;
    c->tot_land[0][l] -= c->cur_land[l][FIGHTERS];
    c->cur_land[l][FIGHTERS] = 0;
    land_ptr(c, l, FIGHTERS)[0] = 0;
  }
  for (int s = 0; s < NS; s++) {
    uint8_t space = (uint8_t)(c->carriers[s] * 2);
    for (int p = 1; p < NP; p++) space = (uint8_t)(space - tpsut(c, p, s)[FIGHTERS]);
    uint8_t* n = &sea_ptr(c, s, FIGHTERS)[0];
    if (space < *n) {
      uint8_t lost = (uint8_t)(*n - space);
      c->tot_sea[0][s] -= lost;
      c->cur_sea[s][FIGHTERS] = (uint8_t)(c->cur_sea[s][FIGHTERS] - lost);
      *n = space;
    }
  }
}
static void reset_units_fully(Ctx* c) {
  for (int s = 0; s < NS; s++) {
    uint8_t dmg = c->cur_sea[s][BS_DAMAGED];
    sea_ptr(c, s, BATTLESHIPS)[0] = (uint8_t)(sea_ptr(c, s, BATTLESHIPS)[0] + dmg);
    c->cur_sea[s][BATTLESHIPS] = (uint8_t)(c->cur_sea[s][BATTLESHIPS] + dmg);
    sea_ptr(c, s, BS_DAMAGED)[0] = 0;
    sea_ptr(c, s, BS_DAMAGED)[1] = 0;
    c->cur_sea[s][BS_DAMAGED] = 0;
  }
}
static void collect_money(Ctx* c) {
  int holds_capital = OWNER(c, T.capital[c->st.player_index]) == 0;
  c->st.money[0] = (uint8_t)(c->st.money[0] + c->income[0] * holds_capital);
}

/* ---- phase 20: rotate_turns, 3895-4029 ---- */
static void rotate_turns(Ctx* c) {
  CaaaGameState* st = &c->st;
  uint8_t tmp_land[NL][NLT], tmp_sea[NS][NST];
  memcpy(tmp_land, c->cur_land, 30);
  memcpy(c->cur_land, st->other_land_units[0], 30);
  memmove(st->other_land_units[0], st->other_land_units[1], 90);
  memcpy(st->other_land_units[3], tmp_land, 30);
  for (int l = 0; l < NL; l++) {
    uint8_t* base = (uint8_t*)&st->land_state[l];
    memset(base + 4, 0, 21);
    for (int u = 0; u < NLT; u++) base[OFF_LAND[u] + STATES_LAND[u] - 1] = c->cur_land[l][u];
  }
  memcpy(tmp_sea, c->cur_sea, 45);
  for (int s = 0; s < NS; s++) memcpy(c->cur_sea[s], st->other_sea_units[0][s], 14); /* 3944: bombers stay stale */
  memmove(st->other_sea_units[0], st->other_sea_units[1], 126);
  memset(st->units_sea, 0, sizeof(st->units_sea));
  for (int s = 0; s < NS; s++) {
    memcpy(st->other_sea_units[3][s], tmp_sea[s], 14);
    uint8_t* base = (uint8_t*)&st->units_sea[s];
    for (int u = 0; u < NST; u++) base[OFF_SEA[u] + STATES_SEA[u] - 1] = c->cur_sea[s][u];
  }
  uint8_t m = st->money[0];
  memmove(&st->money[0], &st->money[1], 4);
  st->money[4] = m;
  c->income[5] = c->income[0];
  memmove(&c->income[0], &c->income[1], sizeof(int) * 5);
  c->nfact[5] = c->nfact[0];
  memmove(&c->nfact[0], &c->nfact[1], sizeof(int) * 5);
  memcpy(c->factloc[5], c->factloc[0], sizeof(c->factloc[0]));
  memmove(c->factloc[0], c->factloc[1], sizeof(c->factloc[0]) * 5);
  memcpy(c->tot_land[5], c->tot_land[0], sizeof(c->tot_land[0]));
  memmove(c->tot_land[0], c->tot_land[1], sizeof(c->tot_land[0]) * 5);
  memcpy(c->tot_sea[5], c->tot_sea[0], sizeof(c->tot_sea[0]));
  memmove(c->tot_sea[0], c->tot_sea[1], sizeof(c->tot_sea[0]) * 5);
  memset(st->flagged_for_combat, 0, NA);
  st->player_index = (uint8_t)((st->player_index + 1) % NP);
  for (int l = 0; l < NL; l++) OWNER(c, l) = (uint8_t)((OWNER(c, l) + NP - 1) % NP);
  for (int f = 0; f < c->nfact[0]; f++) {
    int land = c->factloc[0][f];
    st->builds_left[land] = FMAX(c, land);
    for (int i = 0; i < T.l2s_count[land]; i++)
      st->builds_left[T.l2s_conn[land][i] + NL] = (uint8_t)(st->builds_left[T.l2s_conn[land][i] + NL] + FMAX(c, land));
  }
  refresh_eot_cache(c);
}

/* ---- scoring, 4301-4337 (reads the engine's own state, not its argument, for unit totals) ---- */
static double evaluate_state(Ctx* c) {
  uint8_t starting = c->st.player_index;
  if (starting >= NP) c->st.player_index = (uint8_t)(starting - NP);
  int pi = c->st.player_index;
  int allied = 1, enemy = 1;
  allied += c->st.money[0];
  const uint8_t* other_sea_flat = &c->st.other_sea_units[0][0][0];
  for (int p = 0; p < NP; p++) {
    int score = c->st.money[p];
    for (int l = 0; l < NL; l++)
      for (int u = 0; u < NLT; u++) score += tplut(c, p, l)[u] * COST_LAND[u];
    for (int s = 0; s < NS; s++)
      for (int u = 0; u < NST; u++) { /* 4319: 15 entries of a 14-entry row -> runs into the next row */
        int n = p == 0 ? c->cur_sea[s][u] : other_sea_flat[((p - 1) * NS + s) * 14 + u];
        score += n * COST_SEA[u];
      }
    if (T.is_allied[pi][(pi + p) % NP]) allied += score;
    else enemy += score;
  }
  double score = (double)allied / (double)(enemy + allied);
  c->st.player_index = starting;
  if (starting >= NP) return 1 - score;
  return score;
}

static int run_phase(Ctx* c, int ph) {
  switch (ph) {
  case CAAA_PH_MOVE_FIGHTERS: return move_fighter_units(c);
  case CAAA_PH_MOVE_BOMBERS: return move_bomber_units(c);
  case CAAA_PH_STAGE_TRANSPORTS: return stage_transport_units(c);
  case CAAA_PH_MOVE_TANKS: return move_land_unit_type(c, TANKS);
  case CAAA_PH_MOVE_ARTILLERY: return move_land_unit_type(c, ARTILLERY);
  case CAAA_PH_MOVE_INFANTRY: return move_land_unit_type(c, INFANTRY);
  case CAAA_PH_MOVE_TRANSPORTS: return move_transport_units(c);
  case CAAA_PH_MOVE_SUBS: return move_subs(c);
  case CAAA_PH_MOVE_SHIPS: return move_destroyers_battleships(c);
  case CAAA_PH_SEA_BATTLES: return resolve_sea_battles(c);
  case CAAA_PH_UNLOAD_TRANSPORTS: return unload_transports(c);
  case CAAA_PH_LAND_BATTLES: return resolve_land_battles(c);
  case CAAA_PH_MOVE_AA_GUNS: return move_land_unit_type(c, AA_GUNS);
  case CAAA_PH_LAND_FIGHTERS: return land_fighter_units(c);
  case CAAA_PH_LAND_BOMBERS: return land_bomber_units(c);
  case CAAA_PH_BUY_UNITS: return buy_units(c);
  case CAAA_PH_CRASH_AIR: crash_air_units(c); return 0;
  case CAAA_PH_RESET_UNITS: reset_units_fully(c); return 0;
  case CAAA_PH_BUY_FACTORY: return 0; /* buy_factory is empty, 3889 */
  case CAAA_PH_COLLECT_MONEY: collect_money(c); return 0;
  case CAAA_PH_ROTATE_TURNS: rotate_turns(c); return 0;
  default: return -1;
  }
}

/* random_play_until_terminal, 4184-4242 */
static double random_play(Ctx* c, const void* state670) {
  memcpy(&c->st, state670, CAAA_STATE_BYTES);
  refresh_full_cache(c);
  c->answers_remaining = 100000;
  c->use_selected = 0;
  double score = evaluate_state(c);
  c->max_loops = 1000;
  c->draws = 0; /* P2: stream restarts at (seed, game_id, 0) */
  if (c->auto_game_id) c->game_id = c->playout_counter;
  c->playout_counter++;
  if (c->reset_vm) { /* P1(iv) */
    memset(c->vm, 0, sizeof(c->vm));
    c->nvm = 0;
  }
  while (score > 0.01 && score < 0.99 && c->max_loops-- > 0) {
    for (int ph = 0; ph < CAAA_PH_COUNT; ph++) run_phase(c, ph);
    score = evaluate_state(c);
  }
  if (c->st.player_index % 2 == 1) score = 1 - score;
  return score;
}

/* get_possible_actions / apply_action, 4050-4183: same pipeline, stop at a decision */
static void run_until_decision(Ctx* c) {
  for (;;) {
    int stopped = 0;
    for (int ph = 0; ph <= CAAA_PH_BUY_UNITS; ph++)
      if (run_phase(c, ph)) { stopped = 1; break; }
    if (stopped) break;
    for (int ph = CAAA_PH_CRASH_AIR; ph < CAAA_PH_COUNT; ph++) run_phase(c, ph);
  }
}

static uint64_t fnv1a(const uint8_t* p, int n) {
  uint64_t h = CAAA_FNV_OFFSET;
  for (int i = 0; i < n; i++) { h ^= p[i]; h *= CAAA_FNV_PRIME; }
  return h;
}
static void fill_stats(Ctx* c, double r, CaaaRolloutStats* st) {
  if (!st) return;
  st->result = r;
  st->turns = c->max_loops < 0 ? 1000 : 1000 - c->max_loops;
  st->decisions = 100000 - c->answers_remaining;
  st->draws = (int32_t)c->draws;
  st->final_player = c->st.player_index;
  st->state_hash = fnv1a((const uint8_t*)&c->st, CAAA_STATE_LIVE_BYTES);
}

/* ---- C API (mirrors oracle/ref_shim.cpp) ---- */
void caaa_oracle_init(void) {
  if (g_inited) return;
  caaa_oracle_build_tables(&T);
  memset(&G, 0, sizeof(G));
  G.auto_game_id = 1;
  G.reset_vm = 1;
  g_inited = 1;
}
void caaa_oracle_set_stream(int mode, uint64_t seed) { (void)mode; G.seed = seed; }
void caaa_oracle_set_reset_valid_moves(int on) { G.reset_vm = on != 0; }
void caaa_oracle_set_game_id(int64_t game_id) {
  if (game_id < 0) G.auto_game_id = 1;
  else { G.auto_game_id = 0; G.game_id = (uint64_t)game_id; }
}
double caaa_oracle_rollout(const void* state670, CaaaRolloutStats* st) {
  double r = random_play(&G, state670);
  fill_stats(&G, r, st);
  return r;
}
void caaa_oracle_rollout_many(const void* state670, int64_t first_game_id, int n, CaaaRolloutStats* st, double* results) {
  for (int i = 0; i < n; i++) {
    if (first_game_id >= 0) { G.auto_game_id = 0; G.game_id = (uint64_t)first_game_id + (uint64_t)i; }
    double r = random_play(&G, state670);
    if (results) results[i] = r;
    if (st) fill_stats(&G, r, st + i);
  }
}
void caaa_oracle_get_state(void* out670) { memcpy(out670, &G.st, CAAA_STATE_BYTES); }
int caaa_oracle_get_possible_actions(const void* state670, uint8_t* actions20) {
  Ctx* c = &G;
  c->unlucky = 0;
  memcpy(&c->st, state670, CAAA_STATE_BYTES);
  refresh_full_cache(c);
  c->answers_remaining = 0;
  run_until_decision(c);
  memset(actions20, 0, CAAA_ACTIONS);
  memcpy(actions20, c->vm, c->nvm < CAAA_ACTIONS ? c->nvm : CAAA_ACTIONS);
  return c->nvm;
}
void caaa_oracle_apply_action(void* state670, uint8_t action) {
  Ctx* c = &G;
  memcpy(&c->st, state670, CAAA_STATE_BYTES);
  refresh_full_cache(c);
  c->unlucky = 0;
  c->answers_remaining = 1;
  c->selected_action = action;
  c->use_selected = 1;
  run_until_decision(c);
  memcpy(state670, &c->st, CAAA_STATE_BYTES);
}
int caaa_oracle_is_terminal(const void* state670) { /* 4244-4248; see note in DESIGN.md on the stale-global quirk */
  Ctx* c = &G;
  uint8_t saved_pi = c->st.player_index;
  uint8_t saved_money[NP];
  memcpy(saved_money, c->st.money, NP);
  const CaaaGameState* in = (const CaaaGameState*)state670;
  c->st.player_index = in->player_index;
  memcpy(c->st.money, in->money, NP);
  double s = evaluate_state(c);
  c->st.player_index = saved_pi;
  memcpy(c->st.money, saved_money, NP);
  return s > 0.99 || s < 0.01;
}
void caaa_oracle_load_state(const void* state670, int64_t game_id) {
  Ctx* c = &G;
  memcpy(&c->st, state670, CAAA_STATE_BYTES);
  refresh_full_cache(c);
  c->answers_remaining = 100000;
  c->use_selected = 0;
  c->max_loops = 1000;
  c->auto_game_id = 0;
  c->game_id = (uint64_t)game_id;
  c->draws = 0;
  memset(c->vm, 0, sizeof(c->vm));
  c->nvm = 0;
}
int caaa_oracle_run_phase(int phase) { return run_phase(&G, phase); }
double caaa_oracle_evaluate(void) { return evaluate_state(&G); }
int caaa_oracle_draws(void) { return (int)G.draws; }
void caaa_oracle_get_cache(CaaaCacheDump* d) {
  Ctx* c = &G;
  memset(d, 0, sizeof(*d));
  memcpy(d->income_per_turn, c->income, sizeof(c->income));
  memcpy(d->total_factory_count, c->nfact, sizeof(c->nfact));
  memcpy(d->factory_locations, c->factloc, sizeof(c->factloc));
  memcpy(d->total_player_land_units, c->tot_land, sizeof(c->tot_land));
  memcpy(d->total_player_sea_units, c->tot_sea, sizeof(c->tot_sea));
  memcpy(d->transports_with_large_cargo_space, c->large_cargo, sizeof(c->large_cargo));
  memcpy(d->transports_with_small_cargo_space, c->small_cargo, sizeof(c->small_cargo));
  memcpy(d->allied_carriers, c->carriers, sizeof(c->carriers));
  memcpy(d->enemy_units_count, c->enemy_count, sizeof(c->enemy_count));
  d->answers_remaining = c->answers_remaining;
  memcpy(d->current_player_land_unit_types, c->cur_land, 30);
  memcpy(d->current_player_sea_unit_types, c->cur_sea, 45);
  memcpy(d->enemy_blockade_total, c->blockade, 3);
  memcpy(d->enemy_destroyers_total, c->edestroyers, 3);
  memcpy(d->is_land_path_blocked, c->land_blocked, 25);
  memcpy(d->is_sea_path_blocked, c->sea_blocked, 9);
  memcpy(d->is_sub_path_blocked, c->sub_blocked, 9);
  memcpy(d->enemies_0, c->enemies, 4);
  memcpy(d->is_allied_0, c->allied, 5);
  d->enemies_count_0 = c->n_enemies;
  d->canal_state = c->canal_state;
  memcpy(d->valid_moves, c->vm, CAAA_ACTIONS);
  d->valid_moves_count = c->nvm;
}
void caaa_oracle_get_tables(CaaaTables* t) { *t = T; }
int caaa_oracle_stuck(void) { /* returns and clears the sea-battle guard flag */
  int s = G.stuck;
  G.stuck = 0;
  return s;
}
void caaa_oracle_philox(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) { caaa_philox4x32_10_ref(ctr, key, out); }
uint8_t caaa_oracle_stream_byte(uint64_t seed, uint64_t game_id, uint32_t draw_index) {
  return caaa_philox_byte_ref(seed, game_id, draw_index);
}
