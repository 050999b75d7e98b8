// engine.cuh -- the rules engine that one GPU thread runs on one game slot in shared memory.
//
// B200-native restatement of the reference's turn pipeline (reference src/engine.cpp:642-814 and
// 1291-4337): 21 phases per player-turn, their derived caches, volley-based combat and scoring.
// All parallelism is across games (draw order inside a game is sequential by definition), so the
// code is written for one thread per slot: byte/word accesses into a bank-conflict-free slot
// (slot.cuh), topology tables read from a per-CTA shared-memory copy, unit tables as immediates,
// landing masks as register bitmasks instead of the reference's global bool arrays.
//
// Reference quirks are reproduced deliberately; layout-dependent
// out-of-bounds reads follow the canonical meaning of patch set P1.
//
// The functions are __host__ __device__ so that tests/hostsim can run the *same* code on the CPU
// against the oracle when no GPU is present; the product only ever launches them in kernels.
#pragma once
#include "philox.cuh"
#include "slot.cuh"

namespace caaa {

struct RunCtx {
  const DevTables* T;
  uint32_t seed_lo, seed_hi;
};

// ---- views into the slot that mirror the reference's pointer tables (engine.cpp:521-568) ----
CAAA_HD uint8_t* land_ptr(GameSlot* g, int land, int type) { return (uint8_t*)&g->st.land_state[land] + off_land(type); }
CAAA_HD uint8_t* sea_ptr(GameSlot* g, int sea, int type) { return (uint8_t*)&g->st.units_sea[sea] + off_sea(type); }
CAAA_HD uint8_t* air_fighters(GameSlot* g, int air) { return air < NL ? land_ptr(g, air, FIGHTERS) : sea_ptr(g, air - NL, FIGHTERS); }
CAAA_HD uint8_t* air_bombers(GameSlot* g, int air) { return air < NL ? land_ptr(g, air, BOMBERS_LAND_AIR) : sea_ptr(g, air - NL, BOMBERS_SEA); }
CAAA_HD uint8_t* tplut(GameSlot* g, int player, int land) { return player == 0 ? g->cur_land[land] : g->st.other_land_units[player - 1][land]; }
CAAA_HD uint8_t* tpsut(GameSlot* g, int player, int sea) { return player == 0 ? g->cur_sea[sea] : g->st.other_sea_units[player - 1][sea]; }
CAAA_HD int pidx(const GameSlot* g) { return g->st.player_index % NP; }
CAAA_HD bool allied(const GameSlot* g, const RunCtx& cx, int p) { return (cx.T->allied_mask[pidx(g)] >> p) & 1; }
CAAA_HD bool owner_allied(const GameSlot* g, const RunCtx& cx, int land) { return allied(g, cx, g->st.land_state[land].owner_idx); }

// skipped_moves as a flat array of 64 one-byte flags (P1 ii): out-of-range reads 0, writes dropped
CAAA_HD bool skip_get(const GameSlot* g, unsigned a, unsigned b) {
  const unsigned i = a * 8u + b;
  return i < 64u ? (((const uint8_t*)g->st.skipped_moves)[i] & 1) != 0 : false;
}
CAAA_HD void skip_set(GameSlot* g, unsigned a, unsigned b) {
  const unsigned i = a * 8u + b;
  if (i < 64u) ((uint8_t*)g->st.skipped_moves)[i] |= 1;
}
CAAA_HD void clear_move_history(GameSlot* g) {  // engine.cpp:1330
  uint16_t* p = (uint16_t*)g->st.skipped_moves;   // offset 606: 2-byte aligned
#pragma unroll
  for (int i = 0; i < 32; i++) p[i] = 0;
}

CAAA_HD_NOINLINE uint8_t next_byte(GameSlot* g, const RunCtx& cx) {
  const uint32_t d = g->draws++;
  const uint32_t k = d & 15u;
  if (k == 0) philox4x32_10(d >> 4, g->game_lo, g->game_hi, 0u, cx.seed_lo, cx.seed_hi, g->rng);
  return (uint8_t)(g->rng[k >> 2] >> (8u * (k & 3u)));
}

// ---- cache refresh, engine.cpp:642-814 ----
CAAA_HD void refresh_canals(GameSlot* g, const RunCtx& cx) {  // 726-733 (table copies are read via canal_state)
  uint32_t cs = 0;
#pragma unroll
  for (int k = 0; k < CAAA_CANALS; k++)
    if (owner_allied(g, cx, cx.T->t.canal_lands[k][0]) && owner_allied(g, cx, cx.T->t.canal_lands[k][1])) cs += 1u << k;
  g->canal_state = (uint8_t)cs;
}
CAAA_HD void refresh_enemy_armies_and_fleets(GameSlot* g, const RunCtx& cx) {  // 752-784
  const int pi = pidx(g);
  const int ne = cx.T->n_enemies[pi];
  for (int l = 0; l < NL; l++) {
    int n = 0;
    for (int e = 0; e < ne; e++) n += g->tot_land[cx.T->enemies[pi][e]][l];
    g->enemy_count[l] = n;
  }
  for (int s = 0; s < NS; s++) {
    const uint8_t* mine = g->cur_sea[s];
    int carriers = mine[CARRIERS], ecount = 0;
    uint32_t edest = 0, block = 0;
    for (int p = 1; p < NP; p++) {
      const uint8_t* su = tpsut(g, p, s);
      if (allied(g, cx, p)) {
        carriers += su[CARRIERS];
      } else {
        ecount += g->tot_sea[p][s];
        edest += su[DESTROYERS];
        block += su[DESTROYERS] + su[CARRIERS] + su[CRUISERS] + su[BATTLESHIPS] + su[BS_DAMAGED];
      }
    }
    g->carriers[s] = carriers;
    g->enemy_count[s + NL] = ecount;
    g->edestroyers[s] = (uint8_t)edest;
    g->blockade[s] = (uint8_t)block;
    g->large_cargo[s] = mine[TRANS_EMPTY] + mine[TRANS_1I];
    g->small_cargo[s] = mine[TRANS_EMPTY] + mine[TRANS_1I] + mine[TRANS_1A] + mine[TRANS_1T];
  }
}
CAAA_HD void refresh_paths_blocked(GameSlot* g, const RunCtx& cx) {  // 785-814
  const CaaaTables& t = cx.T->t;
  for (int s = 0; s < NL; s++)
    for (int i = 0; i < t.lands_within_2_count[s]; i++) {
      const int d = t.lands_within_2[s][i];
      const int n1 = t.land_path[s][d], n2 = t.land_path_alt[s][d];
      g->land_blocked[s * NL + d] = (g->enemy_count[n1] > 0 || g->st.land_state[n1].factory_max > 0) &&
                                    (g->enemy_count[n2] > 0 || g->st.land_state[n2].factory_max > 0);
    }
  const int cs = g->canal_state;
  for (int s = 0; s < NS; s++)
    for (int i = 0; i < t.seas_within_2_count[cs][s]; i++) {
      const int d = t.seas_within_2[cs][s][i];
      const int n1 = t.sea_path[cs][s][d], n2 = t.sea_path_alt[cs][s][d];
      g->sea_blocked[s * NS + d] = g->blockade[n1] > 0 && g->blockade[n2] > 0;
      g->sub_blocked[s * NS + d] = g->edestroyers[n1] > 0 && g->edestroyers[n2] > 0;
    }
}
CAAA_HD_NOINLINE void refresh_eot_cache(GameSlot* g, const RunCtx& cx) {  // 653-660 (refresh_allies is tabulated)
  refresh_canals(g, cx);
  refresh_enemy_armies_and_fleets(g, cx);
  refresh_paths_blocked(g, cx);
}
CAAA_HD_NOINLINE void refresh_full_cache(GameSlot* g, const RunCtx& cx) {  // 642-652, 661-715
  for (int p = 0; p < NP; p++) {
    g->income[p] = 0;
    g->nfact[p] = 0;
  }
  for (int p = 0; p < 6; p++) {
    for (int l = 0; l < NL; l++) g->tot_land[p][l] = 0;
    for (int s = 0; s < NS; s++) g->tot_sea[p][s] = 0;
  }
  for (int l = 0; l < NL; l++) {
    const int owner = g->st.land_state[l].owner_idx;
    if (g->st.land_state[l].factory_max > 0) g->factloc[owner][g->nfact[owner]++] = (uint8_t)l;
    g->income[owner] += cx.T->t.land_value[l];
    for (int u = 0; u < NLT; u++) {
      const uint8_t* s = land_ptr(g, l, u);
      uint32_t n = 0;
      for (int k = 0; k < (int)states_land(u); k++) n += s[k];
      g->cur_land[l][u] = (uint8_t)n;
    }
    for (int p = 0; p < NP; p++) {
      const uint8_t* r = tplut(g, p, l);
      int n = 0;
      for (int u = 0; u < NLT; u++) n += r[u];
      g->tot_land[p][l] = n;
    }
  }
  for (int s = 0; s < NS; s++) {
    for (int u = 0; u < NST; u++) {
      const uint8_t* q = sea_ptr(g, s, u);
      uint32_t n = 0;
      for (int k = 0; k < (int)states_sea(u); k++) n += q[k];
      g->cur_sea[s][u] = (uint8_t)n;
    }
    for (int p = 0; p < NP; p++) {
      const uint8_t* r = tpsut(g, p, s);
      int n = 0;
      for (int u = 0; u < NST - 1; u++) n += r[u];
      g->tot_sea[p][s] = n;
    }
  }
  refresh_eot_cache(g, cx);
}

// ---- decisions and dice ----
template <bool EXPAND>
CAAA_HD uint8_t ai_input(GameSlot* g, const RunCtx& cx) {  // getAIInput 1291-1297
  g->answers_remaining--;
  if (EXPAND && g->use_selected) return g->selected_action;
  return g->vm[next_byte(g, cx) % g->nvm];
}
CAAA_HD bool unlucky_team(const GameSlot* g, const RunCtx& cx) { return cx.T->t.is_allied[pidx(g)][g->unlucky % NP] != 0; }
CAAA_HD uint8_t attacker_hits(GameSlot* g, const RunCtx& cx, int dmg) {  // 2552-2565
  if (g->answers_remaining < 2) return (uint8_t)(unlucky_team(g, cx) ? dmg / 6 : dmg / 6 + (1 < dmg % 6 ? 1 : 0));
  return (uint8_t)(dmg / 6 + ((int)(next_byte(g, cx) % 6) < dmg % 6 ? 1 : 0));
}
CAAA_HD uint8_t defender_hits(GameSlot* g, const RunCtx& cx, int dmg) {  // 2567-2580
  if (g->answers_remaining < 2) return (uint8_t)(unlucky_team(g, cx) ? dmg / 6 + (1 < dmg % 6 ? 1 : 0) : dmg / 6);
  return (uint8_t)(dmg / 6 + ((int)(next_byte(g, cx) % 6) < dmg % 6 ? 1 : 0));
}

CAAA_HD_NOINLINE void update_move_history_4air(GameSlot* g, uint8_t src, uint8_t dst) {  // 1300-1328: starts one past the list
  for (;;) {
    const uint8_t va = g->nvm < CAAA_ACTIONS ? g->vm[g->nvm] : (uint8_t)0;
    if (va == dst) break;
    skip_set(g, src, va);
    for (unsigned i = 0; i < (unsigned)NA; i++)  // apply_skip
      if (skip_get(g, va, i)) skip_set(g, src, i);
    g->nvm--;
  }
}

CAAA_HD_NOINLINE bool load_transport(GameSlot* g, int ut, int src_land, int dst_sea, int land_state) {  // 1397-1443
  for (int tt = weight_land(ut) > 2 ? TRANS_1I : TRANS_1T; tt >= TRANS_EMPTY; tt--) {
    uint8_t* arr = sea_ptr(g, dst_sea, tt);
    const int unl = unloading_sea(tt);
    for (int ts = (int)states_sea(tt) - (int)staging_sea(tt); ts >= unl; ts--) {
      if (arr[ts] == 0) continue;
      const int ntt = load_result(ut, tt);
      arr[ts]--;
      if (tt == TRANS_EMPTY && ts == 0) ts = unloading_sea(ntt);
      sea_ptr(g, dst_sea, ntt)[ts]++;
      g->cur_sea[dst_sea][tt]--;
      g->cur_sea[dst_sea][ntt]++;
      g->tot_land[0][src_land]--;
      g->cur_land[src_land][ut]--;
      land_ptr(g, src_land, ut)[land_state]--;
      const uint8_t* mine = g->cur_sea[dst_sea];
      g->large_cargo[dst_sea] = mine[TRANS_EMPTY] + mine[TRANS_1I];
      g->small_cargo[dst_sea] = g->large_cargo[dst_sea] + mine[TRANS_1A] + mine[TRANS_1T];
      return true;
    }
  }
  return false;
}

CAAA_HD_NOINLINE void add_valid_land_moves(GameSlot* g, const RunCtx& cx, int src, int moves, int ut) {  // 1445-1516
  const CaaaTables& t = cx.T->t;
  uint32_t n = g->nvm;
  if (moves == 2) {
    for (int i = 0; i < t.lands_within_2_count[src]; i++) {
      const int d = t.lands_within_2[src][i];
      if (skip_get(g, src, d)) continue;
      if (t.land_dist[src][d] == 2 && g->land_blocked[src * NL + d]) continue;
      g->vm[n++] = (uint8_t)d;
    }
    for (int i = 0; i < t.load_within_2_count[src]; i++) {
      const int ds = t.load_within_2[src][i];
      if (g->large_cargo[ds] == 0) continue;
      const int da = ds + NL;
      if (skip_get(g, src, da)) continue;
      const unsigned flat = (unsigned)(src * NL + da);  // 1474: air index into a 5-wide row (P1 iii)
      if (t.land_dist[src][da] == 1 && (flat < 25u ? g->land_blocked[flat] : 0)) continue;
      g->vm[n++] = (uint8_t)da;
    }
    g->nvm = (uint8_t)n;
    return;
  }
  const bool non_combat = ut == AA_GUNS, unloadable = weight_land(ut) > 5, heavy = weight_land(ut) > 2;
  for (int i = 0; i < t.l2l_count[src]; i++) {
    const int d = t.l2l_conn[src][i];
    if (skip_get(g, src, d)) continue;
    if (non_combat && !owner_allied(g, cx, d)) continue;
    g->vm[n++] = (uint8_t)d;
  }
  if (!unloadable)
    for (int i = 0; i < t.l2s_count[src]; i++) {
      const int ds = t.l2s_conn[src][i];
      if (g->small_cargo[ds] == 0) continue;
      if (heavy && g->large_cargo[ds] == 0) continue;
      const int da = ds + NL;
      if (skip_get(g, src, da)) continue;
      g->vm[n++] = (uint8_t)da;
    }
  g->nvm = (uint8_t)n;
}

CAAA_HD_NOINLINE void add_valid_sea_moves(GameSlot* g, const RunCtx& cx, int src_sea, int moves, const uint8_t* blocked) {  // 1518-1586
  const CaaaTables& t = cx.T->t;
  const int cs = g->canal_state, sa = src_sea + NL;
  uint32_t n = g->nvm;
  if (moves == 2) {
    for (int i = 0; i < t.seas_within_2_count[cs][src_sea]; i++) {
      const int d = t.seas_within_2[cs][src_sea][i];
      if (blocked[src_sea * NS + d]) continue;
      if (skip_get(g, sa, d + NL)) continue;
      g->vm[n++] = (uint8_t)(d + NL);
    }
  } else {
    for (int i = 0; i < t.seas_within_1_count[cs][src_sea]; i++) {
      const int d = t.seas_within_1[cs][src_sea][i];
      if (skip_get(g, sa, d + NL)) continue;
      g->vm[n++] = (uint8_t)(d + NL);
    }
  }
  g->nvm = (uint8_t)n;
}

// ---- conquer_land, 1918-1984 ----
CAAA_HD_NOINLINE void conquer_land(GameSlot* g, const RunCtx& cx, int dst) {
  const CaaaTables& t = cx.T->t;
  const int pi = pidx(g);
  const int old = g->st.land_state[dst].owner_idx;
  if (t.capital[(pi + old) % NP] == dst) {
    g->st.money[0] = (uint8_t)(g->st.money[0] + g->st.money[old]);
    g->st.money[old] = 0;
  }
  g->income[old] -= t.land_value[dst];
  int nw = 0;
  const int orig = (t.original_owner[dst] + NP - pi) % NP;
  if (allied(g, cx, orig)) nw = orig;
  g->st.land_state[dst].owner_idx = (uint8_t)nw;
  g->income[nw] += t.land_value[dst];
  g->factloc[nw][g->nfact[nw]++] = (uint8_t)dst;
  g->nfact[old]--;
  const int n_old = g->nfact[old];
  for (int i = 0; i < n_old; i++)
    if (g->factloc[old][i] == dst) {
      for (int j = i; j < n_old; j++) g->factloc[old][j] = g->factloc[old][j + 1];
      break;
    }
}

// ---- phase 0: move_fighter_units, 1657-1803, 3385-3407 ----
// returns bit masks (bit a = air location a): lo byte canFighterLandHere, next byte canFighterLandIn1Move
CAAA_HD_NOINLINE uint32_t fighter_landing_masks(GameSlot* g, const RunCtx& cx) {  // pre_move_fighter_units 1657-1714
  const CaaaTables& t = cx.T->t;
  uint32_t here = 0;
  for (int l = 0; l < NL; l++) {
    const int owner = g->st.land_state[l].owner_idx;
    if (allied(g, cx, owner) && g->st.flagged_for_combat[l] == 0) here |= 1u << l;
    if (owner == g->st.player_index && g->st.land_state[l].factory_max > 0)  // 1676: relative owner vs absolute player
      for (int i = 0; i < t.l2s_count[l]; i++) here |= 1u << (NL + t.l2s_conn[l][i]);
  }
  for (int s = 0; s < NS; s++) {
    if (g->carriers[s] <= 0) continue;
    here |= 1u << (s + NL);
    if (sea_ptr(g, s, CARRIERS)[2] > 0)  // 1688: carriers[2] aliases cruisers[0]
      for (int i = 0; i < t.s2s_count[s]; i++) {
        const int s1 = t.s2s_conn[s][i];
        here |= 1u << (NL + s1);
        for (int j = 0; j < t.s2s_count[s1]; j++) here |= 1u << (NL + t.s2s_conn[s1][j]);
      }
  }
  uint32_t in1 = 0;
  for (int a = 0; a < NA; a++)
    for (int i = 0; i < t.air_conn_count[a]; i++)
      if ((here >> t.air_conn[a][i]) & 1) {
        in1 |= 1u << a;
        break;
      }
  return here | (in1 << 8);
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_fighter_units(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  bool processed = false;
  uint32_t masks = 0;
  for (int src = 0; src < NA; src++) {
    uint8_t* arr = air_fighters(g, src);
    if (arr[4] == 0) continue;
    if (!processed) {
      processed = true;
      masks = fighter_landing_masks(g, cx);
    }
    g->vm[0] = (uint8_t)src;
    uint32_t n = 1;
    for (int i = 0; i < t.air_within_count[3][src]; i++) {  // add_valid_fighter_moves(src, 4)
      const int d = t.air_within[3][src][i];
      const int dist = t.air_dist[src][d];
      const bool here = (masks >> d) & 1;
      if (dist <= 2 || here || (dist == 3 && ((masks >> (8 + d)) & 1))) {
        if (!here && g->enemy_count[d] == 0) continue;  // waste of a move
        if (!skip_get(g, src, d)) g->vm[n++] = (uint8_t)d;
      }
    }
    g->nvm = (uint8_t)n;
    while (arr[4] > 0) {
      uint8_t dst = g->vm[0];
      if (g->nvm > 1) {
        if (g->answers_remaining == 0) return true;
        dst = ai_input<EXPAND>(g, cx);
      }
      update_move_history_4air(g, (uint8_t)src, dst);
      if (src == dst) {
        arr[0] = (uint8_t)(arr[0] + arr[4]);
        arr[4] = 0;
        continue;
      }
      int dist = t.air_dist[src][dst];
      if (dst < NL) {
        if (!owner_allied(g, cx, dst)) g->st.flagged_for_combat[dst] = 2;
        else dist = 4;  // friendly rebase uses up all moves
        land_ptr(g, dst, FIGHTERS)[(uint8_t)(4 - dist)]++;
        g->cur_land[dst][FIGHTERS]++;
        g->tot_land[0][dst]++;
      } else {
        g->st.flagged_for_combat[dst] = g->enemy_count[dst] > 0;
        sea_ptr(g, dst - NL, FIGHTERS)[(uint8_t)(4 - dist)]++;
        g->cur_sea[dst - NL][FIGHTERS]++;
        g->tot_sea[0][dst - NL]++;
      }
      if (src < NL) {
        g->cur_land[src][FIGHTERS]--;
        g->tot_land[0][src]--;
      } else {
        g->cur_sea[src - NL][FIGHTERS]--;
        g->tot_sea[0][src - NL]--;
      }
      arr[4]--;
    }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 1: move_bomber_units, 1805-1916, 3423-3443 ----
CAAA_HD uint32_t bomber_here_mask(const GameSlot* g, const RunCtx& cx) {  // 1808-1812, 3557-3563 (lands only)
  uint32_t here = 0;
  for (int l = 0; l < NL; l++)
    if (owner_allied(g, cx, l) && g->st.flagged_for_combat[l] == 0) here |= 1u << l;
  return here;
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_bomber_units(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  // has-work shortcut: the landing masks are phase-local in this implementation (the reference keeps
  // them in globals, but re-derives them before every use), so nothing happens without unmoved bombers
  bool any = false;
  for (int l = 0; l < NL; l++) any |= land_ptr(g, l, BOMBERS_LAND_AIR)[6] != 0;
  if (!any) return false;
  const uint32_t here = bomber_here_mask(g, cx);
  uint32_t in1 = 0, in2 = 0;
  for (int l = 0; l < NL; l++)
    for (int i = 0; i < t.l2l_count[l]; i++)
      if ((here >> t.l2l_conn[l][i]) & 1) { in1 |= 1u << l; break; }
  for (int s = 0; s < NS; s++)
    for (int i = 0; i < t.s2l_count[s]; i++)
      if ((here >> t.s2l_conn[s][i]) & 1) { in1 |= 1u << (NL + s); break; }
  for (int a = 0; a < NA; a++)
    for (int i = 0; i < t.air_conn_count[a]; i++)
      if ((in1 >> t.air_conn[a][i]) & 1) { in2 |= 1u << a; break; }
  for (int src = 0; src < NL; src++) {
    uint8_t* arr = land_ptr(g, src, BOMBERS_LAND_AIR);
    if (arr[6] == 0) continue;
    g->vm[0] = (uint8_t)src;
    uint32_t n = 1;
    for (int i = 0; i < t.air_within_count[5][src]; i++) {  // add_valid_bomber_moves(src, 6)
      const int d = t.air_within[5][src][i];
      const int dist = t.air_dist[src][d];
      const bool h = (here >> d) & 1;
      if (dist <= 3 || h || (dist == 4 && ((in2 >> d) & 1)) || (dist == 5 && ((in1 >> d) & 1))) {
        if (!h && g->enemy_count[d] == 0) {
          if (d >= NL || g->st.land_state[d].factory_max == 0 ||
              g->st.land_state[d].factory_hp == -(int)g->st.land_state[d].factory_max)
            continue;
        }
        if (!skip_get(g, src, d)) g->vm[n++] = (uint8_t)d;
      }
    }
    g->nvm = (uint8_t)n;
    while (arr[6] > 0) {
      uint8_t dst = g->vm[0];
      if (g->nvm == 1) {  // 1860: inverted test
        if (g->answers_remaining == 0) return true;
        dst = ai_input<EXPAND>(g, cx);
      }
      if (src == dst) {
        arr[0] = (uint8_t)(arr[0] + arr[6]);
        arr[6] = 0;
        continue;
      }
      const int dist = t.air_dist[src][dst];
      if (dst < NL) {
        if (!owner_allied(g, cx, dst)) g->st.flagged_for_combat[dst] = 2;
        land_ptr(g, dst, BOMBERS_LAND_AIR)[(uint8_t)(6 - dist)]++;
        g->cur_land[dst][BOMBERS_LAND_AIR]++;
        g->tot_land[0][dst]++;
      } else {
        g->st.flagged_for_combat[dst] = g->enemy_count[dst] > 0;
        sea_ptr(g, dst - NL, BOMBERS_SEA)[(uint8_t)(5 - dist)]++;
        g->cur_sea[dst - NL][BOMBERS_SEA]++;
        g->tot_sea[0][dst - NL]++;
      }
      g->cur_land[src][BOMBERS_LAND_AIR]--;
      g->tot_land[0][src]--;
      arr[6]--;
    }
  }
  clear_move_history(g);
  return false;
}

// ---- phase 2: stage_transport_units, 1588-1655 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool stage_transport_units(GameSlot* g, const RunCtx& cx) {
  bool processed = false;
  for (int ut = TRANS_EMPTY; ut <= TRANS_1T; ut++) {
    const int staging = states_sea(ut) - 1, done = staging - 1;
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(g, s, ut);
      if (arr[staging] == 0) continue;
      const uint8_t sa = (uint8_t)(s + NL);
      g->vm[0] = sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, 2, g->sea_blocked);
      while (arr[staging] > 0) {
        processed = true;
        uint8_t da = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          da = ai_input<EXPAND>(g, cx);
        }
        if (sa == da) {
          arr[done] = (uint8_t)(arr[done] + arr[staging]);
          arr[staging] = 0;
          continue;
        }
        const int ds = da - NL;
        const unsigned flat = (unsigned)(s * NS + da);  // 1630: sea_dist indexed with an air id (P1 iii)
        const int dist = flat < 9u ? ((const uint8_t*)cx.T->t.sea_dist[g->canal_state])[flat] : 0;
        if (g->blockade[ds] > 0) {
          g->st.flagged_for_combat[da] = 2;
          continue;  // not moved: the choice is drawn again
        }
        sea_ptr(g, ds, ut)[(uint8_t)(done - dist)]++;
        g->cur_sea[ds][ut]++;
        g->tot_sea[0][ds]++;
        g->small_cargo[ds]++;
        arr[staging]--;
        g->cur_sea[s][ut]--;
        g->tot_sea[0][s]--;
        g->small_cargo[s]--;
        if (ut <= TRANS_1I) {
          g->large_cargo[s]--;
          g->large_cargo[ds]++;
        }
      }
    }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phases 3,4,5,12: move_land_unit_type, 1986-2060 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_land_unit_type(GameSlot* g, const RunCtx& cx, int ut) {
  const CaaaTables& t = cx.T->t;
  bool processed = false;
  for (int src = 0; src < NL; src++)
    for (int moves = max_move_land(ut); moves > 0; moves--) {
      uint8_t* arr = land_ptr(g, src, ut);
      if (arr[moves] == 0) continue;
      processed = true;
      g->vm[0] = (uint8_t)src;
      g->nvm = 1;
      add_valid_land_moves(g, cx, src, moves, ut);
      while (arr[moves] > 0) {
        uint8_t dst = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          dst = ai_input<EXPAND>(g, cx);
        }
        if (src == dst) {
          arr[0] = (uint8_t)(arr[0] + arr[moves]);
          arr[moves] = 0;
          continue;
        }
        if (dst >= NL) {
          if (!load_transport(g, ut, src, dst - NL, moves)) {
            // the reference prints an error and spins for ever here; flag the game and park the units
            g->status |= CAAA_ERR_LOAD_FAILED;
            arr[0] = (uint8_t)(arr[0] + arr[moves]);
            arr[moves] = 0;
            continue;
          }
          g->nvm = 1;
          add_valid_land_moves(g, cx, src, moves, ut);
          continue;
        }
        const bool enemies_there = g->enemy_count[dst] > 0;
        const bool friendly = owner_allied(g, cx, dst);
        g->st.flagged_for_combat[dst] = enemies_there;
        int dist = t.land_dist[src][dst];
        if (friendly || enemies_there) dist = moves;  // not blitzable: the unit's turn ends
        land_ptr(g, dst, ut)[(uint8_t)(moves - dist)]++;
        g->cur_land[dst][ut]++;
        g->tot_land[0][dst]++;
        g->cur_land[src][ut]--;
        g->tot_land[0][src]--;
        arr[moves]--;
        if (!friendly && !enemies_there) {
          conquer_land(g, cx, dst);
          g->st.flagged_for_combat[dst] = 2;
        }
      }
    }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 6: move_transport_units, 2062-2159 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_transport_units(GameSlot* g, const RunCtx& cx) {
  bool processed = false;
  for (int s = 0; s < NS; s++) {  // skip_empty_transports 2062-2075
    uint8_t* e = sea_ptr(g, s, TRANS_EMPTY);
    const uint32_t n = e[1] + e[2] + e[3];
    if (n) {
      e[0] = (uint8_t)(e[0] + n);
      e[1] = e[2] = e[3] = 0;
    }
  }
  for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
    const int max_state = states_sea(ut) - staging_sea(ut);  // always 4
    for (int s = 0; s < NS; s++)
      for (int i = 1; i < 3; i++) {
        const int cur = max_state - i;
        uint8_t* arr = sea_ptr(g, s, ut);
        if (arr[cur] == 0) continue;
        processed = true;
        const uint8_t sa = (uint8_t)(s + NL);
        g->vm[0] = sa;
        g->nvm = 1;
        add_valid_sea_moves(g, cx, s, cur - 1, g->sea_blocked);
        while (arr[cur] > 0) {
          uint8_t da = g->vm[0];
          if (g->nvm > 1) {
            if (g->answers_remaining == 0) return true;
            da = ai_input<EXPAND>(g, cx);
          }
          const int ds = da - NL;
          if (g->blockade[ds] > 0) g->st.flagged_for_combat[da] = 2;
          if (sa == da) {
            arr[1] = (uint8_t)(arr[1] + arr[cur]);
            arr[cur] = 0;
            continue;
          }
          sea_ptr(g, ds, ut)[1]++;
          g->cur_sea[ds][ut]++;
          g->tot_sea[0][ds]++;
          arr[cur]--;
          g->cur_sea[s][ut]--;
          g->tot_sea[0][s]--;
          if (ut <= TRANS_1T) {
            g->small_cargo[ds]++;
            g->small_cargo[s]--;
            if (ut <= TRANS_1I) {
              g->large_cargo[ds]++;
              g->large_cargo[s]--;
            }
          }
        }
      }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 7: move_subs, 2160-2214 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_subs(GameSlot* g, const RunCtx& cx) {
  bool processed = false;
  for (int s = 0; s < NS; s++) {
    uint8_t* arr = sea_ptr(g, s, SUBMARINES);
    if (arr[2] == 0) continue;
    const uint8_t sa = (uint8_t)(s + NL);
    g->vm[0] = sa;
    g->nvm = 1;
    add_valid_sea_moves(g, cx, s, 2, g->sub_blocked);
    while (arr[2] > 0) {
      processed = true;
      uint8_t da = g->vm[0];
      if (g->nvm > 1) {
        if (g->answers_remaining == 0) return true;
        da = ai_input<EXPAND>(g, cx);
      }
      const int ds = da - NL;
      if (g->enemy_count[ds] > 0) g->st.flagged_for_combat[ds] = 2;  // 2188,2194: sea index used as an air index
      if (sa == da) {
        arr[0] = (uint8_t)(arr[0] + arr[2]);
        arr[2] = 0;
        continue;
      }
      sea_ptr(g, ds, SUBMARINES)[0]++;
      g->cur_sea[ds][SUBMARINES]++;
      g->tot_sea[0][ds]++;
      arr[2]--;
      g->cur_sea[s][SUBMARINES]--;
      g->tot_sea[0][s]--;
    }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 8: move_destroyers_battleships, 2216-2310 ----
CAAA_HD void carry_allied_fighters(GameSlot* g, const RunCtx& cx, int src, int dst) {  // 2286-2310
  int moved = 0;
  for (int p = 1; p < NP; p++) {
    if (!allied(g, cx, p)) continue;
    while (tpsut(g, p, src)[FIGHTERS] > 0) {
      tpsut(g, p, src)[FIGHTERS]--;
      g->tot_sea[p][src]--;
      tpsut(g, p, dst)[FIGHTERS]++;
      g->tot_sea[p][dst]++;
      if (moved == 1) return;
      moved++;
    }
  }
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool move_destroyers_battleships(GameSlot* g, const RunCtx& cx) {
  bool processed = false;
  for (int ut = DESTROYERS; ut <= BS_DAMAGED; ut++) {
    const int unmoved = unmoved_sea(ut), done = done_moving_sea(ut);
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(g, s, ut);
      if (arr[unmoved] == 0) continue;
      processed = true;
      const uint8_t sa = (uint8_t)(s + NL);
      g->vm[0] = sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, 2, g->sea_blocked);
      while (arr[unmoved] > 0) {
        uint8_t da = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          da = ai_input<EXPAND>(g, cx);
        }
        if (g->enemy_count[da] > 0) g->st.flagged_for_combat[da] = 2;
        if (sa == da) {
          arr[done] = (uint8_t)(arr[done] + arr[unmoved]);
          arr[unmoved] = 0;
          continue;
        }
        const int ds = da - NL;
        sea_ptr(g, ds, ut)[done]++;
        g->cur_sea[ds][ut]++;
        g->tot_sea[0][ds]++;
        arr[unmoved]--;
        g->cur_sea[s][ut]--;
        g->tot_sea[0][s]--;
        if (ut == CARRIERS) carry_allied_fighters(g, cx, s, ds);
      }
    }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- casualty removal ----
// Take `hits` off a uint8 counter the way every reference casualty loop does: returns true when the
// hits are used up (the reference then returns), false when the counter was wiped and hits remain.
CAAA_HD bool take_hits(uint8_t* n, uint32_t& hits, uint32_t& taken) {
  const uint32_t have = *n;
  if (have < hits) {
    hits -= have;
    taken = have;
    *n = 0;
    return false;
  }
  *n = (uint8_t)(have - hits);
  taken = hits;
  return true;
}
CAAA_HD_NOINLINE void remove_land_defenders(GameSlot* g, const RunCtx& cx, int land, uint32_t hits) {  // 2613-2641
  const int pi = pidx(g), ne = cx.T->n_enemies[pi];
  for (int k = 0; k < 6; k++) {
    const int ut = order_land_def(k);
    for (int e = 0; e < ne; e++) {
      const int p = cx.T->enemies[pi][e];
      uint8_t* n = &g->st.other_land_units[p - 1][land][ut];
      if (*n == 0) continue;
      uint32_t taken;
      const bool done = take_hits(n, hits, taken);
      g->tot_land[p][land] -= (int)taken;
      g->enemy_count[land] -= (int)taken;
      if (done) return;
    }
  }
}
CAAA_HD_NOINLINE void remove_land_attackers(GameSlot* g, int land, uint32_t hits) {  // 2642-2698
  for (int ut = INFANTRY; ut <= TANKS; ut++) {  // ORDER_OF_LAND_ATTACKERS_1, state 0 only
    uint8_t* n = &land_ptr(g, land, ut)[0];
    if (*n == 0) continue;
    uint32_t taken;
    const bool done = take_hits(n, hits, taken);
    g->tot_land[0][land] -= (int)taken;
    if (done) {
      g->cur_land[land][ut] = (uint8_t)(g->cur_land[land][ut] - taken);
      return;
    }
    g->cur_land[land][ut] = 0;
  }
  for (int ut = FIGHTERS; ut <= BOMBERS_LAND_AIR; ut++) {  // ORDER_OF_LAND_ATTACKERS_2, states 1..n-2
    if (g->cur_land[land][ut] == 0) continue;
    for (int s = 1; s < (int)states_land(ut) - 1; s++) {
      uint8_t* n = &land_ptr(g, land, ut)[s];
      if (*n == 0) continue;
      uint32_t taken;
      const bool done = take_hits(n, hits, taken);
      g->cur_land[land][ut] = (uint8_t)(g->cur_land[land][ut] - taken);
      g->tot_land[0][land] -= (int)taken;
      if (done) return;
    }
  }
}
CAAA_HD_NOINLINE void remove_sea_defenders(GameSlot* g, const RunCtx& cx, int sea, uint32_t hits, bool submerged) {  // 2700-2806
  const int pi = pidx(g), ne = cx.T->n_enemies[pi], air = sea + NL;
  for (int e = 0; e < ne; e++) {  // battleships absorb one hit each
    uint8_t* su = g->st.other_sea_units[cx.T->enemies[pi][e] - 1][sea];
    if (su[BATTLESHIPS] == 0) continue;
    uint32_t taken;
    const bool done = take_hits(&su[BATTLESHIPS], hits, taken);
    su[BS_DAMAGED] = (uint8_t)(su[BS_DAMAGED] + taken);
    if (done) return;
  }
  for (int k = submerged ? 1 : 0; k < 13; k++) {
    const int ut = order_sea_def(k);
    const bool in_blockade_range = k >= DESTROYERS && k <= BS_DAMAGED;  // 2783,2791: order index vs type ids
    for (int e = 0; e < ne; e++) {
      const int p = cx.T->enemies[pi][e];
      uint8_t* n = &g->st.other_sea_units[p - 1][sea][ut];
      if (*n == 0) continue;
      uint32_t taken;
      const bool done = take_hits(n, hits, taken);
      g->tot_sea[p][sea] -= (int)taken;
      g->enemy_count[air] -= (int)taken;
      if (in_blockade_range) g->blockade[sea] = (uint8_t)(g->blockade[sea] - taken);
      if (done) return;
    }
  }
}
// attacker casualties among state-0 units of one type; true when all hits are absorbed
CAAA_HD bool hit_sea_state0(GameSlot* g, int sea, int ut, uint32_t& hits) {
  uint8_t* n = &sea_ptr(g, sea, ut)[0];
  if (*n == 0) return false;
  uint32_t taken;
  const bool done = take_hits(n, hits, taken);
  g->tot_sea[0][sea] -= (int)taken;
  if (ut >= TRANS_EMPTY && ut <= TRANS_1T) {
    g->small_cargo[sea] -= (int)taken;
    if (ut <= TRANS_1I) g->large_cargo[sea] -= (int)taken;
  }
  if (done) {
    g->cur_sea[sea][ut] = (uint8_t)(g->cur_sea[sea][ut] - taken);
    if (ut == CARRIERS) g->carriers[sea] -= (int)taken;
  } else {
    g->cur_sea[sea][ut] = 0;
    if (ut == CARRIERS) g->carriers[sea] = 0;
  }
  return done;
}
CAAA_HD_NOINLINE void remove_sea_attackers(GameSlot* g, int sea, uint32_t hits) {  // 2808-2982
  uint8_t* bb = &sea_ptr(g, sea, BATTLESHIPS)[0];
  if (*bb > 0) {
    uint32_t taken;
    const bool done = take_hits(bb, hits, taken);
    uint8_t* bd = &sea_ptr(g, sea, BS_DAMAGED)[0];
    *bd = (uint8_t)(*bd + taken);
    g->cur_sea[sea][BS_DAMAGED] = (uint8_t)(g->cur_sea[sea][BS_DAMAGED] + taken);
    if (done) {
      g->cur_sea[sea][BATTLESHIPS] = (uint8_t)(g->cur_sea[sea][BATTLESHIPS] - taken);
      return;
    }
    g->cur_sea[sea][BATTLESHIPS] = 0;
  }
  if (hit_sea_state0(g, sea, SUBMARINES, hits)) return;
  if (hit_sea_state0(g, sea, DESTROYERS, hits)) return;
  if (hit_sea_state0(g, sea, CARRIERS, hits)) return;
  if (hit_sea_state0(g, sea, CRUISERS, hits)) return;
  for (int k = 0; k < 2; k++) {  // air units in any state, 2915-2942
    const int ut = k == 0 ? (int)FIGHTERS : (int)BOMBERS_SEA;
    if (g->cur_sea[sea][ut] == 0) continue;
    for (int s = 0; s < 5; s++) {
      uint8_t* n = &sea_ptr(g, sea, ut)[s];
      if (*n == 0) continue;
      uint32_t taken;
      const bool done = take_hits(n, hits, taken);
      g->tot_sea[0][sea] -= (int)taken;
      g->cur_sea[sea][ut] = (uint8_t)(g->cur_sea[sea][ut] - taken);
      if (done) return;
    }
  }
  for (int k = 0; k < 8; k++)
    if (hit_sea_state0(g, sea, order_sea_att3(k), hits)) return;
}

// ---- phase 9: resolve_sea_battles, 2312-2550, 2582-2603 ----
CAAA_HD void sea_retreat(GameSlot* g, int src, int dst) {  // 2582-2603
  for (int ut = TRANS_EMPTY; ut <= BS_DAMAGED; ut++) {
    const uint32_t n = g->cur_sea[src][ut];
    uint8_t* d = sea_ptr(g, dst, ut);
    d[0] = (uint8_t)(d[0] + n);
    g->cur_sea[dst][ut] = (uint8_t)(g->cur_sea[dst][ut] + n);
    g->tot_sea[0][dst] += (int)n;
    g->tot_sea[0][src] -= (int)n;
    sea_ptr(g, src, ut)[0] = 0;
    sea_ptr(g, src, ut)[1] = 0;
    g->cur_sea[src][ut] = 0;
  }
  g->st.flagged_for_combat[src + NL] = 0;
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool resolve_sea_battles(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  const int pi = pidx(g), ne = cx.T->n_enemies[pi];
  for (int s = 0; s < NS; s++) {
    const int air = s + NL;
    uint8_t* mine = g->cur_sea[s];
    if (g->st.flagged_for_combat[air] == 0) continue;
    if (g->tot_sea[0][s] == 0) continue;
    const bool submerged = mine[DESTROYERS] == 0;
    if (g->tot_sea[0][s] == 2) {  // 2336: literal "== 2"
      if (submerged) {
        uint32_t subs = 0;
        for (int e = 0; e < ne; e++) subs += g->st.other_sea_units[cx.T->enemies[pi][e] - 1][s][SUBMARINES];
        if ((int)(subs & 0xffu) == g->enemy_count[s]) continue;  // 2344: land index
      }
      for (int ut = CRUISERS; ut <= BS_DAMAGED; ut++) {  // bombardment no longer possible
        uint8_t* a = sea_ptr(g, s, ut);
        a[0] = (uint8_t)(a[0] + a[1]);
        a[1] = 0;
      }
    }
    for (int rounds = 0;; rounds++) {
      if (rounds >= SEA_ROUND_CAP) {  // guard (not in the reference, which would spin for ever)
        g->status |= CAAA_ERR_SEA_ROUND_CAP;
        return true;
      }
      if (g->st.flagged_for_combat[air] == 1) {  // retreat option, 2368-2406
        uint32_t n = 0;
        if (mine[FIGHTERS] + mine[BOMBERS_SEA] + mine[SUBMARINES] + mine[DESTROYERS] + mine[CARRIERS] + mine[CRUISERS] +
                    mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0 ||
            g->blockade[s] == 0)
          g->vm[n++] = (uint8_t)air;
        for (int i = 0; i < t.s2s_count[s]; i++) {
          const int d = t.s2s_conn[s][i];
          if (g->blockade[d] == 0) g->vm[n++] = (uint8_t)(d + NL);
        }
        g->nvm = (uint8_t)n;
        if (n > 0) {
          uint8_t da = g->vm[0];
          if (n > 1) {
            if (g->answers_remaining == 0) return true;
            da = ai_input<EXPAND>(g, cx);
          }
          const int ds = (uint8_t)(da - NL);
          if (ds < NS && t.sea_dist[g->canal_state][s][ds] == 1 && g->st.flagged_for_combat[da] == 0) {
            sea_retreat(g, s, ds);
            break;
          }
        }
      }
      g->st.flagged_for_combat[air] = 1;
      bool targets = false;
      if (mine[DESTROYERS]) {
        targets = g->enemy_count[air] > 0;
      } else if (mine[CARRIERS] + mine[CRUISERS] + mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0) {
        if (g->enemy_count[air] > 0)
          for (int e = 0; e < ne; e++) {
            const int p = cx.T->enemies[pi][e];
            if (g->tot_sea[p][s] - g->st.other_sea_units[p - 1][s][SUBMARINES] > 0) { targets = true; break; }
          }
      } else if (mine[SUBMARINES] > 0) {
        for (int e = 0; e < ne; e++) {
          const int p = cx.T->enemies[pi][e];
          const uint8_t* eu = g->st.other_sea_units[p - 1][s];
          if (g->tot_sea[p][s] - (eu[FIGHTERS] + eu[SUBMARINES]) > 0) { targets = true; break; }
        }
      } else if (mine[FIGHTERS] + mine[BOMBERS_SEA] > 0) {
        for (int e = 0; e < ne; e++) {
          const int p = cx.T->enemies[pi][e];
          if (g->tot_sea[p][s] - g->st.other_sea_units[p - 1][s][SUBMARINES] > 0) { targets = true; break; }
        }
      }
      if (!targets) {  // 2460-2482
        if (g->enemy_count[air] > 0) {
          g->tot_sea[0][s] -= mine[TRANS_EMPTY];
          mine[TRANS_EMPTY] = 0;
          sea_ptr(g, s, TRANS_EMPTY)[0] = 0;
          for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
            g->tot_sea[0][s] -= mine[ut];
            mine[ut] = 0;
            sea_ptr(g, s, ut)[1] = 0;
          }
        }
        g->st.flagged_for_combat[air] = 0;
        break;
      }
      // sub volley: only state-0 attacking subs count (2484)
      int admg = sea_ptr(g, s, SUBMARINES)[0] * 2;
      uint32_t ahits = attacker_hits(g, cx, admg);
      if (!submerged) {
        int ddmg = 0;
        for (int e = 0; e < ne; e++) ddmg += g->st.other_sea_units[cx.T->enemies[pi][e] - 1][s][SUBMARINES];
        const uint32_t dhits = defender_hits(g, cx, ddmg);
        if (dhits > 0) remove_sea_attackers(g, s, dhits);
      }
      if (ahits > 0) remove_sea_defenders(g, cx, s, ahits, submerged);
      // main volley, 2513-2542
      admg = mine[DESTROYERS] * 2 + mine[CARRIERS] * 1 + mine[CRUISERS] * 3 + mine[BATTLESHIPS] * 4 + mine[BS_DAMAGED] * 4 +
             mine[FIGHTERS] * 3 + mine[BOMBERS_SEA] * 4;
      ahits = attacker_hits(g, cx, admg);
      int ddmg = 0;
      for (int e = 0; e < ne; e++) {  // 2527-2530: type ids 0..4 (only fighters have a defence value) + fighters again
        const uint8_t* eu = g->st.other_sea_units[cx.T->enemies[pi][e] - 1][s];
        ddmg += eu[FIGHTERS] * 4 * 2;
      }
      const uint32_t dhits = defender_hits(g, cx, ddmg);
      if (dhits > 0) remove_sea_attackers(g, s, dhits);
      if (ahits > 0) remove_sea_defenders(g, cx, s, ahits, submerged);
      if (g->enemy_count[air] == 0 || g->tot_sea[0][s] == 0) {
        g->st.flagged_for_combat[air] = 0;
        break;
      }
    }
  }
  return false;
}

// ---- phase 10: unload_transports, 2984-3053, 3373-3383 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool unload_transports(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  bool processed = false;
  for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
    const int c1 = unload_cargo1(ut);
    for (int s = 0; s < NS; s++) {
      uint8_t* arr = sea_ptr(g, s, ut);
      if (arr[1] == 0) continue;  // STATES_UNLOADING is 1 for every loaded transport type
      processed = true;
      const uint8_t sa = (uint8_t)(s + NL);
      g->vm[0] = sa;
      uint32_t n = 1;
      for (int i = 0; i < t.s2l_count[s]; i++) {
        const int d = t.s2l_conn[s][i];
        if (!skip_get(g, sa, d)) g->vm[n++] = (uint8_t)d;
      }
      g->nvm = (uint8_t)n;
      while (arr[1] > 0) {
        uint8_t dst = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          dst = ai_input<EXPAND>(g, cx);
        }
        if (sa == dst) {
          arr[0] = (uint8_t)(arr[0] + arr[1]);
          arr[1] = 0;
          continue;
        }
        g->st.land_state[dst].bombard_max++;
        land_ptr(g, dst, c1)[0]++;
        g->cur_land[dst][c1]++;
        g->tot_land[0][dst]++;
        sea_ptr(g, s, TRANS_EMPTY)[0]++;
        g->cur_sea[s][TRANS_EMPTY]++;
        g->cur_sea[s][ut]--;
        arr[1]--;
        if (ut > TRANS_1T) {  // second cargo is always infantry
          g->st.land_state[dst].bombard_max++;
          land_ptr(g, dst, INFANTRY)[0]++;
          g->cur_land[dst][INFANTRY]++;
          g->tot_land[0][dst]++;
        }
        if (!owner_allied(g, cx, dst)) {
          g->st.flagged_for_combat[dst] = 2;
          if (g->enemy_count[dst] == 0) conquer_land(g, cx, dst);
        }
      }
    }
  }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 11: resolve_land_battles, 3055-3371 ----
CAAA_HD void hit_air_states(GameSlot* g, int land, int ut, int lo, int hi, uint32_t& hits, bool stop_on_absorb) {
  // shared shape of the strategic-bombing / AA-fire casualty loops (3103-3117, 3179-3210)
  for (int s = lo; s < hi; s++) {
    uint32_t taken;
    const bool done = take_hits(&land_ptr(g, land, ut)[s], hits, taken);
    g->cur_land[land][ut] = (uint8_t)(g->cur_land[land][ut] - taken);
    g->tot_land[0][land] -= (int)taken;
    if (done) {
      hits = 0;
      if (stop_on_absorb) return;
    }
  }
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool resolve_land_battles(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  const int pi = pidx(g), ne = cx.T->n_enemies[pi];
  for (int l = 0; l < NL; l++) {
    if (g->st.flagged_for_combat[l] == 0) continue;
    uint8_t* mine = g->cur_land[l];
    int* my_total = &g->tot_land[0][l];
    CaaaLandState* ls = &g->st.land_state[l];
    if (*my_total == 0) continue;
    if (g->st.flagged_for_combat[l] == 2) {
      if (mine[BOMBERS_LAND_AIR] > 0 && *my_total == mine[BOMBERS_LAND_AIR]) {  // strategic bombing 3088-3125
        uint32_t ahits = 0;  // uninitialised in the reference when the factory is already at -max; same result
        if (ls->factory_hp > -(int)ls->factory_max) {
          uint32_t dh = defender_hits(g, cx, mine[BOMBERS_LAND_AIR]);
          if (dh > 0) hit_air_states(g, l, BOMBERS_LAND_AIR, 1, 6, dh, false);
          ahits = attacker_hits(g, cx, mine[BOMBERS_LAND_AIR] * 21);
        }
        const int hp = ls->factory_hp - (int)ahits, floor_hp = -(int)ls->factory_max;
        ls->factory_hp = (int8_t)(hp > floor_hp ? hp : floor_hp);
        continue;
      }
      if (ls->bombard_max > 0) {  // shore bombardment 3134-3158
        int admg = 0;
        for (int ut = BS_DAMAGED; ut >= CRUISERS; ut--)
          for (int i = 0; i < t.l2s_count[l]; i++) {
            uint8_t* ships = sea_ptr(g, t.l2s_conn[l][i], ut);
            while (ships[1] > 0 && ls->bombard_max > 0) {
              admg += attack_sea(ut);
              ships[0]++;
              ships[1]--;
              ls->bombard_max--;
            }
          }
        ls->bombard_max = 0;
        const uint32_t ahits = attacker_hits(g, cx, admg);
        if (ahits > 0) remove_land_defenders(g, cx, l, ahits);
      }
      const uint32_t air_units = (uint8_t)(mine[FIGHTERS] + mine[BOMBERS_LAND_AIR]);  // AA fire 3160-3216
      if (air_units > 0) {
        int aa = 0;
        for (int e = 0; e < ne; e++) aa += g->st.other_land_units[cx.T->enemies[pi][e] - 1][l][AA_GUNS];
        if (aa > 0) {
          uint32_t dh = defender_hits(g, cx, (uint8_t)(air_units * 3));
          if (dh > 0) hit_air_states(g, l, FIGHTERS, 0, 5, dh, true);
          if (dh > 0) hit_air_states(g, l, BOMBERS_LAND_AIR, 0, 7, dh, true);
        }
      }
      if (*my_total == 0) continue;
      if (g->enemy_count[l] == 0) {
        if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(g, cx, l);
        continue;
      }
    }
    uint32_t rounds = 0;
    for (;;) {
      rounds = (rounds + 1) & 0xffu;  // uint8 counter in the reference
      if (*my_total == 0) break;
      if (g->st.flagged_for_combat[l] == 1) {  // retreat option 3256-3299
        g->vm[0] = (uint8_t)l;
        uint32_t n = 1;
        for (int i = 0; i < t.l2l_count[l]; i++) {
          const int d = t.l2l_conn[l][i];
          if (g->enemy_count[d] == 0 && g->st.flagged_for_combat[d] == 0 && owner_allied(g, cx, d)) g->vm[n++] = (uint8_t)d;
        }
        g->nvm = (uint8_t)n;
        uint8_t dst = g->vm[0];
        if (n > 1) {
          if (g->answers_remaining == 0) return true;
          dst = ai_input<EXPAND>(g, cx);
        }
        if (l != dst || rounds > 100) {
          const int dl = dst < NL ? dst : l;  // dst is always a land here
          for (int ut = INFANTRY; ut <= TANKS; ut++) {
            const uint32_t k = land_ptr(g, l, ut)[0];
            land_ptr(g, dl, ut)[0] = (uint8_t)(land_ptr(g, dl, ut)[0] + k);
            g->cur_land[dl][ut] = (uint8_t)(g->cur_land[dl][ut] + k);
            g->tot_land[0][dl] += (int)k;
            g->cur_land[l][ut] = (uint8_t)(g->cur_land[l][ut] - k);
            *my_total -= (int)k;
            land_ptr(g, l, ut)[0] = 0;
          }
          g->st.flagged_for_combat[l] = 0;
          break;
        }
      }
      g->st.flagged_for_combat[l] = 1;
      const int inf = mine[INFANTRY], art = mine[ARTILLERY];
      const int admg = mine[FIGHTERS] * 3 + mine[BOMBERS_LAND_AIR] * 4 + inf + art * 2 + mine[TANKS] * 3 + (inf < art ? inf : art);
      const uint32_t ahits = attacker_hits(g, cx, admg);
      int ddmg = 0;
      uint32_t dh = 0;
      for (int e = 0; e < ne; e++) {  // 3319-3327: one draw per enemy player, the last one counts
        const uint8_t* eu = g->st.other_land_units[cx.T->enemies[pi][e] - 1][l];
        ddmg += eu[INFANTRY] * 2 + eu[ARTILLERY] * 2 + eu[TANKS] * 3 + eu[FIGHTERS] * 4 + eu[BOMBERS_LAND_AIR];
        dh = defender_hits(g, cx, ddmg);
      }
      if (dh > 0) remove_land_attackers(g, l, dh);
      if (ahits > 0) remove_land_defenders(g, cx, l, ahits);
      if (*my_total == 0) break;
      if (g->enemy_count[l] == 0) {
        if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(g, cx, l);
        break;
      }
    }
  }
  return false;
}

// ---- phases 13/14: land_fighter_units / land_bomber_units, 3409-3421, 3445-3654 ----
CAAA_HD uint32_t fighter_final_landing_mask(GameSlot* g, const RunCtx& cx) {  // refresh_can_fighter_land_here 3445-3466
  const CaaaTables& t = cx.T->t;
  uint32_t here = 0;
  for (int s = 0; s < NS; s++)
    if (g->carriers[s] > 0) here |= 1u << (s + NL);
  for (int l = 0; l < NL; l++) {
    const int owner = g->st.land_state[l].owner_idx;
    if (allied(g, cx, owner) && g->st.flagged_for_combat[l] == 0) here |= 1u << l;
    if (g->st.land_state[l].factory_max > 0 && owner == g->st.player_index)
      for (int i = 0; i < t.l2s_count[l]; i++) here |= 1u << (NL + t.l2s_conn[l][i]);
  }
  return here;
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool land_fighter_units(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  bool processed = false;
  uint32_t here = 0;
  for (int cur = 1; cur < 4; cur++)
    for (int src = 0; src < NA; src++) {
      uint8_t* arr = air_fighters(g, src);
      if (arr[cur] == 0) continue;
      if (!processed) {
        processed = true;
        here = fighter_final_landing_mask(g, cx);
      }
      uint32_t n = 0;
      for (int i = 0; i < t.air_within_count[cur - 1][src]; i++) {  // add_valid_fighter_landing
        const int d = t.air_within[cur - 1][src][i];
        if (((here >> d) & 1) && !skip_get(g, src, d)) g->vm[n++] = (uint8_t)d;
      }
      if (n == 0 || ((here >> src) & 1)) g->vm[n++] = (uint8_t)src;
      g->nvm = (uint8_t)n;
      while (arr[cur] > 0) {
        uint8_t dst = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          dst = ai_input<EXPAND>(g, cx);
        }
        if (src == dst) {
          arr[0]++;
          arr[cur]--;
          continue;
        }
        air_fighters(g, dst)[0]++;
        if (dst < NL) {
          g->tot_land[0][dst]++;
          g->cur_land[dst][FIGHTERS]++;
        } else {
          g->tot_sea[0][dst - NL]++;
          g->cur_sea[dst - NL][FIGHTERS]++;
        }
        if (src < NL) {
          g->tot_land[0][src]--;
          g->cur_land[src][FIGHTERS]--;
        } else {
          g->tot_sea[0][src - NL]--;
          g->cur_sea[src - NL][FIGHTERS]--;
        }
        arr[cur]--;
      }
    }
  if (processed) clear_move_history(g);
  return false;
}
template <bool EXPAND>
CAAA_HD_NOINLINE bool land_bomber_units(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  bool processed = false;
  uint32_t here = 0;
  for (int cur1 = 0; cur1 < 5; cur1++)
    for (int src = 0; src < NA; src++) {
      const int cur = src < NL ? cur1 + 1 : cur1;
      uint8_t* arr = air_bombers(g, src);
      if (arr[cur] == 0) continue;
      if (!processed) {
        processed = true;
        here = bomber_here_mask(g, cx);
      }
      const int moves = cur + (src < NL ? 0 : 1);
      uint32_t n = 0;
      for (int i = 0; i < t.air_to_land_within_count[moves - 1][src]; i++) {  // add_valid_bomber_landing
        const int d = t.air_to_land_within[moves - 1][src][i];
        if ((here >> d) & 1) g->vm[n++] = (uint8_t)d;
      }
      g->nvm = (uint8_t)n;
      while (arr[cur] > 0) {
        if (g->nvm == 0) g->vm[g->nvm++] = (uint8_t)src;
        uint8_t dst = g->vm[0];
        if (g->nvm > 1) {
          if (g->answers_remaining == 0) return true;
          dst = ai_input<EXPAND>(g, cx);
        }
        if (src == dst) {
          arr[0]++;
          arr[cur]--;
          continue;
        }
        air_bombers(g, dst)[0]++;
        const int dl = dst < NL ? dst : 0;  // bombers only land on land
        g->tot_land[0][dl]++;
        g->cur_land[dl][BOMBERS_LAND_AIR]++;
        if (src < NL) {
          g->tot_land[0][src]--;
          g->cur_land[src][BOMBERS_LAND_AIR]--;
        } else {
          g->tot_sea[0][src - NL]--;
          g->cur_sea[src - NL][BOMBERS_LAND_AIR]--;  // 3638: type index 1 = empty transports
        }
        arr[cur]--;
      }
    }
  if (processed) clear_move_history(g);
  return false;
}

// ---- phase 15: buy_units (purchase and placement), 3663-3834 ----
template <bool EXPAND>
CAAA_HD_NOINLINE bool buy_units(GameSlot* g, const RunCtx& cx) {
  const CaaaTables& t = cx.T->t;
  for (int f = 0; f < g->nfact[0] && f < 256; f++) {
    const int land = g->factloc[0][f];
    CaaaLandState* ls = &g->st.land_state[land];
    if (g->st.builds_left[land] == 0) continue;
    uint32_t repair = 0;
    for (int si = 0; si < t.l2s_count[land]; si++) {
      bool processed = false;
      const int sea = t.l2s_conn[land][si], air = sea + NL;
      g->vm[0] = NST;  // pass
      uint32_t last = 0;
      while (g->st.builds_left[air] > 0) {
        if (g->st.money[0] < 8) {
          g->st.builds_left[air] = 0;
          break;
        }
        const uint32_t built = (uint8_t)(ls->factory_max - g->st.builds_left[land]);
        if (ls->factory_hp <= (int)built) repair = (uint8_t)(1 + built - ls->factory_hp);
        uint32_t n = 1;
        for (int k = 6; k >= 0; k--) {
          const uint32_t ut = buy_sea(k);
          if (skip_get(g, 0, ut)) {
            last = ut;
            break;
          }
          if (g->st.money[0] < cost_sea(ut) + repair) continue;
          if (ut == FIGHTERS) {
            int fighters = 0;
            for (int p = 0; p < NP; p++) fighters += tpsut(g, p, sea)[FIGHTERS];
            if (g->carriers[sea] * 2 <= fighters) continue;
          }
          g->vm[n++] = (uint8_t)ut;
        }
        g->nvm = (uint8_t)n;
        if (n == 1) {
          g->st.builds_left[air] = 0;
          break;
        }
        if (g->answers_remaining == 0) return true;
        processed = true;
        const uint32_t buy = ai_input<EXPAND>(g, cx);
        if (buy == NST) {
          g->st.builds_left[air] = 0;
          break;
        }
        for (int s2 = si; s2 < t.l2s_count[land]; s2++) g->st.builds_left[t.l2s_conn[land][s2] + NL]--;
        g->st.builds_left[land]--;
        ls->factory_hp = (int8_t)(ls->factory_hp + (int)repair);
        const uint32_t b = buy < NST ? buy : 0;  // only listed ids can be drawn; keeps expansion-mode inputs in range
        g->st.money[0] = (uint8_t)(g->st.money[0] - (cost_sea(b) + repair));
        sea_ptr(g, sea, b)[0]++;
        g->tot_sea[0][sea]++;
        g->cur_sea[sea][b]++;
        if (buy > last) {
          for (uint32_t u = last; u < buy; u++) skip_set(g, 0, u);
          last = buy;
        }
      }
      if (processed) clear_move_history(g);
    }
    g->vm[0] = NLT;  // pass
    uint32_t last = 0;
    bool processed = false;
    while (g->st.builds_left[land] > 0) {
      if (g->st.money[0] < 3) {
        g->st.builds_left[land] = 0;
        break;
      }
      const uint32_t built = (uint8_t)(ls->factory_max - g->st.builds_left[land]);
      if (ls->factory_hp <= (int)built) repair = (uint8_t)(1 + built - ls->factory_hp);
      uint32_t n = 1;
      for (int ut = NLT - 1; ut >= 0; ut--) {
        if (skip_get(g, 0, (unsigned)ut)) {
          last = (uint32_t)ut;
          break;
        }
        if (g->st.money[0] < cost_land(ut) + repair) continue;
        g->vm[n++] = (uint8_t)ut;
      }
      g->nvm = (uint8_t)n;
      if (n == 1) {
        g->st.builds_left[land] = 0;
        break;
      }
      if (g->answers_remaining == 0) return true;
      processed = true;
      const uint32_t buy = ai_input<EXPAND>(g, cx);
      if (buy == NLT) {
        g->st.builds_left[land] = 0;
        break;
      }
      g->st.builds_left[land]--;
      ls->factory_hp = (int8_t)(ls->factory_hp + (int)repair);
      const uint32_t b = buy < NLT ? buy : 0;
      g->st.money[0] = (uint8_t)(g->st.money[0] - (cost_land(b) + repair));
      land_ptr(g, land, b)[0]++;
      g->tot_land[0][land]++;
      g->cur_land[land][b]++;
      if (buy > last)
        for (uint32_t u = last; u < buy; u++) skip_set(g, 0, u);
      last = buy;
    }
    if (processed) clear_move_history(g);
  }
  return false;
}

// ---- phases 16-19: crash_air_units, reset_units_fully, collect_money, 3836-3894 ----
CAAA_HD_NOINLINE void crash_air_units(GameSlot* g, const RunCtx& cx) {
  const uint32_t here = fighter_landing_masks(g, cx);
  for (int l = 0; l < NL; l++) {
    if ((here >> l) & 1) continue;
    const uint32_t n = g->cur_land[l][FIGHTERS];
    if (n == 0) continue;
    g->tot_land[0][l] -= (int)n;
    g->cur_land[l][FIGHTERS] = 0;
    land_ptr(g, l, FIGHTERS)[0] = 0;
  }
  for (int s = 0; s < NS; s++) {
    uint32_t space = (uint32_t)(g->carriers[s] * 2);
    for (int p = 1; p < NP; p++) space -= g->st.other_sea_units[p - 1][s][FIGHTERS];
    space &= 0xffu;  // uint8 in the reference
    uint8_t* n = &sea_ptr(g, s, FIGHTERS)[0];
    if (space < *n) {
      const uint32_t lost = *n - space;
      g->tot_sea[0][s] -= (int)lost;
      g->cur_sea[s][FIGHTERS] = (uint8_t)(g->cur_sea[s][FIGHTERS] - lost);
      *n = (uint8_t)space;
    }
  }
}
CAAA_HD void reset_units_fully(GameSlot* g) {
  for (int s = 0; s < NS; s++) {
    const uint32_t dmg = g->cur_sea[s][BS_DAMAGED];
    uint8_t* bb = sea_ptr(g, s, BATTLESHIPS);
    bb[0] = (uint8_t)(bb[0] + dmg);
    g->cur_sea[s][BATTLESHIPS] = (uint8_t)(g->cur_sea[s][BATTLESHIPS] + dmg);
    sea_ptr(g, s, BS_DAMAGED)[0] = 0;
    sea_ptr(g, s, BS_DAMAGED)[1] = 0;
    g->cur_sea[s][BS_DAMAGED] = 0;
  }
}
CAAA_HD void collect_money(GameSlot* g, const RunCtx& cx) {
  if (g->st.land_state[cx.T->t.capital[pidx(g)]].owner_idx == 0) g->st.money[0] = (uint8_t)(g->st.money[0] + g->income[0]);
}

// ---- phase 20: rotate_turns, 3895-4029 ----
CAAA_HD_NOINLINE void rotate_turns(GameSlot* g, const RunCtx& cx) {
  CaaaGameState* st = &g->st;
  // per-player unit totals: current <- other[0] <- other[1] <- other[2] <- other[3] <- current
  for (int l = 0; l < NL; l++)
    for (int u = 0; u < NLT; u++) {
      const uint8_t mine = g->cur_land[l][u];
      g->cur_land[l][u] = st->other_land_units[0][l][u];
      st->other_land_units[0][l][u] = st->other_land_units[1][l][u];
      st->other_land_units[1][l][u] = st->other_land_units[2][l][u];
      st->other_land_units[2][l][u] = st->other_land_units[3][l][u];
      st->other_land_units[3][l][u] = mine;
    }
  for (int l = 0; l < NL; l++) {  // everything of the incoming player starts unmoved
    uint8_t* base = (uint8_t*)&st->land_state[l];
    for (int k = 4; k < 25; k++) base[k] = 0;
    for (int u = 0; u < NLT; u++) base[off_land(u) + states_land(u) - 1] = g->cur_land[l][u];
  }
  for (int s = 0; s < NS; s++) {
    for (int u = 0; u < NST - 1; u++) {  // 3944: only 14 of 15 types, the sea-bomber total stays stale
      const uint8_t mine = g->cur_sea[s][u];
      g->cur_sea[s][u] = st->other_sea_units[0][s][u];
      st->other_sea_units[0][s][u] = st->other_sea_units[1][s][u];
      st->other_sea_units[1][s][u] = st->other_sea_units[2][s][u];
      st->other_sea_units[2][s][u] = st->other_sea_units[3][s][u];
      st->other_sea_units[3][s][u] = mine;
    }
    uint8_t* base = (uint8_t*)&st->units_sea[s];
    for (int k = 0; k < 57; k++) base[k] = 0;
    for (int u = 0; u < NST; u++) base[off_sea(u) + states_sea(u) - 1] = g->cur_sea[s][u];
  }
  const uint8_t m = st->money[0];
  for (int p = 0; p < NP - 1; p++) st->money[p] = st->money[p + 1];
  st->money[NP - 1] = m;
  // derived per-player caches rotate the same way; row 5 keeps a copy of the outgoing player
  g->income[5] = g->income[0];
  g->nfact[5] = g->nfact[0];
  for (int l = 0; l < NL; l++) {
    g->factloc[5][l] = g->factloc[0][l];
    g->tot_land[5][l] = g->tot_land[0][l];
  }
  for (int s = 0; s < NS; s++) g->tot_sea[5][s] = g->tot_sea[0][s];
  for (int p = 0; p < NP; p++) {
    g->income[p] = g->income[p + 1];
    g->nfact[p] = g->nfact[p + 1];
    for (int l = 0; l < NL; l++) {
      g->factloc[p][l] = g->factloc[p + 1][l];
      g->tot_land[p][l] = g->tot_land[p + 1][l];
    }
    for (int s = 0; s < NS; s++) g->tot_sea[p][s] = g->tot_sea[p + 1][s];
  }
  for (int a = 0; a < NA; a++) st->flagged_for_combat[a] = 0;
  st->player_index = (uint8_t)((st->player_index + 1) % NP);
  for (int l = 0; l < NL; l++) st->land_state[l].owner_idx = (uint8_t)((st->land_state[l].owner_idx + NP - 1) % NP);
  for (int f = 0; f < g->nfact[0]; f++) {
    const int land = g->factloc[0][f];
    const uint8_t fm = st->land_state[land].factory_max;
    st->builds_left[land] = fm;
    for (int i = 0; i < cx.T->t.l2s_count[land]; i++) {
      uint8_t* b = &st->builds_left[cx.T->t.l2s_conn[land][i] + NL];
      *b = (uint8_t)(*b + fm);
    }
  }
  refresh_eot_cache(g, cx);
}

// ---- scoring, 4301-4337 ----
CAAA_HD_NOINLINE double evaluate_state(GameSlot* g, const RunCtx& cx) {
  const uint8_t starting = g->st.player_index;
  const int pi = starting >= NP ? starting - NP : starting;
  int allied_score = 1 + g->st.money[0], enemy_score = 1;  // 4309+4311: the mover's money counts twice
  const uint8_t* other_sea_flat = &g->st.other_sea_units[0][0][0];
  for (int p = 0; p < NP; p++) {
    int score = g->st.money[p];
    for (int l = 0; l < NL; l++) {
      const uint8_t* r = tplut(g, p, l);
      for (int u = 0; u < NLT; u++) score += r[u] * (int)cost_land(u);
    }
    for (int s = 0; s < NS; s++)
      for (int u = 0; u < NST; u++) {  // 4319: 15 entries of a 14-entry row run into the next row / flagged_for_combat[0]
        const int n = p == 0 ? g->cur_sea[s][u] : other_sea_flat[((p - 1) * NS + s) * 14 + u];
        score += n * (int)cost_sea(u);
      }
    if (cx.T->t.is_allied[pi % NP][(pi + p) % NP]) allied_score += score;
    else enemy_score += score;
  }
  const double score = (double)allied_score / (double)(enemy_score + allied_score);
  return starting >= NP ? 1 - score : score;
}

// One phase of the pipeline by id (the order is fixed: reference src/engine.cpp:4214-4235).
template <bool EXPAND>
CAAA_HD bool run_phase(GameSlot* g, const RunCtx& cx, int ph) {
  switch (ph) {
    case CAAA_PH_MOVE_FIGHTERS: return move_fighter_units<EXPAND>(g, cx);
    case CAAA_PH_MOVE_BOMBERS: return move_bomber_units<EXPAND>(g, cx);
    case CAAA_PH_STAGE_TRANSPORTS: return stage_transport_units<EXPAND>(g, cx);
    case CAAA_PH_MOVE_TANKS: return move_land_unit_type<EXPAND>(g, cx, TANKS);
    case CAAA_PH_MOVE_ARTILLERY: return move_land_unit_type<EXPAND>(g, cx, ARTILLERY);
    case CAAA_PH_MOVE_INFANTRY: return move_land_unit_type<EXPAND>(g, cx, INFANTRY);
    case CAAA_PH_MOVE_TRANSPORTS: return move_transport_units<EXPAND>(g, cx);
    case CAAA_PH_MOVE_SUBS: return move_subs<EXPAND>(g, cx);
    case CAAA_PH_MOVE_SHIPS: return move_destroyers_battleships<EXPAND>(g, cx);
    case CAAA_PH_SEA_BATTLES: return resolve_sea_battles<EXPAND>(g, cx);
    case CAAA_PH_UNLOAD_TRANSPORTS: return unload_transports<EXPAND>(g, cx);
    case CAAA_PH_LAND_BATTLES: return resolve_land_battles<EXPAND>(g, cx);
    case CAAA_PH_MOVE_AA_GUNS: return move_land_unit_type<EXPAND>(g, cx, AA_GUNS);
    case CAAA_PH_LAND_FIGHTERS: return land_fighter_units<EXPAND>(g, cx);
    case CAAA_PH_LAND_BOMBERS: return land_bomber_units<EXPAND>(g, cx);
    case CAAA_PH_BUY_UNITS: return buy_units<EXPAND>(g, cx);
    case CAAA_PH_CRASH_AIR: crash_air_units(g, cx); return false;
    case CAAA_PH_RESET_UNITS: reset_units_fully(g); return false;
    case CAAA_PH_BUY_FACTORY: return false;  // buy_factory is empty (3889)
    case CAAA_PH_COLLECT_MONEY: collect_money(g, cx); return false;
    case CAAA_PH_ROTATE_TURNS: rotate_turns(g, cx); return false;
    default: return false;
  }
}

// Prepare a slot whose `st` has just been filled: what random_play_until_terminal does before its
// loop (4185-4203, with P1 iv and the P2 stream reset).
CAAA_HD void begin_playout(GameSlot* g, const RunCtx& cx, uint64_t game_id) {
  // entries of the path-blocked tables that no refresh ever writes are read by the flat-index quirk
  // (add_valid_land_moves); they are zero in the reference because its globals are zero-initialised
  for (int i = 0; i < 25; i++) g->land_blocked[i] = 0;
  for (int i = 0; i < 9; i++) g->sea_blocked[i] = g->sub_blocked[i] = 0;
  for (int p = 0; p < 6; p++)
    for (int l = 0; l < NL; l++) g->factloc[p][l] = 0;
  g->income[5] = 0;
  g->nfact[5] = 0;
  refresh_full_cache(g, cx);
  g->answers_remaining = 100000;
  g->use_selected = 0;
  g->selected_action = 0;
  g->unlucky = 0;
  g->draws = 0;
  g->game_lo = (uint32_t)game_id;
  g->game_hi = (uint32_t)(game_id >> 32);
  g->turns = 0;
  g->status = 0;
  for (int i = 0; i < CAAA_ACTIONS; i++) g->vm[i] = 0;
  g->nvm = 0;
}

CAAA_HD uint64_t hash_live_state(const GameSlot* g) {  // FNV-1a-64 over the first 606 bytes
  uint64_t h = CAAA_FNV_OFFSET;
  const uint8_t* p = (const uint8_t*)&g->st;
  for (int i = 0; i < CAAA_STATE_LIVE_BYTES; i++) {
    h ^= p[i];
    h *= CAAA_FNV_PRIME;
  }
  return h;
}

// ---- phase routing -------------------------------------------------------------------------------
// A phase of the pipeline is a no-op on a game unless some unit is still in a state that the phase
// moves (the reference relies on the same property for its stop-and-resume expansion mode: every
// phase is idempotent on "done" units).  phase_has_work(p) is true exactly when
// phase p would touch the slot; next_phase_with_work() lets the scheduler queue a game only for
// phases that do something.  Phases >= CAAA_PH_BUY_UNITS always run (end-of-turn work item).
CAAA_HD bool phase_has_work(const GameSlot* g, int ph) {
  const CaaaGameState& st = g->st;
  uint32_t any = 0;
  switch (ph) {
    case CAAA_PH_MOVE_FIGHTERS:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].fighters[4];
      for (int s = 0; s < NS; s++) any |= st.units_sea[s].fighters[4];
      break;
    case CAAA_PH_MOVE_BOMBERS:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].bombers[6];
      break;
    case CAAA_PH_STAGE_TRANSPORTS:
      for (int s = 0; s < NS; s++)
        any |= st.units_sea[s].trans_empty[3] | st.units_sea[s].trans_1i[4] | st.units_sea[s].trans_1a[4] | st.units_sea[s].trans_1t[4];
      break;
    case CAAA_PH_MOVE_TANKS:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].tanks[1] | st.land_state[l].tanks[2];
      break;
    case CAAA_PH_MOVE_ARTILLERY:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].artillery[1];
      break;
    case CAAA_PH_MOVE_INFANTRY:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].infantry[1];
      break;
    case CAAA_PH_MOVE_TRANSPORTS:
      for (int s = 0; s < NS; s++) {
        const CaaaUnitsSea& u = st.units_sea[s];
        any |= u.trans_empty[1] | u.trans_empty[2] | u.trans_empty[3];  // skip_empty_transports
        any |= u.trans_1i[2] | u.trans_1i[3] | u.trans_1a[2] | u.trans_1a[3] | u.trans_1t[2] | u.trans_1t[3];
        any |= u.trans_2i[2] | u.trans_2i[3] | u.trans_1i_1a[2] | u.trans_1i_1a[3] | u.trans_1i_1t[2] | u.trans_1i_1t[3];
      }
      break;
    case CAAA_PH_MOVE_SUBS:
      for (int s = 0; s < NS; s++) any |= st.units_sea[s].submarines[2];
      break;
    case CAAA_PH_MOVE_SHIPS:
      for (int s = 0; s < NS; s++) {
        const CaaaUnitsSea& u = st.units_sea[s];
        any |= u.destroyers[1] | u.carriers[1] | u.cruisers[2] | u.battleships[2] | u.bs_damaged[2];
      }
      break;
    case CAAA_PH_SEA_BATTLES:
      for (int s = 0; s < NS; s++) any |= (st.flagged_for_combat[NL + s] != 0 && g->tot_sea[0][s] != 0);
      break;
    case CAAA_PH_UNLOAD_TRANSPORTS:
      for (int s = 0; s < NS; s++) {
        const CaaaUnitsSea& u = st.units_sea[s];
        any |= u.trans_1i[1] | u.trans_1a[1] | u.trans_1t[1] | u.trans_2i[1] | u.trans_1i_1a[1] | u.trans_1i_1t[1];
      }
      break;
    case CAAA_PH_LAND_BATTLES:
      for (int l = 0; l < NL; l++) any |= (st.flagged_for_combat[l] != 0 && g->tot_land[0][l] != 0);
      break;
    case CAAA_PH_MOVE_AA_GUNS:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].aa_guns[1];
      break;
    case CAAA_PH_LAND_FIGHTERS:
      for (int l = 0; l < NL; l++) any |= st.land_state[l].fighters[1] | st.land_state[l].fighters[2] | st.land_state[l].fighters[3];
      for (int s = 0; s < NS; s++) any |= st.units_sea[s].fighters[1] | st.units_sea[s].fighters[2] | st.units_sea[s].fighters[3];
      break;
    case CAAA_PH_LAND_BOMBERS:
      for (int l = 0; l < NL; l++)
        for (int k = 1; k <= 5; k++) any |= st.land_state[l].bombers[k];
      for (int s = 0; s < NS; s++)
        for (int k = 0; k <= 4; k++) any |= st.units_sea[s].bombers[k];
      break;
    default:
      return true;
  }
  return any != 0;
}
// first phase >= from that has work; CAAA_PH_BUY_UNITS (the end-of-turn item) if none
CAAA_HD int next_phase_with_work(const GameSlot* g, int from) {
  for (int ph = from; ph < CAAA_PH_BUY_UNITS; ph++)
    if (phase_has_work(g, ph)) return ph;
  return CAAA_PH_BUY_UNITS;
}

}  // namespace caaa
