// engine2.cuh -- the rules engine that one GPU thread runs on one Game (game.cuh) in shared memory.
//
// B200-native restatement of the reference's turn pipeline (reference src/engine.cpp:642-814 and
// 1291-4337): 21 phases per player-turn, their derived caches, volley-based combat and scoring.
// Draw order inside a game is sequential by definition, so all parallelism is across games: the 32
// lanes of a warp run the same phase function on 32 games.  The code is therefore written for SIMT
// convergence rather than for a scalar core:
//   * every phase is ONE flat loop with a per-lane cursor ("which source / unit class am I at") instead
//     of the reference's nested for/while loops, so lanes whose games have units in different
//     territories execute the same decision step together instead of waiting for each other;
//   * a per-game work mask says which phases can have anything to do (set when a unit enters a state a
//     later phase moves, rebuilt at the end of the turn), so a warp skips a phase no lane needs;
//   * per-player data is stored in absolute player order: the reference's end-of-turn rotation of every
//     per-player array is an index increment here;
//   * scoring and the cache sums use 4-byte dot products (DP4A) over the per-player unit vectors.
//
// Reference quirks are reproduced deliberately; layout-dependent
// out-of-bounds reads follow the canonical meaning of patch set P1.
//
// The functions are __host__ __device__ so that tests/hostsim can run the *same* code on the CPU
// against the oracle when no GPU is present; the product only ever launches them in kernels.
#pragma once
#include "game.cuh"

namespace caaa {
namespace e2 {

CAAA_HD constexpr uint32_t W(int ph) { return 1u << ph; }
constexpr uint32_t W_ALL_MOVES = 0x7fffu;  // phases 0..14

CAAA_HD_NOINLINE void refill_rng(Game* g, const RunCtx cx, uint32_t d) {
  philox4x32_10(d >> 4, g->game_lo, g->game_hi, 0u, cx.seed_lo, cx.seed_hi, g->rng);
}
CAAA_HD uint32_t next_byte(Game* g, const RunCtx cx) {
  const uint32_t d = g->draws++;
  const uint32_t k = d & 15u;
  if (k == 0) refill_rng(g, cx, d);  // one Philox block serves 16 draws
  return (g->rng[k >> 2] >> (8u * (k & 3u))) & 0xffu;
}
// byte % n for n in 1..31 without a division: q = byte * ceil(2^16 / n) >> 16 is exact for byte < 256
CAAA_HD uint32_t mod_small(const RunCtx cx, uint32_t byte, uint32_t n) {
  const uint32_t q = (byte * CAAA_TAB(cx)->recip16[n & 31u]) >> 16;
  return byte - q * n;
}

// next non-empty cell of a phase's cell list at or after `cell` (DevTables::scan); `fresh` is raised when the
// cursor moves on, exactly where the reference's for-loops would start a new (type, location, state) block
CAAA_HD uint32_t scan_cells(const Game* g, const uint32_t* tab, int n_cells, int& cell, bool& fresh) {
  const uint8_t* gb = reinterpret_cast<const uint8_t*>(g);
  uint32_t c = 0;
#pragma unroll 1
  for (; cell < n_cells; cell++) {  // (a 4-wide variant with independent loads was measured: 4 % slower)
    c = tab[cell];
    if (gb[c & 1023u] != 0) break;
    fresh = true;
  }
  return c;
}
CAAA_HD int cell_type(uint32_t c) { return (int)((c >> 10) & 15u); }
CAAA_HD int cell_loc(uint32_t c) { return (int)((c >> 14) & 7u); }
CAAA_HD int cell_state(uint32_t c) { return (int)((c >> 17) & 7u); }
CAAA_HD uint8_t* cell_arr(Game* g, uint32_t c) { return reinterpret_cast<uint8_t*>(g) + (c & 1023u) - cell_state(c); }

// ---- cache refresh, engine.cpp:642-814 ----
CAAA_HD_NOINLINE void refresh_eot_cache(Game* g, const RunCtx cx) {  // 653-660 (refresh_allies is tabulated)
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur, ne = CAAA_TAB(cx)->n_enemies[cur];
  const uint8_t* en = CAAA_TAB(cx)->enemies_abs[cur];
  uint32_t cs = 0;  // refresh_canals 726-733 (the canal-specific table copies are read through canal_state)
#pragma unroll
  for (int k = 0; k < CAAA_CANALS; k++)
    if (owner_allied(g, cx, t.canal_lands[k][0]) && owner_allied(g, cx, t.canal_lands[k][1])) cs += 1u << k;
  g->canal_state = (uint8_t)cs;
#pragma unroll 1
  for (int l = 0; l < NL; l++) {  // refresh_enemy_armies 752-760
    int n = 0;
#pragma unroll 1
    for (int e = 0; e < ne; e++) n += g->tot_land[en[e]][l];
    g->enemy_count[l] = (int16_t)n;
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {  // refresh_fleets 761-784
    const uint8_t* mine = uts(g, cur, s);
    int carriers = 0, ecount = 0;
    uint32_t edest = 0, block = 0;
#pragma unroll 1
    for (int pa = 0; pa < NP; pa++) {
      const uint8_t* su = uts(g, pa, s);
      if (allied(g, cx, pa)) {
        carriers += su[CARRIERS];
      } else {
        ecount += g->tot_sea[pa][s];
        edest += su[DESTROYERS];
        block += su[DESTROYERS] + su[CARRIERS] + su[CRUISERS] + su[BATTLESHIPS] + su[BS_DAMAGED];
      }
    }
    g->carriers[s] = (int16_t)carriers;
    g->enemy_count[s + NL] = (int16_t)ecount;
    g->edestroyers[s] = (uint8_t)edest;
    g->blockade[s] = (uint8_t)block;
    g->large_cargo[s] = (int16_t)(mine[TRANS_EMPTY] + mine[TRANS_1I]);
    g->small_cargo[s] = (int16_t)(mine[TRANS_EMPTY] + mine[TRANS_1I] + mine[TRANS_1A] + mine[TRANS_1T]);
  }
  uint32_t lb = g->land_blocked;  // refresh_*_path_blocked 785-814: only the listed entries are ever written
#pragma unroll 1
  for (int s = 0; s < NL; s++)
#pragma unroll 1
    for (int i = 0; i < t.lands_within_2_count[s]; i++) {
      const int d = t.lands_within_2[s][i];
      const int n1 = t.land_path[s][d], n2 = t.land_path_alt[s][d];
      const bool b = (g->enemy_count[n1] > 0 || g->factory_max[n1] > 0) && (g->enemy_count[n2] > 0 || g->factory_max[n2] > 0);
      const uint32_t bit = 1u << (s * NL + d);
      lb = b ? (lb | bit) : (lb & ~bit);
    }
  g->land_blocked = lb;
  uint32_t sb = g->seasub_blocked;
#pragma unroll 1
  for (int s = 0; s < NS; s++)
#pragma unroll 1
    for (int i = 0; i < t.seas_within_2_count[cs][s]; i++) {
      const int d = t.seas_within_2[cs][s][i];
      const int n1 = t.sea_path[cs][s][d], n2 = t.sea_path_alt[cs][s][d];
      const uint32_t bit = 1u << (s * NS + d);
      sb = (g->blockade[n1] > 0 && g->blockade[n2] > 0) ? (sb | bit) : (sb & ~bit);
      sb = (g->edestroyers[n1] > 0 && g->edestroyers[n2] > 0) ? (sb | (bit << 16)) : (sb & ~(bit << 16));
    }
  g->seasub_blocked = sb;
}

CAAA_HD_NOINLINE void refresh_full_cache(Game* g, const RunCtx cx) {  // 642-652, 661-715
  // Runs once per imported state (and per battle in the battle kernels): the per-state counts of a location are read
  // into registers in one go and summed with constant offsets, instead of one dependent byte load per loop trip.
  const int cur = g->cur;
#pragma unroll
  for (int p = 0; p < NP; p++) {
    g->income[p] = 0;
    g->nfact[p] = 0;
  }
#pragma unroll 1
  for (int l = 0; l < NL; l++) {
    const int owner = g->owner[l];
    if (g->factory_max[l] > 0) g->factloc[owner][g->nfact[owner]++] = (uint8_t)l;
    g->income[owner] = (int16_t)(g->income[owner] + CAAA_TAB(cx)->t.land_value[l]);
    uint8_t c[LU_ROW];
#pragma unroll
    for (int k = 0; k < LU_ROW; k++) c[k] = g->lu[l][k];
    uint8_t* mine = utl(g, cur, l);
#pragma unroll
    for (int u = 0; u < NLT; u++) {
      uint32_t n = 0;
#pragma unroll
      for (int k = 0; k < (int)states_land(u); k++) n += c[off_lu(u) + k];
      mine[u] = (uint8_t)n;
    }
    uint8_t r[NP][NLT];
#pragma unroll
    for (int pa = 0; pa < NP; pa++)
#pragma unroll
      for (int u = 0; u < NLT; u++) r[pa][u] = utl(g, pa, l)[u];
#pragma unroll
    for (int pa = 0; pa < NP; pa++) {
      int n = 0;
#pragma unroll
      for (int u = 0; u < NLT; u++) n += r[pa][u];
      g->tot_land[pa][l] = (int16_t)n;
    }
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    uint8_t c[SU_ROW];
#pragma unroll
    for (int k = 0; k < SU_ROW; k++) c[k] = g->su[s][k];
    uint8_t* mine = uts(g, cur, s);
#pragma unroll
    for (int u = 0; u < NST; u++) {
      uint32_t n = 0;
#pragma unroll
      for (int k = 0; k < (int)states_sea(u); k++) n += c[off_sea(u) + k];
      if (u == BOMBERS_SEA) g->cur_sea_bombers[s] = (uint8_t)n;
      else mine[u] = (uint8_t)n;
    }
#pragma unroll 1
    for (int pa = 0; pa < NP; pa++) {
      const uint8_t* q = uts(g, pa, s);
      uint8_t r[NST - 1];
#pragma unroll
      for (int u = 0; u < NST - 1; u++) r[u] = q[u];
      int n = 0;
#pragma unroll
      for (int u = 0; u < NST - 1; u++) n += r[u];
      g->tot_sea[pa][s] = (int16_t)n;
    }
  }
  refresh_eot_cache(g, cx);
}

// ---- decisions and dice ----
template <bool EXPAND>
CAAA_HD uint8_t ai_input(Game* g, const RunCtx cx) {  // getAIInput 1291-1297
  g->answers_remaining--;
  if (EXPAND && g->use_selected) return g->selected_action;
  const uint32_t n = g->nvm;
  const uint32_t b = next_byte(g, cx);
  return g->vm[n < 32u ? mod_small(cx, b, n) : b % n];
}
CAAA_HD bool unlucky_team(const Game* g, const RunCtx cx) { return CAAA_TAB(cx)->t.is_allied[g->cur][g->unlucky % NP] != 0; }
CAAA_HD uint32_t attacker_hits(Game* g, const RunCtx cx, int dmg) {  // 2552-2565
  if (g->answers_remaining < 2) return (uint8_t)(unlucky_team(g, cx) ? dmg / 6 : dmg / 6 + (1 < dmg % 6 ? 1 : 0));
  return (uint8_t)(dmg / 6 + ((int)(next_byte(g, cx) % 6) < dmg % 6 ? 1 : 0));
}
CAAA_HD uint32_t defender_hits(Game* g, const RunCtx cx, int dmg) {  // 2567-2580
  if (g->answers_remaining < 2) return (uint8_t)(unlucky_team(g, cx) ? dmg / 6 + (1 < dmg % 6 ? 1 : 0) : dmg / 6);
  return (uint8_t)(dmg / 6 + ((int)(next_byte(g, cx) % 6) < dmg % 6 ? 1 : 0));
}

CAAA_HD void update_move_history_4air(Game* g, uint32_t src, uint32_t dst) {  // 1300-1328: starts one past the list
#pragma unroll 1
  for (;;) {
    const uint32_t va = g->nvm < CAAA_ACTIONS ? g->vm[g->nvm] : 0u;
    if (va == dst) break;
    skip_set(g, src, va);
    const uint32_t row = skip_row(g, va);  // apply_skip: everything skipped from `va` is skipped from `src` as well
    g->skipped[src >> 2] |= row << ((src & 3u) * 8u);
    g->nvm--;
  }
}

CAAA_HD_NOINLINE bool load_transport(Game* g, int ut, int src_land, int dst_sea, int land_state) {  // 1397-1443
#pragma unroll 1
  for (int tt = weight_land(ut) > 2 ? TRANS_1I : TRANS_1T; tt >= TRANS_EMPTY; tt--) {
    uint8_t* arr = su_ptr(g, dst_sea, tt);
    const int unl = unloading_sea(tt);
#pragma unroll 1
    for (int ts = (int)states_sea(tt) - (int)staging_sea(tt); ts >= unl; ts--) {
      if (arr[ts] == 0) continue;
      const int ntt = load_result(ut, tt);
      arr[ts]--;
      if (tt == TRANS_EMPTY && ts == 0) ts = unloading_sea(ntt);
      su_ptr(g, dst_sea, ntt)[ts]++;
      uint8_t* mine = uts(g, g->cur, dst_sea);
      mine[tt]--;
      mine[ntt]++;
      g->tot_land[g->cur][src_land]--;
      utl(g, g->cur, src_land)[ut]--;
      lu_ptr(g, src_land, ut)[land_state]--;
      g->large_cargo[dst_sea] = (int16_t)(mine[TRANS_EMPTY] + mine[TRANS_1I]);
      g->small_cargo[dst_sea] = (int16_t)(g->large_cargo[dst_sea] + mine[TRANS_1A] + mine[TRANS_1T]);
      g->work |= W(CAAA_PH_MOVE_TRANSPORTS) | W(CAAA_PH_UNLOAD_TRANSPORTS);
      return true;
    }
  }
  return false;
}

CAAA_HD_NOINLINE void add_valid_land_moves(Game* g, const RunCtx cx, int src, int moves, int ut) {  // 1445-1516
  const CaaaTables& t = CAAA_TAB(cx)->t;
  uint32_t n = g->nvm;
  if (moves == 2) {
#pragma unroll 1
    for (int i = 0; i < t.lands_within_2_count[src]; i++) {
      const int d = t.lands_within_2[src][i];
      if (skip_get(g, src, d)) continue;
      if (t.land_dist[src][d] == 2 && land_blocked(g, src * NL + d)) continue;
      g->vm[n++] = (uint8_t)d;
    }
#pragma unroll 1
    for (int i = 0; i < t.load_within_2_count[src]; i++) {
      const int ds = t.load_within_2[src][i];
      if (g->large_cargo[ds] == 0) continue;
      const int da = ds + NL;
      if (skip_get(g, src, da)) continue;
      // 1474: air index into a 5-wide row (P1 iii): flat index, out of range reads 0
      if (t.land_dist[src][da] == 1 && land_blocked(g, (unsigned)(src * NL + da))) continue;
      g->vm[n++] = (uint8_t)da;
    }
    g->nvm = (uint8_t)n;
    return;
  }
  const bool non_combat = ut == AA_GUNS, unloadable = weight_land(ut) > 5, heavy = weight_land(ut) > 2;
#pragma unroll 1
  for (int i = 0; i < t.l2l_count[src]; i++) {
    const int d = t.l2l_conn[src][i];
    if (skip_get(g, src, d)) continue;
    if (non_combat && !owner_allied(g, cx, d)) continue;
    g->vm[n++] = (uint8_t)d;
  }
  if (!unloadable)
#pragma unroll 1
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

// add_valid_sea_moves / add_valid_sub_moves 1518-1586; `blocked_shift` selects the sea (0) or sub (16) flags
CAAA_HD void add_valid_sea_moves(Game* g, const RunCtx cx, int src_sea, int moves, uint32_t blocked_shift) {
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cs = g->canal_state, sa = src_sea + NL;
  uint32_t n = g->nvm;
  const bool two = moves == 2;
  const int cnt = two ? t.seas_within_2_count[cs][src_sea] : t.seas_within_1_count[cs][src_sea];
#pragma unroll 1
  for (int i = 0; i < cnt; i++) {
    const int d = two ? t.seas_within_2[cs][src_sea][i] : t.seas_within_1[cs][src_sea][i];
    if (two && ((g->seasub_blocked >> (blocked_shift + src_sea * NS + d)) & 1u)) continue;
    if (skip_get(g, sa, d + NL)) continue;
    g->vm[n++] = (uint8_t)(d + NL);
  }
  g->nvm = (uint8_t)n;
}

// ---- conquer_land, 1918-1984 ----
CAAA_HD_NOINLINE void conquer_land(Game* g, const RunCtx cx, int dst) {
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  const int old = g->owner[dst];
  if (t.capital[old] == dst) {
    g->money[cur] = (uint8_t)(g->money[cur] + g->money[old]);
    g->money[old] = 0;
  }
  g->income[old] = (int16_t)(g->income[old] - t.land_value[dst]);
  const int orig = t.original_owner[dst];
  const int nw = allied(g, cx, orig) ? orig : cur;
  g->owner[dst] = (uint8_t)nw;
  g->income[nw] = (int16_t)(g->income[nw] + t.land_value[dst]);
  g->factloc[nw][g->nfact[nw]++] = (uint8_t)dst;
  g->nfact[old]--;
  const int n_old = g->nfact[old];
  uint8_t* fl = &g->factloc[0][0];  // flat like the reference's int[6][5]: the shift may read the next row's first entry
#pragma unroll 1
  for (int i = 0; i < n_old; i++)
    if (g->factloc[old][i] == dst) {
#pragma unroll 1
      for (int j = i; j < n_old; j++) fl[old * NL + j] = fl[old * NL + j + 1];
      break;
    }
}

// ---- phase 0: move_fighter_units, 1657-1803, 3385-3407 ----
// returns bit masks (bit a = air location a): lo byte canFighterLandHere, next byte canFighterLandIn1Move
CAAA_HD_NOINLINE uint32_t fighter_landing_masks(Game* g, const RunCtx cx) {  // pre_move_fighter_units 1657-1714
  const CaaaTables& t = CAAA_TAB(cx)->t;
  uint32_t here = 0;
#pragma unroll 1
  for (int l = 0; l < NL; l++) {
    if (owner_allied(g, cx, l) && g->flagged[l] == 0) here |= 1u << l;
    if (owner_rel(g, l) == g->player && g->factory_max[l] > 0)  // 1676: relative owner vs absolute player
#pragma unroll 1
      for (int i = 0; i < t.l2s_count[l]; i++) here |= 1u << (NL + t.l2s_conn[l][i]);
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    if (g->carriers[s] <= 0) continue;
    here |= 1u << (s + NL);
    if (su_ptr(g, s, CARRIERS)[2] > 0)  // 1688: carriers[2] aliases cruisers[0]
#pragma unroll 1
      for (int i = 0; i < t.s2s_count[s]; i++) {
        const int s1 = t.s2s_conn[s][i];
        here |= 1u << (NL + s1);
#pragma unroll 1
        for (int j = 0; j < t.s2s_count[s1]; j++) here |= 1u << (NL + t.s2s_conn[s1][j]);
      }
  }
  uint32_t in1 = 0;
#pragma unroll 1
  for (int a = 0; a < NA; a++)
#pragma unroll 1
    for (int i = 0; i < t.air_conn_count[a]; i++)
      if ((here >> t.air_conn[a][i]) & 1) {
        in1 |= 1u << a;
        break;
      }
  return here | (in1 << 8);
}

template <bool EXPAND>
CAAA_HD_NOINLINE int move_fighter_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_MOVE_FIGHTERS)) != 0);
  if (!(g->work & W(CAAA_PH_MOVE_FIGHTERS))) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  uint32_t masks = 0;
  int src = 0;
  if (g->cur_phase == CAAA_PH_MOVE_FIGHTERS) {
    src = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
    masks = g->cur_b;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_FIGHTERS, SC_FIGHTERS_N, src, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    converge<!EXPAND>(mact);
    if (src >= NA) {
      active = false;
      continue;
    }
    if (fresh) {
      fresh = false;
      if (!processed) {
        processed = true;
        masks = fighter_landing_masks(g, cx);
      }
      g->vm[0] = (uint8_t)src;
      uint32_t n = 1;
#pragma unroll 1
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
    }
    uint32_t dst = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    update_move_history_4air(g, (uint32_t)src, dst);
    if ((uint32_t)src == dst) {
      arr[0] = (uint8_t)(arr[0] + arr[4]);
      arr[4] = 0;
      continue;
    }
    int dist = t.air_dist[src][dst];
    if (dst < (uint32_t)NL) {
      if (!owner_allied(g, cx, dst)) flag_combat(g, dst, 2);
      else dist = 4;  // friendly rebase uses up all moves
      lu_ptr(g, dst, FIGHTERS)[(uint8_t)(4 - dist)]++;
      utl(g, cur, dst)[FIGHTERS]++;
      g->tot_land[cur][dst]++;
    } else {
      flag_combat(g, dst, g->enemy_count[dst] > 0);
      su_ptr(g, dst - NL, FIGHTERS)[(uint8_t)(4 - dist)]++;
      uts(g, cur, dst - NL)[FIGHTERS]++;
      g->tot_sea[cur][dst - NL]++;
    }
    if (src < NL) {
      utl(g, cur, src)[FIGHTERS]--;
      g->tot_land[cur][src]--;
    } else {
      uts(g, cur, src - NL)[FIGHTERS]--;
      g->tot_sea[cur][src - NL]--;
    }
    arr[4]--;
    g->work |= W(CAAA_PH_LAND_FIGHTERS);
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_MOVE_FIGHTERS, src, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, masks);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_MOVE_FIGHTERS);
  return PH_DONE;
}

// ---- phase 1: move_bomber_units, 1805-1916, 3423-3443 ----
CAAA_HD uint32_t bomber_here_mask(const Game* g, const RunCtx cx) {  // 1808-1812, 3557-3563 (lands only)
  uint32_t here = 0;
#pragma unroll 1
  for (int l = 0; l < NL; l++)
    if (owner_allied(g, cx, l) && g->flagged[l] == 0) here |= 1u << l;
  return here;
}
template <bool EXPAND>
CAAA_HD_NOINLINE int move_bomber_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m0 = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_MOVE_BOMBERS)) != 0);
  if (!(g->work & W(CAAA_PH_MOVE_BOMBERS))) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  // the landing masks are phase-local here (the reference keeps them in globals, but re-derives them
  // before every use), so nothing happens without unmoved bombers
  bool any = false;
#pragma unroll 1
  for (int l = 0; l < NL; l++) any |= lu_ptr(g, l, BOMBERS_LAND_AIR)[6] != 0;
  const unsigned m = vote_ballot<!EXPAND>(m0, any);
  if (!any) {
    g->work &= ~W(CAAA_PH_MOVE_BOMBERS);
    return PH_DONE;
  }
  uint32_t here = 0, in1 = 0, in2 = 0;
  bool fresh = true;
  int src = 0;
  if (g->cur_phase == CAAA_PH_MOVE_BOMBERS) {
    src = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    here = g->cur_b & 0xffu;
    in1 = (g->cur_b >> 8) & 0xffu;
    in2 = (g->cur_b >> 16) & 0xffu;
  } else {
    here = bomber_here_mask(g, cx);
  #pragma unroll 1
    for (int l = 0; l < NL; l++)
  #pragma unroll 1
      for (int i = 0; i < t.l2l_count[l]; i++)
        if ((here >> t.l2l_conn[l][i]) & 1) { in1 |= 1u << l; break; }
  #pragma unroll 1
    for (int s = 0; s < NS; s++)
  #pragma unroll 1
      for (int i = 0; i < t.s2l_count[s]; i++)
        if ((here >> t.s2l_conn[s][i]) & 1) { in1 |= 1u << (NL + s); break; }
  #pragma unroll 1
    for (int a = 0; a < NA; a++)
  #pragma unroll 1
      for (int i = 0; i < t.air_conn_count[a]; i++)
        if ((in1 >> t.air_conn[a][i]) & 1) { in2 |= 1u << a; break; }
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_BOMBERS, SC_BOMBERS_N, src, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    converge<!EXPAND>(mact);
    if (src >= NL) {
      active = false;
      continue;
    }
    if (fresh) {
      fresh = false;
      g->vm[0] = (uint8_t)src;
      uint32_t n = 1;
#pragma unroll 1
      for (int i = 0; i < t.air_within_count[5][src]; i++) {  // add_valid_bomber_moves(src, 6)
        const int d = t.air_within[5][src][i];
        const int dist = t.air_dist[src][d];
        const bool h = (here >> d) & 1;
        if (dist <= 3 || h || (dist == 4 && ((in2 >> d) & 1)) || (dist == 5 && ((in1 >> d) & 1))) {
          if (!h && g->enemy_count[d] == 0) {
            if (d >= NL || g->factory_max[d] == 0 || g->factory_hp[d] == -(int)g->factory_max[d]) continue;
          }
          if (!skip_get(g, src, d)) g->vm[n++] = (uint8_t)d;
        }
      }
      g->nvm = (uint8_t)n;
    }
    uint32_t dst = g->vm[0];
    if (g->nvm == 1) {  // 1860: inverted test
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    if ((uint32_t)src == dst) {
      arr[0] = (uint8_t)(arr[0] + arr[6]);
      arr[6] = 0;
      continue;
    }
    const int dist = t.air_dist[src][dst];
    if (dst < (uint32_t)NL) {
      if (!owner_allied(g, cx, dst)) flag_combat(g, dst, 2);
      lu_ptr(g, dst, BOMBERS_LAND_AIR)[(uint8_t)(6 - dist)]++;
      utl(g, cur, dst)[BOMBERS_LAND_AIR]++;
      g->tot_land[cur][dst]++;
    } else {
      flag_combat(g, dst, g->enemy_count[dst] > 0);
      su_ptr(g, dst - NL, BOMBERS_SEA)[(uint8_t)(5 - dist)]++;
      g->cur_sea_bombers[dst - NL]++;
      g->tot_sea[cur][dst - NL]++;
    }
    utl(g, cur, src)[BOMBERS_LAND_AIR]--;
    g->tot_land[cur][src]--;
    arr[6]--;
    g->work |= W(CAAA_PH_LAND_BOMBERS);
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_MOVE_BOMBERS, src, fresh ? CF_FRESH : 0u, 0u, here | (in1 << 8) | (in2 << 16));
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  clear_move_history(g);
  g->work &= ~W(CAAA_PH_MOVE_BOMBERS);
  return PH_DONE;
}

// ---- phase 2: stage_transport_units, 1588-1655 ----
template <bool EXPAND>
CAAA_HD_NOINLINE int stage_transport_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_STAGE_TRANSPORTS)) != 0);
  if (!(g->work & W(CAAA_PH_STAGE_TRANSPORTS))) return PH_DONE;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int cell = 0;  // (ut - TRANS_EMPTY) * NS + sea, in the reference's loop order
  if (g->cur_phase == CAAA_PH_STAGE_TRANSPORTS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_STAGE, SC_STAGE_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    const int ut = cell_type(cc), s = cell_loc(cc), staging = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= 4 * NS) {
      active = false;
      continue;
    }
    const int done = staging - 1;
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      g->vm[0] = (uint8_t)sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, 2, 0);
    }
    processed = true;
    uint32_t da = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      da = ai_input<EXPAND>(g, cx);
    }
    g->work |= W(CAAA_PH_MOVE_TRANSPORTS) | W(CAAA_PH_UNLOAD_TRANSPORTS);
    if (sa == da) {
      arr[done] = (uint8_t)(arr[done] + arr[staging]);
      arr[staging] = 0;
      continue;
    }
    const int ds = (int)da - NL;
    const unsigned flat = (unsigned)(s * NS) + da;  // 1630: sea_dist indexed with an air id (P1 iii)
    const int dist = flat < 9u ? ((const uint8_t*)CAAA_TAB(cx)->t.sea_dist[g->canal_state])[flat] : 0;
    if (g->blockade[ds] > 0) {
      flag_combat(g, da, 2);
      continue;  // not moved: the choice is drawn again
    }
    su_ptr(g, ds, ut)[(uint8_t)(done - dist)]++;
    uts(g, cur, ds)[ut]++;
    g->tot_sea[cur][ds]++;
    g->small_cargo[ds]++;
    arr[staging]--;
    uts(g, cur, s)[ut]--;
    g->tot_sea[cur][s]--;
    g->small_cargo[s]--;
    if (ut <= TRANS_1I) {
      g->large_cargo[s]--;
      g->large_cargo[ds]++;
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_STAGE_TRANSPORTS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_STAGE_TRANSPORTS);
  return PH_DONE;
}

// ---- phases 3,4,5,12: move_land_unit_type, 1986-2060 ----
// Walks the land-unit phases ph_first..ph_last of the pipeline in ONE flat loop (3..5 = tanks, artillery, infantry are
// consecutive in the reference's turn, src/engine.cpp:4217-4219; 12 = AA guns): a lane whose game only has infantry to
// move and a lane that is still moving tanks execute the same decision step together, and a game needs one visit
// instead of up to three.  Per game the order of cells, lists and draws is exactly that of the separate phases; the
// end-of-phase bookkeeping runs when a lane's cursor leaves a phase's cells.
template <bool EXPAND>
CAAA_HD_NOINLINE int move_land_units(Game* g, const RunCtx cx, unsigned min, int ph_first, int ph_last, int yield = 0) {
  const uint32_t range = ((2u << ph_last) - 1u) & ~((1u << ph_first) - 1u);
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & range) != 0);
  if (!(g->work & range)) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int ph = ph_first;
  int cell = 0;  // within the phase: src * max_move + (max_move - moves)
  if (g->cur_phase >= ph_first && g->cur_phase <= ph_last) {
    ph = g->cur_phase;
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const int ut = ph == CAAA_PH_MOVE_TANKS ? (int)TANKS : ph == CAAA_PH_MOVE_ARTILLERY ? (int)ARTILLERY : ph == CAAA_PH_MOVE_INFANTRY ? (int)INFANTRY : (int)AA_GUNS;
    const int sc_first = ph == CAAA_PH_MOVE_TANKS ? (int)SC_TANKS : ph == CAAA_PH_MOVE_ARTILLERY ? (int)SC_ARTILLERY : ph == CAAA_PH_MOVE_INFANTRY ? (int)SC_INFANTRY : (int)SC_AA_GUNS;
    const int n_cells = ph == CAAA_PH_MOVE_TANKS ? 2 * NL : NL;
    const bool has_phase = (g->work & W(ph)) != 0;
    uint32_t cc = 0;
    if (has_phase) cc = scan_cells(g, CAAA_TAB(cx)->scan + sc_first, n_cells, cell, fresh);
    else cell = n_cells;
    uint8_t* const arr = cell_arr(g, cc);
    const int src = cell_loc(cc), moves = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= n_cells) {  // this phase is done for this game: its end-of-phase bookkeeping, then the next phase's cells
      if (has_phase) {
        if (processed) clear_move_history(g);
        g->work &= ~W(ph);
      }
      processed = false;
      fresh = true;
      cell = 0;
      if (ph == ph_last) active = false;
      else ph++;
      continue;
    }
    if (fresh) {
      fresh = false;
      processed = true;
      g->vm[0] = (uint8_t)src;
      g->nvm = 1;
      add_valid_land_moves(g, cx, src, moves, ut);
    }
    uint32_t dst = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    if ((uint32_t)src == dst) {
      arr[0] = (uint8_t)(arr[0] + arr[moves]);
      arr[moves] = 0;
      continue;
    }
    if (dst >= (uint32_t)NL) {
      if (!load_transport(g, ut, src, (int)dst - NL, moves)) {
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
    flag_combat(g, dst, enemies_there);
    int dist = t.land_dist[src][dst];
    if (friendly || enemies_there) dist = moves;  // not blitzable: the unit's turn ends
    lu_ptr(g, dst, ut)[(uint8_t)(moves - dist)]++;
    utl(g, cur, dst)[ut]++;
    g->tot_land[cur][dst]++;
    utl(g, cur, src)[ut]--;
    g->tot_land[cur][src]--;
    arr[moves]--;
    if (!friendly && !enemies_there) {
      conquer_land(g, cx, dst);
      flag_combat(g, dst, 2);
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, ph, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  return PH_DONE;
}
// one land-unit phase (reference move_land_unit_type, src/engine.cpp:1986-2060)
template <bool EXPAND>
CAAA_HD int move_land_unit_type(Game* g, const RunCtx cx, unsigned min, int ut, int ph, int yield = 0) {
  (void)ut;
  return move_land_units<EXPAND>(g, cx, min, ph, ph, yield);
}

// ---- phase 6: move_transport_units, 2062-2159 ----
template <bool EXPAND>
CAAA_HD_NOINLINE int move_transport_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_MOVE_TRANSPORTS)) != 0);
  if (!(g->work & W(CAAA_PH_MOVE_TRANSPORTS))) return PH_DONE;
  const int cur = g->cur;
  if (g->cur_phase != CAAA_PH_MOVE_TRANSPORTS)
#pragma unroll 1
  for (int s = 0; s < NS; s++) {  // skip_empty_transports 2062-2075
    uint8_t* e = su_ptr(g, s, TRANS_EMPTY);
    const uint32_t n = e[1] + e[2] + e[3];
    if (n) {
      e[0] = (uint8_t)(e[0] + n);
      e[1] = e[2] = e[3] = 0;
    }
  }
  bool processed = false, fresh = true;
  int cell = 0;  // ((ut - TRANS_1I) * NS + sea) * 2 + (i - 1); every loaded type has max_state 4
  if (g->cur_phase == CAAA_PH_MOVE_TRANSPORTS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_TRANSPORTS, SC_TRANSPORTS_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    const int ut = cell_type(cc), s = cell_loc(cc), cs = cell_state(cc);  // cs = max_state - i, i = 1, 2
    converge<!EXPAND>(mact);
    if (cell >= 6 * NS * 2) {
      active = false;
      continue;
    }
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      processed = true;
      g->vm[0] = (uint8_t)sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, cs - 1, 0);
    }
    uint32_t da = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      da = ai_input<EXPAND>(g, cx);
    }
    const int ds = (int)da - NL;
    if (g->blockade[ds] > 0) flag_combat(g, da, 2);
    g->work |= W(CAAA_PH_UNLOAD_TRANSPORTS);
    if (sa == da) {
      arr[1] = (uint8_t)(arr[1] + arr[cs]);
      arr[cs] = 0;
      continue;
    }
    su_ptr(g, ds, ut)[1]++;
    uts(g, cur, ds)[ut]++;
    g->tot_sea[cur][ds]++;
    arr[cs]--;
    uts(g, cur, s)[ut]--;
    g->tot_sea[cur][s]--;
    if (ut <= TRANS_1T) {
      g->small_cargo[ds]++;
      g->small_cargo[s]--;
      if (ut <= TRANS_1I) {
        g->large_cargo[ds]++;
        g->large_cargo[s]--;
      }
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_MOVE_TRANSPORTS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_MOVE_TRANSPORTS);
  return PH_DONE;
}

// ---- phase 7: move_subs, 2160-2214 ----
template <bool EXPAND>
CAAA_HD_NOINLINE int move_subs(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_MOVE_SUBS)) != 0);
  if (!(g->work & W(CAAA_PH_MOVE_SUBS))) return PH_DONE;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int s = 0;
  if (g->cur_phase == CAAA_PH_MOVE_SUBS) {
    s = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    uint8_t* arr = nullptr;
#pragma unroll 1
    for (; s < NS; s++) {
      arr = su_ptr(g, s, SUBMARINES);
      if (arr[2] != 0) break;
      fresh = true;
    }
    converge<!EXPAND>(mact);
    if (s >= NS) {
      active = false;
      continue;
    }
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      g->vm[0] = (uint8_t)sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, 2, 16);
    }
    processed = true;
    uint32_t da = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      da = ai_input<EXPAND>(g, cx);
    }
    const int ds = (int)da - NL;
    if (g->enemy_count[ds] > 0) flag_combat(g, ds, 2);  // 2188,2194: sea index used as an air index
    if (sa == da) {
      arr[0] = (uint8_t)(arr[0] + arr[2]);
      arr[2] = 0;
      continue;
    }
    su_ptr(g, ds, SUBMARINES)[0]++;
    uts(g, cur, ds)[SUBMARINES]++;
    g->tot_sea[cur][ds]++;
    arr[2]--;
    uts(g, cur, s)[SUBMARINES]--;
    g->tot_sea[cur][s]--;
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_MOVE_SUBS, s, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_MOVE_SUBS);
  return PH_DONE;
}

// ---- phase 8: move_destroyers_battleships, 2216-2310 ----
CAAA_HD void carry_allied_fighters(Game* g, const RunCtx cx, int src, int dst) {  // 2286-2310
  int moved = 0;
#pragma unroll 1
  for (int p = 1; p < NP; p++) {
    const int pa = (g->cur + p) % NP;
    if (!allied(g, cx, pa)) continue;
#pragma unroll 1
    while (uts(g, pa, src)[FIGHTERS] > 0) {
      uts(g, pa, src)[FIGHTERS]--;
      g->tot_sea[pa][src]--;
      uts(g, pa, dst)[FIGHTERS]++;
      g->tot_sea[pa][dst]++;
      if (moved == 1) return;
      moved++;
    }
  }
}
template <bool EXPAND>
CAAA_HD_NOINLINE int move_destroyers_battleships(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_MOVE_SHIPS)) != 0);
  if (!(g->work & W(CAAA_PH_MOVE_SHIPS))) return PH_DONE;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int cell = 0;  // (ut - DESTROYERS) * NS + sea
  if (g->cur_phase == CAAA_PH_MOVE_SHIPS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_SHIPS, SC_SHIPS_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    const int ut = cell_type(cc), s = cell_loc(cc), unmoved = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= 5 * NS) {
      active = false;
      continue;
    }
    const int done = done_moving_sea(ut);
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      processed = true;
      g->vm[0] = (uint8_t)sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, 2, 0);
    }
    uint32_t da = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      da = ai_input<EXPAND>(g, cx);
    }
    if (g->enemy_count[da] > 0) flag_combat(g, da, 2);
    if (sa == da) {
      arr[done] = (uint8_t)(arr[done] + arr[unmoved]);
      arr[unmoved] = 0;
      continue;
    }
    const int ds = (int)da - NL;
    su_ptr(g, ds, ut)[done]++;
    uts(g, cur, ds)[ut]++;
    g->tot_sea[cur][ds]++;
    arr[unmoved]--;
    uts(g, cur, s)[ut]--;
    g->tot_sea[cur][s]--;
    if (ut == CARRIERS) carry_allied_fighters(g, cx, s, ds);
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_MOVE_SHIPS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_MOVE_SHIPS);
  return PH_DONE;
}

// Walks the three consecutive sea-mover phases 6..8 (loaded transports, submarines, surface ships; reference
// src/engine.cpp:4220-4222) in ONE flat loop, like move_land_units does for the land movers: per game the cells, lists
// and draws come in exactly the order of the separate phases (move_transport_units / move_subs /
// move_destroyers_battleships above), but lanes that are in different phases share every decision step and a game
// needs one visit instead of up to three.
template <bool EXPAND>
CAAA_HD_NOINLINE int move_sea_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  constexpr uint32_t range = W(CAAA_PH_MOVE_TRANSPORTS) | W(CAAA_PH_MOVE_SUBS) | W(CAAA_PH_MOVE_SHIPS);
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & range) != 0);
  if (!(g->work & range)) return PH_DONE;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int ph = CAAA_PH_MOVE_TRANSPORTS, cell = 0;
  if (g->cur_phase >= CAAA_PH_MOVE_TRANSPORTS && g->cur_phase <= CAAA_PH_MOVE_SHIPS) {
    ph = g->cur_phase;
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  } else if (g->work & W(CAAA_PH_MOVE_TRANSPORTS)) {
#pragma unroll 1
    for (int s = 0; s < NS; s++) {  // skip_empty_transports 2062-2075
      uint8_t* e = su_ptr(g, s, TRANS_EMPTY);
      const uint32_t n = e[1] + e[2] + e[3];
      if (n) {
        e[0] = (uint8_t)(e[0] + n);
        e[1] = e[2] = e[3] = 0;
      }
    }
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const bool transports = ph == CAAA_PH_MOVE_TRANSPORTS, subs = ph == CAAA_PH_MOVE_SUBS;
    const int sc_first = transports ? (int)SC_TRANSPORTS : subs ? (int)SC_SUBS : (int)SC_SHIPS;
    const int n_cells = transports ? (int)SC_TRANSPORTS_N : subs ? (int)SC_SUBS_N : (int)SC_SHIPS_N;
    const bool has_phase = (g->work & W(ph)) != 0;
    uint32_t cc = 0;
    if (has_phase) cc = scan_cells(g, CAAA_TAB(cx)->scan + sc_first, n_cells, cell, fresh);
    else cell = n_cells;
    uint8_t* const arr = cell_arr(g, cc);
    const int ut = cell_type(cc), s = cell_loc(cc), st = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= n_cells) {  // this phase is done for this game
      if (has_phase) {
        if (processed) clear_move_history(g);
        g->work &= ~W(ph);
      }
      processed = false;
      fresh = true;
      cell = 0;
      if (ph == CAAA_PH_MOVE_SHIPS) active = false;
      else ph++;
      continue;
    }
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      processed = true;
      g->vm[0] = (uint8_t)sa;
      g->nvm = 1;
      add_valid_sea_moves(g, cx, s, transports ? st - 1 : 2, subs ? 16u : 0u);
    }
    uint32_t da = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      da = ai_input<EXPAND>(g, cx);
    }
    const int ds = (int)da - NL;
    if (transports) {
      if (g->blockade[ds] > 0) flag_combat(g, da, 2);
      g->work |= W(CAAA_PH_UNLOAD_TRANSPORTS);
    } else if (subs) {
      if (g->enemy_count[ds] > 0) flag_combat(g, ds, 2);  // 2188,2194: sea index used as an air index
    } else {
      if (g->enemy_count[da] > 0) flag_combat(g, da, 2);
    }
    const int to = transports ? 1 : subs ? 0 : (int)done_moving_sea(ut);  // the state a moved (or staying) unit ends in
    if (sa == da) {
      arr[to] = (uint8_t)(arr[to] + arr[st]);
      arr[st] = 0;
      continue;
    }
    su_ptr(g, ds, ut)[to]++;
    uts(g, cur, ds)[ut]++;
    g->tot_sea[cur][ds]++;
    arr[st]--;
    uts(g, cur, s)[ut]--;
    g->tot_sea[cur][s]--;
    if (transports && ut <= TRANS_1T) {
      g->small_cargo[ds]++;
      g->small_cargo[s]--;
      if (ut <= TRANS_1I) {
        g->large_cargo[ds]++;
        g->large_cargo[s]--;
      }
    }
    if (ut == CARRIERS) carry_allied_fighters(g, cx, s, ds);
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, ph, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  return PH_DONE;
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
// The casualty loops have no early exits: a lane whose hits are used up leaves through the loop condition.  A `return`
// from inside the loops would make the function's end the only point where the lanes of a warp meet again, and the
// lanes would walk the casualty order one after the other.
CAAA_HD_NOINLINE void remove_land_defenders(Game* g, const RunCtx cx, int land, uint32_t hits) {  // 2613-2641
  const int ne = CAAA_TAB(cx)->n_enemies[g->cur];
  const uint8_t* en = CAAA_TAB(cx)->enemies_abs[g->cur];
  uint32_t removed = 0;
#pragma unroll 1
  for (int k = 0; k < 6 && hits != 0; k++) {
    const int ut = order_land_def(k);
#pragma unroll 1
    for (int e = 0; e < ne; e++) {
      const int pa = en[e];
      uint8_t* n = &utl(g, pa, land)[ut];
      const uint32_t have = *n;
      if (have != 0 && hits != 0) {
        const uint32_t taken = have < hits ? have : hits;
        *n = (uint8_t)(have - taken);
        hits -= taken;
        g->tot_land[pa][land] = (int16_t)(g->tot_land[pa][land] - (int)taken);
        removed += taken;
      }
    }
  }
  g->enemy_count[land] = (int16_t)(g->enemy_count[land] - (int)removed);
}
CAAA_HD_NOINLINE void remove_land_attackers(Game* g, int land, uint32_t hits) {  // 2642-2698
  const int cur = g->cur;
  uint8_t* mine = utl(g, cur, land);
#pragma unroll 1
  for (int ut = INFANTRY; ut <= TANKS && hits != 0; ut++) {  // ORDER_OF_LAND_ATTACKERS_1, state 0 only
    uint8_t* n = &lu_ptr(g, land, ut)[0];
    const uint32_t have = *n;
    if (have != 0) {
      const bool done = have >= hits;  // all remaining hits absorbed by this type
      const uint32_t taken = done ? hits : have;
      *n = (uint8_t)(have - taken);
      hits -= taken;
      g->tot_land[cur][land] = (int16_t)(g->tot_land[cur][land] - (int)taken);
      mine[ut] = done ? (uint8_t)(mine[ut] - taken) : (uint8_t)0;
    }
  }
#pragma unroll 1
  for (int ut = FIGHTERS; ut <= BOMBERS_LAND_AIR && hits != 0; ut++) {  // ORDER_OF_LAND_ATTACKERS_2, states 1..n-2
    if (mine[ut] != 0) {
      uint8_t* n = lu_ptr(g, land, ut);
#pragma unroll 1
      for (int s = 1; s < (int)states_land(ut) - 1 && hits != 0; s++) {
        const uint32_t have = n[s];
        if (have != 0) {
          const uint32_t taken = have < hits ? have : hits;
          n[s] = (uint8_t)(have - taken);
          hits -= taken;
          mine[ut] = (uint8_t)(mine[ut] - taken);
          g->tot_land[cur][land] = (int16_t)(g->tot_land[cur][land] - (int)taken);
        }
      }
    }
  }
}
CAAA_HD_NOINLINE void remove_sea_defenders(Game* g, const RunCtx cx, int sea, uint32_t hits, bool submerged) {  // 2700-2806
  const int ne = CAAA_TAB(cx)->n_enemies[g->cur], air = sea + NL;
  const uint8_t* en = CAAA_TAB(cx)->enemies_abs[g->cur];
#pragma unroll 1
  for (int e = 0; e < ne; e++) {  // battleships absorb one hit each
    uint8_t* su = uts(g, en[e], sea);
    const uint32_t have = su[BATTLESHIPS];
    if (have != 0 && hits != 0) {
      const uint32_t taken = have < hits ? have : hits;
      su[BATTLESHIPS] = (uint8_t)(have - taken);
      su[BS_DAMAGED] = (uint8_t)(su[BS_DAMAGED] + taken);
      hits -= taken;
    }
  }
  uint32_t removed = 0, unblocked = 0;
  uint64_t order = CAAA_PK(SUBMARINES, DESTROYERS, CARRIERS, CRUISERS, FIGHTERS, BS_DAMAGED, TRANS_EMPTY, TRANS_1I) >> (submerged ? 8 : 0);
#pragma unroll 1
  for (int k = submerged ? 1 : 0; k < 13 && hits != 0; k++) {
    if (k == 8) order = CAAA_PK(TRANS_1A, TRANS_1T, TRANS_2I, TRANS_1I_1A, TRANS_1I_1T, 0, 0, 0);
    const int ut = (int)(order & 0xffu);
    order >>= 8;
    const bool in_blockade_range = k >= DESTROYERS && k <= BS_DAMAGED;  // 2783,2791: order index vs type ids
#pragma unroll 1
    for (int e = 0; e < ne; e++) {
      const int pa = en[e];
      uint8_t* n = &uts(g, pa, sea)[ut];
      const uint32_t have = *n;
      if (have != 0 && hits != 0) {
        const uint32_t taken = have < hits ? have : hits;
        *n = (uint8_t)(have - taken);
        hits -= taken;
        g->tot_sea[pa][sea] = (int16_t)(g->tot_sea[pa][sea] - (int)taken);
        removed += taken;
        if (in_blockade_range) unblocked += taken;
      }
    }
  }
  g->enemy_count[air] = (int16_t)(g->enemy_count[air] - (int)removed);
  g->blockade[sea] = (uint8_t)(g->blockade[sea] - unblocked);
}
// attacker casualties among state-0 units of one type
CAAA_HD void hit_sea_state0(Game* g, int sea, int ut, uint32_t& hits) {
  uint8_t* n = &su_ptr(g, sea, ut)[0];
  const uint32_t have = *n;
  if (have == 0 || hits == 0) return;
  const int cur = g->cur;
  const bool done = have >= hits;  // all remaining hits absorbed by this type
  const uint32_t taken = done ? hits : have;
  *n = (uint8_t)(have - taken);
  hits -= taken;
  g->tot_sea[cur][sea] = (int16_t)(g->tot_sea[cur][sea] - (int)taken);
  if (ut >= TRANS_EMPTY && ut <= TRANS_1T) {
    g->small_cargo[sea] = (int16_t)(g->small_cargo[sea] - (int)taken);
    if (ut <= TRANS_1I) g->large_cargo[sea] = (int16_t)(g->large_cargo[sea] - (int)taken);
  }
  uint8_t* c = uts(g, cur, sea) + ut;
  if (done) {
    *c = (uint8_t)(*c - taken);
    if (ut == CARRIERS) g->carriers[sea] = (int16_t)(g->carriers[sea] - (int)taken);
  } else {
    *c = 0;
    if (ut == CARRIERS) g->carriers[sea] = 0;
  }
}
CAAA_HD_NOINLINE void remove_sea_attackers(Game* g, int sea, uint32_t hits) {  // 2808-2982
  const int cur = g->cur;
  uint8_t* mine = uts(g, cur, sea);
  {
    uint8_t* bb = &su_ptr(g, sea, BATTLESHIPS)[0];
    const uint32_t have = *bb;
    if (have > 0) {
      const bool done = have >= hits;
      const uint32_t taken = done ? hits : have;
      *bb = (uint8_t)(have - taken);
      hits -= taken;
      uint8_t* bd = &su_ptr(g, sea, BS_DAMAGED)[0];
      *bd = (uint8_t)(*bd + taken);
      mine[BS_DAMAGED] = (uint8_t)(mine[BS_DAMAGED] + taken);
      mine[BATTLESHIPS] = done ? (uint8_t)(mine[BATTLESHIPS] - taken) : (uint8_t)0;
    }
  }
#pragma unroll 1
  for (int ut = SUBMARINES; ut <= CRUISERS && hits != 0; ut++) hit_sea_state0(g, sea, ut, hits);
#pragma unroll 1
  for (int k = 0; k < 2 && hits != 0; k++) {  // air units in any state, 2915-2942
    const int ut = k == 0 ? (int)FIGHTERS : (int)BOMBERS_SEA;
    uint8_t* c = cur_sea(g, sea, ut);
    if (*c != 0) {
      uint8_t* n = su_ptr(g, sea, ut);
#pragma unroll 1
      for (int s = 0; s < 5 && hits != 0; s++) {
        const uint32_t have = n[s];
        if (have != 0) {
          const uint32_t taken = have < hits ? have : hits;
          n[s] = (uint8_t)(have - taken);
          hits -= taken;
          g->tot_sea[cur][sea] = (int16_t)(g->tot_sea[cur][sea] - (int)taken);
          *c = (uint8_t)(*c - taken);
        }
      }
    }
  }
#pragma unroll 1
  for (int k = 0; k < 8 && hits != 0; k++) hit_sea_state0(g, sea, order_sea_att3(k), hits);
}

// ---- phase 9: resolve_sea_battles, 2312-2550, 2582-2603 ----
CAAA_HD void sea_retreat(Game* g, int src, int dst) {  // 2582-2603
  const int cur = g->cur;
#pragma unroll 1
  for (int ut = TRANS_EMPTY; ut <= BS_DAMAGED; ut++) {
    const uint32_t n = uts(g, cur, src)[ut];
    uint8_t* d = su_ptr(g, dst, ut);
    d[0] = (uint8_t)(d[0] + n);
    uts(g, cur, dst)[ut] = (uint8_t)(uts(g, cur, dst)[ut] + n);
    g->tot_sea[cur][dst] = (int16_t)(g->tot_sea[cur][dst] + (int)n);
    g->tot_sea[cur][src] = (int16_t)(g->tot_sea[cur][src] - (int)n);
    su_ptr(g, src, ut)[0] = 0;
    su_ptr(g, src, ut)[1] = 0;
    uts(g, cur, src)[ut] = 0;
  }
  g->flagged[src + NL] = 0;
}
template <bool EXPAND>
CAAA_HD_NOINLINE int resolve_sea_battles(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W_SEA_BATTLE) != 0);
  if (!(g->work & W_SEA_BATTLE)) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur, ne = CAAA_TAB(cx)->n_enemies[cur];
  const uint8_t* en = CAAA_TAB(cx)->enemies_abs[cur];
  int s = 0, rounds = 0;
  bool in_battle = false, submerged = false;
  if (g->cur_phase == CAAA_PH_SEA_BATTLES) {
    s = g->cur_cell;
    in_battle = (g->cur_flags & CF_IN_BATTLE) != 0;
    submerged = (g->cur_flags & CF_SUBMERGED) != 0;
    rounds = (int)g->cur_b;
  }
  bool active = true, stop = false;
  int steps = 0;
  // One combat round per iteration.  The body has no `continue`: a jump to the loop head from inside a conditional
  // block would make the loop head the first point where the lanes of a warp meet again, and everything after the
  // block would run one lane at a time.  `go` carries "this lane is still in this round" instead, and converge()
  // re-joins the lanes after every block whose length differs from lane to lane.
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;  // (waits at the vote above; the lanes below are exactly `mact`)
    bool go = true;
    if (!in_battle) {  // next flagged sea with own units (one sea per iteration keeps the lanes together)
#pragma unroll 1
      for (; s < NS; s++)
        if (g->flagged[s + NL] != 0 && g->tot_sea[cur][s] != 0) break;
      if (s >= NS) {
        active = false;
        go = false;
      } else {
        const uint8_t* mine0 = uts(g, cur, s);
        submerged = mine0[DESTROYERS] == 0;
        bool skip = false;
        if (g->tot_sea[cur][s] == 2) {  // 2336: literal "== 2"
          if (submerged) {
            uint32_t subs = 0;
#pragma unroll 1
            for (int e = 0; e < ne; e++) subs += uts(g, en[e], s)[SUBMARINES];
            if ((int)(subs & 0xffu) == g->enemy_count[s]) skip = true;  // 2344: land index
          }
          if (!skip)
#pragma unroll 1
            for (int ut = CRUISERS; ut <= BS_DAMAGED; ut++) {  // bombardment no longer possible
              uint8_t* a = su_ptr(g, s, ut);
              a[0] = (uint8_t)(a[0] + a[1]);
              a[1] = 0;
            }
        }
        if (skip) {
          s++;
          go = false;
        } else {
          in_battle = true;
          rounds = 0;
        }
      }
    }
    converge<!EXPAND>(mact);
    // ---- one combat round ----
    const int air = s + NL;
    uint8_t* mine = uts(g, cur, s < NS ? s : 0);
    bool over = false;
    if (go) {
      if (rounds >= SEA_ROUND_CAP) {  // guard (not in the reference, which would spin for ever)
        g->status |= CAAA_ERR_SEA_ROUND_CAP;
        stop = true;
        active = false;
        go = false;
      } else {
        rounds++;
      }
    }
    if (go && g->flagged[air] == 1) {  // retreat option, 2368-2406
      uint32_t n = 0;
      if (g->cur_sea_bombers[s] + mine[FIGHTERS] + mine[SUBMARINES] + mine[DESTROYERS] + mine[CARRIERS] + mine[CRUISERS] +
                  mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0 ||
          g->blockade[s] == 0)
        g->vm[n++] = (uint8_t)air;
#pragma unroll 1
      for (int i = 0; i < t.s2s_count[s]; i++) {
        const int d = t.s2s_conn[s][i];
        if (g->blockade[d] == 0) g->vm[n++] = (uint8_t)(d + NL);
      }
      g->nvm = (uint8_t)n;
      if (n > 0) {
        uint32_t da = g->vm[0];
        if (n > 1) {
          if (g->answers_remaining == 0) {
            stop = true;
            active = false;
            go = false;
          } else {
            da = ai_input<EXPAND>(g, cx);
          }
        }
        if (go) {
          const int ds = (uint8_t)(da - NL);
          if (ds < NS && t.sea_dist[g->canal_state][s][ds] == 1 && g->flagged[da] == 0) {
            sea_retreat(g, s, ds);
            over = true;
          }
        }
      }
    }
    converge<!EXPAND>(mact);
    if (go && !over) {
      g->flagged[air] = 1;
      bool targets = false;
      if (mine[DESTROYERS]) {
        targets = g->enemy_count[air] > 0;
      } else if (mine[CARRIERS] + mine[CRUISERS] + mine[BATTLESHIPS] + mine[BS_DAMAGED] > 0) {
        if (g->enemy_count[air] > 0)
#pragma unroll 1
          for (int e = 0; e < ne; e++)
            if (g->tot_sea[en[e]][s] - uts(g, en[e], s)[SUBMARINES] > 0) { targets = true; break; }
      } else if (mine[SUBMARINES] > 0) {
#pragma unroll 1
        for (int e = 0; e < ne; e++) {
          const uint8_t* eu = uts(g, en[e], s);
          if (g->tot_sea[en[e]][s] - (eu[FIGHTERS] + eu[SUBMARINES]) > 0) { targets = true; break; }
        }
      } else if (mine[FIGHTERS] + g->cur_sea_bombers[s] > 0) {
#pragma unroll 1
        for (int e = 0; e < ne; e++)
          if (g->tot_sea[en[e]][s] - uts(g, en[e], s)[SUBMARINES] > 0) { targets = true; break; }
      }
      if (!targets) {  // 2460-2482
        if (g->enemy_count[air] > 0) {
          g->tot_sea[cur][s] = (int16_t)(g->tot_sea[cur][s] - mine[TRANS_EMPTY]);
          mine[TRANS_EMPTY] = 0;
          su_ptr(g, s, TRANS_EMPTY)[0] = 0;
#pragma unroll 1
          for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
            g->tot_sea[cur][s] = (int16_t)(g->tot_sea[cur][s] - mine[ut]);
            mine[ut] = 0;
            su_ptr(g, s, ut)[1] = 0;
          }
        }
        g->flagged[air] = 0;
        over = true;
      }
    }
    const bool fight = go && !over;
    const unsigned mfight = vote_ballot<!EXPAND>(mact, fight);  // (also re-joins the lanes)
    if (fight) {
      // sub volley: only state-0 attacking subs count (2484)
      int admg = su_ptr(g, s, SUBMARINES)[0] * 2;
      uint32_t ahits = attacker_hits(g, cx, admg);
      uint32_t dhits = 0;
      if (!submerged) {
        int ddmg = 0;
#pragma unroll 1
        for (int e = 0; e < ne; e++) ddmg += uts(g, en[e], s)[SUBMARINES];
        dhits = defender_hits(g, cx, ddmg);
      }
      converge<!EXPAND>(mfight);
      if (dhits > 0) remove_sea_attackers(g, s, dhits);
      converge<!EXPAND>(mfight);
      if (ahits > 0) remove_sea_defenders(g, cx, s, ahits, submerged);
      converge<!EXPAND>(mfight);
      // main volley, 2513-2542
      admg = mine[DESTROYERS] * 2 + mine[CARRIERS] * 1 + mine[CRUISERS] * 3 + mine[BATTLESHIPS] * 4 + mine[BS_DAMAGED] * 4 +
             mine[FIGHTERS] * 3 + g->cur_sea_bombers[s] * 4;
      ahits = attacker_hits(g, cx, admg);
      int ddmg = 0;
#pragma unroll 1
      for (int e = 0; e < ne; e++)  // 2527-2530: type ids 0..4 (only fighters have a defence value) + fighters again
        ddmg += uts(g, en[e], s)[FIGHTERS] * 4 * 2;
      dhits = defender_hits(g, cx, ddmg);
      converge<!EXPAND>(mfight);
      if (dhits > 0) remove_sea_attackers(g, s, dhits);
      converge<!EXPAND>(mfight);
      if (ahits > 0) remove_sea_defenders(g, cx, s, ahits, submerged);
      converge<!EXPAND>(mfight);
      if (g->enemy_count[air] == 0 || g->tot_sea[cur][s] == 0) {
        g->flagged[air] = 0;
        over = true;
      }
    }
    if (go && over) {
      in_battle = false;
      s++;
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_SEA_BATTLES, s, (in_battle ? CF_IN_BATTLE : 0u) | (submerged ? CF_SUBMERGED : 0u), 0u, (uint32_t)rounds);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  g->work &= ~W_SEA_BATTLE;
  return PH_DONE;
}

// ---- phase 10: unload_transports, 2984-3053, 3373-3383 ----
template <bool EXPAND>
CAAA_HD_NOINLINE int unload_transports(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_UNLOAD_TRANSPORTS)) != 0);
  if (!(g->work & W(CAAA_PH_UNLOAD_TRANSPORTS))) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  int cell = 0;  // (ut - TRANS_1I) * NS + sea
  if (g->cur_phase == CAAA_PH_UNLOAD_TRANSPORTS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_UNLOAD, SC_UNLOAD_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);  // STATES_UNLOADING is 1 for every loaded transport type
    const int ut = cell_type(cc), s = cell_loc(cc);
    converge<!EXPAND>(mact);
    if (cell >= 6 * NS) {
      active = false;
      continue;
    }
    const uint32_t sa = (uint32_t)(s + NL);
    if (fresh) {
      fresh = false;
      processed = true;
      g->vm[0] = (uint8_t)sa;
      uint32_t n = 1;
#pragma unroll 1
      for (int i = 0; i < t.s2l_count[s]; i++) {
        const int d = t.s2l_conn[s][i];
        if (!skip_get(g, sa, d)) g->vm[n++] = (uint8_t)d;
      }
      g->nvm = (uint8_t)n;
    }
    uint32_t dst = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    if (sa == dst) {
      arr[0] = (uint8_t)(arr[0] + arr[1]);
      arr[1] = 0;
      continue;
    }
    const int c1 = unload_cargo1(ut);
    uint8_t* mine = utl(g, cur, dst);
    g->bombard_max[dst]++;
    lu_ptr(g, dst, c1)[0]++;
    mine[c1]++;
    g->tot_land[cur][dst]++;
    su_ptr(g, s, TRANS_EMPTY)[0]++;
    uts(g, cur, s)[TRANS_EMPTY]++;
    uts(g, cur, s)[ut]--;
    arr[1]--;
    if (ut > TRANS_1T) {  // second cargo is always infantry
      g->bombard_max[dst]++;
      lu_ptr(g, dst, INFANTRY)[0]++;
      mine[INFANTRY]++;
      g->tot_land[cur][dst]++;
    }
    if (!owner_allied(g, cx, dst)) {
      flag_combat(g, dst, 2);
      if (g->enemy_count[dst] == 0) conquer_land(g, cx, dst);
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_UNLOAD_TRANSPORTS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, 0u);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_UNLOAD_TRANSPORTS);
  return PH_DONE;
}

// ---- phase 11: resolve_land_battles, 3055-3371 ----
CAAA_HD void hit_air_states(Game* g, int land, int ut, int lo, int hi, uint32_t& hits, bool stop_on_absorb) {
  // shared shape of the strategic-bombing / AA-fire casualty loops (3103-3117, 3179-3210)
  const int cur = g->cur;
#pragma unroll 1
  for (int s = lo; s < hi; s++) {
    uint32_t taken;
    const bool done = take_hits(&lu_ptr(g, land, ut)[s], hits, taken);
    utl(g, cur, land)[ut] = (uint8_t)(utl(g, cur, land)[ut] - taken);
    g->tot_land[cur][land] = (int16_t)(g->tot_land[cur][land] - (int)taken);
    if (done) {
      hits = 0;
      if (stop_on_absorb) return;
    }
  }
}
template <bool EXPAND>
CAAA_HD_NOINLINE int resolve_land_battles(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W_LAND_BATTLE) != 0);
  if (!(g->work & W_LAND_BATTLE)) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur, ne = CAAA_TAB(cx)->n_enemies[cur];
  const uint8_t* en = CAAA_TAB(cx)->enemies_abs[cur];
  int l = 0;
  uint32_t rounds = 0;
  bool in_battle = false;
  if (g->cur_phase == CAAA_PH_LAND_BATTLES) {
    l = g->cur_cell;
    in_battle = (g->cur_flags & CF_IN_BATTLE) != 0;
    rounds = g->cur_b;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;  // (waits at the vote above; the lanes below are exactly `mact`)
    // no `continue` below, converge() after the variable-length blocks: see resolve_sea_battles
    bool go = true;
    if (!in_battle) {
#pragma unroll 1
      for (; l < NL; l++)
        if (g->flagged[l] != 0 && g->tot_land[cur][l] != 0) break;
      if (l >= NL) {
        active = false;
        go = false;
      } else {
        uint8_t* mine = utl(g, cur, l);
        int16_t* my_total = &g->tot_land[cur][l];
        bool skip = false;
        if (g->flagged[l] == 2) {
          if (mine[BOMBERS_LAND_AIR] > 0 && *my_total == mine[BOMBERS_LAND_AIR]) {  // strategic bombing 3088-3125
            uint32_t ahits = 0;  // uninitialised in the reference when the factory is already at -max; same result
            if (g->factory_hp[l] > -(int)g->factory_max[l]) {
              uint32_t dh = defender_hits(g, cx, mine[BOMBERS_LAND_AIR]);
              if (dh > 0) hit_air_states(g, l, BOMBERS_LAND_AIR, 1, 6, dh, false);
              ahits = attacker_hits(g, cx, mine[BOMBERS_LAND_AIR] * 21);
            }
            const int hp = g->factory_hp[l] - (int)ahits, floor_hp = -(int)g->factory_max[l];
            g->factory_hp[l] = (int8_t)(hp > floor_hp ? hp : floor_hp);
            skip = true;
          } else {
            if (g->bombard_max[l] > 0) {  // shore bombardment 3134-3158
              int admg = 0;
#pragma unroll 1
              for (int ut = BS_DAMAGED; ut >= CRUISERS; ut--)
#pragma unroll 1
                for (int i = 0; i < t.l2s_count[l]; i++) {
                  uint8_t* ships = su_ptr(g, t.l2s_conn[l][i], ut);
#pragma unroll 1
                  while (ships[1] > 0 && g->bombard_max[l] > 0) {
                    admg += attack_sea(ut);
                    ships[0]++;
                    ships[1]--;
                    g->bombard_max[l]--;
                  }
                }
              g->bombard_max[l] = 0;
              const uint32_t ahits = attacker_hits(g, cx, admg);
              if (ahits > 0) remove_land_defenders(g, cx, l, ahits);
            }
            const uint32_t air_units = (uint8_t)(mine[FIGHTERS] + mine[BOMBERS_LAND_AIR]);  // AA fire 3160-3216
            if (air_units > 0) {
              int aa = 0;
#pragma unroll 1
              for (int e = 0; e < ne; e++) aa += utl(g, en[e], l)[AA_GUNS];
              if (aa > 0) {
                uint32_t dh = defender_hits(g, cx, (uint8_t)(air_units * 3));
                if (dh > 0) hit_air_states(g, l, FIGHTERS, 0, 5, dh, true);
                if (dh > 0) hit_air_states(g, l, BOMBERS_LAND_AIR, 0, 7, dh, true);
              }
            }
            if (*my_total == 0) {
              skip = true;
            } else if (g->enemy_count[l] == 0) {
              if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(g, cx, l);
              skip = true;
            }
          }
        }
        if (skip) {
          l++;
          go = false;
        } else {
          in_battle = true;
          rounds = 0;
        }
      }
    }
    converge<!EXPAND>(mact);
    // ---- one combat round ----
    const int lc = l < NL ? l : 0;  // (a lane that is not `go` may stand past the last land)
    uint8_t* mine = utl(g, cur, lc);
    int16_t* my_total = &g->tot_land[cur][lc];
    bool over = false;
    if (go) {
      rounds = (rounds + 1) & 0xffu;  // uint8 counter in the reference
      if (*my_total == 0) over = true;
    }
    if (go && !over && g->flagged[l] == 1) {  // retreat option 3256-3299
      g->vm[0] = (uint8_t)l;
      uint32_t n = 1;
#pragma unroll 1
      for (int i = 0; i < t.l2l_count[l]; i++) {
        const int d = t.l2l_conn[l][i];
        if (g->enemy_count[d] == 0 && g->flagged[d] == 0 && owner_allied(g, cx, d)) g->vm[n++] = (uint8_t)d;
      }
      g->nvm = (uint8_t)n;
      uint32_t dst = g->vm[0];
      if (n > 1) {
        if (g->answers_remaining == 0) {
          stop = true;
          active = false;
          go = false;
        } else {
          dst = ai_input<EXPAND>(g, cx);
        }
      }
      if (go && ((uint32_t)l != dst || rounds > 100)) {
        const int dl = dst < (uint32_t)NL ? (int)dst : l;  // dst is always a land here
#pragma unroll 1
        for (int ut = INFANTRY; ut <= TANKS; ut++) {
          const uint32_t k = lu_ptr(g, l, ut)[0];
          lu_ptr(g, dl, ut)[0] = (uint8_t)(lu_ptr(g, dl, ut)[0] + k);
          utl(g, cur, dl)[ut] = (uint8_t)(utl(g, cur, dl)[ut] + k);
          g->tot_land[cur][dl] = (int16_t)(g->tot_land[cur][dl] + (int)k);
          mine[ut] = (uint8_t)(mine[ut] - k);
          *my_total = (int16_t)(*my_total - (int)k);
          lu_ptr(g, l, ut)[0] = 0;
        }
        g->flagged[l] = 0;
        over = true;
      }
    }
    const bool fight = go && !over;
    const unsigned mfight = vote_ballot<!EXPAND>(mact, fight);  // (also re-joins the lanes)
    if (fight) {
      g->flagged[l] = 1;
      const int inf = mine[INFANTRY], art = mine[ARTILLERY];
      const int admg = mine[FIGHTERS] * 3 + mine[BOMBERS_LAND_AIR] * 4 + inf + art * 2 + mine[TANKS] * 3 + (inf < art ? inf : art);
      const uint32_t ahits = attacker_hits(g, cx, admg);
      int ddmg = 0;
      uint32_t dh = 0;
#pragma unroll 1
      for (int e = 0; e < ne; e++) {  // 3319-3327: one draw per enemy player, the last one counts
        const uint8_t* eu = utl(g, en[e], l);
        ddmg += eu[INFANTRY] * 2 + eu[ARTILLERY] * 2 + eu[TANKS] * 3 + eu[FIGHTERS] * 4 + eu[BOMBERS_LAND_AIR];
        dh = defender_hits(g, cx, ddmg);
      }
      converge<!EXPAND>(mfight);
      if (dh > 0) remove_land_attackers(g, l, dh);
      converge<!EXPAND>(mfight);
      if (ahits > 0) remove_land_defenders(g, cx, l, ahits);
      converge<!EXPAND>(mfight);
      if (*my_total == 0) {
        over = true;
      } else if (g->enemy_count[l] == 0) {
        if (mine[INFANTRY] + mine[ARTILLERY] + mine[TANKS] > 0) conquer_land(g, cx, l);
        over = true;
      }
    }
    if (go && over) {
      in_battle = false;
      l++;
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_LAND_BATTLES, l, in_battle ? CF_IN_BATTLE : 0u, 0u, rounds);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  g->work &= ~W_LAND_BATTLE;
  return PH_DONE;
}

// ---- phases 13/14: land_fighter_units / land_bomber_units, 3409-3421, 3445-3654 ----
CAAA_HD uint32_t fighter_final_landing_mask(Game* g, const RunCtx cx) {  // refresh_can_fighter_land_here 3445-3466
  const CaaaTables& t = CAAA_TAB(cx)->t;
  uint32_t here = 0;
#pragma unroll 1
  for (int s = 0; s < NS; s++)
    if (g->carriers[s] > 0) here |= 1u << (s + NL);
#pragma unroll 1
  for (int l = 0; l < NL; l++) {
    if (owner_allied(g, cx, l) && g->flagged[l] == 0) here |= 1u << l;
    if (g->factory_max[l] > 0 && owner_rel(g, l) == g->player)
#pragma unroll 1
      for (int i = 0; i < t.l2s_count[l]; i++) here |= 1u << (NL + t.l2s_conn[l][i]);
  }
  return here;
}
template <bool EXPAND>
CAAA_HD_NOINLINE int land_fighter_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_LAND_FIGHTERS)) != 0);
  if (!(g->work & W(CAAA_PH_LAND_FIGHTERS))) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  uint32_t here = 0;
  int cell = 0;  // (state - 1) * NA + src
  if (g->cur_phase == CAAA_PH_LAND_FIGHTERS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
    here = g->cur_b;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_LAND_FIGHTERS, SC_LAND_FIGHTERS_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    const int src = cell_loc(cc), cs = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= 3 * NA) {
      active = false;
      continue;
    }
    if (fresh) {
      fresh = false;
      if (!processed) {
        processed = true;
        here = fighter_final_landing_mask(g, cx);
      }
      uint32_t n = 0;
#pragma unroll 1
      for (int i = 0; i < t.air_within_count[cs - 1][src]; i++) {  // add_valid_fighter_landing
        const int d = t.air_within[cs - 1][src][i];
        if (((here >> d) & 1) && !skip_get(g, src, d)) g->vm[n++] = (uint8_t)d;
      }
      if (n == 0 || ((here >> src) & 1)) g->vm[n++] = (uint8_t)src;
      g->nvm = (uint8_t)n;
    }
    uint32_t dst = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    if ((uint32_t)src == dst) {
      arr[0]++;
      arr[cs]--;
      continue;
    }
    air_fighters(g, dst)[0]++;
    if (dst < (uint32_t)NL) {
      g->tot_land[cur][dst]++;
      utl(g, cur, dst)[FIGHTERS]++;
    } else {
      g->tot_sea[cur][dst - NL]++;
      uts(g, cur, dst - NL)[FIGHTERS]++;
    }
    if (src < NL) {
      g->tot_land[cur][src]--;
      utl(g, cur, src)[FIGHTERS]--;
    } else {
      g->tot_sea[cur][src - NL]--;
      uts(g, cur, src - NL)[FIGHTERS]--;
    }
    arr[cs]--;
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_LAND_FIGHTERS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, here);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_LAND_FIGHTERS);
  return PH_DONE;
}
template <bool EXPAND>
CAAA_HD_NOINLINE int land_bomber_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = vote_ballot<!EXPAND>(min, (g->work & W(CAAA_PH_LAND_BOMBERS)) != 0);
  if (!(g->work & W(CAAA_PH_LAND_BOMBERS))) return PH_DONE;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  bool processed = false, fresh = true;
  uint32_t here = 0;
  int cell = 0;  // cur1 * NA + src
  if (g->cur_phase == CAAA_PH_LAND_BOMBERS) {
    cell = g->cur_cell;
    fresh = g->cur_flags & CF_FRESH;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
    here = g->cur_b;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    const uint32_t cc = scan_cells(g, CAAA_TAB(cx)->scan + SC_LAND_BOMBERS, SC_LAND_BOMBERS_N, cell, fresh);
    uint8_t* const arr = cell_arr(g, cc);
    const int src = cell_loc(cc), cs = cell_state(cc);
    converge<!EXPAND>(mact);
    if (cell >= 5 * NA) {
      active = false;
      continue;
    }
    if (fresh) {
      fresh = false;
      if (!processed) {
        processed = true;
        here = bomber_here_mask(g, cx);
      }
      const int moves = cs + (src < NL ? 0 : 1);
      uint32_t n = 0;
#pragma unroll 1
      for (int i = 0; i < t.air_to_land_within_count[moves - 1][src]; i++) {  // add_valid_bomber_landing
        const int d = t.air_to_land_within[moves - 1][src][i];
        if ((here >> d) & 1) g->vm[n++] = (uint8_t)d;
      }
      g->nvm = (uint8_t)n;
    }
    if (g->nvm == 0) g->vm[g->nvm++] = (uint8_t)src;
    uint32_t dst = g->vm[0];
    if (g->nvm > 1) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      dst = ai_input<EXPAND>(g, cx);
    }
    if ((uint32_t)src == dst) {
      arr[0]++;
      arr[cs]--;
      continue;
    }
    air_bombers(g, dst)[0]++;
    const int dl = dst < (uint32_t)NL ? (int)dst : 0;  // bombers only land on land
    g->tot_land[cur][dl]++;
    utl(g, cur, dl)[BOMBERS_LAND_AIR]++;
    if (src < NL) {
      g->tot_land[cur][src]--;
      utl(g, cur, src)[BOMBERS_LAND_AIR]--;
    } else {
      g->tot_sea[cur][src - NL]--;
      uts(g, cur, src - NL)[BOMBERS_LAND_AIR]--;  // 3638: type index 1 = empty transports
    }
    arr[cs]--;
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_LAND_BOMBERS, cell, (fresh ? CF_FRESH : 0u) | (processed ? CF_PROCESSED : 0u), 0u, here);
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  if (processed) clear_move_history(g);
  g->work &= ~W(CAAA_PH_LAND_BOMBERS);
  return PH_DONE;
}

// ---- phase 15: buy_units (purchase and placement), 3663-3834 ----
// One flat loop: a cell is (factory, adjacent sea) or (factory, the land itself); every iteration is one
// purchase decision of the cell the lane is at.
template <bool EXPAND>
CAAA_HD_NOINLINE int buy_units(Game* g, const RunCtx cx, unsigned min, int yield = 0) {
  const unsigned m = min;
  const CaaaTables& t = CAAA_TAB(cx)->t;
  const int cur = g->cur;
  int f = 0, si = 0;
  bool new_factory = true, new_cell = true, processed = false;
  uint32_t repair = 0, last = 0;
  if (g->cur_phase == CAAA_PH_BUY_UNITS) {
    f = g->cur_cell;
    si = g->cur_a;
    new_factory = (g->cur_flags & CF_NEW_FACTORY) != 0;
    new_cell = (g->cur_flags & CF_NEW_CELL) != 0;
    processed = (g->cur_flags & CF_PROCESSED) != 0;
    repair = g->cur_b & 0xffu;
    last = (g->cur_b >> 8) & 0xffu;
  }
  bool active = true, stop = false;
  int steps = 0;
  for (;;) {
    const unsigned mact = vote_ballot<!EXPAND>(m, active);  // the lanes that take part in this iteration
    if (mact == 0) break;
    if (should_yield<!EXPAND>(m, active, yield, steps)) break;
    if (!active) continue;
    if (f >= g->nfact[cur]) {
      active = false;
      continue;
    }
    const int land = g->factloc[cur][f];
    if (new_factory) {
      if (g->builds_left[land] == 0) {
        f++;
        continue;
      }
      new_factory = false;
      new_cell = true;
      repair = 0;
      si = 0;
    }
    const bool at_sea = si < t.l2s_count[land];
    const int sea = at_sea ? t.l2s_conn[land][si] : 0;
    const int loc = at_sea ? sea + NL : land;
    const uint32_t pass = at_sea ? NST : NLT;
    if (new_cell) {
      new_cell = false;
      g->vm[0] = (uint8_t)pass;
      last = 0;
      processed = false;
    }
    bool close = g->builds_left[loc] == 0;
    if (!close && g->money[cur] < (at_sea ? 8 : 3)) {
      g->builds_left[loc] = 0;
      close = true;
    }
    if (!close) {
      const uint32_t built = (uint8_t)(g->factory_max[land] - g->builds_left[land]);
      if (g->factory_hp[land] <= (int)built) repair = (uint8_t)(1 + built - g->factory_hp[land]);
      uint32_t n = 1;
      if (at_sea) {
#pragma unroll 1
        for (int k = 6; k >= 0; k--) {
          const uint32_t ut = buy_sea(k);
          if (skip_get(g, 0, ut)) {
            last = ut;
            break;
          }
          if (g->money[cur] < cost_sea(ut) + repair) continue;
          if (ut == FIGHTERS) {
            int fighters = 0;
#pragma unroll 1
            for (int pa = 0; pa < NP; pa++) fighters += uts(g, pa, sea)[FIGHTERS];
            if (g->carriers[sea] * 2 <= fighters) continue;
          }
          g->vm[n++] = (uint8_t)ut;
        }
      } else {
#pragma unroll 1
        for (int ut = NLT - 1; ut >= 0; ut--) {
          if (skip_get(g, 0, (unsigned)ut)) {
            last = (uint32_t)ut;
            break;
          }
          if (g->money[cur] < cost_land(ut) + repair) continue;
          g->vm[n++] = (uint8_t)ut;
        }
      }
      g->nvm = (uint8_t)n;
      if (n == 1) {
        g->builds_left[loc] = 0;
        close = true;
      }
    }
    if (!close) {
      if (g->answers_remaining == 0) {
        stop = true;
        active = false;
        continue;
      }
      processed = true;
      const uint32_t buy = ai_input<EXPAND>(g, cx);
      if (buy == pass) {
        g->builds_left[loc] = 0;
        close = true;
      } else {
        if (at_sea) {
#pragma unroll 1
          for (int s2 = si; s2 < t.l2s_count[land]; s2++) g->builds_left[t.l2s_conn[land][s2] + NL]--;
          g->builds_left[land]--;
          g->factory_hp[land] = (int8_t)(g->factory_hp[land] + (int)repair);
          const uint32_t b = buy < (uint32_t)NST ? buy : 0;  // only listed ids can be drawn; keeps expansion-mode inputs in range
          g->money[cur] = (uint8_t)(g->money[cur] - (cost_sea(b) + repair));
          su_ptr(g, sea, b)[0]++;
          g->tot_sea[cur][sea]++;
          (*cur_sea(g, sea, b))++;
          if (buy > last) {
#pragma unroll 1
            for (uint32_t u = last; u < buy; u++) skip_set(g, 0, u);
            last = buy;
          }
        } else {
          g->builds_left[land]--;
          g->factory_hp[land] = (int8_t)(g->factory_hp[land] + (int)repair);
          const uint32_t b = buy < (uint32_t)NLT ? buy : 0;
          g->money[cur] = (uint8_t)(g->money[cur] - (cost_land(b) + repair));
          lu_ptr(g, land, b)[0]++;
          g->tot_land[cur][land]++;
          utl(g, cur, land)[b]++;
          if (buy > last)
#pragma unroll 1
            for (uint32_t u = last; u < buy; u++) skip_set(g, 0, u);
          last = buy;
        }
      }
    }
    if (close) {
      if (processed) clear_move_history(g);
      new_cell = true;
      if (at_sea) {
        si++;
      } else {
        f++;
        new_factory = true;
      }
    }
  }
  if (stop) return PH_STOPPED;
  if (active) {
    save_cursor(g, CAAA_PH_BUY_UNITS, f, (new_factory ? CF_NEW_FACTORY : 0u) | (new_cell ? CF_NEW_CELL : 0u) | (processed ? CF_PROCESSED : 0u),
                (uint32_t)si, repair | (last << 8));
    return PH_YIELDED;
  }
  g->cur_phase = CUR_NONE;
  return PH_DONE;
}

// ---- phases 16-19: crash_air_units, reset_units_fully, collect_money, 3836-3894 ----
CAAA_HD void crash_air_units(Game* g, const RunCtx cx) {
  const int cur = g->cur;
  uint32_t lands_with_fighters = 0;
#pragma unroll 1
  for (int l = 0; l < NL; l++) lands_with_fighters |= utl(g, cur, l)[FIGHTERS];
  if (lands_with_fighters) {  // the landing mask has no side effects: only derive it when a land holds fighters
    const uint32_t here = fighter_landing_masks(g, cx);
#pragma unroll 1
    for (int l = 0; l < NL; l++) {
      if ((here >> l) & 1) continue;
      const uint32_t n = utl(g, cur, l)[FIGHTERS];
      if (n == 0) continue;
      g->tot_land[cur][l] = (int16_t)(g->tot_land[cur][l] - (int)n);
      utl(g, cur, l)[FIGHTERS] = 0;
      lu_ptr(g, l, FIGHTERS)[0] = 0;
    }
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    uint8_t* n = &su_ptr(g, s, FIGHTERS)[0];
    if (*n == 0) continue;  // space is a uint8 >= 0: nothing to lose
    uint32_t space = (uint32_t)(g->carriers[s] * 2);
#pragma unroll 1
    for (int p = 1; p < NP; p++) space -= uts(g, (cur + p) % NP, s)[FIGHTERS];
    space &= 0xffu;  // uint8 in the reference
    if (space < *n) {
      const uint32_t lost = *n - space;
      g->tot_sea[cur][s] = (int16_t)(g->tot_sea[cur][s] - (int)lost);
      uts(g, cur, s)[FIGHTERS] = (uint8_t)(uts(g, cur, s)[FIGHTERS] - lost);
      *n = (uint8_t)space;
    }
  }
}
CAAA_HD void reset_units_fully(Game* g) {
  const int cur = g->cur;
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    uint8_t* mine = uts(g, cur, s);
    const uint32_t dmg = mine[BS_DAMAGED];
    uint8_t* bb = su_ptr(g, s, BATTLESHIPS);
    bb[0] = (uint8_t)(bb[0] + dmg);
    mine[BATTLESHIPS] = (uint8_t)(mine[BATTLESHIPS] + dmg);
    su_ptr(g, s, BS_DAMAGED)[0] = 0;
    su_ptr(g, s, BS_DAMAGED)[1] = 0;
    mine[BS_DAMAGED] = 0;
  }
}
CAAA_HD void collect_money(Game* g, const RunCtx cx) {
  const int cur = g->cur;
  if (g->owner[CAAA_TAB(cx)->t.capital[cur]] == cur) g->money[cur] = (uint8_t)(g->money[cur] + g->income[cur]);
}

// ---- phase 20: rotate_turns, 3895-4029 ----
// Per-player data is stored in absolute order, so the rotation itself is `cur = (cur + 1) % 5`; what
// remains is rebuilding the incoming player's per-state counts with every unit "unmoved".
CAAA_HD_NOINLINE void rotate_turns(Game* g, const RunCtx cx) {
  const int nw = (g->cur + 1) % NP;
  g->player = (uint8_t)((g->player + 1) % NP);
  g->cur = (uint8_t)nw;
  uint32_t* w = reinterpret_cast<uint32_t*>(&g->lu[0][0]);
#pragma unroll 1
  for (int i = 0; i < (NL * LU_ROW + NS * SU_ROW) / 4; i++) w[i] = 0;
  uint32_t work = 0;
#pragma unroll 1
  for (int l = 0; l < NL; l++) {
    const uint8_t* mine = utl(g, nw, l);
    uint8_t* row = g->lu[l];
    const uint32_t f = mine[FIGHTERS], b = mine[BOMBERS_LAND_AIR], i = mine[INFANTRY], a = mine[ARTILLERY], tk = mine[TANKS], aa = mine[AA_GUNS];
    row[off_lu(FIGHTERS) + 4] = (uint8_t)f;
    row[off_lu(BOMBERS_LAND_AIR) + 6] = (uint8_t)b;
    row[off_lu(INFANTRY) + 1] = (uint8_t)i;
    row[off_lu(ARTILLERY) + 1] = (uint8_t)a;
    row[off_lu(TANKS) + 2] = (uint8_t)tk;
    row[off_lu(AA_GUNS) + 1] = (uint8_t)aa;
    work |= (f ? W(CAAA_PH_MOVE_FIGHTERS) : 0u) | (b ? W(CAAA_PH_MOVE_BOMBERS) : 0u) | (i ? W(CAAA_PH_MOVE_INFANTRY) : 0u) |
            (a ? W(CAAA_PH_MOVE_ARTILLERY) : 0u) | (tk ? W(CAAA_PH_MOVE_TANKS) : 0u) | (aa ? W(CAAA_PH_MOVE_AA_GUNS) : 0u);
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    const uint8_t* mine = uts(g, nw, s);
    uint8_t* row = g->su[s];
    uint32_t stage = 0, loaded = 0, ships = 0;
#pragma unroll
    for (int u = 0; u < NST - 1; u++) {
      const uint32_t n = mine[u];
      row[off_sea(u) + states_sea(u) - 1] = (uint8_t)n;
      if (u >= TRANS_EMPTY && u <= TRANS_1T) stage |= n;
      if (u >= TRANS_2I && u <= TRANS_1I_1T) loaded |= n;
      if (u >= DESTROYERS && u <= BS_DAMAGED) ships |= n;
    }
    const uint32_t sb = g->cur_sea_bombers[s];  // 3944: only 14 of 15 types rotate, the sea-bomber total stays
    row[off_sea(BOMBERS_SEA) + 4] = (uint8_t)sb;
    work |= (mine[FIGHTERS] ? W(CAAA_PH_MOVE_FIGHTERS) : 0u) | (stage ? W(CAAA_PH_STAGE_TRANSPORTS) : 0u) |
            (loaded ? W(CAAA_PH_MOVE_TRANSPORTS) : 0u) | (mine[SUBMARINES] ? W(CAAA_PH_MOVE_SUBS) : 0u) |
            (ships ? W(CAAA_PH_MOVE_SHIPS) : 0u) | (sb ? W(CAAA_PH_LAND_BOMBERS) : 0u);
  }
  g->work = work;
  uint32_t* fl = reinterpret_cast<uint32_t*>(g->flagged);
  fl[0] = fl[1] = 0;
#pragma unroll 1
  for (int f = 0; f < g->nfact[nw]; f++) {
    const int land = g->factloc[nw][f];
    const uint8_t fm = g->factory_max[land];
    g->builds_left[land] = fm;
#pragma unroll 1
    for (int i = 0; i < CAAA_TAB(cx)->t.l2s_count[land]; i++) {
      uint8_t* b = &g->builds_left[CAAA_TAB(cx)->t.l2s_conn[land][i] + NL];
      *b = (uint8_t)(*b + fm);
    }
  }
  refresh_eot_cache(g, cx);
}

// ---- scoring, 4301-4337 ----
// cost vectors of one player's 72 unit counts: 5 x land costs (6 types) then 3 x sea costs (types 0..13)
CAAA_HD constexpr uint32_t cost_byte(int i) { return i < 30 ? cost_land(i % 6) : cost_sea((i - 30) % 14); }
CAAA_HD constexpr uint32_t cost_vec(int w) {
  return cost_byte(4 * w) | (cost_byte(4 * w + 1) << 8) | (cost_byte(4 * w + 2) << 16) | (cost_byte(4 * w + 3) << 24);
}
CAAA_HD_NOINLINE double evaluate_state(Game* g, const RunCtx cx) {
  const uint32_t starting = g->player;
  const int pi = (int)(starting >= (uint32_t)NP ? starting - NP : starting) % NP;
  const int cur = g->cur;
  int allied_score = 1 + g->money[cur], enemy_score = 1;  // 4309+4311: the mover's money counts twice
#pragma unroll 1
  for (int p = 0; p < NP; p++) {
    const int pa = (cur + p) % NP;
    uint32_t score = g->money[pa];
    const uint32_t* v = reinterpret_cast<const uint32_t*>(g->ut[pa]);
#pragma unroll
    for (int w = 0; w < UT_BYTES / 4; w++) score = dot4(v[w], cost_vec(w), score);
    // 4319: 15 entries of a 14-entry row: the current player's are its sea bombers, everyone else's run into
    // the next row's first entry (fighters) / flagged_for_combat[0] after the very last row
    uint32_t extra = 0;
    if (p == 0) {
      extra = g->cur_sea_bombers[0] + g->cur_sea_bombers[1] + g->cur_sea_bombers[2];
    } else {
      extra = uts(g, pa, 1)[FIGHTERS] + uts(g, pa, 2)[FIGHTERS];
      extra += p < NP - 1 ? uts(g, (pa + 1) % NP, 0)[FIGHTERS] : g->flagged[0];
    }
    score += extra * cost_sea(BOMBERS_SEA);
    if (CAAA_TAB(cx)->t.is_allied[pi][(pi + p) % NP]) allied_score += (int)score;
    else enemy_score += (int)score;
  }
  const double score = (double)allied_score / (double)(enemy_score + allied_score);
  return starting >= (uint32_t)NP ? 1 - score : score;
}

// One phase of the pipeline by id (the order is fixed: reference src/engine.cpp:4214-4235).
// `m` = the lanes of the warp that make this call together (ignored in stop-at-decision mode and on the host)
template <bool EXPAND>
CAAA_HD int run_phase(Game* g, const RunCtx cx, int ph, unsigned m = 0xffffffffu, int yield = 0) {
  switch (ph) {
    case CAAA_PH_MOVE_FIGHTERS: return move_fighter_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_MOVE_BOMBERS: return move_bomber_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_STAGE_TRANSPORTS: return stage_transport_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_MOVE_TANKS: return move_land_unit_type<EXPAND>(g, cx, m, TANKS, ph, yield);
    case CAAA_PH_MOVE_ARTILLERY: return move_land_unit_type<EXPAND>(g, cx, m, ARTILLERY, ph, yield);
    case CAAA_PH_MOVE_INFANTRY: return move_land_unit_type<EXPAND>(g, cx, m, INFANTRY, ph, yield);
    case CAAA_PH_MOVE_TRANSPORTS: return move_transport_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_MOVE_SUBS: return move_subs<EXPAND>(g, cx, m, yield);
    case CAAA_PH_MOVE_SHIPS: return move_destroyers_battleships<EXPAND>(g, cx, m, yield);
    case CAAA_PH_SEA_BATTLES: return resolve_sea_battles<EXPAND>(g, cx, m, yield);
    case CAAA_PH_UNLOAD_TRANSPORTS: return unload_transports<EXPAND>(g, cx, m, yield);
    case CAAA_PH_LAND_BATTLES: return resolve_land_battles<EXPAND>(g, cx, m, yield);
    case CAAA_PH_MOVE_AA_GUNS: return move_land_unit_type<EXPAND>(g, cx, m, AA_GUNS, ph, yield);
    case CAAA_PH_LAND_FIGHTERS: return land_fighter_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_LAND_BOMBERS: return land_bomber_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_BUY_UNITS: return buy_units<EXPAND>(g, cx, m, yield);
    case CAAA_PH_CRASH_AIR: crash_air_units(g, cx); return PH_DONE;
    case CAAA_PH_RESET_UNITS: reset_units_fully(g); return PH_DONE;
    case CAAA_PH_BUY_FACTORY: return PH_DONE;  // buy_factory is empty (3889)
    case CAAA_PH_COLLECT_MONEY: collect_money(g, cx); return PH_DONE;
    case CAAA_PH_ROTATE_TURNS: rotate_turns(g, cx); return PH_DONE;
    default: return PH_DONE;
  }
}

// Exact work mask of an arbitrary (imported) state: bit ph is set when phase ph would touch the game.  The
// pipeline keeps the mask up to date incrementally afterwards; begin_common() may also start from the
// conservative all-ones mask (every phase then looks once and clears its own bit).
CAAA_HD uint32_t scan_work(const Game* g) {
  uint32_t w = 0;
#pragma unroll 1
  for (int l = 0; l < NL; l++) {
    const uint8_t* r = g->lu[l];
    const uint8_t* f = r + off_lu(FIGHTERS);
    const uint8_t* b = r + off_lu(BOMBERS_LAND_AIR);
    if (f[4]) w |= W(CAAA_PH_MOVE_FIGHTERS);
    if (f[1] | f[2] | f[3]) w |= W(CAAA_PH_LAND_FIGHTERS);
    if (b[6]) w |= W(CAAA_PH_MOVE_BOMBERS);
    if (b[1] | b[2] | b[3] | b[4] | b[5]) w |= W(CAAA_PH_LAND_BOMBERS);
    if (r[off_lu(INFANTRY) + 1]) w |= W(CAAA_PH_MOVE_INFANTRY);
    if (r[off_lu(ARTILLERY) + 1]) w |= W(CAAA_PH_MOVE_ARTILLERY);
    if (r[off_lu(TANKS) + 1] | r[off_lu(TANKS) + 2]) w |= W(CAAA_PH_MOVE_TANKS);
    if (r[off_lu(AA_GUNS) + 1]) w |= W(CAAA_PH_MOVE_AA_GUNS);
    if (g->flagged[l] != 0) w |= W_LAND_BATTLE;  // units may still arrive before the battle phase
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    const uint8_t* r = g->su[s];
    const uint8_t* f = r + off_sea(FIGHTERS);
    const uint8_t* b = r + off_sea(BOMBERS_SEA);
    if (f[4]) w |= W(CAAA_PH_MOVE_FIGHTERS);
    if (f[1] | f[2] | f[3]) w |= W(CAAA_PH_LAND_FIGHTERS);
    if (b[0] | b[1] | b[2] | b[3] | b[4]) w |= W(CAAA_PH_LAND_BOMBERS);
    const uint8_t* e = r + off_sea(TRANS_EMPTY);
    if (e[3]) w |= W(CAAA_PH_STAGE_TRANSPORTS);
    if (e[1] | e[2] | e[3]) w |= W(CAAA_PH_MOVE_TRANSPORTS);  // skip_empty_transports
#pragma unroll 1
    for (int ut = TRANS_1I; ut <= TRANS_1I_1T; ut++) {
      const uint8_t* t = r + off_sea(ut);
      if (ut <= TRANS_1T && t[4]) w |= W(CAAA_PH_STAGE_TRANSPORTS);
      if (t[2] | t[3]) w |= W(CAAA_PH_MOVE_TRANSPORTS);
      if (t[1]) w |= W(CAAA_PH_UNLOAD_TRANSPORTS);
    }
    if (r[off_sea(SUBMARINES) + 2]) w |= W(CAAA_PH_MOVE_SUBS);
    if (r[off_sea(DESTROYERS) + 1] | r[off_sea(CARRIERS) + 1] | r[off_sea(CRUISERS) + 2] | r[off_sea(BATTLESHIPS) + 2] |
        r[off_sea(BS_DAMAGED) + 2])
      w |= W(CAAA_PH_MOVE_SHIPS);
    if (g->flagged[s + NL] != 0) w |= W_SEA_BATTLE;
  }
  return w;
}

// Everything a freshly imported state needs before the pipeline can run on it.
CAAA_HD void begin_common(Game* g, const RunCtx cx) {
  // entries of the path-blocked tables that no refresh ever writes are read by the flat-index quirk
  // (add_valid_land_moves); they are zero in the reference because its globals are zero-initialised
  g->land_blocked = 0;
  g->seasub_blocked = 0;
#pragma unroll 1
  for (int p = 0; p < NP; p++)
#pragma unroll 1
    for (int l = 0; l < NL; l++) g->factloc[p][l] = 0;
  refresh_full_cache(g, cx);
  g->work = scan_work(g);
  g->cur_phase = CUR_NONE;
  g->unlucky = 0;
  g->draws = 0;
  g->turns = 0;
  g->status = 0;
}
// what random_play_until_terminal does before its loop (4185-4203, with P1 iv and the P2 stream reset)
CAAA_HD void begin_playout(Game* g, const RunCtx cx, uint64_t game_id) {
  begin_common(g, cx);
  g->answers_remaining = 100000;
  g->use_selected = 0;
  g->selected_action = 0;
  g->game_lo = (uint32_t)game_id;
  g->game_hi = (uint32_t)(game_id >> 32);
#pragma unroll 1
  for (int i = 0; i < CAAA_ACTIONS; i++) g->vm[i] = 0;
  g->nvm = 0;
}

CAAA_HD uint64_t hash_live_state(const Game* g) {  // FNV-1a-64 over the first 606 bytes of the reference layout
  uint8_t buf[CAAA_STATE_BYTES];
  export_state(g, buf);
  uint64_t h = CAAA_FNV_OFFSET;
#pragma unroll 1
  for (int i = 0; i < CAAA_STATE_LIVE_BYTES; i++) {
    h ^= buf[i];
    h *= CAAA_FNV_PRIME;
  }
  return h;
}

}  // namespace e2
}  // namespace caaa
