// game.cuh -- per-game working set ("Game") of the rollout engine, import/export to the reference layout.
//
// The reference keeps a 670-byte GameState in which every per-player array is *rotated* by one
// player each turn (reference src/engine.cpp:3895-4029) plus ~600 bytes of derived caches in
// file-scope globals (src/engine.cpp:120-159).  On the GPU the rotation is pure data movement that
// every lane would repeat every ply, so a Game stores per-player data in ABSOLUTE player order and
// never rotates; "relative player p" of the reference is absolute player (cur + p) % 5.  The
// reference layout only exists at the ABI: import_state() / export_state() convert.
//
// Layout rules: 932 bytes = 233 words (odd, so the 32 lanes of a warp touching the same field of 32
// games hit 32 different banks); every array that is scanned word-wise starts on a word boundary;
// per-player unit totals are one 72-byte vector per player (30 land + 42 sea counts) so that scoring
// is 18 dp4a per player against a constant cost vector.
#pragma once
#include <cstddef>
#include "philox.cuh"
#include "slot.cuh"

namespace caaa {
namespace e2 {

constexpr int UT_BYTES = NL * NLT + NS * (NST - 1);  // 72: [land][6 types] then [sea][14 types]
constexpr int LU_ROW = 21, SU_ROW = 57;              // per-location bytes of the mover's per-state counts

struct Game {
  uint32_t rng[4];            // current 16-byte Philox block
  uint32_t draws;             // random bytes consumed so far (= Philox draw index)
  uint32_t game_lo, game_hi;
  int32_t answers_remaining;
  uint32_t work;              // bit ph set => pipeline phase ph (0..14) may have something to do
  uint32_t land_blocked;      // is_land_path_blocked, bit src*5+dst (flat, 25 bits)
  uint32_t seasub_blocked;    // bits 0..8 is_sea_path_blocked[src*3+dst], bits 16..24 is_sub_path_blocked
  uint32_t skipped[2];        // skipped_moves as 64 one-bit flags, flat index a*8+b
  uint16_t turns, status;
  int16_t tot_land[NP][NL];   // total_player_land_units, absolute player
  int16_t tot_sea[NP][NS];    // total_player_sea_units
  int16_t enemy_count[NA];    // enemy_units_count
  int16_t income[NP];         // income_per_turn
  int16_t large_cargo[NS], small_cargo[NS], carriers[NS];
  uint8_t ut[NP][UT_BYTES];   // units per type: ut[player][land*6+type], ut[player][30+sea*14+type]
  uint8_t lu[NL][LU_ROW];     // player to move: land units per movement state (reference LandState bytes 4..24)
  uint8_t su[NS][SU_ROW];     // player to move: sea units per movement state (reference UnitsSea)
  uint8_t flagged[NA];        // flagged_for_combat
  uint8_t builds_left[NA];
  uint8_t money[NP];          // absolute player
  uint8_t owner[NL];          // absolute owner
  int8_t factory_hp[NL];
  uint8_t factory_max[NL];
  uint8_t bombard_max[NL];
  int8_t nfact[NP];           // total_factory_count
  uint8_t factloc[NP][NL];    // factory_locations (ordered lists)
  uint8_t cur_sea_bombers[NS];  // current_player_sea_unit_types[s][BOMBERS_SEA]: never rotated (src/engine.cpp:3944)
  uint8_t blockade[NS], edestroyers[NS];
  uint8_t vm[CAAA_ACTIONS];   // valid_moves (persists across phases and turns, src/engine.cpp:158)
  uint8_t nvm;
  uint8_t player;             // player_index as stored (may exceed 4 in caller-supplied states)
  uint8_t cur;                // player % 5
  uint8_t canal_state;
  uint8_t use_selected, selected_action, unlucky;  // expansion mode (src/engine.cpp:156,160,163)
  uint8_t pad_[1];
  // loop cursor of a phase that yielded before it was done (phase-regrouped kernel): lets the phase resume
  // exactly where it stopped, with its valid-moves list (vm/nvm above) and its per-phase masks intact
  uint8_t cur_phase, cur_cell, cur_flags, cur_a;
  uint32_t cur_b;
};
static_assert(sizeof(Game) == 932, "Game layout");
static_assert((sizeof(Game) / 4) % 2 == 1, "slot stride must be an odd number of words");
static_assert(offsetof(Game, ut) % 4 == 0 && offsetof(Game, lu) % 4 == 0 && offsetof(Game, flagged) % 4 == 0, "word alignment");
static_assert((NL * LU_ROW + NS * SU_ROW) % 4 == 0, "per-state counts are cleared word-wise");

struct RunCtx {
  const DevTables* T;
  uint32_t seed_lo, seed_hi;
};

// The read-only tables: on the device always the CTA's copy at the start of dynamic shared memory (every
// kernel stages it there first), so the compiler emits shared-memory loads instead of generic ones.
#if defined(__CUDACC__)
extern __shared__ __align__(16) uint8_t smem_raw[];
#endif
#if defined(__CUDA_ARCH__)
#define CAAA_TAB(cx) (reinterpret_cast<const DevTables*>(smem_raw))
#else
#define CAAA_TAB(cx) ((cx).T)
#endif

// Warp votes that keep the lanes of a phase loop in lockstep: every lane of `m` stays in the loop until the
// last one is done, and all of them meet again at the vote at the top of every iteration, so lanes that took
// different paths through one decision step execute the next one together.  Without the vote a lane that
// finishes its step early runs ahead and the warp never reconverges.  V = false (stop-at-decision mode, where
// every lane stops at its own decision, and the host build) degenerates to the plain per-game loop.
template <bool V>
CAAA_HD bool vote_any(unsigned m, bool p) {
#if defined(__CUDA_ARCH__)
  if (V) return __any_sync(m, p);
#endif
  return p;
}
template <bool V>
CAAA_HD unsigned vote_ballot(unsigned m, bool p) {
#if defined(__CUDA_ARCH__)
  if (V) return __ballot_sync(m, p);
#endif
  return p ? m : 0u;
}

constexpr uint8_t CUR_NONE = 0xff;
constexpr uint32_t CF_FRESH = 1, CF_PROCESSED = 2, CF_IN_BATTLE = 4, CF_SUBMERGED = 8, CF_NEW_FACTORY = 16, CF_NEW_CELL = 32;
CAAA_HD void save_cursor(Game* g, int ph, int cell, uint32_t flags, uint32_t a, uint32_t b) {
  g->cur_phase = (uint8_t)ph;
  g->cur_cell = (uint8_t)cell;
  g->cur_flags = (uint8_t)flags;
  g->cur_a = (uint8_t)a;
  g->cur_b = b;
}
// Phase results: 0 = ran to completion, 1 = stopped at a decision (stop-at-decision mode, or the decision
// budget of a rollout is exhausted), 2 = yielded with its cursor saved (call again to continue).
constexpr int PH_DONE = 0, PH_STOPPED = 1, PH_YIELDED = 2;
// A lockstep loop yields when `yield` of its lanes have finished (device; 0 = never), so that the caller can
// replace them with other games waiting for the same phase instead of idling them until the slowest lane is
// done.  On the host (tests) `yield` is a plain budget of loop iterations.
template <bool V>
CAAA_HD bool should_yield(unsigned m, bool active, int yield, int& steps) {
  if (yield <= 0) return false;
#if defined(__CUDA_ARCH__)
  if (V) return __popc(__ballot_sync(m, !active)) >= yield;
#endif
  return steps++ >= yield;
}

// Re-joins the lanes of `mask` (the lanes that entered this loop iteration) after a loop in which they ran
// different trip counts; the compiler does not place a reconvergence point there by itself.
template <bool V>
CAAA_HD void converge(unsigned mask) {
#if defined(__CUDA_ARCH__)
  if (V) __syncwarp(mask);
#endif
  (void)mask;
}

// ---- accessors ----
CAAA_HD constexpr uint32_t off_lu(int t) { return off_land(t) - 4u; }
CAAA_HD uint8_t* lu_ptr(Game* g, int l, int t) { return g->lu[l] + off_lu(t); }
CAAA_HD uint8_t* su_ptr(Game* g, int s, int t) { return g->su[s] + off_sea(t); }
CAAA_HD uint8_t* utl(Game* g, int pa, int l) { return g->ut[pa] + l * NLT; }
CAAA_HD uint8_t* uts(Game* g, int pa, int s) { return g->ut[pa] + NL * NLT + s * (NST - 1); }
CAAA_HD const uint8_t* utl(const Game* g, int pa, int l) { return g->ut[pa] + l * NLT; }
CAAA_HD const uint8_t* uts(const Game* g, int pa, int s) { return g->ut[pa] + NL * NLT + s * (NST - 1); }
// current_player_sea_unit_types[s][t]; type 14 (sea bombers) lives outside the per-player vectors
CAAA_HD uint8_t* cur_sea(Game* g, int s, int t) { return t == BOMBERS_SEA ? &g->cur_sea_bombers[s] : uts(g, g->cur, s) + t; }
CAAA_HD uint8_t* air_fighters(Game* g, int air) { return air < NL ? lu_ptr(g, air, FIGHTERS) : su_ptr(g, air - NL, FIGHTERS); }
CAAA_HD uint8_t* air_bombers(Game* g, int air) { return air < NL ? lu_ptr(g, air, BOMBERS_LAND_AIR) : su_ptr(g, air - NL, BOMBERS_SEA); }
CAAA_HD bool allied(const Game* g, const RunCtx& cx, int pa) { return (cx.T->allied_abs[g->cur] >> pa) & 1; }
CAAA_HD bool owner_allied(const Game* g, const RunCtx& cx, int land) { return allied(g, cx, g->owner[land]); }
CAAA_HD int owner_rel(const Game* g, int land) { return (g->owner[land] + NP - g->cur) % NP; }

CAAA_HD bool land_blocked(const Game* g, unsigned flat) { return flat < 25u && ((g->land_blocked >> flat) & 1u); }
CAAA_HD bool sea_blocked(const Game* g, unsigned flat) { return (g->seasub_blocked >> flat) & 1u; }
CAAA_HD bool sub_blocked(const Game* g, unsigned flat) { return (g->seasub_blocked >> (16u + flat)) & 1u; }

// skipped_moves as a flat array of 64 one-bit flags (P1 ii): out-of-range reads 0, writes dropped
CAAA_HD bool skip_get(const Game* g, unsigned a, unsigned b) {
  const unsigned i = a * 8u + b;
  return i < 64u && ((g->skipped[i >> 5] >> (i & 31u)) & 1u);
}
CAAA_HD void skip_set(Game* g, unsigned a, unsigned b) {
  const unsigned i = a * 8u + b;
  if (i < 64u) g->skipped[i >> 5] |= 1u << (i & 31u);
}
CAAA_HD uint32_t skip_row(const Game* g, unsigned a) { return a < 8u ? (g->skipped[a >> 2] >> ((a & 3u) * 8u)) & 0xffu : 0u; }
CAAA_HD void clear_move_history(Game* g) { g->skipped[0] = g->skipped[1] = 0; }  // engine.cpp:1330

constexpr uint32_t W_SEA_BATTLE = 1u << CAAA_PH_SEA_BATTLES, W_LAND_BATTLE = 1u << CAAA_PH_LAND_BATTLES;
CAAA_HD void flag_combat(Game* g, int air, uint8_t v) {
  g->flagged[air] = v;
  if (v) g->work |= air < NL ? W_LAND_BATTLE : W_SEA_BATTLE;
}

// 4 x u8 dot product (DP4A on the device)
CAAA_HD uint32_t dot4(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __dp4a(a, b, c);
#else
#pragma unroll 1
  for (int i = 0; i < 4; i++) c += ((a >> (8 * i)) & 0xffu) * ((b >> (8 * i)) & 0xffu);
  return c;
#endif
}

// ---- reference layout <-> Game ----
// N bytes src -> dst with every load issued before the first store: the two pointers may alias as far as the
// compiler knows, and a byte-by-byte copy would wait a full shared-memory round trip per byte.
template <int N>
CAAA_HD void copy_block(uint8_t* dst, const uint8_t* src) {
  uint8_t t[N];
#pragma unroll
  for (int k = 0; k < N; k++) t[k] = src[k];
#pragma unroll
  for (int k = 0; k < N; k++) dst[k] = t[k];
}

// Fills everything that the reference keeps in GameState; derived caches are rebuilt by refresh_full_cache().
CAAA_HD void import_state(Game* g, const uint8_t* src) {
  const CaaaGameState* st = reinterpret_cast<const CaaaGameState*>(src);
  const int player = st->player_index, cur = player % NP;
  g->player = (uint8_t)player;
  g->cur = (uint8_t)cur;
  {
    uint8_t m[NP];
#pragma unroll
    for (int p = 0; p < NP; p++) m[p] = st->money[p];
    int pa = cur;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      g->money[pa] = m[p];
      pa = pa + 1 == NP ? 0 : pa + 1;
    }
  }
  copy_block<NA>(g->builds_left, st->builds_left);
  copy_block<NA>(g->flagged, st->flagged_for_combat);
  {
    uint8_t h[NL][4];  // owner, factory hp / max, bombard max of the five lands
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
      for (int k = 0; k < 4; k++) h[l][k] = reinterpret_cast<const uint8_t*>(&st->land_state[l])[k];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int o = h[l][0] + cur;
      g->owner[l] = (uint8_t)(o >= NP ? o % NP : o);
      g->factory_hp[l] = (int8_t)h[l][1];
      g->factory_max[l] = h[l][2];
      g->bombard_max[l] = h[l][3];
    }
  }
#pragma unroll 1
  for (int l = 0; l < NL; l++) copy_block<LU_ROW>(g->lu[l], reinterpret_cast<const uint8_t*>(&st->land_state[l]) + 4);
#pragma unroll 1
  for (int s = 0; s < NS; s++) copy_block<SU_ROW>(g->su[s], reinterpret_cast<const uint8_t*>(&st->units_sea[s]));
  {
    int pa = cur;
#pragma unroll 1
    for (int p = 1; p < NP; p++) {  // one 30-byte and one 42-byte vector per other player
      pa = pa + 1 == NP ? 0 : pa + 1;
      copy_block<NL * NLT>(g->ut[pa], &st->other_land_units[p - 1][0][0]);
      copy_block<NS*(NST - 1)>(g->ut[pa] + NL * NLT, &st->other_sea_units[p - 1][0][0]);
    }
  }
  const uint8_t* sm = &st->skipped_moves[0][0];
  uint32_t sk[2];
#pragma unroll 1
  for (int w = 0; w < 2; w++) {
    uint8_t t[32];
#pragma unroll
    for (int i = 0; i < 32; i++) t[i] = sm[w * 32 + i];
    uint32_t v = 0;
#pragma unroll
    for (int i = 0; i < 32; i++) v |= (uint32_t)(t[i] & 1u) << i;
    sk[w] = v;
  }
  g->skipped[0] = sk[0];
  g->skipped[1] = sk[1];
}

CAAA_HD void export_state(const Game* g, uint8_t* dst) {
  CaaaGameState* st = reinterpret_cast<CaaaGameState*>(dst);
  const int cur = g->cur;
  st->player_index = g->player;
  {
    uint8_t m[NP];
    int pa = cur;
#pragma unroll
    for (int p = 0; p < NP; p++) {
      m[p] = g->money[pa];
      pa = pa + 1 == NP ? 0 : pa + 1;
    }
#pragma unroll
    for (int p = 0; p < NP; p++) st->money[p] = m[p];
  }
  copy_block<NA>(st->builds_left, g->builds_left);
  copy_block<NA>(st->flagged_for_combat, g->flagged);
  {
    uint8_t h[NL][4];
#pragma unroll
    for (int l = 0; l < NL; l++) {
      const int o = g->owner[l] + NP - cur;
      h[l][0] = (uint8_t)(o >= NP ? o - NP : o);
      h[l][1] = (uint8_t)g->factory_hp[l];
      h[l][2] = g->factory_max[l];
      h[l][3] = g->bombard_max[l];
    }
#pragma unroll
    for (int l = 0; l < NL; l++)
#pragma unroll
      for (int k = 0; k < 4; k++) reinterpret_cast<uint8_t*>(&st->land_state[l])[k] = h[l][k];
  }
#pragma unroll 1
  for (int l = 0; l < NL; l++) copy_block<LU_ROW>(reinterpret_cast<uint8_t*>(&st->land_state[l]) + 4, g->lu[l]);
#pragma unroll 1
  for (int s = 0; s < NS; s++) copy_block<SU_ROW>(reinterpret_cast<uint8_t*>(&st->units_sea[s]), g->su[s]);
  {
    int pa = cur;
#pragma unroll 1
    for (int p = 1; p < NP; p++) {
      pa = pa + 1 == NP ? 0 : pa + 1;
      copy_block<NL * NLT>(&st->other_land_units[p - 1][0][0], g->ut[pa]);
      copy_block<NS*(NST - 1)>(&st->other_sea_units[p - 1][0][0], g->ut[pa] + NL * NLT);
    }
  }
  uint8_t* sm = &st->skipped_moves[0][0];
#pragma unroll 1
  for (int w = 0; w < 2; w++) {
    const uint32_t v = g->skipped[w];
#pragma unroll
    for (int i = 0; i < 32; i++) sm[w * 32 + i] = (uint8_t)((v >> i) & 1u);
  }
}

// the reference's derived globals in the fixed wire layout of CaaaCacheDump (relative player order)
CAAA_HD void export_cache(const Game* g, const DevTables* T, CaaaCacheDump* d) {
  const int cur = g->cur;
#pragma unroll 1
  for (int p = 0; p < 6; p++) {
    const int pa = (cur + p) % NP;
    const bool real = p < NP;
    d->income_per_turn[p] = real ? g->income[pa] : 0;
    d->total_factory_count[p] = real ? g->nfact[pa] : 0;
#pragma unroll 1
    for (int l = 0; l < NL; l++) {
      d->factory_locations[p][l] = real ? g->factloc[pa][l] : 0;
      d->total_player_land_units[p][l] = real ? g->tot_land[pa][l] : 0;
    }
#pragma unroll 1
    for (int s = 0; s < NS; s++) d->total_player_sea_units[p][s] = real ? g->tot_sea[pa][s] : 0;
  }
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
    d->transports_with_large_cargo_space[s] = g->large_cargo[s];
    d->transports_with_small_cargo_space[s] = g->small_cargo[s];
    d->allied_carriers[s] = g->carriers[s];
    d->enemy_blockade_total[s] = g->blockade[s];
    d->enemy_destroyers_total[s] = g->edestroyers[s];
  }
#pragma unroll 1
  for (int a = 0; a < NA; a++) d->enemy_units_count[a] = g->enemy_count[a];
  d->answers_remaining = g->answers_remaining;
#pragma unroll 1
  for (int l = 0; l < NL; l++)
#pragma unroll 1
    for (int u = 0; u < NLT; u++) d->current_player_land_unit_types[l][u] = utl(g, cur, l)[u];
#pragma unroll 1
  for (int s = 0; s < NS; s++) {
#pragma unroll 1
    for (int u = 0; u < NST - 1; u++) d->current_player_sea_unit_types[s][u] = uts(g, cur, s)[u];
    d->current_player_sea_unit_types[s][BOMBERS_SEA] = g->cur_sea_bombers[s];
  }
#pragma unroll 1
  for (unsigned i = 0; i < 25; i++) d->is_land_path_blocked[i] = land_blocked(g, i);
#pragma unroll 1
  for (unsigned i = 0; i < 9; i++) {
    d->is_sea_path_blocked[i] = sea_blocked(g, i);
    d->is_sub_path_blocked[i] = sub_blocked(g, i);
  }
#pragma unroll 1
  for (int e = 0; e < 4; e++) d->enemies_0[e] = e < T->n_enemies[cur] ? T->enemies[cur][e] : 0;
#pragma unroll 1
  for (int p = 0; p < NP; p++) d->is_allied_0[p] = (T->allied_mask[cur] >> p) & 1;
  d->enemies_count_0 = T->n_enemies[cur];
  d->canal_state = g->canal_state;
#pragma unroll 1
  for (int i = 0; i < CAAA_ACTIONS; i++) d->valid_moves[i] = g->vm[i];
  d->valid_moves_count = g->nvm;
  d->pad_[0] = 0;
}

}  // namespace e2
}  // namespace caaa
