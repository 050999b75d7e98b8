// caaa_gpu.cu -- sm_100a kernels and the C-ABI of libcaaa_gpu.so (include/caaa_gpu.h).
//
// Execution model: one persistent CTA per SM; 192 game slots and the topology tables stay
// resident in the CTA's 227 KB of shared memory for the whole playout, so the only HBM traffic of a
// playout is its 670-byte start state in and its result out.  Games are regrouped by phase inside the
// CTA: every slot carries the id of the next pipeline phase that has work for it, and each warp
// repeatedly claims up to 32 slots waiting for the same phase (the fullest queue first), so the 32
// lanes of a warp run the same phase function on 32 games that all have something to do in it.
// Finished games are replaced from a global ticket counter, which absorbs the long tail of game
// lengths.  No tensor cores: the work is branchy uint8 arithmetic.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/caaa_gpu.h"
#include "engine.cuh"
#include "host_tables.h"

namespace caaa {

constexpr int TABLE_BYTES = (int)((sizeof(DevTables) + 15) / 16 * 16);
constexpr int ROLLOUT_THREADS = 192;  // slots per CTA: 1328 + 192 * 1180 = 227,888 B of shared memory
constexpr int SMALL_THREADS = 128;    // non-persistent kernels

extern __shared__ __align__(16) uint8_t smem_raw[];

__device__ __forceinline__ const DevTables* stage_tables(const DevTables* __restrict__ src) {
  // read-only tables: one copy per CTA in shared memory (divergent indices would serialise in the constant cache)
  const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
  uint16_t* d = reinterpret_cast<uint16_t*>(smem_raw);
  for (int i = threadIdx.x; i < (int)(sizeof(DevTables) / 2); i += blockDim.x) d[i] = s[i];
  __syncthreads();
  return reinterpret_cast<const DevTables*>(smem_raw);
}
__device__ __forceinline__ GameSlot* my_slot() { return reinterpret_cast<GameSlot*>(smem_raw + TABLE_BYTES) + threadIdx.x; }

__device__ __forceinline__ void load_state(GameSlot* g, const uint8_t* __restrict__ src) {
  const uint16_t* s = reinterpret_cast<const uint16_t*>(src);  // 670 bytes, 2-byte aligned
  uint16_t* d = reinterpret_cast<uint16_t*>(&g->st);
#pragma unroll 5
  for (int i = 0; i < CAAA_STATE_BYTES / 2; i++) d[i] = s[i];
}
__device__ __forceinline__ void store_state(const GameSlot* g, uint8_t* __restrict__ dst) {
  const uint16_t* s = reinterpret_cast<const uint16_t*>(&g->st);
  uint16_t* d = reinterpret_cast<uint16_t*>(dst);
#pragma unroll 5
  for (int i = 0; i < CAAA_STATE_BYTES / 2; i++) d[i] = s[i];
}

// The fixed pipeline of one player-turn, reference src/engine.cpp:4214-4235.
template <bool EXPAND>
__device__ __forceinline__ void play_turn_rollout(GameSlot* g, const RunCtx& cx) {
  move_fighter_units<EXPAND>(g, cx);
  move_bomber_units<EXPAND>(g, cx);
  stage_transport_units<EXPAND>(g, cx);
  move_land_unit_type<EXPAND>(g, cx, TANKS);
  move_land_unit_type<EXPAND>(g, cx, ARTILLERY);
  move_land_unit_type<EXPAND>(g, cx, INFANTRY);
  move_transport_units<EXPAND>(g, cx);
  move_subs<EXPAND>(g, cx);
  move_destroyers_battleships<EXPAND>(g, cx);
  resolve_sea_battles<EXPAND>(g, cx);
  unload_transports<EXPAND>(g, cx);
  resolve_land_battles<EXPAND>(g, cx);
  move_land_unit_type<EXPAND>(g, cx, AA_GUNS);
  land_fighter_units<EXPAND>(g, cx);
  land_bomber_units<EXPAND>(g, cx);
  buy_units<EXPAND>(g, cx);
  crash_air_units(g, cx);
  reset_units_fully(g);
  collect_money(g, cx);
  rotate_turns(g, cx);
}

struct RolloutArgs {
  const DevTables* tables;
  const uint8_t* states;  // n_states x 670 bytes
  unsigned long long n_games;
  unsigned int rollouts_per_state;
  unsigned long long seed, first_game_id;
  double* results;
  CaaaRolloutStats* stats;  // optional
  CaaaBatchSummary* summary;  // optional
  unsigned long long* ticket;
  int debug;
  int groups;  // barrier sub-groups per CTA (block-synchronous kernel)
};

template <bool STATS>
__global__ void __launch_bounds__(ROLLOUT_THREADS, 1) rollout_kernel(const RolloutArgs a) {
  const DevTables* T = stage_tables(a.tables);
  GameSlot* g = my_slot();
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  bool live = false, exhausted = false;
  unsigned long long gi = 0;
  double score = 0.0;
  unsigned long long n_play = 0, n_plies = 0, n_dec = 0, n_draws = 0, n_flag = 0;
  double sum = 0.0;
  for (;;) {
    if (!live && !exhausted) {  // refill at a turn boundary
      gi = atomicAdd(a.ticket, 1ULL);
      if (gi < a.n_games) {
        load_state(g, a.states + (gi / a.rollouts_per_state) * CAAA_STATE_BYTES);
        begin_playout(g, cx, a.first_game_id + gi);
        score = evaluate_state(g, cx);
        live = true;
      } else {
        exhausted = true;
      }
    }
    if (!__any_sync(0xffffffffu, live)) break;
    if (live) {
      if (score > 0.01 && score < 0.99 && g->turns < 1000) {
        play_turn_rollout<false>(g, cx);
        g->turns++;
        score = evaluate_state(g, cx);
      } else {  // reference src/engine.cpp:4238-4241
        const double r = (g->st.player_index % 2 == 1) ? 1.0 - score : score;
        a.results[gi] = r;
        const int decisions = 100000 - g->answers_remaining;
        if (STATS) {
          CaaaRolloutStats s;
          s.result = r;
          s.turns = g->turns;
          s.decisions = decisions;
          s.draws = (int32_t)g->draws;
          s.final_player = g->st.player_index;
          s.state_hash = hash_live_state(g);
          a.stats[gi] = s;
        }
        n_play++;
        n_plies += (unsigned long long)g->turns;
        n_dec += (unsigned long long)decisions;
        n_draws += g->draws;
        n_flag += g->status != 0;
        sum += r;
        live = false;
      }
    }
  }
  if (a.summary != nullptr) {  // warp-reduce, then one atomic per warp and counter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      n_play += __shfl_down_sync(0xffffffffu, n_play, o);
      n_plies += __shfl_down_sync(0xffffffffu, n_plies, o);
      n_dec += __shfl_down_sync(0xffffffffu, n_dec, o);
      n_draws += __shfl_down_sync(0xffffffffu, n_draws, o);
      n_flag += __shfl_down_sync(0xffffffffu, n_flag, o);
      sum += __shfl_down_sync(0xffffffffu, sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&a.summary->playouts, n_play);
      atomicAdd((unsigned long long*)&a.summary->plies, n_plies);
      atomicAdd((unsigned long long*)&a.summary->decisions, n_dec);
      atomicAdd((unsigned long long*)&a.summary->draws, n_draws);
      atomicAdd((unsigned long long*)&a.summary->flagged, n_flag);
      atomicAdd(&a.summary->sum_result, sum);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Phase-regrouped persistent rollout kernel.
// ------------------------------------------------------------------------------------------------
constexpr int Q_SLOTS = 192;      // game slots per CTA
constexpr int Q_WARPS = 8;        // more warps than slots/32: batches overlap, queues drain faster
constexpr int Q_THREADS = Q_WARPS * 32;
// scheduler states of a slot: 0..14 = waiting for that pipeline phase, then:
constexpr uint32_t ST_EOT = CAAA_PH_BUY_UNITS;  // 15: purchase + end-of-turn bookkeeping + scoring (always runs)
constexpr uint32_t ST_START = 16;               // needs a new game from the ticket counter
constexpr uint32_t ST_BUSY = 30;                // claimed by a lane
constexpr uint32_t ST_FREE = 31;                // no more games
constexpr int Q_SMEM = TABLE_BYTES + Q_SLOTS * (int)sizeof(GameSlot) + Q_SLOTS * 4 + 32 * 4;

struct LaneTotals {
  unsigned long long playouts = 0, plies = 0, decisions = 0, draws = 0, flagged = 0;
  double sum = 0.0;
};

template <bool STATS>
__device__ __forceinline__ void finish_playout(GameSlot* g, double score, unsigned long long gi, const RolloutArgs& a, LaneTotals& tot) {
  const double r = (g->st.player_index % 2 == 1) ? 1.0 - score : score;  // reference src/engine.cpp:4238-4241
  a.results[gi] = r;
  const int decisions = 100000 - g->answers_remaining;
  if (STATS) {
    CaaaRolloutStats s;
    s.result = r;
    s.turns = g->turns;
    s.decisions = decisions;
    s.draws = (int32_t)g->draws;
    s.final_player = g->st.player_index;
    s.state_hash = hash_live_state(g);
    a.stats[gi] = s;
  }
  tot.playouts++;
  tot.plies += (unsigned long long)g->turns;
  tot.decisions += (unsigned long long)decisions;
  tot.draws += g->draws;
  tot.flagged += g->status != 0;
  tot.sum += r;
}

// Block-synchronous variant of the lockstep kernel: a barrier after every phase keeps all warps of the
// CTA inside the same phase function at the same time, so they share its instructions in the SM's
// instruction caches (the whole pipeline is ~240 KB of SASS; one phase is 5-30 KB).
// CTA sub-group barriers (hardware named barriers 1..GROUPS): the warps of one group stay in the same
// phase function; different groups may drift apart by whole phases, which shortens every barrier wait
// (max over fewer games) at the price of a few more phases live in the instruction cache.
__device__ __forceinline__ void group_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ bool group_or(int id, int count, bool pred) {
  int out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, %2, %3, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((int)pred), "r"(id), "r"(count)
      : "memory");
  return out != 0;
}
#define CAAA_SYNC_PHASE(call)    \
  do {                           \
    if (run) { call; }           \
    group_sync(bar_id, bar_cnt); \
  } while (0)
template <bool STATS, int LANES, int SLOTS = ROLLOUT_THREADS>
__global__ void __launch_bounds__(SLOTS / LANES * 32, ROLLOUT_THREADS / SLOTS) rollout_kernel_sync(const RolloutArgs a) {
  const int n_groups = a.groups;
  const int warps_per_group = (SLOTS / LANES) / n_groups;
  const int bar_id = 1 + (int)(threadIdx.x >> 5) / warps_per_group, bar_cnt = warps_per_group * 32;
  // Only the first LANES lanes of every warp own a game slot.  The kernel is bound by per-warp
  // latency and by the serialised union of divergent control flow, not by issue slots, so fewer
  // games per warp (a shorter union) and more warps per SM (more latency hiding) is the better trade.
  const DevTables* T = stage_tables(a.tables);
  const int lane = threadIdx.x & 31;
  const bool owner = lane < LANES;
  GameSlot* g = reinterpret_cast<GameSlot*>(smem_raw + TABLE_BYTES) + ((threadIdx.x >> 5) * LANES + (owner ? lane : 0));
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  bool live = false, exhausted = !owner;
  unsigned long long gi = 0;
  double score = 0.0;
  LaneTotals tot;
  for (;;) {
    if (!live && !exhausted) {
      gi = atomicAdd(a.ticket, 1ULL);
      if (gi < a.n_games) {
        load_state(g, a.states + (gi / a.rollouts_per_state) * CAAA_STATE_BYTES);
        begin_playout(g, cx, a.first_game_id + gi);
        score = evaluate_state(g, cx);
        live = true;
      } else {
        exhausted = true;
      }
    }
    if (!group_or(bar_id, bar_cnt, live)) break;
    const bool run = live && score > 0.01 && score < 0.99 && g->turns < 1000;
    if (live && !run) {
      finish_playout<STATS>(g, score, gi, a, tot);
      live = false;
    }
    CAAA_SYNC_PHASE(move_fighter_units<false>(g, cx));
    CAAA_SYNC_PHASE(move_bomber_units<false>(g, cx));
    CAAA_SYNC_PHASE(stage_transport_units<false>(g, cx));
    CAAA_SYNC_PHASE(move_land_unit_type<false>(g, cx, TANKS));
    CAAA_SYNC_PHASE(move_land_unit_type<false>(g, cx, ARTILLERY));
    CAAA_SYNC_PHASE(move_land_unit_type<false>(g, cx, INFANTRY));
    CAAA_SYNC_PHASE(move_transport_units<false>(g, cx));
    CAAA_SYNC_PHASE(move_subs<false>(g, cx));
    CAAA_SYNC_PHASE(move_destroyers_battleships<false>(g, cx));
    CAAA_SYNC_PHASE(resolve_sea_battles<false>(g, cx));
    CAAA_SYNC_PHASE(unload_transports<false>(g, cx));
    CAAA_SYNC_PHASE(resolve_land_battles<false>(g, cx));
    CAAA_SYNC_PHASE(move_land_unit_type<false>(g, cx, AA_GUNS));
    CAAA_SYNC_PHASE(land_fighter_units<false>(g, cx));
    CAAA_SYNC_PHASE(land_bomber_units<false>(g, cx));
    CAAA_SYNC_PHASE(buy_units<false>(g, cx));
    CAAA_SYNC_PHASE(crash_air_units(g, cx); reset_units_fully(g); collect_money(g, cx));
    CAAA_SYNC_PHASE(rotate_turns(g, cx));
    if (run) {
      g->turns++;
      score = evaluate_state(g, cx);
    }
  }
  if (a.summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot.playouts += __shfl_down_sync(0xffffffffu, tot.playouts, o);
      tot.plies += __shfl_down_sync(0xffffffffu, tot.plies, o);
      tot.decisions += __shfl_down_sync(0xffffffffu, tot.decisions, o);
      tot.draws += __shfl_down_sync(0xffffffffu, tot.draws, o);
      tot.flagged += __shfl_down_sync(0xffffffffu, tot.flagged, o);
      tot.sum += __shfl_down_sync(0xffffffffu, tot.sum, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&a.summary->playouts, tot.playouts);
      atomicAdd((unsigned long long*)&a.summary->plies, tot.plies);
      atomicAdd((unsigned long long*)&a.summary->decisions, tot.decisions);
      atomicAdd((unsigned long long*)&a.summary->draws, tot.draws);
      atomicAdd((unsigned long long*)&a.summary->flagged, tot.flagged);
      atomicAdd(&a.summary->sum_result, tot.sum);
    }
  }
}

// Run scheduler state `p` on slot g; returns the slot's next scheduler state.
template <bool STATS>
__device__ __forceinline__ uint32_t run_item(GameSlot* g, const RunCtx& cx, uint32_t p, const RolloutArgs& a, LaneTotals& tot) {
  switch (p) {
    case ST_START: {
      const unsigned long long gi = atomicAdd(a.ticket, 1ULL);
      if (gi >= a.n_games) return ST_FREE;
      load_state(g, a.states + (gi / a.rollouts_per_state) * CAAA_STATE_BYTES);
      begin_playout(g, cx, a.first_game_id + gi);
      const double score = evaluate_state(g, cx);
      if (!(score > 0.01 && score < 0.99)) {  // already terminal: zero turns
        finish_playout<STATS>(g, score, gi, a, tot);
        return ST_START;
      }
      return (uint32_t)next_phase_with_work(g, 0);
    }
    case CAAA_PH_MOVE_FIGHTERS: move_fighter_units<false>(g, cx); break;
    case CAAA_PH_MOVE_BOMBERS: move_bomber_units<false>(g, cx); break;
    case CAAA_PH_STAGE_TRANSPORTS: stage_transport_units<false>(g, cx); break;
    case CAAA_PH_MOVE_TANKS: move_land_unit_type<false>(g, cx, TANKS); break;
    case CAAA_PH_MOVE_ARTILLERY: move_land_unit_type<false>(g, cx, ARTILLERY); break;
    case CAAA_PH_MOVE_INFANTRY: move_land_unit_type<false>(g, cx, INFANTRY); break;
    case CAAA_PH_MOVE_TRANSPORTS: move_transport_units<false>(g, cx); break;
    case CAAA_PH_MOVE_SUBS: move_subs<false>(g, cx); break;
    case CAAA_PH_MOVE_SHIPS: move_destroyers_battleships<false>(g, cx); break;
    case CAAA_PH_SEA_BATTLES: resolve_sea_battles<false>(g, cx); break;
    case CAAA_PH_UNLOAD_TRANSPORTS: unload_transports<false>(g, cx); break;
    case CAAA_PH_LAND_BATTLES: resolve_land_battles<false>(g, cx); break;
    case CAAA_PH_MOVE_AA_GUNS: move_land_unit_type<false>(g, cx, AA_GUNS); break;
    case CAAA_PH_LAND_FIGHTERS: land_fighter_units<false>(g, cx); break;
    case CAAA_PH_LAND_BOMBERS: land_bomber_units<false>(g, cx); break;
    default: {  // ST_EOT: reference src/engine.cpp:4229-4236
      buy_units<false>(g, cx);
      crash_air_units(g, cx);
      reset_units_fully(g);
      collect_money(g, cx);
      rotate_turns(g, cx);
      g->turns++;
      const double score = evaluate_state(g, cx);
      if (score > 0.01 && score < 0.99 && g->turns < 1000) return (uint32_t)next_phase_with_work(g, 0);
      const unsigned long long gi = ((unsigned long long)g->game_hi << 32 | g->game_lo) - a.first_game_id;
      finish_playout<STATS>(g, score, gi, a, tot);
      return ST_START;
    }
  }
  return (uint32_t)next_phase_with_work(g, (int)p + 1);
}

template <bool STATS>
__global__ void __launch_bounds__(Q_THREADS, 1) rollout_kernel_regrouped(const RolloutArgs a) {
  const DevTables* T = stage_tables(a.tables);
  GameSlot* slots = reinterpret_cast<GameSlot*>(smem_raw + TABLE_BYTES);
  volatile uint32_t* sstate = reinterpret_cast<volatile uint32_t*>(smem_raw + TABLE_BYTES + Q_SLOTS * sizeof(GameSlot));
  volatile int* cnt = reinterpret_cast<volatile int*>(sstate + Q_SLOTS);  // pending slots per scheduler state
  for (int i = threadIdx.x; i < Q_SLOTS; i += blockDim.x) sstate[i] = ST_START;
  if (threadIdx.x < 32) cnt[threadIdx.x] = threadIdx.x == ST_START ? Q_SLOTS : 0;
  __syncthreads();
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  LaneTotals tot;
  for (;;) {
    // pick the fullest queue (ties: the later state, which is closer to finishing a turn)
    const int c = lane <= (int)ST_START ? cnt[lane] : 0;
    const unsigned key = __reduce_max_sync(0xffffffffu, c > 0 ? ((unsigned)c << 5) | (unsigned)lane : 0u);
    if (key == 0) {
      if (cnt[ST_FREE] >= Q_SLOTS) break;  // every slot retired
      __nanosleep(64);
      continue;
    }
    const uint32_t p = key & 31u;
    int my = -1;
#pragma unroll 1
    for (int k = 0; k < Q_SLOTS / 32; k++) {
      const int s = ((k + warp) % (Q_SLOTS / 32)) * 32 + lane;
      if (my < 0 && sstate[s] == p && atomicCAS(const_cast<uint32_t*>(&sstate[s]), p, ST_BUSY) == p) my = s;
    }
    const int n = __popc(__ballot_sync(0xffffffffu, my >= 0));
    if (lane == 0 && n > 0) atomicSub(const_cast<int*>(&cnt[p]), n);
    if (my >= 0) {
      __threadfence_block();
      const uint32_t q = run_item<STATS>(slots + my, cx, p, a, tot);
      __threadfence_block();
      sstate[my] = q;
      atomicAdd(const_cast<int*>(&cnt[q]), 1);
    }
    __syncwarp();
  }
  if (a.summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot.playouts += __shfl_down_sync(0xffffffffu, tot.playouts, o);
      tot.plies += __shfl_down_sync(0xffffffffu, tot.plies, o);
      tot.decisions += __shfl_down_sync(0xffffffffu, tot.decisions, o);
      tot.draws += __shfl_down_sync(0xffffffffu, tot.draws, o);
      tot.flagged += __shfl_down_sync(0xffffffffu, tot.flagged, o);
      tot.sum += __shfl_down_sync(0xffffffffu, tot.sum, o);
    }
    if (lane == 0) {
      atomicAdd((unsigned long long*)&a.summary->playouts, tot.playouts);
      atomicAdd((unsigned long long*)&a.summary->plies, tot.plies);
      atomicAdd((unsigned long long*)&a.summary->decisions, tot.decisions);
      atomicAdd((unsigned long long*)&a.summary->draws, tot.draws);
      atomicAdd((unsigned long long*)&a.summary->flagged, tot.flagged);
      atomicAdd(&a.summary->sum_result, tot.sum);
    }
  }
}

}  // namespace caaa
#include "stream_kernel.cuh"
namespace caaa {

// ---- cache dump shared by the advance kernel ----
__device__ void dump_cache(const GameSlot* g, const DevTables* T, CaaaCacheDump* d) {
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
  for (int i = 0; i < 30; i++) (&d->current_player_land_unit_types[0][0])[i] = (&g->cur_land[0][0])[i];
  for (int i = 0; i < 45; i++) (&d->current_player_sea_unit_types[0][0])[i] = (&g->cur_sea[0][0])[i];
  for (int i = 0; i < 25; i++) d->is_land_path_blocked[i] = g->land_blocked[i];
  for (int i = 0; i < 9; i++) {
    d->is_sea_path_blocked[i] = g->sea_blocked[i];
    d->is_sub_path_blocked[i] = g->sub_blocked[i];
  }
  for (int e = 0; e < 4; e++) d->enemies_0[e] = e < T->n_enemies[pi] ? T->enemies[pi][e] : 0;
  for (int p = 0; p < NP; p++) d->is_allied_0[p] = (T->allied_mask[pi] >> p) & 1;
  d->enemies_count_0 = T->n_enemies[pi];
  d->canal_state = g->canal_state;
  for (int i = 0; i < CAAA_ACTIONS; i++) d->valid_moves[i] = g->vm[i];
  d->valid_moves_count = g->nvm;
  d->pad_[0] = 0;
}

__global__ void __launch_bounds__(SMALL_THREADS) advance_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                unsigned long long seed, unsigned long long first_game_id,
                                                                int n_turns, uint8_t* out_states, CaaaCacheDump* out_caches,
                                                                int32_t* out_draws) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GameSlot* g = my_slot();
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  load_state(g, states + (size_t)i * CAAA_STATE_BYTES);
  begin_playout(g, cx, first_game_id + (unsigned long long)i);
  for (int t = 0; t < n_turns; t++) play_turn_rollout<false>(g, cx);
  if (out_states) store_state(g, out_states + (size_t)i * CAAA_STATE_BYTES);
  if (out_caches) dump_cache(g, T, out_caches + i);
  if (out_draws) out_draws[i] = (int32_t)g->draws;
}

// Isolated combat resolution (BASELINE config 3).
__global__ void __launch_bounds__(SMALL_THREADS) battle_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                               unsigned long long seed, unsigned long long first_game_id,
                                                               uint8_t* out_states, int32_t* out_draws, CaaaBatchSummary* summary) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long draws = 0, flagged = 0, one = 0;
  if (i < n) {
    GameSlot* g = my_slot();
    const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
    load_state(g, states + (size_t)i * CAAA_STATE_BYTES);
    begin_playout(g, cx, first_game_id + (unsigned long long)i);
    resolve_sea_battles<false>(g, cx);
    resolve_land_battles<false>(g, cx);
    if (out_states) store_state(g, out_states + (size_t)i * CAAA_STATE_BYTES);
    if (out_draws) out_draws[i] = (int32_t)g->draws;
    draws = g->draws;
    flagged = g->status != 0;
    one = 1;
  }
  if (summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      draws += __shfl_down_sync(0xffffffffu, draws, o);
      flagged += __shfl_down_sync(0xffffffffu, flagged, o);
      one += __shfl_down_sync(0xffffffffu, one, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&summary->playouts, one);
      atomicAdd((unsigned long long*)&summary->draws, draws);
      atomicAdd((unsigned long long*)&summary->flagged, flagged);
    }
  }
}

// Stop-at-decision mode (reference get_possible_actions / apply_action, src/engine.cpp:4050-4183).
__device__ __forceinline__ void run_until_decision(GameSlot* g, const RunCtx& cx) {
  for (;;) {
    if (move_fighter_units<true>(g, cx)) return;
    if (move_bomber_units<true>(g, cx)) return;
    if (stage_transport_units<true>(g, cx)) return;
    if (move_land_unit_type<true>(g, cx, TANKS)) return;
    if (move_land_unit_type<true>(g, cx, ARTILLERY)) return;
    if (move_land_unit_type<true>(g, cx, INFANTRY)) return;
    if (move_transport_units<true>(g, cx)) return;
    if (move_subs<true>(g, cx)) return;
    if (move_destroyers_battleships<true>(g, cx)) return;
    if (resolve_sea_battles<true>(g, cx)) return;
    if (unload_transports<true>(g, cx)) return;
    if (resolve_land_battles<true>(g, cx)) return;
    if (move_land_unit_type<true>(g, cx, AA_GUNS)) return;
    if (land_fighter_units<true>(g, cx)) return;
    if (land_bomber_units<true>(g, cx)) return;
    if (buy_units<true>(g, cx)) return;
    crash_air_units(g, cx);
    reset_units_fully(g);
    collect_money(g, cx);
    rotate_turns(g, cx);
  }
}
__device__ __forceinline__ void begin_expansion(GameSlot* g, const RunCtx& cx, int answers, int use_selected, uint8_t action) {
  refresh_full_cache(g, cx);
  g->unlucky = 0;
  g->answers_remaining = answers;
  g->use_selected = (uint8_t)use_selected;
  g->selected_action = action;
  g->draws = 0;
  g->game_lo = g->game_hi = 0;
  g->turns = 0;
  g->status = 0;
  for (int i = 0; i < CAAA_ACTIONS; i++) g->vm[i] = 0;
  g->nvm = 0;
}
__global__ void __launch_bounds__(SMALL_THREADS) actions_kernel(const DevTables* tables, const uint8_t* leaves, int n,
                                                                uint8_t* n_actions, uint8_t* actions, uint32_t* status) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GameSlot* g = my_slot();
  const RunCtx cx{T, 0u, 0u};
  load_state(g, leaves + (size_t)i * CAAA_STATE_BYTES);
  begin_expansion(g, cx, 0, 0, 0);
  run_until_decision(g, cx);
  const int k = g->nvm < CAAA_ACTIONS ? g->nvm : CAAA_ACTIONS;
  n_actions[i] = (uint8_t)k;
  for (int j = 0; j < CAAA_ACTIONS; j++) actions[(size_t)i * CAAA_ACTIONS + j] = j < k ? g->vm[j] : 0;
  if (status) status[i] = g->status ? 0x80000000u : 0u;  // bit 31: the action list itself tripped a guard
}
__global__ void __launch_bounds__(SMALL_THREADS) apply_kernel(const DevTables* tables, const uint8_t* leaves, int n,
                                                              const uint8_t* n_actions, const uint8_t* actions,
                                                              uint8_t* children, uint32_t* status) {
  const DevTables* T = stage_tables(tables);
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = idx / CAAA_ACTIONS, k = idx % CAAA_ACTIONS;
  if (i >= n || k >= n_actions[i]) return;
  GameSlot* g = my_slot();
  const RunCtx cx{T, 0u, 0u};
  load_state(g, leaves + (size_t)i * CAAA_STATE_BYTES);
  begin_expansion(g, cx, 1, 1, actions[(size_t)i * CAAA_ACTIONS + k]);
  run_until_decision(g, cx);
  store_state(g, children + (size_t)idx * CAAA_STATE_BYTES);
  if (status && g->status) atomicOr(&status[i], 1u << k);  // bit k: child k tripped a guard
}
__global__ void __launch_bounds__(SMALL_THREADS) evaluate_kernel(const DevTables* tables, const uint8_t* states, int n, double* scores) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  GameSlot* g = my_slot();
  const RunCtx cx{T, 0u, 0u};
  load_state(g, states + (size_t)i * CAAA_STATE_BYTES);
  refresh_full_cache(g, cx);
  scores[i] = evaluate_state(g, cx);
}

// ------------------------------------------------------------------------------------------------
// host side of the C-ABI
// ------------------------------------------------------------------------------------------------
struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  DevTables host_tables;
  DevTables* d_tables = nullptr;
  unsigned long long* d_ticket = nullptr;
  CaaaBatchSummary* d_summary = nullptr;
  // growable scratch for the host-buffer entry points
  void* d_in = nullptr;
  size_t in_cap = 0;
  void* d_out = nullptr;
  size_t out_cap = 0;
  void* d_aux = nullptr;
  size_t aux_cap = 0;
  void* d_aux2 = nullptr;
  size_t aux2_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long seed = 0, next_game_id = 0;
  bool lockstep = false, sync = true, regrouped = false, streaming = false;
  int lanes = 8, ctas_per_sm = 1, groups = 1;
  // device-wide phase queues + game pool of the streaming kernel (allocated on first use)
  PhaseQueues* d_queues = nullptr;
  unsigned long long* d_cells = nullptr;
  PoolSlot* d_pool = nullptr;
  unsigned pool_size = 0;
  int pool_mult = 2;
};
static Context g_ctx;
static std::string g_err;

static int fail(int code, const char* what, cudaError_t e = cudaSuccess) {
  g_err = what;
  if (e != cudaSuccess) {
    g_err += ": ";
    g_err += cudaGetErrorString(e);
  }
  return code;
}
#define CUDA_TRY(expr)                                             \
  do {                                                             \
    cudaError_t e_ = (expr);                                       \
    if (e_ != cudaSuccess) return fail(CAAA_E_CUDA, #expr, e_);    \
  } while (0)

static int grow(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return CAAA_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CUDA_TRY(cudaMalloc(p, need));
  *cap = need;
  return CAAA_OK;
}
static int small_smem() { return TABLE_BYTES + SMALL_THREADS * (int)sizeof(GameSlot); }
static int rollout_smem() { return TABLE_BYTES + ROLLOUT_THREADS * (int)sizeof(GameSlot); }

static int launch_rollout(const void* d_states, unsigned long long n_games, unsigned rollouts_per_state,
                          unsigned long long seed, unsigned long long first_game_id, double* d_results,
                          CaaaRolloutStats* d_stats, CaaaBatchSummary* d_summary, cudaStream_t stream) {
  Context& c = g_ctx;
  CUDA_TRY(cudaMemsetAsync(c.d_ticket, 0, sizeof(unsigned long long), stream));
  RolloutArgs a;
  a.tables = c.d_tables;
  a.states = (const uint8_t*)d_states;
  a.n_games = n_games;
  a.rollouts_per_state = rollouts_per_state;
  a.seed = seed;
  a.first_game_id = first_game_id;
  a.results = d_results;
  a.stats = d_stats;
  a.summary = d_summary;
  a.ticket = c.d_ticket;
  a.debug = std::getenv("CAAA_DEBUG") != nullptr;
  a.groups = c.groups;
  if (c.streaming) {
    const unsigned long long want = (n_games + STREAM_SLOTS - 1) / STREAM_SLOTS;
    const int grid = (int)(want < (unsigned long long)c.sm_count ? want : (unsigned long long)c.sm_count);
    const unsigned need = (unsigned)c.sm_count * STREAM_SLOTS * (unsigned)c.pool_mult;
    if (c.d_pool == nullptr || c.pool_size < need) {
      if (c.d_pool) { cudaFree(c.d_pool); cudaFree(c.d_cells); cudaFree(c.d_queues); c.d_pool = nullptr; }
      unsigned cap = 1;
      while (cap < need) cap <<= 1;
      CUDA_TRY(cudaMalloc(&c.d_pool, (size_t)need * sizeof(PoolSlot)));
      CUDA_TRY(cudaMalloc(&c.d_cells, (size_t)NQ * cap * sizeof(unsigned long long)));
      CUDA_TRY(cudaMalloc(&c.d_queues, sizeof(PhaseQueues)));
      CUDA_TRY(cudaMemset(c.d_cells, 0, (size_t)NQ * cap * sizeof(unsigned long long)));
      PhaseQueues h;
      std::memset(&h, 0, sizeof(h));
      h.capacity = cap;
      h.pool_size = need;
      h.cells = c.d_cells;
      h.pool = c.d_pool;
      CUDA_TRY(cudaMemcpy(c.d_queues, &h, sizeof(h), cudaMemcpyHostToDevice));
      c.pool_size = need;
    }
    // queue positions keep counting across launches (cells are tagged with them); only the per-launch counters reset
    queues_reset_kernel<<<1, 32, 0, stream>>>(c.d_queues);
    // a launch with fewer CTAs than the pool was sized for simply uses fewer pool entries
    if (d_stats)
      rollout_kernel_stream<true><<<grid, STREAM_THREADS + 32, rollout_smem(), stream>>>(a, c.d_queues);
    else
      rollout_kernel_stream<false><<<grid, STREAM_THREADS + 32, rollout_smem(), stream>>>(a, c.d_queues);
    if (a.debug) {
      PhaseQueues h;
      cudaStreamSynchronize(stream);
      cudaMemcpy(&h, c.d_queues, sizeof(h), cudaMemcpyDeviceToHost);
      std::fprintf(stderr, "stream dbg: items %llu empty_polls %llu ticket_wait_polls %llu switches %llu busy_frac %.3f aborted %d\n",
                   h.dbg[0], h.dbg[1], h.dbg[2], h.dbg[3], h.dbg[5] ? (double)h.dbg[4] / (double)h.dbg[5] : 0.0, h.aborted);
    }
  } else if (c.sync) {
    const unsigned long long want = (n_games + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;
    const int grid = (int)(want < (unsigned long long)c.sm_count ? want : (unsigned long long)c.sm_count);
#define CAAA_LAUNCH_SYNC(L)                                                                             \
  do {                                                                                                 \
    if (d_stats)                                                                                       \
      rollout_kernel_sync<true, L><<<grid, ROLLOUT_THREADS / L * 32, rollout_smem(), stream>>>(a);     \
    else                                                                                               \
      rollout_kernel_sync<false, L><<<grid, ROLLOUT_THREADS / L * 32, rollout_smem(), stream>>>(a);    \
  } while (0)
    if (c.ctas_per_sm > 1) {  // several smaller CTAs per SM: one CTA's barrier waits overlap another's work
      const int slots = c.ctas_per_sm == 2 ? 96 : 64;
      const unsigned long long want2 = (n_games + slots - 1) / slots;
      const unsigned long long cap = (unsigned long long)c.sm_count * c.ctas_per_sm;
      const int grid2 = (int)(want2 < cap ? want2 : cap);
      const int smem2 = TABLE_BYTES + slots * (int)sizeof(GameSlot);
      if (c.ctas_per_sm == 2) {
        if (d_stats) rollout_kernel_sync<true, 8, 96><<<grid2, 96 / 8 * 32, smem2, stream>>>(a);
        else rollout_kernel_sync<false, 8, 96><<<grid2, 96 / 8 * 32, smem2, stream>>>(a);
      } else {
        if (d_stats) rollout_kernel_sync<true, 8, 64><<<grid2, 64 / 8 * 32, smem2, stream>>>(a);
        else rollout_kernel_sync<false, 8, 64><<<grid2, 64 / 8 * 32, smem2, stream>>>(a);
      }
    } else if (c.lanes == 32) CAAA_LAUNCH_SYNC(32);
    else if (c.lanes == 16) CAAA_LAUNCH_SYNC(16);
    else if (c.lanes == 6) CAAA_LAUNCH_SYNC(6);
    else CAAA_LAUNCH_SYNC(8);
  } else if (c.lockstep) {  // first-generation kernel, kept for A/B measurements (CAAA_ROLLOUT_KERNEL=lockstep)
    const unsigned long long want = (n_games + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;
    const int grid = (int)(want < (unsigned long long)c.sm_count ? want : (unsigned long long)c.sm_count);
    if (d_stats)
      rollout_kernel<true><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
    else
      rollout_kernel<false><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
  } else {
    const unsigned long long want = (n_games + Q_SLOTS - 1) / Q_SLOTS;
    const int grid = (int)(want < (unsigned long long)c.sm_count ? want : (unsigned long long)c.sm_count);
    if (d_stats)
      rollout_kernel_regrouped<true><<<grid, Q_THREADS, Q_SMEM, stream>>>(a);
    else
      rollout_kernel_regrouped<false><<<grid, Q_THREADS, Q_SMEM, stream>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return CAAA_OK;
}

}  // namespace caaa

using namespace caaa;

extern "C" {

const char* caaa_gpu_last_error(void) { return g_err.c_str(); }

int caaa_gpu_init(int device) {
  Context& c = g_ctx;
  if (c.ready) return CAAA_OK;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(CAAA_E_NO_DEVICE, "no CUDA device available (this library has no CPU path)", e);
  if (device < 0 || device >= count) return fail(CAAA_E_ARG, "device index out of range");
  CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if ((int)prop.sharedMemPerBlockOptin < rollout_smem())
    return fail(CAAA_E_NO_DEVICE, "device does not offer 227 KB of shared memory per block (built for sm_100a)");
  c.device = device;
  c.sm_count = prop.multiProcessorCount;
  build_dev_tables(&c.host_tables);
  CUDA_TRY(cudaMalloc(&c.d_tables, sizeof(DevTables)));
  CUDA_TRY(cudaMemcpy(c.d_tables, &c.host_tables, sizeof(DevTables), cudaMemcpyHostToDevice));
  CUDA_TRY(cudaMalloc(&c.d_ticket, sizeof(unsigned long long)));
  CUDA_TRY(cudaMalloc(&c.d_summary, sizeof(CaaaBatchSummary)));
  CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c.ev0));
  CUDA_TRY(cudaEventCreate(&c.ev1));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_regrouped<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_regrouped<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Q_SMEM));
  {
    const char* k = std::getenv("CAAA_ROLLOUT_KERNEL");
    c.lockstep = k != nullptr && std::strcmp(k, "lockstep") == 0;
    c.regrouped = k != nullptr && std::strcmp(k, "regrouped") == 0;
    c.streaming = k != nullptr && std::strcmp(k, "stream") == 0;
    c.sync = !c.lockstep && !c.regrouped && !c.streaming;  // default: block-synchronous phases
    const char* pm = std::getenv("CAAA_POOL_MULT");
    c.pool_mult = pm ? std::atoi(pm) : 2;
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    const char* l = std::getenv("CAAA_LANES");
    c.lanes = l ? std::atoi(l) : 8;
    const char* gp = std::getenv("CAAA_GROUPS");
    c.groups = gp ? std::atoi(gp) : 1;
    const char* cp = std::getenv("CAAA_CTAS_PER_SM");
    c.ctas_per_sm = cp ? std::atoi(cp) : 1;
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 8, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_BYTES + 96 * (int)sizeof(GameSlot)));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 8, 96>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_BYTES + 96 * (int)sizeof(GameSlot)));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_BYTES + 64 * (int)sizeof(GameSlot)));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 8, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, TABLE_BYTES + 64 * (int)sizeof(GameSlot)));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<true, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
    CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_sync<false, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  }
  CUDA_TRY(cudaFuncSetAttribute(advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(battle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(actions_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(evaluate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  c.ready = true;
  return CAAA_OK;
}

int caaa_gpu_shutdown(void) {
  Context& c = g_ctx;
  if (!c.ready) return CAAA_OK;
  cudaStreamSynchronize(c.stream);
  cudaFree(c.d_tables);
  cudaFree(c.d_ticket);
  cudaFree(c.d_summary);
  if (c.d_in) cudaFree(c.d_in);
  if (c.d_out) cudaFree(c.d_out);
  if (c.d_aux) cudaFree(c.d_aux);
  if (c.d_aux2) cudaFree(c.d_aux2);
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  cudaStreamDestroy(c.stream);
  c = Context();
  return CAAA_OK;
}

int caaa_gpu_get_tables(CaaaTables* out) {
  if (!out) return fail(CAAA_E_ARG, "null argument");
  build_tables(out);  // host computation; identical to what caaa_gpu_init uploads
  return CAAA_OK;
}

int caaa_gpu_device_info(int* sm_count, int* max_smem_per_block, char* name64) {
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  cudaDeviceProp prop;
  CUDA_TRY(cudaGetDeviceProperties(&prop, g_ctx.device));
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (max_smem_per_block) *max_smem_per_block = (int)prop.sharedMemPerBlockOptin;
  if (name64) {
    std::strncpy(name64, prop.name, 63);
    name64[63] = 0;
  }
  return CAAA_OK;
}

int caaa_gpu_rollout_batch_device(const void* d_states, int n_states, int rollouts_per_state, uint64_t seed,
                                  uint64_t first_game_id, double* d_results, CaaaRolloutStats* d_stats,
                                  CaaaBatchSummary* d_summary, void* cuda_stream) {
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!d_states || !d_results || n_states <= 0 || rollouts_per_state <= 0) return fail(CAAA_E_ARG, "bad rollout batch arguments");
  return launch_rollout(d_states, (unsigned long long)n_states * (unsigned long long)rollouts_per_state,
                        (unsigned)rollouts_per_state, seed, first_game_id, d_results, d_stats, d_summary,
                        (cudaStream_t)cuda_stream);
}

int caaa_gpu_rollout_batch(const CaaaGameState* states, int n_states, int rollouts_per_state, uint64_t seed,
                           uint64_t first_game_id, double* results, CaaaRolloutStats* stats, CaaaBatchSummary* summary) {
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!states || !results || n_states <= 0 || rollouts_per_state <= 0) return fail(CAAA_E_ARG, "bad rollout batch arguments");
  const size_t n_games = (size_t)n_states * (size_t)rollouts_per_state;
  const size_t in_bytes = (size_t)n_states * CAAA_STATE_BYTES;
  int rc;
  if ((rc = grow(&c.d_in, &c.in_cap, in_bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_out, &c.out_cap, n_games * sizeof(double))) != CAAA_OK) return rc;
  if (stats && (rc = grow(&c.d_aux, &c.aux_cap, n_games * sizeof(CaaaRolloutStats))) != CAAA_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_in, states, in_bytes, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(cudaMemsetAsync(c.d_summary, 0, sizeof(CaaaBatchSummary), c.stream));
  CUDA_TRY(cudaEventRecord(c.ev0, c.stream));
  rc = launch_rollout(c.d_in, n_games, (unsigned)rollouts_per_state, seed, first_game_id, (double*)c.d_out,
                      stats ? (CaaaRolloutStats*)c.d_aux : nullptr, c.d_summary, c.stream);
  if (rc != CAAA_OK) return rc;
  CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
  CUDA_TRY(cudaMemcpyAsync(results, c.d_out, n_games * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (stats) CUDA_TRY(cudaMemcpyAsync(stats, c.d_aux, n_games * sizeof(CaaaRolloutStats), cudaMemcpyDeviceToHost, c.stream));
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpyAsync(&s, c.d_summary, sizeof(s), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  if (summary) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    s.kernel_ms = ms;
    *summary = s;
  }
  return CAAA_OK;
}

int caaa_gpu_set_stream(uint64_t seed, uint64_t next_game_id) {
  g_ctx.seed = seed;
  g_ctx.next_game_id = next_game_id;
  return CAAA_OK;
}

double caaa_gpu_random_play_until_terminal(const CaaaGameState* state) {
  double r = NAN;
  if (caaa_gpu_rollout_batch(state, 1, 1, g_ctx.seed, g_ctx.next_game_id, &r, nullptr, nullptr) != CAAA_OK) return NAN;
  g_ctx.next_game_id++;
  return r;
}

int caaa_gpu_advance(const CaaaGameState* states, int n, uint64_t seed, uint64_t first_game_id, int n_turns,
                     CaaaGameState* out_states, CaaaCacheDump* out_caches, int32_t* out_draws) {
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!states || n <= 0 || n_turns < 0) return fail(CAAA_E_ARG, "bad advance arguments");
  const size_t bytes = (size_t)n * CAAA_STATE_BYTES;
  int rc;
  if ((rc = grow(&c.d_in, &c.in_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_out, &c.out_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_aux, &c.aux_cap, (size_t)n * sizeof(CaaaCacheDump))) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_aux2, &c.aux2_cap, (size_t)n * sizeof(int32_t))) != CAAA_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_in, states, bytes, cudaMemcpyHostToDevice, c.stream));
  const int grid = (n + SMALL_THREADS - 1) / SMALL_THREADS;
  advance_kernel<<<grid, SMALL_THREADS, small_smem(), c.stream>>>(c.d_tables, (const uint8_t*)c.d_in, n, seed, first_game_id, n_turns,
                                                                  (uint8_t*)c.d_out, (CaaaCacheDump*)c.d_aux, (int32_t*)c.d_aux2);
  CUDA_TRY(cudaGetLastError());
  if (out_states) CUDA_TRY(cudaMemcpyAsync(out_states, c.d_out, bytes, cudaMemcpyDeviceToHost, c.stream));
  if (out_caches) CUDA_TRY(cudaMemcpyAsync(out_caches, c.d_aux, (size_t)n * sizeof(CaaaCacheDump), cudaMemcpyDeviceToHost, c.stream));
  if (out_draws) CUDA_TRY(cudaMemcpyAsync(out_draws, c.d_aux2, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return CAAA_OK;
}

int caaa_gpu_resolve_battles(const CaaaGameState* states, int n, uint64_t seed, uint64_t first_game_id,
                             CaaaGameState* out_states, int32_t* out_draws, CaaaBatchSummary* summary) {
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!states || n <= 0) return fail(CAAA_E_ARG, "bad battle arguments");
  const size_t bytes = (size_t)n * CAAA_STATE_BYTES;
  int rc;
  if ((rc = grow(&c.d_in, &c.in_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_out, &c.out_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_aux2, &c.aux2_cap, (size_t)n * sizeof(int32_t))) != CAAA_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_in, states, bytes, cudaMemcpyHostToDevice, c.stream));
  CUDA_TRY(cudaMemsetAsync(c.d_summary, 0, sizeof(CaaaBatchSummary), c.stream));
  CUDA_TRY(cudaEventRecord(c.ev0, c.stream));
  const int grid = (n + SMALL_THREADS - 1) / SMALL_THREADS;
  battle_kernel<<<grid, SMALL_THREADS, small_smem(), c.stream>>>(c.d_tables, (const uint8_t*)c.d_in, n, seed, first_game_id,
                                                                 (uint8_t*)c.d_out, (int32_t*)c.d_aux2, c.d_summary);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
  if (out_states) CUDA_TRY(cudaMemcpyAsync(out_states, c.d_out, bytes, cudaMemcpyDeviceToHost, c.stream));
  if (out_draws) CUDA_TRY(cudaMemcpyAsync(out_draws, c.d_aux2, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpyAsync(&s, c.d_summary, sizeof(s), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  if (summary) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    s.kernel_ms = ms;
    *summary = s;
  }
  return CAAA_OK;
}

int caaa_gpu_expand(const CaaaGameState* leaves, int n, uint8_t* n_actions, uint8_t* actions, CaaaGameState* children,
                    uint32_t* status) {
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!leaves || !n_actions || !actions || n <= 0) return fail(CAAA_E_ARG, "bad expand arguments");
  const size_t bytes = (size_t)n * CAAA_STATE_BYTES;
  int rc;
  if ((rc = grow(&c.d_in, &c.in_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_out, &c.out_cap, bytes * CAAA_ACTIONS)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_aux, &c.aux_cap, (size_t)n * (CAAA_ACTIONS + 1))) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_aux2, &c.aux2_cap, (size_t)n * sizeof(uint32_t))) != CAAA_OK) return rc;
  uint8_t* d_nact = (uint8_t*)c.d_aux;
  uint8_t* d_act = d_nact + n;
  CUDA_TRY(cudaMemcpyAsync(c.d_in, leaves, bytes, cudaMemcpyHostToDevice, c.stream));
  actions_kernel<<<(n + SMALL_THREADS - 1) / SMALL_THREADS, SMALL_THREADS, small_smem(), c.stream>>>(
      c.d_tables, (const uint8_t*)c.d_in, n, d_nact, d_act, (uint32_t*)c.d_aux2);
  CUDA_TRY(cudaGetLastError());
  if (children) {
    const long long total = (long long)n * CAAA_ACTIONS;
    apply_kernel<<<(int)((total + SMALL_THREADS - 1) / SMALL_THREADS), SMALL_THREADS, small_smem(), c.stream>>>(
        c.d_tables, (const uint8_t*)c.d_in, n, d_nact, d_act, (uint8_t*)c.d_out, (uint32_t*)c.d_aux2);
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMemcpyAsync(children, c.d_out, bytes * CAAA_ACTIONS, cudaMemcpyDeviceToHost, c.stream));
  }
  CUDA_TRY(cudaMemcpyAsync(n_actions, d_nact, (size_t)n, cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaMemcpyAsync(actions, d_act, (size_t)n * CAAA_ACTIONS, cudaMemcpyDeviceToHost, c.stream));
  if (status) CUDA_TRY(cudaMemcpyAsync(status, c.d_aux2, (size_t)n * sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return CAAA_OK;
}

int caaa_gpu_evaluate(const CaaaGameState* states, int n, double* scores) {
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!states || !scores || n <= 0) return fail(CAAA_E_ARG, "bad evaluate arguments");
  const size_t bytes = (size_t)n * CAAA_STATE_BYTES;
  int rc;
  if ((rc = grow(&c.d_in, &c.in_cap, bytes)) != CAAA_OK) return rc;
  if ((rc = grow(&c.d_out, &c.out_cap, (size_t)n * sizeof(double))) != CAAA_OK) return rc;
  CUDA_TRY(cudaMemcpyAsync(c.d_in, states, bytes, cudaMemcpyHostToDevice, c.stream));
  evaluate_kernel<<<(n + SMALL_THREADS - 1) / SMALL_THREADS, SMALL_THREADS, small_smem(), c.stream>>>(
      c.d_tables, (const uint8_t*)c.d_in, n, (double*)c.d_out);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemcpyAsync(scores, c.d_out, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  return CAAA_OK;
}

}  // extern "C"
