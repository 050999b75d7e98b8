// stream_kernel.cuh -- device-wide phase queues + persistent "phase server" CTAs.
//
// Included by caaa_gpu.cu (after RolloutArgs / LaneTotals / finish_playout / load_state).
//
// Why: the turn pipeline is ~240 KB of SASS and nearly scalar per game (1-3 active lanes in the
// decision loops), so the two things that decide its speed are whether the instructions a warp needs
// are already in the SM's instruction cache, and whether a warp ever waits for another game.  Here
//   * all games of a launch live in a pool in global memory (L2-resident: 1184 B per game);
//   * there is one device-wide FIFO of pool indices per pipeline phase (phases 0..14, and 15 = the
//     end-of-turn item: purchase, bookkeeping, rotation, scoring, finish + restart);
//   * every CTA serves ONE phase at a time: its lanes pop a game waiting for that phase, copy its slot
//     into the lane's private shared-memory slot, run the phase function, compute the next phase that
//     has work for the game, copy the slot back and push the index onto that phase's queue.  All warps
//     of the CTA therefore execute the same few KB of code, no lane waits for another game, and a game
//     is only ever queued for phases in which it has something to do;
//   * when its queue runs dry the CTA picks the queue with the most waiting games per serving CTA.
// Finished games are replaced from the ticket counter inside the end-of-turn item, so the pool stays
// full until the batch runs out.
#pragma once

namespace caaa {

constexpr int NQ = 16;                 // scheduler queues: pipeline phases 0..14 and the end-of-turn item (15)
constexpr int STREAM_SLOTS = 192;      // in-flight games per CTA (one private shared-memory slot each)
constexpr int STREAM_LANES = 8;        // game-owning lanes per warp
constexpr int STREAM_THREADS = STREAM_SLOTS / STREAM_LANES * 32;
constexpr int POOL_WORDS = 296;        // 1184 B per pooled game: GameSlot (1180 B) rounded up to 16 B
constexpr uint32_t Q_RETIRED = 0xffffffffu;

struct alignas(16) PoolSlot {
  uint32_t w[POOL_WORDS];
};

struct PhaseQueues {
  unsigned long long head[NQ];
  unsigned long long tail[NQ];
  int servers[NQ];                 // CTAs currently serving each queue
  unsigned long long finished;     // playouts finished so far
  int aborted;                     // set by a scheduler warp whose watchdog saw no progress
  unsigned long long dbg[8];       // diagnostics: items, empty polls, ticket-wait polls, phase switches, busy cycles, total cycles
  unsigned int capacity;           // cells per queue, power of two, >= pool_size
  unsigned int pool_size;
  unsigned long long* cells;       // NQ x capacity: (position + 1) << 32 | pool index
  PoolSlot* pool;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

__device__ __forceinline__ void pool_load(GameSlot* g, const PoolSlot* p) {
  const uint4* s = reinterpret_cast<const uint4*>(p);
  uint32_t* d = reinterpret_cast<uint32_t*>(g);
#pragma unroll 8
  for (int i = 0; i < 73; i++) {
    const uint4 v = __ldcg(s + i);  // written by another SM: read through L2
    d[4 * i] = v.x;
    d[4 * i + 1] = v.y;
    d[4 * i + 2] = v.z;
    d[4 * i + 3] = v.w;
  }
  const uint4 v = __ldcg(s + 73);
  d[292] = v.x;
  d[293] = v.y;
  d[294] = v.z;
}
__device__ __forceinline__ void pool_store(const GameSlot* g, PoolSlot* p) {
  uint4* d = reinterpret_cast<uint4*>(p);
  const uint32_t* s = reinterpret_cast<const uint32_t*>(g);
#pragma unroll 8
  for (int i = 0; i < 73; i++) __stcg(d + i, make_uint4(s[4 * i], s[4 * i + 1], s[4 * i + 2], s[4 * i + 3]));
  __stcg(d + 73, make_uint4(s[292], s[293], s[294], 0u));
}

__device__ __forceinline__ void queue_push(PhaseQueues* Q, uint32_t q, uint32_t idx) {
  const unsigned long long pos = atomicAdd(&Q->tail[q], 1ULL);
  volatile unsigned long long* cell = Q->cells + (size_t)q * Q->capacity + (pos & (Q->capacity - 1));
  *cell = ((pos + 1) << 32) | idx;
}
// Start games in pool entry until one of them needs a turn; returns its first phase, or Q_RETIRED when
// the batch has no games left for this entry.
template <bool STATS>
__device__ __forceinline__ uint32_t start_game(GameSlot* g, const RunCtx& cx, const RolloutArgs& a, PhaseQueues* Q, LaneTotals& tot) {
  for (;;) {
    const unsigned long long gi = atomicAdd(a.ticket, 1ULL);
    if (gi >= a.n_games) return Q_RETIRED;
    load_state(g, a.states + (gi / a.rollouts_per_state) * CAAA_STATE_BYTES);
    begin_playout(g, cx, a.first_game_id + gi);
    const double score = evaluate_state(g, cx);
    if (score > 0.01 && score < 0.99) return (uint32_t)next_phase_with_work(g, 0);
    finish_playout<STATS>(g, score, gi, a, tot);  // already terminal: zero turns
    atomicAdd(&Q->finished, 1ULL);
  }
}

template <bool STATS>
__device__ __forceinline__ uint32_t run_stream_item(GameSlot* g, const RunCtx& cx, uint32_t p, const RolloutArgs& a, PhaseQueues* Q,
                                                    LaneTotals& tot) {
  switch (p) {
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
    default: {  // end-of-turn item, reference src/engine.cpp:4229-4241
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
      atomicAdd(&Q->finished, 1ULL);
      return start_game<STATS>(g, cx, a, Q, tot);
    }
  }
  return (uint32_t)next_phase_with_work(g, (int)p + 1);
}

// Resets the per-launch scheduler state.  Queue positions keep counting across launches (cells are
// tagged with them); positions handed out but never produced at the end of the previous launch are
// skipped by moving both ends of every queue to the larger of the two.
__global__ void queues_reset_kernel(PhaseQueues* Q) {
  const int q = threadIdx.x;
  if (q < NQ) {
    const unsigned long long m = Q->head[q] > Q->tail[q] ? Q->head[q] : Q->tail[q];
    Q->head[q] = m;
    Q->tail[q] = m;
    Q->servers[q] = 0;
  }
  if (q == 0) {
    Q->finished = 0;
    Q->aborted = 0;
  }
  if (q < 8) Q->dbg[q] = 0;
}

template <bool STATS>
__global__ void __launch_bounds__(STREAM_THREADS + 32, 1) rollout_kernel_stream(const RolloutArgs a, PhaseQueues* Q) {
  const DevTables* T = stage_tables(a.tables);
  __shared__ int s_phase_mem;
  volatile int* s_phase = &s_phase_mem;
  if (threadIdx.x == 0) *s_phase = -1;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (warp == STREAM_SLOTS / STREAM_LANES) {
    // ---- scheduler warp: decides which phase this CTA serves; never touches a game ----
    if (lane != 0) return;
    int cur = -1;
    unsigned long long last_sum = ~0ULL;
    long long last_change = clock64();
    for (;;) {
      if (ld_volatile_u64(&Q->finished) >= a.n_games || *reinterpret_cast<volatile int*>(&Q->aborted)) break;
      long long len[NQ];
      unsigned long long sum = 0;
      for (int q = 0; q < NQ; q++) {
        const unsigned long long h = ld_volatile_u64(&Q->head[q]), tl = ld_volatile_u64(&Q->tail[q]);
        len[q] = (long long)(tl - h);
        sum += h + tl;
      }
      int best = -1;
      double best_score = 0.0, cur_score = 0.0;
      for (int q = 0; q < NQ; q++) {
        if (len[q] <= 0) continue;
        const int srv = *reinterpret_cast<volatile int*>(&Q->servers[q]) - (q == cur ? 1 : 0);
        const double sc = (double)len[q] / (double)(srv + 1);
        if (q == cur) cur_score = sc;
        if (sc > best_score) {
          best_score = sc;
          best = q;
        }
      }
      // stay while the current queue still feeds the CTA; move when it is dry or another queue is far more loaded
      if (best >= 0 && best != cur && (cur < 0 || len[cur] <= 0 || best_score > 4.0 * cur_score + (double)STREAM_SLOTS)) {
        if (cur >= 0) atomicSub(&Q->servers[cur], 1);
        atomicAdd(&Q->servers[best], 1);
        cur = best;
        *s_phase = cur;
        if (a.debug) atomicAdd(&Q->dbg[3], 1ULL);
      }
      // watchdog: no queue moved for ~2 s although the batch is unfinished -> give up instead of hanging the GPU
      if (sum != last_sum) {
        last_sum = sum;
        last_change = clock64();
      } else if (clock64() - last_change > 4000000000LL) {
        atomicExch(&Q->aborted, 1);
        break;
      }
      __nanosleep(1000);
    }
    if (cur >= 0) atomicSub(&Q->servers[cur], 1);
    *s_phase = -2;
    return;
  }

  if (lane >= STREAM_LANES) return;  // only the first STREAM_LANES lanes of a worker warp own a slot
  const int slot = warp * STREAM_LANES + lane;
  GameSlot* g = reinterpret_cast<GameSlot*>(smem_raw + TABLE_BYTES) + slot;
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  LaneTotals tot;

  // 1. fill this CTA's share of the pool with fresh games
  for (unsigned e = blockIdx.x * STREAM_SLOTS + slot; e < Q->pool_size; e += gridDim.x * STREAM_SLOTS) {
    const uint32_t q = start_game<STATS>(g, cx, a, Q, tot);
    if (q == Q_RETIRED) break;  // tickets are exhausted: later entries stay empty as well
    pool_store(g, Q->pool + e);
    __threadfence();
    queue_push(Q, q, e);
  }

  // 2. worker loop.  A lane takes a position in its CTA's current phase queue with a fetch-add (no
  // compare-and-swap retries under contention) and keeps that ticket until the item shows up.
  bool have = false;
  uint32_t tq = 0;
  unsigned long long tpos = 0;
  unsigned long long d_items = 0, d_empty = 0, d_wait = 0, d_busy = 0;
  const long long d_t0 = clock64();
  for (;;) {
    if (!have) {
      const int p = *s_phase;
      if (p == -2) break;
      if (p < 0) {
        __nanosleep(200);
        continue;
      }
      if ((long long)(ld_volatile_u64(&Q->tail[p]) - ld_volatile_u64(&Q->head[p])) <= 0) {
        d_empty++;
        __nanosleep(300);
        continue;
      }
      tpos = atomicAdd(&Q->head[p], 1ULL);
      tq = (uint32_t)p;
      have = true;
    }
    const volatile unsigned long long* cell = Q->cells + (size_t)tq * Q->capacity + (tpos & (Q->capacity - 1));
    const unsigned long long v = *cell;
    if ((v >> 32) != ((tpos + 1) & 0xffffffffULL)) {  // position taken ahead of its producer
      if (*s_phase == -2) break;
      d_wait++;
      __nanosleep(100);
      continue;
    }
    const uint32_t e = (uint32_t)v;
    const long long d_b0 = clock64();
    __threadfence();
    pool_load(g, Q->pool + e);
    const uint32_t q = run_stream_item<STATS>(g, cx, tq, a, Q, tot);
    have = false;
    d_items++;
    if (q != Q_RETIRED) {
      pool_store(g, Q->pool + e);
      __threadfence();
      queue_push(Q, q, e);
    }
    d_busy += (unsigned long long)(clock64() - d_b0);
  }
  if (a.debug) {
    atomicAdd(&Q->dbg[0], d_items);
    atomicAdd(&Q->dbg[1], d_empty);
    atomicAdd(&Q->dbg[2], d_wait);
    atomicAdd(&Q->dbg[4], d_busy);
    atomicAdd(&Q->dbg[5], (unsigned long long)(clock64() - d_t0));
  }

  if (a.summary != nullptr) {  // the STREAM_LANES owner lanes of a warp reduce among themselves
    const unsigned mask = (1u << STREAM_LANES) - 1u;
#pragma unroll
    for (int o = STREAM_LANES / 2; o > 0; o >>= 1) {
      tot.playouts += __shfl_down_sync(mask, tot.playouts, o, STREAM_LANES);
      tot.plies += __shfl_down_sync(mask, tot.plies, o, STREAM_LANES);
      tot.decisions += __shfl_down_sync(mask, tot.decisions, o, STREAM_LANES);
      tot.draws += __shfl_down_sync(mask, tot.draws, o, STREAM_LANES);
      tot.flagged += __shfl_down_sync(mask, tot.flagged, o, STREAM_LANES);
      tot.sum += __shfl_down_sync(mask, tot.sum, o, STREAM_LANES);
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
