// resident.cuh -- SM-resident, phase-regrouped rollout kernel (an experiment kept as CAAA_ROLLOUT_KERNEL=resident: measured
// 0.52-0.65 M playouts/s against 1.67 M for the pool kernel in stream.cuh, because every warp of an SM ends up in another phase;
// the same design is what the battle pool kernel in battle.cuh uses, where an SM only has four stages).
//
// Included by caaa_gpu.cu (after RolloutArgs / LaneTotals / finish_playout).
//
// Why: in a player-turn only 10-45 % of the games have anything to do in a given phase and the work per game
// is heavy-tailed, so a warp that walks the pipeline with the same 32 games issues most of its instructions
// for 2-4 lanes.  The games are therefore REGROUPED BY PHASE -- but, unlike round 1's global pool, without
// ever moving a game:
//   * every CTA (one per SM, persistent) keeps RES_SLOTS games in its shared memory for their whole life;
//     a playout touches global memory for its start state and its result only;
//   * per kind of work (pipeline phases 0..14 -- tanks/artillery/infantry and transports/subs/ships merged
//     into one visit each --, 15 = the end-of-turn item, 16 = free slots) the CTA keeps a bitmap of the
//     slots waiting for it, in shared memory;
//   * warps own no slots.  An idle warp picks the kind with the most waiting games, claims up to 32 of them
//     (one shared-memory atomic per bitmap word), runs the phase on them IN PLACE in lockstep (engine2.cuh)
//     and sets each game's bit in the bitmap of the next phase that has work for it.  A hop between phases
//     costs two shared-memory atomics instead of two 932-byte copies through L2;
//   * lane r only ever takes slots with slot % 32 == r (bit r of a bitmap word): the slot stride is an odd
//     number of words, so the 32 games of a batch sit in 32 different banks whatever subset was claimed;
//   * because hops are free, a phase may YIELD once part of its batch is done: the stragglers keep their
//     loop cursor (engine2.cuh), go back to the same queue and meet other stragglers there, instead of
//     holding 30 finished lanes hostage;
//   * finished slots restart from the batch's ticket counter (warp-cooperative copy of the pre-imported
//     start state), which absorbs the 14x tail of game lengths; an SM that runs out of tickets drains its
//     last games with all of its warps polling shared memory, so the tail of a batch is short.
#pragma once

namespace caaa {

#ifndef CAAA_HAVE_STREAM_KERNEL
constexpr int NQ = 17;                   // phases 0..14, 15 = end of turn (purchase, bookkeeping, rotation, scoring,
constexpr int Q_EOT = CAAA_PH_BUY_UNITS; // finish + restart), 16 = free slots
constexpr int Q_FREE = 16;
#endif
constexpr int RES_SLOTS = 240;           // games resident per CTA
constexpr int RES_WORDS = (RES_SLOTS + 31) / 32;

struct ResShared {
  uint32_t bm[NQ][RES_WORDS];  // slots waiting per kind: bit r of word w = slot w * 32 + r
  int retired;                 // slots that found no ticket left
  int progress;                // batches served (watchdog)
  int aborted;
  int pad_;
};
constexpr int RES_SMEM = TABLE_BYTES + RES_SLOTS * (int)sizeof(Game) + (int)((sizeof(ResShared) + 15) / 16 * 16);
static_assert(RES_SMEM <= 232448, "resident kernel: shared memory budget of one sm_100a CTA");

struct ResDbg {
  unsigned long long batches, items, t[4];             // t: select+claim+idle, phase code, restart, route
  unsigned long long kind_cycles[NQ], kind_items[NQ], kind_batches[NQ];
};

#ifndef CAAA_HAVE_STREAM_KERNEL
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }
#endif

// (Re)start the slots of the lanes with `need_start` from the batch's ticket counter: one fetch per warp,
// cooperative copy of the pre-imported start state.  Lanes that find no ticket left retire their slot.
template <bool STATS>
__device__ __noinline__ void restart_resident(const RolloutArgs& a, Game* slots, Game* g, unsigned idx, int lane, bool one_state,
                                              bool& need_start, bool& live, bool& retired, LaneTotals& tot) {
  constexpr int KW = (GAME_WORDS + 31) / 32;
  for (;;) {
    const unsigned need = __ballot_sync(FULL, need_start);
    if (need == 0) break;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(a.ticket, (unsigned long long)__popc(need));
    base = __shfl_sync(FULL, base, 0);
    unsigned m = need;
    while (m) {
      const int j = __ffs(m) - 1;
      m &= m - 1;
      const unsigned sj_slot = __shfl_sync(FULL, idx, j);
      const unsigned long long gj = base + (unsigned)__popc(need & ((1u << j) - 1u));
      if (gj < a.n_games) {
        const unsigned long long sj = one_state ? 0ULL : gj / a.rollouts_per_state;
        const uint32_t* src = a.templates + sj * GAME_WORDS;
        uint32_t* dst = reinterpret_cast<uint32_t*>(slots + sj_slot);
#pragma unroll 1
        for (int k = 0; k < KW; k++)
          if (k * 32 + lane < GAME_WORDS) cp_async4(dst + k * 32 + lane, src + k * 32 + lane);
      }
    }
    cp_async_wait_all();
    __syncwarp();
    if (need_start) {
      const unsigned long long gi = base + (unsigned)__popc(need & ((1u << lane) - 1u));
      if (gi < a.n_games) {
        const unsigned long long id = a.first_game_id + gi;
        g->game_lo = (uint32_t)id;
        g->game_hi = (uint32_t)(id >> 32);
        const double score = a.scores0[one_state ? 0ULL : gi / a.rollouts_per_state];
        if (score > 0.01 && score < 0.99) {
          live = true;
          need_start = false;
        } else {
          finish_playout<STATS>(g, score, gi, a, tot);  // already terminal: zero turns, take another ticket
        }
      } else {
        need_start = false;
        retired = true;
      }
    }
  }
}

// MAXW = most warps the CTA may be launched with (register budget); the launch chooses blockDim <= MAXW * 32.
template <bool STATS, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) rollout_kernel_resident(const RolloutArgs a, int n_slots, ResDbg* dbg) {
  const DevTables* T = stage_tables(a.tables);
  ResShared* S = reinterpret_cast<ResShared*>(smem_raw + TABLE_BYTES + RES_SLOTS * sizeof(Game));
  Game* const slots = reinterpret_cast<Game*>(smem_raw + TABLE_BYTES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < NQ * RES_WORDS) {
    const int q = threadIdx.x / RES_WORDS, w = threadIdx.x % RES_WORDS;
    const int nbits = n_slots - w * 32;
    const uint32_t valid = nbits >= 32 ? 0xffffffffu : nbits <= 0 ? 0u : ((1u << nbits) - 1u);
    S->bm[q][w] = q == Q_FREE ? valid : 0u;
  }
  if (threadIdx.x == 0) {
    S->retired = 0;
    S->progress = 0;
    S->aborted = 0;
  }
  __syncthreads();
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const bool one_state = a.rollouts_per_state >= a.n_games;
  volatile uint32_t* vbm = &S->bm[0][0];
  volatile int* vretired = &S->retired;
  volatile int* vprogress = &S->progress;
  volatile int* vaborted = &S->aborted;
  LaneTotals tot;
  int last_kind = -1;
  int seen_progress = -1, thin_polls = 0;
  long long last_change = clock64();
  unsigned long long d_batches = 0, d_items = 0, d_t[4] = {0, 0, 0, 0};
  long long d_c0 = clock64();
#define CAAA_LAP(k)                                   \
  do {                                                \
    if (dbg != nullptr) {                             \
      const long long c_ = clock64();                 \
      d_t[k] += (unsigned long long)(c_ - d_c0);      \
      d_c0 = c_;                                      \
    }                                                 \
  } while (0)

  for (;;) {
    // ---- choose the kind of work with the most claimable games (one per bank residue); ties go to the later
    // phase, which is closer to finishing a turn; `stick` favours the kind this warp served last (its code is
    // warm in the instruction caches) ----
    uint32_t any = 0;
    if (lane < NQ) {
#pragma unroll
      for (int w = 0; w < RES_WORDS; w++) any |= vbm[lane * RES_WORDS + w];
    }
    const int avail = __popc(any);
    const unsigned key = __reduce_max_sync(FULL, avail ? ((unsigned)(avail + (lane == last_kind ? a.stick : 0)) << 5) | (unsigned)lane : 0u);
    if (key == 0) {
      if (*vretired >= n_slots) break;  // every slot has retired: this CTA is done
      // watchdog: nothing claimable and no batch served by any warp of the CTA for seconds although slots are
      // still live -- give up rather than hang the GPU (it would mean a lost slot, i.e. a bug in this scheduler)
      if (*vaborted) break;
      const int p = *vprogress;
      const long long now = clock64();
      if (p != seen_progress) {
        seen_progress = p;
        last_change = now;
      } else if (now - last_change > 8000000000LL) {
        if (lane == 0) {
          S->aborted = 1;
          if (a.abort_flag != nullptr) atomicExch(a.abort_flag, 1);
        }
        break;
      }
      __nanosleep(a.idle_ns);
      continue;
    }
    // a thin batch may wait a few polls for more games of its kind to arrive (min_batch, off by default)
    if ((int)(key >> 5) < a.min_batch && thin_polls < a.min_batch_polls) {
      thin_polls++;
      __nanosleep(a.idle_ns);
      continue;
    }
    thin_polls = 0;
    const int kind = (int)(key & 31u);
    // ---- claim: bit r of a bitmap word belongs to lane r, so a whole word is claimed with one atomic ----
    unsigned have = 0;
    int my_w = -1;
    {
      int w = warp % RES_WORDS;  // warps start at different words
#pragma unroll 1
      for (int i = 0; i < RES_WORDS; i++, w = (w + 1 == RES_WORDS ? 0 : w + 1)) {
        const uint32_t want = vbm[kind * RES_WORDS + w] & ~have;
        if (want == 0) continue;
        uint32_t old = 0;
        if (lane == 0) old = atomicAnd(&S->bm[kind][w], ~want);
        const uint32_t got = __shfl_sync(FULL, old, 0) & want;
        if ((got >> lane) & 1u) my_w = w;
        have |= got;
        if (have == FULL) break;
      }
    }
    if (have == 0) continue;  // another warp was faster
    __threadfence_block();
    if (lane == 0) S->progress = S->progress + 1;
    last_kind = kind;
    const bool runs = my_w >= 0;
    const unsigned idx = runs ? (unsigned)(my_w * 32 + lane) : 0u;
    Game* const g = slots + idx;
    const int nh = __popc(have);
    d_batches++;
    d_items += (unsigned)nh;
    CAAA_LAP(0);  // idle + select + claim
    // ---- run the phase on the claimed games in lockstep.  yield_pct > 0: the phase yields once that share of
    // the batch is done; the stragglers are queued for the same kind again (cursor saved in the game) ----
    const int pct = (kind == CAAA_PH_SEA_BATTLES || kind == CAAA_PH_LAND_BATTLES) && a.yield_battles > 0 ? a.yield_battles : a.yield_lanes;
    const int yield = (pct > 0 && nh >= a.yield_min) ? (nh * pct + 99) / 100 : 0;
    int res = PH_DONE;
    bool live = runs, need_start = false;
    if (runs) {
      switch (kind) {
        case CAAA_PH_MOVE_FIGHTERS: res = move_fighter_units<false>(g, cx, have, yield); break;
        case CAAA_PH_MOVE_BOMBERS: res = move_bomber_units<false>(g, cx, have, yield); break;
        case CAAA_PH_STAGE_TRANSPORTS: res = stage_transport_units<false>(g, cx, have, yield); break;
        // tanks, artillery and infantry are consecutive phases of one mover: one visit, one merged flat loop
        // (queues 4 and 5 stay empty)
        case CAAA_PH_MOVE_TANKS: res = move_land_units<false>(g, cx, have, CAAA_PH_MOVE_TANKS, CAAA_PH_MOVE_INFANTRY, yield); break;
        // ... and so are loaded transports, submarines and surface ships (queues 7 and 8 stay empty)
        case CAAA_PH_MOVE_TRANSPORTS: res = move_sea_units<false>(g, cx, have, yield); break;
        case CAAA_PH_SEA_BATTLES: res = resolve_sea_battles<false>(g, cx, have, yield); break;
        case CAAA_PH_UNLOAD_TRANSPORTS: res = unload_transports<false>(g, cx, have, yield); break;
        case CAAA_PH_LAND_BATTLES: res = resolve_land_battles<false>(g, cx, have, yield); break;
        case CAAA_PH_MOVE_AA_GUNS: res = move_land_unit_type<false>(g, cx, have, AA_GUNS, CAAA_PH_MOVE_AA_GUNS, yield); break;
        case CAAA_PH_LAND_FIGHTERS: res = land_fighter_units<false>(g, cx, have, yield); break;
        case CAAA_PH_LAND_BOMBERS: res = land_bomber_units<false>(g, cx, have, yield); break;
        case Q_EOT: {  // reference src/engine.cpp:4229-4241
          res = buy_units<false>(g, cx, have, yield);
          if (res != PH_YIELDED) {
            crash_air_units(g, cx);
            reset_units_fully(g);
            collect_money(g, cx);
            rotate_turns(g, cx);
            g->turns++;
            const double score = evaluate_state(g, cx);
            if (!(score > 0.01 && score < 0.99 && g->turns < 1000)) {
              const unsigned long long gi = ((unsigned long long)g->game_hi << 32 | g->game_lo) - a.first_game_id;
              finish_playout<STATS>(g, score, gi, a, tot);
              live = false;
              need_start = true;
            }
          }
          break;
        }
        default:  // Q_FREE: an empty slot
          live = false;
          need_start = true;
          break;
      }
    }
    __syncwarp();
    if (dbg != nullptr && lane == 0) {
      atomicAdd(&dbg->kind_cycles[kind], (unsigned long long)(clock64() - d_c0));
      atomicAdd(&dbg->kind_items[kind], (unsigned long long)nh);
      atomicAdd(&dbg->kind_batches[kind], 1ULL);
    }
    CAAA_LAP(1);  // phase code
    // ---- (re)start finished / free slots from the ticket counter (out of line: once per playout) ----
    bool retired = false;
    if (__any_sync(FULL, need_start)) restart_resident<STATS>(a, slots, g, idx, lane, one_state, need_start, live, retired, tot);
    CAAA_LAP(2);  // restart
    // ---- route every game to the next kind that has work for it: one bit in that kind's bitmap ----
    int nxt = -1;
    if (runs && live && res == PH_YIELDED) {
      nxt = kind;
    } else if (runs && live) {
      uint32_t w = g->work & W_ALL_MOVES;
      const int last = kind == CAAA_PH_MOVE_TANKS ? (int)CAAA_PH_MOVE_INFANTRY : kind == CAAA_PH_MOVE_TRANSPORTS ? (int)CAAA_PH_MOVE_SHIPS : kind;
      if (kind < Q_EOT) w &= ~((2u << last) - 1u);  // (after the end of a turn or a fresh start every phase is open again)
      nxt = w ? __ffs(w) - 1 : Q_EOT;
      if (nxt == CAAA_PH_MOVE_ARTILLERY || nxt == CAAA_PH_MOVE_INFANTRY) nxt = CAAA_PH_MOVE_TANKS;  // the merged land mover
      if (nxt == CAAA_PH_MOVE_SUBS || nxt == CAAA_PH_MOVE_SHIPS) nxt = CAAA_PH_MOVE_TRANSPORTS;      // the merged sea mover
    }
    __threadfence_block();
    if (nxt >= 0) atomicOr(&S->bm[nxt][my_w], 1u << lane);
    const unsigned nret = __popc(__ballot_sync(FULL, retired));
    if (lane == 0 && nret) atomicAdd(&S->retired, (int)nret);
    CAAA_LAP(3);  // route
  }
#undef CAAA_LAP
  if (dbg != nullptr && lane == 0) {
    atomicAdd(&dbg->batches, d_batches);
    atomicAdd(&dbg->items, d_items);
    for (int k = 0; k < 4; k++) atomicAdd(&dbg->t[k], d_t[k]);
  }

  if (a.summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      tot.playouts += __shfl_down_sync(FULL, tot.playouts, o);
      tot.plies += __shfl_down_sync(FULL, tot.plies, o);
      tot.decisions += __shfl_down_sync(FULL, tot.decisions, o);
      tot.draws += __shfl_down_sync(FULL, tot.draws, o);
      tot.flagged += __shfl_down_sync(FULL, tot.flagged, o);
      tot.sum += __shfl_down_sync(FULL, tot.sum, o);
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
