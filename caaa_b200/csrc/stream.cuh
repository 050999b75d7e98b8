// stream.cuh -- phase-regrouped rollout kernel ("phase server" CTAs).
//
// Included by caaa_gpu.cu (after RolloutArgs / LaneTotals / finish_playout).
//
// Why: in a player-turn only 10-45 % of the games have anything to do in a given phase, and the amount of
// work per game is heavy-tailed, so a warp that walks the pipeline with the same 32 games issues most of its
// instructions for 2-4 active lanes.  Here the games are REGROUPED BY PHASE inside every SM:
//   * every CTA (one per SM) owns a private pool of POOL_MULT x 224 games in global memory (924 B each;
//     the whole pool is ~90 MB and stays L2-resident) and steps them until the batch's ticket counter runs out;
//   * per pipeline phase (0..14, 15 = the end-of-turn item: purchase, bookkeeping, rotation, scoring, finish +
//     restart, 16 = free pool entries) the CTA keeps a bitmap of the pool entries waiting for that phase, in
//     shared memory;
//   * a warp claims up to 32 entries that all wait for the SAME phase (atomicAnd on the bitmap words), copies
//     them into its 32 shared-memory slots with cp.async (all lanes move one game at a time: coalesced, fully
//     pipelined), runs the phase with all lanes in lockstep (engine2.cuh), copies every game back and sets its
//     bit in the bitmap of the next phase that has work for it (per-game work mask);
//   * the warps of a CTA prefer the CTA's current phase, so they share its few KB of instructions in the SM's
//     instruction caches; when that bitmap runs low they move to the fullest one.
// No lane waits for a game of another phase, a game is only queued for phases in which it has something to
// do, and finished pool entries restart from the ticket counter, which absorbs the long tail of game lengths.
// Nothing is shared between SMs except the ticket counter, so there is no device-wide queue contention.
#pragma once

namespace caaa {

constexpr int NQ = 17;                   // phases 0..14, 15 = end of turn, 16 = free pool entries
constexpr int Q_EOT = CAAA_PH_BUY_UNITS; // 15
constexpr int Q_FREE = 16;
constexpr int STREAM_SLOTS = 224;        // shared-memory game slots per CTA
constexpr int MAX_POOL_MULT = 4;
constexpr int MAX_BM_WORDS = STREAM_SLOTS * MAX_POOL_MULT / 32;  // 28 bitmap words per phase

struct StreamShared {
  uint32_t bm[NQ][MAX_BM_WORDS];  // pool entries waiting per phase
  int cnt[NQ];                    // their number (selection heuristic; the bitmaps are the truth)
  int kind;                       // the phase this CTA currently prefers
  int retired;                    // pool entries that found no ticket left
  uint16_t batch[STREAM_SLOTS];   // claimed entries, LANES per warp
};
constexpr int STREAM_SMEM = TABLE_BYTES + STREAM_SLOTS * (int)sizeof(Game) + (int)((sizeof(StreamShared) + 15) / 16 * 16);

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

struct StreamDbg {
  unsigned long long batches, items, t[6];
};

// LANES = games per warp.  The kernel is bound by per-warp latency, not by issue slots, and a batch rarely
// fills 32 lanes with equally long work, so fewer games per warp and more warps per SM (224 / LANES) is the
// better trade: more independent instruction streams to hide latency, a shorter wait for the slowest lane.
template <bool STATS, int LANES>
__global__ void __launch_bounds__(STREAM_SLOTS / LANES * 32, 1) rollout_kernel_stream(const RolloutArgs a, uint32_t* pool_all, int pool_mult,
                                                                                       StreamDbg* dbg) {
  const DevTables* T = stage_tables(a.tables);
  StreamShared* S = reinterpret_cast<StreamShared*>(smem_raw + TABLE_BYTES + STREAM_SLOTS * sizeof(Game));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int pool_n = STREAM_SLOTS * pool_mult, bm_words = pool_n / 32;
  uint32_t* const pool = pool_all + (size_t)blockIdx.x * pool_n * GAME_WORDS;
  for (int i = threadIdx.x; i < NQ * MAX_BM_WORDS; i += blockDim.x) {
    const int q = i / MAX_BM_WORDS, w = i % MAX_BM_WORDS;
    S->bm[q][w] = (q == Q_FREE && w < bm_words) ? 0xffffffffu : 0u;
  }
  if (threadIdx.x < NQ) S->cnt[threadIdx.x] = threadIdx.x == Q_FREE ? pool_n : 0;
  if (threadIdx.x == 0) {
    S->kind = Q_FREE;
    S->retired = 0;
  }
  __syncthreads();
  Game* const warp_games = reinterpret_cast<Game*>(smem_raw + TABLE_BYTES) + warp * LANES;
  Game* const g = warp_games + (lane < LANES ? lane : 0);
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const bool one_state = a.rollouts_per_state >= a.n_games;
  volatile int* vcnt = S->cnt;
  volatile int* vkind = &S->kind;
  volatile int* vretired = &S->retired;
  LaneTotals tot;
  unsigned long long d_batches = 0, d_items = 0, d_t[6] = {0, 0, 0, 0, 0, 0};
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
    // ---- choose a phase: the CTA's current one while it can fill most of a warp, else the fullest ----
    const int c = lane < NQ ? vcnt[lane] : 0;
    int kind = *vkind;
    int c_kind = __shfl_sync(FULL, c, kind);
    if (c_kind <= 0) {
      // fullest bitmap; ties go to the later phase, which is closer to finishing a turn
      const unsigned key = __reduce_max_sync(FULL, c > 0 ? ((unsigned)c << 5) | (unsigned)lane : 0u);
      if (key == 0) {
        if (*vretired >= pool_n) break;  // every pool entry has retired: the batch is done
        __nanosleep(200);
        continue;
      }
      kind = (int)(key & 31u);
      c_kind = (int)(key >> 5);
      if (lane == 0) *vkind = kind;
    }
    // share the waiting entries between the warps of the CTA, so that they all serve this phase at the same time
    constexpr int WARPS = STREAM_SLOTS / LANES;
    int take = (c_kind + WARPS - 1) / WARPS;
    take = take < 1 ? 1 : (take > LANES ? LANES : take);
    // ---- claim up to LANES waiting entries: lane w takes bits of bitmap word w ----
    uint32_t word = lane < bm_words ? *reinterpret_cast<volatile uint32_t*>(&S->bm[kind][lane]) : 0u;
    int pc = __popc(word), incl = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    int allow = take - (incl - pc);
    allow = allow < 0 ? 0 : (allow > pc ? pc : allow);
    while (__popc(word) > allow) word &= ~(0x80000000u >> __clz(word));  // keep the lowest `allow` bits
    uint32_t claimed = word ? (atomicAnd(&S->bm[kind][lane], ~word) & word) : 0u;
    pc = __popc(claimed);
    incl = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += v;
    }
    const int n = __shfl_sync(FULL, incl, 31);
    if (n == 0) continue;
    if (lane == 0) atomicSub(&S->cnt[kind], n);
    {
      int o = incl - pc;
      while (claimed) {
        const int b = __ffs(claimed) - 1;
        claimed &= claimed - 1;
        S->batch[warp * LANES + o++] = (uint16_t)(lane * 32 + b);
      }
    }
    __syncwarp();
    const bool have = lane < n;
    const unsigned idx = have ? S->batch[warp * LANES + lane] : 0u;
    __threadfence_block();
    CAAA_LAP(0);  // idle + select + claim
    d_batches++;
    d_items += (unsigned)n;
    const unsigned mhave = __ballot_sync(FULL, have);
    bool need_start = have && kind == Q_FREE;
    bool live = have && kind != Q_FREE;
    if (kind != Q_FREE) {  // ---- copy the games in: all 32 lanes move one game at a time (coalesced, asynchronous) ----
      for (int j = 0; j < n; j++) {
        const unsigned ij = __shfl_sync(FULL, idx, j);
        const uint32_t* src = pool + (size_t)ij * GAME_WORDS;
        uint32_t* dst = reinterpret_cast<uint32_t*>(warp_games + j);
#pragma unroll
        for (int k = 0; k < (GAME_WORDS + 31) / 32; k++)
          if (k * 32 + lane < GAME_WORDS) cp_async4(dst + k * 32 + lane, src + k * 32 + lane);
      }
      cp_async_wait_all();
      __syncwarp();
      CAAA_LAP(1);  // copy in
      // ---- run the phase on all of them in lockstep ----
      if (have) {
        switch (kind) {
          case CAAA_PH_MOVE_FIGHTERS: move_fighter_units<false>(g, cx, mhave); break;
          case CAAA_PH_MOVE_BOMBERS: move_bomber_units<false>(g, cx, mhave); break;
          case CAAA_PH_STAGE_TRANSPORTS: stage_transport_units<false>(g, cx, mhave); break;
          case CAAA_PH_MOVE_TANKS: move_land_unit_type<false>(g, cx, mhave, TANKS, CAAA_PH_MOVE_TANKS); break;
          case CAAA_PH_MOVE_ARTILLERY: move_land_unit_type<false>(g, cx, mhave, ARTILLERY, CAAA_PH_MOVE_ARTILLERY); break;
          case CAAA_PH_MOVE_INFANTRY: move_land_unit_type<false>(g, cx, mhave, INFANTRY, CAAA_PH_MOVE_INFANTRY); break;
          case CAAA_PH_MOVE_TRANSPORTS: move_transport_units<false>(g, cx, mhave); break;
          case CAAA_PH_MOVE_SUBS: move_subs<false>(g, cx, mhave); break;
          case CAAA_PH_MOVE_SHIPS: move_destroyers_battleships<false>(g, cx, mhave); break;
          case CAAA_PH_SEA_BATTLES: resolve_sea_battles<false>(g, cx, mhave); break;
          case CAAA_PH_UNLOAD_TRANSPORTS: unload_transports<false>(g, cx, mhave); break;
          case CAAA_PH_LAND_BATTLES: resolve_land_battles<false>(g, cx, mhave); break;
          case CAAA_PH_MOVE_AA_GUNS: move_land_unit_type<false>(g, cx, mhave, AA_GUNS, CAAA_PH_MOVE_AA_GUNS); break;
          case CAAA_PH_LAND_FIGHTERS: land_fighter_units<false>(g, cx, mhave); break;
          case CAAA_PH_LAND_BOMBERS: land_bomber_units<false>(g, cx, mhave); break;
          default: {  // end of turn, reference src/engine.cpp:4229-4241
            buy_units<false>(g, cx, mhave);
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
        }
      }
      __syncwarp();
      CAAA_LAP(2);  // compute
    }
    // ---- (re)start pool entries from the ticket counter: one fetch per warp, cooperative copy of the start state ----
    bool retired = false;
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
        const unsigned long long gj = base + (unsigned)__popc(need & ((1u << j) - 1u));
        if (gj < a.n_games) {
          const unsigned long long sj = one_state ? 0ULL : gj / a.rollouts_per_state;
          const uint32_t* src = a.templates + sj * GAME_WORDS;
          uint32_t* dst = reinterpret_cast<uint32_t*>(warp_games + j);
#pragma unroll
          for (int k = 0; k < (GAME_WORDS + 31) / 32; k++)
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
    CAAA_LAP(3);  // restart
    // ---- route every live game to the next phase that has work for it and copy it back out ----
    int nxt = -1;
    if (live) {
      uint32_t w = g->work & W_ALL_MOVES;
      if (kind < Q_EOT) w &= ~((2u << kind) - 1u);
      nxt = w ? __ffs(w) - 1 : Q_EOT;
    }
    __syncwarp();
    for (int j = 0; j < n; j++) {
      if (__shfl_sync(FULL, nxt, j) < 0) continue;
      const unsigned ij = __shfl_sync(FULL, idx, j);
      uint32_t* dst = pool + (size_t)ij * GAME_WORDS;
      const uint32_t* src = reinterpret_cast<const uint32_t*>(warp_games + j);
#pragma unroll
      for (int k = 0; k < (GAME_WORDS + 31) / 32; k++)
        if (k * 32 + lane < GAME_WORDS) dst[k * 32 + lane] = src[k * 32 + lane];
    }
    __threadfence_block();
    __syncwarp();
    if (nxt >= 0) {
      atomicOr(&S->bm[nxt][idx >> 5], 1u << (idx & 31u));
      atomicAdd(&S->cnt[nxt], 1);
    }
    const unsigned nret = __popc(__ballot_sync(FULL, retired));
    if (lane == 0 && nret) atomicAdd(&S->retired, (int)nret);
    CAAA_LAP(4);  // copy out + publish
  }
#undef CAAA_LAP
  if (dbg != nullptr && lane == 0) {
    atomicAdd(&dbg->batches, d_batches);
    atomicAdd(&dbg->items, d_items);
    for (int k = 0; k < 6; k++) atomicAdd(&dbg->t[k], d_t[k]);
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
