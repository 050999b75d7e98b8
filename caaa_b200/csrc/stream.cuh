// stream.cuh -- phase-regrouped rollout kernel ("phase server" CTAs).
//
// Included by caaa_gpu.cu (after RolloutArgs / LaneTotals / finish_playout).
//
// Why: in a player-turn only 10-45 % of the games have anything to do in a given phase, and the amount of
// work per game is heavy-tailed, so a warp that walks the pipeline with the same 32 games issues most of its
// instructions for 2-4 active lanes.  Here the games are REGROUPED BY PHASE:
//   * the CTAs (one per SM, persistent) form two groups of 74; every group owns a pool of 2 x 240 games per CTA in
//     global memory (932 B each; 66 MB in total, L2-resident) and steps them until the batch's ticket counter
//     runs out;
//   * per kind of work (pipeline phases 0..14 -- with tanks/artillery/infantry and transports/subs/ships merged
//     into one visit each --, 15 = the end-of-turn item: purchase, bookkeeping, rotation, scoring, finish +
//     restart, 16 = free pool entries) the group keeps a bitmap of the pool entries waiting for it;
//   * a warp claims up to LANES entries that all wait for the SAME kind (atomicAnd on bitmap words), copies them
//     into its shared-memory slots (all 32 lanes move one game at a time: coalesced L2 loads, 4 games in
//     flight), runs the phase on them in lockstep (engine2.cuh), copies every game back and sets its bit in the
//     bitmap of the next phase that has work for it (per-game work mask);
//   * a CTA is STICKY: its warps serve the CTA's current kind until one of them has found that queue empty
//     `patience` times in a row, and only then does the CTA move to the kind with the most entries per serving
//     CTA.  The kernel is bound by instruction delivery (the GPC-level instruction cache ran at 91 % of its request
//     rate when CTAs changed kind every 3 batches, profiles/r2_summary.md): one CTA, one phase's few KB of code.
//     74 SMs pooling their games is what lets every kind keep CTAs of its own;
//   * the three heaviest kinds have a second, "heavy" queue (work-size buckets, see Q_LAND_HEAVY below): a batch runs as
//     long as its longest game, so it should be formed from games of similar length.
// No lane waits for a game of another phase, a game is only queued for phases in which it has something to
// do, and finished pool entries restart from the ticket counter, which absorbs the long tail of game lengths.
// The only device-wide shared word is the ticket counter; queues are per group, so claims do not contend GPU-wide
// (a single set of device-wide FIFO queues was measured in round 1: 92 % of the time went into claim contention).
#pragma once

namespace caaa {

constexpr int NQ = 20;                   // phases 0..14, 15 = end of turn (purchase, bookkeeping, rotation, scoring,
constexpr int Q_EOT = CAAA_PH_BUY_UNITS; // finish + restart), 16 = free pool entries
constexpr int Q_FREE = 16;
// Work-size buckets: the three heaviest kinds have a second queue for games that are expected to need many decision steps
// (many movable units / much money), so that a batch -- which runs as long as its longest game -- is formed from games of
// similar length.  The code a CTA runs for a heavy queue is that of the base kind.
constexpr int Q_LAND_HEAVY = 17, Q_SEA_HEAVY = 18, Q_EOT_HEAVY = 19;
__device__ __forceinline__ int base_kind(int q) {
  return q == Q_LAND_HEAVY ? (int)CAAA_PH_MOVE_TANKS : q == Q_SEA_HEAVY ? (int)CAAA_PH_MOVE_TRANSPORTS : q == Q_EOT_HEAVY ? Q_EOT : q;
}
constexpr int STREAM_SLOTS = 240;        // shared-memory game slots per CTA: 15 warps x 16 games, all of the 227 KB
constexpr int MAX_POOL_MULT = 4;
constexpr int MAX_BM_WORDS = STREAM_SLOTS * MAX_POOL_MULT / 32;  // 28 bitmap words per phase

constexpr int MAX_GROUP = 74;            // most SMs that may share one pool and one set of phase bitmaps: pool entries are
                                         // 16-bit indices, 74 x 480 = 35 520 of them
constexpr int MAX_BM_WORDS_G = MAX_BM_WORDS * MAX_GROUP;

// One per group of CTAs, in global memory (L2): the CTAs of a group pool their games, so that a CTA finds
// enough games waiting for ONE phase to keep all of its warps in that phase's code.
struct GroupState {
  uint32_t bm[NQ][MAX_BM_WORDS_G];  // pool entries waiting per phase
  int cnt[NQ];                      // their number (selection heuristic; the bitmaps are the truth)
  int servers[NQ];                  // CTAs currently serving each phase
  int retired;                      // pool entries that found no ticket left
  int aborted;                      // set by a warp that saw no work for seconds although games are left (watchdog)
  int pad_[2];
};
struct StreamShared {
  int kind;                       // the phase this CTA currently serves
  uint16_t batch[STREAM_SLOTS];   // claimed entries, LANES per warp
};
// A pool entry / shared-memory slot: the Game padded to 936 bytes = 117 double words.  Entries and slots are then 8-byte
// aligned, so a game moves between pool and slot with 64-bit accesses (4 per lane instead of 8: the copies are a quarter
// of the scheduler's instructions), and 234 words per slot keeps the 16 games of a warp in 16 different banks
// (234 j mod 32 is distinct for j < 16; an odd stride, as the 32-lane kernels use, would not be 8-byte aligned).
struct alignas(8) PoolGame {
  Game g;
  uint32_t pad_;
};
static_assert(sizeof(PoolGame) == 936 && (sizeof(PoolGame) / 4) % 4 == 2, "slot stride: 8-byte aligned, bank-conflict free for 16 lanes");
constexpr int POOL_DW = (int)(sizeof(PoolGame) / 8);  // 117
constexpr int STREAM_SMEM = TABLE_BYTES + STREAM_SLOTS * (int)sizeof(PoolGame) + (int)((sizeof(StreamShared) + 15) / 16 * 16);
static_assert(STREAM_SMEM <= 232448, "phase-regrouped kernel: shared memory budget of one sm_100a CTA");
static_assert(MAX_GROUP * STREAM_SLOTS * MAX_POOL_MULT <= 65536 * 2 && MAX_GROUP * STREAM_SLOTS * 2 <= 65536, "pool entries are 16-bit indices at the default pool size");

__global__ void group_init_kernel(GroupState* G, int n_groups, int entries_per_cta, int group_size, int grid) {
  GroupState* gs = G + blockIdx.x;
  if ((int)blockIdx.x >= n_groups) return;
  const int first = (int)blockIdx.x * group_size;
  const int pool_g = entries_per_cta * ((grid - first) < group_size ? (grid - first) : group_size);
  const int words = pool_g / 32;
  for (int i = threadIdx.x; i < NQ * MAX_BM_WORDS_G; i += blockDim.x) {
    const int q = i / MAX_BM_WORDS_G, w = i % MAX_BM_WORDS_G;
    gs->bm[q][w] = (q == Q_FREE && w < words) ? 0xffffffffu : 0u;
  }
  if (threadIdx.x < NQ) {
    gs->cnt[threadIdx.x] = threadIdx.x == Q_FREE ? pool_g : 0;
    gs->servers[threadIdx.x] = 0;
  }
  if (threadIdx.x == 0) {
    gs->retired = 0;
    gs->aborted = 0;
  }
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory"); }

// Orders this thread's pool / bitmap accesses at GPU scope (message passing between CTAs: data, fence, flag).
// The pool is read and written with .cg accesses that bypass L1, so the acquire-release fence is enough;
// __threadfence() (fence.sc + L1 invalidation) was 5 % of the kernel's stall samples in round 1.
__device__ __forceinline__ void fence_gpu() { asm volatile("fence.acq_rel.gpu;" ::: "memory"); }

__device__ __forceinline__ unsigned long long ld_vol_u64(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}

struct StreamDbg {
  unsigned long long batches, items, t[6];
  unsigned long long kind_cycles[NQ], kind_items[NQ], kind_batches[NQ];  // phase-code time, games and batches per queue
};

// (Re)start the pool entries of the lanes with `need_start` from the batch's ticket counter: one fetch per warp,
// cooperative copy of the pre-imported start state.  Lanes that find no ticket left retire their entry.
template <bool STATS>
__device__ __noinline__ void restart_lanes(const RolloutArgs& a, PoolGame* warp_games, Game* g, int lane, bool one_state, bool& need_start,
                                           bool& live, bool& retired, LaneTotals& tot) {
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
      const unsigned long long gj = base + (unsigned)__popc(need & ((1u << j) - 1u));
      if (gj < a.n_games) {
        const unsigned long long sj = one_state ? 0ULL : gj / a.rollouts_per_state;
        const uint32_t* src = a.templates + sj * GAME_WORDS;
        uint32_t* dst = reinterpret_cast<uint32_t*>(&warp_games[j].g);
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

// One visit of `kind` for the lanes in `mask` (all of them call this together).  Returns PH_DONE / PH_YIELDED;
// live = false, need_start = true when the game ended in this visit.
template <bool STATS>
__device__ __forceinline__ int run_kind(int kind, Game* g, const RunCtx cx, unsigned mask, int yield, const RolloutArgs& a, LaneTotals& tot,
                                        bool& live, bool& need_start) {
  int res = PH_DONE;
  switch (kind) {
    case CAAA_PH_MOVE_FIGHTERS: res = move_fighter_units<false>(g, cx, mask, yield); break;
    case CAAA_PH_MOVE_BOMBERS: res = move_bomber_units<false>(g, cx, mask, yield); break;
    case CAAA_PH_STAGE_TRANSPORTS: res = stage_transport_units<false>(g, cx, mask, yield); break;
    // tanks, artillery and infantry are consecutive phases of one mover: one visit, one merged flat loop
    // (queues 4 and 5 stay empty)
    case CAAA_PH_MOVE_TANKS: res = move_land_units<false>(g, cx, mask, CAAA_PH_MOVE_TANKS, CAAA_PH_MOVE_INFANTRY, yield); break;
    // ... and so are loaded transports, submarines and surface ships (queues 7 and 8 stay empty)
    case CAAA_PH_MOVE_TRANSPORTS: res = move_sea_units<false>(g, cx, mask, yield); break;
    case CAAA_PH_SEA_BATTLES: res = resolve_sea_battles<false>(g, cx, mask, yield); break;
    case CAAA_PH_UNLOAD_TRANSPORTS: res = unload_transports<false>(g, cx, mask, yield); break;
    case CAAA_PH_LAND_BATTLES: res = resolve_land_battles<false>(g, cx, mask, yield); break;
    case CAAA_PH_MOVE_AA_GUNS: res = move_land_unit_type<false>(g, cx, mask, AA_GUNS, CAAA_PH_MOVE_AA_GUNS, yield); break;
    case CAAA_PH_LAND_FIGHTERS: res = land_fighter_units<false>(g, cx, mask, yield); break;
    case CAAA_PH_LAND_BOMBERS: res = land_bomber_units<false>(g, cx, mask, yield); break;
    case Q_EOT: {  // reference src/engine.cpp:4229-4241
      res = buy_units<false>(g, cx, mask, yield);
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
    default:  // Q_FREE: an empty pool entry
      live = false;
      need_start = true;
      break;
  }
  return res;
}

// LANES = games per warp (16: 15 warps per SM).  A warp serves ONE kind of work at a time: it claims games waiting
// for it, runs the phase on them in lockstep and routes them on.  Optionally (yield_lanes / yield_battles, off by
// default: measured neutral) a phase yields once a share of its batch is done; the stragglers keep their loop cursor
// in the game (engine2.cuh), are queued for the same phase again and meet other stragglers there.
// DBG = per-section cycle accounting (CAAA_DEBUG); compiled out of the shipped instantiation: the scheduler loop is
// hot in the instruction caches of every SM, and a few dozen instructions there cost 2 % of the throughput.
template <bool STATS, int LANES, bool DBG = false>
__global__ void __launch_bounds__(STREAM_SLOTS / LANES * 32, 1) rollout_kernel_stream(const RolloutArgs a, uint32_t* pool_all, int pool_mult,
                                                                                       GroupState* groups, int group_size, StreamDbg* dbg) {
  const DevTables* T = stage_tables(a.tables);
  StreamShared* S = reinterpret_cast<StreamShared*>(smem_raw + TABLE_BYTES + STREAM_SLOTS * sizeof(PoolGame));
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int group = blockIdx.x / group_size;
  const int first_cta = group * group_size;
  const int ctas_here = ((int)gridDim.x - first_cta) < group_size ? ((int)gridDim.x - first_cta) : group_size;
  const int pool_n = STREAM_SLOTS * pool_mult * ctas_here, bm_words = pool_n / 32;
  GroupState* const GS = groups + group;
  uint2* const pool = reinterpret_cast<uint2*>(pool_all) + (size_t)first_cta * STREAM_SLOTS * pool_mult * POOL_DW;
  if (threadIdx.x == 0) S->kind = -1;
  __syncthreads();
  PoolGame* const warp_games = reinterpret_cast<PoolGame*>(smem_raw + TABLE_BYTES) + warp * LANES;
  Game* const g = &warp_games[lane < LANES ? lane : 0].g;
  uint16_t* const batch = S->batch + warp * LANES;
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const bool one_state = a.rollouts_per_state >= a.n_games;
  volatile int* vcnt = GS->cnt;
  volatile int* vsrv = GS->servers;
  volatile int* vkind = &S->kind;
  volatile int* vretired = &GS->retired;
  unsigned scan_rot = (unsigned)((blockIdx.x * 7 + warp * 3) * 32) % (unsigned)(((bm_words + 31) / 32) * 32);  // where this warp starts scanning
  constexpr int KD = (POOL_DW + 31) / 32;  // 64-bit accesses per lane and game
  LaneTotals tot;
  long long last_work = clock64();
  int seen_kind = -2, empty_polls = 0;
  bool was_idle = false;
  unsigned long long d_batches = 0, d_items = 0, d_t[6] = {0, 0, 0, 0, 0, 0};
  long long d_c0 = clock64();
#define CAAA_LAP(k)                                   \
  do {                                                \
    if (DBG) {                                        \
      const long long c_ = clock64();                 \
      d_t[k] += (unsigned long long)(c_ - d_c0);      \
      d_c0 = c_;                                      \
    }                                                 \
  } while (0)

  // claim up to `want` entries waiting for `kind` into batch[]; each lane takes bits of one bitmap word per round
  auto claim = [&](int kind, int want) -> int {
    int n = 0;
    const int rounds = (bm_words + 31) / 32;
    // a claim looks at no more than `scan_rounds` x 32 words, starting where the previous one stopped: entries it
    // misses are found by a later claim, and a nearly empty phase does not cost a scan of the whole bitmap
    const int look = rounds < a.scan_rounds ? rounds : a.scan_rounds;
    for (int r = 0; r < look && n < want; r++) {
      int wi = (int)scan_rot + r * 32 + lane;  // scan_rot < rounds * 32
      if (wi >= rounds * 32) wi -= rounds * 32;
      uint32_t word = wi < bm_words ? *reinterpret_cast<volatile uint32_t*>(&GS->bm[kind][wi]) : 0u;
      if (!__any_sync(FULL, word != 0)) continue;
      int pc = __popc(word), incl = pc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      int allow = want - n - (incl - pc);
      allow = allow < 0 ? 0 : (allow > pc ? pc : allow);
      while (__popc(word) > allow) word &= ~(0x80000000u >> __clz(word));  // keep the lowest `allow` bits
      uint32_t claimed = word ? (atomicAnd(&GS->bm[kind][wi], ~word) & word) : 0u;
      pc = __popc(claimed);
      incl = pc;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += v;
      }
      int o = n + incl - pc;
      while (claimed) {
        const int b = __ffs(claimed) - 1;
        claimed &= claimed - 1;
        batch[o++] = (uint16_t)(wi * 32 + b);
      }
      n += __shfl_sync(FULL, incl, 31);
    }
    scan_rot += 32u * (unsigned)look;
    while (scan_rot >= (unsigned)(rounds * 32)) scan_rot -= (unsigned)(rounds * 32);
    if (n > 0 && lane == 0) atomicSub(&GS->cnt[kind], n);
    __syncwarp();
    return n;
  };

  for (;;) {
    // ---- choose a phase: the CTA's current one while enough entries wait for it, else the one with the most
    // entries per serving CTA (ties go to the later phase, which is closer to finishing a turn) ----
    // Stickiness matters more than queue length: the SMs of a GPC share one instruction cache behind their own, and
    // round 1's rule (leave a phase as soon as fewer than 8 entries wait for it) made a CTA change phase every 3.2
    // batches -- its 15 warps were in ~5 different phase functions at any time, the SM instruction cache hit rate was
    // 78 % and the GPC-level cache ran at 91 % of its request rate: THE limiter of the kernel (profiles/r2_summary.md).
    // With `patience` > 0 a CTA keeps its phase until a warp has found its queue empty `patience` times in a row
    // (sleeping in between, while other CTAs publish into it), and only then moves on.
    int kind = *vkind;
    if (kind != seen_kind) {
      seen_kind = kind;
      empty_polls = 0;
    }
    if (kind < 0 || empty_polls >= a.patience) {
      const int c = lane < NQ ? vcnt[lane] : 0;
      empty_polls = 0;
      const int srv = lane < NQ ? vsrv[lane] - (lane == kind ? 1 : 0) : 0;
      const unsigned score = c > 0 ? (unsigned)((c << 4) / (srv + 1)) + 1u : 0u;
      const unsigned key = __reduce_max_sync(FULL, score ? (score << 5) | (unsigned)lane : 0u);
      if (key == 0) {
        if (*vretired >= pool_n) break;  // every pool entry has retired: the batch is done
        // watchdog: games are left but NO queue of the group has held an entry for ~4 s without interruption -- give up rather than hang the GPU;
        // the host reports the failure (it would mean a lost pool entry, i.e. a bug in this scheduler)
        if (*reinterpret_cast<volatile int*>(&GS->aborted)) break;
        if (!was_idle) {  // the clock runs from the moment the queues were first seen empty, not from this warp's last claim
          was_idle = true;
          last_work = clock64();
        }
        if (clock64() - last_work > 8000000000LL) {
          if (lane == 0) {
            atomicExch(&GS->aborted, 1);
            if (a.abort_flag != nullptr) atomicExch(a.abort_flag, 1);
          }
          break;
        }
        __nanosleep(200);
        continue;
      }
      was_idle = false;
      const int nk = (int)(key & 31u);
      if (lane == 0) {
        const int old = atomicExch(&S->kind, nk);  // several warps may switch at once: account for the change once
        if (old != nk) {
          if (old >= 0) atomicSub(&GS->servers[old], 1);
          atomicAdd(&GS->servers[nk], 1);
          if (DBG) atomicAdd(&dbg->t[5], 1ULL);  // kind switches of CTAs
        }
      }
      kind = nk;
      seen_kind = nk;
    }
    bool have = false;   // this lane holds a game of phase `kind`
    unsigned idx = 0;    // its pool entry
    {  // ---- serve `kind`: claim a batch, run it, route the games on ----
      const int n_new = claim(kind, LANES);
      if (n_new == 0) {
        empty_polls++;
        __nanosleep(a.idle_ns);
        continue;
      }
      empty_polls = 0;
      const int queue = kind;  // the queue the batch came from; `kind` is the code that runs on it
      kind = base_kind(queue);
      CAAA_LAP(0);  // idle + select + claim
      if (n_new > 0) {
        last_work = clock64();
        if (DBG) {
          d_batches++;
          d_items += (unsigned)n_new;
        }
        fence_gpu();
        // the j-th claimed entry goes to lane j's slot
        if (lane < n_new) {
          idx = batch[lane];
          have = true;
        }
        __syncwarp();
        if (kind != Q_FREE) {  // copy the games in: all 32 lanes move one game at a time (coalesced), 4 games in flight
          for (int j0 = 0; j0 < n_new; j0 += 4) {
            uint2 r[4][KD];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
              const unsigned ij = batch[(j0 + jj) < n_new ? (j0 + jj) : 0];
              const uint2* src = pool + (size_t)ij * POOL_DW;
#pragma unroll
              for (int k = 0; k < KD; k++)
                if (j0 + jj < n_new && k * 32 + lane < POOL_DW) r[jj][k] = __ldcg(src + k * 32 + lane);
            }
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
              uint2* dst = reinterpret_cast<uint2*>(warp_games + ((j0 + jj) < n_new ? (j0 + jj) : 0));
#pragma unroll
              for (int k = 0; k < KD; k++)
                if (j0 + jj < n_new && k * 32 + lane < POOL_DW) dst[k * 32 + lane] = r[jj][k];
            }
          }
        }
        __syncwarp();
        CAAA_LAP(1);  // copy in
      }
      const bool runs = have;
      const unsigned mhave = __ballot_sync(FULL, runs);
      // ---- run the phase on all held games in lockstep ----
      // yield_pct > 0: the phase yields once that share of the batch is done; the stragglers are queued for the same
      // phase again (cursor saved in the game) and meet other stragglers there.  yield_battles_pct applies to the two
      // battle phases only, whose round counts have by far the longest tail.
      const int nh = __popc(mhave);
      const int pct = (kind == CAAA_PH_SEA_BATTLES || kind == CAAA_PH_LAND_BATTLES) && a.yield_battles > 0 ? a.yield_battles : a.yield_lanes;
      const int yield = (pct > 0 && nh >= 6) ? (nh * pct + 99) / 100 : 0;
      int res = PH_DONE;
      bool live = runs, need_start = false;
      if (runs) res = run_kind<STATS>(kind, g, cx, mhave, yield, a, tot, live, need_start);
      __syncwarp();
      if (DBG && lane == 0) {
        atomicAdd(&dbg->kind_cycles[queue], (unsigned long long)(clock64() - d_c0));
        atomicAdd(&dbg->kind_items[queue], (unsigned long long)n_new);
        atomicAdd(&dbg->kind_batches[queue], 1ULL);
      }
      CAAA_LAP(2);  // compute
      // ---- (re)start pool entries from the ticket counter (out of line: it runs once per playout, and the hot loop of
      // this kernel should stay small in the instruction cache) ----
      bool retired = false;
      if (__any_sync(FULL, need_start)) restart_lanes<STATS>(a, warp_games, g, lane, one_state, need_start, live, retired, tot);
      CAAA_LAP(3);  // restart
      // ---- route every finished game to the next phase that has work for it and copy it back out; games that
      // yielded stay in their slots ----
      int nxt = -1;
      if (runs && live && res == PH_YIELDED) {
        nxt = queue;
      } else if (runs && live) {
        uint32_t w = g->work & W_ALL_MOVES;
        const int last = kind == CAAA_PH_MOVE_TANKS ? (int)CAAA_PH_MOVE_INFANTRY : kind == CAAA_PH_MOVE_TRANSPORTS ? (int)CAAA_PH_MOVE_SHIPS : kind;
        if (kind < Q_EOT) w &= ~((2u << last) - 1u);  // (after the end of a turn or a fresh start every phase is open again)
        nxt = w ? __ffs(w) - 1 : Q_EOT;
        if (nxt == CAAA_PH_MOVE_ARTILLERY || nxt == CAAA_PH_MOVE_INFANTRY) nxt = CAAA_PH_MOVE_TANKS;  // the merged land mover
        if (nxt == CAAA_PH_MOVE_SUBS || nxt == CAAA_PH_MOVE_SHIPS) nxt = CAAA_PH_MOVE_TRANSPORTS;      // the merged sea mover
        {  // expected work of the next visit -> light or heavy queue of that kind (a sea-battle bucket by units engaged
           // was measured too: no gain)
          const int cur = g->cur;
          if (nxt == CAAA_PH_MOVE_TANKS) {
            int n = 0;
#pragma unroll
            for (int l = 0; l < NL; l++) {
              const uint8_t* r = g->lu[l];
              n += r[off_lu(TANKS) + 1] + r[off_lu(TANKS) + 2] + r[off_lu(ARTILLERY) + 1] + r[off_lu(INFANTRY) + 1];
            }
            if (n > a.bucket_land) nxt = Q_LAND_HEAVY;
          } else if (nxt == CAAA_PH_MOVE_TRANSPORTS) {
            if (g->tot_sea[cur][0] + g->tot_sea[cur][1] + g->tot_sea[cur][2] > a.bucket_sea) nxt = Q_SEA_HEAVY;
          } else if (nxt == Q_EOT) {
            if (g->money[cur] > a.bucket_money) nxt = Q_EOT_HEAVY;
          }
        }
      }
      unsigned mout = __ballot_sync(FULL, nxt >= 0);
      __syncwarp();
      while (mout) {
        const int j = __ffs(mout) - 1;
        mout &= mout - 1;
        const unsigned ij = __shfl_sync(FULL, idx, j);
        uint2* dst = pool + (size_t)ij * POOL_DW;
        const uint2* src = reinterpret_cast<const uint2*>(warp_games + j);
#pragma unroll
        for (int k = 0; k < KD; k++)
          if (k * 32 + lane < POOL_DW) __stcg(dst + k * 32 + lane, src[k * 32 + lane]);
      }
      fence_gpu();
      __syncwarp();
      {
        const unsigned peers = __match_any_sync(FULL, nxt);
        if (nxt >= 0) {
          atomicOr(&GS->bm[nxt][idx >> 5], 1u << (idx & 31u));
          if (lane == __ffs(peers) - 1) atomicAdd(&GS->cnt[nxt], __popc(peers));
        }
      }
      const unsigned nret = __popc(__ballot_sync(FULL, retired));
      if (lane == 0 && nret) atomicAdd(&GS->retired, (int)nret);
      CAAA_LAP(4);  // copy out + publish
    }
  }
#undef CAAA_LAP
  if (DBG && lane == 0) {
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
