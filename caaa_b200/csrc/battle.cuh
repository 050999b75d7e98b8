// battle.cuh -- isolated combat resolution (BASELINE config 3): resolve_sea_battles + resolve_land_battles
// (reference src/engine.cpp:2312, 3055) on a batch of independent states.
//
// Included by caaa_gpu.cu.  One thread per battle -- but a battle lasts anything from one volley to more than a
// hundred rounds, sea and land battles are different code, and every battle needs its 670-byte record converted to
// and from the working layout.  Two kernels, in the order they were written:
//
//  * battle_kernel (the "refill" kernel, CAAA_BATTLE_KERNEL=refill): a warp owns a run of states; its lanes keep a
//    battle from import to export, the battle phases yield as soon as a quarter of their lanes have finished, the
//    finished lanes export, take the next states of the run and the warp re-enters the phase with stragglers and
//    newcomers together.  41-62 M battles/s: lanes in a sea battle, in a land battle and in the layout conversion
//    never run together, so most instructions issue for 2-8 lanes.
//  * battle_pool_kernel (shipped): the working sets of an SM live in a CTA-wide pool and are regrouped by what they
//    are doing -- see the comment above the kernel.  111 M battles/s.
//
// Memory path of both: states are 670-byte records at the ABI.  Records enter and leave through small staging
// buffers in shared memory with warp-wide coalesced accesses (128-bit loads of the aligned window around a run of
// records; 16-bit stores, a record being only 2-byte aligned); the byte-wise layout conversion (import_state /
// export_state) and the branchy combat code run out of shared memory.  HBM sees the algorithmic bytes: 670 in + 670
// out per battle (measured 1.31 KB of DRAM traffic per battle).
#pragma once

namespace caaa {

// LANES of a warp's 32 lanes own a battle (all 32 move the records): the shared memory of an SM holds ~230 working
// sets whatever the split, so fewer lanes per warp means more warps -- more independent instruction streams to hide
// the latency of the serial, shared-memory-bound combat code behind -- and fewer battles waiting on a warp's slowest.
template <int LANES>
struct BattleCfg {
  static constexpr int STAGE_STATES = LANES / 4;                                  // records per staging pass
  static constexpr int STAGE_BYTES = STAGE_STATES * CAAA_STATE_BYTES;
  static constexpr int WARP_BYTES = LANES * (int)sizeof(Game) + (STAGE_BYTES + 15) / 16 * 16;
  static constexpr int WARPS_FIT = (232448 - TABLE_BYTES) / WARP_BYTES;
  static constexpr int WARPS = WARPS_FIT > 32 ? 32 : WARPS_FIT;
  static constexpr int SMEM = TABLE_BYTES + WARPS * WARP_BYTES;
};

template <int LANES>
__global__ void __launch_bounds__(BattleCfg<LANES>::WARPS * 32, 1) battle_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                                 unsigned long long seed, unsigned long long first_game_id,
                                                                                 uint8_t* out_states, int32_t* out_draws,
                                                                                 CaaaBatchSummary* summary, unsigned int* work_counter,
                                                                                 int run, int yield_div) {
  constexpr int STAGE_STATES = BattleCfg<LANES>::STAGE_STATES;
  const DevTables* T = stage_tables(tables);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* const wbase = smem_raw + TABLE_BYTES + warp * BattleCfg<LANES>::WARP_BYTES;
  Game* const g = reinterpret_cast<Game*>(wbase) + (lane < LANES ? lane : 0);  // this lane's working set (odd word stride: no bank conflicts)
  uint8_t* const raw = wbase + LANES * sizeof(Game);                     // staging: a few records in the reference layout
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  unsigned long long draws = 0, flagged = 0, done = 0;
  // 0 = no battle, 1 = sea battles pending / running, 2 = land battles, 3 = finished, waiting for export; 4 = a lane without a battle of its own
  int stage = lane < LANES ? 0 : 4;
  int idx = -1;    // the state this lane holds
  int next = 0, run_end = 0;  // this warp's run of states [next, run_end)
  bool exhausted = false;

  for (;;) {
    // ---- export the finished battles, STAGE_STATES at a time ----
    unsigned fin = __ballot_sync(FULL, stage == 3);
    while (fin) {
      unsigned grp = 0;
      int k = 0;
      for (unsigned f = fin; f && k < STAGE_STATES; f &= f - 1, k++) grp |= f & (0u - f);
      fin &= ~grp;
      const bool mine = (grp >> lane) & 1u;
      const int slot = __popc(grp & ((1u << lane) - 1u));
      if (mine) {
        if (out_states) export_state(g, raw + slot * CAAA_STATE_BYTES);
        if (out_draws) out_draws[idx] = (int32_t)g->draws;
        draws += g->draws;
        flagged += g->status != 0;
        done += 1;
        stage = 0;
      }
      __syncwarp();
      if (out_states) {
        for (unsigned m = grp; m; m &= m - 1) {  // one record at a time: 335 half-words, coalesced
          const int j = __ffs(m) - 1;
          const int sj = __popc(grp & ((1u << j) - 1u));
          const int ij = __shfl_sync(FULL, idx, j);
          const uint16_t* s16 = reinterpret_cast<const uint16_t*>(raw + sj * CAAA_STATE_BYTES);
          uint16_t* d16 = reinterpret_cast<uint16_t*>(out_states + (size_t)ij * CAAA_STATE_BYTES);
#pragma unroll 1
          for (int h = lane; h < CAAA_STATE_BYTES / 2; h += 32) d16[h] = s16[h];
        }
        __syncwarp();
      }
      if (mine) idx = -1;
    }
    // ---- refill idle lanes with the next states of this warp's run ----
    unsigned idle = __ballot_sync(FULL, stage == 0);
    while (idle && !exhausted) {
      if (next >= run_end) {  // take another run from the device-wide counter
        unsigned r0 = 0;
        if (lane == 0) r0 = atomicAdd(work_counter, (unsigned)run);  // `run` consecutive states per grab
        r0 = __shfl_sync(FULL, r0, 0);
        if (r0 >= (unsigned)n) {
          exhausted = true;
          break;
        }
        next = (int)r0;
        run_end = (int)r0 + run < n ? (int)r0 + run : n;
      }
      int take = __popc(idle);
      if (take > STAGE_STATES) take = STAGE_STATES;
      if (take > run_end - next) take = run_end - next;
      // `take` consecutive records -> staging (coalesced half-word loads; records are 2-byte aligned)
      {
        const uint16_t* s16 = reinterpret_cast<const uint16_t*>(states + (size_t)next * CAAA_STATE_BYTES);
        uint16_t* d16 = reinterpret_cast<uint16_t*>(raw);
        const int halves = take * (CAAA_STATE_BYTES / 2);
#pragma unroll 4
        for (int h = lane; h < halves; h += 32) d16[h] = __ldg(s16 + h);
      }
      __syncwarp();
      const int rank = __popc(idle & ((1u << lane) - 1u));
      const bool mine = ((idle >> lane) & 1u) && rank < take;
      if (mine) {
        idx = next + rank;
        import_state(g, raw + rank * CAAA_STATE_BYTES);
        begin_playout(g, cx, first_game_id + (unsigned long long)idx);
        stage = 1;
      }
      __syncwarp();
      next += take;
      idle &= ~__ballot_sync(FULL, mine);
    }
    // ---- run: sea battles, then land battles; each phase yields once a quarter of its lanes are done ----
    const unsigned msea = __ballot_sync(FULL, stage == 1);
    if (msea) {
      if (stage == 1) {
        const int y = exhausted && next >= run_end ? 0 : (__popc(msea) + yield_div - 1) / yield_div;  // nothing left to refill with: run to the end
#ifdef CAAA_BATTLE_PROBE  // measurement build: layout conversion and cache rebuild only, no combat
        (void)y;
        stage = 2;
#else
        if (resolve_sea_battles<false>(g, cx, msea, y) != PH_YIELDED) stage = 2;
#endif
      }
      __syncwarp();
    }
    const unsigned mland = __ballot_sync(FULL, stage == 2);
    if (mland) {
      if (stage == 2) {
        const int y = exhausted && next >= run_end ? 0 : (__popc(mland) + yield_div - 1) / yield_div;
#ifdef CAAA_BATTLE_PROBE
        (void)y;
        stage = 3;
#else
        if (resolve_land_battles<false>(g, cx, mland, y) != PH_YIELDED) stage = 3;
#endif
      }
      __syncwarp();
    }
    if (__ballot_sync(FULL, stage != 0 && stage != 4) == 0 && exhausted) break;
  }
  if (summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      draws += __shfl_down_sync(FULL, draws, o);
      flagged += __shfl_down_sync(FULL, flagged, o);
      done += __shfl_down_sync(FULL, done, o);
    }
    if (lane == 0 && done) {
      atomicAdd((unsigned long long*)&summary->playouts, done);
      atomicAdd((unsigned long long*)&summary->draws, draws);
      atomicAdd((unsigned long long*)&summary->flagged, flagged);
    }
  }
}

// ---- pool kernel: the battles of an SM regrouped by what they are doing ----
//
// In the refill kernel a lane keeps its battle from import to export, so the lanes of a warp that are in a sea battle,
// in a land battle and in the layout conversion never run together: measured (tools/battle_probe.py), 40 % of the time
// was conversion with 4-8 of 32 lanes active and the combat lanes of a warp were split between two code paths.  Here
// the working sets live in a CTA-wide pool in shared memory and warps own none of them: per stage (free, sea battles,
// land battles, export) the CTA keeps a bitmap of the slots waiting for it; a warp claims up to 32 slots of ONE stage
// (bit r of a bitmap word belongs to lane r, so slot % 32 == lane and the odd word stride keeps the claimed set free
// of bank conflicts), runs that stage on them in lockstep and moves each slot's bit to its next stage.  NLOAD loader
// warps do the layout work at full width -- a coalesced 128-bit read of up to LB consecutive records into their
// staging buffer, import + cache rebuild, and export + coalesced write-back --, NCOMBAT warps the battles; a battle
// phase still yields once 1/yield_div of its lanes are done, and the stragglers meet other stragglers in the queue.
constexpr int BQ_FREE = 0, BQ_SEA = 1, BQ_LAND = 2, BQ_EXPORT = 3, BNQ = 4;
constexpr int BP_MAX_WORDS = 8;

struct BattlePoolCtrl {
  uint32_t bm[BNQ][BP_MAX_WORDS];  // slots waiting per stage: bit r of word w = slot w * 32 + r
  int retired;                     // slots that found no state left to load
  int progress;                    // claims served (watchdog)
  int aborted;
  int pad_;
};

template <int NLOAD, int NCOMBAT, int LB>
struct BattlePoolCfg {
  static constexpr int WARPS = NLOAD + NCOMBAT;
  static constexpr int STAGE_BYTES = (LB * CAAA_STATE_BYTES + 16 + 15) / 16 * 16;  // LB records + alignment slack
  static constexpr int CTRL_BYTES = (int)((sizeof(BattlePoolCtrl) + 15) / 16 * 16);
  static constexpr int SLOTS_FIT = (232448 - TABLE_BYTES - NLOAD * STAGE_BYTES - CTRL_BYTES - 16) / (int)sizeof(Game);
  static constexpr int SLOTS = SLOTS_FIT > BP_MAX_WORDS * 32 ? BP_MAX_WORDS * 32 : SLOTS_FIT;
  static constexpr int WORDS = (SLOTS + 31) / 32;  // the last word may be partial
  static constexpr int POOL_BYTES = (SLOTS * (int)sizeof(Game) + 15) / 16 * 16;  // the staging buffers are 16-byte aligned
  static constexpr int SMEM = TABLE_BYTES + POOL_BYTES + NLOAD * STAGE_BYTES + CTRL_BYTES;
  static_assert(SMEM <= 232448, "battle pool: shared memory budget");
  static_assert(WORDS >= 1 && LB >= 1 && LB <= 32, "battle pool: configuration");
};

template <int NLOAD, int NCOMBAT, int LB>
__global__ void __launch_bounds__((NLOAD + NCOMBAT) * 32, 1) battle_pool_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                                unsigned long long seed, unsigned long long first_game_id,
                                                                                uint8_t* out_states, int32_t* out_draws,
                                                                                CaaaBatchSummary* summary, unsigned int* work_counter,
                                                                                int yield_div, unsigned long long* dbg) {
  using Cfg = BattlePoolCfg<NLOAD, NCOMBAT, LB>;
  constexpr int WORDS = Cfg::WORDS;
  const DevTables* T = stage_tables(tables);
  Game* const slots = reinterpret_cast<Game*>(smem_raw + TABLE_BYTES);
  uint8_t* const stage_base = smem_raw + TABLE_BYTES + Cfg::POOL_BYTES;
  BattlePoolCtrl* S = reinterpret_cast<BattlePoolCtrl*>(stage_base + NLOAD * Cfg::STAGE_BYTES);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x < BNQ * BP_MAX_WORDS) {
    const int q = threadIdx.x / BP_MAX_WORDS, w = threadIdx.x % BP_MAX_WORDS;
    const int nbits = Cfg::SLOTS - w * 32;
    S->bm[q][w] = q != BQ_FREE || nbits <= 0 ? 0u : nbits >= 32 ? FULL : (1u << nbits) - 1u;
  }
  if (threadIdx.x == 0) {
    S->retired = 0;
    S->progress = 0;
    S->aborted = 0;
  }
  __syncthreads();
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  const bool loader = warp < NLOAD;
  uint8_t* const raw = stage_base + (loader ? warp : 0) * Cfg::STAGE_BYTES;
  const int q0 = loader ? BQ_EXPORT : BQ_SEA, q1 = loader ? BQ_FREE : BQ_LAND;
  volatile uint32_t* vbm = &S->bm[0][0];
  volatile int* vretired = &S->retired;
  volatile int* vprogress = &S->progress;
  volatile int* vaborted = &S->aborted;
  unsigned long long draws = 0, flagged = 0, done = 0;
  int last_kind = q0;
  int seen_progress = -1;
  long long last_change = clock64();

  for (;;) {
    // ---- which of this warp's two stages has more claimable slots (one per bank residue)?  Ties and near-ties stay
    // with the stage served last: its code is warm ----
    uint32_t any0 = 0, any1 = 0;
#pragma unroll
    for (int w = 0; w < WORDS; w++) {
      any0 |= vbm[q0 * BP_MAX_WORDS + w];
      any1 |= vbm[q1 * BP_MAX_WORDS + w];
    }
    const long long t_poll = dbg ? clock64() : 0;
    if ((any0 | any1) == 0) {
      if (*vretired >= Cfg::SLOTS || *vaborted) break;
      const int p = *vprogress;  // watchdog: live slots but nothing served by any warp for seconds = a lost slot
      const long long now = clock64();
      if (p != seen_progress) {
        seen_progress = p;
        last_change = now;
      } else if (now - last_change > 8000000000LL) {
        if (lane == 0) S->aborted = 1;
        break;
      }
      __nanosleep(100);
      if (dbg && lane == 0) atomicAdd(&dbg[(loader ? 4 : 5) * 4], (unsigned long long)(clock64() - t_poll));  // idle
      continue;
    }
    const int a0 = __popc(any0) + (last_kind == q0 ? 4 : 0), a1 = __popc(any1) + (last_kind == q1 ? 4 : 0);
    const int kind = any1 == 0 ? q0 : any0 == 0 ? q1 : a0 >= a1 ? q0 : q1;
    const int room_max = loader ? LB : 32;
    // ---- claim: one atomic per bitmap word ----
    unsigned have = 0;
    int my_w = -1;
    {
      int w = warp % WORDS;  // warps start at different words
#pragma unroll 1
      for (int i = 0; i < WORDS; i++, w = (w + 1 == WORDS ? 0 : w + 1)) {
        uint32_t want = vbm[kind * BP_MAX_WORDS + w] & ~have;
        if (want == 0) continue;
        const int room = room_max - __popc(have);
        if (__popc(want) > room) want &= (1u << __fns(want, 0, room + 1)) - 1u;  // keep the lowest `room` bits
        uint32_t old = 0;
        if (lane == 0) old = atomicAnd(&S->bm[kind][w], ~want);
        const uint32_t got = __shfl_sync(FULL, old, 0) & want;
        if ((got >> lane) & 1u) my_w = w;
        have |= got;
        if (__popc(have) >= room_max) break;
      }
    }
    if (have == 0) continue;  // another warp was faster
    __threadfence_block();
    if (lane == 0) S->progress = S->progress + 1;
    last_kind = kind;
    const bool runs = my_w >= 0;
    Game* const g = slots + (runs ? my_w * 32 + lane : 0);
    const int nh = __popc(have);
    const int rank = __popc(have & ((1u << lane) - 1u));
    int nxt = -1;
    bool retire = false;

    if (kind == BQ_FREE) {
      // ---- load: the next nh states of the batch, one coalesced 128-bit pass over the aligned window around them ----
      unsigned r0 = 0;
      if (lane == 0) r0 = atomicAdd(work_counter, (unsigned)nh);
      r0 = __shfl_sync(FULL, r0, 0);
      const int avail = r0 >= (unsigned)n ? 0 : ((unsigned)n - r0 < (unsigned)nh ? (int)((unsigned)n - r0) : nh);
      int off = 0;
      if (avail > 0) {
        const uint8_t* src = states + (size_t)r0 * CAAA_STATE_BYTES;
        off = (int)(reinterpret_cast<uintptr_t>(src) & 15u);
        const uint4* a0p = reinterpret_cast<const uint4*>(src - off);
        const int nvec = (off + avail * CAAA_STATE_BYTES + 15) >> 4;
        uint4* d = reinterpret_cast<uint4*>(raw);
#pragma unroll 4
        for (int v = lane; v < nvec; v += 32) d[v] = __ldg(a0p + v);
      }
      __syncwarp();
      if (runs) {
        if (rank < avail) {
          import_state(g, raw + off + rank * CAAA_STATE_BYTES);
          begin_playout(g, cx, first_game_id + (unsigned long long)r0 + (unsigned)rank);
          nxt = (g->work & W_SEA_BATTLE) ? BQ_SEA : (g->work & W_LAND_BATTLE) ? BQ_LAND : BQ_EXPORT;
        } else {
          retire = true;
        }
      }
      __syncwarp();
    } else if (kind == BQ_EXPORT) {
      // ---- store: results back in the reference layout through the staging buffer, one record at a time ----
      unsigned idx = 0;
      if (runs) {
        idx = g->game_lo - (uint32_t)first_game_id;  // n < 2^31: the low words suffice
        if (out_states) export_state(g, raw + rank * CAAA_STATE_BYTES);
        if (out_draws) out_draws[idx] = (int32_t)g->draws;
        draws += g->draws;
        flagged += g->status != 0;
        done += 1;
        nxt = BQ_FREE;
      }
      __syncwarp();
      if (out_states) {
        for (unsigned m = have; m; m &= m - 1) {
          const int j = __ffs(m) - 1;
          const int sj = __popc(have & ((1u << j) - 1u));
          const unsigned ij = __shfl_sync(FULL, idx, j);
          const uint16_t* s16 = reinterpret_cast<const uint16_t*>(raw + sj * CAAA_STATE_BYTES);
          uint16_t* d16 = reinterpret_cast<uint16_t*>(out_states + (size_t)ij * CAAA_STATE_BYTES);
#pragma unroll 1
          for (int h = lane; h < CAAA_STATE_BYTES / 2; h += 32) d16[h] = s16[h];
        }
        __syncwarp();
      }
    } else {
      // ---- combat: one stage on the claimed battles in lockstep ----
      const int yield = nh >= 4 ? (nh + yield_div - 1) / yield_div : 0;
      if (runs) {
        if (kind == BQ_SEA) {
          const int res = resolve_sea_battles<false>(g, cx, have, yield);
          nxt = res == PH_YIELDED ? BQ_SEA : (g->work & W_LAND_BATTLE) ? BQ_LAND : BQ_EXPORT;
        } else {
          const int res = resolve_land_battles<false>(g, cx, have, yield);
          nxt = res == PH_YIELDED ? BQ_LAND : BQ_EXPORT;
        }
      }
      __syncwarp();
    }
    // ---- route every slot to its next stage: one bit in that stage's bitmap ----
    __threadfence_block();
    if (nxt >= 0) atomicOr(&S->bm[nxt][my_w], 1u << lane);
    const unsigned requeued = dbg ? __ballot_sync(FULL, nxt == kind) : 0u;
    if (dbg && lane == 0) {  // CAAA_BATTLE_DEBUG: cycles / slots / claims per stage
      atomicAdd(&dbg[kind * 4], (unsigned long long)(clock64() - t_poll));
      atomicAdd(&dbg[kind * 4 + 1], (unsigned long long)nh);
      atomicAdd(&dbg[kind * 4 + 2], 1ULL);
      atomicAdd(&dbg[kind * 4 + 3], (unsigned long long)__popc(requeued));  // stragglers back in the same queue
    }
    const unsigned nret = __popc(__ballot_sync(FULL, retire));
    if (lane == 0 && nret) atomicAdd(&S->retired, (int)nret);
  }

  if (summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      draws += __shfl_down_sync(FULL, draws, o);
      flagged += __shfl_down_sync(FULL, flagged, o);
      done += __shfl_down_sync(FULL, done, o);
    }
    if (lane == 0 && done) {
      atomicAdd((unsigned long long*)&summary->playouts, done);
      atomicAdd((unsigned long long*)&summary->draws, draws);
      atomicAdd((unsigned long long*)&summary->flagged, flagged);
    }
  }
}


}  // namespace caaa
