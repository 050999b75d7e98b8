// battle.cuh -- isolated combat resolution (BASELINE config 3): resolve_sea_battles + resolve_land_battles
// (reference src/engine.cpp:2312, 3055) on a batch of independent states.
//
// Included by caaa_gpu.cu.  One thread per battle -- but a battle lasts anything from one volley to more than a
// hundred rounds, and a warp that resolves 32 battles side by side runs as long as its longest one: measured, the
// round-1 kernel and a tile-per-warp rewrite with bulk asynchronous copies both sat at 31 M battles/s with ~2 of
// 32 lanes busy, whatever the memory path.  So the lanes are REFILLED: a warp owns a contiguous run of states and
// a cursor into it; the battle phases yield (engine2.cuh keeps a resumable cursor per game) as soon as a quarter
// of their lanes have finished, the finished lanes export their state, take the next states of the run, and the
// warp re-enters the phase with the stragglers and the newcomers together.
//
// Memory path: states are 670-byte records at the ABI (2-byte aligned).  A refill moves up to 8 consecutive
// records with warp-wide coalesced 16-bit accesses through a small staging buffer in shared memory; the byte-wise
// layout conversion (import_state / export_state) and the branchy combat code run out of shared memory.  HBM sees
// exactly the algorithmic bytes: 670 in + 670 out per battle.
#pragma once

namespace caaa {

constexpr int BATTLE_WARPS = 6;
constexpr int STAGE_STATES = 8;                                       // records per staging pass
constexpr int STAGE_BYTES = STAGE_STATES * CAAA_STATE_BYTES;          // 5 360
constexpr int BATTLE_WARP_BYTES = 32 * (int)sizeof(Game) + (STAGE_BYTES + 15) / 16 * 16;
constexpr int BATTLE_SMEM = TABLE_BYTES + BATTLE_WARPS * BATTLE_WARP_BYTES;
static_assert(BATTLE_SMEM <= 232448, "battle kernel: shared memory budget");

__global__ void __launch_bounds__(BATTLE_WARPS * 32, 1) battle_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                      unsigned long long seed, unsigned long long first_game_id,
                                                                      uint8_t* out_states, int32_t* out_draws, CaaaBatchSummary* summary,
                                                                      unsigned int* work_counter, int run, int yield_div) {
  const DevTables* T = stage_tables(tables);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* const wbase = smem_raw + TABLE_BYTES + warp * BATTLE_WARP_BYTES;
  Game* const g = reinterpret_cast<Game*>(wbase) + lane;                 // this lane's working set (odd word stride: no bank conflicts)
  uint8_t* const raw = wbase + 32 * sizeof(Game);                        // staging: up to 8 records in the reference layout
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  unsigned long long draws = 0, flagged = 0, done = 0;
  int stage = 0;   // 0 = no battle, 1 = sea battles pending / running, 2 = land battles, 3 = finished, waiting for export
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
        if (resolve_sea_battles<false>(g, cx, msea, y) != PH_YIELDED) stage = 2;
      }
      __syncwarp();
    }
    const unsigned mland = __ballot_sync(FULL, stage == 2);
    if (mland) {
      if (stage == 2) {
        const int y = exhausted && next >= run_end ? 0 : (__popc(mland) + yield_div - 1) / yield_div;
        if (resolve_land_battles<false>(g, cx, mland, y) != PH_YIELDED) stage = 3;
      }
      __syncwarp();
    }
    if (__ballot_sync(FULL, stage != 0) == 0 && exhausted) break;
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
