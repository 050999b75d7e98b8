// caaa_gpu.cu -- sm_100a kernels and the C-ABI of libcaaa_gpu.so (include/caaa_gpu.h).
//
// Execution model.  One thread plays one game; a game's working set (`Game`, 932 bytes) lives in
// shared memory while a warp works on it.  Large batches run in rollout_kernel_stream (stream.cuh): persistent, one CTA
// per SM, games regrouped by phase through an L2-resident pool so that a warp's lanes all have work in the phase it
// runs and a CTA's warps share one phase's instructions.  Small batches (<= 12 288 playouts) run in rollout_kernel
// below, which keeps every game in its lane for life and walks the turn pipeline (no pool traffic, lowest latency).
// Finished games are replaced from a global ticket counter by a warp-cooperative copy of the pre-imported start
// state, which absorbs the long tail of game lengths.  No tensor cores: the work is branchy uint8 arithmetic.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/caaa_gpu.h"
#include "engine2.cuh"
#include "host_tables.h"

namespace caaa {
using namespace e2;

constexpr int TABLE_BYTES = (int)((sizeof(DevTables) + 15) / 16 * 16);
constexpr int GAME_WORDS = (int)(sizeof(Game) / 4);
constexpr int ROLLOUT_WARPS = 7;                     // lockstep kernel: 7 x 32 games x 932 B + tables = 210 KB of shared memory
constexpr int ROLLOUT_THREADS = ROLLOUT_WARPS * 32;  // one game per thread
constexpr int SMALL_THREADS = 128;                   // non-persistent kernels
constexpr unsigned FULL = 0xffffffffu;

// smem_raw (dynamic shared memory, tables first) is declared in game.cuh

__device__ __forceinline__ const DevTables* stage_tables(const DevTables* __restrict__ src) {
  // read-only tables: one copy per CTA in shared memory (divergent indices would serialise in the constant cache)
  const uint16_t* s = reinterpret_cast<const uint16_t*>(src);
  uint16_t* d = reinterpret_cast<uint16_t*>(smem_raw);
  for (int i = threadIdx.x; i < (int)(sizeof(DevTables) / 2); i += blockDim.x) d[i] = s[i];
  __syncthreads();
  return reinterpret_cast<const DevTables*>(smem_raw);
}
__device__ __forceinline__ Game* my_game() { return reinterpret_cast<Game*>(smem_raw + TABLE_BYTES) + threadIdx.x; }

struct RolloutArgs {
  const DevTables* tables;
  const uint32_t* templates;  // n_states pre-imported Games (prepare_kernel)
  const double* scores0;      // their first score
  unsigned long long n_games;
  unsigned int rollouts_per_state;
  unsigned long long seed, first_game_id;
  double* results;
  CaaaRolloutStats* stats;    // optional
  CaaaBatchSummary* summary;  // optional
  unsigned long long* ticket;
  int yield_lanes;   // a phase yields once this percentage of its batch is done; stragglers are requeued (0 = never)
  int scan_rounds;   // bitmap words / 32 a claim may look at
  int yield_battles; // like yield_lanes, for the two battle phases only
  // SM-resident kernel (resident.cuh)
  int yield_min;        // batches smaller than this never yield
  int stick;            // bonus (in games) for the kind a warp served last
  int min_batch;        // a warp waits up to min_batch_polls polls for a batch of at least this many games
  int min_batch_polls;
  unsigned idle_ns;     // sleep of an idle warp between two looks at the queues
  int* abort_flag;      // set when the watchdog gave up (global memory, optional)
  int bucket_land, bucket_sea, bucket_money;  // pool kernel: thresholds of the work-size buckets (a huge value switches one off)
  int patience;         // pool kernel: empty claims in a row before a CTA leaves its phase (0 = round 1's queue-length rule)
};

// Import every start / leaf state once: reference layout -> Game with all derived caches, ready to be
// copied into a lane's slot (what random_play_until_terminal does before its loop, src/engine.cpp:4185-4203).
__global__ void __launch_bounds__(SMALL_THREADS) prepare_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                uint32_t* templates, double* scores0) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Game* g = my_game();
  const RunCtx cx{T, 0u, 0u};
  import_state(g, states + (size_t)i * CAAA_STATE_BYTES);
  begin_playout(g, cx, 0);
  scores0[i] = evaluate_state(g, cx);
  const uint32_t* w = reinterpret_cast<const uint32_t*>(g);
  for (int k = 0; k < GAME_WORDS; k++) templates[(size_t)i * GAME_WORDS + k] = w[k];
}

struct LaneTotals {
  unsigned long long playouts = 0, plies = 0, decisions = 0, draws = 0, flagged = 0;
  double sum = 0.0;
};

template <bool STATS>
__device__ __noinline__ void finish_playout(Game* g, double score, unsigned long long gi, const RolloutArgs& a, LaneTotals& tot) {
  const double r = (g->player % 2 == 1) ? 1.0 - score : score;  // reference src/engine.cpp:4238-4241
  a.results[gi] = r;
  const int decisions = 100000 - g->answers_remaining;
  if (STATS) {
    CaaaRolloutStats s;
    s.result = r;
    s.turns = g->turns;
    s.decisions = decisions;
    s.draws = (int32_t)g->draws;
    s.final_player = g->player;
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

__device__ __forceinline__ void cta_sync() { asm volatile("bar.sync 1, %0;" ::"r"(ROLLOUT_THREADS) : "memory"); }
__device__ __forceinline__ bool cta_or(bool pred) {
  int out;
  asm volatile(
      "{\n\t.reg .pred p, q;\n\tsetp.ne.u32 q, %1, 0;\n\tbarrier.red.or.pred p, 1, %2, q;\n\tselp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(out)
      : "r"((int)pred), "r"(ROLLOUT_THREADS)
      : "memory");
  return out != 0;
}

// The fixed pipeline of one player-turn, reference src/engine.cpp:4214-4235.  A phase is only entered by
// the lanes whose game has work for it (`mw`, voted among the lanes that play a turn this iteration); inside,
// those lanes stay in lockstep (engine2.cuh).  With SYNC a CTA barrier after every phase keeps all warps of
// the SM inside the same phase function (they then share its instructions in the SM's instruction caches).
#define CAAA_PHASE(ph, fn, ...)                                                    \
  do {                                                                             \
    if (run) {                                                                     \
      const unsigned mw = __ballot_sync(mrun, (g->work & W(ph)) != 0);             \
      if (g->work & W(ph)) fn<false>(g, cx, mw, ##__VA_ARGS__);                    \
    }                                                                              \
    if (SYNC) cta_sync();                                                          \
  } while (0)

template <bool STATS, bool SYNC>
__global__ void __launch_bounds__(ROLLOUT_THREADS, 1) rollout_kernel(const RolloutArgs a) {
  const DevTables* T = stage_tables(a.tables);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  Game* const warp_games = reinterpret_cast<Game*>(smem_raw + TABLE_BYTES) + warp * 32;
  Game* const g = warp_games + lane;
  const RunCtx cx{T, (uint32_t)a.seed, (uint32_t)(a.seed >> 32)};
  const bool one_state = a.rollouts_per_state >= a.n_games;
  bool live = false, exhausted = false;
  unsigned long long gi = 0;
  double score = 0.0;
  LaneTotals tot;
  for (;;) {
    // ---- refill idle lanes at a turn boundary: one ticket fetch per warp, then all 32 lanes copy each
    // new game's 233 words from the pre-imported state (coalesced, converged) ----
    const unsigned need = __ballot_sync(FULL, !live && !exhausted);
    if (need) {
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
            if (k * 32 + lane < GAME_WORDS) dst[k * 32 + lane] = __ldg(src + k * 32 + lane);
        }
      }
      __syncwarp();
      if ((need >> lane) & 1u) {
        gi = base + (unsigned)__popc(need & ((1u << lane) - 1u));
        if (gi < a.n_games) {
          const unsigned long long id = a.first_game_id + gi;
          g->game_lo = (uint32_t)id;
          g->game_hi = (uint32_t)(id >> 32);
          score = a.scores0[one_state ? 0ULL : gi / a.rollouts_per_state];
          live = true;
        } else {
          exhausted = true;
        }
      }
    }
    if (SYNC) {
      if (!cta_or(live)) break;
    } else {
      if (!__any_sync(FULL, live)) break;
    }
    const bool run = live && score > 0.01 && score < 0.99 && g->turns < 1000;
    if (live && !run) {
      finish_playout<STATS>(g, score, gi, a, tot);
      live = false;
    }
    const unsigned mrun = __ballot_sync(FULL, run);
    CAAA_PHASE(CAAA_PH_MOVE_FIGHTERS, move_fighter_units);
    CAAA_PHASE(CAAA_PH_MOVE_BOMBERS, move_bomber_units);
    CAAA_PHASE(CAAA_PH_STAGE_TRANSPORTS, stage_transport_units);
    CAAA_PHASE(CAAA_PH_MOVE_TANKS, move_land_unit_type, TANKS, CAAA_PH_MOVE_TANKS);
    CAAA_PHASE(CAAA_PH_MOVE_ARTILLERY, move_land_unit_type, ARTILLERY, CAAA_PH_MOVE_ARTILLERY);
    CAAA_PHASE(CAAA_PH_MOVE_INFANTRY, move_land_unit_type, INFANTRY, CAAA_PH_MOVE_INFANTRY);
    CAAA_PHASE(CAAA_PH_MOVE_TRANSPORTS, move_transport_units);
    CAAA_PHASE(CAAA_PH_MOVE_SUBS, move_subs);
    CAAA_PHASE(CAAA_PH_MOVE_SHIPS, move_destroyers_battleships);
    CAAA_PHASE(CAAA_PH_SEA_BATTLES, resolve_sea_battles);
    CAAA_PHASE(CAAA_PH_UNLOAD_TRANSPORTS, unload_transports);
    CAAA_PHASE(CAAA_PH_LAND_BATTLES, resolve_land_battles);
    CAAA_PHASE(CAAA_PH_MOVE_AA_GUNS, move_land_unit_type, AA_GUNS, CAAA_PH_MOVE_AA_GUNS);
    CAAA_PHASE(CAAA_PH_LAND_FIGHTERS, land_fighter_units);
    CAAA_PHASE(CAAA_PH_LAND_BOMBERS, land_bomber_units);
    if (run) {
      buy_units<false>(g, cx, mrun);
      crash_air_units(g, cx);
      reset_units_fully(g);
      collect_money(g, cx);
      rotate_turns(g, cx);
      g->turns++;
      score = evaluate_state(g, cx);
    }
    if (SYNC) cta_sync();
  }
  if (a.summary != nullptr) {  // warp-reduce, then one atomic per warp and counter
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
#include "stream.cuh"
#define CAAA_HAVE_STREAM_KERNEL 1
#include "resident.cuh"
namespace caaa {

__device__ __forceinline__ void play_turn(Game* g, const RunCtx cx, unsigned m) {
#pragma unroll 1
  for (int ph = 0; ph < CAAA_PH_COUNT; ph++) run_phase<false>(g, cx, ph, m);
}

__global__ void __launch_bounds__(SMALL_THREADS) advance_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                unsigned long long seed, unsigned long long first_game_id,
                                                                int n_turns, uint8_t* out_states, CaaaCacheDump* out_caches,
                                                                int32_t* out_draws) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned m = __ballot_sync(FULL, i < n);  // the lanes of this warp that hold a state walk the pipeline in lockstep
  if (i >= n) return;
  Game* g = my_game();
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  import_state(g, states + (size_t)i * CAAA_STATE_BYTES);
  begin_playout(g, cx, first_game_id + (unsigned long long)i);
#pragma unroll 1
  for (int t = 0; t < n_turns; t++) play_turn(g, cx, m);
  if (out_states) export_state(g, out_states + (size_t)i * CAAA_STATE_BYTES);
  if (out_caches) export_cache(g, T, out_caches + i);
  if (out_draws) out_draws[i] = (int32_t)g->draws;
}

}  // namespace caaa
#include "battle.cuh"
namespace caaa {

// Stop-at-decision mode (reference get_possible_actions / apply_action, src/engine.cpp:4050-4183).
__device__ __forceinline__ void run_until_decision(Game* g, const RunCtx cx) {
  // the reference loops `while (true)` here (src/engine.cpp:4068,4135): a degenerate state in which no player ever
  // reaches a decision would spin for ever; cap it, flag the state and report an empty action list
  for (int turn = 0; turn < EXPAND_TURN_CAP; turn++) {
    bool stopped = false;
    for (int ph = 0; ph <= CAAA_PH_BUY_UNITS && !stopped; ph++) stopped = run_phase<true>(g, cx, ph);
    if (stopped) return;
    for (int ph = CAAA_PH_CRASH_AIR; ph < CAAA_PH_COUNT; ph++) run_phase<true>(g, cx, ph);
  }
  g->status |= CAAA_ERR_NO_DECISION;
  g->nvm = 0;
}
__device__ __forceinline__ void begin_expansion(Game* g, const RunCtx cx, int answers, int use_selected, uint8_t action) {
  begin_common(g, cx);
  g->answers_remaining = answers;
  g->use_selected = (uint8_t)use_selected;
  g->selected_action = action;
  g->game_lo = g->game_hi = 0;
  for (int i = 0; i < CAAA_ACTIONS; i++) g->vm[i] = 0;
  g->nvm = 0;
}
__global__ void __launch_bounds__(SMALL_THREADS) actions_kernel(const DevTables* tables, const uint8_t* leaves, int n,
                                                                uint8_t* n_actions, uint8_t* actions, uint32_t* status) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Game* g = my_game();
  const RunCtx cx{T, 0u, 0u};
  import_state(g, leaves + (size_t)i * CAAA_STATE_BYTES);
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
  Game* g = my_game();
  const RunCtx cx{T, 0u, 0u};
  import_state(g, leaves + (size_t)i * CAAA_STATE_BYTES);
  begin_expansion(g, cx, 1, 1, actions[(size_t)i * CAAA_ACTIONS + k]);
  run_until_decision(g, cx);
  export_state(g, children + (size_t)idx * CAAA_STATE_BYTES);
  if (status && g->status) atomicOr(&status[i], 1u << k);  // bit k: child k tripped a guard
}
__global__ void __launch_bounds__(SMALL_THREADS) evaluate_kernel(const DevTables* tables, const uint8_t* states, int n, double* scores) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Game* g = my_game();
  const RunCtx cx{T, 0u, 0u};
  import_state(g, states + (size_t)i * CAAA_STATE_BYTES);
  begin_common(g, cx);
  scores[i] = evaluate_state(g, cx);
}

// ---- AoS <-> SoA (north_star: "structure-of-arrays batch in HBM with coalesced 128-bit loads and stores") ----
// SoA form of a batch of n states: 42 chunks of 16 bytes per state (670 bytes + 2 zero pad bytes), CHUNK-MAJOR:
// chunk c of state i lives at soa[(c * n_pad + i) * 16], n_pad = n rounded up to 32.  A warp that reads chunk c of 32
// consecutive states issues one LDG.128 per lane = 512 contiguous bytes.  One CTA converts a tile of 32 states through
// shared memory so that both sides are coalesced.  (Where this layout pays and where it does not: DESIGN.md section 3.)
constexpr int SOA_CHUNKS = (CAAA_STATE_BYTES + 15) / 16;  // 42
__global__ void __launch_bounds__(128) pack_kernel(const uint8_t* aos, int n, int n_pad, uint4* soa) {
  __shared__ __align__(16) uint8_t tile[32 * SOA_CHUNKS * 16];
  const int base = blockIdx.x * 32, cnt = n - base < 32 ? n - base : 32;
  for (int i = threadIdx.x; i < 32 * SOA_CHUNKS * 4; i += blockDim.x) reinterpret_cast<uint32_t*>(tile)[i] = 0u;
  __syncthreads();
  const uint16_t* s16 = reinterpret_cast<const uint16_t*>(aos + (size_t)base * CAAA_STATE_BYTES);  // records are 2-byte aligned
  for (int h = threadIdx.x; h < cnt * (CAAA_STATE_BYTES / 2); h += blockDim.x) {
    const int g = h / (CAAA_STATE_BYTES / 2), o = h % (CAAA_STATE_BYTES / 2);
    reinterpret_cast<uint16_t*>(tile + g * SOA_CHUNKS * 16)[o] = s16[h];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < SOA_CHUNKS * 32; k += blockDim.x) {
    const int c = k / 32, g = k % 32;
    if (base + g < n_pad) soa[(size_t)c * n_pad + base + g] = reinterpret_cast<const uint4*>(tile + g * SOA_CHUNKS * 16)[c];
  }
}
__global__ void __launch_bounds__(128) unpack_kernel(const uint4* soa, int n, int n_pad, uint8_t* aos) {
  __shared__ __align__(16) uint8_t tile[32 * SOA_CHUNKS * 16];
  const int base = blockIdx.x * 32, cnt = n - base < 32 ? n - base : 32;
  for (int k = threadIdx.x; k < SOA_CHUNKS * 32; k += blockDim.x) {
    const int c = k / 32, g = k % 32;
    if (g < cnt) reinterpret_cast<uint4*>(tile + g * SOA_CHUNKS * 16)[c] = soa[(size_t)c * n_pad + base + g];
  }
  __syncthreads();
  uint16_t* d16 = reinterpret_cast<uint16_t*>(aos + (size_t)base * CAAA_STATE_BYTES);
  for (int h = threadIdx.x; h < cnt * (CAAA_STATE_BYTES / 2); h += blockDim.x) {
    const int g = h / (CAAA_STATE_BYTES / 2), o = h % (CAAA_STATE_BYTES / 2);
    d16[h] = reinterpret_cast<const uint16_t*>(tile + g * SOA_CHUNKS * 16)[o];
  }
}

// Leaf batches (MCTS): the child each visited leaf rolls out, chosen on the device so that the expansion's children
// never leave it before the rollout (reference src/mcts.cpp:84-94: expand, `rand() % num_children`, rollout).
// Terminal leaves and leaves without a playable child roll out from themselves (picked = -1).
__global__ void pick_kernel(const uint8_t* leaves, int n, const double* scores, const uint8_t* n_actions, const uint32_t* status,
                            const uint64_t* pick, const uint8_t* children, uint8_t* targets, int32_t* picked) {
  const int i = blockIdx.x;
  if (i >= n) return;
  __shared__ int s_src;
  if (threadIdx.x == 0) {
    const bool terminal = scores[i] > 0.99 || scores[i] < 0.01;  // reference src/engine.cpp:4244-4248
    int k_sel = -1;
    if (!terminal && !(status[i] & 0x80000000u)) {
      int valid = 0;
      for (int k = 0; k < n_actions[i]; k++) valid += !((status[i] >> k) & 1u);  // children whose apply_action returns
      if (valid > 0) {
        int r = (int)(pick[i] % (uint64_t)valid);
        for (int k = 0; k < n_actions[i]; k++) {
          if ((status[i] >> k) & 1u) continue;
          if (r-- == 0) {
            k_sel = k;
            break;
          }
        }
      }
    }
    picked[i] = k_sel;
    s_src = k_sel;
  }
  __syncthreads();
  const int k_sel = s_src;
  const uint8_t* src = k_sel < 0 ? leaves + (size_t)i * CAAA_STATE_BYTES : children + ((size_t)i * CAAA_ACTIONS + k_sel) * CAAA_STATE_BYTES;
  uint8_t* dst = targets + (size_t)i * CAAA_STATE_BYTES;
  for (int b = threadIdx.x; b < CAAA_STATE_BYTES; b += blockDim.x) dst[b] = src[b];
}

// Per-leaf sum of the rollout results in a FIXED order (lane-strided partial sums, then a shuffle tree), so that a
// search is reproducible: one warp per leaf, results[i * per_leaf .. (i + 1) * per_leaf).  Only n doubles go back to the
// host (reference src/mcts.cpp:96-105 back-propagates one result at a time; the sum of a leaf's results is what its
// path needs).
__global__ void reduce_results_kernel(const double* results, int n, int per_leaf, double* sums) {
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (warp >= n) return;
  const double* r = results + (size_t)warp * per_leaf;
  double acc = 0.0;
  for (int j = lane; j < per_leaf; j += 32) acc += r[j];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(FULL, acc, o);
  if (lane == 0) sums[warp] = acc;
}

// ------------------------------------------------------------------------------------------------
// host side of the C-ABI
// ------------------------------------------------------------------------------------------------
constexpr size_t BATTLE_CHUNK = 131072;  // states per pipeline chunk (88 MB each way), a multiple of 32
struct BattlePipe {  // pinned staging + streams of the host-buffer battle entry point
  static constexpr int NBUF = 3;
  bool ready = false;
  void *h_in[NBUF] = {}, *h_out[NBUF] = {}, *d_in[NBUF] = {}, *d_out[NBUF] = {};
  int32_t *h_draws[NBUF] = {}, *d_draws[NBUF] = {};
  cudaStream_t stream[NBUF] = {};
  cudaEvent_t k0[NBUF] = {}, k1[NBUF] = {}, done[NBUF] = {};
};
// Scratch of ONE rollout launch.  An asynchronous launch (caaa_gpu_rollout_batch_device) owns its slot until its
// `done` event has fired; the next user of the slot waits for that event on the device, so launches issued on
// different streams overlap safely (the drain of one batch then overlaps the start of the next).
constexpr int N_LAUNCH_SLOTS = 3;
struct LaunchSlot {
  void* d_templates = nullptr;  // pre-imported start states + their first scores
  size_t templates_cap = 0;
  void* d_pool = nullptr;       // game pool + group state (+ debug counters) of the phase-regrouped kernel
  size_t pool_cap = 0;
  unsigned long long* d_ticket = nullptr;
  int* d_abort = nullptr;       // set by a kernel's watchdog
  cudaEvent_t done = nullptr;
  bool used = false;
};
// One asynchronous MCTS leaf batch: pinned staging on the host, everything else resident on the device.
constexpr int N_LEAF_BATCHES = 3;
struct LeafBatch {
  bool ready = false, in_flight = false;
  int cap = 0, n = 0;            // leaves allocated / in this batch
  size_t results_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  uint8_t *h_in = nullptr, *h_out = nullptr;   // pinned: leaves + picks in; everything collect() returns out
  uint8_t *d_in = nullptr, *d_out = nullptr;   // device mirrors of the two
  uint8_t* d_children = nullptr;               // n x 20 x 670
  uint8_t* d_targets = nullptr;                // n x 670: the states the rollouts start from
  double* d_results = nullptr;                 // n x rollouts_per_leaf
  CaaaBatchSummary* d_summary = nullptr;
};
struct Context {
  bool ready = false;
  CaaaMap map;
  BattlePipe battle;
  unsigned int* d_battle_counters = nullptr;  // work counters of the battle kernel: one per pipeline buffer + four for device launches
  unsigned next_battle_counter = 0;
  LeafBatch leaf[N_LEAF_BATCHES];
  int rollout_sms = 0;  // SMs a rollout launch may occupy (0 = all): a search leaves a few free for the next batch's expansion
  int device = -1;
  int sm_count = 0;
  DevTables host_tables;
  DevTables* d_tables = nullptr;
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
  bool sync_phases = false;  // CAAA_SYNC=1: CTA barrier after every phase (A/B knob)
  int kernel = 1;            // kernel of the most recent launch: 0 = SM-resident, 1 = phase-regrouped over a global pool, 2 = lockstep
  int small_kernel = 2;      // the kernel for batches of up to small_batch playouts
  int res_warps = 16;        // warps per CTA of the resident kernel (CAAA_WARPS)
  long long small_batch = 12288;  // batches up to this size go to the lockstep kernel (no pool hops: lower latency;
                                  // measured: 4 096 playouts 76 vs 82 ms, 24 576 playouts 100 vs 82 ms);
                                  // CAAA_ROLLOUT_KERNEL=stream sets it to 0
  int pool_mult = 2;         // pool entries per shared-memory slot (CAAA_POOL_MULT); 2 keeps the 62 MB pool L2-resident
  int lanes = 16;            // games per warp of the phase-regrouped kernel (CAAA_LANES = 16 | 20 | 24): 15 warps x 16 games per SM
  int group_size = 74;       // CTAs that share one game pool and one set of phase bitmaps (CAAA_GROUP): two groups per B200
  LaunchSlot slots[N_LAUNCH_SLOTS];  // per-launch scratch: launches on different streams may be in flight together
  int next_slot = 0;
  std::atomic_flag busy = ATOMIC_FLAG_INIT;  // the context is single-caller: concurrent entry is refused (CAAA_E_BUSY)
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
// The context serves one caller at a time (like the reference API, whose engine lives in file-scope globals).
// A second thread entering while a call is in progress is refused instead of corrupting the shared scratch.
struct EntryGuard {
  bool ok;
  EntryGuard();
  ~EntryGuard();
};
#define CAAA_ENTER()   \
  EntryGuard entry_;   \
  if (!entry_.ok) return CAAA_E_BUSY  /* caaa_gpu_last_error() keeps the message of the call in progress */
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
EntryGuard::EntryGuard() : ok(!g_ctx.busy.test_and_set(std::memory_order_acquire)) {}
EntryGuard::~EntryGuard() {
  if (ok) g_ctx.busy.clear(std::memory_order_release);
}

// Every launch slot whose work has completed: did a kernel's watchdog give up?  (Call after a synchronisation.)
static int check_watchdogs() {
  for (LaunchSlot& l : g_ctx.slots) {
    if (!l.used) continue;
    int aborted = 0;
    CUDA_TRY(cudaMemcpy(&aborted, l.d_abort, sizeof(int), cudaMemcpyDeviceToHost));
    if (aborted) return fail(CAAA_E_CUDA, "rollout kernel stopped by its watchdog: no claimable game although the batch was unfinished");
  }
  return CAAA_OK;
}
static int small_smem() { return TABLE_BYTES + SMALL_THREADS * (int)sizeof(Game); }
static int rollout_smem() { return TABLE_BYTES + ROLLOUT_THREADS * (int)sizeof(Game); }

static int launch_rollout(const void* d_states, int n_states, unsigned rollouts_per_state, unsigned long long seed,
                          unsigned long long first_game_id, double* d_results, CaaaRolloutStats* d_stats,
                          CaaaBatchSummary* d_summary, cudaStream_t stream) {
  Context& c = g_ctx;
  LaunchSlot& L = c.slots[c.next_slot];
  c.next_slot = (c.next_slot + 1) % N_LAUNCH_SLOTS;
  const unsigned long long n_games = (unsigned long long)n_states * rollouts_per_state;
  const size_t tmpl_bytes = ((size_t)n_states * sizeof(Game) + 7) / 8 * 8;
  const size_t tmpl_need = tmpl_bytes + (size_t)n_states * sizeof(double);
  const size_t pool_need = (size_t)c.sm_count * STREAM_SLOTS * (size_t)c.pool_mult * sizeof(PoolGame) + 16 +
                           (size_t)((c.sm_count + c.group_size - 1) / c.group_size) * sizeof(GroupState) + sizeof(StreamDbg) + sizeof(ResDbg);
  if (L.used) {
    if (tmpl_need > L.templates_cap || pool_need > L.pool_cap) CUDA_TRY(cudaEventSynchronize(L.done));  // buffers are about to be replaced
    else CUDA_TRY(cudaStreamWaitEvent(stream, L.done, 0));  // the previous launch that used this slot must have drained (device-side wait)
  }
  int rc = grow(&L.d_templates, &L.templates_cap, tmpl_need);
  if (rc != CAAA_OK) return rc;
  if ((rc = grow(&L.d_pool, &L.pool_cap, pool_need)) != CAAA_OK) return rc;
  uint32_t* d_tmpl = (uint32_t*)L.d_templates;
  double* d_scores0 = (double*)((uint8_t*)L.d_templates + tmpl_bytes);
  L.used = true;
  struct RecordDone {  // whatever path returns, the slot's event marks the end of this launch's work on `stream`
    LaunchSlot& l;
    cudaStream_t st;
    ~RecordDone() { cudaEventRecord(l.done, st); }
  } record_done{L, stream};
  prepare_kernel<<<(n_states + SMALL_THREADS - 1) / SMALL_THREADS, SMALL_THREADS, small_smem(), stream>>>(
      c.d_tables, (const uint8_t*)d_states, n_states, d_tmpl, d_scores0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(L.d_ticket, 0, sizeof(unsigned long long), stream));
  CUDA_TRY(cudaMemsetAsync(L.d_abort, 0, sizeof(int), stream));
  RolloutArgs a;
  a.tables = c.d_tables;
  a.templates = d_tmpl;
  a.scores0 = d_scores0;
  a.n_games = n_games;
  a.rollouts_per_state = rollouts_per_state;
  a.seed = seed;
  a.first_game_id = first_game_id;
  a.results = d_results;
  a.stats = d_stats;
  a.summary = d_summary;
  a.ticket = L.d_ticket;
  {
    const char* yl = std::getenv("CAAA_YIELD");
    a.yield_lanes = yl ? std::atoi(yl) : 0;
    const char* yb = std::getenv("CAAA_YIELD_BATTLES");
    a.yield_battles = yb ? std::atoi(yb) : 0;
    const char* sr = std::getenv("CAAA_SCAN_ROUNDS");
    a.scan_rounds = sr ? std::atoi(sr) : 4;
    if (a.scan_rounds < 1) a.scan_rounds = 1;
  }
  auto env_int = [](const char* name, int dflt) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : dflt;
  };
  a.yield_min = env_int("CAAA_YIELD_MIN", 4);
  a.stick = env_int("CAAA_STICK", 0);
  a.min_batch = env_int("CAAA_MIN_BATCH", 1);
  a.min_batch_polls = env_int("CAAA_MIN_BATCH_POLLS", 8);
  a.idle_ns = (unsigned)env_int("CAAA_IDLE_NS", 100);
  a.abort_flag = L.d_abort;
  a.patience = env_int("CAAA_PATIENCE", 6);
  a.bucket_land = env_int("CAAA_BUCKET_LAND", 6);
  a.bucket_sea = env_int("CAAA_BUCKET_SEA", 8);
  a.bucket_money = env_int("CAAA_BUCKET_MONEY", 15);
  if (a.patience < 1) a.patience = 1;
  {  // which kernel: CAAA_ROLLOUT_KERNEL = resident | stream | lockstep forces one; default = by batch size
    const char* rk = std::getenv("CAAA_ROLLOUT_KERNEL");
    const long long small = env_int("CAAA_SMALL_BATCH", (int)c.small_batch);
    if (rk != nullptr && std::strcmp(rk, "resident") == 0) c.kernel = 0;
    else if (rk != nullptr && std::strcmp(rk, "stream") == 0) c.kernel = 1;
    else if (rk != nullptr && std::strcmp(rk, "lockstep") == 0) c.kernel = 2;
    else c.kernel = (long long)n_games > small ? 1 : c.small_kernel;
  }
  if (c.kernel == 0) {
    // SM-resident kernel: spread small batches over all SMs (latency), fill every slot for large ones (regrouping)
    const unsigned long long want32 = (n_games + 31) / 32;
    const unsigned long long sms_r = (unsigned long long)(c.rollout_sms > 0 && c.rollout_sms < c.sm_count ? c.rollout_sms : c.sm_count);
    const int grid_r = (int)(want32 < sms_r ? want32 : sms_r);
    const unsigned long long per_cta = (n_games + grid_r - 1) / grid_r;
    int n_slots = (int)((per_cta + 31) / 32 * 32);
    if (n_slots > RES_SLOTS) n_slots = RES_SLOTS;
    if (n_slots < 32) n_slots = 32;
    ResDbg* d_dbg = nullptr;
    if (std::getenv("CAAA_DEBUG") != nullptr) {
      d_dbg = (ResDbg*)L.d_pool;
      CUDA_TRY(cudaMemsetAsync(d_dbg, 0, sizeof(ResDbg), stream));
    }
    int res_warps = env_int("CAAA_WARPS", c.res_warps);
    if (res_warps < 1) res_warps = 1;
    if (res_warps > 32) res_warps = 32;
    const int threads = res_warps * 32;
    if (res_warps <= 16) {
      if (d_stats) rollout_kernel_resident<true, 16><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
      else rollout_kernel_resident<false, 16><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
    } else if (res_warps <= 24) {
      if (d_stats) rollout_kernel_resident<true, 24><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
      else rollout_kernel_resident<false, 24><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
    } else {
      if (d_stats) rollout_kernel_resident<true, 32><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
      else rollout_kernel_resident<false, 32><<<grid_r, threads, RES_SMEM, stream>>>(a, n_slots, d_dbg);
    }
    CUDA_TRY(cudaGetLastError());
    if (d_dbg != nullptr) {
      ResDbg h;
      cudaStreamSynchronize(stream);
      cudaMemcpy(&h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
      const double tt = (double)(h.t[0] + h.t[1] + h.t[2] + h.t[3]);
      std::fprintf(stderr, "resident dbg: batches %llu games/batch %.2f  idle+claim %.3f phase %.3f restart %.3f route %.3f  Mcyc/warp %.1f\n",
                   h.batches, h.batches ? (double)h.items / (double)h.batches : 0.0, h.t[0] / tt, h.t[1] / tt, h.t[2] / tt, h.t[3] / tt,
                   tt / ((double)grid_r * res_warps) / 1e6);
      double kc = 0, ki = 0;
      for (int q = 0; q < NQ; q++) { kc += (double)h.kind_cycles[q]; ki += (double)h.kind_items[q]; }
      std::fprintf(stderr, "resident dbg: per kind  time share / visit share / games per batch / kcyc per batch:");
      for (int q = 0; q < NQ; q++)
        if (h.kind_items[q])
          std::fprintf(stderr, "  k%d %.3f/%.3f/%.1f/%.1f", q, h.kind_cycles[q] / kc, h.kind_items[q] / ki,
                       (double)h.kind_items[q] / (double)h.kind_batches[q], h.kind_cycles[q] / 1e3 / (double)h.kind_batches[q]);
      std::fprintf(stderr, "\n");
    }
    return CAAA_OK;
  }
  const unsigned long long want = (n_games + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;
  const unsigned long long sms = (unsigned long long)(c.rollout_sms > 0 && c.rollout_sms < c.sm_count ? c.rollout_sms : c.sm_count);
  const int grid = (int)(want < sms ? want : sms);
  if (c.kernel == 1) {
    const size_t need = (size_t)c.sm_count * STREAM_SLOTS * (size_t)c.pool_mult * sizeof(PoolGame);
    const int n_groups = (grid + c.group_size - 1) / c.group_size;
    const size_t need16 = (need + 15) / 16 * 16;
    GroupState* d_groups = (GroupState*)((uint8_t*)L.d_pool + need16);
    group_init_kernel<<<n_groups, 256, 0, stream>>>(d_groups, n_groups, STREAM_SLOTS * c.pool_mult, c.group_size, grid);
    StreamDbg* d_dbg = nullptr;
    if (std::getenv("CAAA_DEBUG") != nullptr) {
      d_dbg = (StreamDbg*)((uint8_t*)d_groups + (size_t)n_groups * sizeof(GroupState));
      CUDA_TRY(cudaMemsetAsync(d_dbg, 0, sizeof(StreamDbg), stream));
    }
#define CAAA_LAUNCH_STREAM(NL)                                                                                                       \
  do {                                                                                                                              \
    if (d_dbg)                                                                                                                      \
      rollout_kernel_stream<false, NL, true><<<grid, stream_threads, STREAM_SMEM, stream>>>(a, (uint32_t*)L.d_pool, c.pool_mult, d_groups, c.group_size, d_dbg); \
    else if (d_stats)                                                                                                               \
      rollout_kernel_stream<true, NL><<<grid, stream_threads, STREAM_SMEM, stream>>>(a, (uint32_t*)L.d_pool, c.pool_mult, d_groups, c.group_size, d_dbg);  \
    else                                                                                                                            \
      rollout_kernel_stream<false, NL><<<grid, stream_threads, STREAM_SMEM, stream>>>(a, (uint32_t*)L.d_pool, c.pool_mult, d_groups, c.group_size, d_dbg); \
  } while (0)
    const int lanes = env_int("CAAA_LANES", c.lanes) == 24 ? 24 : env_int("CAAA_LANES", c.lanes) == 20 ? 20 : 16;
    int stream_warps = env_int("CAAA_STREAM_WARPS", STREAM_SLOTS / lanes);  // A/B knob: fewer warps leave slots unused
    if (stream_warps < 1 || stream_warps > STREAM_SLOTS / lanes) stream_warps = STREAM_SLOTS / lanes;
    const int stream_threads = stream_warps * 32;
    if (lanes == 24) CAAA_LAUNCH_STREAM(24);
    else if (lanes == 20) CAAA_LAUNCH_STREAM(20);
    else CAAA_LAUNCH_STREAM(16);  // (measured with the sticky phase rule: 16 / 20 / 24 games per warp = 1.35 / 1.40 / 1.39 M playouts/s;
                                  //  8 and 32 under round 1's rule: 0.91 M and 1.23 M against 1.30 M)
    if (d_dbg != nullptr) {
      StreamDbg h;
      cudaStreamSynchronize(stream);
      cudaMemcpy(&h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
      const double tt = (double)(h.t[0] + h.t[1] + h.t[2] + h.t[3] + h.t[4]);
      std::fprintf(stderr, "stream dbg: batches %llu items/batch %.1f  idle+claim %.3f in %.3f compute %.3f restart %.3f out %.3f  Mcyc/warp %.1f\n",
                   h.batches, h.batches ? (double)h.items / (double)h.batches : 0.0, h.t[0] / tt, h.t[1] / tt, h.t[2] / tt, h.t[3] / tt,
                   h.t[4] / tt, tt / (grid * (STREAM_SLOTS / c.lanes)) / 1e6);
      std::fprintf(stderr, "stream dbg: CTA kind switches %llu = one per %.1f batches of the CTA\n", h.t[5], h.t[5] ? (double)h.batches / (double)h.t[5] : 0.0);
      std::fprintf(stderr, "stream dbg: per queue  share of phase time / share of game visits / games per batch:");
      double kc = 0, ki = 0;
      for (int q = 0; q < NQ; q++) { kc += (double)h.kind_cycles[q]; ki += (double)h.kind_items[q]; }
      for (int q = 0; q < NQ; q++)
        if (h.kind_items[q])
          std::fprintf(stderr, "  k%d %.3f/%.3f/%.1f", q, h.kind_cycles[q] / kc, h.kind_items[q] / ki,
                       (double)h.kind_items[q] / (double)(h.kind_batches[q] ? h.kind_batches[q] : 1));
      std::fprintf(stderr, "\n");
    }
  } else if (c.sync_phases) {
    if (d_stats) rollout_kernel<true, true><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
    else rollout_kernel<false, true><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
  } else {
    if (d_stats) rollout_kernel<true, false><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
    else rollout_kernel<false, false><<<grid, ROLLOUT_THREADS, rollout_smem(), stream>>>(a);
  }
  CUDA_TRY(cudaGetLastError());
  return CAAA_OK;
}

}  // namespace caaa

using namespace caaa;

extern "C" {

const char* caaa_gpu_last_error(void) { return g_err.c_str(); }

int caaa_gpu_default_map(CaaaMap* out) {
  if (!out) return fail(CAAA_E_ARG, "null argument");
  default_map(out);
  return CAAA_OK;
}

int caaa_gpu_tables_for_map(const CaaaMap* map, CaaaTables* out) {
  if (!map || !out) return fail(CAAA_E_ARG, "null argument");
  if (!valid_map(*map)) return fail(CAAA_E_ARG, "invalid map: an index out of range or more neighbours than the tables hold");
  build_tables(*map, out);
  return CAAA_OK;
}

int caaa_gpu_init(int device) {
  CaaaMap m;
  default_map(&m);
  return caaa_gpu_init_map(device, &m);
}

int caaa_gpu_init_map(int device, const CaaaMap* map) {
  Context& c = g_ctx;
  if (!map) return fail(CAAA_E_ARG, "null argument");
  if (!valid_map(*map)) return fail(CAAA_E_ARG, "invalid map: an index out of range or more neighbours than the tables hold");
  if (c.ready) {
    if (std::memcmp(&c.map, map, sizeof(CaaaMap)) == 0) return CAAA_OK;
    return fail(CAAA_E_ARG, "the library is already initialised with another map: call caaa_gpu_shutdown() first");
  }
  c.map = *map;
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
  build_dev_tables(c.map, &c.host_tables);
  CUDA_TRY(cudaMalloc(&c.d_tables, sizeof(DevTables)));
  CUDA_TRY(cudaMemcpy(c.d_tables, &c.host_tables, sizeof(DevTables), cudaMemcpyHostToDevice));
  for (LaunchSlot& l : c.slots) {
    CUDA_TRY(cudaMalloc(&l.d_ticket, sizeof(unsigned long long)));
    CUDA_TRY(cudaMalloc(&l.d_abort, sizeof(int)));
    CUDA_TRY(cudaMemset(l.d_abort, 0, sizeof(int)));
    CUDA_TRY(cudaEventCreateWithFlags(&l.done, cudaEventDisableTiming));
  }
  CUDA_TRY(cudaMalloc(&c.d_summary, sizeof(CaaaBatchSummary)));
  CUDA_TRY(cudaMalloc(&c.d_battle_counters, (BattlePipe::NBUF + 4) * sizeof(unsigned int)));
  CUDA_TRY(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&c.ev0));
  CUDA_TRY(cudaEventCreate(&c.ev1));
  const char* sy = std::getenv("CAAA_SYNC");
  c.sync_phases = sy != nullptr && std::atoi(sy) != 0;
  const char* rk = std::getenv("CAAA_ROLLOUT_KERNEL");
  (void)rk;
  const char* rw = std::getenv("CAAA_WARPS");
  c.res_warps = rw ? std::atoi(rw) : 16;
  if (c.res_warps < 1) c.res_warps = 1;
  if (c.res_warps > 32) c.res_warps = 32;
  const char* pm = std::getenv("CAAA_POOL_MULT");
  c.pool_mult = pm ? std::atoi(pm) : 2;
  if (c.pool_mult < 1) c.pool_mult = 1;
  if (c.pool_mult > MAX_POOL_MULT) c.pool_mult = MAX_POOL_MULT;
  const char* ln = std::getenv("CAAA_LANES");
  c.lanes = ln ? std::atoi(ln) : 16;
  if (c.lanes != 20 && c.lanes != 24) c.lanes = 16;
  const char* gs = std::getenv("CAAA_GROUP");
  c.group_size = gs ? std::atoi(gs) : 74;
  if (c.group_size < 1) c.group_size = 1;
  if (c.group_size > MAX_GROUP) c.group_size = MAX_GROUP;
  while ((long long)c.group_size * STREAM_SLOTS * c.pool_mult > 65536) c.group_size--;  // pool entries are 16-bit indices
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 24, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<true, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<false, 24>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_resident<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, RES_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
  CUDA_TRY(cudaFuncSetAttribute(prepare_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(advance_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, small_smem()));
  CUDA_TRY(cudaFuncSetAttribute(battle_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, BattleCfg<32>::SMEM));
  CUDA_TRY(cudaFuncSetAttribute(battle_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, BattleCfg<16>::SMEM));
#define CAAA_POOL_ATTR(NLD, NCB, NLB)                                                                                        \
  CUDA_TRY(cudaFuncSetAttribute(battle_pool_kernel<NLD, NCB, NLB>, cudaFuncAttributeMaxDynamicSharedMemorySize,              \
                                BattlePoolCfg<NLD, NCB, NLB>::SMEM))
  CAAA_POOL_ATTR(2, 6, 32);
  CAAA_POOL_ATTR(1, 8, 32);
  CAAA_POOL_ATTR(2, 8, 32);
#undef CAAA_POOL_ATTR
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
  for (LaunchSlot& l : c.slots) {
    if (l.used) cudaEventSynchronize(l.done);
    cudaFree(l.d_ticket);
    cudaFree(l.d_abort);
    if (l.d_templates) cudaFree(l.d_templates);
    if (l.d_pool) cudaFree(l.d_pool);
    cudaEventDestroy(l.done);
  }
  cudaFree(c.d_summary);
  cudaFree(c.d_battle_counters);
  c.d_battle_counters = nullptr;
  for (LeafBatch& b : c.leaf) {
    if (!b.ready) continue;
    cudaStreamSynchronize(b.stream);
    if (b.h_in) cudaFreeHost(b.h_in);
    if (b.h_out) cudaFreeHost(b.h_out);
    if (b.d_in) cudaFree(b.d_in);
    if (b.d_out) cudaFree(b.d_out);
    if (b.d_targets) cudaFree(b.d_targets);
    if (b.d_results) cudaFree(b.d_results);
    cudaFree(b.d_summary);
    cudaEventDestroy(b.done);
    cudaStreamDestroy(b.stream);
    b = LeafBatch();
  }
  if (c.battle.ready)
    for (int b = 0; b < BattlePipe::NBUF; b++) {
      cudaFreeHost(c.battle.h_in[b]);
      cudaFreeHost(c.battle.h_out[b]);
      cudaFreeHost(c.battle.h_draws[b]);
      cudaFree(c.battle.d_in[b]);
      cudaFree(c.battle.d_out[b]);
      cudaFree(c.battle.d_draws[b]);
      cudaStreamDestroy(c.battle.stream[b]);
      cudaEventDestroy(c.battle.k0[b]);
      cudaEventDestroy(c.battle.k1[b]);
      cudaEventDestroy(c.battle.done[b]);
    }
  if (c.d_in) cudaFree(c.d_in);
  if (c.d_out) cudaFree(c.d_out);
  if (c.d_aux) cudaFree(c.d_aux);
  if (c.d_aux2) cudaFree(c.d_aux2);
  cudaEventDestroy(c.ev0);
  cudaEventDestroy(c.ev1);
  cudaStreamDestroy(c.stream);
  c.ready = false;
  c.battle = BattlePipe();
  for (LaunchSlot& l : c.slots) l = LaunchSlot();
  c.next_slot = 0;
  c.d_tables = nullptr;
  c.d_summary = nullptr;
  c.d_in = c.d_out = c.d_aux = c.d_aux2 = nullptr;
  c.in_cap = c.out_cap = c.aux_cap = c.aux2_cap = 0;
  c.stream = nullptr;
  c.ev0 = c.ev1 = nullptr;
  return CAAA_OK;
}

int caaa_gpu_get_tables(CaaaTables* out) {
  if (!out) return fail(CAAA_E_ARG, "null argument");
  // host computation; identical to what caaa_gpu_init uploads (the map the library was initialised with, else the default)
  CaaaMap m;
  if (g_ctx.ready) m = g_ctx.map;
  else default_map(&m);
  build_tables(m, out);
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
  CAAA_ENTER();
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!d_states || !d_results || n_states <= 0 || rollouts_per_state <= 0) return fail(CAAA_E_ARG, "bad rollout batch arguments");
  return launch_rollout(d_states, n_states, (unsigned)rollouts_per_state, seed, first_game_id, d_results, d_stats, d_summary,
                        (cudaStream_t)cuda_stream);
}

int caaa_gpu_rollout_batch(const CaaaGameState* states, int n_states, int rollouts_per_state, uint64_t seed,
                           uint64_t first_game_id, double* results, CaaaRolloutStats* stats, CaaaBatchSummary* summary) {
  CAAA_ENTER();
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
  rc = launch_rollout(c.d_in, n_states, (unsigned)rollouts_per_state, seed, first_game_id, (double*)c.d_out,
                      stats ? (CaaaRolloutStats*)c.d_aux : nullptr, c.d_summary, c.stream);
  if (rc != CAAA_OK) return rc;
  CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
  CUDA_TRY(cudaMemcpyAsync(results, c.d_out, n_games * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (stats) CUDA_TRY(cudaMemcpyAsync(stats, c.d_aux, n_games * sizeof(CaaaRolloutStats), cudaMemcpyDeviceToHost, c.stream));
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpyAsync(&s, c.d_summary, sizeof(s), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  {
    const int rc_w = check_watchdogs();
    if (rc_w != CAAA_OK) return rc_w;
  }
  if (s.playouts != (unsigned long long)n_games) return fail(CAAA_E_CUDA, "rollout kernel finished fewer playouts than requested");
  if (summary) {
    float ms = 0.f;
    CUDA_TRY(cudaEventElapsedTime(&ms, c.ev0, c.ev1));
    s.kernel_ms = ms;
    *summary = s;
  }
  return CAAA_OK;
}

int caaa_gpu_synchronize(void) {
  CAAA_ENTER();
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  for (LaunchSlot& l : g_ctx.slots)
    if (l.used) CUDA_TRY(cudaEventSynchronize(l.done));
  return check_watchdogs();
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
  CAAA_ENTER();
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

// memcpy of a large block split over a few threads: one core moves ~8 GB/s between pageable and pinned memory, which
// is less than half of what the PCIe link carries in each direction
static void parallel_memcpy(void* dst, const void* src, size_t bytes, int threads) {
  if (threads <= 1 || bytes < (size_t)(8u << 20)) {
    std::memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> pool;
  const size_t per = (bytes / (size_t)threads + 4095) / 4096 * 4096;
  for (int k = 1; k < threads; k++) {
    const size_t lo = (size_t)k * per;
    if (lo >= bytes) break;
    const size_t cnt = std::min(per, bytes - lo);
    pool.emplace_back([=] { std::memcpy((uint8_t*)dst + lo, (const uint8_t*)src + lo, cnt); });
  }
  std::memcpy(dst, src, std::min(per, bytes));
  for (std::thread& th : pool) th.join();
}

static int launch_battles(const void* d_states, int n, uint64_t seed, uint64_t first_game_id, void* d_out, int32_t* d_draws,
                          CaaaBatchSummary* d_summary, unsigned int* d_counter, cudaStream_t stream) {
  Context& c = g_ctx;
  // pool kernel (default): one persistent CTA per SM; loader warps take the next <= 32 states from a device-wide counter
  const char* yd0 = std::getenv("CAAA_BATTLE_YIELD_DIV");
  const char* bk = std::getenv("CAAA_BATTLE_KERNEL");  // "pool" (regrouped by stage) or "refill" (battle.cuh)
  if (bk == nullptr || std::strcmp(bk, "refill") != 0) {
    const char* pc = std::getenv("CAAA_BATTLE_POOL");
    const int cfg = pc ? std::atoi(pc) : 0;
    int yield_div = yd0 ? std::atoi(yd0) : 4;
    if (yield_div < 1) yield_div = 1;
    int grid = (n + 31) / 32;
    if (grid > c.sm_count) grid = c.sm_count;
    CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream));
#define CAAA_LAUNCH_POOL(NLD, NCB, NLB)                                                                                     \
  battle_pool_kernel<NLD, NCB, NLB><<<grid, BattlePoolCfg<NLD, NCB, NLB>::WARPS * 32, BattlePoolCfg<NLD, NCB, NLB>::SMEM, stream>>>( \
      c.d_tables, (const uint8_t*)d_states, n, seed, first_game_id, (uint8_t*)d_out, d_draws, d_summary, d_counter, yield_div, d_dbg)
    unsigned long long* d_dbg = nullptr;
    const bool debug = std::getenv("CAAA_BATTLE_DEBUG") != nullptr;  // per-stage cycles / slots / claims on stderr (synchronises)
    if (debug) {
      CUDA_TRY(cudaMalloc((void**)&d_dbg, 24 * sizeof(unsigned long long)));
      CUDA_TRY(cudaMemsetAsync(d_dbg, 0, 24 * sizeof(unsigned long long), stream));
    }
    switch (cfg) {
      case 1: CAAA_LAUNCH_POOL(1, 8, 32); break;
      case 2: CAAA_LAUNCH_POOL(2, 8, 32); break;
      default: CAAA_LAUNCH_POOL(2, 6, 32); break;  // measured best: 2 loader + 6 combat warps, 199 resident battles per SM
    }
#undef CAAA_LAUNCH_POOL
    CUDA_TRY(cudaGetLastError());
    if (debug) {
      unsigned long long h[24];
      CUDA_TRY(cudaStreamSynchronize(stream));
      CUDA_TRY(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
      cudaFree(d_dbg);
      static const char* names[6] = {"load", "sea", "land", "export", "idle(loaders)", "idle(combat)"};
      for (int k = 0; k < 6; k++)
        std::fprintf(stderr, "battle pool %-14s cycles %12llu  slots %10llu  claims %9llu  requeued %9llu  slots/claim %.1f\n", names[k],
                     h[k * 4], h[k * 4 + 1], h[k * 4 + 2], h[k * 4 + 3], h[k * 4 + 2] ? (double)h[k * 4 + 1] / (double)h[k * 4 + 2] : 0.0);
    }
    return CAAA_OK;
  }
  // refill kernel: every warp takes `run` consecutive states at a time from the device-wide counter: long enough for
  // coalesced refills, short enough that the warps of all SMs finish together
  const char* bl = std::getenv("CAAA_BATTLE_LANES");  // refill kernel: battle-owning lanes per warp, 32 or 16 (battle.cuh)
  const int lanes = bl ? std::atoi(bl) : 16;
  const int warps = lanes == 32 ? BattleCfg<32>::WARPS : BattleCfg<16>::WARPS;
  int run = n / (c.sm_count * warps * 8);
  run = run < 32 ? 32 : run > 1024 ? 1024 : run / 8 * 8;
  const int want = (n + run - 1) / run;  // warps that can get a run at all
  int grid = (want + warps - 1) / warps;
  if (grid > c.sm_count) grid = c.sm_count;
  const char* yd = std::getenv("CAAA_BATTLE_YIELD_DIV");  // a phase yields once 1/yield_div of its lanes are done
  int yield_div = yd ? std::atoi(yd) : 4;
  if (yield_div < 1) yield_div = 1;
  CUDA_TRY(cudaMemsetAsync(d_counter, 0, sizeof(unsigned int), stream));
#define CAAA_LAUNCH_BATTLES(NLANES)                                                                                          \
  battle_kernel<NLANES><<<grid, BattleCfg<NLANES>::WARPS * 32, BattleCfg<NLANES>::SMEM, stream>>>(                            \
      c.d_tables, (const uint8_t*)d_states, n, seed, first_game_id, (uint8_t*)d_out, d_draws, d_summary, d_counter, run, yield_div)
  if (lanes == 32) CAAA_LAUNCH_BATTLES(32);
  else CAAA_LAUNCH_BATTLES(16);
#undef CAAA_LAUNCH_BATTLES
  CUDA_TRY(cudaGetLastError());
  return CAAA_OK;
}

int caaa_gpu_resolve_battles_device(const void* d_states, int n, uint64_t seed, uint64_t first_game_id, void* d_out_states,
                                    int32_t* d_out_draws, CaaaBatchSummary* d_summary, void* cuda_stream) {
  CAAA_ENTER();
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!d_states || n <= 0) return fail(CAAA_E_ARG, "bad battle arguments");
  Context& c = g_ctx;
  unsigned int* counter = c.d_battle_counters + BattlePipe::NBUF + (c.next_battle_counter++ % 4);  // up to four device launches in flight
  return launch_battles(d_states, n, seed, first_game_id, d_out_states, d_out_draws, d_summary, counter, (cudaStream_t)cuda_stream);
}

// Host-buffer form: chunks of BATTLE_CHUNK states flow through three pinned staging buffers and three streams, so the
// host->pinned copy of chunk c+1, the H2D / kernel / D2H of chunk c and the pinned->host copy of chunk c-1 overlap
// (pageable user buffers cannot be copied asynchronously themselves).
int caaa_gpu_resolve_battles(const CaaaGameState* states, int n, uint64_t seed, uint64_t first_game_id,
                             CaaaGameState* out_states, int32_t* out_draws, CaaaBatchSummary* summary) {
  CAAA_ENTER();
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!states || n <= 0) return fail(CAAA_E_ARG, "bad battle arguments");
  BattlePipe& bp = c.battle;
  const size_t chunk = BATTLE_CHUNK;
  if (!bp.ready) {
    for (int b = 0; b < BattlePipe::NBUF; b++) {
      CUDA_TRY(cudaHostAlloc(&bp.h_in[b], chunk * CAAA_STATE_BYTES, cudaHostAllocDefault));
      CUDA_TRY(cudaHostAlloc(&bp.h_out[b], chunk * CAAA_STATE_BYTES, cudaHostAllocDefault));
      CUDA_TRY(cudaHostAlloc((void**)&bp.h_draws[b], chunk * sizeof(int32_t), cudaHostAllocDefault));
      CUDA_TRY(cudaMalloc(&bp.d_in[b], chunk * CAAA_STATE_BYTES));
      CUDA_TRY(cudaMalloc(&bp.d_out[b], chunk * CAAA_STATE_BYTES));
      CUDA_TRY(cudaMalloc((void**)&bp.d_draws[b], chunk * sizeof(int32_t)));
      CUDA_TRY(cudaStreamCreateWithFlags(&bp.stream[b], cudaStreamNonBlocking));
      CUDA_TRY(cudaEventCreate(&bp.k0[b]));
      CUDA_TRY(cudaEventCreate(&bp.k1[b]));
      CUDA_TRY(cudaEventCreateWithFlags(&bp.done[b], cudaEventDisableTiming));
    }
    bp.ready = true;
  }
  CUDA_TRY(cudaMemsetAsync(c.d_summary, 0, sizeof(CaaaBatchSummary), c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  const int n_chunks = (int)(((size_t)n + chunk - 1) / chunk);
  const unsigned hw = std::thread::hardware_concurrency();
  int copy_threads = hw >= 16 ? 8 : hw >= 8 ? 4 : hw >= 4 ? 2 : 1;  // per direction
  if (const char* ct = std::getenv("CAAA_COPY_THREADS")) copy_threads = std::max(1, std::atoi(ct));
  // buffers the caller has page-locked (cudaHostAlloc / cudaHostRegister) are moved by the copy engines directly,
  // without the pass through the staging buffers
  auto page_locked = [](const void* p) {
    cudaPointerAttributes a;
    if (p == nullptr || cudaPointerGetAttributes(&a, p) != cudaSuccess) {
      cudaGetLastError();
      return false;
    }
    return a.type == cudaMemoryTypeHost;
  };
  const bool in_direct = page_locked(states), out_direct = page_locked(out_states), draws_direct = page_locked(out_draws);
  std::mutex mu;
  std::condition_variable cv;
  int issued = 0, drained = 0;
  bool failed = false;
  double kernel_ms = 0.0;
  // drains chunk k: waits for its D2H, copies the pinned results to the caller's buffers
  auto drain = [&](int k) -> bool {
    const int b = k % BattlePipe::NBUF;
    const size_t lo = (size_t)k * chunk, cnt = std::min(chunk, (size_t)n - lo);
    if (cudaEventSynchronize(bp.done[b]) != cudaSuccess) return false;
    if (out_states && !out_direct)
      parallel_memcpy((uint8_t*)out_states + lo * CAAA_STATE_BYTES, bp.h_out[b], cnt * CAAA_STATE_BYTES, copy_threads);
    if (out_draws && !draws_direct) std::memcpy(out_draws + lo, bp.h_draws[b], cnt * sizeof(int32_t));
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, bp.k0[b], bp.k1[b]) == cudaSuccess) kernel_ms += ms;
    return true;
  };
  std::thread drainer;
  const bool threaded = n_chunks > 1;
  if (threaded)
    drainer = std::thread([&] {
      for (int k = 0; k < n_chunks; k++) {
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&] { return issued > k || failed; });
          if (failed) return;
        }
        const bool ok = drain(k);
        std::lock_guard<std::mutex> lk(mu);
        if (!ok) failed = true;
        drained = k + 1;
        cv.notify_all();
        if (!ok) return;
      }
    });
  int rc = CAAA_OK;
  for (int k = 0; k < n_chunks && rc == CAAA_OK; k++) {
    const int b = k % BattlePipe::NBUF;
    const size_t lo = (size_t)k * chunk, cnt = std::min(chunk, (size_t)n - lo);
    if (threaded && k >= BattlePipe::NBUF) {  // buffer b is free once chunk k - NBUF has been drained
      std::unique_lock<std::mutex> lk(mu);
      cv.wait(lk, [&] { return drained > k - BattlePipe::NBUF || failed; });
      if (failed) break;
    }
    const uint8_t* h_src = (const uint8_t*)states + lo * CAAA_STATE_BYTES;
    if (!in_direct) {
      parallel_memcpy(bp.h_in[b], h_src, cnt * CAAA_STATE_BYTES, copy_threads);
      h_src = (const uint8_t*)bp.h_in[b];
    }
    cudaStream_t st = bp.stream[b];
    cudaError_t e = cudaMemcpyAsync(bp.d_in[b], h_src, cnt * CAAA_STATE_BYTES, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaEventRecord(bp.k0[b], st);
    if (e == cudaSuccess && launch_battles(bp.d_in[b], (int)cnt, seed, first_game_id + lo, out_states ? bp.d_out[b] : nullptr, bp.d_draws[b],
                                           c.d_summary, c.d_battle_counters + b, st) != CAAA_OK)
      e = cudaErrorLaunchFailure;
    if (e == cudaSuccess) e = cudaEventRecord(bp.k1[b], st);
    if (e == cudaSuccess && out_states)
      e = cudaMemcpyAsync(out_direct ? (void*)((uint8_t*)out_states + lo * CAAA_STATE_BYTES) : bp.h_out[b], bp.d_out[b],
                          cnt * CAAA_STATE_BYTES, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out_draws)
      e = cudaMemcpyAsync(draws_direct ? (void*)(out_draws + lo) : (void*)bp.h_draws[b], bp.d_draws[b], cnt * sizeof(int32_t),
                          cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaEventRecord(bp.done[b], st);
    if (e != cudaSuccess) rc = fail(CAAA_E_CUDA, "battle pipeline", e);
    if (threaded) {
      std::lock_guard<std::mutex> lk(mu);
      if (rc != CAAA_OK) failed = true;
      else issued = k + 1;
      cv.notify_all();
    } else if (rc == CAAA_OK && !drain(k)) {
      rc = fail(CAAA_E_CUDA, "battle pipeline: device error", cudaGetLastError());
    }
  }
  if (threaded) {
    {
      std::lock_guard<std::mutex> lk(mu);
      if (rc != CAAA_OK) failed = true;
      cv.notify_all();
    }
    drainer.join();
    if (failed && rc == CAAA_OK) rc = fail(CAAA_E_CUDA, "battle pipeline: device error", cudaGetLastError());
  }
  if (rc != CAAA_OK) return rc;
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpy(&s, c.d_summary, sizeof(s), cudaMemcpyDeviceToHost));
  if (s.playouts != (unsigned long long)n) return fail(CAAA_E_CUDA, "battle kernel resolved fewer states than requested");
  if (summary) {
    s.kernel_ms = kernel_ms;
    *summary = s;
  }
  return CAAA_OK;
}

int caaa_gpu_expand(const CaaaGameState* leaves, int n, uint8_t* n_actions, uint8_t* actions, CaaaGameState* children,
                    uint32_t* status) {
  CAAA_ENTER();
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
  CAAA_ENTER();
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

// AoS <-> SoA on device-resident batches (device pointers, asynchronous on `cuda_stream`).
int caaa_gpu_soa_bytes(int n, size_t* bytes) {
  if (n <= 0 || !bytes) return fail(CAAA_E_ARG, "bad arguments");
  *bytes = (size_t)SOA_CHUNKS * (size_t)((n + 31) / 32 * 32) * 16;
  return CAAA_OK;
}
int caaa_gpu_pack(const void* d_states_aos, int n, void* d_states_soa, void* cuda_stream) {
  CAAA_ENTER();
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!d_states_aos || !d_states_soa || n <= 0 || (reinterpret_cast<uintptr_t>(d_states_soa) & 15u)) return fail(CAAA_E_ARG, "bad pack arguments");
  pack_kernel<<<(n + 31) / 32, 128, 0, (cudaStream_t)cuda_stream>>>((const uint8_t*)d_states_aos, n, (n + 31) / 32 * 32, (uint4*)d_states_soa);
  CUDA_TRY(cudaGetLastError());
  return CAAA_OK;
}
int caaa_gpu_unpack(const void* d_states_soa, int n, void* d_states_aos, void* cuda_stream) {
  CAAA_ENTER();
  if (!g_ctx.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (!d_states_aos || !d_states_soa || n <= 0 || (reinterpret_cast<uintptr_t>(d_states_soa) & 15u)) return fail(CAAA_E_ARG, "bad unpack arguments");
  unpack_kernel<<<(n + 31) / 32, 128, 0, (cudaStream_t)cuda_stream>>>((const uint4*)d_states_soa, n, (n + 31) / 32 * 32, (uint8_t*)d_states_aos);
  CUDA_TRY(cudaGetLastError());
  return CAAA_OK;
}

// ---- asynchronous MCTS leaf batches ----
// Layout of the pinned / device output block of a batch of n leaves.
struct LeafOut {
  size_t scores, sums, status, picked, n_actions, actions, children, total;
  explicit LeafOut(size_t n) {
    size_t o = 0;
    scores = o; o += n * sizeof(double);
    sums = o; o += n * sizeof(double);
    status = o; o += n * sizeof(uint32_t);
    picked = o; o += n * sizeof(int32_t);
    n_actions = o; o += (n + 15) / 16 * 16;
    actions = o; o += n * CAAA_ACTIONS;
    o = (o + 15) / 16 * 16;
    children = o; o += n * CAAA_ACTIONS * CAAA_STATE_BYTES;
    total = o;
  }
};
static size_t leaf_in_bytes(size_t n) { return (n * CAAA_STATE_BYTES + 15) / 16 * 16 + n * sizeof(uint64_t); }

static int leaf_batch_reserve(LeafBatch& b, int n, size_t n_results) {
  if (!b.ready) {
    CUDA_TRY(cudaStreamCreateWithFlags(&b.stream, cudaStreamNonBlocking));
    CUDA_TRY(cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming));
    CUDA_TRY(cudaMalloc(&b.d_summary, sizeof(CaaaBatchSummary)));
    b.ready = true;
  }
  if (n > b.cap) {
    int cap = b.cap ? b.cap : 64;
    while (cap < n) cap *= 2;
    if (b.h_in) cudaFreeHost(b.h_in);
    if (b.h_out) cudaFreeHost(b.h_out);
    if (b.d_in) cudaFree(b.d_in);
    if (b.d_out) cudaFree(b.d_out);
    if (b.d_targets) cudaFree(b.d_targets);
    b.h_in = b.h_out = b.d_in = b.d_out = b.d_targets = nullptr;
    b.cap = 0;
    const LeafOut lo((size_t)cap);
    CUDA_TRY(cudaHostAlloc((void**)&b.h_in, leaf_in_bytes((size_t)cap), cudaHostAllocDefault));
    CUDA_TRY(cudaHostAlloc((void**)&b.h_out, lo.total, cudaHostAllocDefault));
    CUDA_TRY(cudaMalloc((void**)&b.d_in, leaf_in_bytes((size_t)cap)));
    CUDA_TRY(cudaMalloc((void**)&b.d_out, lo.total));
    CUDA_TRY(cudaMalloc((void**)&b.d_targets, (size_t)cap * CAAA_STATE_BYTES));
    b.cap = cap;
  }
  if (n_results > b.results_cap) {
    if (b.d_results) cudaFree(b.d_results);
    b.d_results = nullptr;
    b.results_cap = 0;
    CUDA_TRY(cudaMalloc((void**)&b.d_results, n_results * sizeof(double)));
    b.results_cap = n_results;
  }
  return CAAA_OK;
}

int caaa_gpu_set_rollout_sms(int sms) {
  CAAA_ENTER();
  g_ctx.rollout_sms = sms < 0 ? 0 : sms;
  return CAAA_OK;
}

int caaa_gpu_leaf_batch_submit(int handle, const CaaaGameState* leaves, int n, int rollouts_per_leaf, uint64_t seed,
                               uint64_t first_game_id, const uint64_t* child_pick) {
  CAAA_ENTER();
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (handle < 0 || handle >= N_LEAF_BATCHES || !leaves || !child_pick || n <= 0 || rollouts_per_leaf <= 0)
    return fail(CAAA_E_ARG, "bad leaf batch arguments");
  LeafBatch& b = c.leaf[handle];
  if (b.in_flight) return fail(CAAA_E_ARG, "leaf batch handle is still in flight: collect it first");
  int rc = leaf_batch_reserve(b, n, (size_t)n * (size_t)rollouts_per_leaf);
  if (rc != CAAA_OK) return rc;
  const LeafOut lo((size_t)b.cap);
  const size_t state_bytes = (size_t)n * CAAA_STATE_BYTES, pick_off = ((size_t)b.cap * CAAA_STATE_BYTES + 15) / 16 * 16;
  std::memcpy(b.h_in, leaves, state_bytes);
  std::memcpy(b.h_in + pick_off, child_pick, (size_t)n * sizeof(uint64_t));
  cudaStream_t st = b.stream;
  CUDA_TRY(cudaMemcpyAsync(b.d_in, b.h_in, state_bytes, cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemcpyAsync(b.d_in + pick_off, b.h_in + pick_off, (size_t)n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  CUDA_TRY(cudaMemsetAsync(b.d_summary, 0, sizeof(CaaaBatchSummary), st));
  double* d_scores = (double*)(b.d_out + lo.scores);
  double* d_sums = (double*)(b.d_out + lo.sums);
  uint32_t* d_status = (uint32_t*)(b.d_out + lo.status);
  int32_t* d_picked = (int32_t*)(b.d_out + lo.picked);
  uint8_t* d_nact = b.d_out + lo.n_actions;
  uint8_t* d_act = b.d_out + lo.actions;
  uint8_t* d_children = b.d_out + lo.children;
  const int blocks = (n + SMALL_THREADS - 1) / SMALL_THREADS;
  evaluate_kernel<<<blocks, SMALL_THREADS, small_smem(), st>>>(c.d_tables, b.d_in, n, d_scores);
  CUDA_TRY(cudaMemsetAsync(d_status, 0, (size_t)n * sizeof(uint32_t), st));
  actions_kernel<<<blocks, SMALL_THREADS, small_smem(), st>>>(c.d_tables, b.d_in, n, d_nact, d_act, d_status);
  const long long total = (long long)n * CAAA_ACTIONS;
  apply_kernel<<<(int)((total + SMALL_THREADS - 1) / SMALL_THREADS), SMALL_THREADS, small_smem(), st>>>(c.d_tables, b.d_in, n, d_nact, d_act,
                                                                                                     d_children, d_status);
  pick_kernel<<<n, 128, 0, st>>>(b.d_in, n, d_scores, d_nact, d_status, (const uint64_t*)(b.d_in + pick_off), d_children, b.d_targets, d_picked);
  CUDA_TRY(cudaGetLastError());
  rc = launch_rollout(b.d_targets, n, (unsigned)rollouts_per_leaf, seed, first_game_id, b.d_results, nullptr, b.d_summary, st);
  if (rc != CAAA_OK) return rc;
  reduce_results_kernel<<<(n * 32 + 127) / 128, 128, 0, st>>>(b.d_results, n, rollouts_per_leaf, d_sums);
  CUDA_TRY(cudaGetLastError());
  // one D2H for everything the host needs (children are only read up to n_actions per leaf, but arrive in one copy)
  const LeafOut ln((size_t)b.cap);
  CUDA_TRY(cudaMemcpyAsync(b.h_out, b.d_out, ln.children, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaMemcpyAsync(b.h_out + ln.children, b.d_out + ln.children, (size_t)n * CAAA_ACTIONS * CAAA_STATE_BYTES, cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaEventRecord(b.done, st));
  b.n = n;
  b.in_flight = true;
  return CAAA_OK;
}

int caaa_gpu_leaf_batch_collect(int handle, double* scores, uint8_t* n_actions, uint8_t* actions, CaaaGameState* children,
                                uint32_t* status, int32_t* picked, double* sum_results, CaaaBatchSummary* summary) {
  CAAA_ENTER();
  Context& c = g_ctx;
  if (!c.ready) return fail(CAAA_E_NOT_INIT, "caaa_gpu_init has not been called");
  if (handle < 0 || handle >= N_LEAF_BATCHES) return fail(CAAA_E_ARG, "bad leaf batch handle");
  LeafBatch& b = c.leaf[handle];
  if (!b.in_flight) return fail(CAAA_E_ARG, "leaf batch handle has nothing in flight");
  CUDA_TRY(cudaEventSynchronize(b.done));
  b.in_flight = false;
  const size_t n = (size_t)b.n;
  const LeafOut lo((size_t)b.cap);
  if (scores) std::memcpy(scores, b.h_out + lo.scores, n * sizeof(double));
  if (sum_results) std::memcpy(sum_results, b.h_out + lo.sums, n * sizeof(double));
  if (status) std::memcpy(status, b.h_out + lo.status, n * sizeof(uint32_t));
  if (picked) std::memcpy(picked, b.h_out + lo.picked, n * sizeof(int32_t));
  if (n_actions) std::memcpy(n_actions, b.h_out + lo.n_actions, n);
  if (actions) std::memcpy(actions, b.h_out + lo.actions, n * CAAA_ACTIONS);
  if (children) std::memcpy(children, b.h_out + lo.children, n * CAAA_ACTIONS * CAAA_STATE_BYTES);
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpy(&s, b.d_summary, sizeof(s), cudaMemcpyDeviceToHost));
  const int rc_w = check_watchdogs();
  if (rc_w != CAAA_OK) return rc_w;
  if (summary) *summary = s;
  return CAAA_OK;
}

}  // extern "C"
