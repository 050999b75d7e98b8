// caaa_gpu.cu -- sm_100a kernels and the C-ABI of libcaaa_gpu.so (include/caaa_gpu.h).
//
// Execution model: one persistent CTA per SM; 224 games (7 warps x 32 lanes, one game per
// lane) and the topology tables stay resident in the CTA's 227 KB of shared memory for the whole
// playout, so the only HBM traffic of a playout is its start state in and its result out.  All 32 lanes
// of a warp walk the turn pipeline together (engine2.cuh is written so that they stay converged:
// flat cursor loops, per-game work masks).  Finished games are replaced from a global ticket counter by
// a warp-cooperative copy of the pre-imported start state, which absorbs the long tail of game
// lengths.  No tensor cores: the work is branchy uint8 arithmetic.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>

#include "../../include/caaa_gpu.h"
#include "engine2.cuh"
#include "host_tables.h"

namespace caaa {
using namespace e2;

constexpr int TABLE_BYTES = (int)((sizeof(DevTables) + 15) / 16 * 16);
constexpr int GAME_WORDS = (int)(sizeof(Game) / 4);
constexpr int ROLLOUT_WARPS = 7;                     // 7 x 32 games x 924 B + tables = 208,336 B of shared memory
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
  int switch_below;
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
    // new game's 231 words from the pre-imported state (coalesced, converged) ----
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
  if (i >= n) return;
  const unsigned m = __activemask();  // the lanes of this warp that hold a state walk the pipeline in lockstep
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

// Isolated combat resolution (BASELINE config 3).
__global__ void __launch_bounds__(ROLLOUT_THREADS, 1) battle_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                               unsigned long long seed, unsigned long long first_game_id,
                                                               uint8_t* out_states, int32_t* out_draws, CaaaBatchSummary* summary) {
  const DevTables* T = stage_tables(tables);
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned long long draws = 0, flagged = 0, one = 0;
  const unsigned m = __ballot_sync(FULL, i < n);
  if (i < n) {
    Game* g = my_game();
    const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
    import_state(g, states + (size_t)i * CAAA_STATE_BYTES);
    begin_playout(g, cx, first_game_id + (unsigned long long)i);
    resolve_sea_battles<false>(g, cx, m);
    resolve_land_battles<false>(g, cx, m);
    if (out_states) export_state(g, out_states + (size_t)i * CAAA_STATE_BYTES);
    if (out_draws) out_draws[i] = (int32_t)g->draws;
    draws = g->draws;
    flagged = g->status != 0;
    one = 1;
  }
  if (summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      draws += __shfl_down_sync(FULL, draws, o);
      flagged += __shfl_down_sync(FULL, flagged, o);
      one += __shfl_down_sync(FULL, one, o);
    }
    if ((threadIdx.x & 31) == 0) {
      atomicAdd((unsigned long long*)&summary->playouts, one);
      atomicAdd((unsigned long long*)&summary->draws, draws);
      atomicAdd((unsigned long long*)&summary->flagged, flagged);
    }
  }
}

// Stop-at-decision mode (reference get_possible_actions / apply_action, src/engine.cpp:4050-4183).
__device__ __forceinline__ void run_until_decision(Game* g, const RunCtx cx) {
  for (;;) {
    bool stopped = false;
    for (int ph = 0; ph <= CAAA_PH_BUY_UNITS && !stopped; ph++) stopped = run_phase<true>(g, cx, ph);
    if (stopped) return;
    for (int ph = CAAA_PH_CRASH_AIR; ph < CAAA_PH_COUNT; ph++) run_phase<true>(g, cx, ph);
  }
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
  void* d_templates = nullptr;  // pre-imported start states + their first scores
  size_t templates_cap = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  unsigned long long seed = 0, next_game_id = 0;
  bool sync_phases = false;  // CAAA_SYNC=1: CTA barrier after every phase (A/B knob)
  int kernel = 1;            // kernel of the most recent launch: 0 = SM-resident, 1 = phase-regrouped over a global pool, 2 = lockstep
  int small_kernel = 2;      // the kernel for batches of up to small_batch playouts
  int res_warps = 16;        // warps per CTA of the resident kernel (CAAA_WARPS)
  int* d_abort = nullptr;    // watchdog flag of the resident kernel
  long long small_batch = 32768;  // batches up to this size go to the lockstep kernel (no pool hops: lower latency);
                                  // CAAA_ROLLOUT_KERNEL=stream sets it to 0
  int pool_mult = 2;         // pool entries per shared-memory slot (CAAA_POOL_MULT); 2 keeps the 62 MB pool L2-resident
  int lanes = 16;            // games per warp of the phase-regrouped kernel (CAAA_LANES = 8 | 16 | 32)
  int group_size = 16;       // CTAs that share one game pool and one set of phase bitmaps (CAAA_GROUP)
  void* d_pool = nullptr;    // game pools + group state of the phase-regrouped kernel
  size_t pool_cap = 0;
  GroupState* d_groups_last = nullptr;  // group state of the most recent launch (watchdog flag)
  int n_groups_last = 0;
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
static int small_smem() { return TABLE_BYTES + SMALL_THREADS * (int)sizeof(Game); }
static int rollout_smem() { return TABLE_BYTES + ROLLOUT_THREADS * (int)sizeof(Game); }

static int launch_rollout(const void* d_states, int n_states, unsigned rollouts_per_state, unsigned long long seed,
                          unsigned long long first_game_id, double* d_results, CaaaRolloutStats* d_stats,
                          CaaaBatchSummary* d_summary, cudaStream_t stream) {
  Context& c = g_ctx;
  const unsigned long long n_games = (unsigned long long)n_states * rollouts_per_state;
  const size_t tmpl_bytes = (size_t)n_states * sizeof(Game);
  int rc = grow(&c.d_templates, &c.templates_cap, tmpl_bytes + (size_t)n_states * sizeof(double));
  if (rc != CAAA_OK) return rc;
  uint32_t* d_tmpl = (uint32_t*)c.d_templates;
  double* d_scores0 = (double*)((uint8_t*)c.d_templates + (tmpl_bytes + 7) / 8 * 8);
  if ((tmpl_bytes + 7) / 8 * 8 + (size_t)n_states * sizeof(double) > c.templates_cap) {
    rc = grow(&c.d_templates, &c.templates_cap, (tmpl_bytes + 7) / 8 * 8 + (size_t)n_states * sizeof(double));
    if (rc != CAAA_OK) return rc;
    d_tmpl = (uint32_t*)c.d_templates;
    d_scores0 = (double*)((uint8_t*)c.d_templates + (tmpl_bytes + 7) / 8 * 8);
  }
  prepare_kernel<<<(n_states + SMALL_THREADS - 1) / SMALL_THREADS, SMALL_THREADS, small_smem(), stream>>>(
      c.d_tables, (const uint8_t*)d_states, n_states, d_tmpl, d_scores0);
  CUDA_TRY(cudaGetLastError());
  CUDA_TRY(cudaMemsetAsync(c.d_ticket, 0, sizeof(unsigned long long), stream));
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
  a.ticket = c.d_ticket;
  {
    const char* sb = std::getenv("CAAA_SWITCH_BELOW");
    a.switch_below = sb ? std::atoi(sb) : 8;
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
  a.abort_flag = c.d_abort;
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
    const int grid_r = (int)(want32 < (unsigned long long)c.sm_count ? want32 : (unsigned long long)c.sm_count);
    const unsigned long long per_cta = (n_games + grid_r - 1) / grid_r;
    int n_slots = (int)((per_cta + 31) / 32 * 32);
    if (n_slots > RES_SLOTS) n_slots = RES_SLOTS;
    if (n_slots < 32) n_slots = 32;
    ResDbg* d_dbg = nullptr;
    if (std::getenv("CAAA_DEBUG") != nullptr) {
      int rc2 = grow(&c.d_pool, &c.pool_cap, sizeof(ResDbg));
      if (rc2 != CAAA_OK) return rc2;
      d_dbg = (ResDbg*)c.d_pool;
      CUDA_TRY(cudaMemsetAsync(d_dbg, 0, sizeof(ResDbg), stream));
    }
    CUDA_TRY(cudaMemsetAsync(c.d_abort, 0, sizeof(int), stream));
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
  const int grid = (int)(want < (unsigned long long)c.sm_count ? want : (unsigned long long)c.sm_count);
  if (c.kernel == 1) {
    const size_t need = (size_t)c.sm_count * STREAM_SLOTS * (size_t)c.pool_mult * sizeof(Game);
    const int n_groups = (grid + c.group_size - 1) / c.group_size;
    const size_t need16 = (need + 15) / 16 * 16;
    int rc2 = grow(&c.d_pool, &c.pool_cap, need16 + (size_t)n_groups * sizeof(GroupState) + sizeof(StreamDbg));
    if (rc2 != CAAA_OK) return rc2;
    GroupState* d_groups = (GroupState*)((uint8_t*)c.d_pool + need16);
    group_init_kernel<<<n_groups, 256, 0, stream>>>(d_groups, n_groups, STREAM_SLOTS * c.pool_mult, c.group_size, grid);
    StreamDbg* d_dbg = nullptr;
    if (std::getenv("CAAA_DEBUG") != nullptr) {
      d_dbg = (StreamDbg*)((uint8_t*)d_groups + (size_t)n_groups * sizeof(GroupState));
      CUDA_TRY(cudaMemsetAsync(d_dbg, 0, sizeof(StreamDbg), stream));
    }
#define CAAA_LAUNCH_STREAM(L)                                                                                                        \
  do {                                                                                                                              \
    if (d_stats)                                                                                                                    \
      rollout_kernel_stream<true, L><<<grid, STREAM_SLOTS / L * 32, STREAM_SMEM, stream>>>(a, (uint32_t*)c.d_pool, c.pool_mult, d_groups, c.group_size, d_dbg);  \
    else                                                                                                                            \
      rollout_kernel_stream<false, L><<<grid, STREAM_SLOTS / L * 32, STREAM_SMEM, stream>>>(a, (uint32_t*)c.d_pool, c.pool_mult, d_groups, c.group_size, d_dbg); \
  } while (0)
    if (c.lanes == 32) CAAA_LAUNCH_STREAM(32);
    else if (c.lanes == 8) CAAA_LAUNCH_STREAM(8);
    else CAAA_LAUNCH_STREAM(16);
    c.d_groups_last = d_groups;
    c.n_groups_last = n_groups;
    if (d_dbg != nullptr) {
      StreamDbg h;
      cudaStreamSynchronize(stream);
      cudaMemcpy(&h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
      const double tt = (double)(h.t[0] + h.t[1] + h.t[2] + h.t[3] + h.t[4]);
      std::fprintf(stderr, "stream dbg: batches %llu items/batch %.1f  idle+claim %.3f in %.3f compute %.3f restart %.3f out %.3f  Mcyc/warp %.1f\n",
                   h.batches, h.batches ? (double)h.items / (double)h.batches : 0.0, h.t[0] / tt, h.t[1] / tt, h.t[2] / tt, h.t[3] / tt,
                   h.t[4] / tt, tt / (grid * (STREAM_SLOTS / c.lanes)) / 1e6);
      std::fprintf(stderr, "stream dbg: per kind  share of phase time / share of game visits:");
      double kc = 0, ki = 0;
      for (int q = 0; q < NQ; q++) { kc += (double)h.kind_cycles[q]; ki += (double)h.kind_items[q]; }
      for (int q = 0; q < NQ; q++)
        if (h.kind_items[q]) std::fprintf(stderr, "  k%d %.3f/%.3f", q, h.kind_cycles[q] / kc, h.kind_items[q] / ki);
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
  CUDA_TRY(cudaMalloc(&c.d_abort, sizeof(int)));
  CUDA_TRY(cudaMemset(c.d_abort, 0, sizeof(int)));
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
  if (c.lanes != 8 && c.lanes != 32) c.lanes = 16;
  const char* gs = std::getenv("CAAA_GROUP");
  c.group_size = gs ? std::atoi(gs) : 16;
  if (c.group_size < 1) c.group_size = 1;
  if (c.group_size > MAX_GROUP) c.group_size = MAX_GROUP;
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
  CUDA_TRY(cudaFuncSetAttribute(rollout_kernel_stream<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, STREAM_SMEM));
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
  CUDA_TRY(cudaFuncSetAttribute(battle_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, rollout_smem()));
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
  cudaFree(c.d_abort);
  if (c.d_in) cudaFree(c.d_in);
  if (c.d_out) cudaFree(c.d_out);
  if (c.d_aux) cudaFree(c.d_aux);
  if (c.d_aux2) cudaFree(c.d_aux2);
  if (c.d_templates) cudaFree(c.d_templates);
  if (c.d_pool) cudaFree(c.d_pool);
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
  return launch_rollout(d_states, n_states, (unsigned)rollouts_per_state, seed, first_game_id, d_results, d_stats, d_summary,
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
  rc = launch_rollout(c.d_in, n_states, (unsigned)rollouts_per_state, seed, first_game_id, (double*)c.d_out,
                      stats ? (CaaaRolloutStats*)c.d_aux : nullptr, c.d_summary, c.stream);
  if (rc != CAAA_OK) return rc;
  CUDA_TRY(cudaEventRecord(c.ev1, c.stream));
  CUDA_TRY(cudaMemcpyAsync(results, c.d_out, n_games * sizeof(double), cudaMemcpyDeviceToHost, c.stream));
  if (stats) CUDA_TRY(cudaMemcpyAsync(stats, c.d_aux, n_games * sizeof(CaaaRolloutStats), cudaMemcpyDeviceToHost, c.stream));
  CaaaBatchSummary s;
  CUDA_TRY(cudaMemcpyAsync(&s, c.d_summary, sizeof(s), cudaMemcpyDeviceToHost, c.stream));
  CUDA_TRY(cudaStreamSynchronize(c.stream));
  if (c.kernel == 0) {
    int aborted = 0;
    CUDA_TRY(cudaMemcpy(&aborted, c.d_abort, sizeof(int), cudaMemcpyDeviceToHost));
    if (aborted) return fail(CAAA_E_CUDA, "rollout kernel stopped by its watchdog: no claimable game although the batch was unfinished");
  }
  for (int gidx = 0; c.kernel == 1 && gidx < c.n_groups_last; gidx++) {  // watchdog of the phase-regrouped kernel
    int aborted = 0;
    CUDA_TRY(cudaMemcpy(&aborted, &c.d_groups_last[gidx].aborted, sizeof(int), cudaMemcpyDeviceToHost));
    if (aborted) return fail(CAAA_E_CUDA, "rollout kernel stopped by its watchdog: no claimable game although the batch was unfinished");
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
  const int grid = (n + ROLLOUT_THREADS - 1) / ROLLOUT_THREADS;  // 7 warps per SM: the sweep is bound by per-warp latency
  battle_kernel<<<grid, ROLLOUT_THREADS, rollout_smem(), c.stream>>>(c.d_tables, (const uint8_t*)c.d_in, n, seed, first_game_id,
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
