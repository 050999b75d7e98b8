/* caaa_gpu.h -- C-ABI of the B200 rollout library (libcaaa_gpu.so).
 *
 * Drop-in boundary for the reference's rollout call and its two tree-expansion calls.  The
 * reference binds these through plain C++ externs (reference src/mcts.cpp:16-22, declared in
 * include/engine.h:150-159); INTEGRATION.md shows the shim a maintainer adds so that mcts.cpp
 * links unchanged.  Plain pointers and sizes only; no torch types.
 *
 * All functions return 0 on success and a negative code on failure; the message of the last
 * failure is available from caaa_gpu_last_error().  One context per process/device; calls are
 * serialised by the caller (the reference API is single-threaded as well, src/engine.cpp:102-163).
 * Unless a function says "device pointer", buffers are caller-owned host memory.
 *
 * There is no CPU fallback: every entry point fails with CAAA_E_NO_DEVICE when no CUDA device can
 * be used.
 */
#ifndef CAAA_GPU_H
#define CAAA_GPU_H

#include <stddef.h>

#include "caaa_types.h"

#ifdef __cplusplus
extern "C" {
#endif

#define CAAA_OK 0
#define CAAA_E_NO_DEVICE (-1)
#define CAAA_E_CUDA (-2)
#define CAAA_E_ARG (-3)
#define CAAA_E_NOT_INIT (-4)
#define CAAA_E_BUSY (-5) /* another thread is inside the library: the context serves one caller at a time */

/* Whole-batch counters, accumulated on the device by the rollout kernel. */
typedef struct CaaaBatchSummary {
  uint64_t playouts;
  uint64_t plies;      /* player-turns stepped */
  uint64_t decisions;  /* move / purchase / retreat draws */
  uint64_t draws;      /* random bytes consumed (decisions + dice) */
  uint64_t flagged;    /* playouts whose status word is non-zero (guards tripped) */
  double sum_result;   /* sum of results, Allies' view */
  double kernel_ms;    /* device time of the kernel(s), CUDA events on the launch stream */
} CaaaBatchSummary;

/* replaces initialize_constants() + the engine's global set-up (reference include/engine.h:29,
 * src/engine.cpp:165-178): builds the topology tables on the host and uploads them. */
int caaa_gpu_init(int device);
/* The same with the board as data (CaaaMap, caaa_types.h: what the reference compiles in from src/land.cpp, sea.cpp,
 * canal.cpp, player.cpp).  Dimensions stay 5 lands / 3 seas / 2 canals / 5 players; connectivity, land values, original
 * owners, capitals, canals and teams are the caller's.  caaa_gpu_init(device) == caaa_gpu_init_map(device, default map).
 * CAAA_E_ARG for an invalid map, or when the library is already initialised with another one. */
int caaa_gpu_init_map(int device, const CaaaMap* map);
int caaa_gpu_default_map(CaaaMap* out);                              /* host only */
int caaa_gpu_tables_for_map(const CaaaMap* map, CaaaTables* out);    /* host only: initialize_constants() for that board */
int caaa_gpu_shutdown(void);
const char* caaa_gpu_last_error(void);
int caaa_gpu_get_tables(CaaaTables* out);
int caaa_gpu_device_info(int* sm_count, int* max_smem_per_block, char* name64);

/* Batched form of `double random_play_until_terminal(GameState*)` (reference include/engine.h:158,
 * src/engine.cpp:4184-4242, called from src/mcts.cpp:94).  Plays rollouts_per_state playouts from
 * each of n_states states.  Playout i (0 <= i < n_states*rollouts_per_state) starts from state
 * i / rollouts_per_state and draws the byte stream of game id first_game_id + i under `seed`.
 * results[i] is what the reference returns; stats (optional, one per playout) carries the parity
 * side outputs; summary (optional) the whole-batch counters.  The input states are not modified. */
int caaa_gpu_rollout_batch(const CaaaGameState* states, int n_states, int rollouts_per_state, uint64_t seed,
                           uint64_t first_game_id, double* results, CaaaRolloutStats* stats,
                           CaaaBatchSummary* summary);

/* Same, with every buffer already resident in device memory (device pointers; stats/summary may be
 * NULL) and asynchronous on `cuda_stream` (a cudaStream_t, NULL = default stream).  d_summary must be
 * zeroed by the caller; kernel_ms is not filled in.  d_summary->playouts == n_states * rollouts_per_state once
 * the stream has drained says the batch is complete; caaa_gpu_synchronize() reports a watchdog abort. */
int caaa_gpu_rollout_batch_device(const void* d_states, int n_states, int rollouts_per_state, uint64_t seed,
                                  uint64_t first_game_id, double* d_results, CaaaRolloutStats* d_stats,
                                  CaaaBatchSummary* d_summary, void* cuda_stream);

/* Waits for every asynchronous launch issued through the *_device entry points and reports what they could not:
 * CAAA_E_CUDA if a kernel's watchdog stopped a batch (caaa_gpu_last_error() says so).  Up to three device
 * launches may be in flight at a time, on the same or on different streams; each owns its own scratch, and a
 * fourth launch waits on the device for the first to drain. */
int caaa_gpu_synchronize(void);

/* Single-state shim with the reference's signature semantics (include/engine.h:158): one playout
 * from *state with the next game id of an internal counter (caaa_gpu_set_stream sets seed and
 * counter).  Returns the score, or NaN on failure. */
int caaa_gpu_set_stream(uint64_t seed, uint64_t next_game_id);
double caaa_gpu_random_play_until_terminal(const CaaaGameState* state);

/* Step n states through `n_turns` whole player-turns of random play without a terminal test (the
 * body of the loop at reference src/engine.cpp:4204-4237) and return the states and the derived
 * caches (either output may be NULL).  Game id of state i is first_game_id + i.  Used for parity
 * tests of the caches and to produce randomised mid-game leaves. */
int caaa_gpu_advance(const CaaaGameState* states, int n, uint64_t seed, uint64_t first_game_id, int n_turns,
                     CaaaGameState* out_states, CaaaCacheDump* out_caches, int32_t* out_draws);

/* Isolated combat resolution (BASELINE config 3): for each state runs refresh_full_cache,
 * resolve_sea_battles, resolve_land_battles (reference src/engine.cpp:642, 2312, 3055) with random
 * dice and retreat decisions, game id first_game_id + i.  out_states / out_draws may be NULL.  Pageable buffers go
 * through the library's page-locked staging buffers (a pipeline of 131 072-state chunks); buffers the caller has
 * page-locked itself (cudaHostAlloc / cudaHostRegister) are read and written by the copy engines directly. */
int caaa_gpu_resolve_battles(const CaaaGameState* states, int n, uint64_t seed, uint64_t first_game_id,
                             CaaaGameState* out_states, int32_t* out_draws, CaaaBatchSummary* summary);

/* Same, with every buffer already resident in device memory (16-byte aligned device pointers; outputs and
 * d_summary may be NULL; d_summary is accumulated into, not zeroed) and asynchronous on `cuda_stream`. */
int caaa_gpu_resolve_battles_device(const void* d_states, int n, uint64_t seed, uint64_t first_game_id, void* d_out_states,
                                    int32_t* d_out_draws, CaaaBatchSummary* d_summary, void* cuda_stream);

/* Tree expansion (reference get_possible_actions + apply_action per action, include/engine.h:153-154,
 * src/engine.cpp:4050-4183; what expand_node does, src/mcts.cpp:110-126): for each leaf the action
 * list (n_actions[i] <= 20, actions[i*20 ..]) and, if children != NULL, the state reached by every
 * action (children[i*20 + k]).  status[i] (optional) is non-zero when a guard tripped, i.e. the
 * reference would not return from that call ("expansion-mode hangs"): bit k = applying
 * action k, bit 31 = computing the action list itself. */
int caaa_gpu_expand(const CaaaGameState* leaves, int n, uint8_t* n_actions, uint8_t* actions,
                    CaaaGameState* children, uint32_t* status);

/* AoS <-> SoA (SURVEY.md 8b `caaa_gpu_pack/unpack`): the chunk-major structure-of-arrays form of a batch of states --
 * 42 chunks of 16 bytes per state; chunk c of state i at byte (c * n_pad + i) * 16, n_pad = n rounded up to 32 -- in
 * which 32 consecutive states are read with one 512-byte LDG.128 per chunk.  Device pointers (SoA 16-byte aligned,
 * caaa_gpu_soa_bytes(n) bytes), asynchronous on `cuda_stream`.  DESIGN.md section 3 records where this layout was
 * measured against the array-of-records form the kernels use. */
int caaa_gpu_soa_bytes(int n, size_t* bytes);
int caaa_gpu_pack(const void* d_states_aos, int n, void* d_states_soa, void* cuda_stream);
int caaa_gpu_unpack(const void* d_states_soa, int n, void* d_states_aos, void* cuda_stream);

/* ---- Asynchronous leaf batches: one MCTS iteration's device work for n leaves at once (reference src/mcts.cpp:78-94:
 * is_terminal_state, expand_node, `rand() % num_children`, random_play_until_terminal), without a host round trip
 * between the steps.  submit() enqueues, on the handle's own stream: terminal test -> action lists -> all children ->
 * the child to roll out (child_pick[i] % number of playable children; terminal / childless leaves roll out from
 * themselves) -> rollouts_per_leaf playouts from each picked child (game ids first_game_id + i * rollouts_per_leaf + r)
 * -> per-leaf sums of the results in a fixed order.  The children stay on the device for the rollout; collect() waits
 * and returns what the host tree needs: scores, action lists, children, status (as caaa_gpu_expand), picked[i] (index
 * into leaf i's action list, -1 = the leaf itself) and sum_results[i].  Handles 0..2 are independent: a search keeps
 * two or three batches in flight.  Any output pointer may be NULL. */
int caaa_gpu_leaf_batch_submit(int handle, const CaaaGameState* leaves, int n, int rollouts_per_leaf, uint64_t seed,
                               uint64_t first_game_id, const uint64_t* child_pick);
int caaa_gpu_leaf_batch_collect(int handle, double* scores, uint8_t* n_actions, uint8_t* actions, CaaaGameState* children,
                                uint32_t* status, int32_t* picked, double* sum_results, CaaaBatchSummary* summary);
/* SMs a rollout launch may occupy (0 = all).  The persistent rollout kernel otherwise holds every SM until its batch
 * has drained; a pipelined search leaves a few SMs free so that the next batch's short expansion kernels run beside it. */
int caaa_gpu_set_rollout_sms(int sms);

/* is_terminal_state / evaluate_state on caller-supplied states, with the unit totals taken from
 * the states themselves (the reference reads stale globals here, src/engine.cpp:4244-4248,4314). */
int caaa_gpu_evaluate(const CaaaGameState* states, int n, double* scores);

/* ---- Root-parallel MCTS (BASELINE config 5): the one exchange step.  Every rank searches its own tree
 * (reference src/mcts.cpp:65-108 per rank, seed = base + rank); the root children's {visits, value} are then summed
 * over the ranks with ncclAllReduce over NVLink / NVSwitch.  Root actions are in the same order on every rank because
 * get_possible_actions is deterministic.  NCCL is bound at run time (dlopen), not at load time. */
/* `nccl_comm` is the caller's ncclComm_t; device pointers, in place, asynchronous on `cuda_stream`; both reductions
 * are one NCCL group. */
int caaa_root_allreduce(void* nccl_comm, int n_children, int64_t* d_visits, double* d_values, void* cuda_stream);
/* The library's own communicator, for hosts that have none: rank 0 makes the 128-byte id, every rank (one per GPU,
 * after cudaSetDevice / caaa_gpu_init) joins with it, then sums host buffers in place. */
int caaa_root_unique_id(uint8_t id128[128]);
int caaa_root_comm_init(const uint8_t id128[128], int rank, int world);
int caaa_root_allreduce_host(int n_children, int64_t* visits, double* values, double* elapsed_us);
int caaa_root_comm_destroy(void);
const char* caaa_root_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* CAAA_GPU_H */
