/* caaa_host.h -- host-side C-ABI that needs no GPU: game_data.json load/save in the reference's
 * wire format (reference src/serialize_data.cpp:75-404; the reference depends on the system cJSON
 * package for this, this library carries its own small parser) and the MCTS driver that replaces
 * the reference's search loop (reference src/mcts.cpp:65-126) with batched GPU expansion and
 * rollouts.
 */
#ifndef CAAA_HOST_H
#define CAAA_HOST_H

#include "caaa_types.h"

#ifdef __cplusplus
extern "C" {
#endif

/* load_game_data_from_json (reference include/serialize_data.h:7, src/serialize_data.cpp:10-20,75-256).
 * Fields missing from the file keep their value in *state (the reference zero-fills the state
 * first, src/engine.cpp:633-641; pass a zeroed state for the same effect).  0 = ok. */
int caaa_state_from_json_file(const char* filename, CaaaGameState* state);
int caaa_state_from_json_text(const char* text, CaaaGameState* state);
/* serialize_game_data_to_json + cJSON_Print (reference src/serialize_data.cpp:258-322): writes at
 * most cap bytes (NUL-terminated) and returns the length needed, or a negative error. */
int caaa_state_to_json_text(const CaaaGameState* state, char* out, int cap);
int caaa_state_to_json_file(const char* filename, const CaaaGameState* state);

/* Root statistics of a search: one entry per root action, in get_possible_actions order. */
typedef struct CaaaRootStats {
  int32_t n_children;
  uint8_t actions[CAAA_ACTIONS];
  int64_t visits[CAAA_ACTIONS];
  double values[CAAA_ACTIONS];
  int64_t iterations;      /* tree iterations performed */
  int64_t playouts;        /* rollouts played */
  int64_t nodes;           /* tree nodes allocated */
} CaaaRootStats;

/* mcts_search + select_best_action (reference src/mcts.cpp:65-108,247-258) on the GPU engine:
 * UCT selection with the reference's formula and +1 smoothing, expansion of every visited leaf,
 * a uniformly random child, `rollouts_per_leaf` playouts of that child in one device batch
 * (1 = the reference's behaviour), back-propagation of the summed results.  `leaves_per_batch`
 * leaves are selected per device launch with virtual visits (1 = the reference's sequential loop).
 * Needs caaa_gpu_init().  best_action may be NULL. */
int caaa_mcts_search(const CaaaGameState* root_state, int64_t iterations, int rollouts_per_leaf,
                     int leaves_per_batch, uint64_t seed, CaaaRootStats* out, uint8_t* best_action);
/* Same with an explicit pipeline depth (1..3 leaf batches in flight; caaa_mcts_search uses 1 for leaves_per_batch == 1,
 * else 2).  Selection for batch k+1 runs while batch k is on the device, with virtual visits for its pending results. */
int caaa_mcts_search_pipelined(const CaaaGameState* root_state, int64_t iterations, int rollouts_per_leaf,
                               int leaves_per_batch, int depth, uint64_t seed, CaaaRootStats* out, uint8_t* best_action);

#ifdef __cplusplus
}
#endif
#endif /* CAAA_HOST_H */
