/* caaa_oracle.h -- TEST INFRASTRUCTURE.  Plain-C CPU restatement of the reference's rollout
 * path (reference src/engine.cpp).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may build, link or call anything declared here; the
 * product library (caaa_b200/csrc) never does.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this
 * port is pinned against the reference itself, compiled here as oracle/_ref (patch sets
 * P0/P1/P2, oracle/build_ref.py) -- bit-exact final states, results, turn/decision/draw counts
 * and per-phase caches on seeded playouts -- and against the fixtures under tests/golden/
 * generated from that build.
 */
#ifndef CAAA_ORACLE_H
#define CAAA_ORACLE_H
#include "../include/caaa_types.h"

#ifdef __cplusplus
extern "C" {
#endif

void caaa_oracle_build_tables(CaaaTables* t);

void caaa_oracle_init(void);
void caaa_oracle_set_stream(int mode, uint64_t seed); /* mode must be 1 (Philox) */
void caaa_oracle_set_reset_valid_moves(int on);
void caaa_oracle_set_game_id(int64_t game_id);
double caaa_oracle_rollout(const void* state670, CaaaRolloutStats* st);
void caaa_oracle_rollout_many(const void* state670, int64_t first_game_id, int n, CaaaRolloutStats* st,
                              double* results);
void caaa_oracle_get_state(void* out670);
int caaa_oracle_get_possible_actions(const void* state670, uint8_t* actions20);
void caaa_oracle_apply_action(void* state670, uint8_t action);
int caaa_oracle_is_terminal(const void* state670);
void caaa_oracle_load_state(const void* state670, int64_t game_id);
int caaa_oracle_run_phase(int phase);
double caaa_oracle_evaluate(void);
int caaa_oracle_draws(void);
void caaa_oracle_get_cache(CaaaCacheDump* c);
void caaa_oracle_get_tables(CaaaTables* t);
int caaa_oracle_stuck(void);

#ifdef __cplusplus
}
#endif
#endif
