// ref_cli.cpp -- TEST/BENCH INFRASTRUCTURE.  Runs the patched reference's own CPU code
// (random_play_until_terminal, src/engine.cpp:4184-4242; mcts_search, src/mcts.cpp:65-108) in one process.
// bench.py and the full-size parity tests start one of these per host core because the reference engine keeps
// all state in file-scope globals and therefore cannot be threaded.
//
// usage: ref_cli <states.bin> <mode 0=glibc|1=philox> <seed> <first_game_id> <n_playouts> [stats_out.bin [per_state]]
//          states.bin holds one or more 670-byte states.  Playout i (0 <= i < n_playouts) has game id
//          first_game_id + i and starts from state (first_game_id + i - id0) / per_state, where id0 is the game id of
//          the first playout of state 0 (= 0 unless CAAA_REF_ID0 is set); per_state = 0 / absent: always state 0.
//          stats_out.bin receives one 32-byte CaaaRolloutStats per playout.
//        prints one line: playouts plies decisions draws seconds sum_result
//        ref_cli mcts <state.bin> <iterations> [literal]    (literal: without the per-playout valid_moves reset, P1 iv)
//          prints: iterations seconds  then one line per root child: action visits value   and "best <action>"
#include "../include/caaa_types.h"
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>

extern "C" {
void caaa_ref_init(void);
void caaa_ref_set_stream(int mode, uint64_t seed);
void caaa_ref_rollout_many(const void* state670, int64_t first_game_id, int n, CaaaRolloutStats* st,
                           double* results);
int caaa_ref_mcts(const void* state670, long iterations, uint8_t* actions, int32_t* visits, double* values,
                  uint8_t* best_action);
void caaa_ref_set_reset_valid_moves(int on);
}

static unsigned char* read_states(const char* path, long* n_states) {
  FILE* f = fopen(path, "rb");
  if (!f) return nullptr;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < CAAA_STATE_BYTES || sz % CAAA_STATE_BYTES) {
    fclose(f);
    return nullptr;
  }
  unsigned char* buf = (unsigned char*)malloc((size_t)sz);
  if (fread(buf, 1, (size_t)sz, f) != (size_t)sz) {
    fclose(f);
    free(buf);
    return nullptr;
  }
  fclose(f);
  *n_states = sz / CAAA_STATE_BYTES;
  return buf;
}

int main(int argc, char** argv) {
  if (argc >= 4 && strcmp(argv[1], "mcts") == 0) {
    long ns = 0;
    unsigned char* st = read_states(argv[2], &ns);
    if (!st) {
      fprintf(stderr, "cannot read %s\n", argv[2]);
      return 2;
    }
    long iters = atol(argv[3]);
    caaa_ref_init();
    caaa_ref_set_stream(0, 0);  // the reference's own glibc stream
    if (argc > 4 && strcmp(argv[4], "literal") == 0) caaa_ref_set_reset_valid_moves(0);  // P0 behaviour: SURVEY Appendix E
    uint8_t actions[CAAA_ACTIONS], best = 255;
    int32_t visits[CAAA_ACTIONS];
    double values[CAAA_ACTIONS];
    fflush(stdout);
    FILE* keep = stdout;
    (void)keep;
    auto t0 = std::chrono::steady_clock::now();
    int n = caaa_ref_mcts(st, iters, actions, visits, values, &best);
    auto t1 = std::chrono::steady_clock::now();
    printf("\nMCTS %ld %.6f\n", iters, std::chrono::duration<double>(t1 - t0).count());
    for (int i = 0; i < n; i++) printf("CHILD %d %d %.6f\n", (int)actions[i], (int)visits[i], values[i]);
    printf("BEST %d\n", (int)best);
    return 0;
  }
  if (argc < 6) {
    fprintf(stderr, "usage: %s states.bin mode seed first_game_id n [stats_out.bin [per_state]]\n       %s mcts state.bin iterations\n", argv[0], argv[0]);
    return 2;
  }
  long n_states = 0;
  unsigned char* st = read_states(argv[1], &n_states);
  if (!st) {
    fprintf(stderr, "cannot read %s\n", argv[1]);
    return 2;
  }
  int mode = atoi(argv[2]);
  unsigned long long seed = strtoull(argv[3], nullptr, 0);
  long long first = atoll(argv[4]);
  int n = atoi(argv[5]);
  const char* out_path = argc > 6 ? argv[6] : nullptr;
  long long per_state = argc > 7 ? atoll(argv[7]) : 0;
  const char* id0s = getenv("CAAA_REF_ID0");
  long long id0 = id0s ? atoll(id0s) : 0;
  caaa_ref_init();
  caaa_ref_set_stream(mode, seed);
  CaaaRolloutStats* stats = (CaaaRolloutStats*)calloc((size_t)n, sizeof(CaaaRolloutStats));
  auto t0 = std::chrono::steady_clock::now();
  if (per_state <= 0) {
    caaa_ref_rollout_many(st, first, n, stats, nullptr);
  } else {
    int i = 0;
    while (i < n) {  // runs of playouts that share a start state
      long long s = (first + i - id0) / per_state;
      long long run_end = (s + 1) * per_state + id0 - first;
      int m = (int)((run_end < n ? run_end : n) - i);
      if (s < 0 || s >= n_states) {
        fprintf(stderr, "game id %lld maps to state %lld of %ld\n", first + i, s, n_states);
        return 2;
      }
      caaa_ref_rollout_many(st + (size_t)s * CAAA_STATE_BYTES, first + i, m, stats + i, nullptr);
      i += m;
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  double secs = std::chrono::duration<double>(t1 - t0).count();
  long long plies = 0, decisions = 0, draws = 0;
  double sum = 0;
  for (int i = 0; i < n; i++) {
    plies += stats[i].turns;
    decisions += stats[i].decisions;
    draws += stats[i].draws;
    sum += stats[i].result;
  }
  if (out_path) {
    FILE* o = fopen(out_path, "wb");
    if (!o || fwrite(stats, sizeof(CaaaRolloutStats), (size_t)n, o) != (size_t)n) {
      fprintf(stderr, "cannot write %s\n", out_path);
      return 2;
    }
    fclose(o);
  }
  printf("%d %lld %lld %lld %.6f %.9f\n", n, plies, decisions, draws, secs, sum);
  free(stats);
  free(st);
  return 0;
}
