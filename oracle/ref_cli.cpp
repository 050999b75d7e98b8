// ref_cli.cpp -- TEST/BENCH INFRASTRUCTURE.  Times the patched reference's own CPU rollout loop
// (random_play_until_terminal, src/engine.cpp:4184-4242) in one process.  bench.py starts one
// of these per host core because the reference engine keeps all state in file-scope globals
// and therefore cannot be threaded.
//
// usage: ref_cli <state670.bin> <mode 0=glibc|1=philox> <seed> <first_game_id> <n_playouts>
// prints one line: playouts plies decisions draws seconds sum_result
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
}

int main(int argc, char** argv) {
  if (argc < 6) {
    fprintf(stderr, "usage: %s state.bin mode seed first_game_id n\n", argv[0]);
    return 2;
  }
  unsigned char st[CAAA_STATE_BYTES];
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(st, 1, sizeof(st), f) != sizeof(st)) {
    fprintf(stderr, "cannot read %s\n", argv[1]);
    return 2;
  }
  fclose(f);
  int mode = atoi(argv[2]);
  unsigned long long seed = strtoull(argv[3], nullptr, 0);
  long long first = atoll(argv[4]);
  int n = atoi(argv[5]);
  caaa_ref_init();
  caaa_ref_set_stream(mode, seed);
  CaaaRolloutStats* stats = (CaaaRolloutStats*)calloc((size_t)n, sizeof(CaaaRolloutStats));
  auto t0 = std::chrono::steady_clock::now();
  caaa_ref_rollout_many(st, first, n, stats, nullptr);
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
  printf("%d %lld %lld %lld %.6f %.9f\n", n, plies, decisions, draws, secs, sum);
  free(stats);
  return 0;
}
