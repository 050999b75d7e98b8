// engine_compat.cpp -- reference-named entry points forwarding to the C-ABI (see caaa_engine_compat.h).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/caaa_engine_compat.h"
#include "../../include/caaa_gpu.h"
#include "../../include/caaa_host.h"

static_assert(sizeof(GameState) == sizeof(CaaaGameState) && alignof(GameState) == alignof(CaaaGameState), "GameState layout");
static inline CaaaGameState* C(GameState* s) { return reinterpret_cast<CaaaGameState*>(s); }

static GameState g_current;  // the reference's global `state` (src/engine.cpp:102) as far as callers can see it

static void die(const char* what) {  // the reference API has no error returns; a failed device call cannot be ignored
  std::fprintf(stderr, "caaa_gpu: %s: %s\n", what, caaa_gpu_last_error());
  std::abort();
}

void initialize_constants() {
  const char* d = std::getenv("CAAA_DEVICE");
  if (caaa_gpu_init(d ? std::atoi(d) : 0) != CAAA_OK) die("initialize_constants");
}
void load_game_data(char const* filename) {  // reference src/engine.cpp:633-641
  std::memset(&g_current, 0, sizeof(g_current));
  if (caaa_state_from_json_file(filename, C(&g_current)) != 0) {
    std::perror("Failed to open file");  // reference src/serialize_data.cpp:38-41
    std::exit(1);
  }
}
GameState* get_game_state_copy() { return clone_state(&g_current); }
GameState* clone_state(GameState* game_state) {
  GameState* s = (GameState*)std::malloc(sizeof(GameState));
  std::memcpy(s, game_state, sizeof(GameState));
  return s;
}
void free_state(GameState* game_state) { std::free(game_state); }
void get_possible_actions(GameState* game_state, uint8_t* num_actions, ActionsPtr actions) {
  uint8_t n = 0, acts[CAAA_ACTIONS];
  if (caaa_gpu_expand(C(game_state), 1, &n, acts, nullptr, nullptr) != CAAA_OK) die("get_possible_actions");
  *num_actions = n;
  std::memcpy(*actions, acts, n);
}
void apply_action(GameState* game_state, uint8_t action) {
  uint8_t n = 0, acts[CAAA_ACTIONS];
  static GameState children[CAAA_ACTIONS];
  if (caaa_gpu_expand(C(game_state), 1, &n, acts, C(children), nullptr) != CAAA_OK) die("apply_action");
  for (int k = 0; k < n; k++)
    if (acts[k] == action) {
      *game_state = children[k];
      return;
    }
  std::fprintf(stderr, "caaa_gpu: apply_action: action %d is not offered in this state\n", (int)action);
}
double evaluate_state(GameState* game_state) {
  double s = NAN;
  if (caaa_gpu_evaluate(C(game_state), 1, &s) != CAAA_OK) die("evaluate_state");
  return s;
}
bool is_terminal_state(GameState* game_state) {
  const double s = evaluate_state(game_state);
  return s > 0.99 || s < 0.01;
}
double random_play_until_terminal(GameState* game_state) {
  const double r = caaa_gpu_random_play_until_terminal(C(game_state));
  if (std::isnan(r)) die("random_play_until_terminal");
  return r;
}
void refresh_full_cache() {}  // no global caches to rebuild: every call derives them from the state it is given
void load_single_game() {
  std::fprintf(stderr, "caaa_gpu: load_single_game (console replay harness, reference src/engine.cpp:4250-4299) is not part of the GPU library\n");
  std::abort();
}
