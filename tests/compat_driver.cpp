// compat_driver.cpp -- TEST: drives libcaaa_gpu.so through the reference's own API names
// (include/caaa_engine_compat.h), in the order the reference's main.cpp / mcts.cpp use them
// (reference src/main.cpp:16-21, src/mcts.cpp:84-94,110-126), and prints what it gets.
#include <cstdio>
#include <cstdlib>

#include "../include/caaa_engine_compat.h"
#include "../include/caaa_gpu.h"

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  initialize_constants();
  load_game_data(argv[1]);
  GameState* root = get_game_state_copy();
  uint8_t n = 0;
  Actions acts = {0};
  get_possible_actions(root, &n, &acts);
  std::printf("actions");
  for (int i = 0; i < n; i++) std::printf(" %d", acts[i]);
  std::printf("\n");
  for (int i = 0; i < n; i++) {
    GameState* child = clone_state(root);
    apply_action(child, acts[i]);
    std::printf("child %d player %d money0 %d terminal %d\n", acts[i], child->player_index, child->money[0], (int)is_terminal_state(child));
    free_state(child);
  }
  caaa_gpu_set_stream(0xCAAA2024ull, 0);
  for (int i = 0; i < 4; i++) std::printf("rollout %.17g\n", random_play_until_terminal(root));
  std::printf("evaluate %.17g\n", evaluate_state(root));
  free_state(root);
  return 0;
}
