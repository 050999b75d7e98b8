// caaa_engine_compat.h -- the reference's MCTS-facing engine API, with the reference's own names,
// signatures and C++ linkage (reference include/engine.h:29,75,85,150-159, src/mcts.cpp:16-22,
// src/main.cpp:16-21,38-39), so that the reference's own mcts.cpp / main.cpp objects link against
// libcaaa_gpu.so (tests/test_reference_links.py does exactly that).
// Every call forwards to the C-ABI in caaa_gpu.h / caaa_host.h; there is no CPU implementation behind it.
//
// The reference declares `typedef struct {...} GameState;` (include/game_state.h:59-68): an unnamed
// struct whose typedef name is its linkage name, so `random_play_until_terminal(GameState*)` mangles
// as _Z26random_play_until_terminalP9GameState.  This header declares GameState the same way (same
// 670 bytes as CaaaGameState) so that the library exports exactly those symbols.
#pragma once
#include "caaa_types.h"

typedef struct {
  CAAA_GAME_STATE_MEMBERS
} GameState;
typedef uint8_t Actions[CAAA_ACTIONS];
typedef Actions* ActionsPtr;

void initialize_constants();                 // -> caaa_gpu_init(device from $CAAA_DEVICE, default 0)
void load_game_data(char const* filename);   // -> caaa_state_from_json_file into the library's current state
GameState* get_game_state_copy();            // malloc'ed copy of the current state (reference src/engine.cpp:4031)
GameState* clone_state(GameState* game_state);
void free_state(GameState* game_state);
void get_possible_actions(GameState* game_state, uint8_t* num_actions, ActionsPtr actions);
void apply_action(GameState* game_state, uint8_t action);
bool is_terminal_state(GameState* game_state);
double evaluate_state(GameState* game_state);
double random_play_until_terminal(GameState* game_state);
// referenced only by the reference's dead `main2` (src/main.cpp:35-42): the library keeps no global derived
// caches, so refresh_full_cache() has nothing to do; load_single_game() (the console replay harness,
// src/engine.cpp:4250-4299, out of scope) reports that and aborts.
void refresh_full_cache();
void load_single_game();
