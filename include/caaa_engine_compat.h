// caaa_engine_compat.h -- the reference's MCTS-facing engine API, with the reference's own names,
// signatures and C++ linkage (reference include/engine.h:29,75,150-159 and src/mcts.cpp:16-22), so
// that the reference's mcts.cpp / main.cpp link against libcaaa_gpu.so unchanged.
// Every call forwards to the C-ABI in caaa_gpu.h / caaa_host.h; there is no CPU implementation behind it.
#pragma once
#include "caaa_types.h"

typedef CaaaGameState GameState;
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
