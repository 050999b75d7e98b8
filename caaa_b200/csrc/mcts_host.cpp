// mcts_host.cpp -- host-side UCT search that drives the GPU engine (product code, plain C++).
//
// Mirrors the reference's search loop (reference src/mcts.cpp:29-126,247-258): UCT with
// EXPLORATION_CONSTANT 1.414 and +1 smoothing (mcts.cpp:24,47-63), expansion of every visited
// leaf (mcts.cpp:84-87,110-126), a uniformly random child, rollout, and back-propagation in which a
// node accumulates `result` when its parent's player_index is even and `1 - result` otherwise
// (mcts.cpp:96-105).  What changes is where the work runs: expansion (get_possible_actions +
// apply_action per action) and rollouts are batched device calls, `rollouts_per_leaf` playouts are
// played per visited leaf (BASELINE config 4: 4096), `leaves_per_batch` leaves are selected per launch with
// virtual visits, and two such batches are kept in flight so one host thread keeps the GPU busy.  With both set to 1 the loop is
// the reference's, except that is_terminal_state is evaluated on the node's own unit totals (the
// reference reads stale globals there) and the random stream is Philox.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cstring>
#include <deque>
#include <random>
#include <vector>

#include "../../include/caaa_gpu.h"
#include "../../include/caaa_host.h"

namespace {

constexpr double kExploration = 1.414;  // reference src/mcts.cpp:24

struct Node {
  CaaaGameState state;
  uint8_t action = 0;
  Node* parent = nullptr;
  std::vector<Node*> children;
  int64_t visits = 0;
  int64_t pending = 0;  // virtual visits of leaves selected in the current batch
  double value = 0.0;
  bool in_batch = false;
};

struct Tree {
  std::deque<Node> pool;  // stable addresses
  Node* make(const CaaaGameState& s, uint8_t action, Node* parent) {
    pool.emplace_back();
    Node* n = &pool.back();
    n->state = s;
    n->action = action;
    n->parent = parent;
    return n;
  }
};

Node* select_best_leaf(Node* node) {  // reference src/mcts.cpp:47-63, with virtual visits
  while (!node->children.empty()) {
    double best = -INFINITY;
    Node* best_child = nullptr;
    const double n_parent = (double)(node->visits + node->pending);
    for (Node* c : node->children) {
      const double n = (double)(c->visits + c->pending);
      const double uct = c->value / (n + 1) + kExploration * std::sqrt(std::log(n_parent + 1) / (n + 1));
      if (uct > best) {
        best = uct;
        best_child = c;
      }
    }
    node = best_child;
  }
  return node;
}

void backpropagate(Node* node, double sum_result, int64_t n) {  // reference src/mcts.cpp:96-105
  while (node != nullptr) {
    node->visits += n;
    if (node->parent != nullptr && node->parent->state.player_index % 2 == 0) node->value += sum_result;
    else node->value += (double)n - sum_result;
    node = node->parent;
  }
}

}  // namespace

// Pipelined driver: up to `depth` leaf batches are in flight (caaa_gpu_leaf_batch_submit / _collect).  While batch k rolls
// out on the device, the host back-propagates batch k-1 and selects batch k+1 with virtual visits standing in for the
// results that are still outstanding; expansion, the random child and the per-leaf sums all stay on the device.
extern "C" int caaa_mcts_search_pipelined(const CaaaGameState* root_state, int64_t iterations, int rollouts_per_leaf,
                                          int leaves_per_batch, int depth, uint64_t seed, CaaaRootStats* out, uint8_t* best_action) {
  if (!root_state || !out || iterations < 0 || rollouts_per_leaf <= 0 || leaves_per_batch <= 0 || depth < 1 || depth > 3) return CAAA_E_ARG;
  Tree tree;
  Node* root = tree.make(*root_state, 0, nullptr);
  std::mt19937_64 rng(seed ^ 0x9E3779B97F4A7C15ull);
  uint64_t next_game_id = 0;
  int64_t done = 0, selected = 0, playouts = 0;
  struct Flight {
    int handle;
    std::vector<Node*> leaves;
  };
  std::deque<Flight> flights;
  std::vector<int> free_handles;
  for (int h = depth - 1; h >= 0; h--) free_handles.push_back(h);
  std::vector<CaaaGameState> states, children;
  std::vector<double> scores, sums;
  std::vector<uint8_t> n_actions, actions;
  std::vector<uint32_t> status;
  std::vector<int32_t> picked;
  std::vector<uint64_t> picks;
  int sm_count = 0;
  if (depth > 1 && caaa_gpu_device_info(&sm_count, nullptr, nullptr) == CAAA_OK && sm_count > 32)
    caaa_gpu_set_rollout_sms((sm_count - 4) / 16 * 16);  // a few SMs stay free for the other batch's expansion kernels
  int rc = CAAA_OK;
  const bool trace = std::getenv("CAAA_MCTS_TRACE") != nullptr;  // per-batch host timeline on stderr
  const auto t_start = std::chrono::steady_clock::now();
  auto ms_now = [&] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_start).count(); };
  while (rc == CAAA_OK && (selected < iterations || !flights.empty())) {
    // 1. keep `depth` batches in flight: select distinct leaves, marking virtual visits along their paths
    while (!free_handles.empty() && selected < iterations) {
      Flight f;
      const int64_t want = std::min<int64_t>(leaves_per_batch, iterations - selected);
      for (int64_t b = 0; b < want; b++) {
        Node* leaf = select_best_leaf(root);
        if (leaf->in_batch) break;  // the tree offers no further distinct leaf until a batch in flight has been collected
        leaf->in_batch = true;
        for (Node* n = leaf; n != nullptr; n = n->parent) n->pending += rollouts_per_leaf;  // one visit = rollouts_per_leaf playouts
        f.leaves.push_back(leaf);
      }
      const int nl = (int)f.leaves.size();
      if (nl == 0) break;
      states.resize(nl);
      picks.resize(nl);
      for (int i = 0; i < nl; i++) {
        states[i] = f.leaves[i]->state;
        picks[i] = rng();  // `rand() % num_children`, reference src/mcts.cpp:86 (the modulo is taken on the device)
      }
      f.handle = free_handles.back();
      if (trace) std::fprintf(stderr, "mcts %9.2f ms  submit  handle %d  %d leaves\n", ms_now(), f.handle, nl);
      rc = caaa_gpu_leaf_batch_submit(f.handle, states.data(), nl, rollouts_per_leaf, seed, next_game_id, picks.data());
      if (rc != CAAA_OK) break;
      free_handles.pop_back();
      next_game_id += (uint64_t)nl * (uint64_t)rollouts_per_leaf;
      selected += nl;
      flights.push_back(std::move(f));
    }
    if (rc != CAAA_OK || flights.empty()) break;
    // 2. collect the oldest batch: grow the tree, clear its virtual visits, back-propagate the per-leaf sums
    Flight f = std::move(flights.front());
    flights.pop_front();
    const int nl = (int)f.leaves.size();
    scores.resize(nl);
    sums.resize(nl);
    n_actions.resize(nl);
    actions.resize((size_t)nl * CAAA_ACTIONS);
    children.resize((size_t)nl * CAAA_ACTIONS);
    status.resize(nl);
    picked.resize(nl);
    if (trace) std::fprintf(stderr, "mcts %9.2f ms  wait    handle %d\n", ms_now(), f.handle);
    rc = caaa_gpu_leaf_batch_collect(f.handle, scores.data(), n_actions.data(), actions.data(), children.data(), status.data(),
                                     picked.data(), sums.data(), nullptr);
    if (trace) std::fprintf(stderr, "mcts %9.2f ms  collect handle %d  %d leaves\n", ms_now(), f.handle, nl);
    free_handles.push_back(f.handle);
    if (rc != CAAA_OK) break;
    for (int i = 0; i < nl; i++) {
      Node* leaf = f.leaves[i];
      Node* target = leaf;
      const bool terminal = scores[i] > 0.99 || scores[i] < 0.01;  // reference src/engine.cpp:4244-4248
      if (!terminal && !(status[i] & 0x80000000u)) {
        for (int k = 0; k < n_actions[i]; k++) {
          if (status[i] & (1u << k)) continue;  // the reference would never return from this apply_action
          Node* child = tree.make(children[(size_t)i * CAAA_ACTIONS + k], actions[(size_t)i * CAAA_ACTIONS + k], leaf);
          leaf->children.push_back(child);
          if (k == picked[i]) target = child;
        }
      }
      for (Node* n = leaf; n != nullptr; n = n->parent) n->pending -= rollouts_per_leaf;
      leaf->in_batch = false;
      backpropagate(target, sums[i], rollouts_per_leaf);
    }
    done += nl;
    playouts += (int64_t)nl * rollouts_per_leaf;
  }
  if (depth > 1) caaa_gpu_set_rollout_sms(0);
  if (rc != CAAA_OK) {
    for (const Flight& f : flights) caaa_gpu_leaf_batch_collect(f.handle, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
    return rc;
  }
  std::memset(out, 0, sizeof(*out));
  out->n_children = (int32_t)std::min<size_t>(root->children.size(), CAAA_ACTIONS);
  double best = -INFINITY;
  uint8_t best_a = 255;
  for (int i = 0; i < out->n_children; i++) {
    const Node* c = root->children[i];
    out->actions[i] = c->action;
    out->visits[i] = c->visits;
    out->values[i] = c->value;
    if (c->value > best) {  // select_best_action, reference src/mcts.cpp:247-258
      best = c->value;
      best_a = c->action;
    }
  }
  out->iterations = done;
  out->playouts = playouts;
  out->nodes = (int64_t)tree.pool.size();
  if (best_action) *best_action = best_a;
  return CAAA_OK;
}

// leaves_per_batch == 1 is the reference's sequential loop (every selection sees every earlier result); wider
// batches keep two in flight.
extern "C" int caaa_mcts_search(const CaaaGameState* root_state, int64_t iterations, int rollouts_per_leaf,
                                int leaves_per_batch, uint64_t seed, CaaaRootStats* out, uint8_t* best_action) {
  return caaa_mcts_search_pipelined(root_state, iterations, rollouts_per_leaf, leaves_per_batch, leaves_per_batch > 1 ? 2 : 1, seed, out,
                                    best_action);
}
