// root_allreduce.cpp -- the ONE exchange step of root-parallel MCTS (BASELINE config 5).
//
// Every GPU grows its own tree from the same root with its own seed; nothing is exchanged while searching.  At
// the end (or every K iterations) the root children's statistics -- int64 visits[n] and double values[n],
// n <= 20, i.e. <= 320 bytes -- are summed over the ranks with ncclAllReduce over NVLink / NVSwitch.  The
// message is latency-bound (a few microseconds); both reductions go out in one NCCL group.
//
// NCCL is bound at run time (dlopen of libnccl.so.2, reusing the copy a host program such as PyTorch has already
// loaded), so libcaaa_gpu.so itself has no load-time dependency on it: the library still loads, and the rollout
// path still runs, on a machine without NCCL.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/caaa_gpu.h"

namespace {

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;
std::string g_root_err;

struct RootComm {
  ncclComm_t comm = nullptr;
  cudaStream_t stream = nullptr;
  int64_t* d_visits = nullptr;
  double* d_values = nullptr;
  int world = 0, rank = 0;
} g_root;

int root_fail(int code, const std::string& what) {
  g_root_err = what;
  return code;
}

int load_nccl() {
  if (g_nccl.lib) return CAAA_OK;
  void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy the process already uses, if any
  if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!lib) return root_fail(CAAA_E_NO_DEVICE, std::string("NCCL is not available: ") + dlerror());
#define CAAA_SYM(field, name)                                                                   \
  *(void**)(&g_nccl.field) = dlsym(lib, name);                                                  \
  if (!g_nccl.field) return root_fail(CAAA_E_NO_DEVICE, std::string("NCCL symbol missing: ") + name)
  CAAA_SYM(GetUniqueId, "ncclGetUniqueId");
  CAAA_SYM(CommInitRank, "ncclCommInitRank");
  CAAA_SYM(CommDestroy, "ncclCommDestroy");
  CAAA_SYM(AllReduce, "ncclAllReduce");
  CAAA_SYM(GroupStart, "ncclGroupStart");
  CAAA_SYM(GroupEnd, "ncclGroupEnd");
  CAAA_SYM(GetErrorString, "ncclGetErrorString");
#undef CAAA_SYM
  g_nccl.lib = lib;
  return CAAA_OK;
}

#define NCCL_TRY(expr)                                                                             \
  do {                                                                                             \
    ncclResult_t r_ = (expr);                                                                      \
    if (r_ != ncclSuccess) return root_fail(CAAA_E_CUDA, std::string(#expr ": ") + g_nccl.GetErrorString(r_)); \
  } while (0)
#define CU_TRY(expr)                                                                               \
  do {                                                                                             \
    cudaError_t e_ = (expr);                                                                       \
    if (e_ != cudaSuccess) return root_fail(CAAA_E_CUDA, std::string(#expr ": ") + cudaGetErrorString(e_)); \
  } while (0)

}  // namespace

extern "C" {

const char* caaa_root_last_error(void) { return g_root_err.c_str(); }

int caaa_root_unique_id(uint8_t id128[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  if (!id128) return root_fail(CAAA_E_ARG, "null argument");
  int rc = load_nccl();
  if (rc != CAAA_OK) return rc;
  ncclUniqueId id;
  NCCL_TRY(g_nccl.GetUniqueId(&id));
  std::memcpy(id128, &id, 128);
  return CAAA_OK;
}

int caaa_root_comm_init(const uint8_t id128[128], int rank, int world) {
  if (!id128 || world < 1 || rank < 0 || rank >= world) return root_fail(CAAA_E_ARG, "bad communicator arguments");
  if (g_root.comm) return root_fail(CAAA_E_ARG, "root communicator already initialised");
  int rc = load_nccl();
  if (rc != CAAA_OK) return rc;
  ncclUniqueId id;
  std::memcpy(&id, id128, 128);
  NCCL_TRY(g_nccl.CommInitRank(&g_root.comm, world, id, rank));  // on the calling thread's current device
  CU_TRY(cudaStreamCreateWithFlags(&g_root.stream, cudaStreamNonBlocking));
  CU_TRY(cudaMalloc(&g_root.d_visits, CAAA_ACTIONS * sizeof(int64_t)));
  CU_TRY(cudaMalloc(&g_root.d_values, CAAA_ACTIONS * sizeof(double)));
  g_root.world = world;
  g_root.rank = rank;
  return CAAA_OK;
}

int caaa_root_comm_destroy(void) {
  if (!g_root.comm) return CAAA_OK;
  cudaStreamSynchronize(g_root.stream);
  g_nccl.CommDestroy(g_root.comm);
  cudaFree(g_root.d_visits);
  cudaFree(g_root.d_values);
  cudaStreamDestroy(g_root.stream);
  g_root = RootComm();
  return CAAA_OK;
}

// sum the root statistics over the ranks of `nccl_comm` (an ncclComm_t), in place, device pointers,
// asynchronous on `cuda_stream`.  Both reductions are issued as one NCCL group (one launch).
int caaa_root_allreduce(void* nccl_comm, int n_children, int64_t* d_visits, double* d_values, void* cuda_stream) {
  if (!nccl_comm || !d_visits || !d_values || n_children < 0 || n_children > CAAA_ACTIONS) return root_fail(CAAA_E_ARG, "bad all-reduce arguments");
  int rc = load_nccl();
  if (rc != CAAA_OK) return rc;
  if (n_children == 0) return CAAA_OK;
  NCCL_TRY(g_nccl.GroupStart());
  NCCL_TRY(g_nccl.AllReduce(d_visits, d_visits, (size_t)n_children, ncclInt64, ncclSum, (ncclComm_t)nccl_comm, (cudaStream_t)cuda_stream));
  NCCL_TRY(g_nccl.AllReduce(d_values, d_values, (size_t)n_children, ncclFloat64, ncclSum, (ncclComm_t)nccl_comm, (cudaStream_t)cuda_stream));
  NCCL_TRY(g_nccl.GroupEnd());
  return CAAA_OK;
}

// Host-buffer form on the library's own communicator (caaa_root_comm_init): H2D, the grouped all-reduce, D2H.
// elapsed_us (optional) = wall time of the call on this rank.
int caaa_root_allreduce_host(int n_children, int64_t* visits, double* values, double* elapsed_us) {
  if (!g_root.comm) return root_fail(CAAA_E_NOT_INIT, "caaa_root_comm_init has not been called");
  if (!visits || !values || n_children < 0 || n_children > CAAA_ACTIONS) return root_fail(CAAA_E_ARG, "bad all-reduce arguments");
  const auto t0 = std::chrono::steady_clock::now();
  CU_TRY(cudaMemcpyAsync(g_root.d_visits, visits, (size_t)n_children * sizeof(int64_t), cudaMemcpyHostToDevice, g_root.stream));
  CU_TRY(cudaMemcpyAsync(g_root.d_values, values, (size_t)n_children * sizeof(double), cudaMemcpyHostToDevice, g_root.stream));
  int rc = caaa_root_allreduce(g_root.comm, n_children, g_root.d_visits, g_root.d_values, g_root.stream);
  if (rc != CAAA_OK) return rc;
  CU_TRY(cudaMemcpyAsync(visits, g_root.d_visits, (size_t)n_children * sizeof(int64_t), cudaMemcpyDeviceToHost, g_root.stream));
  CU_TRY(cudaMemcpyAsync(values, g_root.d_values, (size_t)n_children * sizeof(double), cudaMemcpyDeviceToHost, g_root.stream));
  CU_TRY(cudaStreamSynchronize(g_root.stream));
  if (elapsed_us) *elapsed_us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
  return CAAA_OK;
}

}  // extern "C"
