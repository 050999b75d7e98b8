// soa_ab.cu -- A/B of the game-pool layout (tools; not part of the product).  nvcc -O3 -gencode arch=compute_100a,code=sm_100a
//
// The phase-regrouped rollout kernel moves games between an L2-resident pool and shared memory: a warp claims 16
// pool entries that wait for the same phase -- an ARBITRARY subset of the pool -- copies them in, and copies them
// back out.  This measures exactly that traffic pattern (16 random entries in, the same 16 out, per warp, 15 warps
// per SM, 148 SMs, 66 k entries of 932 bytes) for three layouts of the pool:
//   aos32    array of 932-byte records, 32-bit accesses, all 32 lanes on one record at a time (what ships)
//   aos128   records padded to 944 bytes (16-byte aligned), 128-bit accesses
//   soa128   north_star's chunk-major structure of arrays: 59 chunks of 16 bytes per game, chunk c of entry i at
//            (c * n + i) * 16; 128-bit accesses, one lane per (entry, chunk)
// Output: microseconds per batch of 16 games (in + out) and the pool bandwidth it corresponds to.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int GAME_BYTES = 932, GAME_WORDS = 233, CHUNKS = 59, LANES = 16, WARPS = 15;
extern __shared__ __align__(16) unsigned char smem[];

__global__ void __launch_bounds__(WARPS * 32, 1) k_aos32(unsigned* pool, const unsigned short* idx, int batches_per_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned* slots = reinterpret_cast<unsigned*>(smem) + warp * LANES * GAME_WORDS;
  const unsigned short* my = idx + ((size_t)blockIdx.x * WARPS + warp) * batches_per_warp * LANES;
  for (int b = 0; b < batches_per_warp; b++) {
    for (int j0 = 0; j0 < LANES; j0 += 4) {
      unsigned r[4][8];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const unsigned* src = pool + (size_t)my[b * LANES + j0 + jj] * GAME_WORDS;
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k * 32 + lane < GAME_WORDS) r[jj][k] = __ldcg(src + k * 32 + lane);
      }
#pragma unroll
      for (int jj = 0; jj < 4; jj++)
#pragma unroll
        for (int k = 0; k < 8; k++)
          if (k * 32 + lane < GAME_WORDS) slots[(j0 + jj) * GAME_WORDS + k * 32 + lane] = r[jj][k] + 1u;
    }
    __syncwarp();
    for (int j = 0; j < LANES; j++) {
      unsigned* dst = pool + (size_t)my[b * LANES + j] * GAME_WORDS;
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (k * 32 + lane < GAME_WORDS) __stcg(dst + k * 32 + lane, slots[j * GAME_WORDS + k * 32 + lane]);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(WARPS * 32, 1) k_aos128(uint4* pool, const unsigned short* idx, int batches_per_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* slots = reinterpret_cast<uint4*>(smem) + warp * LANES * CHUNKS;
  const unsigned short* my = idx + ((size_t)blockIdx.x * WARPS + warp) * batches_per_warp * LANES;
  for (int b = 0; b < batches_per_warp; b++) {
    for (int j0 = 0; j0 < LANES; j0 += 4) {
      uint4 r[4][2];
#pragma unroll
      for (int jj = 0; jj < 4; jj++) {
        const uint4* src = pool + (size_t)my[b * LANES + j0 + jj] * CHUNKS;
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (k * 32 + lane < CHUNKS) r[jj][k] = __ldcg(src + k * 32 + lane);
      }
#pragma unroll
      for (int jj = 0; jj < 4; jj++)
#pragma unroll
        for (int k = 0; k < 2; k++)
          if (k * 32 + lane < CHUNKS) {
            r[jj][k].x += 1u;
            slots[(j0 + jj) * CHUNKS + k * 32 + lane] = r[jj][k];
          }
    }
    __syncwarp();
    for (int j = 0; j < LANES; j++) {
      uint4* dst = pool + (size_t)my[b * LANES + j] * CHUNKS;
#pragma unroll
      for (int k = 0; k < 2; k++)
        if (k * 32 + lane < CHUNKS) __stcg(dst + k * 32 + lane, slots[j * CHUNKS + k * 32 + lane]);
    }
    __syncwarp();
  }
}

__global__ void __launch_bounds__(WARPS * 32, 1) k_soa128(uint4* pool, int n_entries, const unsigned short* idx, int batches_per_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint4* slots = reinterpret_cast<uint4*>(smem) + warp * LANES * CHUNKS;
  const unsigned short* my = idx + ((size_t)blockIdx.x * WARPS + warp) * batches_per_warp * LANES;
  for (int b = 0; b < batches_per_warp; b++) {
    // LANES * CHUNKS = 944 (entry, chunk) pairs; pair p = c * 16 + j: lanes 0..15 fetch chunk c of the 16 entries
    for (int p0 = 0; p0 < LANES * CHUNKS; p0 += 32 * 8) {
      uint4 r[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int p = p0 + u * 32 + lane;
        if (p < LANES * CHUNKS) r[u] = __ldcg(pool + (size_t)(p / LANES) * n_entries + my[b * LANES + p % LANES]);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int p = p0 + u * 32 + lane;
        if (p < LANES * CHUNKS) {
          r[u].x += 1u;
          slots[(p % LANES) * CHUNKS + p / LANES] = r[u];
        }
      }
    }
    __syncwarp();
    for (int p = lane; p < LANES * CHUNKS; p += 32)
      __stcg(pool + (size_t)(p / LANES) * n_entries + my[b * LANES + p % LANES], slots[(p % LANES) * CHUNKS + p / LANES]);
    __syncwarp();
  }
}

int main() {
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int n_entries = 148 * 240 * 2 < 65536 ? 148 * 240 * 2 : 65535;  // 16-bit indices
  const int bpw = 400;
  const size_t n_idx = (size_t)sms * WARPS * bpw * LANES;
  std::vector<unsigned short> h(n_idx);
  unsigned long long s = 88172645463325252ull;
  for (size_t i = 0; i < n_idx; i += LANES) {  // 16 distinct random entries per batch
    for (int j = 0; j < LANES; j++) {
      bool dup;
      unsigned short v;
      do {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        v = (unsigned short)(s % n_entries);
        dup = false;
        for (int q = 0; q < j; q++) dup |= h[i + q] == v;
      } while (dup);
      h[i + j] = v;
    }
  }
  unsigned short* d_idx;
  void* d_pool;
  cudaMalloc(&d_idx, n_idx * 2);
  cudaMemcpy(d_idx, h.data(), n_idx * 2, cudaMemcpyHostToDevice);
  cudaMalloc(&d_pool, (size_t)n_entries * 944 + 1024);
  cudaMemset(d_pool, 0, (size_t)n_entries * 944 + 1024);
  const int smem_bytes = WARPS * LANES * 944;
  cudaFuncSetAttribute(k_aos32, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_aos128, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaFuncSetAttribute(k_soa128, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const char* names[3] = {"aos32", "aos128", "soa128"};
  for (int v = 0; v < 3; v++) {
    float best = 1e30f;
    for (int rep = 0; rep < 4; rep++) {
      cudaEventRecord(e0);
      if (v == 0) k_aos32<<<sms, WARPS * 32, smem_bytes>>>((unsigned*)d_pool, d_idx, bpw);
      if (v == 1) k_aos128<<<sms, WARPS * 32, smem_bytes>>>((uint4*)d_pool, d_idx, bpw);
      if (v == 2) k_soa128<<<sms, WARPS * 32, smem_bytes>>>((uint4*)d_pool, n_entries, d_idx, bpw);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    const double batches = (double)sms * WARPS * bpw;
    printf("{\"layout\": \"%s\", \"ms\": %.3f, \"us_per_batch_per_warp\": %.3f, \"pool_GBps\": %.1f, \"status\": \"%s\"}\n", names[v], best,
           best * 1e3 / bpw, batches * LANES * GAME_BYTES * 2 / (best * 1e-3) / 1e9, cudaGetErrorString(e));
  }
  return 0;
}
