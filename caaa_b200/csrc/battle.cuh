// battle.cuh -- isolated combat resolution (BASELINE config 3): resolve_sea_battles + resolve_land_battles
// (reference src/engine.cpp:2312, 3055) on a batch of independent states.
//
// Included by caaa_gpu.cu.  One thread per battle, one warp per tile of 32 consecutive states.  The states are an
// array of 670-byte records at the ABI, so a tile is 21 440 contiguous bytes: the warp moves it between global and
// shared memory with ONE bulk asynchronous copy each way (cp.async.bulk, completion on an mbarrier / bulk group;
// SASS UBLKCP), and every lane imports / exports its own record from shared memory.  HBM sees exactly the
// algorithmic bytes (670 in + 670 out per battle, fully coalesced); the byte-wise layout conversion and the
// branchy combat code run out of shared memory.
#pragma once

namespace caaa {

constexpr int BATTLE_WARPS = 4;
constexpr int TILE_STATES = 32;
constexpr int TILE_BYTES = TILE_STATES * CAAA_STATE_BYTES;  // 21 440, a multiple of 16
static_assert(TILE_BYTES % 16 == 0, "a tile must be a whole number of 16-byte chunks");
constexpr int BATTLE_WARP_BYTES = TILE_BYTES + TILE_STATES * (int)sizeof(Game);
constexpr int BATTLE_SMEM = TABLE_BYTES + BATTLE_WARPS * BATTLE_WARP_BYTES + 16 * BATTLE_WARPS;
static_assert(BATTLE_SMEM <= 232448, "battle kernel: shared memory budget");

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* mbar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(mbar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* mbar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(mbar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* mbar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n"
      "DONE_%=:\n\t}" ::"r"(smem_u32(mbar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gmem_src, unsigned bytes, void* mbar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gmem_src), "r"(bytes), "r"(smem_u32(mbar))
               : "memory");
}
__device__ __forceinline__ void bulk_store(void* gmem_dst, const void* smem_src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_proxy() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__global__ void __launch_bounds__(BATTLE_WARPS * 32, 1) battle_kernel(const DevTables* tables, const uint8_t* states, int n,
                                                                      unsigned long long seed, unsigned long long first_game_id,
                                                                      uint8_t* out_states, int32_t* out_draws, CaaaBatchSummary* summary) {
  const DevTables* T = stage_tables(tables);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint8_t* const wbase = smem_raw + TABLE_BYTES + warp * BATTLE_WARP_BYTES;
  uint8_t* const raw = wbase;                                             // the tile in the reference layout (16-byte aligned)
  Game* const g = reinterpret_cast<Game*>(wbase + TILE_BYTES) + lane;     // this lane's working set
  unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(smem_raw + TABLE_BYTES + BATTLE_WARPS * BATTLE_WARP_BYTES) + 2 * warp;
  if (lane == 0) mbar_init(mbar, 1);
  fence_async_proxy();
  __syncwarp();
  const RunCtx cx{T, (uint32_t)seed, (uint32_t)(seed >> 32)};
  const int n_tiles = (n + TILE_STATES - 1) / TILE_STATES;
  const int warps_total = (int)gridDim.x * BATTLE_WARPS;
  unsigned parity = 0;
  unsigned long long draws = 0, flagged = 0, one = 0;
  // 16-byte alignment of both arrays is what the bulk copies need (device allocations are 256-byte aligned)
  const bool aligned = ((reinterpret_cast<uintptr_t>(states) | reinterpret_cast<uintptr_t>(out_states)) & 15u) == 0;
  for (int tile = (int)blockIdx.x * BATTLE_WARPS + warp; tile < n_tiles; tile += warps_total) {
    const int base = tile * TILE_STATES;
    const int cnt = n - base < TILE_STATES ? n - base : TILE_STATES;
    const unsigned bytes = (unsigned)cnt * CAAA_STATE_BYTES;
    const uint8_t* src = states + (size_t)base * CAAA_STATE_BYTES;
    const bool bulk = aligned && (bytes & 15u) == 0;
    // ---- tile in ----
    if (bulk) {
      if (lane == 0) {
        mbar_expect_tx(mbar, bytes);
        bulk_load(raw, src, bytes, mbar);
      }
      mbar_wait(mbar, parity);
      parity ^= 1u;
    } else {  // ragged last tile / unaligned caller buffer
      for (unsigned i = lane; i < bytes; i += 32) raw[i] = src[i];
      __syncwarp();
    }
    // ---- one battle per lane ----
    const bool mine = lane < cnt;
    const unsigned m = __ballot_sync(FULL, mine);
    if (mine) {
      import_state(g, raw + lane * CAAA_STATE_BYTES);
      begin_playout(g, cx, first_game_id + (unsigned long long)(base + lane));
      resolve_sea_battles<false>(g, cx, m);
      resolve_land_battles<false>(g, cx, m);
      if (out_states) export_state(g, raw + lane * CAAA_STATE_BYTES);
      if (out_draws) out_draws[base + lane] = (int32_t)g->draws;
      draws += g->draws;
      flagged += g->status != 0;
      one += 1;
    }
    __syncwarp();
    // ---- tile out ----
    if (out_states) {
      uint8_t* dst = out_states + (size_t)base * CAAA_STATE_BYTES;
      if (bulk) {
        fence_async_proxy();  // the lanes' generic-proxy writes of the tile before the async-proxy read
        __syncwarp();
        if (lane == 0) {
          bulk_store(dst, raw, bytes);
          bulk_commit();
          bulk_wait_read();   // the tile buffer is reused by the next load
        }
        __syncwarp();
      } else {
        for (unsigned i = lane; i < bytes; i += 32) dst[i] = raw[i];
        __syncwarp();
      }
    }
  }
  if (summary != nullptr) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      draws += __shfl_down_sync(FULL, draws, o);
      flagged += __shfl_down_sync(FULL, flagged, o);
      one += __shfl_down_sync(FULL, one, o);
    }
    if (lane == 0 && one) {
      atomicAdd((unsigned long long*)&summary->playouts, one);
      atomicAdd((unsigned long long*)&summary->draws, draws);
      atomicAdd((unsigned long long*)&summary->flagged, flagged);
    }
  }
}

}  // namespace caaa
