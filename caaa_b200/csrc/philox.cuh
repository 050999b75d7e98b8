// philox.cuh -- counter-based per-game random byte stream (product side).
//
// Philox4x32-10 (Salmon, Moraes, Dror, Shaw, SC'11).  One playout's byte stream is
//   key = (seed_lo, seed_hi), counter = (draw_index >> 4, game_id_lo, game_id_hi, 0),
//   byte(draw_index) = byte (draw_index & 15) of the four 32-bit outputs, little-endian,
// so a trajectory is a pure function of (seed, game_id) and reproducible on any device count.
// It replaces the reference's 64 KiB rand() table + wrapping uint16 cursor
// (reference src/engine.cpp:98-110,516-520,1296,2562,2577,4202).
#pragma once
#include "slot.cuh"

namespace caaa {

CAAA_HD void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1,
                           uint32_t out[4]) {
#pragma unroll
  for (int r = 0; r < 10; r++) {
    const uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    const uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    c1 = (uint32_t)p1;
    c3 = (uint32_t)p0;
    c0 = n0;
    c2 = n2;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0;
  out[1] = c1;
  out[2] = c2;
  out[3] = c3;
}

}  // namespace caaa
