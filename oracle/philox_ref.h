/* philox_ref.h -- TEST INFRASTRUCTURE (oracle side).  Independent restatement of
 * Philox4x32-10 (Salmon et al., "Parallel random numbers: as easy as 1, 2, 3", SC'11) used by
 * the patched reference (ref_shim.cpp) and the C oracle.  The product has its own copy in
 * caaa_b200/csrc/philox.cuh; both are pinned against the Random123 known-answer vectors in
 * tests/test_philox.py.
 *
 * Byte stream of one playout: key = (seed_lo, seed_hi); counter = (draw_index >> 4,
 * game_id_lo, game_id_hi, 0); the byte for draw_index is byte (draw_index & 15) of the four
 * 32-bit outputs laid out little-endian.
 */
#ifndef CAAA_PHILOX_REF_H
#define CAAA_PHILOX_REF_H
#include <stdint.h>

static inline void caaa_philox4x32_10_ref(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
  uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
  uint32_t k0 = key[0], k1 = key[1];
  for (int round = 0; round < 10; round++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

static inline uint8_t caaa_philox_byte_ref(uint64_t seed, uint64_t game_id, uint32_t draw_index) {
  uint32_t ctr[4] = {draw_index >> 4, (uint32_t)game_id, (uint32_t)(game_id >> 32), 0u};
  uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  uint32_t out[4];
  caaa_philox4x32_10_ref(ctr, key, out);
  uint32_t k = draw_index & 15u;
  return (uint8_t)(out[k >> 2] >> (8u * (k & 3u)));
}
#endif
