"""Philox4x32-10: both implementations (oracle side and product side) against the Random123
known-answer vectors, and the per-playout byte stream definition."""
import ctypes as C

import numpy as np
import pytest

# Random123 kat_vectors, philox4x32 10 rounds: (counter, key) -> output
KAT = [
    ((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
    ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
    ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
     (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)),
]


def _philox(lib, name, ctr, key):
    fn = getattr(lib, name)
    c = (C.c_uint32 * 4)(*ctr)
    k = (C.c_uint32 * 2)(*key)
    o = (C.c_uint32 * 4)()
    fn(c, k, o)
    return tuple(o)


@pytest.mark.parametrize("which", ["port", "hostsim"])
def test_known_answers(which, port, hostsim):
    eng = port if which == "port" else hostsim
    name = "caaa_oracle_philox" if which == "port" else "caaa_hostsim_philox"
    for ctr, key, out in KAT:
        assert _philox(eng.lib, name, ctr, key) == out


def test_stream_bytes_agree(port, hostsim):
    fp, fh = port.lib.caaa_oracle_stream_byte, hostsim.lib.caaa_hostsim_stream_byte
    for f in (fp, fh):
        f.restype = C.c_uint8
        f.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32]
    rng = np.random.default_rng(1)
    for _ in range(300):
        seed = int(rng.integers(0, 2**63))
        gid = int(rng.integers(0, 2**63))
        d = int(rng.integers(0, 100000))
        assert fp(seed, gid, d) == fh(seed, gid, d)
    # byte k of block j is byte (k & 3) of output word k >> 2, little-endian
    ctr, key = (5, 7, 0, 0), (0xCAAA2024, 0)
    out = _philox(port.lib, "caaa_oracle_philox", ctr, key)
    for k in range(16):
        assert fp(0xCAAA2024, 7, 5 * 16 + k) == (out[k >> 2] >> (8 * (k & 3))) & 0xff
