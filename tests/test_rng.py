"""Counter-based RNG: Random123 known-answer vectors for Philox4x32-10 and agreement of the three
CPU statements of the draw mapping (oracle/philox_ref.py spec, oracle C)."""
import ctypes as C
import math
import random

import numpy as np

import philox_ref
from oracle import lib

KAT = [  # Random123 kat_vectors, philox4x32 10
    ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]


def c_philox(ctr, key):
    c = (C.c_uint32 * 4)(*ctr); k = (C.c_uint32 * 2)(*key); o = (C.c_uint32 * 4)()
    lib().orc_philox4x32_10(c, k, o)
    return tuple(o)


def test_philox_kat():
    for ctr, key, out in KAT:
        assert philox_ref.philox4x32_10(ctr, key) == out
        assert c_philox(ctr, key) == out


def test_philox_c_matches_spec_random():
    rnd = random.Random(0)
    for _ in range(2000):
        ctr = tuple(rnd.getrandbits(32) for _ in range(4)); key = tuple(rnd.getrandbits(32) for _ in range(2))
        assert c_philox(ctr, key) == philox_ref.philox4x32_10(ctr, key)


def test_neglog_bit_exact_and_accurate():
    rnd = random.Random(1)
    ws = [0, 1, 2, 2**31 - 1, 2**31, 2**32 - 2, 2**32 - 1] + [rnd.getrandbits(32) for _ in range(20000)]
    for w in ws:
        a = philox_ref.neglog_u32(w); b = lib().orc_neglog_u32(w)
        assert a == b, w
        exact = -math.log((w + 1) / 2**32)
        assert abs(a - exact) <= 1e-15 * max(exact, 1e-300) + 1e-300 or abs(a - exact) < 2e-16
