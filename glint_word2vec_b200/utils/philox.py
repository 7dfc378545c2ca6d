"""Counter-based RNG shared by the CPU oracle and the sm_100a kernels.

Every random decision of training (sub-sampling, window radius, negative
samples, synthetic Zipf tokens, weight init) is a pure function of
``(seed, stream, position, sub)`` evaluated with Philox4x32-10.  Because the
function is pure, every column shard (rank) regenerates the *same* negatives
and windows without any index ever crossing NVLink -- the property that makes
the Ordentlich et al. scheme network efficient (reference: the seed-only
``dotprod`` request at MLLIB:420-421).

This module is the numpy replica of ``csrc/philox.cuh``; the two must stay
bit-identical (tests/test_philox.py, and the GPU numerics tests rely on it).
"""
from __future__ import annotations

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85

# stream ids (low 8 bits of counter word 3; the training iteration lives above)
STREAM_SUBSAMPLE = 1
STREAM_WINDOW = 2
STREAM_NEG = 3
STREAM_ZIPF = 4
STREAM_INIT = 5

_MASK32 = np.uint64(0xFFFFFFFF)


def philox4x32(c0, c1, c2, c3, k0, k1, rounds: int = 10):
    """Vectorised Philox4x32.  All inputs broadcastable uint32-valued arrays.

    Returns four uint32 arrays.
    """
    c0 = np.asarray(c0, dtype=np.uint64) & _MASK32
    c1 = np.asarray(c1, dtype=np.uint64) & _MASK32
    c2 = np.asarray(c2, dtype=np.uint64) & _MASK32
    c3 = np.asarray(c3, dtype=np.uint64) & _MASK32
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(rounds):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0 = p0 >> np.uint64(32)
        lo0 = p0 & _MASK32
        hi1 = p1 >> np.uint64(32)
        lo1 = p1 & _MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n1 = lo1
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        n3 = lo0
        c0, c1, c2, c3 = n0, n1, n2, n3
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32),
            c2.astype(np.uint32), c3.astype(np.uint32))


def seed_key(seed: int):
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    return seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF


def stream_word(stream: int, iteration: int = 0) -> int:
    """Counter word 3: stream id in bits 0..7, training iteration above."""
    return (int(stream) & 0xFF) | ((int(iteration) & 0xFFFFFF) << 8)


def rand4(seed: int, stream: int, pos, sub=0, iteration: int = 0):
    """4 x uint32 for every ``pos`` (uint64 positions) and ``sub`` counter."""
    pos = np.asarray(pos, dtype=np.uint64)
    k0, k1 = seed_key(seed)
    return philox4x32(pos & _MASK32, pos >> np.uint64(32), sub,
                      stream_word(stream, iteration), k0, k1)


def mulhi32(a, b):
    """floor(a * b / 2**32) -- maps a uint32 uniformly onto [0, b)."""
    return ((np.asarray(a, dtype=np.uint64) * np.asarray(b, dtype=np.uint64))
            >> np.uint64(32)).astype(np.uint32)


def u32_to_unit_float(r):
    """uint32 -> float32 in [0, 1) using the top 24 bits (matches the device)."""
    return ((np.asarray(r, dtype=np.uint32) >> np.uint32(8)).astype(np.float32)
            * np.float32(1.0 / 16777216.0))
