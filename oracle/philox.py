"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11), restated in
numpy.  TEST INFRASTRUCTURE for the opt-in counter-based sampler: the CUDA path draws the
permutation stream of a call from this generator instead of the torch CPU MT19937 when the module is configured with
``sampler='philox'``; parity then holds against the sampling oracle fed the same stream (``PhiloxStream`` has the
``draw(k)`` interface of ``oracle.mt19937.MT19937``).  Pinned by the known-answer vectors of the Random123
distribution (tests/test_oracle_golden.py).

Stream layout of one call: word j = philox4x32_10(counter = (j >> 2, 0, call & 0xffffffff, call >> 32),
key = (seed & 0xffffffff, seed >> 32))[j & 3].
"""
import numpy as np

_M0, _M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_W0, _W1 = np.uint32(0x9E3779B9), np.uint32(0xBB67AE85)
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(ctr, key):
    """ctr: 4 uint32 arrays (broadcastable), key: 2 uint32 scalars/arrays -> 4 uint32 arrays."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint32) for c in ctr)
    k0, k1 = (np.asarray(k, dtype=np.uint32) for k in key)
    with np.errstate(over="ignore"):
        for _ in range(10):
            p0 = c0.astype(np.uint64) * _M0
            p1 = c2.astype(np.uint64) * _M1
            hi0, lo0 = (p0 >> np.uint64(32)).astype(np.uint32), (p0 & _MASK).astype(np.uint32)
            hi1, lo1 = (p1 >> np.uint64(32)).astype(np.uint32), (p1 & _MASK).astype(np.uint32)
            c0, c1, c2, c3 = hi1 ^ c1 ^ k0, lo1, hi0 ^ c3 ^ k1, lo0
            k0, k1 = (k0 + _W0).astype(np.uint32), (k1 + _W1).astype(np.uint32)
    return c0, c1, c2, c3


def stream_words(seed, call, start, count):
    """Words [start, start + count) of the stream of call ``call`` under ``seed`` (uint32 array)."""
    j = np.arange(start, start + count, dtype=np.uint64)
    blk = j >> np.uint64(2)
    out = philox4x32_10(((blk & _MASK).astype(np.uint32), (blk >> np.uint64(32)).astype(np.uint32),
                         np.uint32(call & 0xFFFFFFFF), np.uint32((call >> 32) & 0xFFFFFFFF)),
                        (np.uint32(seed & 0xFFFFFFFF), np.uint32((seed >> 32) & 0xFFFFFFFF)))
    lane = (j & np.uint64(3)).astype(np.int64)
    return np.stack([np.broadcast_to(o, blk.shape) for o in out], 0)[lane, np.arange(count)]


class PhiloxStream:
    """Sequential view of one call's stream: ``draw(k)`` like ``oracle.mt19937.MT19937``."""

    def __init__(self, seed, call):
        self.seed, self.call, self.pos = int(seed), int(call), 0

    def draw(self, k):
        out = stream_words(self.seed, self.call, self.pos, k)
        self.pos += k
        return out
