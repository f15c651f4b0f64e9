"""MT19937 and the CPU ``torch.randperm`` algorithm, restated in numpy.  TEST INFRASTRUCTURE.

Why: the reference draws one ``torch.randperm(count)`` per kept (image, class) pair from the
CPU default generator (``losses/DenseContrastiveLossV2.py:121``).  ATen is a third-party
dependency not vendored under /root/reference (torch 2.11.0 in this image); its CPU generator
is the published MT19937 (Matsumoto & Nishimura 1998) and ``randperm`` on CPU for n < 2**32/20
is a forward Fisher-Yates shuffle ``z = random() % (n - i); swap(r[i], r[i + z])`` for
i = 0..n-2.  Both are pinned by ``tests/test_oracle_golden.py`` against ``torch.randperm`` itself.
"""
import struct

import numpy as np

_N, _M = 624, 397
_UPPER, _LOWER, _MAG = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)


def _twist(u, v):
    y = (u & _UPPER) | (v & _LOWER)
    return (y >> np.uint32(1)) ^ np.where((v & np.uint32(1)) != 0, _MAG, np.uint32(0))


def next_block(mt):
    """One full 624-word state regeneration (three dependent vector phases)."""
    new = np.empty_like(mt)
    new[0:227] = mt[397:624] ^ _twist(mt[0:227], mt[1:228])
    new[227:454] = new[0:227] ^ _twist(mt[227:454], mt[228:455])
    new[454:623] = new[227:396] ^ _twist(mt[454:623], mt[455:624])
    new[623] = new[396] ^ _twist(mt[623], new[0])
    return new


def temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


class MT19937:
    """State = 624 words + ``pos`` (words of the current block already consumed, 624 = exhausted)."""

    def __init__(self, seed=None, state=None, pos=None):
        if state is not None:
            self.mt = np.asarray(state, dtype=np.uint32).copy()
            self.pos = int(pos)
        else:
            mt = np.empty(_N, dtype=np.uint64)
            mt[0] = seed & 0xFFFFFFFF
            for i in range(1, _N):
                mt[i] = (1812433253 * (int(mt[i - 1]) ^ (int(mt[i - 1]) >> 30)) + i) & 0xFFFFFFFF
            self.mt = mt.astype(np.uint32)
            self.pos = _N

    # -- torch CPU generator state <-> (mt, pos) -------------------------------------------
    @classmethod
    def from_torch_state(cls, state_bytes):
        """Parse ``torch.get_rng_state()``: [seed u64][left i32][seeded i32][next u64][624 x u64]..."""
        b = bytes(state_bytes)
        _seed, left, _seeded, nxt = struct.unpack_from("<QiiQ", b, 0)
        st = np.frombuffer(b, dtype=np.uint64, count=_N, offset=24).astype(np.uint32)
        pos = _N if left == 1 else int(nxt)   # left==1: the next draw regenerates the block
        return cls(state=st, pos=pos)

    def draw(self, k):
        """Next ``k`` tempered 32-bit outputs as uint32 array."""
        out = np.empty(k, dtype=np.uint32)
        done = 0
        while done < k:
            if self.pos >= _N:
                self.mt = next_block(self.mt)
                self.pos = 0
            take = min(k - done, _N - self.pos)
            out[done:done + take] = temper(self.mt[self.pos:self.pos + take])
            self.pos += take
            done += take
        return out


def randperm_full(n, gen):
    """Literal restatement: consumes n-1 draws, returns the whole permutation (int64)."""
    r = np.arange(n, dtype=np.int64)
    if n <= 1:
        return r
    u = gen.draw(n - 1).astype(np.int64)
    for i in range(n - 1):
        z = int(u[i] % (n - i))
        r[i], r[i + z] = r[i + z], r[i]
    return r


def randperm_prefix(n, v, gen):
    """First ``v`` entries of ``randperm_full`` (only the first v swaps matter) -- still
    consumes n-1 draws so the stream stays aligned with the reference."""
    if n <= 1:
        return np.arange(min(n, v), dtype=np.int64)
    u = gen.draw(n - 1)
    v = min(v, n)
    moved = {}
    out = np.empty(v, dtype=np.int64)
    for i in range(min(v, n - 1)):
        t = i + int(u[i]) % (n - i)
        a_i = moved.get(i, i)
        out[i] = moved.get(t, t)
        moved[t] = a_i
    if v == n:
        out[n - 1] = moved.get(n - 1, n - 1)
    return out
