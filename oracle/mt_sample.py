"""Oracle (test infrastructure only): the replay buffer's index stream.

The reference samples with ``random.sample(self.memory, batch_size)``
(/root/reference/sac/replay_buffer.py:39) from CPython's *global* Mersenne
Twister, seeded by ``random.seed(train.seed)`` (/root/reference/sac/agent.py:121).
``random.sample(deque, k)`` returns ``[deque[j] for j in J]`` where ``J`` is the
index stream ``random.sample(range(len(deque)), k)`` under the same generator
state.  The arithmetic therefore lives in a third-party
dependency that is not under /root/reference: CPython 3.12 ``Lib/random.py``
(``Random.sample``, ``Random._randbelow_with_getrandbits``) on top of
``Modules/_randommodule.c`` (MT19937, ``init_by_array`` seeding,
``getrandbits(k) = genrand_uint32() >> (32 - k)`` for k <= 32).

This file restates that published algorithm from scratch so the oracle does
not depend on the interpreter's own implementation; ``tests/test_oracle_sampling.py``
pins it against the stdlib ``random`` module (which *is* the reference's
dependency and is present on every box) for the reference's call pattern.
"""
from __future__ import annotations

import math
from typing import List

import numpy as np

_N, _M = 624, 397
_UPPER, _LOWER = 0x80000000, 0x7FFFFFFF
_MATRIX_A = 0x9908B0DF


class MT19937:
    """MT19937 with CPython's integer seeding (``random.seed(int)``)."""

    def __init__(self, seed: int = 0):
        self.mt = np.zeros(_N, dtype=np.uint64)
        self.idx = _N
        self.seed(seed)

    # -- seeding: _randommodule.c random_seed() -> init_by_array(key) --------
    def _init_genrand(self, s: int) -> None:
        mt = [0] * _N
        mt[0] = s & 0xFFFFFFFF
        for i in range(1, _N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt = np.array(mt, dtype=np.uint64)
        self.idx = _N

    def seed(self, a: int) -> None:
        a = abs(int(a))
        key: List[int] = []
        while True:  # little-endian 32-bit words; seed 0 -> [0]
            key.append(a & 0xFFFFFFFF)
            a >>= 32
            if a == 0:
                break
        self._init_genrand(19650218)
        mt = [int(x) for x in self.mt]
        i, j = 1, 0
        for _ in range(max(_N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & 0xFFFFFFFF
            i += 1
            j += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(_N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & 0xFFFFFFFF
            i += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
        mt[0] = 0x80000000
        self.mt = np.array(mt, dtype=np.uint64)
        self.idx = _N

    # -- generation ------------------------------------------------------------
    def _twist(self) -> None:
        mt = [int(x) for x in self.mt]
        for kk in range(_N):
            y = (mt[kk] & _UPPER) | (mt[(kk + 1) % _N] & _LOWER)
            mt[kk] = mt[(kk + _M) % _N] ^ (y >> 1) ^ (_MATRIX_A if (y & 1) else 0)
        self.mt = np.array(mt, dtype=np.uint64)
        self.idx = 0

    def genrand_uint32(self) -> int:
        if self.idx >= _N:
            self._twist()
        y = int(self.mt[self.idx])
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF

    def getrandbits(self, k: int) -> int:
        if not 0 < k <= 32:
            raise ValueError("oracle getrandbits supports 1..32 bits (buffers < 2**32 transitions)")
        return self.genrand_uint32() >> (32 - k)

    # -- Lib/random.py ---------------------------------------------------------
    def randbelow(self, n: int) -> int:
        """``Random._randbelow_with_getrandbits``: rejection on bit_length(n) bits."""
        k = n.bit_length()
        r = self.getrandbits(k)
        while r >= n:
            r = self.getrandbits(k)
        return r

    def sample_indices(self, n: int, k: int) -> List[int]:
        """``Random.sample(range(n), k)``: k distinct logical positions in [0, n).

        Follows both branches of CPython 3.12 ``Random.sample``: the pool-swap
        branch when ``n <= setsize`` and the set-rejection branch otherwise.
        """
        if not 0 <= k <= n:
            raise ValueError("Sample larger than population or is negative")
        result = [0] * k
        setsize = 21
        if k > 5:
            setsize += 4 ** math.ceil(math.log(k * 3, 4))
        if n <= setsize:
            pool = list(range(n))
            for i in range(k):
                j = self.randbelow(n - i)
                result[i] = pool[j]
                pool[j] = pool[n - i - 1]
        else:
            selected = set()
            for i in range(k):
                j = self.randbelow(n)
                while j in selected:
                    j = self.randbelow(n)
                selected.add(j)
                result[i] = j
        return result


def logical_to_slot(logical: np.ndarray, pushes: int, capacity: int) -> np.ndarray:
    """Map deque positions (0 = oldest survivor) to physical ring slots.

    ``deque(maxlen=capacity)`` (/root/reference/sac/replay_buffer.py:19,30) keeps
    the last ``capacity`` pushes; a ring that writes push number ``p`` (0-based)
    into slot ``p % capacity`` therefore holds deque position ``j`` at slot
    ``(max(pushes - capacity, 0) + j) % capacity``.
    """
    oldest = max(pushes - capacity, 0)
    return (oldest + np.asarray(logical, dtype=np.int64)) % capacity
