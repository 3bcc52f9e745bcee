"""CPython-compatible MT19937 (oracle; test infrastructure only).

Restates what the reference gets from the stdlib: ``random.Random(seed)``
(llm_rl_scripts/wordle/env/env.py:53), ``rng.choice`` (wordle/env/game.py:179),
the process-global ``random.seed`` / ``random.choice`` of the maze reset
(maze/env/randomness.py:9-14, maze/env/env.py:193-202) and ``getrandbits(64)`` of
``seed_generator`` (maze/env/randomness.py:35-41).  The algorithm is the
published MT19937 of Matsumoto & Nishimura with CPython's ``init_by_array``
seeding of 32-bit little-endian chunks of ``abs(seed)`` and CPython's
``_randbelow_with_getrandbits`` rejection loop.  Checked against the stdlib
itself in tests/test_oracle_rng.py.
"""
from __future__ import annotations

N, M = 624, 397
MASK32 = 0xFFFFFFFF


def seed_key(seed: int):
    """abs(seed) as little-endian 32-bit words (at least one word)."""
    s = abs(int(seed))
    key = []
    while True:
        key.append(s & MASK32)
        s >>= 32
        if s == 0:
            break
    return key


class MT19937:
    def __init__(self, seed: int):
        mt = [0] * N
        mt[0] = 19650218
        for i in range(1, N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & MASK32
        key = seed_key(seed)
        i, j = 1, 0
        for _ in range(max(N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & MASK32
            i += 1
            j += 1
            if i >= N:
                mt[0] = mt[N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & MASK32
            i += 1
            if i >= N:
                mt[0] = mt[N - 1]
                i = 1
        mt[0] = 0x80000000
        self.mt = mt
        self.idx = N
        self.draws = 0  # number of 32-bit outputs consumed (diagnostic)

    def _twist(self):
        mt = self.mt
        for k in range(N):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % N] & 0x7FFFFFFF)
            mt[k] = mt[(k + M) % N] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
        self.idx = 0

    def u32(self) -> int:
        if self.idx >= N:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        self.draws += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & MASK32

    def getrandbits(self, k: int) -> int:
        assert k > 0
        if k <= 32:
            return self.u32() >> (32 - k)
        out, shift = 0, 0
        while k > 0:
            w = self.u32()
            if k < 32:
                w >>= 32 - k
            out |= w << shift
            shift += 32
            k -= 32
        return out

    def randbelow(self, n: int) -> int:
        assert n > 0
        k = n.bit_length()
        r = self.getrandbits(k)
        while r >= n:
            r = self.getrandbits(k)
        return r

    def choice_index(self, n: int) -> int:
        return self.randbelow(n)

    def random(self) -> float:
        a = self.u32() >> 5
        b = self.u32() >> 6
        return (a * 67108864.0 + b) / 9007199254740992.0


def seed_stream(seed: int):
    """maze/env/randomness.py:35-41 — successive getrandbits(64) of one MT state."""
    g = MT19937(seed)
    while True:
        yield g.getrandbits(64)
