"""ORACLE / TEST INFRASTRUCTURE ONLY -- never imported by the product path.

numpy restatement of the random stream the reference consumes on CPU: torch's
default CPU generator is a 32-bit MT19937 (``at::mt19937``), seeded by
``torch.manual_seed(seed + 2)`` in ``explainer.py:14-22``.

Draw rules restated here (each is checked against torch itself in
``tests/test_oracle_rng.py``):

* ``torch.randint(0, 2, shape, dtype=bool)`` (``masks.py:130-132,258``;
  ``pathways.py:267-281``): one u32 per element, row-major, value = u32 % 2.
* ``torch.randperm(n)`` (``masks.py:385``; ``pathways.py:318``): Fisher-Yates,
  for i in 0..n-2: z = u32 % (n - i); swap(r[i], r[i + z]).
* ``nn.Linear(N, 1, bias=False)`` kaiming-uniform init (``wlm.py:45``):
  bound = 1/sqrt(N) (fp32); w = fp32(fp64((u32 & 0xFFFFFF) * 2^-24) * fp64(hi - lo) + fp64(lo))
* ``iter(DataLoader)`` (``wlm.py:210``): one int64 base seed = 2 u32 draws.

Parity: pinned against torch's own CPU generator (same library the reference
runs on), not against reference golden vectors -- the reference's tests hold no
RNG goldens (``tests/test_mask.py`` checks invariants only).
"""
import struct

import numpy as np

N, M = 624, 397
_UPPER, _LOWER, _MATRIX_A = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)


class MT19937:
    """32-bit Mersenne Twister with the state layout torch exposes."""

    def __init__(self, seed=5489):
        self.seed_value = int(seed)
        st = np.empty(N, dtype=np.uint64)
        st[0] = seed & 0xFFFFFFFF
        for j in range(1, N):
            prev = int(st[j - 1])
            st[j] = (1812433253 * (prev ^ (prev >> 30)) + j) & 0xFFFFFFFF
        self.state = st.astype(np.uint32)
        self.pos = N  # next output index; N means "twist first" (torch: left == 1)
        self.consumed = 0

    # -- torch interop ----------------------------------------------------
    @classmethod
    def from_torch_state(cls, blob):
        """blob: uint8 tensor/array from ``torch.get_rng_state()`` (5056 bytes)."""
        b = np.asarray(blob, dtype=np.uint8).tobytes()
        seed, left, seeded, nxt = struct.unpack("<QiiQ", b[:24])
        self = cls.__new__(cls)
        self.seed_value = seed
        self.state = np.frombuffer(b[24 : 24 + N * 8], dtype=np.uint64).astype(np.uint32)
        self.pos = N if left == 1 else int(nxt)
        self.consumed = 0
        return self

    def to_torch_state(self, template):
        """Write this state into a copy of ``template`` (a ``get_rng_state`` blob)."""
        b = bytearray(np.asarray(template, dtype=np.uint8).tobytes())
        left = 1 if self.pos == N else N + 1 - self.pos
        b[:24] = struct.pack("<QiiQ", self.seed_value, left, 1, self.pos)
        b[24 : 24 + N * 8] = self.state.astype(np.uint64).tobytes()
        return np.frombuffer(bytes(b), dtype=np.uint8).copy()

    # -- core -------------------------------------------------------------
    def _twist(self):
        s = self.state

        def mix(cur, nxt, far):
            y = (cur & _UPPER) | (nxt & _LOWER)
            return far ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), _MATRIX_A, np.uint32(0))

        new = np.empty_like(s)
        # new[i] needs old[i], old[i+1] and (old|new)[(i+M) % N]; do it in dependency order
        new[0 : N - M] = mix(s[0 : N - M], s[1 : N - M + 1], s[M:N])
        i = N - M
        while i < N - 1:
            j = min(i + (N - M), N - 1)
            new[i:j] = mix(s[i:j], s[i + 1 : j + 1], new[i - (N - M) : j - (N - M)])
            i = j
        new[N - 1] = mix(s[N - 1 : N], new[0:1], new[M - 1 : M])[0]
        self.state = new
        self.pos = 0

    @staticmethod
    def _temper(y):
        y = y ^ (y >> np.uint32(11))
        y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
        y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
        return y ^ (y >> np.uint32(18))

    def raw(self, n):
        """Next ``n`` tempered 32-bit outputs."""
        out = np.empty(n, dtype=np.uint32)
        done = 0
        while done < n:
            if self.pos == N:
                self._twist()
            take = min(n - done, N - self.pos)
            out[done : done + take] = self._temper(self.state[self.pos : self.pos + take])
            self.pos += take
            done += take
        self.consumed += n
        return out

    # -- torch draw rules -------------------------------------------------
    def randint_bool(self, rows, cols):
        return (self.raw(rows * cols) & np.uint32(1)).astype(bool).reshape(rows, cols)

    def randperm(self, n):
        r = np.arange(n, dtype=np.int64)
        if n < 2:
            return r
        u = self.raw(n - 1).astype(np.int64)
        for i in range(n - 1):
            z = int(u[i] % (n - i))
            r[i], r[i + z] = r[i + z], r[i]
        return r

    def linear_init(self, n):
        """Weights of ``nn.Linear(n, 1, bias=False)`` after its default init."""
        hi = np.float32(1.0) / np.sqrt(np.float32(n))  # gain*sqrt(3/fan_in) == 1/sqrt(n)
        hi = np.float32(_kaiming_bound(n))
        lo = np.float32(-hi)
        u = (self.raw(n) & np.uint32(0xFFFFFF)).astype(np.float64) * 2.0 ** -24
        return (u * np.float64(np.float32(hi - lo)) + np.float64(lo)).astype(np.float32)

    def dataloader_iter(self):
        """``iter(DataLoader)`` draws one int64 base seed (two u32)."""
        self.raw(2)


def _kaiming_bound(fan_in):
    """``kaiming_uniform_(a=sqrt(5))`` bound, evaluated like torch does (python doubles)."""
    import math

    gain = math.sqrt(2.0 / (1 + math.sqrt(5) ** 2))
    std = gain / math.sqrt(fan_in)
    return math.sqrt(3.0) * std
