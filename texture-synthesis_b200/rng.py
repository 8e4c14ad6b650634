"""Host-side Pcg32 exactly as the reference's RNG stack produces it.

rand_pcg 0.3.1 `Lcg64Xsh32`, rand_core 0.6.3 `SeedableRng::seed_from_u64`, rand 0.8.5 `gen_range`
(call sites: reference lib/src/ms.rs:386,454-458,549-564,803-804).  Used for schedule planning
in Python helpers and for the deterministic synthetic example textures of the benchmark.
"""
M64 = (1 << 64) - 1
PCG_MUL = 6364136223846793005
SEED_INC = 11634580027462260723


def _rotr32(x, r):
    r &= 31
    return ((x >> r) | (x << (32 - r))) & 0xFFFFFFFF if r else x


class Pcg32:
    __slots__ = ("state", "inc")

    def __init__(self, state, inc):
        self.state, self.inc = state & M64, inc & M64

    @classmethod
    def from_state_incr(cls, state, inc):
        r = cls((state + inc) & M64, inc)
        r._step()
        return r

    @classmethod
    def new(cls, state, stream):
        return cls.from_state_incr(state, ((stream << 1) | 1) & M64)

    @classmethod
    def from_seed(cls, seed16):
        a = int.from_bytes(bytes(seed16[:8]), "little")
        b = int.from_bytes(bytes(seed16[8:16]), "little")
        return cls.from_state_incr(a, b | 1)

    @classmethod
    def seed_from_u64(cls, st):
        st &= M64
        seed = bytearray()
        for _ in range(4):
            st = (st * PCG_MUL + SEED_INC) & M64
            xs = (((st >> 18) ^ st) >> 27) & 0xFFFFFFFF
            seed += _rotr32(xs, st >> 59).to_bytes(4, "little")
        return cls.from_seed(seed)

    def _step(self):
        self.state = (self.state * PCG_MUL + self.inc) & M64

    def next_u32(self):
        s = self.state
        self._step()
        return _rotr32((((s >> 18) ^ s) >> 27) & 0xFFFFFFFF, s >> 59)

    def next_u64(self):
        lo = self.next_u32()
        return (self.next_u32() << 32) | lo

    def gen_range_u32(self, n):
        zone = ((n << (32 - n.bit_length())) & 0xFFFFFFFF) - 1
        while True:
            m = self.next_u32() * n
            if (m & 0xFFFFFFFF) <= zone:
                return m >> 32

    def gen_range_usize(self, n):
        zone = ((n << (64 - n.bit_length())) & M64) - 1
        while True:
            m = self.next_u64() * n
            if (m & M64) <= zone:
                return m >> 64
