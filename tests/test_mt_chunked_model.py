"""CPU: the lazy, chunk-wise regeneration of the mt19937 table that the device RNG uses
(envpool_b200/csrc/common.cuh, `struct Mt`) produces std::mt19937's
sequence.  The scheme is restated here in numpy, 8-word chunks and all, and run for three
full table cycles against the oracle's block-twist generator (itself pinned to libstdc++
through the reference build and the C++ standard's 10000th-draw known answer).

Device layout: chunk c of an env = the 8 words 8c..8c+7.  A chunk is regenerated when the
read position enters it, from values it finds in the table at that moment:
  word i needs  old[i], old[i+1]  and  table[(i+397) % 624]
  * i+1 inside the chunk: its OLD value (the whole chunk is computed from loaded old words);
  * i+1 = first word of the next chunk: not regenerated yet (old) -- except after word 623,
    whose successor is word 0 of the NEW block, exactly what the block twist uses;
  * i+397 < 624: a later chunk, still old;  i+397 >= 624: an earlier chunk, already new --
    again what the block twist uses.
"""
import numpy as np

N, M = 624, 397
UPPER, LOWER, MAG = np.uint32(0x80000000), np.uint32(0x7FFFFFFF), np.uint32(0x9908B0DF)


def init_genrand(seed):
    mt = np.zeros(N, dtype=np.uint32)
    s = np.uint64(seed & 0xFFFFFFFF)
    mt[0] = s
    for i in range(1, N):
        s = (np.uint64(1812433253) * (s ^ (s >> np.uint64(30))) + np.uint64(i)) \
            & np.uint64(0xFFFFFFFF)
        mt[i] = s
    return mt


def temper(y):
    y = np.uint32(y)
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    return int(y)


class ChunkedMt:
    """table[78][8] + read position, regenerated one chunk at a time (device scheme)."""

    def __init__(self, seed):
        self.table = init_genrand(seed).reshape(N // 8, 8)
        self.idx = 0          # seed_kernel: "the first draw regenerates word 0"
        self.loads = 0

    def enter_chunk(self, c):
        flat = self.table.reshape(-1)
        base = 8 * c
        own = flat[base:base + 8].copy()                 # sector 1: own chunk (old)
        nxt = flat[(base + 8) % N]                       # sector 2: first word of the next
        far = np.array([flat[(base + k + M) % N] for k in range(8)], dtype=np.uint32)
        self.loads += 4                                  # sectors 3,4: the i+397 window
        new = np.empty(8, dtype=np.uint32)
        for k in range(8):
            succ = own[k + 1] if k < 7 else nxt
            y = (own[k] & UPPER) | (succ & LOWER)
            new[k] = far[k] ^ (y >> np.uint32(1)) ^ (MAG if (y & np.uint32(1)) else np.uint32(0))
        flat[base:base + 8] = new                        # one sector write

    def next(self):
        if self.idx % 8 == 0:
            self.enter_chunk((self.idx % N) // 8)
        v = self.table.reshape(-1)[self.idx % N]
        self.idx = (self.idx + 1) % N if self.idx + 1 != N else 0
        return temper(v)


def test_chunked_regeneration_equals_std_mt19937():
    from oracle.oracle_lib import OraclePool

    for seed in (0, 7, 5489, 2**31 - 1):
        orc = OraclePool("CartPole", 1, seed=seed, max_episode_steps=500)
        dev = ChunkedMt(seed)
        got = [dev.next() for _ in range(3 * N + 17)]
        want = [orc.draw(0) for _ in range(3 * N + 17)]
        assert got == want, (seed, next(i for i, (a, b) in enumerate(zip(got, want)) if a != b))
        assert dev.loads == 4 * ((3 * N + 17 + 7) // 8)   # 4 sector reads per 8 draws


def test_window_never_straddles_regenerated_and_old_words_wrongly():
    """The 8-word i+397 window of a chunk lies in at most two chunks, and for every word the
    table holds the value the block twist would use at that moment (old if i+397 < 624, new
    otherwise) -- checked structurally: within one chunk regeneration no needed word has
    been overwritten earlier in the same chunk."""
    for c in range(N // 8):
        need = {(8 * c + k + M) % N for k in range(8)} | {(8 * c + 8) % N}
        own = set(range(8 * c, 8 * c + 8))
        assert not (need & own), c
        assert len({w // 8 for w in need if w != (8 * c + 8) % N}) <= 2, c
