"""CPU: the oracle's restated RNG recipes against the toolchain's libstdc++ itself.

The reference draws through std::mt19937 + std::uniform_int_distribution /
uniform_real_distribution / normal_distribution (envpool/core/env.h:75,113; env headers), so
libstdc++ <random> is the ground truth of SURVEY.md 8(a8).  Env trajectories already pin the
recipes on sampled streams; this test pins the branches sampling never reaches by loading
CRAFTED engine states (the 624 words + position that operator<< prints) into both
oracle/ep_oracle.c and a real std::mt19937 (oracle/ref_harness/std_rng.cc):

  * Lemire rejection in uniform_int (range 3/5: only the word 0; range 13: words k * 13^-1
    mod 2^32 for k < 9 -- 2e-9 per draw), single and back to back;
  * generate_canonical's `>= 1 -> nextafter(1, 0)` clamp (both words 0xFFFFFFFF) and exact 0;
  * normal_distribution's polar rejection, saved second value and reset.
The CUDA kernels restate the same recipes and are compared with the oracle on the GPU.
"""
import ctypes
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def untemper(y):
    """Inverse of mt19937's output tempering (so a table word can be chosen to make the
    engine emit a wanted value)."""
    y = int(y) & 0xFFFFFFFF
    y ^= y >> 18
    y ^= (y << 15) & 0xEFC60000
    t = y
    for _ in range(5):
        t = y ^ ((t << 7) & 0x9D2C5680)
    y = t & 0xFFFFFFFF
    t = y
    for _ in range(3):
        t = y ^ (t >> 11)
    return t & 0xFFFFFFFF


def temper(y):
    y ^= y >> 11
    y ^= (y << 7) & 0x9D2C5680
    y ^= (y << 15) & 0xEFC60000
    y ^= y >> 18
    return y & 0xFFFFFFFF


class Pair:
    """The same engine state in the oracle and in a real std::mt19937."""

    def __init__(self):
        from oracle import oracle_lib

        self.L = oracle_lib.lib()
        self.L.epo_debug_set_rng.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                             ctypes.c_int]
        self.L.epo_debug_uniform_int.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_int]
        self.L.epo_debug_uniform_real.restype = ctypes.c_double
        self.L.epo_debug_uniform_real.argtypes = [ctypes.c_void_p, ctypes.c_int,
                                                  ctypes.c_double, ctypes.c_double]
        self.L.epo_debug_normal.restype = ctypes.c_double
        self.L.epo_debug_normal.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_double,
                                            ctypes.c_double]
        self.pool = oracle_lib.OraclePool("CartPole", 1, seed=1, max_episode_steps=500)
        so = os.path.join(ROOT, "oracle", "libstd_rng.so")
        if not os.path.exists(so):
            pytest.skip("oracle/libstd_rng.so not built (make -C oracle oracle)")
        S = ctypes.CDLL(so)
        S.stdrng_create.restype = ctypes.c_void_p
        S.stdrng_create.argtypes = [ctypes.c_uint32]
        S.stdrng_set.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        S.stdrng_next.restype = ctypes.c_uint32
        S.stdrng_next.argtypes = [ctypes.c_void_p]
        S.stdrng_uniform_int.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int]
        S.stdrng_uniform_real.restype = ctypes.c_double
        S.stdrng_uniform_real.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        S.stdrng_normal.restype = ctypes.c_double
        S.stdrng_normal.argtypes = [ctypes.c_void_p, ctypes.c_double, ctypes.c_double]
        self.S = S
        self.h = S.stdrng_create(1)

    def load_outputs(self, outputs):
        """Make both engines emit `outputs` next (at most 623 words)."""
        assert len(outputs) <= 623
        mt = np.random.default_rng(len(outputs)).integers(0, 2**32, size=624, dtype=np.uint32)
        for i, v in enumerate(outputs):
            mt[1 + i] = untemper(v)
        assert self.S.stdrng_set(self.h, mt.ctypes.data, 1) == 0
        self.L.epo_debug_set_rng(self.pool.h, 0, mt.ctypes.data, 1)

    def both(self, fn, *args):
        a = getattr(self.L, "epo_debug_" + fn)(self.pool.h, 0, *args)
        b = getattr(self.S, "stdrng_" + fn)(self.h, *args)
        return a, b

    def same_position(self):
        return self.pool.draw(0) == self.S.stdrng_next(self.h)


@pytest.fixture(scope="module")
def pair():
    return Pair()


def test_untemper_inverts_temper():
    rng = np.random.default_rng(0)
    for v in rng.integers(0, 2**32, size=200):
        assert temper(untemper(int(v))) == int(v)


def test_seeded_streams_agree(pair):
    for seed in (0, 42, 5489):
        p = Pair()
        mt = np.zeros(624, dtype=np.uint32)
        s = seed
        mt[0] = s
        for i in range(1, 624):
            s = (1812433253 * (s ^ (s >> 30)) + i) & 0xFFFFFFFF
            mt[i] = s
        p.S.stdrng_set(p.h, mt.ctypes.data, 624)
        p.L.epo_debug_set_rng(p.pool.h, 0, mt.ctypes.data, 624)
        for k in range(2000):
            kind = k % 4
            if kind == 0:
                a, b = p.both("uniform_int", 1, 13)
            elif kind == 1:
                a, b = p.both("uniform_real", -0.05, 0.05)
            elif kind == 2:
                a, b = p.both("normal", 0.0, 0.1)
            else:
                a, b = p.both("uniform_int", -1, 1)
            assert a == b, (seed, k, kind, a, b)
        assert p.same_position()


@pytest.mark.parametrize("a,b", [(-1, 1), (0, 4), (0, 3), (0, 2), (1, 13), (0, 5), (0, 2**31 - 2)])
def test_lemire_rejections(pair, a, b):
    r = b - a + 1
    thr = (2**32 - r) % r
    inv = pow(r, -1, 2**32) if r % 2 else None
    rng = np.random.default_rng(r)
    rejected = []
    if inv is not None:
        rejected = [(k * inv) % 2**32 for k in range(thr)]     # w*r mod 2^32 = k < thr
    else:                                                      # even range: search a few
        w = 0
        while len(rejected) < min(thr, 4) and w < 10**6:
            if (w * r) % 2**32 < thr:
                rejected.append(w)
            w += 1
    for w in rejected:
        assert (w * r) % 2**32 < thr
    stream = []
    for k in range(150):
        stream.append(int(rng.integers(0, 2**32)))
        if rejected and k % 3 == 0:
            stream.extend(int(x) for x in rng.choice(rejected, size=1 + k % 3))
    stream = stream[:600]
    pair.load_outputs(stream)
    consumed_results = []
    for _ in range(100):
        x, y = pair.both("uniform_int", a, b)
        assert x == y and a <= x <= b
        consumed_results.append(x)
    assert pair.same_position()
    if thr:
        assert rejected, (a, b)


def test_canonical_clamp_and_zero(pair):
    pair.load_outputs([0xFFFFFFFF, 0xFFFFFFFF, 0, 0, 0xFFFFFFFF, 0, 0, 0xFFFFFFFF, 12345, 678])
    got = [pair.both("uniform_real", 0.0, 1.0) for _ in range(5)]
    for a, b in got:
        assert a == b
    assert got[0][0] == np.nextafter(1.0, 0.0) and got[1][0] == 0.0
    assert pair.same_position()
    # scaled: (canonical * (b - a)) + a, one rounding each
    pair.load_outputs([0xFFFFFFFF, 0xFFFFFFFF, 1, 0, 0x80000000, 0x7FFFFFFF])
    for lo, hi in ((-0.05, 0.05), (-np.pi, np.pi), (-0.6, -0.4)):
        a, b = pair.both("uniform_real", lo, hi)
        assert a == b


def test_normal_rejection_saved_value_and_reset(pair):
    big, small = 0xFFFFFFF0, 0x00000010
    # first candidate: x, y both near +1 -> r2 > 1 -> rejected; then an accepted pair
    stream = [big, big, big, big, small, 0x40000000, 0x12345678, 0x9ABCDEF0]
    stream += [int(v) for v in np.random.default_rng(2).integers(0, 2**32, size=200)]
    pair.load_outputs(stream)
    vals = [pair.both("normal", 0.0, 0.1) for _ in range(41)]    # odd count: one value saved
    for a, b in vals:
        assert a == b
    assert pair.same_position()
    # loading a state clears the saved second value on both sides (mj reset does not, the
    # distribution object lives as long as the env: covered by the env-level goldens)
    pair.load_outputs(stream[4:])
    a, b = pair.both("normal", 1.0, 2.0)
    assert a == b
