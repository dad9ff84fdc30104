"""RNG pins: the only part of this path for which the reference holds golden data.

* Source/Tools/FalcorTest/Tests/Sampling/PseudorandomTests.cpp:94-155 checks xoshiro128** / SplitMix64 against the vendored
  public-domain C (Source/Externals/xoshiro) for 256 instances x 64 draws; here the oracle's generator is checked against
  oracle/_ref/libxoshiro_ref.so, which is that same C compiled from the reference checkout (oracle/Makefile).
* SURVEY.md section 8c KATs for the seeding formula (Morton16(pixel) | sampleNumber << 32 -> SplitMix64 x2 -> xoshiro128**).
"""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import vro


def words(px, py, n, count=4):
    w = (C.c_uint32 * count)()
    f = (C.c_float * count)()
    vro.lib().vro_rng_words(px, py, n, count, w, f)
    return list(w), list(f)


def test_seeding_kats():
    assert words(0, 0, 0)[0] == [0x1e93e34b, 0xa6a5a9ba, 0x24a3a744, 0x9ccb9af4]
    assert words(1, 2, 3)[0] == [0x0a080175, 0x9055b158, 0xec757185, 0x29ca8e5e]
    assert words(1919, 1079, 28)[0] == [0xf94d630f, 0x7b264d49, 0x9017eac2, 0xbbaeb807]
    np.testing.assert_allclose(words(0, 0, 0)[1], [0.11944407, 0.65096527, 0.14312214, 0.61248171], rtol=0, atol=5e-9)


def test_sample_next_1d_is_upper_24_bits():
    w, f = words(123, 456, 7, 64)
    for wi, fi in zip(w, f):
        assert fi == np.float32((wi >> 8) * 2.0 ** -24)
        assert 0.0 <= fi < 1.0


def test_morton():
    m = vro.lib().vro_morton
    assert m(0, 0) == 0 and m(1, 0) == 1 and m(0, 1) == 2 and m(3, 3) == 15
    assert m(0xFFFF, 0) == 0x55555555 and m(0, 0xFFFF) == 0xAAAAAAAA
    assert m(0x1FFFF, 5) == m(0xFFFF, 5)      # only the low 16 bits of each coordinate enter the seed


@pytest.mark.skipif(not os.path.exists(vro.REF_PRNG_PATH), reason="oracle/_ref not built (reference checkout absent)")
def test_against_reference_vendored_prng():
    ref = C.CDLL(vro.REF_PRNG_PATH)
    ref.ref_splitmix64_next.restype = C.c_uint64
    ref.ref_splitmix64_seed.argtypes = [C.c_uint64]
    ref.ref_xoshiro128ss_next.restype = C.c_uint32
    rng = np.random.default_rng(1234)
    for _ in range(256):
        px, py, n = (int(x) for x in rng.integers(0, 1 << 16, 3))
        seed = (n << 32) | vro.lib().vro_morton(px, py)
        ref.ref_splitmix64_seed(seed)
        s0, s1 = ref.ref_splitmix64_next(), ref.ref_splitmix64_next()
        st = (C.c_uint32 * 4)(s0 & 0xFFFFFFFF, s0 >> 32, s1 & 0xFFFFFFFF, s1 >> 32)
        ref.ref_xoshiro128ss_seed(st)
        expect = [ref.ref_xoshiro128ss_next() for _ in range(64)]
        assert words(px, py, n, 64)[0] == expect
