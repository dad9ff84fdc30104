"""Independent numpy restatement of the reference's alias-table builder.  TEST INFRASTRUCTURE ONLY (tests/ import it).

Follows F/Utils/Sampling/AliasTable.cpp:46-126 (constructor) and F/Experimental/Scene/Lights/EmissivePowerSampler.cpp:64-79
(weights = triangle flux, RNG = std::default_random_engine(123), which is std::mt19937 on the reference's only toolchain,
MSVC; the constructor's signature takes std::mt19937&).  Triangle flux = luminance(Le) * area * pi
(F/Experimental/Scene/Lights/FinalizeIntegration.cs.slang:73).  Written from those files, not from the product's builder.

Pins: MT19937 is checked against the C++ standard's known answer (10000th output of a default-seeded engine = 4123659995)
in tests/test_alias_tables.py; the reference holds no fixture for the table itself, so beyond that the checks are the
sampling distribution the table encodes (exactly weight_i / sum) and item-for-item equality with the product's table.
std::sort is not stable: tables are only comparable for weights without ties (the tests use such weights).
"""
import numpy as np


class MT19937:
    """std::mt19937 (32-bit Mersenne twister), seeded like the C++ constructor (init_genrand)."""

    def __init__(self, seed=5489):
        mt = [0] * 624
        mt[0] = seed & 0xFFFFFFFF
        for i in range(1, 624):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & 0xFFFFFFFF
        self.mt, self.idx = mt, 624

    def _twist(self):
        mt = self.mt
        for i in range(624):
            y = (mt[i] & 0x80000000) | (mt[(i + 1) % 624] & 0x7FFFFFFF)
            v = mt[(i + 397) % 624] ^ (y >> 1)
            if y & 1:
                v ^= 0x9908B0DF
            mt[i] = v
        self.idx = 0

    def __call__(self):
        if self.idx >= 624:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & 0xFFFFFFFF


def build_alias_table(weights, seed=123):
    """AliasTable::AliasTable.  Returns (items uint32 [count, 4] = {threshold bits, indexA, indexB, 0}, weight sum as float32)."""
    w = np.asarray(weights, dtype=np.float32).copy()
    count = len(w)
    weight_sum = 0.0
    for f in w:                      # double accumulator, float addends
        weight_sum += float(f)
    factor = count / weight_sum
    w = (w.astype(np.float64) * factor).astype(np.float32)
    permutation = [int(i) for i in np.argsort(w, kind="stable")]
    thresholds = np.zeros(count, dtype=np.float32)
    redirect = [0] * count
    one = np.float32(1.0)
    head, tail = 0, count - 1
    while head != tail:
        i, j = permutation[head], permutation[tail]
        thresholds[i] = w[i]
        redirect[i] = j
        w[j] = np.float32(w[j] - np.float32(one - w[i]))
        if head == tail - 1:
            thresholds[j] = one
            redirect[j] = j
            break
        elif w[j] < one:
            permutation[head], permutation[tail] = permutation[tail], permutation[head]
            tail -= 1
        else:
            head += 1
    permutation = list(range(count))
    thresholds = [np.float32(t) for t in thresholds]
    rng = MT19937(seed)
    for i in range(count):
        dst = i + (rng() % (count - i))
        thresholds[i], thresholds[dst] = thresholds[dst], thresholds[i]
        redirect[i], redirect[dst] = redirect[dst], redirect[i]
        permutation[i], permutation[dst] = permutation[dst], permutation[i]
    items = np.zeros((count, 4), dtype=np.uint32)
    items[:, 0] = np.array(thresholds, dtype=np.float32).view(np.uint32)
    items[:, 1] = np.array(redirect, dtype=np.uint32)
    items[:, 2] = np.array(permutation, dtype=np.uint32)
    return items, np.float32(weight_sum)


def table_distribution(items):
    """Probability of every index under AliasTable::sample (F/Utils/Sampling/AliasTable.slang:55-69):
    slot uniform, `rnd >= threshold ? indexA : indexB`."""
    thr = items[:, 0].copy().view(np.float32).astype(np.float64)
    count = len(thr)
    p = np.zeros(count)
    np.add.at(p, items[:, 1], (1.0 - thr) / count)
    np.add.at(p, items[:, 2], thr / count)
    return p


def triangle_flux(tris):
    """tris: ctypes array of vrestir_emissive_triangle {posW[3][3], normal[3], area, Le[3]} -> float32 flux per triangle."""
    out = np.zeros(len(tris), dtype=np.float32)
    pi = np.float32(3.14159265358979323846)
    for i, t in enumerate(tris):
        lum = np.float32(np.float32(np.float32(0.2126) * np.float32(t.Le[0])) + np.float32(np.float32(0.7152) * np.float32(t.Le[1])))
        lum = np.float32(lum + np.float32(np.float32(0.0722) * np.float32(t.Le[2])))
        out[i] = np.float32(np.float32(lum * np.float32(t.area)) * pi)
    return out


def emissive_alias_table(tris):
    """What vrestir_get_emissive_alias returns, rebuilt from the reference's rules: (items, weights, weight sum)."""
    w = triangle_flux(tris)
    items, ws = build_alias_table(w)
    return items, w, float(ws)


def env_alias_distribution(thresholds, redirect):
    """Texel probabilities of the env-map alias table (product extension, no reference counterpart: slot uniform, keep own texel
    with probability `threshold`, else the redirect)."""
    thr = np.asarray(thresholds, dtype=np.float64)
    n = len(thr)
    p = thr / n
    np.add.at(p, np.asarray(redirect, dtype=np.int64), (1.0 - thr) / n)
    return p
