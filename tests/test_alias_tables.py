"""The alias tables the product builds on the host (emissive triangles: the reference's AliasTable; env map: product
extension) against independent witnesses (oracle/alias_oracle.py, written from F/Utils/Sampling/AliasTable.cpp).  CPU only."""
import ctypes as C

import numpy as np
import pytest

from oracle import alias_oracle as ao
from volumetricrestirrelease_b200 import Scene, capi


def test_mt19937_known_answer():
    """ISO C++ [rand.predef]: the 10000th consecutive invocation of a default-constructed mt19937 produces 4123659995."""
    r = ao.MT19937()
    for _ in range(9999):
        r()
    assert r() == 4123659995


def _product_table(w):
    w = np.ascontiguousarray(w, dtype=np.float32)
    items = np.zeros((len(w), 4), dtype=np.uint32)
    ws = C.c_float()
    capi.check(capi.lib().vrestir_build_alias_table(w.ctypes.data, len(w), items.ctypes.data, C.byref(ws)))
    return items, ws.value


@pytest.mark.parametrize("n,seed", [(1, 0), (2, 1), (3, 2), (17, 3), (500, 4), (10000, 5)])
def test_emissive_alias_table_equals_reference_restatement(n, seed):
    rng = np.random.default_rng(seed)
    w = np.unique(rng.lognormal(mean=0.0, sigma=1.5, size=2 * n + 8).astype(np.float32))   # std::sort is not stable: tie-free weights
    w = rng.permutation(w)[:n].copy()
    assert len(w) == n and len(np.unique(w)) == n
    got, ws = _product_table(w)
    want, ws_want = ao.build_alias_table(w)
    assert np.array_equal(got, want)
    assert np.float32(ws) == ws_want
    # the distribution the table encodes is exactly weight_i / sum (up to the fp32 threshold arithmetic)
    p = ao.table_distribution(got)
    np.testing.assert_allclose(p, w.astype(np.float64) / w.astype(np.float64).sum(), rtol=1e-4, atol=1e-9)   # fp32 threshold arithmetic: a heavy item is decremented once per light one


def test_emissive_shell_flux_and_table():
    """Flux = luminance(Le) * area * pi (FinalizeIntegration.cs.slang:73) for the config-4 style triangle shell; the table the
    numpy restatement builds from it is item-for-item the product's."""
    sc = Scene()
    tris = sc.addEmissiveShell(2000, (0.0, 0.0, 0.0), 30.0, seed=4)
    items_w, w, ws = ao.emissive_alias_table(tris)
    got, ws_got = _product_table(w)
    assert np.array_equal(got, items_w) and np.float32(ws_got) == np.float32(ws)
    assert (w > 0).all()


def test_env_alias_table_distribution():
    """The env-map alias table (north-star extension) must encode texel_weight / sum exactly; also on a map with zero texels."""
    rng = np.random.default_rng(9)
    for n, zeros in ((64, 0), (4096, 0.3), (512 * 512, 0.05)):
        w = rng.random(n).astype(np.float32) ** 4
        w[rng.random(n) < zeros] = 0.0
        thr = np.zeros(n, dtype=np.float32)
        red = np.zeros(n, dtype=np.uint32)
        capi.check(capi.lib().vrestir_build_env_alias(w.ctypes.data, n, thr.ctypes.data, red.ctypes.data))
        assert (thr >= 0).all() and (thr <= 1).all() and (red < n).all()
        p = ao.env_alias_distribution(thr, red)
        want = w.astype(np.float64) / w.astype(np.float64).sum()
        np.testing.assert_allclose(p, want, rtol=0, atol=3e-7 / n * 4 + 1e-12)
        assert (p[w == 0] < 1e-12).all()      # a zero texel is never returned
