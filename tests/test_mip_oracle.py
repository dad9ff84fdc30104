"""CPU: the mip / conservative-mip rule.  (1) the numpy oracle against hand-computed cases, (2) the C++ host builder
(csrc/vr_scene.cpp, the thing the CUDA builder is compared with on the GPU) against the numpy oracle, bit for bit."""
import numpy as np
import pytest

from oracle import mip_oracle as mo
from volumetricrestirrelease_b200 import Scene


def test_conservative0_known_answers():
    a = np.zeros((3, 3, 3), np.float32)
    a[1, 1, 1] = 2.7
    c = mo.conservative0(a)
    assert c[1, 1, 1] == np.float32(2.7)                        # non-zero voxels are kept
    assert np.all(c[a == 0] == np.float32(2.7) / np.float32(27))   # every zero voxel sees the one positive neighbour
    b = np.zeros((5, 5, 5), np.float32); b[0, 0, 0] = 1.0
    cb = mo.conservative0(b)
    assert cb[2, 2, 2] == 0 and cb[1, 1, 1] == np.float32(1) / np.float32(27) and cb[0, 0, 2] == 0


def test_downsample_known_answers():
    a = np.arange(4 * 4 * 4, dtype=np.float32).reshape(4, 4, 4)
    d = mo.downsample(a)
    assert d.shape == (2, 2, 2)
    assert np.allclose(d[0, 0, 0], a[:2, :2, :2].mean()) and np.allclose(d[1, 1, 1], a[2:, 2:, 2:].mean())   # even axes: 2x box
    o = np.ones((5, 5, 5), np.float32)
    do = mo.downsample(o)
    assert do.shape == (2, 2, 2) and np.allclose(do, 1.0, atol=1e-6)                                           # odd axes: weights sum to 1
    x = np.zeros((1, 1, 5), np.float32); x[0, 0, :] = [1, 0, 0, 0, 0]
    dx = mo.downsample(np.broadcast_to(x, (2, 2, 5)).copy())
    assert np.allclose(dx[0, 0], [2 / 5, 0.0])                                                                 # taps (cur-i, cur, 1+i) / (2 cur + 1), cur = 2


def test_store_rules():
    raw = np.array([[[0.0, 1e-12, 0.001, 0.5, 1.0]]], np.float32)
    v, mx = mo.store(raw, True, False)
    assert mx == 1.0 and v[0, 0, 1] == 0 and v[0, 0, 2] == np.float32(0.001)          # 1e-9 flush, fp32 otherwise untouched
    q, _ = mo.store(raw, False, False)
    assert q[0, 0, 2] == 0 and np.isclose(q[0, 0, 3], 128 / 255)                      # 0.001 * 255 rounds to 0; 127.5 rounds away from zero
    qc, _ = mo.store(raw, False, True)
    assert np.isclose(qc[0, 0, 2], 1 / 255) and qc[0, 0, 1] == 0                      # conservative: positive never becomes 0 (but the flush wins)


@pytest.mark.parametrize("dim", [(40, 36, 32), (41, 35, 29)])
def test_host_builder_matches_numpy_oracle(dim):
    nx, ny, nz = dim
    rng = np.random.default_rng(2)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    blob = np.exp(-(((x - nx / 2) / (nx / 4)) ** 2 + ((y - ny / 2) / (ny / 4)) ** 2 + ((z - nz / 2) / (nz / 4)) ** 2))
    dense = (np.clip(blob + 0.3 * rng.random((nz, ny, nx)) - 0.55, 0, None) * 2.0).astype(np.float32)
    dense[dense < 0.05] = 0.0
    dense[nz // 2, ny // 2, nx // 2] = 1e-12
    vol = Scene().addGVDBVolume(dense=dense, numMips=3)
    want = mo.chain(dense, 3)
    assert len(want) == 3
    for m, (normal, cons) in enumerate(want):
        for c, ref in ((False, normal), (True, cons)):
            got = vol.dense_mip(m, c)
            assert got.shape == ref.shape
            assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), f"mip {m} conservative {c}: {(got != ref).sum()} of {got.size} voxels differ"
