"""Host-side data contract: mip / conservative rule, tree + brick pool consistency, brick bounds, quantisation."""
import numpy as np

from volumetricrestirrelease_b200 import Scene, capi


def _scene(dense, mips=3):
    sc = Scene()
    sc.addGVDBVolume(dense=dense, numMips=mips, densityScale=1.0)
    return sc


def _rand_sparse(shape, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.random(shape, dtype=np.float32)
    d[d < 0.7] = 0
    d[: shape[0] // 3] = 0          # an empty slab -> inactive bricks
    return d


def test_mip0_round_trip_and_sparsity():
    d = _rand_sparse((40, 24, 33))
    sc = _scene(d)
    np.testing.assert_array_equal(sc.volume.dense_mip(0), d)
    bricks, nbytes = sc.volume.stats(0)
    assert nbytes == bricks * 1000 * 4
    assert bricks < (40 // 8) * (24 // 8) * ((33 + 7) // 8)      # the empty slab is not stored


def test_conservative_mip0_rule():
    """zero voxels become the mean of the positive 27-neighbourhood / 27 (gvdb_volume_gvdb.cpp:2753-2801); 8-bit, never 0."""
    d = _rand_sparse((16, 16, 16), seed=3)
    sc = _scene(d, mips=2)
    c = sc.volume.dense_mip(0, conservative=True)
    pad = np.pad(d, 1)
    acc = np.zeros_like(d)
    for dz in range(3):
        for dy in range(3):
            for dx in range(3):
                acc += pad[dz:dz + 16, dy:dy + 16, dx:dx + 16]
    expect = np.where(d == 0, acc / 27.0, d)
    mx = expect.max()
    q = np.clip(np.round(255.0 * expect / mx), 0, 255)
    q[(q == 0) & (expect > 0)] = 1
    np.testing.assert_allclose(c, q.astype(np.float32) * np.float32(0.003921568859368563) * mx, rtol=1e-6)
    assert ((c > 0) == (expect > 0)).all()          # conservative: support is never lost by quantisation


def test_box_filter_mips_even_and_odd():
    d = _rand_sparse((16, 18, 21), seed=5)             # z even, y even, x odd (21 -> 10 with the 3-tap polyphase filter)
    sc = _scene(d, mips=2)
    m1 = sc.volume.dense_mip(1)
    assert m1.shape == (8, 9, 10)
    nx = 10
    wx = np.zeros((10, 21), np.float32)
    for i in range(nx):
        den = np.float32(2 * nx + 1)
        wx[i, 2 * i:2 * i + 3] = [np.float32(nx - i) / den, np.float32(nx) / den, np.float32(1 + i) / den]
    t = np.einsum("ix,zyx->zyi", wx, d)
    t = 0.5 * (t[:, 0::2] + t[:, 1::2])
    t = 0.5 * (t[0::2] + t[1::2])
    mx = t.max()
    np.testing.assert_allclose(m1, np.round(255 * t / mx) / 255 * mx, atol=mx / 255 * 0.51)


def test_tree_levels_and_bounds():
    d = np.zeros((150, 20, 140), np.float32)
    d[5:9, 3:7, 130:135] = 2.0
    d[140:145, 10:12, 2:6] = 1.0
    sc = _scene(d, mips=1)
    g = sc.volume.grid.contents.slots[0]
    assert g.top_lev == 2 and g.valid == 1
    assert list(g.bmax) == [140.0, 20.0, 150.0]
    nodes = np.ctypeslib.as_array(C_ptr(g.nodes[0]), shape=(g.node_count[0], 8)) if False else None
    bricks = g.brick_count
    assert 2 <= bricks <= 16
    # every brick's bounds cover its stored block: max >= any interior voxel, avg = sum/512
    arr = np.frombuffer((capi.Node * g.node_count[0]).from_address(C_addr(g.nodes[0])), dtype=np.uint8).reshape(-1, 32)
    pos = arr[:, :12].view(np.int32)
    bounds = arr[:, 16:].view(np.float32)
    dm = sc.volume.dense_mip(0)
    for b in range(bricks):
        x, y, z = pos[b]
        blk = dm[z:z + 8, y:y + 8, x:x + 8]
        assert bounds[b, 1] >= blk.max() and bounds[b, 0] <= blk.min() + 1e-6
        assert x % 8 == 0 and y % 8 == 0 and z % 8 == 0


def C_addr(ptr):
    import ctypes
    return ctypes.addressof(ptr.contents)


def C_ptr(ptr):
    return ptr


def test_transforms_are_inverse_and_mips_cover_same_world_box():
    sc = Scene()
    sc.addGVDBVolume(dataFile="sphere", dim=(40, 32, 24), numMips=3, voxelSize=0.25, worldScaling=2.0, worldTranslation=(1, 2, 3))
    g = sc.volume.grid.contents
    for s in (0, 1, 2, 8, 9):
        sl = g.slots[s]
        m2w = np.array(list(sl.medium_to_world), np.float64).reshape(4, 4)
        w2m = np.array(list(sl.world_to_medium), np.float64).reshape(4, 4)
        np.testing.assert_allclose(m2w @ w2m, np.eye(4), atol=1e-5)
        lo = np.array([0, 0, 0, 1.0]) @ m2w
        hi = np.array([sl.bmax[0], sl.bmax[1], sl.bmax[2], 1.0]) @ m2w
        np.testing.assert_allclose(lo[:3], np.array([1, 2, 3]) - 2.0 * 0.5 * 0.25 * np.array([40, 32, 24]), atol=1e-4)
        np.testing.assert_allclose(hi[:3] - lo[:3], 2.0 * 0.25 * np.array([40, 32, 24]), atol=1e-4)
    v = g.volume
    assert abs(v.sigma_t - 10.0) < 1e-6 and abs(v.tStep - 0.25) < 1e-6 and abs(v.densityScaleFactorByScaling - 0.5) < 1e-6
