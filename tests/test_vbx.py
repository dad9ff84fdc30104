"""GVDB .vbx writer / reader: a scene saved as version-1.12 .vbx files and loaded back must give
byte-identical grids (tree nodes, child lists, brick pools after the load-time UNORM8 quantisation, transforms, VolumeDesc).
No reference test pins .vbx parsing and no .vbx asset is available offline (the 7.87 GB scene pack), so the pin is the
round trip plus the structural checks of the file against GVDB_FILESPEC / VolumeGVDB::LoadVBX."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from volumetricrestirrelease_b200 import Scene, capi


def _slot_bytes(g):
    out = {}
    for l in range(3):
        n = g.node_count[l]
        out[f"nodes{l}"] = bytes(C.string_at(g.nodes[l], n * 32)) if n and g.nodes[l] else b""
        c = g.childlist_count[l]
        out[f"child{l}"] = bytes(C.string_at(g.childlist[l], c * 4)) if c and g.childlist[l] else b""
    bpv = 1 if g.atlas_format == 1 else 4
    out["atlas"] = bytes(C.string_at(g.atlas, g.brick_count * g.atlas_channels * 1000 * bpv)) if g.brick_count else b""
    return out


SCALARS = ("valid", "top_lev", "max_value", "compress_scale", "atlas_format", "atlas_channels", "brick_count")
ARRAYS = ("dim", "res", "vdel", "noderange", "bmin", "bmax", "xform", "invxform", "world_to_medium", "medium_to_world")


@pytest.mark.parametrize("kind,dim,extra", [("bunny", (72, 64, 56), {}), ("bunny", (200, 150, 140), {}),
                                            ("plume", (48, 64, 48), dict(hasVelocity=True, hasEmission=True))])
def test_vbx_round_trip(tmp_path, kind, dim, extra):
    sc = Scene()
    kw = dict(sigma_a=(1, 2, 3), sigma_s=(9, 8, 7), g=0.3, numMips=3, densityScale=0.4, worldTranslation=(1.0, -2.0, 0.5), worldScaling=1.5)
    sc.addGVDBVolume(dataFile=kind, dim=dim, seed=5, voxelSize=0.25, **kw, **extra)
    prefix = os.path.join(str(tmp_path), "vol", "vol")
    os.makedirs(os.path.dirname(prefix))
    sc.volume.save_vbx(prefix)
    files = sorted(os.listdir(os.path.dirname(prefix)))
    assert "vol_mip0.vbx" in files and "vol_mip0c.vbx" in files and "vol_mip2c.vbx" in files
    if extra:
        assert {"vol_temperature.vbx", "vol_velocity_x.vbx", "vol_velocity_y.vbx", "vol_velocity_z.vbx"} <= set(files)
    # header of the reference's custom version: 1.12, then 12 floats of transform, 32 floats xform / inverse, bounds, range, 1 grid
    raw = open(prefix + "_mip0.vbx", "rb").read(2 + 48 + 128 + 24 + 8 + 4)
    assert raw[0] == 1 and raw[1] == 12
    assert struct.unpack_from("<i", raw, 2 + 48 + 128 + 24 + 8)[0] == 1
    vmax = struct.unpack_from("<3i", raw, 2 + 48 + 128 + 12)
    assert tuple(vmax) == tuple(dim)

    sc2 = Scene()
    sc2.loadGVDBVolume(prefix, **kw)
    a, b = sc.volume.grid.contents, sc2.volume.grid.contents
    for name, _ in capi.VolumeDesc._fields_:
        va, vb = getattr(a.volume, name), getattr(b.volume, name)
        va = list(va) if hasattr(va, "__len__") else va
        vb = list(vb) if hasattr(vb, "__len__") else vb
        assert va == vb, name
    for slot in range(capi.MAX_SLOTS):
        ga, gb = a.slots[slot], b.slots[slot]
        assert ga.valid == gb.valid, slot
        if not ga.valid:
            continue
        for f in SCALARS:
            assert getattr(ga, f) == getattr(gb, f), (slot, f)
        for f in ARRAYS:
            top = ga.top_lev + 1
            la, lb = list(getattr(ga, f)), list(getattr(gb, f))
            if f in ("dim", "res", "vdel", "noderange"):
                la, lb = la[:top], lb[:top]
            if f in ("world_to_medium", "medium_to_world"):
                # derived from the fp32 xform stored in the file (the builder multiplies in double before rounding): 1 ulp
                np.testing.assert_allclose(la, lb, rtol=3e-7, atol=1e-7, err_msg=str((slot, f)))
            else:
                assert la == lb, (slot, f)
        for l in range(3):
            assert ga.node_count[l] == gb.node_count[l] and ga.childlist_count[l] == gb.childlist_count[l], (slot, l)
        sa, sb = _slot_bytes(ga), _slot_bytes(gb)
        for k in sa:
            assert sa[k] == sb[k], (slot, k)


def test_vbx_errors(tmp_path):
    sc = Scene()
    with pytest.raises(capi.VRestirError):
        sc.loadGVDBVolume(os.path.join(str(tmp_path), "missing"))
    bad = os.path.join(str(tmp_path), "bad_mip0.vbx")
    open(bad, "wb").write(b"\x01\x0b" + b"\0" * 100)      # version 1.11: no xform block
    with pytest.raises(capi.VRestirError):
        sc.loadGVDBVolume(os.path.join(str(tmp_path), "bad"))


def test_vbx_malformed_files_are_rejected(tmp_path):
    """A hostile / corrupted .vbx must come back as an error code: no exception through the C ABI, no out-of-range tree id
    handed to the device.  Header layout: 2 B version, 48 B transform, 128 B xform + inverse, 24 B bounds, 8 B range, numGrids,
    grid offset table, then per grid: name[256], dtype/comps/compress (3 B), voxel size (12 B), leafcnt, leafdim[3], apron, numChan,
    atlasSz (8 B), topo/reuse/layout (1 + 4 + 1 B), axiscnt[3], axisres[3], levels, root (8 B), level table (9 ints per level)."""
    sc = Scene()
    sc.addGVDBVolume(dataFile="bunny", dim=(40, 40, 40), seed=5, numMips=1)
    prefix = os.path.join(str(tmp_path), "v")
    sc.volume.save_vbx(prefix)
    good = open(prefix + "_mip0.vbx", "rb").read()
    hdr = 2 + 48 + 128 + 24 + 8
    grid = struct.unpack_from("<Q", good, hdr + 4)[0]
    o_leafcnt = grid + 256 + 3 + 12
    o_axisres = o_leafcnt + 4 + 12 + 4 + 4 + 8 + 1 + 4 + 1 + 12
    o_levels = o_axisres + 12
    o_table = o_levels + 4 + 8
    levels = struct.unpack_from("<i", good, o_levels)[0]
    assert levels == 2 and struct.unpack_from("<i", good, o_leafcnt)[0] == struct.unpack_from("<i", good, o_table + 5 * 4)[0]

    def attempt(name, patch):
        raw = bytearray(good)
        patch(raw)
        d = os.path.join(str(tmp_path), name)
        os.makedirs(d)
        open(os.path.join(d, "v_mip0.vbx"), "wb").write(bytes(raw))
        with pytest.raises(capi.VRestirError):
            Scene().loadGVDBVolume(os.path.join(d, "v"))

    attempt("grids", lambda r: struct.pack_into("<i", r, hdr, 0x7FFFFFF0))                       # numGrids
    attempt("negcnt", lambda r: (struct.pack_into("<i", r, o_leafcnt, -5), struct.pack_into("<i", r, o_table + 5 * 4, -5)))
    attempt("hugecnt", lambda r: (struct.pack_into("<i", r, o_leafcnt, 0x7FFFFFFF), struct.pack_into("<i", r, o_table + 5 * 4, 0x7FFFFFFF)))
    attempt("hugelist", lambda r: struct.pack_into("<i", r, o_table + 9 * 4 + 7 * 4, 0x7FFFFFFF))   # cnt1 of level 1
    attempt("axisres", lambda r: struct.pack_into("<3i", r, o_axisres, 1 << 19, 1 << 19, 1 << 19))
    attempt("negaxis", lambda r: struct.pack_into("<3i", r, o_axisres, -10, 10, 10))
    attempt("truncated", lambda r: r.__delitem__(slice(len(r) // 2, len(r))))
    # a child id beyond the brick pool: first entry of the level-1 child list (after the node pools)
    cnt = [struct.unpack_from("<i", good, o_table + n * 36 + 5 * 4)[0] for n in range(levels)]
    o_child = o_table + levels * 36 + sum(cnt) * 64
    attempt("childid", lambda r: struct.pack_into("<Q", r, o_child, (cnt[0] + 7) << 16))
    # the untouched file still loads
    Scene().loadGVDBVolume(prefix)


def test_hand_assembled_vbx_fixture_loads_voxel_exact():
    """Second witness for the .vbx reader: the asset under tests/golden/vbx_fixture was assembled byte by byte by a Python
    script written from GVDB_FILESPEC.txt and the reference loader's read order (tests/golden/make_vbx_fixture.py), with bricks
    stored in reverse atlas order.  Every voxel of every level must come back: mip 0 exactly (fp32), the coarse / conservative
    levels within half a UNORM8 code of the level's maximum (the loader quantises them like F/Scene/Scene.cpp:3164-3174), and a
    conservative code is never 0 where the value is positive."""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("make_vbx_fixture", os.path.join(here, "golden", "make_vbx_fixture.py"))
    fx = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fx)
    prefix = os.path.join(here, "golden", "vbx_fixture", "blobs")
    # the committed files are what the script writes today (the fixture cannot drift from its generator)
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        for m, c, dense in fx.levels():
            name = f"blobs_mip{m}{'c' if c else ''}.vbx"
            fx.write_vbx(os.path.join(tmp, name), dense, fx.DIM)
            assert open(os.path.join(tmp, name), "rb").read() == open(os.path.join(here, "golden", "vbx_fixture", name), "rb").read(), name
    sc = Scene()
    vol = sc.loadGVDBVolume(prefix, numMips=fx.NUM_MIPS, densityScale=0.5)
    g = vol.grid.contents
    assert g.volume.numMips == fx.NUM_MIPS
    for m, c, dense in fx.levels():
        slot = m + (8 if c else 0)
        s = g.slots[slot]
        assert s.valid and s.top_lev == 1 and tuple(int(v) for v in s.bmax) == dense.shape[::-1]
        got = vol.dense_mip(m, c)
        assert got.shape == dense.shape
        if m == 0 and not c:
            assert np.array_equal(got.view(np.uint32), dense.view(np.uint32))
            assert s.atlas_format == 0
        else:
            mx = float(dense.max())
            assert s.atlas_format == 1 and abs(s.max_value - mx) <= 1e-6 * mx
            # rounding to the nearest code; a conservative level lifts a positive value below half a code to code 1
            assert np.abs(got - dense).max() <= (1.0 if c else 0.5) * mx / 255 * 1.0001
            if c:
                assert ((got > 0) == (dense / mx >= 1e-9)).all()      # conservative: positive stays positive
        # transforms of the level: voxel (0,0,0) corner and the far corner map to the same model-space box for every mip
        M = np.array(list(s.xform), dtype=np.float64).reshape(4, 4)
        lo = np.array([0, 0, 0, 1.0]) @ M
        hi = np.array([dense.shape[2], dense.shape[1], dense.shape[0], 1.0]) @ M
        np.testing.assert_allclose(lo[:3], [-0.5 * d * fx.VOXEL for d in fx.DIM], rtol=1e-6)
        np.testing.assert_allclose(hi[:3], [0.5 * d * fx.VOXEL for d in fx.DIM], rtol=1e-6)
    # brick bounds of mip 0 from the raw floats (F/Scene/Scene.cpp:2981-3012): max over all bricks = the grid maximum
    n0 = g.slots[0].node_count[0]
    mx = max(g.slots[0].nodes[0][i].bounds[1] for i in range(n0))
    assert mx == float(fx.field().max())
