"""Hand-assembles a small GVDB asset (<name>_mip<k>[c].vbx) byte by byte, independently of the product's C++ reader / writer.

Written from gvdb-voxel-src/GVDB_FILESPEC.txt and the read order of the reference's own loader (VolumeGVDB::LoadVBX,
gvdb-voxel-src/source/gvdb_library/src/gvdb_volume_gvdb.cpp:532-739: version 1.12 = the reference's custom minor version with
xform, inverse, effective voxel bounds and value range after the 1.11 transform block), the 64-byte nvdb::Node
(src/gvdb_node.h:40-53) and the pool-reference encoding grp | lev << 8 | ndx << 16 (src/gvdb_allocator.h:73-76).  The dense
levels come from the numpy restatement of the converter's mip rule (oracle/mip_oracle.py).  The reference pack (7.87 GB) is
not available offline and its writer needs the GVDB library (CUDA/OpenVDB) to run, so this is the second witness for the
.vbx reader: tests/test_vbx.py loads these files and compares every voxel with the dense arrays below.

  python tests/golden/make_vbx_fixture.py          (re)writes tests/golden/vbx_fixture/*
"""
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import mip_oracle as mo   # noqa: E402

UNDEF64 = 0xFFFFFFFFFFFFFFFF
DIM = (36, 28, 20)          # x, y, z voxels of mip 0 (odd halves exercise the 3-tap down-sampling)
VOXEL = 0.25
NUM_MIPS = 3


def field():
    """Two blobs with empty space between them (so that some bricks are missing), values in [0, 1.7]."""
    nx, ny, nz = DIM
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    a = np.exp(-(((x - 9.3) / 5.0) ** 2 + ((y - 10.1) / 4.0) ** 2 + ((z - 9.7) / 4.5) ** 2))
    b = 0.7 * np.exp(-(((x - 27.5) / 3.5) ** 2 + ((y - 18.2) / 3.0) ** 2 + ((z - 8.1) / 3.2) ** 2))
    d = (1.7 * np.maximum(a, b)).astype(np.float32)
    d[d < 0.12] = 0.0
    return d


def elem(grp, lev, ndx):
    return grp | (lev << 8) | (ndx << 16)


def write_vbx(path, dense, dim0):
    """One grid, float, 1 component, uncompressed, GVDB topology <3,4,5> (2 levels: every test grid fits one 128^3 node),
    atlas layout with a 1-voxel apron."""
    nz, ny, nx = dense.shape
    pad = np.zeros((nz + 2 + 8, ny + 2 + 8, nx + 2 + 8), dtype=np.float32)     # zeros outside the grid (and beyond the last partial brick)
    pad[1:nz + 1, 1:ny + 1, 1:nx + 1] = dense
    bx, by, bz = (nx + 7) // 8, (ny + 7) // 8, (nz + 7) // 8
    bricks = []
    for k in range(bz):
        for j in range(by):
            for i in range(bx):
                blk = pad[k * 8:k * 8 + 10, j * 8:j * 8 + 10, i * 8:i * 8 + 10]
                if np.any(blk != 0):
                    bricks.append((i, j, k, blk.copy()))
    assert bricks, "empty grid"
    n = len(bricks)
    ac = 1
    while ac ** 3 < n:
        ac += 1
    axisres = ac * 10
    atlas = np.zeros((axisres, axisres, axisres), dtype=np.float32)
    # bricks are placed in REVERSE order in the atlas: the reader must follow node.value, not the brick index
    slots = list(range(n))[::-1]
    nodes0 = []
    for b, (i, j, k, blk) in enumerate(bricks):
        s = slots[b]
        ax, ay, az = (s % ac) * 10, ((s // ac) % ac) * 10, (s // (ac * ac)) * 10
        atlas[az:az + 10, ay:ay + 10, ax:ax + 10] = blk
        nodes0.append(dict(lev=0, pos=(i * 8, j * 8, k * 8), value=(ax + 1, ay + 1, az + 1), vrange=(float(blk.min()), float(blk.max()), float(blk.mean())),
                           parent=elem(0, 1, 0), child=UNDEF64))
    child_row = [UNDEF64] * 4096
    for b, (i, j, k, _) in enumerate(bricks):
        child_row[(((k << 4) + j) << 4) + i] = elem(0, 0, b)
    node1 = dict(lev=1, pos=(0, 0, 0), value=(-1, -1, -1), vrange=(0.0, float(dense.max()), 0.0), parent=UNDEF64, child=elem(1, 1, 0))

    def node_bytes(nd):
        return struct.pack("<4B3i3i3f3Q", nd["lev"], 1, 0, 0, *nd["pos"], *nd["value"], *nd["vrange"], nd["parent"], nd["child"], 0)

    sx, sy, sz = VOXEL * dim0[0] / nx, VOXEL * dim0[1] / ny, VOXEL * dim0[2] / nz     # per-axis prescale of a mip (gvdb_volume_gvdb.cpp:2731)
    org = [-0.5 * dim0[a] * VOXEL for a in range(3)]
    xform = np.eye(4, dtype=np.float32); inv = np.eye(4, dtype=np.float32)
    for a, s in enumerate((sx, sy, sz)):
        xform[a, a] = s; xform[3, a] = org[a]
        inv[a, a] = 1.0 / s; inv[3, a] = -org[a] / s
    out = bytearray()
    out += struct.pack("<2B", 1, 12)
    out += struct.pack("<12f", 0, 0, 0, 0, 0, 0, 1, 1, 1, 0, 0, 0)          # pre-translation, Euler angles, scale, translation
    out += xform.astype("<f4").tobytes() + inv.astype("<f4").tobytes()
    out += struct.pack("<3i3i", 0, 0, 0, nx, ny, nz)                         # effective voxel bounds
    out += struct.pack("<2f", 0.0, float(dense.max()))                       # value range
    out += struct.pack("<i", 1)                                              # number of grids
    out += struct.pack("<Q", len(out) + 8)                                   # grid offset table
    name = b"density"
    out += name + b"\0" * (256 - len(name))
    out += struct.pack("<3B", ord("f"), 1, 0)
    out += struct.pack("<3f", 1, 1, 1)
    out += struct.pack("<i3iii", n, 8, 8, 8, 1, 1)                           # bricks, brick dims, apron, channels
    out += struct.pack("<Q", atlas.size * 4)
    out += struct.pack("<BiB", 2, 0, 0)                                      # topology GVDB, reuse 0, atlas layout
    out += struct.pack("<3i3i", ac, ac, ac, axisres, axisres, axisres)
    out += struct.pack("<iQ", 2, elem(0, 1, 0))                              # levels, root
    out += struct.pack("<ii3iiiii", 3, 8, 8, 8, 8, n, 64, 0, 0)              # level 0: log2 dim, res, range, node count, P0 width, P1 count, P1 width
    out += struct.pack("<ii3iiiii", 4, 16, 128, 128, 128, 1, 64, 1, 4096 * 8)
    for nd in nodes0:
        out += node_bytes(nd)
    out += node_bytes(node1)
    out += struct.pack("<4096Q", *child_row)
    out += struct.pack("<ii", 3, 4)                                          # channel type T_FLOAT, stride
    out += atlas.astype("<f4").tobytes()
    with open(path, "wb") as f:
        f.write(bytes(out))
    return n


def levels():
    """[(mip, conservative, dense (Z, Y, X) float32)] exactly as the converter would produce them."""
    d0 = field()
    out = []
    cur, cons = d0, mo.conservative0(d0)
    for m in range(NUM_MIPS):
        out.append((m, False, cur)); out.append((m, True, cons))
        cur, cons = mo.downsample(cur), mo.downsample(cons)
    return out


def main():
    d = os.path.join(HERE, "vbx_fixture")
    os.makedirs(d, exist_ok=True)
    for m, c, dense in levels():
        n = write_vbx(os.path.join(d, f"blobs_mip{m}{'c' if c else ''}.vbx"), dense, DIM)
        print(f"mip {m}{' conservative' if c else ''}: {dense.shape[::-1]} voxels, {n} bricks")


if __name__ == "__main__":
    main()
