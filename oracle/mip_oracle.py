"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's mip / conservative-mip rule and of the brick pool's
storage rules — an implementation independent of the C++ host builder (csrc/vr_scene.cpp) and of the CUDA builder
(csrc/vr_mipbuild.cu), which are both checked against it.  Only tests/ may import this.

  conservative0   gvdb-voxel-src/source/gvdb_library/src/gvdb_volume_gvdb.cpp:2753-2801  (zero voxels take the mean of the
                  positive values of their 27-neighbourhood, x offset outermost, z innermost, then / 27)
  downsample      gvdb_volume_gvdb.cpp:2803-2862  (2x box on even axes, 3-tap polyphase on odd axes; accumulation order
                  x tap outermost, z tap innermost; negative / zero results stored as 0)
  store           F/Scene/Scene.cpp:3139-3174  (1e-9 flush relative to the level's max, fp32 for mip 0 of the normal chain,
                  UNORM8 codes of v / max elsewhere, conservative codes never round a positive value to 0)

Arrays are (Z, Y, X) float32.  Pinned by hand-computed cases in tests/test_mip_oracle.py (the reference holds no fixtures
for its converter)."""
import numpy as np

F = np.float32


def _shifted(a, dx, dy, dz):
    """a[z+dz, y+dy, x+dx] with zeros outside the grid."""
    nz, ny, nx = a.shape
    out = np.zeros_like(a)
    zs, ze = max(0, -dz), min(nz, nz - dz)
    ys, ye = max(0, -dy), min(ny, ny - dy)
    xs, xe = max(0, -dx), min(nx, nx - dx)
    if zs < ze and ys < ye and xs < xe:
        out[zs:ze, ys:ye, xs:xe] = a[zs + dz:ze + dz, ys + dy:ye + dy, xs + dx:xe + dx]
    return out


def conservative0(src):
    src = np.asarray(src, dtype=F)
    avg = np.zeros_like(src)
    for ii in (-1, 0, 1):
        for jj in (-1, 0, 1):
            for kk in (-1, 0, 1):
                t = _shifted(src, ii, jj, kk)
                avg = (avg + np.where(t > 0, t, F(0))).astype(F)
    avg = (avg / F(27)).astype(F)
    return np.where((src == 0) & (avg > 0), avg, src).astype(F)


def _weights(n, cur):
    """(3, cur) tap weights of one axis: n = 2 (even source axis) or 3 (odd)."""
    i = np.arange(cur, dtype=np.int64)
    if n == 2:
        h = np.full(cur, 0.5, dtype=F)
        return np.stack([h, h, h])
    den = F(2 * cur + 1)
    return np.stack([((cur - i).astype(F) / den).astype(F), np.full(cur, F(cur) / den, dtype=F), ((1 + i).astype(F) / den).astype(F)])


def downsample(prev):
    prev = np.asarray(prev, dtype=F)
    pz, py, px = prev.shape
    nx, ny, nz = max(1, px // 2), max(1, py // 2), max(1, pz // 2)
    ni, nj, nk = (2 if px % 2 == 0 else 3), (2 if py % 2 == 0 else 3), (2 if pz % 2 == 0 else 3)
    wi, wj, wk = _weights(ni, nx), _weights(nj, ny), _weights(nk, nz)
    pad = np.zeros((2 * nz + 2, 2 * ny + 2, 2 * nx + 2), dtype=F)     # taps past the grid read 0
    pad[:pz, :py, :px] = prev
    res = np.zeros((nz, ny, nx), dtype=F)
    for ii in range(ni):
        for jj in range(nj):
            for kk in range(nk):
                w = ((wi[ii][None, None, :] * wj[jj][None, :, None]).astype(F) * wk[kk][:, None, None]).astype(F)
                v = pad[kk:kk + 2 * nz:2, jj:jj + 2 * ny:2, ii:ii + 2 * nx:2]
                res = (res + (w * v).astype(F)).astype(F)
    return np.where(res > 0, res, F(0)).astype(F)


def store(raw, fp32, conservative):
    """(stored values as float32, max_value): what a brick pool holds for this level, dequantised."""
    raw = np.asarray(raw, dtype=F)
    maxv = F(np.abs(raw).max()) if raw.size else F(0)
    if not maxv > 0:
        maxv = F(1)
    rel = (raw / maxv).astype(F)
    v = np.where((rel < F(1e-9)) & (raw >= 0), F(0), raw).astype(F)
    if fp32:
        return v, maxv
    rel = (v / maxv).astype(F)
    x = 255.0 * rel.astype(np.float64)
    q = np.clip(np.sign(x) * np.floor(np.abs(x) + 0.5), 0, 255).astype(np.int64)      # lround: half away from zero
    if conservative:
        q = np.where((q == 0) & (v > 0), 1, q)
    return (q.astype(F) * F(0.003921568859368563) * maxv).astype(F), maxv


def chain(dense, num_mips):
    """[(normal stored, conservative stored)] per level, like the volume's slots m and 8 + m."""
    cur = np.asarray(dense, dtype=F)
    cons = conservative0(cur)
    out = []
    for m in range(num_mips):
        out.append((store(cur, m == 0, False)[0], store(cons, False, True)[0]))
        if m + 1 < num_mips:
            if min(cur.shape) < 2:
                break
            cur, cons = downsample(cur), downsample(cons)
    return out
