"""Second, independent witness for the two deterministic transmittance estimators.  TEST INFRASTRUCTURE ONLY.

Pure Python / numpy-float32 restatement, written directly from the reference's Slang and NOT from the product's CUDA or the
C++ oracle, of:
  * the iterative hierarchical DDA          VR/VolumeUtils.slang:171-282  + F/Scene/GVDB/gvdbDda.slang:86-157
  * ray-marched transmittance               VR/VolumeTrackingAdapterGVDB.slang:140-208, VR/VolumeUtils.slang:350-362
  * analytic (regular-tracking) transmittance, trilinear and point      VR/VolumeTrackingAdapterGVDB.slang:20-136
  * free-flight distance sampling (K1's candidate generation), point and trilinear, with the pixel's own xoshiro128** stream
                                            VR/VolumeTrackingAdapterGVDB.slang:210-437, F/Utils/Sampling/UniformSampleGenerator.slang
  * WorldToMedium / IntersectVolumeBound / DensityInAtlas / FetchEightVoxelsInAtlas     VR/VolumeBase.slang:103-175,234-263
  * getNode / getChild                      F/Scene/GVDB/gvdbNodes.slang:98-114
for one ray at a time over a grid slot in the layout of include/vrestir.h (32-byte nodes, dense child lists, brick pool of
10^3 blocks standing in for the 3-D atlas texture).  The texture unit is not in the reference source; it is restated as
DESIGN.md section 2 pins it: exact fp32 trilinear filtering as fma-lerps in x, then y, then z over the 8 texels around
(coordinate - 0.5), border colour 0, UNORM8 texels decoded as code * fl(1/255) after filtering.

Every arithmetic step is done in np.float32 in the order the Slang expression tree gives, so the accumulated optical depth
agrees with the C++ oracle to the last bit except where a fused multiply-add had to be emulated in double (one rounding in
double, one to float).  tests/test_march_witness.py compares the two on random rays.
"""
import ctypes as C

import math

import numpy as np

F = np.float32
ID_UNDEFL = 0xFFFFFFFF
MAX_BRICK_STEPS = 128
K_UNORM8 = F(0.003921568859368563)     # fl(1/255)

NODE = np.dtype([("pos", "<i4", 3), ("link", "<u4"), ("bounds", "<f4", 4)])


def slot_arrays(g):
    """numpy views of one vrestir_grid_slot (ctypes)."""
    s = {"top_lev": int(g.top_lev), "dim": [int(v) for v in g.dim], "res": [int(v) for v in g.res],
         "vdel": [F(v) for v in g.vdel], "bmin": np.array(list(g.bmin), dtype=F), "bmax": np.array(list(g.bmax), dtype=F),
         "w2m": np.array(list(g.world_to_medium), dtype=F).reshape(4, 4), "format": int(g.atlas_format),
         "channels": int(g.atlas_channels), "compress_scale": F(g.compress_scale), "max_value": F(g.max_value)}
    s["nodes"], s["child"] = [], []
    for l in range(3):
        n = int(g.node_count[l])
        s["nodes"].append(np.frombuffer(C.string_at(g.nodes[l], n * 32), dtype=NODE) if n and g.nodes[l] else np.zeros(0, NODE))
        c = int(g.childlist_count[l])
        s["child"].append(np.frombuffer(C.string_at(g.childlist[l], c * 4), dtype="<u4") if c and g.childlist[l] else np.zeros(0, "<u4"))
    nvox = int(g.brick_count) * s["channels"] * 1000
    if s["format"] == 1:
        s["atlas"] = np.frombuffer(C.string_at(g.atlas, nvox), dtype=np.uint8)
    else:
        s["atlas"] = np.frombuffer(C.string_at(g.atlas, nvox * 4), dtype="<f4")
    return s


def _fma(a, b, c):
    return F(np.float64(a) * np.float64(b) + np.float64(c))


def _lerp(a, b, t):
    return _fma(t, F(b - a), a)


def _v3(x, y, z):
    return np.array([x, y, z], dtype=F)


def _mul_point(p, M):   # row vector times matrix, w = 1 (the grids are affine: o.w == 1)
    return _v3(*[F(F(F(p[0] * M[0, j]) + F(p[1] * M[1, j])) + F(p[2] * M[2, j])) + M[3, j] for j in range(3)])


def _mul_vec(v, M):
    return _v3(*[F(F(v[0] * M[0, j]) + F(v[1] * M[1, j])) + F(v[2] * M[2, j]) for j in range(3)])


class _Texture:
    """The atlas sampler over one brick: p is the coordinate relative to the brick's interior min corner (voxel centres at +0.5)."""

    def __init__(self, slot):
        self.s = slot

    def texel(self, brick, ix, iy, iz, ch=0):
        if not (-1 <= ix <= 8 and -1 <= iy <= 8 and -1 <= iz <= 8):
            return F(0)   # only reachable through the checked fetches; inside the apron block by construction otherwise
        return F(self.s["atlas"][(brick * self.s["channels"] + ch) * 1000 + ((iz + 1) * 10 + (iy + 1)) * 10 + (ix + 1)])

    def decode(self, raw):
        return F(raw * K_UNORM8) if self.s["format"] == 1 else raw

    def point(self, brick, p):
        return self.decode(self.texel(brick, int(np.floor(p[0])), int(np.floor(p[1])), int(np.floor(p[2]))))

    def linear(self, brick, p, ch=0):
        q = p - F(0.5)
        f0 = np.floor(q)
        i = [int(f0[0]), int(f0[1]), int(f0[2])]
        fr = q - f0
        v = [[[self.texel(brick, i[0] + dx, i[1] + dy, i[2] + dz, ch) for dx in (0, 1)] for dy in (0, 1)] for dz in (0, 1)]
        c = [[_lerp(v[dz][dy][0], v[dz][dy][1], fr[0]) for dy in (0, 1)] for dz in (0, 1)]
        d = [_lerp(c[dz][0], c[dz][1], fr[1]) for dz in (0, 1)]
        return self.decode(_lerp(d[0], d[1], fr[2]))


class _HDDA:   # F/Scene/GVDB/gvdbDda.slang:86-157
    def set_from_ray(self, pos, dirv, tx, ty):
        self.pos, self.dir = pos, dirv
        self.pstep = np.array([1 if d > 0 else (-1 if d < 0 else 0) for d in dirv], dtype=np.int32)   # isign3
        self.tx, self.ty = F(tx), F(ty)

    def prepare(self, vmin, vdel):
        with np.errstate(divide="ignore", invalid="ignore"):
            self.tDel = np.abs(F(vdel) / self.dir)
            pflt = (self.pos + self.tx * self.dir - vmin) / F(vdel)
            fl = np.floor(pflt)
            self.tSide = ((fl - pflt + F(0.5)) * self.pstep.astype(F) + F(0.5)) * self.tDel + self.tx
        self.p = fl.astype(np.int32)

    def prepare_leaf(self, vmin):
        with np.errstate(divide="ignore", invalid="ignore"):
            self.tDel = np.abs(F(1.0) / self.dir)
            pflt = self.pos + self.tx * self.dir - vmin
            fl = np.floor(pflt)
            self.tSide = ((fl - pflt + F(0.5)) * self.pstep.astype(F) + F(0.5)) * self.tDel + self.tx
        self.p = fl.astype(np.int32)

    def next(self):
        t = self.tSide
        self.mask = np.array([int((t[0] < t[1]) & (t[0] <= t[2])), int((t[1] < t[2]) & (t[1] <= t[0])), int((t[2] < t[0]) & (t[2] <= t[1]))], dtype=np.int32)
        self.ty = t[0] if self.mask[0] else (t[1] if self.mask[1] else t[2])

    def step(self):
        self.tx = self.ty
        # `tSide += float3(mask) * tDel`; an axis with a zero direction has tDel = inf and mask 0 (0 * inf would be NaN:
        # the repo documents its select-form deviation for that case; the witness rays never have a zero component)
        self.tSide = (self.tSide + self.mask.astype(F) * self.tDel).astype(F)
        self.p = self.p + self.mask * self.pstep

    def copy(self):
        o = _HDDA()
        o.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in self.__dict__.items()})
        return o


class RayMarchingAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:140-208
    def __init__(self, linear, tstep):
        self.Tr, self.linear, self.tStep, self.initialized = F(0), linear, F(tstep), False

    def start(self):
        self.initialized = True

    def main(self, W, dda, vmin_leaf, brick, t):
        tS = self.tStep
        t = self.tNear + (np.floor((t - self.tNear) / tS) + F(0.5)) * tS
        if t < dda.tx:
            t = t + tS
        mult = F(1.0)
        wp = self.ray_o + t * self.ray_d
        p = wp - vmin_leaf
        wpt = mult * tS * self.ray_d
        res = F(W.slot["res"][0])
        it = 0
        while it < MAX_BRICK_STEPS and bool(np.all((p >= 0) & (p < res))):
            if t >= self.tFar:
                self.Tr = np.exp(self.Tr)
                return True, t
            density = W.density_in_atlas(brick, p, self.linear)
            sigma_t = density * W.sigma_t
            self.Tr = self.Tr + F(F(-sigma_t) * (F(1.0) if it == 0 else mult)) * tS
            p = p + wpt
            t = t + mult * tS
            it += 1
        return False, t

    def end(self):
        self.Tr = np.exp(self.Tr) if self.initialized else F(1.0)


class FeatureMarchAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:518-602 (ReservoirFeatureRayMarchingAdapterGVDB)
    """K0's total camera-ray transmittance: the same stepping as RayMarchingAdapter without the far-end test, stopping once the
    transmittance so far is below 1 %."""

    def __init__(self, linear, tstep):
        self.Tr, self.linear, self.tStep, self.initialized, self.accu = F(0), linear, F(tstep), False, F(1)

    def start(self):
        self.initialized = True

    def main(self, W, dda, vmin_leaf, brick, t):
        tS = self.tStep
        t = self.tNear + (np.floor((t - self.tNear) / tS) + F(0.5)) * tS
        if t < dda.tx:
            t = t + tS
        p = (self.ray_o + t * self.ray_d) - vmin_leaf
        wpt = tS * self.ray_d
        res = F(W.slot["res"][0])
        it = 0
        while it < MAX_BRICK_STEPS and bool(np.all((p >= 0) & (p < res))):
            if np.exp(self.Tr) < F(0.01):
                self.end()
                return True, t
            sigma_t = W.density_in_atlas(brick, p, self.linear) * W.sigma_t
            self.Tr = self.Tr + F(-sigma_t) * tS
            p = p + wpt
            t = t + tS
            it += 1
        return False, t

    def end(self):
        if self.initialized:
            self.accu = np.exp(self.Tr)


class AnalyticAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:20-136
    def __init__(self, linear):
        self.Tr, self.linear = F(0), linear

    def start(self):
        pass

    def main(self, W, dda, vmin_leaf, brick, t):
        leaf = dda.copy()
        leaf.prepare_leaf(vmin_leaf)
        res0 = W.slot["res"][0]
        it = 0
        while it < MAX_BRICK_STEPS and bool(np.all((leaf.p >= 0) & (leaf.p < res0))):
            leaf.next()
            maxDeltaT = leaf.ty - t
            if self.linear:
                v = W.fetch_eight(brick, leaf.p)
                v = [x * W.sigma_t for x in v]
                v000, v100, v010, v110, v001, v101, v011, v111 = v
                mxyz = v111 - v011 - v101 - v110 + v100 + v010 + v001 - v000
                mxy = v000 - v100 - v010 + v110
                mxz = v000 - v100 - v001 + v101
                myz = v000 - v010 - v001 + v011
                mx, my, mz = v100 - v000, v010 - v000, v001 - v000
                d = self.ray_d
                p0 = leaf.pos + leaf.tx * leaf.dir - (leaf.p.astype(F) + vmin_leaf)
                c3 = mxyz * d[0] * d[1] * d[2]
                c2 = (p0[2] * d[0] * d[1] + p0[1] * d[0] * d[2] + p0[0] * d[1] * d[2]) * mxyz + mxy * d[0] * d[1] + mxz * d[0] * d[2] + myz * d[1] * d[2]
                c1 = ((p0[1] * p0[2] * d[0] + p0[0] * p0[2] * d[1] + p0[0] * p0[1] * d[2]) * mxyz + mx * d[0] + my * d[1] + mz * d[2]
                      + (p0[1] * d[0] + p0[0] * d[1]) * mxy + (p0[2] * d[0] + p0[0] * d[2]) * mxz + (p0[2] * d[1] + p0[1] * d[2]) * myz)
                c0 = (p0[0] * p0[1] * p0[2] * mxyz + p0[0] * p0[1] * mxy + p0[0] * p0[2] * mxz + p0[1] * p0[2] * myz + p0[0] * mx + p0[1] * my
                      + p0[2] * mz + v000)
                t_dist = min(self.tFar - t, maxDeltaT)
                t2 = t_dist * t_dist
                t3 = t2 * t_dist
                t4 = t2 * t2
                self.Tr = self.Tr + -(c3 * t4 / F(4) + c2 * t3 / F(3) + c1 * t2 / F(2) + c0 * t_dist)
            else:
                density = W.density_in_atlas(brick, leaf.p.astype(F) + F(0.5), False)
                sigma_t = density * W.sigma_t
                self.Tr = self.Tr + F(-min(self.tFar - t, maxDeltaT)) * sigma_t
            if t + maxDeltaT >= self.tFar:
                self.Tr = np.exp(self.Tr)
                return True, t
            t = t + maxDeltaT
            leaf.step()
            it += 1
        return False, t

    def end(self):
        self.Tr = np.exp(self.Tr)


class Xoshiro:
    """UniformSampleGenerator (F/Utils/Sampling/UniformSampleGenerator.slang:49-72): SplitMix64 seeded with
    (interleave_32bit(pixel), sampleNumber) fills the state of xoshiro128** (Pseudorandom/Xoshiro.slang:47-66,
    SplitMix64.slang:52-58, F/Utils/Math/BitTricks.slang:45-61); sampleNext1D = (next() >> 8) * 2^-24."""
    M32, M64 = 0xFFFFFFFF, 0xFFFFFFFFFFFFFFFF

    def __init__(self, px, py, sample_number):
        def spread(v):
            v &= 0xFFFF
            v = (v | (v << 8)) & 0x00FF00FF
            v = (v | (v << 4)) & 0x0F0F0F0F
            v = (v | (v << 2)) & 0x33333333
            return (v | (v << 1)) & 0x55555555
        self._sm = ((sample_number & self.M32) << 32) | (spread(px) | (spread(py) << 1))
        s0, s1 = self._splitmix(), self._splitmix()
        self.s = [s0 & self.M32, s0 >> 32, s1 & self.M32, s1 >> 32]

    def _splitmix(self):
        self._sm = (self._sm + 0x9E3779B97F4A7C15) & self.M64
        z = self._sm
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & self.M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & self.M64
        return z ^ (z >> 31)

    @staticmethod
    def _rotl(x, k):
        return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF

    def next(self):
        s = self.s
        result = (self._rotl((s[0] * 5) & self.M32, 7) * 9) & self.M32
        t = (s[1] << 9) & self.M32
        s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]
        s[2] ^= t
        s[3] = self._rotl(s[3], 11)
        return result

    def next1d(self):
        return F(self.next() >> 8) * F(2.0 ** -24)


class DistanceSamplingAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:210-437 (SampleMediumAnalyticAdapterGVDB)
    """Free-flight distance sampling of up to 4 samples along one ray: per voxel cell an exponential step per pending sample (point
    sampler) or regula falsi on the cubic optical depth of the trilinear interpolant against pre-drawn targets (linear sampler)."""
    K_RAY_TMAX = F(3.402823466e+38)

    def __init__(self, num_samples, linear, rng):
        self.n, self.linear, self.rng = num_samples, linear, rng
        self.hit = [F(0)] * 4; self.outTr = [F(0)] * 4; self.pdf = [F(0)] * 4
        self.opt = F(0); self.initialized = False

    def start(self):
        for i in range(self.n):
            self.hit[i] = F(-1)
        self.initialized = True

    @staticmethod
    def _tau(t, c3, c2, c1, c0):
        t2 = t * t; t3 = t2 * t; t4 = t2 * t2
        return c3 * t4 / F(4) + c2 * t3 / F(3) + c1 * t2 / F(2) + c0 * t

    def main(self, W, dda, vmin_leaf, brick, t):
        leaf = dda.copy()
        leaf.prepare_leaf(vmin_leaf)
        res0 = W.slot["res"][0]
        it = 0
        while it < MAX_BRICK_STEPS and bool(np.all((leaf.p >= 0) & (leaf.p < res0))):
            leaf.next()
            maxDeltaT = min(self.tFar - t, leaf.ty - t)
            currentTMax = min(self.tFar, leaf.ty)
            finished = 0
            if self.linear:
                v = [x * W.sigma_t for x in W.fetch_eight(brick, leaf.p)]
                v000, v100, v010, v110, v001, v101, v011, v111 = v
                mxyz = v111 - v011 - v101 - v110 + v100 + v010 + v001 - v000
                mxy = v000 - v100 - v010 + v110
                mxz = v000 - v100 - v001 + v101
                myz = v000 - v010 - v001 + v011
                mx, my, mz = v100 - v000, v010 - v000, v001 - v000
                d = self.ray_d
                p0 = leaf.pos + leaf.tx * leaf.dir - (leaf.p.astype(F) + vmin_leaf)
                c3 = mxyz * d[0] * d[1] * d[2]
                c2 = (p0[2] * d[0] * d[1] + p0[1] * d[0] * d[2] + p0[0] * d[1] * d[2]) * mxyz + mxy * d[0] * d[1] + mxz * d[0] * d[2] + myz * d[1] * d[2]
                c1 = ((p0[1] * p0[2] * d[0] + p0[0] * p0[2] * d[1] + p0[0] * p0[1] * d[2]) * mxyz + mx * d[0] + my * d[1] + mz * d[2]
                      + (p0[1] * d[0] + p0[0] * d[1]) * mxy + (p0[2] * d[0] + p0[0] * d[2]) * mxz + (p0[2] * d[1] + p0[1] * d[2]) * myz)
                c0 = (p0[0] * p0[1] * p0[2] * mxyz + p0[0] * p0[1] * mxy + p0[0] * p0[2] * mxz + p0[1] * p0[2] * myz + p0[0] * mx + p0[1] * my
                      + p0[2] * mz + v000)
                delta = self._tau(maxDeltaT, c3, c2, c1, c0)
                for i in range(self.n):
                    if self.hit[i] == F(-1):
                        if self.opt + delta >= self.outTr[i]:
                            target = self.outTr[i] - self.opt
                            t_low, t_high, tau_low, tau_high, t_sol = F(0), maxDeltaT, F(0), delta, F(0)
                            k = 0
                            while k < 32 and t_high - t_low > maxDeltaT * F(0.001):
                                k += 1
                                t_sol = t_low + (t_high - t_low) * (target - tau_low) / (tau_high - tau_low)
                                tau = self._tau(t_sol, c3, c2, c1, c0)
                                if tau < target:
                                    t_low, tau_low = t_sol, tau
                                else:
                                    t_high, tau_high = t_sol, tau
                            self.hit[i] = t + t_sol
                            self.outTr[i] = np.exp(-self.outTr[i])
                            t2 = t_sol * t_sol
                            self.pdf[i] = (c3 * (t2 * t_sol) + c2 * t2 + c1 * t_sol + c0) * self.outTr[i]
                            finished += 1
                    else:
                        finished += 1
            else:
                density = W.density_in_atlas(brick, leaf.p.astype(F) + F(0.5), False)
                sigma_t = density * W.sigma_t
                for i in range(self.n):
                    if self.hit[i] == F(-1):
                        with np.errstate(divide="ignore", invalid="ignore"):
                            dT = -np.log(F(1) - self.rng.next1d()) / sigma_t
                            curT = t + dT
                        if np.isnan(curT) or np.isinf(curT):
                            curT = self.K_RAY_TMAX
                        if curT < currentTMax:
                            self.hit[i] = curT
                            self.outTr[i] = np.exp(-(dT * sigma_t + self.opt))
                            self.pdf[i] = sigma_t * self.outTr[i]
                            finished += 1
                    else:
                        finished += 1
                delta = maxDeltaT * sigma_t
            if finished == self.n:
                return True, t
            t = currentTMax
            self.opt = self.opt + delta
            if t >= self.tFar:
                self.end()
                return True, t
            leaf.step()
            it += 1
        return False, t

    def end(self):
        for i in range(self.n):
            if self.initialized:
                if self.hit[i] == F(-1):
                    self.hit[i] = self.K_RAY_TMAX
                    self.outTr[i] = np.exp(-self.opt)
                    self.pdf[i] = self.outTr[i]
            else:
                self.hit[i], self.outTr[i], self.pdf[i] = self.K_RAY_TMAX, F(1), F(1)


class CellSamplingAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:439-515 (SampleVolumeCellByDensityAdapterGVDB)
    """One voxel cell of the ray, chosen with probability proportional to exp(-optical depth so far) * sigma_t of the cell (weighted
    reservoir sampling with one draw per cell whose running sum is positive): the reprojection point of K2 for background samples."""

    def __init__(self, rng):
        self.rng = rng
        self.bound, self.interval, self.running, self.Tr = F(0), (F(-1), F(-1)), F(0), F(0)

    def start(self):
        self.interval = (F(-1), F(-1))

    def main(self, W, dda, vmin_leaf, brick, t):
        leaf = dda.copy()
        leaf.prepare_leaf(vmin_leaf)
        res0 = W.slot["res"][0]
        it = 0
        while it < MAX_BRICK_STEPS and bool(np.all((leaf.p >= 0) & (leaf.p < res0))):
            density = W.density_in_atlas(brick, leaf.p.astype(F) + F(0.5), False)
            leaf.next()
            maxDeltaT = leaf.ty - t
            sigma_t = density * W.sigma_t
            weight = np.exp(self.Tr) * sigma_t
            self.running = self.running + weight
            if self.running > 0 and self.rng.next1d() < weight / self.running:
                self.bound = density
                self.interval = (t, min(self.tFar, t + maxDeltaT))
            self.Tr = self.Tr + F(-maxDeltaT) * sigma_t
            if t + maxDeltaT >= self.tFar:
                return True, t
            t = t + maxDeltaT
            leaf.step()
            it += 1
        return False, t

    def end(self):
        pass


class ResidualRatioAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:710-794 (ResidualRatioTrackingGVDBAdapter)
    """Ratio tracking (global majorant), residual ratio tracking (per-brick control density from the brick's min / max / avg) and
    its analog variant (control = the brick minimum): transmittance estimate = prod over bricks of T_control * T_residual."""

    def __init__(self, analog, global_majorant, rng):
        self.Tr, self.analog, self.glob, self.rng = F(1), analog or global_majorant, global_majorant, rng

    def start(self):
        pass

    def main(self, W, dda, vmin_leaf, brick, t):
        st, b = W.sigma_t, self.bounds
        mu_min = F(0) if self.glob else b[0] * st
        mu_max = (W.slot["max_value"] * W.dsf) * st if self.glob else b[1] * st
        mu_avg = b[2] * st
        maxDeltaT = min(self.tFar - t, dda.ty - t)
        currentTMax = min(self.tFar, dda.ty)
        mu_r_temp = F(mu_max - mu_min)
        D = W.superVoxelDiagonal
        if mu_r_temp == 0 or self.analog:
            mu_c = mu_min
        else:
            # pow through float64 and one rounding (what a correctly rounded powf returns): numpy's float32 SIMD pow can be an ulp
            # off, and an ulp in mu_c decides whether a collision in a saturated voxel (mu == mu_max) multiplies T_r by exactly 0
            # or by 6e-8 — which in turn decides whether later segments are evaluated at all, i.e. how many numbers are drawn
            e = float(F(1) / (D * mu_r_temp))
            p2 = F(math.pow(2.0, e)) if e < 128.0 else F(np.inf)     # 2^(1 / small) overflows to inf; min() then picks mu_avg, as in the shader
            mu_c = min(mu_avg, max(mu_min, mu_min + mu_r_temp * (p2 - F(1))))
        mu_r = max(F(mu_c - mu_min), F(mu_max - mu_c))
        with np.errstate(divide="ignore"):
            inv_mu_r = F(1) / mu_r
        T_c = np.exp(F(-mu_c) * min(self.tFar - t, maxDeltaT))
        T_r = F(1)
        if mu_r > 0:
            while True:
                t = t - np.log(F(1) - self.rng.next1d()) * inv_mu_r
                if t >= currentTMax:
                    break
                p = (self.ray_o + t * self.ray_d) - vmin_leaf
                mu = W.density_in_atlas(brick, p, True) * st
                T_r = T_r * (F(1) - (mu - mu_c) * inv_mu_r)
        self.Tr = self.Tr * (T_c * T_r)
        return bool(currentTMax >= self.tFar), currentTMax

    def end(self):
        pass


class DecompositionAdapter:   # VR/VolumeTrackingAdapterGVDB.slang:606-707 (DecompositionTrackingGVDBAdapter)
    """Free-flight distance by decomposition tracking: an analytic flight through the brick's minimum density raced against delta
    tracking of the residual; hit = the medium-space parameter of the interaction (== the world parameter), None when the ray leaves."""

    def __init__(self, rng):
        self.rng, self.hit = rng, None

    def start(self):
        pass

    def main(self, W, dda, vmin_leaf, brick, t):
        st, b = W.sigma_t, self.bounds
        lo, hi = b[0], b[1]
        currentTMax = min(self.tFar, dda.ty)
        if lo == 0:
            t_control = DistanceSamplingAdapter.K_RAY_TMAX
        else:
            t_control = t - np.log(F(1) - self.rng.next1d()) / (lo * st)
        if hi - lo > 0:
            inv = F(1) / F(hi - lo)
            while True:
                t = t - np.log(F(1) - self.rng.next1d()) * inv / st
                if t >= t_control or t >= currentTMax:
                    break
                p = (self.ray_o + t * self.ray_d) - vmin_leaf
                density = W.density_in_atlas(brick, p, True)
                if (density - lo) * inv > self.rng.next1d():
                    self.hit = t
                    return True, t
            t = min(t_control, t)
            if t < currentTMax:
                self.hit = t
                return True, t
            t = currentTMax
            if t >= self.tFar:
                self.hit = None
                return True, t
            return False, t
        if t_control < currentTMax:
            self.hit = t_control
            return True, t_control
        return False, currentTMax

    def end(self):
        self.hit = None


class Witness:
    def __init__(self, grid_desc, slot_index, volume_desc=None, mip=None):
        """volume_desc / mip: for the previous frame's grids, which the shaders bind at slot offsets 19 (density) / 11 (temperature,
        velocity) under the CURRENT frame's volume description."""
        self.slot = slot_arrays(grid_desc.slots[slot_index])
        v = grid_desc.volume if volume_desc is None else volume_desc
        self.sigma_t = F(v.sigma_t)
        self.dsf = F(v.densityScaleFactorByScaling)
        self.tStepBase = F(v.tStep) * F(v.volumeWorldScaling)
        self.tex = _Texture(self.slot)
        self.mip = slot_index if mip is None else mip
        # F/Scene/Scene.cpp:3077-3080: 8 voxels times the length of the medium-to-world scale, from slot 0
        m2w = np.linalg.inv(np.array(list(grid_desc.slots[0].world_to_medium), dtype=np.float64).reshape(4, 4))
        self.superVoxelDiagonal = F(8.0 * np.sqrt((m2w[:3, :3] ** 2).sum()))

    # VR/VolumeBase.slang:252-263
    def density_in_atlas(self, brick, p, linear):
        s = self.tex.linear(brick, p) if linear else self.tex.point(brick, p)
        return s * self.slot["compress_scale"] * self.dsf

    def fetch_eight(self, brick, cell):
        out = []
        for i in range(8):
            raw = self.tex.texel(brick, int(cell[0]) + i % 2, int(cell[1]) + (i % 4) // 2, int(cell[2]) + i // 4)
            out.append(self.tex.decode(raw) * self.slot["compress_scale"] * self.dsf)
        return out

    # F/Scene/GVDB/gvdbNodes.slang:141-280 (getNode by position) + F/Scene/GVDB/gvdb.slang:6-31 (getValueAtPoint) +
    # VR/VolumeBase.slang:232-241 (Density / DensityWorldSpace): trilinear point query of this slot at a world-space position
    def density_world(self, p_world):
        s = self.slot
        pos = _mul_point(np.asarray(p_world, dtype=F), s["w2m"])
        if bool(np.any(pos < s["bmin"])) or bool(np.any(pos >= s["bmax"])):
            return F(0)
        return self._value_at(pos) * self.dsf

    def value_world(self, p_world, ch=0):
        """getValueAtPoint of this slot at a world-space position without the density wrapper (temperature / velocity grids: no
        bounding-box test, no density scale)."""
        return self._value_at(_mul_point(np.asarray(p_world, dtype=F), self.slot["w2m"]), ch)

    def _value_at(self, pos, ch=0):
        s = self.slot
        lev = s["top_lev"]
        vmin, link = self._node(lev, 0)
        while lev > 0:
            span = F(s["res"][lev] * s["vdel"][lev])            # noderange of the level: 4096 / 128 voxels
            if bool(np.any(pos < vmin)) or bool(np.any(pos >= vmin + span)):
                return F(0)
            pc = ((pos - vmin) / F(s["vdel"][lev])).astype(np.int64)
            dm = s["dim"][lev]
            b = (((int(pc[2]) << dm) + int(pc[1])) << dm) + int(pc[0])
            child = self._child(lev, link, b)
            if child == ID_UNDEFL:
                return F(0)
            lev -= 1
            vmin, link = self._node(lev, child)
        return self.tex.linear(link, pos - vmin, ch)

    def _node(self, lev, idx):
        n = self.slot["nodes"][lev][idx]
        return n["pos"].astype(F), int(n["link"])

    def _child(self, lev, link, b):
        if link == ID_UNDEFL:
            return ID_UNDEFL
        r3 = self.slot["res"][lev] ** 3
        idx = link * r3 + b
        lst = self.slot["child"][lev]
        return int(lst[idx]) if 0 <= idx < len(lst) else 0     # ByteAddressBuffer.Load outside the buffer returns 0

    # VR/VolumeUtils.slang:171-282
    def track(self, origin_w, dir_w, tmax, adapter, vertex_center):
        s = self.slot
        eps = F(0.01)
        lev = top = s["top_lev"]
        nodeid = [0, 0, 0]
        tMaxL = [F(0), F(0), F(0)]
        o = _mul_point(np.asarray(origin_w, dtype=F), s["w2m"])
        d = _mul_vec(np.asarray(dir_w, dtype=F), s["w2m"])
        if vertex_center:
            o = o - F(0.5)
        bmin, bmax = s["bmin"].copy(), s["bmax"].copy()
        if vertex_center:
            bmin, bmax = bmin - F(0.5), bmax - F(0.5)
        # Bounds3f::IntersectP (VR/VolumeBase.slang:137-158)
        t0, t1 = F(0), F(tmax)
        for i in range(3):
            inv = F(1) / d[i]
            tn = (bmin[i] - o[i]) * inv
            tf = (bmax[i] - o[i]) * inv
            if tn > tf:
                tn, tf = tf, tn
            t0 = tn if tn > t0 else t0
            t1 = tf if tf < t1 else t1
            if t0 > t1:
                adapter.end()
                return
        tNear, tFar = t0, t1
        adapter.tNear, adapter.tFar, adapter.ray_o, adapter.ray_d = tNear, tFar, o, d
        adapter.start()
        vmin, link = self._node(lev, 0)
        tMaxL[lev] = tFar
        dda = _HDDA()
        dda.set_from_ray(o, d, tNear + eps, tFar)
        dda.prepare(vmin, s["vdel"][lev])
        if vertex_center:
            it = 0
            while it < 3 and bool(np.any((dda.p < 0) | (dda.p > s["res"][lev]))):
                it += 1
                dda.next(); dda.step(); dda.tx = dda.tx + eps
        t = tNear
        links = {lev: link}
        vmins = {lev: vmin}
        it = 0
        while it < 4096 and 0 < lev <= top and bool(np.all((dda.p >= 0) & (dda.p <= s["res"][lev]))):
            dda.next()
            dm = s["dim"][lev]
            b = (((int(dda.p[2]) << dm) + int(dda.p[1])) << dm) + int(dda.p[0])
            child = self._child(lev, links[lev], b)
            if child != ID_UNDEFL:
                if lev == 1:
                    nodeid[0] = child
                    t = dda.tx - eps
                    vmin_leaf, brick = self._node(0, child)
                    adapter.bounds = self.slot["nodes"][0][child]["bounds"].astype(F) * self.dsf     # VR/VolumeUtils.slang:251
                    stop, t = adapter.main(self, dda, vmin_leaf, brick, t)
                    if stop:
                        return
                    dda.step(); dda.tx = dda.tx + eps
                else:
                    lev -= 1
                    nodeid[lev] = child
                    vmins[lev], links[lev] = self._node(lev, child)
                    tMaxL[lev] = dda.ty
                    dda.prepare(vmins[lev], s["vdel"][lev])
            else:
                dda.step(); dda.tx = dda.tx + eps
            while lev <= top and dda.tx > tMaxL[lev]:
                lev += 1
                if lev <= top:
                    dda.prepare(vmins[lev], s["vdel"][lev])
            it += 1
        adapter.end()

    def ray_marching(self, origin_w, dir_w, tmax, linear=True, tstep_scale=1.0):
        eff = self.mip - 19 if self.mip >= 19 else self.mip
        eff = eff - 8 if eff >= 8 else eff
        a = RayMarchingAdapter(linear, self.tStepBase * F(tstep_scale) * F(eff + 1))
        self.track(origin_w, dir_w, tmax, a, False)
        return float(a.Tr)

    def feature_transmittance(self, origin_w, dir_w, tstep_scale=1.0, linear=True):
        """ReservoirFeatureRayMarchingGeneric (VR/VolumeUtils.slang:366-380) on this mip."""
        a = FeatureMarchAdapter(linear, self.tStepBase * F(tstep_scale))
        self.track(origin_w, dir_w, DistanceSamplingAdapter.K_RAY_TMAX, a, False)
        return float(a.accu)

    def sample_distances(self, origin_w, dir_w, num_samples, linear, rng):
        """SampleMediumAnalytic (VR/VolumeUtils.slang: the linear sampler draws its optical-depth targets first and walks the
        vertex-centred grid, the point sampler draws per cell): returns (hit distances, pdfs, transmittances) of the samples."""
        a = DistanceSamplingAdapter(num_samples, linear, rng)
        if linear:
            for i in range(num_samples):
                a.outTr[i] = -np.log(F(1) - rng.next1d())
        self.track(origin_w, dir_w, DistanceSamplingAdapter.K_RAY_TMAX, a, linear)
        return [float(x) for x in a.hit], [float(x) for x in a.pdf], [float(x) for x in a.outTr]

    def residual_ratio_tracking(self, origin_w, dir_w, tmax, rng, analog=False, global_majorant=False):
        """MediumTrResidualRatioTrackingGeneric (VR/VolumeUtils.slang:328-339)."""
        a = ResidualRatioAdapter(analog, global_majorant, rng)
        self.track(origin_w, dir_w, tmax, a, False)
        return float(a.Tr)

    def sample_supervoxel(self, origin_w, dir_w, rng):
        """SampleMediumSuperVoxelGeneric (VR/VolumeUtils.slang:315-326): the free-flight distance of the reference path tracer."""
        a = DecompositionAdapter(rng)
        self.track(origin_w, dir_w, DistanceSamplingAdapter.K_RAY_TMAX, a, False)
        return None if a.hit is None else float(a.hit)

    def rejection_sample_point(self, origin_w, dir_w, rng):
        """RejectionSampleRandomPointByDensity (VR/VolumeUtils.slang:572-581): a depth inside the selected cell, or kRayTMax."""
        a = CellSamplingAdapter(rng)
        self.track(origin_w, dir_w, DistanceSamplingAdapter.K_RAY_TMAX, a, False)
        if a.interval[0] == F(-1):
            return DistanceSamplingAdapter.K_RAY_TMAX
        depth = a.interval[0] + (a.interval[1] - a.interval[0]) * rng.next1d()
        rng.next1d()          # sampledY (unused)
        return F(depth)

    def analytic(self, origin_w, dir_w, tmax, linear=True):
        a = AnalyticAdapter(linear)
        self.track(origin_w, dir_w, tmax, a, linear)
        return float(a.Tr)
