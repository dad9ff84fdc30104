"""The C++ oracle's two deterministic transmittance estimators (ray marching = every p-hat of the reuse stages; analytic
tracking = final shading) against an independent Python restatement written from the Slang (oracle/march_witness.py), on random
rays through 2- and 3-level grids, fp32 and UNORM8 pools, trilinear and point sampling.  CPU only."""
import numpy as np
import pytest

from common import env_scene
from oracle import vro
from oracle.march_witness import Witness
from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi


def _rays(scene, n, seed):
    lo, hi = scene.volume_bounds_world()
    c, ext = 0.5 * (lo + hi), hi - lo
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        o = c + (rng.random(3) - 0.5) * ext * (2.2 if len(out) % 3 else 0.6)     # a third of the rays start inside the volume
        tgt = c + (rng.random(3) - 0.5) * ext * 0.8
        d = tgt - o
        d /= np.linalg.norm(d)
        if np.abs(d).min() < 1e-3:
            continue
        tmax = float(3.4e38 if len(out) % 2 else np.linalg.norm(tgt - o) * rng.uniform(0.3, 1.2))
        out.append((o.astype(np.float32), d.astype(np.float32), tmax))
    return out


@pytest.mark.parametrize("dim,three_level", [((64, 64, 56), False), ((200, 150, 140), True)])
def test_oracle_transmittance_matches_slang_witness(dim, three_level):
    sc = env_scene(dim=dim, density_scale=0.004 if not three_level else 0.0015, num_mips=3)   # optical depths of order 1: the sums matter
    grid = sc.volume.grid.contents
    assert (grid.slots[0].top_lev == 2) == three_level
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16, importance=np.zeros(349525, np.float32))
    n = 24 if three_level else 40
    cases = [("march", 1, True, 1.0), ("march", 0, True, 2.0), ("march", 2, False, 1.0), ("march", 9, True, 1.0),
             ("analytic", 0, True, 0.0), ("analytic", 9, False, 0.0)]
    nontrivial = 0
    for kind, mip, linear, scale in cases:
        w = Witness(grid, mip)
        for o, d, tmax in _rays(sc, n, seed=mip * 10 + int(linear)):
            if kind == "march":
                got = op.transmittance(o, d, tmax, capi.kRayMarching, mip, linear, scale)
                want = w.ray_marching(o, d, tmax, linear, scale)
            else:
                got = op.transmittance(o, d, tmax, capi.kAnalyticTracking, mip, linear)
                want = w.analytic(o, d, tmax, linear)
            assert np.isclose(got, want, rtol=4e-6, atol=1e-30), (kind, mip, linear, o, d, tmax, got, want)
            nontrivial += 0.02 < want < 0.98
    assert nontrivial > len(cases) * n * 0.4      # most rays end with an intermediate transmittance: the accumulated sums are compared


def test_xoshiro_witness_matches_the_oracle_stream():
    """The witness's own generator (written from the Slang / the public-domain xoshiro128**) against the oracle's words."""
    from oracle.march_witness import Xoshiro
    for px, py, n in ((0, 0, 0), (17, 5, 3), (1919, 1079, 4095), (65535, 65535, 2 ** 31)):
        words = np.zeros(16, dtype=np.uint32); floats = np.zeros(16, dtype=np.float32)
        vro.lib().vro_rng_words(px, py, n, 16, words.ctypes.data, floats.ctypes.data)
        g = Xoshiro(px, py, n)
        assert [g.next() for _ in range(16)] == [int(w) for w in words]
        g = Xoshiro(px, py, n)
        assert [float(g.next1d()) for _ in range(16)] == [float(f) for f in floats]


@pytest.mark.parametrize("dim,three_level", [((64, 64, 56), False), ((200, 150, 140), True)])
def test_oracle_distance_sampling_matches_slang_witness(dim, three_level):
    """K1's candidate generation: free-flight distances along a ray (up to 4 samples, point sampler on the conservative mip and
    trilinear sampler on the plain mip), each ray with its own random-number stream: hit distances, pdfs, transmittances and the
    number of draws consumed (the generator state after the call) must agree."""
    from oracle.march_witness import Xoshiro
    sc = env_scene(dim=dim, density_scale=0.02 if not three_level else 0.008, num_mips=3)
    grid = sc.volume.grid.contents
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16, importance=np.zeros(349525, np.float32))
    n_rays = 14 if three_level else 24
    hits = exits = 0
    for mip, linear, ns in ((9, False, 4), (8, False, 1), (1, True, 4), (10, False, 2)):
        w = Witness(grid, mip)
        for k, (o, d, _) in enumerate(_rays(sc, n_rays, seed=100 + mip)):
            seed = (k * 7 + 1, k * 3 + 2, k + mip)
            hd, pd, ot, state = op.sample_distances(o, d, mip, linear, ns, seed)
            rng = Xoshiro(*seed)
            whd, wpd, wot = w.sample_distances(o, d, ns, linear, rng)
            assert [int(x) for x in state] == rng.s, (mip, linear, k)          # same number of draws
            for i in range(ns):
                if whd[i] > 1e37:
                    assert hd[i] > 1e37; exits += 1
                else:
                    assert np.isclose(hd[i], whd[i], rtol=3e-6), (mip, linear, k, i, hd[i], whd[i]); hits += 1
                assert np.isclose(pd[i], wpd[i], rtol=2e-5, atol=1e-30) and np.isclose(ot[i], wot[i], rtol=2e-5, atol=1e-30), (mip, linear, k, i)
            assert all(hd[i] == 0 and pd[i] == 0 and ot[i] == 0 for i in range(ns, 4))
    assert hits > 40 and exits > 10          # both outcomes are exercised


@pytest.mark.parametrize("dim,three_level", [((64, 64, 56), False), ((200, 150, 140), True)])
def test_oracle_stochastic_trackers_match_slang_witness(dim, three_level):
    """The random-walk estimators: ratio tracking (global majorant), residual ratio tracking and its analog variant (per-brick
    control densities from the node bounds) as transmittance estimators, and the decomposition tracker that draws the reference
    path tracer's free-flight distances — same random-number stream, so the same walk: estimates agree to rounding, the
    decomposition tracker also in the number of draws it consumed."""
    from oracle.march_witness import Xoshiro
    sc = env_scene(dim=dim, density_scale=0.004 if not three_level else 0.0015, num_mips=3)     # optical depths of order 1
    grid = sc.volume.grid.contents
    assert float(Witness(grid, 0).superVoxelDiagonal) == pytest.approx(float(grid.volume.superVoxelWorldSpaceDiagonalLength), rel=1e-6)
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16, importance=np.zeros(349525, np.float32))
    n_rays = 12 if three_level else 24
    nontrivial = 0
    for method, mip in ((capi.kRatioTracking, 0), (capi.kResidualRatioTracking, 0), (capi.kAnalogResidualRatioTracking, 1), (capi.kResidualRatioTracking, 2)):
        w = Witness(grid, mip)
        for k, (o, d, tmax) in enumerate(_rays(sc, n_rays, seed=200 + mip + method)):
            seed = (k * 5 + 3, k * 11 + 1, k + method)
            got = op.transmittance(o, d, tmax, method, mip, True, 1.0, seed)
            want = w.residual_ratio_tracking(o, d, tmax, Xoshiro(*seed), analog=method == capi.kAnalogResidualRatioTracking,
                                             global_majorant=method == capi.kRatioTracking)
            assert np.isclose(got, want, rtol=2e-5, atol=1e-7), (method, mip, k, got, want)
            nontrivial += 0.02 < want < 0.98
    assert nontrivial > 4 * n_rays * 0.35
    sc = env_scene(dim=dim, density_scale=0.02 if not three_level else 0.008, num_mips=3)       # dense enough to scatter
    grid = sc.volume.grid.contents
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16, importance=np.zeros(349525, np.float32))
    hits = exits = 0
    for mip in (0, 1):
        w = Witness(grid, mip)
        for k, (o, d, _) in enumerate(_rays(sc, n_rays, seed=300 + mip)):
            seed = (k * 13 + 2, k * 3 + 5, k + 40 + mip)
            t, state = op.sample_supervoxel(o, d, mip, seed)
            rng = Xoshiro(*seed)
            wt = w.sample_supervoxel(o, d, rng)
            assert [int(x) for x in state] == rng.s, (mip, k)
            if wt is None:
                assert t is None; exits += 1
            else:
                assert t == pytest.approx(wt, rel=5e-6, abs=1e-6), (mip, k); hits += 1     # the hook returns |p - origin|
    assert hits > n_rays // 2 and exits > 2
