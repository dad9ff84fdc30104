"""Self-made pins of the oracle (SURVEY.md section 8c "additional self-made pins"): closed forms, round trips, unbiasedness."""
import ctypes as C
import math

import numpy as np
import pytest

from common import capi, config1_params, config1_scene, env_scene, rel_mse
from oracle import vro
from volumetricrestirrelease_b200 import Scene, VolumetricReSTIRParams


def _box_scene(n=32, value=0.5, density_scale=0.1):
    sc = Scene()
    d = np.full((n, n, n), value, np.float32)
    sc.addGVDBVolume(dense=d, numMips=2, densityScale=density_scale, sigma_a=(1, 1, 1), sigma_s=(9, 9, 9))
    sc.addDirectionalLight((0, -1, 0), (1, 1, 1))
    sc.frame_camera(1.5)
    return sc


def test_closed_form_transmittance_constant_box():
    """T = exp(-sigma_t * rho * L) through a constant-density box for every tracker (analytic exact, ray march by step
    count, ratio / residual-ratio in expectation)."""
    n, rho, ds = 32, 0.5, 0.01
    sc = _box_scene(n, rho, ds)
    op = vro.OraclePass(VolumetricReSTIRParams(mUseEnvironmentLights=0, mUseAnalyticLights=1))
    op.setScene(sc, 16, 16)
    o, d = (-40.0, 0.3, 0.7), (1.0, 0.0, 0.0)
    mu = 10.0 * rho * ds
    # the 1-voxel apron means trilinear values fall to 0 half a voxel outside; analytic-trilinear integrates the ramp exactly
    expect_point = math.exp(-mu * n)
    t_point = op.transmittance(o, d, 3.4e38, capi.kAnalyticTracking, 0, linear=False)
    assert abs(t_point - expect_point) / expect_point < 1e-4
    t_lin = op.transmittance(o, d, 3.4e38, capi.kAnalyticTracking, 0, linear=True)
    # vertex-centred traversal: the entry half cell [0, 0.5) is skipped by the "step until inside" loop (VR/VolumeUtils.slang:216-225),
    # the interior contributes n - 1 voxels and the exit ramp [n - 0.5, n] integrates to 0.375
    assert abs(t_lin - math.exp(-mu * (n - 0.625))) / t_lin < 1e-4
    t_march = op.transmittance(o, d, 3.4e38, capi.kRayMarching, 0, linear=False, tstep_scale=0.25)
    assert abs(t_march - expect_point) / expect_point < 0.05
    for method in (capi.kRatioTracking, capi.kResidualRatioTracking, capi.kAnalogResidualRatioTracking):
        vals = [op.transmittance(o, d, 3.4e38, method, 0, seed=(i, 7, 3)) for i in range(4000)]
        m, s = float(np.mean(vals)), float(np.std(vals) / math.sqrt(len(vals)))
        assert abs(m - t_lin) <= 5 * s + 1e-4 * t_lin, (method, m, t_lin, s)


def test_analytic_vs_ratio_tracking_heterogeneous():
    """Analytic trilinear tracking equals the mean of 1e4 ratio-tracking estimates on one ray through the fBm sphere."""
    sc = config1_scene()
    op = vro.OraclePass(config1_params())
    op.setScene(sc, 16, 16)
    cam = np.array(sc.camera.position)
    tgt = np.array(sc.camera.target)
    d = (tgt - cam) / np.linalg.norm(tgt - cam)
    ta = op.transmittance(tuple(cam), tuple(d), 3.4e38, capi.kAnalyticTracking, 0, linear=True)
    vals = [op.transmittance(tuple(cam), tuple(d), 3.4e38, capi.kResidualRatioTracking, 0, seed=(i, 1, 2)) for i in range(10000)]
    m, s = float(np.mean(vals)), float(np.std(vals) / math.sqrt(len(vals)))
    assert 0.0 < ta < 1.0
    assert abs(m - ta) <= 5 * s + 2e-3 * ta


def test_axis_parallel_rays_do_not_spin():
    """Rays with an exactly-zero direction component (0*inf in the reference's DDA step) traverse correctly."""
    sc = _box_scene()
    op = vro.OraclePass(VolumetricReSTIRParams(mUseEnvironmentLights=0, mUseAnalyticLights=1))
    op.setScene(sc, 16, 16)
    xyz, t = op.brick_visits((-40.0, 0.3, 0.7), (1.0, 0.0, 0.0), 0)
    assert len(t) == 4 and np.all(np.diff(t) > 0)
    a = op.transmittance((-40.0, 0.3, 0.7), (1.0, 0.0, 0.0), 3.4e38, capi.kRayMarching, 1, linear=True)
    b = op.transmittance((-40.0, 0.3, 0.7), (1.0, 1e-7, 0.0), 3.4e38, capi.kRayMarching, 1, linear=True)
    assert np.isfinite(a) and abs(a - b) < 1e-3


def test_hdda_brick_visits_are_ordered_and_inside_the_ray_box():
    sc = env_scene(dim=(200, 180, 150), density_scale=0.2, num_mips=2)
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16)
    cam = np.array(sc.camera.position)
    for k in range(8):
        tgt = np.array(sc.camera.target) + np.array([k - 4, 2 * k - 7, k]) * 3.0
        d = (tgt - cam) / np.linalg.norm(tgt - cam)
        for vc in (False, True):
            xyz, t = op.brick_visits(tuple(cam), tuple(d), 0, vertex_center=vc)
            assert np.all(np.diff(t) > 0)
            assert len({tuple(p) for p in xyz}) == len(xyz)        # no brick entered twice
            assert np.all(xyz % 8 == 0)


def test_bit_packing_round_trips():
    L = vro.lib()
    for bounces in (0, 1, 3, 7, 2047):
        for storage in (0, 0x12345, 0xFFFFF, 0x7A5A5):
            s = L.vro_encode_max_indirect_bounces(storage, bounces)
            assert L.vro_decode_max_indirect_bounces(s, 4) == bounces
            assert (s & 0xFFFFF) == (storage & 0xFFFFF)
            assert L.vro_decode_max_indirect_bounces(s, 1) == 0      # MAX_BOUNCES == 1 ignores the field
    for tag in range(16):
        s = L.vro_encode_path_tag(0x00300000 | 0xBEEF, tag)
        assert L.vro_decode_path_tag(s) == tag and (s >> 20) == 3 and (s & 0xFFFF) == 0xBEEF
    rng = np.random.default_rng(0)
    for _ in range(200):
        v = rng.normal(size=3)
        v /= np.linalg.norm(v)
        dist = float(rng.uniform(0.1, 50))
        i4 = (C.c_float * 4)(v[0], v[1], v[2], dist)
        o3 = (C.c_float * 3)()
        o4 = (C.c_float * 4)()
        L.vro_encode_wi_dist(C.byref(i4), C.byref(o3))
        L.vro_decode_wi_dist(C.byref(o3), C.byref(o4))
        np.testing.assert_allclose(list(o4), [v[0], v[1], v[2], dist], atol=2e-3 / max(abs(v[2]), 0.05))
        assert math.copysign(1, o4[2]) == math.copysign(1, v[2]) or abs(v[2]) < 1e-3


def test_neighbor_offsets_r2_and_hammersley():
    p = VolumetricReSTIRParams()
    op = vro.OraclePass(p)
    off = op.neighbor_offsets(3, 0, 4)
    assert tuple(off[0]) == (0, 0)                               # sample 0 is the centre pixel itself
    assert np.all(np.abs(off) <= 10)
    seed = ((1 + 1) * 3 + 0) % 16
    for i in (1, 2, 3):
        m = float(seed * 4 + i)
        ux, uy = (0.754877669 * m) % 1.0, (0.569840296 * m) % 1.0
        r, phi = np.float32(np.sqrt(np.float32(ux))), np.float32(2 * np.pi) * np.float32(uy)
        ex, ey = int(np.float32(10) * (r * np.cos(phi))), int(np.float32(10) * (r * np.sin(phi)))
        assert abs(off[i][0] - ex) <= 1 and abs(off[i][1] - ey) <= 1
    assert not np.array_equal(op.neighbor_offsets(4, 0, 4), off)
    p2 = VolumetricReSTIRParams(mRandomSamplerType=capi.kHammersley)
    assert np.all(np.abs(vro.OraclePass(p2).neighbor_offsets(0, 0, 4)) <= 10)


@pytest.mark.parametrize("params", [
    VolumetricReSTIRParams(),
    VolumetricReSTIRParams(mMaxBounces=2),
    VolumetricReSTIRParams(mEnableTemporalReuse=0, mEnableSpatialReuse=0),
    VolumetricReSTIRParams(mMaxBounces=3, mVertexReuse=1, mVertexReuseStartBounce=1),    # VERTEX_REUSE: reconnection at world-space vertices
    VolumetricReSTIRParams(mMaxBounces=3, mVertexReuse=1, mVertexReuseStartBounce=2),
])
def test_restir_is_unbiased_against_the_reference_path_tracer(params):
    """The pass's own notion of ground truth: mUseReference (brute-force volumetric path tracer)."""
    w, h, n = 48, 40, 192
    sc = env_scene(dim=(48, 48, 40), density_scale=0.25, env_size=(128, 64))

    def mean_image(p):
        op = vro.OraclePass(p)
        op.setScene(sc, w, h)
        acc = np.zeros((h, w, 4), np.float64)
        for _ in range(n):
            acc += op.execute()
        return (acc / n).astype(np.float32)

    ref = mean_image(VolumetricReSTIRParams(mUseReference=1, mMaxBounces=params.mMaxBounces))
    img = mean_image(params)
    ratio = img[..., :3].mean() / ref[..., :3].mean()
    assert abs(ratio - 1) < 0.01, ratio
    assert rel_mse(img, ref) < 0.02


def test_options_change_resets_frame_counter_and_history():
    sc = config1_scene(dim=32, num_mips=3)
    op = vro.OraclePass(config1_params())
    op.setScene(sc, 16, 16)
    a = op.execute()
    b = op.execute()
    assert not np.array_equal(a, b)                  # frame counter advanced -> different seeds
    op.updateDict({"mInitialM": 4})                  # updateDict resets mFrameCount (VR/VolumetricReSTIR.cpp:1339,349-359)
    c = op.execute()
    assert np.array_equal(a, c)
    with pytest.raises(RuntimeError):
        op.updateDict({"mInitialLightingMipLevel": 7})
        op.execute()
