"""The oracle's K1 (initial candidates + reservoir streaming + final p-hat) against oracle/stage_witness.py: TraceRays.cs.slang /
ComputeInitialSample.slang restated in Python for one bounce and an env-map light, on top of the independent witnesses for
free-flight sampling, the hierarchical env sampler, the phase function, transmittance and the xoshiro stream."""
import numpy as np
import pytest

from common import FEAT, RES, env_scene
from oracle import light_witness as lw
from oracle import stage_witness as sw
from oracle import vro
from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi


@pytest.mark.parametrize("kw", [dict(), dict(mInitialM=6, mInitialLightingMipLevel=1, mInitialBaseMipLevel=2),
                                dict(mInitialVisibilityUseLinearSampler=1, mInitialM=3, mInitialLightingUseLinearSampler=0),
                                dict(mInitialLightingTrackingMethod=capi.kResidualRatioTracking, mInitialLightSamples=2, mInitialM=3,
                                     mSpatialVisibilityTrackingMethod=capi.kRatioTracking, mSpatialLightingTrackingMethod=capi.kAnalyticTracking),
                                dict(mInitialLightingTrackingMethod=capi.kAnalyticTracking, mInitialLightingMipLevel=0, mInitialLightSamples=0)])
def test_initial_sampling_matches_the_slang_witness(kw):
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()                                                # frame 0, so that the frame counter is not 0
    frame_count = op.frame_count()
    op.record_k1_generator()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(3)
    vy, vx = np.nonzero((res["runningSum"] > 0) & (res["depth"] < 1e37))     # the streamed sample scatters in the medium
    by, bx = np.nonzero((res["runningSum"] > 0) & (res["depth"] > 1e37))     # ... or is the background behind it
    picks = [(int(vx[k]), int(vy[k])) for k in rng.permutation(len(vy))[:9]] + [(int(bx[k]), int(by[k])) for k in rng.permutation(len(by))[:3]]
    # A random-walk tracker inside K1 makes the DRAW COUNT depend on the last bit of powf (the control density of residual ratio
    # tracking): an estimate that is exactly 0 in one implementation and 1e-40 in the other adds or skips one reservoir draw, and
    # every later draw of that pixel shifts.  Such pixels are recognised by their generator and bounded, not compared.
    stochastic_k1 = kw.get("mInitialLightingTrackingMethod", capi.kRayMarching) in (capi.kRatioTracking, capi.kResidualRatioTracking, capi.kAnalogResidualRatioTracking)
    checked = volume = diverged = 0
    for x, y in picks:
        got = res[y, x]
        info = {}
        want = sw.initial_sampling_pixel(frame, x, y, frame_count, mips, info)
        after_candidates, after_p_hat = op.k1_generator(x, y)
        if stochastic_k1 and after_candidates != info["generator_after_candidates"]:
            diverged += 1
            continue
        assert after_candidates == info["generator_after_candidates"] and after_p_hat == info["generator_after_p_hat"], (x, y)   # same number of draws
        assert int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=1e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=5e-5, abs=1e-12), (x, y)
        checked += 1; volume += float(got["depth"]) < 1e37
    assert checked + diverged == 12 and diverged <= 1 and volume >= 5


@pytest.mark.parametrize("kw,density", [(dict(), 0.06), (dict(mInitialVisibilityTStepScale=2.0), 0.6)])
def test_features_match_the_slang_witness(kw, density):
    """K0 (GenerateFeatures.cs.slang:57-102): per-pixel camera-ray transmittance by ray marching mip 0 with the trilinear sampler and
    the 1 % early out; pixels that miss the volume keep 1."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=density, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    feat = op.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(h, w)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(5)
    vy, vx = np.nonzero(feat["transmittance"] != 1.0)
    my, mx = np.nonzero(feat["transmittance"] == 1.0)
    picks = [(int(vx[k]), int(vy[k])) for k in rng.permutation(len(vy))[:16]] + [(int(mx[k]), int(my[k])) for k in rng.permutation(len(my))[:4]]
    opaque = 0
    for x, y in picks:
        want = frame.wit(0).feature_transmittance(frame.origin, frame.ray_dir(x, y), params.mInitialVisibilityTStepScale)
        assert int(feat[y, x]["noReflectiveSurface"]) == 1
        got = float(feat[y, x]["transmittance"])                 # compared as optical depths: the sum is what rounds
        assert np.log(got) == pytest.approx(np.log(want), rel=5e-6, abs=2e-6), (x, y)
        opaque += want < 0.01
    assert density < 0.1 or opaque >= 4            # the dense variant reaches the early out


@pytest.mark.parametrize("kw,g", [(dict(mMaxBounces=3), 0.0), (dict(mMaxBounces=4, mInitialUseCoarserGridForIndirectBounce=0, mInitialM=3), 0.5),
                                  (dict(mMaxBounces=3, mInitialVisibilityUseLinearSampler=1, mInitialM=2), -0.3)])
def test_multi_bounce_paths_match_the_slang_witness(kw, g):
    """MAX_BOUNCES > 1: the bounce loop of ComputeInitialSample (a light sample at every vertex, phase-sampled continuation, one
    free-flight sample per bounce on the coarser grid, Russian roulette from the third vertex on, per-path reservoir), the
    extra-bounce records, the vertex loop of evaluate_F_ for the stored path's p-hat, and its final shading."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=g)
    params = VolumetricReSTIRParams(mEnableSpatialReuse=0, **kw)
    B = params.mMaxBounces
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy()
    op.execute_stage(5, 0, color)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(6)
    depth_of = res["sampledPixel"] >> 20
    picks = []
    for k, n in ((0, 3), (1, 4), (2, 4)) + (((3, 3),) if B > 3 else ()):
        ys, xs = np.nonzero((res["runningSum"] > 0) & (depth_of == k) & (res["depth"] < 1e37))
        assert len(ys) >= n, (k, len(ys))
        picks += [(int(xs[i]), int(ys[i])) for i in rng.permutation(len(ys))[:n]]
    for x, y in picks:
        got = res[y, x]
        want, want_extra = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, mips)
        assert int(got["sampledPixel"]) == want["sampledPixel"] and int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        k = int(got["sampledPixel"]) >> 20
        np.testing.assert_allclose(extra[y, x, :k], want_extra[:k], rtol=1e-5, atol=3e-6)
        np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        rad = sw.final_shading_path(frame, x, y, got, extra[y, x])
        np.testing.assert_allclose(color[y, x, :3], rad, rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("B", [1, 3])
def test_no_reuse_mode_matches_the_slang_witness(B):
    """Both reuse passes off (gNoReuse, BASELINE's configuration 1): candidates are paths whose vertices come from decomposition
    tracking on mip 0, densities and segment transmittances cancel (pdf = Tr = 1, albedo per vertex), and the final shading
    evaluates F in the same mode."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=0.3)
    params = VolumetricReSTIRParams(mEnableTemporalReuse=0, mEnableSpatialReuse=0, mMaxBounces=B, mInitialM=2)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy() if B > 1 else np.zeros((h, w, 1, 3), np.float32)
    op.execute_stage(5, 0, color)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(8)
    ys, xs = np.nonzero((res["runningSum"] > 0) & (res["depth"] < 1e37))
    by, bx = np.nonzero((res["runningSum"] > 0) & (res["depth"] > 1e37))
    picks = [(int(xs[i]), int(ys[i])) for i in rng.permutation(len(ys))[:10]] + [(int(bx[i]), int(by[i])) for i in rng.permutation(len(by))[:2]]
    deep = 0
    for x, y in picks:
        got = res[y, x]
        want, want_extra = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, mips)
        assert int(got["sampledPixel"]) == want["sampledPixel"] and int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        k = int(got["sampledPixel"]) >> 20
        deep += k > 0
        np.testing.assert_allclose(extra[y, x, :k], want_extra[:k], rtol=1e-5, atol=3e-6)
        np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        np.testing.assert_allclose(color[y, x, :3], sw.final_shading_path(frame, x, y, got, extra[y, x]), rtol=2e-4, atol=1e-9)
    assert B == 1 or deep >= 3


@pytest.mark.parametrize("B", [1, 3])
def test_analytic_and_emissive_lights_match_the_slang_witness(B):
    """All three light types at once (BASELINE's configuration 4 in small: emissive triangles drawn from the power alias table, a
    point and a directional light, the env map): type selection, the per-type samples and pdfs (area measure for triangles, Dirac
    lights with pdf 1), the offset shadow-ray origin on the triangle, and evaluate_L_in_volume of the stored (lightID, lightUV)
    for p-hat and the final shading."""
    from oracle import alias_oracle
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.1, env_size=(128, 64), g=0.3)
    lo, hi = sc.volume_bounds_world()
    sc.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
    sc.addDirectionalLight((0.3, -1.0, 0.2), (2.0, 1.5, 1.0))
    tris = sc.addEmissiveShell(300, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
    params = VolumetricReSTIRParams(mEnableSpatialReuse=0, mMaxBounces=B, mInitialM=4, mUseAnalyticLights=1, mUseEmissiveLights=1)
    alias = alias_oracle.emissive_alias_table(tris)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h, emissive_alias=alias)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy() if B > 1 else np.zeros((h, w, 1, 3), np.float32)
    op.execute_stage(5, 0, color)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    frame.lights = sw.Lights(sc, alias)
    rng = np.random.default_rng(9)
    vol = (res["runningSum"] > 0) & (res["depth"] < 1e37)
    picks = []
    for mask, n in ((vol & (res["lightID"] < 0), 4), (vol & (res["lightID"] >= 0) & (res["lightID"] < 2), 5), (vol & (res["lightID"] >= 2), 6)):
        ys, xs = np.nonzero(mask)
        assert len(ys) >= n, len(ys)
        picks += [(int(xs[i]), int(ys[i])) for i in rng.permutation(len(ys))[:n]]
    for x, y in picks:
        got = res[y, x]
        want, want_extra = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, mips) if B > 1 else (sw.initial_sampling_pixel(frame, x, y, frame_count, mips), None)
        assert int(got["sampledPixel"]) == want["sampledPixel"] and int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        k = int(got["sampledPixel"]) >> 20
        if k:
            np.testing.assert_allclose(extra[y, x, :k], want_extra[:k], rtol=1e-5, atol=3e-6)
        np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        np.testing.assert_allclose(color[y, x, :3], sw.final_shading_path(frame, x, y, got, extra[y, x]), rtol=2e-4, atol=1e-9)


@pytest.mark.parametrize("B", [1, 3])
def test_volume_emission_matches_the_slang_witness(B):
    """A plume with a temperature grid (BASELINE's configuration 3 in small): black-body emission at the scatter point, the
    emission-vs-scatter draw, self-emission samples (sigma_a * Le) and — with several bounces — emissive scatter vertices stored as
    positions in (lightID, lightUV) with the path tag, through p-hat and the final shading."""
    from volumetricrestirrelease_b200 import Scene
    w, h = 40, 30
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=0.0, dataFile="plume", numMips=4, densityScale=0.1, hasVelocity=True,
                     hasEmission=True, LeScale=0.3, temperatureCutoff=1.0, temperatureScale=100.0, dim=(64, 96, 64), seed=3, voxelSize=1.0)
    sc.setEnvMap((128, 64), seed=7)
    sc.setEnvMapIntensity(0.5)
    sc.frame_camera(1.0)
    params = VolumetricReSTIRParams(mEnableSpatialReuse=0, mMaxBounces=B, mInitialM=4)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy() if B > 1 else np.zeros((h, w, 1, 3), np.float32)
    op.execute_stage(5, 0, color)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(10)
    vol = (res["runningSum"] > 0) & (res["depth"] < 1e37)
    tag = (res["sampledPixel"] >> 16) & 0xF
    depth_of = res["sampledPixel"] >> 20
    groups = [(vol & (depth_of == 0) & (res["lightID"] == -3), 5), (vol & (tag == 0) & (res["lightID"] != -3), 5)]
    if B > 1:
        groups.append((vol & (tag == 1), 5))
    picks = []
    for mask, n in groups:
        ys, xs = np.nonzero(mask)
        assert len(ys) >= n, len(ys)
        picks += [(int(xs[i]), int(ys[i])) for i in rng.permutation(len(ys))[:n]]
    for x, y in picks:
        got = res[y, x]
        want, want_extra = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, mips) if B > 1 else (sw.initial_sampling_pixel(frame, x, y, frame_count, mips), None)
        assert int(got["sampledPixel"]) == want["sampledPixel"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        k = int(got["sampledPixel"]) >> 20
        if k:
            np.testing.assert_allclose(extra[y, x, :k], want_extra[:k], rtol=1e-5, atol=3e-6)
        if (int(got["sampledPixel"]) >> 16) & 0xF == 1:          # the stored vertex position (z as bits in lightID)
            np.testing.assert_allclose(sw.decode_emissive_position(int(got["lightID"]), got["lightUV"]),
                                       sw.decode_emissive_position(want["lightID"], want["lightUV"]), rtol=1e-5, atol=1e-5)
        else:
            assert int(got["lightID"]) == want["lightID"]
            np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=3e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=2e-4, abs=1e-12), (x, y)
        np.testing.assert_allclose(color[y, x, :3], sw.final_shading_path(frame, x, y, got, extra[y, x]), rtol=3e-4, atol=1e-9)


@pytest.mark.parametrize("B,emission", [(1, False), (4, False), (3, True)])
def test_reference_path_tracer_matches_the_slang_witness(B, emission):
    """mUseReference (the ground truth of the unbiasedness tests): paths by decomposition tracking, next-event estimation with
    residual-ratio-tracked shadow rays at every vertex, emission along the way, Russian roulette — same random-number stream, so the
    same paths: the radiance of single pixels agrees to rounding."""
    from volumetricrestirrelease_b200 import Scene
    w, h = 40, 30
    if emission:
        sc = Scene()
        sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=0.2, dataFile="plume", numMips=4, densityScale=0.1, hasVelocity=True,
                         hasEmission=True, LeScale=0.3, temperatureCutoff=1.0, temperatureScale=100.0, dim=(64, 96, 64), seed=3, voxelSize=1.0)
        sc.setEnvMap((128, 64), seed=7); sc.setEnvMapIntensity(0.5); sc.frame_camera(1.0)
    else:
        sc = env_scene(dim=(64, 64, 56), density_scale=0.1, env_size=(128, 64), g=0.4)
    lo, hi = sc.volume_bounds_world()
    sc.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
    params = VolumetricReSTIRParams(mUseReference=1, mMaxBounces=B, mBaselineSamplePerPixel=2, mUseAnalyticLights=1)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    frame_count = op.frame_count()
    color = op.execute()
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    mask_pass = vro.OraclePass(VolumetricReSTIRParams())          # only for K0's mask of the pixels whose ray meets the medium
    mask_pass.setScene(sc, w, h)
    mask_pass.execute_stage(0, 0, np.zeros((h, w, 4), np.float32))
    inside = mask_pass.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(h, w)["transmittance"] < 0.9
    rng = np.random.default_rng(15)
    picks = []
    for mask, n in ((inside, 12), (~inside, 2)):
        ys, xs = np.nonzero(mask)
        picks += [(int(xs[k]), int(ys[k])) for k in rng.permutation(len(ys))[:n]]
    for x, y in picks:
        want = sw.path_trace_pixel(frame, x, y, frame_count, mips)
        np.testing.assert_allclose(color[y, x, :3], want, rtol=3e-4, atol=1e-7, err_msg=str((x, y)))
        assert color[y, x, 3] == 1.0


@pytest.mark.parametrize("B,S", [(4, 2), (3, 1)])
def test_vertex_reuse_matches_the_slang_witness(B, S):
    """VERTEX_REUSE: from bounce S on the extra-bounce records hold world-space vertices, the path density switches to area measure
    at vertex S, and the p-hat evaluation after K1 leaves the suffix past that vertex in p_partial (the light transmittance when the
    path ends at vertex S, luminance(F / prefix) when it goes on); the final shading re-evaluates the whole path."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=0.3)
    params = VolumetricReSTIRParams(mEnableSpatialReuse=0, mMaxBounces=B, mVertexReuse=1, mVertexReuseStartBounce=S, mInitialM=3)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy()
    pp = op.get_buffer(capi.BUF_PPARTIAL_0).view(np.float32).reshape(h, w).copy()
    op.execute_stage(5, 0, color)
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    mips, off, dim = [], 0, 512
    while dim >= 1:
        mips.append(imp[off:off + dim * dim].reshape(dim, dim).copy()); off += dim * dim; dim //= 2
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(17)
    depth_of = res["sampledPixel"] >> 20
    picks = []
    for k in range(B):
        ys, xs = np.nonzero((res["runningSum"] > 0) & (depth_of == k) & (res["depth"] < 1e37))
        assert len(ys) >= 3, (k, len(ys))
        picks += [(int(xs[i]), int(ys[i])) for i in rng.permutation(len(ys))[:4]]
    for x, y in picks:
        got = res[y, x]
        want, want_extra = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, mips)
        assert int(got["sampledPixel"]) == want["sampledPixel"] and int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        k = int(got["sampledPixel"]) >> 20
        np.testing.assert_allclose(extra[y, x, :k], want_extra[:k], rtol=1e-5, atol=1e-5)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        if k >= S:
            assert float(pp[y, x]) == pytest.approx(float(want["p_partial"]), rel=2e-4, abs=1e-12), (x, y, k)
        np.testing.assert_allclose(color[y, x, :3], sw.final_shading_path(frame, x, y, dict(want), extra[y, x]), rtol=3e-4, atol=1e-9)


def test_configuration_1_matches_the_slang_witness():
    """BASELINE's configuration 1 as the reference runs it (64^3 sphere, one directional light, no env map, initial RIS only — so
    gNoReuse: decomposition-tracked candidates, albedo weights): K1 reservoirs and the frame's radiance."""
    from common import config1_params, config1_scene
    w, h = 48, 48
    sc = config1_scene()
    params = config1_params(M=4)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    op.execute_stage(5, 0, color)
    frame = sw.Frame(sc, params, w, h)
    frame.lights = sw.Lights(sc)
    rng = np.random.default_rng(19)
    ys, xs = np.nonzero((res["runningSum"] > 0) & (res["depth"] < 1e37))
    assert len(ys) > 100
    lit = 0
    for k in rng.permutation(len(ys))[:14]:
        x, y = int(xs[k]), int(ys[k])
        got = res[y, x]
        want, _ = sw.initial_sampling_pixel_paths(frame, x, y, frame_count, None)
        assert int(got["lightID"]) == want["lightID"] == 0 and float(got["M"]) == float(want["M"]) == 4.0, (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        rad = sw.final_shading_path(frame, x, y, got, np.zeros((1, 3), np.float32))
        np.testing.assert_allclose(color[y, x, :3], rad, rtol=2e-4, atol=1e-9)
        lit += bool(rad.sum() > 0)
    assert lit >= 8
