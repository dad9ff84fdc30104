"""The oracle's spatial-reuse pass (K3: neighbour offsets, resampling at the centre ray, Talbot MIS over all taps, weighted reservoir
streaming with the pixel's random-number stream) against oracle/stage_witness.py, a Python restatement written from
VR/SpatialReuse.cs.slang on top of the independent transmittance / light / RNG witnesses."""
import numpy as np
import pytest

from common import FEAT, RES, env_scene
from oracle import stage_witness as sw
from oracle import vro
from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi


@pytest.mark.parametrize("kw", [dict(), dict(mSpatialMISMethod=capi.kMISNone, mSampleRadius=6.0, mSpatialSampleCount=3),
                                dict(mRandomSamplerType=capi.kHammersley, mSpatialSampleCount=5, mSampleRadius=7.0),
                                dict(mSpatialVisibilityTrackingMethod=capi.kResidualRatioTracking, mSpatialLightingTrackingMethod=capi.kRatioTracking, mSpatialSampleCount=3),
                                dict(mSpatialVisibilityTrackingMethod=capi.kAnalyticTracking, mSpatialLightingTrackingMethod=capi.kAnalogResidualRatioTracking,
                                     mSpatialVisibilityUseLinearSampler=0, mSpatialLightingMipLevel=2)])
def test_spatial_reuse_matches_the_slang_witness(kw):
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()                                                 # frame 0 (fills the history)
    frame_count = op.frame_count()
    for stage in (0, 1, 2):
        op.execute_stage(stage, 0, color)
    res_in = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    feat = op.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(h, w).copy()
    op.execute_stage(3, 0, color)
    res_out = op.get_buffer(capi.BUF_RESERVOIR_1).view(RES).reshape(h, w)
    frame = sw.Frame(sc, params, w, h)
    assert frame_count >= 1
    rng = np.random.default_rng(11)
    ys, xs = np.nonzero(feat["transmittance"] != 1.0)
    # With a random-walk tracker in p-hat the number of draws a pixel consumes depends on the last bit of powf (an estimate that is
    # exactly 0 in one implementation and 1e-40 in the other adds or skips a reservoir draw and shifts every later draw of the
    # pixel): those cases bound the number of diverged pixels instead of demanding every one.
    stochastic = any(kw.get(k, capi.kRayMarching) in (capi.kRatioTracking, capi.kResidualRatioTracking, capi.kAnalogResidualRatioTracking)
                     for k in ("mSpatialVisibilityTrackingMethod", "mSpatialLightingTrackingMethod"))
    checked = changed = diverged = 0
    for k in rng.permutation(len(ys))[:12 if stochastic else 9]:
        x, y = int(xs[k]), int(ys[k])
        got = res_out[y, x]
        want = sw.spatial_reuse_pixel(frame, res_in, feat, x, y, frame_count)
        same = (int(got["lightID"]) == want["lightID"] and int(got["sampledPixel"]) == want["sampledPixel"] and float(got["M"]) == float(want["M"])
                and float(got["depth"]) == float(want["depth"]) and np.array_equal(np.asarray(got["lightUV"], np.float32), want["lightUV"])
                and float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=5e-5, abs=1e-12)
                and float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=5e-5, abs=1e-12))
        if stochastic and not same:
            diverged += 1
            continue
        assert same, (x, y, got, want)
        checked += 1
        changed += float(got["depth"]) != float(res_in[y, x]["depth"])
    assert checked >= 9 and diverged <= 1 and changed >= 2   # some pixels ended up with a neighbour's sample
    # a pixel whose ray misses the medium passes through
    bys, bxs = np.nonzero(feat["transmittance"] == 1.0)
    if len(bys):
        x, y = int(bxs[0]), int(bys[0])
        assert res_out[y, x] == res_in[y, x]


def test_final_shading_matches_the_slang_witness():
    """K5: F of the stored sample under the final options (exact transmittance of the trilinear mip-0 interpolant for the camera and
    the light ray) times the RIS weight runningSum / (p_y M), against the radiance the oracle wrote."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, env_size=(128, 64))
    params = VolumetricReSTIRParams()
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    color = np.zeros((h, w, 4), np.float32)
    for stage in (0, 1, 2, 3, 4):
        op.execute_stage(stage, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(RES).reshape(h, w).copy()      # the frame's final reservoirs (history after K4)
    op.execute_stage(5, 0, color)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(2)
    ys, xs = np.nonzero((res["runningSum"] > 0) & (res["depth"] < 1e37))
    lit = 0
    for k in rng.permutation(len(ys))[:10]:
        x, y = int(xs[k]), int(ys[k])
        want = frame.final_shading(x, y, res[y, x])
        np.testing.assert_allclose(color[y, x, :3], want, rtol=5e-5, atol=1e-9)
        lit += bool(want.sum() > 0)
    assert lit >= 6
    by, bx = np.nonzero((res["runningSum"] > 0) & (res["depth"] > 1e37))
    for k in range(min(3, len(by))):
        np.testing.assert_allclose(color[by[k], bx[k], :3], frame.final_shading(int(bx[k]), int(by[k]), res[by[k], bx[k]]), rtol=5e-5, atol=1e-9)


@pytest.mark.parametrize("kw,move", [(dict(), False), (dict(mTemporalMISMethod=capi.kMISNone, mTemporalReuseMThreshold=2.0), False), (dict(), True),
                                     (dict(mTemporalReprojectionMode=capi.kReprojectionNone), True),
                                     (dict(mTemporalReprojectionMode=capi.kReprojectionNoBackground), True),
                                     (dict(mSpatialVisibilityTrackingMethod=capi.kRatioTracking, mSpatialLightingTrackingMethod=capi.kResidualRatioTracking), True)])
def test_temporal_reuse_matches_the_slang_witness(kw, move):
    """K2 on a frame with history: reprojection of the stored depth (or of a density-sampled point for a background sample) through
    the previous frame's view-projection, resampling of the history sample on the current ray, Talbot MIS between the two samples,
    M-capped reservoir streaming — static camera and a camera that moved between the frames."""
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.updateDict({"mOutputMotionVec": 1})
    op.execute()
    cam0 = sc.camera.data(w, h)
    prev_cam = tuple(np.array(getattr(cam0, k)[:], dtype=np.float32) for k in ("posW", "cameraU", "cameraV", "cameraW", "viewMat", "projMat"))
    if move:
        pos = np.array(sc.camera.position); sc.camera.position = tuple(pos + np.array([6.0, -4.0, 2.0])); sc.camera.target = tuple(np.array(sc.camera.target) + np.array([6.0, -4.0, 2.0]))   # a pan: every pixel shifts
        op.updateCamera()
    frame_count = op.frame_count()
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res_cur = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    res_prev = op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(RES).reshape(h, w).copy()
    feat_cur = op.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(h, w).copy()
    feat_prev = op.get_buffer(capi.BUF_FEATURES_TEMPORAL).view(FEAT).reshape(h, w).copy()
    mvec = np.zeros((h, w, 2), np.float32)
    op.execute_stage(2, 0, color, mvec)
    res_out = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(4)
    vol = feat_cur["transmittance"] != 1.0
    escaped = res_cur["depth"] > 1e30                 # K1 kept a background sample: K2 reprojects a density-sampled point of the ray
    picks = []
    for mask, n in ((vol & ~escaped, 8), (vol & escaped, 4)):
        ys, xs = np.nonzero(mask)
        assert len(ys) >= n
        picks += [(int(xs[k]), int(ys[k])) for k in rng.permutation(len(ys))[:n]]
    from_history = shifted = diverged = 0
    stochastic = "mSpatialVisibilityTrackingMethod" in kw
    for x, y in picks:
        got = res_out[y, x]
        info = {}
        want = sw.temporal_reuse_pixel(frame, res_cur, res_prev, feat_cur, feat_prev, x, y, frame_count, prev_cam, info=info)
        assert tuple(mvec[y, x]) == tuple(float(v) for v in info["mvec"]), (x, y)          # the motion vector output
        shifted += bool(np.any(mvec[y, x] != 0))
        same = (int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"])
                and float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6)
                and np.allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=1e-7)
                and float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=1e-4, abs=1e-12)
                and float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=5e-5, abs=1e-12))
        if stochastic and not same:                         # a random walk whose draw count hangs on the last bit of powf (see the spatial test)
            diverged += 1
            continue
        assert same, (x, y, got, want)
        from_history += float(got["M"]) > float(res_cur[y, x]["M"])
    assert from_history >= 8 - diverged and diverged <= 1            # the history really took part
    assert shifted >= 4 - diverged if (move and kw.get("mTemporalReprojectionMode", 0) != capi.kReprojectionNone) else shifted == 0


def test_spatial_reuse_of_multi_bounce_paths_matches_the_slang_witness():
    """K3 with MAX_BOUNCES = 3: the targets of the taps are whole paths (vertex loop of evaluate_F_ on the tap's extra-bounce
    records, re-evaluated from the centre pixel's and every other tap's camera ray), and the selected sample's records travel with it."""
    w, h, B = 40, 30, 3
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=0.3)
    params = VolumetricReSTIRParams(mMaxBounces=B, mSpatialSampleCount=3)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()
    frame_count = op.frame_count()
    for stage in (0, 1, 2):
        op.execute_stage(stage, 0, color)
    res_in = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra_in = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy()
    feat = op.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(h, w).copy()
    op.execute_stage(3, 0, color)
    res_out = op.get_buffer(capi.BUF_RESERVOIR_1).view(RES).reshape(h, w)
    extra_out = op.get_buffer(capi.BUF_EXTRA_1).view(np.float32).reshape(h, w, B - 1, 3)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(12)
    ys, xs = np.nonzero((feat["transmittance"] != 1.0) & ((res_out["sampledPixel"] >> 20) > 0))
    assert len(ys) >= 8
    changed = 0
    for k in rng.permutation(len(ys))[:8]:
        x, y = int(xs[k]), int(ys[k])
        got = res_out[y, x]
        want, want_extra = sw.spatial_reuse_pixel(frame, res_in, feat, x, y, frame_count, extra_in=extra_in)
        assert int(got["lightID"]) == want["lightID"] and int(got["sampledPixel"]) == want["sampledPixel"], (x, y)
        assert float(got["M"]) == float(want["M"]) and float(got["depth"]) == float(want["depth"]), (x, y)
        assert np.array_equal(np.asarray(got["lightUV"], np.float32), want["lightUV"]), (x, y)
        n = int(got["sampledPixel"]) >> 20
        assert np.array_equal(extra_out[y, x, :n], want_extra[:n]), (x, y)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=1e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        changed += float(got["depth"]) != float(res_in[y, x]["depth"])
    assert changed >= 2


def test_temporal_reuse_of_multi_bounce_paths_matches_the_slang_witness():
    """K2 with MAX_BOUNCES = 3 and a camera that moved: the history sample's path is re-evaluated from the current camera ray, the
    current sample's path from the previous frame's ray (primary depth converted between the two rays), and the records of the
    selected path end up in the current frame's extra-bounce buffer."""
    w, h, B = 40, 30, 3
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=0.3)
    params = VolumetricReSTIRParams(mMaxBounces=B)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    cam0 = sc.camera.data(w, h)
    prev_cam = tuple(np.array(getattr(cam0, k)[:], dtype=np.float32) for k in ("posW", "cameraU", "cameraV", "cameraW", "viewMat", "projMat"))
    pos = np.array(sc.camera.position); sc.camera.position = tuple(pos + np.array([0.2, -0.1, 0.1]))
    op.updateCamera()
    frame_count = op.frame_count()
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    get = lambda b, t: op.get_buffer(b).view(t)
    res_cur, res_prev = get(capi.BUF_RESERVOIR_0, RES).reshape(h, w).copy(), get(capi.BUF_RESERVOIR_TEMPORAL, RES).reshape(h, w).copy()
    extra_cur = get(capi.BUF_EXTRA_0, np.float32).reshape(h, w, B - 1, 3).copy()
    extra_prev = get(capi.BUF_EXTRA_TEMPORAL, np.float32).reshape(h, w, B - 1, 3).copy()
    feat_cur, feat_prev = get(capi.BUF_FEATURES, FEAT).reshape(h, w).copy(), get(capi.BUF_FEATURES_TEMPORAL, FEAT).reshape(h, w).copy()
    op.execute_stage(2, 0, color)
    res_out = get(capi.BUF_RESERVOIR_0, RES).reshape(h, w)
    extra_out = get(capi.BUF_EXTRA_0, np.float32).reshape(h, w, B - 1, 3)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(13)
    ys, xs = np.nonzero((feat_cur["transmittance"] != 1.0) & ((res_out["sampledPixel"] >> 20) > 0))
    assert len(ys) >= 10
    from_history = 0
    for k in rng.permutation(len(ys))[:10]:
        x, y = int(xs[k]), int(ys[k])
        got = res_out[y, x]
        want, want_extra = sw.temporal_reuse_pixel(frame, res_cur, res_prev, feat_cur, feat_prev, x, y, frame_count, prev_cam, extra_cur, extra_prev)
        assert int(got["lightID"]) == want["lightID"] and int(got["sampledPixel"]) == want["sampledPixel"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        n = int(got["sampledPixel"]) >> 20
        assert np.array_equal(extra_out[y, x, :n], want_extra[:n]), (x, y)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        from_history += not np.array_equal(extra_out[y, x, :n], extra_cur[y, x, :n])
    assert from_history >= 2


@pytest.mark.parametrize("use_prev", [1, 0])
def test_temporal_reuse_on_an_animated_volume_matches_the_slang_witness(use_prev):
    """BASELINE's configuration 3 in small: the volume advances between the frames (density, temperature and velocity grids), so K2
    moves the reprojection point back along the velocity field and — with mUsePrevVolumeForReproj — evaluates the current sample on
    the previous frame's ray in the previous frame's grids (density slots 19.., temperature slot 27)."""
    from volumetricrestirrelease_b200 import Scene

    def plume(t):
        sc = Scene()
        sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=0.0, dataFile="plume", numMips=4, densityScale=0.1, hasVelocity=True,
                         hasEmission=True, LeScale=0.3, temperatureCutoff=1.0, temperatureScale=100.0, dim=(64, 96, 64), seed=3, voxelSize=1.0,
                         frameTime=t)
        sc.setEnvMap((128, 64), seed=7)
        sc.setEnvMapIntensity(0.5)
        sc.frame_camera(1.0)
        return sc
    w, h = 40, 30
    sc, nxt = plume(0.0), plume(0.6)
    params = VolumetricReSTIRParams(mUsePrevVolumeForReproj=use_prev)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    cam0 = sc.camera.data(w, h)
    prev_cam = tuple(np.array(getattr(cam0, k)[:], dtype=np.float32) for k in ("posW", "cameraU", "cameraV", "cameraW", "viewMat", "projMat"))
    prev_volume = sc.volume                                     # advanceVolume rebinds sc.volume
    op.advanceVolume(nxt.volume)
    pos = np.array(sc.camera.position); sc.camera.position = tuple(pos + np.array([0.6, -0.4, 0.3]))
    op.updateCamera()
    frame_count = op.frame_count()
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    get = lambda b, t: op.get_buffer(b).view(t).reshape(h, w).copy()
    res_cur, res_prev = get(capi.BUF_RESERVOIR_0, RES), get(capi.BUF_RESERVOIR_TEMPORAL, RES)
    feat_cur, feat_prev = get(capi.BUF_FEATURES, FEAT), get(capi.BUF_FEATURES_TEMPORAL, FEAT)
    op.execute_stage(2, 0, color)
    res_out = get(capi.BUF_RESERVOIR_0, RES)
    frame = sw.Frame(sc, params, w, h)
    frame.grid, frame.prev_grid = nxt.volume.grid.contents, prev_volume.grid.contents
    rng = np.random.default_rng(14)
    vol = feat_cur["transmittance"] != 1.0
    picks = []
    for mask, n in ((vol & (res_out["lightID"] == -3), 4), (vol & (res_out["lightID"] != -3) & (res_out["depth"] < 1e37), 6), (vol & (res_cur["depth"] > 1e37), 2)):
        ys, xs = np.nonzero(mask)
        assert len(ys) >= n, len(ys)
        picks += [(int(xs[k]), int(ys[k])) for k in rng.permutation(len(ys))[:n]]
    from_history = 0
    for x, y in picks:
        got = res_out[y, x]
        want = sw.temporal_reuse_pixel(frame, res_cur, res_prev, feat_cur, feat_prev, x, y, frame_count, prev_cam)
        assert int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=2e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=1e-4, abs=1e-12), (x, y)
        from_history += float(got["M"]) > float(res_cur[y, x]["M"])
    assert from_history >= 8


@pytest.mark.parametrize("kw", [
    dict(mFinalVisibilityTrackingMethod=capi.kResidualRatioTracking, mFinalLightTrackingMethod=capi.kRatioTracking, mFinalVisibilitySamples=2, mFinalLightSamples=3),
    dict(mFinalVisibilityTrackingMethod=capi.kRayMarching, mFinalLightTrackingMethod=capi.kAnalogResidualRatioTracking, mFinalLightSamples=2, mMaxBounces=3)])
def test_final_shading_with_stochastic_trackers_matches_the_slang_witness(kw):
    """K5 with random-walk transmittance estimators: the camera segment, (with several bounces) every path segment and the light
    segment are estimated in that order from the pixel's last-round generator, n samples averaged each."""
    w, h = 40, 30
    B = kw.get("mMaxBounces", 1)
    sc = env_scene(dim=(64, 64, 56), density_scale=0.08, env_size=(128, 64), g=0.2)
    params = VolumetricReSTIRParams(mEnableSpatialReuse=0, **kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    frame_count = op.frame_count()
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w).copy()
    extra = op.get_buffer(capi.BUF_EXTRA_0).view(np.float32).reshape(h, w, B - 1, 3).copy() if B > 1 else np.zeros((h, w, 1, 3), np.float32)
    op.execute_stage(5, 0, color)
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(16)
    ys, xs = np.nonzero((res["runningSum"] > 0) & (res["depth"] < 1e37))
    lit = 0
    for k in rng.permutation(len(ys))[:12]:
        x, y = int(xs[k]), int(ys[k])
        want = sw.final_shading_path(frame, x, y, res[y, x], extra[y, x], frame_count)
        np.testing.assert_allclose(color[y, x, :3], want, rtol=2e-4, atol=1e-9, err_msg=str((x, y)))
        lit += bool(want.sum() > 0)
    assert lit >= 8


def test_reuse_with_vertex_reuse_matches_the_slang_witness():
    """K2 and K3 with VERTEX_REUSE (B = 4, S = 2): temporal reuse re-evaluates whole paths and rewrites p_partial of the history
    sample; spatial reuse evaluates only the prefix up to the reuse vertex on the new camera ray and multiplies by the stored
    p_partial; the selected sample's p_partial travels with it."""
    w, h, B, S = 40, 30, 4, 2
    sc = env_scene(dim=(64, 64, 56), density_scale=0.12, env_size=(128, 64), g=0.3)
    params = VolumetricReSTIRParams(mMaxBounces=B, mVertexReuse=1, mVertexReuseStartBounce=S, mSpatialSampleCount=3)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    op.execute()
    cam0 = sc.camera.data(w, h)
    prev_cam = tuple(np.array(getattr(cam0, k)[:], dtype=np.float32) for k in ("posW", "cameraU", "cameraV", "cameraW", "viewMat", "projMat"))
    pos = np.array(sc.camera.position); sc.camera.position = tuple(pos + np.array([0.2, -0.1, 0.1]))
    op.updateCamera()
    frame_count = op.frame_count()
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)
    res = lambda b: op.get_buffer(b).view(RES).reshape(h, w).copy()
    ext = lambda b: op.get_buffer(b).view(np.float32).reshape(h, w, B - 1, 3).copy()
    ppl = lambda b: op.get_buffer(b).view(np.float32).reshape(h, w).copy()
    feat = lambda b: op.get_buffer(b).view(FEAT).reshape(h, w).copy()
    cur = (res(capi.BUF_RESERVOIR_0), ext(capi.BUF_EXTRA_0), ppl(capi.BUF_PPARTIAL_0))
    prev = (res(capi.BUF_RESERVOIR_TEMPORAL), ext(capi.BUF_EXTRA_TEMPORAL), ppl(capi.BUF_PPARTIAL_TEMPORAL))
    feat_cur, feat_prev = feat(capi.BUF_FEATURES), feat(capi.BUF_FEATURES_TEMPORAL)
    op.execute_stage(2, 0, color)
    k2 = (res(capi.BUF_RESERVOIR_0), ext(capi.BUF_EXTRA_0), ppl(capi.BUF_PPARTIAL_0))
    op.execute_stage(3, 0, color)
    k3 = (res(capi.BUF_RESERVOIR_1), ext(capi.BUF_EXTRA_1), ppl(capi.BUF_PPARTIAL_1))
    frame = sw.Frame(sc, params, w, h)
    rng = np.random.default_rng(18)

    def check(got_planes, x, y, want, want_extra, tag):
        got = got_planes[0][y, x]
        assert int(got["lightID"]) == want["lightID"] and int(got["sampledPixel"]) == want["sampledPixel"] and float(got["M"]) == float(want["M"]), (tag, x, y)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (tag, x, y)
        n = int(got["sampledPixel"]) >> 20
        assert np.array_equal(got_planes[1][y, x, :n], want_extra[:n]), (tag, x, y)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=3e-4, abs=1e-12), (tag, x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=2e-4, abs=1e-12), (tag, x, y)
        if n >= S:
            assert float(got_planes[2][y, x]) == pytest.approx(float(want["p_partial"]), rel=3e-4, abs=1e-12), (tag, x, y)
        return n

    for tag, out in (("temporal", k2), ("spatial", k3)):
        ys, xs = np.nonzero((feat_cur["transmittance"] != 1.0) & ((out[0]["sampledPixel"] >> 20) >= S))
        assert len(ys) >= 8, tag
        for k in rng.permutation(len(ys))[:8]:
            x, y = int(xs[k]), int(ys[k])
            if tag == "temporal":
                want, we = sw.temporal_reuse_pixel(frame, cur[0], prev[0], feat_cur, feat_prev, x, y, frame_count, prev_cam, cur[1], prev[1], cur[2], prev[2])
            else:
                want, we = sw.spatial_reuse_pixel(frame, k2[0], feat_cur, x, y, frame_count, extra_in=k2[1], pp_in=k2[2])
            assert check(out, x, y, want, we, tag) >= S
