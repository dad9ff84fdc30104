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


@pytest.mark.parametrize("kw", [dict(), dict(mInitialM=6, mInitialLightingMipLevel=1, mInitialBaseMipLevel=2)])
def test_initial_sampling_matches_the_slang_witness(kw):
    w, h = 40, 30
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute()                                                # frame 0, so that the frame counter is not 0
    frame_count = op.frame_count()
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
    checked = volume = 0
    for x, y in picks:
        got = res[y, x]
        want = sw.initial_sampling_pixel(frame, x, y, frame_count, mips)
        assert int(got["lightID"]) == want["lightID"] and float(got["M"]) == float(want["M"]), (x, y, got, want)
        assert float(got["depth"]) == pytest.approx(float(want["depth"]), rel=3e-6), (x, y)
        np.testing.assert_allclose(np.asarray(got["lightUV"], np.float32), want["lightUV"], rtol=0, atol=3e-6)
        assert float(got["runningSum"]) == pytest.approx(float(want["runningSum"]), rel=1e-4, abs=1e-12), (x, y)
        assert float(got["p_y"]) == pytest.approx(float(want["p_y"]), rel=5e-5, abs=1e-12), (x, y)
        checked += 1; volume += float(got["depth"]) < 1e37
    assert checked == 12 and volume >= 6


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
