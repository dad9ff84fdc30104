"""CPU: edge cases of the oracle (empty and ragged inputs, extreme option values).  The GPU suite repeats the
ragged frame (tests/test_gpu_parity.py::test_ragged_frame_size_staged)."""
import numpy as np
import pytest

from common import capi, env_scene
from oracle import vro
from volumetricrestirrelease_b200 import Scene, VolumetricReSTIRParams

kRayTMax = np.float32(3.4028234663852886e38)


def _run(params, scene, w, h, frames=2):
    op = vro.OraclePass(params)
    op.setScene(scene, w, h)
    img = None
    for _ in range(frames):
        img = op.execute()
    return op, img


def test_empty_volume_is_pure_background():
    """No density anywhere: every reservoir ends as a background sample (depth = FLT_MAX), the image is the env map seen
    through transmittance 1, and nothing is NaN."""
    sc = Scene()
    sc.addGVDBVolume(dense=np.zeros((24, 24, 24), np.float32), numMips=3, densityScale=1.0)
    sc.setEnvMap((64, 32), seed=3)
    sc.frame_camera(1.5)
    op, img = _run(VolumetricReSTIRParams(), sc, 24, 16)
    assert np.isfinite(img).all() and (img[..., :3].sum(-1) > 0).all()
    res = op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(np.float32).reshape(-1, 8)
    assert np.all(res[:, 2] == kRayTMax)
    feat = op.get_buffer(capi.BUF_FEATURES_TEMPORAL).view(np.dtype([("n", np.int32), ("t", np.float32)]))
    assert np.all(feat["t"] == 1.0)


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (33, 9)])
def test_tiny_and_ragged_frames(size):
    w, h = size
    op, img = _run(VolumetricReSTIRParams(), env_scene(), w, h)
    assert img.shape == (h, w, 4) and np.isfinite(img).all()


@pytest.mark.parametrize("kw", [dict(mInitialM=1), dict(mSpatialSampleCount=32, mSampleRadius=30.0), dict(mSpatialReuseRounds=3),
                                dict(mEnableTemporalReuse=0), dict(mEnableSpatialReuse=0), dict(mTemporalReuseMThreshold=1.0)])
def test_extreme_option_values_run_and_stay_finite(kw):
    op, img = _run(VolumetricReSTIRParams(**kw), env_scene(), 40, 24, frames=3)
    assert np.isfinite(img).all() and (img[..., :3].sum(-1) > 0).mean() > 0.5
    res = op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(np.float32).reshape(-1, 8)
    assert np.isfinite(res[:, 0]).all() and (res[:, 1] >= 0).all()      # running sums finite, M non-negative
