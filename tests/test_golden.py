"""Committed golden fixtures (tests/golden/, generated from the oracle by tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest

from common import (FEAT, capi, compare_reservoirs, config1_params, config1_scene, env_scene, gpu_frame, make_pair,
                    rel_err_image, vro)
from volumetricrestirrelease_b200 import VolumetricReSTIRParams

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_oracle_reproduces_config1_golden():
    g = np.load(os.path.join(G, "config1_64.npz"))
    op = vro.OraclePass(config1_params(4))
    op.setScene(config1_scene(), 64, 64)
    img = op.execute()
    flips, err = compare_reservoirs(op.get_buffer(capi.BUF_RESERVOIR_0), g["reservoirs"])
    assert flips.mean() <= 1e-3 and err <= 1e-4          # bit-exact on the generating host; libm may differ elsewhere
    e = rel_err_image(img, g["image"], mask=~flips.reshape(64, 64))
    assert (e.max() if e.size else 0) <= 1e-4
    np.testing.assert_allclose(op.get_buffer(capi.BUF_FEATURES).view(FEAT)["transmittance"], g["features"].view(FEAT)["transmittance"], rtol=2e-5)


def test_oracle_reproduces_env_reuse_golden():
    g = np.load(os.path.join(G, "env_reuse_64x48.npz"))
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15, env_size=(128, 64))
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 64, 48)
    op.execute()
    img = op.execute()
    np.testing.assert_allclose(op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)[:4096], g["importance_head"], rtol=1e-5)
    e = rel_err_image(img, g["image"])
    assert (e > 1e-4).mean() <= 5e-3


@pytest.mark.gpu
def test_gpu_matches_config1_golden():
    g = np.load(os.path.join(G, "config1_64.npz"))
    gp, _ = make_pair(config1_scene(), config1_params(4), 64, 64)
    img = gpu_frame(gp, 64, 64)
    flips, err = compare_reservoirs(gp.get_buffer(capi.BUF_RESERVOIR_0), g["reservoirs"])
    e = rel_err_image(img, g["image"], mask=~flips.reshape(64, 64))
    print(f"[golden config1] flips {int(flips.sum())}/{flips.size} rel err {err:.3g} radiance {float(e.max()) if e.size else 0:.3g}")
    assert flips.mean() <= 1e-3 and err <= 1e-4 and (e.max() if e.size else 0) <= 1e-4
