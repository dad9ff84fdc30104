"""The C++ oracle against oracle/light_witness.py — a numpy restatement, written from the reference's shader sources, of the
env-map importance map, the hierarchical env sampler, env evaluation and the Henyey-Greenstein phase function / sampler."""
import numpy as np
import pytest

from oracle import light_witness as lw
from oracle import vro
from volumetricrestirrelease_b200 import Scene, VolumetricReSTIRParams, capi


def _scene(env_size=(256, 128)):
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile="sphere", numMips=3, densityScale=0.05, dim=(32, 32, 32), seed=1, voxelSize=1.0)
    sc.setEnvMap(env_size, seed=7)
    sc.setEnvMapIntensity(1.5)
    sc.frame_camera(1.0)
    return sc


@pytest.fixture(scope="module")
def oracle_and_mips():
    sc = _scene()
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16)          # the oracle builds its own importance map
    return sc, op, lw.importance_mips(sc.envMap)


def test_importance_map_matches(oracle_and_mips):
    sc, op, mips = oracle_and_mips
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    off = 0
    for m in mips:
        got = imp[off:off + m.size].reshape(m.shape)
        np.testing.assert_allclose(got, m, rtol=2e-5, atol=1e-7)
        off += m.size
    assert off == imp.size and mips[-1].shape == (1, 1)


def test_hierarchical_env_sampling_matches(oracle_and_mips):
    sc, op, mips = oracle_and_mips
    # the oracle's own map feeds the witness sampler here, so that a 1-ulp difference of a texel cannot flip a 2x2 decision
    imp = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    own, off = [], 0
    for m in mips:
        own.append(imp[off:off + m.size].reshape(m.shape).copy()); off += m.size
    rng = np.random.default_rng(5)
    seen = set()
    for u0, u1 in rng.random((400, 2)).astype(np.float32):
        d_ref, pdf_ref, le_ref = op.env_sample(float(u0), float(u1))
        d, pdf, pos = lw.env_sample(own, u0, u1)
        seen.add(pos)
        np.testing.assert_allclose(d_ref, d, rtol=0, atol=3e-6)
        assert pdf_ref == pytest.approx(pdf, rel=1e-6)
        np.testing.assert_allclose(le_ref, lw.env_eval(sc.envMap, d_ref, sc.envMapIntensity), rtol=2e-5, atol=1e-6)
    assert len(seen) > 300                       # the samples really spread over the map
    # importance sampling: the pdf is proportional to the chosen texel
    pdfs = np.array([lw.env_sample(own, a, b)[1] for a, b in rng.random((200, 2)).astype(np.float32)])
    assert pdfs.min() > 0 and np.isfinite(pdfs).all()


def test_env_eval_matches(oracle_and_mips):
    sc, op, _ = oracle_and_mips
    rng = np.random.default_rng(9)
    for v in rng.normal(size=(200, 3)).astype(np.float32):
        v = v / np.linalg.norm(v)
        np.testing.assert_allclose(op.env_eval(v), lw.env_eval(sc.envMap, v, sc.envMapIntensity), rtol=2e-5, atol=1e-6)


@pytest.mark.parametrize("g", [0.0, 0.6, -0.3, 0.0005, 0.95])
def test_henyey_greenstein_matches(g):
    rng = np.random.default_rng(3)
    for _ in range(200):
        wo = rng.normal(size=3).astype(np.float32); wo /= np.linalg.norm(wo)
        u0, u1 = rng.random(2).astype(np.float32)
        wi_ref, pdf_ref = vro.sample_phase(g, wo, float(u0), float(u1))
        wi, pdf = lw.sample_phase(g, wo, u0, u1)
        np.testing.assert_allclose(wi_ref, wi, rtol=0, atol=2e-6)
        assert pdf_ref == pytest.approx(pdf, rel=2e-6)
        c = float(rng.uniform(-1, 1))
        assert vro.phase_hg(c, g) == pytest.approx(lw.phase_hg(c, g), rel=2e-6)
    # the sampler's pdf integrates to one over the sphere (quadrature over cos theta)
    c = np.linspace(-1, 1, 200001)
    denom = 1 + g * g + 2 * g * c
    assert np.trapezoid((1 - g * g) / (denom * np.sqrt(denom)) / (4 * np.pi) * 2 * np.pi, c) == pytest.approx(1.0, rel=1e-4)
