"""p-hat, the target function every reuse stage evaluates (VR/ReSTIRHelper.slang:91-200,426-496 for one bounce and an env-map light):
luminance(camera transmittance x density x sigma_s x env radiance x phase x light transmittance).  Composed here from the
independent witnesses (oracle/march_witness.py: ray-marched transmittance + point query; oracle/light_witness.py: env evaluation,
phase function) with the option -> mip / sampler mapping of VR/VolumetricReSTIR.cpp:455-500, and compared with the C++ oracle's
evaluate_P_hat AND with the p_y that the oracle's K1 stored for the reservoirs of a rendered frame."""
import numpy as np
import pytest

from common import RES, env_scene
from oracle import light_witness as lw
from oracle import vro
from oracle.march_witness import Witness
from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi

F = np.float32
K_RAY_TMAX = F(3.402823466e+38)


def _witness_p_hat(sc, params, grid, w, h, px, py, depth, uv, light_id, cache):
    cam = sc.camera.data(w, h)
    o = np.array(cam.posW[:], dtype=F)
    U, V, Wv = (np.array(getattr(cam, k)[:], dtype=F) for k in ("cameraU", "cameraV", "cameraW"))
    p = np.array([(F(px) + F(0.5)) / F(w), (F(py) + F(0.5)) / F(h)], dtype=F)        # F/Scene/Camera/Camera.slang:160-228
    ndc = np.array([F(2) * p[0] + F(-1), F(-2) * p[1] + F(1)], dtype=F)
    d = ndc[0] * U + ndc[1] * V + Wv
    d = (d / np.sqrt(np.dot(d, d))).astype(F)

    def wit(mip):
        if mip not in cache:
            cache[mip] = Witness(grid, mip)
        return cache[mip]
    vis_mip, vis_lin, vis_tss = params.mSpatialVisibilityMipLevel, bool(params.mSpatialVisibilityUseLinearSampler), params.mSpatialVisibilityTStepScale
    lig_mip, lig_lin, lig_tss = params.mSpatialLightingMipLevel, bool(params.mSpatialLightingUseLinearSampler), params.mSpatialLightingTStepScale
    vol = grid.volume
    if depth == K_RAY_TMAX:                                           # background sample: transmittance of the whole ray x env radiance
        vis = F(wit(vis_mip).ray_marching(o, d, float(K_RAY_TMAX), vis_lin, vis_tss))
        Fv = vis * lw.env_eval(sc.envMap, d, sc.envMapIntensity)
        return float(lw.luminance(Fv))
    pw = (o + d * F(depth)).astype(F)
    density = wit(0).density_world(pw)
    if density == 0:
        return 0.0
    vis = F(wit(vis_mip).ray_marching(o, d, float(depth), vis_lin, vis_tss))
    sigma_s = np.array(vol.sigma_s[:], dtype=F)
    Fv = (vis * density * sigma_s).astype(F)
    # evaluate_L_in_volume, env light: direction from the stored (x, y) and the hemisphere bit of the light id
    z = np.sqrt(F(1) - uv[0] * uv[0] - uv[1] * uv[1]).astype(F) if F(1) - uv[0] * uv[0] - uv[1] * uv[1] >= 0 else F(0)
    wi = np.array([uv[0], uv[1], -z if light_id == -2 else z], dtype=F)
    Ld = lw.env_eval(sc.envMap, wi, sc.envMapIntensity) * F(lw.phase_hg(float(np.dot(-d, wi)), vol.PhaseFunctionConstantG))
    tr = F(wit(lig_mip).ray_marching(pw, wi, float(K_RAY_TMAX), lig_lin, lig_tss))
    return float(lw.luminance((Fv * (tr * Ld)).astype(F)))


@pytest.mark.parametrize("g,kw", [(0.0, {}), (0.6, dict(mSpatialVisibilityMipLevel=2, mSpatialLightingTStepScale=1.5))])
def test_p_hat_composed_from_the_witnesses_matches_the_oracle(g, kw):
    w, h = 48, 36
    sc = env_scene(dim=(64, 64, 56), density_scale=0.06, g=g, env_size=(128, 64))
    params = VolumetricReSTIRParams(**kw)
    op = vro.OraclePass(params)
    op.setScene(sc, w, h)
    color = np.zeros((h, w, 4), np.float32)
    op.execute_stage(0, 0, color); op.execute_stage(1, 0, color)       # K0 + K1: reservoirs whose p_y is K1's final p-hat
    res = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w)
    grid = sc.volume.grid.contents
    cache, checked, volume_samples, background_samples = {}, 0, 0, 0
    ys, xs = np.nonzero(res["runningSum"] > 0)
    order = np.random.default_rng(1).permutation(len(ys))
    for k in order:
        y, x = int(ys[k]), int(xs[k])
        r = res[y, x]
        is_bg = r["depth"] > 1e37
        if not is_bg and (r["lightID"] >= 0 or r["lightID"] == -3):     # env-light samples only (light id -1 / -2)
            continue
        if is_bg and background_samples >= 4 or not is_bg and volume_samples >= 16:
            continue
        want = _witness_p_hat(sc, params, grid, w, h, x, y, r["depth"], r["lightUV"], int(r["lightID"]), cache)
        got = op.p_hat(x, y, float(r["depth"]), r["lightUV"], int(r["lightID"]))
        assert got == pytest.approx(want, rel=2e-5, abs=1e-12), (x, y, got, want)
        assert float(r["p_y"]) == pytest.approx(want, rel=2e-5, abs=1e-12), (x, y, float(r["p_y"]), want)   # what K1 stored
        checked += 1; volume_samples += not is_bg; background_samples += is_bg
        if volume_samples >= 16 and background_samples >= 4:
            break
    assert volume_samples >= 12 and background_samples >= 2, (volume_samples, background_samples)
