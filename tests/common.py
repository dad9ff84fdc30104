"""Shared scene/config factories for the tests (SURVEY.md section 8d configurations at test sizes)."""
import numpy as np

from oracle import vro
from volumetricrestirrelease_b200 import Scene, VolumetricReSTIR, VolumetricReSTIRParams, capi

RES = vro.RES_DTYPE
FEAT = vro.FEAT_DTYPE


def config1_scene(dim=64, num_mips=3, density_scale=0.03):
    """64^3 sphere x fBm, sigma_a=1, sigma_s=9, one directional light (config 1)."""
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile="sphere", numMips=num_mips,
                     densityScale=density_scale, dim=(dim, dim, dim), seed=1, voxelSize=1.0)
    sc.addDirectionalLight((-1, -1, -0.5), (5, 5, 5))
    sc.frame_camera(1.1)
    return sc


def config1_params(M=4, **kw):
    return VolumetricReSTIRParams(mEnableTemporalReuse=0, mEnableSpatialReuse=0, mUseEnvironmentLights=0,
                                  mUseAnalyticLights=1, mInitialM=M, **kw)


def env_scene(kind="bunny", dim=(96, 96, 80), num_mips=4, density_scale=0.08, env_size=(256, 128), g=0.0, seed=2,
              distance=1.0, **kw):
    """Small env-lit scene for reuse-stage parity (config 2 at test size)."""
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=g, dataFile=kind, numMips=num_mips,
                     densityScale=density_scale, dim=dim, seed=seed, voxelSize=1.0, **kw)
    sc.setEnvMap(env_size, seed=7)
    sc.setEnvMapIntensity(1.5)
    sc.frame_camera(distance)
    return sc


def make_pair(scene, params, w, h, dict_=None):
    """(product pass on cuda:0, oracle pass) sharing the scene, the GPU-built importance map and alias tables."""
    d = dict(dict_ or {})
    gp = VolumetricReSTIR.create(dict({"mParams": params}, **d))
    gp.setScene(scene, w, h)
    imp = None
    env_alias = None
    if scene.envMap is not None:
        imp = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
        env_alias = gp.env_alias()
    em = None
    if scene.emissiveTriangles is not None:
        em = gp.emissive_alias(len(scene.emissiveTriangles))
    op = vro.OraclePass(params)
    op.setScene(scene, w, h, importance=imp, emissive_alias=em, env_alias=env_alias)
    if d:
        op.updateDict(d)
    return gp, op


def gpu_frame(gp, w, h, want_mvec=False):
    import torch
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    mvec = torch.zeros((h, w, 2), dtype=torch.float32, device="cuda") if want_mvec else None
    gp.execute(color.data_ptr(), mvec.data_ptr() if want_mvec else None)
    torch.cuda.synchronize()
    return (color.cpu().numpy(), mvec.cpu().numpy()) if want_mvec else color.cpu().numpy()


def compare_reservoirs(a, b, rel=1e-4):
    """Returns (flip mask, max relative error of the float fields on non-flipped pixels).

    North-star protocol: integer bookkeeping must be bit-exact wherever no candidate selection flipped.  A pixel counts
    as a flip when an integer field (lightID, sampledPixel, M) or the identity of the selected sample (depth, lightUV)
    differs, or when a weight (runningSum, p_y) differs by more than `rel` (a flipped selection *inside* the candidate
    generation, e.g. between bounce reservoirs, shows up only there).  Flips are counted against the budget by the caller."""
    a = a.view(RES)
    b = b.view(RES)
    flips = (a["lightID"] != b["lightID"]) | (a["sampledPixel"] != b["sampledPixel"]) | (a["M"] != b["M"])
    fa, fb = a["depth"], b["depth"]
    flips |= ~np.isclose(fa, fb, rtol=1e-5, atol=0) & ~(fa == fb)
    flips |= (~np.isclose(a["lightUV"], b["lightUV"], rtol=1e-4, atol=1e-6)).any(axis=-1)
    errs = np.zeros(a.shape, dtype=np.float64)
    for f in ("runningSum", "p_y"):
        fa, fb = a[f].astype(np.float64), b[f].astype(np.float64)
        e = np.abs(fa - fb) / np.maximum(np.abs(fb), 1e-30)
        e[fa == fb] = 0
        e[~np.isfinite(e)] = np.inf
        errs = np.maximum(errs, e)
    flips |= errs > rel
    ok = ~flips
    return flips, float(errs[ok].max()) if ok.any() else 0.0


def rel_err_image(a, b, mask=None, floor=1e-6):
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    e = np.abs(a - b) / np.maximum(np.abs(b), floor)
    e[(a == b)] = 0
    e = e.max(axis=-1)
    if mask is not None:
        e = e[mask]
    return e


def rel_mse(a, b):
    a = a[..., :3].astype(np.float64)
    b = b[..., :3].astype(np.float64)
    eps = 1e-2 * np.mean(b) ** 2
    return float(np.mean((a - b) ** 2 / (b ** 2 + eps)))
