"""Shared scene/config factories for the tests (SURVEY.md section 8d configurations at test sizes)."""
import os

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


def make_pair(scene, params, w, h, dict_=None, own_tables=False):
    """(product pass on cuda:0, oracle pass) sharing the scene.  By default the oracle receives the GPU-built importance map
    (checked against the oracle's own in test_env_importance_map) and the product's alias tables (checked against independent
    numpy restatements of the reference builders in tests/test_alias_tables.py); own_tables=True hands the oracle the emissive
    alias table rebuilt by that numpy restatement of F/Utils/Sampling/AliasTable.cpp instead of the product's."""
    d = dict(dict_ or {})
    gp = VolumetricReSTIR.create(dict({"mParams": params}, **d))
    gp.setScene(scene, w, h)
    imp = None
    env_alias = None
    if scene.envMap is not None:
        imp = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
        env_alias = gp.env_alias()   # product extension (no reference counterpart); its distribution is checked in test_alias_tables.py
    em = None
    if scene.emissiveTriangles is not None:
        em = gp.emissive_alias(len(scene.emissiveTriangles))
        if own_tables:
            from oracle import alias_oracle
            em = alias_oracle.emissive_alias_table(scene.emissiveTriangles)
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


def compare_reservoirs(a, b, rel=1e-4, pp=None):
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
    if pp is not None:     # (gpu, oracle) p_partial planes (vertex reuse): compared where the reservoir holds a sample
        fa, fb = pp[0].astype(np.float64).reshape(a.shape), pp[1].astype(np.float64).reshape(a.shape)
        e = np.abs(fa - fb) / np.maximum(np.abs(fb), 1e-30)
        e[(fa == fb) | (np.isnan(fa) & np.isnan(fb)) | ~(b["runningSum"] > 0)] = 0
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


FLIP_BUDGET = float(os.environ.get("VRESTIR_FLIP_BUDGET", "1e-3"))   # north star: flips <= 0.1 % of pixels (test_gpu_fast_build.py loosens it for the contraction build)
RADIANCE_RTOL = 1e-4      # north star: radiance within 1e-4 relative per pixel (non-flipped)
# K0 transmittance, exact build: libm-ulp level; the contraction-enabled build is held to the north-star 1e-4 (test_gpu_fast_build.py)
FEATURE_RTOL = float(os.environ.get("VRESTIR_FEATURE_RTOL", "2e-5"))


def staged(params, scene, w, h, frames=2, want_mvec=False, dict_=None, camera_path=None, own_tables=False):
    """Run `frames` frames stage by stage; after every stage the GPU state is overwritten with the oracle's, so each kernel
    is compared on identical inputs.  Returns per-stage (flips, err) of the last frame + the final images.
    camera_path: optional list of camera positions, one per frame (K2 reprojection with a moving camera).
    own_tables: the oracle builds its own alias tables / importance map instead of receiving the product's."""
    import torch
    d = dict(dict_ or {})
    if want_mvec:
        d["mOutputMotionVec"] = 1
    gp, op = make_pair(scene, params, w, h, d or None, own_tables=own_tables)
    out = {}
    color_g = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    mvec_g = torch.zeros((h, w, 2), dtype=torch.float32, device="cuda")
    color_c = np.zeros((h, w, 4), np.float32)
    mvec_c = np.zeros((h, w, 2), np.float32)
    B = params.mMaxBounces
    rounds = params.mSpatialReuseRounds if params.mEnableSpatialReuse else 0
    vr = bool(params.mVertexReuse) and B > 1
    twin = {capi.BUF_RESERVOIR_0: capi.BUF_PPARTIAL_0, capi.BUF_RESERVOIR_1: capi.BUF_PPARTIAL_1, capi.BUF_RESERVOIR_TEMPORAL: capi.BUF_PPARTIAL_TEMPORAL}

    def sync(buf_ids):
        for b in buf_ids:
            gp.set_buffer(b, op.get_buffer(b))
            if vr and b in twin:
                gp.set_buffer(twin[b], op.get_buffer(twin[b]))

    def pp(bid):
        return (gp.get_buffer(twin[bid]).view(np.float32), op.get_buffer(twin[bid]).view(np.float32)) if vr else None

    for f in range(frames):
        last = f == frames - 1
        if camera_path is not None:
            if len(camera_path[f]) == 2:      # (position, target): a pan shifts the whole image
                scene.camera.position, scene.camera.target = tuple(camera_path[f][0]), tuple(camera_path[f][1])
            else:
                scene.camera.position = tuple(camera_path[f])
            gp.updateCamera(); op.updateCamera()
        for stage, arg in [(0, 0), (1, 0), (2, 0)] + [(3, r) for r in range(rounds)] + [(4, 0), (5, 0), (6, 0)]:
            gp.execute_stage(stage, arg, color_g.data_ptr(), mvec_g.data_ptr())
            op.execute_stage(stage, arg, color_c, mvec_c)
            torch.cuda.synchronize()
            if stage == 0:
                fg = gp.get_buffer(capi.BUF_FEATURES).view(FEAT)
                fc = op.get_buffer(capi.BUF_FEATURES).view(FEAT)
                np.testing.assert_allclose(fg["transmittance"], fc["transmittance"], rtol=FEATURE_RTOL, atol=1e-7)
                sync([capi.BUF_FEATURES])
            elif stage in (1, 2):
                bid = capi.BUF_RESERVOIR_0
                flips, err = compare_reservoirs(gp.get_buffer(bid), op.get_buffer(bid), pp=pp(bid))
                if last:
                    out["initial" if stage == 1 else "temporal"] = (flips, err)
                sync([bid] + ([capi.BUF_EXTRA_0] if B > 1 else []))
            elif stage == 3:
                bid = capi.BUF_RESERVOIR_1 if arg % 2 == 0 else capi.BUF_RESERVOIR_0
                flips, err = compare_reservoirs(gp.get_buffer(bid), op.get_buffer(bid), pp=pp(bid))
                if last:
                    out[f"spatial{arg}"] = (flips, err)
                sync([bid] + ([capi.BUF_EXTRA_1 if arg % 2 == 0 else capi.BUF_EXTRA_0] if B > 1 else []))
            elif stage == 4:
                if params.mEnableTemporalReuse:
                    sync([capi.BUF_RESERVOIR_TEMPORAL] + ([capi.BUF_EXTRA_TEMPORAL] if B > 1 else []))
            elif stage == 5 and last:
                out["final"] = (color_g.cpu().numpy(), color_c.copy())
                out["mvec"] = (mvec_g.cpu().numpy(), mvec_c.copy())
    out["launches"] = gp.launch_count()
    return out


def check_staged(out, w, h, name, budget=FLIP_BUDGET):
    budget = max(budget, FLIP_BUDGET)
    for k, v in out.items():
        if k in ("final", "mvec", "launches"):
            continue
        flips, err = v
        print(f"[{name}:{k}] flips {int(flips.sum())}/{flips.size} ({flips.mean():.2e}) rel err {err:.3g}")
        assert flips.mean() <= budget, k
        assert err <= RADIANCE_RTOL, k
    g, c = out["final"]
    e = rel_err_image(g, c)
    bad = (e > RADIANCE_RTOL).mean()
    print(f"[{name}:final] radiance rel err max {float(e.max()):.3g}, frac > 1e-4: {bad:.2e}")
    assert bad <= budget
    assert (c[..., :3].sum(-1) > 0).mean() > 0.02, "the test image is (almost) black: nothing was compared"
