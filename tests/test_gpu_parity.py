"""GPU parity proper: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs."""
import numpy as np
import pytest

from common import (FEAT, FEATURE_RTOL, FLIP_BUDGET, RADIANCE_RTOL, RES, capi, compare_reservoirs, config1_params, config1_scene, env_scene, gpu_frame,
                    make_pair, rel_err_image)
from volumetricrestirrelease_b200 import VolumetricReSTIR, VolumetricReSTIRParams

pytestmark = pytest.mark.gpu





def _report(name, flips, err, e):
    print(f"[{name}] flips {int(flips.sum())}/{flips.size} ({flips.mean():.2e}), reservoir rel err {err:.3g}, "
          f"radiance rel err max {float(e.max()) if e.size else 0:.3g}")


@pytest.mark.parametrize("M", [4, 1])
def test_config1_single_frame_no_reuse(M):
    """64^3 grid, 256^2, one directional light, initial RIS only, frame 0."""
    w = h = 256
    gp, op = make_pair(config1_scene(), config1_params(M), w, h)
    img_gpu = gpu_frame(gp, w, h)
    img_cpu = op.execute()
    fg = gp.get_buffer(capi.BUF_FEATURES).view(FEAT)
    fc = op.get_buffer(capi.BUF_FEATURES).view(FEAT)
    assert (fg["noReflectiveSurface"] == fc["noReflectiveSurface"]).all()
    np.testing.assert_allclose(fg["transmittance"], fc["transmittance"], rtol=FEATURE_RTOL, atol=1e-7)
    flips, err = compare_reservoirs(gp.get_buffer(capi.BUF_RESERVOIR_0), op.get_buffer(capi.BUF_RESERVOIR_0))
    e = rel_err_image(img_gpu, img_cpu, mask=~flips.reshape(h, w))
    _report(f"config1 M={M}", flips, err, e)
    assert flips.mean() <= FLIP_BUDGET
    assert err <= RADIANCE_RTOL
    assert (e.max() if e.size else 0) <= RADIANCE_RTOL
    assert (img_cpu[..., :3].sum(-1) > 0).mean() > 0.1      # the test actually lights pixels


def test_env_importance_map():
    """K6: importance map (512^2 + mips) vs the oracle's own computation."""
    from oracle import vro
    sc = env_scene()
    gp, _ = make_pair(sc, VolumetricReSTIRParams(), 32, 32)
    g = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 32, 32)      # computes its own
    c = op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
    np.testing.assert_allclose(g, c, rtol=2e-5, atol=1e-7)


from common import check_staged as _check_staged, staged as _staged  # noqa: E402


def test_full_reuse_staged_env():
    """Config 2 at test size: env light, temporal + spatial reuse, stage-by-stage parity on the second frame."""
    w, h = 160, 96
    out = _staged(VolumetricReSTIRParams(), env_scene(), w, h, frames=2, want_mvec=True)
    _check_staged(out, w, h, "env full reuse")
    g, c = out["mvec"]
    assert (np.abs(g - c) > 1e-6).mean() <= FLIP_BUDGET


def test_full_reuse_staged_moving_camera():
    """Temporal reprojection with a camera that moves between the two frames (K2 reprojection + depth conversion)."""
    import torch
    w, h = 128, 96
    sc = env_scene()
    params = VolumetricReSTIRParams()
    gp, op = make_pair(sc, params, w, h)
    color_g = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    color_c = np.zeros((h, w, 4), np.float32)
    gp.execute(color_g.data_ptr()); color_c = op.execute()
    pos = np.array(sc.camera.position)
    sc.camera.position = tuple(pos + np.array([0.7, 0.2, -0.3]) * 6.0)
    gp.updateCamera(); op.updateCamera()
    for b in (capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_FEATURES_TEMPORAL):
        gp.set_buffer(b, op.get_buffer(b))
    for stage in (0, 1):
        gp.execute_stage(stage, 0, color_g.data_ptr()); op.execute_stage(stage, 0, color_c)
    gp.set_buffer(capi.BUF_RESERVOIR_0, op.get_buffer(capi.BUF_RESERVOIR_0))
    gp.set_buffer(capi.BUF_FEATURES, op.get_buffer(capi.BUF_FEATURES))
    gp.execute_stage(2, 0, color_g.data_ptr()); op.execute_stage(2, 0, color_c)
    flips, err = compare_reservoirs(gp.get_buffer(capi.BUF_RESERVOIR_0), op.get_buffer(capi.BUF_RESERVOIR_0))
    print(f"[moving camera temporal] flips {int(flips.sum())}/{flips.size} rel err {err:.3g}")
    assert flips.mean() <= FLIP_BUDGET and err <= RADIANCE_RTOL


@pytest.mark.parametrize("B", [2, 4])
def test_multibounce_staged(B):
    w, h = 96, 64
    p = VolumetricReSTIRParams(mMaxBounces=B)
    out = _staged(p, env_scene(dim=(64, 64, 56), density_scale=0.15), w, h, frames=2)
    # multi-bounce paths amplify libm ulp differences (sinf/cosf in Sample_p, the 1-(x^2+y^2) cancellation in decodeWiDist),
    # so a few more candidate selections flip than in the single-bounce configurations
    _check_staged(out, w, h, f"B={B}", budget=5e-3)


@pytest.mark.parametrize("B,S,scene", [(3, 1, "env"), (4, 2, "env"), (3, 1, "plume"), (3, 2, "plume")])
def test_vertex_reuse_staged(B, S, scene):
    """mVertexReuse (VERTEX_REUSE of the reference: paths are reconnected at world-space vertices from bounce S on, the spatial
    pass reuses the stored suffix p_partial instead of re-marching it), every stage against the oracle on identical inputs,
    including the p_partial plane.  The plume scene exercises the emissive scatter vertices on both sides of S."""
    w, h = 96, 64
    p = VolumetricReSTIRParams(mMaxBounces=B, mVertexReuse=1, mVertexReuseStartBounce=S)
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15) if scene == "env" else _plume_scene(0.0)
    out = _staged(p, sc, w, h, frames=2)
    _check_staged(out, w, h, f"vertex reuse B={B} S={S} {scene}", budget=5e-3)


def test_vertex_reuse_changes_the_frame_and_needs_two_bounces():
    """With one bounce the flag is inert (the reference cannot even compile that combination); with more it renders a different
    (equally unbiased: tests/test_oracle_properties.py) frame, and the p_partial buffers exist only then."""
    import torch
    w, h = 96, 64
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    imgs = {}
    for B, vr in ((1, 0), (1, 1), (3, 0), (3, 1)):
        gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams(mMaxBounces=B, mVertexReuse=vr)})
        gp.setScene(sc, w, h)
        for _ in range(2):
            imgs[B, vr] = gpu_frame(gp, w, h)
        if B > 1 and vr:
            pp = gp.get_buffer(capi.BUF_PPARTIAL_TEMPORAL).view(np.float32)
            assert pp.size == w * h and (pp != 0).mean() > 0.05
        else:
            with pytest.raises(capi.VRestirError):
                gp.get_buffer(capi.BUF_PPARTIAL_0)
    assert np.array_equal(imgs[1, 0].view(np.uint32), imgs[1, 1].view(np.uint32))
    assert not np.array_equal(imgs[3, 0], imgs[3, 1])
    assert abs(imgs[3, 1][..., :3].mean() / imgs[3, 0][..., :3].mean() - 1) < 0.1


def test_emissive_triangles_and_env():
    w, h = 96, 64
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    lo, hi = sc.volume_bounds_world()
    sc.addEmissiveShell(500, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
    p = VolumetricReSTIRParams(mUseEmissiveLights=1, mMaxBounces=2)
    out = _staged(p, sc, w, h, frames=2)
    _check_staged(out, w, h, "emissive")


@pytest.mark.parametrize("method", [capi.kRatioTracking, capi.kResidualRatioTracking, capi.kAnalogResidualRatioTracking,
                                    capi.kRayMarching])
def test_final_tracking_methods(method):
    w, h = 96, 64
    p = VolumetricReSTIRParams(mFinalVisibilityTrackingMethod=method, mFinalLightTrackingMethod=method,
                               mEnableTemporalReuse=0)
    out = _staged(p, env_scene(dim=(64, 64, 56), density_scale=0.15), w, h, frames=1)
    _check_staged(out, w, h, f"final method {method}")


def test_reference_path_tracer_mode():
    """mUseReference: the brute-force volumetric path tracer (the reference's own ground-truth mode)."""
    w, h = 96, 64
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    gp, op = make_pair(sc, VolumetricReSTIRParams(mUseReference=1, mMaxBounces=3), w, h)
    g = gpu_frame(gp, w, h)
    c = op.execute()
    e = rel_err_image(g, c)
    print(f"[reference mode] frac > 1e-4: {(e > 1e-4).mean():.2e}, max {e.max():.3g}")
    assert (e > RADIANCE_RTOL).mean() <= 5e-3   # delta/ratio tracking amplify libm ulps through many accept tests


def test_determinism_and_row_bands():
    """Same RNG keys regardless of launch geometry: two half-frame bands == one full frame, run twice == identical."""
    import torch
    w, h = 160, 96
    sc = env_scene()
    p = VolumetricReSTIRParams(mEnableTemporalReuse=0, mEnableSpatialReuse=0)
    gp, _ = make_pair(sc, p, w, h)
    a = gpu_frame(gp, w, h)
    gp.updateDict({})
    b = gpu_frame(gp, w, h)
    assert np.array_equal(a, b)
    from volumetricrestirrelease_b200 import VolumetricReSTIR
    halves = np.zeros_like(a)
    for r0, r1 in ((0, 40), (40, 96)):
        q = VolumetricReSTIR.create({"mParams": p})
        q.setScene(sc, w, h, r0, r1)
        color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        q.execute(color.data_ptr()); torch.cuda.synchronize()
        halves[r0:r1] = color.cpu().numpy()[r0:r1]
    assert np.array_equal(a, halves)


def test_three_level_tree_full_reuse():
    """Grid larger than 128^3 in every axis -> level-2 root (top_lev == 2) on all but the coarsest mips."""
    w, h = 160, 96
    sc = env_scene(dim=(200, 180, 150), density_scale=0.2, num_mips=4, distance=0.9)
    assert sc.volume.grid.contents.slots[0].top_lev == 2
    out = _staged(VolumetricReSTIRParams(), sc, w, h, frames=2)
    _check_staged(out, w, h, "three-level tree")


def _frames_buffers(params, scene, w, h, wavefront, frames=3, extra=None):
    """Run `frames` frames through the public execute() and return (image, final reservoirs) of the last one."""
    import torch
    d = {"mParams": params, "mUseWavefront": int(wavefront is not False)}
    d.update(extra or {})
    if wavefront == 0:
        d["mInitialMode"] = 0
    gp = VolumetricReSTIR.create(d)
    gp.setScene(scene, w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for _ in range(frames):
        gp.execute(color.data_ptr())
    torch.cuda.synchronize()
    return color.cpu().numpy(), gp.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).copy()


@pytest.mark.parametrize("variant", ["default", "no_mis", "three_level", "two_rounds", "analytic_light"])
def test_wavefront_equals_per_pixel(variant):
    """The task-stream (wavefront) form of the reuse stages must be BIT-identical to the per-pixel kernels: it reschedules
    the same arithmetic (shared multi-depth camera marches, compacted light marches), it does not approximate."""
    w, h = 192, 112
    kw, sc = {}, None
    if variant == "no_mis":
        kw = dict(mSpatialMISMethod=capi.kMISNone)
    elif variant == "three_level":
        sc = env_scene(dim=(200, 180, 150), density_scale=0.2, num_mips=4, distance=0.9)
    elif variant == "two_rounds":
        kw = dict(mSpatialReuseRounds=2, mSpatialSampleCount=3)
    elif variant == "analytic_light":
        sc = env_scene()
        sc.addPointLight((40.0, 160.0, 30.0), (4000.0, 3000.0, 2000.0))
        kw = dict(mUseAnalyticLights=1)
    sc = sc or env_scene()
    p = VolumetricReSTIRParams(**kw)
    img_s, res_s = _frames_buffers(p, sc, w, h, False)
    for mode in (True, 0):   # True: default wavefront pipeline (lock-step K1); 0: wavefront K2/K3/K5 with the per-pixel K1
        img_w, res_w = _frames_buffers(p, sc, w, h, mode)
        assert np.array_equal(res_w.view(np.uint32), res_s.view(np.uint32)), f"wavefront reservoirs (mode {mode}) differ from the per-pixel kernels"
        assert np.array_equal(img_w.view(np.uint32), img_s.view(np.uint32))
    assert (img_w[..., :3].sum(-1) > 0).mean() > 0.05


@pytest.mark.parametrize("variant", ["two_bounces", "four_bounces", "three_bounces_emissive_point_light", "four_bounces_no_mis_two_rounds",
                                     "final_ray_marching", "final_mixed_two_bounces", "spatial_analytic_two_bounces", "three_level_three_bounces",
                                     "three_bounces_vertex_reuse", "four_bounces_vertex_reuse_from_2", "three_bounces_trilinear_initial"])
def test_generic_task_streams_equal_per_pixel(variant):
    """The generic task-stream path (the stage bodies run as an emit pass and a consume pass around the march engine: multi-bounce
    option sets, ray-marched / mixed final shading, analytic spatial tracking) must be BIT-identical to the per-pixel kernels,
    also when the stages run in several row chunks (small scratch budget).  Result blocks start as NaN (mDebugPoisonResults), so a
    march the emit pass failed to foresee would surface in the image."""
    w, h = 160, 96
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    kw = {}
    if variant == "two_bounces":
        kw = dict(mMaxBounces=2)
    elif variant == "four_bounces":
        kw = dict(mMaxBounces=4)
    elif variant == "three_bounces_emissive_point_light":
        lo, hi = sc.volume_bounds_world()
        sc.addEmissiveShell(400, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
        sc.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
        kw = dict(mMaxBounces=3, mUseEmissiveLights=1, mUseAnalyticLights=1)
    elif variant == "four_bounces_no_mis_two_rounds":
        kw = dict(mMaxBounces=4, mSpatialMISMethod=capi.kMISNone, mTemporalMISMethod=capi.kMISNone, mSpatialReuseRounds=2, mSpatialSampleCount=3)
    elif variant == "final_ray_marching":
        kw = dict(mFinalVisibilityTrackingMethod=capi.kRayMarching, mFinalLightTrackingMethod=capi.kRayMarching)
    elif variant == "final_mixed_two_bounces":
        kw = dict(mMaxBounces=2, mFinalVisibilityTrackingMethod=capi.kRayMarching, mFinalLightTrackingMethod=capi.kAnalyticTracking, mFinalTStepScale=0.5)
    elif variant == "spatial_analytic_two_bounces":
        kw = dict(mMaxBounces=2, mSpatialLightingTrackingMethod=capi.kAnalyticTracking, mSpatialLightingMipLevel=2, mSpatialVisibilityTStepScale=1.5)
    elif variant == "three_level_three_bounces":
        sc = env_scene(dim=(200, 180, 150), density_scale=0.2, num_mips=4, distance=0.9)
        kw = dict(mMaxBounces=3)
    elif variant == "three_bounces_vertex_reuse":
        kw = dict(mMaxBounces=3, mVertexReuse=1, mVertexReuseStartBounce=1)
    elif variant == "four_bounces_vertex_reuse_from_2":
        kw = dict(mMaxBounces=4, mVertexReuse=1, mVertexReuseStartBounce=2)
    elif variant == "three_bounces_trilinear_initial":   # the indirect bounces' free-flight sampling runs in k_initial_mb_bounce_traverse, not the engine
        kw = dict(mMaxBounces=3, mInitialVisibilityUseLinearSampler=1)
    p = VolumetricReSTIRParams(**kw)
    img_s, res_s = _frames_buffers(p, sc, w, h, False)
    for budget_mb in (4096, 3):
        gp_extra = {"mDebugPoisonResults": 1, "mScratchBudgetMB": budget_mb}
        img_w, res_w = _frames_buffers(p, sc, w, h, True, extra=gp_extra)
        assert np.array_equal(res_w.view(np.uint32), res_s.view(np.uint32)), f"generic task-stream reservoirs differ from the per-pixel kernels (budget {budget_mb} MB)"
        assert np.array_equal(img_w.view(np.uint32), img_s.view(np.uint32)), f"generic task-stream image differs (budget {budget_mb} MB)"
    assert np.isfinite(img_w).all() and (img_w[..., :3].sum(-1) > 0).mean() > 0.05


def test_generic_task_streams_are_used_and_chunked():
    """The multi-bounce frame really runs through the emit / march / consume kernels (launch count), and a small scratch budget
    splits the stages into row chunks."""
    import torch
    w, h = 160, 96
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    counts = []
    for budget_mb in (4096, 3):
        gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams(mMaxBounces=3), "mScratchBudgetMB": budget_mb})
        gp.setScene(sc, w, h)
        gp.execute(color.data_ptr()); gp.execute(color.data_ptr())
        n0 = gp.launch_count()
        gp.execute(color.data_ptr())
        torch.cuda.synchronize()
        counts.append(gp.launch_count() - n0)
    # K0 + K1 {traverse, first step, M * (2B - 1) x (march, bounce traverse, step), p-hat emit / march / consume} + K2 {emit, 1 march, consume}
    # + K3 {emit, camera march, 1 march, consume} + K5 {emit, 1 march, consume}
    assert counts[0] == 1 + (2 + 3 * 4 * (2 * 3 - 1) + 3) + 3 + 4 + 3, counts
    assert counts[1] > counts[0], counts


def _plume_scene(frame_time, dim=(64, 96, 64)):
    """Config 3 at test size: plume-like column with a temperature grid (black-body emission) and a velocity grid."""
    from volumetricrestirrelease_b200 import Scene
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=0.0, dataFile="plume", numMips=4, densityScale=0.1,
                     hasVelocity=True, hasEmission=True, LeScale=0.01, temperatureCutoff=1.0, temperatureScale=100.0,
                     dim=dim, seed=3, voxelSize=1.0, frameTime=frame_time)
    sc.setEnvMap((256, 128), seed=7)
    sc.setEnvMapIntensity(0.5)
    sc.frame_camera(1.0)
    return sc


@pytest.mark.parametrize("wavefront", [1, 0])
def test_plume_animated_sequence_staged(wavefront):
    """emissive temperature grid + velocity-driven reprojection + previous-frame grids.  Three frames of
    an animated sequence (the volume advances every frame, the camera shakes), every stage compared with the oracle on
    identical inputs; run on the wavefront path and on the per-pixel kernels."""
    import torch
    w, h = 128, 96
    params = VolumetricReSTIRParams()
    sc = _plume_scene(0.0)
    gp, op = make_pair(sc, params, w, h, {"mOutputMotionVec": 1, "mUseWavefront": wavefront})
    assert sc.volume.grid.contents.volume.hasEmission and sc.volume.grid.contents.volume.hasVelocity
    color_g = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    mvec_g = torch.zeros((h, w, 2), dtype=torch.float32, device="cuda")
    color_c = np.zeros((h, w, 4), np.float32)
    mvec_c = np.zeros((h, w, 2), np.float32)
    keep = [sc.volume]
    pos0 = np.array(sc.camera.position)
    worst = {}
    self_emission = 0
    for f in range(3):
        if f > 0:
            nxt = _plume_scene(0.35 * f).volume
            keep.append(nxt)
            gp.advanceVolume(nxt); op.advanceVolume(nxt)
            sc.camera.position = tuple(pos0 + np.array([0.6, -0.4, 0.3]) * 2.0 * f)
            gp.updateCamera(); op.updateCamera()
        for stage, arg in [(0, 0), (1, 0), (2, 0), (3, 0), (4, 0), (5, 0), (6, 0)]:
            gp.execute_stage(stage, arg, color_g.data_ptr(), mvec_g.data_ptr())
            op.execute_stage(stage, arg, color_c, mvec_c)
            torch.cuda.synchronize()
            if stage == 0:
                gp.set_buffer(capi.BUF_FEATURES, op.get_buffer(capi.BUF_FEATURES))
            elif stage in (1, 2, 3):
                bid = capi.BUF_RESERVOIR_1 if stage == 3 else capi.BUF_RESERVOIR_0
                a, b = gp.get_buffer(bid), op.get_buffer(bid)
                flips, err = compare_reservoirs(a, b)
                key = ("initial", "temporal", "spatial")[stage - 1]
                worst[key] = max(worst.get(key, 0.0), float(flips.mean()))
                assert err <= RADIANCE_RTOL, (f, key, err)
                if stage == 1:
                    self_emission += int((b.view(RES)["lightID"] == -3).sum())
                gp.set_buffer(bid, b)
            elif stage == 4:
                gp.set_buffer(capi.BUF_RESERVOIR_TEMPORAL, op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL))
            elif stage == 5:
                e = rel_err_image(color_g.cpu().numpy(), color_c)
                worst["final"] = max(worst.get("final", 0.0), float((e > RADIANCE_RTOL).mean()))
                if f > 0:
                    assert (np.abs(mvec_g.cpu().numpy() - mvec_c) > 1e-6).mean() <= 5e-3
    print(f"[plume wavefront={wavefront}] worst flip / mismatch fractions {worst}, self-emission samples {self_emission}")
    assert self_emission > 0, "the emissive path was not exercised"
    for k, v in worst.items():
        assert v <= 5e-3, (k, v)   # libm-ulp differences in the black-body / velocity lookups flip a few more selections


def test_resident_volume_frames_match_streamed_advance():
    """Animated sequence with every frame resident on the device (vrestir_volume_frame_add + vrestir_advance_volume_resident,
    the reference's protocol: F/Scene/Scene.cpp:825-863) renders the same bits as uploading each frame's grids on advance;
    frames are revisited (cycled), bound twice in a row, and cleared while bound."""
    import torch
    w, h = 128, 96
    params = VolumetricReSTIRParams()
    vols = [_plume_scene(0.35 * f).volume for f in range(3)]
    imgs = []
    for resident in (False, True):
        sc = _plume_scene(0.0)
        gp = VolumetricReSTIR.create({"mParams": params, "mOutputMotionVec": 1})
        gp.setScene(sc, w, h)
        ids = [gp.addVolumeFrame(v) for v in vols] if resident else None
        color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        out = []
        for f, k in enumerate([1, 2, 0, 0, 1]):
            if resident:
                gp.advanceVolumeResident(ids[k])
            else:
                gp.advanceVolume(vols[k])
            gp.execute(color.data_ptr()); torch.cuda.synchronize()
            out.append(color.cpu().numpy().copy())
        imgs.append(out)
        if resident:
            gp.clearVolumeFrames()
            with pytest.raises(capi.VRestirError):
                gp.execute(color.data_ptr())          # the bound frame is gone: a volume has to be set again
            with pytest.raises(capi.VRestirError):
                gp.advanceVolumeResident(0)
            gp.setScene(sc, w, h)
            gp.execute(color.data_ptr()); torch.cuda.synchronize()
    for a, b in zip(*imgs):
        assert (a[..., :3].sum(-1) > 0).mean() > 0.02
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))


def test_accumulated_full_reuse_relmse():
    """North-star convergence check: frames accumulated with full temporal + spatial reuse on the GPU (default wavefront path,
    whole-frame execute) against the oracle's accumulation over the same frames, relMSE = mean((a-b)^2 / (b^2 + eps)) with
    eps = 1e-2 * mean(b)^2.  Both run the same RNG streams, so the accumulations differ only by the (counted)
    selection flips; the bound is the north star's 1e-3."""
    import torch
    from common import rel_mse
    w, h, frames = 96, 64, 48
    sc = env_scene()
    gp, op = make_pair(sc, VolumetricReSTIRParams(), w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    acc_g = np.zeros((h, w, 4), np.float64)
    acc_c = np.zeros((h, w, 4), np.float64)
    for _ in range(frames):
        gp.execute(color.data_ptr())
        torch.cuda.synchronize()
        acc_g += color.cpu().numpy()
        acc_c += op.execute()
    acc_g /= frames; acc_c /= frames
    r = rel_mse(acc_g, acc_c)
    print(f"[accumulated {frames} frames] relMSE gpu vs oracle {r:.3e}, mean gpu {acc_g[..., :3].mean():.6f} cpu {acc_c[..., :3].mean():.6f}")
    assert r <= 1e-3
    assert abs(acc_g[..., :3].mean() / acc_c[..., :3].mean() - 1) < 2e-3


def test_converged_4096_frames_relmse():
    """North star: "converged 4096-frame accumulations with full reuse must match within relMSE 1e-3".  4096 frames of the default
    pass (temporal + spatial reuse) accumulated by the product's AccumulatePass (double precision) against the oracle's frames
    accumulated in numpy, at reduced resolution so that the oracle finishes in seconds; plus the unbiasedness of the converged
    image against the reference's own ground-truth mode (mUseReference, 4096 spp): mean ratio within 1 %."""
    import torch
    from common import rel_mse
    from volumetricrestirrelease_b200.post import AccumulatePass
    w, h, frames = 64, 40, 4096
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15)
    gp, op = make_pair(sc, VolumetricReSTIRParams(), w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros_like(color)
    acc = AccumulatePass.create({"precisionMode": "Double"}, w, h)
    acc_c = np.zeros((h, w, 4), np.float64)
    for _ in range(frames):
        gp.execute(color.data_ptr())
        acc.execute(color.data_ptr(), out.data_ptr())
        acc_c += op.execute()
    torch.cuda.synchronize()
    acc_g = out.cpu().numpy().astype(np.float64)
    acc_c /= frames
    r = rel_mse(acc_g, acc_c)
    print(f"[converged {frames} frames] relMSE gpu vs oracle {r:.3e}, mean gpu {acc_g[..., :3].mean():.6f} cpu {acc_c[..., :3].mean():.6f}")
    assert r <= 1e-3
    assert abs(acc_g[..., :3].mean() / acc_c[..., :3].mean() - 1) < 1e-3
    ref = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams(mUseReference=1, mBaselineSamplePerPixel=64)})
    ref.setScene(sc, w, h)
    acc_r = np.zeros((h, w, 4), np.float64)
    for _ in range(64):
        ref.execute(color.data_ptr())
        torch.cuda.synchronize()
        acc_r += color.cpu().numpy()
    acc_r /= 64
    ratio = acc_g[..., :3].mean() / acc_r[..., :3].mean()
    print(f"[converged vs mUseReference 4096 spp] mean ratio {ratio:.4f}, relMSE {rel_mse(acc_g, acc_r):.3e}")
    assert abs(ratio - 1) < 0.01


def test_full_size_crop_with_history_1080p():
    """BASELINE.json configs[1] at full size, a frame WITH history (frame 2: K2 merges the previous frame's reservoirs, K3 reads
    merged neighbours): the GPU renders frames 0 and 1 through the public call; its history (temporal reservoirs, features,
    previous camera, frame counter) is handed to the oracle, and frame 2 is compared on two 64x64 crops (cloud centre and a
    silhouette region), every stage of the oracle run on the crop + the halo the later stages read."""
    import torch
    import bench
    from oracle import vro

    class A:
        pass
    args = A()
    args.width, args.height, args.dim, args.kind, args.mips, args.bounces = 1920, 1080, [577, 572, 438], "bunny", 4, 1
    sc = bench.build_scene(args)
    p = bench.make_params(args)
    w, h = args.width, args.height
    gp = VolumetricReSTIR.create({"mParams": p})
    gp.setScene(sc, w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    for _ in range(2):
        gp.execute(color.data_ptr())
    torch.cuda.synchronize()
    op = vro.OraclePass(p)
    op.setScene(sc, w, h, importance=gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32), env_alias=gp.env_alias())
    c0 = np.zeros((h, w, 4), np.float32)
    op.execute_stage(6, 0, c0)                      # saves the (static) camera as the previous frame's
    op.set_frame_count(2, 1)
    assert gp.frame_count() == 2
    history = {b: gp.get_buffer(b).copy() for b in (capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_FEATURES_TEMPORAL)}
    gp.execute(color.data_ptr()); torch.cuda.synchronize()
    g2 = color.cpu().numpy()
    gres = gp.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(RES).reshape(h, w)
    T, halo = 64, 10
    merged = 0
    for x0, y0 in ((w // 2 - 32, h // 2 - 32), (w // 2 - 420, h // 2 - 200)):
        for b, raw in history.items():      # the oracle's K4 / feature copies are whole-buffer: hand the history over per crop
            op.set_buffer(b, raw)
        for stage in (0, 1, 2):
            op.set_crop(x0 - halo, y0 - halo, x0 + T + halo, y0 + T + halo)
            op.execute_stage(stage, 0, c0)
        m = op.get_buffer(capi.BUF_RESERVOIR_0).view(RES).reshape(h, w)["M"][y0:y0 + T, x0:x0 + T]
        merged += int((m > 4).sum())              # K2 merged history: M = 4 candidates + min(history M, 4 * 4)
        for stage in (3, 4, 5):
            op.set_crop(x0, y0, x0 + T, y0 + T)
            op.execute_stage(stage, 0, c0)
        e = rel_err_image(g2[y0:y0 + T, x0:x0 + T], c0[y0:y0 + T, x0:x0 + T])
        bad = (e > RADIANCE_RTOL).mean()
        print(f"[1080p frame 2 crop at ({x0},{y0}) vs oracle] frac > 1e-4: {bad:.2e}, max {e.max():.3g}")
        assert bad <= 5e-3
        flips, err = compare_reservoirs(gres[y0:y0 + T, x0:x0 + T].copy(), op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).view(RES).reshape(h, w)[y0:y0 + T, x0:x0 + T].copy())
        print(f"[1080p frame 2 crop at ({x0},{y0})] final reservoirs: flips {int(flips.sum())}/{flips.size}, rel err {err:.3g}")
        assert flips.mean() <= 5e-3 and err <= RADIANCE_RTOL
        op.set_frame_count(2, 1)                  # stage 0 of the next crop must not restart the epoch
    assert merged > T * T // 2, "the crops did not contain merged history"


def test_full_size_properties_1080p():
    """BASELINE.json configs[1] at full size (577x572x438 bunny grid, 1920x1080, full reuse), through size-independent
    properties the oracle is too slow to check directly: (1) the wavefront path and the per-pixel kernels give the same
    bits for reservoirs and image after three frames, (2) a second run is bit-identical (no race in the task streams /
    atomics), (3) the image is finite, non-trivial and its mean agrees with the oracle on a 64x64 centre crop."""
    import torch
    import bench

    class A:
        pass
    args = A()
    args.width, args.height, args.dim, args.kind, args.mips, args.bounces = 1920, 1080, [577, 572, 438], "bunny", 4, 1
    sc = bench.build_scene(args)
    p = bench.make_params(args)
    w, h = args.width, args.height
    img_w, res_w = _frames_buffers(p, sc, w, h, True)
    img_w2, res_w2 = _frames_buffers(p, sc, w, h, True)
    img_s, res_s = _frames_buffers(p, sc, w, h, False)
    assert np.array_equal(res_w.view(np.uint32), res_w2.view(np.uint32)) and np.array_equal(img_w.view(np.uint32), img_w2.view(np.uint32))
    assert np.array_equal(res_w.view(np.uint32), res_s.view(np.uint32)), "wavefront reservoirs differ from the per-pixel kernels at 1080p"
    assert np.array_equal(img_w.view(np.uint32), img_s.view(np.uint32))
    assert np.isfinite(img_w).all() and (img_w[..., :3].sum(-1) > 0).mean() > 0.3
    # oracle on a centre crop of frame 0 (no history): same pixels, same RNG
    from oracle import vro
    gp = VolumetricReSTIR.create({"mParams": p})
    gp.setScene(sc, w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    gp.execute(color.data_ptr()); torch.cuda.synchronize()
    g0 = color.cpu().numpy()
    op = vro.OraclePass(p)
    op.setScene(sc, w, h, importance=gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32), env_alias=gp.env_alias())
    x0, y0, T, halo = w // 2 - 32, h // 2 - 32, 64, 10
    c0 = np.zeros((h, w, 4), np.float32)
    for stage in (0, 1, 2):
        op.set_crop(x0 - halo, y0 - halo, x0 + T + halo, y0 + T + halo)
        op.execute_stage(stage, 0, c0)
    for stage in (3, 4, 5):
        op.set_crop(x0, y0, x0 + T, y0 + T)
        op.execute_stage(stage, 0, c0)
    e = rel_err_image(g0[y0:y0 + T, x0:x0 + T], c0[y0:y0 + T, x0:x0 + T])
    bad = (e > RADIANCE_RTOL).mean()
    print(f"[1080p crop vs oracle] frac > 1e-4: {bad:.2e}, max {e.max():.3g}")
    assert bad <= 5e-3


@pytest.mark.parametrize("level", [1, 2])
@pytest.mark.parametrize("motion", ["static", "announced", "unannounced"])
def test_pipelined_frames_equal_serial(motion, level):
    """Frame pipelining ("mPipelineFrames": K0/K1 of frame f+1 run ahead on their own stream, next to K2..K5 of frame f) must not
    change a single bit of any frame — whether the prefetch is adopted (static camera, or a moving camera announced one frame
    ahead with setNextCamera) or discarded (camera moved without notice).  Level 2 also defers K5 to a third stream
    (the consumer orders itself with wait_output)."""
    import copy
    import torch
    w, h, frames = 192, 112, 5
    sc = env_scene()
    pos0 = np.array(sc.camera.position)
    path = [tuple(pos0 + np.array([0.6, -0.4, 0.3]) * 2.0 * f) if motion != "static" else tuple(pos0) for f in range(frames + 1)]

    def run(pipelined):
        gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams(), "mPipelineFrames": level if pipelined else 0})
        sc.camera.position = path[0]
        gp.setScene(sc, w, h)
        color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        imgs = []
        for f in range(frames):
            sc.camera.position = path[f]
            gp.updateCamera()
            if pipelined and motion == "announced":
                nxt = copy.copy(sc.camera)
                nxt.position = path[f + 1]
                gp.setNextCamera(nxt)
            gp.execute(color.data_ptr())
            gp.wait_output()                     # default stream waits for the (possibly deferred) final shading
            imgs.append(color.cpu().numpy().copy())   # .cpu() is ordered on the default stream
        torch.cuda.synchronize()
        return imgs, gp.get_buffer(capi.BUF_RESERVOIR_TEMPORAL).copy(), gp.pipeline_stats()

    ref_imgs, ref_res, st0 = run(False)
    imgs, res, st = run(True)
    sc.camera.position = tuple(pos0)
    assert st0["adopted"] == 0 and st0["discarded"] == 0
    if motion == "unannounced":
        assert st["discarded"] == frames - 1 and st["adopted"] == 0
    else:
        assert st["adopted"] == frames - 1 and st["discarded"] == 0
    for f in range(frames):
        assert np.array_equal(imgs[f].view(np.uint32), ref_imgs[f].view(np.uint32)), f"frame {f} differs with pipelining ({motion})"
    assert np.array_equal(res.view(np.uint32), ref_res.view(np.uint32))
    assert (imgs[-1][..., :3].sum(-1) > 0).mean() > 0.05


@pytest.mark.parametrize("mode", ["Double", "Single", "SingleCompensated"])
def test_accumulate_pass_matches_oracle(mode):
    """AccumulatePass on rendered frames: every running mean bit-identical to the numpy restatement; band
    calls, pass-through, subFrameCount and reset follow AccumulatePass.cpp:128-205."""
    import torch
    from oracle import post_oracle as po
    from volumetricrestirrelease_b200.post import AccumulatePass
    w, h, frames = 96, 64, 6
    sc = env_scene()
    gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams()})
    gp.setScene(sc, w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros_like(color)
    acc = AccumulatePass.create({"precisionMode": mode}, w, h)
    ref = po.Accumulator(mode)
    for f in range(frames):
        gp.execute(color.data_ptr())
        acc.execute(color.data_ptr(), out.data_ptr(), 0, 40)      # two band owners, one frame
        acc2_rows = (40, h)
        # the second band of the same frame must use the same frame counter: a second accumulator instance, as a second rank would own
        if f == 0:
            acc_b = AccumulatePass.create({"precisionMode": mode}, w, h)
        acc_b.execute(color.data_ptr(), out.data_ptr(), *acc2_rows)
        torch.cuda.synchronize()
        want = ref.add(color.cpu().numpy())
        assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32)), f"{mode}: frame {f}"
    assert acc.frameCount == frames
    # subFrameCount: accumulation stops after N frames, the output keeps the finished mean
    acc.reset(); acc.updateDict({"subFrameCount": 2})
    for f in range(4):
        gp.execute(color.data_ptr())
        acc.execute(color.data_ptr(), out.data_ptr())
        torch.cuda.synchronize()
        if f == 1:
            kept = out.cpu().numpy().copy()
    assert acc.frameCount == -1 and np.array_equal(out.cpu().numpy(), kept)
    # disabled: pass-through
    acc.updateDict({"enableAccumulation": False, "subFrameCount": 0}); acc.reset()
    acc.execute(color.data_ptr(), out.data_ptr()); torch.cuda.synchronize()
    assert torch.equal(out, color)


def test_error_measure_pass_matches_oracle():
    """ErrorMeasurePass: the per-pixel difference image bit-identical, the reduced numbers within 1e-6 relative (the sum order
    differs from numpy's; the reference's own float4 tree reduction has no defined order either)."""
    import torch
    from oracle import post_oracle as po
    from volumetricrestirrelease_b200.post import ErrorMeasurePass
    w, h = 200, 120
    rng = np.random.default_rng(5)
    src = rng.random((h, w, 4)).astype(np.float32) * 3
    ref = rng.random((h, w, 4)).astype(np.float32) * 3
    wp = np.ones((h, w, 4), np.float32); wp[rng.random((h, w)) < 0.3, 3] = 0
    ts, tr, tw = (torch.from_numpy(a).cuda() for a in (src, ref, wp))
    diff = torch.zeros_like(ts)
    for kw in ({}, {"ComputeSquaredDifference": False}, {"ComputeAverage": True}, {"IgnoreBackground": False}):
        em = ErrorMeasurePass(kw)
        m = em.execute(ts.data_ptr(), tr.data_ptr(), w, h, tw.data_ptr(), diff.data_ptr())
        d, err, avg = po.error_measure(src, ref, wp, em.IgnoreBackground, em.ComputeSquaredDifference, em.ComputeAverage)
        assert np.array_equal(diff.cpu().numpy()[..., :3].view(np.uint32), d.view(np.uint32)), kw
        assert np.allclose(m["error"], err, rtol=1e-6, atol=0) and np.isclose(m["avgError"], avg, rtol=1e-6), kw
        m2 = em.execute(ts.data_ptr(), tr.data_ptr(), w, h, tw.data_ptr())
        assert m2["error"] == m["error"]                      # deterministic reduction
        assert np.isclose(em.runningAvgError, avg, rtol=1e-6)    # EMA of two equal measurements
    m = ErrorMeasurePass().execute(ts.data_ptr(), tr.data_ptr(), w, h)   # unbound world position: no background test
    assert np.allclose(m["error"], po.error_measure(src, ref, None)[1], rtol=1e-6)


@pytest.mark.parametrize("op", ["Linear", "Reinhard", "ReinhardModified", "HejiHableAlu", "HableUc2", "Aces"])
def test_tone_mapper_matches_oracle(op):
    """ToneMapper (the display end of the reference's graphs) on a rendered frame: every operator, manual and
    auto exposure, white balance, clamp on / off, against the numpy restatement (fp32 kernel vs fp64 numpy: 2e-5 relative)."""
    import torch
    from oracle import post_oracle as po
    from volumetricrestirrelease_b200.post import ToneMapper
    w, h = 150, 91       # not a power of two: the luminance target is 128 x 64
    sc = env_scene()
    gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams()})
    gp.setScene(sc, w, h)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    out = torch.zeros_like(color)
    gp.execute(color.data_ptr()); torch.cuda.synchronize()
    img = color.cpu().numpy()
    for kw in ({}, {"autoExposure": True}, {"whiteBalance": True, "whitePoint": 4000.0, "exposureCompensation": 0.7, "clamp": False},
               {"fNumber": 1.4, "shutter": 2.0, "filmSpeed": 200.0, "whiteScale": 6.0, "whiteMaxLuminance": 2.5}):
        tm = ToneMapper(dict(kw, operator=op))
        avg = tm.execute(color.data_ptr(), out.data_ptr(), w, h, want_average=True)
        torch.cuda.synchronize()
        M = po.tonemap_color_transform(**{k: v for k, v in kw.items() if k in ("exposureCompensation", "autoExposure", "filmSpeed", "whiteBalance", "whitePoint", "fNumber", "shutter")})
        want = po.tonemap(img, M, op, kw.get("autoExposure", False), kw.get("clamp", True), kw.get("whiteScale", 11.2), kw.get("whiteMaxLuminance", 1.0))
        if kw.get("autoExposure"):
            assert abs(avg - po.tonemap_avg_log_luminance(img)) < 2e-5 * max(1.0, abs(avg))
        got = out.cpu().numpy()
        np.testing.assert_allclose(got[..., :3], want[..., :3], rtol=3e-5, atol=2e-6, err_msg=str((op, kw)), equal_nan=True)
        assert np.array_equal(got[..., 3], img[..., 3])
        assert np.nanstd(got[..., :3]) > 1e-3


@pytest.mark.parametrize("dim", [(96, 80, 72), (97, 63, 45)])
def test_gpu_mip_builder_matches_host_builder(dim):
    """the mip / conservative-mip chain built on the GPU from a dense grid stores exactly what the host
    builder's brick pools store (fp32 mip 0 after the 1e-9 flush, UNORM8 codes + scale elsewhere), on even and odd
    (3-tap polyphase) dimensions, normal and conservative chain."""
    import torch
    from volumetricrestirrelease_b200 import Scene
    from volumetricrestirrelease_b200.mipbuild import build_mips
    nx, ny, nz = dim
    rng = np.random.default_rng(11)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    blob = np.exp(-(((x - nx / 2) / (nx / 4)) ** 2 + ((y - ny / 2) / (ny / 4)) ** 2 + ((z - nz / 2) / (nz / 4)) ** 2))
    dense = (np.clip(blob + 0.3 * rng.random((nz, ny, nx)) - 0.55, 0, None) * 2.0).astype(np.float32)   # sparse: zeros outside the blob
    dense[dense < 0.05] = 0.0
    dense[nz // 2, ny // 2, nx // 2] = 1e-12                                                             # below the 1e-9 flush
    sc = Scene()
    vol = sc.addGVDBVolume(dense=dense, numMips=4)
    chain = build_mips(torch.from_numpy(dense).cuda(), 4)
    built = 0
    for m in range(chain.num_mips):
        for cons in (False, True):
            want = vol.dense_mip(m, cons)
            t, scale = chain.level(m, cons)
            got = t.cpu().numpy()
            if got.dtype == np.uint8:
                got = got.astype(np.float32) * np.float32(0.003921568859368563) * np.float32(scale)
            assert got.shape == want.shape, (m, cons, got.shape, want.shape)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), f"mip {m} conservative {cons}: {(got != want).sum()} voxels differ"
            built += 1
    assert built == 8 and (dense == 0).mean() > 0.3
    chain.close()


def _blob_dense(dim, seed, shift=0.0):
    nx, ny, nz = dim
    rng = np.random.default_rng(seed)
    z, y, x = np.meshgrid(np.arange(nz), np.arange(ny), np.arange(nx), indexing="ij")
    blob = np.exp(-(((x - nx / 2 - shift) / (nx / 4)) ** 2 + ((y - ny / 2) / (ny / 4)) ** 2 + ((z - nz / 2) / (nz / 4)) ** 2))
    d = (np.clip(blob + 0.3 * rng.random((nz, ny, nx)) - 0.55, 0, None) * 2.0).astype(np.float32)
    d[d < 0.05] = 0.0
    return d


def test_volume_from_gpu_chain_renders_identically():
    """SURVEY 8f rank 2, second half: density slots bound from a GPU-built chain (brick pools, quad repacks, brick bounds made on
    the device, tree from the GPU activity map) render the same bits as the host-built volume — also as the next frame of an
    animated sequence (advance) and with a tracking method that reads the brick bounds."""
    import torch
    from volumetricrestirrelease_b200 import Scene
    from volumetricrestirrelease_b200.mipbuild import build_mips
    dim, w, h = (97, 88, 75), 128, 96
    dA, dB = _blob_dense(dim, 3), _blob_dense(dim, 4, shift=6.0)

    def scene_of(dense):
        sc = Scene()
        sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), dense=dense, numMips=4, densityScale=0.6, voxelSize=0.05)
        sc.setEnvMap((256, 128), seed=7)
        sc.frame_camera(1.1)
        return sc

    scA, scB = scene_of(dA), scene_of(dB)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")

    def frames(gp, n):
        out = []
        for _ in range(n):
            gp.execute(color.data_ptr())
            torch.cuda.synchronize()
            out.append(color.cpu().numpy().copy())
        return out

    for kw in ({}, {"mFinalVisibilityTrackingMethod": capi.kResidualRatioTracking, "mFinalLightTrackingMethod": capi.kResidualRatioTracking}):
        p = VolumetricReSTIRParams(**kw)
        # (1) static: host-built B  vs  template A + chain(B)
        ref = VolumetricReSTIR.create({"mParams": p}); ref.setScene(scB, w, h)
        want = frames(ref, 2)
        gp = VolumetricReSTIR.create({"mParams": p}); gp.setScene(scA, w, h)
        chainB = build_mips(torch.from_numpy(dB).cuda(), 4)
        gp.setVolumeFromChain(chainB)
        got = frames(gp, 2)
        for a, b in zip(got, want):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"static volume from chain differs ({kw})"
        # (2) animated: A then B as the next frame (advanceVolume rebinds the Scene object's volume: one Scene per pass)
        ref = VolumetricReSTIR.create({"mParams": p}); ref.setScene(scene_of(dA), w, h)
        want = frames(ref, 1)
        ref.advanceVolume(scB.volume)
        want += frames(ref, 2)
        gp = VolumetricReSTIR.create({"mParams": p}); gp.setScene(scene_of(dA), w, h)
        got = frames(gp, 1)
        gp.setVolumeFromChain(chainB, advance=True)
        got += frames(gp, 2)
        for f, (a, b) in enumerate(zip(got, want)):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), f"animated volume from chain differs at frame {f} ({kw})"
        chainB.close()
    assert (want[-1][..., :3].sum(-1) > 0).mean() > 0.05


@pytest.mark.parametrize("kind,dim", [("bunny", (96, 80, 72)), ("shells", (150, 140, 136)), ("cloud", (64, 64, 64))])
def test_device_built_volume_equals_host_built(kind, dim):
    """SURVEY 8d config 5 machinery at test size: the procedural field evaluated on the device, the chain built there and bound
    over a voxel-less template must give (1) the host generator's voxels bit for bit, (2) slot for slot the host builder's
    tree, child lists and brick pools (downloaded back with vrestir_download_volume), (3) identical frames."""
    import ctypes as C
    import torch
    from volumetricrestirrelease_b200 import Scene
    w, h = 128, 96
    kw = dict(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), dataFile=kind, numMips=3, densityScale=0.2, dim=dim, seed=5, voxelSize=0.5)
    host = Scene(); host.addGVDBVolume(**kw)
    dev = Scene(); dev.addGVDBVolumeDevice(0, keep_dense=True, **kw)
    assert np.array_equal(dev.volume.dense.cpu().numpy().view(np.uint32), host.volume.dense_mip(0).view(np.uint32))
    for sc in (host, dev):
        sc.setEnvMap((256, 128), seed=7); sc.frame_camera(1.1)
    color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
    imgs = []
    passes = []
    for sc in (host, dev):
        gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams()}); gp.setScene(sc, w, h)
        passes.append(gp)
        frames = []
        for _ in range(3):
            gp.execute(color.data_ptr()); torch.cuda.synchronize()
            frames.append(color.cpu().numpy().copy())
        imgs.append(frames)
    for a, b in zip(*imgs):
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
    assert (imgs[0][-1][..., :3].sum(-1) > 0).mean() > 0.05
    dev.volume.release_chain()
    back = passes[1].downloadVolume()
    ga, gb = host.volume.grid.contents, back.grid.contents
    for slot in list(range(3)) + list(range(8, 11)):
        a, b = ga.slots[slot], gb.slots[slot]
        assert a.valid and b.valid and a.top_lev == b.top_lev and a.brick_count == b.brick_count and a.max_value == b.max_value, slot
        for l in range(3):
            assert a.node_count[l] == b.node_count[l] and a.childlist_count[l] == b.childlist_count[l]
            if a.node_count[l]:
                assert C.string_at(a.nodes[l], a.node_count[l] * 32) == C.string_at(b.nodes[l], b.node_count[l] * 32), (slot, l)
            if a.childlist_count[l]:
                assert C.string_at(a.childlist[l], a.childlist_count[l] * 4) == C.string_at(b.childlist[l], b.childlist_count[l] * 4), (slot, l)
        nb = a.brick_count * 1000 * (1 if a.atlas_format == 1 else 4)
        assert C.string_at(a.atlas, nb) == C.string_at(b.atlas, nb), slot
    # the downloaded grid drives the oracle: same frame 0 as the GPU
    from oracle import vro
    import copy
    sc2 = copy.copy(dev); sc2.volume = back
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc2, w, h, importance=passes[1].get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32), env_alias=passes[1].env_alias())
    e = rel_err_image(imgs[1][0], op.execute())
    assert (e > RADIANCE_RTOL).mean() <= 5e-3


def test_ragged_frame_size_staged():
    """A frame whose width / height are not multiples of the 16x8 CTA tile or the 8x4 warp tile (partial tiles on the right and
    bottom edges): staged parity against the oracle through every stage, default options."""
    w, h = 150, 91
    out = _staged(VolumetricReSTIRParams(), env_scene(), w, h, frames=2)
    _check_staged(out, w, h, "ragged 150x91")
