"""GPU-vs-oracle staged parity on RANDOM COMBINATIONS of the option surface (tests/test_gpu_options.py varies one option at a
time): every seed draws each option independently from its table, builds the scene the draw asks for (bunny cloud / emissive plume /
three-level tree, phase anisotropy, a point light, emissive triangles, full or ragged frame), picks the task-stream path or the
per-pixel kernels, and runs two frames stage by stage with a moving camera.  Seeds are fixed, so a failure reproduces."""
import numpy as np
import pytest

from common import FLIP_BUDGET, capi, check_staged, env_scene, staged
from volumetricrestirrelease_b200 import Scene, VolumetricReSTIRParams

pytestmark = pytest.mark.gpu


TABLE = {
    "mMaxBounces": [1, 2, 3, 4],
    "mInitialM": [1, 3, 4, 6],
    "mInitialBaseMipLevel": [0, 1, 2],
    "mInitialVisibilityUseLinearSampler": [0, 0, 1],
    "mInitialLightingMipLevel": [1, 2],
    "mInitialLightingTrackingMethod": [capi.kRayMarching, capi.kRayMarching, capi.kAnalyticTracking, capi.kResidualRatioTracking],
    "mInitialLightSamples": [0, 1, 1],
    "mInitialUseRussianRoulette": [0, 1],
    "mInitialUseCoarserGridForIndirectBounce": [0, 1],
    "mTemporalMISMethod": [capi.kMISNone, capi.kMISTalbot],
    "mTemporalReuseMThreshold": [1.0, 4.0, 10.0],
    "mTemporalReprojectionMode": [capi.kReprojectionLinear, capi.kReprojectionLinear, capi.kReprojectionNone, capi.kReprojectionNoBackground],
    "mSpatialMISMethod": [capi.kMISNone, capi.kMISTalbot],
    "mSpatialReuseRounds": [1, 2],
    "mSpatialSampleCount": [2, 4, 5],
    "mSampleRadius": [4.0, 10.0, 17.0],
    "mRandomSamplerType": [capi.kR2, capi.kHammersley],
    "mSpatialVisibilityTrackingMethod": [capi.kRayMarching, capi.kRayMarching, capi.kAnalyticTracking, capi.kRatioTracking],
    "mSpatialLightingTrackingMethod": [capi.kRayMarching, capi.kRayMarching, capi.kAnalyticTracking, capi.kResidualRatioTracking],
    "mSpatialVisibilityMipLevel": [1, 2],
    "mSpatialLightingMipLevel": [1, 2],
    "mSpatialVisibilityUseLinearSampler": [0, 1, 1],
    "mSpatialLightingUseLinearSampler": [0, 1, 1],
    "mSpatialVisibilityTStepScale": [0.5, 1.0, 2.0],
    "mFinalVisibilityTrackingMethod": [capi.kAnalyticTracking, capi.kAnalyticTracking, capi.kRayMarching, capi.kResidualRatioTracking],
    "mFinalLightTrackingMethod": [capi.kAnalyticTracking, capi.kAnalyticTracking, capi.kRatioTracking],
    "mFinalLightSamples": [1, 2],
    "mVertexReuse": [0, 0, 1],
    "mVertexReuseStartBounce": [1, 2],
}


def _draw(seed):
    rng = np.random.default_rng(1000 + seed)
    kw = {k: v[int(rng.integers(len(v)))] for k, v in TABLE.items()}
    scene = dict(g=[0.0, 0.5, -0.3][int(rng.integers(3))], point_light=bool(rng.integers(2)), emissive=bool(rng.integers(3) == 0))
    if kw["mVertexReuseStartBounce"] >= kw["mMaxBounces"]:
        kw["mVertexReuseStartBounce"] = 1
    kw["mUseAnalyticLights"], kw["mUseEmissiveLights"] = int(scene["point_light"]), int(scene["emissive"])
    rng2 = np.random.default_rng(5000 + seed)              # a second stream, so that the option draws above stay what they were
    scene["kind"] = ["bunny", "bunny", "plume", "three_level"][int(rng2.integers(4))]
    scene["path"] = ["task streams", "task streams", "per-pixel kernels"][int(rng2.integers(3))]
    scene["frame"] = [(96, 64), (96, 64), (83, 47)][int(rng2.integers(3))]       # a ragged frame: partial tiles on both axes
    return kw, scene


@pytest.mark.parametrize("seed", range(48))
def test_random_option_combination_staged(seed):
    kw, scene = _draw(seed)
    if scene["kind"] == "plume":                            # temperature grid: volume emission and self-emission samples
        sc = Scene()
        sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=scene["g"], dataFile="plume", numMips=4, densityScale=0.1, hasVelocity=True,
                         hasEmission=True, LeScale=0.05, temperatureCutoff=1.0, temperatureScale=100.0, dim=(64, 96, 64), seed=3, voxelSize=1.0)
        sc.setEnvMap((256, 128), seed=7); sc.setEnvMapIntensity(0.5); sc.frame_camera(1.0)
    elif scene["kind"] == "three_level":                    # a level-2 root above the 128-voxel nodes
        sc = env_scene(dim=(200, 150, 140), density_scale=0.05, g=scene["g"])
        assert sc.volume.grid.contents.slots[0].top_lev == 2
    else:
        sc = env_scene(dim=(64, 64, 56), density_scale=0.15, g=scene["g"])
    W, H = scene["frame"]
    lo, hi = sc.volume_bounds_world()
    if scene["point_light"]:
        sc.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
    if scene["emissive"]:
        sc.addEmissiveShell(300, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
    p0 = np.array(sc.camera.position)
    path = [tuple(p0 + np.array((0.7, 0.2, -0.3)) * 12.0 * f) for f in range(2)]
    print(f"[combo {seed}] {kw} {scene}")
    out = staged(VolumetricReSTIRParams(**kw), sc, W, H, frames=2, camera_path=path, own_tables=scene["emissive"],
                 dict_={"mUseWavefront": int(scene["path"] == "task streams")})
    check_staged(out, W, H, f"combo{seed}", budget=5 * FLIP_BUDGET)


@pytest.mark.parametrize("seed", range(0, 48, 4))
def test_random_option_combination_whole_frames_are_path_independent(seed):
    """The same random combinations through the un-staged product call over four frames with an announced moving camera: the
    pipelined frame (K0/K1 of frame f+1 ahead on their own stream, deferred K5), the serial frame and the per-pixel kernels must
    produce the same bits on every frame."""
    import copy
    import torch
    from volumetricrestirrelease_b200 import VolumetricReSTIR
    kw, scene = _draw(seed)
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15, g=scene["g"])
    lo, hi = sc.volume_bounds_world()
    if scene["point_light"]:
        sc.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
    if scene["emissive"]:
        sc.addEmissiveShell(300, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
    w, h, frames = scene["frame"][0] * 2, scene["frame"][1] * 2, 4
    p0 = np.array(sc.camera.position)
    path = [tuple(p0 + np.array((0.7, 0.2, -0.3)) * 6.0 * f) for f in range(frames + 1)]

    def run(extra):
        gp = VolumetricReSTIR.create(dict({"mParams": VolumetricReSTIRParams(**kw)}, **extra))
        sc.camera.position = path[0]
        gp.setScene(sc, w, h)
        color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
        imgs = []
        for f in range(frames):
            sc.camera.position = path[f]
            gp.updateCamera()
            if extra.get("mPipelineFrames"):
                nxt = copy.copy(sc.camera)
                nxt.position = path[f + 1]
                gp.setNextCamera(nxt)
            gp.execute(color.data_ptr())
            gp.wait_output()
            imgs.append(color.cpu().numpy().view(np.uint32).copy())
        torch.cuda.synchronize()
        return imgs, gp.pipeline_stats()

    serial, _ = run({"mPipelineFrames": 0})
    pipelined, st = run({"mPipelineFrames": 2})
    per_pixel, _ = run({"mPipelineFrames": 0, "mUseWavefront": 0})
    assert st["adopted"] in (0, frames - 1) and st["discarded"] == 0      # K1 runs ahead only on its specialised task-stream path
    for f in range(frames):
        assert np.array_equal(pipelined[f], serial[f]), f"frame {f}: pipelined != serial"
        assert np.array_equal(per_pixel[f], serial[f]), f"frame {f}: per-pixel kernels != task streams"
    assert (serial[-1].view(np.float32)[..., :3].sum(-1) > 0).mean() > 0.05
