"""Stage-level witnesses (TEST INFRASTRUCTURE): every stage of the pass — K0 features (march_witness), K1 initial candidates, K2
temporal reuse, K3 spatial reuse, K5 final shading, the target function p-hat and the reference path tracer — written from the
reference's Slang and composed from the lower-level witnesses (march_witness: trackers, distance sampling, point query, xoshiro;
light_witness: env map, phase function; alias_oracle: emissive power table) — not from oracle/vr_oracle.cpp or the CUDA kernels.
Covered: 1-4 bounces, reuse and no-reuse mode, env-map / point / directional / emissive-triangle lights, volume emission, vertex
reuse, animated volumes (velocity reprojection, previous-frame grids), every tracker.  Not covered: surface scenes (not built).
Pure-Python float32, one pixel at a time: tests compare a dozen pixels per case with the oracle's buffers.

  camera ray                    F/Scene/Camera/Camera.slang:160-228 (computeRayPinholeScaled(pixel, 1, 0.5))
  evaluate_F_ / evaluate_P_hat  VR/ReSTIRHelper.slang:91-200,426-441 + evaluate_L_in_volume :442-496 (env light)
  option -> mip / sampler       VR/VolumetricReSTIR.cpp:455-500 (gSpatialSamplingOptions)
  initial candidates (K1)       VR/TraceRays.cs.slang:64-183, VR/ComputeInitialSample.slang:4-395, SampleDirectLighting /
                                sampleSceneLights VR/VolumeUtils.slang:12-66,454-492 (env-map light), gInitialSamplingOptions
                                VR/VolumetricReSTIR.cpp:457-470
  temporal reuse (K2)           VR/TemporalReuse.cs.slang:80-377 (reprojection modes, velocity grid, Talbot MIS or none),
                                resampleNeighbor VR/ReSTIRHelper.slang:582-597, simpleResampleStepWithMaxM VR/Reservoir.slang:57-87
  multi-bounce paths            the bounce loop of VR/ComputeInitialSample.slang:29-386 (phase-sampled continuation, one free-flight
                                sample per bounce on the coarser grid, Russian roulette, per-bounce reservoir streaming), the
                                extra-bounce records VR/ReSTIRHelper.slang:21-52,66-79 and the vertex loop of evaluate_F_ :205-385
  analytic + emissive lights    sampleSceneLights VR/VolumeUtils.slang:12-149 (type selection, point / directional F/Experimental/Scene/Lights/
                                LightHelpers.slang:196-244, emissive triangles F/.../EmissivePowerSampler.slang:57-92 +
                                EmissiveLightSamplerHelpers.slang:56-101, alias draw F/Utils/Sampling/AliasTable.slang:56-70,
                                computeRayOrigin F/Utils/Helpers.slang:92-105), evaluate_L_in_volume VR/ReSTIRHelper.slang:443-496
  volume emission               EmissionWorldSpace / ConvertTempToColor VR/VolumeBase.slang:196-232 (temperature grid = slot 16, 128-texel
                                black-body table F/Scene/Scene.cpp:2877-2892), the emission-vs-scatter draw and the emissive scatter
                                vertex of VR/ComputeInitialSample.slang:267-334, encodeEmissivePosition VR/ReSTIRHelper.slang:9-19
  reference path tracer         IntegrateByVolumePathTracing VR/VolumePathTracingFunctions.slang:3-131, directLighting
                                VR/VolumeUtils.slang:417-447, the gUseReference branch of VR/TraceRays.cs.slang:85-103
  vertex reuse (VERTEX_REUSE)   VR/ReSTIRHelper.slang:21-27,222-240,289-296,345-352,392-420, VR/ComputeInitialSample.slang:88-94,116-125,
                                324-331, VR/Reservoir.slang:46-47: records hold world-space vertices from bounce S on, the path density turns
                                to area measure at vertex S, p_partial = the suffix of p-hat past that vertex
  final shading (K5)            VR/FinalShading.cs.slang:95-141, finalOptions VR/VolumetricReSTIR.cpp:484-496
  spatial reuse                 VR/SpatialReuse.cs.slang:64-265, resampleNeighborSpatialReuse VR/ReSTIRHelper.slang:563-580,
                                simpleResampleStep VR/Reservoir.slang:26-55, sample_disk F/Utils/Math/MathHelpers.slang:242-250,
                                R2Params / round seeds VR/VolumetricReSTIR.cpp:452-455,674-680
"""
import math

import numpy as np

from . import light_witness as lw
from .march_witness import Witness, Xoshiro

F = np.float32
K_RAY_TMAX = F(3.402823466e+38)
SELF_EMISSION = -3


def _env(frame, direction):
    """EnvMapSampler.eval: radiance of the env map along a direction (black without an env map)."""
    if frame.sc.envMap is None:
        return np.zeros(3, F)
    return lw.env_eval(frame.sc.envMap, direction, frame.sc.envMapIntensity)


class Frame:
    """Scene + options + camera of one frame, with the per-mip march witnesses cached."""

    def __init__(self, scene, params, width, height):
        self.sc, self.P, self.w, self.h = scene, params, width, height
        self.grid = scene.volume.grid.contents
        cam = scene.camera.data(width, height)
        self.origin = np.array(cam.posW[:], dtype=F)
        self.U, self.V, self.Wv = (np.array(getattr(cam, k)[:], dtype=F) for k in ("cameraU", "cameraV", "cameraW"))
        self._wit = {}
        self.lights = None              # a Lights instance when the scene has analytic / emissive lights
        self.prev_grid = None           # the previous frame's grid description of an animated volume (slots 19.. / 27 / 28)
        self.last_frame = False         # isLastFrame of the evaluation in progress
        self.final_rng = None           # the final shading's generator (stochastic trackers only)
        self.stage_rng = None           # the generator of the K1 / K2 / K3 invocation in progress (stochastic trackers in p-hat)

    def seed_final(self, px, py, frame_count):
        """FinalShading.cs.slang:85: the pixel's generator of the last round of the frame."""
        P = self.P
        total_rounds = (P.mSpatialReuseRounds if P.mEnableSpatialReuse else 0) + int(bool(P.mEnableTemporalReuse)) + 1 + 1
        self.final_rng = Xoshiro(px, py, total_rounds * frame_count + total_rounds - 1)

    def wit(self, mip):
        """The march witness of a slot; while an isLastFrame evaluation with usePrevGridForReproj is in progress, density slots
        resolve to the previous frame's grids (offset 19) and the temperature grid to slot 27."""
        if self.last_frame and self.prev_grid is not None and self.P.mUsePrevVolumeForReproj:
            mip = mip + (11 if mip >= 16 else 19)
        if mip not in self._wit:
            if mip >= 19:
                src = mip - 11 if mip >= 27 else mip - 19
                self._wit[mip] = Witness(self.prev_grid, src, self.grid.volume, mip)
            else:
                self._wit[mip] = Witness(self.grid, mip)
        return self._wit[mip]

    def velocity_world(self, p):
        """VelocityWorld (VR/VolumeBase.slang:183-194): trilinear 3-channel point query of slot 17, rotated to world space."""
        v = np.array([self.wit(17).value_world(p, ch) for ch in range(3)], dtype=F)
        M = np.array(self.grid.volume.externalModelToWorld[:], dtype=F).reshape(4, 4)
        return np.array([v[0] * M[0, j] + v[1] * M[1, j] + v[2] * M[2, j] + F(0) * M[3, j] for j in range(3)], dtype=F)

    def ray_dir(self, px, py):
        p = np.array([(F(px) + F(0.5)) / F(self.w), (F(py) + F(0.5)) / F(self.h)], dtype=F)
        ndc = np.array([F(2) * p[0] + F(-1), F(-2) * p[1] + F(1)], dtype=F)
        d = ndc[0] * self.U + ndc[1] * self.V + self.Wv
        return (d / np.sqrt(np.dot(d, d))).astype(F)

    def compute_visibility(self, origin, direction, tmax, samples, mip, linear, method, tstep_scale, rng):
        """computeVisibility (VR/VolumeUtils.slang:549-572): `samples` estimates of the chosen tracker, averaged.  The random-walk
        trackers (ratio 0 / residual ratio 3 / analog residual ratio 4) always sample trilinearly and draw from `rng`."""
        total = F(0)
        for _ in range(samples):
            if method in (0, 3, 4):
                v = self.wit(mip).residual_ratio_tracking(origin, direction, tmax, rng, method == 4, method == 0)
            elif method == 2:
                v = self.wit(mip).ray_marching(origin, direction, tmax, linear, tstep_scale)
            else:
                v = self.wit(mip).analytic(origin, direction, tmax, linear)
            total = F(total + F(v))
        return F(total / F(samples))

    def _transmittance(self, final, which, origin, direction, tmax):
        """The segment transmittance of a p-hat evaluation (spatial options: VR/VolumetricReSTIR.cpp:471-482, one sample) or of the
        final shading (finalOptions :484-496: mip 0, trilinear, the configured tracker and sample count).  Stochastic trackers draw
        from the generator of the stage in progress (self.stage_rng / self.final_rng), in shader order."""
        P = self.P
        if final:
            cam = which == "camera"
            method = P.mFinalVisibilityTrackingMethod if cam else P.mFinalLightTrackingMethod
            n = 1 if method in (1, 2) else (P.mFinalVisibilitySamples if cam else P.mFinalLightSamples)
            return self.compute_visibility(origin, direction, tmax, n, 0, True, method, P.mFinalTStepScale, self.final_rng)
        if which == "camera":
            return self.compute_visibility(origin, direction, tmax, 1, P.mSpatialVisibilityMipLevel, bool(P.mSpatialVisibilityUseLinearSampler),
                                           P.mSpatialVisibilityTrackingMethod, P.mSpatialVisibilityTStepScale, self.stage_rng)
        return self.compute_visibility(origin, direction, tmax, 1, P.mSpatialLightingMipLevel, bool(P.mSpatialLightingUseLinearSampler),
                                       P.mSpatialLightingTrackingMethod, P.mSpatialLightingTStepScale, self.stage_rng)

    def eval_F(self, d, depth, light_uv, light_id, final=False):
        """evaluate_F_ of a single-bounce sample (depth along direction d from the camera, env light stored as (uv, id)): float3."""
        vol = self.grid.volume
        o = self.origin
        if depth == K_RAY_TMAX:
            vis = self._transmittance(final, "camera", o, d, float(K_RAY_TMAX))
            return (vis * _env(self, d)).astype(F)
        pw = (o + d * F(depth)).astype(F)
        density = self.wit(0).density_world(pw)
        if density == 0:
            return np.zeros(3, F)
        vis = self._transmittance(final, "camera", o, d, float(depth))
        Fv = (vis * density * np.array((vol.sigma_a if light_id == SELF_EMISSION else vol.sigma_s)[:], dtype=F)).astype(F)
        if not bool(np.any(Fv > 0)):
            return Fv
        if light_id == SELF_EMISSION:
            return (Fv * emission_world(self, pw)).astype(F)
        if light_id >= 0:
            return (Fv * eval_L_in_volume(self, self.lights, pw, -d, light_id, light_uv, final)).astype(F)
        zz = F(1) - light_uv[0] * light_uv[0] - light_uv[1] * light_uv[1]
        z = np.sqrt(zz).astype(F) if zz >= 0 else F(0)
        wi = np.array([light_uv[0], light_uv[1], -z if light_id == -2 else z], dtype=F)
        Ld = _env(self, wi) * F(lw.phase_hg(float(np.dot(-d, wi)), vol.PhaseFunctionConstantG))
        tr = self._transmittance(final, "light", pw, wi, float(K_RAY_TMAX))
        return (Fv * (tr * Ld)).astype(F)

    def p_hat(self, d, depth, light_uv, light_id):
        return F(lw.luminance(self.eval_F(d, depth, light_uv, light_id)))

    def final_shading(self, px, py, r, frame_count=None):
        """FinalShading.cs.slang:95-141 for one pixel: F of the stored sample under the final options times the RIS weight W.
        frame_count is needed only when a final tracker draws random numbers."""
        if frame_count is not None:
            self.seed_final(px, py, frame_count)
        if not r["runningSum"] > 0:
            return np.zeros(3, F)
        col = self.eval_F(self.ray_dir(px, py), F(r["depth"]), np.asarray(r["lightUV"], dtype=F), int(r["lightID"]), final=True)
        W = F(1) if r["p_y"] == 0 else F(F(r["runningSum"]) / F(F(r["p_y"]) * F(r["M"])))
        out = (col * W).astype(F)
        return np.zeros(3, F) if bool(np.any(np.isnan(out) | np.isinf(out))) else out


def neighbor_offset(P, sample_id, frame_id):
    """generateNeighborOffset: R2 sequence (fp64 multiplier) or Hammersley points (F/Utils/Helpers.slang:62-75) + sample_disk,
    truncated to int2."""
    if P.mRandomSamplerType == 0:                         # kHammersley (0; kR2 = 1): (i / N, bit-reversed i)
        i = sample_id
        i = ((i & 0x55555555) << 1) | ((i & 0xAAAAAAAA) >> 1)
        i = ((i & 0x33333333) << 2) | ((i & 0xCCCCCCCC) >> 2)
        i = ((i & 0x0F0F0F0F) << 4) | ((i & 0xF0F0F0F0) >> 4)
        i = ((i & 0x00FF00FF) << 8) | ((i & 0xFF00FF00) >> 8)
        i = ((i << 16) | (i >> 16)) & 0xFFFFFFFF
        u = (F(F(sample_id) / F(P.mSpatialSampleCount)), F(F(i) * F(2.3283064365386963e-10)))
    elif sample_id == 0:
        u = (F(0), F(0))
    else:
        m = float(frame_id * P.mSpatialSampleCount + sample_id)
        u = (F((0.754877669 * m) % 1.0), F((0.569840296 * m) % 1.0))
    r = np.sqrt(u[0]).astype(F)
    phi = F(6.28318530717958647692) * u[1]
    px, py = F(r * np.cos(phi).astype(F)), F(r * np.sin(phi).astype(F))
    rad = F(P.mSampleRadius)
    return int(np.trunc(rad * px)), int(np.trunc(rad * py))


def _resample_step(tap, state, rng):
    """simpleResampleStep: tap / state are dicts with runningSum, M, depth, p_y, lightUV, lightID, sampledPixel."""
    w = tap["runningSum"]
    state["M"] = F(state["M"] + tap["M"])
    if not w > 0:
        return False
    state["runningSum"] = F(state["runningSum"] + w)
    sel = bool(rng.next1d() * state["runningSum"] < w)
    if sel:
        for k in ("depth", "p_y", "lightUV", "lightID", "sampledPixel") + tuple(x for x in ("src", "p_partial") if x in tap):
            state[k] = tap[k]
    return sel


def spatial_reuse_pixel(frame, res_in, features, px, py, frame_count, round_id=0, extra_in=None, pp_in=None):
    """SpatialReuse.cs.slang main() for one pixel (Talbot MIS or none, R2 sampler).  res_in: (H, W) structured reservoirs;
    features: (H, W) structured {noReflectiveSurface, transmittance}.  Returns the output reservoir as a dict.  With extra_in
    ((H, W, B-1, 3) extra-bounce records: MAX_BOUNCES > 1) the targets are evaluated on whole paths and the result is
    (reservoir, the extra-bounce records of the pixel the selected sample came from)."""
    P, w, h = frame.P, frame.w, frame.h
    talbot = P.mSpatialMISMethod == 1
    rec = lambda r: dict(runningSum=F(r["runningSum"]), M=F(r["M"]), depth=F(r["depth"]), p_y=F(r["p_y"]),
                         lightUV=np.array(r["lightUV"], dtype=F), lightID=int(r["lightID"]), sampledPixel=int(r["sampledPixel"]))
    if pp_in is not None:                                   # vertex reuse: the p_partial plane travels with the reservoirs
        plain = res_in

        class _WithPP:
            def __getitem__(self, idx):
                return _RecWithPP(plain[idx], pp_in[idx])
        res_in = _WithPP()
        rec = lambda r: dict(runningSum=F(r["runningSum"]), M=F(r["M"]), depth=F(r["depth"]), p_y=F(r["p_y"]), lightUV=np.array(r["lightUV"], dtype=F),
                             lightID=int(r["lightID"]), sampledPixel=int(r["sampledPixel"]), p_partial=F(r["p_partial"]))
    if extra_in is not None:
        out = _spatial_reuse_pixel(_PathTargets(frame, extra_in), rec, res_in, features, px, py, frame_count, round_id)
        src = out.pop("src", (px, py))
        return out, extra_in[src[1], src[0]].copy()
    return _spatial_reuse_pixel(frame, rec, res_in, features, px, py, frame_count, round_id)


class _RecWithPP:
    """A structured reservoir record plus its p_partial."""

    def __init__(self, r, pp):
        self.r, self.pp = r, pp

    def __getitem__(self, k):
        return self.pp if k == "p_partial" else self.r[k]


class _PathTargets:
    """p-hat over whole paths: the extra-bounce records travel with the tap's source pixel."""

    def __init__(self, frame, extra):
        self.frame, self.extra = frame, extra
        self.P, self.w, self.h, self.ray_dir = frame.P, frame.w, frame.h, frame.ray_dir

    def p_hat_tap(self, d, tap):
        x, y = tap["src"]
        return F(lw.luminance(eval_F_path(self.frame, d, tap, self.extra[y, x], spatial_reuse=True)))


def _spatial_reuse_pixel(frame, rec, res_in, features, px, py, frame_count, round_id):
    P, w, h = frame.P, frame.w, frame.h
    talbot = P.mSpatialMISMethod == 1
    paths = isinstance(frame, _PathTargets)
    p_hat = (lambda d, tap: frame.p_hat_tap(d, tap)) if paths else (lambda d, tap: frame.p_hat(d, tap["depth"], tap["lightUV"], tap["lightID"]))
    r2_seed = ((P.mSpatialReuseRounds + 1) * frame_count + round_id) % 16
    round_offset = int(bool(P.mEnableTemporalReuse)) + 1
    num_rounds = P.mSpatialReuseRounds + round_offset + 1
    rng = Xoshiro(px, py, num_rounds * frame_count + round_id + round_offset)
    getattr(frame, "frame", frame).stage_rng = rng
    center = rec(res_in[py, px])
    if paths:
        center["src"] = (px, py)
    if features[py, px]["transmittance"] == 1.0:          # IsSelfBackground: passed through
        return center
    output = dict(runningSum=F(0), M=F(0), depth=K_RAY_TMAX, p_y=F(0), lightUV=np.zeros(2, F), lightID=0, sampledPixel=0) if talbot else center
    d0 = frame.ray_dir(px, py)
    S = P.mSpatialSampleCount
    offs = [neighbor_offset(P, s, r2_seed) for s in range(S)]
    inside = lambda x, y: 0 <= x < w and 0 <= y < h
    for s in range(0 if talbot else 1, S):
        tx, ty = px + offs[s][0], py + offs[s][1]
        if not inside(tx, ty):
            continue
        tap = rec(res_in[ty, tx])
        if paths:
            tap["src"] = (tx, ty)
        neighbor_py = tap["p_y"]
        if s > 0 and tap["runningSum"] != 0:            # resampleNeighborSpatialReuse
            ph = p_hat(d0, tap)
            with np.errstate(divide="ignore", invalid="ignore"):
                weight = F(ph / tap["p_y"])
            if np.isinf(weight) or np.isnan(weight):
                weight = F(0)
            tap["runningSum"] = F(tap["runningSum"] * weight)
            tap["p_y"] = ph
        mis = F(1)
        if talbot and tap["runningSum"] > 0:
            p_sum, p_qi, k = F(0), F(0), F(0)
            for j in range(S):
                jx, jy = px + offs[j][0], py + offs[j][1]
                if not inside(jx, jy):
                    continue
                tap2 = res_in[jy, jx]
                m2 = F(tap2["M"])
                k = F(k + m2)
                if j == 0:
                    p_qi = tap["p_y"]; p_sum = F(p_sum + tap["p_y"] * m2)
                elif s == j:
                    p_qi = F(tap2["p_y"]); p_sum = F(p_sum + F(tap2["p_y"]) * m2)
                else:
                    p_y = p_hat(frame.ray_dir(jx, jy), tap)
                    if np.isinf(p_y) or np.isnan(p_y):
                        p_y = F(0)
                    p_sum = F(p_sum + p_y * m2)
            if p_sum > 0:
                mis = F(p_qi * k / p_sum)
        tap["runningSum"] = F(tap["runningSum"] * mis)
        _resample_step(tap, output, rng)
    return output


def _new_reservoir():
    return dict(runningSum=F(0), M=F(0), depth=K_RAY_TMAX, p_y=F(0), lightUV=np.zeros(2, F), lightID=0, sampledPixel=0, p_partial=F(0))


def _initial_candidate(frame, d, hd, pd, tr, rng, mips):
    """ComputeInitialSample for one bounce: the candidate at distance hd along the camera ray (pdf pd, transmittance tr)."""
    return _initial_path(frame, d, hd, pd, tr, rng, mips)[0]


def initial_sampling_pixel(frame, px, py, frame_count, importance_mips, info=None):
    """TraceRays.cs.slang main() for one pixel (one bounce, reuse on, env-map light): M candidates by free-flight sampling along the
    camera ray, each with one importance-sampled env direction and its ray-marched shadow, streamed through a reservoir; then the
    p-hat of the streamed sample under the spatial options replaces the candidate-time target.  Returns the stored reservoir."""
    P = frame.P
    total_rounds = (P.mSpatialReuseRounds if P.mEnableSpatialReuse else 0) + int(bool(P.mEnableTemporalReuse)) + 1 + 1
    rng = Xoshiro(px, py, total_rounds * frame_count)
    frame.stage_rng = rng
    d = frame.ray_dir(px, py)
    final = _new_reservoir()
    linear = bool(P.mInitialVisibilityUseLinearSampler)
    vis_mip = P.mInitialBaseMipLevel if linear else P.mInitialBaseMipLevel + 8
    rounds = (P.mInitialM + 3) // 4
    for r in range(rounds):
        n = P.mInitialM - 4 * (rounds - 1) if r == rounds - 1 else 4
        hd, pd, ot = frame.wit(vis_mip).sample_distances(frame.origin, d, n, linear, rng)
        for s in range(n):
            cand = _initial_candidate(frame, d, F(hd[s]), F(pd[s]), F(ot[s]), rng, importance_mips)
            _resample_step(cand, final, rng)
    if info is not None:
        info["generator_after_candidates"] = list(rng.s)
    p_hat = frame.p_hat(d, final["depth"], final["lightUV"], final["lightID"])
    if info is not None:
        info["generator_after_p_hat"] = list(rng.s)
    if final["runningSum"] > 0:
        final["runningSum"] = F(final["runningSum"] * (F(0) if final["p_y"] == 0 else F(p_hat / final["p_y"])))
        final["p_y"] = p_hat
    return final


def _resample_step_max_m(tap, max_m, state, rng):
    """simpleResampleStepWithMaxM."""
    corrected = F(min(max_m, tap["M"]))
    w = F(0) if corrected == 0 else F(F(corrected / tap["M"]) * tap["runningSum"])
    state["M"] = F(state["M"] + corrected)
    if not w > 0:
        return False
    state["runningSum"] = F(state["runningSum"] + w)
    sel = bool(rng.next1d() * state["runningSum"] < w)
    if sel:
        for k in ("depth", "p_y", "lightUV", "lightID", "sampledPixel") + tuple(x for x in ("src", "p_partial") if x in tap):
            state[k] = tap[k]
    return sel


def temporal_reuse_pixel(frame, res_cur, res_prev, feat_cur, feat_prev, px, py, frame_count, prev_cam=None, extra_cur=None, extra_prev=None,
                         pp_cur=None, pp_prev=None, info=None):
    """TemporalReuse.cs.slang main() for one pixel of a frame with history.  prev_cam: (posW, U, V, W, view[16], proj[16]) of the
    previous frame (default: the current camera, i.e. a static camera).  Returns the reservoir K2 leaves in the current buffer; with
    extra_cur / extra_prev ((H, W, B-1, 3) extra-bounce records of the two frames) the targets are whole paths and the result is
    (reservoir, extra-bounce records of the selected sample)."""
    P, w, h = frame.P, frame.w, frame.h
    paths = extra_cur is not None
    if info is None:
        info = {}
    info["mvec"] = (F(0), F(0))                             # gMotionVec = (reprojected pixel - pixel) / resolution

    def target(direction, tap, depth=None):
        if not paths:
            return frame.p_hat(direction, tap["depth"] if depth is None else depth, tap["lightUV"], tap["lightID"])
        t = tap if depth is None else dict(tap, depth=depth)       # resampleNeighbor takes the tap inout (p_partial is rewritten),
        return F(lw.luminance(eval_F_path(frame, direction, t, tap["extra"])))     # evaluatePHatReadOnly a copy
    rec = lambda r: dict(runningSum=F(r["runningSum"]), M=F(r["M"]), depth=F(r["depth"]), p_y=F(r["p_y"]),
                         lightUV=np.array(r["lightUV"], dtype=F), lightID=int(r["lightID"]), sampledPixel=int(r["sampledPixel"]))
    cam = frame.sc.camera.data(w, h)
    if prev_cam is None:
        prev_cam = (frame.origin, frame.U, frame.V, frame.Wv, np.array(cam.viewMat[:], dtype=F), np.array(cam.projMat[:], dtype=F))
    p_pos, pU, pV, pW, p_view, p_proj = prev_cam
    talbot = P.mTemporalMISMethod == 1
    total_rounds = (P.mSpatialReuseRounds if P.mEnableSpatialReuse else 0) + int(bool(P.mEnableTemporalReuse)) + 1 + 1
    rng = Xoshiro(px, py, total_rounds * frame_count + 1)          # gRoundOffset = numInitialSamplingRounds
    frame.stage_rng = rng
    taps = [rec(res_cur[py, px]), None]
    if paths:
        taps[0]["extra"] = extra_cur[py, px].copy()
    if pp_cur is not None:
        taps[0]["p_partial"] = F(pp_cur[py, px])
    d = frame.ray_dir(px, py)
    o = frame.origin
    output = _new_reservoir() if talbot else dict(taps[0])
    is_bg = feat_cur[py, px]["transmittance"] == 1.0 and bool(feat_cur[py, px]["noReflectiveSurface"])
    fallback, reproj = True, (0, 0)
    if P.mTemporalReprojectionMode != 1:                          # != kReprojectionNone
        depth = taps[0]["depth"]
        if depth == K_RAY_TMAX and P.mTemporalReprojectionMode != 2 and not is_bg:
            depth = frame.wit(8 + P.mTemporalReprojectionMipLevel).rejection_sample_point(o, d, rng)
        pw = (o + d * depth).astype(F)
        vol = frame.grid.volume
        if vol.hasVelocity and frame.prev_grid is not None:       # hasAnimation: where was this point one frame ago
            pw = (pw - (frame.velocity_world(pw) * F(vol.velocityScale)).astype(F)).astype(F)
        view = np.array([pw[0] * p_view[0 + j] + pw[1] * p_view[4 + j] + pw[2] * p_view[8 + j] + F(1) * p_view[12 + j] for j in range(4)], dtype=F)
        clip = np.array([view[0] * p_proj[0 + j] + view[1] * p_proj[4 + j] + view[2] * p_proj[8 + j] + view[3] * p_proj[12 + j] for j in range(4)], dtype=F)
        if depth == K_RAY_TMAX:
            scr = (F(px) + F(0.5), F(py) + F(0.5)); scr_i = (px, py)
        else:
            with np.errstate(divide="ignore", invalid="ignore"):
                sx, sy = F(clip[0] / clip[3]), F(clip[1] / clip[3])
            scr = (F(F(F(0.5) * sx + F(0.5)) * F(w)), F(F(F(-0.5) * sy + F(0.5)) * F(h)))
            scr_i = (int(np.trunc(scr[0])), int(np.trunc(scr[1])))
        idx = (scr_i[1] * w + scr_i[0]) & 0xFFFFFFFF
        tf = feat_prev.reshape(-1)[idx] if idx < w * h else None          # out-of-range structured-buffer reads return 0
        tap_bg = tf is not None and tf["transmittance"] == 1.0 and bool(tf["noReflectiveSurface"])
        if is_bg and not tap_bg:
            info["mvec"] = (F(F(0 - px) / F(w)), F(F(0 - py) / F(h)))    # reprojScreenPos is still (0, 0) at this return
            return (taps[0], taps[0].pop("extra")) if paths else taps[0]     # K1's reservoir stays
        scr_i = (int(np.trunc(scr[0])), int(np.trunc(scr[1])))
        reproj = scr_i
        if 0 <= scr_i[0] < w and 0 <= scr_i[1] < h:
            taps[1] = rec(res_prev[scr_i[1], scr_i[0]]); fallback = False
    if fallback:
        reproj = (px, py); taps[1] = rec(res_prev[py, px])
    info["mvec"] = (F(F(reproj[0] - px) / F(w)), F(F(reproj[1] - py) / F(h)))
    if paths:
        taps[1]["extra"] = extra_prev[reproj[1], reproj[0]].copy()
    if pp_prev is not None:
        taps[1]["p_partial"] = F(pp_prev[reproj[1], reproj[0]])
    max_prev_m = F(F(P.mTemporalReuseMThreshold) * taps[0]["M"])

    def prev_dir(x, y):
        p = np.array([(F(x) + F(0.5)) / F(w), (F(y) + F(0.5)) / F(h)], dtype=F)
        ndc = np.array([F(2) * p[0] + F(-1), F(-2) * p[1] + F(1)], dtype=F)
        v = ndc[0] * pU + ndc[1] * pV + pW
        return (v / np.sqrt(np.dot(v, v))).astype(F)
    temporal_original_depth = taps[1]["depth"]
    if taps[1]["depth"] != K_RAY_TMAX:
        world = (p_pos + taps[1]["depth"] * prev_dir(*reproj)).astype(F)
        diff = world - o
        taps[1]["depth"] = F(np.sqrt(np.dot(diff, diff)))
    center_prev_depth = taps[0]["depth"]
    if center_prev_depth != K_RAY_TMAX:
        diff = (o + d * center_prev_depth).astype(F) - p_pos
        center_prev_depth = F(np.sqrt(np.dot(diff, diff)))
    saved_origin = frame.origin
    for i in range(0 if talbot else 1, 2):
        mis, neighbor_py = F(1), F(0)
        if taps[i]["p_y"] > 0:
            neighbor_py = taps[i]["p_y"]
            if np.isnan(taps[i]["runningSum"]) or np.isinf(taps[i]["runningSum"]):
                taps[i]["runningSum"] = F(0)
            if i > 0 and taps[i]["runningSum"] != 0:               # resampleNeighbor on the current ray
                ph = target(d, taps[i])
                with np.errstate(divide="ignore", invalid="ignore"):
                    weight = F(ph / taps[i]["p_y"])
                if np.isinf(weight) or np.isnan(weight):
                    weight = F(0)
                taps[i]["runningSum"] = F(taps[i]["runningSum"] * weight); taps[i]["p_y"] = ph
        else:
            taps[i]["p_y"] = F(0); taps[i]["runningSum"] = F(0)
        if talbot and taps[i]["runningSum"] > 0:
            p_sum, p_qi, k = F(0), F(0), F(0)
            for j in range(2):
                cm = F(min(max_prev_m, taps[j]["M"]))
                k = F(k + cm)
                if j == 0:
                    p_qi = taps[i]["p_y"]; p_sum = F(p_sum + taps[i]["p_y"] * cm)
                elif i == j:
                    p_qi = neighbor_py; p_sum = F(p_sum + neighbor_py * cm)
                else:                                             # i == 0, j == 1: the current sample seen from the previous frame's ray
                    used_depth = center_prev_depth
                    frame.origin, frame.last_frame = p_pos, True
                    try:
                        p_y = target(prev_dir(*reproj), taps[i], used_depth)
                    finally:
                        frame.origin, frame.last_frame = saved_origin, False
                    if np.isinf(p_y) or np.isnan(p_y):
                        p_y = F(0)
                    p_sum = F(p_sum + p_y * cm)
            if p_sum > 0:
                mis = F(p_qi * k / p_sum)
        taps[i]["runningSum"] = F(taps[i]["runningSum"] * mis)
        if _resample_step_max_m(taps[i], max_prev_m, output, rng) and paths:
            output["extra"] = taps[i]["extra"]
    if paths:
        return output, output.pop("extra", taps[0]["extra"])
    return output


# ---------------------------------------------------------------- multi-bounce paths (MAX_BOUNCES > 1) ----------------------------------------------------------------

def encode_wi_dist(wi, dist):
    """encodeWiDist: direction xy + distance carrying the sign of direction z."""
    return np.array([wi[0], wi[1], -dist if wi[2] < 0 else dist], dtype=F)


def decode_wi_dist(rec):
    """decodeWiDist(input, false): (direction, distance)."""
    x, y, z = F(rec[0]), F(rec[1]), F(rec[2])
    with np.errstate(invalid="ignore"):
        wz = np.sqrt(F(1) - (x * x + y * y)).astype(F)
    if np.isnan(wz):
        wz = F(0)
    if z < 0:
        wz, z = -wz, -z
    return np.array([x, y, wz], dtype=F), F(z)


def _reuse_start(P):
    """options.vertexReuseStartBounce, or None without VERTEX_REUSE (one bounce cannot be compiled with it)."""
    return P.mVertexReuseStartBounce if (P.mVertexReuse and P.mMaxBounces > 1) else None


def eval_F_path(frame, d, r, extra, final=False, no_reuse=False, spatial_reuse=False):
    """evaluate_F_ for a reservoir whose sample may be a multi-bounce path: r = the reservoir record, extra = its extra-bounce
    records (wi_dist), env-map light at the last vertex.  no_reuse (gNoReuse, final shading only): the path was drawn by
    decomposition tracking, so densities and segment transmittances cancel against its pdf and only the albedo remains per vertex.
    With vertex reuse (S = mVertexReuseStartBounce): records from bounce S on are vertices; spatial_reuse multiplies the prefix up to
    vertex S by r["p_partial"] instead of evaluating the suffix; otherwise r["p_partial"] is UPDATED (r must be a dict) to that suffix."""
    bounces = int(r["sampledPixel"]) >> 20
    depth, light_uv, light_id = F(r["depth"]), np.asarray(r["lightUV"], dtype=F), int(r["lightID"])
    S = _reuse_start(frame.P)
    if not no_reuse and (bounces == 0 or depth == K_RAY_TMAX):
        return frame.eval_F(d, depth, light_uv, light_id, final)
    vol, o = frame.grid.volume, frame.origin
    sig_s, g = np.array(vol.sigma_s[:], dtype=F), vol.PhaseFunctionConstantG
    background = depth == K_RAY_TMAX
    p = (o + d * depth).astype(F)
    if no_reuse:
        Fv = (F(1) * F(1) * (np.ones(3, F) if background else (sig_s / F(vol.sigma_t)).astype(F))).astype(F)
    else:
        density = frame.wit(0).density_world(p)
        if density == 0:
            return np.zeros(3, F)
        Fv = (frame._transmittance(final, "camera", o, d, float(depth)) * density * sig_s).astype(F)
    if not bool(np.any(Fv > 0)):
        return Fv
    if background:
        return (Fv * _env(frame, d)).astype(F)
    emissive_path = (int(r["sampledPixel"]) >> 16) & 0xF == 1
    if no_reuse and bounces == 0 and light_id == SELF_EMISSION:
        raise NotImplementedError
    sig_a = np.array(vol.sigma_a[:], dtype=F)
    wo = -d
    prefix = np.ones(3, F)
    for b in range(bounces):
        emissive_vertex = emissive_path and b == bounces - 1
        vertex = None
        if S is not None and b + 1 >= S:                   # decodeWiDist(record, reuseAsVertex): a world-space vertex, or a miss
            if extra[b][0] == K_RAY_TMAX:
                return np.zeros(3, F)
            vertex = np.array(extra[b], dtype=F)
        else:
            wi, dist = decode_wi_dist(extra[b])
            if emissive_vertex:                            # the vertex itself is stored, in (lightID, lightUV)
                vertex = decode_emissive_position(light_id, light_uv)
            elif dist == K_RAY_TMAX:
                return np.zeros(3, F)
        if vertex is not None:
            disp = (vertex - p).astype(F)
            dist = np.sqrt(F(np.dot(disp, disp))).astype(F)
            wi = (disp / dist).astype(F)
        Fv = (Fv * F(lw.phase_hg(float(np.dot(wo, wi)), g))).astype(F)
        if bool(np.all(Fv == 0)):
            return np.zeros(3, F)
        if S is not None and b == S:
            if spatial_reuse:
                return (Fv * F(r["p_partial"])).astype(F)
            prefix = Fv
        origin = p
        p = vertex if vertex is not None else (origin + wi * dist).astype(F)
        sig = sig_a if emissive_vertex else sig_s
        if no_reuse:
            Fv = (Fv * (F(1) * (sig / F(vol.sigma_t)).astype(F))).astype(F)
        else:
            scatter_density = max(F(0), frame.wit(0).density_world(p))
            Fv = (Fv * (scatter_density * sig)).astype(F)
        if (S is not None and b + 1 == S) or (emissive_vertex and (S is None or b + 1 < S)):
            Fv = (Fv * (F(1) / F(dist * dist))).astype(F)
        if bool(np.all(Fv == 0)):
            return np.zeros(3, F)
        if not no_reuse:
            Fv = (Fv * frame._transmittance(final, "camera", origin, wi, float(dist))).astype(F)
        wo = -wi
        if bool(np.all(Fv == 0)):
            return np.zeros(3, F)
    if emissive_path:
        Fv = (Fv * emission_world(frame, p)).astype(F)
    elif bool(np.any(Fv > 0)):
        at_reuse_vertex = S is not None and bounces == S
        report = {}
        Fv = (Fv * eval_L_in_volume(frame, frame.lights, p, wo, light_id, light_uv, final,
                                    tr_reuse=F(r["p_partial"]) if (at_reuse_vertex and spatial_reuse) else None, report=report)).astype(F)
        if at_reuse_vertex and not spatial_reuse and "Tr" in report:
            r["p_partial"] = report["Tr"]
    if S is not None and bounces > S and not spatial_reuse:
        with np.errstate(divide="ignore", invalid="ignore"):
            r["p_partial"] = F(lw.luminance((Fv / prefix).astype(F)))
    return Fv


def _sample_direct_lighting(frame, p, wo, rng, mips):
    """SampleDirectLighting with only an env-map light: (Ld, light pdf, lightID, lightUV)."""
    if frame.lights is not None:
        return sample_direct_lighting(frame, frame.lights, p, wo, rng, mips)
    P, g = frame.P, frame.grid.volume.PhaseFunctionConstantG
    rng.next1d()                                           # light-type selection: env lights only
    u0 = rng.next1d(); u1 = rng.next1d()
    wi, pdf, _ = lw.env_sample(mips, u0, u1)
    pdf = F(F(1) * F(pdf))
    Le = _env(frame, wi)
    Li = (Le / pdf).astype(F) if pdf > 0 else np.zeros(3, F)
    light_id, light_uv = (-2 if wi[2] < 0 else -1), np.array([wi[0], wi[1]], dtype=F)
    if bool(np.any(np.isnan(wi))):
        return np.zeros(3, F), F(0), light_id, light_uv
    if P.mInitialLightSamples != 0:
        vis = frame.compute_visibility(p, wi, float(K_RAY_TMAX), P.mInitialLightSamples, P.mInitialLightingMipLevel, bool(P.mInitialLightingUseLinearSampler),
                                       P.mInitialLightingTrackingMethod, P.mInitialLightingTStepScale, rng)
        Li = (Li * vis).astype(F)
    return (F(lw.phase_hg(float(np.dot(wo, wi)), g)) * Li / F(1)).astype(F), pdf, light_id, light_uv


def _initial_path(frame, d, hd, pd, tr, rng, mips, no_reuse=False):
    """ComputeInitialSample with MAX_BOUNCES > 1: one candidate PATH — a light sample at every vertex, streamed into one reservoir per
    path.  Returns (combined reservoir, extra-bounce records written so far)."""
    P, vol = frame.P, frame.grid.volume
    B = P.mMaxBounces
    sig_s = np.array(vol.sigma_s[:], dtype=F)
    albedo = (sig_s / F(vol.sigma_t)).astype(F)
    linear = bool(P.mInitialVisibilityUseLinearSampler)
    vis_mip = P.mInitialBaseMipLevel if linear else P.mInitialBaseMipLevel + 8
    path_pdf, path_phat = F(1), F(1)
    origin, direction = frame.origin, d
    combined = _new_reservoir()
    extra = np.zeros((max(B - 1, 1), 3), F)
    primary_depth = F(0)
    for bounce in range(B):
        out = _new_reservoir(); out["M"] = F(1)
        p = None
        if no_reuse:                                       # free flight by decomposition tracking on mip 0, pdf and Tr folded into 1
            t = frame.wit(0).sample_supervoxel(origin, direction, rng)
            pdf_dist, Tr = F(1), F(1)
            if t is None:
                cur = K_RAY_TMAX
            else:
                p = (origin + F(t) * direction).astype(F)
                cur = F(np.sqrt(np.dot(p - origin, p - origin)))
        elif bounce >= 1:
            mip = vis_mip
            if P.mInitialUseCoarserGridForIndirectBounce:
                mip = min((8 if vis_mip >= 8 else 0) + vol.numMips - 1, vis_mip + 1)
            h, q, t = frame.wit(mip).sample_distances(origin, direction, 1, linear, rng)
            cur, pdf_dist, Tr = F(h[0]), F(q[0]), F(t[0])
        else:
            cur, pdf_dist, Tr = hd, pd, tr
        valid = cur != K_RAY_TMAX
        if p is None:
            p = (origin + direction * cur).astype(F)
        path_pdf = F(path_pdf * pdf_dist)
        S = _reuse_start(P)
        if S is not None and bounce == S and valid:        # area measure at the reuse vertex
            path_pdf = F(path_pdf / F(cur * cur)); path_phat = F(path_phat / F(cur * cur))
        if bounce == 0:
            out["depth"] = cur if valid else K_RAY_TMAX
            primary_depth = out["depth"]
        else:
            out["depth"] = primary_depth
            out["sampledPixel"] = bounce << 20
            if S is None or bounce < S:
                extra[bounce - 1] = encode_wi_dist(direction, cur if valid else K_RAY_TMAX)
            else:
                extra[bounce - 1] = p if valid else np.full(3, K_RAY_TMAX, F)
        out["p_y"] = path_pdf
        density = (F(1) if no_reuse else frame.wit(0).density_world(p)) if valid else F(0)
        hit_empty = False
        if (not valid and bounce > 0) or (valid and density == 0):
            out["p_y"] = F(0); out["runningSum"] = F(0); hit_empty = True
        elif valid:
            wo = -direction
            Ld, light_pdf, out["lightID"], out["lightUV"] = _sample_direct_lighting(frame, p, wo, rng, mips)
            wi, pdf_dir = None, F(1)
            if B > 1:
                wi, pdf_dir = lw.sample_phase(vol.PhaseFunctionConstantG, wo, rng.next1d(), rng.next1d())
                pdf_dir = F(pdf_dir)
            Le = emission_world(frame, p) if (vol.hasEmission and density > 0) else np.zeros(3, F)
            lum_e = lw.luminance(((F(1) - albedo) * Le).astype(F))
            with np.errstate(divide="ignore", invalid="ignore"):
                ratio = F(lum_e / (lum_e + lw.luminance((albedo * Ld).astype(F))))
            if np.isnan(ratio):
                ratio = F(0)
            if rng.next1d() < ratio:                       # emission vs in-scattering: an RIS step of its own
                p_src = F(out["p_y"] * ratio); out["lightID"] = SELF_EMISSION
            else:
                p_src = F(out["p_y"] * (light_pdf * (F(1) - ratio)))
            out["runningSum"] = F(0) if p_src == 0 else F(1)
            out["p_y"] = p_src
            path_phat = F(F(path_phat * Tr) * density)
            if out["lightID"] == SELF_EMISSION:
                p_y = F(path_phat * lw.luminance((np.array(vol.sigma_a[:], dtype=F) * Le).astype(F)))
            else:
                p_y = F(path_phat * lw.luminance(((sig_s * Ld) * light_pdf).astype(F)))
            path_phat = F(path_phat * F(lw.luminance(sig_s) * pdf_dir))
            if no_reuse:
                p_y = F(p_y / F(vol.sigma_t)); path_phat = F(path_phat / F(vol.sigma_t))
            if out["runningSum"] > 0:
                out["runningSum"] = F(0) if out["p_y"] == 0 else F(p_y / out["p_y"])
                if out["lightID"] == SELF_EMISSION and bounce > 0:   # an emissive scatter vertex is stored as a position: area measure
                    if S is None or bounce < S:
                        out["lightID"], out["lightUV"] = encode_emissive_position(p)
                        p_y = F(p_y / F(cur * cur))
                    out["sampledPixel"] = (1 << 16) | (out["sampledPixel"] & ~0xF0000)
                out["p_y"] = p_y
            path_pdf = F(path_pdf * pdf_dir)
            if bounce < B - 1:
                origin, direction = p, wi
                if P.mInitialUseRussianRoulette and bounce >= 2:
                    if rng.next1d() < albedo[0]:
                        path_pdf = F(path_pdf * albedo[0])
                    else:
                        hit_empty = True; combined["M"] = F(combined["M"] + 1)
        else:                                              # the camera ray left the volume: the background is the sample
            Le = _env(frame, direction)
            path_phat = F(path_phat * Tr)
            p_y = F(path_phat * lw.luminance(Le))
            out["runningSum"] = F(0) if out["p_y"] == 0 else F(p_y / out["p_y"])
            out["p_y"] = p_y
            hit_empty = True
        if B == 1:
            return out, extra
        _resample_step(out, combined, rng)
        if hit_empty:
            break
    combined["M"] = F(1)
    return combined, extra


def initial_sampling_pixel_paths(frame, px, py, frame_count, importance_mips, info=None):
    """TraceRays.cs.slang main() with MAX_BOUNCES > 1, or with both reuse passes off (gNoReuse: every candidate is a path drawn by
    decomposition tracking): (stored reservoir, its extra-bounce records)."""
    P = frame.P
    no_reuse = not P.mEnableSpatialReuse and not P.mEnableTemporalReuse
    total_rounds = (P.mSpatialReuseRounds if P.mEnableSpatialReuse else 0) + int(bool(P.mEnableTemporalReuse)) + 1 + 1
    rng = Xoshiro(px, py, total_rounds * frame_count)
    frame.stage_rng = rng
    d = frame.ray_dir(px, py)
    final, final_extra = _new_reservoir(), np.zeros((max(P.mMaxBounces - 1, 1), 3), F)
    linear = bool(P.mInitialVisibilityUseLinearSampler)
    vis_mip = P.mInitialBaseMipLevel if linear else P.mInitialBaseMipLevel + 8
    rounds = (P.mInitialM + 3) // 4
    for r in range(rounds):
        n = P.mInitialM - 4 * (rounds - 1) if r == rounds - 1 else 4
        hd, pd, ot = ([0] * 4,) * 3 if no_reuse else frame.wit(vis_mip).sample_distances(frame.origin, d, n, linear, rng)
        for s in range(n):
            cand, extra = _initial_path(frame, d, F(hd[s]), F(pd[s]), F(ot[s]), rng, importance_mips, no_reuse)
            if _resample_step(cand, final, rng):
                k = int(final["sampledPixel"]) >> 20
                final_extra[:k] = extra[:k]
    if info is not None:
        info["generator_after_candidates"] = list(rng.s)
    p_hat = F(lw.luminance(eval_F_path(frame, d, final, final_extra)))
    if info is not None:
        info["generator_after_p_hat"] = list(rng.s)
    if final["runningSum"] > 0:
        final["runningSum"] = F(final["runningSum"] * (F(0) if final["p_y"] == 0 else F(p_hat / final["p_y"])))
        final["p_y"] = p_hat
    return final, final_extra


def final_shading_path(frame, px, py, r, extra, frame_count=None):
    """FinalShading.cs.slang for a multi-bounce reservoir."""
    if frame_count is not None:
        frame.seed_final(px, py, frame_count)
    if not r["runningSum"] > 0:
        return np.zeros(3, F)
    no_reuse = not frame.P.mEnableSpatialReuse and not frame.P.mEnableTemporalReuse
    col = eval_F_path(frame, frame.ray_dir(px, py), r, extra, final=True, no_reuse=no_reuse)
    W = F(1) if r["p_y"] == 0 else F(F(r["runningSum"]) / F(F(r["p_y"]) * F(r["M"])))
    out = (col * W).astype(F)
    return np.zeros(3, F) if bool(np.any(np.isnan(out) | np.isinf(out))) else out


# ---------------------------------------------------------------- analytic and emissive lights ----------------------------------------------------------------
FLT_MIN = F(1.17549435e-38)


def compute_ray_origin(pos, normal):
    """computeRayOrigin: integer offset of the fp32 bit pattern along the normal (fixed offset close to the origin)."""
    pos, normal = np.asarray(pos, dtype=F), np.asarray(normal, dtype=F)
    i_off = np.trunc(normal * F(256)).astype(np.int32)
    bits = pos.view(np.int32) + np.where(pos < 0, -i_off, i_off).astype(np.int32)
    i_pos = bits.astype(np.int32).view(F)
    f_off = (normal * F(1.0 / 65536.0)).astype(F)
    return np.where(np.abs(pos) < F(1.0 / 32.0), pos + f_off, i_pos).astype(F)


class Lights:
    """The scene's light lists as the witness needs them: analytic lights, emissive triangles and their alias table
    (items = [threshold bits, indexA, indexB], weights, weight sum — built by oracle/alias_oracle.py from AliasTable.cpp)."""

    def __init__(self, scene, emissive_alias=None):
        self.analytic = list(scene.lights)
        self.tris = scene.emissiveTriangles
        self.mult = F(scene.emissiveIntensityMultiplier)
        if emissive_alias is not None:
            self.items, self.weights, self.weight_sum = emissive_alias[0], np.asarray(emissive_alias[1], dtype=F), F(emissive_alias[2])

    def sample_triangle(self, p, tri_index, u):
        """sampleTriangle: dict(valid, dir, distance, posW, normal, Le, pdf, pdfArea, cos) for the shading point p."""
        t = self.tris[tri_index]
        su = np.sqrt(F(u[0])).astype(F)
        b = (F(1) - su, F(u[1]) * su)
        bary = (F(F(1) - b[0] - b[1]), b[0], b[1])
        v = [np.array(t.posW[k][:], dtype=F) for k in range(3)]
        normal = np.array(t.normal[:], dtype=F)
        pos = compute_ray_origin((v[0] * bary[0] + v[1] * bary[1] + v[2] * bary[2]).astype(F), normal)
        to_light = (pos - p).astype(F)
        dist_sqr = max(FLT_MIN, F(np.dot(to_light, to_light)))
        dist = np.sqrt(dist_sqr).astype(F)
        direction = (to_light / dist).astype(F)
        cos = F(np.dot(normal, -direction))
        out = dict(valid=bool(cos > 0), dir=direction, distance=dist, posW=pos, normal=normal, cos=cos, pdf=F(0), pdfArea=F(0), Le=np.zeros(3, F))
        if out["valid"]:
            out["Le"] = np.array(t.Le[:], dtype=F)
            out["pdf"] = F(dist_sqr / max(FLT_MIN, F(cos * F(t.area))))
            out["pdfArea"] = F(F(1) / F(t.area))
        return out


def sample_scene_lights(frame, lights, p, rng, mips):
    """sampleSceneLights: dict(valid, dir, rayDir, rayDistance, Li (pre-divided by pdf), pdfArea, lightID, lightUV)."""
    P = frame.P
    sel = [F(1) if P.mUseEnvironmentLights else F(0), F(1) if P.mUseAnalyticLights else F(0), F(1) if P.mUseEmissiveLights else F(0)]
    total = F(sel[0] + sel[1] + sel[2])
    invalid = dict(valid=False, pdfArea=F(0), lightID=-1, lightUV=np.zeros(2, F), Li=np.zeros(3, F))
    if total == 0:
        return invalid
    inv = F(1) / total
    sel = [F(x * inv) for x in sel]
    u = rng.next1d()
    if P.mUseEnvironmentLights:
        if u < sel[0]:
            u0 = rng.next1d(); u1 = rng.next1d()
            wi, pdf, _ = lw.env_sample(mips, u0, u1)
            pdf = F(sel[0] * F(pdf))
            Le = _env(frame, wi)
            return dict(valid=not bool(np.any(np.isnan(wi))), dir=wi, rayDir=wi, rayDistance=K_RAY_TMAX, pdfArea=pdf,
                        Li=(Le / pdf).astype(F) if pdf > 0 else np.zeros(3, F), lightID=-2 if wi[2] < 0 else -1, lightUV=np.array([wi[0], wi[1]], dtype=F))
        u = F(u - sel[0])
    if P.mUseAnalyticLights:
        if u < sel[1]:
            u = F(u / sel[1])
            count = len(lights.analytic)
            idx = min(int(np.trunc(F(u * F(count)))), count - 1)
            direction, distance, Li = analytic_light_sample(lights.analytic[idx], p)
            pdf = F(F(sel[1] / F(count)) * F(1))
            return dict(valid=True, dir=direction, rayDir=direction, rayDistance=distance, pdfArea=pdf, Li=(Li / pdf).astype(F), lightID=idx, lightUV=np.zeros(2, F))
        u = F(u - sel[1])
    if P.mUseEmissiveLights and u < sel[2]:
        if lights.tris is None or len(lights.tris) == 0:
            return invalid
        x, y = rng.next1d(), rng.next1d()
        count = len(lights.items)
        slot = min(count - 1, int(np.trunc(F(x * F(count)))))
        thr = lights.items[slot, 0:1].copy().view(F)[0]
        tri = int(lights.items[slot, 1] if y >= thr else lights.items[slot, 2])
        sel_pdf = F(lights.weights[tri] / lights.weight_sum)
        uv = np.array([rng.next1d(), rng.next1d()], dtype=F)
        ls = lights.sample_triangle(p, tri, uv)
        light_id = tri + len(lights.analytic)
        if not ls["valid"]:
            return dict(valid=False, pdfArea=F(0), lightID=light_id, lightUV=uv, Li=np.zeros(3, F))
        pdf = F(sel[2] * F(ls["pdf"] * sel_pdf))
        pdf_area = F(sel[2] * F(ls["pdfArea"] * sel_pdf))
        to_light = (compute_ray_origin(ls["posW"], ls["normal"]) - p).astype(F)
        ray_dist = np.sqrt(F(np.dot(to_light, to_light))).astype(F)
        return dict(valid=True, dir=ls["dir"], rayDir=(to_light / ray_dist).astype(F), rayDistance=ray_dist, pdfArea=pdf_area,
                    Li=((ls["Le"] * lights.mult) / pdf).astype(F) if pdf > 0 else np.zeros(3, F), lightID=light_id, lightUV=uv)
    return invalid


def analytic_light_sample(light, p):
    """samplePointLight / sampleDirectionalLight: (direction, distance, Li)."""
    if light["type"] == 1:
        return (-np.array(light["dirW"], dtype=F)).astype(F), K_RAY_TMAX, np.array(light["intensity"], dtype=F)
    to_light = (np.array(light["posW"], dtype=F) - p).astype(F)
    dist_sqr = max(F(np.dot(to_light, to_light)), F(1e-9))
    dist = np.sqrt(dist_sqr).astype(F)
    return (to_light / dist).astype(F), dist, (np.array(light["intensity"], dtype=F) / dist_sqr).astype(F)


def sample_direct_lighting(frame, lights, p, wo, rng, mips):
    """SampleDirectLighting under the initial options: (Ld, pdf, lightID, lightUV)."""
    P, g = frame.P, frame.grid.volume.PhaseFunctionConstantG
    ls = sample_scene_lights(frame, lights, p, rng, mips)
    if not ls["valid"]:
        return np.zeros(3, F), F(0), ls["lightID"], ls["lightUV"]
    Li = ls["Li"]
    if P.mInitialLightSamples != 0:
        vis = frame.compute_visibility(p, ls["rayDir"], float(ls["rayDistance"]), P.mInitialLightSamples, P.mInitialLightingMipLevel, bool(P.mInitialLightingUseLinearSampler),
                                       P.mInitialLightingTrackingMethod, P.mInitialLightingTStepScale, rng)
        Li = (Li * vis).astype(F)
    return (F(lw.phase_hg(float(np.dot(wo, ls["dir"])), g)) * Li / F(1)).astype(F), ls["pdfArea"], ls["lightID"], ls["lightUV"]


def eval_L_in_volume(frame, lights, p, wo, light_id, light_uv, final=False, tr_reuse=None, report=None):
    """evaluate_L_in_volume: transmittance to the stored light sample times its radiance times the phase function, float3.
    tr_reuse: lightVisibilityReuse (the stored transmittance replaces the march); report["Tr"] receives the transmittance used."""
    g = frame.grid.volume.PhaseFunctionConstantG
    if light_id < 0:
        zz = F(1) - light_uv[0] * light_uv[0] - light_uv[1] * light_uv[1]
        z = np.sqrt(zz).astype(F) if zz >= 0 else F(0)
        ray_dir = np.array([light_uv[0], light_uv[1], -z if light_id == -2 else z], dtype=F)
        ray_dist = K_RAY_TMAX
        Ld = _env(frame, ray_dir) * F(lw.phase_hg(float(np.dot(wo, ray_dir)), g))
    elif light_id < len(lights.analytic):
        ray_dir, ray_dist, Li = analytic_light_sample(lights.analytic[light_id], p)
        Ld = (Li * F(lw.phase_hg(float(np.dot(wo, ray_dir)), g))).astype(F)
    else:
        ls = lights.sample_triangle(p, light_id - len(lights.analytic), light_uv)
        if not ls["valid"]:
            return np.zeros(3, F)
        ray_dir, ray_dist = ls["dir"], ls["distance"]
        Ld = ((((ls["Le"] * lights.mult) * F(lw.phase_hg(float(np.dot(wo, ray_dir)), g))) * ls["cos"]) / F(ray_dist * ray_dist)).astype(F)
    tr = frame._transmittance(final, "light", p, ray_dir, float(ray_dist)) if tr_reuse is None else F(tr_reuse)
    if report is not None:
        report["Tr"] = tr
    return (tr * Ld).astype(F)


# ---------------------------------------------------------------- volume emission ----------------------------------------------------------------

def emission_world(frame, p):
    """EmissionWorldSpace: black-body colour of the temperature at p (trilinear point query of the temperature grid, linear lookup
    in the 128-texel table with border 0), times LeScale."""
    vol = frame.grid.volume
    use_prev = frame.last_frame and frame.prev_grid is not None and bool(frame.P.mUsePrevVolumeForReproj)
    if not (frame.prev_grid.volume.hasEmission if use_prev else vol.hasEmission):
        return np.zeros(3, F)
    temp = frame.wit(16).value_world(p)
    temp = min(F(6400), F(F(temp - F(vol.temperatureCutOff)) * F(vol.temperatureScale)))
    q = F(F(temp - F(25)) / F(6400))
    lut = np.ctypeslib.as_array(frame.grid.blackbody_lut, shape=(128, 4))
    x = F(q * F(128) - F(0.5))
    x0 = np.floor(x)
    fx, i0 = F(x - x0), int(x0)
    tex = lambda i: np.zeros(3, F) if i < 0 or i > 127 else lut[i, :3].astype(F)
    a, b = tex(i0), tex(i0 + 1)
    rgb = np.array([F(np.float64(fx) * np.float64(F(b[k] - a[k])) + np.float64(a[k])) for k in range(3)], dtype=F)
    return (F(vol.LeScale) * rgb).astype(F)


def encode_emissive_position(pos):
    """encodeEmissivePosition: (lightID = bits of z, lightUV = xy)."""
    return int(np.array([pos[2]], dtype=F).view(np.int32)[0]), np.array([pos[0], pos[1]], dtype=F)


def decode_emissive_position(light_id, light_uv):
    return np.array([light_uv[0], light_uv[1], np.array([light_id], dtype=np.int32).view(F)[0]], dtype=F)


# ---------------------------------------------------------------- the reference path tracer (mUseReference) ----------------------------------------------------------------

def path_trace_pixel(frame, px, py, frame_count, importance_mips):
    """TraceRays.cs.slang with gUseReference: gBaselineSamplePerPixel paths by decomposition tracking on mip 0 with next-event
    estimation at every vertex (residual ratio tracking for the shadow rays), emission collected along the way, Russian roulette
    from the third vertex on.  Returns the pixel's radiance (float3)."""
    P, vol = frame.P, frame.grid.volume
    lights = frame.lights if frame.lights is not None else Lights(frame.sc)
    total_rounds = (P.mSpatialReuseRounds if P.mEnableSpatialReuse else 0) + int(bool(P.mEnableTemporalReuse)) + 1 + 1
    rng = Xoshiro(px, py, total_rounds * frame_count)
    g = vol.PhaseFunctionConstantG
    albedo = (np.array(vol.sigma_s[:], dtype=F) / F(vol.sigma_t)).astype(F)
    n_light = max(1, P.mInitialLightSamples)
    avg = np.zeros(3, F)
    for _ in range(P.mBaselineSamplePerPixel):
        origin, direction = frame.origin, frame.ray_dir(px, py)
        beta, L = np.ones(3, F), np.zeros(3, F)
        bounce, B = 0, P.mMaxBounces
        while bounce < B:
            t = frame.wit(0).sample_supervoxel(origin, direction, rng)
            if t is not None:
                p = (origin + direction * F(t)).astype(F)
                L = (L + (emission_world(frame, p) * (F(1) - albedo)) * beta).astype(F)
                beta = (beta * albedo).astype(F)
                Ld = np.zeros(3, F)
                for _s in range(n_light):
                    ls = sample_scene_lights(frame, lights, p, rng, importance_mips)
                    if not ls["valid"]:
                        continue
                    vis = F(frame.wit(0).residual_ratio_tracking(p, ls["rayDir"], float(ls["rayDistance"]), rng))
                    Li = (ls["Li"] * vis).astype(F)
                    Ld = (Ld + (F(lw.phase_hg(float(np.dot(-direction, ls["dir"])), g)) * Li) / F(1)).astype(F)
                Ld = (Ld / F(n_light)).astype(F)
                L = (L + beta * Ld).astype(F)
                wi = None
                if B > 1:
                    wi, _pdf = lw.sample_phase(g, -direction, rng.next1d(), rng.next1d())
                if bounce < B - 1:
                    origin, direction = p, wi
                    if P.mInitialUseRussianRoulette and bounce >= 2:
                        if rng.next1d() < albedo[0]:
                            beta = (beta / albedo[0]).astype(F)
                        else:
                            bounce = B
            else:
                if bounce == 0:
                    L = (L + beta * _env(frame, direction)).astype(F)
                bounce = B
            bounce += 1
        avg = (avg + L).astype(F)
    return (avg / F(P.mBaselineSamplePerPixel)).astype(F)
