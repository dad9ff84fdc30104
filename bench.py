#!/usr/bin/env python
"""bench.py — ms/frame of the VolumetricReSTIR hot path on B200 (BASELINE.json metric), one JSON line.

A "step" is one frame of the pass (K0 features, K1 initial RIS, K2 temporal reuse, K3 spatial reuse, [K4 history],
K5 final shading) over one synthetic scene.  --config selects the BASELINE.json configuration (default 2, the one the metric
is quoted on; 1 = the CPU-runnable parity case, 3 = animated plume, 4 = dense 4-bounce cloud with 10k emissive triangles at 4K,
5 = ~2048^3 grid built on the device, larger than L2, at 4K).

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (row-sharded frame)
  python bench.py --impl reference ...      (the CPU oracle = the only runnable implementation of the reference's logic)

What the line reports:
  value              device-timed ms/frame, frames pipelined (K0/K1 of frame f+1 next to K2..K5 of frame f), camera on a slow
                     orbit that is announced one frame ahead (vrestir_set_next_camera) — every frame's stages run exactly once
  value_serial       the same frames without pipelining (one dependent chain per frame); config.stage_ms are its stage times
  value_static       pipelined, static camera (the round-1 headline, for continuity)
  value_streamed     config 3 only: every step uploads its volume from host memory (value binds frames resident on the device)
  e2e                the same metric through the C ABI's host-buffer call (vrestir_execute_host_async: camera + scene constants
                     up, the rendered frame down into pinned host memory, every step, inside the timed region)
  parity             relMSE and flip fraction of a frame WITH history against the CPU oracle on crops of the same frame
  roofline           the march engine in the spatial-reuse round: algorithmic bytes / launch time against the L2 or HBM peak
  aux_4k             the 3840x2160 frame of the same scene (north-star scaling target), same protocol as `value`
"""
import argparse
import copy
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np

CONFIGS = {
    1: dict(name="config 1: 64^3 sphere x fBm, one directional light, initial RIS only (no reuse)", kind="sphere", dim=[64, 64, 64], width=256, height=256,
            mips=3, bounces=1, density_scale=0.03, voxel=1.0),
    2: dict(name="config 2: bunny-cloud fBm sparse grid, env-map lighting, temporal + spatial reuse", kind="bunny", dim=[577, 572, 438], width=1920, height=1080,
            mips=4, bounces=1, density_scale=1.0, voxel=0.05),
    3: dict(name="config 3: plume-like animated sequence (temperature emission, velocity reprojection, volume advances every frame)", kind="plume",
            dim=[200, 300, 200], width=1920, height=1080, mips=4, bounces=1, density_scale=0.1, voxel=0.1),
    4: dict(name="config 4: dense cloud, 4 bounces, 10k emissive triangles + env map", kind="cloud", dim=[512, 512, 512], width=3840, height=2160,
            mips=4, bounces=4, density_scale=1.0, voxel=0.05),
    5: dict(name="config 5: ~2048^3 thin-shell grid built on the device (larger than L2)", kind="shells", dim=[2048, 2048, 2048], width=3840, height=2160,
            mips=4, bounces=1, density_scale=1.0, voxel=0.0125),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--dim", type=int, nargs=3, default=None)
    ap.add_argument("--kind", default=None)
    ap.add_argument("--mips", type=int, default=None)
    ap.add_argument("--bounces", type=int, default=None)
    ap.add_argument("--camera", default="orbit", choices=["orbit", "static"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--no-aux", action="store_true", help="skip the 3840x2160 frame of the same scene (aux_4k)")
    ap.add_argument("--no-extras", action="store_true", help="skip value_static")
    ap.add_argument("--cpu-tiles", type=int, default=15)
    ap.add_argument("--no-pipeline", action="store_true", help="do not overlap K0/K1 of frame f+1 with K2..K5 of frame f")
    ap.add_argument("--pipeline-level", type=int, default=2, help="1: K0/K1 of the next frame run ahead; 2: K5 additionally deferred to a third stream")
    ap.add_argument("--per-pixel", action="store_true", help="force the per-pixel kernels (A/B against the task-stream path)")
    ap.add_argument("--set", action="append", default=[], metavar="KEY=VALUE", help="extra pass-dictionary entry (A/B of a tuning switch, e.g. mPrimaryDistanceEngine=0)")
    a = ap.parse_args()
    c = CONFIGS[a.config]
    for k in ("width", "height", "dim", "kind", "mips", "bounces"):
        if getattr(a, k) is None:
            setattr(a, k, copy.copy(c[k]))
    return a


# ------------------------------------------------------------------------------------------------ scenes
PLUME_T0 = 15.0     # config 3: the plume has risen through the whole grid (3.7 k of 23.7 k bricks)


def build_scene(args, device=0, frame_time=None):
    """The synthetic scene of the selected configuration.  Config 5 is generated and mip-mapped on the device."""
    from volumetricrestirrelease_b200 import Scene
    cfg = CONFIGS[getattr(args, "config", 2)]
    sc = Scene()
    n = getattr(args, "config", 2)
    if n == 1:
        sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile=args.kind, numMips=args.mips, densityScale=cfg["density_scale"],
                         dim=tuple(args.dim), seed=1, voxelSize=cfg["voxel"])
        sc.addDirectionalLight((-1, -1, -0.5), (5, 5, 5))
        sc.frame_camera(1.1)
        return sc
    if n == 3:
        sc.addGVDBVolume(sigma_a=(6, 6, 6), sigma_s=(14, 14, 14), g=0.0, dataFile=args.kind, numMips=args.mips, densityScale=cfg["density_scale"],
                         hasVelocity=True, hasEmission=True, LeScale=0.01, temperatureCutoff=1.0, temperatureScale=100.0,
                         dim=tuple(args.dim), seed=3, voxelSize=cfg["voxel"], frameTime=PLUME_T0 if frame_time is None else frame_time)
        sc.setEnvMap((2048, 1024), seed=7)
        sc.setEnvMapIntensity(0.5)
        sc.frame_camera(0.9, direction=(0.35, 0.15, 1.0))
        return sc
    if n == 5:
        sc.addGVDBVolumeDevice(device, sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile=args.kind, numMips=args.mips, densityScale=cfg["density_scale"],
                               dim=tuple(args.dim), seed=5, voxelSize=cfg["voxel"])
    else:
        sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile=args.kind, numMips=args.mips, densityScale=cfg["density_scale"],
                         dim=tuple(args.dim), seed=2 if n == 2 else 4, voxelSize=cfg["voxel"])
    sc.setEnvMap((2048, 1024), seed=7)
    sc.setEnvMapIntensity(1.5)
    if n == 4:
        lo, hi = sc.volume_bounds_world()
        sc.addEmissiveShell(10000, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
    sc.frame_camera(0.95 if n != 5 else 0.8, direction=(0.35, 0.22, 1.0))
    return sc


def make_params(args):
    from volumetricrestirrelease_b200 import VolumetricReSTIRParams
    n = getattr(args, "config", 2)
    if n == 1:
        return VolumetricReSTIRParams(mEnableTemporalReuse=0, mEnableSpatialReuse=0, mUseEnvironmentLights=0, mUseAnalyticLights=1, mInitialM=4)
    if n == 4:
        return VolumetricReSTIRParams(mMaxBounces=args.bounces, mUseEmissiveLights=1)
    return VolumetricReSTIRParams(mMaxBounces=args.bounces)


def workload(args):
    cfg = CONFIGS[args.config]
    p = make_params(args)
    reuse = ("temporal + spatial reuse (M=%d, %d round, %d taps, radius %g)" % (p.mInitialM, p.mSpatialReuseRounds, p.mSpatialSampleCount, p.mSampleRadius)
             if p.mEnableSpatialReuse else "initial RIS only (M=%d)" % p.mInitialM)
    return ("%s: synthetic %s grid %dx%dx%d (8^3 bricks, %d mips + conservative twins), %dx%d, %s, %d bounce(s)"
            % (cfg["name"], args.kind, args.dim[0], args.dim[1], args.dim[2], args.mips, args.width, args.height, reuse, args.bounces))


def orbit_positions(scene, count, step_rad=0.004):
    """Camera positions on a slow orbit around the target (about 2 pixels of image motion per frame at 1080p)."""
    p0 = np.array(scene.camera.position, dtype=np.float64)
    t = np.array(scene.camera.target, dtype=np.float64)
    d = p0 - t
    out = []
    for f in range(count):
        a = step_rad * f
        c, s = math.cos(a), math.sin(a)
        out.append(tuple(t + np.array([c * d[0] + s * d[2], d[1], -s * d[0] + c * d[2]])))
    return out


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ CPU side (checker / baseline)
def cpu_sample(args, scene, params, importance, env_alias, tiles, steps=1, warm_frames=1, emissive_alias=None):
    """Oracle on a bounded sample: `tiles` 64x64 tiles spread over the frame (K0-K2 on the tile + 10 px halo, K3-K5 on the
    tile), frame 0 to build history then `steps` timed frames.  Returns (ms_per_frame extrapolated, stage ms, counters/px)."""
    from oracle import vro
    W, H = args.width, args.height
    op = vro.OraclePass(params)
    op.setScene(scene, W, H, importance=importance, env_alias=env_alias, emissive_alias=emissive_alias)
    T, halo = 64, 10
    T = min(T, W, H)
    nx = max(1, int(round(np.sqrt(tiles * W / H))))
    ny = max(1, int(np.ceil(tiles / nx)))
    rects = []
    for j in range(ny):
        for i in range(nx):
            if len(rects) >= tiles:
                break
            cx = int((i + 0.5) / nx * W); cy = int((j + 0.5) / ny * H)
            x0 = max(0, min(W - T, cx - T // 2)); y0 = max(0, min(H - T, cy - T // 2))
            rects.append((x0, y0, x0 + T, y0 + T))
    rounds = params.mSpatialReuseRounds if params.mEnableSpatialReuse else 0
    color = np.zeros((H, W, 4), np.float32)

    def frame():
        t_stage = {}
        cnt = {}
        for stage_group, expand in (((0, 1, 2), True), (tuple([3] * rounds) + (4, 5), False)):
            for k, stage in enumerate(stage_group):
                arg = k if stage == 3 else 0
                t0 = time.perf_counter()
                op.counters(reset=True)
                for (x0, y0, x1, y1) in rects:
                    if expand:
                        op.set_crop(x0 - halo, y0 - halo, x1 + halo, y1 + halo)
                    else:
                        op.set_crop(x0, y0, x1, y1)
                    op.execute_stage(stage, arg, color)
                dt = (time.perf_counter() - t0) * 1e3
                t_stage[stage] = t_stage.get(stage, 0.0) + dt
                c = op.counters(reset=True)
                cnt[stage] = {k2: cnt.get(stage, {}).get(k2, 0) + v for k2, v in c.items()}
        op.execute_stage(6, 0, color)
        return t_stage, cnt

    # execute_stage(0) resets the frame counter on the first call only (options changed); later tiles keep it
    for _ in range(warm_frames):
        frame()
    tot = {}
    cnts = {}
    t0 = time.perf_counter()
    for _ in range(steps):
        ts, cn = frame()
        for k, v in ts.items():
            tot[k] = tot.get(k, 0.0) + v
        for k, v in cn.items():
            cnts[k] = {k2: cnts.get(k, {}).get(k2, 0) + v2 for k2, v2 in v.items()}
    wall = (time.perf_counter() - t0) * 1e3 / steps
    tile_px = len(rects) * T * T
    exp_px = sum((min(W, x1 + halo) - max(0, x0 - halo)) * (min(H, y1 + halo) - max(0, y0 - halo)) for x0, y0, x1, y1 in rects)
    scale_tile, scale_exp = W * H / tile_px, W * H / exp_px
    stage_ms = {k: v / steps * (scale_exp if k in (0, 1, 2) else scale_tile) for k, v in tot.items()}
    per_px = {k: {k2: v2 / steps / (exp_px if k in (0, 1, 2) else tile_px) for k2, v2 in v.items()} for k, v in cnts.items()}
    sample = (f"{len(rects)} tiles of {T}x{T} px (+{halo} px halo for K0-K2) of the {W}x{H} frame, {steps} frame(s) after {warm_frames} history frame(s), "
              "extrapolated by pixel count")
    return sum(stage_ms.values()), stage_ms, per_px, op.threads(), sample, wall


RES = np.dtype([("runningSum", "<f4"), ("M", "<f4"), ("depth", "<f4"), ("p_y", "<f4"), ("lightUV", "<f4", 2), ("lightID", "<i4"), ("sampledPixel", "<i4")])


def parity_with_history(gp, scene, params, W, H, color, oracle_scene=None, frame_ids=None, volumes=None):
    """relMSE / flip fraction of a frame WITH history against the oracle (BASELINE metric's second half).  The GPU renders
    frames 0 and 1 of a fresh epoch; its history is handed to the oracle; frame 2 is compared on two 64x64 crops (the oracle
    runs K0-K2 on crop + 10 px halo, K3-K5 on the crop).  relMSE = mean((a-b)^2 / (b^2 + 1e-2 mean(b)^2))."""
    import torch
    from oracle import vro
    from volumetricrestirrelease_b200 import capi
    B = params.mMaxBounces
    gp.updateDict({"mPipelineFrames": 0})
    for f in range(2):
        if frame_ids:          # animated sequence: every frame binds the next resident volume
            gp.advanceVolumeResident(frame_ids[f])
        gp.execute(color.data_ptr())
    torch.cuda.synchronize()
    op = vro.OraclePass(params)
    if frame_ids:              # the oracle renders frame 2 with volume 1 as the previous frame's grids and volume 2 as the current
        oracle_scene = copy.copy(scene)
        oracle_scene.volume = volumes[1]
    em = gp.emissive_alias(len(scene.emissiveTriangles)) if scene.emissiveTriangles is not None else None
    imp = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32) if scene.envMap is not None else None
    op.setScene(oracle_scene or scene, W, H, importance=imp, env_alias=gp.env_alias() if scene.envMap is not None else None, emissive_alias=em)
    if frame_ids:
        op.advanceVolume(volumes[2])
        gp.advanceVolumeResident(frame_ids[2])
    c0 = np.zeros((H, W, 4), np.float32)
    reuse = bool(params.mEnableTemporalReuse)
    op.execute_stage(6, 0, c0)
    op.set_frame_count(gp.frame_count(), 1)
    history = {}
    if reuse:
        for b in (capi.BUF_RESERVOIR_TEMPORAL, capi.BUF_FEATURES_TEMPORAL) + ((capi.BUF_EXTRA_TEMPORAL,) if B > 1 else ()):
            history[b] = gp.get_buffer(b).copy()
    fc = gp.frame_count()
    gp.execute(color.data_ptr()); torch.cuda.synchronize()
    g = color.cpu().numpy()
    final_buf = capi.BUF_RESERVOIR_TEMPORAL if reuse else capi.BUF_RESERVOIR_0
    gres = gp.get_buffer(final_buf).view(RES).reshape(H, W)
    T, halo = min(64, W // 2, H // 2), 10
    rounds = params.mSpatialReuseRounds if params.mEnableSpatialReuse else 0
    crops = [(W // 2 - T // 2, H // 2 - T // 2), (max(0, W // 2 - int(0.22 * W)), max(0, H // 2 - int(0.19 * H)))]
    num = den = 0.0
    gs, cs, flips, px = [], [], 0, 0
    fields = {}
    for x0, y0 in crops:
        op.set_frame_count(fc, 1)
        for b, raw in history.items():      # the oracle's history copies (K4, feature copy) are whole-buffer: hand the GPU's history over per crop
            op.set_buffer(b, raw)
        for stage in (0, 1, 2):
            op.set_crop(x0 - halo, y0 - halo, x0 + T + halo, y0 + T + halo)
            op.execute_stage(stage, 0, c0)
        for stage in [3] * rounds + [4, 5]:
            op.set_crop(x0, y0, x0 + T, y0 + T)
            op.execute_stage(stage, 0, c0)
        gs.append(g[y0:y0 + T, x0:x0 + T, :3].astype(np.float64)); cs.append(c0[y0:y0 + T, x0:x0 + T, :3].astype(np.float64))
        cres = op.get_buffer(final_buf).view(RES).reshape(H, W)[y0:y0 + T, x0:x0 + T]
        a = gres[y0:y0 + T, x0:x0 + T]
        f = (a["lightID"] != cres["lightID"]) | (a["sampledPixel"] != cres["sampledPixel"]) | (a["M"] != cres["M"])
        f |= ~np.isclose(a["depth"], cres["depth"], rtol=1e-5, atol=0) & ~(a["depth"] == cres["depth"])
        flips += int(f.sum()); px += f.size
        per_field = {k: int((a[k] != cres[k]).sum()) for k in ("lightID", "sampledPixel", "M", "depth")}
        fields = {k: fields.get(k, 0) + v for k, v in per_field.items()}
    ga, ca = np.concatenate(gs), np.concatenate(cs)
    eps = 1e-2 * np.mean(ca) ** 2
    relmse = float(np.mean((ga - ca) ** 2 / (ca ** 2 + eps)))
    rel = np.abs(ga - ca) / np.maximum(np.abs(ca), 1e-6)
    return {"relmse_vs_oracle": relmse, "flip_frac": flips / max(1, px), "pixels_with_differing_field": fields, "frac_pixels_rel_err_gt_1e-4": float((rel.max(axis=-1) > 1e-4).mean()),
            "frame": int(fc), "crops": [[x0, y0, T, T] for x0, y0 in crops], "mean_gpu": float(ga.mean()), "mean_oracle": float(ca.mean())}


def run_reference(args):
    """--impl reference: the reference's logic on the host cores (oracle port; the reference itself is D3D12-only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if args.config == 5:
        print(json.dumps({"impl": "reference", "unavailable": "config 5's grid only exists on the device; the CPU oracle runs on its download inside the ours arm (cpu_baseline)"}))
        return
    scene = build_scene(args)
    params = make_params(args)
    vals = []
    info = None
    ran = max(1, min(args.steps, 3))
    for _ in range(ran):
        ms, stage_ms, per_px, threads, sample, wall = cpu_sample(args, scene, params, None, None, args.cpu_tiles, steps=1, warm_frames=1,
                                                                 emissive_alias=_host_emissive_alias(scene))
        vals.append(ms)
        info = (threads, sample)
    v = float(np.median(vals))
    line = {"impl": "reference", "metric": "ms/frame", "value": v, "unit": "ms/frame", "n_gpus": args.gpus, "steps": ran, "steps_requested": args.steps, "warmup": 1,
            "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload(args)},
            "cpu_baseline": {"value": v, "unit": "ms/frame", "cores": info[0], "kind": "port", "sample": info[1] + f"; median of {ran} such frame(s)"},
            "e2e": {"value": v, "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def _host_emissive_alias(scene):
    """Emissive alias table through the ABI's host-only builder (no device needed)."""
    if scene.emissiveTriangles is None:
        return None
    import ctypes as C
    from volumetricrestirrelease_b200 import capi
    n = len(scene.emissiveTriangles)
    w = np.zeros(n, dtype=np.float32)
    for i, t in enumerate(scene.emissiveTriangles):
        lum = np.float32(0.2126) * np.float32(t.Le[0]) + np.float32(0.7152) * np.float32(t.Le[1]) + np.float32(0.0722) * np.float32(t.Le[2])
        w[i] = np.float32(np.float32(lum) * np.float32(t.area)) * np.float32(3.14159265358979323846)
    items = np.zeros((n, 4), dtype=np.uint32)
    ws = C.c_float()
    capi.check(capi.lib().vrestir_build_alias_table(w.ctypes.data, n, items.ctypes.data, C.byref(ws)))
    return items, w, ws.value


# ------------------------------------------------------------------------------------------------ the measured arm
class Runner:
    """One pass (or one rank's band of it) + the per-frame protocol: [advance the volume], move the camera, announce the next
    camera, execute."""

    def __init__(self, args, W, H, scene, params, level, rank, world, local, volumes=None):
        import torch
        from volumetricrestirrelease_b200 import VolumetricReSTIR
        from volumetricrestirrelease_b200.multi_gpu import ShardedPass
        self.torch = torch
        self.args, self.W, self.H, self.scene, self.world, self.rank = args, W, H, scene, world, rank
        d = {"mParams": params, "mPipelineFrames": level}
        if args.per_pixel:
            d["mUseWavefront"] = 0
        for kv in getattr(args, "set", []) or []:
            d[kv.split("=")[0]] = float(kv.split("=")[1])
        self.gp = VolumetricReSTIR.create(d, device=local)
        self.sp = ShardedPass(self.gp, W, H, rank, world, torch.device("cuda", local))
        self.gp.setScene(scene, W, H)
        self.color = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
        self.band = self.sp.balance(refine=0)
        self.gp.setRowBand(*self.band)
        self.volumes = volumes or []
        # animated sequences: every frame is uploaded once and stays resident, as in the reference (F/Scene/Scene.cpp:825-863)
        self.frame_ids = [self.gp.addVolumeFrame(v) for v in self.volumes]
        self.streamed = False          # True: every step uploads its volume from host memory instead (vrestir_advance_volume)
        self.frame = 0
        self.path = orbit_positions(scene, 4096) if args.camera == "orbit" else None
        self.cam_index = 0

    def set_level(self, level):
        self.gp.updateDict({"mPipelineFrames": level})

    def step(self, out_ptr=None, moving=True):
        gp, sc = self.gp, self.scene
        self.advance()
        if self.path is not None and moving:
            sc.camera.position = self.path[self.cam_index % len(self.path)]
            gp.updateCamera()
            nxt = copy.copy(sc.camera)
            nxt.position = self.path[(self.cam_index + 1) % len(self.path)]
            gp.setNextCamera(nxt)
            self.cam_index += 1
        self.sp.execute(out_ptr or self.color.data_ptr())
        self.frame += 1

    def advance(self):
        if not self.volumes:
            return
        k = self.frame % len(self.volumes)
        if self.streamed:
            self.gp.advanceVolume(self.volumes[k])
        else:
            self.gp.advanceVolumeResident(self.frame_ids[k])

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            self.torch.cuda.synchronize()

    def timed(self, steps, warmup, moving=True):
        torch = self.torch
        for _ in range(max(3, warmup)):
            self.step(moving=moving)
        self.barrier()
        l0 = self.gp.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            self.step(moving=moving)
        self.gp.wait_output()          # the last frame's deferred final shading belongs to the timed region
        e1.record()
        self.barrier()
        return e0.elapsed_time(e1) / steps, self.gp.launch_count() - l0

    def stage_times(self, steps, moving=True):
        acc = {}
        for _ in range(steps):
            self.step(moving=moving)
            self.torch.cuda.synchronize()
            for k, v in self.gp.timings().items():
                acc[k] = acc.get(k, 0.0) + v / steps
        return acc


def allmax(world, vals):
    if world == 1:
        return [float(v) for v in vals]
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(vals), device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t]


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from volumetricrestirrelease_b200 import capi

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = args.width, args.height
    t0 = time.time()
    scene = build_scene(args, device=local)
    params = make_params(args)
    volumes = []
    if args.config == 3:      # animated sequence: 6 prebuilt frames of the fully risen plume, cycled; every step advances the volume
        volumes = [build_scene(args, frame_time=PLUME_T0 + 0.35 * f).volume for f in range(1, 7)]
    # a new volume every frame invalidates K0/K1 computed ahead (they would be discarded): animated sequences run un-pipelined
    pipelined = not args.no_pipeline and not volumes
    level = args.pipeline_level if pipelined else 0
    R = Runner(args, W, H, scene, params, level, rank, world, local, volumes)
    gp = R.gp
    if args.config == 5:
        scene.volume.release_chain()      # the dense chain (tens of GB) is no longer needed once the brick pools are bound
    if rank == 0:
        st = [scene.volume.stats(s) if args.config != 5 else None for s in (0, 1, 9, 2)]
        print(f"[bench] scene + upload {time.time() - t0:.1f}s; bricks mip0/mip1/cons1/mip2 = {st}", file=sys.stderr)

    # ---- device-resident timing: `value` (pipelined frames, camera on its announced orbit)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches = R.timed(args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    stage_acc = R.stage_times(min(args.steps, 10))
    pstats = gp.pipeline_stats()
    static_ms = None
    streamed_ms = None
    if volumes and not args.no_extras:      # the same frames with the volume uploaded from host memory every step
        R.streamed = True
        streamed_ms, _ = R.timed(args.steps, args.warmup)
        R.streamed = False
    if pipelined and not args.no_extras and not volumes and R.path is not None:
        gp.setNextCamera(None)
        static_ms, _ = R.timed(args.steps, args.warmup, moving=False)

    # ---- the same frames without pipelining: `value_serial` + the per-stage times
    serial_ms, serial_stage, mt_alone = None, None, None
    if pipelined:
        R.set_level(0)
        serial_ms, _ = R.timed(args.steps, args.warmup)
        serial_stage = R.stage_times(min(args.steps, 10))
        try:
            mt_alone = gp.march_timings()   # the march launches timed alone on the GPU (no second chain next to them): roofline input
        except capi.VRestirError:
            mt_alone = None
        R.set_level(level)
    else:
        serial_ms, serial_stage = ms, stage_acc
        try:
            mt_alone = gp.march_timings()
        except capi.VRestirError:
            mt_alone = None

    # ---- end to end through the C ABI's host-buffer call: every step uploads its inputs (camera + scene constants, from host
    # memory), renders, and its frame lands in pinned host memory (vrestir_execute_host_async: the read-back of frame f travels
    # while frame f+1 renders; the timed region ends when the last frame has landed).  N > 1: the sharded driver renders the band
    # into a device image and torch copies the band out on a copy stream (same protocol, the ABI call is per stage there).
    hosts = [torch.zeros((H, W, 4), dtype=torch.float32).pin_memory() for _ in range(2)]
    r0, r1 = R.band
    if world == 1:
        def e2e_step(i):
            R.advance()
            if R.path is not None:
                scene.camera.position = R.path[R.cam_index % len(R.path)]
                nxt = copy.copy(scene.camera)
                nxt.position = R.path[(R.cam_index + 1) % len(R.path)]
                gp.setNextCamera(nxt)
                R.cam_index += 1
            gp.updateCamera()                                        # H2D: camera + 8 KB scene constants
            gp.execute_host_async(hosts[i & 1].data_ptr())            # render + D2H of the frame
            R.frame += 1
        for i in range(3):
            e2e_step(i)
        gp.host_wait(); R.barrier()
        t_e2e0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        gp.host_wait()
        R.barrier()
        e2e_ms = (time.perf_counter() - t_e2e0) * 1e3 / args.steps
    else:
        colors = [R.color, torch.zeros_like(R.color)]
        copy_stream = torch.cuda.Stream()
        rendered = [torch.cuda.Event(), torch.cuda.Event()]
        landed = [torch.cuda.Event(), torch.cuda.Event()]

        def e2e_step(i):
            b = i & 1
            torch.cuda.current_stream().wait_event(landed[b])
            gp.updateCamera()
            R.step(colors[b].data_ptr())
            rendered[b].record()
            copy_stream.wait_event(rendered[b])
            gp.wait_output(copy_stream.cuda_stream)
            with torch.cuda.stream(copy_stream):
                hosts[b][r0:r1].copy_(colors[b][r0:r1], non_blocking=True)
                landed[b].record()
        for ev in landed:
            ev.record()
        for i in range(2):
            e2e_step(i)
        R.barrier()
        t_e2e0 = time.perf_counter()
        for i in range(args.steps):
            e2e_step(i)
        R.barrier()
        e2e_ms = (time.perf_counter() - t_e2e0) * 1e3 / args.steps
    ms, e2e_ms, serial_ms, static_v = allmax(world, [ms, e2e_ms, serial_ms or 0.0, static_ms or 0.0])
    static_ms = static_v if static_ms is not None else None

    # ---- aux: the 3840x2160 frame of the same scene (north-star scaling target), pipelined, same camera protocol
    aux = None
    if not args.no_aux and args.config == 2 and (W, H) != (3840, 2160):
        a2 = copy.copy(args)
        a2.width, a2.height = 3840, 2160
        del R.sp, R.gp
        gp = None
        torch.cuda.empty_cache()
        R4 = Runner(a2, 3840, 2160, scene, params, level, rank, world, local)
        ms4, _ = R4.timed(max(5, args.steps // 2), 3)
        ms4, = allmax(world, [ms4])
        aux = {"ms_per_frame": ms4, "width": 3840, "height": 2160, "bands": [list(b) for b in R4.sp.bands], "steps": max(5, args.steps // 2)}
        del R4
        torch.cuda.empty_cache()
        R = Runner(args, W, H, scene, params, level, rank, world, local, volumes)   # back to the configured frame for the checker legs
        gp = R.gp
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = {}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peaks = json.load(open(pk))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    # L2 -> SM read peak (32 MiB resident buffer): not in MEASURED_PEAKS.json, measured here (profiles/r02_l2_read_peak.txt)
    l2 = capi.C.c_float(0.0)
    capi.check(capi.lib().vrestir_debug_read_bandwidth(local, 32 << 20, 50, capi.C.byref(l2)))
    l2_peak = float(l2.value)

    parity = None
    oracle_scene = None
    if world == 1 and (not args.no_parity or not args.no_cpu_baseline) and args.config == 5:
        oracle_scene = copy.copy(scene)
        oracle_scene.volume = gp.downloadVolume()      # the device-built grid, for the CPU checker
    if not args.no_parity and world == 1:
        parity = parity_with_history(gp, scene, params, W, H, R.color, oracle_scene, R.frame_ids, volumes)

    cpu, roof = None, None
    if not args.no_cpu_baseline and world == 1:
        imp = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32) if scene.envMap is not None else None
        em = gp.emissive_alias(len(scene.emissiveTriangles)) if scene.emissiveTriangles is not None else None
        cms, cstage, per_px, threads, sample, wall = cpu_sample(args, oracle_scene or scene, params, imp, gp.env_alias() if scene.envMap is not None else None,
                                                               args.cpu_tiles, emissive_alias=em)
        cpu = {"value": cms, "unit": "ms/frame", "cores": threads, "kind": "port", "sample": sample,
               "stage_ms": {str(k): round(v, 1) for k, v in cstage.items()}}
        mt = mt_alone
        if mt and params.mEnableSpatialReuse:
            # Dominant kernel: k_march, the transmittance-march engine, in its launches of the spatial-reuse round (camera stream
            # + light / scatter stream).  ALGORITHMIC bytes = what the reference's K3 needs for its p-hat evaluations:
            # stored voxel bytes per density tap + 36 B per node visit, counted by the instrumented oracle on the sample of the
            # same frame, + the task / result records the engine moves.
            c3 = per_px.get(3, {})
            alg_px = c3.get("voxel_bytes", 0) + 36.0 * c3.get("node_visits", 0)
            task_bytes = mt["spatial_cam_tasks"] * (32 + 12) + mt["spatial_light_tasks"] * (48 + 4)
            alg = alg_px * W * H / max(1, params.mSpatialReuseRounds) + task_bytes
            t3 = mt["spatial_cam_ms"] + mt["spatial_light_ms"]
            traffic = None
            tp = os.path.join(ROOT, "profiles", "traffic_k_march.json")
            if os.path.exists(tp):
                tj = json.load(open(tp))
                traffic = tj.get("config%d" % args.config, tj.get("dram_bytes_per_spatial_round") if args.config == 2 else None)
            if t3 > 0:
                ach = alg / (t3 * 1e-3) / 1e9
                # which roof: the reuse mips of configs 1-4 are L2 resident by design (DRAM traffic << algorithmic bytes, ncu);
                # a grid larger than L2 (config 5) streams from HBM
                # (measured DRAM bytes of these launches, profiles/traffic_k_march.json, against the algorithmic bytes: even the
                # 2048^3 grid of config 5, whose pools are 25x L2, is served mostly by L2 — neighbouring rays share bricks)
                l2_bound = traffic is None and args.config != 5 or (traffic is not None and traffic < 0.5 * alg)
                peak = l2_peak if l2_bound else hbm_peak
                roof = {"kernel": "k_march (march engine; camera + light launches of one spatial-reuse round)", "bound": "l2" if l2_bound else "hbm",
                        "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak if peak > 0 else None, "traffic": traffic,
                        "peak_source": ("L2 read peak measured in this run (vrestir_debug_read_bandwidth, 32 MiB resident buffer; MEASURED_PEAKS.json has no L2 figure)"
                                        if l2_bound else peak_src),
                        "hbm_peak_gbs": hbm_peak, "l2_read_peak_gbs": l2_peak, "frac_of_hbm_peak": ach / hbm_peak,
                        "dram_frac": (traffic / (t3 * 1e-3) / 1e9 / hbm_peak) if traffic else None,
                        "algorithmic_bytes_per_launch": alg, "taps_per_px": c3.get("density_taps", 0), "node_visits_per_px": c3.get("node_visits", 0),
                        "launch_ms": t3, "camera_launch_ms": mt["spatial_cam_ms"], "light_launch_ms": mt["spatial_light_ms"],
                        "camera_tasks": mt["spatial_cam_tasks"], "light_tasks": mt["spatial_light_tasks"]}
                Rb = 32 if params.mMaxBounces == 1 else 36 + 12 * (params.mMaxBounces - 1)
                res_bytes = {0: 8, 1: Rb, 2: 2 * Rb + 16 + Rb, 3: params.mSpatialSampleCount * Rb + 8 + Rb, 5: Rb + 16}
                names = {0: "features_ms", 1: "initial_ms", 2: "temporal_ms", 3: "spatial_ms", 5: "final_ms"}
                stages = {}
                for k, nm in names.items():
                    c = per_px.get(k, {})
                    b = (c.get("voxel_bytes", 0) + 36.0 * c.get("node_visits", 0) + res_bytes[k]) * W * H
                    t = serial_stage.get(nm, 0.0)
                    if t > 0:
                        stages["K%d" % k] = {"algorithmic_GB": round(b / 1e9, 3), "ms": round(t, 3), "GBps": round(b / (t * 1e-3) / 1e9, 1),
                                              "frac_hbm": round(b / (t * 1e-3) / 1e9 / hbm_peak, 3), "frac_l2": round(b / (t * 1e-3) / 1e9 / l2_peak, 3) if l2_peak > 0 else None}
                roof["stages"] = stages
    cfg = {"workload": workload(args), "parallelism": f"rows/{world}" + (f" (cost-balanced bands {R.sp.bands}, reservoir halo exchange over NCCL send/recv)" if world > 1 else ""),
           "l2": "inputs larger than L2 (fp32 mip-0 brick pool + ~0.9 GB/frame of reservoir traffic stream through every frame; the reuse mip of configs 1-4 is pinned in L2 by design)",
           "camera": "orbit of 0.004 rad/frame around the target, announced one frame ahead" if R.path is not None else "static",
           "stage_ms": {k: round(v, 3) for k, v in (serial_stage or stage_acc).items()},
           "stage_ms_note": "stage times of the un-pipelined frame (one dependent chain; in the pipelined frame the stages of two frames overlap)",
           "pipelining": ({"on": True, "level": level, "main_stream_stage_ms": {k: round(v, 3) for k, v in stage_acc.items()},
                           "prefetch_chain_ms": round(pstats["prefetch_ms"], 3), "deferred_final_ms": round(pstats["deferred_final_ms"], 3),
                           "adopted": int(pstats["adopted"]), "discarded": int(pstats["discarded"])} if pipelined else {"on": False})}
    if volumes:
        cfg["animation"] = (f"every step advances the volume: {len(volumes)} frames resident on the device, cycled (vrestir_advance_volume_resident; the reference "
                            f"keeps a sequence's frames resident too); value_streamed = the same frames with every step's grids uploaded from pageable host memory "
                            f"inside the timed region (vrestir_advance_volume)")
    line = {"metric": "ms/frame", "value": ms, "unit": "ms/frame", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "value_pipelined": ms, "value_serial": serial_ms, "value_static": static_ms, "value_streamed": streamed_ms,
            "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": "ms/frame", "h2d_bytes_per_step": int(capi.C.sizeof(capi.Camera) + 8192),
                    "d2h_bytes_per_step": int((r1 - r0) * W * 16),
                    "through": "vrestir_execute_host_async (C ABI, pinned host buffers)" if world == 1 else "sharded driver + band read-back on a copy stream"},
            "gpu_launches": int(launches), "parity": parity, "aux_4k": aux,
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
