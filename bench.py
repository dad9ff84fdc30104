#!/usr/bin/env python
"""bench.py — ms/frame of the VolumetricReSTIR hot path on B200 (BASELINE.json metric), one JSON line.

A "step" is one frame of the pass (K0 features, K1 initial RIS, K2 temporal reuse, K3 spatial reuse, [K4 history],
K5 final shading) over one synthetic scene.  Default workload = BASELINE.json configs[1]: bunny-cloud-shaped fBm sparse
grid 577x572x438 (8^3 bricks, 4 mips), 1920x1080, env-map lighting, temporal + spatial reuse, single bounce.

  python bench.py --gpus 1 --steps 20 --warmup 5
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...     (row-sharded frame)
  python bench.py --impl reference ...      (the CPU oracle = the only runnable implementation of the reference's logic)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--dim", type=int, nargs=3, default=[577, 572, 438])
    ap.add_argument("--kind", default="bunny")
    ap.add_argument("--mips", type=int, default=4)
    ap.add_argument("--bounces", type=int, default=1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-tiles", type=int, default=15)
    ap.add_argument("--stage-breakdown", action="store_true", help="also print per-stage ms to stderr")
    ap.add_argument("--no-pipeline", action="store_true", help="do not overlap K0/K1 of frame f+1 with K2..K5 of frame f")
    ap.add_argument("--pipeline-level", type=int, default=2, help="1: K0/K1 of the next frame run ahead; 2: K5 additionally deferred to a third stream")
    return ap.parse_args()


def build_scene(args):
    from volumetricrestirrelease_b200 import Scene
    sc = Scene()
    sc.addGVDBVolume(sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile=args.kind, numMips=args.mips, densityScale=1.0,
                     dim=tuple(args.dim), seed=2, voxelSize=0.05)
    sc.setEnvMap((2048, 1024), seed=7)
    sc.setEnvMapIntensity(1.5)
    sc.frame_camera(0.95, direction=(0.35, 0.22, 1.0))
    return sc


def make_params(args):
    from volumetricrestirrelease_b200 import VolumetricReSTIRParams
    return VolumetricReSTIRParams(mMaxBounces=args.bounces)


WORKLOAD = ("synthetic bunny-cloud fBm sparse grid {d[0]}x{d[1]}x{d[2]} (8^3 bricks, {m} mips + conservative twins), {w}x{h}, "
            "env-map lighting 2048x1024, temporal + spatial reuse (M=4, 1 round, 4 taps, radius 10), {b} bounce(s)")


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


def cpu_sample(args, scene, params, importance, env_alias, tiles, steps=1, warm_frames=1):
    """Oracle on a bounded sample: `tiles` 64x64 tiles spread over the frame (K0-K2 on the tile + 10 px halo, K3-K5 on the
    tile), frame 0 to build history then `steps` timed frames.  Returns (ms_per_frame extrapolated, stage ms, counters/px)."""
    from oracle import vro
    W, H = args.width, args.height
    op = vro.OraclePass(params)
    op.setScene(scene, W, H, importance=importance, env_alias=env_alias)
    T, halo = 64, 10
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

    def frame(timed):
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
        frame(False)
    tot = {}
    cnts = {}
    t0 = time.perf_counter()
    for _ in range(steps):
        ts, cn = frame(True)
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
    return sum(stage_ms.values()), stage_ms, per_px, op.threads(), f"{len(rects)} tiles of 64x64 px (+10 px halo for K0-K2) of the {W}x{H} frame, {steps} frame(s) after {warm_frames} history frame(s), extrapolated by pixel count", wall


def run_reference(args):
    """--impl reference: the reference's logic on the host cores (oracle port; the reference itself is D3D12-only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene = build_scene(args)
    params = make_params(args)
    vals = []
    info = None
    for _ in range(max(1, min(args.steps, 3))):
        ms, stage_ms, per_px, threads, sample, wall = cpu_sample(args, scene, params, None, None, args.cpu_tiles, steps=1, warm_frames=1)
        vals.append(ms)
        info = (threads, sample)
    v = float(np.median(vals))
    line = {"impl": "reference", "metric": "ms/frame", "value": v, "unit": "ms/frame", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": v, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(d=args.dim, m=args.mips, w=args.width, h=args.height, b=args.bounces)},
            "cpu_baseline": {"value": v, "unit": "ms/frame", "cores": info[0], "kind": "port", "sample": info[1]},
            "e2e": {"value": v, "unit": "ms/frame", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    from volumetricrestirrelease_b200 import VolumetricReSTIR, capi
    from volumetricrestirrelease_b200.multi_gpu import ShardedPass

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = args.width, args.height
    t0 = time.time()
    scene = build_scene(args)
    params = make_params(args)
    # frame pipelining: the camera of the bench is static, i.e. known one frame ahead; every frame still runs its own K0/K1
    # (keyed by its frame counter), just next to the previous frame's K2..K5 instead of after them
    pipelined = not args.no_pipeline
    level = args.pipeline_level if pipelined else 0
    gp = VolumetricReSTIR.create({"mParams": params, "mPipelineFrames": level}, device=local)
    sp = ShardedPass(gp, W, H, rank, world, torch.device("cuda", local))
    gp.setScene(scene, W, H)
    color = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    # world > 1: cost-balanced row bands from a full-frame K0 cost model (the measured-time refinement is off: with the fixed
    # per-launch latency of a short band it over-corrects, 7.06 vs 6.49 ms at 4K on 8 GPUs)
    r0, r1 = sp.balance(refine=0)
    gp.setRowBand(r0, r1)
    if rank == 0:
        print(f"[bench] scene + upload {time.time() - t0:.1f}s; bricks mip0={scene.volume.stats(0)} mip1={scene.volume.stats(1)} "
              f"cons1={scene.volume.stats(9)} mip2={scene.volume.stats(2)}", file=sys.stderr)
    host_color = torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing (value) ----
    for _ in range(max(3, args.warmup)):
        sp.execute(color.data_ptr())
    barrier()
    launches0 = gp.launch_count()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        sp.execute(color.data_ptr())
    gp.wait_output()          # the last frame's deferred final shading belongs to the timed region
    ev1.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    ms = ev0.elapsed_time(ev1) / args.steps
    launches = gp.launch_count() - launches0
    # per-stage timings (pass-internal CUDA events on the launching stream), averaged over `steps` more frames
    stage_acc = {}
    for _ in range(args.steps):
        sp.execute(color.data_ptr())
        torch.cuda.synchronize()
        for k, v in gp.timings().items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v / args.steps

    pstats = gp.pipeline_stats()
    serial_ms = None
    mt_alone = None
    if pipelined:   # the same frames without pipelining, for the record (option change: history restarts, so warm up again)
        gp.updateDict({"mPipelineFrames": 0})
        for _ in range(max(3, args.warmup)):
            sp.execute(color.data_ptr())
        barrier()
        ev0.record()
        for _ in range(args.steps):
            sp.execute(color.data_ptr())
        ev1.record()
        barrier()
        serial_ms = ev0.elapsed_time(ev1) / args.steps
        serial_stage = {}
        for _ in range(args.steps):
            sp.execute(color.data_ptr())
            torch.cuda.synchronize()
            for k, v in gp.timings().items():
                serial_stage[k] = serial_stage.get(k, 0.0) + v / args.steps
        mt_alone = gp.march_timings()   # the march launches timed alone on the GPU (no second chain next to them): roofline input
        gp.updateDict({"mPipelineFrames": level})
        for _ in range(max(3, args.warmup)):
            sp.execute(color.data_ptr())
        barrier()

    # ---- end to end with HOST buffers: every step uploads its inputs (camera + scene constants, from host memory) through
    # the public call, renders, and reads its frame back into pinned host memory.  The read-back of frame i runs on a copy
    # stream while frame i+1 renders (two device frames, two host frames): a renderer that streams frames out, not a
    # different metric — every step's H2D and D2H are inside the timed region and the region ends when the last frame has
    # landed in host memory.
    colors = [color, torch.zeros_like(color)]
    hosts = [host_color, torch.zeros((H, W, 4), dtype=torch.float32).pin_memory()]
    copy_stream = torch.cuda.Stream()
    rendered = [torch.cuda.Event(), torch.cuda.Event()]
    landed = [torch.cuda.Event(), torch.cuda.Event()]

    def e2e_step(i):
        b = i & 1
        torch.cuda.current_stream().wait_event(landed[b])    # the device frame is free again once its last read-back landed
        gp.updateCamera()                                    # H2D: camera + 8 KB scene constants
        sp.execute(colors[b].data_ptr())
        rendered[b].record()
        copy_stream.wait_event(rendered[b])
        gp.wait_output(copy_stream.cuda_stream)              # pipelining level 2: the image is complete when the deferred K5 is
        with torch.cuda.stream(copy_stream):
            hosts[b][r0:r1].copy_(colors[b][r0:r1], non_blocking=True)   # D2H: the band of the frame
            landed[b].record()

    for ev in landed:
        ev.record()
    for i in range(2):
        e2e_step(i)
    barrier()
    t_e2e0 = time.perf_counter()
    for i in range(args.steps):
        e2e_step(i)
    barrier()                                                # includes the copy stream: torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t_e2e0) * 1e3 / args.steps
    if world > 1:
        t = torch.tensor([ms, e2e_ms, serial_ms or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms = float(t[0]), float(t[1])
        serial_ms = float(t[2]) if pipelined else None
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

    cpu = None
    roof = None
    mt = mt_alone or gp.march_timings()
    # L2 -> SM read peak (32 MiB resident buffer) next to the HBM copy peak: the reuse mips are L2 resident by design
    l2 = capi.C.c_float(0.0)
    capi.check(capi.lib().vrestir_debug_read_bandwidth(local, 32 << 20, 50, capi.C.byref(l2)))
    l2_peak = float(l2.value)
    if not args.no_cpu_baseline and world == 1:
        imp = gp.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)
        cms, cstage, per_px, threads, sample, wall = cpu_sample(args, scene, params, imp, gp.env_alias(), args.cpu_tiles)
        cpu = {"value": cms, "unit": "ms/frame", "cores": threads, "kind": "port", "sample": sample,
               "stage_ms": {str(k): round(v, 1) for k, v in cstage.items()}}
        # Dominant kernel: k_march, the transmittance-march engine, in its two launches of the spatial-reuse stage (camera
        # stream + light stream).  ALGORITHMIC bytes = what the reference's K3 needs for its p-hat evaluations:
        # voxel bytes (8 voxels x 1 B UNORM8 per trilinear tap) + 36 B per node visit, counted by the instrumented oracle
        # on the sample of the same frame, + the task / result records the engine moves.
        c3 = per_px.get(3, {})
        alg_px = c3.get("voxel_bytes", 0) + 36.0 * c3.get("node_visits", 0)
        task_bytes = mt["spatial_cam_tasks"] * (32 + 12) + mt["spatial_light_tasks"] * (48 + 4)
        alg = alg_px * W * H / max(1, params.mSpatialReuseRounds) + task_bytes   # the oracle counters cover all rounds, the timings one
        t3 = mt["spatial_cam_ms"] + mt["spatial_light_ms"]
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic_k_march.json")
        if os.path.exists(tp):
            traffic = json.load(open(tp)).get("dram_bytes_per_spatial_round")
        if t3 > 0:
            ach = alg / (t3 * 1e-3) / 1e9
            roof = {"kernel": "k_march (march engine; camera + light launches of one spatial-reuse round)", "bound": "hbm", "achieved": ach,
                    "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": alg, "taps_per_px": c3.get("density_taps", 0), "node_visits_per_px": c3.get("node_visits", 0),
                    "launch_ms": t3, "camera_launch_ms": mt["spatial_cam_ms"], "light_launch_ms": mt["spatial_light_ms"],
                    "camera_tasks": mt["spatial_cam_tasks"], "light_tasks": mt["spatial_light_tasks"],
                    "l2_read_peak_gbs": l2_peak, "frac_of_l2_peak": ach / l2_peak if l2_peak > 0 else None,
                    "note": "working set (mip-1 UNORM8 quads, 29 MB) is L2 resident: DRAM traffic << algorithmic bytes; the kernel is issue-bound "
                            "(SIMT divergence between DDA stepping and in-brick sampling), see profiles/"}
            # achieved GB/s per stage = (B_vox + B_node + B_res) / stage time.  B_vox, B_node from the instrumented oracle
            # (stored bytes per voxel of the mip each tap reads, 36 B per node visit), B_res = the reservoir / feature / colour
            # bytes of the stage (R = 32 B); stage times = the unpipelined frame's (a stage timed alone on the GPU).
            R = 32
            res_bytes = {0: 8, 1: R, 2: 2 * R + 16 + R, 3: params.mSpatialSampleCount * R + 8 + R, 5: R + 16}
            st_ms = serial_stage if pipelined else stage_acc
            names = {0: "features_ms", 1: "initial_ms", 2: "temporal_ms", 3: "spatial_ms", 5: "final_ms"}
            stages = {}
            for k, nm in names.items():
                c = per_px.get(k, {})
                b = (c.get("voxel_bytes", 0) + 36.0 * c.get("node_visits", 0) + res_bytes[k]) * W * H
                t = st_ms.get(nm, 0.0)
                if t > 0:
                    stages["K%d" % k] = {"algorithmic_GB": round(b / 1e9, 3), "ms": round(t, 3), "GBps": round(b / (t * 1e-3) / 1e9, 1),
                                          "frac_hbm": round(b / (t * 1e-3) / 1e9 / hbm_peak, 3), "frac_l2": round(b / (t * 1e-3) / 1e9 / l2_peak, 3) if l2_peak > 0 else None}
            roof["stages"] = stages
    line = {"metric": "ms/frame", "value": ms, "unit": "ms/frame", "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": ms, "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD.format(d=args.dim, m=args.mips, w=W, h=H, b=args.bounces), "parallelism": f"rows/{world}" + (f" (cost-balanced bands {sp.bands})" if world > 1 else ""),
                       "l2": "inputs larger than L2 (fp32 mip-0 brick pool + ~0.9 GB/frame of reservoir traffic stream through every frame; the reuse mip is pinned in L2 by design)",
                       "camera": "static", "stage_ms": {k: round(v, 3) for k, v in stage_acc.items()},
                       "pipelining": ({"on": True, "level": level,
                                       "what": "K0+K1 of frame f+1 run on a second stream next to K2..K5 of frame f; at level 2 K5 of frame f runs on a third stream next to "
                                               "K2/K3 of frame f+1 (every frame's stages run exactly once, inside the timed region, which ends after the last frame's K5); "
                                               "stage_ms above are the main stream's (features/initial = wait for the prefetched chain, final = 0 when deferred)",
                                       "prefetch_chain_ms": round(pstats["prefetch_ms"], 3), "deferred_final_ms": round(pstats["deferred_final_ms"], 3), "adopted": int(pstats["adopted"]), "discarded": int(pstats["discarded"]),
                                       "unpipelined_ms_per_frame": round(serial_ms, 3), "unpipelined_stage_ms": {k: round(v, 3) for k, v in serial_stage.items()}}
                                      if pipelined else {"on": False})},
            "clocks": clocks,
            "e2e": {"value": e2e_ms, "unit": "ms/frame", "h2d_bytes_per_step": int(capi.C.sizeof(capi.Camera) + 8192),
                    "d2h_bytes_per_step": int((r1 - r0) * W * 16)},
            "gpu_launches": int(launches),
            "roofline": roof, "cpu_baseline": cpu}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
