"""Single-GPU probe of the row-band cost model: renders each band of an N-way partition alone (history everywhere from a few
full frames first) and prints its device time, next to a minimal 8-row band (the fixed per-frame latency of the launch chain).
  python tools/band_probe.py --width 3840 --height 2160 --ways 8"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from volumetricrestirrelease_b200 import VolumetricReSTIR, capi  # noqa: E402
from volumetricrestirrelease_b200.multi_gpu import balanced_row_bands, row_bands, row_cost_from_features  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=3840)
    ap.add_argument("--height", type=int, default=2160)
    ap.add_argument("--ways", type=int, default=8)
    ap.add_argument("--frames", type=int, default=10)
    ap.add_argument("--bgw", type=float, default=0.05)
    ap.add_argument("--only", type=int, nargs=2, default=None, help="render just this band (for an ncu launch list)")
    ap.add_argument("--level", type=int, default=0, help="mPipelineFrames level")
    a = ap.parse_args()
    args = argparse.Namespace(width=a.width, height=a.height, dim=[577, 572, 438], kind="bunny", mips=4, bounces=1)
    W, H = a.width, a.height
    scene = bench.build_scene(args)
    gp = VolumetricReSTIR.create({"mParams": bench.make_params(args), "mPipelineFrames": a.level, "mPrefetchPriority": int(os.environ.get("VR_PF_PRIO", "1"))}, device=0)
    gp.setScene(scene, W, H)
    color = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    gp.setRowBand(0, H)
    gp.execute_stage(0)
    FEAT = np.dtype([("noReflectiveSurface", np.int32), ("transmittance", np.float32)])
    feat = gp.get_buffer(capi.BUF_FEATURES).view(FEAT).reshape(H, W)
    cost = row_cost_from_features(feat, a.bgw)
    active_rows = (feat["transmittance"] != 1.0).sum(axis=1)

    def run(r0, r1, frames=a.frames):
        gp.setRowBand(r0, r1)
        for _ in range(3):
            gp.execute(color.data_ptr())
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(frames):
            gp.execute(color.data_ptr())
        gp.wait_output()
        e1.record()
        torch.cuda.synchronize()
        st = {k: round(v, 3) for k, v in gp.timings().items()}
        return e0.elapsed_time(e1) / frames, st

    if a.only:
        for _ in range(3):
            gp.execute(color.data_ptr())
        ms, st = run(a.only[0], a.only[1], frames=a.frames)
        print(json.dumps({"band": a.only, "ms": round(ms, 3), "stage": st}))
        return
    full, st = run(0, H)
    print(json.dumps({"band": [0, H], "ms": round(full, 3), "stage": st}))
    # fixed latency: 8 rows of sky, 8 rows through the middle of the cloud
    mid = int(np.argmax(active_rows)) // 8 * 8
    for r0 in (0, mid):
        ms, st = run(r0, r0 + 8)
        print(json.dumps({"band": [r0, r0 + 8], "active_px": int(active_rows[r0:r0 + 8].sum()), "ms": round(ms, 3), "stage": st}))
    for name, bands in (("balanced", balanced_row_bands(cost, a.ways, min_rows=16)), ("uniform", row_bands(H, a.ways))):
        tot = []
        for (r0, r1) in bands:
            ms, st = run(r0, r1)
            tot.append(ms)
            print(json.dumps({"part": name, "band": [r0, r1], "active_px": int(active_rows[r0:r1].sum()), "px": (r1 - r0) * W, "ms": round(ms, 3), "stage": st}))
        print(json.dumps({"part": name, "max_ms": round(max(tot), 3), "sum_ms": round(sum(tot), 3), "ideal_ms": round(full / a.ways, 3)}))


if __name__ == "__main__":
    main()
