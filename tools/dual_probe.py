"""How much do two independent launch chains fill each other's tails?  Two pass instances render the same frame on two
streams, phase-shifted; compares ms per frame with one chain alone.  (Upper bound for overlapping K0/K1 of frame f+1 with
K2..K5 of frame f.)
  python tools/dual_probe.py --width 1920 --height 1080"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from volumetricrestirrelease_b200 import VolumetricReSTIR  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=20)
    ap.add_argument("--band", type=int, nargs=2, default=None)
    a = ap.parse_args()
    args = argparse.Namespace(width=a.width, height=a.height, dim=[577, 572, 438], kind="bunny", mips=4, bounces=1)
    W, H = a.width, a.height
    scene = bench.build_scene(args)
    passes, colors, streams = [], [], []
    for i in range(2):
        gp = VolumetricReSTIR.create({"mParams": bench.make_params(args)}, device=0)
        gp.setScene(scene, W, H)
        if a.band:
            gp.setRowBand(*a.band)
        passes.append(gp)
        colors.append(torch.zeros((H, W, 4), dtype=torch.float32, device="cuda"))
        streams.append(torch.cuda.Stream())

    def timed(fn, n):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(n)
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)

    def single(n):
        streams[0].wait_stream(torch.cuda.current_stream())
        for _ in range(n):
            passes[0].execute(colors[0].data_ptr(), None, streams[0].cuda_stream)

    def dual(n):
        for s in streams:
            s.wait_stream(torch.cuda.current_stream())
        # phase shift: chain 0 is one K0+K1 ahead
        passes[0].execute_stage(0, 0, colors[0].data_ptr(), None, streams[0].cuda_stream)
        passes[0].execute_stage(1, 0, colors[0].data_ptr(), None, streams[0].cuda_stream)
        for _ in range(n):
            passes[1].execute(colors[1].data_ptr(), None, streams[1].cuda_stream)
            for st in (2, 3, 4, 5, 6, 0, 1):
                passes[0].execute_stage(st, 0, colors[0].data_ptr(), None, streams[0].cuda_stream)

    for _ in range(2):
        timed(single, 5)
        timed(dual, 5)
    t1 = timed(single, a.frames) / a.frames
    t2 = timed(dual, a.frames) / (2 * a.frames)
    print(json.dumps({"single_ms_per_frame": round(t1, 3), "dual_ms_per_frame": round(t2, 3), "gain": round(t1 / t2, 3)}))


if __name__ == "__main__":
    main()
