"""One steady-state frame of a bench configuration between cudaProfilerStart / Stop, for ncu launch lists and captures:

  ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
      --log-file gpurun_out/launches_c4.csv python tools/profile_frame.py --config 4
  ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:k_march -o gpurun_out/prof python tools/profile_frame.py

Frames are un-pipelined (one dependent chain) unless --pipeline is given."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--warm", type=int, default=3)
    ap.add_argument("--frames", type=int, default=1)
    ap.add_argument("--pipeline", type=int, default=0)
    ap.add_argument("--per-pixel", action="store_true")
    ap.add_argument("--set", action="append", default=[], help="pass dictionary entry key=value (e.g. mPrimaryDistanceEngine=0)")
    a = ap.parse_args()
    sys.argv = [sys.argv[0], "--config", str(a.config)] + (["--width", str(a.width)] if a.width else []) + (["--height", str(a.height)] if a.height else [])
    args = bench.parse()
    args.per_pixel = a.per_pixel
    import torch
    scene = bench.build_scene(args)
    params = bench.make_params(args)
    volumes = [bench.build_scene(args, frame_time=bench.PLUME_T0 + 0.35 * f).volume for f in range(1, 4)] if args.config == 3 else []
    R = bench.Runner(args, args.width, args.height, scene, params, a.pipeline, 0, 1, 0, volumes)
    if args.config == 5:
        scene.volume.release_chain()
    if a.set:
        R.gp.updateDict({kv.split("=")[0]: float(kv.split("=")[1]) for kv in a.set})
    for _ in range(a.warm):
        R.step()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStart()
    for _ in range(a.frames):
        R.step()
    R.gp.wait_output()
    torch.cuda.synchronize()
    torch.cuda.cudart().cudaProfilerStop()
    print("launches per frame:", R.gp.launch_count())


if __name__ == "__main__":
    main()
