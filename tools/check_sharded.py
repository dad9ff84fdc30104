"""Multi-GPU bit-equality check (run under torchrun, one rank per GPU): every rank renders its row band of `frames` frames
through ShardedPass (balanced bands, halo exchanges, frame pipelining as in bench.py) and, on the same GPU, the whole frame
through a plain un-sharded, un-pipelined pass; the band must match the full frame bit for bit on every frame.
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_sharded.py"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from volumetricrestirrelease_b200 import VolumetricReSTIR  # noqa: E402
from volumetricrestirrelease_b200.multi_gpu import ShardedPass  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=360)
    ap.add_argument("--frames", type=int, default=4)
    ap.add_argument("--level", type=int, default=2, help="mPipelineFrames level of the sharded pass")
    ap.add_argument("--motion", type=float, default=0.0, help="camera step per frame in world units (large: the history all-gather fallback must kick in)")
    ap.add_argument("--bounces", type=int, default=1)
    ap.add_argument("--vertex-reuse", type=int, default=0, help="mVertexReuse with this start bounce (0: off); the p_partial plane travels with the halos")
    ap.add_argument("--emissive", type=int, default=0, help="number of emissive triangles around the volume (mUseEmissiveLights)")
    ap.add_argument("--scratch-mb", type=int, default=0, help="mScratchBudgetMB of the sharded pass (small: the generic stages run in row chunks)")
    ap.add_argument("--combo", type=int, default=-1, help="seed of a random option combination (the table of tests/test_gpu_option_combos.py)")
    a = ap.parse_args()
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    args = argparse.Namespace(width=a.width, height=a.height, dim=[289, 286, 219], kind="bunny", mips=4, bounces=a.bounces)
    W, H = a.width, a.height
    scene = bench.build_scene(args)
    params = bench.make_params(args)
    if a.vertex_reuse:
        params.mVertexReuse, params.mVertexReuseStartBounce = 1, a.vertex_reuse
    if a.emissive:
        lo, hi = scene.volume_bounds_world()
        scene.addEmissiveShell(a.emissive, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
        params.mUseEmissiveLights = 1
    combo = ""
    if a.combo >= 0:
        sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
        from test_gpu_option_combos import _draw
        kw, extra = _draw(a.combo)
        for k, v in kw.items():
            setattr(params, k, v)
        lo, hi = scene.volume_bounds_world()
        if extra["point_light"]:
            scene.addPointLight(tuple(0.5 * (lo + hi) + np.array([0.2, 1.1, 0.4]) * (hi - lo)), (9000.0, 7000.0, 5000.0))
        if extra["emissive"]:
            scene.addEmissiveShell(300, tuple(0.5 * (lo + hi)), float(np.linalg.norm(hi - lo)) * 0.75, seed=4)
        combo = f" combo {a.combo} (bounces {params.mMaxBounces}, rounds {params.mSpatialReuseRounds}, taps {params.mSpatialSampleCount}, radius {params.mSampleRadius}, " \
                f"vertex reuse {params.mVertexReuse}/{params.mVertexReuseStartBounce}, reprojection {params.mTemporalReprojectionMode}, lights env{'+point' if extra['point_light'] else ''}{'+emissive' if extra['emissive'] else ''})"
    d = {"mParams": params, "mPipelineFrames": a.level}
    if a.scratch_mb:
        d["mScratchBudgetMB"] = a.scratch_mb
    gp = VolumetricReSTIR.create(d, device=local)
    sp = ShardedPass(gp, W, H, rank, world, torch.device("cuda", local))
    gp.setScene(scene, W, H)
    r0, r1 = sp.balance(refine=0)
    full = VolumetricReSTIR.create({"mParams": params}, device=local)
    full.setScene(scene, W, H)
    c_band = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    c_full = torch.zeros((H, W, 4), dtype=torch.float32, device="cuda")
    bad = 0
    ref = []
    pos0 = np.array(scene.camera.position)
    path = [tuple(pos0 + np.array([0.3, 1.0, 0.1]) * a.motion * f) for f in range(a.frames)]
    for f in range(a.frames):   # the reference frames first: two passes alternating would re-upload the constant bank every frame
        scene.camera.position = path[f]; full.updateCamera()
        full.execute(c_full.data_ptr())   # and (correctly) throw every prefetch away
        torch.cuda.synchronize()
        ref.append(c_full[r0:r1].cpu().numpy().view(np.uint32).copy())
    for f in range(a.frames):
        scene.camera.position = path[f]; gp.updateCamera()
        sp.execute(c_band.data_ptr())
        gp.wait_output()
        x = c_band[r0:r1].cpu().numpy().view(np.uint32)
        bad += int((x != ref[f]).any(axis=-1).sum())
    lit = float((c_full[..., :3].sum(-1) > 0).float().mean())
    t = torch.tensor([bad], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        print(f"[check_sharded] world {world} bands {sp.bands} frames {a.frames} pipelining level {a.level} bounces {a.bounces} vertex reuse {a.vertex_reuse} "
              f"emissive {a.emissive}{combo}: mismatching pixels = {int(t[0])}, lit fraction {lit:.3f}, "
              f"pipeline {gp.pipeline_stats()}, history all-gathers {sp.history_gathers}")
    dist.destroy_process_group()
    sys.exit(1 if int(t[0]) else 0)


if __name__ == "__main__":
    main()
