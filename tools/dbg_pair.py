import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
class A: pass
args = A(); args.width, args.height, args.dim, args.kind, args.mips, args.bounces = int(sys.argv[1]), int(sys.argv[2]), [577, 572, 438], "bunny", 4, 1
from volumetricrestirrelease_b200 import VolumetricReSTIR
sc = bench.build_scene(args)
w, h = args.width, args.height
gp = VolumetricReSTIR.create({"mParams": bench.make_params(args)})
gp.setScene(sc, w, h)
color = torch.zeros((h, w, 4), dtype=torch.float32, device="cuda")
for f in range(3):
    gp.execute(color.data_ptr()); torch.cuda.synchronize()
    print("frame", f, "mean", float(color[..., :3].mean()), "counters", gp.wavefront_counters(), flush=True)
