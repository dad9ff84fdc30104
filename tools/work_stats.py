"""Per-stage work statistics (marches / taps / node visits per ACTIVE pixel) from the oracle's counters on a centre crop
of the bench frame.  Test/diagnostic tooling: uses oracle/ as the counter source only."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from oracle import vro

class A: pass
args = A(); args.width, args.height, args.dim, args.kind, args.mips, args.bounces = 1920, 1080, [577, 572, 438], "bunny", 4, 1
scene = bench.build_scene(args); params = bench.make_params(args)
W, H = args.width, args.height
op = vro.OraclePass(params); op.setScene(scene, W, H)
cx, cy, T = W // 2, H // 2, int(sys.argv[1]) if len(sys.argv) > 1 else 96
color = np.zeros((H, W, 4), np.float32)
for fr in range(2):
    for stage in (0, 1, 2, 3, 4, 5):
        halo = 10 if stage in (0, 1, 2) else 0
        op.set_crop(cx - T // 2 - halo, cy - T // 2 - halo, cx + T // 2 + halo, cy + T // 2 + halo)
        op.counters(reset=True)
        t0 = time.time(); op.execute_stage(stage, 0, color); dt = time.time() - t0
        c = op.counters(reset=True)
        npx = (T + 2 * halo) ** 2
        if fr == 1 and c.get("marches", 0):
            print(f"stage {stage}: px={npx} marches/px={c['marches']/npx:.2f} taps/march={c['density_taps']/c['marches']:.2f} "
                  f"visits/march={c['node_visits']/c['marches']:.2f} voxels/tap={c['voxels_fetched']/max(1,c['density_taps']):.2f} rng/px={c['rng_draws']/npx:.1f}  {dt*1e3:.0f} ms")
    op.execute_stage(6, 0, color)
