"""The C++ oracle's two deterministic transmittance estimators (ray marching = every p-hat of the reuse stages; analytic
tracking = final shading) against an independent Python restatement written from the Slang (oracle/march_witness.py), on random
rays through 2- and 3-level grids, fp32 and UNORM8 pools, trilinear and point sampling.  CPU only."""
import numpy as np
import pytest

from common import env_scene
from oracle import vro
from oracle.march_witness import Witness
from volumetricrestirrelease_b200 import VolumetricReSTIRParams, capi


def _rays(scene, n, seed):
    lo, hi = scene.volume_bounds_world()
    c, ext = 0.5 * (lo + hi), hi - lo
    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        o = c + (rng.random(3) - 0.5) * ext * (2.2 if len(out) % 3 else 0.6)     # a third of the rays start inside the volume
        tgt = c + (rng.random(3) - 0.5) * ext * 0.8
        d = tgt - o
        d /= np.linalg.norm(d)
        if np.abs(d).min() < 1e-3:
            continue
        tmax = float(3.4e38 if len(out) % 2 else np.linalg.norm(tgt - o) * rng.uniform(0.3, 1.2))
        out.append((o.astype(np.float32), d.astype(np.float32), tmax))
    return out


@pytest.mark.parametrize("dim,three_level", [((64, 64, 56), False), ((200, 150, 140), True)])
def test_oracle_transmittance_matches_slang_witness(dim, three_level):
    sc = env_scene(dim=dim, density_scale=0.004 if not three_level else 0.0015, num_mips=3)   # optical depths of order 1: the sums matter
    grid = sc.volume.grid.contents
    assert (grid.slots[0].top_lev == 2) == three_level
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, 16, 16, importance=np.zeros(349525, np.float32))
    n = 24 if three_level else 40
    cases = [("march", 1, True, 1.0), ("march", 0, True, 2.0), ("march", 2, False, 1.0), ("march", 9, True, 1.0),
             ("analytic", 0, True, 0.0), ("analytic", 9, False, 0.0)]
    nontrivial = 0
    for kind, mip, linear, scale in cases:
        w = Witness(grid, mip)
        for o, d, tmax in _rays(sc, n, seed=mip * 10 + int(linear)):
            if kind == "march":
                got = op.transmittance(o, d, tmax, capi.kRayMarching, mip, linear, scale)
                want = w.ray_marching(o, d, tmax, linear, scale)
            else:
                got = op.transmittance(o, d, tmax, capi.kAnalyticTracking, mip, linear)
                want = w.analytic(o, d, tmax, linear)
            assert np.isclose(got, want, rtol=4e-6, atol=1e-30), (kind, mip, linear, o, d, tmax, got, want)
            nontrivial += 0.02 < want < 0.98
    assert nontrivial > len(cases) * n * 0.4      # most rays end with an intermediate transmittance: the accumulated sums are compared
