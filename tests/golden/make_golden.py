"""Regenerates the committed golden fixtures from the CPU oracle:  python tests/golden/make_golden.py

The reference has no golden vectors for this path, and it cannot run here (D3D12-only), so these
fixtures pin the *oracle* (regression) and give the GPU tests a second, committed comparison target."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from common import capi, config1_params, config1_scene, env_scene, vro   # noqa: E402
from volumetricrestirrelease_b200 import VolumetricReSTIRParams          # noqa: E402


def config1():
    w = h = 64
    op = vro.OraclePass(config1_params(4))
    op.setScene(config1_scene(), w, h)
    img = op.execute()
    np.savez_compressed(os.path.join(HERE, "config1_64.npz"), image=img, reservoirs=op.get_buffer(capi.BUF_RESERVOIR_0),
                        features=op.get_buffer(capi.BUF_FEATURES))


def env_reuse():
    w, h = 64, 48
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15, env_size=(128, 64))
    op = vro.OraclePass(VolumetricReSTIRParams())
    op.setScene(sc, w, h)
    op.execute()
    img = op.execute()
    np.savez_compressed(os.path.join(HERE, "env_reuse_64x48.npz"), image=img, reservoirs=op.get_buffer(capi.BUF_RESERVOIR_TEMPORAL),
                        importance_head=op.get_buffer(capi.BUF_ENV_IMPORTANCE).view(np.float32)[:4096].copy())


if __name__ == "__main__":
    config1()
    env_reuse()
    print("golden fixtures written to", HERE)
