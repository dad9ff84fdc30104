"""The contraction-enabled build (csrc `make fast` -> libvrestir_fast.so: FMUL + FADD pairs of the render kernels may fuse) is
not bit-identical to the exact build, so it is held to the north-star tolerances directly: the staged GPU-vs-oracle parity
tests (flips <= 0.1 % of pixels per stage, radiance within 1e-4 relative on non-flipped pixels, accumulated relMSE) are run
again in a child process that loads it instead of libvrestir.so.  Measured (B200, config 2): 7.52 -> 7.34 ms/frame pipelined,
8.30 -> 8.12 serial (k_march<1,true>: 1584 -> 1536 SASS instructions); single-bounce configurations stay within the
tolerances (worst stage 1.3e-4 flips, 9e-5 relative), multi-bounce K1 does NOT (0.68 % of pixels beyond 1e-4 at B = 2: the
1 - (x^2 + y^2) cancellation of the packed bounce directions amplifies the contraction), which is why the exact build stays
the default and this one is opt-in (VRESTIR_LIB) for single-bounce use."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAST = os.path.join(ROOT, "volumetricrestirrelease_b200", "libvrestir_fast.so")
CASES = ("test_config1_single_frame_no_reuse or test_full_reuse_staged_env or test_full_reuse_staged_moving_camera or "
         "test_three_level_tree_full_reuse or test_accumulated_full_reuse_relmse")


@pytest.mark.gpu
def test_fast_build_meets_the_parity_tolerances():
    assert os.path.exists(FAST), "libvrestir_fast.so is missing: run __graft_entry__.build()"
    env = dict(os.environ, VRESTIR_LIB=FAST, VRESTIR_FEATURE_RTOL="1e-4")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-s", "-k", CASES],
                       env=env, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = "\n".join(r.stdout.splitlines()[-40:])
    print(tail)
    assert r.returncode == 0, tail + r.stderr[-2000:]


def test_fast_build_exports_the_same_symbols():
    if not os.path.exists(FAST):
        pytest.skip("libvrestir_fast.so not built")
    import ctypes
    from volumetricrestirrelease_b200 import capi
    lib = ctypes.CDLL(FAST)
    assert [s for s in capi.SYMBOLS if not hasattr(lib, s)] == []
