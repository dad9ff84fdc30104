"""The contraction-enabled build (csrc `make fast` -> libvrestir_fast.so: FMUL + FADD pairs of the render kernels may fuse) is
not bit-identical to the exact build.  The staged GPU-vs-oracle parity tests are run again in a child process that loads it
instead of libvrestir.so, with the radiance tolerance of the north star (1e-4 relative on non-flipped pixels, accumulated
relMSE) and a flip budget of 1 % per stage instead of 0.1 %.

Measured (B200, config 2): 7.52 -> 7.34 ms/frame pipelined, 8.30 -> 8.12 serial (k_march<1,true>: 1584 -> 1536 SASS
instructions).  Most stages stay inside the 0.1 % budget too (worst 3.3e-4), but K1 on a three-level tree reaches 0.33 % and
multi-bounce K1 0.68 % (B = 2: the 1 - (x^2 + y^2) cancellation of the packed bounce directions amplifies the contraction):
pixels whose weights moved by more than 1e-4 count as flips.  That is outside the north-star bar, which is why the exact build
is the default and the parity reference, and this one is opt-in (VRESTIR_LIB)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FAST = os.path.join(ROOT, "volumetricrestirrelease_b200", "libvrestir_fast.so")
CASES = ("test_config1_single_frame_no_reuse or test_full_reuse_staged_env or test_full_reuse_staged_moving_camera or "
         "test_three_level_tree_full_reuse or test_accumulated_full_reuse_relmse or test_emissive_triangles_and_env")


@pytest.mark.gpu
def test_fast_build_stays_within_bounded_tolerances():
    assert os.path.exists(FAST), "libvrestir_fast.so is missing: run __graft_entry__.build()"
    env = dict(os.environ, VRESTIR_LIB=FAST, VRESTIR_FEATURE_RTOL="1e-4", VRESTIR_FLIP_BUDGET="1e-2")
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
