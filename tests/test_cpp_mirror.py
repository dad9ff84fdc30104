"""CPU: the C++ host mirror (include/VolumetricReSTIR.hpp) compiles with g++, links against libvrestir.so and behaves like
the reference interface at the edges that need no GPU: default parameters, reflect(), and create() throwing (not falling
back to a CPU path) when there is no CUDA device."""
import os
import shutil
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "volumetricrestirrelease_b200")

SRC = r'''
#include <cstdio>
#include <cstring>
#include "VolumetricReSTIR.hpp"
int main() {
    vrestir::VolumetricReSTIRParams p;
    int n = 0;
    const char* const* outs = vrestir::VolumetricReSTIR::reflect(&n);
    std::printf("defaults %d %d %d %g %d\n", p.mMaxBounces, p.mInitialM, p.mSpatialSampleCount, (double)p.mSampleRadius, p.mEnableTemporalReuse);
    std::printf("reflect %d %s %s\n", n, outs[0], outs[1]);
    std::printf("version %s\n", vrestir_version());
    try {
        auto pass = vrestir::VolumetricReSTIR::create(p);
        std::printf("created\n");
    } catch (const std::exception& e) {
        std::printf("threw %s\n", e.what());
    }
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_mirror_compiles_links_and_has_no_cpu_fallback(tmp_path):
    src = tmp_path / "mirror.cpp"
    src.write_text(SRC)
    exe = tmp_path / "mirror"
    r = subprocess.run(["g++", "-std=c++17", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", PKG, "-lvrestir", f"-Wl,-rpath,{PKG}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120).stdout
    assert "defaults 1 4 4 10 1" in out                                       # VR/VolumetricReSTIR.h:130-207 defaults
    assert "reflect 2 accumulated_color:RGBA32Float mvec:RG32Float" in out    # VR/VolumetricReSTIR.cpp:39-43
    assert "version vrestir-b200" in out
    if torch.cuda.is_available():
        assert "created" in out
    else:
        assert "threw" in out and "no CPU fallback" in out
