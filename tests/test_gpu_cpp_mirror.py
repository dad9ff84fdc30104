"""GPU: a C++ program that uses only include/VolumetricReSTIR.hpp + include/vrestir.h (the host side a Falcor-style C++ caller
would write: create, setScene, updateDict, setCamera, execute with host buffers, getScriptingDictionary) renders the same frames
as the Python mirror — bit for bit, including after an option change and a camera move — and throws where the reference throws."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np
import pytest

from common import env_scene
from volumetricrestirrelease_b200 import VolumetricReSTIR, VolumetricReSTIRParams, capi
from volumetricrestirrelease_b200.scene import _scene_params

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "volumetricrestirrelease_b200")

SRC = r'''
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "VolumetricReSTIR.hpp"

template <class T> static void readFile(const char* path, T* dst, size_t count) {
    FILE* f = std::fopen(path, "rb");
    if (!f || std::fread(dst, sizeof(T), count, f) != count) { std::fprintf(stderr, "cannot read %s\n", path); std::exit(2); }
    std::fclose(f);
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const std::string dir = argv[1];
    const int W = std::atoi(argv[2]), H = std::atoi(argv[3]), frames = std::atoi(argv[4]);
    vrestir_scene_params sp; readFile((dir + "/scene_params.bin").c_str(), &sp, 1);
    std::vector<vrestir_camera> cams(frames); readFile((dir + "/cameras.bin").c_str(), cams.data(), (size_t)frames);
    vrestir_envmap_desc env; readFile((dir + "/env_desc.bin").c_str(), &env, 1);
    std::vector<float> texels((size_t)env.width * env.height * 4); readFile((dir + "/env_texels.bin").c_str(), texels.data(), texels.size());
    env.texels = texels.data();

    vrestir_scene* scene = nullptr;
    if (vrestir_scene_create(&sp, &scene) < 0) { std::fprintf(stderr, "%s\n", vrestir_last_error()); return 3; }
    try {
        vrestir::VolumetricReSTIRParams params;                          // reference defaults
        auto pass = vrestir::VolumetricReSTIR::create(params, {{"mPipelineFrames", 0}});
        pass->setScene(*vrestir_scene_grid(scene), cams[0], W, H, &env);
        std::vector<float> color((size_t)W * H * 4);
        FILE* out = std::fopen((dir + "/cpp_frames.bin").c_str(), "wb");
        for (int f = 0; f < frames; f++) {
            if (f == 2) pass->updateDict({{"mSpatialSampleCount", 3}, {"mMaxBounces", 2}});   // an option change resets the history
            pass->setCamera(cams[f]);
            pass->executeHost(color.data());
            std::fwrite(color.data(), sizeof(float), color.size(), out);
        }
        std::fclose(out);
        vrestir::VolumetricReSTIRParams now = pass->getScriptingDictionary();
        std::printf("ok bounces %d taps %d launches ", now.mMaxBounces, now.mSpatialSampleCount);
        unsigned long long n = 0; vrestir_get_launch_count(pass->handle(), (uint64_t*)&n); std::printf("%llu\n", n);
        bool threw = false;
        try { pass->updateDict({{"mMaxBounces", 99}}); pass->executeHost(color.data()); } catch (const std::exception& e) { threw = true; std::printf("threw %s\n", e.what()); }
        if (!threw) std::printf("no exception\n");
    } catch (const std::exception& e) { std::fprintf(stderr, "%s\n", e.what()); return 4; }
    vrestir_scene_destroy(scene);
    return 0;
}
'''


@pytest.mark.skipif(shutil.which("g++") is None, reason="g++ not available")
def test_cpp_mirror_renders_the_same_frames_as_the_python_mirror(tmp_path):
    w, h, frames = 160, 96, 4
    sc = env_scene(dim=(64, 64, 56), density_scale=0.15, env_size=(128, 64))
    sp = _scene_params("bunny", (64, 64, 56), 4, 2, (1, 1, 1), (9, 9, 9), 0.0, 0.15, 1.0, (0, 0, 0), 1.0, False, False, 0.005, 1.0, 100.0, 0.0)
    p0 = np.array(sc.camera.position)
    path = [tuple(p0 + np.array((0.7, 0.2, -0.3)) * 8.0 * f) for f in range(frames)]
    cams = (capi.Camera * frames)()
    for f in range(frames):
        sc.camera.position = path[f]
        cams[f] = sc.camera.data(w, h)
    env = sc.envmap_desc()
    open(tmp_path / "scene_params.bin", "wb").write(bytes(sp))
    open(tmp_path / "cameras.bin", "wb").write(bytes(cams))
    open(tmp_path / "env_desc.bin", "wb").write(bytes(env))
    sc.envMap.tofile(tmp_path / "env_texels.bin")
    src = tmp_path / "render.cpp"
    src.write_text(SRC)
    exe = tmp_path / "render"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                        "-L", PKG, "-lvrestir", f"-Wl,-rpath,{PKG}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    run = subprocess.run([str(exe), str(tmp_path), str(w), str(h), str(frames)], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    assert "ok bounces 2 taps 3 launches" in run.stdout and int(run.stdout.split("launches")[1].split()[0]) > 0
    assert "threw VolumetricReSTIR:" in run.stdout                 # failures throw, as in the reference (VR/VolumetricReSTIR.cpp:196)
    got = np.fromfile(tmp_path / "cpp_frames.bin", dtype=np.uint32).reshape(frames, h, w, 4)

    gp = VolumetricReSTIR.create({"mParams": VolumetricReSTIRParams(), "mPipelineFrames": 0})
    sc.camera.position = path[0]
    gp.setScene(sc, w, h)
    for f in range(frames):
        if f == 2:
            gp.updateDict({"mSpatialSampleCount": 3, "mMaxBounces": 2})
        sc.camera.position = path[f]
        gp.updateCamera()
        want = gp.execute_host().view(np.uint32)
        assert np.array_equal(got[f], want), f"frame {f}: the C++ mirror and the Python mirror disagree"
    assert (got[-1].view(np.float32)[..., :3].sum(-1) > 0).mean() > 0.05
