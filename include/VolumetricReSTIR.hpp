// VolumetricReSTIR.hpp — C++ host-side mirror of the reference pass interface over the C ABI (vrestir.h).
//
// Same method names and argument meaning as `class VolumetricReSTIR : public RenderPass`
// (Source/RenderPasses/VolumetricReSTIR/VolumetricReSTIR.h:41-62): create / reflect / setScene / execute / updateDict /
// getScriptingDictionary.  Falcor types are replaced by plain ones: Dictionary -> std::map<std::string,double> (+ the
// params struct), RenderData outputs -> device pointers, Scene -> the vrestir_* descriptors.  Errors follow the
// reference's convention: failures throw std::runtime_error (VR/VolumetricReSTIR.cpp:196), unknown dictionary keys only
// warn (VR/VolumetricReSTIR.h:309).  Header-only; link with libvrestir.so.
#pragma once
#include <cstdio>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>

#include "vrestir.h"

namespace vrestir {

using Dictionary = std::map<std::string, double>;

struct VolumetricReSTIRParams : vrestir_params {
    VolumetricReSTIRParams() { vrestir_default_params(this); }
};

class VolumetricReSTIR {
public:
    using SharedPtr = std::shared_ptr<VolumetricReSTIR>;

    /// RenderPassLibrary factory signature analogue: create(pRenderContext, dict)
    static SharedPtr create(const VolumetricReSTIRParams& params = VolumetricReSTIRParams(), const Dictionary& dict = {}, int device = 0) {
        SharedPtr p(new VolumetricReSTIR());
        check(vrestir_create(&params, device, &p->mpPass));
        p->updateDict(dict, /*initial*/ true);
        return p;
    }
    ~VolumetricReSTIR() { if (mpPass) vrestir_destroy(mpPass); }

    /// reflect(): the two outputs the pass declares (VR/VolumetricReSTIR.cpp:39-43,149-155)
    static const char* const* reflect(int* count) {
        static const char* kOutputs[] = {"accumulated_color:RGBA32Float", "mvec:RG32Float"};
        if (count) *count = 2;
        return kOutputs;
    }

    /// setScene(): volume (VDBInfo + VolumeDesc), camera, env map, analytic lights, emissive triangles
    void setScene(const vrestir_grid_desc& volume, const vrestir_camera& camera, int width, int height, const vrestir_envmap_desc* env = nullptr,
                  const vrestir_light* lights = nullptr, int lightCount = 0, const vrestir_emissive_triangle* tris = nullptr, int triCount = 0,
                  float emissiveIntensityMultiplier = 1.f, int rowBegin = 0, int rowEnd = -1) {
        check(vrestir_set_frame(mpPass, width, height, rowBegin, rowEnd < 0 ? height : rowEnd));
        check(vrestir_set_volume(mpPass, &volume));
        check(vrestir_set_camera(mpPass, &camera));
        if (env) check(vrestir_set_envmap(mpPass, env));
        check(vrestir_set_analytic_lights(mpPass, lights, lightCount));
        if (tris) check(vrestir_set_emissive_triangles(mpPass, tris, triCount, emissiveIntensityMultiplier));
    }
    void setCamera(const vrestir_camera& camera) { check(vrestir_set_camera(mpPass, &camera)); }
    /// frame pipelining ("mPipelineFrames"): the camera of the NEXT frame, announced before execute() of the current one
    void setNextCamera(const vrestir_camera* camera) { check(vrestir_set_next_camera(mpPass, camera)); }
    /// "mPipelineFrames" = 2: order `cudaStream` after the (deferred) final shading of the last execute()
    void waitOutput(void* cudaStream = nullptr) { check(vrestir_wait_output(mpPass, cudaStream)); }
    /// Scene::update for animated volumes: current grids become the previous-frame slots
    void advanceVolume(const vrestir_grid_desc& volume) { check(vrestir_advance_volume(mpPass, &volume)); }
    /// animated sequences whose frames stay on the device (the reference's protocol, F/Scene/Scene.cpp:825-863): upload once, bind per frame
    int addVolumeFrame(const vrestir_grid_desc& volume) { int index = -1; check(vrestir_volume_frame_add(mpPass, &volume, &index)); return index; }
    void advanceVolumeResident(int index) { check(vrestir_advance_volume_resident(mpPass, index)); }
    void clearVolumeFrames() { check(vrestir_volume_frames_clear(mpPass)); }

    /// execute(pRenderContext, renderData): renderData["accumulated_color"] / ["mvec"] are device pointers here
    void execute(float* accumulated_color, float* mvec = nullptr, void* cudaStream = nullptr) { check(vrestir_execute(mpPass, accumulated_color, mvec, cudaStream)); }
    void executeHost(float* accumulated_color_host, float* mvec_host = nullptr) { check(vrestir_execute_host(mpPass, accumulated_color_host, mvec_host)); }
    /// the same without blocking: the read-back of this frame travels while the next one renders; hostWait() before reading the buffer
    void executeHostAsync(float* accumulated_color_host, float* mvec_host = nullptr) { check(vrestir_execute_host_async(mpPass, accumulated_color_host, mvec_host)); }
    void hostWait() { check(vrestir_host_wait(mpPass)); }

    /// updateDict(): any key resets the frame counter and the temporal history (VR/VolumetricReSTIR.cpp:1339)
    void updateDict(const Dictionary& dict, bool initial = false) {
        for (const auto& kv : dict) {
            int rc = vrestir_update(mpPass, kv.first.c_str(), kv.second);
            if (rc == VRESTIR_WARN_UNKNOWN_KEY) std::fprintf(stderr, "(Warning) Unknown field '%s' in a VolumetricReSTIR dictionary\n", kv.first.c_str());
            else check(rc);
        }
        if (dict.empty() && !initial) { VolumetricReSTIRParams p = getParams(); check(vrestir_set_params(mpPass, &p)); }
    }
    void setParams(const VolumetricReSTIRParams& p) { check(vrestir_set_params(mpPass, &p)); }
    VolumetricReSTIRParams getParams() const { VolumetricReSTIRParams p; check(vrestir_get_params(mpPass, &p)); return p; }
    /// getScriptingDictionary(): the serialisable state is the params struct
    VolumetricReSTIRParams getScriptingDictionary() const { return getParams(); }

    vrestir_timings getTimings() { vrestir_timings t; check(vrestir_get_timings(mpPass, &t)); return t; }
    vrestir_pass* handle() { return mpPass; }

private:
    VolumetricReSTIR() = default;
    VolumetricReSTIR(const VolumetricReSTIR&) = delete;
    static void check(int rc) { if (rc < 0) throw std::runtime_error(std::string("VolumetricReSTIR: ") + vrestir_last_error()); }
    vrestir_pass* mpPass = nullptr;
};

}  // namespace vrestir
