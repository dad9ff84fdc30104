// vr_types.h — plain structs shared by the kernels (vr_kernels.cu) and the host pass (vr_pass.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/vrestir.h"

namespace vrd {

// ------------------------------------------------------------------------------------------------ device scene
struct DSlot {
    int valid, top_lev;
    int dim[3], res[3];
    float vdel[3];
    const vrestir_node* nodes[3];
    const uint32_t* child[3];
    unsigned long long childCount[3];
    float bmin[3], bmax[3];
    float w2m[16];
    float max_value, compress_scale;
    int format, channels;
    const void* atlas;
    const uint32_t* quads;   // UNORM8 single-channel slots: [brick][10][9][9] words of 2x2 xy-neighbours (trilinear = 2 loads)
    unsigned childCount32[3];
    float ivdel[3]; unsigned res3[3];   // 1 / vdel (IEEE, computed on the host) and res^3
    int rootPos[3]; uint32_t rootLink;   // nodes[top_lev][0]: the root never changes during a launch, the march engine reads it from constants
};

struct DScene {
    DSlot slots[VRESTIR_MAX_SLOTS];
    vrestir_volume_desc vol;
    const float4* lut;
    // the cameras are not here: the current one travels in FrameParams (kernel parameter), the previous frame's in DPrevCam —
    // so this block only changes when the scene does, and the kernels of two frames in flight can share it
    int haveEnv, envW, envH;
    const float4* envTexels;
    float envIntensity; float3 envTint;
    float envT[9], envInvT[9], envPrevT[9], envPrevInvT[9];
    const float* importance; int impDim, impBaseMip; unsigned impOffset[12];
    int envSamplerType; const float* envAliasThr; const uint32_t* envAliasRedirect; unsigned envAliasCount;
    int lightCount; const vrestir_light* lights;
    int triCount; const vrestir_emissive_triangle* tris; const uint4* alias; const float* aliasWeights;
    float aliasWeightSum, emissiveMul;
};

// previous frame's camera (K2 only); uploaded on the main stream every frame
struct DPrevCam {
    float prevView[16], prevProj[16];
    float3 prevU, prevV, prevW, prevPos;
};

struct SamplingOptions {   // VR/HostDeviceSharedDefinitions.h:82-138
    uint32_t visibilityTrackingMethod, lightingTrackingMethod;
    int lightSamples, lightingMipLevel, visibilitySamples, visibilityMipLevel;
    int visibilityUseLinearSampler, lightingUseLinearSampler;
    float visibilityTStepScale, lightingTStepScale;
    int useEnvironmentLights, useAnalyticLights, useEmissiveLights;
    int vertexReuseStartBounce;
};

// SoA reservoir storage: plane 0 = (runningSum, M, depth, p_y), plane 1 = (lightUV.x, lightUV.y, lightID, sampledPixel);
// extraBounceStartId is implied by the pixel the record is read from (pixel * (B-1)).
struct ResBuf { float4* p0; float4* p1; float* p2; };   // p2: Reservoir::p_partial, nullptr unless mVertexReuse && mMaxBounces > 1
#define VR_NO_VERTEX_REUSE (1 << 30)                  // SamplingOptions::vertexReuseStartBounce when vertex reuse is off

struct FrameParams {
    int W, H, rowBegin, rowEnd;
    float3 camPos, camU, camV, camW;   // camera of the frame these parameters belong to (the frame in flight or the prefetched next one)
    int frameCount, numTotalRounds, maxBounces;
    int useReference, baselineSpp, initialM, useRussianRoulette, noReuse, useCoarserGrid;
    int visualizeTransmittance, outputMotionVec;
    // temporal
    float temporalMThreshold; uint32_t temporalMIS, reprojectionMode; int reprojectionMip;
    // spatial
    int spatialRounds, roundId, roundOffset; uint32_t spatialMIS; int sampleCount;
    int2 offsets[32];
    SamplingOptions initial, spatial, fin;
    ResBuf cur, out, temporal;
    float3* extCur; float3* extOut; float3* extTemporal;   // packed 12-byte records
    int2* features; const int2* featuresTemporal;          // (noReflectiveSurface, transmittance bits)
    float4* refColor;
    float4* outColor; float2* outMvec;
};

// ------------------------------------------------------------------------------------------------ wavefront streams
// Explicit tasks (originMode 0) are 3 x uint4, prepared by the emitter: (pos.xyz, tNear | dir.xyz, tFar | result index) with
// the ray in the index space of the march mip and clipped to the volume box.  Camera tasks (originMode 1/2: origin = current /
// previous camera position) are 2 x uint4: (thr0, thr1, thr2, threshold mask | world dir.xyz, result index).
struct MarchKind { int mip, linear; float tStepScale; int originMode; float origin[3]; };   // origin: camera-task ray origin (originMode != 0)
struct WfStream { uint4* tasks; unsigned* count; unsigned* cursor; unsigned capacity; };
// per-pixel result block (floats): D[i*4+j] density of tap i's sample seen from ray j, C[j*3+k] camera transmittance along
// ray j to the depth of tap i (k = i - (i > j)), L[i*4+j] light transmittance from that point
enum { WF_BLOCK = 48, WF_D = 0, WF_C = 16, WF_L = 28 };
struct WfBufs { WfStream cam, light; float* results; };
// K1 lock-step candidate state: K1_STRIDE floats per pixel (vr_wavefront.cu)
enum { K1_WORDS = 20, K1_STRIDE = 80, K1_EVAL_BLOCK = 4, K5_BLOCK = 4 };   // K5_BLOCK: floats per pixel of K5's results (density, camera Tr, light Tr)   // K1_EVAL_BLOCK: floats per pixel of K1's p-hat results (density, camera Tr, light Tr)
struct WfInitial { WfStream light; float* state; uint8_t* done; WfStream evalCam, evalLight; float* results; };
// K2: four explicit-origin streams {current camera, current light, previous-frame camera, previous-frame light};
// streams whose march configuration is identical alias the same buffer
struct WfBufs4 { WfStream s[4]; int mip[4]; float* results; };

// Generic (multi-bounce) task-stream path: the stage bodies of vr_stages.cuh run as an emit pass and a consume pass.  Every
// march configuration a stage can ask for is one stream; a pixel's results live at [pixel * stride + eval * MARCH_SLOTS + slot]
// (+ MB_CAM for the shared multi-threshold camera marches of K3).
struct MarchStreams { WfStream s[4]; int mip[4], linear[4], analytic[4]; float scale[4]; int n; };
enum { MB_K2_STRIDE = 20, MB_K3_STRIDE = 96, MB_K3_CAM = 80, MB_K5_STRIDE = 8, MB_K1_EVAL_STRIDE = 8 };
// Multi-bounce K1 as lock-step waves (vr_wavefront.cu): per-pixel state block of K1MB_STRIDE floats
//   [0,4) RNG  [4,16) hd/pd/ot of the <= 4 distance candidates  [16,24) streamed (final) reservoir  [24,32) per-candidate
//   (combined) reservoir  [32,41) path: ray origin, ray dir, pathPdf, pathPHat, primary depth  [41,43) candidate / bounce cursor
//   [44,53) extra-bounce records of the final reservoir  [56,65) of the current candidate  [68,86) the bounce waiting for its
//   shadow march  [95] the marched visibility.  After the first wave the kernels iterate over the previous wave's task list
//   instead of the pixel grid, so warps stay full while most pixels have already finished.
enum { MBK_SG = 0, MBK_HD = 4, MBK_FIN = 16, MBK_COMB = 24, MBK_PATH = 32, MBK_CUR = 41, MBK_KIND = 43, MBK_FINX = 44, MBK_EXTRA = 56, MBK_PEND = 68, MBK_TRAV = 86,
       MBK_VIS = 95, K1MB_STRIDE = 96 };   // KIND: what the pixel waits for (0 shadow march, 1 bounce traversal); TRAV: hit distance / pdf / transmittance of that traversal
// light: the stream this wave's shadow marches go to; prev: the stream of the previous wave, whose task list doubles as the
// compacted list of the pixels that are still running (every pixel that emitted a march continues in the next wave)
// travList / travCount: the pixels (band-local indices) this wave leaves waiting for a bounce traversal; prevTrav*: the previous wave's
// trav: the prepared free-flight tasks of those pixels for the march engine (point sampler; unused with the trilinear one)
struct WfInitialMB { WfStream light, prev, trav; float* state; unsigned* travList; unsigned* travCount; const unsigned* prevTravList; const unsigned* prevTravCount; };

}  // namespace vrd
