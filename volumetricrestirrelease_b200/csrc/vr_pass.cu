// vr_pass.cu — host side of the pass behind the C ABI (include/vrestir.h): option surface, device residency of the
// scene, reservoir buffers, the per-frame dispatch sequence of VR/VolumetricReSTIR.cpp:303-772, buffer access.
//
// B200-first choices made here (details in DESIGN.md):
//   * reservoir / feature history is kept by rotating three device buffers instead of K4's copy and the two
//     copyResource calls (VR/VolumetricReSTIR.cpp:629-638,699-718): zero bytes moved per frame;
//   * the whole scene descriptor lives in __constant__ memory (one 8 KB upload when something changed);
//   * the atlas of the reuse mip is pinned in L2 with an access-policy window on the pass's stream.
// There is no CPU fallback anywhere: every entry point that needs the GPU fails with VRESTIR_ERR_CUDA without one.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <random>
#include <string>
#include <vector>

#include "vr_host.h"
#include "vr_kernels.h"
#include "vr_mipbuild.h"

using namespace vrd;

namespace vr {
static thread_local std::string g_lastError;
int setError(int code, const std::string& msg) { g_lastError = msg; return code; }
int caughtException() {
    try { throw; }
    catch (const std::bad_alloc&) { return setError(VRESTIR_ERR_INVALID_ARGUMENT, "out of host memory (a size in the input is implausibly large)"); }
    catch (const std::exception& e) { return setError(VRESTIR_ERR_INVALID_ARGUMENT, std::string("exception in the host code: ") + e.what()); }
    catch (...) { return setError(VRESTIR_ERR_INVALID_ARGUMENT, "unknown exception in the host code"); }
}
}  // namespace vr
using vr::setError;

#define CK(call)                                                                                                   \
    do {                                                                                                           \
        cudaError_t e__ = (call);                                                                                  \
        if (e__ != cudaSuccess) return setError(VRESTIR_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

namespace {

struct DevSlot {
    void* nodes[3] = {nullptr, nullptr, nullptr}; void* child[3] = {nullptr, nullptr, nullptr}; void* atlas = nullptr; size_t atlasBytes = 0; void* quads = nullptr; size_t quadBytes = 0;
    vrestir_grid_slot meta{};   // what the slot was uploaded from, host pointers cleared (vrestir_download_volume)
    bool borrowed = false;      // the allocations belong to a resident animation frame (vrestir_volume_frame_add)
};

struct KeyDesc { const char* name; size_t off; int type; };
#define K_I(f) {#f, offsetof(vrestir_params, f), 0}
#define K_U(f) {#f, offsetof(vrestir_params, f), 1}
#define K_F(f) {#f, offsetof(vrestir_params, f), 2}
const KeyDesc kKeys[] = {
    K_I(mMaxBounces), K_I(mEnableTemporalReuse), K_I(mEnableSpatialReuse), K_I(mVertexReuse), K_I(mVertexReuseStartBounce), K_I(mUseReference),
    K_I(mUseEnvironmentLights), K_I(mUseAnalyticLights), K_I(mUseEmissiveLights), K_I(mBaselineSamplePerPixel), K_I(mVisualizeTotalTransmittance),
    K_I(mUseSurfaceScene), K_I(mUsePrevVolumeForReproj), K_I(mInitialBaseMipLevel), K_I(mInitialM), K_I(mInitialLightSamples), K_I(mInitialLightingMipLevel),
    K_I(mInitialVisibilityUseLinearSampler), K_I(mInitialLightingUseLinearSampler), K_U(mInitialLightingTrackingMethod), K_F(mInitialVisibilityTStepScale),
    K_F(mInitialLightingTStepScale), K_I(mInitialUseRussianRoulette), K_I(mInitialUseCoarserGridForIndirectBounce), K_F(mTemporalReuseMThreshold),
    K_U(mTemporalReprojectionMode), K_U(mTemporalMISMethod), K_I(mTemporalReprojectionMipLevel), K_I(mSpatialReuseRounds), K_I(mSpatialVisibilityMipLevel),
    K_I(mSpatialLightingMipLevel), K_I(mSpatialVisibilityUseLinearSampler), K_I(mSpatialLightingUseLinearSampler), K_F(mSpatialVisibilityTStepScale),
    K_F(mSpatialLightingTStepScale), K_U(mSpatialVisibilityTrackingMethod), K_U(mSpatialLightingTrackingMethod), K_U(mRandomSamplerType), K_F(mSampleRadius),
    K_I(mSpatialSampleCount), K_I(mEnableVisibilitySimilarityRejection), K_U(mSpatialMISMethod), K_I(mFinalLightSamples), K_I(mFinalVisibilitySamples),
    K_U(mFinalVisibilityTrackingMethod), K_U(mFinalLightTrackingMethod), K_U(mFinalRandomSamplerType), K_F(mFinalTStepScale)};

}  // namespace

struct ScratchArena { void* base = nullptr; size_t bytes = 0; unsigned* counters = nullptr; };

struct vrestir_pass {
    int device = 0;
    vrestir_params P{};
    bool mOutputMotionVec = false, mFreezeFrame = false, mRandomizeFrameSeed = false;
    float densityExtra = -1.f, albedoExtra = -1.f, anisotropyExtra = -1.f;
    int envSamplerType = VRESTIR_ENV_SAMPLER_HIERARCHICAL;
    unsigned randState = 1;

    DScene scene{};
    vrestir_volume_desc volBase{};
    bool sceneDirty = true, haveVolume = false, haveCamera = false;
    DevSlot dslots[VRESTIR_MAX_SLOTS];
    // resident animation frames (vrestir_volume_frame_add): the current-frame slots of each, bound by pointer on advance
    struct ResidentFrame { DevSlot d[VRESTIR_PREV_DENSITY_GRID_OFFSET]; DSlot s[VRESTIR_PREV_DENSITY_GRID_OFFSET]; vrestir_volume_desc vol{}; };
    std::vector<std::unique_ptr<ResidentFrame>> volumeFrames;
    void* d_lut = nullptr; void* d_lutPrev = nullptr;
    vrestir_camera cam{};
    // env
    void* d_env = nullptr; float* d_importance = nullptr; size_t importanceCount = 0;
    float* d_envAliasThr = nullptr; uint32_t* d_envAliasRedirect = nullptr;
    std::vector<float> envAliasThr; std::vector<uint32_t> envAliasRedirect;
    // lights
    void* d_lights = nullptr; void* d_tris = nullptr; void* d_alias = nullptr; void* d_aliasWeights = nullptr;
    std::vector<uint32_t> aliasItems; std::vector<float> aliasWeights; float aliasWeightSum = 0.f;

    int W = 0, H = 0, rowBegin = 0, rowEnd = 0;
    int allocW = 0, allocH = 0, allocB = 0; bool allocVR = false;
    float4* res[4] = {nullptr, nullptr, nullptr, nullptr};
    float3* ext[4] = {nullptr, nullptr, nullptr, nullptr};
    int2* feat[3] = {nullptr, nullptr, nullptr};
    float4* refColor = nullptr;
    int ia = 0, ib = 1, it = 2, in = 3;   // physical indices of ping-pong buffers 0/1, the temporal history and the prefetch target
    int finalPhys = 0;                 // physical buffer holding the final reservoirs of the frame in flight
    int featCur = 0, featPrev = 1, featNext = 2;
    // VR/VolumetricReSTIR.cpp:636 copies the features into the history after every temporal frame.  Here the two buffers swap
    // roles instead, and only when the NEXT active frame starts: until then feat[featCur] stays what the reference's feature
    // buffer holds (K5's transmittance view on a freeze frame, BUF_FEATURES) and doubles as the history (BUF_FEATURES_TEMPORAL).
    bool featSwapPending = false;
    DPrevCam prevCam{};
    // Frame pipelining (option mPipelineFrames): K0 + K1 of frame f+1 read no history, so they run on a stream of their own while
    // K2..K5 of frame f run on the caller's stream (the tail of every launch of one chain is filled by the other chain).
    // The prefetched frame is adopted by the next execute when its key (camera, frame counter, band, options) still matches.
    // Level 2 additionally defers K5 (final shading only reads the frame's final reservoirs, which the next frame reads too but
    // never writes): it runs on a third stream next to K2/K3 of the next frame; consumers order themselves after it with
    // vrestir_wait_output.
    int mPipelineFrames = 0; bool pfValid = false, haveNextCam = false, mPrefetchPriority = true;
    cudaStream_t outStream = nullptr; cudaEvent_t evOutGo = nullptr, evOutDone[2] = {nullptr, nullptr}, evOut0 = nullptr, evOut1 = nullptr;
    uint64_t outSeq = 0; bool outTimed = false;
    uint4* k5Tasks = nullptr; float* k5Results = nullptr; unsigned* k5Counters = nullptr; size_t k5Pixels = 0;
    vrestir_camera nextCam{};
    struct PrefetchKey { float cam[12]; int frameCount, W, H, rowBegin, rowEnd; vrestir_params P; } pfKey{};
    cudaStream_t pfStream = nullptr; cudaEvent_t evPfGo = nullptr, evPfDone = nullptr, evPf0 = nullptr, evPf1 = nullptr;
    bool pfTimed = false; uint64_t pfAdopted = 0, pfDiscarded = 0;
    // K1's own scratch (its kernels may run next to the other stages')
    uint4* k1LightTasks = nullptr; uint4* k1EvalTasks = nullptr; float* k1Results = nullptr; unsigned* k1Counters = nullptr;
    int mFrameCount = 0, mTemporalSampleAccumulated = 0; bool mOptionsChanged = true;
    cudaEvent_t ev[8] = {};
    bool evValid[8] = {};
    // K0 (features) has no consumer before K2: vrestir_execute runs it on an auxiliary stream next to K1
    bool mOverlapFeatures = true, framesOverlapped = false;
    cudaStream_t auxStream = nullptr; cudaEvent_t evFork = nullptr, evJoin = nullptr;
    cudaEvent_t evMarch[3] = {}; bool evMarchValid = false;
                                                                      // evMarch: around the two march launches of the last spatial round
    cudaEvent_t evMainTail = nullptr; bool mainTailValid = false;   // recorded after every stage call: the constant banks are shared per device
    // host-buffer path (vrestir_execute_host[_async]): two device frames, render on hostStream, read-back on hostCopyStream
    cudaStream_t hostStream = nullptr, hostCopyStream = nullptr;
    float4* d_hostColor[2] = {nullptr, nullptr}; float2* d_hostMvec[2] = {nullptr, nullptr}; size_t hostColorPixels = 0;
    cudaEvent_t evHostRendered[2] = {nullptr, nullptr}, evHostLanded[2] = {nullptr, nullptr}; uint64_t hostSeq = 0;
    uint64_t launches = 0;
    vrestir_timings timings{};
    void* persistBase = nullptr; size_t persistBytes = 0; void* persistBasePf = nullptr; size_t persistBytesPf = 0;
    // wavefront (task-stream) path: march-task streams, result blocks, counters {cam.count, cam.cursor, light.count, light.cursor}
    bool mUseWavefront = true;
    int mInitialMode = 1;   // K1: 0 per-pixel kernel, 1 lock-step wavefront (default)
    uint4* wfCamTasks = nullptr; uint4* wfLightTasks = nullptr; float* wfResults = nullptr; unsigned* wfCounters = nullptr;
    // off by default — measured: 13.2 -> 13.8 lanes per instruction, light march 1.67 -> 1.61 ms, the sort itself 0.17 ms: a net loss
    // (profiles/r02_sorted_light_marches.txt)
    // "mPrimaryDistanceEngine": K1's free-flight sampling along the camera rays runs on the march engine (point sampler) instead of the per-pixel traversal kernel
    bool mPrimaryDistanceEngine = true; int primaryBlocks = 0;
    size_t wfPixels = 0;
    int marchBlocks1 = 0, marchBlocks3 = 0, analyticBlocks = 0;
    float* wfInitialState = nullptr; size_t wfInitialPixels = 0;   // lock-step wavefront K1
    // generic task-stream path (multi-bounce option sets): one grow-only arena {tasks | results}, carved per stage and per row chunk
    void* mbArena = nullptr; size_t mbArenaBytes = 0; unsigned* mbCounters = nullptr;
    ScratchArena k1mb, k1mbEval;   // multi-bounce K1 owns its scratch (it may run on the prefetch stream next to the other stages)
    size_t mScratchBudget = (size_t)4 << 30;   // "mScratchBudgetMB": a stage whose worst-case task scratch exceeds this runs in row chunks
    // march launches of the last spatial round when it ran in row chunks (generic path): one event triple + task counts per chunk
    std::vector<cudaEvent_t> evMarchChunks; int marchChunksUsed = 0; unsigned* d_marchCounts = nullptr;
    uint64_t mbChunks = 0; bool mDebugPoison = false;   // "mDebugPoisonResults": result blocks start as NaN, so a march the emit pass missed shows up in the image
};

namespace {

size_t N(const vrestir_pass* p) { return (size_t)p->W * p->H; }
bool vertexReuseOn(const vrestir_pass* p) { return p->P.mVertexReuse && p->P.mMaxBounces > 1; }   // VR/VolumetricReSTIR.cpp:363,376
// planes of one reservoir buffer: p0 | p1 (16 B per pixel each) | p2 = p_partial (4 B per pixel, only with vertex reuse)
ResBuf resView(const vrestir_pass* p, int phys) {
    ResBuf b; b.p0 = p->res[phys]; b.p1 = p->res[phys] + N(p); b.p2 = vertexReuseOn(p) ? (float*)(p->res[phys] + 2 * N(p)) : nullptr;
    return b;
}

int ensureBuffers(vrestir_pass* p) {
    const int B = p->P.mMaxBounces;
    const bool vr = vertexReuseOn(p);
    if (p->allocW == p->W && p->allocH == p->H && p->allocB == B && p->allocVR == vr && p->res[0]) return VRESTIR_OK;
    const size_t n = N(p);
    if (p->pfStream) CK(cudaStreamSynchronize(p->pfStream));     // a prefetch may be writing, a deferred K5 reading them
    if (p->outStream) CK(cudaStreamSynchronize(p->outStream));
    p->pfValid = false;
    for (int i = 0; i < 4; i++) { if (p->res[i]) cudaFree(p->res[i]); p->res[i] = nullptr; }
    for (int i = 0; i < 4; i++) { if (p->ext[i]) cudaFree(p->ext[i]); p->ext[i] = nullptr; }
    for (int i = 0; i < 3; i++) { if (p->feat[i]) cudaFree(p->feat[i]); p->feat[i] = nullptr; }
    if (p->refColor) { cudaFree(p->refColor); p->refColor = nullptr; }
    const size_t resBytes = n * (vr ? 36 : 32);
    for (int i = 0; i < 4; i++) { CK(cudaMalloc(&p->res[i], resBytes)); CK(cudaMemset(p->res[i], 0, resBytes)); }
    for (int i = 0; i < 4; i++) {   // one per reservoir buffer (ping, pong, history, prefetch target)
        if (B > 1) { CK(cudaMalloc(&p->ext[i], n * (size_t)(B - 1) * 12)); CK(cudaMemset(p->ext[i], 0, n * (size_t)(B - 1) * 12)); }
    }
    for (int i = 0; i < 3; i++) { CK(cudaMalloc(&p->feat[i], n * 8)); CK(cudaMemset(p->feat[i], 0, n * 8)); }
    CK(cudaMalloc(&p->refColor, n * 16)); CK(cudaMemset(p->refColor, 0, n * 16));
    p->allocW = p->W; p->allocH = p->H; p->allocB = B; p->allocVR = vr;
    p->ia = 0; p->ib = 1; p->it = 2; p->in = 3; p->finalPhys = 0; p->featCur = 0; p->featPrev = 1; p->featNext = 2; p->featSwapPending = false;
    return VRESTIR_OK;
}

// camera tasks: <= 4 per pixel, light tasks: <= 12 per pixel (S <= 4 taps), result block: WF_BLOCK floats per pixel of the band
int ensureWavefront(vrestir_pass* p) {
    const size_t n = (size_t)(p->rowEnd - p->rowBegin) * p->W;
    if (p->wfPixels == n && p->wfResults) return VRESTIR_OK;
    if (p->wfCamTasks) cudaFree(p->wfCamTasks);
    if (p->wfLightTasks) cudaFree(p->wfLightTasks);
    if (p->wfResults) cudaFree(p->wfResults);
    p->wfCamTasks = p->wfLightTasks = nullptr; p->wfResults = nullptr; p->wfPixels = 0;
    if (n * 12 >= (1ull << 32) || n * WF_BLOCK >= (1ull << 32)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "row band too large for 32-bit task indices; shard the frame");
    CK(cudaMalloc(&p->wfCamTasks, n * 4 * 32));
    CK(cudaMalloc(&p->wfLightTasks, n * 12 * 48));   // explicit (prepared) tasks are 48 B
    CK(cudaMalloc(&p->wfResults, n * WF_BLOCK * sizeof(float)));
    if (!p->wfCounters) CK(cudaMalloc(&p->wfCounters, 128));   // 32 counters: two chains x {stream count / cursor pairs}
    p->wfPixels = n;
    if (!p->marchBlocks1) {
        int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        p->marchBlocks1 = sms * marchBlocksPerSM(1); p->marchBlocks3 = sms * marchBlocksPerSM(3);
    }
    return VRESTIR_OK;
}
WfBufs wfView(const vrestir_pass* p) {
    WfBufs w;
    w.cam.tasks = p->wfCamTasks; w.cam.count = p->wfCounters; w.cam.cursor = p->wfCounters + 1; w.cam.capacity = (unsigned)(p->wfPixels * 4);
    w.light.tasks = p->wfLightTasks; w.light.count = p->wfCounters + 2; w.light.cursor = p->wfCounters + 3; w.light.capacity = (unsigned)(p->wfPixels * 12);
    w.results = p->wfResults;
    return w;
}
// the wavefront forms cover the default option family (single bounce, ray-marched p-hat under the spatial options);
// everything else runs the per-pixel kernels
bool wavefrontEvalOk(const vrestir_pass* p) {
    const vrestir_params& m = p->P;
    return p->mUseWavefront && m.mMaxBounces == 1 && m.mSpatialVisibilityTrackingMethod == VRESTIR_RAY_MARCHING &&
           m.mSpatialLightingTrackingMethod == VRESTIR_RAY_MARCHING;
}
bool wavefrontSpatialOk(const vrestir_pass* p) { return wavefrontEvalOk(p) && p->P.mSpatialSampleCount <= 4; }
void wavefrontKinds(const vrestir_pass* p, MarchKind& cam, MarchKind& light) {
    const vrestir_params& m = p->P;
    cam = MarchKind{m.mSpatialVisibilityMipLevel, m.mSpatialVisibilityUseLinearSampler, m.mSpatialVisibilityTStepScale, 1, {p->cam.posW[0], p->cam.posW[1], p->cam.posW[2]}};
    light = MarchKind{m.mSpatialLightingMipLevel, m.mSpatialLightingUseLinearSampler, m.mSpatialLightingTStepScale, 0, {0.f, 0.f, 0.f}};
}

// ---- generic task-stream path ---------------------------------------------------------------------------------------------
// which stages of the current option set can run as emit / march / consume passes (deterministic tracking only: a stochastic
// tracker draws from the pixel's random-number stream inside the march)
bool genericStageOk(const vrestir_pass* p, int stage) {
    const vrestir_params& m = p->P;
    if (!p->mUseWavefront || m.mUseReference || m.mMaxBounces < 1 || m.mMaxBounces > 4) return false;
    auto det = [](uint32_t method) { return method == VRESTIR_RAY_MARCHING || method == VRESTIR_ANALYTIC_TRACKING; };
    if (stage == 5) return !m.mVisualizeTotalTransmittance && det(m.mFinalVisibilityTrackingMethod) && det(m.mFinalLightTrackingMethod);
    // K2 / K3 evaluate under the spatial options; analytic tracking with the point sampler has no march kernel
    auto ok = [&](uint32_t method, int linear) { return method == VRESTIR_RAY_MARCHING || (method == VRESTIR_ANALYTIC_TRACKING && linear); };
    if (!ok(m.mSpatialVisibilityTrackingMethod, m.mSpatialVisibilityUseLinearSampler) || !ok(m.mSpatialLightingTrackingMethod, m.mSpatialLightingUseLinearSampler)) return false;
    // the shared camera marches of K3 are ray-marched (multi-threshold march engine)
    if (stage == 3) return m.mSpatialSampleCount <= 4 && m.mSpatialVisibilityTrackingMethod == VRESTIR_RAY_MARCHING;
    return true;
}
struct MarchRole { int mip, linear; float scale; uint32_t method; int perPixel; };

// One stage (2 temporal, 3 spatial, 5 final) over the band of `fp`, in row chunks whose worst-case scratch fits the budget.
int growArena(ScratchArena& a, size_t need, cudaStream_t st) {
    if (a.bytes < need) {
        CK(cudaStreamSynchronize(st));
        if (a.base) cudaFree(a.base);
        a.base = nullptr; a.bytes = 0;
        CK(cudaMalloc(&a.base, need));
        a.bytes = need;
    }
    if (!a.counters) CK(cudaMalloc(&a.counters, 64));
    return VRESTIR_OK;
}
int runStageGeneric(vrestir_pass* p, int stage, const FrameParams& fp, cudaStream_t st, ScratchArena* arenaOverride = nullptr, size_t reservedBytesPerPixel = 0, int forcedChunkRows = 0) {
    const vrestir_params& m = p->P;
    const int B = m.mMaxBounces;
    const bool prevGrid = p->scene.vol.usePrevGridForReproj && p->scene.vol.hasAnimation;
    const int off = prevGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
    // worst-case marches per pixel and configuration (an evaluation = 1 camera + (B-1) scatter segments under the visibility options + 1 light march)
    std::vector<MarchRole> roles;
    const SamplingOptions& o = stage == 5 ? fp.fin : fp.spatial;
    int stride = 0, camTasks = 0;
    if (stage == 1) {   // K1's final p-hat of the streamed reservoir (VR/TraceRays.cs.slang:176-183)
        stride = MB_K1_EVAL_STRIDE;
        roles.push_back({o.visibilityMipLevel, o.visibilityUseLinearSampler, o.visibilityTStepScale, o.visibilityTrackingMethod, B});
        roles.push_back({o.lightingMipLevel, o.lightingUseLinearSampler, o.lightingTStepScale, o.lightingTrackingMethod, 1});
    } else if (stage == 2) {
        stride = MB_K2_STRIDE;
        roles.push_back({o.visibilityMipLevel, o.visibilityUseLinearSampler, o.visibilityTStepScale, o.visibilityTrackingMethod, B});        // E(1,0): history sample on the current ray
        roles.push_back({o.lightingMipLevel, o.lightingUseLinearSampler, o.lightingTStepScale, o.lightingTrackingMethod, 1});
        roles.push_back({o.visibilityMipLevel + off, o.visibilityUseLinearSampler, o.visibilityTStepScale, o.visibilityTrackingMethod, B});  // E(0,1): current sample on the previous frame's ray
        roles.push_back({o.lightingMipLevel + off, o.lightingUseLinearSampler, o.lightingTStepScale, o.lightingTrackingMethod, 1});
    } else if (stage == 3) {
        stride = MB_K3_STRIDE; camTasks = 4;
        roles.push_back({o.visibilityMipLevel, o.visibilityUseLinearSampler, o.visibilityTStepScale, o.visibilityTrackingMethod, 12 * (B - 1)});
        roles.push_back({o.lightingMipLevel, o.lightingUseLinearSampler, o.lightingTStepScale, o.lightingTrackingMethod, 12});
    } else {
        stride = MB_K5_STRIDE;
        roles.push_back({o.visibilityMipLevel, o.visibilityUseLinearSampler, o.visibilityTStepScale, o.visibilityTrackingMethod, B});
        roles.push_back({0, o.lightingUseLinearSampler, o.lightingTStepScale, o.lightingTrackingMethod, 1});   // final shading marches the light ray through mip 0
    }
    // identical configurations share a stream
    std::vector<MarchRole> uniq;
    for (const MarchRole& r : roles) {
        if (r.perPixel <= 0) continue;
        bool merged = false;
        for (MarchRole& u : uniq)
            if (u.mip == r.mip && u.linear == r.linear && u.scale == r.scale && (u.method == VRESTIR_ANALYTIC_TRACKING) == (r.method == VRESTIR_ANALYTIC_TRACKING)) { u.perPixel += r.perPixel; merged = true; break; }
        if (!merged) uniq.push_back(r);
    }
    if (uniq.size() > 4) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "internal: more than four march configurations in one stage");
    for (const MarchRole& u : uniq)
        if (u.mip < 0 || u.mip >= VRESTIR_MAX_SLOTS || !p->scene.slots[u.mip].valid) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "a reuse stage names a grid slot that is not bound");
    size_t tasksPerPixel = 0; for (const MarchRole& u : uniq) tasksPerPixel += (size_t)u.perPixel;
    const size_t bytesPerPixel = tasksPerPixel * 48 + (size_t)camTasks * 32 + (size_t)stride * 4;
    const int bandRows = fp.rowEnd - fp.rowBegin;
    size_t maxRows = p->mScratchBudget / ((bytesPerPixel + reservedBytesPerPixel) * (size_t)fp.W);
    maxRows = std::min<size_t>(maxRows, ((size_t)1 << 32) / ((size_t)fp.W * (size_t)std::max<size_t>(stride, 1)) - 1);   // 32-bit result indices
    int chunkRows = (int)std::min<size_t>((size_t)bandRows, std::max<size_t>(8, maxRows / 8 * 8));
    if (forcedChunkRows) chunkRows = forcedChunkRows;   // the caller already runs on a chunk of its own
    const size_t chunkPixels = (size_t)chunkRows * fp.W;
    const size_t need = chunkPixels * bytesPerPixel + 256;
    ScratchArena mainArena{p->mbArena, p->mbArenaBytes, p->mbCounters};
    ScratchArena& arena = arenaOverride ? *arenaOverride : mainArena;
    { int rcA = growArena(arena, need, st); if (rcA) return rcA; }
    if (!arenaOverride) { p->mbArena = arena.base; p->mbArenaBytes = arena.bytes; p->mbCounters = arena.counters; }
    unsigned* const counters = arena.counters;
    if (!p->marchBlocks1 || !p->analyticBlocks) {
        int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        p->marchBlocks1 = sms * marchBlocksPerSM(1); p->marchBlocks3 = sms * marchBlocksPerSM(3); p->analyticBlocks = sms * analyticBlocksPerSM();
    }
    // carve: [results | camera tasks | stream 0 | stream 1 | ...]
    char* base = (char*)arena.base;
    float* results = (float*)base; base += (chunkPixels * stride * 4 + 15) / 16 * 16;
    WfStream cam{}; cam.tasks = (uint4*)base; cam.count = counters; cam.cursor = counters + 1; cam.capacity = (unsigned)(chunkPixels * camTasks);
    base += chunkPixels * camTasks * 32;
    MarchStreams ms{}; ms.n = (int)uniq.size();
    for (int k = 0; k < ms.n; k++) {
        ms.s[k].tasks = (uint4*)base; ms.s[k].count = counters + 2 + 2 * k; ms.s[k].cursor = counters + 3 + 2 * k;
        ms.s[k].capacity = (unsigned)std::min<size_t>(chunkPixels * (size_t)uniq[k].perPixel, 0xffffffffull);
        base += chunkPixels * (size_t)uniq[k].perPixel * 48;
        ms.mip[k] = uniq[k].mip; ms.linear[k] = uniq[k].linear ? 1 : 0; ms.scale[k] = uniq[k].scale; ms.analytic[k] = uniq[k].method == VRESTIR_ANALYTIC_TRACKING ? 1 : 0;
    }
    for (int r0 = fp.rowBegin; r0 < fp.rowEnd; r0 += chunkRows) {
        FrameParams fc = fp;
        fc.rowBegin = r0; fc.rowEnd = std::min(fp.rowEnd, r0 + chunkRows);
        CK(cudaMemsetAsync(counters, 0, 64, st));
        if (p->mDebugPoison) CK(cudaMemsetAsync(results, 0xFF, chunkPixels * stride * 4, st));
        CK(launchStageEmit(stage, fc, ms, cam, results, st)); p->launches++;
        const int chunkIdx = (r0 - fp.rowBegin) / chunkRows;
        const bool timeMarches = stage == 3 && chunkIdx < 64;   // the march launches of the spatial round, per chunk: roofline input
        cudaEvent_t* evc = nullptr;
        if (timeMarches) {
            while ((int)p->evMarchChunks.size() < 3 * (chunkIdx + 1)) { cudaEvent_t e; CK(cudaEventCreate(&e)); p->evMarchChunks.push_back(e); }
            if (!p->d_marchCounts) CK(cudaMalloc(&p->d_marchCounts, 128 * sizeof(unsigned)));
            evc = &p->evMarchChunks[3 * chunkIdx];
            CK(cudaEventRecord(evc[0], st));
        }
        if (camTasks) {
            const MarchKind kc = {o.visibilityMipLevel, o.visibilityUseLinearSampler, o.visibilityTStepScale, 1, {fp.camPos.x, fp.camPos.y, fp.camPos.z}};
            CK(launchMarch(cam, results, kc, p->scene.slots[kc.mip], 3, p->marchBlocks3, st)); p->launches++;
        }
        if (timeMarches) CK(cudaEventRecord(evc[1], st));
        for (int k = 0; k < ms.n; k++) {
            const MarchKind kk = {ms.mip[k], ms.linear[k], ms.scale[k], 0, {0.f, 0.f, 0.f}};
            if (ms.analytic[k]) CK(launchMarchAnalytic(ms.s[k], results, kk, p->scene.slots[kk.mip], p->analyticBlocks, st));
            else CK(launchMarch(ms.s[k], results, kk, p->scene.slots[kk.mip], 1, p->marchBlocks1, st));
            p->launches++;
        }
        if (timeMarches) {
            CK(cudaEventRecord(evc[2], st));
            p->evMarchValid = true; p->marchChunksUsed = chunkIdx + 1;
            CK(cudaMemcpyAsync(p->d_marchCounts + 2 * chunkIdx, counters, 4, cudaMemcpyDeviceToDevice, st));
            CK(cudaMemcpyAsync(p->d_marchCounts + 2 * chunkIdx + 1, counters + 2, 4, cudaMemcpyDeviceToDevice, st));
        }
        CK(launchStageConsume(stage, fc, results, st)); p->launches++;
        p->mbChunks++;
    }
    return VRESTIR_OK;
}

// Device blocks of the grid slots come from a per-device cache: an animated sequence that uploads a new volume every frame
// (vrestir_advance_volume, vrestir_set_volume_from_chain) releases and requests ~100 blocks of nearly the same sizes per frame,
// and cudaMalloc / cudaFree (a device-wide synchronisation each) cost more than the copies.  A released block is handed out
// again for a request of up to its capacity and at least 3/4 of it.  Callers synchronise the device before releasing blocks
// that in-flight work may still read (they did so before their cudaFree, too).
struct SlotBlockCache {
    std::mutex mu;
    std::multimap<size_t, void*> free;      // capacity -> block
    std::map<void*, size_t> capacity;       // every live or cached block handed out by slotAlloc
    size_t cachedBytes = 0;
};
std::map<int, std::unique_ptr<SlotBlockCache>> g_slotCaches;
std::mutex g_slotCachesMu;
constexpr size_t kSlotCacheLimit = (size_t)4 << 30;
SlotBlockCache& slotCache() {
    int dev = 0; cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(g_slotCachesMu);
    auto& c = g_slotCaches[dev];
    if (!c) c.reset(new SlotBlockCache());
    return *c;
}
cudaError_t slotAlloc(void** out, size_t bytes) {
    SlotBlockCache& c = slotCache();
    {
        std::lock_guard<std::mutex> lock(c.mu);
        auto it = c.free.lower_bound(bytes);
        if (it != c.free.end() && it->first - bytes <= it->first / 4) {
            *out = it->second; c.cachedBytes -= it->first; c.free.erase(it);
            return cudaSuccess;
        }
    }
    const cudaError_t e = cudaMalloc(out, bytes);
    if (e == cudaSuccess) { std::lock_guard<std::mutex> lock(c.mu); c.capacity[*out] = bytes; }
    return e;
}
void slotRelease(void* ptr) {
    if (!ptr) return;
    SlotBlockCache& c = slotCache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto it = c.capacity.find(ptr);
    if (it == c.capacity.end()) { cudaFree(ptr); return; }
    if (c.cachedBytes + it->second > kSlotCacheLimit) { c.capacity.erase(it); cudaFree(ptr); return; }
    c.free.emplace(it->second, ptr); c.cachedBytes += it->second;
}
void slotCacheTrim() {   // vrestir_destroy: cached blocks go back to the driver
    SlotBlockCache& c = slotCache();
    std::lock_guard<std::mutex> lock(c.mu);
    for (auto& kv : c.free) { c.capacity.erase(kv.second); cudaFree(kv.second); }
    c.free.clear(); c.cachedBytes = 0;
}

void freeSlot(DevSlot& d) {
    if (!d.borrowed) {
        for (int l = 0; l < 3; l++) { slotRelease(d.nodes[l]); slotRelease(d.child[l]); }
        slotRelease(d.atlas);
        slotRelease(d.quads);
    }
    d = DevSlot{};
}
bool ownsMemory(const DevSlot& d) { return !d.borrowed && (d.atlas || d.quads || d.nodes[0] || d.nodes[1] || d.nodes[2]); }

int uploadSlotTo(DevSlot& d, DSlot& s, const vrestir_grid_slot& g) {
    freeSlot(d);
    memset(&s, 0, sizeof(s));
    if (!g.valid) return VRESTIR_OK;
    if (g.top_lev < 1 || g.top_lev > 2) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "top_lev must be 1 or 2");
    s.valid = 1; s.top_lev = g.top_lev;
    for (int l = 0; l < 3; l++) {
        s.dim[l] = g.dim[l]; s.res[l] = g.res[l]; s.vdel[l] = g.vdel[l];
        if (g.node_count[l] && g.nodes[l]) {
            CK(slotAlloc(&d.nodes[l], (size_t)g.node_count[l] * sizeof(vrestir_node)));
            CK(cudaMemcpy(d.nodes[l], g.nodes[l], (size_t)g.node_count[l] * sizeof(vrestir_node), cudaMemcpyHostToDevice));
        }
        if (g.childlist_count[l] && g.childlist[l]) {
            CK(slotAlloc(&d.child[l], (size_t)g.childlist_count[l] * 4));
            CK(cudaMemcpy(d.child[l], g.childlist[l], (size_t)g.childlist_count[l] * 4, cudaMemcpyHostToDevice));
        }
        if (g.childlist_count[l] >= (1ull << 31)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "child list too large");
        s.nodes[l] = (const vrestir_node*)d.nodes[l]; s.child[l] = (const uint32_t*)d.child[l]; s.childCount[l] = g.childlist_count[l];
        s.childCount32[l] = (unsigned)g.childlist_count[l];
        s.ivdel[l] = 1.0f / g.vdel[l]; s.res3[l] = (unsigned)(g.res[l] * g.res[l] * g.res[l]);
    }
    if (!g.nodes[g.top_lev] || !g.node_count[g.top_lev]) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "slot has no root node");
    for (int i = 0; i < 3; i++) s.rootPos[i] = g.nodes[g.top_lev][0].pos[i];
    s.rootLink = g.nodes[g.top_lev][0].link;
    if ((unsigned long long)g.brick_count * g.atlas_channels * VRESTIR_BRICK_VOXELS >= (1ull << 32)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "brick pool too large for 32-bit voxel indices");
    for (int i = 0; i < 3; i++) { s.bmin[i] = g.bmin[i]; s.bmax[i] = g.bmax[i]; }
    memcpy(s.w2m, g.world_to_medium, 64);
    s.max_value = g.max_value; s.compress_scale = g.compress_scale; s.format = g.atlas_format; s.channels = g.atlas_channels;
    const size_t bytes = (size_t)g.brick_count * g.atlas_channels * VRESTIR_BRICK_VOXELS * (g.atlas_format == VRESTIR_ATLAS_UNORM8 ? 1 : 4);
    if (bytes) {
        CK(slotAlloc(&d.atlas, bytes + 16));
        if (g.atlas) CK(cudaMemcpy(d.atlas, g.atlas, bytes, cudaMemcpyHostToDevice));   // NULL: the caller fills the pool on the device
        d.atlasBytes = bytes;
    }
    s.atlas = d.atlas;
    d.meta = g;
    for (int l = 0; l < 3; l++) { d.meta.nodes[l] = nullptr; d.meta.childlist[l] = nullptr; }
    d.meta.atlas = nullptr;
    s.quads = nullptr;
    if (bytes && g.atlas && g.atlas_format == VRESTIR_ATLAS_UNORM8 && g.atlas_channels == 1) {
        // device-only repack for trilinear fetches: per brick [10][9][9] words, word(z,y,x) = codes (x,y) (x+1,y) (x,y+1) (x+1,y+1) of plane z
        d.quadBytes = (size_t)g.brick_count * 810 * 4;
        CK(slotAlloc(&d.quads, d.quadBytes));
        CK(vr::launchQuadRepack((const uint8_t*)d.atlas, g.brick_count, (uint32_t*)d.quads, 0));
        s.quads = (const uint32_t*)d.quads;
    }
    return VRESTIR_OK;
}
int uploadSlot(vrestir_pass* p, int slot, const vrestir_grid_slot& g) { return uploadSlotTo(p->dslots[slot], p->scene.slots[slot], g); }

void applyOverrides(vrestir_pass* p) {   // VR/VolumetricReSTIR.cpp:211-235
    vrestir_volume_desc v = p->volBase;
    if (p->densityExtra > 0) v.densityScaleFactor = p->densityExtra;
    if (p->anisotropyExtra > 0) v.PhaseFunctionConstantG = p->anisotropyExtra;
    if (p->albedoExtra > 0) for (int i = 0; i < 3; i++) { v.sigma_s[i] = v.sigma_t * p->albedoExtra; v.sigma_a[i] = v.sigma_t - v.sigma_s[i]; }
    v.usePrevGridForReproj = p->P.mUsePrevVolumeForReproj;
    if (memcmp(&v, &p->scene.vol, sizeof(v)) != 0) { p->scene.vol = v; p->sceneDirty = true; }
}

SamplingOptions mkOpt(uint32_t vt, uint32_t lt, int ls, int lm, int vs, int vm, int vl, int ll, float vts, float lts, const vrestir_params& m) {
    SamplingOptions o;
    o.visibilityTrackingMethod = vt; o.lightingTrackingMethod = lt; o.lightSamples = ls; o.lightingMipLevel = lm; o.visibilitySamples = vs; o.visibilityMipLevel = vm;
    o.visibilityUseLinearSampler = vl; o.lightingUseLinearSampler = ll; o.visibilityTStepScale = vts; o.lightingTStepScale = lts;
    o.useEnvironmentLights = m.mUseEnvironmentLights; o.useAnalyticLights = m.mUseAnalyticLights; o.useEmissiveLights = m.mUseEmissiveLights;
    o.vertexReuseStartBounce = (m.mVertexReuse && m.mMaxBounces > 1) ? m.mVertexReuseStartBounce : VR_NO_VERTEX_REUSE;
    return o;
}

// F/Utils/Math/MathHelpers.slang:178-197,356-361; VR/SpatialReuse.cs.slang:64-81 — offsets depend only on (sampleId, frame seed): host-side, fp64 stays off the GPU
float radicalInverse(uint32_t i) {
    i = (i & 0x55555555u) << 1 | (i & 0xAAAAAAAAu) >> 1; i = (i & 0x33333333u) << 2 | (i & 0xCCCCCCCCu) >> 2;
    i = (i & 0x0F0F0F0Fu) << 4 | (i & 0xF0F0F0F0u) >> 4; i = (i & 0x00FF00FFu) << 8 | (i & 0xFF00FF00u) >> 8;
    i = (i << 16) | (i >> 16);
    return (float)i * 2.3283064365386963e-10f;
}
int f2iHost(float v) { if (std::isnan(v)) return 0; if (v >= 2147483648.f) return INT32_MAX; if (v <= -2147483648.f) return INT32_MIN; return (int)v; }
int2 neighborOffset(const vrestir_params& m, int sampleId, int frameId) {
    float ux, uy;
    if (m.mRandomSamplerType == VRESTIR_SAMPLER_HAMMERSLEY) { ux = (float)sampleId / (float)m.mSpatialSampleCount; uy = radicalInverse((uint32_t)sampleId); }
    else if (sampleId == 0) { ux = 0; uy = 0; }
    else {
        double mult = (double)(frameId * m.mSpatialSampleCount + sampleId);
        double a = 0.754877669 * mult, b = 0.569840296 * mult;
        ux = (float)(a - std::floor(a)); uy = (float)(b - std::floor(b));
    }
    // sample_disk: the two libm calls below run on the host for every caller (oracle and product do the same)
    float r = sqrtf(ux), phi = 6.28318530717958647693f * uy;
    return make_int2(f2iHost(m.mSampleRadius * (r * cosf(phi))), f2iHost(m.mSampleRadius * (r * sinf(phi))));
}

void buildFrameParams(vrestir_pass* p, FrameParams& fp, float* out_color, float* out_mvec) {
    const vrestir_params& m = p->P;
    memset(&fp, 0, sizeof(fp));
    fp.W = p->W; fp.H = p->H; fp.rowBegin = p->rowBegin; fp.rowEnd = p->rowEnd;
    auto f3of = [](const float* a) { return make_float3(a[0], a[1], a[2]); };
    fp.camPos = f3of(p->cam.posW); fp.camU = f3of(p->cam.cameraU); fp.camV = f3of(p->cam.cameraV); fp.camW = f3of(p->cam.cameraW);
    fp.frameCount = p->mFrameCount;
    fp.numTotalRounds = (m.mEnableSpatialReuse ? m.mSpatialReuseRounds : 0) + (m.mEnableTemporalReuse ? 1 : 0) + 1 + 1;   // VR/VolumetricReSTIR.cpp:452-453
    fp.maxBounces = m.mMaxBounces;
    fp.useReference = m.mUseReference; fp.baselineSpp = m.mBaselineSamplePerPixel; fp.initialM = m.mInitialM;
    fp.useRussianRoulette = m.mInitialUseRussianRoulette; fp.noReuse = !m.mEnableSpatialReuse && !m.mEnableTemporalReuse;
    fp.useCoarserGrid = m.mInitialUseCoarserGridForIndirectBounce;
    fp.visualizeTransmittance = m.mVisualizeTotalTransmittance; fp.outputMotionVec = p->mOutputMotionVec;
    fp.temporalMThreshold = m.mTemporalReuseMThreshold; fp.temporalMIS = m.mTemporalMISMethod; fp.reprojectionMode = m.mTemporalReprojectionMode;
    fp.reprojectionMip = m.mTemporalReprojectionMipLevel;
    fp.spatialRounds = m.mSpatialReuseRounds; fp.roundOffset = (m.mEnableTemporalReuse ? 1 : 0) + 1; fp.spatialMIS = m.mSpatialMISMethod; fp.sampleCount = m.mSpatialSampleCount;
    // VR/VolumetricReSTIR.cpp:457-496
    fp.initial = mkOpt(VRESTIR_ANALYTIC_TRACKING, m.mInitialLightingTrackingMethod, m.mInitialLightSamples, m.mInitialLightingMipLevel, 1,
                       m.mInitialVisibilityUseLinearSampler ? m.mInitialBaseMipLevel : m.mInitialBaseMipLevel + VRESTIR_NUM_MAX_MIPS, m.mInitialVisibilityUseLinearSampler,
                       m.mInitialLightingUseLinearSampler, m.mInitialVisibilityTStepScale, m.mInitialLightingTStepScale, m);
    fp.spatial = mkOpt(m.mSpatialVisibilityTrackingMethod, m.mSpatialLightingTrackingMethod, 1, m.mSpatialLightingMipLevel, 1, m.mSpatialVisibilityMipLevel,
                       m.mSpatialVisibilityUseLinearSampler, m.mSpatialLightingUseLinearSampler, m.mSpatialVisibilityTStepScale, m.mSpatialLightingTStepScale, m);
    const bool lightDet = m.mFinalLightTrackingMethod == VRESTIR_ANALYTIC_TRACKING || m.mFinalLightTrackingMethod == VRESTIR_RAY_MARCHING;
    const bool visDet = m.mFinalVisibilityTrackingMethod == VRESTIR_ANALYTIC_TRACKING || m.mFinalVisibilityTrackingMethod == VRESTIR_RAY_MARCHING;
    fp.fin = mkOpt(m.mFinalVisibilityTrackingMethod, m.mFinalLightTrackingMethod, lightDet ? 1 : m.mFinalLightSamples, 0, visDet ? 1 : m.mFinalVisibilitySamples, 0, 1, 1,
                   m.mFinalTStepScale, m.mFinalTStepScale, m);
    fp.features = p->feat[p->featCur]; fp.featuresTemporal = p->feat[p->featPrev];
    fp.refColor = p->refColor;
    fp.outColor = (float4*)out_color; fp.outMvec = (float2*)out_mvec;
}

// The __constant__ banks (scene block, previous camera) exist once per device and are shared by every pass of the process on
// it.  The registry remembers what each device holds and which passes live there, so that an upload (a) happens whenever the
// bank may hold another pass's data and (b) is ordered after EVERYTHING the other passes still have in flight that reads the
// old contents: their main-stream work, a prefetched K0/K1 (which is then discarded: it would be adopted next to constants it
// was not computed with only if they were identical, but its owner re-uploads anyway) and a deferred K5.
struct Uploaded { DScene scene; DPrevCam prev; bool have = false; const vrestir_pass* owner = nullptr; std::vector<vrestir_pass*> passes; };
std::mutex g_constMu;
std::map<int, std::unique_ptr<Uploaded>> g_constPerDevice;
void registerPass(vrestir_pass* p) {
    std::lock_guard<std::mutex> lock(g_constMu);
    std::unique_ptr<Uploaded>& slot = g_constPerDevice[p->device];
    if (!slot) slot.reset(new Uploaded());
    slot->passes.push_back(p);
}
void unregisterPass(vrestir_pass* p) {
    std::lock_guard<std::mutex> lock(g_constMu);
    auto it = g_constPerDevice.find(p->device);
    if (it == g_constPerDevice.end() || !it->second) return;
    auto& v = it->second->passes;
    v.erase(std::remove(v.begin(), v.end(), p), v.end());
    if (it->second->owner == p) { it->second->owner = nullptr; it->second->have = false; }
}
// make `st` wait for everything the OTHER passes of the device have in flight (they may read the constants about to change)
int waitForOtherPasses(vrestir_pass* p, Uploaded& slot, cudaStream_t st) {
    for (vrestir_pass* q : slot.passes) {
        if (q == p) continue;
        if (q->mainTailValid) CK(cudaStreamWaitEvent(st, q->evMainTail, 0));
        if (q->pfValid) { CK(cudaStreamWaitEvent(st, q->evPfDone, 0)); q->pfValid = false; q->pfDiscarded++; }
        if (q->outSeq) CK(cudaStreamWaitEvent(st, q->evOutDone[(q->outSeq - 1) & 1], 0));
    }
    return VRESTIR_OK;
}

int syncScene(vrestir_pass* p, cudaStream_t st) {
    applyOverrides(p);
    DScene& s = p->scene;
    s.envSamplerType = p->envSamplerType;
    std::lock_guard<std::mutex> lock(g_constMu);
    std::unique_ptr<Uploaded>& slot = g_constPerDevice[p->device];
    if (!slot) slot.reset(new Uploaded());
    DScene& lastScene = slot->scene; DPrevCam& lastPrev = slot->prev; bool& haveLast = slot->have;
    const bool sceneDiffers = !haveLast || p->sceneDirty || memcmp(&lastScene, &s, sizeof(DScene)) != 0;
    const bool prevDiffers = !haveLast || memcmp(&lastPrev, &p->prevCam, sizeof(DPrevCam)) != 0;
    if ((sceneDiffers || prevDiffers) && slot->passes.size() > 1) { int rc = waitForOtherPasses(p, *slot, st); if (rc) return rc; }
    if (sceneDiffers || prevDiffers) slot->owner = p;
    if (sceneDiffers) {
        // a prefetched K0/K1 was computed with the old constants and may still be reading them
        if (p->pfValid) { CK(cudaStreamWaitEvent(st, p->evPfDone, 0)); p->pfValid = false; p->pfDiscarded++; }
        if (p->outSeq) CK(cudaStreamWaitEvent(st, p->evOutDone[(p->outSeq - 1) & 1], 0));
        CK(uploadScene(s, st));
        CK(uploadSceneWavefront(s, st));
        // the async copy reads `s` at enqueue time only when the source is pageable (staged); keep a private copy alive
        lastScene = s; p->sceneDirty = false;
        if (p->pfStream) {   // later work on the prefetch stream must see the new constants
            CK(cudaEventRecord(p->evPfGo, st)); CK(cudaStreamWaitEvent(p->pfStream, p->evPfGo, 0));
        }
    }
    if (prevDiffers) {   // read by K2 only, which runs on `st`
        CK(uploadPrevCam(p->prevCam, st));
        CK(uploadPrevCamWavefront(p->prevCam, st));
        lastPrev = p->prevCam;
    }
    haveLast = true;
    return VRESTIR_OK;
}

void recordEv(vrestir_pass* p, int i, cudaStream_t st) { if (!p->ev[i]) cudaEventCreate(&p->ev[i]); cudaEventRecord(p->ev[i], st); p->evValid[i] = true; }

int setPersistingWindow(vrestir_pass* p, cudaStream_t st, bool prefetchStream = false) {
    // "a coarse mip pinned in B200's L2": the atlas of the reuse mip (mSpatialVisibilityMipLevel) gets a persisting
    // access-policy window on the stream that runs K2/K3.
    int slot = p->P.mSpatialVisibilityMipLevel;
    if (slot < 0 || slot >= VRESTIR_MAX_SLOTS || !p->dslots[slot].atlas) return VRESTIR_OK;
    void* base = p->dslots[slot].quads ? p->dslots[slot].quads : p->dslots[slot].atlas;
    size_t bytes = p->dslots[slot].quads ? p->dslots[slot].quadBytes : p->dslots[slot].atlasBytes;
    void*& cachedBase = prefetchStream ? p->persistBasePf : p->persistBase;
    size_t& cachedBytes = prefetchStream ? p->persistBytesPf : p->persistBytes;
    if (base == cachedBase && bytes == cachedBytes) return VRESTIR_OK;
    int maxWin = 0, maxPersist = 0;
    cudaDeviceGetAttribute(&maxWin, cudaDevAttrMaxAccessPolicyWindowSize, p->device);
    cudaDeviceGetAttribute(&maxPersist, cudaDevAttrMaxPersistingL2CacheSize, p->device);
    if (maxWin <= 0 || maxPersist <= 0) return VRESTIR_OK;
    size_t want = std::min(bytes, (size_t)maxPersist);
    cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want);
    cudaStreamAttrValue attr{};
    attr.accessPolicyWindow.base_ptr = base;
    attr.accessPolicyWindow.num_bytes = std::min(bytes, (size_t)maxWin);
    attr.accessPolicyWindow.hitRatio = bytes <= want ? 1.0f : (float)((double)want / (double)bytes);
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(st, cudaStreamAttributeAccessPolicyWindow, &attr);
    cudaGetLastError();
    cachedBase = base; cachedBytes = bytes;
    return VRESTIR_OK;
}

// lock-step wavefront K1 (vr_wavefront.cu): single bounce, reuse on, <= 4 candidates, ray-marched light visibility
bool initialWavefrontOk(const vrestir_pass* p) {
    const vrestir_params& m = p->P;
    const bool noReuse = !m.mEnableSpatialReuse && !m.mEnableTemporalReuse;
    return wavefrontEvalOk(p) && !m.mUseReference && !noReuse && m.mInitialM <= 4 && m.mInitialLightingTrackingMethod == VRESTIR_RAY_MARCHING &&
           m.mInitialLightSamples <= 1 && p->mInitialMode == 1;
}

void makePrefetchKey(const vrestir_pass* p, const vrestir_camera& c, int frameCount, vrestir_pass::PrefetchKey& k) {
    memset(&k, 0, sizeof(k));
    memcpy(k.cam, c.posW, 12); memcpy(k.cam + 3, c.cameraU, 12); memcpy(k.cam + 6, c.cameraV, 12); memcpy(k.cam + 9, c.cameraW, 12);
    k.frameCount = frameCount; k.W = p->W; k.H = p->H; k.rowBegin = p->rowBegin; k.rowEnd = p->rowEnd;
    k.P = p->P;
}
// does the prefetched K0/K1 belong to the frame that is starting now?
bool prefetchMatches(const vrestir_pass* p) {
    if (!p->pfValid || p->mFreezeFrame) return false;
    vrestir_pass::PrefetchKey k;
    makePrefetchKey(p, p->cam, p->mFrameCount, k);
    return memcmp(&k, &p->pfKey, sizeof(k)) == 0;
}

int ensureInitialScratch(vrestir_pass* p) {
    const size_t n = (size_t)(p->rowEnd - p->rowBegin) * p->W;
    if (p->wfInitialPixels == n && p->wfInitialState) return VRESTIR_OK;
    if (p->pfStream) CK(cudaStreamSynchronize(p->pfStream));
    p->pfValid = false;
    void* old[] = {p->wfInitialState, p->k1LightTasks, p->k1EvalTasks, p->k1Results};
    for (void* q : old) if (q) cudaFree(q);
    p->wfInitialState = nullptr; p->k1LightTasks = p->k1EvalTasks = nullptr; p->k1Results = nullptr; p->wfInitialPixels = 0;
    if (n * K1_STRIDE >= (1ull << 32)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "row band too large for 32-bit record indices; shard the frame");
    CK(cudaMalloc(&p->wfInitialState, n * K1_STRIDE * sizeof(float) + n));   // + one done flag per pixel
    CK(cudaMalloc(&p->k1LightTasks, n * 48));                                // one light task per pixel and candidate wave
    CK(cudaMalloc(&p->k1EvalTasks, 2 * n * 48));                             // final p-hat: <= one camera + one light task per pixel
    CK(cudaMalloc(&p->k1Results, n * K1_EVAL_BLOCK * sizeof(float)));
    if (!p->k1Counters) CK(cudaMalloc(&p->k1Counters, 64));
    p->wfInitialPixels = n;
    if (!p->marchBlocks1) {
        int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        p->marchBlocks1 = sms * marchBlocksPerSM(1); p->marchBlocks3 = sms * marchBlocksPerSM(3);
    }
    return VRESTIR_OK;
}

// K5's own scratch: <= 2 analytic march tasks (camera, light) and K5_BLOCK result floats per pixel of the band
int ensureFinalScratch(vrestir_pass* p) {
    const size_t n = (size_t)(p->rowEnd - p->rowBegin) * p->W;
    if (p->k5Pixels == n && p->k5Tasks) return VRESTIR_OK;
    if (p->outStream) CK(cudaStreamSynchronize(p->outStream));
    if (p->k5Tasks) cudaFree(p->k5Tasks);
    if (p->k5Results) cudaFree(p->k5Results);
    p->k5Tasks = nullptr; p->k5Results = nullptr; p->k5Pixels = 0;
    if (n * 2 >= (1ull << 32)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "row band too large for 32-bit task indices; shard the frame");
    CK(cudaMalloc(&p->k5Tasks, 2 * n * 48));
    CK(cudaMalloc(&p->k5Results, n * K5_BLOCK * sizeof(float)));
    if (!p->k5Counters) CK(cudaMalloc(&p->k5Counters, 16));
    p->k5Pixels = n;
    return VRESTIR_OK;
}

// K1 as M + 1 lock-step waves over the band: traverse, then per candidate s {finish candidate s-1 / emit the light march of
// candidate s, march}, the p-hat evaluation of the surviving sample under the spatial options, finish.  Works entirely in
// its own scratch, on `st` (the caller's stream or the prefetch stream).  fp.cur is the output reservoir buffer.
int runInitialWavefront(vrestir_pass* p, const FrameParams& fp, cudaStream_t st) {
    const vrestir_params& m = p->P;
    int rc = ensureInitialScratch(p); if (rc) return rc;
    const size_t n = p->wfInitialPixels;
    MarchKind kc, klp; wavefrontKinds(p, kc, klp); kc.originMode = 0; kc.origin[0] = kc.origin[1] = kc.origin[2] = 0.f;   // explicit-origin tasks
    const bool oneEval = memcmp(&kc, &klp, sizeof(MarchKind)) == 0;
    const MarchKind kl = {m.mInitialLightingMipLevel, m.mInitialLightingUseLinearSampler, m.mInitialLightingTStepScale, 0, {0.f, 0.f, 0.f}};
    unsigned* cnt = p->k1Counters;
    CK(cudaMemsetAsync(cnt, 0, 64, st));
    WfInitial wi;
    wi.light.tasks = p->k1LightTasks; wi.light.count = cnt; wi.light.cursor = cnt + 1; wi.light.capacity = (unsigned)n;
    wi.state = p->wfInitialState; wi.done = (uint8_t*)(p->wfInitialState + n * K1_STRIDE);
    wi.results = p->k1Results;
    wi.evalCam.tasks = p->k1EvalTasks; wi.evalCam.count = cnt + 4; wi.evalCam.cursor = cnt + 5; wi.evalCam.capacity = (unsigned)(oneEval ? 2 * n : n);
    if (oneEval) wi.evalLight = wi.evalCam;
    else { wi.evalLight.tasks = p->k1EvalTasks + 3 * n; wi.evalLight.count = cnt + 6; wi.evalLight.cursor = cnt + 7; wi.evalLight.capacity = (unsigned)n; }
    const bool primaryEngine = p->mPrimaryDistanceEngine && !m.mInitialVisibilityUseLinearSampler;
    if (primaryEngine && !p->primaryBlocks) { int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device); p->primaryBlocks = sms * primaryDistanceBlocksPerSM(); }
    for (int s = 0; s <= m.mInitialM; s++) {
        if (s > 0 && s < m.mInitialM) CK(cudaMemsetAsync(cnt, 0, 8, st));
        if (s == 0 && primaryEngine) {
            CK(launchPrimaryDistance(fp, wi.state, K1_STRIDE, 24 /* K1_HD */, cnt + 12, p->scene.slots[fp.initial.visibilityMipLevel], p->primaryBlocks, st));
            CK(launchInitialStepOnly(fp, wi, 0, st));
        } else CK(launchInitialStep(fp, wi, s, st));
        p->launches += s == 0 ? 2 : 1;
        if (s < m.mInitialM) { CK(launchMarch(wi.light, wi.state, kl, p->scene.slots[kl.mip], 1, p->marchBlocks1, st)); p->launches++; }
    }
    CK(launchMarch(wi.evalCam, wi.results, kc, p->scene.slots[kc.mip], 1, p->marchBlocks1, st)); p->launches++;
    if (!oneEval) { CK(launchMarch(wi.evalLight, wi.results, klp, p->scene.slots[klp.mip], 1, p->marchBlocks1, st)); p->launches++; }
    CK(launchInitialFinish(fp, wi, st)); p->launches++;
    return VRESTIR_OK;
}

// lock-step wavefront K1 for 2-4 bounces (vr_wavefront.cu): reuse on, <= 4 candidates, ray-marched light visibility, deterministic
// tracking under the spatial options (the final p-hat)
bool initialWavefrontMBOk(const vrestir_pass* p) {
    const vrestir_params& m = p->P;
    const bool noReuse = !m.mEnableSpatialReuse && !m.mEnableTemporalReuse;
    return p->mUseWavefront && m.mMaxBounces >= 2 && m.mMaxBounces <= 4 && !m.mUseReference && !noReuse && m.mInitialM <= 4 &&
           m.mInitialLightingTrackingMethod == VRESTIR_RAY_MARCHING && m.mInitialLightSamples <= 1 && p->mInitialMode == 1 && genericStageOk(p, 2);
}
// K1 as <= M * B waves over the band (in row chunks under the scratch budget): traverse, then per wave {advance every pixel to
// its next shadow march, march}, then the p-hat of the streamed reservoir as an emit / march / consume pass.
int runInitialWavefrontMB(vrestir_pass* p, const FrameParams& fp, cudaStream_t st) {
    const vrestir_params& m = p->P;
    const int B = m.mMaxBounces, M = m.mInitialM;
    const size_t stateBytes = (size_t)K1MB_STRIDE * 4 + 3 * 48 + 2 * 4;                       // state block, one light task in each of the two (ping-pong) streams, one free-flight task, one entry in each traversal list
    const size_t evalBytes = (size_t)(B + 1) * 48 + (size_t)MB_K1_EVAL_STRIDE * 4;
    size_t maxRows = p->mScratchBudget / ((stateBytes + evalBytes) * (size_t)fp.W);
    maxRows = std::min<size_t>(maxRows, ((size_t)1 << 32) / ((size_t)fp.W * K1MB_STRIDE) - 1);   // 32-bit record indices
    const int bandRows = fp.rowEnd - fp.rowBegin;
    const int chunkRows = (int)std::min<size_t>((size_t)bandRows, std::max<size_t>(8, maxRows / 8 * 8));
    const size_t chunkPixels = (size_t)chunkRows * fp.W;
    { int rc = growArena(p->k1mb, chunkPixels * stateBytes + 256, st); if (rc) return rc; }
    if (!p->marchBlocks1) {
        int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
        p->marchBlocks1 = sms * marchBlocksPerSM(1); p->marchBlocks3 = sms * marchBlocksPerSM(3);
    }
    const MarchKind kl = {m.mInitialLightingMipLevel, m.mInitialLightingUseLinearSampler, m.mInitialLightingTStepScale, 0, {0.f, 0.f, 0.f}};
    // two task streams in turn: the step kernel of wave w walks the task list of wave w-1 (= the pixels still running) and fills the other one
    WfStream streams[2];
    float* const state = (float*)p->k1mb.base;
    for (int k = 0; k < 2; k++) {
        streams[k].tasks = (uint4*)(state + chunkPixels * K1MB_STRIDE) + (size_t)k * 3 * chunkPixels;
        streams[k].count = p->k1mb.counters + 2 * k; streams[k].cursor = p->k1mb.counters + 2 * k + 1; streams[k].capacity = (unsigned)chunkPixels;
    }
    WfStream travStream;                                                                     // free-flight tasks of the indirect bounces (march engine, point sampler)
    travStream.tasks = streams[1].tasks + 3 * chunkPixels; travStream.count = p->k1mb.counters + 4; travStream.cursor = p->k1mb.counters + 5; travStream.capacity = (unsigned)chunkPixels;
    unsigned* const travLists = (unsigned*)(travStream.tasks + 3 * chunkPixels);             // two lists of chunkPixels entries behind the task streams
    const bool travEngine = !m.mInitialVisibilityUseLinearSampler;
    int travMip = fp.initial.visibilityMipLevel;                                             // VR/ComputeInitialSample.slang:60-66 (mbBounceMip)
    if (m.mInitialUseCoarserGridForIndirectBounce) travMip = std::min((travMip >= VRESTIR_NUM_MAX_MIPS ? VRESTIR_NUM_MAX_MIPS : 0) + p->scene.vol.numMips - 1, travMip + 1);
    const MarchKind kt = {travMip, 0, 1.f, 0, {0.f, 0.f, 0.f}};
    static int distBlocksPerSM = 0;
    if (!distBlocksPerSM) distBlocksPerSM = distanceBlocksPerSM();
    static int stepBlocksPerSM = 0, travBlocksPerSM = 0;
    if (!stepBlocksPerSM) { stepBlocksPerSM = initialMBStepBlocksPerSM(); travBlocksPerSM = initialMBBounceTraverseBlocksPerSM(); }
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device);
    const int stepBlocks = (int)std::min<size_t>((chunkPixels + 127) / 128, (size_t)sms * stepBlocksPerSM);
    const int travBlocks = (int)std::min<size_t>((chunkPixels + 127) / 128, (size_t)sms * travBlocksPerSM);
    for (int r0 = fp.rowBegin; r0 < fp.rowEnd; r0 += chunkRows) {
        FrameParams fc = fp;
        fc.rowBegin = r0; fc.rowEnd = std::min(fp.rowEnd, r0 + chunkRows);
        WfInitialMB wi; wi.state = state; wi.light = streams[0]; wi.prev = streams[1]; wi.trav = travStream;
        unsigned* const travCounts = p->k1mb.counters + 8;
        int cur = 0;
        wi.travList = travLists; wi.travCount = travCounts; wi.prevTravList = travLists + chunkPixels; wi.prevTravCount = travCounts + 1;
        CK(cudaMemsetAsync(p->k1mb.counters, 0, 64, st));
        if (p->mPrimaryDistanceEngine && travEngine) {
            if (!p->primaryBlocks) p->primaryBlocks = sms * primaryDistanceBlocksPerSM();
            CK(launchPrimaryDistance(fc, wi.state, K1MB_STRIDE, MBK_HD, p->k1mb.counters + 12, p->scene.slots[fc.initial.visibilityMipLevel], p->primaryBlocks, st));
        } else CK(launchInitialMBTraverse(fc, wi, st));
        CK(launchInitialMBStep(fc, wi, 1, 0, st));
        p->launches += 2;
        // every wave advances every running pixel to its next suspension: the shadow march of a bounce (<= B per candidate) or the
        // free-flight traversal of an indirect bounce (<= B - 1 per candidate)
        // (running the two marches of a wave on two streams was measured: 143.1 -> 142.3 ms, either one fills the GPU — kept serial)
        for (int w = 0; w < M * (2 * B - 1); w++) {
            CK(launchMarch(wi.light, wi.state, kl, p->scene.slots[kl.mip], 1, p->marchBlocks1, st));
            std::swap(wi.light, wi.prev);
            cur ^= 1;
            wi.prevTravList = wi.travList; wi.prevTravCount = wi.travCount;
            wi.travList = travLists + (size_t)cur * chunkPixels; wi.travCount = travCounts + cur;
            if (travEngine) {
                CK(launchMarchDistance(wi.trav, wi.state, kt, p->scene.slots[kt.mip], sms * distBlocksPerSM, st));
                CK(cudaMemsetAsync(wi.trav.count, 0, 8, st));
            } else CK(launchInitialMBBounceTraverse(fc, wi, travBlocks, st));
            CK(cudaMemsetAsync(wi.light.count, 0, 8, st));
            CK(cudaMemsetAsync(wi.travCount, 0, 4, st));
            CK(launchInitialMBStep(fc, wi, 0, stepBlocks, st));
            p->launches += 3;
        }
        int rc = runStageGeneric(p, 1, fc, st, &p->k1mbEval, 0, fc.rowEnd - fc.rowBegin); if (rc) return rc;
    }
    return VRESTIR_OK;
}

int runStageBody(vrestir_pass* p, int stage, int arg, float* out_color, float* out_mvec, cudaStream_t st);
int runStage(vrestir_pass* p, int stage, int arg, float* out_color, float* out_mvec, cudaStream_t st) {
    const int rc = runStageBody(p, stage, arg, out_color, out_mvec, st);
    if (p && rc == VRESTIR_OK && stage != 6) {   // what other passes of the device must wait for before they touch the constant banks
        if (!p->evMainTail) CK(cudaEventCreateWithFlags(&p->evMainTail, cudaEventDisableTiming));
        CK(cudaEventRecord(p->evMainTail, st));
        p->mainTailValid = true;
    }
    return rc;
}
int runStageBody(vrestir_pass* p, int stage, int arg, float* out_color, float* out_mvec, cudaStream_t st) {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (!p->haveVolume || !p->haveCamera || p->W <= 0) return setError(VRESTIR_ERR_NOT_READY, "volume/camera/frame not set");
    const vrestir_params& m = p->P;
    if (m.mUseSurfaceScene) return setError(VRESTIR_ERR_UNSUPPORTED, "mUseSurfaceScene: surface scenes are outside the hot-path scope (SURVEY.md 8f rank 4)");
    if (m.mMaxBounces < 1 || m.mMaxBounces > VRESTIR_MAX_BOUNCES) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "mMaxBounces must be 1..4");
    if (m.mSpatialSampleCount < 1 || m.mSpatialSampleCount > 32) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "mSpatialSampleCount must be 1..32");
    if (m.mInitialM < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "mInitialM must be >= 1");
    if (m.mUseEnvironmentLights && !p->scene.haveEnv) return setError(VRESTIR_ERR_NOT_READY, "mUseEnvironmentLights without an env map");
    if (m.mUseAnalyticLights && p->scene.lightCount == 0) return setError(VRESTIR_ERR_NOT_READY, "mUseAnalyticLights without lights");
    {   // every mip the options name must be bound
        auto ok = [&](int slot) { return slot >= 0 && slot < VRESTIR_MAX_SLOTS && p->scene.slots[slot].valid; };
        const int initVis = m.mInitialVisibilityUseLinearSampler ? m.mInitialBaseMipLevel : m.mInitialBaseMipLevel + VRESTIR_NUM_MAX_MIPS;
        const bool reuse = m.mEnableSpatialReuse || m.mEnableTemporalReuse;
        bool good = ok(0) && ok(m.mInitialLightingMipLevel) && (reuse ? ok(initVis) : true) && ok(m.mSpatialVisibilityMipLevel) && ok(m.mSpatialLightingMipLevel) &&
                    (m.mEnableTemporalReuse && m.mTemporalReprojectionMode != VRESTIR_REPROJECTION_NONE ? ok(VRESTIR_NUM_MAX_MIPS + m.mTemporalReprojectionMipLevel) : true);
        if (!good) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "an option names a mip level that the volume does not have");
    }
    CK(cudaSetDevice(p->device));
    int rc = ensureBuffers(p); if (rc) return rc;
    if ((stage == 0 || stage == -1) && p->mOptionsChanged) {   // VR/VolumetricReSTIR.cpp:349-359
        if (p->mRandomizeFrameSeed) p->mFrameCount = rand_r(&p->randState) % 65536; else p->mFrameCount = 0;
        p->mTemporalSampleAccumulated = 0; p->mOptionsChanged = false;
    }
    if ((stage == 0 || stage == -1) && p->outSeq >= 2) CK(cudaStreamWaitEvent(st, p->evOutDone[p->outSeq & 1], 0));   // deferred K5 of frame f-2
    rc = syncScene(p, st); if (rc) return rc;
    const bool active = !p->mFreezeFrame;
    if ((stage == 0 || stage == -1) && active && p->featSwapPending) { std::swap(p->featCur, p->featPrev); p->featSwapPending = false; }
    FrameParams fp; buildFrameParams(p, fp, out_color, out_mvec);
    switch (stage) {
        case -1:   // frame begin on the main stream when K0 runs concurrently on the auxiliary stream (vrestir_execute)
            recordEv(p, 7, st);
            setPersistingWindow(p, st);
            break;
        case 0:
            recordEv(p, 0, st);
            setPersistingWindow(p, st);
            if (prefetchMatches(p)) {
                // K0 of this frame ran ahead on the prefetch stream into feat[featNext]: adopt it (stage 1 waits for the chain)
                std::swap(p->featCur, p->featNext);
            } else if (active && !m.mUseReference) { CK(launchFeatures(fp, st)); p->launches++; }
            recordEv(p, 1, st);
            break;
        case 1:
            if (active) {
                if (prefetchMatches(p)) {
                    // K1 of this frame ran ahead into res[in]: it becomes ping-pong buffer 0, the old one the next prefetch target
                    CK(cudaStreamWaitEvent(st, p->evPfDone, 0));
                    std::swap(p->ia, p->in);
                    p->pfValid = false; p->pfAdopted++;
                    p->finalPhys = p->ia;
                    recordEv(p, 2, st);
                    break;
                }
                if (p->pfValid) {   // stale prefetch (camera / options moved on): its scratch must be idle before K1 reuses it
                    CK(cudaStreamWaitEvent(st, p->evPfDone, 0)); p->pfValid = false; p->pfDiscarded++;
                }
                fp.cur = resView(p, p->ia); fp.extCur = p->ext[p->ia];
                if (initialWavefrontOk(p) || initialWavefrontMBOk(p)) {
                    rc = initialWavefrontOk(p) ? runInitialWavefront(p, fp, st) : runInitialWavefrontMB(p, fp, st); if (rc) return rc;
                    p->finalPhys = p->ia;
                    recordEv(p, 2, st);
                    break;
                }
                CK(launchInitial(fp, st)); p->launches++;
                p->finalPhys = p->ia;
            }
            recordEv(p, 2, st);
            break;
        case 7: {
            // Prefetch: K0 + K1 of the NEXT frame on the prefetch stream.  Call after stage 1 of the frame in flight (its K1 has
            // released the K1 scratch) and before stage 2, so that the chain overlaps K2..K5.  No-op unless mPipelineFrames.
            if (!p->mPipelineFrames || !active || m.mUseReference || !(initialWavefrontOk(p) || initialWavefrontMBOk(p)) || p->pfValid) break;
            if (!p->pfStream) {
                // High priority: the chain is ~13 short dependent launches; at default priority each of them queues behind the
                // resident persistent CTAs of the main stream's march kernels and the chain stretches over the whole frame
                // (measured: 2.3 ms alone -> 4.4 ms next to K2..K5 on a 1/8 band of a 4K frame), which makes it the critical path.
                int prLeast = 0, prGreatest = 0;
                CK(cudaDeviceGetStreamPriorityRange(&prLeast, &prGreatest));
                CK(cudaStreamCreateWithPriority(&p->pfStream, cudaStreamNonBlocking, p->mPrefetchPriority ? prGreatest : prLeast));
                CK(cudaEventCreateWithFlags(&p->evPfGo, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&p->evPfDone, cudaEventDisableTiming));
                CK(cudaEventCreate(&p->evPf0)); CK(cudaEventCreate(&p->evPf1));
            }
            FrameParams fn = fp;
            const vrestir_camera& c = p->haveNextCam ? p->nextCam : p->cam;
            auto f3of = [](const float* a) { return make_float3(a[0], a[1], a[2]); };
            fn.camPos = f3of(c.posW); fn.camU = f3of(c.cameraU); fn.camV = f3of(c.cameraV); fn.camW = f3of(c.cameraW);
            fn.frameCount = p->mFrameCount + 1;
            fn.features = p->feat[p->featNext];
            fn.cur = resView(p, p->in); fn.extCur = p->ext[p->in];
            // everything enqueued so far on `st` (the previous frame, this frame's K1) is done with the target buffers / the scratch
            CK(cudaEventRecord(p->evPfGo, st));
            CK(cudaStreamWaitEvent(p->pfStream, p->evPfGo, 0));
            setPersistingWindow(p, p->pfStream, true);
            CK(cudaEventRecord(p->evPf0, p->pfStream));
            CK(launchFeatures(fn, p->pfStream)); p->launches++;
            rc = initialWavefrontOk(p) ? runInitialWavefront(p, fn, p->pfStream) : runInitialWavefrontMB(p, fn, p->pfStream); if (rc) return rc;
            CK(cudaEventRecord(p->evPf1, p->pfStream));
            CK(cudaEventRecord(p->evPfDone, p->pfStream));
            p->pfTimed = true;
            makePrefetchKey(p, c, fn.frameCount, p->pfKey);
            p->pfValid = true;
            p->haveNextCam = false;   // an announcement covers one frame; without a new one the camera is assumed to stay
            break;
        }
        case 2:
            if (active && !m.mUseReference && m.mEnableTemporalReuse) {
                if (p->mTemporalSampleAccumulated != 0) {   // gIsFirstFrame skips the kernel (VR/TemporalReuse.cs.slang:93)
                    fp.cur = resView(p, p->ia); fp.extCur = p->ext[p->ia];
                    fp.temporal = resView(p, p->it); fp.extTemporal = p->ext[p->it];
                    if (wavefrontEvalOk(p)) {
                        rc = ensureWavefront(p); if (rc) return rc;
                        // four explicit-origin streams {camera, light} x {current, previous frame}; identical march configurations share one stream
                        const bool prevGrid = p->scene.vol.usePrevGridForReproj && p->scene.vol.hasAnimation;
                        const int off = prevGrid ? VRESTIR_PREV_DENSITY_GRID_OFFSET : 0;
                        MarchKind kinds[4] = {{m.mSpatialVisibilityMipLevel, m.mSpatialVisibilityUseLinearSampler, m.mSpatialVisibilityTStepScale, 0},
                                              {m.mSpatialLightingMipLevel, m.mSpatialLightingUseLinearSampler, m.mSpatialLightingTStepScale, 0},
                                              {m.mSpatialVisibilityMipLevel + off, m.mSpatialVisibilityUseLinearSampler, m.mSpatialVisibilityTStepScale, 0},
                                              {m.mSpatialLightingMipLevel + off, m.mSpatialLightingUseLinearSampler, m.mSpatialLightingTStepScale, 0}};
                        for (int k = 0; k < 4; k++)
                            if (kinds[k].mip < 0 || kinds[k].mip >= VRESTIR_MAX_SLOTS || !p->scene.slots[kinds[k].mip].valid) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "temporal reuse names a grid slot that is not bound");
                        WfBufs4 wf{};
                        wf.results = p->wfResults;
                        const size_t n = p->wfPixels;
                        // one buffer (16 n tasks) is plenty: <= 4 tasks per pixel in this stage
                        uint4* bufs[4] = {p->wfLightTasks, p->wfLightTasks + 3 * (2 * n), p->wfLightTasks + 3 * (4 * n), p->wfLightTasks + 3 * (6 * n)};
                        bool unique[4];
                        for (int k = 0; k < 4; k++) {
                            int alias = -1;
                            for (int j = 0; j < k && alias < 0; j++) if (memcmp(&kinds[j], &kinds[k], sizeof(MarchKind)) == 0) alias = j;
                            unique[k] = alias < 0;
                            wf.mip[k] = kinds[k].mip;
                            if (alias >= 0) wf.s[k] = wf.s[alias];
                            else { wf.s[k].tasks = bufs[k]; wf.s[k].count = p->wfCounters + 2 * k; wf.s[k].cursor = p->wfCounters + 2 * k + 1; wf.s[k].capacity = (unsigned)(2 * n); }
                        }
                        // aliased streams can receive up to 4 tasks per pixel: they own the whole remaining buffer
                        if (!unique[1] && !unique[2] && !unique[3]) wf.s[0].capacity = wf.s[1].capacity = wf.s[2].capacity = wf.s[3].capacity = (unsigned)(12 * n);
                        else if (!unique[2] && !unique[3]) { wf.s[0].capacity = wf.s[2].capacity = (unsigned)(2 * n); wf.s[1].capacity = wf.s[3].capacity = (unsigned)(2 * n); }
                        CK(cudaMemsetAsync(p->wfCounters, 0, 32, st));
                        CK(launchTemporalGather(fp, wf, st));
                        for (int k = 0; k < 4; k++)
                            if (unique[k]) { CK(launchMarch(wf.s[k], wf.results, kinds[k], p->scene.slots[kinds[k].mip], 1, p->marchBlocks1, st)); p->launches++; }
                        CK(launchTemporalCombine(fp, wf, st));
                        p->launches += 2;
                    } else if (genericStageOk(p, 2)) { rc = runStageGeneric(p, 2, fp, st); if (rc) return rc; }
                    else { CK(launchTemporal(fp, st)); p->launches++; }
                }
                p->finalPhys = p->ia;
            }
            recordEv(p, 3, st);
            break;
        case 3:
            if (active && !m.mUseReference && m.mEnableSpatialReuse) {
                const int in = (arg % 2 == 0) ? p->ia : p->ib, out = (arg % 2 == 0) ? p->ib : p->ia;
                fp.cur = resView(p, in); fp.out = resView(p, out); fp.extCur = p->ext[in]; fp.extOut = p->ext[out];
                fp.roundId = arg;
                const int r2TimeSeed = ((m.mSpatialReuseRounds + 1) * p->mFrameCount + arg) % 16;   // VR/SpatialReuse.cs.slang:109
                for (int s = 0; s < m.mSpatialSampleCount; s++) fp.offsets[s] = neighborOffset(m, s, r2TimeSeed);
                if (wavefrontSpatialOk(p)) {
                    rc = ensureWavefront(p); if (rc) return rc;
                    const WfBufs wf = wfView(p);
                    CK(cudaMemsetAsync(p->wfCounters, 0, 16, st));
                    CK(launchSpatialGather(fp, wf, st));
                    MarchKind kc, kl;
                    wavefrontKinds(p, kc, kl);
                    for (auto& e : p->evMarch) if (!e) CK(cudaEventCreate(&e));
                    CK(cudaEventRecord(p->evMarch[0], st));
                    CK(launchMarch(wf.cam, wf.results, kc, p->scene.slots[kc.mip], 3, p->marchBlocks3, st));
                    CK(cudaEventRecord(p->evMarch[1], st));
                    CK(launchMarch(wf.light, wf.results, kl, p->scene.slots[kl.mip], 1, p->marchBlocks1, st));
                    CK(cudaEventRecord(p->evMarch[2], st));
                    p->evMarchValid = true; p->marchChunksUsed = 0;
                    CK(cudaMemcpyAsync(p->wfCounters + 8, p->wfCounters, 4, cudaMemcpyDeviceToDevice, st));       // task counts of this round (diagnostics)
                    CK(cudaMemcpyAsync(p->wfCounters + 9, p->wfCounters + 2, 4, cudaMemcpyDeviceToDevice, st));
                    CK(launchSpatialCombine(fp, wf, st));
                    p->launches += 4;
                } else if (genericStageOk(p, 3)) { rc = runStageGeneric(p, 3, fp, st); if (rc) return rc; }
                else { CK(launchSpatial(fp, st)); p->launches++; }
                p->finalPhys = out;
            }
            if (!(m.mEnableSpatialReuse && arg + 1 < m.mSpatialReuseRounds)) recordEv(p, 4, st);
            break;
        case 4:
            // K4 CopyReservoirs / copyResource(temporal <- cur): the final buffer *becomes* the history, the old history
            // takes over its ping-pong role.  No bytes move.
            if (active && !m.mUseReference && m.mEnableTemporalReuse) {
                const int f = p->finalPhys, oldT = p->it;
                if (f == p->ia) p->ia = oldT; else if (f == p->ib) p->ib = oldT;
                p->it = f;
            }
            recordEv(p, 5, st);
            break;
        case 5:
            fp.cur = resView(p, p->finalPhys); fp.extCur = p->ext[p->finalPhys];
            if (p->mFreezeFrame) fp.frameCount = p->mFrameCount - 1;
            if (!out_color) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "out_color is null");
            if (p->mUseWavefront && m.mMaxBounces == 1 && !m.mUseReference && !m.mVisualizeTotalTransmittance &&
                m.mFinalVisibilityTrackingMethod == VRESTIR_ANALYTIC_TRACKING && m.mFinalLightTrackingMethod == VRESTIR_ANALYTIC_TRACKING) {
                rc = ensureFinalScratch(p); if (rc) return rc;
                if (!p->analyticBlocks) { int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p->device); p->analyticBlocks = sms * analyticBlocksPerSM(); }
                const bool deferred = p->mPipelineFrames >= 2 && active;
                cudaStream_t so = st;
                if (deferred) {
                    if (!p->outStream) {
                        CK(cudaStreamCreateWithFlags(&p->outStream, cudaStreamNonBlocking));
                        CK(cudaEventCreateWithFlags(&p->evOutGo, cudaEventDisableTiming));
                        for (auto& e : p->evOutDone) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                        CK(cudaEventCreate(&p->evOut0)); CK(cudaEventCreate(&p->evOut1));
                    }
                    so = p->outStream;
                    CK(cudaEventRecord(p->evOutGo, st));
                    CK(cudaStreamWaitEvent(so, p->evOutGo, 0));
                    CK(cudaEventRecord(p->evOut0, so));
                }
                WfStream s; s.tasks = p->k5Tasks; s.count = p->k5Counters; s.cursor = p->k5Counters + 1; s.capacity = (unsigned)(2 * p->k5Pixels);
                CK(cudaMemsetAsync(p->k5Counters, 0, 8, so));
                CK(launchFinalGather(fp, s, p->k5Results, so));
                const MarchKind k = {0, 1, m.mFinalTStepScale, 0, {0.f, 0.f, 0.f}};
                CK(launchMarchAnalytic(s, p->k5Results, k, p->scene.slots[0], p->analyticBlocks, so));
                CK(launchFinalCombine(fp, p->k5Results, so));
                p->launches += 3;
                if (deferred) {
                    CK(cudaEventRecord(p->evOut1, so));
                    CK(cudaEventRecord(p->evOutDone[p->outSeq & 1], so));
                    p->outSeq++; p->outTimed = true;
                }
            } else if (genericStageOk(p, 5)) { rc = runStageGeneric(p, 5, fp, st); if (rc) return rc; }
            else { CK(launchFinal(fp, st)); p->launches++; }
            recordEv(p, 6, st);
            break;
        case 6: {   // VR/VolumetricReSTIR.cpp:765-772 (+ :636 feature history, as a swap)
            if (active && !m.mUseReference && m.mEnableTemporalReuse) p->featSwapPending = true;   // the swap happens when the next active frame starts
            p->mTemporalSampleAccumulated = 1;
            memcpy(p->prevCam.prevView, p->cam.viewMat, 64); memcpy(p->prevCam.prevProj, p->cam.projMat, 64);
            p->prevCam.prevU = fp.camU; p->prevCam.prevV = fp.camV; p->prevCam.prevW = fp.camW; p->prevCam.prevPos = fp.camPos;
            if (!p->mFreezeFrame) p->mFrameCount++;
            break;
        }
        default: return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad stage");
    }
    return VRESTIR_OK;
}

// F/Utils/Sampling/AliasTable.cpp:46-126 (std::mt19937 seeded with 123 as in EmissivePowerSampler.cpp:76)
void buildAliasTable(std::vector<float> weights, std::vector<uint32_t>& itemsOut, float& weightSumOut) {
    const uint32_t count = (uint32_t)weights.size();
    std::mt19937 rng(123);
    double weightSum = 0.0;
    for (float f : weights) weightSum += f;
    double factor = count / weightSum;
    for (float& f : weights) f = (float)(f * factor);
    std::vector<uint32_t> permutation(count);
    for (uint32_t i = 0; i < count; ++i) permutation[i] = i;
    std::sort(permutation.begin(), permutation.end(), [&](uint32_t a, uint32_t b) { return weights[a] < weights[b]; });
    std::vector<float> thresholds(count);
    std::vector<uint32_t> redirect(count);
    uint32_t head = 0, tail = count - 1;
    // count == 1: the loop does not run and the single item stays {threshold 0, indexA 0, indexB 0}, as in the reference
    while (head != tail) {
        int i = permutation[head], j = permutation[tail];
        thresholds[i] = weights[i];
        redirect[i] = j;
        weights[j] -= 1.f - weights[i];
        if (head == tail - 1) { thresholds[j] = 1.f; redirect[j] = j; break; }
        else if (weights[j] < 1.f) { std::swap(permutation[head], permutation[tail]); tail--; }
        else head++;
    }
    for (uint32_t i = 0; i < count; ++i) permutation[i] = i;
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t dst = i + ((uint32_t)rng() % (count - i));
        std::swap(thresholds[i], thresholds[dst]); std::swap(redirect[i], redirect[dst]); std::swap(permutation[i], permutation[dst]);
    }
    itemsOut.resize((size_t)count * 4);
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t tb; memcpy(&tb, &thresholds[i], 4);
        itemsOut[i * 4 + 0] = tb; itemsOut[i * 4 + 1] = redirect[i]; itemsOut[i * 4 + 2] = permutation[i]; itemsOut[i * 4 + 3] = 0;
    }
    weightSumOut = (float)weightSum;
}

// Vose alias table over the finest importance mip (env-map alias sampling, north star); threshold = P(keep own texel)
void buildEnvAlias(const float* w, uint32_t count, std::vector<float>& thr, std::vector<uint32_t>& redirect) {
    thr.assign(count, 1.f); redirect.resize(count);
    double sum = 0; for (uint32_t i = 0; i < count; i++) sum += w[i];
    std::vector<double> q(count);
    std::vector<uint32_t> small, large;
    for (uint32_t i = 0; i < count; i++) { q[i] = sum > 0 ? (double)w[i] * count / sum : 1.0; redirect[i] = i; (q[i] < 1.0 ? small : large).push_back(i); }
    while (!small.empty() && !large.empty()) {
        uint32_t s = small.back(); small.pop_back();
        uint32_t l = large.back();
        thr[s] = (float)q[s]; redirect[s] = l;
        q[l] -= 1.0 - q[s];
        if (q[l] < 1.0) { large.pop_back(); small.push_back(l); }
    }
    for (uint32_t i : small) thr[i] = 1.f;
    for (uint32_t i : large) thr[i] = 1.f;
}

}  // namespace

// ================================================================================================ C API
extern "C" {

const char* vrestir_last_error(void) { return vr::g_lastError.c_str(); }
const char* vrestir_version(void) { return "vrestir-b200 0.1 (sm_100a)"; }

void vrestir_default_params(vrestir_params* o) {
    if (!o) return;
    memset(o, 0, sizeof(*o));
    o->mMaxBounces = 1; o->mEnableTemporalReuse = 1; o->mEnableSpatialReuse = 1; o->mVertexReuse = 0; o->mVertexReuseStartBounce = 1; o->mUseReference = 0;
    o->mUseEnvironmentLights = 1; o->mUseAnalyticLights = 0; o->mUseEmissiveLights = 0; o->mBaselineSamplePerPixel = 1; o->mVisualizeTotalTransmittance = 0;
    o->mUseSurfaceScene = 0; o->mUsePrevVolumeForReproj = 1;
    o->mInitialBaseMipLevel = 1; o->mInitialM = 4; o->mInitialLightSamples = 1; o->mInitialLightingMipLevel = 2; o->mInitialVisibilityUseLinearSampler = 0;
    o->mInitialLightingUseLinearSampler = 1; o->mInitialLightingTrackingMethod = VRESTIR_RAY_MARCHING; o->mInitialVisibilityTStepScale = 1.f;
    o->mInitialLightingTStepScale = 2.f; o->mInitialUseRussianRoulette = 1; o->mInitialUseCoarserGridForIndirectBounce = 1;
    o->mTemporalReuseMThreshold = 4.f; o->mTemporalReprojectionMode = VRESTIR_REPROJECTION_LINEAR; o->mTemporalMISMethod = VRESTIR_MIS_TALBOT; o->mTemporalReprojectionMipLevel = 1;
    o->mSpatialReuseRounds = 1; o->mSpatialVisibilityMipLevel = 1; o->mSpatialLightingMipLevel = 1; o->mSpatialVisibilityUseLinearSampler = 1; o->mSpatialLightingUseLinearSampler = 1;
    o->mSpatialVisibilityTStepScale = 1.f; o->mSpatialLightingTStepScale = 1.f; o->mSpatialVisibilityTrackingMethod = VRESTIR_RAY_MARCHING;
    o->mSpatialLightingTrackingMethod = VRESTIR_RAY_MARCHING; o->mRandomSamplerType = VRESTIR_SAMPLER_R2; o->mSampleRadius = 10.f; o->mSpatialSampleCount = 4;
    o->mEnableVisibilitySimilarityRejection = 0; o->mSpatialMISMethod = VRESTIR_MIS_TALBOT;
    o->mFinalLightSamples = 1; o->mFinalVisibilitySamples = 1; o->mFinalVisibilityTrackingMethod = VRESTIR_ANALYTIC_TRACKING; o->mFinalLightTrackingMethod = VRESTIR_ANALYTIC_TRACKING;
    o->mFinalRandomSamplerType = VRESTIR_SAMPLER_R2; o->mFinalTStepScale = 0.2f;
}

int vrestir_create(const vrestir_params* params, int device, vrestir_pass** out) try {
    if (!out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null out");
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) return setError(VRESTIR_ERR_CUDA, std::string("no CUDA device: the VolumetricReSTIR pass has no CPU fallback (") + cudaGetErrorString(e) + ")");
    if (device < 0 || device >= count) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad device index");
    CK(cudaSetDevice(device));
    auto* p = new vrestir_pass();
    p->device = device;
    if (params) p->P = *params; else vrestir_default_params(&p->P);
    memset(&p->scene, 0, sizeof(p->scene));
    registerPass(p);
    *out = p;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_destroy(vrestir_pass* p) try {
    if (!p) return VRESTIR_OK;
    cudaSetDevice(p->device);
    cudaDeviceSynchronize();
    unregisterPass(p);
    if (p->evMainTail) cudaEventDestroy(p->evMainTail);
    for (auto& d : p->dslots) freeSlot(d);
    for (auto& fr : p->volumeFrames) for (auto& d : fr->d) freeSlot(d);
    slotCacheTrim();
    for (int i = 0; i < 4; i++) if (p->res[i]) cudaFree(p->res[i]);
    for (int i = 0; i < 4; i++) if (p->ext[i]) cudaFree(p->ext[i]);
    for (int i = 0; i < 3; i++) if (p->feat[i]) cudaFree(p->feat[i]);
    void* ptrs[] = {p->refColor, p->d_lut, p->d_lutPrev, p->d_env, p->d_importance, p->d_envAliasThr, p->d_envAliasRedirect, p->d_lights, p->d_tris, p->d_alias, p->d_aliasWeights, p->d_hostColor[0], p->d_hostColor[1], p->d_hostMvec[0], p->d_hostMvec[1], p->wfCamTasks, p->wfLightTasks, p->wfResults, p->wfCounters, p->wfInitialState, p->k1LightTasks, p->k1EvalTasks, p->k1Results, p->k1Counters, p->k5Tasks, p->k5Results, p->k5Counters, p->mbArena, p->mbCounters, p->d_marchCounts, p->k1mb.base, p->k1mb.counters, p->k1mbEval.base, p->k1mbEval.counters};
    for (void* q : ptrs) if (q) cudaFree(q);
    for (auto& e : p->ev) if (e) cudaEventDestroy(e);
    if (p->hostStream) cudaStreamDestroy(p->hostStream);
    if (p->hostCopyStream) cudaStreamDestroy(p->hostCopyStream);
    for (cudaEvent_t e : {p->evHostRendered[0], p->evHostRendered[1], p->evHostLanded[0], p->evHostLanded[1]}) if (e) cudaEventDestroy(e);
    if (p->auxStream) cudaStreamDestroy(p->auxStream);
    if (p->evFork) cudaEventDestroy(p->evFork);
    if (p->evJoin) cudaEventDestroy(p->evJoin);
    for (auto& e : p->evMarch) if (e) cudaEventDestroy(e);
    for (auto& e : p->evMarchChunks) if (e) cudaEventDestroy(e);
    if (p->outStream) cudaStreamDestroy(p->outStream);
    for (cudaEvent_t e : {p->evOutGo, p->evOutDone[0], p->evOutDone[1], p->evOut0, p->evOut1}) if (e) cudaEventDestroy(e);
    if (p->pfStream) cudaStreamDestroy(p->pfStream);
    for (cudaEvent_t e : {p->evPfGo, p->evPfDone, p->evPf0, p->evPf1}) if (e) cudaEventDestroy(e);
    delete p;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_volume(vrestir_pass* p, const vrestir_grid_desc* g) try {
    if (!p || !g) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    if (!g->slots[0].valid) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "slot 0 (density mip 0) must be valid");
    for (int s = 0; s < VRESTIR_MAX_SLOTS; s++) { int rc = uploadSlot(p, s, g->slots[s]); if (rc) return rc; }
    p->volBase = g->volume;
    if (p->d_lut) { cudaFree(p->d_lut); p->d_lut = nullptr; }
    if (g->blackbody_lut) { CK(cudaMalloc(&p->d_lut, 2048)); CK(cudaMemcpy(p->d_lut, g->blackbody_lut, 2048, cudaMemcpyHostToDevice)); }
    p->scene.lut = (const float4*)p->d_lut;
    p->haveVolume = true; p->sceneDirty = true; p->mOptionsChanged = true; p->persistBase = p->persistBasePf = nullptr;
    applyOverrides(p);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// The current density / temperature / velocity grids become the previous frame's (F/Scene/Scene.cpp:825-863).  Memory owned by
// the displaced previous-frame slots is released, so everything in flight has to be done first; borrowed (resident) slots just
// change hands.
static int shiftSlotsToPrev(vrestir_pass* p) {
    auto prevOf = [](int s) { return s < VRESTIR_NUM_MAX_MIPS ? VRESTIR_PREV_DENSITY_GRID_OFFSET + s : s + VRESTIR_PREV_EXTRA_GRID_OFFSET; };
    const int from[] = {VRESTIR_TEMPERATURE_GRID_ID, VRESTIR_VELOCITY_GRID_ID};
    bool frees = false;
    for (int i = 0; i < VRESTIR_NUM_MAX_MIPS; i++) if (prevOf(i) < VRESTIR_MAX_SLOTS) frees |= ownsMemory(p->dslots[prevOf(i)]);
    for (int f : from) frees |= ownsMemory(p->dslots[prevOf(f)]);
    if (frees) CK(cudaDeviceSynchronize());
    auto moveSlot = [&](int a, int b) {
        freeSlot(p->dslots[b]);
        p->dslots[b] = p->dslots[a]; p->dslots[a] = DevSlot{};
        p->scene.slots[b] = p->scene.slots[a]; memset(&p->scene.slots[a], 0, sizeof(DSlot));
    };
    for (int i = 0; i < VRESTIR_NUM_MAX_MIPS; i++) if (prevOf(i) < VRESTIR_MAX_SLOTS) moveSlot(i, prevOf(i));
    for (int f : from) moveSlot(f, prevOf(f));
    return VRESTIR_OK;
}

int vrestir_advance_volume(vrestir_pass* p, const vrestir_grid_desc* g) try {
    if (!p || !g || !p->haveVolume) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "advance_volume before set_volume");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    { int rc = shiftSlotsToPrev(p); if (rc) return rc; }
    const int lastHasEmission = p->volBase.hasEmission;
    for (int s = 0; s < VRESTIR_PREV_DENSITY_GRID_OFFSET - 1; s++) { int rc = uploadSlot(p, s, g->slots[s]); if (rc) return rc; }
    p->volBase = g->volume; p->volBase.lastFrameHasEmission = lastHasEmission; p->volBase.hasAnimation = 1;
    if (g->blackbody_lut && !p->d_lut) { CK(cudaMalloc(&p->d_lut, 2048)); CK(cudaMemcpy(p->d_lut, g->blackbody_lut, 2048, cudaMemcpyHostToDevice)); p->scene.lut = (const float4*)p->d_lut; }
    p->sceneDirty = true; p->persistBase = p->persistBasePf = nullptr;
    applyOverrides(p);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Resident animation frames.  The reference keeps every frame of an animated sequence on the GPU and switches the bound grids
// and the volume description per frame (F/Scene/Scene.cpp:825-863, mVolumeDescArray[mVDBAnimationFrameId]); these two calls
// are that protocol: add uploads a frame once, advance_resident rebinds pointers (no copy, no allocation, no device-wide wait).
int vrestir_volume_frame_add(vrestir_pass* p, const vrestir_grid_desc* g, int* out_index) try {
    if (!p || !g) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (!g->slots[0].valid) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "slot 0 (density mip 0) must be valid");
    CK(cudaSetDevice(p->device));
    auto fr = std::make_unique<vrestir_pass::ResidentFrame>();
    for (int s = 0; s < VRESTIR_PREV_DENSITY_GRID_OFFSET - 1; s++) {
        int rc = uploadSlotTo(fr->d[s], fr->s[s], g->slots[s]);
        if (rc) { for (auto& d : fr->d) freeSlot(d); return rc; }
    }
    CK(cudaDeviceSynchronize());
    fr->vol = g->volume;
    if (g->blackbody_lut && !p->d_lut) { CK(cudaMalloc(&p->d_lut, 2048)); CK(cudaMemcpy(p->d_lut, g->blackbody_lut, 2048, cudaMemcpyHostToDevice)); p->scene.lut = (const float4*)p->d_lut; p->sceneDirty = true; }
    p->volumeFrames.push_back(std::move(fr));
    if (out_index) *out_index = (int)p->volumeFrames.size() - 1;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_advance_volume_resident(vrestir_pass* p, int index) try {
    if (!p || !p->haveVolume) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "advance_volume_resident before set_volume");
    if (index < 0 || index >= (int)p->volumeFrames.size()) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "no resident frame with this index");
    CK(cudaSetDevice(p->device));
    { int rc = shiftSlotsToPrev(p); if (rc) return rc; }
    const vrestir_pass::ResidentFrame& fr = *p->volumeFrames[index];
    bool frees = false;
    for (int s = 0; s < VRESTIR_PREV_DENSITY_GRID_OFFSET - 1; s++) frees |= ownsMemory(p->dslots[s]);
    if (frees) CK(cudaDeviceSynchronize());
    for (int s = 0; s < VRESTIR_PREV_DENSITY_GRID_OFFSET - 1; s++) {
        freeSlot(p->dslots[s]);
        p->dslots[s] = fr.d[s]; p->dslots[s].borrowed = true;
        p->scene.slots[s] = fr.s[s];
    }
    const int lastHasEmission = p->volBase.hasEmission;
    p->volBase = fr.vol; p->volBase.lastFrameHasEmission = lastHasEmission; p->volBase.hasAnimation = 1;
    p->sceneDirty = true; p->persistBase = p->persistBasePf = nullptr;
    applyOverrides(p);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Releases the resident frames; slots still bound to one of them are unbound first (a volume must be set again before rendering).
int vrestir_volume_frames_clear(vrestir_pass* p) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->volumeFrames.empty()) return VRESTIR_OK;
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    bool unbound = false;
    for (int s = 0; s < VRESTIR_MAX_SLOTS; s++) if (p->dslots[s].borrowed) { p->dslots[s] = DevSlot{}; memset(&p->scene.slots[s], 0, sizeof(DSlot)); unbound = true; }
    for (auto& fr : p->volumeFrames) for (auto& d : fr->d) freeSlot(d);
    p->volumeFrames.clear();
    if (unbound) { p->haveVolume = false; p->sceneDirty = true; p->persistBase = p->persistBasePf = nullptr; }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Device-resident volume update: the density slots come from a GPU-built mip chain, everything else
// (transforms, volume description, temperature / velocity grids) from `tmpl`, the grid description of a host-built volume
// of the same dimensions (typically frame 0 of the sequence).  Only the brick-activity maps (1 byte per brick) visit the
// host, where the tree over them is built; brick pools, quad repacks and brick bounds are produced on the device.
int vrestir_set_volume_from_chain(vrestir_pass* p, const vrestir_mip_chain* chain, const vrestir_grid_desc* tmpl, int advance) try {
    if (!p || !chain || !tmpl) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (vr::chainDevice(chain) != p->device) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "the chain lives on another device");
    if (advance && !p->haveVolume) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "advance before a volume was set");
    if (!tmpl->slots[0].valid) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "template slot 0 (density mip 0) must be valid");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    if (advance) { int rc = shiftSlotsToPrev(p); if (rc) return rc; }   // like vrestir_advance_volume
    int built = 0;
    if (vrestir_mips_count(chain, &built)) return VRESTIR_ERR_INVALID_ARGUMENT;
    const int lastSlot = advance ? VRESTIR_PREV_DENSITY_GRID_OFFSET - 1 : VRESTIR_MAX_SLOTS;
    for (int s = 0; s < lastSlot; s++) {
        const bool density = s < 2 * VRESTIR_NUM_MAX_MIPS;
        const int mip = s % VRESTIR_NUM_MAX_MIPS, cons = s / VRESTIR_NUM_MAX_MIPS;
        const vrestir_grid_slot& t = tmpl->slots[s];
        if (!density || !t.valid) { int rc = uploadSlot(p, s, t); if (rc) return rc; continue; }   // non-density grids: host data of the template
        if (mip >= built) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "the chain has fewer levels than the template volume");
        vr::ChainLevelView lv;
        int rc = vr::chainLevelView(chain, mip, cons, lv); if (rc) return rc;
        if (lv.dim[0] != (int)t.bmax[0] || lv.dim[1] != (int)t.bmax[1] || lv.dim[2] != (int)t.bmax[2] || lv.format != t.atlas_format || t.atlas_channels != 1)
            return setError(VRESTIR_ERR_INVALID_ARGUMENT, "chain level and template slot differ in size or format");
        const int BX = (lv.dim[0] + 7) / 8, BY = (lv.dim[1] + 7) / 8, BZ = (lv.dim[2] + 7) / 8;
        std::vector<uint8_t> active((size_t)BX * BY * BZ);
        CK(cudaMemcpy(active.data(), lv.active, active.size(), cudaMemcpyDeviceToHost));
        vr::Topology topo;
        vr::buildTopology(active, lv.dim[0], lv.dim[1], lv.dim[2], topo);
        vrestir_grid_slot g = t;   // dims, res, vdel, bounds, transform of the template
        g.top_lev = topo.topLev;
        for (int l = 0; l < 3; l++) {
            g.node_count[l] = (uint32_t)topo.nodes[l].size(); g.nodes[l] = topo.nodes[l].empty() ? nullptr : topo.nodes[l].data();
            g.childlist[l] = topo.child[l].empty() ? nullptr : topo.child[l].data(); g.childlist_count[l] = topo.child[l].size();
        }
        g.brick_count = topo.brickCount; g.atlas = nullptr;
        g.max_value = lv.maxValue; g.compress_scale = lv.format == VRESTIR_ATLAS_UNORM8 ? lv.maxValue : 1.f;
        rc = uploadSlot(p, s, g); if (rc) return rc;
        DevSlot& d = p->dslots[s];
        CK(vr::launchPackBricks(lv.data, lv.dim, lv.format, (const vrestir_node*)d.nodes[0], topo.brickCount, d.atlas, 0));
        CK(vr::launchBrickBounds(d.atlas, lv.format, lv.maxValue, (vrestir_node*)d.nodes[0], topo.brickCount, 0));
        if (lv.format == VRESTIR_ATLAS_UNORM8) {
            d.quadBytes = (size_t)topo.brickCount * 810 * 4;
            CK(slotAlloc(&d.quads, d.quadBytes));
            CK(vr::launchQuadRepack((const uint8_t*)d.atlas, topo.brickCount, (uint32_t*)d.quads, 0));
            p->scene.slots[s].quads = (const uint32_t*)d.quads;
        }
        p->launches += 3;
    }
    CK(cudaDeviceSynchronize());
    const int lastHasEmission = p->volBase.hasEmission;
    p->volBase = tmpl->volume;
    if (advance) { p->volBase.lastFrameHasEmission = lastHasEmission; p->volBase.hasAnimation = 1; }
    if (!advance || !p->d_lut) {
        if (p->d_lut) { cudaFree(p->d_lut); p->d_lut = nullptr; }
        if (tmpl->blackbody_lut) { CK(cudaMalloc(&p->d_lut, 2048)); CK(cudaMemcpy(p->d_lut, tmpl->blackbody_lut, 2048, cudaMemcpyHostToDevice)); }
        p->scene.lut = (const float4*)p->d_lut;
    }
    p->haveVolume = true; p->sceneDirty = true; p->persistBase = p->persistBasePf = nullptr;
    if (!advance) p->mOptionsChanged = true;
    applyOverrides(p);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Host copy of the volume as it is bound on the device (nodes with the device-computed brick bounds, child lists, brick pools):
// what a checker needs to evaluate the same grid on the CPU when the grid was built on the device (vrestir_set_volume_from_chain).
int vrestir_download_volume(vrestir_pass* p, vrestir_scene** out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (!p->haveVolume) return setError(VRESTIR_ERR_NOT_READY, "no volume bound");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    vrestir_scene* sc = vr::newHostScene(p->volBase);
    for (int slot = 0; slot < VRESTIR_MAX_SLOTS; slot++) {
        const DevSlot& d = p->dslots[slot];
        if (!p->scene.slots[slot].valid) continue;
        vr::HostSlotVectors h = vr::hostSlotVectors(sc, slot);
        *h.desc = d.meta;
        for (int l = 0; l < 3; l++) {
            h.nodes[l]->resize(d.meta.node_count[l]);
            if (d.meta.node_count[l] && d.nodes[l]) CK(cudaMemcpy(h.nodes[l]->data(), d.nodes[l], (size_t)d.meta.node_count[l] * sizeof(vrestir_node), cudaMemcpyDeviceToHost));
            h.child[l]->resize(d.meta.childlist_count[l]);
            if (d.meta.childlist_count[l] && d.child[l]) CK(cudaMemcpy(h.child[l]->data(), d.child[l], (size_t)d.meta.childlist_count[l] * 4, cudaMemcpyDeviceToHost));
            h.desc->nodes[l] = h.nodes[l]->empty() ? nullptr : h.nodes[l]->data();
            h.desc->childlist[l] = h.child[l]->empty() ? nullptr : h.child[l]->data();
        }
        h.atlas->resize(d.atlasBytes);
        if (d.atlasBytes) CK(cudaMemcpy(h.atlas->data(), d.atlas, d.atlasBytes, cudaMemcpyDeviceToHost));
        h.desc->atlas = h.atlas->empty() ? nullptr : h.atlas->data();
    }
    if (p->d_lut) vr::attachBlackbodyLut(sc);
    *out = sc;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_camera(vrestir_pass* p, const vrestir_camera* cam) try {
    if (!p || !cam) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    p->cam = *cam; p->haveCamera = true;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_envmap(vrestir_pass* p, const vrestir_envmap_desc* env) try {
    if (!p || !env || !env->texels || env->width < 1 || env->height < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad env map");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    if (p->d_env) cudaFree(p->d_env);
    const size_t bytes = (size_t)env->width * env->height * 16;
    CK(cudaMalloc(&p->d_env, bytes));
    CK(cudaMemcpy(p->d_env, env->texels, bytes, cudaMemcpyHostToDevice));
    DScene& s = p->scene;
    s.haveEnv = 1; s.envW = env->width; s.envH = env->height; s.envTexels = (const float4*)p->d_env;
    s.envIntensity = env->intensity; s.envTint = make_float3(env->tint[0], env->tint[1], env->tint[2]);
    memcpy(s.envT, env->transform, 36); memcpy(s.envInvT, env->invTransform, 36); memcpy(s.envPrevT, env->prevTransform, 36); memcpy(s.envPrevInvT, env->prevInvTransform, 36);
    // K6: importance map 512^2, 64 spp/texel, + mip chain (EnvMapSampler.cpp:83-116)
    const int dim = 512, spp = 64;
    int mips = 0; for (int d = dim; d >= 1; d >>= 1) mips++;
    size_t total = 0; for (int i = 0; i < mips; i++) { s.impOffset[i] = (unsigned)total; total += (size_t)(dim >> i) * (dim >> i); }
    s.impDim = dim; s.impBaseMip = mips - 1;
    if (p->d_importance) cudaFree(p->d_importance);
    CK(cudaMalloc(&p->d_importance, total * 4 + 16));   // + slack: the top of the chain is staged with 16-byte bulk copies
    p->importanceCount = total;
    s.importance = p->d_importance;
    CK(uploadScene(s, 0)); CK(cudaDeviceSynchronize());
    const int sx = std::max(1, (int)std::sqrt((double)spp)), sy = spp / sx;
    CK(launchImportance(p->d_importance, dim, sx, sy, 0)); p->launches++;
    for (int mip = 1; mip < mips; mip++) { CK(launchImportanceMip(p->d_importance + s.impOffset[mip - 1], p->d_importance + s.impOffset[mip], dim >> mip, 0)); p->launches++; }
    CK(cudaDeviceSynchronize());
    // alias table over the finest mip (used when mEnvSamplerType == VRESTIR_ENV_SAMPLER_ALIAS)
    std::vector<float> base((size_t)dim * dim);
    CK(cudaMemcpy(base.data(), p->d_importance, base.size() * 4, cudaMemcpyDeviceToHost));
    buildEnvAlias(base.data(), (uint32_t)base.size(), p->envAliasThr, p->envAliasRedirect);
    if (p->d_envAliasThr) cudaFree(p->d_envAliasThr);
    if (p->d_envAliasRedirect) cudaFree(p->d_envAliasRedirect);
    CK(cudaMalloc(&p->d_envAliasThr, base.size() * 4)); CK(cudaMalloc(&p->d_envAliasRedirect, base.size() * 4));
    CK(cudaMemcpy(p->d_envAliasThr, p->envAliasThr.data(), base.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(p->d_envAliasRedirect, p->envAliasRedirect.data(), base.size() * 4, cudaMemcpyHostToDevice));
    s.envAliasThr = p->d_envAliasThr; s.envAliasRedirect = p->d_envAliasRedirect; s.envAliasCount = (unsigned)base.size();
    p->sceneDirty = true; p->mOptionsChanged = true;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_analytic_lights(vrestir_pass* p, const vrestir_light* lights, int count) try {
    if (!p || count < 0 || (count > 0 && !lights)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad lights");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    for (int i = 0; i < count; i++) if (lights[i].type > VRESTIR_LIGHT_DIRECTIONAL) return setError(VRESTIR_ERR_UNSUPPORTED, "only point and directional analytic lights are supported");
    if (p->d_lights) { cudaFree(p->d_lights); p->d_lights = nullptr; }
    if (count) { CK(cudaMalloc(&p->d_lights, (size_t)count * sizeof(vrestir_light))); CK(cudaMemcpy(p->d_lights, lights, (size_t)count * sizeof(vrestir_light), cudaMemcpyHostToDevice)); }
    p->scene.lights = (const vrestir_light*)p->d_lights; p->scene.lightCount = count;
    p->sceneDirty = true; p->mOptionsChanged = true;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_emissive_triangles(vrestir_pass* p, const vrestir_emissive_triangle* tris, int count, float mul) try {
    if (!p || count < 0 || (count > 0 && !tris)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad triangles");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    void** ptrs[] = {&p->d_tris, &p->d_alias, &p->d_aliasWeights};
    for (void** q : ptrs) if (*q) { cudaFree(*q); *q = nullptr; }
    p->aliasItems.clear(); p->aliasWeights.clear(); p->aliasWeightSum = 0.f;
    if (count) {
        p->aliasWeights.resize(count);
        for (int i = 0; i < count; i++) {   // flux = luminance(Le) * area * pi (F/Experimental/Scene/Lights/FinalizeIntegration.cs.slang:73)
            float lum = 0.2126f * tris[i].Le[0] + 0.7152f * tris[i].Le[1] + 0.0722f * tris[i].Le[2];
            p->aliasWeights[i] = lum * tris[i].area * 3.14159265358979323846f;
        }
        buildAliasTable(p->aliasWeights, p->aliasItems, p->aliasWeightSum);
        CK(cudaMalloc(&p->d_tris, (size_t)count * sizeof(vrestir_emissive_triangle)));
        CK(cudaMemcpy(p->d_tris, tris, (size_t)count * sizeof(vrestir_emissive_triangle), cudaMemcpyHostToDevice));
        CK(cudaMalloc(&p->d_alias, (size_t)count * 16)); CK(cudaMemcpy(p->d_alias, p->aliasItems.data(), (size_t)count * 16, cudaMemcpyHostToDevice));
        CK(cudaMalloc(&p->d_aliasWeights, (size_t)count * 4)); CK(cudaMemcpy(p->d_aliasWeights, p->aliasWeights.data(), (size_t)count * 4, cudaMemcpyHostToDevice));
    }
    DScene& s = p->scene;
    s.tris = (const vrestir_emissive_triangle*)p->d_tris; s.triCount = count; s.alias = (const uint4*)p->d_alias; s.aliasWeights = (const float*)p->d_aliasWeights;
    s.aliasWeightSum = p->aliasWeightSum; s.emissiveMul = mul;
    p->sceneDirty = true; p->mOptionsChanged = true;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_get_emissive_alias(const vrestir_pass* p, uint32_t* items, float* weights, float* weight_sum) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (items && !p->aliasItems.empty()) memcpy(items, p->aliasItems.data(), p->aliasItems.size() * 4);
    if (weights && !p->aliasWeights.empty()) memcpy(weights, p->aliasWeights.data(), p->aliasWeights.size() * 4);
    if (weight_sum) *weight_sum = p->aliasWeightSum;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_build_alias_table(const float* weights, int count, uint32_t* items4, float* weight_sum) try {
    if (!weights || count < 1 || !items4) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    std::vector<uint32_t> items; float ws = 0.f;
    buildAliasTable(std::vector<float>(weights, weights + count), items, ws);
    memcpy(items4, items.data(), items.size() * 4);
    if (weight_sum) *weight_sum = ws;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_build_env_alias(const float* texel_weights, int count, float* thresholds, uint32_t* redirect) try {
    if (!texel_weights || count < 1 || !thresholds || !redirect) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    std::vector<float> thr; std::vector<uint32_t> red;
    buildEnvAlias(texel_weights, (uint32_t)count, thr, red);
    memcpy(thresholds, thr.data(), thr.size() * 4); memcpy(redirect, red.data(), red.size() * 4);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_env_alias(const vrestir_pass* p, float* thresholds, uint32_t* redirect, int* count) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (count) *count = (int)p->envAliasThr.size();
    if (thresholds && !p->envAliasThr.empty()) memcpy(thresholds, p->envAliasThr.data(), p->envAliasThr.size() * 4);
    if (redirect && !p->envAliasRedirect.empty()) memcpy(redirect, p->envAliasRedirect.data(), p->envAliasRedirect.size() * 4);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_frame(vrestir_pass* p, int w, int h, int row_begin, int row_end) try {
    if (!p || w <= 0 || h <= 0 || row_begin < 0 || row_end > h || row_begin >= row_end) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad frame / row band");
    if ((long long)w * h > (1ll << 30)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "frame too large");
    if (p->W != w || p->H != h) p->mOptionsChanged = true;
    p->W = w; p->H = h; p->rowBegin = row_begin; p->rowEnd = row_end;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_update(vrestir_pass* p, const char* key, double value) try {
    if (!p || !key) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    std::string k(key);
    if (k.rfind("mParams.", 0) == 0) k = k.substr(8);
    bool found = false;
    for (const auto& kd : kKeys)
        if (k == kd.name) {
            char* base = (char*)&p->P + kd.off;
            if (kd.type == 0) *(int32_t*)base = (int32_t)value; else if (kd.type == 1) *(uint32_t*)base = (uint32_t)value; else *(float*)base = (float)value;
            found = true; break;
        }
    if (!found) {
        found = true;
        if (k == "mOutputMotionVec") p->mOutputMotionVec = value != 0;
        else if (k == "mFreezeFrame") p->mFreezeFrame = value != 0;
        else if (k == "volumeDensityScaleExtraControl") p->densityExtra = (float)value;
        else if (k == "volumeAlbedoExtraControl") p->albedoExtra = (float)value;
        else if (k == "volumeAnisotropyExtraControl") p->anisotropyExtra = (float)value;
        else if (k == "mEnvSamplerType") p->envSamplerType = (int)value;
        else if (k == "mUseWavefront") p->mUseWavefront = value != 0;
        else if (k == "mInitialMode") p->mInitialMode = (int)value;
        else if (k == "mOverlapFeatures") p->mOverlapFeatures = value != 0;
        else if (k == "mPipelineFrames") p->mPipelineFrames = (int)value;
        else if (k == "mPrimaryDistanceEngine") p->mPrimaryDistanceEngine = value != 0;
        else if (k == "mDebugPoisonResults") p->mDebugPoison = value != 0;
        else if (k == "mScratchBudgetMB") p->mScratchBudget = (size_t)std::max(1.0, value) << 20;
        else if (k == "mPrefetchPriority") p->mPrefetchPriority = value != 0;
        else if (k == "randomizeFrameSeed") { if (!p->mRandomizeFrameSeed) p->randState = 123; p->mRandomizeFrameSeed = true; }
        else found = false;
    }
    p->mOptionsChanged = true;   // VR/VolumetricReSTIR.cpp:1339
    p->sceneDirty = true;
    if (!found) return setError(VRESTIR_WARN_UNKNOWN_KEY, std::string("Unknown field '") + key + "' in a VolumetricReSTIR dictionary");
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_set_params(vrestir_pass* p, const vrestir_params* params) try {
    if (!p || !params) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    p->P = *params; p->mOptionsChanged = true; p->sceneDirty = true;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_params(const vrestir_pass* p, vrestir_params* out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *out = p->P; return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_set_frame_count(vrestir_pass* p, int fc, int acc) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    p->mFrameCount = fc; p->mTemporalSampleAccumulated = acc; p->mOptionsChanged = false; return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_frame_count(const vrestir_pass* p, int* fc) try {
    if (!p || !fc) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *fc = p->mFrameCount; return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
/* previous-frame camera for staged parity tests of K2 (normally saved by stage 6) */
int vrestir_set_prev_camera(vrestir_pass* p, const vrestir_camera* cam) try {
    if (!p || !cam) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    DPrevCam& s = p->prevCam;
    memcpy(s.prevView, cam->viewMat, 64); memcpy(s.prevProj, cam->projMat, 64);
    s.prevU = make_float3(cam->cameraU[0], cam->cameraU[1], cam->cameraU[2]); s.prevV = make_float3(cam->cameraV[0], cam->cameraV[1], cam->cameraV[2]);
    s.prevW = make_float3(cam->cameraW[0], cam->cameraW[1], cam->cameraW[2]); s.prevPos = make_float3(cam->posW[0], cam->posW[1], cam->posW[2]);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_execute_stage(vrestir_pass* p, int stage, int arg, float* out_color, float* out_mvec, void* stream) try {
    return runStage(p, stage, arg, out_color, out_mvec, (cudaStream_t)stream);
} catch (...) { return vr::caughtException(); }

int vrestir_execute(vrestir_pass* p, float* out_color, float* out_mvec, void* stream) try {
    cudaStream_t st = (cudaStream_t)stream;
    int rc;
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    // a matching prefetch already holds K0/K1 of this frame: stages 0/1 only adopt it (matching is re-checked there after the
    // options / scene bookkeeping of the frame start)
    const bool overlap = p->mOverlapFeatures && !p->P.mUseReference && !p->mFreezeFrame && !p->pfValid;
    p->framesOverlapped = overlap;
    if (overlap) {
        CK(cudaSetDevice(p->device));
        if (!p->auxStream) { CK(cudaStreamCreateWithFlags(&p->auxStream, cudaStreamNonBlocking)); CK(cudaEventCreateWithFlags(&p->evFork, cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&p->evJoin, cudaEventDisableTiming)); }
        if ((rc = runStage(p, -1, 0, out_color, out_mvec, st))) return rc;   // options / frame counter / scene constants, on the main stream
        CK(cudaEventRecord(p->evFork, st));
        CK(cudaStreamWaitEvent(p->auxStream, p->evFork, 0));
        if ((rc = runStage(p, 0, 0, out_color, out_mvec, p->auxStream))) return rc;
        CK(cudaEventRecord(p->evJoin, p->auxStream));
        if ((rc = runStage(p, 1, 0, out_color, out_mvec, st))) return rc;
        CK(cudaStreamWaitEvent(st, p->evJoin, 0));
    } else {
        if ((rc = runStage(p, 0, 0, out_color, out_mvec, st))) return rc;
        if ((rc = runStage(p, 1, 0, out_color, out_mvec, st))) return rc;
    }
    if ((rc = runStage(p, 7, 0, out_color, out_mvec, st))) return rc;   // mPipelineFrames: K0 + K1 of the next frame start now
    if ((rc = runStage(p, 2, 0, out_color, out_mvec, st))) return rc;
    if (p->P.mEnableSpatialReuse) { for (int r = 0; r < p->P.mSpatialReuseRounds; r++) if ((rc = runStage(p, 3, r, out_color, out_mvec, st))) return rc; }
    else recordEv(p, 4, st);
    if ((rc = runStage(p, 4, 0, out_color, out_mvec, st))) return rc;
    if ((rc = runStage(p, 5, 0, out_color, out_mvec, st))) return rc;
    if ((rc = runStage(p, 6, 0, out_color, out_mvec, st))) return rc;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

// Host-buffer entry points: what a CPU-side caller of the plugin sees.  The frame renders into one of two device images on the
// pass's own stream; its read-back into the caller's buffer runs on a copy stream, so with the _async form the next frame
// renders while the previous one travels (pinned host memory makes the copy truly asynchronous).
int vrestir_execute_host_async(vrestir_pass* p, float* out_color_host, float* out_mvec_host) try {
    if (!p || !out_color_host) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->W <= 0) return setError(VRESTIR_ERR_NOT_READY, "frame not set");
    CK(cudaSetDevice(p->device));
    if (!p->hostStream) {
        CK(cudaStreamCreateWithFlags(&p->hostStream, cudaStreamNonBlocking)); CK(cudaStreamCreateWithFlags(&p->hostCopyStream, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) { CK(cudaEventCreateWithFlags(&p->evHostRendered[i], cudaEventDisableTiming)); CK(cudaEventCreateWithFlags(&p->evHostLanded[i], cudaEventDisableTiming)); }
    }
    const size_t n = N(p);
    if (p->hostColorPixels != n) {
        CK(cudaStreamSynchronize(p->hostStream)); CK(cudaStreamSynchronize(p->hostCopyStream));
        for (int i = 0; i < 2; i++) {
            if (p->d_hostColor[i]) cudaFree(p->d_hostColor[i]);
            if (p->d_hostMvec[i]) cudaFree(p->d_hostMvec[i]);
            CK(cudaMalloc(&p->d_hostColor[i], n * 16)); CK(cudaMalloc(&p->d_hostMvec[i], n * 8));
            CK(cudaMemsetAsync(p->d_hostColor[i], 0, n * 16, p->hostStream)); CK(cudaMemsetAsync(p->d_hostMvec[i], 0, n * 8, p->hostStream));
        }
        p->hostColorPixels = n; p->hostSeq = 0;
    }
    const int b = (int)(p->hostSeq & 1);
    if (p->hostSeq >= 2) CK(cudaStreamWaitEvent(p->hostStream, p->evHostLanded[b], 0));   // the device image is free once its last read-back landed
    int rc = vrestir_execute(p, (float*)p->d_hostColor[b], out_mvec_host ? (float*)p->d_hostMvec[b] : nullptr, p->hostStream);
    if (rc) return rc;
    CK(cudaEventRecord(p->evHostRendered[b], p->hostStream));
    CK(cudaStreamWaitEvent(p->hostCopyStream, p->evHostRendered[b], 0));
    if ((rc = vrestir_wait_output(p, p->hostCopyStream))) return rc;                         // deferred final shading (mPipelineFrames 2)
    const size_t off = (size_t)p->rowBegin * p->W, cnt = (size_t)(p->rowEnd - p->rowBegin) * p->W;
    CK(cudaMemcpyAsync(out_color_host + off * 4, p->d_hostColor[b] + off, cnt * 16, cudaMemcpyDeviceToHost, p->hostCopyStream));
    if (out_mvec_host) CK(cudaMemcpyAsync(out_mvec_host + off * 2, p->d_hostMvec[b] + off, cnt * 8, cudaMemcpyDeviceToHost, p->hostCopyStream));
    CK(cudaEventRecord(p->evHostLanded[b], p->hostCopyStream));
    p->hostSeq++;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_host_wait(vrestir_pass* p) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (p->hostCopyStream) { CK(cudaSetDevice(p->device)); CK(cudaStreamSynchronize(p->hostCopyStream)); }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_execute_host(vrestir_pass* p, float* out_color_host, float* out_mvec_host) try {
    const int rc = vrestir_execute_host_async(p, out_color_host, out_mvec_host);
    return rc ? rc : vrestir_host_wait(p);
} catch (...) { return vr::caughtException(); }

int vrestir_get_timings(vrestir_pass* p, vrestir_timings* out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    for (int i = 0; i <= 6; i++) if (!p->evValid[i]) return setError(VRESTIR_ERR_NOT_READY, "no completed frame");
    CK(cudaEventSynchronize(p->ev[6]));
    float* dst[6] = {&out->features_ms, &out->initial_ms, &out->temporal_ms, &out->spatial_ms, &out->copy_ms, &out->final_ms};
    for (int i = 0; i < 6; i++) CK(cudaEventElapsedTime(dst[i], p->ev[i], p->ev[i + 1]));
    CK(cudaEventElapsedTime(&out->total_ms, p->ev[0], p->ev[6]));
    if (p->framesOverlapped && p->evValid[7]) {   // K0 ran next to K1: initial / total are measured from the frame start on the main stream
        CK(cudaEventElapsedTime(&out->initial_ms, p->ev[7], p->ev[2]));
        CK(cudaEventElapsedTime(&out->total_ms, p->ev[7], p->ev[6]));
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_march_timings(vrestir_pass* p, vrestir_march_timings* out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (!p->evMarchValid) return setError(VRESTIR_ERR_NOT_READY, "no wavefront spatial round has run");
    CK(cudaSetDevice(p->device));
    if (p->marchChunksUsed > 0) {   // generic path: sum over the row chunks of the round
        CK(cudaDeviceSynchronize());
        std::vector<unsigned> cnt(2 * p->marchChunksUsed);
        CK(cudaMemcpy(cnt.data(), p->d_marchCounts, cnt.size() * sizeof(unsigned), cudaMemcpyDeviceToHost));
        out->spatial_cam_ms = out->spatial_light_ms = 0.f; out->spatial_cam_tasks = out->spatial_light_tasks = 0;
        for (int c = 0; c < p->marchChunksUsed; c++) {
            float a = 0.f, b = 0.f;
            CK(cudaEventElapsedTime(&a, p->evMarchChunks[3 * c], p->evMarchChunks[3 * c + 1]));
            CK(cudaEventElapsedTime(&b, p->evMarchChunks[3 * c + 1], p->evMarchChunks[3 * c + 2]));
            out->spatial_cam_ms += a; out->spatial_light_ms += b;
            out->spatial_cam_tasks += cnt[2 * c]; out->spatial_light_tasks += cnt[2 * c + 1];
        }
        return VRESTIR_OK;
    }
    CK(cudaEventSynchronize(p->evMarch[2]));
    CK(cudaEventElapsedTime(&out->spatial_cam_ms, p->evMarch[0], p->evMarch[1]));
    CK(cudaEventElapsedTime(&out->spatial_light_ms, p->evMarch[1], p->evMarch[2]));
    uint32_t c[2] = {0, 0};
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(c, p->wfCounters + 8, 8, cudaMemcpyDeviceToHost));
    out->spatial_cam_tasks = c[0]; out->spatial_light_tasks = c[1];
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_debug_read_bandwidth(int device, size_t bytes, int iters, float* gbs) try {
    if (!gbs || bytes < 4096 || iters < 1) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad argument");
    CK(cudaSetDevice(device));
    void* buf = nullptr; unsigned* sink = nullptr;
    CK(cudaMalloc(&buf, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(buf, 1, bytes));
    int sms = 0; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    CK(launchReadBandwidth(buf, bytes, 2, sms * 8, sink, nullptr));   // warm-up: brings the buffer into L2 when it fits
    CK(cudaEventRecord(e0, nullptr));
    CK(launchReadBandwidth(buf, bytes, iters, sms * 8, sink, nullptr));
    CK(cudaEventRecord(e1, nullptr));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f; CK(cudaEventElapsedTime(&ms, e0, e1));
    *gbs = (float)((double)(bytes / 16 * 16) * iters / (ms * 1e-3) / 1e9);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(buf); cudaFree(sink);
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_set_next_camera(vrestir_pass* p, const vrestir_camera* cam) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (cam) { p->nextCam = *cam; p->haveNextCam = true; } else p->haveNextCam = false;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_wait_output(vrestir_pass* p, void* stream) try {
    if (!p) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null pass");
    if (p->outSeq) { CK(cudaSetDevice(p->device)); CK(cudaStreamWaitEvent((cudaStream_t)stream, p->evOutDone[(p->outSeq - 1) & 1], 0)); }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_pipeline_stats(vrestir_pass* p, vrestir_pipeline_stats* out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    memset(out, 0, sizeof(*out));
    out->adopted = p->pfAdopted; out->discarded = p->pfDiscarded;
    if (p->pfTimed) {
        CK(cudaSetDevice(p->device));
        CK(cudaEventSynchronize(p->evPf1));
        CK(cudaEventElapsedTime(&out->prefetch_ms, p->evPf0, p->evPf1));
    }
    if (p->outTimed) {
        CK(cudaSetDevice(p->device));
        CK(cudaEventSynchronize(p->evOut1));
        CK(cudaEventElapsedTime(&out->deferred_final_ms, p->evOut0, p->evOut1));
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_launch_count(const vrestir_pass* p, uint64_t* out) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *out = p->launches; return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

static int physOf(const vrestir_pass* p, int buffer) {
    switch (buffer) {
        case VRESTIR_BUF_RESERVOIR_0: case VRESTIR_BUF_EXTRA_0: case VRESTIR_BUF_PPARTIAL_0: return p->ia;
        case VRESTIR_BUF_RESERVOIR_1: case VRESTIR_BUF_EXTRA_1: case VRESTIR_BUF_PPARTIAL_1: return p->ib;
        default: return p->it;
    }
}
int vrestir_buffer_bytes(const vrestir_pass* p, int buffer, size_t* bytes) try {
    if (!p || !bytes) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    const size_t n = N(p); const int B = p->P.mMaxBounces;
    switch (buffer) {
        case VRESTIR_BUF_RESERVOIR_0: case VRESTIR_BUF_RESERVOIR_1: case VRESTIR_BUF_RESERVOIR_TEMPORAL: *bytes = n * sizeof(vrestir_reservoir); break;
        case VRESTIR_BUF_EXTRA_0: case VRESTIR_BUF_EXTRA_1: case VRESTIR_BUF_EXTRA_TEMPORAL: *bytes = n * (size_t)std::max(0, B - 1) * 12; break;
        case VRESTIR_BUF_FEATURES: case VRESTIR_BUF_FEATURES_TEMPORAL: *bytes = n * 8; break;
        case VRESTIR_BUF_ENV_IMPORTANCE: *bytes = p->importanceCount * 4; break;
        case VRESTIR_BUF_PPARTIAL_0: case VRESTIR_BUF_PPARTIAL_1: case VRESTIR_BUF_PPARTIAL_TEMPORAL:
            if (!vertexReuseOn(p)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "p_partial exists only with mVertexReuse and mMaxBounces > 1");
            *bytes = n * 4; break;
        default: return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad buffer id");
    }
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_get_buffer(vrestir_pass* p, int buffer, void* dst, size_t bytes) try {
    if (!p || (!dst && bytes)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(p->device));
    size_t need; int rc = vrestir_buffer_bytes(p, buffer, &need); if (rc) return rc;
    if (need != bytes) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "size mismatch");
    if (buffer != VRESTIR_BUF_ENV_IMPORTANCE) { if (p->W <= 0) return setError(VRESTIR_ERR_NOT_READY, "frame not set"); rc = ensureBuffers(p); if (rc) return rc; }
    CK(cudaDeviceSynchronize());
    if (!bytes) return VRESTIR_OK;
    if (buffer >= VRESTIR_BUF_PPARTIAL_0 && buffer <= VRESTIR_BUF_PPARTIAL_TEMPORAL) CK(cudaMemcpy(dst, resView(p, physOf(p, buffer)).p2, bytes, cudaMemcpyDeviceToHost));
    else if (buffer <= VRESTIR_BUF_RESERVOIR_TEMPORAL) {
        vrestir_reservoir* tmp; CK(cudaMalloc(&tmp, bytes));
        cudaError_t e = launchResToAos(resView(p, physOf(p, buffer)), tmp, (int)N(p), 0); p->launches++;
        if (e == cudaSuccess) e = cudaMemcpy(dst, tmp, bytes, cudaMemcpyDeviceToHost);
        cudaFree(tmp); CK(e);
    } else if (buffer <= VRESTIR_BUF_EXTRA_TEMPORAL) CK(cudaMemcpy(dst, p->ext[physOf(p, buffer)], bytes, cudaMemcpyDeviceToHost));
    else if (buffer == VRESTIR_BUF_FEATURES) CK(cudaMemcpy(dst, p->feat[p->featCur], bytes, cudaMemcpyDeviceToHost));
    else if (buffer == VRESTIR_BUF_FEATURES_TEMPORAL) CK(cudaMemcpy(dst, p->feat[p->featSwapPending ? p->featCur : p->featPrev], bytes, cudaMemcpyDeviceToHost));
    else CK(cudaMemcpy(dst, p->d_importance, bytes, cudaMemcpyDeviceToHost));
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_set_buffer(vrestir_pass* p, int buffer, const void* src, size_t bytes) try {
    if (!p || (!src && bytes)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(p->device));
    size_t need; int rc = vrestir_buffer_bytes(p, buffer, &need); if (rc) return rc;
    if (need != bytes) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "size mismatch");
    if (buffer != VRESTIR_BUF_ENV_IMPORTANCE) { if (p->W <= 0) return setError(VRESTIR_ERR_NOT_READY, "frame not set"); rc = ensureBuffers(p); if (rc) return rc; }
    CK(cudaDeviceSynchronize());
    if (!bytes) return VRESTIR_OK;
    if (buffer >= VRESTIR_BUF_PPARTIAL_0 && buffer <= VRESTIR_BUF_PPARTIAL_TEMPORAL) CK(cudaMemcpy(resView(p, physOf(p, buffer)).p2, src, bytes, cudaMemcpyHostToDevice));
    else if (buffer <= VRESTIR_BUF_RESERVOIR_TEMPORAL) {
        vrestir_reservoir* tmp; CK(cudaMalloc(&tmp, bytes));
        cudaError_t e = cudaMemcpy(tmp, src, bytes, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) { e = launchResFromAos(resView(p, physOf(p, buffer)), tmp, (int)N(p), 0); p->launches++; }
        if (e == cudaSuccess) e = cudaDeviceSynchronize();
        cudaFree(tmp); CK(e);
    } else if (buffer <= VRESTIR_BUF_EXTRA_TEMPORAL) CK(cudaMemcpy(p->ext[physOf(p, buffer)], src, bytes, cudaMemcpyHostToDevice));
    else if (buffer == VRESTIR_BUF_FEATURES) CK(cudaMemcpy(p->feat[p->featCur], src, bytes, cudaMemcpyHostToDevice));
    else if (buffer == VRESTIR_BUF_FEATURES_TEMPORAL) {
        // while the swap is pending one buffer plays both roles: the history gets a buffer of its own before it is overwritten
        p->featSwapPending = false;
        CK(cudaMemcpy(p->feat[p->featPrev], src, bytes, cudaMemcpyHostToDevice));
    }
    else CK(cudaMemcpy(p->d_importance, src, bytes, cudaMemcpyHostToDevice));
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_device_buffer(vrestir_pass* p, int buffer, void** base, size_t* plane_stride_bytes, int* planes) try {
    if (!p || !base) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    if (p->W <= 0) return setError(VRESTIR_ERR_NOT_READY, "frame not set");
    int rc = ensureBuffers(p); if (rc) return rc;
    const size_t n = N(p);
    if (buffer >= VRESTIR_BUF_PPARTIAL_0 && buffer <= VRESTIR_BUF_PPARTIAL_TEMPORAL) {
        if (!vertexReuseOn(p)) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "p_partial exists only with mVertexReuse and mMaxBounces > 1");
        *base = resView(p, physOf(p, buffer)).p2; if (plane_stride_bytes) *plane_stride_bytes = 0; if (planes) *planes = 1;
    }
    else if (buffer <= VRESTIR_BUF_RESERVOIR_TEMPORAL) { *base = p->res[physOf(p, buffer)]; if (plane_stride_bytes) *plane_stride_bytes = n * 16; if (planes) *planes = 2; }
    else if (buffer <= VRESTIR_BUF_EXTRA_TEMPORAL) { *base = p->ext[physOf(p, buffer)]; if (plane_stride_bytes) *plane_stride_bytes = 0; if (planes) *planes = 1; }
    else if (buffer == VRESTIR_BUF_FEATURES) { *base = p->feat[p->featCur]; if (plane_stride_bytes) *plane_stride_bytes = 0; if (planes) *planes = 1; }
    else if (buffer == VRESTIR_BUF_FEATURES_TEMPORAL) { *base = p->feat[p->featSwapPending ? p->featCur : p->featPrev]; if (plane_stride_bytes) *plane_stride_bytes = 0; if (planes) *planes = 1; }
    else return setError(VRESTIR_ERR_INVALID_ARGUMENT, "bad buffer id");
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_spatial_input_buffer(const vrestir_pass* p, int round, int* buffer) try {
    if (!p || !buffer) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    *buffer = (round % 2 == 0) ? VRESTIR_BUF_RESERVOIR_0 : VRESTIR_BUF_RESERVOIR_1;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

int vrestir_debug_wavefront_counters(vrestir_pass* p, uint32_t out[16]) try {
    if (!p || !out) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    memset(out, 0, 64);
    if (!p->wfCounters) return VRESTIR_OK;
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, p->wfCounters, 64, cudaMemcpyDeviceToHost));
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }
int vrestir_debug_long_rays(vrestir_pass* p, float* out64x8, uint32_t* count) try {
    if (!p || !out64x8 || !count) return setError(VRESTIR_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(p->device));
    CK(cudaDeviceSynchronize());
    unsigned c = 0;
    CK(readDebugRays(out64x8, &c));
    *count = c;
    return VRESTIR_OK;
} catch (...) { return vr::caughtException(); }

}  // extern "C"
