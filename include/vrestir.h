/*
 * vrestir.h — C ABI of the B200-native VolumetricReSTIR render-pass hot path.
 *
 * Every entry point below replaces one piece of the reference's plugin surface for this path
 * (paths relative to the reference checkout; VR/ = Source/RenderPasses/VolumetricReSTIR/,
 * F/ = Source/Falcor/):
 *
 *   vrestir_create / vrestir_destroy      VR/VolumetricReSTIR.cpp:56-64 (getPasses -> registerClass -> create),
 *                                         VR/VolumetricReSTIR.cpp:66-140 (ctor)
 *   vrestir_set_volume                    F/Scene/Scene.cpp:2754-3326 (addGVDBVolume -> VDBInfo + VolumeDesc),
 *                                         F/Scene/Scene.cpp:2659-2707 (bindParameterBlock)
 *   vrestir_set_camera                    F/Scene/Camera/Camera.cpp:150-189 (CameraData U,V,W, view/proj)
 *   vrestir_set_envmap                    F/Experimental/Scene/Lights/EnvMap.cpp:48-110, EnvMapSampler.cpp:83-116
 *   vrestir_set_analytic_lights           F/Scene/Lights/LightData.slang:52-71
 *   vrestir_set_emissive_triangles        F/Experimental/Scene/Lights/EmissivePowerSampler.cpp:60-80,
 *                                         F/Utils/Sampling/AliasTable.cpp:46-126
 *   vrestir_update                        VR/VolumetricReSTIR.cpp:1280-1347 (updateDict), VR/VolumetricReSTIR.h:278-312
 *   vrestir_get_params / vrestir_set_params   VR/VolumetricReSTIR.cpp:142-147 (getScriptingDictionary), :1349-1430
 *   vrestir_execute                       VR/VolumetricReSTIR.cpp:303-772 (execute: K0..K5 + history copies)
 *   vrestir_get_buffer / vrestir_set_buffer   no reference counterpart (reservoir dump/load for staged parity tests,
 *                                         SURVEY.md section 5 "checkpoint / resume")
 *
 * Plain pointers and sizes only; no torch, CUDA or C++ types.  `stream` arguments are cudaStream_t passed as void*.
 * All functions return VRESTIR_OK (0) on success, a positive warning code, or a negative error code;
 * vrestir_last_error() returns the message of the most recent non-OK return on the calling thread.
 * One host thread per pass handle (same contract as the reference's render thread).
 */
#ifndef VRESTIR_H_
#define VRESTIR_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- constants shared with the shaders: VR/HostDeviceSharedConstants.slang:5-32 ---- */
enum { VRESTIR_MIS_NONE = 0, VRESTIR_MIS_TALBOT = 1 };
enum { VRESTIR_SAMPLER_HAMMERSLEY = 0, VRESTIR_SAMPLER_R2 = 1 };
enum {
    VRESTIR_RATIO_TRACKING = 0,
    VRESTIR_ANALYTIC_TRACKING = 1,
    VRESTIR_RAY_MARCHING = 2,
    VRESTIR_RESIDUAL_RATIO_TRACKING = 3,
    VRESTIR_ANALOG_RESIDUAL_RATIO_TRACKING = 4
};
enum { VRESTIR_REPROJECTION_LINEAR = 0, VRESTIR_REPROJECTION_NONE = 1, VRESTIR_REPROJECTION_NO_BACKGROUND = 2 };
enum { VRESTIR_ENV_SAMPLER_HIERARCHICAL = 0, VRESTIR_ENV_SAMPLER_ALIAS = 1 };

#define VRESTIR_MAX_INITIAL_SAMPLE_COUNT 4
#define VRESTIR_SELF_EMISSION_LIGHT_ID (-3)
#define VRESTIR_NUM_MAX_MIPS 8
#define VRESTIR_MAX_LEVELS 3
#define VRESTIR_MAX_SLOTS 30                     /* MAX_MIPS, F/Scene/GVDB/gvdbNodes.slang:37 */
#define VRESTIR_TEMPERATURE_GRID_ID 16           /* 2*kNumMaxMips     */
#define VRESTIR_VELOCITY_GRID_ID 17              /* 2*kNumMaxMips + 1 */
#define VRESTIR_PREV_DENSITY_GRID_OFFSET 19      /* 2*kNumMaxMips + 3 */
#define VRESTIR_PREV_EXTRA_GRID_OFFSET 11
#define VRESTIR_MAX_BOUNCES 4
#define VRESTIR_BRICK_RES 8
#define VRESTIR_BRICK_APRON 1
#define VRESTIR_BRICK_STRIDE 10                  /* 8 + 2*apron */
#define VRESTIR_BRICK_VOXELS 1000

/* ---- status codes ---- */
enum {
    VRESTIR_OK = 0,
    VRESTIR_WARN_UNKNOWN_KEY = 1,        /* mirrors logWarning on unknown dict key, VR/VolumetricReSTIR.h:309 */
    VRESTIR_ERR_INVALID_ARGUMENT = -1,
    VRESTIR_ERR_UNSUPPORTED = -2,        /* option outside the hot-path scope (surface scene, ...) */
    VRESTIR_ERR_CUDA = -3,
    VRESTIR_ERR_NOT_READY = -4,          /* execute() before set_volume()/set_camera() */
    VRESTIR_ERR_IO = -5
};

/* ---- VolumetricReSTIRParams: VR/VolumetricReSTIR.h:130-207 (same field names, same defaults) ---- */
typedef struct vrestir_params {
    int32_t mMaxBounces;                          /* 1 */
    int32_t mEnableTemporalReuse;                 /* true */
    int32_t mEnableSpatialReuse;                  /* true */
    int32_t mVertexReuse;                         /* false */
    int32_t mVertexReuseStartBounce;              /* 1 */
    int32_t mUseReference;                        /* false */
    int32_t mUseEnvironmentLights;                /* true */
    int32_t mUseAnalyticLights;                   /* false */
    int32_t mUseEmissiveLights;                   /* false */
    int32_t mBaselineSamplePerPixel;              /* 1 */
    int32_t mVisualizeTotalTransmittance;         /* false */
    int32_t mUseSurfaceScene;                     /* false */
    int32_t mUsePrevVolumeForReproj;              /* true */
    /* initial sampling */
    int32_t mInitialBaseMipLevel;                 /* 1 */
    int32_t mInitialM;                            /* 4 */
    int32_t mInitialLightSamples;                 /* 1 */
    int32_t mInitialLightingMipLevel;             /* 2 */
    int32_t mInitialVisibilityUseLinearSampler;   /* false */
    int32_t mInitialLightingUseLinearSampler;     /* true */
    uint32_t mInitialLightingTrackingMethod;      /* kRayMarching */
    float mInitialVisibilityTStepScale;           /* 1 */
    float mInitialLightingTStepScale;             /* 2 */
    int32_t mInitialUseRussianRoulette;           /* true */
    int32_t mInitialUseCoarserGridForIndirectBounce; /* true */
    /* temporal reuse */
    float mTemporalReuseMThreshold;               /* 4 */
    uint32_t mTemporalReprojectionMode;           /* kReprojectionLinear */
    uint32_t mTemporalMISMethod;                  /* kMISTalbot */
    int32_t mTemporalReprojectionMipLevel;        /* 1 */
    /* spatial reuse */
    int32_t mSpatialReuseRounds;                  /* 1 */
    int32_t mSpatialVisibilityMipLevel;           /* 1 */
    int32_t mSpatialLightingMipLevel;             /* 1 */
    int32_t mSpatialVisibilityUseLinearSampler;   /* true */
    int32_t mSpatialLightingUseLinearSampler;     /* true */
    float mSpatialVisibilityTStepScale;           /* 1 */
    float mSpatialLightingTStepScale;             /* 1 */
    uint32_t mSpatialVisibilityTrackingMethod;    /* kRayMarching */
    uint32_t mSpatialLightingTrackingMethod;      /* kRayMarching */
    uint32_t mRandomSamplerType;                  /* kR2 */
    float mSampleRadius;                          /* 10 */
    int32_t mSpatialSampleCount;                  /* 4 */
    int32_t mEnableVisibilitySimilarityRejection; /* false (accepted, unused) */
    uint32_t mSpatialMISMethod;                   /* kMISTalbot */
    /* final shading */
    int32_t mFinalLightSamples;                   /* 1 */
    int32_t mFinalVisibilitySamples;              /* 1 */
    uint32_t mFinalVisibilityTrackingMethod;      /* kAnalyticTracking */
    uint32_t mFinalLightTrackingMethod;           /* kAnalyticTracking */
    uint32_t mFinalRandomSamplerType;             /* kR2 (unused) */
    float mFinalTStepScale;                       /* 0.2 */
} vrestir_params;

/* ---- GVDB-style sparse grid, one per slot: F/Scene/GVDB/gvdbNodes.slang:38-95, F/Scene/Scene.cpp:2898-3240 ----
 *
 * Node (32 B, one L2 sector).  Replaces the reference's repacked VDBNode {int3 packedPosValue; uint childList;
 * float4 densityBounds} (F/Scene/Scene.cpp:2908-3023): pos is the node's index-space minimum corner (unpackPos);
 * `link` is the child-list row for levels >= 1 (mChildList) and the brick-pool index for level 0 (unpackValue:
 * the brick's atlas location); bounds = (min, max, avg, 0) density over the 10^3 apron-inclusive block with avg
 * divided by 512 (F/Scene/Scene.cpp:2981-3012).
 */
typedef struct vrestir_node {
    int32_t pos[3];
    uint32_t link;
    float bounds[4];
} vrestir_node;

enum { VRESTIR_ATLAS_F32 = 0, VRESTIR_ATLAS_UNORM8 = 1 };

typedef struct vrestir_grid_slot {
    int32_t valid;                               /* 0 = slot not bound */
    int32_t top_lev;                             /* level whose node count is 1 (1 or 2) */
    int32_t dim[VRESTIR_MAX_LEVELS];             /* log2 of children per axis: 3,4,5 */
    int32_t res[VRESTIR_MAX_LEVELS];             /* children per axis: 8,16,32 */
    float vdel[VRESTIR_MAX_LEVELS];              /* voxels covered by one child: 1,8,128 */
    int32_t noderange[VRESTIR_MAX_LEVELS];       /* voxels covered by one node: 8,128,4096 */
    uint32_t node_count[VRESTIR_MAX_LEVELS];
    const vrestir_node* nodes[VRESTIR_MAX_LEVELS];
    const uint32_t* childlist[VRESTIR_MAX_LEVELS]; /* [lev>=1]: node.link*res^3 + b -> child node id, 0xFFFFFFFF = none */
    uint64_t childlist_count[VRESTIR_MAX_LEVELS];
    float bmin[3];                               /* inclusive AABB min in voxels (index space of this slot) */
    float bmax[3];                               /* exclusive AABB max */
    float xform[16];                             /* index -> model, row-vector convention p' = p * M, row-major */
    float invxform[16];
    float world_to_medium[16];                   /* volumeExternalWorldToModel * invxform (VR/VolumeBase.slang:103-116) */
    float medium_to_world[16];                   /* xform * volumeExternalModelToWorld    (VR/VolumeBase.slang:119-130) */
    float max_value;                             /* maxValue[slot] */
    float compress_scale;                        /* densityCompressScaleFactor[slot] (1 for fp32 atlases) */
    int32_t atlas_format;                        /* VRESTIR_ATLAS_* */
    int32_t atlas_channels;                      /* 1 (density, temperature) or 3 (velocity) */
    uint32_t brick_count;
    const void* atlas;                           /* brick pool: [brick][channel][10*10*10], x fastest, apron included */
} vrestir_grid_slot;

/* VolumeDesc: F/Scene/SceneTypes.slang:85-110, filled like F/Scene/Scene.cpp:3246-3298 */
typedef struct vrestir_volume_desc {
    float sigma_t;
    float sigma_s[3];
    float sigma_a[3];
    float PhaseFunctionConstantG;
    float densityScaleFactor;
    float densityScaleFactorByScaling;
    float tStep;
    int32_t hasEmission;
    int32_t hasVelocity;
    int32_t hasAnimation;
    int32_t lastFrameHasEmission;
    float LeScale;
    float temperatureCutOff;
    float temperatureScale;
    float velocityScale;
    int32_t numMips;
    int32_t usePrevGridForReproj;
    float volumeWorldScaling;                    /* gScene.volumeWorldScaling */
    float superVoxelWorldSpaceDiagonalLength;    /* F/Scene/Scene.cpp:3077-3080 */
    float externalModelToWorld[16];              /* gScene.volumeExternalModelToWorldMatrix */
    float externalWorldToModel[16];
} vrestir_volume_desc;

typedef struct vrestir_grid_desc {
    vrestir_volume_desc volume;
    vrestir_grid_slot slots[VRESTIR_MAX_SLOTS];
    const float* blackbody_lut;                  /* 128 x RGBA32F (gBlackBodyRadiationTex) or NULL */
} vrestir_grid_desc;

/* CameraData subset used by the path: F/Scene/Camera/CameraData.slang:35-65 */
typedef struct vrestir_camera {
    float posW[3];
    float cameraU[3];
    float cameraV[3];
    float cameraW[3];
    float viewMat[16];                           /* row-vector convention */
    float projMat[16];
    float nearZ, farZ;
} vrestir_camera;

typedef struct vrestir_envmap_desc {
    const float* texels;                         /* lat-long map, RGBA32F, width*height*4 */
    int32_t width, height;
    float intensity;
    float tint[3];
    float transform[9];                          /* local -> world rotation, row-vector convention */
    float invTransform[9];
    float prevTransform[9];
    float prevInvTransform[9];
} vrestir_envmap_desc;

enum { VRESTIR_LIGHT_POINT = 0, VRESTIR_LIGHT_DIRECTIONAL = 1 };
typedef struct vrestir_light {                   /* LightData subset, F/Scene/Lights/LightData.slang:52-71 */
    uint32_t type;
    float posW[3];
    float dirW[3];
    float intensity[3];
} vrestir_light;

typedef struct vrestir_emissive_triangle {       /* EmissiveTriangle, F/.../LightCollectionShared.slang:41-71 */
    float posW[3][3];
    float normal[3];
    float area;
    float Le[3];                                 /* emitted radiance (replaces the material lookup) */
} vrestir_emissive_triangle;

/* ---- buffers addressable through get/set_buffer ---- */
enum {
    VRESTIR_BUF_RESERVOIR_0 = 0,                 /* mPerPixelReservoirBuffer[0], AoS `vrestir_reservoir` view */
    VRESTIR_BUF_RESERVOIR_1 = 1,
    VRESTIR_BUF_RESERVOIR_TEMPORAL = 2,          /* mTemporalReservoirBuffer */
    VRESTIR_BUF_EXTRA_0 = 3,                     /* mPerPixelExtraBounceReservoirBuffer[0], float3 x (B-1) per pixel */
    VRESTIR_BUF_EXTRA_1 = 4,
    VRESTIR_BUF_EXTRA_TEMPORAL = 5,
    VRESTIR_BUF_FEATURES = 6,                    /* mReservoirFeatureBuffer: {int noReflectiveSurface; float transmittance} */
    VRESTIR_BUF_FEATURES_TEMPORAL = 7,
    VRESTIR_BUF_ENV_IMPORTANCE = 8,              /* importance map, all mips, finest first (512^2 + 256^2 + ... + 1) */
    VRESTIR_BUF_PPARTIAL_0 = 9,                  /* Reservoir::p_partial of buffer 0 / 1 / temporal (VERTEX_REUSE, HostDeviceSharedDefinitions.h:29-31), */
    VRESTIR_BUF_PPARTIAL_1 = 10,                 /* one float per pixel; present when mVertexReuse && mMaxBounces > 1 */
    VRESTIR_BUF_PPARTIAL_TEMPORAL = 11,
    VRESTIR_BUF_COUNT = 12
};

/* Host-visible reservoir record (VR/HostDeviceSharedDefinitions.h:16-45 + extraBounceStartId). Device storage is SoA. */
typedef struct vrestir_reservoir {
    float runningSum;
    float M;
    float depth;
    float p_y;
    float lightUV[2];
    int32_t lightID;
    int32_t sampledPixel;
} vrestir_reservoir;

typedef struct vrestir_pass vrestir_pass;

/* per-stage device times of the last vrestir_execute (ms), Profiler events VR/VolumetricReSTIR.cpp:501..761 */
typedef struct vrestir_timings {
    float features_ms, initial_ms, temporal_ms, spatial_ms, copy_ms, final_ms, total_ms;
} vrestir_timings;

const char* vrestir_last_error(void);
const char* vrestir_version(void);

void vrestir_default_params(vrestir_params* out);

int vrestir_create(const vrestir_params* params, int device, vrestir_pass** out);
int vrestir_destroy(vrestir_pass* pass);

int vrestir_set_volume(vrestir_pass* pass, const vrestir_grid_desc* grid);
/* Animated volumes: rebinds the previous call's grids to the prev-frame slots (19..28) and uploads `grid` as current
 * (F/Scene/Scene.cpp:825-863). */
int vrestir_advance_volume(vrestir_pass* pass, const vrestir_grid_desc* grid);
/* Resident animation frames.  The reference keeps every frame of an animated sequence on the GPU and switches the bound grids
 * and the volume description once per frame (F/Scene/Scene.cpp:825-863: mVolumeDescArray[mVDBAnimationFrameId]).
 * vrestir_volume_frame_add uploads the current-frame slots of `grid` once and returns the frame's index;
 * vrestir_advance_volume_resident(index) has the semantics of vrestir_advance_volume with that frame as the new volume but only
 * rebinds pointers (no copy, no allocation, no device-wide wait).  vrestir_volume_frames_clear releases the frames (a volume
 * must be set again if the pass was still bound to one). */
int vrestir_volume_frame_add(vrestir_pass* pass, const vrestir_grid_desc* grid, int* out_index);
int vrestir_advance_volume_resident(vrestir_pass* pass, int index);
int vrestir_volume_frames_clear(vrestir_pass* pass);
int vrestir_set_camera(vrestir_pass* pass, const vrestir_camera* camera);
int vrestir_set_envmap(vrestir_pass* pass, const vrestir_envmap_desc* env);
int vrestir_set_analytic_lights(vrestir_pass* pass, const vrestir_light* lights, int count);
int vrestir_set_emissive_triangles(vrestir_pass* pass, const vrestir_emissive_triangle* tris, int count,
                                   float emissiveIntensityMultiplier);

/* Alias tables the pass built (host copies): emissive items = {threshold bits, indexA, indexB, pad} x count
 * (F/Utils/Sampling/AliasTable.cpp:110-124) + original weights; env = per-texel keep-threshold + redirect texel. */
int vrestir_get_emissive_alias(const vrestir_pass* pass, uint32_t* items4, float* weights, float* weight_sum);
int vrestir_get_env_alias(const vrestir_pass* pass, float* thresholds, uint32_t* redirect, int* count);
/* The host-side table builders on their own (no device needed): the emissive table exactly as AliasTable::AliasTable builds it
 * from per-triangle flux with std::mt19937(123) (F/Utils/Sampling/AliasTable.cpp:46-126, EmissivePowerSampler.cpp:64-79), and
 * the env-map table over `count` texel weights (what vrestir_set_envmap builds over the finest importance mip). */
int vrestir_build_alias_table(const float* weights, int count, uint32_t* items4, float* weight_sum);
int vrestir_build_env_alias(const float* texel_weights, int count, float* thresholds, uint32_t* redirect);

/* Image size and the row band this pass instance owns ([row_begin,row_end) of `height`; whole frame = 0,height). */
int vrestir_set_frame(vrestir_pass* pass, int width, int height, int row_begin, int row_end);

/* updateDict: `key` is a VolumetricReSTIRParams field ("mInitialM", ...), a top-level key ("mOutputMotionVec",
 * "mFreezeFrame", "volumeDensityScaleExtraControl", "volumeAlbedoExtraControl", "volumeAnisotropyExtraControl",
 * "mEnvSamplerType") or a verb ("randomizeFrameSeed").  Resets the frame counter and temporal history like
 * VR/VolumetricReSTIR.cpp:1339.  Unknown key -> VRESTIR_WARN_UNKNOWN_KEY. */
int vrestir_update(vrestir_pass* pass, const char* key, double value);
int vrestir_set_params(vrestir_pass* pass, const vrestir_params* params);
int vrestir_get_params(const vrestir_pass* pass, vrestir_params* out);
int vrestir_set_frame_count(vrestir_pass* pass, int frame_count, int temporal_sample_accumulated);
/* Previous-frame camera (mPrevViewMat/mPrevProjMat/mPrevCameraU..PosW, VR/VolumetricReSTIR.cpp:767-769); normally saved by
 * the end-of-frame stage, settable for staged parity tests of temporal reuse. */
int vrestir_set_prev_camera(vrestir_pass* pass, const vrestir_camera* camera);
int vrestir_get_frame_count(const vrestir_pass* pass, int* frame_count);

/* One frame. out_color: device pointer, width*height float4 (accumulated_color, RGBA32F); out_mvec: device pointer,
 * width*height float2 or NULL (mvec, RG32F).  Only rows [row_begin,row_end) are written.  Asynchronous on `stream`. */
int vrestir_execute(vrestir_pass* pass, float* out_color, float* out_mvec, void* stream);
/* Same frame through host buffers (pinned or pageable): runs execute and copies the band back, synchronous. */
int vrestir_execute_host(vrestir_pass* pass, float* out_color_host, float* out_mvec_host);
/* The same without blocking: returns once the frame and its read-back are enqueued (the read-back of frame f overlaps the
 * rendering of frame f+1); vrestir_host_wait blocks until every enqueued frame has landed in its host buffer.  Use pinned
 * host memory, and do not touch a buffer between the call that fills it and the wait. */
int vrestir_execute_host_async(vrestir_pass* pass, float* out_color_host, float* out_mvec_host);
int vrestir_host_wait(vrestir_pass* pass);

/* Individual stages (for staged parity tests and multi-GPU drivers that interleave halo exchanges).
 * stage: 0 features, 1 initial, 2 temporal, 3 spatial round `arg`, 4 copy-to-history, 5 final shading,
 * 6 end-of-frame bookkeeping (saves prev camera, frameCount++),
 * 7 prefetch (frame pipelining, see below): call between stage 1 and stage 2; no-op unless "mPipelineFrames" is on. */
int vrestir_execute_stage(vrestir_pass* pass, int stage, int arg, float* out_color, float* out_mvec, void* stream);

/* Frame pipelining (updateDict key "mPipelineFrames", off by default).  K0 (VR/GenerateFeatures.cs.slang) and K1
 * (VR/TraceRays.cs.slang) read no history, so with the option on vrestir_execute(f) also starts K0 + K1 of frame f+1 on an
 * internal stream, next to K2..K5 of frame f; vrestir_execute(f+1) adopts them when the camera, frame counter, row band,
 * options and scene of frame f+1 are what the prefetch assumed, and otherwise discards them and runs K0/K1 itself (same
 * results either way, bit for bit).  The prefetch assumes the camera stays where it is unless the application announces the
 * next frame's camera here before vrestir_execute(f) (one announcement covers one frame; NULL clears it). */
int vrestir_set_next_camera(vrestir_pass* pass, const vrestir_camera* camera);
typedef struct vrestir_pipeline_stats {
    uint64_t adopted, discarded;   /* prefetched frames used / thrown away since create */
    float prefetch_ms;             /* device time of the last prefetch chain (K0 + K1 of the next frame) on its stream */
    float deferred_final_ms;       /* device time of the last deferred K5 on its stream (level 2) */
} vrestir_pipeline_stats;
/* "mPipelineFrames" = 2 additionally defers K5 (VR/FinalShading.cs.slang): it only reads the frame's final reservoirs, so it
 * runs on a third internal stream next to K2/K3 of the NEXT frame.  out_color / out_mvec of vrestir_execute are then complete
 * once the work enqueued here has run: call it with the stream (0 = default stream) that consumes the image.  A no-op at
 * levels 0 and 1, where the image is complete in `stream` order of vrestir_execute itself.  vrestir_execute_host waits by itself. */
int vrestir_wait_output(vrestir_pass* pass, void* stream);
int vrestir_get_pipeline_stats(vrestir_pass* pass, vrestir_pipeline_stats* out);

int vrestir_get_timings(vrestir_pass* pass, vrestir_timings* out);
/* The two march launches of the last spatial-reuse round (the dominant kernels of a frame): CUDA-event times on the launching
 * stream and the number of tasks each stream held.  Feeds bench.py's roofline object. */
typedef struct vrestir_march_timings {
    float spatial_cam_ms, spatial_light_ms;
    uint32_t spatial_cam_tasks, spatial_light_tasks;
} vrestir_march_timings;
int vrestir_get_march_timings(vrestir_pass* pass, vrestir_march_timings* out);
/* Read-bandwidth microbenchmark: `bytes` are read `iters` times by every SM with 16-byte loads (CUDA events); a buffer that
 * fits L2 (e.g. 32 MiB) gives the L2 -> SM peak the volume-fetch kernels are compared with, a large one the HBM read peak. */
int vrestir_debug_read_bandwidth(int device, size_t bytes, int iters, float* gb_per_s);
/* number of kernels this library launched since create (claim for bench.py's gpu_launches) */
int vrestir_get_launch_count(const vrestir_pass* pass, uint64_t* out);

/* ---- the passes behind VolumetricReSTIR.accumulated_color in the reference's scripts ---------------
 * AccumulatePass (Source/RenderPasses/AccumulatePass/AccumulatePass.cpp:128-205, Accumulate.cs.slang:57-122): running mean
 * of the frames, three precision modes (AccumulatePass.h:66-71; default Double), pass-through when "enableAccumulation" is
 * off, "subFrameCount" stops after N frames while "autoReset" is on.  Scene / camera / refresh-flag changes reset the
 * reference's counter through the render graph; here the owner calls vrestir_accum_reset. */
enum { VRESTIR_ACCUM_DOUBLE = 0, VRESTIR_ACCUM_SINGLE = 1, VRESTIR_ACCUM_SINGLE_COMPENSATED = 2 };
typedef struct vrestir_accumulator vrestir_accumulator;
int vrestir_accum_create(int device, int width, int height, vrestir_accumulator** out);
int vrestir_accum_destroy(vrestir_accumulator* acc);
/* keys: "enableAccumulation", "autoReset", "precisionMode" (VRESTIR_ACCUM_*), "subFrameCount"; unknown -> VRESTIR_WARN_UNKNOWN_KEY */
int vrestir_accum_update(vrestir_accumulator* acc, const char* key, double value);
int vrestir_accum_reset(vrestir_accumulator* acc);
int vrestir_accum_resize(vrestir_accumulator* acc, int width, int height);   /* a new resolution restarts the accumulation */
int vrestir_accum_frame_count(const vrestir_accumulator* acc, int* out);
/* input / output: device pointers, width*height float4; rows [row_begin,row_end) are accumulated (one call per frame and
 * band owner).  Asynchronous on `stream`. */
int vrestir_accum_execute(vrestir_accumulator* acc, const float* input, float* output, int row_begin, int row_end, void* stream);
/* ErrorMeasurePass (Source/RenderPasses/ErrorMeasurePass/ErrorMeasurer.cs.slang:41-61, ErrorMeasurePass.cpp:217-259):
 * per-pixel |source - reference| (squared / rgb-averaged on request, background pixels = world_position.w == 0 skipped when
 * ignore_background and world_position is bound), summed and divided by the pixel count.  error_rgb_avg = {r, g, b,
 * (r+g+b)/3}.  difference_out (device, float4 per pixel) may be NULL.  Synchronous (returns the numbers). */
int vrestir_error_measure(int device, const float* source, const float* reference, const float* world_position, int width, int height,
                          int ignore_background, int compute_squared_difference, int compute_average, float* difference_out,
                          float error_rgb_avg[4], void* stream);

/* ToneMapper (Source/RenderPasses/ToneMapper/ToneMapper.cpp, ToneMapping.ps.slang, Luminance.ps.slang): exposure (manual from
 * f-number / shutter / film speed, or auto from the average log-luminance of the frame), exposure compensation, CAT02 white
 * balance (F/Utils/Color/ColorUtils.h:200-216), one of six operators, optional clamp.  Defaults = the reference's
 * (ToneMapper.h:111-125, operator Aces).  vrestir_tonemap_params_from_settings is host-only (no device needed). */
enum { VRESTIR_TONEMAP_LINEAR = 0, VRESTIR_TONEMAP_REINHARD = 1, VRESTIR_TONEMAP_REINHARD_MODIFIED = 2, VRESTIR_TONEMAP_HEJI_HABLE_ALU = 3,
       VRESTIR_TONEMAP_HABLE_UC2 = 4, VRESTIR_TONEMAP_ACES = 5 };
typedef struct vrestir_tonemap_settings {
    float exposureCompensation; int32_t autoExposure; float filmSpeed; int32_t whiteBalance; float whitePoint;
    uint32_t op; int32_t clamp; float whiteMaxLuminance, whiteScale, fNumber, shutter;
} vrestir_tonemap_settings;
typedef struct vrestir_tonemap_params {   /* what the shader's constant buffer holds (ToneMapperParams.slang:52-59) + the pass's defines */
    uint32_t op; int32_t autoExposure, clamp; float whiteScale, whiteMaxLuminance; float colorTransform[9];
} vrestir_tonemap_params;
void vrestir_tonemap_default_settings(vrestir_tonemap_settings* out);
int vrestir_tonemap_params_from_settings(const vrestir_tonemap_settings* settings, vrestir_tonemap_params* out);
/* src / dst: device pointers, width*height float4 (may alias).  avg_log_luminance_out (host, nullable): the auto-exposure
 * average (log2) — requesting it makes the call synchronous.  Asynchronous on `stream` otherwise. */
int vrestir_tonemap_execute(int device, const vrestir_tonemap_params* params, const float* src, float* dst, int width, int height,
                            float* avg_log_luminance_out, void* stream);

/* ---- mip / conservative-mip chain on the GPU (the converter's job in the reference) ------------------
 * Builds, from a dense device-resident density grid ([z][y][x] floats), every level of the normal chain and of the
 * conservative chain by the rule of gvdb-voxel-src/source/gvdb_library/src/gvdb_volume_gvdb.cpp:2703-2885 (conservative mip 0
 * :2753-2801, down-sampling :2803-2862) and stores them the way the brick pool does (F/Scene/Scene.cpp:3139-3174): mip 0 of
 * the normal chain as fp32 after the 1e-9 flush, every other level as UNORM8 codes of value / max_value (conservative codes
 * never round a positive value to 0).  Bit-identical to the host builder behind vrestir_scene_create_from_dense. */
typedef struct vrestir_mip_chain vrestir_mip_chain;
typedef struct vrestir_mip_level {
    const void* data;      /* device pointer: dim[0]*dim[1]*dim[2] floats (VRESTIR_ATLAS_F32) or bytes (VRESTIR_ATLAS_UNORM8), x fastest */
    size_t bytes;
    int32_t dim[3];
    int32_t format;        /* VRESTIR_ATLAS_F32 / VRESTIR_ATLAS_UNORM8 */
    float max_value;       /* max |v| of the level = the UNORM8 scale (densityCompressScaleFactor) */
} vrestir_mip_level;
int vrestir_mips_build_device(int device, const float* dense_mip0, const int32_t dim[3], int num_mips, vrestir_mip_chain** out, void* stream);
int vrestir_mips_count(const vrestir_mip_chain* chain, int* out);
int vrestir_mips_level(const vrestir_mip_chain* chain, int mip, int conservative, vrestir_mip_level* out);
int vrestir_mips_destroy(vrestir_mip_chain* chain);
/* Binds the density slots (every mip, normal and conservative) of `pass` to a GPU-built chain without a round trip of the
 * voxels through the host: brick pools, quad repacks and brick bounds are produced on the device, only the brick-activity
 * maps (1 byte per brick) visit the host, where the tree over them is built (like the host builder, bit-identical slots).
 * `tmpl` = grid description of a host-built volume of the SAME dimensions (e.g. frame 0 of an animated sequence): it
 * supplies the transforms, the volume description and the temperature / velocity grids.  advance != 0 has the semantics of
 * vrestir_advance_volume (the current grids become the previous frame's, F/Scene/Scene.cpp:825-863). */
int vrestir_set_volume_from_chain(vrestir_pass* pass, const vrestir_mip_chain* chain, const vrestir_grid_desc* tmpl, int advance);

/* Diagnostics: world-space rays whose hierarchical DDA ran >= 1024 outer iterations since the last call
 * (8 floats each: origin, dir, mip (+100 when vertex-centred), iterations; first 64) and their total count. */
int vrestir_debug_long_rays(vrestir_pass* pass, float* out64x8, uint32_t* count);
/* Diagnostics of the wavefront path: out[0..7] = task-stream counters {count, cursor} x 4 of the last stage run,
 * out[8..15] reserved. */
int vrestir_debug_wavefront_counters(vrestir_pass* pass, uint32_t out[16]);

/* Buffer access.  Host copies use the AoS views documented at the enum; `bytes` must match vrestir_buffer_bytes. */
int vrestir_buffer_bytes(const vrestir_pass* pass, int buffer, size_t* bytes);
int vrestir_get_buffer(vrestir_pass* pass, int buffer, void* host_dst, size_t bytes);
int vrestir_set_buffer(vrestir_pass* pass, int buffer, const void* host_src, size_t bytes);
/* Raw device pointer + layout of the SoA reservoir planes, for halo exchange by the multi-GPU driver:
 * plane p of buffer b lives at base + p*plane_stride_bytes, element (y*width+x)*16 bytes. */
int vrestir_device_buffer(vrestir_pass* pass, int buffer, void** base, size_t* plane_stride_bytes, int* planes);
/* which ping-pong buffer holds the input of spatial round `round` / the final reservoirs of the frame */
int vrestir_spatial_input_buffer(const vrestir_pass* pass, int round, int* buffer);

/* ---- host-side scene helpers (synthetic grids; F/Scene/Scene.cpp + GV/src/gvdb_volume_gvdb.cpp data contract) ---- */
typedef struct vrestir_scene vrestir_scene;

typedef struct vrestir_scene_params {
    int32_t kind;                 /* 0 sphere-fBm (config 1), 1 bunny-cloud blob, 2 plume frame, 3 dense cloud, 4 thin shells */
    int32_t dim[3];               /* mip-0 voxel dimensions */
    int32_t num_mips;             /* 1..8 */
    uint32_t seed;
    float frame_time;             /* plume animation time (kind 2) */
    float sigma_a[3], sigma_s[3], g;
    float density_scale;          /* addGVDBVolume densityScale */
    float voxel_size;             /* model units per mip-0 voxel */
    float world_translation[3];
    float world_scaling;
    int32_t with_temperature;     /* build slot 16 */
    int32_t with_velocity;        /* build slot 17 */
    float LeScale, temperatureCutOff, temperatureScale;
} vrestir_scene_params;

int vrestir_scene_create(const vrestir_scene_params* p, vrestir_scene** out);
/* Build from a caller-supplied dense density array (dim x*y*z floats, x fastest); temperature/velocity may be NULL. */
int vrestir_scene_create_from_dense(const vrestir_scene_params* p, const float* density, const float* temperature,
                                    const float* velocity_xyz, vrestir_scene** out);
/* Description without voxels (dimensions, formats, transforms, VolumeDesc of every density level): the template of
 * vrestir_set_volume_from_chain for grids that only ever exist on the device. */
int vrestir_scene_create_template(const vrestir_scene_params* p, vrestir_scene** out);
/* The procedural density field of `p` evaluated on the device into dense_out (dim x*y*z floats, x fastest; device memory):
 * identical voxels to vrestir_scene_create's host generator for every kind but the plume. */
int vrestir_make_procedural_device(int device, const vrestir_scene_params* p, float* dense_out, void* stream);
/* Host copy of the volume bound to `pass` (tree with the device-computed brick bounds, child lists, brick pools) as a scene:
 * lets a CPU checker evaluate a grid that was built on the device.  Destroy with vrestir_scene_destroy. */
int vrestir_download_volume(vrestir_pass* pass, vrestir_scene** out);
int vrestir_scene_destroy(vrestir_scene* s);
const vrestir_grid_desc* vrestir_scene_grid(const vrestir_scene* s);
/* dense copy of one mip (debug / tests): conservative = 0/1 */
int vrestir_scene_dense_mip(const vrestir_scene* s, int mip, int conservative, float* out, int32_t out_dim[3]);
int vrestir_scene_stats(const vrestir_scene* s, int slot, uint32_t* bricks, uint64_t* atlas_bytes);

/* Camera helper: Falcor's Camera::calculateCameraParameters (F/Scene/Camera/Camera.cpp:150-189) */
int vrestir_camera_look_at(const float pos[3], const float target[3], const float up[3], float fovY_radians,
                           float aspect, float nearZ, float farZ, vrestir_camera* out);
/* Procedural HDR sky (sun + gradient) lat-long map, RGBA32F */
int vrestir_make_sky_envmap(int width, int height, uint32_t seed, float* out_texels);
/* Procedural emissive-triangle shell (config 4) */
int vrestir_make_emissive_shell(int count, uint32_t seed, const float center[3], float radius,
                                vrestir_emissive_triangle* out);
/* Planck-law RGB LUT, 128 entries RGBA32F (stand-in for F/Data/LUT/BlackBodyRadiationRGB_50K-6400K.txt) */
int vrestir_make_blackbody_lut(float* out_128x4);

/* GVDB .vbx reader (GV/src/gvdb_volume_gvdb.cpp:532-739): loads <dir>/<name>_mip<k>[c].vbx etc. into a scene */
int vrestir_scene_load_vbx(const char* dir_and_prefix, int num_mips, const vrestir_scene_params* p, vrestir_scene** out);
/* GVDB .vbx writer (GV/src/gvdb_volume_gvdb.cpp:1682-1831, version 1.12): one file per grid of the scene, same naming */
int vrestir_scene_save_vbx(const vrestir_scene* scene, const char* dir_and_prefix);

#ifdef __cplusplus
}
#endif
#endif /* VRESTIR_H_ */
