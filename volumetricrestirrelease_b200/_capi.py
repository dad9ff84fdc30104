"""ctypes view of include/vrestir.h and loader of the in-tree CUDA library.

There is no Python or CPU fallback: if ``libvrestir.so`` is missing the import fails loudly, and every compute entry
point fails with VRESTIR_ERR_CUDA when no B200 is visible.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VRESTIR_LIB") or os.path.join(_HERE, "libvrestir.so")   # VRESTIR_LIB: developer override (tuning variants)

MAX_SLOTS = 30
NUM_MAX_MIPS = 8
TEMPERATURE_GRID_ID = 16
VELOCITY_GRID_ID = 17

OK = 0
WARN_UNKNOWN_KEY = 1
ERR_INVALID_ARGUMENT = -1
ERR_UNSUPPORTED = -2
ERR_CUDA = -3
ERR_NOT_READY = -4

# VR/HostDeviceSharedConstants.slang:5-32
kMISNone, kMISTalbot = 0, 1
kHammersley, kR2 = 0, 1
kRatioTracking, kAnalyticTracking, kRayMarching, kResidualRatioTracking, kAnalogResidualRatioTracking = 0, 1, 2, 3, 4
kReprojectionLinear, kReprojectionNone, kReprojectionNoBackground = 0, 1, 2

(BUF_RESERVOIR_0, BUF_RESERVOIR_1, BUF_RESERVOIR_TEMPORAL, BUF_EXTRA_0, BUF_EXTRA_1, BUF_EXTRA_TEMPORAL, BUF_FEATURES,
 BUF_FEATURES_TEMPORAL, BUF_ENV_IMPORTANCE, BUF_PPARTIAL_0, BUF_PPARTIAL_1, BUF_PPARTIAL_TEMPORAL) = range(12)

_I, _U, _F = C.c_int32, C.c_uint32, C.c_float

PARAM_FIELDS = [
    ("mMaxBounces", _I, 1), ("mEnableTemporalReuse", _I, 1), ("mEnableSpatialReuse", _I, 1), ("mVertexReuse", _I, 0),
    ("mVertexReuseStartBounce", _I, 1), ("mUseReference", _I, 0), ("mUseEnvironmentLights", _I, 1),
    ("mUseAnalyticLights", _I, 0), ("mUseEmissiveLights", _I, 0), ("mBaselineSamplePerPixel", _I, 1),
    ("mVisualizeTotalTransmittance", _I, 0), ("mUseSurfaceScene", _I, 0), ("mUsePrevVolumeForReproj", _I, 1),
    ("mInitialBaseMipLevel", _I, 1), ("mInitialM", _I, 4), ("mInitialLightSamples", _I, 1),
    ("mInitialLightingMipLevel", _I, 2), ("mInitialVisibilityUseLinearSampler", _I, 0),
    ("mInitialLightingUseLinearSampler", _I, 1), ("mInitialLightingTrackingMethod", _U, kRayMarching),
    ("mInitialVisibilityTStepScale", _F, 1.0), ("mInitialLightingTStepScale", _F, 2.0),
    ("mInitialUseRussianRoulette", _I, 1), ("mInitialUseCoarserGridForIndirectBounce", _I, 1),
    ("mTemporalReuseMThreshold", _F, 4.0), ("mTemporalReprojectionMode", _U, kReprojectionLinear),
    ("mTemporalMISMethod", _U, kMISTalbot), ("mTemporalReprojectionMipLevel", _I, 1),
    ("mSpatialReuseRounds", _I, 1), ("mSpatialVisibilityMipLevel", _I, 1), ("mSpatialLightingMipLevel", _I, 1),
    ("mSpatialVisibilityUseLinearSampler", _I, 1), ("mSpatialLightingUseLinearSampler", _I, 1),
    ("mSpatialVisibilityTStepScale", _F, 1.0), ("mSpatialLightingTStepScale", _F, 1.0),
    ("mSpatialVisibilityTrackingMethod", _U, kRayMarching), ("mSpatialLightingTrackingMethod", _U, kRayMarching),
    ("mRandomSamplerType", _U, kR2), ("mSampleRadius", _F, 10.0), ("mSpatialSampleCount", _I, 4),
    ("mEnableVisibilitySimilarityRejection", _I, 0), ("mSpatialMISMethod", _U, kMISTalbot),
    ("mFinalLightSamples", _I, 1), ("mFinalVisibilitySamples", _I, 1),
    ("mFinalVisibilityTrackingMethod", _U, kAnalyticTracking), ("mFinalLightTrackingMethod", _U, kAnalyticTracking),
    ("mFinalRandomSamplerType", _U, kR2), ("mFinalTStepScale", _F, 0.2),
]


class Params(C.Structure):
    _fields_ = [(n, t) for n, t, _ in PARAM_FIELDS]


class Node(C.Structure):
    _fields_ = [("pos", _I * 3), ("link", _U), ("bounds", _F * 4)]


class GridSlot(C.Structure):
    _fields_ = [
        ("valid", _I), ("top_lev", _I), ("dim", _I * 3), ("res", _I * 3), ("vdel", _F * 3), ("noderange", _I * 3),
        ("node_count", _U * 3), ("nodes", C.POINTER(Node) * 3), ("childlist", C.POINTER(_U) * 3),
        ("childlist_count", C.c_uint64 * 3), ("bmin", _F * 3), ("bmax", _F * 3), ("xform", _F * 16),
        ("invxform", _F * 16), ("world_to_medium", _F * 16), ("medium_to_world", _F * 16), ("max_value", _F),
        ("compress_scale", _F), ("atlas_format", _I), ("atlas_channels", _I), ("brick_count", _U),
        ("atlas", C.c_void_p),
    ]


class VolumeDesc(C.Structure):
    _fields_ = [
        ("sigma_t", _F), ("sigma_s", _F * 3), ("sigma_a", _F * 3), ("PhaseFunctionConstantG", _F),
        ("densityScaleFactor", _F), ("densityScaleFactorByScaling", _F), ("tStep", _F), ("hasEmission", _I),
        ("hasVelocity", _I), ("hasAnimation", _I), ("lastFrameHasEmission", _I), ("LeScale", _F),
        ("temperatureCutOff", _F), ("temperatureScale", _F), ("velocityScale", _F), ("numMips", _I),
        ("usePrevGridForReproj", _I), ("volumeWorldScaling", _F), ("superVoxelWorldSpaceDiagonalLength", _F),
        ("externalModelToWorld", _F * 16), ("externalWorldToModel", _F * 16),
    ]


class GridDesc(C.Structure):
    _fields_ = [("volume", VolumeDesc), ("slots", GridSlot * MAX_SLOTS), ("blackbody_lut", C.POINTER(_F))]


class Camera(C.Structure):
    _fields_ = [("posW", _F * 3), ("cameraU", _F * 3), ("cameraV", _F * 3), ("cameraW", _F * 3), ("viewMat", _F * 16),
                ("projMat", _F * 16), ("nearZ", _F), ("farZ", _F)]


class EnvMapDesc(C.Structure):
    _fields_ = [("texels", C.POINTER(_F)), ("width", _I), ("height", _I), ("intensity", _F), ("tint", _F * 3),
                ("transform", _F * 9), ("invTransform", _F * 9), ("prevTransform", _F * 9), ("prevInvTransform", _F * 9)]


class Light(C.Structure):
    _fields_ = [("type", _U), ("posW", _F * 3), ("dirW", _F * 3), ("intensity", _F * 3)]


class EmissiveTriangle(C.Structure):
    _fields_ = [("posW", (_F * 3) * 3), ("normal", _F * 3), ("area", _F), ("Le", _F * 3)]


class Reservoir(C.Structure):
    _fields_ = [("runningSum", _F), ("M", _F), ("depth", _F), ("p_y", _F), ("lightUV", _F * 2), ("lightID", _I),
                ("sampledPixel", _I)]


class Timings(C.Structure):
    _fields_ = [("features_ms", _F), ("initial_ms", _F), ("temporal_ms", _F), ("spatial_ms", _F), ("copy_ms", _F),
                ("final_ms", _F), ("total_ms", _F)]


class MarchTimings(C.Structure):
    _fields_ = [("spatial_cam_ms", _F), ("spatial_light_ms", _F), ("spatial_cam_tasks", _U), ("spatial_light_tasks", _U)]


class PipelineStats(C.Structure):
    _fields_ = [("adopted", C.c_uint64), ("discarded", C.c_uint64), ("prefetch_ms", _F), ("deferred_final_ms", _F)]


class TonemapSettings(C.Structure):
    _fields_ = [("exposureCompensation", _F), ("autoExposure", _I), ("filmSpeed", _F), ("whiteBalance", _I), ("whitePoint", _F), ("op", _U),
                ("clamp", _I), ("whiteMaxLuminance", _F), ("whiteScale", _F), ("fNumber", _F), ("shutter", _F)]


class TonemapParams(C.Structure):
    _fields_ = [("op", _U), ("autoExposure", _I), ("clamp", _I), ("whiteScale", _F), ("whiteMaxLuminance", _F), ("colorTransform", _F * 9)]


class MipLevel(C.Structure):
    _fields_ = [("data", C.c_void_p), ("bytes", C.c_size_t), ("dim", _I * 3), ("format", _I), ("max_value", _F)]


class SceneParams(C.Structure):
    _fields_ = [("kind", _I), ("dim", _I * 3), ("num_mips", _I), ("seed", _U), ("frame_time", _F), ("sigma_a", _F * 3),
                ("sigma_s", _F * 3), ("g", _F), ("density_scale", _F), ("voxel_size", _F), ("world_translation", _F * 3),
                ("world_scaling", _F), ("with_temperature", _I), ("with_velocity", _I), ("LeScale", _F),
                ("temperatureCutOff", _F), ("temperatureScale", _F)]


# every symbol include/vrestir.h declares (checked by tests/test_capi_symbols.py against the header text)
SYMBOLS = [
    "vrestir_last_error", "vrestir_version", "vrestir_default_params", "vrestir_create", "vrestir_destroy",
    "vrestir_set_volume", "vrestir_advance_volume", "vrestir_volume_frame_add", "vrestir_advance_volume_resident", "vrestir_volume_frames_clear", "vrestir_set_camera", "vrestir_set_envmap",
    "vrestir_set_analytic_lights", "vrestir_set_emissive_triangles", "vrestir_get_emissive_alias",
    "vrestir_get_env_alias", "vrestir_build_alias_table", "vrestir_build_env_alias", "vrestir_set_frame", "vrestir_update", "vrestir_set_params", "vrestir_get_params",
    "vrestir_set_frame_count", "vrestir_set_prev_camera", "vrestir_get_frame_count", "vrestir_execute",
    "vrestir_execute_host", "vrestir_execute_host_async", "vrestir_host_wait", "vrestir_execute_stage", "vrestir_set_next_camera", "vrestir_get_pipeline_stats", "vrestir_wait_output", "vrestir_get_timings", "vrestir_get_march_timings", "vrestir_debug_read_bandwidth", "vrestir_debug_long_rays", "vrestir_debug_wavefront_counters", "vrestir_get_launch_count",
    "vrestir_buffer_bytes", "vrestir_get_buffer", "vrestir_set_buffer", "vrestir_device_buffer",
    "vrestir_spatial_input_buffer", "vrestir_scene_create", "vrestir_scene_create_template", "vrestir_make_procedural_device", "vrestir_download_volume", "vrestir_scene_create_from_dense", "vrestir_scene_destroy",
    "vrestir_scene_grid", "vrestir_scene_dense_mip", "vrestir_scene_stats", "vrestir_camera_look_at",
    "vrestir_set_volume_from_chain", "vrestir_mips_build_device", "vrestir_mips_count", "vrestir_mips_level", "vrestir_mips_destroy",
    "vrestir_accum_create", "vrestir_accum_destroy", "vrestir_accum_update", "vrestir_accum_reset", "vrestir_accum_resize",
    "vrestir_accum_frame_count", "vrestir_accum_execute", "vrestir_error_measure",
    "vrestir_tonemap_default_settings", "vrestir_tonemap_params_from_settings", "vrestir_tonemap_execute",
    "vrestir_make_sky_envmap", "vrestir_make_emissive_shell", "vrestir_make_blackbody_lut", "vrestir_scene_load_vbx", "vrestir_scene_save_vbx",
]

_lib = None


def lib():
    """Load libvrestir.so (built in-tree by __graft_entry__.build()).  Fails loudly when absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). The VolumetricReSTIR pass has no CPU / PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    L.vrestir_last_error.restype = C.c_char_p
    L.vrestir_version.restype = C.c_char_p
    L.vrestir_scene_grid.restype = C.POINTER(GridDesc)
    vp = C.c_void_p
    L.vrestir_create.argtypes = [C.POINTER(Params), C.c_int, C.POINTER(vp)]
    L.vrestir_destroy.argtypes = [vp]
    L.vrestir_set_volume.argtypes = [vp, C.POINTER(GridDesc)]
    L.vrestir_advance_volume.argtypes = [vp, C.POINTER(GridDesc)]
    L.vrestir_volume_frame_add.argtypes = [vp, C.POINTER(GridDesc), C.POINTER(C.c_int)]
    L.vrestir_advance_volume_resident.argtypes = [vp, C.c_int]
    L.vrestir_volume_frames_clear.argtypes = [vp]
    L.vrestir_set_camera.argtypes = [vp, C.POINTER(Camera)]
    L.vrestir_set_prev_camera.argtypes = [vp, C.POINTER(Camera)]
    L.vrestir_set_envmap.argtypes = [vp, C.POINTER(EnvMapDesc)]
    L.vrestir_set_analytic_lights.argtypes = [vp, C.POINTER(Light), C.c_int]
    L.vrestir_set_emissive_triangles.argtypes = [vp, C.POINTER(EmissiveTriangle), C.c_int, C.c_float]
    L.vrestir_get_emissive_alias.argtypes = [vp, vp, vp, C.POINTER(C.c_float)]
    L.vrestir_get_env_alias.argtypes = [vp, vp, vp, C.POINTER(C.c_int)]
    L.vrestir_build_alias_table.argtypes = [vp, C.c_int, vp, C.POINTER(C.c_float)]
    L.vrestir_build_env_alias.argtypes = [vp, C.c_int, vp, vp]
    L.vrestir_set_frame.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
    L.vrestir_update.argtypes = [vp, C.c_char_p, C.c_double]
    L.vrestir_set_params.argtypes = [vp, C.POINTER(Params)]
    L.vrestir_get_params.argtypes = [vp, C.POINTER(Params)]
    L.vrestir_set_frame_count.argtypes = [vp, C.c_int, C.c_int]
    L.vrestir_get_frame_count.argtypes = [vp, C.POINTER(C.c_int)]
    L.vrestir_execute.argtypes = [vp, vp, vp, vp]
    L.vrestir_execute_host.argtypes = [vp, vp, vp]
    L.vrestir_execute_host_async.argtypes = [vp, vp, vp]
    L.vrestir_host_wait.argtypes = [vp]
    L.vrestir_execute_stage.argtypes = [vp, C.c_int, C.c_int, vp, vp, vp]
    L.vrestir_get_timings.argtypes = [vp, C.POINTER(Timings)]
    L.vrestir_mips_build_device.argtypes = [C.c_int, vp, C.POINTER(_I * 3), C.c_int, C.POINTER(vp), vp]
    L.vrestir_mips_count.argtypes = [vp, C.POINTER(C.c_int)]
    L.vrestir_mips_level.argtypes = [vp, C.c_int, C.c_int, C.POINTER(MipLevel)]
    L.vrestir_mips_destroy.argtypes = [vp]
    L.vrestir_set_volume_from_chain.argtypes = [vp, vp, C.POINTER(GridDesc), C.c_int]
    L.vrestir_accum_create.argtypes = [C.c_int, C.c_int, C.c_int, C.POINTER(vp)]
    L.vrestir_accum_destroy.argtypes = [vp]
    L.vrestir_accum_update.argtypes = [vp, C.c_char_p, C.c_double]
    L.vrestir_accum_reset.argtypes = [vp]
    L.vrestir_accum_resize.argtypes = [vp, C.c_int, C.c_int]
    L.vrestir_accum_frame_count.argtypes = [vp, C.POINTER(C.c_int)]
    L.vrestir_accum_execute.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
    L.vrestir_error_measure.argtypes = [C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.POINTER(C.c_float * 4), vp]
    L.vrestir_tonemap_default_settings.argtypes = [C.POINTER(TonemapSettings)]
    L.vrestir_tonemap_default_settings.restype = None
    L.vrestir_tonemap_params_from_settings.argtypes = [C.POINTER(TonemapSettings), C.POINTER(TonemapParams)]
    L.vrestir_tonemap_execute.argtypes = [C.c_int, C.POINTER(TonemapParams), vp, vp, C.c_int, C.c_int, C.POINTER(C.c_float), vp]
    L.vrestir_set_next_camera.argtypes = [vp, C.POINTER(Camera)]
    L.vrestir_get_pipeline_stats.argtypes = [vp, C.POINTER(PipelineStats)]
    L.vrestir_wait_output.argtypes = [vp, vp]
    L.vrestir_get_march_timings.argtypes = [vp, C.POINTER(MarchTimings)]
    L.vrestir_debug_read_bandwidth.argtypes = [C.c_int, C.c_size_t, C.c_int, C.POINTER(C.c_float)]
    L.vrestir_debug_long_rays.argtypes = [vp, vp, C.POINTER(C.c_uint32)]
    L.vrestir_debug_wavefront_counters.argtypes = [vp, C.POINTER(C.c_uint32)]
    L.vrestir_get_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.vrestir_buffer_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
    L.vrestir_get_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.vrestir_set_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t]
    L.vrestir_device_buffer.argtypes = [vp, C.c_int, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_int)]
    L.vrestir_spatial_input_buffer.argtypes = [vp, C.c_int, C.POINTER(C.c_int)]
    L.vrestir_scene_create.argtypes = [C.POINTER(SceneParams), C.POINTER(vp)]
    L.vrestir_scene_create_template.argtypes = [C.POINTER(SceneParams), C.POINTER(vp)]
    L.vrestir_make_procedural_device.argtypes = [C.c_int, C.POINTER(SceneParams), vp, vp]
    L.vrestir_download_volume.argtypes = [vp, C.POINTER(vp)]
    L.vrestir_scene_create_from_dense.argtypes = [C.POINTER(SceneParams), vp, vp, vp, C.POINTER(vp)]
    L.vrestir_scene_destroy.argtypes = [vp]
    L.vrestir_scene_grid.argtypes = [vp]
    L.vrestir_scene_dense_mip.argtypes = [vp, C.c_int, C.c_int, vp, C.POINTER(_I * 3)]
    L.vrestir_scene_stats.argtypes = [vp, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    L.vrestir_camera_look_at.argtypes = [C.POINTER(_F * 3), C.POINTER(_F * 3), C.POINTER(_F * 3), C.c_float, C.c_float,
                                         C.c_float, C.c_float, C.POINTER(Camera)]
    L.vrestir_make_sky_envmap.argtypes = [C.c_int, C.c_int, C.c_uint32, vp]
    L.vrestir_make_emissive_shell.argtypes = [C.c_int, C.c_uint32, C.POINTER(_F * 3), C.c_float,
                                              C.POINTER(EmissiveTriangle)]
    L.vrestir_make_blackbody_lut.argtypes = [vp]
    L.vrestir_scene_load_vbx.argtypes = [C.c_char_p, C.c_int, C.POINTER(SceneParams), C.POINTER(vp)]
    L.vrestir_scene_save_vbx.argtypes = [vp, C.c_char_p]
    _lib = L
    return L


class VRestirError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vrestir error {code}: {msg}")
        self.code = code


def check(rc):
    """Raise on negative status (errors); return positive status (warnings) to the caller."""
    if rc < 0:
        raise VRestirError(rc, lib().vrestir_last_error().decode())
    return rc
