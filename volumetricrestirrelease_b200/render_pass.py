"""Host-side mirror of the reference's ``VolumetricReSTIR`` render pass (VR/VolumetricReSTIR.{h,cpp}) over the C ABI.

Same names and argument meaning as the plugin interface the reference exposes to the render graph and to scripts:
``VolumetricReSTIRParams(...)`` (VR/VolumetricReSTIR.cpp:1349-1430), ``VolumetricReSTIR.create(dict)`` (:66-140),
``reflect`` (:149-155), ``setScene`` (:1218-1256), ``execute`` (:303-772), ``updateDict`` (:1280-1347),
``getScriptingDictionary`` (:142-147).  Error behaviour: unknown dictionary keys warn (VR/VolumetricReSTIR.h:309),
out-of-scope options raise.
"""
import ctypes as C
import warnings

import numpy as np

from . import _capi as capi

kOutputChannels = {"accumulated_color": "RGBA32Float", "mvec": "RG32Float"}   # VR/VolumetricReSTIR.cpp:39-43

TOP_LEVEL_KEYS = ("mOutputMotionVec", "mFreezeFrame", "volumeDensityScaleExtraControl", "volumeAlbedoExtraControl",
                  "volumeAnisotropyExtraControl", "mEnvSamplerType", "mUseWavefront", "mInitialMode", "mOverlapFeatures", "mPipelineFrames", "mPrefetchPriority", "mScratchBudgetMB", "mDebugPoisonResults", "mPrimaryDistanceEngine")
# accepted for script compatibility, camera / env-light animation and UI live outside the hot path
IGNORED_KEYS = ("mCameraMoveScale", "mCameraForwardScale", "mCameraFrameInterval", "mCameraPauseInterval",
                "mCameraShakeTotalRounds", "mCameraShakeRoundsBeforePause", "mCameraAnimationMode", "mAnimateEnvLight",
                "mAnimationFreezedFrame", "mEnvLightRotationSpeed", "mVolumeAnimationSelectedFrameId",
                "mEmissiveSamplerTypeId")


class VolumetricReSTIRParams:
    """Same fields and defaults as ``VolumetricReSTIR::VolumetricReSTIRParams`` (VR/VolumetricReSTIR.h:130-207)."""

    _names = [n for n, _, _ in capi.PARAM_FIELDS]

    def __init__(self, **kwargs):
        for n, _, default in capi.PARAM_FIELDS:
            setattr(self, n, default)
        for k, v in kwargs.items():
            if k not in self._names:
                raise AttributeError(f"VolumetricReSTIRParams has no field '{k}'")
            setattr(self, k, v)

    def to_c(self):
        p = capi.Params()
        for n, t, _ in capi.PARAM_FIELDS:
            v = getattr(self, n)
            setattr(p, n, float(v) if t is C.c_float else int(v))
        return p

    @classmethod
    def from_c(cls, p):
        o = cls()
        for n in cls._names:
            setattr(o, n, getattr(p, n))
        return o

    def as_dict(self):
        return {n: getattr(self, n) for n in self._names}

    def __repr__(self):
        return "VolumetricReSTIRParams(" + ", ".join(f"{n}={getattr(self, n)!r}" for n in self._names) + ")"


class VolumetricReSTIR:
    """One pass instance = one ``vrestir_pass`` handle on one GPU (optionally one row band of the frame)."""

    def __init__(self, dict_=None, device=0):
        dict_ = dict(dict_ or {})
        params = dict_.pop("mParams", None) or VolumetricReSTIRParams()
        self._lib = capi.lib()
        self._h = C.c_void_p()
        cp = params.to_c()
        capi.check(self._lib.vrestir_create(C.byref(cp), int(device), C.byref(self._h)))
        self.device = device
        self._scene = None
        self._frame = None
        self._keep = []
        self._apply_dict(dict_)

    def wavefront_counters(self):
        out = (C.c_uint32 * 16)()
        capi.check(self._lib.vrestir_debug_wavefront_counters(self._h, out))
        return list(out)

    # RenderPassLibrary registers `create(RenderContext*, const Dictionary&)` (VR/VolumetricReSTIR.cpp:61-70)
    @classmethod
    def create(cls, dict_=None, device=0):
        return cls(dict_, device)

    def reflect(self):
        return dict(kOutputChannels)

    # ------------------------------------------------------------------------------------------------ dictionary
    def _apply_dict(self, d):
        for k, v in d.items():
            if k == "mParams":
                cp = v.to_c()
                capi.check(self._lib.vrestir_set_params(self._h, C.byref(cp)))
            elif k in IGNORED_KEYS:
                continue
            else:
                rc = capi.check(self._lib.vrestir_update(self._h, k.encode(), float(v)))
                if rc == capi.WARN_UNKNOWN_KEY:
                    warnings.warn(f"Unknown field '{k}' in a VolumetricReSTIR dictionary")

    def updateDict(self, d):
        """Live option update; resets the frame counter and the temporal history (VR/VolumetricReSTIR.cpp:1339)."""
        d = dict(d)
        if not d:
            capi.check(self._lib.vrestir_update(self._h, b"mMaxBounces", float(self.params.mMaxBounces)))
        self._apply_dict(d)

    def getScriptingDictionary(self):
        return {"mParams": self.params}

    @property
    def params(self):
        cp = capi.Params()
        capi.check(self._lib.vrestir_get_params(self._h, C.byref(cp)))
        return VolumetricReSTIRParams.from_c(cp)

    # ------------------------------------------------------------------------------------------------ scene
    def setScene(self, scene, width, height, row_begin=0, row_end=None):
        self._scene = scene
        self._frame = (int(width), int(height))
        L = self._lib
        capi.check(L.vrestir_set_frame(self._h, int(width), int(height), int(row_begin), int(height if row_end is None else row_end)))
        if getattr(scene.volume, "chain", None) is not None:     # device-built volume: bind the GPU chain over the voxel-less template
            capi.check(L.vrestir_set_volume_from_chain(self._h, scene.volume.chain._h, scene.volume.grid, 0))
        else:
            capi.check(L.vrestir_set_volume(self._h, scene.volume.grid))
        self.updateCamera()
        env = scene.envmap_desc()
        if env is not None:
            capi.check(L.vrestir_set_envmap(self._h, C.byref(env)))
        arr, n = scene.lights_array()
        capi.check(L.vrestir_set_analytic_lights(self._h, arr, n))
        if scene.emissiveTriangles is not None:
            capi.check(L.vrestir_set_emissive_triangles(self._h, scene.emissiveTriangles, len(scene.emissiveTriangles),
                                                        float(scene.emissiveIntensityMultiplier)))

    def setRowBand(self, row_begin, row_end):
        """Restrict the pass to rows [row_begin, row_end) of the frame (multi-GPU row sharding)."""
        w, h = self._frame
        capi.check(self._lib.vrestir_set_frame(self._h, w, h, int(row_begin), int(row_end)))

    def updateCamera(self):
        cam = self._scene.camera.data(*self._frame)
        capi.check(self._lib.vrestir_set_camera(self._h, C.byref(cam)))
        return cam

    def setVolumeFromChain(self, chain, advance=False, template=None):
        """Bind the density grids to a GPU-built mip chain (``mipbuild.build_mips``) without moving the voxels through the host.
        ``template``: a host-built Volume of the same dimensions (default: the scene's volume) for transforms / description;
        ``advance=True`` = ``advanceVolume`` semantics (current grids become the previous frame's)."""
        tmpl = (template or self._scene.volume).grid
        capi.check(self._lib.vrestir_set_volume_from_chain(self._h, chain._h, tmpl, int(advance)))

    def downloadVolume(self):
        """Host copy (a scene ``Volume``) of the grids bound on the device, e.g. to hand a device-built volume to a CPU checker."""
        from .scene import Volume
        h = C.c_void_p()
        capi.check(self._lib.vrestir_download_volume(self._h, C.byref(h)))
        return Volume(h)

    def setNextCamera(self, camera=None):
        """Frame pipelining ("mPipelineFrames"): announce the camera of the NEXT frame before executing the current one, so that
        its K0/K1 can run ahead (include/vrestir.h).  `camera`: a scene Camera, or None to clear (= the camera stays put)."""
        if camera is None:
            capi.check(self._lib.vrestir_set_next_camera(self._h, None))
            return
        cam = camera.data(*self._frame)
        capi.check(self._lib.vrestir_set_next_camera(self._h, C.byref(cam)))

    def wait_output(self, stream=None):
        """"mPipelineFrames" = 2 (deferred final shading): make `stream` (a raw cudaStream_t; None = the default stream) wait
        until the image of the last execute() is complete.  A no-op at the other levels."""
        capi.check(self._lib.vrestir_wait_output(self._h, C.c_void_p(stream) if stream else None))

    def pipeline_stats(self):
        t = capi.PipelineStats()
        capi.check(self._lib.vrestir_get_pipeline_stats(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in capi.PipelineStats._fields_}

    def advanceVolume(self, volume):
        """Animated sequences: current grids become the prev-frame slots (F/Scene/Scene.cpp:825-863)."""
        self._keep = [self._scene.volume, volume]
        self._scene.volume = volume
        capi.check(self._lib.vrestir_advance_volume(self._h, volume.grid))

    def addVolumeFrame(self, volume):
        """Uploads one frame of an animated sequence and keeps it on the device (the reference holds all frames of a sequence
        resident: F/Scene/Scene.cpp:825-863); returns its index for advanceVolumeResident."""
        idx = C.c_int(-1)
        capi.check(self._lib.vrestir_volume_frame_add(self._h, volume.grid, C.byref(idx)))
        self._frames = getattr(self, "_frames", [])
        self._frames.append(volume)
        return idx.value

    def advanceVolumeResident(self, index):
        """advanceVolume with a resident frame as the new volume: pointers are rebound, nothing is copied."""
        capi.check(self._lib.vrestir_advance_volume_resident(self._h, int(index)))
        self._scene.volume = self._frames[index]

    def clearVolumeFrames(self):
        capi.check(self._lib.vrestir_volume_frames_clear(self._h))
        self._frames = []

    # ------------------------------------------------------------------------------------------------ execution
    def execute(self, out_color_ptr, out_mvec_ptr=None, stream=None):
        """Device-pointer path: `out_color_ptr` is a CUDA address of width*height float4 (e.g. tensor.data_ptr())."""
        capi.check(self._lib.vrestir_execute(self._h, C.c_void_p(out_color_ptr),
                                             C.c_void_p(out_mvec_ptr) if out_mvec_ptr else None,
                                             C.c_void_p(stream) if stream else None))

    def execute_stage(self, stage, arg=0, out_color_ptr=None, out_mvec_ptr=None, stream=None):
        capi.check(self._lib.vrestir_execute_stage(self._h, int(stage), int(arg),
                                                   C.c_void_p(out_color_ptr) if out_color_ptr else None,
                                                   C.c_void_p(out_mvec_ptr) if out_mvec_ptr else None,
                                                   C.c_void_p(stream) if stream else None))

    def execute_host(self, out_color=None, out_mvec=None):
        """Host-buffer path (what a CPU-side caller of the plugin sees): returns (H, W, 4) float32."""
        w, h = self._frame
        if out_color is None:
            out_color = np.zeros((h, w, 4), dtype=np.float32)
        capi.check(self._lib.vrestir_execute_host(self._h, out_color.ctypes.data,
                                                  out_mvec.ctypes.data if out_mvec is not None else None))
        return out_color

    def execute_host_async(self, out_color_ptr, out_mvec_ptr=None):
        """Host-buffer path without blocking: `out_color_ptr` is the address of a (pinned) host buffer of width*height float4;
        the read-back of this frame overlaps the next frame.  Call host_wait() before reading the buffer."""
        capi.check(self._lib.vrestir_execute_host_async(self._h, C.c_void_p(out_color_ptr), C.c_void_p(out_mvec_ptr) if out_mvec_ptr else None))

    def host_wait(self):
        capi.check(self._lib.vrestir_host_wait(self._h))

    # ------------------------------------------------------------------------------------------------ introspection
    def timings(self):
        t = capi.Timings()
        capi.check(self._lib.vrestir_get_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in capi.Timings._fields_}

    def march_timings(self):
        t = capi.MarchTimings()
        capi.check(self._lib.vrestir_get_march_timings(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in capi.MarchTimings._fields_}

    def debug_long_rays(self):
        out = np.zeros((64, 8), dtype=np.float32)
        n = C.c_uint32()
        capi.check(self._lib.vrestir_debug_long_rays(self._h, out.ctypes.data, C.byref(n)))
        return out[:min(64, n.value)], int(n.value)

    def launch_count(self):
        n = C.c_uint64()
        capi.check(self._lib.vrestir_get_launch_count(self._h, C.byref(n)))
        return int(n.value)

    def set_frame_count(self, frame_count, temporal_sample_accumulated=1):
        capi.check(self._lib.vrestir_set_frame_count(self._h, int(frame_count), int(temporal_sample_accumulated)))

    def frame_count(self):
        n = C.c_int()
        capi.check(self._lib.vrestir_get_frame_count(self._h, C.byref(n)))
        return n.value

    def set_prev_camera(self, cam):
        capi.check(self._lib.vrestir_set_prev_camera(self._h, C.byref(cam)))

    def _buffer_array(self, buffer):
        n = C.c_size_t()
        capi.check(self._lib.vrestir_buffer_bytes(self._h, buffer, C.byref(n)))
        return np.zeros(n.value, dtype=np.uint8)

    def get_buffer(self, buffer):
        raw = self._buffer_array(buffer)
        capi.check(self._lib.vrestir_get_buffer(self._h, buffer, raw.ctypes.data, raw.size))
        return raw

    def set_buffer(self, buffer, raw):
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        capi.check(self._lib.vrestir_set_buffer(self._h, buffer, raw.ctypes.data, raw.size))

    def device_buffer(self, buffer):
        base, stride, planes = C.c_void_p(), C.c_size_t(), C.c_int()
        capi.check(self._lib.vrestir_device_buffer(self._h, buffer, C.byref(base), C.byref(stride), C.byref(planes)))
        return base.value, stride.value, planes.value

    def spatial_input_buffer(self, rnd):
        b = C.c_int()
        capi.check(self._lib.vrestir_spatial_input_buffer(self._h, int(rnd), C.byref(b)))
        return b.value

    def emissive_alias(self, count):
        items = np.zeros((count, 4), dtype=np.uint32)
        weights = np.zeros(count, dtype=np.float32)
        ws = C.c_float()
        capi.check(self._lib.vrestir_get_emissive_alias(self._h, items.ctypes.data, weights.ctypes.data, C.byref(ws)))
        return items, weights, ws.value

    def env_alias(self):
        n = C.c_int()
        capi.check(self._lib.vrestir_get_env_alias(self._h, None, None, C.byref(n)))
        thr = np.zeros(n.value, dtype=np.float32)
        red = np.zeros(n.value, dtype=np.uint32)
        capi.check(self._lib.vrestir_get_env_alias(self._h, thr.ctypes.data, red.ctypes.data, C.byref(n)))
        return thr, red

    def close(self):
        if self._h:
            self._lib.vrestir_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
