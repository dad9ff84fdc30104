"""Host-side mirrors of the two passes behind ``VolumetricReSTIR.accumulated_color`` in the reference's render graphs
(VR/Scripts/run_bunny_tree.py:9-21): ``AccumulatePass`` and ``ErrorMeasurePass`` — same names, dictionary keys and
defaults (AccumulatePass.h:91-94, ErrorMeasurePass.h:93-98).  The work happens in csrc/vr_post.cu through the C ABI."""
import ctypes as C

from . import _capi as capi

PRECISION = {"Double": 0, "Single": 1, "SingleCompensated": 2}


class AccumulatePass:
    """``createPass("AccumulatePass", {"enableAccumulation": ..., "autoReset": ..., "precisionMode": ..., "subFrameCount": ...})``"""

    KEYS = ("enableAccumulation", "autoReset", "precisionMode", "subFrameCount")

    def __init__(self, handle, width, height):
        self._h = handle
        self._lib = capi.lib()
        self._dim = (width, height)

    @classmethod
    def create(cls, d=None, width=1, height=1, device=0):
        h = C.c_void_p()
        capi.check(capi.lib().vrestir_accum_create(int(device), int(width), int(height), C.byref(h)))
        self = cls(h, width, height)
        self.updateDict(d or {})
        return self

    def updateDict(self, d):
        for k, v in dict(d).items():
            if k == "precisionMode" and isinstance(v, str):
                v = PRECISION[v]
            rc = self._lib.vrestir_accum_update(self._h, k.encode(), float(v))
            if rc == capi.WARN_UNKNOWN_KEY:
                import warnings
                warnings.warn(f"Unknown field '{k}' in an AccumulatePass dictionary")
            else:
                capi.check(rc)

    def reset(self):
        """What a scene / camera change or a refresh flag does to the reference pass (AccumulatePass.cpp:141-161)."""
        capi.check(self._lib.vrestir_accum_reset(self._h))

    def resize(self, width, height):
        capi.check(self._lib.vrestir_accum_resize(self._h, int(width), int(height)))
        self._dim = (width, height)

    @property
    def frameCount(self):
        n = C.c_int(0)
        capi.check(self._lib.vrestir_accum_frame_count(self._h, C.byref(n)))
        return n.value

    def execute(self, input_ptr, output_ptr, row_begin=0, row_end=None, stream=None):
        """input / output: device pointers of width*height float4 images (``tensor.data_ptr()``)."""
        row_end = self._dim[1] if row_end is None else row_end
        capi.check(self._lib.vrestir_accum_execute(self._h, C.c_void_p(input_ptr), C.c_void_p(output_ptr), int(row_begin), int(row_end),
                                                   C.c_void_p(stream) if stream else None))

    def __del__(self):
        try:
            if self._h:
                self._lib.vrestir_accum_destroy(self._h)
                self._h = None
        except Exception:
            pass


class ErrorMeasurePass:
    """``createPass("ErrorMeasurePass", {...})``: keys IgnoreBackground, ComputeSquaredDifference, ComputeAverage,
    ReportRunningError, RunningErrorSigma (ErrorMeasurePass.cpp:40-51).  ``execute`` returns the measurements of the frame
    and keeps the exponential moving average of ErrorMeasurePass.cpp:246-256."""

    def __init__(self, d=None, device=0):
        self.IgnoreBackground = True
        self.ComputeSquaredDifference = True
        self.ComputeAverage = False
        self.ReportRunningError = True
        self.RunningErrorSigma = 0.995
        self.device = device
        self.runningAvgError = -1.0
        self.runningError = (0.0, 0.0, 0.0)
        self.measurements = None
        self.updateDict(d or {})

    def updateDict(self, d):
        for k, v in dict(d).items():
            if k in ("IgnoreBackground", "ComputeSquaredDifference", "ComputeAverage", "ReportRunningError"):
                setattr(self, k, bool(v))
            elif k == "RunningErrorSigma":
                self.RunningErrorSigma = float(v)
            else:
                import warnings
                warnings.warn(f"Unknown field '{k}' in ErrorMeasurePass dictionary")

    def execute(self, source_ptr, reference_ptr, width, height, world_position_ptr=None, difference_ptr=None, stream=None):
        import numpy as np
        out = (C.c_float * 4)()
        capi.check(capi.lib().vrestir_error_measure(int(self.device), C.c_void_p(source_ptr), C.c_void_p(reference_ptr),
                                                    C.c_void_p(world_position_ptr) if world_position_ptr else None, int(width), int(height),
                                                    int(self.IgnoreBackground), int(self.ComputeSquaredDifference), int(self.ComputeAverage),
                                                    C.c_void_p(difference_ptr) if difference_ptr else None, C.byref(out),
                                                    C.c_void_p(stream) if stream else None))
        err = tuple(np.float32(out[i]) for i in range(3))
        avg = np.float32(out[3])
        self.measurements = {"error": err, "avgError": avg}
        s = np.float32(self.RunningErrorSigma)
        if self.runningAvgError < 0:
            self.runningError, self.runningAvgError = err, avg
        else:
            self.runningError = tuple(s * r + (np.float32(1) - s) * e for r, e in zip(self.runningError, err))
            self.runningAvgError = s * self.runningAvgError + (np.float32(1) - s) * avg
        return self.measurements


TONEMAP_OPERATORS = {"Linear": 0, "Reinhard": 1, "ReinhardModified": 2, "HejiHableAlu": 3, "HableUc2": 4, "Aces": 5}


class ToneMapper:
    """``createPass("ToneMapper", {"autoExposure": ..., "exposureCompensation": ..., "operator": ToneMapOp.Aces, ...})`` — the
    display end of the reference's graphs (Source/RenderPasses/ToneMapper/ToneMapper.cpp:36-49 keys, ToneMapper.h:111-125 defaults).
    Clamps of the setters as in ToneMapper.cpp:401-479."""

    KEYS = ("exposureCompensation", "autoExposure", "exposureValue", "filmSpeed", "whiteBalance", "whitePoint", "operator", "clamp",
            "whiteMaxLuminance", "whiteScale", "fNumber", "shutter")

    def __init__(self, d=None, device=0):
        self._lib = capi.lib()
        self.device = device
        self._s = capi.TonemapSettings()
        self._lib.vrestir_tonemap_default_settings(C.byref(self._s))
        self.avgLogLuminance = None
        for k, v in dict(d or {}).items():
            self.set(k, v)

    def set(self, key, value):
        s = self._s
        clampf = lambda v, lo, hi: max(lo, min(hi, float(v)))
        if key == "exposureCompensation": s.exposureCompensation = clampf(value, -12.0, 12.0)
        elif key == "autoExposure": s.autoExposure = int(bool(value))
        elif key == "filmSpeed": s.filmSpeed = clampf(value, 1.0, 6400.0)
        elif key == "whiteBalance": s.whiteBalance = int(bool(value))
        elif key == "whitePoint": s.whitePoint = clampf(value, 1905.0, 25000.0)
        elif key == "operator": s.op = TONEMAP_OPERATORS[value] if isinstance(value, str) else int(value)
        elif key == "clamp": s.clamp = int(bool(value))
        elif key == "whiteMaxLuminance": s.whiteMaxLuminance = float(value)
        elif key == "whiteScale": s.whiteScale = max(0.001, float(value))
        elif key == "fNumber": s.fNumber = clampf(value, 0.1, 100.0)
        elif key == "shutter": s.shutter = clampf(value, 0.1, 10000.0)
        elif key == "exposureValue":   # aperture priority (the default exposure mode): the shutter follows the EV (ToneMapper.cpp:294-303)
            import math
            ev = clampf(value, math.log2(0.1 * 0.1 * 0.1), math.log2(10000.0 * 100.0 * 100.0))
            s.shutter = clampf(2.0 ** ev / (s.fNumber * s.fNumber), 0.1, 10000.0)
        else:
            import warnings
            warnings.warn(f"Unknown field '{key}' in a ToneMapper dictionary")

    @property
    def exposureValue(self):
        import math
        return math.log2(self._s.shutter * self._s.fNumber * self._s.fNumber)

    def params(self):
        p = capi.TonemapParams()
        capi.check(self._lib.vrestir_tonemap_params_from_settings(C.byref(self._s), C.byref(p)))
        return p

    def execute(self, src_ptr, dst_ptr, width, height, stream=None, want_average=False):
        """src / dst: CUDA addresses of width*height float4.  want_average: also return the auto-exposure average log2 luminance."""
        p = self.params()
        avg = C.c_float(0.0)
        capi.check(self._lib.vrestir_tonemap_execute(int(self.device), C.byref(p), C.c_void_p(src_ptr), C.c_void_p(dst_ptr), int(width), int(height),
                                                     C.byref(avg) if (want_average and self._s.autoExposure) else None, C.c_void_p(stream) if stream else None))
        if want_average and self._s.autoExposure:
            self.avgLogLuminance = avg.value
        return self.avgLogLuminance if want_average else None
