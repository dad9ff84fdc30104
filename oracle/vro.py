"""ctypes wrapper of oracle/libvr_oracle.so.  TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package."""
import ctypes as C
import os
import subprocess

import numpy as np

from volumetricrestirrelease_b200 import _capi as capi   # struct layouts of include/vrestir.h only

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libvr_oracle.so")
REF_PRNG_PATH = os.path.join(_HERE, "_ref", "libxoshiro_ref.so")


class Counters(C.Structure):
    _fields_ = [("density_taps", C.c_uint64), ("voxels_fetched", C.c_uint64), ("voxel_bytes", C.c_uint64),
                ("node_visits", C.c_uint64), ("rng_draws", C.c_uint64), ("marches", C.c_uint64)]


def build():
    subprocess.check_call(["make", "-C", _HERE, "-s"])


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            build()
        L = C.CDLL(LIB_PATH)
        L.vro_last_error.restype = C.c_char_p
        L.vro_transmittance.restype = C.c_float
        L.vro_density_world.restype = C.c_float
        L.vro_morton.restype = C.c_uint32
        vp = C.c_void_p
        L.vro_create.argtypes = [C.POINTER(capi.Params), C.POINTER(vp)]
        for name in ("vro_destroy", "vro_get_threads"):
            getattr(L, name).argtypes = [vp]
        L.vro_set_threads.argtypes = [vp, C.c_int]
        L.vro_set_volume.argtypes = [vp, C.POINTER(capi.GridDesc)]
        L.vro_advance_volume.argtypes = [vp, C.POINTER(capi.GridDesc)]
        L.vro_set_camera.argtypes = [vp, C.POINTER(capi.Camera)]
        L.vro_set_envmap.argtypes = [vp, C.POINTER(capi.EnvMapDesc), vp]
        L.vro_set_env_alias.argtypes = [vp, vp, vp, vp, C.c_int]
        L.vro_set_analytic_lights.argtypes = [vp, C.POINTER(capi.Light), C.c_int]
        L.vro_set_emissive_triangles.argtypes = [vp, C.POINTER(capi.EmissiveTriangle), C.c_int, vp, vp, C.c_float, C.c_float]
        L.vro_set_frame.argtypes = [vp, C.c_int, C.c_int]
        L.vro_set_crop.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int]
        L.vro_update.argtypes = [vp, C.c_char_p, C.c_double]
        L.vro_set_params.argtypes = [vp, C.POINTER(capi.Params)]
        L.vro_get_params.argtypes = [vp, C.POINTER(capi.Params)]
        L.vro_set_frame_count.argtypes = [vp, C.c_int, C.c_int]
        L.vro_get_frame_count.argtypes = [vp, C.POINTER(C.c_int)]
        L.vro_execute.argtypes = [vp, vp, vp]
        L.vro_execute_stage.argtypes = [vp, C.c_int, C.c_int, vp, vp]
        L.vro_buffer_bytes.argtypes = [vp, C.c_int, C.POINTER(C.c_size_t)]
        L.vro_get_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.vro_set_buffer.argtypes = [vp, C.c_int, vp, C.c_size_t]
        L.vro_get_counters.argtypes = [vp, C.POINTER(Counters), C.c_int]
        L.vro_get_stage_ms.argtypes = [vp, C.POINTER(capi.Timings)]
        L.vro_rng_words.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, vp, vp]
        L.vro_morton.argtypes = [C.c_uint32, C.c_uint32]
        L.vro_transmittance.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.c_float, C.c_uint32, C.c_uint32, C.c_uint32]
        L.vro_density_world.argtypes = [vp, C.POINTER(C.c_float * 3), C.c_int]
        L.vro_dump_brick_visits.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_int, C.c_int, C.c_int, vp, vp]
        L.vro_env_eval.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3)]
        L.vro_env_sample.argtypes = [vp, C.c_float, C.c_float, C.POINTER(C.c_float * 3), C.POINTER(C.c_float), C.POINTER(C.c_float * 3)]
        L.vro_neighbor_offsets.argtypes = [vp, C.c_int, C.c_int, vp]
        L.vro_p_hat.argtypes = [vp, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_int]; L.vro_p_hat.restype = C.c_float
        L.vro_record_k1_generator.argtypes = [vp, C.c_int]
        L.vro_get_k1_generator.argtypes = [vp, C.c_int, C.c_int, vp]
        L.vro_visibility_state.restype = C.c_float
        L.vro_visibility_state.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_float, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                           C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.vro_sample_supervoxel.restype = C.c_float
        L.vro_sample_supervoxel.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.vro_sample_distances.argtypes = [vp, C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 3), C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_uint32, vp, vp]
        L.vro_phase_hg.argtypes = [C.c_float, C.c_float]; L.vro_phase_hg.restype = C.c_float
        L.vro_sample_phase.argtypes = [C.c_float, C.POINTER(C.c_float * 3), C.c_float, C.c_float, C.POINTER(C.c_float * 3)]; L.vro_sample_phase.restype = C.c_float
        L.vro_encode_wi_dist.argtypes = [C.POINTER(C.c_float * 4), C.POINTER(C.c_float * 3)]
        L.vro_decode_wi_dist.argtypes = [C.POINTER(C.c_float * 3), C.POINTER(C.c_float * 4)]
        _lib = L
    return _lib


def phase_hg(cos_theta, g):
    return float(lib().vro_phase_hg(cos_theta, g))


def sample_phase(g, wo, u0, u1):
    w = (C.c_float * 3)(*wo); out = (C.c_float * 3)()
    pdf = float(lib().vro_sample_phase(g, C.byref(w), u0, u1, C.byref(out)))
    return np.array(out[:], dtype=np.float32), pdf


def check(rc):
    if rc < 0:
        raise RuntimeError(f"oracle error {rc}: {lib().vro_last_error().decode()}")
    return rc


class OraclePass:
    """Same call sequence as the product pass (setScene / execute / updateDict / buffers) on the CPU oracle."""

    def __init__(self, params, threads=0):
        self._h = C.c_void_p()
        cp = params.to_c()
        check(lib().vro_create(C.byref(cp), C.byref(self._h)))
        lib().vro_set_threads(self._h, threads)
        self._scene = None
        self._frame = None

    def threads(self):
        return lib().vro_get_threads(self._h)

    def setScene(self, scene, width, height, importance=None, emissive_alias=None, env_alias=None):
        L = lib()
        self._scene, self._frame = scene, (width, height)
        check(L.vro_set_frame(self._h, width, height))
        check(L.vro_set_volume(self._h, scene.volume.grid))
        self.updateCamera()
        env = scene.envmap_desc()
        if env is not None:
            imp = None if importance is None else np.ascontiguousarray(importance, dtype=np.float32)
            check(L.vro_set_envmap(self._h, C.byref(env), None if imp is None else imp.ctypes.data))
            if env_alias is not None:
                thr, red = env_alias
                check(L.vro_set_env_alias(self._h, thr.ctypes.data, red.ctypes.data, None, len(thr)))
        arr, n = scene.lights_array()
        check(L.vro_set_analytic_lights(self._h, arr, n))
        if scene.emissiveTriangles is not None:
            items, weights, ws = emissive_alias
            check(L.vro_set_emissive_triangles(self._h, scene.emissiveTriangles, len(scene.emissiveTriangles), items.ctypes.data,
                                               weights.ctypes.data, float(ws), float(scene.emissiveIntensityMultiplier)))

    def updateCamera(self):
        cam = self._scene.camera.data(*self._frame)
        check(lib().vro_set_camera(self._h, C.byref(cam)))
        return cam

    def advanceVolume(self, volume):
        self._scene.volume = volume
        check(lib().vro_advance_volume(self._h, volume.grid))

    def set_crop(self, x0, y0, x1, y1):
        check(lib().vro_set_crop(self._h, x0, y0, x1, y1))

    def updateDict(self, d):
        for k, v in d.items():
            if k == "mParams":
                cp = v.to_c()
                check(lib().vro_set_params(self._h, C.byref(cp)))
            else:
                check(lib().vro_update(self._h, k.encode(), float(v)))

    def frame_count(self):
        fc = C.c_int()
        check(lib().vro_get_frame_count(self._h, C.byref(fc)))
        return fc.value

    def set_frame_count(self, fc, acc=1):
        check(lib().vro_set_frame_count(self._h, fc, acc))

    def execute(self, want_mvec=False):
        w, h = self._frame
        color = np.zeros((h, w, 4), dtype=np.float32)
        mvec = np.zeros((h, w, 2), dtype=np.float32) if want_mvec else None
        check(lib().vro_execute(self._h, color.ctypes.data, None if mvec is None else mvec.ctypes.data))
        return (color, mvec) if want_mvec else color

    def execute_stage(self, stage, arg=0, color=None, mvec=None):
        check(lib().vro_execute_stage(self._h, stage, arg, None if color is None else color.ctypes.data,
                                      None if mvec is None else mvec.ctypes.data))

    def get_buffer(self, buffer):
        n = C.c_size_t()
        check(lib().vro_buffer_bytes(self._h, buffer, C.byref(n)))
        raw = np.zeros(n.value, dtype=np.uint8)
        check(lib().vro_get_buffer(self._h, buffer, raw.ctypes.data, raw.size))
        return raw

    def set_buffer(self, buffer, raw):
        raw = np.ascontiguousarray(raw).view(np.uint8).ravel()
        check(lib().vro_set_buffer(self._h, buffer, raw.ctypes.data, raw.size))

    def counters(self, reset=True):
        c = Counters()
        check(lib().vro_get_counters(self._h, C.byref(c), int(reset)))
        return {n: getattr(c, n) for n, _ in Counters._fields_}

    def stage_ms(self):
        t = capi.Timings()
        check(lib().vro_get_stage_ms(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in capi.Timings._fields_}

    def transmittance(self, origin, direction, tmax, method, mip, linear=True, tstep_scale=1.0, seed=(0, 0, 0)):
        o = (C.c_float * 3)(*origin)
        d = (C.c_float * 3)(*direction)
        return float(lib().vro_transmittance(self._h, C.byref(o), C.byref(d), tmax, method, mip, int(linear), tstep_scale, *seed))

    def env_eval(self, direction):
        d = (C.c_float * 3)(*direction); out = (C.c_float * 3)()
        lib().vro_env_eval(self._h, C.byref(d), C.byref(out))
        return np.array(out[:], dtype=np.float32)

    def env_sample(self, u0, u1):
        """(direction, pdf, Le) of EnvMapSampler::sample for the random pair (u0, u1)."""
        d = (C.c_float * 3)(); le = (C.c_float * 3)(); pdf = C.c_float()
        lib().vro_env_sample(self._h, u0, u1, C.byref(d), C.byref(pdf), C.byref(le))
        return np.array(d[:], dtype=np.float32), float(pdf.value), np.array(le[:], dtype=np.float32)

    def sample_distances(self, origin, direction, mip, linear, n, seed):
        """SampleMediumAnalyticGeneric along one ray: (hit distances[4], pdfs[4], transmittances[4], generator state after)."""
        o = (C.c_float * 3)(*origin); d = (C.c_float * 3)(*direction)
        out = np.zeros(12, dtype=np.float32); st = np.zeros(4, dtype=np.uint32)
        lib().vro_sample_distances(self._h, C.byref(o), C.byref(d), mip, int(linear), n, seed[0], seed[1], seed[2], out.ctypes.data, st.ctypes.data)
        return out[0:4], out[4:8], out[8:12], st

    def record_k1_generator(self, on=True):
        """From the next K1 on, keep every pixel's generator after the candidate loop and after the p-hat evaluation."""
        check(lib().vro_record_k1_generator(self._h, int(on)))

    def k1_generator(self, px, py):
        out = np.zeros(8, dtype=np.uint32)
        check(lib().vro_get_k1_generator(self._h, px, py, out.ctypes.data))
        return [int(x) for x in out[:4]], [int(x) for x in out[4:]]

    def visibility_state(self, origin, direction, tmax, method, mip, linear, tstep_scale, samples, seed):
        """computeVisibility with `samples` estimates: (value, generator state after)."""
        o = (C.c_float * 3)(*origin); d = (C.c_float * 3)(*direction)
        st = np.zeros(4, dtype=np.uint32)
        if len(seed) == 4:                                  # a raw generator state instead of (pixel x, pixel y, sample number)
            st[:] = seed
            seed = (0xFFFFFFFF, 0xFFFFFFFF, 0)
        v = float(lib().vro_visibility_state(self._h, C.byref(o), C.byref(d), tmax, method, mip, int(linear), tstep_scale, samples, seed[0], seed[1], seed[2], st.ctypes.data))
        return v, st

    def sample_supervoxel(self, origin, direction, mip, seed):
        """SampleMediumSuperVoxelGeneric along one ray: (distance or None, generator state after)."""
        o = (C.c_float * 3)(*origin); d = (C.c_float * 3)(*direction)
        st = np.zeros(4, dtype=np.uint32)
        t = float(lib().vro_sample_supervoxel(self._h, C.byref(o), C.byref(d), mip, seed[0], seed[1], seed[2], st.ctypes.data))
        return (None if t < 0 else t), st

    def p_hat(self, px, py, depth, light_uv, light_id):
        """evaluate_P_hat of a single-bounce reservoir of pixel (px, py) under the spatial sampling options."""
        return float(lib().vro_p_hat(self._h, px, py, depth, light_uv[0], light_uv[1], light_id))

    def density_world(self, pos, mip=0):
        o = (C.c_float * 3)(*pos)
        return float(lib().vro_density_world(self._h, C.byref(o), mip))

    def brick_visits(self, origin, direction, mip, vertex_center=False, max_cells=4096):
        o = (C.c_float * 3)(*origin)
        d = (C.c_float * 3)(*direction)
        xyz = np.zeros((max_cells, 3), dtype=np.int32)
        t = np.zeros(max_cells, dtype=np.float32)
        n = lib().vro_dump_brick_visits(self._h, C.byref(o), C.byref(d), mip, int(vertex_center), max_cells, xyz.ctypes.data, t.ctypes.data)
        return xyz[:n], t[:n]

    def neighbor_offsets(self, frame_count, rnd, count):
        out = np.zeros((count, 2), dtype=np.int32)
        lib().vro_neighbor_offsets(self._h, frame_count, rnd, out.ctypes.data)
        return out

    def close(self):
        if self._h:
            lib().vro_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


RES_DTYPE = np.dtype([("runningSum", "<f4"), ("M", "<f4"), ("depth", "<f4"), ("p_y", "<f4"), ("lightUV", "<f4", 2),
                      ("lightID", "<i4"), ("sampledPixel", "<i4")])
FEAT_DTYPE = np.dtype([("noReflectiveSurface", "<i4"), ("transmittance", "<f4")])
