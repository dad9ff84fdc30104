"""Host-side scene for the VolumetricReSTIR pass: the subset of Falcor's ``Scene`` the hot path consumes.

Mirrors the script-facing calls of the reference (Source/Mogwai/MogwaiScripting.cpp:127-131, VR/Scripts/run_*.py):
``addGVDBVolume`` (here fed by procedural generators because the 7.87 GB scene pack is not available offline),
``setEnvMap`` / ``setEnvMapIntensity``, ``camera.position/target/up``, analytic lights and emissive triangles.
All heavy lifting (tree/atlas/mip construction) is native code in csrc/vr_scene.cpp behind the C ABI.
"""
import ctypes as C
import math

import numpy as np

from . import _capi as capi

PROCEDURAL_KINDS = {"sphere": 0, "bunny": 1, "plume": 2, "cloud": 3, "shells": 4}


def _f3(v):
    return (C.c_float * 3)(*[float(x) for x in v])


class CameraState:
    """position / target / up like ``m.scene.camera`` in the reference scripts; fovY in degrees."""

    def __init__(self):
        self.position = (0.0, 0.0, 5.0)
        self.target = (0.0, 0.0, 0.0)
        self.up = (0.0, 1.0, 0.0)
        self.fovY = 45.0
        self.nearZ = 0.1
        self.farZ = 1000.0

    def data(self, width, height):
        cam = capi.Camera()
        capi.check(capi.lib().vrestir_camera_look_at(_f3(self.position), _f3(self.target), _f3(self.up),
                                                     math.radians(self.fovY), float(width) / float(height),
                                                     self.nearZ, self.farZ, C.byref(cam)))
        return cam


class Volume:
    """Owns one native ``vrestir_scene`` (tree + brick pools + VolumeDesc)."""

    def __init__(self, handle):
        self._h = handle

    @property
    def grid(self):
        return capi.lib().vrestir_scene_grid(self._h)

    def stats(self, slot):
        b, n = C.c_uint32(), C.c_uint64()
        capi.check(capi.lib().vrestir_scene_stats(self._h, slot, C.byref(b), C.byref(n)))
        return int(b.value), int(n.value)

    def dense_mip(self, mip, conservative=False):
        dim = (C.c_int32 * 3)()
        capi.check(capi.lib().vrestir_scene_dense_mip(self._h, mip, int(conservative), None, C.byref(dim)))
        out = np.zeros((dim[2], dim[1], dim[0]), dtype=np.float32)
        capi.check(capi.lib().vrestir_scene_dense_mip(self._h, mip, int(conservative), out.ctypes.data, C.byref(dim)))
        return out

    def save_vbx(self, dir_and_prefix):
        """Write every grid of the volume as GVDB .vbx files (<prefix>_mip<k>[c].vbx, _temperature.vbx, _velocity_{x,y,z}.vbx)."""
        capi.check(capi.lib().vrestir_scene_save_vbx(self._h, str(dir_and_prefix).encode()))

    def release_chain(self):
        """Device-built volumes: free the dense mip chain (tens of GB for a 2048^3 grid) once a pass has bound its brick pools."""
        ch = getattr(self, "chain", None)
        if ch is not None:
            ch.close()
            self.chain = None
        self.dense = None

    def close(self):
        self.release_chain()
        if self._h:
            capi.lib().vrestir_scene_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def _scene_params(kind, dim, numMips, seed, sigma_a, sigma_s, g, densityScale, voxelSize, worldTranslation,
                  worldScaling, hasVelocity, hasEmission, LeScale, temperatureCutoff, temperatureScale, frameTime):
    sp = capi.SceneParams()
    sp.kind = PROCEDURAL_KINDS[kind] if isinstance(kind, str) else int(kind)
    sp.dim[:] = [int(d) for d in dim]
    sp.num_mips = int(numMips)
    sp.seed = int(seed)
    sp.frame_time = float(frameTime)
    sp.sigma_a[:] = [float(x) for x in sigma_a]
    sp.sigma_s[:] = [float(x) for x in sigma_s]
    sp.g = float(g)
    sp.density_scale = float(densityScale)
    sp.voxel_size = float(voxelSize)
    sp.world_translation[:] = [float(x) for x in worldTranslation]
    sp.world_scaling = float(worldScaling)
    sp.with_temperature = int(bool(hasEmission))
    sp.with_velocity = int(bool(hasVelocity))
    sp.LeScale = float(LeScale)
    sp.temperatureCutOff = float(temperatureCutoff)
    sp.temperatureScale = float(temperatureScale)
    return sp


class Scene:
    def __init__(self):
        self.camera = CameraState()
        self.volume = None
        self.envMap = None          # (H, W, 4) float32
        self.envMapIntensity = 1.0
        self.envMapTint = (1.0, 1.0, 1.0)
        self.envMapRotation = (0.0, 0.0, 0.0)   # degrees XYZ like EnvMap::setRotation
        self.lights = []            # list of dicts: type, posW, dirW, intensity
        self.emissiveTriangles = None
        self.emissiveIntensityMultiplier = 1.0

    # --- volumes -------------------------------------------------------------------------------------------------
    def addGVDBVolume(self, sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile="bunny", numMips=4, densityScale=1.0,
                      hasVelocity=False, hasEmission=False, LeScale=0.005, temperatureCutoff=1.0,
                      temperatureScale=100.0, worldTranslation=(0, 0, 0), worldRotation=(0, 0, 0), worldScaling=1.0,
                      dim=(128, 128, 128), seed=1, voxelSize=1.0, frameTime=0.0, dense=None, temperature=None,
                      velocity=None):
        """Same leading arguments as ``m.addGVDBVolume`` (Source/Mogwai/MogwaiScripting.cpp:127-131).

        ``dataFile`` names a procedural generator ("sphere", "bunny", "plume", "cloud", "shells") unless ``dense`` (a
        (Z, Y, X) float32 array) is given."""
        if any(abs(r) > 0 for r in worldRotation):
            raise capi.VRestirError(capi.ERR_UNSUPPORTED, "worldRotation is not supported")
        h = C.c_void_p()
        if dense is not None:
            dense = np.ascontiguousarray(dense, dtype=np.float32)
            sp = _scene_params(0, dense.shape[::-1], numMips, seed, sigma_a, sigma_s, g, densityScale, voxelSize,
                               worldTranslation, worldScaling, velocity is not None, temperature is not None, LeScale,
                               temperatureCutoff, temperatureScale, frameTime)
            t = None if temperature is None else np.ascontiguousarray(temperature, dtype=np.float32)
            v = None if velocity is None else np.ascontiguousarray(velocity, dtype=np.float32)
            capi.check(capi.lib().vrestir_scene_create_from_dense(
                C.byref(sp), dense.ctypes.data, None if t is None else t.ctypes.data,
                None if v is None else v.ctypes.data, C.byref(h)))
        else:
            sp = _scene_params(dataFile, dim, numMips, seed, sigma_a, sigma_s, g, densityScale, voxelSize,
                               worldTranslation, worldScaling, hasVelocity, hasEmission, LeScale, temperatureCutoff,
                               temperatureScale, frameTime)
            capi.check(capi.lib().vrestir_scene_create(C.byref(sp), C.byref(h)))
        self.volume = Volume(h)
        return self.volume

    def addGVDBVolumeDevice(self, device=0, sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, dataFile="shells", numMips=4, densityScale=1.0,
                            worldTranslation=(0, 0, 0), worldScaling=1.0, dim=(2048, 2048, 2048), seed=5, voxelSize=1.0, keep_dense=False):
        """A procedural volume that only ever exists on the GPU (grids too large for the host builder):
        the density field is evaluated per voxel on the device, the mip / conservative-mip chain is built there
        (``mipbuild.build_mips``) and ``VolumetricReSTIR.setScene`` binds it with ``vrestir_set_volume_from_chain``.
        ``self.volume`` is a voxel-less template (dimensions, transforms, VolumeDesc) carrying the chain."""
        import torch
        from .mipbuild import build_mips
        sp = _scene_params(dataFile, dim, numMips, seed, sigma_a, sigma_s, g, densityScale, voxelSize, worldTranslation, worldScaling,
                           False, False, 0.005, 1.0, 100.0, 0.0)
        h = C.c_void_p()
        capi.check(capi.lib().vrestir_scene_create_template(C.byref(sp), C.byref(h)))
        vol = Volume(h)
        dev = torch.device("cuda", device)
        dense = torch.empty((int(dim[2]), int(dim[1]), int(dim[0])), dtype=torch.float32, device=dev)
        capi.check(capi.lib().vrestir_make_procedural_device(int(device), C.byref(sp), C.c_void_p(dense.data_ptr()), None))
        torch.cuda.synchronize(dev)
        vol.chain = build_mips(dense, numMips)
        vol.dense = dense if keep_dense else None
        del dense
        self.volume = vol
        return vol

    def loadGVDBVolume(self, dir_and_prefix, sigma_a=(1, 1, 1), sigma_s=(9, 9, 9), g=0.0, numMips=4, densityScale=1.0,
                       LeScale=0.005, temperatureCutoff=1.0, temperatureScale=100.0, worldTranslation=(0, 0, 0),
                       worldScaling=1.0):
        """``m.addGVDBVolume(..., dataFile=<folder>/<name>, ...)`` for real assets: reads <prefix>_mip<k>[c].vbx (+ _temperature,
        _velocity_{x,y,z}) like F/Scene/Scene.cpp:2806-2815 / GV/src/gvdb_volume_gvdb.cpp:532-739."""
        sp = _scene_params(0, (8, 8, 8), numMips, 0, sigma_a, sigma_s, g, densityScale, 1.0, worldTranslation, worldScaling,
                           False, False, LeScale, temperatureCutoff, temperatureScale, 0.0)
        h = C.c_void_p()
        capi.check(capi.lib().vrestir_scene_load_vbx(str(dir_and_prefix).encode(), int(numMips), C.byref(sp), C.byref(h)))
        self.volume = Volume(h)
        return self.volume

    # --- env map -------------------------------------------------------------------------------------------------
    def setEnvMap(self, texels_or_size=(2048, 1024), seed=7):
        """Accepts an (H, W, 3|4) float array or a (W, H) size for the procedural HDR sky."""
        if isinstance(texels_or_size, np.ndarray):
            t = np.asarray(texels_or_size, dtype=np.float32)
            if t.shape[-1] == 3:
                t = np.concatenate([t, np.ones(t.shape[:2] + (1,), np.float32)], axis=-1)
            self.envMap = np.ascontiguousarray(t)
        else:
            w, h = texels_or_size
            out = np.zeros((h, w, 4), dtype=np.float32)
            capi.check(capi.lib().vrestir_make_sky_envmap(int(w), int(h), int(seed), out.ctypes.data))
            self.envMap = out

    def setEnvMapIntensity(self, v):
        self.envMapIntensity = float(v)

    def envmap_desc(self, prevRotation=None):
        if self.envMap is None:
            return None
        d = capi.EnvMapDesc()
        d.texels = self.envMap.ctypes.data_as(C.POINTER(C.c_float))
        d.height, d.width = self.envMap.shape[0], self.envMap.shape[1]
        d.intensity = self.envMapIntensity
        d.tint[:] = self.envMapTint

        def rot(deg):   # EnvMap::setRotation: rotZ * rotY * rotX (F/Experimental/Scene/Lights/EnvMap.cpp:48-63)
            x, y, z = [math.radians(a) for a in deg]
            rx = np.array([[1, 0, 0], [0, math.cos(x), -math.sin(x)], [0, math.sin(x), math.cos(x)]])
            ry = np.array([[math.cos(y), 0, math.sin(y)], [0, 1, 0], [-math.sin(y), 0, math.cos(y)]])
            rz = np.array([[math.cos(z), -math.sin(z), 0], [math.sin(z), math.cos(z), 0], [0, 0, 1]])
            m = rz @ ry @ rx          # column-vector convention
            return m.T, m             # row-vector convention: dir * M^T ; inverse = M (rotation)

        t, ti = rot(self.envMapRotation)
        pt, pti = rot(prevRotation if prevRotation is not None else self.envMapRotation)
        d.transform[:] = t.astype(np.float32).ravel()
        d.invTransform[:] = ti.astype(np.float32).ravel()
        d.prevTransform[:] = pt.astype(np.float32).ravel()
        d.prevInvTransform[:] = pti.astype(np.float32).ravel()
        return d

    # --- lights --------------------------------------------------------------------------------------------------
    def addDirectionalLight(self, direction, intensity):
        d = np.asarray(direction, dtype=np.float32)
        d = d / np.float32(np.sqrt(np.sum(d * d, dtype=np.float32)))
        self.lights.append(dict(type=1, posW=(0, 0, 0), dirW=tuple(float(x) for x in d), intensity=tuple(intensity)))

    def addPointLight(self, position, intensity):
        self.lights.append(dict(type=0, posW=tuple(position), dirW=(0, -1, 0), intensity=tuple(intensity)))

    def lights_array(self):
        arr = (capi.Light * max(1, len(self.lights)))()
        for i, l in enumerate(self.lights):
            arr[i].type = l["type"]
            arr[i].posW[:] = [float(x) for x in l["posW"]]
            arr[i].dirW[:] = [float(x) for x in l["dirW"]]
            arr[i].intensity[:] = [float(x) for x in l["intensity"]]
        return arr, len(self.lights)

    def addEmissiveShell(self, count, center, radius, seed=4):
        arr = (capi.EmissiveTriangle * count)()
        capi.check(capi.lib().vrestir_make_emissive_shell(int(count), int(seed), _f3(center), float(radius), arr))
        self.emissiveTriangles = arr
        return arr

    def volume_bounds_world(self):
        """World-space AABB of density mip 0 (F/Scene/Scene.cpp:3311-3314)."""
        g = self.volume.grid.contents.slots[0]
        m = np.array(list(g.medium_to_world), dtype=np.float64).reshape(4, 4)
        lo = np.array([g.bmin[0], g.bmin[1], g.bmin[2], 1.0]) @ m
        hi = np.array([g.bmax[0], g.bmax[1], g.bmax[2], 1.0]) @ m
        return lo[:3], hi[:3]

    def frame_camera(self, distance_in_diagonals=2.5, direction=(0.3, 0.25, 1.0), fovY=45.0):
        """Pinhole camera at `distance_in_diagonals` x bbox diagonal looking at the centre."""
        lo, hi = self.volume_bounds_world()
        c = 0.5 * (lo + hi)
        diag = float(np.linalg.norm(hi - lo))
        d = np.asarray(direction, dtype=np.float64)
        d = d / np.linalg.norm(d)
        self.camera.position = tuple(c + d * diag * distance_in_diagonals)
        self.camera.target = tuple(c)
        self.camera.up = (0.0, 1.0, 0.0)
        self.camera.fovY = fovY
