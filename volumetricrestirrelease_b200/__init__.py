"""volumetricrestirrelease_b200 — B200-native VolumetricReSTIR render-pass hot path.

csrc/      hand-written sm_100a CUDA kernels + the extern "C" boundary (include/vrestir.h) -> libvrestir.so
scene.py   host-side scene mirror (addGVDBVolume / setEnvMap / camera / lights)
render_pass.py  host-side mirror of the reference pass interface (VolumetricReSTIR, VolumetricReSTIRParams)
"""
from . import _capi as capi
from .render_pass import VolumetricReSTIR, VolumetricReSTIRParams
from .scene import Scene

__all__ = ["capi", "VolumetricReSTIR", "VolumetricReSTIRParams", "Scene"]
