"""volumetricrestirrelease_b200 — B200-native VolumetricReSTIR render-pass hot path.

csrc/      hand-written sm_100a CUDA kernels + the extern "C" boundary (include/vrestir.h) -> libvrestir.so
scene.py   host-side scene mirror (addGVDBVolume / setEnvMap / camera / lights)
render_pass.py  host-side mirror of the reference pass interface (VolumetricReSTIR, VolumetricReSTIRParams)
post.py    AccumulatePass / ErrorMeasurePass, the passes behind accumulated_color in the reference's render graphs
multi_gpu.py  row-band sharding with reservoir-halo exchange (one process per GPU)
"""
from . import _capi as capi
from .render_pass import VolumetricReSTIR, VolumetricReSTIRParams
from .post import AccumulatePass, ErrorMeasurePass
from .scene import Scene

__all__ = ["capi", "VolumetricReSTIR", "VolumetricReSTIRParams", "Scene", "AccumulatePass", "ErrorMeasurePass"]
