"""GPU mip / conservative-mip chain builder (csrc/vr_mipbuild.cu behind ``vrestir_mips_*``).

``build_mips(dense)`` takes a (Z, Y, X) float32 CUDA tensor and returns, per level, the stored form of the normal and the
conservative chain as torch tensors that alias the builder's device memory (fp32 for mip 0 of the normal chain, UNORM8
codes + scale for everything else) — what the brick pool of each ``.vbx`` level holds."""
import ctypes as C

import torch

from . import _capi as capi
from .multi_gpu import device_view


class MipChain:
    def __init__(self, handle, device):
        self._h = handle
        self.device = device
        n = C.c_int(0)
        capi.check(capi.lib().vrestir_mips_count(self._h, C.byref(n)))
        self.num_mips = n.value

    def level(self, mip, conservative=False):
        """(tensor (Z, Y, X) float32 | uint8, max_value).  Stored value of a UNORM8 level = code / 255 * max_value."""
        lv = capi.MipLevel()
        capi.check(capi.lib().vrestir_mips_level(self._h, int(mip), int(conservative), C.byref(lv)))
        t = device_view(lv.data, lv.bytes, self.device)
        shape = (lv.dim[2], lv.dim[1], lv.dim[0])
        t = t.view(torch.float32).view(shape) if lv.format == 0 else t.view(shape)
        return t, float(lv.max_value)

    def close(self):
        if self._h:
            capi.lib().vrestir_mips_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def build_mips(dense, num_mips=4, stream=None):
    if not (dense.is_cuda and dense.dtype == torch.float32 and dense.dim() == 3 and dense.is_contiguous()):
        raise ValueError("dense must be a contiguous (Z, Y, X) float32 CUDA tensor")
    dim = (C.c_int32 * 3)(dense.shape[2], dense.shape[1], dense.shape[0])
    h = C.c_void_p()
    capi.check(capi.lib().vrestir_mips_build_device(dense.device.index or 0, C.c_void_p(dense.data_ptr()), C.byref(dim), int(num_mips),
                                                    C.byref(h), C.c_void_p(stream) if stream else None))
    return MipChain(h, dense.device)
