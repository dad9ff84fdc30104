"""The C-ABI library loads on a CPU-only box and exports every symbol include/vrestir.h declares (no compute calls)."""
import ctypes as C
import os
import re

from volumetricrestirrelease_b200 import capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "vrestir.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vrestir_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported():
    lib = capi.lib()
    names = declared_symbols()
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == names


def test_struct_layouts_match_header_sizes():
    # sizes the C side relies on (32-byte node = one L2 sector; 32-byte host reservoir record)
    assert C.sizeof(capi.Node) == 32
    assert C.sizeof(capi.Reservoir) == 32
    assert C.sizeof(capi.Params) == 4 * len(capi.PARAM_FIELDS)
    assert C.sizeof(capi.EmissiveTriangle) == 64


def test_default_params_match_python_mirror():
    p = capi.Params()
    capi.lib().vrestir_default_params(C.byref(p))
    for name, _, default in capi.PARAM_FIELDS:
        assert abs(getattr(p, name) - default) < 1e-6, name


def test_create_fails_loudly_without_gpu_or_works_with_one():
    import torch
    h = C.c_void_p()
    p = capi.Params()
    capi.lib().vrestir_default_params(C.byref(p))
    rc = capi.lib().vrestir_create(C.byref(p), 0, C.byref(h))
    if torch.cuda.is_available():
        assert rc == 0
        capi.lib().vrestir_destroy(h)
    else:
        assert rc == capi.ERR_CUDA          # no CPU fallback
        assert b"no CPU fallback" in capi.lib().vrestir_last_error()
