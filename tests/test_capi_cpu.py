"""CPU tests: the C-ABI library loads, exports every symbol include/ffvd_b200.h declares, and the
product path fails loudly (no CPU fallback) when no B200 is visible."""
import ctypes
import os
import re

import numpy as np
import pytest

import ffvd_b200
from ffvd_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ffvd_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ffvd_[a-z_0-9A-Z]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = ffvd_b200.load_library()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(lib, s), "libffvd_b200.so does not export %s" % s
    assert lib.ffvd_version() >= 100
    assert lib.ffvd_status_string(0) == b"ok"
    assert b"positive definite" in lib.ffvd_status_string(7)


def test_struct_layout_matches_header():
    # 11 pointers each, in the header's order
    assert ctypes.sizeof(_capi._Problem) == 11 * ctypes.sizeof(ctypes.c_void_p)
    assert ctypes.sizeof(_capi._Outputs) == 11 * ctypes.sizeof(ctypes.c_void_p)
    assert [f[0] for f in _capi._Problem._fields_] == ["X", "Z", "U", "logv", "logl", "logQ", "C", "d", "logR", "Y", "ctrl"]


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(ffvd_b200.FFVDError):
        ffvd_b200.Context(0)
    from ffvd_b200.kernels_multi_output import SquaredExponential
    k = SquaredExponential(3, variance=0.5, lengthscales=np.ones(3), ARD=True)
    with pytest.raises(Exception):
        k.K(np.zeros((4, 3)))                     # must not silently compute on the CPU


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ffvd_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert "import oracle" not in src and "from oracle" not in src, fn


def test_dlpack_capsule_roundtrip_pointer():
    a = np.arange(6, dtype=np.float64).reshape(2, 3)
    b = _capi._Borrow()
    p = b.ptr(a)
    assert p
    class DLDevice(ctypes.Structure):
        _fields_ = [("device_type", ctypes.c_int32), ("device_id", ctypes.c_int32)]
    class DLTensor(ctypes.Structure):
        _fields_ = [("data", ctypes.c_void_p), ("device", DLDevice), ("ndim", ctypes.c_int32), ("code", ctypes.c_uint8),
                    ("bits", ctypes.c_uint8), ("lanes", ctypes.c_uint16), ("shape", ctypes.POINTER(ctypes.c_int64)),
                    ("strides", ctypes.POINTER(ctypes.c_int64)), ("byte_offset", ctypes.c_uint64)]
    t = ctypes.cast(p, ctypes.POINTER(DLTensor)).contents
    assert t.device.device_type == 1 and t.ndim == 2 and t.bits == 64 and t.code == 2
    assert t.shape[0] == 2 and t.shape[1] == 3 and t.data == a.ctypes.data
    b.release()
