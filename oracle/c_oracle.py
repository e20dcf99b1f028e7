"""ctypes loader for the plain-C NF4 oracle (oracle/nf4_ref.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libnf4_oracle.so")
DTYPE_CODE = {"float32": 0, "float16": 1, "bfloat16": 2}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "nf4_ref.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE], check=True, capture_output=True)
    return _SO


def _lib():
    lib = ctypes.CDLL(build())
    lib.nf4_quantize_ref.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    lib.nf4_quantize_ref.restype = ctypes.c_int
    lib.nf4_dequantize_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    lib.nf4_dequantize_ref.restype = ctypes.c_int
    lib.absmax_nest_ref.argtypes = [ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.absmax_nest_ref.restype = ctypes.c_int
    lib.absmax_denest_ref.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_float, ctypes.c_int64, ctypes.c_int, ctypes.c_void_p]
    lib.absmax_denest_ref.restype = ctypes.c_int
    return lib


def absmax_nest(absmax, code, blocksize2: int = 256):
    """fp32 absmax -> (absmax8, absmax2, offset) through the C restatement."""
    a = np.ascontiguousarray(absmax, dtype=np.float32).reshape(-1)
    c = np.ascontiguousarray(code, dtype=np.float32)
    n = a.size
    q = np.zeros(n, np.uint8)
    a2 = np.zeros((n + blocksize2 - 1) // blocksize2, np.float32)
    off = np.zeros(1, np.float32)
    rc = _lib().absmax_nest_ref(a.ctypes.data, n, blocksize2, c.ctypes.data, q.ctypes.data, a2.ctypes.data, off.ctypes.data)
    assert rc == 0
    return q, a2, off[0]


def absmax_denest(q, absmax2, offset, code, blocksize2: int = 256):
    q = np.ascontiguousarray(q, dtype=np.uint8).reshape(-1)
    a2 = np.ascontiguousarray(absmax2, dtype=np.float32)
    c = np.ascontiguousarray(code, dtype=np.float32)
    out = np.zeros(q.size, np.float32)
    rc = _lib().absmax_denest_ref(q.ctypes.data, a2.ctypes.data, c.ctypes.data, float(offset), q.size, blocksize2, out.ctypes.data)
    assert rc == 0
    return out


def quantize(w_torch, blocksize: int = 64):
    """w_torch: CPU torch tensor (fp32/fp16/bf16). Returns (packed uint8 [ceil(n/2),1], absmax f32)."""
    import torch

    w = w_torch.detach().contiguous().cpu()
    n = w.numel()
    packed = np.zeros(((n + 1) // 2, 1), np.uint8)
    absmax = np.zeros(((n + blocksize - 1) // blocksize,), np.float32)
    code = DTYPE_CODE[str(w.dtype).replace("torch.", "")]
    rc = _lib().nf4_quantize_ref(w.data_ptr(), code, n, blocksize, packed.ctypes.data, absmax.ctypes.data)
    assert rc == 0
    return packed, absmax


def dequantize(packed, absmax, n: int, dtype: str = "bfloat16", blocksize: int = 64):
    import torch

    tdt = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}[dtype]
    out = torch.empty(n, dtype=tdt)
    p = np.ascontiguousarray(packed, dtype=np.uint8)
    a = np.ascontiguousarray(absmax, dtype=np.float32)
    rc = _lib().nf4_dequantize_ref(p.ctypes.data, a.ctypes.data, n, blocksize, out.data_ptr(), DTYPE_CODE[dtype])
    assert rc == 0
    return out
