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
    return lib


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
