"""ctypes binding of libvft_b200.so (declared in include/vft_b200.h).

The library is the product: there is no Python/PyTorch/CPU fallback behind these
calls.  If the shared object is missing the import of this module fails loudly
(``VftLibraryError``) instead of degrading.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# VFT_LIB: A/B measurements against another build of the same library (tools/*_probe.py); symbols that build lacks
# are then skipped and its ABI version is not checked.  Never set in the product path.
_ALT_LIB = os.environ.get("VFT_LIB")
LIB_PATH = _ALT_LIB if _ALT_LIB else os.path.join(_HERE, "libvft_b200.so")

ABI_VERSION = 7
LORA_LD = 64
F32, F16, BF16 = 0, 1, 2
PATH_NONE, PATH_TCGEN05, PATH_SIMT, PATH_GEMV = 0, 1, 2, 3
OP_FWD, OP_BWD_DX, OP_BWD_DAB, OP_ABSMAX_NEST, OP_BWD = 0, 1, 2, 3, 4

# every symbol include/vft_b200.h declares: (restype, argtypes)
_c = ctypes
_p, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float
SYMBOLS = {
    "vft_abi_version": (_i, []),
    "vft_last_error": (_c.c_char_p, []),
    "vft_last_path": (_i, []),
    "vft_force_path": (None, [_i]),
    "vft_reload_env": (None, []),
    "vft_debug_tc_timeline": (_i, [_p, _i]),
    "vft_debug_tc2_timeline": (_i, [_p, _i]),
    "vft_debug_side_timeline": (_i, [_p, _i]),
    "vft_debug_tc2_p0dump": (_i, [_p, _i]),
    "vft_nf4_quantize": (_i, [_p, _i, _i64, _i, _p, _p, _p]),
    "vft_nf4_quantize_many": (_i, [_i, _p, _i, _p, _i, _p, _p, _p]),
    "vft_nf4_dequantize": (_i, [_p, _p, _i64, _i, _p, _i, _p]),
    "vft_nf4_quantize_host": (_i, [_p, _i, _i64, _i, _p, _p]),
    "vft_nf4_tiled_bytes": (_i64, [_i64, _i64, _i]),
    "vft_nf4_tile_weight": (_i, [_p, _p, _i64, _i64, _i, _p, _p, _p]),
    "vft_absmax_nest": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _i64, _p]),
    "vft_absmax_nest_at": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p]),
    "vft_absmax_denest": (_i, [_p, _p, _p, _f, _i64, _i, _p, _p]),
    "vft_workspace_bytes": (_i64, [_i, _i64, _i64, _i64, _i]),
    "vft_qlora_fwd": (_i, [_p, _i64, _p, _p, _i64, _i64, _i, _i, _i, _p, _p, _p, _i, _f, _p, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "vft_qlora_bwd": (_i, [_p, _p, _i64, _p, _p, _i64, _i64, _i, _i, _i, _p, _p, _i, _f, _p, _p, _p, _p, _p, _p, _p, _p, _i64,
                           _p, _p, _p]),
    "vft_qlora_bwd_dx": (_i, [_p, _i64, _p, _p, _i64, _i64, _i, _i, _i, _p, _p, _i, _f, _p, _p, _p, _p, _i64, _p, _p, _p]),
    "vft_lora_bwd_dab": (_i, [_p, _p, _p, _p, _i64, _i64, _i64, _i, _i, _f, _p, _p, _p, _i64, _p]),
}


class VftLibraryError(RuntimeError):
    pass


class VftError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"vft_b200 error {status}: {message}")
        self.status = status


def _load() -> ctypes.CDLL:
    if not os.path.exists(LIB_PATH):
        raise VftLibraryError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py build` "
            "(nvcc, sm_100a). There is no fallback for the QLoRA hot path."
        )
    lib = ctypes.CDLL(LIB_PATH)
    for name, (restype, argtypes) in SYMBOLS.items():
        if _ALT_LIB and not hasattr(lib, name):
            continue
        fn = getattr(lib, name)  # AttributeError if a declared symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if not _ALT_LIB and lib.vft_abi_version() != ABI_VERSION:
        raise VftLibraryError(f"ABI mismatch: library {lib.vft_abi_version()} != binding {ABI_VERSION}")
    return lib


lib = _load()


def check(status: int) -> None:
    if status != 0:
        raise VftError(status, (lib.vft_last_error() or b"").decode("utf-8", "replace"))
