"""CPU: the C-ABI library loads, exports every symbol include/vft_b200.h declares, and validates arguments
without touching a GPU."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def cabi():
    from vft_b200 import build

    build.build()  # no-op when up to date
    from vft_b200 import _cabi

    return _cabi


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "vft_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vft_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(cabi):
    declared = _declared_symbols()
    assert len(declared) >= 11
    raw = ctypes.CDLL(cabi.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/vft_b200.h but not exported"
    assert sorted(cabi.SYMBOLS) == declared, "ctypes binding and header disagree"


def test_every_exported_symbol_is_declared(cabi):
    """The reverse direction: nothing named vft_* leaves the library without a declaration in the header."""
    import subprocess

    lib_path = os.path.join(os.path.dirname(cabi.__file__), "libvft_b200.so")
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = sorted({ln.split()[-1] for ln in out.splitlines() if ln.split() and ln.split()[-1].startswith("vft_")})
    assert exported == _declared_symbols()


def test_abi_version(cabi):
    assert cabi.lib.vft_abi_version() == cabi.ABI_VERSION == 7


def test_argument_validation_needs_no_gpu(cabi):
    lib = cabi.lib
    # null pointers -> VFT_ERR_INVALID with a message
    assert lib.vft_nf4_quantize(None, cabi.BF16, 64, 64, None, None, None) == -1
    assert b"null" in lib.vft_last_error()
    assert lib.vft_nf4_quantize(None, cabi.BF16, -1, 64, None, None, None) == -1
    # n == 0 is a no-op
    assert lib.vft_nf4_quantize(None, cabi.BF16, 0, 64, None, None, None) == 0
    # LoRA rank out of range / missing adapter pointers
    rc = lib.vft_qlora_fwd(1, 4, 1, 1, 64, 64, 64, cabi.BF16, cabi.BF16, None, None, None, 65, 1.0, 1, 1, None, None, None, 0, None, None, None)
    assert rc == -1 and b"rank" in lib.vft_last_error()
    rc = lib.vft_qlora_fwd(1, 4, 1, 1, 64, 64, 64, cabi.BF16, cabi.BF16, None, None, None, 16, 1.0, 1, 1, None, None, None, 0, None, None, None)
    assert rc == -1
    # the micro-tiled copy comes as a pair
    rc = lib.vft_qlora_fwd(1, 4, 1, 1, 64, 64, 64, cabi.BF16, cabi.BF16, None, None, None, 0, 1.0, 1, None, None, None, None, 0, 1, None, None)
    assert rc == -1 and b"together" in lib.vft_last_error()
    assert lib.vft_nf4_tiled_bytes(100, 128, 0) == 128 * 128 // 2 and lib.vft_nf4_tiled_bytes(100, 128, 1) == 128 * 2 * 4
    assert lib.vft_nf4_tiled_bytes(100, 100, 0) == 0
    # workspace contract
    assert lib.vft_workspace_bytes(cabi.OP_BWD_DAB, 4096, 3072, 3072, 16) == 4 * (3072 + 3072) * 16
    rc = lib.vft_lora_bwd_dab(1, 1, 1, 1, 8, 64, 64, 16, cabi.BF16, 1.0, 1, 1, None, 0, None)
    assert rc == -4 and b"workspace" in lib.vft_last_error()


def test_nested_statistics_argument_validation(cabi):
    lib = cabi.lib
    assert lib.vft_absmax_nest(None, 256, 256, None, None, None, None, None, 0, None) == -1 and b"null" in lib.vft_last_error()
    # only bitsandbytes' nested blocksize is implemented; the workspace contract is enforced before any launch
    assert lib.vft_absmax_nest(16, 256, 128, 16, 16, 16, 16, None, 0, None) == -1 and b"256" in lib.vft_last_error()
    need = lib.vft_workspace_bytes(cabi.OP_ABSMAX_NEST, 0, 0, 0, 0)
    assert need == 128 * 8
    assert lib.vft_absmax_nest(16, 256, 256, 16, 16, 16, 16, None, 0, None) == -4 and b"workspace" in lib.vft_last_error()
    assert lib.vft_absmax_denest(None, None, None, 0.0, 0, 256, None, None) == 0  # empty vector: no-op
    assert lib.vft_absmax_denest(None, None, None, 0.0, 5, 256, None, None) == -1


def test_cpu_tensors_are_rejected(cabi):
    import torch
    from vft_b200 import ops

    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.nf4_quantize(torch.zeros(64))
