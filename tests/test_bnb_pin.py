"""Pin against real bitsandbytes output (tests/golden/bnb_vectors.safetensors, written by tests/golden/make_bnb_golden.py
through tools/pin_bnb.sh).  The package cannot be installed in the offline image (profiles/r02_bnb_pin_attempt.log), so
the fixture does not exist yet and these tests SKIP: the NF4 / nested-statistics half of the oracle stays "parity
unpinned"  until someone runs the script where bitsandbytes is available."""
import os

import numpy as np
import pytest
import torch
from safetensors.torch import load_file

from oracle import nf4_oracle

FIXTURE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "bnb_vectors.safetensors")
pytestmark = pytest.mark.skipif(not os.path.exists(FIXTURE), reason="no bitsandbytes fixture (package not installable offline)")


def _cases():
    v = load_file(FIXTURE)
    return v, sorted({k.rsplit(".", 2)[0] for k in v if k.endswith(".plain.packed")})


def test_oracle_matches_bitsandbytes():
    v, names = _cases()
    g = load_file(os.path.join(os.path.dirname(FIXTURE), "nf4_vectors.safetensors"))
    for name in names:
        if name.startswith("seeded3072_"):
            dt = {"bfloat16": torch.bfloat16, "float16": torch.float16}[name.split("_", 1)[1]]
            w = (torch.randn(3072, 3072, generator=torch.Generator().manual_seed(0)) * 0.02).to(dt)
        else:
            w = g["probe_f32"] if name == "probe" else g[f"{name}_w"]
        p, a = nf4_oracle.nf4_quantize(w)
        assert np.array_equal(p, v[f"{name}.plain.packed"].numpy()), name
        assert np.array_equal(a, v[f"{name}.plain.absmax"].numpy()), name
        off = float(v[f"{name}.nested.offset"][0])
        q8, a2, _, code = nf4_oracle.absmax_nest(a, offset=off)
        assert np.array_equal(code, v[f"{name}.nested.nested_code"].numpy()), name
        assert np.array_equal(q8, v[f"{name}.nested.absmax"].numpy()), name
        assert np.array_equal(a2, v[f"{name}.nested.nested_absmax"].numpy()), name


@pytest.mark.gpu
def test_cuda_matches_bitsandbytes():
    from vft_b200 import nn as vnn

    v, names = _cases()
    g = load_file(os.path.join(os.path.dirname(FIXTURE), "nf4_vectors.safetensors"))
    for name in names:
        if name.startswith("seeded3072_"):
            continue
        w = g["probe_f32"] if name == "probe" else g[f"{name}_w"]
        packed, qs = vnn.quantize_4bit(w.cuda(), compress_statistics=True)
        assert torch.equal(packed.cpu(), v[f"{name}.nested.packed"]), name
        assert torch.equal(qs.absmax.cpu(), v[f"{name}.nested.absmax"]), name
        assert float(qs.offset) == float(v[f"{name}.nested.offset"][0]), name
        if f"{name}.nested.dequant" in v:
            assert torch.equal(vnn.dequantize_4bit(packed, qs).cpu().reshape(-1), v[f"{name}.nested.dequant"].reshape(-1)), name
