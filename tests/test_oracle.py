"""CPU: the oracle against the committed golden vectors and against its own C restatement."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch
from safetensors.torch import load_file

from oracle import c_oracle, nf4_oracle, qlora_oracle


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def nf4_vec(golden_dir):
    return load_file(os.path.join(golden_dir, "nf4_vectors.safetensors"))


@pytest.fixture(scope="module")
def lora_vec(golden_dir):
    return load_file(os.path.join(golden_dir, "lora_vectors.safetensors"))


def test_codebook_thresholds_are_midpoints():
    """Thresholds = bitsandbytes' f-suffixed literals (decimal -> fp32 directly): the fp32 neighbour of the
    float64 midpoint of adjacent code-book entries, at most 1 ulp from float32(midpoint)."""
    import ctypes

    cb = nf4_oracle.NF4_CODEBOOK.astype(np.float64)
    mid64 = (cb[:-1] + cb[1:]) / 2
    thr = nf4_oracle.NF4_THRESHOLDS
    assert thr.dtype == np.float32 and (np.diff(thr) > 0).all()
    ulp = np.spacing(np.abs(thr))
    assert (np.abs(thr.astype(np.float64) - mid64) <= ulp).all()
    libc = ctypes.CDLL(None)
    libc.strtof.restype = ctypes.c_float
    libc.strtof.argtypes = [ctypes.c_char_p, ctypes.c_void_p]
    for t, m in zip(thr, mid64):
        assert np.float32(libc.strtof(repr(float(m)).encode(), None)) == t  # the literal bnb compiles


def test_probe_vectors(nf4_vec):
    probe = nf4_vec["probe_f32"]
    packed, absmax = nf4_oracle.nf4_quantize(probe)
    assert np.array_equal(packed, nf4_vec["probe_packed"].numpy())
    assert np.array_equal(absmax, nf4_vec["probe_absmax"].numpy())
    pc, ac = c_oracle.quantize(probe)  # gcc parses the f-literals: must agree on every +-1 ulp probe
    assert np.array_equal(pc, packed) and np.array_equal(ac, absmax)
    # semantic anchors: a code-book value encodes to its own index, a threshold to the index below
    codes = nf4_oracle.nf4_unpack(packed, probe.numel())
    p = probe.numpy()
    for j, v in enumerate(nf4_oracle.NF4_CODEBOOK):
        assert (codes[p == v] == j).all()
    for j, t in enumerate(nf4_oracle.NF4_THRESHOLDS):
        assert (codes[p == t] == j).all()
        assert (codes[p == np.nextafter(t, np.float32(2))] == j + 1).all()


@pytest.mark.parametrize("name", ["tail1", "tail63", "tail64", "tail65", "tail127", "odd_rows", "k16", "zero_block"])
def test_tail_vectors_numpy_and_c(nf4_vec, name):
    w = nf4_vec[f"{name}_w"]
    exp_p, exp_a = nf4_vec[f"{name}_packed"].numpy(), nf4_vec[f"{name}_absmax"].numpy()
    p, a = nf4_oracle.nf4_quantize(w)
    assert np.array_equal(p, exp_p) and np.array_equal(a, exp_a)
    pc, ac = c_oracle.quantize(w)
    assert np.array_equal(pc, exp_p) and np.array_equal(ac, exp_a)
    assert p.shape == ((w.numel() + 1) // 2, 1)
    dt = str(w.dtype).replace("torch.", "")
    d_np = nf4_oracle.nf4_dequantize(p, a, w.shape, dt)
    d_c = c_oracle.dequantize(p, a, w.numel(), dt).reshape(w.shape)
    assert torch.equal(d_np, d_c)


def test_zero_block_encodes_to_code_zero(nf4_vec):
    p = nf4_vec["zero_block_packed"].numpy().reshape(-1)
    assert (p[32:64] == 0).all()
    assert nf4_vec["zero_block_absmax"].numpy()[1] == 0.0


@pytest.mark.parametrize("dt_name,dt", [("bfloat16", torch.bfloat16), ("float16", torch.float16)])
def test_seeded_3072_hashes(golden_dir, dt_name, dt):
    with open(os.path.join(golden_dir, "nf4_hashes.json")) as f:
        ref = json.load(f)[dt_name]
    g = torch.Generator().manual_seed(ref["seed"])
    w = (torch.randn(*ref["shape"], generator=g) * ref["std"]).to(dt)
    p, a = c_oracle.quantize(w)
    assert _sha(p) == ref["packed_sha256"] and _sha(a) == ref["absmax_sha256"]
    p2, a2 = nf4_oracle.nf4_quantize(w)
    assert _sha(p2) == ref["packed_sha256"] and _sha(a2) == ref["absmax_sha256"]


def test_quantize_roundtrip_error_bound():
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(256, 192, generator=g) * 0.02).to(torch.bfloat16)
    p, a = nf4_oracle.nf4_quantize(w)
    d = nf4_oracle.nf4_dequantize(p, a, w.shape, "bfloat16").float()
    # largest code-book gap is 0.3038 -> error <= half of it (+ bf16 rounding) times the block absmax
    bound = torch.from_numpy(a).repeat_interleave(64)[: w.numel()].reshape(w.shape) * (0.3038 / 2 + 0.01)
    assert ((d - w.float()).abs() <= bound).all()
    # idempotence: re-quantizing the dequantized tensor reproduces the codes
    p2, a2 = nf4_oracle.nf4_quantize(d.to(torch.bfloat16))
    assert np.array_equal(p, p2)


@pytest.mark.parametrize("name", ["r16", "r4_bias"])
def test_lora_oracle_pinned_to_reference(lora_vec, name):
    """qlora_linear_ref must reproduce, bit for bit, what the REFERENCE's LoRALinear computed on CPU
    (fixtures frozen by tests/golden/make_golden.py from /root/reference/src/modules/peft/lora.py)."""
    v = lora_vec
    N, K = v[f"{name}_w"].shape
    w_deq = qlora_oracle.dequant_weight(v[f"{name}_packed"].numpy(), v[f"{name}_absmax"].numpy(), (N, K), "bfloat16")
    out = qlora_oracle.qlora_linear_ref(
        v[f"{name}_x"], w_deq, v.get(f"{name}_bias"), v[f"{name}_a"], v[f"{name}_b"], float(v[f"{name}_alpha"]), v[f"{name}_dy"]
    )
    assert torch.equal(out["y"], v[f"{name}_y"])
    assert torch.equal(out["dx"], v[f"{name}_dx"])
    assert torch.equal(out["da"], v[f"{name}_da"])
    assert torch.equal(out["db"], v[f"{name}_db"])
    # and the fp64 truth stays within bf16 rounding of it
    truth = qlora_oracle.qlora_linear_truth(
        v[f"{name}_x"], w_deq, v.get(f"{name}_bias"), v[f"{name}_a"], v[f"{name}_b"], float(v[f"{name}_alpha"]), v[f"{name}_dy"]
    )
    for k in ("y", "dx", "da", "db"):
        assert qlora_oracle.rel_l2(out[k], truth[k]) < 1e-2


def test_quant_state_blob_roundtrip():
    blob = nf4_oracle.pack_quant_state_blob((32, 16), "float16")
    meta = nf4_oracle.unpack_quant_state_blob(blob)
    assert meta == {"quant_type": "nf4", "blocksize": 64, "dtype": "float16", "shape": [32, 16]}


# ----------------------------------------------------------------------------- nested ("double quant") statistics
def test_dynamic_map_shape_and_hash(golden_dir):
    m = nf4_oracle.dynamic_map()
    assert m.dtype == np.float32 and m.shape == (256,) and (np.diff(m) > 0).all()
    assert m[0] == np.float32(-0.99296874) and m[127] == 0.0 and m[255] == 1.0  # 127 negatives, 0, 128 positives
    assert np.isclose(m[128], 5.5e-7) and np.isclose(m[1], -0.9789063)
    with open(os.path.join(golden_dir, "nf4_hashes.json")) as f:
        assert _sha(m) == json.load(f)["dynamic_map_sha256"]


def test_dquantize8_is_nearest_entry():
    m = nf4_oracle.dynamic_map()
    rng = np.random.default_rng(0)
    x = np.concatenate([rng.uniform(-1, 1, 4000), rng.normal(0, 1e-4, 2000), m, [-1.0, 1.0, 0.0]]).astype(np.float32)
    got = nf4_oracle.dquantize8(m, x)
    scalar = np.array([nf4_oracle._dquantize8(m, v) for v in x[:1500]])
    assert np.array_equal(got[:1500], scalar)
    d = np.abs(x[:, None].astype(np.float64) - m[None, :].astype(np.float64))
    assert np.array_equal(d[np.arange(x.size), got], d.min(axis=1))  # a nearest entry (ties: either side)
    assert np.array_equal(nf4_oracle.dquantize8(m, m), np.arange(256))  # every code value maps to itself


@pytest.mark.parametrize("n", [1, 5, 255, 256, 257, 1000, 3072 * 3072 // 64])
def test_absmax_nest_numpy_and_c(n):
    rng = np.random.default_rng(n)
    a = (np.abs(rng.normal(0, 0.02, n)) * 3 + 0.05).astype(np.float32)
    q, a2, off, code = nf4_oracle.absmax_nest(a)
    qc, a2c, offc = c_oracle.absmax_nest(a, code)
    assert np.array_equal(q, qc) and np.array_equal(a2, a2c) and off == offc
    d = nf4_oracle.absmax_denest(q, a2, off, code)
    assert np.array_equal(d, c_oracle.absmax_denest(q, a2, off, code))
    if n > 1:  # the 8-bit map resolves (absmax - mean) / absmax2 to ~1.5 % of full scale near +-1
        assert np.abs(d - a).max() <= 0.016 * np.abs(a - off).max() + 1e-9
    # idempotence of the decoded statistics under re-encoding with the same offset/scale grid
    assert q.dtype == np.uint8 and a2.shape == ((n + 255) // 256,)


@pytest.mark.parametrize("dt_name,dt", [("bfloat16", torch.bfloat16), ("float16", torch.float16)])
def test_nested_seeded_3072_hashes(golden_dir, dt_name, dt):
    with open(os.path.join(golden_dir, "nf4_hashes.json")) as f:
        ref = json.load(f)[dt_name]
    g = torch.Generator().manual_seed(ref["seed"])
    w = (torch.randn(*ref["shape"], generator=g) * ref["std"]).to(dt)
    _, a = c_oracle.quantize(w)
    q, a2, off, code = nf4_oracle.absmax_nest(a)
    assert _sha(q) == ref["nested_absmax8_sha256"] and _sha(a2) == ref["nested_absmax2_sha256"]
    assert float(off) == ref["nested_offset"]
    assert _sha(nf4_oracle.absmax_denest(q, a2, off, code)) == ref["denested_absmax_sha256"]


def test_product_dynamic_map_matches_oracle():
    """vft_b200.nn.create_dynamic_map (the product's host-side table, what lands in ``nested_quant_map``) against the
    oracle's restatement of bitsandbytes.functional.create_dynamic_map -- two independent writings of the same table."""
    from vft_b200.nn import create_dynamic_map

    m = create_dynamic_map()
    assert m.dtype == torch.float32 and np.array_equal(m.numpy(), nf4_oracle.dynamic_map())
    m[0] = 123.0  # callers get a private copy of the cached table
    assert create_dynamic_map()[0] == np.float32(-0.99296874)
