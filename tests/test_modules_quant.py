"""Module surgery + checkpoint behaviour of the quant API, in the style of
/root/reference/tests/test_modules_quant.py (same model shapes and assertions; NF4 only)."""
import pytest
import torch
import torch.nn as nn

from src.modules.quant import (
    BnbLinear4bit,
    quantize_inplace,
    quantize_state_dict,
    replace_by_prequantized_weights,
    replace_to_quant_linear,
    validate_quant_type,
)
from src.modules.quant.functional import collect_children_dict, get_quant_type_from_children_dict


def test_collect_children_keys():
    cases = [
        ("abc.def.", {"abc.def.0": 0, "abc.def.1": 0, "abc.def.2": 0, "abc.ooo": 0}, ["0", "1", "2"]),
        ("abc.def.", {"abc.def.ghi.jkl": 0}, ["ghi.jkl"]),
        ("abc.def.", {"XYZ": 0}, []),
    ]
    for prefix, sd, expected in cases:
        assert set(collect_children_dict(prefix, sd).keys()) == set(expected)


def test_get_quant_type_from_children_keys():
    z = torch.zeros(1)
    cases = [
        ({"absmax": z, "quant_state.bitsandbytes__nf4": z}, "bnb_nf4"),
        ({"quant_map": z, "quant_state.bitsandbytes__fp4": z}, "bnb_fp4"),
        ({"weight_format": z}, "bnb_int8"),
        ({"_data": torch.zeros(1, dtype=torch.int8)}, "quanto_int8"),
        ({"_data._data": torch.zeros(1, dtype=torch.uint8)}, "quanto_int4"),
    ]
    for keys, expected in cases:
        assert get_quant_type_from_children_dict(keys) == expected
    with pytest.raises(ValueError):
        get_quant_type_from_children_dict({"foo": z})


def test_validate_quant_type():
    for q in ["fp8_e4m3fn", "bnb_int8", "bnb_fp4", "bnb_nf4", "quanto_int4", "quanto_int8", "ao_nf4", "ao_fp8"]:
        validate_quant_type(q)
    with pytest.raises(ValueError):
        validate_quant_type("bnb_nf8")


class _Model(nn.Module):
    def __init__(self, dtype=None):
        super().__init__()
        self.linear = nn.Linear(128, 256, bias=True, dtype=dtype)
        self.non_quant = nn.Linear(128, 256, bias=True, dtype=dtype)


@torch.no_grad()
def test_replace_to_quant_linear():
    model = replace_to_quant_linear(_Model(), quant_type="bnb_nf4", include_keys=["linear"], exclude_keys=[])
    assert isinstance(model.linear, BnbLinear4bit) and isinstance(model.linear, nn.Linear)
    assert not isinstance(model.non_quant, BnbLinear4bit)
    assert model.linear.weight.device.type == "meta" and model.linear.quant_type == "nf4"
    assert (model.linear.in_features, model.linear.out_features) == (128, 256)
    assert not any(p.requires_grad for p in model.linear.parameters())


@torch.no_grad()
def test_quantize_inplace_keeps_weights_until_cuda():
    model = _Model()
    w = model.linear.weight.detach().clone()
    quantize_inplace(model, quant_type="bnb_nf4", include_keys=["linear"], exclude_keys=[])
    assert isinstance(model.linear, BnbLinear4bit)
    assert torch.equal(model.linear.weight.data, w) and not model.linear.weight.bnb_quantized
    assert model.linear.compress_statistics is True and model.linear.quant_storage == torch.uint8


@torch.no_grad()
def test_other_backends_raise_not_implemented():
    for q in ("bnb_fp4", "bnb_int8", "ao_nf4", "ao_fp8", "quanto_int8"):
        with pytest.raises(NotImplementedError):
            replace_to_quant_linear(_Model(), quant_type=q, include_keys=["linear"])
    with pytest.raises(NotImplementedError):
        quantize_state_dict({}, "bnb_int8", ["x"])
    with pytest.raises(ValueError):
        replace_to_quant_linear(_Model(), quant_type="nope", include_keys=["linear"])


@torch.no_grad()
def test_prequantized_checkpoint_loads_on_cpu():
    """bnb-format keys -> BnbLinear4bit with packed uint8 weight + quant state; state_dict round-trips the keys."""
    from oracle import nf4_oracle  # fixture generator only

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(16, 32, bias=True)
            self.non_quant = nn.Linear(16, 32, bias=True)

    src = M()
    sd = src.state_dict()
    p, a = nf4_oracle.nf4_quantize(sd["linear.weight"])
    sd["linear.weight"] = torch.from_numpy(p)
    sd["linear.weight.absmax"] = torch.from_numpy(a)
    sd["linear.weight.quant_map"] = torch.from_numpy(nf4_oracle.NF4_CODEBOOK.copy())
    sd["linear.weight.quant_state.bitsandbytes__nf4"] = torch.from_numpy(nf4_oracle.pack_quant_state_blob((32, 16), "float32"))

    model = M()
    replace_by_prequantized_weights(model, sd)
    assert isinstance(model.linear, BnbLinear4bit) and not isinstance(model.non_quant, BnbLinear4bit)
    model.load_state_dict(sd)
    assert model.linear.weight.dtype == torch.uint8 and model.linear.weight.bnb_quantized
    assert tuple(model.linear.weight.quant_state.shape) == (32, 16)
    out = model.state_dict()
    for k in sd:
        assert k in out, k
    assert torch.equal(out["linear.weight"], sd["linear.weight"])
    assert torch.equal(out["linear.weight.absmax"], sd["linear.weight.absmax"])
    assert bytes(out["linear.weight.quant_state.bitsandbytes__nf4"].tolist()) == bytes(
        sd["linear.weight.quant_state.bitsandbytes__nf4"].tolist()
    )


def _nested_stats(n_out=128, n_in=256, seed=0):
    """A compress_statistics=True checkpoint entry built by the oracle (tools/quantize_model.py output format)."""
    import json

    import numpy as np

    from oracle import nf4_oracle

    w = torch.randn(n_out, n_in, generator=torch.Generator().manual_seed(seed)).to(torch.bfloat16)
    p, a = nf4_oracle.nf4_quantize(w)
    q8, a2, off, code2 = nf4_oracle.absmax_nest(a)
    meta = {"quant_type": "nf4", "blocksize": 64, "dtype": "bfloat16", "shape": [n_out, n_in],
            "nested_blocksize": 256, "nested_dtype": "float32", "nested_offset": float(off)}
    stats = {
        "absmax": torch.from_numpy(q8),
        "quant_map": torch.from_numpy(nf4_oracle.NF4_CODEBOOK.copy()),
        "nested_absmax": torch.from_numpy(a2),
        "nested_quant_map": torch.from_numpy(np.array(code2)),
        "quant_state.bitsandbytes__nf4": torch.tensor(list(json.dumps(meta).encode()), dtype=torch.uint8),
    }
    return w, torch.from_numpy(p), stats, nf4_oracle.absmax_denest(q8, a2, off, code2)


@torch.no_grad()
def test_nested_absmax_checkpoint_stays_nested_on_cpu():
    """compress_statistics=True checkpoints: absmax uint8 + nested_* keys survive load -> state_dict unchanged (the
    fp32 statistics are derived on the device, never stored)."""
    from vft_b200.nn import Params4bit

    _, packed, stats, _ = _nested_stats()
    w = Params4bit.from_prequantized(packed, stats, device="cpu")
    qs = w.quant_state
    assert qs.nested and qs.absmax.dtype == torch.uint8 and qs.state2.blocksize == 256
    out = qs.as_dict(packed=True)
    assert set(out) == set(stats)
    for k in ("absmax", "nested_absmax", "nested_quant_map", "quant_map"):
        assert torch.equal(out[k], stats[k]), k
    import json

    blob = lambda t: json.loads(bytes(t.tolist()).decode())
    assert blob(out["quant_state.bitsandbytes__nf4"]) == blob(stats["quant_state.bitsandbytes__nf4"])
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        qs.absmax_f32()


@pytest.mark.gpu
@torch.no_grad()
def test_nested_absmax_checkpoint_forward_matches_oracle():
    """Prequantized nested checkpoint -> module -> forward with the de-nested statistics (bit-exact vs the oracle)."""
    from oracle import qlora_oracle

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(256, 128, bias=False)

    w, packed, stats, absmax_eff = _nested_stats()
    sd = {"linear.weight": packed, **{f"linear.weight.{k}": v for k, v in stats.items()}}
    model = M()
    replace_by_prequantized_weights(model, sd)
    model.load_state_dict(sd)
    model.cuda()
    qs = model.linear.weight.quant_state
    assert qs.nested and torch.equal(qs.absmax_f32().cpu(), torch.from_numpy(absmax_eff))
    out = model.state_dict()
    for k, v in sd.items():
        assert torch.equal(out[k].cpu(), v), k
    x = torch.randn(70, 256, dtype=torch.bfloat16)
    y = model.linear(x.cuda())
    ref = qlora_oracle.qlora_linear_ref(x, qlora_oracle.dequant_weight(packed.numpy(), absmax_eff, (128, 256)), None, None, None, 1.0)["y"]
    assert qlora_oracle.rel_l2(y.cpu(), ref) < 6e-3


# ----------------------------------------------------------------------------- GPU: the reference's numeric tests
@pytest.mark.gpu
@torch.no_grad()
def test_bnb_load_prequantized():
    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(16, 32, bias=True)
            self.non_quant = nn.Linear(16, 32, bias=True)

    model = M()
    state_dict = quantize_state_dict(model.state_dict(), quant_type="bnb_nf4", include_keys=["linear.weight"], exclude_keys=[])
    assert state_dict["linear.weight"].dtype == torch.uint8 and state_dict["linear.weight"].shape == (256, 1)
    assert "linear.weight.quant_state.bitsandbytes__nf4" in state_dict and "non_quant.weight.absmax" not in state_dict
    model = M()
    replace_by_prequantized_weights(model, state_dict)
    model.load_state_dict(state_dict)
    assert isinstance(model.linear, BnbLinear4bit)


@pytest.mark.gpu
@torch.no_grad()
def test_bnb_quantize_inplace_and_load():
    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(16, 32, bias=True, dtype=torch.float16)
            self.non_quant = nn.Linear(32, 32, bias=True, dtype=torch.float16)

        def forward(self, x):
            return self.non_quant(self.linear(x))

    model = M()
    quantize_inplace(model, quant_type="bnb_nf4", include_keys=["linear"], exclude_keys=[])
    model.cuda()  # do quantization
    state_dict = model.state_dict()
    assert state_dict["linear.weight"].dtype == torch.uint8

    inputs = torch.randn(1, 16, dtype=torch.float16).to("cuda")
    output = model(inputs)
    assert output.dtype == torch.float16

    del model
    model = M()
    assert "linear.weight.quant_state.bitsandbytes__nf4" in state_dict.keys()
    replace_by_prequantized_weights(model, state_dict)
    model.load_state_dict(state_dict)
    model.cuda()
    assert model.linear.weight.dtype == torch.uint8

    output_2 = model(inputs)
    assert torch.allclose(output, output_2)


@pytest.mark.gpu
@torch.no_grad()
def test_quantized_module_matches_oracle_and_moves_between_devices():
    from oracle import nf4_oracle, qlora_oracle

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(128, 256, bias=True, dtype=torch.bfloat16)

    model = M()
    w, b = model.linear.weight.detach().clone(), model.linear.bias.detach().clone()
    quantize_inplace(model, "bnb_nf4", include_keys=["linear"])
    model.cuda()
    p, a = nf4_oracle.nf4_quantize(w)
    assert torch.equal(model.linear.weight.data.cpu(), torch.from_numpy(p))
    # module default compress_statistics=True (bnb.py:44): nested statistics, each piece bit-exact vs the oracle
    qs = model.linear.weight.quant_state
    # the module encodes around torch's fp32 absmax.mean() on the device (what bitsandbytes does); it must sit within
    # a few ulps of the correctly rounded mean, and every other piece follows from it bit for bit
    exact_mean = float(a.astype("float64").mean())
    assert abs(float(qs.offset) - exact_mean) <= 4 * 2.0 ** -24 * exact_mean
    q8, a2, off, code2 = nf4_oracle.absmax_nest(a, offset=float(qs.offset))
    assert qs.nested and torch.equal(qs.absmax.cpu(), torch.from_numpy(q8))
    assert torch.equal(qs.state2.absmax.cpu(), torch.from_numpy(a2)) and float(qs.offset) == float(off)
    assert torch.equal(qs.state2.code.cpu(), torch.from_numpy(code2))
    a = nf4_oracle.absmax_denest(q8, a2, off, code2)
    assert torch.equal(qs.absmax_f32().cpu(), torch.from_numpy(a))
    sd = model.state_dict()
    assert sd["linear.weight.absmax"].dtype == torch.uint8 and "linear.weight.nested_absmax" in sd and "linear.weight.nested_quant_map" in sd
    x = torch.randn(3, 50, 128, dtype=torch.bfloat16)
    y = model.linear(x.cuda())
    ref = qlora_oracle.qlora_linear_ref(x, qlora_oracle.dequant_weight(p, a, (256, 128)), b, None, None, 1.0)["y"]
    assert y.shape == (3, 50, 256) and qlora_oracle.rel_l2(y.cpu(), ref) < 6e-3
    # cuda -> cpu -> cuda keeps the packed weight and its state
    model.cpu()
    assert model.linear.weight.dtype == torch.uint8 and model.linear.weight.device.type == "cpu"
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.linear(x)
    model.cuda()
    assert torch.equal(model.linear(x.cuda()), y)


@pytest.mark.gpu
@torch.no_grad()
def test_deepcopy_owns_its_quant_state_and_inplace_updates_invalidate_derived_buffers():
    """(a) a deep-copied module (EMA / reference copy) must not share its QuantState with the original --
    moving the copy must leave the original on its device; (b) packed bytes / statistics overwritten IN PLACE (same
    storage, same pointers) must rebuild the micro-tiled copy and the decoded statistics: the tcgen05 path (tiled copy)
    and the few-token path (checkpoint layout) have to agree."""
    import copy

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(256, 512, bias=False, dtype=torch.bfloat16)

    torch.manual_seed(7)
    model = M()
    quantize_inplace(model, "bnb_nf4", include_keys=["linear"])
    model.cuda()
    x_many = torch.randn(96, 256, dtype=torch.bfloat16, device="cuda")   # tcgen05 path
    x_few = x_many[:2].contiguous()                                       # few-token streaming path
    y0 = model.linear(x_many)

    # (a)
    clone = copy.deepcopy(model)
    assert clone.linear.weight.quant_state is not model.linear.weight.quant_state
    assert clone.linear.weight.module is clone.linear and clone.linear.quant_state is clone.linear.weight.quant_state
    assert clone.linear.weight.quant_state.absmax.data_ptr() != model.linear.weight.quant_state.absmax.data_ptr()
    assert torch.equal(clone.linear(x_many), y0)
    clone.cpu()
    assert model.linear.weight.quant_state.absmax.is_cuda and model.linear.weight.quant_state.state2.absmax.is_cuda
    assert torch.equal(model.linear(x_many), y0)

    # (b) a second weight quantized into the SAME storage
    other = M()
    quantize_inplace(other, "bnb_nf4", include_keys=["linear"])
    other.cuda()
    y_other_many, y_other_few = other.linear(x_many), other.linear(x_few)
    assert not torch.equal(y_other_many, y0)
    w, qs = model.linear.weight, model.linear.weight.quant_state
    ptrs = (w.data_ptr(), qs.absmax.data_ptr(), qs.state2.absmax.data_ptr())
    w.data.copy_(other.linear.weight.data)
    qs.absmax.copy_(other.linear.weight.quant_state.absmax)
    qs.state2.absmax.copy_(other.linear.weight.quant_state.state2.absmax)
    qs.offset = other.linear.weight.quant_state.offset.clone()
    assert ptrs == (w.data_ptr(), qs.absmax.data_ptr(), qs.state2.absmax.data_ptr())
    assert torch.equal(model.linear(x_many), y_other_many)
    assert torch.equal(model.linear(x_few), y_other_few)
