"""LoRA surgery in the style of /root/reference/tests/test_peft.py:26-126,234-292 (Linear adapters)."""
import pytest
import torch
import torch.nn as nn

from src.modules.peft import LoRAConfig, LoRALinear, PeftTargetConfig, get_adapter_parameters, load_peft_weight
from src.modules.peft import while_peft_disabled
from src.modules.peft.functional import detect_peft_method, extract_peft_layers
from src.utils.state_dict import RegexMatch


class ChildLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.child_1 = nn.Linear(10, 10)  # <- target
        self.child_extra = nn.Linear(10, 10)

    def forward(self, x):
        return self.child_extra(self.child_1(x))


class TestModel(nn.Module):
    __test__ = False

    def __init__(self):
        super().__init__()
        self.layer1 = nn.Sequential(nn.Linear(10, 10), nn.ReLU(), nn.Linear(10, 10))  # .0 <- target
        self.layer2 = nn.Sequential(nn.Linear(10, 10), nn.ReLU(), nn.Linear(10, 10))
        self.child = ChildLayer()
        self.last_layer = nn.ModuleList([nn.Linear(10, 20)])  # <- target

    def forward(self, x):
        return self.last_layer[0](self.layer2(self.layer1(x)))


def _config(rank=4, dtype="float16", **kw):
    return PeftTargetConfig(
        config=LoRAConfig(type="lora", dtype=dtype, rank=rank, alpha=1.0, dropout=0.0, use_bias=False),
        include_keys=[".0", RegexMatch(regex=r".*\.child_\d+")],
        exclude_keys=["layer2"],
        **kw,
    )


@torch.no_grad()
def test_replace_lora_linear():
    model = TestModel().to(torch.float16)
    inputs = torch.randn(1, 10, dtype=torch.float16)
    original_output = model(inputs)

    _config().replace_to_peft_layer(model, freeze_base=True)

    assert isinstance(model.layer1[0], LoRALinear)
    assert model.layer1[0].lora_down.weight.T.shape == torch.Size([10, 4])
    assert model.layer1[0].lora_up.weight.T.shape == torch.Size([4, 10])
    assert isinstance(model.layer1[2], nn.Linear) and not isinstance(model.layer1[2], LoRALinear)
    assert isinstance(model.layer2[0], nn.Linear) and isinstance(model.layer2[2], nn.Linear)
    assert isinstance(model.child.child_1, LoRALinear)
    assert isinstance(model.child.child_extra, nn.Linear)
    assert isinstance(model.last_layer[0], LoRALinear)
    assert model.last_layer[0].lora_up.weight.T.shape == torch.Size([4, 20])

    # initial LoRA output is zero (B = 0) -> identical outputs
    assert torch.equal(original_output, model(inputs))

    for name, param in model.named_parameters():
        assert param.requires_grad is ("lora_" in name), name

    adapter_params = get_adapter_parameters(model)
    assert sorted(adapter_params.keys()) == sorted(
        f"{m}.{k}" for m in ("layer1.0", "child.child_1", "last_layer.0") for k in ("lora_down.weight", "lora_up.weight", "alpha")
    )


def test_include_keys_must_not_be_empty():
    with pytest.raises(ValueError):
        PeftTargetConfig(config=LoRAConfig(rank=4), include_keys=[])


@torch.no_grad()
def test_second_replace_skips_existing_adapters_and_use_bias():
    model = TestModel()
    _config(dtype="float32").replace_to_peft_layer(model)
    first = model.layer1[0]
    PeftTargetConfig(config=LoRAConfig(rank=8, use_bias=True, dtype="float32"), include_keys=["layer"]).replace_to_peft_layer(model)
    assert model.layer1[0] is first and first.rank == 4          # untouched
    assert isinstance(model.layer2[0], LoRALinear) and model.layer2[0].rank == 8
    assert model.layer2[0].lora_up.bias is not None
    assert "layer2.0.lora_up.bias" in get_adapter_parameters(model)


@torch.no_grad()
def test_enable_switch_and_train_mode():
    model = TestModel()
    _config(dtype="float32").replace_to_peft_layer(model, freeze_base=True)
    x = torch.randn(2, 10)
    for layer in extract_peft_layers(model).values():
        layer.lora_up.weight.normal_()
    with_lora = model(x)
    with while_peft_disabled(model):
        without = model(x)
    assert not torch.equal(with_lora, without)
    assert torch.equal(with_lora, model(x))
    model.train()
    assert model.layer1[0].lora_down.training and not model.layer1[0].linear.training


@torch.no_grad()
def test_adapter_state_dict_roundtrip():
    src = TestModel()
    _config(dtype="float32").replace_to_peft_layer(src)
    for layer in extract_peft_layers(src).values():
        layer.lora_up.weight.normal_()
    sd = get_adapter_parameters(src)
    assert detect_peft_method(sd) == "lora"
    dst = TestModel()
    dst.load_state_dict({k: v for k, v in src.state_dict().items() if "lora_" not in k and ".alpha" not in k and ".linear." not in k}
                        | {k.replace(".linear.", "."): v for k, v in src.state_dict().items() if ".linear." in k})
    load_peft_weight(dst, sd)  # wraps the plain Linears on the fly
    assert isinstance(dst.layer1[0], LoRALinear) and dst.layer1[0].rank == 4
    x = torch.randn(3, 10)
    assert torch.allclose(src(x), dst(x))
    with pytest.raises(ValueError):
        load_peft_weight(TestModel(), {"foo": torch.zeros(1)})


# ----------------------------------------------------------------------------- GPU: LoRA over the NF4 layer
@pytest.mark.gpu
def test_lora_over_nf4_layer_fused_path():
    from oracle import nf4_oracle, qlora_oracle
    from src.modules.quant import quantize_inplace
    from vft_b200 import ops

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(256, 384, bias=False, dtype=torch.bfloat16)

        def forward(self, x):
            return self.linear(x)

    torch.manual_seed(0)
    model = M()
    w = model.linear.weight.detach().clone()
    quantize_inplace(model, "bnb_nf4", include_keys=["linear"])
    model.cuda()
    x = torch.randn(2, 70, 256, dtype=torch.bfloat16)
    base_out = model(x.cuda())
    PeftTargetConfig(config=LoRAConfig(rank=16, alpha=1.0, dtype="bfloat16"), include_keys=["linear"]).replace_to_peft_layer(
        model, freeze_base=True
    )
    layer = model.linear
    assert isinstance(layer, LoRALinear) and layer.lora_down.weight.is_cuda
    # B = 0 at init: output equals the base output exactly (tests/test_peft.py:98-101 of the reference)
    assert torch.equal(model(x.cuda()), base_out)

    with torch.no_grad():
        layer.lora_up.weight.normal_(std=0.02)
    xg = x.cuda().requires_grad_(True)
    y = model(xg)
    assert ops.last_path() == 1  # tcgen05
    dy = torch.randn_like(y)
    y.backward(dy)
    p, a = nf4_oracle.nf4_quantize(w)
    ref = qlora_oracle.qlora_linear_ref(x, qlora_oracle.dequant_weight(p, a, (384, 256)), None,
                                        layer.lora_down.weight.detach().cpu(), layer.lora_up.weight.detach().cpu(), 1.0, dy.cpu())
    assert qlora_oracle.rel_l2(y.detach().cpu(), ref["y"]) < 6e-3
    assert qlora_oracle.rel_l2(xg.grad.cpu(), ref["dx"]) < 6e-3
    assert qlora_oracle.rel_l2(layer.lora_down.weight.grad.cpu(), ref["da"]) < 2e-2
    assert qlora_oracle.rel_l2(layer.lora_up.weight.grad.cpu(), ref["db"]) < 2e-2
    assert layer.linear.weight.grad is None
    with while_peft_disabled(model):
        assert torch.equal(model(x.cuda()), base_out)


@pytest.mark.gpu
def test_gradient_checkpointing_reenters_the_function():
    from src.modules.quant import quantize_inplace
    from torch.utils.checkpoint import checkpoint

    class Block(nn.Module):
        def __init__(self):
            super().__init__()
            self.fc1 = nn.Linear(128, 256, bias=False, dtype=torch.bfloat16)
            self.fc2 = nn.Linear(256, 128, bias=False, dtype=torch.bfloat16)

        def forward(self, x):
            return self.fc2(torch.nn.functional.gelu(self.fc1(x)))

    torch.manual_seed(1)
    model = Block()
    quantize_inplace(model, "bnb_nf4", include_keys=["fc"])
    model.cuda()
    PeftTargetConfig(config=LoRAConfig(rank=8, dtype="bfloat16"), include_keys=["fc"]).replace_to_peft_layer(model, freeze_base=True)
    with torch.no_grad():
        for m in (model.fc1, model.fc2):
            m.lora_up.weight.normal_(std=0.02)
    x = torch.randn(64, 128, dtype=torch.bfloat16, device="cuda", requires_grad=True)
    model(x).float().pow(2).mean().backward()
    g_plain = [p.grad.clone() for p in model.parameters() if p.requires_grad] + [x.grad.clone()]
    for p in model.parameters():
        p.grad = None
    x.grad = None
    checkpoint(model, x, use_reentrant=False).float().pow(2).mean().backward()
    g_ckpt = [p.grad for p in model.parameters() if p.requires_grad] + [x.grad]
    assert torch.equal(g_plain[-1], g_ckpt[-1])  # dX: deterministic kernels
    for a, b in zip(g_plain[:-1], g_ckpt[:-1]):  # dA/dB: fp32 atomics -> summation order varies
        assert float((a.float() - b.float()).norm() / b.float().norm().clamp_min(1e-20)) < 5e-3


@pytest.mark.gpu
def test_use_bias_adapter_takes_the_fused_path():
    """LoRAConfig(use_bias=True): lora_up carries a bias (/root/reference/src/modules/peft/lora.py:52-60).  s * bias rides
    the fused epilogue's per-feature scalar; y, dx and the gradients of A, B and the bias against the reference's own
    composition base + lora_up(lora_down(x)) * (alpha / rank) in fp64."""
    from src.modules.quant import quantize_inplace
    from vft_b200 import ops

    class M(nn.Module):
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(256, 384, bias=True, dtype=torch.bfloat16)

        def forward(self, x):
            return self.linear(x)

    torch.manual_seed(1)
    model = M()
    quantize_inplace(model, "bnb_nf4", include_keys=["linear"])
    model.cuda()
    PeftTargetConfig(config=LoRAConfig(rank=8, alpha=4.0, use_bias=True, dtype="bfloat16"), include_keys=["linear"]).replace_to_peft_layer(
        model, freeze_base=True)
    layer = model.linear
    assert layer.lora_up.bias is not None
    with torch.no_grad():
        layer.lora_up.weight.normal_(std=0.05)
        layer.lora_up.bias.normal_(std=0.5)
    x = torch.randn(3, 50, 256, dtype=torch.bfloat16, device="cuda", requires_grad=True)
    assert layer._can_fuse(x)
    y = model(x)
    assert ops.last_path() == 1  # tcgen05
    dy = torch.randn_like(y)
    y.backward(dy)
    # fp64 truth of the reference's composition on the dequantized weight
    wd = layer.linear.dequantized_weight().double() if hasattr(layer.linear, "dequantized_weight") else None
    if wd is None:
        from vft_b200.nn import dequantize_4bit
        wd = dequantize_4bit(layer.linear.weight.data, layer.linear.weight.quant_state).double()
    xd = x.detach().double().requires_grad_(True)
    A = layer.lora_down.weight.detach().double().requires_grad_(True)
    B = layer.lora_up.weight.detach().double().requires_grad_(True)
    b = layer.lora_up.bias.detach().double().requires_grad_(True)
    s = 4.0 / 8
    yt = xd @ wd.t() + layer.linear.bias.detach().double() + ((xd @ A.t()) @ B.t() + b) * s
    yt.backward(dy.double())
    rel = lambda a_, b_: float((a_.double() - b_).norm() / b_.norm())
    assert rel(y.detach(), yt.detach()) < 4e-3
    assert rel(x.grad, xd.grad) < 4e-3
    assert rel(layer.lora_down.weight.grad, A.grad) < 6e-3
    assert rel(layer.lora_up.weight.grad, B.grad) < 6e-3
    assert rel(layer.lora_up.bias.grad, b.grad) < 6e-3
    assert layer.linear.bias.grad is None and layer.linear.weight.grad is None
