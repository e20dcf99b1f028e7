"""A full QLoRA training step through the module API vs the CPU oracle (BASELINE north_star: "a full training
step's loss must match").

The model is a small transformer-style block stack built from the pieces the reference trains with
(/root/reference/src/trainer/common.py:169-198: quantize -> replace_to_peft_layer -> optimizer over the adapter
parameters; :287-365: forward, loss, backward, step), at the layer shapes of the census (square attention-style
projections, an up/down MLP pair with bias, one NF4-only layer without adapter).  The oracle model is the same graph in
plain PyTorch on the CPU with the weights dequantized by oracle/nf4_oracle.py and the adapter composed exactly as
/root/reference/src/modules/peft/lora.py:92-104 does.
"""
import copy

import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from oracle import nf4_oracle, qlora_oracle
from src.modules.peft import LoRAConfig, PeftTargetConfig, get_adapter_parameters
from src.modules.quant import quantize_inplace

H, FF, T, R = 256, 640, 384, 16


class Block(nn.Module):
    def __init__(self):
        super().__init__()
        self.q = nn.Linear(H, H, bias=False)
        self.o = nn.Linear(H, H, bias=False)
        self.up = nn.Linear(H, FF, bias=True)
        self.down = nn.Linear(FF, H, bias=True)
        self.mod = nn.Linear(H, H, bias=True)  # NF4 only, no adapter (like the modulation layers)

    def forward(self, x):
        x = x + self.o(F.gelu(self.q(x)))
        x = x + self.down(F.silu(self.up(x)))
        return x * (1 + 0.1 * torch.tanh(self.mod(x)))


class Net(nn.Module):
    def __init__(self, n=2):
        super().__init__()
        self.blocks = nn.ModuleList([Block() for _ in range(n)])

    def forward(self, x):
        for b in self.blocks:
            x = b(x)
        return x


class OracleLoRALinear(nn.Module):
    """Dequantized base (frozen) + adapter, op for op as the reference composes them (bf16 roundings per op)."""

    def __init__(self, w_deq, bias, a, b, scale):
        super().__init__()
        self.w = w_deq
        self.bias = bias
        self.a = None if a is None else nn.Parameter(a.clone())
        self.b = None if b is None else nn.Parameter(b.clone())
        self.scale = scale

    def forward(self, x):
        y = F.linear(x, self.w, self.bias)
        if self.a is None:
            return y
        return y + F.linear(F.linear(x, self.a), self.b) * self.scale


@pytest.mark.gpu
def test_one_training_step_matches_the_oracle():
    torch.manual_seed(0)
    ref_fp = Net().to(torch.bfloat16)
    with torch.no_grad():
        for p in ref_fp.parameters():
            if p.dim() == 2:
                p.normal_(std=0.05)
            else:
                p.normal_(std=0.02)
    model = copy.deepcopy(ref_fp)

    # ---- product path: NF4 base + LoRA through the mirrored reference API, on the GPU
    quantize_inplace(model, "bnb_nf4", include_keys=["blocks"])
    model.to("cuda")
    PeftTargetConfig(config=LoRAConfig(rank=R, alpha=8.0, dtype="bfloat16"), include_keys=[".q", ".o", ".up", ".down"]
                     ).replace_to_peft_layer(model, freeze_base=True)
    gen = torch.Generator().manual_seed(1)
    adapters = {}
    for name, mod in model.named_modules():
        if hasattr(mod, "lora_up"):
            a = ((torch.rand(R, mod.lora_down.weight.shape[1], generator=gen) * 2 - 1) * 0.1).to(torch.bfloat16)
            b = (torch.randn(mod.lora_up.weight.shape[0], R, generator=gen) * 0.05).to(torch.bfloat16)
            with torch.no_grad():
                mod.lora_down.weight.copy_(a)
                mod.lora_up.weight.copy_(b)
            adapters[name] = (a, b)
    assert len([k for k in get_adapter_parameters(model) if "lora_" in k]) == 2 * 4 * 2
    params = [p for n, p in model.named_parameters() if p.requires_grad and "lora_" in n]
    assert len(params) == 2 * 4 * 2
    opt = torch.optim.SGD(params, lr=0.5)

    # ---- oracle model on the CPU from the same packed weights
    oracle = copy.deepcopy(ref_fp)
    o_params = []
    for name, mod in list(model.named_modules()):
        base = getattr(mod, "linear", None) if hasattr(mod, "lora_up") else (mod if type(mod).__name__ == "BnbLinear4bit" else None)
        if base is None or (not hasattr(mod, "lora_up") and name.endswith(".linear")):
            continue
        packed = base.weight.data.cpu().numpy()
        absmax = nf4_oracle.quant_state_absmax_f32(base.weight.quant_state)
        n_out, n_in = base.out_features, base.in_features
        w_deq = qlora_oracle.dequant_weight(packed, absmax, (n_out, n_in), "bfloat16")
        bias = None if base.bias is None else base.bias.detach().cpu()
        a, b = adapters.get(name, (None, None))
        om = OracleLoRALinear(w_deq, bias, a, b, 8.0 / R)
        parent = oracle
        *path, leaf = name.split(".")
        for part in path:
            parent = parent[int(part)] if part.isdigit() else getattr(parent, part)
        setattr(parent, leaf, om)
        if a is not None:
            o_params += [om.a, om.b]
    o_opt = torch.optim.SGD(o_params, lr=0.5)

    x = torch.randn(2, T // 2, H, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    target = torch.randn(2, T // 2, H, generator=torch.Generator().manual_seed(3)).to(torch.bfloat16)

    losses, o_losses = [], []
    for _ in range(2):  # two steps: the second one sees the updated adapters
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model(x.cuda()).float(), target.cuda().float())
        loss.backward()
        opt.step()
        losses.append(float(loss.detach()))
        o_opt.zero_grad(set_to_none=True)
        o_loss = F.mse_loss(oracle(x).float(), target.float())
        o_loss.backward()
        o_opt.step()
        o_losses.append(float(o_loss))

    for l, ol in zip(losses, o_losses):
        assert abs(l - ol) <= 5e-3 * abs(ol), (losses, o_losses)
    assert o_losses[1] < o_losses[0] and losses[1] < losses[0]
    # updated adapter weights after two steps
    got = {n: (m.lora_down.weight.detach().cpu(), m.lora_up.weight.detach().cpu()) for n, m in model.named_modules() if hasattr(m, "lora_up")}
    for name, (ga, gb) in got.items():
        parent = oracle
        for part in name.split("."):
            parent = parent[int(part)] if part.isdigit() else getattr(parent, part)
        assert qlora_oracle.rel_l2(ga, parent.a.detach()) < 2e-2, name
        assert qlora_oracle.rel_l2(gb, parent.b.detach()) < 2e-2, name
