"""Host-side cost of one fused-layer call through the module API (kernels are tiny here, so the loop is host-bound).
Not part of the product."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch, torch.nn as nn
from src.modules.peft import LoRAConfig, PeftTargetConfig
from src.modules.quant import quantize_inplace

class M(nn.Module):
    def __init__(self):
        super().__init__()
        self.linear = nn.Linear(256, 256, bias=False, dtype=torch.bfloat16)
        self.plain = nn.Linear(256, 256, bias=False, dtype=torch.bfloat16)
m = M(); quantize_inplace(m, "bnb_nf4", include_keys=["linear", "plain"]); m.cuda()
PeftTargetConfig(config=LoRAConfig(rank=16, alpha=1.0, dtype="bfloat16"), include_keys=["linear"]).replace_to_peft_layer(m, freeze_base=True)
x = torch.randn(64, 256, device="cuda", dtype=torch.bfloat16, requires_grad=True)
ref = nn.Linear(256, 256, bias=False, dtype=torch.bfloat16, device="cuda")
def bench(fn, n=2000):
    for _ in range(50): fn()
    torch.cuda.synchronize(); t = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t) / n * 1e6
with torch.no_grad():
    print(f"forward, no grad : NF4 {bench(lambda: m.plain(x)):6.1f} us   NF4+LoRA {bench(lambda: m.linear(x)):6.1f} us   torch nn.Linear {bench(lambda: ref(x)):6.1f} us")
def fb(layer):
    def f():
        x.grad = None
        layer(x).sum().backward()
    return f
print(f"fwd+bwd          : NF4 {bench(fb(m.plain), 1000):6.1f} us   NF4+LoRA {bench(fb(m.linear), 1000):6.1f} us   torch nn.Linear {bench(fb(ref), 1000):6.1f} us")
