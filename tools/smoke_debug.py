import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch, torch.nn as nn
from oracle import nf4_oracle, qlora_oracle
from src.modules.peft import LoRAConfig, PeftTargetConfig
from src.modules.quant import quantize_inplace
torch.manual_seed(0)
K, N, T, r = 256, 384, 320, 16
class M(nn.Module):
    def __init__(self):
        super().__init__(); self.linear = nn.Linear(K, N, bias=False, dtype=torch.bfloat16)
model = M(); w = model.linear.weight.detach().clone()
quantize_inplace(model, "bnb_nf4", include_keys=["linear"]); model.to("cuda:0")
p, a = nf4_oracle.nf4_quantize(w)
PeftTargetConfig(config=LoRAConfig(rank=r, alpha=1.0, dtype="bfloat16"), include_keys=["linear"]).replace_to_peft_layer(model, freeze_base=True)
layer = model.linear
with torch.no_grad(): layer.lora_up.weight.normal_(std=0.02)
x = torch.randn(T, K, dtype=torch.bfloat16); dy = torch.randn(T, N, dtype=torch.bfloat16)
xg = x.to("cuda:0").requires_grad_(True)
y = layer(xg); y.backward(dy.to("cuda:0")); torch.cuda.synchronize()
ref = qlora_oracle.qlora_linear_ref(x, qlora_oracle.dequant_weight(p, a, (N, K)), None, layer.lora_down.weight.detach().cpu(), layer.lora_up.weight.detach().cpu(), 1.0, dy)
ga, gb = layer.lora_down.weight.grad.cpu(), layer.lora_up.weight.grad.cpu()
print("dA norms got/ref", float(ga.float().norm()), float(ref["da"].float().norm()), "max diff", float((ga.float()-ref["da"].float()).abs().max()))
print("dB norms got/ref", float(gb.float().norm()), float(ref["db"].float().norm()), "max diff", float((gb.float()-ref["db"].float()).abs().max()))
print("n differing elements dA", int((ga != ref["da"]).sum()), "of", ga.numel(), " dB", int((gb != ref["db"]).sum()), "of", gb.numel())
