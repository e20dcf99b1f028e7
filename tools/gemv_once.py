"""One few-token forward launch per weight copy (for ncu captures).  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import ops
N, K, T = (int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (18432, 3072, 2)))
w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
p0, a0 = ops.nf4_quantize(w)
packs = [(p0.clone(), a0.clone()) for _ in range(6)]
x = torch.randn(T, K, device="cuda", dtype=torch.bfloat16)
for p, a in packs:
    y = ops.qlora_linear(x, p, a, None, None, None, 0.0, N, K, 64, torch.bfloat16)
torch.cuda.synchronize()
print("ok", ops.last_path())
