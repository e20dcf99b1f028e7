import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import ops
w = [(torch.randn(18432, 3072, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(3)]
for t in w: ops.nf4_quantize(t)
torch.cuda.synchronize(); print("ok")
