"""ncu launch stub: one vft_nf4_quantize_many call over 96 x [3072, 3072] fp16 tensors (one grouped launch)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import ops
ws = [(torch.randn(3072, 3072, device="cuda") * 0.02).to(torch.float16) for _ in range(96)]
for _ in range(2): ops.nf4_quantize_many(ws)
torch.cuda.synchronize(); print("ok")
