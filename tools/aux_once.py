"""Launch the adapter side kernels once (ncu target).  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops
T, N, K, r = 4096, 3072, 3072, 16
dev = torch.device("cuda"); bf = torch.bfloat16
w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
packed, absmax = ops.nf4_quantize(w)
x = torch.randn(T, K, device=dev, dtype=bf); g = torch.randn(T, N, device=dev, dtype=bf)
A = (torch.randn(r, K, device=dev) * 0.02).to(bf); B = (torch.randn(N, r, device=dev) * 0.02).to(bf)
y = torch.empty(T, N, device=dev, dtype=bf); dx = torch.empty(T, K, device=dev, dtype=bf)
ts = torch.empty(T, 64, device=dev, dtype=bf); dts = torch.empty(T, 64, device=dev, dtype=bf)
dA = torch.empty_like(A); dB = torch.empty_like(B)
wsb = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD_DAB, T, N, K, r)
ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
st = torch.cuda.current_stream().cuda_stream
L = _cabi.lib
for _ in range(3):
    _cabi.check(L.vft_qlora_fwd(x.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, A.data_ptr(), B.data_ptr(), r, 1.0 / r, y.data_ptr(), ts.data_ptr(), None, None, None, 0, None, None, st))
    _cabi.check(L.vft_qlora_bwd_dx(g.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, A.data_ptr(), B.data_ptr(), r, 1.0 / r, None, dts.data_ptr(), None, None, 0, None, None, st))
    _cabi.check(L.vft_lora_bwd_dab(g.data_ptr(), x.data_ptr(), ts.data_ptr(), dts.data_ptr(), T, N, K, r, 2, 1.0 / r, dA.data_ptr(), dB.data_ptr(), ws.data_ptr(), wsb, st))
torch.cuda.synchronize()
print("ok")
