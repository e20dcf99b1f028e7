"""Launch the fused kernels a few times for one shape (ncu target).  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

T, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 3072, 3072))]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda")
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
packed, absmax = ops.nf4_quantize(w)
tiles = None if os.environ.get('VFT_NOTILE') else ops.nf4_tile_weight(packed, absmax, N, K)
TC, TA = (tiles[0].data_ptr(), tiles[1].data_ptr()) if tiles else (None, None)
x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
g = torch.randn(T, N, device=dev, dtype=torch.bfloat16)
y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
dx = torch.empty(T, K, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
wfb, wbb = _cabi.lib.vft_workspace_bytes(_cabi.OP_FWD, T, N, K, 0), _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD_DX, T, N, K, 0)
wf = torch.empty(max(wfb, 4), dtype=torch.uint8, device=dev); wb = torch.empty(max(wbb, 4), dtype=torch.uint8, device=dev)
for _ in range(reps):
    _cabi.check(_cabi.lib.vft_qlora_fwd(x.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, None, None, 0, 0.0, y.data_ptr(), None, None, None, wf.data_ptr() if wfb else None, wfb, TC, TA, st))
    _cabi.check(_cabi.lib.vft_qlora_bwd_dx(g.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, None, 0, 0.0, dx.data_ptr(), None, None, wb.data_ptr() if wbb else None, wbb, TC, TA, st))
torch.cuda.synchronize()
print("ok")
