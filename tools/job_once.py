"""One forward + one-call backward through the C ABI, checked against torch (triage of the in-launch dA/dB job)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops
T, N, K, r = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (4096, 3072, 3072, 16))]
dev = torch.device("cuda"); bf = torch.bfloat16
torch.manual_seed(0)
w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
packed, absmax = ops.nf4_quantize(w)
wd = ops.nf4_dequantize(packed, absmax, (N, K), bf).float()
tiles = ops.nf4_tile_weight(packed, absmax, N, K)
x = torch.randn(T, K, device=dev, dtype=bf); dy = torch.randn(T, N, device=dev, dtype=bf)
A = (torch.randn(r, K, device=dev) * 0.05).to(bf); B = (torch.randn(N, r, device=dev) * 0.05).to(bf)
y = torch.empty(T, N, device=dev, dtype=bf); dx = torch.empty(T, K, device=dev, dtype=bf)
ts = torch.zeros(T, 64, device=dev, dtype=bf); dts = torch.zeros(T, 64, device=dev, dtype=bf)
rp = 16 * ((r + 15) // 16)
bt = torch.empty(rp, N, device=dev, dtype=bf); tt = torch.empty(rp, T, device=dev, dtype=bf)
dA = torch.full_like(A, float("nan")); dB = torch.full_like(B, float("nan"))
L = _cabi.lib; st = torch.cuda.current_stream().cuda_stream; s = 1.0 / r
_cabi.check(L.vft_qlora_fwd(x.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, A.data_ptr(), B.data_ptr(), r, s,
                            y.data_ptr(), ts.data_ptr(), bt.data_ptr(), tt.data_ptr(), None, 0, tiles[0].data_ptr(), tiles[1].data_ptr(), st))
torch.cuda.synchronize()
t_ref = (x.float() @ A.float().t())
print("fwd ok; tt err", float((tt[:r].float() - t_ref.t().to(bf).float()).abs().max()), "bt err", float((bt[:r].float() - (s * B.float()).t().to(bf).float()).abs().max()))
wsb = L.vft_workspace_bytes(_cabi.OP_BWD, T, N, K, r); ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
t0 = time.time()
_cabi.check(L.vft_qlora_bwd(dy.data_ptr(), x.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, A.data_ptr(), B.data_ptr(), r, s,
                            ts.data_ptr(), tt.data_ptr(), bt.data_ptr(), dx.data_ptr(), dA.data_ptr(), dB.data_ptr(), dts.data_ptr(),
                            ws.data_ptr(), wsb, tiles[0].data_ptr(), tiles[1].data_ptr(), st))
try:
    torch.cuda.synchronize()
finally:
    print(f"bwd returned after {time.time() - t0:.2f} s")
dt_ref = (s * dy.float() @ B.float()).to(bf).float()
dA_ref = dt_ref.t() @ x.float(); dB_ref = s * dy.float().t() @ t_ref.to(bf).float()
dx_ref = dy.float() @ wd + dt_ref @ A.float()
rel = lambda a, b: float((a.float() - b).norm() / b.norm())
print("rel-L2  dx %.2e  dA %.2e  dB %.2e  dt %.2e" % (rel(dx, dx_ref), rel(dA, dA_ref), rel(dB, dB_ref), rel(dts[:, :r], dt_ref)))
