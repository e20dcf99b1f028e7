"""Where does tcgen05.mma cta_group::2 with M = 128 put its 64 rows per CTA in tensor memory?  Runs the forward launch
with the side product forced to M = 128 (VFT_TC_DEBUG=512), dumps the raw accumulator lanes of the first pair and
matches them against the expected rows of t = x . A^T.  Not part of the product."""
import ctypes, os, sys
os.environ["VFT_TC_DEBUG"] = "512"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

T, N, K, r = 4096, 3072, 3072, 16
dev = torch.device("cuda")
w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
packed, absmax = ops.nf4_quantize(w)
tiles = ops.nf4_tile_weight(packed, absmax, N, K)
x = torch.randn(T, K, device=dev, dtype=torch.bfloat16)
A = (torch.randn(r, K, device=dev) * 0.05).to(torch.bfloat16)
B = (torch.randn(N, r, device=dev) * 0.05).to(torch.bfloat16)
y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
ts = torch.zeros(T, 64, device=dev, dtype=torch.bfloat16)
st = torch.cuda.current_stream().cuda_stream
_cabi.check(_cabi.lib.vft_qlora_fwd(x.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, A.data_ptr(), B.data_ptr(),
                                    r, 1.0 / r, y.data_ptr(), ts.data_ptr(), None, None, None, 0, tiles[0].data_ptr(), tiles[1].data_ptr(), st))
torch.cuda.synchronize()
buf = (ctypes.c_float * (2 * 128 * 32))()
fn = _cabi.lib.vft_debug_tc2_p0dump
fn.argtypes = [ctypes.c_void_p, ctypes.c_int]
fn(buf, 2 * 128 * 32)
dump = torch.tensor(list(buf)).view(2, 128, 32)
t = (x.float() @ A.float().t()).cpu()  # [T, 16]
rows = 32  # p0_rows at this shape: ceil(4096 / 144) -> 32
for cta in range(2):
    print(f"CTA {cta}: token rows {cta * rows}..{cta * rows + rows - 1}")
    for i in range(rows):
        want = t[cta * rows + i]
        # find (lane, column offset) whose 16 (or 8 + 8) values match
        hits = []
        for lane in range(128):
            for c0 in range(0, 17):
                got = dump[cta, lane, c0:c0 + 16]
                if got.numel() == 16 and torch.allclose(got, want, rtol=2e-2, atol=2e-2):
                    hits.append((lane, c0, 16))
            for c0 in range(0, 25):
                got = dump[cta, lane, c0:c0 + 8]
                if torch.allclose(got, want[:8], rtol=2e-2, atol=2e-2):
                    hits.append((lane, c0, "lo8"))
                if torch.allclose(got, want[8:], rtol=2e-2, atol=2e-2):
                    hits.append((lane, c0, "hi8"))
        print(f"  row {i:2d}: {hits[:6]}")
