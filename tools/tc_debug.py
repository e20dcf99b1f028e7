"""GPU debugging aid: run the tcgen05 path on structured inputs and print where it differs from the
dequantize kernel / generic kernels.  Not part of the product or the tests."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)

import torch

from vft_b200 import ops


def summarize(name, got, want):
    got, want = got.float(), want.float()
    diff = (got - want).abs()
    bad = diff > 1e-3 * want.abs().max().clamp_min(1e-6)
    print(f"[{name}] shape={tuple(got.shape)} max_abs={diff.max().item():.4e} bad={int(bad.sum())}/{bad.numel()} "
          f"ref_max={want.abs().max().item():.3e} got_max={got.abs().max().item():.3e} nan={int(torch.isnan(got).sum())}")
    if bad.any():
        rows = bad.any(dim=1).nonzero().flatten()
        cols = bad.any(dim=0).nonzero().flatten()
        print(f"   bad rows: n={rows.numel()} first={rows[:12].tolist()}  bad cols: n={cols.numel()} first={cols[:12].tolist()}")
        r, c = bad.nonzero()[0].tolist()
        print(f"   first bad [{r},{c}] got={got[r, c].item():.5f} want={want[r, c].item():.5f}")
        print("   got [r, c:c+8]  =", [round(v, 4) for v in got[r, c:c + 8].tolist()])
        print("   want[r, c:c+8]  =", [round(v, 4) for v in want[r, c:c + 8].tolist()])


def run(N, K, T, r=0, backward=False):
    g = torch.Generator(device="cuda").manual_seed(N * 7 + K)
    w = (torch.randn(N, K, generator=g, device="cuda") * 0.02).to(torch.bfloat16)
    packed, absmax = ops.nf4_quantize(w)
    wd = ops.nf4_dequantize(packed, absmax, (N, K), torch.bfloat16)
    x = torch.randn(T, K, generator=g, device="cuda").to(torch.bfloat16)
    dy = torch.randn(T, N, generator=g, device="cuda").to(torch.bfloat16)
    a = (torch.randn(r, K, generator=g, device="cuda") * 0.05).to(torch.bfloat16) if r else None
    b = (torch.randn(N, r, generator=g, device="cuda") * 0.05).to(torch.bfloat16) if r else None
    res = {}
    for path in (2, 1):
        ops.force_path(path)
        xg = x.clone().requires_grad_(True)
        ag = a.clone().requires_grad_(True) if r else None
        bg = b.clone().requires_grad_(True) if r else None
        try:
            y = ops.qlora_linear(xg, packed, absmax, None, ag, bg, 0.25, N, K, 64, torch.bfloat16)
            used = ops.last_path()
            y.backward(dy)
            torch.cuda.synchronize()
            res[path] = (y.detach(), xg.grad, used)
        except Exception as e:
            print(f"path {path} FAILED: {type(e).__name__}: {e}")
            ops.force_path(0)
            return
    ops.force_path(0)
    ref_y = x.float() @ wd.float().t()
    ref_dx = dy.float() @ wd.float()
    if r:
        t = (x.float() @ a.float().t()).to(torch.bfloat16).float()
        ref_y = ref_y + 0.25 * (t @ b.float().t())
        dt = (0.25 * (dy.float() @ b.float())).to(torch.bfloat16).float()
        ref_dx = ref_dx + dt @ a.float()
    tag = f"N{N} K{K} T{T} r{r}"
    print(f"== {tag}: paths used simt={res[2][2]} tc={res[1][2]}")
    summarize(tag + " simt y ", res[2][0], ref_y)
    summarize(tag + " tc   y ", res[1][0], ref_y)
    summarize(tag + " simt dx", res[2][1], ref_dx)
    summarize(tag + " tc   dx", res[1][1], ref_dx)


if __name__ == "__main__":
    torch.manual_seed(0)
    print(torch.cuda.get_device_name(0))
    run(128, 64, 64)
    run(128, 128, 64)
    run(256, 256, 300)
    run(384, 256, 300, r=16)
    run(3072, 3072, 4096, r=16)
