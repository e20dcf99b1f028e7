"""Time vft_nf4_quantize (device-resident weights, preallocated outputs, CUDA-graph replay) on AuraFlow DiT shapes.
Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi

SHAPES = {  # AuraFlow 6.8B DiT, include "denoiser.", exclude t_embedder/final_linear/modF (SURVEY.md 8d cfg 2)
    (3072, 3072): 4 * 8 + 32 * 4, (8192, 3072): 4 * 4 + 32 * 2, (3072, 8192): 4 * 2 + 32 * 1,
    (18432, 3072): 4 * 2 + 32 * 1, (3072, 2048): 1, (3072, 16): 1,
}

def whole_set(dt=torch.float16, code=None, reps=6, verbose=False):
    """GB/s of vft_nf4_quantize over the whole synthetic AuraFlow DiT set (322 tensors, weights device-resident):
    every distinct shape timed from a CUDA graph, weighted by its count.  Returns (ms, GB/s)."""
    code = _cabi.F16 if dt == torch.float16 else _cabi.BF16
    dev = torch.device("cuda")
    total_bytes = total_us = 0.0
    for (n_, k_), count in SHAPES.items():
        n = n_ * k_
        ws = [(torch.randn(n_, k_, device=dev) * 0.02).to(dt) for _ in range(min(max(2, int(300e6 // (2 * n)) + 1), 6))]
        packed = torch.empty((n + 1) // 2, dtype=torch.uint8, device=dev)
        absmax = torch.empty((n + 63) // 64, dtype=torch.float32, device=dev)
        side = torch.cuda.Stream()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st = side.cuda_stream
            for w in ws: _cabi.check(_cabi.lib.vft_nf4_quantize(w.data_ptr(), code, n, 64, packed.data_ptr(), absmax.data_ptr(), st))
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for i in range(reps):
                    _cabi.check(_cabi.lib.vft_nf4_quantize(ws[i % len(ws)].data_ptr(), code, n, 64, packed.data_ptr(), absmax.data_ptr(), st))
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): g.replay()
        b.record(); torch.cuda.synchronize()
        us = a.elapsed_time(b) / (3 * reps) * 1e3
        total_bytes += 2.5625 * n * count; total_us += us * count
        del ws, g
    return total_us / 1e3, total_bytes / total_us / 1e3


def whole_set_batched(dt=torch.float16, reps=3):
    """The same set through vft_nf4_quantize_many (96 tensors per launch): every tensor of the model resident at once
    (13.6 GB of fp16 weights + 3.8 GB of outputs), timed with CUDA events around the whole call.  Returns (ms, GB/s)."""
    import ctypes
    code = _cabi.F16 if dt == torch.float16 else _cabi.BF16
    dev = torch.device("cuda")
    ws = []
    for (n_, k_), count in SHAPES.items():
        base = (torch.randn(n_, k_, device=dev) * 0.02).to(dt)
        ws += [base] + [base.clone() for _ in range(count - 1)]
    outs = [(torch.empty((w.numel() + 1) // 2, dtype=torch.uint8, device=dev), torch.empty((w.numel() + 63) // 64, dtype=torch.float32, device=dev)) for w in ws]
    k = len(ws)
    src = (ctypes.c_void_p * k)(*[w.data_ptr() for w in ws]); ns = (ctypes.c_int64 * k)(*[w.numel() for w in ws])
    pk = (ctypes.c_void_p * k)(*[o[0].data_ptr() for o in outs]); am = (ctypes.c_void_p * k)(*[o[1].data_ptr() for o in outs])
    st = torch.cuda.current_stream().cuda_stream
    call = lambda: _cabi.check(_cabi.lib.vft_nf4_quantize_many(k, src, code, ns, 64, pk, am, st))
    call(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): call()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / reps
    total = 2.5625 * sum(w.numel() for w in ws)
    return ms, total / ms / 1e6, k


def main():
    dev = torch.device("cuda")
    total_bytes = total_us = 0.0
    for dt, code in ((torch.bfloat16, _cabi.BF16), (torch.float16, _cabi.F16)):
        total_bytes = total_us = 0.0
        for (n_, k_), count in SHAPES.items():
            n = n_ * k_
            nset = max(2, int(300e6 // (2 * n)) + 1)
            ws = [(torch.randn(n_, k_, device=dev) * 0.02).to(dt) for _ in range(min(nset, 6))]
            packed = torch.empty((n + 1) // 2, dtype=torch.uint8, device=dev)
            absmax = torch.empty((n + 63) // 64, dtype=torch.float32, device=dev)
            side = torch.cuda.Stream()
            g = torch.cuda.CUDAGraph()
            reps = 12
            with torch.cuda.stream(side):
                st = side.cuda_stream
                for w in ws: _cabi.check(_cabi.lib.vft_nf4_quantize(w.data_ptr(), code, n, 64, packed.data_ptr(), absmax.data_ptr(), st))
                side.synchronize()
                with torch.cuda.graph(g, stream=side):
                    for i in range(reps):
                        _cabi.check(_cabi.lib.vft_nf4_quantize(ws[i % len(ws)].data_ptr(), code, n, 64, packed.data_ptr(), absmax.data_ptr(), st))
            g.replay(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): g.replay()
            b.record(); torch.cuda.synchronize()
            us = a.elapsed_time(b) / (5 * reps) * 1e3
            byts = 2.5625 * n
            print(f"  {str(dt)[6:]:9s} [{n_:5d},{k_:5d}] x{count:3d}: {us:8.2f} us  {byts / us / 1e3:7.1f} GB/s")
            total_bytes += byts * count; total_us += us * count
        print(f"{str(dt)[6:]}: AuraFlow DiT set (322 tensors, 6.80 G elements): {total_us / 1e3:.2f} ms, {total_bytes / total_us / 1e3:.0f} GB/s "
              f"({total_bytes / total_us / 1e3 / 6452.2 * 100:.1f} % of measured HBM copy bandwidth)")

if __name__ == "__main__":
    main()
    ms, gbs, k = whole_set_batched(torch.float16)
    print(f"float16: AuraFlow DiT set through vft_nf4_quantize_many ({k} tensors, one launch per <= 96 tensors of equal size): {ms:.2f} ms, {gbs:.0f} GB/s "
          f"({gbs / 6452.2 * 100:.1f} % of measured HBM copy bandwidth)")
