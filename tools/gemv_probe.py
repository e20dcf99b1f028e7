"""Few-token forward (T <= 8): achieved GB/s on the packed-weight stream, HBM-cold.

The weight is replicated so that the copies visited between two uses of the same copy exceed the 126 MB L2
(>= 256 MB in rotation); launches are replayed from a CUDA graph through the C ABI.  Compares the streaming kernel
(auto path for T <= 8, qlora_gemv.cu) with the tcgen05 split-K form of the persistent kernel (forced path 1).
Not part of the product."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

def probe(N, K, T, path, bias=False, dev="cuda"):
    bf = torch.bfloat16
    wbytes = N * K * 0.5625
    copies = max(2, int(270e6 // wbytes) + 1)
    w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
    p0, a0 = ops.nf4_quantize(w)
    packs = [p0.clone() for _ in range(copies)]
    ams = [a0.clone() for _ in range(copies)]
    tiles = [ops.nf4_tile_weight(p, a, N, K) for p, a in zip(packs, ams)] if path == 1 else [None] * copies
    x = torch.randn(T, K, device=dev, dtype=bf)
    y = torch.empty(T, N, device=dev, dtype=bf)
    bv = torch.randn(N, device=dev, dtype=bf) if bias else None
    wsb = _cabi.lib.vft_workspace_bytes(_cabi.OP_FWD, T, N, K, 0)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
    _cabi.lib.vft_force_path(path)
    side = torch.cuda.Stream()
    def call(i, st):
        tl = tiles[i % copies]
        _cabi.check(_cabi.lib.vft_qlora_fwd(x.data_ptr(), T, packs[i % copies].data_ptr(), ams[i % copies].data_ptr(), N, K, 64, 2, 2,
                                            None if bv is None else bv.data_ptr(), None, None, 0, 0.0, y.data_ptr(), None, None, None,
                                            ws.data_ptr() if wsb else None, wsb, tl[0].data_ptr() if tl else None,
                                            tl[1].data_ptr() if tl else None, st))
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        st = side.cuda_stream
        for i in range(copies): call(i, st)
        side.synchronize()
        used = _cabi.lib.vft_last_path()
        with torch.cuda.graph(g, stream=side):
            for i in range(copies): call(i, st)
    g.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): g.replay()
    b.record(); torch.cuda.synchronize()
    _cabi.lib.vft_force_path(0)
    us = a.elapsed_time(b) / (5 * copies) * 1e3
    return {"N": N, "K": K, "T": T, "path": used, "us": round(us, 2), "weight_GBs": round(wbytes / us / 1e3, 1), "copies": copies}

def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    rows = []
    for N, K in ((18432, 3072), (9216, 1024), (3072, 3072), (3072, 8192)):
        for T in (1, 2, 4, 8):
            for path in (0, 1):
                r = probe(N, K, T, path)
                r["frac_of_hbm_peak"] = round(r["weight_GBs"] / peak, 3)
                rows.append(r)
                print(r, flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "gemv_probe.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
