"""Time the three C-ABI calls of one NF4+LoRA layer step (fwd, bwd_dx, bwd_dab), with and without the adapter,
so that the cost of the rank-r side kernels is visible.  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

def main():
    T, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 3072, 3072))]
    r = int(sys.argv[4]) if len(sys.argv) > 4 else 16
    global NSET
    NSET = int(os.environ.get("NSET", "4"))
    dev = torch.device("cuda")
    bf = torch.bfloat16
    w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
    packed, absmax = ops.nf4_quantize(w)
    tiles = ops.nf4_tile_weight(packed, absmax, N, K)
    TC, TA = tiles[0].data_ptr(), tiles[1].data_ptr()
    xs = [torch.randn(T, K, device=dev, dtype=bf) for _ in range(4)]
    gs = [torch.randn(T, N, device=dev, dtype=bf) for _ in range(4)]
    A = (torch.randn(r, K, device=dev) * 0.02).to(bf)
    B = (torch.randn(N, r, device=dev) * 0.02).to(bf)
    y = torch.empty(T, N, device=dev, dtype=bf); dx = torch.empty(T, K, device=dev, dtype=bf)
    ts = torch.empty(T, 64, device=dev, dtype=bf); dts = torch.empty(T, 64, device=dev, dtype=bf)
    bt = torch.empty(16 * ((r + 15) // 16), N, device=dev, dtype=bf)
    BT = None if os.environ.get("VFT_NO_BT") else bt.data_ptr()
    tt = torch.empty(16 * ((r + 15) // 16), T, device=dev, dtype=bf)
    TT = None if os.environ.get("VFT_NO_BT") else tt.data_ptr()
    wsb2 = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD, T, N, K, r)
    ws2 = torch.empty(max(wsb2, 4), dtype=torch.uint8, device=dev)
    dA = torch.empty_like(A); dB = torch.empty_like(B)
    wsb = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD_DAB, T, N, K, r)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L = _cabi.lib
    calls = {
        "fwd  (NF4 only)": lambda i: L.vft_qlora_fwd(xs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, None, None, 0, 0.0, y.data_ptr(), None, None, None, None, 0, TC, TA, st),
        "fwd  (+LoRA)   ": lambda i: L.vft_qlora_fwd(xs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, A.data_ptr(), B.data_ptr(), r, 1.0 / r, y.data_ptr(), ts.data_ptr(), BT, TT, None, 0, TC, TA, st),
        "bwd  (NF4 only)": lambda i: L.vft_qlora_bwd_dx(gs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, None, 0, 0.0, dx.data_ptr(), None, None, None, 0, TC, TA, st),
        "bwd  (+LoRA)   ": lambda i: L.vft_qlora_bwd_dx(gs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, A.data_ptr(), B.data_ptr(), r, 1.0 / r, dx.data_ptr(), dts.data_ptr(), BT, None, 0, TC, TA, st),
        "dt only        ": lambda i: L.vft_qlora_bwd_dx(gs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, A.data_ptr(), B.data_ptr(), r, 1.0 / r, None, dts.data_ptr(), None, None, 0, TC, TA, st),
        "bwd  (one call)": lambda i: L.vft_qlora_bwd(gs[i % NSET].data_ptr(), xs[i % NSET].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, A.data_ptr(), B.data_ptr(), r, 1.0 / r, ts.data_ptr(), TT, BT, dx.data_ptr(), dA.data_ptr(), dB.data_ptr(), dts.data_ptr(), ws2.data_ptr(), wsb2, TC, TA, st),
        "dA/dB          ": lambda i: L.vft_lora_bwd_dab(gs[i % NSET].data_ptr(), xs[i % NSET].data_ptr(), ts.data_ptr(), dts.data_ptr(), T, N, K, r, 2, 1.0 / r, dA.data_ptr(), dB.data_ptr(), ws.data_ptr(), wsb, st),
    }
    tot = {}
    REP = 20
    for name, fn in calls.items():
        for i in range(3): _cabi.check(fn(i))
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            sst = side.cuda_stream
            # the lambdas read `st` from the enclosing scope at call time
            st_saved = st
            st = sst
            with torch.cuda.graph(graph, stream=side):
                for i in range(REP): _cabi.check(fn(i))
            st = st_saved
        graph.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5): graph.replay()
        b.record(); torch.cuda.synchronize()
        tot[name] = a.elapsed_time(b) / (5 * REP) * 1e3
        print(f"  {name}: {tot[name]:7.1f} us")
    if int(os.environ.get("VFT_TC_DEBUG", "0")) & 16:
        import ctypes
        names = ["entry", "setup done", "first full", "last full", "acc in regs", "after cluster sync", "stored", "exit"]
        for nm in ("dt only        ", "dA/dB          "):
            _cabi.check(calls[nm](0)); torch.cuda.synchronize()
            buf = (ctypes.c_ulonglong * 16)()
            L.vft_debug_side_timeline(buf, 16)
            print(f"  side-kernel timeline [{nm.strip()}] (ns since entry): " + ", ".join(f"{n} {buf[i]-buf[0]}" for i, n in enumerate(names)))
    print(f"  step with the one-call backward: {tot['fwd  (+LoRA)   '] + tot['bwd  (one call)']:.1f} us")
    step = tot["fwd  (+LoRA)   "] + tot["bwd  (+LoRA)   "] + tot["dA/dB          "]
    fl = 4 * T * N * K + 6 * T * r * (N + K)
    print(f"T={T} N={N} K={K} r={r}: step (3 calls, CUDA-graph replay) {step:.1f} us -> {fl / step / 1e6:.0f} TF/s; "
          f"adapter side kernels {step - tot['fwd  (NF4 only)'] - tot['bwd  (NF4 only)']:.1f} us")

main()
