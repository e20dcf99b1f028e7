"""Time the fused tcgen05 kernel alone (forward and backward launches, NF4-only) for a shape; used with the
VFT_TC_DEBUG triage mask.  Not part of the product."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

def main():
    T, N, K = [int(v) for v in (sys.argv[1:4] if len(sys.argv) > 3 else (4096, 3072, 3072))]
    dev = torch.device("cuda")
    w = (torch.randn(N, K, device=dev) * 0.02).to(torch.bfloat16)
    packed, absmax = ops.nf4_quantize(w)
    tiles = None if os.environ.get('VFT_NOTILE') else ops.nf4_tile_weight(packed, absmax, N, K)
    TC, TA = (tiles[0].data_ptr(), tiles[1].data_ptr()) if tiles else (None, None)
    xs = [torch.randn(T, K, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    gs = [torch.randn(T, N, device=dev, dtype=torch.bfloat16) for _ in range(4)]
    y = torch.empty(T, N, device=dev, dtype=torch.bfloat16)
    dx = torch.empty(T, K, device=dev, dtype=torch.bfloat16)
    st = torch.cuda.current_stream().cuda_stream
    r = int(os.environ.get("R", "0"))
    A = (torch.randn(max(r, 1), K, device=dev) * 0.02).to(torch.bfloat16)
    B = (torch.randn(N, max(r, 1), device=dev) * 0.02).to(torch.bfloat16)
    ts = torch.empty(T, 64, device=dev, dtype=torch.bfloat16)
    bt = torch.empty(16 * ((max(r, 1) + 15) // 16), N, device=dev, dtype=torch.bfloat16)
    AP, BP, TS, BT = (A.data_ptr(), B.data_ptr(), ts.data_ptr(), bt.data_ptr()) if r else (None, None, None, None)
    BT_F = None if os.environ.get("VFT_NO_BT") else BT  # forward: do not ask for bt_save (the backward still reads the buffer)
    WFB = _cabi.lib.vft_workspace_bytes(0, T, N, K, 0); WBB = _cabi.lib.vft_workspace_bytes(1, T, N, K, 0)
    wf_t = torch.empty(max(WFB, 4), dtype=torch.uint8, device=dev); wb_t = torch.empty(max(WBB, 4), dtype=torch.uint8, device=dev)
    WF = wf_t.data_ptr() if WFB else None; WB = wb_t.data_ptr() if WBB else None
    def fwd(i):
        _cabi.check(_cabi.lib.vft_qlora_fwd(xs[i % 4].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, AP, BP, r, 1.0 / max(r, 1), y.data_ptr(), TS, BT_F, None, WF, WFB, TC, TA, st))
    def bwd(i):
        _cabi.check(_cabi.lib.vft_qlora_bwd_dx(gs[i % 4].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, AP, BP, r, 1.0 / max(r, 1), dx.data_ptr(), TS, BT, WB, WBB, TC, TA, st))
    res = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for i in range(5): fn(i)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for i in range(40): fn(i)
        b.record(); torch.cuda.synchronize()
        res[name] = a.elapsed_time(b) / 40 * 1e3
    if False:
        import ctypes
        fwd(0); torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 512)()
        _cabi.lib.vft_debug_tc_timeline(buf, 512)
        names = ["entry", "setup done", "first stage ready", "block 8 ready", "last block ready", "accum ready", "epilogue done", "exit"]
        c0, n0 = buf[0], buf[1]
        for i, nm in enumerate(names):
            print(f"   {nm:18s} +{buf[2*i]-c0:8d} cyc  +{(buf[2*i+1]-n0)/1e3:8.2f} us")
    if int(os.environ.get("VFT_TC_DEBUG", "0")) & 16 and hasattr(_cabi.lib, "vft_debug_tc2_timeline"):
        import ctypes
        for nm, fn in (("fwd", fwd), ("bwd", bwd)):
            fn(0); torch.cuda.synchronize()
            R, C = 7, 256
            buf = (ctypes.c_ulonglong * (R * C))()
            _cabi.lib.vft_debug_tc2_timeline(buf, R * C)
            rows = [[buf[r * C + c] for c in range(C)] for r in range(R)]
            t0 = min(v for v in rows[5][:4] if v) if any(rows[5][:4]) else rows[0][0]
            def rel(v): return (v - t0) if v else -1
            print(f"  [{nm}] pair-kernel timeline (cycles since first producer wait), leader CTA of pair 0")
            print("   step: mma_full_seen  mma_commit   | prod_empty_seen")
            for g in list(range(0, 12)) + list(range(40, 56)) + list(range(90, 100)):
                print(f"   {g:4d}: {rel(rows[0][g]):10d} {rel(rows[1][g]):10d}   | {rel(rows[5][g]):10d}")
            print("   decode group0 (steps 0,4,8..): empty_seen, stores issued, arrived")
            for i in list(range(0, 6)) + list(range(10, 14)):
                print(f"   {4*i:4d}: {rel(rows[2][i]):10d} {rel(rows[6][i]):10d} {rel(rows[3][i]):10d}")
            print("   epilogue per tile (t_full seen, t box written, acc_full seen, drained):", [rel(v) for v in rows[4][:12]])
            print("   split exchange: stores issued, fenced, all items arrived, reduced:", [rel(v) for v in rows[6][240:244]])
            print("   side product: done seen by epilogue, published, generation seen by producer:", [rel(v) for v in rows[6][250:253]])
            steps = [v for v in rows[0] if v]
            if len(steps) > 20:
                d = [steps[i + 1] - steps[i] for i in range(8, len(steps) - 1)]
                print(f"   mma steps seen {len(steps)}, median step interval {sorted(d)[len(d)//2]} cyc, mean {sum(d)/len(d):.0f}")
    fl = 2 * T * N * K
    print(f"VFT_TC_DEBUG={os.environ.get('VFT_TC_DEBUG','0'):>2} T={T} N={N} K={K}: fwd {res['fwd']:.1f} us ({fl/res['fwd']/1e6:.0f} TF/s)  bwd {res['bwd']:.1f} us ({fl/res['bwd']/1e6:.0f} TF/s)")

main()
