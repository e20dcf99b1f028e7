"""Layer census of the three model families: every NF4(+LoRA) Linear of one training step at its own
token count, forward + backward through the C ABI, CUDA-graph timed.  Reports per-shape time / TFLOP/s / packed-weight
GB/s and the sum over the step (linear layers only, gradient checkpointing = one extra forward).  Not part of the product."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    sys.path.insert(0, p)
import torch
from vft_b200 import _cabi, ops

R = 16
def auraflow(B=2):
    t1, tx, tc = B * 4360, B * 4096, B * 264
    L = []
    L += [("single.attn.w1qkvo", 3072, 3072, t1, True, False, 32 * 4), ("single.mlp.c_fc1/2", 8192, 3072, t1, True, False, 32 * 2),
          ("single.mlp.c_proj", 3072, 8192, t1, True, False, 32)]
    L += [("double.attn.w2qkvo", 3072, 3072, tx, True, False, 4 * 4), ("double.mlpX.c_fc1/2", 8192, 3072, tx, True, False, 4 * 2),
          ("double.mlpX.c_proj", 3072, 8192, tx, True, False, 4)]
    L += [("double.attn.w1qkvo (NF4 only)", 3072, 3072, tc, False, False, 4 * 4), ("double.mlpC.c_fc1/2 (NF4 only)", 8192, 3072, tc, False, False, 4 * 2),
          ("double.mlpC.c_proj (NF4 only)", 3072, 8192, tc, False, False, 4)]
    L += [("mod*.1 (NF4 only, T = batch)", 18432, 3072, B, False, False, 40), ("cond_seq_linear", 3072, 2048, B * 256, False, False, 1)]
    return L
def lumina2(B=1):
    t, tc = B * (4096 + 256), B * 256
    shapes = [("attention.qkv", 3840, 2304), ("attention.out", 2304, 2304), ("feed_forward.w1/w3", 9216, 2304), ("feed_forward.w2", 2304, 9216)]
    cnt = {"attention.qkv": 1, "attention.out": 1, "feed_forward.w1/w3": 2, "feed_forward.w2": 1}
    L = []
    for nm, n, k in shapes:
        L.append((f"layers.{nm}", n, k, t, True, False, 26 * cnt[nm]))
        L.append((f"context_refiner.{nm}", n, k, tc, True, False, 2 * cnt[nm]))
        L.append((f"noise_refiner.{nm} (NF4 only)", n, k, B * 4096, False, False, 2 * cnt[nm]))
    L.append(("adaLN_modulation.1 (NF4 only, bias)", 9216, 1024, B, False, True, 28))
    return L
def sdxl(B=2, w=1024, h=1024):
    L = []
    for C, blocks in ((640, 10), (1280, 60)):
        T = B * (w // (16 if C == 640 else 32)) * (h // (16 if C == 640 else 32))
        L += [(f"C{C}.attn1.to_q/k/v", C, C, T, True, False, 3 * blocks), (f"C{C}.attn1.to_out", C, C, T, True, True, blocks),
              (f"C{C}.attn2.to_q", C, C, T, True, False, blocks), (f"C{C}.attn2.to_k/v (77 text tokens)", C, 2048, B * 77, True, False, 2 * blocks),
              (f"C{C}.attn2.to_out", C, C, T, True, True, blocks), (f"C{C}.ff.net.0.proj", 8 * C, C, T, True, True, blocks),
              (f"C{C}.ff.net.2", C, 4 * C, T, True, True, blocks)]
    return L

def time_layer(N, K, T, lora, bias, dev):
    bf = torch.bfloat16
    w = (torch.randn(N, K, device=dev) * 0.02).to(bf)
    packed, absmax = ops.nf4_quantize(w)
    tiles = ops.nf4_tile_weight(packed, absmax, N, K)
    TC, TA = (tiles[0].data_ptr(), tiles[1].data_ptr()) if tiles else (None, None)
    r_ = R if lora else 0
    nset = 3
    xs = [torch.randn(T, K, device=dev, dtype=bf) for _ in range(nset)]
    gs = [torch.randn(T, N, device=dev, dtype=bf) for _ in range(nset)]
    r = R if lora else 0
    A = (torch.randn(R, K, device=dev) * 0.02).to(bf) if lora else None
    B_ = (torch.randn(N, R, device=dev) * 0.02).to(bf) if lora else None
    bv = (torch.randn(N, device=dev) * 0.1).to(bf) if bias else None
    y = torch.empty(T, N, device=dev, dtype=bf); dx = torch.empty(T, K, device=dev, dtype=bf)
    ts = torch.empty(T, 64, device=dev, dtype=bf); dts = torch.empty(T, 64, device=dev, dtype=bf)
    dA = torch.empty(R, K, device=dev, dtype=bf); dB = torch.empty(N, R, device=dev, dtype=bf)
    bt = torch.empty(16 * ((R + 15) // 16), N, device=dev, dtype=bf)
    tt = torch.empty(16 * ((R + 15) // 16), T, device=dev, dtype=bf)
    wsb2 = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD, T, N, K, R)
    ws2 = torch.empty(max(wsb2, 4), dtype=torch.uint8, device=dev)
    wsb = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD_DAB, T, N, K, R)
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)
    wf_b = _cabi.lib.vft_workspace_bytes(_cabi.OP_FWD, T, N, K, r_) if True else 0
    wb_b = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD_DX, T, N, K, r_)
    wf = torch.empty(max(wf_b, 4), dtype=torch.uint8, device=dev)
    wb = torch.empty(max(wb_b, 4), dtype=torch.uint8, device=dev)
    L = _cabi.lib
    P = lambda t: None if t is None else t.data_ptr()
    side = torch.cuda.Stream()
    def fwd(i, st):
        _cabi.check(L.vft_qlora_fwd(xs[i % nset].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, P(bv), P(A), P(B_), r, 1.0 / R, y.data_ptr(), P(ts) if lora else None, P(bt) if lora else None, P(tt) if lora else None, wf.data_ptr() if wf_b else None, wf_b, TC, TA, st))
    def bwd(i, st):
        if lora:  # the whole backward (dx, dt, dA, dB) in one C-ABI call: one launch where the persistent kernel takes it
            _cabi.check(L.vft_qlora_bwd(gs[i % nset].data_ptr(), xs[i % nset].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2,
                                        P(A), P(B_), R, 1.0 / R, ts.data_ptr(), tt.data_ptr(), bt.data_ptr(), dx.data_ptr(), dA.data_ptr(),
                                        dB.data_ptr(), dts.data_ptr(), ws2.data_ptr(), wsb2, TC, TA, st))
        else:
            _cabi.check(L.vft_qlora_bwd_dx(gs[i % nset].data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, 64, 2, 2, None, None, 0, 0.0,
                                           dx.data_ptr(), None, None, wb.data_ptr() if wb_b else None, wb_b, TC, TA, st))
    out = {}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        reps = 6
        g = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            st = side.cuda_stream
            for i in range(2): fn(i, st)
            side.synchronize()
            with torch.cuda.graph(g, stream=side):
                for i in range(reps): fn(i, st)
        g.replay(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): g.replay()
        b.record(); torch.cuda.synchronize()
        out[name] = a.elapsed_time(b) / (3 * reps) * 1e3
    return out["fwd"], out["bwd"], L.vft_last_path()

def model_step(layers, dev, verbose=None):
    """Sum of (2 x forward + backward) over the NF4(+LoRA) Linear layers of one training step."""
    tot_us = tot_fl = 0.0
    cache, rows = {}, []
    for nm, N, K, T, lora, bias, count in layers:
        key = (N, K, T, lora, bias)
        if key not in cache:
            cache[key] = time_layer(N, K, T, lora, bias, dev)
        f, b, _ = cache[key]
        fl_f = 2 * T * N * K + (2 * T * R * (N + K) if lora else 0)
        fl_b = 2 * T * N * K + (4 * T * R * (N + K) if lora else 0)
        tot_us += (2 * f + b) * count
        tot_fl += (2 * fl_f + fl_b) * count
        rows.append((nm, N, K, T, lora, count, f, b, fl_f, fl_b))
    return tot_us / 1e3, tot_fl / tot_us / 1e6, rows


def main():
    dev = torch.device("cuda")
    models = {"auraflow_6.8B_B2_1024": auraflow(2), "lumina2_2.6B_B1_1024": lumina2(1), "sdxl_unet_B2_1024": sdxl(2)}
    sel = sys.argv[1:] or list(models)
    report = {}
    for mname in sel:
        rows, tot_us, tot_fl = [], 0.0, 0.0
        cache = {}
        for nm, N, K, T, lora, bias, count in models[mname]:
            key = (N, K, T, lora, bias)
            if key not in cache: cache[key] = time_layer(N, K, T, lora, bias, dev)
            f, b, path = cache[key]
            fl_f = 2 * T * N * K + (2 * T * R * (N + K) if lora else 0)
            fl_b = 2 * T * N * K + (4 * T * R * (N + K) if lora else 0)
            step_us = (2 * f + b) * count           # gradient checkpointing: forward runs twice
            tot_us += step_us; tot_fl += (2 * fl_f + fl_b) * count
            wbytes = N * K * 0.5625
            rows.append({"layer": nm, "N": N, "K": K, "T": T, "lora": lora, "count": count, "fwd_us": round(f, 1), "bwd_us": round(b, 1),
                         "fwd_tflops": round(fl_f / f / 1e6, 0), "bwd_tflops": round(fl_b / b / 1e6, 0), "fwd_weight_GBs": round(wbytes / f / 1e3, 0)})
            print(f"  {mname:24s} {nm:40s} N{N:6d} K{K:5d} T{T:6d} x{count:3d}  fwd {f:8.1f} us {fl_f / f / 1e6:6.0f} TF/s {wbytes / f / 1e3:6.0f} GB/s(W)   bwd {b:8.1f} us {fl_b / b / 1e6:6.0f} TF/s")
        print(f"{mname}: NF4(+LoRA) Linear layers of one training step (2x fwd + bwd): {tot_us / 1e3:.2f} ms, {tot_fl / tot_us / 1e6:.0f} TFLOP/s average")
        report[mname] = {"linear_ms_per_step": tot_us / 1e3, "avg_tflops": tot_fl / tot_us / 1e6, "layers": rows}
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", "census.json"), "w"), indent=1)

if __name__ == "__main__":
    main()
