#!/usr/bin/env python
"""Benchmark of the QLoRA hot path on B200 (contract: one JSON line on stdout from rank 0).

Workload (BASELINE.json configs[0], the configuration the metric is quoted on): a single NF4 Linear
3072x3072 + LoRA r=16, forward + backward (dX, dA, dB) on 4096 tokens per GPU, bf16.  One "step" is one
forward+backward pass of that layer over one batch of synthetic activations.

  value     whole-job TFLOP/s with inputs resident in HBM, CUDA-event timed, max over ranks; a step is TWO launches
            (forward / backward, the adapter's side products and dA, dB inside them), replayed from CUDA graphs that
            hold a round of four steps each
  e2e       same metric through the module API with HOST (pinned) inputs: H2D of x, dy every step (uploaded on a copy
            stream one step ahead), D2H of dA, dB every step (consumed by the host one step behind)
  roofline  the two launches of the step (fused NF4-decode tcgen05 GEMM with the adapter inside) timed alone through the
            C ABI; the NF4-only launches beside them
  cpu_baseline  the oracle port (oracle/qlora_oracle.py) on the host cores, bounded sample, rank 0 at N=1 only

  extra     rank 0: NF4 quantize/pack GB/s (one tensor, and the whole AuraFlow weight set = BASELINE configs[1]);
            few-token (T = 2) weight-stream GB/s, HBM-cold; at N = 1 the layer censuses of an AuraFlow / Lumina2 / SDXL
            step (configs[3], [2], [4]), sibling projections as one launch (projection_groups), the Lumina2 step
            harness and the AuraFlow step with SDPA.  All ranks: `auraflow_qlora_step_dp`, the AuraFlow-6.8B QLoRA step
            harness (tools/auraflow_step.py: steps/s, samples/s, exposed all-reduce time) -- the second half of
            BASELINE's metric.

`--impl reference` times the reference's CPU path for the same layer (the oracle port: bitsandbytes itself is
not installable here) on the host cores.  Multi-GPU: weak scaling, each rank steps its own 4096 tokens and the
LoRA gradients are all-reduced over NCCL on a side stream through a 2-CTA communicator (the fused GEMM is a
persistent kernel on 72 of the 74 SM pairs: a collective that holds more SMs costs it a wave; measurements in
DESIGN.md section 5), overlapped with the next step.  stdout carries exactly one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_FEAT = K_FEAT = 3072
RANK = 16
TOKENS = 4096
METRIC = "nf4_lora_linear_fwd_bwd_tflops"
UNIT = "TFLOP/s"


def layer_flops(T, N=N_FEAT, K=K_FEAT, r=RANK):
    """4*T*N*K + 6*T*r*(N+K)."""
    return 4 * T * N * K + 6 * T * r * (N + K)


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"bf16_tflops": d["bf16_tflops"], "bf16_tflops_sustained": d.get("bf16_tflops_sustained"),
                "hbm_gbs": d["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


# ----------------------------------------------------------------------------- CPU arm
def make_cpu_case(T):
    import torch
    from oracle import nf4_oracle, qlora_oracle

    w = (torch.randn(N_FEAT, K_FEAT, generator=torch.Generator().manual_seed(0)) * 0.02).to(torch.bfloat16)
    p, a = nf4_oracle.nf4_quantize(w)
    w_deq = qlora_oracle.dequant_weight(p, a, (N_FEAT, K_FEAT), "bfloat16")
    x = torch.randn(T, K_FEAT, generator=torch.Generator().manual_seed(1)).to(torch.bfloat16)
    dy = torch.randn(T, N_FEAT, generator=torch.Generator().manual_seed(2)).to(torch.bfloat16)
    la = ((torch.rand(RANK, K_FEAT, generator=torch.Generator().manual_seed(3)) * 2 - 1) * (6.0 / K_FEAT) ** 0.5).to(torch.bfloat16)
    lb = (torch.randn(N_FEAT, RANK, generator=torch.Generator().manual_seed(4)) * 0.02).to(torch.bfloat16)
    return p, a, w_deq, x, dy, la, lb


def cpu_step(case, T):
    """What the reference does per call on its CPU path: dequantize W to bf16 (forward), F.linear + LoRA,
    autograd backward (MatMul4Bit.backward dequantizes again).  The dequantisation is the plain-C restatement
    (oracle/nf4_ref.c, OpenMP over all host threads -- bitsandbytes' CPU kernel is multi-threaded C++ too), the GEMMs
    are torch's CPU bf16 matmuls on all host threads."""
    from oracle import c_oracle, qlora_oracle

    p, a, _, x, dy, la, lb = case
    n = N_FEAT * K_FEAT
    w_deq = c_oracle.dequantize(p, a, n, "bfloat16").reshape(N_FEAT, K_FEAT)  # forward dequant
    out = qlora_oracle.qlora_linear_ref(x, w_deq, None, la, lb, 1.0, dy)
    c_oracle.dequantize(p, a, n, "bfloat16")  # backward dequant
    return out


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_cpu(T, steps, warmup):
    import torch

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = make_cpu_case(T)
    for _ in range(warmup):
        cpu_step(case, T)
    times = []
    for _ in range(steps):
        t0 = time.perf_counter()
        cpu_step(case, T)
        times.append(time.perf_counter() - t0)
    mean = sum(times) / len(times)
    return {"value": layer_flops(T) / mean / 1e12, "ms": mean * 1e3, "cores": torch.get_num_threads(),
            "cpu_model": cpu_model()}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    T = TOKENS  # the same configuration as the GPU arm: the full 4096-token batch per step (~1 s of host work each)
    steps = max(1, min(args.steps, 20))  # bounded: the whole run stays within a couple of minutes
    res = time_cpu(T, steps, max(1, min(args.warmup, 3)))
    line = {
        "impl": "reference", "metric": METRIC, "value": res["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": max(1, min(args.warmup, 3)), "ms_per_step": res["ms"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "single NF4 Linear 3072x3072 + LoRA r=16 fwd+bwd on 4096 tokens, bf16 (BASELINE configs[0])",
                   "N": N_FEAT, "K": K_FEAT, "r": RANK, "tokens_per_gpu": T, "tokens_per_step": T, "same_config": True,
                   "note": "CPU path: dequantize W -> bf16 (C, OpenMP), F.linear + LoRA, autograd backward, second dequantize"},
        "cpu_baseline": {"value": res["value"], "unit": UNIT, "cores": res["cores"], "kind": "port",
                         "cpu_model": res["cpu_model"],
                         "sample": f"all {TOKENS} tokens per step, {steps} timed steps; oracle port of the bitsandbytes+LoRA "
                                   "CPU path (bitsandbytes 0.48.2 is not installable offline)"},
        "e2e": {"value": res["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    REASONS = {
        0x1: "gpu_idle", 0x2: "applications_clocks_setting", 0x4: "sw_power_cap", 0x8: "hw_slowdown",
        0x10: "sync_boost", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake_slowdown",
        0x100: "display_clock_setting",
    }

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                self.samples.append(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM))
                mask = self.nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self.REASONS.items():
                    if mask & bit and name != "gpu_idle":
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread is not None:
            self._thread.join()
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


# ----------------------------------------------------------------------------- GPU arm
def build_layer(device):
    import torch
    import torch.nn as nn
    from src.modules.peft import LoRAConfig, PeftTargetConfig
    from src.modules.quant import quantize_inplace

    class Model(nn.Module):  # tests/test_modules_quant.py-style (SURVEY.md 8d cfg 1)
        def __init__(self):
            super().__init__()
            self.linear = nn.Linear(K_FEAT, N_FEAT, bias=False, dtype=torch.bfloat16)

    model = Model()
    with torch.no_grad():
        model.linear.weight.copy_((torch.randn(N_FEAT, K_FEAT, generator=torch.Generator().manual_seed(0)) * 0.02))
    quantize_inplace(model, "bnb_nf4", include_keys=["linear"])
    model.to(device)
    PeftTargetConfig(config=LoRAConfig(rank=RANK, alpha=1.0, dtype="bfloat16"), include_keys=["linear"]).replace_to_peft_layer(
        model, freeze_base=True
    )
    with torch.no_grad():
        model.linear.lora_up.weight.copy_(
            (torch.randn(N_FEAT, RANK, generator=torch.Generator().manual_seed(4)) * 0.02).to(device))
    return model


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist

    from vft_b200 import _cabi, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly ONE JSON line: everything libraries print while the run lasts (NCCL's version banner
    # goes to the stdout file descriptor from C) is sent to stderr; the line is written to the saved descriptor.
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the QLoRA hot path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime

        # stdout carries exactly one JSON line: keep NCCL's version banner (NCCL_DEBUG=VERSION) off it
        # The fused GEMM is a persistent kernel on 72 of the 74 SM pairs.  The LoRA-gradient exchange is small
        # (196 KB for this layer, 67 MB for a whole AuraFlow step) and latency-bound, so NCCL gets at most 2 CTAs: it
        # runs on the SMs the GEMM leaves free instead of holding back the next step's cluster launch.  Measured at
        # 8 GPUs (tools/dp_probe.py, profiles/r01_dp_probe_n8.json), us per step: no exchange 147.9; side stream with
        # NCCL's default channels 171.2, capped at 4 CTAs 157.2, at 2 CTAs 154.1; in stream order 182-190.
        pg_opts = None
        try:
            pg_opts = dist.ProcessGroupNCCL.Options()
            pg_opts.config.max_ctas = int(os.environ.get("VFT_NCCL_MAX_CTAS", "2"))
            pg_opts.config.min_ctas = 1
        except Exception:  # pragma: no cover - older torch: default channel count
            pg_opts = None
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120), pg_options=pg_opts)
    comm_group = None  # the default group: the 2-CTA communicator configured above
    peaks = measured_peaks()
    model = build_layer(dev)
    layer = model.linear
    params = [layer.lora_down.weight, layer.lora_up.weight]
    T = TOKENS

    # input sets: rotate so the footprint between two uses of a set exceeds L2 (126 MB)
    n_sets = 4
    gens = [torch.Generator(device=dev).manual_seed(1000 * rank + i) for i in range(n_sets)]
    xs = [torch.randn(2, T // 2, K_FEAT, generator=g, device=dev, dtype=torch.bfloat16).requires_grad_(True) for g in gens]
    dys = [torch.randn(2, T // 2, N_FEAT, generator=g, device=dev, dtype=torch.bfloat16) for g in gens]

    def step(i):
        x = xs[i % n_sets]
        x.grad = None
        for p in params:
            p.grad = None
        y = layer(x)
        y.backward(dys[i % n_sets])

    # eager warm-up (also compiles nothing: kernels are prebuilt)
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    assert ops.last_path() in (_cabi.PATH_TCGEN05, _cabi.PATH_SIMT)

    # CUDA graphs: one per input set, so launch overhead is off the device timeline
    graphs, launch_mode = [], "cuda_graph"
    round_graph = None
    grad_bufs = []
    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for i in range(n_sets):
                step(i)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        # N > 1, --exchange graph (default): the all-reduce of the PREVIOUS step's gradient bucket is captured inside the
        # step's graph on a forked stream, i.e. the exchange overlaps the next step's forward/backward and costs no host
        # time per step (an eager NCCL call + event bookkeeping per 150-us step made the host loop the pacing item at 8
        # ranks).  The buckets are allocated up front so that graph i can name the bucket graph i-1 fills.
        n_grad = sum(p.numel() for p in params)
        grad_bufs = [torch.zeros(n_grad, device=dev, dtype=torch.bfloat16) if world > 1 else None for _ in range(n_sets)]
        in_graph = world > 1 and args.exchange == "graph"
        fork = torch.cuda.Stream() if in_graph else None
        def captured_step(i):
            cur = torch.cuda.current_stream()
            if in_graph:
                fork.wait_stream(cur)
                with torch.cuda.stream(fork):
                    dist.all_reduce(grad_bufs[(i - 1) % n_sets], group=comm_group)
            step(i)
            if world > 1:  # the flat bucket exists for the NCCL exchange only: a single rank has nothing to pack
                torch.cat([p.grad.reshape(-1) for p in params], out=grad_bufs[i])
            if in_graph:
                cur.wait_stream(fork)

        for i in range(n_sets):
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                captured_step(i)
            graphs.append(g)
        # One graph launch per STEP puts a graph-launch boundary between every backward and the next forward: ~5 us per
        # 130-us step during which the device idles, and the programmatic dependent launch of the next kernel (its
        # barrier / TMEM / tensor-map set-up under the previous kernel's tail) is lost.  A training loop captures more
        # than one layer call per graph; here one graph holds a whole round of the n_sets input sets (same kernels, same
        # order, same work per step), and a remainder of steps % n_sets steps runs from the single-step graphs.
        if world == 1 or in_graph:
            round_graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(round_graph):
                for i in range(n_sets):
                    captured_step(i)
        torch.cuda.synchronize()
    except Exception as e:  # pragma: no cover - reported, not hidden
        print(f"[bench] CUDA graph capture failed ({type(e).__name__}: {e}); timing eager launches", file=sys.stderr)
        graphs, launch_mode, round_graph = [], "eager", None
        if args.exchange == "graph":
            args.exchange = "overlap"
        torch.cuda.synchronize()
    if round_graph is not None:
        launch_mode += f" ({n_sets} steps per graph launch)"
    if world > 1:
        launch_mode += f"+allreduce:{args.exchange}"

    comm = torch.cuda.Stream() if (world > 1 and args.exchange == "overlap") else None
    comm_events = [None] * n_sets

    def run_step(i, exchange=True):
        s = i % n_sets
        if comm is not None and comm_events[s] is not None:
            torch.cuda.current_stream().wait_event(comm_events[s])  # the set's gradients are about to be overwritten
        if graphs:
            graphs[s].replay()
            flat = grad_bufs[s]
        else:
            step(i)
            flat = torch.cat([p.grad.reshape(-1) for p in params]) if world > 1 else None
        if comm is not None and exchange:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(comm):
                comm.wait_event(ev)
                dist.all_reduce(flat)  # LoRA-gradient exchange, overlapped with the next step's kernels
                done = torch.cuda.Event()
                done.record()
            comm_events[s] = done
        elif world > 1 and exchange and args.exchange == "inorder":
            dist.all_reduce(flat)  # in stream order, right behind the step that produced the gradients

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # Everything with a host-side cost that differs between ranks happens BEFORE the barrier: nvmlInit() serialises the
    # ranks of a node on a driver lock (8 processes: milliseconds), and a rank that starts its timed region late makes
    # every other rank wait for it inside the per-step collective -- round 1's 8-GPU line measured exactly that skew.
    sampler = ClockSampler(local)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    align = torch.zeros(1, device=dev)
    def run_steps(n):
        i = 0
        while i < n:
            if round_graph is not None and i % n_sets == 0 and i + n_sets <= n:
                round_graph.replay()  # steps i .. i + n_sets - 1
                i += n_sets
            else:
                run_step(i)
                i += 1

    run_steps(max(args.warmup, 3))
    if comm is not None:
        torch.cuda.current_stream().wait_stream(comm)
    barrier()
    sampler.start()
    if world > 1:
        dist.all_reduce(align)  # ranks leave this collective together ON THE DEVICE, right in front of the first event
    e0.record()
    run_steps(args.steps)  # exactly args.steps steps
    if comm is not None:
        torch.cuda.current_stream().wait_stream(comm)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    # keep the GPU under the same load a little longer if the region was too short to sample clocks
    t_end = time.time() + 0.25
    while len(sampler.samples) < 8 and time.time() < t_end:
        step(0)  # rank-local padding (eager launches): the number of iterations differs per rank, so no collective
    torch.cuda.synchronize()
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    ms_per_step = ms / args.steps
    value = world * layer_flops(T) / (ms_per_step * 1e-3) / 1e12

    # ---- e2e: host (pinned) inputs, H2D every step, D2H of the adapter gradients every step
    hx = [torch.randn(2, T // 2, K_FEAT, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    hdy = [torch.randn(2, T // 2, N_FEAT, dtype=torch.bfloat16).pin_memory() for _ in range(2)]

    # The user-level loop a trainer runs with host-side batches: a copy stream uploads step i+1 (two device buffers)
    # while step i computes; the adapter gradients of every step are read back to pinned host memory and the host
    # waits for them one step behind (so it never idles the GPU), then once more at the end.
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    dx_in = [torch.empty(2, T // 2, K_FEAT, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ddy_in = [torch.empty(2, T // 2, N_FEAT, device=dev, dtype=torch.bfloat16) for _ in range(2)]
    ev_in = [torch.cuda.Event() for _ in range(2)]
    ev_free = [torch.cuda.Event() for _ in range(2)]
    ev_done = [torch.cuda.Event() for _ in range(2)]
    hga2 = [torch.empty(RANK, K_FEAT, dtype=torch.bfloat16).pin_memory() for _ in range(2)]
    hgb2 = [torch.empty(N_FEAT, RANK, dtype=torch.bfloat16).pin_memory() for _ in range(2)]

    def e2e_upload(i):
        b = i % 2
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(ev_free[b])  # the step that last used this buffer has consumed it
            dx_in[b].copy_(hx[b], non_blocking=True)
            ddy_in[b].copy_(hdy[b], non_blocking=True)
            ev_in[b].record(copy_stream)

    def e2e_compute(i):
        b = i % 2
        main_stream.wait_event(ev_in[b])
        x = dx_in[b].detach().requires_grad_(True)  # a fresh leaf over the uploaded buffer
        for p in params:
            p.grad = None
        layer(x).backward(ddy_in[b])
        ev_free[b].record(main_stream)
        if world > 1:
            flat = torch.cat([p.grad.reshape(-1) for p in params])
            dist.all_reduce(flat)
        hga2[b].copy_(params[0].grad, non_blocking=True)
        hgb2[b].copy_(params[1].grad, non_blocking=True)
        ev_done[b].record(main_stream)

    def e2e_run(n):
        for b in range(2):
            ev_free[b].record(main_stream)
        e2e_upload(0)
        for i in range(n):
            if i + 1 < n:
                e2e_upload(i + 1)
            e2e_compute(i)
            if i >= 1:
                ev_done[(i - 1) % 2].synchronize()  # the host consumes step i-1's gradients
        ev_done[(n - 1) % 2].synchronize()

    e2e_run(3)
    barrier()
    e2e_steps = max(5, min(args.steps, 50))
    t0 = time.perf_counter()
    e2e_run(e2e_steps)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps  # host-visible time per step
    barrier()
    if world > 1:
        t = torch.tensor([e2e_ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t.item())
    e2e_val = world * layer_flops(T) / (e2e_ms * 1e-3) / 1e12
    h2d = hx[0].numel() * 2 + hdy[0].numel() * 2
    d2h = hga2[0].numel() * 2 + hgb2[0].numel() * 2

    # ---- roofline: the two launches of the step, timed alone through the C ABI: forward with the adapter's side
    # product inside (vft_qlora_fwd, r = 16) and the whole backward in one launch (vft_qlora_bwd: dx, dt, dA, dB); the
    # NF4-only launches (r = 0) are timed next to them
    roof = None
    if rank == 0:
        w = layer.linear.weight
        qs = w.quant_state
        xk = [torch.randn(T, K_FEAT, device=dev, dtype=torch.bfloat16) for _ in range(n_sets)]
        gk = [torch.randn(T, N_FEAT, device=dev, dtype=torch.bfloat16) for _ in range(n_sets)]
        yk = torch.empty(T, N_FEAT, device=dev, dtype=torch.bfloat16)
        dxk = torch.empty(T, K_FEAT, device=dev, dtype=torch.bfloat16)
        st = torch.cuda.current_stream().cuda_stream
        absmax = qs.absmax_f32()
        tiles = ops.nf4_tile_weight(w.data, absmax, N_FEAT, K_FEAT)  # same derived copy the module path uses
        tc_ptr, ta_ptr = (tiles[0].data_ptr(), tiles[1].data_ptr()) if tiles else (None, None)
        la, lb = layer.lora_down.weight.detach(), layer.lora_up.weight.detach()
        sc = float(layer._scale_value())
        rp = 16 * ((RANK + 15) // 16)
        tsk = torch.zeros(T, 64, device=dev, dtype=torch.bfloat16)
        dtsk = torch.zeros(T, 64, device=dev, dtype=torch.bfloat16)
        btk = torch.empty(rp, N_FEAT, device=dev, dtype=torch.bfloat16)
        ttk = torch.empty(rp, T, device=dev, dtype=torch.bfloat16)
        dak, dbk = torch.empty_like(la), torch.empty_like(lb)
        wsb = _cabi.lib.vft_workspace_bytes(_cabi.OP_BWD, T, N_FEAT, K_FEAT, RANK)
        wsk = torch.empty(max(wsb, 4), dtype=torch.uint8, device=dev)

        def k_fwd0(i):
            _cabi.check(_cabi.lib.vft_qlora_fwd(xk[i % n_sets].data_ptr(), T, w.data_ptr(), absmax.data_ptr(), N_FEAT, K_FEAT, 64,
                                                _cabi.BF16, _cabi.BF16, None, None, None, 0, 0.0, yk.data_ptr(), None, None, None, None, 0, tc_ptr, ta_ptr, st))

        def k_bwd0(i):
            _cabi.check(_cabi.lib.vft_qlora_bwd_dx(gk[i % n_sets].data_ptr(), T, w.data_ptr(), absmax.data_ptr(), N_FEAT, K_FEAT, 64,
                                                   _cabi.BF16, _cabi.BF16, None, None, 0, 0.0, dxk.data_ptr(), None, None, None, 0, tc_ptr, ta_ptr, st))

        def k_fwd(i):
            _cabi.check(_cabi.lib.vft_qlora_fwd(xk[i % n_sets].data_ptr(), T, w.data_ptr(), absmax.data_ptr(), N_FEAT, K_FEAT, 64,
                                                _cabi.BF16, _cabi.BF16, None, la.data_ptr(), lb.data_ptr(), RANK, sc, yk.data_ptr(),
                                                tsk.data_ptr(), btk.data_ptr(), ttk.data_ptr(), None, 0, tc_ptr, ta_ptr, st))

        def k_bwd(i):
            _cabi.check(_cabi.lib.vft_qlora_bwd(gk[i % n_sets].data_ptr(), xk[i % n_sets].data_ptr(), T, w.data_ptr(), absmax.data_ptr(),
                                                N_FEAT, K_FEAT, 64, _cabi.BF16, _cabi.BF16, la.data_ptr(), lb.data_ptr(), RANK, sc,
                                                tsk.data_ptr(), ttk.data_ptr(), btk.data_ptr(), dxk.data_ptr(), dak.data_ptr(), dbk.data_ptr(),
                                                dtsk.data_ptr(), wsk.data_ptr(), wsb, tc_ptr, ta_ptr, st))

        def time_kernel(fn, iters=50):
            for i in range(5):
                fn(i)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for i in range(iters):
                fn(i)
            b.record()
            torch.cuda.synchronize()
            return a.elapsed_time(b) / iters

        t_f0, t_b0 = time_kernel(k_fwd0), time_kernel(k_bwd0)
        t_f, t_b = time_kernel(k_fwd), time_kernel(k_bwd)
        assert ops.last_path() == _cabi.PATH_TCGEN05, "roofline kernel is not the tcgen05 path"
        flops_launch = 2 * T * N_FEAT * K_FEAT
        lora_f, lora_b = 2 * T * RANK * (N_FEAT + K_FEAT), 4 * T * RANK * (N_FEAT + K_FEAT)
        achieved = (2 * flops_launch + lora_f + lora_b) / ((t_f + t_b) * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": UNIT,
                "frac": achieved / peaks["bf16_tflops"],
                # dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture of these two
                # launches (profiles/r02_ncu_tc2_fused.txt: forward 31.1 MB read + 0.9 MB written, backward 58.1 + 3.3 MB --
                # it also reads x for dA; the 25 MB of output stay in the 126 MB L2 past the end of the launch).
                # Algorithmic minimum: forward x + y + packed W = 55.9 MB, backward dy + x + dx + packed W = 81 MB.
                "traffic": 46.7e6, "traffic_source": "ncu dram bytes per launch (read + write, mean of the forward and the backward launch), profiles/r02_ncu_tc2_fused.txt",
                "peak_source": peaks["source"] + ", burst",
                "kernel": "qlora_tc2_kernel as launched in the step: forward <adapter side product inside>, backward <side product + dA/dB job inside> (persistent CTA-pair tcgen05 GEMM)",
                "fwd_us": t_f * 1e3, "bwd_us": t_b * 1e3, "flops_per_launch": flops_launch + (lora_f + lora_b) / 2,
                "nf4_only": {"fwd_us": t_f0 * 1e3, "bwd_us": t_b0 * 1e3, "tflops": 2 * flops_launch / ((t_f0 + t_b0) * 1e-3) / 1e12,
                             "frac": 2 * flops_launch / ((t_f0 + t_b0) * 1e-3) / 1e12 / peaks["bf16_tflops"]}}

    # ---- secondary figures of merit (same run, rank 0): quantize/pack GB/s, small-T weight-stream GB/s
    extra = {}
    if rank == 0:
        # NF4 quantize/pack (BASELINE configs[1] building block): the largest AuraFlow DiT weight [18432, 3072] bf16,
        # device-resident, outputs preallocated, C-ABI calls replayed from a CUDA graph (host overhead off the timeline)
        qn, qk = 18432, 3072
        wq2 = [(torch.randn(qn, qk, device=dev) * 0.02).to(torch.bfloat16) for _ in range(3)]
        n = qn * qk
        q_packed = torch.empty(n // 2, dtype=torch.uint8, device=dev)
        q_absmax = torch.empty(n // 64, dtype=torch.float32, device=dev)
        qside = torch.cuda.Stream()
        qgraph = torch.cuda.CUDAGraph()
        q_reps = 9
        with torch.cuda.stream(qside):
            qst = qside.cuda_stream
            for t in wq2:
                _cabi.check(_cabi.lib.vft_nf4_quantize(t.data_ptr(), _cabi.BF16, n, 64, q_packed.data_ptr(), q_absmax.data_ptr(), qst))
            qside.synchronize()
            with torch.cuda.graph(qgraph, stream=qside):
                for i in range(q_reps):
                    _cabi.check(_cabi.lib.vft_nf4_quantize(wq2[i % 3].data_ptr(), _cabi.BF16, n, 64, q_packed.data_ptr(),
                                                           q_absmax.data_ptr(), qst))
        qgraph.replay()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(5):
            qgraph.replay()
        b.record()
        torch.cuda.synchronize()
        q_ms = a.elapsed_time(b) / (5 * q_reps)
        gbs = (2 * n + n / 2 + n / 16) / (q_ms * 1e-3) / 1e9
        extra["nf4_quantize_pack"] = {"GB/s": gbs, "frac_of_hbm_peak": gbs / peaks["hbm_gbs"], "elements": n,
                                      "bytes_per_element": 2.5625, "us_per_tensor": q_ms * 1e3,
                                      "note": "bf16 [18432, 3072], device-resident, CUDA-graph replay"}
        del wq2, qgraph
        try:  # BASELINE configs[1]: the whole synthetic AuraFlow DiT weight set (322 tensors, 6.80 G elements, fp16)
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import quant_probe

            q_ms_1, q_gbs_1 = quant_probe.whole_set(torch.float16)
            q_ms_set, q_gbs_set, q_k = quant_probe.whole_set_batched(torch.float16)
            extra["nf4_quantize_pack_auraflow_set"] = {
                "ms": q_ms_set, "GB/s": q_gbs_set, "frac_of_hbm_peak": q_gbs_set / peaks["hbm_gbs"], "tensors": q_k,
                "elements": 6.80e9, "hbm_floor_ms": 17.4e9 / peaks["hbm_gbs"] / 1e6, "launches": 6,
                "one_launch_per_tensor": {"ms": q_ms_1, "GB/s": q_gbs_1, "frac_of_hbm_peak": q_gbs_1 / peaks["hbm_gbs"]},
                "note": "fp16 weights device-resident (13.6 GB), vft_nf4_quantize_many: up to 96 equal-size tensors per launch (6 launches for this set), CUDA events around the call"}
        except Exception as e:  # pragma: no cover - reported, not hidden
            extra["nf4_quantize_pack_auraflow_set"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    if rank == 0:
        # few-token forward (T = per-GPU batch 2) of the largest AuraFlow modulation weight [18432, 3072]: achieved
        # GB/s on the packed-weight stream (0.5625 B/parameter), HBM-cold (9 weight copies in rotation > 2 x L2)
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import gemv_probe

            r = gemv_probe.probe(18432, 3072, 2, 0)
            extra["few_token_weight_stream"] = {
                "GB/s": r["weight_GBs"], "frac_of_hbm_peak": r["weight_GBs"] / peaks["hbm_gbs"], "us_per_launch": r["us"],
                "N": 18432, "K": 3072, "T": 2, "path": r["path"], "bytes_per_param": 0.5625,
                "note": "qlora_gemv.cu (path 3); HBM-cold: weight copies rotate; tools/gemv_probe.py sweeps T and shapes"}
        except Exception as e:  # pragma: no cover - reported, not hidden
            extra["few_token_weight_stream"] = {"error": f"{type(e).__name__}: {e}"}

    # (the censuses run BEFORE the step harnesses: ten seconds of whole-model steps leave the GPU heat-soaked and under its
    #  power cap, and the per-layer timings that follow then read 3-10 % slower than the same library measured by
    #  tools/census.py alone -- profiles/README.md)
    if rank == 0 and world == 1 and not args.no_census:
        # second half of BASELINE's metric: every NF4(+LoRA) Linear of one AuraFlow-6.8B QLoRA training step (per-GPU
        # batch 2 at 1024^2, LoRA r=16 on attention + MLP projections, gradient checkpointing = forward twice), each at
        # its own token count, timed through the C ABI (tools/census.py).  Attention, norms and the optimizer are not
        # part of the hot path and are not included: this bounds steps/s from above.
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import census

            ms, avg_tf, _ = census.model_step(census.auraflow(2), dev)
            extra["auraflow_qlora_step_linear_layers"] = {
                "ms_per_step": ms, "steps_per_s_upper_bound": 1e3 / ms, "avg_tflops": avg_tf,
                "frac_of_measured_bf16_peak": avg_tf / peaks["bf16_tflops"], "per_gpu_batch": 2,
                "note": "322 quantized Linears, forward x2 (checkpointing) + backward; data-parallel ranks step independently"}
            # BASELINE configs[2] and [4]: the same census for Lumina2 NextDiT-2.6B (batch 1, 1024^2) and the SDXL UNet
            # transformer Linears (batch 2, 1024^2: small K, 77-token cross-attention projections)
            for key, layers, what in (("lumina2_qlora_step_linear_layers", census.lumina2(1), "178 quantized Linears of NextDiT-2.6B, batch 1"),
                                      ("sdxl_qlora_step_linear_layers", census.sdxl(2), "SDXL UNet attention/FF Linears (C = 640, 1280), batch 2")):
                ms, avg_tf, _ = census.model_step(layers, dev)
                extra[key] = {"ms_per_step": ms, "avg_tflops": avg_tf, "frac_of_measured_bf16_peak": avg_tf / peaks["bf16_tflops"],
                              "note": what + "; forward x2 (checkpointing) + backward, each layer at its own token count"}
        except Exception as e:  # pragma: no cover - reported, not hidden
            extra["auraflow_qlora_step_linear_layers"] = {"error": f"{type(e).__name__}: {e}"}
        try:  # sibling projections as one launch (vft_b200/group.py), module API, forward + backward
            import group_probe

            rows = [group_probe.time_case(nm, k, ns, t, r, verbose=False) for nm, k, ns, t, r in (
                ("sdxl C1280 attn1 to_q/k/v", 1280, [1280] * 3, 2048, 4), ("sdxl C1280 attn1 to_q/k/v", 1280, [1280] * 3, 2048, 16),
                ("sdxl C640 attn1 to_q/k/v", 640, [640] * 3, 8192, 4), ("sdxl C1280 attn2 to_k/v (2 x 77 text tokens)", 2048, [1280] * 2, 154, 4))]
            extra["projection_groups"] = {
                "cases": [{**r_, "speedup": r_["members_us"] / r_["group_us"]} for r_ in rows],
                "note": "q/k/v (k/v) LoRALinear-over-Linear4bit siblings, forward + backward through the module API in a CUDA graph: "
                        "one launch per member vs one ProjectionGroup launch per direction; LoRA rank 4 is the reference's shipped rank"}
        except Exception as e:  # pragma: no cover - reported, not hidden
            extra["projection_groups"] = {"error": f"{type(e).__name__}: {e}"}

    if not args.no_aura_step:
        # second half of BASELINE's metric, measured as a job: the AuraFlow-6.8B QLoRA step over the Linear skeleton of
        # the MMDiT (tools/auraflow_step.py: module API, 322 NF4 Linears, LoRA r=16, checkpointing, fused AdamW,
        # element-wise glue in torch, attention stand-in), every rank its own batch of 2, LoRA gradients all-reduced
        # over NCCL from backward hooks.  All ranks take part; timings are the max over ranks.
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import auraflow_step

            # the 67 MB exchange of a whole step is bandwidth-bound: its own process group, NCCL's default channels
            aura_group = dist.new_group(pg_options=dist.ProcessGroupNCCL.Options()) if world > 1 else None
            res = auraflow_step.measure(B=2, steps=4, warmup=2, attention="stub", world=world, rank=rank, group=aura_group)
            if rank == 0:
                extra["auraflow_qlora_step_dp"] = res
            if world == 1:
                # the same step with torch's SDPA (a library call outside the hot path) in place of the attention stand-in:
                # the figure a user of the whole model would see per GPU
                torch.cuda.empty_cache()
                r2 = auraflow_step.measure(B=2, steps=3, warmup=1, attention="sdpa", world=1, rank=0, group=None)
                extra["auraflow_qlora_step_sdpa"] = {k: r2[k] for k in ("workload", "launch", "ms_per_step", "steps_per_s", "samples_per_s",
                                                                        "hot_path_tflops_per_gpu", "hot_path_frac_of_sustained_bf16_peak", "mem_gb") if k in r2}
        except Exception as e:  # pragma: no cover - reported, not hidden
            if rank == 0:
                extra["auraflow_qlora_step_dp"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
        if world == 1:
            # BASELINE configs[2]: the Lumina2 NextDiT-2.6B QLoRA step (batch 1 at 1024^2) as a job, same kind of harness
            # (tools/lumina2_step.py: 178 NF4 Linears, LoRA r=16, checkpointing, fused AdamW, whole step as one CUDA graph)
            try:
                import lumina2_step

                torch.cuda.empty_cache()
                extra["lumina2_qlora_step"] = {att: {k: v for k, v in lumina2_step.measure(1, 4, 2, att).items() if k not in ("adapter_params", "loss")}
                                               for att in ("stub", "sdpa")}
            except Exception as e:  # pragma: no cover - reported, not hidden
                extra["lumina2_qlora_step"] = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            r = time_cpu(TOKENS, 5, 1)  # ~10 s of host work on the box's cores, the full 4096-token batch per run
            cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "cpu_model": r["cpu_model"],
                   "sample": f"all {TOKENS} tokens, 1 warm-up + 5 runs of the oracle port (C/OpenMP dequant + F.linear + LoRA, "
                             "autograd, second dequant)"}
        # qlora_tc2<fwd, side product inside> | qlora_tc2<bwd, side product + dA/dB job inside>; the triage switches
        # bring the side kernels back: lora_side<x.A^T>, <dy.B> (VFT_TC2_FUSE=0) and lora_side<dA,dB> (VFT_TC2_JOB=0)
        kernels_per_step = 5 if os.environ.get("VFT_TC2_FUSE") == "0" else (3 if os.environ.get("VFT_TC2_JOB") == "0" else 2)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "single NF4 Linear 3072x3072 + LoRA r=16 fwd+bwd on 4096 tokens per GPU, bf16 (BASELINE configs[0])",
                       "N": N_FEAT, "K": K_FEAT, "r": RANK, "tokens_per_gpu": T, "parallelism": f"dp{world}",
                       "launch": launch_mode,
                       "l2": f"inputs rotate over {n_sets} sets (x,dy,y,dx = {4 * T * 3072 * 2 * n_sets / 1e6:.0f} MB > 126 MB L2)",
                       "flops_per_step_per_gpu": layer_flops(T),
                       "frac_of_measured_bf16_peak": value / world / peaks["bf16_tflops"]},
            "roofline": roof, "cpu_baseline": cpu,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms},
            "gpu_launches": kernels_per_step * args.steps, "clocks": clocks, "extra": extra,
        }
        sys.stdout.flush()
        os.write(json_fd, (json.dumps(line) + "\n").encode())
    if world > 1:
        dist.barrier()
        torch.cuda.synchronize()
        # Tear-down: the captured step graphs hold NCCL work; destroying the process group under them was seen to block
        # for ever (torch 2.11 / NCCL 2.28).  Drop the graphs first, give destroy_process_group() ten seconds in a helper
        # thread, and leave with exit code 0 either way -- the JSON line is already on stdout.
        import gc

        graphs.clear()
        round_graph = None
        gc.collect()
        t = threading.Thread(target=dist.destroy_process_group, daemon=True)
        t.start()
        t.join(10.0)
        sys.stderr.flush()
        os._exit(0)


def main():
    if os.environ.get("BENCH_WATCHDOG"):  # triage: dump every thread's stack and exit if the run takes longer than this
        import faulthandler

        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-census", action="store_true")
    ap.add_argument("--no-aura-step", action="store_true")
    ap.add_argument("--exchange", default="graph", choices=["inorder", "overlap", "graph"],
                    help="N > 1: LoRA-gradient all-reduce in stream order behind each step, on a side stream (eager NCCL "
                         "call per step), or captured in the next step's CUDA graph on a forked stream (default)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
