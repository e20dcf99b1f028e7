"""AuraFlow-6.8B QLoRA training step on the hot path (BASELINE.json config #4), 1..8 GPUs data-parallel.

What runs: the Linear skeleton of the AuraFlow MMDiT (/root/reference/src/models/auraflow/denoiser.py:107-109,160-163,
233-242,351-362,442-445,561-569: 4 double + 32 single blocks, 322 NF4 Linears, 6.80 G parameters) built through the
reference-facing module API -- ``quantize_inplace`` + ``PeftTargetConfig.replace_to_peft_layer`` with the include keys
of /root/reference/tests/assets/debug_dataset.yml:16-19 (attention + MLP projections, r = 16; SURVEY.md 8e case ii) --
with per-block gradient checkpointing (denoiser.py:826-853), a fused AdamW step on the 33.6 M adapter parameters and,
for N > 1, ``vft_b200.dp.LoraGradReducer``: bucketed NCCL all-reduce of the LoRA gradients launched from backward
hooks on a side stream.

What does NOT run: the model code around the Linears is out of this repo's scope, so layer norms / modulation / gates
/ SwiGLU products are plain torch element-wise ops and joint attention is either torch SDPA (``--attention sdpa``, a
library call, reported separately) or an element-wise stand-in (``--attention stub``, default: the step is then
"every NF4(+LoRA) Linear of the step at its real token count + glue").  Per-GPU batch 2 at 1024^2: 4096 patch + 264
condition tokens per sample (configs/auraflow/lora.yml:29).

Reports steps/s, samples/s, the hot-path FLOP/s fraction and (N > 1) the exposed all-reduce time = step time minus
the step time with the exchange switched off.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "vision-ft_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.utils.checkpoint import checkpoint

D, HID, HEADS, JOINT, PATCH_IN = 3072, 8192, 12, 2048, 16
N_PATCH, N_TEXT, N_REG = 4096, 256, 8
R = 16


def _lin(k, n, bias=False):
    return nn.Linear(k, n, bias=bias, dtype=torch.bfloat16, device="cuda")


class MLP(nn.Module):
    def __init__(self):
        super().__init__()
        self.c_fc1, self.c_fc2, self.c_proj = _lin(D, HID), _lin(D, HID), _lin(HID, D)

    def forward(self, x):
        return self.c_proj(F.silu(self.c_fc1(x)) * self.c_fc2(x))


def _attend(q, k, v, mode):
    if mode == "sdpa":
        B, T, _ = q.shape
        sp = lambda t: t.view(B, T, HEADS, D // HEADS).transpose(1, 2)
        o = F.scaled_dot_product_attention(sp(q), sp(k), sp(v))
        return o.transpose(1, 2).reshape(B, T, D)
    return q * torch.sigmoid(k) + v  # element-wise stand-in: keeps the q/k/v/o Linears and their gradients live


def _mod(x, shift, scale):
    return F.layer_norm(x, (D,)) * (1 + scale[:, None]) + shift[:, None]


class SingleAttn(nn.Module):
    def __init__(self):
        super().__init__()
        self.w1q, self.w1k, self.w1v, self.w1o = _lin(D, D), _lin(D, D), _lin(D, D), _lin(D, D)

    def forward(self, c, mode):
        return self.w1o(_attend(self.w1q(c), self.w1k(c), self.w1v(c), mode))


class DoubleAttn(nn.Module):
    def __init__(self):
        super().__init__()
        for nm in ("w1q", "w1k", "w1v", "w1o", "w2q", "w2k", "w2v", "w2o"):
            setattr(self, nm, _lin(D, D))

    def forward(self, c, x, mode):
        nc = c.shape[1]
        q = torch.cat([self.w1q(c), self.w2q(x)], 1)
        k = torch.cat([self.w1k(c), self.w2k(x)], 1)
        v = torch.cat([self.w1v(c), self.w2v(x)], 1)
        o = _attend(q, k, v, mode)
        return self.w1o(o[:, :nc]), self.w2o(o[:, nc:])


class DoubleBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.mlpC, self.mlpX, self.attn = MLP(), MLP(), DoubleAttn()
        self.modC = nn.Sequential(nn.SiLU(), _lin(D, 6 * D))
        self.modX = nn.Sequential(nn.SiLU(), _lin(D, 6 * D))

    def forward(self, c, x, g, mode):
        cs = self.modC(g).chunk(6, dim=1)
        xs = self.modX(g).chunk(6, dim=1)
        ca, xa = self.attn(_mod(c, cs[0], cs[1]), _mod(x, xs[0], xs[1]), mode)
        c = c + cs[2][:, None] * ca
        x = x + xs[2][:, None] * xa
        c = c + cs[5][:, None] * self.mlpC(_mod(c, cs[3], cs[4]))
        x = x + xs[5][:, None] * self.mlpX(_mod(x, xs[3], xs[4]))
        return c, x


class SingleBlock(nn.Module):
    def __init__(self):
        super().__init__()
        self.modCX = nn.Sequential(nn.SiLU(), _lin(D, 6 * D))
        self.attn, self.mlp = SingleAttn(), MLP()

    def forward(self, c, g, mode):
        s = self.modCX(g).chunk(6, dim=1)
        c = c + s[2][:, None] * self.attn(_mod(c, s[0], s[1]), mode)
        return c + s[5][:, None] * self.mlp(_mod(c, s[3], s[4]))


class Denoiser(nn.Module):
    def __init__(self, n_double=4, n_single=32):
        super().__init__()
        self.cond_seq_linear = _lin(JOINT, D)
        self.init_x_linear = _lin(PATCH_IN, D, bias=True)
        self.register_tokens = nn.Parameter(torch.randn(1, N_REG, D, dtype=torch.bfloat16, device="cuda") * 0.02, requires_grad=False)
        self.double_layers = nn.ModuleList([DoubleBlock() for _ in range(n_double)])
        self.single_layers = nn.ModuleList([SingleBlock() for _ in range(n_single)])
        self.final_linear = _lin(D, PATCH_IN)

    def forward(self, patches, text, g, mode="stub", ckpt=True):
        x = self.init_x_linear(patches)
        c = torch.cat([self.register_tokens.expand(text.shape[0], -1, -1), self.cond_seq_linear(text)], 1)
        # no dropout anywhere: no RNG state to preserve (and reading it is not allowed during CUDA-graph capture)
        run = (lambda f, *a: checkpoint(f, *a, use_reentrant=False, preserve_rng_state=False)) if ckpt else (lambda f, *a: f(*a))
        for blk in self.double_layers:
            c, x = run(blk, c, x, g, mode)
        ctx = torch.cat([c, x], 1)
        for blk in self.single_layers:
            ctx = run(blk, ctx, g, mode)
        return self.final_linear(ctx[:, c.shape[1]:])


class Model(nn.Module):
    def __init__(self, n_double=4, n_single=32):
        super().__init__()
        self.denoiser = Denoiser(n_double, n_single)


def build(n_double=4, n_single=32, seed=0):
    from src.modules.peft import LoRAConfig, PeftTargetConfig
    from src.modules.quant import quantize_inplace
    from src.utils.state_dict import RegexMatch

    torch.manual_seed(seed)
    model = Model(n_double, n_single)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.02)
    # tools/quantize_model.py keys (:20-21); the fp weights already sit on the device, .to() packs them in place
    quantize_inplace(model, "bnb_nf4", include_keys=["denoiser."], exclude_keys=["t_embedder", "final_linear", "modF"])
    model.to("cuda")
    torch.cuda.empty_cache()
    PeftTargetConfig(
        config=LoRAConfig(rank=R, alpha=1.0, dtype="bfloat16"),
        include_keys=[RegexMatch(regex=r".*\.attn\.w2[qkvo]"), RegexMatch(regex=r".*\.mlp[X]?\."),
                      RegexMatch(regex=r".*single_layers\.\d+\.attn\.w1[qkvo]")],
        exclude_keys=["text_encoder", "vae", "t_embedder", "final_linear", RegexMatch(regex=r".*\.mod[CX]{1,2}")],
    ).replace_to_peft_layer(model, freeze_base=True)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("lora_up.weight"):
                p.normal_(0, 0.02)
    return model


def hot_path_flops(model, B):
    """4TNK + 6Tr(N+K) per NF4(+LoRA) Linear, + one more forward (2TNK + 2Tr(N+K)) for checkpointing."""
    from src.modules.peft import LoRALinear
    from vft_b200.nn import Linear4bit

    tok = {"double": (B * (N_TEXT + N_REG), B * N_PATCH), "single": B * (N_PATCH + N_TEXT + N_REG)}
    total = 0
    for name, m in model.named_modules():
        base = m.linear if isinstance(m, LoRALinear) else m
        if not isinstance(base, Linear4bit) or (name.endswith(".linear") and not isinstance(m, LoRALinear)):
            continue
        r = R if isinstance(m, LoRALinear) else 0
        n, k = base.out_features, base.in_features
        if ".mod" in name:
            T = B
        elif "cond_seq_linear" in name:
            T = B * N_TEXT
        elif "init_x_linear" in name:
            T = B * N_PATCH
        elif "double_layers" in name:
            T = tok["double"][0] if (".w1" in name or ".mlpC." in name) else tok["double"][1]
        else:
            T = tok["single"]
        ck = 0 if ("cond_seq_linear" in name or "init_x_linear" in name) else 1
        total += 4 * T * n * k + 6 * T * r * (n + k) + ck * (2 * T * n * k + 2 * T * r * (n + k))
    return total


def run(args):
    import torch.distributed as dist

    from vft_b200.dp import LoraGradReducer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    own_pg = False
    if world > 1 and not dist.is_initialized():
        import datetime

        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=180))
        own_pg = True
    res = measure(args.batch, args.steps, args.warmup, args.attention, args.double, args.single, world, rank,
                  overlap=args.overlap, graph=not args.eager)
    if rank == 0:
        print(json.dumps(res), flush=True)
    if own_pg:
        dist.barrier()
        dist.destroy_process_group()


def measure(B=2, steps=5, warmup=2, attention="stub", n_double=4, n_single=32, world=1, rank=0, exposed=True,
            overlap=False, group=None, graph=True):
    """One process per GPU (the caller has set the device and, for world > 1, initialised NCCL).  Returns a dict on
    every rank (timings are the max over ranks)."""
    import torch.distributed as dist

    from vft_b200.dp import LoraGradReducer

    dev = torch.device("cuda", torch.cuda.current_device())
    model = build(n_double, n_single)
    params = [p for p in model.parameters() if p.requires_grad]
    n_adapter = sum(p.numel() for p in params)
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True, capturable=graph)
    reducer = LoraGradReducer(params, bucket_bytes=8 << 20, overlap=overlap, group=group) if world > 1 else None
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    patches = torch.randn(B, N_PATCH, PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)
    text = torch.randn(B, N_TEXT, JOINT, generator=g, device=dev, dtype=torch.bfloat16)
    gc = torch.randn(B, D, generator=g, device=dev, dtype=torch.bfloat16)
    target = torch.randn(B, N_PATCH, PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)

    def step(exchange=True):
        opt.zero_grad(set_to_none=True)
        out = model.denoiser(patches, text, gc, attention)
        loss = F.mse_loss(out.float(), target.float())
        if reducer is not None and not exchange:
            with reducer.no_sync():
                loss.backward()
        else:
            loss.backward()
            if reducer is not None:
                reducer.wait()
        opt.step()
        return loss

    # The whole step (forward, checkpointed backward, exchange, optimizer) as ONE CUDA graph: ~5000 launches per step
    # leave the host (tools/host_overhead.py: 40-110 us of Python per fused-layer call, which a B=1 step cannot hide).
    graphs, launch_mode = {}, "eager"

    def capture(exchange):
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            step(exchange)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr, capture_error_mode="thread_local"):
            out_loss = step(exchange)
        return gr, out_loss

    def run_step(exchange=True):
        if exchange in graphs:
            graphs[exchange][0].replay()
            return graphs[exchange][1]
        return step(exchange)

    def timed(n, exchange=True):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            loss = run_step(exchange)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), float(loss.item())

    for _ in range(max(warmup, 1)):
        step()
    if graph:
        try:
            graphs[True] = capture(True)
            if world > 1 and exposed:
                graphs[False] = capture(False)
            launch_mode = "cuda_graph"
        except Exception as e:  # reported, not hidden: the eager numbers are still valid
            print(f"[auraflow_step] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            graphs.clear()
            torch.cuda.synchronize()
    run_step()
    ms, loss = timed(steps)
    ms_local = None
    if world > 1 and exposed:
        # the first timed run of a process is a few ms slower than the following ones (allocator, clocks): time the
        # exchange run on both sides of the no-exchange run and keep the faster one
        run_step(False)
        ms_local, _ = timed(steps, exchange=False)
        run_step()
        ms2, loss = timed(steps)
        ms = min(ms, ms2)
    flops = hot_path_flops(model, B)
    peak = 1691.7
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    res = {
        "workload": f"AuraFlow-6.8B QLoRA step, Linear skeleton ({n_double} double + {n_single} single blocks), per-GPU batch {B} at 1024^2, "
                    f"LoRA r={R} on attention+MLP projections, gradient checkpointing, fused AdamW, attention={attention}",
        "n_gpus": world, "launch": launch_mode, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "samples_per_s": world * B * 1e3 / ms,
        "hot_path_tflops_per_gpu": flops / (ms * 1e-3) / 1e12,
        "hot_path_frac_of_sustained_bf16_peak": flops / (ms * 1e-3) / 1e12 / peak,
        "hot_path_flops_per_step_per_gpu": flops, "adapter_params": n_adapter,
        "allreduce_bytes_per_step": 2 * n_adapter if world > 1 else 0,
        "exchange": ("overlapped with backward (hooks)" if overlap else "after backward (deferred buckets)") if world > 1 else None,
        "ms_per_step_without_exchange": ms_local,
        "exposed_allreduce_ms": (ms - ms_local) if ms_local is not None else None,
        "loss": loss, "mem_gb": torch.cuda.max_memory_allocated() / 1e9,
    }
    if reducer is not None:
        reducer.remove()
    del model, opt
    torch.cuda.empty_cache()
    return res


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=2)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--attention", default="stub", choices=["stub", "sdpa"])
    ap.add_argument("--double", type=int, default=4)
    ap.add_argument("--single", type=int, default=32)
    ap.add_argument("--overlap", action="store_true")
    ap.add_argument("--eager", action="store_true", help="no CUDA graph: every launch goes through Python")
    run(ap.parse_args())
