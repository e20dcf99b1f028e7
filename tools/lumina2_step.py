"""Lumina Image 2.0 (NextDiT 2.6B) QLoRA training step on the hot path (BASELINE.json config #3), one B200.

Same idea as tools/auraflow_step.py: the Linear skeleton of the reference model
(/root/reference/src/models/lumina2/denoiser.py:85-95 qkv/out, :190-205 w1/w2/w3, :258-262 adaLN_modulation,
:283-339 block forward, :536-561 refiners; /root/reference/src/models/lumina2/config.py:9-30: hidden 2304, 24 heads +
8 KV heads of 96, 26 blocks + 2 noise-refiner + 2 context-refiner blocks) built through the reference-facing module API
-- ``quantize_inplace`` on every ``attention`` / ``feed_forward`` / ``adaLN_modulation`` Linear, LoRA r = 16 on
``attention`` + ``feed_forward`` except the noise refiner (/root/reference/configs/lumina2/text_to_image/lora.yml:23-27)
-- with per-block gradient checkpointing and a fused AdamW step.  1024^2 bucket, batch 1: 4096 image tokens + 256
caption tokens.  Norms / modulation / gates / SwiGLU products are torch element-wise ops, attention is an element-wise
stand-in (``--attention stub``) or torch SDPA with grouped KV heads (``--attention sdpa``): the model code around the
Linears is outside this repository's scope.  Prints one JSON object.
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

D, HEADS, KV_HEADS, HEAD_DIM, FF, T_EMB = 2304, 24, 8, 96, 9216, 1024
N_IMG, N_CAP, PATCH_IN = 4096, 256, 64
R = 16


def _lin(k, n, bias=False):
    return nn.Linear(k, n, bias=bias, dtype=torch.bfloat16, device="cuda")


class Attention(nn.Module):
    def __init__(self):
        super().__init__()
        self.qkv = _lin(D, (HEADS + 2 * KV_HEADS) * HEAD_DIM)
        self.out = _lin(HEADS * HEAD_DIM, D)

    def forward(self, x, mode):
        B, T, _ = x.shape
        q, k, v = self.qkv(x).split([HEADS * HEAD_DIM, KV_HEADS * HEAD_DIM, KV_HEADS * HEAD_DIM], dim=-1)
        if mode == "sdpa":
            q = q.view(B, T, HEADS, HEAD_DIM).transpose(1, 2)
            k = k.view(B, T, KV_HEADS, HEAD_DIM).transpose(1, 2)
            v = v.view(B, T, KV_HEADS, HEAD_DIM).transpose(1, 2)
            o = F.scaled_dot_product_attention(q, k, v, enable_gqa=True).transpose(1, 2).reshape(B, T, HEADS * HEAD_DIM)
        else:  # element-wise stand-in that keeps q, k, v and their gradients live
            rep = HEADS // KV_HEADS
            o = q * torch.sigmoid(k.repeat(1, 1, rep)) + v.repeat(1, 1, rep)
        return self.out(o)


class FeedForward(nn.Module):
    def __init__(self):
        super().__init__()
        self.w1, self.w2, self.w3 = _lin(D, FF), _lin(FF, D), _lin(D, FF)

    def forward(self, x):
        return self.w2(F.silu(self.w1(x)) * self.w3(x))


class Block(nn.Module):
    def __init__(self, modulation=True):
        super().__init__()
        self.attention, self.feed_forward = Attention(), FeedForward()
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), _lin(T_EMB, 4 * D, bias=True)) if modulation else None

    @staticmethod
    def _norm(x):
        return F.rms_norm(x, (D,))

    def forward(self, x, c, mode):
        if self.adaLN_modulation is not None:
            s_a, g_a, s_m, g_m = self.adaLN_modulation(c).chunk(4, dim=1)
            a = self._norm(self.attention(self._norm(x) * (1 + s_a[:, None]), mode))
            x = x + g_a[:, None].tanh() * a
            m = self._norm(self.feed_forward(self._norm(x) * (1 + s_m[:, None])))
            return x + g_m[:, None].tanh() * m
        x = x + self._norm(self.attention(self._norm(x), mode))
        return x + self._norm(self.feed_forward(self._norm(x)))


class Denoiser(nn.Module):
    def __init__(self, depth=26, refiner_depth=2):
        super().__init__()
        self.x_embedder = _lin(PATCH_IN, D, bias=True)
        self.noise_refiner = nn.ModuleList([Block(True) for _ in range(refiner_depth)])
        self.context_refiner = nn.ModuleList([Block(False) for _ in range(refiner_depth)])
        self.layers = nn.ModuleList([Block(True) for _ in range(depth)])
        self.final_linear = _lin(D, PATCH_IN)

    def forward(self, patches, caption, c, mode="stub"):
        run = lambda f, *a: checkpoint(f, *a, use_reentrant=False, preserve_rng_state=False)  # no dropout; capture-safe
        cap = caption
        for blk in self.context_refiner:
            cap = run(blk, cap, c, mode)
        x = self.x_embedder(patches)
        for blk in self.noise_refiner:
            x = run(blk, x, c, mode)
        h = torch.cat([cap, x], 1)
        for blk in self.layers:
            h = run(blk, h, c, mode)
        return self.final_linear(h[:, cap.shape[1]:])


class Model(nn.Module):
    def __init__(self, depth=26, refiner_depth=2):
        super().__init__()
        self.denoiser = Denoiser(depth, refiner_depth)


def build(depth=26, refiner_depth=2):
    from src.modules.peft import LoRAConfig, PeftTargetConfig
    from src.modules.quant import quantize_inplace

    torch.manual_seed(0)
    model = Model(depth, refiner_depth)
    with torch.no_grad():
        for p in model.parameters():
            if p.dim() == 2:
                p.normal_(0, 0.02)
    quantize_inplace(model, "bnb_nf4", include_keys=["attention", "feed_forward", "adaLN_modulation"])
    model.to("cuda")
    torch.cuda.empty_cache()
    PeftTargetConfig(config=LoRAConfig(rank=R, alpha=1.0, dtype="bfloat16"), include_keys=["attention", "feed_forward"],
                     exclude_keys=["text_encoder", "vae", "noise_refiner"]).replace_to_peft_layer(model, freeze_base=True)
    with torch.no_grad():
        for n, p in model.named_parameters():
            if n.endswith("lora_up.weight"):
                p.normal_(0, 0.02)
    return model


def hot_path_flops(model, B):
    """4TNK + 6Tr(N+K) per NF4(+LoRA) Linear, + one more forward for checkpointing."""
    from src.modules.peft import LoRALinear
    from vft_b200.nn import Linear4bit

    total = 0
    for name, m in model.named_modules():
        base = m.linear if isinstance(m, LoRALinear) else m
        if not isinstance(base, Linear4bit) or (name.endswith(".linear") and not isinstance(m, LoRALinear)):
            continue
        r = R if isinstance(m, LoRALinear) else 0
        n, k = base.out_features, base.in_features
        T = B if "adaLN_modulation" in name else B * (N_CAP if "context_refiner" in name else N_IMG if "noise_refiner" in name else N_IMG + N_CAP)
        total += 6 * T * n * k + 8 * T * r * (n + k)
    return total


def measure(batch=1, steps=5, warmup=2, attention="stub", eager=False):
    """One Lumina2 QLoRA step on the current CUDA device; returns the result dict (bench.py's extra, main())."""
    args = argparse.Namespace(batch=batch, steps=steps, warmup=warmup, attention=attention, eager=eager)
    dev = torch.device("cuda", torch.cuda.current_device())
    model = build()
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-4, fused=True, capturable=not args.eager)
    g = torch.Generator(device=dev).manual_seed(1)
    B = args.batch
    patches = torch.randn(B, N_IMG, PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)
    caption = torch.randn(B, N_CAP, D, generator=g, device=dev, dtype=torch.bfloat16)
    c = torch.randn(B, T_EMB, generator=g, device=dev, dtype=torch.bfloat16)
    target = torch.randn(B, N_IMG, PATCH_IN, generator=g, device=dev, dtype=torch.bfloat16)

    def step():
        opt.zero_grad(set_to_none=True)
        loss = F.mse_loss(model.denoiser(patches, caption, c, args.attention).float(), target.float())
        loss.backward()
        opt.step()
        return loss

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launch_mode, run = "eager", step
    if not args.eager:  # the whole step as one CUDA graph (see tools/auraflow_step.py)
        try:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                step()
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                g_loss = step()

            def run():
                gr.replay()
                return g_loss

            launch_mode = "cuda_graph"
        except Exception as e:
            print(f"[lumina2_step] CUDA graph capture failed ({type(e).__name__}: {e}); eager launches", file=sys.stderr)
            torch.cuda.synchronize()
            run = step
    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    flops = hot_path_flops(model, B)
    peak = 1415.6
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"]
    except Exception:
        pass
    n_q = sum(1 for m in model.modules() if type(m).__name__ == "BnbLinear4bit")
    return {
        "workload": f"Lumina2 NextDiT-2.6B QLoRA step, Linear skeleton (26 + 2 + 2 blocks, {n_q} NF4 Linears), batch {B} at 1024^2 "
                    f"(4096 image + 256 caption tokens), LoRA r={R} on attention + feed_forward (not the noise refiner), "
                    f"gradient checkpointing, fused AdamW, attention={args.attention}",
        "launch": launch_mode, "ms_per_step": ms, "steps_per_s": 1e3 / ms, "hot_path_tflops": flops / (ms * 1e-3) / 1e12,
        "hot_path_frac_of_sustained_bf16_peak": flops / (ms * 1e-3) / 1e12 / peak, "hot_path_flops_per_step": flops,
        "adapter_params": sum(p.numel() for p in params), "loss": float(loss.item()),
        "mem_gb": torch.cuda.max_memory_allocated() / 1e9}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--attention", default="stub", choices=["stub", "sdpa"])
    ap.add_argument("--eager", action="store_true", help="no CUDA graph: every launch goes through Python")
    args = ap.parse_args()
    torch.cuda.set_device(0)
    print(json.dumps(measure(args.batch, args.steps, args.warmup, args.attention, args.eager)), flush=True)


if __name__ == "__main__":
    main()
