"""CPU oracle for the NF4 base Linear + LoRA adapter, forward and backward.

TEST INFRASTRUCTURE ONLY (see oracle/nf4_oracle.py header for who may import it).

What it restates, with the per-op roundings of the reference:

  * base path -- ``bnb.nn.Linear4bit.forward`` -> ``bnb.matmul_4bit`` ->
    ``MatMul4Bit`` (third-party, bitsandbytes 0.48.2; PARITY UNPINNED, see
    nf4_oracle.py), reached from /root/reference/src/modules/peft/lora.py:93:
        W~ = dequantize_4bit(codes, absmax)      rounded to quant_state.dtype
        Y0 = F.linear(x, W~.to(x.dtype), bias)   fp32 accumulate, one rounding
        dX0 = dY @ W~.to(dY.dtype)               (base frozen: no dW)
  * adapter path -- /root/reference/src/modules/peft/lora.py:92-104 (PINNED:
    tests/golden/make_lora_golden.py imports that file and freezes its outputs):
        down = lora_down(dropout(x)); up = lora_up(down)
        Y = Y0 + up * (alpha / rank)             each op rounded to the LoRA dtype
    backward is autograd of exactly that graph.

``qlora_linear_ref`` runs those torch ops on CPU ("port" of the reference);
``qlora_linear_truth`` is the same math in float64 with no intermediate
rounding, used to calibrate the tolerance the GPU tests state.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import nf4_oracle


def dequant_weight(packed, absmax, shape, qdtype: str = "bfloat16") -> torch.Tensor:
    return nf4_oracle.nf4_dequantize(packed, absmax, shape, qdtype)


def qlora_linear_ref(
    x: torch.Tensor,
    w_deq: torch.Tensor,
    bias: torch.Tensor | None,
    lora_a: torch.Tensor | None,  # lora_down.weight [r, K]
    lora_b: torch.Tensor | None,  # lora_up.weight   [N, r]
    alpha: float,
    dy: torch.Tensor | None = None,
):
    """Reference-rounding forward (+ backward when ``dy`` is given).

    Returns dict(y, dx, da, db); gradients are None without ``dy``.
    """
    x = x.detach().clone().requires_grad_(dy is not None)
    w = w_deq.to(x.dtype)
    a = b = None
    y = F.linear(x, w, None if bias is None else bias.to(x.dtype))
    if lora_a is not None:
        a = lora_a.detach().clone().requires_grad_(dy is not None)
        b = lora_b.detach().clone().requires_grad_(dy is not None)
        rank = a.shape[0]
        alpha_t = torch.tensor(alpha, dtype=a.dtype)  # lora.py:48-51 (0-dim, LoRA dtype)
        down = F.linear(x, a)
        up = F.linear(down, b)
        y = y + up * (alpha_t / rank)
    out = {"y": y.detach(), "dx": None, "da": None, "db": None}
    if dy is not None:
        y.backward(dy)
        out["dx"] = x.grad.detach()
        if a is not None:
            out["da"] = a.grad.detach()
            out["db"] = b.grad.detach()
    return out


def qlora_linear_truth(x, w_deq, bias, lora_a, lora_b, alpha, dy=None):
    """float64, no intermediate rounding.  Same return layout as qlora_linear_ref."""
    lead = x.shape[:-1]
    xd = x.double().reshape(-1, x.shape[-1])
    wd = w_deq.double()
    y = xd @ wd.t()
    if bias is not None:
        y = y + bias.double()
    s = None
    if lora_a is not None:
        ad, bd = lora_a.double(), lora_b.double()
        s = float(alpha) / ad.shape[0]
        t = xd @ ad.t()
        y = y + s * (t @ bd.t())
    out = {"y": y.reshape(*lead, -1), "dx": None, "da": None, "db": None}
    if dy is not None:
        g = dy.double().reshape(-1, dy.shape[-1])
        dx = g @ wd
        if lora_a is not None:
            dt = s * (g @ bd)  # [T, r]
            dx = dx + dt @ ad
            out["da"] = dt.t() @ xd
            out["db"] = s * (g.t() @ t)
        out["dx"] = dx.reshape(*lead, -1)
    return out


def rel_l2(a: torch.Tensor, b: torch.Tensor) -> float:
    a, b = a.double().flatten(), b.double().flatten()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def max_abs(a: torch.Tensor, b: torch.Tensor) -> float:
    return float((a.double() - b.double()).abs().max())
