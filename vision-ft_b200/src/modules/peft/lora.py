"""LoRA adapter over a (quantized) Linear.

Mirror of the reference's ``LoRAConfig`` / ``LoRALinear``
(/root/reference/src/modules/peft/lora.py:11-164): same attributes (``lora_down``,
``lora_up``, ``alpha``, ``dropout``, ``linear``, ``enabled``, ``rank``), same
initialisation (A kaiming-uniform, B zero, :68-76), same adapter key names.  What
differs is ``forward``: when the wrapped layer is the NF4 ``Linear4bit`` on a CUDA
device, base GEMM + adapter run as ONE fused kernel sequence with on-the-fly NF4
decode (vft_b200.ops.qlora_linear) instead of "base(x) + up(down(x)) * s"
(:92-104).  LoRAConv2d / LoHa are outside the hot path (OUT).
"""
from __future__ import annotations

from typing import Literal

import torch
import torch.nn as nn

from vft_b200.nn import Linear4bit

from ...utils.dtype import str_to_dtype
from .config import PeftConfigMixin
from .util import PeftLayer


class LoRAConfig(PeftConfigMixin):
    type: Literal["lora"] = "lora"
    rank: int
    alpha: float = 1.0
    dropout: float = 0.0
    use_bias: bool = False


class LoRALinear(PeftLayer):
    adapter_param_names = ["lora_up", "lora_down", "alpha"]
    adapter_weight_names = ["lora_up.weight", "lora_up.bias", "lora_down.weight", "alpha"]

    def __init__(self, config: LoRAConfig, original_linear: nn.Linear) -> None:
        super().__init__()
        self.config = config
        dtype = str_to_dtype(config.dtype)
        fan_in, fan_out = original_linear.in_features, original_linear.out_features

        self.lora_down = nn.Linear(fan_in, config.rank, bias=False, dtype=dtype)
        self.lora_up = nn.Linear(config.rank, fan_out, bias=False, dtype=dtype)
        self.dropout = self._make_dropout()
        self.alpha = nn.Parameter(torch.tensor(config.alpha, dtype=dtype), requires_grad=False)
        self.rank = config.rank
        if config.use_bias:
            self.lora_up.bias = nn.Parameter(torch.zeros(fan_out, dtype=dtype))
        self.enabled = True

        # the wrapped layer stays frozen (lora.py:60-64)
        self.linear = original_linear
        self.linear.weight.requires_grad_(False)
        if self.linear.bias is not None:
            self.linear.bias.requires_grad_(False)

        self.init_weights()

    def _make_dropout(self) -> nn.Module:
        return nn.Dropout(self.config.dropout) if self.config.dropout > 0 else nn.Identity()

    def init_weights(self) -> None:
        device = self.linear.weight.device
        for mod in (self.lora_down, self.lora_up, self.dropout):
            mod.to_empty(device=device)
        nn.init.kaiming_uniform_(self.lora_down.weight)
        nn.init.zeros_(self.lora_up.weight)
        if self.lora_up.bias is not None:
            nn.init.zeros_(self.lora_up.bias)
        self.alpha = nn.Parameter(
            torch.tensor(self.config.alpha, dtype=self.lora_down.weight.dtype), requires_grad=False
        )
        self.dropout = self._make_dropout()
        if self.alpha.device.type != "meta":
            self._scale_value()  # resolve alpha / rank on the host once, now (never on the hot path, never in a trace)

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = enabled

    # ------------------------------------------------------------------ forward
    def _can_fuse(self, x: torch.Tensor) -> bool:
        if not (isinstance(self.linear, Linear4bit) and x.is_cuda):
            return False
        if isinstance(self.dropout, nn.Dropout) and self.dropout.training and self.dropout.p > 0:
            return False  # adapter sees dropout(x), base sees x: operands differ
        if self.rank > 64:
            return False
        if self.lora_up.bias is not None and self.lora_up.bias.dtype != self.lora_up.weight.dtype:
            return False
        act = self.linear._cast_input(x).dtype
        return self.lora_down.weight.dtype == act and self.lora_up.weight.dtype == act

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        if not self.enabled:
            return self.linear(x)
        if self._can_fuse(x):
            scale = self._scale_value()
            # use_bias (lora.py:52-60 of the reference: a bias on lora_up): (t . B^T + b) * s = s * t . B^T + s * b, so
            # s * b joins the base bias as the per-feature scalar the fused epilogue adds; its gradient (the column sum
            # of dy) comes back through the operator's bias input
            up_bias = self.lora_up.bias * scale if self.lora_up.bias is not None else None
            return self.linear.forward_with_lora(x, self.lora_down.weight, self.lora_up.weight, scale, up_bias)
        # composition of the reference (lora.py:92-104) for bases/dtypes the fused path does not take
        base = self.linear(x)
        down = self.lora_down(self.dropout(x))
        return base + self.lora_up(down) * (self.alpha / self.rank)

    def _scale_value(self) -> float:
        # alpha is a frozen 0-dim parameter; cache its host value so the hot path never syncs
        cached = getattr(self, "_scale_cache", None)
        if cached is not None and torch.compiler.is_compiling():
            return cached[1]  # a torch.compile trace must not read pointers / sync: the value of the last eager call
        key = (self.alpha.data_ptr(), self.alpha._version, self.rank)
        if cached is None or cached[0] != key:
            # same rounding as the reference: (alpha / rank) evaluated in the adapter dtype
            value = float((self.alpha.detach() / self.rank).cpu())
            self._scale_cache = (key, value)
            return value
        return cached[1]

    # ------------------------------------------------------------------ nn.Module protocol
    def train(self, mode: bool = True) -> "LoRALinear":
        # exactly the reference's override (lora.py:106-113): only the two adapter Linears follow `mode`
        self.lora_down.train(mode)
        self.lora_up.train(mode)
        self.linear.train(False)  # the wrapped layer never trains
        return self

    def requires_grad_(self, requires_grad: bool = True) -> "LoRALinear":
        self.lora_down.requires_grad_(requires_grad)
        self.lora_up.requires_grad_(requires_grad)
        self.linear.weight.requires_grad_(False)
        return self

    # ------------------------------------------------------------------ adapter (de)serialisation
    @classmethod
    def from_weights(cls, adapter_weights: dict[str, torch.Tensor], original_layer: nn.Linear) -> "LoRALinear":
        rank = adapter_weights["lora_down.weight"].shape[0]
        alpha = adapter_weights["alpha"].item()
        module = cls(LoRAConfig(rank=rank, alpha=alpha), original_layer)
        device = original_layer.weight.device
        module.lora_down.weight = nn.Parameter(adapter_weights["lora_down.weight"].to(device))
        module.lora_up.weight = nn.Parameter(adapter_weights["lora_up.weight"].to(device))
        module.alpha = nn.Parameter(adapter_weights["alpha"].to(device))
        if adapter_weights.get("lora_up.bias") is not None:
            module.lora_up.bias = nn.Parameter(adapter_weights["lora_up.bias"].to(device))
        return module

    def load_weights(self, adapter_weights: dict[str, torch.Tensor | None]) -> None:
        device = self.lora_down.weight.device
        slots = {
            "lora_down.weight": (self.lora_down, "weight"),
            "lora_up.weight": (self.lora_up, "weight"),
            "lora_up.bias": (self.lora_up, "bias"),
            "alpha": (self, "alpha"),
        }
        for key, (owner, attr) in slots.items():
            tensor = adapter_weights.get(key)
            if tensor is not None:
                setattr(owner, attr, nn.Parameter(tensor.to(device)))
