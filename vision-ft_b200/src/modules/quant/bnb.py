"""``BnbLinear4bit``: the NF4 Linear the trainer and the quantize tool instantiate.

Same constructor, attributes and checkpoint behaviour as the reference class
(/root/reference/src/modules/quant/bnb.py:37-129), but its base is
``vft_b200.nn.Linear4bit`` (hand-written sm_100a kernels) instead of
``bitsandbytes.nn.Linear4bit``.
"""
from __future__ import annotations

from typing import Literal

import torch
import torch.nn as nn

from vft_b200.nn import Linear4bit, Params4bit


def collect_quantized_stats(
    prefix: str,
    state_dict: dict[str, torch.Tensor],
    remove_prefix: bool = True,
) -> dict[str, torch.Tensor]:
    """Entries of ``state_dict`` under ``prefix`` (bnb.py:11-24)."""
    cut = len(prefix) if remove_prefix else 0
    return {k[cut:]: v for k, v in state_dict.items() if k.startswith(prefix)}


def _get_bnb_4bit_quant_type_from_stats(quantized_stats: dict[str, torch.Tensor]) -> Literal["fp4", "nf4"]:
    marker = "quant_state.bitsandbytes__"
    for key in quantized_stats:
        if "quant_state" in key:
            quant_type = key[len(marker):]
            assert quant_type in ("nf4", "fp4")
            return quant_type  # type: ignore[return-value]
    raise ValueError("quant_type not found")


class BnbLinear4bit(Linear4bit):
    def __init__(
        self,
        input_features,
        output_features,
        bias=True,
        compute_dtype=None,
        compress_statistics=True,
        quant_type="fp4",
        quant_storage=torch.uint8,
        device=None,
    ):
        if quant_type != "nf4":
            raise NotImplementedError(
                f"BnbLinear4bit(quant_type={quant_type!r}): only 'nf4' is implemented by the B200-native path"
            )
        # The fp weight nn.Linear would allocate and initialise is replaced two lines below by 'meta' placeholders
        # (bnb.py:56-69), so it is never materialised: 322 AuraFlow Linears would otherwise cost 27 GB of host RNG.
        super().__init__(input_features, output_features, bias=bias, device="meta" if device is None else device)
        self.weight = Params4bit(
            torch.empty(output_features, input_features, dtype=compute_dtype, device="meta"),
            requires_grad=False,
            compress_statistics=compress_statistics,
            quant_type=quant_type,
            quant_storage=quant_storage,
            module=self,
        )
        if bias:
            self.bias = nn.Parameter(
                torch.empty(output_features, dtype=compute_dtype, device="meta"), requires_grad=False
            )
        self.compute_dtype = compute_dtype
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_storage = quant_storage

    def _load_from_state_dict(
        self,
        state_dict: dict[str, torch.Tensor],
        prefix: str,
        local_metadata: dict,
        strict: bool,
        missing_keys: list[str],
        unexpected_keys: list[str],
        error_msgs: list[str],
    ):
        stats = collect_quantized_stats(f"{prefix}weight.", state_dict)
        if stats:
            # pre-quantized checkpoint: packed bytes + quant state (bnb.py:91-107)
            quant_type = _get_bnb_4bit_quant_type_from_stats(stats)
            self.weight = Params4bit.from_prequantized(
                data=state_dict[f"{prefix}weight"], quantized_stats=stats, quant_type=quant_type, module=self
            )
            if self.bias is not None:
                self.bias = nn.Parameter(state_dict[f"{prefix}bias"], requires_grad=False)
            return

        # full-precision checkpoint: load normally, then re-wrap so the first move to a CUDA
        # device quantizes it (bnb.py:108-129)
        super()._load_from_state_dict(
            state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs
        )
        self.weight = Params4bit(
            self.weight.data,
            requires_grad=False,
            compress_statistics=self.compress_statistics,
            quant_type=self.quant_type,
            quant_storage=self.quant_storage,
            module=self,
        )
