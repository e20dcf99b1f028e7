"""Model surgery for quantized Linears + state-dict quantization.

Drop-in for /root/reference/src/modules/quant/functional.py (public names and
signatures unchanged: ``QUANT_TYPE``, ``validate_quant_type``,
``replace_to_quant_linear`` :127-147, ``quantize_inplace`` :238-253,
``replace_by_prequantized_weights`` :332-339, ``quantize_state_dict`` :342-371,
``collect_children_dict`` :261-274, ``get_quant_type_from_children_dict`` :277-300).
Only ``bnb_nf4`` is executable here; every other quant type is recognised (so
``validate_quant_type`` and checkpoint detection behave as before) and raises
NotImplementedError when it would have to run -- no multi-backend dispatch.
"""
from __future__ import annotations

from typing import Callable, Literal, get_args

import torch
import torch.nn as nn

from vft_b200 import ops
from vft_b200.nn import QuantState

from ...utils.state_dict import get_target_keys
from .bnb import BnbLinear4bit

QUANT_TYPE = Literal[
    "fp8_e4m3fn",
    "bnb_int8",
    "bnb_fp4",
    "bnb_nf4",
    "quanto_int4",
    "quanto_int8",
    "ao_nf4",
    "ao_fp8",
]
_KNOWN = get_args(QUANT_TYPE)
_IMPLEMENTED = ("bnb_nf4",)


def validate_quant_type(quant_type: str) -> None:
    if quant_type not in _KNOWN:
        raise ValueError(f"Unknown quant_type: {quant_type}")


def _require_implemented(quant_type: str) -> None:
    validate_quant_type(quant_type)
    if quant_type not in _IMPLEMENTED:
        raise NotImplementedError(
            f"quant_type {quant_type!r} is outside the B200-native hot path; implemented: {_IMPLEMENTED}"
        )


def _get_quant_linear(module: nn.Linear, quant_type: QUANT_TYPE) -> nn.Module:
    _require_implemented(quant_type)
    return BnbLinear4bit(
        module.in_features,
        module.out_features,
        bias=module.bias is not None,
        quant_type=quant_type[len("bnb_"):],
    )


def _walk_linears(root: nn.Module, visit: Callable[[nn.Module, str, str, nn.Linear], None], prefix: str = "") -> None:
    """Depth-first over ``named_children``; ``visit(parent, child_name, full_name, linear)`` on every nn.Linear
    (Linears are leaves for this purpose, exactly as in the reference's recursive helpers)."""
    for name, child in list(root.named_children()):
        full = f"{prefix}{name}"
        if isinstance(child, nn.Linear):
            visit(root, name, full, child)
        else:
            _walk_linears(child, visit, f"{full}.")


def replace_to_quant_linear(
    model: nn.Module,
    quant_type: QUANT_TYPE,
    include_keys: list[str],
    exclude_keys: list[str] = [],
) -> nn.Module:
    """Swap matching ``nn.Linear`` modules for empty quantized Linears (weights come later from a state dict)."""
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))

    def visit(parent: nn.Module, name: str, full: str, layer: nn.Linear) -> None:
        if full in targets:
            q = _get_quant_linear(layer, quant_type)
            q.requires_grad_(False)
            setattr(parent, name, q)

    _walk_linears(model, visit)
    return model


def quantize_inplace(
    model: nn.Module,
    quant_type: QUANT_TYPE,
    include_keys: list[str],
    exclude_keys: list[str] = [],
) -> None:
    """Swap matching Linears for quantized ones carrying the current weights (quantized on ``.cuda()``)."""
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))

    def visit(parent: nn.Module, name: str, full: str, layer: nn.Linear) -> None:
        if full in targets:
            q = _get_quant_linear(layer, quant_type)
            q.load_state_dict(layer.state_dict(), assign=True)
            setattr(parent, name, q)

    _walk_linears(model, visit)


def collect_children_dict(
    prefix: str,
    state_dict: dict[str, torch.Tensor],
    remove_prefix: bool = True,
) -> dict[str, torch.Tensor]:
    cut = len(prefix) if remove_prefix else 0
    return {k[cut:]: v for k, v in state_dict.items() if k.startswith(prefix)}


def get_quant_type_from_children_dict(children_dict: dict[str, torch.Tensor]) -> QUANT_TYPE:
    """Detect the checkpoint flavour from the sub-keys of ``<layer>.weight.`` (functional.py:277-300)."""
    for key, tensor in children_dict.items():
        if "quant_state" in key:
            flavour = key[len("quant_state.bitsandbytes__"):]
            if flavour in "nf4":
                return "bnb_nf4"
            if flavour in "fp4":
                return "bnb_fp4"
        elif "weight_format" in key:
            return "bnb_int8"
        elif "_data" in key:
            if tensor.dtype == torch.int8:
                return "quanto_int8"
            if tensor.dtype == torch.uint8:
                return "quanto_int4"
    raise ValueError("quant_type not found")


def replace_by_prequantized_weights(model: nn.Module, state_dict: dict[str, torch.Tensor]) -> None:
    """Swap every Linear that has ``<name>.weight.*`` entries in ``state_dict`` for its quantized class."""

    def visit(parent: nn.Module, name: str, full: str, layer: nn.Linear) -> None:
        children = collect_children_dict(f"{full}.weight.", state_dict)
        if children:
            q = _get_quant_linear(layer, get_quant_type_from_children_dict(children))
            q.requires_grad_(False)
            setattr(parent, name, q)

    _walk_linears(model, visit)


def quantize_state_dict(
    state_dict: dict[str, torch.Tensor],
    quant_type: QUANT_TYPE,
    include_keys: list[str],
    exclude_keys: list[str] = [],
) -> dict[str, torch.Tensor]:
    """Quantize the matching tensors of a state dict in place and add their quant-state entries."""
    if quant_type not in ("bnb_nf4", "bnb_fp4", "fp8_e4m3fn"):
        raise NotImplementedError("Only bitsandbytes 4bit quantization is supported")
    _require_implemented(quant_type)
    targets = set(get_target_keys(include_keys, exclude_keys, list(state_dict.keys())))
    keys = [k for k in list(state_dict.keys()) if k in targets]
    # The reference quantizes tensor by tensor (.cuda() -> quantize_4bit -> .cpu()).  Here the uploads of a group of
    # tensors of one dtype (at most ~2 GB of weights) are followed by ONE batched launch (vft_nf4_quantize_many, 96
    # tensors per kernel, bit-identical to the per-tensor calls): a checkpoint is mostly small weights, and per-tensor
    # launches leave the HBM-bound kernel waiting on launch latency.
    group: list[str] = []
    group_bytes = 0

    def flush() -> None:
        nonlocal group, group_bytes
        if not group:
            return
        dev_w = [state_dict[k].cuda() for k in group]
        for k, w, (packed, absmax) in zip(group, dev_w, ops.nf4_quantize_many(dev_w)):
            state = QuantState(absmax=absmax, shape=w.shape, dtype=w.dtype, blocksize=64, quant_type="nf4")
            state_dict[k] = packed.cpu()
            for state_key, state_value in state.as_dict(packed=True).items():
                state_dict[f"{k}.{state_key}"] = state_value.cpu()
        group, group_bytes = [], 0

    for key in keys:
        w = state_dict[key]
        if group and (w.dtype != state_dict[group[0]].dtype or group_bytes + w.numel() * w.element_size() > (2 << 30)):
            flush()
        group.append(key)
        group_bytes += w.numel() * w.element_size()
    flush()
    return state_dict
