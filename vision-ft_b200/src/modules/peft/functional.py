"""Adapter surgery, adapter state dicts and enable/disable switches.

Drop-in for the LoRA-on-Linear part of /root/reference/src/modules/peft/functional.py:
``replace_to_peft_layer`` :96-111, ``get_adapter_parameters`` :114-125,
``extract_peft_internal_modules`` :128-141, ``extract_peft_layers`` :144-153,
``detect_peft_method`` :156-160, ``load_peft_weight`` :229-233,
``calculate_trainable_parameters`` / ``print_trainable_parameters`` :243-296,
``while_peft_disabled`` / ``while_peft_enabled`` :302-360.
Conv2d and LoHa adapters are outside the hot path and raise.
"""
from __future__ import annotations

import warnings
from contextlib import contextmanager
from typing import Callable, NamedTuple

import torch
from torch import nn

from ...utils.state_dict import RegexMatch, get_target_keys
from ...utils.tensor import remove_orig_mod_prefix
from .config import PEFT_TYPE, PeftConfigMixin
from .lora import LoRAConfig, LoRALinear
from .util import PeftLayer


def _get_peft_linear(module: nn.Linear, config: PeftConfigMixin) -> PeftLayer:
    if config.type == "none":
        raise ValueError("peft type 'none' is not parameter efficient training")
    if config.type == "lora":
        return LoRALinear(config=LoRAConfig.model_validate(config.model_dump()), original_linear=module)
    if config.type == "loha":
        raise NotImplementedError("LoHa is outside the B200-native hot path (SURVEY.md section 2)")
    raise ValueError(f"Unknown peft type: {config.type}")


def get_peft_linear_class(peft_type: PEFT_TYPE) -> type[PeftLayer]:
    if peft_type == "lora":
        return LoRALinear
    if peft_type == "loha":
        raise NotImplementedError("LoHa is outside the B200-native hot path (SURVEY.md section 2)")
    raise ValueError(f"Unknown peft type: {peft_type}")


def _children(model: nn.Module, prefix: str = ""):
    """(parent, name, full_name, child) in the reference's traversal order; adapters and Linears are leaves."""
    for name, child in list(model.named_children()):
        full = f"{prefix}{name}"
        yield model, name, full, child
        if not isinstance(child, (PeftLayer, nn.Linear, nn.Conv2d)):
            yield from _children(child, f"{full}.")


def replace_to_peft_layer(
    model: nn.Module,
    include_keys: list[str | RegexMatch],
    exclude_keys: list[str | RegexMatch],
    config: PeftConfigMixin,
    freeze_base: bool = False,
) -> None:
    targets = set(get_target_keys(include_keys, exclude_keys, [n for n, _ in model.named_modules()]))
    if freeze_base:
        for _, module in model.named_modules():
            module.requires_grad_(False)
    for parent, name, full, child in list(_children(model)):
        if isinstance(child, PeftLayer) or full not in targets:
            continue
        if isinstance(child, nn.Linear):
            setattr(parent, name, _get_peft_linear(child, config))
        elif isinstance(child, nn.Conv2d):
            raise NotImplementedError("LoRAConv2d is outside the B200-native hot path (SURVEY.md section 2)")


def get_adapter_parameters(model: nn.Module) -> dict[str, torch.Tensor]:
    found: dict[str, torch.Tensor] = {}
    for mod_name, module in model.named_modules():
        names = getattr(module, "adapter_param_names", None)
        if names is None:
            continue
        for key, value in module.state_dict().items():
            if any(key.startswith(n) for n in names):
                found[remove_orig_mod_prefix(f"{mod_name}.{key}")] = value
    return found


def extract_peft_internal_modules(model: nn.Module) -> dict[str, nn.Module]:
    out: dict[str, nn.Module] = {}
    for mod_name, module in model.named_modules():
        for pname in getattr(module, "adapter_param_names", None) or ():
            sub = getattr(module, pname, None)
            if isinstance(sub, nn.Module):
                out[f"{mod_name}.{pname}"] = sub
    return out


def extract_peft_layers(model: nn.Module) -> dict[str, PeftLayer]:
    return {name: m for name, m in model.named_modules() if isinstance(m, PeftLayer)}


def detect_peft_method(state_dict: dict[str, torch.Tensor]) -> PEFT_TYPE:
    return "lora" if any(k.endswith(".lora_up.weight") for k in state_dict) else "none"


def load_peft_weight(model: nn.Module, state_dict: dict[str, torch.Tensor]) -> None:
    """Load adapter tensors; Linears that have adapter weights but no adapter yet get wrapped on the fly."""
    peft_type = detect_peft_method(state_dict)
    if peft_type == "none":
        raise ValueError("Failed to detect peft method from state_dict")
    peft_class = get_peft_linear_class(peft_type)

    for parent, name, full, child in list(_children(model)):
        if not isinstance(child, (PeftLayer, nn.Linear)):
            continue
        weights = {w: state_dict.get(f"{full}.{w}") for w in peft_class.adapter_weight_names}
        complete = all(v is not None for k, v in weights.items() if "bias" not in k)
        if not complete:
            continue
        if isinstance(child, PeftLayer):
            child.load_weights(weights)
        else:
            setattr(parent, name, peft_class.from_weights(weights, child))  # type: ignore[arg-type]


class TrainableParameters(NamedTuple):
    trainable_params: int
    all_param: int
    trainable_percent: float


def calculate_trainable_parameters(model: nn.Module) -> TrainableParameters:
    total = trainable = 0
    for _, p in model.named_parameters():
        total += p.numel()
        trainable += p.numel() if p.requires_grad else 0
    return TrainableParameters(trainable, total, 100 * trainable / total)


def human_readable_param(param_size: int) -> str:
    for unit, value in (("T", 10**12), ("B", 10**9), ("M", 10**6), ("K", 10**3)):
        if param_size >= value:
            return f"{param_size / value:.2f}{unit}"
    return f"{param_size}"


def print_trainable_parameters(model: nn.Module, print_fn: Callable = print):
    trainable, total, pct = calculate_trainable_parameters(model)
    print_fn(
        f"Trainable params: {human_readable_param(trainable)}, "
        f"All params: {human_readable_param(total)}, Trainable%: {pct:.4f}%"
    )
    if trainable == 0:
        warnings.warn("!!!!!No trainable parameters found!!!!!")
        warnings.warn("!!!!!If this is not intended, check your peft config!!!!!")


def set_peft_layer_enabled(model: nn.Module, enabled: bool) -> None:
    for _, module in model.named_modules():
        if hasattr(module, "set_enabled"):
            module.set_enabled(enabled)


@contextmanager
def while_peft_disabled(model: nn.Module):
    """Run the body with every adapter bypassed (base model behaviour); adapters are re-enabled afterwards."""
    try:
        set_peft_layer_enabled(model, False)
        yield
    finally:
        set_peft_layer_enabled(model, True)


@contextmanager
def while_peft_enabled(model: nn.Module):
    """Run the body with every adapter active; adapters are disabled again afterwards."""
    try:
        set_peft_layer_enabled(model, True)
        yield
    finally:
        set_peft_layer_enabled(model, False)
