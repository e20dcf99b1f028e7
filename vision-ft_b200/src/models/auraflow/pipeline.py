"""LoRA adapter export / import with the key layouts the AuraFlow trainer writes.

What the reference does at the end of a QLoRA run (/root/reference/src/models/auraflow/train/text_to_image.py via
``get_state_dict_to_save``, /root/reference/train/auraflow/text_to_image.py:145-153): collect the adapter tensors
(``get_adapter_parameters``), rename ``denoiser.`` to ComfyUI's ``diffusion_model.`` and hand the dict to safetensors.
The three renames follow /root/reference/src/models/auraflow/pipeline.py:35-54; the prefixes are the constants of
``denoiser.py:32``, ``vae.py:36`` and ``text_encoder.py:50``.  ``tests/test_peft.py:295-343`` of the reference is the
behaviour pinned in ``tests/test_adapter_export.py`` here.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ...modules.peft.functional import get_adapter_parameters, load_peft_weight

DENOISER_TENSOR_PREFIX = "model."
VAE_TENSOR_PREFIX = "vae."
TEXT_ENCODER_TENSOR_PREFIX = "text_encoders.pile_t5xl.transformer."

# (module-tree prefix, original-checkpoint prefix, ComfyUI prefix); applied in this order, as the reference does
_RENAMES = (
    ("denoiser.", DENOISER_TENSOR_PREFIX, "diffusion_model."),
    ("vae.", VAE_TENSOR_PREFIX, VAE_TENSOR_PREFIX),
    ("text_encoder.model.", TEXT_ENCODER_TENSOR_PREFIX, TEXT_ENCODER_TENSOR_PREFIX),
)


def convert_to_original_key(key: str) -> str:
    for ours, orig, _ in _RENAMES:
        key = key.replace(ours, orig)
    return key


def convert_to_comfy_key(key: str) -> str:
    for ours, _, comfy in _RENAMES:
        key = key.replace(ours, comfy)
    return key


def convert_from_original_key(key: str) -> str:
    key = key.replace("diffusion_model.", "denoiser.")
    for ours, orig, _ in _RENAMES:
        key = key.replace(orig, ours)
    return key


def adapter_state_dict_to_save(model: nn.Module, layout: str = "comfy") -> dict[str, torch.Tensor]:
    """The dict ``get_state_dict_to_save`` returns for a peft run: adapter tensors only, keys in ``layout``
    ("comfy": ``diffusion_model.*``, "original": ``model.*``, "module": the module tree's own names)."""
    convert = {"comfy": convert_to_comfy_key, "original": convert_to_original_key, "module": lambda k: k}[layout]
    return {convert(k): v for k, v in get_adapter_parameters(model).items()}


def save_adapter_file(model: nn.Module, path: str, layout: str = "comfy", metadata: dict[str, str] | None = None) -> None:
    from safetensors.torch import save_file

    tensors = {k: v.detach().to("cpu").contiguous() for k, v in adapter_state_dict_to_save(model, layout).items()}
    save_file(tensors, path, metadata=metadata)


def load_adapter_file(model: nn.Module, path: str) -> None:
    """Load an adapter file in any of the three key layouts into ``model`` (wrapping bare Linears on the fly)."""
    from safetensors.torch import load_file

    state = {convert_from_original_key(k): v for k, v in load_file(path).items()}
    load_peft_weight(model, state)
