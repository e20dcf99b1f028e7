"""Checkpoint-key side of the AuraFlow pipeline only (adapter export / import): the model itself is out of scope
(DESIGN.md section 6).  Same import path as the reference's ``src.models.auraflow.pipeline``."""
from .pipeline import (  # noqa: F401
    DENOISER_TENSOR_PREFIX,
    TEXT_ENCODER_TENSOR_PREFIX,
    VAE_TENSOR_PREFIX,
    adapter_state_dict_to_save,
    convert_from_original_key,
    convert_to_comfy_key,
    convert_to_original_key,
    load_adapter_file,
    save_adapter_file,
)
