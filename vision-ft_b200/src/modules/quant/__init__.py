"""Drop-in for /root/reference/src/modules/quant/__init__.py:1-11 (NF4 path only).

``BnbLinear8bit``, ``AOLinearNF4``, ``AOLinearFP8`` and ``QuantoLinear`` are outside the
hot path (OUT); their quant types raise NotImplementedError.
"""
from .bnb import BnbLinear4bit
from .functional import (
    QUANT_TYPE,
    quantize_inplace,
    quantize_state_dict,
    replace_to_quant_linear,
    replace_by_prequantized_weights,
    validate_quant_type,
)

__all__ = [
    "BnbLinear4bit",
    "QUANT_TYPE",
    "quantize_inplace",
    "quantize_state_dict",
    "replace_to_quant_linear",
    "replace_by_prequantized_weights",
    "validate_quant_type",
]
