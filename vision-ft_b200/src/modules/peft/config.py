"""Fields shared by every adapter configuration.

Interface mirror of /root/reference/src/modules/peft/config.py:1-10: the YAML configs of the reference
(``peft: {config: {type: lora, rank: .., dtype: bfloat16}, include_keys: [...]}``) validate against these names, so
the two field names, the literal values and the default must not change.  Only ``type == "lora"`` has kernels behind
it here; ``"loha"`` is recognised so that configs parse and then raises where it would have to run.
"""
from __future__ import annotations

from typing import Literal

import torch
from pydantic import BaseModel

from ...utils.dtype import str_to_dtype

PEFT_TYPE = Literal["lora", "loha", "none"]


class PeftConfigMixin(BaseModel):
    type: PEFT_TYPE
    dtype: str = "bfloat16"  # dtype of the adapter parameters (and of the fused kernels' adapter operands)

    def torch_dtype(self) -> torch.dtype:
        """``dtype`` as a torch dtype; unknown names raise ValueError (src/utils/dtype.py)."""
        return str_to_dtype(self.dtype)

    def is_adapter(self) -> bool:
        """False for the ``none`` placeholder type (full fine-tuning entries of a target list)."""
        return self.type != "none"
