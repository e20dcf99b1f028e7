"""PEFT config base (mirror of /root/reference/src/modules/peft/config.py:1-10)."""
from typing import Literal

from pydantic import BaseModel

PEFT_TYPE = Literal["lora", "loha", "none"]


class PeftConfigMixin(BaseModel):
    type: PEFT_TYPE
    dtype: str = "bfloat16"
