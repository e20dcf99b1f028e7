"""dtype names used in configs (mirror of /root/reference/src/utils/dtype.py:4-15)."""
import torch

_ALIASES = {
    torch.bfloat16: ("bfloat16", "bf16"),
    torch.float16: ("float16", "fp16"),
    torch.float32: ("float32", "fp32", "float"),
}
_BY_NAME = {name: dt for dt, names in _ALIASES.items() for name in names}


def str_to_dtype(dtype: str) -> torch.dtype:
    try:
        return _BY_NAME[dtype.lower()]
    except KeyError:
        raise ValueError(f"Unknown dtype: {dtype}") from None
