"""Only the helper the peft module needs (reference: /root/reference/src/utils/tensor.py:131-135)."""


def remove_orig_mod_prefix(name: str) -> str:
    """torch.compile wraps modules as ``_orig_mod.<name>``; strip the first such prefix."""
    return name.replace("_orig_mod.", "", 1)
