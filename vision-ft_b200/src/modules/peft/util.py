"""The contract an adapter layer fulfils towards the surgery helpers in ``functional.py``.

Interface mirror of /root/reference/src/modules/peft/util.py:10-49 (same class name, attributes and method
signatures: ``replace_to_peft_layer``, ``get_adapter_parameters``, ``load_peft_weight`` and the enable/disable
context managers only ever talk to this surface).
"""
from __future__ import annotations

import abc

import torch
from torch import nn


class PeftLayer(abc.ABC, nn.Module):
    # names of the sub-modules / parameters that belong to the adapter (what is trained and saved) ...
    adapter_param_names: list[str]
    # ... and the keys of one layer's entry in an adapter checkpoint
    adapter_weight_names: list[str]
    # False: forward() is the wrapped layer alone
    enabled: bool

    def set_enabled(self, enabled: bool) -> None:
        self.enabled = enabled

    @abc.abstractmethod
    def init_weights(self) -> None:
        """(Re-)initialise the adapter so that the wrapped layer's output is unchanged at step 0."""

    @abc.abstractmethod
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """Wrapped layer + adapter (or the wrapped layer alone when disabled)."""

    @classmethod
    @abc.abstractmethod
    def from_weights(cls, adapter_weights: dict[str, torch.Tensor], original_layer: nn.Module) -> "PeftLayer":
        """Build the adapter around ``original_layer`` from one layer's checkpoint entries."""

    @abc.abstractmethod
    def load_weights(self, adapter_weights: dict[str, torch.Tensor | None]) -> None:
        """Overwrite the adapter tensors named in ``adapter_weight_names``; missing / None entries are skipped."""
