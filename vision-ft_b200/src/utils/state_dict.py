"""Key matching that decides which Linears get quantized / LoRA-wrapped.

Mirror of the reference's ``RegexMatch`` / ``get_target_keys``
(/root/reference/src/utils/state_dict.py:8-42; semantics pinned by
/root/reference/tests/test_utils.py:12-127): a plain string matches as a substring,
a ``RegexMatch`` matches with ``re.match`` (anchored at the start), and every
exclude pattern wins over every include pattern.
"""
from __future__ import annotations

import re
from typing import Iterable, Sequence

from pydantic import BaseModel


class RegexMatch(BaseModel):
    regex: str

    def __call__(self, value: str) -> bool:
        return re.match(self.regex, value) is not None


def _hits(pattern: "str | RegexMatch", keys: Iterable[str]) -> set[str]:
    if isinstance(pattern, RegexMatch):
        rx = re.compile(pattern.regex)
        return {k for k in keys if rx.match(k)}
    if isinstance(pattern, str):
        return {k for k in keys if pattern in k}
    return set()


def get_target_keys(
    include: Sequence["str | RegexMatch"],
    exclude: Sequence["str | RegexMatch"],
    keys: list[str],
) -> list[str]:
    selected: set[str] = set()
    for pattern in include:
        selected |= _hits(pattern, keys)
    for pattern in exclude:
        selected -= _hits(pattern, keys)
    return list(selected)
