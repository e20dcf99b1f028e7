"""Block-level fusion of sibling projections: q/k/v -- or SwiGLU's fc1/fc2 -- as ONE launch.

The reference's blocks apply several Linears to the same activations
(/root/reference/src/models/auraflow/denoiser.py:113-117 ``w1q/w1k/w1v``, ``:160-163`` ``c_fc1/c_fc2``;
/root/reference/src/models/lumina2/denoiser.py:212-219; /root/reference/src/models/sdxl/denoiser.py:184-186
``to_q/to_k/to_v``).  Each of them is an NF4 ``Linear4bit`` (optionally under a ``LoRALinear``), i.e. one fused-kernel
launch per projection and direction: the same ``x`` is read three times, three ``dx`` are written and added by two more
element-wise kernels, and at SDXL's sizes (6.7 GFLOP per projection) the ~8 us every launch costs before its first MMA
is most of its run time.

NF4 quantizes along the in-feature axis in blocks of 64, so stacking the members' packed weights along the out-feature
axis is EXACT: ``[N1 + N2 + N3, K]`` is a valid NF4 matrix whose rows decode to what the members' rows decode to.  The
adapters stack the same way: ``A = [A1; A2; A3]`` (``[r1 + r2 + r3, K]``) and ``B = blockdiag(B1, B2, B3)``
(off-diagonal zeros contribute exact zeros), so the group is served by the SAME C-ABI calls as one layer
(``vft_qlora_fwd`` / ``vft_qlora_bwd``): one forward launch that reads ``x`` once, one backward launch that reads the
stacked ``dy`` and leaves the summed ``dx`` -- no kernel changes, no new numerics.

Drop-in: the model code is not touched.  ``fuse_projection_groups(model, [("w1q", "w1k", "w1v"), ...])`` finds parents
that own all the named children and routes the members' ``forward`` through the group: the first member called with
an input computes all outputs, the siblings called with THE SAME tensor pick theirs up (the group keeps the input alive
until then, so an address can not be recycled under it).  Anything the group cannot take -- a torch.compile trace, a
sibling called on a different tensor (cross-attention's k/v), mixed adapter switches, dropout in training -- falls back
to the member's own forward.  Parameters, state_dict keys, optimizers and DDP see the unchanged member modules.
"""
from __future__ import annotations

from typing import Iterable, Sequence

import torch
import torch.nn as nn

from . import ops
from .nn import Linear4bit

LORA_MAX_RANK = 64  # VFT_LORA_LD: the adapter step carries at most 64 rank columns
# One launch instead of three pays while a launch's fixed cost (~8 us before its first MMA) matters: measured on B200
# (tools/group_probe.py, forward + backward through the module API, rank 4 / 16): SDXL C1280 q/k/v at 2048 tokens
# x1.37 / x0.99, C640 at 8192 tokens x1.33 / x1.13, text-token k/v x1.34 / x1.21; at AuraFlow's sizes (494 GFLOP per
# q/k/v forward) the members already run near the tensor roofline and the group LOSES 1-8 %: autograd hands the
# backward three separate dy tensors, and stacking them costs a pass over T x (N1+N2+N3) that the summed dx does not
# win back.  Groups above this much forward work therefore leave their members alone.
MAX_GROUP_GFLOP = 80.0


def _base_of(m: nn.Module) -> Linear4bit | None:
    if isinstance(m, Linear4bit):
        return m
    inner = getattr(m, "linear", None)
    return inner if isinstance(inner, Linear4bit) else None


def _is_adapter(m: nn.Module) -> bool:
    return not isinstance(m, Linear4bit) and hasattr(m, "lora_down") and hasattr(m, "lora_up")


class ProjectionGroup:
    """Sibling NF4(+LoRA) projections over one input, one fused launch per direction."""

    def __init__(self, members: Sequence[nn.Module], names: Sequence[str] | None = None):
        if len(members) < 2:
            raise ValueError("a projection group needs at least two members")
        bases = [_base_of(m) for m in members]
        if any(b is None for b in bases):
            raise TypeError("every member must be a Linear4bit or an adapter over one")
        k = {b.in_features for b in bases}
        if len(k) != 1:
            raise ValueError(f"members read different input widths: {sorted(k)}")
        self.members = list(members)
        self.names = list(names) if names is not None else [str(i) for i in range(len(members))]
        self.in_features = bases[0].in_features
        self.sizes = [b.out_features for b in bases]
        self.out_features = sum(self.sizes)
        self._stacked = None  # (key, (packed, absmax, blocksize, qdtype, tiled, bias))
        self._pending = None  # (x kept alive, outputs, taken flags, grad mode)
        self.launches = 0     # group launches served (tests / census)
        self.fallbacks = 0
        self.max_gflop = MAX_GROUP_GFLOP

    @property
    def bases(self) -> "list[Linear4bit]":
        return [_base_of(m) for m in self.members]  # type: ignore[misc]

    # ------------------------------------------------------------------ stacked operands (derived, rebuilt on change)
    def _operands(self):
        bases = self.bases
        per = [b._operands() for b in bases]
        key = tuple((p[0].data_ptr(), p[0]._version, p[1].data_ptr(), p[1]._version) for p in per) + tuple(
            (None if b.bias is None else (b.bias.data_ptr(), b.bias._version)) for b in bases)
        if self._stacked is not None and self._stacked[0] == key:
            return self._stacked[1]
        blocksize, qdtype = per[0][2], per[0][3]
        if any(p[2] != blocksize or p[3] != qdtype for p in per):
            raise ValueError("members were quantized with different block sizes / dtypes")
        if any((n * self.in_features) % 2 for n in self.sizes):
            raise ValueError("a member's packed weight does not end on a byte boundary")
        packed = torch.cat([p[0].reshape(-1) for p in per]).reshape(-1, 1)
        absmax = torch.cat([p[1].reshape(-1) for p in per])
        tiled = ops.nf4_tile_weight(packed, absmax, self.out_features, self.in_features, blocksize)
        bias = None
        if any(b.bias is not None for b in bases):
            ref = next(b.bias for b in bases if b.bias is not None)
            bias = torch.cat([b.bias.detach() if b.bias is not None else ref.new_zeros(n) for b, n in zip(bases, self.sizes)])
        out = (packed, absmax, blocksize, qdtype, tiled, bias)
        self._stacked = (key, out)
        return out

    # ------------------------------------------------------------------ can the group serve this call?
    def _mode(self, x: torch.Tensor) -> str | None:
        """'lora' / 'base' when one launch can serve every member for this input, None when it cannot."""
        if torch.compiler.is_compiling() or not x.is_cuda:
            return None
        tokens = x.numel() // max(x.shape[-1], 1)
        if 2.0 * tokens * self.in_features * self.out_features > self.max_gflop * 1e9:
            return None
        adapters = [_is_adapter(m) for m in self.members]
        if not any(adapters):
            return "base"
        if not all(adapters):
            return None
        enabled = [bool(getattr(m, "enabled", True)) for m in self.members]
        if not any(enabled):
            return "base"
        if not all(enabled):
            return None
        if sum(m.rank for m in self.members) > LORA_MAX_RANK:
            return None
        if any(m.lora_up.bias is not None or not m._can_fuse(x) for m in self.members):
            return None
        return "lora"

    def _compute(self, x: torch.Tensor, mode: str) -> tuple[torch.Tensor, ...]:
        packed, absmax, blocksize, qdtype, tiled, bias = self._operands()
        base0 = self.bases[0]
        inp_dtype = x.dtype
        xc = base0._cast_input(x)
        a = b = None
        scale = 0.0
        if mode == "lora":
            scales = [m._scale_value() for m in self.members]
            scale = scales[0]
            a = torch.cat([m.lora_down.weight for m in self.members], dim=0)
            ups = [m.lora_up.weight if s == scale else m.lora_up.weight * (s / scale) for m, s in zip(self.members, scales)]
            b = torch.block_diag(*ups)
        y = ops.qlora_linear(xc, packed, absmax, bias, a, b, scale, self.out_features, self.in_features, blocksize,
                             qdtype, tiled)
        if y.dtype != inp_dtype:
            y = y.to(inp_dtype)
        self.launches += 1
        return y.split(self.sizes, dim=-1)

    # ------------------------------------------------------------------ the members' forward
    def member_forward(self, idx: int, x: torch.Tensor) -> torch.Tensor:
        pend = self._pending
        if pend is not None:
            px, outs, taken, grad_mode = pend
            same = px is x or (px.data_ptr() == x.data_ptr() and px._version == x._version and px.shape == x.shape
                               and px.dtype == x.dtype and px.stride() == x.stride())
            if same and not taken[idx] and grad_mode == torch.is_grad_enabled():
                taken[idx] = True
                if all(taken):
                    self._pending = None
                return outs[idx]
            self._pending = None  # a different input (or a member asked twice): whatever was left is dropped
        mode = self._mode(x)
        if mode is None:
            self.fallbacks += 1
            return self._own_forward(idx, x)
        outs = self._compute(x, mode)
        taken = [False] * len(self.members)
        taken[idx] = True
        self._pending = (x, outs, taken, torch.is_grad_enabled())
        return outs[idx]

    def forward(self, x: torch.Tensor) -> tuple[torch.Tensor, ...]:
        """All outputs at once (for callers that hold the group itself)."""
        mode = self._mode(x)
        if mode is None:
            self.fallbacks += 1
            return tuple(self._own_forward(i, x) for i in range(len(self.members)))
        return self._compute(x, mode)

    __call__ = forward

    def _own_forward(self, idx: int, x: torch.Tensor) -> torch.Tensor:
        m = self.members[idx]
        return type(m).__mro__[1].forward(m, x)  # the class the member had before install()

    def install(self) -> "ProjectionGroup":
        """Route the members' forward through the group: each member's class is swapped for a one-method subclass
        (isinstance checks, parameters and state_dict keys are unchanged; a deep copy of the model copies the group
        with it, bound to the copied members)."""
        for i, m in enumerate(self.members):
            m.__class__ = _grouped_class(type(m))
            m.__dict__["_vft_group"] = (self, i)
        return self

    def remove(self) -> None:
        for m in self.members:
            if "_vft_group" in m.__dict__:
                del m.__dict__["_vft_group"]
                m.__class__ = type(m).__mro__[1]
        self._pending = None
        self._stacked = None

    def __deepcopy__(self, memo):
        import copy

        new = ProjectionGroup.__new__(ProjectionGroup)
        memo[id(self)] = new
        new.members = [copy.deepcopy(m, memo) for m in self.members]
        new.names = list(self.names)
        new.in_features, new.sizes, new.out_features = self.in_features, list(self.sizes), self.out_features
        new._stacked = None
        new._pending = None
        new.launches = new.fallbacks = 0
        new.max_gflop = self.max_gflop
        return new


_GROUPED_CLASSES: dict[type, type] = {}


def _grouped_class(cls: type) -> type:
    sub = _GROUPED_CLASSES.get(cls)
    if sub is None:
        def forward(self, x):
            entry = self.__dict__.get("_vft_group")
            if entry is None:
                return cls.forward(self, x)
            return entry[0].member_forward(entry[1], x)

        sub = type(f"Grouped{cls.__name__}", (cls,), {"forward": forward, "__module__": cls.__module__})
        _GROUPED_CLASSES[cls] = sub
    return sub


def fuse_projection_groups(model: nn.Module, groups: Iterable[Sequence[str]]) -> list[ProjectionGroup]:
    """Install a :class:`ProjectionGroup` on every module of ``model`` that owns ALL children named in one of
    ``groups`` (e.g. ``[("w1q", "w1k", "w1v"), ("c_fc1", "c_fc2")]``), provided they are NF4 projections of one input
    width.  Returns the groups installed.  Call after the peft surgery (the members are then the adapter wrappers)."""
    made: list[ProjectionGroup] = []
    for pname, parent in list(model.named_modules()):
        for names in groups:
            kids = [getattr(parent, n, None) for n in names]
            if any(not isinstance(k, nn.Module) or _base_of(k) is None for k in kids):
                continue
            if any("_vft_group" in k.__dict__ for k in kids):
                continue
            if len({_base_of(k).in_features for k in kids}) != 1:
                continue
            made.append(ProjectionGroup(kids, [f"{pname}.{n}" if pname else n for n in names]).install())
    return made


def unfuse_projection_groups(model: nn.Module) -> int:
    groups = {id(e[0]): e[0] for m in model.modules() if (e := m.__dict__.get("_vft_group")) is not None}
    for g in groups.values():
        g.remove()
    return len(groups)
