"""Module-level containers of the NF4 base layer: ``QuantState``, ``Params4bit``, ``Linear4bit``.

These stand where ``bitsandbytes.functional.QuantState``, ``bitsandbytes.nn.Params4bit``
and ``bitsandbytes.nn.Linear4bit`` stand under the reference
(/root/reference/src/modules/quant/bnb.py:6,37,56-64,94-99,122-129): same attribute
names, same checkpoint key format (``weight.absmax``, ``weight.quant_map``,
``weight.quant_state.bitsandbytes__nf4`` [+ ``nested_*``]), same "quantize when the
fp weight first moves to a CUDA device" behaviour -- but backed by libvft_b200.so.
Only NF4 is implemented (BASELINE.json north_star); fp4 raises.
"""
from __future__ import annotations

import json
from typing import Any

import torch
import torch.nn as nn

from . import ops

NF4_QUANT_MAP = (
    -1.0, -0.6961928009986877, -0.5250730514526367, -0.39491748809814453,
    -0.28444138169288635, -0.18477343022823334, -0.09105003625154495, 0.0,
    0.07958029955625534, 0.16093020141124725, 0.24611230194568634, 0.33791524171829224,
    0.44070982933044434, 0.5626170039176941, 0.7229568362236023, 1.0,
)
NESTED_BLOCKSIZE = 256
_DTYPE_BY_NAME = {"float32": torch.float32, "float16": torch.float16, "bfloat16": torch.bfloat16}
_PACKED_KEY = "quant_state.bitsandbytes__"


def _dtype_name(dt: torch.dtype) -> str:
    return str(dt).replace("torch.", "")


_DYNAMIC_MAP: torch.Tensor | None = None


def create_dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> torch.Tensor:
    """The 256-entry ascending 8-bit "dynamic" code of the nested statistics (``nested_quant_map`` in a checkpoint;
    stands where bitsandbytes.functional.create_dynamic_map stands).  Exponent-like layout: 10^-6 .. 10^0 decades,
    each with linearly spaced fraction means, mirrored in sign, plus 0 and 1.  Host-side table construction."""
    global _DYNAMIC_MAP
    if (signed, max_exponent_bits, total_bits) == (True, 7, 8) and _DYNAMIC_MAP is not None:
        return _DYNAMIC_MAP.clone()
    non_sign_bits = total_bits - 1
    extra = 2 ** (non_sign_bits - max_exponent_bits) - 1
    values: list[float] = []

    def decade(exponent: int, n_items: int) -> None:
        edges = torch.linspace(0.1, 1, n_items, dtype=torch.float32)
        means = (edges[:-1] + edges[1:]) / 2.0
        mag = ((10 ** exponent) * means).tolist()
        values.extend(mag)
        if signed:
            values.extend((-(10 ** exponent) * means).tolist())

    e = -(max_exponent_bits - 1)
    for i in range(max_exponent_bits):
        e = -(max_exponent_bits - 1) + i
        decade(e, int(2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed
                      else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1))
    if extra > 0:
        decade(e, extra + 1)
    values += [0.0, 1.0]
    if len(values) != 2 ** total_bits:
        raise ValueError("dynamic map does not fill the code space")
    table = torch.tensor(sorted(values), dtype=torch.float32)
    if (signed, max_exponent_bits, total_bits) == (True, 7, 8):
        _DYNAMIC_MAP = table
        return table.clone()
    return table


class QuantState:
    """Everything needed to decode a packed 4-bit weight (mirror of bitsandbytes' QuantState)."""

    def __init__(self, absmax, shape, dtype, blocksize=64, quant_type="nf4", code=None, offset=None, state2=None):
        self.absmax = absmax  # fp32 [nblocks] (or uint8 when nested, with state2/offset)
        self.shape = torch.Size(shape)
        self.dtype = dtype  # dtype of the ORIGINAL weight; dequantisation rounds to it first
        self.blocksize = int(blocksize)
        self.quant_type = quant_type
        self.code = code if code is not None else torch.tensor(NF4_QUANT_MAP, dtype=torch.float32, device=absmax.device)
        self.offset = offset
        self.state2 = state2
        self.nested = state2 is not None

    # ---- nested ("double quant") statistics: absmax is uint8 indices into state2.code, scaled per 256 by
    # state2.absmax, shifted by offset.  The checkpoint form is kept as loaded / as encoded; the fused kernels read
    # the fp32 vector below, decoded ONCE per weight on the device (bitsandbytes decodes it in every forward).
    def absmax_f32(self) -> torch.Tensor:
        if not self.nested:
            return self.absmax
        # keyed on identity + in-place version of everything it is decoded from: an absmax / state2 that is
        # reassigned, or overwritten in place (weight.data.copy_, a reload into the same storage), invalidates it
        s2 = self.state2
        key = (self.absmax.data_ptr(), self.absmax._version, s2.absmax.data_ptr(), s2.absmax._version,
               float(self.offset), self.absmax.device)
        cached = self.__dict__.get("_absmax_f32")
        if cached is None or cached[0] != key:
            cached = (key, ops.absmax_denest(self.absmax, s2.absmax, s2.code, float(self.offset), s2.blocksize))
            self.__dict__["_absmax_f32"] = cached
        return cached[1]

    def __deepcopy__(self, memo):
        """Tensors cloned, derived buffers dropped (bitsandbytes deep-copies its QuantState too: a copied model --
        EMA, reference copy -- must not share statistics that .to() moves in place)."""
        import copy

        new = QuantState(
            absmax=self.absmax.clone(), shape=self.shape, dtype=self.dtype, blocksize=self.blocksize,
            quant_type=self.quant_type, code=self.code.clone(),
            offset=self.offset.clone() if torch.is_tensor(self.offset) else self.offset,
            state2=copy.deepcopy(self.state2, memo),
        )
        memo[id(self)] = new
        return new

    def to(self, device) -> "QuantState":
        self.absmax = self.absmax.to(device)
        self.code = self.code.to(device)
        self.__dict__.pop("_absmax_f32", None)
        if self.nested:
            self.offset = self.offset.to(device) if torch.is_tensor(self.offset) else self.offset
            self.state2.absmax = self.state2.absmax.to(device)
            self.state2.code = self.state2.code.to(device)
        return self

    def as_dict(self, packed: bool = False) -> dict[str, Any]:
        d: dict[str, Any] = {
            "quant_type": self.quant_type,
            "absmax": self.absmax,
            "blocksize": self.blocksize,
            "quant_map": self.code,
            "dtype": _dtype_name(self.dtype),
            "shape": tuple(self.shape),
        }
        if self.nested:
            d.update(
                {
                    "nested_absmax": self.state2.absmax,
                    "nested_blocksize": self.state2.blocksize,
                    "nested_quant_map": self.state2.code.clone(),
                    "nested_dtype": _dtype_name(self.state2.dtype),
                    "nested_offset": float(self.offset),
                }
            )
        if not packed:
            return d
        tensors = {k: v for k, v in d.items() if torch.is_tensor(v)}
        meta = {k: (list(v) if isinstance(v, tuple) else v) for k, v in d.items() if not torch.is_tensor(v)}
        blob = torch.tensor(list(json.dumps(meta).encode("utf-8")), dtype=torch.uint8)
        tensors[_PACKED_KEY + self.quant_type] = blob
        return tensors

    @classmethod
    def from_dict(cls, qs_dict: dict[str, Any], device) -> "QuantState":
        qs = dict(qs_dict)
        blob_keys = [k for k in qs if _PACKED_KEY in k and torch.is_tensor(qs[k])]
        if len(blob_keys) > 1 or (not blob_keys and "quant_type" not in qs):
            raise ValueError("expected exactly one packed 'quant_state.bitsandbytes__*' entry or an unpacked dict")
        if blob_keys:
            blob = qs.pop(blob_keys[0])
            qs.update(json.loads(bytes(blob.detach().cpu().to(torch.uint8).tolist()).decode("utf-8")))
        qs = {k.split(".")[-1]: v for k, v in qs.items()}
        if qs["quant_type"] != "nf4":
            raise NotImplementedError(f"vft_b200 implements nf4 only, checkpoint has quant_type={qs['quant_type']!r}")
        state2 = offset = None
        if "nested_absmax" in qs:
            offset = torch.tensor(float(qs["nested_offset"]), device=device)
            state2 = QuantState(
                absmax=qs["nested_absmax"].to(device),
                shape=qs["absmax"].shape,
                dtype=_DTYPE_BY_NAME[qs["nested_dtype"]],
                blocksize=qs["nested_blocksize"],
                quant_type="dynamic8",
                code=qs["nested_quant_map"].to(device),
            )
        return cls(
            absmax=qs["absmax"].to(device),
            shape=qs["shape"],
            dtype=_DTYPE_BY_NAME[qs["dtype"]],
            blocksize=qs["blocksize"],
            quant_type=qs["quant_type"],
            code=qs["quant_map"].to(device),
            offset=offset,
            state2=state2,
        )


def quantize_4bit(w: torch.Tensor, blocksize: int = 64, compress_statistics: bool = False, quant_type: str = "nf4",
                  quant_storage: torch.dtype = torch.uint8) -> tuple[torch.Tensor, QuantState]:
    """CUDA NF4 quantize/pack (stands where bitsandbytes.functional.quantize_4bit stands,
    /root/reference/src/modules/quant/functional.py:12,362-365)."""
    if quant_type != "nf4":
        raise NotImplementedError(f"vft_b200 implements nf4 only, got quant_type={quant_type!r}")
    if quant_storage != torch.uint8:
        raise NotImplementedError("only uint8 quant_storage is implemented")
    if w.dtype not in _DTYPE_BY_NAME.values():
        raise ValueError(f"cannot quantize dtype {w.dtype}")
    packed, absmax = ops.nf4_quantize(w, blocksize)
    if not compress_statistics:
        return packed, QuantState(absmax=absmax, shape=w.shape, dtype=w.dtype, blocksize=blocksize, quant_type="nf4")
    # nested statistics: offset = mean(absmax); 8-bit dynamic-map blockwise code of (absmax - offset), blocksize 256
    code2 = create_dynamic_map().to(absmax.device)
    # offset exactly as bitsandbytes takes it -- torch's fp32 absmax.mean() on the device -- so that a checkpoint written
    # here can be byte-identical to one written by the reference on the same machine (nested_offset sits in the JSON blob
    # and every nested index depends on it); the library's own fp64-accumulated mean (offset=None) differs from it by
    # at most the rounding of torch's summation order
    absmax8, absmax2, offset = ops.absmax_nest(absmax, code2, NESTED_BLOCKSIZE, offset=absmax.mean())
    state2 = QuantState(absmax=absmax2, shape=absmax.shape, dtype=torch.float32, blocksize=NESTED_BLOCKSIZE,
                        quant_type="dynamic8", code=code2)
    return packed, QuantState(absmax=absmax8, shape=w.shape, dtype=w.dtype, blocksize=blocksize, quant_type="nf4",
                              offset=offset, state2=state2)


def dequantize_4bit(packed: torch.Tensor, quant_state: QuantState) -> torch.Tensor:
    return ops.nf4_dequantize(packed, quant_state.absmax_f32(), quant_state.shape, quant_state.dtype, quant_state.blocksize)


class Params4bit(torch.nn.Parameter):
    """Packed 4-bit weight.  Holds fp data until it first reaches a CUDA device, where it is
    quantized in place (``bnb_quantized`` flips to True and ``quant_state`` appears)."""

    def __new__(cls, data=None, requires_grad=False, quant_state=None, blocksize=64, compress_statistics=True,
                quant_type="fp4", quant_storage=torch.uint8, module=None, bnb_quantized=False):
        if data is None:
            data = torch.empty(0)
        self = torch.Tensor._make_subclass(cls, data, requires_grad)
        self.blocksize = blocksize
        self.compress_statistics = compress_statistics
        self.quant_type = quant_type
        self.quant_state = quant_state
        self.quant_storage = quant_storage
        self.bnb_quantized = bnb_quantized
        self.module = module
        return self

    def __deepcopy__(self, memo):
        import copy

        # the owning module is resolved through memo (deepcopy of a model reaches it before its parameters); the
        # quant state is copied, not shared
        state = copy.deepcopy(self.quant_state, memo)
        module = memo.get(id(self.module), self.module) if self.module is not None else None
        new = type(self).__new__(
            type(self), self.data.clone(), self.requires_grad, state, self.blocksize,
            self.compress_statistics, self.quant_type, self.quant_storage, module, self.bnb_quantized,
        )
        memo[id(self)] = new
        if module is not None and module is not self.module and state is not None:
            module.quant_state = state
        return new

    @classmethod
    def from_prequantized(cls, data, quantized_stats, requires_grad=False, device="cuda", module=None, **kwargs):
        if str(device).startswith("cuda") and not torch.cuda.is_available():
            device = data.device  # nothing to run on; keep the packed bytes where they are
        self = torch.Tensor._make_subclass(cls, data.to(device), requires_grad)
        self.quant_state = QuantState.from_dict(quantized_stats, device=device)
        self.blocksize = self.quant_state.blocksize
        self.compress_statistics = self.quant_state.nested
        self.quant_type = self.quant_state.quant_type
        self.quant_storage = data.dtype
        self.bnb_quantized = True
        self.module = module
        if module is not None:
            module.quant_state = self.quant_state
        return self

    def _quantize(self, device):
        w = self.data.contiguous().to(device)
        packed, state = quantize_4bit(w, blocksize=self.blocksize, compress_statistics=self.compress_statistics,
                                      quant_type=self.quant_type, quant_storage=self.quant_storage)
        self.data = packed
        self.quant_state = state
        if self.module is not None:
            self.module.quant_state = state
        self.bnb_quantized = True
        return self

    def cuda(self, device=None, non_blocking=False):
        return self.to(device="cuda" if device is None else device, non_blocking=non_blocking)

    def cpu(self):
        return self.to(device="cpu")

    def to(self, *args, **kwargs):
        device, dtype, non_blocking, _ = torch._C._nn._parse_to(*args, **kwargs)
        if device is not None and device.type == "cuda" and not self.bnb_quantized and self.data.device.type != "meta":
            return self._quantize(device)
        if self.bnb_quantized:
            dtype = None  # packed bytes never change dtype
        new = Params4bit(
            super().to(device=device, dtype=dtype, non_blocking=non_blocking), requires_grad=self.requires_grad,
            quant_state=self.quant_state, blocksize=self.blocksize, compress_statistics=self.compress_statistics,
            quant_type=self.quant_type, quant_storage=self.quant_storage, module=self.module,
            bnb_quantized=self.bnb_quantized,
        )
        if self.quant_state is not None and device is not None:
            self.quant_state.to(device)
        return new


class Linear4bit(nn.Linear):
    """NF4 base layer; forward = fused dequant GEMM (stands where bnb.nn.Linear4bit stands)."""

    def __init__(self, input_features, output_features, bias=True, compute_dtype=None, compress_statistics=True,
                 quant_type="fp4", quant_storage=torch.uint8, device=None):
        super().__init__(input_features, output_features, bias, device)
        self.weight = Params4bit(self.weight.data, requires_grad=False, compress_statistics=compress_statistics,
                                 quant_type=quant_type, quant_storage=quant_storage, module=self)
        self.compute_dtype = compute_dtype
        self.quant_state = None
        self.quant_storage = quant_storage

    # derived device buffers (instance attributes once built; never parameters, buffers or state_dict entries)
    _vft_operands = None
    _vft_tiled = None

    def _packed(self):
        w = self.weight
        qs = getattr(w, "quant_state", None) or self.quant_state
        if qs is None or not getattr(w, "bnb_quantized", False):
            raise RuntimeError(
                "Linear4bit weight is not quantized yet: move the module to a CUDA device first "
                "(the NF4 path has no CPU implementation)"
            )
        return w.data, qs

    def _tiled(self, packed: torch.Tensor, qs: "QuantState"):
        """Micro-tiled copy of the packed weight for the fused kernels, built once per (storage, device) on first
        use.  Derived data: not a parameter, not a buffer, never in state_dict()."""
        absmax = qs.absmax_f32()
        # identity AND in-place version: packed bytes rewritten in the same storage must rebuild the copy (the
        # few-token kernel reads the checkpoint layout directly: a stale copy would make the two paths disagree)
        key = (packed.data_ptr(), packed._version, absmax.data_ptr(), absmax._version, packed.device)
        cache = self._vft_tiled
        if cache is None or cache[0] != key:
            tiles = ops.nf4_tile_weight(packed, absmax, self.out_features, self.in_features, qs.blocksize)
            cache = (key, tiles)
            self.__dict__["_vft_tiled"] = cache
        return cache[1]

    def _cast_input(self, x: torch.Tensor) -> torch.Tensor:
        if torch.is_autocast_enabled("cuda") and x.is_cuda:
            return x.to(torch.get_autocast_dtype("cuda"))
        if self.compute_dtype is not None:
            return x.to(self.compute_dtype)
        return x

    def _apply(self, fn, *args, **kwargs):
        # .to() / .cuda() / .cpu() replace the weight: the derived device buffers (micro-tiled copy, decoded
        # statistics) of the old one must not outlive it
        self.__dict__.pop("_vft_operands", None)
        self.__dict__.pop("_vft_tiled", None)
        out = super()._apply(fn, *args, **kwargs)
        # build them right away when the (quantized) weight now lives on a CUDA device: the first forward then finds
        # them, and so does a torch.compile trace (which must not look at data pointers / version counters)
        w = self.weight
        if getattr(w, "bnb_quantized", False) and w.is_cuda and getattr(w, "quant_state", None) is not None:
            self._operands()
        return out

    def _operands(self):
        """(packed, fp32 absmax, blocksize, quant dtype, tiled copy) of the current weight, resolved once per
        (weight storage, quant state) and then served from a single cache hit: this runs in every forward, and
        host time per call is what bounds small-batch steps (tools/host_overhead.py)."""
        w = self.weight
        cache = self._vft_operands
        if torch.compiler.is_compiling():
            # under torch.compile the buffers prepared by .cuda() / the first eager call are used as they are
            if cache is None:
                raise RuntimeError("Linear4bit: move the module to a CUDA device before compiling it")
            return cache[2]
        if cache is not None and cache[0] is w and cache[1] == (w.data_ptr(), w._version, cache[3].absmax._version):
            return cache[2]
        packed, qs = self._packed()
        ops_ = (packed, qs.absmax_f32(), qs.blocksize, qs.dtype, self._tiled(packed, qs))
        self.__dict__["_vft_operands"] = (w, (w.data_ptr(), w._version, qs.absmax._version), ops_, qs)
        return ops_

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        packed, absmax, blocksize, qdtype, tiled = self._operands()
        inp_dtype = x.dtype
        x = self._cast_input(x)
        out = ops.qlora_linear(x, packed, absmax, self.bias, None, None, 0.0, self.out_features, self.in_features,
                               blocksize, qdtype, tiled)
        return out if out.dtype == inp_dtype else out.to(inp_dtype)

    def forward_with_lora(self, x: torch.Tensor, lora_a: torch.Tensor, lora_b: torch.Tensor, scale: float,
                          extra_bias: torch.Tensor | None = None) -> torch.Tensor:
        """One fused kernel sequence for base + adapter (used by LoRALinear).  ``extra_bias`` [out_features] (may
        require grad) is added to the layer's own bias: one per-feature scalar in the kernel's epilogue."""
        packed, absmax, blocksize, qdtype, tiled = self._operands()
        inp_dtype = x.dtype
        x = self._cast_input(x)
        bias = self.bias
        if extra_bias is not None:
            bias = extra_bias.to(x.dtype) if bias is None else bias.to(x.dtype) + extra_bias.to(x.dtype)
        out = ops.qlora_linear(x, packed, absmax, bias, lora_a, lora_b, scale, self.out_features,
                               self.in_features, blocksize, qdtype, tiled)
        return out if out.dtype == inp_dtype else out.to(inp_dtype)

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        qs = getattr(self.weight, "quant_state", None)
        if qs is not None:
            for k, v in qs.as_dict(packed=True).items():
                destination[prefix + "weight." + k] = v if keep_vars else v.detach()
