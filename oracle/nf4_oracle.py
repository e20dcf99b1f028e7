"""CPU oracle for the NF4 quantize/pack + dequantize half of the QLoRA hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported by the
product path (``vision-ft_b200/``); only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s CPU-baseline / ``--impl reference`` legs use it, as the
checker or as the timed CPU baseline.

PARITY UNPINNED for this file: the arithmetic lives in the third-party
dependency ``bitsandbytes==0.48.2`` (/root/reference/uv.lock:307-308), which is
not vendored under /root/reference and is not installable here (no network).
The reference only *calls* it:

  * /root/reference/src/modules/quant/functional.py:12,362-368
        ``quantize_4bit(t.cuda(), quant_type="nf4")`` + ``state.as_dict(packed=True)``
  * /root/reference/src/modules/quant/bnb.py:56-64,94-99,122-129
        ``bnb.nn.Params4bit`` / ``Params4bit.from_prequantized``
  * /root/reference/src/modules/peft/lora.py:93  ``self.linear(x)`` -> ``bnb.matmul_4bit``

and its tests hold no golden vectors for codes / absmax / dequantized values
(/root/reference/tests/test_modules_quant.py:154-193 is a round-trip
self-consistency check).  What follows restates the *published* bitsandbytes
blockwise-NF4 algorithm ("[bnb-recall]"):

  absmax_b = max_i |float32(w_i)|            over each 64-element block of the
                                             FLATTENED weight
  s_b      = 1.0f / absmax_b                 (IEEE round-to-nearest fp32)
  x_i      = float32(w_i) * s_b              (one fp32 multiply)
  code_i   = #{ t in THRESHOLDS : t < x_i }  (== bnb's dQuantizeNF4 '>' tree)
  byte_j   = code_{2j} << 4 | code_{2j+1}    (odd tail: low nibble 0)
  dequant  = float32(CODEBOOK[code_i]) * absmax_b  -> rounded to the weight dtype

Edge case adopted (unverified against bnb): an all-zero block has absmax 0,
s = +inf, x = NaN, every '>' is false -> code 0; it decodes to -1.0 * 0 = -0.0.
"""
from __future__ import annotations

import json

import numpy as np

BLOCKSIZE = 64

# fp32-exact NF4 code book (also the ``quant_map`` bitsandbytes stores).
NF4_CODEBOOK = np.array(
    [
        -1.0,
        -0.6961928009986877,
        -0.5250730514526367,
        -0.39491748809814453,
        -0.28444138169288635,
        -0.18477343022823334,
        -0.09105003625154495,
        0.0,
        0.07958029955625534,
        0.16093020141124725,
        0.24611230194568634,
        0.33791524171829224,
        0.44070982933044434,
        0.5626170039176941,
        0.7229568362236023,
        1.0,
    ],
    dtype=np.float32,
)

# The 15 decision thresholds of bitsandbytes' dQuantizeNF4.  In the CUDA source they are
# f-suffixed literals, i.e. the decimal string rounded DIRECTLY to fp32.  Each decimal is the
# float64 midpoint of two adjacent code-book entries, which sits (almost) exactly halfway
# between two fp32 values, so float32(float64(decimal)) -- what numpy would do -- lands 1 ulp
# away from the literal in 4 of the 15 cases (t0, t8, t12, t14).  The bit patterns below are
# strtof() of each literal (== what gcc/nvcc emit for "<decimal>f").
NF4_THRESHOLDS = np.array(
    [
        0xBF591CD9,  # -0.8480964004993439f
        0xBF1C5270,  # -0.6106329262256622f
        0xBEEB8480,  # -0.4599952697753906f
        0xBEADEA76,  # -0.33967943489551544f
        0xBE703CEC,  # -0.23460740596055984f
        0xBE0D38BC,  # -0.13791173323988914f
        0xBD3A7871,  # -0.045525018125772476f
        0x3D22FAFF,  # 0.03979014977812767f
        0x3DF64863,  # 0.1202552504837513f
        0x3E5067E0,  # 0.2035212516784668f
        0x3E9582D4,  # 0.2920137718319893f
        0x3EC753F9,  # 0.3893125355243683f
        0x3F006D03,  # 0.5016634166240692f
        0x3F248DAF,  # 0.6427869200706482f
        0x3F5C89D9,  # 0.8614784181118011f
    ],
    dtype=np.uint32,
).view(np.float32)


def _as_f32(w) -> np.ndarray:
    """Flatten any array-like (numpy or torch, any float dtype) to float32."""
    if hasattr(w, "detach"):  # torch tensor; bf16 has no numpy dtype
        w = w.detach().to("cpu").float().numpy()
    return np.ascontiguousarray(np.asarray(w, dtype=np.float32).reshape(-1))


def nf4_quantize(w, blocksize: int = BLOCKSIZE):
    """Blockwise NF4 encode + pack.

    Returns ``(packed uint8[(n+1)//2, 1], absmax float32[ceil(n/blocksize)])``
    exactly as ``bitsandbytes.functional.quantize_4bit`` lays them out
    (call site: /root/reference/src/modules/quant/functional.py:362-366).
    """
    x = _as_f32(w)
    n = x.size
    nblocks = (n + blocksize - 1) // blocksize
    pad = nblocks * blocksize - n
    xp = np.concatenate([x, np.zeros(pad, np.float32)]) if pad else x
    blocks = xp.reshape(nblocks, blocksize)
    absmax = np.abs(blocks).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        scale = (np.float32(1.0) / absmax).astype(np.float32)
        scaled = (blocks * scale[:, None]).astype(np.float32)
    # number of thresholds strictly below x  (NaN -> 0, matching the '>' tree)
    flat = scaled.reshape(-1)
    codes = np.searchsorted(NF4_THRESHOLDS, flat, side="left").astype(np.uint8)
    codes[np.isnan(flat)] = 0
    codes = codes[:n]
    if n % 2:
        codes = np.concatenate([codes, np.zeros(1, np.uint8)])
    packed = ((codes[0::2] << 4) | codes[1::2]).astype(np.uint8).reshape(-1, 1)
    return packed, absmax


def nf4_unpack(packed, n: int) -> np.ndarray:
    p = np.asarray(packed, dtype=np.uint8).reshape(-1)
    codes = np.empty(p.size * 2, np.uint8)
    codes[0::2] = p >> 4
    codes[1::2] = p & 0xF
    return codes[:n]


def nf4_dequantize_f32(packed, absmax, n: int, blocksize: int = BLOCKSIZE) -> np.ndarray:
    """fp32 product ``CODEBOOK[code] * absmax[i // blocksize]`` (before the
    rounding to ``quant_state.dtype`` that bitsandbytes applies)."""
    codes = nf4_unpack(packed, n)
    am = np.asarray(absmax, dtype=np.float32).reshape(-1)
    idx = np.arange(n) // blocksize
    return (NF4_CODEBOOK[codes] * am[idx]).astype(np.float32)


def nf4_dequantize(packed, absmax, shape, dtype: str = "bfloat16", blocksize: int = BLOCKSIZE):
    """Dequantize to a torch tensor of ``dtype`` (the stored ``quant_state.dtype``)
    -- one rounding fp32 -> dtype, as bitsandbytes' kDequantizeBlockwise does."""
    import torch

    n = int(np.prod(shape))
    f = nf4_dequantize_f32(packed, absmax, n, blocksize)
    tdt = {"bfloat16": torch.bfloat16, "float16": torch.float16, "float32": torch.float32}[dtype]
    return torch.from_numpy(f).to(tdt).reshape(tuple(shape))


# ---------------------------------------------------------------------------
# quant_state (de)serialisation: bitsandbytes ``QuantState.as_dict(packed=True)``
# key names pinned by /root/reference/tests/test_modules_quant.py:54,183 and
# corroborated by transformers/quantizers/quantizer_bnb_4bit.py.
# ---------------------------------------------------------------------------
def pack_quant_state_blob(shape, dtype: str, blocksize: int = BLOCKSIZE, nested: dict | None = None) -> np.ndarray:
    meta = {"quant_type": "nf4", "blocksize": int(blocksize), "dtype": dtype, "shape": [int(s) for s in shape]}
    if nested:
        meta.update(nested)
    return np.frombuffer(json.dumps(meta).encode("utf-8"), dtype=np.uint8).copy()


def unpack_quant_state_blob(blob) -> dict:
    if hasattr(blob, "detach"):
        blob = blob.detach().cpu().numpy()
    return json.loads(bytes(np.asarray(blob, dtype=np.uint8)).decode("utf-8"))


# ---------------------------------------------------------------------------
# Nested ("double quant") block statistics -- PARITY UNPINNED like the rest of this file.
# bitsandbytes.functional.quantize_4bit(compress_statistics=True), the path Params4bit.cuda() takes under the
# reference's default (/root/reference/src/modules/quant/bnb.py:44,122-129; tools/quantize_model.py:33-54):
#     offset = absmax.mean(); absmax -= offset
#     qabsmax, state2 = quantize_blockwise(absmax, blocksize=256)     # 8-bit "dynamic" map, kQuantizeBlockwise
# and on the way back  absmax = dequantize_blockwise(qabsmax, state2) + offset.
# ---------------------------------------------------------------------------
NESTED_BLOCKSIZE = 256


def dynamic_map(signed: bool = True, max_exponent_bits: int = 7, total_bits: int = 8) -> np.ndarray:
    """bitsandbytes.functional.create_dynamic_map restated: the 256-entry ascending 8-bit code of the nested
    statistics (``nested_quant_map`` in a checkpoint).  torch.linspace in fp32, like the original."""
    import torch

    data: list[float] = []
    non_sign_bits = total_bits - 1
    additional_items = 2 ** (non_sign_bits - max_exponent_bits) - 1
    i = 0
    for i in range(max_exponent_bits):
        fraction_items = int(
            2 ** (i + non_sign_bits - max_exponent_bits) + 1 if signed else 2 ** (i + non_sign_bits - max_exponent_bits + 1) + 1
        )
        boundaries = torch.linspace(0.1, 1, fraction_items, dtype=torch.float32)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    if additional_items > 0:
        boundaries = torch.linspace(0.1, 1, additional_items + 1, dtype=torch.float32)
        means = (boundaries[:-1] + boundaries[1:]) / 2.0
        data += ((10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
        if signed:
            data += (-(10 ** (-(max_exponent_bits - 1) + i)) * means).tolist()
    data.append(0)
    data.append(1.0)
    assert len(data) == 2**total_bits
    data.sort()
    return torch.tensor(data, dtype=torch.float32).numpy()


def _dquantize8(code: np.ndarray, x: np.float32) -> int:
    """bitsandbytes' ``dQuantize<0>`` (kernels.cu), scalar restatement: seven-step bisection from pivot 127, then the
    nearer of the pivot and the neighbour on x's side, all comparisons strict, fp32 midpoints."""
    pivot, upper_pivot, lower_pivot = 127, 255, 0
    lower, upper = np.float32(-1.0), np.float32(1.0)
    val = code[pivot]
    i = 64
    while i > 0:
        if x > val:
            lower_pivot, lower = pivot, val
            pivot += i
        else:
            upper_pivot, upper = pivot, val
            pivot -= i
        val = code[pivot]
        i >>= 1
    if upper_pivot == 255:
        upper = code[upper_pivot]
    if lower_pivot == 0:
        lower = code[lower_pivot]
    if x > val:
        mid = np.float32(np.float32(upper + val) * np.float32(0.5))
        return upper_pivot if x > mid else pivot
    mid = np.float32(np.float32(lower + val) * np.float32(0.5))
    return lower_pivot if x < mid else pivot


def dquantize8(code: np.ndarray, x: np.ndarray) -> np.ndarray:
    """Vectorised form of :func:`_dquantize8` (same decisions; checked against the scalar loop in tests)."""
    code = np.asarray(code, np.float32)
    x = np.asarray(x, np.float32).reshape(-1)
    n = x.size
    pivot = np.full(n, 127, np.int64)
    upper_pivot = np.full(n, 255, np.int64)
    lower_pivot = np.zeros(n, np.int64)
    lower = np.full(n, -1.0, np.float32)
    upper = np.full(n, 1.0, np.float32)
    val = code[pivot]
    i = 64
    with np.errstate(invalid="ignore"):
        while i > 0:
            gt = x > val
            lower_pivot = np.where(gt, pivot, lower_pivot)
            lower = np.where(gt, val, lower)
            upper_pivot = np.where(gt, upper_pivot, pivot)
            upper = np.where(gt, upper, val)
            pivot = np.where(gt, pivot + i, pivot - i)
            val = code[pivot]
            i >>= 1
        upper = np.where(upper_pivot == 255, code[255], upper).astype(np.float32)
        lower = np.where(lower_pivot == 0, code[0], lower).astype(np.float32)
        gt = x > val
        mid_hi = ((upper + val).astype(np.float32) * np.float32(0.5)).astype(np.float32)
        mid_lo = ((lower + val).astype(np.float32) * np.float32(0.5)).astype(np.float32)
        out = np.where(gt, np.where(x > mid_hi, upper_pivot, pivot), np.where(x < mid_lo, lower_pivot, pivot))
    return out.astype(np.uint8)


def absmax_nest(absmax, code: np.ndarray | None = None, blocksize2: int = NESTED_BLOCKSIZE, offset=None):
    """fp32 absmax -> (absmax8 uint8[n], absmax2 fp32[ceil(n/256)], offset fp32, code fp32[256]).

    ``offset`` is the correctly rounded fp32 mean (float64 accumulation).  bitsandbytes takes torch's fp32
    ``absmax.mean()`` whose summation order depends on the device and the torch build; the float64 mean is the value
    all of those approximate, and is what the CUDA kernel computes (csrc/absmax_nest.cu)."""
    a = np.ascontiguousarray(np.asarray(absmax, np.float32).reshape(-1))
    code = dynamic_map() if code is None else np.asarray(code, np.float32)
    n = a.size
    # ``offset`` given: encode around that value (the module path passes torch's ``absmax.mean()``, as bitsandbytes does)
    offset = np.float32(a.astype(np.float64).sum() / n) if offset is None else np.float32(offset)
    v = (a - offset).astype(np.float32)
    nb = (n + blocksize2 - 1) // blocksize2
    pad = nb * blocksize2 - n
    vp = np.concatenate([v, np.zeros(pad, np.float32)]) if pad else v
    blocks = vp.reshape(nb, blocksize2)
    absmax2 = np.abs(blocks).max(axis=1).astype(np.float32)
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = (np.float32(1.0) / absmax2).astype(np.float32)
        scaled = (blocks * inv[:, None]).astype(np.float32)
    q = dquantize8(code, scaled.reshape(-1))[:n]
    return q, absmax2, offset, code


def absmax_denest(absmax8, absmax2, offset, code: np.ndarray, blocksize2: int = NESTED_BLOCKSIZE) -> np.ndarray:
    """``dequantize_blockwise(absmax8, state2) + offset``: code[q] * absmax2[i // 256] rounded to fp32, then + offset
    rounded to fp32 (formula corroborated by vllm's bitsandbytes loader)."""
    q = np.asarray(absmax8, np.uint8).reshape(-1)
    s = np.asarray(absmax2, np.float32).reshape(-1)
    code = np.asarray(code, np.float32)
    idx = np.arange(q.size) // blocksize2
    prod = (code[q] * s[idx]).astype(np.float32)
    return (prod + np.float32(offset)).astype(np.float32)


def quant_state_absmax_f32(quant_state) -> np.ndarray:
    """fp32 block statistics a (possibly nested) module-level quant state decodes to -- checker-side restatement of
    what Linear4bit feeds the kernels.  ``quant_state``: the product's QuantState (read-only: tensors -> numpy)."""
    a = quant_state.absmax.detach().cpu().numpy()
    if not getattr(quant_state, "nested", False):
        return a.astype(np.float32)
    s2 = quant_state.state2
    return absmax_denest(a, s2.absmax.detach().cpu().numpy(), np.float32(float(quant_state.offset)),
                         s2.code.detach().cpu().numpy(), int(s2.blocksize))
