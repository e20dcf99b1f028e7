"""torch-facing wrappers over the C ABI (libvft_b200.so).

torch is used for device memory, streams and autograd bookkeeping only; every
arithmetic step of the hot path runs in the hand-written CUDA kernels.  There is
no CPU path: tensors that are not on a CUDA device raise.
"""
from __future__ import annotations

import torch

from . import _cabi
from ._cabi import lib, check

LORA_LD = _cabi.LORA_LD
_DT = {torch.float32: _cabi.F32, torch.float16: _cabi.F16, torch.bfloat16: _cabi.BF16}
_DT_NAME = {"float32": _cabi.F32, "float16": _cabi.F16, "bfloat16": _cabi.BF16}
_TORCH_DT = {v: k for k, v in _DT.items()}


def dtype_code(dt) -> int:
    if isinstance(dt, str):
        return _DT_NAME[dt.replace("torch.", "")]
    try:
        return _DT[dt]
    except KeyError:
        raise TypeError(f"vft_b200 supports float32/float16/bfloat16, got {dt}") from None


def _require_cuda(*tensors: torch.Tensor | None) -> torch.device:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError(
                "vft_b200: the NF4/LoRA hot path runs on CUDA (sm_100a) only and has no CPU fallback; "
                f"got a tensor on {t.device}"
            )
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError(f"vft_b200: tensors on different devices ({dev} vs {t.device})")
    assert dev is not None
    return dev


class _NoGuard:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_GUARD = _NoGuard()


def _on_device(dev: torch.device):
    """Device guard only when the tensors do not live on the current device (one process per GPU: the common case is
    a no-op, and torch.cuda.device() costs ~10 us of host time per call on the hot path)."""
    if dev.index is None or dev.index == torch.cuda.current_device():
        return _NO_GUARD
    return torch.cuda.device(dev)


_WS_BYTES: dict[tuple, int] = {}


def _workspace_bytes(op: int, T: int, N: int, K: int, r: int) -> int:
    key = (op, T, N, K, r)
    n = _WS_BYTES.get(key)
    if n is None:
        n = _WS_BYTES[key] = int(lib.vft_workspace_bytes(op, T, N, K, r))
    return n


def bt_rows(r: int) -> int:
    """Rows of bt_save = s * lora_up^T: the rank padded to the MMA's K step (include/vft_b200.h)."""
    return 16 * ((int(r) + 15) // 16)


def _ptr(t: torch.Tensor | None) -> int | None:
    return None if t is None else t.data_ptr()


def _stream() -> int:
    # raw handle of the current stream of the current device: torch.cuda.current_stream() builds a Python Stream
    # object through several layers (~19 us per call on the hot path, measured with cProfile on the GPU box)
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


def last_path() -> int:
    return lib.vft_last_path()


def force_path(path: int) -> None:
    lib.vft_force_path(path)


# ----------------------------------------------------------------------------- NF4 quantize / dequantize
def nf4_quantize(w: torch.Tensor, blocksize: int = 64) -> tuple[torch.Tensor, torch.Tensor]:
    """Blockwise NF4 encode + pack of a CUDA tensor -> (packed uint8 [(n+1)//2, 1], absmax fp32 [ceil(n/bs)])."""
    dev = _require_cuda(w)
    w = w.contiguous()
    n = w.numel()
    packed = torch.empty(((n + 1) // 2, 1), dtype=torch.uint8, device=dev)
    absmax = torch.empty(((n + blocksize - 1) // blocksize,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.vft_nf4_quantize(w.data_ptr(), dtype_code(w.dtype), n, blocksize, packed.data_ptr(), absmax.data_ptr(), _stream()))
    return packed, absmax


def nf4_quantize_many(ws: "list[torch.Tensor]", blocksize: int = 64) -> "list[tuple[torch.Tensor, torch.Tensor]]":
    """``nf4_quantize`` for a list of CUDA tensors of ONE dtype on one device: up to 96 equal-size tensors per
    launch (``vft_nf4_quantize_many``), bit-identical to the per-tensor calls.  This is what a checkpoint wants: the small
    weights of a model otherwise spend more time in launch latency than in the kernel."""
    import ctypes

    if not ws:
        return []
    dev = _require_cuda(*ws)
    dt = ws[0].dtype
    if any(w.dtype != dt or w.device != dev for w in ws):
        raise ValueError("nf4_quantize_many: all tensors must share dtype and device")
    ws = [w.contiguous() for w in ws]
    outs = []
    for w in ws:
        n = w.numel()
        outs.append((torch.empty(((n + 1) // 2, 1), dtype=torch.uint8, device=dev),
                     torch.empty(((n + blocksize - 1) // blocksize,), dtype=torch.float32, device=dev)))
    k = len(ws)
    src = (ctypes.c_void_p * k)(*[w.data_ptr() for w in ws])
    ns = (ctypes.c_int64 * k)(*[w.numel() for w in ws])
    pk = (ctypes.c_void_p * k)(*[o[0].data_ptr() for o in outs])
    am = (ctypes.c_void_p * k)(*[o[1].data_ptr() for o in outs])
    with torch.cuda.device(dev):
        check(lib.vft_nf4_quantize_many(k, src, dtype_code(dt), ns, blocksize, pk, am, _stream()))
    return outs


def nf4_dequantize(packed: torch.Tensor, absmax: torch.Tensor, shape, dtype: torch.dtype, blocksize: int = 64) -> torch.Tensor:
    dev = _require_cuda(packed, absmax)
    n = 1
    for s in shape:
        n *= int(s)
    out = torch.empty(tuple(shape), dtype=dtype, device=dev)
    packed = packed.contiguous()
    absmax = absmax.contiguous().float()
    with torch.cuda.device(dev):
        check(lib.vft_nf4_dequantize(packed.data_ptr(), absmax.data_ptr(), n, blocksize, out.data_ptr(), dtype_code(dtype), _stream()))
    return out


def absmax_nest(absmax: torch.Tensor, code256: torch.Tensor, blocksize2: int = 256, offset: torch.Tensor | None = None):
    """Nested statistics encode: fp32 absmax -> (absmax8 uint8 [n], absmax2 fp32 [ceil(n/256)], offset fp32 0-dim).
    ``offset``: a 0-dim fp32 device tensor to encode around (bitsandbytes: ``absmax.mean()``); None = the correctly
    rounded mean, accumulated in fp64 in a fixed order by the library."""
    dev = _require_cuda(absmax, code256)
    absmax = absmax.contiguous().float()
    code256 = code256.contiguous().float()
    if code256.numel() != 256:
        raise ValueError("the nested code map must have 256 entries")
    n = absmax.numel()
    absmax8 = torch.empty((n,), dtype=torch.uint8, device=dev)
    absmax2 = torch.empty(((n + blocksize2 - 1) // blocksize2,), dtype=torch.float32, device=dev)
    if offset is not None:
        offset = offset.detach().to(device=dev, dtype=torch.float32).reshape(()).contiguous()
        with torch.cuda.device(dev):
            check(lib.vft_absmax_nest_at(absmax.data_ptr(), n, blocksize2, code256.data_ptr(), offset.data_ptr(),
                                         absmax8.data_ptr(), absmax2.data_ptr(), _stream()))
        return absmax8, absmax2, offset
    offset = torch.empty((), dtype=torch.float32, device=dev)
    ws_bytes = lib.vft_workspace_bytes(_cabi.OP_ABSMAX_NEST, 0, 0, 0, 0)
    ws = torch.empty((ws_bytes,), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        check(lib.vft_absmax_nest(absmax.data_ptr(), n, blocksize2, code256.data_ptr(), absmax8.data_ptr(),
                                  absmax2.data_ptr(), offset.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return absmax8, absmax2, offset


def absmax_denest(absmax8: torch.Tensor, absmax2: torch.Tensor, code256: torch.Tensor, offset: float,
                  blocksize2: int = 256) -> torch.Tensor:
    """Nested statistics decode: code256[absmax8] * absmax2[i // blocksize2] + offset -> fp32 [n]."""
    dev = _require_cuda(absmax8, absmax2, code256)
    absmax8 = absmax8.contiguous().reshape(-1)
    if absmax8.dtype != torch.uint8:
        raise TypeError(f"nested absmax must be uint8, got {absmax8.dtype}")
    n = absmax8.numel()
    out = torch.empty((n,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.vft_absmax_denest(absmax8.data_ptr(), absmax2.contiguous().float().data_ptr(),
                                    code256.contiguous().float().data_ptr(), float(offset), n, blocksize2,
                                    out.data_ptr(), _stream()))
    return out


def nf4_quantize_host(w: torch.Tensor, blocksize: int = 64) -> tuple[torch.Tensor, torch.Tensor]:
    """Host-buffer entry (H2D + kernel + D2H inside the C call): what quantize_state_dict amounts to per tensor."""
    if w.is_cuda:
        raise RuntimeError("nf4_quantize_host expects a host tensor")
    w = w.contiguous()
    n = w.numel()
    packed = torch.empty(((n + 1) // 2, 1), dtype=torch.uint8)
    absmax = torch.empty(((n + blocksize - 1) // blocksize,), dtype=torch.float32)
    check(lib.vft_nf4_quantize_host(w.data_ptr(), dtype_code(w.dtype), n, blocksize, packed.data_ptr(), absmax.data_ptr()))
    return packed, absmax


def nf4_tile_weight(packed: torch.Tensor, absmax: torch.Tensor, out_features: int, in_features: int,
                    blocksize: int = 64) -> tuple[torch.Tensor, torch.Tensor] | None:
    """Kernel-friendly 64x64 micro-tiled copy of a packed weight (see csrc/nf4_quant.cu); None when the layout does
    not apply (blocksize != 64 or in_features % 64 != 0).  Derived data: never part of a checkpoint."""
    dev = _require_cuda(packed, absmax)
    N, K = int(out_features), int(in_features)
    if blocksize != 64 or K % 64 != 0 or packed.data_ptr() % 16 != 0:
        return None
    codes_t = torch.empty((lib.vft_nf4_tiled_bytes(N, K, 0),), dtype=torch.uint8, device=dev)
    absmax_t = torch.empty((lib.vft_nf4_tiled_bytes(N, K, 1) // 4,), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        check(lib.vft_nf4_tile_weight(packed.data_ptr(), absmax.contiguous().float().data_ptr(), N, K, blocksize,
                                      codes_t.data_ptr(), absmax_t.data_ptr(), _stream()))
    return codes_t, absmax_t


def _workspace(op: int, T: int, N: int, K: int, r: int, dev) -> tuple[torch.Tensor | None, int]:
    """Scratch the C side asks for (split-K partial sums of small problems); allocated from torch's caching allocator."""
    n = _workspace_bytes(op, T, N, K, r)
    if n <= 0:
        return None, 0
    return torch.empty((n,), dtype=torch.uint8, device=dev), n


# ----------------------------------------------------------------------------- fused layer
class QLoRALinearFunction(torch.autograd.Function):
    """y = x . W~^T (+bias) + scale * (x . A^T) . B^T with W~ decoded from NF4 inside the GEMM.

    Replaces bitsandbytes.matmul_4bit (MatMul4Bit) under LoRALinear.forward
    (/root/reference/src/modules/peft/lora.py:92-104).  The base weight is frozen
    (/root/reference/src/modules/quant/functional.py:115), so only dX, dA, dB exist.
    """

    @staticmethod
    def forward(ctx, x, packed, absmax, bias, lora_a, lora_b, scale, out_features, in_features, blocksize, qdtype,
                tiled=None):
        dev = _require_cuda(x, packed, absmax, bias, lora_a, lora_b)
        codes_t, absmax_t = tiled if tiled is not None else (None, None)
        N, K = int(out_features), int(in_features)
        if x.shape[-1] != K:
            raise RuntimeError(f"input feature size {x.shape[-1]} does not match in_features {K}")
        act = dtype_code(x.dtype)
        x2 = x.reshape(-1, K)
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        T = x2.shape[0]
        r = 0 if lora_a is None else int(lora_a.shape[0])
        if r:
            if lora_a.dtype != x.dtype or lora_b.dtype != x.dtype:
                raise RuntimeError("fused LoRA needs adapter weights in the activation dtype")
            if not lora_a.is_contiguous():
                lora_a = lora_a.contiguous()
            if not lora_b.is_contiguous():
                lora_b = lora_b.contiguous()
        if bias is not None and (bias.dtype != x.dtype or not bias.is_contiguous()):
            bias = bias.to(x.dtype).contiguous()
        y = torch.empty((*x.shape[:-1], N), dtype=x.dtype, device=dev)  # final shape: no view between us and autograd
        t_save = torch.empty((T, LORA_LD), dtype=x.dtype, device=dev) if r else None
        # s * B^T, K-major: lets the backward launch compute dt = s * dy . B itself (include/vft_b200.h)
        bt_save = torch.empty((bt_rows(r), N), dtype=x.dtype, device=dev) if r and T > 0 else None
        # t^T: with it the backward launch computes dA, dB as well (one launch for the whole backward)
        tt_save = torch.empty((bt_rows(r), T), dtype=x.dtype, device=dev) if r and T > 0 else None
        ws, ws_bytes = _workspace(_cabi.OP_FWD, T, N, K, r, dev)
        with _on_device(dev):
            if T > 0:  # an empty batch (ragged bucket on one rank) launches nothing
                check(
                    lib.vft_qlora_fwd(
                        x2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, act, dtype_code(qdtype),
                        _ptr(bias), _ptr(lora_a), _ptr(lora_b), r, float(scale), y.data_ptr(), _ptr(t_save), _ptr(bt_save),
                        _ptr(tt_save), _ptr(ws), ws_bytes, _ptr(codes_t), _ptr(absmax_t), _stream(),
                    )
                )
        ctx.meta = (N, K, blocksize, act, dtype_code(qdtype), r, float(scale), x.shape)
        ctx.tiled = (codes_t, absmax_t, bt_save, tt_save)  # frozen derived buffers, not autograd-tracked
        ctx.save_for_backward(x2 if r else None, packed, absmax, lora_a, lora_b, t_save)
        return y

    @staticmethod
    def backward(ctx, dy):
        N, K, blocksize, act, qd, r, scale, x_shape = ctx.meta
        x2, packed, absmax, lora_a, lora_b, t_save = ctx.saved_tensors
        codes_t, absmax_t, bt_save, tt_save = ctx.tiled
        dev = dy.device
        dy2 = dy.reshape(-1, N)
        if dy2.dtype != _TORCH_DT[act]:
            dy2 = dy2.to(_TORCH_DT[act])
        if not dy2.is_contiguous():
            dy2 = dy2.contiguous()
        T = dy2.shape[0]
        need_dx = ctx.needs_input_grad[0]
        need_ab = r > 0 and (ctx.needs_input_grad[4] or ctx.needs_input_grad[5])
        # a trainable bias (LoRA use_bias: s * lora_up.bias rides the epilogue's bias): column sum of dy, fp32 accumulation
        dbias = dy2.sum(0, dtype=torch.float32).to(dy2.dtype) if ctx.needs_input_grad[3] else None
        # allocated in the input's shape: a reshaped view would make AccumulateGrad clone it (25 MB at config #1)
        dx = torch.empty(x_shape, dtype=dy2.dtype, device=dev) if need_dx else None
        dt_save = torch.empty((T, LORA_LD), dtype=dy2.dtype, device=dev) if r else None
        da = db = None
        if T == 0:  # empty batch: no launches; the adapter gradients of an empty sum are zeros
            if need_ab:
                da, db = torch.zeros_like(lora_a), torch.zeros_like(lora_b)
            return dx, None, None, dbias, da, db, None, None, None, None, None, None
        if need_dx and need_ab:
            # the whole backward in one C-ABI call (one launch when the persistent tcgen05 kernel takes it)
            da = torch.empty_like(lora_a)
            db = torch.empty_like(lora_b)
            ws, ws_bytes = _workspace(_cabi.OP_BWD, T, N, K, r, dev)
            with _on_device(dev):
                check(
                    lib.vft_qlora_bwd(
                        dy2.data_ptr(), x2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, act, qd,
                        _ptr(lora_a), _ptr(lora_b), r, scale, t_save.data_ptr(), _ptr(tt_save), _ptr(bt_save), _ptr(dx),
                        da.data_ptr(), db.data_ptr(), dt_save.data_ptr(), _ptr(ws), ws_bytes, _ptr(codes_t), _ptr(absmax_t),
                        _stream(),
                    )
                )
            return dx, None, None, dbias, da, db, None, None, None, None, None, None
        with _on_device(dev):
            if need_dx or need_ab:
                ws, ws_bytes = _workspace(_cabi.OP_BWD_DX, T, N, K, r, dev) if need_dx else (None, 0)
                check(
                    lib.vft_qlora_bwd_dx(
                        dy2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, act, qd,
                        _ptr(lora_a), _ptr(lora_b), r, scale, _ptr(dx), _ptr(dt_save), _ptr(bt_save), _ptr(ws), ws_bytes,
                        _ptr(codes_t), _ptr(absmax_t), _stream(),
                    )
                )
            if need_ab:
                da = torch.empty_like(lora_a)
                db = torch.empty_like(lora_b)
                ws_bytes = _workspace_bytes(_cabi.OP_BWD_DAB, T, N, K, r)
                ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
                check(
                    lib.vft_lora_bwd_dab(
                        dy2.data_ptr(), x2.data_ptr(), t_save.data_ptr(), dt_save.data_ptr(), T, N, K, r, act, scale,
                        da.data_ptr(), db.data_ptr(), ws.data_ptr(), ws_bytes, _stream(),
                    )
                )
        return dx, None, None, dbias, da, db, None, None, None, None, None, None


# ----------------------------------------------------------------------------- torch.library registration
# The reference compiles its denoiser with torch.compile(fullgraph=True) (configs/auraflow/lora.yml:82-85,
# /root/reference/src/models/for_training.py:60-65).  An autograd.Function whose forward goes through ctypes cannot be
# traced without a graph break, so the same two C-ABI call sequences are ALSO registered as custom operators with fake
# (meta) implementations and an autograd formula; qlora_linear() switches to them while a compiler is tracing and keeps
# the leaner autograd.Function for eager calls (an operator dispatch costs tens of microseconds of host time per call).
def _empty(like: torch.Tensor) -> torch.Tensor:
    return like.new_empty((0,))


@torch.library.custom_op("vft_b200::qlora_fwd", mutates_args=())
def _qlora_fwd_op(x: torch.Tensor, packed: torch.Tensor, absmax: torch.Tensor, bias: torch.Tensor | None,
                  lora_a: torch.Tensor | None, lora_b: torch.Tensor | None, scale: float, out_features: int,
                  in_features: int, blocksize: int, qdtype: int, codes_t: torch.Tensor | None,
                  absmax_t: torch.Tensor | None) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = _require_cuda(x, packed, absmax, bias, lora_a, lora_b)
    N, K = out_features, in_features
    if x.shape[-1] != K:
        raise RuntimeError(f"input feature size {x.shape[-1]} does not match in_features {K}")
    x2 = x.reshape(-1, K).contiguous()
    T = x2.shape[0]
    r = 0 if lora_a is None else int(lora_a.shape[0])
    if r and (lora_a.dtype != x.dtype or lora_b.dtype != x.dtype):
        raise RuntimeError("fused LoRA needs adapter weights in the activation dtype")
    la = lora_a.contiguous() if r else None
    lb = lora_b.contiguous() if r else None
    if bias is not None:
        bias = bias.to(x.dtype).contiguous()
    y = torch.empty((*x.shape[:-1], N), dtype=x.dtype, device=dev)
    # (zeros: the kernels write the first 16 * ceil(r / 16) columns only, and an operator's outputs must be reproducible)
    t_save = torch.zeros((T, LORA_LD), dtype=x.dtype, device=dev) if r else _empty(x)
    bt_save = torch.empty((bt_rows(r), N), dtype=x.dtype, device=dev) if r else _empty(x)
    tt_save = torch.empty((bt_rows(r), T), dtype=x.dtype, device=dev) if r else _empty(x)
    ws, ws_bytes = _workspace(_cabi.OP_FWD, T, N, K, r, dev)
    with _on_device(dev):
        if T > 0:
            check(lib.vft_qlora_fwd(x2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, dtype_code(x.dtype),
                                    qdtype, _ptr(bias), _ptr(la), _ptr(lb), r, float(scale), y.data_ptr(),
                                    t_save.data_ptr() if r else None, bt_save.data_ptr() if r else None,
                                    tt_save.data_ptr() if r else None, _ptr(ws), ws_bytes,
                                    _ptr(codes_t), _ptr(absmax_t), _stream()))
    return y, t_save, bt_save, tt_save


@_qlora_fwd_op.register_fake
def _(x, packed, absmax, bias, lora_a, lora_b, scale, out_features, in_features, blocksize, qdtype, codes_t, absmax_t):
    T = x.numel() // in_features
    y = x.new_empty((*x.shape[:-1], out_features))
    t_save = x.new_empty((T, LORA_LD)) if lora_a is not None else x.new_empty((0,))
    bt_save = x.new_empty((bt_rows(lora_a.shape[0]), out_features)) if lora_a is not None else x.new_empty((0,))
    tt_save = x.new_empty((bt_rows(lora_a.shape[0]), T)) if lora_a is not None else x.new_empty((0,))
    return y, t_save, bt_save, tt_save


@torch.library.custom_op("vft_b200::qlora_bwd", mutates_args=())
def _qlora_bwd_op(dy: torch.Tensor, x: torch.Tensor, packed: torch.Tensor, absmax: torch.Tensor,
                  lora_a: torch.Tensor | None, lora_b: torch.Tensor | None, t_save: torch.Tensor, bt_save: torch.Tensor,
                  tt_save: torch.Tensor, scale: float, out_features: int, in_features: int, blocksize: int, qdtype: int, codes_t: torch.Tensor | None,
                  absmax_t: torch.Tensor | None, need_dx: bool, need_ab: bool) -> tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    dev = dy.device
    N, K = out_features, in_features
    dy2 = dy.reshape(-1, N).to(x.dtype).contiguous()
    x2 = x.reshape(-1, K).contiguous()
    T = dy2.shape[0]
    r = 0 if lora_a is None else int(lora_a.shape[0])
    need_ab = need_ab and r > 0
    la = lora_a.contiguous() if r else None
    lb = lora_b.contiguous() if r else None
    dx = torch.empty(x.shape, dtype=x.dtype, device=dev) if need_dx else _empty(x)
    da = torch.zeros_like(la) if need_ab else _empty(x)
    db = torch.zeros_like(lb) if need_ab else _empty(x)
    if T == 0 or not (need_dx or need_ab):
        return dx, da, db
    dt_save = torch.empty((T, LORA_LD), dtype=x.dtype, device=dev) if r else None
    act = dtype_code(x.dtype)
    if need_dx and need_ab:
        ws, ws_bytes = _workspace(_cabi.OP_BWD, T, N, K, r, dev)
        with _on_device(dev):
            check(lib.vft_qlora_bwd(dy2.data_ptr(), x2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, act,
                                    qdtype, _ptr(la), _ptr(lb), r, float(scale), t_save.data_ptr(), tt_save.data_ptr(),
                                    bt_save.data_ptr(), dx.data_ptr(), da.data_ptr(), db.data_ptr(), dt_save.data_ptr(),
                                    _ptr(ws), ws_bytes, _ptr(codes_t), _ptr(absmax_t), _stream()))
        return dx, da, db
    with _on_device(dev):
        ws, ws_bytes = _workspace(_cabi.OP_BWD_DX, T, N, K, r, dev) if need_dx else (None, 0)
        check(lib.vft_qlora_bwd_dx(dy2.data_ptr(), T, packed.data_ptr(), absmax.data_ptr(), N, K, blocksize, act, qdtype,
                                   _ptr(la), _ptr(lb), r, float(scale), dx.data_ptr() if need_dx else None, _ptr(dt_save),
                                   bt_save.data_ptr() if r else None, _ptr(ws), ws_bytes, _ptr(codes_t), _ptr(absmax_t), _stream()))
        if need_ab:
            ws_bytes = _workspace_bytes(_cabi.OP_BWD_DAB, T, N, K, r)
            ws = torch.empty((max(ws_bytes, 4),), dtype=torch.uint8, device=dev)
            check(lib.vft_lora_bwd_dab(dy2.data_ptr(), x2.data_ptr(), t_save.data_ptr(), dt_save.data_ptr(), T, N, K, r, act,
                                       float(scale), da.data_ptr(), db.data_ptr(), ws.data_ptr(), ws_bytes, _stream()))
    return dx, da, db


@_qlora_bwd_op.register_fake
def _(dy, x, packed, absmax, lora_a, lora_b, t_save, bt_save, tt_save, scale, out_features, in_features, blocksize, qdtype,
      codes_t, absmax_t, need_dx, need_ab):
    dx = x.new_empty(x.shape) if need_dx else x.new_empty((0,))
    ab = need_ab and lora_a is not None
    da = lora_a.new_empty(lora_a.shape) if ab else x.new_empty((0,))
    db = lora_b.new_empty(lora_b.shape) if ab else x.new_empty((0,))
    return dx, da, db


def _op_setup_context(ctx, inputs, output):
    x, packed, absmax, bias, lora_a, lora_b, scale, N, K, blocksize, qdtype, codes_t, absmax_t = inputs
    _, t_save, bt_save, tt_save = output
    ctx.meta = (float(scale), N, K, blocksize, qdtype)
    ctx.save_for_backward(x, packed, absmax, lora_a, lora_b, t_save, bt_save, tt_save, codes_t, absmax_t)


def _op_backward(ctx, dy, _dt_save, _dbt_save, _dtt_save):
    scale, N, K, blocksize, qdtype = ctx.meta
    x, packed, absmax, lora_a, lora_b, t_save, bt_save, tt_save, codes_t, absmax_t = ctx.saved_tensors
    need_dx = ctx.needs_input_grad[0]
    need_ab = lora_a is not None and (ctx.needs_input_grad[4] or ctx.needs_input_grad[5])
    dx, da, db = _qlora_bwd_op(dy, x, packed, absmax, lora_a, lora_b, t_save, bt_save, tt_save, scale, N, K, blocksize, qdtype,
                               codes_t, absmax_t, need_dx, need_ab)
    dbias = dy.reshape(-1, N).sum(0, dtype=torch.float32).to(dy.dtype) if ctx.needs_input_grad[3] else None
    return (dx if need_dx else None, None, None, dbias, da if need_ab else None, db if need_ab else None,
            None, None, None, None, None, None, None)


_qlora_fwd_op.register_autograd(_op_backward, setup_context=_op_setup_context)


def qlora_linear(x, packed, absmax, bias, lora_a, lora_b, scale, out_features, in_features, blocksize=64,
                 qdtype=torch.bfloat16, tiled=None):
    """``tiled``: optional (codes_t, absmax_t) from :func:`nf4_tile_weight` for the same weight."""
    if torch.compiler.is_compiling():  # traced by torch.compile: the registered operator (no graph break)
        codes_t, absmax_t = tiled if tiled is not None else (None, None)
        y, _, _, _ = _qlora_fwd_op(x, packed, absmax, bias, lora_a, lora_b, float(scale), int(out_features), int(in_features),
                             int(blocksize), dtype_code(qdtype), codes_t, absmax_t)
        return y
    return QLoRALinearFunction.apply(x, packed, absmax, bias, lora_a, lora_b, scale, out_features, in_features,
                                     blocksize, qdtype, tiled)
