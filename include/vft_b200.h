/* vft_b200.h -- C ABI of the B200-native QLoRA hot path (NF4 base Linear + LoRA,
 * forward / backward, and NF4 quantize/pack).
 *
 * This is the boundary a maintainer of p1atdev/vision-ft binds instead of
 * bitsandbytes for this path (see INTEGRATION.md for the ctypes stub).  The
 * reference has no FFI of its own: the path sits behind Python classes, so each
 * entry point cites the reference call site (or the bitsandbytes function that
 * call site reaches) it replaces.
 *
 * Conventions
 *   - plain C, no torch types: raw DEVICE pointers + sizes + an explicit
 *     cudaStream_t passed as void* (NULL = legacy default stream);
 *   - nothing is allocated, nothing is freed; every buffer (outputs, saved LoRA
 *     activations, workspace) belongs to the caller and is borrowed for the
 *     duration of the call on the given stream.  The only process-wide state is
 *     the triage switches at the end of this header (forced kernel family,
 *     VFT_* environment variables read once at first use) and, on the device,
 *     a pool of 4096 self-resetting {arrivals, generation} counters that the
 *     launches with an in-kernel side product take in turn;
 *   - re-entrant and callable from any host thread (autograd worker threads call
 *     the backward entry points); the current CUDA device is the caller's;
 *   - every function returns 0 on success or a negative vft_status; the message
 *     is in vft_last_error() (thread-local);
 *   - there is NO CPU fallback: without a CUDA device the compute entry points
 *     return VFT_ERR_CUDA.
 *
 * Shapes: T tokens (batch*seq, flattened), K in_features, N out_features,
 * r LoRA rank, blocksize 64.  W is [N, K] row-major, quantized over the FLATTENED
 * tensor: packed[(N*K+1)/2] (element 2j in the HIGH nibble of byte j),
 * absmax[ceil(N*K/blocksize)] fp32.
 */
#ifndef VFT_B200_H_
#define VFT_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VFT_ABI_VERSION 7
#define VFT_LORA_LD 64 /* leading dimension (elements) of the saved LoRA activations t_save / dt_save */

enum vft_dtype { VFT_F32 = 0, VFT_F16 = 1, VFT_BF16 = 2 };

enum vft_status {
  VFT_OK = 0,
  VFT_ERR_INVALID = -1,     /* bad argument (null pointer, size, alignment, dtype) */
  VFT_ERR_UNSUPPORTED = -2, /* valid request this build does not implement */
  VFT_ERR_CUDA = -3,        /* CUDA runtime / driver error (no device, launch failure) */
  VFT_ERR_WORKSPACE = -4    /* workspace too small: see vft_workspace_bytes */
};

/* Which kernel family served the last compute call on this thread. */
enum vft_path {
  VFT_PATH_NONE = 0,
  VFT_PATH_TCGEN05 = 1, /* fused dequant -> tcgen05.mma (TMEM accumulators), sm_100a */
  VFT_PATH_SIMT = 2,    /* generic CUDA-core kernels: shapes the tensor path does not take */
  VFT_PATH_GEMV = 3     /* few-token forward (T <= 8): packed-weight streaming kernel, HBM/decode bound */
};

enum vft_op { VFT_OP_FWD = 0, VFT_OP_BWD_DX = 1, VFT_OP_BWD_DAB = 2, VFT_OP_ABSMAX_NEST = 3, VFT_OP_BWD = 4 };

int vft_abi_version(void);
const char* vft_last_error(void);
int vft_last_path(void);
/* Force a kernel family for subsequent calls, process-wide (tests/bench): 0 = auto. */
void vft_force_path(int path);
/* Triage (tests / profiling tools, not thread-safe): re-read the VFT_* environment switches, which are otherwise
 * read once per process; SM-clock timelines of the last launch made with VFT_TC_DEBUG & 16 (rows x 256 / 16 stamps). */
void vft_reload_env(void);
int vft_debug_tc_timeline(unsigned long long* out, int n);  /* one-tile kernel (csrc/qlora_tc.cu): 16 stamps */
int vft_debug_tc2_timeline(unsigned long long* out, int n);
int vft_debug_side_timeline(unsigned long long* out, int n);
/* Raw side-product accumulator lanes of CTA pair 0 of the last launch made with VFT_TC_DEBUG & 512 (the TMEM layout
 * probe of tools/p0_layout_probe.py): out[2 CTAs][128 lanes][32 columns] fp32. */
int vft_debug_tc2_p0dump(float* out, int n);

/* NF4 quantize/pack.  Replaces bitsandbytes.functional.quantize_4bit(quant_type="nf4")
 * as called at /root/reference/src/modules/quant/functional.py:362-365 and, lazily, by
 * Params4bit.cuda() from /root/reference/tools/quantize_model.py:53.
 *   w       [n] device, dtype in {F32,F16,BF16}
 *   packed  [(n+1)/2] device uint8 out;  absmax [ceil(n/blocksize)] device fp32 out
 * Bit-exact contract: absmax = max|float(w)| per block; code = number of NF4
 * thresholds strictly below float(w) * (1.0f/absmax) (IEEE fp32). */
int vft_nf4_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax, void* stream);

/* The same for a batch of tensors of one dtype in as few launches as possible (the tables are HOST arrays of `count`
 * device pointers / element counts; up to 96 tensors of EQUAL element count ride one launch -- a model checkpoint
 * repeats a handful of shapes).  This is the loop of quantize_state_dict()
 * (/root/reference/src/modules/quant/functional.py:342-371) and of tools/quantize_model.py:33-54 over a checkpoint:
 * per-tensor launches leave an HBM-bound kernel waiting on launch latency for the small weights.  Results are
 * bit-identical to `count` calls of vft_nf4_quantize. */
int vft_nf4_quantize_many(int count, const void* const* w, int dtype, const int64_t* n, int blocksize,
                          uint8_t* const* packed, float* const* absmax, void* stream);

/* NF4 dequantize (debug / checker entry; the fused kernels never materialise W).
 * Replaces bitsandbytes.functional.dequantize_4bit.  out[n] in `dtype`. */
int vft_nf4_dequantize(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, void* out, int dtype,
                       void* stream);

/* Same two operations with HOST buffers: copies in, kernel, copies out, synchronises.
 * This is the "reference-facing call with host buffers" used for the e2e numbers and
 * what quantize_state_dict() (functional.py:342-371: .cuda() ... .cpu()) amounts to. */
int vft_nf4_quantize_host(const void* w_host, int dtype, int64_t n, int blocksize, uint8_t* packed_host,
                          float* absmax_host);

/* Optional kernel-friendly copy of a packed weight (blocksize 64, K % 64 == 0): 64 x 64 micro-tiles so that the
 * 32 lanes of a decode warp read contiguous bytes (layout: csrc/nf4_quant.cu).  The checkpoint format -- what
 * Params4bit holds and state_dict() emits, /root/reference/src/modules/quant/bnb.py:91-107 -- is unchanged; this is
 * a derived, caller-owned device buffer that the fused entry points accept next to packed/absmax.
 *   codes_t  [vft_nf4_tiled_bytes(N,K,0)] bytes, 16-byte aligned;  absmax_t [vft_nf4_tiled_bytes(N,K,1)] bytes */
int64_t vft_nf4_tiled_bytes(int64_t N, int64_t K, int which);
int vft_nf4_tile_weight(const uint8_t* packed, const float* absmax, int64_t N, int64_t K, int blocksize,
                        uint8_t* codes_t, float* absmax_t, void* stream);

/* Nested ("double quant") block statistics.  Replaces the tail of bitsandbytes.functional.quantize_4bit(
 * compress_statistics=True) -- offset = absmax.mean(); quantize_blockwise(absmax - offset, blocksize=256) with the
 * 8-bit "dynamic" map -- which is the path Params4bit.cuda() takes under BnbLinear4bit's default
 * (/root/reference/src/modules/quant/bnb.py:44,122-129; /root/reference/tools/quantize_model.py:33-54).
 *   absmax   [nblocks] device fp32 in (16-byte aligned)      code256  [256] device fp32, ascending (the dynamic map)
 *   absmax8  [nblocks] device uint8 out                      absmax2  [ceil(nblocks/256)] device fp32 out
 *   offset   device fp32 scalar out = fp32(sum_fp64(absmax) / nblocks)
 *   ws       vft_workspace_bytes(VFT_OP_ABSMAX_NEST, ...) bytes
 * Bit-exact contract: v = absmax - offset (fp32); absmax2 = max|v| per 256; index = bitsandbytes' dQuantize
 * bisection of v * (1.0f/absmax2) over code256 (nearest entry; a value exactly on a midpoint keeps the pivot). */
int vft_absmax_nest(const float* absmax, int64_t nblocks, int blocksize2, const float* code256, uint8_t* absmax8,
                    float* absmax2, float* offset, void* ws, int64_t ws_bytes, void* stream);

/* Same encode around a GIVEN offset (device fp32 scalar, read on the stream): bitsandbytes takes
 * offset = absmax.mean() from torch on the device, whose summation order is torch's; a caller that wants checkpoints
 * byte-identical to bitsandbytes' computes that scalar the same way and passes it here.  No workspace. */
int vft_absmax_nest_at(const float* absmax, int64_t nblocks, int blocksize2, const float* code256, const float* offset,
                       uint8_t* absmax8, float* absmax2, void* stream);

/* Decode side: absmax_out[i] = code256[absmax8[i]] * absmax2[i / blocksize2] + offset (two fp32 roundings), i.e.
 * bitsandbytes.functional.dequantize_blockwise(absmax8, state2) + offset -- what every dequantize_4bit of a nested
 * checkpoint starts with (prequantized branch, /root/reference/src/modules/quant/bnb.py:91-107). */
int vft_absmax_denest(const uint8_t* absmax8, const float* absmax2, const float* code256, float offset,
                      int64_t nblocks, int blocksize2, float* absmax_out, void* stream);

/* Scratch the caller must provide for an op (bytes; 0 is possible). */
int64_t vft_workspace_bytes(int op, int64_t T, int64_t N, int64_t K, int r);

/* Fused forward.  Replaces Linear4bit.forward -> matmul_4bit -> MatMul4Bit.forward
 * (dequant to a bf16 copy + cuBLAS) reached from
 * /root/reference/src/modules/peft/lora.py:93, plus the adapter arithmetic of
 * lora.py:100-104:
 *     y[T,N] = x[T,K] . W~^T (+ bias) + scale * (x . A^T) . B^T
 *   act_dtype  dtype of x, y, bias, A, B, t_save: VFT_BF16 or VFT_F16
 *   qdtype     quant_state.dtype: W~ is rounded to it first, then to act_dtype
 *   lora_a [r,K] (lora_down.weight), lora_b [N,r] (lora_up.weight); both NULL and
 *   r = 0 for an NF4-only layer.  scale = alpha / rank.
 *   t_save [T, VFT_LORA_LD] out: x . A^T rounded to act_dtype in the first 16*ceil(r/16) columns (columns
 *   [r, 16*ceil(r/16)) are zeros, the ones behind are not written and never read); required when r > 0 (the
 *   backward reads it).
 *   bt_save [16*ceil(r/16), N] out, optional (NULL: not wanted): scale * lora_b^T rounded to act_dtype, rows >= r
 *   zero -- the K-major form of the adapter's up-projection that lets vft_qlora_bwd_dx compute dt inside its launch.
 *   tt_save [16*ceil(r/16), T] out, optional: t_save transposed (rows >= r zero) -- with it vft_qlora_bwd computes
 *   the adapter's weight gradients inside its launch; 16-byte aligned, T % 8 == 0 for that to apply.
 *   codes_t / absmax_t: the micro-tiled copy of the same weight, or both NULL. */
int vft_qlora_fwd(const void* x, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                  int blocksize, int act_dtype, int qdtype, const void* bias, const void* lora_a, const void* lora_b,
                  int r, float scale, void* y, void* t_save, void* bt_save, void* tt_save, void* ws, int64_t ws_bytes,
                  const uint8_t* codes_t, const float* absmax_t, void* stream);

/* Fused backward w.r.t. the input.  Replaces MatMul4Bit.backward (second dequant +
 * cuBLAS) and the dX half of the adapter's autograd:
 *     dt_save[T, VFT_LORA_LD] = scale * dy . B        (rounded to act_dtype)
 *     dx[T,K] = dy[T,N] . W~ + dt . A
 * dx may be NULL when only dt_save is wanted (input does not require grad).
 * bt_save: what the forward call left (see there), or NULL; with it dt is computed inside the GEMM launch
 * (as dy . bt^T: the scale is folded into the 16-bit rows of bt) instead of by a kernel in front of it. */
int vft_qlora_bwd_dx(const void* dy, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                     int blocksize, int act_dtype, int qdtype, const void* lora_a, const void* lora_b, int r,
                     float scale, void* dx, void* dt_save, const void* bt_save, void* ws, int64_t ws_bytes,
                     const uint8_t* codes_t, const float* absmax_t, void* stream);

/* The whole backward of the layer in one call -- MatMul4Bit.backward plus the autograd of lora.py:100-104:
 *     dt_save = scale * dy . B;  dx = dy . W~ + dt . A;  dA = dt^T . x;  dB = scale * dy^T . t
 * With bt_save and tt_save from the forward call (and a shape the persistent tcgen05 kernel takes unsplit) this is ONE
 * launch: the side product and the two token contractions ride the GEMM (csrc/qlora_tc2.cu).  Otherwise it is the two
 * calls below, in order.  ws: vft_workspace_bytes(VFT_OP_BWD, T, N, K, r) bytes, 256-byte aligned.  dx may be NULL. */
int vft_qlora_bwd(const void* dy, const void* x, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                  int blocksize, int act_dtype, int qdtype, const void* lora_a, const void* lora_b, int r, float scale,
                  const void* t_save, const void* tt_save, const void* bt_save, void* dx, void* dA, void* dB,
                  void* dt_save, void* ws, int64_t ws_bytes, const uint8_t* codes_t, const float* absmax_t, void* stream);

/* Adapter weight gradients (autograd of lora.py:100-104):
 *     dA[r,K] = dt^T . x        dB[N,r] = scale * dy^T . t
 * with t_save / dt_save as produced by the two calls above.  dA, dB in act_dtype. */
int vft_lora_bwd_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N,
                     int64_t K, int r, int act_dtype, float scale, void* dA, void* dB, void* ws, int64_t ws_bytes,
                     void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VFT_B200_H_ */
