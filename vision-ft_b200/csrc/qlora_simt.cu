// Generic CUDA-core kernels of the QLoRA layer.
//
// They take every shape (any K, N, T, blocksize, rank <= 64, quantization blocks that span
// rows as in AuraFlow's init_x_linear with K = 16) and are what the dispatcher uses when the
// tcgen05 path (qlora_tc.cu) does not apply.  The skinny rank-r contractions of the adapter
// (x.A^T, dy.B, dA, dB) are bandwidth-bound and are served from here for both families.
#include "vft_common.cuh"

namespace vft {

// ---------------------------------------------------------------------------
// Tiled GEMM with on-the-fly NF4 decode:   out[T, OUT] = act[T, RED] . W~  (+ extras)
//   forward : OUT = N, RED = K, W~(out=n, red=k) = flat element n*K + k
//   backward: OUT = K, RED = N, W~(out=k, red=n) = flat element n*K + k
// extras: + bias[out] (forward) + sum_j left[t, j] * right(out, j) (the rank-r adapter term).
// ---------------------------------------------------------------------------
constexpr int kTile = 64;
constexpr int kRedStep = 32;

template <typename ActT, bool kBackward>
__global__ void __launch_bounds__(256)
simt_nf4_gemm_kernel(const ActT* __restrict__ act, const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                     int64_t T, int64_t N, int64_t K, int blocksize, int qdtype, const ActT* __restrict__ bias,
                     const ActT* __restrict__ lora_left /* [T, VFT_LORA_LD] */,
                     const ActT* __restrict__ lora_right /* fwd: B [N, r]; bwd: A [r, K] */, int r, float scale,
                     ActT* __restrict__ out) {
  const int64_t OUT = kBackward ? K : N;
  const int64_t RED = kBackward ? N : K;
  __shared__ float s_act[kTile][kRedStep + 1];
  __shared__ float s_w[kTile][kRedStep + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int64_t t0 = (int64_t)blockIdx.y * kTile, o0 = (int64_t)blockIdx.x * kTile;
  float acc[4][4] = {};

  for (int64_t r0 = 0; r0 < RED; r0 += kRedStep) {
    for (int i = threadIdx.x; i < kTile * kRedStep; i += 256) {
      const int row = i / kRedStep, col = i % kRedStep;
      const int64_t t = t0 + row, red = r0 + col;
      s_act[row][col] = (t < T && red < RED) ? to_f32<ActT>(act[t * RED + red]) : 0.0f;
      const int64_t o = o0 + row;
      float w = 0.0f;
      if (o < OUT && red < RED) {
        const int64_t flat = kBackward ? (red * K + o) : (o * K + red);
        const uint8_t byte = packed[flat >> 1];
        const unsigned c = (flat & 1) ? (byte & 15u) : (byte >> 4);
        w = round_through<ActT>(__fmul_rn(nf4_code_value(c), absmax[flat / blocksize]), qdtype);
      }
      s_w[row][col] = w;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < kRedStep; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = s_act[ty * 4 + i][kk];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = s_w[tx * 4 + j][kk];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int64_t t = t0 + ty * 4 + i;
    if (t >= T) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t o = o0 + tx * 4 + j;
      if (o >= OUT) continue;
      float v = acc[i][j];
      if (!kBackward && bias != nullptr) v += to_f32<ActT>(bias[o]);
      if (r > 0) {
        float l = 0.0f;
        for (int jj = 0; jj < r; ++jj) {
          const float left = to_f32<ActT>(lora_left[t * VFT_LORA_LD + jj]);
          // forward: the operand the tensor path feeds the MMA is ActT(scale * B); keep the same rounding here
          const float right = kBackward ? to_f32<ActT>(lora_right[(int64_t)jj * K + o])
                                        : to_f32<ActT>(from_f32<ActT>(scale * to_f32<ActT>(lora_right[o * r + jj])));
          l = fmaf(left, right, l);
        }
        v += l;
      }
      out[t * OUT + o] = from_f32<ActT>(v);
    }
  }
}

template <typename ActT>
static int simt_gemm_typed(const LayerArgs& a, bool backward, const void* act, void* out, const void* lora_left,
                           cudaStream_t st) {
  const int64_t OUT = backward ? a.K : a.N;
  dim3 grid((unsigned)ceil_div64(OUT, kTile), (unsigned)ceil_div64(a.T, kTile));
  const ActT* right = static_cast<const ActT*>(backward ? a.lora_a : a.lora_b);
  if (backward)
    simt_nf4_gemm_kernel<ActT, true><<<grid, 256, 0, st>>>(static_cast<const ActT*>(act), a.packed, a.absmax, a.T, a.N,
                                                           a.K, a.blocksize, a.qdtype, nullptr,
                                                           static_cast<const ActT*>(lora_left), right, a.r, a.scale,
                                                           static_cast<ActT*>(out));
  else
    simt_nf4_gemm_kernel<ActT, false><<<grid, 256, 0, st>>>(static_cast<const ActT*>(act), a.packed, a.absmax, a.T, a.N,
                                                            a.K, a.blocksize, a.qdtype, static_cast<const ActT*>(a.bias),
                                                            static_cast<const ActT*>(lora_left), right, a.r, a.scale,
                                                            static_cast<ActT*>(out));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

#define VFT_DISPATCH_ACT(dtype, ...)                                          \
  switch (dtype) {                                                            \
    case VFT_BF16: { using ActT = __nv_bfloat16; return __VA_ARGS__; }        \
    case VFT_F16: { using ActT = __half; return __VA_ARGS__; }                \
    case VFT_F32: { using ActT = float; return __VA_ARGS__; }                 \
    default: set_error("unsupported activation dtype %d", dtype); return VFT_ERR_INVALID; \
  }

int simt_fwd(const LayerArgs& a, const void* x, void* y, const void* t_save, cudaStream_t st) {
  if (a.T == 0) return VFT_OK;
  VFT_DISPATCH_ACT(a.act_dtype, simt_gemm_typed<ActT>(a, false, x, y, t_save, st));
}

int simt_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st) {
  if (a.T == 0) return VFT_OK;
  VFT_DISPATCH_ACT(a.act_dtype, simt_gemm_typed<ActT>(a, true, dy, dx, dt_save, st));
}

// ---------------------------------------------------------------------------
// Skinny adapter contractions.  One warp per token row:
//   rowdot<false>:  out[t, j] = ActT(         sum_c m[t, c] * v[j, c] )   (x . A^T ; v = A [r, C])
//   rowdot<true> :  out[t, j] = ActT( scale * sum_c m[t, c] * v[c, j] )   (dy . B  ; v = B [C, r])
// out has leading dimension VFT_LORA_LD and is zero padded from r to VFT_LORA_LD.
// ---------------------------------------------------------------------------
template <typename ActT, bool kVTransposed>
__global__ void __launch_bounds__(256)
simt_rowdot_kernel(const ActT* __restrict__ m, const ActT* __restrict__ v, int64_t T, int64_t C, int r, float scale,
                   ActT* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t t = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= T) return;
  const ActT* row = m + t * C;
  for (int j0 = 0; j0 < VFT_LORA_LD; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    if (j0 < r) {
      for (int64_t c = lane; c < C; c += 32) {
        const float mv = to_f32<ActT>(row[c]);
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          if (j0 + j < r) {
            const float vv = kVTransposed ? to_f32<ActT>(v[c * r + j0 + j]) : to_f32<ActT>(v[(int64_t)(j0 + j) * C + c]);
            acc[j] = fmaf(mv, vv, acc[j]);
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 16; ++j)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
    }
    if (lane < 16) {
      float val = 0.0f;
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (lane == j) val = acc[j];
      out[t * VFT_LORA_LD + j0 + lane] = from_f32<ActT>(kVTransposed ? scale * val : val);
    }
  }
}

template <typename ActT>
static int rowdot_typed(const void* m, const void* v, int64_t T, int64_t C, int r, float scale, bool v_transposed,
                        void* out, cudaStream_t st) {
  const int warps = 8;
  const unsigned blocks = (unsigned)ceil_div64(T, warps);
  if (v_transposed)
    simt_rowdot_kernel<ActT, true><<<blocks, warps * 32, 0, st>>>(static_cast<const ActT*>(m),
                                                                  static_cast<const ActT*>(v), T, C, r, scale,
                                                                  static_cast<ActT*>(out));
  else
    simt_rowdot_kernel<ActT, false><<<blocks, warps * 32, 0, st>>>(static_cast<const ActT*>(m),
                                                                   static_cast<const ActT*>(v), T, C, r, scale,
                                                                   static_cast<ActT*>(out));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

// bt[j, n] = scale * B[n, j] for j < r, 0 for r <= j < 16 * ceil(r / 16): what the persistent tcgen05 forward writes on its
// way (qlora_tc2.cu); the other forward paths launch this
template <typename ActT>
__global__ void lora_bt_kernel(const ActT* __restrict__ b, int64_t N, int r, int rows, float scale, ActT* __restrict__ bt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * rows) return;
  const int64_t j = i / N, n = i % N;
  bt[i] = from_f32<ActT>(j < r ? scale * to_f32<ActT>(b[n * r + j]) : 0.0f);
}

int simt_lora_bt(const void* b, int64_t N, int r, float scale, int act_dtype, void* bt, cudaStream_t st) {
  const int rows = ((r + 15) / 16) * 16;
  const unsigned grid = (unsigned)ceil_div64(N * rows, 256);
  if (act_dtype == VFT_BF16)
    lora_bt_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(b), N, r, rows, scale, static_cast<__nv_bfloat16*>(bt));
  else if (act_dtype == VFT_F16)
    lora_bt_kernel<<<grid, 256, 0, st>>>(static_cast<const __half*>(b), N, r, rows, scale, static_cast<__half*>(bt));
  else
    lora_bt_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(b), N, r, rows, scale, static_cast<float*>(bt));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

// tt[j, t] = t_save[t, j] for j < 16 * ceil(r / 16): the transposed copy the backward's dA/dB job reads; written by the
// persistent tcgen05 forward on its way when it computes t itself, by this kernel otherwise
template <typename ActT>
__global__ void lora_tt_kernel(const ActT* __restrict__ t_save, int64_t T, int rows, ActT* __restrict__ tt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * rows) return;
  const int64_t j = i / T, t = i % T;
  tt[i] = t_save[t * VFT_LORA_LD + j];
}

int simt_lora_tt(const void* t_save, int64_t T, int r, int act_dtype, void* tt, cudaStream_t st) {
  const int rows = ((r + 15) / 16) * 16;
  if (T == 0) return VFT_OK;
  const unsigned grid = (unsigned)ceil_div64(T * rows, 256);
  if (act_dtype == VFT_F32)
    lora_tt_kernel<<<grid, 256, 0, st>>>(static_cast<const float*>(t_save), T, rows, static_cast<float*>(tt));
  else  // 16-bit types: a plain copy of two-byte elements
    lora_tt_kernel<<<grid, 256, 0, st>>>(static_cast<const uint16_t*>(t_save), T, rows, static_cast<uint16_t*>(tt));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

int simt_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                   cudaStream_t st) {
  if (T == 0) return VFT_OK;
  VFT_DISPATCH_ACT(act_dtype, rowdot_typed<ActT>(x, a, T, K, r, 1.0f, false, t_save, st));
}

int simt_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
                 cudaStream_t st) {
  if (T == 0) return VFT_OK;
  VFT_DISPATCH_ACT(act_dtype, rowdot_typed<ActT>(dy, b, T, N, r, scale, true, dt_save, st));
}

// ---------------------------------------------------------------------------
// Adapter weight gradients: acc[c, j] += sum_{t in split} m[t, c] * v[t, j]   (fp32 atomics into ws)
//   dA^T: m = x  [T, K], v = dt_save ; dB: m = dy [T, N], v = t_save (scaled afterwards)
// ---------------------------------------------------------------------------
constexpr int kDabCols = 128;
constexpr int kDabTok = 32;

template <typename ActT>
__global__ void __launch_bounds__(kDabCols)
simt_colsum_kernel(const ActT* __restrict__ m, const ActT* __restrict__ v, int64_t T, int64_t C, int r,
                   int64_t tokens_per_split, float* __restrict__ acc_out /* [C, r] fp32 */) {
  __shared__ float s_v[kDabTok][VFT_LORA_LD];
  const int64_t c = (int64_t)blockIdx.x * kDabCols + threadIdx.x;
  const int64_t t_begin = (int64_t)blockIdx.y * tokens_per_split;
  const int64_t t_end = (t_begin + tokens_per_split < T) ? t_begin + tokens_per_split : T;
  for (int j0 = 0; j0 < r; j0 += 16) {
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.0f;
    for (int64_t tb = t_begin; tb < t_end; tb += kDabTok) {
      __syncthreads();
      for (int i = threadIdx.x; i < kDabTok * 16; i += kDabCols) {
        const int tt = i >> 4, j = i & 15;
        const int64_t t = tb + tt;
        s_v[tt][j] = (t < t_end && j0 + j < r) ? to_f32<ActT>(v[t * VFT_LORA_LD + j0 + j]) : 0.0f;
      }
      __syncthreads();
      if (c < C) {
        const int lim = (int)((t_end - tb < kDabTok) ? (t_end - tb) : kDabTok);
        for (int tt = 0; tt < lim; ++tt) {
          const float mv = to_f32<ActT>(m[(tb + tt) * C + c]);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = fmaf(mv, s_v[tt][j], acc[j]);
        }
      }
    }
    if (c < C) {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (j0 + j < r) atomicAdd(acc_out + c * r + j0 + j, acc[j]);
    }
  }
}

// dA[j, k] = ActT(accA[k, j]);  dB[n, j] = ActT(scale * accB[n, j])
template <typename ActT>
__global__ void simt_dab_finalize_kernel(const float* __restrict__ accA, const float* __restrict__ accB, int64_t N,
                                         int64_t K, int r, float scale, ActT* __restrict__ dA, ActT* __restrict__ dB) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * r) {
    const int64_t j = i / K, k = i % K;
    dA[i] = from_f32<ActT>(accA[k * r + j]);
  }
  if (i < N * r) dB[i] = from_f32<ActT>(scale * accB[i]);
}

template <typename ActT>
static int dab_typed(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N,
                     int64_t K, int r, float scale, void* dA, void* dB, float* ws, cudaStream_t st) {
  float* accA = ws;          // [K, r]
  float* accB = ws + K * r;  // [N, r]
  VFT_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)(K + N) * r, st));
  if (T > 0) {
    // enough token splits to put ~4 CTAs on every SM
    auto splits_for = [&](int64_t C) {
      const int64_t col_blocks = ceil_div64(C, kDabCols);
      int64_t s = ceil_div64(148 * 4, col_blocks);
      const int64_t max_s = ceil_div64(T, kDabTok);
      if (s > max_s) s = max_s;
      return s < 1 ? (int64_t)1 : s;
    };
    {
      const int64_t s = splits_for(K);
      const int64_t per = ceil_div64(ceil_div64(T, s), kDabTok) * kDabTok;
      dim3 grid((unsigned)ceil_div64(K, kDabCols), (unsigned)ceil_div64(T, per));
      simt_colsum_kernel<ActT><<<grid, kDabCols, 0, st>>>(static_cast<const ActT*>(x),
                                                         static_cast<const ActT*>(dt_save), T, K, r, per, accA);
      VFT_CUDA_OK(cudaGetLastError());
    }
    {
      const int64_t s = splits_for(N);
      const int64_t per = ceil_div64(ceil_div64(T, s), kDabTok) * kDabTok;
      dim3 grid((unsigned)ceil_div64(N, kDabCols), (unsigned)ceil_div64(T, per));
      simt_colsum_kernel<ActT><<<grid, kDabCols, 0, st>>>(static_cast<const ActT*>(dy),
                                                         static_cast<const ActT*>(t_save), T, N, r, per, accB);
      VFT_CUDA_OK(cudaGetLastError());
    }
  }
  const int64_t total = (K > N ? K : N) * r;
  simt_dab_finalize_kernel<ActT><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
      accA, accB, N, K, r, scale, static_cast<ActT*>(dA), static_cast<ActT*>(dB));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

int simt_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
             int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st) {
  VFT_DISPATCH_ACT(act_dtype, dab_typed<ActT>(dy, x, t_save, dt_save, T, N, K, r, scale, dA, dB, ws, st));
}

}  // namespace vft
