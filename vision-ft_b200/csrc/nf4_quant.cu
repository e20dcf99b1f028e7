// NF4 blockwise quantize/pack (Q1) and dequantize kernels.
//
// Replaces bitsandbytes' kQuantizeBlockwise<T,64,2,0,NF4> / kDequantizeBlockwise
// reached from /root/reference/src/modules/quant/functional.py:362-365 and from
// Params4bit.cuda() (/root/reference/tools/quantize_model.py:53).
//
// Both kernels are pure streaming work and are bounded by HBM:
//   quantize  : 2 B/elem read (16-bit weights) + 0.5 B codes + 0.0625 B absmax written
//   dequantize: 0.5625 B/elem read + 2 B/elem written
// Design for the HBM roofline: 16-byte vector loads (8 weights per thread, 8 threads =
// one 64-element block, absmax by three xor-shuffles), 4 independent loads in flight per
// thread, one 4-byte coalesced store of 8 packed codes per thread.  The encode is the
// instruction-count bottleneck (a 4-level compare tree costs ~19 ALU ops/element, which
// would cap the kernel at ~30 % of HBM speed), so it is done with a 35-cell lookup that
// is EXACT: cell = round(16 x) (one FFMA against 2^23+17, mantissa bits), each cell holds
// at most one NF4 threshold (cells are 1/16 wide, thresholds >= 0.0805 apart), and
// code = base[cell] + (x > thr[cell]).  The table lives in shared memory replicated per
// lane ([cell][lane]) so the lookup never bank-conflicts.
#include <math.h>
#include <string.h>

#include "vft_common.cuh"

namespace vft {

constexpr int kCells = 36;
constexpr float kCellMagic = 8388625.0f;  // 2^23 + 17

struct CellTable {
  float thr[kCells];
  float base[kCells];  // small integers stored as float bit patterns of uint32 (see below)
};

static inline unsigned host_cell(float x) {
  float y = fmaf(x, 16.0f, kCellMagic);
  uint32_t u;
  memcpy(&u, &y, 4);
  return u & 63u;
}

static CellTable make_cell_table() {
  const float thr[15] = VFT_NF4_THRESHOLDS;
  CellTable t;
  int owner[kCells];
  for (int g = 0; g < kCells; ++g) owner[g] = -1;
  for (int i = 0; i < 15; ++i) {
    unsigned g = host_cell(thr[i]);
    // by construction every threshold owns a distinct cell; checked in tests through bit-exactness
    if (g < (unsigned)kCells && owner[g] < 0) owner[g] = i;
  }
  int below = 0;
  for (int g = 0; g < kCells; ++g) {
    uint32_t b = (uint32_t)below;
    memcpy(&t.base[g], &b, 4);
    if (owner[g] >= 0) {
      t.thr[g] = thr[owner[g]];
      ++below;
    } else {
      t.thr[g] = INFINITY;
    }
  }
  return t;
}

__device__ __forceinline__ unsigned encode_cell(float x, const float2* __restrict__ tab /* [cell][32] + lane */) {
  const float y = __fmaf_rn(x, 16.0f, kCellMagic);
  const unsigned g = __float_as_uint(y) & 63u;
  const float2 e = tab[g * 32];
  return __float_as_uint(e.y) + (x > e.x ? 1u : 0u);
}

template <typename T>
struct Vec8 {};
template <>
struct Vec8<__nv_bfloat16> {
  uint4 raw;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) { raw = __ldcs(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __uint_as_float(w[i] << 16);
      f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
  }
};
template <>
struct Vec8<__half> {
  uint4 raw;
  __device__ __forceinline__ void load(const __half* p) { raw = __ldcs(reinterpret_cast<const uint4*>(p)); }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
      f[2 * i] = v.x;
      f[2 * i + 1] = v.y;
    }
  }
};
template <>
struct Vec8<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    a = __ldcs(reinterpret_cast<const float4*>(p));
    b = __ldcs(reinterpret_cast<const float4*>(p) + 1);
  }
  __device__ __forceinline__ void get(float (&f)[8]) const {
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
    f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
  }
};

constexpr int kQuantThreads = 256;
constexpr int kQuantUnroll = 4;                         // independent 16-byte loads in flight per thread
constexpr int kWarpElems = 32 * 8;                      // elements per warp per load
constexpr int kChunkElems = kWarpElems * kQuantUnroll;  // 1024

// Fast path: blocksize == 64, n_main a multiple of 1024, w 16-byte aligned.
template <typename T>
__global__ void __launch_bounds__(kQuantThreads)
nf4_quantize64_kernel(const T* __restrict__ w, int64_t n_chunks, uint32_t* __restrict__ packed_words,
                      float* __restrict__ absmax, const CellTable table) {
  __shared__ float2 s_tab[kCells * 32];
  for (int i = threadIdx.x; i < kCells * 32; i += kQuantThreads) {
    const int g = i >> 5;
    s_tab[i] = make_float2(table.thr[g], table.base[g]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const float2* tab = s_tab + lane;
  const int64_t warp_global = (int64_t)blockIdx.x * (kQuantThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kQuantThreads / 32);

  for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warp_stride) {
    const int64_t base = chunk * kChunkElems + lane * 8;
    Vec8<T> v[kQuantUnroll];
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) v[u].load(w + base + u * kWarpElems);
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) {
      float f[8];
      v[u].get(f);
      float am = 0.0f;
#pragma unroll
      for (int i = 0; i < 8; ++i) am = fmaxf(am, fabsf(f[i]));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 1));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 2));
      am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, 4));
      const float s = __fdiv_rn(1.0f, am);  // IEEE reciprocal; +inf for an all-zero block
      uint32_t word = 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const unsigned c = encode_cell(__fmul_rn(f[i], s), tab);
        // element e=2j in the high nibble of byte j; bytes little-endian inside the word
        word |= c << (8 * (i >> 1) + ((i & 1) ? 0 : 4));
      }
      if (am == 0.0f) word = 0;  // 0 * inf = NaN -> every '>' false -> code 0
      const int64_t e0 = base + u * kWarpElems;
      packed_words[e0 >> 3] = word;
      if ((lane & 7) == 0) absmax[e0 >> 6] = am;
    }
  }
}

// Generic path: any blocksize, any n, any alignment.  One warp per quantization block.
template <typename T>
__global__ void nf4_quantize_generic_kernel(const T* __restrict__ w, int64_t start, int64_t n, int blocksize,
                                            uint8_t* __restrict__ packed, float* __restrict__ absmax) {
  constexpr float kThr[15] = VFT_NF4_THRESHOLDS;
  const int lane = threadIdx.x & 31;
  const int64_t warp = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int64_t lo = start + warp * blocksize;
  if (lo >= n) return;
  const int64_t hi = (lo + blocksize < n) ? lo + blocksize : n;
  float am = 0.0f;
  for (int64_t i = lo + lane; i < hi; i += 32) am = fmaxf(am, fabsf(to_f32<T>(w[i])));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) am = fmaxf(am, __shfl_xor_sync(0xffffffffu, am, o));
  if (lane == 0) absmax[lo / blocksize] = am;
  const float s = __fdiv_rn(1.0f, am);
  for (int64_t i = lo + 2 * lane; i < hi; i += 64) {
    unsigned c[2] = {0u, 0u};
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (i + h < hi) {
        const float x = __fmul_rn(to_f32<T>(w[i + h]), s);
#pragma unroll
        for (int t = 0; t < 15; ++t) c[h] += (x > kThr[t]) ? 1u : 0u;
      }
    }
    packed[i >> 1] = (uint8_t)((c[0] << 4) | c[1]);
  }
}

template <typename T>
static int quantize_typed(const T* w, int64_t n, int blocksize, uint8_t* packed, float* absmax, cudaStream_t st) {
  int64_t n_main = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(w) % 16 == 0) && (reinterpret_cast<uintptr_t>(packed) % 4 == 0);
  if (blocksize == 64 && aligned) n_main = (n / kChunkElems) * kChunkElems;
  if (n_main > 0) {
    static const CellTable table = make_cell_table();
    const int64_t n_chunks = n_main / kChunkElems;
    const int warps_per_block = kQuantThreads / 32;
    int64_t blocks = ceil_div64(n_chunks, warps_per_block);
    const int64_t max_blocks = 148 * 8;  // 8 resident CTAs of 256 threads per SM
    if (blocks > max_blocks) blocks = max_blocks;
    nf4_quantize64_kernel<T><<<(unsigned)blocks, kQuantThreads, 0, st>>>(
        w, n_chunks, reinterpret_cast<uint32_t*>(packed), absmax, table);
    VFT_CUDA_OK(cudaGetLastError());
  }
  if (n_main < n) {
    const int64_t rest_blocks = ceil_div64(n - n_main, blocksize);
    const int warps = 4;
    nf4_quantize_generic_kernel<T><<<(unsigned)ceil_div64(rest_blocks, warps), warps * 32, 0, st>>>(
        w, n_main, n, blocksize, packed, absmax);
    VFT_CUDA_OK(cudaGetLastError());
  }
  return VFT_OK;
}

int launch_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax,
                    cudaStream_t st) {
  VFT_REQUIRE(blocksize >= 2 && blocksize % 2 == 0 && blocksize <= 4096, "blocksize %d must be even and in [2, 4096]",
              blocksize);
  if (n == 0) return VFT_OK;
  switch (dtype) {
    case VFT_F32: return quantize_typed(static_cast<const float*>(w), n, blocksize, packed, absmax, st);
    case VFT_F16: return quantize_typed(static_cast<const __half*>(w), n, blocksize, packed, absmax, st);
    case VFT_BF16: return quantize_typed(static_cast<const __nv_bfloat16*>(w), n, blocksize, packed, absmax, st);
    default: set_error("unknown dtype %d", dtype); return VFT_ERR_INVALID;
  }
}

// ---------------------------------------------------------------------------
// dequantize: out[i] = T( codebook[code_i] * absmax[i / blocksize] )
// ---------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
nf4_dequantize_vec_kernel(const uint32_t* __restrict__ packed_words, const float* __restrict__ absmax, int64_t n_words,
                          int blocksize, T* __restrict__ out) {
  // code book replicated per lane ([code][lane]) so the lookup is bank-conflict free
  __shared__ float s_code[16 * 32];
  for (int i = threadIdx.x; i < 16 * 32; i += 256) s_code[i] = nf4_code_value(i >> 5);
  __syncthreads();
  const float* code = s_code + (threadIdx.x & 31);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t wi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; wi < n_words; wi += stride) {
    const uint32_t word = __ldcs(packed_words + wi);
    const float am = __ldg(absmax + (wi * 8) / blocksize);
    T vals[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const unsigned c = (word >> (8 * (i >> 1) + ((i & 1) ? 0 : 4))) & 15u;
      vals[i] = from_f32<T>(__fmul_rn(code[c * 32], am));
    }
    if constexpr (sizeof(T) == 2) {
      *reinterpret_cast<uint4*>(out + wi * 8) = *reinterpret_cast<const uint4*>(vals);
    } else {
      reinterpret_cast<uint4*>(out + wi * 8)[0] = reinterpret_cast<const uint4*>(vals)[0];
      reinterpret_cast<uint4*>(out + wi * 8)[1] = reinterpret_cast<const uint4*>(vals)[1];
    }
  }
}

template <typename T>
__global__ void nf4_dequantize_scalar_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax,
                                             int64_t start, int64_t n, int blocksize, T* __restrict__ out) {
  const int64_t i = start + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t byte = packed[i >> 1];
  const unsigned c = (i & 1) ? (byte & 15u) : (byte >> 4);
  out[i] = from_f32<T>(__fmul_rn(nf4_code_value(c), absmax[i / blocksize]));
}

template <typename T>
static int dequantize_typed(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, T* out,
                            cudaStream_t st) {
  int64_t n_main = 0;
  const bool aligned = (reinterpret_cast<uintptr_t>(packed) % 4 == 0) && (reinterpret_cast<uintptr_t>(out) % 16 == 0);
  if (aligned && blocksize % 8 == 0) n_main = (n / 8) * 8;
  if (n_main > 0) {
    const int64_t n_words = n_main / 8;
    int64_t blocks = ceil_div64(n_words, 256);
    if (blocks > 148 * 16) blocks = 148 * 16;
    nf4_dequantize_vec_kernel<T><<<(unsigned)blocks, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(packed), absmax,
                                                                    n_words, blocksize, out);
    VFT_CUDA_OK(cudaGetLastError());
  }
  if (n_main < n) {
    const int64_t rest = n - n_main;
    nf4_dequantize_scalar_kernel<T><<<(unsigned)ceil_div64(rest, 128), 128, 0, st>>>(packed, absmax, n_main, n,
                                                                                     blocksize, out);
    VFT_CUDA_OK(cudaGetLastError());
  }
  return VFT_OK;
}

int launch_dequantize(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, void* out, int dtype,
                      cudaStream_t st) {
  VFT_REQUIRE(blocksize >= 1, "blocksize %d must be positive", blocksize);
  if (n == 0) return VFT_OK;
  switch (dtype) {
    case VFT_F32: return dequantize_typed(packed, absmax, n, blocksize, static_cast<float*>(out), st);
    case VFT_F16: return dequantize_typed(packed, absmax, n, blocksize, static_cast<__half*>(out), st);
    case VFT_BF16: return dequantize_typed(packed, absmax, n, blocksize, static_cast<__nv_bfloat16*>(out), st);
    default: set_error("unknown dtype %d", dtype); return VFT_ERR_INVALID;
  }
}

// ---------------------------------------------------------------------------
// Micro-tiled copy of a packed weight for the fused kernels.
//
// bitsandbytes' layout keeps a weight row contiguous, so the 32 lanes of a decode warp -- one weight row each -- read
// 32 different 128-byte lines per load instruction (ncu, profiles/r01_*: 425 L1 wavefronts per pipeline step for the
// codes + absmax of one 128 x 64 tile, more than the tile's shared-memory stores).  The tiled copy groups
// 64 rows x 64 columns (one quantization block per row):
//   codes_t  [N/64][K/64][2 halves][64 rows][16 bytes]   half h = bytes 16h..16h+15 of the row's 32-byte block
//   absmax_t [N/64][K/64][64 rows] fp32
// so a warp's 32 rows are 512 contiguous bytes per LDG.128 and one 128-byte line of absmax.  Rows >= N are zero
// (code 0, absmax 0 -> decodes to 0).  The same copy serves forward (128 rows x 1 block) and backward (64 rows x
// 2 blocks).  Only defined for blocksize 64 and K % 64 == 0 (quantization blocks do not span rows).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
nf4_tile_kernel(const uint8_t* __restrict__ packed, const float* __restrict__ absmax, int64_t N, int64_t K,
                uint8_t* __restrict__ codes_t, float* __restrict__ absmax_t) {
  const int64_t KB = K >> 6;
  const int64_t kb = blockIdx.x, nb = blockIdx.y;
  const int r = threadIdx.x & 63, h = threadIdx.x >> 6;
  const int64_t row = nb * 64 + r;
  uint4 v = make_uint4(0, 0, 0, 0);
  if (row < N) v = *reinterpret_cast<const uint4*>(packed + ((row * K + kb * 64) >> 1) + h * 16);
  *reinterpret_cast<uint4*>(codes_t + ((nb * KB + kb) * 2 + h) * 1024 + r * 16) = v;
  if (h == 0) absmax_t[(nb * KB + kb) * 64 + r] = row < N ? absmax[row * KB + kb] : 0.0f;
}

int launch_tile_weight(const uint8_t* packed, const float* absmax, int64_t N, int64_t K, int blocksize,
                       uint8_t* codes_t, float* absmax_t, cudaStream_t st) {
  VFT_REQUIRE(blocksize == 64 && K % 64 == 0 && N > 0 && K > 0, "tiled weights need blocksize 64 and K %% 64 == 0");
  VFT_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15u) == 0 && (reinterpret_cast<uintptr_t>(codes_t) & 15u) == 0,
              "packed / codes_t must be 16-byte aligned");
  const int64_t nb = ceil_div64(N, 64), kb = K / 64;
  VFT_REQUIRE(nb <= 65535, "N too large for the tiling grid");
  nf4_tile_kernel<<<dim3((unsigned)kb, (unsigned)nb), 128, 0, st>>>(packed, absmax, N, K, codes_t, absmax_t);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

}  // namespace vft
