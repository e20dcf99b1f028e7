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

#include <vector>

#include "vft_common.cuh"

namespace vft {

// Cell table.  y = fma(x, 16, 2^15 + 17) has ulp 2^-8 for x in [-1, 1], so mantissa bits [8, 14) of y hold
// k = floor(16 x + 17) (x rounded to a 2^-12 grid first): cell k covers x in [(k - 17) / 16, (k - 16) / 16).  Cells
// are 1/16 wide and NF4 thresholds are >= 0.0805 apart, so a cell holds at most one threshold, and
//     code(x) = base[k] + (x > thr[k])        base[k] = number of thresholds below the cell
// is exact PROVIDED no threshold sits within the rounding distance below a cell boundary (checked when the table
// is built: a value there could be pushed into the next cell and skip its compare).  64 cells are allocated so
// that the NaN of an all-zero block (0 * inf) still indexes inside the table; its word is forced to 0 afterwards.
constexpr int kCells = 64;
constexpr float kCellMagic = 32768.0f + 17.0f;

struct CellTable {
  float thr[kCells];
  float base[kCells];  // number of thresholds below the cell, as a float (codes are accumulated in fp32)
};

static CellTable make_cell_table() {
  const float thr[15] = VFT_NF4_THRESHOLDS;
  CellTable t;
  for (int k = 0; k < kCells; ++k) {
    const double lo = (k - 17) / 16.0, hi = (k - 16) / 16.0;
    int below = 0, owner = -1;
    for (int i = 0; i < 15; ++i) {
      if ((double)thr[i] < lo) ++below;
      else if ((double)thr[i] < hi) owner = i;  // at most one (spacing > cell width)
    }
    t.base[k] = (float)below;
    t.thr[k] = owner >= 0 ? thr[owner] : INFINITY;
    // a threshold within 2^-8 / 16 (+ slack) below the upper boundary would break exactness: none of the 15 is
    if (owner >= 0 && hi - (double)thr[owner] < 2.0 / 4096.0) t.thr[k] = NAN;  // poison: tests fail loudly
  }
  return t;
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
// The kernel is bound by instruction issue, not by memory (ncu of the first version: 18 instructions per element,
// 82 % issue-active at 62 % of HBM speed), so the per-element sequence is kept to
//     unpack, FMUL (x = w * s), FFMA (cell), LOP3 (table address), LDS.64, FSET (x > thr -> 1.0/0.0), FADD, FFMA
// with the 4-bit codes accumulated as fp32 (two 16-bit halves per 8 elements, exact) so that the packing runs on
// the FMA pipe, and the table address is a single LOP3: the table is 16 KB-aligned in shared memory, replicated per
// lane ([cell][lane], 8 bytes each: conflict-free), and the cell index sits at bits [8, 14) of the FFMA result.
constexpr int kTabBytes = kCells * 32 * 8;  // 16 KB

// One launch quantizes a GROUP of tensors of equal size (vft_nf4_quantize_many: a checkpoint is hundreds of weights in a
// handful of shapes, and a launch per tensor costs ~2 us of an HBM-bound kernel that needs 4-30 us for one -- the
// 322-tensor AuraFlow set ran at 57 % of HBM bandwidth against 73 % for its largest tensor).  blockIdx.y selects the
// tensor, so the pointers are CTA-uniform and the hot loop is the single-tensor loop.  (A walk over one chunk index
// space across tensors of any size -- table look-up when a boundary is crossed -- cost 16 more registers, a fourth
// resident CTA per SM and 18 % of the single-tensor speed: the kernel is issue-bound.)  The table rides in the kernel
// parameters (constant bank: no workspace, capturable), kQuantBatch tensors per launch.
constexpr int kQuantBatch = 96;
struct QuantBatch {
  const void* w[kQuantBatch];
  uint32_t* packed[kQuantBatch];
  float* absmax[kQuantBatch];
};

template <typename T>
__global__ void __launch_bounds__(kQuantThreads)
nf4_quantize64_kernel(const __grid_constant__ QuantBatch batch, int64_t n_chunks, const __grid_constant__ CellTable table) {
  const T* __restrict__ w = static_cast<const T*>(batch.w[blockIdx.y]);
  uint32_t* __restrict__ packed_words = batch.packed[blockIdx.y];
  float* __restrict__ absmax = batch.absmax[blockIdx.y];
  extern __shared__ uint8_t q_smem[];
  const uint32_t raw = static_cast<uint32_t>(__cvta_generic_to_shared(q_smem));
  const uint32_t tab_base = (raw + (uint32_t)kTabBytes - 1u) & ~((uint32_t)kTabBytes - 1u);
  float2* s_tab = reinterpret_cast<float2*>(q_smem + (tab_base - raw));
  for (int i = threadIdx.x; i < kCells * 32; i += kQuantThreads) {
    const int g = i >> 5;
    s_tab[i] = make_float2(table.thr[g], table.base[g]);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const uint32_t lane_addr = tab_base + (uint32_t)lane * 8u;
  const int64_t warp_global = (int64_t)blockIdx.x * (kQuantThreads / 32) + (threadIdx.x >> 5);
  const int64_t warp_stride = (int64_t)gridDim.x * (kQuantThreads / 32);

  // register double buffering: the loads of the warp's NEXT chunk are in flight while this one is encoded
  Vec8<T> v[kQuantUnroll], nxt[kQuantUnroll];
  if (warp_global < n_chunks) {
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) v[u].load(w + warp_global * kChunkElems + lane * 8 + u * kWarpElems);
  }
  for (int64_t chunk = warp_global; chunk < n_chunks; chunk += warp_stride) {
    const int64_t base = chunk * kChunkElems + lane * 8;
    if (chunk + warp_stride < n_chunks) {
#pragma unroll
      for (int u = 0; u < kQuantUnroll; ++u)
        nxt[u].load(w + (chunk + warp_stride) * kChunkElems + lane * 8 + u * kWarpElems);
    }
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
      const float s = __frcp_rn(am);  // correctly rounded 1.0f / am (same value as IEEE division); +inf for am = 0
      // element 2j goes to the HIGH nibble of byte j: weights 16, 1, 4096, 256 inside each 16-bit half
      float half_acc[2] = {0.0f, 0.0f};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float x = __fmul_rn(f[i], s);
        const float y = __fmaf_rn(x, 16.0f, kCellMagic);
        uint32_t addr;  // (bits(y) & 0x3F00) | lane_addr as ONE LOP3 (the compiler otherwise emits AND + ADD)
        asm("lop3.b32 %0, %1, 0x3F00, %2, 0xEA;" : "=r"(addr) : "r"(__float_as_uint(y)), "r"(lane_addr));
        float2 e;
        asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(e.x), "=f"(e.y) : "r"(addr));
        const float code = e.y + (x > e.x ? 1.0f : 0.0f);
        const float wgt = (float)(1u << (8 * ((i & 3) >> 1) + ((i & 1) ? 0 : 4)));
        half_acc[i >> 2] = __fmaf_rn(code, wgt, half_acc[i >> 2]);
      }
      uint32_t word = __float2uint_rz(half_acc[0]) | (__float2uint_rz(half_acc[1]) << 16);
      if (am == 0.0f) word = 0;  // 0 * inf = NaN -> every '>' false -> code 0
      const int64_t e0 = base + u * kWarpElems;
      packed_words[e0 >> 3] = word;
      if ((lane & 7) == 0) absmax[e0 >> 6] = am;
    }
#pragma unroll
    for (int u = 0; u < kQuantUnroll; ++u) v[u] = nxt[u];
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
static int launch_quant_batch(const QuantBatch& b, int count, int64_t n_chunks, cudaStream_t st) {
  static const CellTable table = make_cell_table();
  const int warps_per_block = kQuantThreads / 32;
  // grid.x CTAs walk one tensor with the grid stride; about eight waves of resident CTAs over the whole group
  const int64_t resident = 148 * 6;  // 6 CTAs of 256 threads per SM by shared memory (32 KB each)
  int64_t gx = ceil_div64(n_chunks, warps_per_block);
  int64_t cap = count == 1 ? resident : (8 * resident) / count;
  if (cap < 8) cap = 8;
  if (gx > cap) gx = cap;
  nf4_quantize64_kernel<T><<<dim3((unsigned)gx, (unsigned)count), kQuantThreads, 2 * kTabBytes, st>>>(b, n_chunks, table);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

// tail of a tensor the fast path cannot take (blocksize != 64, n % 1024, misaligned pointers): one warp per block
template <typename T>
static int quantize_rest(const T* w, int64_t n_main, int64_t n, int blocksize, uint8_t* packed, float* absmax,
                         cudaStream_t st) {
  if (n_main >= n) return VFT_OK;
  const int64_t rest_blocks = ceil_div64(n - n_main, blocksize);
  const int warps = 4;
  nf4_quantize_generic_kernel<T><<<(unsigned)ceil_div64(rest_blocks, warps), warps * 32, 0, st>>>(
      w, n_main, n, blocksize, packed, absmax);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

static int64_t fast_elems(const void* w, int64_t n, int blocksize, const uint8_t* packed) {
  const bool aligned = (reinterpret_cast<uintptr_t>(w) % 16 == 0) && (reinterpret_cast<uintptr_t>(packed) % 4 == 0);
  return (blocksize == 64 && aligned) ? (n / kChunkElems) * kChunkElems : 0;
}

template <typename T>
static int quantize_many_typed(int count, const void* const* ws, const int64_t* ns, int blocksize, uint8_t* const* packed,
                               float* const* absmax, cudaStream_t st) {
  // groups of equal fast-path size, in order of first appearance (model checkpoints repeat a handful of shapes)
  std::vector<char> done((size_t)count, 0);
  for (int i = 0; i < count; ++i) {
    if (done[i]) continue;
    const int64_t n_main = ns[i] > 0 ? fast_elems(ws[i], ns[i], blocksize, packed[i]) : 0;
    QuantBatch b;
    int k = 0;
    for (int j = i; j < count; ++j) {
      if (done[j] || ns[j] == 0) { done[j] = 1; continue; }
      if (fast_elems(ws[j], ns[j], blocksize, packed[j]) != n_main) continue;
      done[j] = 1;
      if (n_main > 0) {
        b.w[k] = ws[j];
        b.packed[k] = reinterpret_cast<uint32_t*>(packed[j]);
        b.absmax[k] = absmax[j];
        if (++k == kQuantBatch) {
          const int rc = launch_quant_batch<T>(b, k, n_main / kChunkElems, st);
          if (rc != VFT_OK) return rc;
          k = 0;
        }
      }
      const int rc = quantize_rest(static_cast<const T*>(ws[j]), n_main, ns[j], blocksize, packed[j], absmax[j], st);
      if (rc != VFT_OK) return rc;
    }
    if (k > 0) {
      const int rc = launch_quant_batch<T>(b, k, n_main / kChunkElems, st);
      if (rc != VFT_OK) return rc;
    }
  }
  return VFT_OK;
}

int launch_quantize_many(int count, const void* const* w, int dtype, const int64_t* n, int blocksize,
                         uint8_t* const* packed, float* const* absmax, cudaStream_t st) {
  VFT_REQUIRE(blocksize >= 2 && blocksize % 2 == 0 && blocksize <= 4096, "blocksize %d must be even and in [2, 4096]",
              blocksize);
  VFT_REQUIRE(count >= 0 && (count == 0 || (w && n && packed && absmax)), "null table");
  for (int i = 0; i < count; ++i)
    VFT_REQUIRE(n[i] >= 0 && (n[i] == 0 || (w[i] && packed[i] && absmax[i])), "tensor %d: null pointer or negative size", i);
  switch (dtype) {
    case VFT_F32: return quantize_many_typed<float>(count, w, n, blocksize, packed, absmax, st);
    case VFT_F16: return quantize_many_typed<__half>(count, w, n, blocksize, packed, absmax, st);
    case VFT_BF16: return quantize_many_typed<__nv_bfloat16>(count, w, n, blocksize, packed, absmax, st);
    default: set_error("unknown dtype %d", dtype); return VFT_ERR_INVALID;
  }
}

int launch_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax,
                    cudaStream_t st) {
  if (n == 0) {
    VFT_REQUIRE(blocksize >= 2 && blocksize % 2 == 0 && blocksize <= 4096, "blocksize %d must be even and in [2, 4096]",
                blocksize);
    return VFT_OK;
  }
  return launch_quantize_many(1, &w, dtype, &n, blocksize, &packed, &absmax, st);
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
