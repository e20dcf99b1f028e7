// Nested ("double quant") block statistics of the NF4 format: encode and decode of the fp32 absmax vector.
//
// Stands where bitsandbytes' quantize_4bit(compress_statistics=True) tail stands --
//     offset = absmax.mean(); absmax -= offset
//     absmax8, state2 = quantize_blockwise(absmax, blocksize=256)        (8-bit "dynamic" map)
// reached from Params4bit.cuda() (/root/reference/src/modules/quant/bnb.py:44,122-129; the path
// /root/reference/tools/quantize_model.py:33-54 takes) -- and where dequantize_blockwise(absmax8, state2) + offset
// stands on the decode side (every Linear4bit forward of a nested checkpoint in the reference; once per weight here).
//
// HBM-bound byte work on a small vector (N*K/64 floats: 0.6 MB for 3072x3072, 3.5 MB for 18432x3072): three
// launches, coalesced 16-byte loads, no atomics, bit-reproducible run to run.
//   1. absmax_sum_kernel     fixed grid of kSumParts CTAs, fp64 partial sums in a fixed order -> workspace
//   2. absmax_nest_kernel    every CTA re-adds the kSumParts partials in the same order (so all agree on the mean,
//                            rounded once to fp32), then one warp per 256-block: max |a - offset|, reciprocal,
//                            binary search in the 256-entry sorted map (shared memory), nearest of the two neighbours
//   3. absmax_denest_kernel  out = map[q] * absmax2[i / 256] + offset   (two separately rounded fp32 operations)
#include "vft_common.cuh"

namespace vft {

namespace {

constexpr int kSumParts = 128;    // partial sums (fixed: the reduction order never depends on the device)
constexpr int kSumThreads = 256;
constexpr int kNestThreads = 256;  // 8 warps = 8 statistics blocks per CTA pass

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void __launch_bounds__(kSumThreads) absmax_sum_kernel(const float* __restrict__ a, int64_t n,
                                                                 double* __restrict__ parts) {
  // contiguous slice per CTA; inside it a thread walks with a fixed stride, so the order of additions is a function
  // of (n, kSumParts, kSumThreads) only
  const int64_t per = (n + kSumParts - 1) / kSumParts;
  const int64_t lo = (int64_t)blockIdx.x * per;
  const int64_t hi = lo + per < n ? lo + per : n;
  double s = 0.0;
  for (int64_t i = lo + threadIdx.x; i < hi; i += kSumThreads) s += (double)a[i];
  __shared__ double sh[kSumThreads / 32];
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < kSumThreads / 32; ++w) t += sh[w];
    parts[blockIdx.x] = t;
  }
}

// bitsandbytes' dQuantize<0>: seven-step bisection from pivot 127, then the nearer of the pivot and the neighbour on
// x's side (strict comparisons: a value exactly on a midpoint stays on the pivot).
__device__ __forceinline__ unsigned quantize_dynamic8(const float* __restrict__ code, float x) {
  int pivot = 127, upper_pivot = 255, lower_pivot = 0;
  float lower = -1.0f, upper = 1.0f;
  float val = code[pivot];
#pragma unroll
  for (int i = 64; i > 0; i >>= 1) {
    if (x > val) {
      lower_pivot = pivot;
      lower = val;
      pivot += i;
    } else {
      upper_pivot = pivot;
      upper = val;
      pivot -= i;
    }
    val = code[pivot];
  }
  if (upper_pivot == 255) upper = code[upper_pivot];
  if (lower_pivot == 0) lower = code[lower_pivot];
  if (x > val) {
    const float mid = __fmul_rn(__fadd_rn(upper, val), 0.5f);
    return x > mid ? upper_pivot : pivot;
  }
  const float mid = __fmul_rn(__fadd_rn(lower, val), 0.5f);
  return x < mid ? lower_pivot : pivot;
}

__global__ void __launch_bounds__(kNestThreads) absmax_nest_kernel(const float* __restrict__ a, int64_t n,
                                                                   const float* __restrict__ code,
                                                                   const double* __restrict__ parts,
                                                                   uint8_t* __restrict__ q, float* __restrict__ absmax2,
                                                                   float* __restrict__ offset_out,
                                                                   const float* __restrict__ offset_in) {
  __shared__ float s_code[256];
  __shared__ float s_offset;
  s_code[threadIdx.x] = code[threadIdx.x];
  if (threadIdx.x == 0) {
    if (offset_in != nullptr) {  // the caller's offset (bitsandbytes: torch's absmax.mean() on the device)
      s_offset = *offset_in;
    } else {
      double t = 0.0;
      for (int p = 0; p < kSumParts; ++p) t += parts[p];
      const float off = (float)(t / (double)n);
      s_offset = off;
      if (blockIdx.x == 0) *offset_out = off;
    }
  }
  __syncthreads();
  const float off = s_offset;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t nblk = (n + 255) / 256;
  for (int64_t b = (int64_t)blockIdx.x * (kNestThreads / 32) + warp; b < nblk; b += (int64_t)gridDim.x * (kNestThreads / 32)) {
    const int64_t base = b * 256 + lane * 8;
    float v[8];
    if (base + 8 <= n) {
      const float4 p0 = *reinterpret_cast<const float4*>(a + base);
      const float4 p1 = *reinterpret_cast<const float4*>(a + base + 4);
      v[0] = p0.x; v[1] = p0.y; v[2] = p0.z; v[3] = p0.w;
      v[4] = p1.x; v[5] = p1.y; v[6] = p1.z; v[7] = p1.w;
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = base + j < n ? a[base + j] : off;  // padding -> 0 after the subtraction
    }
    float m = 0.0f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      v[j] = __fsub_rn(v[j], off);
      m = fmaxf(m, fabsf(v[j]));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    const float inv = __frcp_rn(m);  // 1.0f / m, IEEE
    uint32_t w0 = 0, w1 = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) w0 |= quantize_dynamic8(s_code, __fmul_rn(v[j], inv)) << (8 * j);
#pragma unroll
    for (int j = 0; j < 4; ++j) w1 |= quantize_dynamic8(s_code, __fmul_rn(v[4 + j], inv)) << (8 * j);
    if (base + 8 <= n) {
      *reinterpret_cast<uint2*>(q + base) = make_uint2(w0, w1);
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (base + j < n) q[base + j] = (uint8_t)((j < 4 ? w0 >> (8 * j) : w1 >> (8 * (j - 4))) & 0xffu);
    }
    if (lane == 0) absmax2[b] = m;
  }
}

__global__ void __launch_bounds__(256) absmax_denest_kernel(const uint8_t* __restrict__ q,
                                                            const float* __restrict__ absmax2,
                                                            const float* __restrict__ code, float offset, int64_t n,
                                                            int bs2, float* __restrict__ out) {
  __shared__ float s_code[256];
  s_code[threadIdx.x] = code[threadIdx.x];
  __syncthreads();
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n && (bs2 & 3) == 0) {
      const uint32_t w = *reinterpret_cast<const uint32_t*>(q + i);
      const float s = absmax2[i / bs2];
      float4 o;
      o.x = __fadd_rn(__fmul_rn(s_code[w & 0xffu], s), offset);
      o.y = __fadd_rn(__fmul_rn(s_code[(w >> 8) & 0xffu], s), offset);
      o.z = __fadd_rn(__fmul_rn(s_code[(w >> 16) & 0xffu], s), offset);
      o.w = __fadd_rn(__fmul_rn(s_code[w >> 24], s), offset);
      *reinterpret_cast<float4*>(out + i) = o;
    } else {
      for (int j = 0; j < 4 && i + j < n; ++j)
        out[i + j] = __fadd_rn(__fmul_rn(s_code[q[i + j]], absmax2[(i + j) / bs2]), offset);
    }
  }
}

}  // namespace

int64_t absmax_nest_workspace_bytes() { return (int64_t)sizeof(double) * kSumParts; }

int launch_absmax_nest(const float* absmax, int64_t n, int blocksize2, const float* code256, uint8_t* absmax8,
                       float* absmax2, float* offset_out, void* ws, int64_t ws_bytes, cudaStream_t st,
                       const float* offset_in) {
  VFT_REQUIRE(blocksize2 == 256, "nested statistics use blocksize 256 (bitsandbytes), got %d", blocksize2);
  VFT_REQUIRE(n > 0, "empty statistics vector");
  VFT_REQUIRE((reinterpret_cast<uintptr_t>(absmax) & 15u) == 0 && (reinterpret_cast<uintptr_t>(absmax8) & 7u) == 0,
              "absmax must be 16-byte aligned, absmax8 8-byte aligned");
  double* parts = nullptr;
  if (offset_in == nullptr) {
    if (ws == nullptr || ws_bytes < absmax_nest_workspace_bytes()) {
      set_error("workspace too small: need %lld bytes, got %lld", (long long)absmax_nest_workspace_bytes(),
                (long long)ws_bytes);
      return VFT_ERR_WORKSPACE;
    }
    parts = static_cast<double*>(ws);
    absmax_sum_kernel<<<kSumParts, kSumThreads, 0, st>>>(absmax, n, parts);
    VFT_CUDA_OK(cudaGetLastError());
  }
  const int64_t nblk = ceil_div64(n, 256);
  const int grid = (int)(ceil_div64(nblk, kNestThreads / 32) < 148 * 4 ? ceil_div64(nblk, kNestThreads / 32) : 148 * 4);
  absmax_nest_kernel<<<grid, kNestThreads, 0, st>>>(absmax, n, code256, parts, absmax8, absmax2, offset_out, offset_in);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

int launch_absmax_denest(const uint8_t* absmax8, const float* absmax2, const float* code256, float offset, int64_t n,
                         int blocksize2, float* out, cudaStream_t st) {
  VFT_REQUIRE(blocksize2 > 0, "bad nested blocksize %d", blocksize2);
  if (n == 0) return VFT_OK;
  VFT_REQUIRE((reinterpret_cast<uintptr_t>(out) & 15u) == 0 && (reinterpret_cast<uintptr_t>(absmax8) & 3u) == 0,
              "out must be 16-byte aligned, absmax8 4-byte aligned");
  const int64_t groups = ceil_div64(n, 4 * 256);
  const int grid = (int)(groups < 148 * 8 ? groups : 148 * 8);
  absmax_denest_kernel<<<grid, 256, 0, st>>>(absmax8, absmax2, code256, offset, n, blocksize2, out);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

}  // namespace vft
