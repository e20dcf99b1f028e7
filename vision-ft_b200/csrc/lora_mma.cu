// Rank-r adapter contractions on the warp-level tensor path (mma.sync m16n8k16), cp.async-pipelined.
//
//   rowproj :  out[T, 64]  = ActT( scale * M[T, C] . V )      x.A^T (V = A [r, C], "NT")  /  dy.B (V = B [C, r], "NN")
//   colproj :  acc[C, r]  += M[t0:t1, C]^T . V[t0:t1, r]      dA^T = x^T.dt   and   dB = dy^T.t   (fp32 atomics)
//
// These are the autograd pieces of /root/reference/src/modules/peft/lora.py:100-104 that do not
// fit the big fused GEMM: r <= 64 wide, 0.4-0.8 GFLOP each for the 3072x3072 layer, i.e. bound by
// reading x / dy once from HBM (25 MB each) -- but still ~80 TFLOP/s of math at that speed, which is
// beyond the CUDA cores, hence mma.sync.  tcgen05 would need M >= 64 rows of accumulator per
// instruction and a TMEM round trip for a 16-wide result; the warp-level MMA keeps the accumulators
// in registers and lets 128+ small CTAs keep enough bytes in flight to stream at HBM speed.
// Tiles are staged with 16-byte cp.async into XOR-swizzled shared memory (conflict-free ldmatrix).
#include "vft_common.cuh"

namespace vft {
namespace {

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_8(uint32_t dst, const void* src, bool valid) {
  const int n = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <typename ActT>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
__device__ __forceinline__ void mma_16816<__nv_bfloat16>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
__device__ __forceinline__ void mma_16816<__half>(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename ActT>
__device__ __forceinline__ uint32_t pack_pair(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack_pair<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <>
__device__ __forceinline__ uint32_t pack_pair<__half>(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

// ---------------------------------------------------------------------------------------------
// rowproj: CTA = 2 warps = 32 token rows; contraction streamed in chunks of 128 through a 4-stage ring
// ---------------------------------------------------------------------------------------------
constexpr int kRpTok = 32, kRpChunk = 128, kRpStages = 8, kRpThreads = 64;  // 7 x 8 KB of x in flight per CTA

template <int kNT>
struct RowProjSmem {
  static constexpr int rp = kNT * 8;                   // padded rank
  static constexpr int m_bytes = kRpTok * kRpChunk * 2;  // 8 KB
  static constexpr int v_bytes = rp * kRpChunk * 2;
  static constexpr int stage = m_bytes + v_bytes;
  static constexpr int total = stage * kRpStages;
};

template <typename ActT, int kNT, bool kVIsCxR>
__global__ void __launch_bounds__(kRpThreads)
lora_rowproj_kernel(const ActT* __restrict__ M, const ActT* __restrict__ V, int64_t T, int64_t C, int r, float scale,
                    ActT* __restrict__ out) {
  using S = RowProjSmem<kNT>;
  constexpr int RP = S::rp;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_addr(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t t0 = (int64_t)blockIdx.x * kRpTok;
  const int n_chunks = (int)((C + kRpChunk - 1) / kRpChunk);

  if (kVIsCxR) {  // columns r..RP of the [c][RP] tiles are never written by cp.async: zero them once
    for (int i = threadIdx.x; i < S::total / 16; i += kRpThreads)
      reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    __syncthreads();
  }

  auto issue = [&](int ch) {
    const uint32_t mb = sbase + (ch % kRpStages) * S::stage;
    const uint32_t vb = mb + S::m_bytes;
    const int64_t c0 = (int64_t)ch * kRpChunk;
    for (int i = threadIdx.x; i < kRpTok * 16; i += kRpThreads) {
      const int row = i >> 4, chunk = i & 15;
      const int64_t t = t0 + row, c = c0 + chunk * 8;
      const bool ok = t < T && c < C;
      cp_async_16(mb + row * 256 + ((chunk ^ (row & 7)) << 4), ok ? (const void*)(M + t * C + c) : (const void*)M, ok);
    }
    if (!kVIsCxR) {  // V = [r, C]: rows j, contraction contiguous
      for (int i = threadIdx.x; i < RP * 16; i += kRpThreads) {
        const int j = i >> 4, chunk = i & 15;
        const int64_t c = c0 + chunk * 8;
        const bool ok = j < r && c < C;
        cp_async_16(vb + j * 256 + ((chunk ^ (j & 7)) << 4), ok ? (const void*)(V + (int64_t)j * C + c) : (const void*)V, ok);
      }
    } else if ((r & 7) == 0) {  // V = [C, r]: rows c, r values each, copied in 16-byte units
      const int units = r >> 3;
      for (int i = threadIdx.x; i < kRpChunk * units; i += kRpThreads) {
        const int row = i / units, u = i - row * units;
        const int64_t c = c0 + row;
        const bool ok = c < C;
        cp_async_16(vb + row * (RP * 2) + u * 16, ok ? (const void*)(V + c * r + u * 8) : (const void*)V, ok);
      }
    } else {  // ... or in 8-byte units (r = 4, 12, ...)
      const int units = r >> 2;
      for (int i = threadIdx.x; i < kRpChunk * units; i += kRpThreads) {
        const int row = i / units, u = i - row * units;
        const int64_t c = c0 + row;
        const bool ok = c < C;
        cp_async_8(vb + row * (RP * 2) + u * 8, ok ? (const void*)(V + c * r + u * 4) : (const void*)V, ok);
      }
    }
  };

  float acc[kNT][4];
#pragma unroll
  for (int i = 0; i < kNT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int s = 0; s < kRpStages - 1; ++s) {
    if (s < n_chunks) issue(s);
    cp_async_commit();
  }
  for (int ch = 0; ch < n_chunks; ++ch) {
    cp_async_wait<kRpStages - 2>();
    __syncthreads();
    if (ch + kRpStages - 1 < n_chunks) issue(ch + kRpStages - 1);
    cp_async_commit();
    const uint32_t mb = sbase + (ch % kRpStages) * S::stage;
    const uint32_t vb = mb + S::m_bytes;
#pragma unroll
    for (int ks = 0; ks < kRpChunk / 16; ++ks) {
      uint32_t a[4];
      {
        const int row = warp * 16 + (lane & 15), chunk = ks * 2 + (lane >> 4);
        ldsm_x4(mb + row * 256 + ((chunk ^ (row & 7)) << 4), a);
      }
#pragma unroll
      for (int np = 0; np < kNT / 2; ++np) {
        uint32_t b[4];
        const int mi = lane >> 3, nt = 2 * np + (mi >> 1), kh = mi & 1;
        if (!kVIsCxR) {
          const int j = nt * 8 + (lane & 7), chunk = ks * 2 + kh;
          ldsm_x4(vb + j * 256 + ((chunk ^ (j & 7)) << 4), b);
        } else {
          const int crow = ks * 16 + kh * 8 + (lane & 7);
          ldsm_x4_trans(vb + crow * (RP * 2) + nt * 16, b);
        }
        mma_16816<ActT>(acc[2 * np], a, b[0], b[1]);
        mma_16816<ActT>(acc[2 * np + 1], a, b[2], b[3]);
      }
    }
  }
  cp_async_wait<0>();

  // epilogue: rows (lane / 4) and (lane / 4 + 8) of this warp's 16 tokens; all 64 columns of the padded output
  const int col = 2 * (lane & 3);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int64_t t = t0 + warp * 16 + (lane >> 2) + 8 * h;
    if (t >= T) continue;
    uint32_t* orow = reinterpret_cast<uint32_t*>(out + t * VFT_LORA_LD);
#pragma unroll
    for (int nt = 0; nt < VFT_LORA_LD / 8; ++nt) {
      float v0 = 0.0f, v1 = 0.0f;
      if (nt < kNT) {
        const int j = nt * 8 + col;
        v0 = (j < r) ? scale * acc[nt < kNT ? nt : 0][2 * h] : 0.0f;
        v1 = (j + 1 < r) ? scale * acc[nt < kNT ? nt : 0][2 * h + 1] : 0.0f;
      }
      orow[(nt * 8 + col) >> 1] = pack_pair<ActT>(v0, v1);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// colproj: CTA = 4 warps = 128 output columns (32 per warp); tokens of this split streamed 32 at a time
// ---------------------------------------------------------------------------------------------
constexpr int kCpCols = 128, kCpTok = 32, kCpStages = 6, kCpThreads = 128;

template <int kNT>
struct ColProjSmem {
  static constexpr int m_bytes = kCpTok * kCpCols * 2;   // 8 KB: [32 tokens][128 cols]
  static constexpr int v_bytes = kCpTok * VFT_LORA_LD * 2;  // 4 KB: [32 tokens][64]
  static constexpr int stage = m_bytes + v_bytes;
  static constexpr int total = stage * kCpStages;
};

template <typename ActT, int kNT>
__global__ void __launch_bounds__(kCpThreads)
lora_colproj_kernel(const ActT* __restrict__ X, const ActT* __restrict__ dT, int64_t K, float* __restrict__ accA,
                    const ActT* __restrict__ dY, const ActT* __restrict__ Tm, int64_t N, float* __restrict__ accB,
                    int64_t T, int r, int64_t tokens_per_split) {
  using S = ColProjSmem<kNT>;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sbase = smem_addr(smem);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tilesK = (int)((K + kCpCols - 1) / kCpCols);
  const bool second = (int)blockIdx.x >= tilesK;  // block-uniform: which of the two products
  const ActT* __restrict__ M = second ? dY : X;
  const ActT* __restrict__ V = second ? Tm : dT;
  const int64_t C = second ? N : K;
  float* __restrict__ acc_out = second ? accB : accA;
  const int64_t c0 = (int64_t)(second ? blockIdx.x - tilesK : blockIdx.x) * kCpCols;
  const int64_t tb = (int64_t)blockIdx.y * tokens_per_split;
  const int64_t te = (tb + tokens_per_split < T) ? tb + tokens_per_split : T;
  const int n_chunks = (int)((te - tb + kCpTok - 1) / kCpTok);

  auto issue = [&](int ch) {
    const uint32_t mb = sbase + (ch % kCpStages) * S::stage;
    const uint32_t vb = mb + S::m_bytes;
    const int64_t t_first = tb + (int64_t)ch * kCpTok;
    for (int i = threadIdx.x; i < kCpTok * 16; i += kCpThreads) {
      const int row = i >> 4, chunk = i & 15;
      const int64_t t = t_first + row, c = c0 + chunk * 8;
      const bool ok = t < te && c < C;
      cp_async_16(mb + row * 256 + ((chunk ^ (row & 7)) << 4), ok ? (const void*)(M + t * C + c) : (const void*)M, ok);
    }
    for (int i = threadIdx.x; i < kCpTok * kNT; i += kCpThreads) {
      const int row = i / kNT, chunk = i - row * kNT;
      const int64_t t = t_first + row;
      const bool ok = t < te;
      cp_async_16(vb + row * 128 + ((chunk ^ (row & 7)) << 4),
                  ok ? (const void*)(V + t * VFT_LORA_LD + chunk * 8) : (const void*)V, ok);
    }
  };

  float acc[2][kNT][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int i = 0; i < kNT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][i][j] = 0.0f;

  for (int s = 0; s < kCpStages - 1; ++s) {
    if (s < n_chunks) issue(s);
    cp_async_commit();
  }
  for (int ch = 0; ch < n_chunks; ++ch) {
    cp_async_wait<kCpStages - 2>();
    __syncthreads();
    if (ch + kCpStages - 1 < n_chunks) issue(ch + kCpStages - 1);
    cp_async_commit();
    const uint32_t mb = sbase + (ch % kCpStages) * S::stage;
    const uint32_t vb = mb + S::m_bytes;
#pragma unroll
    for (int ks = 0; ks < kCpTok / 16; ++ks) {
      const int mi = lane >> 3;
      uint32_t b[kNT / 2][4];
#pragma unroll
      for (int np = 0; np < kNT / 2; ++np) {
        const int t = ks * 16 + (mi & 1) * 8 + (lane & 7), nt = 2 * np + (mi >> 1);
        ldsm_x4_trans(vb + t * 128 + ((nt ^ (t & 7)) << 4), b[np]);
      }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t a[4];
        // A(c, t) = M[t][c]: source 8x8 blocks are [t rows][8 cols]; .trans hands each thread (c = lane/4, t pair)
        const int t = ks * 16 + (mi >> 1) * 8 + (lane & 7);
        const int chunk = (warp * 32 + mt * 16) / 8 + (mi & 1);
        ldsm_x4_trans(mb + t * 256 + ((chunk ^ (t & 7)) << 4), a);
#pragma unroll
        for (int np = 0; np < kNT / 2; ++np) {
          mma_16816<ActT>(acc[mt][2 * np], a, b[np][0], b[np][1]);
          mma_16816<ActT>(acc[mt][2 * np + 1], a, b[np][2], b[np][3]);
        }
      }
    }
  }
  cp_async_wait<0>();

#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int64_t c = c0 + warp * 32 + mt * 16 + (lane >> 2) + 8 * h;
      if (c >= C) continue;
#pragma unroll
      for (int nt = 0; nt < kNT; ++nt) {
        const int j = nt * 8 + 2 * (lane & 3);
        if (j < r) atomicAdd(acc_out + c * r + j, acc[mt][nt][2 * h]);
        if (j + 1 < r) atomicAdd(acc_out + c * r + j + 1, acc[mt][nt][2 * h + 1]);
      }
    }
}

// dA[j, k] = ActT(accA[k, j]);  dB[n, j] = ActT(scale * accB[n, j])
template <typename ActT>
__global__ void lora_dab_finalize_kernel(const float* __restrict__ accA, const float* __restrict__ accB, int64_t N,
                                         int64_t K, int r, float scale, ActT* __restrict__ dA, ActT* __restrict__ dB) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < K * r) {
    const int64_t j = i / K, k = i - j * K;
    dA[i] = from_f32<ActT>(accA[k * r + j]);
  }
  if (i < N * r) dB[i] = from_f32<ActT>(scale * accB[i]);
}

template <typename ActT, int kNT, bool kVIsCxR>
int launch_rowproj(const void* M, const void* V, int64_t T, int64_t C, int r, float scale, void* out, cudaStream_t st) {
  using S = RowProjSmem<kNT>;
  auto kern = lora_rowproj_kernel<ActT, kNT, kVIsCxR>;
  if (S::total > 48 * 1024) VFT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total));
  kern<<<(unsigned)ceil_div64(T, kRpTok), kRpThreads, S::total, st>>>(static_cast<const ActT*>(M), static_cast<const ActT*>(V),
                                                                     T, C, r, scale, static_cast<ActT*>(out));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT, bool kVIsCxR>
int rowproj_rank(const void* M, const void* V, int64_t T, int64_t C, int r, float scale, void* out, cudaStream_t st) {
  if (r <= 16) return launch_rowproj<ActT, 2, kVIsCxR>(M, V, T, C, r, scale, out, st);
  if (r <= 32) return launch_rowproj<ActT, 4, kVIsCxR>(M, V, T, C, r, scale, out, st);
  return launch_rowproj<ActT, 8, kVIsCxR>(M, V, T, C, r, scale, out, st);
}

template <typename ActT, int kNT>
int launch_colproj(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N,
                   int64_t K, int r, float* accA, float* accB, cudaStream_t st) {
  using S = ColProjSmem<kNT>;
  auto kern = lora_colproj_kernel<ActT, kNT>;
  if (S::total > 48 * 1024) VFT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::total));
  const int64_t col_tiles = ceil_div64(K, kCpCols) + ceil_div64(N, kCpCols);
  int64_t splits = ceil_div64(148 * 4, col_tiles);  // ~4 CTAs per SM keep enough loads in flight
  const int64_t max_splits = ceil_div64(T, kCpTok);
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  const int64_t per = ceil_div64(ceil_div64(T, splits), kCpTok) * kCpTok;
  dim3 grid((unsigned)col_tiles, (unsigned)ceil_div64(T, per));
  kern<<<grid, kCpThreads, S::total, st>>>(static_cast<const ActT*>(x), static_cast<const ActT*>(dt_save), K, accA,
                                           static_cast<const ActT*>(dy), static_cast<const ActT*>(t_save), N, accB, T, r,
                                           per);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT>
int dab_typed(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
              int r, float scale, void* dA, void* dB, float* ws, cudaStream_t st) {
  float* accA = ws;          // [K, r]
  float* accB = ws + K * r;  // [N, r]
  VFT_CUDA_OK(cudaMemsetAsync(ws, 0, sizeof(float) * (size_t)(K + N) * r, st));
  if (T > 0) {
    int rc;
    if (r <= 16) rc = launch_colproj<ActT, 2>(dy, x, t_save, dt_save, T, N, K, r, accA, accB, st);
    else if (r <= 32) rc = launch_colproj<ActT, 4>(dy, x, t_save, dt_save, T, N, K, r, accA, accB, st);
    else rc = launch_colproj<ActT, 8>(dy, x, t_save, dt_save, T, N, K, r, accA, accB, st);
    if (rc != VFT_OK) return rc;
  }
  const int64_t total = (K > N ? K : N) * r;
  lora_dab_finalize_kernel<ActT><<<(unsigned)ceil_div64(total, 256), 256, 0, st>>>(
      accA, accB, N, K, r, scale, static_cast<ActT*>(dA), static_cast<ActT*>(dB));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

bool mma_lora_supported(int act_dtype, int64_t C, int r) {
  return (act_dtype == VFT_BF16 || act_dtype == VFT_F16) && C % 8 == 0 && r >= 1 && r <= VFT_LORA_LD;
}

// t_save[T, 64] = x[T, K] . A[r, K]^T
int mma_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                  cudaStream_t st) {
  if (T == 0) return VFT_OK;
  if (!mma_lora_supported(act_dtype, K, r) || !aligned16(x) || !aligned16(a) || !aligned16(t_save))
    return simt_lora_down(x, a, T, K, r, act_dtype, t_save, st);
  if (act_dtype == VFT_BF16) return rowproj_rank<__nv_bfloat16, false>(x, a, T, K, r, 1.0f, t_save, st);
  return rowproj_rank<__half, false>(x, a, T, K, r, 1.0f, t_save, st);
}

// dt_save[T, 64] = scale * dy[T, N] . B[N, r]
int mma_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
                cudaStream_t st) {
  if (T == 0) return VFT_OK;
  if (!mma_lora_supported(act_dtype, N, r) || r % 4 != 0 || !aligned16(dy) || !aligned16(b) || !aligned16(dt_save))
    return simt_lora_dt(dy, b, T, N, r, scale, act_dtype, dt_save, st);
  if (act_dtype == VFT_BF16) return rowproj_rank<__nv_bfloat16, true>(dy, b, T, N, r, scale, dt_save, st);
  return rowproj_rank<__half, true>(dy, b, T, N, r, scale, dt_save, st);
}

int mma_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
            int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st) {
  if (!mma_lora_supported(act_dtype, K, r) || N % 8 != 0 || !aligned16(dy) || !aligned16(x) || !aligned16(t_save) ||
      !aligned16(dt_save))
    return simt_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, ws, st);
  if (act_dtype == VFT_BF16) return dab_typed<__nv_bfloat16>(dy, x, t_save, dt_save, T, N, K, r, scale, dA, dB, ws, st);
  return dab_typed<__half>(dy, x, t_save, dt_save, T, N, K, r, scale, dA, dB, ws, st);
}

}  // namespace vft
