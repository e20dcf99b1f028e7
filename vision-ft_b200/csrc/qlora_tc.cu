// Fused NF4-dequant + LoRA GEMM on the 5th-generation tensor cores (tcgen05, sm_100a).
//
//   forward  (kBackward = false):  Y^T[n, t]  = sum_k W~[n, k] X[t, k]   + sum_j (s B[n, j]) Tm[t, j]  (+ bias[n])
//   backward (kBackward = true ):  dX^T[k, t] = sum_n W~[n, k] dY[t, n]  + sum_j A[j, k] dTm[t, j]
//
// Replaces bitsandbytes' MatMul4Bit.forward/backward (dequantize W to a bf16 copy in HBM, then
// cuBLAS) reached from /root/reference/src/modules/peft/lora.py:93, and the two rank-r GEMMs +
// scale + add of lora.py:100-104.  No bf16 copy of W ever exists outside shared memory.
//
// Tile: 128 "features" (MMA M; out-features forward, in-features backward) x BN tokens (MMA N)
// x 64 contraction elements per pipeline stage.  The weight is the A operand, the activations
// are the B operand ("swap-AB"): the accumulator D[feature, token] lives in tensor memory with
// one feature per TMEM lane, so bias is a per-thread scalar and the epilogue thread that owns a
// lane also owns the quantization row it decodes.
//
// Warp roles (256 threads):
//   warp 0      TMA producer: activation tile [BN x 64] (SWIZZLE_128B, K-major) and the packed
//               4-bit codes of the weight tile (4 KB) into a 4-stage shared-memory ring
//   warp 1      MMA issuer: one elected thread issues tcgen05.mma.kind::f16 (M=128, N=BN, K=16),
//               tcgen05.commit releases the stage / publishes the accumulator
//   warp 2      TMEM allocation / deallocation
//   warps 4-7   decode: 128 threads, each owns one 64-element quantization block per stage:
//               builds the absmax-scaled 16-entry table in registers (nf4_lut.cuh), decodes 64
//               codes with byte permutes and writes the bf16 operand tile in the canonical
//               UMMA shared-memory layout (forward: K-major, backward: MN-major, both 128B
//               swizzle).  After the main loop the same warps run the epilogue
//               (tcgen05.ld -> bias -> convert -> global).
// The adapter enters as ONE extra pipeline stage: activations = saved x.A^T (or dy.B), weight
// tile = scale*B rows (or A rows); its MMAs accumulate into the same TMEM tile.
#include <type_traits>

#include "nf4_lut.cuh"
#include "ptx_sm100.cuh"
#include "vft_common.cuh"

namespace vft {
namespace {

constexpr int kBM = 128;       // features per CTA tile (MMA M)
constexpr int kBK = 64;        // contraction elements per stage (= NF4 blocksize)
constexpr int kStages = 4;
constexpr int kThreads = 256;
constexpr int kDecodeWarp0 = 4;  // warps 4..7 decode + epilogue
constexpr int kATileBytes = kBM * kBK * 2;  // 16 KB operand tile produced by the decode warps
constexpr int kCodeTileBytes = kBM * kBK / 2;  // 4 KB of packed codes per stage

template <int BN>
struct SmemLayout {
  static constexpr int act_bytes = BN * kBK * 2;
  static constexpr int act_off = 0;
  static constexpr int a_off = act_off + kStages * act_bytes;
  static constexpr int code_off = a_off + kStages * kATileBytes;
  static constexpr int bar_off = code_off + kStages * kCodeTileBytes;
  // barriers: full_act[kStages], full_a[kStages], empty[kStages], accum_full, then the TMEM base address
  static constexpr int total = bar_off + (3 * kStages + 1) * 8 + 16;
  static constexpr int dyn_bytes = total + 1024;  // slack to align the base to 1024 B
};

struct TcParams {
  int64_t T, N, K;
  int r;
  int qdtype;
  float scale;
  const float* absmax;
  const void* bias;
  const void* lora_w;  // forward: B [N, r]; backward: A [r, K]
  void* out;           // forward: Y [T, N]; backward: dX [T, K]
};

template <typename ActT, bool kBackward, int BN>
__global__ void __launch_bounds__(kThreads, 1)
qlora_tc_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_codes,
                const __grid_constant__ CUtensorMap map_lora, const TcParams p) {
  using L = SmemLayout<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t OUT = kBackward ? p.K : p.N;  // feature dimension of this GEMM
  const int64_t RED = kBackward ? p.N : p.K;  // contraction dimension
  const int64_t f0 = (int64_t)blockIdx.x * kBM;
  const int64_t t0 = (int64_t)blockIdx.y * BN;
  const int n_main = (int)((RED + kBK - 1) / kBK);
  const int n_blocks = n_main + (p.r > 0 ? 1 : 0);
  const int KB = (int)(p.K / kBK);  // absmax entries per weight row

  auto bar_full_act = [&](int s) { return smem_base + L::bar_off + 8 * s; };
  auto bar_full_a = [&](int s) { return smem_base + L::bar_off + 8 * (kStages + s); };
  auto bar_empty = [&](int s) { return smem_base + L::bar_off + 8 * (2 * kStages + s); };
  const uint32_t bar_accum = smem_base + L::bar_off + 8 * (3 * kStages);
  const uint32_t tmem_slot = bar_accum + 8;
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + L::bar_off + 8 * (3 * kStages) + 8);

  if (warp == 0 && ptx::elect_one()) {
    ptx::tma_prefetch_desc(&map_act);
    ptx::tma_prefetch_desc(&map_codes);
    if (p.r > 0) ptx::tma_prefetch_desc(&map_lora);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(bar_full_act(s), 1);
      ptx::mbar_init(bar_full_a(s), 4);  // one arrive per decode warp
      ptx::mbar_init(bar_empty(s), 1);
    }
    ptx::mbar_init(bar_accum, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<BN>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_gen;

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t phase = 0;
      for (int b = 0; b < n_blocks; ++b) {
        ptx::mbar_wait(bar_empty(s), phase ^ 1u);
        const uint32_t dst_act = smem_base + L::act_off + s * L::act_bytes;
        if (b < n_main) {
          ptx::mbar_arrive_expect_tx(bar_full_act(s), L::act_bytes + kCodeTileBytes);
          ptx::tma_load_2d(&map_act, dst_act, bar_full_act(s), b * kBK, (int)t0);
          const uint32_t dst_code = smem_base + L::code_off + s * kCodeTileBytes;
          if (kBackward)  // codes of W[n-block b (64 rows), k = f0 .. f0+127]  -> [64 rows][64 bytes]
            ptx::tma_load_2d(&map_codes, dst_code, bar_full_act(s), (int)(f0 / 2), b * kBK);
          else            // codes of W[n = f0 .. f0+127, k-block b (64 cols)]  -> [128 rows][32 bytes]
            ptx::tma_load_2d(&map_codes, dst_code, bar_full_act(s), b * (kBK / 2), (int)f0);
        } else {
          ptx::mbar_arrive_expect_tx(bar_full_act(s), L::act_bytes);
          ptx::tma_load_2d(&map_lora, dst_act, bar_full_act(s), 0, (int)t0);
        }
        if (++s == kStages) { s = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_f16(sizeof(ActT) == 2 && std::is_same<ActT, __nv_bfloat16>::value,
                                                     /*a_mn_major=*/kBackward, /*b_mn_major=*/false, kBM, BN);
      int s = 0;
      uint32_t phase = 0;
      for (int b = 0; b < n_blocks; ++b) {
        ptx::mbar_wait(bar_full_act(s), phase);
        ptx::mbar_wait(bar_full_a(s), phase);
        ptx::tc_fence_after();
        const uint32_t a_addr = smem_base + L::a_off + s * kATileBytes;
        const uint32_t b_addr = smem_base + L::act_off + s * L::act_bytes;
        // A: forward  K-major  [128 rows x 128 B], 8-row groups 1024 B apart
        //    backward MN-major [2 atoms of 64 features][64 contraction rows x 128 B]: atoms 8192 B apart,
        //             8-row groups 1024 B apart
        const uint64_t a_desc = kBackward ? ptx::make_smem_desc_sw128(a_addr, 8192, 1024)
                                          : ptx::make_smem_desc_sw128(a_addr, 16, 1024);
        const uint64_t b_desc = ptx::make_smem_desc_sw128(b_addr, 16, 1024);
        const int ksteps = (b < n_main) ? (kBK / 16) : ((p.r + 15) / 16);
        for (int k = 0; k < ksteps; ++k) {
          // advance 16 contraction elements: 32 B inside a K-major swizzle row, 16 rows (2048 B) MN-major
          const uint64_t a_k = a_desc + (uint64_t)(kBackward ? (k * 2048) >> 4 : (k * 32) >> 4);
          const uint64_t b_k = b_desc + (uint64_t)((k * 32) >> 4);
          ptx::umma_ss(tmem_d, a_k, b_k, idesc, (b | k) != 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar_empty(s));  // stage reusable once these MMAs have read it
        if (++s == kStages) { s = 0; phase ^= 1u; }
      }
      ptx::umma_commit(bar_accum);
    }
  } else if (warp >= kDecodeWarp0) {
    // ------------------------------------------------------------- decode warps
    const int m = threadIdx.x - kDecodeWarp0 * 32;  // 0..127
    // forward : thread m owns weight row n = f0 + m, one 64-wide k-block per stage
    // backward: thread m owns contraction row n = 64 b + (m & 63), in-feature half (m >> 6)
    const int row = kBackward ? (m & 63) : m;
    const int half = kBackward ? (m >> 6) : 0;
    const uint32_t code_row_off = kBackward ? (uint32_t)(row * 64 + half * 32) : (uint32_t)(row * 32);
    // swizzle applied by TMA to the code tile (SWIZZLE_64B backward, SWIZZLE_32B forward): 16-byte chunk
    // index ^= address bits [7, 7 + log2(span/16))
    const uint32_t code_xor = kBackward ? (uint32_t)(((row >> 1) & 3) << 4) : (uint32_t)(((row >> 2) & 1) << 4);
    const uint32_t a_row_off = kBackward ? (uint32_t)(half * 8192 + (row >> 3) * 1024 + (row & 7) * 128)
                                         : (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const uint32_t a_xor = (uint32_t)(row & 7) << 4;

    // absmax of the block this thread decodes in stage b
    const int64_t am_feature = kBackward ? (f0 / kBK + half) : 0;  // k-block index (backward)
    auto absmax_at = [&](int b) -> float {
      if (kBackward) {
        const int64_t n = (int64_t)b * kBK + row;
        return (n < p.N && am_feature < KB) ? __ldg(p.absmax + n * KB + am_feature) : 0.0f;
      } else {
        const int64_t n = f0 + row;
        return (n < p.N) ? __ldg(p.absmax + n * KB + b) : 0.0f;
      }
    };

    int s = 0;
    uint32_t phase = 0;
    float am_next = absmax_at(0);
    for (int b = 0; b < n_blocks; ++b) {
      const float am = am_next;
      if (b + 1 < n_main) am_next = absmax_at(b + 1);
      ptx::mbar_wait(bar_full_act(s), phase);
      const uint32_t a_tile = smem_base + L::a_off + s * kATileBytes + a_row_off;
      if (b < n_main) {
        const uint32_t code_addr = smem_base + L::code_off + s * kCodeTileBytes + code_row_off;
        const uint4 q0 = ptx::lds128(code_addr ^ code_xor);
        const uint4 q1 = ptx::lds128((code_addr + 16) ^ code_xor);
        Nf4Lut lut;
        nf4_build_lut<ActT>(am, p.qdtype, lut);
        const uint32_t words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t v[4];
          nf4_decode_word(words[c], lut, v);
          ptx::sts128(a_tile + (((uint32_t)c << 4) ^ a_xor), v[0], v[1], v[2], v[3]);
        }
      } else {
        // adapter stage: forward row n of scale*B (r values), backward row j of A (64 in-features)
        const ActT* lw = static_cast<const ActT*>(p.lora_w);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint32_t v[4] = {0u, 0u, 0u, 0u};
          if (kBackward) {
            const int64_t k = f0 + half * 64 + c * 8;
            if (row < p.r && k < p.K) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(lw + (int64_t)row * p.K + k));
              v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
            }
          } else {
            const int64_t n = f0 + row;
            if (n < p.N) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = c * 8 + 2 * e;
                const float b0 = (j < p.r) ? p.scale * to_f32<ActT>(lw[n * p.r + j]) : 0.0f;
                const float b1 = (j + 1 < p.r) ? p.scale * to_f32<ActT>(lw[n * p.r + j + 1]) : 0.0f;
                v[e] = pack2<ActT>(b0, b1);
              }
            }
          }
          ptx::sts128(a_tile + (((uint32_t)c << 4) ^ a_xor), v[0], v[1], v[2], v[3]);
        }
      }
      ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_full_a(s));
      if (++s == kStages) { s = 0; phase ^= 1u; }
    }

    // ----------------------------------------------------------- epilogue (same 4 warps)
    ptx::mbar_wait(bar_accum, 0);
    ptx::tc_fence_after();
    const int64_t feat = f0 + m;
    const uint32_t lane_base = tmem_d + ((uint32_t)((warp & 3) * 32) << 16);
    float bias_v = 0.0f;
    if (!kBackward && p.bias != nullptr && feat < OUT) bias_v = to_f32<ActT>(static_cast<const ActT*>(p.bias)[feat]);
    ActT* out = static_cast<ActT*>(p.out);
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(lane_base + (uint32_t)c0, v);
      ptx::tmem_ld_wait();
      if (feat < OUT) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const int64_t t = t0 + c0 + j;
          if (t < p.T) out[t * OUT + feat] = from_f32<ActT>(__uint_as_float(v[j]) + bias_v);
        }
      }
    }
    ptx::tc_fence_before();
  }

  __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<BN>(tmem_d);
  }
}

// ---------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = []() -> EncodeTiledFn {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) != cudaSuccess ||
        qres != cudaDriverEntryPointSuccess)
      return nullptr;
    return reinterpret_cast<EncodeTiledFn>(sym);
  }();
  return fn;
}

static int make_map_2d(CUtensorMap* map, CUtensorMapDataType dt, const void* base, uint64_t inner, uint64_t outer,
                       uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer, CUtensorMapSwizzle sw) {
  // The driver entry point needs a current context; a thread whose first CUDA call is this one (an autograd
  // worker running an NF4-only backward) has none until a runtime call binds the primary context.
  static thread_local bool ctx_bound = false;
  if (!ctx_bound) {
    if (cudaFree(nullptr) != cudaSuccess) {
      set_error("no usable CUDA context: %s", cudaGetErrorString(cudaGetLastError()));
      return VFT_ERR_CUDA;
    }
    ctx_bound = true;
  }
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("cuTensorMapEncodeTiled is not available from the CUDA driver");
    return VFT_ERR_CUDA;
  }
  const cuuint64_t dims[2] = {inner, outer};
  const cuuint64_t strides[1] = {row_stride_bytes};
  const cuuint32_t box[2] = {box_inner, box_outer};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult rc = fn(map, dt, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (rc != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner=%llu outer=%llu stride=%llu box=%ux%u)", (int)rc,
              (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)row_stride_bytes, box_inner,
              box_outer);
    return VFT_ERR_CUDA;
  }
  return VFT_OK;
}

template <typename ActT, bool kBackward, int BN>
static int launch_tc(const LayerArgs& a, const void* act, void* out, const void* lora_act, cudaStream_t st) {
  using L = SmemLayout<BN>;
  const int64_t OUT = kBackward ? a.K : a.N;
  const int64_t RED = kBackward ? a.N : a.K;
  const CUtensorMapDataType dt =
      std::is_same<ActT, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap map_act, map_codes, map_lora;
  int rc = make_map_2d(&map_act, dt, act, (uint64_t)RED, (uint64_t)a.T, (uint64_t)RED * 2, kBK, BN,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  if (kBackward)
    rc = make_map_2d(&map_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.packed, (uint64_t)a.K / 2, (uint64_t)a.N,
                     (uint64_t)a.K / 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_64B);
  else
    rc = make_map_2d(&map_codes, CU_TENSOR_MAP_DATA_TYPE_UINT8, a.packed, (uint64_t)a.K / 2, (uint64_t)a.N,
                     (uint64_t)a.K / 2, 32, 128, CU_TENSOR_MAP_SWIZZLE_32B);
  if (rc != VFT_OK) return rc;
  if (a.r > 0) {
    rc = make_map_2d(&map_lora, dt, lora_act, VFT_LORA_LD, (uint64_t)a.T, VFT_LORA_LD * 2, kBK, BN,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
  } else {
    map_lora = map_act;
  }
  TcParams p;
  p.T = a.T; p.N = a.N; p.K = a.K; p.r = a.r; p.qdtype = a.qdtype; p.scale = a.scale;
  p.absmax = a.absmax;
  p.bias = kBackward ? nullptr : a.bias;
  p.lora_w = kBackward ? a.lora_a : a.lora_b;
  p.out = out;
  auto kern = qlora_tc_kernel<ActT, kBackward, BN>;
  VFT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::dyn_bytes));
  dim3 grid((unsigned)ceil_div64(OUT, kBM), (unsigned)ceil_div64(a.T, BN));
  kern<<<grid, kThreads, L::dyn_bytes, st>>>(map_act, map_codes, map_lora, p);
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT, bool kBackward>
static int launch_tc_bn(const LayerArgs& a, const void* act, void* out, const void* lora_act, cudaStream_t st) {
  if (a.T > 128) return launch_tc<ActT, kBackward, 256>(a, act, out, lora_act, st);
  if (a.T > 64) return launch_tc<ActT, kBackward, 128>(a, act, out, lora_act, st);
  return launch_tc<ActT, kBackward, 64>(a, act, out, lora_act, st);
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

bool tc_supported(const LayerArgs& a, bool backward) {
  if (a.act_dtype != VFT_BF16 && a.act_dtype != VFT_F16) return false;
  if (a.blocksize != 64 || a.K % 64 != 0) return false;  // quantization blocks must not span rows
  if (a.T <= 0 || a.T > 0x7fffffff || a.N > 0x7fffffff) return false;
  if (backward && a.N % 8 != 0) return false;  // dY row stride must be a multiple of 16 bytes (TMA)
  if (!aligned16(a.packed)) return false;
  if (a.r > 0 && backward && !aligned16(a.lora_a)) return false;
  return true;
}

int tc_fwd(const LayerArgs& a, const void* x, void* y, const void* t_save, cudaStream_t st) {
  if (!aligned16(x)) { set_error("x must be 16-byte aligned for the TMA path"); return VFT_ERR_INVALID; }
  if (a.act_dtype == VFT_BF16) return launch_tc_bn<__nv_bfloat16, false>(a, x, y, t_save, st);
  return launch_tc_bn<__half, false>(a, x, y, t_save, st);
}

int tc_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st) {
  if (!aligned16(dy)) { set_error("dy must be 16-byte aligned for the TMA path"); return VFT_ERR_INVALID; }
  if (a.act_dtype == VFT_BF16) return launch_tc_bn<__nv_bfloat16, true>(a, dy, dx, dt_save, st);
  return launch_tc_bn<__half, true>(a, dy, dx, dt_save, st);
}

}  // namespace vft
