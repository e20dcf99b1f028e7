// Rank-r adapter side products on tcgen05 + TMA (sm_100a):
//
//   kRowNT :  out[T, 64] = ActT( scale * M[T, C] . V[r, C]^T )        t  = x . A^T        (forward)
//   kRowNN :  out[T, 64] = ActT( scale * M[T, C] . V[C, r]   )        dt = s * dy . B     (backward)
//   kCol   :  acc[C, 64] = M[T, C]^T . V[T, 64]                       dA^T = x^T . dt,  dB = s * dy^T . t
//
// These are the pieces of the autograd of /root/reference/src/modules/peft/lora.py:100-104 that do not fit the
// big fused GEMM.  Each is one pass over x or dy (25 MB at 3072 x 3072 x 4096 tokens) against a <= 64-wide
// second operand: bandwidth-bound by construction.  The mma.sync versions (lora_mma.cu) needed 20-29 us per pass
// because two-warp CTAs spend ~1250 cycles of instruction latency per 8 KB chunk (address arithmetic + cp.async
// issue + ldmatrix + dependent MMAs; ncu: 5 warps per SM, 22 % issue-active, profiles/r01_ncu_adapter_*).  Here
// one thread issues TMA boxes, one thread issues tcgen05.mma (M = 128, N = 64, K = 16: 32 cycles per 64-deep step
// against ~260 cycles of load time), and nothing else runs until the epilogue, so the copy engine streams at full
// rate.  The contraction is split over the CTAs of a cluster (gridDim.y = 1, 2, 4 or 8) to put ~150-300 CTAs on the
// machine; partial [128 x 64] fp32 tiles meet in the rank-0 CTA through distributed shared memory (no workspace,
// no atomics, deterministic).
//
// Layouts (all SWIZZLE_128B, 128-byte rows, TMA zero-fills everything outside the tensors):
//   kRowNT/kRowNN  A = M tile [128 tokens x 64 c], K-major.   B = V: NT [64 rows j x 64 c] K-major (rows >= r are
//                  out of bounds -> 0); NN [64 c rows x 64 j] MN-major (columns >= r out of bounds -> 0).
//   kCol           A = M^T tile: two atoms [64 tokens x 64 c], MN-major.   B = V tile [64 tokens x 64 j], MN-major.
#include <stdlib.h>

#include <type_traits>

#include "nf4_lut.cuh"
#include "ptx_sm100.cuh"
#include "tensor_map.cuh"
#include "vft_common.cuh"

namespace vft {
namespace {

enum SideMode { kRowNT = 0, kRowNN = 1, kCol = 2 };

constexpr int kTile = 128;                    // MMA M: tokens (row modes) or columns of M (kCol)
constexpr int kStep = 64;                     // contraction elements per ring stage
constexpr int kABytes = kTile * kStep * 2;    // 16 KB
constexpr int kBBytes = 64 * kStep * 2;       // 8 KB
constexpr int kStageBytes = kABytes + kBBytes;
constexpr int kSideStages = 8;
constexpr int kSideThreads = 192;             // warp 0 TMA, warp 1 MMA, warps 2..5 epilogue (TMEM quadrants 2,3,0,1)
constexpr int kSideTmemCols = 64;
constexpr int kSideBars = 2 * kSideStages + 1;
constexpr int kSideDyn = kSideStages * kStageBytes + kSideBars * 8 + 16 + 1024;
constexpr int kMaxSplit = 8;

struct SideParams {
  int64_t T, C;       // rows of M, columns of M
  int r;
  float scale;
  void* out;          // row modes: [T, 64] ActT.  kCol: dA [r, C] (transposed store) or dB [C, r]
  int col_transposed; // kCol: 1 = store acc[c][j] to out[j * C + c] (dA), 0 = out[c * r + j] (dB)
};

// second problem of a kCol launch (dB next to dA): CTAs with blockIdx.x >= tiles_first take it
struct SidePair {
  SideParams p[2];
  int tiles_first;
  int debug;
};

// triage: wall-clock stamps (ns) of CTA (0,0), enabled with VFT_TC_DEBUG & 16
__device__ unsigned long long g_side_timeline[16];
__device__ __forceinline__ void side_mark(int dbg, int slot) {
  if ((dbg & 16) && blockIdx.x == 0 && blockIdx.y == 0) {
    unsigned long long ns;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns));
    g_side_timeline[slot] = ns;
  }
}

__device__ __forceinline__ uint32_t cluster_ctaid_y() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctaid.y;" : "=r"(r));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t local_addr, uint32_t rank) {
  float4 v;
  const uint32_t remote = ptx::mapa(local_addr, rank);
  asm volatile("ld.shared::cluster.v4.f32 {%0, %1, %2, %3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(remote)
               : "memory");
  return v;
}
__device__ __forceinline__ void st_smem_f4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

template <typename ActT, int kMode, int RP>  // RP: padded rank = MMA N = accumulator columns that are reduced / stored
__global__ void __launch_bounds__(kSideThreads, 1)
lora_side_kernel(const __grid_constant__ CUtensorMap map_m0, const __grid_constant__ CUtensorMap map_v0,
                 const __grid_constant__ CUtensorMap map_m1, const __grid_constant__ CUtensorMap map_v1,
                 const SidePair pp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) side_mark(pp.debug, 0);

  const bool second = kMode == kCol && (int)blockIdx.x >= pp.tiles_first;
  const SideParams& p = pp.p[second ? 1 : 0];
  const CUtensorMap* map_m = second ? &map_m1 : &map_m0;
  const CUtensorMap* map_v = second ? &map_v1 : &map_v0;
  const int tile = second ? (int)blockIdx.x - pp.tiles_first : (int)blockIdx.x;
  const int64_t base = (int64_t)tile * kTile;  // first token (row modes) / first column of M (kCol)

  // contraction range of this CTA: ring steps [k_begin, k_end) of the cluster's total
  const int64_t red = kMode == kCol ? p.T : p.C;
  const int n_total = (int)((red + kStep - 1) / kStep);
  const int n_split = (int)gridDim.y;
  const uint32_t split = cluster_ctaid_y();
  const int per = (n_total + n_split - 1) / n_split;
  const int k_begin = (int)split * per;
  const int k_end = min(n_total, k_begin + per);
  const int n_steps = max(0, k_end - k_begin);

  const uint32_t bar_base = smem_base + kSideStages * kStageBytes;
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (kSideStages + s); };
  const uint32_t bar_acc = bar_base + 8u * (2 * kSideStages);
  const uint32_t tmem_slot = bar_acc + 8;
  volatile uint32_t* tmem_slot_gen =
      reinterpret_cast<volatile uint32_t*>(smem_gen + kSideStages * kStageBytes + 8 * kSideBars);

  if (warp == 0 && ptx::elect_one()) {
    ptx::tma_prefetch_desc(map_m);
    ptx::tma_prefetch_desc(map_v);
  }
  if (warp == 1 && ptx::elect_one()) {
    for (int s = 0; s < kSideStages; ++s) {
      ptx::mbar_init(bar_full(s), 1);
      ptx::mbar_init(bar_empty(s), 1);
    }
    ptx::mbar_init(bar_acc, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 2) ptx::tmem_alloc<kSideTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_gen;
  if (threadIdx.x == 0) side_mark(pp.debug, 1);
  ptx::griddep_launch_dependents();  // programmatic dependent launch: see qlora_tc2.cu
  ptx::griddep_wait();

  if (warp == 0) {
    // ------------------------------------------------------------- TMA producer
    if (ptx::elect_one()) {
      int s = 0;
      uint32_t par = 1;
      for (int i = 0; i < n_steps; ++i) {
        const int kk = (k_begin + i) * kStep;  // first contraction element of this step
        ptx::mbar_wait(bar_empty(s), par);
        const uint32_t a_dst = smem_base + (uint32_t)(s * kStageBytes);
        const uint32_t b_dst = a_dst + kABytes;
        ptx::mbar_arrive_expect_tx(bar_full(s), kStageBytes);
        if (kMode == kCol) {
          ptx::tma_load_2d(map_m, a_dst, bar_full(s), (int)base, kk);              // [64 tokens x 64 c], atom 0
          ptx::tma_load_2d(map_m, a_dst + 8192, bar_full(s), (int)base + 64, kk);  // atom 1
          ptx::tma_load_2d(map_v, b_dst, bar_full(s), 0, kk);                      // [64 tokens x 64 j]
        } else {
          ptx::tma_load_2d(map_m, a_dst, bar_full(s), kk, (int)base);  // [128 tokens x 64 c]
          if (kMode == kRowNT) ptx::tma_load_2d(map_v, b_dst, bar_full(s), kk, 0);  // [64 j x 64 c]
          else ptx::tma_load_2d(map_v, b_dst, bar_full(s), 0, kk);                  // [64 c x 64 j]
        }
        if (++s == kSideStages) {
          s = 0;
          par ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------- MMA issuer (warp-uniform loop, one lane issues)
    constexpr bool kAMn = kMode == kCol;
    constexpr bool kBMn = kMode != kRowNT;
    constexpr uint32_t idesc =
        ptx::make_idesc_f16(std::is_same<ActT, __nv_bfloat16>::value, kAMn, kBMn, kTile, RP);
    constexpr uint32_t kAStep = kAMn ? (2048u >> 4) : (32u >> 4);
    constexpr uint32_t kBStep = kBMn ? (2048u >> 4) : (32u >> 4);
    int s = 0;
    uint32_t par = 0;
    for (int i = 0; i < n_steps; ++i) {
      ptx::mbar_wait(bar_full(s), par);
      ptx::tc_fence_after();
      if (lane == 0 && i == 0) side_mark(pp.debug, 2);
      if (lane == 0 && i == n_steps - 1) side_mark(pp.debug, 3);
      const uint32_t a_addr = smem_base + (uint32_t)(s * kStageBytes);
      const uint64_t a_desc = kAMn ? ptx::make_smem_desc_sw128(a_addr, 8192, 1024)
                                   : ptx::make_smem_desc_sw128(a_addr, 16, 1024);
      const uint64_t b_desc = kBMn ? ptx::make_smem_desc_sw128(a_addr + kABytes, 8192, 1024)
                                   : ptx::make_smem_desc_sw128(a_addr + kABytes, 16, 1024);
      if (ptx::elect_one()) {
#pragma unroll
        for (int k = 0; k < kStep / 16; ++k)
          ptx::umma_ss(tmem_d, a_desc + k * kAStep, b_desc + k * kBStep, idesc, (i | k) != 0 ? 1u : 0u);
        ptx::umma_commit(bar_empty(s));
        if (i == n_steps - 1) ptx::umma_commit(bar_acc);
      }
      __syncwarp();
      if (++s == kSideStages) {
        s = 0;
        par ^= 1u;
      }
    }
  }

  // ------------------------------------------------------------- epilogue warps (one TMEM lane quadrant each)
  const bool epi = warp >= 2;
  const int quad = warp & 3;
  const int row = quad * 32 + lane;  // accumulator lane: token (row modes) / column of M (kCol)
  const uint32_t my_row = smem_base + (uint32_t)row * (uint32_t)(RP * 4);
  float acc[RP];
#pragma unroll
  for (int j = 0; j < RP; ++j) acc[j] = 0.0f;
  if (epi) {
    if (n_steps > 0) {
      ptx::mbar_wait(bar_acc, 0);
      ptx::tc_fence_after();
      uint32_t v[32];
      ptx::tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(quad * 32) << 16), v);
      ptx::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < (RP < 32 ? RP : 32); ++j) acc[j] = __uint_as_float(v[j]);  // columns >= RP: never written
      if (RP > 32) {
        ptx::tmem_ld_32x32b_x32(tmem_d + ((uint32_t)(quad * 32) << 16) + 32u, v);
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 32; j < RP; ++j) acc[j] = __uint_as_float(v[j - 32]);
      }
      ptx::tc_fence_before();
      if (warp == 2 && lane == 0) side_mark(pp.debug, 4);
    }
    if (n_split > 1 && split != 0) {
      // partial tile [128 rows][RP] fp32 at the start of shared memory (the ring is drained: every issued load was
      // consumed by an MMA that has completed); 16-byte chunks XOR-swizzled by the row to spread the banks
#pragma unroll
      for (int c = 0; c < RP / 4; ++c)
        st_smem_f4(my_row + (uint32_t)((c ^ (row & (RP / 4 - 1))) << 4), acc[4 * c], acc[4 * c + 1], acc[4 * c + 2],
                   acc[4 * c + 3]);
    }
  }
  if (n_split > 1) ptx::cluster_sync();  // every thread of every CTA: the partial tiles are visible cluster-wide
  if (warp == 2 && lane == 0) side_mark(pp.debug, 5);

  if (epi && split == 0) {
    for (int rk = 1; rk < n_split; ++rk) {
#pragma unroll
      for (int c = 0; c < RP / 4; ++c) {
        const float4 q = ld_dsmem_f4(my_row + (uint32_t)((c ^ (row & (RP / 4 - 1))) << 4), (uint32_t)rk);
        acc[4 * c] += q.x; acc[4 * c + 1] += q.y; acc[4 * c + 2] += q.z; acc[4 * c + 3] += q.w;
      }
    }
    ActT* out = static_cast<ActT*>(p.out);
    if (kMode == kCol) {
      const int64_t c = base + row;
      if (c < p.C) {
        if (p.col_transposed) {  // dA[j, c]: for each j the warp writes 32 consecutive elements
#pragma unroll
          for (int j = 0; j < RP; ++j)
            if (j < p.r) out[(int64_t)j * p.C + c] = from_f32<ActT>(p.scale * acc[j]);
        } else {  // dB[c, j]: r contiguous values per thread
#pragma unroll
          for (int j = 0; j < RP; ++j)
            if (j < p.r) out[c * p.r + j] = from_f32<ActT>(p.scale * acc[j]);
        }
      }
    } else {
      const int64_t t = base + row;
      if (t < p.T) {  // all 64 columns of the padded row; columns >= r are exact zeros (TMA zero fill)
        uint4* orow = reinterpret_cast<uint4*>(out + t * VFT_LORA_LD);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4 q = make_uint4(0u, 0u, 0u, 0u);
          if (8 * c < RP) {
            q.x = pack2<ActT>(p.scale * acc[(8 * c) % RP], p.scale * acc[(8 * c + 1) % RP]);
            q.y = pack2<ActT>(p.scale * acc[(8 * c + 2) % RP], p.scale * acc[(8 * c + 3) % RP]);
            q.z = pack2<ActT>(p.scale * acc[(8 * c + 4) % RP], p.scale * acc[(8 * c + 5) % RP]);
            q.w = pack2<ActT>(p.scale * acc[(8 * c + 6) % RP], p.scale * acc[(8 * c + 7) % RP]);
          }
          orow[c] = q;
        }
      }
    }
  }

  if (warp == 2 && lane == 0) side_mark(pp.debug, 6);
  // keep every CTA's shared memory alive until rank 0 has read the partial tiles; then release TMEM
  ptx::tc_fence_before();
  if (gridDim.y > 1) ptx::cluster_sync();
  else __syncthreads();
  if (warp == 2) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc<kSideTmemCols>(tmem_d);
  }
  if (threadIdx.x == 0) side_mark(pp.debug, 7);
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
// Cluster size along the contraction: the largest S <= 4 whose clusters all fit in ONE wave (1 CTA per SM: 197 KB
// of shared memory), asked from the occupancy API once per kernel and size (size 2 packs all 74 SM pairs of a B200,
// size 4 places 33 clusters -- GPC sizes strand 16 SMs --, size 8 only 15); the rank-0 reduction is serial in S.
template <typename Kern>
static int max_active_clusters(Kern kern, int split) {
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3(1, (unsigned)split);
  lc.blockDim = dim3(kSideThreads);
  lc.dynamicSmemBytes = kSideDyn;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)split;
  attr[0].val.clusterDim.z = 1;
  lc.attrs = attr;
  lc.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &lc) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

template <typename Kern>
static int pick_split(Kern kern, int64_t tiles, int64_t n_steps, int n_sm) {
  if (const int s = env().side_split) {  // triage override
    if (s >= 1 && s <= kMaxSplit) return s <= n_steps ? s : 1;
  }
  static int cap[5] = {0, -1, -1, -1, -1};  // per kernel instantiation: clusters of size S that fit at once
  for (int s = 4; s >= 2; --s) {
    if (s > n_steps) continue;
    if (cap[s] < 0) cap[s] = max_active_clusters(kern, s);
    if (tiles <= cap[s]) return s;
  }
  (void)n_sm;
  return 1;
}

static int sm_count() {
  int dev = 0, n = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

template <typename ActT, int kMode, int RP>
static int launch_side(const CUtensorMap& m0, const CUtensorMap& v0, const CUtensorMap& m1, const CUtensorMap& v1,
                       const SidePair& pp, int64_t tiles, int64_t n_steps, cudaStream_t st) {
  auto kern = lora_side_kernel<ActT, kMode, RP>;
  VFT_OPT_IN_SMEM_ONCE(kern, kSideDyn);
  const int split = pick_split(kern, tiles, n_steps, sm_count());
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)tiles, (unsigned)split);
  lc.blockDim = dim3(kSideThreads);
  lc.dynamicSmemBytes = kSideDyn;
  lc.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = (unsigned)split;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_enabled() ? 2 : 1;
  VFT_CUDA_OK(cudaLaunchKernelEx(&lc, kern, m0, v0, m1, v1, pp));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT>
static CUtensorMapDataType tm_dtype() {
  return std::is_same<ActT, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
}

template <typename ActT, int kMode>
static int launch_side_rank(const CUtensorMap& m0, const CUtensorMap& v0, const CUtensorMap& m1, const CUtensorMap& v1,
                            const SidePair& pp, int64_t tiles, int64_t n_steps, int r, cudaStream_t st) {
  if (r <= 16) return launch_side<ActT, kMode, 16>(m0, v0, m1, v1, pp, tiles, n_steps, st);
  if (r <= 32) return launch_side<ActT, kMode, 32>(m0, v0, m1, v1, pp, tiles, n_steps, st);
  return launch_side<ActT, kMode, 64>(m0, v0, m1, v1, pp, tiles, n_steps, st);
}

template <typename ActT, int kMode>
static int rowproj_tc(const void* M, const void* V, int64_t T, int64_t C, int r, float scale, void* out,
                      cudaStream_t st) {
  CUtensorMap mm, mv;
  int rc = make_map_2d(&mm, tm_dtype<ActT>(), M, (uint64_t)C, (uint64_t)T, (uint64_t)C * 2, kStep, kTile,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  if (kMode == kRowNT)  // V [r, C]: box [64 rows j x 64 c]
    rc = make_map_2d(&mv, tm_dtype<ActT>(), V, (uint64_t)C, (uint64_t)r, (uint64_t)C * 2, kStep, 64,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  else  // V [C, r]: box [64 rows c x 64 j]
    rc = make_map_2d(&mv, tm_dtype<ActT>(), V, (uint64_t)r, (uint64_t)C, (uint64_t)r * 2, 64, kStep,
                     CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  SidePair pp = {};
  pp.p[0] = SideParams{T, C, r, scale, out, 0};
  pp.p[1] = pp.p[0];
  pp.tiles_first = 0x7fffffff;
  pp.debug = env().tc_debug;
  const int64_t tiles = ceil_div64(T, kTile);
  return launch_side_rank<ActT, kMode>(mm, mv, mm, mv, pp, tiles, ceil_div64(C, kStep), r, st);
}

template <typename ActT>
static int dab_tc(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N,
                  int64_t K, int r, float scale, void* dA, void* dB, cudaStream_t st) {
  CUtensorMap mx, mdt, mdy, mt;
  int rc = make_map_2d(&mx, tm_dtype<ActT>(), x, (uint64_t)K, (uint64_t)T, (uint64_t)K * 2, 64, kStep,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  rc = make_map_2d(&mdt, tm_dtype<ActT>(), dt_save, VFT_LORA_LD, (uint64_t)T, VFT_LORA_LD * 2, 64, kStep,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  rc = make_map_2d(&mdy, tm_dtype<ActT>(), dy, (uint64_t)N, (uint64_t)T, (uint64_t)N * 2, 64, kStep,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  rc = make_map_2d(&mt, tm_dtype<ActT>(), t_save, VFT_LORA_LD, (uint64_t)T, VFT_LORA_LD * 2, 64, kStep,
                   CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  SidePair pp = {};
  pp.p[0] = SideParams{T, K, r, 1.0f, dA, 1};   // dA[j, k] = sum_t dt[t, j] x[t, k]   (dt already carries s)
  pp.p[1] = SideParams{T, N, r, scale, dB, 0};  // dB[n, j] = s * sum_t dy[t, n] t[t, j]
  pp.tiles_first = (int)ceil_div64(K, kTile);
  pp.debug = env().tc_debug;
  const int64_t tiles = ceil_div64(K, kTile) + ceil_div64(N, kTile);
  return launch_side_rank<ActT, kCol>(mx, mdt, mdy, mt, pp, tiles, ceil_div64(T, kStep), r, st);
}

static bool al16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace
}  // namespace vft

extern "C" int vft_debug_side_timeline(unsigned long long* out, int n) {
  if (n > 16) n = 16;
  return cudaMemcpyFromSymbol(out, vft::g_side_timeline, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : -3;
}

namespace vft {

// The tcgen05 side kernels take 16-bit activations, 16-byte aligned rows (TMA global strides) and r <= 64.
bool tc_side_supported(int act_dtype, int64_t C, int r) {
  return (act_dtype == VFT_BF16 || act_dtype == VFT_F16) && C % 8 == 0 && C >= 64 && r >= 1 && r <= VFT_LORA_LD;
}

int tc_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                 cudaStream_t st) {
  if (T == 0) return VFT_OK;
  if (!tc_side_supported(act_dtype, K, r) || !al16(x) || !al16(a) || !al16(t_save))
    return mma_lora_down(x, a, T, K, r, act_dtype, t_save, st);
  if (act_dtype == VFT_BF16) return rowproj_tc<__nv_bfloat16, kRowNT>(x, a, T, K, r, 1.0f, t_save, st);
  return rowproj_tc<__half, kRowNT>(x, a, T, K, r, 1.0f, t_save, st);
}

int tc_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
               cudaStream_t st) {
  if (T == 0) return VFT_OK;
  if (!tc_side_supported(act_dtype, N, r) || r % 8 != 0 || !al16(dy) || !al16(b) || !al16(dt_save))
    return mma_lora_dt(dy, b, T, N, r, scale, act_dtype, dt_save, st);
  if (act_dtype == VFT_BF16) return rowproj_tc<__nv_bfloat16, kRowNN>(dy, b, T, N, r, scale, dt_save, st);
  return rowproj_tc<__half, kRowNN>(dy, b, T, N, r, scale, dt_save, st);
}

int tc_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
           int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st) {
  if (T == 0 || !tc_side_supported(act_dtype, K, r) || !tc_side_supported(act_dtype, N, r) || !al16(dy) || !al16(x) ||
      !al16(t_save) || !al16(dt_save))
    return mma_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, ws, st);
  if (act_dtype == VFT_BF16) return dab_tc<__nv_bfloat16>(dy, x, t_save, dt_save, T, N, K, r, scale, dA, dB, st);
  return dab_tc<__half>(dy, x, t_save, dt_save, T, N, K, r, scale, dA, dB, st);
}

}  // namespace vft
