// Few-token forward of the NF4 (+LoRA) Linear: a packed-weight STREAMING kernel (T <= 8).
//
//     y[t, n] = sum_b absmax[n, b] . ( sum_{k in block b} code[n, k] . x[t, k] )  (+ bias[n])
//               (+ sum_j round(s.B[n, j]) . t_save[t, j])
//
// The adaLN / modulation Linears of the DiTs run on one token per sample (T = batch: AuraFlow modC/modX/modCX
// [18432, 3072], /root/reference/src/models/auraflow/denoiser.py:351-362,442-445 -- 31.9 MB of packed weight each;
// Lumina2 adaLN_modulation [9216, 1024]), i.e. far below the ~74-token crossover of SURVEY.md 8d: the bound is the
// 0.5625 B/parameter weight stream, not the tensor pipe.  The persistent tcgen05 kernel spends such a launch on
// TMEM / split-K bookkeeping (1.0-1.4 TB/s).  A first streaming kernel that kept the register-LUT decode of the
// tcgen05 kernels (bit-identical W~, 2.6 byte permutes per weight) measured 1.6 TB/s with the ALU pipe 58 % busy
// (ncu, profiles/): the permutes, not HBM, were its ceiling (~3 TB/s).  This version moves the decode off the ALU pipe:
//
//   * one shared-memory look-up per packed BYTE: a 256-entry table of (ActT(code[hi nibble]), ActT(code[lo nibble]))
//     pairs, replicated per lane (32 KB: entry v of lane l at word v*32 + l, so a warp's 32 look-ups never conflict),
//     gives an MMA A-operand register directly; address = one IMAD/shift (FMA pipe) + one mask (ALU pipe) per byte;
//   * the block scale is applied AFTER the contraction over the block (fp32): a quad of threads shares one 64-element
//     block per 4 MMAs, so acc += absmax[row, block] * acc_block costs 4 FMAs per thread per block.  Numerically
//     this is W~ = ActT(code) * absmax in fp32 instead of bitsandbytes' ActT(code * absmax): both sit within one ActT
//     rounding of the exact product, the outputs agree to ~1e-3 relative (tolerances in tests/test_gpu_parity.py);
//     the few-token path is therefore NOT bit-identical to the tcgen05 kernels, it is slightly closer to fp64;
//   * multiply-adds on the warp-level tensor path (mma.sync m16n8k16: 16 weight rows as M, <= 8 tokens as N); the
//     contraction index is only a label, so thread t of a quad feeds its k-slots from ITS OWN 8 contiguous bytes of
//     the block (one 8-byte load per row and block, a quad reads the 32 bytes of a block contiguously);
//   * x is staged once per CTA into shared memory in exactly the fragment order ([block][mma][t][token] x 8 bytes);
//   * weights are read ONCE, straight from the checkpoint layout (row-major packed codes), evict-first.  Two feeds:
//     - TMA (default when K % 1024 == 0 and the rings fit in shared memory; qlora_gemv_tma_kernel): 32 warps per SM,
//       each with a private two-stage ring of [16 rows x 128 B] code boxes (128-byte swizzle) + [16 x 4] statistics
//       boxes completing on the stage's mbarrier; nothing is staged in registers, so the kernel fits 64 registers;
//     - registers (fallback; qlora_gemv_kernel): three 4-block items in flight per thread, 16 warps per SM.
//     Measured on [N, 3072] at T = 2, HBM-cold: TMA 14.7 / 42.0 / 77.2 us at N = 18432 / 73728 / 147456, registers
//     15.5 / 52.8 / 100.2 us -- i.e. 3.6 TB/s (56 % of HBM) vs 2.7 TB/s asymptotically, plus ~6 us per launch that
//     does not scale (launch, table build, x staging, pipeline fill and drain): at the 31.9 MB of an AuraFlow
//     modulation weight the fixed part is 40 % of the launch.  Tried without effect on that size: L2 bulk prefetch of
//     the CTA's future tiles, 2 vs 3 items in flight, one accumulator per MMA instead of a chain of four, 16 vs 32
//     warps, K-slices interleaved item by item (DRAM locality) -- see DESIGN.md 4.4;
//   * a CTA = 4 independent row-tile groups x 4 contraction slices (16 warps); the slices of a tile meet in shared
//     memory (parity double buffer, one named barrier per group), bias and the rank-r adapter term are added there.
//
// Algorithmic bytes per launch: N*K*0.5625 (codes + absmax); x, y, bias, B are negligible.
#include <stdlib.h>

#include <type_traits>

#include "nf4_lut.cuh"
#include "ptx_sm100.cuh"
#include "tensor_map.cuh"
#include "vft_common.cuh"

namespace vft {
namespace {

constexpr int kGemvThreads = 512;
constexpr int kKSlices = 4;                               // warps sharing one 16-row tile along the contraction
constexpr int kGroups = kGemvThreads / 32 / kKSlices;     // 4 row-tile groups per CTA
constexpr int kItemBlocks = 4;                            // quantization blocks per in-flight item
constexpr int kLutBytes = 256 * 32 * 4;
constexpr int kRedFloats = 2 * kGroups * kKSlices * 16 * 8;

template <typename ActT>
__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                         uint32_t b1) {
  if (std::is_same<ActT, __nv_bfloat16>::value) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  } else {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
}

struct GemvArgs {
  const void* x;
  const uint8_t* packed;
  const float* absmax;
  const void* bias;
  const void* lora_b;
  const void* t_save;
  void* y;
  int T, N, K, r, qdtype;
  float scale;
};

struct Item {          // one thread's share of 4 consecutive blocks of rows g and g+8
  uint2 lo[kItemBlocks], hi[kItemBlocks];
  float am_lo, am_hi;  // lane t of the quad holds the statistics of block t of the item
};

__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}

// TP: token slots staged (power of two >= T)
// kFull: every item holds 4 valid blocks (K % 1024 == 0) -- no per-block branches, so the four blocks of an item
// (four independent MMA chains, 64 independent look-ups) are scheduled together
template <typename ActT, int TP, bool kFull>
__global__ void __launch_bounds__(kGemvThreads, 1) qlora_gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // the table sits on a 32 KB boundary of the shared window, so "base + byte*128 + lane*4" is one AND-OR
  const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + (kLutBytes - 1)) & ~(uint32_t)(kLutBytes - 1)) - raw_addr);
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem);                       // [256][32 lanes]
  float* red = reinterpret_cast<float*>(smem + kLutBytes);                 // [2 parity][groups][slices][16][8]
  uint2* xs = reinterpret_cast<uint2*>(smem + kLutBytes + kRedFloats * 4); // [K/64][4 mma][4 t][TP] x 8 bytes
  const int K = a.K, N = a.N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int group = warp / kKSlices, slice = warp % kKSlices;
  const int n_tiles = N / 16;
  const int nbs = K / 64 / kKSlices;                         // blocks per slice (K % 256 == 0)
  const int ipt = (nbs + kItemBlocks - 1) / kItemBlocks;     // items per tile per warp
  // tile visited by this group at iteration it: (it * kGroups + group) * gridDim.x + blockIdx.x
  const int first_tile = group * (int)gridDim.x + (int)blockIdx.x;
  const int tile_stride = kGroups * (int)gridDim.x;
  const int my_tiles = first_tile < n_tiles ? (n_tiles - 1 - first_tile) / tile_stride + 1 : 0;
  const int n_items = my_tiles * ipt;
  const size_t row_bytes = (size_t)(K / 2), row_blocks = (size_t)(K / 64);
  const int blk0 = slice * nbs;                              // first block of this warp's slice
  const uint8_t* col_codes = a.packed + (size_t)blk0 * 32 + 8 * t;
  const float* col_absmax = a.absmax + blk0 + t;

  auto load_item = [&](int it, int q, Item& b) {  // q: item index inside the tile
    const int row = (first_tile + it * tile_stride) * 16 + g;
    const uint8_t* p0 = col_codes + (size_t)row * row_bytes + (size_t)q * (kItemBlocks * 32);
    const uint8_t* p1 = p0 + 8 * row_bytes;
#pragma unroll
    for (int i = 0; i < kItemBlocks; ++i) {
      if (kFull || q * kItemBlocks + i < nbs) {
        b.lo[i] = __ldcs(reinterpret_cast<const uint2*>(p0 + 32 * i));
        b.hi[i] = __ldcs(reinterpret_cast<const uint2*>(p1 + 32 * i));
      } else {
        b.lo[i] = b.hi[i] = make_uint2(0u, 0u);
      }
    }
    const float* q0 = col_absmax + (size_t)row * row_blocks + q * kItemBlocks;
    const bool ok = kFull || q * kItemBlocks + t < nbs;
    b.am_lo = ok ? __ldcs(q0) : 0.0f;
    b.am_hi = ok ? __ldcs(q0 + 8 * row_blocks) : 0.0f;
  };

  // Programmatic dependent launch: the CTAs of this kernel are scheduled under the tail of the previous one and
  // park here.  Nothing is read before the wait: the packed weight and its statistics may have been written by the
  // kernel right before this one (first use after quantize / de-nest).
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();
  // ---- the first items leave for HBM before the table is built and x is staged
  Item bufA, bufB, bufC;
  if (n_items > 0) load_item(0, 0, bufA);
  if (n_items > 1) load_item(1 / ipt, 1 % ipt, bufB);
  if (n_items > 2) load_item(2 / ipt, 2 % ipt, bufC);

  // ---- byte -> (code[hi nibble], code[lo nibble]) pairs in ActT, one copy per lane
  {
    constexpr float kCode[16] = VFT_NF4_CODEBOOK;
    float* s_code = reinterpret_cast<float*>(xs);  // scratch: x is staged after the table is built
    if (threadIdx.x < 16) {
      float v = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) v = threadIdx.x == i ? kCode[i] : v;
      s_code[threadIdx.x] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 256 * 32; e += kGemvThreads) {
      const int v = e >> 5;  // element 2j sits in the HIGH nibble: it is the first (low) half of the pair
      lut[e] = pack2<ActT>(s_code[v >> 4], s_code[v & 15]);
    }
    __syncthreads();
  }
  // ---- stage x in fragment order: unit u = ((blk*4 + j)*4 + t)*TP + g  <-  x[g][blk*64 + 16t + 4j .. +3]
  {
    const ActT* x = static_cast<const ActT*>(a.x);
    const int units = (K / 4) * TP;
#pragma unroll 4
    for (int u = threadIdx.x; u < units; u += kGemvThreads) {
      const int gg = u % TP, q = u / TP;
      const int tt = q & 3, jj = (q >> 2) & 3, bb = q >> 4;
      uint2 v = make_uint2(0u, 0u);
      if (gg < a.T) v = __ldg(reinterpret_cast<const uint2*>(x + (size_t)gg * K + bb * 64 + 16 * tt + 4 * jj));
      xs[u] = v;
    }
  }
  __syncthreads();

  float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  // lanes whose token slot does not exist read slot g % TP: their accumulator columns are never stored, and the
  // columns of an MMA are independent, so no predicate is needed
  const uint2* x_lane = xs + (size_t)t * TP + (g & (TP - 1));
  const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + 4u * (uint32_t)lane;
  // table address of byte i of a word: lut_lane | byte * 128 (lut_lane has bits 7..14 clear: 32 KB alignment)
  auto e0 = [&](uint32_t w) { return lds32(((w << 7) & 0x7f80u) | lut_lane); };
  auto e1 = [&](uint32_t w) { return lds32(((w >> 1) & 0x7f80u) | lut_lane); };
  auto e2 = [&](uint32_t w) { return lds32(((w >> 9) & 0x7f80u) | lut_lane); };
  auto e3 = [&](uint32_t w) { return lds32(((w >> 17) & 0x7f80u) | lut_lane); };

  auto process_item = [&](int it, int q, const Item& b) {
#pragma unroll
    for (int i = 0; i < kItemBlocks; ++i) {
      const int blk_in_slice = q * kItemBlocks + i;
      if (kFull || blk_in_slice < nbs) {  // warp-uniform
        const float am_lo = __shfl_sync(0xffffffffu, b.am_lo, (lane & ~3) | i);
        const float am_hi = __shfl_sync(0xffffffffu, b.am_hi, (lane & ~3) | i);
        const uint2* xg = x_lane + (size_t)(blk0 + blk_in_slice) * (16 * TP);  // + j * 4 * TP per mma
        float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
        const uint2 x0 = xg[0], x1 = xg[4 * TP], x2 = xg[8 * TP], x3 = xg[12 * TP];
        mma16816<ActT>(d, e0(b.lo[i].x), e0(b.hi[i].x), e1(b.lo[i].x), e1(b.hi[i].x), x0.x, x0.y);
        mma16816<ActT>(d, e2(b.lo[i].x), e2(b.hi[i].x), e3(b.lo[i].x), e3(b.hi[i].x), x1.x, x1.y);
        mma16816<ActT>(d, e0(b.lo[i].y), e0(b.hi[i].y), e1(b.lo[i].y), e1(b.hi[i].y), x2.x, x2.y);
        mma16816<ActT>(d, e2(b.lo[i].y), e2(b.hi[i].y), e3(b.lo[i].y), e3(b.hi[i].y), x3.x, x3.y);
        c[0] = fmaf(am_lo, d[0], c[0]);
        c[1] = fmaf(am_lo, d[1], c[1]);
        c[2] = fmaf(am_hi, d[2], c[2]);
        c[3] = fmaf(am_hi, d[3], c[3]);
      }
    }
    // ---- last item of a tile: the four slices of the row tile meet in shared memory
    if (q == ipt - 1) {
      const int tile = first_tile + it * tile_stride;
      float* mine = red + ((((it & 1) * kGroups + group) * kKSlices + slice) * 16) * 8;
      *reinterpret_cast<float2*>(mine + g * 8 + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(mine + (g + 8) * 8 + 2 * t) = make_float2(c[2], c[3]);
      c[0] = c[1] = c[2] = c[3] = 0.0f;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(32 * kKSlices) : "memory");
      const int tid = slice * 32 + lane;  // 0..127 within the row tile
      const int row = tid >> 3, col = tid & 7;
      const int n = tile * 16 + row;
      if (col < a.T) {
        const float* base = red + (((it & 1) * kGroups + group) * kKSlices * 16) * 8 + row * 8 + col;
        float v = 0.0f;
#pragma unroll
        for (int sl = 0; sl < kKSlices; ++sl) v += base[sl * 16 * 8];
        if (a.r > 0) {  // adapter term with the same operand roundings as the tcgen05 kernels: ActT(s.B) x ActT(t)
          const ActT* bw = static_cast<const ActT*>(a.lora_b) + (size_t)n * a.r;
          const ActT* ts = static_cast<const ActT*>(a.t_save) + (size_t)col * VFT_LORA_LD;
          float acc = 0.0f;
          for (int jj = 0; jj < a.r; ++jj)
            acc = fmaf(to_f32<ActT>(from_f32<ActT>(a.scale * to_f32<ActT>(bw[jj]))), to_f32<ActT>(ts[jj]), acc);
          v += acc;
        }
        if (a.bias) v += to_f32<ActT>(static_cast<const ActT*>(a.bias)[n]);
        static_cast<ActT*>(a.y)[(size_t)col * N + n] = from_f32<ActT>(v);
      }
    }
  };

  // ---- three items in flight per thread
  for (int i = 0; i < n_items; i += 3) {
    process_item(i / ipt, i % ipt, bufA);
    if (i + 3 < n_items) load_item((i + 3) / ipt, (i + 3) % ipt, bufA);
    if (i + 1 < n_items) {
      process_item((i + 1) / ipt, (i + 1) % ipt, bufB);
      if (i + 4 < n_items) load_item((i + 4) / ipt, (i + 4) % ipt, bufB);
    }
    if (i + 2 < n_items) {
      process_item((i + 2) / ipt, (i + 2) % ipt, bufC);
      if (i + 5 < n_items) load_item((i + 5) / ipt, (i + 5) % ipt, bufC);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// TMA-fed form (K % 1024 == 0 and the rings fit): the same decode and MMAs, but the packed codes and their statistics
// arrive through the copy engine instead of register loads.  Each warp owns a private ring of kRing stages; a stage is
// one 4-block item of its 16 rows: a [16 rows x 128 B] box of codes (128-byte swizzle, so the 8-byte reads of a
// warp -- rows g / g+8, chunk 2j + t/2 -- spread over all banks: 2 wavefronts per 256 B) and a [16 x 4] box of fp32
// statistics, both completing on the stage's mbarrier.  Lane 0 refills a stage as soon as the warp has consumed it.
// Register loads of this layout cost 80 L1 wavefronts per item (10 instructions x 8 lines); the boxes cost none.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kTmaThreads = 1024;     // 32 warps: nothing is staged in registers, so the kernel fits 64 per thread
constexpr int kRing = 2;
constexpr int kBoxCodes = 16 * 128;   // bytes per stage
constexpr int kBoxStats = 16 * 4 * 4;
constexpr int kWarps = kTmaThreads / 32;
constexpr int kTmaGroups = kWarps / kKSlices;  // 8 row-tile groups per CTA
constexpr int kTmaRedFloats = 2 * kTmaGroups * kKSlices * 16 * 8;

template <typename ActT, int TP>
__global__ void __launch_bounds__(kTmaThreads, 1)
qlora_gemv_tma_kernel(const __grid_constant__ CUtensorMap map_codes, const __grid_constant__ CUtensorMap map_stats,
                      const GemvArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_addr = (uint32_t)__cvta_generic_to_shared(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  // [codes rings: warps x kRing x 2 KB][table 32 KB][stats rings][red][barriers][x]
  uint8_t* ring_codes = smem;
  uint32_t* lut = reinterpret_cast<uint32_t*>(smem + kWarps * kRing * kBoxCodes);
  uint8_t* ring_stats = reinterpret_cast<uint8_t*>(lut) + kLutBytes;
  float* red = reinterpret_cast<float*>(ring_stats + kWarps * kRing * kBoxStats);
  uint64_t* bars = reinterpret_cast<uint64_t*>(red + kTmaRedFloats);
  uint2* xs = reinterpret_cast<uint2*>(bars + kWarps * kRing);
  const int K = a.K, N = a.N;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int group = warp / kKSlices, slice = warp % kKSlices;
  const int n_tiles = N / 16;
  const int nbs = K / 64 / kKSlices;
  const int ipt = nbs / kItemBlocks;  // K % 1024 == 0: every item is full
  const int first_tile = group * (int)gridDim.x + (int)blockIdx.x;
  const int tile_stride = kTmaGroups * (int)gridDim.x;
  const int my_tiles = first_tile < n_tiles ? (n_tiles - 1 - first_tile) / tile_stride + 1 : 0;
  const int n_items = my_tiles * ipt;

  const uint32_t my_codes = (uint32_t)__cvta_generic_to_shared(ring_codes) + (uint32_t)(warp * kRing * kBoxCodes);
  const uint32_t my_stats = (uint32_t)__cvta_generic_to_shared(ring_stats) + (uint32_t)(warp * kRing * kBoxStats);
  const uint32_t my_bars = (uint32_t)__cvta_generic_to_shared(bars) + (uint32_t)(warp * kRing * 8);

  auto issue = [&](int i) {  // lane 0 only: item i of this warp -> stage i % kRing
    const int it = i / ipt, q = i - it * ipt;
    const int st = i % kRing;
    const int row0 = (first_tile + it * tile_stride) * 16;
    const int blk = (q * kKSlices + slice) * kItemBlocks;  // slices interleave item by item: the four warps of a
                                                            // group read 512 adjacent bytes of every row at a time
    const uint32_t bar = my_bars + 8u * st;
    ptx::mbar_arrive_expect_tx(bar, kBoxCodes + kBoxStats);
    ptx::tma_load_2d(&map_codes, my_codes + (uint32_t)(st * kBoxCodes), bar, blk * 32, row0);
    ptx::tma_load_2d(&map_stats, my_stats + (uint32_t)(st * kBoxStats), bar, blk, row0);
  };

  if (lane == 0) {
    for (int st = 0; st < kRing; ++st) ptx::mbar_init(my_bars + 8u * st, 1);
    ptx::fence_mbar_init();
  }
  __syncwarp();
  ptx::griddep_launch_dependents();
  ptx::griddep_wait();  // nothing is read before the wait (see the register-fed kernel)
  if (lane == 0) {
    ptx::tma_prefetch_desc(&map_codes);
    ptx::tma_prefetch_desc(&map_stats);
    for (int i = 0; i < kRing && i < n_items; ++i) issue(i);
  }

  // ---- x: the first units of every thread leave for L2 now and are stored after the table is built (their latency
  // hides under the table build); unit u = ((blk*4 + j)*4 + t)*TP + g  <-  x[g][blk*64 + 16t + 4j .. +3]
  constexpr int kPre = 4;
  const ActT* xsrc = static_cast<const ActT*>(a.x);
  const int units = (K / 4) * TP;
  auto x_unit = [&](int u) {
    const int gg = u % TP, q = u / TP;
    const int tt = q & 3, jj = (q >> 2) & 3, bb = q >> 4;
    uint2 v = make_uint2(0u, 0u);
    if (gg < a.T) v = __ldg(reinterpret_cast<const uint2*>(xsrc + (size_t)gg * K + bb * 64 + 16 * tt + 4 * jj));
    return v;
  };
  uint2 xpre[kPre];
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int u = (int)threadIdx.x + i * kTmaThreads;
    xpre[i] = u < units ? x_unit(u) : make_uint2(0u, 0u);
  }

  // ---- byte -> (code[hi nibble], code[lo nibble]) pairs in ActT, one copy per lane
  {
    constexpr float kCode[16] = VFT_NF4_CODEBOOK;
    float* s_code = reinterpret_cast<float*>(red);  // scratch: the reduction buffer is first used after a barrier
    if (threadIdx.x < 16) {
      float v = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) v = threadIdx.x == i ? kCode[i] : v;
      s_code[threadIdx.x] = v;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < 256 * 32; e += kTmaThreads) {
      const int v = e >> 5;
      lut[e] = pack2<ActT>(s_code[v >> 4], s_code[v & 15]);
    }
  }
#pragma unroll
  for (int i = 0; i < kPre; ++i) {
    const int u = (int)threadIdx.x + i * kTmaThreads;
    if (u < units) xs[u] = xpre[i];
  }
  for (int u = (int)threadIdx.x + kPre * kTmaThreads; u < units; u += kTmaThreads) xs[u] = x_unit(u);
  __syncthreads();

  float c[4] = {0.0f, 0.0f, 0.0f, 0.0f};
  const uint2* x_lane = xs + (size_t)t * TP + (g & (TP - 1));
  const uint32_t lut_lane = (uint32_t)__cvta_generic_to_shared(lut) + 4u * (uint32_t)lane;
  auto e0 = [&](uint32_t w) { return lds32(((w << 7) & 0x7f80u) + lut_lane); };
  auto e1 = [&](uint32_t w) { return lds32(((w >> 1) & 0x7f80u) + lut_lane); };
  auto e2 = [&](uint32_t w) { return lds32(((w >> 9) & 0x7f80u) + lut_lane); };
  auto e3 = [&](uint32_t w) { return lds32(((w >> 17) & 0x7f80u) + lut_lane); };
  // this lane's 8 bytes of block j sit in 16-byte chunk 2j + t/2 of its row, XOR-swizzled by the row (rows g and g+8
  // share g & 7), second half of the chunk for odd t
  const uint32_t row_off = (uint32_t)(g * 128 + ((t & 1) << 3));
  const uint32_t sw = (uint32_t)(g & 7);

  for (int i = 0; i < n_items; ++i) {
    const int it = i / ipt, q = i - it * ipt;
    const int st = i % kRing;
    ptx::mbar_wait(my_bars + 8u * st, (uint32_t)((i / kRing) & 1));
    const uint32_t cs = my_codes + (uint32_t)(st * kBoxCodes) + row_off;
    const uint32_t ss = my_stats + (uint32_t)(st * kBoxStats);
#pragma unroll
    for (int j = 0; j < kItemBlocks; ++j) {
      const uint32_t chunk = ((uint32_t)(2 * j + (t >> 1)) ^ sw) << 4;
      uint2 lo, hi;
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(lo.x), "=r"(lo.y) : "r"(cs + chunk));
      asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(hi.x), "=r"(hi.y) : "r"(cs + 8 * 128 + chunk));
      float am_lo, am_hi;
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(am_lo) : "r"(ss + (uint32_t)((g * 4 + j) * 4)));
      asm volatile("ld.shared.f32 %0, [%1];" : "=f"(am_hi) : "r"(ss + (uint32_t)(((g + 8) * 4 + j) * 4)));
      const uint2* xg = x_lane + (size_t)((q * kKSlices + slice) * kItemBlocks + j) * (16 * TP);
      float d[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const uint2 x0 = xg[0], x1 = xg[4 * TP], x2 = xg[8 * TP], x3 = xg[12 * TP];
      mma16816<ActT>(d, e0(lo.x), e0(hi.x), e1(lo.x), e1(hi.x), x0.x, x0.y);
      mma16816<ActT>(d, e2(lo.x), e2(hi.x), e3(lo.x), e3(hi.x), x1.x, x1.y);
      mma16816<ActT>(d, e0(lo.y), e0(hi.y), e1(lo.y), e1(hi.y), x2.x, x2.y);
      mma16816<ActT>(d, e2(lo.y), e2(hi.y), e3(lo.y), e3(hi.y), x3.x, x3.y);
      c[0] = fmaf(am_lo, d[0], c[0]);
      c[1] = fmaf(am_lo, d[1], c[1]);
      c[2] = fmaf(am_hi, d[2], c[2]);
      c[3] = fmaf(am_hi, d[3], c[3]);
    }
    __syncwarp();  // every lane has consumed the stage: it can be refilled
    if (lane == 0 && i + kRing < n_items) issue(i + kRing);

    if (q == ipt - 1) {  // last item of a tile: the four slices of the row tile meet in shared memory
      const int tile = first_tile + it * tile_stride;
      float* mine = red + ((((it & 1) * kTmaGroups + group) * kKSlices + slice) * 16) * 8;
      *reinterpret_cast<float2*>(mine + g * 8 + 2 * t) = make_float2(c[0], c[1]);
      *reinterpret_cast<float2*>(mine + (g + 8) * 8 + 2 * t) = make_float2(c[2], c[3]);
      c[0] = c[1] = c[2] = c[3] = 0.0f;
      asm volatile("bar.sync %0, %1;" ::"r"(1 + group), "n"(32 * kKSlices) : "memory");
      const int tid = slice * 32 + lane;
      const int row = tid >> 3, col = tid & 7;
      const int n = tile * 16 + row;
      if (col < a.T) {
        const float* base = red + (((it & 1) * kTmaGroups + group) * kKSlices * 16) * 8 + row * 8 + col;
        float v = 0.0f;
#pragma unroll
        for (int sl = 0; sl < kKSlices; ++sl) v += base[sl * 16 * 8];
        if (a.r > 0) {
          const ActT* bw = static_cast<const ActT*>(a.lora_b) + (size_t)n * a.r;
          const ActT* ts = static_cast<const ActT*>(a.t_save) + (size_t)col * VFT_LORA_LD;
          float acc = 0.0f;
          for (int jj = 0; jj < a.r; ++jj)
            acc = fmaf(to_f32<ActT>(from_f32<ActT>(a.scale * to_f32<ActT>(bw[jj]))), to_f32<ActT>(ts[jj]), acc);
          v += acc;
        }
        if (a.bias) v += to_f32<ActT>(static_cast<const ActT*>(a.bias)[n]);
        static_cast<ActT*>(a.y)[(size_t)col * N + n] = from_f32<ActT>(v);
      }
    }
  }
}

static size_t gemv_tma_smem(int K, int TP) {
  return 1024 + (size_t)kWarps * kRing * (kBoxCodes + kBoxStats) + kLutBytes + sizeof(float) * kTmaRedFloats +
         (size_t)kWarps * kRing * 8 + (size_t)(K / 4) * TP * 8;
}

template <typename ActT, int TP>
int launch_gemv_tma(const GemvArgs& g, int n_sm, cudaStream_t st) {
  CUtensorMap mc, ms;
  int rc = make_map_2d(&mc, CU_TENSOR_MAP_DATA_TYPE_UINT8, g.packed, (uint64_t)(g.K / 2), (uint64_t)g.N,
                       (uint64_t)(g.K / 2), 128, 16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  rc = make_map_2d(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, g.absmax, (uint64_t)(g.K / 64), (uint64_t)g.N,
                   (uint64_t)(g.K / 64) * 4, 4, 16, CU_TENSOR_MAP_SWIZZLE_NONE);
  if (rc != VFT_OK) return rc;
  auto kern = qlora_gemv_tma_kernel<ActT, TP>;
  VFT_OPT_IN_SMEM_ONCE(kern, VFT_MAX_DYN_SMEM);
  const int n_tiles = g.N / 16;
  const int grid = n_tiles < n_sm ? n_tiles : n_sm;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(kTmaThreads);
  lc.dynamicSmemBytes = gemv_tma_smem(g.K, TP);
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_enabled() ? 1 : 0;
  VFT_CUDA_OK(cudaLaunchKernelEx(&lc, kern, mc, ms, g));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT, int TP>
int launch_gemv_tp(const GemvArgs& g, int n_sm, cudaStream_t st) {
  const int use_tma = env().gemv_tma;  // triage switch
  if (use_tma && g.K % (64 * kKSlices * kItemBlocks) == 0 && gemv_tma_smem(g.K, TP) <= VFT_MAX_DYN_SMEM &&
      (reinterpret_cast<uintptr_t>(g.absmax) & 15u) == 0)
    return launch_gemv_tma<ActT, TP>(g, n_sm, st);
  // (two instead of three items in flight, or one accumulator per MMA instead of a chain of four, measured the same
  // 16.2-17.0 us on [18432, 3072]: neither prefetch depth nor the MMA chain is what bounds the kernel)
  const bool full = g.K % (64 * kKSlices * kItemBlocks) == 0;
  auto kern = full ? qlora_gemv_kernel<ActT, TP, true> : qlora_gemv_kernel<ActT, TP, false>;
  const size_t smem = (size_t)2 * kLutBytes + sizeof(float) * kRedFloats + (size_t)(g.K / 4) * TP * 8;  // incl. alignment slack
  if (full) VFT_OPT_IN_SMEM_ONCE((qlora_gemv_kernel<ActT, TP, true>), VFT_MAX_DYN_SMEM);
  else VFT_OPT_IN_SMEM_ONCE((qlora_gemv_kernel<ActT, TP, false>), VFT_MAX_DYN_SMEM);
  const int n_tiles = g.N / 16;
  int grid = n_tiles < n_sm ? n_tiles : n_sm;  // one CTA per SM; tiles go round-robin over CTAs first, groups second
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)grid);
  lc.blockDim = dim3(kGemvThreads);
  lc.dynamicSmemBytes = smem;
  lc.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_enabled() ? 1 : 0;
  VFT_CUDA_OK(cudaLaunchKernelEx(&lc, kern, g));
  VFT_CUDA_OK(cudaGetLastError());
  return VFT_OK;
}

template <typename ActT>
int launch_gemv(const GemvArgs& g, cudaStream_t st) {
  int dev = 0, n_sm = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  if (g.T <= 1) return launch_gemv_tp<ActT, 1>(g, n_sm, st);
  if (g.T <= 2) return launch_gemv_tp<ActT, 2>(g, n_sm, st);
  if (g.T <= 4) return launch_gemv_tp<ActT, 4>(g, n_sm, st);
  return launch_gemv_tp<ActT, 8>(g, n_sm, st);
}

}  // namespace

// Shapes the streaming kernel takes: 16-bit activations, T <= 8, blocksize 64, K a multiple of 256 (one block per
// quad thread), N a multiple of 16, 16-byte aligned packed rows, x rows 8-byte aligned, staged x within shared memory.
bool gemv_supported(const LayerArgs& a) {
  if (a.act_dtype != VFT_BF16 && a.act_dtype != VFT_F16) return false;
  if (a.T < 1 || a.T > 8 || a.blocksize != 64) return false;
  if (a.K % (64 * kKSlices) != 0 || a.N % 16 != 0 || a.K > 8192 || a.N > (1 << 26)) return false;
  if ((reinterpret_cast<uintptr_t>(a.packed) & 15u) != 0) return false;
  return true;
}

int gemv_fwd(const LayerArgs& a, const void* x, void* y, const void* t_save, cudaStream_t st) {
  if ((reinterpret_cast<uintptr_t>(x) & 7u) != 0) {
    set_error("x must be 8-byte aligned for the streaming kernel");
    return VFT_ERR_INVALID;
  }
  GemvArgs g{x, a.packed, a.absmax, a.bias, a.lora_b, t_save, y, (int)a.T, (int)a.N, (int)a.K, a.r, a.qdtype, a.scale};
  if (a.act_dtype == VFT_BF16) return launch_gemv<__nv_bfloat16>(g, st);
  return launch_gemv<__half>(g, st);
}

}  // namespace vft
