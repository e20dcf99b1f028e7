// Persistent CTA-pair form of the fused NF4-dequant + LoRA GEMM (tcgen05 cta_group::2, sm_100a), used for
// token counts large enough to fill the machine.  Same arithmetic and operand layouts as qlora_tc.cu
// (read its header first); what changes is the schedule:
//
//   * One cluster of two CTAs per TPC, launched once per SM pair and kept resident: each pair walks a static
//     list of output tiles (tile = 256 features x n_acc*N_acc tokens), so barrier/TMEM set-up, the first
//     cold global loads and the tensor-map fetches are paid once per SM instead of once per tile.
//   * A decoded weight tile (128 features x 64 contraction elements per CTA) feeds up to TWO accumulators of
//     N_acc <= 256 tokens each (all 512 TMEM columns).  The first kernel was bound by the decode warps' ALU
//     work (~480 ALU-pipe cycles per tile against 512 tensor cycles at 256 tokens; profiles/r01_*): doubling
//     the tokens per decoded tile halves that ratio, and pairing the CTAs halves the activation traffic through
//     shared memory (each CTA stages N_acc/2 token rows per accumulator; the tensor cores of both SMs read both
//     halves), which is what keeps 128 B/clk of shared-memory bandwidth sufficient.
//   * N_acc is a run-time multiple of 16 chosen on the host so that the tile count is a near-multiple of the
//     number of SM pairs (T = 4096, 3072 features: 2 x 176 tokens -> 144 tiles on 74 pairs, 2 full waves,
//     instead of 96 tiles of 512 tokens = 1.3 waves).
//   * The epilogue has its own four warps and drains accumulator a of tile i (tcgen05.ld -> bias -> convert
//     -> global) while the decode warps and the MMA issuer are already on tile i+1; the issuer only waits for
//     "accumulator a drained" before its first MMA into a.
//
//   * (round 2) The adapter rides the same launch: the rank-r side product (x . A^T / s * dy . B) is computed by every CTA
//     for its share of the token rows next to the first tile's main loop (kSide), the adapter's weight gradients as a
//     column-tile job behind it in the backward (kJob) -- two launches per layer step.  The forward also knows an
//     asymmetric tile (192 + 176 tokens) so that the side product's columns do not cost a wave; few-token problems
//     split the contraction into per-item fp32 slices that a dependent launch adds in split order.
//
// Warp roles (768 threads per CTA, both CTAs of a pair run all roles except the MMA issuer):
//   warp 20     TMA producer: for every stage, n_acc boxes [N_acc/2 tokens x 64] of the activations
//               (SWIZZLE_128B); completion bytes of BOTH CTAs are counted on the leader's barrier
//   warp 23     MMA issuer (leader CTA only): tcgen05.mma.cta_group::2.kind::f16, M = 256, N = N_acc, K = 16;
//               multicast tcgen05.commit frees the stage in both CTAs / publishes the accumulators
//   warp 21     TMEM allocation (512 columns per CTA)
//   warps 16-19 epilogue, one TMEM lane quadrant each
//   warps 0-15  decode: four groups of 128 threads, group g owns pipeline steps g, g+4, ...
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <type_traits>

#include "nf4_lut.cuh"
#include "ptx_sm100.cuh"
#include "tensor_map.cuh"
#include "vft_common.cuh"

namespace vft {
namespace {

constexpr int kBM = 128;   // features per CTA (256 per pair = MMA M)
constexpr int kBK = 64;    // contraction elements per pipeline step (= NF4 blocksize)
constexpr int kGroups = 4;
// Warp numbering follows the issue arbiter, which favours HIGHER warp ids within a scheduler (measured: with the
// MMA issuer in warp 1 under 16 ALU-saturating decode warps it needed ~1400 cycles per pipeline step for 8 MMAs,
// one wait and one commit; profiles/r01_*): the latency-critical single-thread roles sit on top, the decode
// warps at the bottom.
constexpr int kDecWarp0 = 0;                    // warps 0..15: decode
constexpr int kEpiWarp0 = 4 * kGroups;          // warps 16..19: epilogue (warp % 4 = TMEM lane quadrant)
constexpr int kTmaWarp = kEpiWarp0 + 4;         // warp 20: TMA producer
constexpr int kAllocWarp = kEpiWarp0 + 5;       // warp 21: TMEM allocation
constexpr int kStoreWarp = kEpiWarp0 + 6;       // warp 22: issues the TMA stores of the staged output tiles
constexpr int kMmaWarp = kEpiWarp0 + 7;         // warp 23: MMA issuer
constexpr int kThreads = (kEpiWarp0 + 8) * 32;  // 768
// registers per thread after the role split (launch: 80 x 768): 5 x 128 x (88 - 80) <= 128 x (80 - 40)
constexpr int kCtrlRegs = 40, kEpiRegs = 88, kDecRegs = 88;
constexpr int kATileBytes = kBM * kBK * 2;                // 16 KB decoded weight tile
constexpr int kMaxStages = 8;
constexpr int kMaxAcc = 2;
constexpr int kMaxSideRP = 32;  // largest padded rank whose side product (x.A^T / s.dy.B) is computed inside the launch
constexpr int kMaxP0 = 8;       // slots of the side product's own little ring
// TMEM columns.  Backward (weights staged in shared memory): two accumulators of up to 256 columns.  Forward
// (kTmemA): the decoded weight tile itself lives in tensor memory -- 4 ring stages of 32 columns (64 16-bit values
// per lane) above two accumulators of up to 192 columns -- so it costs no shared-memory bandwidth at all: the
// tcgen05.mma reads A from TMEM, and shared memory only carries the activation boxes.  (Measured before this:
// with A in shared memory the pipe ran at ~138 B/clk of demand against 128 B/clk, and the decode warps' stores
// queued for 2-3 k cycles per block.)
constexpr int kTmemAStages = 4;
constexpr int kTmemACol0 = 384;
template <bool kTmemA>
struct AccLayout {
  static constexpr int pitch = kTmemA ? 192 : 256;  // column pitch between the accumulators = max N_acc
};
constexpr int kTmemCols = 512;
constexpr int kSmemLimit = 227 * 1024;
constexpr int kMaxStg = 16;  // output staging tiles
constexpr int kMaxJobUnits = 2;  // dA/dB column tiles one pair can take (each in its own block of r_pad TMEM columns)
constexpr int kBarBytes = (2 * kMaxStages + 2 * kMaxAcc + 2 * kMaxStg + 2 * kMaxP0 + 1 + kMaxJobUnits) * 8 + 16;
constexpr int kJobBox = 64 * 128;  // dA/dB job: one [64 tokens x 64 columns] box of x / dy per CTA and ring step
constexpr int kStgBytes = 32 * kBM * 2;  // one staging tile [32 tokens][128 features] of 16-bit outputs (8 KB)
constexpr int kEpiBytes = 2 * kStgBytes;  // the minimum: two tiles

struct Tc2Params {
  int64_t T, N, K;
  int r;
  int qdtype;
  float scale;
  const uint8_t* packed;  // bitsandbytes layout, or the micro-tiled copy when tiled != 0 (nf4_quant.cu)
  const float* absmax;
  int tiled;
  const void* bias;
  const void* lora_w;  // forward: B [N, r]; backward: A [r, K]
  void* out;           // forward: Y [T, N]; backward: dX [T, K]
  int n_acc, N_acc;    // accumulators per tile, tokens per accumulator (multiple of 16, <= 256)
  int pingpong;        // one accumulator per tile (n_acc == 1), tile i of a pair in accumulator i & 1: the drain of a tile
                       // runs under the next tile's contraction (short contractions, several tiles per pair)
  int N_acc1;          // tokens of accumulator 1 (= N_acc, or narrower: the forward's side product then sits behind it)
  int stages;
  int n_stg;        // output staging tiles (2..16): as many as fit, so that the accumulators drain at the epilogue
                    // warps' speed while the TMA stores trickle out under the next tile's main loop
  int b_bytes;      // bytes of one accumulator's activation box in one CTA: (N_acc / 2) * 128
  int stage_bytes;  // (backward: kATileBytes +) n_acc * b_bytes
  int n_fblk;       // feature blocks of 256
  int n_tiles;      // n_fblk * token blocks
  // split-K (small problems: fewer tiles than SM pairs; all work items resident at once): a work item is
  // (tile, split); split i reduces ring steps [i * k_per, (i + 1) * k_per) of the contraction (the adapter step belongs
  // to the last split) and writes its fp32 partial tile to ITS slice of `partial` [item][token of the tile][256
  // features]; qlora_tc2_finalize_kernel (launched behind this kernel, programmatic dependent launch) adds the slices
  // of every output element IN SPLIT ORDER, adds the bias and converts.  No memset, no atomics: reproducible sums.
  // (Measured alternatives: fp32 atomicAdd into one [T, OUT] buffer + memset + finalize -- three launches, sums that
  // change from run to run; meeting inside the launch on a per-tile counter and reducing there -- the items of a tile
  // finish up to 15 k cycles apart, and the store -> fence -> atomic -> poll -> load chain cost 20 us at T = 528.)
  int n_split, k_per;
  float* partial;
  int debug;        // VFT_TC_DEBUG triage mask (results are garbage when non-zero): 1 = no decode stores,
                    // 2 = no MMAs, 4 = no epilogue stores, 8 = no TMA loads
  // Side product inside the launch ("phase 0", n_split == 1).  The adapter step of every tile needs the rank-r
  // projection of ITS tokens (forward: t = x . A^T, backward: dt = s * dy . B) -- 25 MB of activations against a
  // 16-row operand, a kernel of its own until now (7.9 / 8.5 us at config #1 plus a launch gap each).  Here every CTA
  // of the grid computes the projection of T / gridDim.x token rows ONCE (no redundancy between the feature blocks
  // that share a token block), next to the main loop of its first work item: the producer thread interleaves the
  // loads of [p0_rows x 64] activation boxes + [r_pad/2 x 64] boxes of the rank-r operand into a small ring of
  // their own, the issuer interleaves M = 256 (128 rows per CTA, p0_rows of them real), N = r_pad MMAs into spare
  // TMEM columns, the (otherwise idle) epilogue warps round the result and write it to t_save / dt_save, and a
  // grid-wide counter tells the producers when every row is there -- long before the first adapter step loads it.
  int side;           // 0: t_save / dt_save were written by a side kernel before this launch
  int r_pad;          // 16 or 32: MMA N of the side product
  int la_bytes;       // bytes of one CTA's box of the rank-r operand: (r_pad / 2) * 128
  int p0_rows;        // token rows per CTA (multiple of 8, <= 128): rows [cta * p0_rows, +p0_rows)
  int p0_slots;       // ring slots (<= kMaxP0)
  int p0_slot_bytes;  // p0_rows * 128 + la_bytes
  int p0_per_step;    // side-product steps issued per main ring step (2: done half-way through the first item)
  void* save;         // t_save / dt_save [T, VFT_LORA_LD] (written when side != 0; the adapter step reads it)
  unsigned* sync;     // {arrivals, generation} of the grid-wide counter (self-resetting; csrc pool, one pair per launch)
  void* bt_out;       // forward, optional: s * lora_up.weight^T as [16 * ceil(r / 16), N] for the backward's side product
  void* save_t;       // optional: the side product once more, transposed [r_pad, T] (K-major operand of the dA/dB job)
  // dA/dB job (backward launch with its side product inside): the adapter's weight gradients dA = dt^T . x [r, K] and
  // dB = s * dy^T . t [N, r] are contractions over the TOKENS -- one pass over x and one over dy, a kernel of its own
  // until now (12.9 us + a launch gap at config #1).  Here pair u takes 128 columns of x (dA) or dy (dB) -- unit u, and,
  // when there are more units than pairs (N + K > 128 * pairs: the 8192-wide MLP layers), unit u + pairs as well:
  // after the side product, the same two threads stream [64 tokens x 64 columns] boxes (MN-major A operand) and
  // [r_pad/2 x 64 tokens] boxes of dt^T / t^T (K-major B operand) through the side product's ring and accumulate
  // M = 128, N = r_pad MMAs in spare TMEM columns; the epilogue warps write the result before the last tile's drain.
  int job;            // 0: adapter gradients by vft_lora_bwd_dab after this launch
  int job_units_a;    // units [0, job_units_a): dA column tiles; [job_units_a, job_units): dB column tiles
  int job_units;
  void* job_da;       // dA [r, K]
  void* job_db;       // dB [N, r]
};

// Timeline of the leader CTA of pair 0 for performance triage (VFT_TC_DEBUG & 16): SM clock per event.
constexpr int kTlRows = 7, kTlCols = 256;
__device__ unsigned long long g_tc2_timeline[kTlRows * kTlCols];
__device__ float g_tc2_p0dump[2 * 128 * 32];  // VFT_TC_DEBUG & 512: raw side-product accumulator lanes of pair 0
__device__ __forceinline__ void tl_mark(const Tc2Params& p, int row, int col) {
  if ((p.debug & 16) && blockIdx.x == 0 && col < kTlCols) g_tc2_timeline[row * kTlCols + col] = (unsigned long long)clock64();
}

// Pin a loop-invariant address in a register: without it the compiler rebuilds shared-memory addresses from the
// cluster/CTA special registers and the kernel parameters inside the hot loops (chains of S2UR / LDCU / ULEA with
// their latencies in series, ~150 cycles per barrier operation in the accumulator drain).
__device__ __forceinline__ uint32_t pinned(uint32_t v) {
  asm volatile("mov.b32 %0, %0;" : "+r"(v));
  return v;
}

__device__ __forceinline__ uint4 ldg_stream_u4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

struct JobMaps {  // dA/dB job: x / dy in [64 columns x 64 tokens] boxes, dt^T / t^T in [64 tokens x r_pad/2 rows] boxes
  CUtensorMap m_a, m_b, v_a, v_b;
};

template <typename ActT, bool kBackward, bool kSide, bool kJob = false>
__global__ void __launch_bounds__(kThreads, 1)
qlora_tc2_kernel(const __grid_constant__ CUtensorMap map_act, const __grid_constant__ CUtensorMap map_lora,
                 const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_out16,
                 const __grid_constant__ CUtensorMap map_p0a, const __grid_constant__ CUtensorMap map_p0w,
                 const __grid_constant__ JobMaps jm, const Tc2Params p) {
  static_assert(!kJob || (kBackward && kSide), "the dA/dB job rides the backward launch's side-product machinery");
  constexpr bool kTmemA = !kBackward;  // forward: decoded weights go to tensor memory, backward: shared memory
  constexpr int kAccCols = AccLayout<kTmemA>::pitch;
  constexpr int kAOff = kTmemA ? 0 : kATileBytes;  // offset of the activation boxes inside a stage
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (ptx::smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - ptx::smem_u32(smem_raw));

  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader (issues the MMAs, owns the full/acc_empty barriers)
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t OUT = kBackward ? p.K : p.N;  // feature dimension of this GEMM
  const int64_t RED = kBackward ? p.N : p.K;  // contraction dimension
  const int n_main = (int)((RED + kBK - 1) / kBK);
  const int KB = (int)(p.K / kBK);  // absmax entries per weight row
  const int S = p.stages;
  const int n_pairs = (int)(gridDim.x >> 1);
  const int pair = (int)(blockIdx.x >> 1);
  const int tok_tile = p.N_acc + (p.n_acc == 2 ? p.N_acc1 : 0);
  auto nacc_of = [&](int a) { return a == 0 ? p.N_acc : p.N_acc1; };   // tokens of accumulator a
  auto tok_off = [&](int a) { return a == 0 ? 0 : p.N_acc; };          // its first token inside the tile

  // shared memory: [S stages: decoded weight tile | n_acc activation boxes][epilogue staging][barriers, TMEM slot]
  // (side product: [p0_slots x (activation rows | rank-r operand box)] between the ring and the staging tiles; its
  //  MMAs read 128 rows from the start of a slot whatever p0_rows is -- into the next slots / the staging tiles)
  const int epi_bytes = p.n_stg * kStgBytes;
  const int p0_bytes = kSide ? p.p0_slots * p.p0_slot_bytes : 0;
  const uint32_t p0_base = smem_base + (uint32_t)(S * p.stage_bytes);
  const uint32_t bar_base = smem_base + (uint32_t)(S * p.stage_bytes + p0_bytes + epi_bytes);
  auto bar_full = [&](int s) { return bar_base + 8u * s; };
  auto bar_empty = [&](int s) { return bar_base + 8u * (kMaxStages + s); };
  // accumulators: full[i] (ping-pong: accumulator i; otherwise full[0] for the whole tile) / empty[a]
  auto bar_acc_full = [&](int i) { return bar_base + 8u * (2 * kMaxStages + i); };
  auto bar_acc_empty = [&](int a) { return bar_base + 8u * (2 * kMaxStages + kMaxAcc + a); };
  // output staging tiles: full[b] (the four epilogue warps have written tile b) / empty[b] (its TMA store has read it)
  auto bar_stg_full = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc + b); };
  auto bar_stg_empty = [&](int b) { return bar_base + 8u * (2 * kMaxStages + 2 * kMaxAcc + kMaxStg + b); };
  // side product: full / empty per ring slot (same protocol as the main ring: bytes of both CTAs are counted on the
  // leader's barrier, a multicast commit frees the slot in both) and "done" (multicast commit behind its last MMA)
  constexpr int kBarP0 = 2 * kMaxStages + 2 * kMaxAcc + 2 * kMaxStg;
  auto bar_p0_full = [&](int i) { return bar_base + 8u * (kBarP0 + i); };
  auto bar_p0_empty = [&](int i) { return bar_base + 8u * (kBarP0 + kMaxP0 + i); };
  const uint32_t bar_p0_done = bar_base + 8u * (kBarP0 + 2 * kMaxP0);
  auto bar_job_done = [&](int j) { return bar_base + 8u * (kBarP0 + 2 * kMaxP0 + 1 + j); };  // unit j of this pair
  const uint32_t tmem_slot = bar_base + 8u * (kBarP0 + 2 * kMaxP0 + 1 + kMaxJobUnits);
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem_base));
  auto stage_a = [&](int s) { return smem_base + (uint32_t)(s * p.stage_bytes); };
  auto stage_b = [&](int s, int a) { return smem_base + (uint32_t)(s * p.stage_bytes + kAOff + a * p.b_bytes); };
  auto p0_slot = [&](int i) { return p0_base + (uint32_t)(i * p.p0_slot_bytes); };
  // TMEM columns of the side product: the tail of accumulator 0's pitch (two accumulators: N_acc <= pitch - r_pad)
  // or, with one accumulator, the tail of the second pitch
  // (two accumulators of equal width: the tail of pitch 0; one accumulator, or a narrower second one: of pitch 1)
  const uint32_t p0_col = (uint32_t)((p.n_acc == 2 && p.N_acc1 == p.N_acc ? 1 : 2) * kAccCols - p.r_pad);
  const int n_p0 = kSide ? n_main : 0;  // side-product steps: one per 64 contraction elements
  const bool p0_m128 = p.p0_rows <= 64;
  // dA/dB job of this pair (kJob): unit, whether it is a dA unit, first column of this CTA's 64, contraction steps
  // (unit j of this pair is global unit pair + j * n_pairs; units [0, job_units_a) are dA column tiles, the rest dB)
  const bool job_on = kJob && pair < p.job_units;
  const int job_mine = job_on ? (p.job_units - pair + n_pairs - 1) / n_pairs : 0;  // 1 or 2 (kMaxJobUnits)
  auto job_unit_is_a = [&](int j) { return pair + j * n_pairs < p.job_units_a; };
  auto job_unit_col0 = [&](int j) {
    const int u = pair + j * n_pairs;
    return (u < p.job_units_a ? u : u - p.job_units_a) * kBM + (int)rank * 64;
  };
  const int job_steps = (int)((p.T + 63) / 64);
  // TMEM columns of unit j: [p0_col - (j + 1) * r_pad, + r_pad / 2)
  auto job_col = [&](int j) { return p0_col - (uint32_t)((j + 1) * p.r_pad); };
  // grid-wide counter: generation before anybody of this launch can have arrived (read by the one thread that waits)
  const int n_ctas = (int)gridDim.x;
  // accumulators of a tile that hold at least one real token (all roles derive it the same way)
  auto accs_of = [&](int64_t t0) -> int {
    const int64_t left = p.T - t0;
    return left >= tok_tile ? p.n_acc : (left > p.N_acc ? 2 : 1);
  };

  // work items: item -> (tile, split) -> ring steps [b0, b1) of that tile (step n_main = the adapter step)
  const int n_items = p.n_tiles * p.n_split;
  auto item_tile = [&](int item) { return item / p.n_split; };
  auto item_b0 = [&](int item) { return (item % p.n_split) * p.k_per; };
  auto item_b1 = [&](int item) {
    const int sp = item % p.n_split;
    const int e = min(n_main, (sp + 1) * p.k_per);
    return sp == p.n_split - 1 ? e + (p.r > 0 ? 1 : 0) : e;
  };

  if (warp == kTmaWarp && ptx::elect_one()) {
    if (kSide) {
      ptx::tma_prefetch_desc(&map_p0a);
      ptx::tma_prefetch_desc(&map_p0w);
    }
    if (kJob) {
      ptx::tma_prefetch_desc(&jm.m_a);
      ptx::tma_prefetch_desc(&jm.v_a);
      ptx::tma_prefetch_desc(&jm.m_b);
      ptx::tma_prefetch_desc(&jm.v_b);
    }
    ptx::tma_prefetch_desc(&map_act);
    if (p.r > 0) ptx::tma_prefetch_desc(&map_lora);
    ptx::tma_prefetch_desc(&map_out);
    ptx::tma_prefetch_desc(&map_out16);
  }
  if (warp == kMmaWarp && ptx::elect_one()) {
    for (int s = 0; s < S; ++s) {
      ptx::mbar_init(bar_full(s), 1 + 2 * 4);  // leader's producer (expect_tx) + the decode warps of both CTAs
      ptx::mbar_init(bar_empty(s), 1);         // multicast tcgen05.commit
    }
    for (int a = 0; a < kMaxAcc; ++a) ptx::mbar_init(bar_acc_full(a), 1);
    if (kSide) {
      for (int i = 0; i < p.p0_slots; ++i) {
        ptx::mbar_init(bar_p0_full(i), 1);   // the leader's producer (expect_tx for both CTAs)
        ptx::mbar_init(bar_p0_empty(i), 1);  // multicast tcgen05.commit
      }
      ptx::mbar_init(bar_p0_done, 1);
      for (int j = 0; j < kMaxJobUnits; ++j) ptx::mbar_init(bar_job_done(j), 1);
    }
    for (int a = 0; a < kMaxAcc; ++a) ptx::mbar_init(bar_acc_empty(a), 2 * 4);  // epilogue warps of both CTAs
    for (int b = 0; b < p.n_stg; ++b) {
      ptx::mbar_init(bar_stg_full(b), 4);   // one arrive per epilogue warp
      ptx::mbar_init(bar_stg_empty(b), 1);  // the store-issuing thread
    }
    ptx::fence_mbar_init();
  }
  if (warp == kAllocWarp) ptx::tmem_alloc_pair<kTmemCols>(tmem_slot);
  ptx::tc_fence_before();
  ptx::cluster_sync();  // barrier inits of both CTAs visible before any remote arrive / multicast commit
  ptx::tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_gen;
  // programmatic dependent launch: everything above (barriers, TMEM, tensor-map fetch) overlapped the tail of the
  // previous kernel.  Each role waits for that kernel itself, as late as it can: the decode warps first put the
  // loads of their first weight block in flight (frozen weights: no dependency), the MMA issuer touches no global
  // memory at all.
  ptx::griddep_launch_dependents();

  // Register budget: 768 threads start with 80 registers each.  The control warpgroup (warps 20-23) gives most of
  // its share back so that the four decode warpgroups can hold a fully decoded block (32 registers) while they
  // wait for their ring stage: decode then overlaps the MMAs that still read the stage.
  // (setmaxnreg sits INSIDE each role's exclusive branch so that ptxas applies the limit to that role only; the
  //  registers a warpgroup gains must have been released by another warpgroup of the SAME CTA.)
  if (warp >= kTmaWarp) {
  ptx::setmaxnreg_dec<kCtrlRegs>();
  if (warp == kTmaWarp) {
    // ------------------------------------------------------------- TMA producer (activations)
    if (ptx::elect_one()) {
      ptx::griddep_wait();  // activations / saved adapter products come from the previous kernels
      int g = 0, s = 0;
      uint32_t empty_par = 1;
      unsigned gen0 = 0;  // generation of the grid-wide counter before anybody of this launch can have arrived
      if (kSide) gen0 = *reinterpret_cast<volatile unsigned*>(p.sync + 1);
      for (int item = pair; item < n_items; item += n_pairs) {
        const int tile = item_tile(item);
        const int64_t t0 = (int64_t)(tile / p.n_fblk) * tok_tile;
        const int na = accs_of(t0);
        const int b1 = item_b1(item);
        for (int b = item_b0(item); b < b1; ++b, ++g) {
          ptx::mbar_wait(bar_empty(s), empty_par);
          tl_mark(p, 5, g);
          if (p.debug & 8) {
            if (rank == 0) ptx::mbar_arrive(bar_full(s));
          } else {
            if (kSide && b == n_main && g == n_main) {
              // first adapter step of this pair: its boxes are rows of t_save / dt_save that CTAs all over the grid
              // are writing in this very launch.  Wait until the generation of the grid-wide counter moves (every
              // epilogue warp of every CTA has published its rows; that happened a tile's worth of time ago).
              const long long t_start = clock64();
              unsigned gen;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(p.sync + 1) : "memory");
                if (clock64() - t_start > 4000000000LL) __trap();
              } while (gen == gen0);
              tl_mark(p, 6, 252);
              asm volatile("fence.proxy.async;" ::: "memory");  // generic-proxy writes of other SMs -> TMA reads
            }
            // the leader arms its barrier for the bytes of BOTH CTAs; each CTA loads its N_acc/2 token rows
            if (rank == 0) ptx::mbar_arrive_expect_tx(bar_full(s), (uint32_t)(2 * na * p.b_bytes));
            const uint32_t leader_bar = ptx::mapa(bar_full(s), 0);
            for (int a = 0; a < na; ++a) {
              // (the box always has N_acc / 2 rows; a narrower accumulator 1 ignores the rows past its own half)
              const int trow = (int)(t0 + tok_off(a)) + (int)rank * (nacc_of(a) >> 1);
              if (b < n_main)
                ptx::tma_load_2d_pair(&map_act, stage_b(s, a), leader_bar, b * kBK, trow);
              else
                ptx::tma_load_2d_pair(&map_lora, stage_b(s, a), leader_bar, 0, trow);
            }
          }
          if (++s == S) {
            s = 0;
            empty_par ^= 1u;
          }
        }
      }
    }
  } else if (kSide && warp == kAllocWarp) {
    // ------------------------------------------------------------- side product: its own issuer (leader CTA)
    // Threads of their own -- this warp (the TMEM-allocation warp has nothing else to do) issues the MMAs, the store
    // warp issues the loads before its first output tile exists.  With the side-product steps interleaved into the
    // main producer and issuer loops every main ring step of the first tile took 1500-2200 cycles instead of 712
    // (an MMA costs the issuing thread ~40 cycles whatever its size, barrier waits come on top); with ONE thread
    // polling both duties a step took ~900 cycles and the slowest CTA of the grid published its rows 43 k cycles
    // into the launch, after the first tile's contraction had ended everywhere.
    if ((rank == 0 || job_on) && ptx::elect_one()) {
      unsigned gen0 = 0;  // generation of the grid-wide counter before anybody of this launch can have arrived
      if (kJob) gen0 = *reinterpret_cast<volatile unsigned*>(p.sync + 1);
      int mm_s = 0;
      uint32_t mm_par = 0;
      if (rank == 0) {
        // M = 128 over the pair (64 rows per CTA) whenever the CTA's token rows fit: the MMA reads M/2 rows of 32
        // bytes per CTA from shared memory whatever number of them is real
        const uint32_t idesc_p0 = ptx::make_idesc_f16(std::is_same<ActT, __nv_bfloat16>::value, false, false,
                                                      p0_m128 ? kBM : 2 * kBM, p.r_pad);
        for (int mm = 0; mm < n_p0; ++mm) {
          ptx::mbar_wait(bar_p0_full(mm_s), mm_par);
          ptx::tc_fence_after();
          const uint64_t xa = ptx::make_smem_desc_sw128(p0_slot(mm_s), 16, 1024);
          const uint64_t wa = ptx::make_smem_desc_sw128(p0_slot(mm_s) + (uint32_t)(p.p0_rows * 128), 16, 1024);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            ptx::umma_ss_pair(tmem_d + p0_col, xa + k * (32u >> 4), wa + k * (32u >> 4), idesc_p0, (mm | k) != 0 ? 1u : 0u);
          ptx::umma_commit_pair(bar_p0_empty(mm_s));
          if (mm == n_p0 - 1) ptx::umma_commit_pair(bar_p0_done);  // -> epilogue warps, both CTAs
          if (++mm_s == p.p0_slots) {
            mm_s = 0;
            mm_par ^= 1u;
          }
        }
      }
      if (job_on) {
        // ---- dA/dB job: this thread issues the loads of its CTA and (leader) the MMAs, polling both duties.  The
        // ring continues where the side product left it: slot and parity after n_p0 steps.
        // One slow thread (~900 cycles per step) is the schedule that measured best (config #1 step through the module
        // API, us): this form 134.3; loads and MMAs on threads of their own, as fast as the data arrives, 145.7 (the
        // backward's main loop is bound by shared-memory bandwidth: every byte the job's MMAs read while it runs is
        // taken from it, and a job that runs early competes with the side product and the cold start); the peer CTA's
        // thread loading for both CTAs through the leader's cluster address 158.8 (that traffic shares the SM-to-SM
        // path with the operand exchange of the cta_group::2 MMAs); MMAs issued by the main issuer thread while it
        // waits for the accumulators to drain, the rest behind its last MMA, 143.0 (an MMA costs the issuing thread
        // ~40 cycles: the job's 256 do not fit into the drain bubbles, and what is left becomes the launch's tail).
        // A smaller ring (28 KB instead of 40: three slots) costs 5 us: the job must not fall behind.
        ptx::griddep_wait();
        // (the peer CTA's thread has issued nothing so far: the ring is the side product's until its last MMA is done)
        if (rank != 0) ptx::mbar_wait(bar_p0_done, 0);
        bool gen_seen = false;
        int ld_s = n_p0 % p.p0_slots;
        uint32_t ld_par = 1u ^ (uint32_t)((n_p0 / p.p0_slots) & 1);
        mm_s = ld_s;
        mm_par = (uint32_t)((n_p0 / p.p0_slots) & 1);
        // M = 128 over the pair: 64 columns per CTA = one MN-major atom [64 tokens x 128 bytes]; N = r_pad; K = tokens
        const uint32_t idesc_job = ptx::make_idesc_f16(std::is_same<ActT, __nv_bfloat16>::value, /*a_mn_major=*/true,
                                                       /*b_mn_major=*/false, kBM, p.r_pad);
        const int total = job_mine * job_steps;  // the units of this pair, one after the other through the same ring
        int ld = 0, mm = 0;
        int ld_j = 0, ld_i = 0, mm_j = 0, mm_i = 0;  // (unit, step inside the unit) of the next load / MMA group
        while (ld < total || (rank == 0 && mm < total)) {
          if (ld < total && ptx::mbar_try_wait(bar_p0_empty(ld_s), ld_par)) {
            const bool is_a = job_unit_is_a(ld_j);
            if (is_a && !gen_seen) {  // dt^T is being written by every CTA of the grid in this very launch
              const long long t_start = clock64();
              unsigned gen;
              do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(p.sync + 1) : "memory");
                if (clock64() - t_start > 4000000000LL) __trap();
              } while (gen == gen0);
              asm volatile("fence.proxy.async;" ::: "memory");
              gen_seen = true;
            }
            const uint32_t dst = p0_slot(ld_s);
            if (rank == 0) ptx::mbar_arrive_expect_tx(bar_p0_full(ld_s), (uint32_t)(2 * (kJobBox + p.la_bytes)));
            const uint32_t leader_bar = ptx::mapa(bar_p0_full(ld_s), 0);
            ptx::tma_load_2d_pair(is_a ? &jm.m_a : &jm.m_b, dst, leader_bar, job_unit_col0(ld_j), ld_i * 64);
            ptx::tma_load_2d_pair(is_a ? &jm.v_a : &jm.v_b, dst + (uint32_t)kJobBox, leader_bar, ld_i * 64,
                                  (int)rank * (p.r_pad >> 1));
            ++ld;
            if (++ld_i == job_steps) {
              ld_i = 0;
              ++ld_j;
            }
            if (++ld_s == p.p0_slots) {
              ld_s = 0;
              ld_par ^= 1u;
            }
          }
          if (rank == 0 && mm < total && ptx::mbar_try_wait(bar_p0_full(mm_s), mm_par)) {
            ptx::tc_fence_after();
            const uint64_t ma = ptx::make_smem_desc_sw128(p0_slot(mm_s), 8192, 1024);
            const uint64_t va = ptx::make_smem_desc_sw128(p0_slot(mm_s) + (uint32_t)kJobBox, 16, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16 tokens: 16 rows of 128 bytes (A), 32 bytes inside a row (B)
              ptx::umma_ss_pair(tmem_d + job_col(mm_j), ma + k * (2048u >> 4), va + k * (32u >> 4), idesc_job,
                                (mm_i | k) != 0 ? 1u : 0u);
            ptx::umma_commit_pair(bar_p0_empty(mm_s));
            ++mm;
            if (++mm_i == job_steps) {  // unit complete -> epilogue warps, both CTAs
              ptx::umma_commit_pair(bar_job_done(mm_j));
              mm_i = 0;
              ++mm_j;
            }
            if (++mm_s == p.p0_slots) {
              mm_s = 0;
              mm_par ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == kMmaWarp) {
    // ------------------------------------------------------------- MMA issuer (leader CTA only)
    // The whole warp walks the loop with warp-uniform state (descriptors stay in uniform registers; inside an
    // elect_one region the compiler has to treat them as per-thread values and pays an R2UR chain per MMA);
    // only the tcgen05 instructions themselves are issued by one elected lane.  Ring stage and parity are
    // tracked incrementally -- the first version spent ~1400 cycles per step in this thread (8 MMAs, two integer
    // divisions, descriptor rebuilds), twice the 704 cycles the MMAs of a 2 x 176-token step need.
    if (rank == 0) {
      const uint32_t idesc = ptx::make_idesc_f16(std::is_same<ActT, __nv_bfloat16>::value,
                                                 /*a_mn_major=*/kBackward, /*b_mn_major=*/false, 2 * kBM, p.N_acc);
      const uint32_t idesc1 = ptx::make_idesc_f16(std::is_same<ActT, __nv_bfloat16>::value,
                                                  /*a_mn_major=*/kBackward, /*b_mn_major=*/false, 2 * kBM, p.N_acc1);
      // advance 16 contraction elements: 32 B inside a K-major swizzle row, 16 rows (2048 B) MN-major
      constexpr uint32_t kAStep = kBackward ? (2048u >> 4) : (32u >> 4);
      constexpr uint32_t kBStep = 32u >> 4;
      const int k_lora = (p.r + 15) / 16;
      const bool do_mma = !(p.debug & 2);
      int g = 0, s = 0;
      uint32_t full_par = 0, acc_par = 1, it_local = 0;
      for (int item = pair; item < n_items; item += n_pairs, ++it_local) {
        const int tile = item_tile(item);
        const int64_t t0 = (int64_t)(tile / p.n_fblk) * tok_tile;
        const int na = accs_of(t0);
        const int b0 = item_b0(item), b1 = item_b1(item);
        // ping-pong: this item's accumulator, and the parity of ITS "drained" barrier (it is used by every other item)
        const int ab = p.pingpong ? (int)(it_local & 1u) : 0;
        const uint32_t d_base = tmem_d + (uint32_t)(ab * kAccCols);
        for (int b = b0; b < b1; ++b, ++g) {
          ptx::mbar_wait(bar_full(s), full_par);
          if (b == b0) {  // the epilogue must have drained the accumulators this work item writes
            if (p.pingpong) ptx::mbar_wait(bar_acc_empty(ab), ((it_local >> 1) & 1u) ^ 1u);
            else for (int a = 0; a < na; ++a) ptx::mbar_wait(bar_acc_empty(a), acc_par);
          }
          ptx::tc_fence_after();
          tl_mark(p, 0, g);
          // A: forward  K-major  [128 rows x 128 B], 8-row groups 1024 B apart
          //    backward MN-major [2 atoms of 64 features][64 contraction rows x 128 B]: atoms 8192 B apart
          const uint64_t a_desc = kBackward ? ptx::make_smem_desc_sw128(stage_a(s), 8192, 1024)
                                            : ptx::make_smem_desc_sw128(stage_a(s), 16, 1024);
          const uint64_t b_desc0 = ptx::make_smem_desc_sw128(stage_b(s, 0), 16, 1024);
          const uint64_t b_desc1 = ptx::make_smem_desc_sw128(stage_b(s, 1), 16, 1024);
          const uint32_t acc0 = b > b0 ? 1u : 0u;
          // forward: A = TMEM stage s, 8 columns (16 packed 16-bit values per lane) per MMA
          const uint32_t a_tmem = tmem_d + (uint32_t)(kTmemACol0 + 32 * s);
          auto mma = [&](uint32_t d, int k, uint64_t b_desc, uint32_t accumulate) {
            const uint32_t id = d == d_base ? idesc : idesc1;
            if (kTmemA) ptx::umma_ts_pair(d, a_tmem + (uint32_t)(8 * k), b_desc + k * kBStep, id, accumulate);
            else ptx::umma_ss_pair(d, a_desc + k * kAStep, b_desc + k * kBStep, id, accumulate);
          };
          if (ptx::elect_one()) {
            if (do_mma) {
              if (b < n_main) {
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k) mma(d_base, k, b_desc0, k > 0 ? 1u : acc0);
                if (na > 1) {
#pragma unroll
                  for (int k = 0; k < kBK / 16; ++k) mma(d_base + kAccCols, k, b_desc1, k > 0 ? 1u : acc0);
                }
              } else {  // adapter step: ceil(r / 16) MMAs per accumulator
                for (int k = 0; k < k_lora; ++k) mma(d_base, k, b_desc0, k > 0 ? 1u : acc0);
                if (na > 1) {
                  for (int k = 0; k < k_lora; ++k) mma(d_base + kAccCols, k, b_desc1, k > 0 ? 1u : acc0);
                }
              }
            }
            ptx::umma_commit_pair(bar_empty(s));  // the stage is reusable in both CTAs once these MMAs have read it
            if (b == b1 - 1) ptx::umma_commit_pair(bar_acc_full(ab));  // item complete -> epilogue warps, both CTAs
          }
          __syncwarp();
          tl_mark(p, 1, g);
          if (++s == S) {
            s = 0;
            full_par ^= 1u;
          }
        }
        acc_par ^= 1u;
      }
    }
  } else if (warp == kStoreWarp) {
    // ------------------------------------------------------------- output store issuer
    // Issuing a TMA store costs one thread ~350 cycles; done by an epilogue thread it sat on the critical path of
    // every 32-token chunk.  This warp mirrors the epilogue's chunk sequence and does nothing but wait for a staged
    // tile, hand it to the TMA engine and release it once the engine has read it.
    if (p.n_split == 1 && ptx::elect_one()) {
      ptx::griddep_wait();  // the output buffer may still be read by the previous kernel
      if (kSide) {
        // side product: this thread is free until the first output tile is staged -- it issues all the loads (step i
        // goes to slot i % p0_slots once the MMAs of step i - p0_slots have read it; bytes of both CTAs are counted
        // on the leader's barrier, as in the main ring)
        int ld_s = 0;
        uint32_t ld_par = 1;
        for (int ld = 0; ld < n_p0; ++ld) {
          ptx::mbar_wait(bar_p0_empty(ld_s), ld_par);
          const uint32_t dst = p0_slot(ld_s);
          // (bytes of the two boxes, not the slot pitch: the dA/dB job's larger boxes can widen the slots)
          if (rank == 0) ptx::mbar_arrive_expect_tx(bar_p0_full(ld_s), (uint32_t)(2 * (p.p0_rows * 128 + p.la_bytes)));
          const uint32_t leader_bar = ptx::mapa(bar_p0_full(ld_s), 0);
          ptx::tma_load_2d_pair(&map_p0a, dst, leader_bar, ld * kBK, (int)blockIdx.x * p.p0_rows);
          ptx::tma_load_2d_pair(&map_p0w, dst + (uint32_t)(p.p0_rows * 128), leader_bar, ld * kBK,
                                (int)rank * (p.r_pad >> 1));
          if (++ld_s == p.p0_slots) {
            ld_s = 0;
            ld_par ^= 1u;
          }
        }
      }
      const uint32_t stg = bar_base - (uint32_t)epi_bytes;
      uint32_t chunk = 0, sb = 0, sphase = 0;  // staging tile of this chunk and its use parity
      for (int item = pair; item < n_items; item += n_pairs) {
        const int tile = item_tile(item);
        const int64_t t0 = (int64_t)(tile / p.n_fblk) * tok_tile;
        const int na = accs_of(t0);
        const int64_t feat0 = (int64_t)(tile % p.n_fblk) * (2 * kBM) + (int64_t)rank * kBM;
        for (int a = 0; a < na; ++a) {
          const int64_t ta = t0 + tok_off(a);
          const int na_cols = nacc_of(a);
          for (int c0 = 0; c0 < na_cols && ta + c0 < p.T; c0 += 32, ++chunk) {
            const uint32_t b = sb;
            ptx::mbar_wait(bar_stg_full(b), sphase);
            if (feat0 < OUT && !(p.debug & 4)) {
              // the staged tile is two halves [32 tokens][64 features] (128-byte rows, SWIZZLE_128B); a chunk cut
              // short by N_acc (a multiple of 16) leaves through the 16-token boxes
              const uint32_t src = stg + b * (uint32_t)kStgBytes;
              const CUtensorMap* m = (c0 + 32 <= na_cols) ? &map_out : &map_out16;
              ptx::tma_store_2d(m, src, (int)feat0, (int)(ta + c0));
              if (feat0 + 64 < OUT) ptx::tma_store_2d(m, src + 4096u, (int)feat0 + 64, (int)(ta + c0));
            }
            ptx::bulk_commit_group();
            if (chunk >= 1) {  // the group committed one chunk ago has finished reading ITS tile
              ptx::bulk_wait_group_read<1>();
              ptx::mbar_arrive(bar_stg_empty(b == 0 ? (uint32_t)p.n_stg - 1u : b - 1u));
            }
            if (++sb == (uint32_t)p.n_stg) {
              sb = 0;
              sphase ^= 1u;
            }
          }
        }
      }
      ptx::bulk_wait_group<0>();  // all output rows written before the CTA retires
    }
  }
  } else if (warp >= kEpiWarp0) {
    // ------------------------------------------------------------- epilogue warps
    ptx::setmaxnreg_inc<kEpiRegs>();
    // TMEM holds D[feature (lane), token (column)] but the output is [token, feature]: each warp converts its
    // 32 features x 32 tokens to 16-bit and writes them TRANSPOSED into a staging tile [32 tokens][128 features]
    // (64 contiguous bytes per token per warp: conflict-free); one thread then hands the tile to the TMA store
    // engine, which writes full 256-byte rows and clips rows >= T / features >= OUT.  Two staging tiles: the
    // store of chunk c overlaps the TMEM load + conversion of chunk c + 1.
    const int quad = warp & 3;  // TMEM lane quadrant this warp may access
    const int et = (warp - kEpiWarp0) * 32 + lane;
    const uint32_t lane_base = pinned(tmem_d + ((uint32_t)(quad * 32) << 16));
    // stmatrix row address of this thread inside a staging tile: matrix k = lane / 8 (features 8k..8k+7 of the
    // quadrant's 32), row i = lane % 8 (token 8q + i); half (quad / 2), 16-byte chunk (quad % 2) * 4 + k of the
    // 128-byte row, XOR-swizzled with the token (SWIZZLE_128B, matching the store's tensor map)
    const int sm_k = lane >> 3, sm_i = lane & 7;
    const uint32_t st_off = (uint32_t)((quad >> 1) * 4096 + sm_i * 128 + ((((quad & 1) * 4 + sm_k) ^ sm_i) << 4));
    const uint32_t epi_smem = pinned(bar_base - (uint32_t)epi_bytes + st_off);
    const uint32_t stg_full0 = pinned(bar_stg_full(0)), stg_empty0 = pinned(bar_stg_empty(0));
    const uint32_t acc_empty_leader = pinned(ptx::mapa(bar_acc_empty(0), 0));  // + 8 a: the leader's barrier
    uint32_t it = 0, sb = 0, sphase = 1;  // staging tile of the next live chunk; parity of its "empty" wait
    ptx::griddep_wait();  // bias / split-K workspace (zeroed by a memset node) / output ordering
    if (kSide) {
      // Side product of this CTA's token rows: lane = row, r_pad fp32 columns.  Round to the activation dtype, write
      // the rows to t_save / dt_save (first r_pad columns; nothing reads the others) and count this warp in: the
      // warp that completes the count resets it and bumps the generation the producers of all CTAs are watching.
      ptx::mbar_wait(bar_p0_done, 0);
      ptx::tc_fence_after();
      if (et == 0) tl_mark(p, 6, 250);
      // M = 256: lane = row, r_pad columns.  M = 128 (measured with tools/p0_layout_probe.py: tcgen05.mma
      // cta_group::2 with M = 128 stores a CTA's 64 rows x N as 128 lanes x N/2 columns): lanes 0..63 hold columns
      // [0, N/2) of rows 0..63, lanes 64..127 columns [N/2, N) of the same rows.
      if ((p.debug & 512) && blockIdx.x < 2) {  // layout probe (tools/p0_layout_probe.py): every lane's 32 columns
        uint32_t v[16];
        for (int c16 = 0; c16 < 32; c16 += 16) {
          ptx::tmem_ld_32x32b_x16(lane_base + p0_col + (uint32_t)c16, v);
          ptx::tmem_ld_wait();
          for (int e = 0; e < 16; ++e)
            g_tc2_p0dump[(blockIdx.x * 128 + quad * 32 + lane) * 32 + c16 + e] = __uint_as_float(v[e]);
        }
      }
      const int half = p0_m128 ? (quad >> 1) : 0;                     // which half of the r_pad columns this warp holds
      const int row = (p0_m128 ? (quad & 1) : quad) * 32 + lane;      // token row of this CTA's slice
      const int n_col = p0_m128 ? (p.r_pad >> 1) : p.r_pad;           // accumulator columns per lane: 8, 16 or 32
      if (row - lane < p.p0_rows) {
        const int64_t tok = (int64_t)blockIdx.x * p.p0_rows + row;
        uint32_t pk[kMaxSideRP / 2];
#pragma unroll
        for (int c16 = 0; c16 < kMaxSideRP; c16 += 16) {
          if (c16 < n_col) {  // (a 16-column load of an 8-column half reads 8 columns nobody wrote: unused)
            uint32_t v[16];
            ptx::tmem_ld_32x32b_x16(lane_base + p0_col + (uint32_t)c16, v);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int e = 0; e < 8; ++e)
              pk[(c16 >> 1) + e] = pack2<ActT>(__uint_as_float(v[2 * e]), __uint_as_float(v[2 * e + 1]));
          }
        }
        if (row < p.p0_rows && tok < p.T) {
          uint4* orow = reinterpret_cast<uint4*>(static_cast<ActT*>(p.save) + tok * VFT_LORA_LD) + half * (n_col >> 3);
#pragma unroll
          for (int c = 0; c < kMaxSideRP / 8; ++c)
            if (c < (n_col >> 3)) orow[c] = make_uint4(pk[4 * c], pk[4 * c + 1], pk[4 * c + 2], pk[4 * c + 3]);
          if (p.save_t != nullptr) {  // the same values as rows [r_pad, T] (K-major operand of the dA/dB job)
            uint16_t* tcol = static_cast<uint16_t*>(p.save_t) + (int64_t)(half * n_col) * p.T + tok;
#pragma unroll
            for (int c = 0; c < kMaxSideRP / 2; ++c) {
              if (2 * c < n_col) {
                tcol[(int64_t)(2 * c) * p.T] = (uint16_t)(pk[c] & 0xffffu);
                tcol[(int64_t)(2 * c + 1) * p.T] = (uint16_t)(pk[c] >> 16);
              }
            }
          }
        }
      }
      __threadfence();
      __syncwarp();
      if (lane == 0) {
        const unsigned total = 4u * (unsigned)n_ctas;
        if (atomicAdd(p.sync, 1u) == total - 1u) {
          atomicExch(p.sync, 0u);  // ready for the next launch that is handed this pair
          __threadfence();
          atomicAdd(p.sync + 1, 1u);
        }
      }
      if (et == 0) tl_mark(p, 6, 251);
    }
    for (int item = pair; item < n_items; item += n_pairs, ++it) {
      const int tile = item_tile(item);
      const int64_t t0 = (int64_t)(tile / p.n_fblk) * tok_tile;
      const int na = accs_of(t0);
      const int64_t feat0 = (int64_t)(tile % p.n_fblk) * (2 * kBM) + (int64_t)rank * kBM;
      // bias of the four feature rows this thread holds fragments of: quadrant base + lane / 4 + {0, 8, 16, 24}
      float bias_v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      const bool has_bias = !kBackward && p.bias != nullptr;
      if (has_bias) {
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int64_t f = feat0 + quad * 32 + (lane >> 2) + 8 * u;
          if (f < OUT) bias_v[u] = to_f32<ActT>(static_cast<const ActT*>(p.bias)[f]);
        }
      }
      if (kJob && job_on && item + n_pairs >= n_items) {
        // dA/dB units of this pair: 64 columns x r_pad accumulators (M = 128 layout: lanes 0..63 hold the first half of
        // the r_pad columns of rows 0..63, lanes 64..127 the second half) -- written while the last tile's MMAs run
        for (int j = 0; j < job_mine; ++j) {
          ptx::mbar_wait(bar_job_done(j), 0);
          ptx::tc_fence_after();
          const bool is_a = job_unit_is_a(j);
          const int L = quad * 32 + lane;
          const int64_t col = job_unit_col0(j) + (L & 63);
          const int jh = (L >> 6) * (p.r_pad >> 1);
          const int64_t C = is_a ? p.K : p.N;
          uint32_t v[16];
          ptx::tmem_ld_32x32b_x16(lane_base + job_col(j), v);
          ptx::tmem_ld_wait();
          if (col < C) {
            if (is_a) {  // dA[j, col]: per j the warp writes 32 consecutive elements
              ActT* o = static_cast<ActT*>(p.job_da) + col;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (e < (p.r_pad >> 1) && jh + e < p.r) o[(int64_t)(jh + e) * p.K] = from_f32<ActT>(__uint_as_float(v[e]));
            } else {     // dB[col, j]: this thread's half row
              ActT* o = static_cast<ActT*>(p.job_db) + col * p.r + jh;
#pragma unroll
              for (int e = 0; e < 16; ++e)
                if (e < (p.r_pad >> 1) && jh + e < p.r) o[e] = from_f32<ActT>(p.scale * __uint_as_float(v[e]));
            }
          }
        }
      }
      if (!kBackward && p.bt_out != nullptr && tile / p.n_fblk == 0) {
        // bt = s * B^T [16 * ceil(r / 16), N], the K-major form of the adapter's up-projection that the backward launch
        // reads for its side product: the tiles of token block 0 cover every out-feature exactly once, and this warp
        // has nothing to do until the tile's contraction ends.  (Written by the decode warps on their way through the
        // adapter step, the 16 two-byte stores per thread cost the tiles of token block 0 ~6 k cycles.)
        const int64_t n = feat0 + quad * 32 + lane;
        if (n < p.N) {
          const ActT* brow = static_cast<const ActT*>(p.lora_w) + n * p.r;
          ActT* bt = static_cast<ActT*>(p.bt_out) + n;
          const int bt_rows = ((p.r + 15) >> 4) << 4;
          for (int j = 0; j < bt_rows; ++j)
            bt[(int64_t)j * p.N] = from_f32<ActT>(j < p.r ? p.scale * to_f32<ActT>(brow[j]) : 0.0f);
        }
      }
      const int ab = p.pingpong ? (int)(it & 1u) : 0;  // ping-pong: this item's accumulator
      ptx::mbar_wait(bar_acc_full(ab), p.pingpong ? ((it >> 1) & 1u) : (it & 1u));
      ptx::tc_fence_after();
      if (et == 0) tl_mark(p, 4, 4 * (int)it + 2);
      if (p.n_split > 1) {
        // split-K: this item's fp32 partial tile goes to its own slice of the workspace, rows = tokens of the tile,
        // 256 features per row (a warp's 32 lanes = 32 consecutive features: 128-byte stores)
        float* slice = p.partial + (int64_t)item * tok_tile * (2 * kBM) + (int)rank * kBM + quad * 32 + lane;
        for (int a = 0; a < na; ++a) {
          const int64_t ta = t0 + tok_off(a);
          const int na_cols = nacc_of(a);
#pragma unroll 1
          for (int c0 = 0; c0 < na_cols; c0 += 16) {  // (a multiple of 16)
            const bool live = ta + c0 < p.T;  // warp-uniform; dead chunks still release the accumulator below
            uint32_t v0[16];  // lane = feature, registers = 16 consecutive token columns
            if (live) {
              ptx::tmem_ld_32x32b_x16(lane_base + (uint32_t)(a * kAccCols + c0), v0);
              ptx::tmem_ld_wait();
            }
            if (c0 + 16 >= na_cols) {  // accumulator a is in registers / not needed: the issuer may overwrite it
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(bar_acc_empty(a), 0));
            }
            if (!live || (p.debug & 4)) continue;
            float* row = slice + (int64_t)(tok_off(a) + c0) * (2 * kBM);
#pragma unroll
            for (int e = 0; e < 16; ++e)
              if (ta + c0 + e < p.T) __stcg(row + e * (2 * kBM), __uint_as_float(v0[e]));
          }
        }
      } else {
        // Drain: per 32-column chunk one load -> convert -> stmatrix -> proxy fence -> publish chain (~520 cycles per
        // chunk and warp: a single warp issues it at a few cycles per instruction).  Measured alternatives, all slower:
        // loads issued one chunk ahead (convert first, then the next tcgen05.ld, then the chain: 6.7 k cycles per
        // 2 x 176-token tile against 5.7 k), a warp-private transpose patch + st.global.v4 (8.2 k), one st.global.b16
        // per token from the 32x32b load shape (21 k).  Letting the idle decode warps take four fifths of the LAST
        // tile's chunks (staging tiles handed out through a counter instead of per-tile mbarriers) shortened that
        // drain from 7.1 k to 3.4 k cycles on the epilogue warp's clock and left the launch where it was (58.0 us with,
        // 58.2 us without): every CTA of the grid drains at the same moment, and 13 MB of output leave in one burst.
        for (int a = 0; a < na; ++a) {
          const int64_t ta = t0 + tok_off(a);
          const int na_cols = nacc_of(a);
#pragma unroll 1
          for (int c0 = 0; c0 < na_cols; c0 += 32) {
            const bool live = ta + c0 < p.T;  // warp-uniform; dead chunks still release the accumulator below
            // 16x256b: mma-style fragments -- r[4q], r[4q+1] = (lane t/4, tokens 8q + 2(t%4), +1); r[4q+2], r[4q+3] = lane + 8
            uint32_t v0[16], v1[16];  // lanes 0..15 / 16..31 of the quadrant, 32 token columns
            if (live) {
              ptx::tmem_ld_16x256b_x4(lane_base + (uint32_t)((ab + a) * kAccCols + c0), v0);
              ptx::tmem_ld_16x256b_x4(lane_base + (16u << 16) + (uint32_t)((ab + a) * kAccCols + c0), v1);
              ptx::tmem_ld_wait();
            }
            if (c0 + 32 >= na_cols) {  // accumulator a is in registers / not needed: the issuer may overwrite it
              ptx::tc_fence_before();
              __syncwarp();
              if (lane == 0) ptx::mbar_arrive_cluster(acc_empty_leader + 8u * (uint32_t)(ab + a));
            }
            if (!live) continue;
            const uint32_t buf = epi_smem + sb * (uint32_t)kStgBytes;
            ptx::mbar_wait(stg_empty0 + 8u * sb, sphase);  // the tile's previous store has read it
            // tokens 8q..8q+7: four 8x8 matrices = feature groups 0-7, 8-15, 16-23, 24-31 (the 32 bias adds are
            // skipped when there is no bias: this chain is issue-bound)
            if (has_bias) {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t m0 = pack2<ActT>(__uint_as_float(v0[4 * q]) + bias_v[0], __uint_as_float(v0[4 * q + 1]) + bias_v[0]);
                const uint32_t m1 = pack2<ActT>(__uint_as_float(v0[4 * q + 2]) + bias_v[1], __uint_as_float(v0[4 * q + 3]) + bias_v[1]);
                const uint32_t m2 = pack2<ActT>(__uint_as_float(v1[4 * q]) + bias_v[2], __uint_as_float(v1[4 * q + 1]) + bias_v[2]);
                const uint32_t m3 = pack2<ActT>(__uint_as_float(v1[4 * q + 2]) + bias_v[3], __uint_as_float(v1[4 * q + 3]) + bias_v[3]);
                ptx::stmatrix_x4_trans(buf + (uint32_t)(q * 1024), m0, m1, m2, m3);
              }
            } else {
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const uint32_t m0 = pack2<ActT>(__uint_as_float(v0[4 * q]), __uint_as_float(v0[4 * q + 1]));
                const uint32_t m1 = pack2<ActT>(__uint_as_float(v0[4 * q + 2]), __uint_as_float(v0[4 * q + 3]));
                const uint32_t m2 = pack2<ActT>(__uint_as_float(v1[4 * q]), __uint_as_float(v1[4 * q + 1]));
                const uint32_t m3 = pack2<ActT>(__uint_as_float(v1[4 * q + 2]), __uint_as_float(v1[4 * q + 3]));
                ptx::stmatrix_x4_trans(buf + (uint32_t)(q * 1024), m0, m1, m2, m3);
              }
            }
            ptx::fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(stg_full0 + 8u * sb);  // the store warp takes it from here
            if (++sb == (uint32_t)p.n_stg) {  // live chunks walk round the staging tiles
              sb = 0;
              sphase ^= 1u;
            }
          }
        }
      }
      if (et == 0) tl_mark(p, 4, 4 * (int)it + 3);
      for (int a = na; a < p.n_acc; ++a) {  // unused accumulators of a ragged last token block: keep phases in step
        if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(bar_acc_empty(a), 0));
      }
    }
  } else {
    // ------------------------------------------------------------- decode warps
    ptx::setmaxnreg_inc<kDecRegs>();
    const int dw = warp - kDecWarp0;  // 0..15
    const int group = dw >> 2;        // owns pipeline steps group, group + 4, ...
    const int quad = dw & 3;
    const int m = quad * 32 + lane;  // 0..127
    // forward : thread m owns weight row n = f0 + m, one 64-wide k-block per step
    // backward: thread m owns contraction row n = 64 b + (m & 63), in-feature half (m >> 6)
    const int row = kBackward ? (m & 63) : m;
    const int half = kBackward ? (m >> 6) : 0;
    const uint32_t a_row_off = kBackward ? (uint32_t)(half * 8192 + (row >> 3) * 1024 + (row & 7) * 128)
                                         : (uint32_t)((row >> 3) * 1024 + (row & 7) * 128);
    const uint32_t a_xor = (uint32_t)(row & 7) << 4;

    // Position in this pair's step sequence, with the per-item part of the weight addresses cached: the integer
    // divisions (item -> tile -> feature block) run once per work item instead of once per 64-element block.
    struct Pos {
      int b;             // ring step of the tile's contraction (n_main = the adapter step)
      int item;          // work item (tile, split)
      int b1;            // end of the item's step range
      int64_t f0;        // first feature of this CTA's half of the tile (forward: out-feature n; backward: in-feature k)
      const uint8_t* c;  // codes of this thread's block at step 0 (nullptr: outside the weight)
      const float* a;    // its absmax
    };
    // per-step strides of the two pointers and the offset of the second 16 bytes of a 32-byte block
    const int64_t c_step = p.tiled ? (kBackward ? (int64_t)KB * 2048 : 2048) : (kBackward ? 32 * p.K : 32);
    const int64_t a_step = p.tiled ? (kBackward ? (int64_t)KB * 64 : 64) : (kBackward ? (int64_t)kBK * KB : 1);
    const int c_second = p.tiled ? 1024 : 16;
    auto enter_item = [&](Pos& q) {
      q.b1 = 0;
      q.c = nullptr;
      q.a = nullptr;
      q.f0 = 0;
      if (q.item >= n_items) return;
      q.b1 = item_b1(q.item);
      q.f0 = (int64_t)(item_tile(q.item) % p.n_fblk) * (2 * kBM) + (int64_t)rank * kBM;
      // element coordinates of this thread's quantization block at step 0 in W [N, K]
      const int64_t wrow = kBackward ? row : q.f0 + row;
      const int64_t wcol = kBackward ? q.f0 + half * 64 : 0;
      if (kBackward ? (wcol >= p.K) : (wrow >= p.N)) return;
      if (p.tiled) {  // 64 x 64 micro-tiles: the warp's 32 rows are contiguous (512 B per load, one line of absmax)
        const int64_t mt = (wrow >> 6) * KB + (wcol >> 6);
        q.c = p.packed + mt * 2048 + (wrow & 63) * 16;
        q.a = p.absmax + mt * 64 + (wrow & 63);
      } else {
        q.c = p.packed + ((wrow * p.K + wcol) >> 1);
        q.a = p.absmax + wrow * KB + (wcol >> 6);
      }
    };
    auto normalize = [&](Pos& q) {  // carry an overflow past the item's last step into the pair's next item(s)
      while (q.item < n_items && q.b >= q.b1) {
        const int over = q.b - q.b1;
        q.item += n_pairs;
        enter_item(q);
        q.b = (q.item < n_items ? item_b0(q.item) : 0) + over;
      }
    };

    uint4 q0 = make_uint4(0, 0, 0, 0), q1 = q0;
    float am = 0.0f;
    auto prefetch = [&](const Pos& q) {
      q0 = q1 = make_uint4(0, 0, 0, 0);
      am = 0.0f;
      if (q.c == nullptr || q.b >= n_main) return;
      if (kBackward && (int64_t)q.b * kBK + row >= p.N) return;
      const uint8_t* c = q.c + q.b * c_step;
      q0 = ldg_stream_u4(c);
      q1 = ldg_stream_u4(c + c_second);
      am = __ldg(q.a + q.b * a_step);
    };

    Pos cur;
    cur.item = pair;
    enter_item(cur);
    cur.b = (pair < n_items ? item_b0(pair) : 0) + group;
    normalize(cur);
    prefetch(cur);
    ptx::griddep_wait();  // the adapter weights read in the adapter step are written by the optimizer's kernels
    // forward: this thread's lane of the TMEM weight ring (warp % 4 = lane quadrant the warp may access)
    const uint32_t a_tmem_lane = tmem_d + ((uint32_t)(quad * 32) << 16) + (uint32_t)kTmemACol0;
    int s = group;  // S >= kGroups: at most one ring wrap per step of kGroups
    uint32_t empty_parity = 1;
    for (int g = group; cur.item < n_items; g += kGroups) {
      const uint32_t a_tile = stage_a(s) + a_row_off;
      Pos nxt = cur;
      nxt.b += kGroups;
      normalize(nxt);
      uint32_t v[8][4];  // this thread's 64 decoded 16-bit values, in contraction order
      if (cur.b < n_main) {
        Nf4Lut lut;
        nf4_build_lut<ActT>(am, p.qdtype, lut);
        const uint32_t words[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        prefetch(nxt);  // next block's codes are in flight while this one is decoded
        // Decode into registers BEFORE waiting for the stage: the permute work then overlaps the MMAs still
        // reading the stage, and only the stores + a fence sit between "stage free" and "stage full".
#pragma unroll
        for (int c = 0; c < 8; ++c) nf4_decode_word(words[c], lut, v[c]);
      } else {
        // adapter step: forward row n of scale*B (r values), backward row j of A (64 in-features)
        const ActT* lw = static_cast<const ActT*>(p.lora_w);
        const int64_t f0 = cur.f0;
        prefetch(nxt);
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          v[c][0] = v[c][1] = v[c][2] = v[c][3] = 0u;
          if (kBackward) {
            const int64_t k = f0 + half * 64 + c * 8;
            if (row < p.r && k < p.K) {
              const uint4 q = __ldg(reinterpret_cast<const uint4*>(lw + (int64_t)row * p.K + k));
              v[c][0] = q.x; v[c][1] = q.y; v[c][2] = q.z; v[c][3] = q.w;
            }
          } else {
            const int64_t n = f0 + row;
            if (n < p.N && (p.r & 7) == 0 && (reinterpret_cast<uintptr_t>(lw) & 15u) == 0) {
              // ranks that are multiples of 8 (16 in every shipped config but one): one 16-byte load per 8 columns
              // instead of eight 2-byte loads in the path of the tile's last ring step
              if (c * 8 < p.r) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(lw + n * p.r + c * 8));
                const ActT* e8 = reinterpret_cast<const ActT*>(&q);
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  v[c][e] = pack2<ActT>(p.scale * to_f32<ActT>(e8[2 * e]), p.scale * to_f32<ActT>(e8[2 * e + 1]));
              }
            } else if (n < p.N) {
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = c * 8 + 2 * e;
                const float b0 = (j < p.r) ? p.scale * to_f32<ActT>(lw[n * p.r + j]) : 0.0f;
                const float b1 = (j + 1 < p.r) ? p.scale * to_f32<ActT>(lw[n * p.r + j + 1]) : 0.0f;
                v[c][e] = pack2<ActT>(b0, b1);
              }
            }
          }
        }
      }
      ptx::mbar_wait(bar_empty(s), empty_parity);
      if (dw == 0 && lane == 0) tl_mark(p, 2, g >> 2);
      if (!(p.debug & 1)) {
        if (kTmemA) {
          uint32_t flat[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            flat[4 * c] = v[c][0]; flat[4 * c + 1] = v[c][1]; flat[4 * c + 2] = v[c][2]; flat[4 * c + 3] = v[c][3];
          }
          ptx::tmem_st_32x32b_x32(a_tmem_lane + (uint32_t)(32 * s), flat);
        } else {
#pragma unroll
          for (int c = 0; c < 8; ++c)
            ptx::sts128(a_tile + (((uint32_t)c << 4) ^ a_xor), v[c][0], v[c][1], v[c][2], v[c][3]);
        }
      }
      if (dw == 0 && lane == 0) tl_mark(p, 6, g >> 2);
      if (kTmemA) {
        ptx::tmem_st_wait();     // the tile is in tensor memory ...
        ptx::tc_fence_before();  // ... before the arrive that lets the issuer's tcgen05.mma read it
      } else {
        ptx::fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core (async proxy)
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive_cluster(ptx::mapa(bar_full(s), 0));
      if (dw == 0 && lane == 0) tl_mark(p, 3, g >> 2);
      cur = nxt;
      s += kGroups;
      if (s >= S) {
        s -= S;
        empty_parity ^= 1u;
      }
    }
  }

  ptx::tc_fence_before();
  ptx::cluster_sync();  // the peer may still be reading this CTA's shared memory / arriving on its barriers
  if (warp == kAllocWarp) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_pair<kTmemCols>(tmem_d);
  }
}

// split contraction, second half: out[t, f] = ActT(bias[f] + sum over splits, in split order, of the fp32 slices);
// one thread per 4 consecutive features of one token row
template <typename ActT>
__global__ void __launch_bounds__(256)
qlora_tc2_finalize_kernel(const float* __restrict__ partial, const ActT* __restrict__ bias, int64_t T, int64_t OUT,
                          int tok_tile, int n_fblk, int n_split, ActT* __restrict__ out) {
  ptx::griddep_wait();  // the slices are written by the GEMM launch in front of this one
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t per_row = OUT >> 2;
  if (q >= T * per_row) return;
  const int64_t t = q / per_row, f = (q - t * per_row) << 2;
  const int64_t tile = (t / tok_tile) * n_fblk + (f >> 8);
  const float* src = partial + ((tile * n_split) * tok_tile + (t % tok_tile)) * (2 * kBM) + (f & 255);
  float4 acc = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (bias != nullptr) {
    acc.x = to_f32<ActT>(bias[f]); acc.y = to_f32<ActT>(bias[f + 1]); acc.z = to_f32<ActT>(bias[f + 2]); acc.w = to_f32<ActT>(bias[f + 3]);
  }
  // four slices in flight per thread (the loads come from other SMs' L2 lines; one at a time the loop measured
  // 12 k cycles for six slices), added in split order
  const int64_t slice_stride = (int64_t)tok_tile * (2 * kBM);
  int s = 0;
  for (; s + 4 <= n_split; s += 4) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = __ldcs(reinterpret_cast<const float4*>(src + (s + u) * slice_stride));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w;
    }
  }
  for (; s < n_split; ++s) {
    const float4 v = __ldcs(reinterpret_cast<const float4*>(src + s * slice_stride));
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  uint2 o;
  o.x = pack2<ActT>(acc.x, acc.y);
  o.y = pack2<ActT>(acc.z, acc.w);
  *reinterpret_cast<uint2*>(out + t * OUT + f) = o;
}

// ---------------------------------------------------------------------------
// host side: tile shape selection + launch
// ---------------------------------------------------------------------------
struct Tc2Config {
  int n_acc, N_acc, stages;
  double cost;
  int N_acc1 = 0;  // tokens of accumulator 1 when it is narrower than accumulator 0 (0: the same)
  int pingpong = 0;  // n_acc == 1: consecutive tiles of a pair alternate between the two accumulator pitches
  int tok() const { return n_acc == 2 ? N_acc + (N_acc1 > 0 ? N_acc1 : N_acc) : N_acc; }
};

// Cycle model per pipeline step (one CTA): tensor pipe 2*N_acc cycles per accumulator (M = 256 over the pair,
// K = 64), decode ~600 ALU-pipe cycles per 128 x 64 weight tile, shared memory 128 B/clk over the activation
// boxes (written by TMA, read by the MMA) and -- backward only -- the decoded tile (written once, read once per
// accumulator).
// `rp`: padded rank of a side product computed inside the launch (0: none) -- costs its ring (p0_budget() bytes of
// shared memory) and the last rp columns of an accumulator pitch.
static int p0_budget() { return (env().tc2_p0kb >= 10 && env().tc2_p0kb <= 64 ? env().tc2_p0kb : 40) * 1024; }
static int max_stages(bool tmem_a, int n_acc, int N_acc, int rp) {
  const int stage_bytes = (tmem_a ? 0 : kATileBytes) + n_acc * (N_acc / 2) * 128;
  int stages = (kSmemLimit - kBarBytes - kEpiBytes - 1024 - (rp > 0 ? p0_budget() : 0)) / stage_bytes;
  if (stages > kMaxStages) stages = kMaxStages;
  if (tmem_a && stages > kTmemAStages) stages = kTmemAStages;  // the weight ring in tensor memory has 4 slots
  // four stages (one per decode group) keep the tensor pipe fed; shared memory beyond that is worth more as output
  // staging (see n_stg), which shortens the accumulator drain at tile boundaries
  const int cap = env().tc2_max_stages >= kGroups ? env().tc2_max_stages : kGroups;
  if (stages > cap) stages = cap;
  return stages;
}

static bool config_ok(bool tmem_a, int n_acc, int N_acc, int rp) {
  if (n_acc < 1 || n_acc > kMaxAcc || N_acc < 16 || N_acc > 256 || N_acc % 16 != 0) return false;
  if (tmem_a && n_acc == 2 && N_acc > AccLayout<true>::pitch) return false;
  if (rp > 0) {  // side-product columns: the tail of the first pitch (two accumulators) / of the second (one)
    const int pitch = tmem_a ? AccLayout<true>::pitch : AccLayout<false>::pitch;
    if (n_acc == 2 ? N_acc > pitch - rp : N_acc > 2 * pitch - rp) return false;
  }
  return max_stages(tmem_a, n_acc, N_acc, rp) >= kGroups;  // a decode group may run at most one ring phase ahead
}

// Split-K plan: with fewer tiles than half the SM pairs, the contraction of every tile is divided over n_split work
// items so that about one item per pair exists (never more items than pairs: the items of a tile wait for each other
// inside the launch); their fp32 partial tiles meet in a caller-provided workspace, one slice per item.
struct Tc2Plan {
  Tc2Config cfg;
  int n_tiles, n_split, k_per;
  int64_t ws_bytes;  // 0 when no split is used
};

// Tile shape and split are chosen together from a cycle model of one CTA: per ring step the tensor pipe needs
// 2*N_acc cycles per accumulator (M = 256 over the pair, K = 64), decode ~620 ALU-pipe cycles per 128 x 64 weight
// tile, shared memory 128 B/clk over the activation boxes (TMA write + MMA read) and -- backward only -- the decoded
// tile (written once, read once per accumulator); per work item ~2500 cycles of fill + ~20 cycles per token of
// epilogue; a split item instead writes its fp32 partial tile (one 128-byte store per token and warp: ~50 cycles per
// token, measured) and a second small launch adds the slices (~5 k cycles of launch + the slices at ~1.5 KB per cycle
// out of L2), so the split pays for few tokens only (adaLN / modulation layers, text-token projections).
// cycles per wave that do not scale with the tile: pipeline fill + the part of the accumulator drain the next tile cannot
// hide.  Fitted on 9216 x 2304 at T = 4352 (forward): 7 waves of 2 x 160 tokens run 143.2 us, 6 waves of 2 x 192 run
// 135.5 us -- 5-7 k cycles per wave more than the 2500 of fill alone.
constexpr double kWaveFixed = 7000.0;
static Tc2Plan plan_tc2(int64_t T, int64_t OUT, int64_t RED, int r, bool tmem_a, int n_pairs, bool allow_split = true,
                        int rp = 0) {
  Tc2Plan best = {};
  double best_cost = 1e300;
  const int n_main = (int)ceil_div64(RED, kBK);
  const int64_t n_f = ceil_div64(OUT, 2 * kBM);
  auto consider = [&](int n_acc, int N_acc, int N_acc1, int pp = 0) {  // N_acc1 = 0: both accumulators N_acc wide
      const int b_bytes = (N_acc / 2) * 128;
      const int64_t tok = n_acc == 2 ? N_acc + (N_acc1 > 0 ? N_acc1 : N_acc) : N_acc;
      const int64_t tiles = n_f * ceil_div64(T, tok);
      const double mma = 2.0 * (double)tok;
      const double smem = ((tmem_a ? 0.0 : kATileBytes * (1.0 + n_acc)) + 2.0 * n_acc * b_bytes) / 128.0;
      double step = mma > smem ? mma : smem;
      if (step < 620.0) step = 620.0;
      int split = 1, k_per = n_main;
      if (allow_split && tiles * 2 <= n_pairs && n_main >= 4) {
        int want = (int)(n_pairs / tiles);
        if (want > n_main / 2) want = n_main / 2;  // at least two ring steps per item
        if (want > 1) {
          k_per = (n_main + want - 1) / want;
          split = (n_main + k_per - 1) / k_per;
          if (split <= 1) { split = 1; k_per = n_main; }
        }
      }
      const int64_t ws = tiles * split * tok * (2 * kBM) * (int64_t)sizeof(float);  // one slice per work item
      if (split > 1 && (ws > (64ll << 20) || tiles * split > n_pairs)) { split = 1; k_per = n_main; }
      const double waves = (double)ceil_div64(tiles * split, n_pairs);
      const double tok_live = (double)(tok < T ? tok : T);
      const double epi = split > 1 ? 50.0 * tok_live + 5000.0 + (double)(split * T * OUT) * 4.0 / 1500.0 : 20.0 * tok;
      // (the adapter's single ring step is left out on purpose: the plan -- above all whether and where the contraction
      //  is split, i.e. the order of the fp32 sums -- must not depend on r, so that a layer whose lora_up is still zero
      //  returns exactly what the bare base layer returns: /root/reference/tests/test_peft.py:98-101)
      (void)r;
      double cost = waves * (k_per * step + kWaveFixed + epi);
      if (pp) {
        // ping-pong: a tile's drain and the next tile's fill run under the contraction; what is left per tile is the
        // commit / barrier round trip, and one fill + one drain per launch.  Needs several tiles per pair and no split.
        if (split > 1 || waves < 2.0) return;
        cost = waves * (k_per * step + 1500.0) + kWaveFixed + epi;
      }
      if (cost < best_cost * 0.999 || (cost < best_cost * 1.001 && tok > best.cfg.tok())) {
        best_cost = cost;
        best.cfg = {n_acc, N_acc, max_stages(tmem_a, n_acc, N_acc, rp), cost, N_acc1, pp};
        best.n_tiles = (int)tiles;
        best.n_split = split;
        best.k_per = k_per;
        best.ws_bytes = split > 1 ? ws : 0;
      }
  };
  for (int n_acc = 1; n_acc <= kMaxAcc; ++n_acc)
    for (int N_acc = 32; N_acc <= 256; N_acc += 16)
      if (config_ok(tmem_a, n_acc, N_acc, rp)) consider(n_acc, N_acc, 0);
  // one accumulator per tile, alternating between the two pitches (same column budget per pitch as two accumulators).
  // OFF unless VFT_TC2_PINGPONG=1: measured on the short-K SDXL layers it was meant for, the planner's choices with
  // it ran SLOWER than two accumulators sharing every decoded tile (10240 x 1280 at T = 2048 forward 66.7 vs 54.8 us,
  // 5120 x 640 at T = 8192 75.3 vs 63.9 us, 640 x 2560 backward 41.2 vs 32.8 us): halving the tokens per decoded tile
  // costs more decode than the hidden drain wins back.  Kept as a tested schedule (tests force it), not as a default.
  if (env().tc2_pingpong == 1)
    for (int N_acc = 32; N_acc <= 256; N_acc += 16)
      if (config_ok(tmem_a, 2, N_acc, rp)) consider(1, N_acc, 0, 1);
  // forward with its side product inside: accumulator 0 keeps the full pitch, accumulator 1 gives up the rp columns
  // (2 x 176 tokens per tile is 300 tiles = 5 waves at T = 8720, 3072 features; 192 + 176 is 288 = 4 waves)
  if (tmem_a && rp > 0 && !allow_split) {
    const int pitch = AccLayout<true>::pitch;
    if (max_stages(tmem_a, 2, pitch, rp) >= kGroups) consider(2, pitch, pitch - rp);
  }
  return best;
}

static int device_pairs() {
  // (asked several times per call: one attribute query per device and process)
  static std::atomic<int> cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) {
    cudaGetLastError();
    return 74;
  }
  if (dev >= 0 && dev < 64) {
    const int c = cached[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int n_sm = 148;
  if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) cudaGetLastError();
  const int pairs = n_sm / 2 > 0 ? n_sm / 2 : 1;
  if (dev >= 0 && dev < 64) cached[dev].store(pairs, std::memory_order_relaxed);
  return pairs;
}

// Padded rank of the side product (forward: t = x . A^T, backward: dt = s * dy . B) that the launch computes itself,
// 0 if it cannot: the rank-r operand must be a K-major [r_pad, contraction] matrix that TMA can address -- forward:
// lora_down.weight as it is; backward: `bt` = s * lora_up.weight^T, written by the forward call -- and the rank has to
// fit the spare accumulator columns.  VFT_TC2_FUSE=0 keeps the side kernels of lora_tc.cu.
static int side_rank(const LayerArgs& a, bool backward) {
  if (a.r <= 0 || a.r > kMaxSideRP || env().tc2_fuse == 0) return 0;
  const void* w = backward ? a.bt_save : a.lora_a;
  if (w == nullptr || (reinterpret_cast<uintptr_t>(w) & 15u) != 0) return 0;
  return a.r <= 16 ? 16 : 32;
}

struct Tc2Choice {
  Tc2Plan plan;
  bool job;     // backward: dA / dB inside the launch too
  int rp;       // > 0: side product inside the launch
  int pairs;    // CTA pairs launched
  int p0_rows;  // token rows per CTA of the side product
};

// the fewest pairs that still finish in the same number of waves (T = 4096, 3072 features: 144 tiles -> 72 pairs
// of 2 tiles instead of 74): identical run time, and the SMs left over stay free for a concurrent NCCL all-reduce
// of the LoRA gradients, which otherwise delays the launch of the last cluster until it has drained
static int pairs_for(const Tc2Plan& plan, int n_pairs) {
  const int n_items = plan.n_tiles * plan.n_split;
  const int waves = (n_items + n_pairs - 1) / n_pairs;
  return (n_items + waves - 1) / waves;
}

// Everything the launch is decided by, in one place (the API asks the same function whether the side kernel can be
// skipped): tile shape, split, side product, grid.
static Tc2Choice choose_tc2_uncached(const LayerArgs& a, bool backward, int n_pairs);

// One C-ABI call asks up to three times (does the launch compute the side product itself? dA/dB too? then the launch):
// the last answer of this thread is kept, keyed on everything the decision reads.
static Tc2Choice choose_tc2(const LayerArgs& a, bool backward, int n_pairs) {
  struct Key {
    int64_t T, N, K, ws_bytes;
    const void *lora_a, *bt_save, *ws, *job_x, *tt_save, *job_dtt, *job_da, *job_db;
    int r, backward, n_pairs;
    unsigned env_gen;
  };
  static thread_local Key last_key = {};
  static thread_local Tc2Choice last_choice = {};
  static thread_local bool have = false;
  Key k = {};  // (zero-initialised including padding: compared with memcmp)
  k.T = a.T; k.N = a.N; k.K = a.K; k.ws_bytes = a.ws_bytes;
  k.lora_a = a.lora_a; k.bt_save = a.bt_save; k.ws = a.ws; k.job_x = a.job_x; k.tt_save = a.tt_save;
  k.job_dtt = a.job_dtt; k.job_da = a.job_da; k.job_db = a.job_db;
  k.r = a.r; k.backward = backward ? 1 : 0; k.n_pairs = n_pairs; k.env_gen = env_generation();
  if (have && memcmp(&k, &last_key, sizeof(Key)) == 0) return last_choice;
  last_choice = choose_tc2_uncached(a, backward, n_pairs);
  last_key = k;
  have = true;
  return last_choice;
}

static Tc2Choice choose_tc2_uncached(const LayerArgs& a, bool backward, int n_pairs) {
  const int64_t OUT = backward ? a.K : a.N;
  const int64_t RED = backward ? a.N : a.K;
  const bool tmem_a = !backward;
  const VftEnv& ev = env();
  Tc2Choice c;
  c.p0_rows = 0;
  c.job = false;
  // triage override VFT_TC2_NACC="<n_acc>x<N_acc>" (clamped to what the path allows): forces the shape, never splits
  auto forced = [&](int rp, Tc2Plan& plan) -> bool {
    if (ev.tc2_force_na <= 0) return false;
    int na = ev.tc2_force_na, nn = ev.tc2_force_nn;
    if (tmem_a && na == 2 && nn > AccLayout<true>::pitch) nn = AccLayout<true>::pitch;
    int nn1 = 0;
    if (tmem_a && rp > 0 && na == 2 && nn == AccLayout<true>::pitch && max_stages(tmem_a, na, nn, rp) >= kGroups) {
      nn1 = nn - rp;  // "2x192" with a side product inside: accumulator 1 gives up the columns
    } else if (!config_ok(tmem_a, na, nn, rp)) {
      return false;
    }
    plan.cfg = {na, nn, max_stages(tmem_a, na, nn, rp), plan.cfg.cost, nn1};
    plan.cfg.pingpong = (na == 1 && ev.tc2_pingpong == 1 && config_ok(tmem_a, 2, nn, rp)) ? 1 : 0;  // triage: force it
    plan.n_tiles = (int)(ceil_div64(OUT, 2 * kBM) * ceil_div64(a.T, (int64_t)plan.cfg.tok()));
    plan.n_split = 1;
    plan.k_per = (int)ceil_div64(RED, kBK);
    plan.ws_bytes = 0;
    return true;
  };
  c.rp = side_rank(a, backward);
  const bool ws_ok = a.ws != nullptr && (reinterpret_cast<uintptr_t>(a.ws) & 15u) == 0 && !ev.tc2_nosplit;
  if (c.rp > 0) {
    // Inside the launch or next to it?  Compared on the cycle model, with the side kernels of lora_tc.cu at what they
    // measured (3.5 us + bytes at 3.2 TB/s for t / dt, 4 us + bytes at 3.9 TB/s for dA/dB: 7.9 / 8.5 / 12.9 us at
    // config #1) and the fused launch at 1.08 x its plan (the side product shares the tensor pipe, the TMA queue and
    // 40 KB of the staging tiles' shared memory with the main loop).  What decides in practice is the wave count:
    // setting the side product's columns aside caps the forward's accumulators at 2 x 176 tokens instead of 2 x 192,
    // and at T = 8720 (AuraFlow, batch 2) that is 300 tiles = 5 waves instead of 276 = 4 (measured, 3072 x 3072:
    // 139.5 us fused against 131.8 us with the side kernel; at T = 4096 and 8192 the fused launch wins).
    const double cyc_per_us = 1900.0;
    const double side_cyc = cyc_per_us * (3.5 + (double)a.T * (double)RED * 2.0 / 3.2e6);
    const double dab_cyc = backward ? cyc_per_us * (4.0 + (double)a.T * (double)(a.N + a.K) * 2.0 / 3.9e6) : 0.0;
    const Tc2Plan base = plan_tc2(a.T, OUT, RED, a.r, tmem_a, n_pairs);  // (whether or not a workspace was passed)
    const double base_cost = base.cfg.cost + side_cyc + dab_cyc;
    // dA/dB job: everything it reads and writes is there, TMA can address the transposed side products (row stride
    // T * 2 bytes), and a second block of r_pad accumulator columns is free
    const bool job_ok = backward && ev.tc2_job != 0 && a.job_x && a.tt_save && a.job_dtt && a.job_da && a.job_db &&
                        a.T % 8 == 0 && ((reinterpret_cast<uintptr_t>(a.job_x) | reinterpret_cast<uintptr_t>(a.tt_save) |
                                          reinterpret_cast<uintptr_t>(a.job_dtt)) & 15u) == 0;
    Tc2Choice best_c = c;
    double best_cost = 1e300;
    // candidates: the job with one unit per pair, with two (more column tiles than pairs: N + K > 128 * pairs, the
    // 8192-wide MLP layers -- each unit has its own block of accumulator columns), and no job at all
    for (int upp = job_ok ? kMaxJobUnits : 0; upp >= 0; --upp) {
      Tc2Choice k = c;
      const int cols = c.rp * (1 + upp);  // accumulator columns set aside
      k.plan = plan_tc2(a.T, OUT, RED, a.r, tmem_a, n_pairs, /*allow_split=*/false, cols);
      forced(cols, k.plan);
      if (k.plan.cfg.N_acc <= 0) continue;
      k.pairs = pairs_for(k.plan, n_pairs);
      const int64_t rows = ceil_div64(ceil_div64(a.T, 2 * k.pairs), 8) * 8;
      if (rows > kBM) continue;  // more tokens than one pass of 128 rows per CTA holds
      const int64_t units = ceil_div64(a.K, kBM) + ceil_div64(a.N, kBM);
      // the job streams all T tokens through ONE pair per 128 columns at ~900 cycles per 64 tokens: it has to end
      // before the pair's last tile does (C640 at T = 8192: 115 k cycles of job against 21 k of GEMM -- 50 us, measured)
      const double job_cyc = 900.0 * (double)ceil_div64(a.T, 64) * upp;
      if (upp > 0 && (units > (int64_t)upp * k.pairs || (upp > 1 && units <= k.pairs) || job_cyc > 0.85 * k.plan.cfg.cost))
        continue;
      k.p0_rows = (int)rows;
      k.job = upp != 0;
      const double cost = 1.08 * k.plan.cfg.cost + (upp ? 0.0 : dab_cyc);
      if (cost < best_cost) {
        best_cost = cost;
        best_c = k;
      }
    }
    // a problem small enough to be split along the contraction keeps its (then tiny) side kernels: with and without
    // an adapter it must add its fp32 partial sums in the same order (see plan_tc2)
    if (best_cost < 1e300 && (ev.tc2_fuse == 1 || (best_cost <= base_cost && base.n_split == 1))) return best_c;
    c.rp = 0;
  }
  c.plan = plan_tc2(a.T, OUT, RED, a.r, tmem_a, n_pairs);
  if (c.plan.n_split > 1 && (!ws_ok || a.ws_bytes < c.plan.ws_bytes))
    c.plan = plan_tc2(a.T, OUT, RED, a.r, tmem_a, n_pairs, /*allow_split=*/false);  // no workspace: unsplit shape
  forced(0, c.plan);
  c.pairs = pairs_for(c.plan, n_pairs);
  return c;
}

// {arrivals, generation} pairs of the grid-wide counters of the side product, handed out round-robin per launch: zero at
// module load, every launch leaves its pair at arrivals == 0.  Two launches share a pair only if kSyncSlots launches
// were issued in between, i.e. the earlier one has long finished (same-stream launches run in order); the one thing
// to avoid is replaying ONE captured graph concurrently with itself on two streams.
constexpr int kSyncSlots = 4096;
__device__ unsigned g_tc2_sync[2 * kSyncSlots];
static unsigned* next_sync_pair(unsigned n = 1) {  // n consecutive pairs
  static std::atomic<unsigned> next{0};
  static unsigned* base = nullptr;  // (the symbol's address does not change; a failed lookup is retried)
  if (base == nullptr && cudaGetSymbolAddress(reinterpret_cast<void**>(&base), g_tc2_sync) != cudaSuccess) {
    base = nullptr;
    return nullptr;
  }
  unsigned first = next.fetch_add(n, std::memory_order_relaxed) % kSyncSlots;
  if (first + n > kSyncSlots) first = next.fetch_add(n, std::memory_order_relaxed) % kSyncSlots;  // no wrap inside a run
  if (first + n > kSyncSlots) first = 0;
  return base + 2 * first;
}

template <typename ActT, bool kBackward>
static int launch_tc2(const LayerArgs& a, const void* act, void* out, void* lora_act, cudaStream_t st) {
  const int64_t OUT = kBackward ? a.K : a.N;
  const int64_t RED = kBackward ? a.N : a.K;
  const int n_pairs = device_pairs();
  constexpr bool kTmemA = !kBackward;
  const VftEnv& ev = env();
  const Tc2Choice choice = choose_tc2(a, kBackward, n_pairs);
  const Tc2Plan& plan = choice.plan;
  const int rp = choice.rp;
  Tc2Config cfg = plan.cfg;
  if (ev.tc2_stages >= kGroups && ev.tc2_stages <= cfg.stages) cfg.stages = ev.tc2_stages;

  Tc2Params p;
  p.T = a.T; p.N = a.N; p.K = a.K; p.r = a.r; p.qdtype = a.qdtype; p.scale = a.scale;
  p.tiled = a.codes_t != nullptr;
  p.packed = p.tiled ? a.codes_t : a.packed;
  p.absmax = p.tiled ? a.absmax_t : a.absmax;
  p.bias = kBackward ? nullptr : a.bias;
  p.lora_w = kBackward ? a.lora_a : a.lora_b;
  p.out = out;
  p.n_acc = cfg.n_acc;
  p.N_acc = cfg.N_acc;
  p.N_acc1 = cfg.N_acc1 > 0 ? cfg.N_acc1 : cfg.N_acc;
  p.pingpong = cfg.n_acc == 1 ? cfg.pingpong : 0;
  p.stages = cfg.stages;
  p.side = rp > 0 ? 1 : 0;
  p.r_pad = rp > 0 ? rp : 16;
  p.la_bytes = (p.r_pad / 2) * 128;
  p.p0_rows = rp > 0 ? choice.p0_rows : 8;
  p.p0_slot_bytes = p.p0_rows * 128 + p.la_bytes;
  p.p0_slots = rp > 0 ? p0_budget() / p.p0_slot_bytes : 0;
  if (p.p0_slots > kMaxP0) p.p0_slots = kMaxP0;
  p.p0_per_step = 2;
  p.save = lora_act;
  p.bt_out = kBackward ? nullptr : a.bt_save;
  p.save_t = rp > 0 ? (kBackward ? (choice.job ? a.job_dtt : nullptr) : a.tt_save) : nullptr;
  p.job = choice.job ? 1 : 0;
  p.job_units_a = (int)ceil_div64(a.K, kBM);
  p.job_units = choice.job ? p.job_units_a + (int)ceil_div64(a.N, kBM) : 0;
  p.job_da = a.job_da;
  p.job_db = a.job_db;
  if (choice.job && p.p0_slot_bytes < kJobBox + p.la_bytes) {  // the job's boxes ride the side product's ring
    p.p0_slot_bytes = kJobBox + p.la_bytes;
    p.p0_slots = p0_budget() / p.p0_slot_bytes;
    if (p.p0_slots > kMaxP0) p.p0_slots = kMaxP0;
  }
  p.sync = nullptr;
  if (rp > 0) {
    p.sync = next_sync_pair();
    if (p.sync == nullptr) {
      set_error("cudaGetSymbolAddress(g_tc2_sync) failed: %s", cudaGetErrorString(cudaGetLastError()));
      return VFT_ERR_CUDA;
    }
  }
  p.b_bytes = (cfg.N_acc / 2) * 128;
  p.stage_bytes = (kTmemA ? 0 : kATileBytes) + cfg.n_acc * p.b_bytes;
  p.n_fblk = (int)ceil_div64(OUT, 2 * kBM);
  p.n_tiles = p.n_fblk * (int)ceil_div64(a.T, (int64_t)cfg.tok());
  p.n_split = 1;
  p.k_per = (int)ceil_div64(RED, kBK);
  p.partial = nullptr;
  if (plan.n_split > 1) {
    p.n_split = plan.n_split;
    p.k_per = plan.k_per;
    p.partial = static_cast<float*>(a.ws);
  }
  p.debug = ev.tc_debug;

  const CUtensorMapDataType dt =
      std::is_same<ActT, __nv_bfloat16>::value ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16;
  CUtensorMap map_act, map_lora, map_out, map_out16, map_p0a, map_p0w;
  JobMaps jm;
  int rc = make_map_2d(&map_act, dt, act, (uint64_t)RED, (uint64_t)a.T, (uint64_t)RED * 2, kBK, cfg.N_acc / 2,
                       CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  // output rows leave through TMA stores of [32 (or 16) tokens x 64 features] boxes, clipped at T / OUT
  rc = make_map_2d(&map_out, dt, out, (uint64_t)OUT, (uint64_t)a.T, (uint64_t)OUT * 2, 64, 32, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  rc = make_map_2d(&map_out16, dt, out, (uint64_t)OUT, (uint64_t)a.T, (uint64_t)OUT * 2, 64, 16, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc != VFT_OK) return rc;
  map_p0a = map_p0w = map_act;
  if (rp > 0) {
    // side product: [p0_rows x 64] boxes of the activations; the rank-r operand as a K-major [rows, contraction] matrix
    // in boxes of [rp/2 rows x 64] (forward: lora_down.weight [r, K], rows >= r zero-filled; backward: bt [rp, N])
    rc = make_map_2d(&map_p0a, dt, act, (uint64_t)RED, (uint64_t)a.T, (uint64_t)RED * 2, kBK, (uint32_t)p.p0_rows,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
    rc = make_map_2d(&map_p0w, dt, kBackward ? a.bt_save : a.lora_a, (uint64_t)RED, (uint64_t)(kBackward ? rp : a.r),
                     (uint64_t)RED * 2, kBK, (uint32_t)(rp / 2), CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
  }
  jm.m_a = jm.m_b = jm.v_a = jm.v_b = map_act;
  if (choice.job) {
    rc = make_map_2d(&jm.m_a, dt, a.job_x, (uint64_t)a.K, (uint64_t)a.T, (uint64_t)a.K * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
    rc = make_map_2d(&jm.m_b, dt, act, (uint64_t)a.N, (uint64_t)a.T, (uint64_t)a.N * 2, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
    rc = make_map_2d(&jm.v_a, dt, a.job_dtt, (uint64_t)a.T, (uint64_t)rp, (uint64_t)a.T * 2, 64, (uint32_t)(rp / 2),
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
    rc = make_map_2d(&jm.v_b, dt, a.tt_save, (uint64_t)a.T, (uint64_t)rp, (uint64_t)a.T * 2, 64, (uint32_t)(rp / 2),
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
  }
  if (a.r > 0) {
    rc = make_map_2d(&map_lora, dt, lora_act, VFT_LORA_LD, (uint64_t)a.T, VFT_LORA_LD * 2, kBK, cfg.N_acc / 2,
                     CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc != VFT_OK) return rc;
  } else {
    map_lora = map_act;
  }

  // output staging tiles from what the ring leaves over (a whole tile of 2 x 176 tokens is 11 of them)
  const int p0_bytes = p.p0_slots * p.p0_slot_bytes;
  p.n_stg = (kSmemLimit - kBarBytes - 1024 - cfg.stages * p.stage_bytes - p0_bytes) / kStgBytes;
  if (p.n_stg > kMaxStg) p.n_stg = kMaxStg;
  if (ev.tc2_n_stg >= 2 && ev.tc2_n_stg <= p.n_stg) p.n_stg = ev.tc2_n_stg;
  const int dyn_bytes = cfg.stages * p.stage_bytes + p0_bytes + p.n_stg * kStgBytes + kBarBytes + 1024;  // + 1024-B alignment slack
  auto kern_plain = qlora_tc2_kernel<ActT, kBackward, false, false>;
  auto kern_side = qlora_tc2_kernel<ActT, kBackward, true, false>;
  auto kern_job = qlora_tc2_kernel<ActT, kBackward, true, kBackward>;  // (forward: the same function as kern_side)
  VFT_OPT_IN_SMEM_ONCE(kern_plain, VFT_MAX_DYN_SMEM);  // dyn_bytes depends on the plan: opt in to the limit
  VFT_OPT_IN_SMEM_ONCE(kern_side, VFT_MAX_DYN_SMEM);
  VFT_OPT_IN_SMEM_ONCE(kern_job, VFT_MAX_DYN_SMEM);
  auto kern = rp > 0 ? (choice.job ? kern_job : kern_side) : kern_plain;
  const int pairs = choice.pairs;
  cudaLaunchConfig_t lc = {};
  lc.gridDim = dim3((unsigned)(2 * pairs));
  lc.blockDim = dim3(kThreads);
  lc.dynamicSmemBytes = (size_t)dyn_bytes;
  lc.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  lc.attrs = attr;
  lc.numAttrs = pdl_enabled() ? 2 : 1;
  VFT_CUDA_OK(cudaLaunchKernelEx(&lc, kern, map_act, map_lora, map_out, map_out16, map_p0a, map_p0w, jm, p));
  VFT_CUDA_OK(cudaGetLastError());
  if (p.n_split > 1) {
    const int64_t quads = a.T * (OUT / 4);
    cudaLaunchConfig_t fc = {};
    fc.gridDim = dim3((unsigned)ceil_div64(quads, 256));
    fc.blockDim = dim3(256);
    fc.stream = st;
    fc.attrs = attr + 1;  // programmatic stream serialization only: its launch latency hides under the GEMM's tail
    fc.numAttrs = pdl_enabled() ? 1 : 0;
    VFT_CUDA_OK(cudaLaunchKernelEx(&fc, qlora_tc2_finalize_kernel<ActT>, static_cast<const float*>(p.partial),
                                   static_cast<const ActT*>(p.bias), a.T, OUT, cfg.tok(), p.n_fblk, p.n_split,
                                   static_cast<ActT*>(out)));
    VFT_CUDA_OK(cudaGetLastError());
  }
  return VFT_OK;
}

}  // namespace
}  // namespace vft

extern "C" int vft_debug_tc2_p0dump(float* out, int n) {
  if (n > 2 * 128 * 32) n = 2 * 128 * 32;
  return cudaMemcpyFromSymbol(out, vft::g_tc2_p0dump, sizeof(float) * n) == cudaSuccess ? 0 : -3;
}

// Debug export (not part of the public ABI): timeline of the last VFT_TC_DEBUG&16 launch of the pair kernel.
extern "C" int vft_debug_tc2_timeline(unsigned long long* out, int n) {
  if (n > vft::kTlRows * vft::kTlCols) n = vft::kTlRows * vft::kTlCols;
  return cudaMemcpyFromSymbol(out, vft::g_tc2_timeline, sizeof(unsigned long long) * n) == cudaSuccess ? 0 : -3;
}

namespace vft {

// The persistent pair kernel pays off once there are enough tokens to give every SM pair work and to make the
// GEMM tensor-bound; below that the one-tile-per-CTA kernel of qlora_tc.cu (more, smaller CTAs) is used.
bool tc2_preferred(const LayerArgs& a, bool backward) {
  const int64_t OUT = backward ? a.K : a.N;
  if (OUT % 8 != 0) return false;  // row stride of the output must be a multiple of 16 bytes (TMA store)
  if (env().tc2 >= 0) return env().tc2 != 0 && OUT >= kBM + 1 && a.T >= 1;
  return OUT >= 2 * kBM;  // every token count: small problems are split along the contraction
}

// fp32 workspace the split-K form wants for this call (0: none); the call still works without it, unsplit
int64_t tc2_workspace_bytes(int64_t T, int64_t N, int64_t K, int r, bool backward) {
  const int64_t OUT = backward ? K : N, RED = backward ? N : K;
  if (T <= 0 || OUT % 8 != 0 || OUT < 2 * kBM || K % 64 != 0) return 0;
  return plan_tc2(T, OUT, RED, r, !backward, device_pairs()).ws_bytes;  // (a fused plan never splits)
}

// True when the launch computes its adapter side product (t_save / dt_save) itself: the caller then skips the side kernel.
bool tc2_fuses_side(const LayerArgs& a, bool backward) {
  if (a.T <= 0 || !tc2_preferred(a, backward)) return false;
  return choose_tc2(a, backward, device_pairs()).rp > 0;
}

// True when the backward launch also computes the adapter's weight gradients (no vft_lora_bwd_dab afterwards)
bool tc2_fuses_dab(const LayerArgs& a) {
  if (a.T <= 0 || !tc2_preferred(a, true)) return false;
  return choose_tc2(a, true, device_pairs()).job;
}

int tc2_fwd(const LayerArgs& a, const void* x, void* y, void* t_save, cudaStream_t st) {
  if (a.act_dtype == VFT_BF16) return launch_tc2<__nv_bfloat16, false>(a, x, y, t_save, st);
  return launch_tc2<__half, false>(a, x, y, t_save, st);
}

int tc2_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st) {
  if (a.act_dtype == VFT_BF16) return launch_tc2<__nv_bfloat16, true>(a, dy, dx, const_cast<void*>(dt_save), st);
  return launch_tc2<__half, true>(a, dy, dx, const_cast<void*>(dt_save), st);
}

}  // namespace vft
