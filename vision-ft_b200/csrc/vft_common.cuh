// Shared helpers for the vft_b200 CUDA sources (sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/vft_b200.h"

namespace vft {

// ---------------------------------------------------------------------------
// error plumbing (thread-local message, read through vft_last_error()).
// ---------------------------------------------------------------------------
void set_error(const char* fmt, ...);
void set_path(int path);
int forced_path();

#define VFT_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      ::vft::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return VFT_ERR_CUDA;                                                                  \
    }                                                                                       \
  } while (0)

#define VFT_REQUIRE(cond, ...)          \
  do {                                  \
    if (!(cond)) {                      \
      ::vft::set_error(__VA_ARGS__);    \
      return VFT_ERR_INVALID;           \
    }                                   \
  } while (0)

// ---------------------------------------------------------------------------
// NF4 constants.  Code book = bitsandbytes' NF4 `quant_map`; thresholds = the 15
// literals of its dQuantizeNF4 decision tree.
// ---------------------------------------------------------------------------
#define VFT_NF4_CODEBOOK                                                                       \
  {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,                  \
   -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,                 \
   0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,    \
   0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f}

#define VFT_NF4_THRESHOLDS                                                                     \
  {-0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f,   \
   -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f, \
   0.1202552504837513f, 0.2035212516784668f, 0.2920137718319893f, 0.3893125355243683f,        \
   0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f}

__device__ __forceinline__ float nf4_code_value(unsigned c) {
  // Compile-time table folded into a select chain / constant bank by nvcc.
  constexpr float kCode[16] = VFT_NF4_CODEBOOK;
  return kCode[c & 15u];
}

// ---------------------------------------------------------------------------
// dtype helpers
// ---------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f32(T v);
template <>
__device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }
template <>
__device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__device__ __forceinline__ T from_f32(float v);
template <>
__device__ __forceinline__ float from_f32<float>(float v) { return v; }
template <>
__device__ __forceinline__ __half from_f32<__half>(float v) { return __float2half_rn(v); }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f32<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Rounding W~ the way bitsandbytes does: fp32 product -> quant_state.dtype -> activation dtype.
template <typename ActT>
__device__ __forceinline__ float round_through(float v, int qdtype) {
  if (qdtype == VFT_F16) v = __half2float(__float2half_rn(v));
  else if (qdtype == VFT_BF16) v = __bfloat162float(__float2bfloat16_rn(v));
  return to_f32<ActT>(from_f32<ActT>(v));
}

inline int64_t ceil_div64(int64_t a, int64_t b) { return (a + b - 1) / b; }

// Opt a kernel instantiation in to large dynamic shared memory ONCE per device instead of on every launch (the call is
// ~1 us of host time, and host time per layer call is what bounds small-batch steps).  `bytes` is the most the
// instantiation will ever ask for; each expansion site owns its flag word (one bit per device ordinal).
#define VFT_MAX_DYN_SMEM 232448 /* 227 KB: the per-block opt-in limit of sm_100 */
#define VFT_OPT_IN_SMEM_ONCE(kern, bytes)                                                                  \
  do {                                                                                                     \
    static unsigned long long _vft_done = 0;                                                               \
    int _vft_dev = 0;                                                                                      \
    cudaGetDevice(&_vft_dev);                                                                              \
    const unsigned long long _vft_bit = 1ull << (_vft_dev & 63);                                           \
    if (!(__atomic_load_n(&_vft_done, __ATOMIC_RELAXED) & _vft_bit)) {                                     \
      VFT_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes)));  \
      __atomic_fetch_or(&_vft_done, _vft_bit, __ATOMIC_RELAXED);                                           \
    }                                                                                                      \
  } while (0)

// Triage switches.  The environment is read ONCE per process (getenv on every launch is host time on the hot path);
// vft_reload_env() re-reads it (tests that flip a switch between calls; not thread-safe).
struct VftEnv {
  bool pdl = true;         // VFT_PDL=0: no programmatic dependent launch
  bool side_mma = false;   // VFT_SIDE_MMA=1: mma.sync adapter kernels (lora_mma.cu) instead of lora_tc.cu
  int side_split = 0;      // VFT_SIDE_SPLIT: cluster size of the adapter side kernels
  int gemv_tma = 1;        // VFT_GEMV_TMA=0: register-fed few-token kernel
  bool tc_pair = false;    // VFT_TC_PAIR=1: CTA-pair form of the one-tile-per-CTA kernel
  int tc_debug = 0;        // VFT_TC_DEBUG: bit mask, see qlora_tc2.cu (results are garbage when non-zero)
  int tc2 = -1;            // VFT_TC2: -1 unset, 0 persistent pair kernel off, 1 on for every shape it can take
  int tc2_max_stages = 0, tc2_stages = 0, tc2_n_stg = 0;   // VFT_TC2_MAXSTAGES / _STAGES / _NSTG
  int tc2_force_na = 0, tc2_force_nn = 0;                   // VFT_TC2_NACC="<n_acc>x<N_acc>"
  bool tc2_nosplit = false;                                 // VFT_TC2_NOSPLIT=1
  int tc2_p0kb = 0;   // VFT_TC2_P0KB: shared memory (KB) of the side product's ring (default 40)
  int tc2_pingpong = -1;  // VFT_TC2_PINGPONG=1: let the planner alternate accumulators between the tiles of a pair (off by default: measured slower)
  int tc2_job = -1;   // VFT_TC2_JOB=0: adapter weight gradients by vft_lora_bwd_dab's kernel, not inside the backward launch
  int tc2_fuse = -1;  // VFT_TC2_FUSE=1: adapter down-projection fused into the forward launch whenever the shape
                      // allows it (never splits the contraction); otherwise a side kernel (see fuse_rank())
  void load();
};
const VftEnv& env();
void reload_env();
unsigned env_generation();  // bumped by reload_env()
inline bool pdl_enabled() { return env().pdl; }

// ---------------------------------------------------------------------------
// launchers implemented across the .cu files (all return vft_status)
// ---------------------------------------------------------------------------
int launch_quantize_many(int count, const void* const* w, int dtype, const int64_t* n, int blocksize,
                         uint8_t* const* packed, float* const* absmax, cudaStream_t st);
int launch_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax,
                    cudaStream_t st);
int launch_dequantize(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, void* out, int dtype,
                      cudaStream_t st);
int launch_tile_weight(const uint8_t* packed, const float* absmax, int64_t N, int64_t K, int blocksize,
                       uint8_t* codes_t, float* absmax_t, cudaStream_t st);

// nested block statistics (absmax_nest.cu)
int64_t absmax_nest_workspace_bytes();
int launch_absmax_nest(const float* absmax, int64_t n, int blocksize2, const float* code256, uint8_t* absmax8,
                       float* absmax2, float* offset_out, void* ws, int64_t ws_bytes, cudaStream_t st,
                       const float* offset_in = nullptr);
int launch_absmax_denest(const uint8_t* absmax8, const float* absmax2, const float* code256, float offset, int64_t n,
                         int blocksize2, float* out, cudaStream_t st);

struct LayerArgs {
  int64_t T, N, K;
  int blocksize, act_dtype, qdtype, r;
  float scale;
  const uint8_t* packed;
  const float* absmax;
  const void* bias;
  const void* lora_a;
  const void* lora_b;
  const uint8_t* codes_t = nullptr;  // optional micro-tiled copy (vft_nf4_tile_weight); both or neither
  const float* absmax_t = nullptr;
  void* ws = nullptr;  // optional workspace (vft_workspace_bytes) for the split-K form of small problems
  int64_t ws_bytes = 0;
  // s * lora_up.weight^T as a K-major [16 * ceil(r / 16), N] matrix: written by the forward call (when asked for), read
  // by the backward call, which can then compute dt = s * dy . B inside its launch; nullptr = not available
  void* bt_save = nullptr;
  // t^T [16 * ceil(r / 16), T]: forward output (optional), backward input -- with it, x, a dt^T scratch of the same
  // size and the two gradient buffers, the backward launch also computes dA = dt^T . x and dB = s * dy^T . t itself
  void* tt_save = nullptr;
  const void* job_x = nullptr;
  void* job_dtt = nullptr;
  void* job_da = nullptr;
  void* job_db = nullptr;
};

// generic CUDA-core family
int simt_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                   cudaStream_t st);
int simt_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
                 cudaStream_t st);
int simt_lora_bt(const void* b, int64_t N, int r, float scale, int act_dtype, void* bt, cudaStream_t st);
int simt_fwd(const LayerArgs& a, const void* x, void* y, const void* t_save, cudaStream_t st);
int simt_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st);
int simt_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
             int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st);

// warp-level tensor path (mma.sync) for the rank-r adapter contractions; falls back to simt_* when unsupported
int mma_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                  cudaStream_t st);
int mma_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
                cudaStream_t st);
int mma_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
            int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st);

// tcgen05 + TMA form of the same three products (lora_tc.cu); falls back to mma_* when unsupported
int tc_lora_down(const void* x, const void* a, int64_t T, int64_t K, int r, int act_dtype, void* t_save,
                 cudaStream_t st);
int tc_lora_dt(const void* dy, const void* b, int64_t T, int64_t N, int r, float scale, int act_dtype, void* dt_save,
               cudaStream_t st);
int tc_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N, int64_t K,
           int r, int act_dtype, float scale, void* dA, void* dB, float* ws, cudaStream_t st);

// few-token forward: packed-weight streaming kernel (qlora_gemv.cu)
bool gemv_supported(const LayerArgs& a);
int gemv_fwd(const LayerArgs& a, const void* x, void* y, const void* t_save, cudaStream_t st);

// tcgen05 family
bool tc_supported(const LayerArgs& a, bool backward);
int tc_fwd(const LayerArgs& a, const void* x, void* y, void* t_save, cudaStream_t st);
bool tc_fuses_side(const LayerArgs& a, bool backward);  // tc_fwd / tc_bwd_dx will write t_save / dt_save itself
int tc_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st);
// persistent CTA-pair form (qlora_tc2.cu); same contract, used by tc_fwd / tc_bwd_dx for large token counts
bool tc2_preferred(const LayerArgs& a, bool backward);
int64_t tc2_workspace_bytes(int64_t T, int64_t N, int64_t K, int r, bool backward);
int tc2_fwd(const LayerArgs& a, const void* x, void* y, void* t_save, cudaStream_t st);
bool tc2_fuses_side(const LayerArgs& a, bool backward);  // the launch computes t_save / dt_save itself (no side kernel)
bool tc2_fuses_dab(const LayerArgs& a);                  // the backward launch also computes dA, dB (no vft_lora_bwd_dab)
int simt_lora_tt(const void* t_save, int64_t T, int r, int act_dtype, void* tt, cudaStream_t st);
int tc2_bwd_dx(const LayerArgs& a, const void* dy, void* dx, const void* dt_save, cudaStream_t st);

}  // namespace vft
