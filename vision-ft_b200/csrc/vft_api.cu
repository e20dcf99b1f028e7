// extern "C" entry points of libvft_b200.so (declared in include/vft_b200.h) and the
// dispatcher between the tcgen05 family (qlora_tc.cu) and the generic family (qlora_simt.cu).
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>

#include "vft_common.cuh"

namespace vft {

static thread_local char g_error[512] = "";
static thread_local int g_path = VFT_PATH_NONE;
static std::atomic<int> g_forced{0};  // process-wide: backward runs on autograd worker threads

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_error, sizeof(g_error), fmt, ap);
  va_end(ap);
}
void set_path(int path) { g_path = path; }
int forced_path() { return g_forced.load(std::memory_order_relaxed); }

static int check_layer(const LayerArgs& a, const void* act, const void* out, bool out_optional = false) {
  VFT_REQUIRE(a.T >= 0 && a.N > 0 && a.K > 0, "bad shape T=%lld N=%lld K=%lld", (long long)a.T, (long long)a.N,
              (long long)a.K);
  VFT_REQUIRE(a.blocksize >= 2 && a.blocksize % 2 == 0, "bad blocksize %d", a.blocksize);
  VFT_REQUIRE(a.act_dtype == VFT_BF16 || a.act_dtype == VFT_F16 || a.act_dtype == VFT_F32, "bad act_dtype %d",
              a.act_dtype);
  VFT_REQUIRE(a.qdtype == VFT_BF16 || a.qdtype == VFT_F16 || a.qdtype == VFT_F32, "bad qdtype %d", a.qdtype);
  VFT_REQUIRE(a.packed && a.absmax, "packed/absmax must not be null");
  VFT_REQUIRE(a.T == 0 || (act && (out || out_optional)), "activation/output pointers must not be null");
  VFT_REQUIRE(a.r >= 0 && a.r <= VFT_LORA_LD, "LoRA rank %d outside [0, %d]", a.r, VFT_LORA_LD);
  VFT_REQUIRE((a.r == 0) == (a.lora_a == nullptr) && (a.r == 0) == (a.lora_b == nullptr),
              "lora_a/lora_b must be given exactly when r > 0");
  return VFT_OK;
}

// `act`, `out`: the activation / output pointers of the call -- TMA needs them 16-byte aligned; a misaligned view falls
// back to the generic kernels like every other shape the tensor maps cannot describe
static bool use_tc(const LayerArgs& a, bool backward, const void* act, const void* out, int* status) {
  *status = VFT_OK;
  const int forced = forced_path();
  const bool ok = tc_supported(a, backward) && ((reinterpret_cast<uintptr_t>(act) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0;
  if (forced == VFT_PATH_TCGEN05 && !ok) {
    set_error("tcgen05 path forced but shape/dtype not supported (T=%lld N=%lld K=%lld blocksize=%d dtype=%d)",
              (long long)a.T, (long long)a.N, (long long)a.K, a.blocksize, a.act_dtype);
    *status = VFT_ERR_UNSUPPORTED;
    return false;
  }
  if (forced == VFT_PATH_SIMT) return false;
  return ok;
}

void VftEnv::load() {
  *this = VftEnv();
  auto num = [](const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
  };
  auto is = [](const char* name, char c) {
    const char* e = getenv(name);
    return e && e[0] == c;
  };
  pdl = !is("VFT_PDL", '0');
  side_mma = is("VFT_SIDE_MMA", '1');
  side_split = num("VFT_SIDE_SPLIT", 0);
  gemv_tma = num("VFT_GEMV_TMA", 1);
  tc_pair = is("VFT_TC_PAIR", '1');
  tc_debug = num("VFT_TC_DEBUG", 0);
  if (const char* e = getenv("VFT_TC2")) tc2 = e[0] != '0';
  tc2_max_stages = num("VFT_TC2_MAXSTAGES", 0);
  tc2_stages = num("VFT_TC2_STAGES", 0);
  tc2_n_stg = num("VFT_TC2_NSTG", 0);
  if (const char* e = getenv("VFT_TC2_NACC")) {
    if (sscanf(e, "%dx%d", &tc2_force_na, &tc2_force_nn) != 2) tc2_force_na = tc2_force_nn = 0;
  }
  tc2_nosplit = is("VFT_TC2_NOSPLIT", '1');
  tc2_fuse = num("VFT_TC2_FUSE", -1);
  tc2_pingpong = num("VFT_TC2_PINGPONG", -1);
  tc2_job = num("VFT_TC2_JOB", -1);
  tc2_p0kb = num("VFT_TC2_P0KB", 0);
}
static VftEnv& env_mut() {
  static VftEnv e = [] { VftEnv v; v.load(); return v; }();
  return e;
}
const VftEnv& env() { return env_mut(); }
static std::atomic<unsigned> g_env_generation{0};
unsigned env_generation() { return g_env_generation.load(std::memory_order_relaxed); }
void reload_env() {
  env_mut().load();
  g_env_generation.fetch_add(1, std::memory_order_relaxed);  // invalidates per-thread memoised launch plans
}
static bool side_mma() { return env().side_mma; }

}  // namespace vft

using namespace vft;

extern "C" {

int vft_abi_version(void) { return VFT_ABI_VERSION; }
const char* vft_last_error(void) { return g_error; }
int vft_last_path(void) { return g_path; }
void vft_force_path(int path) { g_forced.store(path, std::memory_order_relaxed); }
void vft_reload_env(void) { reload_env(); }

int vft_nf4_quantize(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax,
                     void* stream) {
  VFT_REQUIRE(n >= 0, "n must be >= 0");
  VFT_REQUIRE(n == 0 || (w && packed && absmax), "null pointer");
  return launch_quantize(w, dtype, n, blocksize, packed, absmax, static_cast<cudaStream_t>(stream));
}

int vft_nf4_quantize_many(int count, const void* const* w, int dtype, const int64_t* n, int blocksize,
                          uint8_t* const* packed, float* const* absmax, void* stream) {
  return launch_quantize_many(count, w, dtype, n, blocksize, packed, absmax, static_cast<cudaStream_t>(stream));
}

int vft_nf4_dequantize(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, void* out, int dtype,
                       void* stream) {
  VFT_REQUIRE(n >= 0, "n must be >= 0");
  VFT_REQUIRE(n == 0 || (packed && absmax && out), "null pointer");
  return launch_dequantize(packed, absmax, n, blocksize, out, dtype, static_cast<cudaStream_t>(stream));
}

int vft_nf4_quantize_host(const void* w_host, int dtype, int64_t n, int blocksize, uint8_t* packed_host,
                          float* absmax_host) {
  VFT_REQUIRE(n >= 0 && blocksize >= 2, "bad n/blocksize");
  if (n == 0) return VFT_OK;
  VFT_REQUIRE(w_host && packed_host && absmax_host, "null pointer");
  const size_t esz = dtype == VFT_F32 ? 4 : 2;
  const size_t nb_packed = (size_t)(n + 1) / 2, nb_absmax = sizeof(float) * (size_t)((n + blocksize - 1) / blocksize);
  void *d_w = nullptr, *d_p = nullptr, *d_a = nullptr;
  cudaStream_t st = nullptr;
  int rc = VFT_OK;
  auto cleanup = [&]() {
    if (d_w) cudaFree(d_w);
    if (d_p) cudaFree(d_p);
    if (d_a) cudaFree(d_a);
    if (st) cudaStreamDestroy(st);
  };
#define VFT_HOST_OK(expr)                                                              \
  do {                                                                                 \
    cudaError_t _e = (expr);                                                           \
    if (_e != cudaSuccess) {                                                           \
      set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                       \
      cleanup();                                                                       \
      return VFT_ERR_CUDA;                                                             \
    }                                                                                  \
  } while (0)
  VFT_HOST_OK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  VFT_HOST_OK(cudaMalloc(&d_w, esz * (size_t)n));
  VFT_HOST_OK(cudaMalloc(&d_p, nb_packed));
  VFT_HOST_OK(cudaMalloc(&d_a, nb_absmax));
  VFT_HOST_OK(cudaMemcpyAsync(d_w, w_host, esz * (size_t)n, cudaMemcpyHostToDevice, st));
  rc = launch_quantize(d_w, dtype, n, blocksize, static_cast<uint8_t*>(d_p), static_cast<float*>(d_a), st);
  if (rc != VFT_OK) {
    cleanup();
    return rc;
  }
  VFT_HOST_OK(cudaMemcpyAsync(packed_host, d_p, nb_packed, cudaMemcpyDeviceToHost, st));
  VFT_HOST_OK(cudaMemcpyAsync(absmax_host, d_a, nb_absmax, cudaMemcpyDeviceToHost, st));
  VFT_HOST_OK(cudaStreamSynchronize(st));
#undef VFT_HOST_OK
  cleanup();
  return VFT_OK;
}

int64_t vft_nf4_tiled_bytes(int64_t N, int64_t K, int which) {
  if (N <= 0 || K <= 0 || K % 64 != 0) return 0;
  const int64_t rows = ceil_div64(N, 64) * 64;
  return which == 0 ? rows * K / 2 : rows * (K / 64) * (int64_t)sizeof(float);
}

int vft_nf4_tile_weight(const uint8_t* packed, const float* absmax, int64_t N, int64_t K, int blocksize,
                        uint8_t* codes_t, float* absmax_t, void* stream) {
  VFT_REQUIRE(packed && absmax && codes_t && absmax_t, "null pointer");
  return launch_tile_weight(packed, absmax, N, K, blocksize, codes_t, absmax_t, static_cast<cudaStream_t>(stream));
}

int vft_absmax_nest(const float* absmax, int64_t nblocks, int blocksize2, const float* code256, uint8_t* absmax8,
                    float* absmax2, float* offset, void* ws, int64_t ws_bytes, void* stream) {
  VFT_REQUIRE(absmax && code256 && absmax8 && absmax2 && offset, "null pointer");
  return launch_absmax_nest(absmax, nblocks, blocksize2, code256, absmax8, absmax2, offset, ws, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

int vft_absmax_nest_at(const float* absmax, int64_t nblocks, int blocksize2, const float* code256, const float* offset,
                       uint8_t* absmax8, float* absmax2, void* stream) {
  VFT_REQUIRE(absmax && code256 && offset && absmax8 && absmax2, "null pointer");
  return launch_absmax_nest(absmax, nblocks, blocksize2, code256, absmax8, absmax2, nullptr, nullptr, 0,
                            static_cast<cudaStream_t>(stream), offset);
}

int vft_absmax_denest(const uint8_t* absmax8, const float* absmax2, const float* code256, float offset,
                      int64_t nblocks, int blocksize2, float* absmax_out, void* stream) {
  VFT_REQUIRE(nblocks >= 0, "nblocks must be >= 0");
  VFT_REQUIRE(nblocks == 0 || (absmax8 && absmax2 && code256 && absmax_out), "null pointer");
  return launch_absmax_denest(absmax8, absmax2, code256, offset, nblocks, blocksize2, absmax_out,
                              static_cast<cudaStream_t>(stream));
}

static int64_t bwd_dtt_bytes(int64_t T, int r) {
  const int64_t rows = ((r > 0 ? r : 0) + 15) / 16 * 16;
  return (rows * T * 2 + 255) / 256 * 256;
}

int64_t vft_workspace_bytes(int op, int64_t T, int64_t N, int64_t K, int r) {
  if (op == VFT_OP_ABSMAX_NEST) return absmax_nest_workspace_bytes();
  if (op == VFT_OP_BWD_DAB) return (int64_t)sizeof(float) * (N + K) * (r > 0 ? r : 0);
  if (op == VFT_OP_FWD) return tc2_workspace_bytes(T, N, K, r, false);
  if (op == VFT_OP_BWD_DX) return tc2_workspace_bytes(T, N, K, r, true);
  if (op == VFT_OP_BWD) {  // [dt^T scratch, 256-byte aligned][the larger of the two-call workspaces]
    const int64_t dx = tc2_workspace_bytes(T, N, K, r, true), dab = (int64_t)sizeof(float) * (N + K) * (r > 0 ? r : 0);
    return bwd_dtt_bytes(T, r) + (dx > dab ? dx : dab);
  }
  return 0;
}

int vft_qlora_fwd(const void* x, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                  int blocksize, int act_dtype, int qdtype, const void* bias, const void* lora_a, const void* lora_b,
                  int r, float scale, void* y, void* t_save, void* bt_save, void* tt_save, void* ws, int64_t ws_bytes,
                  const uint8_t* codes_t, const float* absmax_t, void* stream) {
  VFT_REQUIRE((codes_t == nullptr) == (absmax_t == nullptr), "codes_t / absmax_t must be given together");
  LayerArgs a{T, N, K, blocksize, act_dtype, qdtype, r, scale, packed, absmax, bias, lora_a, lora_b, codes_t, absmax_t,
              ws, ws_bytes, r > 0 ? bt_save : nullptr};
  a.tt_save = r > 0 ? tt_save : nullptr;
  int rc = check_layer(a, x, y);
  if (rc != VFT_OK) return rc;
  VFT_REQUIRE(r == 0 || t_save != nullptr, "t_save is required when r > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool gemv = (forced_path() == 0 || forced_path() == VFT_PATH_GEMV) && gemv_supported(a) &&
                    ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15u) == 0;
  if (!gemv && forced_path() == VFT_PATH_GEMV) {
    set_error("streaming path forced but shape/dtype not supported (T=%lld N=%lld K=%lld)", (long long)T, (long long)N,
              (long long)K);
    return VFT_ERR_UNSUPPORTED;
  }
  bool tc = false;
  if (!gemv) {
    tc = use_tc(a, false, x, y, &rc);
    if (rc != VFT_OK) return rc;
  }
  // t_save = x . A^T: inside the GEMM launch when the persistent tcgen05 kernel takes the call unsplit, else a kernel
  // of its own in front of it
  if (r > 0 && !(tc && tc_fuses_side(a, false))) {
    rc = (forced_path() == VFT_PATH_SIMT) ? simt_lora_down(x, lora_a, T, K, r, act_dtype, t_save, st)
                                          : side_mma() ? mma_lora_down(x, lora_a, T, K, r, act_dtype, t_save, st)
                                                       : tc_lora_down(x, lora_a, T, K, r, act_dtype, t_save, st);
    if (rc != VFT_OK) return rc;
    // t^T for the backward's dA/dB job (the fused launch writes it itself).  A backward whose contraction is split keeps
    // its side kernels and never reads it (choose_tc2: the plan does not depend on r or on the workspace): no launch then
    if (a.tt_save != nullptr && tc2_workspace_bytes(T, N, K, r, /*backward=*/true) == 0) {
      rc = simt_lora_tt(t_save, T, r, act_dtype, a.tt_save, st);
      if (rc != VFT_OK) return rc;
    }
  }
  // bt_save = s * B^T for the backward call: the persistent tcgen05 forward writes it on its way, every other path
  // gets a small kernel
  if (a.bt_save != nullptr && !(tc && tc2_preferred(a, false))) {
    rc = simt_lora_bt(lora_b, N, r, scale, act_dtype, a.bt_save, st);
    if (rc != VFT_OK) return rc;
  }
  if (gemv) {
    set_path(VFT_PATH_GEMV);  // T <= 8: the launch is a weight stream, not a GEMM
    return gemv_fwd(a, x, y, t_save, st);
  }
  set_path(tc ? VFT_PATH_TCGEN05 : VFT_PATH_SIMT);
  return tc ? tc_fwd(a, x, y, t_save, st) : simt_fwd(a, x, y, t_save, st);
}

int vft_qlora_bwd_dx(const void* dy, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                     int blocksize, int act_dtype, int qdtype, const void* lora_a, const void* lora_b, int r,
                     float scale, void* dx, void* dt_save, const void* bt_save, void* ws, int64_t ws_bytes,
                     const uint8_t* codes_t, const float* absmax_t, void* stream) {
  VFT_REQUIRE((codes_t == nullptr) == (absmax_t == nullptr), "codes_t / absmax_t must be given together");
  LayerArgs a{T, N, K, blocksize, act_dtype, qdtype, r, scale, packed, absmax, nullptr, lora_a, lora_b, codes_t, absmax_t,
              ws, ws_bytes, r > 0 ? const_cast<void*>(bt_save) : nullptr};
  int rc = check_layer(a, dy, dx, /*out_optional=*/true);  // dx == NULL: only dt_save is wanted
  if (rc != VFT_OK) return rc;
  VFT_REQUIRE(r == 0 || dt_save != nullptr, "dt_save is required when r > 0");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  bool tc = false;
  if (dx != nullptr) {
    tc = use_tc(a, true, dy, dx, &rc);
    if (rc != VFT_OK) return rc;
  }
  // dt_save = s * dy . B: inside the GEMM launch when the persistent tcgen05 kernel takes the call unsplit and the
  // forward left bt_save, else a kernel of its own in front of it
  if (r > 0 && !(tc && tc_fuses_side(a, true))) {
    rc = (forced_path() == VFT_PATH_SIMT) ? simt_lora_dt(dy, lora_b, T, N, r, scale, act_dtype, dt_save, st)
                                          : side_mma() ? mma_lora_dt(dy, lora_b, T, N, r, scale, act_dtype, dt_save, st)
                                                       : tc_lora_dt(dy, lora_b, T, N, r, scale, act_dtype, dt_save, st);
    if (rc != VFT_OK) return rc;
  }
  if (dx == nullptr) {
    set_path(VFT_PATH_SIMT);
    return VFT_OK;
  }
  set_path(tc ? VFT_PATH_TCGEN05 : VFT_PATH_SIMT);
  return tc ? tc_bwd_dx(a, dy, dx, dt_save, st) : simt_bwd_dx(a, dy, dx, dt_save, st);
}

int vft_qlora_bwd(const void* dy, const void* x, int64_t T, const uint8_t* packed, const float* absmax, int64_t N, int64_t K,
                  int blocksize, int act_dtype, int qdtype, const void* lora_a, const void* lora_b, int r, float scale,
                  const void* t_save, const void* tt_save, const void* bt_save, void* dx, void* dA, void* dB,
                  void* dt_save, void* ws, int64_t ws_bytes, const uint8_t* codes_t, const float* absmax_t, void* stream) {
  VFT_REQUIRE(r > 0 && r <= VFT_LORA_LD, "LoRA rank %d outside [1, %d]", r, VFT_LORA_LD);
  VFT_REQUIRE(dA && dB && dt_save && (T == 0 || (dy && x && t_save)), "null pointer");
  const int64_t need = vft_workspace_bytes(VFT_OP_BWD, T, N, K, r);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)need, (long long)ws_bytes);
    return VFT_ERR_WORKSPACE;
  }
  const int64_t dtt_bytes = bwd_dtt_bytes(T, r);
  void* ws_rest = static_cast<char*>(ws) + dtt_bytes;
  const int64_t rest_bytes = ws_bytes - dtt_bytes;
  if (T > 0 && dx != nullptr) {
    // one launch: dx, dt (side product) and dA / dB (column-tile job) inside the persistent tcgen05 kernel
    LayerArgs a{T, N, K, blocksize, act_dtype, qdtype, r, scale, packed, absmax, nullptr, lora_a, lora_b, codes_t, absmax_t,
                ws_rest, rest_bytes, const_cast<void*>(bt_save)};
    a.tt_save = const_cast<void*>(tt_save);
    a.job_x = x;
    a.job_dtt = ws;
    a.job_da = dA;
    a.job_db = dB;
    int rc = check_layer(a, dy, dx);
    if (rc != VFT_OK) return rc;
    const bool tc = use_tc(a, true, dy, dx, &rc);
    if (rc != VFT_OK) return rc;
    if (tc && tc_fuses_side(a, true) && tc2_fuses_dab(a)) {
      set_path(VFT_PATH_TCGEN05);
      return tc_bwd_dx(a, dy, dx, dt_save, static_cast<cudaStream_t>(stream));
    }
  }
  // two calls: input gradient (+ dt), then the adapter's weight gradients
  int rc = vft_qlora_bwd_dx(dy, T, packed, absmax, N, K, blocksize, act_dtype, qdtype, lora_a, lora_b, r, scale, dx, dt_save,
                            bt_save, ws_rest, rest_bytes, codes_t, absmax_t, stream);
  if (rc != VFT_OK) return rc;
  const int path = g_path;
  rc = vft_lora_bwd_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, ws_rest, rest_bytes, stream);
  set_path(path);
  return rc;
}

int vft_lora_bwd_dab(const void* dy, const void* x, const void* t_save, const void* dt_save, int64_t T, int64_t N,
                     int64_t K, int r, int act_dtype, float scale, void* dA, void* dB, void* ws, int64_t ws_bytes,
                     void* stream) {
  VFT_REQUIRE(r > 0 && r <= VFT_LORA_LD, "LoRA rank %d outside [1, %d]", r, VFT_LORA_LD);
  VFT_REQUIRE(T >= 0 && N > 0 && K > 0, "bad shape");
  VFT_REQUIRE(dA && dB && (T == 0 || (dy && x && t_save && dt_save)), "null pointer");
  const int64_t need = vft_workspace_bytes(VFT_OP_BWD_DAB, T, N, K, r);
  if (ws == nullptr || ws_bytes < need) {
    set_error("workspace too small: need %lld bytes, got %lld", (long long)need, (long long)ws_bytes);
    return VFT_ERR_WORKSPACE;
  }
  set_path(VFT_PATH_SIMT);
  if (forced_path() == VFT_PATH_SIMT)
    return simt_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, static_cast<float*>(ws),
                    static_cast<cudaStream_t>(stream));
  if (side_mma())
    return mma_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, static_cast<float*>(ws),
                   static_cast<cudaStream_t>(stream));
  return tc_dab(dy, x, t_save, dt_save, T, N, K, r, act_dtype, scale, dA, dB, static_cast<float*>(ws),
                static_cast<cudaStream_t>(stream));
}

}  // extern "C"
