/* Plain-C restatement of the blockwise NF4 quantize/pack and dequantize steps.
 *
 * TEST INFRASTRUCTURE ONLY: checker for tests/ and the timed CPU baseline of
 * bench.py ("port"); never linked into the product library.
 *
 * PARITY UNPINNED: the algorithm is bitsandbytes 0.48.2's (kQuantizeBlockwise /
 * kDequantizeBlockwise with the NF4 data type), a dependency that is pinned at
 * /root/reference/uv.lock:307-308 but not vendored in /root/reference.  Call
 * sites in the reference: src/modules/quant/functional.py:362-368
 * (quantize_4bit), src/modules/quant/bnb.py:94-99,122-129 (Params4bit).
 * Semantics follow oracle/nf4_oracle.py (same header, same edge cases) and are
 * cross-checked against it in tests/test_oracle.py.
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

enum { VFT_F32 = 0, VFT_F16 = 1, VFT_BF16 = 2 };

static const float kCodebook[16] = {
    -1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
    -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
    0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
    0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};

static const float kThresholds[15] = {
    -0.8480964004993439f, -0.6106329262256622f, -0.4599952697753906f, -0.33967943489551544f,
    -0.23460740596055984f, -0.13791173323988914f, -0.045525018125772476f, 0.03979014977812767f,
    0.1202552504837513f, 0.2035212516784668f, 0.2920137718319893f, 0.3893125355243683f,
    0.5016634166240692f, 0.6427869200706482f, 0.8614784181118011f};

static inline float bf16_to_f32(uint16_t h) {
  uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* quiet NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static inline float load_elem(const void* w, int dtype, int64_t i) {
  switch (dtype) {
    case VFT_F32: return ((const float*)w)[i];
    case VFT_F16: return (float)((const _Float16*)w)[i];
    default: return bf16_to_f32(((const uint16_t*)w)[i]);
  }
}

/* code = number of thresholds strictly below x (NaN compares false -> 0). */
static inline unsigned encode_nf4(float x) {
  unsigned c = 0;
  for (int t = 0; t < 15; ++t) c += (x > kThresholds[t]);
  return c;
}

int nf4_quantize_ref(const void* w, int dtype, int64_t n, int blocksize, uint8_t* packed, float* absmax) {
  if (blocksize <= 0 || (blocksize & 1)) return -1;
  const int64_t nblocks = (n + blocksize - 1) / blocksize;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < nblocks; ++b) {
    const int64_t lo = b * blocksize;
    const int64_t hi = lo + blocksize < n ? lo + blocksize : n;
    float am = 0.0f;
    for (int64_t i = lo; i < hi; ++i) {
      float a = fabsf(load_elem(w, dtype, i));
      am = a > am ? a : am; /* NaN inputs are outside the contract */
    }
    absmax[b] = am;
    const volatile float one = 1.0f;
    const float s = one / am; /* IEEE division; +inf for an all-zero block */
    for (int64_t i = lo; i < hi; i += 2) {
      unsigned c0 = encode_nf4(load_elem(w, dtype, i) * s);
      unsigned c1 = (i + 1 < n) ? encode_nf4(load_elem(w, dtype, i + 1) * s) : 0u;
      packed[i >> 1] = (uint8_t)((c0 << 4) | c1);
    }
  }
  return 0;
}

int nf4_dequantize_ref(const uint8_t* packed, const float* absmax, int64_t n, int blocksize, void* out, int dtype) {
  if (blocksize <= 0) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const uint8_t byte = packed[i >> 1];
    const unsigned c = (i & 1) ? (byte & 0xFu) : (byte >> 4);
    const float v = kCodebook[c] * absmax[i / blocksize];
    switch (dtype) {
      case VFT_F32: ((float*)out)[i] = v; break;
      case VFT_F16: ((_Float16*)out)[i] = (_Float16)v; break;
      default: ((uint16_t*)out)[i] = f32_to_bf16_rne(v); break;
    }
  }
  return 0;
}

/* ---- nested ("double quant") block statistics: quantize_blockwise(absmax - mean, 256) with the 8-bit dynamic map.
 * Same status (parity unpinned) and same semantics as oracle/nf4_oracle.py absmax_nest / absmax_denest. */
static unsigned dquantize8(const float* code, float x) { /* bitsandbytes dQuantize<0> */
  int pivot = 127, upper_pivot = 255, lower_pivot = 0;
  float lower = -1.0f, upper = 1.0f, val = code[pivot];
  for (int i = 64; i > 0; i >>= 1) {
    if (x > val) { lower_pivot = pivot; lower = val; pivot += i; }
    else { upper_pivot = pivot; upper = val; pivot -= i; }
    val = code[pivot];
  }
  if (upper_pivot == 255) upper = code[upper_pivot];
  if (lower_pivot == 0) lower = code[lower_pivot];
  if (x > val) {
    const float mid = (upper + val) * 0.5f;
    return (unsigned)(x > mid ? upper_pivot : pivot);
  }
  const float mid = (lower + val) * 0.5f;
  return (unsigned)(x < mid ? lower_pivot : pivot);
}

int absmax_nest_ref(const float* absmax, int64_t n, int blocksize2, const float* code, uint8_t* q, float* absmax2,
                    float* offset_out) {
  if (n <= 0 || blocksize2 <= 0) return -1;
  double s = 0.0;
  for (int64_t i = 0; i < n; ++i) s += (double)absmax[i];
  const float off = (float)(s / (double)n);
  *offset_out = off;
  const int64_t nb = (n + blocksize2 - 1) / blocksize2;
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < nb; ++b) {
    const int64_t lo = b * blocksize2, hi = lo + blocksize2 < n ? lo + blocksize2 : n;
    float m = 0.0f;
    for (int64_t i = lo; i < hi; ++i) {
      const float a = fabsf(absmax[i] - off);
      m = a > m ? a : m;
    }
    absmax2[b] = m;
    const volatile float one = 1.0f;
    const float inv = one / m;
    for (int64_t i = lo; i < hi; ++i) q[i] = (uint8_t)dquantize8(code, (absmax[i] - off) * inv);
  }
  return 0;
}

int absmax_denest_ref(const uint8_t* q, const float* absmax2, const float* code, float offset, int64_t n,
                      int blocksize2, float* out) {
  if (blocksize2 <= 0) return -1;
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; ++i) {
    const volatile float prod = code[q[i]] * absmax2[i / blocksize2];
    out[i] = prod + offset;
  }
  return 0;
}
