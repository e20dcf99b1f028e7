// Register-resident NF4 look-up used by the fused tcgen05 kernels.
//
// For one 64-element quantization block the 16 possible dequantized values
//     v_j = ActT( round_qdtype( codebook[j] * absmax ) )
// are computed ONCE (16 fp32 multiplies + 8 pack-converts: bit-identical to what bitsandbytes'
// dequantize_4bit + .to(x.dtype) would store in its bf16 copy of W) and kept in 8 registers,
// split into a low-byte table L[0..3] and a high-byte table H[0..3].  Eight 4-bit codes (one
// 32-bit word of the packed weight) are then decoded with byte permutes only:
//     sel   = w & 0x7777                  low 3 bits of each code -> PRMT selector nibbles
//     X     = prmt(L0, L1, sel)           candidates from entries 0..7
//     Y     = prmt(L2, L3, sel)           candidates from entries 8..15
//     pick  = ((w >> 1) & 0x4444) | 0x3210  bit 3 of each code picks X or Y per byte
//     lo4   = prmt(X, Y, pick)            4 low bytes; same again with H for 4 high bytes
//     out   = prmt(lo4, hi4, 0x4051 / 0x6273)   interleave into two packed 16-bit pairs
// i.e. 2.6 ALU instructions per weight and no shared-memory traffic for the table.
//
// Everything is __host__ __device__ (PRMT emulated on the host) so tests/test_lut_host.py can
// check the permute network against the plain decode on the CPU.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <string.h>

#include <type_traits>

#include "../../include/vft_b200.h"

#define VFT_HD __host__ __device__ __forceinline__

namespace vft {

VFT_HD uint32_t prmt_hd(uint32_t a, uint32_t b, uint32_t sel) {
#ifdef __CUDA_ARCH__
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
#else
  const uint64_t src = ((uint64_t)b << 32) | a;
  uint32_t d = 0;
  for (int i = 0; i < 4; ++i) {
    const uint32_t s = (sel >> (4 * i)) & 0xF;
    uint32_t byte = (uint32_t)(src >> (8 * (s & 7))) & 0xFF;
    if (s & 8) byte = (byte & 0x80) ? 0xFF : 0x00;
    d |= byte << (8 * i);
  }
  return d;
#endif
}

template <typename ActT>
VFT_HD uint32_t pack2(float lo, float hi);
template <>
VFT_HD uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);  // one cvt.rn.bf16x2.f32 (F2FP), not two F2F
  return *reinterpret_cast<const uint32_t*>(&v);
}
template <>
VFT_HD uint32_t pack2<__half>(float lo, float hi) {
  const __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}

VFT_HD float round_to_qdtype(float v, int qdtype) {
  if (qdtype == VFT_F16) return __half2float(__float2half_rn(v));
  if (qdtype == VFT_BF16) return __bfloat162float(__float2bfloat16_rn(v));
  return v;
}

// Round a pair through quant_state.dtype when it differs from the activation dtype (double rounding, as
// bitsandbytes' dequantize_4bit(...).to(x.dtype) does).  Packed conversions only: the scalar F2F path runs on
// the XU pipe at 1/8 rate and was THE bottleneck of the first version of the fused kernel (74 % XU busy).
template <typename ActT>
VFT_HD void round_pair_through(float& a, float& b, int qdtype) {
  constexpr int kAct = sizeof(ActT) == 2 ? (std::is_same<ActT, __half>::value ? VFT_F16 : VFT_BF16) : VFT_F32;
  if (qdtype == kAct || qdtype == VFT_F32) return;  // a second rounding to the same format changes nothing
  if (qdtype == VFT_F16) {
    const float2 f = __half22float2(__floats2half2_rn(a, b));
    a = f.x;
    b = f.y;
  } else {
    const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
    const uint32_t u = *reinterpret_cast<const uint32_t*>(&v);
#ifdef __CUDA_ARCH__
    a = __uint_as_float(u << 16);
    b = __uint_as_float(u & 0xffff0000u);
#else
    const uint32_t ua = u << 16, ub = u & 0xffff0000u;
    memcpy(&a, &ua, 4);
    memcpy(&b, &ub, 4);
#endif
  }
}

struct Nf4Lut {
  uint32_t L[4];  // low bytes of the 16 scaled values, entry j in byte j%4 of L[j/4]
  uint32_t H[4];  // high bytes
};

template <typename ActT>
VFT_HD void nf4_build_lut(float absmax, int qdtype, Nf4Lut& t) {
  const float kCode[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                           -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                           0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                           0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};
  uint32_t p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float a = kCode[2 * i] * absmax, b = kCode[2 * i + 1] * absmax;
    round_pair_through<ActT>(a, b, qdtype);
    p[i] = pack2<ActT>(a, b);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    t.L[q] = prmt_hd(p[2 * q], p[2 * q + 1], 0x6420);
    t.H[q] = prmt_hd(p[2 * q], p[2 * q + 1], 0x7531);
  }
}

// Decode the 8 codes of one packed word (flat elements e..e+7, element 2j in the high nibble of
// byte j) into 4 registers of two 16-bit values each, in element order.
VFT_HD void nf4_decode_word(uint32_t w, const Nf4Lut& t, uint32_t (&out)[4]) {
  const uint32_t sel = w & 0x77777777u;
  const uint32_t pick = ((w >> 1) & 0x44444444u) | 0x32103210u;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const uint32_t s = h ? (sel >> 16) : sel;
    const uint32_t pk = h ? (pick >> 16) : pick;
    const uint32_t lo4 = prmt_hd(prmt_hd(t.L[0], t.L[1], s), prmt_hd(t.L[2], t.L[3], s), pk);
    const uint32_t hi4 = prmt_hd(prmt_hd(t.H[0], t.H[1], s), prmt_hd(t.H[2], t.H[3], s), pk);
    // result byte i of lo4/hi4 belongs to nibble i of the half-word = element (i ^ 1) of this half
    out[2 * h] = prmt_hd(lo4, hi4, 0x4051);      // elements 0,1 of the half
    out[2 * h + 1] = prmt_hd(lo4, hi4, 0x6273);  // elements 2,3 of the half
  }
}

}  // namespace vft
