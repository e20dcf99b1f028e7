// Host-side check of the PRMT look-up network in vision-ft_b200/csrc/nf4_lut.cuh:
// decode every packed word pattern class against the plain "codebook[code] * absmax, round" decode.
// Built and run on the CPU by tests/test_lut_host.py (no GPU needed).
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../vision-ft_b200/csrc/nf4_lut.cuh"

using namespace vft;

template <typename ActT>
static uint16_t plain(unsigned code, float am, int qdtype) {
  const float kCode[16] = {-1.0f, -0.6961928009986877f, -0.5250730514526367f, -0.39491748809814453f,
                           -0.28444138169288635f, -0.18477343022823334f, -0.09105003625154495f, 0.0f,
                           0.07958029955625534f, 0.16093020141124725f, 0.24611230194568634f, 0.33791524171829224f,
                           0.44070982933044434f, 0.5626170039176941f, 0.7229568362236023f, 1.0f};
  float v = kCode[code] * am;
  if (qdtype != VFT_F32) v = round_to_qdtype(v, qdtype);
  return (uint16_t)(pack2<ActT>(v, 0.0f) & 0xFFFF);
}

template <typename ActT>
static long run(int qdtype) {
  long bad = 0;
  uint32_t seed = 12345u;
  for (int trial = 0; trial < 20000; ++trial) {
    seed = seed * 1664525u + 1013904223u;
    uint32_t w = seed;
    seed = seed * 1664525u + 1013904223u;
    float am = (float)((seed >> 8) & 0xFFFF) / 65536.0f * 0.2f;
    if (trial % 97 == 0) am = 0.0f;
    if (trial < 16) w = 0x11111111u * trial;  // all-equal codes
    Nf4Lut t;
    nf4_build_lut<ActT>(am, qdtype, t);
    uint32_t out[4];
    nf4_decode_word(w, t, out);
    for (int e = 0; e < 8; ++e) {
      const unsigned byte = (w >> (8 * (e >> 1))) & 0xFF;
      const unsigned code = (e & 1) ? (byte & 15u) : (byte >> 4);
      const uint16_t got = (uint16_t)((out[e >> 1] >> (16 * (e & 1))) & 0xFFFF);
      if (got != plain<ActT>(code, am, qdtype)) ++bad;
    }
  }
  return bad;
}

int main() {
  long bad = 0;
  bad += run<__nv_bfloat16>(VFT_BF16);
  bad += run<__nv_bfloat16>(VFT_F16);  // fp16 checkpoint used with bf16 activations: double rounding
  bad += run<__half>(VFT_F16);
  bad += run<__half>(VFT_F32);
  printf("mismatches=%ld\n", bad);
  return bad == 0 ? 0 : 1;
}
