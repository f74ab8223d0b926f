// planes.cuh — 16-bit operand planes for the tensor-core GEMMs.
// An fp32 activation x is handed to the tcgen05 GEMM in one of three forms (a row of C channels is stored as
// [plane 0 | plane 1 | ...], i.e. element (p, c) at p*C + c, two bytes per element in every form):
//   parts = 1  bf16(x)                                                    LDS_PREC_BF16
//   parts = 2  SPLIT-F16 (fp32-accurate mode, round 2): two fp16 planes of the scaled value x*16,
//              h1 = f16(16x), h2 = f16(16x - h1): 22 significant bits; the GEMM evaluates h1*w1 + h1*w2 + h2*w1 — THREE
//              tensor-core products per logical product instead of six.  The power-of-two scale keeps the second plane of
//              O(1) activations / O(0.05) weights out of fp16's subnormal range (without it the split loses 2-4 bits, which is
//              why round 1 dismissed fp16 planes); measured operand error of the 3-product sum: relative L2 7.6e-8 for
//              K = 256 ... 3072, a third of an fp32 FFMA GEMM's own rounding error (profiles/r02_split_f16_operand_error.txt).
//              Values beyond +-4094 saturate (cvt.satfinite) instead of overflowing to inf.
//   parts = 3  three bf16 planes hi, mid, lo with hi + mid + lo == x to 24 bits (round 1's split; still used for the
//              attention operands Q, K, V^T and P, whose kernel evaluates the six significant plane products).
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace lds {

constexpr float PLANE_SCALE = 16.f;              // activation planes of the split-f16 form hold x * PLANE_SCALE
constexpr float PLANE_INV_SCALE = 1.f / 16.f;

// Packs (a, b) to bf16x2 with round-to-nearest and leaves the residuals a - bf16(a), b - bf16(b) in place: one
// cvt.rn.bf16x2.f32 (full-rate F2FP) + two integer ops + two FADDs per pair instead of per-element F2F conversions.
__device__ __forceinline__ uint32_t planes_split_pair(float& a, float& b) {
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));   // low half <- a, high half <- b
  a -= __uint_as_float(w << 16);
  b -= __uint_as_float(w & 0xffff0000u);
  return w;
}
// fp16 form of the same: (a, b) are ALREADY scaled; F2FP.SATFINITE + two HADD2.F32 + two FADDs
__device__ __forceinline__ uint32_t planes_split_pair_f16(float& a, float& b) {
  uint32_t w;
  float lo, hi;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  asm("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(w));
  a -= lo;
  b -= hi;
  return w;
}
__device__ __forceinline__ uint32_t planes_pack_pair_f16(float a, float b) {
  uint32_t w;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  return w;
}

__device__ __forceinline__ void store_planes4(__nv_bfloat16* row, int c, int C, int parts, float v0, float v1, float v2, float v3) {
  if (parts == 2) {
    v0 *= PLANE_SCALE; v1 *= PLANE_SCALE; v2 *= PLANE_SCALE; v3 *= PLANE_SCALE;
    uint2 w;
    w.x = planes_split_pair_f16(v0, v1);
    w.y = planes_split_pair_f16(v2, v3);
    *reinterpret_cast<uint2*>(row + c) = w;
    w.x = planes_pack_pair_f16(v0, v1);
    w.y = planes_pack_pair_f16(v2, v3);
    *reinterpret_cast<uint2*>(row + (size_t)C + c) = w;
    return;
  }
  for (int p = 0; p < parts; ++p) {
    uint2 w;
    w.x = planes_split_pair(v0, v1);
    w.y = planes_split_pair(v2, v3);
    *reinterpret_cast<uint2*>(row + (size_t)p * C + c) = w;
  }
}

}  // namespace lds
