// planes.cuh — bf16 operand planes for the tensor-core GEMMs.
// An fp32 activation x is handed to the tcgen05 GEMM either rounded to bf16 (parts = 1, LDS_PREC_BF16) or as three
// bf16 planes hi, mid, lo with hi + mid + lo == x to 24 bits (parts = 3, fp32-accurate split mode).  A row of C
// channels is stored as [plane 0 | plane 1 | plane 2], i.e. element (p, c) at p*C + c.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace lds {

// Packs (a, b) to bf16x2 with round-to-nearest and leaves the residuals a - bf16(a), b - bf16(b) in place: one
// cvt.rn.bf16x2.f32 (full-rate F2FP) + two integer ops + two FADDs per pair instead of per-element F2F conversions.
__device__ __forceinline__ uint32_t planes_split_pair(float& a, float& b) {
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));   // low half <- a, high half <- b
  a -= __uint_as_float(w << 16);
  b -= __uint_as_float(w & 0xffff0000u);
  return w;
}

__device__ __forceinline__ void store_planes4(__nv_bfloat16* row, int c, int C, int parts, float v0, float v1, float v2, float v3) {
  for (int p = 0; p < parts; ++p) {
    uint2 w;
    w.x = planes_split_pair(v0, v1);
    w.y = planes_split_pair(v2, v3);
    *reinterpret_cast<uint2*>(row + (size_t)p * C + c) = w;
  }
}

}  // namespace lds
