// planes.cuh — bf16 operand planes for the tensor-core GEMMs.
// An fp32 activation x is handed to the tcgen05 GEMM either rounded to bf16 (parts = 1, LDS_PREC_BF16) or as three
// bf16 planes hi, mid, lo with hi + mid + lo == x to 24 bits (parts = 3, fp32-accurate split mode).  A row of C
// channels is stored as [plane 0 | plane 1 | plane 2], i.e. element (p, c) at p*C + c.
#pragma once
#include <cuda_bf16.h>
#include <stdint.h>

namespace lds {

__device__ __forceinline__ void store_planes4(__nv_bfloat16* row, int c, int C, int parts, float v0, float v1, float v2, float v3) {
  float r[4] = {v0, v1, v2, v3};
  for (int p = 0; p < parts; ++p) {
    __nv_bfloat16 h[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      h[j] = __float2bfloat16_rn(r[j]);
      r[j] -= __bfloat162float(h[j]);
    }
    uint2 w;
    w.x = (uint32_t)__bfloat16_as_ushort(h[0]) | ((uint32_t)__bfloat16_as_ushort(h[1]) << 16);
    w.y = (uint32_t)__bfloat16_as_ushort(h[2]) | ((uint32_t)__bfloat16_as_ushort(h[3]) << 16);
    *reinterpret_cast<uint2*>(row + (size_t)p * C + c) = w;
  }
}

}  // namespace lds
