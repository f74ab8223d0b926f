// host_pack.h — host-side repacking of fp32 weights into the 16-bit operand planes of the tensor-core GEMMs (planes.cuh) and
// into aligned fp32 arenas; shared by the sampler (lds_api.cu) and the units front-end (units.cu).  Internal header.
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <cstdint>
#include <cstring>
#include <vector>

namespace {

inline uint16_t host_f2bf(float f) {     // round-to-nearest-even, identical to __float2bfloat16_rn
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
inline float host_bf2f(uint16_t b) {
  const uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// fp32 [rows][cin] -> 16-bit planes [rows][parts][cin] (planes.cuh): parts 1 / 3 bf16 planes; parts 2 split-f16 planes of
// the scaled value (h1 = f16(w * scale), h2 = f16(w * scale - h1)), bit-identical to what the device conversion would give
struct PlanePacker {
  std::vector<uint16_t> host;
  float scale = 1.f;
  static uint16_t f2h_sat(float f) {
    if (f > 65504.f) f = 65504.f;
    if (f < -65504.f) f = -65504.f;
    const __half_raw r = __float2half_rn(f);
    return r.x;
  }
  static float h2f(uint16_t b) {
    __half_raw r;
    r.x = b;
    return __half2float(__half(r));
  }
  size_t add(const float* src, size_t rows, int cin, int parts) {
    size_t off = (host.size() + 127) / 128 * 128;   // 256-byte alignment
    host.resize(off + rows * parts * cin);
    for (size_t r = 0; r < rows; ++r)
      for (int c = 0; c < cin; ++c) {
        float v = src[r * cin + c];
        if (parts == 2) {
          v *= scale;
          const uint16_t h1 = f2h_sat(v);
          host[off + (r * 2 + 0) * cin + c] = h1;
          host[off + (r * 2 + 1) * cin + c] = f2h_sat(v - h2f(h1));
          continue;
        }
        for (int p = 0; p < parts; ++p) {
          const uint16_t b = host_f2bf(v);
          host[off + (r * parts + p) * cin + c] = b;
          v -= host_bf2f(b);
        }
      }
    return off;
  }
};

struct Packer {
  std::vector<float> host;
  size_t add(const float* src, size_t n) {
    size_t off = host.size();
    off = (off + 63) / 64 * 64;       // 256-byte alignment of every tensor
    host.resize(off + n);
    if (src) memcpy(host.data() + off, src, n * sizeof(float));
    return off;
  }
};


}  // namespace
