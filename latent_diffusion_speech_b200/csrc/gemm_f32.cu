// gemm_f32.cu — IEEE-fp32 (FFMA) implicit-GEMM for Conv1d(k=3, stride 1/2, optional fused
// nearest-upsample), Conv1d(k=1) and nn.Linear on channels-last activations.
//
// This is the fp32-accurate path (LDS_PREC_FP32): the parity bar of BASELINE.json (max-abs
// <= 1e-3 on outputs of magnitude ~7e2) sits on the fp32 round-off floor of the reference
// itself, so plain-TF32/bf16 tensor-core products cannot be used there.  The bf16 mode uses
// the tcgen05 kernels in gemm_tc.cu instead.
//
// Reference ops replaced: F.conv1d (diffusion/unet1d/lora.py:102), nn.Linear
// (attention_processor.py:1012-1040, attention.py:291,247), F.interpolate(nearest) feeding the
// upsampler conv (resnet.py:157-169), stride-2 downsample conv (resnet.py:200,221), GEGLU
// (attention.py:299-301), residual adds (resnet.py:639, attention.py:161-201, transformer_1d.py:295).
//
// Tiling: 128x128x16 CTA tile, 256 threads, 8x8 register micro-tile per thread (two 4-wide
// strips 64 apart so shared-memory reads are conflict-free float4s), double-buffered shared
// memory with register prefetch of the next K slice.
#include "lds_kernels.h"

namespace lds {
namespace {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;
constexpr int SLD = BM + 4;  // padded leading dim of the transposed tiles (keeps float4 alignment)

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

__global__ void __launch_bounds__(NT, 2) gemm_f32_kernel(const GemmArgs p) {
  __shared__ __align__(16) float As[2][BK][SLD];
  __shared__ __align__(16) float Bs[2][BK][SLD];

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ld_row = tid >> 2, ld_kv = (tid & 3) * 4;

  // ---- per-thread source-row bookkeeping for the two A rows / two W rows it stages ----
  int a_b[2], a_u0[2];
  bool a_ok[2];
  const float* w_ptr[2];
  bool w_ok[2];
  const int pad = (p.taps == 3) ? 1 : 0;
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int m = m0 + ld_row + 64 * r;
    a_ok[r] = m < p.M;
    const int mm = a_ok[r] ? m : 0;
    a_b[r] = mm / p.t_out;
    a_u0[r] = (mm - a_b[r] * p.t_out) * p.stride - pad;
    const int n = n0 + ld_row + 64 * r;
    w_ok[r] = n < p.N;
    w_ptr[r] = p.W + (size_t)(w_ok[r] ? n : 0) * p.K + ld_kv;
  }

  float4 ra[2], rb[2];
  auto fetch = [&](int k0) {
    const int tap = k0 / p.cin;
    const int c0 = k0 - tap * p.cin;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int u = a_u0[r] + tap;
      bool ok = a_ok[r] && u >= 0 && u < p.t_conv;
      int src = u;
      if (p.upsample) src = min((int)floorf((float)u * p.up_scale), p.t_in - 1);
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok) ra[r] = __ldg(reinterpret_cast<const float4*>(p.A + ((size_t)a_b[r] * p.t_in + src) * p.a_ld + c0 + ld_kv));
      rb[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (w_ok[r]) rb[r] = __ldg(reinterpret_cast<const float4*>(w_ptr[r] + k0));
    }
  };
  auto stage = [&](int buf) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int row = ld_row + 64 * r;
      As[buf][ld_kv + 0][row] = ra[r].x; As[buf][ld_kv + 1][row] = ra[r].y;
      As[buf][ld_kv + 2][row] = ra[r].z; As[buf][ld_kv + 3][row] = ra[r].w;
      Bs[buf][ld_kv + 0][row] = rb[r].x; Bs[buf][ld_kv + 1][row] = rb[r].y;
      Bs[buf][ld_kv + 2][row] = rb[r].z; Bs[buf][ld_kv + 3][row] = rb[r].w;
    }
  };

  const int tx = tid & 15, ty = tid >> 4;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = p.K / BK;
  fetch(0);
  stage(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) fetch((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[cur][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[cur][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[cur][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[cur][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) stage(cur ^ 1);
    __syncthreads();
  }

  // ---- epilogue ----
  const int cq0 = n0 + tx * 4, cq1 = n0 + 64 + tx * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= p.M) continue;
    const float* bias = p.bias;
    const size_t rrow = (size_t)(m / p.r_div) * p.r_ld;
    if (p.epilogue == EPI_GEGLU) {
      if (cq1 + 3 < p.N) {
        float o[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float v = acc[i][j], g = acc[i][4 + j];
          if (bias) { v += bias[cq0 + j]; g += bias[cq1 + j]; }
          o[j] = v * gelu_erf(g);
        }
        const int oc = (n0 >> 1) + tx * 4;
        float* dst = p.C + (size_t)m * p.c_ld + oc;
        if (p.R) {
          const float4 r = *reinterpret_cast<const float4*>(p.R + rrow + oc);
          o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
        }
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      }
      continue;
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int c = q ? cq1 : cq0;
      if (c + 3 >= p.N) continue;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[i][q * 4 + j];
        if (bias) v += bias[c + j];
        if (p.epilogue == EPI_SILU) v = silu_f(v);
        o[j] = v;
      }
      if (p.R) {
        const float4 r = *reinterpret_cast<const float4*>(p.R + rrow + c);
        o[0] += r.x; o[1] += r.y; o[2] += r.z; o[3] += r.w;
      }
      *reinterpret_cast<float4*>(p.C + (size_t)m * p.c_ld + c) = make_float4(o[0], o[1], o[2], o[3]);
    }
  }
}

}  // namespace

cudaError_t launch_gemm_f32(const GemmArgs& a, cudaStream_t s) {
  if (a.M <= 0 || a.N <= 0) return cudaSuccess;
  if (a.K % BK || a.cin % BK || a.N % 4 || a.a_ld % 4 || a.c_ld % 4 || (a.R && a.r_ld % 4) || a.taps * a.cin != a.K)
    return cudaErrorInvalidValue;
  if (a.epilogue == EPI_GEGLU && a.N % BN) return cudaErrorInvalidValue;
  dim3 grid((a.N + BN - 1) / BN, (a.M + BM - 1) / BM);
  gemm_f32_kernel<<<grid, NT, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace lds
