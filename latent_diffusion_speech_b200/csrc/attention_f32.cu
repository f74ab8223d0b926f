// attention_f32.cu — fp32 flash-style self-attention over the time axis (LDS_PREC_FP32 path).
//
// Replaces F.scaled_dot_product_attention at diffusion/unet1d/attention_processor.py:1032-1034
// (non-causal, no mask, dropout 0, scale 1/sqrt(d)); heads are contiguous d-wide channel slices
// (attention_processor.py:1025-1028).  Input is the fused QKV projection [B*T, 3C]; the score
// matrix is never materialised (T=2584 would need 27 GB at B=256).
//
// CTA = 64 queries of one (utterance, head); 256 threads as a 16x16 grid, each owning a 4x4
// block of the 64x64 score tile and a 4 x (d/16) block of the output tile.  Keys/values stream
// through shared memory in tiles of 64; online softmax (running max / sum) in registers with
// half-warp shuffles.  P is handed to the PV product through shared memory with a key
// permutation (row = (key%4)*16 + key/4) that makes both the P stores and the V/P loads
// bank-conflict free; V rows are stored with the same permutation so the product is unchanged.
#include "lds_kernels.h"
#include "planes.cuh"
#include <math.h>

namespace lds {
namespace {

constexpr int BQ = 64, BKV = 64, ATT_THREADS = 256, TLD = 68;  // TLD: padded leading dim of transposed tiles

template <int D>
__global__ void __launch_bounds__(ATT_THREADS) attention_f32_kernel(const float* __restrict__ qkv, float* __restrict__ out,
                                                                     __nv_bfloat16* __restrict__ outb, int parts, int T, int C) {
  constexpr int DC = D / 16;  // output columns per thread
  extern __shared__ __align__(16) float smem[];
  float* Qs = smem;                 // [D][TLD]   Qs[k][i]
  float* Ks = Qs + D * TLD;         // [D][TLD]   Ks[k][j]
  float* Vs = Ks + D * TLD;         // [BKV][D]   row = perm(key)
  float* Ps = Vs + BKV * D;         // [BKV][TLD] Ps[perm(key)][i]

  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * BQ;
  const int ld = 3 * C;
  const float* base = qkv + (size_t)b * T * ld + h * D;
  const float scale = rsqrtf((float)D);

  // ---- stage Q (transposed) ----
  constexpr int QV = D / 4;  // float4 per row
  for (int e = tid; e < BQ * QV; e += ATT_THREADS) {
    const int r = e / QV, kv = (e - r * QV) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < T) v = __ldg(reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * ld + kv));
    Qs[(kv + 0) * TLD + r] = v.x; Qs[(kv + 1) * TLD + r] = v.y;
    Qs[(kv + 2) * TLD + r] = v.z; Qs[(kv + 3) * TLD + r] = v.w;
  }

  float m_run[4], l_run[4], o[4][DC];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
    for (int c = 0; c < DC; ++c) o[i][c] = 0.f;
  }

  for (int k0 = 0; k0 < T; k0 += BKV) {
    __syncthreads();  // previous tile fully consumed (also orders the Q staging before first use)
    for (int e = tid; e < BKV * QV; e += ATT_THREADS) {
      const int r = e / QV, kv = (e - r * QV) * 4;
      float4 kq = make_float4(0.f, 0.f, 0.f, 0.f), vq = kq;
      if (k0 + r < T) {
        const float* row = base + (size_t)(k0 + r) * ld + kv;
        kq = __ldg(reinterpret_cast<const float4*>(row + C));
        vq = __ldg(reinterpret_cast<const float4*>(row + 2 * C));
      }
      Ks[(kv + 0) * TLD + r] = kq.x; Ks[(kv + 1) * TLD + r] = kq.y;
      Ks[(kv + 2) * TLD + r] = kq.z; Ks[(kv + 3) * TLD + r] = kq.w;
      const int pr = (r & 3) * 16 + (r >> 2);
      *reinterpret_cast<float4*>(&Vs[pr * D + kv]) = vq;
    }
    __syncthreads();

    // ---- S = Q K^T (4x4 per thread) ----
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
    for (int kk = 0; kk < D; ++kk) {
      const float4 a = *reinterpret_cast<const float4*>(&Qs[kk * TLD + ty * 4]);
      const float4 bq = *reinterpret_cast<const float4*>(&Ks[kk * TLD + tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {bq.x, bq.y, bq.z, bq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], bv[j], s[i][j]);
    }

    // ---- online softmax ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = (k0 + tx * 4 + j < T) ? s[i][j] * scale : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[i], mx);
      const float corr = expf(m_run[i] - m_new);  // exp(-inf) = 0 on the first tile
      float sum = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        s[i][j] = expf(s[i][j] - m_new);
        sum += s[i][j];
      }
#pragma unroll
      for (int off = 1; off < 16; off <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
      l_run[i] = l_run[i] * corr + sum;
      m_run[i] = m_new;
#pragma unroll
      for (int c = 0; c < DC; ++c) o[i][c] *= corr;
    }
    // P -> shared (key tx*4+j lives in row j*16+tx)
#pragma unroll
    for (int j = 0; j < 4; ++j)
      *reinterpret_cast<float4*>(&Ps[(j * 16 + tx) * TLD + ty * 4]) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
    __syncthreads();

    // ---- O += P V ----
#pragma unroll 8
    for (int j = 0; j < BKV; ++j) {
      const float4 p4 = *reinterpret_cast<const float4*>(&Ps[j * TLD + ty * 4]);
      const float pv[4] = {p4.x, p4.y, p4.z, p4.w};
      float vv[DC];
      if constexpr (DC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(&Vs[j * D + tx * 4]);
        vv[0] = t.x; vv[1] = t.y; vv[2] = t.z; vv[3] = t.w;
      } else if constexpr (DC == 2) {
        const float2 t = *reinterpret_cast<const float2*>(&Vs[j * D + tx * 2]);
        vv[0] = t.x; vv[1] = t.y;
      } else {
#pragma unroll
        for (int c = 0; c < DC; ++c) vv[c] = Vs[j * D + tx * DC + c];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int c = 0; c < DC; ++c) o[i][c] = fmaf(pv[i], vv[c], o[i][c]);
    }
  }

  // ---- normalise and store ----
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int q = q0 + ty * 4 + i;
    if (q >= T) continue;
    const float inv = 1.f / l_run[i];
    if (outb) {
      __nv_bfloat16* rowb = outb + ((size_t)b * T + q) * (size_t)(parts * C);
#pragma unroll
      for (int c = 0; c < DC; ++c) {
        float r = o[i][c] * inv;
        for (int pl = 0; pl < parts; ++pl) {
          const __nv_bfloat16 hv = __float2bfloat16_rn(r);
          r -= __bfloat162float(hv);
          rowb[(size_t)pl * C + h * D + tx * DC + c] = hv;
        }
      }
    } else {
      float* dst = out + ((size_t)b * T + q) * C + h * D + tx * DC;
#pragma unroll
      for (int c = 0; c < DC; ++c) dst[c] = o[i][c] * inv;
    }
  }
}

template <int D>
cudaError_t launch_d(const float* qkv, float* out, __nv_bfloat16* outb, int parts, int B, int T, int C, int heads, cudaStream_t s) {
  const size_t smem = (size_t)(2 * D * TLD + BKV * D + BKV * TLD) * sizeof(float);
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(attention_f32_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  dim3 grid((T + BQ - 1) / BQ, heads, B);
  attention_f32_kernel<D><<<grid, ATT_THREADS, smem, s>>>(qkv, out, outb, parts, T, C);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_f32(const float* qkv, float* out, __nv_bfloat16* outb, int parts, int B, int T, int C, int heads,
                                 cudaStream_t s) {
  if (heads <= 0 || C % heads) return cudaErrorInvalidValue;
  switch (C / heads) {
    case 32: return launch_d<32>(qkv, out, outb, parts, B, T, C, heads, s);
    case 48: return launch_d<48>(qkv, out, outb, parts, B, T, C, heads, s);
    case 64: return launch_d<64>(qkv, out, outb, parts, B, T, C, heads, s);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace lds
