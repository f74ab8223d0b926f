// norm.cu — GroupNorm (statistics + fused affine / scale-shift / SiLU apply) and LayerNorm on
// channels-last activations.  HBM-bound kernels: float4 accesses, warp-shuffle / shared-memory
// reductions, Chan/Welford merging of (count, mean, M2) so that variance never suffers the
// E[x^2]-E[x]^2 cancellation (activations reach |x| ~ 1e2 with random-init weights).
//
// Reference ops replaced:
//   nn.GroupNorm(8, C, eps)                resnet.py:536,557 ; transformer_1d.py:134 ; unet_1d_condition.py:546
//   h*(1+scale)+shift, SiLU                resnet.py:597-631
//   torch.cat([hidden, skip], dim=1)       unet_1d_blocks.py:2084,2186  (virtual: two source pointers)
//   nn.LayerNorm(C)                        attention.py:83,102,118
#include "lds_kernels.h"
#include "planes.cuh"

namespace lds {
namespace {

struct Wf { float n, mean, m2; };

__device__ __forceinline__ Wf wf_merge(Wf a, Wf b) {
  if (b.n == 0.f) return a;
  if (a.n == 0.f) return b;
  const float n = a.n + b.n;
  const float d = b.mean - a.mean;
  const float f = b.n / n;
  Wf r;
  r.n = n;
  r.mean = a.mean + d * f;
  r.m2 = a.m2 + b.m2 + d * d * a.n * f;
  return r;
}


__device__ __forceinline__ float4 load_cat(const float* x1, int c1, const float* x2, int c2, size_t row, int c) {
  // channel c of the virtual concat [x1 | x2]; c and c1 are multiples of 4
  return (c < c1) ? __ldg(reinterpret_cast<const float4*>(x1 + row * c1 + c))
                  : __ldg(reinterpret_cast<const float4*>(x2 + row * c2 + (c - c1)));
}

// Merge of partial (count, mean, M2) triples held by the lanes of one warp (lane-strided entries already folded into
// n / nm / per-entry arrays by the caller): two shuffle reductions, no division inside the loops.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  return v;
}

// grid (nchunk, B); block = (C/4) * rpar threads, thread (v, r) owns channel quad v and frames r, r+rpar, ...
// Phase 1: every thread reduces its own values in registers (exact two-pass mean / M2 over <= GN_ROWS/rpar float4s).
// Phase 2: warp g merges the per-thread triples of group g with the count-weighted Chan formula.
constexpr int GN_UNROLL = 8;
__global__ void gn_stats_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2, int T,
                                int groups, int rpar, int chunk_rows, float* __restrict__ part) {
  extern __shared__ float sh[];  // [nthreads][3]
  pdl_trigger();
  pdl_wait();
  const int C = c1 + c2, V = C >> 2, cg = C / groups;
  const int v = threadIdx.x % V, r0 = threadIdx.x / V;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int t0 = chunk * chunk_rows, t1 = min(t0 + chunk_rows, T);
  Wf acc{0.f, 0.f, 0.f};
  if (r0 < rpar) {
    for (int tb = t0 + r0; tb < t1; tb += rpar * GN_UNROLL) {
      float4 x[GN_UNROLL];
      int cnt = 0;
#pragma unroll
      for (int i = 0; i < GN_UNROLL; ++i) {
        const int t = tb + i * rpar;
        if (t < t1) { x[i] = load_cat(x1, c1, x2, c2, (size_t)b * T + t, v * 4); ++cnt; }
      }
      float sum = 0.f;
#pragma unroll
      for (int i = 0; i < GN_UNROLL; ++i)
        if (i < cnt) sum += (x[i].x + x[i].y) + (x[i].z + x[i].w);
      const float n = 4.f * cnt, mean = sum / n;
      float m2 = 0.f;
#pragma unroll
      for (int i = 0; i < GN_UNROLL; ++i)
        if (i < cnt) {
          const float d0 = x[i].x - mean, d1 = x[i].y - mean, d2 = x[i].z - mean, d3 = x[i].w - mean;
          m2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
        }
      acc = wf_merge(acc, Wf{n, mean, m2});
    }
  }
  sh[threadIdx.x * 3 + 0] = acc.n; sh[threadIdx.x * 3 + 1] = acc.mean; sh[threadIdx.x * 3 + 2] = acc.m2;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  const int vq = cg >> 2, entries = rpar * vq;   // quads per group; per-thread triples of one group
  for (int g = warp; g < groups; g += nwarps) {
    float n = 0.f, nm = 0.f;
    for (int e = lane; e < entries; e += 32) {
      const int th = (e / vq) * V + g * vq + (e % vq);
      n += sh[th * 3];
      nm += sh[th * 3] * sh[th * 3 + 1];
    }
    n = warp_sum(n);
    nm = warp_sum(nm);
    const float mean = n > 0.f ? nm / n : 0.f;
    float m2 = 0.f;
    for (int e = lane; e < entries; e += 32) {
      const int th = (e / vq) * V + g * vq + (e % vq);
      const float d = sh[th * 3 + 1] - mean;
      m2 += sh[th * 3 + 2] + sh[th * 3] * d * d;
    }
    m2 = warp_sum(m2);
    if (lane == 0) {
      float* dst = part + (((size_t)b * gridDim.x + chunk) * groups + g) * 3;
      dst[0] = n; dst[1] = mean; dst[2] = m2;
    }
  }
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
// bf16 mode only (the result is rounded to bf16): ex2.approx + rcp.approx, ~4 ulp, a third of the instructions
__device__ __forceinline__ float silu_fast(float x) { return __fdividef(x, 1.f + __expf(-x)); }

// grid (ceil(T / slab), B); block = (C/4) * rpar threads: thread (v, r) owns channel quad v — its gamma / beta / scale /
// shift live in registers — and frames r, r+rpar, ... of the [slab, C] slab, eight 16-byte loads in flight per thread
// (bytes in flight per SM, not instruction issue, is what bounds this kernel).  The host picks slab = 32 frames when
// that still yields several waves of blocks and 16 on the small U-Net levels.
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2,
                                                        int c2, int T, int groups, int rpar, int nchunk, int slab, const float* __restrict__ part,
                                                        float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ ss,
                                                        int silu, float* __restrict__ y, __nv_bfloat16* __restrict__ yb,
                                                        int parts, __nv_bfloat16* __restrict__ rawb) {
  __shared__ float s_mean[32], s_rstd[32];
  pdl_trigger();
  pdl_wait();
  const int C = c1 + c2, V = C >> 2, cg = C / groups;
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int g = warp; g < groups; g += nwarps) {   // merge the per-chunk partials (nchunk statistics chunks) of group g
    const float* src = part + ((size_t)b * nchunk * groups + g) * 3;
    float n = 0.f, nm = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      const float* e = src + (size_t)c * groups * 3;
      n += e[0];
      nm += e[0] * e[1];
    }
    n = warp_sum(n);
    nm = warp_sum(nm);
    const float mean = nm / n;
    float m2 = 0.f;
    for (int c = lane; c < nchunk; c += 32) {
      const float* e = src + (size_t)c * groups * 3;
      const float d = e[1] - mean;
      m2 += e[2] + e[0] * d * d;
    }
    m2 = warp_sum(m2);
    if (lane == 0) { s_mean[g] = mean; s_rstd[g] = rsqrtf(m2 / n + eps); }
  }
  __syncthreads();
  const int v = threadIdx.x % V, r0 = threadIdx.x / V;
  if (r0 >= rpar) return;
  const int c = v * 4, g = c / cg;
  const float mean = s_mean[g], rstd = s_rstd[g];
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sf = sc;
  if (ss) {
    sc = __ldg(reinterpret_cast<const float4*>(ss + c));
    sf = __ldg(reinterpret_cast<const float4*>(ss + C + c));
  }
  const int t0 = chunk * slab, t1 = min(t0 + slab, T);
  constexpr int U = 8;
  for (int tb = t0 + r0; tb < t1; tb += rpar * U) {
    float4 xv[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int t = tb + i * rpar;
      if (t < t1) xv[i] = load_cat(x1, c1, x2, c2, (size_t)b * T + t, c);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int t = tb + i * rpar;
      if (t >= t1) break;
      const size_t row = (size_t)b * T + t;
      float o[4] = {(xv[i].x - mean) * rstd * ga.x + be.x, (xv[i].y - mean) * rstd * ga.y + be.y,
                    (xv[i].z - mean) * rstd * ga.z + be.z, (xv[i].w - mean) * rstd * ga.w + be.w};
      if (ss) {
        o[0] = o[0] * (1.f + sc.x) + sf.x; o[1] = o[1] * (1.f + sc.y) + sf.y;
        o[2] = o[2] * (1.f + sc.z) + sf.z; o[3] = o[3] * (1.f + sc.w) + sf.w;
      }
      if (silu) { o[0] = silu_f(o[0]); o[1] = silu_f(o[1]); o[2] = silu_f(o[2]); o[3] = silu_f(o[3]); }
      if (yb) store_planes4(yb + row * (size_t)(parts * C), c, C, parts, o[0], o[1], o[2], o[3]);
      else *reinterpret_cast<float4*>(y + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
      if (rawb) store_planes4(rawb + row * (size_t)(parts * C), c, C, parts, xv[i].x, xv[i].y, xv[i].z, xv[i].w);
    }
  }
}

// Single-pass GroupNorm: one CTA per (utterance, group) keeps its whole [T, C/G] slab in shared memory, so x is read
// from HBM once (statistics + apply: 4 B read + 2*parts B written per element instead of 8 + 2*parts) and the two
// launches become one.  Exact two-pass mean / M2 over the slab.  Used when the slab fits (T * C/G * 4 B <= ~200 KB),
// else the stats + apply pair above.  thread (v, r): channel quad v of the group, frames r, r+R, ...
constexpr int GNF_THREADS = 512;
__global__ void __launch_bounds__(GNF_THREADS) gn_fused_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2,
                                                               int T, int groups, float eps, const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, const float* __restrict__ ss, int silu,
                                                               float* __restrict__ y, __nv_bfloat16* __restrict__ yb, int parts,
                                                               __nv_bfloat16* __restrict__ rawb) {
  extern __shared__ float4 slab[];               // [T][q]
  __shared__ float red[GNF_THREADS / 32];
  __shared__ float s_bcast;
  pdl_trigger();
  pdl_wait();
  const int C = c1 + c2, cg = C / groups, q = cg >> 2;
  const int g = blockIdx.x, b = blockIdx.y;
  const int R = GNF_THREADS / q;                 // frames in flight per pass over the block's threads
  const int v = threadIdx.x % q, r0 = threadIdx.x / q;
  const int c = g * cg + 4 * v;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool active = r0 < R;
  auto block_sum = [&](float val) {
    val = warp_sum(val);
    if (lane == 0) red[warp] = val;
    __syncthreads();
    if (warp == 0) {
      float t = lane < GNF_THREADS / 32 ? red[lane] : 0.f;
      t = warp_sum(t);
      if (lane == 0) s_bcast = t;
    }
    __syncthreads();
    return s_bcast;
  };
  float sum = 0.f;
  if (active) {
    constexpr int U = 8;                         // loads in flight per thread
    for (int tb = r0; tb < T; tb += R * U) {
      float4 xv[U];
#pragma unroll
      for (int i = 0; i < U; ++i)
        if (tb + i * R < T) xv[i] = load_cat(x1, c1, x2, c2, (size_t)b * T + tb + i * R, c);
#pragma unroll
      for (int i = 0; i < U; ++i)
        if (tb + i * R < T) {
          slab[(tb + i * R) * q + v] = xv[i];
          sum += (xv[i].x + xv[i].y) + (xv[i].z + xv[i].w);
        }
    }
  }
  const float n = (float)T * (float)cg;
  const float mean = block_sum(sum) / n;
  float m2 = 0.f;
  if (active) {
    for (int t = r0; t < T; t += R) {
      const float4 xv = slab[t * q + v];
      const float d0 = xv.x - mean, d1 = xv.y - mean, d2 = xv.z - mean, d3 = xv.w - mean;
      m2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
  }
  const float rstd = rsqrtf(block_sum(m2) / n + eps);
  if (!active) return;
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma + c));
  const float4 be = __ldg(reinterpret_cast<const float4*>(beta + c));
  float4 sc = make_float4(0.f, 0.f, 0.f, 0.f), sf = sc;
  if (ss) {
    sc = __ldg(reinterpret_cast<const float4*>(ss + c));
    sf = __ldg(reinterpret_cast<const float4*>(ss + C + c));
  }
  for (int t = r0; t < T; t += R) {
    const float4 xv = slab[t * q + v];
    const size_t row = (size_t)b * T + t;
    float o[4] = {(xv.x - mean) * rstd * ga.x + be.x, (xv.y - mean) * rstd * ga.y + be.y,
                  (xv.z - mean) * rstd * ga.z + be.z, (xv.w - mean) * rstd * ga.w + be.w};
    if (ss) {
      o[0] = o[0] * (1.f + sc.x) + sf.x; o[1] = o[1] * (1.f + sc.y) + sf.y;
      o[2] = o[2] * (1.f + sc.z) + sf.z; o[3] = o[3] * (1.f + sc.w) + sf.w;
    }
    if (silu) {
      if (yb && parts == 1) { o[0] = silu_fast(o[0]); o[1] = silu_fast(o[1]); o[2] = silu_fast(o[2]); o[3] = silu_fast(o[3]); }
      else { o[0] = silu_f(o[0]); o[1] = silu_f(o[1]); o[2] = silu_f(o[2]); o[3] = silu_f(o[3]); }
    }
    if (yb) store_planes4(yb + row * (size_t)(parts * C), c, C, parts, o[0], o[1], o[2], o[3]);
    else *reinterpret_cast<float4*>(y + row * C + c) = make_float4(o[0], o[1], o[2], o[3]);
    if (rawb) store_planes4(rawb + row * (size_t)(parts * C), c, C, parts, xv.x, xv.y, xv.z, xv.w);
  }
}

// One warp per LN_ROWS consecutive rows (C <= 32*4*LN_MAXV, LN_MAXV instantiated for C <= 256 / 384 / 512): all loads of the rows are issued before the first
// reduction, so that LN_ROWS x C x 4 bytes per warp are in flight (bytes in flight, not issue, bound this kernel).
constexpr int LN_ROWS = 4;
template <int LN_MAXV>
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float eps, int rows, int C,
                                                        float* __restrict__ y, __nv_bfloat16* __restrict__ yb, int parts) {
  pdl_trigger();
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * LN_ROWS;
  if (row0 >= rows) return;
  const int V = C >> 2;
  float4 v[LN_ROWS][LN_MAXV];
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    const float4* src = reinterpret_cast<const float4*>(x + (size_t)(row0 + r) * C);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < V && row0 + r < rows) v[r][i] = __ldg(src + q);
    }
  }
  float4 g[LN_MAXV], be[LN_MAXV];
#pragma unroll
  for (int i = 0; i < LN_MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < V) {
      g[i] = __ldg(reinterpret_cast<const float4*>(gamma) + q);
      be[i] = __ldg(reinterpret_cast<const float4*>(beta) + q);
    }
  }
#pragma unroll
  for (int r = 0; r < LN_ROWS; ++r) {
    if (row0 + r >= rows) break;
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (lane + 32 * i < V) sum += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
    sum = warp_sum(sum);
    const float mean = sum / (float)C;
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i)
      if (lane + 32 * i < V) {
        const float a = v[r][i].x - mean, b = v[r][i].y - mean, c = v[r][i].z - mean, d = v[r][i].w - mean;
        sq += (a * a + b * b) + (c * c + d * d);
      }
    sq = warp_sum(sq);
    const float rstd = rsqrtf(sq / (float)C + eps);
    const size_t row = (size_t)(row0 + r);
#pragma unroll
    for (int i = 0; i < LN_MAXV; ++i) {
      const int q = lane + 32 * i;
      if (q < V) {
        const float4 o = make_float4((v[r][i].x - mean) * rstd * g[i].x + be[i].x, (v[r][i].y - mean) * rstd * g[i].y + be[i].y,
                                     (v[r][i].z - mean) * rstd * g[i].z + be[i].z, (v[r][i].w - mean) * rstd * g[i].w + be[i].w);
        if (yb) store_planes4(yb + row * (parts * C), q * 4, C, parts, o.x, o.y, o.z, o.w);
        else reinterpret_cast<float4*>(y + row * C)[q] = o;
      }
    }
  }
}

}  // namespace

// Frames per statistics chunk: 64 when that still gives several waves of blocks (the merge phase of a block is then
// amortised over twice the loads), GN_ROWS = 32 on the small U-Net levels.  `part` is sized for 32-frame chunks.
static int gn_chunk_rows(int B, int T) { return (int64_t)((T + 63) / 64) * B >= 4 * 148 ? 64 : GN_ROWS; }

cudaError_t launch_gn_stats(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float* part,
                            cudaStream_t s) {
  const int C = c1 + c2;
  if (C % (4 * groups) || c1 % 4 || groups > 32 || C / 4 > 1024) return cudaErrorInvalidValue;
  const int V = C / 4;
  int rpar = 256 / V;
  if (rpar < 1) rpar = 1;
  if (rpar > GN_ROWS) rpar = GN_ROWS;
  const int threads = ((V * rpar + 31) / 32) * 32;
  const int chunk_rows = gn_chunk_rows(B, T);
  dim3 grid((T + chunk_rows - 1) / chunk_rows, B);
  return launch_pdl(gn_stats_kernel, grid, dim3(threads), threads * 3 * sizeof(float), s, 1, x1, c1, x2, c2, T, groups, rpar, chunk_rows, part);
}

cudaError_t launch_gn_apply(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups,
                            const float* part, float eps, const float* gamma, const float* beta, const float* ss,
                            int silu, float* y, __nv_bfloat16* yb, int parts, __nv_bfloat16* rawb, cudaStream_t s) {
  const int C = c1 + c2;
  if (C % (4 * groups) || c1 % 4 || groups > 32 || C / 4 > 256) return cudaErrorInvalidValue;
  const int V = C / 4;
  int rpar = 256 / V;
  if (rpar < 1) rpar = 1;
  const int slab = (int64_t)((T + 31) / 32) * B >= 4 * 148 ? 32 : 16;
  if (rpar > slab) rpar = slab;
  const int threads = ((V * rpar + 31) / 32) * 32;
  dim3 grid((T + slab - 1) / slab, B);
  const int chunk_rows = gn_chunk_rows(B, T);
  return launch_pdl(gn_apply_kernel, grid, dim3(threads), 0, s, 1, x1, c1, x2, c2, T, groups, rpar, (T + chunk_rows - 1) / chunk_rows, slab, part, eps, gamma, beta, ss,
                                           silu, y, yb, parts, rawb);
}

// cudaErrorNotSupported when the slab of one (utterance, group) does not fit shared memory: use stats + apply then.
cudaError_t launch_gn_fused(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                            const float* gamma, const float* beta, const float* ss, int silu, float* y, __nv_bfloat16* yb,
                            int parts, __nv_bfloat16* rawb, cudaStream_t s) {
  const int C = c1 + c2;
  if (C % (4 * groups) || c1 % 4 || groups > 32) return cudaErrorInvalidValue;
  const int cg = C / groups, q = cg / 4;
  const size_t smem = (size_t)T * cg * sizeof(float);
  if (q > GNF_THREADS || smem > 200 * 1024) return cudaErrorNotSupported;
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(gn_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(gn_fused_kernel, dim3(groups, B), dim3(GNF_THREADS), smem, s, 1, x1, c1, x2, c2, T, groups, eps, gamma, beta, ss,
                    silu, y, yb, parts, rawb);
}

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int C, float* y,
                             __nv_bfloat16* yb, int parts, cudaStream_t s) {
  if (C % 4 || C > 512) return cudaErrorInvalidValue;
  const int warps_per_block = 8, rows_per_block = warps_per_block * LN_ROWS;
  const int grid = (rows + rows_per_block - 1) / rows_per_block;
  const dim3 g(grid), b(warps_per_block * 32);
  if (C <= 256) return launch_pdl(layernorm_kernel<2>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
  if (C <= 384) return launch_pdl(layernorm_kernel<3>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
  return launch_pdl(layernorm_kernel<4>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
}

}  // namespace lds
