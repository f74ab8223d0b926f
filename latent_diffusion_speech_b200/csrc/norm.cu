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
#include <algorithm>

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


// branch-free variant (empty partials allowed), approximate reciprocal: f only weights the mean correction
__device__ __forceinline__ Wf wf_merge_fast(Wf a, Wf b) {
  const float n = a.n + b.n;
  const float f = n > 0.f ? __fdividef(b.n, n) : 0.f;
  const float d = b.mean - a.mean;
  Wf r;
  r.n = n;
  r.mean = fmaf(d, f, a.mean);
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
                                                        int parts, __nv_bfloat16* __restrict__ rawb, long long ss_b) {
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
    sc = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + c));
    sf = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + C + c));
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
                                                               __nv_bfloat16* __restrict__ rawb, long long ss_b) {
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
    sc = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + c));
    sf = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + C + c));
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

// Cluster form of the single-pass GroupNorm: the [T, C/G] slab of one (utterance, group) is split along time over the
// CL CTAs of a thread-block cluster (slices of <= 36 KB), so that several small CTAs share an SM and slabs that do not
// fit one SM (640-channel concats, T = 2584 at 256 channels) stay single-pass.  The per-CTA (count, mean, M2) partials
// travel through distributed shared memory: every CTA reads the CL partials in rank order (identical mean / rstd in
// all of them).  CL depends on (T, C/G) only, never on the batch: results are batch-composition invariant.
// History of the kernel and the ncu evidence behind each step: profiles/r01_groupnorm_layernorm_investigation.md.
constexpr int GNC_THREADS = 128;   // 13.5 instead of 6.75 channel quads per thread and item at T = 864: the per-item fixed cost halves
__device__ __forceinline__ uint32_t gnc_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t gnc_nctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
// Relaxed arrive: the one thread that published data fences first (fence + relaxed arrive = release); a
// barrier.cluster.arrive.release in every thread costs an ERRBAR / membar per warp (ncu: 9 % of the stall samples).
__device__ __forceinline__ void gnc_cluster_arrive() { asm volatile("barrier.cluster.arrive.relaxed.aligned;" ::: "memory"); }
__device__ __forceinline__ void gnc_cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void gnc_publish_fence() { asm volatile("fence.acq_rel.cluster;" ::: "memory"); }
__device__ __forceinline__ float gnc_ld_peer(const float* p, uint32_t rank) {   // *p of CTA `rank` of this cluster
  uint32_t a = (uint32_t)__cvta_generic_to_shared(p), ra;
  float v;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(rank));
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}
// x * sigmoid(x) with the reciprocal from rcp.approx + one Newton step (<= 1 ulp) instead of the IEEE division and its
// slow-path branch; x is clamped at -80 so that 1 + e^-x stays finite (silu(-80) ~ -1e-33 either way).
__device__ __forceinline__ float silu_newton(float x) {
  const float t = 1.f + expf(-fmaxf(x, -80.f));
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(t));
  r = fmaf(r, fmaf(-t, r, 1.f), r);
  return x * r;
}
template <int PARTS>
__device__ __forceinline__ void store_planes4_ct(uint2* dst, int plane_stride_u2, float v0, float v1, float v2, float v3) {
  if constexpr (PARTS == 2) {                    // split-f16: two fp16 planes of the scaled value (planes.cuh)
    v0 *= PLANE_SCALE; v1 *= PLANE_SCALE; v2 *= PLANE_SCALE; v3 *= PLANE_SCALE;
    uint2 w;
    w.x = planes_split_pair_f16(v0, v1);
    w.y = planes_split_pair_f16(v2, v3);
    dst[0] = w;
    w.x = planes_pack_pair_f16(v0, v1);
    w.y = planes_pack_pair_f16(v2, v3);
    dst[plane_stride_u2] = w;
  } else {
#pragma unroll
    for (int p = 0; p < PARTS; ++p) {
      uint2 w;
      if (p == PARTS - 1) {
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.x) : "f"(v1), "f"(v0));
        asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w.y) : "f"(v3), "f"(v2));
      } else {
        w.x = planes_split_pair(v0, v1);
        w.y = planes_split_pair(v2, v3);
      }
      dst[p * plane_stride_u2] = w;
    }
  }
}
__device__ __forceinline__ void cp_async16(uint32_t dst_smem, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_smem), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
// PARTS: 0 fp32 output y, 1 / 3 bf16 operand planes yb.  SILU / SS (scale-shift) / RAW (also copy x as planes) are
// compile-time so that the loops are branch-free; all addresses advance by constant strides (the first version spent
// 3/4 of its instructions on index arithmetic and flag tests: 58 instructions per element).
// Persistent and double-buffered: cluster k handles the (utterance, group) items k, k + n_clusters, ...; the slab of the
// next item streams into the other shared-memory buffer with cp.async (no registers) while the current one is reduced,
// normalised and stored — without that, the CTAs of an SM run their load / barrier / store phases in lockstep and
// HBM idles during the barriers (measured 3.7 TB/s effective whatever the instruction count).
template <int PARTS, bool SILU, bool SS, bool RAW>
__global__ void __launch_bounds__(GNC_THREADS, 8) gn_cluster_kernel(const float* __restrict__ x1, int c1, const float* __restrict__ x2, int c2,
                                                                    int T, int groups, int n_items, int tc, float eps,
                                                                    const float* __restrict__ gamma, const float* __restrict__ beta,
                                                                    const float* __restrict__ ss, float* __restrict__ y,
                                                                    __nv_bfloat16* __restrict__ yb, __nv_bfloat16* __restrict__ rawb, long long ss_b) {
  extern __shared__ float4 slab[];               // [2][tc][q]: frames [rank*tc, rank*tc + tc) of the group, two items
  __shared__ float red[GNC_THREADS / 32][3];     // per-warp (count, mean, M2)
  __shared__ float s_part[2][3];                 // this CTA's (count, mean, M2), read by its cluster peers; per item parity
  pdl_trigger();
  const uint32_t CL = gnc_nctarank(), rank = gnc_ctarank();
  const int C = c1 + c2, cg = C / groups, q = cg >> 2;
  const int cluster_id = blockIdx.x / CL, n_clusters = gridDim.x / CL;
  const int t_lo = min(T, (int)rank * tc), nt = min(T, t_lo + tc) - t_lo;
  const int R = GNC_THREADS / q;                 // frames per pass over the block's threads
  const int v = threadIdx.x % q, r0 = threadIdx.x / q;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool active = r0 < R;
  const int sstep = R * q;                       // slab step (float4) between two frames of this thread
  const int buf_f4 = tc * q;                     // one slab buffer in float4
  const uint32_t slab_u32 = (uint32_t)__cvta_generic_to_shared(slab);
  // streams the frames of item `item` owned by this thread into buffer `buf`
  const uint32_t dst0 = slab_u32 + (uint32_t)(r0 * q + v) * 16u, dstep = (uint32_t)sstep * 16u;
  const int nfr = (active && r0 < nt) ? (nt - r0 + R - 1) / R : 0;   // frames of this thread in every item
  auto prefetch = [&](int item, int buf) {
    if (nfr > 0) {
      const int g = item % groups, b = item / groups;
      const int c = g * cg + 4 * v;
      // the virtual concat [x1 | x2] resolves per thread (its channel quad is fixed), not per load
      const float* src = c < c1 ? x1 + c : x2 + (c - c1);
      const int ld = c < c1 ? c1 : c2;
      const char* p = reinterpret_cast<const char*>(src + ((size_t)b * T + t_lo + r0) * (size_t)ld);
      const size_t pstep = (size_t)R * ld * sizeof(float);
      uint32_t dst = dst0 + (uint32_t)(buf * buf_f4) * 16u;
#pragma unroll 4
      for (int i = 0; i < nfr; ++i, p += pstep, dst += dstep) cp_async16(dst, p);
    }
    cp_async_commit();
  };
  const float n = (float)T * (float)cg;
  constexpr int PW = PARTS > 0 ? PARTS : 1;
  const int plane_u2 = C >> 2;                   // one plane of a row in uint2 (4 x bf16) units
  const int orow = PARTS > 0 ? PW * (C >> 2) : (C >> 2);   // one output row in uint2 (planes) / float4 (fp32) units
  const size_t ostep = (size_t)R * orow;
  pdl_wait();
  if (cluster_id < n_items) prefetch(cluster_id, 0);
  int buf = 0;
  for (int item = cluster_id; item < n_items; item += n_clusters, buf ^= 1) {
    const int g = item % groups, b = item / groups;
    const int c = g * cg + 4 * v;
    cp_async_wait_all();
    __syncthreads();                             // slab[buf] complete and visible; every thread is done with slab[buf ^ 1]
    if (item + n_clusters < n_items) prefetch(item + n_clusters, buf ^ 1);
    // coefficients of this thread's channel quad: requested now, needed after the two reductions
    float4 A = make_float4(0.f, 0.f, 0.f, 0.f), Bc = A, sc = A, sf = A;
    if (active) {
      A = __ldg(reinterpret_cast<const float4*>(gamma + c));
      Bc = __ldg(reinterpret_cast<const float4*>(beta + c));
      if (SS) {
        sc = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + c));
        sf = __ldg(reinterpret_cast<const float4*>(ss + (size_t)b * ss_b + C + c));
      }
    }
    const float4* sp0 = slab + buf * buf_f4 + r0 * q + v;
    // One exchange per item.  Each warp takes the mean of its lanes' first channel quads as a pivot k (within
    // sigma/sqrt(128) of the mean: the shifted moments below lose no bits), accumulates sum(x-k) and sum((x-k)^2) in ONE
    // pass over shared memory, and turns them into an exact-enough (count, mean, M2) triple; the 4 warp triples and then
    // the CL CTA triples merge with the count-weighted Chan formula (one shared-memory hop, one cluster hop).  The
    // first versions (two block sums + two cluster barriers, then a per-thread Chan tree) spent more instructions on
    // the reductions than on the normalisation itself: 44 instructions per element, issue-bound (ncu).
    float s1 = 0.f, s2 = 0.f, k = 0.f;
    {
      float pv = 0.f;
      if (nfr > 0) { const float4 x0 = *sp0; pv = (x0.x + x0.y) + (x0.z + x0.w); }
      const float cntl = nfr > 0 ? 4.f : 0.f;
      const float pvs = warp_sum(pv), cnts = warp_sum(cntl);
      k = cnts > 0.f ? pvs / cnts : 0.f;
    }
    {
      const float4* sp = sp0;
#pragma unroll 4
      for (int i = 0; i < nfr; ++i, sp += sstep) {
        const float4 xv = *sp;
        const float d0 = xv.x - k, d1 = xv.y - k, d2 = xv.z - k, d3 = xv.w - k;
        s1 += (d0 + d1) + (d2 + d3);
        s2 += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
    }
    float nw = 4.f * (float)nfr;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s1 += __shfl_xor_sync(0xffffffffu, s1, off);
      s2 += __shfl_xor_sync(0xffffffffu, s2, off);
      nw += __shfl_xor_sync(0xffffffffu, nw, off);
    }
    if (lane == 0) {
      const float dm = nw > 0.f ? s1 / nw : 0.f;           // mean - k of this warp's elements
      red[warp][0] = nw; red[warp][1] = k + dm; red[warp][2] = fmaxf(s2 - s1 * dm, 0.f);
    }
    __syncthreads();
    const int par = buf;                         // s_part slot alternates per item (a fast peer may already publish the next item)
    if (warp == 0) {
      Wf w{0.f, 0.f, 0.f};
      if (lane < GNC_THREADS / 32) { w.n = red[lane][0]; w.mean = red[lane][1]; w.m2 = red[lane][2]; }
#pragma unroll
      for (int off = GNC_THREADS / 64; off > 0; off >>= 1) {
        Wf o;
        o.n = __shfl_xor_sync(0xffffffffu, w.n, off);
        o.mean = __shfl_xor_sync(0xffffffffu, w.mean, off);
        o.m2 = __shfl_xor_sync(0xffffffffu, w.m2, off);
        w = wf_merge_fast(w, o);
      }
      if (lane == 0) {
        s_part[par][0] = w.n; s_part[par][1] = w.mean; s_part[par][2] = w.m2;
        gnc_publish_fence();
      }
    }
    gnc_cluster_arrive();
    gnc_cluster_wait();
    Wf tot{0.f, 0.f, 0.f};
    for (uint32_t r = 0; r < CL; ++r) {          // rank order: identical statistics in every CTA of the cluster
      Wf o;
      o.n = gnc_ld_peer(&s_part[par][0], r);
      o.mean = gnc_ld_peer(&s_part[par][1], r);
      o.m2 = gnc_ld_peer(&s_part[par][2], r);
      tot = wf_merge_fast(tot, o);
    }
    const float mean = tot.mean;
    const float rstd = rsqrtf(tot.m2 / n + eps);
    if (active) {
      // o = (x - mean) * A + Bc with A = rstd*gamma [*(1+scale)], Bc = beta [*(1+scale) + shift]
      A.x *= rstd; A.y *= rstd; A.z *= rstd; A.w *= rstd;
      if (SS) {
        const float k0 = 1.f + sc.x, k1 = 1.f + sc.y, k2 = 1.f + sc.z, k3 = 1.f + sc.w;
        A.x *= k0; A.y *= k1; A.z *= k2; A.w *= k3;
        Bc.x = fmaf(Bc.x, k0, sf.x); Bc.y = fmaf(Bc.y, k1, sf.y); Bc.z = fmaf(Bc.z, k2, sf.z); Bc.w = fmaf(Bc.w, k3, sf.w);
      }
      const float4* sp = sp0;
      const size_t o0 = ((size_t)b * T + t_lo + r0) * (size_t)orow + (c >> 2);
      uint2* ybp = PARTS > 0 ? reinterpret_cast<uint2*>(yb) + o0 : nullptr;
      uint2* rbp = (PARTS > 0 && RAW) ? reinterpret_cast<uint2*>(rawb) + o0 : nullptr;
      float4* yp = PARTS == 0 ? reinterpret_cast<float4*>(y) + o0 : nullptr;
#pragma unroll 2
      for (int i = 0; i < nfr; ++i, sp += sstep) {
        const float4 xv = *sp;
        float o0v = fmaf(xv.x - mean, A.x, Bc.x), o1v = fmaf(xv.y - mean, A.y, Bc.y), o2v = fmaf(xv.z - mean, A.z, Bc.z),
              o3v = fmaf(xv.w - mean, A.w, Bc.w);
        if (SILU) {
          if (PARTS == 1) { o0v = silu_fast(o0v); o1v = silu_fast(o1v); o2v = silu_fast(o2v); o3v = silu_fast(o3v); }
          else { o0v = silu_newton(o0v); o1v = silu_newton(o1v); o2v = silu_newton(o2v); o3v = silu_newton(o3v); }
        }
        if (PARTS > 0) {
          store_planes4_ct<PW>(ybp, plane_u2, o0v, o1v, o2v, o3v);
          ybp += ostep;
          if (RAW) {
            store_planes4_ct<PW>(rbp, plane_u2, xv.x, xv.y, xv.z, xv.w);
            rbp += ostep;
          }
        } else {
          *yp = make_float4(o0v, o1v, o2v, o3v);
          yp += ostep;
        }
      }
    }
  }
  gnc_cluster_arrive();                          // no CTA of the cluster exits while a peer may still read its s_part
  gnc_cluster_wait();
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


// LayerNorm fast path for C == 128 * NV (the denoiser's 256 / 384 / 512): one warp per ROWS consecutive rows, every
// load of the rows issued before the first reduction, the reductions of the rows interleaved (independent shuffle
// chains), gamma / beta staged in shared memory before the dependency wait (they are weights), compile-time plane count.
// ~12 instructions per element and <= 64 registers (the generic kernel: 31 and 128 registers -> two blocks per SM).
template <int NV, int PARTS, int ROWS>
__global__ void __launch_bounds__(256, 4) layernorm_fast_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                                const float* __restrict__ beta, float eps, int rows,
                                                                float* __restrict__ y, __nv_bfloat16* __restrict__ yb) {
  constexpr int C = 128 * NV, V = 32 * NV;
  __shared__ float4 s_g[V], s_b[V];
  pdl_trigger();
  if (threadIdx.x < V) {
    s_g[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(gamma) + threadIdx.x);
    s_b[threadIdx.x] = __ldg(reinterpret_cast<const float4*>(beta) + threadIdx.x);
  }
  pdl_wait();
  __syncthreads();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row0 = warp * ROWS;
  if (row0 >= rows) return;
  float4 v[ROWS][NV];
  const float4* src = reinterpret_cast<const float4*>(x) + (size_t)row0 * V + lane;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const bool ok = row0 + r < rows;
#pragma unroll
    for (int i = 0; i < NV; ++i) v[r][i] = ok ? __ldg(src + r * V + 32 * i) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float sum[ROWS], sq[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    sum[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) sum[r] += (v[r][i].x + v[r][i].y) + (v[r][i].z + v[r][i].w);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) sum[r] += __shfl_xor_sync(0xffffffffu, sum[r], off);
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    sum[r] = sum[r] / (float)C;                   // mean
    sq[r] = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      v[r][i].x -= sum[r]; v[r][i].y -= sum[r]; v[r][i].z -= sum[r]; v[r][i].w -= sum[r];
      sq[r] += (v[r][i].x * v[r][i].x + v[r][i].y * v[r][i].y) + (v[r][i].z * v[r][i].z + v[r][i].w * v[r][i].w);
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1)
#pragma unroll
    for (int r = 0; r < ROWS; ++r) sq[r] += __shfl_xor_sync(0xffffffffu, sq[r], off);
  constexpr int PW = PARTS > 0 ? PARTS : 1;
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    if (row0 + r >= rows) break;
    const float rstd = rsqrtf(sq[r] / (float)C + eps);
    const size_t row = (size_t)(row0 + r);
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float4 g = s_g[lane + 32 * i], be = s_b[lane + 32 * i];
      const float o0 = v[r][i].x * rstd * g.x + be.x, o1 = v[r][i].y * rstd * g.y + be.y, o2 = v[r][i].z * rstd * g.z + be.z,
                  o3 = v[r][i].w * rstd * g.w + be.w;
      if (PARTS > 0) store_planes4_ct<PW>(reinterpret_cast<uint2*>(yb) + row * (PW * V) + lane + 32 * i, V, o0, o1, o2, o3);
      else reinterpret_cast<float4*>(y)[row * V + lane + 32 * i] = make_float4(o0, o1, o2, o3);
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
                            int silu, float* y, __nv_bfloat16* yb, int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride) {
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
                                           silu, y, yb, parts, rawb, (long long)ss_bstride);
}

// cudaErrorNotSupported when the slab of one (utterance, group) does not fit shared memory: use stats + apply then.
cudaError_t launch_gn_fused(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                            const float* gamma, const float* beta, const float* ss, int silu, float* y, __nv_bfloat16* yb,
                            int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride) {
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
                    silu, y, yb, parts, rawb, (long long)ss_bstride);
}

// Cluster single-pass GroupNorm.  cudaErrorNotSupported when even an eighth of the slab does not fit shared memory.
namespace {
template <int PARTS, bool SILU, bool SS, bool RAW>
cudaError_t launch_gnc(dim3 grid, size_t smem, int cl, cudaStream_t s, const float* x1, int c1, const float* x2, int c2, int T, int groups,
                       int n_items, int tc, float eps, const float* gamma, const float* beta, const float* ss, float* y, __nv_bfloat16* yb,
                       __nv_bfloat16* rawb, long long ss_b) {
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(gn_cluster_kernel<PARTS, SILU, SS, RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gn_cluster_kernel<PARTS, SILU, SS, RAW>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(gn_cluster_kernel<PARTS, SILU, SS, RAW>, grid, dim3(GNC_THREADS), smem, s, cl, x1, c1, x2, c2, T, groups, n_items, tc, eps,
                    gamma, beta, ss, y, yb, rawb, ss_b);
}
template <int PARTS, typename... A>
cudaError_t launch_gnc_flags(bool silu, bool has_ss, bool raw, A... a) {
  if (PARTS == 0) raw = false;
  const int f = (silu ? 4 : 0) | (has_ss ? 2 : 0) | (raw ? 1 : 0);
  switch (f) {
    case 0: return launch_gnc<PARTS, false, false, false>(a...);
    case 1: return launch_gnc<PARTS, false, false, PARTS != 0>(a...);
    case 2: return launch_gnc<PARTS, false, true, false>(a...);
    case 3: return launch_gnc<PARTS, false, true, PARTS != 0>(a...);
    case 4: return launch_gnc<PARTS, true, false, false>(a...);
    case 5: return launch_gnc<PARTS, true, false, PARTS != 0>(a...);
    case 6: return launch_gnc<PARTS, true, true, false>(a...);
    default: return launch_gnc<PARTS, true, true, PARTS != 0>(a...);
  }
}
}  // namespace
cudaError_t launch_gn_cluster(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                              const float* gamma, const float* beta, const float* ss, int silu, float* y, __nv_bfloat16* yb,
                              int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride) {
  const int C = c1 + c2;
  const long long ss_b = (long long)ss_bstride;
  if (C % (4 * groups) || c1 % 4 || groups > 32 || B <= 0 || T <= 0) return cudaErrorInvalidValue;
  if (yb ? (parts < 1 || parts > 3) : (y == nullptr)) return cudaErrorInvalidValue;
  const int cg = C / groups, q = cg / 4;
  if (q > GNC_THREADS) return cudaErrorNotSupported;
  // cluster size: a function of (T, cg) only (batch-composition invariance); slices of <= 36 KB, two buffers per CTA
  int cl = 1;
  while (cl < 8 && (size_t)((T + cl - 1) / cl) * cg * sizeof(float) > 36 * 1024) cl *= 2;
  const int tc = (T + cl - 1) / cl;
  const size_t smem = 2 * (size_t)tc * cg * sizeof(float);
  if (smem > 200 * 1024) return cudaErrorNotSupported;
  // persistent grid: at most eight CTAs of 128 threads per SM (register bound), every cluster gets the same number of items when possible
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  const int per_sm = (int)std::max<size_t>(1, std::min<size_t>(8, (226 * 1024) / (smem + 1200)));
  const int n_items = B * groups;
  const int max_clusters = std::max(1, sms * per_sm / cl);
  const int rounds = (n_items + max_clusters - 1) / max_clusters;
  const int n_clusters = (n_items + rounds - 1) / rounds;
  const dim3 grid(n_clusters * cl);
  const bool raw = yb && rawb;
  if (!yb) return launch_gnc_flags<0>(silu != 0, ss != nullptr, false, grid, smem, cl, s, x1, c1, x2, c2, T, groups, n_items, tc, eps, gamma, beta, ss, y, yb, rawb, ss_b);
  if (parts == 1) return launch_gnc_flags<1>(silu != 0, ss != nullptr, raw, grid, smem, cl, s, x1, c1, x2, c2, T, groups, n_items, tc, eps, gamma, beta, ss, y, yb, rawb, ss_b);
  if (parts == 2) return launch_gnc_flags<2>(silu != 0, ss != nullptr, raw, grid, smem, cl, s, x1, c1, x2, c2, T, groups, n_items, tc, eps, gamma, beta, ss, y, yb, rawb, ss_b);
  return launch_gnc_flags<3>(silu != 0, ss != nullptr, raw, grid, smem, cl, s, x1, c1, x2, c2, T, groups, n_items, tc, eps, gamma, beta, ss, y, yb, rawb, ss_b);
}

// Wide rows (512 < C <= 2048; the Whisper encoder's 1280): one warp per row, the row held in registers (up to 16 float4 per lane),
// two-pass statistics (mean, then centred sum of squares) like the narrow kernels, gamma / beta read through the read-only path.
template <int MAXV>
__global__ void __launch_bounds__(256) layernorm_wide_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                             const float* __restrict__ beta, float eps, int rows, int C,
                                                             float* __restrict__ y, __nv_bfloat16* __restrict__ yb, int parts) {
  pdl_trigger();
  pdl_wait();
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int V = C >> 2;
  const float4* src = reinterpret_cast<const float4*>(x + (size_t)row * C);
  float4 v[MAXV];
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (lane + 32 * i < V) v[i] = __ldg(src + lane + 32 * i);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (lane + 32 * i < V) sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  sum = warp_sum(sum);
  const float mean = sum / (float)C;
  float sq = 0.f;
#pragma unroll
  for (int i = 0; i < MAXV; ++i)
    if (lane + 32 * i < V) {
      const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
      sq += (a * a + b * b) + (c * c + d * d);
    }
  sq = warp_sum(sq);
  const float rstd = rsqrtf(sq / (float)C + eps);
#pragma unroll
  for (int i = 0; i < MAXV; ++i) {
    const int q = lane + 32 * i;
    if (q < V) {
      const float4 g = __ldg(reinterpret_cast<const float4*>(gamma) + q), be = __ldg(reinterpret_cast<const float4*>(beta) + q);
      const float4 o = make_float4((v[i].x - mean) * rstd * g.x + be.x, (v[i].y - mean) * rstd * g.y + be.y,
                                   (v[i].z - mean) * rstd * g.z + be.z, (v[i].w - mean) * rstd * g.w + be.w);
      if (yb) store_planes4(yb + (size_t)row * (parts * C), q * 4, C, parts, o.x, o.y, o.z, o.w);
      else reinterpret_cast<float4*>(y + (size_t)row * C)[q] = o;
    }
  }
}

cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int C, float* y,
                             __nv_bfloat16* yb, int parts, cudaStream_t s) {
  if (C % 4 || C > 2048) return cudaErrorInvalidValue;
  if (C > 512) {
    const dim3 g((rows + 7) / 8), b(256);
    if (C <= 1024) return launch_pdl(layernorm_wide_kernel<8>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
    if (C <= 1536) return launch_pdl(layernorm_wide_kernel<12>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
    return launch_pdl(layernorm_wide_kernel<16>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
  }
  if (yb && (parts < 1 || parts > 3)) return cudaErrorInvalidValue;
  if (C == 256 || C == 384 || C == 512) {        // the denoiser's widths: specialised kernel
    const int pk = yb ? parts : 0;
    auto go = [&](auto kernel, int rows_per_warp) {
      const int rows_per_block = 8 * rows_per_warp;
      return launch_pdl(kernel, dim3((rows + rows_per_block - 1) / rows_per_block), dim3(256), 0, s, 1, x, gamma, beta, eps, rows, y, yb);
    };
    if (C == 256) return pk == 0 ? go(layernorm_fast_kernel<2, 0, 4>, 4) : pk == 1 ? go(layernorm_fast_kernel<2, 1, 4>, 4)
                                 : pk == 2 ? go(layernorm_fast_kernel<2, 2, 4>, 4) : go(layernorm_fast_kernel<2, 3, 4>, 4);
    if (C == 384) return pk == 0 ? go(layernorm_fast_kernel<3, 0, 2>, 2) : pk == 1 ? go(layernorm_fast_kernel<3, 1, 2>, 2)
                                 : pk == 2 ? go(layernorm_fast_kernel<3, 2, 2>, 2) : go(layernorm_fast_kernel<3, 3, 2>, 2);
    return pk == 0 ? go(layernorm_fast_kernel<4, 0, 2>, 2) : pk == 1 ? go(layernorm_fast_kernel<4, 1, 2>, 2)
                   : pk == 2 ? go(layernorm_fast_kernel<4, 2, 2>, 2) : go(layernorm_fast_kernel<4, 3, 2>, 2);
  }
  const int warps_per_block = 8, rows_per_block = warps_per_block * LN_ROWS;
  const int grid = (rows + rows_per_block - 1) / rows_per_block;
  const dim3 g(grid), b(warps_per_block * 32);
  if (C <= 256) return launch_pdl(layernorm_kernel<2>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
  if (C <= 384) return launch_pdl(layernorm_kernel<3>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
  return launch_pdl(layernorm_kernel<4>, g, b, 0, s, 1, x, gamma, beta, eps, rows, C, y, yb, parts);
}

}  // namespace lds
