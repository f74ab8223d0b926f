// attention_tc.cu — tcgen05 / TMEM / TMA flash attention over the time axis (self-attention, no mask).
//
// Replaces F.scaled_dot_product_attention (diffusion/unet1d/attention_processor.py:1032-1034) on the tensor-core
// path.  Operands come from the fused QKV projection whose epilogue (gemm_tc.cu, out_kind 3) writes
//   Q, K : bf16 planes [B*T][parts][H][dpad]          (head dim padded to 32/64, pad columns are exact zeros)
//   V^T  : bf16 planes [B][parts][H][dpad][T_pad]      (keys contiguous -> K-major B operand of the P*V product)
// so that every MMA operand is a K-major, TMA-swizzled tile.  parts = 1: bf16 mode.  parts = 3: fp32-accurate mode,
// both products are evaluated as the six significant plane products (split-bf16), softmax in fp32 with expf.
//
// CTA = 128 queries of one (utterance, head); keys stream in tiles of 64.  Two passes over the keys:
//   pass A: S = Q K^T -> TMEM, softmax warps reduce the row maximum m (no exponentials)
//   pass B: S again, P = exp(S*scale - m) (fp32), row sums in registers, P planes -> shared memory (128B-swizzled,
//           K-major), O += P V accumulated in TMEM across all key tiles with the accumulate flag (m is final, so O
//           never needs rescaling)
// Pipeline: K tiles in an NK-stage TMA ring, V^T tiles in an NV-stage ring, S double-buffered in TMEM so that
// Q K^T of tile j+1 is issued before the softmax of tile j has finished and P V of tile j overlaps softmax j+1.
// Warp roles (10 warps): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..9 softmax / epilogue — two
// warps per TMEM lane quarter, each thread owns one query row and half of the key columns of a tile.
#include "lds_kernels.h"
#include "tc_ptx.cuh"
#include <math.h>

namespace lds {

cudaError_t tc_make_map_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                             int swizzle_bytes, CUtensorMap* out);

namespace {
using namespace ptx;

constexpr int AQ = 128, AKV = 64, ATT_TC_THREADS = 320, NSOFT = 256;

struct AttnTcParams {
  int T, H, d, C, parts, n_pairs;
  int pair_a[6], pair_w[6];
  float scale;
  __nv_bfloat16* out;   // planes [B*T][parts*C]
};

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }

// Packs (a, b) to bf16x2 with round-to-nearest and leaves the residuals a - bf16(a), b - bf16(b) in place
// (one cvt.rn.bf16x2 + two integer ops + two FADDs per pair instead of per-element conversions).
__device__ __forceinline__ uint32_t split_pair(float& a, float& b) {
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));   // low half <- a, high half <- b
  a -= __uint_as_float(w << 16);
  b -= __uint_as_float(w & 0xffff0000u);
  return w;
}
__device__ __forceinline__ uint32_t pack_pair(float a, float b) {
  uint32_t w;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(w) : "f"(b), "f"(a));
  return w;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int DPAD, int NK, int NV, int NP>
__global__ void __launch_bounds__(ATT_TC_THREADS)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapVT, const AttnTcParams p) {
  constexpr int SWZ = DPAD * 2;                    // swizzle width of the Q/K tiles (bytes per row)
  constexpr int QB = AQ * DPAD * 2, KB = AKV * DPAD * 2, VB = DPAD * AKV * 2, PB = AQ * AKV * 2;
  constexpr uint32_t IDESC_QK = umma_idesc_bf16(AQ, AKV), IDESC_PV = umma_idesc_bf16(AQ, DPAD);
  constexpr int TMEM_COLS = 256;                   // S0: [0,64), S1: [64,128), O: [128, 128+DPAD)
  constexpr int OC = DPAD / 2;                     // output columns per softmax thread

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int parts = p.parts;
  const uint32_t q_s = base, k_s = q_s + parts * QB, v_s = k_s + NK * parts * KB, p_s = v_s + NV * parts * VB;
  const uint32_t bar0 = p_s + NP * parts * PB;
  uint8_t* p_ptr = smem + (p_s - base);
  // barriers (8 B each)
  const uint32_t q_full = bar0;
  auto k_full = [&](int s) { return bar0 + 8u * (1 + s); };
  auto k_empty = [&](int s) { return bar0 + 8u * (1 + NK + s); };
  auto v_full = [&](int s) { return bar0 + 8u * (1 + 2 * NK + s); };
  auto v_empty = [&](int s) { return bar0 + 8u * (1 + 2 * NK + NV + s); };
  const uint32_t bar1 = bar0 + 8u * (1 + 2 * NK + 2 * NV);
  auto s_full = [&](int s) { return bar1 + 8u * s; };
  auto s_free = [&](int s) { return bar1 + 8u * (2 + s); };
  auto p_full = [&](int s) { return bar1 + 32 + 8u * s; };
  auto pv_done = [&](int s) { return bar1 + 48 + 8u * s; };
  uint8_t* misc = smem + (bar1 - base) + 64;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc);
  float* red = reinterpret_cast<float*>(misc + 16);          // [2][128] exchange of row max / row sum between the halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (p.T + AKV - 1) / AKV;
  const int HD = p.H * DPAD;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < NK; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    for (int s = 0; s < NV; ++s) { mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(s_free(s), NSOFT); }
    for (int s = 0; s < NP; ++s) { mbar_init(p_full(s), NSOFT); mbar_init(pv_done(s), 1); }
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  const uint32_t tmem_o = tmem0 + 128;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, parts * QB);
      for (int pl = 0; pl < parts; ++pl) tma_load_3d(q_s + pl * QB, &mapQ, q_full, pl * HD + h * DPAD, q0, b);
      for (int it = 0; it < 2 * nt; ++it) {
        const int jt = it < nt ? it : it - nt;
        const int ks = it % NK;
        mbar_wait(k_empty(ks), ((uint32_t)(it / NK) & 1u) ^ 1u);
        mbar_expect_tx(k_full(ks), parts * KB);
        for (int pl = 0; pl < parts; ++pl)
          tma_load_3d(k_s + (ks * parts + pl) * KB, &mapK, k_full(ks), pl * HD + h * DPAD, jt * AKV, b);
        if (it >= nt) {
          const int vs = jt % NV;
          mbar_wait(v_empty(vs), ((uint32_t)(jt / NV) & 1u) ^ 1u);
          mbar_expect_tx(v_full(vs), parts * VB);
          for (int pl = 0; pl < parts; ++pl)
            tma_load_2d(v_s + (vs * parts + pl) * VB, &mapVT, v_full(vs), jt * AKV, ((b * parts + pl) * p.H + h) * DPAD);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      // S = Q K^T of global iteration `it` into S buffer it&1
      auto issue_qk = [&](int it) {
        const int ks = it % NK, sb = it & 1, u = it >> 1;
        mbar_wait(k_full(ks), (uint32_t)(it / NK) & 1u);
        mbar_wait(s_free(sb), ((uint32_t)u & 1u) ^ 1u);        // previous use of this S buffer has been read
        tc_fence_after();
        bool first = true;
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          const uint64_t qd = umma_desc_kmajor(q_s + p.pair_a[pr] * QB, SWZ);
          const uint64_t kd = umma_desc_kmajor(k_s + (ks * parts + p.pair_w[pr]) * KB, SWZ);
#pragma unroll
          for (int k = 0; k < DPAD / 16; ++k) {
            umma_bf16(tmem0 + sb * 64, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), IDESC_QK, first ? 0u : 1u);
            first = false;
          }
        }
        umma_commit(k_empty(ks));
        umma_commit(s_full(sb));
      };
      for (int it = 0; it < nt; ++it) issue_qk(it);            // pass A
      issue_qk(nt);                                            // pass B runs Q K^T one tile ahead of P V
      for (int jb = 0; jb < nt; ++jb) {
        if (jb + 1 < nt) issue_qk(nt + jb + 1);
        const int vs = jb % NV, pb = jb % NP;
        mbar_wait(v_full(vs), (uint32_t)(jb / NV) & 1u);
        mbar_wait(p_full(pb), (uint32_t)(jb / NP) & 1u);
        tc_fence_after();
        bool pfirst = jb == 0;
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          const uint64_t pd = umma_desc_kmajor(p_s + (pb * parts + p.pair_a[pr]) * PB, 128);
          const uint64_t vd = umma_desc_kmajor(v_s + (vs * parts + p.pair_w[pr]) * VB, 128);
#pragma unroll
          for (int k = 0; k < AKV / 16; ++k) {
            umma_bf16(tmem_o, pd + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), IDESC_PV, pfirst ? 0u : 1u);
            pfirst = false;
          }
        }
        umma_commit(v_empty(vs));
        umma_commit(pv_done(pb));
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = half * 32;                            // this thread's key columns inside a tile
    const float sl2 = p.scale * 1.4426950408889634f;    // softmax scale * log2(e): exp(x*scale - m) = 2^(x*sl2 - m*sl2)
    float m = -INFINITY;
    // ---- pass A: row maximum of the scaled, masked scores ----
    for (int it = 0; it < nt; ++it) {
      const int sb = it & 1;
      mbar_wait(s_full(sb), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem0 + lane_off + sb * 64 + c0, s);
      tc_fence_before();
      mbar_arrive(s_free(sb));
      const int k0 = it * AKV + c0;
      if (k0 + 32 <= p.T) {
#pragma unroll
        for (int i = 0; i < 32; ++i) m = fmaxf(m, s[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + i < p.T) m = fmaxf(m, s[i]);
      }
    }
    m *= sl2;                                            // row maximum in the log2 domain (sl2 > 0)
    red[half * 128 + row] = m;
    softmax_bar();
    m = fmaxf(m, red[(half ^ 1) * 128 + row]);
    softmax_bar();
    // ---- pass B: probabilities, row sum, P planes to shared memory ----
    float l = 0.f;
    uint8_t* prow = p_ptr + (row >> 3) * 1024 + (row & 7) * 128;     // 8-row / 1024 B swizzle atoms, 128 B per row
    for (int jb = 0; jb < nt; ++jb) {
      const int it = nt + jb, sb = it & 1;
      mbar_wait(s_full(sb), (uint32_t)(it >> 1) & 1u);
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem0 + lane_off + sb * 64 + c0, s);
      tc_fence_before();
      mbar_arrive(s_free(sb));
      const int k0 = jb * AKV + c0;
#pragma unroll
      for (int i = 0; i < 32; ++i) s[i] = ex2_approx(fmaf(s[i], sl2, -m));
      if (k0 + 32 > p.T) {                                             // ragged last tile: keys beyond T contribute nothing
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + i >= p.T) s[i] = 0.f;
      }
      float l0 = 0.f, l1 = 0.f;
#pragma unroll
      for (int i = 0; i < 32; i += 2) { l0 += s[i]; l1 += s[i + 1]; }
      l += l0 + l1;
      // split into bf16 planes in registers first, so that only the stores sit behind the P-buffer hand-off
      uint32_t w[3][16];
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        if (pl < parts) {
          const bool last_plane = pl == parts - 1;
#pragma unroll
          for (int i = 0; i < 16; ++i) w[pl][i] = last_plane ? pack_pair(s[2 * i], s[2 * i + 1]) : split_pair(s[2 * i], s[2 * i + 1]);
        }
      }
      const int pb = jb % NP;
      if (jb >= NP) mbar_wait(pv_done(pb), (uint32_t)(jb / NP - 1) & 1u);   // the P*V that last used this P buffer is done
#pragma unroll
      for (int pl = 0; pl < 3; ++pl) {
        if (pl < parts) {
          uint8_t* dst = prow + (pb * parts + pl) * PB;
#pragma unroll
          for (int ch = 0; ch < 4; ++ch)                                // 16-byte chunk = 8 keys
            *reinterpret_cast<uint4*>(dst + (((half * 4 + ch) ^ (row & 7)) << 4)) =
                make_uint4(w[pl][4 * ch], w[pl][4 * ch + 1], w[pl][4 * ch + 2], w[pl][4 * ch + 3]);
        }
      }
      fence_async_smem();
      mbar_arrive(p_full(pb));
    }
    red[half * 128 + row] = l;
    softmax_bar();
    l += red[(half ^ 1) * 128 + row];
    // ---- epilogue: O / l -> bf16 planes; this thread writes output columns [half*OC, half*OC + OC) ----
    mbar_wait(pv_done((nt - 1) % NP), (uint32_t)((nt - 1) / NP) & 1u);
    tc_fence_after();
    const int q = q0 + row;
    const float inv = 1.f / l;
    float o[OC];
    if constexpr (OC == 32) tmem_ld32(tmem_o + lane_off + half * OC, o);
    else tmem_ld16(tmem_o + lane_off + half * OC, o);
    const int ncol = min(OC, p.d - half * OC);                        // d = 48: 16 valid columns in the upper half
    if (q < p.T && ncol > 0) {
      __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q) * (size_t)(parts * p.C) + h * p.d + half * OC;
#pragma unroll
      for (int i = 0; i < OC; ++i) o[i] *= inv;
      for (int pl = 0; pl < parts; ++pl) {
        uint32_t w[OC / 2];
#pragma unroll
        for (int i = 0; i < OC / 2; ++i) w[i] = split_pair(o[2 * i], o[2 * i + 1]);
        uint4* dst = reinterpret_cast<uint4*>(orow + (size_t)pl * p.C);
#pragma unroll
        for (int i = 0; i < OC / 8; ++i)
          if (i * 8 < ncol) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem0, TMEM_COLS);
  }
}

template <int DPAD, int NK, int NV, int NP>
cudaError_t launch_attn(const AttnTcArgs& a, cudaStream_t s) {
  const int parts = a.parts;
  const size_t smem = (size_t)parts * (AQ * DPAD * 2 + NK * AKV * DPAD * 2 + NV * DPAD * AKV * 2 + NP * AQ * AKV * 2) + 1024 +
                      8 * (1 + 2 * NK + 2 * NV) + 64 + 16 + 2 * 128 * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<DPAD, NK, NV, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const uint64_t HD = (uint64_t)a.H * DPAD;
  CUtensorMap mQ, mK, mV;
  {
    const uint64_t dims[3] = {(uint64_t)parts * HD, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t str[2] = {(uint64_t)parts * HD * 2, (uint64_t)parts * HD * 2 * a.T};
    const uint32_t boxq[3] = {(uint32_t)DPAD, (uint32_t)AQ, 1}, boxk[3] = {(uint32_t)DPAD, (uint32_t)AKV, 1};
    if ((e = tc_make_map_bf16(a.q, 3, dims, str, boxq, DPAD * 2, &mQ)) != cudaSuccess) return e;
    if ((e = tc_make_map_bf16(a.k, 3, dims, str, boxk, DPAD * 2, &mK)) != cudaSuccess) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.T, (uint64_t)a.B * parts * HD};
    const uint64_t str[1] = {(uint64_t)a.T_pad * 2};
    const uint32_t box[2] = {(uint32_t)AKV, (uint32_t)DPAD};
    if ((e = tc_make_map_bf16(a.vt, 2, dims, str, box, 128, &mV)) != cudaSuccess) return e;
  }
  AttnTcParams p;
  p.T = a.T; p.H = a.H; p.d = a.d; p.C = a.H * a.d; p.parts = parts;
  p.scale = 1.0f / sqrtf((float)a.d);
  p.out = a.out;
  if (parts == 3) {
    static const int pa[6] = {2, 0, 1, 1, 0, 0}, pw[6] = {0, 2, 1, 0, 1, 0};
    p.n_pairs = 6;
    for (int i = 0; i < 6; ++i) { p.pair_a[i] = pa[i]; p.pair_w[i] = pw[i]; }
  } else {
    p.n_pairs = 1;
    for (int i = 0; i < 6; ++i) p.pair_a[i] = p.pair_w[i] = 0;
  }
  dim3 grid((a.T + AQ - 1) / AQ, a.H, a.B);
  attention_tc_kernel<DPAD, NK, NV, NP><<<grid, ATT_TC_THREADS, smem, s>>>(mQ, mK, mV, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_tc(const AttnTcArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if ((a.parts != 1 && a.parts != 3) || a.d > a.dpad || a.d % 8 || a.T_pad % 8 || a.T_pad < a.T) return cudaErrorInvalidValue;
  // shared memory per CTA: split, dpad 32: 24+3*12+2*12+48 = 132 KB; split, dpad 64: 48+3*24+2*24+48 = 216 KB;
  //                        bf16: a third of that (two CTAs per SM, bounded by 2 x 256 TMEM columns)
  if (a.parts == 3) {
    if (a.dpad == 32) return launch_attn<32, 3, 2, 2>(a, s);   // 24 + 36 + 24 + 96 = 180 KB
    if (a.dpad == 64) return launch_attn<64, 3, 2, 1>(a, s);   // 48 + 72 + 48 + 48 = 216 KB
  } else {
    if (a.dpad == 32) return launch_attn<32, 4, 3, 2>(a, s);
    if (a.dpad == 64) return launch_attn<64, 4, 3, 2>(a, s);
  }
  return cudaErrorNotSupported;
}

}  // namespace lds
