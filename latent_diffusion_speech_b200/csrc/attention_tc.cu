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
//           K-major), O += P V accumulated in TMEM across all key tiles with the accumulate flag (m is final, so no
//           rescaling of O is ever needed)
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..5 softmax / epilogue (one query row per
// thread, tcgen05.ld 32x32b).  Single-buffered tiles; several CTAs per SM overlap one CTA's softmax with another's MMAs.
#include "lds_kernels.h"
#include "tc_ptx.cuh"
#include <math.h>

namespace lds {

cudaError_t tc_make_map_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                             int swizzle_bytes, CUtensorMap* out);

namespace {
using namespace ptx;

constexpr int AQ = 128, AKV = 64, ATT_TC_THREADS = 192;

struct AttnTcParams {
  int T, H, d, C, parts, n_pairs;
  int pair_a[6], pair_w[6];
  float scale;
  __nv_bfloat16* out;   // planes [B*T][parts*C]
};

template <int DPAD>
__global__ void __launch_bounds__(ATT_TC_THREADS)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapVT, const AttnTcParams p) {
  constexpr int SWZ = DPAD * 2;                    // swizzle width of the Q/K tiles (bytes per row)
  constexpr int QB = AQ * DPAD * 2, KB = AKV * DPAD * 2, VB = DPAD * AKV * 2, PB = AQ * AKV * 2;
  constexpr uint32_t IDESC_QK = umma_idesc_bf16(AQ, AKV), IDESC_PV = umma_idesc_bf16(AQ, DPAD);
  constexpr int TMEM_COLS = 128;                   // S: columns [0,64), O: columns [64, 64+DPAD)

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const int parts = p.parts;
  const uint32_t q_s = base, k_s = q_s + parts * QB, v_s = k_s + parts * KB, p_s = v_s + parts * VB;
  const uint32_t bar0 = p_s + parts * PB;
  uint8_t* p_ptr = smem + (p_s - base);
  const uint32_t q_full = bar0, k_full = bar0 + 8, k_empty = bar0 + 16, v_full = bar0 + 24, pv_done = bar0 + 32,
                 s_full = bar0 + 40, s_free = bar0 + 48, p_full = bar0 + 56;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + (bar0 - base) + 64);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;
  const int nt = (p.T + AKV - 1) / AKV;
  const int HD = p.H * DPAD;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapVT);
    mbar_init(q_full, 1); mbar_init(k_full, 1); mbar_init(k_empty, 1); mbar_init(v_full, 1); mbar_init(pv_done, 1);
    mbar_init(s_full, 1); mbar_init(s_free, 128); mbar_init(p_full, 128);
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  const uint32_t tmem_s = tmem0, tmem_o = tmem0 + 64;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, parts * QB);
      for (int pl = 0; pl < parts; ++pl) tma_load_3d(q_s + pl * QB, &mapQ, q_full, pl * HD + h * DPAD, q0, b);
      for (int it = 0; it < 2 * nt; ++it) {
        const int jt = it < nt ? it : it - nt;
        mbar_wait(k_empty, ((uint32_t)it & 1u) ^ 1u);
        mbar_expect_tx(k_full, parts * KB);
        for (int pl = 0; pl < parts; ++pl) tma_load_3d(k_s + pl * KB, &mapK, k_full, pl * HD + h * DPAD, jt * AKV, b);
        if (it >= nt) {
          mbar_wait(pv_done, ((uint32_t)jt & 1u) ^ 1u);
          mbar_expect_tx(v_full, parts * VB);
          for (int pl = 0; pl < parts; ++pl)
            tma_load_2d(v_s + pl * VB, &mapVT, v_full, jt * AKV, ((b * parts + pl) * p.H + h) * DPAD);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      mbar_wait(q_full, 0);
      for (int it = 0; it < 2 * nt; ++it) {
        mbar_wait(k_full, (uint32_t)it & 1u);
        if (it >= 1 && it - 1 < nt) mbar_wait(s_free, (uint32_t)(it - 1) & 1u);   // pass-A readers released S
        tc_fence_after();
        bool first = true;
        for (int pr = 0; pr < p.n_pairs; ++pr) {
          const uint64_t qd = umma_desc_kmajor(q_s + p.pair_a[pr] * QB, SWZ);
          const uint64_t kd = umma_desc_kmajor(k_s + p.pair_w[pr] * KB, SWZ);
#pragma unroll
          for (int k = 0; k < DPAD / 16; ++k) {
            umma_bf16(tmem_s, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), IDESC_QK, first ? 0u : 1u);
            first = false;
          }
        }
        umma_commit(k_empty);
        umma_commit(s_full);
        if (it >= nt) {
          const int jb = it - nt;
          mbar_wait(v_full, (uint32_t)jb & 1u);
          mbar_wait(p_full, (uint32_t)jb & 1u);
          tc_fence_after();
          bool pfirst = jb == 0;
          for (int pr = 0; pr < p.n_pairs; ++pr) {
            const uint64_t pd = umma_desc_kmajor(p_s + p.pair_a[pr] * PB, 128);
            const uint64_t vd = umma_desc_kmajor(v_s + p.pair_w[pr] * VB, 128);
#pragma unroll
            for (int k = 0; k < AKV / 16; ++k) {
              umma_bf16(tmem_o, pd + (uint64_t)(2 * k), vd + (uint64_t)(2 * k), IDESC_PV, pfirst ? 0u : 1u);
              pfirst = false;
            }
          }
          umma_commit(pv_done);
        }
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    float m = -INFINITY;
    // ---- pass A: row maximum of the scaled, masked scores ----
    for (int it = 0; it < nt; ++it) {
      mbar_wait(s_full, (uint32_t)it & 1u);
      tc_fence_after();
      const int k0 = it * AKV;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float s[32];
        tmem_ld32(tmem_s + lane_off + c * 32, s);
#pragma unroll
        for (int i = 0; i < 32; ++i)
          if (k0 + c * 32 + i < p.T) m = fmaxf(m, s[i] * p.scale);
      }
      tc_fence_before();
      mbar_arrive(s_free);
    }
    // ---- pass B: probabilities, row sum, P planes to shared memory ----
    float l = 0.f;
    uint8_t* prow = p_ptr + (row >> 3) * 1024 + (row & 7) * 128;     // 8-row / 1024 B swizzle atoms, 128 B per row
    for (int jb = 0; jb < nt; ++jb) {
      const int it = nt + jb, k0 = jb * AKV;
      mbar_wait(s_full, (uint32_t)it & 1u);
      tc_fence_after();
      float s0[32], s1[32];
      tmem_ld32(tmem_s + lane_off, s0);
      tmem_ld32(tmem_s + lane_off + 32, s1);
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        s0[i] = (k0 + i < p.T) ? expf(s0[i] * p.scale - m) : 0.f;
        s1[i] = (k0 + 32 + i < p.T) ? expf(s1[i] * p.scale - m) : 0.f;
        l += s0[i];
      }
#pragma unroll
      for (int i = 0; i < 32; ++i) l += s1[i];
      if (jb > 0) mbar_wait(pv_done, (uint32_t)(jb - 1) & 1u);       // previous P*V has consumed the P tile
      for (int pl = 0; pl < parts; ++pl) {
        uint8_t* dst = prow + pl * PB;
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {                               // 16-byte chunk = 8 keys
          float* src = ch < 4 ? &s0[ch * 8] : &s1[(ch - 4) * 8];
          uint32_t w[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const __nv_bfloat16 a = __float2bfloat16_rn(src[2 * i]), bb = __float2bfloat16_rn(src[2 * i + 1]);
            src[2 * i] -= __bfloat162float(a);
            src[2 * i + 1] -= __bfloat162float(bb);
            w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(bb) << 16);
          }
          *reinterpret_cast<uint4*>(dst + ((ch ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_async_smem();
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // ---- epilogue: O / l -> bf16 planes ----
    mbar_wait(pv_done, (uint32_t)(nt - 1) & 1u);
    tc_fence_after();
    const int q = q0 + row;
    const float inv = 1.f / l;
#pragma unroll 1
    for (int c = 0; c < DPAD / 32; ++c) {
      float o[32];
      tmem_ld32(tmem_o + lane_off + c * 32, o);
      if (q < p.T) {
        __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q) * (size_t)(parts * p.C) + h * p.d + c * 32;
        const int ncol = min(32, p.d - c * 32);                       // d = 48: only 16 valid columns in the 2nd chunk
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv;
        for (int pl = 0; pl < parts; ++pl) {
          uint32_t w[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const __nv_bfloat16 a = __float2bfloat16_rn(o[2 * i]), bb = __float2bfloat16_rn(o[2 * i + 1]);
            o[2 * i] -= __bfloat162float(a);
            o[2 * i + 1] -= __bfloat162float(bb);
            w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(bb) << 16);
          }
          uint4* dst = reinterpret_cast<uint4*>(orow + (size_t)pl * p.C);
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (i * 8 < ncol) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
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

template <int DPAD>
cudaError_t launch_attn(const AttnTcArgs& a, cudaStream_t s) {
  const int parts = a.parts;
  const size_t smem = (size_t)parts * (AQ * DPAD * 2 + AKV * DPAD * 2 + DPAD * AKV * 2 + AQ * AKV * 2) + 1024 + 128;
  cudaError_t e = cudaFuncSetAttribute(attention_tc_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
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
  attention_tc_kernel<DPAD><<<grid, ATT_TC_THREADS, smem, s>>>(mQ, mK, mV, p);
  return cudaGetLastError();
}

}  // namespace

cudaError_t launch_attention_tc(const AttnTcArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if ((a.parts != 1 && a.parts != 3) || a.d > a.dpad || a.d % 8 || a.T_pad % 8 || a.T_pad < a.T) return cudaErrorInvalidValue;
  switch (a.dpad) {
    case 32: return launch_attn<32>(a, s);
    case 64: return launch_attn<64>(a, s);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace lds
