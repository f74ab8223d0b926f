// attention_tc.cu — tcgen05 / TMEM / TMA flash attention over the time axis (self-attention, no mask).
//
// Replaces F.scaled_dot_product_attention (diffusion/unet1d/attention_processor.py:1032-1034) on the tensor-core
// path.  Operands come from the fused QKV projection whose epilogue (gemm_tc.cu, out_kind 3) writes
//   Q, K : bf16 planes [B*T][parts][H][dpad]          (head dim padded to 32/64, pad columns are exact zeros)
//   V^T  : bf16 planes [B][parts][H][dpad][T_pad]      (keys contiguous -> K-major B operand of the P*V product)
// so that every MMA operand is a K-major, TMA-swizzled tile.  parts = 1: bf16 mode.  parts = 2: fp32-accurate mode (split-f16,
// planes.cuh): Q, K, V^T and the probabilities P are two fp16 planes of the scaled value each, both products are evaluated as the
// three plane products x1*y1 + x1*y2 + x2*y1, softmax in fp32.  parts = 3: round 1's three bf16 planes / six products (kept for A/B
// through lds_op_qkv_attention_tc).
//
// Work item = 128 queries of one (utterance, head); persistent CTAs stride over the items; keys stream in tiles of 64.
// Two passes over the keys:
//   pass A: S ~ Q_1 K_1^T (first planes only: the maximum is only needed to ~1 %) -> TMEM, softmax warps reduce the
//           row maximum m (no exponentials)
//   pass B: S again, P = exp(S*scale - m) (fp32), row sums in registers, P planes -> shared memory (128B-swizzled,
//           K-major), O += P V accumulated in TMEM across all key tiles with the accumulate flag (m is final, so O
//           never needs rescaling)
// Pipeline: K tiles in an NK-stage TMA ring, V^T tiles in an NV-stage ring, S double-buffered in TMEM so that
// Q K^T of tile j+1 is issued before the softmax of tile j has finished and P V of tile j overlaps softmax j+1.
// Warp roles (10 warps): warp 0 TMA producer, warp 1 MMA issuer (+TMEM alloc), warps 2..9 softmax / epilogue — two
// warps per TMEM lane quarter, each thread owns one query row and half of the key columns of a tile.
#include "lds_kernels.h"
#include "planes.cuh"
#include "tc_ptx.cuh"
#include <math.h>
#include <stdlib.h>

namespace lds {

cudaError_t tc_make_map_bf16_cached(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                    int swizzle_bytes, CUtensorMap* out);

namespace {
using namespace ptx;

constexpr int AQ = 128, ATT_TC_THREADS = 320, NSOFT = 256;

struct AttnTcParams {
  int B, T, H, d, C, parts;
  int out_parts;        // operand planes of the output for the following GEMM: 1 bf16, 2 split-f16 (planes.cuh), 3 bf16 hi/mid/lo
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

// MMA work lists.  tcgen05.mma costs ~92 cycles for any N <= 128 (M = 128, K = 16; measured, tests/micro/bench_umma.cu),
// 96 for N = 192 and 128 for N = 256, so the plane products of the split mode are batched along N: one instruction
// multiplies an A plane with up to three B planes that sit in consecutive shared-memory rows and writes one
// accumulator column block per B plane; the softmax / epilogue threads add the blocks.
struct MmaItem { int a_plane, b_plane0, n_planes, blk; };

// DUAL (split modes): 256 TMEM columns per CTA (and, with three bf16 planes, two S accumulator blocks instead of three), so
// that two CTAs fit on one SM and the softmax of one overlaps the tensor-pipe work of the other.
//
// Ring depths are compile-time: NK / NV stages of the K / V^T rings (shared memory), NSB S buffers (TMEM), NPB P buffers (shared
// memory).  PA = 2: a pass-A step covers two 64-key tiles (the two hi-plane tiles fill plane slots 0 and 1 of a K stage -> one
// N = 128 instruction).  Everything the single MMA-issuing thread computes between two tcgen05.mma must be cheap: the ncu source
// page of the round-1 kernel (runtime ring depths -> integer divisions, ~440 instructions per key tile) showed that thread
// executing 96 % of its time while the tensor pipe was busy 67 % — the pipe was waiting for its issuer, not the other way round.
template <int DPAD, int AKV, int PARTS, bool DUAL, int NK, int NV, int NSB, int NPB>
__global__ void __launch_bounds__(ATT_TC_THREADS, (PARTS == 1 || DUAL) ? 2 : 1)
attention_tc_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                    const __grid_constant__ CUtensorMap mapVT, const AttnTcParams p) {
  constexpr int SWZ = DPAD * 2;                    // swizzle width of the Q/K tiles (bytes per row)
  constexpr int KBLK = AKV / 64;                   // 64-key (128-byte) K-blocks of the P and V^T operands
  constexpr int QB = AQ * DPAD * 2, KB = AKV * DPAD * 2, VBK = DPAD * 128, PBK = AQ * 128;
  constexpr int OC = DPAD / 2;                     // output columns per softmax thread
  constexpr int NCH = AKV / 64;                    // 32-column chunks per softmax thread and tile

  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  constexpr int parts = PARTS;
  constexpr int PA = 2;                            // 64-key tiles per pass-A step
  constexpr int KSL = parts > PA ? parts : PA;     // K tiles (plane slots) per stage of the K ring
  const uint32_t q_s = base, k_s = q_s + parts * QB, v_s = k_s + NK * KSL * KB, p_s = v_s + NV * parts * KBLK * VBK;
  const uint32_t bar0 = p_s + NPB * parts * KBLK * PBK;
  uint8_t* p_ptr = smem + (p_s - base);
  const uint32_t q_full = bar0;
  auto s_full = [&](int s) { return bar0 + 8 + 8u * s; };
  auto s_free = [&](int s) { return bar0 + 24 + 8u * s; };
  auto p_full = [&](int s) { return bar0 + 40 + 8u * s; };
  auto pv_done = [&](int s) { return bar0 + 56 + 8u * s; };
  auto k_full = [&](int s) { return bar0 + 72 + 8u * s; };
  auto k_empty = [&](int s) { return bar0 + 72 + 8u * (4 + s); };
  auto v_full = [&](int s) { return bar0 + 72 + 8u * (8 + s); };
  auto v_empty = [&](int s) { return bar0 + 72 + 8u * (12 + s); };
  uint8_t* misc = smem + (bar0 - base) + 72 + 8 * 16 + 16;        // + q_empty, o_free
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc);
  float* red = reinterpret_cast<float*>(misc + 16);          // [2][128] exchange of row max / row sum between the halves

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;   // provably warp-uniform: the role branches below do not diverge
  pdl_trigger();
  // Persistent CTA: work items (utterance, head, query tile) blockIdx.x, blockIdx.x + gridDim.x, ...; every barrier
  // phase follows a running use counter, so the TMA producer and the MMA warp run ahead into the next item (its Q / K
  // loads and its pass A overlap the softmax tail, the O read-out and the stores of the current one).
  const int nq = (p.T + AQ - 1) / AQ, n_items = nq * p.H * p.B;
  const int nt = (p.T + AKV - 1) / AKV;
  const int ntA = (nt + PA - 1) / PA;              // pass-A steps of PA key tiles
  const int HD = p.H * DPAD;
  constexpr int tmem_cols = (parts >= 2 && !DUAL) ? 512 : 256;
  // accumulator blocks: an instruction multiplies an A plane with up to `n_planes` B planes in consecutive shared-memory rows
  // and writes one block per B plane; the softmax / epilogue threads add the blocks.
  //   parts 2 (split-f16: h1, h2 of q, k, v^T and P): S = q1*[k1|k2] + q2*[k1]  ->  2 blocks, 2 instructions per k-step
  //   parts 3 (three bf16 planes, round 1): 3 blocks (one N = 192 instruction) or 2 blocks in DUAL mode
  constexpr int n_sblk = parts == 3 ? ((AKV == 64 && !DUAL) ? 3 : 2) : (parts == 2 ? 2 : 1);
  constexpr int s_stride = n_sblk * AKV;           // TMEM columns per S buffer
  constexpr int n_oblk = parts == 3 ? 3 : (parts == 2 ? 2 : 1);
  // split-f16 scales: q, k, v^T planes hold 16 x value (written by the QKV projection), P planes 2048 x P (P <= 1)
  constexpr float P_SCALE = parts == 2 ? 2048.f : 1.f;
  constexpr float QK_SCALE = parts == 2 ? PLANE_SCALE * PLANE_SCALE : 1.f, PV_SCALE = parts == 2 ? P_SCALE * PLANE_SCALE : 1.f;
  // S buffers start at column 0, the O blocks (and the alternate 128-column pass-A buffer) at o_col
  constexpr int o_col = NSB * s_stride <= 128 ? 128 : (NSB * s_stride <= 256 ? 256 : 384);
  static_assert(o_col + (n_oblk * DPAD > 128 ? n_oblk * DPAD : 128) <= tmem_cols, "TMEM layout");
  const uint32_t q_empty = bar0 + 72 + 8 * 16, o_free = q_empty + 8;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapVT);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    mbar_init(o_free, NSOFT);
    for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(s_free(s), NSOFT); mbar_init(p_full(s), NSOFT); mbar_init(pv_done(s), 1); }
    for (int s = 0; s < 4; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); mbar_init(v_full(s), 1); mbar_init(v_empty(s), 1); }
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), tmem_cols);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  const uint32_t tmem_o = tmem0 + o_col;
  pdl_wait();                                    // the set-up above overlapped the previous kernel's tail

  if (warp == 0) {
    if (lane == 0) {
      int ks = 0, vs = 0;                        // next stage of the K / V^T ring
      uint32_t kph = 0, vph = 0, local = 0;      // ring phases; items started
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
        const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
        mbar_wait(q_empty, (local & 1u) ^ 1u);                   // every Q K^T of the previous item has read Q
        mbar_expect_tx(q_full, parts * QB);
        for (int pl = 0; pl < parts; ++pl) tma_load_3d(q_s + pl * QB, &mapQ, q_full, pl * HD + h * DPAD, qt * AQ, b);
        for (int it = 0; it < ntA + nt; ++it) {
          const int jt = it < ntA ? it : it - ntA;
          mbar_wait(k_empty(ks), kph ^ 1u);
          if (it < ntA) {                                        // pass A needs the hi plane only: PA consecutive key tiles per stage
            mbar_expect_tx(k_full(ks), PA * KB);
            for (int u = 0; u < PA; ++u)                         // (a tile beyond T is all out-of-bounds: zero fill, full byte count)
              tma_load_3d(k_s + (ks * KSL + u) * KB, &mapK, k_full(ks), h * DPAD, (jt * PA + u) * AKV, b);
          } else {
            mbar_expect_tx(k_full(ks), parts * KB);
            for (int pl = 0; pl < parts; ++pl)
              tma_load_3d(k_s + (ks * KSL + pl) * KB, &mapK, k_full(ks), pl * HD + h * DPAD, jt * AKV, b);
          }
          if (it >= ntA) {
            mbar_wait(v_empty(vs), vph ^ 1u);
            mbar_expect_tx(v_full(vs), parts * KBLK * VBK);
            for (int kb = 0; kb < KBLK; ++kb)
              for (int pl = 0; pl < parts; ++pl)
                tma_load_2d(v_s + ((vs * KBLK + kb) * parts + pl) * VBK, &mapVT, v_full(vs), jt * AKV + kb * 64,
                            ((b * parts + pl) * p.H + h) * DPAD);
            if (++vs == NV) { vs = 0; vph ^= 1u; }
          }
          if (++ks == NK) { ks = 0; kph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    {   // the WHOLE warp runs this loop in lock step (all lanes poll the barriers); one elected lane issues each MMA / commit
      // Work lists (widest instruction first: it initialises every accumulator block it covers).  They are compile-time
      // constants, the loops over them unroll completely and every descriptor below lives in a (uniform) register.
      constexpr bool WIDE_QK = parts == 3 && AKV == 64 && !DUAL;   // three S accumulator blocks: one N = 192 instruction
      constexpr int n_qk = parts == 3 ? (WIDE_QK ? 3 : 4) : (parts == 2 ? 2 : 1), n_pv = parts == 3 ? 3 : (parts == 2 ? 2 : 1);
      constexpr MmaItem qk_f16[4] = {{0, 0, 2, 0}, {1, 0, 1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};     // q1*[k1|k2], q2*[k1]
      constexpr MmaItem pv_f16[3] = {{0, 0, 2, 0}, {1, 0, 1, 0}, {0, 0, 0, 0}};                    // p1*[v1|v2], p2*[v1]
      constexpr MmaItem qk_wide[4] = {{0, 0, 3, 0}, {1, 0, 2, 0}, {2, 0, 1, 0}, {0, 0, 0, 0}};
      constexpr MmaItem qk_dual[4] = {{0, 0, 2, 0}, {1, 0, 2, 0}, {0, 2, 1, 1}, {2, 0, 1, 0}};
      constexpr MmaItem qk_one[4] = {{0, 0, 1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
      constexpr MmaItem pv_split[3] = {{0, 0, 3, 0}, {1, 0, 2, 0}, {2, 0, 1, 0}};
      constexpr MmaItem pv_one[3] = {{0, 0, 1, 0}, {0, 0, 0, 0}, {0, 0, 0, 0}};
      auto qk = [&](int e) -> MmaItem { return parts == 3 ? (WIDE_QK ? qk_wide[e] : qk_dual[e]) : (parts == 2 ? qk_f16[e] : qk_one[e]); };
      auto pv = [&](int e) -> MmaItem { return parts == 3 ? pv_split[e] : (parts == 2 ? pv_f16[e] : pv_one[e]); };
      auto idesc = [&](int n) -> uint32_t { return parts == 2 ? umma_idesc_f16(AQ, n) : umma_idesc_bf16(AQ, n); };
      const uint64_t q_desc0 = umma_desc_kmajor(q_s, SWZ), k_desc0 = umma_desc_kmajor(k_s, SWZ);
      const uint64_t p_desc0 = umma_desc_kmajor(p_s, 128), v_desc0 = umma_desc_kmajor(v_s, 128);
      constexpr uint64_t k_stage_step = (uint64_t)((KSL * KB) >> 4), v_stage_step = (uint64_t)((parts * KBLK * VBK) >> 4),
                         v_kb_step = (uint64_t)((parts * VBK) >> 4), p_buf_step = (uint64_t)((parts * KBLK * PBK) >> 4),
                         p_kb_step = (uint64_t)(PBK >> 4);
      constexpr uint32_t hi_idesc = parts == 2 ? umma_idesc_f16(AQ, PA * AKV) : umma_idesc_bf16(AQ, PA * AKV);
      const int ksteps = (p.d + 15) / 16;        // head dims beyond d are zero padding (d = 48 in a 64-wide tile): skip their K slices
      // S buffer of an evaluation: pass B rotates over the NSB buffers of the S region; pass A (PA == 2) alternates between
      // columns [0, 128) and the O region (idle until the first P V of the item), so that the hi*hi product of step j+1 runs
      // while the softmax threads still reduce step j.
      uint32_t su0 = 0u, su1 = 0u, pu0 = 0u, pu1 = 0u, local = 0u;   // uses of the S / P buffers so far
      int ks = 0, vs = 0;                                            // oldest live stage of the K / V^T ring
      uint32_t kph = 0u, vph = 0u;
      bool o_claimed = false;
      auto claim_o = [&]() {                     // the softmax threads have read the previous item's O
        if (!o_claimed) {
          mbar_wait(o_free, (local & 1u) ^ 1u);
          o_claimed = true;
        }
      };
      auto k_advance = [&]() { if (++ks == NK) { ks = 0; kph ^= 1u; } };
      // pass-A step `it`: hi*hi of PA key tiles (the row maximum is only needed to ~1 %: any m close to it gives the same softmax)
      auto issue_qk_a = [&](int it) {
        const int sb = it & 1;
        if (sb) claim_o();
        mbar_wait(k_full(ks), kph);
        mbar_wait(s_free(sb), ((sb ? su1 : su0) & 1u) ^ 1u);   // the softmax threads have read the previous use of this S buffer
        // the 128 columns [0, 128) also cover pass-B buffer 1 ([64, 128) when NSB == 2): wait for its last use too
        if (NSB == 2 && sb == 0) mbar_wait(s_free(1), (su1 & 1u) ^ 1u);
        if (sb) ++su1; else ++su0;
        tc_fence_after();
        const uint32_t dst = tmem0 + (sb ? (uint32_t)o_col : 0u);
        const uint64_t kd = k_desc0 + (uint64_t)ks * k_stage_step;
#pragma unroll
        for (int k = 0; k < DPAD / 16; ++k)
          if (k < ksteps) umma_bf16_elect(dst, q_desc0 + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), hi_idesc, k == 0 ? 0u : 1u);
        umma_commit_elect(k_empty(ks));
        umma_commit_elect(s_full(sb));
        k_advance();
      };
      // pass-B tile `jb`: the plane products of Q K^T (three in split-f16 mode, issued as two instructions; one in bf16 mode)
      auto issue_qk_b = [&](int jb, bool last) {
        const int sb = NSB == 1 ? 0 : (jb & 1);
        mbar_wait(k_full(ks), kph);
        mbar_wait(s_free(sb), ((sb ? su1 : su0) & 1u) ^ 1u);
        if (sb) ++su1; else ++su0;
        tc_fence_after();
        const uint32_t dst = tmem0 + (uint32_t)(sb * s_stride);
        const uint64_t kd = k_desc0 + (uint64_t)ks * k_stage_step;
#pragma unroll
        for (int k = 0; k < DPAD / 16; ++k) {
          if (k < ksteps) {
#pragma unroll
            for (int e = 0; e < n_qk; ++e)
              umma_bf16_elect(dst + (uint32_t)(qk(e).blk * AKV), q_desc0 + (uint64_t)((qk(e).a_plane * QB) >> 4) + (uint64_t)(2 * k),
                        kd + (uint64_t)((qk(e).b_plane0 * KB) >> 4) + (uint64_t)(2 * k), idesc(qk(e).n_planes * AKV),
                        (k == 0 && e == 0) ? 0u : 1u);
          }
        }
        umma_commit_elect(k_empty(ks));
        umma_commit_elect(s_full(sb));
        if (last) umma_commit_elect(q_empty);                          // last Q K^T of the item: Q may be overwritten
        k_advance();
      };
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++local) {
        o_claimed = false;
        mbar_wait(q_full, local & 1u);
        for (int it = 0; it < ntA; ++it) issue_qk_a(it);         // pass A
        issue_qk_b(0, nt == 1);                                  // pass B: Q K^T runs one tile ahead of P V
        for (int jb = 0; jb < nt; ++jb) {
          if (jb + 1 < nt) issue_qk_b(jb + 1, jb + 2 == nt);
          const int pb = NPB == 1 ? 0 : (jb & 1);
          mbar_wait(v_full(vs), vph);
          mbar_wait(p_full(pb), (pb ? pu1 : pu0) & 1u);
          if (pb) ++pu1; else ++pu0;
          if (jb == 0) claim_o();
          tc_fence_after();
          const uint64_t pd = p_desc0 + (uint64_t)pb * p_buf_step, vd = v_desc0 + (uint64_t)vs * v_stage_step;
#pragma unroll
          for (int k = 0; k < AKV / 16; ++k) {
            const uint64_t pk = pd + (uint64_t)(k >> 2) * p_kb_step + (uint64_t)(2 * (k & 3));
            const uint64_t vk = vd + (uint64_t)(k >> 2) * v_kb_step + (uint64_t)(2 * (k & 3));
#pragma unroll
            for (int e = 0; e < n_pv; ++e)
              umma_bf16_elect(tmem_o + (uint32_t)(pv(e).blk * DPAD), pk + (uint64_t)((pv(e).a_plane * KBLK * PBK) >> 4),
                        vk + (uint64_t)((pv(e).b_plane0 * VBK) >> 4), idesc(pv(e).n_planes * DPAD),
                        (jb == 0 && k == 0 && e == 0) ? 0u : 1u);
          }
          umma_commit_elect(v_empty(vs));
          umma_commit_elect(pv_done(pb));
          if (++vs == NV) { vs = 0; vph ^= 1u; }
        }
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = half * (AKV / 2);                     // this thread's key columns inside a tile
    // softmax scale * log2(e) (/ the operand scales of the split-f16 planes): exp(x*scale - m) = 2^(x*sl2 - m*sl2)
    const float sl2 = p.scale * 1.4426950408889634f / QK_SCALE;
    // sum of the accumulator blocks of 32 columns starting at column c of the S tile
    auto load_s32 = [&](int sb, int c, float* s) {
      const uint32_t sbase = tmem0 + lane_off + sb * s_stride + c;
      tmem_ld32(sbase + (n_sblk - 1) * AKV, s);
      for (int blk = n_sblk - 2; blk >= 0; --blk) {
        float t[32];
        tmem_ld32(sbase + blk * AKV, t);
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] += t[i];
      }
    };
    uint32_t su0 = 0u, su1 = 0u;                         // per-buffer use counters (same sequence as the MMA warp)
    uint32_t pw0 = 0u, pw1 = 0u;                         // P buffer uses == P V completions to expect per buffer
    uint8_t* prow = p_ptr + (row >> 3) * 1024 + (row & 7) * 128;     // 8-row / 1024 B swizzle atoms, 128 B per row
    for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
      const int qt = item % nq, h = (item / nq) % p.H, b = item / (nq * p.H);
      float m = -INFINITY;
      // ---- pass A: row maximum of the scaled, masked scores ----
      const int nchA = NCH * PA, c0A = c0 * PA;            // this thread's half of a pass-A step of PA * AKV keys
      for (int it = 0; it < ntA; ++it) {
        const int sb = it & 1;
        const uint32_t soff = (NSB == 1 || PA == 2) ? (uint32_t)(sb * o_col) : (uint32_t)(sb * s_stride);
        mbar_wait(s_full(sb), (sb ? su1 : su0) & 1u);
        if (sb) ++su1; else ++su0;
        tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < nchA; ++ch) {
          float s[32];
          tmem_ld32(tmem0 + lane_off + soff + c0A + ch * 32, s);  // hi*hi scores live in block 0
          if (ch == nchA - 1) { tc_fence_before(); mbar_arrive(s_free(sb)); }
          const int k0 = it * PA * AKV + c0A + ch * 32;
          if (k0 + 32 <= p.T) {
#pragma unroll
            for (int i = 0; i < 32; ++i) m = fmaxf(m, s[i]);
          } else {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (k0 + i < p.T) m = fmaxf(m, s[i]);
          }
        }
      }
      m *= sl2;                                            // row maximum in the log2 domain (sl2 > 0)
      red[half * 128 + row] = m;
      softmax_bar();
      m = fmaxf(m, red[(half ^ 1) * 128 + row]);
      softmax_bar();
      // ---- pass B: probabilities, row sum, P planes to shared memory ----
      float l = 0.f;
      for (int jb = 0; jb < nt; ++jb) {
        const int sb = jb % NSB, pb = jb % NPB;
        mbar_wait(s_full(sb), (sb ? su1 : su0) & 1u);
        if (sb) ++su1; else ++su0;
        tc_fence_after();
        uint32_t w[PARTS][16 * NCH];
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
          float s[32];
          load_s32(sb, c0 + ch * 32, s);
          if (ch == NCH - 1) { tc_fence_before(); mbar_arrive(s_free(sb)); }
          const int k0 = jb * AKV + c0 + ch * 32;
#pragma unroll
          for (int i = 0; i < 32; ++i) s[i] = ex2_approx(fmaf(s[i], sl2, -m));
          if (k0 + 32 > p.T) {                                           // ragged last tile: keys beyond T contribute nothing
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (k0 + i >= p.T) s[i] = 0.f;
          }
          float l0 = 0.f, l1 = 0.f;
#pragma unroll
          for (int i = 0; i < 32; i += 2) { l0 += s[i]; l1 += s[i + 1]; }
          l += l0 + l1;
          // split into operand planes in registers, so that only the stores sit behind the P-buffer hand-off
          if constexpr (PARTS == 2) {                                    // split-f16: h1 = f16(P_SCALE * P), h2 = f16(rest)
#pragma unroll
            for (int i = 0; i < 32; ++i) s[i] *= P_SCALE;
#pragma unroll
            for (int i = 0; i < 16; ++i) w[0][ch * 16 + i] = planes_split_pair_f16(s[2 * i], s[2 * i + 1]);
#pragma unroll
            for (int i = 0; i < 16; ++i) w[1][ch * 16 + i] = planes_pack_pair_f16(s[2 * i], s[2 * i + 1]);
          } else {
#pragma unroll
            for (int pl = 0; pl < PARTS; ++pl) {
#pragma unroll
              for (int i = 0; i < 16; ++i)
                w[pl][ch * 16 + i] = (pl == PARTS - 1) ? pack_pair(s[2 * i], s[2 * i + 1]) : split_pair(s[2 * i], s[2 * i + 1]);
            }
          }
        }
        {                                                                // the P*V that last read this P buffer is done
          const uint32_t uses = pb ? pw1 : pw0;
          if (uses > 0) mbar_wait(pv_done(pb), (uses - 1) & 1u);
          if (pb) ++pw1; else ++pw0;
        }
#pragma unroll
        for (int pl = 0; pl < PARTS; ++pl) {
#pragma unroll
          for (int cc = 0; cc < 4 * NCH; ++cc) {                         // 16-byte chunk = 8 keys
            const int key_chunk = (c0 >> 3) + cc;                        // chunk index inside the AKV-key row
            uint8_t* dst = prow + ((pb * parts + pl) * KBLK + (key_chunk >> 3)) * PBK;
            *reinterpret_cast<uint4*>(dst + (((key_chunk & 7) ^ (row & 7)) << 4)) =
                make_uint4(w[pl][4 * cc], w[pl][4 * cc + 1], w[pl][4 * cc + 2], w[pl][4 * cc + 3]);
          }
        }
        fence_async_smem();
        mbar_arrive(p_full(pb));
      }
      red[half * 128 + row] = l;
      softmax_bar();
      l += red[(half ^ 1) * 128 + row];
      // ---- epilogue: O / l -> bf16 planes; this thread writes output columns [half*OC, half*OC + OC) ----
      {
        const int pb = (nt - 1) % NPB;
        mbar_wait(pv_done(pb), ((pb ? pw1 : pw0) - 1) & 1u);           // the last P V of the item
      }
      tc_fence_after();
      const int q = qt * AQ + row;
      const float inv = 1.f / l;
      float o[OC];
      {
        auto ld = [&](uint32_t addr, float* v) {
          if constexpr (OC == 32) tmem_ld32(addr, v); else tmem_ld16(addr, v);
        };
        ld(tmem_o + lane_off + (n_oblk - 1) * DPAD + half * OC, o);
        for (int blk = n_oblk - 2; blk >= 0; --blk) {
          float t[OC];
          ld(tmem_o + lane_off + blk * DPAD + half * OC, t);
#pragma unroll
          for (int i = 0; i < OC; ++i) o[i] += t[i];
        }
      }
      tc_fence_before();
      mbar_arrive(o_free);                                               // the next item may overwrite O
      const int ncol = min(OC, p.d - half * OC);                        // d = 48: 16 valid columns in the upper half
      if (q < p.T && ncol > 0) {
        __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q) * (size_t)(p.out_parts * p.C) + h * p.d + half * OC;
        const float oscale = (p.out_parts == 2 ? inv * PLANE_SCALE : inv) * (1.f / PV_SCALE);
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] *= oscale;
        for (int pl = 0; pl < p.out_parts; ++pl) {
          uint32_t w[OC / 2];
#pragma unroll
          for (int i = 0; i < OC / 2; ++i)
            w[i] = p.out_parts == 2 ? (pl == 0 ? planes_split_pair_f16(o[2 * i], o[2 * i + 1]) : planes_pack_pair_f16(o[2 * i], o[2 * i + 1]))
                                    : split_pair(o[2 * i], o[2 * i + 1]);
          uint4* dst = reinterpret_cast<uint4*>(orow + (size_t)pl * p.C);
#pragma unroll
          for (int i = 0; i < OC / 8; ++i)
            if (i * 8 < ncol) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
        }
      }
      softmax_bar();                                                     // red[] is reused by the next item
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem0, tmem_cols);
  }
}

template <int DPAD, int AKV, int PARTS, bool DUAL, int NK, int NV, int NSB, int NPB>
cudaError_t launch_attn(const AttnTcArgs& a, cudaStream_t s) {
  constexpr int PA = 2, KSL = PARTS > PA ? PARTS : PA;
  auto kernel = attention_tc_kernel<DPAD, AKV, PARTS, DUAL, NK, NV, NSB, NPB>;
  constexpr size_t smem = (size_t)PARTS * (AQ * DPAD * 2 + NV * DPAD * AKV * 2 + NPB * AQ * AKV * 2) + (size_t)KSL * NK * AKV * DPAD * 2 + 1024 +
                          72 + 8 * 16 + 16 + 16 + 2 * 128 * 4 + 64;
  static_assert(smem <= 227 * 1024, "attention tile configuration exceeds shared memory");
  cudaError_t e = cudaSuccess;
  static unsigned long long configured = 0;      // per template instance
  if (first_use_on_this_device(configured)) {
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int parts = PARTS;
  const uint64_t HD = (uint64_t)a.H * DPAD;
  CUtensorMap mQ, mK, mV;
  {
    const uint64_t dims[3] = {(uint64_t)parts * HD, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t str[2] = {(uint64_t)parts * HD * 2, (uint64_t)parts * HD * 2 * a.T};
    const uint32_t boxq[3] = {(uint32_t)DPAD, (uint32_t)AQ, 1}, boxk[3] = {(uint32_t)DPAD, (uint32_t)AKV, 1};
    if ((e = tc_make_map_bf16_cached(a.q, 3, dims, str, boxq, DPAD * 2, &mQ)) != cudaSuccess) return e;
    if ((e = tc_make_map_bf16_cached(a.k, 3, dims, str, boxk, DPAD * 2, &mK)) != cudaSuccess) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.T, (uint64_t)a.B * parts * HD};
    const uint64_t str[1] = {(uint64_t)a.T_pad * 2};
    const uint32_t box[2] = {64u, (uint32_t)DPAD};
    if ((e = tc_make_map_bf16_cached(a.vt, 2, dims, str, box, 128, &mV)) != cudaSuccess) return e;
  }
  AttnTcParams p;
  p.B = a.B; p.T = a.T; p.H = a.H; p.d = a.d; p.C = a.H * a.d; p.parts = parts;
  p.out_parts = a.out_parts > 0 ? a.out_parts : parts;
  p.scale = 1.0f / sqrtf((float)a.d);
  p.out = a.out;
  // persistent: one CTA per resident slot (two per SM in bf16 / DUAL mode), items strided over them
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms < 1) sms = 148;
  const int n_items = ((a.T + AQ - 1) / AQ) * a.H * a.B;
  const int slots = sms * ((PARTS == 1 || DUAL) ? 2 : 1);
  // (bf16, d <= 32: one item per CTA measured 3 % faster than the persistent grid — its items are too short to amortise anything)
  const bool persist = !(PARTS == 1 && DPAD == 32);
  dim3 grid(persist && n_items > slots ? slots : n_items);
  return launch_pdl(kernel, grid, dim3(ATT_TC_THREADS), smem, s, 1, mQ, mK, mV, p);
}

}  // namespace

cudaError_t launch_attention_tc(const AttnTcArgs& a, cudaStream_t s) {
  if (a.B <= 0 || a.T <= 0) return cudaSuccess;
  if (a.parts < 1 || a.parts > 3 || a.d > a.dpad || a.d % 8 || a.T_pad % 8 || a.T_pad < a.T) return cudaErrorInvalidValue;
  // shared memory per CTA (KB): Q + nk*K + nv*V^T + P; a K stage holds max(parts, 2) 64-key tiles (pass A loads two
  // hi-plane tiles per stage), two CTAs per SM in bf16 and DUAL mode (bounded by 2 x 256 TMEM columns)
  if (a.parts == 2) {
    // split-f16 operands (h1, h2 of q, k, v^T; P split the same way): 2 S blocks (128 columns) + 2 O blocks (2*dpad columns).
    // d <= 32: two CTAs per SM (DUAL): 16 (Q) + 2*8 (K) + 2*8 (V^T) + 32 (P) = 80 KB, TMEM 128 + 64.
    // d > 32: two CTAs per SM as well, with single-stage K and V^T rings — the second CTA's work covers the load
    //         latency the ring would: 32 (Q) + 16 (K) + 16 (V^T) + 32 (P) = 96 KB, TMEM 128 (S) + 128 (O).  Two stages of
    //         either ring (112 KB + barriers) is 1.3 KB too much for two CTAs per SM.  Measured against one CTA per SM with
    //         S and P double-buffered (160 KB, TMEM 2*128 + 128; LDS_ATT_DUAL64=0): op time -6% at T=432, -1% at T=216,
    //         attention -4.4% on the headline (profiles/r02_attention_dual64_call166.txt)
    if (a.dpad == 32) return launch_attn<32, 64, 2, true, 2, 2, 1, 1>(a, s);
    if (a.dpad == 64) {
      if (knobs().att_dual64) return launch_attn<64, 64, 2, true, 1, 1, 1, 1>(a, s);
      return launch_attn<64, 64, 2, false, 2, 2, 2, 2>(a, s);
    }
  } else if (a.parts == 3) {
    // three bf16 planes, six plane products (round 1's fp32-accurate form; kept for A/B through lds_op_qkv_attention_tc)
    if (a.dpad == 32) return launch_attn<32, 64, 3, true, 2, 1, 1, 1>(a, s);
    if (a.dpad == 64) return launch_attn<64, 64, 3, false, 2, 2, 1, 1>(a, s);
  } else {                                                             // bf16: S and P double-buffered, two CTAs per SM
    // K stages hold two 64-key tiles (pass A runs over 128-key steps): 4 (2) stages of 8 (16) KB
    if (a.dpad == 32) return launch_attn<32, 64, 1, false, 4, 3, 2, 2>(a, s);    //  8 + 4*8 + 3*4 + 2*16 = 84, TMEM 2*64 (S) + 32 (O)
    if (a.dpad == 64) return launch_attn<64, 64, 1, false, 2, 3, 2, 2>(a, s);    // 16 + 2*16 + 3*8 + 2*16 = 104
  }
  return cudaErrorNotSupported;
}

}  // namespace lds
