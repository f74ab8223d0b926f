// gemm_tc.cu — persistent tcgen05 / TMEM / TMA implicit-GEMM for Conv1d(k=3), Conv1d(k=1) and nn.Linear.
//
// One kernel serves both precision modes of the library:
//   * LDS_PREC_BF16 : A and W are bf16, one tcgen05.mma (kind::f16, fp32 accumulate in TMEM) per K slice.
//   * LDS_PREC_FP32 : "split-f16" (round 2) — every fp32 operand is stored as two fp16 planes of the scaled value (planes.cuh:
//     h1 = f16(s x), h2 = f16(s x - h1), 22 significant bits) and the product is evaluated as h1*w1 + (h1*w2 + h2*w1): THREE MMAs
//     per K slice instead of the six of round 1's three-bf16-plane split, and 4 instead of 6 operand bytes per element.  The
//     operand error of the three-product sum is a third of an fp32 FFMA GEMM's own rounding error
//     (profiles/r02_split_f16_operand_error.txt); what dominates the result error is the tensor core's truncating fp32
//     accumulate, as before.
//
// Structure (B200, sm_100a).  Measured on B200 (tests/micro/bench_umma.cu, tests/gpu_gemm_bench.py, ncu captures
// under profiles/): one tcgen05.mma M128 x K16 costs ~92 cycles for any N <= 128 but 96 / 128 cycles at N = 192 / 256,
// so only N >= 192 instructions run the pipe at its rate; and the main loop is bound by the bytes the L2 can DELIVER
// to one SM (~40 B/clk/SM; TMA multicast between two CTAs does not lower it, keeping a smaller share of W per SM does).
// Hence:
//   * CTA PAIR (cluster of 2, tcgen05.mma.cta_group::2): the pair computes a 256 (M) x BN (N) tile, each CTA holding
//     its own 128 accumulator rows in TMEM, its own A tile and HALF of the W tile in shared memory; rank 0 issues every
//     MMA for both SMs.  BN = 256 / 192 / 128 chosen per GEMM so that N % BN == 0.  Pairs are persistent, items
//     (M-tile pair, N tile) strided over the clusters with the N tiles of one row block adjacent (A rows stay hot in L2).
//     An odd M tile count leaves one ghost tile (TMA zero fill, no stores).
//   * K blocks of 64 elements (128-byte rows, 128B swizzle).  Operand tiles live in TWO shared-memory rings — A slots of
//     16 KB, W slots of BN/2 x 128 B — filled by each CTA's producer in the order the MMA warp needs them and released
//     tile by tile.  Split-f16 mode, per K block: loads A_h1 W_h2 A_h2 W_h1 (every plane tile once), products h1*w2 and
//     h2*w1 into the SMALL accumulator, h1*w1 into the MAIN one.  The two accumulators are separate TMEM column ranges:
//     the tensor core's fp32 accumulate truncates at every step, and the main accumulator should see as few steps as the
//     bf16 mode's (four per K block); the small one holds terms 2^-11 of the main, its truncation is irrelevant.  The epilogue
//     adds the two once, in fp32, and multiplies by 1 / (scale_A * scale_W).
//   * Barriers: "full" barriers live in the leader and count the TMA bytes of BOTH CTAs (the peer's loads complete on
//     the leader's barrier, cp.async.bulk.tensor.cta_group::2); "empty" and "accumulator full" barriers exist in both
//     CTAs and are signalled by the leader's multicast tcgen05.commit; "accumulator empty" lives in the leader and
//     collects one relaxed remote arrive per epilogue warp of both CTAs.
//   * TMEM (512 columns): bf16 mode double-buffers its BN-column accumulator (the epilogue of tile i overlaps the main loop
//     of tile i+1).  Split-f16 mode needs main + small = 2 BN columns per buffer: two buffers at BN = 128 (short K loops),
//     one at BN = 256 / 192 (long K loops, where the exposed epilogue is a small share).
//   warp 0      TMA producer  — cp.async.bulk.tensor (3-D map over [channels, frames, utterances] for A so that the
//                               three taps of a k=3 convolution are three shifted loads of the same tensor and the
//                               zero padding is TMA out-of-bounds fill; 2-D map for W), mbarrier tx counts
//   warp 1      MMA issuer    — tcgen05.alloc (cta_group::2), one elected lane of the leader issues tcgen05.mma and tcgen05.commit to
//                               release smem stages / publish an accumulator buffer
//   warps 2..9  epilogue      — tcgen05.ld 32x32b (one accumulator row per thread, two warps per TMEM lane quarter
//                               interleaving the 32-column chunks), + bias, SiLU / GEGLU, transpose through a per-warp
//                               shared-memory tile, + fp32 residual, coalesced stores: fp32 / bf16 / 3-plane split bf16 /
//                               attention operands
// Reference ops replaced: F.conv1d (lora.py:102), nn.Linear (attention_processor.py:1012-1040,
// attention.py:291,247), GEGLU (attention.py:299-301), residual adds (resnet.py:639, attention.py:161-201).
#include <cuda.h>
#include <cstdlib>

#include <mutex>
#include <unordered_map>

#include "lds_kernels.h"
#include "planes.cuh"
#include "tc_ptx.cuh"

namespace lds {
namespace {

#ifndef LDS_N_EPI_WARPS
#define LDS_N_EPI_WARPS 8      // 16 only in experiment builds (debug-knob library; the TMA epilogue's tile rings do not fit then)
#endif
constexpr int TBM = 128, TBK = 64, N_EPI_WARPS = LDS_N_EPI_WARPS, TC_THREADS = 64 + 32 * N_EPI_WARPS;
constexpr int EPI_PARTS = N_EPI_WARPS / 4;          // epilogue warps per TMEM lane quarter: they interleave the 32-column chunks
constexpr int A_SLOT_BYTES = TBM * TBK * 2;         // 16 KB: 128 rows x 128 B
constexpr int MAX_SLOTS = 8;                        // per ring
constexpr int TMEM_COLS = 512;                        // accumulator buffers: see TcParams::n_buf
constexpr int STAGE_BYTES = 32 * 32 * 4;              // per epilogue warp (and per TMA epilogue buffer): one 32 x 32 fp32 block
constexpr int BAR_BYTES = 1024;                       // mbarriers + TMEM slot
constexpr int SMEM_BUDGET = 227 * 1024 - 1024 - BAR_BYTES;    // rings + per-warp staging tiles / TMA epilogue buffers
constexpr int MAX_EPI_BUFS = 4;                       // TMA epilogue buffers per epilogue warp

struct TcParams {
  int rows, batches, tiles_per_batch, m_tiles, n_tiles, total_items;   // A: [batches][rows][parts*cin]; item = (M-tile group, N tile)
  int blk_tiling, blks_per_batch, total_blks;   // flat tiling of batched convolutions in 32-row blocks (rows % 32 == 0)
  int cin, taps, parts;
  short tap_row[24];                    // row offset of tap t relative to the output row (uniform: (t - (taps-1)/2) * dil)
  int BN, na, nw, w_slot_bytes;         // ring depths (A slots, W slots)
  int nkb, kb_per_tap;                  // K blocks of TBK per plane (all taps) / per tap
  int N;
  const float* bias;
  const float* R; int r_ld, r_div;
  void* C; int c_ld; int out_kind;      // 0 fp32, 1 bf16, 2 split bf16 (3 planes of c_ld/3 columns), 3 attention operands
  int epilogue, dbg;
  float out_scale;                      // accumulator scale (split-f16: 1 / (scale_A * scale_W)); applied before bias / activation
  float act_slope;                      // EPI_LRELU
  int small_off;                        // split-f16: TMEM column offset of the "small" accumulator (h1*w2 + h2*w1) behind the main one; 0 = none
  int n_buf, buf_stride;                // accumulator buffers in TMEM (2: epilogue of tile i overlaps the main loop of tile i+1) and their stride
  int att_parts;                        // planes of the attention operands (out_kind 3)
  int tma_epi, nbuf;                    // fp32 output through TMA (residual tile TMA-loaded, result tile TMA-stored); buffers per epilogue warp
  int red_add;                          // in-place residual (R == C): the result tile is TMA-REDUCED (+=) into C, no residual load at all
  __nv_bfloat16 *q_out, *k_out, *vt_out; int att_T, att_H, att_dpad, att_Tpad;
};

using namespace ptx;

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
// GELU(x) = 0.5 x (1 + erf(x / sqrt 2)) with erfc(z) ~ t P(t) exp(-z^2), t = 1 / (1 + 0.39 z), P of degree 5 in t (a minimax refit
// of the Abramowitz-Stegun 7.1.26 form with one more term: approximation error 8.3e-9 absolute).  One MUFU.RCP + one MUFU.EX2 + 11
// FMA-class instructions instead of the ~27 of erff (whose branch-free form costs nine FSELs per element) — the GEGLU epilogue is
// bound by exactly this arithmetic.  Evaluated in fp32 the result is within 4e-7 absolute of the exact GELU (torch's own fp32
// F.gelu: 1.2e-6), tests/test_gpu_kernels.py::test_tc_gemm_geglu_and_output_kinds; used in both precision modes.
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float ex2_approx_ftz(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  const float t = rcp_approx(fmaf(0.39f, z, 1.f));
  float poly = fmaf(-0.22753774338131172f, t, 0.8866668986234079f);
  poly = fmaf(poly, t, -0.6354043903488646f);
  poly = fmaf(poly, t, 0.6495627199979991f);
  poly = fmaf(poly, t, 0.09138708187290995f);
  poly = fmaf(poly, t, 0.23532543152903582f);
  const float e = poly * t * ex2_approx_ftz(-z * z * 1.4426950408889634f);    // 1 - erf(|x|/sqrt2)
  const float one_plus_erf = x >= 0.f ? 2.f - e : e;                    // 1 + erf(x/sqrt2)
  return 0.5f * x * one_plus_erf;
}
// adds 32 consecutive bias values (16-byte loads; the address is warp-uniform)
__device__ __forceinline__ void add_bias32(float* v, const float* bias) {
  const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 q = __ldg(b4 + i);
    v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
  }
}

// Direct path (bf16 mode): store 32 consecutive fp32 values of one output row in the requested representation (v is clobbered by the split)
__device__ __forceinline__ void store_row32(const TcParams& p, size_t row, int col, int n_out, float* v) {
  if (p.out_kind == 0) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + row * p.c_ld + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if (p.out_kind == 1) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + col);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = make_uint4(pack_pair_bf16(v[8 * i], v[8 * i + 1]), pack_pair_bf16(v[8 * i + 2], v[8 * i + 3]),
                          pack_pair_bf16(v[8 * i + 4], v[8 * i + 5]), pack_pair_bf16(v[8 * i + 6], v[8 * i + 7]));
  } else {                                  // out_kind 2: split-f16 planes [h1 | h2] of the scaled value (planes.cuh)
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + col;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= PLANE_SCALE;
#pragma unroll
    for (int pl = 0; pl < 2; ++pl) {
      uint4* dst = reinterpret_cast<uint4*>(base + (size_t)pl * n_out);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) w[i] = pl == 1 ? planes_pack_pair_f16(v[2 * i], v[2 * i + 1]) : planes_split_pair_f16(v[2 * i], v[2 * i + 1]);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
}

// ---- epilogue -------------------------------------------------------------------------------------------------------
// tcgen05.ld hands every thread ONE accumulator row (TMEM lane), so a direct store makes each warp instruction touch 32
// different rows with 16 bytes each — half-written sectors that cost the L2 as much as full ones, on the same SM <-> L2
// path the main loop is bound by.  Each epilogue warp therefore transposes its 32 x 32 block through a private 4 KB
// shared-memory tile (16-byte chunks XOR-swizzled by the row: conflict-free in both directions) and writes it back
// with 8 lanes per row: full 128-byte lines of C (fp32), full sectors of the bf16 planes, and the residual is read the
// same way.  V^T of a fused QKV projection (keys contiguous) reads the tile by columns instead.
struct EpiCtx {
  float* stage;           // this warp's [32][32] fp32 tile
  int lane;
  size_t grow0;           // global row of tile row 0 of this warp's block (b * rows + t)
  int nvalid;             // rows of the block that exist (0..32)
};

__device__ __forceinline__ void stage_write(const EpiCtx& e, const float* v) {
  float4* row = reinterpret_cast<float4*>(e.stage + e.lane * 32);
#pragma unroll
  for (int j = 0; j < 8; ++j) row[j ^ (e.lane & 7)] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
}

// Residual of this lane's eight 16-byte positions of a 32 x 32 block, requested before the accumulator is waited for so
// that its latency hides behind the TMEM load and the activation math.  Two lane layouts:
//   WIDE = false (fp32 output): lane -> columns [4j, 4j+4), j = lane & 7, rows it*4 + (lane >> 3), it = 0..7
//   WIDE = true  (bf16 outputs): lane -> columns [8j, 8j+8), j = lane & 3, rows it*8 + (lane >> 2), it = 0..3  (two per row)
struct Res8 { float4 q[8]; };
template <bool WIDE>
__device__ __forceinline__ void load_res8(const TcParams& p, const EpiCtx& e, int col, Res8& r) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int rr = WIDE ? (i >> 1) * 8 + (e.lane >> 2) : i * 4 + (e.lane >> 3);
    const int c4 = col + (WIDE ? 8 * (e.lane & 3) + 4 * (i & 1) : 4 * (e.lane & 7));
    r.q[i] = rr < e.nvalid ? __ldg(reinterpret_cast<const float4*>(p.R + ((e.grow0 + rr) / p.r_div) * p.r_ld + c4))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// fp32 output: every instruction writes four full 128-byte lines of C
__device__ __forceinline__ void stage_store_f32(const TcParams& p, const EpiCtx& e, int col, const Res8& res) {
  const int j = e.lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int rr = it * 4 + (e.lane >> 3);
    if (rr >= e.nvalid) continue;
    float4 x = reinterpret_cast<const float4*>(e.stage + rr * 32)[j ^ (rr & 7)];
    if (p.R) { x.x += res.q[it].x; x.y += res.q[it].y; x.z += res.q[it].z; x.w += res.q[it].w; }
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + (e.grow0 + rr) * p.c_ld + col + 4 * j) = x;
  }
}

// bf16 outputs (plain, 3-plane split, q / k attention operands): 16-byte stores, every instruction writes eight rows x
// 64 bytes (two full sectors each) per plane
__device__ __forceinline__ void stage_store_bf16(const TcParams& p, const EpiCtx& e, int col, int n_out, int q_region, const Res8& res) {
  const int j = e.lane & 3;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int rr = it * 8 + (e.lane >> 2);
    if (rr >= e.nvalid) continue;
    const float4* srow = reinterpret_cast<const float4*>(e.stage + rr * 32);
    float4 a = srow[(2 * j) ^ (rr & 7)], b = srow[(2 * j + 1) ^ (rr & 7)];
    if (p.R) {
      const float4 ra = res.q[2 * it], rb = res.q[2 * it + 1];
      a.x += ra.x; a.y += ra.y; a.z += ra.z; a.w += ra.w;
      b.x += rb.x; b.y += rb.y; b.z += rb.z; b.w += rb.w;
    }
    const size_t row = e.grow0 + rr;
    const int c8 = col + 8 * j;
    __nv_bfloat16* base;
    int parts, pstride;
    if (p.out_kind == 1) {
      base = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + c8; parts = 1; pstride = 0;
    } else if (p.out_kind == 2) {         // split-f16 planes [h1 | h2] of the scaled value (planes.cuh)
      __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + c8;
      a.x *= PLANE_SCALE; a.y *= PLANE_SCALE; a.z *= PLANE_SCALE; a.w *= PLANE_SCALE;
      b.x *= PLANE_SCALE; b.y *= PLANE_SCALE; b.z *= PLANE_SCALE; b.w *= PLANE_SCALE;
      uint4 w;
      w.x = planes_split_pair_f16(a.x, a.y); w.y = planes_split_pair_f16(a.z, a.w);
      w.z = planes_split_pair_f16(b.x, b.y); w.w = planes_split_pair_f16(b.z, b.w);
      *reinterpret_cast<uint4*>(o2) = w;
      w = make_uint4(planes_pack_pair_f16(a.x, a.y), planes_pack_pair_f16(a.z, a.w), planes_pack_pair_f16(b.x, b.y), planes_pack_pair_f16(b.z, b.w));
      *reinterpret_cast<uint4*>(o2 + n_out) = w;
      continue;
    } else {                              // q / k operands of the attention kernel: planes [row][parts][H*dpad]
      const int HD = p.att_H * p.att_dpad;
      base = (q_region == 0 ? p.q_out : p.k_out) + row * (size_t)(p.att_parts * HD) + (c8 - q_region * HD); parts = p.att_parts; pstride = HD;
    }
    if (p.out_kind == 3 && parts == 2) {  // split-f16 attention operands: [h1 | h2] of the scaled value
      a.x *= PLANE_SCALE; a.y *= PLANE_SCALE; a.z *= PLANE_SCALE; a.w *= PLANE_SCALE;
      b.x *= PLANE_SCALE; b.y *= PLANE_SCALE; b.z *= PLANE_SCALE; b.w *= PLANE_SCALE;
      uint4 w;
      w.x = planes_split_pair_f16(a.x, a.y); w.y = planes_split_pair_f16(a.z, a.w);
      w.z = planes_split_pair_f16(b.x, b.y); w.w = planes_split_pair_f16(b.z, b.w);
      *reinterpret_cast<uint4*>(base) = w;
      w = make_uint4(planes_pack_pair_f16(a.x, a.y), planes_pack_pair_f16(a.z, a.w), planes_pack_pair_f16(b.x, b.y), planes_pack_pair_f16(b.z, b.w));
      *reinterpret_cast<uint4*>(base + pstride) = w;
      continue;
    }
    for (int pl = 0; pl < parts; ++pl) {
      uint4 w;
      if (pl == parts - 1) {
        w = make_uint4(pack_pair_bf16(a.x, a.y), pack_pair_bf16(a.z, a.w), pack_pair_bf16(b.x, b.y), pack_pair_bf16(b.z, b.w));
      } else {
        w.x = split_pair_bf16(a.x, a.y); w.y = split_pair_bf16(a.z, a.w);
        w.z = split_pair_bf16(b.x, b.y); w.w = split_pair_bf16(b.z, b.w);
      }
      *reinterpret_cast<uint4*>(base + (size_t)pl * pstride) = w;
    }
  }
}

// q / k operands straight from the registers (direct path): planes [row][parts][H*dpad]
__device__ __forceinline__ void store_qk32(const TcParams& p, size_t row, int col, int region, float* v) {
  const int HD = p.att_H * p.att_dpad;
  __nv_bfloat16* base = (region == 0 ? p.q_out : p.k_out) + row * (size_t)(p.att_parts * HD) + (col - region * HD);
  for (int pl = 0; pl < p.att_parts; ++pl) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) w[i] = pl == p.att_parts - 1 ? pack_pair_bf16(v[2 * i], v[2 * i + 1]) : split_pair_bf16(v[2 * i], v[2 * i + 1]);
    uint4* dst = reinterpret_cast<uint4*>(base + (size_t)pl * HD);
#pragma unroll
    for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
  }
}

// V^T of a fused QKV projection (planes [B][parts][H][dpad][T_pad], frames contiguous): written straight from the
// accumulator registers — lanes hold consecutive frames, so every store instruction is one fully written 64-byte segment.
__device__ __forceinline__ void store_vt32(const TcParams& p, size_t row, int col, float* v) {
  const int HD = p.att_H * p.att_dpad, rem = col - 2 * HD;
  const int bb = (int)(row / p.att_T), tt = (int)(row - (size_t)bb * p.att_T);
  const int hh = rem / p.att_dpad, j0 = rem - hh * p.att_dpad;
  if (p.att_parts == 2) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= PLANE_SCALE;
  }
  for (int pl = 0; pl < p.att_parts; ++pl) {
    __nv_bfloat16* dst = p.vt_out + ((size_t)((bb * p.att_parts + pl) * p.att_H + hh) * p.att_dpad + j0) * p.att_Tpad + tt;
#pragma unroll
    for (int i = 0; i < 32; i += 2) {
      const uint32_t w = p.att_parts == 2 ? (pl == 0 ? planes_split_pair_f16(v[i], v[i + 1]) : planes_pack_pair_f16(v[i], v[i + 1]))
                                          : (pl == p.att_parts - 1 ? pack_pair_bf16(v[i], v[i + 1]) : split_pair_bf16(v[i], v[i + 1]));
      dst[(size_t)i * p.att_Tpad] = __ushort_as_bfloat16((unsigned short)(w & 0xffffu));
      dst[(size_t)(i + 1) * p.att_Tpad] = __ushort_as_bfloat16((unsigned short)(w >> 16));
    }
  }
}

// Accumulator read-out: 32 columns of the main accumulator; in split-f16 mode plus the same columns of the "small" accumulator
// (h1*w2 + h2*w1, kept apart so that its tiny terms are not truncated against the large h1*w1 sums), times out_scale.
__device__ __forceinline__ void acc_finish32(const TcParams& p, uint32_t taddr, float* v) {
  if (p.small_off) {
    float s[32];
    tmem_ld32(taddr + p.small_off, s);
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (v[i] + s[i]) * p.out_scale;
  } else if (p.out_scale != 1.f) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] *= p.out_scale;
  }
}

// Epilogue of this warp's 32 accumulator rows over its share of the BN columns [n0, n0+BN) (32-column chunks part,
// part + EPI_PARTS, ...): + bias, SiLU / GEGLU, + fp32 residual, store in the requested representation.
// trow = TMEM address of column n0 of this thread's lane.
template <bool STAGED>
__device__ __forceinline__ void epilogue_block(const TcParams& p, uint32_t trow, int n0, const EpiCtx& e, int part, bool skip_store) {
  // Split mode: the main loop is bound by the SM <-> L2 path, so the coalesced (staged) stores pay.  bf16 mode: main
  // loops are short and the epilogue sits on the critical path, so the lowest-latency direct stores win (measured
  // in situ: staged -2 % / +10 % step time in split / bf16 mode).
  constexpr bool staged = STAGED;
  const bool row_ok = e.lane < e.nvalid;
  const size_t grow = e.grow0 + (row_ok ? e.lane : 0);
  const float* rrow = p.R ? p.R + (grow / p.r_div) * p.r_ld : nullptr;
  if (p.epilogue == EPI_GEGLU) {          // 128-column groups of [64 value | 64 gate]
    const int n_out = p.N >> 1, units = p.BN >> 6;   // (group, 32-column chunk) units of the tile
#pragma unroll 1
    for (int u = part; u < units; u += EPI_PARTS) {
      const int grp = u >> 1, c = u & 1;
      const int col_v = n0 + grp * 128 + c * 32, col_g = col_v + 64, oc = ((n0 + grp * 128) >> 1) + c * 32;
      float v[32], g[32];
      tmem_ld32(trow + grp * 128 + c * 32, v);
      acc_finish32(p, trow + grp * 128 + c * 32, v);
      tmem_ld32(trow + grp * 128 + 64 + c * 32, g);
      acc_finish32(p, trow + grp * 128 + 64 + c * 32, g);
      if (p.bias) { add_bias32(v, p.bias + col_v); add_bias32(g, p.bias + col_g); }
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= gelu_fast(g[i]);
      if (!staged) {
        if (!skip_store && row_ok) {
          if (rrow) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += rrow[oc + i];
          }
          store_row32(p, grow, oc, n_out, v);
        }
        continue;
      }
      stage_write(e, v);
      if (!skip_store) {
        Res8 res;
        if (p.out_kind == 0) {
          if (p.R) load_res8<false>(p, e, oc, res);
          stage_store_f32(p, e, oc, res);
        } else {
          if (p.R) load_res8<true>(p, e, oc, res);
          stage_store_bf16(p, e, oc, n_out, 0, res);
        }
      }
      __syncwarp();
    }
  } else {
    const int chunks = p.BN >> 5;
    const int HD = p.att_H * p.att_dpad;
#pragma unroll 1
    for (int c = part; c < chunks; c += EPI_PARTS) {
      const int col = n0 + c * 32;
      const int region = p.out_kind == 3 ? col / HD : 0;
      uint32_t raw[32];
      tmem_ld32_issue(trow + c * 32, raw);
      Res8 res;
      if (p.R && region != 2) {
        if (!staged) {                      // direct path: this lane's own row, 8 x 16 B
          if (row_ok) {
#pragma unroll
            for (int i = 0; i < 8; ++i) res.q[i] = __ldg(reinterpret_cast<const float4*>(rrow + col) + i);
          }
        } else if (p.out_kind == 0) load_res8<false>(p, e, col, res);
        else load_res8<true>(p, e, col, res);
      }
      tmem_ld32_wait(raw);
      float v[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
      acc_finish32(p, trow + c * 32, v);
      if (p.bias) add_bias32(v, p.bias + col);
      if (p.epilogue == EPI_SILU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
      } else if (p.epilogue == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
      } else if (p.epilogue == EPI_LRELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.act_slope;
      }
      if (region == 2) {                  // V^T: straight from the registers
        if (!skip_store && e.lane < e.nvalid) store_vt32(p, e.grow0 + e.lane, col, v);
        continue;
      }
      if (!staged) {
        if (!skip_store && row_ok) {
          if (p.R) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              v[4 * i] += res.q[i].x; v[4 * i + 1] += res.q[i].y; v[4 * i + 2] += res.q[i].z; v[4 * i + 3] += res.q[i].w;
            }
          }
          if (p.out_kind == 3) store_qk32(p, grow, col, region, v);
          else store_row32(p, grow, col, p.N, v);
        }
        continue;
      }
      stage_write(e, v);
      if (!skip_store) {
        if (p.out_kind == 0) stage_store_f32(p, e, col, res);
        else stage_store_bf16(p, e, col, p.N, region, res);
      }
      __syncwarp();
    }
  }
}

// ---- TMA epilogue (fp32 output, out_kind 0) ---------------------------------------------------------------------------
// The per-thread ld.global / st.global epilogue above is LATENCY bound where the main loop is short: every warp walks its
// 32-column chunks one after the other and each chunk waits for its own residual load (the C x C projections moved 14-22 B/clk
// per SM against a ~40 B/clk path; profiles/r02_gemm_bounds_call8.md: 0.45-0.67 of their bound).  Here each epilogue warp owns
// a small ring of 4 KB shared-memory tiles: the fp32 RESIDUAL block of a chunk is TMA-loaded into a tile ahead of time (the
// loads of the next chunks / the next output tile fly while the main loop of that tile still runs), the thread that holds a
// TMEM row adds bias + residual IN PLACE in the tile (the 128-byte TMA swizzle is the same XOR pattern the staged path used,
// so both the row-per-thread accesses and the TMA engine see conflict-free / linear data), and one elected lane hands the tile
// to a TMA STORE — no thread ever waits for a global load or issues a global store, ragged tile edges are clipped by the
// tensor map (a 3-D map [columns, frames, utterances] like the A operand's, so a block never spills into the next utterance).
struct TmaEpi {
  uint32_t buf0, bar0;     // shared-memory addresses of this warp's tiles / mbarriers
  uint32_t q, pq;          // chunks consumed / residual loads issued so far (ring position = count % nbuf)
  int p_item, p_c;         // prefetch iterator: next (item, chunk) whose residual has not been requested yet
};

__device__ __forceinline__ void chunk_coords(const TcParams& p, int item, int c, int crank, int quarter, int& col, int& t, int& b) {
  const int it = item / p.n_tiles, nt = item - it * p.n_tiles, mt = it * 2 + crank;
  col = nt * p.BN + c * 32;
  if (p.blk_tiling) {
    const int gb = mt * 4 + quarter;
    b = gb < p.total_blks ? gb / p.blks_per_batch : p.batches;          // ghost block: out of bounds (zero fill / nothing stored)
    t = gb < p.total_blks ? (gb - b * p.blks_per_batch) * 32 : 0;
  } else {
    b = mt < p.m_tiles ? mt / p.tiles_per_batch : p.batches;
    t = mt < p.m_tiles ? (mt - b * p.tiles_per_batch) * TBM + quarter * 32 : 0;
  }
}

// lane 0 of an epilogue warp: request the residual block of the next chunk of this warp's sequence
__device__ __forceinline__ void tma_epi_prefetch(const TcParams& p, const CUtensorMap* mapR, TmaEpi& st, int crank, int quarter, int part,
                                                 int ncl) {
  if (st.p_item >= p.total_items) return;
  int col, t, b;
  chunk_coords(p, st.p_item, st.p_c, crank, quarter, col, t, b);
  const uint32_t k = st.pq % (uint32_t)p.nbuf;
  mbar_expect_tx(st.bar0 + 8u * k, STAGE_BYTES);
  tma_load_3d(st.buf0 + k * STAGE_BYTES, mapR, st.bar0 + 8u * k, col, t, b);
  ++st.pq;
  st.p_c += EPI_PARTS;
  if (st.p_c >= (p.BN >> 5)) { st.p_c = part; st.p_item += ncl; }
}

__device__ __forceinline__ void epilogue_tma(const TcParams& p, const CUtensorMap* mapC, const CUtensorMap* mapR, TmaEpi& st, uint32_t trow,
                                             int item, int crank, int quarter, int part, int lane, int nvalid, int ncl, uint8_t* smem_gen,
                                             uint32_t smem_gen_addr) {
  const int chunks = p.BN >> 5;
#pragma unroll 1
  for (int c = part; c < chunks; c += EPI_PARTS) {
    const uint32_t k = st.q % (uint32_t)p.nbuf;
    const uint32_t buf = st.buf0 + k * STAGE_BYTES;
    int col, t, b;
    chunk_coords(p, item, c, crank, quarter, col, t, b);
    uint32_t raw[32];
    tmem_ld32_issue(trow + c * 32, raw);
    if (p.R) {
      mbar_wait(st.bar0 + 8u * k, (st.q / (uint32_t)p.nbuf) & 1u);          // the residual block has landed in this tile
    } else {
      if (lane == 0) {                                                       // the store that last used this tile (nbuf chunks ago) has read it
        if (p.nbuf == 2) bulk_wait_read<1>(); else if (p.nbuf == 3) bulk_wait_read<2>(); else bulk_wait_read<3>();
      }
      __syncwarp();
    }
    tmem_ld32_wait(raw);
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(raw[i]);
    acc_finish32(p, trow + c * 32, v);
    if (p.bias) add_bias32(v, p.bias + col);
    if (p.epilogue == EPI_SILU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
    } else if (p.epilogue == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_fast(v[i]);
    } else if (p.epilogue == EPI_LRELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * p.act_slope;
    }
    float4* row = reinterpret_cast<float4*>(smem_gen + (buf - smem_gen_addr) + lane * 128);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 x = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
      const int jj = j ^ (lane & 7);                                         // TMA 128-byte swizzle of an 8-row x 128 B atom
      if (p.R) { const float4 r = row[jj]; x.x += r.x; x.y += r.y; x.z += r.z; x.w += r.w; }
      row[jj] = x;
    }
    fence_async_smem();
    __syncwarp();
    ++st.q;
    if (lane == 0) {
      if (nvalid > 0) {
        if (p.red_add) tma_reduce_add_3d(mapC, buf, col, t, b);
        else tma_store_3d(mapC, buf, col, t, b);
      }
      bulk_commit();
      if (p.R) {
        bulk_wait_read<1>();        // every store but the one just committed has read its tile: the tile of the previous chunk is free
        tma_epi_prefetch(p, mapR, st, crank, quarter, part, ncl);
      }
    }
  }
}

template <bool STAGED>
__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const __grid_constant__ CUtensorMap mapC,
               const __grid_constant__ CUtensorMap mapR, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;              // 1024 B alignment required by the 128B swizzle atoms
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_ring = base, w_ring = base + p.na * A_SLOT_BYTES;
  const uint32_t ring_bytes = p.na * A_SLOT_BYTES + p.nw * p.w_slot_bytes;
  const uint32_t bar_base = base + ring_bytes;    // a_full[8], a_empty[8], w_full[8], w_empty[8], tmem_full[2], tmem_empty[2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + ring_bytes + 8 * (4 * MAX_SLOTS + 4));
  const uint32_t epi_bar_base = bar_base + 8u * (4 * MAX_SLOTS + 8);        // [N_EPI_WARPS][MAX_EPI_BUFS] residual-tile barriers
  float* stage_base = reinterpret_cast<float*>(smem + ring_bytes + BAR_BYTES);   // 1024-byte aligned (slots are multiples of 1 KB)
  const uint32_t stage_addr = base + ring_bytes + BAR_BYTES;
  auto a_full = [&](int s) { return bar_base + 8u * s; };
  auto a_empty = [&](int s) { return bar_base + 8u * (MAX_SLOTS + s); };
  auto w_full = [&](int s) { return bar_base + 8u * (2 * MAX_SLOTS + s); };
  auto w_empty = [&](int s) { return bar_base + 8u * (3 * MAX_SLOTS + s); };
  auto tmem_full_bar = [&](int b) { return bar_base + 8u * (4 * MAX_SLOTS + b); };
  auto tmem_empty_bar = [&](int b) { return bar_base + 8u * (4 * MAX_SLOTS + 2 + b); };

  const int warp = warp_id_uniform(), lane = threadIdx.x & 31;   // provably warp-uniform role branches
  pdl_trigger();                                 // the next kernel may start launching behind our last wave
  // CTA pair: rank 0 (leader) issues every MMA for both SMs.  full barriers live in the leader and count the bytes
  // of both CTAs' loads; empty / tmem_full barriers exist in both CTAs and are signalled by the leader's multicast
  // commits; tmem_empty lives in the leader and collects one arrive per epilogue warp of both CTAs.
  const uint32_t crank = (uint32_t)__shfl_sync(0xffffffffu, (int)cluster_ctarank(), 0);   // uniform for the compiler too
  const bool leader = crank == 0;
  const int cid = blockIdx.x >> 1, ncl = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapW);
    for (int s = 0; s < 4 * MAX_SLOTS; ++s) mbar_init(bar_base + 8u * s, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(tmem_full_bar(b), 1);
      mbar_init(tmem_empty_bar(b), 2 * N_EPI_WARPS);
    }
    for (int s = 0; s < N_EPI_WARPS * MAX_EPI_BUFS; ++s) mbar_init(epi_bar_base + 8u * s, 1);
    if (p.tma_epi) { prefetch_tensormap(&mapC); if (p.R) prefetch_tensormap(&mapR); }
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // the peer's barriers are initialised before any remote signal
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                    // everything above overlapped the previous kernel's tail; its data is needed from here on

  const bool split = p.parts == 2;               // split-f16: planes h1, h2 of both operands, three products per K block

  if (warp == 0) {
    {   // the whole warp runs the producer loop in lock step (uniform operands); one elected lane issues each TMA load (tc_ptx.cuh)
      int sa = 0, sw = 0;                        // next slot of each ring
      uint32_t pha = 0, phw = 0;
      const int w_row = (int)crank * (p.BN >> 1);
      // Batched convolutions whose utterance length is a multiple of 32 frames are tiled over the FLAT sequence of
      // 32-frame blocks (a 128-row tile = four blocks, possibly of two utterances; each block is its own TMA box, so the
      // zero padding at the utterance edges is still out-of-bounds fill): 864-frame utterances then fill 432 tiles
      // instead of 7 x 64 = 448, which is 3 rounds of the 74 CTA pairs instead of 4.
      int blk_b[4], blk_t[4];
      auto load_a = [&](int col, int row, int b) {
        mbar_wait(a_empty(sa), pha ^ 1u);
        if (leader) mbar_expect_tx_elect(a_full(sa), 2 * A_SLOT_BYTES);
        const uint32_t dst = a_ring + sa * A_SLOT_BYTES, bar = mapa_u32(a_full(sa), 0);
        if (p.blk_tiling) {
#pragma unroll
          for (int q = 0; q < 4; ++q) tma_load_3d_pair_elect(dst + q * (A_SLOT_BYTES / 4), &mapA, bar, col, blk_t[q] + row, blk_b[q]);
        } else {
          tma_load_3d_pair_elect(dst, &mapA, bar, col, row, b);
        }
        if (++sa == p.na) { sa = 0; pha ^= 1u; }
      };
      auto load_w = [&](int col, int n0) {
        mbar_wait(w_empty(sw), phw ^ 1u);
        if (leader) mbar_expect_tx_elect(w_full(sw), 2 * p.w_slot_bytes);
        tma_load_2d_pair_elect(w_ring + sw * p.w_slot_bytes, &mapW, mapa_u32(w_full(sw), 0), col, n0 + w_row);
        if (++sw == p.nw) { sw = 0; phw ^= 1u; }
      };
      for (int item = cid; item < p.total_items; item += ncl) {
        const int nt = item % p.n_tiles, mt = (item / p.n_tiles) * 2 + (int)crank;
        int b = mt / p.tiles_per_batch;
        int t0 = (mt - b * p.tiles_per_batch) * TBM;
        if (p.blk_tiling) {                      // rows passed to load_a become offsets (tap - pad) relative to each block
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int gb = mt * 4 + q;
            blk_b[q] = gb < p.total_blks ? gb / p.blks_per_batch : p.batches;   // ghost block: out of bounds -> zeros
            blk_t[q] = (gb % p.blks_per_batch) * 32;
          }
          b = 0; t0 = 0;
        }
        const int n0 = nt * p.BN;
        int tap = 0, cb = 0;
        for (int kb = 0; kb < p.nkb; ++kb) {
          const int row = t0 + p.tap_row[tap], acol = cb * TBK, wcol = tap * p.parts * p.cin + cb * TBK;
          if (split) {                           // need order of the MMA warp: A_h1, W_h2 | A_h2, W_h1
            load_a(acol, row, b);
            load_w(wcol + p.cin, n0);
            load_a(p.cin + acol, row, b);
            load_w(wcol, n0);
          } else {
            load_a(acol, row, b);
            load_w(wcol, n0);
          }
          if (++cb == p.kb_per_tap) { cb = 0; ++tap; }
        }
      }
    }
  } else if (warp == 1) {
    if (leader) {   // the whole warp runs the issue loop in lock step; one elected lane issues each MMA / commit (tc_ptx.cuh)
      const uint32_t idesc = split ? umma_idesc_f16(2 * TBM, p.BN) : umma_idesc_bf16(2 * TBM, p.BN);
      int sa = 0, sw = 0;                        // oldest live slot of each ring
      uint32_t pha = 0, phw = 0;
      int local = 0;
      auto next_a = [&]() { if (++sa == p.na) { sa = 0; pha ^= 1u; } };
      auto next_w = [&]() { if (++sw == p.nw) { sw = 0; phw ^= 1u; } };
      for (int item = cid; item < p.total_items; item += ncl, ++local) {
        const int buf = p.n_buf == 2 ? (local & 1) : 0;
        const uint32_t use = p.n_buf == 2 ? ((uint32_t)local >> 1) : (uint32_t)local;     // uses of this buffer so far
        mbar_wait(tmem_empty_bar(buf), (use & 1u) ^ 1u);   // both epilogues have drained this buffer
        tc_fence_after();
        const uint32_t acc_main = tmem_base + buf * p.buf_stride, acc_small = acc_main + p.small_off;
        uint32_t acc_m = 0, acc_s = 0;             // accumulate flags of the two accumulators
        auto product = [&](int a_slot, int w_slot, uint32_t acc, uint32_t& accumulate) {
          const uint64_t ad = umma_desc_kmajor(a_ring + a_slot * A_SLOT_BYTES, 128);
          const uint64_t wd = umma_desc_kmajor(w_ring + w_slot * p.w_slot_bytes, 128);
#pragma unroll
          for (int k = 0; k < TBK / 16; ++k) {   // +32 B (16 elements) along K inside the swizzle atom per UMMA_K step
            umma_bf16_pair_elect(acc, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), idesc, accumulate);
            accumulate = 1;
          }
        };
        for (int kb = 0; kb < p.nkb; ++kb) {
          if (split) {
            // slots (ring order): A: sa -> h1, sa+1 -> h2 ; W: sw -> h2, sw+1 -> h1.  The two small products go to their own
            // accumulator: their terms are 2^-11 of the main ones and would be truncated against the large h1*w1 partial sums
            // by the tensor core's fp32 accumulate (the epilogue adds the two accumulators once, in fp32).
            const int a_h1 = sa, w_h2 = sw;
            mbar_wait(a_full(a_h1), pha);
            mbar_wait(w_full(w_h2), phw);
            tc_fence_after();
            product(a_h1, w_h2, acc_small, acc_s);         // h1 * w2
            umma_commit_pair_elect(w_empty(w_h2), 3);
            next_a();
            next_w();
            const int a_h2 = sa, w_h1 = sw;
            mbar_wait(a_full(a_h2), pha);
            mbar_wait(w_full(w_h1), phw);
            tc_fence_after();
            product(a_h2, w_h1, acc_small, acc_s);         // h2 * w1
            umma_commit_pair_elect(a_empty(a_h2), 3);
            product(a_h1, w_h1, acc_main, acc_m);          // h1 * w1
            umma_commit_pair_elect(a_empty(a_h1), 3);
            umma_commit_pair_elect(w_empty(w_h1), 3);
            next_a();
            next_w();
          } else {
            mbar_wait(a_full(sa), pha);
            mbar_wait(w_full(sw), phw);
            tc_fence_after();
            product(sa, sw, acc_main, acc_m);
            umma_commit_pair_elect(a_empty(sa), 3);
            umma_commit_pair_elect(w_empty(sw), 3);
            next_a();
            next_w();
          }
        }
        umma_commit_pair_elect(tmem_full_bar(buf), 3);
      }
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32); the EPI_PARTS warps of a quarter interleave the chunks ----
    const int quarter = warp & 3, part = (warp - 2) >> 2;
    EpiCtx e;
    e.stage = stage_base + (warp - 2) * (STAGE_BYTES / 4);
    e.lane = lane;
    TmaEpi te;
    te.buf0 = stage_addr + (uint32_t)((warp - 2) * p.nbuf) * STAGE_BYTES;
    te.bar0 = epi_bar_base + 8u * (uint32_t)((warp - 2) * MAX_EPI_BUFS);
    te.q = te.pq = 0u;
    te.p_item = cid; te.p_c = part;
    if (p.tma_epi && p.R && lane == 0 && !p.dbg)   // residual blocks of the first nbuf - 1 chunks
      for (int i = 0; i + 1 < p.nbuf; ++i) tma_epi_prefetch(p, &mapR, te, (int)crank, quarter, part, ncl);
    int local = 0;
    for (int item = cid; item < p.total_items; item += ncl, ++local) {
      const int nt = item % p.n_tiles, mt = (item / p.n_tiles) * 2 + (int)crank;
      // the last pair may carry a ghost M tile (TMA zero fill, no stores)
      if (p.blk_tiling) {                        // this warp's 32 rows are block mt*4 + quarter of the flat block sequence
        const int gb = mt * 4 + quarter;
        e.nvalid = gb < p.total_blks ? 32 : 0;
        e.grow0 = (size_t)(gb < p.total_blks ? gb : 0) * 32;
      } else {
        const int b = mt / p.tiles_per_batch;
        const int t_blk = (mt - b * p.tiles_per_batch) * TBM + quarter * 32;     // first frame of this warp's 32 rows
        e.nvalid = mt < p.m_tiles ? max(0, min(32, p.rows - t_blk)) : 0;
        e.grow0 = (size_t)(mt < p.m_tiles ? b : 0) * p.rows + t_blk;
      }
      const int buf = p.n_buf == 2 ? (local & 1) : 0;
      mbar_wait(tmem_full_bar(buf), (p.n_buf == 2 ? ((uint32_t)local >> 1) : (uint32_t)local) & 1u);
      tc_fence_after();
      const uint32_t trow = tmem_base + buf * p.buf_stride + ((uint32_t)(quarter * 32) << 16);
#ifdef LDS_DEBUG_KNOBS   // timing experiments only (results are wrong by construction); never compiled into the product library
      if (p.dbg & 2) {
      } else if (p.tma_epi && !(p.dbg & 1))
        epilogue_tma(p, &mapC, &mapR, te, trow, item, (int)crank, quarter, part, lane, e.nvalid, ncl, reinterpret_cast<uint8_t*>(stage_base), stage_addr);
      else
        epilogue_block<STAGED>(p, trow, nt * p.BN, e, part, (p.dbg & 1) != 0);
#else
      if (p.tma_epi)
        epilogue_tma(p, &mapC, &mapR, te, trow, item, (int)crank, quarter, part, lane, e.nvalid, ncl, reinterpret_cast<uint8_t*>(stage_base), stage_addr);
      else
        epilogue_block<STAGED>(p, trow, nt * p.BN, e, part, false);
#endif
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(tmem_empty_bar(buf), 0));
    }
    if (p.tma_epi && lane == 0) bulk_wait<0>();   // every TMA store of this warp has completed before the CTA may exit
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();                                // no CTA exits (or frees TMEM) while the pair still works on it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------------------
// host side: tensor-map construction (driver entry point fetched through the runtime; no libcuda link dependency)
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                   const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                   CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr; uint64_t d0, d1, d2; uint32_t box1;
  bool operator==(const MapKey& o) const { return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && d2 == o.d2 && box1 == o.box1; }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    for (uint64_t v : {k.d0, k.d1, k.d2, (uint64_t)k.box1}) h = h * 1000003u ^ (size_t)v;
    return h;
  }
};

// bf16 tensor [d2][d1][d0] (d0 contiguous), box = TBK x box1 x 1, 128B swizzle, zero OOB fill.
cudaError_t get_map(const void* ptr, uint64_t d0, uint64_t d1, uint64_t d2, uint32_t box1, CUtensorMap* out) {
  static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> cache;
  static std::mutex mu;
  const MapKey key{ptr, d0, d1, d2, box1};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) {
    *out = it->second;
    return cudaSuccess;
  }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  CUtensorMap m;
  CUresult r;
  if (d2 == 0) {
    const cuuint64_t dims[2] = {d0, d1};
    const cuuint64_t strides[1] = {d0 * 2};
    const cuuint32_t box[2] = {TBK, box1};
    const cuuint32_t es[2] = {1, 1};
    r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    const cuuint32_t box[3] = {TBK, box1, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return cudaSuccess;
}

// fp32 tensor [d2][d1][d0 valid columns, row stride ld], box = 32 columns x 32 rows x 1, 128B swizzle (TMA epilogue tiles)
cudaError_t get_map_f32(const void* ptr, uint64_t d0, uint64_t ld, uint64_t d1, uint64_t d2, CUtensorMap* out) {
  struct Key {
    const void* ptr; uint64_t d0, ld, d1, d2;
    bool operator==(const Key& o) const { return ptr == o.ptr && d0 == o.d0 && ld == o.ld && d1 == o.d1 && d2 == o.d2; }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr);
      for (uint64_t v : {k.d0, k.ld, k.d1, k.d2}) h = h * 1000003u ^ (size_t)v;
      return h;
    }
  };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  const Key key{ptr, d0, ld, d1, d2};
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return cudaSuccess; }
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t strides[2] = {ld * 4, ld * 4 * d1};
  const cuuint32_t box[3] = {32, 32, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUtensorMap m;
  const CUresult r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  *out = m;
  return cudaSuccess;
}

int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& v = n[dev & 63];
  if (!v && (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v < 1)) v = 148;
  return v;
}

// co-resident clusters of two persistent CTAs (GPC boundaries can strand an SM; asked from the occupancy calculator)
int max_clusters2() {
  static int n[64] = {0};
  int dev = 0;
  cudaGetDevice(&dev);
  int& slot = n[dev & 63];
  if (!slot) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * sm_count());
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = 227 * 1024;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int v = 0;
    if (cudaOccupancyMaxActiveClusters(&v, gemm_tc_kernel<true>, &cfg) != cudaSuccess || v < 1) v = sm_count() / 2;
    (void)cudaGetLastError();
    slot = v;
  }
  return slot;
}

}  // namespace

// generic bf16 tiled tensor map (rank 2 or 3) with a 32/64/128-byte swizzle and zero out-of-bounds fill
cudaError_t tc_make_map_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                             int swizzle_bytes, CUtensorMap* out) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  if (rank < 2 || rank > 3) return cudaErrorInvalidValue;
  cuuint64_t d[3], st[2];
  cuuint32_t bx[3], es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  const CUtensorMapSwizzle sw = swizzle_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                                : (swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
  const CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr), d, st, bx, es,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// cached form (attention operand maps: the same scratch buffers and shapes recur every evaluation)
cudaError_t tc_make_map_bf16_cached(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                                    int swizzle_bytes, CUtensorMap* out) {
  struct Key {
    const void* ptr; int rank, swz; uint64_t d[3], st[2]; uint32_t bx[3];
    bool operator==(const Key& o) const {
      if (ptr != o.ptr || rank != o.rank || swz != o.swz) return false;
      for (int i = 0; i < 3; ++i) if (d[i] != o.d[i] || bx[i] != o.bx[i]) return false;
      return st[0] == o.st[0] && st[1] == o.st[1];
    }
  };
  struct KeyHash {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.ptr) ^ ((size_t)k.rank << 1) ^ ((size_t)k.swz << 8);
      for (int i = 0; i < 3; ++i) h = (h * 1000003u ^ (size_t)k.d[i]) * 1000003u ^ (size_t)k.bx[i];
      return (h * 1000003u ^ (size_t)k.st[0]) * 1000003u ^ (size_t)k.st[1];
    }
  };
  static std::unordered_map<Key, CUtensorMap, KeyHash> cache;
  static std::mutex mu;
  if (rank < 2 || rank > 3) return cudaErrorInvalidValue;
  Key key{ptr, rank, swizzle_bytes, {0, 0, 0}, {0, 0}, {0, 0, 0}};
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.st[i] = strides_bytes[i];
  std::lock_guard<std::mutex> lock(mu);
  auto it = cache.find(key);
  if (it != cache.end()) { *out = it->second; return cudaSuccess; }
  const cudaError_t e = tc_make_map_bf16(ptr, rank, dims, strides_bytes, box, swizzle_bytes, out);
  if (e != cudaSuccess) return e;
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, *out);
  return cudaSuccess;
}

cudaError_t launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s) {
  if (a.batches <= 0 || a.rows <= 0 || a.N <= 0) return cudaSuccess;
  const bool split = a.a_parts == 2 && a.w_parts == 2 && a.n_pairs == 3;      // split-f16 (fp32-accurate)
  const bool plain = a.a_parts == 1 && a.w_parts == 1 && a.n_pairs == 1;
  if (a.cin % 64 || a.N % 64 || (a.N % 128 && (a.epilogue == EPI_GEGLU || a.out_kind == 3)) || a.taps < 1 || (a.tap_rows ? a.taps > 24 : (a.taps > 11 || a.taps % 2 == 0)) || a.dil < 1 || (!split && !plain) || (a.out_kind != 3 && a.c_ld % 8) ||
      (a.R && a.r_ld % 4) || a.r_div < 1)
    return cudaErrorInvalidValue;
  if (a.out_kind == 3 && (!a.q_out || !a.k_out || !a.vt_out || a.att_T < 1 || a.att_dpad % 32 || a.N != 3 * a.att_H * a.att_dpad ||
                          a.epilogue != EPI_NONE))
    return cudaErrorInvalidValue;
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
  }
  TcParams p;
  p.BN = a.N % 256 == 0 ? 256 : ((a.N % 192 == 0 && a.epilogue != EPI_GEGLU) ? 192 : (a.N % 128 == 0 ? 128 : 64));   // GEGLU groups are 128 wide;
  // BN = 64 (N = 64 + 128 j: the vocoder's 64-channel level): an N = 64 instruction costs the same 92 cycles as N = 128
  {
    // Small problems (single utterances, the deep levels of small batches) leave most CTA pairs idle: narrower N tiles double
    // the number of items, and an N = 128 instruction costs 92 cycles against 128 for N = 256, so every item is also 28 %
    // shorter (one 5 s utterance, 20 evaluations: 152 -> 127.5 ms).  The accumulation order per output element does not change:
    // results are bit-identical to the wide tiling, batch-composition invariance is untouched.
    // (Split-K across clusters was built and measured for the same case — partial tiles in an L2 workspace, last arriver sums
    // in slice order and writes the sum back to TMEM: parity-green but SLOWER, 139.8 vs 127.5 ms at B=1 and 185.6 vs 175.0 ms at
    // B=8: the finishing CTA reads S x 128 KB through one SM, about one K block of work per slice.  Removed.)
    const int blks = a.rows / 32 * a.batches;
    const bool flat = a.batches > 1 && a.rows % 32 == 0 && a.rows % TBM != 0;
    const int m_tiles = flat ? (blks + 3) / 4 : ((a.rows + TBM - 1) / TBM) * a.batches;
    if (p.BN == 256 && (a.N / 256) * ((m_tiles + 1) / 2) * 4 <= max_clusters2()) p.BN = 128;
    // split-f16: main + small accumulator of a 256-column tile fill the 512 TMEM columns, so the epilogue of a tile cannot
    // overlap the next tile's main loop; with 128-column tiles two buffers fit.  Short K loops (<= 8 K blocks)
    // take the narrow, double-buffered form, long ones the wide one.
    // take the narrow, double-buffered form, long ones the wide one (thresholds 0 / 4 / 8 / 16 K blocks measured 534.0 / 531.0 /
    // 529.3 / 531.6 ms per headline step).
    if (split && p.BN == 256 && a.taps * (a.cin / TBK) <= 8) p.BN = 128;
  }
  p.w_slot_bytes = (p.BN / 2) * TBK * 2;          // each CTA of the pair holds half of the W tile
  // fp32 outputs (with or without an fp32 residual) go through the TMA epilogue
  const bool aligned16 = (reinterpret_cast<uintptr_t>(a.C) & 15) == 0 && (!a.R || (reinterpret_cast<uintptr_t>(a.R) & 15) == 0);
  p.tma_epi = (a.out_kind == 0 && a.epilogue != EPI_GEGLU && (!a.R || a.r_div == 1) && a.c_ld % 4 == 0 && (!a.R || a.r_ld % 4 == 0) &&
               aligned16 && knobs().tma_epi) ? 1 : 0;
  // (three tiles per warp with shallower operand rings, and an L2 prefetch of the residual blocks at tile start, measured the
  // same as two tiles: 115.1-116.0 k frames/s for all four combinations, GPU call 26 of round 2.  Also measured and rejected,
  // profiles/r02_gemm_experiments_call133.md: an L2 prefetch stream of the A operand 4 / 8 K blocks ahead of the ring loads
  // (cp.async.bulk.prefetch.tensor) — 8-90 % SLOWER per shape: the main loop is bound by L2 -> SM throughput, not latency, and
  // the prefetches compete for it; A ring of 4 instead of 6 slots: +2 %; the staged epilogue instead of TMA for long K: +5 %.
  // bf16 mode, GPU call 141: the staged epilogue for the fused QKV projection 228.4 k -> 209.9 k frames/s, for every 16-bit output
  // 190.4 k: the direct stores stay.)
  p.nbuf = split ? 2 : 3;
  p.na = (split || !p.tma_epi) ? 6 : 4;
#ifdef LDS_DEBUG_KNOBS
  {
    static const int e_na = getenv("LDS_TC_NA") ? atoi(getenv("LDS_TC_NA")) : 0,
                     e_nbuf = getenv("LDS_TC_NBUF") ? atoi(getenv("LDS_TC_NBUF")) : 0,
                     e_maxkb = getenv("LDS_TC_TMAEPI_MAXKB") ? atoi(getenv("LDS_TC_TMAEPI_MAXKB")) : 1 << 30;
    if (a.taps * (a.cin / TBK) > e_maxkb) p.tma_epi = 0;
    if (e_nbuf) p.nbuf = e_nbuf;
    if (e_na) p.na = e_na;
  }
#endif
  const int stage_bytes = p.tma_epi ? N_EPI_WARPS * p.nbuf * STAGE_BYTES : (split ? N_EPI_WARPS * STAGE_BYTES : 0);
  p.nw = (SMEM_BUDGET - stage_bytes - p.na * A_SLOT_BYTES) / p.w_slot_bytes;
  if (p.nw > MAX_SLOTS) p.nw = MAX_SLOTS;
  if (!split && p.tma_epi && p.nw > 4) p.nw = 4;
  CUtensorMap mA, mW, mC, mR;
  p.blk_tiling = (a.batches > 1 && a.rows % 32 == 0 && a.rows % TBM != 0) ? 1 : 0;
  p.blks_per_batch = a.rows / 32; p.total_blks = p.blks_per_batch * a.batches;
  cudaError_t e = get_map(a.A, (uint64_t)a.a_parts * a.cin, (uint64_t)a.rows, (uint64_t)a.batches, p.blk_tiling ? 32 : TBM, &mA);
  if (e != cudaSuccess) return e;
  p.m_tiles = p.blk_tiling ? (p.total_blks + 3) / 4 : ((a.rows + TBM - 1) / TBM) * a.batches;
  const int csize = 2;                            // always a CTA pair; an odd M tile count leaves one ghost tile
  e = get_map(a.W, (uint64_t)a.taps * a.w_parts * a.cin, (uint64_t)a.N, 0, (uint32_t)(p.BN / csize), &mW);
  if (e != cudaSuccess) return e;
  p.rows = a.rows; p.batches = a.batches; p.tiles_per_batch = (a.rows + TBM - 1) / TBM;
  p.n_tiles = a.N / p.BN; p.total_items = p.n_tiles * ((p.m_tiles + csize - 1) / csize);
  p.cin = a.cin; p.taps = a.taps; p.parts = a.a_parts;
  for (int t = 0; t < 24; ++t) p.tap_row[t] = t < a.taps ? (short)(a.tap_rows ? a.tap_rows[t] : (t - (a.taps - 1) / 2) * a.dil) : 0;
  p.kb_per_tap = a.cin / TBK; p.nkb = a.taps * p.kb_per_tap;
  p.N = a.N; p.bias = a.bias; p.R = a.R; p.r_ld = a.r_ld; p.r_div = a.r_div;
  p.C = a.C; p.c_ld = a.c_ld; p.out_kind = a.out_kind; p.epilogue = a.epilogue;
#ifdef LDS_DEBUG_KNOBS
  { static const int dbg = getenv("LDS_TC_DEBUG") ? atoi(getenv("LDS_TC_DEBUG")) : 0; p.dbg = dbg; }
#else
  p.dbg = 0;
#endif
  p.out_scale = a.out_scale; p.act_slope = a.act_slope; p.att_parts = a.att_parts;
  // in-place residual (attention out-projection, conv2 behind a shortcut, the Whisper blocks): C += tile through a TMA reduce-add store
  // instead of TMA-load R / add / TMA-store — a quarter less L2 traffic on these launches and no residual barrier in the epilogue;
  // (acc + bias) + C rounds once either way, and every element is reduced exactly once: bit-identical, deterministic
  p.red_add = (p.tma_epi && a.R && a.R == (const float*)a.C && a.r_ld == a.c_ld && knobs().red_add) ? 1 : 0;
  if (p.red_add) p.R = nullptr;
  // TMEM: main accumulator (BN columns) [+ small accumulator (BN columns) in split-f16 mode] per buffer; two buffers when they fit
  p.small_off = split ? p.BN : 0;
  p.buf_stride = split ? 2 * p.BN : p.BN;
  p.n_buf = 2 * p.buf_stride <= TMEM_COLS ? 2 : 1;
  if (p.n_buf == 2) p.buf_stride = TMEM_COLS / 2;
  p.q_out = a.q_out; p.k_out = a.k_out; p.vt_out = a.vt_out;
  p.att_T = a.att_T; p.att_H = a.att_H; p.att_dpad = a.att_dpad; p.att_Tpad = a.att_Tpad;
  const size_t smem = (size_t)p.na * A_SLOT_BYTES + (size_t)p.nw * p.w_slot_bytes + 1024 + BAR_BYTES + stage_bytes;
  if (p.tma_epi) {
    e = get_map_f32(a.C, (uint64_t)a.N, (uint64_t)a.c_ld, (uint64_t)a.rows, (uint64_t)a.batches, &mC);
    if (e != cudaSuccess) return e;
    if (a.R) {
      e = get_map_f32(a.R, (uint64_t)a.N, (uint64_t)a.r_ld, (uint64_t)a.rows, (uint64_t)a.batches, &mR);
      if (e != cudaSuccess) return e;
    } else {
      mR = mC;
    }
  } else {
    mC = mA; mR = mA;            // unused by the kernel
  }
  const int max_cl = max_clusters2();
  const int ncl = p.total_items < max_cl ? p.total_items : max_cl;
  return split ? launch_pdl(gemm_tc_kernel<true>, dim3(ncl * csize), dim3(TC_THREADS), smem, s, csize, mA, mW, mC, mR, p)
               : launch_pdl(gemm_tc_kernel<false>, dim3(ncl * csize), dim3(TC_THREADS), smem, s, csize, mA, mW, mC, mR, p);
}

}  // namespace lds
