// attention_pair.cu — CTA-pair (tcgen05.mma.cta_group::2) flash attention for head dim <= 32 in split (fp32-accurate) mode.
//
// STATUS: measured experiment, off by default (LDS_ATT_PAIR=1 enables it; tests/test_gpu_kernels.py covers it that way).
// It is parity-green but ~8 % SLOWER than attention_tc.cu: the hypothesis below is wrong in one point — a cta_group::2
// instruction keeps the tensor pipes of BOTH SMs busy for its ~92 cycles (tests/micro/bench_umma.cu, "PAIR" rows), so
// the issue cost per query row AND SM does not drop; pairs only save operand bytes per SM (what gemm_tc.cu uses them
// for).  Kept as the record of that measurement and as a working cta_group::2 attention skeleton.
//
// Why (hypothesis): with d = 32 every P*V instruction has N <= 96 and every Q*K^T instruction N <= 128, and one tcgen05.mma costs
// ~92 cycles whatever N <= 128 (tests/micro/bench_umma.cu) — attention_tc.cu is bound by tensor-pipe ISSUE at ~40 % of the
// pipe's rate (ncu: the softmax warps wait for S).  A cta_group::2 instruction covers 256 query rows (128 per SM) for
// the same issue cost, so a CTA pair halves the instructions per query.  Same algorithm as attention_tc.cu (two passes:
// row max from the hi*hi product, then exp / P*V with all six plane products; S and O in TMEM), same operand layouts
// (written by the fused QKV projection of gemm_tc.cu), replaces F.scaled_dot_product_attention
// (diffusion/unet1d/attention_processor.py:1032-1034).
//
// Pair layout.  The pair handles 256 consecutive queries of one (utterance, head); CTA r owns queries [128 r, 128 r + 128):
// its Q planes (A operand rows), its S / O accumulator rows in its own TMEM, its softmax threads and its P planes.  The B
// operand of a pair MMA is split by ROWS between the two CTAs (first N/2 rows from rank 0, last N/2 from rank 1), and an
// instruction names ONE shared-memory offset for both, so every N-batched product gets its own region, filled per CTA:
//   K tile (64 keys), SW64 rows of 32 dims:   R1 64 rows: rank0 K_hi        | rank1 K_mid        -> Q_hi, Q_mid x [K_hi;K_mid]  N=128
//                                             R2 32 rows: rank0 K_hi[0:32]  | rank1 K_hi[32:64]  -> Q_lo x K_hi (and pass A)   N=64
//                                             R3 32 rows: rank0 K_lo[0:32]  | rank1 K_lo[32:64]  -> Q_hi x K_lo                N=64
//   V^T tile (64 keys), SW128 rows of 64 keys: RV1 48 rows: rank0 V_hi, V_mid[0:16] | rank1 V_mid[16:32], V_lo -> P_hi  N=96
//                                             RV2 32 rows: rank0 V_hi              | rank1 V_mid              -> P_mid N=64
//                                             RV3 16 rows: rank0 V_hi[0:16]        | rank1 V_hi[16:32]        -> P_lo  N=32
// The accumulator columns come out in stacked order ([.K_hi | .K_mid], [V_hi | V_mid | V_lo]), exactly the blocks the
// softmax / epilogue threads of attention_tc.cu add up.
// Barriers: "full" barriers (q, k, v) live in the leader and count the TMA bytes of both CTAs; k/v "empty", "S full" and
// "PV done" exist in both CTAs and are signalled by the leader's multicast commits; "S free" and "P full" live in the
// leader and collect one arrive per softmax warp of both CTAs (remote for the peer).
// Two pairs share an SM pair (256 TMEM columns and ~100 KB of shared memory per CTA), so the softmax of one overlaps
// the tensor work of the other.
#include "lds_kernels.h"
#include "tc_ptx.cuh"
#include <math.h>

namespace lds {

cudaError_t tc_make_map_bf16(const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes, const uint32_t* box,
                             int swizzle_bytes, CUtensorMap* out);

namespace {
using namespace ptx;

constexpr int AQ = 128, AKV = 64, DPAD = 32, PARTS = 3, THREADS = 320, NSOFT_WARPS = 8;
constexpr int QB = AQ * DPAD * 2;                 // one Q plane tile: 128 rows x 64 B
constexpr int K_R1 = 0, K_R2 = 64 * 64, K_R3 = K_R2 + 32 * 64, K_STAGE = K_R3 + 32 * 64;        // 8 KB
constexpr int V_R1 = 0, V_R2 = 48 * 128, V_R3 = V_R2 + 32 * 128, V_STAGE = V_R3 + 16 * 128;     // 12 KB
constexpr int PBK = AQ * 128;                     // one P plane: 128 rows x 128 B (64 keys)
constexpr int NK = 2, NV = 1;
constexpr int S_COLS = 128, O_COL = 128, TMEM_COLS = 256;     // S: blk0 [0,64) blk1 [64,128); O: [128, 224)
constexpr int SMEM_BYTES = PARTS * QB + NK * K_STAGE + NV * V_STAGE + PARTS * PBK + 1024 + 256 + 2 * 128 * 4 + 64;

struct PairParams {
  int T, H, d, C;
  float scale;
  __nv_bfloat16* out;   // planes [B*T][parts*C]
};

__device__ __forceinline__ void softmax_bar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void mbar_arrive_cluster_release(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}

__global__ void __launch_bounds__(THREADS, 2)
attention_pair_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapVT, const PairParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t q_s = base, k_s = q_s + PARTS * QB, v_s = k_s + NK * K_STAGE, p_s = v_s + NV * V_STAGE;
  const uint32_t bar0 = p_s + PARTS * PBK;
  uint8_t* p_ptr = smem + (p_s - base);
  const uint32_t q_full = bar0;
  auto s_full = [&](int s) { return bar0 + 8 + 8u * s; };
  auto s_free = [&](int s) { return bar0 + 24 + 8u * s; };
  const uint32_t p_full = bar0 + 40, pv_done = bar0 + 48;
  auto k_full = [&](int s) { return bar0 + 56 + 8u * s; };
  auto k_empty = [&](int s) { return bar0 + 72 + 8u * s; };
  const uint32_t v_full = bar0 + 88, v_empty = bar0 + 96;
  uint8_t* misc = smem + (bar0 - base) + 128;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc);
  float* red = reinterpret_cast<float*>(misc + 16);          // [2][128] exchange of row max / row sum between the halves

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t crank = cluster_ctarank();
  const bool leader = crank == 0;
  const int q0 = (int)blockIdx.x * AQ, h = blockIdx.y, b = blockIdx.z;     // blockIdx.x = 2 * pair + rank
  const int nt = (p.T + AKV - 1) / AKV;
  const int HD = p.H * DPAD;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapQ);
    prefetch_tensormap(&mapK);
    prefetch_tensormap(&mapVT);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) { mbar_init(s_full(s), 1); mbar_init(s_free(s), 2 * NSOFT_WARPS); }
    mbar_init(p_full, 2 * NSOFT_WARPS);
    mbar_init(pv_done, 1);
    for (int s = 0; s < NK; ++s) { mbar_init(k_full(s), 1); mbar_init(k_empty(s), 1); }
    mbar_init(v_full, 1);
    mbar_init(v_empty, 1);
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc_pair(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();
  tc_fence_after();
  const uint32_t tmem0 = *tmem_slot;
  const uint32_t tmem_o = tmem0 + O_COL;

  if (warp == 0) {
    if (lane == 0) {
      // every load completes on the LEADER's barrier; the leader arms it with the bytes of both CTAs
      const uint32_t qf = mapa_u32(q_full, 0);
      if (leader) mbar_expect_tx(q_full, 2 * PARTS * QB);
      for (int pl = 0; pl < PARTS; ++pl) tma_load_3d_pair(q_s + pl * QB, &mapQ, qf, pl * HD + h * DPAD, q0, b);
      const int r = (int)crank;
      for (int it = 0; it < 2 * nt; ++it) {
        const bool pass_a = it < nt;
        const int jt = pass_a ? it : it - nt;
        const int ks = it % NK;
        const uint32_t kdst = k_s + ks * K_STAGE, kf = mapa_u32(k_full(ks), 0);
        const int key0 = jt * AKV, hcol = h * DPAD;
        mbar_wait(k_empty(ks), ((uint32_t)(it / NK) & 1u) ^ 1u);
        if (leader) mbar_expect_tx(k_full(ks), 2 * (pass_a ? 32 * 64 : K_STAGE));
        tma_load_3d_pair(kdst + K_R2, &mapK, kf, hcol, key0 + 32 * r, b);                          // K_hi half (pass A: all it needs)
        if (!pass_a) {
          tma_load_3d_pair(kdst + K_R1, &mapK, kf, r * HD + hcol, key0, b);                        // rank0 K_hi, rank1 K_mid
          tma_load_3d_pair(kdst + K_R1 + 32 * 64, &mapK, kf, r * HD + hcol, key0 + 32, b);
          tma_load_3d_pair(kdst + K_R3, &mapK, kf, 2 * HD + hcol, key0 + 32 * r, b);               // K_lo half
          // V^T rows of plane pl, dims [j0, j0+16): global row ((b*parts + pl)*H + h)*DPAD + j0
          const uint32_t vf = mapa_u32(v_full, 0);
          auto vrow = [&](int pl, int j0) { return ((b * PARTS + pl) * p.H + h) * DPAD + j0; };
          mbar_wait(v_empty, ((uint32_t)jt & 1u) ^ 1u);
          if (leader) mbar_expect_tx(v_full, 2 * V_STAGE);
          if (r == 0) {
            tma_load_2d_pair(v_s + V_R1, &mapVT, vf, key0, vrow(0, 0));
            tma_load_2d_pair(v_s + V_R1 + 2048, &mapVT, vf, key0, vrow(0, 16));
            tma_load_2d_pair(v_s + V_R1 + 4096, &mapVT, vf, key0, vrow(1, 0));
            tma_load_2d_pair(v_s + V_R2, &mapVT, vf, key0, vrow(0, 0));
            tma_load_2d_pair(v_s + V_R2 + 2048, &mapVT, vf, key0, vrow(0, 16));
            tma_load_2d_pair(v_s + V_R3, &mapVT, vf, key0, vrow(0, 0));
          } else {
            tma_load_2d_pair(v_s + V_R1, &mapVT, vf, key0, vrow(1, 16));
            tma_load_2d_pair(v_s + V_R1 + 2048, &mapVT, vf, key0, vrow(2, 0));
            tma_load_2d_pair(v_s + V_R1 + 4096, &mapVT, vf, key0, vrow(2, 16));
            tma_load_2d_pair(v_s + V_R2, &mapVT, vf, key0, vrow(1, 0));
            tma_load_2d_pair(v_s + V_R2 + 2048, &mapVT, vf, key0, vrow(1, 16));
            tma_load_2d_pair(v_s + V_R3, &mapVT, vf, key0, vrow(0, 16));
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      const uint32_t id128 = umma_idesc_bf16(2 * AQ, 128), id96 = umma_idesc_bf16(2 * AQ, 96), id64 = umma_idesc_bf16(2 * AQ, 64),
                     id32 = umma_idesc_bf16(2 * AQ, 32);
      const uint64_t q_hi = umma_desc_kmajor(q_s, 64), q_mid = umma_desc_kmajor(q_s + QB, 64), q_lo = umma_desc_kmajor(q_s + 2 * QB, 64);
      const uint64_t p_hi = umma_desc_kmajor(p_s, 128), p_mid = umma_desc_kmajor(p_s + PBK, 128), p_lo = umma_desc_kmajor(p_s + 2 * PBK, 128);
      const uint64_t v1 = umma_desc_kmajor(v_s + V_R1, 128), v2 = umma_desc_kmajor(v_s + V_R2, 128), v3 = umma_desc_kmajor(v_s + V_R3, 128);
      uint32_t su0 = 0u, su1 = 0u;
      mbar_wait(q_full, 0);
      auto issue_qk = [&](int it) {
        const bool pass_a = it < nt;
        const int ks = it % NK;
        const int sb = pass_a ? (it & 1) : 0;
        const uint32_t sdst = tmem0 + (pass_a && sb ? O_COL : 0);      // pass A alternates with the idle O columns
        mbar_wait(k_full(ks), (uint32_t)(it / NK) & 1u);
        mbar_wait(s_free(sb), ((sb ? su1 : su0) & 1u) ^ 1u);
        if (sb) ++su1; else ++su0;
        tc_fence_after();
        const uint32_t kb = k_s + ks * K_STAGE;
        const uint64_t r1 = umma_desc_kmajor(kb + K_R1, 64), r2 = umma_desc_kmajor(kb + K_R2, 64), r3 = umma_desc_kmajor(kb + K_R3, 64);
        if (pass_a) {
#pragma unroll
          for (int k = 0; k < DPAD / 16; ++k) umma_bf16_pair(sdst, q_hi + 2 * k, r2 + 2 * k, id64, k == 0 ? 0u : 1u);
        } else {
#pragma unroll
          for (int k = 0; k < DPAD / 16; ++k) {
            umma_bf16_pair(sdst, q_hi + 2 * k, r1 + 2 * k, id128, k == 0 ? 0u : 1u);       // blk0 += hi*hi, blk1 += hi*mid
            umma_bf16_pair(sdst, q_mid + 2 * k, r1 + 2 * k, id128, 1u);                    // blk0 += mid*hi, blk1 += mid*mid
            umma_bf16_pair(sdst, q_lo + 2 * k, r2 + 2 * k, id64, 1u);                      // blk0 += lo*hi
            umma_bf16_pair(sdst + AKV, q_hi + 2 * k, r3 + 2 * k, id64, 1u);                // blk1 += hi*lo
          }
        }
        umma_commit_pair(k_empty(ks), 3);
        umma_commit_pair(s_full(sb), 3);
      };
      for (int it = 0; it < nt; ++it) issue_qk(it);            // pass A
      issue_qk(nt);                                            // pass B: Q K^T runs one tile ahead of P V
      for (int jb = 0; jb < nt; ++jb) {
        if (jb + 1 < nt) issue_qk(nt + jb + 1);
        mbar_wait(v_full, (uint32_t)jb & 1u);
        mbar_wait(p_full, (uint32_t)jb & 1u);
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < AKV / 16; ++k) {
          umma_bf16_pair(tmem_o, p_hi + 2 * k, v1 + 2 * k, id96, (jb == 0 && k == 0) ? 0u : 1u);   // O blk0..2 += P_hi x [V_hi|V_mid|V_lo]
          umma_bf16_pair(tmem_o, p_mid + 2 * k, v2 + 2 * k, id64, 1u);                             // O blk0..1 += P_mid x [V_hi|V_mid]
          umma_bf16_pair(tmem_o, p_lo + 2 * k, v3 + 2 * k, id32, 1u);                              // O blk0    += P_lo x V_hi
        }
        umma_commit_pair(v_empty, 3);
        umma_commit_pair(pv_done, 3);
      }
    }
  } else {
    const int quarter = warp & 3, half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;                 // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const int c0 = half * (AKV / 2);                     // this thread's 32 key columns inside a tile
    const float sl2 = p.scale * 1.4426950408889634f;    // softmax scale * log2(e)
    const uint32_t s_free_leader[2] = {mapa_u32(s_free(0), 0), mapa_u32(s_free(1), 0)};
    const uint32_t p_full_leader = mapa_u32(p_full, 0);
    float m = -INFINITY;
    uint32_t su0 = 0u, su1 = 0u;
    // ---- pass A: row maximum of the hi*hi scores ----
    for (int it = 0; it < nt; ++it) {
      const int sb = it & 1;
      mbar_wait(s_full(sb), (sb ? su1 : su0) & 1u);
      if (sb) ++su1; else ++su0;
      tc_fence_after();
      float s[32];
      tmem_ld32(tmem0 + lane_off + (sb ? O_COL : 0) + c0, s);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(sb ? s_free_leader[1] : s_free_leader[0]);
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
    m *= sl2;
    red[half * 128 + row] = m;
    softmax_bar();
    m = fmaxf(m, red[(half ^ 1) * 128 + row]);
    softmax_bar();
    // ---- pass B: probabilities, row sum, P planes to shared memory ----
    float l = 0.f;
    uint8_t* prow = p_ptr + (row >> 3) * 1024 + (row & 7) * 128;     // 8-row / 1024 B swizzle atoms, 128 B per row
    for (int jb = 0; jb < nt; ++jb) {
      mbar_wait(s_full(0), su0 & 1u);
      ++su0;
      tc_fence_after();
      float s[32];
      {
        float t[32];
        tmem_ld32(tmem0 + lane_off + AKV + c0, s);                     // blk1
        tmem_ld32(tmem0 + lane_off + c0, t);                           // blk0
#pragma unroll
        for (int i = 0; i < 32; ++i) s[i] += t[i];
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(s_free_leader[0]);
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
      uint32_t w[PARTS][16];
#pragma unroll
      for (int pl = 0; pl < PARTS; ++pl) {
#pragma unroll
        for (int i = 0; i < 16; ++i) w[pl][i] = (pl == PARTS - 1) ? pack_pair_bf16(s[2 * i], s[2 * i + 1]) : split_pair_bf16(s[2 * i], s[2 * i + 1]);
      }
      if (jb >= 1) mbar_wait(pv_done, (uint32_t)(jb - 1) & 1u);        // the P*V that last read the P buffer is done
#pragma unroll
      for (int pl = 0; pl < PARTS; ++pl) {
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {                               // 16-byte chunk = 8 keys
          const int key_chunk = (c0 >> 3) + cc;
          *reinterpret_cast<uint4*>(prow + pl * PBK + (((key_chunk & 7) ^ (row & 7)) << 4)) =
              make_uint4(w[pl][4 * cc], w[pl][4 * cc + 1], w[pl][4 * cc + 2], w[pl][4 * cc + 3]);
        }
      }
      fence_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster_release(p_full_leader);
    }
    red[half * 128 + row] = l;
    softmax_bar();
    l += red[(half ^ 1) * 128 + row];
    // ---- epilogue: O / l -> bf16 planes; this thread writes output columns [half*16, half*16 + 16) ----
    mbar_wait(pv_done, (uint32_t)(nt - 1) & 1u);
    tc_fence_after();
    const int q = q0 + row;
    const float inv = 1.f / l;
    constexpr int OC = DPAD / 2;
    float o[OC];
    {
      uint32_t r16[16];
      auto ld16 = [&](uint32_t addr, float* v) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
            : "=r"(r16[0]), "=r"(r16[1]), "=r"(r16[2]), "=r"(r16[3]), "=r"(r16[4]), "=r"(r16[5]), "=r"(r16[6]), "=r"(r16[7]),
              "=r"(r16[8]), "=r"(r16[9]), "=r"(r16[10]), "=r"(r16[11]), "=r"(r16[12]), "=r"(r16[13]), "=r"(r16[14]), "=r"(r16[15])
            : "r"(addr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r16[i]);
      };
      ld16(tmem_o + lane_off + 2 * DPAD + half * OC, o);               // smallest block first
      for (int blk = 1; blk >= 0; --blk) {
        float t[OC];
        ld16(tmem_o + lane_off + blk * DPAD + half * OC, t);
#pragma unroll
        for (int i = 0; i < OC; ++i) o[i] += t[i];
      }
    }
    const int ncol = min(OC, p.d - half * OC);
    if (q < p.T && ncol > 0) {
      __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q) * (size_t)(PARTS * p.C) + h * p.d + half * OC;
#pragma unroll
      for (int i = 0; i < OC; ++i) o[i] *= inv;
      for (int pl = 0; pl < PARTS; ++pl) {
        uint32_t w[OC / 2];
#pragma unroll
        for (int i = 0; i < OC / 2; ++i) w[i] = split_pair_bf16(o[2 * i], o[2 * i + 1]);
        uint4* dst = reinterpret_cast<uint4*>(orow + (size_t)pl * p.C);
#pragma unroll
        for (int i = 0; i < OC / 8; ++i)
          if (i * 8 < ncol) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  cluster_sync();                                // no CTA exits (or frees TMEM) while the pair still works on it
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem0, TMEM_COLS);
  }
}

}  // namespace

// d <= 32, split mode.  Operand layouts as attention_tc.cu (a.dpad == 32, a.parts == 3).
cudaError_t launch_attention_pair(const AttnTcArgs& a, cudaStream_t s) {
  if (a.parts != PARTS || a.dpad != DPAD || a.d > DPAD || a.d % 8 || a.T_pad % 8 || a.T_pad < a.T) return cudaErrorInvalidValue;
  static unsigned long long configured = 0;
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(attention_pair_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return e;
  }
  const uint64_t HD = (uint64_t)a.H * DPAD;
  CUtensorMap mQ, mK, mV;
  cudaError_t e;
  {
    const uint64_t dims[3] = {(uint64_t)PARTS * HD, (uint64_t)a.T, (uint64_t)a.B};
    const uint64_t str[2] = {(uint64_t)PARTS * HD * 2, (uint64_t)PARTS * HD * 2 * a.T};
    const uint32_t boxq[3] = {(uint32_t)DPAD, (uint32_t)AQ, 1}, boxk[3] = {(uint32_t)DPAD, 32u, 1};
    if ((e = tc_make_map_bf16(a.q, 3, dims, str, boxq, 64, &mQ)) != cudaSuccess) return e;
    if ((e = tc_make_map_bf16(a.k, 3, dims, str, boxk, 64, &mK)) != cudaSuccess) return e;
  }
  {
    const uint64_t dims[2] = {(uint64_t)a.T, (uint64_t)a.B * PARTS * HD};
    const uint64_t str[1] = {(uint64_t)a.T_pad * 2};
    const uint32_t box[2] = {64u, 16u};
    if ((e = tc_make_map_bf16(a.vt, 2, dims, str, box, 128, &mV)) != cudaSuccess) return e;
  }
  PairParams p;
  p.T = a.T; p.H = a.H; p.d = a.d; p.C = a.H * a.d;
  p.scale = 1.0f / sqrtf((float)a.d);
  p.out = a.out;
  const int pairs = (a.T + 2 * AQ - 1) / (2 * AQ);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(2 * pairs, a.H, a.B);
  cfg.blockDim = dim3(THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, attention_pair_kernel, mQ, mK, mV, p);
}

}  // namespace lds
