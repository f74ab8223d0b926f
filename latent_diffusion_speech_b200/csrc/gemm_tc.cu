// gemm_tc.cu — tcgen05 / TMEM / TMA implicit-GEMM for Conv1d(k=3), Conv1d(k=1) and nn.Linear.
//
// One kernel serves both precision modes of the library:
//   * LDS_PREC_BF16 : A and W are bf16, one tcgen05.mma (kind::f16, fp32 accumulate in TMEM) per K slice.
//   * LDS_PREC_FP32 : "split-bf16" — every fp32 operand is stored as three bf16 planes (hi, mid, lo with
//     hi+mid+lo == x to 24 bits); the K loop runs the six significant plane products
//     (lo*hi, hi*lo, mid*mid, mid*hi, hi*mid, hi*hi — smallest first) into the same fp32 TMEM accumulator,
//     which recovers fp32-level accuracy on the tensor pipe at 6 MMAs per logical K slice.
//
// Structure (B200, sm_100a): CTA tile 128(M) x 128(N) x 64(K); 6 warps, warp-specialised:
//   warp 0      TMA producer  — cp.async.bulk.tensor (3-D map over [channels, frames, utterances] for A so that the
//                               three taps of a k=3 convolution are three shifted loads of the same tensor and the
//                               zero padding is TMA out-of-bounds fill; 2-D map for W), 128B swizzle, mbarrier tx
//   warp 1      MMA issuer    — tcgen05.alloc (128 TMEM columns), one elected lane issues tcgen05.mma.cta_group::1
//                               and tcgen05.commit to release smem stages / publish the accumulator
//   warps 2..5  epilogue      — tcgen05.ld 32x32b (one accumulator row per thread), + bias, SiLU / GEGLU, + fp32
//                               residual, store fp32 / bf16 / 3-plane split bf16
// 3-stage smem ring (96 KB) so that two CTAs are resident per SM: one CTA's epilogue overlaps the other's
// main loop.  Reference ops replaced: F.conv1d (lora.py:102), nn.Linear (attention_processor.py:1012-1040,
// attention.py:291,247), GEGLU (attention.py:299-301), residual adds (resnet.py:639, attention.py:161-201).
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "lds_kernels.h"
#include "tc_ptx.cuh"

namespace lds {
namespace {

constexpr int TBM = 128, TBN = 128, TBK = 64, TSTAGES = 3, TC_THREADS = 192;
constexpr int A_STAGE_BYTES = TBM * TBK * 2, W_STAGE_BYTES = TBN * TBK * 2;
constexpr int TC_SMEM_BYTES = TSTAGES * (A_STAGE_BYTES + W_STAGE_BYTES) + 1024 + 256;
constexpr int TMEM_COLS = 128;

struct TcParams {
  int rows, batches, tiles_per_batch;   // A: [batches][rows][a_parts*cin]
  int cin, taps, pad;
  int w_parts, n_pairs;
  int pair_a[6], pair_w[6];
  int N;
  const float* bias;
  const float* R; int r_ld, r_div;
  void* C; int c_ld; int out_kind;      // 0 fp32, 1 bf16, 2 split bf16 (3 planes of c_ld/3 columns), 3 attention operands
  int epilogue;
  __nv_bfloat16 *q_out, *k_out, *vt_out; int att_T, att_H, att_dpad, att_Tpad;
};

using namespace ptx;
constexpr uint32_t kIdesc = umma_idesc_bf16(TBM, TBN);
__device__ __forceinline__ uint64_t umma_smem_desc(uint32_t saddr) { return umma_desc_kmajor(saddr, 128); }

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

// store 32 consecutive fp32 values of one output row in the requested representation
__device__ __forceinline__ void store_row32(const TcParams& p, size_t row, int col, int n_out, const float* v) {
  if (p.out_kind == 0) {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.C) + row * p.c_ld + col);
#pragma unroll
    for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else if (p.out_kind == 1) {
    uint4* dst = reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + col);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      dst[i] = make_uint4(pack_bf16(v[8 * i], v[8 * i + 1]), pack_bf16(v[8 * i + 2], v[8 * i + 3]),
                          pack_bf16(v[8 * i + 4], v[8 * i + 5]), pack_bf16(v[8 * i + 6], v[8 * i + 7]));
  } else {
    float r[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = v[i];
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(p.C) + row * p.c_ld + col;
#pragma unroll
    for (int pl = 0; pl < 3; ++pl) {
      uint4* dst = reinterpret_cast<uint4*>(base + (size_t)pl * n_out);
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __nv_bfloat16 a = __float2bfloat16_rn(r[2 * i]), b = __float2bfloat16_rn(r[2 * i + 1]);
        r[2 * i] -= __bfloat162float(a);
        r[2 * i + 1] -= __bfloat162float(b);
        w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  }
}

// out_kind 3: 32 consecutive columns of the fused [q | k | v] projection (one head, one of q/k/v) as attention operands
__device__ __forceinline__ void store_qkv32(const TcParams& p, size_t row, int col, const float* v) {
  const int parts = p.w_parts, HD = p.att_H * p.att_dpad;
  const int region = col / HD, rem = col - region * HD;
  float r[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) r[i] = v[i];
  if (region < 2) {
    __nv_bfloat16* base = (region == 0 ? p.q_out : p.k_out) + row * (size_t)(parts * HD) + rem;
    for (int pl = 0; pl < parts; ++pl) {
      uint32_t w[16];
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const __nv_bfloat16 a = __float2bfloat16_rn(r[2 * i]), b = __float2bfloat16_rn(r[2 * i + 1]);
        r[2 * i] -= __bfloat162float(a);
        r[2 * i + 1] -= __bfloat162float(b);
        w[i] = (uint32_t)__bfloat16_as_ushort(a) | ((uint32_t)__bfloat16_as_ushort(b) << 16);
      }
      uint4* dst = reinterpret_cast<uint4*>(base + (size_t)pl * HD);
#pragma unroll
      for (int i = 0; i < 4; ++i) dst[i] = make_uint4(w[4 * i], w[4 * i + 1], w[4 * i + 2], w[4 * i + 3]);
    }
  } else {
    const int bb = (int)(row / p.att_T), tt = (int)(row - (size_t)bb * p.att_T);
    const int hh = rem / p.att_dpad, j0 = rem - hh * p.att_dpad;
    for (int pl = 0; pl < parts; ++pl) {
      __nv_bfloat16* dst = p.vt_out + ((size_t)((bb * parts + pl) * p.att_H + hh) * p.att_dpad + j0) * p.att_Tpad + tt;
#pragma unroll
      for (int i = 0; i < 32; ++i) {          // lanes hold consecutive frames: each store is a coalesced 64 B row segment
        const __nv_bfloat16 a = __float2bfloat16_rn(r[i]);
        r[i] -= __bfloat162float(a);
        dst[(size_t)i * p.att_Tpad] = a;
      }
    }
  }
}

// Epilogue of one accumulator row (this thread's TMEM lane) over the 128 columns [n0, n0+128): + bias, SiLU / GEGLU,
// + fp32 residual, store in the requested representation.  trow = TMEM address of column n0 of this lane.
__device__ __forceinline__ void epilogue_128(const TcParams& p, uint32_t trow, int n0, size_t grow, bool valid) {
    const float* rrow = p.R ? p.R + (grow / p.r_div) * p.r_ld : nullptr;
    if (p.epilogue == EPI_GEGLU) {
      const int n_out = p.N >> 1;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        float v[32], g[32];
        tmem_ld32(trow + c * 32, v);
        tmem_ld32(trow + 64 + c * 32, g);
        const int col_v = n0 + c * 32, col_g = n0 + 64 + c * 32, oc = (n0 >> 1) + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float a = v[i], gg = g[i];
          if (p.bias) { a += __ldg(p.bias + col_v + i); gg += __ldg(p.bias + col_g + i); }
          v[i] = a * gelu_erf(gg);
        }
        if (valid) {
          if (rrow) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] += rrow[oc + i];
          }
          store_row32(p, grow, oc, n_out, v);
        }
      }
    } else {
#pragma unroll 1
      for (int c = 0; c < TBN / 32; ++c) {
        float v[32];
        tmem_ld32(trow + c * 32, v);
        const int col = n0 + c * 32;
        if (p.bias) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += __ldg(p.bias + col + i);
        }
        if (p.epilogue == EPI_SILU) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = silu_f(v[i]);
        }
        if (valid) {
          if (rrow) {
            const float4* r4 = reinterpret_cast<const float4*>(rrow + col);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 q = r4[i];
              v[4 * i] += q.x; v[4 * i + 1] += q.y; v[4 * i + 2] += q.z; v[4 * i + 3] += q.w;
            }
          }
          if (p.out_kind == 3) store_qkv32(p, grow, col, v);
          else store_row32(p, grow, col, p.N, v);
        }
      }
    }
}

__global__ void __launch_bounds__(TC_THREADS, 2)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapW, const TcParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;              // 1024 B alignment required by the 128B swizzle atoms
  uint8_t* smem = smem_raw + (base - raw);
  const uint32_t a_base = base, w_base = base + TSTAGES * A_STAGE_BYTES;
  const uint32_t bar_base = w_base + TSTAGES * W_STAGE_BYTES;   // full[S], empty[S], tmem_full
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + TSTAGES * (A_STAGE_BYTES + W_STAGE_BYTES) + 8 * (2 * TSTAGES + 1));
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (TSTAGES + s); };
  const uint32_t tmem_full_bar = bar_base + 8u * (2 * TSTAGES);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * TBN;
  const int b = blockIdx.y / p.tiles_per_batch;
  const int t0 = (blockIdx.y - b * p.tiles_per_batch) * TBM;

  if (warp == 0 && lane == 0) {
    prefetch_tensormap(&mapA);
    prefetch_tensormap(&mapW);
    for (int s = 0; s < TSTAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    mbar_fence_init();
  } else if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = *tmem_slot;

  const int kblocks_per_tap = p.cin / TBK;
  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int pr = 0; pr < p.n_pairs; ++pr) {
        const int a_col0 = p.pair_a[pr] * p.cin;
        for (int tap = 0; tap < p.taps; ++tap) {
          const int w_col0 = (tap * p.w_parts + p.pair_w[pr]) * p.cin;
          for (int cb = 0; cb < kblocks_per_tap; ++cb, ++it) {
            const int s = it % TSTAGES;
            const uint32_t ph = (uint32_t)(it / TSTAGES) & 1u;
            mbar_wait(empty_bar(s), ph ^ 1u);
            mbar_expect_tx(full_bar(s), A_STAGE_BYTES + W_STAGE_BYTES);
            tma_load_3d(a_base + s * A_STAGE_BYTES, &mapA, full_bar(s), a_col0 + cb * TBK, t0 - p.pad + tap, b);
            tma_load_2d(w_base + s * W_STAGE_BYTES, &mapW, full_bar(s), w_col0 + cb * TBK, n0);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const int total = p.n_pairs * p.taps * kblocks_per_tap;
      for (int it = 0; it < total; ++it) {
        const int s = it % TSTAGES;
        const uint32_t ph = (uint32_t)(it / TSTAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint64_t ad = umma_smem_desc(a_base + s * A_STAGE_BYTES);
        const uint64_t wd = umma_smem_desc(w_base + s * W_STAGE_BYTES);
#pragma unroll
        for (int k = 0; k < TBK / 16; ++k)   // +32 B (16 bf16) along K inside the swizzle atom per UMMA_K step
          umma_bf16(tmem_acc, ad + (uint64_t)(2 * k), wd + (uint64_t)(2 * k), kIdesc, (it > 0 || k > 0) ? 1u : 0u);
        umma_commit(empty_bar(s));
      }
      umma_commit(tmem_full_bar);
    }
  } else {
    // ---- epilogue: warp w may touch TMEM lanes [32*(w%4), 32*(w%4)+32) ----
    const int quarter = warp & 3;
    const int r_in_tile = quarter * 32 + lane;
    const int t = t0 + r_in_tile;
    const bool valid = t < p.rows;
    const size_t grow = (size_t)b * p.rows + (valid ? t : 0);
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    epilogue_128(p, tmem_acc + ((uint32_t)(quarter * 32) << 16), n0, grow, valid);
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, TMEM_COLS);
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

// bf16 tensor [d2][d1][d0] (d0 contiguous), box = 64 x box1 x 1, 128B swizzle, zero OOB fill.
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
    const cuuint32_t box[2] = {64, box1};
    const cuuint32_t es[2] = {1, 1};
    r = fn(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  } else {
    const cuuint64_t dims[3] = {d0, d1, d2};
    const cuuint64_t strides[2] = {d0 * 2, d0 * d1 * 2};
    const cuuint32_t box[3] = {64, box1, 1};
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

cudaError_t launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s) {
  if (a.batches <= 0 || a.rows <= 0 || a.N <= 0) return cudaSuccess;
  if (a.cin % TBK || a.N % TBN || (a.taps != 1 && a.taps != 3) || a.n_pairs < 1 || a.n_pairs > 6 || a.a_parts < 1 ||
      a.w_parts < 1 || (a.out_kind != 3 && a.c_ld % 8) || (a.R && a.r_ld % 4) || a.r_div < 1)
    return cudaErrorInvalidValue;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  CUtensorMap mA, mW;
  cudaError_t e = get_map(a.A, (uint64_t)a.a_parts * a.cin, (uint64_t)a.rows, (uint64_t)a.batches, TBM, &mA);
  if (e != cudaSuccess) return e;
  e = get_map(a.W, (uint64_t)a.taps * a.w_parts * a.cin, (uint64_t)a.N, 0, TBN, &mW);
  if (e != cudaSuccess) return e;
  TcParams p;
  p.rows = a.rows; p.batches = a.batches; p.tiles_per_batch = (a.rows + TBM - 1) / TBM;
  p.cin = a.cin; p.taps = a.taps; p.pad = a.taps == 3 ? 1 : 0;
  p.w_parts = a.w_parts; p.n_pairs = a.n_pairs;
  for (int i = 0; i < 6; ++i) { p.pair_a[i] = a.pair_a[i]; p.pair_w[i] = a.pair_w[i]; }
  p.N = a.N; p.bias = a.bias; p.R = a.R; p.r_ld = a.r_ld; p.r_div = a.r_div;
  p.C = a.C; p.c_ld = a.c_ld; p.out_kind = a.out_kind; p.epilogue = a.epilogue;
  p.q_out = a.q_out; p.k_out = a.k_out; p.vt_out = a.vt_out;
  p.att_T = a.att_T; p.att_H = a.att_H; p.att_dpad = a.att_dpad; p.att_Tpad = a.att_Tpad;
  if (a.out_kind == 3 && (!a.q_out || !a.k_out || !a.vt_out || a.att_T < 1 || a.att_dpad % 32 || a.N != 3 * a.att_H * a.att_dpad ||
                          a.epilogue != EPI_NONE))
    return cudaErrorInvalidValue;
  dim3 grid(a.N / TBN, p.tiles_per_batch * a.batches);
  gemm_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, s>>>(mA, mW, p);
  return cudaGetLastError();
}

}  // namespace lds
