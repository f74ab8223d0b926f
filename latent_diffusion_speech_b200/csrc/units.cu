// units.cu — the units front-end: the step BEFORE the diffusion sampler, audio -> units[B,T,1280].
//
// Replaces, behind the lds_units_* entry points of include/lds_b200.h (file:line relative to the reference tree):
//   encoder/whisper/audio.py:60-80     log_mel_spectrogram: hann-windowed STFT (n_fft 400, hop 160, centre / reflect padding), power
//                                      spectrum, mel filterbank, log10, clamp to (global max - 8), (x + 4) / 4
//   encoder/whisper/model.py:112-131   AudioEncoder.forward: conv1 k3 + GELU, conv2 k3 / stride 2 + GELU, + sinusoids,
//                                      n_layer x ResidualAttentionBlock (:89-110; MultiHeadAttention :43-87), ln_post
//   tools/tools.py:193-223             units_forced_alignment ('nearest' / 'left': a row gather along time)
//   quantize/kmeans_codebook.py:29-31  EuclideanCodebook.dequantize (F.embedding: the same row gather)
// The transformer runs on the sampler's own kernels: gemm_tc (tcgen05 / TMEM / TMA; split-f16 fp32-accurate or bf16 operands;
// bias + exact-erf GELU + fp32 residual in the epilogue; the k3 convolutions as implicit GEMMs), the fused QKV projection
// writing attention operands, attention_tc (flash attention, d = 64) and the LayerNorm kernels (norm.cu, wide-row form).
// New here: the log-mel kernels (direct DFT in fp64 — 0.5 GFLOP per 30 s of audio, and more accurate than an fp32 FFT where the
// log10 amplifies errors of quiet bins) and the row gather.  Activations are channels-last [B*T, C]; the residual stream x stays
// fp32, every GEMM A operand is written as 16-bit planes (planes.cuh) by the kernel that produces it.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/lds_b200.h"
#include "lds_kernels.h"
#include "planes.cuh"
#include "host_pack.h"

namespace {

using namespace lds;

thread_local std::string g_units_error;

int ufail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_units_error = buf;
  return code;
}

// ---- log-mel spectrogram -----------------------------------------------------------------------------------------------
constexpr int N_FFT = 400, HOP = 160, N_BINS = N_FFT / 2 + 1, FPB = 8, MEL_THREADS = 256;

// float max through integer atomics (the slot starts at -inf): non-negative floats order like signed ints, negative ones like
// reversed unsigned ints
__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// One CTA = FPB consecutive frames of one audio row.  Frame f covers samples [f*160 - 200, f*160 + 200) of the reflect-padded signal
// (torch.stft centre = True); the frame after the last full hop (index L / 160) is the one audio.py:70 drops.
// out_log[b][m][f] = log10(max(sum_k filt[m][k] * |X_f[k]|^2, 1e-10)); *gmax = max over everything (for the finish kernel).
__global__ void __launch_bounds__(MEL_THREADS) logmel_power_kernel(const float* __restrict__ audio, int L, int n_frames,
                                                                  const float* __restrict__ filt, int n_mels,
                                                                  float* __restrict__ out_log, float* __restrict__ gmax) {
  __shared__ float xw[FPB][N_FFT];
  __shared__ double tw_c[N_FFT], tw_s[N_FFT];
  __shared__ float pw[FPB][N_BINS + 3];
  __shared__ float wmax[MEL_THREADS / 32];
  const int tid = threadIdx.x, f0 = blockIdx.x * FPB, b = blockIdx.y;
  const float* a = audio + (size_t)b * L;
  for (int n = tid; n < N_FFT; n += MEL_THREADS) {
    double s, c;
    sincospi(2.0 * n / N_FFT, &s, &c);
    tw_c[n] = c; tw_s[n] = s;
  }
  for (int i = tid; i < FPB * N_FFT; i += MEL_THREADS) {
    const int f = i / N_FFT, n = i - f * N_FFT;
    float v = 0.f;
    if (f0 + f < n_frames) {
      int src = (f0 + f) * HOP + n - N_FFT / 2;
      if (src < 0) src = -src;
      if (src >= L) src = 2 * (L - 1) - src;
      const float hann = (float)(0.5 - 0.5 * cospi(2.0 * n / N_FFT));      // torch.hann_window(400), periodic
      v = a[src] * hann;
    }
    xw[f][n] = v;
  }
  __syncthreads();
  for (int i = tid; i < FPB * N_BINS; i += MEL_THREADS) {
    const int f = i / N_BINS, k = i - f * N_BINS;
    double re = 0.0, im = 0.0;
    int ph = 0;                                   // (k * n) mod 400
    for (int n = 0; n < N_FFT; ++n) {
      const double x = (double)xw[f][n];
      re = fma(x, tw_c[ph], re);
      im = fma(x, tw_s[ph], im);
      ph += k;
      if (ph >= N_FFT) ph -= N_FFT;
    }
    pw[f][k] = (float)(re * re + im * im);
  }
  __syncthreads();
  float local = -INFINITY;
  if (tid < n_mels) {
    float acc[FPB];
#pragma unroll
    for (int f = 0; f < FPB; ++f) acc[f] = 0.f;
    const float* fr = filt + (size_t)tid * N_BINS;
    for (int k = 0; k < N_BINS; ++k) {
      const float w = __ldg(fr + k);
      if (w != 0.f) {
#pragma unroll
        for (int f = 0; f < FPB; ++f) acc[f] = fmaf(w, pw[f][k], acc[f]);
      }
    }
    float* dst = out_log + ((size_t)b * n_mels + tid) * n_frames + f0;
#pragma unroll
    for (int f = 0; f < FPB; ++f)
      if (f0 + f < n_frames) {
        const float v = log10f(fmaxf(acc[f], 1e-10f));
        dst[f] = v;
        local = fmaxf(local, v);
      }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) local = fmaxf(local, __shfl_xor_sync(0xffffffffu, local, o));
  if ((tid & 31) == 0) wmax[tid >> 5] = local;
  __syncthreads();
  if (tid == 0) {
    float m = wmax[0];
    for (int w = 1; w < MEL_THREADS / 32; ++w) m = fmaxf(m, wmax[w]);
    if (m > -INFINITY) atomic_max_float(gmax, m);
  }
}

__global__ void logmel_finish_kernel(float* __restrict__ x, int64_t n, const float* __restrict__ gmax) {
  const float floor_v = *gmax - 8.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = (fmaxf(x[i], floor_v) + 4.0f) / 4.0f;
}

__global__ void fill_kernel(float* p, float v) { *p = v; }

// out[r][:] = table[(r / out_per_batch) * in_per_batch + idx[r % out_per_batch]][:]   (float4 rows)
__global__ void gather_rows_kernel(const float4* __restrict__ table, const int64_t* __restrict__ idx, float4* __restrict__ out,
                                   int64_t n_rows, int V, int64_t out_per_batch, int64_t in_per_batch) {
  const int64_t total = n_rows * V;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / V;
    const int q = (int)(i - r * V);
    const int64_t bb = r / out_per_batch, src = bb * in_per_batch + idx[r - bb * out_per_batch];
    out[i] = __ldg(table + src * V + q);
  }
}

// EuclideanCodebook.quantize (quantize/kmeans_codebook.py:15-23): argmax_v -(|x|^2 - 2 x.e_v + |e_v|^2) = argmax_v (x.e_v - |e_v|^2 / 2).
// half_norm[v] = -|e_v|^2 / 2 (the bias of the x E^T GEMM); one warp per codeword
__global__ void codebook_half_norm_kernel(const float* __restrict__ embed, int V, int C, float* __restrict__ out) {
  const int v = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (v >= V) return;
  float s = 0.f;
  for (int c = lane; c < C; c += 32) { const float e = embed[(size_t)v * C + c]; s = fmaf(e, e, s); }
#pragma unroll
  for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) out[v] = -0.5f * s;
}
// first index of the row maximum (torch.max(dim=-1).indices semantics on ties); one warp per row
__global__ void row_argmax_kernel(const float* __restrict__ score, int64_t M, int V, int64_t* __restrict__ idx) {
  const int64_t m = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (m >= M) return;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int v = lane; v < V; v += 32) {
    const float x = score[m * V + v];
    if (x > best) { best = x; bi = v; }
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  if (lane == 0) idx[m] = bi;
}

struct BlockW {
  const __nv_bfloat16 *qkv_h = nullptr, *out_h = nullptr, *fc1_h = nullptr, *fc2_h = nullptr;
  const float *qkv_b = nullptr, *out_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
};

}  // namespace

struct lds_units {
  lds_units_config cfg{};
  int device = 0;
  bool finalized = false;
  std::map<std::string, std::pair<std::vector<float>, std::vector<int64_t>>> raw;
  int parts = 1;
  float wscale = 1.f;
  float* warena = nullptr;
  __nv_bfloat16* wharena = nullptr;
  const __nv_bfloat16 *conv1_h = nullptr, *conv2_h = nullptr;
  const float *conv1_b = nullptr, *conv2_b = nullptr, *lnp_g = nullptr, *lnp_b = nullptr;
  std::vector<BlockW> blocks;
  float* arena = nullptr;
  size_t arena_cap = 0;
  __nv_bfloat16* barena = nullptr;
  size_t barena_cap = 0;
  int64_t launches = 0;
  double last_flops = 0;
};

extern "C" {

const char* lds_units_last_error(void) { return g_units_error.c_str(); }

int lds_units_create(const lds_units_config* cfg, int device, lds_units** out) {
  if (!cfg || !out) return ufail(LDS_ERR_INVALID, "null argument");
  if (cfg->precision != LDS_PREC_FP32 && cfg->precision != LDS_PREC_BF16) return ufail(LDS_ERR_INVALID, "precision must be LDS_PREC_FP32 or LDS_PREC_BF16");
  if (cfg->n_layer < 1 || cfg->n_head < 1 || cfg->n_state < 1 || cfg->n_mels < 1) return ufail(LDS_ERR_INVALID, "dimensions must be positive");
  if (cfg->n_state % cfg->n_head) return ufail(LDS_ERR_INVALID, "n_state %d is not a multiple of n_head %d", cfg->n_state, cfg->n_head);
  const int d = cfg->n_state / cfg->n_head;
  if (d != 32 && d != 48 && d != 64) return ufail(LDS_ERR_UNSUPPORTED, "head dim %d not in {32,48,64} (every Whisper size has 64)", d);
  if (cfg->n_state % 128 || cfg->n_state > 2048) return ufail(LDS_ERR_UNSUPPORTED, "n_state %d must be a multiple of 128, at most 2048", cfg->n_state);
  if (cfg->n_mels % 64) return ufail(LDS_ERR_UNSUPPORTED, "n_mels %d must be a multiple of 64 (whisper large-v3: 128)", cfg->n_mels);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev)
    return ufail(LDS_ERR_CUDA, "CUDA device %d not available (%s); the units encoder has no CPU fallback", device,
                 e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return ufail(LDS_ERR_UNSUPPORTED, "device %d is not sm_100; this library is built for B200 only", device);
  (void)lds::knobs();
  lds_units* u = new lds_units();
  u->cfg = *cfg;
  u->device = device;
  u->parts = cfg->precision == LDS_PREC_FP32 ? 2 : 1;
  *out = u;
  return LDS_OK;
}

void lds_units_destroy(lds_units* u) {
  if (!u) return;
  cudaSetDevice(u->device);
  cudaDeviceSynchronize();
  if (u->warena) cudaFree(u->warena);
  if (u->wharena) cudaFree(u->wharena);
  if (u->arena) cudaFree(u->arena);
  if (u->barena) cudaFree(u->barena);
  delete u;
}

int lds_units_load_weight(lds_units* u, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
  if (!u || !key || !data || !shape || ndim < 1 || ndim > 3) return ufail(LDS_ERR_INVALID, "bad argument to lds_units_load_weight");
  if (u->finalized) return ufail(LDS_ERR_INVALID, "weights already finalized");
  if (cudaSetDevice(u->device) != cudaSuccess) return ufail(LDS_ERR_CUDA, "cudaSetDevice failed");
  size_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) { n *= (size_t)shape[i]; shp.push_back(shape[i]); }
  std::vector<float> buf(n);
  if (dtype == LDS_DTYPE_F32) {
    const cudaError_t e = cudaMemcpy(buf.data(), data, n * sizeof(float), cudaMemcpyDefault);
    if (e != cudaSuccess) return ufail(LDS_ERR_CUDA, "copy of '%s' failed: %s", key, cudaGetErrorString(e));
  } else if (dtype == LDS_DTYPE_F16 || dtype == LDS_DTYPE_BF16) {      // the whisper checkpoints are stored in fp16
    std::vector<uint16_t> tmp(n);
    const cudaError_t e = cudaMemcpy(tmp.data(), data, n * 2, cudaMemcpyDefault);
    if (e != cudaSuccess) return ufail(LDS_ERR_CUDA, "copy of '%s' failed: %s", key, cudaGetErrorString(e));
    for (size_t i = 0; i < n; ++i) buf[i] = dtype == LDS_DTYPE_F16 ? PlanePacker::h2f(tmp[i]) : host_bf2f(tmp[i]);
  } else {
    return ufail(LDS_ERR_INVALID, "unknown dtype %d", dtype);
  }
  u->raw[key] = std::make_pair(std::move(buf), std::move(shp));
  return LDS_OK;
}

int lds_units_finalize(lds_units* u) {
  if (!u) return ufail(LDS_ERR_INVALID, "null handle");
  if (u->finalized) return LDS_OK;
  if (cudaSetDevice(u->device) != cudaSuccess) return ufail(LDS_ERR_CUDA, "cudaSetDevice failed");
  const int C = u->cfg.n_state, NM = u->cfg.n_mels, parts = u->parts;
  Packer pk;
  PlanePacker pkh;
  if (parts == 2) {      // one power-of-two scale for every packed weight (lds_api.cu does the same for the denoiser)
    float wmax = 0.f;
    for (const auto& kv : u->raw)
      if (kv.second.second.size() >= 2)
        for (float v : kv.second.first) wmax = std::max(wmax, std::fabs(v));
    float sc = 4096.f;
    while (sc > 1.f && wmax * sc >= 16384.f) sc *= 0.5f;
    u->wscale = pkh.scale = sc;
  }
  int rc = LDS_OK;
  auto need = [&](const std::string& key, std::vector<int64_t> shape) -> const float* {
    auto it = u->raw.find(key);
    if (it == u->raw.end()) { rc = ufail(LDS_ERR_MISSING, "weight '%s' was not loaded", key.c_str()); return nullptr; }
    if (it->second.second != shape) { rc = ufail(LDS_ERR_INVALID, "weight '%s' has an unexpected shape", key.c_str()); return nullptr; }
    return it->second.first.data();
  };
  std::vector<std::pair<const float**, size_t>> fix;
  std::vector<std::pair<const __nv_bfloat16**, size_t>> fixh;
  auto vec = [&](const float** slot, const std::string& key, int64_t n) {
    const float* p = need(key, {n});
    if (p) fix.emplace_back(slot, pk.add(p, (size_t)n));
    return p != nullptr;
  };
  auto lin = [&](const __nv_bfloat16** slot, const std::string& key, int64_t n, int64_t k) {
    const float* p = need(key, {n, k});
    if (p) fixh.emplace_back(slot, pkh.add(p, (size_t)n, (int)k, parts));
    return p != nullptr;
  };
  // conv1 [C, n_mels, 3] -> rows (n, tap) of n_mels channels (implicit-GEMM k3 convolution: three shifted TMA loads)
  {
    const float* w = need("conv1.weight", {C, NM, 3});
    if (!w) return rc;
    std::vector<float> t((size_t)C * 3 * NM);
    for (int n = 0; n < C; ++n)
      for (int c = 0; c < NM; ++c)
        for (int k = 0; k < 3; ++k) t[((size_t)n * 3 + k) * NM + c] = w[((size_t)n * NM + c) * 3 + k];
    fixh.emplace_back(&u->conv1_h, pkh.add(t.data(), (size_t)C * 3, NM, parts));
  }
  // conv2 [C, C, 3], stride 2 -> one row of 3C per output channel, K index = tap * C + c (the im2col layout of cast_gather mode 2)
  {
    const float* w = need("conv2.weight", {C, C, 3});
    if (!w) return rc;
    std::vector<float> t((size_t)C * 3 * C);
    for (int n = 0; n < C; ++n)
      for (int c = 0; c < C; ++c)
        for (int k = 0; k < 3; ++k) t[(size_t)n * 3 * C + (size_t)k * C + c] = w[((size_t)n * C + c) * 3 + k];
    fixh.emplace_back(&u->conv2_h, pkh.add(t.data(), (size_t)C, 3 * C, parts));
  }
  if (!vec(&u->conv1_b, "conv1.bias", C) || !vec(&u->conv2_b, "conv2.bias", C)) return rc;
  u->blocks.assign(u->cfg.n_layer, BlockW());
  for (int i = 0; i < u->cfg.n_layer; ++i) {
    BlockW& b = u->blocks[i];
    const std::string P = "blocks." + std::to_string(i) + ".";
    // fused QKV projection: rows [query | key | value]; the key projection has no bias (model.py:47)
    const float *wq = need(P + "attn.query.weight", {C, C}), *wk = need(P + "attn.key.weight", {C, C}), *wv = need(P + "attn.value.weight", {C, C});
    const float *bq = need(P + "attn.query.bias", {C}), *bv = need(P + "attn.value.bias", {C});
    if (!wq || !wk || !wv || !bq || !bv) return rc;
    const int H = u->cfg.n_head, d = C / H, dpad = d <= 32 ? 32 : 64;
    std::vector<float> t((size_t)3 * H * dpad * C, 0.f), tb((size_t)3 * H * dpad, 0.f);
    for (int r = 0; r < 3; ++r) {
      const float* src = r == 0 ? wq : (r == 1 ? wk : wv);
      const float* bsrc = r == 0 ? bq : (r == 1 ? nullptr : bv);
      for (int hh = 0; hh < H; ++hh)
        for (int j = 0; j < d; ++j) {
          const size_t row = ((size_t)r * H + hh) * dpad + j;
          memcpy(&t[row * C], src + ((size_t)hh * d + j) * C, sizeof(float) * C);
          if (bsrc) tb[row] = bsrc[hh * d + j];
        }
    }
    fixh.emplace_back(&b.qkv_h, pkh.add(t.data(), (size_t)3 * H * dpad, C, parts));
    fix.emplace_back(&b.qkv_b, pk.add(tb.data(), tb.size()));
    if (!lin(&b.out_h, P + "attn.out.weight", C, C) || !vec(&b.out_b, P + "attn.out.bias", C) ||
        !lin(&b.fc1_h, P + "mlp.0.weight", 4 * C, C) || !vec(&b.fc1_b, P + "mlp.0.bias", 4 * C) ||
        !lin(&b.fc2_h, P + "mlp.2.weight", C, 4 * C) || !vec(&b.fc2_b, P + "mlp.2.bias", C) ||
        !vec(&b.ln1_g, P + "attn_ln.weight", C) || !vec(&b.ln1_b, P + "attn_ln.bias", C) ||
        !vec(&b.ln2_g, P + "mlp_ln.weight", C) || !vec(&b.ln2_b, P + "mlp_ln.bias", C))
      return rc;
  }
  if (!vec(&u->lnp_g, "ln_post.weight", C) || !vec(&u->lnp_b, "ln_post.bias", C)) return rc;
  if (cudaMalloc(&u->warena, pk.host.size() * sizeof(float)) != cudaSuccess ||
      cudaMalloc(&u->wharena, pkh.host.size() * sizeof(uint16_t)) != cudaSuccess)
    return ufail(LDS_ERR_CUDA, "cudaMalloc of the encoder weights (%zu MB) failed", (pk.host.size() * 4 + pkh.host.size() * 2) >> 20);
  if (cudaMemcpy(u->warena, pk.host.data(), pk.host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(u->wharena, pkh.host.data(), pkh.host.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess)
    return ufail(LDS_ERR_CUDA, "upload of the encoder weights failed");
  for (auto& f : fix) *f.first = u->warena + f.second;
  for (auto& f : fixh) *f.first = u->wharena + f.second;
  u->raw.clear();
  u->finalized = true;
  return LDS_OK;
}

int lds_units_out_frames(int L) { return L < 1 ? 0 : (L - 1) / 2 + 1; }

int lds_units_encode(lds_units* u, const float* mel_BML, int B, int L, const float* pos_TC, float* out_BTC, void* stream) {
  if (!u || !u->finalized) return ufail(LDS_ERR_INVALID, "lds_units_encode: weights not finalized");
  if (!mel_BML || !pos_TC || !out_BTC || B < 1 || L < 1) return ufail(LDS_ERR_INVALID, "lds_units_encode: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaSetDevice(u->device) != cudaSuccess) return ufail(LDS_ERR_CUDA, "cudaSetDevice failed");
  const int C = u->cfg.n_state, NM = u->cfg.n_mels, H = u->cfg.n_head, d = C / H, dpad = d <= 32 ? 32 : 64, parts = u->parts;
  const int T = lds_units_out_frames(L), t_pad = (T + 7) / 8 * 8;
  const int64_t M = (int64_t)B * T, ML = (int64_t)B * L;
  // ---- workspace (grow-only; growing synchronises because work in flight may still use the old arena) ----
  auto r64 = [](size_t n) { return (n + 63) / 64 * 64; };
  const size_t f_need = r64((size_t)ML * NM) + r64((size_t)ML * C) + 2 * r64((size_t)M * C) + 64;
  const size_t b_need = r64((size_t)ML * parts * NM) + r64((size_t)M * parts * 3 * C) + 2 * r64((size_t)M * parts * C) +
                        r64((size_t)M * parts * 4 * C) + 2 * r64((size_t)M * parts * H * dpad) + r64((size_t)B * parts * H * dpad * t_pad) + 64;
  if (f_need > u->arena_cap || b_need > u->barena_cap) {
    if (cudaDeviceSynchronize() != cudaSuccess) return ufail(LDS_ERR_CUDA, "synchronize before workspace growth failed");
    if (f_need > u->arena_cap) {
      if (u->arena) cudaFree(u->arena);
      u->arena = nullptr; u->arena_cap = 0;
      if (cudaMalloc(&u->arena, f_need * sizeof(float)) != cudaSuccess) return ufail(LDS_ERR_CUDA, "cudaMalloc of %zu MB encoder workspace failed", f_need * 4 >> 20);
      u->arena_cap = f_need;
    }
    if (b_need > u->barena_cap) {
      if (u->barena) cudaFree(u->barena);
      u->barena = nullptr; u->barena_cap = 0;
      if (cudaMalloc(&u->barena, b_need * 2) != cudaSuccess) return ufail(LDS_ERR_CUDA, "cudaMalloc of %zu MB encoder operand workspace failed", b_need * 2 >> 20);
      u->barena_cap = b_need;
    }
  }
  size_t fo = 0, bo = 0;
  auto takef = [&](size_t n) { float* p = u->arena + fo; fo += r64(n); return p; };
  auto takeb = [&](size_t n) { __nv_bfloat16* p = u->barena + bo; bo += r64(n); return p; };
  float* mel_t = takef((size_t)ML * NM);
  float* c1 = takef((size_t)ML * C);
  float* x = takef((size_t)M * C);
  float* pos_rep = takef((size_t)M * C);
  __nv_bfloat16* mel_p = takeb((size_t)ML * parts * NM);
  __nv_bfloat16* col_p = takeb((size_t)M * parts * 3 * C);
  __nv_bfloat16* xn_p = takeb((size_t)M * parts * C);
  __nv_bfloat16* att_p = takeb((size_t)M * parts * C);
  __nv_bfloat16* h_p = takeb((size_t)M * parts * 4 * C);
  __nv_bfloat16* q_p = takeb((size_t)M * parts * H * dpad);
  __nv_bfloat16* k_p = takeb((size_t)M * parts * H * dpad);
  __nv_bfloat16* vt_p = takeb((size_t)B * parts * H * dpad * t_pad);

  double flops = 0;
  auto ck = [&](cudaError_t e, const char* what) -> int {
    ++u->launches;
    return e == cudaSuccess ? LDS_OK : ufail(LDS_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  };
#define UTRY(expr) do { int rc__ = (expr); if (rc__ != LDS_OK) return rc__; } while (0)
  auto base = [&](const __nv_bfloat16* A, int batches, int rows, int cin, int taps, const __nv_bfloat16* W, const float* bias, int N) {
    TcGemmArgs g;
    g.A = A; g.batches = batches; g.rows = rows; g.cin = cin; g.taps = taps; g.W = W; g.N = N; g.bias = bias;
    if (parts == 2) { tc_set_split_pairs(g); g.out_scale = 1.f / (PLANE_SCALE * u->wscale); }
    flops += 2.0 * batches * rows * (double)N * taps * cin;
    return g;
  };
  auto out_f32 = [&](TcGemmArgs& g, float* Cp, int ld) { g.C = Cp; g.c_ld = ld; g.out_kind = 0; };
  auto out_planes = [&](TcGemmArgs& g, __nv_bfloat16* Cp, int n_out) { g.C = Cp; g.c_ld = parts * n_out; g.out_kind = parts == 2 ? 2 : 1; };

  // x + sinusoids(T, C) (model.py:123): the table repeats per utterance -> the fp32 residual operand of conv2's epilogue
  for (int b = 0; b < B; ++b)
    if (cudaMemcpyAsync(pos_rep + (size_t)b * T * C, pos_TC, sizeof(float) * (size_t)T * C, cudaMemcpyDeviceToDevice, s) != cudaSuccess)
      return ufail(LDS_ERR_CUDA, "copy of the positional table failed");
  // conv1 + GELU (model.py:120): mel [B, n_mels, L] -> channels-last planes -> implicit-GEMM k3 convolution
  UTRY(ck(launch_transpose_bct_to_btc(mel_BML, mel_t, B, NM, L, 1.f, s), "transpose"));
  UTRY(ck(launch_split_cast(mel_t, mel_p, ML, NM, parts, s), "split_cast"));
  {
    TcGemmArgs g = base(mel_p, B, L, NM, 3, u->conv1_h, u->conv1_b, C);
    g.epilogue = EPI_GELU;
    out_f32(g, c1, C);
    UTRY(ck(launch_gemm_tc(g, s), "conv1"));
  }
  // conv2 k3 / stride 2 + GELU, + positional table (model.py:121-123): im2col gather-cast, then one GEMM with K = 3C
  UTRY(ck(launch_cast_gather(c1, col_p, B, L, T, C, parts, 2, 0.f, s), "cast_im2col_s2"));
  {
    TcGemmArgs g = base(col_p, 1, (int)M, 3 * C, 1, u->conv2_h, u->conv2_b, C);
    g.epilogue = EPI_GELU;
    out_f32(g, x, C);
    g.R = pos_rep; g.r_ld = C;
    UTRY(ck(launch_gemm_tc(g, s), "conv2"));
  }
  for (const BlockW& w : u->blocks) {   // ResidualAttentionBlock.forward (model.py:104-110)
    UTRY(ck(launch_layernorm(x, w.ln1_g, w.ln1_b, 1e-5f, (int)M, C, nullptr, xn_p, parts, s), "attn_ln"));
    {
      TcGemmArgs g = base(xn_p, 1, (int)M, C, 1, w.qkv_h, w.qkv_b, 3 * H * dpad);
      g.out_kind = 3; g.q_out = q_p; g.k_out = k_p; g.vt_out = vt_p;
      g.att_T = T; g.att_H = H; g.att_dpad = dpad; g.att_Tpad = t_pad; g.att_parts = parts;
      UTRY(ck(launch_gemm_tc(g, s), "qkv"));
    }
    {   // softmax((q d^-1/4)(k d^-1/4)^T) v (model.py:70-87) = softmax(q k^T / sqrt(d)) v
      AttnTcArgs at;
      at.q = q_p; at.k = k_p; at.vt = vt_p; at.out = att_p;
      at.B = B; at.T = T; at.T_pad = t_pad; at.H = H; at.d = d; at.dpad = dpad; at.parts = parts; at.out_parts = parts;
      UTRY(ck(launch_attention_tc(at, s), "attention"));
      flops += 4.0 * B * (double)T * T * C;
    }
    {
      TcGemmArgs g = base(att_p, 1, (int)M, C, 1, w.out_h, w.out_b, C);
      out_f32(g, x, C);
      g.R = x; g.r_ld = C;
      UTRY(ck(launch_gemm_tc(g, s), "attn.out"));
    }
    UTRY(ck(launch_layernorm(x, w.ln2_g, w.ln2_b, 1e-5f, (int)M, C, nullptr, xn_p, parts, s), "mlp_ln"));
    {
      TcGemmArgs g = base(xn_p, 1, (int)M, C, 1, w.fc1_h, w.fc1_b, 4 * C);
      g.epilogue = EPI_GELU;
      out_planes(g, h_p, 4 * C);
      UTRY(ck(launch_gemm_tc(g, s), "mlp.0"));
    }
    {
      TcGemmArgs g = base(h_p, 1, (int)M, 4 * C, 1, w.fc2_h, w.fc2_b, C);
      out_f32(g, x, C);
      g.R = x; g.r_ld = C;
      UTRY(ck(launch_gemm_tc(g, s), "mlp.2"));
    }
  }
  UTRY(ck(launch_layernorm(x, u->lnp_g, u->lnp_b, 1e-5f, (int)M, C, out_BTC, nullptr, 1, s), "ln_post"));
#undef UTRY
  u->last_flops = flops;
  return LDS_OK;
}

int64_t lds_units_launches(const lds_units* u) { return u ? u->launches : 0; }
double lds_units_last_flops(const lds_units* u) { return u ? u->last_flops : 0; }
int64_t lds_units_workspace_bytes(const lds_units* u) { return u ? (int64_t)(u->arena_cap * 4 + u->barena_cap * 2) : 0; }

int lds_units_log_mel(const float* audio_BL, int B, int L, const float* filters, int n_mels, float* mel_out, float* scratch, void* stream) {
  if (!audio_BL || !filters || !mel_out || !scratch || B < 1) return ufail(LDS_ERR_INVALID, "lds_units_log_mel: bad argument");
  if (L <= N_FFT / 2) return ufail(LDS_ERR_INVALID, "lds_units_log_mel: %d samples are too few for reflect padding by %d", L, N_FFT / 2);
  if (n_mels < 1 || n_mels > MEL_THREADS) return ufail(LDS_ERR_UNSUPPORTED, "lds_units_log_mel: n_mels must be in [1,%d]", MEL_THREADS);
  const int n_frames = L / HOP;
  if (n_frames < 1) return ufail(LDS_ERR_INVALID, "lds_units_log_mel: fewer than %d samples", HOP);
  cudaStream_t s = (cudaStream_t)stream;
  fill_kernel<<<1, 1, 0, s>>>(scratch, -INFINITY);
  logmel_power_kernel<<<dim3((n_frames + FPB - 1) / FPB, B), MEL_THREADS, 0, s>>>(audio_BL, L, n_frames, filters, n_mels, mel_out, scratch);
  const int64_t n = (int64_t)B * n_mels * n_frames;
  logmel_finish_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, 1184), 256, 0, s>>>(mel_out, n, scratch);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? LDS_OK : ufail(LDS_ERR_CUDA, "lds_units_log_mel: %s", cudaGetErrorString(e));
}

int lds_units_gather_rows(const float* table, const int64_t* idx, int64_t n_batches, int64_t out_per_batch, int64_t in_per_batch, int C,
                          float* out, void* stream) {
  if (!table || !idx || !out || n_batches < 0 || out_per_batch < 0 || C < 4 || C % 4) return ufail(LDS_ERR_INVALID, "lds_units_gather_rows: bad argument");
  const int64_t rows = n_batches * out_per_batch;
  if (rows == 0) return LDS_OK;
  const int64_t total = rows * (C / 4);
  gather_rows_kernel<<<(unsigned)std::min<int64_t>((total + 255) / 256, 4736), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const float4*>(table), idx, reinterpret_cast<float4*>(out), rows, C / 4, out_per_batch, in_per_batch);
  const cudaError_t e = cudaGetLastError();
  return e == cudaSuccess ? LDS_OK : ufail(LDS_ERR_CUDA, "lds_units_gather_rows: %s", cudaGetErrorString(e));
}


int lds_units_quantize(const float* x_MC, const float* embed_VC, int64_t M, int V, int C, float* scratch, int64_t* idx_out, void* stream) {
  if (!x_MC || !embed_VC || !scratch || !idx_out || M < 0 || V < 1 || C < 16 || C % 16 || V % 4)
    return ufail(LDS_ERR_INVALID, "lds_units_quantize: bad argument (C must be a multiple of 16, V of 4)");
  if (M == 0) return LDS_OK;
  if (M > 0x7fffffff) return ufail(LDS_ERR_UNSUPPORTED, "lds_units_quantize: more than 2^31 rows");
  cudaStream_t s = (cudaStream_t)stream;
  float* half_norm = scratch;                         // [V], then the scores [M, V]
  float* score = scratch + ((size_t)V + 63) / 64 * 64;
  codebook_half_norm_kernel<<<(V * 32 + 255) / 256, 256, 0, s>>>(embed_VC, V, C, half_norm);
  GemmArgs g;                                          // score = x E^T - |e|^2 / 2 on the fp32 FFMA kernel (IEEE products: index decisions)
  g.A = x_MC; g.a_ld = C; g.W = embed_VC; g.C = score; g.c_ld = V; g.bias = half_norm; g.M = (int)M; g.N = V; g.K = C; g.taps = 1; g.cin = C;
  g.t_out = g.t_in = g.t_conv = (int)M;
  cudaError_t e = launch_gemm_f32(g, s);
  if (e != cudaSuccess) return ufail(LDS_ERR_CUDA, "lds_units_quantize: gemm: %s", cudaGetErrorString(e));
  row_argmax_kernel<<<(unsigned)((M * 32 + 255) / 256), 256, 0, s>>>(score, M, V, idx_out);
  e = cudaGetLastError();
  return e == cudaSuccess ? LDS_OK : ufail(LDS_ERR_CUDA, "lds_units_quantize: %s", cudaGetErrorString(e));
}

}  // extern "C"
