// vocoder.cu — HiFi-VAEGAN `Generator` decode (latent / mel frames -> waveform) on B200, the step that follows the Unit2Mel
// sampling path.  C ABI: lds_vocoder_* / lds_vocode (include/lds_b200.h).
//
// Reference (file:line relative to the reference tree):
//   encoder/hifi_vaegan/modules/models.py:224-266   Generator: conv_pre k7 -> [leaky_relu -> ConvTranspose1d -> mean of the
//                                                    resblocks] x n_ups -> leaky_relu(0.01) -> conv_post k7 -> tanh
//   encoder/hifi_vaegan/modules/models.py:161-221   ResBlock1 / ResBlock2 (dilated k in {3,7,11} convolutions with residuals)
//   encoder/hifi_vaegan/hifi_vaegan.py:52-65        Hifi_VAEGAN.forward: [B,T,C] -> [B,C,T], remove_weight_norm, Generator
//
// Three kinds of layers:
//   * ResBlock levels whose channel count is a multiple of 64 (256, 128 and 64 channels in the HiFi-GAN V1 layout) run on the
//     tensor cores: channels-LAST fp32 [B*L, C] state, every dilated k in {3..11} convolution an implicit GEMM of gemm_tc.cu (tap t =
//     the TMA row coordinate shifted by (t - (k-1)/2) * dilation, zero padding = TMA out-of-bounds fill, split-f16 operand planes:
//     fp32-accurate at three tcgen05 products per logical product), leaky_relu fused into the operand cast / the first convolution's
//     epilogue, bias + residual in the second one's;
//   * the transposed convolutions (kernel 2u, stride u) are one 3-tap implicit GEMM each with N = u * C_out (lds_vocoder_finalize);
//   * the 32-channel level is time-folded onto the same kernel (fold_of below): two frames per GEMM row, block-sparse folded taps;
//   * conv_pre (128 -> 512, k7) is a plain 7-tap implicit GEMM on the mel frames, which already are its channels-last operand;
//   * conv_post (and any layer the GEMM forms do not cover) stays channels-FIRST fp32 [B, C, L] on the CUDA cores: IEEE
//     FFMA, register-blocked direct convolution (8 output channels x 8 time steps per thread, input slab with its dilated halo and the [ci][tap][co] weight slab staged in shared memory,
//     leaky_relu applied once while staging).
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
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

thread_local std::string g_voc_error;
int vfail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_voc_error = buf;
  return code;
}

constexpr int L_T = 128;      // time steps per CTA
constexpr int CI_T = 8;       // input channels per shared-memory stage
constexpr int MAX_HALO = 10 * 5;   // (k - 1) * dilation for k = 11, d = 5

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : v * slope; }

// y[b, co, l] = epi( bias[co] + sum_ci sum_k W[ci][k][co] * lrelu_in(x[b, ci, l + k*dil - pad]) )
//   in_slope: leaky_relu slope applied to the input (1 = identity)
//   res: optional residual added to the conv result
//   acc_mode 0: y = v ; 1: y = y_old + v ; 2: y = (y_old + v) / acc_div          (mean over the resblocks, models.py:243-251)
//   out_tanh: tanh on the way out (conv_post)
// CO_T output channels per CTA: 128 threads = 16 time lanes x 8 channel lanes; a thread owns CO_T/8 channels x 8 time steps
// (time steps interleaved by 16 so that shared-memory reads of x are conflict-free and global stores are 64-byte runs).
template <int KS, int CO_T>
__global__ void __launch_bounds__(128)
voc_conv1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, const float* __restrict__ res,
                  float* __restrict__ y, int C_in, int C_out, int L, int dil, int pad, float in_slope, int acc_mode, float acc_div, int out_tanh) {
  constexpr int CPT = CO_T / 8;                                   // channels per thread
  __shared__ float xs[CI_T][L_T + MAX_HALO];
  __shared__ __align__(16) float ws[CI_T][KS][CO_T];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int l0 = blockIdx.x * L_T, co0 = blockIdx.y * CO_T, b = blockIdx.z;
  const int span = L_T + (KS - 1) * dil;
  const float* xb = x + (size_t)b * C_in * L;
  float acc[CPT][8];
#pragma unroll
  for (int c = 0; c < CPT; ++c)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[c][j] = 0.f;

  for (int ci0 = 0; ci0 < C_in; ci0 += CI_T) {
    __syncthreads();
    for (int i = threadIdx.x; i < CI_T * span; i += 128) {        // input slab with halo, activation applied once
      const int ci = i / span, s = i - ci * span;
      const int l = l0 + s - pad;
      float v = 0.f;
      if (ci0 + ci < C_in && l >= 0 && l < L) v = lrelu(__ldg(xb + (size_t)(ci0 + ci) * L + l), in_slope);
      xs[ci][s] = v;
    }
    for (int i = threadIdx.x; i < CI_T * KS * CO_T; i += 128) {   // weights are packed [C_in][KS][C_out]: contiguous slab per ci
      const int ci = i / (KS * CO_T), r = i - ci * (KS * CO_T);
      const int k = r / CO_T, co = r - k * CO_T;
      ws[ci][k][co] = (ci0 + ci < C_in && co0 + co < C_out) ? __ldg(w + ((size_t)(ci0 + ci) * KS + k) * C_out + co0 + co) : 0.f;
    }
    __syncthreads();
#pragma unroll 2
    for (int ci = 0; ci < CI_T; ++ci) {
#pragma unroll
      for (int k = 0; k < KS; ++k) {
        float wv[CPT], xv[8];
#pragma unroll
        for (int c = 0; c < CPT; ++c) wv[c] = ws[ci][k][ty * CPT + c];
        const float* xr = &xs[ci][tx + k * dil];
#pragma unroll
        for (int j = 0; j < 8; ++j) xv[j] = xr[16 * j];
#pragma unroll
        for (int c = 0; c < CPT; ++c)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[c][j] = fmaf(wv[c], xv[j], acc[c][j]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < CPT; ++c) {
    const int co = co0 + ty * CPT + c;
    if (co >= C_out) continue;
    const float bv = bias ? __ldg(bias + co) : 0.f;
    const size_t rowoff = ((size_t)b * C_out + co) * L;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int l = l0 + tx + 16 * j;
      if (l >= L) continue;
      float v = acc[c][j] + bv;
      if (res) v = v + res[rowoff + l];                    // x = xt + x
      if (acc_mode == 1) v = y[rowoff + l] + v;            // xs += resblock(x)
      else if (acc_mode == 2) v = __fdiv_rn(y[rowoff + l] + v, acc_div);   // x = xs / num_kernels
      if (out_tanh) v = tanhf(v);
      y[rowoff + l] = v;
    }
  }
}

// ConvTranspose1d(stride u, kernel KS, padding pad) on lrelu(x):  y[b, co, lo] = bias[co] + sum_ci sum_k W[ci][co][k] * f(x[b, ci, li]),
// lo = li*u - pad + k  =>  for an output lo the taps are k = (lo + pad) mod u + m*u with li = (lo + pad - k) / u.
// Weights packed [C_in][KS][C_out].  A thread owns 4 output channels x 4 consecutive outputs; ~3 % of the generator's FLOPs.
__global__ void __launch_bounds__(128)
voc_convtr1d_kernel(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ y,
                    int C_in, int C_out, int L_in, int L_out, int KS, int u, int pad, float in_slope) {
  const int lo0 = (blockIdx.x * 32 + (threadIdx.x & 31)) * 4;
  const int co0 = (blockIdx.y * 4 + (threadIdx.x >> 5)) * 4;
  const int b = blockIdx.z;
  if (lo0 >= L_out || co0 >= C_out) return;
  const float* xb = x + (size_t)b * C_in * L_in;
  float acc[4][4];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[c][j] = 0.f;
  const int taps = KS / u + (KS % u ? 1 : 0);
  for (int ci = 0; ci < C_in; ++ci) {
    const float* xr = xb + (size_t)ci * L_in;
    const float* wr = w + (size_t)ci * KS * C_out;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int lo = lo0 + j;
      if (lo >= L_out) continue;
      const int r = (lo + pad) % u;
      for (int m = 0; m < taps; ++m) {
        const int k = r + m * u;
        if (k >= KS) break;
        const int li = (lo + pad - k) / u;
        if (li < 0 || li >= L_in) continue;
        const float xv = lrelu(__ldg(xr + li), in_slope);
        const float4 wv = __ldg(reinterpret_cast<const float4*>(wr + (size_t)k * C_out + co0));
        acc[0][j] = fmaf(wv.x, xv, acc[0][j]); acc[1][j] = fmaf(wv.y, xv, acc[1][j]);
        acc[2][j] = fmaf(wv.z, xv, acc[2][j]); acc[3][j] = fmaf(wv.w, xv, acc[3][j]);
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float bv = bias ? __ldg(bias + co0 + c) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (lo0 + j < L_out) y[((size_t)b * C_out + co0 + c) * L_out + lo0 + j] = acc[c][j] + bv;
  }
}

// [B, T, C] -> [B, C, T]
__global__ void voc_transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int T, int C) {
  __shared__ float tile[32][33];
  const int b = blockIdx.z, c0 = blockIdx.x * 32, t0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int t = t0 + j, c = c0 + threadIdx.x;
    if (t < T && c < C) tile[j][threadIdx.x] = in[((size_t)b * T + t) * C + c];
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    const int c = c0 + j, t = t0 + threadIdx.x;
    if (t < T && c < C) out[((size_t)b * C + c) * T + t] = tile[threadIdx.x][j];
  }
}

struct ConvP {
  const float* w = nullptr; const float* b = nullptr; int cin = 0, cout = 0, k = 0;
  const __nv_bfloat16* wh = nullptr;      // tensor-core form (gemm_tc operand planes)
  const float* b_rep = nullptr;           // bias repeated over the N columns of a folded / transposed-convolution GEMM
  int fold = 1, ntaps = 0, tap_off[24] = {0};   // time-folded form of a narrow level: `fold` frames per GEMM row, coarse tap offsets
};

// resblock levels that run as implicit GEMMs on the tensor cores (gemm_tc: K blocks of 64 channels, N tiles of 64 ... 256)
inline bool tc_level(int ch) { return (ch >= 64 && ch % 64 == 0) || ch == 32 || ch == 16; }
// Narrow levels are TIME-FOLDED onto the 64-wide K blocks: `fold` = 64 / ch consecutive frames form one GEMM row of 64 "channels"
// (the channels-last tensor [B, L, ch] IS [B, L / fold, 64] in memory), and the dilated convolution becomes a convolution over
// folded rows with block-sparse 64 x 64 taps at the (non-uniform) folded offsets floor((s + delta) / fold).
inline int fold_of(int ch) { return ch < 64 ? 64 / ch : 1; }
inline int floordiv(int a, int b) { return a >= 0 ? a / b : -((-a + b - 1) / b); }

// transposed convolutions that run as a 3-tap implicit GEMM (kernel 2u, even stride u; K blocks of 64 channels, N = u * cout tiles of 64)
inline bool up_tc_ok(int cin, int cout, int k, int u) { return u >= 2 && u % 2 == 0 && k == 2 * u && cin % 64 == 0 && (u * cout) % 64 == 0; }

// xs (+)= r over n floats: mode 0 xs = r ; 1 xs = xs + r ; 2 xs = (xs + r) / div     (mean over the resblocks, models.py:243-251)
__global__ void voc_acc_kernel(float4* __restrict__ xs, const float4* __restrict__ r, int64_t n4, int mode, float div) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    float4 a = r[i];
    if (mode) {
      const float4 o = xs[i];
      a.x = o.x + a.x; a.y = o.y + a.y; a.z = o.z + a.z; a.w = o.w + a.w;
      if (mode == 2) { a.x = __fdiv_rn(a.x, div); a.y = __fdiv_rn(a.y, div); a.z = __fdiv_rn(a.z, div); a.w = __fdiv_rn(a.w, div); }
    }
    xs[i] = a;
  }
}

template <int KS>
cudaError_t launch_conv_ks(const float* x, const ConvP& c, const float* res, float* y, int B, int L, int dil, int pad, float in_slope,
                           int acc_mode, float acc_div, int out_tanh, cudaStream_t s) {
  if (c.cout >= 64) {
    dim3 grid((L + L_T - 1) / L_T, (c.cout + 63) / 64, B);
    voc_conv1d_kernel<KS, 64><<<grid, 128, 0, s>>>(x, c.w, c.b, res, y, c.cin, c.cout, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh);
  } else if (c.cout >= 16) {
    dim3 grid((L + L_T - 1) / L_T, (c.cout + 31) / 32, B);
    voc_conv1d_kernel<KS, 32><<<grid, 128, 0, s>>>(x, c.w, c.b, res, y, c.cin, c.cout, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh);
  } else {
    dim3 grid((L + L_T - 1) / L_T, (c.cout + 7) / 8, B);
    voc_conv1d_kernel<KS, 8><<<grid, 128, 0, s>>>(x, c.w, c.b, res, y, c.cin, c.cout, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh);
  }
  return cudaGetLastError();
}

cudaError_t launch_conv(const float* x, const ConvP& c, const float* res, float* y, int B, int L, int dil, float in_slope, int acc_mode,
                        float acc_div, int out_tanh, cudaStream_t s) {
  const int pad = (c.k * dil - dil) / 2;            // get_padding (commons.py:13-14)
  if ((c.k - 1) * dil > MAX_HALO) return cudaErrorInvalidValue;
  switch (c.k) {
    case 3: return launch_conv_ks<3>(x, c, res, y, B, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh, s);
    case 5: return launch_conv_ks<5>(x, c, res, y, B, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh, s);
    case 7: return launch_conv_ks<7>(x, c, res, y, B, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh, s);
    case 11: return launch_conv_ks<11>(x, c, res, y, B, L, dil, pad, in_slope, acc_mode, acc_div, out_tanh, s);
    default: return cudaErrorNotSupported;
  }
}

}  // namespace

struct lds_vocoder {
  lds_vocoder_config cfg{};
  int device = 0;
  bool finalized = false;
  std::map<std::string, std::pair<std::vector<float>, std::vector<int64_t>>> raw;
  float* warena = nullptr;
  size_t warena_floats = 0;
  __nv_bfloat16* wharena = nullptr;          // split-f16 operand planes of the tensor-core levels' resblock filters
  float wscale = 1.f;
  ConvP conv_pre, conv_post;
  std::vector<ConvP> ups;
  std::vector<std::vector<ConvP>> rb1, rb2;      // per resblock: convs1 / convs2 (ResBlock2: rb1 = convs, rb2 empty)
  float* arena = nullptr;
  size_t arena_cap = 0;
  int64_t launches = 0;
  double flops_last = 0;
};

extern "C" {

const char* lds_vocoder_last_error(void) { return g_voc_error.c_str(); }

int lds_vocoder_create(const lds_vocoder_config* cfg, int device, lds_vocoder** out) {
  if (!cfg || !out) return vfail(LDS_ERR_INVALID, "null argument");
  if (cfg->n_ups < 1 || cfg->n_ups > LDS_VOC_MAX || cfg->n_kernels < 1 || cfg->n_kernels > LDS_VOC_MAX)
    return vfail(LDS_ERR_INVALID, "n_ups and n_kernels must be in [1,%d]", LDS_VOC_MAX);
  if (cfg->resblock_kind != 1 && cfg->resblock_kind != 2) return vfail(LDS_ERR_INVALID, "resblock_kind must be 1 or 2");
  if (cfg->inter_channels < 1 || cfg->upsample_initial_channel < 1) return vfail(LDS_ERR_INVALID, "channel counts must be positive");
  int ch = cfg->upsample_initial_channel;
  for (int i = 0; i < cfg->n_ups; ++i) {
    const int u = cfg->upsample_rates[i], k = cfg->upsample_kernel_sizes[i];
    if (u < 1 || k < u || ch % 2) return vfail(LDS_ERR_INVALID, "upsample stage %d: rate %d, kernel %d, channels %d", i, u, k, ch);
    ch /= 2;
    if (ch % 4) return vfail(LDS_ERR_UNSUPPORTED, "channel count %d after stage %d must be a multiple of 4", ch, i);
  }
  for (int j = 0; j < cfg->n_kernels; ++j) {
    const int k = cfg->resblock_kernel_sizes[j];
    if (k != 3 && k != 5 && k != 7 && k != 11) return vfail(LDS_ERR_UNSUPPORTED, "resblock kernel size %d not in {3,5,7,11}", k);
    for (int n = 0; n < 3; ++n) {
      const int d = cfg->resblock_dilations[j][n];
      if (d < 1 || (k - 1) * d > MAX_HALO) return vfail(LDS_ERR_UNSUPPORTED, "resblock dilation %d with kernel %d exceeds the staged halo", d, k);
    }
  }
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || device < 0 || device >= ndev)
    return vfail(LDS_ERR_CUDA, "CUDA device %d not available (%s); the vocoder has no CPU fallback", device,
                 e == cudaSuccess ? "index out of range" : cudaGetErrorString(e));
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major != 10)
    return vfail(LDS_ERR_UNSUPPORTED, "device %d is not sm_100; this library is built for B200 only", device);
  lds_vocoder* v = new lds_vocoder();
  v->cfg = *cfg;
  v->device = device;
  *out = v;
  return LDS_OK;
}

void lds_vocoder_destroy(lds_vocoder* v) {
  if (!v) return;
  cudaSetDevice(v->device);
  cudaDeviceSynchronize();
  if (v->warena) cudaFree(v->warena);
  if (v->wharena) cudaFree(v->wharena);
  if (v->arena) cudaFree(v->arena);
  delete v;
}

int lds_vocoder_load_weight(lds_vocoder* v, const char* key, const void* data, const int64_t* shape, int ndim, int dtype) {
  if (!v || !key || !data || !shape || ndim < 1 || ndim > 3) return vfail(LDS_ERR_INVALID, "bad argument to lds_vocoder_load_weight");
  if (v->finalized) return vfail(LDS_ERR_INVALID, "weights already finalized");
  if (dtype != LDS_DTYPE_F32) return vfail(LDS_ERR_UNSUPPORTED, "vocoder weights must be fp32");
  if (cudaSetDevice(v->device) != cudaSuccess) return vfail(LDS_ERR_CUDA, "cudaSetDevice failed");
  size_t n = 1;
  std::vector<int64_t> shp;
  for (int i = 0; i < ndim; ++i) { n *= (size_t)shape[i]; shp.push_back(shape[i]); }
  std::vector<float> buf(n);
  cudaError_t e = cudaMemcpy(buf.data(), data, n * sizeof(float), cudaMemcpyDefault);
  if (e != cudaSuccess) return vfail(LDS_ERR_CUDA, "copy of '%s' failed: %s", key, cudaGetErrorString(e));
  v->raw[key] = std::make_pair(std::move(buf), std::move(shp));
  return LDS_OK;
}

int lds_vocoder_finalize(lds_vocoder* v) {
  if (!v) return vfail(LDS_ERR_INVALID, "null handle");
  if (v->finalized) return LDS_OK;
  if (cudaSetDevice(v->device) != cudaSuccess) return vfail(LDS_ERR_CUDA, "cudaSetDevice failed");
  const lds_vocoder_config& c = v->cfg;
  std::vector<float> host;
  std::vector<std::pair<const float**, size_t>> fix;
  auto put = [&](const float** slot, const float* src, size_t n) {
    const size_t off = (host.size() + 63) / 64 * 64;
    host.resize(off + n);
    memcpy(host.data() + off, src, n * sizeof(float));
    fix.emplace_back(slot, off);
  };
  int rc = LDS_OK;
  // tensor-core levels: one power-of-two scale for their resblock filters (planes.cuh; lds_api.cu does the same for the denoiser)
  PlanePacker pkh;
  std::vector<std::pair<const __nv_bfloat16**, size_t>> fixh;
  {
    float wmax = 0.f;
    for (const auto& kv : v->raw)
      if ((kv.first.rfind("resblocks.", 0) == 0 || kv.first.rfind("ups.", 0) == 0 || kv.first.rfind("conv_pre.", 0) == 0) && kv.second.second.size() == 3)
        for (float x : kv.second.first) wmax = std::max(wmax, std::fabs(x));
    float sc = 4096.f;
    while (sc > 1.f && wmax * sc >= 16384.f) sc *= 0.5f;
    v->wscale = pkh.scale = sc;
  }
  // Conv1d weight [cout, cin, k] -> [cin][k][cout];  ConvTranspose1d weight [cin, cout, k] -> [cin][k][cout]
  auto conv = [&](ConvP& p, const std::string& key, int cin, int cout, int k, bool transposed, int up_rate = 0, int dil = 1) -> bool {
    auto itw = v->raw.find(key + ".weight"), itb = v->raw.find(key + ".bias");
    if (itw == v->raw.end() || itb == v->raw.end()) { rc = vfail(LDS_ERR_MISSING, "weight '%s.weight' / '.bias' was not loaded", key.c_str()); return false; }
    const std::vector<int64_t> want = transposed ? std::vector<int64_t>{cin, cout, k} : std::vector<int64_t>{cout, cin, k};
    if (itw->second.second != want || itb->second.second != std::vector<int64_t>{cout}) {
      rc = vfail(LDS_ERR_INVALID, "weight '%s' has an unexpected shape (expected [%d,%d,%d])", key.c_str(), (int)want[0], (int)want[1], k);
      return false;
    }
    const float* src = itw->second.first.data();
    std::vector<float> t((size_t)cin * k * cout);
    for (int ci = 0; ci < cin; ++ci)
      for (int kk = 0; kk < k; ++kk)
        for (int co = 0; co < cout; ++co)
          t[((size_t)ci * k + kk) * cout + co] = transposed ? src[((size_t)ci * cout + co) * k + kk] : src[((size_t)co * cin + ci) * k + kk];
    put(&p.w, t.data(), t.size());
    put(&p.b, itb->second.first.data(), (size_t)cout);
    p.cin = cin; p.cout = cout; p.k = k;
    if (transposed && up_tc_ok(cin, cout, k, up_rate)) {
      // ConvTranspose1d(stride u, kernel 2u, padding u/2) as ONE 3-tap implicit GEMM with N = u * cout: output frame lo = q*u + s of
      // input frame q is  s <  u - pad:  W[.., s+pad] x[q] + W[.., s+pad+u] x[q-1]
      //                  s >= u - pad:  W[.., s+pad-u] x[q+1] + W[.., s+pad] x[q]
      // (lo = li*u - pad + kk  =>  kk = (lo + pad) mod u (+u), li = (lo + pad - kk) / u), so a GEMM row q of width [u][cout] IS the u
      // output frames of q in channels-last order; taps (x[q-1], x[q], x[q+1]) = gemm_tc's k3 convolution, a third of the blocks zero.
      const int u = up_rate, pad = u / 2;
      std::vector<float> tt((size_t)u * cout * 3 * cin, 0.f), bb((size_t)u * cout);
      for (int sidx = 0; sidx < u; ++sidx)
        for (int co = 0; co < cout; ++co) {
          const size_t n = (size_t)sidx * cout + co;
          bb[n] = itb->second.first[co];
          for (int ci = 0; ci < cin; ++ci) {
            const float* wk = src + ((size_t)ci * cout + co) * k;
            if (sidx < u - pad) {
              tt[(n * 3 + 1) * cin + ci] = wk[sidx + pad];
              tt[(n * 3 + 0) * cin + ci] = wk[sidx + pad + u];
            } else {
              tt[(n * 3 + 2) * cin + ci] = wk[sidx + pad - u];
              tt[(n * 3 + 1) * cin + ci] = wk[sidx + pad];
            }
          }
        }
      fixh.emplace_back(&p.wh, pkh.add(tt.data(), (size_t)u * cout * 3, cin, 2));
      put(&p.b_rep, bb.data(), bb.size());
    }
    if (!transposed && key.rfind("resblocks.", 0) == 0 && tc_level(cin) && cin == cout && cin < 64) {   // time-folded form
      const int f = fold_of(cin), cc = (k - 1) / 2;
      std::vector<int> offs;
      for (int so = 0; so < f; ++so)
        for (int tp = 0; tp < k; ++tp) {
          const int o = floordiv(so + (tp - cc) * dil, f);
          if (std::find(offs.begin(), offs.end(), o) == offs.end()) offs.push_back(o);
        }
      std::sort(offs.begin(), offs.end());
      if ((int)offs.size() <= 24) {
        const int nt = (int)offs.size(), W = 64;
        std::vector<float> tt((size_t)W * nt * W, 0.f), bb((size_t)W);
        for (int so = 0; so < f; ++so)
          for (int co = 0; co < cout; ++co) {
            const size_t n = (size_t)so * cout + co;
            bb[n] = itb->second.first[co];
            for (int j = 0; j < nt; ++j)
              for (int si = 0; si < f; ++si) {
                const int delta = f * offs[j] + si - so;
                if (delta % dil) continue;
                const int tp = delta / dil + cc;
                if (tp < 0 || tp >= k) continue;
                for (int ci = 0; ci < cin; ++ci) tt[(n * nt + j) * W + (size_t)si * cin + ci] = src[((size_t)co * cin + ci) * k + tp];
              }
          }
        fixh.emplace_back(&p.wh, pkh.add(tt.data(), (size_t)W * nt, W, 2));
        put(&p.b_rep, bb.data(), bb.size());
        p.fold = f; p.ntaps = nt;
        for (int j = 0; j < nt; ++j) p.tap_off[j] = offs[j];
      }
    } else if (!transposed && ((key.rfind("resblocks.", 0) == 0 && tc_level(cin) && cin == cout) ||
                               (key == "conv_pre" && cin % 64 == 0 && cout % 64 == 0 && k % 2 == 1 && k <= 11))) {   // [cout][tap][plane][cin] for gemm_tc
      std::vector<float> tt((size_t)cout * k * cin);
      for (int co = 0; co < cout; ++co)
        for (int ci = 0; ci < cin; ++ci)
          for (int kk = 0; kk < k; ++kk) tt[((size_t)co * k + kk) * cin + ci] = src[((size_t)co * cin + ci) * k + kk];
      fixh.emplace_back(&p.wh, pkh.add(tt.data(), (size_t)cout * k, cin, 2));
    }
    return true;
  };
  const int nrb = c.n_ups * c.n_kernels;
  v->ups.assign(c.n_ups, ConvP());
  v->rb1.assign(nrb, std::vector<ConvP>());
  v->rb2.assign(nrb, std::vector<ConvP>());
  bool ok = conv(v->conv_pre, "conv_pre", c.inter_channels, c.upsample_initial_channel, 7, false);
  int ch = c.upsample_initial_channel;
  for (int i = 0; ok && i < c.n_ups; ++i) {
    ok = conv(v->ups[i], "ups." + std::to_string(i), ch, ch / 2, c.upsample_kernel_sizes[i], true, c.upsample_rates[i]);
    ch /= 2;
    for (int j = 0; ok && j < c.n_kernels; ++j) {
      const int n = i * c.n_kernels + j, k = c.resblock_kernel_sizes[j];
      const std::string rk = "resblocks." + std::to_string(n);
      const int nconv = c.resblock_kind == 1 ? 3 : 2;
      v->rb1[n].assign(nconv, ConvP());
      if (c.resblock_kind == 1) v->rb2[n].assign(nconv, ConvP());
      for (int q = 0; ok && q < nconv; ++q) {
        if (c.resblock_kind == 1) {
          ok = conv(v->rb1[n][q], rk + ".convs1." + std::to_string(q), ch, ch, k, false, 0, c.resblock_dilations[j][q]) &&
               conv(v->rb2[n][q], rk + ".convs2." + std::to_string(q), ch, ch, k, false, 0, 1);
        } else {
          ok = conv(v->rb1[n][q], rk + ".convs." + std::to_string(q), ch, ch, k, false, 0, c.resblock_dilations[j][q]);
        }
      }
    }
  }
  ok = ok && conv(v->conv_post, "conv_post", ch, 1, 7, false);
  if (!ok) return rc != LDS_OK ? rc : vfail(LDS_ERR_MISSING, "vocoder weights incomplete");
  v->warena_floats = host.size();
  if (cudaMalloc(&v->warena, host.size() * sizeof(float)) != cudaSuccess) return vfail(LDS_ERR_CUDA, "cudaMalloc of the vocoder weights failed");
  if (cudaMemcpy(v->warena, host.data(), host.size() * sizeof(float), cudaMemcpyHostToDevice) != cudaSuccess)
    return vfail(LDS_ERR_CUDA, "upload of the vocoder weights failed");
  for (auto& f : fix) *f.first = v->warena + f.second;
  if (!pkh.host.empty()) {
    if (cudaMalloc(&v->wharena, pkh.host.size() * sizeof(uint16_t)) != cudaSuccess ||
        cudaMemcpy(v->wharena, pkh.host.data(), pkh.host.size() * sizeof(uint16_t), cudaMemcpyHostToDevice) != cudaSuccess)
      return vfail(LDS_ERR_CUDA, "upload of the vocoder's tensor-core filters failed");
    for (auto& f : fixh) *f.first = v->wharena + f.second;
  }
  v->raw.clear();
  v->finalized = true;
  return LDS_OK;
}

int lds_vocoder_hop(const lds_vocoder* v) {
  if (!v) return 0;
  int hop = 1;
  for (int i = 0; i < v->cfg.n_ups; ++i) hop *= v->cfg.upsample_rates[i];
  return hop;
}

int lds_vocode(lds_vocoder* v, const float* mel_BTC, int B, int T, float* wav_BL, void* stream) {
  if (!v || !v->finalized) return vfail(LDS_ERR_INVALID, "lds_vocode: weights not finalized");
  if (!mel_BTC || !wav_BL || B < 1 || T < 1) return vfail(LDS_ERR_INVALID, "lds_vocode: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (cudaSetDevice(v->device) != cudaSuccess) return vfail(LDS_ERR_CUDA, "cudaSetDevice failed");
  const lds_vocoder_config& c = v->cfg;
  // workspace: z^T and conv_pre output, then per level: level input x, xs, and three scratch tensors of the level's size
  size_t need = (size_t)B * c.inter_channels * T + (size_t)B * c.upsample_initial_channel * T;
  size_t lvl_max = 0, tc_max = 0;
  {
    int ch = c.upsample_initial_channel;
    int64_t L = T;
    for (int i = 0; i < c.n_ups; ++i) {
      ch /= 2; L *= c.upsample_rates[i];
      lvl_max = std::max(lvl_max, (size_t)B * ch * (size_t)L);
    }
    // channels-last tensors of the tensor-core path: a level's input / output, or conv_pre's output ahead of the first transposed convolution
    tc_max = std::max(lvl_max, (size_t)B * c.upsample_initial_channel * (size_t)T);
  }
  need += 6 * (lvl_max + 64) + 4 * (tc_max + 64) + 1024;   // tensor-core path: two channels-last fp32 tensors + two split-f16 plane buffers
  if (need > v->arena_cap) {           // grow-only; growing synchronises (work in flight may still use the old arena)
    if (cudaDeviceSynchronize() != cudaSuccess) return vfail(LDS_ERR_CUDA, "synchronize before workspace growth failed");
    if (v->arena) { cudaFree(v->arena); v->arena = nullptr; v->arena_cap = 0; }
    if (cudaMalloc(&v->arena, need * sizeof(float)) != cudaSuccess) return vfail(LDS_ERR_CUDA, "cudaMalloc of %zu MB vocoder workspace failed", need * 4 >> 20);
    v->arena_cap = need;
  }
  size_t off = 0;
  auto take = [&](size_t n) { float* p = v->arena + off; off += (n + 63) / 64 * 64; return p; };
  float* zt = take((size_t)B * c.inter_channels * T);
  float* pre = take((size_t)B * c.upsample_initial_channel * T);
  float* lv[6];
  for (int i = 0; i < 6; ++i) lv[i] = take(lvl_max);
  float* x_cl = take(tc_max);
  float* xs_cl = take(tc_max);
  __nv_bfloat16* a_p = reinterpret_cast<__nv_bfloat16*>(take(tc_max));      // 2 planes x 2 bytes = one float per element
  __nv_bfloat16* t_p = reinterpret_cast<__nv_bfloat16*>(take(tc_max));
  double flops = 0;
  auto ck = [&](cudaError_t e, const char* what) -> int {
    ++v->launches;
    return e == cudaSuccess ? LDS_OK : vfail(LDS_ERR_CUDA, "launch of %s failed: %s", what, cudaGetErrorString(e));
  };
#define VTRY(expr) do { int rc__ = (expr); if (rc__ != LDS_OK) return rc__; } while (0)
  if (!(v->conv_pre.wh && v->ups[0].wh)) {  // Hifi_VAEGAN.forward: z = mel.transpose(-1, -2)
    dim3 grid((c.inter_channels + 31) / 32, (T + 31) / 32, B), block(32, 8);
    voc_transpose_kernel<<<grid, block, 0, s>>>(mel_BTC, zt, T, c.inter_channels);
    VTRY(ck(cudaGetLastError(), "voc_transpose"));
  }
  const bool pre_tc = v->conv_pre.wh && v->ups[0].wh;     // conv_pre on the tensor cores: the mel [B, T, C] IS its channels-last operand
  if (!pre_tc) VTRY(ck(launch_conv(zt, v->conv_pre, nullptr, pre, B, T, 1, 1.f, 0, 1.f, 0, s), "conv_pre"));
  flops += 2.0 * B * T * c.inter_channels * c.upsample_initial_channel * 7;
  // The state between levels is either channels-first (x_cf: the FFMA kernels) or channels-last (x_cl_cur: the tensor-core kernels);
  // a transpose is inserted only where the form changes (after conv_pre, and before the first level that stays on the CUDA cores).
  const float* x_cf = pre;
  const float* x_cl_cur = nullptr;
  int ch = c.upsample_initial_channel;
  int64_t L = T;
  auto conv_tc = [&](const __nv_bfloat16* A, int rows_per_utt, int cin, int taps, int dil, const __nv_bfloat16* W, int N, const float* bias, int epi,
                     const float* R, float* out_f32, __nv_bfloat16* out_planes, const int* tap_rows = nullptr) {
    TcGemmArgs g;
    g.A = A; g.batches = B; g.rows = rows_per_utt; g.cin = cin; g.taps = taps; g.dil = dil; g.W = W; g.N = N; g.bias = bias; g.tap_rows = tap_rows;
    tc_set_split_pairs(g);
    g.out_scale = 1.f / (PLANE_SCALE * v->wscale);
    g.epilogue = epi; g.act_slope = 0.1f;
    if (out_planes) { g.C = out_planes; g.c_ld = 2 * N; g.out_kind = 2; }
    else { g.C = out_f32; g.c_ld = N; g.out_kind = 0; g.R = R; g.r_ld = N; }
    return launch_gemm_tc(g, s);
  };
  if (pre_tc) {                  // x = conv_pre(z) as a k7 implicit GEMM straight from the mel frames (no transpose either side)
    VTRY(ck(launch_split_cast(mel_BTC, a_p, (int64_t)B * T, c.inter_channels, 2, s), "split_cast"));
    VTRY(ck(conv_tc(a_p, T, c.inter_channels, v->conv_pre.k, 1, v->conv_pre.wh, ch, v->conv_pre.b, EPI_NONE, nullptr, xs_cl, nullptr), "conv_pre (tc)"));
    x_cl_cur = xs_cl;
    x_cf = nullptr;
  }
  for (int i = 0; i < c.n_ups; ++i) {
    const int u = c.upsample_rates[i], k = c.upsample_kernel_sizes[i];
    const int64_t Lo = (L - 1) * u - 2 * ((k - u + 1) / 2) + k;
    float* xin = lv[0];          // level input, channels-first (output of the FFMA transposed convolution)
    float* xs = lv[1];           // running sum / mean of the resblocks, channels-first
    bool level_tc = tc_level(ch / 2) && Lo % fold_of(ch / 2) == 0;
    for (int j = 0; level_tc && j < c.n_kernels; ++j) {            // every convolution of the level has its tensor-core form
      for (const ConvP& w : v->rb1[(size_t)i * c.n_kernels + j]) level_tc = level_tc && w.wh;
      for (const ConvP& w : v->rb2[(size_t)i * c.n_kernels + j]) level_tc = level_tc && w.wh;
    }
    bool xin_is_cl = false;      // the level input lives in x_cl (channels-last) instead of xin
    if (v->ups[i].wh) {          // x = ups[i](leaky_relu(x, 0.1)) as a 3-tap implicit GEMM: [B*L, ch] -> [B*L, u*ch/2] == [B*Lo, ch/2]
      if (!x_cl_cur) {
        VTRY(ck(launch_transpose_bct_to_btc(x_cf, xs_cl, B, ch, (int)L, 1.f, s), "transpose to channels-last"));
        x_cl_cur = xs_cl;
      }
      VTRY(ck(launch_lrelu_split_cast(x_cl_cur, a_p, (int64_t)B * L, ch, 2, 0.1f, s), "lrelu_split_cast"));
      VTRY(ck(conv_tc(a_p, (int)L, ch, 3, 1, v->ups[i].wh, u * (ch / 2), v->ups[i].b_rep, EPI_NONE, nullptr, x_cl, nullptr), "conv_transpose (tc)"));
      flops += 2.0 * B * Lo * ch * (ch / 2) * ((double)k / u);
      xin_is_cl = true;
    } else {
      if (x_cl_cur) {
        VTRY(ck(launch_transpose_btc_to_bct(x_cl_cur, lv[5], B, ch, (int)L, 1.f, s), "transpose to channels-first"));
        x_cf = lv[5];
      }
      dim3 grid((unsigned)((Lo + 127) / 128), (ch / 2 + 15) / 16, B);
      voc_convtr1d_kernel<<<grid, 128, 0, s>>>(x_cf, v->ups[i].w, v->ups[i].b, xin, ch, ch / 2, (int)L, (int)Lo, k, u, (k - u + 1) / 2, 0.1f);
      VTRY(ck(cudaGetLastError(), "conv_transpose"));
      flops += 2.0 * B * Lo * ch * (ch / 2) * ((double)k / u);
    }
    ch /= 2;
    L = Lo;
    if (level_tc) {
      // ---- tensor-core level: channels-last state, dilated convolutions as implicit GEMMs (gemm_tc.cu) ----
      const int f = fold_of(ch), Cw = ch * f, Lw = (int)(L / f);          // GEMM view: [B * Lw, Cw] (Cw = 64 on a folded level)
      const int64_t rows = (int64_t)B * Lw;
      if (!xin_is_cl) VTRY(ck(launch_transpose_bct_to_btc(xin, x_cl, B, ch, (int)L, 1.f, s), "transpose to channels-last"));
      for (int j = 0; j < c.n_kernels; ++j) {
        const int n = i * c.n_kernels + j;
        const float* cur = x_cl;
        const int nconv = (int)v->rb1[n].size();
        for (int q = 0; q < nconv; ++q) {
          const int d = c.resblock_dilations[j][q];
          const bool last = q == nconv - 1;
          // resblock j keeps its running x in ONE buffer (xs_cl for j = 0, whose result starts the mean; lv[4] otherwise): the first
          // convolution pair reads the level input x_cl as its residual, every later one updates the buffer in place — its residual is
          // then a TMA reduce-add in gemm_tc (no residual load)
          float* dst = j == 0 ? xs_cl : lv[4];
          (void)last;
          VTRY(ck(launch_lrelu_split_cast(cur, a_p, rows, Cw, 2, 0.1f, s), "lrelu_split_cast"));
          auto rconv = [&](const __nv_bfloat16* A, const ConvP& w, int dil, int epi, const float* R, float* out_f32, __nv_bfloat16* out_planes) {
            return f > 1 ? conv_tc(A, Lw, Cw, w.ntaps, 1, w.wh, Cw, w.b_rep, epi, R, out_f32, out_planes, w.tap_off)
                         : conv_tc(A, Lw, Cw, w.k, dil, w.wh, Cw, w.b, epi, R, out_f32, out_planes);
          };
          if (c.resblock_kind == 1) {     // xt = c2(lrelu(c1(lrelu(x)))); x = xt + x   (models.py:186-193)
            VTRY(ck(rconv(a_p, v->rb1[n][q], d, EPI_LRELU, nullptr, nullptr, t_p), "resblock conv1 (tc)"));
            VTRY(ck(rconv(t_p, v->rb2[n][q], 1, EPI_NONE, cur, dst, nullptr), "resblock conv2 (tc)"));
            flops += 2.0 * 2.0 * B * L * ch * ch * c.resblock_kernel_sizes[j];
          } else {                        // xt = c(lrelu(x)); x = xt + x   (models.py:214-217)
            VTRY(ck(rconv(a_p, v->rb1[n][q], d, EPI_NONE, cur, dst, nullptr), "resblock conv (tc)"));
            flops += 2.0 * B * L * ch * ch * c.resblock_kernel_sizes[j];
          }
          cur = dst;
        }
        if (j > 0) {                      // xs += r_j ; the last one also divides by num_kernels (models.py:243-251)
          const int64_t n4 = rows * Cw / 4;
          voc_acc_kernel<<<(unsigned)std::min<int64_t>((n4 + 255) / 256, 2368), 256, 0, s>>>(
              reinterpret_cast<float4*>(xs_cl), reinterpret_cast<const float4*>(lv[4]), n4, j == c.n_kernels - 1 ? 2 : 1, (float)c.n_kernels);
          VTRY(ck(cudaGetLastError(), "resblock mean"));
        }
      }
      x_cl_cur = xs_cl;                   // stays channels-last for the next transposed convolution
      x_cf = nullptr;
      continue;
    }
    // ---- CUDA-core level (channels-first) ----
    if (xin_is_cl) VTRY(ck(launch_transpose_btc_to_bct(x_cl, xin, B, ch, (int)L, 1.f, s), "transpose to channels-first"));
    for (int j = 0; j < c.n_kernels; ++j) {
      const int n = i * c.n_kernels + j;
      // xs = r_0; xs += r_j; x = xs / num_kernels   (models.py:243-251; a single resblock: x = r_0 / 1 = r_0)
      const int last_mode = j == 0 ? 0 : (j == c.n_kernels - 1 ? 2 : 1);
      const float* cur = xin;
      const int nconv = (int)v->rb1[n].size();
      for (int q = 0; q < nconv; ++q) {
        const int d = c.resblock_dilations[j][q];
        const bool last = q == nconv - 1;
        float* dst = last ? xs : (cur == lv[2] ? lv[3] : lv[2]);
        if (c.resblock_kind == 1) {
          VTRY(ck(launch_conv(cur, v->rb1[n][q], nullptr, lv[4], B, (int)L, d, 0.1f, 0, 1.f, 0, s), "resblock conv1"));
          VTRY(ck(launch_conv(lv[4], v->rb2[n][q], cur, dst, B, (int)L, 1, 0.1f, last ? last_mode : 0, (float)c.n_kernels, 0, s), "resblock conv2"));
          flops += 2.0 * 2.0 * B * L * ch * ch * c.resblock_kernel_sizes[j];
        } else {
          VTRY(ck(launch_conv(cur, v->rb1[n][q], cur, dst, B, (int)L, d, 0.1f, last ? last_mode : 0, (float)c.n_kernels, 0, s), "resblock conv"));
          flops += 2.0 * B * L * ch * ch * c.resblock_kernel_sizes[j];
        }
        cur = dst;
      }
    }
    // next level reads xs; keep it out of the way of the next level's scratch by swapping roles
    std::swap(lv[1], lv[5]);
    x_cf = lv[5];
    x_cl_cur = nullptr;
  }
  const float* x = x_cf;
  if (!x) {                               // the last level ran on the tensor cores: conv_post (FFMA) reads channels-first
    VTRY(ck(launch_transpose_btc_to_bct(x_cl_cur, lv[0], B, ch, (int)L, 1.f, s), "transpose to channels-first"));
    x = lv[0];
  }
  VTRY(ck(launch_conv(x, v->conv_post, nullptr, wav_BL, B, (int)L, 1, 0.01f, 0, 1.f, 1, s), "conv_post"));
  flops += 2.0 * B * L * ch * 7;
  v->flops_last = flops;
#undef VTRY
  return LDS_OK;
}

int64_t lds_vocoder_launches(const lds_vocoder* v) { return v ? v->launches : 0; }
double lds_vocoder_last_flops(const lds_vocoder* v) { return v ? v->flops_last : 0; }
int64_t lds_vocoder_workspace_bytes(const lds_vocoder* v) { return v ? (int64_t)((v->arena_cap + v->warena_floats) * sizeof(float)) : 0; }

}  // extern "C"
