// lds_kernels.h — launch wrappers of the hand-written sm_100a kernels (internal header).
// Activations are channels-last: a tensor of a U-Net level is [B*T_l, C] row-major fp32
// (bf16 for tensor-core operands in LDS_PREC_BF16 mode).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

namespace lds {

// Programmatic dependent launch (PDL).  The hot kernels of the denoiser call pdl_trigger() first (the next kernel of the
// stream may start launching as soon as every CTA of this one has started) and pdl_wait() before their first access to
// global memory (blocks until the previous kernel has completed and flushed), and are launched with
// cudaLaunchAttributeProgrammaticStreamSerialization: launch latency and the next kernel's prologue (barrier init, TMEM
// allocation, tensor-map fetch) then hide behind the tail wave of the previous kernel.  LDS_PDL=0 turns the attribute off.
//
// Run-time switches (A/B measurements only; none selects a CPU path or changes results): read ONCE per process, on the
// first call of knobs() — lds_create calls it — from the environment of the process that loads the library.
struct Knobs {
  int gn_mode = 2;      // LDS_GN_MODE   2 persistent cluster GroupNorm, 1 one CTA per slab, 0 stats + apply
  bool pdl = true;      // LDS_PDL       programmatic dependent launch attribute on the hot kernels
  bool tma_epi = true;  // LDS_TMA_EPI   fp32 GEMM outputs through the TMA epilogue (0: per-thread ld.global / st.global epilogue)
  bool ff2_inplace = true;    // LDS_FF2_INPLACE  FF2 as an in-place fp32 reduce-add GEMM + one cast pass (0: 16-bit plane output with a register-path residual)
  bool att_dual64 = true;     // LDS_ATT_DUAL64   split-f16 attention at head dim 33..64: two CTAs per SM with single-stage K / V^T rings (0: one CTA, double-buffered)
  bool red_add = true;  // LDS_RED_ADD   in-place residual GEMMs store through a TMA reduce-add (0: TMA-load the residual, add, TMA-store)
};
const Knobs& knobs();
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
inline bool pdl_enabled() { return knobs().pdl; }
// Kernel attributes (opt-in shared memory) are per device: true exactly once per (call site, device of the calling thread).
inline bool first_use_on_this_device(unsigned long long& seen) {
  int d = 0;
  cudaGetDevice(&d);
  const unsigned long long bit = 1ull << (d & 63);
  if (seen & bit) return false;
  seen |= bit;
  return true;
}
// cudaLaunchKernelEx with the PDL attribute (and an optional cluster size)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

enum Epilogue : int { EPI_NONE = 0, EPI_SILU = 1, EPI_GEGLU = 2, EPI_GELU = 3, EPI_LRELU = 4 };   // EPI_GELU (exact-erf) / EPI_LRELU (slope act_slope): tensor-core path only

// C[M,N] = epi( A (*) W^T + bias ) + R.
//  A: source rows [*, a_ld]; logical K = taps * cin.  taps==1: plain GEMM row m <- source row m.
//  taps==3: 1-D convolution along time, output row (b,to) reads source frames
//           u = to*stride - 1 + tap (zero outside [0, t_conv)), where the convolution input is the
//           source itself (t_conv = t_in) or its nearest-neighbour upsampling to t_conv frames
//           (src = min(floor(u*up_scale), t_in-1), F.interpolate semantics, resnet.py:157-160).
//  W: [N, K] row-major, K index = tap*cin + c.
//  bias: [N].
//  R (optional): row (m / r_div) of [*, r_ld] is added after the activation (residual / shortcut /
//                partial sums; r_div = frames per utterance broadcasts one row per utterance).
//  EPI_GEGLU: W rows are interleaved per 128-column tile as [64 value | 64 gate]; output has N/2 columns.
struct GemmArgs {
  const float* A = nullptr; int a_ld = 0;
  const float* W = nullptr;
  float* C = nullptr; int c_ld = 0;
  const float* bias = nullptr;
  const float* R = nullptr; int r_ld = 0; int r_div = 1;
  int M = 0, N = 0, K = 0;
  int taps = 1, cin = 0;
  int t_out = 1, t_in = 1, t_conv = 1, stride = 1, upsample = 0; float up_scale = 1.f;
  int epilogue = EPI_NONE;
};
cudaError_t launch_gemm_f32(const GemmArgs& a, cudaStream_t s);

// tcgen05/TMEM/TMA implicit GEMM on bf16 operands (gemm_tc.cu).
//  A: 16-bit planes [batches][rows][a_parts*cin]; W: [N][taps*w_parts*cin] (planes.cuh).  Plain bf16: parts = 1, one product.
//  Split-f16 (fp32-accurate): parts = 2 (fp16 planes h1, h2 of the scaled operand) and the three products h1*w2, h2*w1 (into a
//  "small" TMEM accumulator) and h1*w1 (into the main one); the epilogue adds the two and multiplies by out_scale =
//  1 / (scale_A * scale_W) — three tensor-core products per logical product.
//  taps in {3, 5, 7, 9, 11}: stride-1 'same' convolution along `rows` with dilation `dil` (tap t reads row r + (t - (taps-1)/2) * dil;
//  TMA out-of-bounds fill = zero padding).
//  out_kind: 0 fp32 [*, c_ld], 1 bf16 [*, c_ld], 2 split-f16 planes [h1 | h2] of c_ld/2 columns each,
//            3 attention operands of a fused QKV projection with N = 3*H*dpad columns ([q | k | v], head dim padded
//              to dpad): q_out/k_out planes [rows][parts][H*dpad], vt_out planes [B][parts][H][dpad][T_pad] (V transposed,
//              keys contiguous) — the layouts attention_tc.cu loads with TMA.
struct TcGemmArgs {
  const __nv_bfloat16* A = nullptr; int batches = 1, rows = 0, cin = 0, a_parts = 1;
  const __nv_bfloat16* W = nullptr; int N = 0, taps = 1, w_parts = 1, dil = 1;
  const int* tap_rows = nullptr;          // optional (host): row offset of every tap, taps <= 24 of any parity — replaces (t - (taps-1)/2) * dil
                                          // (time-folded convolutions of the vocoder's narrow levels: non-uniform tap spacing)
  int n_pairs = 1; int pair_a[6] = {0, 0, 0, 0, 0, 0}; int pair_w[6] = {0, 0, 0, 0, 0, 0};
  const float* bias = nullptr;
  const float* R = nullptr; int r_ld = 0, r_div = 1;
  void* C = nullptr; int c_ld = 0, out_kind = 0, epilogue = EPI_NONE;
  float out_scale = 1.f;                  // applied to the accumulator before bias / activation (split-f16: 1 / (scale_A * scale_W))
  float act_slope = 0.f;                  // EPI_LRELU: negative-side slope
  __nv_bfloat16 *q_out = nullptr, *k_out = nullptr, *vt_out = nullptr; int att_T = 0, att_H = 0, att_dpad = 0, att_Tpad = 0;
  int att_parts = 1;                      // planes of the attention operands written by out_kind 3 (bf16: 1 or 3)
};
cudaError_t launch_gemm_tc(const TcGemmArgs& a, cudaStream_t s);

// tcgen05 flash attention (attention_tc.cu) on the operands written by a QKV projection with out_kind 3.
// out: operand planes [B*T][out_parts*H*d] for the following GEMM (out_parts 1: bf16; 2: split-f16; 3: three bf16 planes).
struct AttnTcArgs {
  const __nv_bfloat16 *q = nullptr, *k = nullptr, *vt = nullptr;
  __nv_bfloat16* out = nullptr;
  int B = 0, T = 0, T_pad = 0, H = 0, d = 0, dpad = 0, parts = 1, out_parts = 0;   // out_parts 0: same as parts
};
cudaError_t launch_attention_tc(const AttnTcArgs& a, cudaStream_t s);
inline void tc_set_split_pairs(TcGemmArgs& a) {   // h1*w2, h2*w1, h1*w1
  static const int pa[3] = {0, 1, 0}, pw[3] = {1, 0, 0};
  a.a_parts = a.w_parts = 2; a.n_pairs = 3;
  for (int i = 0; i < 3; ++i) { a.pair_a[i] = pa[i]; a.pair_w[i] = pw[i]; }
}
// fp32 [rows, C] -> bf16 [rows, parts*C]: parts=1 plain rounding; parts=3 hi/mid/lo planes (x == hi+mid+lo to 24 bits)
cudaError_t launch_split_cast(const float* in, __nv_bfloat16* out, int64_t rows, int C, int parts, cudaStream_t s);
// the same with leaky_relu(x, slope) applied first (the pre-activation of the vocoder's dilated convolutions)
cudaError_t launch_lrelu_split_cast(const float* in, __nv_bfloat16* out, int64_t rows, int C, int parts, float slope, cudaStream_t s);
// Gathering casts for the two resampling convolutions (channels-last [B, t, C] fp32 -> bf16 planes):
//  mode 1: nearest upsample to t_out frames, out [B, t_out, parts*C], src = min(floor(u*scale), t_in-1)   (resnet.py:157-160)
//  mode 2: k=3 / stride-2 / pad-1 im2col, out [B, t_out, parts*3C] with element (p, tap, c) at p*3C + tap*C + c  (resnet.py:200)
cudaError_t launch_cast_gather(const float* in, __nv_bfloat16* out, int B, int t_in, int t_out, int C, int parts, int mode,
                               float scale, cudaStream_t s);

// Flash-style self-attention on a fused QKV buffer [B*T, 3C]: head h uses columns
// [h*d,(h+1)*d) of each C-wide third.  out [B*T, C].  softmax(q k^T / sqrt(d)) v, no mask.
// Output: fp32 `out` [B*T, C], or (outb != nullptr) bf16 operand planes [B*T, parts*C] for the following tensor-core GEMM.
cudaError_t launch_attention_f32(const float* qkv, float* out, __nv_bfloat16* outb, int parts, int B, int T, int C, int heads,
                                 cudaStream_t s);

// GroupNorm over a (virtually concatenated) channels-last tensor [x1 | x2].
//  stats : partial (count, mean, M2) per (b, chunk of GN_ROWS frames, group) -> part[B][nchunk][G][3]
//  apply : y = GN(x)*gamma+beta ; optional y = y*(1+scale[c])+shift[c] (ss = [scale(C) | shift(C)]); optional SiLU
constexpr int GN_ROWS = 32;
cudaError_t launch_gn_stats(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float* part,
                            cudaStream_t s);
cudaError_t launch_gn_apply(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups,
                            const float* part, float eps, const float* gamma, const float* beta, const float* ss,
                            int silu, float* y, __nv_bfloat16* yb, int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride = 0);
// (ss_bstride: floats between the scale-shift rows of consecutive utterances — per-utterance timesteps of the training-loss forward;
//  0 = one row for the whole batch, the sampler's case)
// single-pass variant (slab of one (utterance, group) in shared memory); cudaErrorNotSupported when it does not fit
cudaError_t launch_gn_fused(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                            const float* gamma, const float* beta, const float* ss, int silu, float* y, __nv_bfloat16* yb,
                            int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride = 0);
// cluster variant: the slab is split along time over a thread-block cluster of 1..8 CTAs (partials exchanged through
// distributed shared memory) — small CTAs, several per SM, and slabs up to 8 x 200 KB stay single-pass
cudaError_t launch_gn_cluster(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                              const float* gamma, const float* beta, const float* ss, int silu, float* y, __nv_bfloat16* yb,
                              int parts, __nv_bfloat16* rawb, cudaStream_t s, int64_t ss_bstride = 0);
//  (yb != nullptr: write bf16 operand planes [B*T, parts*C] instead of fp32 y; rawb: also copy the un-normalised
//   concat input as planes — the A operand of the 1x1 shortcut convolution)

// LayerNorm over the last dim of [rows, C].
cudaError_t launch_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int C, float* y,
                             __nv_bfloat16* yb, int parts, cudaStream_t s);

// [B, C, T] <-> [B, T, C] tiled transposes; `scale` multiplies on the way.
cudaError_t launch_transpose_bct_to_btc(const float* in, float* out, int B, int C, int T, float scale, cudaStream_t s);
cudaError_t launch_transpose_btc_to_bct(const float* in, float* out, int B, int C, int T, float scale, cudaStream_t s);
cudaError_t launch_div_copy(const float* in, float* out, int64_t n, float divisor, cudaStream_t s);
cudaError_t launch_silu(const float* in, float* out, int64_t n, cudaStream_t s);
// out[b][:] = table[idx[b]-1][:]   (nn.Embedding gather of spk_embed(spk_id - 1), unit2mel.py:82)
cudaError_t launch_spk_gather(const float* table, const int64_t* idx, int n_rows, int B, int H, float* out,
                              cudaStream_t s);

// Solver updates on channels-last state tensors of n = B*T*M elements (coefficient rows: lds_b200.h).
// x0 prediction: m = (x - sigma*eps)/alpha.
cudaError_t launch_x0_pred(const float* x, const float* eps, float sigma, float alpha, float* m, int64_t n, cudaStream_t s);
// DPM-Solver++: order 1: x = cx*x - cm*m0 ; order 2: x = cx*x - cm*m0 - hcm*(ir0*(m0-m1))
cudaError_t launch_dpm_update(float* x, const float* m0, const float* m1, float cx, float cm, float hcm, float ir0,
                              int order, int64_t n, cudaStream_t s);
// UniPC predictor: xb = cx*x - cmE*m0 ; xp = xb - aB*(rho_p*((m1-m0)/rk)) (order 2) | xb (order 1)
cudaError_t launch_unipc_predict(const float* x, const float* m0, const float* m1, float cx, float cmE, float aB,
                                 float rk, float rho_p, int order, float* xb, float* xp, int64_t n, cudaStream_t s);
// UniPC corrector: x = xb - aB*(rho_c0*((m1-m0)/rk) + rho_c1*(mt-m0))  (order 1: the first term is absent)
cudaError_t launch_unipc_correct(const float* xb, const float* m0, const float* m1, const float* mt, float aB, float rk,
                                 float rho_c0, float rho_c1, int order, float* x, int64_t n, cudaStream_t s);
// DDPM ancestral step; noise is in the reference layout [B, M, T] and is transposed on the fly.
cudaError_t launch_ddpm_step(float* x, const float* eps, const float* noise_BMT, float c_recip, float c_recipm1,
                             float pm1, float pm2, float sig, int B, int T, int M, cudaStream_t s);

// Shallow-diffusion start: x[B,T,M] = sqrt_acp*(gt[B,T,M]*acoustic_scale) + sqrt_1m_acp*noise[B,M,T]   (diffusion.py:169-171,208-212)
cudaError_t launch_q_sample(float* x, const float* gt_BTM, const float* noise_BMT, float acoustic_scale, float sqrt_acp,
                            float sqrt_1m_acp, int B, int T, int M, cudaStream_t s);

// Training loss (diffusion.py:181-185): mean |noise - eps| (l1) or mean (noise - eps)^2 (l2); eps [B,T,M] channels-last, noise [B,M,T];
// partial: scratch of B * ceil(T/32) * ceil(M/32) doubles; deterministic two-stage reduction, loss: one device float
cudaError_t launch_diffusion_loss(const float* eps_BTM, const float* noise_BMT, int B, int T, int M, int l1, double* partial, float* loss,
                                  cudaStream_t s);
// DDIM step: x = sqrt_aprev * (x / sqrt_at + coef * eps)                                   (diffusion.py:131)
cudaError_t launch_ddim_step(float* x, const float* eps, float sqrt_at, float coef, float sqrt_aprev, int64_t n, cudaStream_t s);
// PLMS step: out = x + d*(k1*x - k2*e'), e' = Adams-Bashforth combination selected by mode (solver.cu) (diffusion.py:134-167)
cudaError_t launch_pndm_update(const float* x, const float* e, const float* h1, const float* h2, const float* h3, float d,
                               float k1, float k2, int mode, float* out, int64_t n, cudaStream_t s);

}  // namespace lds
