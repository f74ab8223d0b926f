/* lds_b200.h — C ABI of the B200-native Unit2Mel diffusion-sampling library (liblds_b200.so).
 *
 * The reference (bfloat16/latent-diffusion-speech) is pure Python/PyTorch and has NO FFI or
 * plugin interface for this path; the seam is two Python call sites.  Each entry point below
 * names the reference code it replaces (file:line relative to the reference tree).  A host in
 * any language binds these with plain pointers and sizes; the shipped Python host
 * (latent_diffusion_speech_b200/capi.py) binds them with ctypes — see INTEGRATION.md.
 *
 * Conventions
 *   - every function returns LDS_OK (0) or a negative lds_status; the message of the last
 *     failure on the calling thread is available from lds_last_error().
 *   - all tensor pointers are DEVICE pointers unless a parameter says "host"; tensors are
 *     contiguous; the caller keeps ownership and must keep them alive until the stream work
 *     enqueued by the call has completed.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  All work is
 *     enqueued asynchronously on it; no call synchronises the device except lds_finalize_weights,
 *     lds_destroy and an lds_plan that has to GROW the workspace or change the sampler program
 *     (a re-plan for a batch that fits the existing workspace neither synchronises nor allocates;
 *     callers that switch streams between calls order them themselves).
 *   - a handle is bound to one device and is not thread-safe.
 *   - there is no random number generation inside the library: initial noise and per-step
 *     DDPM noise are produced by the caller (reference: torch.randn at diffusion.py:207,170,118).
 */
#ifndef LDS_B200_H_
#define LDS_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define LDS_API __attribute__((visibility("default")))
#else
#define LDS_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define LDS_VERSION 100 /* 0.1.0 */
#define LDS_MAX_BLOCKS 8

typedef enum {
  LDS_OK = 0,
  LDS_ERR_INVALID = -1,   /* bad argument / shape / state */
  LDS_ERR_CUDA = -2,      /* CUDA runtime failure (sticky per handle) */
  LDS_ERR_MISSING = -3,   /* a required weight was never loaded */
  LDS_ERR_UNSUPPORTED = -4
} lds_status;

/* LDS_PREC_FP32      : fp32-accurate ("split-f16").  Every GEMM / attention operand is held as two fp16 planes of the value
 *                      times a power of two (h1 = f16(s x), h2 = f16(s x - h1): 22 significant bits, csrc/planes.cuh) and every
 *                      product is evaluated by tcgen05.mma as h1*w1 + (h1*w2 + h2*w1) with fp32 accumulation in TMEM — three
 *                      tensor-core products per logical product; norms, softmax statistics and the solver run in fp32.
 *                      Meets max-abs <= 1e-3 against the reference fp32 sampler (tests/test_gpu_parity*.py).
 * LDS_PREC_BF16      : bf16 GEMM operands (one tcgen05.mma per K slice), fp32 accumulation, fp32 norms / solver state.
 * LDS_PREC_FP32_FFMA : IEEE fp32 FFMA kernels on the CUDA cores (the first, reference-grade implementation). */
typedef enum { LDS_PREC_FP32 = 0, LDS_PREC_BF16 = 1, LDS_PREC_FP32_FFMA = 2 } lds_precision;
typedef enum { LDS_DTYPE_F32 = 0, LDS_DTYPE_BF16 = 1, LDS_DTYPE_F16 = 2 } lds_dtype;

/* Sampler kinds.  The per-step scalar coefficients are computed by the host (they are batch
 * invariant) and handed to lds_plan as rows of LDS_COEF_STRIDE floats:
 *  LDS_SAMPLER_DPMPP_2M   (diffusion/dpm_solver_pytorch.py:1171-1213,547-576,796-831)
 *     n_rows = steps+1.  row k: [0]=sigma_k [1]=alpha_k  (x0 = (x - sigma*eps)/alpha at NFE k, k<steps)
 *                               [2]=sigma_k/sigma_{k-1} [3]=alpha_k*phi1 [4]=0.5*alpha_k*phi1 [5]=1/r0 [6]=order (k>=1)
 *  LDS_SAMPLER_UNIPC_BH2  (diffusion/uni_pc.py:471-588,606-658)
 *     n_rows = steps+1.  row k: [0]=sigma_k [1]=alpha_k [2]=sigma_k/sigma_{k-1} [3]=alpha_k*h_phi_1
 *                               [4]=alpha_k*B_h [5]=r_k [6]=order [7]=use_corrector [8]=rho_c[0] [9]=rho_c[-1] [10]=rho_p
 *  LDS_SAMPLER_DDPM       (diffusion/diffusion.py:104-121,335-341)
 *     n_rows = n_nfe.    row j: [0]=sqrt_recip_acp [1]=sqrt_recipm1_acp [2]=post_mean_coef1 [3]=post_mean_coef2
 *                               [4]=[t>0]*exp(0.5*post_log_var)   (t = k_step-1-j)
 *  LDS_SAMPLER_DDIM       (diffusion/diffusion.py:123-132,317-332)   t_j = reversed(range(0, t_total, interval))
 *     n_rows = n_nfe.    row j: [0]=sqrt(a_t) [1]=sqrt((1-a_prev)/a_prev)-sqrt((1-a_t)/a_t) [2]=sqrt(a_prev),
 *                               a_prev = alphas_cumprod[max(t-interval,0)]
 *  LDS_SAMPLER_PNDM       (diffusion/diffusion.py:134-167,300-316)   PLMS, same timesteps
 *     n_rows = n_nfe-1 (the first step evaluates the denoiser twice: at t_0 and at max(t_0-interval,0)).
 *                        row j: [0]=a_prev-a_t [1]=1/(sqrt(a_t)*(sqrt(a_t)+sqrt(a_prev)))
 *                               [2]=1/(sqrt(a_t)*(sqrt((1-a_prev)*a_t)+sqrt((1-a_t)*a_prev)))
 */
typedef enum { LDS_SAMPLER_DPMPP_2M = 0, LDS_SAMPLER_UNIPC_BH2 = 1, LDS_SAMPLER_DDPM = 2, LDS_SAMPLER_DDIM = 3,
               LDS_SAMPLER_PNDM = 4 } lds_sampler;
#define LDS_COEF_STRIDE 12

/* Mirrors Unit2Mel.__init__ (diffusion/unit2mel.py:52-71). */
typedef struct {
  int32_t input_channel;                 /* units feature dim (1280 for whisper_large_v3) */
  int32_t n_spk;                         /* <=1: no speaker embedding */
  int32_t out_dims;                      /* mel / latent bins (128) */
  int32_t n_layers;                      /* resnets per down block (2) */
  int32_t n_blocks;                      /* len(block_out_channels) (4) */
  int32_t block_out_channels[LDS_MAX_BLOCKS];
  int32_t n_heads;                       /* 8 */
  int32_t n_hidden;                      /* conditioning width (256) */
  int32_t norm_groups;                   /* 8 */
  float acoustic_scale;                  /* data.acoustic_scale */
  int32_t precision;                     /* lds_precision */
} lds_config;

typedef struct lds_handle lds_handle;

LDS_API int lds_version(void);
LDS_API const char* lds_last_error(void);

/* Replaces Unit2Mel.__init__ / UNet1DConditionModel.__init__ (unit2mel.py:52, unet_1d_condition.py:151). */
LDS_API int lds_create(const lds_config* cfg, int device, lds_handle** out);
LDS_API void lds_destroy(lds_handle* h);

/* Replaces model.load_state_dict(ckpt['model']) (unit2mel.py:31-33).  `key` is the reference
 * state_dict key (e.g. "decoder.denoise_fn.conv_in.weight"); data may be a host or device
 * pointer (copied; the caller keeps ownership).  Schedule buffers / spec_min/max are ignored. */
LDS_API int lds_load_weight(lds_handle* h, const char* key, const void* data, const int64_t* shape, int ndim, int dtype);
/* Checks that every parameter arrived and repacks them (tap-major conv filters, fused QKV,
 * value/gate-interleaved GEGLU projection, x/cond split of conv_in). */
LDS_API int lds_finalize_weights(lds_handle* h);

/* Allocates workspaces for a [B, T] batch and installs the sampler program.
 *   t_sinusoid : host, [n_nfe, block_out_channels[0]] timestep sinusoids of the n_nfe denoiser
 *                evaluations in execution order (embeddings.py:24-64; batch invariant)
 *   coefs      : host, [n_rows, LDS_COEF_STRIDE]
 * Replaces the per-call construction of NoiseScheduleVP / DPM_Solver / UniPC
 * (diffusion.py:215-299) and evaluates the time-embedding MLP and all time_emb_proj layers
 * for every step once (unet_1d_condition.py:841-848, resnet.py:614-617). */
LDS_API int lds_plan(lds_handle* h, int B, int T, int sampler, int n_nfe, const float* t_sinusoid, int n_rows,
             const float* coefs);

/* cond[B,T,n_hidden] = unit_embed(units[B,T,input_channel]) + spk_embed[spk_id-1]
 * (unit2mel.py:79-82).  spk_id: device int64 [B] (may be NULL when n_spk<=1). */
LDS_API int lds_cond(lds_handle* h, const float* units, const int64_t* spk_id, float* cond, void* stream);

/* One denoiser evaluation (UNet1DConditionModel.forward, unet_1d_condition.py:743-1036):
 * eps[B,out_dims,T] = unet(cat(x[B,out_dims,T], cond^T), t).  t_sinusoid: host [block_out_channels[0]]. */
LDS_API int lds_denoise(lds_handle* h, const float* x_BMT, const float* cond_BTH, const float* t_sinusoid, float* eps_BMT,
                void* stream);

/* The sampling loop (GaussianDiffusion.forward, infer branch, diffusion.py:203-343).
 *   lds_sample_begin : x <- x_init[B,out_dims,T] (the reference's [B,1,M,T] state), binds cond
 *   lds_sample_steps : runs program steps [k0, k1) ; step_noise[(k1-k0), B, out_dims, T] for DDPM, else NULL
 *   lds_sample_end   : mel[B,T,out_dims] = x^T / acoustic_scale   (diffusion.py:342-343)
 *   lds_sample       : begin + all steps + end
 * Step indices: DPM/UniPC k in [0, steps] (k=0 is the first evaluation), DDPM / DDIM j in [0, n_nfe), PNDM j in [0, n_nfe-1). */
LDS_API int lds_sample_begin(lds_handle* h, const float* cond_BTH, const float* x_init_BMT, void* stream);
/* Shallow-diffusion start (diffusion.py:208-212 + q_sample :169-171 + norm_spec :86): one fused kernel computes
 *   x <- sqrt_acp * (gt_spec[B,T,out_dims] * acoustic_scale) + sqrt_1m_acp * noise[B,out_dims,T]
 * (sqrt_acp = sqrt_alphas_cumprod[k_step-1], sqrt_1m_acp = sqrt_one_minus_alphas_cumprod[k_step-1]) and binds cond. */
LDS_API int lds_sample_begin_shallow(lds_handle* h, const float* cond_BTH, const float* gt_spec_BTM, const float* noise_BMT,
                             float sqrt_acp, float sqrt_1m_acp, void* stream);
LDS_API int lds_sample_steps(lds_handle* h, int k0, int k1, const float* step_noise, void* stream);
LDS_API int lds_sample_end(lds_handle* h, float* mel_BTM, void* stream);
LDS_API int lds_sample(lds_handle* h, const float* cond_BTH, const float* x_init_BMT, const float* step_noise,
               float* mel_BTM, void* stream);
LDS_API int lds_num_steps(const lds_handle* h);

/* Training-loss FORWARD (GaussianDiffusion.forward with infer=False -> p_losses, diffusion/diffusion.py:173-201; the validation loss of
 * diffusion/solver.py:56-62): one denoiser evaluation with PER-UTTERANCE timesteps and the l1 / l2 loss against the injected noise.
 *   x_noisy[b] = sqrt_acp[b] * (gt_spec[b] * acoustic_scale) + sqrt_1m_acp[b] * noise[b]         (q_sample, :169-171)
 *   eps        = unet(cat(x_noisy, cond^T), t)        t_sinusoid: host [B, block_out_channels[0]], one row per utterance
 *   *loss      = mean |noise - eps|  (loss_type 1)   or   mean (noise - eps)^2  (loss_type 2, the reference's default)
 * sqrt_acp / sqrt_1m_acp: host [B] = sqrt_alphas_cumprod[t_b] / sqrt_one_minus_alphas_cumprod[t_b]; noise [B, out_dims, T]; loss: one device
 * float; eps_BMT (optional) receives the prediction in the reference layout.  Needs lds_plan(B, T, ...) (any sampler program).  The
 * reduction is deterministic (fp64 partial sums in a fixed order).  Backward / optimiser steps are not part of this library. */
LDS_API int lds_train_loss(lds_handle* h, const float* cond_BTH, const float* gt_spec_BTM, const float* noise_BMT, const float* t_sinusoid,
                   const float* sqrt_acp, const float* sqrt_1m_acp, int loss_type, float* loss, float* eps_BMT, void* stream);

/* Introspection used by bench.py / tests. */
LDS_API int64_t lds_workspace_bytes(const lds_handle* h);
LDS_API int64_t lds_kernel_launches(const lds_handle* h);   /* kernels launched by this handle so far */
/* Per-kernel-class device time (CUDA events on the launch stream) of the LAST lds_sample /
 * lds_denoise when profiling is enabled with lds_set_profiling(h, 1); classes are listed by
 * lds_profile_class_name.  Returns milliseconds; n launches via lds_profile_class_launches. */
LDS_API int lds_set_profiling(lds_handle* h, int enabled);
LDS_API int lds_profile_num_classes(void);
LDS_API const char* lds_profile_class_name(int cls);
LDS_API double lds_profile_class_ms(lds_handle* h, int cls);
LDS_API int64_t lds_profile_class_launches(lds_handle* h, int cls);
LDS_API double lds_profile_class_flops(lds_handle* h, int cls);   /* algorithmic FLOPs issued in that class */
LDS_API double lds_profile_class_bytes(lds_handle* h, int cls);   /* algorithmic bytes moved in that class */


/* Stateless operator entry points (no handle): the individual kernels behind the sampler, exposed
 * so that each one can be parity-tested against the PyTorch op it replaces.  Activations are
 * channels-last [B*T, C] fp32; `w` is [N, taps*cin] with K index = tap*cin + c.
 *   lds_op_gemm      : F.conv1d k=1/k=3 (stride 1/2, optional fused nearest upsample) and nn.Linear
 *                      (lora.py:102, resnet.py:157-169,200-221); epilogue 0 none, 1 SiLU, 2 GEGLU
 *                      (attention.py:299-301; rows of `w`/`bias` interleaved [64 value|64 gate] per 128)
 *   lds_op_attention : F.scaled_dot_product_attention on fused qkv [B*T,3C] (attention_processor.py:1032)
 *   lds_op_groupnorm : nn.GroupNorm over the virtual concat [x1|x2] (+ (1+scale)/shift, + SiLU)
 *                      (resnet.py:597-631); part: scratch of B*ceil(T/32)*groups*3 floats
 *   lds_op_layernorm : nn.LayerNorm(C) (attention.py:83) */
LDS_API int lds_op_gemm(const float* A, int a_ld, const float* w, const float* bias, const float* R, int r_ld, int r_div,
                float* C, int c_ld, int M, int N, int K, int taps, int cin, int t_out, int t_in, int t_conv,
                int stride, int upsample, float up_scale, int epilogue, void* stream);
LDS_API int lds_op_attention(const float* qkv, float* out, int B, int T, int C, int heads, void* stream);
LDS_API int lds_op_groupnorm(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                     const float* gamma, const float* beta, const float* scale_shift, int silu, float* part,
                     float* y, void* stream);
/* Single-pass form of lds_op_groupnorm: one CTA per (utterance, group) keeps its [T, C/groups] slab in shared memory
 * (x read once).  LDS_ERR_UNSUPPORTED when the slab exceeds 200 KB; the sampler picks it automatically when it fits. */
LDS_API int lds_op_groupnorm_fused(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                           const float* gamma, const float* beta, const float* scale_shift, int silu, float* y,
                           void* stream);
/* Cluster form of the single pass: the slab is split along time over a thread-block cluster of 1..8 CTAs that exchange
 * their partial sums through distributed shared memory (<= ~40 KB per CTA, several CTAs per SM).  This is the form the
 * sampler uses; LDS_ERR_UNSUPPORTED when an eighth of the slab exceeds 200 KB (the sampler then uses stats + apply). */
LDS_API int lds_op_groupnorm_cluster(const float* x1, int c1, const float* x2, int c2, int B, int T, int groups, float eps,
                             const float* gamma, const float* beta, const float* scale_shift, int silu, float* y,
                             void* stream);
LDS_API int lds_op_layernorm(const float* x, const float* gamma, const float* beta, float eps, int rows, int C, float* y,
                     void* stream);
/* Tensor-core (tcgen05/TMEM/TMA) form of lds_op_gemm on 16-bit operand planes (csrc/planes.cuh).
 *   lds_op_split_cast : fp32 [rows, C] -> planes [rows, parts*C] (2 bytes per element): parts=1 rounds to bf16; parts=2 writes the
 *                       SPLIT-F16 form h1 = f16(16 x), h2 = f16(16 x - h1) (fp32-accurate mode); parts=3 three bf16 planes hi/mid/lo
 *   lds_op_gemm_tc    : A planes [batches][rows][parts*cin], w planes [N][taps*parts*cin] (tap-major, then plane, then channel),
 *                       both produced by lds_op_split_cast; parts=1: one bf16 product; parts=2: the three products h1*w1 (main
 *                       TMEM accumulator) and h1*w2 + h2*w1 (small accumulator), added and rescaled in the epilogue — fp32-accurate
 *                       at three tensor-core products per logical product.  taps=3: k=3/stride-1/pad-1 conv along rows.
 *                       out_kind 0 fp32, 1 bf16, 2 split-f16 planes [h1 | h2]. */
LDS_API int lds_op_split_cast(const float* in, void* out_bf16, int64_t rows, int C, int parts, void* stream);
LDS_API int lds_op_gemm_tc(const void* A_bf16, int batches, int rows, int cin, int parts, const void* w_bf16, int N, int taps,
                   const float* bias, const float* R, int r_ld, int r_div, void* C, int c_ld, int out_kind,
                   int epilogue, void* stream);
/* The dilated form of lds_op_gemm_tc: a stride-1 'same' convolution along `rows` with an odd number of taps <= 11 and dilation `dil`
 * (tap t reads row r + (t - (taps-1)/2) * dil, zero outside the utterance) — the ResBlock convolutions of the HiFi-VAEGAN generator
 * (encoder/hifi_vaegan/modules/models.py:166-184).  epilogue 0 none, 1 SiLU, 3 exact-erf GELU, 4 leaky_relu(act_slope). */
LDS_API int lds_op_conv1d_tc(const void* A_planes, int batches, int rows, int cin, int parts, const void* w_planes, int N, int taps,
                     int dil, const float* bias, const float* R, int r_ld, void* C, int c_ld, int out_kind, int epilogue,
                     float act_slope, void* stream);
/* Fused QKV projection + tcgen05 flash attention (attention_processor.py:1012-1034) on operand planes:
 *   x planes [B*T][parts*C] -> q/k/v^T scratch (layouts in lds_kernels.h) -> out planes [B*T][parts*C].  parts 1 (bf16) or 2 (split-f16:
 *   Q, K, V^T and the probabilities P are each two fp16 planes, every product is evaluated as three plane products).
 *   w_qkv: planes [3*H*dpad][parts*C], rows = [q | k | v] x heads x dpad (head dim zero-padded to dpad in {32,64}).
 *   scratch sizes (16-bit elements): q, k: B*T*parts*H*dpad; vt: B*parts*H*dpad*T_pad, T_pad = round_up(T, 8). */
LDS_API int lds_op_qkv_attention_tc(const void* x_planes, const void* w_qkv, int B, int T, int C, int H, int dpad, int parts,
                            void* q_scratch, void* k_scratch, void* vt_scratch, void* out_planes, void* stream);

/* The per-step sampler arithmetic and the layout kernels (csrc/solver.cu), one entry point per kernel, so that each is
 * testable bit-for-bit against the PyTorch expressions it replaces.  State tensors are flat fp32 arrays of n elements
 * (n % 4 == 0) in the channels-last layout [B,T,M]; the arithmetic uses round-to-nearest mul/add/sub/div in the
 * reference's evaluation order (no FMA contraction), so results are torch.equal to the eager expressions:
 *   lds_op_x0_pred       m = (x - sigma*eps)/alpha                               dpm_solver_pytorch.py:433-442, uni_pc.py:285-294
 *   lds_op_dpm_update    order 1: x = cx*x - cm*m0                                dpm_solver_pytorch.py:569-576
 *                        order 2: x = cx*x - cm*m0 - hcm*(ir0*(m0-m1))            dpm_solver_pytorch.py:813-831
 *   lds_op_unipc_predict xb = cx*x - cmE*m0; xp = xb - aB*(rho_p*((m1-m0)/rk))     uni_pc.py:545-556 (order 1: xp = xb)
 *   lds_op_unipc_correct x = xb - aB*(rho_c0*((m1-m0)/rk) + rho_c1*(mt-m0))        uni_pc.py:557-568 (order 1: first term absent)
 *   lds_op_ddpm_step     x = pm1*clamp(cr*x - crm1*eps, -1, 1) + pm2*x + sig*noise  diffusion.py:95-121 (noise [B,M,T])
 *   lds_op_ddim_step     x = sqrt_aprev*(x/sqrt_at + coef*eps)                     diffusion.py:123-132
 *   lds_op_pndm_update   out = x + d*(k1*x - k2*e'), e' by mode 0..4               diffusion.py:134-167
 *   lds_op_q_sample      x = sqrt_acp*(gt*acoustic_scale) + sqrt_1m_acp*noise^T     diffusion.py:169-171,208-212
 *   lds_op_cast_gather   mode 1 nearest upsample, mode 2 k3/s2/p1 im2col -> bf16 planes   resnet.py:157-160,200
 *   lds_op_transpose     [B,C,T] <-> [B,T,C] (* scale)                             diffusion.py:225,342
 *   lds_op_div_copy      out = in / divisor                                         diffusion.py:343 (denorm_spec :87) */
LDS_API int lds_op_x0_pred(const float* x, const float* eps, float sigma, float alpha, float* m, int64_t n, void* stream);
LDS_API int lds_op_dpm_update(float* x, const float* m0, const float* m1, float cx, float cm, float hcm, float ir0, int order,
                      int64_t n, void* stream);
LDS_API int lds_op_unipc_predict(const float* x, const float* m0, const float* m1, float cx, float cmE, float aB, float rk,
                         float rho_p, int order, float* xb, float* xp, int64_t n, void* stream);
LDS_API int lds_op_unipc_correct(const float* xb, const float* m0, const float* m1, const float* mt, float aB, float rk,
                         float rho_c0, float rho_c1, int order, float* x, int64_t n, void* stream);
LDS_API int lds_op_ddpm_step(float* x_BTM, const float* eps_BTM, const float* noise_BMT, float c_recip, float c_recipm1, float pm1,
                     float pm2, float sig, int B, int T, int M, void* stream);
LDS_API int lds_op_ddim_step(float* x, const float* eps, float sqrt_at, float coef, float sqrt_aprev, int64_t n, void* stream);
LDS_API int lds_op_pndm_update(const float* x, const float* e, const float* h1, const float* h2, const float* h3, float d, float k1,
                       float k2, int mode, float* out, int64_t n, void* stream);
LDS_API int lds_op_q_sample(float* x_BTM, const float* gt_BTM, const float* noise_BMT, float acoustic_scale, float sqrt_acp,
                    float sqrt_1m_acp, int B, int T, int M, void* stream);
LDS_API int lds_op_cast_gather(const float* in, void* out_bf16, int B, int t_in, int t_out, int C, int parts, int mode, float scale,
                       void* stream);
LDS_API int lds_op_transpose(const float* in, float* out, int B, int C, int T, float scale, int to_channels_last, void* stream);
LDS_API int lds_op_div_copy(const float* in, float* out, int64_t n, float divisor, void* stream);

/* ---- HiFi-VAEGAN Generator decode: latent / mel frames -> waveform -------------------------------
 * Replaces Vocoder.infer (diffusion/vocoder.py:31-32) -> Hifi_VAEGAN.forward (encoder/hifi_vaegan/hifi_vaegan.py:52-65) ->
 * Generator.forward (encoder/hifi_vaegan/modules/models.py:248-256).  The configuration mirrors the generator's `h` dictionary
 * (stored in the vocoder checkpoint, hifi_vaegan.py:6-8).  Weights are the Generator's state_dict AFTER remove_weight_norm
 * (keys "conv_pre.weight", "ups.0.weight", "resblocks.3.convs1.1.bias", "conv_post.weight", ...; the Python host folds
 * weight_g / weight_v pairs before loading).  fp32 throughout; same ownership / stream / error conventions as above,
 * messages through lds_vocoder_last_error().  lds_vocode grows its workspace on demand (that call then synchronises). */
#define LDS_VOC_MAX 8
typedef struct {
  int32_t inter_channels;                       /* h["inter_channels"] = Vocoder.dimension (128) */
  int32_t upsample_initial_channel;             /* 512 */
  int32_t n_ups;                                /* len(h["upsample_rates"]) */
  int32_t upsample_rates[LDS_VOC_MAX];          /* product = hop size (512) */
  int32_t upsample_kernel_sizes[LDS_VOC_MAX];
  int32_t resblock_kind;                        /* 1: ResBlock1 (models.py:161-200), 2: ResBlock2 (:203-221) */
  int32_t n_kernels;                            /* len(h["resblock_kernel_sizes"]) */
  int32_t resblock_kernel_sizes[LDS_VOC_MAX];   /* in {3,5,7,11} */
  int32_t resblock_dilations[LDS_VOC_MAX][3];   /* h["resblock_dilation_sizes"][j][:3] (ResBlock2 uses the first two) */
} lds_vocoder_config;
typedef struct lds_vocoder lds_vocoder;
LDS_API const char* lds_vocoder_last_error(void);
LDS_API int lds_vocoder_create(const lds_vocoder_config* cfg, int device, lds_vocoder** out);     /* Generator.__init__ (models.py:225-246) */
LDS_API void lds_vocoder_destroy(lds_vocoder* v);
LDS_API int lds_vocoder_load_weight(lds_vocoder* v, const char* key, const void* data, const int64_t* shape, int ndim, int dtype);
LDS_API int lds_vocoder_finalize(lds_vocoder* v);              /* load_state_dict + remove_weight_norm (hifi_vaegan.py:56-61) */
/* wav[B, T*hop] = tanh(conv_post(...Generator(mel[B,T,inter_channels]^T)))  (hifi_vaegan.py:52-65; the reference returns [B,1,T*hop]) */
LDS_API int lds_vocode(lds_vocoder* v, const float* mel_BTC, int B, int T, float* wav_BL, void* stream);
LDS_API int lds_vocoder_hop(const lds_vocoder* v);
LDS_API int64_t lds_vocoder_launches(const lds_vocoder* v);
LDS_API double lds_vocoder_last_flops(const lds_vocoder* v);   /* algorithmic FLOPs of the last lds_vocode call */
LDS_API int64_t lds_vocoder_workspace_bytes(const lds_vocoder* v);

/* ---- Units front-end: audio -> Whisper encoder units ---------------------------------------------
 * The step before the sampler: Units_Encoder.encode -> WhisperLargeV3.__call__ (tools/tools.py:77-126) = log_mel_spectrogram
 * (encoder/whisper/audio.py:60-80) + AudioEncoder.forward (encoder/whisper/model.py:112-131), then units_forced_alignment
 * (tools/tools.py:193-223) onto the mel frame grid.  The configuration mirrors ModelDimensions' audio fields (model.py:11-21);
 * weights are AudioEncoder's state_dict ("conv1.weight", "blocks.7.attn.query.bias", "blocks.7.mlp.2.weight", "ln_post.weight", ...;
 * fp32, or the fp16 the whisper checkpoints are stored in).  precision: LDS_PREC_FP32 (split-f16 tensor-core GEMMs, fp32-accurate)
 * or LDS_PREC_BF16.  Same ownership / stream / error conventions as above; messages through lds_units_last_error().
 * lds_units_encode grows its workspace on demand (that call then synchronises). */
typedef struct {
  int32_t n_mels;        /* 128 for large-v3 (multiple of 64) */
  int32_t n_state;       /* n_audio_state (1280) */
  int32_t n_head;        /* n_audio_head (20); head dim n_state / n_head in {32, 48, 64} */
  int32_t n_layer;       /* n_audio_layer (32) */
  int32_t precision;     /* lds_precision: LDS_PREC_FP32 or LDS_PREC_BF16 */
} lds_units_config;
typedef struct lds_units lds_units;
LDS_API const char* lds_units_last_error(void);
LDS_API int lds_units_create(const lds_units_config* cfg, int device, lds_units** out);      /* AudioEncoder.__init__ (model.py:113-118) */
LDS_API void lds_units_destroy(lds_units* u);
LDS_API int lds_units_load_weight(lds_units* u, const char* key, const void* data, const int64_t* shape, int ndim, int dtype);
LDS_API int lds_units_finalize(lds_units* u);                 /* model.load_state_dict (tools/tools.py:118) + repacking into operand planes */
/* units[B, T, n_state] = AudioEncoder(mel[B, n_mels, L]), T = lds_units_out_frames(L) = (L - 1) / 2 + 1 (model.py:120-131).
 * pos_TC: device [T, n_state] = sinusoids(T, n_state) (model.py:32-38; computed by the host, it depends on T only). */
LDS_API int lds_units_encode(lds_units* u, const float* mel_BML, int B, int L, const float* pos_TC, float* out_BTC, void* stream);
LDS_API int lds_units_out_frames(int L);
LDS_API int64_t lds_units_launches(const lds_units* u);
LDS_API double lds_units_last_flops(const lds_units* u);      /* algorithmic FLOPs of the last lds_units_encode call */
LDS_API int64_t lds_units_workspace_bytes(const lds_units* u);
/* mel[B, n_mels, L / 160] = log_mel_spectrogram(audio[B, L]) (audio.py:60-80): hann window 400, hop 160, centre / reflect padding,
 * the last STFT frame dropped, |.|^2, filters[n_mels, 201] (audio.py:54-58), log10(max(., 1e-10)), max(., global max - 8), (. + 4) / 4.
 * The DFT is evaluated directly with fp64 accumulation.  L > 200; scratch: one float on the device. */
LDS_API int lds_units_log_mel(const float* audio_BL, int B, int L, const float* filters, int n_mels, float* mel_out, float* scratch,
                      void* stream);
/* out[b * out_per_batch + i, :] = table[b * in_per_batch + idx[i], :] (rows of C floats, C % 4 == 0; idx: device int64 [out_per_batch]).
 * units_forced_alignment 'nearest' / 'left' (tools/tools.py:193-223; idx = the source frame of every output frame, in_per_batch = input
 * frames) and EuclideanCodebook.dequantize / F.embedding (quantize/kmeans_codebook.py:29-31; n_batches = 1, in_per_batch = 0). */
/* idx[m] = argmax_v -(|x_m|^2 - 2 x_m.e_v + |e_v|^2), the first index on ties: EuclideanCodebook.quantize / .encode
 * (quantize/kmeans_codebook.py:15-23,37-46).  x [M, C], embed [V, C] (C % 16 == 0, V % 4 == 0); scratch: round_up(V, 64) + M * V floats
 * on the device; idx: device int64 [M].  fp32 FFMA products (IEEE): the reference's fp32 matmul can only differ on near-ties. */
LDS_API int lds_units_quantize(const float* x_MC, const float* embed_VC, int64_t M, int V, int C, float* scratch, int64_t* idx_out,
                       void* stream);
LDS_API int lds_units_gather_rows(const float* table, const int64_t* idx, int64_t n_batches, int64_t out_per_batch, int64_t in_per_batch,
                          int C, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LDS_B200_H_ */
