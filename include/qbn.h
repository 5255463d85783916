/*
 * libqbn — C ABI of the B200-native stochastic-layer hot path.
 *
 * The reference (martinferianc/quantised-bayesian-nets) has NO FFI layer: its boundary is the
 * Python nn.Module protocol.  This header is therefore the interface a
 * maintainer would bind with ctypes; each entry point cites the reference
 * lines whose arithmetic it replaces.  Paths are relative to the reference repo root.
 *
 * Conventions
 *   - every function returns 0 on success or a negative QBN_ERR_*; qbn_last_error() gives text.
 *     Nothing is thrown across the ABI and no entry point falls back to the CPU.
 *   - all tensor pointers are DEVICE pointers owned by the caller (outputs and workspaces too);
 *     `stream` is a cudaStream_t passed as void*; launches are asynchronous on that stream.
 *   - activations are dense NHWC ("channels last"): [B][H][W][C].  A linear layer is the
 *     degenerate conv H=W=R=S=1 (rows = batch, C = in_features).
 *   - "packed" weights are OHWI: [N][R][S][C] (K = R*S*C contiguous per output channel);
 *     parameters in the nn.Module stay OIHW like the reference (state-dict compatible) and
 *     are packed by qbn_weight_prep.
 *   - RNG: Philox4x32-10, stream = (seed, stream_a = layer/purpose id, stream_b = GLOBAL
 *     Monte-Carlo sample index or training step), counter = element index / 4.  Passing an
 *     explicit noise pointer ("injected noise") bypasses Philox; that is how the parity tests
 *     replay the reference's torch.Generator draws.
 */
#ifndef QBN_H_
#define QBN_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QBN_OK 0
#define QBN_ERR_INVALID_ARG (-1)
#define QBN_ERR_UNSUPPORTED (-2)
#define QBN_ERR_CUDA (-3)
#define QBN_ERR_WORKSPACE (-4)

/* math_mode for the contraction kernels */
#define QBN_MATH_FP32 0 /* CUDA-core FFMA, fp32 exact ordering-insensitive parity (rtol 1e-5) */
#define QBN_MATH_TF32 1 /* tcgen05.mma kind::tf32, fp32 accumulate in TMEM (rtol 1e-3)        */

typedef struct qbn_conv_desc {
  int32_t B, H, W, C;           /* input  [B][H][W][C]                      */
  int32_t N, R, S;              /* filter [N][R][S][C]                      */
  int32_t stride_h, stride_w, pad_h, pad_w, dil_h, dil_w;
  int32_t Ho, Wo;               /* output [B][Ho][Wo][N]                    */
  int32_t out_pad_h, out_pad_w; /* tcgen05 path only: the output tensor is [B][Ho+2*out_pad_h][Wo+2*out_pad_w][N]
                                   and only its interior is written (zero-bordered layout consumed by
                                   qbn_conv_s1_fwd).  pad_h/pad_w may be negative: reading the interior of a
                                   zero-bordered input with a 1x1 stride-2 filter is pad = -border.       */
} qbn_conv_desc;

const char* qbn_last_error(void);
int qbn_version(void);
/* sm_count / compute capability of the current device; fails (QBN_ERR_CUDA) without a GPU. */
int qbn_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---- Philox test hooks (statistical tests; oracle/philox.py is the bit-exact restatement) ---- */
int qbn_philox_u32(uint32_t* out, int64_t n, uint64_t seed, uint32_t stream_a, uint32_t stream_b,
                   void* stream);
int qbn_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t stream_a, uint32_t stream_b,
                      void* stream);
/* out[i] = 1.0f with probability keep_prob else 0 — dropout.py:19-30 (bernoulli_(1-p)) */
int qbn_philox_bernoulli(float* out, int64_t n, float keep_prob, uint64_t seed, uint32_t stream_a,
                         uint32_t stream_b, void* stream);

/* ---- parameter packing ------------------------------------------------------------------- */
/* mu, second: OIHW [N][C][R][S].  second is rho (second_is_sigma=0: sigma=softplus(rho),
 * linear.py:35, conv.py:26) or sigma itself (=1, QAT path where sigma was fake-quantised,
 * conv_qat.py:28,143).  chan_scale[N] (nullable) multiplies mu and sigma per output channel
 * (BN fold, conv.py:70-80 / conv_qat.py:140-143).  Outputs (each nullable) are packed OHWI. */
int qbn_weight_prep(const float* mu, const float* second, int second_is_sigma, int N, int C, int R,
                    int S, const float* chan_scale, float* mu_p, float* sigma_p, float* sigma2_p,
                    int round_tf32 /* round mu_p and sigma2_p to TF32 (LRT operands of the tcgen05 path) */,
                    void* stream);
/* chain rule back to the nn.Module parameters: d_mu[OIHW] = dmu_p ; d_rho = dsig2_p * 2*sigma *
 * sigmoid(rho)  (or d_sigma = dsig2_p * 2*sigma when second_is_sigma) — SURVEY §8a row A3.
 * accumulate!=0 adds into the outputs (used to fold the KL gradient in). */
int qbn_weight_grad_post(const float* dmu_p, const float* dsig2_p, const float* second,
                         int second_is_sigma, int N, int C, int R, int S, float* d_mu,
                         float* d_second, int accumulate, void* stream);

/* ---- A1/A2: local-reparametrisation forward (linear.py:32-40, conv.py:24-32) ---------------
 * out = x*mu + sqrt(1e-8 + x^2*sigma^2) .* eps + bias, both contractions in ONE pass over x.
 * eps: injected noise laid out like out ([B][Ho][Wo][N]) or NULL -> Philox(seed, stream_a,
 * stream_b, offset in out).  std_out (nullable) receives sqrt(1e-8+v) for the backward.       */
int qbn_lrt_fwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p,
                const float* bias, const float* eps, uint64_t seed, uint32_t stream_a,
                uint32_t stream_b, float* out, float* std_out, int math_mode, void* stream);

/* ---- A3: backward of A1/A2 (autograd of linear.py:32-40 / conv.py:24-32, trainer.py:104) ----
 * g = dL/dout; dv = g*eps/(2*std); dmu_p = g^T xcol; dsig2_p = dv^T xcol^2;
 * dx = g*mu + 2x .* (dv*sigma^2); dbias = sum g.  dx/dbias nullable.  eps as in forward.      */
size_t qbn_lrt_bwd_workspace_bytes(const qbn_conv_desc* d);
int qbn_lrt_bwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p,
                const float* grad_out, const float* std_saved, const float* eps, uint64_t seed,
                uint32_t stream_a, uint32_t stream_b, float* dx, float* dmu_p, float* dsig2_p,
                float* dbias, void* workspace, size_t workspace_bytes, int math_mode,
                void* stream);

/* ---- A4: eval-time weight sampling (linear.py:42-50, conv.py:33-39) ------------------------
 * w[s][i] = mu_p[i] + sigma_p[i]*eps, i in [0,n): one draw per (sample, layer, weight).
 * eps: injected [n_samples][n] (same packed order) or NULL -> Philox(seed, layer_id,
 * sample0+s, i).  The buffer is a few MB and is consumed straight out of L2 by qbn_conv_fwd.  */
int qbn_sample_weights(const float* mu_p, const float* sigma_p, int64_t n, int n_samples,
                       const float* eps, uint64_t seed, uint32_t layer_id, uint32_t sample0,
                       float* w, int round_tf32 /* store W rounded to TF32 (RNA) for the tcgen05 path */,
                       void* stream);

/* y = conv(x, w[s]) for every Monte-Carlo sample s in one launch, with the caller-side glue of
 * models_bbb.py:170-183,226-245 folded into the epilogue: per-channel affine (eval BatchNorm or
 * bias), residual add (src/utils.py:49-55), ReLU.  x: [n_samples][B].. or, if x_shared, [B]..
 * read by every sample.  in_mask (nullable) is the A8 MC-Dropout mask [n_samples*B][C] applied to
 * the operand load with multiplier in_mult (dropout.py:15-40).  out: [n_samples][B][Ho][Wo][N]. */
#define QBN_FLAG_RELU 1            /* ReLU in the epilogue                                      */
#define QBN_FLAG_A_TF32_READY 2    /* TF32 mode: x is already TF32-exact (e.g. written by a previous
                                      launch with OUT_ROUND_TF32) -> operand goes global->smem by
                                      cp.async; otherwise it is rounded (cvt.rna) in registers   */
#define QBN_FLAG_OUT_ROUND_TF32 4  /* TF32 mode: round the stored activations to TF32 (RNA) so the
                                      next layer can take the cp.async path                      */
#define QBN_FLAG_OUT_PHASE_SPLIT 8 /* qbn_conv_p4_fwd: write the output phase-split for a stride-2 consumer */
#define QBN_FLAG_X_SHARED_STACKED 32 /* qbn_conv_p4_fwd: x holds ONE set of B maps read by all n_samples samples (first layer);
                                      w is ONE blocked tensor whose N rows are the samples' weights stacked (row s*N + n), so
                                      the input is staged once and one accumulator tile holds every sample (n_samples*N <= 256) */
#define QBN_FLAG_RELU_PRE 64        /* qbn_conv_p4_fwd: ReLU right after the affine, before the output mask / residual */
#define QBN_FLAG_OUT_P4 16         /* qbn_conv_fwd (TF32): out and residual are planar-C4 with out_pad zero rows on top /
                                      zero columns on the left of every map and a zero tail (see below)          */
int qbn_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const float* x,
                 const float* w, int w_shared, const float* scale, const float* shift,
                 const float* residual, int flags, const float* in_mask, float in_mult, float* out,
                 int math_mode, void* stream);

/* Stride-1 "same" convolution (odd RxS, pad (R-1)/2,(S-1)/2) on the zero-bordered layout, TF32
 * tcgen05, persistent: x [n_samples][B][Hp][Wp][C] with Hp = H+R-1, Wp = W+S-1 and a zero border,
 * TF32-exact values; w [n_samples][N][R][S][C] (qbn_sample_weights with round_tf32); out and
 * residual [n_samples][B][Hp][Wp][N] in the same layout (the border of out is written as zeros).
 * Same epilogue/flags as qbn_conv_fwd.  One smem tile serves all R*S taps (zero-copy im2col). */
int qbn_conv_s1_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, const float* x,
                    const float* w, int w_shared, const float* scale, const float* shift,
                    const float* residual, int flags, float* out, void* stream);

/* A5 for every Bayesian layer of a model in ONE launch: jobs_dev = device array of qbn_kl_job; kl_out (zeroed by the caller)
 * receives the sum over layers (models_bbb.py:254-259), d_mu / d_rho (nullable) the gradients times grad_scale (overwritten). */
typedef struct qbn_kl_job { const float* mu; const float* rho; float* d_mu; float* d_rho; int64_t n; float sigma_prior; int32_t pad_; } qbn_kl_job;
int qbn_kl_multi(const void* jobs_dev, int n_jobs, int64_t max_n, float* kl_out, float grad_scale, void* stream);
/* trainer.py:105-107 (NaN gradients -> 0, per parameter) for every gradient tensor of a model in ONE launch per 128 tensors;
 * jobs_host: HOST array (it travels in the kernel parameters: no device table, capture-safe) */
typedef struct qbn_scrub_job { float* grad; int64_t n; } qbn_scrub_job;
int qbn_scrub_nan_multi(const qbn_scrub_job* jobs_host, int n_jobs, void* stream);

/* ---- planar-C4 path: the S-batched eval convolution (A4 + A11 glue) with NO operand handling by threads.
 * Activations "planar C4": [C/4 chunk planes][plane_rows][4 floats]; a plane holds the pixels of the zero-bordered maps
 * [n_samples*B][Hp][Wp] followed by a zero tail.  The zeros are SHARED between neighbours: ph zero rows on TOP of every map
 * (= the bottom padding of the map above), pw zero columns on the LEFT of every row (= the right padding of the row above),
 * Hp = H + ph, Wp = W + pw, tail = ph*Wp + pw pixels after the last map; *_plane_rows = rows per chunk plane of that tensor.
 * Values TF32-exact; the caller zero-initialises buffers once (kernels write border pixels as zeros, never the tail).  Sampled weights
 * "blocked": [sample][C/CB][R*S][CB/4][n_pad][4] (CB depends on C and the stride, n_pad = N rounded up to 16: qbn_p4_weight_floats), written by
 * qbn_sample_weights_blocked from mu/sigma blocked once by qbn_p4_block_weights.  A tile's operands are then
 * contiguous runs moved by bulk copies, every filter tap is a row-shifted UMMA descriptor on one smem image
 * (zero-copy im2col), and the epilogue's 16-byte stores are contiguous across a warp.
 *   stride 1: odd RxS "same" conv, Hp = H + (R-1)/2.
 *   stride 2: 3x3 pad 1 or 1x1 pad 0; Hp = H_out + 1; x is PHASE-SPLIT: [C/4][4 phases x n_samples*B*Hp*Wp (+ tail)][4]
 *             where phase (a,b) holds pixels (2i+a, 2j+b) of the full-resolution map at (i+1, j+1), as written
 *             by the producing qbn_conv_p4_fwd with QBN_FLAG_OUT_PHASE_SPLIT (into a pre-zeroed buffer).
 * out / residual: planar C4 [N/4][out_plane_rows][4] with the output geometry; border pixels of out are written as zeros. */
int qbn_p4_weight_floats(int C, int N, int R, int S, int stride, long long* out_floats /* host */);
int qbn_p4_block_weights(const float* w_ohwi /* [n_mats][N][taps][C] */, int n_mats, int N, int C, int taps,
                         int stride, int cb_override /* 0: layout rule */, float* out, void* stream);
int qbn_sample_weights_blocked(const float* mu_b, const float* sigma_b, int N, int C, int taps, int stride, int n_samples,
                               const float* eps /* nullable, canonical [n_samples][N][taps][C] */, uint64_t seed,
                               uint32_t layer_id, uint32_t sample0, float* w, int round_tf32, void* stream);
/* the same for several layers in ONE launch: jobs_dev = device array of qbn_p4_sample_job (blockIdx.z = job) */
typedef struct qbn_p4_sample_job {
  const float* mu_b; const float* sigma_b; const float* eps /* nullable */; float* w;
  int32_t N, C, taps, stride; uint32_t layer_id;
  int32_t n_stack;   /* > 0: write the n_stack samples stacked along N (row s*N + n of ONE blocked tensor, QBN_FLAG_X_SHARED_STACKED) */
  const float* chan_scale; /* nullable [N]: W[n][.] *= chan_scale[n] before rounding (BatchNorm scale folded into the weights) */
  int32_t cb_override;     /* > 0: channels per block (qbn_p4_shortcut_block_channels) instead of the layout rule */
  int32_t w_sample_stride4;/* > 0: float4 units between consecutive samples in w (several jobs filling one tensor) */
  int32_t s_off;           /* stacked jobs: the job covers the chunk's samples [s_off, s_off + n_stack) */
  int32_t pad_;
} qbn_p4_sample_job;
int qbn_sample_weights_blocked_multi(const void* jobs_dev, int n_jobs, int64_t max_floats_per_sample, int n_samples,
                                     uint64_t seed, uint32_t sample0, int round_tf32, void* stream);
/* epilogue order: affine -> [RELU_PRE] -> [out_mask: MC-Dropout of the OUTPUT, x*mask[img][n]*mult, A8] -> [+residual]
 * -> [RELU] -> [RNA]; that is conv-BN-ReLU-dropout (models_mc.py:125-129) and conv-BN-dropout-add-ReLU (:130-157) */
int qbn_conv_p4_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, int stride, const float* x,
                    long long x_plane_rows, const float* w, int w_shared, const float* scale, const float* shift,
                    const float* residual, long long res_plane_rows, const float* out_mask /* nullable [n_samples*B][N] */,
                    float out_mask_mult, int flags, float* out, long long out_plane_rows, void* stream);
/* Stride-1 conv with the BasicBlock's 1x1 stride-2 shortcut FUSED as extra K blocks (models_bbb.py:163-178):
 *   out = act( conv_RxS(x, W) + conv_1x1,stride2(x_block, Wsc) + shift ),  W and Wsc carrying their BatchNorm scales
 * (qbn_p4_sample_job.chan_scale), so both branches accumulate in ONE TMEM tile: no shortcut launch, no residual round trip.
 * x2 = the block input, phase-split (its phase (0,0) IS the stride-2 sampling grid), C2 channels; every sample's weight tensor
 * is [main blocks][shortcut blocks of qbn_p4_shortcut_block_channels(C, C2) channels, one tap]. */
int qbn_p4_shortcut_block_channels(int C, int C2);
int qbn_conv_p4_shortcut_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, const float* x,
                             long long x_plane_rows, const float* w, const float* x2, long long x2_plane_rows, int C2,
                             const float* scale, const float* shift, int flags, float* out, long long out_plane_rows,
                             void* stream);
/* Bernoulli(keep) masks of several dropout sites in ONE launch: jobs_dev = device array of qbn_mask_job; the mask of
 * site j, sample s is out[j][s][elems], Philox(seed, site_id, sample0 + s, element) */
typedef struct qbn_mask_job { float* out; int64_t elems; uint32_t site_id; int32_t pad_; } qbn_mask_job;
int qbn_dropout_masks_multi(const void* jobs_dev, int n_jobs, int64_t max_elems, int n_samples, float keep_prob,
                            uint64_t seed, uint32_t sample0, void* stream);
/* global average pool of planar-C4 maps -> [n_img][C] (divisor = interior pixel count) */
int qbn_avgpool_p4(const float* x, int64_t n_img, int HW /* Hp*Wp */, int64_t plane_rows, int C, float divisor, float* out, void* stream);

/* ---- A8 standalone: x[b,h,w,c] * mask[b,c] * mult (dropout.py:35-39); mask NULL -> Philox ---- */
int qbn_dropout_fwd(const float* x, int64_t rows /*B*/, int64_t hw, int64_t C, const float* mask,
                    float keep_prob, float mult, uint64_t seed, uint32_t stream_a,
                    uint32_t stream_b, float* out, float* mask_out, void* stream);

/* ---- A5: KL(q||p) closed form + gradient (utils_bbb.py:3-5, linear.py:24-28, conv.py:43-47) --
 * kl_out (device scalar) += 0.5*sum(2*log(sp/sigma) - 1 + (sigma/sp)^2 + (mu/sp)^2).
 * If d_mu/d_rho are non-NULL they are ACCUMULATED with grad_scale * dKL/d(.).                */
int qbn_kl_fwd_bwd(const float* mu, const float* rho, int64_t n, float sigma_prior, float* kl_out,
                   float* d_mu, float* d_rho, float grad_scale, void* stream);
/* utils_bbb.py:3-5 with the reference's own arguments: sigma (not rho), scalar mu_prior and sigma_prior.
 * kl_out += 0.5*sum(2*log(sp/sigma) - 1 + (sigma/sp)^2 + ((mu_prior-mu)/sp)^2); d_mu / d_sigma accumulated if non-NULL. */
int qbn_kl_sigma_fwd_bwd(const float* mu, const float* sigma, int64_t n, float mu_prior, float sigma_prior,
                         float* kl_out, float* d_mu, float* d_sigma, float grad_scale, void* stream);

/* ---- A7: fake quantisation + MovingAverageMinMax observer (linear_qat.py:18-41,
 * conv_qat.py:26-52,139-170; torch/ao/quantization/observer.py:374-410,668-683) -------------
 * state = {min, max, initialised} on device.  observe!=0: one pass computes (min,max) of x and
 * updates the EMA (c = averaging_const; first call initialises), then scale/zero_point.
 * y = (clamp(rint(x*(1/scale)) + zp, qmin, qmax) - zp) * scale; mask (nullable) = STE pass mask. */
int qbn_fake_quant_fwd(const float* x, int64_t n, float* state, float averaging_const, int observe,
                       int qmin, int qmax, float* scale, int32_t* zero_point, float* y,
                       uint8_t* mask, void* workspace /* >= 2*sizeof(float)*1024 */, void* stream);
int qbn_fake_quant_bwd(const float* grad_y, const uint8_t* mask, int64_t n, float* grad_x,
                       void* stream);

/* ---- A6: true int8 path (linear_q.py:80-94,154-173; conv_q.py:107-125,189-209) ---------------
 * quantize: q = clamp(rint(x*(1/scale)) + zp, qmin, qmax)  (torch.quantize_per_tensor)         */
int qbn_quantize_u8(const float* x, int64_t n, float scale, int32_t zp, int qmin, int qmax,
                    uint8_t* q, void* stream);
int qbn_quantize_s8(const float* x, int64_t n, float scale, int32_t zp, int qmin, int qmax,
                    int8_t* q, void* stream);
int qbn_dequantize_u8(const uint8_t* q, int64_t n, float scale, int32_t zp, float* x, void* stream);

typedef struct qbn_i8_sample_params {
  float s_mu;    int32_t z_mu;    /* qint8 mu tensor qparams                                  */
  float s_sigma; int32_t z_sigma; /* qint8 sigma tensor qparams                               */
  float s_eps;   int32_t z_eps;   /* NOISE_SCALE=3/127, NOISE_ZERO_POINT=0 (quantized/__init__.py) */
  float s_mul;   int32_t z_mul;   /* mul_noise QFunctional output qparams                     */
  float s_add;   int32_t z_add;   /* add_weight QFunctional output qparams                    */
  int32_t w_min, w_max;           /* clamp_weight INT_BOUNDS (src/utils.py:18-20,32-37)       */
  int64_t n_vec;                  /* elements [0,n_vec) use ATen's vector-body dequantise
                                     (fma(scale, q, -zp*scale)), the rest its scalar-tail form
                                     ((q-zp)*scale); ATen's split is n_vec = (n/64)*64; <0 = that */
} qbn_i8_sample_params;
/* w[s][i] = clamp_weight(qadd(mu_q, qmul(sigma_q, quantize(eps)))) — SURVEY §8a A6 steps 1-4.
 * eps injected fp32 [n_samples][n] or NULL -> Philox.  All arithmetic reproduces ATen's
 * QuantizedCPU kernels bit for bit (fp32 multipliers, round-half-even).                       */
int qbn_i8_sample_weights(const int8_t* mu_q, const int8_t* sigma_q, int64_t n, int n_samples,
                          const qbn_i8_sample_params* p, const float* eps, uint64_t seed,
                          uint32_t layer_id, uint32_t sample0, int8_t* w, void* stream);

/* u8 x s8 -> s32 contraction with FBGEMM's requantisation (A6 step 6):
 * acc = sum (x-z_x)(w-z_w); y = clamp(rint((fp32(acc) + bias/(s_x*s_w)) * (s_x*s_w/s_out)) + z_out,
 * lo, hi), lo = max(z_out if relu else 0, act_min), hi = min(255, act_max) (clamp_activation,
 * src/utils.py:25-30).  [act_min, act_max] is the MODEL's activation range (activation_precision): the
 * caller guarantees that x lies in it too (the reference clamps after every module, models_bbb.py:125-127);
 * with act_max <= 127 QBN_I8_AUTO picks the tcgen05 kernel, which feeds (x - z_x) as s8.  Pass 255 when
 * the input range is unknown.  acc_dump (nullable, int32 like out) exposes raw accumulators to tests.  */
int qbn_i8_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const uint8_t* x,
                    float s_x, int32_t z_x, const int8_t* w, int w_shared, float s_w, int32_t z_w,
                    const float* bias, float s_out, int32_t z_out, int relu, int act_min,
                    int act_max, uint8_t* out, int32_t* acc_dump, int path, void* stream);
#define QBN_I8_AUTO 0  /* tcgen05 kind::i8 when C % 8 == 0 and activations are <= 7 bit, else IMAD */
#define QBN_I8_IMAD 1  /* CUDA-core integer FMA kernel (any shape)                               */
#define QBN_I8_UMMA 2  /* force the tcgen05 kernel (QBN_ERR_UNSUPPORTED if the shape is ragged)   */
/* quantized::add (+optional relu) — src/utils.py:49-55 via QFunctional.add; A6 step 3 formula
 * on quint8: f = (a-za)*sa + (b-zb)*sb ; q = clamp(rint(f*(1/so)) + zo, lo, hi)              */
int qbn_i8_add(const uint8_t* a, float sa, int32_t za, const uint8_t* b, float sb, int32_t zb,
               int64_t n, int64_t n_vec, float so, int32_t zo, int lo, int hi, uint8_t* out,
               void* stream);
/* torch.relu on quint8 (BasicBlock.end, models_bbb.py:186) fused with clamp_activation (src/utils.py:25-30):
 * out = clamp(max(x, z_x), lo, hi); qparams unchanged.                                          */
int qbn_i8_relu(const uint8_t* x, int64_t n, int32_t z_x, int lo, int hi, uint8_t* out, void* stream);
/* nn.AvgPool2d(k) on a quint8 map (models_bbb.py:211; ATen qavg_pool2d, output qparams = input's), x NHWC
 * [B][H][W][C] -> out [B][H/k][W/k][C]: q = clamp(rint(fp32(sum - k*k*z) * fp32(1/(k*k))) + z, 0, 255), then
 * clamp_activation to [lo, hi].                                                                 */
int qbn_i8_avgpool(const uint8_t* x, int64_t B, int H, int W, int C, int k, int32_t z_x, int lo, int hi,
                   uint8_t* out, void* stream);
/* int8 MC-Dropout (dropout.py:31-39): mask quantised at (s_m,z_m) then quantized::mul with the
 * output at the same (s_m,z_m); mul_scalar only rescales.  mask fp32 {0,1} [rows][C] injected
 * or NULL -> Philox.                                                                         */
int qbn_i8_dropout(const uint8_t* x, float s_x, int32_t z_x, int64_t rows, int64_t hw, int64_t C,
                   const float* mask, float keep_prob, float s_m, int32_t z_m, uint64_t seed,
                   uint32_t stream_a, uint32_t stream_b, int lo, int hi, uint8_t* out,
                   void* stream);
/* the same for a chunk of Monte-Carlo samples in one launch: x [n_samples][rows_per_sample][hw][C], sample s draws its
 * mask from Philox(seed, site, sample0 + s) — equal to n_samples calls of qbn_i8_dropout with stream_b = sample0 + s */
int qbn_i8_dropout_mc(const uint8_t* x, float s_x, int32_t z_x, int n_samples, int64_t rows_per_sample, int64_t hw,
                      int64_t C, float keep_prob, float s_m, int32_t z_m, uint64_t seed, uint32_t site,
                      uint32_t sample0, int lo, int hi, uint8_t* out, void* stream);

/* ---- A6 on the planar zero-copy kernel ("planar C16", csrc/p4_layout.cuh) ---------------------------------------------
 * Activations: int8 maps holding (q - zero_point) — quint8 activations of at most 7 bits (quant_utils.py:120) — laid out
 * [C_pad/16 chunk planes][n_img*(H+1)*(W+1) pixels + tail][16], one shared zero row on top / zero column on the left of every
 * map (the zero IS the zero-point padding of a quint8 convolution), channel count zero-padded to a multiple of 32.  A tensor
 * feeding a stride-2 conv is stored phase-split like its float twin.  Sampled weights: blocked
 * [sample][C_pad/CB][tap][CB/16][n_pad][16] with row N = ones over the real channels (accumulator column N = sum_k (x - z_x),
 * the z_w correction), n_pad = ceil16(N + 1).
 *
 * qbn_i8_conv_p16_fwd = conv_q.py:107-125,189-209 step 6 for a chunk of Monte-Carlo samples (x_shared: all samples read the
 * same input maps) + the glue that follows in models_bbb.py:170-183: clamp_activation, quantized::add with `residual`
 * (add_relu: the block's final ReLU) — all in the epilogue.  q = clamp(rint((fp32(acc) + bias/(s_x*s_w)) * (s_x*s_w/s_out))
 * + z_out, relu ? z_out : 0, act_max); with a residual r: clamp(rint((s_out*(q - z_out) + s_res*(r - z_res)) / s_add) + z_add,
 * add_relu ? z_add : 0, act_max) in ATen's vector-body arithmetic (maps whose element count is a multiple of 64, as every
 * ResNet map is).  The output map holds q - z_out (or q - z_add).  acc_dump (nullable): [n_samples*B*Hp*Wp][N] accumulators. */
typedef struct qbn_i8_requant {
  float s_x, s_w; int32_t z_w; float s_out; int32_t z_out; int32_t relu; int32_t act_max;
  float s_res; int32_t z_res; float s_add; int32_t z_add; int32_t add_relu;
} qbn_i8_requant;
int qbn_i8_conv_p16_fwd(int n_samples, int B, int Hp, int Wp, int C_pad, int N, int R, int S, int stride,
                        const int8_t* x, long long x_plane_rows, int x_shared, const int8_t* w_blocked, int w_shared,
                        const float* bias, const qbn_i8_requant* rq, const int8_t* residual, long long res_plane_rows,
                        int flags /* QBN_FLAG_OUT_PHASE_SPLIT */, int8_t* out, long long out_plane_rows,
                        int32_t* acc_dump, void* stream);
int qbn_p16_weight_bytes(int C_pad, int N, int R, int S, int stride, long long* out_bytes);
/* sampled weights [n_samples][N][C][taps] (the sampler's OIHW order) -> blocked operands of qbn_i8_conv_p16_fwd */
int qbn_i8_p16_block_weights(const int8_t* w_oihw, int n_samples, int N, int C, int C_pad, int taps, int stride,
                             int8_t* out, void* stream);
/* entry / exit of the layout: quint8 NHWC [n_img][H][W][C] <-> planar C16 (border 1), and the global average pool
 * (nn.AvgPool2d(H), ATen qavg_pool2d rounding) to quint8 [n_img][C] clamped to [lo, hi] */
int qbn_i8_p16_from_nhwc(const uint8_t* x, int64_t n_img, int H, int W, int C, int C_pad, int32_t z_x,
                         int64_t plane_rows, int8_t* out, void* stream);
int qbn_i8_p16_to_nhwc(const int8_t* x, int64_t n_img, int H, int W, int C, int32_t z_x, int64_t plane_rows,
                       uint8_t* out, void* stream);
int qbn_i8_p16_avgpool(const int8_t* x, int64_t n_img, int H, int W, int C, int32_t z_x, int64_t plane_rows,
                       int lo, int hi, uint8_t* out, void* stream);

/* ---- A1-A3 on the planar zero-copy kernels: LRT training in TF32 mode (linear.py:32-40, conv.py:24-32 and their autograd,
 * trainer.py:104).  The module boundary stays dense NHWC; per layer the operands are staged once into planar-C4 zero-bordered
 * maps (phase-split for a stride-2 layer) and every contraction — forward (mean and variance side by side in TMEM), input
 * gradient (the same kernel on flipped / transposed weights; four phase launches for a stride-2 layer), weight gradients
 * (pixels as the reduction dimension, both operands MN-major from a swizzled 32-channel-block copy of the maps) — runs on tcgen05
 * without a gather.
 * Eligible: C_pad % 8 == 0, N % 8 == 0, N <= 256, stride 1 with an odd 'same' filter or stride 2 with 3x3 pad 1 / 1x1 pad 0.
 * Every plane must hold ceil(rows / 128) * 128 + 128 + 2 * (Wp + 1) + 8 rows (rows = phases * n_img * Hp * Wp).            */
/* eps[i] = element i of the Philox stream (seed, stream_a, stream_b + draw offset): the noise the LRT kernels draw in their epilogue
 * when eps is NULL, materialised (the planar training path generates it at full occupancy and keeps it for the backward) */
int qbn_lrt_noise(float* out, int64_t n, uint64_t seed, uint32_t stream_a, uint32_t stream_b, void* stream);
/* x NHWC [n_img][H][W][C] -> x_p4 = tf32(x), xsq_p4 = tf32(x * x) (nullable), planes [C_pad/4][plane_rows][4], border (bh, bw) */
int qbn_p4_stage_input(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                       long long plane_rows, float* x_p4, float* xsq_p4, void* stream);
/* g_out, std_saved, eps (NULL -> the forward's Philox draw) NHWC [n_img][H][W][N] -> g_p4 = tf32(g), dv_p4 = tf32(g * eps / (2 std)) */
int qbn_p4_stage_grad(const float* g_out, const float* std_saved, const float* eps, uint64_t seed, uint32_t stream_a,
                      uint32_t stream_b, int64_t n_img, int H, int W, int N, int bh, int bw, long long plane_rows, float* g_p4,
                      float* dv_p4, void* stream);
/* the two staging passes with the W32 copies (qbn_w32_from_p4 below) written in the same pass: what ops.LRTFunction uses */
int qbn_lrt_stage_input(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                        long long plane_rows, float* x_p4, float* xsq_p4, float* x_w32, float* xsq_w32, void* stream);
int qbn_lrt_stage_grad(const float* g_out, const float* std_saved, const float* eps, uint64_t seed, uint32_t stream_a,
                       uint32_t stream_b, int64_t n_img, int H, int W, int N, int bh, int bw, long long plane_rows, float* g_p4,
                       float* dv_p4, float* g_w32, float* dv_w32, void* stream);
/* qbn_lrt_stage_input and qbn_lrt_noise in ONE launch: noise_out[0 .. noise_n) = the layer's eps tensor (conv.py:29-30: one draw per
 * output element; noise_n = B * Ho * Wo * N, a multiple of 4), the same values qbn_lrt_noise writes for (seed, stream_a, stream_b) */
int qbn_lrt_stage_input_noise(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                              long long plane_rows, float* x_p4, float* xsq_p4, float* x_w32, float* xsq_w32, float* noise_out,
                              int64_t noise_n, uint64_t seed, uint32_t stream_a, uint32_t stream_b, void* stream);
/* OIHW (mu, rho | sigma) -> blocked [mu | sigma^2] operand, TF32-rounded.  mode 0: forward (input channels zero-padded to C_pad,
 * blocked for `stride`); 1: input gradient of a stride-1 layer (taps reversed, channels swapped); 2: one phase of a stride-2
 * layer's input gradient (parameter taps tap_list[0..n_taps)); 3: the four phases of a 3x3 stride-2 layer, four taps each (absent
 * taps zeroed), one [mu | sigma^2] pair per phase.  out: 2 * qbn_p4_weight_floats(...) floats per pair (total in *out_floats). */
int qbn_lrt_p4_weight_prep(const float* mu, const float* second, int second_is_sigma, int N, int C, int C_pad, int R, int S,
                           int stride, int mode, const int* tap_list, int n_taps, float* out, long long* out_floats, void* stream);
/* out = conv(x, mu) + sqrt(1e-8 + conv(x_sq, sigma^2)) .* eps + bias ; std_out = the square root.  Hp, Wp: padded extent of the
 * OUTPUT maps.  out / std_out / eps: dense NHWC.  eps NULL -> Philox(seed, stream_a, stream_b, offset in out / 4) like qbn_lrt_fwd. */
int qbn_lrt_conv_p4_fwd(int B, int Hp, int Wp, int C_pad, int N, int R, int S, int stride, const float* x_p4, const float* xsq_p4,
                        long long x_plane_rows, const float* w_blocked, const float* bias, const float* eps, uint64_t seed,
                        uint32_t stream_a, uint32_t stream_b, float* out, float* std_out, void* stream);
/* dx = convT(g, mu) + 2 x .* convT(dv, sigma^2) of a stride-1 layer.  C: channels of g, N: channels of dx; xin, dx dense NHWC */
int qbn_lrt_conv_p4_dgrad(int B, int Hp, int Wp, int C, int N, int R, int S, const float* g_p4, const float* dv_p4,
                          long long g_plane_rows, const float* w_flipped_blocked, const float* xin, float* dx, void* stream);
/* all four phases of a 3x3 stride-2 layer's input gradient in one launch (weights: qbn_lrt_p4_weight_prep mode 3) */
int qbn_lrt_conv_p4_dgrad_s2(int B, int Hp, int Wp, int C, int N, const float* g_p4, const float* dv_p4, long long g_plane_rows,
                             const float* w_phases_blocked, const float* xin, float* dx, void* stream);
/* phase (a, b) of a stride-2 layer's input gradient: dx[2i+a][2j+b] = sum_t g[i + di_t][j + dj_t] * w[tap_t], shifts[t] =
 * di_t * Wp + dj_t (di, dj in {0, 1}); xin, dx dense NHWC [B][2(Hp-1)][2(Wp-1)][N] */
int qbn_lrt_conv_p4_dgrad_phase(int B, int Hp, int Wp, int C, int N, int n_taps, const int* shifts, int phase_a, int phase_b,
                                const float* g_p4, const float* dv_p4, long long g_plane_rows, const float* w_phase_blocked,
                                const float* xin, float* dx, void* stream);
/* planar-C4 maps [C_pad/4][plane_rows][4] (as staged above, borders and tail included) -> "W32": [ceil(C_pad/32)][plane_rows][32]
 * floats, the four 32-byte chunks of row r XORed with r & 3 — the global image of the one shared-memory layout in which
 * tcgen05.mma kind::tf32 accepts MN-major operands (SWIZZLE_128B_BASE32B).  src1 / dst1 nullable: a second tensor, same launch. */
int qbn_w32_from_p4(const float* src0, const float* src1, int C_pad, long long plane_rows, float* dst0, float* dst1, void* stream);
/* dmu_p / dsig2_p [N][R*S][C_real] (packed OHWI, overwritten) = sum over pixels of g (x) x and dv (x) x^2, operands in W32 */
int qbn_lrt_wgrad_p4(int B, int Hp, int Wp, int C_real, int N, int R, int S, int stride, const float* g_w32,
                     const float* dv_w32, long long g_plane_rows, const float* x_w32, const float* xsq_w32, long long x_plane_rows,
                     float* dmu_p, float* dsig2_p, void* stream);

/* Programmatic dependent launch of the planar convolution kernels (qbn_conv_p4_fwd, qbn_conv_p4_shortcut_fwd, qbn_i8_conv_p16_fwd,
 * qbn_lrt_conv_p4_*): with enabled != 0 every such launch carries cudaLaunchAttributeProgrammaticStreamSerialization, so its
 * prologue (barrier init, TMEM allocation, operand tables) overlaps the tail of the previous kernel of the stream; the kernel
 * touches global memory only after that kernel has completed (griddepcontrol.wait).  Process-wide, off by default. */
int qbn_set_pdl(int enabled);

/* Unit window of the sample-sharded evaluation (qbn_b200.dist.shard_units; SURVEY 8e-i: MC samples partitioned over the GPUs):
 * a rank's share is a contiguous range of (sample, image) units, so of the n_samples samples of a chunk the FIRST may only need
 * images [first_img, B) and the LAST only [0, end_img).  While a window is set, the qbn_conv_p4_fwd / qbn_conv_p4_shortcut_fwd /
 * qbn_i8_conv_p16_fwd launches over exactly n_samples samples cover the tiles of those rows only (rows outside keep their previous
 * contents; the caller masks them: qbn_softmax_accumulate_window).  end_img 0 = the whole batch; n_samples 0 switches the window
 * off.  Launches over another sample count (a fixed-weight layer on the shared input), launches with the samples stacked along N
 * (shared-input first layer) and the LRT kinds ignore it.  Process-wide and sticky like qbn_set_pdl. */
int qbn_p4_set_window(int first_img, int end_img, int n_samples);

/* int8 MC-Dropout (dropout.py:31-39) of a chunk of Monte-Carlo samples on planar-C16 maps, optionally followed by the BasicBlock's
 * quantized::add[_relu] with `residual` (models_mc.py:143-157: the dropout sits between the second conv and the add).  Same integers
 * as qbn_i8_dropout_mc + qbn_i8_add.  x holds q - z_x at scale s_x (x_shared: B images shared by all samples); mask fp32 {0,1}
 * [n_samples*B][C] (qbn_dropout_masks_multi); out holds q - z_m at scale s_drop_out = s_m * multiplier, or q - z_add after the add
 * (rq: s_res, z_res, s_add, z_add, add_relu).  x, residual: normal layout (Hp x Wp maps); out: normal or phase-split. */
int qbn_i8_p16_dropout(const int8_t* x, long long x_plane_rows, int x_shared, int n_samples, int B, int Hp, int Wp,
                       int out_phase_split, int C, float s_x, const float* mask, float s_m, int32_t z_m, float s_drop_out, int act_max,
                       const int8_t* residual, long long res_plane_rows, const qbn_i8_requant* rq, int8_t* out,
                       long long out_plane_rows, void* stream);

/* Draw offset of the Monte-Carlo samplers, kept on the DEVICE: after qbn_set_sample_base(p) every sampler launch
 * (qbn_sample_weights*, qbn_i8_sample_weights, qbn_dropout_masks_multi, qbn_i8_dropout_mc) uses the Philox stream index
 * *p + sample0 + s instead of sample0 + s, reading *p when the kernel runs.  One captured CUDA graph then serves every batch
 * with fresh noise (the reference redraws per batch, experiments/utils.py:342-347): bump the scalar between replays.
 * NULL switches it off.  The pointer must stay valid while launches that captured it can still run. */
int qbn_set_sample_base(const uint32_t* base_dev);

/* ---- A9: Monte-Carlo aggregation (experiments/utils.py:344-355) ----------------------------
 * logits [n_samples][B][K] -> psum[B][K] (+)= sum_s softmax(logits_s)  (models_bbb.py:131,243)
 * accumulate==0 overwrites.  The caller divides by the GLOBAL S after the allreduce.           */
int qbn_softmax_accumulate(const float* logits, int n_samples, int B, int K, float* psum,
                           int accumulate, void* stream);
/* the same with the unit window of the sample-sharded evaluation (qbn_b200.dist.shard_units): sample 0 contributes images
 * [first_img, B) only, sample n_samples - 1 images [0, end_img) only */
int qbn_softmax_accumulate_window(const float* logits, int n_samples, int B, int K, int first_img, int end_img, float* psum,
                                  int accumulate, void* stream);
/* probs [n_samples][B][K] -> mean over samples (stack(...).mean(dim=1)) */
int qbn_mc_mean(const float* probs, int n_samples, int64_t BK, float* mean, void* stream);
/* regression (experiments/utils.py:349-353): mean_s mu, Var_s(mu) (unbiased) + mean_s var */
int qbn_reg_mc_reduce(const float* mu, const float* var, int n_samples, int64_t B, float* mean_out,
                      float* var_out, void* stream);

/* ---- A10: metric reductions (src/metrics.py:20-29,48-57,76-85,104-112,381-383) -------------
 * out[0..3] += {sum[argmax!=t], sum -log(p_t+1e-8), sum (p-onehot)^2, sum -p log(p+1e-8)};
 * out[4 + 3*b + {0,1,2}] += {sum conf, sum correct, count} of confidence bin b (n_bins equal-
 * width bins, ECE l1).  scale multiplies probs first (1/S after an allreduce of sums).        */
int qbn_cls_metrics(const float* probs, const int64_t* target, int B, int K, float scale,
                    int n_bins, float* out, void* stream);
/* out[0..2] += {gaussian nll sum, squared error sum, abs error sum} (metrics.py:135-157,176-225) */
int qbn_reg_metrics(const float* mean, const float* var, const float* target, int64_t B,
                    float* out, void* stream);

/* ---- N4: the classification ELBO (src/losses.py:14-29, trainer.py:96-104), value and gradient in one launch ----
 * out3 = {loss, data, kl_term}: data = data_scale * mean_b -log(probs[b][target[b]] + 1e-8), kl_term = *kl * kl_scale,
 * loss = data + gamma * kl_term; d_probs [B][K] (nullable) = d data / d probs.  'batch' scaling: data_scale 1, kl_scale
 * 1 / (B * n_batches); 'whole': data_scale n_points * loss_multiplier, kl_scale 1 / n_batches.                       */
int qbn_elbo_cls(const float* probs, const int64_t* target, const float* kl, int B, int K, float data_scale,
                 float kl_scale, float gamma, float* out3, float* d_probs, void* stream);

/* ---- SGHMC / SGLD parameter update (utils_sgld.py:30-92), one fused pass per parameter tensor:
 * grad += weight_decay * p (in place, like the reference); burn_in: tau, g, V_hat preconditioner update; resample_momentum:
 * v = z_m * sqrt(lr^2 / (sqrt(V_hat) + eps)); v += -lr^2/(sqrt(V_hat)+eps) * grad - base_C * v + z_n * sqrt(max(2 lr^2/(sqrt(V_hat)+eps)
 * base_C - lr^4, 1e-16)); NaN / inf momentum -> 0; p += v.  z_momentum / z_noise: injected standard normals or NULL -> Philox
 * (seed, stream_a, stream_b [+1 for the noise], element).                                                                   */
int qbn_sghmc_step(float* p, float* grad, float* tau, float* g, float* V_hat, float* v_momentum, int64_t n,
                   float weight_decay, float lr, float base_C, float eps, int burn_in, int resample_momentum,
                   const float* z_momentum, const float* z_noise, uint64_t seed, uint32_t stream_a,
                   uint32_t stream_b, void* stream);

/* ---- A11 glue kernels that survive fusion only at resolution changes ------------------------ */
int qbn_maxpool2x2(const float* x, int64_t B, int H, int W, int C, float* out, void* stream);
/* global average pool HxW -> 1 (nn.AvgPool2d(4) on the 4x4 map, models_bbb.py:209) */
int qbn_avgpool_all(const float* x, int64_t B, int HW, int C, float divisor /* <=0: HW */, float* out, void* stream);
/* NCHW <-> NHWC (entry/exit of the NHWC domain) */
int qbn_nchw_to_nhwc(const float* x, int64_t B, int C, int HW, float* out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* QBN_H_ */
