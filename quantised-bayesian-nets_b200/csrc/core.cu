// HBM-bound kernels of the stochastic-layer path: Philox hooks, parameter packing, eval-time
// weight sampling (float A4 and int8 A6 steps 1-4), MC-Dropout (A8), KL (A5), fake-quant +
// observer (A7), quantise/requantise glue (A6/A11).  All are grid-stride kernels with 128-bit
// accesses where the layout allows, sized in multiples of the SM count.
#include <stdarg.h>
#include <string.h>
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// error / device plumbing
// ---------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void qbn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

// Device-side draw offset of the Monte-Carlo samplers (qbn_set_sample_base): every Philox stream index `sample0 + s` becomes
// `*base + sample0 + s`, read by the kernel at run time.  A CUDA graph captured once therefore serves every batch: the host
// bumps the device scalar between replays instead of re-capturing with a new sample0 (experiments/utils.py:342-347 redraws
// the noise for every batch).  One process drives one GPU (torch.distributed, one rank per device), hence a process-wide pointer.
static const uint32_t* g_sample_base = nullptr;
extern "C" int qbn_set_sample_base(const uint32_t* base_dev) {
  g_sample_base = base_dev;
  return QBN_OK;
}
const uint32_t* qbn_sample_base_ptr() { return g_sample_base; }

int qbn_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

extern "C" const char* qbn_last_error(void) { return g_err; }
extern "C" int qbn_version(void) { return 100; }

extern "C" int qbn_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  int dev = 0;
  QBN_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  QBN_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// Philox test hooks
// ---------------------------------------------------------------------------------------------
__global__ void philox_u32_kernel(uint32_t* __restrict__ out, int64_t n, uint64_t seed, uint32_t sa, uint32_t sb) {
  int64_t n4 = (n + 3) >> 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n4; c += (int64_t)gridDim.x * blockDim.x) {
    Philox4 p = philox_at(seed, sa, sb, (uint64_t)c);
    uint32_t v[4] = {p.x, p.y, p.z, p.w};
    int64_t base = c << 2;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j < n) out[base + j] = v[j];
  }
}

__global__ void philox_normal_kernel(float* __restrict__ out, int64_t n, uint64_t seed, uint32_t sa, uint32_t sb, const uint32_t* __restrict__ sbase) {
  if (sbase) sb += *sbase;           // device-side draw offset (qbn_set_sample_base)
  int64_t n4 = (n + 3) >> 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n4; c += (int64_t)gridDim.x * blockDim.x) {
    float z[4];
    philox_normal4(seed, sa, sb, (uint64_t)c, z);
    int64_t base = c << 2;
    if (base + 3 < n && ((reinterpret_cast<uintptr_t>(out + base) & 15) == 0)) {
      *reinterpret_cast<float4*>(out + base) = make_float4(z[0], z[1], z[2], z[3]);
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j < n) out[base + j] = z[j];
    }
  }
}

__global__ void philox_bernoulli_kernel(float* __restrict__ out, int64_t n, float keep, uint64_t seed, uint32_t sa, uint32_t sb) {
  int64_t n4 = (n + 3) >> 2;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n4; c += (int64_t)gridDim.x * blockDim.x) {
    Philox4 p = philox_at(seed, sa, sb, (uint64_t)c);
    uint32_t v[4] = {p.x, p.y, p.z, p.w};
    int64_t base = c << 2;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (base + j < n) out[base + j] = u01(v[j]) < keep ? 1.0f : 0.0f;
  }
}

extern "C" int qbn_philox_u32(uint32_t* out, int64_t n, uint64_t seed, uint32_t sa, uint32_t sb, void* stream) {
  QBN_CHECK_ARG(out && n >= 0, "out/n");
  if (n == 0) return QBN_OK;
  philox_u32_kernel<<<qbn_grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, sa, sb);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
extern "C" int qbn_philox_normal(float* out, int64_t n, uint64_t seed, uint32_t sa, uint32_t sb, void* stream) {
  QBN_CHECK_ARG(out && n >= 0, "out/n");
  if (n == 0) return QBN_OK;
  philox_normal_kernel<<<qbn_grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, sa, sb, nullptr);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
// The LRT noise of one layer and forward (linear.py:37-38, conv.py:29-30) as a tensor: eps[i] = element i of the Philox stream
// (seed, stream_a, stream_b + *sample_base) — exactly the draw the LRT kernels make in their epilogue when eps is NULL.  At the
// small per-launch sizes of a training step the conv epilogue (8 warps per SM) is latency-bound on the Philox / Box-Muller chain;
// this kernel runs at full occupancy, and the backward reads the tensor instead of regenerating it.
extern "C" int qbn_lrt_noise(float* out, int64_t n, uint64_t seed, uint32_t sa, uint32_t sb, void* stream) {
  QBN_CHECK_ARG(out && n >= 0, "out/n");
  if (n == 0) return QBN_OK;
  philox_normal_kernel<<<qbn_grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, seed, sa, sb, qbn_sample_base_ptr());
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
extern "C" int qbn_philox_bernoulli(float* out, int64_t n, float keep_prob, uint64_t seed, uint32_t sa, uint32_t sb, void* stream) {
  QBN_CHECK_ARG(out && n >= 0, "out/n");
  QBN_CHECK_ARG(keep_prob >= 0.f && keep_prob <= 1.f, "keep_prob in [0,1]");
  if (n == 0) return QBN_OK;
  philox_bernoulli_kernel<<<qbn_grid_for((n + 3) / 4, 256), 256, 0, (cudaStream_t)stream>>>(out, n, keep_prob, seed, sa, sb);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// parameter packing OIHW -> OHWI (+softplus, +BN channel scale) and the chain rule back
// ---------------------------------------------------------------------------------------------
__global__ void weight_prep_kernel(const float* __restrict__ mu, const float* __restrict__ second, int second_is_sigma,
                                   int N, int C, int R, int S, const float* __restrict__ chan_scale,
                                   float* __restrict__ mu_p, float* __restrict__ sigma_p, float* __restrict__ sigma2_p,
                                   int round_tf32) {
  int64_t total = (int64_t)N * C * R * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes the packed OHWI order
    int c = (int)(i % C);
    int64_t t = i / C;
    int s = (int)(t % S);
    t /= S;
    int r = (int)(t % R);
    int n = (int)(t / R);
    int64_t src = (((int64_t)n * C + c) * R + r) * S + s;
    float cs = chan_scale ? chan_scale[n] : 1.0f;
    float m = mu[src];
    float sg = second_is_sigma ? second[src] : softplus_f(second[src]);
    if (chan_scale) {
      m = __fmul_rn(m, cs);
      sg = __fmul_rn(sg, cs);
    }
    float s2 = __fmul_rn(sg, sg);
    if (round_tf32) { m = tf32_round(m); s2 = tf32_round(s2); }
    if (mu_p) mu_p[i] = m;
    if (sigma_p) sigma_p[i] = sg;
    if (sigma2_p) sigma2_p[i] = s2;
  }
}

extern "C" int qbn_weight_prep(const float* mu, const float* second, int second_is_sigma, int N, int C, int R, int S,
                               const float* chan_scale, float* mu_p, float* sigma_p, float* sigma2_p, int round_tf32,
                               void* stream) {
  QBN_CHECK_ARG(mu && second, "mu/second");
  QBN_CHECK_ARG(N > 0 && C > 0 && R > 0 && S > 0, "N,C,R,S > 0");
  int64_t total = (int64_t)N * C * R * S;
  weight_prep_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(mu, second, second_is_sigma, N, C, R, S,
                                                                               chan_scale, mu_p, sigma_p, sigma2_p, round_tf32);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void weight_grad_post_kernel(const float* __restrict__ dmu_p, const float* __restrict__ dsig2_p,
                                        const float* __restrict__ second, int second_is_sigma, int N, int C, int R, int S,
                                        float* __restrict__ d_mu, float* __restrict__ d_second, int accumulate) {
  int64_t total = (int64_t)N * C * R * S;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    // i indexes OIHW (coalesced parameter-gradient writes); gather from packed OHWI
    int s = (int)(i % S);
    int64_t t = i / S;
    int r = (int)(t % R);
    t /= R;
    int c = (int)(t % C);
    int n = (int)(t / C);
    int64_t src = (((int64_t)n * R + r) * S + s) * C + c;
    if (d_mu && dmu_p) {
      float v = dmu_p[src];
      d_mu[i] = accumulate ? d_mu[i] + v : v;
    }
    if (d_second && dsig2_p) {
      float sec = second[i];
      float sg = second_is_sigma ? sec : softplus_f(sec);
      float v = dsig2_p[src] * 2.0f * sg;
      if (!second_is_sigma) v *= sigmoid_f(sec);
      d_second[i] = accumulate ? d_second[i] + v : v;
    }
  }
}

extern "C" int qbn_weight_grad_post(const float* dmu_p, const float* dsig2_p, const float* second, int second_is_sigma,
                                    int N, int C, int R, int S, float* d_mu, float* d_second, int accumulate, void* stream) {
  QBN_CHECK_ARG(N > 0 && C > 0 && R > 0 && S > 0, "N,C,R,S > 0");
  QBN_CHECK_ARG(!(d_second && dsig2_p) || second, "second required for d_second");
  int64_t total = (int64_t)N * C * R * S;
  weight_grad_post_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(dmu_p, dsig2_p, second, second_is_sigma, N, C,
                                                                                    R, S, d_mu, d_second, accumulate);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A4: eval-time weight sampling.  grid = (chunks of 4 weights, samples)
// ---------------------------------------------------------------------------------------------
__global__ void sample_weights_kernel(const float* __restrict__ mu_p, const float* __restrict__ sigma_p, int64_t n,
                                      const float* __restrict__ eps, uint64_t seed, uint32_t layer_id, uint32_t sample0,
                                      float* __restrict__ w, int round_tf32, const uint32_t* __restrict__ sbase) {
  if (sbase) sample0 += *sbase;      // per-call draw offset read on the device: one CUDA graph serves every batch

  int s = blockIdx.y;
  int64_t n4 = (n + 3) >> 2;
  const bool vec_ok = (n & 3) == 0;
  float* ws = w + (int64_t)s * n;
  const float* es = eps ? eps + (int64_t)s * n : nullptr;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n4; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t base = c << 2;
    float z[4];
    if (es) {
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = (base + j < n) ? es[base + j] : 0.f;
    } else {
      philox_normal4(seed, layer_id, sample0 + (uint32_t)s, (uint64_t)c, z);
    }
    if (vec_ok) {
      float4 m = *reinterpret_cast<const float4*>(mu_p + base);
      float4 sg = *reinterpret_cast<const float4*>(sigma_p + base);
      // linear.py:46-47: std = mul(noise, softplus) ; weight = add(weight, std) -> two roundings
      float4 o;
      o.x = __fadd_rn(m.x, __fmul_rn(z[0], sg.x));
      o.y = __fadd_rn(m.y, __fmul_rn(z[1], sg.y));
      o.z = __fadd_rn(m.z, __fmul_rn(z[2], sg.z));
      o.w = __fadd_rn(m.w, __fmul_rn(z[3], sg.w));
      if (round_tf32) { o.x = tf32_round(o.x); o.y = tf32_round(o.y); o.z = tf32_round(o.z); o.w = tf32_round(o.w); }
      *reinterpret_cast<float4*>(ws + base) = o;
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (base + j < n) {
          float o = __fadd_rn(mu_p[base + j], __fmul_rn(z[j], sigma_p[base + j]));
          ws[base + j] = round_tf32 ? tf32_round(o) : o;
        }
    }
  }
}

extern "C" int qbn_sample_weights(const float* mu_p, const float* sigma_p, int64_t n, int n_samples, const float* eps,
                                  uint64_t seed, uint32_t layer_id, uint32_t sample0, float* w, int round_tf32, void* stream) {
  QBN_CHECK_ARG(mu_p && sigma_p && w, "null pointer");
  QBN_CHECK_ARG(n > 0 && n_samples > 0 && n_samples <= 65535, "n>0, 0<n_samples<=65535");
  int64_t n4 = (n + 3) / 4;
  int gx = (int)((n4 + 255) / 256);
  int cap = (qbn_sm_count() * 8 + n_samples - 1) / n_samples;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  sample_weights_kernel<<<dim3(gx, n_samples), 256, 0, (cudaStream_t)stream>>>(mu_p, sigma_p, n, eps, seed, layer_id, sample0, w, round_tf32, g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A8: MC-Dropout, mask per (row, channel), broadcast over hw
// ---------------------------------------------------------------------------------------------
__global__ void dropout_mask_kernel(float* __restrict__ mask_out, int64_t n, float keep, uint64_t seed, uint32_t sa, uint32_t sb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    mask_out[i] = philox_uniform1(seed, sa, sb, (uint64_t)i) < keep ? 1.0f : 0.0f;
}

__global__ void dropout_apply_kernel(const float* __restrict__ x, int64_t rows, int64_t hw, int64_t C,
                                     const float* __restrict__ mask, float mult, float* __restrict__ out) {
  int64_t total = rows * hw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = i % C;
    int64_t b = i / (hw * C);
    // dropout.py:38-39: x = mul(x, mask) ; x = mul_scalar(x, multiplier)
    out[i] = __fmul_rn(__fmul_rn(x[i], mask[b * C + c]), mult);
  }
}

// C % 4 == 0: one float4 per thread, the mask row is indexed by the image of the pixel (no 64-bit division per element)
__global__ void dropout_apply4_kernel(const float4* __restrict__ x, int64_t rows, int hw, int C4, const float4* __restrict__ mask, float mult,
                                      float4* __restrict__ out) {
  const int64_t n_pix = rows * hw;
  for (int64_t pix = blockIdx.x * (int64_t)blockDim.y + threadIdx.y; pix < n_pix; pix += (int64_t)gridDim.x * blockDim.y) {
    const int64_t b = pix / hw;
    for (int c = threadIdx.x; c < C4; c += blockDim.x) {
      const float4 v = x[pix * C4 + c], m = mask[b * C4 + c];
      // dropout.py:38-39: x = mul(x, mask) ; x = mul_scalar(x, multiplier)
      out[pix * C4 + c] = make_float4(__fmul_rn(__fmul_rn(v.x, m.x), mult), __fmul_rn(__fmul_rn(v.y, m.y), mult),
                                      __fmul_rn(__fmul_rn(v.z, m.z), mult), __fmul_rn(__fmul_rn(v.w, m.w), mult));
    }
  }
}

extern "C" int qbn_dropout_fwd(const float* x, int64_t rows, int64_t hw, int64_t C, const float* mask, float keep_prob,
                               float mult, uint64_t seed, uint32_t sa, uint32_t sb, float* out, float* mask_out, void* stream) {
  QBN_CHECK_ARG(x && out, "x/out");
  QBN_CHECK_ARG(rows > 0 && hw > 0 && C > 0, "rows,hw,C > 0");
  QBN_CHECK_ARG(mask || mask_out, "either an injected mask or a mask_out buffer for the Philox mask");
  const float* m = mask;
  if (!mask) {
    dropout_mask_kernel<<<qbn_grid_for(rows * C, 256), 256, 0, (cudaStream_t)stream>>>(mask_out, rows * C, keep_prob, seed, sa, sb);
    QBN_CHECK_LAUNCH();
    m = mask_out;
  }
  if (C % 4 == 0 && hw < (1 << 30) && ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(m)) & 15) == 0) {
    const int C4 = (int)(C / 4);
    int tx = 1;
    while (tx < C4 && tx < 32) tx <<= 1;                 // lanes over channel chunks, the rest of the block over pixels
    const dim3 block(tx, 256 / tx);
    const int64_t n_pix = rows * hw;
    int64_t gx = (n_pix + block.y - 1) / block.y;
    const int64_t cap = (int64_t)qbn_sm_count() * 16;
    if (gx > cap) gx = cap;
    dropout_apply4_kernel<<<(unsigned)gx, block, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), rows, (int)hw, C4,
                                                                            reinterpret_cast<const float4*>(m), mult,
                                                                            reinterpret_cast<float4*>(out));
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
  dropout_apply_kernel<<<qbn_grid_for(rows * hw * C, 256), 256, 0, (cudaStream_t)stream>>>(x, rows, hw, C, m, mult, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A5: KL + gradient, one pass
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float kl_term(float m, float r, float sp, float inv_sp, float inv_sp2, float gscale, float& gm, float& gr) {
  const float sg = softplus_f(r);
  const float a = sg * inv_sp, b = m * inv_sp;
  gm = gscale * m * inv_sp2;
  gr = gscale * (sg * inv_sp2 - 1.0f / sg) * sigmoid_f(r);
  return 2.0f * logf(sp / sg) - 1.0f + a * a + b * b;
}
__global__ void kl_kernel(const float* __restrict__ mu, const float* __restrict__ rho, int64_t n, float sp,
                          float* __restrict__ kl_out, float* __restrict__ d_mu, float* __restrict__ d_rho, float gscale) {
  double acc = 0.0;
  const float inv_sp = 1.0f / sp;
  const float inv_sp2 = inv_sp * inv_sp;
  const bool vec = ((reinterpret_cast<uintptr_t>(mu) | reinterpret_cast<uintptr_t>(rho) | reinterpret_cast<uintptr_t>(d_mu) |
                     reinterpret_cast<uintptr_t>(d_rho)) & 15) == 0;
  const int64_t n4 = vec ? (n >> 2) : 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 m = reinterpret_cast<const float4*>(mu)[i], r = reinterpret_cast<const float4*>(rho)[i];
    float4 gm, gr;
    float t = kl_term(m.x, r.x, sp, inv_sp, inv_sp2, gscale, gm.x, gr.x);      // per-thread partial in fp32 over 4 terms, then fp64
    t += kl_term(m.y, r.y, sp, inv_sp, inv_sp2, gscale, gm.y, gr.y);
    t += kl_term(m.z, r.z, sp, inv_sp, inv_sp2, gscale, gm.z, gr.z);
    t += kl_term(m.w, r.w, sp, inv_sp, inv_sp2, gscale, gm.w, gr.w);
    acc += (double)t;
    if (d_mu) {
      float4 o = reinterpret_cast<float4*>(d_mu)[i];
      reinterpret_cast<float4*>(d_mu)[i] = make_float4(o.x + gm.x, o.y + gm.y, o.z + gm.z, o.w + gm.w);
    }
    if (d_rho) {
      float4 o = reinterpret_cast<float4*>(d_rho)[i];
      reinterpret_cast<float4*>(d_rho)[i] = make_float4(o.x + gr.x, o.y + gr.y, o.z + gr.z, o.w + gr.w);
    }
  }
  for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float gm, gr;
    acc += (double)kl_term(mu[i], rho[i], sp, inv_sp, inv_sp2, gscale, gm, gr);
    if (d_mu) d_mu[i] += gm;
    if (d_rho) d_rho[i] += gr;
  }
  // block reduction in double, one atomic per block
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(kl_out, (float)(0.5 * acc));
  }
}

extern "C" int qbn_kl_fwd_bwd(const float* mu, const float* rho, int64_t n, float sigma_prior, float* kl_out, float* d_mu,
                              float* d_rho, float grad_scale, void* stream) {
  QBN_CHECK_ARG(mu && rho && kl_out, "null pointer");
  QBN_CHECK_ARG(n > 0 && sigma_prior > 0.f, "n>0, sigma_prior>0");
  kl_kernel<<<qbn_grid_for((n + 3) / 4, 256, 8), 256, 0, (cudaStream_t)stream>>>(mu, rho, n, sigma_prior, kl_out, d_mu, d_rho, grad_scale);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// utils_bbb.py:3-5 with its own argument list: sigma given (not rho), scalar mu_prior / sigma_prior.  The layers call the
// rho form above; this one serves callers that follow the reference's signature literally.
__global__ void kl_sigma_kernel(const float* __restrict__ mu, const float* __restrict__ sigma, int64_t n, float mp, float sp,
                                float* __restrict__ kl_out, float* __restrict__ d_mu, float* __restrict__ d_sigma, float gscale) {
  double acc = 0.0;
  const float inv_sp = 1.0f / sp, inv_sp2 = inv_sp * inv_sp;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float sg = sigma[i], dm = mp - mu[i];
    const float a = sg * inv_sp, b = dm * inv_sp;
    acc += (double)(2.0f * logf(sp / sg) - 1.0f + a * a + b * b);
    if (d_mu) d_mu[i] += -gscale * dm * inv_sp2;
    if (d_sigma) d_sigma[i] += gscale * (sg * inv_sp2 - 1.0f / sg);
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) atomicAdd(kl_out, (float)(0.5 * acc));
  }
}
extern "C" int qbn_kl_sigma_fwd_bwd(const float* mu, const float* sigma, int64_t n, float mu_prior, float sigma_prior, float* kl_out,
                                    float* d_mu, float* d_sigma, float grad_scale, void* stream) {
  QBN_CHECK_ARG(mu && sigma && kl_out, "null pointer");
  QBN_CHECK_ARG(n > 0 && sigma_prior > 0.f, "n>0, sigma_prior>0");
  kl_sigma_kernel<<<qbn_grid_for(n, 256, 8), 256, 0, (cudaStream_t)stream>>>(mu, sigma, n, mu_prior, sigma_prior, kl_out, d_mu, d_sigma,
                                                                            grad_scale);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A7: fused observer (min/max + EMA + qparams) and fake-quantise
//   workspace: [0] uint32 ticket counter (zero on entry, reset on exit), [16..] float2 partials
// ---------------------------------------------------------------------------------------------
#define FQ_MAX_BLOCKS 1024

__global__ void fq_observe_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ state, float c, int qmin, int qmax,
                                  float* __restrict__ scale, int32_t* __restrict__ zp, uint32_t* __restrict__ ticket,
                                  float2* __restrict__ partial) {
  float mn = INFINITY, mx = -INFINITY;
  int64_t n4 = n >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  const bool al = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (al) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = x4[i];
      mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
      mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
    }
    for (int64_t i = (n4 << 2) + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      mn = fminf(mn, x[i]);
      mx = fmaxf(mx, x[i]);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      mn = fminf(mn, x[i]);
      mx = fmaxf(mx, x[i]);
    }
  }
  __shared__ float smn[32], smx[32];
  __shared__ bool is_last;
  mn = warp_min(mn);
  mx = warp_max(mx);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    mn = lane < (blockDim.x >> 5) ? smn[lane] : INFINITY;
    mx = lane < (blockDim.x >> 5) ? smx[lane] : -INFINITY;
    mn = warp_min(mn);
    mx = warp_max(mx);
    if (lane == 0) {
      partial[blockIdx.x] = make_float2(mn, mx);
      __threadfence();
      uint32_t t = atomicAdd(ticket, 1u);
      is_last = (t == gridDim.x - 1);
    }
  }
  __syncthreads();
  if (!is_last) return;
  // last block: fold partials, EMA (observer.py:668-683), qparams (observer.py:374-410)
  __threadfence();
  mn = INFINITY; mx = -INFINITY;
  for (int i = threadIdx.x; i < (int)gridDim.x; i += blockDim.x) {
    float2 p = __ldcg(&partial[i]);
    mn = fminf(mn, p.x);
    mx = fmaxf(mx, p.y);
  }
  mn = warp_min(mn);
  mx = warp_max(mx);
  __syncthreads();
  if (lane == 0) { smn[wid] = mn; smx[wid] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = fminf(mn, smn[w]); mx = fmaxf(mx, smx[w]); }
    float omin = state[0], omax = state[1];
    if (state[2] == 0.0f) { omin = mn; omax = mx; }
    else {
      omin = __fadd_rn(omin, __fmul_rn(c, __fsub_rn(mn, omin)));
      omax = __fadd_rn(omax, __fmul_rn(c, __fsub_rn(mx, omax)));
    }
    state[0] = omin; state[1] = omax; state[2] = 1.0f;
    float mneg = fminf(omin, 0.0f), mpos = fmaxf(omax, 0.0f);
    float sc = __fdiv_rn(__fsub_rn(mpos, mneg), (float)(qmax - qmin));
    sc = fmaxf(sc, 1.1920928955078125e-07f);
    int z = qmin - (int)rintf(__fdiv_rn(mneg, sc));
    z = max(qmin, min(qmax, z));
    *scale = sc;
    *zp = z;
    *ticket = 0u;
  }
}

__global__ void fq_quant_kernel(const float* __restrict__ x, int64_t n, const float* __restrict__ scale,
                                const int32_t* __restrict__ zp, int qmin, int qmax, float* __restrict__ y, uint8_t* __restrict__ mask) {
  const float sc = *scale;
  const float inv = __fdiv_rn(1.0f, sc);
  const float z = (float)(*zp);
  const float lo = (float)qmin, hi = (float)qmax;
  int64_t n4 = n >> 2;
  const bool al = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 &&
                  (!mask || (reinterpret_cast<uintptr_t>(mask) & 3) == 0);
  if (al) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 v = reinterpret_cast<const float4*>(x)[i];
      float q[4] = {rintf(__fmul_rn(v.x, inv)) + z, rintf(__fmul_rn(v.y, inv)) + z, rintf(__fmul_rn(v.z, inv)) + z,
                    rintf(__fmul_rn(v.w, inv)) + z};
      float o[4];
      uint32_t mk = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        bool in = q[j] >= lo && q[j] <= hi;
        mk |= (in ? 1u : 0u) << (8 * j);
        o[j] = __fmul_rn(fminf(fmaxf(q[j], lo), hi) - z, sc);
      }
      reinterpret_cast<float4*>(y)[i] = make_float4(o[0], o[1], o[2], o[3]);
      if (mask) reinterpret_cast<uint32_t*>(mask)[i] = mk;
    }
  }
  int64_t start = al ? (n4 << 2) : 0;
  for (int64_t i = start + blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float q = rintf(__fmul_rn(x[i], inv)) + z;
    bool in = q >= lo && q <= hi;
    y[i] = __fmul_rn(fminf(fmaxf(q, lo), hi) - z, sc);
    if (mask) mask[i] = in ? 1 : 0;
  }
}

extern "C" int qbn_fake_quant_fwd(const float* x, int64_t n, float* state, float averaging_const, int observe, int qmin, int qmax,
                                  float* scale, int32_t* zero_point, float* y, uint8_t* mask, void* workspace, void* stream) {
  QBN_CHECK_ARG(x && scale && zero_point, "null pointer");
  QBN_CHECK_ARG(n > 0 && qmin < qmax, "n>0, qmin<qmax");
  cudaStream_t st = (cudaStream_t)stream;
  if (observe) {
    QBN_CHECK_ARG(state && workspace, "observe needs state and workspace");
    int blocks = qbn_grid_for((n + 3) / 4, 256, 4);
    if (blocks > FQ_MAX_BLOCKS) blocks = FQ_MAX_BLOCKS;
    uint32_t* ticket = reinterpret_cast<uint32_t*>(workspace);
    float2* partial = reinterpret_cast<float2*>(reinterpret_cast<char*>(workspace) + 16);
    fq_observe_kernel<<<blocks, 256, 0, st>>>(x, n, state, averaging_const, qmin, qmax, scale, zero_point, ticket, partial);
    QBN_CHECK_LAUNCH();
  }
  if (y) {
    fq_quant_kernel<<<qbn_grid_for((n + 3) / 4, 256), 256, 0, st>>>(x, n, scale, zero_point, qmin, qmax, y, mask);
    QBN_CHECK_LAUNCH();
  }
  return QBN_OK;
}

__global__ void fq_bwd_kernel(const float* __restrict__ g, const uint8_t* __restrict__ mask, int64_t n, float* __restrict__ gx) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    gx[i] = mask[i] ? g[i] : 0.0f;
}

extern "C" int qbn_fake_quant_bwd(const float* grad_y, const uint8_t* mask, int64_t n, float* grad_x, void* stream) {
  QBN_CHECK_ARG(grad_y && mask && grad_x && n > 0, "null pointer / n");
  fq_bwd_kernel<<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(grad_y, mask, n, grad_x);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A6 glue: quantise / dequantise, int8 weight sampling, quantised add, int8 dropout
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void quantize_kernel(const float* __restrict__ x, int64_t n, float inv, int zp, int qmin, int qmax, T* __restrict__ q) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int v = (int)rintf(__fmul_rn(x[i], inv)) + zp;
    q[i] = (T)max(qmin, min(qmax, v));
  }
}

extern "C" int qbn_quantize_u8(const float* x, int64_t n, float scale, int32_t zp, int qmin, int qmax, uint8_t* q, void* stream) {
  QBN_CHECK_ARG(x && q && n > 0 && scale > 0.f, "x/q/n/scale");
  QBN_CHECK_ARG(qmin >= 0 && qmax <= 255 && qmin <= qmax, "0<=qmin<=qmax<=255");
  quantize_kernel<uint8_t><<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, 1.0f / scale, zp, qmin, qmax, q);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
extern "C" int qbn_quantize_s8(const float* x, int64_t n, float scale, int32_t zp, int qmin, int qmax, int8_t* q, void* stream) {
  QBN_CHECK_ARG(x && q && n > 0 && scale > 0.f, "x/q/n/scale");
  QBN_CHECK_ARG(qmin >= -128 && qmax <= 127 && qmin <= qmax, "-128<=qmin<=qmax<=127");
  quantize_kernel<int8_t><<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, 1.0f / scale, zp, qmin, qmax, q);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void dequantize_u8_kernel(const uint8_t* __restrict__ q, int64_t n, float scale, int zp, float* __restrict__ x) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    x[i] = __fmul_rn((float)((int)q[i] - zp), scale);
}
extern "C" int qbn_dequantize_u8(const uint8_t* q, int64_t n, float scale, int32_t zp, float* x, void* stream) {
  QBN_CHECK_ARG(x && q && n > 0, "x/q/n");
  dequantize_u8_kernel<<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(q, n, scale, zp, x);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

struct I8SampleConsts {
  float inv_eps;   // 1.0f / s_eps
  float mul_mult;  // s_sigma * s_eps * (1.0f / s_mul)   (ATen qmul)
  float s_mu, premul_mu;    // vector-body dequantise: fma(s, q, -z*s)
  float s_mul, premul_mul;
  float inv_add;   // 1.0f / s_add
  int z_mu, z_sigma, z_eps, z_mul, z_add, w_min, w_max;
  int64_t n_vec;
};

QBN_DEVINL int clampi(int v, int lo, int hi) { return max(lo, min(hi, v)); }

__global__ void i8_sample_weights_kernel(const int8_t* __restrict__ mu_q, const int8_t* __restrict__ sigma_q, int64_t n,
                                         I8SampleConsts k, const float* __restrict__ eps, uint64_t seed, uint32_t layer_id,
                                         uint32_t sample0, int8_t* __restrict__ w, const uint32_t* __restrict__ sbase) {
  if (sbase) sample0 += *sbase;      // per-call draw offset read on the device: one CUDA graph serves every batch

  int s = blockIdx.y;
  int64_t n4 = (n + 3) >> 2;
  int8_t* ws = w + (int64_t)s * n;
  const float* es = eps ? eps + (int64_t)s * n : nullptr;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < n4; c += (int64_t)gridDim.x * blockDim.x) {
    int64_t base = c << 2;
    float z[4];
    if (es) {
#pragma unroll
      for (int j = 0; j < 4; ++j) z[j] = (base + j < n) ? es[base + j] : 0.f;
    } else {
      philox_normal4(seed, layer_id, sample0 + (uint32_t)s, (uint64_t)c, z);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int64_t i = base + j;
      if (i >= n) break;
      // (1) quantize_per_tensor(eps, NOISE_SCALE, 0, qint8)
      int eq = clampi((int)rintf(__fmul_rn(z[j], k.inv_eps)) + k.z_eps, -128, 127);
      // (2) quantized::mul(sigma_q, eps_q)
      int prod = ((int)sigma_q[i] - k.z_sigma) * (eq - k.z_eps);
      int r = clampi((int)rintf(__fmul_rn((float)prod, k.mul_mult)) + k.z_mul, -128, 127);
      // (3) quantized::add(mu_q, r)
      float da, db;
      if (i < k.n_vec) {
        da = __fmaf_rn(k.s_mu, (float)mu_q[i], k.premul_mu);
        db = __fmaf_rn(k.s_mul, (float)r, k.premul_mul);
      } else {
        da = __fmul_rn((float)((int)mu_q[i] - k.z_mu), k.s_mu);
        db = __fmul_rn((float)(r - k.z_mul), k.s_mul);
      }
      int wq = clampi((int)rintf(__fmul_rn(__fadd_rn(da, db), k.inv_add)) + k.z_add, -128, 127);
      // (4) clamp_weight
      ws[i] = (int8_t)clampi(wq, k.w_min, k.w_max);
    }
  }
}

extern "C" int qbn_i8_sample_weights(const int8_t* mu_q, const int8_t* sigma_q, int64_t n, int n_samples,
                                     const qbn_i8_sample_params* p, const float* eps, uint64_t seed, uint32_t layer_id,
                                     uint32_t sample0, int8_t* w, void* stream) {
  QBN_CHECK_ARG(mu_q && sigma_q && p && w, "null pointer");
  QBN_CHECK_ARG(n > 0 && n_samples > 0 && n_samples <= 65535, "n>0, 0<n_samples<=65535");
  QBN_CHECK_ARG(p->s_mu > 0 && p->s_sigma > 0 && p->s_eps > 0 && p->s_mul > 0 && p->s_add > 0, "scales must be > 0");
  I8SampleConsts k;
  k.inv_eps = 1.0f / p->s_eps;
  k.mul_mult = p->s_sigma * p->s_eps * (1.0f / p->s_mul);
  k.s_mu = p->s_mu;
  k.premul_mu = p->s_mu * (float)(-p->z_mu);
  k.s_mul = p->s_mul;
  k.premul_mul = p->s_mul * (float)(-p->z_mul);
  k.inv_add = 1.0f / p->s_add;
  k.z_mu = p->z_mu; k.z_sigma = p->z_sigma; k.z_eps = p->z_eps; k.z_mul = p->z_mul; k.z_add = p->z_add;
  k.w_min = p->w_min; k.w_max = p->w_max;
  k.n_vec = p->n_vec < 0 ? (n / 64) * 64 : p->n_vec;
  int64_t n4 = (n + 3) / 4;
  int gx = (int)((n4 + 255) / 256);
  int cap = (qbn_sm_count() * 8 + n_samples - 1) / n_samples;
  if (cap < 1) cap = 1;
  if (gx > cap) gx = cap;
  i8_sample_weights_kernel<<<dim3(gx, n_samples), 256, 0, (cudaStream_t)stream>>>(mu_q, sigma_q, n, k, eps, seed, layer_id, sample0, w, eps ? nullptr : g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void i8_add_kernel(const uint8_t* __restrict__ a, float sa, float pa, int za, const uint8_t* __restrict__ b, float sb,
                              float pb, int zb, int64_t n, int64_t n_vec, float inv_so, int zo, int lo, int hi,
                              uint8_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float da, db;
    if (i < n_vec) {
      da = __fmaf_rn(sa, (float)a[i], pa);
      db = __fmaf_rn(sb, (float)b[i], pb);
    } else {
      da = __fmul_rn((float)((int)a[i] - za), sa);
      db = __fmul_rn((float)((int)b[i] - zb), sb);
    }
    int q = (int)rintf(__fmul_rn(__fadd_rn(da, db), inv_so)) + zo;
    out[i] = (uint8_t)clampi(q, lo, hi);
  }
}

extern "C" int qbn_i8_add(const uint8_t* a, float sa, int32_t za, const uint8_t* b, float sb, int32_t zb, int64_t n,
                          int64_t n_vec, float so, int32_t zo, int lo, int hi, uint8_t* out, void* stream) {
  QBN_CHECK_ARG(a && b && out && n > 0, "null pointer / n");
  QBN_CHECK_ARG(sa > 0 && sb > 0 && so > 0, "scales must be > 0");
  QBN_CHECK_ARG(lo >= 0 && hi <= 255 && lo <= hi, "0<=lo<=hi<=255");
  if (n_vec < 0) n_vec = (n / 64) * 64;
  i8_add_kernel<<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(a, sa, sa * (float)(-za), za, b, sb, sb * (float)(-zb), zb, n,
                                                                      n_vec, 1.0f / so, zo, lo, hi, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// quint8 ReLU (zero point is the floor) + clamp_activation, and k x k average pooling of an NHWC quint8 map
// (ATen qavg_pool2d: acc = sum - k*k*z ; q = clamp(rint(fp32(acc) * fp32(1/(k*k))) + z, 0, 255))
__global__ void i8_relu_kernel(const uint8_t* __restrict__ x, int64_t n, int floor_q, int lo, int hi, uint8_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    int q = (int)x[i];
    q = q < floor_q ? floor_q : q;
    out[i] = (uint8_t)clampi(q, lo, hi);
  }
}
extern "C" int qbn_i8_relu(const uint8_t* x, int64_t n, int32_t z_x, int lo, int hi, uint8_t* out, void* stream) {
  QBN_CHECK_ARG(x && out && n > 0, "null pointer / n");
  QBN_CHECK_ARG(lo >= 0 && hi <= 255 && lo <= hi && z_x >= 0 && z_x <= 255, "0<=lo<=hi<=255, zero point in range");
  i8_relu_kernel<<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(x, n, z_x, lo, hi, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void i8_avgpool_kernel(const uint8_t* __restrict__ x, int64_t B, int H, int W, int C, int k, int zx, float inv_area,
                                  int lo, int hi, uint8_t* __restrict__ out) {
  const int Ho = H / k, Wo = W / k;
  const int64_t total = B * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int64_t b = t / Ho;
    int acc = 0;
    for (int r = 0; r < k; ++r)
      for (int q = 0; q < k; ++q) acc += (int)x[((b * H + ho * k + r) * W + wo * k + q) * C + c];
    acc -= k * k * zx;
    int v = (int)rintf(__fmul_rn((float)acc, inv_area)) + zx;
    v = clampi(v, 0, 255);
    out[i] = (uint8_t)clampi(v, lo, hi);
  }
}
extern "C" int qbn_i8_avgpool(const uint8_t* x, int64_t B, int H, int W, int C, int k, int32_t z_x, int lo, int hi, uint8_t* out,
                              void* stream) {
  QBN_CHECK_ARG(x && out && B > 0 && H > 0 && W > 0 && C > 0, "args");
  QBN_CHECK_ARG(k > 0 && H % k == 0 && W % k == 0, "k must divide H and W (AvgPool2d(k), stride k, no padding)");
  QBN_CHECK_ARG(lo >= 0 && hi <= 255 && lo <= hi, "0<=lo<=hi<=255");
  i8_avgpool_kernel<<<qbn_grid_for(B * (H / k) * (W / k) * C, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, k, z_x,
                                                                                                 1.0f / (float)(k * k), lo, hi, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void i8_dropout_kernel(const uint8_t* __restrict__ x, int zx, int64_t rows, int64_t hw, int64_t C,
                                  const float* __restrict__ mask, float keep, float inv_sm, int zm, float mult, uint64_t seed,
                                  uint32_t sa, uint32_t sb, int64_t rows_per_sample, int lo, int hi, uint8_t* __restrict__ out, const uint32_t* __restrict__ sbase) {
  if (sbase) sb += *sbase;

  int64_t total = rows * hw * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t c = i % C;
    int64_t b = i / (hw * C);
    // rows_per_sample > 0: the rows are n Monte-Carlo samples x rows_per_sample images; sample s draws stream sb + s
    const int64_t s_idx = rows_per_sample > 0 ? b / rows_per_sample : 0;
    const int64_t bl = rows_per_sample > 0 ? b - s_idx * rows_per_sample : b;
    float m = mask ? mask[b * C + c] : (philox_uniform1(seed, sa, sb + (uint32_t)s_idx, (uint64_t)(bl * C + c)) < keep ? 1.0f : 0.0f);
    int mq = clampi((int)rintf(__fmul_rn(m, inv_sm)) + zm, 0, 255);  // dropout.py:34
    int prod = ((int)x[i] - zx) * (mq - zm);
    int q = (int)rintf(__fmul_rn((float)prod, mult)) + zm;           // quantized::mul, out at (s_m, z_m)
    out[i] = (uint8_t)clampi(q, lo, hi);
  }
}

extern "C" int qbn_i8_dropout(const uint8_t* x, float s_x, int32_t z_x, int64_t rows, int64_t hw, int64_t C, const float* mask,
                              float keep_prob, float s_m, int32_t z_m, uint64_t seed, uint32_t sa, uint32_t sb, int lo, int hi,
                              uint8_t* out, void* stream) {
  QBN_CHECK_ARG(x && out, "x/out");
  QBN_CHECK_ARG(rows > 0 && hw > 0 && C > 0 && s_x > 0 && s_m > 0, "sizes/scales");
  float mult = s_x * s_m * (1.0f / s_m);
  i8_dropout_kernel<<<qbn_grid_for(rows * hw * C, 256), 256, 0, (cudaStream_t)stream>>>(x, z_x, rows, hw, C, mask, keep_prob,
                                                                                     1.0f / s_m, z_m, mult, seed, sa, sb, 0, lo, hi, out, nullptr);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_i8_dropout_mc(const uint8_t* x, float s_x, int32_t z_x, int n_samples, int64_t rows_per_sample, int64_t hw, int64_t C,
                                 float keep_prob, float s_m, int32_t z_m, uint64_t seed, uint32_t site, uint32_t sample0, int lo, int hi,
                                 uint8_t* out, void* stream) {
  QBN_CHECK_ARG(x && out, "x/out");
  QBN_CHECK_ARG(n_samples > 0 && rows_per_sample > 0 && hw > 0 && C > 0 && s_x > 0 && s_m > 0, "sizes/scales");
  float mult = s_x * s_m * (1.0f / s_m);
  const int64_t rows = (int64_t)n_samples * rows_per_sample;
  i8_dropout_kernel<<<qbn_grid_for(rows * hw * C, 256), 256, 0, (cudaStream_t)stream>>>(x, z_x, rows, hw, C, nullptr, keep_prob, 1.0f / s_m, z_m,
                                                                                     mult, seed, site, sample0, rows_per_sample, lo, hi, out, g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// A11 glue at resolution changes
// ---------------------------------------------------------------------------------------------
__global__ void maxpool2x2_kernel(const float* __restrict__ x, int64_t B, int H, int W, int C, float* __restrict__ out) {
  int Ho = H >> 1, Wo = W >> 1;
  int64_t total = B * Ho * Wo * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t t = i / C;
    int wo = (int)(t % Wo);
    t /= Wo;
    int ho = (int)(t % Ho);
    int64_t b = t / Ho;
    const float* p = x + ((b * H + 2 * ho) * W + 2 * wo) * C + c;
    out[i] = fmaxf(fmaxf(p[0], p[C]), fmaxf(p[(int64_t)W * C], p[(int64_t)W * C + C]));
  }
}
extern "C" int qbn_maxpool2x2(const float* x, int64_t B, int H, int W, int C, float* out, void* stream) {
  QBN_CHECK_ARG(x && out && B > 0 && H > 1 && W > 1 && C > 0, "args");
  maxpool2x2_kernel<<<qbn_grid_for(B * (H / 2) * (W / 2) * C, 256), 256, 0, (cudaStream_t)stream>>>(x, B, H, W, C, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void avgpool_all_kernel(const float* __restrict__ x, int64_t B, int HW, int C, float divisor, float* __restrict__ out) {
  int64_t total = B * C;
  const float inv = 1.0f / divisor;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int c = (int)(i % C);
    int64_t b = i / C;
    float acc = 0.f;
    for (int p = 0; p < HW; ++p) acc += x[(b * HW + p) * C + c];
    out[i] = acc * inv;
  }
}
extern "C" int qbn_avgpool_all(const float* x, int64_t B, int HW, int C, float divisor, float* out, void* stream) {
  QBN_CHECK_ARG(x && out && B > 0 && HW > 0 && C > 0, "args");
  if (divisor <= 0.f) divisor = (float)HW;   // zero-bordered maps: sum over the padded plane / interior size
  avgpool_all_kernel<<<qbn_grid_for(B * C, 256), 256, 0, (cudaStream_t)stream>>>(x, B, HW, C, divisor, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ x, int64_t B, int C, int HW, float* __restrict__ out) {
  // 32x32 smem transpose of the [C][HW] plane of each image
  __shared__ float tile[32][33];
  int64_t b = blockIdx.z;
  int c0 = blockIdx.y * 32, p0 = blockIdx.x * 32;
  const float* xb = x + b * (int64_t)C * HW;
  float* ob = out + b * (int64_t)C * HW;
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int c = c0 + j, p = p0 + threadIdx.x;
    tile[j][threadIdx.x] = (c < C && p < HW) ? xb[(int64_t)c * HW + p] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += blockDim.y) {
    int p = p0 + j, c = c0 + threadIdx.x;
    if (c < C && p < HW) ob[(int64_t)p * C + c] = tile[threadIdx.x][j];
  }
}
extern "C" int qbn_nchw_to_nhwc(const float* x, int64_t B, int C, int HW, float* out, void* stream) {
  QBN_CHECK_ARG(x && out && B > 0 && B <= 65535 && C > 0 && HW > 0, "args");
  dim3 grid((HW + 31) / 32, (C + 31) / 32, (unsigned)B);
  nchw_to_nhwc_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(x, B, C, HW, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// planar-C4 path (p4_layout.cuh): weight blocking, blocked A4 sampler, pooling
// ---------------------------------------------------------------------------------------------
#include "p4_layout.cuh"

__global__ void p4_block_weights_kernel(const float* __restrict__ w, P4Block g, float* __restrict__ out) {
  const float* ws = w + (int64_t)blockIdx.y * g.N * g.K;
  float4* os = reinterpret_cast<float4*>(out) + (int64_t)blockIdx.y * g.total4;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < g.total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = p4_canonical(g, i);
    os[i] = idx < 0 ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(ws + idx);
  }
}
extern "C" int qbn_p4_block_weights(const float* w_ohwi, int n_mats, int N, int C, int taps, int stride, int cb_override, float* out,
                                    void* stream) {
  QBN_CHECK_ARG(w_ohwi && out && n_mats > 0 && n_mats <= 65535, "args");
  P4Block g;
  if (!p4_block_geom(N, C, taps, stride, g, cb_override)) {
    qbn_set_error("qbn_p4_block_weights: needs C %% 8 == 0 and N <= 256 (C=%d N=%d)", C, N);
    return QBN_ERR_UNSUPPORTED;
  }
  int gx = (int)((g.total4 + 255) / 256);
  int cap = (qbn_sm_count() * 8 + n_mats - 1) / n_mats;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  p4_block_weights_kernel<<<dim3(gx, n_mats), 256, 0, (cudaStream_t)stream>>>(w_ohwi, g, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// A4 on blocked operands: W[s] = mu + sigma * eps_s written straight into the smem image the conv kernel
// bulk-copies.  The Philox counter is the CANONICAL OHWI element index / 4, so the sampled values are
// bit-identical to qbn_sample_weights' whatever the blocking.
__global__ void sample_weights_blocked_kernel(const float* __restrict__ mu_b, const float* __restrict__ sigma_b, P4Block g,
                                              const float* __restrict__ eps, uint64_t seed, uint32_t layer_id, uint32_t sample0,
                                              float* __restrict__ w, int round_tf32, const uint32_t* __restrict__ sbase) {
  if (sbase) sample0 += *sbase;      // per-call draw offset read on the device: one CUDA graph serves every batch

  const int s = blockIdx.y;
  float4* ws = reinterpret_cast<float4*>(w) + (int64_t)s * g.total4;
  const float* es = eps ? eps + (int64_t)s * g.N * g.K : nullptr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < g.total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = p4_canonical(g, i);
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx >= 0) {
      float z[4];
      if (es) {
        const float4 e = *reinterpret_cast<const float4*>(es + idx);
        z[0] = e.x; z[1] = e.y; z[2] = e.z; z[3] = e.w;
      } else {
        philox_normal4(seed, layer_id, sample0 + (uint32_t)s, (uint64_t)(idx >> 2), z);
      }
      const float4 m = reinterpret_cast<const float4*>(mu_b)[i];
      const float4 sg = reinterpret_cast<const float4*>(sigma_b)[i];
      o.x = __fadd_rn(m.x, __fmul_rn(z[0], sg.x));
      o.y = __fadd_rn(m.y, __fmul_rn(z[1], sg.y));
      o.z = __fadd_rn(m.z, __fmul_rn(z[2], sg.z));
      o.w = __fadd_rn(m.w, __fmul_rn(z[3], sg.w));
      if (round_tf32) { o.x = tf32_round(o.x); o.y = tf32_round(o.y); o.z = tf32_round(o.z); o.w = tf32_round(o.w); }
    }
    ws[i] = o;
  }
}
extern "C" int qbn_sample_weights_blocked(const float* mu_b, const float* sigma_b, int N, int C, int taps, int stride, int n_samples,
                                          const float* eps, uint64_t seed, uint32_t layer_id, uint32_t sample0, float* w,
                                          int round_tf32, void* stream) {
  QBN_CHECK_ARG(mu_b && sigma_b && w, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && n_samples <= 65535, "0<n_samples<=65535");
  P4Block g;
  if (!p4_block_geom(N, C, taps, stride, g)) {
    qbn_set_error("qbn_sample_weights_blocked: needs C %% 8 == 0 and N <= 256 (C=%d N=%d)", C, N);
    return QBN_ERR_UNSUPPORTED;
  }
  int gx = (int)((g.total4 + 255) / 256);
  int cap = (qbn_sm_count() * 8 + n_samples - 1) / n_samples;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  sample_weights_blocked_kernel<<<dim3(gx, n_samples), 256, 0, (cudaStream_t)stream>>>(mu_b, sigma_b, g, eps, seed, layer_id, sample0, w,
                                                                                     round_tf32, g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// global average pool of planar-C4 maps: x [C/4][n_img * HW][4] -> out [n_img][C].  The zero border contributes nothing; divisor =
// interior size.  Small maps (HW <= 64: the 5x5 padded map that ends the ResNet): one THREAD per (image, chunk) — HW independent
// 16-byte loads of one contiguous run in flight per thread, consecutive threads on consecutive images of a plane, no shuffles
// (the warp-per-map version ran at 2 TB/s: 25 of 32 lanes busy, one load per 20 shuffles).  Larger maps: one warp per (image, chunk).
__global__ void avgpool_p4_small_kernel(const float* __restrict__ x, int64_t n_img, int HW, int64_t plane, int C, float inv, float* __restrict__ out) {
  const int chunks = C >> 2;
  const int64_t total = n_img * chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j = (int)(i / n_img);
    const int64_t img = i - (int64_t)j * n_img;
    const float4* src = reinterpret_cast<const float4*>(x) + (int64_t)j * plane + img * HW;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 5
    for (int r = 0; r < HW; ++r) {
      const float4 v = src[r];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    *reinterpret_cast<float4*>(out + img * C + 4 * j) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
}
__global__ void avgpool_p4_kernel(const float* __restrict__ x, int64_t n_img, int HW, int64_t plane, int C, float inv, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int chunks = C >> 2;
  for (int64_t wi = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wi < n_img * chunks; wi += warps) {
    const int64_t img = wi / chunks;
    const int j = (int)(wi - img * chunks);
    const float4* src = reinterpret_cast<const float4*>(x) + (int64_t)j * plane + img * HW;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = lane; r < HW; r += 32) {
      const float4 v = src[r];
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a.x += __shfl_xor_sync(0xffffffffu, a.x, o);
      a.y += __shfl_xor_sync(0xffffffffu, a.y, o);
      a.z += __shfl_xor_sync(0xffffffffu, a.z, o);
      a.w += __shfl_xor_sync(0xffffffffu, a.w, o);
    }
    if (lane == 0) *reinterpret_cast<float4*>(out + img * C + 4 * j) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
  }
}
extern "C" int qbn_avgpool_p4(const float* x, int64_t n_img, int HW, int64_t plane_rows, int C, float divisor, float* out, void* stream) {
  QBN_CHECK_ARG(x && out && n_img > 0 && HW > 0 && C > 0 && C % 4 == 0 && divisor > 0.f && plane_rows >= n_img * HW, "args");
  if (HW <= 64) {
    avgpool_p4_small_kernel<<<qbn_grid_for(n_img * (C / 4), 256, 16), 256, 0, (cudaStream_t)stream>>>(x, n_img, HW, plane_rows, C, 1.0f / divisor, out);
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
  const int64_t warps = n_img * (C / 4);
  int64_t blocks = (warps + 7) / 8;
  const int64_t cap = (int64_t)qbn_sm_count() * 16;
  if (blocks > cap) blocks = cap;
  avgpool_p4_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, n_img, HW, plane_rows, C, 1.0f / divisor, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// All planar layers of one Monte-Carlo chunk in ONE launch (21 launches -> 1): blockIdx.z = layer job.
struct qbn_p4_sample_job_dev {
  const float* mu_b; const float* sigma_b; const float* eps; float* w;
  int N, C, taps, stride; uint32_t layer_id; int n_stack;
  const float* chan_scale; int cb_override; int w_sample_stride4;
  int s_off; int pad_;          // stacked jobs: this job covers the chunk's samples [s_off, s_off + n_stack)
};
__global__ void sample_weights_blocked_multi_kernel(const qbn_p4_sample_job_dev* __restrict__ jobs, uint64_t seed, uint32_t sample0,
                                                    int round_tf32, const uint32_t* __restrict__ sbase) {
  if (sbase) sample0 += *sbase;      // per-call draw offset read on the device: one CUDA graph serves every batch

  const qbn_p4_sample_job_dev jb = jobs[blockIdx.z];
  P4Block g;
  g.N = jb.N; g.C = jb.C; g.taps = jb.taps;
  g.CB = jb.cb_override > 0 ? jb.cb_override : qbn_p4_block_channels(jb.C, jb.stride, jb.taps);
  g.cbc = g.CB / 4; g.n_pad = qbn_p4_n_pad(jb.N); g.K = jb.taps * jb.C;
  g.total4 = (int64_t)(jb.C / g.CB) * jb.taps * g.cbc * g.n_pad;
  const int s = blockIdx.y;
  // stacked: ONE blocked tensor of n_stack*N rows; sample s fills rows [(s-s_off)*N, (s-s_off+1)*N) of every (cb, tap, chunk)
  // column; a chunk larger than one accumulator tile is covered by several stacked jobs (groups of samples)
  if (jb.n_stack > 0 && (s < jb.s_off || s >= jb.s_off + jb.n_stack)) return;
  const int n_pad_out = jb.n_stack > 0 ? qbn_p4_n_pad(jb.n_stack * jb.N) : g.n_pad;
  float4* ws = reinterpret_cast<float4*>(jb.w) +
               (jb.n_stack > 0 ? (int64_t)(s - jb.s_off) * g.N : (int64_t)s * (jb.w_sample_stride4 > 0 ? (int64_t)jb.w_sample_stride4 : g.total4));
  const float* es = jb.eps ? jb.eps + (int64_t)s * g.N * g.K : nullptr;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < g.total4; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t idx = p4_canonical(g, i);
    if (jb.n_stack > 0 && idx < 0) continue;               // padding rows of the stacked tensor are zeroed by the caller
    float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx >= 0) {
      float z[4];
      if (es) {
        const float4 e = *reinterpret_cast<const float4*>(es + idx);
        z[0] = e.x; z[1] = e.y; z[2] = e.z; z[3] = e.w;
      } else {
        philox_normal4(seed, jb.layer_id, sample0 + (uint32_t)s, (uint64_t)(idx >> 2), z);
      }
      const float4 m = reinterpret_cast<const float4*>(jb.mu_b)[i];
      const float4 sg = reinterpret_cast<const float4*>(jb.sigma_b)[i];
      o.x = __fadd_rn(m.x, __fmul_rn(z[0], sg.x));
      o.y = __fadd_rn(m.y, __fmul_rn(z[1], sg.y));
      o.z = __fadd_rn(m.z, __fmul_rn(z[2], sg.z));
      o.w = __fadd_rn(m.w, __fmul_rn(z[3], sg.w));
      if (jb.chan_scale) {     // eval BatchNorm scale folded into the sampled weights (two branches share one accumulator)
        const float cs = jb.chan_scale[(int)(i % g.n_pad)];
        o.x *= cs; o.y *= cs; o.z *= cs; o.w *= cs;
      }
      if (round_tf32) { o.x = tf32_round(o.x); o.y = tf32_round(o.y); o.z = tf32_round(o.z); o.w = tf32_round(o.w); }
    }
    const int64_t col = i / g.n_pad;                         // (cb, tap, chunk) column, row n = i % n_pad
    ws[jb.n_stack > 0 ? col * n_pad_out + (i - col * g.n_pad) : i] = o;
  }
}
extern "C" int qbn_sample_weights_blocked_multi(const void* jobs_dev, int n_jobs, int64_t max_floats_per_sample, int n_samples,
                                                uint64_t seed, uint32_t sample0, int round_tf32, void* stream) {
  QBN_CHECK_ARG(jobs_dev && n_jobs > 0 && n_jobs <= 65535 && n_samples > 0 && n_samples <= 65535 && max_floats_per_sample > 0, "args");
  // grid.x is sized for the largest layer and shared by all jobs: 16 float4 items per thread there keeps the number of blocks that
  // find nothing to do in the small layers low (4 items: 98.6 us for 13 samples of the ResNet, 16: 88.1, 64: 128.5)
  const int64_t per_block = 256 * 16;
  int64_t gx = (max_floats_per_sample / 4 + per_block - 1) / per_block;
  if (gx < 1) gx = 1;
  if (gx > 4096) gx = 4096;
  sample_weights_blocked_multi_kernel<<<dim3((unsigned)gx, n_samples, n_jobs), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const qbn_p4_sample_job_dev*>(jobs_dev), seed, sample0, round_tf32, g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// A8 masks of every dropout site of a Monte-Carlo chunk in one launch (blockIdx.z = site, blockIdx.y = sample)
struct qbn_mask_job_dev { float* out; int64_t elems; uint32_t site_id; int pad_; };
__global__ void dropout_masks_multi_kernel(const qbn_mask_job_dev* __restrict__ jobs, float keep, uint64_t seed, uint32_t sample0, const uint32_t* __restrict__ sbase) {
  if (sbase) sample0 += *sbase;      // per-call draw offset read on the device: one CUDA graph serves every batch

  const qbn_mask_job_dev jb = jobs[blockIdx.z];
  const int s = blockIdx.y;
  float* o = jb.out + (int64_t)s * jb.elems;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < jb.elems; i += (int64_t)gridDim.x * blockDim.x)
    o[i] = philox_uniform1(seed, jb.site_id, sample0 + (uint32_t)s, (uint64_t)i) < keep ? 1.0f : 0.0f;
}
extern "C" int qbn_dropout_masks_multi(const void* jobs_dev, int n_jobs, int64_t max_elems, int n_samples, float keep_prob, uint64_t seed,
                                       uint32_t sample0, void* stream) {
  QBN_CHECK_ARG(jobs_dev && n_jobs > 0 && n_jobs <= 65535 && n_samples > 0 && n_samples <= 65535 && max_elems > 0, "args");
  int64_t gx = (max_elems + 255) / 256;
  if (gx > 1024) gx = 1024;
  dropout_masks_multi_kernel<<<dim3((unsigned)gx, n_samples, n_jobs), 256, 0, (cudaStream_t)stream>>>(
      reinterpret_cast<const qbn_mask_job_dev*>(jobs_dev), keep_prob, seed, sample0, g_sample_base);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// A5 for a whole model in ONE launch (21 launches -> 1): blockIdx.y = layer job; the total goes to one scalar
struct qbn_kl_job_dev { const float* mu; const float* rho; float* d_mu; float* d_rho; int64_t n; float sigma_prior; int pad_; };
__global__ void kl_multi_kernel(const qbn_kl_job_dev* __restrict__ jobs, float* __restrict__ kl_out, float gscale) {
  const qbn_kl_job_dev jb = jobs[blockIdx.y];
  const float sp = jb.sigma_prior, inv_sp = 1.0f / sp, inv_sp2 = inv_sp * inv_sp;
  double acc = 0.0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < jb.n; i += (int64_t)gridDim.x * blockDim.x) {
    float gm, gr;
    acc += (double)kl_term(jb.mu[i], jb.rho[i], sp, inv_sp, inv_sp2, gscale, gm, gr);
    if (jb.d_mu) jb.d_mu[i] = gm;
    if (jb.d_rho) jb.d_rho[i] = gr;
  }
  __shared__ double sh[32];
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (blockDim.x >> 5) ? sh[lane] : 0.0;
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0 && acc != 0.0) atomicAdd(kl_out, (float)(0.5 * acc));
  }
}
extern "C" int qbn_kl_multi(const void* jobs_dev, int n_jobs, int64_t max_n, float* kl_out, float grad_scale, void* stream) {
  QBN_CHECK_ARG(jobs_dev && kl_out && n_jobs > 0 && n_jobs <= 65535 && max_n > 0, "args");
  int64_t gx = (max_n + 1023) / 1024;
  if (gx > 64) gx = 64;
  kl_multi_kernel<<<dim3((unsigned)gx, n_jobs), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const qbn_kl_job_dev*>(jobs_dev), kl_out,
                                                                                grad_scale);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}


// trainer.py:105-107 (`p.grad[p.grad != p.grad] = 0` for every parameter) for a whole model in ONE launch: blockIdx.y = tensor job.
// The job list travels in the kernel parameters (no device table, nothing to copy: safe inside a CUDA graph capture, and gradient
// tensors may move between steps).
struct qbn_scrub_job_dev { float* g; int64_t n; };
constexpr int SCRUB_MAX_JOBS = 128;
struct ScrubBatch { qbn_scrub_job_dev j[SCRUB_MAX_JOBS]; };
__global__ void scrub_nan_multi_kernel(const __grid_constant__ ScrubBatch jobs) {
  const qbn_scrub_job_dev jb = jobs.j[blockIdx.y];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < jb.n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = jb.g[i];
    if (v != v) jb.g[i] = 0.0f;
  }
}
extern "C" int qbn_scrub_nan_multi(const qbn_scrub_job* jobs_host, int n_jobs, void* stream) {
  QBN_CHECK_ARG(jobs_host && n_jobs > 0, "args");
  for (int j0 = 0; j0 < n_jobs; j0 += SCRUB_MAX_JOBS) {
    const int nj = n_jobs - j0 < SCRUB_MAX_JOBS ? n_jobs - j0 : SCRUB_MAX_JOBS;
    ScrubBatch b;
    int64_t max_n = 1;
    for (int j = 0; j < nj; ++j) {
      QBN_CHECK_ARG(jobs_host[j0 + j].grad && jobs_host[j0 + j].n >= 0, "job");
      b.j[j].g = jobs_host[j0 + j].grad; b.j[j].n = jobs_host[j0 + j].n;
      if (b.j[j].n > max_n) max_n = b.j[j].n;
    }
    int64_t gx = (max_n + 1023) / 1024;
    if (gx > 64) gx = 64;
    scrub_nan_multi_kernel<<<dim3((unsigned)gx, nj), 256, 0, (cudaStream_t)stream>>>(b);
    QBN_CHECK_LAUNCH();
  }
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// SGHMC / SGLD parameter update (src/models/stochastic/sgld/utils_sgld.py:30-92), one fused pass per parameter tensor
// instead of ~35 elementwise launches: weight decay into the gradient, burn-in preconditioner (tau, g, V_hat), optional momentum
// resampling, friction + injected Gaussian noise, NaN/inf scrub of the momentum, parameter step.  z_mom / z_noise: standard
// normals injected by the caller (parity tests) or NULL -> Philox(seed, stream_a, stream_b / stream_b + 1, element).
// Every fp32 operation is rounded separately, in the order torch evaluates the reference's expressions.
// ---------------------------------------------------------------------------------------------
__global__ void sghmc_step_kernel(float* __restrict__ p, float* __restrict__ grad, float* __restrict__ tau, float* __restrict__ g,
                                  float* __restrict__ V, float* __restrict__ v, int64_t n, float wd, float lr2, float lr4, float base_C,
                                  float eps, int burn_in, int resample, const float* __restrict__ z_mom, const float* __restrict__ z_noise,
                                  uint64_t seed, uint32_t sa, uint32_t sb) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float d = grad[i];
    if (wd != 0.f) d = __fmaf_rn(wd, p[i], d);                          // d_p.add_(p.data, alpha=weight_decay): ATen's add-with-alpha is one fused multiply-add; in place, like the reference
    grad[i] = d;
    float Vh = V[i];
    if (burn_in) {
      float t = tau[i], gg = g[i];
      t = __fadd_rn(t, __fadd_rn(__fdiv_rn(__fmul_rn(-t, __fmul_rn(gg, gg)), __fadd_rn(Vh, eps)), 1.0f));
      const float ti = __fdiv_rn(1.0f, __fadd_rn(t, eps));
      gg = __fadd_rn(gg, __fadd_rn(__fmul_rn(-ti, gg), __fmul_rn(ti, d)));
      Vh = __fadd_rn(Vh, __fadd_rn(__fmul_rn(-ti, Vh), __fmul_rn(ti, __fmul_rn(d, d))));
      tau[i] = t; g[i] = gg; V[i] = Vh;
    }
    const float vis = __fdiv_rn(1.0f, __fadd_rn(sqrtf(Vh), eps));        // V_inv_sqrt
    float vm = v[i];
    if (resample) {
      const float z = z_mom ? z_mom[i] : philox_normal1(seed, sa, sb, (uint64_t)i);
      vm = __fmul_rn(z, sqrtf(__fmul_rn(lr2, vis)));                     // torch.normal(0, sqrt(lr^2 * V_inv_sqrt))
    }
    const float nvar = __fadd_rn(__fmul_rn(__fmul_rn(__fmul_rn(2.0f, lr2), vis), base_C), -lr4);
    const float nstd = sqrtf(fmaxf(nvar, 1e-16f));
    const float z = z_noise ? z_noise[i] : philox_normal1(seed, sa, sb + 1u, (uint64_t)i);
    const float ns = __fmul_rn(z, nstd);
    // v.add_(-(lr^2) * V_inv_sqrt * d_p - base_C * v + noise)
    vm = __fadd_rn(vm, __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(-lr2, vis), d), -__fmul_rn(base_C, vm)), ns));
    if (vm != vm || isinf(vm)) vm = 0.f;                                 // utils_sgld.py:86-88
    v[i] = vm;
    p[i] = __fadd_rn(p[i], vm);
  }
}
extern "C" int qbn_sghmc_step(float* p, float* grad, float* tau, float* g, float* V_hat, float* v_momentum, int64_t n, float weight_decay,
                              float lr, float base_C, float eps, int burn_in, int resample_momentum, const float* z_momentum,
                              const float* z_noise, uint64_t seed, uint32_t stream_a, uint32_t stream_b, void* stream) {
  QBN_CHECK_ARG(p && grad && tau && g && V_hat && v_momentum && n > 0, "null pointer / n");
  const float lr2 = (float)((double)lr * (double)lr), lr4 = (float)((double)lr * lr * lr * lr);       // python: lr ** 2, lr ** 4 in double
  sghmc_step_kernel<<<qbn_grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>(p, grad, tau, g, V_hat, v_momentum, n, weight_decay, lr2, lr4, base_C,
                                                                           eps, burn_in, resample_momentum, z_momentum, z_noise, seed, stream_a,
                                                                           stream_b);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
