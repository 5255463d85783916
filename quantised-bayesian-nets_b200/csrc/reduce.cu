// A9/A10: Monte-Carlo aggregation and metric reductions as warp-shuffle kernels.
//   experiments/utils.py:344-355 (stack + mean over S), src/metrics.py:20-29,48-57,76-85,104-112
//   (error / NLL / Brier / entropy sums), :381-383 (ECE, 10 equal-width bins, l1),
//   :135-157,176-225 (regression NLL / MSE / MAE).
// The tensors are tiny ([S,B,10] = 1 MB at S=100,B=256): the kernels are latency-bound, so each
// is a single pass with deterministic in-block reductions and no host synchronisation.
#include "common.cuh"

namespace {

// one block (8 warps) per image; warps stride over samples, lanes over classes
__global__ void softmax_accumulate_kernel(const float* __restrict__ logits, int S, int B, int K, float* __restrict__ psum,
                                          int accumulate, int b_first, int b_end) {
  extern __shared__ float sh[];  // [8][K]
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int per = (K + 31) / 32;  // classes per lane (K <= 128)
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int s = wid; s < S; s += nw) {
    if ((s == 0 && b < b_first) || (s == S - 1 && b >= b_end)) continue;   // unit window: warp-uniform
    const float* row = logits + ((int64_t)s * B + b) * K;
    float v[4];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int k = lane + 32 * j;
      v[j] = (j < per && k < K) ? row[k] : -INFINITY;
      mx = fmaxf(mx, v[j]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[j] = (v[j] == -INFINITY) ? 0.f : expf(v[j] - mx);
      sum += v[j];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[j] += v[j] / sum;
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int k = lane + 32 * j;
    if (k < K) sh[wid * K + k] = acc[j];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sh[w * K + k];
    int64_t o = (int64_t)b * K + k;
    psum[o] = accumulate ? psum[o] + t : t;
  }
}

__global__ void mc_mean_kernel(const float* __restrict__ probs, int S, int64_t n, float* __restrict__ mean) {
  const float fs = (float)S;
  const bool vec = (n & 3) == 0 && ((reinterpret_cast<uintptr_t>(probs) | reinterpret_cast<uintptr_t>(mean)) & 15) == 0;
  if (vec) {
    int64_t n4 = n >> 2;
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
      float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
      for (int s = 0; s < S; ++s) {
        float4 v = reinterpret_cast<const float4*>(probs + (int64_t)s * n)[i];
        a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
      }
      reinterpret_cast<float4*>(mean)[i] = make_float4(a.x / fs, a.y / fs, a.z / fs, a.w / fs);
    }
  } else {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
      float a = 0.f;
      for (int s = 0; s < S; ++s) a += probs[(int64_t)s * n + i];
      mean[i] = a / fs;
    }
  }
}

__global__ void reg_mc_reduce_kernel(const float* __restrict__ mu, const float* __restrict__ var, int S, int64_t B,
                                     float* __restrict__ mean_out, float* __restrict__ var_out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    float sm = 0.f, sv = 0.f;
    for (int s = 0; s < S; ++s) { sm += mu[(int64_t)s * B + i]; sv += var[(int64_t)s * B + i]; }
    float m = sm / (float)S;
    float ss = 0.f;
    for (int s = 0; s < S; ++s) { float d = mu[(int64_t)s * B + i] - m; ss += d * d; }
    mean_out[i] = m;
    // torch.var (unbiased) of the means + mean of the variances (experiments/utils.py:353)
    var_out[i] = ss / (float)(S > 1 ? S - 1 : 1) + sv / (float)S;
  }
}

struct Bounds { float b[33]; };

// single block: warps stride over rows; per-warp partials in smem; fixed-order final reduction
__global__ void cls_metrics_kernel(const float* __restrict__ probs, const int64_t* __restrict__ target, int B, int K, float scale,
                                   int n_bins, Bounds bounds, float* __restrict__ out) {
  extern __shared__ float sh[];  // [nw][4 + 3*n_bins]
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int stride = 4 + 3 * n_bins;
  float* mine = sh + wid * stride;
  for (int i = lane; i < stride; i += 32) mine[i] = 0.f;
  __syncwarp();
  float err = 0.f, nll = 0.f, brier = 0.f, ent = 0.f;
  for (int r = wid; r < B; r += nw) {
    const float* row = probs + (int64_t)r * K;
    int t = (int)target[r];
    float best = -INFINITY; int besti = 0x7fffffff;
    float lb = 0.f, le = 0.f, pt = 0.f;
    for (int k = lane; k < K; k += 32) {
      float p = row[k] * scale;
      if (p > best) { best = p; besti = k; }      // first max within the lane (k ascending)
      float oh = (k == t) ? 1.f : 0.f;
      lb += (p - oh) * (p - oh);
      le += -p * logf(p + 1e-8f);
      if (k == t) pt = p;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {  // argmax with lowest-index tie break (torch.argmax)
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, besti, o);
      if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    lb = warp_sum(lb); le = warp_sum(le); pt = warp_sum(pt);
    if (lane == 0) {
      bool correct = besti == t;
      err += correct ? 0.f : 1.f;
      nll += -logf(pt + 1e-8f);
      brier += lb;
      ent += le;
      int bin = 0;
      for (int q = 1; q < n_bins; ++q) bin += (best >= bounds.b[q]) ? 1 : 0;   // bucketize(right=True) - 1
      mine[4 + 3 * bin + 0] += best;
      mine[4 + 3 * bin + 1] += correct ? 1.f : 0.f;
      mine[4 + 3 * bin + 2] += 1.f;
    }
  }
  if (lane == 0) { mine[0] = err; mine[1] = nll; mine[2] = brier; mine[3] = ent; }
  __syncthreads();
  for (int i = threadIdx.x; i < stride; i += blockDim.x) {
    float t = 0.f;
    for (int w = 0; w < nw; ++w) t += sh[w * stride + i];
    out[i] += t;
  }
}

__global__ void reg_metrics_kernel(const float* __restrict__ mean, const float* __restrict__ var, const float* __restrict__ target,
                                   int64_t B, float* __restrict__ out) {
  __shared__ float sh[32][3];
  float nll = 0.f, se = 0.f, ae = 0.f;
  for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
    float m = mean[i], v = var[i], t = target[i];
    float d = t - m;
    // metrics.py:144: 0.5*log(2*pi*var + 1e-8) + (t-m)^2 / (2*var + 1e-8)
    nll += 0.5f * logf(6.283185307179586f * v + 1e-8f) + d * d / (2.0f * v + 1e-8f);
    se += d * d;
    ae += fabsf(d);
  }
  nll = warp_sum(nll); se = warp_sum(se); ae = warp_sum(ae);
  int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sh[wid][0] = nll; sh[wid][1] = se; sh[wid][2] = ae; }
  __syncthreads();
  if (threadIdx.x < 3) {
    float t = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += sh[w][threadIdx.x];
    out[threadIdx.x] += t;
  }
}

}  // namespace

// K <= 16 (the 10-class nets): one thread per (image, sample group) keeps its row in registers — consecutive lanes read
// consecutive images, i.e. contiguous rows — and the sample groups are reduced through shared memory in a fixed order.
template <int SG>
__global__ void softmax_accumulate_smallk_kernel(const float* __restrict__ logits, int S, int B, int K, float* __restrict__ psum,
                                                 int accumulate, int b_first, int b_end) {
  __shared__ float sh[SG][32][17];
  const int lane = threadIdx.x & 31, sg = threadIdx.x >> 5;
  const int b = blockIdx.x * 32 + lane;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  if (b < B) {
    for (int s = sg; s < S; s += SG) {
      if ((s == 0 && b < b_first) || (s == S - 1 && b >= b_end)) continue;   // unit window (dist.shard_units)
      const float* row = logits + ((int64_t)s * B + b) * K;
      float v[16];
      float mx = -INFINITY;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k] = k < K ? __ldg(row + k) : -INFINITY;
        mx = fmaxf(mx, v[k]);
      }
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        v[k] = k < K ? expf(v[k] - mx) : 0.f;
        sum += v[k];
      }
      const float inv = 1.0f / sum;
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[k] += v[k] * inv;
    }
  }
#pragma unroll
  for (int k = 0; k < 16; ++k) sh[sg][lane][k] = acc[k];
  __syncthreads();
  for (int i = threadIdx.x; i < 32 * K; i += blockDim.x) {
    const int l = i / K, k = i - l * K;
    const int bb = blockIdx.x * 32 + l;
    if (bb >= B) continue;
    float t = 0.f;
#pragma unroll
    for (int g = 0; g < SG; ++g) t += sh[g][l][k];
    const int64_t o = (int64_t)bb * K + k;
    psum[o] = accumulate ? psum[o] + t : t;
  }
}

// first_img / end_img: the unit window of the sample-sharded evaluation — sample 0 contributes images [first_img, B) only, sample
// n_samples - 1 images [0, end_img) only (dist.shard_units)
extern "C" int qbn_softmax_accumulate_window(const float* logits, int n_samples, int B, int K, int first_img, int end_img, float* psum,
                                             int accumulate, void* stream) {
  QBN_CHECK_ARG(logits && psum, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && B > 0 && K > 0 && K <= 128, "S,B > 0 and 0 < K <= 128");
  QBN_CHECK_ARG(first_img >= 0 && first_img < B && end_img > 0 && end_img <= B, "unit window outside the batch");
  if (K <= 16) {
    softmax_accumulate_smallk_kernel<8><<<(B + 31) / 32, 256, 0, (cudaStream_t)stream>>>(logits, n_samples, B, K, psum, accumulate, first_img, end_img);
    QBN_CHECK_LAUNCH();
    return QBN_OK;
  }
  softmax_accumulate_kernel<<<B, 256, 8 * K * sizeof(float), (cudaStream_t)stream>>>(logits, n_samples, B, K, psum, accumulate, first_img, end_img);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_softmax_accumulate(const float* logits, int n_samples, int B, int K, float* psum, int accumulate, void* stream) {
  return qbn_softmax_accumulate_window(logits, n_samples, B, K, 0, B, psum, accumulate, stream);
}

extern "C" int qbn_mc_mean(const float* probs, int n_samples, int64_t BK, float* mean, void* stream) {
  QBN_CHECK_ARG(probs && mean && n_samples > 0 && BK > 0, "args");
  mc_mean_kernel<<<qbn_grid_for((BK + 3) / 4, 128), 128, 0, (cudaStream_t)stream>>>(probs, n_samples, BK, mean);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_reg_mc_reduce(const float* mu, const float* var, int n_samples, int64_t B, float* mean_out, float* var_out,
                                 void* stream) {
  QBN_CHECK_ARG(mu && var && mean_out && var_out && n_samples > 0 && B > 0, "args");
  reg_mc_reduce_kernel<<<qbn_grid_for(B, 128), 128, 0, (cudaStream_t)stream>>>(mu, var, n_samples, B, mean_out, var_out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_cls_metrics(const float* probs, const int64_t* target, int B, int K, float scale, int n_bins, float* out,
                               void* stream) {
  QBN_CHECK_ARG(probs && target && out, "null pointer");
  QBN_CHECK_ARG(B > 0 && K > 0 && n_bins > 0 && n_bins <= 32, "B,K > 0, 0 < n_bins <= 32");
  // bin edges exactly as torch.linspace(0, 1, n_bins+1) computes them in fp32
  Bounds bd;
  int steps = n_bins + 1;
  float step = (1.0f - 0.0f) / (float)(steps - 1);
  int halfway = steps / 2;
  for (int i = 0; i < steps; ++i) bd.b[i] = i < halfway ? 0.0f + step * (float)i : 1.0f - step * (float)(steps - i - 1);
  int threads = 1024;
  size_t smem = (threads / 32) * (4 + 3 * n_bins) * sizeof(float);
  cls_metrics_kernel<<<1, threads, smem, (cudaStream_t)stream>>>(probs, target, B, K, scale, n_bins, bd, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_reg_metrics(const float* mean, const float* var, const float* target, int64_t B, float* out, void* stream) {
  QBN_CHECK_ARG(mean && var && target && out && B > 0, "args");
  reg_metrics_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(mean, var, target, B, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// ---------------------------------------------------------------------------------------------
// N4: the classification ELBO of src/losses.py:14-29 in ONE launch, value and gradient:
//   data = data_scale * mean_b -log(p[b, t_b] + 1e-8);  kl_term = kl * kl_scale;  loss = data + gamma * kl_term
//   d_probs[b, k] = -data_scale / (B * (p[b, t_b] + 1e-8)) at k == t_b, else 0        (d loss / d probs; the caller scales it)
// out[0..2] = {loss, data, kl_term}.  One CTA: the [B, K] output of a training batch is a few KB.
// ---------------------------------------------------------------------------------------------
__global__ void elbo_cls_kernel(const float* __restrict__ probs, const int64_t* __restrict__ target, const float* __restrict__ kl, int B, int K,
                                float data_scale, float kl_scale, float gamma, float* __restrict__ out, float* __restrict__ d_probs) {
  float acc = 0.f;
  for (int b = threadIdx.x; b < B; b += blockDim.x) {
    const int64_t t = target[b];
    const bool ok = t >= 0 && t < K;
    const float p = ok ? probs[(int64_t)b * K + t] + 1e-8f : 1.0f;
    acc += -logf(p);
    if (d_probs) {
      for (int k = 0; k < K; ++k) d_probs[(int64_t)b * K + k] = 0.f;
      if (ok) d_probs[(int64_t)b * K + t] = -data_scale / ((float)B * p);
    }
  }
  __shared__ float sh[32];
  acc = warp_sum(acc);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    acc = lane < (int)(blockDim.x >> 5) ? sh[lane] : 0.f;
    acc = warp_sum(acc);
    if (lane == 0) {
      const float data = data_scale * (acc / (float)B);
      const float klt = kl[0] * kl_scale;
      out[0] = data + gamma * klt;
      out[1] = data;
      out[2] = klt;
    }
  }
}
extern "C" int qbn_elbo_cls(const float* probs, const int64_t* target, const float* kl, int B, int K, float data_scale, float kl_scale,
                            float gamma, float* out3, float* d_probs, void* stream) {
  QBN_CHECK_ARG(probs && target && kl && out3 && B > 0 && K > 0, "args");
  elbo_cls_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(probs, target, kl, B, K, data_scale, kl_scale, gamma, out3, d_probs);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
