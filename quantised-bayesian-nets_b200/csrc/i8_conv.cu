// A6 step 6: u8 x s8 -> s32 implicit-GEMM convolution / linear with FBGEMM's requantisation
// epilogue (linear_q.py:93-94,168-172; conv_q.py:120-125,206-209), batched over Monte-Carlo
// samples (one sampled int8 weight tensor per sample).  Integer accumulation is exact, so any
// tiling is bit-identical to FBGEMM's; the epilogue reproduces its fp32 rounding sequence:
//   y = clamp(rint((fp32(acc) + bias/(s_x*s_w)) * ((s_x*s_w)/s_out)) + z_out, lo, hi)
// This file is the CUDA-core (IMAD) version used for ragged shapes; umma_conv.cu holds the
// tcgen05 kind::i8 version for the aligned ResNet shapes.
#include "common.cuh"

namespace {

constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

struct G8 {
  int B, H, W, C, N, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo, K;
  int64_t M;
};

struct RowI {
  int64_t base;
  int h0, w0, valid;
};

__global__ void __launch_bounds__(NT) i8_conv_kernel(G8 g, const uint8_t* __restrict__ x, int x_shared, int z_x,
                                                     const int8_t* __restrict__ w, int w_shared, int z_w,
                                                     const float* __restrict__ bias, float act_times_w, float mult, int z_out,
                                                     int lo, int hi, uint8_t* __restrict__ out, int32_t* __restrict__ acc_dump) {
  __shared__ int As[BK][BM + 4];
  __shared__ int Bs[BK][BN + 4];
  __shared__ RowI rows[BM];
  const int tid = threadIdx.x, z = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  const uint8_t* xs = x + (x_shared ? 0 : (int64_t)z * g.B * g.H * g.W * g.C);
  const int8_t* ws = w + (w_shared ? 0 : (int64_t)z * g.N * g.K);

  for (int i = tid; i < BM; i += NT) {
    int64_t m = m0 + i;
    RowI r;
    r.valid = m < g.M;
    int64_t mm = r.valid ? m : 0;
    int wo = (int)(mm % g.Wo);
    int64_t t = mm / g.Wo;
    int ho = (int)(t % g.Ho);
    int b = (int)(t / g.Ho);
    r.h0 = ho * g.sh - g.ph;
    r.w0 = wo * g.sw - g.pw;
    r.base = (int64_t)b * g.H * g.W * g.C;
    rows[i] = r;
  }
  __syncthreads();

  const int tx = tid % 16, ty = tid / 16;
  int acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0;
  const int lk = tid % BK, lr = tid / BK;

  for (int k0 = 0; k0 < g.K; k0 += BK) {
    int k = k0 + lk;
    bool kv = k < g.K;
    int kk = kv ? k : 0;
    int c = kk % g.C, rs = kk / g.C;
    int ds = (rs % g.S) * g.dw, dr = (rs / g.S) * g.dh;
#pragma unroll
    for (int i = 0; i < BM / 16; ++i) {
      int row = lr + 16 * i;
      RowI r = rows[row];
      int v = 0;
      int hi_ = r.h0 + dr, wi_ = r.w0 + ds;
      // zero padding pads with the zero point: (x - z_x) = 0 outside the image
      if (kv && r.valid && hi_ >= 0 && hi_ < g.H && wi_ >= 0 && wi_ < g.W)
        v = (int)xs[r.base + ((int64_t)hi_ * g.W + wi_) * g.C + c] - z_x;
      As[lk][row] = v;
    }
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) {
      int col = lr + 16 * i;
      int n = n0 + col;
      Bs[lk][col] = (kv && n < g.N) ? (int)ws[(int64_t)n * g.K + k] - z_w : 0;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < BK; ++q) {
      int a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[q][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[q][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * b[j];
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      int64_t o = ((int64_t)z * g.M + m) * g.N + n;
      float xf = (float)acc[i][j];
      if (bias) xf = __fadd_rn(xf, __fdiv_rn(bias[n], act_times_w));
      int q = (int)rintf(__fmul_rn(xf, mult)) + z_out;
      out[o] = (uint8_t)max(lo, min(hi, q));
      if (acc_dump) acc_dump[o] = acc[i][j];
    }
  }
}

}  // namespace

int qbn_umma_i8_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const uint8_t* x, int z_x, const int8_t* w,
                         int w_shared, int z_w, const float* bias, float act_times_w, float mult, int z_out, int lo, int hi,
                         uint8_t* out, int32_t* acc_dump, cudaStream_t st);  // umma_conv.cu; returns QBN_ERR_UNSUPPORTED if ragged

extern "C" int qbn_i8_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const uint8_t* x, float s_x, int32_t z_x,
                               const int8_t* w, int w_shared, float s_w, int32_t z_w, const float* bias, float s_out,
                               int32_t z_out, int relu, int act_min, int act_max, uint8_t* out, int32_t* acc_dump,
                               int path, void* stream) {
  QBN_CHECK_ARG(d && x && w && out, "null pointer");
  QBN_CHECK_ARG(d->B > 0 && d->H > 0 && d->W > 0 && d->C > 0 && d->N > 0 && d->R > 0 && d->S > 0 && d->Ho > 0 && d->Wo > 0, "sizes");
  QBN_CHECK_ARG(n_samples > 0 && n_samples <= 65535, "0 < n_samples <= 65535");
  QBN_CHECK_ARG(s_x > 0 && s_w > 0 && s_out > 0, "scales must be > 0");
  G8 g;
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C; g.N = d->N; g.R = d->R; g.S = d->S;
  g.sh = d->stride_h; g.sw = d->stride_w; g.ph = d->pad_h; g.pw = d->pad_w; g.dh = d->dil_h; g.dw = d->dil_w;
  g.Ho = d->Ho; g.Wo = d->Wo; g.K = d->R * d->S * d->C; g.M = (int64_t)d->B * d->Ho * d->Wo;
  // ATen qlinear/qconv (fbgemm): act_times_w = s_x*s_w ; multiplier = act_times_w / s_out, all fp32
  float atw = s_x * s_w;
  float mult = atw / s_out;
  int lo = relu ? z_out : 0;
  if (lo < act_min) lo = act_min;
  int hi = act_max < 255 ? act_max : 255;
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(path == QBN_I8_AUTO || path == QBN_I8_IMAD || path == QBN_I8_UMMA, "path");
  // the tcgen05 kernel feeds (x - z_x) as s8, which needs 7-bit activations (quant_utils.py:120)
  const bool umma_ok = d->C % 8 == 0 && z_x >= 0 && z_x <= 127 && hi <= 127 && d->N + 1 <= 256;
  if (path == QBN_I8_UMMA || (path == QBN_I8_AUTO && umma_ok))
    return qbn_umma_i8_conv_fwd(d, n_samples, x_shared, x, z_x, w, w_shared, z_w, bias, atw, mult, z_out, lo, hi, out, acc_dump, st);
  dim3 grid((unsigned)ceil_div64(g.M, BM), (unsigned)ceil_div64(g.N, BN), (unsigned)n_samples);
  i8_conv_kernel<<<grid, NT, 0, st>>>(g, x, x_shared, z_x, w, w_shared, z_w, bias, atw, mult, z_out, lo, hi, out, acc_dump);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
