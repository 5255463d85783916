// Layout helpers of the int8 planar path (p4_layout.cuh, "planar C16"): sampled-weight blocking, entry/exit of the layout.
// The convolution itself is umma_conv_p4.cu's kernel instantiated for kind::i8 (qbn_i8_conv_p16_fwd).
#include "common.cuh"
#include "p4_layout.cuh"

namespace {

// ---- blocked weights: [sample][channel block][tap][chunk j][n_pad rows][16 s8], one thread per 16-byte row ------------------
// Source: the sampler's output in the reference's own order, [sample][N][C][taps] (OIHW, conv_q.py:113-119), so the Philox
// stream, the vector-body / tail split of quantized::add and therefore every sampled integer are those of qbn_i8_sample_weights.
__global__ void i8_p16_block_kernel(const int8_t* __restrict__ w, int N, int C, int C_pad, int taps, int CB, int n_pad, int64_t rows_per_sample,
                                    int8_t* __restrict__ out) {
  const int s = blockIdx.y;
  const int8_t* ws = w + (int64_t)s * N * C * taps;
  uint4* os = reinterpret_cast<uint4*>(out) + (int64_t)s * rows_per_sample;
  const int cbc = CB / 16;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < rows_per_sample; i += (int64_t)gridDim.x * blockDim.x) {
    const uint32_t u = (uint32_t)i;
    const uint32_t t1 = u / (uint32_t)n_pad, n = u - t1 * (uint32_t)n_pad;
    const uint32_t t2 = t1 / (uint32_t)cbc, j = t1 - t2 * (uint32_t)cbc;
    const uint32_t cb = t2 / (uint32_t)taps, t = t2 - cb * (uint32_t)taps;
    const int c0 = (int)(cb * CB + j * 16);
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if ((int)n <= N) {
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int c = c0 + k;
        int v = 0;
        if (c < C) v = (int)n < N ? (int)ws[((int64_t)n * C + c) * taps + t] : 1;      // row N: ones over the real channels
        pk[k >> 2] |= ((uint32_t)v & 0xFFu) << ((k & 3) * 8);
      }
    }
    os[i] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---- quint8 NHWC [n_img][H][W][C] -> planar C16 s8 (q - z) with the shared zero border (1 row on top, 1 column on the left) ----
__global__ void i8_p16_from_nhwc_kernel(const uint8_t* __restrict__ x, int64_t n_img, int H, int W, int C, int C_pad, int z, int64_t plane_rows,
                                        int8_t* __restrict__ out) {
  const int Hp = H + 1, Wp = W + 1;
  const int64_t rows = n_img * Hp * Wp;
  const int n_chunks = C_pad / 16;
  const int64_t total = rows * n_chunks;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t row = i % rows;
    const int ch = (int)(i / rows);
    const int wp = (int)(row % Wp);
    const int64_t t = row / Wp;
    const int hp = (int)(t % Hp);
    const int64_t b = t / Hp;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    if (hp >= 1 && wp >= 1) {
      const uint8_t* src = x + ((b * H + (hp - 1)) * W + (wp - 1)) * C;
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const int c = ch * 16 + k;
        const int v = c < C ? (int)src[c] - z : 0;
        pk[k >> 2] |= ((uint32_t)v & 0xFFu) << ((k & 3) * 8);
      }
    }
    reinterpret_cast<uint4*>(out)[(int64_t)ch * plane_rows + row] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
  }
}

// ---- planar C16 s8 -> quint8 NHWC interior (exit of the layout: tests, and the flatten in front of the linear layer) ----------
__global__ void i8_p16_to_nhwc_kernel(const int8_t* __restrict__ x, int64_t n_img, int H, int W, int C, int z, int64_t plane_rows,
                                      uint8_t* __restrict__ out) {
  const int Hp = H + 1, Wp = W + 1;
  const int64_t total = n_img * H * W * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int64_t b = t / H;
    const int64_t row = (b * Hp + h + 1) * Wp + w + 1;
    out[i] = (uint8_t)((int)x[((int64_t)(c >> 4) * plane_rows + row) * 16 + (c & 15)] + z);
  }
}

// ---- nn.AvgPool2d(H) over the whole map (models_bbb.py:209-211; ATen qavg_pool2d): the maps hold q - z, so the interior sum IS
// sum(q) - H*W*z; q_out = clamp(rint(fp32(acc) * fp32(1/(H*W))) + z, 0, 255), then clamp_activation.  Output quint8 [n_img][C].
__global__ void i8_p16_avgpool_kernel(const int8_t* __restrict__ x, int64_t n_img, int H, int W, int C, int z, int64_t plane_rows, float inv_area,
                                      int lo, int hi, uint8_t* __restrict__ out) {
  const int Hp = H + 1, Wp = W + 1;
  const int64_t total = n_img * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / C;
    const int8_t* src = x + ((int64_t)(c >> 4) * plane_rows + b * Hp * Wp) * 16 + (c & 15);
    int acc = 0;
    for (int h = 1; h < Hp; ++h)
      for (int w = 1; w < Wp; ++w) acc += (int)src[(int64_t)(h * Wp + w) * 16];
    const int q = (int)rintf(__fmul_rn((float)acc, inv_area)) + z;
    out[i] = (uint8_t)max(lo, min(hi, max(0, min(255, q))));
  }
}

}  // namespace

extern "C" int qbn_i8_p16_block_weights(const int8_t* w_oihw, int n_samples, int N, int C, int C_pad, int taps, int stride, int8_t* out,
                                        void* stream) {
  QBN_CHECK_ARG(w_oihw && out && n_samples > 0 && n_samples <= 65535 && N > 0 && C > 0 && taps > 0, "args");
  const int CB = qbn_p16_block_channels(C_pad, stride, taps);
  if (C_pad % 32 != 0 || C_pad < C || CB == 0 || N + 1 > 256) {
    qbn_set_error("qbn_i8_p16_block_weights: needs C_pad %% 32 == 0, C_pad >= C and N <= 255 (C=%d C_pad=%d N=%d)", C, C_pad, N);
    return QBN_ERR_UNSUPPORTED;
  }
  const int n_pad = qbn_p16_n_pad(N);
  const int64_t rows = (int64_t)(C_pad / CB) * taps * (CB / 16) * n_pad;
  int gx = (int)((rows + 255) / 256);
  int cap = (qbn_sm_count() * 8 + n_samples - 1) / n_samples;
  if (gx > cap) gx = cap < 1 ? 1 : cap;
  i8_p16_block_kernel<<<dim3(gx, n_samples), 256, 0, (cudaStream_t)stream>>>(w_oihw, N, C, C_pad, taps, CB, n_pad, rows, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_i8_p16_from_nhwc(const uint8_t* x, int64_t n_img, int H, int W, int C, int C_pad, int32_t z_x, int64_t plane_rows,
                                    int8_t* out, void* stream) {
  QBN_CHECK_ARG(x && out && n_img > 0 && H > 0 && W > 0 && C > 0 && C_pad % 16 == 0 && C_pad >= C, "args");
  QBN_CHECK_ARG(plane_rows >= n_img * (H + 1) * (W + 1), "plane too small");
  QBN_CHECK_ARG(z_x >= 0 && z_x <= 127, "zero point must fit 7 bits");
  const int64_t total = n_img * (H + 1) * (W + 1) * (C_pad / 16);
  i8_p16_from_nhwc_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, n_img, H, W, C, C_pad, z_x, plane_rows, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_i8_p16_to_nhwc(const int8_t* x, int64_t n_img, int H, int W, int C, int32_t z_x, int64_t plane_rows, uint8_t* out,
                                  void* stream) {
  QBN_CHECK_ARG(x && out && n_img > 0 && H > 0 && W > 0 && C > 0, "args");
  i8_p16_to_nhwc_kernel<<<qbn_grid_for(n_img * H * W * C, 256), 256, 0, (cudaStream_t)stream>>>(x, n_img, H, W, C, z_x, plane_rows, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_i8_p16_avgpool(const int8_t* x, int64_t n_img, int H, int W, int C, int32_t z_x, int64_t plane_rows, int lo, int hi,
                                  uint8_t* out, void* stream) {
  QBN_CHECK_ARG(x && out && n_img > 0 && H > 0 && W > 0 && C > 0, "args");
  QBN_CHECK_ARG(lo >= 0 && hi <= 255 && lo <= hi, "0<=lo<=hi<=255");
  i8_p16_avgpool_kernel<<<qbn_grid_for(n_img * C, 256), 256, 0, (cudaStream_t)stream>>>(x, n_img, H, W, C, z_x, plane_rows,
                                                                                       1.0f / (float)(H * W), lo, hi, out);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
