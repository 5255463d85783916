// Layout helpers of the int8 planar path (p4_layout.cuh, "planar C16"): sampled-weight blocking, entry/exit of the layout.
// The convolution itself is umma_conv_p4.cu's kernel instantiated for kind::i8 (qbn_i8_conv_p16_fwd).
#include <string.h>
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

// ---- int8 MC-Dropout on planar-C16 maps (dropout.py:31-39) for a chunk of Monte-Carlo samples, optionally followed by the
// BasicBlock's quantized::add[_relu] (models_mc.py:143-157: the dropout sits between the second conv and the residual add, so the
// add cannot ride the conv epilogue).  Per element: m_q = clamp(rint(mask / s_m) + z_m, 0, 255);
// q = clamp(rint(fp32((x - z_x)(m_q - z_m)) * fp32(s_x * s_m * (1 / s_m))) + z_m, lo, hi) at (s_m * multiplier, z_m) — the arithmetic
// of qbn_i8_dropout_mc on this layout; then, with a residual r at (s_r, z_r): clamp(rint((fma(s_a, q, -s_a z_m) + fma(s_r, r, -s_r z_r))
// * fp32(1 / s_add)) + z_add, add_lo, hi) (ATen's vector body, as in the conv epilogue).  mask: fp32 {0, 1} [n_samples * B][C]
// (qbn_dropout_masks_multi: the draws of qbn_i8_dropout_mc).  Maps hold q - zero point: a zero (border, tail, padding channel)
// stays zero through every step.  x and the residual are stored in the normal layout; the output in the normal layout or
// phase-split for a stride-2 consumer (like the conv epilogue's QBN_FLAG_OUT_PHASE_SPLIT).
// x_shared: the input holds B images shared by all samples (the first conv of a deterministic network).
struct P16Drop {
  const int8_t* x; long long x_plane; int x_shared;
  const int8_t* res; long long res_plane;
  int8_t* out; long long out_plane;
  const float* mask;
  long long rows;            // n_img * map_rows (input rows: x and residual are stored in the normal layout)
  int map_rows, n_img, B, C, chunks;
  int out_split, Wp, Hp2, Wp2;   // phase-split OUTPUT for a stride-2 consumer: pixel (h, w) -> map (h&1, w&1), position (h>>1, w>>1)
  long long q2_total;
  float inv_sm, mult; int z_m, lo, hi;
  int has_add; float s_a, p_a, s_b, p_b, inv_s_add; int z_res, z_add, add_lo, add_hi;
};
__global__ void i8_p16_dropout_kernel(const __grid_constant__ P16Drop p) {
  const long long total = (long long)p.chunks * p.rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int ch = (int)(i / p.rows);
    const long long row = i - (long long)ch * p.rows;
    const int img = (int)(row / p.map_rows);
    const int pix = (int)(row - (long long)img * p.map_rows);
    const long long src_row = p.x_shared ? (long long)(img % p.B) * p.map_rows + pix : row;
    long long orow = row;
    if (p.out_split) {
      const int hh = pix / p.Wp, ww = pix - hh * p.Wp;
      if (hh < 1 || ww < 1) continue;                          // the border of a phase-split map is never written (pre-zeroed)
      const int h = hh - 1, w = ww - 1;
      orow = (long long)((h & 1) * 2 + (w & 1)) * p.q2_total + ((long long)img * p.Hp2 + (h >> 1) + 1) * p.Wp2 + (w >> 1) + 1;
    }
    const uint4 xv = reinterpret_cast<const uint4*>(p.x)[(long long)ch * p.x_plane + src_row];
    uint4 rv = make_uint4(0, 0, 0, 0);
    if (p.has_add) rv = reinterpret_cast<const uint4*>(p.res)[(long long)ch * p.res_plane + row];
    const uint32_t xw[4] = {xv.x, xv.y, xv.z, xv.w}, rw[4] = {rv.x, rv.y, rv.z, rv.w};
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    const float* mrow = p.mask + (long long)img * p.C;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int c = ch * 16 + k;
      if (c < p.C) {
        const int xs = (int)(int8_t)(xw[k >> 2] >> ((k & 3) * 8));                  // q_x - z_x
        const int mq = max(0, min(255, (int)rintf(__fmul_rn(mrow[c], p.inv_sm)) + p.z_m));
        const int prod = xs * (mq - p.z_m);
        int q = max(p.lo, min(p.hi, (int)rintf(__fmul_rn((float)prod, p.mult)) + p.z_m));
        int st = q - p.z_m;
        if (p.has_add) {
          const int rb = (int)(int8_t)(rw[k >> 2] >> ((k & 3) * 8)) + p.z_res;
          const float da = __fmaf_rn(p.s_a, (float)q, p.p_a);
          const float db = __fmaf_rn(p.s_b, (float)rb, p.p_b);
          q = max(p.add_lo, min(p.add_hi, (int)rintf(__fmul_rn(__fadd_rn(da, db), p.inv_s_add)) + p.z_add));
          st = q - p.z_add;
        }
        pk[k >> 2] |= ((uint32_t)st & 0xFFu) << ((k & 3) * 8);
      }
    }
    reinterpret_cast<uint4*>(p.out)[(long long)ch * p.out_plane + orow] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
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

// int8 MC-Dropout of a chunk of samples on planar-C16 maps (+ optional residual add / ReLU); see the kernel comment.
// x: [C_pad/16][x_plane_rows][16] holding q - z_x at scale s_x, n_in = x_shared ? B : n_samples * B images of map_rows rows per
// (normal layout); out: n_samples * B images (normal or phase-split), holding q - z_m (no residual) or q - z_add.  rq (nullable): the residual add —
// s_res / z_res of `residual`, s_add / z_add / add_relu of the output; the dropped operand enters at scale s_drop_out = s_m * multiplier.
extern "C" int qbn_i8_p16_dropout(const int8_t* x, long long x_plane_rows, int x_shared, int n_samples, int B, int Hp, int Wp, int out_phase_split,
                                  int C, float s_x, const float* mask, float s_m, int32_t z_m, float s_drop_out, int act_max, const int8_t* residual,
                                  long long res_plane_rows, const qbn_i8_requant* rq, int8_t* out, long long out_plane_rows, void* stream) {
  QBN_CHECK_ARG(x && mask && out && n_samples > 0 && B > 0 && Hp > 1 && Wp > 1 && C > 0, "args");
  QBN_CHECK_ARG(s_x > 0 && s_m > 0 && s_drop_out > 0 && z_m >= 0 && z_m <= 127 && act_max >= 1 && act_max <= 127, "scales / zero point / 7-bit activations");
  QBN_CHECK_ARG(!residual || rq, "residual add parameters");
  P16Drop p;
  memset(&p, 0, sizeof(p));
  p.x = x; p.x_plane = x_plane_rows; p.x_shared = x_shared ? 1 : 0; p.out = out; p.out_plane = out_plane_rows; p.mask = mask;
  p.n_img = n_samples * B; p.B = B; p.map_rows = Hp * Wp; p.Wp = Wp; p.C = C; p.chunks = (C + 15) / 16;
  p.rows = (long long)p.n_img * p.map_rows;
  long long out_rows = p.rows;
  if (out_phase_split) {
    QBN_CHECK_ARG(((Hp - 1) % 2 == 0) && ((Wp - 1) % 2 == 0), "phase-split output needs even H, W");
    p.out_split = 1; p.Hp2 = (Hp - 1) / 2 + 1; p.Wp2 = (Wp - 1) / 2 + 1;
    p.q2_total = (long long)p.n_img * p.Hp2 * p.Wp2;
    out_rows = 4 * p.q2_total;
  }
  QBN_CHECK_ARG(out_plane_rows >= out_rows && x_plane_rows >= (long long)(x_shared ? B : p.n_img) * p.map_rows, "planes too small");
  p.inv_sm = 1.0f / s_m; p.mult = s_x * s_m * (1.0f / s_m); p.z_m = z_m; p.lo = 0; p.hi = act_max;
  if (residual) {
    QBN_CHECK_ARG(rq->s_res > 0 && rq->s_add > 0 && rq->z_res >= 0 && rq->z_res <= 127 && rq->z_add >= 0 && rq->z_add <= 127, "residual add parameters");
    QBN_CHECK_ARG(res_plane_rows >= p.rows, "residual plane too small");
    p.has_add = 1; p.res = residual; p.res_plane = res_plane_rows;
    p.s_a = s_drop_out; p.p_a = s_drop_out * (float)(-z_m);
    p.s_b = rq->s_res; p.p_b = rq->s_res * (float)(-rq->z_res);
    p.inv_s_add = 1.0f / rq->s_add; p.z_res = rq->z_res; p.z_add = rq->z_add;
    p.add_lo = rq->add_relu ? rq->z_add : 0; p.add_hi = act_max;
  }
  i8_p16_dropout_kernel<<<qbn_grid_for((long long)p.chunks * p.rows, 256), 256, 0, (cudaStream_t)stream>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
