// Staging for the planar LRT training kernels (umma_conv_p4.cu KIND_LRT, umma_wgrad_p4.cu): the module boundary is dense NHWC
// (ops.LRTFunction), the tcgen05 kernels read planar-C4 zero-bordered maps (p4_layout.cuh).
//   * qbn_p4_stage_input : x NHWC -> planar x and x^2 (TF32-rounded, RNA), phase-split for a stride-2 layer (SURVEY 8a A1/A2:
//                          the two operands of linear.py:32-35 / conv.py:24-27)
//   * qbn_p4_stage_grad  : g NHWC (+ std, eps) -> planar g and dv = g * eps / (2 std)   (d out / d var)
//   * qbn_lrt_p4_weight_prep : OIHW (mu, rho) -> the blocked [mu | sigma^2] operand of the forward, of the flipped / transposed
//                          input-gradient convolution, or of one phase of a stride-2 layer's input gradient
// Every kernel writes its whole output (borders, padding channels and the zero tail included): buffers come from torch.empty.
#include <string.h>
#include "common.cuh"
#include "p4_layout.cuh"

namespace {

struct StageGeom {
  int H, W, C, chunks;               // NHWC tensor: [n_img][H][W][C]; chunks = planes written (C_pad / 4)
  int Hp, Wp, bh, bw, split;         // planar maps
  uint32_t map_rows, body_rows;      // Hp * Wp ; phases * n_img * Hp * Wp
  uint32_t n_img_rows;               // n_img * Hp * Wp (rows of one phase)
  long long plane_rows;
};

// row of a chunk plane -> NHWC pixel index, or -1 for border / tail rows
__device__ __forceinline__ long long stage_pixel(const StageGeom& g, uint32_t r) {
  if (r >= g.body_rows) return -1;
  uint32_t ph = 0;
  if (g.split) { ph = r / g.n_img_rows; r -= ph * g.n_img_rows; }
  const uint32_t img = r / g.map_rows;
  const uint32_t rem = r - img * g.map_rows;
  const int hh = (int)(rem / (uint32_t)g.Wp), ww = (int)(rem - (uint32_t)hh * (uint32_t)g.Wp);
  if (hh < g.bh || ww < g.bw) return -1;
  int h = hh - g.bh, w = ww - g.bw;
  if (g.split) { h = 2 * h + (int)(ph >> 1); w = 2 * w + (int)(ph & 1); }
  return ((long long)img * g.H + h) * g.W + w;
}

__device__ __forceinline__ float4 ld4_guard(const float* base, long long pix, int C, int c0) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  const float* p = base + pix * C + c0;
  if ((C & 3) == 0) {
    if (c0 < C) v = *reinterpret_cast<const float4*>(p);
  } else {
    if (c0 + 0 < C) v.x = p[0];
    if (c0 + 1 < C) v.y = p[1];
    if (c0 + 2 < C) v.z = p[2];
    if (c0 + 3 < C) v.w = p[3];
  }
  return v;
}

__global__ void p4_stage_input_kernel(const float* __restrict__ x, StageGeom g, float4* __restrict__ xp, float4* __restrict__ xsq) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent planar conv launch (qbn_set_pdl) may start its prologue now
  const long long total = (long long)g.chunks * g.plane_rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / g.plane_rows);
    const uint32_t r = (uint32_t)(i - (long long)j * g.plane_rows);
    const long long pix = stage_pixel(g, r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), q = v;
    if (pix >= 0) {
      v = ld4_guard(x, pix, g.C, 4 * j);
      q = make_float4(tf32_round(__fmul_rn(v.x, v.x)), tf32_round(__fmul_rn(v.y, v.y)), tf32_round(__fmul_rn(v.z, v.z)), tf32_round(__fmul_rn(v.w, v.w)));
      v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
    }
    xp[i] = v;
    if (xsq) xsq[i] = q;
  }
}

__global__ void p4_stage_grad_kernel(const float* __restrict__ gout, const float* __restrict__ sd, const float* __restrict__ eps, StageGeom g,
                                     uint64_t seed, uint32_t sa, uint32_t sb, const uint32_t* __restrict__ sbase, float4* __restrict__ gp,
                                     float4* __restrict__ dvp) {
  if (sbase) sb += *sbase;           // the forward's draw (device-side offset, qbn_set_sample_base)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent planar conv launch (qbn_set_pdl) may start its prologue now
  const long long total = (long long)g.chunks * g.plane_rows;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int j = (int)(i / g.plane_rows);
    const uint32_t r = (uint32_t)(i - (long long)j * g.plane_rows);
    const long long pix = stage_pixel(g, r);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), d = v;
    if (pix >= 0) {
      const long long off = pix * g.C + 4 * j;
      v = *reinterpret_cast<const float4*>(gout + off);
      const float4 s = *reinterpret_cast<const float4*>(sd + off);
      float e[4];
      if (eps) {
        const float4 t = *reinterpret_cast<const float4*>(eps + off);
        e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
      } else {
        philox_normal4(seed, sa, sb, (uint64_t)(off >> 2), e);      // the forward's draw: counter = offset in out / 4
      }
      // lrt_dv_kernel's expression (gemm_fp32.cu): g * eps / (2 std)
      d = make_float4(tf32_round(v.x * e[0] / (2.0f * s.x)), tf32_round(v.y * e[1] / (2.0f * s.y)), tf32_round(v.z * e[2] / (2.0f * s.z)),
                      tf32_round(v.w * e[3] / (2.0f * s.w)));
      v = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
    }
    gp[i] = v;
    dvp[i] = d;
  }
}

// ---- fused staging: ONE pass over the NHWC tensor writes both layouts the planar training kernels read — planar C4 (forward /
// input gradient, K-major operands) and W32 (weight gradients, MN-major operands: 128-byte rows of 32 channels, the 32-byte chunks of
// row r XORed with r & 3, see umma_wgrad_p4.cu).  One thread = one 16-byte piece of a W32 row: a warp writes four whole W32 rows (512
// contiguous bytes), 64 contiguous bytes into each of 8 chunk planes, and reads four pixels' channels.
template <bool GRAD>
__global__ void lrt_stage_fused_kernel(const float* __restrict__ src, const float* __restrict__ sd, const float* __restrict__ eps, StageGeom g,
                                       uint64_t seed, uint32_t sa, uint32_t sb, const uint32_t* __restrict__ sbase, float4* __restrict__ a_p4,
                                       float4* __restrict__ b_p4, float4* __restrict__ a_w32, float4* __restrict__ b_w32,
                                       float4* __restrict__ noise_out, long long noise_n4) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent planar conv launch (qbn_set_pdl) may start its prologue now
  if (sbase) sb += *sbase;
  const int n_blk = (g.chunks + 7) / 8;
  const long long total = (long long)n_blk * g.plane_rows * 8;
  // forward staging with the layer's noise tensor in the same launch (qbn_lrt_stage_input_noise): the Philox / Box-Muller fill is
  // ALU work, the staging is memory traffic — odd blocks draw first and stage second, so both kinds are in flight on every SM
  auto draw = [&]() {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < noise_n4; i += (long long)gridDim.x * blockDim.x) {
      float z[4];
      philox_normal4(seed, sa, sb, (uint64_t)i, z);
      noise_out[i] = make_float4(z[0], z[1], z[2], z[3]);
    }
  };
  const bool with_noise = !GRAD && noise_out != nullptr;
  if (with_noise && (blockIdx.x & 1)) draw();
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int piece = (int)(i & 7);
    const long long br = i >> 3;
    const int blk = (int)(br / g.plane_rows);
    const uint32_t r = (uint32_t)(br - (long long)blk * g.plane_rows);
    const int j = blk * 8 + (((piece >> 1) ^ (int)(r & 3)) << 1) + (piece & 1);      // the 4-channel chunk stored at this position of row r
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
    if (j < g.chunks) {
      const long long pix = stage_pixel(g, r);
      if (pix >= 0) {
        if (!GRAD) {
          const float4 v = ld4_guard(src, pix, g.C, 4 * j);
          b = make_float4(tf32_round(__fmul_rn(v.x, v.x)), tf32_round(__fmul_rn(v.y, v.y)), tf32_round(__fmul_rn(v.z, v.z)), tf32_round(__fmul_rn(v.w, v.w)));
          a = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        } else {
          const long long off = pix * g.C + 4 * j;
          const float4 v = *reinterpret_cast<const float4*>(src + off);
          const float4 s = *reinterpret_cast<const float4*>(sd + off);
          float e[4];
          if (eps) {
            const float4 t = *reinterpret_cast<const float4*>(eps + off);
            e[0] = t.x; e[1] = t.y; e[2] = t.z; e[3] = t.w;
          } else {
            philox_normal4(seed, sa, sb, (uint64_t)(off >> 2), e);
          }
          b = make_float4(tf32_round(v.x * e[0] / (2.0f * s.x)), tf32_round(v.y * e[1] / (2.0f * s.y)), tf32_round(v.z * e[2] / (2.0f * s.z)),
                          tf32_round(v.w * e[3] / (2.0f * s.w)));
          a = make_float4(tf32_round(v.x), tf32_round(v.y), tf32_round(v.z), tf32_round(v.w));
        }
      }
      const long long pdst = (long long)j * g.plane_rows + r;
      a_p4[pdst] = a;
      b_p4[pdst] = b;
    }
    a_w32[i] = a;
    b_w32[i] = b;
  }
  if (with_noise && !(blockIdx.x & 1)) draw();
}

static int stage_geom(StageGeom& g, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int split, long long plane_rows) {
  memset(&g, 0, sizeof(g));
  g.H = H; g.W = W; g.C = C; g.chunks = C_pad / 4; g.bh = bh; g.bw = bw; g.split = split ? 1 : 0;
  if (split) {
    if ((H & 1) || (W & 1) || bh != 1 || bw != 1) { qbn_set_error("phase-split staging needs even H, W and a (1, 1) border"); return QBN_ERR_INVALID_ARG; }
    g.Hp = H / 2 + 1; g.Wp = W / 2 + 1;
  } else {
    g.Hp = H + bh; g.Wp = W + bw;
  }
  const long long body = (long long)(split ? 4 : 1) * n_img * g.Hp * g.Wp;
  if (body >= (1ll << 31) || plane_rows < body) { qbn_set_error("staging: plane of %lld rows for %lld map rows", plane_rows, body); return QBN_ERR_INVALID_ARG; }
  g.map_rows = (uint32_t)(g.Hp * g.Wp); g.n_img_rows = (uint32_t)(n_img * g.Hp * g.Wp); g.body_rows = (uint32_t)body;
  g.plane_rows = plane_rows;
  return QBN_OK;
}

}  // namespace

extern "C" int qbn_p4_stage_input(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                                  long long plane_rows, float* x_p4, float* xsq_p4, void* stream) {
  QBN_CHECK_ARG(x && x_p4, "null pointer");
  QBN_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && C > 0 && C_pad >= C && C_pad % 4 == 0 && bh >= 0 && bw >= 0, "sizes");
  StageGeom g;
  int rc = stage_geom(g, n_img, H, W, C, C_pad, bh, bw, phase_split, plane_rows);
  if (rc != QBN_OK) return rc;
  const long long total = (long long)g.chunks * plane_rows;
  p4_stage_input_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(x, g, reinterpret_cast<float4*>(x_p4), reinterpret_cast<float4*>(xsq_p4));
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_p4_stage_grad(const float* g_out, const float* std_saved, const float* eps, uint64_t seed, uint32_t stream_a,
                                 uint32_t stream_b, int64_t n_img, int H, int W, int N, int bh, int bw, long long plane_rows, float* g_p4,
                                 float* dv_p4, void* stream) {
  QBN_CHECK_ARG(g_out && std_saved && g_p4 && dv_p4, "null pointer");
  QBN_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && N > 0 && N % 4 == 0 && bh >= 0 && bw >= 0, "sizes (N % 4 == 0)");
  StageGeom g;
  int rc = stage_geom(g, n_img, H, W, N, N, bh, bw, 0, plane_rows);
  if (rc != QBN_OK) return rc;
  const long long total = (long long)g.chunks * plane_rows;
  p4_stage_grad_kernel<<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(g_out, std_saved, eps, g, seed, stream_a, stream_b, qbn_sample_base_ptr(),
                                                                                  reinterpret_cast<float4*>(g_p4), reinterpret_cast<float4*>(dv_p4));
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

// The two staging entry points with the W32 copies produced in the same pass (ops.LRTFunction uses these; the separate
// qbn_p4_stage_* + qbn_w32_from_p4 pair gives identical bytes).  *_w32: [ceil(C_pad/32)][plane_rows][32] floats.
extern "C" int qbn_lrt_stage_input(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                                   long long plane_rows, float* x_p4, float* xsq_p4, float* x_w32, float* xsq_w32, void* stream) {
  QBN_CHECK_ARG(x && x_p4 && xsq_p4 && x_w32 && xsq_w32, "null pointer");
  QBN_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && C > 0 && C_pad >= C && C_pad % 4 == 0 && bh >= 0 && bw >= 0, "sizes");
  StageGeom g;
  int rc = stage_geom(g, n_img, H, W, C, C_pad, bh, bw, phase_split, plane_rows);
  if (rc != QBN_OK) return rc;
  const long long total = (long long)((g.chunks + 7) / 8) * plane_rows * 8;
  lrt_stage_fused_kernel<false><<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      x, nullptr, nullptr, g, 0, 0, 0, nullptr, reinterpret_cast<float4*>(x_p4), reinterpret_cast<float4*>(xsq_p4), reinterpret_cast<float4*>(x_w32),
      reinterpret_cast<float4*>(xsq_w32), nullptr, 0);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
// qbn_lrt_stage_input and qbn_lrt_noise (core.cu: the same values, counter = element / 4 of the layer's NHWC output) in one launch
extern "C" int qbn_lrt_stage_input_noise(const float* x, int64_t n_img, int H, int W, int C, int C_pad, int bh, int bw, int phase_split,
                                         long long plane_rows, float* x_p4, float* xsq_p4, float* x_w32, float* xsq_w32, float* noise_out,
                                         int64_t noise_n, uint64_t seed, uint32_t stream_a, uint32_t stream_b, void* stream) {
  QBN_CHECK_ARG(x && x_p4 && xsq_p4 && x_w32 && xsq_w32 && noise_out, "null pointer");
  QBN_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && C > 0 && C_pad >= C && C_pad % 4 == 0 && bh >= 0 && bw >= 0, "sizes");
  QBN_CHECK_ARG(noise_n > 0 && noise_n % 4 == 0, "noise_n: a positive multiple of 4");
  StageGeom g;
  int rc = stage_geom(g, n_img, H, W, C, C_pad, bh, bw, phase_split, plane_rows);
  if (rc != QBN_OK) return rc;
  const long long total = (long long)((g.chunks + 7) / 8) * plane_rows * 8;
  const long long work = total > noise_n / 4 ? total : noise_n / 4;
  lrt_stage_fused_kernel<false><<<qbn_grid_for(work, 256), 256, 0, (cudaStream_t)stream>>>(
      x, nullptr, nullptr, g, seed, stream_a, stream_b, qbn_sample_base_ptr(), reinterpret_cast<float4*>(x_p4), reinterpret_cast<float4*>(xsq_p4),
      reinterpret_cast<float4*>(x_w32), reinterpret_cast<float4*>(xsq_w32), reinterpret_cast<float4*>(noise_out), noise_n / 4);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
extern "C" int qbn_lrt_stage_grad(const float* g_out, const float* std_saved, const float* eps, uint64_t seed, uint32_t stream_a,
                                  uint32_t stream_b, int64_t n_img, int H, int W, int N, int bh, int bw, long long plane_rows, float* g_p4,
                                  float* dv_p4, float* g_w32, float* dv_w32, void* stream) {
  QBN_CHECK_ARG(g_out && std_saved && g_p4 && dv_p4 && g_w32 && dv_w32, "null pointer");
  QBN_CHECK_ARG(n_img > 0 && H > 0 && W > 0 && N > 0 && N % 4 == 0 && bh >= 0 && bw >= 0, "sizes (N % 4 == 0)");
  StageGeom g;
  int rc = stage_geom(g, n_img, H, W, N, N, bh, bw, 0, plane_rows);
  if (rc != QBN_OK) return rc;
  const long long total = (long long)((g.chunks + 7) / 8) * plane_rows * 8;
  lrt_stage_fused_kernel<true><<<qbn_grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      g_out, std_saved, eps, g, seed, stream_a, stream_b, qbn_sample_base_ptr(), reinterpret_cast<float4*>(g_p4), reinterpret_cast<float4*>(dv_p4),
      reinterpret_cast<float4*>(g_w32), reinterpret_cast<float4*>(dv_w32), nullptr, 0);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

namespace {

struct WPrep {
  P4Block g;                  // geometry of the blocked OUTPUT
  int N, C, taps;             // the layer's parameters: OIHW [N][C][taps]
  int transposed;             // output row n' = input channel c, output channel c' = output channel n (input gradient)
  int tap_map[25];            // output tap -> parameter tap (-1: a zero tap)
  int n_sets;                 // mode 3: four tap maps (tap_map[4 * z + t]), one blocked [mu | sigma^2] pair per phase z
  int second_is_sigma;
};

// one thread = one float4 of the blocked tensor, for both halves ([mu | sigma^2]) of every tap set
__global__ void lrt_p4_weight_prep_kernel(const float* __restrict__ mu, const float* __restrict__ second, WPrep w, float4* __restrict__ out) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");      // a dependent planar conv launch (qbn_set_pdl) may start its prologue now
  const int n_sets = w.n_sets > 0 ? w.n_sets : 1;
  for (int64_t ii = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; ii < w.g.total4 * n_sets; ii += (int64_t)gridDim.x * blockDim.x) {
    const int set = (int)(ii / w.g.total4);
    const int64_t i = ii - (int64_t)set * w.g.total4;
    const int64_t idx = p4_canonical(w.g, i);            // n' * K + t' * C' + c'   (c' a multiple of 4)
    float m[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (idx >= 0) {
      const int np = (int)(idx / w.g.K);
      const int rem = (int)(idx - (int64_t)np * w.g.K);
      const int tp = rem / w.g.C, cp = rem - tp * w.g.C;
      const int t = w.tap_map[set * w.g.taps + tp];
      if (t >= 0) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int n = w.transposed ? cp + k : np, c = w.transposed ? np : cp + k;
          if (n < w.N && c < w.C) {
            const int64_t src = ((int64_t)n * w.C + c) * w.taps + t;
            const float sg = w.second_is_sigma ? second[src] : softplus_f(second[src]);
            m[k] = tf32_round(mu[src]);
            s2[k] = tf32_round(__fmul_rn(sg, sg));
          }
        }
      }
    }
    float4* o = out + (int64_t)set * 2 * w.g.total4;
    o[i] = make_float4(m[0], m[1], m[2], m[3]);
    o[w.g.total4 + i] = make_float4(s2[0], s2[1], s2[2], s2[3]);
  }
}

}  // namespace

// mode 0: forward operand, output channels N, input channels C_pad (channels >= C are zero), taps R*S, blocked for `stride`.
// mode 1: input gradient of a stride-1 layer: rows = the layer's input channels, K = its output channels, taps reversed.
// mode 2: one phase of a stride-2 layer's input gradient: like mode 1 with the taps `tap_list[0..n_taps)` (parameter tap indices).
// out: 2 x qbn_p4_weight_floats(...) floats — the mu blocks, then the sigma^2 blocks (sigma = softplus(rho) or `second` itself).
extern "C" int qbn_lrt_p4_weight_prep(const float* mu, const float* second, int second_is_sigma, int N, int C, int C_pad, int R, int S,
                                      int stride, int mode, const int* tap_list, int n_taps, float* out, long long* out_floats, void* stream) {
  QBN_CHECK_ARG(mu && second && out, "null pointer");
  QBN_CHECK_ARG(N > 0 && C > 0 && R > 0 && S > 0 && R * S <= 25 && mode >= 0 && mode <= 3, "sizes / mode");
  WPrep w;
  memset(&w, 0, sizeof(w));
  w.N = N; w.C = C; w.taps = R * S; w.second_is_sigma = second_is_sigma;
  bool ok;
  if (mode == 0) {
    QBN_CHECK_ARG(C_pad >= C, "C_pad");
    ok = p4_block_geom(N, C_pad, R * S, stride, w.g);
    for (int t = 0; t < R * S; ++t) w.tap_map[t] = t;
  } else if (mode == 1) {
    w.transposed = 1;
    ok = p4_block_geom(C, N, R * S, 1, w.g);
    for (int t = 0; t < R * S; ++t) w.tap_map[t] = R * S - 1 - t;
  } else if (mode == 3) {
    // the four phases (a, b) of a 3x3 stride-2 layer's input gradient, four taps (dr, ds) each: tap (dr, ds) of phase (a, b) exists
    // iff (dr == 0 or a == 1) and (ds == 0 or b == 1) and is the parameter tap (dr ? 0 : (a ? 2 : 1), ds ? 0 : (b ? 2 : 1))
    QBN_CHECK_ARG(R == 3 && S == 3 && stride == 2, "mode 3: 3x3 stride-2 layers");
    w.transposed = 1;
    w.n_sets = 4;
    ok = p4_block_geom(C, N, 4, 1, w.g);
    for (int z = 0; z < 4; ++z)
      for (int t = 0; t < 4; ++t) {
        const int a = z >> 1, b = z & 1, dr = t >> 1, ds = t & 1;
        const bool valid = (dr == 0 || a == 1) && (ds == 0 || b == 1);
        w.tap_map[4 * z + t] = valid ? (dr ? 0 : (a ? 2 : 1)) * 3 + (ds ? 0 : (b ? 2 : 1)) : -1;
      }
  } else {
    QBN_CHECK_ARG(tap_list && n_taps > 0 && n_taps <= 4, "tap list");
    w.transposed = 1;
    ok = p4_block_geom(C, N, n_taps, 1, w.g);
    for (int t = 0; t < n_taps; ++t) {
      QBN_CHECK_ARG(tap_list[t] >= 0 && tap_list[t] < R * S, "tap index");
      w.tap_map[t] = tap_list[t];
    }
  }
  if (!ok) {
    qbn_set_error("qbn_lrt_p4_weight_prep: no blocking (mode %d N=%d C=%d C_pad=%d)", mode, N, C, C_pad);
    return QBN_ERR_UNSUPPORTED;
  }
  const int n_sets = w.n_sets > 0 ? w.n_sets : 1;
  if (out_floats) *out_floats = 2 * w.g.total4 * 4 * n_sets;
  lrt_p4_weight_prep_kernel<<<qbn_grid_for(w.g.total4 * n_sets, 256), 256, 0, (cudaStream_t)stream>>>(mu, second, w, reinterpret_cast<float4*>(out));
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
