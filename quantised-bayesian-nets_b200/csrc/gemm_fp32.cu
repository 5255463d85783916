// fp32 ("exact") mode of the contraction kernels: CUDA-core FFMA implicit GEMM.
//
// One generic tiled kernel, specialised by a Problem policy that says how a row of A, a column of
// B and an index of the reduction dimension map onto NHWC activations / OHWI weights:
//   FwdLRT   A1/A2 forward  : rows = output pixels, cols = out-channels, red = (r,s,c);
//                             second accumulator uses x^2 (formed in registers) and sigma^2.
//   FwdEval  A4 forward     : same geometry, one accumulator, per-MC-sample weights, fused
//                             dropout mask on the operand and affine/residual/ReLU epilogue.
//   Dgrad    A3 dx          : rows = input pixels, cols = in-channels, red = (r,s,n).
//   Wgrad    A3 dmu/dsigma2 : rows = out-channels, cols = (r,s,c), red = output pixels (split
//                             over blockIdx.z into a workspace, then reduced deterministically).
// This is the QBN_MATH_FP32 path (rtol 1e-5 parity).  The tensor-core path is umma_conv.cu.
#include "common.cuh"

namespace {

constexpr int BM_DEFAULT = 64;
constexpr int BK = 16;
constexpr int NT = 256;

struct Geom {
  int B, H, W, C, N, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo;
  int K;       // R*S*C
  int64_t M;   // B*Ho*Wo
};

static Geom make_geom(const qbn_conv_desc* d) {
  Geom g;
  g.B = d->B; g.H = d->H; g.W = d->W; g.C = d->C; g.N = d->N; g.R = d->R; g.S = d->S;
  g.sh = d->stride_h; g.sw = d->stride_w; g.ph = d->pad_h; g.pw = d->pad_w; g.dh = d->dil_h; g.dw = d->dil_w;
  g.Ho = d->Ho; g.Wo = d->Wo;
  g.K = d->R * d->S * d->C;
  g.M = (int64_t)d->B * d->Ho * d->Wo;
  return g;
}

static int check_desc(const qbn_conv_desc* d) {
  if (!d) return 0;
  if (d->B <= 0 || d->H <= 0 || d->W <= 0 || d->C <= 0 || d->N <= 0 || d->R <= 0 || d->S <= 0) return 0;
  if (d->stride_h <= 0 || d->stride_w <= 0 || d->dil_h <= 0 || d->dil_w <= 0 || d->pad_h < -8 || d->pad_w < -8) return 0;
  if (d->out_pad_h < 0 || d->out_pad_w < 0) return 0;
  int ho = (d->H + 2 * d->pad_h - d->dil_h * (d->R - 1) - 1) / d->stride_h + 1;
  int wo = (d->W + 2 * d->pad_w - d->dil_w * (d->S - 1) - 1) / d->stride_w + 1;
  return ho == d->Ho && wo == d->Wo && ho > 0 && wo > 0;
}

struct Info {  // generic per-row / per-col / per-reduction-index decoded coordinates
  int64_t base;
  int a, b, c, valid;
};

// ---------------------------------------------------------------------------------------------
// generic kernel
// ---------------------------------------------------------------------------------------------
template <class P>
__global__ void __launch_bounds__(NT) igemm_kernel(P p) {
  constexpr int BN = P::BN, BM = P::BM;
  constexpr int TN = BN / 16;  // 16x16 thread grid
  constexpr int TM = BM / 16;
  constexpr bool DUAL = P::DUAL;
  constexpr bool SEP_A2 = DUAL && !P::A2_SQUARE;
  constexpr bool SEP_B2 = DUAL && !P::B2_SQUARE;

  __shared__ __align__(16) float As1[BK][BM + 4];
  __shared__ __align__(16) float As2[SEP_A2 ? BK : 1][BM + 4];
  __shared__ __align__(16) float Bs1[BK][BN + 4];
  __shared__ __align__(16) float Bs2[SEP_B2 ? BK : 1][BN + 4];
  __shared__ Info rows[BM];
  __shared__ Info cols[BN];

  const int tid = threadIdx.x;
  const int z = blockIdx.z;
  const int64_t m0 = (int64_t)blockIdx.x * BM;
  const int n0 = blockIdx.y * BN;
  p.begin(z);

  for (int i = tid; i < BM; i += NT) rows[i] = p.row_info(m0 + i);
  for (int i = tid; i < BN; i += NT) cols[i] = p.col_info(n0 + i);
  __syncthreads();

  const int tx = tid % 16, ty = tid / 16;
  float acc1[TM][TN], acc2[DUAL ? TM : 1][DUAL ? TN : 1];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc1[i][j] = 0.f;
  if constexpr (DUAL) {
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
      for (int j = 0; j < TN; ++j) acc2[i][j] = 0.f;
  }

  const int lk = tid % BK;   // this thread's reduction offset inside a tile
  const int lr = tid / BK;   // 0..15
  const int64_t rbeg = p.red_begin(z), rend = p.red_end(z);

  // Software pipeline: the gather loads of reduction tile t+1 are issued into registers before tile t is multiplied,
  // so their latency hides behind the FMAs (one smem stage, two barriers per tile).
  float pa1[BM / 16], pa2[SEP_A2 ? BM / 16 : 1], pb1[BN / 16], pb2[SEP_B2 ? BN / 16 : 1];
  auto fetch = [&](int64_t r0) {
    const int64_t kred = r0 + lk;
    Info red = p.red_info(kred, kred < rend);
#pragma unroll
    for (int i = 0; i < BM / 16; ++i) {
      float a1 = 0.f, a2 = 0.f;
      p.load_a(rows[lr + 16 * i], red, a1, a2);
      pa1[i] = a1;
      if constexpr (SEP_A2) pa2[i] = a2;
    }
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) {
      float b1 = 0.f, b2 = 0.f;
      p.load_b(cols[lr + 16 * i], red, b1, b2);
      pb1[i] = b1;
      if constexpr (SEP_B2) pb2[i] = b2;
    }
  };
  if (rbeg < rend) fetch(rbeg);
  for (int64_t r0 = rbeg; r0 < rend; r0 += BK) {
#pragma unroll
    for (int i = 0; i < BM / 16; ++i) {
      As1[lk][lr + 16 * i] = pa1[i];
      if constexpr (SEP_A2) As2[lk][lr + 16 * i] = pa2[i];
    }
#pragma unroll
    for (int i = 0; i < BN / 16; ++i) {
      Bs1[lk][lr + 16 * i] = pb1[i];
      if constexpr (SEP_B2) Bs2[lk][lr + 16 * i] = pb2[i];
    }
    __syncthreads();
    if (r0 + BK < rend) fetch(r0 + BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[TM], b[TN], a2[TM], b2[TN];
#pragma unroll
      for (int i = 0; i < TM; ++i) a[i] = As1[k][ty * TM + i];
#pragma unroll
      for (int j = 0; j < TN; ++j) b[j] = Bs1[k][tx * TN + j];
      if constexpr (DUAL) {
#pragma unroll
        for (int i = 0; i < TM; ++i) {
          if constexpr (SEP_A2) a2[i] = As2[k][ty * TM + i]; else a2[i] = a[i] * a[i];
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          if constexpr (SEP_B2) b2[j] = Bs2[k][tx * TN + j]; else b2[j] = b[j] * b[j];
        }
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          acc1[i][j] = fmaf(a[i], b[j], acc1[i][j]);
          if constexpr (DUAL) acc2[i][j] = fmaf(a2[i], b2[j], acc2[i][j]);
        }
    }
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < TM; ++i) {
    int64_t m = m0 + ty * TM + i;
    if (m >= p.rows_total()) continue;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      int n = n0 + tx * TN + j;
      if (n >= p.cols_total()) continue;
      float second = 0.f;
      if constexpr (DUAL) second = acc2[i][j];
      p.store(z, m, n, acc1[i][j], second);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// shared geometry helpers
// ---------------------------------------------------------------------------------------------
struct PixelRows {  // rows = output pixels of the conv
  Geom g;
  __device__ Info pixel_row(int64_t m) const {
    Info r;
    r.valid = m < g.M;
    int64_t mm = r.valid ? m : 0;
    int wo = (int)(mm % g.Wo);
    int64_t t = mm / g.Wo;
    int ho = (int)(t % g.Ho);
    int b = (int)(t / g.Ho);
    r.a = ho * g.sh - g.ph;  // h0
    r.b = wo * g.sw - g.pw;  // w0
    r.c = b;
    r.base = (int64_t)b * g.H * g.W * g.C;
    return r;
  }
  __device__ Info tap_red(int64_t k, bool ok) const {  // k -> (r, s, c)
    Info q;
    q.valid = ok && k < g.K;
    int kk = q.valid ? (int)k : 0;
    q.c = kk % g.C;
    int rs = kk / g.C;
    q.b = (rs % g.S) * g.dw;
    q.a = (rs / g.S) * g.dh;
    q.base = kk;
    return q;
  }
  // offset of input element for (row, tap) or -1
  __device__ int64_t in_offset(const Info& row, const Info& red) const {
    if (!(row.valid && red.valid)) return -1;
    int hi = row.a + red.a, wi = row.b + red.b;
    if (hi < 0 || hi >= g.H || wi < 0 || wi >= g.W) return -1;
    return row.base + ((int64_t)hi * g.W + wi) * g.C + red.c;
  }
};

// ---- A1/A2 forward ------------------------------------------------------------------------
template <int BN_>
struct FwdLRT : PixelRows {
  static constexpr int BN = BN_, BM = BM_DEFAULT;
  static constexpr bool DUAL = true, A2_SQUARE = true, B2_SQUARE = false;
  const float* x; const float* mu; const float* sig2; const float* bias; const float* eps;
  float* out; float* std_out;
  uint64_t seed; uint32_t sa, sb; const uint32_t* sbase;      // sbase: device-side draw offset (qbn_set_sample_base)
  __device__ void begin(int) {}
  __device__ int64_t rows_total() const { return g.M; }
  __device__ int cols_total() const { return g.N; }
  __device__ int64_t red_begin(int) const { return 0; }
  __device__ int64_t red_end(int) const { return g.K; }
  __device__ Info row_info(int64_t m) const { return pixel_row(m); }
  __device__ Info col_info(int n) const { Info c; c.valid = n < g.N; c.base = (int64_t)(c.valid ? n : 0) * g.K; c.a = c.b = c.c = 0; return c; }
  __device__ Info red_info(int64_t k, bool ok) const { return tap_red(k, ok); }
  __device__ void load_a(const Info& row, const Info& red, float& a1, float&) const {
    int64_t o = in_offset(row, red);
    a1 = o >= 0 ? x[o] : 0.f;
  }
  __device__ void load_b(const Info& col, const Info& red, float& b1, float& b2) const {
    if (col.valid && red.valid) { b1 = mu[col.base + red.base]; b2 = sig2[col.base + red.base]; }
  }
  __device__ void store(int, int64_t m, int n, float mean, float var) const {
    int64_t o = m * g.N + n;
    float sd = sqrtf(1e-8f + var);
    float e = eps ? eps[o] : philox_normal1(seed, sa, sb + (sbase ? *sbase : 0u), (uint64_t)o);
    // linear.py:40 / conv.py:31-32: mean + std*noise (+ bias)
    float v = __fadd_rn(mean, __fmul_rn(sd, e));
    if (bias) v = __fadd_rn(v, bias[n]);
    out[o] = v;
    if (std_out) std_out[o] = sd;
  }
};

// ---- A4 forward (per-sample weights) ------------------------------------------------------------
template <int BN_>
struct FwdEval : PixelRows {
  static constexpr int BN = BN_, BM = BM_DEFAULT;
  static constexpr bool DUAL = false, A2_SQUARE = false, B2_SQUARE = false;
  const float* x; const float* w; const float* scale; const float* shift; const float* residual; const float* in_mask;
  float* out;
  float in_mult;
  int x_shared, w_shared, flags;
  int64_t x_sample_stride, w_sample_stride;
  const float* xs; const float* ws; const float* ms; int zs;
  __device__ void begin(int z) {
    zs = z;
    xs = x + (x_shared ? 0 : (int64_t)z * x_sample_stride);
    ws = w + (w_shared ? 0 : (int64_t)z * w_sample_stride);
    ms = in_mask ? in_mask + (int64_t)z * g.B * g.C : nullptr;
  }
  __device__ int64_t rows_total() const { return g.M; }
  __device__ int cols_total() const { return g.N; }
  __device__ int64_t red_begin(int) const { return 0; }
  __device__ int64_t red_end(int) const { return g.K; }
  __device__ Info row_info(int64_t m) const { return pixel_row(m); }
  __device__ Info col_info(int n) const { Info c; c.valid = n < g.N; c.base = (int64_t)(c.valid ? n : 0) * g.K; c.a = c.b = c.c = 0; return c; }
  __device__ Info red_info(int64_t k, bool ok) const { return tap_red(k, ok); }
  __device__ void load_a(const Info& row, const Info& red, float& a1, float&) const {
    int64_t o = in_offset(row, red);
    float v = o >= 0 ? xs[o] : 0.f;
    if (ms && o >= 0) v = __fmul_rn(__fmul_rn(v, ms[(int64_t)row.c * g.C + red.c]), in_mult);  // dropout.py:38-39
    a1 = v;
  }
  __device__ void load_b(const Info& col, const Info& red, float& b1, float&) const {
    if (col.valid && red.valid) b1 = ws[col.base + red.base];
  }
  __device__ void store(int z, int64_t m, int n, float acc, float) const {
    int64_t o = ((int64_t)z * g.M + m) * g.N + n;
    float v = acc;
    if (scale) v = __fmul_rn(v, scale[n]);
    if (shift) v = __fadd_rn(v, shift[n]);
    if (residual) v = __fadd_rn(v, residual[o]);
    if (flags & QBN_FLAG_RELU) v = fmaxf(v, 0.f);
    out[o] = v;
  }
};

// ---- A3 dx ----------------------------------------------------------------------------------
template <int BN_>
struct Dgrad {
  static constexpr int BN = BN_, BM = BM_DEFAULT;   // (128-row tiles were measured: no gain, 147 registers)
  static constexpr bool DUAL = true, A2_SQUARE = false, B2_SQUARE = false;
  Geom g;
  const float* gout; const float* dv; const float* mu; const float* sig2; const float* x; float* dx;
  __device__ void begin(int) {}
  __device__ int64_t rows_total() const { return (int64_t)g.B * g.H * g.W; }
  __device__ int cols_total() const { return g.C; }
  __device__ int64_t red_begin(int) const { return 0; }
  __device__ int64_t red_end(int) const { return (int64_t)g.R * g.S * g.N; }
  __device__ Info row_info(int64_t m) const {  // input pixel
    Info r;
    r.valid = m < rows_total();
    int64_t mm = r.valid ? m : 0;
    int w = (int)(mm % g.W);
    int64_t t = mm / g.W;
    int h = (int)(t % g.H);
    int b = (int)(t / g.H);
    r.a = h + g.ph; r.b = w + g.pw; r.c = b;
    r.base = (int64_t)b * g.Ho * g.Wo * g.N;
    return r;
  }
  __device__ Info col_info(int c) const { Info q; q.valid = c < g.C; q.c = q.valid ? c : 0; q.a = q.b = 0; q.base = 0; return q; }
  __device__ Info red_info(int64_t k, bool ok) const {  // k -> (r, s, n)
    Info q;
    q.valid = ok && k < red_end(0);
    int kk = q.valid ? (int)k : 0;
    q.c = kk % g.N;             // n
    int rs = kk / g.N;
    int s = rs % g.S, r = rs / g.S;
    q.a = r * g.dh; q.b = s * g.dw;
    q.base = (int64_t)q.c * g.K + (int64_t)rs * g.C;  // offset of W[n][r][s][0]
    return q;
  }
  __device__ void load_a(const Info& row, const Info& red, float& a1, float& a2) const {
    if (!(row.valid && red.valid)) return;
    int hn = row.a - red.a, wn = row.b - red.b;
    if (hn < 0 || wn < 0 || hn % g.sh || wn % g.sw) return;
    int ho = hn / g.sh, wo = wn / g.sw;
    if (ho >= g.Ho || wo >= g.Wo) return;
    int64_t o = row.base + ((int64_t)ho * g.Wo + wo) * g.N + red.c;
    a1 = gout[o];
    a2 = dv[o];
  }
  __device__ void load_b(const Info& col, const Info& red, float& b1, float& b2) const {
    if (col.valid && red.valid) { b1 = mu[red.base + col.c]; b2 = sig2[red.base + col.c]; }
  }
  __device__ void store(int, int64_t m, int c, float a1, float a2) const {
    int64_t o = m * g.C + c;
    dx[o] = a1 + 2.0f * x[o] * a2;
  }
};

// ---- A3 dmu / dsigma^2 (split over output pixels) ---------------------------------------------
template <int BN_>
struct Wgrad : PixelRows {
  static constexpr int BN = BN_, BM = BM_DEFAULT;
  static constexpr bool DUAL = true, A2_SQUARE = false, B2_SQUARE = true;
  const float* gout; const float* dv; const float* x;
  float* part1; float* part2;  // [splits][N][K]
  int64_t chunk;
  __device__ void begin(int) {}
  __device__ int64_t rows_total() const { return g.N; }
  __device__ int cols_total() const { return g.K; }
  __device__ int64_t red_begin(int z) const { return (int64_t)z * chunk; }
  __device__ int64_t red_end(int z) const { int64_t e = (int64_t)(z + 1) * chunk; return e < g.M ? e : g.M; }
  __device__ Info row_info(int64_t n) const { Info r; r.valid = n < g.N; r.c = r.valid ? (int)n : 0; r.a = r.b = 0; r.base = 0; return r; }
  __device__ Info col_info(int k) const { return tap_red(k, true); }
  __device__ Info red_info(int64_t m, bool ok) const { Info r = pixel_row(m); r.valid = r.valid && ok; return r; }
  __device__ void load_a(const Info& row, const Info& red, float& a1, float& a2) const {
    // red carries the output pixel (h0, w0, b); recover its linear index
    if (!(row.valid && red.valid)) return;
    int ho = (red.a + g.ph) / g.sh, wo = (red.b + g.pw) / g.sw;
    int64_t m = ((int64_t)red.c * g.Ho + ho) * g.Wo + wo;
    a1 = gout[m * g.N + row.c];
    a2 = dv[m * g.N + row.c];
  }
  __device__ void load_b(const Info& col, const Info& red, float& b1, float&) const {
    int64_t o = in_offset(red, col);
    b1 = o >= 0 ? x[o] : 0.f;
  }
  __device__ void store(int z, int64_t n, int k, float a1, float a2) const {
    int64_t o = ((int64_t)z * g.N + n) * g.K + k;
    part1[o] = a1;
    part2[o] = a2;
  }
};

__global__ void lrt_dv_kernel(const float* __restrict__ g, const float* __restrict__ sd, const float* __restrict__ eps, int64_t n,
                              uint64_t seed, uint32_t sa, uint32_t sb, const uint32_t* __restrict__ sbase, float* __restrict__ dv) {
  if (sbase) sb += *sbase;           // the forward's draw (device-side offset, qbn_set_sample_base)
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float e = eps ? eps[i] : philox_normal1(seed, sa, sb, (uint64_t)i);
    dv[i] = g[i] * e / (2.0f * sd[i]);
  }
}

__global__ void split_reduce_kernel(const float* __restrict__ part, int splits, int64_t n, float* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += part[(int64_t)s * n + i];
    out[i] = acc;
  }
}

// column sums of g [M][N] -> dbias[N]; one block per 32 columns, deterministic
__global__ void colsum_kernel(const float* __restrict__ g, int64_t M, int N, float* __restrict__ out) {
  __shared__ float sh[8][33];
  int n = blockIdx.x * 32 + threadIdx.x;
  float acc = 0.f;
  if (n < N)
    for (int64_t m = threadIdx.y; m < M; m += 8) acc += g[m * N + n];
  sh[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    for (int j = 1; j < 8; ++j) acc += sh[j][threadIdx.x];
    out[n] = acc;
  }
}

static int wgrad_splits(const Geom& g) {
  // enough CTAs to fill the machine, but keep each slice >= 256 pixels
  int tiles = (int)(ceil_div64(g.N, BM_DEFAULT) * ceil_div64(g.K, 64));
  int want = (2 * qbn_sm_count() + tiles - 1) / tiles;
  int64_t maxs = ceil_div64(g.M, 256);
  if (want > maxs) want = (int)maxs;
  if (want < 1) want = 1;
  if (want > 512) want = 512;
  return want;
}

template <class P>
static void launch(P& p, int64_t rows, int cols, int nz, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(rows, P::BM), (unsigned)ceil_div64(cols, P::BN), (unsigned)nz);
  igemm_kernel<P><<<grid, NT, 0, st>>>(p);
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
int qbn_umma_lrt_fwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p, const float* bias,
                     const float* eps, uint64_t seed, uint32_t sa, uint32_t sb, float* out, float* std_out, cudaStream_t st);
int qbn_umma_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const float* x, const float* w, int w_shared,
                      const float* scale, const float* shift, const float* residual, int flags, const float* in_mask,
                      float in_mult, float* out, cudaStream_t st);

extern "C" int qbn_lrt_fwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p, const float* bias,
                           const float* eps, uint64_t seed, uint32_t sa, uint32_t sb, float* out, float* std_out, int math_mode,
                           void* stream) {
  QBN_CHECK_ARG(check_desc(d), "conv descriptor inconsistent");
  QBN_CHECK_ARG(x && mu_p && sig2_p && out, "null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode == QBN_MATH_TF32) return qbn_umma_lrt_fwd(d, x, mu_p, sig2_p, bias, eps, seed, sa, sb, out, std_out, st);
  QBN_CHECK_ARG(math_mode == QBN_MATH_FP32, "math_mode");
  QBN_CHECK_ARG(d->out_pad_h == 0 && d->out_pad_w == 0, "out_pad is a tcgen05-path feature");
  Geom g = make_geom(d);
  if (g.N <= 32) {
    FwdLRT<32> p; p.g = g; p.x = x; p.mu = mu_p; p.sig2 = sig2_p; p.bias = bias; p.eps = eps; p.out = out; p.std_out = std_out;
    p.seed = seed; p.sa = sa; p.sb = sb; p.sbase = qbn_sample_base_ptr();
    launch(p, g.M, g.N, 1, st);
  } else {
    FwdLRT<64> p; p.g = g; p.x = x; p.mu = mu_p; p.sig2 = sig2_p; p.bias = bias; p.eps = eps; p.out = out; p.std_out = std_out;
    p.seed = seed; p.sa = sa; p.sb = sb; p.sbase = qbn_sample_base_ptr();
    launch(p, g.M, g.N, 1, st);
  }
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

extern "C" int qbn_conv_fwd(const qbn_conv_desc* d, int n_samples, int x_shared, const float* x, const float* w, int w_shared,
                            const float* scale, const float* shift, const float* residual, int flags, const float* in_mask,
                            float in_mult, float* out, int math_mode, void* stream) {
  QBN_CHECK_ARG(check_desc(d), "conv descriptor inconsistent");
  QBN_CHECK_ARG(x && w && out, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && n_samples <= 65535, "0 < n_samples <= 65535");
  cudaStream_t st = (cudaStream_t)stream;
  if (math_mode == QBN_MATH_TF32)
    return qbn_umma_conv_fwd(d, n_samples, x_shared, x, w, w_shared, scale, shift, residual, flags, in_mask, in_mult, out, st);
  QBN_CHECK_ARG(math_mode == QBN_MATH_FP32, "math_mode");
  QBN_CHECK_ARG(d->out_pad_h == 0 && d->out_pad_w == 0, "out_pad is a tcgen05-path feature");
  Geom g = make_geom(d);
#define QBN_FILL_EVAL(p)                                                                                        \
  p.g = g; p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.residual = residual; p.in_mask = in_mask;      \
  p.out = out; p.in_mult = in_mult; p.x_shared = x_shared; p.w_shared = w_shared; p.flags = flags;                 \
  p.x_sample_stride = (int64_t)g.B * g.H * g.W * g.C; p.w_sample_stride = (int64_t)g.N * g.K;
  if (g.N <= 32) {
    FwdEval<32> p; QBN_FILL_EVAL(p);
    launch(p, g.M, g.N, n_samples, st);
  } else {
    FwdEval<64> p; QBN_FILL_EVAL(p);
    launch(p, g.M, g.N, n_samples, st);
  }
#undef QBN_FILL_EVAL
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}

int qbn_umma_lrt_dgrad(const qbn_conv_desc* d, const float* g, const float* dv, const float* mu_t, const float* sig2_t, const float* x,
                       float* dx, cudaStream_t st);
int qbn_umma_lrt_wgrad(const qbn_conv_desc* d, const float* x, const float* g, const float* dv, float* part1, float* part2, int splits,
                       cudaStream_t st);

// tcgen05 dgrad: undilated layers (any stride) whose channel counts fit the 16-byte K chunks / one TMEM tile
static bool dgrad_tf32_ok(const qbn_conv_desc* d) {
  return d->dil_h == 1 && d->dil_w == 1 && d->N % 4 == 0 && d->C <= 256 &&
         d->pad_h <= d->R - 1 && d->pad_w <= d->S - 1 && d->pad_h >= 0 && d->pad_w >= 0 && d->out_pad_h == 0 && d->out_pad_w == 0;
}

// mu_p / sig2_p [N][R][S][C] -> [C][R'][S'][N] with (r', s') = (R-1-r, S-1-s), RNA-rounded to TF32
__global__ void dgrad_weights_kernel(const float* __restrict__ mu_p, const float* __restrict__ sig2_p, int N, int R, int S, int C,
                                     float* __restrict__ mu_t, float* __restrict__ sig2_t) {
  const int64_t total = (int64_t)N * R * S * C;
  for (int64_t o = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; o < total; o += (int64_t)gridDim.x * blockDim.x) {
    const int n = (int)(o % N);
    int64_t t = o / N;
    const int s2 = (int)(t % S);
    t /= S;
    const int r2 = (int)(t % R);
    const int c = (int)(t / R);
    const int64_t src = (((int64_t)n * R + (R - 1 - r2)) * S + (S - 1 - s2)) * C + c;
    mu_t[o] = tf32_round(mu_p[src]);
    sig2_t[o] = tf32_round(sig2_p[src]);
  }
}

extern "C" size_t qbn_lrt_bwd_workspace_bytes(const qbn_conv_desc* d) {
  if (!check_desc(d)) return 0;
  Geom g = make_geom(d);
  int splits = wgrad_splits(g);
  size_t dv = (size_t)g.M * g.N * sizeof(float);
  size_t parts = (size_t)2 * splits * g.N * g.K * sizeof(float);
  size_t wt = (size_t)2 * g.N * g.K * sizeof(float);     // flipped / transposed weights of the tcgen05 dgrad
  return dv + parts + wt + 512;
}

extern "C" int qbn_lrt_bwd(const qbn_conv_desc* d, const float* x, const float* mu_p, const float* sig2_p, const float* grad_out,
                           const float* std_saved, const float* eps, uint64_t seed, uint32_t sa, uint32_t sb, float* dx,
                           float* dmu_p, float* dsig2_p, float* dbias, void* workspace, size_t workspace_bytes, int math_mode,
                           void* stream) {
  QBN_CHECK_ARG(check_desc(d), "conv descriptor inconsistent");
  QBN_CHECK_ARG(x && mu_p && sig2_p && grad_out && std_saved && dmu_p && dsig2_p && workspace, "null pointer");
  QBN_CHECK_ARG(math_mode == QBN_MATH_FP32 || math_mode == QBN_MATH_TF32, "math_mode");
  if (workspace_bytes < qbn_lrt_bwd_workspace_bytes(d)) {
    qbn_set_error("qbn_lrt_bwd: workspace too small (%zu < %zu)", workspace_bytes, qbn_lrt_bwd_workspace_bytes(d));
    return QBN_ERR_WORKSPACE;
  }
  cudaStream_t st = (cudaStream_t)stream;
  Geom g = make_geom(d);
  const int64_t MN = g.M * g.N;
  float* dv = reinterpret_cast<float*>(workspace);
  size_t dv_bytes = ((size_t)MN * sizeof(float) + 255) / 256 * 256;
  int splits = wgrad_splits(g);
  float* part1 = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + dv_bytes);
  float* part2 = part1 + (int64_t)splits * g.N * g.K;

  lrt_dv_kernel<<<qbn_grid_for(MN, 256), 256, 0, st>>>(grad_out, std_saved, eps, MN, seed, sa, sb, qbn_sample_base_ptr(), dv);
  QBN_CHECK_LAUNCH();
  if (dx && math_mode == QBN_MATH_TF32 && dgrad_tf32_ok(d)) {
    float* mu_t = part2 + (int64_t)splits * g.N * g.K;
    mu_t = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(mu_t) + 255) / 256 * 256);
    float* sig2_t = mu_t + (int64_t)g.N * g.K;
    dgrad_weights_kernel<<<qbn_grid_for((int64_t)g.N * g.K, 256), 256, 0, st>>>(mu_p, sig2_p, g.N, g.R, g.S, g.C, mu_t, sig2_t);
    QBN_CHECK_LAUNCH();
    int rc = qbn_umma_lrt_dgrad(d, grad_out, dv, mu_t, sig2_t, x, dx, st);
    if (rc != QBN_OK) return rc;
  } else if (dx) {
    if (g.C <= 32) {
      Dgrad<32> p; p.g = g; p.gout = grad_out; p.dv = dv; p.mu = mu_p; p.sig2 = sig2_p; p.x = x; p.dx = dx;
      launch(p, (int64_t)g.B * g.H * g.W, g.C, 1, st);
    } else {
      Dgrad<64> p; p.g = g; p.gout = grad_out; p.dv = dv; p.mu = mu_p; p.sig2 = sig2_p; p.x = x; p.dx = dx;
      launch(p, (int64_t)g.B * g.H * g.W, g.C, 1, st);
    }
    QBN_CHECK_LAUNCH();
  }
  if (math_mode == QBN_MATH_TF32) {
    // tcgen05 weight gradients: pixels as the reduction dimension, split over CTAs into the same workspace
    const int kt = g.K >= 128 ? 128 : (g.K + 15) / 16 * 16;
    const int tiles = (int)(ceil_div64(g.N, 128) * ceil_div64(g.K, kt));
    int splits_t = (2 * qbn_sm_count() + tiles - 1) / tiles;
    const int64_t maxs = ceil_div64(g.M, 256);
    if (splits_t > maxs) splits_t = (int)maxs;
    if (splits_t > splits) splits_t = splits;
    if (splits_t < 1) splits_t = 1;
    int rc = qbn_umma_lrt_wgrad(d, x, grad_out, dv, part1, part2, splits_t, st);
    if (rc != QBN_OK) return rc;
    int64_t nk = (int64_t)g.N * g.K;
    split_reduce_kernel<<<qbn_grid_for(nk, 256), 256, 0, st>>>(part1, splits_t, nk, dmu_p);
    split_reduce_kernel<<<qbn_grid_for(nk, 256), 256, 0, st>>>(part2, splits_t, nk, dsig2_p);
    QBN_CHECK_LAUNCH();
  } else {
    Wgrad<64> p; p.g = g; p.gout = grad_out; p.dv = dv; p.x = x; p.part1 = part1; p.part2 = part2;
    p.chunk = ceil_div64(ceil_div64(g.M, splits), BK) * BK;
    launch(p, g.N, g.K, splits, st);
    QBN_CHECK_LAUNCH();
    int64_t nk = (int64_t)g.N * g.K;
    split_reduce_kernel<<<qbn_grid_for(nk, 256), 256, 0, st>>>(part1, splits, nk, dmu_p);
    split_reduce_kernel<<<qbn_grid_for(nk, 256), 256, 0, st>>>(part2, splits, nk, dsig2_p);
    QBN_CHECK_LAUNCH();
  }
  if (dbias) {
    colsum_kernel<<<(g.N + 31) / 32, dim3(32, 8), 0, st>>>(grad_out, g.M, g.N, dbias);
    QBN_CHECK_LAUNCH();
  }
  return QBN_OK;
}
