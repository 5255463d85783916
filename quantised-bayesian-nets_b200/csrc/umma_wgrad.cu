// A3 weight gradients on the tensor cores (QBN_MATH_TF32):
//
//   dmu[n][k]     = sum_m g [m][n] * X  [m][k]          k = (r, s, c), X = im2col(x), m = output pixels
//   dsigma2[n][k] = sum_m dv[m][n] * X^2[m][k]
//
// as D[128 rows = n][KT cols = k] (+)= A[n][m] * B[k][m]^T with the PIXELS as the tcgen05 reduction dimension (K = 8 pixels
// per kind::tf32 instruction) and both contractions accumulating side by side in TMEM.  Four producer warps stage the
// operands K-major through registers: A = g^T / dv^T (four consecutive pixels of one channel = one 16-byte chunk), B = the
// im2col gather of x with x^2 formed in registers — neither transposed tensor nor x^2 nor the im2col matrix exists in HBM.
// The pixel range is split over blockIdx.z; partial sums go to the workspace and are reduced deterministically by
// gemm_fp32.cu's split_reduce_kernel.  RNA rounding to TF32 on the way into shared memory.
#include <string.h>
#include "umma_common.cuh"

namespace {

constexpr int WM = 128;           // accumulator rows = output channels per CTA
constexpr int WCH = 8;            // 16-byte chunks (of 4 pixels) per stage: 32 pixels
constexpr int WPIX = WCH * 4;
constexpr int WPROD = 128;
constexpr int WTHREADS = 160;

struct WParams {
  int B, H, W, C, N, R, S, sh, sw, ph, pw, dh, dw, Ho, Wo, K;
  long long M;              // B*Ho*Wo
  int KT;                   // k columns per CTA (multiple of 16, <= 128)
  int stages, a_pitch, b_pitch, tmem_cols;
  long long chunk;          // pixels per split (multiple of 32)
  uint32_t idesc;
  const float* x; const float* g; const float* dv;
  float* part1; float* part2;     // [splits][N][K]
};

struct PixInfo { int h0, w0, base, valid; };

__global__ void __launch_bounds__(WTHREADS) umma_wgrad_kernel(const WParams p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int k0 = blockIdx.x * p.KT;
  const int n0 = blockIdx.y * WM;
  const long long m_begin = (long long)blockIdx.z * p.chunk;
  long long m_end = m_begin + p.chunk;
  if (m_end > p.M) m_end = p.M;
  const int n_stage = m_begin < m_end ? (int)((m_end - m_begin + WPIX - 1) / WPIX) : 0;

  const uint32_t a_bytes = (uint32_t)WCH * p.a_pitch * 16;
  const uint32_t b_bytes = (uint32_t)WCH * p.b_pitch * 16;
  const uint32_t stage_bytes = 2 * (a_bytes + b_bytes);
  uint8_t* ring = smem;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)p.stages * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* accum_bar = bars + 2 * p.stages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * p.stages + 1);
  PixInfo* pix = reinterpret_cast<PixInfo*>(tmem_slot + 4);          // [stages][32]

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(smem_u32(&full_bar[s]), WPROD);
      mbar_init(smem_u32(&empty_bar[s]), 1);
    }
    mbar_init(smem_u32(accum_bar), 1);
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp < 4) {
    // =========================== PRODUCERS ======================================================
    // this thread's k columns (static over the stages): k = k0 + tid (+128 would exceed KT <= 128)
    const int kk = tid;
    const int k = k0 + kk;
    const bool kvalid = kk < p.KT && k < p.K;
    int kc = 0, kr = 0, ks = 0;
    if (kvalid) {
      kc = k % p.C;
      const int rs = k / p.C;
      ks = (rs % p.S) * p.dw;
      kr = (rs / p.S) * p.dh;
    }
    int stage = 0;
    uint32_t phase = 0;
    for (int st = 0; st < n_stage; ++st) {
      if (lane == 0) mbar_wait(smem_u32(&empty_bar[stage]), phase ^ 1);
      __syncwarp();
      const long long mb = m_begin + (long long)st * WPIX;
      // ---- pixel table of the stage (32 pixels): decoded once, used by every k column
      if (tid < WPIX) {
        const long long m = mb + tid;
        PixInfo pi;
        pi.valid = m < m_end;
        const long long mm = pi.valid ? m : 0;
        const int wo = (int)(mm % p.Wo);
        const long long t2 = mm / p.Wo;
        const int ho = (int)(t2 % p.Ho);
        const int b = (int)(t2 / p.Ho);
        pi.h0 = ho * p.sh - p.ph;
        pi.w0 = wo * p.sw - p.pw;
        pi.base = b * p.H * p.W * p.C;
        pix[stage * WPIX + tid] = pi;
      }
      asm volatile("bar.sync 1, 128;" ::: "memory");
      uint8_t* sa1 = ring + (size_t)stage * stage_bytes;
      uint8_t* sa2 = sa1 + a_bytes;
      uint8_t* sb1 = sa2 + a_bytes;
      uint8_t* sb2 = sb1 + b_bytes;
      // ---- A = g^T, dv^T: four consecutive pixels of one channel = one 16-byte chunk.  N % 4 == 0: one item = 4 channels x
      //      4 pixels, loaded as four float4 (channels contiguous) and transposed in registers
      const int nrows = min(WM, p.N - n0);
      if ((p.N & 3) == 0) {
        const int nq = nrows >> 2;
        for (int item = tid; item < WCH * nq; item += WPROD) {
          const int q = item % nq, j = item / nq;
          float4 a[4], d2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const long long m = mb + 4 * j + e;
            const bool ok = m < m_end;
            const long long off = (ok ? m : 0) * p.N + n0 + 4 * q;
            a[e] = ok ? __ldg(reinterpret_cast<const float4*>(p.g + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
            d2[e] = ok ? __ldg(reinterpret_cast<const float4*>(p.dv + off)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const size_t so = ((size_t)j * p.a_pitch + 4 * q) * 16;
          *reinterpret_cast<uint4*>(sa1 + so) = make_uint4(tf32_rna(a[0].x), tf32_rna(a[1].x), tf32_rna(a[2].x), tf32_rna(a[3].x));
          *reinterpret_cast<uint4*>(sa1 + so + 16) = make_uint4(tf32_rna(a[0].y), tf32_rna(a[1].y), tf32_rna(a[2].y), tf32_rna(a[3].y));
          *reinterpret_cast<uint4*>(sa1 + so + 32) = make_uint4(tf32_rna(a[0].z), tf32_rna(a[1].z), tf32_rna(a[2].z), tf32_rna(a[3].z));
          *reinterpret_cast<uint4*>(sa1 + so + 48) = make_uint4(tf32_rna(a[0].w), tf32_rna(a[1].w), tf32_rna(a[2].w), tf32_rna(a[3].w));
          *reinterpret_cast<uint4*>(sa2 + so) = make_uint4(tf32_rna(d2[0].x), tf32_rna(d2[1].x), tf32_rna(d2[2].x), tf32_rna(d2[3].x));
          *reinterpret_cast<uint4*>(sa2 + so + 16) = make_uint4(tf32_rna(d2[0].y), tf32_rna(d2[1].y), tf32_rna(d2[2].y), tf32_rna(d2[3].y));
          *reinterpret_cast<uint4*>(sa2 + so + 32) = make_uint4(tf32_rna(d2[0].z), tf32_rna(d2[1].z), tf32_rna(d2[2].z), tf32_rna(d2[3].z));
          *reinterpret_cast<uint4*>(sa2 + so + 48) = make_uint4(tf32_rna(d2[0].w), tf32_rna(d2[1].w), tf32_rna(d2[2].w), tf32_rna(d2[3].w));
        }
      } else {
        for (int idx = tid; idx < WCH * nrows; idx += WPROD) {
          const int n = idx % nrows, j = idx / nrows;
          float a[4], d2[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const long long m = mb + 4 * j + e;
            const bool ok = m < m_end;
            const long long off = (ok ? m : 0) * p.N + n0 + n;
            a[e] = ok ? __ldg(p.g + off) : 0.f;
            d2[e] = ok ? __ldg(p.dv + off) : 0.f;
          }
          const size_t so = ((size_t)j * p.a_pitch + n) * 16;
          *reinterpret_cast<uint4*>(sa1 + so) = make_uint4(tf32_rna(a[0]), tf32_rna(a[1]), tf32_rna(a[2]), tf32_rna(a[3]));
          *reinterpret_cast<uint4*>(sa2 + so) = make_uint4(tf32_rna(d2[0]), tf32_rna(d2[1]), tf32_rna(d2[2]), tf32_rna(d2[3]));
        }
      }
      // ---- B = X^T, (X^2)^T.  C % 4 == 0: one item = 4 consecutive channels of one tap x 4 pixels (four float4 loads,
      //      register transpose); otherwise one k column per thread with scalar loads (the 3-channel first layer)
      if ((p.C & 3) == 0) {
        const int nq = p.KT >> 2;
        for (int item = tid; item < WCH * nq; item += WPROD) {
          const int q = item % nq, j = item / nq;
          const int kq = k0 + 4 * q;
          const bool qv = kq < p.K;
          const int kk4 = qv ? kq : 0;
          const int c4 = kk4 % p.C, rs4 = kk4 / p.C;
          const int ds4 = (rs4 % p.S) * p.dw, dr4 = (rs4 / p.S) * p.dh;
          float4 v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const PixInfo pi = pix[stage * WPIX + 4 * j + e];
            const int hi = pi.h0 + dr4, wi = pi.w0 + ds4;
            const bool ok = qv && pi.valid && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            v[e] = ok ? __ldg(reinterpret_cast<const float4*>(p.x + pi.base + (hi * p.W + wi) * p.C + c4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          const size_t so = ((size_t)j * p.b_pitch + 4 * q) * 16;
          *reinterpret_cast<uint4*>(sb1 + so) = make_uint4(tf32_rna(v[0].x), tf32_rna(v[1].x), tf32_rna(v[2].x), tf32_rna(v[3].x));
          *reinterpret_cast<uint4*>(sb1 + so + 16) = make_uint4(tf32_rna(v[0].y), tf32_rna(v[1].y), tf32_rna(v[2].y), tf32_rna(v[3].y));
          *reinterpret_cast<uint4*>(sb1 + so + 32) = make_uint4(tf32_rna(v[0].z), tf32_rna(v[1].z), tf32_rna(v[2].z), tf32_rna(v[3].z));
          *reinterpret_cast<uint4*>(sb1 + so + 48) = make_uint4(tf32_rna(v[0].w), tf32_rna(v[1].w), tf32_rna(v[2].w), tf32_rna(v[3].w));
          *reinterpret_cast<uint4*>(sb2 + so) = make_uint4(tf32_rna(v[0].x * v[0].x), tf32_rna(v[1].x * v[1].x), tf32_rna(v[2].x * v[2].x), tf32_rna(v[3].x * v[3].x));
          *reinterpret_cast<uint4*>(sb2 + so + 16) = make_uint4(tf32_rna(v[0].y * v[0].y), tf32_rna(v[1].y * v[1].y), tf32_rna(v[2].y * v[2].y), tf32_rna(v[3].y * v[3].y));
          *reinterpret_cast<uint4*>(sb2 + so + 32) = make_uint4(tf32_rna(v[0].z * v[0].z), tf32_rna(v[1].z * v[1].z), tf32_rna(v[2].z * v[2].z), tf32_rna(v[3].z * v[3].z));
          *reinterpret_cast<uint4*>(sb2 + so + 48) = make_uint4(tf32_rna(v[0].w * v[0].w), tf32_rna(v[1].w * v[1].w), tf32_rna(v[2].w * v[2].w), tf32_rna(v[3].w * v[3].w));
        }
      } else if (kk < p.KT) {
#pragma unroll 2
        for (int j = 0; j < WCH; ++j) {
          float v[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const PixInfo pi = pix[stage * WPIX + 4 * j + e];
            const int hi = pi.h0 + kr, wi = pi.w0 + ks;
            const bool ok = kvalid && pi.valid && hi >= 0 && hi < p.H && wi >= 0 && wi < p.W;
            v[e] = ok ? __ldg(p.x + pi.base + (hi * p.W + wi) * p.C + kc) : 0.f;
          }
          const size_t so = ((size_t)j * p.b_pitch + kk) * 16;
          *reinterpret_cast<uint4*>(sb1 + so) = make_uint4(tf32_rna(v[0]), tf32_rna(v[1]), tf32_rna(v[2]), tf32_rna(v[3]));
          *reinterpret_cast<uint4*>(sb2 + so) =
              make_uint4(tf32_rna(v[0] * v[0]), tf32_rna(v[1] * v[1]), tf32_rna(v[2] * v[2]), tf32_rna(v[3] * v[3]));
        }
      }
      fence_proxy_async();
      mbar_arrive(smem_u32(&full_bar[stage]));
      if (++stage == p.stages) { stage = 0; phase ^= 1; }
    }
  } else {
    // =========================== MMA ISSUER (one elected thread) ==================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      const uint32_t lbo_a = (uint32_t)p.a_pitch * 16, lbo_b = (uint32_t)p.b_pitch * 16;
      const uint64_t adesc_hi = make_smem_desc(0, lbo_a, 128), bdesc_hi = make_smem_desc(0, lbo_b, 128);
      for (int st = 0; st < n_stage; ++st) {
        mbar_wait(smem_u32(&full_bar[stage]), phase);
        fence_proxy_async();
        tc_fence_after();
        const uint32_t sa1 = smem_u32(ring + (size_t)stage * stage_bytes);
        const uint32_t sa2 = sa1 + a_bytes, sb1 = sa2 + a_bytes, sb2 = sb1 + b_bytes;
#pragma unroll 1
        for (int j = 0; j < WCH / 2; ++j) {
          const uint32_t acc = (st > 0 || j > 0) ? 1u : 0u;
          umma_mma<MODE_EVAL>(tmem_base, adesc_hi | (uint64_t)(((sa1 + 2 * j * lbo_a) >> 4) & 0x3FFF),
                              bdesc_hi | (uint64_t)(((sb1 + 2 * j * lbo_b) >> 4) & 0x3FFF), p.idesc, acc);
          umma_mma<MODE_EVAL>(tmem_base + (uint32_t)p.KT, adesc_hi | (uint64_t)(((sa2 + 2 * j * lbo_a) >> 4) & 0x3FFF),
                              bdesc_hi | (uint64_t)(((sb2 + 2 * j * lbo_b) >> 4) & 0x3FFF), p.idesc, acc);
        }
        umma_commit(smem_u32(&empty_bar[stage]));
        if (st == n_stage - 1) umma_commit(smem_u32(accum_bar));
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  }

  // =============================== EPILOGUE (warps 0-3): lane = output channel ===================
  if (warp < 4) {
    const int n = n0 + warp * 32 + lane;
    const bool nv = n < p.N;
    float* o1 = p.part1 + ((size_t)blockIdx.z * p.N + (nv ? n : 0)) * p.K;
    float* o2 = p.part2 + ((size_t)blockIdx.z * p.N + (nv ? n : 0)) * p.K;
    if (n_stage > 0) {
      if (lane == 0) mbar_wait(smem_u32(accum_bar), 0);
      __syncwarp();
      tc_fence_after();
    }
    const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16);
    for (int c0 = 0; c0 < p.KT; c0 += 8) {
      uint32_t v1[8], v2[8];
      if (n_stage > 0) {
        tmem_ld8(tlane + (uint32_t)c0, v1);
        tmem_ld8(tlane + (uint32_t)(p.KT + c0), v2);
        tmem_ld_wait();
      }
      if (!nv) continue;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int kq = k0 + c0 + j;
        if (kq < p.K) {
          o1[kq] = n_stage > 0 ? __uint_as_float(v1[j]) : 0.f;
          o2[kq] = n_stage > 0 ? __uint_as_float(v2[j]) : 0.f;
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

// part1 / part2: [splits][N][K] partial sums (every element written).  Returns QBN_ERR_UNSUPPORTED for shapes the caller
// should route to the fp32 kernels.
int qbn_umma_lrt_wgrad(const qbn_conv_desc* d, const float* x, const float* g, const float* dv, float* part1, float* part2, int splits,
                       cudaStream_t st) {
  WParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B; p.H = d->H; p.W = d->W; p.C = d->C; p.N = d->N; p.R = d->R; p.S = d->S;
  p.sh = d->stride_h; p.sw = d->stride_w; p.ph = d->pad_h; p.pw = d->pad_w; p.dh = d->dil_h; p.dw = d->dil_w;
  p.Ho = d->Ho; p.Wo = d->Wo; p.K = d->R * d->S * d->C;
  p.M = (long long)d->B * d->Ho * d->Wo;
  if ((long long)d->B * d->H * d->W * d->C >= (1ll << 31)) {
    qbn_set_error("qbn_lrt_bwd(TF32 wgrad): input too large for 32-bit offsets");
    return QBN_ERR_UNSUPPORTED;
  }
  p.KT = p.K >= 128 ? 128 : (p.K + 15) / 16 * 16;
  p.a_pitch = WM + 1;
  p.b_pitch = p.KT + 1;
  p.tmem_cols = 32;
  while (p.tmem_cols < 2 * p.KT) p.tmem_cols <<= 1;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.KT >> 3) << 17) | ((uint32_t)(WM >> 4) << 24);
  const size_t stage_bytes = (size_t)2 * WCH * 16 * (p.a_pitch + p.b_pitch);
  p.stages = (int)((196 * 1024) / stage_bytes);
  if (p.stages > 3) p.stages = 3;
  if (p.stages < 2) {
    qbn_set_error("qbn_lrt_bwd(TF32 wgrad): stage does not fit shared memory");
    return QBN_ERR_UNSUPPORTED;
  }
  const long long per = (p.M + splits - 1) / splits;
  p.chunk = (per + WPIX - 1) / WPIX * WPIX;
  p.x = x; p.g = g; p.dv = dv; p.part1 = part1; p.part2 = part2;
  const size_t smem = p.stages * stage_bytes + (2 * p.stages + 1) * 8 + 16 + (size_t)p.stages * WPIX * sizeof(PixInfo);
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  dim3 grid((unsigned)((p.K + p.KT - 1) / p.KT), (unsigned)((p.N + WM - 1) / WM), (unsigned)splits);
  umma_wgrad_kernel<<<grid, WTHREADS, smem, st>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
