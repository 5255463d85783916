// tcgen05 "zero-copy im2col" kernel for the stride-1 same-padding convolutions (16 of the 21
// stochastic layers of the ResNet, ~85 % of its FLOPs), persistent and warp-specialised.
//
// Activations live in HBM in a zero-bordered ("padded") NHWC layout [B][H+2ph][W+2pw][C]; the output
// uses the same layout with N channels.  With q = the flat pixel index of that padded grid, the
// input row needed by output row q for filter tap (r,s) is simply q + (r-ph)*Wp + (s-pw): a constant
// shift.  So ONE shared-memory tile holding rows [q0-D, q0+128+D) of a channel block serves all R*S
// taps — each tap is the same tile addressed through a UMMA descriptor whose start address is moved
// by the shift (the canonical K-major no-swizzle layout keeps 8 consecutive rows in one 128-byte
// core matrix, and consecutive rows are consecutive q).  Every input element is fetched from
// L2/HBM once per tile instead of R*S times, with no bounds tests (the zero border IS the padding)
// and no per-element address arithmetic.  Border rows are computed but stored as zeros, which keeps
// the border of the next layer's input intact.
//
//   warps 0-3  epilogue   : TMEM -> registers -> affine(BN)/residual/ReLU/TF32-round -> global
//   warp  4    MMA issuer : tcgen05.mma kind::tf32, double-buffered TMEM accumulators
//   warps 5-8  producers  : cp.async (LDGSTS) into two rings — A (activation tile, one slot per
//                           channel block) and B (sampled weights, one slot per (channel block, tap))
// CTAs are persistent (grid = SMs x occupancy) and walk the tile list round-robin, so TMEM
// allocation, barrier setup and the epilogue of tile i overlap the loads/MMAs of tile i+1.
#include <string.h>
#include "umma_common.cuh"

namespace {

constexpr int TM = 128;
constexpr int N_EPI = 128, N_PROD = 128;
constexpr int NTHREADS_S1 = N_EPI + 32 + N_PROD;   // 288

struct S1Params {
  int Hp, Wp, C, N, R, S, ph, pw;
  int Qs;                         // rows (padded pixels) per Monte-Carlo sample
  int tiles_per_sample, total_tiles;
  int n_pad, CB, cbc, n_cb, K;    // MMA N, channels per block, 16-byte chunks per block (even), #blocks, R*S*C
  int D, RA, RA_p, b_pitch;       // halo rows, A rows per slot, pitches (in 16-byte chunks)
  int SA, SB;                     // ring depths
  int tmem_cols, flags, w_shared;
  uint32_t idesc;
  const float* x; const float* w; const float* scale; const float* shift; const float* residual; float* out;
};

__global__ void __launch_bounds__(NTHREADS_S1) umma_conv_s1_kernel(const S1Params p) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t a_bytes = (uint32_t)p.cbc * p.RA_p * 16;
  const uint32_t b_bytes = (uint32_t)p.cbc * p.b_pitch * 16;
  uint8_t* a_ring = smem;
  uint8_t* b_ring = smem + (size_t)p.SA * a_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(b_ring + (size_t)p.SB * b_bytes);
  uint64_t* a_full = bars;
  uint64_t* a_empty = a_full + p.SA;
  uint64_t* b_full = a_empty + p.SA;
  uint64_t* b_empty = b_full + p.SB;
  uint64_t* acc_full = b_empty + p.SB;     // [2]
  uint64_t* acc_empty = acc_full + 2;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);

  if (tid == 0) {
    for (int i = 0; i < p.SA; ++i) { mbar_init(smem_u32(&a_full[i]), N_PROD); mbar_init(smem_u32(&a_empty[i]), 1); }
    for (int i = 0; i < p.SB; ++i) { mbar_init(smem_u32(&b_full[i]), N_PROD); mbar_init(smem_u32(&b_empty[i]), 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(smem_u32(&acc_full[i]), 1); mbar_init(smem_u32(&acc_empty[i]), N_EPI); }
    fence_mbar_init();
    fence_proxy_async();
  }
  if (warp == 4) tmem_alloc(smem_u32(tmem_slot), (uint32_t)p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const int taps = p.R * p.S;

  if (warp >= 5) {
    // ======================================= PRODUCERS ==========================================
    const int pt = tid - (N_EPI + 32);            // 0..127
    const int CH = p.cbc;
    const int j = pt % CH;                        // this thread's 16-byte chunk inside every slot
    const int lane_row = pt / CH;
    const int RS = N_PROD / CH;                   // rows advanced per pass
    const bool active = lane_row < RS;            // (128 % CH) threads idle but still arrive
    int sa = 0, sb = 0;
    uint32_t pa = 0, pb = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int z = tile / p.tiles_per_sample;
      const int q0 = (tile - z * p.tiles_per_sample) * TM;
      const float* xs = p.x + (size_t)z * p.Qs * p.C;
      const float* ws = p.w + (p.w_shared ? 0 : (size_t)z * p.N * p.K);
      for (int cb = 0; cb < p.n_cb; ++cb) {
        const int c = cb * p.CB + 4 * j;
        const bool cv = active && c < p.C && 4 * j < p.CB;
        // ---- A slot: rows [q0-D, q0+TM+D) of channel block cb, one contiguous strip of the padded map
        mbar_wait(smem_u32(&a_empty[sa]), pa ^ 1);
        if (active) {
          const uint32_t dst0 = smem_u32(a_ring + (size_t)sa * a_bytes) + (uint32_t)(j * p.RA_p) * 16;
          for (int rho = lane_row; rho < p.RA; rho += RS) {
            const int q = q0 - p.D + rho;
            const bool ok = cv && q >= 0 && q < p.Qs;
            cp_async16(dst0 + (uint32_t)rho * 16, ok ? xs + (size_t)q * p.C + c : p.x, ok ? 16u : 0u);
          }
        }
        cp_async_arrive_noinc(smem_u32(&a_full[sa]));
        if (++sa == p.SA) { sa = 0; pa ^= 1; }
        // ---- B slots: sampled weights W[n][tap][cb block] for every tap
        for (int t = 0; t < taps; ++t) {
          mbar_wait(smem_u32(&b_empty[sb]), pb ^ 1);
          const uint32_t dst0 = smem_u32(b_ring + (size_t)sb * b_bytes) + (uint32_t)(j * p.b_pitch) * 16;
          const float* wt = ws + (size_t)t * p.C + c;
          for (int n = lane_row; active && n < p.n_pad; n += RS) {
            const bool ok = cv && n < p.N;
            cp_async16(dst0 + (uint32_t)n * 16, ok ? wt + (size_t)n * p.K : p.w, ok ? 16u : 0u);
          }
          cp_async_arrive_noinc(smem_u32(&b_full[sb]));
          if (++sb == p.SB) { sb = 0; pb ^= 1; }
        }
      }
    }
  } else if (warp == 4) {
    // ======================================= MMA ISSUER =========================================
    int sa = 0, sb = 0, as = 0;
    uint32_t pa = 0, pb = 0, pacc = 0;
    const uint32_t lbo_a = (uint32_t)p.RA_p * 16, lbo_b = (uint32_t)p.b_pitch * 16;
    const int nk = p.cbc / 2;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      mbar_wait(smem_u32(&acc_empty[as]), pacc ^ 1);      // epilogue has drained this accumulator
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(as * p.n_pad);
      uint32_t first = 1;
      for (int cb = 0; cb < p.n_cb; ++cb) {
        mbar_wait(smem_u32(&a_full[sa]), pa);
        const uint32_t abase = smem_u32(a_ring + (size_t)sa * a_bytes);
        for (int t = 0; t < taps; ++t) {
          mbar_wait(smem_u32(&b_full[sb]), pb);
          fence_proxy_async();          // cp.async data (generic proxy) -> ordered before async-proxy reads
          tc_fence_after();
          if (lane == 0) {
            const int r = t / p.S, s = t - r * p.S;
            const int shift = p.D + (r - p.ph) * p.Wp + (s - p.pw);     // row of the slot that output row 0 reads
            const uint32_t bbase = smem_u32(b_ring + (size_t)sb * b_bytes);
            for (int jj = 0; jj < nk; ++jj) {
              const uint64_t ad = make_smem_desc(abase + (uint32_t)(2 * jj) * lbo_a + (uint32_t)shift * 16, lbo_a, 128);
              const uint64_t bd = make_smem_desc(bbase + (uint32_t)(2 * jj) * lbo_b, lbo_b, 128);
              umma_mma<MODE_EVAL>(tacc, ad, bd, p.idesc, first ? 0u : 1u);
              first = 0;
            }
            umma_commit(smem_u32(&b_empty[sb]));
            if (t == taps - 1) {
              umma_commit(smem_u32(&a_empty[sa]));
              if (cb == p.n_cb - 1) umma_commit(smem_u32(&acc_full[as]));
            }
          }
          __syncwarp();
          if (++sb == p.SB) { sb = 0; pb ^= 1; }
        }
        if (++sa == p.SA) { sa = 0; pa ^= 1; }
      }
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
  } else {
    // ======================================= EPILOGUE ===========================================
    int as = 0;
    uint32_t pacc = 0;
    const int plane = p.Hp * p.Wp;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
      const int z = tile / p.tiles_per_sample;
      const int q = (tile - z * p.tiles_per_sample) * TM + warp * 32 + lane;
      const bool qv = q < p.Qs;
      const int rem = qv ? q % plane : 0;
      const int hh = rem / p.Wp, ww = rem - hh * p.Wp;
      const bool interior = qv && hh >= p.ph && hh < p.Hp - p.ph && ww >= p.pw && ww < p.Wp - p.pw;
      const size_t orow = ((size_t)z * p.Qs + (qv ? q : 0)) * p.N;
      mbar_wait(smem_u32(&acc_full[as]), pacc);
      tc_fence_after();
      const uint32_t tlane = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(as * p.n_pad);
      for (int c0 = 0; c0 < p.N; c0 += 8) {
        uint32_t v[8];
        tmem_ld8(tlane + (uint32_t)c0, v);
        tmem_ld_wait();
        if (!qv) continue;
        const int nvalid = min(8, p.N - c0);
        float o[8];
#pragma unroll
        for (int jx = 0; jx < 8; ++jx) {
          float a = 0.f;
          if (interior && jx < nvalid) {
            a = __uint_as_float(v[jx]);
            if (p.scale) a = __fmul_rn(a, __ldg(p.scale + c0 + jx));
            if (p.shift) a = __fadd_rn(a, __ldg(p.shift + c0 + jx));
            if (p.residual) a = __fadd_rn(a, __ldg(p.residual + orow + c0 + jx));
            if (p.flags & QBN_FLAG_RELU) a = fmaxf(a, 0.f);
            if (p.flags & QBN_FLAG_OUT_ROUND_TF32) a = tf32_round(a);
          }
          o[jx] = a;
        }
        if (nvalid == 8 && ((orow + c0) & 3) == 0) {
          *reinterpret_cast<float4*>(p.out + orow + c0) = make_float4(o[0], o[1], o[2], o[3]);
          *reinterpret_cast<float4*>(p.out + orow + c0 + 4) = make_float4(o[4], o[5], o[6], o[7]);
        } else {
          for (int jx = 0; jx < nvalid; ++jx) p.out[orow + c0 + jx] = o[jx];
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&acc_empty[as]));               // accumulator free for tile i+2
      if (++as == 2) { as = 0; pacc ^= 1; }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 4) {
    tc_fence_after();
    tmem_dealloc(tmem_base, (uint32_t)p.tmem_cols);
  }
}

}  // namespace

extern "C" int qbn_conv_s1_fwd(int n_samples, int B, int Hp, int Wp, int C, int N, int R, int S, const float* x, const float* w,
                               int w_shared, const float* scale, const float* shift, const float* residual, int flags, float* out,
                               void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  QBN_CHECK_ARG(x && w && out, "null pointer");
  QBN_CHECK_ARG(n_samples > 0 && B > 0 && Hp > 0 && Wp > 0 && C > 0 && N > 0 && R > 0 && S > 0, "sizes");
  QBN_CHECK_ARG(Hp > R - 1 && Wp > S - 1, "padded extent must exceed the halo");
  if (C % 4 != 0 || (R & 1) == 0 || (S & 1) == 0 || N > 256) {
    qbn_set_error("qbn_conv_s1_fwd: needs C %% 4 == 0, odd kernel, N <= 256 (C=%d R=%d S=%d N=%d)", C, R, S, N);
    return QBN_ERR_UNSUPPORTED;
  }
  S1Params p;
  memset(&p, 0, sizeof(p));
  p.Hp = Hp; p.Wp = Wp; p.C = C; p.N = N; p.R = R; p.S = S; p.ph = (R - 1) / 2; p.pw = (S - 1) / 2;
  p.Qs = B * Hp * Wp;
  p.tiles_per_sample = (p.Qs + TM - 1) / TM;
  p.total_tiles = p.tiles_per_sample * n_samples;
  p.n_pad = (N + 15) / 16 * 16;
  p.K = R * S * C;
  // channel blocking: whole C when small, else 32-channel blocks; chunks per block rounded to even
  p.CB = C <= 48 ? C : 32;
  p.n_cb = (C + p.CB - 1) / p.CB;
  p.cbc = ((p.CB + 7) / 8) * 2;
  p.D = p.ph * Wp + p.pw;
  p.RA = TM + 2 * p.D;
  p.RA_p = p.RA | 1;                // odd pitch: conflict-free LDGSTS for the (row, chunk) thread map
  p.b_pitch = p.n_pad + 1;
  p.flags = flags; p.w_shared = w_shared;
  p.x = x; p.w = w; p.scale = scale; p.shift = shift; p.residual = residual; p.out = out;
  p.idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(p.n_pad >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);
  int cols = 2 * p.n_pad;
  p.tmem_cols = 32;
  while (p.tmem_cols < cols) p.tmem_cols <<= 1;
  const size_t a_bytes = (size_t)p.cbc * p.RA_p * 16, b_bytes = (size_t)p.cbc * p.b_pitch * 16;
  p.SA = 2;
  p.SB = 4;
  size_t smem = p.SA * a_bytes + p.SB * b_bytes + (2 * p.SA + 2 * p.SB + 4) * 8 + 16;
  const size_t cap = 220 * 1024;
  if (smem > cap) { p.SB = 2; smem = p.SA * a_bytes + p.SB * b_bytes + (2 * p.SA + 2 * p.SB + 4) * 8 + 16; }
  if (smem > cap) {
    qbn_set_error("qbn_conv_s1_fwd: tile does not fit shared memory (%zu bytes)", smem);
    return QBN_ERR_UNSUPPORTED;
  }
  // deepen the weight ring while two CTAs still fit per SM
  while (p.SB < 8 && smem + b_bytes <= 100 * 1024) { p.SB++; smem += b_bytes + 16; }
  static bool attr_set = false;
  if (!attr_set) {
    QBN_CUDA(cudaFuncSetAttribute(umma_conv_s1_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    attr_set = true;
  }
  int occ = (int)((227 * 1024) / (smem + 1024));
  int occ_t = 512 / p.tmem_cols;
  if (occ > occ_t) occ = occ_t;
  if (occ > 3) occ = 3;
  if (occ < 1) occ = 1;
  int grid = qbn_sm_count() * occ;
  if (grid > p.total_tiles) grid = p.total_tiles;
  umma_conv_s1_kernel<<<grid, NTHREADS_S1, smem, st>>>(p);
  QBN_CHECK_LAUNCH();
  return QBN_OK;
}
